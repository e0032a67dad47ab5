#!/bin/bash
# visit I (1 GPU): chains sweep at K = 20, ncu launch list + full captures of the headline and DNS kernels
set -u
out=gpurun_out; mkdir -p $out
for c in 4 5 10 20; do
  python bench.py --steps 20 --warmup 5 --quick --no-cpu --chains $c > $out/r2i_k20_c$c.json 2> $out/r2i_k20_c$c.err
  python -c "
import json; d=json.loads(open('$out/r2i_k20_c$c.json').read().strip().splitlines()[-1]); print('K=20 chains=$c us/step %.3f value %.3e' % (d['ms_per_step']*1e3, d['value']))"
done
CMD="python bench.py --steps 48 --warmup 5 --quick --no-cpu --chains 1 --pool 4"
$CMD > $out/r2i_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r2i_launches.csv $CMD > $out/r2i_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:burgers_warp -s 60 -c 2 -f -o $out/r2i_prof_burgers $CMD > $out/r2i_ncu2.log 2>&1; tail -2 $out/r2i_ncu2.log
ncu --set full --clock-control none --import-source on -k regex:dns1024 -c 1 -f -o $out/r2i_prof_dns python tools/dns_run.py 60 1 > $out/r2i_ncu3.log 2>&1; tail -2 $out/r2i_ncu3.log
