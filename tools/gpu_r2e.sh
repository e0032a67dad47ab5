#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
python tools/dns_run.py 500 1 > $out/r2e_dns_hist.log 2>&1; cat $out/r2e_dns_hist.log
python tools/dns_run.py 500 0 > $out/r2e_dns_nohist.log 2>&1; cat $out/r2e_dns_nohist.log
ncu --set full --clock-control none --import-source on -k regex:dns1024 -c 1 -f -o $out/r2e_prof_dns python tools/dns_run.py 60 1 > $out/r2e_ncu.log 2>&1; tail -3 $out/r2e_ncu.log
python bench.py --steps 20 --warmup 5 --quick --no-cpu > $out/r2e_bench_k20.json 2> $out/r2e_bench.err; python -c "
import json; d=json.loads(open('$out/r2e_bench_k20.json').read().strip().splitlines()[-1]); print('k20 us/step', d['ms_per_step']*1e3, d['value'])"
