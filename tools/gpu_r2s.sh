#!/bin/bash
# visit S (1 GPU): ncu capture of the two-warp DNS kernel + the complete default bench line of the final code
set -u
out=gpurun_out; mkdir -p $out
python tools/dns_run.py 500 1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:dns1024x2 -c 1 -f -o $out/r2s_prof_dns python tools/dns_run.py 60 1 > $out/r2s_ncu.log 2>&1; tail -2 $out/r2s_ncu.log
python bench.py --steps 20 --warmup 5 > $out/r2s_bench_k20.json 2> $out/r2s_bench_k20.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s_bench_k20.json').read().strip().splitlines()[-1])
print('K=%d value=%.3e us/step=%.3f e2e=%.3e (%.1f us, floor %.1f us) alive=%s frac=%.3f fp64=%.3f' % (d['steps'], d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['e2e']['us_per_step'], d['e2e']['pcie_floor_us_per_step'], d['all_envs_alive'], d['roofline']['frac'], d['roofline_fp64']['frac']))
print(json.dumps(d.get('other_configs'))[:2000])
PY
