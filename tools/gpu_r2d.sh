#!/bin/bash
# round 2, visit D: full GPU suite (new defaults, DNS warp kernel) + full bench line
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $out/r2d_pytest.log
python bench.py --steps 20 --warmup 5 > $out/r2d_bench_k20.json 2> $out/r2d_bench_k20.err; echo "bench rc=$?"
tail -c 800 $out/r2d_bench_k20.err
MPDE_DNS_WARP=0 python - > $out/r2d_dns_old.log 2>&1 <<'PY'
import sys, torch, json
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.other_configs(torch, torch.device('cuda', 0))['c4_dns_n1024_x512']))
PY
cat $out/r2d_dns_old.log | tail -2
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench_k20.json').read().strip().splitlines()[-1])
print('value=%.3e us/step=%.3f e2e=%.3e alive=%s' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['all_envs_alive']))
print(json.dumps(d.get('roofline'))[:600])
print(json.dumps(d.get('sweep'))[:2500])
print(json.dumps(d.get('other_configs'))[:2500])
print(d.get('cpu_baseline'))
PY
