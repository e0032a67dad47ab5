#!/bin/bash
# visit O (1 GPU): full suite (device spline fit on the MSE-reward path) + the complete default bench line
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/r2o_pytest.log
python bench.py --steps 20 --warmup 5 > $out/r2o_bench_k20.json 2> $out/r2o_bench_k20.err; echo "bench rc=$?"; tail -c 600 $out/r2o_bench_k20.err
python bench.py --impl reference --steps 20 --warmup 5 > $out/r2o_ref.json 2> $out/r2o_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench_k20.json').read().strip().splitlines()[-1])
print('value=%.3e us/step=%.3f e2e=%.3e alive=%s frac=%.3f fp64=%.3f' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['all_envs_alive'], d['roofline']['frac'], d['roofline_fp64']['frac']))
print(json.dumps(d.get('sweep'))[:3000])
print(json.dumps(d.get('other_configs'))[:2000])
print(d.get('cpu_baseline'))
r=json.loads(open('gpurun_out/r2o_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])
PY
