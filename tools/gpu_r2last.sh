#!/bin/bash
# last validation visit of round 2 (1 GPU): smoke, full suite, both bench arms with the driver's arguments
set -u
out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" > $out/r2last_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/r2last_smoke.log
python -m pytest tests -m gpu -x -q > $out/r2last_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/r2last_pytest.log
python bench.py --impl reference --steps 20 --warmup 5 > $out/r2last_ref.json 2> $out/r2last_ref.err; echo "reference arm rc=$?"
python bench.py --steps 20 --warmup 5 > $out/r2last_bench.json 2> $out/r2last_bench.err; echo "bench rc=$?"; tail -c 200 $out/r2last_bench.err
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r2last_ref.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/r2last_bench.json').read().strip().splitlines()[-1])
print('reference arm: value=%.3e e2e=%.3e cores=%d kind=%s' % (r['value'], r['e2e']['value'], r['cpu_baseline']['cores'], r['cpu_baseline']['kind']))
print('K=%d value=%.3e us/step=%.3f e2e=%.3e (%d steps, %.1f us, floor %.1f us) alive=%s frac=%.3f fp64=%.3f cpu=%.3e launches=%d clocks=%s' % (d['steps'], d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['e2e']['steps'], d['e2e']['us_per_step'], d['e2e']['pcie_floor_us_per_step'], d['all_envs_alive'], d['roofline']['frac'], d['roofline_fp64']['frac'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks']))
print({k: (round(v['launch_us'],2), '%.3e' % v['value']) for k, v in d['other_configs'].items()})
PY
