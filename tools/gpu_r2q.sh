#!/bin/bash
# visit Q (1 GPU): smoke(), full suite, full default bench line (after the gather / KS trims)
set -u
out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" > $out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/r2q_smoke.log
python -m pytest tests -m gpu -x -q > $out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/r2q_pytest.log
python bench.py --no-cpu > $out/r2q_bench.json 2> $out/r2q_bench.err; echo "bench rc=$?"; tail -c 400 $out/r2q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
print('K=%d value=%.3e us/step=%.3f e2e=%.3e (%.1f us, floor %.1f us) alive=%s frac=%.3f fp64=%.3f' % (d['steps'], d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['e2e']['us_per_step'], d['e2e']['pcie_floor_us_per_step'], d['all_envs_alive'], d['roofline']['frac'], d['roofline_fp64']['frac']))
print(json.dumps(d.get('other_configs'))[:2000])
PY
