#!/bin/bash
# visit T (1 GPU): multi-agent (MSE) specialisation of the step kernel: parity tests, configuration 5 before/after, full suite
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_burger.py tests/test_gpu_fullsize.py -m gpu -x -q > $out/r2t_pytest_a.log 2>&1; echo "pytest (burger, fullsize) rc=$?"; tail -4 $out/r2t_pytest_a.log
python bench.py --workload c5 --quick --no-cpu --steps 240 --warmup 24 > $out/r2t_c5.json 2> $out/r2t_c5.err; echo "c5 rc=$?"; tail -c 300 $out/r2t_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_c5.json').read().strip().splitlines()[-1])
print('c5: K=%d value=%.3e us/step=%.3f alive=%s frac=%.3f fp64=%.3f' % (d['steps'], d['value'], d['ms_per_step']*1e3, d['all_envs_alive'], d['roofline']['frac'], d['roofline_fp64']['frac']))
PY
python -m pytest tests -m gpu -x -q > $out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r2t_pytest.log
python bench.py --no-cpu > $out/r2t_bench.json 2> $out/r2t_bench.err; echo "bench rc=$?"; tail -c 300 $out/r2t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print('K=%d value=%.3e us/step=%.3f e2e=%.3e alive=%s frac=%.3f fp64=%.3f' % (d['steps'], d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['all_envs_alive'], d['roofline']['frac'], d['roofline_fp64']['frac']))
print(json.dumps(d.get('other_configs'))[:2000])
PY
