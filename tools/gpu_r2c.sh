#!/bin/bash
# round 2, visit C: new parity tests + lanes x chains experiment (1 GPU)
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_long.py tests/test_gpu_burger.py -x -q > $out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $out/r2c_pytest.log
for lanes in 8 4 16; do for chains in 1 2 3 4 6; do
  python bench.py --steps 240 --warmup 24 --quick --no-cpu --lanes $lanes --chains $chains > $out/r2c_l${lanes}_c${chains}.json 2> $out/r2c_l${lanes}_c${chains}.err
  python - <<PY
import json
try:
    d=json.loads(open('$out/r2c_l${lanes}_c${chains}.json').read().strip().splitlines()[-1])
    print('lanes=$lanes chains=$chains us/step=%.3f value=%.3e alive=%s' % (d['ms_per_step']*1e3, d['value'], d['all_envs_alive']))
except Exception as e: print('lanes=$lanes chains=$chains ERR', e)
PY
done; done
