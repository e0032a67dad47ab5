#!/bin/bash
# visit X (1 GPU): whole-warp collectives in the training kernels: parity + headline timing
set -u
out=gpurun_out; mkdir -p $out
tag=${1:-x}
python -m pytest tests/test_gpu_burger.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py tests/test_gpu_environment.py -m gpu -x -q > $out/r2${tag}_pytest_a.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2${tag}_pytest_a.log
for k in 20 240; do
  w=$([ $k = 20 ] && echo 5 || echo 24)
  python bench.py --steps $k --warmup $w --quick --no-cpu > $out/r2${tag}_k$k.json 2> $out/r2${tag}_k$k.err
  python -c "
import json; d=json.loads(open('$out/r2${tag}_k$k.json').read().strip().splitlines()[-1]); print('K=$k chains %d us/step %.3f value %.3e alive %s' % (d['timing']['batches_in_flight'], d['ms_per_step']*1e3, d['value'], d['all_envs_alive']))"
done
python bench.py --workload c5 --quick --no-cpu --steps 240 --warmup 24 > $out/r2${tag}_c5.json 2> $out/r2${tag}_c5.err
python -c "
import json; d=json.loads(open('$out/r2${tag}_c5.json').read().strip().splitlines()[-1]); print('c5 us/step %.3f value %.3e alive %s' % (d['ms_per_step']*1e3, d['value'], d['all_envs_alive']))"
python tools/step_run.py 32768 1 | tail -1
python tools/step_run.py 32768 10 | tail -1
