#!/bin/bash
# visit (2 GPUs): the two state-row store paths of the fused gather (parity) + N = 2 with the driver's arguments
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "fused-rows or fused-learner-rows or fused-learner or test_two_gpu_marl" > $out/r2m4_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/r2m4_pytest.log
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((30100 + n)) bench.py --gpus $n "$@"; }
run 2 --steps 20 --warmup 5 --quick --no-cpu > $out/r2m4_n2.json 2> $out/r2m4_n2.err; echo "n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m4_n2.json').read().strip().splitlines()[-1])
print('N=%d value=%.3e us/step=%.3f e2e=%.3e parity=%s %s' % (d['n_gpus'], d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d.get('gather_parity'), d.get('transport')))
PY
