#!/bin/bash
# visit (2 GPUs): staged whole-row stores vs direct per-lane stores in the fused gather
set -u
out=gpurun_out; mkdir -p $out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((30100 + n)) bench.py --gpus $n "$@"; }
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "fused" > $out/r2rows_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/r2rows_pytest.log
for rows in 1 0; do for g in learner all; do
  MPDE_PEER_ROW_STORES=$rows run 2 --steps 20 --warmup 5 --quick --no-cpu --gather $g > $out/r2rows_${rows}_$g.json 2> $out/r2rows_${rows}_$g.err
  python -c "
import json; d=json.loads(open('$out/r2rows_${rows}_$g.json').read().strip().splitlines()[-1]); print('row_stores=$rows gather=$g us/step %.3f value %.3e parity %s' % (d['ms_per_step']*1e3, d['value'], d.get('gather_parity')))"
done; done
MPDE_PEER_ROW_STORES=0 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "fused" > $out/r2rows_pytest0.log 2>&1; echo "pytest (direct) rc=$?"; tail -2 $out/r2rows_pytest0.log
