#!/bin/bash
# visit J (2 GPUs): KS transposed transform (parity + speed), gather-to-learner 2-GPU test and bench
set -u
out=gpurun_out; mkdir -p $out
MPDE_KS_TS=-8 python -m pytest tests/test_gpu_ks.py tests/test_gpu_long.py tests/test_gpu_fp32.py tests/test_gpu_environment.py -x -q -k "ks or KS" > $out/r2j_pytest_ks.log 2>&1; echo "ks(-8) pytest rc=$?"; tail -5 $out/r2j_pytest_ks.log
for ts in 16 -8; do for mb in 4 5; do
MPDE_KS_TS=$ts MPDE_KS_MINB=$mb python - <<'PY'
import os, sys, torch, json
sys.path.insert(0, '.')
import bench
r = bench.other_configs(torch, torch.device('cuda', 0), only='c3')
print('KS_TS', os.environ['MPDE_KS_TS'], 'MINB', os.environ['MPDE_KS_MINB'], json.dumps(r))
PY
done; done
python -m pytest tests/test_gpu_multi.py -x -q > $out/r2j_pytest_multi.log 2>&1; echo "multi pytest rc=$?"; tail -4 $out/r2j_pytest_multi.log
for g in all learner; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus 2 --steps 20 --warmup 5 --gather $g > $out/r2j_n2_$g.json 2> $out/r2j_n2_$g.err; echo "n2 $g rc=$?"
python -c "
import json; d=json.loads(open('$out/r2j_n2_$g.json').read().strip().splitlines()[-1]); print('$g', 'us/step %.3f value %.3e parity %s %s' % (d['ms_per_step']*1e3, d['value'], d['gather_parity'], d['transport']), d.get('gather_bytes_per_step'))"
done
