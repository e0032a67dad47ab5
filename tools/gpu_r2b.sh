#!/bin/bash
# round 2, visit B (2 GPUs): multi-GPU tests, driver-style scaling bench N=1,2, e2e comparison with the r1 bench
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_multi.py tests/test_gpu_reset.py -x -q > $out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $out/r2b_pytest.log
python bench.py --steps 20 --warmup 5 --quick --no-cpu > $out/r2b_n1.json 2> $out/r2b_n1.err; echo "n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > $out/r2b_n2.json 2> $out/r2b_n2.err; echo "n2 rc=$?"
tail -c 1500 $out/r2b_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 240 --warmup 24 > $out/r2b_n2_k240.json 2>> $out/r2b_n2.err; echo "n2 k240 rc=$?"
python tools/bench_r1.py --steps 20 --warmup 5 --no-cpu > $out/r2b_r1bench.json 2> $out/r2b_r1bench.err; echo "r1bench rc=$?"
python bench.py --steps 20 --warmup 5 --quick --no-cpu > $out/r2b_n1_again.json 2>> $out/r2b_n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value=%.3e'%d['value'], 'ms/step=%.5f'%d['ms_per_step'], 'e2e=%.3e'%d['e2e']['value'], d['e2e'].get('us_per_step_per_rank'), d.get('gather_parity'), d.get('transport'), d.get('all_envs_alive'))
    except Exception as e: print(f, 'ERR', e)
PY
