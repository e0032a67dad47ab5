#!/bin/bash
# visit (2 GPUs): multi-GPU tests + N = 1 / 2 with the driver's arguments after the epilogue changes
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $out/r2m2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2m2_pytest.log
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((30100 + n)) bench.py --gpus $n "$@"; }
python bench.py --steps 20 --warmup 5 --quick --no-cpu > $out/r2m2_n1.json 2> $out/r2m2_n1.err; echo "n1 rc=$?"
run 2 --steps 20 --warmup 5 --quick --no-cpu > $out/r2m2_n2.json 2> $out/r2m2_n2.err; echo "n2 rc=$?"
run 2 --steps 20 --warmup 5 --quick --no-cpu --gather all > $out/r2m2_n2_all.json 2> $out/r2m2_n2_all.err; echo "n2 all rc=$?"
python - <<'PY'
import json
base=None
for f in ['r2m2_n1','r2m2_n2','r2m2_n2_all']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        if base is None: base=d['value']
        print(f, 'N=%d value=%.3e eff=%.3f us/step=%.3f e2e=%.3e parity=%s %s' % (d['n_gpus'], d['value'], d['value']/(d['n_gpus']*base), d['ms_per_step']*1e3, d['e2e']['value'], d.get('gather_parity'), d.get('transport')))
    except Exception as e: print(f, 'ERR', e)
PY
