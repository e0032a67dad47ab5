#!/bin/bash
# visit (2 GPUs): N = 1 / 2 with the driver's arguments, full lines (final bench.py)
set -u
out=gpurun_out; mkdir -p $out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((30100 + n)) bench.py --gpus $n "$@"; }
python bench.py --steps 20 --warmup 5 --no-cpu > $out/r2m3_n1.json 2> $out/r2m3_n1.err; echo "n1 rc=$?"
run 2 --steps 20 --warmup 5 > $out/r2m3_n2.json 2> $out/r2m3_n2.err; echo "n2 rc=$?"; tail -3 $out/r2m3_n2.err
run 2 --impl reference --steps 20 --warmup 5 > $out/r2m3_ref2.json 2> $out/r2m3_ref2.err; echo "ref n2 rc=$?"; tail -c 200 $out/r2m3_ref2.json
python - <<'PY'
import json
base=None
for f in ['r2m3_n1','r2m3_n2']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        if base is None: base=d['value']
        print(f, 'N=%d value=%.3e eff=%.3f us/step=%.3f e2e=%.3e (%d steps, %s us/rank) parity=%s %s' % (d['n_gpus'], d['value'], d['value']/(d['n_gpus']*base), d['ms_per_step']*1e3, d['e2e']['value'], d['e2e']['steps'], d['e2e']['us_per_step_per_rank'], d.get('gather_parity'), d.get('transport')))
    except Exception as e: print(f, 'ERR', e)
PY
