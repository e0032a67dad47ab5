#!/bin/bash
# visit N (8 GPUs): scaling with whole-row state stores -- the driver's arguments, gather to the learner (default) and all-gather
set -u
out=gpurun_out; mkdir -p $out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((30100 + n)) bench.py --gpus $n "$@"; }
python bench.py --steps 20 --warmup 5 --quick --no-cpu > $out/r2n_n1.json 2> $out/r2n_n1.err; echo "n1 rc=$?"
run 2 --steps 20 --warmup 5 > $out/r2n_n2.json 2> $out/r2n_n2.err; echo "n2 rc=$?"
run 4 --steps 20 --warmup 5 > $out/r2n_n4.json 2> $out/r2n_n4.err; echo "n4 rc=$?"
run 8 --steps 20 --warmup 5 > $out/r2n_n8.json 2> $out/r2n_n8.err; echo "n8 rc=$?"
run 8 --steps 20 --warmup 5 --gather all > $out/r2n_n8_all.json 2> $out/r2n_n8_all.err; echo "n8 all rc=$?"
run 8 --steps 240 --warmup 24 > $out/r2n_n8_k240.json 2>> $out/r2n_n8.err; echo "n8 k240 rc=$?"
python - <<'PY'
import json
base=None
for f in ['r2n_n1','r2n_n2','r2n_n4','r2n_n8','r2n_n8_all','r2n_n8_k240']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        if base is None: base=d['value']
        print(f, 'N=%d value=%.3e eff=%.3f us/step=%.3f e2e=%.3e parity=%s %s %s' % (d['n_gpus'], d['value'], d['value']/(d['n_gpus']*base), d['ms_per_step']*1e3, d['e2e']['value'], d.get('gather_parity'), d.get('transport'), d.get('gather_bytes_per_step')))
    except Exception as e: print(f, 'ERR', e)
PY
