#!/bin/bash
# round 2, visit G (8 GPUs): the driver's scaling run N = 2, 4, 8, 1 with --steps 20 --warmup 5 (+ the c5 configuration on 8 GPUs)
set -u
out=gpurun_out; mkdir -p $out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n "$@"; }
run 2 --steps 20 --warmup 5 > $out/r2g_n2.json 2> $out/r2g_n2.err; rc=$?; echo "n2 rc=$rc"
if [ $rc -ne 0 ]; then tail -c 2000 $out/r2g_n2.err; exit 1; fi
run 4 --steps 20 --warmup 5 > $out/r2g_n4.json 2> $out/r2g_n4.err; echo "n4 rc=$?"
run 8 --steps 20 --warmup 5 > $out/r2g_n8.json 2> $out/r2g_n8.err; echo "n8 rc=$?"
python bench.py --steps 20 --warmup 5 --quick --no-cpu > $out/r2g_n1.json 2> $out/r2g_n1.err; echo "n1 rc=$?"
run 8 --steps 240 --warmup 24 > $out/r2g_n8_k240.json 2>> $out/r2g_n8.err; echo "n8 k240 rc=$?"
run 8 --steps 240 --warmup 24 --workload c5 --rewards-only > $out/r2g_c5_n8.json 2> $out/r2g_c5.err; echo "c5 n8 rc=$?"
tail -c 600 $out/r2g_c5.err
python -m pytest tests/test_gpu_multi.py -x -q > $out/r2g_pytest.log 2>&1; tail -3 $out/r2g_pytest.log
python - <<'PY'
import json,glob
base=None
for f in ['gpurun_out/r2g_n1.json','gpurun_out/r2g_n2.json','gpurun_out/r2g_n4.json','gpurun_out/r2g_n8.json','gpurun_out/r2g_n8_k240.json','gpurun_out/r2g_c5_n8.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if base is None: base=d['value']
        print(f, 'N=%d value=%.3e eff=%.3f us/step=%.3f e2e=%.3e per-rank=%s parity=%s %s alive=%s' % (d['n_gpus'], d['value'], d['value']/(d['n_gpus']*base), d['ms_per_step']*1e3, d['e2e']['value'], d['e2e'].get('us_per_step_per_rank'), d.get('gather_parity'), d.get('transport'), d.get('all_envs_alive')))
    except Exception as e: print(f, 'ERR', e)
PY
