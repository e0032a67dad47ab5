#!/bin/bash
# visit K (8 GPUs): gather-to-learner vs all-gather at N = 8 / 4 with the driver's arguments, K = 240, and configuration 5
set -u
out=gpurun_out; mkdir -p $out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29900 + n)) bench.py --gpus $n "$@"; }
run 8 --steps 20 --warmup 5 --gather learner > $out/r2k_n8_learner.json 2> $out/r2k_n8_learner.err; echo "n8 learner rc=$?"
run 8 --steps 20 --warmup 5 --gather all > $out/r2k_n8_all.json 2> $out/r2k_n8_all.err; echo "n8 all rc=$?"
run 4 --steps 20 --warmup 5 --gather learner > $out/r2k_n4_learner.json 2> $out/r2k_n4_learner.err; echo "n4 learner rc=$?"
run 2 --steps 20 --warmup 5 --gather learner > $out/r2k_n2_learner.json 2> $out/r2k_n2_learner.err; echo "n2 learner rc=$?"
python bench.py --steps 20 --warmup 5 --quick --no-cpu > $out/r2k_n1.json 2> $out/r2k_n1.err; echo "n1 rc=$?"
run 8 --steps 240 --warmup 24 --gather learner > $out/r2k_n8_learner_k240.json 2>> $out/r2k_n8_learner.err; echo "n8 learner k240 rc=$?"
run 8 --steps 240 --warmup 24 --workload c5 --rewards-only --gather all > $out/r2k_c5_n8.json 2> $out/r2k_c5.err; echo "c5 n8 rc=$?"
tail -c 400 $out/r2k_c5.err
python - <<'PY'
import json
base=None
for f in ['r2k_n1','r2k_n2_learner','r2k_n4_learner','r2k_n8_learner','r2k_n8_all','r2k_n8_learner_k240','r2k_c5_n8']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        if base is None: base=d['value']
        print(f, 'N=%d value=%.3e eff=%.3f us/step=%.3f e2e=%.3e per-rank=%s parity=%s %s alive=%s %s' % (d['n_gpus'], d['value'], d['value']/(d['n_gpus']*base), d['ms_per_step']*1e3, d['e2e']['value'], d['e2e'].get('us_per_step_per_rank'), d.get('gather_parity'), d.get('transport'), d.get('all_envs_alive'), d.get('gather_bytes_per_step')))
    except Exception as e: print(f, 'ERR', e)
PY
