#!/bin/bash
# visit Z (1 GPU): full suite + headline / c5 / KS timing after the epilogue changes
set -u
out=gpurun_out; mkdir -p $out
tag=${1:-z}
python -m pytest tests -m gpu -x -q > $out/r2${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2${tag}_pytest.log
for k in 20 240; do
  w=$([ $k = 20 ] && echo 5 || echo 24)
  python bench.py --steps $k --warmup $w --quick --no-cpu > $out/r2${tag}_k$k.json 2> $out/r2${tag}_k$k.err
  python -c "
import json; d=json.loads(open('$out/r2${tag}_k$k.json').read().strip().splitlines()[-1]); print('K=$k chains %d us/step %.3f value %.3e alive %s' % (d['timing']['batches_in_flight'], d['ms_per_step']*1e3, d['value'], d['all_envs_alive']))"
done
python - <<'PY'
import json, torch, bench
dev = torch.device('cuda:0')
r = bench.other_configs(torch, dev)
for k, v in r.items():
    print(k, 'launch_us %.2f value %.3e fp64 %.3f alive %s' % (v['launch_us'], v['value'], v['fp64_frac'], v['alive']))
PY
