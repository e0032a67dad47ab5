#!/bin/bash
# visit M (1 GPU): full suite after the raw split step / whole-row state stores + bench
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/r2m_pytest.log
for k in 20 240; do
python bench.py --steps $k --warmup 5 --quick --no-cpu > $out/r2m_k$k.json 2> $out/r2m_k$k.err; python -c "
import json; d=json.loads(open('$out/r2m_k$k.json').read().strip().splitlines()[-1]); print('K=$k us/step %.3f value %.3e chains %s alive %s' % (d['ms_per_step']*1e3, d['value'], d['timing']['batches_in_flight'], d['all_envs_alive']))"
done
python bench.py --steps 20 --warmup 5 --quick --no-cpu --fused-single --gather all > $out/r2m_fs.json 2> $out/r2m_fs.err; python -c "
import json; d=json.loads(open('$out/r2m_fs.json').read().strip().splitlines()[-1]); print('fused-single K=20 us/step %.3f' % (d['ms_per_step']*1e3))"
