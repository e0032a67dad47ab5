#!/bin/bash
# visit H (1 GPU): full suite + register-budget experiment for the 4-lane kernel (5 / 6 CTAs per SM)
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/r2h_pytest.log
for v in base lb5 lb6; do
  if [ $v = base ]; then unset MPDE_LIB_PATH; else export MPDE_LIB_PATH=$PWD/marlpde_b200/libmarlpde_b200_$v.so; fi
  for k in 20 240; do
    python bench.py --steps $k --warmup 5 --quick --no-cpu > $out/r2h_${v}_k$k.json 2> $out/r2h_${v}_k$k.err
    python -c "
import json; d=json.loads(open('$out/r2h_${v}_k$k.json').read().strip().splitlines()[-1]); print('$v K=$k us/step %.3f value %.3e one-at-a-time %s' % (d['ms_per_step']*1e3, d['value'], d['roofline']['launch_us_one_batch_at_a_time']))"
  done
done
