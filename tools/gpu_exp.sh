#!/bin/bash
# experiment visit: bench headline with experimental builds of the library (exp_libs/*.so, MPDE_LIB_PATH)
set -u
out=gpurun_out; mkdir -p $out
for lib in "" $(ls exp_libs/*.so); do
  name=$(basename "${lib:-default}" .so)
  for k in 240; do
    MPDE_LIB_PATH=${lib:+$PWD/$lib} python bench.py --steps $k --warmup 24 --quick --no-cpu > $out/exp_${name}_k$k.json 2> $out/exp_${name}_k$k.err
    python -c "
import json; d=json.loads(open('$out/exp_${name}_k$k.json').read().strip().splitlines()[-1]); print('$name K=$k us/step %.3f value %.3e alive %s' % (d['ms_per_step']*1e3, d['value'], d['all_envs_alive']))" || tail -3 $out/exp_${name}_k$k.err
  done
done
