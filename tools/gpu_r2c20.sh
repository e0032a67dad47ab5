#!/bin/bash
# visit (1 GPU): batches in flight at the driver's K = 20 with the final kernel
set -u
out=gpurun_out; mkdir -p $out
for c in 4 5 7 10 20; do
  python bench.py --steps 20 --warmup 5 --quick --no-cpu --chains $c > $out/r2c20_c$c.json 2> $out/r2c20_c$c.err
  python -c "
import json; d=json.loads(open('$out/r2c20_c$c.json').read().strip().splitlines()[-1]); print('K=20 chains=$c us/step %.3f value %.3e' % (d['ms_per_step']*1e3, d['value']))"
done
for c in 3 4 6 8; do
  python bench.py --steps 240 --warmup 24 --quick --no-cpu --chains $c > $out/r2c240_c$c.json 2> $out/r2c240_c$c.err
  python -c "
import json; d=json.loads(open('$out/r2c240_c$c.json').read().strip().splitlines()[-1]); print('K=240 chains=$c us/step %.3f value %.3e' % (d['ms_per_step']*1e3, d['value']))"
done
