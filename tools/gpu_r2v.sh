#!/bin/bash
# visit V (1 GPU): programmatic dependent launch on / off with several batches in flight
set -u
out=gpurun_out; mkdir -p $out
for pdl in 1 0; do for k in 20 240; do
  w=$([ $k = 20 ] && echo 5 || echo 24)
  MPDE_PDL=$pdl python bench.py --steps $k --warmup $w --quick --no-cpu > $out/r2v_pdl${pdl}_k$k.json 2> $out/r2v_pdl${pdl}_k$k.err
  python -c "
import json; d=json.loads(open('$out/r2v_pdl${pdl}_k$k.json').read().strip().splitlines()[-1]); print('PDL=$pdl K=$k chains %d us/step %.3f value %.3e' % (d['timing']['batches_in_flight'], d['ms_per_step']*1e3, d['value']))"
done; done
for c in 1 2; do
  MPDE_PDL=0 python bench.py --steps 240 --warmup 24 --quick --no-cpu --chains $c > $out/r2v_pdl0_c$c.json 2> $out/r2v_pdl0_c$c.err
  python -c "
import json; d=json.loads(open('$out/r2v_pdl0_c$c.json').read().strip().splitlines()[-1]); print('PDL=0 K=240 chains $c us/step %.3f' % (d['ms_per_step']*1e3))"
  MPDE_PDL=1 python bench.py --steps 240 --warmup 24 --quick --no-cpu --chains $c > $out/r2v_pdl1_c$c.json 2> $out/r2v_pdl1_c$c.err
  python -c "
import json; d=json.loads(open('$out/r2v_pdl1_c$c.json').read().strip().splitlines()[-1]); print('PDL=1 K=240 chains $c us/step %.3f' % (d['ms_per_step']*1e3))"
done
