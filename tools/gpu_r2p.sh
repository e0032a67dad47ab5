#!/bin/bash
# visit P (1 GPU): end-to-end pipeline depth sweep against the PCIe floor of the same box
set -u
out=gpurun_out; mkdir -p $out
python tools/e2e_floor.py | tee $out/r2p_floor.log
for d in 4 8 12 16; do
python bench.py --steps 240 --warmup 5 --quick --no-cpu --depth $d > $out/r2p_d$d.json 2> $out/r2p_d$d.err; python -c "
import json; d=json.loads(open('$out/r2p_d$d.json').read().strip().splitlines()[-1]); print('depth $d e2e %.3e (%.1f us/step)  device %.3f us/step' % (d['e2e']['value'], 40960/d['e2e']['value']*1e6, d['ms_per_step']*1e3))"
done
