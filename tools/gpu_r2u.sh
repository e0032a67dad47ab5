#!/bin/bash
# visit U (1 GPU): ncu --set full of the configuration-5 (multi-agent, MSE) step kernel
set -u
out=gpurun_out; mkdir -p $out
CMD="python bench.py --workload c5 --steps 24 --warmup 5 --quick --no-cpu --chains 1 --pool 4"
$CMD > $out/r2u_plain.log 2>&1; echo "plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:burgers_warp -s 40 -c 1 -f -o $out/r2u_prof_c5 $CMD > $out/r2u_ncu.log 2>&1; tail -2 $out/r2u_ncu.log
