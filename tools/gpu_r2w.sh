#!/bin/bash
# visit W (1 GPU): where the per-launch fixed cost goes -- ncu source-level capture of a loaded chip (B = 32768) at 1 and 10 fused steps
set -u
out=gpurun_out; mkdir -p $out


ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:burgers_warp -s 4 -c 1 -f -o $out/r2w_prof_n1 python tools/step_run.py 32768 1 3 2 > $out/r2w_ncu1.log 2>&1; tail -1 $out/r2w_ncu1.log
ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:burgers_warp -s 4 -c 1 -f -o $out/r2w_prof_n10 python tools/step_run.py 32768 10 3 2 > $out/r2w_ncu10.log 2>&1; tail -1 $out/r2w_ncu10.log
