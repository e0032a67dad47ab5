#!/bin/bash
# One GPU-box visit: GPU test suite, default bench (both arms), team-size sweep, ncu launch list + full capture.
# usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
set -u
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 40 --warmup 3 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err; echo "ref rc=$?"
CMD="python bench.py --steps 96 --warmup 24 --no-cpu --pool 4 --no-graph"
$CMD > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $CMD > $out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:burgers_warp -s 30 -c 3 -f -o $out/${tag}_prof_burgers $CMD > $out/${tag}_ncu2.log 2>&1
echo done
