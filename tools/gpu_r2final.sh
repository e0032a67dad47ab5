#!/bin/bash
# final visit of round 2 (1 GPU): smoke, full suite, both bench arms with the driver's arguments, launch list + ncu of the headline kernel
set -u
out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" > $out/r2fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/r2fin_smoke.log
python -m pytest tests -m gpu -x -q > $out/r2fin_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2fin_pytest.log
python bench.py --impl reference --steps 20 --warmup 5 > $out/r2fin_ref.json 2> $out/r2fin_ref.err; echo "reference arm rc=$?"; tail -c 600 $out/r2fin_ref.json
python bench.py --steps 20 --warmup 5 > $out/r2fin_bench.json 2> $out/r2fin_bench.err; echo "bench rc=$?"; tail -c 300 $out/r2fin_bench.err
python bench.py > $out/r2fin_bench_default.json 2> $out/r2fin_bench_default.err; echo "bench (defaults) rc=$?"
python - <<'PY'
import json
for f in ('r2fin_bench', 'r2fin_bench_default'):
    d=json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
    print(f, 'K=%d value=%.3e us/step=%.3f e2e=%.3e (%.1f us, floor %.1f us) alive=%s frac=%.3f fp64=%.3f cpu=%.3e' % (d['steps'], d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['e2e']['us_per_step'], d['e2e']['pcie_floor_us_per_step'], d['all_envs_alive'], d['roofline']['frac'], d['roofline_fp64']['frac'], d['cpu_baseline']['value']))
    print(json.dumps(d.get('other_configs'))[:1500])
    print(json.dumps(d.get('sweep'))[:2500])
PY
CMD="python bench.py --steps 48 --warmup 5 --quick --no-cpu --chains 1 --pool 4"
$CMD > $out/r2fin_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r2fin_launches.csv $CMD > $out/r2fin_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:burgers_warp -s 60 -c 1 -f -o $out/r2fin_prof_burgers $CMD > $out/r2fin_ncu2.log 2>&1; tail -2 $out/r2fin_ncu2.log
ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:burgers_warp -s 4 -c 1 -f -o $out/r2fin_prof_b32768 python tools/step_run.py 32768 10 3 2 > $out/r2fin_ncu3.log 2>&1; tail -1 $out/r2fin_ncu3.log
