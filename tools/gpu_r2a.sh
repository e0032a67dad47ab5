#!/bin/bash
# round 2, visit A: GPU tests + driver-style bench (both arms) on 1 GPU
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/r2a_pytest.log
tail -5 $out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > $out/r2a_bench_k20.json 2> $out/r2a_bench_k20.err; echo "bench k20 rc=$?"
tail -c 600 $out/r2a_bench_k20.err
python bench.py --steps 20 --warmup 5 --quick --no-cpu --chains 1 > $out/r2a_bench_k20_c1.json 2> $out/r2a_c1.err; echo "bench c1 rc=$?"
python bench.py --quick --no-cpu > $out/r2a_bench_default.json 2> $out/r2a_def.err; echo "bench default rc=$?"
python bench.py --steps 20 --warmup 5 --quick --no-cpu --fused-single > $out/r2a_bench_fs.json 2> $out/r2a_fs.err; echo "bench fused-single rc=$?"
tail -c 400 $out/r2a_fs.err
python bench.py --impl reference --steps 20 --warmup 5 > $out/r2a_ref.json 2> $out/r2a_ref.err; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_bench*.json'))+['gpurun_out/r2a_ref.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value=%.3e'%d['value'], 'ms/step=%.5f'%d['ms_per_step'], 'e2e=%.3e'%d['e2e']['value'], d.get('roofline',{}).get('launch_us_one_batch_at_a_time'), d.get('all_envs_alive'))
        for k in ('sweep','other_configs'):
            if k in d: print(json.dumps(d[k])[:3000])
    except Exception as e: print(f, 'ERR', e)
PY
