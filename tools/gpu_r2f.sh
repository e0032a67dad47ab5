#!/bin/bash
# round 2, visit F: full GPU suite (new FD envs, sgs, Burger_fd cases, DNS tweaks) + bench with the per-env-seed workload
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 $out/r2f_pytest.log
python tools/dns_run.py 500 1 > $out/r2f_dns_hist.log 2>&1; tail -1 $out/r2f_dns_hist.log
python tools/dns_run.py 500 0 > $out/r2f_dns_nohist.log 2>&1; tail -1 $out/r2f_dns_nohist.log
python bench.py --steps 20 --warmup 5 --no-cpu > $out/r2f_bench_k20.json 2> $out/r2f_bench_k20.err; echo "bench rc=$?"
tail -c 1000 $out/r2f_bench_k20.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_k20.json').read().strip().splitlines()[-1])
print('value=%.3e us/step=%.3f e2e=%.3e alive=%s %s ref=%s' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['all_envs_alive'], d['alive_fraction'], d['reward_reference']))
print(json.dumps(d.get('other_configs'))[:1500])
PY
