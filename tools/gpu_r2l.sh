#!/bin/bash
# visit L (1 GPU): DNS kernel with the spectrum in shared memory vs in registers; DNS parity tests on the new default
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_dns.py tests/test_gpu_sgs.py tests/test_gpu_burger.py -x -q > $out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2l_pytest.log
echo "== v in shared memory (default)"; python tools/dns_run.py 500 1 | tail -1; python tools/dns_run.py 500 0 | tail -1
echo "== v in registers"; MPDE_DNS_VREG=1 python tools/dns_run.py 500 1 | tail -1; MPDE_DNS_VREG=1 python tools/dns_run.py 500 0 | tail -1
python bench.py --steps 20 --warmup 5 --quick --no-cpu > $out/r2l_k20.json 2> $out/r2l_k20.err; python -c "
import json; d=json.loads(open('$out/r2l_k20.json').read().strip().splitlines()[-1]); print('K=20 us/step %.3f value %.3e chains %s' % (d['ms_per_step']*1e3, d['value'], d['timing']['batches_in_flight']))"
python bench.py --quick --no-cpu > $out/r2l_k240.json 2> $out/r2l_k240.err; python -c "
import json; d=json.loads(open('$out/r2l_k240.json').read().strip().splitlines()[-1]); print('K=240 us/step %.3f value %.3e chains %s' % (d['ms_per_step']*1e3, d['value'], d['timing']['batches_in_flight']))"
