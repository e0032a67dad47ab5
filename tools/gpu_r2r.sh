#!/bin/bash
# visit R (1 GPU): two-warp DNS kernel -- parity (DNS goldens, sgs, environment episodes use N = 1024 / 512 DNS) and speed vs the one-warp kernel
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_dns.py tests/test_gpu_sgs.py tests/test_gpu_fullsize.py -x -q > $out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r2r_pytest.log
echo "== two warps per environment (default)"; python tools/dns_run.py 500 1 | tail -1; python tools/dns_run.py 500 0 | tail -1
echo "== one warp per environment"; MPDE_DNS_WARPS=1 python tools/dns_run.py 500 1 | tail -1; MPDE_DNS_WARPS=1 python tools/dns_run.py 500 0 | tail -1
