#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + source page) into the handful of numbers we track.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_profiled_launches]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
nl = len(rows) - 2
vals = rows[2]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
print(f"kernel: {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}  ({nl} launches in report)")
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:75s} {units[i]:10s} {vals[i]}")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        v = float(vals[i])
        if v >= 0.15:
            print(f"stall {h.split('stalled_')[1].split('_per_issue')[0]:22s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ia, isrc = h.index("Instructions Executed"), h.index("Source")
ops = collections.Counter()
for r in rows[2:]:
    try:
        n = int(r[ia])
    except Exception:
        continue
    s = re.sub(r'^@!?U?P\w+\s+', '', r[isrc].strip())
    ops[s.split()[0].split('.')[0] if s else '?'] += n
grid = int(vals[hdr.index("launch__grid_size")]); blk = int(vals[hdr.index("launch__block_size")])
warps = grid * blk // 32
tot = sum(ops.values())
div = float(vals[hdr.index("smsp__inst_executed.sum")])
scale = div / tot if tot else 1
print(f"warp-instructions per warp per launch: {div / warps:.0f}")
print("opcode mix per warp per launch: " + ", ".join(f"{o} {n * scale / warps:.0f}" for o, n in ops.most_common(16)))
