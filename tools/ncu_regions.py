#!/usr/bin/env python3
"""Split the SASS of a profiled kernel into prologue / hot loop / epilogue and report, per region, the warp-instructions
executed and the stall samples (source page of an .ncu-rep captured with --import-source on).

The hot loop is the address range whose instructions ran most often (>= `frac` x the maximum executed count, default 0.5);
everything before its first instruction is the prologue, everything after its last the epilogue.
usage: tools/ncu_regions.py gpurun_out/prof.ncu-rep [frac] [top_n_epilogue_instructions]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
print(rows[0][1] if len(rows[0]) > 1 else "")
h = rows[1]
ia, isrc, ismp, iad = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples"), h.index("Address")
ins, last = [], -1
for r in rows[2:]:
    try:
        ad = int(r[iad], 16)
        if ad < last:
            break                                  # a second profiled launch follows: the first one is enough
        last = ad
        ins.append((int(r[ia]), int(r[ismp]), r[isrc].strip()))
    except Exception:
        if ins:
            break
mx = max(n for n, _, _ in ins)
hot = [i for i, (n, _, _) in enumerate(ins) if n >= frac * mx]
lo, hi = hot[0], hot[-1]
tot_i, tot_s = sum(n for n, _, _ in ins), sum(s for _, s, _ in ins)
warps = min(n for n, _, _ in ins[:4]) or 1          # the first instructions run once per warp


def opname(s):
    s = re.sub(r'^@!?U?P\w+\s+', '', s)
    return s.split()[0].split('.')[0] if s else '?'


for name, a, b in (("prologue", 0, lo), ("loop", lo, hi + 1), ("epilogue", hi + 1, len(ins))):
    seg = ins[a:b]
    ni, ns = sum(n for n, _, _ in seg), sum(s for _, s, _ in seg)
    ops = collections.Counter()
    for n, _, s in seg:
        ops[opname(s)] += n
    print(f"{name:9s} SASS lines {len(seg):5d}  warp-instr/warp {ni / warps:8.0f} ({100 * ni / tot_i:4.1f} %)  stall samples {100 * ns / max(tot_s, 1):4.1f} %  :: "
          + ", ".join(f"{o} {c / warps:.0f}" for o, c in ops.most_common(10)))
if topn:
    seg = sorted(ins[hi + 1:], key=lambda t: -t[1])[:topn]
    print("epilogue instructions with the most samples:")
    for n, s, t in seg:
        print(f"  {s:6d} samples  x{n / warps:5.1f}  {t}")
