#!/usr/bin/env python3
"""Opcode histogram of the hot loop(s) of a kernel, read from the SASS of an object file (no GPU needed).
usage: tools/sass_loop.py <object.o> <mangled-kernel-substring> [min_loop_len]"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for line in f.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, s) in enumerate(ins):
        m2 = re.search(r"BRA.*0x([0-9a-f]+)", s)
        if m2:
            t = int(m2.group(1), 16)
            if t in addr and addr[t] < i:
                loops.append((i - addr[t] + 1, addr[t], i))
    print(f"== {name}: {len(ins)} instructions, {len(loops)} backward branches")
    for n, lo, hi in sorted(loops, reverse=True):
        if n < minlen:
            continue
        ops = collections.Counter()
        for _, s in ins[lo:hi + 1]:
            s = re.sub(r"^@!?U?P\w+\s+", "", s)
            ops[s.split()[0].split(".")[0]] += 1
        fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
        mov = sum(v for k, v in ops.items() if k in ("IMAD", "MOV", "LOP3"))
        print(f"loop of {n} instr: FP64 {fp64}, SHFL {ops['SHFL']}, IMAD/MOV/LOP3 {mov} :: " +
              ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
