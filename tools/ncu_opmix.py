#!/usr/bin/env python
"""Opcode mix of a kernel from an ncu report's source page: executed warp-instructions by opcode (top N) and stall
samples by opcode.   python tools/ncu_opmix.py <report.ncu-rep> [kernel-regex] [top]"""
import csv, io, re, subprocess, sys, collections

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kernels = out.split('"Kernel Name",')
for blk in kernels[1:]:
    name, rest = blk.split("\n", 1)
    if pat and not pat.search(name):
        continue
    rows = list(csv.reader(io.StringIO(rest)))
    hdr = rows[0]
    i_src, i_exec, i_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    mix, samp = collections.Counter(), collections.Counter()
    total = 0
    for r in rows[1:]:
        if len(r) <= i_exec or not r[i_exec].isdigit():
            continue
        src = r[i_src].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        op = op.split(".")[0] if not op.startswith(("MUFU", "LDTM", "STTM", "UTC")) else op
        n = int(r[i_exec])
        mix[op] += n
        samp[op] += int(r[i_samp]) if r[i_samp].isdigit() else 0
        total += n
    print(name.strip().strip('",'), "total warp-instr", total)
    for op, n in mix.most_common(top):
        print(f"  {op:16s} {n:12d} {100.0*n/total:6.2f}%   samples {samp[op]}")
    break
