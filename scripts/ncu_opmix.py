#!/usr/bin/env python
"""Instruction mix / stall samples by SASS opcode from `ncu -i X.ncu-rep --page source --csv`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iSt = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ops, stall, tot, tst = collections.Counter(), collections.Counter(), 0, 0
for r in rows[2:]:
    try:
        n, st = int(r[iE]), int(r[iSt])
    except (ValueError, IndexError):
        continue
    tok = r[iS].strip().split()
    op = tok[1] if tok[0].startswith("@") else tok[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "LDS", "STG", "LDG", "STS", "BAR", "SYNCS")) else op.split(".")[0]
    ops[op] += n; stall[op] += st; tot += n; tst += st
print(f"{tot} warp-instructions, {tst} stall samples")
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{k:14s} {v:11d} {100 * v / tot:5.1f}%   stall samples {stall[k]:7d} {100 * stall[k] / max(1, tst):5.1f}%")
