#!/usr/bin/env python
"""Lists the local-memory (spill) instructions of a kernel by source line from `nvdisasm -g -c` output.
Usage: sass_spills.py listing.sass mangled_kernel_name"""
import re
import sys
from collections import Counter

listing, kernel = sys.argv[1:3]
lines = open(listing).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kernel + ":"))
cur = ("?", 0)
pat_line = re.compile(r'//## File "([^"]+)", line (\d+)(.*)')
pat_inst = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
cnt, total, ops = Counter(), 0, Counter()
for l in lines[start + 1:]:
    if l.startswith("//--------------------- .text.") or l.startswith(".text."):
        break
    m = pat_line.search(l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = pat_inst.match(l)
    if m:
        total += 1
        t = m.group(2).split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
        if op.startswith(("LDL", "STL")):
            cnt[(cur, op.split(".")[0])] += 1
print("instructions", total, "LDL", ops["LDL"], "STL", ops["STL"])
for (loc, op), n in sorted(cnt.items(), key=lambda kv: (kv[0][0][0], kv[0][0][1])):
    print(f"{loc[0]}:{loc[1]:<5d} {op} x{n}")
print(ops.most_common(25))
