#!/usr/bin/env python
"""Joins the per-instruction stall samples of an ncu report (`--page source --csv`, SASS view) with the line table of
the same kernel in an nvdisasm listing (`nvdisasm -g -c file.cubin`), and prints the hottest source lines and the
stall reasons of each.  Usage: ncu_by_line.py samples.csv listing.sass mangled_kernel_name [top_n]"""
import csv
import re
import sys
from collections import defaultdict

samples, listing, kernel = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 50
rows = list(csv.reader(open(samples)))
hdr = rows[1]
inst = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
# listing: instructions of the kernel in order, each with the (file, line) in effect
lines = open(listing).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kernel + ":"))
cur, seq = ("?", 0), []
pat_line = re.compile(r'//## File "([^"]+)", line (\d+)(.*)')
pat_inst = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
for l in lines[start + 1:]:
    if l.startswith("//--------------------- .text.") or l.startswith(".text."):
        break
    m = pat_line.search(l)
    if m:
        # "inlined at" chains: keep the innermost location but remember the outermost adjoint3 line
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = pat_inst.match(l)
    if m:
        seq.append((cur, m.group(2)))
assert len(seq) == len(inst), (len(seq), len(inst))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_line = defaultdict(lambda: defaultdict(float))
tot = 0
for (loc, text), d in zip(seq, inst):
    n = int(d["# Samples"] or 0)
    tot += n
    by_line[loc]["samples"] += n
    by_line[loc]["inst"] += int(d["Instructions Executed"] or 0)
    for c in stall_cols:
        by_line[loc][c] += float(d[c] or 0)
    op = text.split()[0] if not text.startswith("@") else text.split()[1]
    if op.startswith(("LDL", "STL")):
        by_line[loc]["local_inst"] += int(d["Instructions Executed"] or 0)
print("total samples", tot, "instructions", len(inst))
for loc, v in sorted(by_line.items(), key=lambda kv: -kv[1]["samples"])[:top_n]:
    st = sorted(((c[6:], v[c]) for c in stall_cols if v[c] > 0), key=lambda x: -x[1])[:4]
    print(f"{loc[0]}:{loc[1]:<5d} {100 * v['samples'] / tot:5.1f}%  inst {int(v['inst']):>10d}  local {int(v['local_inst']):>9d}  " +
          " ".join(f"{k}={int(x)}" for k, x in st))
