#!/usr/bin/env python
"""One-page text summary of an `ncu --set full` report (the numbers DESIGN.md and the judge cite): duration, occupancy,
issue / FP64 / LSU pipe utilisation, stall reasons per issue, instruction mix by opcode, local-memory and DRAM traffic,
SASS evidence of tensor-memory use.   python tools/ncu_summary.py report.ncu-rep > profiles/<name>_ncu_summary.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def g(name):
    v, u = m.get(name, ("n/a", ""))
    return f"{v} {u}".strip()


print(f"report: {rep}")
print(f"kernel: {m.get('Kernel Name', ('?',))[0]}   grid {g('Grid Size')} block {g('Block Size')}")
for k in ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_local_ld.sum",
          "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum",
          "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
          "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"):
    print(f"  {k:78s} {g(k)}")
print("stall reasons (warps stalled per issue-active cycle):")
st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v))
      for h, v in zip(hdr, vals) if "issue_stalled" in h and "per_issue_active" in h]
for k, v in sorted(st, key=lambda x: -x[1]):
    if v > 0.05:
        print(f"  {k:24s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
r2 = list(csv.reader(io.StringIO(src)))
h2 = r2[1]
c, tot = Counter(), 0
for r in r2[2:]:
    if len(r) != len(h2):
        continue
    d = dict(zip(h2, r))
    t = d["Source"].split()
    if not t:
        continue
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    n = int(d["Instructions Executed"] or 0)
    c[op] += n
    tot += n
print(f"executed warp instructions by opcode ({len(r2) - 2} SASS instructions in the kernel):")
print("  " + "  ".join(f"{k} {100 * v / tot:.1f}%" for k, v in c.most_common(22)))
fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"  FP64 arithmetic share {100 * fp64 / tot:.1f}%   local (spill) share {100 * (c['LDL'] + c['STL']) / tot:.1f}%   "
      f"tensor-memory share {100 * (c['LDTM'] + c['STTM']) / tot:.1f}%")
