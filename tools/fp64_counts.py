#!/usr/bin/env python
"""Turns an ncu metrics CSV of a `bench.py` run into profiles/r02_fp64_instruction_counts.json: the FP64 instructions
the adjoint kernel really executes per algorithmic flop of the model, and its DRAM traffic.

On the GPU box:
  ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,\\
smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
      --clock-control none -k regex:adjoint --csv --log-file gpurun_out/fp64_counts.csv \\
      python bench.py --designs 148 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --horizon-scale 0.25 > gpurun_out/fp64_bench.json
Here:
  python tools/fp64_counts.py gpurun_out/fp64_counts.csv gpurun_out/fp64_bench.json
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
counts_csv, bench_json = sys.argv[1:3]
rows = [r for r in csv.reader(l for l in open(counts_csv) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    launches.setdefault(r[ix["ID"]], {"kernel": r[ix["Kernel Name"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
last = list(launches.values())[-1]  # the timed step's adjoint launch
line = json.loads([l for l in open(bench_json) if l.startswith("{")][-1])
dfma = last["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
dadd = last["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
dmul = last["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
executed = 2 * dfma + dadd + dmul
model = line["roofline"]["achieved"] * 1e12 * line["roofline"]["adjoint_ms"] * 1e-3
designs = line["config"]["designs_total"]
dram = last["dram__bytes_read.sum"] + last["dram__bytes_write.sum"]
out = {"kernel": last["kernel"], "capture": f"{designs} designs, horizon scale {line['config'].get('PROFILING_ONLY_horizon_scale', 1.0)}",
       "dfma": dfma, "dadd": dadd, "dmul": dmul, "executed_flops": executed, "model_flops": model,
       "executed_over_model": executed / model,
       "dram_bytes": dram, "algorithmic_bytes": 10.7e6 * designs * line["config"].get("PROFILING_ONLY_horizon_scale", 1.0),
       "dram_note": f"profiles/r02_fp64_instruction_counts.json: {dram / 1e9:.2f} GB DRAM traffic for {designs} designs at horizon scale "
                    f"{line['config'].get('PROFILING_ONLY_horizon_scale', 1.0)} (ncu, this build)"}
with open(os.path.join(ROOT, "profiles", "r02_fp64_instruction_counts.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
