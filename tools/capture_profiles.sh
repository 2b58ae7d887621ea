#!/bin/bash
# Round-end measurement batch for one B200 (run under gpurun from the repo root); everything lands in gpurun_out/.
#   bench line of the default command, ncu launch list of a short bench run, FP64 instruction counters of the adjoint
#   kernel, one `ncu --set full` capture of the adjoint and of the forward kernel (quarter horizon, one wave).
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
tail -c 400 gpurun_out/bench_r02_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --designs 296 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launch_bench.log 2>&1
ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:adjoint --csv --log-file gpurun_out/fp64_counts.csv \
    python bench.py --designs 148 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --horizon-scale 0.25 > gpurun_out/fp64_bench.json 2> gpurun_out/fp64_bench.err
ncu --set full --clock-control none --import-source on -k regex:adjoint3 -c 1 -f -o gpurun_out/prof_a3_r02_end \
    python bench.py --designs 148 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --horizon-scale 0.25 > gpurun_out/ncu_a3_end.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:forward2 -c 1 -f -o gpurun_out/prof_f2_r02_end \
    python bench.py --designs 296 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --horizon-scale 0.25 > gpurun_out/ncu_f2_end.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
