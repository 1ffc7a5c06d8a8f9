#!/bin/bash
# GPU session AT: profile refresh on the final build (G1): launch list of the bench command + per-kernel table.
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-airs --inflight 2"
$BENCH > gpurun_out/r2at_bench_plain.json 2> gpurun_out/r2at_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2at_launches_bench.csv $BENCH > gpurun_out/r2at_ncu_launch.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed
python tools/profile_g1.py 1 > gpurun_out/r2at_prof_plain.log 2>&1 || exit 1
ncu --metrics $M --clock-control none -c 4000 --csv --log-file gpurun_out/r2at_kernel_metrics_g1.csv python tools/profile_g1.py 1 > gpurun_out/r2at_ncu_table.log 2>&1
ls -la gpurun_out/r2at_*
