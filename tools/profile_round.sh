# The ncu captures behind profiles/r02_* (run on the GPU box: gpurun -- 'bash tools/profile_round.sh'): each capture only after the plain
# run of the same command has exited 0; summaries are made with tools/launch_list_summary.py, tools/kernel_table.py, tools/ncu_summary.py.
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-airs --inflight 2"
$BENCH > gpurun_out/r2d_bench_plain.json 2> gpurun_out/r2d_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2d_launches_bench.csv $BENCH > gpurun_out/r2d_ncu_launch.log 2>&1
python tools/profile_g1.py 1 > gpurun_out/r2d_prof_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 4000 --csv --log-file gpurun_out/r2d_kernel_metrics_g1.csv python tools/profile_g1.py 1 > gpurun_out/r2d_ncu_table.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_leaf_hash" -c 1 -o gpurun_out/r2d_prof_leaf -f python tools/profile_g1.py 1 > gpurun_out/r2d_ncu_leaf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ntt_pass1|k_ntt_pass2" -s 40 -c 2 -o gpurun_out/r2d_prof_ntt -f python tools/profile_g1.py 1 > gpurun_out/r2d_ncu_ntt.log 2>&1
