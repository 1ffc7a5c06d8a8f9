# The ncu captures behind profiles/ (run on the GPU box: gpurun -- 'bash tools/profile_round.sh'): each capture only after the plain run
# of the same command has exited 0; summaries are made here with tools/launch_list_summary.py, tools/kernel_table.py, tools/ncu_summary.py.
set -x
python tools/profile_g1.py 2 > gpurun_out/prof_plain_r1z.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1z.csv python tools/profile_g1.py 2 > gpurun_out/ncu_launch_r1z.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 4000 --csv --log-file gpurun_out/kernel_metrics_g1_r1z.csv python tools/profile_g1.py 1 > gpurun_out/ncu_table_r1z.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_leaf_hash" -c 1 -o gpurun_out/prof_leaf_r1z -f python tools/profile_g1.py 1 > gpurun_out/ncu_leaf_r1z.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ntt_pass1|k_ntt_pass2" -s 40 -c 2 -o gpurun_out/prof_ntt_r1z -f python tools/profile_g1.py 1 > gpurun_out/ncu_ntt_r1z.log 2>&1
python bench.py > gpurun_out/bench_g1_final_r1.json 2> gpurun_out/bench_g1_final_r1.err
