set -x
python tools/profile_g1.py 1 > gpurun_out/prof_plain_r1v.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 4000 --csv --log-file gpurun_out/kernel_metrics_g1_r1v.csv python tools/profile_g1.py 1 > gpurun_out/ncu_table_r1v.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_leaf_hash" -c 1 -o gpurun_out/prof_leaf_r1v -f python tools/profile_g1.py 1 > gpurun_out/ncu_leaf_r1v.log 2>&1
