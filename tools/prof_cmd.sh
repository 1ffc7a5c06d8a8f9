set -x
python tools/profile_g1.py 2 > gpurun_out/prof_plain_r1s.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_ntt_pass1|k_ntt_pass2" -s 40 -c 2 -o gpurun_out/prof_ntt_r1s -f python tools/profile_g1.py 1 > gpurun_out/ncu_ntt_r1s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_leaf_hash" -c 1 -o gpurun_out/prof_leaf_r1s -f python tools/profile_g1.py 1 > gpurun_out/ncu_leaf_r1s.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1s.csv python tools/profile_g1.py 2 > gpurun_out/ncu_launch_r1s.log 2>&1
