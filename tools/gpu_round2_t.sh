#!/bin/bash
# GPU session T: full ncu capture of the third version of the scan-based lookup kernel.
mkdir -p gpurun_out
python tools/profile_g1.py 1 > gpurun_out/r2t_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_lookup_walk_parallel" -c 1 -o gpurun_out/r2t_prof_walk -f python tools/profile_g1.py 1 > gpurun_out/r2t_ncu_walk.log 2>&1
tail -3 gpurun_out/r2t_ncu_walk.log
ls -la gpurun_out/r2t_prof_walk.ncu-rep
