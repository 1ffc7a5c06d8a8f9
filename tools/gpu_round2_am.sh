#!/bin/bash
# GPU session AM: full ncu capture of k_g1_chain (what binds the exponentiation chain?).
mkdir -p gpurun_out
python tools/profile_g1.py 1 > gpurun_out/r2am_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_g1_chain" -c 1 -o gpurun_out/r2am_prof_chain -f python tools/profile_g1.py 1 > gpurun_out/r2am_ncu_chain.log 2>&1
tail -3 gpurun_out/r2am_ncu_chain.log
ls -la gpurun_out/r2am_prof_chain.ncu-rep
