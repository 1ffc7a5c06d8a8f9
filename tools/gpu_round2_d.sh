#!/bin/bash
# GPU session D of round 2: gpu tests (U1/U3 switches), ncu captures of the round-2 build, default bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.txt
timeout 1500 bash tools/profile_round.sh > gpurun_out/r2d_profile_round.log 2>&1
timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?" >> gpurun_out/r2d_bench.err
tail -6 gpurun_out/r2d_pytest.txt; tail -3 gpurun_out/r2d_bench.err; cut -c1-200 gpurun_out/r2d_bench.json; ls -la gpurun_out/r2d_*
