#!/bin/bash
# GPU session B of round 2: tests, Poseidon A/B with the corrected stitch, default bench (batch API, all AIRs, computed roofline).
mkdir -p gpurun_out
cd tools/microbench
{ ./pb_r01 17 1676; ./pb_cur 17 1676; ./pb_cur 14 9808; } > ../../gpurun_out/r2b_poseidon_ab.txt 2>&1
./pb_int_throughput > ../../gpurun_out/r2b_int_throughput.txt 2>&1
cd ../..
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.txt
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?" >> gpurun_out/r2b_bench.err
tail -15 gpurun_out/r2b_pytest.txt; cat gpurun_out/r2b_poseidon_ab.txt; tail -5 gpurun_out/r2b_bench.err; cut -c1-1500 gpurun_out/r2b_bench.json
