#!/bin/bash
# GPU session A of round 2: toolchain probe, Poseidon before/after micro-benchmark, pipe throughput (incl. FP64), tests, bench.
mkdir -p gpurun_out
{ echo "== cargo probe"; which cargo rustc 2>&1; ls -d ~/.cargo ~/.cargo/registry ~/.cargo/git ~/.rustup 2>&1; echo "nproc $(nproc)"; nvidia-smi -L; } > gpurun_out/r2a_probe.txt 2>&1
cd tools/microbench
{ ./pb_r01 17 1676; ./pb_cur 17 1676; ./pb_r01 14 9808; ./pb_cur 14 9808; } > ../../gpurun_out/r2a_poseidon_ab.txt 2>&1
./pb_int_throughput > ../../gpurun_out/r2a_int_throughput.txt 2>&1
cd ../..
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.txt
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_g1.json 2> gpurun_out/r2a_bench_g1.err
timeout 600 python bench.py --air fq12 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_fq12.json 2> gpurun_out/r2a_bench_fq12.err
tail -3 gpurun_out/r2a_pytest.txt; cat gpurun_out/r2a_poseidon_ab.txt; cut -c1-400 gpurun_out/r2a_bench_g1.json
