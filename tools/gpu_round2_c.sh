#!/bin/bash
# GPU session C of round 2: full gpu test suite, default bench, compute-sanitizer, config-5 sweep incl. 2^22 rows (streamed).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.txt
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?" >> gpurun_out/r2c_bench.err
gcc -std=c99 -I include tools/sanitize/prove_file.c -o /tmp/prove_file -L starky-bn254_b200 -lstarkybn254_b200 -Wl,-rpath,$PWD/starky-bn254_b200
{ echo "== memcheck ModularStark 512"; timeout 300 compute-sanitizer --tool memcheck /tmp/prove_file 0 512 tools/sanitize/modular_512.ios 2>&1 | tail -6;
  echo "== memcheck G1ExpStark 128"; timeout 300 compute-sanitizer --tool memcheck /tmp/prove_file 2 128 tools/sanitize/g1_128.ios 2>&1 | tail -6;
  echo "== memcheck Fq12ExpStark 2"; timeout 300 compute-sanitizer --tool memcheck /tmp/prove_file 4 2 tools/sanitize/fq12_2.ios 2>&1 | tail -6;
  echo "== memcheck batch (2 lanes) G1ExpStark 128"; timeout 300 compute-sanitizer --tool memcheck /tmp/prove_file 2 128 tools/sanitize/g1_128.ios 2 2>&1 | tail -6;
  echo "== racecheck ModularStark 512"; timeout 300 compute-sanitizer --tool racecheck /tmp/prove_file 0 512 tools/sanitize/modular_512.ios 2>&1 | tail -6;
  echo "== racecheck G1ExpStark 128"; timeout 420 compute-sanitizer --tool racecheck /tmp/prove_file 2 128 tools/sanitize/g1_128.ios 2>&1 | tail -6;
  echo "== racecheck G1ExpStark 128, SBN_LOOKUP_SEQUENTIAL=1"; SBN_LOOKUP_SEQUENTIAL=1 timeout 420 compute-sanitizer --tool racecheck /tmp/prove_file 2 128 tools/sanitize/g1_128.ios 2>&1 | tail -6;
  echo "== synccheck G1ExpStark 128"; timeout 300 compute-sanitizer --tool synccheck /tmp/prove_file 2 128 tools/sanitize/g1_128.ios 2>&1 | tail -6;
} > gpurun_out/r2c_sanitizer.txt 2>&1
SBN_SWEEP_MIN_LOG=20 SBN_SWEEP_MAX_LOG=22 timeout 1500 python bench.py --sweep modular > gpurun_out/r2c_sweep_20_22.jsonl 2> gpurun_out/r2c_sweep.err
tail -6 gpurun_out/r2c_pytest.txt; tail -3 gpurun_out/r2c_bench.err; cut -c1-300 gpurun_out/r2c_bench.json; cat gpurun_out/r2c_sanitizer.txt | tail -50; cat gpurun_out/r2c_sweep_20_22.jsonl; tail -5 gpurun_out/r2c_sweep.err
