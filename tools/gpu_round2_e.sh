#!/bin/bash
# GPU session E: gpu tests with the low-latency Montgomery product in the chain kernels, default bench, G2/Fq12 kernel timings.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.txt
timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?" >> gpurun_out/r2e_bench.err
timeout 600 python bench.py --air g2 --steps 6 --no-cpu-baseline > gpurun_out/r2e_bench_g2.json 2> gpurun_out/r2e_bench_g2.err
timeout 600 python bench.py --air fq12 --steps 6 --no-cpu-baseline > gpurun_out/r2e_bench_fq12.json 2> gpurun_out/r2e_bench_fq12.err
tail -4 gpurun_out/r2e_pytest.txt; tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
for f in ("r2e_bench", "r2e_bench_g2", "r2e_bench_fq12"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read())
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), {k: v for k, v in list(d["kernel_ms_per_proof"].items())[:8]})
    except Exception as e:
        print(f, "ERR", e)
PY
