#!/bin/bash
# GPU session H: fully unrolled modular-op constraint segments (limb arrays in registers) -- tests + kernel timings for G1 / G2 / Fq12.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?" >> gpurun_out/r2h_bench.err
timeout 600 python bench.py --air g2 --steps 6 --no-cpu-baseline > gpurun_out/r2h_bench_g2.json 2> gpurun_out/r2h_bench_g2.err
timeout 600 python bench.py --air fq12 --steps 6 --no-cpu-baseline > gpurun_out/r2h_bench_fq12.json 2> gpurun_out/r2h_bench_fq12.err
tail -4 gpurun_out/r2h_pytest.txt; tail -3 gpurun_out/r2h_bench.err
python - <<'PY'
import json
for f in ("r2h_bench", "r2h_bench_g2", "r2h_bench_fq12"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read())
        k = d["kernel_ms_per_proof"]
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), {a: b for a, b in k.items() if a.startswith("q_")}, {a: round(v.get("value", 0), 2) for a, v in d.get("airs", {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
