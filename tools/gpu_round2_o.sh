#!/bin/bash
# GPU session O: final single-GPU verification of the round-2 build: gpu tests, smoke, default bench, reference arm (2 full proofs).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2o_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?" >> gpurun_out/r2o_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2o_reference.json 2> gpurun_out/r2o_reference.err; echo "ref rc=$?" >> gpurun_out/r2o_reference.err
tail -3 gpurun_out/r2o_pytest.txt; tail -2 gpurun_out/r2o_smoke.txt | cut -c1-200; tail -2 gpurun_out/r2o_bench.err; cut -c1-400 gpurun_out/r2o_reference.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_bench.json").read().strip().split("\n")[-1])
print(d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"], 2), {k: (round(v.get("value", 0), 2), round(v.get("e2e", {}).get("value", 0), 2)) for k, v in d.get("airs", {}).items()}, round(d["roofline"].get("frac") or 0, 3), d["cpu_baseline"]["value"], d["gpu_launches"])
PY
