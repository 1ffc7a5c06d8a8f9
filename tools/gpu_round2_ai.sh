#!/bin/bash
# GPU session AI: verification of the build with the loads-first openings kernel: GPU suite, smoke, default bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ai_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ai_pytest.txt
tail -3 gpurun_out/r2ai_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ai_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2ai_smoke.txt; tail -1 gpurun_out/r2ai_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ai_bench.json").read().strip().split("\n")[-1])
km = d["kernel_ms_per_proof"]
print(d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), {k: (v.get("lanes"), v.get("steps"), round(v.get("value", 0), 2), round(v.get("e2e", {}).get("value", 0), 2)) if "error" not in v else v for k, v in d.get("airs", {}).items()}, round(d["roofline"]["frac"], 3), d["cpu_baseline"]["value"], "sum", round(sum(km.values()), 1), "openings", km.get("openings_eval"))
PY
