#!/bin/bash
# GPU session AP: final verification of the round-2 build (clean rebuild): GPU suite, smoke, default bench with the driver's flags, reference arm (1 step).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ap_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ap_pytest.txt
tail -3 gpurun_out/r2ap_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ap_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2ap_smoke.txt; tail -2 gpurun_out/r2ap_smoke.txt | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ap_bench.json 2> gpurun_out/r2ap_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2ap_reference.json 2> gpurun_out/r2ap_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ap_bench.json").read().strip().split("\n")[-1])
km = d["kernel_ms_per_proof"]
print(d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), {k: (v.get("lanes"), v.get("steps"), round(v.get("value", 0), 2), round(v.get("e2e", {}).get("value", 0), 2)) if "error" not in v else v for k, v in d.get("airs", {}).items()}, round(d["roofline"]["frac"], 3), d["cpu_baseline"]["value"], "sum", round(sum(km.values()), 1), d["gpu_launches"])
r = json.loads(open("gpurun_out/r2ap_reference.json").read().strip().split("\n")[-1])
print("reference", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"])
PY
