#!/bin/bash
# GPU session AS: two-lane leaf hash extended to trees of up to 2^16 leaves: whole GPU suite + sharded latency on one GPU (threads) + default bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2as_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2as_pytest.txt
tail -3 gpurun_out/r2as_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2as_bench.json 2> gpurun_out/r2as_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2as_bench.json").read().strip().split("\n")[-1])
print(d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), {k: (v.get("lanes"), v.get("steps"), round(v.get("value", 0), 2), round(v.get("e2e", {}).get("value", 0), 2)) if "error" not in v else v for k, v in d.get("airs", {}).items()}, round(d["roofline"]["frac"], 3))
PY
