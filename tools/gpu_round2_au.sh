#!/bin/bash
# GPU session AU: default bench after the sched_getaffinity change.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2au_bench.json 2> gpurun_out/r2au_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2au_bench.json").read().strip().split("\n")[-1])
print(d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), {k: (v.get("lanes"), v.get("steps"), round(v.get("value", 0), 2), round(v.get("e2e", {}).get("value", 0), 2)) if "error" not in v else v for k, v in d.get("airs", {}).items()}, round(d["roofline"]["frac"], 3))
PY
