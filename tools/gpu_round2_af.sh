#!/bin/bash
# GPU session AF: split-u8 columns in one launch (was one launch per target: 1 332 per Fq12 proof): whole GPU suite + Fq12 bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2af_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2af_pytest.txt
tail -3 gpurun_out/r2af_pytest.txt
timeout 600 python bench.py --air fq12 --no-cpu-baseline --no-other-airs --steps 32 --warmup 5 > gpurun_out/r2af_fq12.json 2> gpurun_out/r2af_fq12.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2af_fq12.json").read().strip().split("\n")[-1])
km = d["kernel_ms_per_proof"]
print(d["steps"], d.get("inflight_per_gpu"), round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), d["gpu_launches"], "sum", round(sum(km.values()), 1), {k: v for k, v in list(km.items())[:8]})
PY
