#!/bin/bash
# GPU session Q: second version of the scan-based lookup-walk kernel (swizzled heights, warp-interleaved placement): parity + timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lookup or golden or skewed or modular_trace" > gpurun_out/r2q_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.txt
tail -3 gpurun_out/r2q_pytest.txt
timeout 600 python bench.py --no-cpu-baseline --no-other-airs --steps 20 --warmup 5 > gpurun_out/r2q_g1.json 2> gpurun_out/r2q_g1.err
timeout 600 python bench.py --air g2 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2q_g2.json 2> gpurun_out/r2q_g2.err
timeout 600 python bench.py --air fq --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2q_fq.json 2> gpurun_out/r2q_fq.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2q_*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        km = d.get("kernel_ms_per_proof", {})
        print(f.split("r2q_")[1], d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), "walk", km.get("lookup_walk"), "hist", km.get("lookup_hist"), "leaf", km.get("merkle_leaf_hash"))
    except Exception as e:
        print(f, "ERR", e)
PY
