#!/bin/bash
# GPU session N: leaf hash at 64 registers (launch bounds (128, 8): one free block slot per SM beside a resident leaf-hash wave),
# k_g1_chain capped at 128 registers, trace-generation gate: default bench + long batch.
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?" >> gpurun_out/r2n_bench.err
timeout 900 python bench.py --steps 256 --no-cpu-baseline --no-other-airs > gpurun_out/r2n_g1_batch256.json 2> gpurun_out/r2n_g1_batch256.err
timeout 600 python -m pytest tests -m gpu -x -q -k "poseidon or g1_trace or prove_batch or commit_columns" > gpurun_out/r2n_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.txt
tail -3 gpurun_out/r2n_pytest.txt; tail -2 gpurun_out/r2n_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n_*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"], 2), {k: round(v.get("value", 0), 2) for k, v in d.get("airs", {}).items()}, round(d["roofline"].get("frac") or 0, 3), d["kernel_ms_per_proof"].get("merkle_leaf_hash"), d["kernel_ms_per_proof"].get("g1_chain"), round(d["serial_ms_per_step"], 1))
    except Exception as e:
        print(f, "ERR", e)
PY
