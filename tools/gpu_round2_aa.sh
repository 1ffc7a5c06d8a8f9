#!/bin/bash
# GPU session AA: leaf hash with two lanes per permutation for small trees: whole GPU suite + Fq12 / G1 bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_pytest.txt
tail -3 gpurun_out/r2aa_pytest.txt
timeout 600 python bench.py --air fq12 --no-cpu-baseline --no-other-airs --steps 32 --warmup 5 > gpurun_out/r2aa_fq12.json 2> gpurun_out/r2aa_fq12.err
timeout 600 env SBN_LEAF_HASH_ONE_THREAD=1 python bench.py --air fq12 --no-cpu-baseline --no-other-airs --steps 32 --warmup 5 > gpurun_out/r2aa_fq12_one_thread.json 2> gpurun_out/r2aa_fq12_one_thread.err
timeout 600 python bench.py --air fq12 --no-cpu-baseline --no-other-airs --steps 32 --warmup 5 > gpurun_out/r2aa_fq12_b.json 2> gpurun_out/r2aa_fq12_b.err
timeout 600 env SBN_LEAF_HASH_ONE_THREAD=1 python bench.py --air fq12 --no-cpu-baseline --no-other-airs --steps 32 --warmup 5 > gpurun_out/r2aa_fq12_one_thread_b.json 2> gpurun_out/r2aa_fq12_one_thread_b.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2aa_*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        km = d.get("kernel_ms_per_proof", {})
        print(f.split("r2aa_")[1], d["steps"], d.get("inflight_per_gpu"), round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), "leaf", km.get("merkle_leaf_hash"))
    except Exception as e:
        print(f, "ERR", e)
PY
