#!/bin/bash
# GPU session M: NTT tile A/B (32-wide vs 16-wide tiles for the 2^8 sub-transform), leaf hash with launch bounds (128, 6), batch with the
# trace-generation gate and memory-limited lanes (long batch, large shapes).
mkdir -p gpurun_out
python tools/ntt_ab.py starky-bn254_b200/libstarkybn254_b200.so starky-bn254_b200/libstarkybn254_b200_t16.so > gpurun_out/r2m_ntt_ab.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?" >> gpurun_out/r2m_bench.err
timeout 900 python bench.py --steps 256 --no-cpu-baseline --no-other-airs > gpurun_out/r2m_g1_batch256.json 2> gpurun_out/r2m_g1_batch256.err
timeout 900 python bench.py --num-io 512 --steps 6 --no-cpu-baseline > gpurun_out/r2m_g1_n512.json 2> gpurun_out/r2m_g1_n512.err
timeout 900 python bench.py --air fq12 --num-io 128 --steps 6 --no-cpu-baseline > gpurun_out/r2m_fq12_n128.json 2> gpurun_out/r2m_fq12_n128.err
cat gpurun_out/r2m_ntt_ab.txt; tail -3 gpurun_out/r2m_pytest.txt; tail -2 gpurun_out/r2m_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m_*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"], 2), {k: round(v.get("value", 0), 2) for k, v in d.get("airs", {}).items()}, round(d["roofline"].get("frac") or 0, 3), d["kernel_ms_per_proof"].get("merkle_leaf_hash"))
    except Exception as e:
        print(f, "ERR", e)
PY
