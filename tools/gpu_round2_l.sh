#!/bin/bash
# GPU session L: final single-GPU check of the round-2 build (gpu tests, smoke, default bench) + BASELINE config 2 (batches of B = 8 / 64 / 512
# independent G1 proofs in ONE sbn_prove_batch call; 256 / 512 instances per proof) + Fq12 with 128 instances.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2l_smoke.txt
timeout 900 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?" >> gpurun_out/r2l_bench.err
for B in 8 64 512; do timeout 900 python bench.py --steps $B --no-cpu-baseline --no-other-airs > gpurun_out/r2l_g1_batch$B.json 2> gpurun_out/r2l_g1_batch$B.err; done
for n in 256 512; do timeout 900 python bench.py --num-io $n --steps 8 --no-cpu-baseline > gpurun_out/r2l_g1_n$n.json 2> gpurun_out/r2l_g1_n$n.err; done
timeout 900 python bench.py --air fq12 --num-io 128 --steps 6 --no-cpu-baseline > gpurun_out/r2l_fq12_n128.json 2> gpurun_out/r2l_fq12_n128.err
tail -3 gpurun_out/r2l_pytest.txt; tail -2 gpurun_out/r2l_smoke.txt; tail -2 gpurun_out/r2l_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2l_*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"], 2), round(d.get("instances_per_s", 0), 1), {k: round(v.get("value", 0), 2) for k, v in d.get("airs", {}).items()}, round(d["roofline"].get("frac") or 0, 3))
    except Exception as e:
        print(f, "ERR", e)
PY
