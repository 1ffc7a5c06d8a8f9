#!/bin/bash
# GPU session AD: small Merkle levels with two lanes per parent: whole GPU suite, smoke, serial latency / kernel numbers.
mkdir -p gpurun_out
nproc > gpurun_out/r2ad_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ad_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ad_pytest.txt
tail -3 gpurun_out/r2ad_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ad_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2ad_smoke.txt; tail -1 gpurun_out/r2ad_smoke.txt
timeout 600 python bench.py --no-cpu-baseline --no-other-airs --steps 20 --warmup 5 > gpurun_out/r2ad_g1.json 2> gpurun_out/r2ad_g1.err
timeout 600 env SBN_LEAF_HASH_ONE_THREAD=1 python bench.py --no-cpu-baseline --no-other-airs --steps 20 --warmup 5 > gpurun_out/r2ad_g1_one_thread.json 2> gpurun_out/r2ad_g1_one_thread.err
python - <<'PY'
import json
for f in ("g1", "g1_one_thread"):
    d = json.loads(open("gpurun_out/r2ad_%s.json" % f).read().strip().split("\n")[-1])
    km = d["kernel_ms_per_proof"]
    print(f, d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), "levels", km.get("merkle_tree_levels"), "leaf", km.get("merkle_leaf_hash"), "sum", round(sum(km.values()), 1))
PY
