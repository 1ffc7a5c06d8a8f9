#!/bin/bash
# GPU session AG: openings evaluation kernel variants (columns per block x iterations in flight); huge-list switch test of the lookup kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "parallel_lookup" > gpurun_out/r2ag_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ag_pytest.txt
tail -2 gpurun_out/r2ag_pytest.txt
for v in 0 1 2 3 4; do
  timeout 600 env SBN_EVAL_VARIANT=$v python bench.py --no-cpu-baseline --no-other-airs --steps 12 --warmup 3 > gpurun_out/r2ag_v$v.json 2> gpurun_out/r2ag_v$v.err
done
python - <<'PY'
import json
for v in range(5):
    d = json.loads(open("gpurun_out/r2ag_v%d.json" % v).read().strip().split("\n")[-1])
    km = d["kernel_ms_per_proof"]
    print(v, round(d["value"], 2), round(d["e2e"]["value"], 2), "openings_eval", km.get("openings_eval"), d["proof_sha256_per_rank"])
PY
