#!/bin/bash
# GPU session AK: does a lane's allocator still grow inside the timed region (one warm-up proof per lane)?  + k_reduce_columns with lazy sums.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or sharded or oracle_bytes" > gpurun_out/r2ak_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ak_pytest.txt
tail -2 gpurun_out/r2ak_pytest.txt
for w in 5 12 18; do
  timeout 600 python bench.py --no-cpu-baseline --no-other-airs --steps 20 --warmup $w > gpurun_out/r2ak_w$w.json 2> gpurun_out/r2ak_w$w.err
done
python - <<'PY'
import json
for w in (5, 12, 18):
    d = json.loads(open("gpurun_out/r2ak_w%d.json" % w).read().strip().split("\n")[-1])
    km = d["kernel_ms_per_proof"]
    print(w, round(d["value"], 2), round(d["e2e"]["value"], 2), "growth", d["allocator_growth_bytes_in_timed_region"], "reduce", km.get("fri_reduce_columns"))
PY
