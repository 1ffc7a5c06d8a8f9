#!/bin/bash
# GPU session F: lanes sweep for the small-trace AIR (Fq12: 2^14 leaves per commitment), G1 lanes 8.
mkdir -p gpurun_out
for L in 9 12 16; do timeout 600 python bench.py --air fq12 --steps 16 --inflight $L --no-cpu-baseline > gpurun_out/r2f_fq12_l$L.json 2> gpurun_out/r2f_fq12_l$L.err; done
for L in 4 8; do timeout 600 python bench.py --steps 16 --inflight $L --no-cpu-baseline --no-other-airs > gpurun_out/r2f_g1_l$L.json 2> gpurun_out/r2f_g1_l$L.err; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2f_*.json")):
    try:
        d = json.loads(open(f).read())
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), d["inflight_per_gpu"])
    except Exception as e:
        print(f, "ERR", e)
PY
