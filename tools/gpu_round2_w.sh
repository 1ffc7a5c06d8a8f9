#!/bin/bash
# GPU session W: lanes per batch for the AIRs with small proofs (Fq12: 2^13 rows, leaf hash starved by 128-block launches; Fq).
mkdir -p gpurun_out
for spec in fq12:6 fq12:10 fq12:16 fq12:24 fq:6 fq:8 fq:12 g2:6 g2:8 g1:7; do
  a=${spec%%:*}; n=${spec##*:}
  timeout 600 python bench.py --air $a --inflight $n --no-cpu-baseline --no-other-airs --steps 24 --warmup 5 > gpurun_out/r2w_${a}_$n.json 2> gpurun_out/r2w_${a}_$n.err || echo "$spec failed"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2w_*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f.split("r2w_")[1], d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1))
    except Exception as e:
        print(f, "ERR", e)
PY
