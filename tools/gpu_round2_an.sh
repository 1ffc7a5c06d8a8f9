#!/bin/bash
# GPU session AN: exponentiation chains as doubling chain + prefix-sum scan: whole GPU suite, kernel times / serial latency G1, G2.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2an_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2an_pytest.txt
tail -3 gpurun_out/r2an_pytest.txt
for a in g1 g2; do timeout 600 python bench.py --air $a --no-cpu-baseline --no-other-airs --steps 16 --warmup 5 > gpurun_out/r2an_$a.json 2> gpurun_out/r2an_$a.err; done
python - <<'PY'
import json
for a in ("g1", "g2"):
    d = json.loads(open("gpurun_out/r2an_%s.json" % a).read().strip().split("\n")[-1])
    km = d["kernel_ms_per_proof"]
    print(a, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), "sum", round(sum(km.values()), 1), {k: v for k, v in km.items() if "chain" in k or "rows" in k or "affine" in k})
PY
