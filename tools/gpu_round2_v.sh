#!/bin/bash
# GPU session V: row kernels with one step type per warp (parity remap): whole GPU suite, default bench, per-kernel numbers.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.txt
tail -3 gpurun_out/r2v_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
for a in g2 fq; do timeout 600 python bench.py --air $a --no-cpu-baseline --no-other-airs --steps 10 --warmup 3 > gpurun_out/r2v_$a.json 2> gpurun_out/r2v_$a.err; done
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v_bench.json").read().strip().split("\n")[-1])
km = d["kernel_ms_per_proof"]
print(d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), {k: (round(v.get("value", 0), 2), round(v.get("e2e", {}).get("value", 0), 2)) for k, v in d.get("airs", {}).items()}, "rows", km.get("g1_rows"), "walk", km.get("lookup_walk"))
for a in ("g2", "fq"):
    d = json.loads(open("gpurun_out/r2v_%s.json" % a).read().strip().split("\n")[-1])
    km = d["kernel_ms_per_proof"]
    print(a, round(d["value"], 2), round(d["serial_ms_per_step"], 1), {k: v for k, v in km.items() if "rows" in k or "walk" in k or "chain" in k})
PY
