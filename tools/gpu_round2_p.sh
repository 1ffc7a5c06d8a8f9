#!/bin/bash
# GPU session P: scan-based lookup-walk kernel (parity with the two walk kernels, effect on the G1 / G2 step) and one stream
# priority level per batch lane (SBN_LANE_PRIORITIES=1) against equal priorities, 20- and 64-proof batches, two runs each.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lookup or golden or skewed or modular_trace" > gpurun_out/r2p_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.txt
tail -3 gpurun_out/r2p_pytest.txt
run() { # name, env..., -- args
  local name=$1; shift
  timeout 600 env "$@" python bench.py --no-cpu-baseline --no-other-airs --steps ${STEPS:-20} --warmup 5 > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err || echo "$name failed"
}
STEPS=20 run walk_chunked_a SBN_LOOKUP_WALK=chunked
STEPS=20 run par_a SBN_X=0
STEPS=20 run pri_a SBN_LANE_PRIORITIES=1
STEPS=20 run walk_chunked_b SBN_LOOKUP_WALK=chunked
STEPS=20 run par_b SBN_X=0
STEPS=20 run pri_b SBN_LANE_PRIORITIES=1
STEPS=64 run par_64 SBN_X=0
STEPS=64 run pri_64 SBN_LANE_PRIORITIES=1
timeout 600 python bench.py --air g2 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2p_g2.json 2> gpurun_out/r2p_g2.err
timeout 600 env SBN_LANE_PRIORITIES=1 python bench.py --air g2 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2p_g2_pri.json 2> gpurun_out/r2p_g2_pri.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2p_*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        km = d.get("kernel_ms_per_proof", {})
        print(f.split("r2p_")[1], d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["serial_ms_per_step"], 1), "walk", km.get("lookup_walk"), "hist", km.get("lookup_hist"), "leaf", km.get("merkle_leaf_hash"))
    except Exception as e:
        print(f, "ERR", e)
PY
