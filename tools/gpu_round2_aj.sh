#!/bin/bash
# GPU session AJ (2 or 4 GPUs): the driver's scaling command on the final build, headline AIR only (replicas + one proof on N GPUs).
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 20 --warmup 5 --no-other-airs --no-cpu-baseline > gpurun_out/r2aj_bench_${N}gpu.json 2> gpurun_out/r2aj_bench_${N}gpu.err; echo "bench rc=$?"
python - $N <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2aj_bench_%sgpu.json" % sys.argv[1]).read().strip().split("\n")[-1])
ip = d.get("intra_proof") or {}
print(d["n_gpus"], d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), "intra", ip.get("ms_per_proof"), ip.get("phase_ms_rank0"))
PY
