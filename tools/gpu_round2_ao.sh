#!/bin/bash
# GPU session AO (8 GPUs): the driver's scaling command on the build with the scan-based chains (replicas + one proof on 8 GPUs).
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2ao_bench_8gpu.json 2> gpurun_out/r2ao_bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ao_bench_8gpu.json").read().strip().split("\n")[-1])
ip = d.get("intra_proof") or {}
print(d["n_gpus"], d["steps"], round(d["value"], 2), round(d["e2e"]["value"], 2), {k: (v.get("lanes"), round(v.get("value", 0), 2)) if "error" not in v else v for k, v in d.get("airs", {}).items()}, "intra", ip.get("ms_per_proof"), ip.get("phase_ms_rank0"))
PY
