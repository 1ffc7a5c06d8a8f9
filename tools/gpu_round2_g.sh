#!/bin/bash
# GPU session G (2 GPUs): gpu tests, then the driver's multi-GPU launch of bench.py (throughput + intra-proof latency) and the reference arm's rank handling.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err; echo "bench rc=$?" >> gpurun_out/r2g_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --air fq12 > gpurun_out/r2g_bench_2gpu_fq12.json 2> gpurun_out/r2g_bench_2gpu_fq12.err
tail -4 gpurun_out/r2g_pytest.txt; tail -4 gpurun_out/r2g_bench_2gpu.err
python - <<'PY'
import json
for f in ("r2g_bench_2gpu", "r2g_bench_2gpu_fq12"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().split("\n")[-1])
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), json.dumps(d["intra_proof"])[:600], {k: round(v.get("value", 0), 2) for k, v in d.get("airs", {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
