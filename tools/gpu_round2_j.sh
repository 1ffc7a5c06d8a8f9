#!/bin/bash
# GPU session J (8 GPUs): bench.py with the sleep-backoff host waits (8-GPU weak-scaling efficiency), 1-GPU line on the same box for the ratio.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-other-airs > gpurun_out/r2j_bench_1gpu.json 2> gpurun_out/r2j_bench_1gpu.err; echo "bench1 rc=$?" >> gpurun_out/r2j_bench_1gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 12 --warmup 3 > gpurun_out/r2j_bench_8gpu.json 2> gpurun_out/r2j_bench_8gpu.err; echo "bench8 rc=$?" >> gpurun_out/r2j_bench_8gpu.err
tail -2 gpurun_out/r2j_bench_1gpu.err; tail -2 gpurun_out/r2j_bench_8gpu.err
python - <<'PY'
import json
for f in ("r2j_bench_1gpu", "r2j_bench_8gpu"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().split("\n")[-1])
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"], 2), round(d["serial_ms_per_step"], 1), json.dumps(d["intra_proof"])[:300], {k: round(v.get("value", 0), 2) for k, v in d.get("airs", {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
