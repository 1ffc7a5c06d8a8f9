#!/bin/bash
# GPU session I (8 GPUs): two-device test, the driver's 8-GPU launch of bench.py, and the reference arm's multi-rank behaviour.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_contexts or prove_batch" > gpurun_out/r2i_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 12 --warmup 3 > gpurun_out/r2i_bench_8gpu.json 2> gpurun_out/r2i_bench_8gpu.err; echo "bench rc=$?" >> gpurun_out/r2i_bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 12 --warmup 3 --no-other-airs > gpurun_out/r2i_bench_4gpu.json 2> gpurun_out/r2i_bench_4gpu.err; echo "bench rc=$?" >> gpurun_out/r2i_bench_4gpu.err
tail -4 gpurun_out/r2i_pytest.txt; tail -3 gpurun_out/r2i_bench_8gpu.err; tail -3 gpurun_out/r2i_bench_4gpu.err; nproc
python - <<'PY'
import json
for f in ("r2i_bench_8gpu", "r2i_bench_4gpu"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().split("\n")[-1])
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"], 2), json.dumps(d["intra_proof"])[:900], {k: round(v.get("value", 0), 2) for k, v in d.get("airs", {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
