#!/bin/bash
# GPU session K (8 GPUs): config 5 at 2^22 rows x rate_bits 1..3 -- ONE proof on 8 GPUs (sbn_prove_sharded) vs one GPU (streamed), digests compared.
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/sweep_modular_sharded.py 22 22 1,2,3 > gpurun_out/r2k_sweep_sharded_8gpu_2p22.jsonl 2> gpurun_out/r2k_sweep.err; echo "rc=$?" >> gpurun_out/r2k_sweep.err
tail -3 gpurun_out/r2k_sweep.err; cut -c1-700 gpurun_out/r2k_sweep_sharded_8gpu_2p22.jsonl
