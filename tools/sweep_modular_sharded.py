#!/usr/bin/env python3
"""BASELINE.json config 5 on several GPUs: ONE ModularStark proof (2^min .. 2^max rows, rate_bits 1) computed by all ranks
together through sbn_prove_sharded -- the LDE, the Merkle cap subtrees and the constraint evaluation are
split by LDE class, the exchanges are NCCL all-gathers.  Rank 0 prints one JSON object per size: the sharded wall time per proof
(max over ranks), the per-phase device milliseconds of rank 0, and -- when the trace fits -- the unsharded time on one GPU.
    python -m torch.distributed.run --nproc-per-node N tools/sweep_modular_sharded.py [max_log_rows=20] [min_log_rows=18] [rates=1]
(rates: comma-separated rate_bits; sizes whose LDE does not fit one GPU are proved sharded only)"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as entry  # noqa: E402
from sweep_modular import fast_ios  # noqa: E402


def main():
    max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    min_log = int(sys.argv[2]) if len(sys.argv) > 2 else 18
    rates = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1]
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    json_fd = os.dup(1); os.dup2(2, 1)   # NCCL's banner goes to stderr
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sbn = entry.load_package()
    from starky_bn254_b200 import sharding
    ctx = sbn.Context(local)
    ag = sharding.dist_allgather(device=torch.device("cuda", local)) if world > 1 else None
    agd = sharding.dist_allgather_device(torch.device("cuda", local)) if world > 1 else None
    empty = np.zeros(0, dtype=np.uint64)
    for logn in range(min_log, max_log + 1):
        n = 1 << logn
        ios = fast_ios(n)
        stark = sbn.ModularStark(n, ctx)
        for rate_bits in rates:
            cfg = stark.config()
            cfg.rate_bits = rate_bits
            res = {"rows_log2": logn, "rate_bits": rate_bits, "world": world}
            # one GPU always works since round 2: when coefficients + LDE do not fit, sbn_prove streams its commitments
            modes = (["sharded"] if world > 1 else []) + ["single"]
            for mode in modes:
                if mode == "single" and rank != 0:
                    continue
                best = None
                for rep in range(2 if (logn <= 21 or mode == "sharded") else 1):
                    tr = stark.generate_trace(ios)
                    ctx.synchronize()
                    if world > 1 and mode == "sharded":
                        dist.barrier()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    proof = sbn.prove_sharded(stark, cfg, tr, empty, rank, world, ag, allgather_device=agd) if mode == "sharded" else sbn.prove(stark, cfg, tr, empty)
                    dt = (time.perf_counter() - t0) * 1e3
                    tr.free()
                    best = (dt, proof.timings, hashlib.sha256(proof.to_bytes()).hexdigest()[:16])
                    del proof
                dt, ph, dig = best
                if mode == "sharded":
                    dt = sharding.max_over_ranks([dt], device="cuda")[0]
                    res["device_gb_per_rank"] = round(ctx.device_bytes / 1e9, 1)
                res[mode + "_prove_ms"] = round(dt, 2)
                res[mode + "_phases_rank0"] = {k: round(v, 2) for k, v in ph.items()}
                res[mode + "_proof_sha"] = dig
            if rank == 0:
                if "sharded_prove_ms" in res and "single_prove_ms" in res:
                    res["identical"] = res["sharded_proof_sha"] == res["single_proof_sha"]
                    res["speedup"] = round(res["single_prove_ms"] / res["sharded_prove_ms"], 2)
                os.write(json_fd, (json.dumps(res) + "\n").encode())
            if world > 1:
                dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
