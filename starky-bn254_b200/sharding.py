"""Multi-GPU mode of the prover path: independent proofs sharded across ranks.

Proofs share no state (every `XExpStark` is a `Copy` value: reference src/curves/g1/exp.rs:232), so the
data path needs no collective: rank r proves the batches `assign(num_batches, world, r)` on its own GPU.
`torch.distributed` is used only for plumbing -- the max-over-ranks timing reduce and the gather of proof
digests / lengths on rank 0 (NCCL on the GPU box, gloo in the CPU tests)."""
import hashlib

import torch
import torch.distributed as dist


def assign(num_batches, world, rank):
    """Round-robin: batch b goes to rank b % world (fixed per-rank work as N grows = weak scaling)."""
    return [b for b in range(num_batches) if b % world == rank]


def max_over_ranks(values, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (timings are reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def gather_digests(local, num_batches, device="cpu"):
    """`local`: {batch index: proof bytes} proved by this rank.  Returns on every rank the list of
    (sha256 digest, length) for all `num_batches` batches, in batch order."""
    buf = torch.zeros((num_batches, 5), dtype=torch.int64, device=device)
    for b, proof in local.items():
        d = hashlib.sha256(proof).digest()
        for k in range(4):
            buf[b, k] = int.from_bytes(d[8 * k:8 * k + 8], "little", signed=True)
        buf[b, 4] = len(proof)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)   # every batch is written by exactly one rank
    out = []
    for b in range(num_batches):
        d = b"".join(int(buf[b, k]).to_bytes(8, "little", signed=True) for k in range(4))
        out.append((d.hex(), int(buf[b, 4])))
    return out
