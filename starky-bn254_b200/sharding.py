"""Multi-GPU mode of the prover path: independent proofs sharded across ranks.

Proofs share no state (every `XExpStark` is a `Copy` value: reference src/curves/g1/exp.rs:232), so the
data path needs no collective: rank r proves the batches `assign(num_batches, world, r)` on its own GPU.
`torch.distributed` is used only for plumbing -- the max-over-ranks timing reduce and the gather of proof
digests / lengths on rank 0 (NCCL on the GPU box, gloo in the CPU tests).

Intra-proof sharding (`prove_sharded`): the ranks of a group compute ONE proof together; the only exchanges
are all-gathers of Merkle cap digests, quotient values and opened rows.  `dist_allgather` runs them over torch.distributed
(NCCL between GPUs, gloo on CPU); `ThreadGroup` is the in-process version for ranks simulated as threads on one GPU."""
import hashlib
import threading

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


def dist_allgather(group=None, device=None):
    """all-gather of equal-length byte strings over torch.distributed for `prove_sharded`: `device` = the rank's CUDA device
    under NCCL (the bytes are staged through a device tensor so the transfer runs over NVLink), None under gloo.
    The returned function also carries `.raw(send_ptr, nbytes, recv_ptr)`, which `prove_sharded` prefers: it works on the
    library's host buffers in place (no Python-level copies; the opened rows of a wide AIR are megabytes per rank)."""
    import ctypes
    world = dist.get_world_size(group)

    def allgather(data):
        t = torch.frombuffer(bytearray(data), dtype=torch.uint8)
        if device is not None:
            t = t.to(device)
        out = torch.empty(world * len(data), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(out, t, group=group)
        raw = out.cpu().numpy().tobytes()
        return [raw[i * len(data):(i + 1) * len(data)] for i in range(world)]

    def raw(send_ptr, nbytes, recv_ptr):
        s = torch.frombuffer((ctypes.c_ubyte * nbytes).from_address(send_ptr), dtype=torch.uint8)
        r = torch.frombuffer((ctypes.c_ubyte * (nbytes * world)).from_address(recv_ptr), dtype=torch.uint8)
        if device is None:
            dist.all_gather_into_tensor(r, s, group=group)
        else:
            out = torch.empty(world * nbytes, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(out, s.to(device), group=group)
            r.copy_(out)

    allgather.raw = raw
    return allgather


class _DevView:
    """A raw device pointer as a 1-D uint8 array for torch.as_tensor (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def dist_allgather_device(device, group=None):
    """Device-to-device all-gather for `prove_sharded(..., allgather_device=...)`: NCCL all_gather_into_tensor directly on the
    library's buffers (viewed through __cuda_array_interface__), no host staging."""
    world = dist.get_world_size(group)

    def allgather(send_ptr, nbytes, recv_ptr):
        s = torch.as_tensor(_DevView(send_ptr, nbytes), device=device)
        r = torch.as_tensor(_DevView(recv_ptr, nbytes * world), device=device)
        dist.all_gather_into_tensor(r, s, group=group)
        torch.cuda.current_stream(device).synchronize()

    return allgather


class ThreadGroup:
    """In-process exchange for `world` ranks running as threads (one sbn context each, same or different GPUs)."""

    def __init__(self, world):
        self.world = world
        self.slots = [None] * world
        self.barrier = threading.Barrier(world)

    def allgather_device(self, rank, device=0):
        """Device-buffer all-gather between the threads (all contexts on `device`): every rank copies the published send
        buffers into its own receive buffer with torch (views over the raw pointers), so the single-GPU tests exercise the same
        library path NCCL serves in production (quotient values scattered on the device, column-split FRI reduction)."""
        dev = torch.device("cuda", device)

        def fn(send_ptr, nbytes, recv_ptr):
            self.slots[rank] = (send_ptr, nbytes)
            self.barrier.wait(timeout=600)
            r = torch.as_tensor(_DevView(recv_ptr, nbytes * self.world), device=dev)
            for k, (p, n) in enumerate(self.slots):
                assert n == nbytes
                r[k * nbytes:(k + 1) * nbytes].copy_(torch.as_tensor(_DevView(p, n), device=dev))
            torch.cuda.synchronize(dev)
            self.barrier.wait(timeout=600)
        return fn

    def allgather(self, rank):
        def fn(data):
            self.slots[rank] = data
            self.barrier.wait(timeout=600)
            parts = list(self.slots)
            self.barrier.wait(timeout=600)   # nobody overwrites a slot before everyone has read it
            return parts
        return fn


def lde_class_of_rank(rank, world):
    """The LDE class (natural LDE index mod world) rank `rank` commits to: bit-reversal of the rank, because plonky2 stores
    leaves in bit-reversed order and a rank owns a contiguous run of Merkle cap entries (prover.cu `Shard::rho`)."""
    m = world.bit_length() - 1
    return int(format(rank, "0%db" % m)[::-1], 2) if m else 0


def owner_of_leaf(leaf_index, log_lde_size, world):
    """Rank whose cap subtrees contain leaf `leaf_index` (a FRI query index) and the leaf's index inside that rank's class."""
    m = world.bit_length() - 1
    return leaf_index >> (log_lde_size - m), leaf_index & ((1 << (log_lde_size - m)) - 1)
