"""World-size-2 `gloo` test of the multi-GPU host logic (proof sharding, timing reduce, digest gather); no GPU."""
import hashlib
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_proof(batch):
    return hashlib.sha512(b"proof %d" % batch).digest() * (1 + batch % 3)


def _worker(rank, world, port, num_batches, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as g
    sbn = g.load_package()
    from starky_bn254_b200 import sharding
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    mine = sharding.assign(num_batches, world, rank)
    local = {b: _fake_proof(b) for b in mine}
    digests = sharding.gather_digests(local, num_batches)
    tmax = sharding.max_over_ranks([10.0 + rank, 5.0 - rank])
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, digests, tmax, sbn.synthetic.g1_exp_ios(1, seed=0x5EED0001 + 1000 * rank)[:8].hex()))


@pytest.mark.parametrize("world,num_batches", [(2, 7), (2, 2)])
def test_round_robin_sharding_and_gather(world, num_batches):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + num_batches
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_batches, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [(hashlib.sha256(_fake_proof(b)).hexdigest(), len(_fake_proof(b))) for b in range(num_batches)]
    covered = []
    for rank, mine, digests, tmax, seed_probe in res:
        assert mine == [b for b in range(num_batches) if b % world == rank]
        assert digests == want                      # every rank sees every proof's digest, in batch order
        assert tmax == [10.0 + world - 1, 5.0]      # slowest rank wins
        covered += mine
    assert sorted(covered) == list(range(num_batches))          # no batch lost or proved twice
    assert len({r[4] for r in res}) == world                    # ranks draw distinct synthetic inputs
