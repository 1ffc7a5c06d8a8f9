"""GPU tests (-m gpu) of intra-proof sharding (sbn_prove_sharded): `world` cooperating ranks -- here
threads with one sbn context (CUDA stream) each on cuda:0, exchanging through sharding.ThreadGroup -- must return, on every
rank, the byte-identical proof the unsharded prover returns (which the other GPU tests pin to the oracle and the goldens)."""
import hashlib
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sharded(sbn, world, make, rate_bits=1, device_exchange=False):
    """make(ctx) -> (stark, trace, public_inputs); returns the proof bytes of every rank."""
    from starky_bn254_b200 import sharding
    grp = sharding.ThreadGroup(world)
    out, err = [None] * world, []

    def run(rank):
        ctx = None
        try:
            ctx = sbn.Context(0)
            stark, trace, pi = make(ctx)
            cfg = stark.config(); cfg.rate_bits = rate_bits
            agd = grp.allgather_device(rank) if device_exchange else None
            out[rank] = sbn.prove_sharded(stark, cfg, trace, pi, rank, world, grp.allgather(rank), allgather_device=agd).to_bytes()
            trace.free()
        except BaseException as e:
            err.append((rank, e))
            grp.barrier.abort()
        finally:
            if ctx is not None:
                ctx.close()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not err, err
    return out


@pytest.mark.parametrize("world", [2, 4, 8, 16])
def test_modular_sharded_proof_is_byte_identical(ctx, sbn, orc, golden, world):
    n = 512
    ios = sbn.synthetic.modular_ios(n)

    def make(c):
        stark = sbn.ModularStark(n, c)
        return stark, stark.generate_trace(ios), np.zeros(0, dtype=np.uint64)

    stark, trace, pi = make(ctx)
    want = sbn.prove(stark, stark.config(), trace, pi).to_bytes()
    assert hashlib.sha256(want).hexdigest() == golden["modular_512"]["proof_sha256"]
    for rank, got in enumerate(_sharded(sbn, world, make)):
        assert got == want, (world, rank)
    assert orc.Air(orc.AIR_MODULAR, n).verify(want) == (True, "")


@pytest.mark.parametrize("world", [2, 8])
def test_g1_sharded_proof_matches_golden(sbn, golden, world):
    """The G1 n = 128 proof (BASELINE config 2) from 2 and 8 ranks: same bytes as the committed golden of the oracle prover."""
    n = 128
    syn = sbn.synthetic
    ios = syn.g1_exp_ios(n)

    def make(c):
        stark = sbn.G1ExpStark(n, c)
        trace = stark.generate_trace(ios)
        full = syn.fill_g1_outputs(ios, trace.results())
        return stark, trace, stark.generate_public_inputs(full)

    for rank, got in enumerate(_sharded(sbn, world, make)):
        assert len(got) == golden["g1_128"]["proof_len"], (world, rank)
        assert hashlib.sha256(got).hexdigest() == golden["g1_128"]["proof_sha256"], (world, rank)


def test_fq12_sharded_proof_is_byte_identical(ctx, sbn):
    """Fq12 (the wide AIR: Fq12 products through the scratch batch, split range checks, 584 public inputs per instance)."""
    n = 2
    syn = sbn.synthetic
    ios = syn.fq12_exp_ios(n)

    def make(c):
        stark = sbn.Fq12ExpStark(n, c)
        trace = stark.generate_trace(ios)
        full = syn.fill_outputs(ios, trace.results(), stark.io_size, stark.io_size - 8 * stark.result_words)
        return stark, trace, stark.generate_public_inputs(full)

    stark, trace, pi = make(ctx)
    want = sbn.prove(stark, stark.config(), trace, pi).to_bytes()
    for world in (2, 4):
        for rank, got in enumerate(_sharded(sbn, world, make)):
            assert got == want, (world, rank)


@pytest.mark.parametrize("rate_bits,world", [(2, 2), (2, 4), (2, 8), (3, 2), (3, 8), (3, 16)])
def test_modular_sharded_higher_rates(ctx, sbn, rate_bits, world):
    """rate_bits 2 and 3 (BASELINE config 5): a rank's Merkle class is then several whole sub-cosets (world <= 2^rate_bits) or a
    folded one, and its quotient class is evaluated separately from the coefficients; the proof must not change."""
    n = 1024
    ios = sbn.synthetic.modular_ios(n, seed=11)

    def make(c):
        stark = sbn.ModularStark(n, c)
        return stark, stark.generate_trace(ios), np.zeros(0, dtype=np.uint64)

    stark, trace, pi = make(ctx)
    cfg = stark.config(); cfg.rate_bits = rate_bits
    want = sbn.prove(stark, cfg, trace, pi).to_bytes()
    for rank, got in enumerate(_sharded(sbn, world, make, rate_bits)):
        assert got == want, (rate_bits, world, rank)


def test_sharded_argument_errors(ctx, sbn):
    n = 512
    stark = sbn.ModularStark(n, ctx)
    trace = stark.generate_trace(sbn.synthetic.modular_ios(n))
    with pytest.raises(sbn.SbnError, match="bad shard"):
        sbn.prove_sharded(stark, stark.config(), trace, np.zeros(0, dtype=np.uint64), 0, 3, lambda b: [b, b, b])
    # world = 1 is the plain prover (the callback is never used)
    one = sbn.prove_sharded(stark, stark.config(), trace, np.zeros(0, dtype=np.uint64), 0, 1, None).to_bytes()
    assert one == sbn.prove(stark, stark.config(), trace, np.zeros(0, dtype=np.uint64)).to_bytes()


@pytest.mark.parametrize("name,cls_name,n,gen,world", [
    ("g2_128", "G2ExpStark", 128, "g2_exp_ios", 4),
    ("fq_128", "FqExpStark", 128, "fq_exp_ios", 8),
    ("fq12u64_16", "Fq12ExpU64Stark", 16, "fq12_exp_u64_ios", 2),
])
def test_other_exp_airs_sharded_match_golden(sbn, golden, name, cls_name, n, gen, world):
    """Every exponentiation AIR through the sharded prover: same bytes as the committed golden of the oracle prover."""
    syn = sbn.synthetic
    ios = getattr(syn, gen)(n)

    def make(c):
        stark = getattr(sbn, cls_name)(n, c)
        trace = stark.generate_trace(ios)
        full = syn.fill_outputs(ios, trace.results(), stark.io_size, stark.io_size - 8 * stark.result_words)
        return stark, trace, stark.generate_public_inputs(full)

    for rank, got in enumerate(_sharded(sbn, world, make)):
        assert len(got) == golden[name]["proof_len"], (world, rank)
        assert hashlib.sha256(got).hexdigest() == golden[name]["proof_sha256"], (world, rank)


@pytest.mark.parametrize("name,cls_name,gen", [("g1_muladd_512", "G1Stark", "g1_muladd_ios"), ("fq12_mul_512", "Fq12Stark", "fq12_mul_ios")])
def test_gadget_airs_sharded_match_golden(sbn, golden, name, cls_name, gen):
    n = 512
    ios = getattr(sbn.synthetic, gen)(n)

    def make(c):
        stark = getattr(sbn, cls_name)(n, c)
        return stark, stark.generate_trace(ios), np.zeros(0, dtype=np.uint64)

    for rank, got in enumerate(_sharded(sbn, 4, make)):
        assert hashlib.sha256(got).hexdigest() == golden[name]["proof_sha256"], rank


@pytest.mark.parametrize("world,rate_bits", [(2, 1), (8, 1), (4, 2)])
def test_sharded_with_device_exchange(ctx, sbn, world, rate_bits):
    """The device-buffer exchange path (NCCL all_gather on the library's buffers in production; torch copies between the threads'
    buffers here): quotient values scattered by k_scatter_classes, FRI batch reduction split by columns."""
    n = 1024
    ios = sbn.synthetic.modular_ios(n, seed=21)

    def make(c):
        stark = sbn.ModularStark(n, c)
        return stark, stark.generate_trace(ios), np.zeros(0, dtype=np.uint64)

    stark, trace, pi = make(ctx)
    cfg = stark.config(); cfg.rate_bits = rate_bits
    want = sbn.prove(stark, cfg, trace, pi).to_bytes()
    for rank, got in enumerate(_sharded(sbn, world, make, rate_bits, device_exchange=True)):
        assert got == want, (world, rank)


def test_g1_sharded_with_device_exchange_matches_golden(sbn, golden):
    n = 128
    syn = sbn.synthetic
    ios = syn.g1_exp_ios(n)

    def make(c):
        stark = sbn.G1ExpStark(n, c)
        trace = stark.generate_trace(ios)
        return stark, trace, stark.generate_public_inputs(syn.fill_g1_outputs(ios, trace.results()))

    for rank, got in enumerate(_sharded(sbn, 4, make, device_exchange=True)):
        assert hashlib.sha256(got).hexdigest() == golden["g1_128"]["proof_sha256"], rank
