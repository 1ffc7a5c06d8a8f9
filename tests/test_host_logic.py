"""CPU tests of the product's host logic and of its __host__ __device__ per-point / per-row code run
through the tests/emu harness (no GPU): constraint segments vs the oracle consumer at random points,
witness rows vs the oracle trace, the C ABI surface, and loud failure without a device."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**64 - 2**32 + 1
vp = lambda a: a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def emu(entry):
    return C.CDLL(os.path.join(ROOT, "tests", "emu", "libsbn_emu.so"))


def test_abi_exports_every_declared_symbol(sbn):
    hdr = open(os.path.join(ROOT, "include", "starky_bn254_b200.h")).read()
    names = set(re.findall(r"\b(sbn_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    lib = C.CDLL(sbn.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), "library does not export %s" % n


def test_no_cpu_fallback_when_device_missing(sbn):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sbn.SbnError) as e:
        sbn.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_air_info_and_argument_errors(sbn):
    m = sbn.ModularStark(512)
    assert (m.num_columns, m.num_rows, m.num_permutation_pairs, m.io_size) == (812, 512, 444, 64)
    for bad in (0, 100, 384):      # not a power of two / below the table size
        with pytest.raises(sbn.SbnError):
            sbn.ModularStark(bad)
    with pytest.raises(sbn.SbnError):
        sbn.G1ExpStark(64)         # u16 lookup table needs 2^16 rows (reference range_check.rs:26)
    cfg = sbn.StarkConfig.standard_fast_config()
    assert (cfg.security_bits, cfg.num_challenges, cfg.rate_bits, cfg.cap_height, cfg.pow_bits, cfg.fri_arity_bits, cfg.fri_final_poly_bits,
            cfg.num_query_rounds, cfg.coset_shift) == (100, 2, 1, 4, 16, 4, 5, 84, 7)


def test_public_inputs_match_oracle(sbn, orc):
    n = 128
    ios = sbn.synthetic.g1_exp_ios(n)
    res = np.arange(n * 8, dtype=np.uint64).reshape(n, 8)
    ios = sbn.synthetic.fill_g1_outputs(ios, res)
    assert (sbn.G1ExpStark(n).generate_public_inputs(ios) == orc.Air(orc.AIR_G1_EXP, n).generate_public_inputs(ios)).all()


@pytest.mark.parametrize("air_id,num_io", [(0, 512), (2, 128)])
def test_constraint_segments_match_oracle_consumer(emu, orc, air_id, num_io):
    air = orc.Air(air_id, num_io)
    rng = random.Random(11 + air_id)
    rv = lambda n: np.array([rng.randrange(P) for _ in range(n)], dtype=np.uint64)
    for _ in range(3):
        lv, nv, pi, al = rv(air.num_columns), rv(air.num_columns), rv(max(air.num_public_inputs, 1)), rv(2)
        zl, lf, ll = (rng.randrange(P) for _ in range(3))
        want, cnt = orc.eval_constraints(air, lv, nv, pi, al, zl, lf, ll)
        got = np.zeros(2, dtype=np.uint64)
        ncon = C.c_size_t()
        rc = emu.emu_eval_air(C.c_int(air_id), C.c_size_t(num_io), vp(lv), vp(nv), vp(pi), vp(al), C.c_uint64(zl), C.c_uint64(lf), C.c_uint64(ll),
                              vp(got), C.byref(ncon))
        assert rc == 0 and ncon.value == cnt and (got == want).all()


def test_modular_witness_rows_match_oracle(emu, orc, sbn):
    n = 512
    ios = sbn.synthetic.modular_ios(n)
    trace, _ = orc.Air(orc.AIR_MODULAR, n).generate_trace(ios)
    io = np.frombuffer(ios, dtype=np.uint64).reshape(n, 8)
    for r in range(n):
        row = np.zeros(145, dtype=np.uint64)
        emu.emu_modular_row(vp(np.ascontiguousarray(io[r])), vp(row))
        assert (row == trace[:145, r]).all(), r


def test_modular_witness_edge_inputs(emu, orc):
    """0, 1, p-1 operands (quotient 0 / maximal) through both implementations."""
    q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    vals = [0, 1, 2, q - 1, q - 2, (1 << 253), (1 << 128) - 1]
    rows = [(a, b) for a in vals for b in vals]
    while len(rows) < 256:
        rows.append((q - 1, q - 1))
    ios = b"".join(a.to_bytes(32, "little") + b.to_bytes(32, "little") for a, b in rows)
    trace, _ = orc.Air(orc.AIR_MODULAR, 256).generate_trace(ios)
    io = np.frombuffer(ios, dtype=np.uint64).reshape(256, 8)
    for r in range(256):
        row = np.zeros(145, dtype=np.uint64)
        emu.emu_modular_row(vp(np.ascontiguousarray(io[r])), vp(row))
        assert (row == trace[:145, r]).all(), rows[r]


def test_flags_closed_form_matches_sequential_generation(emu, sbn):
    """flags.rs:46-134 generates rows sequentially; the product uses a closed form per row."""
    rng = random.Random(13)
    for e in ([0] * 8, [0xFFFFFFFF] * 8, [rng.getrandbits(32) for _ in range(8)], [1, 0, 0, 0, 0, 0, 0, 0x80000000]):
        # sequential model (the reference's algorithm, in Python)
        lim = list(e)
        bit = lim[0] & 1; lim[0] >>= 1
        rows = [[0, 0, 0, 1, bit, bit] + lim]
        for cur in range(511):
            lv = rows[-1]
            nv = [0] * 14
            nv[2], nv[3] = 1 - lv[2], 1 - lv[3]
            nv[0] = 1 if cur == 510 else 0
            nv[1] = 1 if cur % 64 == 61 else 0
            if lv[2] == 1:
                nv[5] = lv[6] & 1; nv[6] = lv[6] >> 1
            else:
                nv[5] = lv[5]; nv[6] = lv[6]
            if lv[1] == 1:
                for c in range(7, 14):
                    nv[c - 1] = lv[c]
                nv[13] = 0
            else:
                for c in range(7, 14):
                    nv[c] = lv[c]
            nv[4] = nv[5] * nv[3]
            rows.append(nv)
        ea = np.array(e, dtype=np.uint32)
        for r in range(512):
            out = np.zeros(14, dtype=np.uint64)
            emu.emu_flags_row(vp(ea), C.c_int(r), vp(out))
            assert [int(x) for x in out] == rows[r], (e, r)


def test_g1_chain_and_rows_match_bigint_arithmetic(emu, sbn):
    """Jacobian chain + per-row witness against plain big-int BN254 arithmetic (arkworks' role in the reference)."""
    syn = sbn.synthetic
    ios = syn.g1_exp_ios(1, seed=77)
    b = ios[:224]
    x = (int.from_bytes(b[0:32], "little"), int.from_bytes(b[32:64], "little"))
    off = (int.from_bytes(b[64:96], "little"), int.from_bytes(b[96:128], "little"))
    e = int.from_bytes(b[128:160], "little")
    arr = lambda lo, hi, dt=np.uint64: np.frombuffer(b[lo:hi], dtype=dt).copy()
    A = np.zeros((257, 16), dtype=np.uint32); B = np.zeros((257, 16), dtype=np.uint32)
    emu.emu_g1_chain(vp(arr(0, 32)), vp(arr(32, 64)), vp(arr(64, 96)), vp(arr(96, 128)), vp(arr(128, 160, np.uint32)), vp(A), vp(B))
    words = lambda w: sum(int(v) << (32 * i) for i, v in enumerate(w))
    want = syn.g1_add(syn.g1_mul(x, e), off)
    assert (words(B[256][:8]), words(B[256][8:])) == want
    assert (words(A[3][:8]), words(A[3][8:])) == syn.g1_mul(x, 8)
    # one add row and one double row: new point must be the group law's result, limbs canonical
    for op, p1, p2 in ((1, x, off), (2, x, x)):
        row = np.zeros(384, dtype=np.uint64)
        w8 = lambda v: np.array([(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)
        assert emu.emu_g1_row(vp(w8(p1[0])), vp(w8(p1[1])), vp(w8(p2[0])), vp(w8(p2[1])), C.c_int(op), vp(row)) == 1
        lim = lambda c0: sum(int(row[c0 + i]) << (16 * i) for i in range(16))
        assert (lim(64 + 16), lim(64 + 32)) == syn.g1_add(p1, p2)
        assert int(row[:64 + 317].max()) < 65536
    # equal x with different y: the reference panics (division by zero); the product reports it
    neg = (x[0], syn.BN254_P - x[1])
    row = np.zeros(384, dtype=np.uint64)
    assert emu.emu_g1_row(vp(w8(x[0])), vp(w8(x[1])), vp(w8(neg[0])), vp(w8(neg[1])), C.c_int(1), vp(row)) == 0


def test_product_poseidon_host_path_kat(emu):
    st = np.zeros(12, dtype=np.uint64)
    emu.emu_poseidon(vp(st))
    assert int(st[0]) == 0x3c18a9786cb0b359 and int(st[11]) == 0x1792b1c4342109d7
