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


@pytest.mark.parametrize("air_id,num_io", [(0, 512), (2, 128), (1, 128), (3, 128), (4, 1), (4, 4), (5, 2)])
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


def test_device_poseidon_mds_digit_model_matches_plain_layer():
    """Integer model of the device MDS layer of csrc/poseidon.cuh (16-bit digits, dp2a sums, BIASED constant digits from
    poseidon_rc_dig16.inc, the 10-instruction stitch with 32-bit wrap-around semantics) against sum_i circ[i] s[i+k] + rc:
    the stitch must give the same residue for arbitrary 64-bit representatives, including the extreme digit sums."""
    P = 2**64 - 2**32 + 1
    M32 = 2**32 - 1
    circ = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
    inc = open(os.path.join(ROOT, "starky-bn254_b200", "csrc", "poseidon_rc_dig16.inc")).read()
    dig = [int(x) for x in re.findall(r"\b\d+\b", inc.split("SBN_POSEIDON_RC_DIG16_LIST")[1])]
    assert len(dig) == 372 * 4
    rc_txt = open(os.path.join(ROOT, "starky-bn254_b200", "csrc", "poseidon_rc.inc")).read()
    rc = [int(x, 16) for x in re.findall(r"0x[0-9a-f]{16}", rc_txt)] + [0] * 12
    for i in range(372):   # the biased digits represent the same constants and keep digit 0 above the largest fold term
        d = dig[4 * i:4 * i + 4]
        assert sum(x << (16 * k) for k, x in enumerate(d)) % P == rc[i] and d[0] >= 1024 and max(d) < 65536 + 1024

    def stitch(c):
        assert all(0 <= x < 2**25 for x in c)
        rot = ((c[3] << 16) | (c[3] >> 16)) & M32
        a = c[0] - (c[3] >> 16)
        assert a >= 0
        t1 = (c[1] << 16) & M32
        b = (c[1] >> 16) + c[2]
        s0 = a + t1
        s1 = b + rot + (s0 >> 32)
        e = (-(s1 >> 32)) & M32
        r0 = (s0 & M32) + e
        r1 = (s1 & M32) + (r0 >> 32)
        assert r1 < 2**32
        return (r1 << 32) | (r0 & M32)

    def layer(s, off):
        out = []
        for r in range(12):
            c = list(dig[4 * (off + r):4 * (off + r) + 4])
            for l in range(12):
                coef = circ[(l - r) % 12] + (8 if l == 0 and r == 0 else 0)
                for d in range(4):
                    c[d] += ((s[l] >> (16 * d)) & 0xFFFF) * coef
            out.append(stitch(c))
        return out

    rng = random.Random(11)
    cases = [[rng.getrandbits(64) for _ in range(12)] for _ in range(300)]
    cases += [[2**64 - 1] * 12, [0] * 12, [P - 1] * 12, [0xFFFF_0000_FFFF_FFFF] * 12, [0xFFFF_FFFF_0000_0000] * 12]
    cases += [[rng.choice([0, 2**64 - 1, 0xFFFF, 0xFFFF << 48, P]) for _ in range(12)] for _ in range(300)]
    for s in cases:
        for off in (12, 60, 360):
            want = [(sum(circ[i] * s[(i + k) % 12] for i in range(12)) + (8 * s[0] if k == 0 else 0) + rc[off + k]) % P for k in range(12)]
            got = layer(s, off)
            assert [g % P for g in got] == want and all(g < 2**64 for g in got)


def _w8(v):
    return np.array([(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)


def _ops_from_flags(trace, sf, r, u64_variant=False):
    """0 none, 1 mul/add (filtered_bit), 2 square/double (a) -- EXP_OP_* of witness.cuh."""
    a_col, fbit_col = (sf + 1, sf + 3) if u64_variant else (sf + 2, sf + 4)
    return 2 if trace[a_col, r] == 1 else (1 if trace[fbit_col, r] == 1 else 0)


def test_fq_exp_rows_match_oracle(emu, orc, sbn):
    n = 128
    ios = sbn.synthetic.fq_exp_ios(n, seed=5)
    trace, res = orc.Air(orc.AIR_FQ_EXP, n).generate_trace(ios)
    limbs = lambda r, c0: np.array([int(trace[c0 + 2 * i, r]) | (int(trace[c0 + 2 * i + 1, r]) << 16) for i in range(8)], dtype=np.uint32)
    for r in list(range(0, 40)) + [510, 511, 512, 513, 65535]:
        row = np.zeros(144, dtype=np.uint64)
        emu.emu_fq_exp_row(vp(limbs(r, 0)), vp(limbs(r, 16)), C.c_int(_ops_from_flags(trace, 144, r)), vp(row))
        assert (row == trace[:144, r]).all(), r
    # semantic: chain result = offset * x^e (reference fq/exp.rs:241-244)
    b = ios[:128]
    I = lambda o: int.from_bytes(b[o:o + 32], "little")
    q = sbn.synthetic.BN254_P
    assert int.from_bytes(res[0][:4].tobytes(), "little") == I(32) * pow(I(0), I(64), q) % q


def test_g2_chain_and_rows_match_bigint_arithmetic(emu, sbn):
    syn = sbn.synthetic
    ios = syn.g2_exp_ios(1, seed=78)
    b = ios[:syn.G2_IO_SIZE]
    I = lambda o: int.from_bytes(b[o:o + 32], "little")
    x = ((I(0), I(32)), (I(64), I(96))); off = ((I(128), I(160)), (I(192), I(224))); e = I(256) & ((1 << 40) - 1)
    A = np.zeros((41, 32), dtype=np.uint32); B = np.zeros((41, 32), dtype=np.uint32)
    ew = np.array([(e >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)
    emu.emu_g2_chain(vp(np.frombuffer(b[0:128], dtype=np.uint64).copy()), vp(np.frombuffer(b[128:256], dtype=np.uint64).copy()), vp(ew), C.c_int(40), vp(A), vp(B))
    words = lambda w: sum(int(v) << (32 * i) for i, v in enumerate(w))
    pt = lambda w: ((words(w[0:8]), words(w[8:16])), (words(w[16:24]), words(w[24:32])))
    assert pt(B[40]) == syn.g2_add(syn.g2_mul(x, e), off)
    assert pt(A[5]) == syn.g2_mul(x, 32)
    w16 = lambda c: np.concatenate([_w8(c[0]), _w8(c[1])])
    for op, p1, p2 in ((1, x, off), (2, x, x)):
        row = np.zeros(768, dtype=np.uint64)
        assert emu.emu_g2_row(vp(w16(p1[0])), vp(w16(p1[1])), vp(w16(p2[0])), vp(w16(p2[1])), C.c_int(op), vp(row)) == 1
        lim = lambda c0: sum(int(row[c0 + i]) << (16 * i) for i in range(16))
        got = ((lim(128 + 32), lim(128 + 48)), (lim(128 + 64), lim(128 + 80)))
        assert got == syn.g2_add(p1, p2)
        assert int(row[:128 + 634].max()) < 65536
    neg = (x[0], (-x[1][0] % syn.BN254_P, -x[1][1] % syn.BN254_P))
    row = np.zeros(768, dtype=np.uint64)
    assert emu.emu_g2_row(vp(w16(x[0])), vp(w16(x[1])), vp(w16(neg[0])), vp(w16(neg[1])), C.c_int(1), vp(row)) == 0


@pytest.mark.parametrize("air_id,num_io,rows", [(4, 1, 512), (5, 2, 128)])
def test_fq12_rows_match_oracle(emu, orc, sbn, air_id, num_io, rows):
    syn = sbn.synthetic
    u64v = air_id == 5
    ios = syn.fq12_exp_u64_ios(num_io, seed=9) if u64v else syn.fq12_exp_ios(num_io, seed=9)
    air = orc.Air(air_id, num_io)
    trace, res = air.generate_trace(ios)
    sf = 108 * 16
    words = lambda r, c0: np.array([int(trace[c0 + 2 * i, r]) | (int(trace[c0 + 2 * i + 1, r]) << 16) for i in range(96)], dtype=np.uint32)
    for r in list(range(0, 12)) + [rows - 2, rows - 1, rows, rows + 1]:
        if r >= trace.shape[1]:
            continue
        op = _ops_from_flags(trace, sf, r, u64v)
        a, b, out = words(r, 0), words(r, 192), words(r, 384)
        row = np.zeros(1728, dtype=np.uint64)
        emu.emu_fq12_row(vp(a), vp(b), vp(out), C.c_int(op), vp(row))
        assert (row == trace[:1728, r]).all(), (r, np.nonzero(row != trace[:1728, r])[0][:8])
        if op:
            got = np.zeros(96, dtype=np.uint32)
            emu.emu_fq12_mul(vp(a), vp(a if op == 2 else b), vp(got))
            assert (got == out).all(), r
            got2 = np.zeros(96, dtype=np.uint32)
            emu.emu_fq12_mul_pairwise(vp(a), vp(a if op == 2 else b), vp(got2))
            assert (got2 == out).all(), r
        # flag columns
        if u64v:
            e = int.from_bytes(ios[(r // 128) * syn.FQ12_U64_IO_SIZE + 768:(r // 128) * syn.FQ12_U64_IO_SIZE + 776], "little")
            fl = np.zeros(6, dtype=np.uint64)
            emu.emu_flags_u64_row(C.c_uint64(e), C.c_int(r % 128), vp(fl))
            assert (fl == trace[sf:sf + 6, r]).all(), r
    # semantic check of the chain result (reference fq12/exp.rs:277-279)
    size = syn.FQ12_U64_IO_SIZE if u64v else syn.FQ12_IO_SIZE
    b0 = ios[:size]
    c12 = lambda o: [int.from_bytes(b0[o + 32 * i:o + 32 * i + 32], "little") for i in range(12)]
    e = int.from_bytes(b0[768:776 if u64v else 800], "little")
    want = syn.fq12_pow_mul(c12(0), e, c12(384))
    assert [int.from_bytes(res[0][4 * i:4 * i + 4].tobytes(), "little") for i in range(12)] == want


def test_flags_u64_closed_form_matches_sequential_generation(emu):
    rng = random.Random(17)
    for e in (0, 1, (1 << 64) - (1 << 32), rng.getrandbits(64) % P, 1 << 63):
        bit = e & 1; val = e >> 1
        rows = [[0, 0, 1, bit, bit, val]]
        for cur in range(127):
            lv = rows[-1]
            nv = [1 if cur == 126 else 0, 1 - lv[1], 1 - lv[2], 0, 0, 0]
            if lv[1] == 1:
                nv[4] = lv[5] & 1; nv[5] = lv[5] >> 1
            else:
                nv[4] = lv[4]; nv[5] = lv[5]
            nv[3] = nv[4] * nv[2]
            rows.append(nv)
        for r in range(128):
            out = np.zeros(6, dtype=np.uint64)
            emu.emu_flags_u64_row(C.c_uint64(e), C.c_int(r), vp(out))
            assert [int(x) for x in out] == rows[r], (e, r)


def test_air_info_new_airs(sbn, orc):
    for cls, air_id, n in ((sbn.FqExpStark, 1, 128), (sbn.G2ExpStark, 3, 128), (sbn.Fq12ExpStark, 4, 16), (sbn.Fq12ExpU64Stark, 5, 16)):
        st = cls(n)
        o = orc.Air(air_id, n)
        assert (st.num_columns, st.num_public_inputs, st.num_rows, st.num_permutation_pairs, st.io_size, st.result_words) == \
               (o.num_columns, o.num_public_inputs, o.num_rows, o.num_pairs, o.io_size, o.result_words)
    assert sbn.G2ExpStark(128).num_columns == 2822 and sbn.Fq12ExpStark(16).num_columns == 9802 and sbn.FqExpStark(128).num_columns == 960


def test_public_inputs_new_airs_match_oracle(sbn, orc):
    syn = sbn.synthetic
    for cls, air_id, n, ios in ((sbn.FqExpStark, 1, 128, syn.fq_exp_ios(128)), (sbn.G2ExpStark, 3, 128, syn.g2_exp_ios(128)),
                                (sbn.Fq12ExpStark, 4, 2, syn.fq12_exp_ios(2)), (sbn.Fq12ExpU64Stark, 5, 2, syn.fq12_exp_u64_ios(2))):
        st = cls(n)
        res = (np.arange(n * st.result_words, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)).reshape(n, st.result_words)
        ios2 = syn.fill_outputs(ios, res, st.io_size, st.io_size - 8 * st.result_words)
        assert (st.generate_public_inputs(ios2) == orc.Air(air_id, n).generate_public_inputs(ios2)).all(), air_id


def test_rust_sys_crate_declares_the_whole_c_abi():
    """bindings/rust (not compiled here: no Rust toolchain) must declare every function of the C header, with the same number of
    parameters, and every input record with the header's size."""
    hdr = open(os.path.join(ROOT, "include", "starky_bn254_b200.h")).read()
    rs = open(os.path.join(ROOT, "bindings", "rust", "starky-bn254-b200-sys", "src", "lib.rs")).read()
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    c_funcs = {m.group(1): m.group(2) for m in re.finditer(r"\b(sbn_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr_nc) if "typedef" not in m.group(0)}
    ext = rs[rs.index('extern "C" {'):]
    ext = ext[:ext.index("\n}\n")]
    r_funcs = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (sbn_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->[^;]*)?;", ext, flags=re.S)}
    nargs = lambda s: 0 if s.strip() in ("", "void") else s.count(",") + 1
    assert set(c_funcs) == set(r_funcs), (sorted(set(c_funcs) - set(r_funcs)), sorted(set(r_funcs) - set(c_funcs)))
    for name in c_funcs:
        assert nargs(c_funcs[name]) == nargs(r_funcs[name]), name
    for air, const in re.findall(r"(SBN_AIR_[A-Z0-9_]+) = (\d+)", hdr):
        assert re.search(r"pub const %s: i32 = %s;" % (air, const), rs), air


def test_wire_format_decodes_as_documented(orc, sbn):
    """The proof wire format of DESIGN.md section 7, decoded the way bindings/rust `decode_proof` does it, on an oracle proof:
    every byte is consumed and the shapes are those of StarkProofWithPublicInputs for ModularStark at 512 rows."""
    n = 512
    air = orc.Air(orc.AIR_MODULAR, n)
    trace, _ = air.generate_trace(sbn.synthetic.modular_ios(n))
    b = air.prove(trace, np.zeros(0, dtype=np.uint64))
    at = 0

    def u32():
        nonlocal at
        v = int.from_bytes(b[at:at + 4], "little"); at += 4
        return v

    def words(k):
        nonlocal at
        v = [int.from_bytes(b[at + 8 * i:at + 8 * i + 8], "little") for i in range(k)]; at += 8 * k
        return v
    hashes = lambda: words(4 * u32())
    evec = lambda: words(2 * u32())
    fvec = lambda: words(u32())
    assert len(hashes()) == 64                      # trace cap: 16 digests
    tag = b[at]; at += 1
    assert tag == 1 and len(hashes()) == 64         # Option<permutation_zs_cap> = Some
    assert len(hashes()) == 64                      # quotient cap
    assert [len(evec()) // 2 for _ in range(5)] == [812, 812, 444, 444, 4]   # local, next, zs, zs_next, quotient
    nlayers = u32()
    assert nlayers == 1 and len(hashes()) == 64     # 2^9 rows: one arity-16 reduction
    nq = u32()
    assert nq == 84
    for _ in range(nq):
        assert u32() == 3
        for ncols in (812, 444, 4):
            assert len(fvec()) == ncols and len(hashes()) == 4 * (10 - 4)   # path up to (excluding) the cap level
        assert u32() == nlayers
        assert len(evec()) == 2 * 16 and len(hashes()) == 4 * (6 - 4)
    assert len(evec()) == 2 * 32                    # final polynomial: 2^5 coefficients
    pow_witness = words(1)[0]
    assert pow_witness < P
    assert fvec() == [] and at == len(b)            # no public inputs, nothing left
