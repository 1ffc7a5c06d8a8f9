"""Seeded synthetic inputs for the BN254 STARK provers: the reference's tests draw
unseeded random inputs (`rand::thread_rng()`, reference src/curves/g1/exp.rs:792-809); here the same
distributions come from SplitMix64 so that the CPU oracle and the CUDA path see identical bytes."""
import struct

BN254_P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & _M64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def below(self, n):
        """Uniform integer in [0, n) by rejection on 256-bit draws."""
        bits = n.bit_length()
        while True:
            v = 0
            for i in range((bits + 63) // 64):
                v |= self.next() << (64 * i)
            v &= (1 << bits) - 1
            if v < n:
                return v


def _le32(v):
    return v.to_bytes(32, "little")


def random_g1(rng):
    """Random affine point on y^2 = x^3 + 3 (cofactor 1, so every curve point is in G1)."""
    while True:
        x = rng.below(BN254_P)
        rhs = (x * x * x + 3) % BN254_P
        y = pow(rhs, (BN254_P + 1) // 4, BN254_P)
        if y * y % BN254_P == rhs and y != 0:
            if rng.next() & 1:
                y = BN254_P - y
            return x, y


def modular_ios(num_rows, seed=0x5EED0005):
    """ModularStark rows: two uniform Fq residues each (reference src/modular/modular.rs:385-390)."""
    rng = SplitMix64(seed)
    return b"".join(_le32(rng.below(BN254_P)) for _ in range(2 * num_rows))


G1_IO_SIZE = 224


def g1_muladd_ios(num_rows, seed=0x5EED0006, distinct=64):
    """G1Stark rows (reference src/curves/g1/muladd.rs:486-492 draws two random points per row): a.x a.y b.x b.y.  Square roots
    in pure Python are slow, so `distinct` random points are drawn once and paired pseudo-randomly (a != b in every row)."""
    rng = SplitMix64(seed)
    pts = [random_g1(rng) for _ in range(distinct)]
    out = bytearray()
    for _ in range(num_rows):
        i = rng.below(distinct)
        j = (i + 1 + rng.below(distinct - 1)) % distinct
        out += _le32(pts[i][0]) + _le32(pts[i][1]) + _le32(pts[j][0]) + _le32(pts[j][1])
    return bytes(out)


def fq12_mul_ios(num_rows, seed=0x5EED0007):
    """Fq12Stark rows (reference src/fields/fq12/mul.rs:381-382): x, y uniform in Fq12, 12 residues each (MyFq12 order)."""
    rng = SplitMix64(seed)
    return b"".join(_le32(rng.below(BN254_P)) for _ in range(24 * num_rows))


def g1_exp_ios(num_io, seed=0x5EED0001):
    """G1ExpIONative records (x, offset, exp_val[8 x u32], output left zero -- filled from the chain
    result by `fill_g1_outputs`); exp_val is a full 256-bit value (reference g1/exp.rs:796)."""
    rng = SplitMix64(seed)
    out = bytearray()
    for _ in range(num_io):
        x = random_g1(rng)
        off = random_g1(rng)
        out += _le32(x[0]) + _le32(x[1]) + _le32(off[0]) + _le32(off[1])
        out += struct.pack("<8I", *[rng.next() & 0xFFFFFFFF for _ in range(8)])
        out += bytes(64)
    return bytes(out)


def fill_g1_outputs(ios, results):
    """Copy per-io chain results (num_io x >=8 u64: x, y) into the io records' output field."""
    b = bytearray(ios)
    n = len(b) // G1_IO_SIZE
    for i in range(n):
        b[i * G1_IO_SIZE + 160:i * G1_IO_SIZE + 224] = bytes(results[i][:8].tobytes())
    return bytes(b)


# -------- plain big-int BN254 G1 arithmetic (used by tests to check outputs semantically) --------
def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % BN254_P == 0:
            return None
        lam = 3 * p[0] * p[0] * pow(2 * p[1], -1, BN254_P) % BN254_P
    else:
        lam = (q[1] - p[1]) * pow(q[0] - p[0], -1, BN254_P) % BN254_P
    x = (lam * lam - p[0] - q[0]) % BN254_P
    return x, (lam * (p[0] - x) - p[1]) % BN254_P


def g1_mul(p, e):
    r = None
    while e:
        if e & 1:
            r = g1_add(r, p)
        p = g1_add(p, p)
        e >>= 1
    return r


def fill_outputs(ios, results, io_size, out_off):
    """Generic form of fill_g1_outputs: copy each io's chain result into the record's output field."""
    b = bytearray(ios)
    n = len(b) // io_size
    w = (io_size - out_off) // 8
    for i in range(n):
        b[i * io_size + out_off:(i + 1) * io_size] = bytes(results[i][:w].tobytes())
    return bytes(b)


# ---------------- Fq (reference src/fields/fq/exp.rs:88-93 `FqExpIONative`) ----------------
FQ_IO_SIZE = 128


def fq_exp_ios(num_io, seed=0x5EED0000):
    """x, offset uniform non-zero residues; exp_val a full 256-bit value (reference fq/exp.rs:575-590)."""
    rng = SplitMix64(seed)
    out = bytearray()
    for _ in range(num_io):
        out += _le32(1 + rng.below(BN254_P - 1)) + _le32(1 + rng.below(BN254_P - 1))
        out += struct.pack("<8I", *[rng.next() & 0xFFFFFFFF for _ in range(8)])
        out += bytes(32)
    return bytes(out)


# ---------------- Fq2 / G2 (reference src/curves/g2/exp.rs:90-95 `G2ExpIONative`) ----------------
def fq2_add(a, b):
    return (a[0] + b[0]) % BN254_P, (a[1] + b[1]) % BN254_P


def fq2_sub(a, b):
    return (a[0] - b[0]) % BN254_P, (a[1] - b[1]) % BN254_P


def fq2_mul(a, b):
    return (a[0] * b[0] - a[1] * b[1]) % BN254_P, (a[0] * b[1] + a[1] * b[0]) % BN254_P


def fq2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, BN254_P)
    return a[0] * n % BN254_P, -a[1] * n % BN254_P


def _fq_sqrt(a):
    r = pow(a, (BN254_P + 1) // 4, BN254_P)
    return r if r * r % BN254_P == a % BN254_P else None


def fq2_sqrt(a):
    if a[1] == 0:
        r = _fq_sqrt(a[0])
        if r is not None:
            return r, 0
        r = _fq_sqrt(-a[0] % BN254_P)   # sqrt(-c) * u
        return (0, r) if r is not None else None
    s = _fq_sqrt((a[0] * a[0] + a[1] * a[1]) % BN254_P)
    if s is None:
        return None
    inv2 = pow(2, -1, BN254_P)
    for t in ((a[0] + s) * inv2 % BN254_P, (a[0] - s) * inv2 % BN254_P):
        x0 = _fq_sqrt(t)
        if x0:
            x1 = a[1] * pow(2 * x0, -1, BN254_P) % BN254_P
            if fq2_mul((x0, x1), (x0, x1)) == (a[0] % BN254_P, a[1] % BN254_P):
                return x0, x1
    return None


G2_B = fq2_mul((3, 0), fq2_inv((9, 1)))   # twist y^2 = x^3 + 3/(9+u)


def random_g2(rng):
    """Random affine point on the twist (need not be in the r-torsion: the AIR only needs curve points)."""
    while True:
        x = (rng.below(BN254_P), rng.below(BN254_P))
        rhs = fq2_add(fq2_mul(fq2_mul(x, x), x), G2_B)
        y = fq2_sqrt(rhs)
        if y is not None and y != (0, 0):
            if rng.next() & 1:
                y = (-y[0] % BN254_P, -y[1] % BN254_P)
            return x, y


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if fq2_add(p[1], q[1]) == (0, 0):
            return None
        lam = fq2_mul(fq2_mul((3, 0), fq2_mul(p[0], p[0])), fq2_inv(fq2_mul((2, 0), p[1])))
    else:
        lam = fq2_mul(fq2_sub(q[1], p[1]), fq2_inv(fq2_sub(q[0], p[0])))
    x = fq2_sub(fq2_sub(fq2_mul(lam, lam), p[0]), q[0])
    return x, fq2_sub(fq2_mul(lam, fq2_sub(p[0], x)), p[1])


def g2_mul(p, e):
    r = None
    while e:
        if e & 1:
            r = g2_add(r, p)
        p = g2_add(p, p)
        e >>= 1
    return r


G2_IO_SIZE = 416


def _g2_bytes(p):
    return _le32(p[0][0]) + _le32(p[0][1]) + _le32(p[1][0]) + _le32(p[1][1])


def g2_exp_ios(num_io, seed=0x5EED0002):
    """G2ExpIONative records: x, offset as (x.c0, x.c1, y.c0, y.c1), exp_val[8 x u32], output left zero."""
    rng = SplitMix64(seed)
    out = bytearray()
    for _ in range(num_io):
        out += _g2_bytes(random_g2(rng)) + _g2_bytes(random_g2(rng))
        out += struct.pack("<8I", *[rng.next() & 0xFFFFFFFF for _ in range(8)])
        out += bytes(128)
    return bytes(out)


# ---------------- Fq12 in the flat MyFq12 basis (reference src/fields/fq12/exp.rs:90-95) ----------------
def fq12_mul(a, b):
    """sum_{i<6} (c[i] + c[i+6] u) w^i with u^2 = -1, w^6 = 9 + u (reference fq12/mul.rs:24-87)."""
    re, im = [0] * 11, [0] * 11
    for i in range(6):
        for j in range(6):
            re[i + j] += a[i] * b[j] - a[i + 6] * b[j + 6]
            im[i + j] += a[i] * b[j + 6] + a[i + 6] * b[j]
    out = [0] * 12
    for i in range(6):
        if i < 5:
            out[i] = (re[i] + 9 * re[i + 6] - im[i + 6]) % BN254_P
            out[i + 6] = (im[i] + re[i + 6] + 9 * im[i + 6]) % BN254_P
        else:
            out[i], out[i + 6] = re[i] % BN254_P, im[i] % BN254_P
    return out


def fq12_pow_mul(x, e, offset):
    acc = list(offset)
    while e:
        if e & 1:
            acc = fq12_mul(acc, x)
        x = fq12_mul(x, x)
        e >>= 1
    return acc


FQ12_IO_SIZE = 1184
FQ12_U64_IO_SIZE = 1160


def fq12_exp_ios(num_io, seed=0x5EED0003):
    """x, offset uniform in Fq12 (12 residues each); exponent uniform below r (`Fr::rand`, reference fq12/exp.rs:649)."""
    rng = SplitMix64(seed)
    out = bytearray()
    for _ in range(num_io):
        for _ in range(24):
            out += _le32(rng.below(BN254_P))
        out += _le32(rng.below(BN254_R))
        out += bytes(384)
    return bytes(out)


def fq12_exp_u64_ios(num_io, seed=0x5EED0004):
    rng = SplitMix64(seed)
    out = bytearray()
    for _ in range(num_io):
        for _ in range(24):
            out += _le32(rng.below(BN254_P))
        out += struct.pack("<Q", rng.below(P_GOLDILOCKS))
        out += bytes(384)
    return bytes(out)


P_GOLDILOCKS = 2**64 - 2**32 + 1
