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
