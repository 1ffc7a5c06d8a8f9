"""ctypes wrapper over oracle/liboracle.so (TEST INFRASTRUCTURE -- may only be imported from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs)."""
import ctypes as C
import json
import os
import subprocess
import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "liboracle.so")

AIR_MODULAR, AIR_FQ_EXP, AIR_G1_EXP, AIR_G2_EXP, AIR_FQ12_EXP, AIR_FQ12_EXP_U64, AIR_G1_MULADD, AIR_FQ12_MUL = range(8)


class Config(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("security_bits", "num_challenges", "rate_bits", "cap_height", "pow_bits",
                                          "fri_arity_bits", "fri_final_poly_bits", "num_query_rounds")] + [("coset_shift", C.c_uint64), ("fri_degree_hack", C.c_uint32), ("reserved", C.c_uint32)]

    @staticmethod
    def standard_fast_config(rate_bits=1, coset_shift=7, fri_degree_hack=0):
        return Config(100, 2, rate_bits, 4, 16, 4, 5, 84, coset_shift, fri_degree_hack, 0)


def build(force=False):
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".cpp", ".hpp", ".inc"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _DIR, "liboracle.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.orc_last_error.restype = C.c_char_p
        L.orc_air_create.restype = C.c_void_p
        L.orc_air_create.argtypes = [C.c_int, C.c_size_t]
        for f in ("orc_air_num_columns", "orc_air_num_public_inputs", "orc_air_num_rows", "orc_air_num_permutation_pairs", "orc_air_io_size", "orc_air_result_words"):
            getattr(L, f).restype = C.c_size_t
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_air_destroy.argtypes = [C.c_void_p]
        L.orc_generate_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.orc_generate_public_inputs.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_prove.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_verify.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_gl_mul.restype = C.c_uint64
        L.orc_gl_mul.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_gl_inv.restype = C.c_uint64
        L.orc_gl_inv.argtypes = [C.c_uint64]
        L.orc_root_of_unity.restype = C.c_uint64
        L.orc_dbg_num_z.restype = C.c_size_t
        L.orc_dbg_challenges.restype = C.c_size_t
        L.orc_dbg_timings.restype = C.c_size_t
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def select_field(coset_shift=7):
    """U1: generator pair used by the stage entry points without a config (fft, commit_columns, root of unity); 7 = default."""
    if lib().orc_select_field(C.c_uint64(coset_shift)) != 0:
        raise RuntimeError(lib().orc_last_error().decode())


def set_threads(n):
    """OpenMP thread count of the oracle (torchrun sets OMP_NUM_THREADS=1 for its workers)."""
    lib().orc_set_threads(C.c_int(int(n)))


def poseidon(state):
    s = np.ascontiguousarray(state, dtype=np.uint64).copy()
    lib().orc_poseidon(_p(s))
    return s


def poseidon_fast(state):
    s = np.ascontiguousarray(state, dtype=np.uint64).copy()
    lib().orc_poseidon_fast(_p(s))
    return s


def hash_or_noop(vals):
    v = np.ascontiguousarray(vals, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_hash_or_noop(_p(v), C.c_size_t(len(v)), _p(out))
    return out


def fft(col, inverse=False):
    v = np.ascontiguousarray(col, dtype=np.uint64).copy()
    lib().orc_fft(_p(v), C.c_int(int(np.log2(len(v)))), C.c_int(1 if inverse else 0))
    return v


def commit_columns(values, rate_bits=1, cap_height=4, want_lde=True):
    """values: (ncols, n) uint64.  Returns (coeffs, lde (natural order) or None, cap (2^cap_height, 4))."""
    values = np.ascontiguousarray(values, dtype=np.uint64)
    ncols, n = values.shape
    coeffs = np.zeros_like(values)
    lde = np.zeros((ncols, n << rate_bits), dtype=np.uint64) if want_lde else None
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    lib().orc_commit_columns(_p(values), C.c_size_t(ncols), C.c_int(int(np.log2(n))), C.c_int(rate_bits), C.c_int(cap_height),
                             _p(coeffs), _p(lde) if want_lde else None, _p(cap))
    return coeffs, lde, cap


def merkle(leaves, cap_height, prove_index=None):
    leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
    n, w = leaves.shape
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    path = None
    if prove_index is not None:
        path = np.zeros((int(np.log2(n)) - cap_height, 4), dtype=np.uint64)
    lib().orc_merkle(_p(leaves), C.c_size_t(n), C.c_size_t(w), C.c_int(cap_height), _p(cap), C.c_size_t(prove_index or 0),
                     _p(path) if path is not None else None)
    return cap, path


def challenger(observed, n_out):
    o = np.ascontiguousarray(observed, dtype=np.uint64)
    out = np.zeros(n_out, dtype=np.uint64)
    lib().orc_challenger(_p(o), C.c_size_t(len(o)), _p(out), C.c_size_t(n_out))
    return out


class Air:
    def __init__(self, air_id, num_io):
        self.id, self.num_io = air_id, num_io
        self.h = lib().orc_air_create(air_id, num_io)
        if not self.h:
            raise RuntimeError(lib().orc_last_error().decode())
        self.num_columns = lib().orc_air_num_columns(self.h)
        self.num_public_inputs = lib().orc_air_num_public_inputs(self.h)
        self.num_rows = lib().orc_air_num_rows(self.h)
        self.num_pairs = lib().orc_air_num_permutation_pairs(self.h)
        self.io_size = lib().orc_air_io_size(self.h)
        self.result_words = lib().orc_air_result_words(self.h)

    def __del__(self):
        try:
            lib().orc_air_destroy(self.h)
        except Exception:
            pass

    def generate_trace(self, ios: bytes):
        """Returns (trace columns (C, N) uint64, per-io chain results (num_io, result_words) uint64)."""
        assert len(ios) == self.io_size * self.num_io
        cols = np.zeros((self.num_columns, self.num_rows), dtype=np.uint64)
        res = np.zeros((self.num_io, max(self.result_words, 1)), dtype=np.uint64)
        buf = C.create_string_buffer(bytes(ios), len(ios))
        rc = lib().orc_generate_trace(self.h, buf, self.num_io, _p(cols), _p(res))
        if rc != 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return cols, res

    def generate_public_inputs(self, ios: bytes):
        pi = np.zeros(max(self.num_public_inputs, 1), dtype=np.uint64)
        buf = C.create_string_buffer(bytes(ios), len(ios))
        rc = lib().orc_generate_public_inputs(self.h, buf, self.num_io, _p(pi))
        if rc != 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return pi[:self.num_public_inputs]

    def prove(self, trace, public_inputs, cfg=None):
        cfg = cfg or Config.standard_fast_config()
        trace = np.ascontiguousarray(trace, dtype=np.uint64)
        pi = np.ascontiguousarray(public_inputs, dtype=np.uint64)
        out = C.c_void_p()
        n = C.c_size_t()
        rc = lib().orc_prove(self.h, _p(trace), trace.shape[1], _p(pi), len(pi), C.byref(cfg), C.byref(out), C.byref(n))
        if rc != 0:
            raise RuntimeError(lib().orc_last_error().decode())
        data = C.string_at(out, n.value)
        lib().orc_free(out)
        return data

    def verify(self, proof: bytes, cfg=None):
        """Returns (ok, reason)."""
        cfg = cfg or Config.standard_fast_config()
        rc = lib().orc_verify(self.h, bytes(proof), len(proof), C.byref(cfg))
        return rc == 0, ("" if rc == 0 else lib().orc_last_error().decode())


def dbg_z_polys(n):
    nz = lib().orc_dbg_num_z()
    out = np.zeros((nz, n), dtype=np.uint64)
    lib().orc_dbg_z_polys(_p(out))
    return out


def dbg_quotient_chunks(n, count=4):
    out = np.zeros((count, n), dtype=np.uint64)
    lib().orc_dbg_quotient_chunks(_p(out))
    return out


def dbg_challenges():
    out = np.zeros(64, dtype=np.uint64)
    k = lib().orc_dbg_challenges(_p(out), C.c_size_t(64))
    return out[:k]


def dbg_timings():
    buf = C.create_string_buffer(4096)
    lib().orc_dbg_timings(buf, C.c_size_t(4096))
    return json.loads(buf.value.decode())


def permuted_cols(inputs, table):
    i = np.ascontiguousarray(inputs, dtype=np.uint64)
    t = np.ascontiguousarray(table, dtype=np.uint64)
    so, pe = np.zeros_like(i), np.zeros_like(i)
    lib().orc_permuted_cols(_p(i), _p(t), C.c_size_t(len(i)), _p(so), _p(pe))
    return so, pe


def eval_constraints(air, lv, nv, pis, alphas, z_last, l_first, l_last):
    lv, nv = np.ascontiguousarray(lv, dtype=np.uint64), np.ascontiguousarray(nv, dtype=np.uint64)
    pis = np.ascontiguousarray(pis, dtype=np.uint64)
    al = np.ascontiguousarray(alphas, dtype=np.uint64)
    out = np.zeros(len(al), dtype=np.uint64)
    cnt = C.c_size_t()
    lib().orc_eval_constraints(C.c_void_p(air.h), _p(lv), _p(nv), _p(pis), C.c_size_t(air.num_public_inputs), _p(al), C.c_size_t(len(al)),
                               C.c_uint64(z_last), C.c_uint64(l_first), C.c_uint64(l_last), _p(out), C.byref(cnt))
    return out, cnt.value


def check_trace(air, trace, pis):
    """Number of (row, constraint) pairs violated by the trace itself, first violation, constraints per row."""
    trace = np.ascontiguousarray(trace, dtype=np.uint64)
    pis = np.ascontiguousarray(pis, dtype=np.uint64)
    fb = (C.c_long * 2)()
    nc = C.c_size_t()
    L = lib()
    L.orc_check_trace.restype = C.c_long
    bad = L.orc_check_trace(C.c_void_p(air.h), _p(trace), C.c_size_t(trace.shape[1]), _p(pis), C.c_size_t(len(pis)), fb, C.byref(nc))
    return bad, (fb[0], fb[1]), nc.value


def time_sample(air, ios, shift=3, cfg=None):
    """Bounded-sample CPU timing; returns dict of phase -> estimated FULL-proof milliseconds (sampled phases scaled by 2^shift)."""
    cfg = cfg or Config.standard_fast_config()
    out = (C.c_double * 7)()
    buf = C.create_string_buffer(bytes(ios), len(ios)) if ios else None
    rc = lib().orc_time_sample(C.c_void_p(air.h), buf, C.c_size_t(air.num_io), C.byref(cfg), C.c_int(shift), out)
    if rc != 0:
        raise RuntimeError(lib().orc_last_error().decode())
    names = ["tracegen", "commit", "zpoly", "quotient", "openings", "reduce", "fri_tail"]
    est = {k: v * ((1 << shift) if k != "fri_tail" else 1) for k, v in zip(names, list(out))}
    return est
