"""Host-side mirror of the reference's prover surface for the B200 path.

The reference is Rust (`stark.generate_trace(&inputs)`, `stark.generate_public_inputs(&inputs)`,
`prove::<F, C, S, D>(stark, &config, trace, public_inputs, &mut timing)` -- reference
src/curves/g1/exp.rs:811-826); no Rust toolchain exists in this image, so this module is the thin
Python binding over the C ABI (`include/starky_bn254_b200.h`) with the same names and argument
meaning.  There is NO CPU fallback: importing works anywhere, but every compute call raises unless
the CUDA library is built and a GPU is present.
"""
import ctypes as C
import json
import os

import numpy as np

from . import synthetic  # noqa: F401  (seeded inputs, shared by tests and bench)

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libstarkybn254_b200.so")

AIR_MODULAR, AIR_FQ_EXP, AIR_G1_EXP, AIR_G2_EXP, AIR_FQ12_EXP, AIR_FQ12_EXP_U64, AIR_G1_MULADD, AIR_FQ12_MUL = range(8)


class SbnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[sbn error {code}] {msg}")
        self.code = code


class StarkConfig(C.Structure):
    """starky::config::StarkConfig (+ FriConfig).  `standard_fast_config` is what every reference call
    site uses (reference src/curves/g1/exp.rs:250-253)."""
    _fields_ = [(n, C.c_uint32) for n in ("security_bits", "num_challenges", "rate_bits", "cap_height", "pow_bits",
                                          "fri_arity_bits", "fri_final_poly_bits", "num_query_rounds")] + [("coset_shift", C.c_uint64), ("fri_degree_hack", C.c_uint32), ("reserved", C.c_uint32)]

    @staticmethod
    def standard_fast_config(num_columns=None, num_public_inputs=None):
        cfg = StarkConfig()
        rc = lib().sbn_config_standard_fast(C.byref(cfg))
        if rc != 0:
            raise SbnError(rc, "sbn_config_standard_fast failed")
        return cfg


_lib = None


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


class Shard(C.Structure):
    """sbn_shard (include/starky_bn254_b200.h): this rank's place in an intra-proof sharding group."""
    _fields_ = [("rank", C.c_uint32), ("world", C.c_uint32), ("allgather", ALLGATHER_FN), ("user", C.c_void_p), ("allgather_device", ALLGATHER_FN)]


def lib():
    """Loads the CUDA library; fails loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, sz, u64p = C.c_void_p, C.c_size_t, C.c_void_p
        L.sbn_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
        L.sbn_ctx_destroy.argtypes = [vp]
        L.sbn_last_error.restype = C.c_char_p
        L.sbn_last_error.argtypes = [vp]
        L.sbn_ctx_synchronize.argtypes = [vp]
        L.sbn_ctx_select_field.argtypes = [vp, C.c_uint64]
        L.sbn_ctx_launch_count.restype = C.c_uint64
        L.sbn_ctx_launch_count.argtypes = [vp]
        L.sbn_ctx_device_bytes.restype = C.c_uint64
        L.sbn_ctx_device_bytes.argtypes = [vp]
        L.sbn_ctx_trim.argtypes = [vp]
        L.sbn_batch_trim.argtypes = [vp]
        L.sbn_ctx_kernel_timing.argtypes = [vp, C.c_int]
        L.sbn_ctx_kernel_stats.argtypes = [vp, C.c_char_p, sz]
        L.sbn_config_standard_fast.argtypes = [vp]
        L.sbn_air_info.argtypes = [C.c_int, sz] + [C.POINTER(sz)] * 6
        L.sbn_trace_generate.argtypes = [vp, C.c_int, vp, sz, C.POINTER(vp)]
        L.sbn_trace_generate_device.argtypes = [vp, C.c_int, vp, sz, C.POINTER(vp)]
        L.sbn_trace_upload.argtypes = [vp, C.c_int, sz, u64p, sz, sz, C.POINTER(vp)]
        L.sbn_trace_download.argtypes = [vp, u64p]
        L.sbn_trace_results.argtypes = [vp, u64p]
        L.sbn_trace_free.argtypes = [vp]
        L.sbn_public_inputs.argtypes = [C.c_int, vp, sz, u64p, sz]
        L.sbn_prove.argtypes = [vp, vp, vp, u64p, sz, C.POINTER(vp)]
        L.sbn_prove_sharded.argtypes = [vp, vp, vp, u64p, sz, C.POINTER(Shard), C.POINTER(vp)]
        L.sbn_batch_create.argtypes = [C.c_int, C.c_uint32, C.POINTER(vp)]
        L.sbn_batch_destroy.argtypes = [vp]
        L.sbn_batch_last_error.restype = C.c_char_p
        L.sbn_batch_last_error.argtypes = [vp]
        L.sbn_prove_batch.argtypes = [vp, C.c_int, sz, vp, C.POINTER(vp), sz, C.c_uint32, C.POINTER(vp)]
        L.sbn_batch_launch_count.restype = C.c_uint64
        L.sbn_batch_launch_count.argtypes = [vp]
        L.sbn_batch_device_bytes.restype = C.c_uint64
        L.sbn_batch_device_bytes.argtypes = [vp]
        L.sbn_proof_serialize.argtypes = [vp, vp, C.POINTER(sz)]
        L.sbn_proof_timings.argtypes = [vp, C.c_char_p, sz]
        L.sbn_proof_debug.argtypes = [vp, C.c_int, u64p, sz, C.POINTER(sz)]
        L.sbn_proof_free.argtypes = [vp]
        L.sbn_poseidon_permute.argtypes = [vp, u64p, sz]
        L.sbn_commit_columns.argtypes = [vp, u64p, sz, C.c_int, C.c_int, C.c_int, u64p, u64p, u64p]
        L.sbn_bench_commit.argtypes = [vp, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One per GPU (sbn_ctx).  `stream` may be a raw cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device=0, stream=None):
        self.h = C.c_void_p()
        rc = lib().sbn_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self.h))
        if rc != 0:
            raise SbnError(rc, lib().sbn_last_error(None).decode())

    def check(self, rc):
        if rc != 0:
            raise SbnError(rc, lib().sbn_last_error(self.h).decode())

    def synchronize(self):
        self.check(lib().sbn_ctx_synchronize(self.h))

    def select_field(self, coset_shift=7):
        """U1: generator pair for the stage entry points without a config (commit_columns); prove() takes it from the config."""
        self.check(lib().sbn_ctx_select_field(self.h, coset_shift))

    @property
    def launch_count(self):
        return int(lib().sbn_ctx_launch_count(self.h))

    @property
    def device_bytes(self):
        return int(lib().sbn_ctx_device_bytes(self.h))

    def trim(self):
        """Return the allocator's cached (unused) device blocks."""
        self.check(lib().sbn_ctx_trim(self.h))

    def kernel_timing(self, enable=True):
        self.check(lib().sbn_ctx_kernel_timing(self.h, 1 if enable else 0))

    def kernel_stats(self):
        buf = C.create_string_buffer(1 << 16)
        self.check(lib().sbn_ctx_kernel_stats(self.h, buf, 1 << 16))
        return json.loads(buf.value.decode() or "{}")

    def close(self):
        if self.h:
            lib().sbn_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stage entry points (parity tests / micro-benchmarks) ----
    def poseidon_permute(self, states):
        s = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, 12).copy()
        self.check(lib().sbn_poseidon_permute(self.h, _ptr(s), s.shape[0]))
        return s

    def commit_columns(self, values, rate_bits=1, cap_height=4, want_lde=True):
        values = np.ascontiguousarray(values, dtype=np.uint64)
        ncols, n = values.shape
        logn = int(n).bit_length() - 1
        coeffs = np.zeros_like(values)
        lde = np.zeros((ncols, n << rate_bits), dtype=np.uint64) if want_lde else None
        cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
        self.check(lib().sbn_commit_columns(self.h, _ptr(values), ncols, logn, rate_bits, cap_height, _ptr(coeffs),
                                            _ptr(lde) if want_lde else None, _ptr(cap)))
        return coeffs, lde, cap

    def bench_commit(self, ncols, logn, rate_bits=1, cap_height=4, iters=3):
        ms = (C.c_float * 3)()
        self.check(lib().sbn_bench_commit(self.h, ncols, logn, rate_bits, cap_height, iters, ms))
        return {"ntt_lde_ms": ms[0], "leaf_hash_ms": ms[1], "tree_ms": ms[2]}


class Trace:
    """Device-resident column-major trace (`Vec<PolynomialValues<F>>` of the reference)."""

    def __init__(self, ctx, handle, stark):
        self.ctx, self.h, self.stark = ctx, handle, stark

    def download(self):
        out = np.zeros((self.stark.num_columns, self.stark.num_rows), dtype=np.uint64)
        self.ctx.check(lib().sbn_trace_download(self.h, _ptr(out)))
        return out

    def results(self):
        out = np.zeros((self.stark.num_io, max(self.stark.result_words, 1)), dtype=np.uint64)
        self.ctx.check(lib().sbn_trace_results(self.h, _ptr(out)))
        return out[:, :self.stark.result_words]

    def free(self):
        if self.h:
            lib().sbn_trace_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class StarkProofWithPublicInputs:
    """Proof in the canonical wire format (DESIGN.md "Proof wire format")."""

    def __init__(self, ctx, handle):
        n = C.c_size_t()
        if lib().sbn_proof_serialize(handle, None, C.byref(n)) != 0:
            raise SbnError(-1, "sbn_proof_serialize failed")
        buf = C.create_string_buffer(n.value)
        if lib().sbn_proof_serialize(handle, buf, C.byref(n)) != 0:
            raise SbnError(-1, "sbn_proof_serialize failed")
        self.bytes = buf.raw[:n.value]
        tb = C.create_string_buffer(4096)
        lib().sbn_proof_timings(handle, tb, 4096)
        self.timings = json.loads(tb.value.decode() or "{}")
        self._debug = {}
        for which, name in ((0, "z_polys"), (1, "quotient_chunks"), (2, "challenges")):
            w = C.c_size_t()
            lib().sbn_proof_debug(handle, which, None, 0, C.byref(w))
            arr = np.zeros(w.value, dtype=np.uint64)
            if w.value:
                lib().sbn_proof_debug(handle, which, _ptr(arr), w.value, C.byref(w))
            self._debug[name] = arr
        lib().sbn_proof_free(handle)

    def to_bytes(self):
        return self.bytes

    def debug(self, name):
        return self._debug[name]


class _Stark:
    AIR = None

    def __init__(self, num_io, ctx=None):
        self.num_io = num_io
        self.ctx = ctx
        v = [C.c_size_t() for _ in range(6)]
        rc = lib().sbn_air_info(self.AIR, num_io, *[C.byref(x) for x in v])
        if rc != 0:
            raise SbnError(rc, lib().sbn_last_error(None).decode())
        (self.num_columns, self.num_public_inputs, self.num_rows, self.io_size, self.result_words, self.num_permutation_pairs) = [x.value for x in v]

    def config(self):
        return StarkConfig.standard_fast_config(self.num_columns, self.num_public_inputs)

    def _ctx(self, ctx):
        c = ctx or self.ctx
        if c is None:
            raise ValueError("a Context is required")
        return c

    def generate_trace(self, inputs: bytes, ctx=None):
        """K1 on the GPU.  `inputs` = num_io packed input records (see include/starky_bn254_b200.h)."""
        c = self._ctx(ctx)
        if len(inputs) != self.io_size * self.num_io:
            raise ValueError("inputs has the wrong length")
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(inputs), len(inputs))
        c.check(lib().sbn_trace_generate(c.h, self.AIR, buf, self.num_io, C.byref(h)))
        return Trace(c, h, self)

    def generate_trace_ptr(self, host_ptr, nbytes, ctx=None):
        """K1 from a raw host pointer (e.g. a pinned torch tensor's data_ptr()); no extra host copy."""
        c = self._ctx(ctx)
        if nbytes != self.io_size * self.num_io:
            raise ValueError("inputs has the wrong length")
        h = C.c_void_p()
        c.check(lib().sbn_trace_generate(c.h, self.AIR, C.c_void_p(host_ptr), self.num_io, C.byref(h)))
        return Trace(c, h, self)

    def generate_trace_device(self, device_ptr, ctx=None):
        """K1 with the input records already resident in HBM (device pointer)."""
        c = self._ctx(ctx)
        h = C.c_void_p()
        c.check(lib().sbn_trace_generate_device(c.h, self.AIR, C.c_void_p(device_ptr), self.num_io, C.byref(h)))
        return Trace(c, h, self)

    def upload_trace(self, cols, ctx=None):
        """Host-generated trace path (column-major (num_columns, num_rows) uint64)."""
        c = self._ctx(ctx)
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        h = C.c_void_p()
        c.check(lib().sbn_trace_upload(c.h, self.AIR, self.num_io, _ptr(cols), cols.shape[0], cols.shape[1], C.byref(h)))
        return Trace(c, h, self)

    def generate_public_inputs(self, inputs: bytes):
        if len(inputs) != self.io_size * self.num_io:   # the C side reads io_size * num_io bytes
            raise ValueError("inputs has the wrong length")
        out = np.zeros(max(self.num_public_inputs, 1), dtype=np.uint64)
        buf = C.create_string_buffer(bytes(inputs), len(inputs))
        rc = lib().sbn_public_inputs(self.AIR, buf, self.num_io, _ptr(out), self.num_public_inputs)
        if rc != 0:
            raise SbnError(rc, lib().sbn_last_error(None).decode())
        return out[:self.num_public_inputs]


class ModularStark(_Stark):
    """reference src/modular/modular.rs:361-537 (rows = num_io)."""
    AIR = AIR_MODULAR


class G1Stark(_Stark):
    """The reference's gadget test AIR, one G1 addition per row: src/curves/g1/muladd.rs:462-624 (rows = num_io;
    inputs: packed sbn_g1_muladd_io records)."""
    AIR = AIR_G1_MULADD


class Fq12Stark(_Stark):
    """The reference's gadget test AIR, one Fq12 product per row: src/fields/fq12/mul.rs:355-484 (rows = num_io;
    inputs: packed sbn_fq12_mul_io records)."""
    AIR = AIR_FQ12_MUL


class G1ExpStark(_Stark):
    """reference src/curves/g1/exp.rs:232-742."""
    AIR = AIR_G1_EXP


class FqExpStark(_Stark):
    """reference src/fields/fq/exp.rs (inputs: packed sbn_fq_exp_io records)."""
    AIR = AIR_FQ_EXP


class G2ExpStark(_Stark):
    """reference src/curves/g2/exp.rs (inputs: packed sbn_g2_exp_io records)."""
    AIR = AIR_G2_EXP


class Fq12ExpStark(_Stark):
    """reference src/fields/fq12/exp.rs (inputs: packed sbn_fq12_exp_io records, MyFq12 coefficient order)."""
    AIR = AIR_FQ12_EXP


class Fq12ExpU64Stark(_Stark):
    """reference src/fields/fq12_u64/exp_u64.rs (inputs: packed sbn_fq12_exp_u64_io records)."""
    AIR = AIR_FQ12_EXP_U64


def prove(stark, config, trace, public_inputs, timing=None):
    """`starky::prover::prove(stark, &config, trace_poly_values, public_inputs, &mut timing)` on the GPU."""
    ctx = trace.ctx
    pi = np.ascontiguousarray(public_inputs, dtype=np.uint64)
    h = C.c_void_p()
    ctx.check(lib().sbn_prove(ctx.h, C.byref(config), trace.h, _ptr(pi), len(pi), C.byref(h)))
    proof = StarkProofWithPublicInputs(ctx, h)
    if timing is not None:
        timing.update(proof.timings)
    return proof


BATCH_IOS_ON_DEVICE, BATCH_FILL_OUTPUTS = 1, 2


class Batch:
    """sbn_batch: `lanes` worker contexts on one GPU (a CUDA stream, a device allocator and a native host thread each) sharing one
    set of device tables.  `prove_batch` runs trace generation + prove for a list of independent input batches in ONE call
    (SURVEY.md 8d config 2); the caller is a single host thread."""

    def __init__(self, device=0, lanes=6):
        self.h = C.c_void_p()
        self.lanes = lanes
        rc = lib().sbn_batch_create(device, lanes, C.byref(self.h))
        if rc != 0:
            raise SbnError(rc, lib().sbn_batch_last_error(None).decode())

    @property
    def launch_count(self):
        return int(lib().sbn_batch_launch_count(self.h))

    @property
    def device_bytes(self):
        return int(lib().sbn_batch_device_bytes(self.h))

    def trim(self):
        """Return every lane's cached device blocks (before switching to an AIR of another size)."""
        lib().sbn_batch_trim(self.h)

    def close(self):
        if self.h:
            lib().sbn_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def prove_batch(stark, config, batch, inputs, on_device=False, fill_outputs=True):
    """`count` independent proofs of `stark` in one call: for every element of `inputs` -- the packed input records of one proof,
    as `bytes`, or a raw pointer (int) to host (pinned) or device memory -- trace generation, public inputs and prove.
    fill_outputs: take each record's `output` field from the trace's chain result (sbn_trace_results) rather than from the record.
    Returns the proofs in input order, each byte-identical to `prove`'s for the same inputs."""
    n = len(inputs)
    keep, ptrs = [], (C.c_void_p * n)()
    for j, x in enumerate(inputs):
        if isinstance(x, int):
            ptrs[j] = x
        else:
            if on_device:
                raise ValueError("device inputs must be raw pointers")
            if len(x) != stark.io_size * stark.num_io:
                raise ValueError("inputs has the wrong length")
            buf = C.create_string_buffer(bytes(x), len(x))
            keep.append(buf)
            ptrs[j] = C.cast(buf, C.c_void_p).value
    out = (C.c_void_p * n)()
    flags = (BATCH_IOS_ON_DEVICE if on_device else 0) | (BATCH_FILL_OUTPUTS if fill_outputs else 0)
    rc = lib().sbn_prove_batch(batch.h, stark.AIR, stark.num_io, C.byref(config), ptrs, n, flags, out)
    if rc != 0:
        raise SbnError(rc, lib().sbn_batch_last_error(batch.h).decode())
    return [StarkProofWithPublicInputs(None, C.c_void_p(h)) for h in out]


def prove_sharded(stark, config, trace, public_inputs, rank, world, allgather, timing=None, allgather_device=None):
    """One proof computed by `world` (2, 4, 8, 16) cooperating ranks, one GPU each: `sbn_prove_sharded`.  Every rank passes the
    same (replicated) trace and public inputs and gets the full proof, byte-identical to `prove`'s.  `allgather(data: bytes)`
    returns the list of every rank's `data` in rank order (see sharding.dist_allgather / sharding.ThreadGroup).
    `allgather_device(send_ptr, nbytes, recv_ptr)` (optional) does the same between device buffers (sharding.dist_allgather_device):
    the quotient values then never pass through the host."""
    ctx = trace.ctx
    pi = np.ascontiguousarray(public_inputs, dtype=np.uint64)
    failure = []

    def cb(user, send, nbytes, recv):
        try:
            if hasattr(allgather, "raw"):   # in place on the library's buffers
                allgather.raw(send, nbytes, recv)
                return 0
            parts = allgather(C.string_at(send, nbytes))
            if len(parts) != world or any(len(x) != nbytes for x in parts):
                raise ValueError("all-gather returned %s parts of sizes %s, expected %d x %d" % (len(parts), [len(x) for x in parts][:4], world, nbytes))
            C.memmove(recv, b"".join(parts), nbytes * world)
            return 0
        except BaseException as e:  # never let an exception cross the C frames
            failure.append(e)
            return 1

    def cb_dev(user, send, nbytes, recv):
        try:
            allgather_device(send, nbytes, recv)
            return 0
        except BaseException as e:
            failure.append(e)
            return 1

    shard = Shard(rank, world, ALLGATHER_FN(cb), None, ALLGATHER_FN(cb_dev) if allgather_device is not None else ALLGATHER_FN())
    h = C.c_void_p()
    rc = lib().sbn_prove_sharded(ctx.h, C.byref(config), trace.h, _ptr(pi), len(pi), C.byref(shard), C.byref(h))
    if failure:
        raise failure[0]
    ctx.check(rc)
    proof = StarkProofWithPublicInputs(ctx, h)
    if timing is not None:
        timing.update(proof.timings)
    return proof
