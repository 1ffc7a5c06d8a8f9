//! FFI to `libstarkybn254_b200.so` (include/starky_bn254_b200.h): the B200 replacement of
//! `stark.generate_trace` / `generate_public_inputs` / `starky::prover::prove` for the AIRs of qope/starky-bn254
//! (reference call sites: src/curves/g1/exp.rs:816-826, src/curves/g1/circuit.rs:187-201 and the 17 others listed in
//! SURVEY.md section 8b).  Written against the C header; not compiled in the build image (no Rust toolchain).
#![allow(non_camel_case_types)]
use std::ffi::{c_char, c_void, CStr};
use std::ptr;

#[repr(C)] pub struct sbn_ctx { _private: [u8; 0] }
#[repr(C)] pub struct sbn_trace { _private: [u8; 0] }
#[repr(C)] pub struct sbn_proof { _private: [u8; 0] }
#[repr(C)] pub struct sbn_batch { _private: [u8; 0] }
pub const SBN_BATCH_IOS_ON_DEVICE: u32 = 1;
pub const SBN_BATCH_FILL_OUTPUTS: u32 = 2;

/// `StarkConfig` (+ the coset shift, 0 = the field's multiplicative generator).
#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct sbn_config {
    pub security_bits: u32, pub num_challenges: u32, pub rate_bits: u32, pub cap_height: u32, pub pow_bits: u32,
    pub fri_arity_bits: u32, pub fri_final_poly_bits: u32, pub num_query_rounds: u32,
    /// U1: `F::coset_shift()`; the two-adic generator follows as g^((p-1)/2^32).  0 / 7 = (7, 1753635133440165772); 14293326489335486720 = the other candidate pair.
    pub coset_shift: u64,
    /// U3: 1 = FRI polynomial multiplied by X (the older "max-degree hack"); 0 = padded batch quotients.
    pub fri_degree_hack: u32, pub reserved: u32,
}
/// AIR identifiers (a generic `Stark` callback cannot cross to CUDA).
pub const SBN_AIR_MODULAR: i32 = 0;
pub const SBN_AIR_FQ_EXP: i32 = 1;
pub const SBN_AIR_G1_EXP: i32 = 2;
pub const SBN_AIR_G2_EXP: i32 = 3;
pub const SBN_AIR_FQ12_EXP: i32 = 4;
pub const SBN_AIR_FQ12_EXP_U64: i32 = 5;
pub const SBN_AIR_G1_MULADD: i32 = 6;
pub const SBN_AIR_FQ12_MUL: i32 = 7;

/// `G1ExpIONative` (src/curves/g1/exp.rs:88-93): canonical little-endian residues, `Fq::into_bigint().0`.
#[repr(C)] #[derive(Clone, Copy)]
pub struct sbn_g1_exp_io { pub x_x: [u64; 4], pub x_y: [u64; 4], pub offset_x: [u64; 4], pub offset_y: [u64; 4], pub exp_val: [u32; 8], pub output_x: [u64; 4], pub output_y: [u64; 4] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct sbn_fq_exp_io { pub x: [u64; 4], pub offset: [u64; 4], pub exp_val: [u32; 8], pub output: [u64; 4] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct sbn_g2_exp_io { pub x: [u64; 16], pub offset: [u64; 16], pub exp_val: [u32; 8], pub output: [u64; 16] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct sbn_fq12_exp_io { pub x: [u64; 48], pub offset: [u64; 48], pub exp_val: [u32; 8], pub output: [u64; 48] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct sbn_fq12_exp_u64_io { pub x: [u64; 48], pub offset: [u64; 48], pub exp_val: u64, pub output: [u64; 48] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct sbn_modular_io { pub input0: [u64; 4], pub input1: [u64; 4] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct sbn_g1_muladd_io { pub a_x: [u64; 4], pub a_y: [u64; 4], pub b_x: [u64; 4], pub b_y: [u64; 4] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct sbn_fq12_mul_io { pub x: [u64; 48], pub y: [u64; 48] }

pub type sbn_allgather_fn = Option<unsafe extern "C" fn(user: *mut c_void, send: *const c_void, nbytes: usize, recv: *mut c_void) -> i32>;
/// One rank of an intra-proof sharding group (`sbn_prove_sharded`).
#[repr(C)]
pub struct sbn_shard { pub rank: u32, pub world: u32, pub allgather: sbn_allgather_fn, pub user: *mut c_void, pub allgather_device: sbn_allgather_fn }

extern "C" {
    pub fn sbn_ctx_create(device: i32, cuda_stream: *mut c_void, out: *mut *mut sbn_ctx) -> i32;
    pub fn sbn_ctx_destroy(ctx: *mut sbn_ctx);
    pub fn sbn_last_error(ctx: *const sbn_ctx) -> *const c_char;
    pub fn sbn_ctx_synchronize(ctx: *mut sbn_ctx) -> i32;
    pub fn sbn_ctx_select_field(ctx: *mut sbn_ctx, coset_shift: u64) -> i32;
    pub fn sbn_ctx_launch_count(ctx: *const sbn_ctx) -> u64;
    pub fn sbn_ctx_device_bytes(ctx: *const sbn_ctx) -> u64;
    pub fn sbn_ctx_trim(ctx: *mut sbn_ctx) -> i32;
    pub fn sbn_ctx_kernel_timing(ctx: *mut sbn_ctx, enable: i32) -> i32;
    pub fn sbn_ctx_kernel_stats(ctx: *mut sbn_ctx, buf: *mut c_char, cap: usize) -> i32;
    pub fn sbn_config_standard_fast(out: *mut sbn_config) -> i32;
    pub fn sbn_air_info(air: i32, num_io: usize, num_columns: *mut usize, num_public_inputs: *mut usize, num_rows: *mut usize, io_size: *mut usize,
                        result_words: *mut usize, num_permutation_pairs: *mut usize) -> i32;
    pub fn sbn_trace_generate(ctx: *mut sbn_ctx, air: i32, ios: *const c_void, num_io: usize, out: *mut *mut sbn_trace) -> i32;
    pub fn sbn_trace_generate_device(ctx: *mut sbn_ctx, air: i32, d_ios: *const c_void, num_io: usize, out: *mut *mut sbn_trace) -> i32;
    pub fn sbn_trace_upload(ctx: *mut sbn_ctx, air: i32, num_io: usize, cols: *const u64, ncols: usize, nrows: usize, out: *mut *mut sbn_trace) -> i32;
    pub fn sbn_trace_download(trace: *const sbn_trace, cols_out: *mut u64) -> i32;
    pub fn sbn_trace_results(trace: *const sbn_trace, out: *mut u64) -> i32;
    pub fn sbn_trace_free(trace: *mut sbn_trace);
    pub fn sbn_public_inputs(air: i32, ios: *const c_void, num_io: usize, out: *mut u64, out_len: usize) -> i32;
    pub fn sbn_prove(ctx: *mut sbn_ctx, config: *const sbn_config, trace: *const sbn_trace, public_inputs: *const u64, num_public_inputs: usize,
                     out: *mut *mut sbn_proof) -> i32;
    pub fn sbn_prove_sharded(ctx: *mut sbn_ctx, config: *const sbn_config, trace: *const sbn_trace, public_inputs: *const u64, num_public_inputs: usize,
                             shard: *const sbn_shard, out: *mut *mut sbn_proof) -> i32;
    pub fn sbn_batch_create(device: i32, lanes: u32, out: *mut *mut sbn_batch) -> i32;
    pub fn sbn_batch_destroy(batch: *mut sbn_batch);
    pub fn sbn_batch_last_error(batch: *const sbn_batch) -> *const c_char;
    pub fn sbn_prove_batch(batch: *mut sbn_batch, air: i32, num_io: usize, config: *const sbn_config, ios: *const *const c_void, count: usize, flags: u32,
                           proofs_out: *mut *mut sbn_proof) -> i32;
    pub fn sbn_batch_launch_count(batch: *const sbn_batch) -> u64;
    pub fn sbn_batch_device_bytes(batch: *const sbn_batch) -> u64;
    pub fn sbn_batch_trim(batch: *mut sbn_batch) -> i32;
    pub fn sbn_proof_serialize(proof: *const sbn_proof, buf: *mut u8, len: *mut usize) -> i32;
    pub fn sbn_proof_timings(proof: *const sbn_proof, buf: *mut c_char, cap: usize) -> i32;
    pub fn sbn_proof_debug(proof: *const sbn_proof, which: i32, out: *mut u64, cap_words: usize, written: *mut usize) -> i32;
    pub fn sbn_proof_free(proof: *mut sbn_proof);
    pub fn sbn_poseidon_permute(ctx: *mut sbn_ctx, states: *mut u64, n: usize) -> i32;
    pub fn sbn_commit_columns(ctx: *mut sbn_ctx, values: *const u64, ncols: usize, logn: i32, rate_bits: i32, cap_height: i32,
                              coeffs_out: *mut u64, lde_out: *mut u64, cap_out: *mut u64) -> i32;
    pub fn sbn_bench_commit(ctx: *mut sbn_ctx, ncols: usize, logn: i32, rate_bits: i32, cap_height: i32, iters: i32, ms: *mut f32) -> i32;
}

/// Error of a failed call: the negative return code and `sbn_last_error`.
#[derive(Debug)]
pub struct SbnError { pub code: i32, pub message: String }
impl std::fmt::Display for SbnError { fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "[sbn error {}] {}", self.code, self.message) } }
impl std::error::Error for SbnError {}

/// One context per (thread, GPU): a CUDA stream and a caching device allocator.  Thread-compatible, not thread-safe.
pub struct Context { raw: *mut sbn_ctx }
impl Context {
    pub fn new(device: i32) -> Result<Self, SbnError> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sbn_ctx_create(device, ptr::null_mut(), &mut raw) };
        if rc != 0 { return Err(SbnError { code: rc, message: unsafe { last_error(ptr::null()) } }); }
        Ok(Context { raw })
    }
    fn check(&self, rc: i32) -> Result<(), SbnError> { if rc == 0 { Ok(()) } else { Err(SbnError { code: rc, message: unsafe { last_error(self.raw) } }) } }
    /// `stark.generate_trace(&inputs)` on the GPU; `ios` = the AIR's packed input records (`sbn_*_io`).
    /// The record type must be the AIR's: its size is checked against `sbn_air_info` (the C side reads `io_size * num_io` bytes).
    pub fn generate_trace<T: Copy>(&self, air: i32, ios: &[T]) -> Result<Trace<'_>, SbnError> {
        let info = air_info(air, ios.len())?;
        if std::mem::size_of::<T>() != info.io_size { return Err(SbnError { code: -1, message: format!("record type is {} bytes, the AIR's input record is {}", std::mem::size_of::<T>(), info.io_size) }); }
        let mut t = ptr::null_mut();
        self.check(unsafe { sbn_trace_generate(self.raw, air, ios.as_ptr().cast(), ios.len(), &mut t) })?;
        Ok(Trace { ctx: self, raw: t, result_words: info.result_words * ios.len() })
    }
    /// A trace made on the host (`Vec<PolynomialValues<F>>` flattened column-major, canonical u64 values).
    pub fn upload_trace(&self, air: i32, num_io: usize, cols: &[u64], ncols: usize, nrows: usize) -> Result<Trace<'_>, SbnError> {
        assert_eq!(cols.len(), ncols * nrows);
        let mut t = ptr::null_mut();
        self.check(unsafe { sbn_trace_upload(self.raw, air, num_io, cols.as_ptr(), ncols, nrows, &mut t) })?;
        Ok(Trace { ctx: self, raw: t, result_words: 0 })
    }
    /// `starky::prover::prove`: returns the proof in the canonical wire format.
    pub fn prove(&self, config: &sbn_config, trace: &Trace<'_>, public_inputs: &[u64]) -> Result<Vec<u8>, SbnError> {
        let mut p = ptr::null_mut();
        self.check(unsafe { sbn_prove(self.raw, config, trace.raw, public_inputs.as_ptr(), public_inputs.len(), &mut p) })?;
        let mut len = 0usize;
        unsafe { sbn_proof_serialize(p, ptr::null_mut(), &mut len) };
        let mut buf = vec![0u8; len];
        let rc = unsafe { sbn_proof_serialize(p, buf.as_mut_ptr(), &mut len) };
        unsafe { sbn_proof_free(p) };
        self.check(rc)?;
        Ok(buf)
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sbn_ctx_destroy(self.raw) } } }

pub struct Trace<'a> { ctx: &'a Context, raw: *mut sbn_trace, result_words: usize }
impl Trace<'_> {
    /// Per-instance chain results (G1: x, y as 8 words each) to fill the `output` field of the input records:
    /// `result_words * num_io` words, the size the C side writes (empty for an uploaded trace).
    pub fn results(&self) -> Result<Vec<u64>, SbnError> {
        let mut out = vec![0u64; self.result_words];
        if self.result_words > 0 { self.ctx.check(unsafe { sbn_trace_results(self.raw, out.as_mut_ptr()) })?; }
        Ok(out)
    }
}
impl Drop for Trace<'_> { fn drop(&mut self) { unsafe { sbn_trace_free(self.raw) } } }

/// `constants(num_io)` of an AIR (`sbn_air_info`).
#[derive(Clone, Copy, Debug, Default)]
pub struct AirInfo { pub num_columns: usize, pub num_public_inputs: usize, pub num_rows: usize, pub io_size: usize, pub result_words: usize, pub num_permutation_pairs: usize }
pub fn air_info(air: i32, num_io: usize) -> Result<AirInfo, SbnError> {
    let mut i = AirInfo::default();
    let rc = unsafe { sbn_air_info(air, num_io, &mut i.num_columns, &mut i.num_public_inputs, &mut i.num_rows, &mut i.io_size, &mut i.result_words, &mut i.num_permutation_pairs) };
    if rc != 0 { return Err(SbnError { code: rc, message: unsafe { last_error(ptr::null()) } }); }
    Ok(i)
}
/// `stark.generate_public_inputs(&inputs)`; the record type is checked like in `Context::generate_trace`.
pub fn public_inputs<T: Copy>(air: i32, ios: &[T]) -> Result<Vec<u64>, SbnError> {
    let info = air_info(air, ios.len())?;
    if std::mem::size_of::<T>() != info.io_size { return Err(SbnError { code: -1, message: "record type does not match the AIR".into() }); }
    let mut out = vec![0u64; info.num_public_inputs];
    let rc = unsafe { sbn_public_inputs(air, ios.as_ptr().cast(), ios.len(), out.as_mut_ptr(), out.len()) };
    if rc != 0 { return Err(SbnError { code: rc, message: unsafe { last_error(ptr::null()) } }); }
    Ok(out)
}

/// `lanes` worker contexts on one GPU sharing one set of device tables (`sbn_batch`): B independent proofs in one call.
pub struct Batch { raw: *mut sbn_batch }
unsafe impl Send for Batch {}
impl Batch {
    pub fn new(device: i32, lanes: u32) -> Result<Self, SbnError> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sbn_batch_create(device, lanes, &mut raw) };
        if rc != 0 { return Err(SbnError { code: rc, message: unsafe { CStr::from_ptr(sbn_batch_last_error(ptr::null())).to_string_lossy().into_owned() } }); }
        Ok(Batch { raw })
    }
    /// Trace generation + public inputs + prove for every element of `inputs` (each: the `num_io` records of one proof).
    /// `fill_outputs`: take every record's `output` from the trace's chain result.  Returns the serialized proofs in order.
    pub fn prove<T: Copy>(&self, air: i32, config: &sbn_config, inputs: &[&[T]], fill_outputs: bool) -> Result<Vec<Vec<u8>>, SbnError> {
        if inputs.is_empty() { return Ok(vec![]); }
        let num_io = inputs[0].len();
        let info = air_info(air, num_io)?;
        if std::mem::size_of::<T>() != info.io_size || inputs.iter().any(|x| x.len() != num_io) { return Err(SbnError { code: -1, message: "record type / batch shape does not match the AIR".into() }); }
        let ptrs: Vec<*const c_void> = inputs.iter().map(|x| x.as_ptr().cast()).collect();
        let mut proofs: Vec<*mut sbn_proof> = vec![ptr::null_mut(); inputs.len()];
        let rc = unsafe { sbn_prove_batch(self.raw, air, num_io, config, ptrs.as_ptr(), ptrs.len(), if fill_outputs { SBN_BATCH_FILL_OUTPUTS } else { 0 }, proofs.as_mut_ptr()) };
        if rc != 0 { return Err(SbnError { code: rc, message: unsafe { CStr::from_ptr(sbn_batch_last_error(self.raw)).to_string_lossy().into_owned() } }); }
        Ok(proofs.into_iter().map(|p| unsafe { let mut len = 0usize; sbn_proof_serialize(p, ptr::null_mut(), &mut len); let mut b = vec![0u8; len]; sbn_proof_serialize(p, b.as_mut_ptr(), &mut len); sbn_proof_free(p); b }).collect())
    }
}
impl Drop for Batch { fn drop(&mut self) { unsafe { sbn_batch_destroy(self.raw) } } }

unsafe fn last_error(ctx: *const sbn_ctx) -> String { CStr::from_ptr(sbn_last_error(ctx)).to_string_lossy().into_owned() }
pub fn standard_fast_config() -> sbn_config { let mut c = sbn_config::default(); unsafe { sbn_config_standard_fast(&mut c) }; c }

// ---- proof wire format: F = canonical LE u64, extension = (a0, a1), Vec = u32 LE length prefix, Option = u8 tag ----
pub type Hash = [u64; 4];
pub type Ext = [u64; 2];
#[derive(Debug, Default)] pub struct MerkleProof { pub siblings: Vec<Hash> }
#[derive(Debug, Default)] pub struct QueryStep { pub evals: Vec<Ext>, pub merkle_proof: MerkleProof }
#[derive(Debug, Default)] pub struct QueryRound { pub initial: Vec<(Vec<u64>, MerkleProof)>, pub steps: Vec<QueryStep> }
/// Field-for-field `StarkProofWithPublicInputs` (starky 0.1.1): the consumer maps it onto the plonky2 types.
#[derive(Debug, Default)]
pub struct ProofParts {
    pub trace_cap: Vec<Hash>, pub permutation_zs_cap: Option<Vec<Hash>>, pub quotient_polys_cap: Vec<Hash>,
    pub local_values: Vec<Ext>, pub next_values: Vec<Ext>, pub permutation_zs: Option<Vec<Ext>>, pub permutation_zs_next: Option<Vec<Ext>>, pub quotient_polys: Vec<Ext>,
    pub commit_phase_merkle_caps: Vec<Vec<Hash>>, pub query_round_proofs: Vec<QueryRound>, pub final_poly: Vec<Ext>, pub pow_witness: u64,
    pub public_inputs: Vec<u64>,
}
struct Reader<'a> { b: &'a [u8], at: usize }
impl Reader<'_> {
    fn u8(&mut self) -> u8 { let v = self.b[self.at]; self.at += 1; v }
    fn u32(&mut self) -> usize { let v = u32::from_le_bytes(self.b[self.at..self.at + 4].try_into().unwrap()); self.at += 4; v as usize }
    fn f(&mut self) -> u64 { let v = u64::from_le_bytes(self.b[self.at..self.at + 8].try_into().unwrap()); self.at += 8; v }
    fn fvec(&mut self) -> Vec<u64> { let n = self.u32(); (0..n).map(|_| self.f()).collect() }
    fn evec(&mut self) -> Vec<Ext> { let n = self.u32(); (0..n).map(|_| [self.f(), self.f()]).collect() }
    fn hashes(&mut self) -> Vec<Hash> { let n = self.u32(); (0..n).map(|_| [self.f(), self.f(), self.f(), self.f()]).collect() }
}
/// Inverse of `sbn_proof_serialize`.
pub fn decode_proof(bytes: &[u8]) -> ProofParts {
    let mut r = Reader { b: bytes, at: 0 };
    let mut p = ProofParts::default();
    p.trace_cap = r.hashes();
    let has_perm = r.u8() == 1;
    if has_perm { p.permutation_zs_cap = Some(r.hashes()); }
    p.quotient_polys_cap = r.hashes();
    p.local_values = r.evec();
    p.next_values = r.evec();
    if has_perm { p.permutation_zs = Some(r.evec()); p.permutation_zs_next = Some(r.evec()); }
    p.quotient_polys = r.evec();
    let nlayers = r.u32();
    p.commit_phase_merkle_caps = (0..nlayers).map(|_| r.hashes()).collect();
    let nq = r.u32();
    for _ in 0..nq {
        let mut q = QueryRound::default();
        let noracles = r.u32();
        for _ in 0..noracles { let evals = r.fvec(); let siblings = r.hashes(); q.initial.push((evals, MerkleProof { siblings })); }
        let nsteps = r.u32();
        for _ in 0..nsteps { let evals = r.evec(); let siblings = r.hashes(); q.steps.push(QueryStep { evals, merkle_proof: MerkleProof { siblings } }); }
        p.query_round_proofs.push(q);
    }
    p.final_poly = r.evec();
    p.pow_witness = r.f();
    p.public_inputs = r.fvec();
    assert_eq!(r.at, bytes.len(), "trailing bytes in the proof");
    p
}
