/* C ABI of the B200 prover path for qope/starky-bn254.
 *
 * The reference has no FFI: its seam is the Rust generic call
 *     starky::prover::prove::<F, C, S, D>(stark, &config, trace_poly_values, public_inputs, &mut timing)
 * made at reference src/curves/g1/exp.rs:818, src/curves/g1/circuit.rs:192, src/curves/g2/exp.rs:870,
 * src/fields/fq12/exp.rs:671, src/fields/fq/exp.rs:618, src/modular/modular.rs:550 (19 call sites,
 * SURVEY.md section 8b), preceded by `stark.generate_trace(&inputs)` / `stark.generate_public_inputs(&inputs)`
 * (e.g. src/curves/g1/exp.rs:816-817).  A generic callback cannot cross to CUDA, so the replacement is
 * keyed by an AIR identifier.  All functions return 0 on success and a negative code on failure; they
 * never abort or throw across the boundary.  Plain pointers and sizes only.
 */
#ifndef STARKY_BN254_B200_H
#define STARKY_BN254_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Supported envelope (checked, SBN_ERR_INVALID otherwise): constraint degree 3 (quotient degree factor 2, as every AIR of the
 * reference), 1 <= rate_bits <= 4, num_challenges <= 2, trace rows 2^8 .. 2^24 with rows << rate_bits <= 2^30, sharded world <= 16.
 * Handle lifetime: a trace holds buffers of its context; sbn_ctx_destroy with live traces defers the destruction to the last
 * sbn_trace_free.  Proofs are plain host objects and outlive their context. */
typedef struct sbn_ctx sbn_ctx;     /* one per GPU / stream; not thread-safe, thread-compatible */
typedef struct sbn_trace sbn_trace; /* device-resident, column-major trace (Vec<PolynomialValues<F>>) */
typedef struct sbn_proof sbn_proof; /* StarkProofWithPublicInputs in the canonical wire format */

/* mirrors starky::config::StarkConfig (+ FriConfig); see sbn_config_standard_fast */
typedef struct {
  uint32_t security_bits, num_challenges, rate_bits, cap_height, pow_bits, fri_arity_bits, fri_final_poly_bits, num_query_rounds;
  /* U1 -- F::coset_shift() = MULTIPLICATIVE_GROUP_GENERATOR g of plonky2_field's Goldilocks; the two-adic
   * generator follows as POWER_OF_TWO_GENERATOR = g^((p - 1) / 2^32).  0 or 7: pair (B) = (7, 1753635133440165772), the default;
   * 14293326489335486720: pair (A) = (.., 7277203076849721926).  Any generator of F* is accepted. */
  uint64_t coset_shift;
  /* U3 -- 0: batch quotients padded with `quotient.coeffs.push(0)` (default); 1: the older "max-degree hack", the FRI polynomial
   * multiplied by X (prover `final_poly.coeffs.insert(0, 0)`, verifier `sum * subgroup_x`). */
  uint32_t fri_degree_hack;
  uint32_t reserved; /* must be 0 */
} sbn_config;

/* AIR identifiers: the `Stark` implementations of the reference */
enum {
  SBN_AIR_MODULAR = 0,      /* ModularStark        src/modular/modular.rs:361-537 (num_io = rows) */
  SBN_AIR_FQ_EXP = 1,       /* FqExpStark          src/fields/fq/exp.rs */
  SBN_AIR_G1_EXP = 2,       /* G1ExpStark          src/curves/g1/exp.rs:232-742 */
  SBN_AIR_G2_EXP = 3,       /* G2ExpStark          src/curves/g2/exp.rs */
  SBN_AIR_FQ12_EXP = 4,     /* Fq12ExpStark        src/fields/fq12/exp.rs */
  SBN_AIR_FQ12_EXP_U64 = 5, /* Fq12ExpU64Stark     src/fields/fq12_u64/exp_u64.rs */
  SBN_AIR_G1_MULADD = 6,    /* G1Stark (gadget test AIR)   src/curves/g1/muladd.rs:462-624 (num_io = rows, one addition per row) */
  SBN_AIR_FQ12_MUL = 7      /* Fq12Stark (gadget test AIR) src/fields/fq12/mul.rs:355-484  (num_io = rows, one product per row) */
};

#define SBN_OK 0
#define SBN_ERR_INVALID (-1)
#define SBN_ERR_CUDA (-2)
#define SBN_ERR_UNSUPPORTED (-3)
#define SBN_ERR_INTERNAL (-4)

/* Input record of G1ExpStark, replaces `G1ExpIONative` (src/curves/g1/exp.rs:88-93): canonical
 * 256-bit little-endian coordinates, exponent as 8 u32 limbs (least significant first). */
typedef struct {
  uint64_t x_x[4], x_y[4], offset_x[4], offset_y[4];
  uint32_t exp_val[8];
  uint64_t output_x[4], output_y[4];
} sbn_g1_exp_io;
/* Input record of FqExpStark, replaces `FqExpIONative` (src/fields/fq/exp.rs:88-93): output = offset * x^exp_val. */
typedef struct { uint64_t x[4], offset[4]; uint32_t exp_val[8]; uint64_t output[4]; } sbn_fq_exp_io;
/* Input record of G2ExpStark, replaces `G2ExpIONative` (src/curves/g2/exp.rs:90-95).  A point is x.c0, x.c1, y.c0, y.c1
 * (four canonical 256-bit residues, the order of `g2_exp_io_to_columns`, src/curves/g2/exp.rs:139-156). */
typedef struct { uint64_t x[16], offset[16]; uint32_t exp_val[8]; uint64_t output[16]; } sbn_g2_exp_io;
/* Input record of Fq12ExpStark, replaces `Fq12ExpIONative` (src/fields/fq12/exp.rs:90-95).  An Fq12 element is its 12
 * coefficients in the flat `MyFq12` order the reference converts to before writing columns (src/fields/fq12/exp.rs:107-124,
 * src/utils/utils.rs:174-183): element = sum_{i<6} (c[i] + c[i+6] u) w^i, u^2 = -1, w^6 = 9 + u.  output = offset * x^exp_val. */
typedef struct { uint64_t x[48], offset[48]; uint32_t exp_val[8]; uint64_t output[48]; } sbn_fq12_exp_io;
/* Input record of Fq12ExpU64Stark, replaces `Fq12ExpU64IONative` (src/fields/fq12_u64/exp_u64.rs:85-90); exp_val < 2^64 - 2^32 + 1. */
typedef struct { uint64_t x[48], offset[48]; uint64_t exp_val; uint64_t output[48]; } sbn_fq12_exp_u64_io;
/* Input record of G1Stark: one row = the two affine points it adds, a.x != b.x (src/curves/g1/muladd.rs:486-492). */
typedef struct { uint64_t a_x[4], a_y[4], b_x[4], b_y[4]; } sbn_g1_muladd_io;
/* Input record of Fq12Stark: one row = the two Fq12 elements it multiplies, flat MyFq12 order (src/fields/fq12/mul.rs:381-386). */
typedef struct { uint64_t x[48], y[48]; } sbn_fq12_mul_io;
/* Input record of ModularStark: one row = two canonical Fq residues (src/modular/modular.rs:385-390). */
typedef struct { uint64_t input0[4], input1[4]; } sbn_modular_io;

/* `cuda_stream` is a cudaStream_t (may be NULL: the library creates its own stream). */
int sbn_ctx_create(int device, void* cuda_stream, sbn_ctx** out);
void sbn_ctx_destroy(sbn_ctx* ctx);
const char* sbn_last_error(const sbn_ctx* ctx); /* ctx may be NULL: last error of a failed sbn_ctx_create */
int sbn_ctx_synchronize(sbn_ctx* ctx);
/* Generator pair (sbn_config.coset_shift semantics) for the stage entry points that take no config (sbn_commit_columns);
 * sbn_prove / sbn_prove_batch select it from their config. */
int sbn_ctx_select_field(sbn_ctx* ctx, uint64_t coset_shift);
uint64_t sbn_ctx_launch_count(const sbn_ctx* ctx);  /* kernels launched so far through this context */
uint64_t sbn_ctx_device_bytes(const sbn_ctx* ctx);  /* bytes held by the context's caching allocator */
int sbn_ctx_trim(sbn_ctx* ctx);                     /* return the allocator's cached (unused) blocks to the device */
/* Optional CUDA-event timing of the kernel families on the context's stream.  enable != 0 starts (and
 * resets) the collection; sbn_ctx_kernel_stats writes {"family":{"ms":total,"count":launch groups},...}. */
int sbn_ctx_kernel_timing(sbn_ctx* ctx, int enable);
int sbn_ctx_kernel_stats(sbn_ctx* ctx, char* buf, size_t cap);

/* StarkConfig::standard_fast_config (every call site, e.g. src/curves/g1/exp.rs:250-253) */
int sbn_config_standard_fast(sbn_config* out);
/* `constants(num_io)` of each AIR (e.g. src/curves/g1/exp.rs:6-34) */
int sbn_air_info(int air, size_t num_io, size_t* num_columns, size_t* num_public_inputs, size_t* num_rows, size_t* io_size,
                 size_t* result_words, size_t* num_permutation_pairs);

/* K1: `stark.generate_trace(&inputs)` (src/curves/g1/exp.rs:290-318) on the GPU; `ios` is a host array of
 * num_io input records of the AIR's type. */
int sbn_trace_generate(sbn_ctx* ctx, int air, const void* ios, size_t num_io, sbn_trace** out);
/* Same with the input records already resident in device memory (`d_ios` is a device pointer). */
int sbn_trace_generate_device(sbn_ctx* ctx, int air, const void* d_ios, size_t num_io, sbn_trace** out);
/* Host-generated trace path: `cols` is column-major (ncols x nrows), the layout `prove` takes. */
int sbn_trace_upload(sbn_ctx* ctx, int air, size_t num_io, const uint64_t* cols, size_t ncols, size_t nrows, sbn_trace** out);
int sbn_trace_download(const sbn_trace* trace, uint64_t* cols_out);
/* per-io result of the exponentiation chain (`b` on the last row of each block, src/curves/g1/exp.rs:273-281);
 * result_words u64 per io */
int sbn_trace_results(const sbn_trace* trace, uint64_t* out);
void sbn_trace_free(sbn_trace* trace);
/* `stark.generate_public_inputs(&inputs)` (src/curves/g1/exp.rs:320-327); host-side formatting only */
int sbn_public_inputs(int air, const void* ios, size_t num_io, uint64_t* out, size_t out_len);

/* K2-K6: `starky::prover::prove` */
int sbn_prove(sbn_ctx* ctx, const sbn_config* config, const sbn_trace* trace, const uint64_t* public_inputs, size_t num_public_inputs,
              sbn_proof** out);
/* The same proof (byte-identical) computed by `world` = 2, 4, 8 or 16 cooperating contexts, one per GPU: intra-proof sharding of
 * the LDE, the Merkle cap subtrees, the quotient evaluation and the query openings.  Every rank calls
 * this with the same config, the same (replicated) trace and public inputs and its own rank; every rank returns the full proof.
 * `allgather(user, send, nbytes, recv)` must deliver the `nbytes` of every rank, in rank order, into recv[world * nbytes] on
 * every rank (host memory; NCCL / gloo all_gather in the Python mirror) and return 0.  It is called the same number of times
 * with the same sizes on every rank: once per commitment (cap digests), once for the quotient values, once for the opened rows.
 * Needs world <= 2^cap_height (any rate_bits).  world = 1 is sbn_prove. */
typedef int (*sbn_allgather_fn)(void* user, const void* send, size_t nbytes, void* recv);
/* `allgather_device` (optional, may be NULL): the same exchange on DEVICE buffers of this rank's GPU (ncclAllGather), used for
 * the quotient values, the FRI partial sums and the opened rows so that they never pass through the host before the exchange; the library has synchronised its stream before the call and
 * the callback must return only when `recv` is complete.  NULL: the values are staged through host memory and `allgather`. */
typedef struct { uint32_t rank, world; sbn_allgather_fn allgather; void* user; sbn_allgather_fn allgather_device; } sbn_shard;
int sbn_prove_sharded(sbn_ctx* ctx, const sbn_config* config, const sbn_trace* trace, const uint64_t* public_inputs, size_t num_public_inputs,
                      const sbn_shard* shard, sbn_proof** out);
/* Batched proofs ("batches of B independent G1 proofs"): `count` independent proofs of the same
 * AIR in ONE call.  Replaces the loop a caller of the reference writes around
 *     let trace = stark.generate_trace(&inputs); let pi = stark.generate_public_inputs(&inputs); prove(stark, &config, trace, pi, ..)
 * (reference src/curves/g1/exp.rs:816-818, the five `*StarkyProofGenerator::run_once`, e.g. src/curves/g1/circuit.rs:187-201).
 * A batch owns `lanes` worker contexts on one GPU (one CUDA stream, one device allocator and one host thread each) that share
 * one set of read-only device tables; proofs are handed to the lanes as they become free, so the serial sections of one proof
 * (exponentiation chains, lookup walk, host-side Fiat-Shamir) overlap with the wide kernels of the others, and the kernels of
 * small traces (Fq12: 2^14 leaves) run concurrently; fewer lanes run at once when the proofs' footprint (trace + coefficients +
 * LDE) would not fit the device that many times.  Every proof is byte-identical to the one sbn_prove returns for the same
 * inputs.  ios[j] points to the num_io input records of proof j (host memory, or device memory with SBN_BATCH_IOS_ON_DEVICE).
 * SBN_BATCH_FILL_OUTPUTS: the `output` field of every record is taken from the trace (the chain result, sbn_trace_results)
 * instead of being read from the caller's record -- the caller then need not compute the BN254 results natively first.
 * proofs_out[count] receives the proofs (caller frees each with sbn_proof_free); on failure nothing is returned. */
typedef struct sbn_batch sbn_batch;
#define SBN_BATCH_IOS_ON_DEVICE 1u
#define SBN_BATCH_FILL_OUTPUTS 2u
int sbn_batch_create(int device, uint32_t lanes, sbn_batch** out);
void sbn_batch_destroy(sbn_batch* batch);
const char* sbn_batch_last_error(const sbn_batch* batch);
int sbn_prove_batch(sbn_batch* batch, int air, size_t num_io, const sbn_config* config, const void* const* ios, size_t count, uint32_t flags,
                    sbn_proof** proofs_out);
uint64_t sbn_batch_launch_count(const sbn_batch* batch); /* kernels launched so far by all lanes */
uint64_t sbn_batch_device_bytes(const sbn_batch* batch);  /* bytes held by the lanes' allocators */
int sbn_batch_trim(sbn_batch* batch); /* return every lane's cached blocks to the device (e.g. before proving a different AIR) */

/* Canonical little-endian wire format (DESIGN.md "Proof wire format").  Call with buf == NULL to get the length. */
int sbn_proof_serialize(const sbn_proof* proof, uint8_t* buf, size_t* len);
/* JSON object of per-phase device milliseconds of the sbn_prove call that produced `proof` */
int sbn_proof_timings(const sbn_proof* proof, char* buf, size_t cap);
/* Intermediates kept for stage-by-stage parity tests: which = 0 permutation Z columns (nz x N values),
 * 1 quotient chunk coefficients (2*num_challenges x N), 2 challenges (alphas, zeta, fri_alpha, permutation sets). */
int sbn_proof_debug(const sbn_proof* proof, int which, uint64_t* out, size_t cap_words, size_t* written);
void sbn_proof_free(sbn_proof* proof);

/* Stage entry points (parity tests and micro-benchmarks; host buffers in, host buffers out) */
int sbn_poseidon_permute(sbn_ctx* ctx, uint64_t* states, size_t n);  /* n states of 12 u64, in place */
/* PolynomialBatch::from_values: values (ncols x 2^logn, column-major) -> coeffs, LDE (ncols x 2^(logn+rate_bits),
 * natural order i -> shift*w^i) and Merkle cap (2^cap_height x 4).  Output pointers may be NULL. */
int sbn_commit_columns(sbn_ctx* ctx, const uint64_t* values, size_t ncols, int logn, int rate_bits, int cap_height,
                       uint64_t* coeffs_out, uint64_t* lde_out, uint64_t* cap_out);
/* Device-resident micro-benchmark of the commitment kernels on synthetic data: fills ms[0..2] with the
 * average milliseconds of (iNTT+LDE, leaf hashing, upper tree levels) over `iters` runs. */
int sbn_bench_commit(sbn_ctx* ctx, size_t ncols, int logn, int rate_bits, int cap_height, int iters, float* ms);

#ifdef __cplusplus
}
#endif
#endif
