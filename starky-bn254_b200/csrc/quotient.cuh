#pragma once
#include "air.cuh"
#include "constraints.cuh"

// One permutation-argument instance list per Z polynomial (singleton column pairs only, as in every
// `permutation_pairs()` of the reference: src/utils/range_check.rs:96-113, :230-246).
struct PermInstances {
  int batch_size = 0;           // instances per Z
  std::vector<u32> lhs, rhs;    // [nz * batch_size]
  std::vector<u64> gamma;       // [nz * batch_size]
  std::vector<int> count;       // [nz] (last batch may be short)
  size_t nz() const { return count.size(); }
};

// Where the quotient is evaluated.  Unsharded (m = 0): the whole size-2N coset as two half-cosets (segments) read from the
// LDE batches lde[col][b][k].  Sharded (m >= 1, rate_bits = 1): class sigma of G = 2^m, the points x = s w_2N^(sigma + G t),
// t < 2N / G, one segment; "next" (index + 2 on the 2N coset) is the same class shifted by one point for G = 2 and class
// sigma + 2 (mod G) for G > 2, which the caller evaluates as a second batch.
struct QDomain {
  int m = 0; u32 sigma = 0;
  const u64* trace = nullptr; const u64* trace_next = nullptr;   // class batches [col][2N / G] when m >= 1
  const u64* zs = nullptr; const u64* zs_next = nullptr;
};

struct QArgs {
  const u64* trace; const u64* trace_next; size_t trace_stride;   // local rows / next rows; stride between columns
  const u64* zs; const u64* zs_next; size_t zs_stride;            // permutation Z batch, same layout
  int logn;                                  // log2(trace rows)
  int logm, nseg; size_t npoints;            // nseg segments of 2^logm points; npoints = nseg << logm
  size_t seg_off[2];                         // offset of a segment inside a column
  size_t next_shift;                         // next row = (k + next_shift) mod 2^logm inside the "next" batch
  u64 seg_shift[2];                          // x = seg_shift[sg] * wpow[k]
  const u64* wpow;                           // w_M^k, M = 2^logm
  u64 w_inv;                                 // w_N^-1 (last subgroup element)
  const u64* lagrange; size_t lagrange_stride;  // [2 cols: first,last], column layout as `trace`
  const u64* pi;
  u64* acc;                                  // [challenge][point]
  u64 alpha[SBN_MAX_CHALLENGES], alpha_m[SBN_MAX_CHALLENGES];
  int first;                                 // first segment: accumulators start from zero
  // permutation segment
  const u32* perm_lhs; const u32* perm_rhs; const u64* perm_gamma; int perm_batch; int nz;
  u64* scratch;                              // Fq12 product limb polynomials [12*31][npoints] (SEG_FQ12_MUL only)
  const u64* pi_lde; int pi_per_chal;        // public-input binding columns on the evaluation points [col][point] (core segments)
  u64 pi_skip[SBN_MAX_CHALLENGES];           // alpha^((num_io - 1) * io_len)
};

// Evaluates every constraint of `air` (+ the permutation checks) on the size-2N quotient coset and
// returns the 2*num_challenges quotient chunk polynomials in coefficient form: out[(2c+h)*N + j].
void compute_quotient_chunks(sbn_ctx* ctx, const AirDesc& air, const u64* trace_lde, const u64* zs_lde, const PermInstances& perm,
                             const u64* d_public_inputs, const u64* alphas, int num_challenges, int logn, int rate_bits, u64* out_chunks);
// The two halves of compute_quotient_chunks.  quotient_eval: constraint accumulators divided by Z_H on the points of `dom`,
// acc[challenge][point] (SBN_MAX_CHALLENGES x npoints, device); quotient_finish: from the full [challenge][bq][k] values to the
// chunk polynomials.  A sharded prover gathers the classes between the two.
size_t quotient_points(const QDomain& dom, int logn);
void quotient_eval(sbn_ctx* ctx, const AirDesc& air, const QDomain& dom, const u64* trace_lde, const u64* zs_lde, const PermInstances& perm,
                   const u64* d_public_inputs, const u64* alphas, int num_challenges, int logn, int rate_bits, u64* d_acc);
void quotient_finish(sbn_ctx* ctx, const u64* d_acc, int num_challenges, int logn, u64* out_chunks);
