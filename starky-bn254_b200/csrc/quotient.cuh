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

struct QArgs {
  const u64* trace; size_t trace_stride;     // LDE batch lde[col][b][k]; stride between columns
  const u64* zs; size_t zs_stride;           // permutation Z LDE batch
  int logn;
  size_t coset_off[2];                       // offset of the two quotient cosets inside a column
  u64 coset_shift[2];                        // x = coset_shift[bq] * w_N^k
  const u64* wpow;                           // w_N^k
  u64 w_inv;                                 // w_N^-1 (last subgroup element)
  const u64* lagrange; size_t lagrange_stride;  // [2 cols: first,last][b][k]
  const u64* pi;
  u64* acc;                                  // [challenge][bq][k]
  u64 alpha[SBN_MAX_CHALLENGES], alpha_m[SBN_MAX_CHALLENGES];
  int first;                                 // first segment: accumulators start from zero
  // permutation segment
  const u32* perm_lhs; const u32* perm_rhs; const u64* perm_gamma; int perm_batch; int nz;
  u64* scratch;                              // Fq12 product limb polynomials [12*31][2N] (SEG_FQ12_MUL only)
  const u64* pi_lde; int pi_per_chal;        // public-input binding columns on the quotient cosets [col][bq][k] (core segments)
  u64 pi_skip[SBN_MAX_CHALLENGES];           // alpha^((num_io - 1) * io_len)
};

// Evaluates every constraint of `air` (+ the permutation checks) on the size-2N quotient coset and
// returns the 2*num_challenges quotient chunk polynomials in coefficient form: out[(2c+h)*N + j].
void compute_quotient_chunks(sbn_ctx* ctx, const AirDesc& air, const u64* trace_lde, const u64* zs_lde, const PermInstances& perm,
                             const u64* d_public_inputs, const u64* alphas, int num_challenges, int logn, int rate_bits, u64* out_chunks);
