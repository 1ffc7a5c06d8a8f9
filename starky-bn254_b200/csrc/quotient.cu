// K5: quotient evaluation.  Replaces starky's `compute_quotient_polys` (external dependency;
// SURVEY.md App. B.7) including the call-back into the reference's `eval_packed_generic`
// (e.g. reference src/curves/g1/exp.rs:331-495) and `eval_permutation_checks`.
//
// One thread per point of the size-2N quotient coset; LDE columns are column-major so a warp reads
// 32 consecutive values of each column it touches (the "next" row is the same run shifted by one).
#include "quotient.cuh"
#include "ntt.cuh"

HD void qpoint_begin(const QArgs& a, size_t idx, QPoint& q) {
  const size_t M = size_t(1) << a.logm;
  const int sg = (int)(idx >> a.logm);
  const size_t k = idx & (M - 1), kn = (k + a.next_shift) & (M - 1);
  q.lp = a.trace + a.seg_off[sg] + k;
  q.np = a.trace_next + a.seg_off[sg] + kn;
  q.stride = a.trace_stride;
  q.pi = a.pi;
  F x(gl_mul(a.seg_shift[sg], a.wpow[k]));
  q.z_last = x - F(a.w_inv);
  q.l_first = F(a.lagrange[a.seg_off[sg] + k]);
  q.l_last = F(a.lagrange[a.lagrange_stride + a.seg_off[sg] + k]);
  for (int c = 0; c < SBN_MAX_CHALLENGES; c++) { q.alpha[c] = F(a.alpha[c]); q.acc[c] = F(); q.pi_skip[c] = F(a.pi_skip[c]); }
  q.pic = a.pi_lde + idx; q.pic_stride = a.npoints; q.pic_per_chal = a.pi_per_chal;
}
HD void qpoint_end(const QArgs& a, size_t idx, const QPoint& q) {
  for (int c = 0; c < SBN_MAX_CHALLENGES; c++) {
    u64* p = a.acc + (size_t)c * a.npoints + idx;
    F prev = a.first ? F() : F(*p) * F(a.alpha_m[c]);
    *p = f_canon(prev + q.acc[c]);
  }
}

// starky `eval_permutation_checks`: first-row Z-1 for every Z, then one product
// constraint per Z.  2*nz constraints.
HD void eval_permutation_checks(const QArgs& a, size_t idx, QPoint& q) {
  const size_t M = size_t(1) << a.logm;
  const int sg = (int)(idx >> a.logm);
  const size_t k = idx & (M - 1), kn = (k + a.next_shift) & (M - 1);
  const u64* zl = a.zs + a.seg_off[sg] + k;
  const u64* zn = a.zs_next + a.seg_off[sg] + kn;
  for (int i = 0; i < a.nz; i++) q.first_row(F(zl[(size_t)i * a.zs_stride]) - F(1));
  for (int i = 0; i < a.nz; i++) {
    F lhs(1), rhs(1);
    for (int j = 0; j < a.perm_batch; j++) {
      int e = i * a.perm_batch + j;
      u32 l = a.perm_lhs[e];
      if (l == 0xFFFFFFFFu) break;
      F g(a.perm_gamma[e]);
      lhs = lhs * (q.lv(l) + g);
      rhs = rhs * (q.lv(a.perm_rhs[e]) + g);
    }
    q.constraint(F(zn[(size_t)i * a.zs_stride]) * rhs - F(zl[(size_t)i * a.zs_stride]) * lhs);
  }
}

// runs of the same list over the Z range [z0, z1): run 0 = first-row constraints, run 1 = product constraints
HD void eval_permutation_checks_run(const QArgs& a, size_t idx, QPoint& q, int z0, int z1, int run) {
  const size_t M = size_t(1) << a.logm;
  const int sg = (int)(idx >> a.logm);
  const size_t k = idx & (M - 1), kn = (k + a.next_shift) & (M - 1);
  const u64* zl = a.zs + a.seg_off[sg] + k;
  const u64* zn = a.zs_next + a.seg_off[sg] + kn;
  if (run == 0) {
    for (int i = z0; i < z1; i++) q.first_row(F(zl[(size_t)i * a.zs_stride]) - F(1));
    return;
  }
  for (int i = z0; i < z1; i++) {
    F lhs(1), rhs(1);
    for (int j = 0; j < a.perm_batch; j++) {
      int e = i * a.perm_batch + j;
      u32 l = a.perm_lhs[e];
      if (l == 0xFFFFFFFFu) break;
      F g(a.perm_gamma[e]);
      lhs = lhs * (q.lv(l) + g);
      rhs = rhs * (q.lv(a.perm_rhs[e]) + g);
    }
    q.constraint(F(zn[(size_t)i * a.zs_stride]) * rhs - F(zl[(size_t)i * a.zs_stride]) * lhs);
  }
}

// Small evaluation domains (Fq12 at 2^13 rows, a rank's class of a sharded proof): one thread per point leaves most of the
// GPU idle while each thread walks thousands of items (5 328 Z polynomials for Fq12).  The long item loops of the permutation
// and split-range-check segments are cut into `nchunks` ranges (blockIdx.y); a chunk folds each of its runs with Horner from
// zero and scales it by alpha^(number of constraints that follow the run in the full list), so that the sum over chunks is the
// segment's Horner value S -- the same field element the one-thread-per-point kernel produces.
struct ChunkPlan { int nchunks, nitems, per; };   // chunk j covers items [j per, min(nitems, (j + 1) per))
__global__ void __launch_bounds__(128) k_segment_chunked(QArgs a, Segment s, ChunkPlan plan, const u64* __restrict__ weights /* [chunk][3][chal] */,
                                                         u64* __restrict__ partial /* [chunk][chal][npoints] */) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= a.npoints) return;
  const int ch = blockIdx.y, i0 = ch * plan.per, i1 = min(plan.nitems, i0 + plan.per);
  QPoint q;
  qpoint_begin(a, idx, q);
  F total[SBN_MAX_CHALLENGES];
  for (int c = 0; c < SBN_MAX_CHALLENGES; c++) total[c] = F();
  const int nruns = (s.kind == SEG_SPLIT_RANGE_CHECK && ch == plan.nchunks - 1) ? 3 : 2;
  for (int run = 0; run < nruns; run++) {
    for (int c = 0; c < SBN_MAX_CHALLENGES; c++) q.acc[c] = F();
    if (s.kind == SEG_PERMUTATION) eval_permutation_checks_run(a, idx, q, i0, i1, run);
    else eval_split_u16_range_check_run(q, s.p0, s.p1, i0, i1, run);
    for (int c = 0; c < SBN_MAX_CHALLENGES; c++) total[c] = total[c] + q.acc[c] * F(weights[((size_t)ch * 3 + run) * SBN_MAX_CHALLENGES + c]);
  }
  for (int c = 0; c < SBN_MAX_CHALLENGES; c++) partial[((size_t)ch * SBN_MAX_CHALLENGES + c) * a.npoints + idx] = f_canon(total[c]);
}
__global__ void k_combine_chunks(QArgs a, int nchunks, const u64* __restrict__ partial) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= a.npoints) return;
  for (int c = 0; c < SBN_MAX_CHALLENGES; c++) {
    u64* p = a.acc + (size_t)c * a.npoints + idx;
    u64 v = a.first ? 0 : gl_mul(*p, a.alpha_m[c]);
    for (int ch = 0; ch < nchunks; ch++) v = gl_add(v, partial[((size_t)ch * SBN_MAX_CHALLENGES + c) * a.npoints + idx]);
    *p = v;
  }
}

HD void eval_segment(const QArgs& a, const Segment& s, size_t idx, QPoint& q) {
  switch (s.kind) {
    case SEG_SPLIT_RANGE_CHECK: eval_split_u16_range_check(q, s.p0, s.p1, s.p2); break;
    case SEG_MODULAR_CORE: eval_modular_stark_core(q); break;
    case SEG_G1_CORE: eval_exp_core_u32<2>(q, s.p0, s.p1, s.p2); break;
    case SEG_FLAGS: eval_flags(q, s.p0); break;
    case SEG_G1_ADD: eval_g1_add(q, q.lv(s.p1), s.p0); break;
    case SEG_G1_DOUBLE: eval_g1_double(q, q.lv(s.p1), s.p0); break;
    case SEG_PERIODIC_PULSE: eval_periodic_pulse(q, s.p0, s.p1, s.p2, s.p3); break;
    case SEG_PULSE: eval_pulse(q, s.p0, s.p1, s.p2); break;
    case SEG_U16_RANGE_CHECK: eval_u16_range_check(q, s.p0, s.p1); break;
    case SEG_PERMUTATION: eval_permutation_checks(a, idx, q); break;
    case SEG_FQ_CORE: eval_exp_core_u32<1>(q, s.p0, s.p1, s.p2); break;
    case SEG_G2_CORE: eval_exp_core_u32<4>(q, s.p0, s.p1, s.p2); break;
    case SEG_FQ_MUL: eval_fq_mul(q, q.lv(s.p0), s.p1 != 0); break;
    case SEG_G2_ADD: eval_g2_add(q, q.lv(s.p1), s.p0); break;
    case SEG_G2_DOUBLE: eval_g2_double(q, q.lv(s.p1), s.p0); break;
    case SEG_FQ12_CORE: eval_fq12_exp_core(q, s.p0, s.p1, s.p2 != 0); break;
    case SEG_FQ12_MUL: eval_fq12_mul(q, q.lv(s.p0), a.scratch + idx, a.npoints); break;
    case SEG_FLAGS_U64: eval_flags_u64(q, s.p0); break;
  }
}

// Limb polynomial of pol_mul_fq12(x, y, 9) (reference src/fields/fq12/mul.rs:24-87) at every quotient point, for the
// SEG_FQ12_MUL segment: prod[(oi * 31 + k) * 2N + idx].  One thread per (point, output coefficient oi); the 31
// coefficients are accumulated in registers while the contributing (x_i, y_j) limb pairs stream through.
//   out[i]   = re[i] + 9 re[i+6] - im[i+6],  out[i+6] = im[i] + re[i+6] + 9 im[i+6]  (i < 5);  out[5] = re[5], out[11] = im[5]
//   re[m] = sum_{i+j=m} (x_i y_j - x_{i+6} y_{j+6}),  im[m] = sum_{i+j=m} (x_i y_{j+6} + x_{i+6} y_j)
__global__ void __launch_bounds__(128) k_fq12_products(QArgs a, int xa, int ya, u64* __restrict__ prod) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t N2 = a.npoints;
  if (idx >= N2) return;
  const int oi = blockIdx.y;
  QPoint q;
  qpoint_begin(a, idx, q);
  F acc[31];
  fq12_product_acc(q, xa, ya, oi, acc);
#pragma unroll
  for (int k = 0; k < 31; k++) prod[((size_t)oi * 31 + k) * N2 + idx] = f_canon(acc[k]);
}

template <int KIND> __global__ void __launch_bounds__(128) k_segment(QArgs a, Segment s) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= a.npoints) return;
  QPoint q;
  qpoint_begin(a, idx, q);
  Segment ss = s; ss.kind = (SegKind)KIND;   // compile-time kind: each instantiation keeps only its own code
  eval_segment(a, ss, idx, q);
  qpoint_end(a, idx, q);
}

// The curve gadgets (zero / new_x / new_y modular operations: 33 + 66 + 66 constraints for G1, twice that for G2) as one
// instantiation per operation: a third of the live limb arrays each (166 registers + 976 B of local arrays for the whole
// G1 addition, 168 + 2 800 B for G2).
template <int KIND, int PART> __global__ void __launch_bounds__(128) k_segment_part(QArgs a, Segment s) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= a.npoints) return;
  QPoint q;
  qpoint_begin(a, idx, q);
  if (KIND == SEG_G1_ADD) eval_g1_add(q, q.lv(s.p1), s.p0, PART);
  else if (KIND == SEG_G1_DOUBLE) eval_g1_double(q, q.lv(s.p1), s.p0, PART);
  else if (KIND == SEG_G2_ADD) eval_g2_add(q, q.lv(s.p1), s.p0, PART);
  else eval_g2_double(q, q.lv(s.p1), s.p0, PART);
  qpoint_end(a, idx, q);
}
template <int KIND> static void launch_gadget_parts(sbn_ctx* ctx, const QArgs& a0, const Segment& s, unsigned blocks) {
  const size_t unit = (KIND == SEG_G1_ADD || KIND == SEG_G1_DOUBLE) ? 33 : 66, m[3] = {unit, 2 * unit, 2 * unit};
  QArgs a = a0;
  for (int part = 1; part <= 3; part++) {
    for (int c = 0; c < SBN_MAX_CHALLENGES; c++) a.alpha_m[c] = gl_pow(a.alpha[c], m[part - 1]);
    if (part == 1) k_segment_part<KIND, 1><<<blocks, 128, 0, ctx->stream>>>(a, s);
    else if (part == 2) k_segment_part<KIND, 2><<<blocks, 128, 0, ctx->stream>>>(a, s);
    else k_segment_part<KIND, 3><<<blocks, 128, 0, ctx->stream>>>(a, s);
    LAUNCH_CHECK(ctx);
    a.first = 0;
  }
}

static const char* seg_name(SegKind k) {
  switch (k) {
    case SEG_SPLIT_RANGE_CHECK: return "q_split_range_check"; case SEG_MODULAR_CORE: return "q_modular_core"; case SEG_G1_CORE: return "q_g1_core";
    case SEG_FLAGS: return "q_flags"; case SEG_G1_ADD: return "q_g1_add"; case SEG_G1_DOUBLE: return "q_g1_double"; case SEG_PERIODIC_PULSE: return "q_periodic_pulse";
    case SEG_PULSE: return "q_pulse"; case SEG_U16_RANGE_CHECK: return "q_u16_range_check"; case SEG_PERMUTATION: return "q_permutation";
    case SEG_FQ_CORE: return "q_fq_core"; case SEG_FQ_MUL: return "q_fq_mul"; case SEG_G2_CORE: return "q_g2_core"; case SEG_G2_ADD: return "q_g2_add";
    case SEG_G2_DOUBLE: return "q_g2_double"; case SEG_FQ12_CORE: return "q_fq12_core"; case SEG_FQ12_MUL: return "q_fq12_mul"; case SEG_FLAGS_U64: return "q_flags_u64";
  }
  return "q_other";
}
static void launch_segment(sbn_ctx* ctx, const QArgs& a, const Segment& s) {
  unsigned blocks = (unsigned)((a.npoints + 127) / 128);
  if ((s.kind == SEG_PERMUTATION || s.kind == SEG_SPLIT_RANGE_CHECK) && a.npoints <= (size_t(1) << 15)) {
    // item ranges: enough chunks for ~2^18 threads, at least 16 items each
    const int nitems = s.kind == SEG_PERMUTATION ? a.nz : s.p2;
    int nchunks = (int)std::min<size_t>((size_t(1) << 18) / a.npoints, (size_t)std::max(1, nitems / 16));
    if (nchunks > 1) {
      ChunkPlan plan; plan.nitems = nitems; plan.per = (nitems + nchunks - 1) / nchunks; plan.nchunks = (nitems + plan.per - 1) / plan.per;
      // constraints that follow each run of chunk j: permutation list = nz first-row + nz products; split list = n + 4 n + 3
      std::vector<u64> w((size_t)plan.nchunks * 3 * SBN_MAX_CHALLENGES, 0);
      for (int j = 0; j < plan.nchunks; j++) {
        const u64 i1 = std::min(nitems, (j + 1) * plan.per), n = nitems;
        u64 after[3];
        if (s.kind == SEG_PERMUTATION) { after[0] = (n - i1) + n; after[1] = n - i1; after[2] = 0; }
        else { after[0] = (n - i1) + 4 * n + 3; after[1] = 4 * (n - i1) + 3; after[2] = 0; }
        for (int r = 0; r < 3; r++) for (int c = 0; c < SBN_MAX_CHALLENGES; c++) w[((size_t)j * 3 + r) * SBN_MAX_CHALLENGES + c] = gl_pow(a.alpha[c], after[r]);
      }
      DevBuf<u64> d_w(ctx, w.size()), partial(ctx, (size_t)plan.nchunks * SBN_MAX_CHALLENGES * a.npoints);
      ctx->upload(d_w, w.data(), w.size() * 8);
      KScope ks(ctx, seg_name(s.kind));
      k_segment_chunked<<<dim3(blocks, plan.nchunks), 128, 0, ctx->stream>>>(a, s, plan, d_w, partial); LAUNCH_CHECK(ctx);
      k_combine_chunks<<<blocks, 128, 0, ctx->stream>>>(a, plan.nchunks, partial); LAUNCH_CHECK(ctx);
      return;
    }
  }
  if (s.kind == SEG_FQ12_MUL) {   // x = a; y = a (square) or b (mul): columns 0 / 192 of the row
    KScope kp(ctx, "q_fq12_products");
    k_fq12_products<<<dim3(blocks, 12), 128, 0, ctx->stream>>>(a, 0, s.p1 ? 0 : 192, a.scratch);
    LAUNCH_CHECK(ctx);
  }
  KScope ks(ctx, seg_name(s.kind));
  if (s.kind == SEG_G1_ADD) { launch_gadget_parts<SEG_G1_ADD>(ctx, a, s, blocks); return; }
  if (s.kind == SEG_G1_DOUBLE) { launch_gadget_parts<SEG_G1_DOUBLE>(ctx, a, s, blocks); return; }
  if (s.kind == SEG_G2_ADD) { launch_gadget_parts<SEG_G2_ADD>(ctx, a, s, blocks); return; }
  if (s.kind == SEG_G2_DOUBLE) { launch_gadget_parts<SEG_G2_DOUBLE>(ctx, a, s, blocks); return; }
#define SEGCASE(K) case K: k_segment<K><<<blocks, 128, 0, ctx->stream>>>(a, s); break;
  switch (s.kind) {
    SEGCASE(SEG_SPLIT_RANGE_CHECK) SEGCASE(SEG_MODULAR_CORE) SEGCASE(SEG_G1_CORE) SEGCASE(SEG_FLAGS) SEGCASE(SEG_G1_ADD)
    SEGCASE(SEG_G1_DOUBLE) SEGCASE(SEG_PERIODIC_PULSE) SEGCASE(SEG_PULSE) SEGCASE(SEG_U16_RANGE_CHECK) SEGCASE(SEG_PERMUTATION)
    SEGCASE(SEG_FQ_CORE) SEGCASE(SEG_FQ_MUL) SEGCASE(SEG_G2_CORE) SEGCASE(SEG_G2_ADD) SEGCASE(SEG_G2_DOUBLE) SEGCASE(SEG_FQ12_CORE) SEGCASE(SEG_FQ12_MUL)
    SEGCASE(SEG_FLAGS_U64)
  }
#undef SEGCASE
  LAUNCH_CHECK(ctx);
}

// Sparse public-input binding columns (constraints.cuh "public-input binding ... folded over the instances"):
// vals[col][row], col 0 = sum of output pulses; per challenge c: 1 + c * (io_len + 2) + {0: S_in, 1: S_out, 2 + u: U_u}.
__global__ void k_pi_columns(u64* __restrict__ vals, size_t N, const u64* __restrict__ pi, const u64* __restrict__ a /* [chal][num_io] */, int num_io,
                             int rows_per_io, int group_count, int num_challenges) {
  const int io_len = pi_io_len(group_count), per = io_len + 2;
  const int w = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (w >= per) return;
  const size_t pos_in = (size_t)i * rows_per_io, pos_out = pos_in + rows_per_io - 1;
  if (w == 0) vals[pos_out] = 1;
  for (int c = 0; c < num_challenges; c++) {
    const u64 ai = a[(size_t)c * num_io + i];
    u64* col = vals + (size_t)(1 + c * per + w) * N;
    if (w == 0) col[pos_in] = ai;
    else if (w == 1) col[pos_out] = ai;
    else {
      int pidx, kind;
      pi_map(group_count, w - 2, pidx, kind);
      col[kind ? pos_out : pos_in] = gl_mul(ai, pi[(size_t)i * io_len + pidx]);
    }
  }
}
// Returns the low-degree extension of the binding columns on the evaluation points: out[col][point].
static void build_pi_binding(sbn_ctx* ctx, QArgs& a, const QDomain& dom, const Segment& core, const u64* d_public_inputs, int num_challenges, DevBuf<u64>& lde) {
  const int group_count = core.kind == SEG_FQ_CORE ? 1 : core.kind == SEG_G1_CORE ? 2 : core.kind == SEG_G2_CORE ? 4 : (core.p2 ? -1 : 0);
  num_challenges = SBN_MAX_CHALLENGES;   // unused challenges have alpha = 0; their columns exist so the kernel never reads out of bounds
  const int num_io = core.p0, io_len = pi_io_len(group_count), per = io_len + 2, ncols = 1 + num_challenges * per;
  const int logn = a.logn; const size_t N = size_t(1) << logn;
  const int rows_per_io = (int)(N / num_io);
  KScope ks(ctx, "q_pi_binding");
  std::vector<u64> ha((size_t)num_challenges * num_io);
  for (int c = 0; c < num_challenges; c++) {
    const u64 step = gl_pow(a.alpha[c], (u64)io_len);
    u64 v = 1;
    for (int i = num_io - 1; i >= 0; i--) { ha[(size_t)c * num_io + i] = v; v = gl_mul(v, step); }   // a_i = alpha^((n-1-i) io_len)
    a.pi_skip[c] = ha[(size_t)c * num_io];                                                            // alpha^((n-1) io_len)
  }
  DevBuf<u64> d_a(ctx, ha.size()), vals(ctx, (size_t)ncols * N), coeffs(ctx, (size_t)ncols * N);
  ctx->upload(d_a, ha.data(), ha.size() * 8);
  CUDA_CHECK(cudaMemsetAsync(vals, 0, (size_t)ncols * N * 8, ctx->stream));
  k_pi_columns<<<dim3((per + 63) / 64, num_io), 64, 0, ctx->stream>>>(vals, N, d_public_inputs, d_a, num_io, rows_per_io, group_count, num_challenges);
  LAUNCH_CHECK(ctx);
  intt_columns(ctx, vals, coeffs, ncols, logn);
  lde = DevBuf<u64>(ctx, (size_t)ncols * a.npoints);
  lde_class(ctx, coeffs, lde, ncols, logn, 1, dom.m, dom.sigma);   // m = 0: both half-cosets, [col][bq][k]
  a.pi_lde = lde; a.pi_per_chal = per;
}

__global__ void k_fill_lagrange_coeffs(u64* coeffs, const u64* wpow, u64 ninv, size_t N) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= N) return;
  coeffs[j] = ninv;                      // ifft(selector(0))      = 1/N
  coeffs[N + j] = gl_mul(ninv, wpow[j]);  // ifft(selector(N - 1))  = w^j / N
}
__global__ void k_scale_cosets(u64* acc, int logm, size_t npoints, int num_challenges, u64 zh0, u64 zh1) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npoints * num_challenges) return;
  int sg = (int)((i % npoints) >> logm);
  acc[i] = gl_mul(acc[i], sg ? zh1 : zh0);
}
// f_lo = (u0 + u1)/2, f_hi = (u0 - u1)/(2 g^N): the two degree-N chunks of a degree-2N quotient
// from its unscaled per-coset interpolants u_b = f_lo + (-1)^b g^N f_hi.
__global__ void k_quotient_split(const u64* u0, const u64* u1, u64* lo, u64* hi, u64 inv2, u64 inv2gn, size_t N) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= N) return;
  u64 a = u0[j], b = u1[j];
  lo[j] = gl_mul(gl_add(a, b), inv2);
  hi[j] = gl_mul(gl_sub(a, b), inv2gn);
}

size_t quotient_points(const QDomain& dom, int logn) { return (size_t(2) << logn) >> dom.m; }

void quotient_eval(sbn_ctx* ctx, const AirDesc& air, const QDomain& dom, const u64* trace_lde, const u64* zs_lde, const PermInstances& perm,
                   const u64* d_public_inputs, const u64* alphas, int num_challenges, int logn, int rate_bits, u64* d_acc) {
  SBN_REQUIRE(air.quotient_degree_factor() == 2, "only constraint_degree 3 (quotient degree factor 2) is supported");
  SBN_REQUIRE(rate_bits >= 1, "constraint degree higher than the rate is not supported");
  SBN_REQUIRE(num_challenges >= 1 && num_challenges <= SBN_MAX_CHALLENGES, "unsupported num_challenges");
  SBN_REQUIRE(dom.m <= logn, "too many quotient classes");
  const size_t N = size_t(1) << logn, R = size_t(1) << rate_bits, L = N * R;
  const size_t step = size_t(1) << (rate_bits - 1);
  const u64 w2n = ctx->root_of_unity(logn + 1);
  QArgs a; memset(&a, 0, sizeof a);
  a.logn = logn;
  if (dom.m == 0) {   // two half-cosets inside the LDE batches
    a.logm = logn; a.nseg = 2; a.npoints = 2 * N; a.next_shift = 1;
    a.trace = a.trace_next = trace_lde; a.trace_stride = L; a.zs = a.zs_next = zs_lde; a.zs_stride = L;
    for (int bq = 0; bq < 2; bq++) { a.seg_off[bq] = (size_t)bq * step * N; a.seg_shift[bq] = gl_mul(ctx->coset_shift(), gl_pow(w2n, bq)); }
  } else {            // one class of 2^m: its own batch, "next" in the same batch (G = 2) or in the class + 2 batch
    const u32 G = 1u << dom.m;
    a.logm = logn + 1 - dom.m; a.nseg = 1; a.npoints = size_t(1) << a.logm;
    a.trace = dom.trace; a.trace_next = dom.trace_next; a.trace_stride = a.npoints;
    a.zs = dom.zs; a.zs_next = dom.zs_next; a.zs_stride = a.npoints;
    a.next_shift = G == 2 ? 1 : (dom.sigma + 2 >= G ? 1 : 0);
    a.seg_off[0] = 0; a.seg_shift[0] = gl_mul(ctx->coset_shift(), gl_pow(w2n, dom.sigma));
  }
  a.wpow = get_ntt_tables(ctx, a.logm).w_fwd; a.w_inv = gl_inv(ctx->root_of_unity(logn));
  // Lagrange selectors on the evaluation points
  const NttTables& tb = get_ntt_tables(ctx, logn);
  DevBuf<u64> lag_coeffs(ctx, 2 * N), lag_lde(ctx, dom.m == 0 ? 2 * L : 2 * a.npoints);
  u64 ninv = gl_inv((u64)N);
  k_fill_lagrange_coeffs<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(lag_coeffs, tb.w_fwd, ninv, N);
  LAUNCH_CHECK(ctx);
  if (dom.m == 0) { lde_columns(ctx, lag_coeffs, lag_lde, 2, logn, rate_bits); a.lagrange_stride = L; }
  else { lde_class(ctx, lag_coeffs, lag_lde, 2, logn, 1, dom.m, dom.sigma); a.lagrange_stride = a.npoints; }
  a.lagrange = lag_lde; a.pi = d_public_inputs; a.acc = d_acc;
  for (int c = 0; c < SBN_MAX_CHALLENGES; c++) a.alpha[c] = c < num_challenges ? alphas[c] : 0;
  DevBuf<u32> d_lhs, d_rhs; DevBuf<u64> d_gamma;
  std::vector<Segment> segs = air.segments;
  if (perm.nz()) {
    size_t ne = perm.lhs.size();
    d_lhs = DevBuf<u32>(ctx, ne); d_rhs = DevBuf<u32>(ctx, ne); d_gamma = DevBuf<u64>(ctx, ne);
    ctx->upload(d_lhs, perm.lhs.data(), ne * 4);
    ctx->upload(d_rhs, perm.rhs.data(), ne * 4);
    ctx->upload(d_gamma, perm.gamma.data(), ne * 8);
    a.perm_lhs = d_lhs; a.perm_rhs = d_rhs; a.perm_gamma = d_gamma; a.perm_batch = perm.batch_size; a.nz = (int)perm.nz();
    segs.push_back({SEG_PERMUTATION, 0, 0, 0, 0, 2 * perm.nz()});
  }
  DevBuf<u64> scratch, pi_lde;
  a.pi_lde = d_acc; a.pi_per_chal = 0;   // valid pointer for AIRs without a public-input block (never dereferenced there)
  for (const Segment& s : segs)
    if (s.kind == SEG_FQ_CORE || s.kind == SEG_G1_CORE || s.kind == SEG_G2_CORE || s.kind == SEG_FQ12_CORE) build_pi_binding(ctx, a, dom, s, d_public_inputs, num_challenges, pi_lde);
  for (const Segment& s : segs) if (s.kind == SEG_FQ12_MUL && !scratch.p) scratch = DevBuf<u64>(ctx, (size_t)12 * 31 * a.npoints);
  a.scratch = scratch.p;
  bool first = true;
  for (const Segment& s : segs) {
    a.first = first ? 1 : 0;
    for (int c = 0; c < SBN_MAX_CHALLENGES; c++) a.alpha_m[c] = gl_pow(a.alpha[c], s.num_constraints);
    launch_segment(ctx, a, s);
    first = false;
  }
  // divide by Z_H(x) = x^N - 1, constant on a segment: (seg_shift w_M^k)^N = seg_shift^N
  u64 zh[2] = {0, 0};
  for (int sg = 0; sg < a.nseg; sg++) zh[sg] = gl_inv(gl_sub(gl_exp_pow2(a.seg_shift[sg], logn), 1));
  size_t tot = a.npoints * num_challenges;
  k_scale_cosets<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(d_acc, a.logm, a.npoints, num_challenges, zh[0], zh[1]);
  LAUNCH_CHECK(ctx);
}

// per-coset interpolation (coset_ifft restricted to each half), then split into the two chunks
void quotient_finish(sbn_ctx* ctx, const u64* d_acc, int num_challenges, int logn, u64* out_chunks) {
  const size_t N = size_t(1) << logn;
  const u64 w2n = ctx->root_of_unity(logn + 1), gn = gl_exp_pow2(ctx->coset_shift(), logn);
  DevBuf<u64> u(ctx, 2 * N);
  u64 inv2 = gl_inv(2), inv2gn = gl_inv(gl_mul(2, gn));
  for (int c = 0; c < num_challenges; c++) {
    for (int bq = 0; bq < 2; bq++) {
      const u64* post = get_pow_table(ctx, gl_inv(gl_mul(ctx->coset_shift(), gl_pow(w2n, bq))), logn);
      ntt_batch(ctx, d_acc + ((size_t)c * 2 + bq) * N, N, u + (size_t)bq * N, N, 1, logn, true, 0, post);
    }
    k_quotient_split<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(u, u + N, out_chunks + (size_t)(2 * c) * N, out_chunks + (size_t)(2 * c + 1) * N, inv2, inv2gn, N);
    LAUNCH_CHECK(ctx);
  }
}

void compute_quotient_chunks(sbn_ctx* ctx, const AirDesc& air, const u64* trace_lde, const u64* zs_lde, const PermInstances& perm,
                             const u64* d_public_inputs, const u64* alphas, int num_challenges, int logn, int rate_bits, u64* out_chunks) {
  DevBuf<u64> acc(ctx, (size_t)SBN_MAX_CHALLENGES * (size_t(2) << logn));
  quotient_eval(ctx, air, QDomain(), trace_lde, zs_lde, perm, d_public_inputs, alphas, num_challenges, logn, rate_bits, acc);
  quotient_finish(ctx, acc, num_challenges, logn, out_chunks);
}
