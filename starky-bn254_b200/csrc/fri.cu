// K6: openings at zeta / g*zeta, FRI batch reduction, folding, commit-phase trees, proof-of-work grind
// and query openings.  Replaces starky's `StarkOpeningSet::new` and plonky2's
// `PolynomialBatch::prove_openings` / `fri_proof` (external dependency), reached
// from the reference through `prove()` (reference src/curves/g1/exp.rs:818).
#include "fri.cuh"
#include "ntt.cuh"
#include "poseidon.cuh"

// out[i] = base^i (extension field, SoA): a thread raises base to the first exponent of its run of 16 and multiplies from there
#define EXTPOW_RUN 16
__global__ void k_ext_powers(u64* out_a, u64* out_b, gl2 base, size_t n) {
  const size_t i0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * EXTPOW_RUN;
  if (i0 >= n) return;
  gl2 r = gl2_pow(base, i0);
  for (size_t i = i0; i < n && i < i0 + EXTPOW_RUN; i++) { out_a[i] = r.a; out_b[i] = r.b; r = gl2_mul(r, base); }
}
static void launch_ext_powers(sbn_ctx* ctx, u64* out_a, u64* out_b, gl2 base, size_t n) {
  const size_t threads = (n + EXTPOW_RUN - 1) / EXTPOW_RUN;
  k_ext_powers<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(out_a, out_b, base, n); LAUNCH_CHECK(ctx);
}

// ---- openings ----
// One block per COLS columns; a column is read once (coalesced) and multiplied by the power tables, which every block streams from
// L2 (the blocks of a wave walk the tables at the same pace).  Shape chosen by measurement on the G1 trace (gpurun_out/r2ag_*, ms
// per proof): all five loads of an iteration issued before the four multiply-accumulates, one column per block, one iteration
// in flight: 1.00; the same loop with the loads written inside the multiply-accumulate calls: 1.41; four iterations in flight:
// 1.61; two / three columns per block sharing each table load: 1.71 / 2.15.
template <int COLS, int U> __global__ void __launch_bounds__(256) k_eval_two_points(const u64* __restrict__ coeffs, size_t N, int ncols,
                                                                                    const u64* __restrict__ pw /* [4][N]: z.a z.b zn.a zn.b */, u64* __restrict__ out) {
  const int c0 = blockIdx.x * COLS;
  const u64* col[COLS];
#pragma unroll
  for (int c = 0; c < COLS; c++) col[c] = coeffs + (size_t)min(c0 + c, ncols - 1) * N;   // a ragged last block recomputes the last column
  gl_acc acc[COLS][4];
#pragma unroll
  for (int c = 0; c < COLS; c++)
#pragma unroll
    for (int k = 0; k < 4; k++) acc[c][k] = gl_acc_zero();   // unreduced sums of products
#pragma unroll 1
  for (size_t j0 = threadIdx.x; j0 < N; j0 += (size_t)blockDim.x * U) {   // U > 1 needs N to be a multiple of 256 U
    u64 v[U][COLS], p[U][4];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t j = j0 + (size_t)u * blockDim.x;
#pragma unroll
      for (int c = 0; c < COLS; c++) v[u][c] = col[c][j];
#pragma unroll
      for (int k = 0; k < 4; k++) p[u][k] = pw[(size_t)k * N + j];
    }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int c = 0; c < COLS; c++)
#pragma unroll
        for (int k = 0; k < 4; k++) gl_acc_mac(acc[c][k], v[u][c], p[u][k]);
  }
  __shared__ u64 red[COLS * 4][256];
#pragma unroll
  for (int c = 0; c < COLS; c++)
#pragma unroll
    for (int k = 0; k < 4; k++) red[c * 4 + k][threadIdx.x] = gl_acc_reduce(acc[c][k]);
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) for (int k = 0; k < COLS * 4; k++) red[k][threadIdx.x] = gl_add(red[k][threadIdx.x], red[k][threadIdx.x + d]);
    __syncthreads();
  }
  if (threadIdx.x < COLS * 4 && c0 + (int)(threadIdx.x >> 2) < ncols) out[(size_t)c0 * 4 + threadIdx.x] = red[threadIdx.x][0];
}

DevBuf<u64> two_point_power_table(sbn_ctx* ctx, int logn, gl2 z0, gl2 z1) {
  size_t N = size_t(1) << logn;
  DevBuf<u64> pw(ctx, 4 * N);
  launch_ext_powers(ctx, pw, pw + N, z0, N);
  launch_ext_powers(ctx, pw + 2 * N, pw + 3 * N, z1, N);
  return pw;
}

void eval_columns_at_two_points(sbn_ctx* ctx, const u64* coeffs, int ncols, int logn, const u64* pw, u64* d_out) {
  if (ncols <= 0) return;
  KScope ks(ctx, "openings_eval");
  const size_t N = size_t(1) << logn;
  k_eval_two_points<1, 1><<<ncols, 256, 0, ctx->stream>>>(coeffs, N, ncols, pw, d_out);
  LAUNCH_CHECK(ctx);
}
void eval_columns_at_two_points(sbn_ctx* ctx, const u64* coeffs, int ncols, int logn, gl2 zeta, gl2 zeta_next, u64* d_out) {
  if (ncols <= 0) return;
  DevBuf<u64> pw = two_point_power_table(ctx, logn, zeta, zeta_next);
  eval_columns_at_two_points(ctx, coeffs, ncols, logn, pw, d_out);
}

// ---- batch reduction: partial[g][2][N] = sum over the g-th slice of columns of alpha^j f_j ----
__global__ void __launch_bounds__(256) k_reduce_columns(const u64* __restrict__ coeffs, size_t N, int ncols, int cols_per_group,
                                                        const u64* __restrict__ apow_a, const u64* __restrict__ apow_b, int apow_off,
                                                        u64* __restrict__ partial) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  int c0 = blockIdx.y * cols_per_group, c1 = min(ncols, c0 + cols_per_group);
  gl_acc sa = gl_acc_zero(), sb = gl_acc_zero();   // unreduced sums of products, one reduction per output
  for (int c = c0; c < c1; c++) {
    const u64 v = coeffs[(size_t)c * N + i];
    gl_acc_mac(sa, v, apow_a[apow_off + c]); gl_acc_mac(sb, v, apow_b[apow_off + c]);
  }
  u64* p = partial + (size_t)blockIdx.y * 2 * N;
  p[i] = gl_acc_reduce(sa); p[N + i] = gl_acc_reduce(sb);
}
__global__ void k_sum_partials(const u64* partial, int ngroups, size_t N, u64* out /* [2][N] */, int accumulate) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= 2 * N) return;
  u64 s = accumulate ? out[i] : 0;
  for (int g = 0; g < ngroups; g++) s = gl_add(s, partial[(size_t)g * 2 * N + i]);
  out[i] = s;
}

// out[i] = base[i] + sum_{r < nparts} parts[r * stride + offset + i], i < n
__global__ void k_sum_strided(const u64* __restrict__ parts, int nparts, size_t stride, size_t offset, size_t n, u64* __restrict__ out, const u64* __restrict__ base) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 s = base ? base[i] : 0;
  for (int r = 0; r < nparts; r++) s = gl_add(s, parts[(size_t)r * stride + offset + i]);
  out[i] = s;
}

// ---- (comp(X) - comp(z)) / (X - z):  Q[i] = z^-(i+1) * sum_{j>i} comp[j] z^j, Q[N-1] = 0 ----
// A suffix sum in three steps: every thread sums its run of DIV_RUN terms, one block turns the run totals into exclusive suffix
// sums, every thread walks its run from the top.  comp [2][N]; zp = z^j, zi = z^-j (ext SoA tables, stride N).
#define DIV_RUN 16
__global__ void __launch_bounds__(256) k_div_run_sums(const u64* __restrict__ comp, size_t N, const u64* __restrict__ zp, u64* __restrict__ tot /* [2][T] */, size_t T) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= T) return;
  const size_t lo = t * DIV_RUN, hi = min(N, lo + DIV_RUN);
  gl2 s = gl2_make(0, 0);
  for (size_t j = lo; j < hi; j++) s = gl2_add(s, gl2_mul(gl2_make(comp[j], comp[N + j]), gl2_make(zp[j], zp[N + j])));
  tot[t] = s.a; tot[T + t] = s.b;
}
__global__ void __launch_bounds__(1024) k_div_suffix(u64* __restrict__ tot /* [2][T]: totals in, exclusive suffix sums out */, size_t T) {
  __shared__ u64 pa[1024], pb[1024];
  const size_t chunk = (T + blockDim.x - 1) / blockDim.x;
  const size_t lo = min(T, (size_t)threadIdx.x * chunk), hi = min(T, lo + chunk);
  gl2 s = gl2_make(0, 0);
  for (size_t j = lo; j < hi; j++) s = gl2_add(s, gl2_make(tot[j], tot[T + j]));
  pa[threadIdx.x] = s.a; pb[threadIdx.x] = s.b;
  __syncthreads();
  gl2 running = gl2_make(0, 0);   // sum over the pieces after mine
  for (int u = threadIdx.x + 1; u < (int)blockDim.x; u++) running = gl2_add(running, gl2_make(pa[u], pb[u]));
  for (size_t j = hi; j-- > lo;) {
    const gl2 v = gl2_make(tot[j], tot[T + j]);
    tot[j] = running.a; tot[T + j] = running.b;
    running = gl2_add(running, v);
  }
}
__global__ void __launch_bounds__(256) k_div_finish(const u64* __restrict__ comp, size_t N, const u64* __restrict__ zp, const u64* __restrict__ zi,
                                                    const u64* __restrict__ suf /* [2][T] */, size_t T, u64* __restrict__ q /* [2][N] */) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= T) return;
  const size_t lo = t * DIV_RUN, hi = min(N, lo + DIV_RUN);
  gl2 running = gl2_make(suf[t], suf[T + t]);   // sum_{j >= hi} comp[j] z^j
  for (size_t i = hi; i-- > lo;) {
    gl2 v;
    if (i + 1 < N) v = gl2_mul(running, gl2_make(zi[i + 1], zi[N + i + 1])); else v = gl2_make(0, 0);
    q[i] = v.a; q[N + i] = v.b;
    running = gl2_add(running, gl2_mul(gl2_make(comp[i], comp[N + i]), gl2_make(zp[i], zp[N + i])));
  }
}
static void divide_by_linear(sbn_ctx* ctx, const u64* comp, size_t N, const u64* zp, const u64* zi, u64* q) {
  const size_t T = (N + DIV_RUN - 1) / DIV_RUN;
  DevBuf<u64> tot(ctx, 2 * T);
  KScope ks(ctx, "fri_divide_by_linear");
  k_div_run_sums<<<(unsigned)((T + 255) / 256), 256, 0, ctx->stream>>>(comp, N, zp, tot, T); LAUNCH_CHECK(ctx);
  k_div_suffix<<<1, 1024, 0, ctx->stream>>>(tot, T); LAUNCH_CHECK(ctx);
  k_div_finish<<<(unsigned)((T + 255) / 256), 256, 0, ctx->stream>>>(comp, N, zp, zi, tot, T, q); LAUNCH_CHECK(ctx);
}
// final = q0 * shift + q1, written into the first N entries of the zero-padded [2][L] coefficient array
__global__ void k_combine_final(const u64* q0, const u64* q1, gl2 shift, size_t N, size_t L, u64* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= L) return;
  gl2 v = gl2_make(0, 0);
  if (i < N) v = gl2_add(gl2_mul(gl2_make(q0[i], q0[N + i]), shift), gl2_make(q1[i], q1[N + i]));
  out[i] = v.a; out[L + i] = v.b;
}

void fri_final_poly(sbn_ctx* ctx, const std::vector<OracleView>& oracles, int logn, int rate_bits, gl2 alpha, gl2 zeta, gl2 zeta_next,
                    u64* d_final_coeffs, const ColumnSplit* split, const u64* pw) {
  const size_t N = size_t(1) << logn, L = N << rate_bits;
  int total = 0; for (auto& o : oracles) total += o.ncols;
  int n1 = total - oracles.back().ncols;   // the zeta_next batch omits the last oracle (quotient polys)
  // alpha^j table (ext SoA)
  DevBuf<u64> apow(ctx, 2 * (size_t)total);
  launch_ext_powers(ctx, apow, apow + total, alpha, total);
  const int G = 16;
  const int world = split ? split->world : 1, rank = split ? split->rank : 0;
  // mine[0] = my share of the oracles both batches contain, mine[1] = my share of the last oracle (zeta batch only)
  DevBuf<u64> partial(ctx, (size_t)G * 2 * N), mine(ctx, 4 * N), comp1(ctx, 2 * N), comp0(ctx, 2 * N);
  CUDA_CHECK(cudaMemsetAsync(mine, 0, 4 * N * 8, ctx->stream));
  int off = 0;
  for (size_t o = 0; o < oracles.size(); o++) {
    const int nc_all = oracles[o].ncols, per = (nc_all + world - 1) / world;
    const int c0 = std::min(nc_all, rank * per), nc = std::min(nc_all, c0 + per) - c0;
    const bool last = o + 1 == oracles.size();
    if (nc > 0) {
      int cpg = (nc + G - 1) / G, ng = (nc + cpg - 1) / cpg;
      dim3 grid((unsigned)((N + 255) / 256), ng);
      KScope ks(ctx, "fri_reduce_columns");
      k_reduce_columns<<<grid, 256, 0, ctx->stream>>>(oracles[o].coeffs + (size_t)c0 * N, N, nc, cpg, apow, apow + total, off + c0, partial); LAUNCH_CHECK(ctx);
      k_sum_partials<<<(unsigned)((2 * N + 255) / 256), 256, 0, ctx->stream>>>(partial, ng, N, mine + (last ? 2 * N : 0), 1); LAUNCH_CHECK(ctx);
    }
    off += nc_all;
  }
  if (world > 1) {
    DevBuf<u64> all(ctx, (size_t)world * 4 * N);
    ctx->sync();
    split->gather_device(mine, 4 * N * 8, all);
    // comp1 = sum_r all[r][0], comp0 = comp1 + sum_r all[r][1]: the blocks are laid out as `2 world` partials of [2][N]
    k_sum_strided<<<(unsigned)((2 * N + 255) / 256), 256, 0, ctx->stream>>>(all, world, 4 * N, 0, 2 * N, comp1, nullptr); LAUNCH_CHECK(ctx);
    k_sum_strided<<<(unsigned)((2 * N + 255) / 256), 256, 0, ctx->stream>>>(all, world, 4 * N, 2 * N, 2 * N, comp0, comp1); LAUNCH_CHECK(ctx);
  } else {
    CUDA_CHECK(cudaMemcpyAsync(comp1, mine, 2 * N * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    k_sum_strided<<<(unsigned)((2 * N + 255) / 256), 256, 0, ctx->stream>>>(mine, 1, 4 * N, 2 * N, 2 * N, comp0, comp1); LAUNCH_CHECK(ctx);
  }
  // quotients of the two batches: z^j from the caller's table when it has one, z^-j built here
  DevBuf<u64> q0(ctx, 2 * N), q1(ctx, 2 * N), own_pw;
  if (!pw) { own_pw = two_point_power_table(ctx, logn, zeta, zeta_next); pw = own_pw; }
  DevBuf<u64> pwi = two_point_power_table(ctx, logn, gl2_inv(zeta), gl2_inv(zeta_next));
  divide_by_linear(ctx, comp0, N, pw, pwi, q0);
  divide_by_linear(ctx, comp1, N, pw + 2 * N, pwi + 2 * N, q1);
  gl2 shift = gl2_pow(alpha, (u64)n1);
  k_combine_final<<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>(q0, q1, shift, N, L, d_final_coeffs); LAUNCH_CHECK(ctx);
}

// ---- commit-phase layer ----
__global__ void __launch_bounds__(128) k_fri_leaves(const u64* __restrict__ values /* [2][n] natural */, int logsize, int arity_bits,
                                                    u64* __restrict__ leaves, u64* __restrict__ digests) {
  const size_t n = size_t(1) << logsize, nleaves = n >> arity_bits;
  const size_t l = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (l >= nleaves) return;
  const int arity = 1 << arity_bits;
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = 0;
  u64* out = leaves + l * 2 * arity;
  // leaf = flatten(values_bitrev[l*arity .. (l+1)*arity)), hashed 8 field elements (4 ext values) at a time
  for (int t = 0; t < arity; t += 4) {
    u64 v[8];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      size_t i = bitrev32((u32)(l * arity + t + e), logsize);
      v[2 * e] = values[i]; v[2 * e + 1] = values[n + i];
    }
#pragma unroll
    for (int e = 0; e < 8; e++) { st[e] = v[e]; out[2 * t + e] = v[e]; }
    poseidon_permute(st);
  }
  ulonglong2* d = reinterpret_cast<ulonglong2*>(digests + l * 4);
  d[0] = make_ulonglong2(st[0], st[1]); d[1] = make_ulonglong2(st[2], st[3]);
}

void fri_commit_layer(sbn_ctx* ctx, const u64* d_values, int logsize, int arity_bits, int cap_height, FriLayer* out) {
  SBN_REQUIRE(arity_bits >= 2, "FRI arity below 4 is not supported");  // leaf must exceed 4 elements and be a multiple of 8
  size_t nleaves = (size_t(1) << logsize) >> arity_bits;
  out->nleaves = nleaves; out->arity_bits = arity_bits;
  out->leaves = DevBuf<u64>(ctx, nleaves * 2 * (size_t(1) << arity_bits));
  merkle_alloc(ctx, &out->tree, nleaves, cap_height);
  k_fri_leaves<<<(unsigned)((nleaves + 127) / 128), 128, 0, ctx->stream>>>(d_values, logsize, arity_bits, out->leaves, out->tree.digests);
  LAUNCH_CHECK(ctx);
  merkle_build_from_leaf_digests(ctx, &out->tree);
}

__global__ void k_fri_fold(const u64* __restrict__ c, size_t n, int arity_bits, gl2 beta, u64* __restrict__ out) {
  size_t m = n >> arity_bits;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  int arity = 1 << arity_bits;
  gl2 acc = gl2_make(0, 0);
  for (int k = arity - 1; k >= 0; k--) {
    size_t j = (i << arity_bits) + k;
    acc = gl2_add(gl2_mul(acc, beta), gl2_make(c[j], c[n + j]));
  }
  out[i] = acc.a; out[m + i] = acc.b;
}
void fri_fold_coeffs(sbn_ctx* ctx, const u64* d_coeffs, size_t n, int arity_bits, gl2 beta, u64* d_out) {
  size_t m = n >> arity_bits;
  k_fri_fold<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(d_coeffs, n, arity_bits, beta, d_out);
  LAUNCH_CHECK(ctx);
}

// ---- proof of work ----
struct PowState { u64 s[12]; };
__global__ void __launch_bounds__(128) k_pow(PowState st0, int pos, int pow_bits, u64 base, unsigned long long* best) {
  u64 cand = base + blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (cand >= GL_P) return;
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = (i == pos) ? cand : st0.s[i];
  poseidon_permute(st);
  if ((st[7] >> (64 - pow_bits)) == 0) atomicMin(best, (unsigned long long)cand);
}
u64 fri_pow_search(sbn_ctx* ctx, const u64 state[12], int pos, int pow_bits) {
  SBN_REQUIRE(pow_bits >= 1 && pow_bits <= 40 && pos >= 0 && pos < 8, "bad proof-of-work parameters");
  PowState st; memcpy(st.s, state, sizeof st.s);
  DevBuf<unsigned long long> best(ctx, 1);
  unsigned long long h = ~0ULL;
  ctx->upload(best, &h, 8);
  const u64 batch = 1ULL << (pow_bits + 2 > 22 ? 22 : pow_bits + 2);
  for (u64 base = 0;; base += batch) {
    KScope ks(ctx, "fri_pow");
    k_pow<<<(unsigned)(batch / 128), 128, 0, ctx->stream>>>(st, pos, pow_bits, base, best);
    LAUNCH_CHECK(ctx);
    CUDA_CHECK(cudaMemcpyAsync(&h, best, 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    if (h != ~0ULL) return h;
    if (base + batch < base) throw SbnError(SBN_ERR_INTERNAL, "proof of work failed");
  }
}

// ---- query openings ----
// Record of one query (u64 words), in proof order:
//   for each initial oracle o: evals[ncols_o], path[proof_len_o * 4]
//   for each FRI layer j:      evals[2 * arity_j], path[proof_len_j * 4]
struct GatherDesc {
  int noracles, nlayers, logn, rate_bits;
  const u64* lde[4]; int ncols[4]; int sub_coset[4]; const u64* odig[4]; int oplen[4]; size_t olevel_off[4][32];
  const u64* lleaves[8]; int larity_bits[8]; const u64* ldig[8]; int lplen[8]; size_t llevel_off[8][32];
};
__global__ void __launch_bounds__(256) k_gather_queries(GatherDesc d, const u64* __restrict__ indices, size_t record_words, u64* __restrict__ out) {
  const size_t x_index0 = indices[blockIdx.x];
  u64* rec = out + (size_t)blockIdx.x * record_words;
  const int logL = d.logn + d.rate_bits;
  const size_t N = size_t(1) << d.logn, L = size_t(1) << logL;
  size_t w = 0;
  for (int o = 0; o < d.noracles; o++) {
    // leaf x_index holds the LDE row of natural index bitrev(x_index): coset b = i mod 2^r, k = i >> r
    size_t i = bitrev32((u32)x_index0, logL);
    const size_t b = i & ((size_t(1) << d.rate_bits) - 1), k = i >> d.rate_bits;
    if (d.sub_coset[o] < 0) {
      for (int c = threadIdx.x; c < d.ncols[o]; c += blockDim.x) rec[w + c] = d.lde[o][(size_t)c * L + b * N + k];
    } else if (d.lde[o] && (size_t)d.sub_coset[o] == b) {
      for (int c = threadIdx.x; c < d.ncols[o]; c += blockDim.x) rec[w + c] = d.lde[o][(size_t)c * N + k];
    }
    w += d.ncols[o];
    for (int t = threadIdx.x; t < d.oplen[o] * 4; t += blockDim.x) {
      int lvl = t >> 2;
      size_t sib = (x_index0 >> lvl) ^ 1;
      rec[w + t] = d.odig[o][d.olevel_off[o][lvl] + sib * 4 + (t & 3)];
    }
    w += (size_t)d.oplen[o] * 4;
  }
  size_t x = x_index0;
  for (int j = 0; j < d.nlayers; j++) {
    int ab = d.larity_bits[j];
    size_t leaf = x >> ab;
    int lw = 2 << ab;
    for (int t = threadIdx.x; t < lw; t += blockDim.x) rec[w + t] = d.lleaves[j][leaf * lw + t];
    w += lw;
    for (int t = threadIdx.x; t < d.lplen[j] * 4; t += blockDim.x) {
      int lvl = t >> 2;
      size_t sib = (leaf >> lvl) ^ 1;
      rec[w + t] = d.ldig[j][d.llevel_off[j][lvl] + sib * 4 + (t & 3)];
    }
    w += (size_t)d.lplen[j] * 4;
    x = leaf;
  }
}
size_t fri_query_record_words(const std::vector<QueryOracle>& oracles, const std::vector<FriLayer*>& layers) {
  size_t w = 0;
  for (auto& o : oracles) w += o.ncols + (size_t)o.tree->proof_len() * 4;
  for (auto* l : layers) w += (size_t(2) << l->arity_bits) + (size_t)l->tree.proof_len() * 4;
  return w;
}
void fri_gather_queries(sbn_ctx* ctx, const std::vector<QueryOracle>& oracles, int logn, int rate_bits, const std::vector<FriLayer*>& layers,
                        const std::vector<u64>& indices, u64* h_out, u64* d_dst) {
  SBN_REQUIRE(oracles.size() <= 4 && layers.size() <= 8, "too many oracles / FRI layers");
  GatherDesc d; memset(&d, 0, sizeof d);
  d.noracles = (int)oracles.size(); d.nlayers = (int)layers.size(); d.logn = logn; d.rate_bits = rate_bits;
  for (size_t o = 0; o < oracles.size(); o++) {
    d.lde[o] = oracles[o].lde; d.ncols[o] = oracles[o].ncols; d.sub_coset[o] = oracles[o].sub_coset; d.odig[o] = oracles[o].tree->digests; d.oplen[o] = oracles[o].tree->proof_len();
    for (int l = 0; l < oracles[o].tree->num_levels(); l++) d.olevel_off[o][l] = oracles[o].tree->level_off[l];
  }
  for (size_t j = 0; j < layers.size(); j++) {
    d.lleaves[j] = layers[j]->leaves; d.larity_bits[j] = layers[j]->arity_bits; d.ldig[j] = layers[j]->tree.digests; d.lplen[j] = layers[j]->tree.proof_len();
    for (int l = 0; l < layers[j]->tree.num_levels(); l++) d.llevel_off[j][l] = layers[j]->tree.level_off[l];
  }
  size_t rw = fri_query_record_words(oracles, layers), nq = indices.size();
  DevBuf<u64> d_idx(ctx, nq), d_own;
  if (!d_dst) { d_own = DevBuf<u64>(ctx, nq * rw); d_dst = d_own; }
  ctx->upload(d_idx, indices.data(), nq * 8);
  KScope ks(ctx, "fri_gather_queries");
  k_gather_queries<<<(unsigned)nq, 256, 0, ctx->stream>>>(d, d_idx, rw, d_dst);
  LAUNCH_CHECK(ctx);
  if (h_out) CUDA_CHECK(cudaMemcpyAsync(h_out, d_dst, nq * rw * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();   // `indices` may be a temporary of the caller
}
