// K2: column-batched Goldilocks NTT / iNTT / coset-LDE.  Replaces plonky2's
// `PolynomialValues::ifft`, `PolynomialCoeffs::coset_fft` and `PolynomialBatch::from_values/from_coeffs`
// LDE step (external dependency), which the reference reaches through
// `prove()` (reference src/curves/g1/exp.rs:818).
//
// Layout: every polynomial is one contiguous column of N u64 in HBM (column-major batch).  A size-N
// transform is a four-step N = N1*N2 decomposition run as two kernels; each kernel keeps a
// [sub-NTT length x T columns-of-the-matrix] tile in shared memory, runs the sub-transform there as radix-16
// register passes, and touches HBM once for the read and once for the write, in >=128-byte contiguous runs.
// Natural order in, natural order out (the transposition is folded into the tile addressing).
#include "common.cuh"
#include "ntt.cuh"

__global__ void k_powers(u64* out, u64 base, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = gl_pow(base, i);
}

// Tables are built on the requesting context's stream and published (under the mutex of the table set, which the lanes of a
// batch share) only after that stream has been synchronised, so every other stream may read them without further ordering.
static const u64* pow_table_locked(sbn_ctx* ctx, u64 base, int logn) {
  SharedTables& T = *ctx->tables;
  auto key = std::make_pair(base, logn);
  auto it = T.pow_tables.find(key);
  if (it != T.pow_tables.end()) return it->second;
  size_t n = size_t(1) << logn;
  u64* p; CUDA_CHECK(cudaMalloc(&p, n * 8));
  k_powers<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, base, n);
  LAUNCH_CHECK(ctx);
  ctx->sync();
  T.pow_tables[key] = p;
  return p;
}
const u64* get_pow_table(sbn_ctx* ctx, u64 base, int logn) {
  std::lock_guard<std::mutex> g(ctx->tables->mu);
  return pow_table_locked(ctx, base, logn);
}

NttTables get_ntt_tables(sbn_ctx* ctx, int logn) {
  std::lock_guard<std::mutex> g(ctx->tables->mu);
  SharedTables& T = *ctx->tables;
  auto it = T.ntt_tables.find(logn);
  if (it != T.ntt_tables.end()) return it->second;
  NttTables t; t.logn = logn;
  u64 w = ctx->root_of_unity(logn);
  t.w_fwd = (u64*)pow_table_locked(ctx, w, logn);
  t.w_inv = (u64*)pow_table_locked(ctx, gl_inv(w), logn);
  T.ntt_tables[logn] = t;
  return t;
}

// ---- in-shared-memory transform: mixed-radix decimation in frequency, up to radix 16 per pass ----
// A pass of radix R = 2^LOGR works on blocks of length Mp (initially the whole sequence): the thread that owns group g
// reads x[g + (Mp/R) a], a < R, into registers, runs the size-R DFT there (R/2 log R butterflies, no memory traffic),
// multiplies output c by w_Mp^(g c) and writes it back to x[g + (Mp/R) c]; block c then holds the length-Mp/R
// sub-problem whose results are X[c + R k'].  Passes are in place and a thread only touches its own R slots, so one
// barrier per pass suffices (a radix-2 sweep needs one per stage and 4x the shared-memory traffic).  After the last
// pass X[k] sits at the digit-reversed position ntt_pos_to_k() inverts; the caller folds that into its global store.
// Arithmetic is lazy (arbitrary 64-bit representatives, gl.cuh); values are canonicalised when they leave the tile.
// Tile layout: s[pos * TS + t], t < T interleaved sequences, row stride TS = T + 1 (T > 1) so that both the
// "t fastest" and the "pos fastest" access patterns are bank-conflict free for 8-byte words.
template <int LOGR> __device__ __forceinline__ void dft_dif_regs(u64* v, const u64* wr /* w_R^m, m < R/2 */) {
  constexpr int R = 1 << LOGR;
#pragma unroll
  for (int st = 0; st < LOGR; st++) {
    const int half = R >> (st + 1);
#pragma unroll
    for (int blk = 0; blk < R; blk += 2 * half) {
#pragma unroll
      for (int i = 0; i < half; i++) {
        const u64 a = v[blk + i], b = v[blk + i + half];
        v[blk + i] = gl_add_nc2(a, b);
        const u64 d = gl_sub_nc2(a, b);
        v[blk + i + half] = (i == 0) ? d : gl_mul_nc(d, wr[i << st]);   // w_(2 half)^i = w_R^(i R / (2 half))
      }
    }
  }
}
__device__ __forceinline__ int ntt_row_stride(int logT) { return (1 << logT) + (logT > 0 ? 1 : 0); }
template <int LOGR> __device__ __forceinline__ void smem_pass(u64* s, const u64* tw, int l, int lp, int logT) {
  constexpr int R = 1 << LOGR;
  const int TS = ntt_row_stride(logT);
  const int total = (1 << (l - LOGR)) << logT;
  const int gsh = lp - LOGR;                 // log2(groups per block) = log2(element stride inside a group)
  u64 wr[R > 1 ? R / 2 : 1];
#pragma unroll
  for (int m = 0; m < R / 2; m++) wr[m] = tw[m << (l - LOGR)];
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int t = idx & ((1 << logT) - 1), ga = idx >> logT;
    const int g = ga & ((1 << gsh) - 1), base = ((ga >> gsh) << lp) + g;
    u64 v[R];
#pragma unroll
    for (int a = 0; a < R; a++) v[a] = s[(base + (a << gsh)) * TS + t];
    dft_dif_regs<LOGR>(v, wr);
#pragma unroll
    for (int c = 0; c < R; c++) {
      u64 y = v[bitrev32(c, LOGR)];
      if (c > 0 && gsh > 0) y = gl_mul_nc(y, tw[(g * c) << (l - lp)]);   // w_Mp^(g c); the last pass has g == 0
      s[(base + (c << gsh)) * TS + t] = y;
    }
  }
}
// radix schedule: 16 while at least 4 bits remain, then the remainder (1..3 bits) in one pass
__device__ __forceinline__ void smem_ntt(u64* s, const u64* tw, int l, int logT) {
  int lp = l;
  while (lp > 0) {
    const int r = lp >= 4 ? 4 : lp;
    if (r == 4) smem_pass<4>(s, tw, l, lp, logT);
    else if (r == 3) smem_pass<3>(s, tw, l, lp, logT);
    else if (r == 2) smem_pass<2>(s, tw, l, lp, logT);
    else smem_pass<1>(s, tw, l, lp, logT);
    __syncthreads();
    lp -= r;
  }
}
template <int L, int LP> __device__ __forceinline__ void smem_ntt_ct(u64* s, const u64* tw, int logT) {   // the same with L known
  if constexpr (LP > 0) {
    constexpr int r = LP >= 4 ? 4 : LP;
    smem_pass<r>(s, tw, L, LP, logT);
    __syncthreads();
    smem_ntt_ct<L, LP - r>(s, tw, logT);
  }
}
HD int ntt_pos_to_k(int pos, int l) {   // position in the tile after smem_ntt -> frequency index
  int k = 0, shift = 0, lp = l;
  while (lp > 0) {
    const int r = lp >= 4 ? 4 : lp;
    k |= ((pos >> (lp - r)) & ((1 << r) - 1)) << shift;
    shift += r; lp -= r;
  }
  return k;
}

// Four-step transform N = N1 * N2 (n = N2 n1 + n2, k = k1 + N1 k2) as two kernels, sub-transform size a template
// parameter so that every loop has a compile-time trip count.  A block keeps 2^L x T elements in shared memory
// (T = 2^LOGT columns of the N1 x N2 matrix); each thread moves its elements in batches of 8 independent loads.
//   pass 1: tile of T consecutive n2: length-N1 transform over n1, then times F[k1][n2], in -> tmp (same [k1 N2 + n2] layout)
//   pass 2: tile of T consecutive k1: length-N2 transform over n2 (contiguous), tmp -> X[k1 + N1 k2]
// F[k1][n2] = w_N^(n2 k1) * c^n2 * scale carries the four-step twiddle, the n2 part of a coset shift c (evaluation on
// c<w>: input times c^n, c^n = (c^N2)^n1 c^n2; the n1 part is the row factor P1[n1] applied on load) and the 1/N of an
// inverse transform, so those cost no instructions; it is read with the same coalesced index the result is stored at.
__host__ __device__ constexpr int ntt_logT_of(int l) { return 13 - l > 5 ? 5 : (13 - l < 0 ? 0 : 13 - l); }
constexpr int NTT_THREADS = 256, NTT_BATCH = 8;

template <int L1, bool PRE> __global__ void __launch_bounds__(NTT_THREADS) k_ntt_pass1(const u64* __restrict__ in, size_t in_stride, u64* __restrict__ tmp,
    const u64* __restrict__ W, const u64* __restrict__ F, const u64* __restrict__ P1, int l2) {
  extern __shared__ u64 smem[];
  constexpr int LOGT = ntt_logT_of(L1), N1 = 1 << L1, T = 1 << LOGT, TS = T + (LOGT > 0 ? 1 : 0), TOTAL = N1 << LOGT;
  static_assert(TOTAL % (NTT_THREADS * NTT_BATCH) == 0, "tile is a whole number of load batches");
  const size_t N = size_t(1) << (L1 + l2);
  u64* s = smem;
  u64* tw = smem + (size_t)N1 * TS;
  const u64* src = in + (size_t)blockIdx.y * in_stride + ((size_t)blockIdx.x << LOGT);
  u64* dst = tmp + (size_t)blockIdx.y * N + ((size_t)blockIdx.x << LOGT);
  const u64* Fb = F + ((size_t)blockIdx.x << LOGT);
  for (int j = threadIdx.x; j < N1; j += NTT_THREADS) tw[j] = W[(size_t)j << l2];
#pragma unroll 1
  for (int base = threadIdx.x; base < TOTAL; base += NTT_THREADS * NTT_BATCH) {
    u64 v[NTT_BATCH], pf[NTT_BATCH];
#pragma unroll
    for (int u = 0; u < NTT_BATCH; u++) {
      const int idx = base + u * NTT_THREADS, n1 = idx >> LOGT, t = idx & (T - 1);
      v[u] = src[((size_t)n1 << l2) + t];
      if (PRE) pf[u] = P1[n1];
    }
#pragma unroll
    for (int u = 0; u < NTT_BATCH; u++) {
      const int idx = base + u * NTT_THREADS, n1 = idx >> LOGT, t = idx & (T - 1);
      s[n1 * TS + t] = PRE ? gl_mul_nc(v[u], pf[u]) : v[u];
    }
  }
  __syncthreads();
  smem_ntt_ct<L1, L1>(s, tw, LOGT);
#pragma unroll 1
  for (int base = threadIdx.x; base < TOTAL; base += NTT_THREADS * NTT_BATCH) {
    u64 f[NTT_BATCH];
#pragma unroll
    for (int u = 0; u < NTT_BATCH; u++) {
      const int idx = base + u * NTT_THREADS, pos = idx >> LOGT, t = idx & (T - 1);
      f[u] = Fb[((size_t)ntt_pos_to_k(pos, L1) << l2) + t];
    }
#pragma unroll
    for (int u = 0; u < NTT_BATCH; u++) {
      const int idx = base + u * NTT_THREADS, pos = idx >> LOGT, t = idx & (T - 1);
      dst[((size_t)ntt_pos_to_k(pos, L1) << l2) + t] = gl_mul_nc(s[pos * TS + t], f[u]);   // lazy representative: pass 2 reduces it again
    }
  }
}

template <int L2> __global__ void __launch_bounds__(NTT_THREADS) k_ntt_pass2(const u64* __restrict__ tmp, u64* __restrict__ out, size_t out_stride,
    const u64* __restrict__ W, const u64* __restrict__ postscale, int l1) {
  extern __shared__ u64 smem[];
  constexpr int LOGT = ntt_logT_of(L2), N2 = 1 << L2, T = 1 << LOGT, TS = T + (LOGT > 0 ? 1 : 0), TOTAL = N2 << LOGT;
  static_assert(TOTAL % (NTT_THREADS * NTT_BATCH) == 0, "tile is a whole number of load batches");
  const size_t N = size_t(1) << (l1 + L2);
  u64* s = smem;
  u64* tw = smem + (size_t)N2 * TS;
  const size_t k1_0 = (size_t)blockIdx.x << LOGT;
  const u64* src = tmp + (size_t)blockIdx.y * N + (k1_0 << L2);
  u64* dst = out + (size_t)blockIdx.y * out_stride + k1_0;
  for (int j = threadIdx.x; j < N2; j += NTT_THREADS) tw[j] = W[(size_t)j << l1];
#pragma unroll 1
  for (int base = threadIdx.x; base < TOTAL; base += NTT_THREADS * NTT_BATCH) {
    u64 v[NTT_BATCH];
#pragma unroll
    for (int u = 0; u < NTT_BATCH; u++) v[u] = src[base + u * NTT_THREADS];        // the T rows of the tile are one contiguous run
#pragma unroll
    for (int u = 0; u < NTT_BATCH; u++) {
      const int idx = base + u * NTT_THREADS, t = idx >> L2, n2 = idx & (N2 - 1);
      s[n2 * TS + t] = v[u];
    }
  }
  __syncthreads();
  smem_ntt_ct<L2, L2>(s, tw, LOGT);
  const bool post = postscale != nullptr;
#pragma unroll 1
  for (int base = threadIdx.x; base < TOTAL; base += NTT_THREADS * NTT_BATCH) {
#pragma unroll
    for (int u = 0; u < NTT_BATCH; u++) {
      const int idx = base + u * NTT_THREADS, pos = idx >> LOGT, t = idx & (T - 1);
      const size_t k = ((size_t)ntt_pos_to_k(pos, L2) << l1) + t;
      u64 v = s[pos * TS + t];
      if (post) v = gl_mul_nc(v, postscale[k1_0 + k]);
      dst[k] = gl_canon(v);
    }
  }
}
__global__ void k_fourstep_table(u64* __restrict__ out, u64 w, u64 c, u64 scale, int l2, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 k1 = i >> l2, n2 = i & ((size_t(1) << l2) - 1);
  out[i] = gl_mul(gl_mul(gl_pow(w, n2 * k1), gl_pow(c, n2)), scale);
}
// (logn, inverse, coset base) -> F table (device, N entries), built once per context
static const u64* get_fourstep_table(sbn_ctx* ctx, int logn, bool inverse, u64 c) {
  std::lock_guard<std::mutex> g(ctx->tables->mu);
  SharedTables& T = *ctx->tables;
  auto key = std::make_tuple(logn, inverse, c);
  auto it = T.fourstep_tables.find(key);
  if (it != T.fourstep_tables.end()) return it->second;
  const size_t n = size_t(1) << logn;
  const int l1 = logn / 2, l2 = logn - l1;
  u64 w = ctx->root_of_unity(logn);
  if (inverse) w = gl_inv(w);
  u64* p; CUDA_CHECK(cudaMalloc(&p, n * 8));
  k_fourstep_table<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, w, c ? c : 1, inverse ? gl_inv((u64)n) : 1, l2, n);
  LAUNCH_CHECK(ctx);
  ctx->sync();
  T.fourstep_tables[key] = p;
  return p;
}

// Whole transform in one block (N <= 2048): one column per block.
__global__ void k_ntt_small(const u64* __restrict__ in, size_t in_stride, u64* __restrict__ out, size_t out_stride,
                            const u64* __restrict__ W, const u64* __restrict__ prescale, const u64* __restrict__ postscale,
                            u64 scale, int logn) {
  extern __shared__ u64 smem[];
  const int N = 1 << logn;
  u64* s = smem;
  u64* tw = smem + N;
  const u64* src = in + (size_t)blockIdx.x * in_stride;
  u64* dst = out + (size_t)blockIdx.x * out_stride;
  for (int j = threadIdx.x; j < N; j += blockDim.x) tw[j] = W[j];
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    u64 v = src[n];
    if (prescale) v = gl_mul_nc(v, prescale[n]);
    s[n] = v;
  }
  __syncthreads();
  smem_ntt(s, tw, logn, 0);
  for (int pos = threadIdx.x; pos < N; pos += blockDim.x) {
    int k = ntt_pos_to_k(pos, logn);
    u64 v = s[pos];
    if (scale != 1) v = gl_mul_nc(v, scale);
    if (postscale) v = gl_mul_nc(v, postscale[k]);
    dst[k] = gl_canon(v);
  }
}

// shared memory of a pass with sub-transform size 2^l: the tile (row stride T + 1) and the 2^l twiddles; 131 KB at l = 12
static constexpr int ntt_tile_bytes(int l) { return (int)((((size_t)1 << l) * (((size_t)1 << ntt_logT_of(l)) + (ntt_logT_of(l) > 0 ? 1 : 0)) + ((size_t)1 << l)) * 8); }
template <int L> static void launch_pass1(sbn_ctx* ctx, dim3 grid, size_t smem, bool pre, const u64* in, size_t in_stride, u64* tmp, const u64* W, const u64* F, const u64* P1, int l2) {
  // the shared-memory opt-in is a per-device function attribute: remembered per context (a context is bound to one device)
  if (!(ctx->ntt_attr_mask & (1u << L))) {
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass1<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ntt_tile_bytes(L)));
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass1<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ntt_tile_bytes(L)));
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass1<L, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass1<L, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    ctx->ntt_attr_mask |= 1u << L;
  }
  if (pre) k_ntt_pass1<L, true><<<grid, NTT_THREADS, smem, ctx->stream>>>(in, in_stride, tmp, W, F, P1, l2);
  else k_ntt_pass1<L, false><<<grid, NTT_THREADS, smem, ctx->stream>>>(in, in_stride, tmp, W, F, P1, l2);
}
template <int L> static void launch_pass2(sbn_ctx* ctx, dim3 grid, size_t smem, const u64* tmp, u64* out, size_t out_stride, const u64* W, const u64* postscale, int l1) {
  if (!(ctx->ntt_attr_mask & (1u << (16 + L)))) {
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass2<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, ntt_tile_bytes(L)));
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass2<L>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    ctx->ntt_attr_mask |= 1u << (16 + L);
  }
  k_ntt_pass2<L><<<grid, NTT_THREADS, smem, ctx->stream>>>(tmp, out, out_stride, W, postscale, l1);
}
#define NTT_DISPATCH_L(l, CALL) \
  switch (l) { case 6: { constexpr int L = 6; CALL; } break; case 7: { constexpr int L = 7; CALL; } break; case 8: { constexpr int L = 8; CALL; } break; \
    case 9: { constexpr int L = 9; CALL; } break; case 10: { constexpr int L = 10; CALL; } break; case 11: { constexpr int L = 11; CALL; } break; \
    case 12: { constexpr int L = 12; CALL; } break; case 13: { constexpr int L = 13; CALL; } break; default: SBN_REQUIRE(false, "ntt: unsupported sub-transform size"); }

// Transform of `ncols` columns.  pre_base c != 0: evaluate on the coset c<w> (input times c^n); postscale: optional table
// the output is multiplied by (output index k); inverse: w^-1 and 1/N.
void ntt_batch(sbn_ctx* ctx, const u64* in, size_t in_stride, u64* out, size_t out_stride, int ncols, int logn, bool inverse,
               u64 pre_base, const u64* postscale) {
  if (ncols <= 0) return;
  SBN_REQUIRE(logn >= 1 && logn <= 26, "ntt: unsupported size");
  if (pre_base == 1) pre_base = 0;
  const NttTables tb = get_ntt_tables(ctx, logn);
  const u64* W = inverse ? tb.w_inv : tb.w_fwd;
  const size_t N = size_t(1) << logn;
  if (logn <= 11) {
    const u64 scale = inverse ? gl_inv((u64)N) : 1;
    const u64* prescale = pre_base ? get_pow_table(ctx, pre_base, logn) : nullptr;
    size_t smem = 2 * N * 8;
    int threads = N >= 4096 ? 256 : (N >= 512 ? (int)(N / 16) : 32);
    KScope ks(ctx, "ntt_small");
    k_ntt_small<<<ncols, threads, smem, ctx->stream>>>(in, in_stride, out, out_stride, W, prescale, postscale, scale, logn);
    LAUNCH_CHECK(ctx);
    return;
  }
  const int l1 = logn / 2, l2 = logn - l1;
  const int logT1 = ntt_logT_of(l1), logT2 = ntt_logT_of(l2);
  const u64* F = get_fourstep_table(ctx, logn, inverse, pre_base);
  const u64* P1 = pre_base ? get_pow_table(ctx, gl_exp_pow2(pre_base, l2), l1) : nullptr;   // (c^N2)^n1
  size_t smem1 = ntt_tile_bytes(l1), smem2 = ntt_tile_bytes(l2);
  // Column chunks sized so the intermediate stays L2-resident between the two passes.
  size_t chunk = (size_t(48) << 20) / (N * 8);
  if (chunk < 1) chunk = 1;
  if (chunk > (size_t)ncols) chunk = ncols;
  if (chunk > 32768) chunk = 32768;
  {  // round the chunk to whole waves of resident blocks (3 blocks of 256 threads per SM): a 1.7-wave launch costs 2 waves
    const size_t slots = (size_t)ctx->num_sms * 3, per_col = size_t(1) << (l2 - logT1);
    size_t waves = (chunk * per_col + slots / 2) / slots;
    if (waves >= 1 && waves * slots / per_col >= 1) chunk = std::min<size_t>(waves * slots / per_col, 32768);
    if (chunk > (size_t)ncols) chunk = ncols;
  }
  DevBuf<u64> tmp(ctx, chunk * N);
  for (size_t c0 = 0; c0 < (size_t)ncols; c0 += chunk) {
    unsigned nc = (unsigned)std::min(chunk, (size_t)ncols - c0);
    dim3 g1((unsigned)(1u << (l2 - logT1)), nc), g2((unsigned)(1u << (l1 - logT2)), nc);
    { KScope ks(ctx, "ntt_pass1");
    NTT_DISPATCH_L(l1, (launch_pass1<L>(ctx, g1, smem1, pre_base != 0, in + c0 * in_stride, in_stride, tmp, W, F, P1, l2)));
    LAUNCH_CHECK(ctx); }
    KScope ks2(ctx, "ntt_pass2");
    NTT_DISPATCH_L(l2, (launch_pass2<L>(ctx, g2, smem2, tmp, out + c0 * out_stride, out_stride, W, postscale, l1)));
    LAUNCH_CHECK(ctx);
  }
}

// values (N) -> coefficients (N) for every column:  PolynomialValues::ifft
void intt_columns(sbn_ctx* ctx, const u64* values, u64* coeffs, int ncols, int logn) {
  size_t N = size_t(1) << logn;
  ntt_batch(ctx, values, N, coeffs, N, ncols, logn, true, 0, nullptr);
}

// coefficients (N) -> LDE values on the coset shift*<w_{N*2^r}>, laid out lde[col][b][k] with natural
// index i = k*2^r + b, i.e. sub-coset b is the size-N NTT of c_n * (shift * w_L^b)^n.
// (PolynomialCoeffs::lde(rate_bits).coset_fft(F::coset_shift()), without the zero-padded half.)
void lde_columns(sbn_ctx* ctx, const u64* coeffs, u64* lde, int ncols, int logn, int rate_bits) {
  size_t N = size_t(1) << logn;
  int R = 1 << rate_bits;
  u64 wL = ctx->root_of_unity(logn + rate_bits);
  for (int b = 0; b < R; b++) {
    u64 sb = gl_mul(ctx->coset_shift(), gl_pow(wL, b));
    ntt_batch(ctx, coeffs, N, lde + (size_t)b * N, N * R, ncols, logn, false, sb, nullptr);
  }
}

// One sub-coset b of the LDE: out[col][k] = value at natural LDE index k 2^r + b (column stride N).
void lde_sub_coset(sbn_ctx* ctx, const u64* coeffs, u64* out, int ncols, int logn, int rate_bits, int b) {
  size_t N = size_t(1) << logn;
  const u64 sb = gl_mul(ctx->coset_shift(), gl_pow(ctx->root_of_unity(logn + rate_bits), (u64)b));
  ntt_batch(ctx, coeffs, N, out, N, ncols, logn, false, sb, nullptr);
}

// ---- one LDE class (intra-proof sharding) ----
// The leaves under Merkle cap entries [r 2^(cap_height - m), (r + 1) 2^(cap_height - m)) are the LDE points of natural index
// i = rho (mod G), G = 2^m, rho = bitrev_m(r): the coset s w_L^rho <w_(L/G)>.  lde_class evaluates every column on that
// coset only: out[col][b'][k'] with i = rho + G (k' B' + b'), B' = max(1, R / G) -- the layout lde_columns would give for a
// domain of L / G points, so the leaf-hash, quotient and query kernels run on it unchanged.
//   G <= R: the class is R / G whole sub-cosets b = rho + G b' of the size-N transform;
//   G >  R: L / G < N points: reduce the polynomial mod x^(L/G) - a, a = (s w_L^rho)^(L/G), then one size-L/G coset transform.
__global__ void k_fold_mod_binomial(const u64* __restrict__ coeffs, size_t N, u64* __restrict__ out, size_t Lp, int f, u64 a) {
  const size_t n = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (n >= Lp) return;
  const u64* c = coeffs + (size_t)blockIdx.y * N + n;
  u64 acc = 0, ap = 1;
  for (int j = 0; j < f; j++) { acc = gl_add(acc, gl_mul(c[(size_t)j * Lp], ap)); ap = gl_mul(ap, a); }
  out[(size_t)blockIdx.y * Lp + n] = acc;
}
void lde_class(sbn_ctx* ctx, const u64* coeffs, u64* out, int ncols, int logn, int rate_bits, int m, u32 rho) {
  const size_t N = size_t(1) << logn;
  const int R = 1 << rate_bits, G = 1 << m;
  const u64 wL = ctx->root_of_unity(logn + rate_bits);
  if (G <= R) {
    const int Bp = R / G;
    for (int bp = 0; bp < Bp; bp++) {
      const u64 sb = gl_mul(ctx->coset_shift(), gl_pow(wL, rho + (u64)G * bp));
      ntt_batch(ctx, coeffs, N, out + (size_t)bp * N, N * Bp, ncols, logn, false, sb, nullptr);
    }
    return;
  }
  const int logLp = logn + rate_bits - m, f = G / R;
  const size_t Lp = size_t(1) << logLp;
  const u64 c = gl_mul(ctx->coset_shift(), gl_pow(wL, rho));
  DevBuf<u64> folded(ctx, (size_t)ncols * Lp);
  { KScope ks(ctx, "lde_fold");
    k_fold_mod_binomial<<<dim3((unsigned)((Lp + 255) / 256), (unsigned)ncols), 256, 0, ctx->stream>>>(coeffs, N, folded, Lp, f, gl_exp_pow2(c, logLp));
    LAUNCH_CHECK(ctx); }
  ntt_batch(ctx, folded, Lp, out, Lp, ncols, logLp, false, c, nullptr);
}
