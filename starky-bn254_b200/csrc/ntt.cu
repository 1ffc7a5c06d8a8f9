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

const u64* get_pow_table(sbn_ctx* ctx, u64 base, int logn) {
  auto key = std::make_pair(base, logn);
  auto it = ctx->pow_tables.find(key);
  if (it != ctx->pow_tables.end()) return it->second;
  size_t n = size_t(1) << logn;
  u64* p; CUDA_CHECK(cudaMalloc(&p, n * 8));
  k_powers<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, base, n);
  LAUNCH_CHECK(ctx);
  ctx->pow_tables[key] = p;
  return p;
}

const NttTables& get_ntt_tables(sbn_ctx* ctx, int logn) {
  auto it = ctx->ntt_tables.find(logn);
  if (it != ctx->ntt_tables.end()) return it->second;
  NttTables t; t.logn = logn;
  u64 w = gl_root_of_unity(logn);
  t.w_fwd = (u64*)get_pow_table(ctx, w, logn);
  t.w_inv = (u64*)get_pow_table(ctx, gl_inv(w), logn);
  ctx->ntt_tables[logn] = t;
  return ctx->ntt_tables[logn];
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
HD int ntt_pos_to_k(int pos, int l) {   // position in the tile after smem_ntt -> frequency index
  int k = 0, shift = 0, lp = l;
  while (lp > 0) {
    const int r = lp >= 4 ? 4 : lp;
    k |= ((pos >> (lp - r)) & ((1 << r) - 1)) << shift;
    shift += r; lp -= r;
  }
  return k;
}

// Pass 1 of the four-step transform: for a tile of T consecutive n2, length-N1 NTT over n1
// (elements N2 apart), then the inter-step twiddle w_N^(n2*k1).  in -> tmp, same [N2*k1 + n2] layout.
__global__ void __launch_bounds__(256) k_ntt_pass1(const u64* __restrict__ in, size_t in_stride, u64* __restrict__ tmp, const u64* __restrict__ W,
                                                   const u64* __restrict__ prescale, int l1, int l2, int logT) {
  extern __shared__ u64 smem[];
  const int N1 = 1 << l1, T = 1 << logT, TS = ntt_row_stride(logT);
  const size_t N = size_t(1) << (l1 + l2);
  u64* s = smem;
  u64* tw = smem + (size_t)N1 * TS;
  const u64* src = in + (size_t)blockIdx.y * in_stride;
  u64* dst = tmp + (size_t)blockIdx.y * N;
  const int n2_0 = blockIdx.x << logT;
  for (int j = threadIdx.x; j < N1; j += blockDim.x) tw[j] = W[(size_t)j << l2];
  for (int idx = threadIdx.x; idx < (N1 << logT); idx += blockDim.x) {
    int n1 = idx >> logT, t = idx & (T - 1);
    size_t n = ((size_t)n1 << l2) + n2_0 + t;
    u64 v = src[n];
    if (prescale) v = gl_mul_nc(v, prescale[n]);
    s[n1 * TS + t] = v;
  }
  __syncthreads();
  smem_ntt(s, tw, l1, logT);
  for (int idx = threadIdx.x; idx < (N1 << logT); idx += blockDim.x) {
    int pos = idx >> logT, t = idx & (T - 1);
    int k1 = ntt_pos_to_k(pos, l1);
    size_t n2 = n2_0 + t;
    u64 v = gl_mul_nc(s[pos * TS + t], W[n2 * k1]);
    dst[((size_t)k1 << l2) + n2] = v;     // lazy representative: pass 2 reduces it again
  }
}

// Pass 2: for a tile of T consecutive k1, length-N2 NTT over n2 (contiguous), output X[k1 + N1*k2].
__global__ void __launch_bounds__(256) k_ntt_pass2(const u64* __restrict__ tmp, u64* __restrict__ out, size_t out_stride, const u64* __restrict__ W,
                                                   const u64* __restrict__ postscale, u64 scale, int l1, int l2, int logT) {
  extern __shared__ u64 smem[];
  const int N2 = 1 << l2, T = 1 << logT, TS = ntt_row_stride(logT);
  const size_t N = size_t(1) << (l1 + l2);
  u64* s = smem;
  u64* tw = smem + (size_t)N2 * TS;
  const u64* src = tmp + (size_t)blockIdx.y * N;
  u64* dst = out + (size_t)blockIdx.y * out_stride;
  const int k1_0 = blockIdx.x << logT;
  for (int j = threadIdx.x; j < N2; j += blockDim.x) tw[j] = W[(size_t)j << l1];
  for (int idx = threadIdx.x; idx < (N2 << logT); idx += blockDim.x) {
    int t = idx >> l2, n2 = idx & (N2 - 1);
    s[n2 * TS + t] = src[((size_t)(k1_0 + t) << l2) + n2];
  }
  __syncthreads();
  smem_ntt(s, tw, l2, logT);
  for (int idx = threadIdx.x; idx < (N2 << logT); idx += blockDim.x) {
    int pos = idx >> logT, t = idx & (T - 1);
    int k2 = ntt_pos_to_k(pos, l2);
    size_t k = (size_t)k1_0 + t + ((size_t)k2 << l1);
    u64 v = s[pos * TS + t];
    if (scale != 1) v = gl_mul_nc(v, scale);
    if (postscale) v = gl_mul_nc(v, postscale[k]);
    dst[k] = gl_canon(v);
  }
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

static int pick_logT(int l) {
  // tile = 2^l * T * 8 bytes <= 64 KiB, T <= 32
  int logT = 13 - l;
  if (logT > 5) logT = 5;
  if (logT < 0) logT = 0;
  return logT;
}

void ntt_batch(sbn_ctx* ctx, const u64* in, size_t in_stride, u64* out, size_t out_stride, int ncols, int logn, bool inverse,
               const u64* prescale, const u64* postscale) {
  if (ncols <= 0) return;
  SBN_REQUIRE(logn >= 1 && logn <= 24, "ntt: unsupported size");
  const NttTables& tb = get_ntt_tables(ctx, logn);
  const u64* W = inverse ? tb.w_inv : tb.w_fwd;
  const size_t N = size_t(1) << logn;
  u64 scale = inverse ? gl_inv((u64)N) : 1;
  if (logn <= 11) {
    size_t smem = 2 * N * 8;
    int threads = N >= 4096 ? 256 : (N >= 512 ? (int)(N / 16) : 32);
    KScope ks(ctx, "ntt_small");
    k_ntt_small<<<ncols, threads, smem, ctx->stream>>>(in, in_stride, out, out_stride, W, prescale, postscale, scale, logn);
    LAUNCH_CHECK(ctx);
    return;
  }
  int l1 = logn / 2, l2 = logn - l1;
  int logT1 = pick_logT(l1), logT2 = pick_logT(l2);
  if (logT1 > l2) logT1 = l2;
  if (logT2 > l1) logT2 = l1;
  auto tile_bytes = [](int l, int logT) { return ((size_t(1) << l) * ((size_t(1) << logT) + (logT > 0 ? 1 : 0)) + (size_t(1) << l)) * 8; };
  size_t smem1 = tile_bytes(l1, logT1), smem2 = tile_bytes(l2, logT2);
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass1, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
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
    k_ntt_pass1<<<g1, 256, smem1, ctx->stream>>>(in + c0 * in_stride, in_stride, tmp, W, prescale, l1, l2, logT1);
    LAUNCH_CHECK(ctx); }
    KScope ks2(ctx, "ntt_pass2");
    k_ntt_pass2<<<g2, 256, smem2, ctx->stream>>>(tmp, out + c0 * out_stride, out_stride, W, postscale, scale, l1, l2, logT2);
    LAUNCH_CHECK(ctx);
  }
}

// values (N) -> coefficients (N) for every column:  PolynomialValues::ifft
void intt_columns(sbn_ctx* ctx, const u64* values, u64* coeffs, int ncols, int logn) {
  size_t N = size_t(1) << logn;
  ntt_batch(ctx, values, N, coeffs, N, ncols, logn, true, nullptr, nullptr);
}

// coefficients (N) -> LDE values on the coset shift*<w_{N*2^r}>, laid out lde[col][b][k] with natural
// index i = k*2^r + b, i.e. sub-coset b is the size-N NTT of c_n * (shift * w_L^b)^n.
// (PolynomialCoeffs::lde(rate_bits).coset_fft(F::coset_shift()), without the zero-padded half.)
void lde_columns(sbn_ctx* ctx, const u64* coeffs, u64* lde, int ncols, int logn, int rate_bits) {
  size_t N = size_t(1) << logn;
  int R = 1 << rate_bits;
  u64 wL = gl_root_of_unity(logn + rate_bits);
  for (int b = 0; b < R; b++) {
    u64 sb = gl_mul(GL_MULT_GENERATOR, gl_pow(wL, b));
    const u64* pre = get_pow_table(ctx, sb, logn);
    ntt_batch(ctx, coeffs, N, lde + (size_t)b * N, N * R, ncols, logn, false, pre, nullptr);
  }
}
