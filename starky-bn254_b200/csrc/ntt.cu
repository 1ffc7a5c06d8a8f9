// K2: column-batched Goldilocks NTT / iNTT / coset-LDE.  Replaces plonky2's
// `PolynomialValues::ifft`, `PolynomialCoeffs::coset_fft` and `PolynomialBatch::from_values/from_coeffs`
// LDE step (external dependency), which the reference reaches through
// `prove()` (reference src/curves/g1/exp.rs:818).
//
// Layout: every polynomial is one contiguous column of N u64 in HBM (column-major batch).  A size-N
// transform is a four-step N = N1*N2 decomposition run as two kernels; each kernel keeps a
// [sub-NTT length x T columns-of-the-matrix] tile in shared memory, runs all of its radix-2 stages
// there, and touches HBM once for the read and once for the write, in >=128-byte contiguous runs.
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

// Radix-2 DIT over `T = 1 << logT` interleaved sequences of length M = 1 << l held in shared memory as
// s[m * T + t]; input in bit-reversed order, output natural.  tw[j] = w_M^j, j < M/2.
__device__ __forceinline__ void smem_ntt(u64* s, const u64* tw, int l, int logT) {
  const int M = 1 << l;
  const int total = (M >> 1) << logT;
  for (int stage = 1; stage <= l; stage++) {
    const int half = 1 << (stage - 1);
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      int t = idx & ((1 << logT) - 1);
      int b = idx >> logT;
      int j = b & (half - 1);
      int i0 = ((b >> (stage - 1)) << stage) + j;
      int i1 = i0 + half;
      u64 w = tw[j << (l - stage)];
      u64 u = s[(i0 << logT) + t];
      u64 v = gl_mul(s[(i1 << logT) + t], w);
      s[(i0 << logT) + t] = gl_add(u, v);
      s[(i1 << logT) + t] = gl_sub(u, v);
    }
    __syncthreads();
  }
}

// Pass 1 of the four-step transform: for a tile of T consecutive n2, length-N1 NTT over n1
// (elements N2 apart), then the inter-step twiddle w_N^(n2*k1).  in -> tmp, same [N2*k1 + n2] layout.
__global__ void k_ntt_pass1(const u64* __restrict__ in, size_t in_stride, u64* __restrict__ tmp, const u64* __restrict__ W,
                            const u64* __restrict__ prescale, int l1, int l2, int logT) {
  extern __shared__ u64 smem[];
  const int N1 = 1 << l1, T = 1 << logT;
  const size_t N = size_t(1) << (l1 + l2);
  u64* s = smem;
  u64* tw = smem + ((size_t)N1 << logT);
  const u64* src = in + (size_t)blockIdx.y * in_stride;
  u64* dst = tmp + (size_t)blockIdx.y * N;
  const int n2_0 = blockIdx.x << logT;
  for (int j = threadIdx.x; j < (N1 >> 1); j += blockDim.x) tw[j] = W[(size_t)j << l2];
  for (int idx = threadIdx.x; idx < (N1 << logT); idx += blockDim.x) {
    int n1 = idx >> logT, t = idx & (T - 1);
    size_t n = ((size_t)n1 << l2) + n2_0 + t;
    u64 v = src[n];
    if (prescale) v = gl_mul(v, prescale[n]);
    s[((size_t)bitrev32(n1, l1) << logT) + t] = v;
  }
  __syncthreads();
  smem_ntt(s, tw, l1, logT);
  for (int idx = threadIdx.x; idx < (N1 << logT); idx += blockDim.x) {
    int k1 = idx >> logT, t = idx & (T - 1);
    size_t n2 = n2_0 + t;
    u64 v = gl_mul(s[((size_t)k1 << logT) + t], W[n2 * k1]);
    dst[((size_t)k1 << l2) + n2] = v;
  }
}

// Pass 2: for a tile of T consecutive k1, length-N2 NTT over n2 (contiguous), output X[k1 + N1*k2].
__global__ void k_ntt_pass2(const u64* __restrict__ tmp, u64* __restrict__ out, size_t out_stride, const u64* __restrict__ W,
                            const u64* __restrict__ postscale, u64 scale, int l1, int l2, int logT) {
  extern __shared__ u64 smem[];
  const int N2 = 1 << l2, T = 1 << logT;
  const size_t N = size_t(1) << (l1 + l2);
  u64* s = smem;
  u64* tw = smem + ((size_t)N2 << logT);
  const u64* src = tmp + (size_t)blockIdx.y * N;
  u64* dst = out + (size_t)blockIdx.y * out_stride;
  const int k1_0 = blockIdx.x << logT;
  for (int j = threadIdx.x; j < (N2 >> 1); j += blockDim.x) tw[j] = W[(size_t)j << l1];
  for (int idx = threadIdx.x; idx < (N2 << logT); idx += blockDim.x) {
    int t = idx >> l2, n2 = idx & (N2 - 1);
    u64 v = src[((size_t)(k1_0 + t) << l2) + n2];
    s[((size_t)bitrev32(n2, l2) << logT) + t] = v;
  }
  __syncthreads();
  smem_ntt(s, tw, l2, logT);
  for (int idx = threadIdx.x; idx < (N2 << logT); idx += blockDim.x) {
    int k2 = idx >> logT, t = idx & (T - 1);
    size_t k = (size_t)k1_0 + t + ((size_t)k2 << l1);
    u64 v = s[((size_t)k2 << logT) + t];
    if (scale != 1) v = gl_mul(v, scale);
    if (postscale) v = gl_mul(v, postscale[k]);
    dst[k] = v;
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
  for (int j = threadIdx.x; j < (N >> 1); j += blockDim.x) tw[j] = W[j];
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    u64 v = src[n];
    if (prescale) v = gl_mul(v, prescale[n]);
    s[bitrev32(n, logn)] = v;
  }
  __syncthreads();
  smem_ntt(s, tw, logn, 0);
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    u64 v = s[k];
    if (scale != 1) v = gl_mul(v, scale);
    if (postscale) v = gl_mul(v, postscale[k]);
    dst[k] = v;
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
    size_t smem = (N + N / 2) * 8;
    int threads = N >= 512 ? 256 : (N >= 64 ? (int)(N / 2) : 32);
    KScope ks(ctx, "ntt_small");
    k_ntt_small<<<ncols, threads, smem, ctx->stream>>>(in, in_stride, out, out_stride, W, prescale, postscale, scale, logn);
    LAUNCH_CHECK(ctx);
    return;
  }
  int l1 = logn / 2, l2 = logn - l1;
  int logT1 = pick_logT(l1), logT2 = pick_logT(l2);
  if (logT1 > l2) logT1 = l2;
  if (logT2 > l1) logT2 = l1;
  size_t smem1 = (((size_t)1 << l1) << logT1) * 8 + ((size_t)1 << l1) * 4;
  size_t smem2 = (((size_t)1 << l2) << logT2) * 8 + ((size_t)1 << l2) * 4;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass1, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set = true;
  }
  // Column chunks sized so the intermediate stays L2-resident between the two passes.
  size_t chunk = (size_t(48) << 20) / (N * 8);
  if (chunk < 1) chunk = 1;
  if (chunk > (size_t)ncols) chunk = ncols;
  if (chunk > 32768) chunk = 32768;
  DevBuf<u64> tmp(ctx, chunk * N);
  for (size_t c0 = 0; c0 < (size_t)ncols; c0 += chunk) {
    unsigned nc = (unsigned)std::min(chunk, (size_t)ncols - c0);
    dim3 g1((unsigned)(1u << (l2 - logT1)), nc), g2((unsigned)(1u << (l1 - logT2)), nc);
    { KScope ks(ctx, "ntt_pass1");
    k_ntt_pass1<<<g1, 512, smem1, ctx->stream>>>(in + c0 * in_stride, in_stride, tmp, W, prescale, l1, l2, logT1);
    LAUNCH_CHECK(ctx); }
    KScope ks2(ctx, "ntt_pass2");
    k_ntt_pass2<<<g2, 512, smem2, ctx->stream>>>(tmp, out + c0 * out_stride, out_stride, W, postscale, scale, l1, l2, logT2);
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
