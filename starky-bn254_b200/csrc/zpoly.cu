// K4: permutation Z polynomials.  Replaces starky's `compute_permutation_z_polys`
// (external dependency) for the singleton column pairs the reference's
// `permutation_pairs()` produce (reference src/utils/range_check.rs:96-113, :230-246).
//
// Z[r] = prod_{i<r} num_i / den_i  is evaluated without per-element inversions as
//        Z[r] = (prod_{i<r} num_i) * (prod_{i>=r} den_i) / (prod_all den_i):
// an exclusive prefix scan of num, an inclusive suffix scan of den, and one field inversion per
// column.  Rows are processed in 256-row tiles (coalesced column reads); tile products are combined by
// a tiny per-column pass.  Exact field arithmetic, so the values equal the sequential definition.
#include "quotient.cuh"
#include "zpoly.cuh"

#define ZT 256

__device__ __forceinline__ void row_num_den(const u64* __restrict__ trace, size_t stride, size_t r, const u32* lhs, const u32* rhs,
                                            const u64* gamma, int batch, int z, u64& num, u64& den) {
  num = 1; den = 1;
  for (int j = 0; j < batch; j++) {
    int e = z * batch + j;
    u32 l = lhs[e];
    if (l == 0xFFFFFFFFu) break;
    u64 g = gamma[e];
    num = gl_mul_nc(num, gl_add_nc(trace[(size_t)l * stride + r], g));
    den = gl_mul_nc(den, gl_add_nc(trace[(size_t)rhs[e] * stride + r], g));
  }
}
// block-wide inclusive multiplicative scans: prefix over v (result in pre) and suffix over w (result in suf).
// Warp scans by shuffle; the 8 warp totals are scanned once by warp 0 and shared, so a thread pays 10 multiplications for
// the warp level and 2 for the block level.
__device__ __forceinline__ void block_scan_mul(u64 v, u64 w, u64& pre, u64& suf) {
  __shared__ u64 sv[ZT / 32], sw[ZT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 a = v, b = w;
  for (int d = 1; d < 32; d <<= 1) {
    u64 ta = __shfl_up_sync(0xffffffffu, a, d), tb = __shfl_down_sync(0xffffffffu, b, d);
    if (lane >= d) a = gl_mul_nc(a, ta);
    if (lane + d < 32) b = gl_mul_nc(b, tb);
  }
  if (lane == 31) sv[warp] = a;
  if (lane == 0) sw[warp] = b;
  __syncthreads();
  if (warp == 0) {   // exclusive prefix of the warp totals of v, exclusive suffix of those of w
    u64 x = lane < ZT / 32 ? sv[lane] : 1, y = lane < ZT / 32 ? sw[lane] : 1;
    u64 px = x, sy = y;
    for (int d = 1; d < ZT / 32; d <<= 1) {
      u64 tx = __shfl_up_sync(0xffffffffu, px, d), ty = __shfl_down_sync(0xffffffffu, sy, d);
      if (lane >= d) px = gl_mul_nc(px, tx);
      if (lane + d < ZT / 32) sy = gl_mul_nc(sy, ty);
    }
    u64 ex = __shfl_up_sync(0xffffffffu, px, 1), ey = __shfl_down_sync(0xffffffffu, sy, 1);
    if (lane < ZT / 32) { sv[lane] = lane == 0 ? 1 : ex; sw[lane] = lane == ZT / 32 - 1 ? 1 : ey; }
  }
  __syncthreads();
  pre = gl_mul_nc(a, sv[warp]); suf = gl_mul_nc(b, sw[warp]);
  __syncthreads();
}
// block-wide products of v and of w (valid in thread 0)
__device__ __forceinline__ void block_reduce_mul(u64 v, u64 w, u64& tv, u64& tw) {
  __shared__ u64 rv[ZT / 32], rw[ZT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = 16; d >= 1; d >>= 1) { v = gl_mul_nc(v, __shfl_xor_sync(0xffffffffu, v, d)); w = gl_mul_nc(w, __shfl_xor_sync(0xffffffffu, w, d)); }
  if (lane == 0) { rv[warp] = v; rw[warp] = w; }
  __syncthreads();
  tv = 1; tw = 1;
  if (threadIdx.x == 0) for (int i = 0; i < ZT / 32; i++) { tv = gl_mul_nc(tv, rv[i]); tw = gl_mul_nc(tw, rw[i]); }
}

__global__ void __launch_bounds__(ZT) k_z_tile_products(const u64* __restrict__ trace, size_t stride, const u32* lhs, const u32* rhs, const u64* gamma,
                                                        int batch, u64* tile_num, u64* tile_den, int ntiles) {
  const int z = blockIdx.y, tile = blockIdx.x;
  size_t r = (size_t)tile * ZT + threadIdx.x;
  u64 num, den;
  row_num_den(trace, stride, r, lhs, rhs, gamma, batch, z, num, den);
  u64 tn, td;
  block_reduce_mul(num, den, tn, td);
  if (threadIdx.x == 0) { tile_num[(size_t)z * ntiles + tile] = gl_canon(tn); tile_den[(size_t)z * ntiles + tile] = gl_canon(td); }
}
// per column: tile_num -> exclusive prefix, tile_den -> exclusive suffix times 1/prod(all den)
__global__ void k_z_tile_scan(u64* tile_num, u64* tile_den, int ntiles, int nz) {
  int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= nz) return;
  u64* tn = tile_num + (size_t)z * ntiles; u64* td = tile_den + (size_t)z * ntiles;
  u64 acc = 1;
  for (int t = 0; t < ntiles; t++) { u64 v = tn[t]; tn[t] = acc; acc = gl_mul(acc, v); }
  u64 tot = 1;
  for (int t = 0; t < ntiles; t++) tot = gl_mul(tot, td[t]);
  acc = gl_inv(tot);
  for (int t = ntiles - 1; t >= 0; t--) { u64 v = td[t]; td[t] = acc; acc = gl_mul(acc, v); }
}
__global__ void __launch_bounds__(ZT) k_z_finish(const u64* __restrict__ trace, size_t stride, const u32* lhs, const u32* rhs, const u64* gamma, int batch,
                                                 const u64* tile_num, const u64* tile_den, int ntiles, u64* __restrict__ zout, size_t N) {
  const int z = blockIdx.y, tile = blockIdx.x;
  size_t r = (size_t)tile * ZT + threadIdx.x;
  u64 num, den;
  row_num_den(trace, stride, r, lhs, rhs, gamma, batch, z, num, den);
  u64 pre, suf;
  block_scan_mul(num, den, pre, suf);
  // exclusive prefix of num within the tile = inclusive prefix of the previous thread
  u64 pre_ex = __shfl_up_sync(0xffffffffu, pre, 1);
  __shared__ u64 warp_last[ZT / 32];
  if ((threadIdx.x & 31) == 31) warp_last[threadIdx.x >> 5] = pre;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) pre_ex = threadIdx.x == 0 ? 1 : warp_last[(threadIdx.x >> 5) - 1];
  u64 v = gl_mul_nc(gl_mul_nc(pre_ex, tile_num[(size_t)z * ntiles + tile]), gl_mul_nc(suf, tile_den[(size_t)z * ntiles + tile]));
  zout[(size_t)z * N + r] = gl_canon(v);   // products ran on lazy representatives (gl.cuh)
}

void compute_z_polys(sbn_ctx* ctx, const u64* trace, int logn, const PermInstances& perm, u64* z_out) {
  size_t N = size_t(1) << logn;
  int nz = (int)perm.nz();
  if (!nz) return;
  SBN_REQUIRE(N % ZT == 0, "trace too short for the Z-polynomial kernel (need >= 256 rows)");
  int ntiles = (int)(N / ZT);
  size_t ne = perm.lhs.size();
  DevBuf<u32> d_lhs(ctx, ne), d_rhs(ctx, ne); DevBuf<u64> d_gamma(ctx, ne);
  ctx->upload(d_lhs, perm.lhs.data(), ne * 4);
  ctx->upload(d_rhs, perm.rhs.data(), ne * 4);
  ctx->upload(d_gamma, perm.gamma.data(), ne * 8);
  DevBuf<u64> tn(ctx, (size_t)nz * ntiles), td(ctx, (size_t)nz * ntiles);
  dim3 grid(ntiles, nz);
  KScope ks(ctx, "zpoly");
  k_z_tile_products<<<grid, ZT, 0, ctx->stream>>>(trace, N, d_lhs, d_rhs, d_gamma, perm.batch_size, tn, td, ntiles);
  LAUNCH_CHECK(ctx);
  k_z_tile_scan<<<(nz + 63) / 64, 64, 0, ctx->stream>>>(tn, td, ntiles, nz);
  LAUNCH_CHECK(ctx);
  k_z_finish<<<grid, ZT, 0, ctx->stream>>>(trace, N, d_lhs, d_rhs, d_gamma, perm.batch_size, tn, td, ntiles, z_out, N);
  LAUNCH_CHECK(ctx);
}
