// Micro-benchmark of the leaf-hash kernel variants (not product code): synthetic LDE, 2^17 leaves.
// -DPOSEIDON_R01: the frozen round-1 permutation (poseidon_r01_baseline.cuh) for before / after numbers on the same box.
#ifdef POSEIDON_R01
#include "poseidon_r01_baseline.cuh"
#define VARIANT "r01-baseline"
#else
#include "../../starky-bn254_b200/csrc/poseidon.cuh"
#define VARIANT "current"
#endif
#include <cstdio>
#include <vector>

template <int BS, int MINB> __global__ void __launch_bounds__(BS, MINB) k_leaf(const u64* __restrict__ lde, size_t L, int ncols, u64* __restrict__ digests) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= L) return;
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = 0;
  const u64* p = lde + idx;
#pragma unroll 1
  for (int c = 0; c < ncols; c += 8) {
    u64 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = p[(size_t)(c + i) * L];
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = v[i];
    poseidon_permute(st);
  }
  ulonglong2* d = reinterpret_cast<ulonglong2*>(digests + idx * 4);
  d[0] = make_ulonglong2(st[0], st[1]); d[1] = make_ulonglong2(st[2], st[3]);
}
#ifndef POSEIDON_R01
// two lanes per leaf (poseidon_permute_pair): role 0 absorbs columns c .. c+5, role 1 columns c+6, c+7
template <int BS> __global__ void __launch_bounds__(BS) k_leaf_pair(const u64* __restrict__ lde, size_t L, int ncols, u64* __restrict__ digests) {
  __shared__ PoseidonPairTables tab;
  poseidon_pair_load_tables(&tab);
  const size_t gt = blockIdx.x * (size_t)blockDim.x + threadIdx.x, idx = gt >> 1;
  const int role = (int)(gt & 1);
  if (idx >= L) return;   // L is a multiple of 16: whole warps leave together
  u64 st[6];
#pragma unroll
  for (int i = 0; i < 6; i++) st[i] = 0;
  const u64* p = lde + idx;
#pragma unroll 1
  for (int c = 0; c < ncols; c += 8) {
    const int c0 = c + 6 * role, n = role ? 2 : 6;
#pragma unroll
    for (int i = 0; i < 6; i++) if (i < n && c0 + i < ncols) st[i] = p[(size_t)(c0 + i) * L];
    poseidon_permute_pair(st, role, &tab);
  }
  if (role == 0) {
    ulonglong2* d = reinterpret_cast<ulonglong2*>(digests + idx * 4);
    d[0] = make_ulonglong2(gl_canon(st[0]), gl_canon(st[1])); d[1] = make_ulonglong2(gl_canon(st[2]), gl_canon(st[3]));
  }
}
template <int BS> float run_pair(const u64* lde, size_t L, int ncols, u64* dig, int iters) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k_leaf_pair<BS><<<(unsigned)((2 * L + BS - 1) / BS), BS>>>(lde, L, ncols, dig);
  cudaEventRecord(a);
  for (int i = 0; i < iters; i++) k_leaf_pair<BS><<<(unsigned)((2 * L + BS - 1) / BS), BS>>>(lde, L, ncols, dig);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}
#endif
__global__ void k_fill(u64* p, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 z = (i + 1) * 0x9E3779B97F4A7C15ULL; z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 32;
  p[i] = z >= GL_P ? z - GL_P : z;
}
template <int BS, int MINB> float run(const u64* lde, size_t L, int ncols, u64* dig, int iters) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k_leaf<BS, MINB><<<(unsigned)((L + BS - 1) / BS), BS>>>(lde, L, ncols, dig);
  cudaEventRecord(a);
  for (int i = 0; i < iters; i++) k_leaf<BS, MINB><<<(unsigned)((L + BS - 1) / BS), BS>>>(lde, L, ncols, dig);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}
int main(int argc, char** argv) {
  int logL = argc > 1 ? atoi(argv[1]) : 17, ncols = argc > 2 ? atoi(argv[2]) : 256;
  size_t L = size_t(1) << logL;
  u64 *lde, *dig;
  cudaMalloc(&lde, L * ncols * 8); cudaMalloc(&dig, L * 32);
  k_fill<<<(unsigned)((L * ncols + 255) / 256), 256>>>(lde, L * ncols);
  std::vector<u64> h(4);
  double perms = (double)L * (ncols / 8);
  float ms;
#define RUN(BS, MINB) ms = run<BS, MINB>(lde, L, ncols, dig, 3); cudaMemcpy(h.data(), dig + 4 * 777, 32, cudaMemcpyDeviceToHost); \
  printf("%s bs=%d minb=%d  %.3f ms  %.1f Mperm/s  dig=%016llx\n", VARIANT, BS, MINB, ms, perms / ms / 1e3, h[0]);
  RUN(128, 1) RUN(128, 6) RUN(128, 8) RUN(64, 14) RUN(256, 3) RUN(32, 1) RUN(64, 1) RUN(32, 16) RUN(64, 4)
#ifndef POSEIDON_R01
#define RUNP(BS) ms = run_pair<BS>(lde, L, ncols, dig, 3); cudaMemcpy(h.data(), dig + 4 * 777, 32, cudaMemcpyDeviceToHost); \
  printf("pair    bs=%d          %.3f ms  %.1f Mperm/s  dig=%016llx\n", BS, ms, perms / ms / 1e3, h[0]);
  RUNP(64) RUNP(128) RUNP(256)
#endif
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
