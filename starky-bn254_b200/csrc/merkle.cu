// K3: Poseidon-Goldilocks Merkle commitment over LDE rows.  Replaces plonky2's `MerkleTree::new`
// with `hash_or_noop` leaves and `two_to_one` nodes (external dependency), reached
// from the reference through `prove()` (reference src/curves/g1/exp.rs:818).
//
// One thread hashes one leaf (= one LDE row across all columns).  The LDE batch is column-major, so
// the 32 threads of a warp read 32 consecutive u64 of the same column: every load is a fully
// coalesced 256-byte run and the row never has to be materialised.  Digests are written to the
// bit-reversed leaf position plonky2 uses (32-byte aligned stores).
#include "merkle.cuh"
#include "poseidon.cuh"

// sub_coset < 0: the whole batch lde[col][b][k] (col_stride = L); sub_coset = b: one sub-coset sub[col][k] (col_stride = N) of a
// streamed commitment -- the digests land at the same positions either way.
__global__ void __launch_bounds__(128) k_leaf_hash(const u64* __restrict__ lde, size_t col_stride, int ncols, int logn, int rate_bits,
                                                   u64* __restrict__ digests, int sub_coset) {
  const size_t L = size_t(1) << (logn + (sub_coset < 0 ? rate_bits : 0));
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= L) return;
  const u32 b = sub_coset < 0 ? (u32)(idx >> logn) : (u32)sub_coset, k = (u32)(idx & ((size_t(1) << logn) - 1));
  const size_t pos = ((size_t)bitrev32(b, rate_bits) << logn) + bitrev32(k, logn);
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = 0;
  const u64* p = lde + idx;
  if (ncols <= 4) {  // hash_or_noop: short rows are copied, not hashed
#pragma unroll
    for (int c = 0; c < 4; c++) if (c < ncols) st[c] = p[(size_t)c * col_stride];
  } else {
    // single loop / single inlined permutation (instruction-cache footprint); the ragged tail keeps the
    // old state in the lanes it does not overwrite (overwrite-mode sponge, no padding)
#pragma unroll 1
    for (int c = 0; c < ncols; c += 8) {
      u64 v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = (c + i < ncols) ? p[(size_t)(c + i) * col_stride] : st[i];
#pragma unroll
      for (int i = 0; i < 8; i++) st[i] = v[i];
      poseidon_permute(st);
    }
  }
  ulonglong2* d = reinterpret_cast<ulonglong2*>(digests + pos * 4);
  d[0] = make_ulonglong2(st[0], st[1]);
  d[1] = make_ulonglong2(st[2], st[3]);
}

// The same leaves with two lanes per permutation (poseidon_permute_pair): role 0 absorbs columns c .. c+5, role 1 columns c+6, c+7.
// For trees of at most 2^16 leaves, where one thread per leaf leaves schedulers empty or with few warps (tools/microbench:
// 9 800 columns, 2^13 leaves 28.0 -> 16.5 ms, 2^14 leaves 28.4 -> 25.6 ms; 2^15 x 4 096 columns 19.7 -> 19.3 ms; 2^16 x 2 816 columns
// 25.4 -> 23.1 ms; at 2^17 leaves the one-thread form is as fast or faster and stays).
template <int BS> __global__ void __launch_bounds__(BS) k_leaf_hash_pair(const u64* __restrict__ lde, size_t col_stride, int ncols, int logn, int rate_bits,
                                                                         u64* __restrict__ digests, int sub_coset) {
  __shared__ PoseidonPairTables tab;
  poseidon_pair_load_tables(&tab);
  const size_t L = size_t(1) << (logn + (sub_coset < 0 ? rate_bits : 0));
  const size_t gt = blockIdx.x * (size_t)blockDim.x + threadIdx.x, idx = gt >> 1;
  const int role = (int)(gt & 1);
  if (idx >= L) return;   // L is a multiple of 16 (the caller checks): whole warps leave together
  const u32 b = sub_coset < 0 ? (u32)(idx >> logn) : (u32)sub_coset, k = (u32)(idx & ((size_t(1) << logn) - 1));
  const size_t pos = ((size_t)bitrev32(b, rate_bits) << logn) + bitrev32(k, logn);
  u64 st[6];
#pragma unroll
  for (int i = 0; i < 6; i++) st[i] = 0;
  const u64* p = lde + idx;
#pragma unroll 1
  for (int c = 0; c < ncols; c += 8) {   // overwrite-mode sponge; the ragged tail keeps the old state in the lanes it does not overwrite
    const int c0 = c + 6 * role, n = role ? 2 : 6;
#pragma unroll
    for (int i = 0; i < 6; i++) if (i < n && c0 + i < ncols) st[i] = p[(size_t)(c0 + i) * col_stride];
    poseidon_permute_pair(st, role, &tab);
  }
  if (role == 0) {
    ulonglong2* d = reinterpret_cast<ulonglong2*>(digests + pos * 4);
    d[0] = make_ulonglong2(gl_canon(st[0]), gl_canon(st[1]));
    d[1] = make_ulonglong2(gl_canon(st[2]), gl_canon(st[3]));
  }
}
static void launch_leaf_hash(sbn_ctx* ctx, const u64* lde, size_t col_stride, size_t nleaves, int ncols, int logn, int rate_bits, u64* digests, int sub_coset) {
  const bool no_pair = getenv("SBN_LEAF_HASH_ONE_THREAD") != nullptr;   // test switch: always one thread per leaf
  if (ncols > 4 && nleaves >= 16 && nleaves <= (size_t(1) << 16) && !no_pair) {
    // block size by measurement (profiles/r02_poseidon_pair_microbench.txt): 256 threads for 2^14 .. 2^15 leaves, 128 otherwise
    if (nleaves > (size_t(1) << 13) && nleaves <= (size_t(1) << 15)) k_leaf_hash_pair<256><<<(unsigned)((2 * nleaves + 255) / 256), 256, 0, ctx->stream>>>(lde, col_stride, ncols, logn, rate_bits, digests, sub_coset);
    else k_leaf_hash_pair<128><<<(unsigned)((2 * nleaves + 127) / 128), 128, 0, ctx->stream>>>(lde, col_stride, ncols, logn, rate_bits, digests, sub_coset);
  } else {
    k_leaf_hash<<<(unsigned)((nleaves + 127) / 128), 128, 0, ctx->stream>>>(lde, col_stride, ncols, logn, rate_bits, digests, sub_coset);
  }
}

__global__ void __launch_bounds__(128) k_merkle_level(const u64* __restrict__ child, u64* __restrict__ parent, size_t nparents) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nparents) return;
  const ulonglong2* c = reinterpret_cast<const ulonglong2*>(child + i * 8);
  ulonglong2 a0 = c[0], a1 = c[1], b0 = c[2], b1 = c[3];
  u64 st[12] = {a0.x, a0.y, a1.x, a1.y, b0.x, b0.y, b1.x, b1.y, 0, 0, 0, 0};
  poseidon_permute(st);
  ulonglong2* d = reinterpret_cast<ulonglong2*>(parent + i * 4);
  d[0] = make_ulonglong2(st[0], st[1]);
  d[1] = make_ulonglong2(st[2], st[3]);
}

// two_to_one with two lanes per parent for the small upper levels (one permutation of latency per level either way: 23 us
// with one thread per parent, about half with two lanes): role 0 holds the left child and half of the right, role 1 the rest.
__global__ void __launch_bounds__(128) k_merkle_level_pair(const u64* __restrict__ child, u64* __restrict__ parent, size_t nparents) {
  __shared__ PoseidonPairTables tab;
  poseidon_pair_load_tables(&tab);
  const size_t gt = blockIdx.x * (size_t)blockDim.x + threadIdx.x, i = gt >> 1;
  const int role = (int)(gt & 1);
  if (i >= nparents) return;   // nparents is a multiple of 16: whole warps leave together
  const ulonglong2* c = reinterpret_cast<const ulonglong2*>(child + i * 8);
  u64 st[6] = {0, 0, 0, 0, 0, 0};
  if (role == 0) { const ulonglong2 a0 = c[0], a1 = c[1], b0 = c[2]; st[0] = a0.x; st[1] = a0.y; st[2] = a1.x; st[3] = a1.y; st[4] = b0.x; st[5] = b0.y; }
  else { const ulonglong2 b1 = c[3]; st[0] = b1.x; st[1] = b1.y; }
  poseidon_permute_pair(st, role, &tab);
  if (role == 0) {
    ulonglong2* d = reinterpret_cast<ulonglong2*>(parent + i * 4);
    d[0] = make_ulonglong2(gl_canon(st[0]), gl_canon(st[1]));
    d[1] = make_ulonglong2(gl_canon(st[2]), gl_canon(st[3]));
  }
}

void merkle_alloc(sbn_ctx* ctx, DevMerkleTree* t, size_t nleaves, int cap_height) {
  SBN_REQUIRE(nleaves >= (size_t(1) << cap_height), "merkle: fewer leaves than cap entries");
  t->nleaves = nleaves; t->cap_height = cap_height;
  t->level_off.clear();
  size_t off = 0;
  for (size_t n = nleaves; n >= (size_t(1) << cap_height); n >>= 1) { t->level_off.push_back(off); off += n * 4; if (n == 1) break; }
  t->digests = DevBuf<u64>(ctx, off);
}

void merkle_build_from_leaf_digests(sbn_ctx* ctx, DevMerkleTree* t) {
  size_t n = t->nleaves;
  KScope ks(ctx, "merkle_tree_levels");
  for (int l = 0; l + 1 < t->num_levels(); l++) {
    size_t np = n >> 1;
    if (np >= 16 && np <= (size_t(1) << 13) && !getenv("SBN_LEAF_HASH_ONE_THREAD"))
      k_merkle_level_pair<<<(unsigned)((2 * np + 127) / 128), 128, 0, ctx->stream>>>(t->digests + t->level_off[l], t->digests + t->level_off[l + 1], np);
    else
      k_merkle_level<<<(unsigned)((np + 127) / 128), 128, 0, ctx->stream>>>(t->digests + t->level_off[l], t->digests + t->level_off[l + 1], np);
    LAUNCH_CHECK(ctx);
    n = np;
  }
  t->cap.resize((size_t(4)) << t->cap_height);
  CUDA_CHECK(cudaMemcpyAsync(t->cap.data(), t->digests + t->level_off.back(), t->cap.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
}

void merkle_leaf_hash_only(sbn_ctx* ctx, const u64* lde, int ncols, int logn, int rate_bits, DevMerkleTree* t) {
  size_t L = size_t(1) << (logn + rate_bits);
  KScope ks(ctx, "merkle_leaf_hash");
  launch_leaf_hash(ctx, lde, L, L, ncols, logn, rate_bits, t->digests, -1);
  LAUNCH_CHECK(ctx);
}
void merkle_leaf_hash_sub_coset(sbn_ctx* ctx, const u64* sub, int ncols, int logn, int rate_bits, int b, DevMerkleTree* t) {
  size_t N = size_t(1) << logn;
  KScope ks(ctx, "merkle_leaf_hash");
  launch_leaf_hash(ctx, sub, N, N, ncols, logn, rate_bits, t->digests, b);
  LAUNCH_CHECK(ctx);
}

void merkle_commit_lde(sbn_ctx* ctx, const u64* lde, int ncols, int logn, int rate_bits, int cap_height, DevMerkleTree* out) {
  size_t L = size_t(1) << (logn + rate_bits);
  merkle_alloc(ctx, out, L, cap_height);
  merkle_leaf_hash_only(ctx, lde, ncols, logn, rate_bits, out);
  merkle_build_from_leaf_digests(ctx, out);
}
