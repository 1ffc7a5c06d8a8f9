// Poseidon-Goldilocks (width 12, rate 8, 4+22+4 rounds, x^7) -- the hash of plonky2's
// PoseidonGoldilocksConfig, which the reference selects at every prove() call site
// (e.g. reference src/curves/g1/exp.rs:788-790).  One permutation per thread, state in registers.
//
// Device path (the hot kernel of the whole prover: ~40 M permutations per G1 proof):
//  * lanes are kept as arbitrary u64 representatives (NOT canonical) between operations; only the
//    values leaving the permutation are canonicalised, so every result equals the canonical algorithm;
//  * S-box products from four (squares: three) explicit 32 x 32 -> 64 multiplications and an 11-instruction
//    2^64 = 2^32 - 1, 2^96 = -1 carry-chain reduction (gl.cuh);
//  * the MDS layer on 16-bit digits with the 2-way dot product IDP.2A (poseidon_mds_dp2a below).
// Variants measured and dropped (numbers in DESIGN.md K3, code in the history): 22/22/20-bit limbs on IMAD (50.9 ms per G1
// proof), byte digits on IDP.4A (48.4 ms), O(t) partial rounds with full 64-bit constants (58.8 ms), unrolled loop shapes.
// Host path (Fiat-Shamir challenger): branch-free arithmetic, partial rounds in the O(t) form (poseidon_fast.inc).
#pragma once
#include "gl.cuh"

static const u64 h_poseidon_rc[360] = {
#include "poseidon_rc.inc"
    SBN_POSEIDON_RC_LIST};
#ifdef __CUDACC__
static __constant__ u64 d_poseidon_rc[372] = {SBN_POSEIDON_RC_LIST, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // + a zero "round 30"
#endif

#ifdef __CUDACC__
static __constant__ __align__(16) u32 d_poseidon_rc_dig16[372 * 4] = {
#include "poseidon_rc_dig16.inc"
    SBN_POSEIDON_RC_DIG16_LIST};
#endif
#ifdef __CUDA_ARCH__
#define POSEIDON_RC(i) d_poseidon_rc[i]
#else
#define POSEIDON_RC(i) h_poseidon_rc[i]
#endif

#ifdef __CUDA_ARCH__
__device__ __forceinline__ u64 poseidon_sbox_nc(u64 x) {
  u64 x2 = gl_sqr_nc(x), x3 = gl_mul_nc(x2, x), x4 = gl_sqr_nc(x2);
  return gl_mul_nc(x3, x4);
}
// s <- MDS * s + rc[rc_off ..] (the next round's constants; offset 360 = zeros), lanes arbitrary u64 in and out.
// M[m] = the four circulant coefficients circ[m .. m+3] in one word: digit d of output lane r is
//   rc_digit + sum over lane pairs of dp2a(A[d][pair], M[(4 q - r) mod 12]).
#define SBN_MDS_WORD(m) ((u32)SBN_CIRC((m)) | ((u32)SBN_CIRC((m) + 1) << 8) | ((u32)SBN_CIRC((m) + 2) << 16) | ((u32)SBN_CIRC((m) + 3) << 24))
#define SBN_CIRC(i) ((i) % 12 == 0 ? 17u : (i) % 12 == 1 ? 15u : (i) % 12 == 2 ? 41u : (i) % 12 == 3 ? 16u : (i) % 12 == 4 ? 2u : (i) % 12 == 5 ? 28u : \
                     (i) % 12 == 6 ? 13u : (i) % 12 == 7 ? 13u : (i) % 12 == 8 ? 39u : (i) % 12 == 9 ? 18u : (i) % 12 == 10 ? 34u : 20u)
__device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 sel) { u32 r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r; }
// MDS layer on 16-bit digits with the 2-way dot product (IDP.2A: two 16-bit values times two bytes): each lane is cut into four
// digits, the digits of two lanes share a word (one PRMT), 12 lanes x 4 digits x 6 lane pairs = 288 dot products where the limb
// form needed 435 multiply-adds on the same pipe.  A digit sum is < 65535 * 264 + 65535 + 1024 < 2^25; the four sums c0..c3 are
// stitched at 16-bit spacing into 74 bits and folded with 2^64 = 2^32 - 1:
//   value = (c0 - h) + c1 2^16 + (c2 + h) 2^32 + (c3 mod 2^16) 2^48,  h = c3 >> 16 <= 264.
// The round constants enter as the initial accumulators, BIASED (tools/gen_poseidon_constants.py): digits of (rc - 1024) mod p
// with 1024 added back to digit 0, so c0 >= 1024 > h and the value above is a sum of non-negative terms that wraps 2^64 at most
// once; the wrap is folded by adding 2^32 - 1, which cannot wrap again (the high word is < 2^27 after a wrap).
// Instruction placement (ptxas lowers plain add / sub / shl / mov to IMAD.* on the multiplier pipe, which is this kernel's
// binding pipe): c0 - h is one signed dp2a against c3's own halves (replaces a PRMT + an IMAD.IADD), the 16-bit
// shifts are PRMTs, (c3 mod 2^16) 2^16 + h is one PRMT (rotation by 16), and every add of the stitch is part of a carry chain.
__device__ __forceinline__ u32 dp2a_lo(u32 a, u32 b, u32 c) { u32 r; asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 dp2a_hi(u32 a, u32 b, u32 c) { u32 r; asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
// the four digit sums of one output lane -> its 64-bit representative (comment above)
__device__ __forceinline__ u64 poseidon_stitch(u32 c0, u32 c1, u32 c2, u32 c3) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 rot, a, b, t1, t, e;\n\t"
      "prmt.b32 rot, %5, %5, 0x1032;\n\t"        // (c3 << 16) | (c3 >> 16) = (c3 mod 2^16) 2^16 + h
      "dp2a.lo.u32.s32 a, %5, 0xFF00, %2;\n\t"    // c0 + 0 * (c3 mod 2^16) - 1 * (c3 >> 16) = c0 - h
      "prmt.b32 t1, %3, 0, 0x1044;\n\t"          // c1 << 16
      "shf.r.clamp.b32 b, %3, 0, 16;\n\t"         // c1 >> 16
      "add.u32 b, b, %4;\n\t"
      "add.cc.u32 %0, a, t1;\n\t"
      "addc.cc.u32 %1, b, rot;\n\t"
      "addc.u32 t, 0x7FFFFFFF, 0;\n\t"            // bit 31 = carry (never mix add.cc with subc: ptxas keeps the carry as NOT borrow)
      "prmt.b32 e, t, 0, 0xBBBB;\n\t"             // sign of byte 3 replicated: 0xFFFFFFFF (= 2^32 - 1 as a 64-bit addend) after a wrap
      "add.cc.u32 %0, %0, e;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(c0), "r"(c1), "r"(c2), "r"(c3));
  return ((u64)r1 << 32) | r0;
}
__device__ __forceinline__ void poseidon_mds_dp2a(u64 s[12], const uint4* __restrict__ rc /* this layer's 12 biased digit quads */) {
  u32 A[4][6];   // A[d][j] = 16-bit digit d of lanes 2j (low half) and 2j+1 (high half)
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const u32 al = (u32)s[2 * j], ah = (u32)(s[2 * j] >> 32), bl = (u32)s[2 * j + 1], bh = (u32)(s[2 * j + 1] >> 32);
    A[0][j] = prmt(al, bl, 0x5410); A[1][j] = prmt(al, bl, 0x7632);
    A[2][j] = prmt(ah, bh, 0x5410); A[3][j] = prmt(ah, bh, 0x7632);
  }
#pragma unroll
  for (int r = 0; r < 12; r++) {
    const uint4 k = rc[r];
    u32 c[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const u32 m = SBN_MDS_WORD((4 * q + 12 - r) % 12) + ((r == 0 && q == 0) ? 8u : 0u);   // coefficients of lanes 4q .. 4q+3
#pragma unroll
      for (int d = 0; d < 4; d++) { c[d] = dp2a_lo(A[d][2 * q], m, c[d]); c[d] = dp2a_hi(A[d][2 * q + 1], m, c[d]); }
    }
    s[r] = poseidon_stitch(c[0], c[1], c[2], c[3]);
  }
}

#endif
#ifdef __CUDACC__
// ---- two lanes per permutation (small trees) ----
// A tree with few leaves cannot fill the machine with one permutation per thread: 2^14 leaves are 0.86 warps per scheduler, and a
// lone warp runs the permutation's dependent chains at ~0.5 instructions per clock (28 ms for the 1 225 absorptions of an Fq12
// trace row whatever the launch shape, tools/microbench).  Here lanes 2t and 2t+1 of a warp share a permutation: role r = lane & 1
// holds state elements 6r .. 6r+5 (local index i <-> global 6r + i), so a tree has twice the warps and each does half the work.
//  * S-boxes: six per lane in a full round; in a partial round both lanes run the S-box of local element 0 and role 1 keeps its
//    old value (same instruction stream, no divergence);
//  * MDS: the matrix is circulant, so in LOCAL indexing (own six elements first, then the partner's six) the rows of a lane's six
//    outputs are the first six rows of the same matrix for either role -- the unsplit code with the partner's digit words fetched
//    by 12 shuffles; only the diagonal term (8 s_0 into output 0) belongs to role 0 alone;
//  * round constants come from a shared-memory copy of the digit table (the index depends on the lane).
// State stays in the lazy representation across the permutations of a sponge; the caller canonicalises what leaves it.
struct PoseidonPairTables { uint4 dig[372]; u64 first[12]; };   // shared memory, filled by poseidon_pair_load_tables
#ifdef __CUDA_ARCH__
__device__ __forceinline__ void poseidon_pair_load_tables(PoseidonPairTables* t) {
  for (int i = threadIdx.x; i < 372; i += blockDim.x) t->dig[i] = reinterpret_cast<const uint4*>(d_poseidon_rc_dig16)[i];
  for (int i = threadIdx.x; i < 12; i += blockDim.x) t->first[i] = d_poseidon_rc[i];
  __syncthreads();
}
__device__ __forceinline__ void poseidon_mds_dp2a_pair(u64 s[6], const uint4* __restrict__ rc /* digit quads of this lane's six outputs */, u32 diag8) {
  u32 A[4][6];   // [d][0..2]: own lane pairs, [d][3..5]: the partner's
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const u32 al = (u32)s[2 * j], ah = (u32)(s[2 * j] >> 32), bl = (u32)s[2 * j + 1], bh = (u32)(s[2 * j + 1] >> 32);
    A[0][j] = prmt(al, bl, 0x5410); A[1][j] = prmt(al, bl, 0x7632);
    A[2][j] = prmt(ah, bh, 0x5410); A[3][j] = prmt(ah, bh, 0x7632);
  }
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int d = 0; d < 4; d++) A[d][3 + j] = __shfl_xor_sync(0xffffffffu, A[d][j], 1);
#pragma unroll
  for (int r = 0; r < 6; r++) {
    const uint4 k = rc[r];
    u32 c[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const u32 m = SBN_MDS_WORD((4 * q + 12 - r) % 12) + ((r == 0 && q == 0) ? diag8 : 0u);
#pragma unroll
      for (int d = 0; d < 4; d++) { c[d] = dp2a_lo(A[d][2 * q], m, c[d]); c[d] = dp2a_hi(A[d][2 * q + 1], m, c[d]); }
    }
    s[r] = poseidon_stitch(c[0], c[1], c[2], c[3]);
  }
}
// s = this lane's six state elements (arbitrary representatives in and out); every lane of the warp must call it.
__device__ __forceinline__ void poseidon_permute_pair(u64 s[6], int role, const PoseidonPairTables* t) {
#pragma unroll
  for (int i = 0; i < 6; i++) s[i] = gl_add_nc(s[i], t->first[6 * role + i]);
  const uint4* rc = t->dig + 12 + 6 * role;
  const u32 diag8 = role ? 0u : 8u;
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int k = 0; k < 4; k++, rc += 12) {
#pragma unroll
      for (int i = 0; i < 6; i++) s[i] = poseidon_sbox_nc(s[i]);
      poseidon_mds_dp2a_pair(s, rc, diag8);
    }
    if (half == 0) {
#pragma unroll 1
      for (int k = 0; k < 22; k++, rc += 12) {
        const u64 y = poseidon_sbox_nc(s[0]);
        s[0] = role ? s[0] : y;
        poseidon_mds_dp2a_pair(s, rc, diag8);
      }
    }
  }
}
#else
__device__ void poseidon_pair_load_tables(PoseidonPairTables* t);
__device__ void poseidon_permute_pair(u64 s[6], int role, const PoseidonPairTables* t);
#endif
#endif

HD u64 poseidon_sbox(u64 x) {
  u64 x2 = gl_sqr(x), x3 = gl_mul(x2, x), x4 = gl_sqr(x2);
  return gl_mul(x3, x4);
}

// out[k] = sum_i s[(i+k) % 12] * circ[i] + (k == 0 ? 8 * s[0] : 0), circ = [17,15,41,16,2,28,13,13,39,18,34,20]
HD void poseidon_mds(u64 s[12]) {
  const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  u32 lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { lo[i] = (u32)s[i]; hi[i] = (u32)(s[i] >> 32); }
#pragma unroll
  for (int k = 0; k < 12; k++) {
    u64 al = 0, ah = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      al += (u64)lo[(i + k) % 12] * C[i];
      ah += (u64)hi[(i + k) % 12] * C[i];
    }
    if (k == 0) { al += (u64)lo[0] * 8u; ah += (u64)hi[0] * 8u; }
    // value = al + ah * 2^32 with al, ah < 2^42
    u64 l = al + (ah << 32);
    u64 h = (ah >> 32) + (l < al ? 1 : 0);
    s[k] = gl_reduce128(l, h);
  }
}

#ifndef __CUDA_ARCH__
// Host path (Fiat-Shamir challenger: every opening is absorbed, ~10^4 field elements per proof).  Full rounds use
// 128-bit accumulators for the MDS layer; the 22 partial rounds run in the O(t)-per-round form whose constants
// tools/gen_poseidon_fast.py derives and checks against the plain permutation (poseidon_fast.inc).
static const u64 h_poseidon_fast[639] = {
#include "poseidon_fast.inc"
    SBN_POSEIDON_FAST_LIST};
// Branch-free arithmetic on arbitrary 64-bit representatives (carry / borrow of random operands is unpredictable, a
// mispredicted branch costs more than the whole reduction); only the permutation's outputs are canonicalised.
static inline u64 h_add(u64 a, u64 b) {
  u64 s, t; u64 c = __builtin_add_overflow(a, b, &s);
  u64 c2 = __builtin_add_overflow(s, (0 - c) & GL_EPS, &t);
  return t + ((0 - c2) & GL_EPS);
}
static inline u64 h_red(unsigned __int128 x) {   // 2^64 = 2^32 - 1, 2^96 = -1 (mod p)
  u64 lo = (u64)x, hi = (u64)(x >> 64), hh = hi >> 32, hl = hi & GL_EPS, t0, s;
  u64 bw = __builtin_sub_overflow(lo, hh, &t0); t0 -= (0 - bw) & GL_EPS;
  u64 c = __builtin_add_overflow(t0, (hl << 32) - hl, &s);
  return s + ((0 - c) & GL_EPS);
}
static inline u64 h_mul(u64 a, u64 b) { return h_red((unsigned __int128)a * b); }
static inline u64 h_sbox(u64 x) { u64 x2 = h_mul(x, x), x3 = h_mul(x2, x), x4 = h_mul(x2, x2); return h_mul(x3, x4); }
static inline void h_mds(u64 s[12]) {
  static const u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  u64 o[12];
  for (int k = 0; k < 12; k++) {
    unsigned __int128 acc = k == 0 ? (unsigned __int128)s[0] * 8 : 0;
    for (int i = 0; i < 12; i++) acc += (unsigned __int128)s[(i + k) % 12] * C[i];   // < 2^64 * 284
    o[k] = h_red(acc);
  }
  for (int k = 0; k < 12; k++) s[k] = o[k];
}
static inline void poseidon_permute_host(u64 s[12]) {
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 12; i++) s[i] = h_sbox(h_add(s[i], h_poseidon_rc[12 * r + i]));
    h_mds(s);
  }
  const u64* f = h_poseidon_fast;
  for (int r = 0; r < 22; r++, f += 23) {
    const u64 x0 = h_sbox(h_add(s[0], f[0]));
    // s0' = 25 x0 + v . s[1..];  s_i' = s_i + u_i x0     (sums of reduced 64-bit products: no 128-bit overflow)
    unsigned __int128 acc = (unsigned __int128)x0 * 25;
    for (int i = 0; i < 11; i++) acc += h_mul(f[1 + i], s[1 + i]);
    for (int i = 0; i < 11; i++) s[1 + i] = h_add(s[1 + i], h_mul(f[12 + i], x0));
    s[0] = h_red(acc);
  }
  {  // true state = diag(1, D^) x + e
    u64 o[11];
    for (int i = 0; i < 11; i++) {
      unsigned __int128 acc = f[121 + 1 + i];
      for (int j = 0; j < 11; j++) acc += h_mul(f[11 * i + j], s[1 + j]);
      o[i] = h_red(acc);
    }
    s[0] = h_add(s[0], f[121]);
    for (int i = 0; i < 11; i++) s[1 + i] = o[i];
  }
  for (int r = 26; r < 30; r++) {
    for (int i = 0; i < 12; i++) s[i] = h_sbox(h_add(s[i], h_poseidon_rc[12 * r + i]));
    h_mds(s);
  }
  for (int i = 0; i < 12; i++) s[i] = s[i] >= GL_P ? s[i] - GL_P : s[i];
}
#endif

// Canonical in, canonical out.
HD void poseidon_permute(u64 s[12]) {
#ifdef __CUDA_ARCH__
  // Two loop bodies -- a full round (12 S-boxes + MDS) and a partial round (1 S-box + MDS) -- so that each is scheduled for its
  // own instruction mix, both small enough to stay in the instruction cache (a fully unrolled permutation, > 90 KB, stalls on
  // instruction fetch; one rolled body with a uniform `full` branch was the round-1 shape).
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add_nc(s[i], d_poseidon_rc[i]);
  // The constant-bank index is kept in a per-thread (vector) register: inputs are canonical (< p), so `zero` is always 0, but not
  // provably so.  With a uniform index ptxas re-materialises it for every one of the 24 constant loads of a layer (24 IMAD.U32
  // per layer on the multiplier pipe); off a vector register they are one base + immediates.
  const u32 zero = s[11] == ~0ULL ? 1u : 0u;
  const uint4* rc = reinterpret_cast<const uint4*>(d_poseidon_rc_dig16) + 12 + zero;
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int k = 0; k < 4; k++, rc += 12) {
#pragma unroll
      for (int i = 0; i < 12; i++) s[i] = poseidon_sbox_nc(s[i]);
      poseidon_mds_dp2a(s, rc);
    }
    if (half == 0) {
#pragma unroll 1
      for (int k = 0; k < 22; k++, rc += 12) {
        s[0] = poseidon_sbox_nc(s[0]);
        poseidon_mds_dp2a(s, rc);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
#else
  poseidon_permute_host(s);
#endif
}
