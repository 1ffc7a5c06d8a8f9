// Poseidon-Goldilocks (width 12, rate 8, 4+22+4 rounds, x^7) -- the hash of plonky2's
// PoseidonGoldilocksConfig, which the reference selects at every prove() call site
// (e.g. reference src/curves/g1/exp.rs:788-790).  One permutation per thread, state in registers.
//
// Device path (the hot kernel of the whole prover: ~40 M permutations per G1 proof):
//  * lanes are kept as arbitrary u64 representatives (NOT canonical) between operations; only the
//    values leaving the permutation are canonicalised, so every result equals the canonical algorithm;
//  * 64x64 multiply and the 2^64 = 2^32 - 1, 2^96 = -1 reduction are PTX carry chains (23 instructions);
//  * the MDS layer works on the 32-bit halves of each lane (every product is a 32 x 9-bit IMAD), the next
//    round's constants are added to the un-reduced accumulators, and one 10-instruction reduction per lane
//    brings the < 2^75 sums back to 64 bits.
// Host path (Fiat-Shamir challenger, a few dozen permutations per proof): plain canonical arithmetic.
#pragma once
#include "gl.cuh"

static const u64 h_poseidon_rc[360] = {
#include "poseidon_rc.inc"
    SBN_POSEIDON_RC_LIST};
#ifdef __CUDACC__
static __constant__ u64 d_poseidon_rc[372] = {SBN_POSEIDON_RC_LIST, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // + a zero "round 30"
#endif

#ifdef __CUDACC__
static __constant__ u32 d_poseidon_rc_l22[372 * 3] = {
#include "poseidon_rc_l22.inc"
    SBN_POSEIDON_RC_L22_LIST};
// MDS coefficients as run-time constant-bank operands: with literal constants ptxas strength-reduces the
// products into 64-bit shift/add chains, which more than doubles the instruction count of the MDS layer.
static __constant__ u32 d_mds_c[13] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20, 8};
#endif
#ifdef __CUDACC__
static __constant__ __align__(16) u32 d_poseidon_rc_dig[372 * 8] = {
#include "poseidon_rc_dig.inc"
    SBN_POSEIDON_RC_DIG_LIST};
#endif
#ifdef __CUDACC__
static __constant__ __align__(16) u32 d_poseidon_rc_dig16[372 * 4] = {
#include "poseidon_rc_dig16.inc"
    SBN_POSEIDON_RC_DIG16_LIST};
#endif
#ifdef __CUDACC__
// O(t)-per-round form of the 22 partial rounds (tools/gen_poseidon_fast.py; the table the host challenger uses):
// per round g0, v[11], u[11]; then D^[11][11]; then e[12].
static __constant__ u64 d_poseidon_fast[639] = {
#include "poseidon_fast.inc"
    SBN_POSEIDON_FAST_LIST};
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
// Each lane is cut into limbs of 22 + 22 + 20 bits, so every product with an MDS coefficient (<= 41, row sum 284)
// and the whole 12-term sum plus the constant's limb stay below 2^32: the layer is 3 * 145 native 32-bit
// multiply-adds (IMAD issues at twice the rate of IMAD.WIDE on sm_100, tools/microbench/int_throughput.cu)
// followed by one carry-chain recombination per lane:
//   V = a0 + a1 2^22 + a2 2^44 < 2^74  ->  (w1:w0) + h 2^64,  h < 2^10,  then  + h (2^32 - 1)  with one wrap fix.
__device__ __forceinline__ void poseidon_mds_nc(u64 s[12], int rc_off) {
  u32 l0[12], l1[12], l2[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    l0[i] = (u32)s[i] & 0x3FFFFFu;
    l1[i] = (u32)(s[i] >> 22) & 0x3FFFFFu;
    l2[i] = (u32)(s[i] >> 44);
  }
#pragma unroll
  for (int k = 0; k < 12; k++) {
    u32 a0 = d_poseidon_rc_l22[3 * (rc_off + k)], a1 = d_poseidon_rc_l22[3 * (rc_off + k) + 1], a2 = d_poseidon_rc_l22[3 * (rc_off + k) + 2];
#pragma unroll
    for (int i = 0; i < 12; i++) {
      asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a0) : "r"(l0[(i + k) % 12]), "r"(d_mds_c[i]));
      asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a1) : "r"(l1[(i + k) % 12]), "r"(d_mds_c[i]));
      asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a2) : "r"(l2[(i + k) % 12]), "r"(d_mds_c[i]));
    }
    if (k == 0) {
      asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a0) : "r"(l0[0]), "r"(d_mds_c[12]));
      asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a1) : "r"(l1[0]), "r"(d_mds_c[12]));
      asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a2) : "r"(l2[0]), "r"(d_mds_c[12]));
    }
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 p, q, y, h, el, eh, c;\n\t"
        "shl.b32 p, %3, 22;\n\t"
        "shr.u32 q, %3, 10;\n\t"
        "shl.b32 y, %4, 12;\n\t"
        "shr.u32 h, %4, 20;\n\t"
        "add.cc.u32 %0, %2, p;\n\t"
        "addc.cc.u32 %1, q, y;\n\t"
        "addc.u32 h, h, 0;\n\t"             // overflow above 2^64, < 2^10
        "mul.lo.u32 el, h, 0xFFFFFFFF;\n\t"  // h * (2^32 - 1) = eh:el
        "mul.hi.u32 eh, h, 0xFFFFFFFF;\n\t"
        "add.cc.u32 %0, %0, el;\n\t"
        "addc.cc.u32 %1, %1, eh;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "neg.s32 c, c;\n\t"                 // wrapped once: + (2^32 - 1); the sum was < 2^64 + 2^42, so no second wrap
        "add.cc.u32 %0, %0, c;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(a0), "r"(a1), "r"(a2));
    s[k] = ((u64)r1 << 32) | r0;
  }
}
// MDS layer on byte digits with the 4-way byte dot product (IDP.4A): lane values are cut into 8 digits of 8 bits, the
// digits of four lanes are packed into one word (a 4x4 byte transpose, 8 PRMT per 4 lanes and half), and digit d of
// output lane r is  rc_digit + sum_q dp4a(A[d][q], M[(4q - r) mod 12])  with M[m] = the four circulant coefficients
// circ[m .. m+3] in one word -- 3 dot-product instructions where the limb form needs 12 multiply-adds on the same pipe.
// A digit sum is < 255 * 292 + 255 < 2^17; the 8 sums are stitched at 8-bit spacing into 74 bits and folded with
// 2^64 = 2^32 - 1 on the ALU pipe.  Lanes arbitrary u64 in and out.
#define SBN_MDS_WORD(m) ((u32)SBN_CIRC((m)) | ((u32)SBN_CIRC((m) + 1) << 8) | ((u32)SBN_CIRC((m) + 2) << 16) | ((u32)SBN_CIRC((m) + 3) << 24))
#define SBN_CIRC(i) ((i) % 12 == 0 ? 17u : (i) % 12 == 1 ? 15u : (i) % 12 == 2 ? 41u : (i) % 12 == 3 ? 16u : (i) % 12 == 4 ? 2u : (i) % 12 == 5 ? 28u : \
                     (i) % 12 == 6 ? 13u : (i) % 12 == 7 ? 13u : (i) % 12 == 8 ? 39u : (i) % 12 == 9 ? 18u : (i) % 12 == 10 ? 34u : 20u)
__device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 sel) { u32 r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r; }
__device__ __forceinline__ u32 dp4a_u(u32 a, u32 b, u32 c) { u32 r; asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ void poseidon_mds_dp4a(u64 s[12], int rc_off) {
  u32 A[8][3];
#pragma unroll
  for (int q = 0; q < 3; q++) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const u32 w0 = (u32)(s[4 * q] >> (32 * h)), w1 = (u32)(s[4 * q + 1] >> (32 * h)), w2 = (u32)(s[4 * q + 2] >> (32 * h)), w3 = (u32)(s[4 * q + 3] >> (32 * h));
      const u32 u0 = prmt(w0, w1, 0x5140), u1 = prmt(w0, w1, 0x7362), v0 = prmt(w2, w3, 0x5140), v1 = prmt(w2, w3, 0x7362);
      A[4 * h + 0][q] = prmt(u0, v0, 0x5410); A[4 * h + 1][q] = prmt(u0, v0, 0x7632);
      A[4 * h + 2][q] = prmt(u1, v1, 0x5410); A[4 * h + 3][q] = prmt(u1, v1, 0x7632);
    }
  }
#pragma unroll
  for (int r = 0; r < 12; r++) {
    u32 c[8];
    const uint4 k0 = reinterpret_cast<const uint4*>(d_poseidon_rc_dig)[2 * (rc_off + r)], k1 = reinterpret_cast<const uint4*>(d_poseidon_rc_dig)[2 * (rc_off + r) + 1];
    const u32 kd[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
    for (int d = 0; d < 8; d++) {
      u32 acc = kd[d];
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const u32 m = SBN_MDS_WORD((4 * q + 12 - r) % 12) + ((r == 0 && q == 0) ? 8u : 0u);   // + diag[0] = 8
        acc = dp4a_u(A[d][q], m, acc);
      }
      c[d] = acc;
    }
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 p0, p1, p2, p3, t1, t2, t3, t4, h, b, d, e, f;\n\t"
        "shl.b32 t1, %3, 8;\n\t"  "add.u32 p0, %2, t1;\n\t"      // digit pairs, < 2^26
        "shl.b32 t1, %5, 8;\n\t"  "add.u32 p1, %4, t1;\n\t"
        "shl.b32 t1, %7, 8;\n\t"  "add.u32 p2, %6, t1;\n\t"
        "shl.b32 t1, %9, 8;\n\t"  "add.u32 p3, %8, t1;\n\t"
        "shl.b32 t1, p1, 16;\n\t" "shr.u32 t2, p1, 16;\n\t" "add.u32 t3, t2, p2;\n\t"
        "shl.b32 t4, p3, 16;\n\t" "shr.u32 h, p3, 16;\n\t"
        "add.cc.u32 %0, p0, t1;\n\t"
        "addc.cc.u32 %1, t3, t4;\n\t"
        "addc.u32 h, h, 0;\n\t"              // value = h 2^64 + (%1:%0), h < 2^11
        "sub.cc.u32 %0, %0, h;\n\t"          // + h (2^32 - 1)
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.u32 b, 0, 0;\n\t"
        "add.cc.u32 %1, %1, h;\n\t"
        "addc.u32 d, b, 0;\n\t"              // net 64-bit wraps, in {-1, 0, 1}
        "neg.s32 e, d;\n\t"
        "shr.s32 f, d, 31;\n\t"
        "add.cc.u32 %0, %0, e;\n\t"
        "addc.u32 %1, %1, f;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]));
    s[r] = ((u64)r1 << 32) | r0;
  }
}
// The same layer on 16-bit digits with the 2-way dot product (IDP.2A: two 16-bit values times two bytes): 4 digit sums per
// output lane instead of 8 (each < 65535 * 292 + 65535 < 2^25), so half the packing and half the stitching; the number of
// dot-product instructions is the same (12 lanes x 4 digits x 6 lane pairs).
__device__ __forceinline__ u32 dp2a_lo(u32 a, u32 b, u32 c) { u32 r; asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 dp2a_hi(u32 a, u32 b, u32 c) { u32 r; asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ void poseidon_mds_dp2a(u64 s[12], int rc_off) {
  u32 A[4][6];   // A[d][j] = 16-bit digit d of lanes 2j (low half) and 2j+1 (high half)
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const u32 al = (u32)s[2 * j], ah = (u32)(s[2 * j] >> 32), bl = (u32)s[2 * j + 1], bh = (u32)(s[2 * j + 1] >> 32);
    A[0][j] = prmt(al, bl, 0x5410); A[1][j] = prmt(al, bl, 0x7632);
    A[2][j] = prmt(ah, bh, 0x5410); A[3][j] = prmt(ah, bh, 0x7632);
  }
#pragma unroll
  for (int r = 0; r < 12; r++) {
    const uint4 k = reinterpret_cast<const uint4*>(d_poseidon_rc_dig16)[rc_off + r];
    u32 c[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const u32 m = SBN_MDS_WORD((4 * q + 12 - r) % 12) + ((r == 0 && q == 0) ? 8u : 0u);   // coefficients of lanes 4q .. 4q+3
#pragma unroll
      for (int d = 0; d < 4; d++) { c[d] = dp2a_lo(A[d][2 * q], m, c[d]); c[d] = dp2a_hi(A[d][2 * q + 1], m, c[d]); }
    }
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 t1, t2, t3, t4, h, b, d, e, f;\n\t"
        "shl.b32 t1, %3, 16;\n\t" "shr.u32 t2, %3, 16;\n\t" "add.u32 t3, t2, %4;\n\t"
        "shl.b32 t4, %5, 16;\n\t" "shr.u32 h, %5, 16;\n\t"
        "add.cc.u32 %0, %2, t1;\n\t"
        "addc.cc.u32 %1, t3, t4;\n\t"
        "addc.u32 h, h, 0;\n\t"              // value = h 2^64 + (%1:%0), h < 2^10
        "sub.cc.u32 %0, %0, h;\n\t"          // + h (2^32 - 1)
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.u32 b, 0, 0;\n\t"
        "add.cc.u32 %1, %1, h;\n\t"
        "addc.u32 d, b, 0;\n\t"              // net 64-bit wraps, in {-1, 0, 1}
        "neg.s32 e, d;\n\t"
        "shr.s32 f, d, 31;\n\t"
        "add.cc.u32 %0, %0, e;\n\t"
        "addc.u32 %1, %1, f;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
    s[r] = ((u64)r1 << 32) | r0;
  }
}
// The 22 partial rounds in the O(t) form: the dense MDS layer (3 * 145 multiply-adds + 12 recombinations per round) becomes
// one 12-term dot product accumulated without reductions and 11 multiply-adds by the S-box output.  Entry: the true state
// after round 3's MDS with no constants added; exit: the true state with round 26's constants added.
__device__ __forceinline__ u64 gl_acc_reduce_nc(const gl_acc& s) { return gl_sub_nc2(gl_reduce128_nc(s.lo, s.hi), (u64)s.top << 32); }
__device__ __forceinline__ void poseidon_partial_rounds_nc(u64 s[12]) {
  int off = 0;
#pragma unroll 1
  for (int r = 0; r < 22; r++, off += 23) {
    const u64 x0 = poseidon_sbox_nc(gl_add_nc(s[0], d_poseidon_fast[off]));
    gl_acc acc = gl_acc_zero();
    gl_acc_mac(acc, x0, 25);
#pragma unroll
    for (int i = 0; i < 11; i++) gl_acc_mac(acc, s[1 + i], d_poseidon_fast[off + 1 + i]);
#pragma unroll
    for (int i = 0; i < 11; i++) s[1 + i] = gl_add_nc2(s[1 + i], gl_mul_nc(x0, d_poseidon_fast[off + 12 + i]));
    s[0] = gl_acc_reduce_nc(acc);
  }
  // true state = diag(1, D^) s + e: one output per iteration of a rolled loop, results shifted through a register window
  u64 o[11];
#pragma unroll
  for (int i = 0; i < 11; i++) o[i] = 0;
#pragma unroll 1
  for (int i = 0; i < 11; i++) {
    gl_acc acc = gl_acc_zero();
    acc.lo = d_poseidon_fast[506 + 122 + i];
#pragma unroll
    for (int j = 0; j < 11; j++) gl_acc_mac(acc, s[1 + j], d_poseidon_fast[506 + 11 * i + j]);
#pragma unroll
    for (int k = 0; k < 10; k++) o[k] = o[k + 1];
    o[10] = gl_acc_reduce_nc(acc);
  }
  s[0] = gl_add_nc(gl_add_nc(s[0], d_poseidon_fast[506 + 121]), d_poseidon_rc[312]);
#pragma unroll
  for (int i = 0; i < 11; i++) s[1 + i] = gl_add_nc(o[i], d_poseidon_rc[313 + i]);
}
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
  // One rolled round loop (uniform `full` branch) keeps the hot body ~24 KB so it stays in the
  // instruction cache; a fully unrolled permutation (>90 KB) stalls on instruction fetch.
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add_nc(s[i], d_poseidon_rc[i]);
#ifndef POSEIDON_MDS
#define POSEIDON_MDS poseidon_mds_dp2a
#endif
#ifndef POSEIDON_LOOP_MODE
#define POSEIDON_LOOP_MODE 0
#endif
#if POSEIDON_LOOP_MODE == 0
#pragma unroll 1
  for (int r = 0; r < 30; r++) {
    s[0] = poseidon_sbox_nc(s[0]);
    if (r < 4 || r >= 26) {
#pragma unroll
      for (int i = 1; i < 12; i++) s[i] = poseidon_sbox_nc(s[i]);
    }
    POSEIDON_MDS(s, 12 * (r + 1));
  }
#elif POSEIDON_LOOP_MODE == 1
  // full x4 | partial x22 | full x4 with one copy of each body
  int r = 0;
#pragma unroll 1
  for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
    for (int k = 0; k < 4; k++, r++) {
#pragma unroll
      for (int i = 0; i < 12; i++) s[i] = poseidon_sbox_nc(s[i]);
      poseidon_mds_nc(s, 12 * (r + 1));
    }
    if (phase == 0) {
#pragma unroll 1
      for (int k = 0; k < 22; k++, r++) {
        s[0] = poseidon_sbox_nc(s[0]);
        poseidon_mds_nc(s, 12 * (r + 1));
      }
    }
  }
#elif POSEIDON_LOOP_MODE == 3
  // full x4 | 22 partial rounds in the O(t) form | full x4, one copy of the full-round body
  int r = 0;
#pragma unroll 1
  for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
    for (int k = 0; k < 4; k++, r++) {
#pragma unroll
      for (int i = 0; i < 12; i++) s[i] = poseidon_sbox_nc(s[i]);
      poseidon_mds_nc(s, r == 3 ? 360 : 12 * (r + 1));
    }
    if (phase == 0) { poseidon_partial_rounds_nc(s); r = 26; }
  }
#elif POSEIDON_LOOP_MODE == 2
  // partial rounds unrolled by two
  int r = 0;
#pragma unroll 1
  for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
    for (int k = 0; k < 4; k++, r++) {
#pragma unroll
      for (int i = 0; i < 12; i++) s[i] = poseidon_sbox_nc(s[i]);
      poseidon_mds_nc(s, 12 * (r + 1));
    }
    if (phase == 0) {
#pragma unroll 1
      for (int k = 0; k < 11; k++, r += 2) {
        s[0] = poseidon_sbox_nc(s[0]);
        poseidon_mds_nc(s, 12 * (r + 1));
        s[0] = poseidon_sbox_nc(s[0]);
        poseidon_mds_nc(s, 12 * (r + 2));
      }
    }
  }
#endif
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
#else
  poseidon_permute_host(s);
#endif
}
