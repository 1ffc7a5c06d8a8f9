// Goldilocks field (p = 2^64 - 2^32 + 1) for device and host code of the B200 prover.
// Values are kept canonical (< p) in memory and across every helper, so device results can be
// compared bit-for-bit with the CPU oracle.  Replaces plonky2_field::goldilocks_field (external
// dependency of the reference, Cargo.lock:591-593).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL
#define HD __host__ __device__ __forceinline__

HD u64 gl_add(u64 a, u64 b) {
  u64 s = a + b;
  if (s < a) s += GL_EPS;  // 2^64 == EPS (mod p)
  if (s >= GL_P) s -= GL_P;
  return s;
}
HD u64 gl_sub(u64 a, u64 b) { return a >= b ? a - b : a - b + GL_P; }
HD u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }
HD u64 gl_reduce128(u64 lo, u64 hi) {
  u64 hh = hi >> 32, hl = hi & GL_EPS;
  u64 t0 = lo - hh;
  if (lo < hh) t0 -= GL_EPS;
  u64 t1 = (hl << 32) - hl;  // hl * EPS
  u64 s = t0 + t1;
  if (s < t0) s += GL_EPS;
  if (s >= GL_P) s -= GL_P;
  return s;
}
HD u64 gl_mul(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
  return gl_reduce128(a * b, __umul64hi(a, b));
#else
  unsigned __int128 x = (unsigned __int128)a * b;
  return gl_reduce128((u64)x, (u64)(x >> 64));
#endif
}
HD u64 gl_sqr(u64 a) { return gl_mul(a, a); }
HD u64 gl_pow(u64 b, u64 e) {
  u64 r = 1;
  while (e) { if (e & 1) r = gl_mul(r, b); b = gl_sqr(b); e >>= 1; }
  return r;
}
HD u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }
HD u64 gl_exp_pow2(u64 a, int k) { for (int i = 0; i < k; i++) a = gl_sqr(a); return a; }
HD u64 gl_from_i64(long long x) { return x >= 0 ? (u64)x : GL_P - (u64)(-x); }

// ---- device-only lazy arithmetic: operands and results are ARBITRARY 64-bit representatives (not necessarily < p). ----
// Used by the hot kernels (Poseidon, NTT) between canonical loads and canonical stores: skipping the conditional
// subtraction of p after every operation saves ~40% of their integer instructions.
#ifdef __CUDACC__
// (x3:x2:x1:x0) mod p as an arbitrary 64-bit representative, using 2^64 = 2^32 - 1 and 2^96 = -1:
//   T = (x1:x0) - (x2 + x3) + x2 * 2^32  lies in (-2^33, 2^65 - 2^33]; its 64-bit wrap count d in {-1, 0, 1} is the
//   carry of the addition plus the (negative) borrow of the subtraction, and T - d * 2^64 + d * (2^32 - 1) cannot wrap again.
// 11 instructions; the product comes as words from gl_mul128_words / gl_sqr128_words below.
__device__ __forceinline__ u64 gl_reduce_words_nc(u32 x0, u32 x1, u32 x2, u32 x3) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 s, cs, b, d, e, f;\n\t"
      "add.cc.u32 s, %4, %5;\n\t"        // x2 + x3 = cs:s
      "addc.u32 cs, 0, 0;\n\t"
      "sub.cc.u32 %0, %2, s;\n\t"
      "subc.cc.u32 %1, %3, cs;\n\t"
      "subc.u32 b, 0, 0;\n\t"            // 0 or -1
      "add.cc.u32 %1, %1, %4;\n\t"       // + x2 * 2^32
      "addc.u32 d, b, 0;\n\t"            // d = carry + b
      "neg.s32 e, d;\n\t"                // d * (2^32 - 1) as a two's-complement 64-bit value f:e
      "shr.s32 f, d, 31;\n\t"
      "add.cc.u32 %0, %0, e;\n\t"
      "addc.u32 %1, %1, f;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3));
  return ((u64)r1 << 32) | r0;
}
__device__ __forceinline__ u64 gl_reduce128_nc(u64 lo, u64 hi) { return gl_reduce_words_nc((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32)); }
// 128-bit product as four 32-bit words from exactly four 32 x 32 -> 64 multiplications (IMAD.WIDE) and five carry adds.  The
// compiler's `a * b` + `__umul64hi(a, b)` pair costs five IMAD.WIDE and two IMAD (it forms the low product twice), i.e.
// 24 cycles of the multiplier pipe per product against 16 here (tools/microbench: IMAD 2 cycles, IMAD.WIDE 4 per warp).
__device__ __forceinline__ void gl_mul128_words(u64 a, u64 b, u32& x0, u32& x1, u32& x2, u32& x3) {
  asm("{\n\t"
      ".reg .u64 p, q, r, s;\n\t"
      ".reg .u32 ph, ql, qh, rl, rh, sl, sh;\n\t"
      "mul.wide.u32 p, %4, %6;\n\t"     // a_lo b_lo
      "mul.wide.u32 q, %4, %7;\n\t"     // a_lo b_hi
      "mul.wide.u32 r, %5, %6;\n\t"     // a_hi b_lo
      "mul.wide.u32 s, %5, %7;\n\t"     // a_hi b_hi
      "mov.b64 {%0, ph}, p;\n\t"
      "mov.b64 {ql, qh}, q;\n\t"
      "mov.b64 {rl, rh}, r;\n\t"
      "mov.b64 {sl, sh}, s;\n\t"
      "add.cc.u32 %1, ph, ql;\n\t"
      "addc.cc.u32 %2, qh, sl;\n\t"
      "addc.u32 %3, sh, 0;\n\t"
      "add.cc.u32 %1, %1, rl;\n\t"
      "addc.cc.u32 %2, %2, rh;\n\t"
      "addc.u32 %3, %3, 0;\n\t"
      "}"
      : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)       // every input is consumed by the four multiplications before an output is written
      : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
}
// a^2: three multiplications, the cross product doubled by a funnel shift.
__device__ __forceinline__ void gl_sqr128_words(u64 a, u32& x0, u32& x1, u32& x2, u32& x3) {
  asm("{\n\t"
      ".reg .u64 p, q, s;\n\t"
      ".reg .u32 ph, ql, qh, sl, sh, d0, d1, d2;\n\t"
      "mul.wide.u32 p, %4, %4;\n\t"     // a_lo^2
      "mul.wide.u32 q, %4, %5;\n\t"     // a_lo a_hi
      "mul.wide.u32 s, %5, %5;\n\t"     // a_hi^2
      "mov.b64 {%0, ph}, p;\n\t"
      "mov.b64 {ql, qh}, q;\n\t"
      "mov.b64 {sl, sh}, s;\n\t"
      "shl.b32 d0, ql, 1;\n\t"                   // 2 q = d2:d1:d0
      "shf.l.clamp.b32 d1, ql, qh, 1;\n\t"
      "shr.u32 d2, qh, 31;\n\t"
      "add.cc.u32 %1, ph, d0;\n\t"
      "addc.cc.u32 %2, sl, d1;\n\t"
      "addc.u32 %3, sh, d2;\n\t"
      "}"
      : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
      : "r"((u32)a), "r"((u32)(a >> 32)));
}
// a * b mod p as an arbitrary 64-bit representative; a, b arbitrary u64.
__device__ __forceinline__ u64 gl_mul_nc(u64 a, u64 b) { u32 x0, x1, x2, x3; gl_mul128_words(a, b, x0, x1, x2, x3); return gl_reduce_words_nc(x0, x1, x2, x3); }
__device__ __forceinline__ u64 gl_sqr_nc(u64 a) { u32 x0, x1, x2, x3; gl_sqr128_words(a, x0, x1, x2, x3); return gl_reduce_words_nc(x0, x1, x2, x3); }
// a + b mod p, a arbitrary u64, b canonical (< p); result arbitrary u64 representative.
__device__ __forceinline__ u64 gl_add_nc(u64 a, u64 b) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 c;\n\t"
      "add.cc.u32 %0, %2, %4;\n\t"
      "addc.cc.u32 %1, %3, %5;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "neg.s32 c, c;\n\t"
      "add.cc.u32 %0, %0, c;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
  return ((u64)r1 << 32) | r0;
}
__device__ __forceinline__ u64 gl_canon(u64 a) { return a >= GL_P ? a - GL_P : a; }
// a + b mod p, both arbitrary u64: a wrapped sum gets + (2^64 mod p); that can wrap once more (only if both operands
// were >= 2^64 - 2^32), after which no further wrap is possible.
__device__ __forceinline__ u64 gl_add_nc2(u64 a, u64 b) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 c;\n\t"
      "add.cc.u32 %0, %2, %4;\n\t"
      "addc.cc.u32 %1, %3, %5;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "neg.s32 c, c;\n\t"
      "add.cc.u32 %0, %0, c;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "neg.s32 c, c;\n\t"
      "add.cc.u32 %0, %0, c;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
  return ((u64)r1 << 32) | r0;
}
// Sum of products without intermediate reductions: a 160-bit accumulator takes up to 2^32 terms a*b (a, b arbitrary u64);
// gl_acc_reduce folds it once with 2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32 (mod p).  7 + 3 instructions per term
// instead of the 18 + 6 of a reduced multiply-add.
struct gl_acc { u64 lo, hi; u32 top; };
__device__ __forceinline__ gl_acc gl_acc_zero() { gl_acc s; s.lo = 0; s.hi = 0; s.top = 0; return s; }
__device__ __forceinline__ void gl_acc_mac(gl_acc& s, u64 a, u64 b) {
  u32 x0, x1, x2, x3;
  gl_mul128_words(a, b, x0, x1, x2, x3);   // four multiplications (the compiler's a * b + __umul64hi pair costs five and two IMAD)
  const u64 pl = ((u64)x1 << 32) | x0, ph = ((u64)x3 << 32) | x2;
  asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u32 %2, %2, 0;" : "+l"(s.lo), "+l"(s.hi), "+r"(s.top) : "l"(pl), "l"(ph));
}
__device__ __forceinline__ u64 gl_sub_nc2(u64 a, u64 b);
__device__ __forceinline__ u64 gl_acc_reduce(const gl_acc& s) {   // canonical result
  return gl_canon(gl_sub_nc2(gl_reduce128_nc(s.lo, s.hi), (u64)s.top << 32));
}
// a - b mod p, both arbitrary u64: a borrowed difference gets - (2^64 mod p), which can borrow once more.
__device__ __forceinline__ u64 gl_sub_nc2(u64 a, u64 b) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 m;\n\t"
      "sub.cc.u32 %0, %2, %4;\n\t"
      "subc.cc.u32 %1, %3, %5;\n\t"
      "subc.u32 m, 0, 0;\n\t"            // 0 or 0xFFFFFFFF (= 2^64 mod p)
      "sub.cc.u32 %0, %0, m;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, m;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(r0), "=&r"(r1)
      : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
  return ((u64)r1 << 32) | r0;
}
#endif

// Quadratic extension F[X]/(X^2 - 7) (plonky2 `QuadraticExtension<GoldilocksField>`, W = 7).
struct gl2 { u64 a, b; };
HD gl2 gl2_make(u64 a, u64 b) { gl2 r; r.a = a; r.b = b; return r; }
HD gl2 gl2_add(gl2 x, gl2 y) { return gl2_make(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
HD gl2 gl2_sub(gl2 x, gl2 y) { return gl2_make(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
HD gl2 gl2_mul(gl2 x, gl2 y) {
  u64 bb = gl_mul(x.b, y.b);
  u64 w = gl_mul(bb, 7);
  return gl2_make(gl_add(gl_mul(x.a, y.a), w), gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a)));
}
HD gl2 gl2_mul_base(gl2 x, u64 y) { return gl2_make(gl_mul(x.a, y), gl_mul(x.b, y)); }
HD gl2 gl2_inv(gl2 x) {
  u64 n = gl_sub(gl_sqr(x.a), gl_mul(7, gl_sqr(x.b)));
  u64 ni = gl_inv(n);
  return gl2_make(gl_mul(x.a, ni), gl_mul(gl_neg(x.b), ni));
}
HD gl2 gl2_pow(gl2 b, u64 e) {
  gl2 r = gl2_make(1, 0);
  while (e) { if (e & 1) r = gl2_mul(r, b); b = gl2_mul(b, b); e >>= 1; }
  return r;
}
HD bool gl2_eq(gl2 x, gl2 y) { return x.a == y.a && x.b == y.b; }

// Field constants of plonky2_field (DESIGN.md "U1"): multiplicative generator (= coset shift) 7 and
// the 2^32-th root of unity 7^((p-1)/2^32).
#define GL_MULT_GENERATOR 7ULL
#define GL_POW2_GENERATOR 1753635133440165772ULL
HD u64 gl_root_of_unity(int logn) { return gl_exp_pow2(GL_POW2_GENERATOR, 32 - logn); }

HD u32 bitrev32(u32 x, int bits) {
#ifdef __CUDA_ARCH__
  return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
  u32 r = 0; for (int i = 0; i < bits; i++) r = (r << 1) | ((x >> i) & 1); return r;
#endif
}
