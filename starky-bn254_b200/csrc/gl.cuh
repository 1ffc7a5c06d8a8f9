// Goldilocks field (p = 2^64 - 2^32 + 1) for device and host code of the B200 prover.
// Values are kept canonical (< p) in memory and across every helper, so device results can be
// compared bit-for-bit with the CPU oracle.  Replaces plonky2_field::goldilocks_field (external
// dependency of the reference, Cargo.lock:591-593).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

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
