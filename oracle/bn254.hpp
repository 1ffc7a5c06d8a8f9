// ORACLE (test infrastructure).  Big-integer and BN254 base-field helpers standing in for the
// reference's `num_bigint::BigInt` and `ark_bn254::Fq` uses (reference src/modular/modular.rs:38-100,
// src/utils/utils.rs:124-167, src/curves/g1/muladd.rs:136,415).  Plain arithmetic: results are
// canonical residues, so any correct implementation yields the reference's limbs.
#pragma once
#include <cstdint>
#include <cstring>
#include <cassert>
#include <array>
#include "gl.hpp"

namespace orc {
typedef long long i64;
static const int N_LIMBS = 16;   // reference src/constants.rs:2
static const int LIMB_BITS = 16; // reference src/constants.rs:1

// BN254 base modulus, 16-bit limbs LSB first (reference modular.rs:298-309).
static const i64 BN254_P_LIMBS[16] = {64839, 55420, 35862, 15392, 51853, 26737, 27281, 38785,
                                      22621, 33153, 17846, 47184, 41001, 57649, 20082, 12388};

// ---- signed magnitude big integer, 32-bit limbs, fixed capacity (enough for |x| < 2^700) ----
struct Big {
  static const int W = 22;
  bool neg = false;
  u32 m[W];
  Big() { memset(m, 0, sizeof m); }
  bool is_zero() const { for (int i = 0; i < W; i++) if (m[i]) return false; return true; }
  int top() const { for (int i = W - 1; i >= 0; i--) if (m[i]) return i + 1; return 0; }
};
static inline int mag_cmp(const u32* a, const u32* b, int n) {
  for (int i = n - 1; i >= 0; i--) { if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1; } return 0;
}
static inline void mag_add(u32* r, const u32* a, const u32* b, int n) {
  u64 c = 0; for (int i = 0; i < n; i++) { c += (u64)a[i] + b[i]; r[i] = (u32)c; c >>= 32; }
}
static inline void mag_sub(u32* r, const u32* a, const u32* b, int n) {  // a >= b
  i64 c = 0; for (int i = 0; i < n; i++) { c += (i64)a[i] - b[i]; r[i] = (u32)c; c >>= 32; }
}
static inline Big big_add(const Big& a, const Big& b) {
  Big r;
  if (a.neg == b.neg) { mag_add(r.m, a.m, b.m, Big::W); r.neg = a.neg; }
  else {
    int c = mag_cmp(a.m, b.m, Big::W);
    if (c == 0) return r;
    if (c > 0) { mag_sub(r.m, a.m, b.m, Big::W); r.neg = a.neg; }
    else { mag_sub(r.m, b.m, a.m, Big::W); r.neg = b.neg; }
  }
  if (r.is_zero()) r.neg = false;
  return r;
}
static inline Big big_neg(Big a) { if (!a.is_zero()) a.neg = !a.neg; return a; }
static inline Big big_sub(const Big& a, const Big& b) { return big_add(a, big_neg(b)); }
// Knuth algorithm D on magnitudes: q = floor(num/den), r = num - q*den.
static inline void mag_divmod(const u32* num, int nn, const u32* den, int dn, u32* q, u32* r) {
  // normalise
  int s = __builtin_clz(den[dn - 1]);
  u32 v[Big::W + 1], u[Big::W + 2];
  memset(v, 0, sizeof v); memset(u, 0, sizeof u);
  for (int i = dn - 1; i > 0; i--) v[i] = s ? (den[i] << s) | (den[i - 1] >> (32 - s)) : den[i];
  v[0] = den[0] << s;
  u[nn] = s ? num[nn - 1] >> (32 - s) : 0;
  for (int i = nn - 1; i > 0; i--) u[i] = s ? (num[i] << s) | (num[i - 1] >> (32 - s)) : num[i];
  u[0] = num[0] << s;
  for (int j = nn - dn; j >= 0; j--) {
    u64 top = ((u64)u[j + dn] << 32) | u[j + dn - 1];
    u64 qhat = top / v[dn - 1], rhat = top % v[dn - 1];
    while (qhat >= (1ULL << 32) || (dn >= 2 && qhat * v[dn - 2] > ((rhat << 32) | u[j + dn - 2]))) {
      qhat--; rhat += v[dn - 1]; if (rhat >= (1ULL << 32)) break;
    }
    i64 borrow = 0; u64 carry = 0;
    for (int i = 0; i < dn; i++) {
      u64 p = qhat * v[i] + carry; carry = p >> 32;
      i64 t = (i64)u[i + j] - borrow - (i64)(p & 0xFFFFFFFFULL);
      u[i + j] = (u32)t; borrow = t < 0 ? 1 : 0;
    }
    i64 t = (i64)u[j + dn] - borrow - (i64)carry;
    u[j + dn] = (u32)t;
    if (t < 0) {  // add back
      qhat--;
      u64 c = 0;
      for (int i = 0; i < dn; i++) { c += (u64)u[i + j] + v[i]; u[i + j] = (u32)c; c >>= 32; }
      u[j + dn] += (u32)c;
    }
    q[j] = (u32)qhat;
  }
  for (int i = 0; i < dn; i++) r[i] = s ? (u[i] >> s) | ((u64)u[i + 1] << (32 - s)) : u[i];
}

// reference utils.rs:124-151 `columns_to_bigint`: sum_i limbs[i] * 2^(16 i) as a signed integer.
template <int N> static inline Big columns_to_bigint(const i64 (&limbs)[N]) {
  // two's-complement accumulation over 16-bit digits, then convert to sign/magnitude
  const int D = 2 * Big::W;
  u32 dig[D];
  i64 carry = 0;
  for (int i = 0; i < D; i++) {
    __int128 t = (__int128)carry + (i < N ? limbs[i] : 0);
    dig[i] = (u32)((u64)t & 0xFFFF);
    carry = (i64)(t >> 16);
  }
  assert(carry == 0 || carry == -1);
  Big r;
  for (int i = 0; i < Big::W; i++) r.m[i] = dig[2 * i] | (dig[2 * i + 1] << 16);
  if (carry == -1) {  // negative: magnitude = 2^(32W) - value
    u64 c = 1;
    for (int i = 0; i < Big::W; i++) { c += (u64)(~r.m[i]); r.m[i] = (u32)c; c >>= 32; }
    r.neg = !r.is_zero();
  }
  return r;
}
// reference utils.rs:153-167 `bigint_to_columns`: 16-bit limbs of |num|, every limb negated if num<0.
template <int N> static inline void bigint_to_columns(const Big& num, i64 (&out)[N]) {
  assert(num.top() * 32 <= 16 * N + 16);
  for (int i = 0; i < N; i++) {
    u32 w = num.m[i / 2];
    i64 l = (i & 1) ? (w >> 16) : (w & 0xFFFF);
    out[i] = num.neg ? -l : l;
  }
  for (int i = N; i < 2 * Big::W; i++) { u32 w = num.m[i / 2]; assert(((i & 1) ? (w >> 16) : (w & 0xFFFF)) == 0); }
}
static inline Big bn254_modulus_big() {
  Big p; for (int i = 0; i < 16; i++) p.m[i / 2] |= (u32)BN254_P_LIMBS[i] << (16 * (i & 1)); return p;
}

// ---- Fq: BN254 base field, Montgomery form on 4x64 limbs ----
struct U256 { u64 w[4]; };
static inline bool u256_geq(const U256& a, const U256& b) {
  for (int i = 3; i >= 0; i--) { if (a.w[i] != b.w[i]) return a.w[i] > b.w[i]; }
  return true;
}
static inline U256 u256_sub(const U256& a, const U256& b) {
  U256 r; u128 br = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a.w[i] - b.w[i] - br; r.w[i] = (u64)t; br = (t >> 64) & 1; }
  return r;
}
static inline U256 u256_add(const U256& a, const U256& b, u64* carry) {
  U256 r; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a.w[i] + b.w[i]; r.w[i] = (u64)c; c >>= 64; }
  *carry = (u64)c; return r;
}
struct FqCtx {
  U256 p, r2, one; u64 n0;
  FqCtx() {
    for (int i = 0; i < 4; i++) { p.w[i] = 0; for (int k = 0; k < 4; k++) p.w[i] |= (u64)BN254_P_LIMBS[4 * i + k] << (16 * k); }
    u64 x = 1; for (int i = 0; i < 6; i++) x *= 2 - p.w[0] * x;  // p^-1 mod 2^64
    n0 = (u64)0 - x;
    U256 t = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; i++) { u64 c; t = u256_add(t, t, &c); if (c || u256_geq(t, p)) t = u256_sub(t, p); if (i == 255) one = t; }
    r2 = t;
  }
};
static const FqCtx& fqctx() { static FqCtx c; return c; }
struct Fq {
  U256 v;  // Montgomery form
  bool operator==(const Fq& o) const { return !memcmp(v.w, o.v.w, 32); }
};
static inline Fq fq_mont_mul(const Fq& a, const Fq& b) {
  const FqCtx& c = fqctx();
  u64 t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 carry = 0;
    for (int j = 0; j < 4; j++) { carry += (u128)a.v.w[j] * b.v.w[i] + t[j]; t[j] = (u64)carry; carry >>= 64; }
    carry += t[4]; t[4] = (u64)carry; t[5] = (u64)(carry >> 64);
    u64 m = t[0] * c.n0;
    carry = (u128)m * c.p.w[0] + t[0]; carry >>= 64;
    for (int j = 1; j < 4; j++) { carry += (u128)m * c.p.w[j] + t[j]; t[j - 1] = (u64)carry; carry >>= 64; }
    carry += t[4]; t[3] = (u64)carry; t[4] = t[5] + (u64)(carry >> 64); t[5] = 0;
  }
  Fq r; for (int i = 0; i < 4; i++) r.v.w[i] = t[i];
  if (t[4] || u256_geq(r.v, c.p)) r.v = u256_sub(r.v, c.p);
  return r;
}
static inline Fq operator*(const Fq& a, const Fq& b) { return fq_mont_mul(a, b); }
static inline Fq operator+(const Fq& a, const Fq& b) {
  u64 c; Fq r; r.v = u256_add(a.v, b.v, &c); if (c || u256_geq(r.v, fqctx().p)) r.v = u256_sub(r.v, fqctx().p); return r;
}
static inline Fq operator-(const Fq& a, const Fq& b) {
  Fq r; if (u256_geq(a.v, b.v)) r.v = u256_sub(a.v, b.v); else { u64 c; r.v = u256_sub(u256_add(a.v, fqctx().p, &c), b.v); } return r;
}
static inline Fq fq_zero() { Fq r; memset(&r, 0, sizeof r); return r; }
static inline Fq fq_one() { Fq r; r.v = fqctx().one; return r; }
static inline Fq operator-(const Fq& a) { return fq_zero() - a; }
static inline bool fq_is_zero(const Fq& a) { return !(a.v.w[0] | a.v.w[1] | a.v.w[2] | a.v.w[3]); }
static inline Fq fq_from_u256(const U256& x) { Fq a; a.v = x; Fq r2; r2.v = fqctx().r2; return fq_mont_mul(a, r2); }
static inline U256 fq_to_u256(const Fq& a) { Fq o; o.v = U256{{1, 0, 0, 0}}; return fq_mont_mul(a, o).v; }
static inline Fq fq_from_u64(u64 x) { return fq_from_u256(U256{{x, 0, 0, 0}}); }
static inline Fq fq_pow(Fq b, const U256& e) {
  Fq r = fq_one();
  for (int i = 255; i >= 0; i--) { r = r * r; if ((e.w[i / 64] >> (i % 64)) & 1) r = r * b; }
  return r;
}
static inline Fq fq_inv(const Fq& a) {
  assert(!fq_is_zero(a));  // arkworks panics on division by zero (reference g1/muladd.rs:136)
  U256 e = u256_sub(fqctx().p, U256{{2, 0, 0, 0}});
  return fq_pow(a, e);
}
// reference utils.rs:169-172 `fq_to_columns` / :189-193 `columns_to_fq`
static inline void fq_to_columns(const Fq& x, i64 (&out)[16]) {
  U256 v = fq_to_u256(x);
  for (int i = 0; i < 16; i++) out[i] = (v.w[i / 4] >> (16 * (i % 4))) & 0xFFFF;
}
static inline Fq columns_to_fq(const i64 (&c)[16]) {
  U256 v = {{0, 0, 0, 0}};
  for (int i = 0; i < 16; i++) { assert(c[i] >= 0 && c[i] < 65536); v.w[i / 4] |= (u64)c[i] << (16 * (i % 4)); }
  assert(!u256_geq(v, fqctx().p));
  return fq_from_u256(v);
}
}  // namespace orc
