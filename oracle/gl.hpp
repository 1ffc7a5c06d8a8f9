// ORACLE (test infrastructure, not product code).  CPU restatement of the Goldilocks field and its
// quadratic extension as used by plonky2_field 0.1.1 (un-vendored dependency of the reference:
// Cargo.lock:591-593).  Parity status: "parity unpinned" at the prover boundary;
// the field itself is pinned by Poseidon's published KAT (tests/test_oracle_core.py).
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <cassert>

namespace orc {
typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

static const u64 GP = 0xFFFFFFFF00000001ULL;  // p = 2^64 - 2^32 + 1
static const u64 GEPS = 0xFFFFFFFFULL;        // 2^64 mod p

struct GF {
  u64 v;  // always canonical (< p)
  GF() : v(0) {}
  explicit GF(u64 x) : v(x >= GP ? x - GP : x) {}
  static GF from_i64(long long x) { return x >= 0 ? GF((u64)x) : GF(GP - (u64)(-x)); }
  static GF zero() { return GF(); }
  static GF one() { return GF(1); }
  bool operator==(const GF& o) const { return v == o.v; }
  bool operator!=(const GF& o) const { return v != o.v; }
};

// (all branch-free: the data is random, so branches would mispredict)
static inline GF operator+(GF a, GF b) {
  u64 s = a.v + b.v;
  s += (0 - (u64)(s < a.v)) & GEPS;  // wrapped past 2^64 == EPS (mod p); cannot wrap twice for canonical inputs
  s -= (0 - (u64)(s >= GP)) & GP;
  GF r; r.v = s; return r;
}
static inline GF operator-(GF a, GF b) {
  u64 d = a.v - b.v;
  d += (0 - (u64)(a.v < b.v)) & GP;
  GF r; r.v = d; return r;
}
static inline GF operator-(GF a) { GF r; r.v = (GP - a.v) & (0 - (u64)(a.v != 0)); return r; }
static inline u64 gl_reduce128(u128 x) {
  u64 lo = (u64)x, hi = (u64)(x >> 64);
  u64 hh = hi >> 32, hl = hi & GEPS;
  u64 t0 = lo - hh;
  t0 -= (0 - (u64)(lo < hh)) & GEPS;  // borrow: subtract 2^64 == EPS
  u64 t1 = hl * GEPS;
  u64 s = t0 + t1;
  s += (0 - (u64)(s < t0)) & GEPS;
  s -= (0 - (u64)(s >= GP)) & GP;
  return s;
}
static inline GF operator*(GF a, GF b) { GF r; r.v = gl_reduce128((u128)a.v * b.v); return r; }
static inline GF& operator+=(GF& a, GF b) { a = a + b; return a; }
static inline GF& operator-=(GF& a, GF b) { a = a - b; return a; }
static inline GF& operator*=(GF& a, GF b) { a = a * b; return a; }

static inline GF gl_pow(GF b, u64 e) {
  GF r = GF::one();
  while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
  return r;
}
static inline GF gl_inv(GF a) { assert(a.v != 0); return gl_pow(a, GP - 2); }
static inline GF gl_exp_pow2(GF a, int k) { for (int i = 0; i < k; i++) a = a * a; return a; }

// Field constants of plonky2_field::goldilocks_field (dependency; DESIGN.md "U1" explains how the
// (generator, 2^32-th root) pair was pinned: EXT_POWER_OF_TWO_GENERATOR^2 must equal it).
struct FieldParams {
  u64 mult_generator = 7;                       // MULTIPLICATIVE_GROUP_GENERATOR == coset_shift()
  u64 pow2_generator = 1753635133440165772ULL;  // POWER_OF_TWO_GENERATOR (order 2^32)
};
extern FieldParams g_field;
static inline GF coset_shift() { return GF(g_field.mult_generator); }
static inline GF root_of_unity(int logn) { return gl_exp_pow2(GF(g_field.pow2_generator), 32 - logn); }

// Quadratic extension F[X]/(X^2 - 7)
struct GF2 {
  GF a, b;
  GF2() {}
  GF2(GF x) : a(x), b() {}
  GF2(GF x, GF y) : a(x), b(y) {}
  explicit GF2(u64 x) : a(x), b() {}
  static GF2 zero() { return GF2(); }
  static GF2 one() { return GF2(GF::one()); }
  bool operator==(const GF2& o) const { return a == o.a && b == o.b; }
  bool operator!=(const GF2& o) const { return !(*this == o); }
};
static inline GF2 operator+(GF2 x, GF2 y) { return GF2(x.a + y.a, x.b + y.b); }
static inline GF2 operator-(GF2 x, GF2 y) { return GF2(x.a - y.a, x.b - y.b); }
static inline GF2 operator-(GF2 x) { return GF2(-x.a, -x.b); }
static inline GF2 operator*(GF2 x, GF2 y) {
  return GF2(x.a * y.a + GF(7) * (x.b * y.b), x.a * y.b + x.b * y.a);
}
static inline GF2 operator*(GF2 x, GF y) { return GF2(x.a * y, x.b * y); }
static inline GF2& operator+=(GF2& x, GF2 y) { x = x + y; return x; }
static inline GF2& operator-=(GF2& x, GF2 y) { x = x - y; return x; }
static inline GF2& operator*=(GF2& x, GF2 y) { x = x * y; return x; }
static inline GF2 gf2_inv(GF2 x) {
  // 1/(a+bX) = (a-bX)/(a^2-7b^2)
  GF n = x.a * x.a - GF(7) * (x.b * x.b);
  GF ni = gl_inv(n);
  return GF2(x.a * ni, (-x.b) * ni);
}
static inline GF2 gf2_pow(GF2 b, u64 e) {
  GF2 r = GF2::one();
  while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
  return r;
}
static inline GF2 gf2_exp_pow2(GF2 a, int k) { for (int i = 0; i < k; i++) a = a * a; return a; }

// Lift helpers so the AIR templates can be written once for P in {GF, GF2}.
template <class P> struct FieldOf;
template <> struct FieldOf<GF> { static GF c(u64 x) { return GF(x); } static GF from(GF x) { return x; } };
template <> struct FieldOf<GF2> { static GF2 c(u64 x) { return GF2(GF(x)); } static GF2 from(GF x) { return GF2(x); } };

static inline int log2_strict(size_t n) { int l = 0; while ((size_t(1) << l) < n) l++; assert((size_t(1) << l) == n); return l; }
static inline size_t reverse_bits(size_t x, int bits) {
  size_t r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | ((x >> i) & 1); } return r;
}
template <class T> static inline void reverse_index_bits_in_place(std::vector<T>& v) {
  int lb = log2_strict(v.size());
  for (size_t i = 0; i < v.size(); i++) { size_t j = reverse_bits(i, lb); if (i < j) std::swap(v[i], v[j]); }
}
}  // namespace orc
