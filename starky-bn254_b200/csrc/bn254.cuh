// BN254 base field Fq on 8 x 32-bit limbs (Montgomery form) and the integer witness arithmetic of the
// reference's modular-reduction gadget, written for one GPU thread per trace row.
//
// Replaces, for trace generation (K1): `ark_bn254::Fq` uses (reference src/curves/g1/muladd.rs:136,415),
// `num_bigint` division in `generate_modular_op` / `generate_modular_zero`
// (reference src/modular/modular.rs:38-100, src/modular/modular_zero.rs:33-80) and the i64 limb-polynomial
// helpers of src/modular/pol_utils.rs.  No heap, no big-integer library: the exact quotient
// (input - output) / p is obtained by multiplying with p^-1 mod 2^288 (the division is exact and
// |quot| < 2^272), which is a 9x9-word truncated product.
#pragma once
#include "gl.cuh"

typedef long long i64;

struct Fq { u32 l[8]; };

#define FQ_P_LIST {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}
#define FQ_N0 0xe4866389u
#define FQ_R2_LIST {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u}
#define FQ_ONE_LIST {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}
#define FQ_PINV288_LIST {0x1b799c77u, 0x782df87du, 0xe1359536u, 0x6121829au, 0xe7cc257fu, 0x2750342fu, 0x6e777394u, 0x0a85dd48u, 0x5b52d390u}
#define FQ_2_256_MINUS_P_LIST {0x278302b9u, 0xc3df73e9u, 0x978e3572u, 0x687e956eu, 0x7e7ea7a2u, 0x47afba49u, 0x1ece5fd6u, 0xcf9bb18du}

HD u32 fq_p(int i) { const u32 P[8] = FQ_P_LIST; return P[i]; }

HD bool fq_geq_p(const u32* t) {
  const u32 P[8] = FQ_P_LIST;
#pragma unroll
  for (int i = 7; i >= 0; i--) { if (t[i] != P[i]) return t[i] > P[i]; }
  return true;
}
HD void fq_sub_p(u32* t) {
  const u32 P[8] = FQ_P_LIST;
  i64 c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (i64)t[i] - P[i]; t[i] = (u32)c; c >>= 32; }
}
HD Fq fq_mul(const Fq& a, const Fq& b) {
  const u32 P[8] = FQ_P_LIST;
  u32 t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u64 c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { u64 s = (u64)a.l[j] * b.l[i] + t[j] + c; t[j] = (u32)s; c = s >> 32; }
    u64 s = (u64)t[8] + c; t[8] = (u32)s; t[9] = (u32)(s >> 32);
    u32 m = t[0] * FQ_N0;
    c = ((u64)m * P[0] + t[0]) >> 32;
#pragma unroll
    for (int j = 1; j < 8; j++) { s = (u64)m * P[j] + t[j] + c; t[j - 1] = (u32)s; c = s >> 32; }
    s = (u64)t[8] + c; t[7] = (u32)s; t[8] = t[9] + (u32)(s >> 32);
  }
  if (t[8] || fq_geq_p(t)) fq_sub_p(t);
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = t[i];
  return r;
}
HD Fq fq_add(const Fq& a, const Fq& b) {
  u32 t[8]; u64 c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (u64)a.l[i] + b.l[i]; t[i] = (u32)c; c >>= 32; }
  if (c || fq_geq_p(t)) fq_sub_p(t);
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = t[i];
  return r;
}
HD Fq fq_sub(const Fq& a, const Fq& b) {
  const u32 P[8] = FQ_P_LIST;
  u32 t[8]; i64 c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (i64)a.l[i] - b.l[i]; t[i] = (u32)c; c >>= 32; }
  if (c) { u64 k = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { k += (u64)t[i] + P[i]; t[i] = (u32)k; k >>= 32; } }
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = t[i];
  return r;
}
HD Fq fq_zero() { Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
HD Fq fq_one() { const u32 O[8] = FQ_ONE_LIST; Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = O[i]; return r; }
HD bool fq_is_zero(const Fq& a) { u32 o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.l[i]; return o == 0; }
HD Fq fq_to_mont(const Fq& a) { const u32 R2[8] = FQ_R2_LIST; Fq r2;
#pragma unroll
  for (int i = 0; i < 8; i++) r2.l[i] = R2[i]; return fq_mul(a, r2); }
HD Fq fq_from_mont(const Fq& a) { Fq o = fq_zero(); o.l[0] = 1; return fq_mul(a, o); }
HD Fq fq_dbl(const Fq& a) { return fq_add(a, a); }
HD Fq fq_sqr(const Fq& a) { return fq_mul(a, a); }
HD Fq fq_inv(const Fq& a) {  // a^(p-2); a != 0
  const u32 P[8] = FQ_P_LIST;
  Fq r = fq_one();
  for (int i = 255; i >= 0; i--) {
    r = fq_sqr(r);
    u32 w = P[i >> 5] - ((i >> 5) == 0 ? 2u : 0u);  // p - 2 only changes the lowest word (0x...47 - 2, no borrow)
    if ((w >> (i & 31)) & 1) r = fq_mul(r, a);
  }
  return r;
}
// canonical 256-bit value <-> sixteen 16-bit limbs (reference src/utils/utils.rs:169-193)
HD void fq_words_to_limbs(const u32* w, i64* limbs) {
#pragma unroll
  for (int i = 0; i < 8; i++) { limbs[2 * i] = w[i] & 0xFFFF; limbs[2 * i + 1] = w[i] >> 16; }
}

// ---- modular-reduction witness (reference src/modular/modular.rs:38-100, modular_zero.rs:33-80) ----
struct ModWitness {
  u32 out_aux_red[16];  // only for modular_op
  u32 quot_abs[17];
  u32 aux_lo[31], aux_hi[31];
  bool negative;        // quot_sign = negative ? p_goldilocks - 1 : 1
};
// pol_input: 31 signed coefficients; output: canonical residue as 8 words (nullptr for the "zero" variant).
// Precondition (checked by the caller's algebra, asserted by the reference): sum_i pol_input[i] 2^(16 i) == output (mod p).
HD void modular_witness(const i64* pol_input, const u32* output_words, ModWitness& w) {
  const u32 PINV[9] = FQ_PINV288_LIST;
  const u32 P[8] = FQ_P_LIST;
  // D = input - output as two's complement over 18 words (576 bits); only the low 288 bits are needed for the quotient
  u32 d[9];
  {
    i64 carry = 0;
#pragma unroll
    for (int k = 0; k < 18; k++) {
      i64 t = carry + (k < 31 ? pol_input[k] : 0);
      if (output_words && k < 16) t -= (i64)((output_words[k >> 1] >> (16 * (k & 1))) & 0xFFFF);
      u32 dig = (u32)(t & 0xFFFF);
      carry = t >> 16;
      if (k & 1) d[k >> 1] |= dig << 16; else d[k >> 1] = dig;
    }
  }
  // q = D * p^-1 mod 2^288 (exact division)
  u32 q[9];
  {
    u64 acc = 0; u32 acc_hi = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) {
#pragma unroll
      for (int i = 0; i <= k; i++) {
        u64 pr = (u64)d[i] * PINV[k - i];
        acc += pr; if (acc < pr) acc_hi++;
      }
      q[k] = (u32)acc;
      acc = (acc >> 32) | ((u64)acc_hi << 32); acc_hi = 0;
    }
  }
  w.negative = (q[8] >> 31) != 0;
  u32 qa[9];
  if (w.negative) { u64 c = 1;
#pragma unroll
    for (int i = 0; i < 9; i++) { c += (u64)(~q[i]); qa[i] = (u32)c; c >>= 32; } }
  else {
#pragma unroll
    for (int i = 0; i < 9; i++) qa[i] = q[i]; }
  i64 quot_limbs[17];
#pragma unroll
  for (int i = 0; i < 17; i++) {
    u32 l = (qa[i >> 1] >> (16 * (i & 1))) & 0xFFFF;
    w.quot_abs[i] = l;
    quot_limbs[i] = w.negative ? -(i64)l : (i64)l;
  }
  if (output_words) {  // out_aux_red = 2^256 - p + output
    const u32 C0[8] = FQ_2_256_MINUS_P_LIST;
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (u64)C0[i] + output_words[i]; u32 v = (u32)c; c >>= 32; w.out_aux_red[2 * i] = v & 0xFFFF; w.out_aux_red[2 * i + 1] = v >> 16; }
  }
  // constr_poly = input - output - quot * p  (limb polynomials), aux = constr / (x - 2^16)
  i64 prev = 0;
#pragma unroll
  for (int k = 0; k < 31; k++) {
    i64 c = pol_input[k];
    if (output_words && k < 16) c -= (i64)((output_words[k >> 1] >> (16 * (k & 1))) & 0xFFFF);
#pragma unroll
    for (int i = 0; i < 17; i++) {
      int j = k - i;
      if (j >= 0 && j < 16) c -= quot_limbs[i] * (i64)((P[j >> 1] >> (16 * (j & 1))) & 0xFFFF);
    }
    i64 a = (k == 0) ? -(c >> 16) : ((prev - c) >> 16);   // pol_remove_root_2exp (pol_utils.rs:390-414)
    prev = a;
    i64 s = a + (1LL << 29);                              // + AUX_COEFF_ABS_MAX (modular.rs:77-79)
    w.aux_lo[k] = (u32)(s & 0xFFFF);
    w.aux_hi[k] = (u32)((s >> 16) & 0xFFFF);
  }
}
// 16x16 schoolbook product of limb arrays (pol_utils.rs:221-232), accumulated with a sign/scale
HD void pol_mul_acc(i64* res /*31*/, const i64* a, const i64* b, i64 scale) {
#pragma unroll
  for (int i = 0; i < 16; i++)
#pragma unroll
    for (int j = 0; j < 16; j++) res[i + j] += scale * a[i] * b[j];
}
