// K1 per-row witness programs (one GPU thread per trace row), restating the reference's
// `generate_*` functions over canonical residues.  `W` is a column writer: w(col, value).
#pragma once
#include "bn254.cuh"

HD u64 sign_to_gl(bool negative) { return negative ? GL_P - 1 : 1; }

template <class W> HD void write_limbs16(W& w, int col, const u32* words) {
#pragma unroll
  for (int i = 0; i < 8; i++) { w(col + 2 * i, words[i] & 0xFFFF); w(col + 2 * i + 1, words[i] >> 16); }
}
// reference src/modular/modular.rs:273-279 `write_modulus_aux`: out_aux_red16 | quot_abs17 | lo31 | hi31
template <class W> HD void write_modulus_aux(W& w, int col, const ModWitness& m) {
  for (int i = 0; i < 16; i++) w(col + i, m.out_aux_red[i]);
  for (int i = 0; i < 17; i++) w(col + 16 + i, m.quot_abs[i]);
  for (int i = 0; i < 31; i++) w(col + 33 + i, m.aux_lo[i]);
  for (int i = 0; i < 31; i++) w(col + 64 + i, m.aux_hi[i]);
}
// reference src/modular/modular_zero.rs:174-179 `write_modulus_aux_zero`: quot_abs17 | lo31 | hi31
template <class W> HD void write_modulus_aux_zero(W& w, int col, const ModWitness& m) {
  for (int i = 0; i < 17; i++) w(col + i, m.quot_abs[i]);
  for (int i = 0; i < 31; i++) w(col + 17 + i, m.aux_lo[i]);
  for (int i = 0; i < 31; i++) w(col + 48 + i, m.aux_hi[i]);
}

HD void u64x4_to_words(const u64* v, u32* w) {
#pragma unroll
  for (int i = 0; i < 4; i++) { w[2 * i] = (u32)v[i]; w[2 * i + 1] = (u32)(v[i] >> 32); }
}
HD Fq fq_from_words(const u32* w) { Fq a;
#pragma unroll
  for (int i = 0; i < 8; i++) a.l[i] = w[i]; return fq_to_mont(a); }
HD void fq_to_words(const Fq& m, u32* w) { Fq a = fq_from_mont(m);
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = a.l[i]; }

// ---- ModularStark row (reference src/modular/modular.rs:385-428): in0 | in1 | out | aux(95) | sign | filter ----
template <class W> HD void modular_stark_row(const u64* io /* input0[4], input1[4] */, W& w) {
  u32 a[8], b[8], o[8];
  u64x4_to_words(io, a); u64x4_to_words(io + 4, b);
  Fq braw;
#pragma unroll
  for (int i = 0; i < 8; i++) braw.l[i] = b[i];
  Fq out = fq_mul(fq_from_words(a), braw);  // (aR) * b / R = a*b
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = out.l[i];
  i64 al[16], bl[16], pol[31];
  fq_words_to_limbs(a, al); fq_words_to_limbs(b, bl);
  for (int i = 0; i < 31; i++) pol[i] = 0;
  pol_mul_acc(pol, al, bl, 1);
  ModWitness m;
  modular_witness(pol, o, m);
  write_limbs16(w, 0, a); write_limbs16(w, 16, b); write_limbs16(w, 32, o);
  write_modulus_aux(w, 48, m);
  w(143, sign_to_gl(m.negative));
  w(144, 1);
}

// ---- flag columns in closed form (reference src/utils/flags.rs:46-134 generates them row by row) ----
// Row r of a 512-row block with exponent limbs e[0..8): segment s = r/64 consumes limb s; local row q = r%64
// handles bit j = q/2.  Columns: is_final, is_rotate, a, b, filtered_bit, bit, limbs[8].
HD void flags_row(const u32* e, int r, u64* out /*14*/) {
  const int s = r >> 6, q = r & 63, j = q >> 1;
  const u32 limb = e[s];
  const u64 bit = (limb >> j) & 1;
  const u64 a = q & 1, b = 1 - a;
  out[0] = (r == 511) ? 1 : 0;
  out[1] = (q == 62) ? 1 : 0;
  out[2] = a; out[3] = b; out[4] = bit * b; out[5] = bit;
  if (q == 63) {  // row after the rotation: every limb moved down one place, the next limb is still whole
    for (int c = 0; c < 8; c++) out[6 + c] = (s + 1 + c < 8) ? e[s + 1 + c] : 0;
  } else {
    out[6] = (j == 31) ? 0 : (limb >> (j + 1));
    for (int c = 1; c < 8; c++) out[6 + c] = (s + c < 8) ? e[s + c] : 0;
  }
}

// ---- G1 row (reference src/curves/g1/muladd.rs:124-177 `generate_g1_add`, :409-460 `generate_g1_double`,
//      :61-75 default; row layout src/curves/g1/exp.rs:165-230): a.x a.y b.x b.y | G1Output(320) ----
#define G1_OP_NONE 0
#define G1_OP_ADD 1
#define G1_OP_DOUBLE 2
// ax..by: canonical coordinate words.  Returns false if the slope denominator is zero (the reference panics there).
template <class W> HD bool g1_row(const u32* ax, const u32* ay, const u32* bx, const u32* by, int op, W& w) {
  write_limbs16(w, 0, ax); write_limbs16(w, 16, ay); write_limbs16(w, 32, bx); write_limbs16(w, 48, by);
  const int o = 64;
  if (op == G1_OP_NONE) {
    for (int i = 0; i < 317; i++) w(o + i, 0);
    w(o + 317, 1); w(o + 318, 1); w(o + 319, 1);
    return true;
  }
  Fq x1 = fq_from_words(ax), y1 = fq_from_words(ay), x2, lambda;
  i64 x1l[16], y1l[16], x2l[16], ll[16], t1[16], pol[31];
  fq_words_to_limbs(ax, x1l); fq_words_to_limbs(ay, y1l);
  u32 lw[8];
  if (op == G1_OP_ADD) {
    x2 = fq_from_words(bx);
    Fq y2 = fq_from_words(by);
    Fq dx = fq_sub(x2, x1);
    if (fq_is_zero(dx)) return false;
    lambda = fq_mul(fq_sub(y2, y1), fq_inv(dx));
    fq_to_words(lambda, lw); fq_words_to_limbs(lw, ll);
    fq_words_to_limbs(bx, x2l);
    i64 y2l[16]; fq_words_to_limbs(by, y2l);
    // zero_pol = lambda * (x2 - x1) - (y2 - y1)
    for (int i = 0; i < 16; i++) t1[i] = x2l[i] - x1l[i];
    for (int i = 0; i < 31; i++) pol[i] = 0;
    pol_mul_acc(pol, ll, t1, 1);
    for (int i = 0; i < 16; i++) pol[i] -= y2l[i] - y1l[i];
  } else {
    x2 = x1;
    Fq den = fq_dbl(y1);
    if (fq_is_zero(den)) return false;
    Fq x1sq = fq_sqr(x1);
    lambda = fq_mul(fq_add(fq_dbl(x1sq), x1sq), fq_inv(den));
    fq_to_words(lambda, lw); fq_words_to_limbs(lw, ll);
    for (int i = 0; i < 16; i++) x2l[i] = x1l[i];
    // zero_pol = 2 * lambda * y - 3 * x * x
    for (int i = 0; i < 31; i++) pol[i] = 0;
    pol_mul_acc(pol, ll, y1l, 2);
    pol_mul_acc(pol, x1l, x1l, -3);
  }
  ModWitness m;
  modular_witness(pol, nullptr, m);
  write_limbs16(w, o, lw);
  write_modulus_aux_zero(w, o + 48, m);
  w(o + 317, sign_to_gl(m.negative));
  // new_x = lambda^2 - x1 - x2
  Fq nx = fq_sub(fq_sub(fq_sqr(lambda), x1), x2);
  u32 nxw[8]; fq_to_words(nx, nxw);
  for (int i = 0; i < 31; i++) pol[i] = 0;
  pol_mul_acc(pol, ll, ll, 1);
  for (int i = 0; i < 16; i++) pol[i] -= x1l[i] + x2l[i];
  modular_witness(pol, nxw, m);
  write_limbs16(w, o + 16, nxw);
  write_modulus_aux(w, o + 127, m);
  w(o + 318, sign_to_gl(m.negative));
  // new_y = lambda * (x1 - new_x) - y1
  Fq ny = fq_sub(fq_mul(lambda, fq_sub(x1, nx)), y1);
  u32 nyw[8]; fq_to_words(ny, nyw);
  i64 nxl[16]; fq_words_to_limbs(nxw, nxl);
  for (int i = 0; i < 16; i++) t1[i] = x1l[i] - nxl[i];
  for (int i = 0; i < 31; i++) pol[i] = 0;
  pol_mul_acc(pol, ll, t1, 1);
  for (int i = 0; i < 16; i++) pol[i] -= y1l[i];
  modular_witness(pol, nyw, m);
  write_limbs16(w, o + 32, nyw);
  write_modulus_aux(w, o + 222, m);
  w(o + 319, sign_to_gl(m.negative));
  return true;
}

// ---- Jacobian arithmetic for the exponentiation chain (y^2 = x^3 + 3, a = 0), Montgomery coordinates ----
struct G1Jac { Fq x, y, z; };
HD G1Jac g1_jac_dbl(const G1Jac& p) {
  Fq A = fq_sqr(p.x), B = fq_sqr(p.y), C = fq_sqr(B);
  Fq t = fq_add(p.x, B);
  Fq D = fq_dbl(fq_sub(fq_sub(fq_sqr(t), A), C));
  Fq E = fq_add(fq_dbl(A), A), Fv = fq_sqr(E);
  G1Jac r;
  r.x = fq_sub(Fv, fq_dbl(D));
  Fq c8 = fq_dbl(fq_dbl(fq_dbl(C)));
  r.y = fq_sub(fq_mul(E, fq_sub(D, r.x)), c8);
  r.z = fq_dbl(fq_mul(p.y, p.z));
  return r;
}
HD G1Jac g1_jac_add(const G1Jac& p, const G1Jac& q) {
  Fq z1z1 = fq_sqr(p.z), z2z2 = fq_sqr(q.z);
  Fq u1 = fq_mul(p.x, z2z2), u2 = fq_mul(q.x, z1z1);
  Fq s1 = fq_mul(fq_mul(p.y, q.z), z2z2), s2 = fq_mul(fq_mul(q.y, p.z), z1z1);
  Fq h = fq_sub(u2, u1);
  Fq i = fq_sqr(fq_dbl(h)), j = fq_mul(h, i);
  Fq rr = fq_dbl(fq_sub(s2, s1)), v = fq_mul(u1, i);
  G1Jac r;
  r.x = fq_sub(fq_sub(fq_sqr(rr), j), fq_dbl(v));
  r.y = fq_sub(fq_mul(rr, fq_sub(v, r.x)), fq_dbl(fq_mul(s1, j)));
  r.z = fq_mul(fq_sub(fq_sub(fq_sqr(fq_add(p.z, q.z)), z1z1), z2z2), h);
  return r;
}
