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
// The chain kernels run one warp per scheduler, so nothing hides instruction fetch: with every multiplication inlined
// one add + double is ~150 KB of code, far beyond the instruction cache.  The formulas therefore call ONE out-of-line
// copy of the Montgomery product.
__host__ __device__ __noinline__ Fq fq_mul_call(const Fq& a, const Fq& b) { return fq_mul(a, b); }
struct G1Jac { Fq x, y, z; };
HD Fq fq_sqr_call(const Fq& a) { return fq_mul_call(a, a); }
HD G1Jac g1_jac_dbl(const G1Jac& p) {
  Fq A = fq_sqr_call(p.x), B = fq_sqr_call(p.y), C = fq_sqr_call(B);
  Fq t = fq_add(p.x, B);
  Fq D = fq_dbl(fq_sub(fq_sub(fq_sqr_call(t), A), C));
  Fq E = fq_add(fq_dbl(A), A), Fv = fq_sqr_call(E);
  G1Jac r;
  r.x = fq_sub(Fv, fq_dbl(D));
  Fq c8 = fq_dbl(fq_dbl(fq_dbl(C)));
  r.y = fq_sub(fq_mul_call(E, fq_sub(D, r.x)), c8);
  r.z = fq_dbl(fq_mul_call(p.y, p.z));
  return r;
}
HD G1Jac g1_jac_add(const G1Jac& p, const G1Jac& q) {
  Fq z1z1 = fq_sqr_call(p.z), z2z2 = fq_sqr_call(q.z);
  Fq u1 = fq_mul_call(p.x, z2z2), u2 = fq_mul_call(q.x, z1z1);
  Fq s1 = fq_mul_call(fq_mul_call(p.y, q.z), z2z2), s2 = fq_mul_call(fq_mul_call(q.y, p.z), z1z1);
  Fq h = fq_sub(u2, u1);
  Fq i = fq_sqr_call(fq_dbl(h)), j = fq_mul_call(h, i);
  Fq rr = fq_dbl(fq_sub(s2, s1)), v = fq_mul_call(u1, i);
  G1Jac r;
  r.x = fq_sub(fq_sub(fq_sqr_call(rr), j), fq_dbl(v));
  r.y = fq_sub(fq_mul_call(rr, fq_sub(v, r.x)), fq_dbl(fq_mul_call(s1, j)));
  r.z = fq_mul_call(fq_sub(fq_sub(fq_sqr_call(fq_add(p.z, q.z)), z1z1), z2z2), h);
  return r;
}

// ---- FqExp row (reference src/fields/fq/exp.rs:128-178, src/fields/fq/mul.rs:34-45): a16 b16 | output16 aux95 sign ----
#define EXP_OP_NONE 0
#define EXP_OP_MUL 1     // output = a * b   (G1/G2: add)
#define EXP_OP_SQUARE 2  // output = a * a   (G1/G2: double)
template <class W> HD void fq_exp_row(const u32* a, const u32* b, int op, W& w) {
  write_limbs16(w, 0, a); write_limbs16(w, 16, b);
  const int o = 32;
  if (op == EXP_OP_NONE) {
    for (int i = 0; i < 111; i++) w(o + i, 0);
    w(o + 111, 1);
    return;
  }
  const u32* y = op == EXP_OP_SQUARE ? a : b;
  Fq yraw;
#pragma unroll
  for (int i = 0; i < 8; i++) yraw.l[i] = y[i];
  Fq out = fq_mul(fq_from_words(a), yraw);   // (aR) * y / R = a * y, canonical
  u32 ow[8];
#pragma unroll
  for (int i = 0; i < 8; i++) ow[i] = out.l[i];
  i64 al[16], yl[16], pol[31];
  fq_words_to_limbs(a, al); fq_words_to_limbs(y, yl);
  for (int i = 0; i < 31; i++) pol[i] = 0;
  pol_mul_acc(pol, al, yl, 1);
  ModWitness m;
  modular_witness(pol, ow, m);
  write_limbs16(w, o, ow);
  write_modulus_aux(w, o + 16, m);
  w(o + 111, sign_to_gl(m.negative));
}

// ---- Fq2 = Fq[u]/(u^2 + 1), Montgomery coordinates ----
struct Fq2 { Fq c0, c1; };
HD Fq2 fq2_add(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = fq_add(a.c0, b.c0); r.c1 = fq_add(a.c1, b.c1); return r; }
HD Fq2 fq2_sub(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = fq_sub(a.c0, b.c0); r.c1 = fq_sub(a.c1, b.c1); return r; }
HD Fq2 fq2_dbl(const Fq2& a) { return fq2_add(a, a); }
HD Fq2 fq2_mul(const Fq2& a, const Fq2& b) {   // Karatsuba: 3 base multiplications
  Fq t0 = fq_mul_call(a.c0, b.c0), t1 = fq_mul_call(a.c1, b.c1);
  Fq s = fq_mul_call(fq_add(a.c0, a.c1), fq_add(b.c0, b.c1));
  Fq2 r; r.c0 = fq_sub(t0, t1); r.c1 = fq_sub(fq_sub(s, t0), t1); return r;
}
HD Fq2 fq2_sqr(const Fq2& a) {   // (c0 + c1)(c0 - c1), 2 c0 c1
  Fq2 r; r.c0 = fq_mul_call(fq_add(a.c0, a.c1), fq_sub(a.c0, a.c1)); r.c1 = fq_dbl(fq_mul_call(a.c0, a.c1)); return r;
}
HD bool fq2_is_zero(const Fq2& a) { return fq_is_zero(a.c0) && fq_is_zero(a.c1); }
HD Fq2 fq2_inv(const Fq2& a) {   // conj(a) / (c0^2 + c1^2)
  Fq n = fq_inv(fq_add(fq_sqr(a.c0), fq_sqr(a.c1)));
  Fq2 r; r.c0 = fq_mul(a.c0, n); r.c1 = fq_sub(fq_zero(), fq_mul(a.c1, n)); return r;
}
HD Fq2 fq2_one() { Fq2 r; r.c0 = fq_one(); r.c1 = fq_zero(); return r; }
HD Fq2 fq2_from_words(const u32* w /*16*/) { Fq2 r; r.c0 = fq_from_words(w); r.c1 = fq_from_words(w + 8); return r; }
HD void fq2_to_words(const Fq2& a, u32* w /*16*/) { fq_to_words(a.c0, w); fq_to_words(a.c1, w + 8); }

// Jacobian doubling / addition on y^2 = x^3 + b (a = 0) over Fq2 (same formulas as g1_jac_dbl / g1_jac_add)
struct G2Jac { Fq2 x, y, z; };
HD G2Jac g2_jac_dbl(const G2Jac& p) {
  Fq2 A = fq2_sqr(p.x), B = fq2_sqr(p.y), C = fq2_sqr(B);
  Fq2 t = fq2_add(p.x, B);
  Fq2 D = fq2_dbl(fq2_sub(fq2_sub(fq2_sqr(t), A), C));
  Fq2 E = fq2_add(fq2_dbl(A), A), Fv = fq2_sqr(E);
  G2Jac r;
  r.x = fq2_sub(Fv, fq2_dbl(D));
  Fq2 c8 = fq2_dbl(fq2_dbl(fq2_dbl(C)));
  r.y = fq2_sub(fq2_mul(E, fq2_sub(D, r.x)), c8);
  r.z = fq2_dbl(fq2_mul(p.y, p.z));
  return r;
}
HD G2Jac g2_jac_add(const G2Jac& p, const G2Jac& q) {
  Fq2 z1z1 = fq2_sqr(p.z), z2z2 = fq2_sqr(q.z);
  Fq2 u1 = fq2_mul(p.x, z2z2), u2 = fq2_mul(q.x, z1z1);
  Fq2 s1 = fq2_mul(fq2_mul(p.y, q.z), z2z2), s2 = fq2_mul(fq2_mul(q.y, p.z), z1z1);
  Fq2 h = fq2_sub(u2, u1);
  Fq2 i = fq2_sqr(fq2_dbl(h)), j = fq2_mul(h, i);
  Fq2 rr = fq2_dbl(fq2_sub(s2, s1)), v = fq2_mul(u1, i);
  G2Jac r;
  r.x = fq2_sub(fq2_sub(fq2_sqr(rr), j), fq2_dbl(v));
  r.y = fq2_sub(fq2_mul(rr, fq2_sub(v, r.x)), fq2_dbl(fq2_mul(s1, j)));
  r.z = fq2_mul(fq2_sub(fq2_sub(fq2_sqr(fq2_add(p.z, q.z)), z1z1), z2z2), h);
  return r;
}

// res(31) += sign * (x0 + x1 u)(y0 + y1 u) component `c` in limb-polynomial form (reference src/fields/fq2.rs:41-58),
// scaled:  c = 0: x0 y0 - x1 y1;  c = 1: x0 y1 + x1 y0
HD void pol_mul_fq2_acc(i64* res, const i64* x0, const i64* x1, const i64* y0, const i64* y1, int c, i64 scale) {
  if (c == 0) { pol_mul_acc(res, x0, y0, scale); pol_mul_acc(res, x1, y1, -scale); }
  else { pol_mul_acc(res, x0, y1, scale); pol_mul_acc(res, x1, y0, scale); }
}
// ---- G2 row (reference src/curves/g2/muladd.rs:118-201 `generate_g2_double`, :330-414 `generate_g2_add`, :42-54 default;
//      row layout src/curves/g2/exp.rs:180-245): a.x a.y b.x b.y (4 x 32) | G2Output(640) ----
// ax..by: canonical words, c0 then c1 (16 words each).  Returns false if the slope denominator is zero.
template <class W> HD bool g2_row(const u32* ax, const u32* ay, const u32* bx, const u32* by, int op, W& w) {
  write_limbs16(w, 0, ax); write_limbs16(w, 16, ax + 8); write_limbs16(w, 32, ay); write_limbs16(w, 48, ay + 8);
  write_limbs16(w, 64, bx); write_limbs16(w, 80, bx + 8); write_limbs16(w, 96, by); write_limbs16(w, 112, by + 8);
  const int o = 128;
  if (op == EXP_OP_NONE) {
    for (int i = 0; i < 634; i++) w(o + i, 0);
    for (int i = 634; i < 640; i++) w(o + i, 1);
    return true;
  }
  Fq2 x1 = fq2_from_words(ax), y1 = fq2_from_words(ay), x2, lambda;
  i64 x1l[2][16], y1l[2][16], x2l[2][16], ll[2][16], t1[2][16], pol[31];
  for (int c = 0; c < 2; c++) { fq_words_to_limbs(ax + 8 * c, x1l[c]); fq_words_to_limbs(ay + 8 * c, y1l[c]); }
  u32 lw[16];
  ModWitness m;
  if (op == EXP_OP_MUL) {   // add
    x2 = fq2_from_words(bx);
    Fq2 y2 = fq2_from_words(by);
    Fq2 dx = fq2_sub(x2, x1);
    if (fq2_is_zero(dx)) return false;
    lambda = fq2_mul(fq2_sub(y2, y1), fq2_inv(dx));
    fq2_to_words(lambda, lw);
    i64 y2l[2][16];
    for (int c = 0; c < 2; c++) { fq_words_to_limbs(lw + 8 * c, ll[c]); fq_words_to_limbs(bx + 8 * c, x2l[c]); fq_words_to_limbs(by + 8 * c, y2l[c]); }
    for (int c = 0; c < 2; c++) for (int i = 0; i < 16; i++) t1[c][i] = x2l[c][i] - x1l[c][i];
    // zero_pol = lambda * (x2 - x1) - (y2 - y1)
    for (int c = 0; c < 2; c++) {
      for (int i = 0; i < 31; i++) pol[i] = 0;
      pol_mul_fq2_acc(pol, ll[0], ll[1], t1[0], t1[1], c, 1);
      for (int i = 0; i < 16; i++) pol[i] -= y2l[c][i] - y1l[c][i];
      modular_witness(pol, nullptr, m);
      write_modulus_aux_zero(w, o + 96 + 79 * c, m);
      w(o + 634 + c, sign_to_gl(m.negative));
    }
  } else {   // double
    x2 = x1;
    Fq2 den = fq2_dbl(y1);
    if (fq2_is_zero(den)) return false;
    Fq2 x1sq = fq2_sqr(x1);
    lambda = fq2_mul(fq2_add(fq2_dbl(x1sq), x1sq), fq2_inv(den));
    fq2_to_words(lambda, lw);
    for (int c = 0; c < 2; c++) { fq_words_to_limbs(lw + 8 * c, ll[c]); for (int i = 0; i < 16; i++) x2l[c][i] = x1l[c][i]; }
    // zero_pol = 2 * lambda * y - 3 * x * x
    for (int c = 0; c < 2; c++) {
      for (int i = 0; i < 31; i++) pol[i] = 0;
      pol_mul_fq2_acc(pol, ll[0], ll[1], y1l[0], y1l[1], c, 2);
      pol_mul_fq2_acc(pol, x1l[0], x1l[1], x1l[0], x1l[1], c, -3);
      modular_witness(pol, nullptr, m);
      write_modulus_aux_zero(w, o + 96 + 79 * c, m);
      w(o + 634 + c, sign_to_gl(m.negative));
    }
  }
  write_limbs16(w, o, lw); write_limbs16(w, o + 16, lw + 8);
  // new_x = lambda^2 - x1 - x2
  Fq2 nx = fq2_sub(fq2_sub(fq2_sqr(lambda), x1), x2);
  u32 nxw[16]; fq2_to_words(nx, nxw);
  for (int c = 0; c < 2; c++) {
    for (int i = 0; i < 31; i++) pol[i] = 0;
    pol_mul_fq2_acc(pol, ll[0], ll[1], ll[0], ll[1], c, 1);
    for (int i = 0; i < 16; i++) pol[i] -= x1l[c][i] + x2l[c][i];
    modular_witness(pol, nxw + 8 * c, m);
    write_limbs16(w, o + 32 + 16 * c, nxw + 8 * c);
    write_modulus_aux(w, o + 254 + 95 * c, m);
    w(o + 636 + c, sign_to_gl(m.negative));
  }
  // new_y = lambda * (x1 - new_x) - y1
  Fq2 ny = fq2_sub(fq2_mul(lambda, fq2_sub(x1, nx)), y1);
  u32 nyw[16]; fq2_to_words(ny, nyw);
  for (int c = 0; c < 2; c++) { i64 nxl[16]; fq_words_to_limbs(nxw + 8 * c, nxl); for (int i = 0; i < 16; i++) t1[c][i] = x1l[c][i] - nxl[i]; }
  for (int c = 0; c < 2; c++) {
    for (int i = 0; i < 31; i++) pol[i] = 0;
    pol_mul_fq2_acc(pol, ll[0], ll[1], t1[0], t1[1], c, 1);
    for (int i = 0; i < 16; i++) pol[i] -= y1l[c][i];
    modular_witness(pol, nyw + 8 * c, m);
    write_limbs16(w, o + 64 + 16 * c, nyw + 8 * c);
    write_modulus_aux(w, o + 254 + 95 * (2 + c), m);
    w(o + 636 + 2 + c, sign_to_gl(m.negative));
  }
  return true;
}

// ---- Fq12 in the flat MyFq12 basis: element = sum_{i<6} (c[i] + c[i+6] u) w^i, u^2 = -1, w^6 = 9 + u ----
// One output coefficient of x * y in Fq (Montgomery) -- the field-level image of `pol_mul_fq12` (reference
// src/fields/fq12/mul.rs:24-87), used by the exponentiation chain.
HD Fq fq12_mul_coeff(const Fq* x, const Fq* y, int oi) {
  const bool imag = oi >= 6; const int i0 = imag ? oi - 6 : oi;
  Fq acc = fq_zero();
  for (int term = 0; term < 3; term++) {
    int m, part, wgt;
    if (term == 0) { m = i0; part = imag ? 1 : 0; wgt = 1; }
    else if (i0 == 5) break;
    else if (term == 1) { m = i0 + 6; part = 0; wgt = imag ? 1 : 9; }
    else { m = i0 + 6; part = 1; wgt = imag ? 9 : -1; }
    Fq sum = fq_zero();
    for (int i = 0; i < 6; i++) {
      const int j = m - i;
      if (j < 0 || j > 5) continue;
      if (part == 0) sum = fq_sub(fq_add(sum, fq_mul_call(x[i], y[j])), fq_mul_call(x[i + 6], y[j + 6]));
      else sum = fq_add(fq_add(sum, fq_mul_call(x[i], y[j + 6])), fq_mul_call(x[i + 6], y[j]));
    }
    if (wgt == 9) { Fq s2 = fq_dbl(sum), s4 = fq_dbl(s2), s8 = fq_dbl(s4); sum = fq_add(s8, sum); }
    acc = wgt < 0 ? fq_sub(acc, sum) : fq_add(acc, sum);
  }
  return acc;
}
// Same coefficient from the 144 pairwise products prod[12 * i + j] = x_i * y_j (the chain kernel computes those with one
// thread each and sums them here).
HD Fq fq12_sum_coeff(const Fq* prod, int oi) {
  const bool imag = oi >= 6; const int i0 = imag ? oi - 6 : oi;
  Fq acc = fq_zero();
  for (int term = 0; term < 3; term++) {
    int m, part, wgt;
    if (term == 0) { m = i0; part = imag ? 1 : 0; wgt = 1; }
    else if (i0 == 5) break;
    else if (term == 1) { m = i0 + 6; part = 0; wgt = imag ? 1 : 9; }
    else { m = i0 + 6; part = 1; wgt = imag ? 9 : -1; }
    Fq sum = fq_zero();
    for (int i = 0; i < 6; i++) {
      const int j = m - i;
      if (j < 0 || j > 5) continue;
      if (part == 0) sum = fq_sub(fq_add(sum, prod[12 * i + j]), prod[12 * (i + 6) + j + 6]);
      else sum = fq_add(fq_add(sum, prod[12 * i + j + 6]), prod[12 * (i + 6) + j]);
    }
    if (wgt == 9) { Fq s2 = fq_dbl(sum), s4 = fq_dbl(s2), s8 = fq_dbl(s4); sum = fq_add(s8, sum); }
    acc = wgt < 0 ? fq_sub(acc, sum) : fq_add(acc, sum);
  }
  return acc;
}
// i64 limb polynomial of output coefficient `oi` of pol_mul_fq12(x, y, 9); xl / yl: 12 x 16 limbs (u16 values)
HD void fq12_pol_input(const unsigned short* xl, const unsigned short* yl, int oi, i64* pol /*31*/) {
  for (int k = 0; k < 31; k++) pol[k] = 0;
  const bool imag = oi >= 6; const int i0 = imag ? oi - 6 : oi;
  for (int term = 0; term < 3; term++) {
    int m, part; i64 wgt;
    if (term == 0) { m = i0; part = imag ? 1 : 0; wgt = 1; }
    else if (i0 == 5) break;
    else if (term == 1) { m = i0 + 6; part = 0; wgt = imag ? 1 : 9; }
    else { m = i0 + 6; part = 1; wgt = imag ? 9 : -1; }
    for (int i = 0; i < 6; i++) {
      const int j = m - i;
      if (j < 0 || j > 5) continue;
      for (int half = 0; half < 2; half++) {
        const int xi = half ? i + 6 : i;
        const int yj = part == 0 ? (half ? j + 6 : j) : (half ? j : j + 6);
        const i64 sc = (part == 0 && half == 1) ? -wgt : wgt;
        const unsigned short* xp = xl + 16 * xi; const unsigned short* yp = yl + 16 * yj;
        for (int s = 0; s < 16; s++) { const i64 xs = sc * (i64)xp[s]; for (int t = 0; t < 16; t++) pol[s + t] += xs * (i64)yp[t]; }
      }
    }
  }
}
// One coefficient of an Fq12 row (reference src/fields/fq12/mul.rs:192-231, src/fields/fq12/exp.rs:142-214): writes
// output[oi] (16), aux[oi] (95) and sign[oi] of the Fq12Output block at column 384.  `out_words`: canonical words of
// coefficient oi of x * y (known from the chain), op NONE writes the default block.
template <class W> HD void fq12_row_coeff(const unsigned short* xl, const unsigned short* yl, const u32* out_words, int oi, int op, W& w) {
  const int o = 384;
  if (op == EXP_OP_NONE) {
    for (int i = 0; i < 16; i++) w(o + 16 * oi + i, 0);
    for (int i = 0; i < 95; i++) w(o + 192 + 95 * oi + i, 0);
    w(o + 192 + 12 * 95 + oi, 1);
    return;
  }
  i64 pol[31];
  fq12_pol_input(xl, yl, oi, pol);
  ModWitness m;
  modular_witness(pol, out_words, m);
  write_limbs16(w, o + 16 * oi, out_words);
  write_modulus_aux(w, o + 192 + 95 * oi, m);
  w(o + 192 + 12 * 95 + oi, sign_to_gl(m.negative));
}
// flag columns of the 64-bit exponent variant in closed form (reference src/fields/fq12_u64/flags_u64.rs:34-94):
// is_final, a, b, filtered_bit, bit, val; row r of a 128-row block handles bit j = r / 2.
HD void flags_u64_row(u64 e, int r, u64* out /*6*/) {
  const int j = r >> 1;
  const u64 bit = (e >> j) & 1, a = r & 1, b = 1 - a;
  out[0] = (r == 127) ? 1 : 0; out[1] = a; out[2] = b; out[3] = bit * b; out[4] = bit;
  out[5] = j == 63 ? 0 : (e >> (j + 1));
}
