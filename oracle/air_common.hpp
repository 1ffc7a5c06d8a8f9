// ORACLE (test infrastructure).  Restatement of the reference's arithmetic gadgets, written once as
// templates over the value type so the same code serves witness generation (T = i64), the prover's
// base-field constraint evaluation (P = GF) and the verifier's extension-field evaluation (P = GF2),
// exactly as the reference instantiates `P: PackedField`.  Each function cites the file:line it follows.
#pragma once
#include "gl.hpp"
#include "bn254.hpp"
#include <array>
#include <vector>
#include <algorithm>

namespace orc {

// starky::constraint_consumer::ConstraintConsumer (dependency): acc = acc*alpha + c.
template <class P> struct Consumer {
  std::vector<P> alphas, accs;
  P z_last, l_first, l_last;
  size_t count = 0;
  std::vector<P>* log = nullptr;  // debug: every constraint value in emission order
  Consumer(const std::vector<P>& al, P zl, P lf, P ll) : alphas(al), accs(al.size()), z_last(zl), l_first(lf), l_last(ll) {}
  void constraint(P c) { for (size_t k = 0; k < accs.size(); k++) accs[k] = accs[k] * alphas[k] + c; count++; if (log) log->push_back(c); }
  void transition(P c) { constraint(c * z_last); }
  void first_row(P c) { constraint(c * l_first); }
  void last_row(P c) { constraint(c * l_last); }
};

template <class T> static inline T tzero() { return T(); }
template <> inline i64 tzero<i64>() { return 0; }

// ---- reference src/modular/pol_utils.rs ----
template <class T, size_t N> using Arr = std::array<T, N>;
template <class T, size_t N> static inline Arr<T, N> pol_zero() { Arr<T, N> a; for (auto& x : a) x = tzero<T>(); return a; }
// pol_utils.rs:25-33 / :195-203 (a.len() >= b.len())
template <class T, size_t N, size_t M> static inline void pol_add_assign(Arr<T, N>& a, const Arr<T, M>& b) { static_assert(N >= M, ""); for (size_t i = 0; i < M; i++) a[i] = a[i] + b[i]; }
template <class T, size_t N, size_t M> static inline void pol_sub_assign(Arr<T, N>& a, const Arr<T, M>& b) { static_assert(N >= M, ""); for (size_t i = 0; i < M; i++) a[i] = a[i] - b[i]; }
// pol_utils.rs:48-57 / :122-131 (results zero-extended to 2N-1)
template <class T> static inline Arr<T, 31> pol_add(const Arr<T, 16>& a, const Arr<T, 16>& b) { auto r = pol_zero<T, 31>(); for (int i = 0; i < 16; i++) r[i] = a[i] + b[i]; return r; }
template <class T> static inline Arr<T, 31> pol_sub(const Arr<T, 16>& a, const Arr<T, 16>& b) { auto r = pol_zero<T, 31>(); for (int i = 0; i < 16; i++) r[i] = a[i] - b[i]; return r; }
// pol_utils.rs:59-68 / :133-142
template <class T, size_t N> static inline Arr<T, N> pol_add_normal(const Arr<T, N>& a, const Arr<T, N>& b) { Arr<T, N> r; for (size_t i = 0; i < N; i++) r[i] = a[i] + b[i]; return r; }
template <class T, size_t N> static inline Arr<T, N> pol_sub_normal(const Arr<T, N>& a, const Arr<T, N>& b) { Arr<T, N> r; for (size_t i = 0; i < N; i++) r[i] = a[i] - b[i]; return r; }
// pol_utils.rs:221-232
template <class T> static inline Arr<T, 31> pol_mul_wide(const Arr<T, 16>& a, const Arr<T, 16>& b) {
  auto r = pol_zero<T, 31>();
  for (int i = 0; i < 16; i++) for (int j = 0; j < 16; j++) r[i + j] = r[i + j] + a[i] * b[j];
  return r;
}
// pol_utils.rs:234-243
template <class T, size_t N> static inline Arr<T, N> pol_mul_scalar(const Arr<T, N>& a, T c) { Arr<T, N> r; for (size_t i = 0; i < N; i++) r[i] = c * a[i]; return r; }
// pol_utils.rs:274-285
template <class T> static inline Arr<T, 32> pol_mul_wide2(const Arr<T, 17>& a, const Arr<T, 16>& b) {
  auto r = pol_zero<T, 32>();
  for (int i = 0; i < 17; i++) for (int j = 0; j < 16; j++) r[i + j] = r[i + j] + a[i] * b[j];
  return r;
}
// pol_utils.rs:348-363: (x - root) * a(x)
template <class T, size_t N> static inline Arr<T, N> pol_adjoin_root(const Arr<T, N>& a, T root) {
  Arr<T, N> r;
  r[0] = tzero<T>() - root * a[0];
  for (size_t d = 1; d < N; d++) r[d] = a[d - 1] - root * a[d];
  return r;
}
// pol_utils.rs:390-414 (i64 only; last element deliberately left zero)
template <size_t N> static inline Arr<i64, N> pol_remove_root_2exp16(const Arr<i64, N>& a) {
  auto q = pol_zero<i64, N>();
  q[0] = -(a[0] >> 16);
  for (size_t d = 1; d < N - 1; d++) q[d] = (q[d - 1] - a[d]) >> 16;
  return q;
}
template <class T> static inline Arr<T, 31> widen16(const Arr<T, 16>& a) { auto r = pol_zero<T, 31>(); for (int i = 0; i < 16; i++) r[i] = a[i]; return r; }

// ---- reference src/modular/modular.rs:31-36, modular_zero.rs:27-31 ----
template <class P> struct ModulusAux { Arr<P, 16> out_aux_red; Arr<P, 17> quot_abs; Arr<P, 31> lo, hi; };
template <class P> struct ModulusAuxZero { Arr<P, 17> quot_abs; Arr<P, 31> lo, hi; };
static const i64 AUX_COEFF_ABS_MAX = 1LL << 29;  // modular.rs:28

static inline Arr<i64, 16> modulus_limbs_i64() { Arr<i64, 16> m; for (int i = 0; i < 16; i++) m[i] = BN254_P_LIMBS[i]; return m; }
template <class P> static inline Arr<P, 16> bn254_base_modulus_packfield() {  // modular.rs:305-309
  Arr<P, 16> m; for (int i = 0; i < 16; i++) m[i] = FieldOf<P>::c((u64)BN254_P_LIMBS[i]); return m;
}

template <size_t N> static inline Big cols_to_big(const Arr<i64, N>& a) { i64 t[N]; for (size_t i = 0; i < N; i++) t[i] = a[i]; return columns_to_bigint<(int)N>(t); }
template <size_t N> static inline Arr<i64, N> big_to_cols(const Big& b) { i64 t[N]; bigint_to_columns<(int)N>(b, t); Arr<i64, N> r; for (size_t i = 0; i < N; i++) r[i] = t[i]; return r; }
static inline Big big_div_exact_or_floor(const Big& num, const Big& den, Big* rem) {
  // truncated division like num_bigint (`/` and `%` truncate toward zero)
  Big q, r;
  int nn = num.top(), dn = den.top();
  assert(dn > 0);
  if (nn < dn) { r = num; if (rem) *rem = r; return q; }
  mag_divmod(num.m, nn, den.m, dn, q.m, r.m);
  q.neg = !q.is_zero() && (num.neg != den.neg);
  r.neg = !r.is_zero() && num.neg;
  if (rem) *rem = r;
  return q;
}

struct ModOpWitness { Arr<GF, 16> output; GF quot_sign; ModulusAux<GF> aux; };
struct ModZeroWitness { GF quot_sign; ModulusAuxZero<GF> aux; };

static inline void split_aux(const Arr<i64, 32>& aux_limbs_in, Arr<GF, 31>& lo, Arr<GF, 31>& hi) {
  Arr<i64, 32> aux_limbs = aux_limbs_in;
  assert(aux_limbs[31] == 0);
  for (auto& c : aux_limbs) c += AUX_COEFF_ABS_MAX;
  for (auto& c : aux_limbs) { assert((c < 0 ? -c : c) <= 2 * AUX_COEFF_ABS_MAX); (void)c; }
  for (int i = 0; i < 31; i++) {
    lo[i] = GF((u64)(uint16_t)aux_limbs[i]);
    hi[i] = GF((u64)(uint16_t)(aux_limbs[i] >> 16));
  }
}
// modular.rs:38-100 `generate_modular_op`
static inline ModOpWitness generate_modular_op(const Arr<i64, 31>& pol_input) {
  Big modulus = bn254_modulus_big();
  Arr<i64, 16> modulus_limbs = modulus_limbs_i64();
  auto constr_poly = pol_zero<i64, 32>();
  for (int i = 0; i < 31; i++) constr_poly[i] = pol_input[i];
  Big input = cols_to_big(constr_poly);
  Big output;
  big_div_exact_or_floor(input, modulus, &output);       // output = input % modulus (truncated)
  if (output.neg) output = big_add(output, modulus);     // modular.rs:50-52
  Arr<i64, 16> output_limbs = big_to_cols<16>(output);
  Big rem;
  Big quot = big_div_exact_or_floor(big_sub(input, output), modulus, &rem);
  assert(rem.is_zero());
  ModOpWitness w;
  w.quot_sign = quot.neg ? -GF::one() : GF::one();       // modular.rs:56-60
  Arr<i64, 17> quot_limbs = big_to_cols<17>(quot);
  Big qa = quot; qa.neg = false;
  Arr<i64, 17> quot_abs_limbs = big_to_cols<17>(qa);
  Big two256; two256.m[8] = 1;
  Arr<i64, 16> out_aux_red = big_to_cols<16>(big_add(big_sub(two256, modulus), output));  // modular.rs:66
  pol_sub_assign(constr_poly, output_limbs);
  Arr<i64, 32> prod = pol_mul_wide2(quot_limbs, modulus_limbs);
  pol_sub_assign(constr_poly, prod);
  Arr<i64, 32> aux_limbs = pol_remove_root_2exp16(constr_poly);   // modular.rs:74
  split_aux(aux_limbs, w.aux.lo, w.aux.hi);
  for (int i = 0; i < 16; i++) { w.output[i] = GF((u64)output_limbs[i]); w.aux.out_aux_red[i] = GF::from_i64(out_aux_red[i]); }
  for (int i = 0; i < 17; i++) w.aux.quot_abs[i] = GF::from_i64(quot_abs_limbs[i]);
  return w;
}
// modular_zero.rs:33-80 `generate_modular_zero`
static inline ModZeroWitness generate_modular_zero(const Arr<i64, 31>& zero_pol) {
  Big modulus = bn254_modulus_big();
  Arr<i64, 16> modulus_limbs = modulus_limbs_i64();
  Big input = cols_to_big(zero_pol);
  Big rem;
  Big quot = big_div_exact_or_floor(input, modulus, &rem);
  assert(rem.is_zero());                                  // modular_zero.rs:39
  ModZeroWitness w;
  w.quot_sign = quot.neg ? -GF::one() : GF::one();
  Arr<i64, 17> quot_limbs = big_to_cols<17>(quot);
  Big qa = quot; qa.neg = false;
  Arr<i64, 17> quot_abs_limbs = big_to_cols<17>(qa);
  auto constr_poly = pol_zero<i64, 32>();
  for (int i = 0; i < 31; i++) constr_poly[i] = zero_pol[i];
  Arr<i64, 32> prod = pol_mul_wide2(quot_limbs, modulus_limbs);
  pol_sub_assign(constr_poly, prod);
  Arr<i64, 32> aux_limbs = pol_remove_root_2exp16(constr_poly);
  split_aux(aux_limbs, w.aux.lo, w.aux.hi);
  for (int i = 0; i < 17; i++) w.aux.quot_abs[i] = GF::from_i64(quot_abs_limbs[i]);
  return w;
}

// ---- reference src/modular/addcy.rs:16-58 ----
static const u64 GOLDILOCKS_INVERSE_65536 = 18446462594437939201ULL;
template <class P> static inline void eval_addcy(Consumer<P>& yc, P filter, const Arr<P, 16>& x, const Arr<P, 16>& y,
                                                 const Arr<P, 16>& z, const Arr<P, 16>& given_cy) {
  P overflow = FieldOf<P>::c(1ULL << 16), overflow_inv = FieldOf<P>::c(GOLDILOCKS_INVERSE_65536);
  P cy = tzero<P>();
  for (int i = 0; i < 16; i++) {
    P t = cy + x[i] + y[i] - z[i];
    yc.constraint(filter * t * (overflow - t));
    cy = t * overflow_inv;
  }
  yc.constraint(filter * given_cy[0] * (given_cy[0] - FieldOf<P>::c(1)));
  yc.constraint(filter * (cy - given_cy[0]));
  for (int i = 1; i < 16; i++) yc.constraint(filter * given_cy[i]);
}
// modular.rs:102-153 `modular_constr_poly`
template <class P> static inline Arr<P, 32> modular_constr_poly(Consumer<P>& yc, P filter, const Arr<P, 16>& modulus,
                                                                const Arr<P, 16>& output, P quot_sign, const ModulusAux<P>& aux) {
  auto is_less_than = pol_zero<P, 16>();
  is_less_than[0] = FieldOf<P>::c(1);
  eval_addcy(yc, filter, modulus, aux.out_aux_red, output, is_less_than);
  yc.constraint(filter * (quot_sign * quot_sign - FieldOf<P>::c(1)));
  Arr<P, 17> quot; for (int i = 0; i < 17; i++) quot[i] = quot_sign * aux.quot_abs[i];
  Arr<P, 32> constr_poly = pol_mul_wide2(quot, modulus);
  pol_add_assign(constr_poly, output);
  P base = FieldOf<P>::c(1ULL << 16), offset = FieldOf<P>::c((u64)AUX_COEFF_ABS_MAX);
  auto aux_poly = pol_zero<P, 32>();
  for (int i = 0; i < 31; i++) { aux_poly[i] = aux.lo[i] - offset; aux_poly[i] = aux_poly[i] + base * aux.hi[i]; }
  pol_add_assign(constr_poly, pol_adjoin_root(aux_poly, base));
  return constr_poly;
}
// modular.rs:215-230 `eval_modular_op`
template <class P> static inline void eval_modular_op(Consumer<P>& yc, P filter, const Arr<P, 16>& modulus, const Arr<P, 31>& input,
                                                      const Arr<P, 16>& output, P quot_sign, const ModulusAux<P>& aux) {
  Arr<P, 32> c = modular_constr_poly(yc, filter, modulus, output, quot_sign, aux);
  pol_sub_assign(c, input);
  for (auto& v : c) yc.constraint(filter * v);
}
// modular_zero.rs:82-120 `eval_modular_zero`
template <class P> static inline void eval_modular_zero(Consumer<P>& yc, P filter, const Arr<P, 16>& modulus, const Arr<P, 31>& input,
                                                        P quot_sign, const ModulusAuxZero<P>& aux) {
  yc.constraint(filter * (quot_sign * quot_sign - FieldOf<P>::c(1)));
  Arr<P, 17> quot; for (int i = 0; i < 17; i++) quot[i] = quot_sign * aux.quot_abs[i];
  Arr<P, 32> constr_poly = pol_mul_wide2(quot, modulus);
  P base = FieldOf<P>::c(1ULL << 16), offset = FieldOf<P>::c((u64)AUX_COEFF_ABS_MAX);
  auto aux_poly = pol_zero<P, 32>();
  for (int i = 0; i < 31; i++) { aux_poly[i] = aux.lo[i] - offset; aux_poly[i] = aux_poly[i] + base * aux.hi[i]; }
  pol_add_assign(constr_poly, pol_adjoin_root(aux_poly, base));
  pol_sub_assign(constr_poly, input);
  for (auto& v : constr_poly) yc.constraint(filter * v);
}

// ---- column (de)serialisers: modular.rs:260-296, modular_zero.rs:174-197 ----
template <class T> static inline void write_u256(T* lv, const Arr<T, 16>& v, size_t& cur) { for (int i = 0; i < 16; i++) lv[cur + i] = v[i]; cur += 16; }
template <class T> static inline Arr<T, 16> read_u256(const T* lv, size_t& cur) { Arr<T, 16> r; for (int i = 0; i < 16; i++) r[i] = lv[cur + i]; cur += 16; return r; }
template <class T> static inline void write_modulus_aux(T* lv, const ModulusAux<T>& a, size_t& cur) {
  for (int i = 0; i < 16; i++) lv[cur + i] = a.out_aux_red[i];
  for (int i = 0; i < 17; i++) lv[cur + 16 + i] = a.quot_abs[i];
  for (int i = 0; i < 31; i++) lv[cur + 33 + i] = a.lo[i];
  for (int i = 0; i < 31; i++) lv[cur + 64 + i] = a.hi[i];
  cur += 95;
}
template <class T> static inline ModulusAux<T> read_modulus_aux(const T* lv, size_t& cur) {
  ModulusAux<T> a;
  for (int i = 0; i < 16; i++) a.out_aux_red[i] = lv[cur + i];
  for (int i = 0; i < 17; i++) a.quot_abs[i] = lv[cur + 16 + i];
  for (int i = 0; i < 31; i++) a.lo[i] = lv[cur + 33 + i];
  for (int i = 0; i < 31; i++) a.hi[i] = lv[cur + 64 + i];
  cur += 95; return a;
}
template <class T> static inline void write_modulus_aux_zero(T* lv, const ModulusAuxZero<T>& a, size_t& cur) {
  for (int i = 0; i < 17; i++) lv[cur + i] = a.quot_abs[i];
  for (int i = 0; i < 31; i++) lv[cur + 17 + i] = a.lo[i];
  for (int i = 0; i < 31; i++) lv[cur + 48 + i] = a.hi[i];
  cur += 79;
}
template <class T> static inline ModulusAuxZero<T> read_modulus_aux_zero(const T* lv, size_t& cur) {
  ModulusAuxZero<T> a;
  for (int i = 0; i < 17; i++) a.quot_abs[i] = lv[cur + i];
  for (int i = 0; i < 31; i++) a.lo[i] = lv[cur + 17 + i];
  for (int i = 0; i < 31; i++) a.hi[i] = lv[cur + 48 + i];
  cur += 79; return a;
}
static inline Arr<i64, 16> positive_column_to_i64(const Arr<GF, 16>& c) { Arr<i64, 16> r; for (int i = 0; i < 16; i++) r[i] = (i64)c[i].v; return r; }
static inline Arr<GF, 16> i64_to_column_positive(const Arr<i64, 16>& c) { Arr<GF, 16> r; for (int i = 0; i < 16; i++) r[i] = GF((u64)c[i]); return r; }
static inline Arr<i64, 16> fq_to_cols(const Fq& x) { i64 t[16]; fq_to_columns(x, t); Arr<i64, 16> r; for (int i = 0; i < 16; i++) r[i] = t[i]; return r; }
static inline Fq cols_to_fq(const Arr<GF, 16>& c) { i64 t[16]; for (int i = 0; i < 16; i++) t[i] = (i64)c[i].v; return columns_to_fq(t); }
static inline Fq cols_to_fq(const Arr<i64, 16>& c) { i64 t[16]; for (int i = 0; i < 16; i++) t[i] = c[i]; return columns_to_fq(t); }

// utils.rs:56-63 `u16_columns_to_u32_columns`
template <class P> static inline Arr<P, 8> u16_columns_to_u32_columns(const Arr<P, 16>& x) {
  P base = FieldOf<P>::c(1ULL << 16); Arr<P, 8> r;
  for (int i = 0; i < 8; i++) r[i] = x[2 * i] + base * x[2 * i + 1];
  return r;
}
// utils.rs:24-34 `fq_to_u32_columns`
static inline Arr<GF, 8> u256_to_u32_columns(const U256& v) { Arr<GF, 8> r; for (int i = 0; i < 8; i++) r[i] = GF((v.w[i / 2] >> (32 * (i & 1))) & 0xFFFFFFFFULL); return r; }

// ---- reference src/utils/equals.rs ----
template <class P, size_t N> static inline void vec_equal(Consumer<P>& yc, P filter, const Arr<P, N>& x, const Arr<P, N>& y) { for (size_t i = 0; i < N; i++) yc.constraint(filter * (x[i] - y[i])); }
template <class P> static inline void fq_equal_transition(Consumer<P>& yc, P filter, const Arr<P, 16>& x, const Arr<P, 16>& y) { for (int i = 0; i < 16; i++) yc.transition(filter * (x[i] - y[i])); }

// ---- reference src/utils/flags.rs ----
static const int NUM_INPUT_LIMBS = 8, INPUT_LIMB_BITS = 32, NUM_FLAGS_COLS = 14;
// flags.rs:46-75
static inline void generate_flags_first_row(GF* lv, size_t sf, const u32 limbs[8]) {
  u32 first_bit = limbs[0] % 2, rest = (limbs[0] - first_bit) / 2;
  lv[sf] = GF(); lv[sf + 1] = GF(); lv[sf + 2] = GF(); lv[sf + 3] = GF(1);
  lv[sf + 4] = GF(first_bit); lv[sf + 5] = GF(first_bit);
  for (int i = 0; i < 8; i++) lv[sf + 6 + i] = GF(i == 0 ? rest : limbs[i]);
}
// flags.rs:77-134
static inline void generate_flags_next_row(const GF* lv, GF* nv, size_t cur_row, size_t sf) {
  size_t is_final = sf, is_rotate = sf + 1, a = sf + 2, b = sf + 3, fbit = sf + 4, bit = sf + 5, sl = sf + 6, el = sl + 8;
  nv[a] = GF(1) - lv[a]; nv[b] = GF(1) - lv[b];
  size_t num_rows = 2 * INPUT_LIMB_BITS * NUM_INPUT_LIMBS;
  nv[is_final] = cur_row == num_rows - 2 ? GF(1) : GF();
  nv[is_rotate] = (cur_row % (2 * INPUT_LIMB_BITS) == 2 * INPUT_LIMB_BITS - 3) ? GF(1) : GF();
  if (lv[a] == GF(1)) { u64 fl = lv[sl].v, nb = fl % 2; nv[bit] = GF(nb); nv[sl] = GF((fl - nb) / 2); }
  else { nv[bit] = lv[bit]; nv[sl] = lv[sl]; }
  if (lv[is_rotate] == GF(1)) { for (size_t c = sl + 1; c < el; c++) nv[c - 1] = lv[c]; nv[el - 1] = GF(); }
  else { for (size_t c = sl + 1; c < el; c++) nv[c] = lv[c]; }
  nv[fbit] = nv[bit] * nv[b];
}
// flags.rs:136-195
template <class P> static inline void eval_flags(Consumer<P>& yc, const P* lv, const P* nv, size_t sf) {
  size_t is_final_c = sf, is_rotate_c = sf + 1, a = sf + 2, b = sf + 3, fbit = sf + 4, bit_c = sf + 5, sl = sf + 6, el = sl + 8;
  P one = FieldOf<P>::c(1);
  yc.first_row(lv[a]);
  yc.first_row(lv[b] - one);
  P bit = lv[bit_c];
  yc.constraint(bit * bit - bit);
  yc.constraint(bit * lv[b] - lv[fbit]);
  yc.constraint(lv[is_rotate_c] * lv[a]);
  yc.constraint(lv[is_final_c] * lv[is_rotate_c]);
  yc.transition(lv[a] + nv[a] - one);
  yc.transition(lv[b] + nv[b] - one);
  P first_limb = lv[sl], next_first_limb = nv[sl], next_bit = nv[bit_c], is_split = lv[a], is_final = lv[is_final_c];
  P is_not_final = one - is_final;
  yc.transition(is_not_final * is_split * (first_limb - FieldOf<P>::c(2) * next_first_limb - next_bit));
  P is_not_split = one - is_split, is_rotate = lv[is_rotate_c], is_not_rotate_nor_final = one - is_rotate - is_final;
  yc.transition(is_not_split * (next_bit - bit));
  yc.transition(is_not_rotate_nor_final * is_not_split * (first_limb - next_first_limb));
  for (size_t c = sl + 1; c < el; c++) yc.transition(is_rotate * (nv[c - 1] - lv[c]));
  yc.transition(is_rotate * nv[el - 1]);
  for (size_t c = sl + 1; c < el; c++) yc.transition(is_not_rotate_nor_final * (nv[c] - lv[c]));
}

// ---- reference src/utils/pulse.rs ----
typedef std::vector<std::vector<GF>> Cols;
static inline size_t get_pulse_col(size_t start, size_t i) { return start + 1 + 2 * i + 1; }    // pulse.rs:10-12
static inline size_t get_witness_col(size_t start, size_t i) { return start + 1 + 2 * i; }      // pulse.rs:14-16
// pulse.rs:20-43
static inline void generate_pulse(Cols& cols, const std::vector<size_t>& positions) {
  size_t rows = cols[0].size();
  std::vector<GF> counter(rows);
  for (size_t r = 0; r < rows; r++) counter[r] = GF((u64)r);
  cols.push_back(counter);
  size_t base = cols.size();
  cols.resize(base + 2 * positions.size());
#pragma omp parallel for schedule(dynamic)
  for (size_t k = 0; k < positions.size(); k++) {
    size_t pos = positions[k];
    std::vector<GF> witness(rows), pulse(rows);
    for (size_t r = 0; r < rows; r++) witness[r] = (r == pos) ? GF() : gl_inv(counter[r] - GF((u64)pos));
    pulse[pos] = GF(1);
    cols[base + 2 * k] = std::move(witness); cols[base + 2 * k + 1] = std::move(pulse);
  }
}
// pulse.rs:45-63
template <class P> static inline void eval_pulse(Consumer<P>& yc, const P* lv, const P* nv, size_t start, const std::vector<size_t>& positions) {
  P one = FieldOf<P>::c(1);
  P counter = lv[start];
  yc.first_row(counter);
  yc.transition(nv[start] - counter - one);
  for (size_t i = 0; i < positions.size(); i++) {
    P cmp = counter - FieldOf<P>::c((u64)positions[i]);
    P witness = lv[get_witness_col(start, i)], pulse = lv[get_pulse_col(start, i)];
    yc.constraint(cmp * witness + pulse - one);
    yc.constraint(cmp * pulse);
  }
}
// pulse.rs:100-144
static inline void generate_periodic_pulse_witness(Cols& cols, size_t pulse_col, size_t period, size_t first_pulse) {
  size_t rows = cols[pulse_col].size();
  std::vector<GF> counter(rows), witness(rows);
  size_t c = period - first_pulse - 1;
  for (size_t r = 0; r < rows; r++) {
    counter[r] = GF((u64)c);
    assert((c == period - 1) == (cols[pulse_col][r] == GF(1)));
    witness[r] = (c == period - 1) ? GF() : gl_inv(counter[r] - GF((u64)(period - 1)));
    c = (c + 1) % period;
  }
  cols.push_back(counter); cols.push_back(witness);
}
// pulse.rs:146-170
template <class P> static inline void eval_periodic_pulse(Consumer<P>& yc, const P* lv, const P* nv, size_t pulse_col, size_t start, size_t period, size_t first_pulse) {
  P one = FieldOf<P>::c(1);
  P counter = lv[start], witness = lv[start + 1], is_reset = lv[pulse_col], next_counter = nv[start];
  yc.first_row(counter - FieldOf<P>::c((u64)(period - first_pulse - 1)));
  P is_not_reset = one - is_reset;
  yc.transition(is_not_reset * (next_counter - counter - one));
  yc.transition(is_reset * next_counter);
  P delta = counter - FieldOf<P>::c((u64)(period - 1));
  yc.constraint(delta * witness + is_reset - one);
  yc.constraint(delta * is_reset);
}

// ---- reference src/utils/lookup.rs ----
// lookup.rs:13-34
template <class P> static inline void eval_lookups(Consumer<P>& yc, const P* lv, const P* nv, size_t col_in, size_t col_tab) {
  P local_perm_input = lv[col_in], next_perm_table = nv[col_tab], next_perm_input = nv[col_in];
  P diff_input_prev = next_perm_input - local_perm_input;
  P diff_input_table = next_perm_input - next_perm_table;
  yc.constraint(diff_input_prev * diff_input_table);
  yc.last_row(diff_input_table);
}
// lookup.rs:60-111 `permuted_cols` (literal transcription)
static inline void permuted_cols(const std::vector<GF>& inputs, const std::vector<GF>& table, std::vector<GF>& sorted_inputs, std::vector<GF>& permuted_table) {
  size_t n = inputs.size();
  sorted_inputs = inputs;
  std::sort(sorted_inputs.begin(), sorted_inputs.end(), [](GF a, GF b) { return a.v < b.v; });
  std::vector<GF> sorted_table = table;
  std::sort(sorted_table.begin(), sorted_table.end(), [](GF a, GF b) { return a.v < b.v; });
  std::vector<size_t> unused_inds; std::vector<GF> unused_vals;
  permuted_table.assign(n, GF());
  size_t i = 0, j = 0;
  while (j < n && i < n) {
    u64 iv = sorted_inputs[i].v, tv = sorted_table[j].v;
    if (iv > tv) { unused_vals.push_back(sorted_table[j]); j++; }
    else if (iv < tv) {
      if (!unused_vals.empty()) { permuted_table[i] = unused_vals.back(); unused_vals.pop_back(); }
      else unused_inds.push_back(i);
      i++;
    } else { permuted_table[i] = sorted_table[j]; i++; j++; }
  }
  for (size_t k = j; k < n; k++) unused_vals.push_back(sorted_table[k]);
  for (size_t k = i; k < n; k++) unused_inds.push_back(k);
  assert(unused_inds.size() == unused_vals.size());
  for (size_t k = 0; k < unused_inds.size(); k++) permuted_table[unused_inds[k]] = unused_vals[k];
}

// ---- reference src/utils/range_check.rs ----
// range_check.rs:20-47
static inline void generate_u16_range_check(size_t t0, size_t t1, Cols& cols) {
  u64 range_max = 1 << 16; size_t num_rows = cols[0].size();
  assert(num_rows >= range_max);
  std::vector<GF> table(num_rows);
  for (size_t i = 0; i < num_rows; i++) table[i] = GF(i < range_max ? (u64)i : range_max - 1);
  cols.push_back(table);
  size_t base = cols.size(); cols.resize(base + 2 * (t1 - t0));
#pragma omp parallel for schedule(dynamic)
  for (size_t i = t0; i < t1; i++) {
    for (auto& x : cols[i]) { assert(x.v < range_max); (void)x; }
    std::vector<GF> cp, tp; permuted_cols(cols[i], table, cp, tp);
    cols[base + 2 * (i - t0)] = std::move(cp); cols[base + 2 * (i - t0) + 1] = std::move(tp);
  }
}
// range_check.rs:49-68
template <class P> static inline void eval_u16_range_check(Consumer<P>& yc, const P* lv, const P* nv, size_t start, size_t ntargets) {
  for (size_t i = start + 1; i < start + 1 + 2 * ntargets; i += 2) eval_lookups(yc, lv, nv, i, i + 1);
  P cur = lv[start], next = nv[start];
  yc.first_row(cur);
  P incr = next - cur;
  yc.transition(incr * incr - incr);
  yc.last_row(cur - FieldOf<P>::c((1 << 16) - 1));
}
// range_check.rs:96-113
static inline std::vector<std::pair<size_t, size_t>> u16_range_check_pairs(size_t start_lookups, size_t t0, size_t t1) {
  std::vector<std::pair<size_t, size_t>> pairs;
  for (size_t i = 0, pos = t0; pos < t1; i++, pos++) {
    pairs.push_back({start_lookups, start_lookups + 1 + 2 * i + 1});
    pairs.push_back({pos, start_lookups + 1 + 2 * i});
  }
  return pairs;
}
// range_check.rs:116-160
static inline void generate_split_u16_range_check(size_t t0, size_t t1, Cols& cols) {
  u64 range_max = 1 << 8; size_t num_rows = cols[0].size();
  assert((num_rows & (num_rows - 1)) == 0 && range_max <= num_rows);
  std::vector<GF> table(num_rows);
  for (size_t i = 0; i < num_rows; i++) table[i] = GF(i < range_max ? (u64)i : range_max - 1);
  cols.push_back(table);
  size_t base = cols.size(); cols.resize(base + 6 * (t1 - t0));
#pragma omp parallel for schedule(dynamic)
  for (size_t i = t0; i < t1; i++) {
    std::vector<GF> lo(num_rows), hi(num_rows), plo, tlo, phi, thi;
    for (size_t r = 0; r < num_rows; r++) { u64 x = cols[i][r].v; assert(x < (1 << 16)); lo[r] = GF(x & 0xFF); hi[r] = GF((x >> 8) & 0xFF); }
    permuted_cols(lo, table, plo, tlo); permuted_cols(hi, table, phi, thi);
    size_t o = base + 6 * (i - t0);
    cols[o] = std::move(lo); cols[o + 1] = std::move(plo); cols[o + 2] = std::move(tlo);
    cols[o + 3] = std::move(hi); cols[o + 4] = std::move(phi); cols[o + 5] = std::move(thi);
  }
}
// range_check.rs:162-192
template <class P> static inline void eval_split_u16_range_check(Consumer<P>& yc, const P* lv, const P* nv, size_t main_col, size_t t0, size_t t1) {
  for (size_t i = 0, col = t0; col < t1; i++, col++) {
    P original = lv[col], lo = lv[main_col + 1 + 6 * i], hi = lv[main_col + 4 + 6 * i];
    yc.constraint(original - (lo + hi * FieldOf<P>::c(1 << 8)));
  }
  for (size_t i = main_col + 1; i < main_col + 1 + 6 * (t1 - t0); i += 6) { eval_lookups(yc, lv, nv, i + 1, i + 2); eval_lookups(yc, lv, nv, i + 4, i + 5); }
  P cur = lv[main_col], next = nv[main_col];
  yc.first_row(cur);
  P incr = next - cur;
  yc.transition(incr * incr - incr);
  yc.last_row(cur - FieldOf<P>::c((1 << 8) - 1));
}
// range_check.rs:230-246
static inline std::vector<std::pair<size_t, size_t>> split_u16_range_check_pairs(size_t main_col, size_t t0, size_t t1) {
  std::vector<std::pair<size_t, size_t>> pairs;
  for (size_t i = main_col + 1; i < main_col + 1 + 6 * (t1 - t0); i += 6) {
    pairs.push_back({main_col, i + 2}); pairs.push_back({main_col, i + 5});
    pairs.push_back({i, i + 1}); pairs.push_back({i + 3, i + 4});
  }
  return pairs;
}

// plonky2::util::transpose of row-major rows into columns (reference g1/exp.rs:298)
static inline Cols transpose_rows(const std::vector<std::vector<GF>>& rows) {
  size_t n = rows.size(), c = rows[0].size();
  Cols cols(c, std::vector<GF>(n));
  for (size_t r = 0; r < n; r++) for (size_t k = 0; k < c; k++) cols[k][r] = rows[r][k];
  return cols;
}

// The `Stark` trait surface the prover needs (starky::stark::Stark).
struct Air {
  virtual ~Air() {}
  virtual size_t num_columns() const = 0;
  virtual size_t num_public_inputs() const = 0;
  virtual int constraint_degree() const { return 3; }
  virtual std::vector<std::pair<size_t, size_t>> permutation_pairs() const = 0;
  virtual void eval(const GF* lv, const GF* nv, const GF* pi, Consumer<GF>& yc) const = 0;
  virtual void eval_ext(const GF2* lv, const GF2* nv, const GF2* pi, Consumer<GF2>& yc) const = 0;
  int quotient_degree_factor() const { int d = constraint_degree() - 1; return d < 1 ? 1 : d; }
};
#define ORC_AIR_EVAL_IMPL                                                                                         \
  void eval(const GF* lv, const GF* nv, const GF* pi, Consumer<GF>& yc) const override { eval_t<GF>(lv, nv, pi, yc); } \
  void eval_ext(const GF2* lv, const GF2* nv, const GF2* pi, Consumer<GF2>& yc) const override { eval_t<GF2>(lv, nv, pi, yc); }
}  // namespace orc
