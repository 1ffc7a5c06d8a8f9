// ORACLE (test infrastructure).  Fq2 limb-polynomial helpers (reference src/fields/fq2.rs), the G2 affine
// add/double gadget (reference src/curves/g2/muladd.rs) and `G2ExpStark` (reference src/curves/g2/exp.rs).
#pragma once
#include "air_common.hpp"
namespace orc {
// ark_bn254::Fq2 = Fq[u]/(u^2 + 1)
struct Fq2 { Fq c0, c1; bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; } };
static inline Fq2 operator+(const Fq2& a, const Fq2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
static inline Fq2 operator-(const Fq2& a, const Fq2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
static inline Fq2 operator*(const Fq2& a, const Fq2& b) { return {a.c0 * b.c0 - a.c1 * b.c1, a.c0 * b.c1 + a.c1 * b.c0}; }
static inline bool fq2_is_zero(const Fq2& a) { return fq_is_zero(a.c0) && fq_is_zero(a.c1); }
static inline Fq2 fq2_inv(const Fq2& a) {
  assert(!fq2_is_zero(a));  // arkworks panics on division by zero (reference g2/muladd.rs:124,342)
  Fq n = fq_inv(a.c0 * a.c0 + a.c1 * a.c1);
  return {a.c0 * n, -(a.c1 * n)};
}
static inline Fq2 fq2_from_u64(u64 x) { return {fq_from_u64(x), fq_zero()}; }

template <class T, size_t N> using Arr2 = std::array<Arr<T, N>, 2>;
// fq2.rs:20-28
template <class T> static inline Arr2<T, 31> to_wide_fq2(const Arr2<T, 16>& x) { return {widen16(x[0]), widen16(x[1])}; }
// fq2.rs:41-58
template <class T> static inline Arr2<T, 31> pol_mul_fq2(const Arr2<T, 16>& x, const Arr2<T, 16>& y) {
  Arr<T, 31> z0 = pol_sub_normal(pol_mul_wide(x[0], y[0]), pol_mul_wide(x[1], y[1]));
  Arr<T, 31> z1 = pol_add_normal(pol_mul_wide(x[0], y[1]), pol_mul_wide(x[1], y[0]));
  return {z0, z1};
}
// fq2.rs:80-92, :109-121, :138-148
template <class T, size_t N> static inline Arr2<T, N> pol_sub_fq2(const Arr2<T, N>& x, const Arr2<T, N>& y) { return {pol_sub_normal(x[0], y[0]), pol_sub_normal(x[1], y[1])}; }
template <class T, size_t N> static inline Arr2<T, N> pol_add_fq2(const Arr2<T, N>& x, const Arr2<T, N>& y) { return {pol_add_normal(x[0], y[0]), pol_add_normal(x[1], y[1])}; }
template <class T, size_t N> static inline Arr2<T, N> pol_mul_scalar_fq2(const Arr2<T, N>& x, T c) { return {pol_mul_scalar(x[0], c), pol_mul_scalar(x[1], c)}; }
// fq2.rs:164-177
template <class T> static inline void write_fq2(T* lv, const Arr2<T, 16>& v, size_t& cur) { write_u256(lv, v[0], cur); write_u256(lv, v[1], cur); }
template <class T> static inline Arr2<T, 16> read_fq2(const T* lv, size_t& cur) { Arr2<T, 16> r; r[0] = read_u256(lv, cur); r[1] = read_u256(lv, cur); return r; }
// utils.rs:184-187 / :195-201
static inline Arr2<i64, 16> fq2_to_cols(const Fq2& x) { return {fq_to_cols(x.c0), fq_to_cols(x.c1)}; }
static inline Fq2 cols_to_fq2(const Arr2<GF, 16>& c) { return {cols_to_fq(c[0]), cols_to_fq(c[1])}; }
static inline Arr2<i64, 16> positive2_to_i64(const Arr2<GF, 16>& c) { return {positive_column_to_i64(c[0]), positive_column_to_i64(c[1])}; }
static inline Arr2<GF, 16> i64_to_positive2(const Arr2<i64, 16>& c) { return {i64_to_column_positive(c[0]), i64_to_column_positive(c[1])}; }

// g2/muladd.rs:32-40
template <class P> struct G2Output {
  Arr2<P, 16> lambda, new_x, new_y; ModulusAuxZero<P> aux_zeros[2]; ModulusAux<P> auxs[4]; P quot_sign_zeros[2], quot_signs[4];
};
// g2/muladd.rs:42-54 `Default`
static inline G2Output<GF> g2_output_default() {
  G2Output<GF> o;
  Arr<GF, 16> z16 = pol_zero<GF, 16>(); Arr<GF, 17> z17 = pol_zero<GF, 17>(); Arr<GF, 31> z31 = pol_zero<GF, 31>();
  o.lambda = {z16, z16}; o.new_x = o.lambda; o.new_y = o.lambda;
  for (auto& a : o.aux_zeros) { a.quot_abs = z17; a.lo = z31; a.hi = z31; }
  for (auto& a : o.auxs) { a.out_aux_red = z16; a.quot_abs = z17; a.lo = z31; a.hi = z31; }
  for (auto& s : o.quot_sign_zeros) s = GF(1);
  for (auto& s : o.quot_signs) s = GF(1);
  return o;
}
// g2/muladd.rs:56-80
template <class T> static inline void write_g2_output(T* lv, const G2Output<T>& o, size_t& cur) {
  size_t orig = cur;
  write_fq2(lv, o.lambda, cur); write_fq2(lv, o.new_x, cur); write_fq2(lv, o.new_y, cur);
  write_modulus_aux_zero(lv, o.aux_zeros[0], cur); write_modulus_aux_zero(lv, o.aux_zeros[1], cur);
  for (int i = 0; i < 4; i++) write_modulus_aux(lv, o.auxs[i], cur);
  lv[cur++] = o.quot_sign_zeros[0]; lv[cur++] = o.quot_sign_zeros[1];
  for (int i = 0; i < 4; i++) lv[cur++] = o.quot_signs[i];
  assert(cur == orig + 40 * 16); (void)orig;
}
// g2/muladd.rs:82-116
template <class T> static inline G2Output<T> read_g2_output(const T* lv, size_t& cur) {
  G2Output<T> o;
  o.lambda = read_fq2(lv, cur); o.new_x = read_fq2(lv, cur); o.new_y = read_fq2(lv, cur);
  o.aux_zeros[0] = read_modulus_aux_zero(lv, cur); o.aux_zeros[1] = read_modulus_aux_zero(lv, cur);
  for (int i = 0; i < 4; i++) o.auxs[i] = read_modulus_aux(lv, cur);
  o.quot_sign_zeros[0] = lv[cur++]; o.quot_sign_zeros[1] = lv[cur++];
  for (int i = 0; i < 4; i++) o.quot_signs[i] = lv[cur++];
  return o;
}
// shared tail of generate_g2_double (muladd.rs:139-200) / generate_g2_add (:357-413)
static inline G2Output<GF> g2_finish(const Arr2<i64, 16>& lambda_i64, const Arr2<i64, 31>& zero_pol, const Arr2<i64, 31>& new_x_input,
                                     const Arr2<i64, 16>& x1_i64, const Arr2<i64, 16>& y1_i64) {
  G2Output<GF> o;
  o.lambda = i64_to_positive2(lambda_i64);
  for (int i = 0; i < 2; i++) { ModZeroWitness w = generate_modular_zero(zero_pol[i]); o.quot_sign_zeros[i] = w.quot_sign; o.aux_zeros[i] = w.aux; }
  for (int i = 0; i < 2; i++) { ModOpWitness w = generate_modular_op(new_x_input[i]); o.new_x[i] = w.output; o.quot_signs[i] = w.quot_sign; o.auxs[i] = w.aux; }
  Arr2<i64, 16> new_x_i64 = positive2_to_i64(o.new_x);
  Arr2<i64, 16> x_minus_new_x = pol_sub_fq2(x1_i64, new_x_i64);
  Arr2<i64, 31> new_y_input = pol_sub_fq2(pol_mul_fq2(lambda_i64, x_minus_new_x), to_wide_fq2(y1_i64));
  for (int i = 0; i < 2; i++) { ModOpWitness w = generate_modular_op(new_y_input[i]); o.new_y[i] = w.output; o.quot_signs[i + 2] = w.quot_sign; o.auxs[i + 2] = w.aux; }
  return o;
}
// g2/muladd.rs:118-201 `generate_g2_double`
static inline G2Output<GF> generate_g2_double(const Arr2<GF, 16>& x, const Arr2<GF, 16>& y) {
  Fq2 xf = cols_to_fq2(x), yf = cols_to_fq2(y);
  Fq2 lambda = (fq2_from_u64(3) * xf * xf) * fq2_inv(fq2_from_u64(2) * yf);
  Arr2<i64, 16> xi = positive2_to_i64(x), yi = positive2_to_i64(y), li = fq2_to_cols(lambda);
  Arr2<i64, 31> lambda_y_double = pol_mul_scalar_fq2(pol_mul_fq2(li, yi), (i64)2);
  Arr2<i64, 31> x_sq_triple = pol_mul_scalar_fq2(pol_mul_fq2(xi, xi), (i64)3);
  Arr2<i64, 31> zero_pol = pol_sub_fq2(lambda_y_double, x_sq_triple);
  Arr2<i64, 31> double_x = to_wide_fq2(pol_mul_scalar_fq2(xi, (i64)2));
  Arr2<i64, 31> new_x_input = pol_sub_fq2(pol_mul_fq2(li, li), double_x);
  return g2_finish(li, zero_pol, new_x_input, xi, yi);
}
// g2/muladd.rs:330-414 `generate_g2_add`
static inline G2Output<GF> generate_g2_add(const Arr2<GF, 16>& a_x, const Arr2<GF, 16>& a_y, const Arr2<GF, 16>& b_x, const Arr2<GF, 16>& b_y) {
  Fq2 ax = cols_to_fq2(a_x), ay = cols_to_fq2(a_y), bx = cols_to_fq2(b_x), by = cols_to_fq2(b_y);
  Fq2 lambda = (by - ay) * fq2_inv(bx - ax);
  Arr2<i64, 16> axi = positive2_to_i64(a_x), ayi = positive2_to_i64(a_y), bxi = positive2_to_i64(b_x), byi = positive2_to_i64(b_y);
  Arr2<i64, 16> li = fq2_to_cols(lambda);
  Arr2<i64, 16> delta_x = pol_sub_fq2(bxi, axi);
  Arr2<i64, 31> delta_y = to_wide_fq2(pol_sub_fq2(byi, ayi));
  Arr2<i64, 31> zero_pol = pol_sub_fq2(pol_mul_fq2(li, delta_x), delta_y);
  Arr2<i64, 31> x1_add_x2 = to_wide_fq2(pol_add_fq2(axi, bxi));
  Arr2<i64, 31> new_x_input = pol_sub_fq2(pol_mul_fq2(li, li), x1_add_x2);
  return g2_finish(li, zero_pol, new_x_input, axi, ayi);
}
template <class P> static inline void eval_g2_tail(Consumer<P>& yc, P filter, const Arr<P, 16>& modulus, const Arr2<P, 31>& zero_pol, const Arr2<P, 31>& new_x_input,
                                                   const Arr2<P, 16>& x1, const Arr2<P, 16>& y1, const G2Output<P>& o) {
  for (int i = 0; i < 2; i++) eval_modular_zero(yc, filter, modulus, zero_pol[i], o.quot_sign_zeros[i], o.aux_zeros[i]);
  for (int i = 0; i < 2; i++) eval_modular_op(yc, filter, modulus, new_x_input[i], o.new_x[i], o.quot_signs[i], o.auxs[i]);
  Arr2<P, 16> x_minus_new_x = pol_sub_fq2(x1, o.new_x);
  Arr2<P, 31> new_y_input = pol_sub_fq2(pol_mul_fq2(o.lambda, x_minus_new_x), to_wide_fq2(y1));
  for (int i = 0; i < 2; i++) eval_modular_op(yc, filter, modulus, new_y_input[i], o.new_y[i], o.quot_signs[i + 2], o.auxs[i + 2]);
}
// g2/muladd.rs:203-261 `eval_g2_double`
template <class P> static inline void eval_g2_double(Consumer<P>& yc, P filter, const Arr2<P, 16>& x, const Arr2<P, 16>& y, const G2Output<P>& o) {
  Arr<P, 16> modulus = bn254_base_modulus_packfield<P>();
  Arr2<P, 31> lambda_y_double = pol_mul_scalar_fq2(pol_mul_fq2(o.lambda, y), FieldOf<P>::c(2));
  Arr2<P, 31> x_sq_triple = pol_mul_scalar_fq2(pol_mul_fq2(x, x), FieldOf<P>::c(3));
  Arr2<P, 31> zero_pol = pol_sub_fq2(lambda_y_double, x_sq_triple);
  Arr2<P, 31> double_x = to_wide_fq2(pol_mul_scalar_fq2(x, FieldOf<P>::c(2)));
  Arr2<P, 31> new_x_input = pol_sub_fq2(pol_mul_fq2(o.lambda, o.lambda), double_x);
  eval_g2_tail(yc, filter, modulus, zero_pol, new_x_input, x, y, o);
}
// g2/muladd.rs:416-472 `eval_g2_add`
template <class P> static inline void eval_g2_add(Consumer<P>& yc, P filter, const Arr2<P, 16>& a_x, const Arr2<P, 16>& a_y, const Arr2<P, 16>& b_x, const Arr2<P, 16>& b_y,
                                                  const G2Output<P>& o) {
  Arr<P, 16> modulus = bn254_base_modulus_packfield<P>();
  Arr2<P, 16> delta_x = pol_sub_fq2(b_x, a_x);
  Arr2<P, 31> delta_y = to_wide_fq2(pol_sub_fq2(b_y, a_y));
  Arr2<P, 31> zero_pol = pol_sub_fq2(pol_mul_fq2(o.lambda, delta_x), delta_y);
  Arr2<P, 31> x1_add_x2 = to_wide_fq2(pol_add_fq2(a_x, b_x));
  Arr2<P, 31> new_x_input = pol_sub_fq2(pol_mul_fq2(o.lambda, o.lambda), x1_add_x2);
  eval_g2_tail(yc, filter, modulus, zero_pol, new_x_input, a_x, a_y, o);
}
// equals.rs:121-130
template <class P> static inline void fq2_equal_transition(Consumer<P>& yc, P filter, const Arr2<P, 16>& x, const Arr2<P, 16>& y) {
  fq_equal_transition(yc, filter, x[0], y[0]); fq_equal_transition(yc, filter, x[1], y[1]);
}

struct G2Point { U256 x0, x1, y0, y1; };  // x = x0 + x1 u, y = y0 + y1 u
// g2/exp.rs:90-95 `G2ExpIONative`
struct G2ExpIONative { G2Point x, offset; u32 exp_val[8]; G2Point output; };

struct G2ExpStark : Air {
  size_t num_io;
  // g2/exp.rs:6-34 `constants`
  size_t start_flags_col = 48 * 16, num_main_cols = start_flags_col + NUM_FLAGS_COLS, start_periodic_pulse_col = num_main_cols,
         start_io_pulses_col = start_periodic_pulse_col + 2, start_lookups_col, start_range_check_col = 0, num_range_check_cols = 48 * 16 - 6,
         end_range_check_col = num_range_check_cols, n_columns, n_public_inputs;
  explicit G2ExpStark(size_t n) : num_io(n) {
    start_lookups_col = start_io_pulses_col + 1 + 4 * num_io;
    n_columns = start_lookups_col + 1 + 2 * num_range_check_cols;
    n_public_inputs = 13 * NUM_INPUT_LIMBS * num_io;
  }
  size_t num_columns() const override { return n_columns; }
  size_t num_public_inputs() const override { return n_public_inputs; }
  std::vector<std::pair<size_t, size_t>> permutation_pairs() const override { return u16_range_check_pairs(start_lookups_col, start_range_check_col, end_range_check_col); }
  static Arr2<GF, 16> coord(const U256& c0, const U256& c1) { return {i64_to_column_positive(fq_to_cols(fq_from_u256(c0))), i64_to_column_positive(fq_to_cols(fq_from_u256(c1)))}; }
  // g2/exp.rs:180-205
  void generate_first_row(GF* lv, const G2Point& x, const G2Point& offset) const {
    Arr2<GF, 16> a_x = coord(x.x0, x.x1), a_y = coord(x.y0, x.y1), b_x = coord(offset.x0, offset.x1), b_y = coord(offset.y0, offset.y1);
    G2Output<GF> out = lv[start_flags_col + 4] == GF(1) ? generate_g2_add(a_x, a_y, b_x, b_y) : g2_output_default();
    size_t cur = 0;
    write_fq2(lv, a_x, cur); write_fq2(lv, a_y, cur); write_fq2(lv, b_x, cur); write_fq2(lv, b_y, cur);
    write_g2_output(lv, out, cur);
  }
  // g2/exp.rs:207-245
  void generate_next_row(const GF* lv, GF* nv) const {
    size_t is_double_col = start_flags_col + 2, is_add_col = start_flags_col + 4;
    size_t cur = 0;
    Arr2<GF, 16> a_x = read_fq2(lv, cur), a_y = read_fq2(lv, cur), b_x = read_fq2(lv, cur), b_y = read_fq2(lv, cur);
    G2Output<GF> output = read_g2_output(lv, cur);
    Arr2<GF, 16> nax = a_x, nay = a_y, nbx = b_x, nby = b_y;
    if (lv[is_double_col] == GF(1)) { nax = output.new_x; nay = output.new_y; }
    else if (lv[is_add_col] == GF(1)) { nbx = output.new_x; nby = output.new_y; }
    G2Output<GF> next_output = nv[is_double_col] == GF(1) ? generate_g2_double(nax, nay)
                               : nv[is_add_col] == GF(1) ? generate_g2_add(nax, nay, nbx, nby) : g2_output_default();
    cur = 0;
    write_fq2(nv, nax, cur); write_fq2(nv, nay, cur); write_fq2(nv, nbx, cur); write_fq2(nv, nby, cur);
    write_g2_output(nv, next_output, cur);
  }
  // g2/exp.rs:270-304; the chain result b on the last row is returned (the reference asserts it equals
  // arkworks' x*e + offset; the caller compares it with big-integer group arithmetic).
  std::vector<std::vector<GF>> generate_trace_for_one_block(const G2Point& x, const G2Point& offset, const u32 exp_val[8], G2Point* result) const {
    size_t num_rows = 2 * INPUT_LIMB_BITS * NUM_INPUT_LIMBS;
    std::vector<GF> lv(num_main_cols);
    generate_flags_first_row(lv.data(), start_flags_col, exp_val);
    generate_first_row(lv.data(), x, offset);
    std::vector<std::vector<GF>> rows; rows.push_back(lv);
    for (size_t i = 0; i + 1 < num_rows; i++) {
      std::vector<GF> nv(lv.size());
      generate_flags_next_row(lv.data(), nv.data(), i, start_flags_col);
      generate_next_row(lv.data(), nv.data());
      rows.push_back(nv); lv = nv;
    }
    size_t cur = 4 * 16;
    Arr2<GF, 16> bx = read_fq2(rows.back().data(), cur), by = read_fq2(rows.back().data(), cur);
    result->x0 = fq_to_u256(cols_to_fq(bx[0])); result->x1 = fq_to_u256(cols_to_fq(bx[1]));
    result->y0 = fq_to_u256(cols_to_fq(by[0])); result->y1 = fq_to_u256(cols_to_fq(by[1]));
    return rows;
  }
  // g2/exp.rs:306-335 (blocks are independent; run under OpenMP)
  Cols generate_trace(const std::vector<G2ExpIONative>& inputs, std::vector<G2Point>* results = nullptr) const {
    assert(inputs.size() == num_io);
    size_t nr = 2 * INPUT_LIMB_BITS * NUM_INPUT_LIMBS;
    std::vector<std::vector<GF>> rows(num_io * nr);
    std::vector<G2Point> res(num_io);
#pragma omp parallel for schedule(dynamic)
    for (size_t k = 0; k < num_io; k++) {
      auto blk = generate_trace_for_one_block(inputs[k].x, inputs[k].offset, inputs[k].exp_val, &res[k]);
      for (size_t r = 0; r < nr; r++) rows[k * nr + r] = std::move(blk[r]);
    }
    if (results) *results = res;
    Cols cols = transpose_rows(rows);
    rows.clear(); rows.shrink_to_fit();
    size_t rotation_period = 2 * INPUT_LIMB_BITS;
    generate_periodic_pulse_witness(cols, start_flags_col + 1, rotation_period, rotation_period - 2);
    generate_pulse(cols, G1ExpStark::get_pulse_positions(num_io));   // g2/exp.rs:50 imports g1's get_pulse_positions
    generate_u16_range_check(start_range_check_col, end_range_check_col, cols);
    return cols;
  }
  // g2/exp.rs:337-342 + :139-156
  std::vector<GF> generate_public_inputs(const std::vector<G2ExpIONative>& inputs) const {
    std::vector<GF> pi;
    auto push = [&](const U256& v) { auto c = u256_to_u32_columns(v); pi.insert(pi.end(), c.begin(), c.end()); };
    auto pushp = [&](const G2Point& p) { push(p.x0); push(p.x1); push(p.y0); push(p.y1); };
    for (auto& in : inputs) {
      pushp(in.x); pushp(in.offset);
      for (int i = 0; i < 8; i++) pi.push_back(GF(in.exp_val[i]));
      pushp(in.output);
    }
    return pi;
  }
  // g2/exp.rs:346-507
  template <class P> void eval_t(const P* lv, const P* nv, const P* pi, Consumer<P>& yc) const {
    P one = FieldOf<P>::c(1);
    size_t is_final_col = start_flags_col, is_double_col = start_flags_col + 2, is_add_col = start_flags_col + 4, start_limbs_col = start_flags_col + 6;
    size_t cur = 0;
    Arr2<P, 16> a_x = read_fq2(lv, cur), a_y = read_fq2(lv, cur), b_x = read_fq2(lv, cur), b_y = read_fq2(lv, cur);
    G2Output<P> output = read_g2_output(lv, cur);
    P is_add = lv[is_add_col], is_double = lv[is_double_col], is_final = lv[is_final_col], is_not_final = one - is_final;
    P sum_is_output = tzero<P>();
    for (size_t i = 1; i < 2 * num_io; i += 2) sum_is_output = sum_is_output + lv[get_pulse_col(start_io_pulses_col, i)];
    yc.constraint(is_final - sum_is_output);
    cur = 0;
    for (size_t i = 0; i < 2 * num_io; i += 2) {
      Arr<P, 8> io[13];   // x[4] offset[4] exp_val output[4]   (g2/exp.rs:158-178 `read_g2_exp_io`)
      for (int k = 0; k < 13; k++) { for (int j = 0; j < 8; j++) io[k][j] = pi[cur + j]; cur += 8; }
      P is_ith_input = lv[get_pulse_col(start_io_pulses_col, i)], is_ith_output = lv[get_pulse_col(start_io_pulses_col, i + 1)];
      Arr<P, 8> row[4] = {u16_columns_to_u32_columns(a_x[0]), u16_columns_to_u32_columns(a_x[1]), u16_columns_to_u32_columns(a_y[0]), u16_columns_to_u32_columns(a_y[1])};
      Arr<P, 8> rowb[4] = {u16_columns_to_u32_columns(b_x[0]), u16_columns_to_u32_columns(b_x[1]), u16_columns_to_u32_columns(b_y[0]), u16_columns_to_u32_columns(b_y[1])};
      for (int k = 0; k < 4; k++) vec_equal(yc, is_ith_input, io[k], row[k]);
      for (int k = 0; k < 4; k++) vec_equal(yc, is_ith_input, io[4 + k], rowb[k]);
      for (int k = 0; k < 4; k++) vec_equal(yc, is_ith_output, io[9 + k], rowb[k]);
      Arr<P, 8> limbs; for (int j = 0; j < 8; j++) limbs[j] = lv[start_limbs_col + j];
      limbs[0] = limbs[0] * FieldOf<P>::c(2) + is_add;
      vec_equal(yc, is_ith_input, io[8], limbs);
    }
    cur = 0;
    Arr2<P, 16> next_a_x = read_fq2(nv, cur), next_a_y = read_fq2(nv, cur), next_b_x = read_fq2(nv, cur), next_b_y = read_fq2(nv, cur);
    { P f = is_not_final * is_double;
      fq2_equal_transition(yc, f, next_a_x, output.new_x); fq2_equal_transition(yc, f, next_a_y, output.new_y);
      fq2_equal_transition(yc, f, next_b_x, b_x); fq2_equal_transition(yc, f, next_b_y, b_y); }
    { P f = is_not_final * is_add;
      fq2_equal_transition(yc, f, next_a_x, a_x); fq2_equal_transition(yc, f, next_a_y, a_y);
      fq2_equal_transition(yc, f, next_b_x, output.new_x); fq2_equal_transition(yc, f, next_b_y, output.new_y); }
    { P f = is_not_final * (one - is_double - is_add);
      fq2_equal_transition(yc, f, next_a_x, a_x); fq2_equal_transition(yc, f, next_a_y, a_y);
      fq2_equal_transition(yc, f, next_b_x, b_x); fq2_equal_transition(yc, f, next_b_y, b_y); }
    eval_flags(yc, lv, nv, start_flags_col);
    eval_g2_add(yc, is_add, a_x, a_y, b_x, b_y, output);
    eval_g2_double(yc, is_double, a_x, a_y, output);
    eval_flags(yc, lv, nv, start_flags_col);   // emitted twice in the reference (g2/exp.rs:474 and :479-484)
    eval_periodic_pulse(yc, lv, nv, start_flags_col + 1, start_periodic_pulse_col, 2 * INPUT_LIMB_BITS, 2 * INPUT_LIMB_BITS - 2);
    eval_pulse(yc, lv, nv, start_io_pulses_col, G1ExpStark::get_pulse_positions(num_io));
    eval_u16_range_check(yc, lv, nv, start_lookups_col, num_range_check_cols);
  }
  ORC_AIR_EVAL_IMPL
};
}  // namespace orc
