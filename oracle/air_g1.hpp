// ORACLE (test infrastructure).  G1 affine add/double gadget (reference src/curves/g1/muladd.rs) and
// `G1ExpStark` (reference src/curves/g1/exp.rs).
#pragma once
#include "air_common.hpp"
namespace orc {
// g1/muladd.rs:49-59
template <class P> struct G1Output {
  Arr<P, 16> lambda, new_x, new_y; ModulusAuxZero<P> aux_zero; ModulusAux<P> aux_x, aux_y; P quot_sign_zero, quot_sign_x, quot_sign_y;
};
// g1/muladd.rs:61-75 `Default`
static inline G1Output<GF> g1_output_default() {
  G1Output<GF> o;
  o.lambda = pol_zero<GF, 16>(); o.new_x = o.lambda; o.new_y = o.lambda;
  o.aux_zero.quot_abs = pol_zero<GF, 17>(); o.aux_zero.lo = pol_zero<GF, 31>(); o.aux_zero.hi = o.aux_zero.lo;
  o.aux_x.out_aux_red = o.lambda; o.aux_x.quot_abs = o.aux_zero.quot_abs; o.aux_x.lo = o.aux_zero.lo; o.aux_x.hi = o.aux_zero.lo; o.aux_y = o.aux_x;
  o.quot_sign_zero = GF(1); o.quot_sign_x = GF(1); o.quot_sign_y = GF(1);
  return o;
}
// g1/muladd.rs:79-94
template <class T> static inline void write_g1_output(T* lv, const G1Output<T>& o, size_t& cur) {
  write_u256(lv, o.lambda, cur); write_u256(lv, o.new_x, cur); write_u256(lv, o.new_y, cur);
  write_modulus_aux_zero(lv, o.aux_zero, cur); write_modulus_aux(lv, o.aux_x, cur); write_modulus_aux(lv, o.aux_y, cur);
  lv[cur++] = o.quot_sign_zero; lv[cur++] = o.quot_sign_x; lv[cur++] = o.quot_sign_y;
}
// g1/muladd.rs:98-122
template <class T> static inline G1Output<T> read_g1_output(const T* lv, size_t& cur) {
  G1Output<T> o;
  o.lambda = read_u256(lv, cur); o.new_x = read_u256(lv, cur); o.new_y = read_u256(lv, cur);
  o.aux_zero = read_modulus_aux_zero(lv, cur); o.aux_x = read_modulus_aux(lv, cur); o.aux_y = read_modulus_aux(lv, cur);
  o.quot_sign_zero = lv[cur++]; o.quot_sign_x = lv[cur++]; o.quot_sign_y = lv[cur++];
  return o;
}
static inline G1Output<GF> g1_finish(const Arr<i64, 16>& lambda_i64, const Arr<i64, 31>& zero_pol, const Arr<i64, 31>& new_x_input,
                                     const Arr<i64, 16>& x1_i64, const Arr<i64, 16>& y1_i64) {
  G1Output<GF> o;
  o.lambda = i64_to_column_positive(lambda_i64);
  ModZeroWitness wz = generate_modular_zero(zero_pol);
  o.quot_sign_zero = wz.quot_sign; o.aux_zero = wz.aux;
  ModOpWitness wx = generate_modular_op(new_x_input);
  o.new_x = wx.output; o.quot_sign_x = wx.quot_sign; o.aux_x = wx.aux;
  Arr<i64, 16> new_x_i64 = positive_column_to_i64(wx.output);
  Arr<i64, 16> x1_minus_new_x = pol_sub_normal(x1_i64, new_x_i64);
  Arr<i64, 31> l = pol_mul_wide(lambda_i64, x1_minus_new_x);
  Arr<i64, 31> new_y_input = pol_sub_normal(l, widen16(y1_i64));
  ModOpWitness wy = generate_modular_op(new_y_input);
  o.new_y = wy.output; o.quot_sign_y = wy.quot_sign; o.aux_y = wy.aux;
  return o;
}
// g1/muladd.rs:124-177 `generate_g1_add`
static inline G1Output<GF> generate_g1_add(const Arr<GF, 16>& a_x, const Arr<GF, 16>& a_y, const Arr<GF, 16>& b_x, const Arr<GF, 16>& b_y) {
  Fq ax = cols_to_fq(a_x), ay = cols_to_fq(a_y), bx = cols_to_fq(b_x), by = cols_to_fq(b_y);
  Fq lambda = (by - ay) * fq_inv(bx - ax);
  Arr<i64, 16> axi = positive_column_to_i64(a_x), ayi = positive_column_to_i64(a_y), bxi = positive_column_to_i64(b_x), byi = positive_column_to_i64(b_y);
  Arr<i64, 16> li = fq_to_cols(lambda);
  Arr<i64, 16> delta_x = pol_sub_normal(bxi, axi);
  Arr<i64, 31> delta_y = pol_sub(byi, ayi);
  Arr<i64, 31> zero_pol = pol_sub_normal(pol_mul_wide(li, delta_x), delta_y);
  Arr<i64, 31> new_x_input = pol_sub_normal(pol_mul_wide(li, li), pol_add(axi, bxi));
  return g1_finish(li, zero_pol, new_x_input, axi, ayi);
}
// g1/muladd.rs:409-460 `generate_g1_double`
static inline G1Output<GF> generate_g1_double(const Arr<GF, 16>& x, const Arr<GF, 16>& y) {
  Fq xf = cols_to_fq(x), yf = cols_to_fq(y);
  Fq lambda = (fq_from_u64(3) * xf * xf) * fq_inv(fq_from_u64(2) * yf);
  Arr<i64, 16> xi = positive_column_to_i64(x), yi = positive_column_to_i64(y), li = fq_to_cols(lambda);
  Arr<i64, 31> lambda_y_double = pol_mul_scalar(pol_mul_wide(li, yi), (i64)2);
  Arr<i64, 31> x_sq_triple = pol_mul_scalar(pol_mul_wide(xi, xi), (i64)3);
  Arr<i64, 31> zero_pol = pol_sub_normal(lambda_y_double, x_sq_triple);
  Arr<i64, 31> double_x = pol_mul_scalar(widen16(xi), (i64)2);
  Arr<i64, 31> new_x_input = pol_sub_normal(pol_mul_wide(li, li), double_x);
  return g1_finish(li, zero_pol, new_x_input, xi, yi);
}
template <class P> static inline void eval_g1_tail(Consumer<P>& yc, P filter, const Arr<P, 16>& modulus, const Arr<P, 31>& zero_pol, const Arr<P, 31>& new_x_input,
                                                   const Arr<P, 16>& x1, const Arr<P, 16>& y1, const G1Output<P>& o) {
  eval_modular_zero(yc, filter, modulus, zero_pol, o.quot_sign_zero, o.aux_zero);
  eval_modular_op(yc, filter, modulus, new_x_input, o.new_x, o.quot_sign_x, o.aux_x);
  Arr<P, 16> x1_minus_new_x = pol_sub_normal(x1, o.new_x);
  Arr<P, 31> new_y_input = pol_sub_normal(pol_mul_wide(o.lambda, x1_minus_new_x), widen16(y1));
  eval_modular_op(yc, filter, modulus, new_y_input, o.new_y, o.quot_sign_y, o.aux_y);
}
// g1/muladd.rs:179-230 `eval_g1_add`
template <class P> static inline void eval_g1_add(Consumer<P>& yc, P filter, const Arr<P, 16>& a_x, const Arr<P, 16>& a_y, const Arr<P, 16>& b_x, const Arr<P, 16>& b_y, const G1Output<P>& o) {
  Arr<P, 16> modulus = bn254_base_modulus_packfield<P>();
  Arr<P, 16> delta_x = pol_sub_normal(b_x, a_x);
  Arr<P, 31> delta_y = pol_sub(b_y, a_y);
  Arr<P, 31> zero_pol = pol_sub_normal(pol_mul_wide(o.lambda, delta_x), delta_y);
  Arr<P, 31> new_x_input = pol_sub_normal(pol_mul_wide(o.lambda, o.lambda), pol_add(a_x, b_x));
  eval_g1_tail(yc, filter, modulus, zero_pol, new_x_input, a_x, a_y, o);
}
// g1/muladd.rs:291-342 `eval_g1_double`
template <class P> static inline void eval_g1_double(Consumer<P>& yc, P filter, const Arr<P, 16>& x, const Arr<P, 16>& y, const G1Output<P>& o) {
  Arr<P, 16> modulus = bn254_base_modulus_packfield<P>();
  Arr<P, 31> lambda_y_double = pol_mul_scalar(pol_mul_wide(o.lambda, y), FieldOf<P>::c(2));
  Arr<P, 31> x_sq_triple = pol_mul_scalar(pol_mul_wide(x, x), FieldOf<P>::c(3));
  Arr<P, 31> zero_pol = pol_sub_normal(lambda_y_double, x_sq_triple);
  Arr<P, 31> new_x_input = pol_sub_normal(pol_mul_wide(o.lambda, o.lambda), pol_add(x, x));
  eval_g1_tail(yc, filter, modulus, zero_pol, new_x_input, x, y, o);
}

struct G1Point { U256 x, y; };
// g1/exp.rs:88-93 `G1ExpIONative`
struct G1ExpIONative { G1Point x, offset; u32 exp_val[8]; G1Point output; };

struct G1ExpStark : Air {
  size_t num_io;
  // g1/exp.rs:6-34 `constants`
  size_t start_flags_col = 24 * 16, num_main_cols = start_flags_col + NUM_FLAGS_COLS, start_periodic_pulse_col = num_main_cols,
         start_io_pulses_col = start_periodic_pulse_col + 2, start_lookups_col, start_range_check_col = 0, num_range_check_cols = 24 * 16 - 3,
         end_range_check_col = num_range_check_cols, n_columns, n_public_inputs;
  explicit G1ExpStark(size_t n) : num_io(n) {
    start_lookups_col = start_io_pulses_col + 1 + 4 * num_io;
    n_columns = start_lookups_col + 1 + 2 * num_range_check_cols;
    n_public_inputs = 7 * NUM_INPUT_LIMBS * num_io;
  }
  size_t num_columns() const override { return n_columns; }
  size_t num_public_inputs() const override { return n_public_inputs; }
  std::vector<std::pair<size_t, size_t>> permutation_pairs() const override { return u16_range_check_pairs(start_lookups_col, start_range_check_col, end_range_check_col); }
  // g1/exp.rs:153-163
  static std::vector<size_t> get_pulse_positions(size_t num_io) {
    size_t nr = 2 * INPUT_LIMB_BITS * NUM_INPUT_LIMBS; std::vector<size_t> p;
    for (size_t i = 0; i < num_io; i++) { p.push_back(i * nr); p.push_back(i * nr + nr - 1); }
    return p;
  }
  // g1/exp.rs:165-190
  void generate_first_row(GF* lv, const G1Point& x, const G1Point& offset) const {
    Arr<GF, 16> a_x = i64_to_column_positive(fq_to_cols(fq_from_u256(x.x))), a_y = i64_to_column_positive(fq_to_cols(fq_from_u256(x.y)));
    Arr<GF, 16> b_x = i64_to_column_positive(fq_to_cols(fq_from_u256(offset.x))), b_y = i64_to_column_positive(fq_to_cols(fq_from_u256(offset.y)));
    G1Output<GF> out = lv[start_flags_col + 4] == GF(1) ? generate_g1_add(a_x, a_y, b_x, b_y) : g1_output_default();
    size_t cur = 0;
    write_u256(lv, a_x, cur); write_u256(lv, a_y, cur); write_u256(lv, b_x, cur); write_u256(lv, b_y, cur);
    write_g1_output(lv, out, cur);
  }
  // g1/exp.rs:192-230
  void generate_next_row(const GF* lv, GF* nv) const {
    size_t is_double_col = start_flags_col + 2, is_add_col = start_flags_col + 4;
    size_t cur = 0;
    Arr<GF, 16> a_x = read_u256(lv, cur), a_y = read_u256(lv, cur), b_x = read_u256(lv, cur), b_y = read_u256(lv, cur);
    G1Output<GF> output = read_g1_output(lv, cur);
    Arr<GF, 16> nax = a_x, nay = a_y, nbx = b_x, nby = b_y;
    if (lv[is_double_col] == GF(1)) { nax = output.new_x; nay = output.new_y; }
    else if (lv[is_add_col] == GF(1)) { nbx = output.new_x; nby = output.new_y; }
    G1Output<GF> next_output = nv[is_double_col] == GF(1) ? generate_g1_double(nax, nay)
                               : nv[is_add_col] == GF(1) ? generate_g1_add(nax, nay, nbx, nby) : g1_output_default();
    cur = 0;
    write_u256(nv, nax, cur); write_u256(nv, nay, cur); write_u256(nv, nbx, cur); write_u256(nv, nby, cur);
    write_g1_output(nv, next_output, cur);
  }
  // g1/exp.rs:255-288; returns rows and the chain result b on the last row (the reference asserts it
  // equals arkworks' x*e+offset; here the caller compares it with io.output).
  std::vector<std::vector<GF>> generate_trace_for_one_block(const G1Point& x, const G1Point& offset, const u32 exp_val[8], G1Point* result) const {
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
    size_t cur = 32;
    Arr<GF, 16> bx = read_u256(rows.back().data(), cur), by = read_u256(rows.back().data(), cur);
    result->x = fq_to_u256(cols_to_fq(bx)); result->y = fq_to_u256(cols_to_fq(by));
    return rows;
  }
  // g1/exp.rs:290-318.  (The reference loops serially over inputs; blocks are independent, so the
  // oracle runs them under OpenMP -- same rows.)
  Cols generate_trace(const std::vector<G1ExpIONative>& inputs, std::vector<G1Point>* results = nullptr) const {
    assert(inputs.size() == num_io);
    size_t nr = 2 * INPUT_LIMB_BITS * NUM_INPUT_LIMBS;
    std::vector<std::vector<GF>> rows(num_io * nr);
    std::vector<G1Point> res(num_io);
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
    generate_pulse(cols, get_pulse_positions(num_io));
    generate_u16_range_check(start_range_check_col, end_range_check_col, cols);
    return cols;
  }
  // g1/exp.rs:320-327 + :124-135
  std::vector<GF> generate_public_inputs(const std::vector<G1ExpIONative>& inputs) const {
    std::vector<GF> pi;
    auto push = [&](const U256& v) { auto c = u256_to_u32_columns(v); pi.insert(pi.end(), c.begin(), c.end()); };
    for (auto& in : inputs) {
      push(in.x.x); push(in.x.y); push(in.offset.x); push(in.offset.y);
      for (int i = 0; i < 8; i++) pi.push_back(GF(in.exp_val[i]));
      push(in.output.x); push(in.output.y);
    }
    return pi;
  }
  // g1/exp.rs:331-495
  template <class P> void eval_t(const P* lv, const P* nv, const P* pi, Consumer<P>& yc) const {
    P one = FieldOf<P>::c(1);
    size_t is_final_col = start_flags_col, is_double_col = start_flags_col + 2, is_add_col = start_flags_col + 4, start_limbs_col = start_flags_col + 6;
    size_t cur = 0;
    Arr<P, 16> a_x = read_u256(lv, cur), a_y = read_u256(lv, cur), b_x = read_u256(lv, cur), b_y = read_u256(lv, cur);
    G1Output<P> output = read_g1_output(lv, cur);
    P is_add = lv[is_add_col], is_double = lv[is_double_col], is_final = lv[is_final_col], is_not_final = one - is_final;
    P sum_is_output = tzero<P>();
    for (size_t i = 1; i < 2 * num_io; i += 2) sum_is_output = sum_is_output + lv[get_pulse_col(start_io_pulses_col, i)];
    yc.constraint(is_final - sum_is_output);
    cur = 0;
    for (size_t i = 0; i < 2 * num_io; i += 2) {
      Arr<P, 8> io[7];
      for (int k = 0; k < 7; k++) { for (int j = 0; j < 8; j++) io[k][j] = pi[cur + j]; cur += 8; }
      P is_ith_input = lv[get_pulse_col(start_io_pulses_col, i)], is_ith_output = lv[get_pulse_col(start_io_pulses_col, i + 1)];
      Arr<P, 8> x_x = u16_columns_to_u32_columns(a_x), x_y = u16_columns_to_u32_columns(a_y), bx32 = u16_columns_to_u32_columns(b_x), by32 = u16_columns_to_u32_columns(b_y);
      vec_equal(yc, is_ith_input, io[0], x_x); vec_equal(yc, is_ith_input, io[1], x_y);
      vec_equal(yc, is_ith_input, io[2], bx32); vec_equal(yc, is_ith_input, io[3], by32);
      vec_equal(yc, is_ith_output, io[5], bx32); vec_equal(yc, is_ith_output, io[6], by32);
      Arr<P, 8> limbs; for (int j = 0; j < 8; j++) limbs[j] = lv[start_limbs_col + j];
      limbs[0] = limbs[0] * FieldOf<P>::c(2) + is_add;
      vec_equal(yc, is_ith_input, io[4], limbs);
    }
    cur = 0;
    Arr<P, 16> next_a_x = read_u256(nv, cur), next_a_y = read_u256(nv, cur), next_b_x = read_u256(nv, cur), next_b_y = read_u256(nv, cur);
    { P f = is_not_final * is_double;
      fq_equal_transition(yc, f, next_a_x, output.new_x); fq_equal_transition(yc, f, next_a_y, output.new_y);
      fq_equal_transition(yc, f, next_b_x, b_x); fq_equal_transition(yc, f, next_b_y, b_y); }
    { P f = is_not_final * is_add;
      fq_equal_transition(yc, f, next_a_x, a_x); fq_equal_transition(yc, f, next_a_y, a_y);
      fq_equal_transition(yc, f, next_b_x, output.new_x); fq_equal_transition(yc, f, next_b_y, output.new_y); }
    { P f = is_not_final * (one - is_double - is_add);
      fq_equal_transition(yc, f, next_a_x, a_x); fq_equal_transition(yc, f, next_a_y, a_y);
      fq_equal_transition(yc, f, next_b_x, b_x); fq_equal_transition(yc, f, next_b_y, b_y); }
    eval_flags(yc, lv, nv, start_flags_col);
    eval_g1_add(yc, is_add, a_x, a_y, b_x, b_y, output);
    eval_g1_double(yc, is_double, a_x, a_y, output);
    eval_flags(yc, lv, nv, start_flags_col);   // emitted twice in the reference (g1/exp.rs:462 and :467-472)
    eval_periodic_pulse(yc, lv, nv, start_flags_col + 1, start_periodic_pulse_col, 2 * INPUT_LIMB_BITS, 2 * INPUT_LIMB_BITS - 2);
    eval_pulse(yc, lv, nv, start_io_pulses_col, get_pulse_positions(num_io));
    eval_u16_range_check(yc, lv, nv, start_lookups_col, num_range_check_cols);
  }
  ORC_AIR_EVAL_IMPL
};
}  // namespace orc
