// ORACLE (test infrastructure).  Fq mul gadget + `FqExpStark` (reference src/fields/fq/{mul,exp}.rs), Fq12 mul
// gadget (reference src/fields/fq12/mul.rs), `Fq12ExpStark` (reference src/fields/fq12/exp.rs) and the 64-bit
// exponent variant `Fq12ExpU64Stark` (reference src/fields/fq12_u64/{flags_u64,exp_u64}.rs).
// Fq12 elements are handled in the flat `MyFq12` coefficient order of plonky2-bn254 (external):
// element = sum_{i<6} (c[i] + c[i+6] u) w^i, u^2 = -1, w^6 = 9 + u -- the order `pol_mul_fq12` forces.
#pragma once
#include "air_common.hpp"
#include "air_g1.hpp"
#include <stdexcept>
#include <string>
namespace orc {
template <class T> using Arr12 = std::array<Arr<T, 16>, 12>;
template <class T> using Wide12 = std::array<Arr<T, 31>, 12>;

// ---------------- Fq mul gadget: fq/mul.rs ----------------
template <class P> struct FqOutput { Arr<P, 16> output; ModulusAux<P> aux; P quot_sign; };
static inline ModulusAux<GF> modulus_aux_default() {
  ModulusAux<GF> a; a.out_aux_red = pol_zero<GF, 16>(); a.quot_abs = pol_zero<GF, 17>(); a.lo = pol_zero<GF, 31>(); a.hi = a.lo; return a;
}
// fq/mul.rs:24-32
static inline FqOutput<GF> fq_output_default() { FqOutput<GF> o; o.output = pol_zero<GF, 16>(); o.aux = modulus_aux_default(); o.quot_sign = GF(1); return o; }
// fq/mul.rs:34-45
static inline FqOutput<GF> generate_fq_mul(const Arr<GF, 16>& x, const Arr<GF, 16>& y) {
  ModOpWitness w = generate_modular_op(pol_mul_wide(positive_column_to_i64(x), positive_column_to_i64(y)));
  FqOutput<GF> o; o.output = w.output; o.aux = w.aux; o.quot_sign = w.quot_sign; return o;
}
// fq/mul.rs:49-67
template <class T> static inline void write_fq_output(T* lv, const FqOutput<T>& o, size_t& cur) { write_u256(lv, o.output, cur); write_modulus_aux(lv, o.aux, cur); lv[cur++] = o.quot_sign; }
template <class T> static inline FqOutput<T> read_fq_output(const T* lv, size_t& cur) { FqOutput<T> o; o.output = read_u256(lv, cur); o.aux = read_modulus_aux(lv, cur); o.quot_sign = lv[cur++]; return o; }
// fq/mul.rs:69-87
template <class P> static inline void eval_fq_mul(Consumer<P>& yc, P filter, const Arr<P, 16>& x, const Arr<P, 16>& y, const FqOutput<P>& o) {
  eval_modular_op(yc, filter, bn254_base_modulus_packfield<P>(), pol_mul_wide(x, y), o.output, o.quot_sign, o.aux);
}

// fq/exp.rs:88-93 `FqExpIONative`
struct FqExpIONative { U256 x, offset; u32 exp_val[8]; U256 output; };

struct FqExpStark : Air {
  size_t num_io;
  // fq/exp.rs:6-34 `constants`
  size_t start_flags_col = 9 * 16, num_main_cols = start_flags_col + NUM_FLAGS_COLS, start_periodic_pulse_col = num_main_cols,
         start_io_pulses_col = start_periodic_pulse_col + 2, start_lookups_col, start_range_check_col = 0, num_range_check_cols = 9 * 16 - 1,
         end_range_check_col = num_range_check_cols, n_columns, n_public_inputs;
  explicit FqExpStark(size_t n) : num_io(n) {
    start_lookups_col = start_io_pulses_col + 1 + 4 * num_io;
    n_columns = start_lookups_col + 1 + 2 * num_range_check_cols;
    n_public_inputs = 4 * NUM_INPUT_LIMBS * num_io;
  }
  size_t num_columns() const override { return n_columns; }
  size_t num_public_inputs() const override { return n_public_inputs; }
  std::vector<std::pair<size_t, size_t>> permutation_pairs() const override { return u16_range_check_pairs(start_lookups_col, start_range_check_col, end_range_check_col); }
  // fq/exp.rs:128-143
  void generate_first_row(GF* lv, const U256& x, const U256& offset) const {
    Arr<GF, 16> a = i64_to_column_positive(fq_to_cols(fq_from_u256(x))), b = i64_to_column_positive(fq_to_cols(fq_from_u256(offset)));
    FqOutput<GF> out = lv[start_flags_col + 4] == GF(1) ? generate_fq_mul(a, b) : fq_output_default();
    size_t cur = 0; write_u256(lv, a, cur); write_u256(lv, b, cur); write_fq_output(lv, out, cur);
  }
  // fq/exp.rs:145-178
  void generate_next_row(const GF* lv, GF* nv) const {
    size_t is_sq_col = start_flags_col + 2, is_mul_col = start_flags_col + 4;
    size_t cur = 0;
    Arr<GF, 16> a = read_u256(lv, cur), b = read_u256(lv, cur);
    FqOutput<GF> output = read_fq_output(lv, cur);
    Arr<GF, 16> na = a, nb = b;
    if (lv[is_sq_col] == GF(1)) na = output.output; else if (lv[is_mul_col] == GF(1)) nb = output.output;
    FqOutput<GF> next = nv[is_sq_col] == GF(1) ? generate_fq_mul(na, na) : nv[is_mul_col] == GF(1) ? generate_fq_mul(na, nb) : fq_output_default();
    cur = 0; write_u256(nv, na, cur); write_u256(nv, nb, cur); write_fq_output(nv, next, cur);
  }
  // fq/exp.rs:215-247
  std::vector<std::vector<GF>> generate_trace_for_one_block(const U256& x, const U256& offset, const u32 exp_val[8], U256* result) const {
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
    size_t cur = 16;
    Arr<GF, 16> b = read_u256(rows.back().data(), cur);
    Fq out = cols_to_fq(b);
    // fq/exp.rs:241-244: output == offset * x^exp_val
    { Fq xf = fq_from_u256(x), acc = fq_from_u256(offset);
      for (int k = 0; k < 256; k++) { if ((exp_val[k >> 5] >> (k & 31)) & 1) acc = acc * xf; xf = xf * xf; }
      if (!(acc == out)) throw std::runtime_error("FqExp chain result differs from offset * x^e"); }
    *result = fq_to_u256(out);
    return rows;
  }
  // fq/exp.rs:249-278
  Cols generate_trace(const std::vector<FqExpIONative>& inputs, std::vector<U256>* results = nullptr) const {
    assert(inputs.size() == num_io);
    size_t nr = 2 * INPUT_LIMB_BITS * NUM_INPUT_LIMBS;
    std::vector<std::vector<GF>> rows(num_io * nr);
    std::vector<U256> res(num_io); std::string err;
#pragma omp parallel for schedule(dynamic)
    for (size_t k = 0; k < num_io; k++) {
      try {
        auto blk = generate_trace_for_one_block(inputs[k].x, inputs[k].offset, inputs[k].exp_val, &res[k]);
        for (size_t r = 0; r < nr; r++) rows[k * nr + r] = std::move(blk[r]);
      } catch (std::exception& e) {
#pragma omp critical
        err = e.what();
      }
    }
    if (!err.empty()) throw std::runtime_error(err);
    if (results) *results = res;
    Cols cols = transpose_rows(rows);
    rows.clear(); rows.shrink_to_fit();
    size_t rotation_period = 2 * INPUT_LIMB_BITS;
    generate_periodic_pulse_witness(cols, start_flags_col + 1, rotation_period, rotation_period - 2);
    generate_pulse(cols, G1ExpStark::get_pulse_positions(num_io));   // fq/exp.rs:180-190 (same positions)
    generate_u16_range_check(start_range_check_col, end_range_check_col, cols);
    return cols;
  }
  // fq/exp.rs:280-285 + :103-111
  std::vector<GF> generate_public_inputs(const std::vector<FqExpIONative>& inputs) const {
    std::vector<GF> pi;
    auto push = [&](const U256& v) { auto c = u256_to_u32_columns(v); pi.insert(pi.end(), c.begin(), c.end()); };
    for (auto& in : inputs) { push(in.x); push(in.offset); for (int i = 0; i < 8; i++) pi.push_back(GF(in.exp_val[i])); push(in.output); }
    return pi;
  }
  // fq/exp.rs:289-394
  template <class P> void eval_t(const P* lv, const P* nv, const P* pi, Consumer<P>& yc) const {
    P one = FieldOf<P>::c(1);
    size_t is_final_col = start_flags_col, is_sq_col = start_flags_col + 2, is_mul_col = start_flags_col + 4, start_limbs_col = start_flags_col + 6;
    size_t cur = 0;
    Arr<P, 16> a = read_u256(lv, cur), b = read_u256(lv, cur);
    FqOutput<P> output = read_fq_output(lv, cur);
    P is_mul = lv[is_mul_col], is_sq = lv[is_sq_col], is_final = lv[is_final_col], is_not_final = one - is_final;
    P sum_is_output = tzero<P>();
    for (size_t i = 1; i < 2 * num_io; i += 2) sum_is_output = sum_is_output + lv[get_pulse_col(start_io_pulses_col, i)];
    yc.constraint(is_final - sum_is_output);
    cur = 0;
    for (size_t i = 0; i < 2 * num_io; i += 2) {
      Arr<P, 8> io[4];   // x offset exp_val output
      for (int k = 0; k < 4; k++) { for (int j = 0; j < 8; j++) io[k][j] = pi[cur + j]; cur += 8; }
      P is_ith_input = lv[get_pulse_col(start_io_pulses_col, i)], is_ith_output = lv[get_pulse_col(start_io_pulses_col, i + 1)];
      Arr<P, 8> a32 = u16_columns_to_u32_columns(a), b32 = u16_columns_to_u32_columns(b);
      vec_equal(yc, is_ith_input, io[0], a32); vec_equal(yc, is_ith_input, io[1], b32); vec_equal(yc, is_ith_output, io[3], b32);
      Arr<P, 8> limbs; for (int j = 0; j < 8; j++) limbs[j] = lv[start_limbs_col + j];
      limbs[0] = limbs[0] * FieldOf<P>::c(2) + is_mul;
      vec_equal(yc, is_ith_input, io[2], limbs);
    }
    cur = 0;
    Arr<P, 16> next_a = read_u256(nv, cur), next_b = read_u256(nv, cur);
    fq_equal_transition(yc, is_not_final * is_sq, next_a, output.output); fq_equal_transition(yc, is_not_final * is_sq, next_b, b);
    fq_equal_transition(yc, is_not_final * is_mul, next_a, a); fq_equal_transition(yc, is_not_final * is_mul, next_b, output.output);
    { P f = is_not_final * (one - is_sq - is_mul); fq_equal_transition(yc, f, next_a, a); fq_equal_transition(yc, f, next_b, b); }
    eval_flags(yc, lv, nv, start_flags_col);
    eval_fq_mul(yc, is_sq, a, a, output);
    eval_fq_mul(yc, is_mul, a, b, output);
    eval_flags(yc, lv, nv, start_flags_col);   // emitted twice (fq/exp.rs:361 and :366-371)
    eval_periodic_pulse(yc, lv, nv, start_flags_col + 1, start_periodic_pulse_col, 2 * INPUT_LIMB_BITS, 2 * INPUT_LIMB_BITS - 2);
    eval_pulse(yc, lv, nv, start_io_pulses_col, G1ExpStark::get_pulse_positions(num_io));
    eval_u16_range_check(yc, lv, nv, start_lookups_col, num_range_check_cols);
  }
  ORC_AIR_EVAL_IMPL
};

// ---------------- Fq12 mul gadget: fq12/mul.rs ----------------
// fq12/mul.rs:24-87 `pol_mul_fq12`
template <class T> static inline Wide12<T> pol_mul_fq12(const Arr12<T>& a, const Arr12<T>& b, T xi) {
  std::array<Arr<T, 31>, 11> a0b0, a0b1, a1b0, a1b1;
  for (auto* v : {&a0b0, &a0b1, &a1b0, &a1b1}) for (auto& p : *v) p = pol_zero<T, 31>();
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
    pol_add_assign(a0b0[i + j], pol_mul_wide(a[i], b[j]));
    pol_add_assign(a0b1[i + j], pol_mul_wide(a[i], b[j + 6]));
    pol_add_assign(a1b0[i + j], pol_mul_wide(a[i + 6], b[j]));
    pol_add_assign(a1b1[i + j], pol_mul_wide(a[i + 6], b[j + 6]));
  }
  std::array<Arr<T, 31>, 11> re, im;   // a0b0_minus_a1b1, a0b1_plus_a1b0
  for (int i = 0; i < 11; i++) { re[i] = pol_sub_normal(a0b0[i], a1b1[i]); im[i] = pol_add_normal(a0b1[i], a1b0[i]); }
  Wide12<T> out;
  for (int i = 0; i < 6; i++) {
    if (i < 5) { Arr<T, 31> c = pol_add_normal(re[i], pol_mul_scalar(re[i + 6], xi)); pol_sub_assign(c, im[i + 6]); out[i] = c; }
    else out[i] = re[i];
  }
  for (int i = 0; i < 6; i++) {
    if (i < 5) { Arr<T, 31> c = pol_add_normal(im[i], re[i + 6]); pol_add_assign(c, pol_mul_scalar(im[i + 6], xi)); out[i + 6] = c; }
    else out[i + 6] = im[i];
  }
  return out;
}
// fq12/mul.rs:159-173
template <class T> static inline void write_fq12(T* lv, const Arr12<T>& v, size_t& cur) { for (int i = 0; i < 12; i++) write_u256(lv, v[i], cur); }
template <class T> static inline Arr12<T> read_fq12(const T* lv, size_t& cur) { Arr12<T> r; for (int i = 0; i < 12; i++) r[i] = read_u256(lv, cur); return r; }
// fq12/mul.rs:176-180
template <class P> struct Fq12Output { Arr12<P> output; ModulusAux<P> auxs[12]; P quot_signs[12]; };
// fq12/mul.rs:182-190
static inline Fq12Output<GF> fq12_output_default() {
  Fq12Output<GF> o;
  for (int i = 0; i < 12; i++) { o.output[i] = pol_zero<GF, 16>(); o.auxs[i] = modulus_aux_default(); o.quot_signs[i] = GF(1); }
  return o;
}
// fq12/mul.rs:192-215 `generate_fq12_mul`
static inline Fq12Output<GF> generate_fq12_mul(const Arr12<GF>& x, const Arr12<GF>& y) {
  Arr12<i64> xi, yi;
  for (int i = 0; i < 12; i++) { xi[i] = positive_column_to_i64(x[i]); yi[i] = positive_column_to_i64(y[i]); }
  Wide12<i64> pol_input = pol_mul_fq12(xi, yi, (i64)9);
  Fq12Output<GF> o;
  for (int i = 0; i < 12; i++) { ModOpWitness w = generate_modular_op(pol_input[i]); o.output[i] = w.output; o.auxs[i] = w.aux; o.quot_signs[i] = w.quot_sign; }
  return o;
}
// fq12/mul.rs:219-231
template <class T> static inline void write_fq12_output(T* lv, const Fq12Output<T>& o, size_t& cur) {
  write_fq12(lv, o.output, cur);
  for (int i = 0; i < 12; i++) write_modulus_aux(lv, o.auxs[i], cur);
  for (int i = 0; i < 12; i++) lv[cur++] = o.quot_signs[i];
}
// fq12/mul.rs:234-252
template <class T> static inline Fq12Output<T> read_fq12_output(const T* lv, size_t& cur) {
  Fq12Output<T> o;
  o.output = read_fq12(lv, cur);
  for (int i = 0; i < 12; i++) o.auxs[i] = read_modulus_aux(lv, cur);
  for (int i = 0; i < 12; i++) o.quot_signs[i] = lv[cur++];
  return o;
}
// fq12/mul.rs:254-275 `eval_fq12_mul`
template <class P> static inline void eval_fq12_mul(Consumer<P>& yc, P filter, const Arr12<P>& x, const Arr12<P>& y, const Fq12Output<P>& o) {
  Wide12<P> input = pol_mul_fq12(x, y, FieldOf<P>::c(9));
  Arr<P, 16> modulus = bn254_base_modulus_packfield<P>();
  for (int i = 0; i < 12; i++) eval_modular_op(yc, filter, modulus, input[i], o.output[i], o.quot_signs[i], o.auxs[i]);
}
// equals.rs:232-245
template <class P> static inline void fq12_equal_transition(Consumer<P>& yc, P filter, const Arr12<P>& x, const Arr12<P>& y) {
  for (int i = 0; i < 12; i++) for (int j = 0; j < 16; j++) yc.transition(filter * (x[i][j] - y[i][j]));
}

// plain Fq12 arithmetic in the flat basis, for the `output == offset * x^e` assertions (fq12/exp.rs:277-279)
struct Fq12Flat { Fq c[12]; bool operator==(const Fq12Flat& o) const { for (int i = 0; i < 12; i++) if (!(c[i] == o.c[i])) return false; return true; } };
static inline Fq12Flat fq12_flat_mul(const Fq12Flat& a, const Fq12Flat& b) {
  Fq re[11], im[11]; for (int i = 0; i < 11; i++) { re[i] = fq_zero(); im[i] = fq_zero(); }
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
    re[i + j] = re[i + j] + a.c[i] * b.c[j] - a.c[i + 6] * b.c[j + 6];
    im[i + j] = im[i + j] + a.c[i] * b.c[j + 6] + a.c[i + 6] * b.c[j];
  }
  Fq nine = fq_from_u64(9); Fq12Flat r;
  for (int i = 0; i < 6; i++) {
    if (i < 5) { r.c[i] = re[i] + nine * re[i + 6] - im[i + 6]; r.c[i + 6] = im[i] + re[i + 6] + nine * im[i + 6]; }
    else { r.c[i] = re[i]; r.c[i + 6] = im[i]; }
  }
  return r;
}
struct Fq12Words { U256 c[12]; };
static inline Fq12Flat fq12_from_words(const Fq12Words& w) { Fq12Flat r; for (int i = 0; i < 12; i++) r.c[i] = fq_from_u256(w.c[i]); return r; }
static inline Arr12<GF> fq12_words_to_cols(const Fq12Words& w) { Arr12<GF> r; for (int i = 0; i < 12; i++) r[i] = i64_to_column_positive(fq_to_cols(fq_from_u256(w.c[i]))); return r; }

// fq12/exp.rs:90-95 `Fq12ExpIONative` (coefficients in MyFq12 order)
struct Fq12ExpIONative { Fq12Words x, offset; u32 exp_val[8]; Fq12Words output; };

// Shared row logic of Fq12ExpStark / Fq12ExpU64Stark (fq12/exp.rs:142-214 == fq12_u64/exp_u64.rs:147-231 up to flag columns)
static inline void fq12_first_row(GF* lv, size_t is_mul_col, const Fq12Words& x, const Fq12Words& offset) {
  Arr12<GF> a = fq12_words_to_cols(x), b = fq12_words_to_cols(offset);
  Fq12Output<GF> out = lv[is_mul_col] == GF(1) ? generate_fq12_mul(a, b) : fq12_output_default();
  size_t cur = 0; write_fq12(lv, a, cur); write_fq12(lv, b, cur); write_fq12_output(lv, out, cur);
}
static inline void fq12_next_row(const GF* lv, GF* nv, size_t is_sq_col, size_t is_mul_col) {
  size_t cur = 0;
  Arr12<GF> a = read_fq12(lv, cur), b = read_fq12(lv, cur);
  Fq12Output<GF> output = read_fq12_output(lv, cur);
  Arr12<GF> na = a, nb = b;
  if (lv[is_sq_col] == GF(1)) na = output.output; else if (lv[is_mul_col] == GF(1)) nb = output.output;
  Fq12Output<GF> next = nv[is_sq_col] == GF(1) ? generate_fq12_mul(na, na) : nv[is_mul_col] == GF(1) ? generate_fq12_mul(na, nb) : fq12_output_default();
  cur = 0; write_fq12(nv, na, cur); write_fq12(nv, nb, cur); write_fq12_output(nv, next, cur);
}
static inline void fq12_check_result(const std::vector<GF>& last_row, const Fq12Words& x, const Fq12Words& offset, const u32* exp_bits_words, int nbits, Fq12Words* result) {
  size_t cur = 12 * 16;
  Arr12<GF> b = read_fq12(last_row.data(), cur);
  Fq12Flat out; for (int i = 0; i < 12; i++) { out.c[i] = cols_to_fq(b[i]); result->c[i] = fq_to_u256(out.c[i]); }
  Fq12Flat xf = fq12_from_words(x), acc = fq12_from_words(offset);
  for (int k = 0; k < nbits; k++) { if ((exp_bits_words[k >> 5] >> (k & 31)) & 1) acc = fq12_flat_mul(acc, xf); xf = fq12_flat_mul(xf, xf); }
  if (!(acc == out)) throw std::runtime_error("Fq12 chain result differs from offset * x^e");
}

struct Fq12ExpStark : Air {
  size_t num_io;
  // fq12/exp.rs:6-34 `constants`
  size_t start_flags_col = 108 * 16, num_main_cols = start_flags_col + NUM_FLAGS_COLS, start_periodic_pulse_col = num_main_cols,
         start_io_pulses_col = start_periodic_pulse_col + 2, start_lookups_col, start_range_check_col = 24 * 16, num_range_check_cols = 84 * 16 - 12,
         end_range_check_col = start_range_check_col + num_range_check_cols, n_columns, n_public_inputs;
  static const size_t IO_LEN = 36 * 16 + 8;  // fq12/exp.rs:97
  explicit Fq12ExpStark(size_t n) : num_io(n) {
    start_lookups_col = start_io_pulses_col + 1 + 4 * num_io;
    n_columns = start_lookups_col + 1 + 6 * num_range_check_cols;
    n_public_inputs = IO_LEN * num_io;
  }
  size_t num_columns() const override { return n_columns; }
  size_t num_public_inputs() const override { return n_public_inputs; }
  std::vector<std::pair<size_t, size_t>> permutation_pairs() const override { return split_u16_range_check_pairs(start_lookups_col, start_range_check_col, end_range_check_col); }
  // fq12/exp.rs:251-282
  std::vector<std::vector<GF>> generate_trace_for_one_block(const Fq12Words& x, const Fq12Words& offset, const u32 exp_val[8], Fq12Words* result) const {
    size_t num_rows = 2 * INPUT_LIMB_BITS * NUM_INPUT_LIMBS;
    std::vector<GF> lv(num_main_cols);
    generate_flags_first_row(lv.data(), start_flags_col, exp_val);
    fq12_first_row(lv.data(), start_flags_col + 4, x, offset);
    std::vector<std::vector<GF>> rows; rows.push_back(lv);
    for (size_t i = 0; i + 1 < num_rows; i++) {
      std::vector<GF> nv(lv.size());
      generate_flags_next_row(lv.data(), nv.data(), i, start_flags_col);
      fq12_next_row(lv.data(), nv.data(), start_flags_col + 2, start_flags_col + 4);
      rows.push_back(nv); lv = nv;
    }
    fq12_check_result(rows.back(), x, offset, exp_val, 256, result);
    return rows;
  }
  // fq12/exp.rs:284-313
  Cols generate_trace(const std::vector<Fq12ExpIONative>& inputs, std::vector<Fq12Words>* results = nullptr) const {
    assert(inputs.size() == num_io);
    size_t nr = 2 * INPUT_LIMB_BITS * NUM_INPUT_LIMBS;
    std::vector<std::vector<GF>> rows(num_io * nr);
    std::vector<Fq12Words> res(num_io); std::string err;
#pragma omp parallel for schedule(dynamic)
    for (size_t k = 0; k < num_io; k++) {
      try {
        auto blk = generate_trace_for_one_block(inputs[k].x, inputs[k].offset, inputs[k].exp_val, &res[k]);
        for (size_t r = 0; r < nr; r++) rows[k * nr + r] = std::move(blk[r]);
      } catch (std::exception& e) {
#pragma omp critical
        err = e.what();
      }
    }
    if (!err.empty()) throw std::runtime_error(err);
    if (results) *results = res;
    Cols cols = transpose_rows(rows);
    rows.clear(); rows.shrink_to_fit();
    size_t rotation_period = 2 * INPUT_LIMB_BITS;
    generate_periodic_pulse_witness(cols, start_flags_col + 1, rotation_period, rotation_period - 2);
    generate_pulse(cols, G1ExpStark::get_pulse_positions(num_io));   // fq12/exp.rs:216-227 (same positions)
    generate_split_u16_range_check(start_range_check_col, end_range_check_col, cols);
    return cols;
  }
  // fq12/exp.rs:315-320 + :107-124 (u16 limbs for the Fq12 coefficients)
  std::vector<GF> generate_public_inputs(const std::vector<Fq12ExpIONative>& inputs) const {
    std::vector<GF> pi;
    auto push12 = [&](const Fq12Words& w) { for (int i = 0; i < 12; i++) for (int j = 0; j < 16; j++) pi.push_back(GF((w.c[i].w[j / 4] >> (16 * (j % 4))) & 0xFFFF)); };
    for (auto& in : inputs) { push12(in.x); push12(in.offset); for (int i = 0; i < 8; i++) pi.push_back(GF(in.exp_val[i])); push12(in.output); }
    return pi;
  }
  // fq12/exp.rs:324-428
  template <class P> void eval_t(const P* lv, const P* nv, const P* pi, Consumer<P>& yc) const {
    P one = FieldOf<P>::c(1);
    size_t is_final_col = start_flags_col, is_sq_col = start_flags_col + 2, is_mul_col = start_flags_col + 4, start_limbs_col = start_flags_col + 6;
    size_t cur = 0;
    Arr12<P> a = read_fq12(lv, cur), b = read_fq12(lv, cur);
    Fq12Output<P> output = read_fq12_output(lv, cur);
    P is_mul = lv[is_mul_col], is_sq = lv[is_sq_col], is_final = lv[is_final_col], is_not_final = one - is_final;
    P sum_is_output = tzero<P>();
    for (size_t i = 1; i < 2 * num_io; i += 2) sum_is_output = sum_is_output + lv[get_pulse_col(start_io_pulses_col, i)];
    yc.constraint(is_final - sum_is_output);
    cur = 0;
    for (size_t i = 0; i < 2 * num_io; i += 2) {
      const P* x = pi + cur; const P* off = x + 192; const P* ev = off + 192; const P* outp = ev + 8; cur += IO_LEN;   // fq12/exp.rs:126-140
      P is_ith_input = lv[get_pulse_col(start_io_pulses_col, i)], is_ith_output = lv[get_pulse_col(start_io_pulses_col, i + 1)];
      for (int k = 0; k < 12; k++) {
        for (int j = 0; j < 16; j++) yc.constraint(is_ith_input * (x[16 * k + j] - a[k][j]));
        for (int j = 0; j < 16; j++) yc.constraint(is_ith_input * (off[16 * k + j] - b[k][j]));
        for (int j = 0; j < 16; j++) yc.constraint(is_ith_output * (outp[16 * k + j] - b[k][j]));
      }
      Arr<P, 8> limbs, evv; for (int j = 0; j < 8; j++) { limbs[j] = lv[start_limbs_col + j]; evv[j] = ev[j]; }
      limbs[0] = limbs[0] * FieldOf<P>::c(2) + is_mul;
      vec_equal(yc, is_ith_input, evv, limbs);
    }
    cur = 0;
    Arr12<P> next_a = read_fq12(nv, cur), next_b = read_fq12(nv, cur);
    fq12_equal_transition(yc, is_not_final * is_sq, next_a, output.output); fq12_equal_transition(yc, is_not_final * is_sq, next_b, b);
    fq12_equal_transition(yc, is_not_final * is_mul, next_a, a); fq12_equal_transition(yc, is_not_final * is_mul, next_b, output.output);
    { P f = is_not_final * (one - is_sq - is_mul); fq12_equal_transition(yc, f, next_a, a); fq12_equal_transition(yc, f, next_b, b); }
    eval_flags(yc, lv, nv, start_flags_col);
    eval_fq12_mul(yc, is_sq, a, a, output);
    eval_fq12_mul(yc, is_mul, a, b, output);
    eval_flags(yc, lv, nv, start_flags_col);   // emitted twice (fq12/exp.rs:395 and :400-405)
    eval_periodic_pulse(yc, lv, nv, start_flags_col + 1, start_periodic_pulse_col, 2 * INPUT_LIMB_BITS, 2 * INPUT_LIMB_BITS - 2);
    eval_pulse(yc, lv, nv, start_io_pulses_col, G1ExpStark::get_pulse_positions(num_io));
    eval_split_u16_range_check(yc, lv, nv, start_lookups_col, start_range_check_col, end_range_check_col);
  }
  ORC_AIR_EVAL_IMPL
};

// ---------------- 64-bit exponent variant: fq12_u64/flags_u64.rs, exp_u64.rs ----------------
static const int NUM_FLAGS_U64_COLS = 6;   // flags_u64.rs:21
// flags_u64.rs:34-56: is_final, a, b, filtered_bit, bit, val
static inline void generate_flags_u64_first_row(GF* lv, size_t sf, u64 exp_val) {
  u64 first_bit = exp_val % 2, rest = (exp_val - first_bit) / 2;
  lv[sf] = GF(); lv[sf + 1] = GF(); lv[sf + 2] = GF(1); lv[sf + 3] = GF(first_bit); lv[sf + 4] = GF(first_bit); lv[sf + 5] = GF(rest);
}
// flags_u64.rs:58-94
static inline void generate_flags_u64_next_row(const GF* lv, GF* nv, size_t cur_row, size_t sf) {
  size_t is_final = sf, a = sf + 1, b = sf + 2, fbit = sf + 3, bit = sf + 4, val = sf + 5;
  nv[a] = GF(1) - lv[a]; nv[b] = GF(1) - lv[b];
  nv[is_final] = cur_row == 2 * 64 - 2 ? GF(1) : GF();
  if (lv[a] == GF(1)) { u64 fl = lv[val].v, nb = fl % 2; nv[bit] = GF(nb); nv[val] = GF((fl - nb) / 2); }
  else { nv[bit] = lv[bit]; nv[val] = lv[val]; }
  nv[fbit] = nv[bit] * nv[b];
}
// flags_u64.rs:96-139
template <class P> static inline void eval_flags_u64(Consumer<P>& yc, const P* lv, const P* nv, size_t sf) {
  size_t is_final_c = sf, a = sf + 1, b = sf + 2, fbit = sf + 3, bit_c = sf + 4, val = sf + 5;
  P one = FieldOf<P>::c(1);
  yc.first_row(lv[a]);
  yc.first_row(lv[b] - one);
  P bit = lv[bit_c];
  yc.constraint(bit * bit - bit);
  yc.constraint(bit * lv[b] - lv[fbit]);
  yc.transition(lv[a] + nv[a] - one);
  yc.transition(lv[b] + nv[b] - one);
  P first_limb = lv[val], next_first_limb = nv[val], next_bit = nv[bit_c], is_split = lv[a], is_final = lv[is_final_c];
  P is_not_final = one - is_final;
  yc.transition(is_not_final * is_split * (first_limb - FieldOf<P>::c(2) * next_first_limb - next_bit));
  P is_not_split = one - is_split;
  yc.transition(is_not_split * (next_bit - bit));
  yc.transition(is_not_final * is_not_split * (first_limb - next_first_limb));
}
// fq12_u64/exp_u64.rs:85-90 `Fq12ExpU64IONative`
struct Fq12ExpU64IONative { Fq12Words x, offset; u64 exp_val; Fq12Words output; };

struct Fq12ExpU64Stark : Air {
  size_t num_io;
  static const size_t ROWS_PER_IO = 2 * 64;
  // fq12_u64/exp_u64.rs:19-45 `constants`
  size_t start_flags_col = 108 * 16, num_main_cols = start_flags_col + NUM_FLAGS_U64_COLS, start_io_pulses_col = num_main_cols, start_lookups_col,
         start_range_check_col = 24 * 16, num_range_check_cols = 84 * 16 - 12, end_range_check_col = start_range_check_col + num_range_check_cols, n_columns, n_public_inputs;
  static const size_t IO_LEN = 36 * 16 + 1;  // exp_u64.rs:92
  explicit Fq12ExpU64Stark(size_t n) : num_io(n) {
    start_lookups_col = start_io_pulses_col + 1 + 4 * num_io;
    n_columns = start_lookups_col + 1 + 6 * num_range_check_cols;
    n_public_inputs = IO_LEN * num_io;
  }
  size_t num_columns() const override { return n_columns; }
  size_t num_public_inputs() const override { return n_public_inputs; }
  std::vector<std::pair<size_t, size_t>> permutation_pairs() const override { return split_u16_range_check_pairs(start_lookups_col, start_range_check_col, end_range_check_col); }
  // exp_u64.rs:234-245
  static std::vector<size_t> get_pulse_u64_positions(size_t num_io) {
    std::vector<size_t> p; for (size_t i = 0; i < num_io; i++) { p.push_back(i * ROWS_PER_IO); p.push_back(i * ROWS_PER_IO + ROWS_PER_IO - 1); } return p;
  }
  // exp_u64.rs:252-279
  std::vector<std::vector<GF>> generate_trace_for_one_block(const Fq12Words& x, const Fq12Words& offset, u64 exp_val, Fq12Words* result) const {
    std::vector<GF> lv(num_main_cols);
    generate_flags_u64_first_row(lv.data(), start_flags_col, exp_val);
    fq12_first_row(lv.data(), start_flags_col + 3, x, offset);
    std::vector<std::vector<GF>> rows; rows.push_back(lv);
    for (size_t i = 0; i + 1 < ROWS_PER_IO; i++) {
      std::vector<GF> nv(lv.size());
      generate_flags_u64_next_row(lv.data(), nv.data(), i, start_flags_col);
      fq12_next_row(lv.data(), nv.data(), start_flags_col + 1, start_flags_col + 3);
      rows.push_back(nv); lv = nv;
    }
    u32 ew[2] = {(u32)exp_val, (u32)(exp_val >> 32)};
    fq12_check_result(rows.back(), x, offset, ew, 64, result);
    return rows;
  }
  // exp_u64.rs:280-305
  Cols generate_trace(const std::vector<Fq12ExpU64IONative>& inputs, std::vector<Fq12Words>* results = nullptr) const {
    assert(inputs.size() == num_io);
    std::vector<std::vector<GF>> rows(num_io * ROWS_PER_IO);
    std::vector<Fq12Words> res(num_io); std::string err;
#pragma omp parallel for schedule(dynamic)
    for (size_t k = 0; k < num_io; k++) {
      try {
        auto blk = generate_trace_for_one_block(inputs[k].x, inputs[k].offset, inputs[k].exp_val, &res[k]);
        for (size_t r = 0; r < ROWS_PER_IO; r++) rows[k * ROWS_PER_IO + r] = std::move(blk[r]);
      } catch (std::exception& e) {
#pragma omp critical
        err = e.what();
      }
    }
    if (!err.empty()) throw std::runtime_error(err);
    if (results) *results = res;
    Cols cols = transpose_rows(rows);
    rows.clear(); rows.shrink_to_fit();
    generate_pulse(cols, get_pulse_u64_positions(num_io));
    generate_split_u16_range_check(start_range_check_col, end_range_check_col, cols);
    return cols;
  }
  // exp_u64.rs:307-312 + :102-122
  std::vector<GF> generate_public_inputs(const std::vector<Fq12ExpU64IONative>& inputs) const {
    std::vector<GF> pi;
    auto push12 = [&](const Fq12Words& w) { for (int i = 0; i < 12; i++) for (int j = 0; j < 16; j++) pi.push_back(GF((w.c[i].w[j / 4] >> (16 * (j % 4))) & 0xFFFF)); };
    for (auto& in : inputs) { push12(in.x); push12(in.offset); if (in.exp_val >= GP) throw std::runtime_error("exp_val is not a canonical field element"); pi.push_back(GF(in.exp_val)); push12(in.output); }
    return pi;
  }
  // exp_u64.rs:315-409
  template <class P> void eval_t(const P* lv, const P* nv, const P* pi, Consumer<P>& yc) const {
    P one = FieldOf<P>::c(1);
    size_t is_final_col = start_flags_col, is_sq_col = start_flags_col + 1, is_mul_col = start_flags_col + 3, exp_val_col = start_flags_col + 5;
    size_t cur = 0;
    Arr12<P> a = read_fq12(lv, cur), b = read_fq12(lv, cur);
    Fq12Output<P> output = read_fq12_output(lv, cur);
    P is_mul = lv[is_mul_col], is_sq = lv[is_sq_col], is_final = lv[is_final_col], is_not_final = one - is_final;
    P sum_is_output = tzero<P>();
    for (size_t i = 1; i < 2 * num_io; i += 2) sum_is_output = sum_is_output + lv[get_pulse_col(start_io_pulses_col, i)];
    yc.constraint(is_final - sum_is_output);
    cur = 0;
    for (size_t i = 0; i < 2 * num_io; i += 2) {
      const P* x = pi + cur; const P* off = x + 192; P ev = off[192]; const P* outp = off + 193; cur += IO_LEN;   // exp_u64.rs:124-145
      P is_ith_input = lv[get_pulse_col(start_io_pulses_col, i)], is_ith_output = lv[get_pulse_col(start_io_pulses_col, i + 1)];
      for (int k = 0; k < 12; k++) {
        for (int j = 0; j < 16; j++) yc.constraint(is_ith_input * (x[16 * k + j] - a[k][j]));
        for (int j = 0; j < 16; j++) yc.constraint(is_ith_input * (off[16 * k + j] - b[k][j]));
        for (int j = 0; j < 16; j++) yc.constraint(is_ith_output * (outp[16 * k + j] - b[k][j]));
      }
      P recovered = lv[exp_val_col] * FieldOf<P>::c(2) + is_mul;
      yc.constraint(is_ith_input * (ev - recovered));
    }
    cur = 0;
    Arr12<P> next_a = read_fq12(nv, cur), next_b = read_fq12(nv, cur);
    fq12_equal_transition(yc, is_not_final * is_sq, next_a, output.output); fq12_equal_transition(yc, is_not_final * is_sq, next_b, b);
    fq12_equal_transition(yc, is_not_final * is_mul, next_a, a); fq12_equal_transition(yc, is_not_final * is_mul, next_b, output.output);
    { P f = is_not_final * (one - is_sq - is_mul); fq12_equal_transition(yc, f, next_a, a); fq12_equal_transition(yc, f, next_b, b); }
    eval_flags_u64(yc, lv, nv, start_flags_col);
    eval_fq12_mul(yc, is_sq, a, a, output);
    eval_fq12_mul(yc, is_mul, a, b, output);
    eval_flags_u64(yc, lv, nv, start_flags_col);   // emitted twice (exp_u64.rs:385 and :390-395)
    eval_pulse(yc, lv, nv, start_io_pulses_col, get_pulse_u64_positions(num_io));
    eval_split_u16_range_check(yc, lv, nv, start_lookups_col, start_range_check_col, end_range_check_col);
  }
  ORC_AIR_EVAL_IMPL
};
}  // namespace orc
