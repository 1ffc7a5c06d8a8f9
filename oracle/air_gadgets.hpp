// ORACLE (test infrastructure).  The reference's test-only gadget AIRs, with the row count a parameter instead of the fixed 512
// (extra sweep points for BASELINE.json config 5 and the split range check at N >> 256):
//   G1Stark   src/curves/g1/muladd.rs:462-624  -- every row an independent G1 addition a + b (is_add = 1, is_double = 0)
//   Fq12Stark src/fields/fq12/mul.rs:355-484   -- every row an independent Fq12 product x * y
// The reference draws the row inputs from rand::thread_rng(); here they are an argument (one record per row).
#pragma once
#include "air_g1.hpp"
#include "air_fq12.hpp"
namespace orc {
struct G1Stark : Air {
  static const size_t MAIN_COLS = 24 * 16 + 2, START_RANGE_CHECK = 4 * 16, NUM_RANGE_CHECKS = 20 * 16 - 4, END_RANGE_CHECK = START_RANGE_CHECK + NUM_RANGE_CHECKS;
  static const size_t COLUMNS = MAIN_COLS + 1 + 6 * NUM_RANGE_CHECKS;   // 2283
  size_t num_columns() const override { return COLUMNS; }
  size_t num_public_inputs() const override { return 0; }
  std::vector<std::pair<size_t, size_t>> permutation_pairs() const override { return split_u16_range_check_pairs(MAIN_COLS, START_RANGE_CHECK, END_RANGE_CHECK); }
  // muladd.rs:481-546; inputs[r] = (a, b) affine points with distinct x (the addition gadget divides by b.x - a.x)
  Cols generate_trace(const std::vector<std::array<G1Point, 2>>& inputs) const {
    std::vector<std::vector<GF>> rows(inputs.size(), std::vector<GF>(MAIN_COLS));
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < inputs.size(); r++) {
      auto cols_of = [](const U256& v) { return i64_to_column_positive(fq_to_cols(fq_from_u256(v))); };
      Arr<GF, 16> a_x = cols_of(inputs[r][0].x), a_y = cols_of(inputs[r][0].y), b_x = cols_of(inputs[r][1].x), b_y = cols_of(inputs[r][1].y);
      G1Output<GF> out = generate_g1_add(a_x, a_y, b_x, b_y);
      GF* lv = rows[r].data(); size_t cur = 0;
      write_u256(lv, a_x, cur); write_u256(lv, a_y, cur); write_u256(lv, b_x, cur); write_u256(lv, b_y, cur);
      write_g1_output(lv, out, cur);
      lv[cur++] = GF(1);   // is_add
      lv[cur++] = GF(0);   // is_double
      assert(cur == MAIN_COLS);
    }
    Cols cols = transpose_rows(rows);
    generate_split_u16_range_check(START_RANGE_CHECK, END_RANGE_CHECK, cols);
    return cols;
  }
  // muladd.rs:549-581
  template <class P> void eval_t(const P* lv, const P* nv, const P*, Consumer<P>& yc) const {
    eval_split_u16_range_check(yc, lv, nv, MAIN_COLS, START_RANGE_CHECK, END_RANGE_CHECK);
    size_t cur = 0;
    Arr<P, 16> a_x = read_u256(lv, cur), a_y = read_u256(lv, cur), b_x = read_u256(lv, cur), b_y = read_u256(lv, cur);
    G1Output<P> output = read_g1_output(lv, cur);
    P is_add = lv[cur++], is_double = lv[cur++];
    assert(cur == MAIN_COLS);
    eval_g1_add(yc, is_add, a_x, a_y, b_x, b_y, output);
    eval_g1_double(yc, is_double, a_x, a_y, output);
  }
  ORC_AIR_EVAL_IMPL
};

struct Fq12Stark : Air {
  static const size_t MAIN_COLS = 108 * 16 + 1, START_RANGE_CHECK = 24 * 16, NUM_RANGE_CHECK = 84 * 16 - 12, END_RANGE_CHECK = START_RANGE_CHECK + NUM_RANGE_CHECK;
  static const size_t COLUMNS = MAIN_COLS + 1 + 6 * NUM_RANGE_CHECK;   // 9722
  size_t num_columns() const override { return COLUMNS; }
  size_t num_public_inputs() const override { return 0; }
  std::vector<std::pair<size_t, size_t>> permutation_pairs() const override { return split_u16_range_check_pairs(MAIN_COLS, START_RANGE_CHECK, END_RANGE_CHECK); }
  // mul.rs:378-419; inputs[r] = (x, y), 12 canonical coefficients each (MyFq12 order)
  Cols generate_trace(const std::vector<std::array<Fq12Words, 2>>& inputs) const {
    std::vector<std::vector<GF>> rows(inputs.size(), std::vector<GF>(MAIN_COLS));
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < inputs.size(); r++) {
      Arr12<GF> x = fq12_words_to_cols(inputs[r][0]), y = fq12_words_to_cols(inputs[r][1]);
      Fq12Output<GF> out = generate_fq12_mul(x, y);
      // mul.rs:383-389: the witness must equal the field product
      Fq12Flat want = fq12_flat_mul(fq12_from_words(inputs[r][0]), fq12_from_words(inputs[r][1]));
      for (int i = 0; i < 12; i++) { assert(cols_to_fq(out.output[i]) == want.c[i]); (void)want; }
      GF* lv = rows[r].data(); size_t cur = 0;
      write_fq12(lv, x, cur); write_fq12(lv, y, cur);
      write_fq12_output(lv, out, cur);
      lv[cur++] = GF(1);   // filter
      assert(cur == MAIN_COLS);
    }
    Cols cols = transpose_rows(rows);
    generate_split_u16_range_check(START_RANGE_CHECK, END_RANGE_CHECK, cols);
    return cols;
  }
  // mul.rs:423-448
  template <class P> void eval_t(const P* lv, const P* nv, const P*, Consumer<P>& yc) const {
    eval_split_u16_range_check(yc, lv, nv, MAIN_COLS, START_RANGE_CHECK, END_RANGE_CHECK);
    size_t cur = 0;
    Arr12<P> x = read_fq12(lv, cur), y = read_fq12(lv, cur);
    Fq12Output<P> output = read_fq12_output(lv, cur);
    P filter = lv[cur++];
    assert(cur == MAIN_COLS);
    eval_fq12_mul(yc, filter, x, y, output);
  }
  ORC_AIR_EVAL_IMPL
};
}  // namespace orc
