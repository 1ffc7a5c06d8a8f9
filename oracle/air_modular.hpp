// ORACLE (test infrastructure).  The reference's test-only `ModularStark`
// (src/modular/modular.rs:361-537), with the row count a parameter instead of the fixed 512
// (BASELINE.json config 5 sweeps 2^16..2^22 rows).
#pragma once
#include "air_common.hpp"
namespace orc {
struct ModularStark : Air {
  static const size_t MAIN_COLS = 9 * 16 + 1, START_RANGE_CHECK = 32, NUM_RANGE_CHECK = 7 * 16 - 1, END_RANGE_CHECK = START_RANGE_CHECK + NUM_RANGE_CHECK;
  static const size_t COLUMNS = MAIN_COLS + 1 + 6 * NUM_RANGE_CHECK;  // 812
  size_t num_columns() const override { return COLUMNS; }
  size_t num_public_inputs() const override { return 0; }
  std::vector<std::pair<size_t, size_t>> permutation_pairs() const override { return split_u16_range_check_pairs(MAIN_COLS, START_RANGE_CHECK, END_RANGE_CHECK); }
  // modular.rs:379-434; inputs[r] = (input0, input1) canonical residues.
  Cols generate_trace(const std::vector<std::array<U256, 2>>& inputs) const {
    std::vector<std::vector<GF>> rows(inputs.size(), std::vector<GF>(MAIN_COLS));
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < inputs.size(); r++) {
      Fq a = fq_from_u256(inputs[r][0]), b = fq_from_u256(inputs[r][1]);
      Fq out_fq = a * b;
      Arr<i64, 16> l0 = fq_to_cols(a), l1 = fq_to_cols(b);
      Arr<i64, 31> pol_input = pol_mul_wide(l0, l1);
      ModOpWitness w = generate_modular_op(pol_input);
      assert(cols_to_fq(w.output) == out_fq);   // modular.rs:405-406
      GF* lv = rows[r].data(); size_t cur = 0;
      write_u256(lv, i64_to_column_positive(l0), cur); write_u256(lv, i64_to_column_positive(l1), cur);
      write_u256(lv, w.output, cur); write_modulus_aux(lv, w.aux, cur);
      lv[cur++] = w.quot_sign; lv[cur++] = GF(1);
      assert(cur == MAIN_COLS);
    }
    Cols cols = transpose_rows(rows);
    generate_split_u16_range_check(START_RANGE_CHECK, END_RANGE_CHECK, cols);
    return cols;
  }
  // modular.rs:440-482
  template <class P> void eval_t(const P* lv, const P* nv, const P*, Consumer<P>& yc) const {
    eval_split_u16_range_check(yc, lv, nv, MAIN_COLS, START_RANGE_CHECK, END_RANGE_CHECK);
    size_t cur = 0;
    Arr<P, 16> in0 = read_u256(lv, cur), in1 = read_u256(lv, cur), output = read_u256(lv, cur);
    ModulusAux<P> aux = read_modulus_aux(lv, cur);
    P quot_sign = lv[cur++], filter = lv[cur++];
    assert(cur == MAIN_COLS);
    eval_modular_op(yc, filter, bn254_base_modulus_packfield<P>(), pol_mul_wide(in0, in1), output, quot_sign, aux);
  }
  ORC_AIR_EVAL_IMPL
};
}  // namespace orc
