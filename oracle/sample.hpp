// ORACLE (test infrastructure): bounded-sample timing of the CPU prover for bench.py's cpu_baseline and
// `--impl reference` legs.  A full G1 proof cannot be shrunk (the u16 lookup table forces 2^16 rows,
// reference src/utils/range_check.rs:26), so the sample runs every heavy phase of trace generation and
// `prove` on 1/2^shift of its columns / instances / evaluation points -- all of these phases are linear in
// that dimension -- and the caller scales the time back by 2^shift.  Small phases (FRI layers, PoW,
// queries) run in full.  The sample does not produce a proof; orc_prove does, and is what parity uses.
#pragma once
#include "stark.hpp"
#include "air_g1.hpp"
#include "air_modular.hpp"
#include <random>

namespace orc {
struct SampleTimes { double tracegen_ms = 0, commit_ms = 0, zpoly_ms = 0, quotient_ms = 0, openings_ms = 0, reduce_ms = 0, fri_ms = 0; };

static inline std::vector<std::vector<GF>> random_cols(size_t ncols, size_t n, u64 seed, u64 bound) {
  std::vector<std::vector<GF>> c(ncols, std::vector<GF>(n));
  std::mt19937_64 rng(seed);
  for (auto& col : c) for (auto& v : col) v = GF(bound ? rng() % bound : rng() % GP);
  return c;
}

// Phases of `prove` for an AIR with `air.num_columns()` columns and nrows rows, on 1/2^shift of the work.
static inline SampleTimes time_prove_sample(const Air& air, size_t nrows, const StarkConfig& cfg, int shift) {
  SampleTimes t; Timer tm;
  const size_t C = air.num_columns(), Cs = std::max<size_t>(1, C >> shift);
  auto pairs = air.permutation_pairs();
  const size_t Z = pairs.size(), Zs = std::max<size_t>(1, Z >> shift);
  const int degree_bits = log2_strict(nrows);
  // trace commitment on Cs columns (iFFT, LDE, transpose, Merkle)
  auto cols = random_cols(Cs, nrows, 1, 0);
  tm.lap();
  PolynomialBatch tc = PolynomialBatch::from_values(cols, cfg.rate_bits, cfg.cap_height);
  t.commit_ms += tm.lap();
  // Z polynomials for Zs batches + their commitment
  if (Z) {
    PermChallengeSets sets(cfg.num_challenges, std::vector<PermChallenge>(air.quotient_degree_factor(), PermChallenge{GF(3), GF(5)}));
    std::vector<std::vector<GF>> zs(Zs);
    tm.lap();
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t b = 0; b < Zs; b++) {
      const auto& l = cols[b % Cs]; const auto& r = cols[(b + 1) % Cs];
      size_t n = nrows;
      std::vector<GF> num(n), den(n), pre(n);
      for (size_t i = 0; i < n; i++) { num[i] = (GF(5) + l[i]) * (GF(7) + l[i]); den[i] = (GF(5) + r[i]) * (GF(7) + r[i]); }
      GF acc = GF::one();
      for (size_t i = 0; i < n; i++) { pre[i] = acc; acc = acc * den[i]; }
      GF inv = gl_inv(acc);
      for (size_t i = n; i-- > 0;) { GF di = inv * pre[i]; inv = inv * den[i]; num[i] = num[i] * di; }
      std::vector<GF> z(n); acc = GF::one();
      for (size_t i = 0; i < n; i++) { z[i] = acc; acc = acc * num[i]; }
      zs[b] = std::move(z);
    }
    t.zpoly_ms += tm.lap();
    PolynomialBatch zc = PolynomialBatch::from_values(zs, cfg.rate_bits, cfg.cap_height);
    t.commit_ms += tm.lap();
  }
  // quotient: every constraint at (2N >> shift) points (row values are arbitrary field elements on the LDE)
  {
    size_t npts = std::max<size_t>(64, (nrows * 2) >> shift);
    std::vector<GF> alphas(cfg.num_challenges, GF(1234567));
    std::vector<GF> pis(air.num_public_inputs(), GF(9));
    PermChallengeSets sets(cfg.num_challenges, std::vector<PermChallenge>(air.quotient_degree_factor(), PermChallenge{GF(3), GF(5)}));
    PermBatches pb; if (Z) pb = get_permutation_batches(pairs, sets, cfg.num_challenges, air.quotient_degree_factor());
    auto rows = random_cols(64, C, 2, 0);       // 64 distinct rows reused round-robin
    auto zrows = random_cols(64, pb.size() ? pb.size() : 1, 3, 0);
    std::vector<GF> out(npts);
    tm.lap();
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < npts; i++) {
      Consumer<GF> yc(alphas, GF(i + 2), GF(i + 3), GF(i + 4));
      const auto& lv = rows[i % 64]; const auto& nv = rows[(i + 1) % 64];
      air.eval(lv.data(), nv.data(), pis.data(), yc);
      if (Z) eval_permutation_checks<GF>(pb, lv.data(), zrows[i % 64].data(), zrows[(i + 1) % 64].data(), pb.size(), yc);
      out[i] = yc.accs[0];
    }
    t.quotient_ms += tm.lap();
    volatile u64 sink = out[npts / 2].v; (void)sink;
  }
  // openings (two points) and FRI batch reduction over Cs (+Zs) coefficient columns
  {
    GF2 zeta(GF(123), GF(456));
    size_t ncol = Cs + (Z ? Zs : 0);
    tm.lap();
    std::vector<GF2> r(2 * ncol);
#pragma omp parallel for schedule(dynamic, 8)
    for (size_t i = 0; i < ncol; i++) { const auto& p = tc.polynomials[i % Cs]; r[2 * i] = eval_poly_ext(p, zeta); r[2 * i + 1] = eval_poly_ext(p, zeta * GF(3)); }
    t.openings_ms += tm.lap();
    GF2 alpha(GF(77), GF(88));
    std::vector<GF2> apow(ncol); { GF2 a = GF2::one(); for (auto& x : apow) { x = a; a = a * alpha; } }
    std::vector<GF2> comp(nrows);
    for (int batch = 0; batch < 2; batch++) {
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < nrows; i++) { GF2 acc; for (size_t j = 0; j < ncol; j++) acc = acc + apow[j] * tc.polynomials[j % Cs][i]; comp[i] = acc; }
    }
    t.reduce_ms += tm.lap();
    // full-size (not sampled) tail: final LDE, commit-phase layers, proof of work
    std::vector<GF2> coeffs = comp; coeffs.resize(nrows << cfg.rate_bits);
    std::vector<GF2> values = coset_fft_ext(coeffs, coset_shift());
    GF shift_ = coset_shift();
    Challenger ch;
    for (int ab : cfg.reduction_arity_bits(degree_bits)) {
      size_t arity = size_t(1) << ab;
      reverse_index_bits_in_place(values);
      std::vector<std::vector<GF>> chunked(values.size() / arity);
      for (size_t i = 0; i < chunked.size(); i++) chunked[i] = flatten_ext(&values[i * arity], arity);
      MerkleTree tree(std::move(chunked), cfg.cap_height);
      ch.observe_cap(tree.cap);
      GF2 beta = ch.get_ext();
      std::vector<GF2> folded(coeffs.size() / arity);
      for (size_t i = 0; i < folded.size(); i++) { GF2 acc; for (size_t k = arity; k-- > 0;) acc = acc * beta + coeffs[i * arity + k]; folded[i] = acc; }
      coeffs.swap(folded);
      shift_ = gl_pow(shift_, arity);
      values = coset_fft_ext(coeffs, shift_);
    }
    fri_proof_of_work(ch, cfg);
    t.fri_ms += tm.lap();
  }
  return t;
}

// Trace generation of G1ExpStark on 1/2^shift of its instances, pulse positions and lookup columns.
static inline double time_g1_tracegen_sample(const G1ExpStark& air, const std::vector<G1ExpIONative>& ios, int shift) {
  Timer tm;
  size_t n = std::max<size_t>(1, ios.size() >> shift), N = air.num_io * 512;
  std::vector<G1Point> res(n);
  tm.lap();
  for (size_t k = 0; k < n; k++) air.generate_trace_for_one_block(ios[k].x, ios[k].offset, ios[k].exp_val, &res[k]);  // serial, like the reference (g1/exp.rs:293-297)
  double ms = tm.lap();
  Cols cols = random_cols(std::max<size_t>(1, air.num_range_check_cols >> shift), N, 4, 65536);
  size_t nc = cols.size();
  std::vector<size_t> positions;
  for (size_t i = 0; i < std::max<size_t>(1, (2 * air.num_io) >> shift); i++) positions.push_back(i * 512 % N);
  tm.lap();
  { Cols c2(1, std::vector<GF>(N)); generate_pulse(c2, positions); }
  generate_u16_range_check(0, nc, cols);
  ms += tm.lap();
  return ms;
}
// Same for the other exponentiation AIRs: `block(k)` generates instance k's rows (serial loop, like the reference).
template <class BlockFn> static inline double time_exp_tracegen_sample(BlockFn block, size_t num_io, size_t rows_per_io, size_t nrc, bool split, int shift) {
  Timer tm;
  size_t n = std::max<size_t>(1, num_io >> shift), N = num_io * rows_per_io;
  tm.lap();
  for (size_t k = 0; k < n; k++) block(k);
  double ms = tm.lap();
  Cols cols = random_cols(std::max<size_t>(1, nrc >> shift), N, 4, 65536);
  size_t nc = cols.size();
  std::vector<size_t> positions;
  for (size_t i = 0; i < std::max<size_t>(1, (2 * num_io) >> shift); i++) positions.push_back(i * rows_per_io % N);
  tm.lap();
  { Cols c2(1, std::vector<GF>(N)); generate_pulse(c2, positions); }
  if (split) generate_split_u16_range_check(0, nc, cols); else generate_u16_range_check(0, nc, cols);
  ms += tm.lap();
  return ms;
}
}  // namespace orc
