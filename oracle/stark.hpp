// ORACLE (test infrastructure).  CPU restatement of `starky::prover::prove` and
// `starky::verifier::verify_stark_proof` (starky 0.1.1 / plonky2 0.1.3 @ InternetMaximalism rev
// 541e127 -- NOT vendored in /root/reference: Cargo.toml:18-21, Cargo.lock:529-531,797-799).  The
// published algorithm is restated from SURVEY.md App. B; the reference's call sites that anchor it:
// src/curves/g1/exp.rs:816-826, src/modular/modular.rs:545-558 (prove then verify_stark_proof).
// PARITY UNPINNED: the reference stores no prover golden vector; this oracle is
// pinned by (i) Poseidon's published KAT, (ii) prover<->verifier round trips and tamper tests.
#pragma once
#include "gl.hpp"
#include "hash.hpp"
#include "fft.hpp"
#include "air_common.hpp"
#include <chrono>
#include <map>
#include <string>
#include <stdexcept>

namespace orc {

struct StarkConfig {  // StarkConfig::standard_fast_config
  int security_bits = 100, num_challenges = 2;
  int rate_bits = 1, cap_height = 4, pow_bits = 16, arity_bits = 4, final_poly_bits = 5, num_query_rounds = 84;
  // U3: false = `quotient.coeffs.push(0)` padding (B.8); true = the older "max-degree hack" -- the batch quotients
  // are not padded and the FRI polynomial is multiplied by X (prover `final_poly.coeffs.insert(0, 0)`, verifier `sum * subgroup_x`).
  bool fri_degree_hack = false;
  std::vector<int> reduction_arity_bits(int degree_bits) const {  // FriReductionStrategy::ConstantArityBits
    std::vector<int> r;
    while (degree_bits > final_poly_bits && degree_bits + rate_bits - arity_bits >= cap_height) { r.push_back(arity_bits); degree_bits -= arity_bits; }
    return r;
  }
};

struct PolynomialBatch {  // plonky2::fri::oracle::PolynomialBatch (blinding = false)
  std::vector<std::vector<GF>> polynomials;  // coefficient form
  MerkleTree tree;                           // leaves = LDE rows in bit-reversed order
  int degree_log = 0, rate_bits = 0;
  static PolynomialBatch from_coeffs(std::vector<std::vector<GF>> coeffs, int rate_bits, int cap_height) {
    PolynomialBatch b;
    b.degree_log = log2_strict(coeffs[0].size()); b.rate_bits = rate_bits;
    size_t ncol = coeffs.size(), L = coeffs[0].size() << rate_bits;
    std::vector<std::vector<GF>> leaves(L, std::vector<GF>(ncol));
    int lb = log2_strict(L);
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t c = 0; c < ncol; c++) {
      std::vector<GF> v = lde_onto_coset(coeffs[c], rate_bits);
      for (size_t i = 0; i < L; i++) leaves[reverse_bits(i, lb)][c] = v[i];   // transpose + reverse_index_bits
    }
    b.polynomials = std::move(coeffs);
    b.tree = MerkleTree(std::move(leaves), cap_height);
    return b;
  }
  static PolynomialBatch from_values(const std::vector<std::vector<GF>>& values, int rate_bits, int cap_height) {
    std::vector<std::vector<GF>> coeffs(values.size());
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t c = 0; c < values.size(); c++) coeffs[c] = ifft(values[c]);
    return from_coeffs(std::move(coeffs), rate_bits, cap_height);
  }
  const std::vector<GF>& get_lde_values(size_t index, size_t step) const {
    return tree.leaves[reverse_bits(index * step, degree_log + rate_bits)];
  }
};

struct Openings { std::vector<GF2> local_values, next_values, permutation_zs, permutation_zs_next, quotient_polys; };
struct FriQueryStep { std::vector<GF2> evals; std::vector<Hash4> merkle_proof; };
struct FriInitialEval { std::vector<GF> evals; std::vector<Hash4> merkle_proof; };
struct FriQueryRound { std::vector<FriInitialEval> initial; std::vector<FriQueryStep> steps; };
struct FriProof { std::vector<std::vector<Hash4>> commit_caps; std::vector<FriQueryRound> rounds; std::vector<GF2> final_poly; GF pow_witness; };
struct Proof {
  std::vector<Hash4> trace_cap; bool has_perm = false; std::vector<Hash4> permutation_zs_cap, quotient_polys_cap;
  Openings openings; FriProof fri; std::vector<GF> public_inputs;
};

// Canonical wire format (defined by us; starky at this version has no serializer -- SURVEY.md B.9):
// fields in struct order, F = canonical LE u64, Ext = (a0, a1), Vec = u32 length prefix, Option = u8 tag.
struct ByteWriter {
  std::vector<uint8_t> b;
  void u8(uint8_t x) { b.push_back(x); }
  void u32_(u32 x) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(x >> (8 * i))); }
  void f(GF x) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(x.v >> (8 * i))); }
  void e(GF2 x) { f(x.a); f(x.b); }
  void h(const Hash4& x) { for (int i = 0; i < 4; i++) f(x.e[i]); }
  void hv(const std::vector<Hash4>& v) { u32_((u32)v.size()); for (auto& x : v) h(x); }
  void fv(const std::vector<GF>& v) { u32_((u32)v.size()); for (auto& x : v) f(x); }
  void ev(const std::vector<GF2>& v) { u32_((u32)v.size()); for (auto& x : v) e(x); }
};
struct ByteReader {
  const uint8_t* p; size_t n, o = 0;
  ByteReader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
  void need(size_t k) { if (o + k > n) throw std::runtime_error("proof truncated"); }
  uint8_t u8() { need(1); return p[o++]; }
  u32 u32_() { need(4); u32 x = 0; for (int i = 0; i < 4; i++) x |= (u32)p[o++] << (8 * i); return x; }
  GF f() { need(8); u64 x = 0; for (int i = 0; i < 8; i++) x |= (u64)p[o++] << (8 * i); if (x >= GP) throw std::runtime_error("non-canonical field element"); GF r; r.v = x; return r; }
  GF2 e() { GF a = f(); GF b = f(); return GF2(a, b); }
  Hash4 h() { Hash4 x; for (int i = 0; i < 4; i++) x.e[i] = f(); return x; }
  std::vector<Hash4> hv() { u32 k = u32_(); need((size_t)k * 32); std::vector<Hash4> v(k); for (auto& x : v) x = h(); return v; }
  std::vector<GF> fv() { u32 k = u32_(); need((size_t)k * 8); std::vector<GF> v(k); for (auto& x : v) x = f(); return v; }
  std::vector<GF2> ev() { u32 k = u32_(); need((size_t)k * 16); std::vector<GF2> v(k); for (auto& x : v) x = e(); return v; }
};
static inline std::vector<uint8_t> serialize_proof(const Proof& p) {
  ByteWriter w;
  w.hv(p.trace_cap);
  w.u8(p.has_perm ? 1 : 0); if (p.has_perm) w.hv(p.permutation_zs_cap);
  w.hv(p.quotient_polys_cap);
  w.ev(p.openings.local_values); w.ev(p.openings.next_values);
  if (p.has_perm) { w.ev(p.openings.permutation_zs); w.ev(p.openings.permutation_zs_next); }
  w.ev(p.openings.quotient_polys);
  w.u32_((u32)p.fri.commit_caps.size()); for (auto& c : p.fri.commit_caps) w.hv(c);
  w.u32_((u32)p.fri.rounds.size());
  for (auto& r : p.fri.rounds) {
    w.u32_((u32)r.initial.size()); for (auto& ie : r.initial) { w.fv(ie.evals); w.hv(ie.merkle_proof); }
    w.u32_((u32)r.steps.size()); for (auto& s : r.steps) { w.ev(s.evals); w.hv(s.merkle_proof); }
  }
  w.ev(p.fri.final_poly); w.f(p.fri.pow_witness);
  w.fv(p.public_inputs);
  return w.b;
}
static inline Proof deserialize_proof(const uint8_t* data, size_t n) {
  ByteReader r(data, n); Proof p;
  p.trace_cap = r.hv();
  p.has_perm = r.u8() != 0; if (p.has_perm) p.permutation_zs_cap = r.hv();
  p.quotient_polys_cap = r.hv();
  p.openings.local_values = r.ev(); p.openings.next_values = r.ev();
  if (p.has_perm) { p.openings.permutation_zs = r.ev(); p.openings.permutation_zs_next = r.ev(); }
  p.openings.quotient_polys = r.ev();
  u32 nc = r.u32_(); if (nc > 64) throw std::runtime_error("bad proof"); p.fri.commit_caps.resize(nc); for (auto& c : p.fri.commit_caps) c = r.hv();
  u32 nr = r.u32_(); if (nr > 4096) throw std::runtime_error("bad proof"); p.fri.rounds.resize(nr);
  for (auto& q : p.fri.rounds) {
    u32 ni = r.u32_(); if (ni > 8) throw std::runtime_error("bad proof"); q.initial.resize(ni); for (auto& ie : q.initial) { ie.evals = r.fv(); ie.merkle_proof = r.hv(); }
    u32 ns = r.u32_(); if (ns > 64) throw std::runtime_error("bad proof"); q.steps.resize(ns); for (auto& s : q.steps) { s.evals = r.ev(); s.merkle_proof = r.hv(); }
  }
  p.fri.final_poly = r.ev(); p.fri.pow_witness = r.f();
  p.public_inputs = r.fv();
  if (r.o != n) throw std::runtime_error("trailing bytes in proof");
  return p;
}

struct PermChallenge { GF beta, gamma; };
typedef std::vector<std::vector<PermChallenge>> PermChallengeSets;  // [num_challenges][batch_size]
// starky::permutation::get_n_permutation_challenge_sets
static inline PermChallengeSets get_n_permutation_challenge_sets(Challenger& ch, int num_challenges, int num_sets) {
  PermChallengeSets s(num_challenges);
  for (int i = 0; i < num_challenges; i++) for (int j = 0; j < num_sets; j++) { PermChallenge c; c.beta = ch.get(); c.gamma = ch.get(); s[i].push_back(c); }
  return s;
}
struct PermInstance { std::pair<size_t, size_t> pair; PermChallenge challenge; };
// starky::permutation::get_permutation_batches: cartesian(pairs, 0..num_challenges) chunked by batch_size;
// the i-th instance of a chunk uses challenge_sets[chal].challenges[i].
static inline std::vector<std::vector<PermInstance>> get_permutation_batches(const std::vector<std::pair<size_t, size_t>>& pairs,
                                                                            const PermChallengeSets& sets, int num_challenges, int batch_size) {
  std::vector<std::vector<PermInstance>> batches;
  std::vector<PermInstance> cur;
  for (auto& pr : pairs) for (int chal = 0; chal < num_challenges; chal++) {
    PermInstance in; in.pair = pr; in.challenge = sets[chal][cur.size()];
    cur.push_back(in);
    if ((int)cur.size() == batch_size) { batches.push_back(cur); cur.clear(); }
  }
  if (!cur.empty()) batches.push_back(cur);
  return batches;
}
// starky::permutation::compute_permutation_z_polys (singleton column pairs: reduced = gamma + col)
static inline std::vector<std::vector<GF>> compute_permutation_z_polys(const Air& air, const StarkConfig& cfg, const std::vector<std::vector<GF>>& trace,
                                                                       const PermChallengeSets& sets) {
  auto batches = get_permutation_batches(air.permutation_pairs(), sets, cfg.num_challenges, air.quotient_degree_factor());
  size_t n = trace[0].size();
  std::vector<std::vector<GF>> zs(batches.size());
#pragma omp parallel for schedule(dynamic, 4)
  for (size_t b = 0; b < batches.size(); b++) {
    std::vector<GF> num(n, GF::one()), den(n, GF::one());
    for (auto& in : batches[b]) {
      const auto& l = trace[in.pair.first]; const auto& r = trace[in.pair.second];
      for (size_t i = 0; i < n; i++) { num[i] = num[i] * (in.challenge.gamma + l[i]); den[i] = den[i] * (in.challenge.gamma + r[i]); }
    }
    // batch inverse of den
    std::vector<GF> pre(n); GF acc = GF::one();
    for (size_t i = 0; i < n; i++) { pre[i] = acc; acc = acc * den[i]; }
    GF inv = gl_inv(acc);
    for (size_t i = n; i-- > 0;) { GF di = inv * pre[i]; inv = inv * den[i]; num[i] = num[i] * di; }
    std::vector<GF> z(n); acc = GF::one();
    for (size_t i = 0; i < n; i++) { z[i] = acc; acc = acc * num[i]; }
    zs[b] = std::move(z);
  }
  return zs;
}
// starky::permutation::eval_permutation_checks (batches = get_permutation_batches(..), hoisted by the caller)
typedef std::vector<std::vector<PermInstance>> PermBatches;
template <class P> static inline void eval_permutation_checks(const PermBatches& batches, const P* lv, const P* local_zs, const P* next_zs,
                                                              size_t nz, Consumer<P>& yc) {
  for (size_t i = 0; i < nz; i++) yc.first_row(local_zs[i] - FieldOf<P>::c(1));
  assert(batches.size() == nz);
  for (size_t i = 0; i < batches.size(); i++) {
    P lhs = FieldOf<P>::c(1), rhs = FieldOf<P>::c(1);
    for (auto& in : batches[i]) {
      lhs = lhs * (lv[in.pair.first] + FieldOf<P>::from(in.challenge.gamma));
      rhs = rhs * (lv[in.pair.second] + FieldOf<P>::from(in.challenge.gamma));
    }
    yc.constraint(next_zs[i] * rhs - local_zs[i] * lhs);
  }
}

struct ProverDebug {  // intermediates exported for parity tests of the CUDA path
  std::vector<std::vector<GF>> z_polys;          // permutation Z columns (values)
  std::vector<std::vector<GF>> quotient_chunks;  // coefficient form
  std::vector<GF> alphas; GF2 zeta, fri_alpha; std::vector<GF2> fri_betas;
  PermChallengeSets perm_sets;
  std::vector<size_t> query_indices;
  std::map<std::string, double> timings_ms;
};
struct Timer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double lap() { auto t1 = std::chrono::steady_clock::now(); double ms = std::chrono::duration<double, std::milli>(t1 - t0).count(); t0 = t1; return ms; }
};

static inline GF2 eval_poly_ext(const std::vector<GF>& coeffs, GF2 z) {
  GF2 acc; for (size_t i = coeffs.size(); i-- > 0;) acc = acc * z + GF2(coeffs[i]); return acc;
}
static inline GF2 eval_poly_ext2(const std::vector<GF2>& coeffs, GF2 z) {
  GF2 acc; for (size_t i = coeffs.size(); i-- > 0;) acc = acc * z + coeffs[i]; return acc;
}

// plonky2 fri::prover::fri_proof_of_work.  Upstream uses rayon `find_any` (any valid witness);
// canonical choice here = smallest valid witness.
static inline GF fri_proof_of_work(Challenger& ch, const StarkConfig& cfg) {
  GF st0[12]; for (int i = 0; i < 12; i++) st0[i] = ch.st[i];
  size_t pos = ch.in.size();
  for (size_t i = 0; i < pos; i++) st0[i] = ch.in[i];
  u64 found = ~0ULL;
  for (u64 base = 0; found == ~0ULL; base += (1 << 15)) {
    u64 best = ~0ULL;
#pragma omp parallel for reduction(min : best)
    for (long long c = (long long)base; c < (long long)(base + (1 << 15)); c++) {
      if ((u64)c > best) continue;
      GF st[12]; for (int i = 0; i < 12; i++) st[i] = st0[i];
      st[pos] = GF((u64)c);
      poseidon_fast(st);
      if (__builtin_clzll(st[7].v | 1) >= cfg.pow_bits && (st[7].v >> (64 - cfg.pow_bits)) == 0) { if ((u64)c < best) best = (u64)c; }
    }
    found = best;
  }
  GF w = GF(found);
  ch.observe(w);
  GF resp = ch.get();
  assert((resp.v >> (64 - cfg.pow_bits)) == 0);
  return w;
}

static inline std::vector<GF> flatten_ext(const GF2* v, size_t n) { std::vector<GF> r(2 * n); for (size_t i = 0; i < n; i++) { r[2 * i] = v[i].a; r[2 * i + 1] = v[i].b; } return r; }

// starky::prover::prove
static inline Proof prove(const Air& air, const StarkConfig& cfg, const std::vector<std::vector<GF>>& trace, const std::vector<GF>& public_inputs,
                          ProverDebug* dbg = nullptr) {
  Timer tm; std::map<std::string, double> tms;
  size_t degree = trace[0].size();
  int degree_bits = log2_strict(degree);
  if (trace.size() != air.num_columns()) throw std::runtime_error("trace width != num_columns");
  if (public_inputs.size() != air.num_public_inputs()) throw std::runtime_error("public input count mismatch");
  std::vector<int> arities = cfg.reduction_arity_bits(degree_bits);
  int total_arities = 0; for (int a : arities) total_arities += a;
  if (total_arities > degree_bits + cfg.rate_bits - cfg.cap_height) throw std::runtime_error("FRI total reduction arity is too large.");
  int rate_bits = cfg.rate_bits;

  PolynomialBatch trace_commitment = PolynomialBatch::from_values(trace, rate_bits, cfg.cap_height);
  tms["trace_commitment"] = tm.lap();
  Proof proof; proof.public_inputs = public_inputs;
  proof.trace_cap = trace_commitment.tree.cap;
  Challenger ch;
  ch.observe_cap(proof.trace_cap);

  auto pairs = air.permutation_pairs();
  bool uses_perm = !pairs.empty();
  PermChallengeSets perm_sets; PolynomialBatch z_commitment;
  if (uses_perm) {
    perm_sets = get_n_permutation_challenge_sets(ch, cfg.num_challenges, air.quotient_degree_factor());
    auto z_polys = compute_permutation_z_polys(air, cfg, trace, perm_sets);
    tms["z_polys"] = tm.lap();
    if (dbg) dbg->z_polys = z_polys;
    z_commitment = PolynomialBatch::from_values(z_polys, rate_bits, cfg.cap_height);
    tms["z_commitment"] = tm.lap();
    proof.has_perm = true; proof.permutation_zs_cap = z_commitment.tree.cap;
    ch.observe_cap(proof.permutation_zs_cap);
  }
  std::vector<GF> alphas(cfg.num_challenges);
  for (auto& a : alphas) a = ch.get();

  // ---- compute_quotient_polys ----
  int qdf = air.quotient_degree_factor();
  int qdb = 0; while ((1 << qdb) < qdf) qdb++;
  if (qdb > rate_bits) throw std::runtime_error("constraint degree higher than the rate is not supported");
  size_t step = size_t(1) << (rate_bits - qdb), next_step = size_t(1) << qdb;
  size_t size = degree << qdb;
  std::vector<GF> sel(degree); sel[0] = GF::one();
  std::vector<GF> lagrange_first = lde_onto_coset(ifft(sel), qdb);
  sel[0] = GF(); sel[degree - 1] = GF::one();
  std::vector<GF> lagrange_last = lde_onto_coset(ifft(sel), qdb);
  // ZeroPolyOnCoset
  GF g_pow_n = gl_exp_pow2(coset_shift(), degree_bits);
  std::vector<GF> zh_inv(size_t(1) << qdb);
  { GF w = root_of_unity(qdb), x = GF::one(); for (auto& v : zh_inv) { v = gl_inv(g_pow_n * x - GF::one()); x = x * w; } }
  GF last = gl_inv(root_of_unity(degree_bits));
  std::vector<GF> coset(size);
  { GF w = root_of_unity(degree_bits + qdb), x = coset_shift(); for (auto& v : coset) { v = x; x = x * w; } }
  size_t nz = uses_perm ? z_commitment.polynomials.size() : 0;
  PermBatches perm_batches; if (uses_perm) perm_batches = get_permutation_batches(pairs, perm_sets, cfg.num_challenges, air.quotient_degree_factor());
  std::vector<std::vector<GF>> quotient_values(cfg.num_challenges, std::vector<GF>(size));
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < size; i++) {
    size_t i_next = (i + next_step) % size;
    Consumer<GF> yc(alphas, coset[i] - last, lagrange_first[i], lagrange_last[i]);
    const std::vector<GF>& lv = trace_commitment.get_lde_values(i, step);
    const std::vector<GF>& nv = trace_commitment.get_lde_values(i_next, step);
    air.eval(lv.data(), nv.data(), public_inputs.data(), yc);
    if (uses_perm) {
      const std::vector<GF>& lz = z_commitment.get_lde_values(i, step);
      const std::vector<GF>& nzv = z_commitment.get_lde_values(i_next, step);
      eval_permutation_checks<GF>(perm_batches, lv.data(), lz.data(), nzv.data(), nz, yc);
    }
    GF dinv = zh_inv[i % zh_inv.size()];
    for (int k = 0; k < cfg.num_challenges; k++) quotient_values[k][i] = yc.accs[k] * dinv;
  }
  std::vector<std::vector<GF>> all_quotient_chunks;
  for (int k = 0; k < cfg.num_challenges; k++) {
    std::vector<GF> qc = coset_ifft(quotient_values[k], coset_shift());
    // trim_to_len(degree * qdf): everything beyond must be zero
    for (size_t i = degree * qdf; i < qc.size(); i++) if (qc[i].v) throw std::runtime_error("Quotient has failed, the vanishing polynomial is not divisible by Z_H");
    for (int c = 0; c < qdf; c++) all_quotient_chunks.emplace_back(qc.begin() + c * degree, qc.begin() + (c + 1) * degree);
  }
  tms["quotient_polys"] = tm.lap();
  if (dbg) dbg->quotient_chunks = all_quotient_chunks;
  PolynomialBatch quotient_commitment = PolynomialBatch::from_coeffs(all_quotient_chunks, rate_bits, cfg.cap_height);
  tms["quotient_commitment"] = tm.lap();
  proof.quotient_polys_cap = quotient_commitment.tree.cap;
  ch.observe_cap(proof.quotient_polys_cap);

  GF2 zeta = ch.get_ext();
  GF g = root_of_unity(degree_bits);
  if (gf2_exp_pow2(zeta, degree_bits) == GF2::one()) throw std::runtime_error("Opening point is in the subgroup.");
  GF2 zeta_next = zeta * g;
  auto eval_commitment = [](GF2 z, const PolynomialBatch& c) {
    std::vector<GF2> r(c.polynomials.size());
#pragma omp parallel for schedule(dynamic, 8)
    for (size_t i = 0; i < r.size(); i++) r[i] = eval_poly_ext(c.polynomials[i], z);
    return r;
  };
  Openings& op = proof.openings;
  op.local_values = eval_commitment(zeta, trace_commitment);
  op.next_values = eval_commitment(zeta_next, trace_commitment);
  if (uses_perm) { op.permutation_zs = eval_commitment(zeta, z_commitment); op.permutation_zs_next = eval_commitment(zeta_next, z_commitment); }
  op.quotient_polys = eval_commitment(zeta, quotient_commitment);
  // challenger.observe_openings(&openings.to_fri_openings())
  for (auto& v : op.local_values) ch.observe(v);
  for (auto& v : op.permutation_zs) ch.observe(v);
  for (auto& v : op.quotient_polys) ch.observe(v);
  for (auto& v : op.next_values) ch.observe(v);
  for (auto& v : op.permutation_zs_next) ch.observe(v);
  tms["openings"] = tm.lap();

  // ---- PolynomialBatch::prove_openings ----
  std::vector<const PolynomialBatch*> oracles; oracles.push_back(&trace_commitment); if (uses_perm) oracles.push_back(&z_commitment); oracles.push_back(&quotient_commitment);
  GF2 alpha = ch.get_ext();
  std::vector<GF2> final_poly;  // PolynomialCoeffs::empty()
  for (int batch = 0; batch < 2; batch++) {
    GF2 point = batch == 0 ? zeta : zeta_next;
    std::vector<const std::vector<GF>*> polys;
    for (size_t o = 0; o < oracles.size(); o++) {
      if (batch == 1 && o + 1 == oracles.size()) continue;  // zeta_next batch has no quotient polys
      for (auto& p : oracles[o]->polynomials) polys.push_back(&p);
    }
    // alpha.reduce_polys_base: sum_j alpha^j f_j
    std::vector<GF2> apow(polys.size()); { GF2 a = GF2::one(); for (auto& x : apow) { x = a; a = a * alpha; } }
    std::vector<GF2> comp(degree);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < degree; i++) { GF2 acc; for (size_t j = 0; j < polys.size(); j++) acc = acc + apow[j] * (*polys[j])[i]; comp[i] = acc; }
    // divide_by_linear(point) then pad back with one zero
    std::vector<GF2> quot(degree);
    { GF2 acc; std::vector<GF2> bs(degree); for (size_t i = degree; i-- > 0;) { acc = acc * point + comp[i]; bs[i] = acc; }
      for (size_t i = 0; i + 1 < degree; i++) { quot[i] = bs[i + 1]; }
      quot[degree - 1] = GF2(); }
    // alpha.shift_poly(&mut final_poly); final_poly += quotient
    GF2 shift = gf2_pow(alpha, polys.size());
    if (final_poly.empty()) final_poly.assign(degree, GF2());
    for (size_t i = 0; i < degree; i++) final_poly[i] = final_poly[i] * shift + quot[i];
  }
  if (cfg.fri_degree_hack) {   // coefficient degree-1 of every padded quotient is zero, so the product with X still has `degree` coefficients
    final_poly.insert(final_poly.begin(), GF2()); final_poly.pop_back();
  }
  std::vector<GF2> lde_coeffs = final_poly; lde_coeffs.resize(degree << rate_bits);
  std::vector<GF2> lde_values = coset_fft_ext(lde_coeffs, coset_shift());
  tms["fri_reduce"] = tm.lap();

  // ---- fri_committed_trees ----
  FriProof& fp = proof.fri;
  std::vector<MerkleTree> trees;
  {
    std::vector<GF2> coeffs = lde_coeffs, values = lde_values;
    GF shift = coset_shift();
    for (int arity_bits : arities) {
      size_t arity = size_t(1) << arity_bits;
      reverse_index_bits_in_place(values);
      std::vector<std::vector<GF>> chunked(values.size() / arity);
      for (size_t i = 0; i < chunked.size(); i++) chunked[i] = flatten_ext(&values[i * arity], arity);
      trees.emplace_back(std::move(chunked), cfg.cap_height);
      ch.observe_cap(trees.back().cap);
      GF2 beta = ch.get_ext();
      if (dbg) dbg->fri_betas.push_back(beta);
      std::vector<GF2> folded(coeffs.size() / arity);
      for (size_t i = 0; i < folded.size(); i++) { GF2 acc; for (size_t k = arity; k-- > 0;) acc = acc * beta + coeffs[i * arity + k]; folded[i] = acc; }
      coeffs.swap(folded);
      shift = gl_pow(shift, arity);
      values = coset_fft_ext(coeffs, shift);
    }
    coeffs.resize(coeffs.size() >> rate_bits);
    for (auto& c : coeffs) ch.observe(c);
    fp.final_poly = coeffs;
  }
  for (auto& t : trees) fp.commit_caps.push_back(t.cap);
  tms["fri_commit"] = tm.lap();
  fp.pow_witness = fri_proof_of_work(ch, cfg);
  tms["fri_pow"] = tm.lap();
  // ---- fri_prover_query_rounds ----
  size_t n = degree << rate_bits;
  std::vector<const MerkleTree*> initial_trees; for (auto* o : oracles) initial_trees.push_back(&o->tree);
  for (int q = 0; q < cfg.num_query_rounds; q++) {
    size_t x_index = ch.get().v % n;
    if (dbg) dbg->query_indices.push_back(x_index);
    FriQueryRound round;
    for (auto* t : initial_trees) { FriInitialEval ie; ie.evals = t->leaves[x_index]; ie.merkle_proof = t->prove(x_index); round.initial.push_back(std::move(ie)); }
    for (size_t i = 0; i < trees.size(); i++) {
      int ab = arities[i];
      const std::vector<GF>& leaf = trees[i].leaves[x_index >> ab];
      FriQueryStep st; st.evals.resize(leaf.size() / 2);
      for (size_t k = 0; k < st.evals.size(); k++) st.evals[k] = GF2(leaf[2 * k], leaf[2 * k + 1]);
      st.merkle_proof = trees[i].prove(x_index >> ab);
      round.steps.push_back(std::move(st));
      x_index >>= ab;
    }
    fp.rounds.push_back(std::move(round));
  }
  tms["fri_queries"] = tm.lap();
  if (dbg) { dbg->alphas = alphas; dbg->zeta = zeta; dbg->fri_alpha = alpha; dbg->perm_sets = perm_sets; dbg->timings_ms = tms; }
  return proof;
}

// starky::verifier::verify_stark_proof.  Returns "" on success, otherwise the reason.
static inline std::string verify_stark_proof(const Air& air, const Proof& proof, const StarkConfig& cfg) {
  size_t ncols = air.num_columns(), cap_n = size_t(1) << cfg.cap_height;
  size_t nz = air.permutation_pairs().size() * cfg.num_challenges;
  int qdf = air.quotient_degree_factor();
  nz = (nz + qdf - 1) / qdf;
  bool uses_perm = nz > 0;
  // validate_proof_shape
  if (proof.public_inputs.size() != air.num_public_inputs()) return "public input count";
  if (proof.trace_cap.size() != cap_n || proof.quotient_polys_cap.size() != cap_n) return "cap size";
  if (proof.has_perm != uses_perm) return "permutation cap presence";
  if (uses_perm && proof.permutation_zs_cap.size() != cap_n) return "cap size";
  const Openings& op = proof.openings;
  if (op.local_values.size() != ncols || op.next_values.size() != ncols) return "opening width";
  if (op.permutation_zs.size() != nz || op.permutation_zs_next.size() != nz) return "Z opening width";
  if (op.quotient_polys.size() != (size_t)qdf * cfg.num_challenges) return "quotient opening width";
  const FriProof& fp = proof.fri;
  if ((int)fp.rounds.size() != cfg.num_query_rounds || fp.rounds.empty()) return "query round count";
  if (fp.rounds[0].initial.empty()) return "initial proofs";
  // recover_degree_bits
  int lde_bits = cfg.cap_height + (int)fp.rounds[0].initial[0].merkle_proof.size();
  int degree_bits = lde_bits - cfg.rate_bits;
  if (degree_bits < 1 || degree_bits > 30) return "degree bits";
  std::vector<int> arities = cfg.reduction_arity_bits(degree_bits);
  if (fp.commit_caps.size() != arities.size()) return "commit phase cap count";
  { int tot = 0; for (int a : arities) tot += a; if (fp.final_poly.size() != (size_t(1) << (degree_bits - tot))) return "final poly length"; }
  size_t noracles = uses_perm ? 3 : 2;
  for (auto& r : fp.rounds) {
    if (r.initial.size() != noracles || r.steps.size() != arities.size()) return "query round shape";
    size_t widths[3] = {ncols, uses_perm ? nz : (size_t)qdf * cfg.num_challenges, (size_t)qdf * cfg.num_challenges};
    for (size_t o = 0; o < noracles; o++) { if (r.initial[o].evals.size() != widths[o]) return "initial eval width"; if ((int)r.initial[o].merkle_proof.size() != lde_bits - cfg.cap_height) return "merkle proof length"; }
    int cur = lde_bits;
    for (size_t i = 0; i < arities.size(); i++) {
      if (r.steps[i].evals.size() != (size_t(1) << arities[i])) return "step eval width";
      cur -= arities[i];
      int plen = cur - cfg.cap_height; if (plen < 0) plen = 0;
      if ((int)r.steps[i].merkle_proof.size() != plen) return "step merkle proof length";
    }
  }
  // get_challenges
  Challenger ch;
  ch.observe_cap(proof.trace_cap);
  PermChallengeSets perm_sets;
  if (uses_perm) { perm_sets = get_n_permutation_challenge_sets(ch, cfg.num_challenges, qdf); ch.observe_cap(proof.permutation_zs_cap); }
  std::vector<GF> alphas(cfg.num_challenges); for (auto& a : alphas) a = ch.get();
  ch.observe_cap(proof.quotient_polys_cap);
  GF2 zeta = ch.get_ext();
  for (auto& v : op.local_values) ch.observe(v);
  for (auto& v : op.permutation_zs) ch.observe(v);
  for (auto& v : op.quotient_polys) ch.observe(v);
  for (auto& v : op.next_values) ch.observe(v);
  for (auto& v : op.permutation_zs_next) ch.observe(v);
  GF2 fri_alpha = ch.get_ext();
  std::vector<GF2> fri_betas;
  for (auto& c : fp.commit_caps) { if (c.size() != cap_n) return "commit cap size"; ch.observe_cap(c); fri_betas.push_back(ch.get_ext()); }
  for (auto& c : fp.final_poly) ch.observe(c);
  ch.observe(fp.pow_witness);
  GF pow_response = ch.get();
  size_t lde_size = size_t(1) << lde_bits;
  std::vector<size_t> query_indices(cfg.num_query_rounds); for (auto& q : query_indices) q = ch.get().v % lde_size;

  // vanishing polynomial at zeta
  GF g = root_of_unity(degree_bits);
  GF2 zeta_pow_deg = gf2_exp_pow2(zeta, degree_bits);
  GF2 z_x = zeta_pow_deg - GF2::one();
  GF nF = GF((u64)1 << degree_bits);
  GF2 l_0 = z_x * gf2_inv((zeta - GF2::one()) * nF), l_last = z_x * gf2_inv((zeta * g - GF2::one()) * nF);
  GF2 z_last = zeta - GF2(gl_inv(g));
  std::vector<GF2> alphas_ext; for (auto a : alphas) alphas_ext.push_back(GF2(a));
  Consumer<GF2> yc(alphas_ext, z_last, l_0, l_last);
  std::vector<GF2> pis; for (auto v : proof.public_inputs) pis.push_back(GF2(v));
  air.eval_ext(op.local_values.data(), op.next_values.data(), pis.data(), yc);
  if (uses_perm) { PermBatches pb = get_permutation_batches(air.permutation_pairs(), perm_sets, cfg.num_challenges, qdf); eval_permutation_checks<GF2>(pb, op.local_values.data(), op.permutation_zs.data(), op.permutation_zs_next.data(), nz, yc); }
  for (int i = 0; i < cfg.num_challenges; i++) {
    GF2 acc; for (int k = qdf; k-- > 0;) acc = acc * zeta_pow_deg + op.quotient_polys[i * qdf + k];
    if (yc.accs[i] != z_x * acc) return "Mismatch between evaluation and opening of quotient polynomial";
  }
  // verify_fri_proof
  if ((pow_response.v >> (64 - cfg.pow_bits)) != 0) return "Invalid proof-of-work witness";
  std::vector<const std::vector<Hash4>*> caps; caps.push_back(&proof.trace_cap); if (uses_perm) caps.push_back(&proof.permutation_zs_cap); caps.push_back(&proof.quotient_polys_cap);
  GF2 zeta_next = zeta * g;
  // PrecomputedReducedOpenings: reduce(batch values) with Horner from the back
  auto reduce = [&](const std::vector<GF2>& v) { GF2 acc; for (size_t i = v.size(); i-- > 0;) acc = acc * fri_alpha + v[i]; return acc; };
  std::vector<GF2> b0, b1;
  b0.insert(b0.end(), op.local_values.begin(), op.local_values.end()); b0.insert(b0.end(), op.permutation_zs.begin(), op.permutation_zs.end()); b0.insert(b0.end(), op.quotient_polys.begin(), op.quotient_polys.end());
  b1.insert(b1.end(), op.next_values.begin(), op.next_values.end()); b1.insert(b1.end(), op.permutation_zs_next.begin(), op.permutation_zs_next.end());
  GF2 red0 = reduce(b0), red1 = reduce(b1);
  for (int q = 0; q < cfg.num_query_rounds; q++) {
    const FriQueryRound& r = fp.rounds[q];
    size_t x_index = query_indices[q];
    for (size_t o = 0; o < noracles; o++)
      if (!verify_merkle_proof_to_cap(r.initial[o].evals.data(), r.initial[o].evals.size(), x_index, *caps[o], r.initial[o].merkle_proof)) return "Invalid Merkle proof (initial tree)";
    GF subgroup_x_b = coset_shift() * gl_pow(root_of_unity(lde_bits), reverse_bits(x_index, lde_bits));
    GF2 subgroup_x = GF2(subgroup_x_b);
    // fri_combine_initial
    GF2 sum;
    for (int batch = 0; batch < 2; batch++) {
      std::vector<GF2> ev;
      for (size_t o = 0; o < noracles; o++) { if (batch == 1 && o + 1 == noracles) continue; for (auto v : r.initial[o].evals) ev.push_back(GF2(v)); }
      GF2 reduced = reduce(ev);
      GF2 numerator = reduced - (batch == 0 ? red0 : red1);
      GF2 denominator = subgroup_x - (batch == 0 ? zeta : zeta_next);
      sum = sum * gf2_pow(fri_alpha, ev.size());
      sum = sum + numerator * gf2_inv(denominator);
    }
    if (cfg.fri_degree_hack) sum = sum * subgroup_x;
    GF2 old_eval = sum;
    for (size_t i = 0; i < arities.size(); i++) {
      int ab = arities[i]; size_t arity = size_t(1) << ab;
      const std::vector<GF2>& evals = r.steps[i].evals;
      size_t coset_index = x_index >> ab, within = x_index & (arity - 1);
      if (evals[within] != old_eval) return "FRI consistency check failed";
      // compute_evaluation: interpolate {(coset_start*g^i, evals_rev[i])} at beta
      GF gg = root_of_unity(ab);
      std::vector<GF2> ev = evals; reverse_index_bits_in_place(ev);
      size_t rev_within = reverse_bits(within, ab);
      GF coset_start = subgroup_x_b * gl_pow(gg, arity - rev_within);
      std::vector<GF2> xs(arity); { GF y = GF::one(); for (size_t k = 0; k < arity; k++) { xs[k] = GF2(coset_start * y); y = y * gg; } }
      GF2 beta = fri_betas[i], res;
      for (size_t k = 0; k < arity; k++) {
        GF2 num = GF2::one(), den = GF2::one();
        for (size_t m = 0; m < arity; m++) if (m != k) { num = num * (beta - xs[m]); den = den * (xs[k] - xs[m]); }
        res = res + ev[k] * num * gf2_inv(den);
      }
      old_eval = res;
      std::vector<GF> flat = flatten_ext(evals.data(), evals.size());
      if (!verify_merkle_proof_to_cap(flat.data(), flat.size(), coset_index, fp.commit_caps[i], r.steps[i].merkle_proof)) return "Invalid Merkle proof (FRI layer)";
      subgroup_x_b = gl_exp_pow2(subgroup_x_b, ab);
      x_index = coset_index;
    }
    if (eval_poly_ext2(fp.final_poly, GF2(subgroup_x_b)) != old_eval) return "Final polynomial evaluation is invalid.";
  }
  return "";
}
}  // namespace orc
