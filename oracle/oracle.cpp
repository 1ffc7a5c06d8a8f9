// ORACLE (test infrastructure, NOT product code): C entry points over the CPU restatement, loaded
// with ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// only.  Parity status: PARITY UNPINNED at the prover boundary (no golden vectors exist in the
// reference) -- pinned by Poseidon's published KAT, arithmetic identities and
// prover<->verifier round trips.
#include "stark.hpp"
#include "air_modular.hpp"
#include "air_g1.hpp"
#include "air_g2.hpp"
#include "air_fq12.hpp"
#include "air_gadgets.hpp"
#include "sample.hpp"
#include <cstdio>
#include <omp.h>
#include <cstdlib>
#include <memory>
#include <sstream>

using namespace orc;
namespace orc { FieldParams g_field; }

enum { AIR_MODULAR = 0, AIR_FQ_EXP = 1, AIR_G1_EXP = 2, AIR_G2_EXP = 3, AIR_FQ12_EXP = 4, AIR_FQ12_EXP_U64 = 5, AIR_G1_MULADD = 6, AIR_FQ12_MUL = 7 };

struct OrcConfig { uint32_t security_bits, num_challenges, rate_bits, cap_height, pow_bits, fri_arity_bits, fri_final_poly_bits, num_query_rounds; uint64_t coset_shift;
                   uint32_t fri_degree_hack, reserved; };
// U1: the coset shift is the field's multiplicative generator g and the two-adic generator is
// g^((p - 1) / 2^32); 0 or 7 selects plonky2_field pair (B) = (7, 1753635133440165772), 14293326489335486720 pair (A).
static void select_field(const OrcConfig* c) {
  const u64 g = (c && c->coset_shift) ? c->coset_shift : 7;
  const GF r = gl_pow(GF(g), 0xFFFFFFFFULL);   // (p - 1) / 2^32 = 2^32 - 1
  if (gl_exp_pow2(r, 31).v != 0xFFFFFFFF00000000ULL) throw std::runtime_error("coset_shift is not a generator of the multiplicative group (its two-adic part has order < 2^32)");
  g_field.mult_generator = g; g_field.pow2_generator = r.v;
}
static StarkConfig to_cfg(const OrcConfig* c) {
  StarkConfig s;
  if (c) { s.security_bits = c->security_bits; s.num_challenges = c->num_challenges; s.rate_bits = c->rate_bits; s.cap_height = c->cap_height; s.pow_bits = c->pow_bits;
           s.arity_bits = c->fri_arity_bits; s.final_poly_bits = c->fri_final_poly_bits; s.num_query_rounds = c->num_query_rounds; s.fri_degree_hack = c->fri_degree_hack != 0; }
  return s;
}
struct AirHandle { int id; size_t num_io; std::unique_ptr<Air> air; };
static ProverDebug g_dbg;
static std::string g_err;

// packed input records: identical to the sbn_*_io structs of include/starky_bn254_b200.h
struct G1IoBlob { u64 x_x[4], x_y[4], off_x[4], off_y[4]; u32 exp[8]; u64 out_x[4], out_y[4]; };
struct FqIoBlob { u64 x[4], off[4]; u32 exp[8]; u64 out[4]; };
struct G2IoBlob { u64 x[16], off[16]; u32 exp[8]; u64 out[16]; };          // point = x.c0 x.c1 y.c0 y.c1
struct Fq12IoBlob { u64 x[48], off[48]; u32 exp[8]; u64 out[48]; };        // 12 coefficients, MyFq12 order
struct Fq12U64IoBlob { u64 x[48], off[48]; u64 exp; u64 out[48]; };
static U256 mk(const u64* p) { U256 r; for (int i = 0; i < 4; i++) r.w[i] = p[i]; return r; }
static G2Point mkg2(const u64* p) { return {mk(p), mk(p + 4), mk(p + 8), mk(p + 12)}; }
static Fq12Words mk12(const u64* p) { Fq12Words w; for (int i = 0; i < 12; i++) w.c[i] = mk(p + 4 * i); return w; }

extern "C" {
const char* orc_last_error() { return g_err.c_str(); }
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_max_threads() { return omp_get_max_threads(); }
void orc_set_field_params(u64 gen, u64 pow2) { g_field.mult_generator = gen; g_field.pow2_generator = pow2; }
void orc_poseidon_fast(u64* st) { GF s[12]; for (int i = 0; i < 12; i++) s[i] = GF(st[i]); poseidon_fast(s); for (int i = 0; i < 12; i++) st[i] = s[i].v; }
void orc_poseidon(u64* st) { GF s[12]; for (int i = 0; i < 12; i++) s[i] = GF(st[i]); poseidon(s); for (int i = 0; i < 12; i++) st[i] = s[i].v; }
void orc_hash_or_noop(const u64* in, size_t n, u64* out4) { std::vector<GF> v(n); for (size_t i = 0; i < n; i++) v[i] = GF(in[i]); Hash4 h = hash_or_noop(v.data(), n); for (int i = 0; i < 4; i++) out4[i] = h.e[i].v; }
void orc_two_to_one(const u64* l, const u64* r, u64* out4) { Hash4 a, b; for (int i = 0; i < 4; i++) { a.e[i] = GF(l[i]); b.e[i] = GF(r[i]); } Hash4 h = two_to_one(a, b); for (int i = 0; i < 4; i++) out4[i] = h.e[i].v; }
u64 orc_gl_mul(u64 a, u64 b) { return (GF(a) * GF(b)).v; }
u64 orc_gl_inv(u64 a) { return gl_inv(GF(a)).v; }
u64 orc_root_of_unity(int logn) { return root_of_unity(logn).v; }
// in-place FFT of one column (natural order in/out)
void orc_fft(u64* data, int logn, int inverse) { std::vector<GF> v(size_t(1) << logn); for (size_t i = 0; i < v.size(); i++) v[i] = GF(data[i]); fft_inplace(v, inverse != 0); for (size_t i = 0; i < v.size(); i++) data[i] = v[i].v; }
// values (ncols x n, column-major) -> coefficients and LDE values (ncols x n<<rate_bits, natural order)
void orc_commit_columns(const u64* values, size_t ncols, int logn, int rate_bits, int cap_height, u64* coeffs_out, u64* lde_out, u64* cap_out) {
  size_t n = size_t(1) << logn, L = n << rate_bits;
  std::vector<std::vector<GF>> cols(ncols, std::vector<GF>(n));
  for (size_t c = 0; c < ncols; c++) for (size_t i = 0; i < n; i++) cols[c][i] = GF(values[c * n + i]);
  PolynomialBatch b = PolynomialBatch::from_values(cols, rate_bits, cap_height);
  if (coeffs_out) for (size_t c = 0; c < ncols; c++) for (size_t i = 0; i < n; i++) coeffs_out[c * n + i] = b.polynomials[c][i].v;
  if (lde_out) for (size_t c = 0; c < ncols; c++) for (size_t i = 0; i < L; i++) lde_out[c * L + i] = b.get_lde_values(i, 1)[c].v;
  if (cap_out) for (size_t k = 0; k < b.tree.cap.size(); k++) for (int j = 0; j < 4; j++) cap_out[4 * k + j] = b.tree.cap[k].e[j].v;
}
// Merkle tree over row-major leaves; writes the cap and (optionally) the sibling path of `prove_index`.
void orc_merkle(const u64* leaves, size_t nleaves, size_t width, int cap_height, u64* cap_out, size_t prove_index, u64* path_out) {
  std::vector<std::vector<GF>> lv(nleaves, std::vector<GF>(width));
  for (size_t i = 0; i < nleaves; i++) for (size_t j = 0; j < width; j++) lv[i][j] = GF(leaves[i * width + j]);
  MerkleTree t(std::move(lv), cap_height);
  for (size_t k = 0; k < t.cap.size(); k++) for (int j = 0; j < 4; j++) cap_out[4 * k + j] = t.cap[k].e[j].v;
  if (path_out) { auto p = t.prove(prove_index); for (size_t k = 0; k < p.size(); k++) for (int j = 0; j < 4; j++) path_out[4 * k + j] = p[k].e[j].v; }
}
// Challenger transcript replay: observe `n_obs` elements then draw `n_out` challenges.
void orc_challenger(const u64* obs, size_t n_obs, u64* out, size_t n_out) { Challenger ch; for (size_t i = 0; i < n_obs; i++) ch.observe(GF(obs[i])); for (size_t i = 0; i < n_out; i++) out[i] = ch.get().v; }

void* orc_air_create(int id, size_t num_io) {
  AirHandle* h = new AirHandle{id, num_io, nullptr};
  switch (id) {
    case AIR_MODULAR: h->air.reset(new ModularStark()); break;
    case AIR_G1_EXP: h->air.reset(new G1ExpStark(num_io)); break;
    case AIR_FQ_EXP: h->air.reset(new FqExpStark(num_io)); break;
    case AIR_G2_EXP: h->air.reset(new G2ExpStark(num_io)); break;
    case AIR_FQ12_EXP: h->air.reset(new Fq12ExpStark(num_io)); break;
    case AIR_FQ12_EXP_U64: h->air.reset(new Fq12ExpU64Stark(num_io)); break;
    case AIR_G1_MULADD: h->air.reset(new G1Stark()); break;
    case AIR_FQ12_MUL: h->air.reset(new Fq12Stark()); break;
    default: delete h; g_err = "unsupported air"; return nullptr;
  }
  return h;
}
void orc_air_destroy(void* p) { delete (AirHandle*)p; }
size_t orc_air_num_columns(void* p) { return ((AirHandle*)p)->air->num_columns(); }
size_t orc_air_num_public_inputs(void* p) { return ((AirHandle*)p)->air->num_public_inputs(); }
size_t orc_air_num_rows(void* p) { AirHandle* h = (AirHandle*)p; return (h->id == AIR_MODULAR || h->id == AIR_G1_MULADD || h->id == AIR_FQ12_MUL) ? h->num_io : (h->id == AIR_FQ12_EXP_U64 ? 128 : 512) * h->num_io; }
size_t orc_air_num_permutation_pairs(void* p) { return ((AirHandle*)p)->air->permutation_pairs().size(); }
size_t orc_air_result_words(void* p) { AirHandle* h = (AirHandle*)p; switch (h->id) { case AIR_FQ_EXP: return 4; case AIR_G1_EXP: return 8; case AIR_G2_EXP: return 16; case AIR_FQ12_EXP: case AIR_FQ12_EXP_U64: return 48; } return 0; }
size_t orc_air_io_size(void* p) { AirHandle* h = (AirHandle*)p; switch (h->id) { case AIR_MODULAR: return 64; case AIR_G1_MULADD: return 128; case AIR_FQ12_MUL: return 768; case AIR_G1_EXP: return sizeof(G1IoBlob); case AIR_FQ_EXP: return sizeof(FqIoBlob); case AIR_G2_EXP: return sizeof(G2IoBlob);
  case AIR_FQ12_EXP: return sizeof(Fq12IoBlob); case AIR_FQ12_EXP_U64: return sizeof(Fq12U64IoBlob); } return 0; }

static std::vector<G1ExpIONative> g1_ios(const void* ios, size_t n) {
  const G1IoBlob* b = (const G1IoBlob*)ios; std::vector<G1ExpIONative> v(n);
  for (size_t i = 0; i < n; i++) { v[i].x = {mk(b[i].x_x), mk(b[i].x_y)}; v[i].offset = {mk(b[i].off_x), mk(b[i].off_y)}; memcpy(v[i].exp_val, b[i].exp, 32); v[i].output = {mk(b[i].out_x), mk(b[i].out_y)}; }
  return v;
}
static std::vector<FqExpIONative> fq_ios(const void* ios, size_t n) {
  const FqIoBlob* b = (const FqIoBlob*)ios; std::vector<FqExpIONative> v(n);
  for (size_t i = 0; i < n; i++) { v[i].x = mk(b[i].x); v[i].offset = mk(b[i].off); memcpy(v[i].exp_val, b[i].exp, 32); v[i].output = mk(b[i].out); }
  return v;
}
static std::vector<G2ExpIONative> g2_ios(const void* ios, size_t n) {
  const G2IoBlob* b = (const G2IoBlob*)ios; std::vector<G2ExpIONative> v(n);
  for (size_t i = 0; i < n; i++) { v[i].x = mkg2(b[i].x); v[i].offset = mkg2(b[i].off); memcpy(v[i].exp_val, b[i].exp, 32); v[i].output = mkg2(b[i].out); }
  return v;
}
static std::vector<Fq12ExpIONative> fq12_ios(const void* ios, size_t n) {
  const Fq12IoBlob* b = (const Fq12IoBlob*)ios; std::vector<Fq12ExpIONative> v(n);
  for (size_t i = 0; i < n; i++) { v[i].x = mk12(b[i].x); v[i].offset = mk12(b[i].off); memcpy(v[i].exp_val, b[i].exp, 32); v[i].output = mk12(b[i].out); }
  return v;
}
static std::vector<Fq12ExpU64IONative> fq12u64_ios(const void* ios, size_t n) {
  const Fq12U64IoBlob* b = (const Fq12U64IoBlob*)ios; std::vector<Fq12ExpU64IONative> v(n);
  for (size_t i = 0; i < n; i++) { v[i].x = mk12(b[i].x); v[i].offset = mk12(b[i].off); v[i].exp_val = b[i].exp; v[i].output = mk12(b[i].out); }
  return v;
}
// Trace generation.  out_cols: num_columns x num_rows column-major.  results (optional): per-io chain
// result (G1: x,y as 8 u64), which the caller may copy into the io blob's output field.
int orc_generate_trace(void* p, const void* ios, size_t num_io, u64* out_cols, u64* results) {
  AirHandle* h = (AirHandle*)p;
  try {
    Cols cols;
    if (h->id == AIR_MODULAR) {
      const u64* b = (const u64*)ios; std::vector<std::array<U256, 2>> in(num_io);
      for (size_t i = 0; i < num_io; i++) { in[i][0] = mk(b + 8 * i); in[i][1] = mk(b + 8 * i + 4); }
      cols = static_cast<ModularStark*>(h->air.get())->generate_trace(in);
    } else if (h->id == AIR_G1_MULADD) {   // one record per row: a.x a.y b.x b.y
      const u64* b = (const u64*)ios; std::vector<std::array<G1Point, 2>> in(num_io);
      for (size_t i = 0; i < num_io; i++) { in[i][0] = {mk(b + 16 * i), mk(b + 16 * i + 4)}; in[i][1] = {mk(b + 16 * i + 8), mk(b + 16 * i + 12)}; }
      cols = static_cast<G1Stark*>(h->air.get())->generate_trace(in);
    } else if (h->id == AIR_FQ12_MUL) {    // one record per row: x[12] y[12]
      const u64* b = (const u64*)ios; std::vector<std::array<Fq12Words, 2>> in(num_io);
      for (size_t i = 0; i < num_io; i++) { in[i][0] = mk12(b + 96 * i); in[i][1] = mk12(b + 96 * i + 48); }
      cols = static_cast<Fq12Stark*>(h->air.get())->generate_trace(in);
    } else if (h->id == AIR_G1_EXP) {
      std::vector<G1Point> res;
      cols = static_cast<G1ExpStark*>(h->air.get())->generate_trace(g1_ios(ios, num_io), &res);
      if (results) for (size_t i = 0; i < num_io; i++) { memcpy(results + 8 * i, res[i].x.w, 32); memcpy(results + 8 * i + 4, res[i].y.w, 32); }
    } else if (h->id == AIR_FQ_EXP) {
      std::vector<U256> res;
      cols = static_cast<FqExpStark*>(h->air.get())->generate_trace(fq_ios(ios, num_io), &res);
      if (results) for (size_t i = 0; i < num_io; i++) memcpy(results + 4 * i, res[i].w, 32);
    } else if (h->id == AIR_G2_EXP) {
      std::vector<G2Point> res;
      cols = static_cast<G2ExpStark*>(h->air.get())->generate_trace(g2_ios(ios, num_io), &res);
      if (results) for (size_t i = 0; i < num_io; i++) { memcpy(results + 16 * i, res[i].x0.w, 32); memcpy(results + 16 * i + 4, res[i].x1.w, 32); memcpy(results + 16 * i + 8, res[i].y0.w, 32); memcpy(results + 16 * i + 12, res[i].y1.w, 32); }
    } else if (h->id == AIR_FQ12_EXP || h->id == AIR_FQ12_EXP_U64) {
      std::vector<Fq12Words> res;
      if (h->id == AIR_FQ12_EXP) cols = static_cast<Fq12ExpStark*>(h->air.get())->generate_trace(fq12_ios(ios, num_io), &res);
      else cols = static_cast<Fq12ExpU64Stark*>(h->air.get())->generate_trace(fq12u64_ios(ios, num_io), &res);
      if (results) for (size_t i = 0; i < num_io; i++) for (int k = 0; k < 12; k++) memcpy(results + 48 * i + 4 * k, res[i].c[k].w, 32);
    } else { g_err = "unsupported air"; return -1; }
    size_t n = cols[0].size();
    for (size_t c = 0; c < cols.size(); c++) for (size_t r = 0; r < n; r++) out_cols[c * n + r] = cols[c][r].v;
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -2; }
}
int orc_generate_public_inputs(void* p, const void* ios, size_t num_io, u64* out) {
  AirHandle* h = (AirHandle*)p;
  std::vector<GF> pi;
  if (h->id == AIR_MODULAR || h->id == AIR_G1_MULADD || h->id == AIR_FQ12_MUL) return 0;
  try {
  if (h->id == AIR_G1_EXP) pi = static_cast<G1ExpStark*>(h->air.get())->generate_public_inputs(g1_ios(ios, num_io));
  else if (h->id == AIR_FQ_EXP) pi = static_cast<FqExpStark*>(h->air.get())->generate_public_inputs(fq_ios(ios, num_io));
  else if (h->id == AIR_G2_EXP) pi = static_cast<G2ExpStark*>(h->air.get())->generate_public_inputs(g2_ios(ios, num_io));
  else if (h->id == AIR_FQ12_EXP) pi = static_cast<Fq12ExpStark*>(h->air.get())->generate_public_inputs(fq12_ios(ios, num_io));
  else if (h->id == AIR_FQ12_EXP_U64) pi = static_cast<Fq12ExpU64Stark*>(h->air.get())->generate_public_inputs(fq12u64_ios(ios, num_io));
  else { g_err = "unsupported air"; return -1; }
  } catch (std::exception& e) { g_err = e.what(); return -2; }
  for (size_t i = 0; i < pi.size(); i++) out[i] = pi[i].v;
  return 0;
}
// prove: trace is column-major (num_columns x nrows).  On success *proof_out is malloc'd (free with orc_free).
int orc_prove(void* p, const u64* trace, size_t nrows, const u64* pis, size_t npis, const OrcConfig* c, uint8_t** proof_out, size_t* len_out) {
  AirHandle* h = (AirHandle*)p;
  try {
    select_field(c);
    size_t nc = h->air->num_columns();
    std::vector<std::vector<GF>> cols(nc, std::vector<GF>(nrows));
    for (size_t k = 0; k < nc; k++) for (size_t r = 0; r < nrows; r++) cols[k][r] = GF(trace[k * nrows + r]);
    std::vector<GF> pi(npis); for (size_t i = 0; i < npis; i++) pi[i] = GF(pis[i]);
    g_dbg = ProverDebug();
    Proof proof = prove(*h->air, to_cfg(c), cols, pi, &g_dbg);
    std::vector<uint8_t> bytes = serialize_proof(proof);
    *proof_out = (uint8_t*)malloc(bytes.size()); memcpy(*proof_out, bytes.data(), bytes.size()); *len_out = bytes.size();
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -2; }
}
void orc_free(void* p) { free(p); }
// selects the generator pair for the stage entry points that take no config (orc_fft, orc_commit_columns, orc_root_of_unity)
int orc_select_field(u64 coset_shift) {
  try { OrcConfig c; memset(&c, 0, sizeof c); c.coset_shift = coset_shift; select_field(&c); return 0; } catch (std::exception& e) { g_err = e.what(); return -1; }
}
// verify: 0 = accepted, 1 = rejected (reason in orc_last_error), <0 = malformed input
int orc_verify(void* p, const uint8_t* proof, size_t len, const OrcConfig* c) {
  AirHandle* h = (AirHandle*)p;
  try {
    select_field(c);
    Proof pr = deserialize_proof(proof, len);
    std::string r = verify_stark_proof(*h->air, pr, to_cfg(c));
    if (r.empty()) return 0;
    g_err = r; return 1;
  } catch (std::exception& e) { g_err = e.what(); return -2; }
}
// Evaluate the AIR constraints on the trace rows themselves (no LDE): returns the number of violated
// (row, constraint) pairs and writes the first one to first_bad[0..1]; num_constraints_out = constraints per row.
long orc_check_trace(void* p, const u64* trace, size_t nrows, const u64* pis, size_t npis, long* first_bad, size_t* num_constraints_out) {
  AirHandle* h = (AirHandle*)p;
  size_t nc = h->air->num_columns();
  std::vector<GF> pi(npis); for (size_t i = 0; i < npis; i++) pi[i] = GF(pis[i]);
  long bad = 0; first_bad[0] = first_bad[1] = -1;
  std::vector<GF> lv(nc), nv(nc);
  for (size_t r = 0; r < nrows; r++) {
    size_t rn = (r + 1) % nrows;
    for (size_t k = 0; k < nc; k++) { lv[k] = GF(trace[k * nrows + r]); nv[k] = GF(trace[k * nrows + rn]); }
    Consumer<GF> yc({}, r + 1 == nrows ? GF() : GF(1), r == 0 ? GF(1) : GF(), r + 1 == nrows ? GF(1) : GF());
    std::vector<GF> log; yc.log = &log;
    h->air->eval(lv.data(), nv.data(), pi.data(), yc);
    if (num_constraints_out) *num_constraints_out = log.size();
    for (size_t k = 0; k < log.size(); k++) if (log[k].v) { if (!bad) { first_bad[0] = (long)r; first_bad[1] = (long)k; } bad++; }
  }
  return bad;
}
// reference src/utils/lookup.rs:60-111 `permuted_cols`
void orc_permuted_cols(const u64* inputs, const u64* table, size_t n, u64* sorted_out, u64* perm_out) {
  std::vector<GF> in(n), tb(n), so, pe;
  for (size_t i = 0; i < n; i++) { in[i] = GF(inputs[i]); tb[i] = GF(table[i]); }
  permuted_cols(in, tb, so, pe);
  for (size_t i = 0; i < n; i++) { sorted_out[i] = so[i].v; perm_out[i] = pe[i].v; }
}
// AIR constraints folded by the consumer at one evaluation point (base field), for unit parity tests.
int orc_eval_constraints(void* p, const u64* lv, const u64* nv, const u64* pis, size_t npis, const u64* alphas, size_t nalpha, u64 z_last, u64 l_first,
                         u64 l_last, u64* out, size_t* count) {
  AirHandle* h = (AirHandle*)p;
  size_t nc = h->air->num_columns();
  std::vector<GF> l(nc), n(nc), pi(npis), al(nalpha);
  for (size_t i = 0; i < nc; i++) { l[i] = GF(lv[i]); n[i] = GF(nv[i]); }
  for (size_t i = 0; i < npis; i++) pi[i] = GF(pis[i]);
  for (size_t i = 0; i < nalpha; i++) al[i] = GF(alphas[i]);
  Consumer<GF> yc(al, GF(z_last), GF(l_first), GF(l_last));
  h->air->eval(l.data(), n.data(), pi.data(), yc);
  for (size_t i = 0; i < nalpha; i++) out[i] = yc.accs[i].v;
  if (count) *count = yc.count;
  return 0;
}
// Bounded-sample timing (see sample.hpp): out_ms = {tracegen, commit, zpoly, quotient, openings, reduce, fri_tail}, each measured on
// 1/2^shift of its work (FRI tail in full); the caller scales the sampled phases by 2^shift.
int orc_time_sample(void* p, const void* ios, size_t num_io, const OrcConfig* c, int shift, double* out_ms) {
  AirHandle* h = (AirHandle*)p;
  try {
    size_t nrows = orc_air_num_rows(p);
    SampleTimes t = time_prove_sample(*h->air, nrows, to_cfg(c), shift);
    if (h->id == AIR_G1_EXP && ios) t.tracegen_ms = time_g1_tracegen_sample(*static_cast<G1ExpStark*>(h->air.get()), g1_ios(ios, num_io), shift);
    if (h->id == AIR_G2_EXP && ios) {
      auto* a = static_cast<G2ExpStark*>(h->air.get()); auto v = g2_ios(ios, num_io); G2Point r;
      t.tracegen_ms = time_exp_tracegen_sample([&](size_t k) { a->generate_trace_for_one_block(v[k].x, v[k].offset, v[k].exp_val, &r); }, num_io, 512, a->num_range_check_cols, false, shift);
    }
    if (h->id == AIR_FQ_EXP && ios) {
      auto* a = static_cast<FqExpStark*>(h->air.get()); auto v = fq_ios(ios, num_io); U256 r;
      t.tracegen_ms = time_exp_tracegen_sample([&](size_t k) { a->generate_trace_for_one_block(v[k].x, v[k].offset, v[k].exp_val, &r); }, num_io, 512, a->num_range_check_cols, false, shift);
    }
    if (h->id == AIR_FQ12_EXP && ios) {
      auto* a = static_cast<Fq12ExpStark*>(h->air.get()); auto v = fq12_ios(ios, num_io); Fq12Words r;
      t.tracegen_ms = time_exp_tracegen_sample([&](size_t k) { a->generate_trace_for_one_block(v[k].x, v[k].offset, v[k].exp_val, &r); }, num_io, 512, a->num_range_check_cols, true, shift);
    }
    out_ms[0] = t.tracegen_ms; out_ms[1] = t.commit_ms; out_ms[2] = t.zpoly_ms; out_ms[3] = t.quotient_ms; out_ms[4] = t.openings_ms; out_ms[5] = t.reduce_ms; out_ms[6] = t.fri_ms;
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -2; }
}
// ---- intermediates of the last orc_prove call (for stage-by-stage parity tests) ----
size_t orc_dbg_num_z() { return g_dbg.z_polys.size(); }
void orc_dbg_z_polys(u64* out) { size_t n = g_dbg.z_polys.empty() ? 0 : g_dbg.z_polys[0].size(); for (size_t c = 0; c < g_dbg.z_polys.size(); c++) for (size_t i = 0; i < n; i++) out[c * n + i] = g_dbg.z_polys[c][i].v; }
void orc_dbg_quotient_chunks(u64* out) { size_t n = g_dbg.quotient_chunks.empty() ? 0 : g_dbg.quotient_chunks[0].size(); for (size_t c = 0; c < g_dbg.quotient_chunks.size(); c++) for (size_t i = 0; i < n; i++) out[c * n + i] = g_dbg.quotient_chunks[c][i].v; }
// challenges: [alphas(nc)] [zeta a,b] [fri_alpha a,b] [perm sets: for chal, for slot: beta,gamma]
size_t orc_dbg_challenges(u64* out, size_t cap) {
  std::vector<u64> v; for (auto a : g_dbg.alphas) v.push_back(a.v);
  v.push_back(g_dbg.zeta.a.v); v.push_back(g_dbg.zeta.b.v); v.push_back(g_dbg.fri_alpha.a.v); v.push_back(g_dbg.fri_alpha.b.v);
  for (auto& s : g_dbg.perm_sets) for (auto& ch : s) { v.push_back(ch.beta.v); v.push_back(ch.gamma.v); }
  for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
  return v.size();
}
size_t orc_dbg_timings(char* out, size_t cap) {
  std::ostringstream os; os << "{"; bool first = true;
  for (auto& kv : g_dbg.timings_ms) { if (!first) os << ","; first = false; os << "\"" << kv.first << "\":" << kv.second; }
  os << "}"; std::string s = os.str(); if (cap) { snprintf(out, cap, "%s", s.c_str()); } return s.size();
}
}
