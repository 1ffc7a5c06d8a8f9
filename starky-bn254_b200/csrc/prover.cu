// Host orchestration of `starky::prover::prove` (external dependency of the reference; SURVEY.md
// sections 3.3 and App. B) over the CUDA kernel families K2-K6, plus the extern "C" boundary declared in
// include/starky_bn254_b200.h.  The Fiat-Shamir challenger runs on the host (a few dozen Poseidon
// permutations per proof); everything that touches trace-sized data runs on the device.
#include "../../include/starky_bn254_b200.h"
#include "common.cuh"
#include "air.cuh"
#include "ntt.cuh"
#include "merkle.cuh"
#include "quotient.cuh"
#include "zpoly.cuh"
#include "fri.cuh"
#include "tracegen.cuh"
#include "poseidon.cuh"
#include <atomic>
#include <cmath>
#include <map>
#include <memory>
#include <sstream>
#include <thread>

// ------------------------------------------------------------------------------------------------
// plonky2::iop::challenger::Challenger: duplex sponge in overwrite mode; challenges
// are popped from the END of the 8-element output buffer.
struct Challenger {
  u64 st[12]; std::vector<u64> in, out;
  Challenger() { memset(st, 0, sizeof st); }
  void duplex() {
    for (size_t i = 0; i < in.size(); i++) st[i] = in[i];
    in.clear();
    poseidon_permute(st);
    out.assign(st, st + 8);
  }
  void observe(u64 x) { out.clear(); in.push_back(x); if (in.size() == 8) duplex(); }
  void observe_n(const u64* p, size_t n) { for (size_t i = 0; i < n; i++) observe(p[i]); }
  u64 get() { if (!in.empty() || out.empty()) duplex(); u64 r = out.back(); out.pop_back(); return r; }
  gl2 get_ext() { u64 a = get(); u64 b = get(); return gl2_make(a, b); }
};

struct PhaseTimer {  // CUDA-event timings of the prover phases (names follow plonky2's `timed!` labels)
  sbn_ctx* ctx; std::vector<std::pair<std::string, cudaEvent_t>> ev;
  explicit PhaseTimer(sbn_ctx* c) : ctx(c) { mark("start"); }
  void mark(const char* name) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, ctx->stream); ev.push_back({name, e}); }
  ~PhaseTimer() { for (auto& e : ev) cudaEventDestroy(e.second); }   // an exception between marks must not leak the events
  std::string json() {
    cudaEventSynchronize(ev.back().second);
    std::ostringstream os; os << "{";
    for (size_t i = 1; i < ev.size(); i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second); os << (i > 1 ? "," : "") << "\"" << ev[i].first << "\":" << ms; }
    float tot = 0; cudaEventElapsedTime(&tot, ev.front().second, ev.back().second);
    os << ",\"total\":" << tot << "}";
    for (auto& e : ev) cudaEventDestroy(e.second);
    ev.clear();
    return os.str();
  }
};

struct sbn_trace {
  sbn_ctx* ctx; AirDesc air; int logn; DevBuf<u64> cols; std::vector<u64> results;
};
struct sbn_proof {
  std::vector<uint8_t> bytes; std::string timings;
  std::vector<u64> dbg_z, dbg_q, dbg_ch;
};

// Intra-proof sharding over world = 2^m ranks: rank r commits to, evaluates the quotient on and
// answers queries for the LDE class rho = bitrev_m(r) (natural LDE indices i = rho mod 2^m), which is exactly the set of leaves
// under Merkle cap entries [r 2^(cap_height - m), (r + 1) 2^(cap_height - m)).  Everything else (trace, coefficients, Z, openings,
// FRI layers, transcript) is replicated and deterministic, so the only exchanges are all-gathers of cap digests, of the
// quotient values (2 N num_challenges field elements in total) and of the opened rows with their paths.
struct Shard {
  int rank = 0, world = 1, m = 0; sbn_allgather_fn allgather = nullptr; void* user = nullptr; sbn_allgather_fn allgather_device = nullptr;
  bool on() const { return world > 1; }
  u32 rho() const { return bitrev32((u32)rank, m); }
  void gather(const void* send, size_t nbytes, std::vector<uint8_t>& recv) const {
    recv.resize(nbytes * world);
    int rc = allgather(user, send, nbytes, recv.data());
    if (rc != 0) throw SbnError(SBN_ERR_INTERNAL, "sharded prove: the all-gather callback failed");
  }
  void gather_device(sbn_ctx* ctx, const void* d_send, size_t nbytes, void* d_recv) const {   // device buffers, stream-ordered on both sides
    ctx->sync();
    if (allgather_device(user, d_send, nbytes, d_recv) != 0) throw SbnError(SBN_ERR_INTERNAL, "sharded prove: the device all-gather callback failed");
  }
};

// Commitment to a batch of polynomials (plonky2 `PolynomialBatch`, blinding = false).  Sharded: `lde` / `tree` hold this rank's
// class only ([col][2N / world], local leaf order), `lde_next` the class the quotient's "next" rows come from when world > 2.
struct Commitment {
  DevBuf<u64> coeffs, lde, lde_next; DevMerkleTree tree; int ncols = 0;
  std::vector<u64> cap;   // the full cap (2^cap_height x 4), identical on every rank
  bool streamed = false;  // the LDE was hashed sub-coset by sub-coset and dropped: `lde` is empty, sub-cosets are recomputed on demand
};
// Streamed commitment (config 5 at 2^22 rows: trace + coefficients + LDE of 812 + 444 columns do not fit 180 GB): every sub-coset
// b of the LDE is evaluated into `scratch` ([ncols][N]), its N leaves are hashed, and the values are dropped; the tree (all digest
// levels) is kept.  The footprint is independent of rate_bits; the same cap, leaf for leaf, as commit_from_coeffs.
static void commit_streamed(sbn_ctx* ctx, Commitment& c, int ncols, int logn, int rate_bits, int cap_height, u64* scratch) {
  c.ncols = ncols; c.streamed = true;
  merkle_alloc(ctx, &c.tree, (size_t(1) << logn) << rate_bits, cap_height);
  for (int b = 0; b < (1 << rate_bits); b++) {
    lde_sub_coset(ctx, c.coeffs, scratch, ncols, logn, rate_bits, b);
    merkle_leaf_hash_sub_coset(ctx, scratch, ncols, logn, rate_bits, b, &c.tree);
  }
  merkle_build_from_leaf_digests(ctx, &c.tree);
  c.cap = c.tree.cap;
}
// Streaming is chosen when the resident footprint of the plain prover (coefficients + LDE of the trace and Z batches + Z values,
// beside the trace the caller holds) would not fit the device; SBN_STREAMING=0/1 forces the choice (tests: byte-identical proofs).
static bool want_streaming(size_t C, size_t Z, size_t N, int rate_bits) {
  if (const char* e = getenv("SBN_STREAMING")) return atoi(e) != 0;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
  const double need = 8.0 * (double)N * ((double)(C + Z) * (1.0 + (double)(1 << rate_bits)) + (double)Z) * 1.1;
  return need > 0.8 * (double)total_b - 8.0 * (double)N * (double)C;
}
static void commit_from_coeffs(sbn_ctx* ctx, const Shard& sh, Commitment& c, int ncols, int logn, int rate_bits, int cap_height, bool need_next) {
  size_t N = size_t(1) << logn;
  c.ncols = ncols;
  if (!sh.on()) {
    c.lde = DevBuf<u64>(ctx, (size_t)ncols * (N << rate_bits));
    lde_columns(ctx, c.coeffs, c.lde, ncols, logn, rate_bits);
    merkle_commit_lde(ctx, c.lde, ncols, logn, rate_bits, cap_height, &c.tree);
    c.cap = c.tree.cap;
    return;
  }
  const size_t Lp = (N << rate_bits) >> sh.m;
  c.lde = DevBuf<u64>(ctx, (size_t)ncols * Lp);
  lde_class(ctx, c.coeffs, c.lde, ncols, logn, rate_bits, sh.m, sh.rho());
  if (need_next && sh.world > 2 && rate_bits == 1) {   // rate 1: the quotient coset is the LDE coset, "next" rows = class rho + 2
    c.lde_next = DevBuf<u64>(ctx, (size_t)ncols * Lp);
    lde_class(ctx, c.coeffs, c.lde_next, ncols, logn, rate_bits, sh.m, (sh.rho() + 2) & (sh.world - 1));
  }
  // the class is laid out like a small LDE: [col][b'][k'] with B' = max(1, R / world) sub-cosets (lde_class)
  const int l_logn = sh.m <= rate_bits ? logn : logn + rate_bits - sh.m, l_rate = sh.m <= rate_bits ? rate_bits - sh.m : 0;
  merkle_commit_lde(ctx, c.lde, ncols, l_logn, l_rate, cap_height - sh.m, &c.tree);
  std::vector<uint8_t> all;
  sh.gather(c.tree.cap.data(), c.tree.cap.size() * 8, all);
  c.cap.resize(all.size() / 8);
  memcpy(c.cap.data(), all.data(), all.size());
}
static void commit_from_values(sbn_ctx* ctx, const Shard& sh, Commitment& c, const u64* values, int ncols, int logn, int rate_bits, int cap_height, bool need_next) {
  size_t N = size_t(1) << logn;
  c.coeffs = DevBuf<u64>(ctx, (size_t)ncols * N);
  intt_columns(ctx, values, c.coeffs, ncols, logn);
  commit_from_coeffs(ctx, sh, c, ncols, logn, rate_bits, cap_height, need_next);
}

struct Writer {  // canonical proof wire format: LE u64 field elements, u32 length prefixes, u8 option tag
  std::vector<uint8_t>& b;
  explicit Writer(std::vector<uint8_t>& v) : b(v) {}
  void u8(uint8_t x) { b.push_back(x); }
  void u32_(uint32_t x) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(x >> (8 * i))); }
  void f(u64 x) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(x >> (8 * i))); }
  void fs(const u64* p, size_t n) { size_t o = b.size(); b.resize(o + 8 * n); memcpy(b.data() + o, p, 8 * n); }  // little-endian host
  void hashes(const u64* p, size_t nh) { u32_((uint32_t)nh); fs(p, 4 * nh); }
  void fvec(const u64* p, size_t n) { u32_((uint32_t)n); fs(p, n); }
  void evec(const u64* p, size_t n) { u32_((uint32_t)n); fs(p, 2 * n); }
};

static std::vector<int> reduction_arity_bits(const sbn_config& c, int degree_bits) {  // FriReductionStrategy::ConstantArityBits
  std::vector<int> r;
  while (degree_bits > (int)c.fri_final_poly_bits && degree_bits + (int)c.rate_bits - (int)c.fri_arity_bits >= (int)c.cap_height) {
    r.push_back(c.fri_arity_bits); degree_bits -= c.fri_arity_bits;
  }
  return r;
}

// quotient values of all classes, parts[rank][challenge][t] at quotient-coset index iq = bitrev_m(rank) + (t << m), into the
// unsharded layout full[challenge][iq & 1][iq >> 1]
__global__ void k_scatter_classes(const u64* __restrict__ parts, u64* __restrict__ full, int m, int nch, int logn) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const int logMp = logn + 1 - m;
  const size_t Mp = size_t(1) << logMp, N = size_t(1) << logn;
  if (i >= ((size_t)nch << (logn + 1))) return;
  const size_t t = i & (Mp - 1), rc = i >> logMp;
  const int c = (int)(rc % nch), r = (int)(rc / nch);
  const size_t iq = bitrev32((u32)r, m) + (t << m);
  full[(size_t)c * 2 * N + (iq & 1) * N + (iq >> 1)] = parts[i];
}

static void prove_impl(sbn_ctx* ctx, const sbn_config& cfg, const sbn_trace* tr, const u64* public_inputs, size_t npis, const Shard& sh, sbn_proof* proof) {
  const AirDesc& air = tr->air;
  const int logn = tr->logn, rate_bits = cfg.rate_bits, cap_height = cfg.cap_height, nch = cfg.num_challenges;
  const size_t N = size_t(1) << logn, L = N << rate_bits;
  SBN_REQUIRE(npis == air.num_public_inputs, "public input count mismatch");
  SBN_REQUIRE(cfg.coset_shift == 0 ? ctx->coset_shift() == GL_MULT_GENERATOR : cfg.coset_shift == ctx->coset_shift(), "internal: generator pair not selected");
  SBN_REQUIRE(nch >= 1 && nch <= SBN_MAX_CHALLENGES, "unsupported num_challenges");
  SBN_REQUIRE(cfg.fri_degree_hack <= 1 && cfg.reserved == 0, "bad config flags");
  SBN_REQUIRE(logn + rate_bits <= 30 && rate_bits >= 1 && rate_bits <= 4, "unsupported trace size / rate");
  SBN_REQUIRE(cap_height <= logn + rate_bits && cfg.pow_bits >= 1, "bad FRI configuration");
  std::vector<int> arities = reduction_arity_bits(cfg, logn);
  int total_arities = 0; for (int a : arities) total_arities += a;
  SBN_REQUIRE(total_arities <= logn + rate_bits - cap_height, "FRI total reduction arity is too large.");
  if (sh.on()) {
    SBN_REQUIRE(sh.allgather && (1 << sh.m) == sh.world && sh.rank >= 0 && sh.rank < sh.world, "sharded prove: world must be a power of two and the all-gather callback set");
    SBN_REQUIRE(sh.m <= cap_height && sh.m < logn, "sharded prove: world must not exceed 2^cap_height");
  }
  for (size_t i = 0; i < npis; i++) SBN_REQUIRE(public_inputs[i] < GL_P, "public input is not a canonical field element");
  PhaseTimer tm(ctx);
  Writer w(proof->bytes);

  // ---- trace commitment ----
  Commitment trace_c;
  const size_t nz_cols = air.perm_pairs.empty() ? 0 : (air.perm_pairs.size() * nch + air.quotient_degree_factor() - 1) / air.quotient_degree_factor();
  const bool streaming = !sh.on() && want_streaming(air.num_columns, nz_cols, N, rate_bits);
  DevBuf<u64> scratch_t, scratch_z;   // one sub-coset of the trace / Z batch (streaming only)
  if (streaming) {
    trace_c.coeffs = DevBuf<u64>(ctx, air.num_columns * N);
    intt_columns(ctx, tr->cols, trace_c.coeffs, (int)air.num_columns, logn);
    scratch_t = DevBuf<u64>(ctx, air.num_columns * N);
    commit_streamed(ctx, trace_c, (int)air.num_columns, logn, rate_bits, cap_height, scratch_t);
  } else
  commit_from_values(ctx, sh, trace_c, tr->cols, (int)air.num_columns, logn, rate_bits, cap_height, true);
  tm.mark("compute trace commitment");
  Challenger ch;
  ch.observe_n(trace_c.cap.data(), trace_c.cap.size());
  w.hashes(trace_c.cap.data(), trace_c.cap.size() / 4);

  // ---- permutation argument ----
  const int qdf = air.quotient_degree_factor();
  PermInstances perm;
  Commitment z_c;
  bool uses_perm = !air.perm_pairs.empty();
  std::vector<u64> perm_challenges;  // [chal][slot] (beta, gamma)
  if (uses_perm) {
    const int batch = qdf;  // Stark::permutation_batch_size
    perm_challenges.resize((size_t)nch * batch * 2);
    for (int i = 0; i < nch; i++) for (int j = 0; j < batch; j++) { perm_challenges[(i * batch + j) * 2] = ch.get(); perm_challenges[(i * batch + j) * 2 + 1] = ch.get(); }
    // get_permutation_batches: cartesian(pairs, 0..num_challenges) chunked by batch; slot s of a chunk uses sets[chal].challenges[s]
    perm.batch_size = batch;
    size_t ninst = air.perm_pairs.size() * nch, nz = (ninst + batch - 1) / batch;
    perm.lhs.assign(nz * batch, 0xFFFFFFFFu); perm.rhs.assign(nz * batch, 0xFFFFFFFFu); perm.gamma.assign(nz * batch, 0); perm.count.assign(nz, 0);
    size_t e = 0;
    for (auto& pr : air.perm_pairs) for (int chal = 0; chal < nch; chal++, e++) {
      size_t z = e / batch; int slot = (int)(e % batch);
      perm.lhs[z * batch + slot] = pr.first; perm.rhs[z * batch + slot] = pr.second;
      perm.gamma[z * batch + slot] = perm_challenges[(chal * batch + slot) * 2 + 1];
      perm.count[z]++;
    }
    DevBuf<u64> zvals(ctx, nz * N);
    compute_z_polys(ctx, tr->cols, logn, perm, zvals);
    tm.mark("compute permutation Z polys");
    if (getenv("SBN_DEBUG_INTERMEDIATES")) {
      proof->dbg_z.resize(nz * N);
      CUDA_CHECK(cudaMemcpyAsync(proof->dbg_z.data(), zvals, nz * N * 8, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->sync();
    }
    if (streaming) {
      z_c.coeffs = DevBuf<u64>(ctx, nz * N);
      intt_columns(ctx, zvals, z_c.coeffs, (int)nz, logn);
      zvals.reset();                          // its block is reused for the sub-coset scratch
      scratch_z = DevBuf<u64>(ctx, nz * N);
      commit_streamed(ctx, z_c, (int)nz, logn, rate_bits, cap_height, scratch_z);
    } else
    commit_from_values(ctx, sh, z_c, zvals, (int)nz, logn, rate_bits, cap_height, true);
    tm.mark("compute permutation Z commitments");
    ch.observe_n(z_c.cap.data(), z_c.cap.size());
    w.u8(1); w.hashes(z_c.cap.data(), z_c.cap.size() / 4);
  } else {
    w.u8(0);
  }

  // ---- quotient ----
  u64 alphas[SBN_MAX_CHALLENGES] = {0, 0};
  for (int i = 0; i < nch; i++) alphas[i] = ch.get();
  DevBuf<u64> d_pis(ctx, npis ? npis : 1);
  if (npis) ctx->upload(d_pis, public_inputs, npis * 8);
  Commitment q_c;
  const int nq_polys = qdf * nch;
  q_c.coeffs = DevBuf<u64>(ctx, (size_t)nq_polys * N);
  if (streaming) {
    // the two half-cosets of the quotient coset one after the other, each evaluated from the coefficients into the scratch
    // buffers (exactly the classes a world-2 sharded proof gives its two ranks)
    const size_t Mp = N;
    DevBuf<u64> acc_local(ctx, (size_t)SBN_MAX_CHALLENGES * Mp), acc_full(ctx, (size_t)nch * 2 * N);
    for (u32 bq = 0; bq < 2; bq++) {
      QDomain dom; dom.m = 1; dom.sigma = bq;
      lde_class(ctx, trace_c.coeffs, scratch_t, trace_c.ncols, logn, 1, 1, bq);
      dom.trace = dom.trace_next = scratch_t;
      if (uses_perm) { lde_class(ctx, z_c.coeffs, scratch_z, z_c.ncols, logn, 1, 1, bq); dom.zs = dom.zs_next = scratch_z; }
      quotient_eval(ctx, air, dom, nullptr, nullptr, perm, d_pis, alphas, nch, logn, rate_bits, acc_local);
      for (int c = 0; c < nch; c++)
        CUDA_CHECK(cudaMemcpyAsync(acc_full + ((size_t)c * 2 + bq) * N, acc_local + (size_t)c * Mp, N * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    quotient_finish(ctx, acc_full, nch, logn, q_c.coeffs);
  } else if (!sh.on()) {
    compute_quotient_chunks(ctx, air, trace_c.lde, uses_perm ? z_c.lde.get() : nullptr, perm, d_pis, alphas, nch, logn, rate_bits, q_c.coeffs);
  } else {
    // this rank's class of the quotient coset (rate_bits = 1: the quotient coset is the LDE coset), then all classes -> [chal][bq][k]
    QDomain dom; dom.m = sh.m; dom.sigma = sh.rho();
    const size_t Mp = quotient_points(dom, logn);
    DevBuf<u64> qt, qtn, qz, qzn;
    if (rate_bits == 1) {   // the quotient coset is the LDE coset: the committed class batches are the quotient's inputs
      dom.trace = trace_c.lde; dom.trace_next = sh.world > 2 ? trace_c.lde_next.get() : trace_c.lde.get();
      if (uses_perm) { dom.zs = z_c.lde; dom.zs_next = sh.world > 2 ? z_c.lde_next.get() : z_c.lde.get(); }
    } else {                // higher rates: the quotient coset is a sub-coset of the LDE coset that other ranks committed to;
                            // evaluate this rank's class of it (and the class its "next" rows lie in) from the coefficients
      const u32 nxt = (sh.rho() + 2) & (sh.world - 1);
      qt = DevBuf<u64>(ctx, (size_t)trace_c.ncols * Mp);
      lde_class(ctx, trace_c.coeffs, qt, trace_c.ncols, logn, 1, sh.m, sh.rho());
      if (sh.world > 2) { qtn = DevBuf<u64>(ctx, (size_t)trace_c.ncols * Mp); lde_class(ctx, trace_c.coeffs, qtn, trace_c.ncols, logn, 1, sh.m, nxt); }
      dom.trace = qt; dom.trace_next = sh.world > 2 ? qtn.get() : qt.get();
      if (uses_perm) {
        qz = DevBuf<u64>(ctx, (size_t)z_c.ncols * Mp);
        lde_class(ctx, z_c.coeffs, qz, z_c.ncols, logn, 1, sh.m, sh.rho());
        if (sh.world > 2) { qzn = DevBuf<u64>(ctx, (size_t)z_c.ncols * Mp); lde_class(ctx, z_c.coeffs, qzn, z_c.ncols, logn, 1, sh.m, nxt); }
        dom.zs = qz; dom.zs_next = sh.world > 2 ? qzn.get() : qz.get();
      }
    }
    DevBuf<u64> acc_local(ctx, (size_t)SBN_MAX_CHALLENGES * Mp), acc_full(ctx, (size_t)nch * 2 * N);
    quotient_eval(ctx, air, dom, nullptr, nullptr, perm, d_pis, alphas, nch, logn, rate_bits, acc_local);
    if (sh.allgather_device) {   // device to device over NVLink, then one scatter kernel
      DevBuf<u64> parts(ctx, (size_t)sh.world * nch * Mp);
      sh.gather_device(ctx, acc_local, (size_t)nch * Mp * 8, parts);
      const size_t tot = (size_t)sh.world * nch * Mp;
      k_scatter_classes<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(parts, acc_full, sh.m, nch, logn);
      LAUNCH_CHECK(ctx);
    } else {
      std::vector<u64> mine((size_t)nch * Mp), full((size_t)nch * 2 * N);
      CUDA_CHECK(cudaMemcpyAsync(mine.data(), acc_local, mine.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
      ctx->sync();
      std::vector<uint8_t> all;
      sh.gather(mine.data(), mine.size() * 8, all);
      const u64* parts = reinterpret_cast<const u64*>(all.data());
      for (int r = 0; r < sh.world; r++) {
        const size_t sigma = bitrev32((u32)r, sh.m);
        for (int c = 0; c < nch; c++) {
          const u64* src = parts + ((size_t)r * nch + c) * Mp;
          u64* dst = full.data() + (size_t)c * 2 * N;
          for (size_t t = 0; t < Mp; t++) { const size_t iq = sigma + ((size_t)t << sh.m); dst[(iq & 1) * N + (iq >> 1)] = src[t]; }
        }
      }
      CUDA_CHECK(cudaMemcpyAsync(acc_full, full.data(), full.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
      ctx->sync();
    }
    quotient_finish(ctx, acc_full, nch, logn, q_c.coeffs);
  }
  tm.mark("compute quotient polys");
  if (getenv("SBN_DEBUG_INTERMEDIATES")) {
    proof->dbg_q.resize((size_t)nq_polys * N);
    CUDA_CHECK(cudaMemcpyAsync(proof->dbg_q.data(), q_c.coeffs, proof->dbg_q.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
  }
  commit_from_coeffs(ctx, sh, q_c, nq_polys, logn, rate_bits, cap_height, false);
  tm.mark("compute quotient commitment");
  ch.observe_n(q_c.cap.data(), q_c.cap.size());
  w.hashes(q_c.cap.data(), q_c.cap.size() / 4);

  // ---- openings ----
  gl2 zeta = ch.get_ext();
  u64 g = ctx->root_of_unity(logn);
  if (gl2_eq(gl2_pow(zeta, (u64)N), gl2_make(1, 0))) throw SbnError(SBN_ERR_INTERNAL, "Opening point is in the subgroup.");
  gl2 zeta_next = gl2_mul_base(zeta, g);
  const int C = (int)air.num_columns, Z = uses_perm ? z_c.ncols : 0;
  std::vector<u64> open((size_t)(C + Z + nq_polys) * 4);
  DevBuf<u64> zpw = two_point_power_table(ctx, logn, zeta, zeta_next);   // zeta^j, (g zeta)^j: shared by the openings and the FRI quotients
  if (!sh.on()) {
    DevBuf<u64> d_open(ctx, (size_t)(C + Z + nq_polys) * 4);
    eval_columns_at_two_points(ctx, trace_c.coeffs, C, logn, zpw, d_open);
    if (Z) eval_columns_at_two_points(ctx, z_c.coeffs, Z, logn, zpw, d_open + (size_t)C * 4);
    eval_columns_at_two_points(ctx, q_c.coeffs, nq_polys, logn, zpw, d_open + (size_t)(C + Z) * 4);
    CUDA_CHECK(cudaMemcpyAsync(open.data(), d_open, open.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
  } else {
    // every rank evaluates its slice of the columns of each commitment; the values (4 words per column) are all-gathered
    const u64* cf[3] = {trace_c.coeffs.get(), Z ? z_c.coeffs.get() : nullptr, q_c.coeffs.get()};
    const int ncs[3] = {C, Z, nq_polys};
    int per[3], tot_per = 0;
    for (int o = 0; o < 3; o++) { per[o] = (ncs[o] + sh.world - 1) / sh.world; tot_per += per[o]; }
    DevBuf<u64> d_mine(ctx, (size_t)tot_per * 4);
    CUDA_CHECK(cudaMemsetAsync(d_mine, 0, (size_t)tot_per * 32, ctx->stream));
    for (int o = 0, at = 0; o < 3; at += per[o], o++) {
      const int c0 = std::min(ncs[o], sh.rank * per[o]), nc = std::min(ncs[o], c0 + per[o]) - c0;
      if (nc > 0) eval_columns_at_two_points(ctx, cf[o] + (size_t)c0 * N, nc, logn, zpw, d_mine + (size_t)at * 4);
    }
    std::vector<u64> mine((size_t)tot_per * 4);
    CUDA_CHECK(cudaMemcpyAsync(mine.data(), d_mine, mine.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    std::vector<uint8_t> all;
    sh.gather(mine.data(), mine.size() * 8, all);
    const u64* parts = reinterpret_cast<const u64*>(all.data());
    for (int o = 0, at = 0, first = 0; o < 3; at += per[o], first += ncs[o], o++)
      for (int c = 0; c < ncs[o]; c++) memcpy(open.data() + (size_t)(first + c) * 4, parts + ((size_t)(c / per[o]) * tot_per + at + c % per[o]) * 4, 32);
  }
  // StarkOpeningSet { local_values, next_values, permutation_zs, permutation_zs_next, quotient_polys }
  auto pick = [&](int first, int count, int which) { std::vector<u64> v((size_t)count * 2); for (int i = 0; i < count; i++) { v[2 * i] = open[(size_t)(first + i) * 4 + 2 * which]; v[2 * i + 1] = open[(size_t)(first + i) * 4 + 2 * which + 1]; } return v; };
  std::vector<u64> local_values = pick(0, C, 0), next_values = pick(0, C, 1), pzs = pick(C, Z, 0), pzs_next = pick(C, Z, 1), qp = pick(C + Z, nq_polys, 0);
  w.evec(local_values.data(), C); w.evec(next_values.data(), C);
  if (uses_perm) { w.evec(pzs.data(), Z); w.evec(pzs_next.data(), Z); }
  w.evec(qp.data(), nq_polys);
  // challenger.observe_openings(to_fri_openings): zeta batch = local | zs | quotient, zeta_next batch = next | zs_next
  ch.observe_n(local_values.data(), local_values.size()); ch.observe_n(pzs.data(), pzs.size()); ch.observe_n(qp.data(), qp.size());
  ch.observe_n(next_values.data(), next_values.size()); ch.observe_n(pzs_next.data(), pzs_next.size());
  tm.mark("compute openings");

  // ---- FRI ----
  gl2 fri_alpha = ch.get_ext();
  std::vector<OracleView> views; views.push_back({trace_c.coeffs, C}); if (Z) views.push_back({z_c.coeffs, Z}); views.push_back({q_c.coeffs, nq_polys});
  const int logL = logn + rate_bits;
  DevBuf<u64> fcoeffs(ctx, 2 * L), fvalues(ctx, 2 * L);
  ColumnSplit split{sh.rank, sh.world, [&](const void* a, size_t n, void* b) { sh.gather_device(ctx, a, n, b); }};
  fri_final_poly(ctx, views, logn, rate_bits, fri_alpha, zeta, zeta_next, fcoeffs, sh.on() && sh.allgather_device ? &split : nullptr, zpw);
  if (cfg.fri_degree_hack) {   // U3: multiply the FRI polynomial by X (its top coefficient is zero: the batch quotients were padded)
    DevBuf<u64> shifted(ctx, 2 * L);
    CUDA_CHECK(cudaMemsetAsync(shifted, 0, 2 * L * 8, ctx->stream));
    for (int comp = 0; comp < 2; comp++)
      CUDA_CHECK(cudaMemcpyAsync(shifted + (size_t)comp * L + 1, fcoeffs + (size_t)comp * L, (N - 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    fcoeffs = std::move(shifted);
  }
  tm.mark("reduce batch of polynomials");
  u64 shift = ctx->coset_shift();
  ntt_batch(ctx, fcoeffs, L, fvalues, L, 2, logL, false, shift, nullptr);
  tm.mark("perform final FFT");
  std::vector<FriLayer> layers(arities.size());
  std::vector<gl2> betas;
  int cur_log = logL;
  DevBuf<u64> cur_coeffs = std::move(fcoeffs), cur_values = std::move(fvalues);
  w.u32_((uint32_t)arities.size());
  for (size_t li = 0; li < arities.size(); li++) {
    int ab = arities[li];
    fri_commit_layer(ctx, cur_values, cur_log, ab, cap_height, &layers[li]);
    ch.observe_n(layers[li].tree.cap.data(), layers[li].tree.cap.size());
    w.hashes(layers[li].tree.cap.data(), layers[li].tree.cap.size() / 4);
    gl2 beta = ch.get_ext(); betas.push_back(beta);
    size_t n = size_t(1) << cur_log, m = n >> ab;
    DevBuf<u64> folded(ctx, 2 * m), vals(ctx, 2 * m);
    fri_fold_coeffs(ctx, cur_coeffs, n, ab, beta, folded);
    shift = gl_pow(shift, (u64)1 << ab);
    cur_log -= ab;
    ntt_batch(ctx, folded, m, vals, m, 2, cur_log, false, shift, nullptr);
    cur_coeffs = std::move(folded); cur_values = std::move(vals);
  }
  // final polynomial: truncate by the rate, observe
  size_t nfinal = (size_t(1) << cur_log) >> rate_bits;
  std::vector<u64> fa(nfinal), fb(nfinal), final_poly(2 * nfinal);
  CUDA_CHECK(cudaMemcpyAsync(fa.data(), cur_coeffs, nfinal * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(fb.data(), cur_coeffs + (size_t(1) << cur_log), nfinal * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  for (size_t i = 0; i < nfinal; i++) { final_poly[2 * i] = fa[i]; final_poly[2 * i + 1] = fb[i]; }
  ch.observe_n(final_poly.data(), final_poly.size());
  tm.mark("fold codewords in the commitment phase");
  // proof of work (canonical choice: smallest valid witness; upstream accepts any -- SURVEY.md U2)
  u64 pst[12]; memcpy(pst, ch.st, sizeof pst);
  for (size_t i = 0; i < ch.in.size(); i++) pst[i] = ch.in[i];
  u64 pow_witness = fri_pow_search(ctx, pst, (int)ch.in.size(), cfg.pow_bits);
  ch.observe(pow_witness);
  u64 pow_response = ch.get();
  if ((pow_response >> (64 - cfg.pow_bits)) != 0) throw SbnError(SBN_ERR_INTERNAL, "proof-of-work self-check failed");
  tm.mark("find proof-of-work witness");
  // queries
  std::vector<u64> indices(cfg.num_query_rounds);
  for (auto& x : indices) x = ch.get() % L;
  std::vector<QueryOracle> qo; qo.push_back({trace_c.lde, C, &trace_c.tree}); if (Z) qo.push_back({z_c.lde, Z, &z_c.tree}); qo.push_back({q_c.lde, nq_polys, &q_c.tree});
  std::vector<FriLayer*> lp; for (auto& l : layers) lp.push_back(&l);
  const size_t rw_o = fri_query_record_words(qo, {}), rw_l = fri_query_record_words({}, lp), nq = indices.size();
  std::vector<u64> rec_o(rw_o * nq), rec_l(rw_l * nq);
  if (streaming) {
    // rows of a streamed commitment: every sub-coset that holds a queried leaf is evaluated again and its rows are gathered
    DevBuf<u64> d_rec(ctx, rw_o * nq);
    const u64 rmask = (u64(1) << rate_bits) - 1;
    for (int b = 0; b < (1 << rate_bits); b++) {
      bool any = false;
      for (u64 x : indices) any = any || ((u64)bitrev32((u32)x, logL) & rmask) == (u64)b;
      if (!any) continue;
      lde_sub_coset(ctx, trace_c.coeffs, scratch_t, C, logn, rate_bits, b);
      if (Z) lde_sub_coset(ctx, z_c.coeffs, scratch_z, Z, logn, rate_bits, b);
      std::vector<QueryOracle> qs; qs.push_back({scratch_t, C, &trace_c.tree, b}); if (Z) qs.push_back({scratch_z, Z, &z_c.tree, b}); qs.push_back({q_c.lde, nq_polys, &q_c.tree});
      fri_gather_queries(ctx, qs, logn, rate_bits, {}, indices, nullptr, d_rec);
    }
    CUDA_CHECK(cudaMemcpyAsync(rec_o.data(), d_rec, rec_o.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
  } else if (!sh.on()) {
    fri_gather_queries(ctx, qo, logn, rate_bits, {}, indices, rec_o.data());
  } else {
    // the rank that owns a leaf opens it: rows and paths live in its class batches, at the leaf's index inside the class
    const int logLp = logL - sh.m;
    std::vector<u64> local_idx;
    for (size_t q = 0; q < nq; q++) if ((int)(indices[q] >> logLp) == sh.rank) local_idx.push_back(indices[q] & ((u64(1) << logLp) - 1));
    // every rank knows all indices, hence how many queries each rank answers: blocks are padded to the largest count only
    std::vector<size_t> count(sh.world, 0), slot(nq);
    for (size_t q = 0; q < nq; q++) slot[q] = count[indices[q] >> logLp]++;
    size_t max_count = 0; for (size_t c : count) max_count = std::max(max_count, c);
    const int l_logn = sh.m <= rate_bits ? logn : logLp, l_rate = sh.m <= rate_bits ? rate_bits - sh.m : 0;   // geometry of a class batch
    std::vector<uint8_t> all;
    if (sh.allgather_device) {   // rows stay on the device until every rank has every block: one device-to-host copy
      DevBuf<u64> d_mine(ctx, rw_o * max_count), d_all(ctx, rw_o * max_count * sh.world);
      CUDA_CHECK(cudaMemsetAsync(d_mine, 0, rw_o * max_count * 8, ctx->stream));
      if (!local_idx.empty()) fri_gather_queries(ctx, qo, l_logn, l_rate, {}, local_idx, nullptr, d_mine);
      sh.gather_device(ctx, d_mine, rw_o * max_count * 8, d_all);
      all.resize(rw_o * max_count * 8 * sh.world);
      CUDA_CHECK(cudaMemcpyAsync(all.data(), d_all, all.size(), cudaMemcpyDeviceToHost, ctx->stream));
      ctx->sync();
    } else {
      std::vector<u64> mine(rw_o * max_count, 0);
      if (!local_idx.empty()) fri_gather_queries(ctx, qo, l_logn, l_rate, {}, local_idx, mine.data());
      sh.gather(mine.data(), mine.size() * 8, all);
    }
    const u64* parts = reinterpret_cast<const u64*>(all.data());
    for (size_t q = 0; q < nq; q++) memcpy(rec_o.data() + q * rw_o, parts + ((size_t)(indices[q] >> logLp) * max_count + slot[q]) * rw_o, rw_o * 8);
  }
  fri_gather_queries(ctx, {}, logn, rate_bits, lp, indices, rec_l.data());
  w.u32_((uint32_t)indices.size());
  for (size_t q = 0; q < indices.size(); q++) {
    const u64* r = rec_o.data() + q * rw_o;
    w.u32_((uint32_t)qo.size());
    for (auto& o : qo) { w.fvec(r, o.ncols); r += o.ncols; w.hashes(r, o.tree->proof_len()); r += (size_t)o.tree->proof_len() * 4; }
    r = rec_l.data() + q * rw_l;
    w.u32_((uint32_t)lp.size());
    for (auto* l : lp) { size_t ne = size_t(1) << l->arity_bits; w.evec(r, ne); r += 2 * ne; w.hashes(r, l->tree.proof_len()); r += (size_t)l->tree.proof_len() * 4; }
  }
  w.evec(final_poly.data(), nfinal);
  w.f(pow_witness);
  w.fvec(public_inputs, npis);
  tm.mark("query rounds");
  proof->timings = tm.json();
  // challenges for parity tests: alphas | zeta | fri_alpha | permutation sets
  proof->dbg_ch.clear();
  for (int i = 0; i < nch; i++) proof->dbg_ch.push_back(alphas[i]);
  proof->dbg_ch.push_back(zeta.a); proof->dbg_ch.push_back(zeta.b); proof->dbg_ch.push_back(fri_alpha.a); proof->dbg_ch.push_back(fri_alpha.b);
  for (u64 v : perm_challenges) proof->dbg_ch.push_back(v);
}

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_create_error;
#define API_BEGIN try {
#define API_END(ctx)                                                                             \
  } catch (const SbnError& e) { if (ctx) (ctx)->last_error = e.what(); else g_create_error = e.what(); return e.code; } \
  catch (const std::exception& e) { if (ctx) (ctx)->last_error = e.what(); else g_create_error = e.what(); return -4; }   \
  return 0;

extern "C" {
static sbn_ctx* ctx_new(int device, void* cuda_stream, std::shared_ptr<SharedTables> tables) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) throw SbnError(-2, std::string("no CUDA device available: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
  SBN_REQUIRE(device >= 0 && device < ndev, "bad device index");
  CUDA_CHECK(cudaSetDevice(device));
  std::unique_ptr<sbn_ctx> c(new sbn_ctx());   // nothing leaks if a later step throws
  c->device = device;
  if (tables) c->tables = tables;
  cudaDeviceProp prop; CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  c->num_sms = prop.multiProcessorCount;
  if (cuda_stream) c->stream = (cudaStream_t)cuda_stream; else { CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->owns_stream = true; }
  return c.release();
}
static void ctx_delete(sbn_ctx* ctx) {
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->release_all();
  ctx->kresolve();
  for (auto e : ctx->kpool) cudaEventDestroy(e);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;   // the table set goes with its last owner
}
int sbn_ctx_create(int device, void* cuda_stream, sbn_ctx** out) {
  sbn_ctx* ctx = nullptr;
  API_BEGIN
  SBN_REQUIRE(out, "null output pointer");
  *out = ctx_new(device, cuda_stream, nullptr);
  API_END(ctx)
}
void sbn_ctx_destroy(sbn_ctx* ctx) {
  if (!ctx) return;
  // traces hold buffers of this context's allocator: with live traces the destruction is deferred to the last sbn_trace_free
  if (ctx->live_handles > 0) { ctx->destroy_pending = true; return; }
  ctx_delete(ctx);
}
const char* sbn_last_error(const sbn_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }
int sbn_ctx_select_field(sbn_ctx* ctx, uint64_t coset_shift) { API_BEGIN SBN_REQUIRE(ctx, "null context"); CUDA_CHECK(cudaSetDevice(ctx->device)); ctx->tables->select_generator(coset_shift); API_END(ctx) }
int sbn_ctx_synchronize(sbn_ctx* ctx) { API_BEGIN SBN_REQUIRE(ctx, "null context"); CUDA_CHECK(cudaSetDevice(ctx->device)); ctx->sync(); API_END(ctx) }
uint64_t sbn_ctx_launch_count(const sbn_ctx* ctx) { return ctx->launches; }
uint64_t sbn_ctx_device_bytes(const sbn_ctx* ctx) { return ctx->bytes_allocated; }

int sbn_ctx_kernel_timing(sbn_ctx* ctx, int enable) {
  if (!ctx) return -1;
  cudaSetDevice(ctx->device);
  ctx->kresolve();
  ctx->ktime_enabled = enable != 0;
  ctx->kstats.clear();
  return 0;
}
int sbn_ctx_kernel_stats(sbn_ctx* ctx, char* buf, size_t cap) {
  if (!ctx || !buf || !cap) return -1;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->kresolve();
  std::ostringstream os; os << "{"; bool first = true;
  for (auto& kv : ctx->kstats) { os << (first ? "" : ",") << "\"" << kv.first << "\":{\"ms\":" << kv.second.ms << ",\"count\":" << kv.second.count << "}"; first = false; }
  os << "}";
  snprintf(buf, cap, "%s", os.str().c_str());
  return 0;
}
int sbn_config_standard_fast(sbn_config* out) {
  if (!out) return -1;
  *out = sbn_config{100, 2, 1, 4, 16, 4, 5, 84, GL_MULT_GENERATOR, 0, 0};
  return 0;
}
int sbn_air_info(int air, size_t num_io, size_t* num_columns, size_t* num_public_inputs, size_t* num_rows, size_t* io_size, size_t* result_words,
                 size_t* num_permutation_pairs) {
  sbn_ctx* ctx = nullptr;
  API_BEGIN
  AirDesc a = make_air(air, num_io);
  if (num_columns) *num_columns = a.num_columns;
  if (num_public_inputs) *num_public_inputs = a.num_public_inputs;
  if (num_rows) *num_rows = a.num_rows;
  if (io_size) *io_size = a.io_size;
  if (result_words) *result_words = a.result_words;
  if (num_permutation_pairs) *num_permutation_pairs = a.perm_pairs.size();
  API_END(ctx)
}

static int trace_generate_impl(sbn_ctx* ctx, int air, const void* ios, bool on_device, size_t num_io, sbn_trace** out) {
  API_BEGIN
  SBN_REQUIRE(ctx && ios && out, "null argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ctx->begin_call();
  std::unique_ptr<sbn_trace> t(new sbn_trace());
  t->ctx = ctx; t->air = make_air(air, num_io); t->logn = ilog2(t->air.num_rows);
  t->cols = DevBuf<u64>(ctx, t->air.num_columns * t->air.num_rows);
  t->results.resize(t->air.result_words * num_io);
  generate_trace(ctx, t->air, ios, on_device, t->cols, t->results.data());
  ctx->live_handles++;
  *out = t.release();
  API_END(ctx)
}
int sbn_trace_generate(sbn_ctx* ctx, int air, const void* ios, size_t num_io, sbn_trace** out) { return trace_generate_impl(ctx, air, ios, false, num_io, out); }
int sbn_trace_generate_device(sbn_ctx* ctx, int air, const void* d_ios, size_t num_io, sbn_trace** out) { return trace_generate_impl(ctx, air, d_ios, true, num_io, out); }
int sbn_trace_upload(sbn_ctx* ctx, int air, size_t num_io, const uint64_t* cols, size_t ncols, size_t nrows, sbn_trace** out) {
  API_BEGIN
  SBN_REQUIRE(ctx && cols && out, "null argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  std::unique_ptr<sbn_trace> t(new sbn_trace());
  t->ctx = ctx; t->air = make_air(air, num_io);
  SBN_REQUIRE(ncols == t->air.num_columns && nrows == t->air.num_rows, "trace shape does not match the AIR");
  t->logn = ilog2(nrows);
  t->cols = DevBuf<u64>(ctx, ncols * nrows);
  CUDA_CHECK(cudaMemcpyAsync(t->cols, cols, ncols * nrows * 8, cudaMemcpyHostToDevice, ctx->stream));
  ctx->sync();
  ctx->live_handles++;
  *out = t.release();
  API_END(ctx)
}
int sbn_trace_download(const sbn_trace* t, uint64_t* cols_out) {
  sbn_ctx* ctx = t ? t->ctx : nullptr;
  API_BEGIN
  SBN_REQUIRE(t && cols_out, "null argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  CUDA_CHECK(cudaMemcpyAsync(cols_out, t->cols, t->air.num_columns * t->air.num_rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  API_END(ctx)
}
int sbn_trace_results(const sbn_trace* t, uint64_t* out) {
  if (!t || !out) return -1;
  memcpy(out, t->results.data(), t->results.size() * 8);
  return 0;
}
void sbn_trace_free(sbn_trace* t) {
  if (!t) return;
  sbn_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  delete t;   // returns the buffer to the context's allocator
  if (--ctx->live_handles == 0 && ctx->destroy_pending) ctx_delete(ctx);
}

int sbn_public_inputs(int air, const void* ios, size_t num_io, uint64_t* out, size_t out_len) {
  sbn_ctx* ctx = nullptr;
  API_BEGIN
  AirDesc a = make_air(air, num_io);
  SBN_REQUIRE(out_len == a.num_public_inputs, "public input buffer has the wrong length");
  format_public_inputs(a, ios, (u64*)out);
  API_END(ctx)
}

int sbn_prove(sbn_ctx* ctx, const sbn_config* config, const sbn_trace* trace, const uint64_t* public_inputs, size_t num_public_inputs, sbn_proof** out) {
  API_BEGIN
  SBN_REQUIRE(ctx && config && trace && out && (public_inputs || num_public_inputs == 0), "null argument");
  SBN_REQUIRE(trace->ctx == ctx, "trace belongs to another context");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ctx->begin_call();
  ctx->tables->select_generator(config->coset_shift);
  std::unique_ptr<sbn_proof> p(new sbn_proof());
  prove_impl(ctx, *config, trace, (const u64*)public_inputs, num_public_inputs, Shard(), p.get());
  *out = p.release();
  API_END(ctx)
}
int sbn_prove_sharded(sbn_ctx* ctx, const sbn_config* config, const sbn_trace* trace, const uint64_t* public_inputs, size_t num_public_inputs,
                      const sbn_shard* shard, sbn_proof** out) {
  API_BEGIN
  SBN_REQUIRE(ctx && config && trace && out && shard && (public_inputs || num_public_inputs == 0), "null argument");
  SBN_REQUIRE(trace->ctx == ctx, "trace belongs to another context");
  SBN_REQUIRE(shard->world >= 1 && shard->world <= 16 && (shard->world & (shard->world - 1)) == 0 && shard->rank < shard->world, "bad shard description");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  Shard sh; sh.rank = (int)shard->rank; sh.world = (int)shard->world; sh.allgather = shard->allgather; sh.user = shard->user; sh.allgather_device = shard->allgather_device;
  while ((1 << sh.m) < sh.world) sh.m++;
  ctx->begin_call();
  ctx->tables->select_generator(config->coset_shift);
  std::unique_ptr<sbn_proof> p(new sbn_proof());
  prove_impl(ctx, *config, trace, (const u64*)public_inputs, num_public_inputs, sh, p.get());
  *out = p.release();
  API_END(ctx)
}
// ---- batched proofs: a pool of lanes (context + host thread each) on one device sharing one table set ----
struct sbn_batch {
  int device = 0;
  std::vector<sbn_ctx*> lanes;
  std::string last_error;
};
int sbn_batch_create(int device, uint32_t lanes, sbn_batch** out) {
  sbn_ctx* ctx = nullptr;
  API_BEGIN
  SBN_REQUIRE(out && lanes >= 1 && lanes <= 64, "sbn_batch_create: 1 <= lanes <= 64");
  std::unique_ptr<sbn_batch> b(new sbn_batch());
  b->device = device;
  auto tables = std::make_shared<SharedTables>();
  try {
    for (uint32_t i = 0; i < lanes; i++) b->lanes.push_back(ctx_new(device, nullptr, tables));
  } catch (...) { for (auto* c : b->lanes) ctx_delete(c); throw; }
  *out = b.release();
  API_END(ctx)
}
void sbn_batch_destroy(sbn_batch* b) {
  if (!b) return;
  for (auto* c : b->lanes) ctx_delete(c);
  delete b;
}
const char* sbn_batch_last_error(const sbn_batch* b) { return b ? b->last_error.c_str() : g_create_error.c_str(); }
int sbn_ctx_trim(sbn_ctx* ctx) { API_BEGIN SBN_REQUIRE(ctx, "null context"); CUDA_CHECK(cudaSetDevice(ctx->device)); ctx->sync(); ctx->trim(); API_END(ctx) }
int sbn_batch_trim(sbn_batch* b) {
  if (!b) return SBN_ERR_INVALID;
  cudaSetDevice(b->device);
  for (auto* c : b->lanes) { cudaStreamSynchronize(c->stream); c->trim(); }
  return 0;
}
uint64_t sbn_batch_launch_count(const sbn_batch* b) { uint64_t n = 0; if (b) for (auto* c : b->lanes) n += c->launches; return n; }
uint64_t sbn_batch_device_bytes(const sbn_batch* b) { uint64_t n = 0; if (b) for (auto* c : b->lanes) n += c->bytes_allocated; return n; }

// one proof on one lane: K1, public inputs, K2-K6
static void batch_job(sbn_ctx* ctx, const AirDesc& air, const sbn_config& cfg, const void* ios, uint32_t flags, sbn_proof* proof) {
  const bool on_device = (flags & SBN_BATCH_IOS_ON_DEVICE) != 0;
  ctx->begin_call();
  sbn_trace t;
  t.ctx = ctx; t.air = air; t.logn = ilog2(air.num_rows);
  t.cols = DevBuf<u64>(ctx, air.num_columns * air.num_rows);
  t.results.resize(air.result_words * air.num_io);
  generate_trace(ctx, air, ios, on_device, t.cols, t.results.data());
  std::vector<u64> pis(air.num_public_inputs);
  if (air.num_public_inputs) {
    const size_t bytes = air.io_size * air.num_io;
    std::vector<uint8_t> host;
    const void* recs = ios;
    if (on_device || (flags & SBN_BATCH_FILL_OUTPUTS)) {
      host.resize(bytes);
      if (on_device) { CUDA_CHECK(cudaMemcpyAsync(host.data(), ios, bytes, cudaMemcpyDeviceToHost, ctx->stream)); ctx->sync(); }
      else memcpy(host.data(), ios, bytes);
      if (flags & SBN_BATCH_FILL_OUTPUTS) {   // output = the chain result: the trailing result_words u64 of every record
        const size_t rb = air.result_words * 8;
        for (size_t i = 0; i < air.num_io; i++) memcpy(host.data() + (i + 1) * air.io_size - rb, t.results.data() + i * air.result_words, rb);
      }
      recs = host.data();
    }
    format_public_inputs(air, recs, pis.data());
  }
  prove_impl(ctx, cfg, &t, pis.data(), pis.size(), Shard(), proof);
}
int sbn_prove_batch(sbn_batch* b, int air_id, size_t num_io, const sbn_config* config, const void* const* ios, size_t count, uint32_t flags,
                    sbn_proof** proofs_out) {
  if (!b) return SBN_ERR_INVALID;
  try {
    SBN_REQUIRE(config && ios && proofs_out, "null argument");
    const AirDesc air = make_air(air_id, num_io);
    for (size_t j = 0; j < count; j++) { SBN_REQUIRE(ios[j], "null input batch"); proofs_out[j] = nullptr; }
    CUDA_CHECK(cudaSetDevice(b->device));
    b->lanes[0]->tables->select_generator(config->coset_shift);   // one table set for all lanes; no lane is running yet
    std::vector<std::unique_ptr<sbn_proof>> proofs(count);
    std::atomic<size_t> next(0);
    std::atomic<bool> failed(false);
    std::mutex err_mu; std::string err; int err_code = 0;
    auto work = [&](sbn_ctx* ctx) {
      try {
        CUDA_CHECK(cudaSetDevice(ctx->device));
        for (;;) {
          const size_t j = next.fetch_add(1);
          if (j >= count || failed.load()) return;
          std::unique_ptr<sbn_proof> p(new sbn_proof());
          batch_job(ctx, air, *config, ios[j], flags, p.get());
          proofs[j] = std::move(p);
        }
      } catch (const SbnError& e) { failed = true; std::lock_guard<std::mutex> g(err_mu); if (!err_code) { err_code = e.code; err = e.what(); } }
      catch (const std::exception& e) { failed = true; std::lock_guard<std::mutex> g(err_mu); if (!err_code) { err_code = SBN_ERR_INTERNAL; err = e.what(); } }
    };
    // lanes that may run at once: each proof in flight holds trace + coefficients + LDE of the trace and Z batches (+ Z values);
    // large shapes (G1 with 512 instances: 2^18 rows; Fq12 with 128: 10 250 columns x 2^16 rows) do not fit six times
    size_t nworkers = std::min(b->lanes.size(), count);
    {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) {
        const double nz = air.perm_pairs.empty() ? 0.0 : (double)((air.perm_pairs.size() * config->num_challenges + 1) / 2);
        const double per_proof = 8.0 * (double)air.num_rows * (((double)air.num_columns + nz) * (2.0 + (double)(1u << config->rate_bits)) + (double)air.num_columns) * 1.15;
        const size_t fit = (size_t)std::max(1.0, std::floor(0.8 * (double)total_b / per_proof));
        nworkers = std::min(nworkers, fit);
      }
    }
    std::vector<std::thread> threads;
    for (size_t w = 1; w < nworkers; w++) threads.emplace_back(work, b->lanes[w]);
    if (nworkers) work(b->lanes[0]);   // the calling thread drives lane 0
    for (auto& t : threads) t.join();
    if (err_code) { b->last_error = err; return err_code; }
    for (size_t j = 0; j < count; j++) proofs_out[j] = proofs[j].release();
  } catch (const SbnError& e) { b->last_error = e.what(); return e.code; }
  catch (const std::exception& e) { b->last_error = e.what(); return SBN_ERR_INTERNAL; }
  return 0;
}

int sbn_proof_serialize(const sbn_proof* proof, uint8_t* buf, size_t* len) {
  if (!proof || !len) return -1;
  if (!buf) { *len = proof->bytes.size(); return 0; }
  if (*len < proof->bytes.size()) { *len = proof->bytes.size(); return -1; }
  memcpy(buf, proof->bytes.data(), proof->bytes.size());
  *len = proof->bytes.size();
  return 0;
}
int sbn_proof_timings(const sbn_proof* proof, char* buf, size_t cap) {
  if (!proof || !buf || cap == 0) return -1;
  snprintf(buf, cap, "%s", proof->timings.c_str());
  return 0;
}
int sbn_proof_debug(const sbn_proof* proof, int which, uint64_t* out, size_t cap_words, size_t* written) {
  if (!proof || !written) return -1;
  const std::vector<u64>* v = which == 0 ? &proof->dbg_z : which == 1 ? &proof->dbg_q : which == 2 ? &proof->dbg_ch : nullptr;
  if (!v) return -1;
  *written = v->size();
  if (out) { if (cap_words < v->size()) return -1; memcpy(out, v->data(), v->size() * 8); }
  return 0;
}
void sbn_proof_free(sbn_proof* p) { delete p; }

__global__ void k_poseidon_batch(u64* states, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 st[12];
#pragma unroll
  for (int k = 0; k < 12; k++) st[k] = states[i * 12 + k];
  poseidon_permute(st);
#pragma unroll
  for (int k = 0; k < 12; k++) states[i * 12 + k] = st[k];
}
int sbn_poseidon_permute(sbn_ctx* ctx, uint64_t* states, size_t n) {
  API_BEGIN
  SBN_REQUIRE(ctx && states, "null argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ctx->begin_call();
  DevBuf<u64> d(ctx, n * 12);
  CUDA_CHECK(cudaMemcpyAsync(d, states, n * 96, cudaMemcpyHostToDevice, ctx->stream));
  k_poseidon_batch<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d, n);
  LAUNCH_CHECK(ctx);
  CUDA_CHECK(cudaMemcpyAsync(states, d, n * 96, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  API_END(ctx)
}

__global__ void k_lde_to_natural(const u64* lde, u64* out, int logn, int rate_bits, size_t ncols) {
  size_t L = size_t(1) << (logn + rate_bits), N = size_t(1) << logn;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= L * ncols) return;
  size_t c = t / L, i = t % L;
  out[t] = lde[c * L + (i & ((size_t(1) << rate_bits) - 1)) * N + (i >> rate_bits)];
}
int sbn_commit_columns(sbn_ctx* ctx, const uint64_t* values, size_t ncols, int logn, int rate_bits, int cap_height, uint64_t* coeffs_out,
                       uint64_t* lde_out, uint64_t* cap_out) {
  API_BEGIN
  SBN_REQUIRE(ctx && values && ncols > 0, "null argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ctx->begin_call();
  size_t N = size_t(1) << logn, L = N << rate_bits;
  DevBuf<u64> d_vals(ctx, ncols * N);
  CUDA_CHECK(cudaMemcpyAsync(d_vals, values, ncols * N * 8, cudaMemcpyHostToDevice, ctx->stream));
  Commitment c;
  commit_from_values(ctx, Shard(), c, d_vals, (int)ncols, logn, rate_bits, cap_height, false);
  if (coeffs_out) CUDA_CHECK(cudaMemcpyAsync(coeffs_out, c.coeffs, ncols * N * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (lde_out) {
    DevBuf<u64> nat(ctx, ncols * L);
    k_lde_to_natural<<<(unsigned)((ncols * L + 255) / 256), 256, 0, ctx->stream>>>(c.lde, nat, logn, rate_bits, ncols);
    LAUNCH_CHECK(ctx);
    CUDA_CHECK(cudaMemcpyAsync(lde_out, nat, ncols * L * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
  }
  if (cap_out) memcpy(cap_out, c.tree.cap.data(), c.tree.cap.size() * 8);
  ctx->sync();
  API_END(ctx)
}

__global__ void k_fill_pseudo(u64* p, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 z = (i + 1) * 0x9E3779B97F4A7C15ULL; z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 32;
  p[i] = z >= GL_P ? z - GL_P : z;
}
int sbn_bench_commit(sbn_ctx* ctx, size_t ncols, int logn, int rate_bits, int cap_height, int iters, float* ms) {
  API_BEGIN
  SBN_REQUIRE(ctx && ms && iters > 0, "bad argument");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ctx->begin_call();
  size_t N = size_t(1) << logn, L = N << rate_bits;
  DevBuf<u64> vals(ctx, ncols * N), coeffs(ctx, ncols * N), lde(ctx, ncols * L);
  k_fill_pseudo<<<(unsigned)((ncols * N + 255) / 256), 256, 0, ctx->stream>>>(vals, ncols * N);
  LAUNCH_CHECK(ctx);
  cudaEvent_t e[4]; for (auto& x : e) CUDA_CHECK(cudaEventCreate(&x));
  ms[0] = ms[1] = ms[2] = 0;
  for (int it = -1; it < iters; it++) {  // one warm-up
    DevMerkleTree tree;
    merkle_alloc(ctx, &tree, L, cap_height);
    CUDA_CHECK(cudaEventRecord(e[0], ctx->stream));
    intt_columns(ctx, vals, coeffs, (int)ncols, logn);
    lde_columns(ctx, coeffs, lde, (int)ncols, logn, rate_bits);
    CUDA_CHECK(cudaEventRecord(e[1], ctx->stream));
    merkle_leaf_hash_only(ctx, lde, (int)ncols, logn, rate_bits, &tree);
    CUDA_CHECK(cudaEventRecord(e[2], ctx->stream));
    merkle_build_from_leaf_digests(ctx, &tree);
    CUDA_CHECK(cudaEventRecord(e[3], ctx->stream));
    CUDA_CHECK(cudaEventSynchronize(e[3]));
    if (it >= 0) for (int k = 0; k < 3; k++) { float t; CUDA_CHECK(cudaEventElapsedTime(&t, e[k], e[k + 1])); ms[k] += t / iters; }
  }
  for (auto& x : e) cudaEventDestroy(x);
  API_END(ctx)
}
}  // extern "C"
