// ORACLE (test infrastructure).  Poseidon-Goldilocks permutation, hash_or_noop / two_to_one,
// MerkleTree with caps, and the duplex Challenger, restated from the published plonky2 0.1.3
// algorithm (dependency absent from /root/reference: Cargo.lock:529-531).
// Call sites that pin the behaviour: every `prove::<F, C, _, D>` with C = PoseidonGoldilocksConfig
// (e.g. reference src/curves/g1/exp.rs:788-825).
#pragma once
#include "gl.hpp"
#include <array>
#include <cstring>

namespace orc {

static const u64 POSEIDON_RC[360] = {
#include "poseidon_rc.inc"
    SBN_POSEIDON_RC_LIST};
static const u64 MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static const u64 MDS_DIAG[12] = {8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

static inline GF sbox7(GF x) { GF x2 = x * x, x3 = x2 * x, x4 = x2 * x2; return x3 * x4; }

// Plain round structure: 4 full, 22 partial, 4 full; constants added to all 12 lanes in every round;
// x^7 on all lanes (full) or lane 0 (partial); circulant+diagonal MDS.  The MDS product is evaluated
// on the 32-bit halves of each lane (12 * 2^32 * 41 < 2^42 per half) and recombined -- same values.
static inline void mds_layer(GF s[12]) {
  alignas(32) u32 lo[24], hi[24];
  alignas(32) u64 al[12], ah[12];
  for (int i = 0; i < 12; i++) { lo[i] = lo[i + 12] = (u32)s[i].v; hi[i] = hi[i + 12] = (u32)(s[i].v >> 32); al[i] = ah[i] = 0; }
  for (int i = 0; i < 12; i++) {
    const u32 c = (u32)MDS_CIRC[i];
    for (int k = 0; k < 12; k++) { al[k] += (u64)lo[i + k] * c; ah[k] += (u64)hi[i + k] * c; }
  }
  al[0] += (u64)lo[0] * MDS_DIAG[0]; ah[0] += (u64)hi[0] * MDS_DIAG[0];
  for (int k = 0; k < 12; k++) s[k].v = gl_reduce128((u128)al[k] + ((u128)ah[k] << 32));
}
static inline void poseidon(GF s[12]) {
  for (int r = 0; r < 30; r++) {
    const u64* rc = POSEIDON_RC + 12 * r;
    for (int i = 0; i < 12; i++) { GF c; c.v = rc[i]; s[i] = s[i] + c; }
    if (r < 4 || r >= 26) { for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]); }
    else s[0] = sbox7(s[0]);
    mds_layer(s);
  }
}

// O(t)-per-round form of the 22 partial rounds (constants derived and checked against the plain rounds by
// tools/gen_poseidon_fast.py; tests/test_oracle_core.py checks poseidon_fast == poseidon).  plonky2's own CPU
// permutation uses the same kind of optimisation, so the CPU baseline timed by bench.py is not handicapped.
static const u64 POSEIDON_FAST[639] = {
#include "poseidon_fast.inc"
    SBN_POSEIDON_FAST_LIST};
static inline void poseidon_fast(GF s[12]) {
  auto full = [&](int r) {
    const u64* rc = POSEIDON_RC + 12 * r;
    for (int i = 0; i < 12; i++) { GF c; c.v = rc[i]; s[i] = sbox7(s[i] + c); }
    mds_layer(s);
  };
  for (int r = 0; r < 4; r++) full(r);
  const u64* f = POSEIDON_FAST;
  for (int r = 0; r < 22; r++, f += 23) {
    GF g0; g0.v = f[0];
    GF x0 = sbox7(s[0] + g0);
    u128 acc = (u128)x0.v * 25;
    for (int i = 0; i < 11; i++) { GF v; v.v = f[1 + i]; acc += (v * s[1 + i]).v; }
    for (int i = 0; i < 11; i++) { GF u; u.v = f[12 + i]; s[1 + i] = s[1 + i] + u * x0; }
    s[0].v = gl_reduce128(acc);
  }
  GF o[11];
  for (int i = 0; i < 11; i++) {
    u128 acc = f[121 + 1 + i];
    for (int j = 0; j < 11; j++) { GF d; d.v = f[11 * i + j]; acc += (d * s[1 + j]).v; }
    o[i].v = gl_reduce128(acc);
  }
  { GF e0; e0.v = f[121]; s[0] = s[0] + e0; }
  for (int i = 0; i < 11; i++) s[1 + i] = o[i];
  for (int r = 26; r < 30; r++) full(r);
}

struct Hash4 { GF e[4]; bool operator==(const Hash4& o) const { return e[0]==o.e[0]&&e[1]==o.e[1]&&e[2]==o.e[2]&&e[3]==o.e[3]; } };

// PoseidonHash::hash_no_pad: overwrite-mode sponge, rate 8, no padding, squeeze 4.
static inline Hash4 hash_no_pad(const GF* in, size_t n) {
  GF st[12];
  for (size_t off = 0; off < n; off += 8) {
    size_t m = n - off < 8 ? n - off : 8;
    for (size_t i = 0; i < m; i++) st[i] = in[off + i];
    poseidon_fast(st);
  }
  Hash4 h; for (int i = 0; i < 4; i++) h.e[i] = st[i]; return h;
}
// Hasher::hash_or_noop: <= 4 elements are copied (zero padded) instead of hashed.
static inline Hash4 hash_or_noop(const GF* in, size_t n) {
  if (n <= 4) { Hash4 h; for (size_t i = 0; i < n; i++) h.e[i] = in[i]; return h; }
  return hash_no_pad(in, n);
}
static inline Hash4 two_to_one(const Hash4& l, const Hash4& r) {
  GF st[12];
  for (int i = 0; i < 4; i++) { st[i] = l.e[i]; st[4 + i] = r.e[i]; }
  poseidon_fast(st);
  Hash4 h; for (int i = 0; i < 4; i++) h.e[i] = st[i]; return h;
}

// MerkleTree::new(leaves, cap_height).  layers[0] = leaf digests, layers[k] = parents; the cap is
// the layer holding 2^cap_height nodes.  prove(i) = siblings from the leaf level up to (excluding)
// the cap level.
struct MerkleTree {
  std::vector<std::vector<GF>> leaves;
  std::vector<std::vector<Hash4>> layers;
  std::vector<Hash4> cap;
  int cap_height = 0;
  MerkleTree() {}
  MerkleTree(std::vector<std::vector<GF>> lv, int ch) : leaves(std::move(lv)), cap_height(ch) {
    size_t n = leaves.size();
    int lg = log2_strict(n);
    assert(lg >= ch);
    std::vector<Hash4> cur(n);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) cur[i] = hash_or_noop(leaves[i].data(), leaves[i].size());
    layers.push_back(cur);
    while (cur.size() > (size_t(1) << ch)) {
      std::vector<Hash4> nxt(cur.size() / 2);
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < nxt.size(); i++) nxt[i] = two_to_one(cur[2 * i], cur[2 * i + 1]);
      layers.push_back(nxt);
      cur.swap(nxt);
    }
    cap = cur;
  }
  std::vector<Hash4> prove(size_t idx) const {
    std::vector<Hash4> sib;
    for (size_t l = 0; l + 1 < layers.size(); l++) { sib.push_back(layers[l][idx ^ 1]); idx >>= 1; }
    return sib;
  }
};

static inline bool verify_merkle_proof_to_cap(const GF* leaf, size_t n, size_t idx, const std::vector<Hash4>& cap,
                                              const std::vector<Hash4>& sib) {
  Hash4 cur = hash_or_noop(leaf, n);
  for (const Hash4& s : sib) {
    cur = (idx & 1) ? two_to_one(s, cur) : two_to_one(cur, s);
    idx >>= 1;
  }
  return idx < cap.size() && cur == cap[idx];
}

// plonky2::iop::challenger::Challenger (duplex sponge, overwrite mode; get_challenge pops from the
// END of the 8-element output buffer).
struct Challenger {
  GF st[12];
  std::vector<GF> in, out;
  void duplex() {
    for (size_t i = 0; i < in.size(); i++) st[i] = in[i];
    in.clear();
    poseidon_fast(st);
    out.assign(st, st + 8);
  }
  void observe(GF x) { out.clear(); in.push_back(x); if (in.size() == 8) duplex(); }
  void observe(const Hash4& h) { for (int i = 0; i < 4; i++) observe(h.e[i]); }
  void observe_cap(const std::vector<Hash4>& c) { for (auto& h : c) observe(h); }
  void observe(GF2 x) { observe(x.a); observe(x.b); }
  GF get() { if (!in.empty() || out.empty()) duplex(); GF r = out.back(); out.pop_back(); return r; }
  GF2 get_ext() { GF a = get(); GF b = get(); return GF2(a, b); }
};
}  // namespace orc
