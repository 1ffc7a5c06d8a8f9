// Poseidon-Goldilocks (width 12, rate 8, 4+22+4 rounds, x^7) -- the hash of plonky2's
// PoseidonGoldilocksConfig, which the reference selects at every prove() call site
// (e.g. reference src/curves/g1/exp.rs:788-790).  One permutation per thread, state in registers;
// the MDS layer works on the 32-bit halves of each lane so every product is a 32x(6-bit) IMAD.
#pragma once
#include "gl.cuh"

static const u64 h_poseidon_rc[360] = {
#include "poseidon_rc.inc"
    SBN_POSEIDON_RC_LIST};
#ifdef __CUDACC__
static __constant__ u64 d_poseidon_rc[360] = {SBN_POSEIDON_RC_LIST};
#endif

#ifdef __CUDA_ARCH__
#define POSEIDON_RC(i) d_poseidon_rc[i]
#else
#define POSEIDON_RC(i) h_poseidon_rc[i]
#endif

HD u64 poseidon_sbox(u64 x) {
  u64 x2 = gl_sqr(x), x3 = gl_mul(x2, x), x4 = gl_sqr(x2);
  return gl_mul(x3, x4);
}

// out[k] = sum_i s[(i+k) % 12] * circ[i] + (k == 0 ? 8 * s[0] : 0), circ = [17,15,41,16,2,28,13,13,39,18,34,20]
HD void poseidon_mds(u64 s[12]) {
  const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  u32 lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { lo[i] = (u32)s[i]; hi[i] = (u32)(s[i] >> 32); }
#pragma unroll
  for (int k = 0; k < 12; k++) {
    u64 al = 0, ah = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      al += (u64)lo[(i + k) % 12] * C[i];
      ah += (u64)hi[(i + k) % 12] * C[i];
    }
    if (k == 0) { al += (u64)lo[0] * 8u; ah += (u64)hi[0] * 8u; }
    // value = al + ah * 2^32 with al, ah < 2^42
    u64 l = al + (ah << 32);
    u64 h = (ah >> 32) + (l < al ? 1 : 0);
    s[k] = gl_reduce128(l, h);
  }
}

HD void poseidon_permute(u64 s[12]) {
  int r = 0;
  for (; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add(s[i], POSEIDON_RC(12 * r + i)));
    poseidon_mds(s);
  }
  for (; r < 26; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], POSEIDON_RC(12 * r + i));
    s[0] = poseidon_sbox(s[0]);
    poseidon_mds(s);
  }
  for (; r < 30; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add(s[i], POSEIDON_RC(12 * r + i)));
    poseidon_mds(s);
  }
}
