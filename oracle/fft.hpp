// ORACLE (test infrastructure).  Radix-2 FFT over Goldilocks in plonky2_field's conventions:
// fft(coeffs)[i] = sum_j c_j w^{ij} with w = primitive_root_of_unity(log n); natural order in and
// out; coset_fft(shift) scales c_j by shift^j first; coset_ifft unscales after.
#pragma once
#include "gl.hpp"

namespace orc {
static inline void fft_inplace(std::vector<GF>& a, bool inverse) {
  size_t n = a.size();
  if (n <= 1) return;
  int lg = log2_strict(n);
  reverse_index_bits_in_place(a);
  GF root = root_of_unity(lg);
  if (inverse) root = gl_inv(root);
  for (int s = 1; s <= lg; s++) {
    size_t m = size_t(1) << s, h = m >> 1;
    GF wm = gl_exp_pow2(root, lg - s);
    std::vector<GF> tw(h);
    GF w = GF::one();
    for (size_t j = 0; j < h; j++) { tw[j] = w; w = w * wm; }
    for (size_t k = 0; k < n; k += m)
      for (size_t j = 0; j < h; j++) {
        GF t = tw[j] * a[k + j + h], u = a[k + j];
        a[k + j] = u + t; a[k + j + h] = u - t;
      }
  }
  if (inverse) { GF ninv = gl_inv(GF((u64)n)); for (auto& x : a) x = x * ninv; }
}
static inline std::vector<GF> ifft(std::vector<GF> v) { fft_inplace(v, true); return v; }
static inline std::vector<GF> fft(std::vector<GF> c) { fft_inplace(c, false); return c; }
static inline std::vector<GF> coset_fft(std::vector<GF> c, GF shift) {
  GF s = GF::one();
  for (auto& x : c) { x = x * s; s = s * shift; }
  fft_inplace(c, false); return c;
}
static inline std::vector<GF> coset_ifft(std::vector<GF> v, GF shift) {
  fft_inplace(v, true);
  GF si = gl_inv(shift), s = GF::one();
  for (auto& x : v) { x = x * s; s = s * si; }
  return v;
}
// PolynomialCoeffs::lde(rate_bits).coset_fft(F::coset_shift())
static inline std::vector<GF> lde_onto_coset(const std::vector<GF>& coeffs, int rate_bits) {
  std::vector<GF> c(coeffs); c.resize(coeffs.size() << rate_bits);
  return coset_fft(std::move(c), coset_shift());
}
// Extension-field versions (the FFT is F-linear, roots and shifts live in the base field).
static inline std::vector<GF2> coset_fft_ext(const std::vector<GF2>& c, GF shift) {
  std::vector<GF> a(c.size()), b(c.size());
  for (size_t i = 0; i < c.size(); i++) { a[i] = c[i].a; b[i] = c[i].b; }
  a = coset_fft(std::move(a), shift); b = coset_fft(std::move(b), shift);
  std::vector<GF2> r(c.size());
  for (size_t i = 0; i < c.size(); i++) r[i] = GF2(a[i], b[i]);
  return r;
}
}  // namespace orc
