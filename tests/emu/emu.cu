// TEST HARNESS (not product code): runs the product's __host__ __device__ per-point / per-row
// functions (constraints.cuh, witness.cuh) on the CPU so their logic can be compared with the oracle in
// the GPU-less container.  The product library never links or loads this file.
#include "../../starky-bn254_b200/csrc/constraints.cuh"
#include "../../starky-bn254_b200/csrc/witness.cuh"
#include "../../starky-bn254_b200/csrc/air.cuh"
#include "../../starky-bn254_b200/csrc/poseidon.cuh"
#include <vector>

struct RowWriter { u64* row; void operator()(int col, u64 v) const { row[col] = v; } };

extern "C" {
// Evaluate all AIR segments (in order, combined exactly like quotient.cu) at one point given as plain rows.
int emu_eval_air(int air_id, size_t num_io, const u64* lv, const u64* nv, const u64* pi, const u64* alphas, u64 z_last, u64 l_first, u64 l_last,
                 u64* out_acc, size_t* num_constraints) {
  try {
    AirDesc air = make_air(air_id, num_io);
    u64 acc[2] = {0, 0};
    size_t total = 0;
    for (const Segment& s : air.segments) {
      QPoint q;
      q.lp = lv; q.np = nv; q.stride = 1; q.pi = pi;
      q.z_last = F(z_last); q.l_first = F(l_first); q.l_last = F(l_last);
      for (int c = 0; c < 2; c++) { q.alpha[c] = F(alphas[c]); q.acc[c] = F(); }
      // public-input binding columns at this point: what the prover's sparse-column LDEs evaluate to, i.e. the same linear
      // combinations of the row's pulse columns (quotient.cu build_pi_binding / k_pi_columns)
      std::vector<u64> pic(1, 0);
      q.pic = pic.data(); q.pic_stride = 1; q.pic_per_chal = 0;
      if (s.kind == SEG_FQ_CORE || s.kind == SEG_G1_CORE || s.kind == SEG_G2_CORE || s.kind == SEG_FQ12_CORE) {
        const int gc = s.kind == SEG_FQ_CORE ? 1 : s.kind == SEG_G1_CORE ? 2 : s.kind == SEG_G2_CORE ? 4 : (s.p2 ? -1 : 0);
        const int sf = s.kind == SEG_FQ12_CORE ? s.p1 : s.p2;
        const int start_pulses = sf + (gc == -1 ? 6 : 16), n = s.p0, io_len = pi_io_len(gc), per = io_len + 2;
        pic.assign(1 + 2 * per, 0);
        for (int i = 0; i < n; i++) pic[0] = gl_add(pic[0], lv[start_pulses + 4 + 4 * i]);
        for (int c = 0; c < 2; c++) {
          u64 step = gl_pow(alphas[c], (u64)io_len);
          q.pi_skip[c] = F(gl_pow(step, (u64)(n - 1)));
          for (int i = 0; i < n; i++) {
            u64 ai = gl_pow(step, (u64)(n - 1 - i));
            u64 pin = lv[start_pulses + 2 + 4 * i], pout = lv[start_pulses + 4 + 4 * i];
            u64* col = pic.data() + 1 + c * per;
            col[0] = gl_add(col[0], gl_mul(ai, pin)); col[1] = gl_add(col[1], gl_mul(ai, pout));
            for (int u = 0; u < io_len; u++) { int pidx, kind; pi_map(gc, u, pidx, kind); col[2 + u] = gl_add(col[2 + u], gl_mul(gl_mul(ai, pi[(size_t)i * io_len + pidx]), kind ? pout : pin)); }
          }
        }
        q.pic = pic.data(); q.pic_per_chal = per;
      }
      switch (s.kind) {
        case SEG_SPLIT_RANGE_CHECK: eval_split_u16_range_check(q, s.p0, s.p1, s.p2); break;
        case SEG_MODULAR_CORE: eval_modular_stark_core(q); break;
        case SEG_G1_CORE: eval_exp_core_u32<2>(q, s.p0, s.p1, s.p2); break;
        case SEG_FLAGS: eval_flags(q, s.p0); break;
        case SEG_G1_ADD: eval_g1_add(q, q.lv(s.p1), s.p0); break;
        case SEG_G1_DOUBLE: eval_g1_double(q, q.lv(s.p1), s.p0); break;
        case SEG_PERIODIC_PULSE: eval_periodic_pulse(q, s.p0, s.p1, s.p2, s.p3); break;
        case SEG_PULSE: eval_pulse(q, s.p0, s.p1, s.p2); break;
        case SEG_U16_RANGE_CHECK: eval_u16_range_check(q, s.p0, s.p1); break;
        case SEG_FQ_CORE: eval_exp_core_u32<1>(q, s.p0, s.p1, s.p2); break;
        case SEG_G2_CORE: eval_exp_core_u32<4>(q, s.p0, s.p1, s.p2); break;
        case SEG_FQ_MUL: eval_fq_mul(q, q.lv(s.p0), s.p1 != 0); break;
        case SEG_G2_ADD: eval_g2_add(q, q.lv(s.p1), s.p0); break;
        case SEG_G2_DOUBLE: eval_g2_double(q, q.lv(s.p1), s.p0); break;
        case SEG_FQ12_CORE: eval_fq12_exp_core(q, s.p0, s.p1, s.p2 != 0); break;
        case SEG_FQ12_MUL: {   // k_fq12_products' per-thread body
          std::vector<u64> prod(12 * 31);
          for (int oi = 0; oi < 12; oi++) { F acc[31]; fq12_product_acc(q, 0, s.p1 ? 0 : 192, oi, acc); for (int k = 0; k < 31; k++) prod[oi * 31 + k] = acc[k].v; }
          eval_fq12_mul(q, q.lv(s.p0), prod.data(), 1);
          break;
        }
        case SEG_FLAGS_U64: eval_flags_u64(q, s.p0); break;
        default: return -1;
      }
      for (int c = 0; c < 2; c++) acc[c] = gl_add(gl_mul(acc[c], gl_pow(alphas[c], s.num_constraints)), q.acc[c].v);
      total += s.num_constraints;
    }
    out_acc[0] = acc[0]; out_acc[1] = acc[1];
    if (num_constraints) *num_constraints = total;
    return 0;
  } catch (...) { return -2; }
}
// ModularStark main columns (145) of one row
void emu_modular_row(const u64* io, u64* row) { RowWriter w{row}; modular_stark_row(io, w); }
// G1 main columns a,b,output (384) of one row; returns 0 if the slope denominator vanished
int emu_g1_row(const u32* ax, const u32* ay, const u32* bx, const u32* by, int op, u64* row) { RowWriter w{row}; return g1_row(ax, ay, bx, by, op, w) ? 1 : 0; }
void emu_flags_row(const u32* e, int r, u64* out) { flags_row(e, r, out); }
void emu_poseidon(u64* st) { poseidon_permute(st); }
// Jacobian chain of one instance -> affine canonical words of A[k], B[k], k = 0..256 (A: 2^k x, B: partial sums)
void emu_g1_chain(const u64* x_x, const u64* x_y, const u64* o_x, const u64* o_y, const u32* e, u32* affA /*257*16*/, u32* affB) {
  u32 w[8]; G1Jac A, B;
  u64x4_to_words(x_x, w); A.x = fq_from_words(w); u64x4_to_words(x_y, w); A.y = fq_from_words(w); A.z = fq_one();
  u64x4_to_words(o_x, w); B.x = fq_from_words(w); u64x4_to_words(o_y, w); B.y = fq_from_words(w); B.z = fq_one();
  auto store = [&](const G1Jac& p, u32* out) {
    Fq zi = fq_inv(p.z), zi2 = fq_sqr(zi);
    Fq x = fq_mul(p.x, zi2), y = fq_mul(p.y, fq_mul(zi2, zi));
    fq_to_words(x, out); fq_to_words(y, out + 8);
  };
  store(A, affA); store(B, affB);
  for (int k = 0; k < 256; k++) {
    if ((e[k >> 5] >> (k & 31)) & 1) B = g1_jac_add(A, B);
    A = g1_jac_dbl(A);
    store(A, affA + (k + 1) * 16); store(B, affB + (k + 1) * 16);
  }
}
// FqExp main columns a,b,output (144) of one row
void emu_fq_exp_row(const u32* a, const u32* b, int op, u64* row) { RowWriter w{row}; fq_exp_row(a, b, op, w); }
// G2 main columns a,b,output (768) of one row; returns 0 if the slope denominator vanished
int emu_g2_row(const u32* ax, const u32* ay, const u32* bx, const u32* by, int op, u64* row) { RowWriter w{row}; return g2_row(ax, ay, bx, by, op, w) ? 1 : 0; }
// Jacobian G2 chain of one instance -> affine canonical words (32 per point: x.c0 x.c1 y.c0 y.c1) of A[k], B[k], k = 0..nsteps
void emu_g2_chain(const u64* x /*16*/, const u64* off /*16*/, const u32* e, int nsteps, u32* affA, u32* affB) {
  u32 w[16]; G2Jac A, B;
  auto load = [&](const u64* p, G2Jac& P) {
    for (int t = 0; t < 4; t++) u64x4_to_words(p + 4 * t, w + 0), (t == 0 ? P.x.c0 : t == 1 ? P.x.c1 : t == 2 ? P.y.c0 : P.y.c1) = fq_from_words(w);
    P.z = fq2_one();
  };
  load(x, A); load(off, B);
  auto store = [&](const G2Jac& p, u32* out) {
    Fq2 zi = fq2_inv(p.z), zi2 = fq2_sqr(zi);
    Fq2 xx = fq2_mul(p.x, zi2), yy = fq2_mul(p.y, fq2_mul(zi2, zi));
    fq2_to_words(xx, out); fq2_to_words(yy, out + 16);
  };
  store(A, affA); store(B, affB);
  for (int k = 0; k < nsteps; k++) {
    if ((e[k >> 5] >> (k & 31)) & 1) B = g2_jac_add(A, B);
    A = g2_jac_dbl(A);
    store(A, affA + (k + 1) * 32); store(B, affB + (k + 1) * 32);
  }
}
// Fq12 product in the flat basis through the chain's per-coefficient routine: canonical words in and out (12 x 8)
void emu_fq12_mul(const u32* x, const u32* y, u32* out) {
  Fq xm[12], ym[12];
  for (int i = 0; i < 12; i++) { xm[i] = fq_from_words(x + 8 * i); ym[i] = fq_from_words(y + 8 * i); }
  for (int oi = 0; oi < 12; oi++) fq_to_words(fq12_mul_coeff(xm, ym, oi), out + 8 * oi);
}
// the same product the way k_fq12_chain forms it: 144 pairwise products, then fq12_sum_coeff per output coefficient
void emu_fq12_mul_pairwise(const u32* x, const u32* y, u32* out) {
  Fq xm[12], ym[12], pr[144];
  for (int i = 0; i < 12; i++) { xm[i] = fq_from_words(x + 8 * i); ym[i] = fq_from_words(y + 8 * i); }
  for (int i = 0; i < 12; i++) for (int j = 0; j < 12; j++) pr[12 * i + j] = fq_mul(xm[i], ym[j]);
  for (int oi = 0; oi < 12; oi++) fq_to_words(fq12_sum_coeff(pr, oi), out + 8 * oi);
}
// Fq12 main columns a,b,output (1728) of one row, written the way k_fq12_rows does (one coefficient at a time)
void emu_fq12_row(const u32* a, const u32* b, const u32* out_words, int op, u64* row) {
  RowWriter w{row};
  unsigned short sx[192], sy[192];
  const bool square = op == EXP_OP_SQUARE;
  for (int oi = 0; oi < 12; oi++) for (int i = 0; i < 8; i++) {
    sx[16 * oi + 2 * i] = a[8 * oi + i] & 0xFFFF; sx[16 * oi + 2 * i + 1] = a[8 * oi + i] >> 16;
    u32 yv = square ? a[8 * oi + i] : b[8 * oi + i];
    sy[16 * oi + 2 * i] = yv & 0xFFFF; sy[16 * oi + 2 * i + 1] = yv >> 16;
  }
  for (int oi = 0; oi < 12; oi++) {
    write_limbs16(w, 16 * oi, a + 8 * oi); write_limbs16(w, 192 + 16 * oi, b + 8 * oi);
    fq12_row_coeff(sx, sy, out_words + 8 * oi, oi, op, w);
  }
}
void emu_flags_u64_row(u64 e, int r, u64* out) { flags_u64_row(e, r, out); }
}
