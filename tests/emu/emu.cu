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
      switch (s.kind) {
        case SEG_SPLIT_RANGE_CHECK: eval_split_u16_range_check(q, s.p0, s.p1, s.p2); break;
        case SEG_MODULAR_CORE: eval_modular_stark_core(q); break;
        case SEG_G1_CORE: eval_g1_exp_core(q, s.p0); break;
        case SEG_FLAGS: eval_flags(q, s.p0); break;
        case SEG_G1_ADD: eval_g1_add(q, q.lv(s.p1), s.p0); break;
        case SEG_G1_DOUBLE: eval_g1_double(q, q.lv(s.p1), s.p0); break;
        case SEG_PERIODIC_PULSE: eval_periodic_pulse(q, s.p0, s.p1, s.p2, s.p3); break;
        case SEG_PULSE: eval_pulse(q, s.p0, s.p1, s.p2); break;
        case SEG_U16_RANGE_CHECK: eval_u16_range_check(q, s.p0, s.p1); break;
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
}
