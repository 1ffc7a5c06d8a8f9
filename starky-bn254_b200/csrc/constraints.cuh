// K5 building blocks: per-point evaluation of the reference's constraint programs
// (`Stark::eval_packed_generic` and the gadget `eval_*` functions), restated for a
// thread-per-LDE-point kernel over column-major LDE data.
//
// The starky consumer folds constraints as acc = acc*alpha + c, so only the ORDER and
// the values of the emitted constraints matter.  The AIR is cut into "segments" (one kernel each);
// a segment folds its own constraints with Horner and the caller combines
// acc <- acc * alpha^m + S (m = #constraints in the segment), which is the same field element.
//
// Every function cites the reference function it replays.  Column indices are relative to the AIR
// row; `lv(c)` / `nv(c)` read column c at the local / next LDE point.
#pragma once
#include "gl.cuh"

struct F {
  u64 v;
  HD F() : v(0) {}
  HD explicit F(u64 x) : v(x) {}
};
// On the device the value is an ARBITRARY 64-bit representative (lazy arithmetic of gl.cuh: no conditional subtraction of p
// after each operation, four-multiplication products); it is canonicalised where it leaves the kernel (f_canon in quotient.cu).
// On the host (tests/emu) the same code runs on canonical values.
#ifdef __CUDA_ARCH__
HD F operator+(F a, F b) { return F(gl_add_nc2(a.v, b.v)); }
HD F operator-(F a, F b) { return F(gl_sub_nc2(a.v, b.v)); }
HD F operator*(F a, F b) { return F(gl_mul_nc(a.v, b.v)); }
HD F operator-(F a) { return F(gl_sub_nc2(0, a.v)); }
HD u64 f_canon(F a) { return gl_canon(a.v); }
#else
HD F operator+(F a, F b) { return F(gl_add(a.v, b.v)); }
HD F operator-(F a, F b) { return F(gl_sub(a.v, b.v)); }
HD F operator*(F a, F b) { return F(gl_mul(a.v, b.v)); }
HD F operator-(F a) { return F(gl_neg(a.v)); }
HD u64 f_canon(F a) { return a.v; }
#endif

#define SBN_MAX_CHALLENGES 2

// Evaluation context of one LDE point (plays the role of StarkEvaluationVars + ConstraintConsumer).
struct QPoint {
  const u64* lp;   // &lde[0][local point]
  const u64* np;   // &lde[0][next point]
  size_t stride;   // distance between consecutive columns
  const u64* pi;   // public inputs
  F z_last, l_first, l_last;
  F alpha[SBN_MAX_CHALLENGES];
  F acc[SBN_MAX_CHALLENGES];
  // public-input binding columns at this point (see PiBinding below): pic[col * pic_stride]
  const u64* pic; size_t pic_stride; int pic_per_chal;
  F pi_skip[SBN_MAX_CHALLENGES];   // alpha^((num_io - 1) * io_len)
  HD F picol(int col) const { return F(pic[(size_t)col * pic_stride]); }
  HD F lv(int c) const { return F(lp[(size_t)c * stride]); }
  HD F nv(int c) const { return F(np[(size_t)c * stride]); }
  HD void constraint(F c) {
#pragma unroll
    for (int k = 0; k < SBN_MAX_CHALLENGES; k++) acc[k] = acc[k] * alpha[k] + c;
  }
  // one constraint whose value differs per challenge (the folded public-input binding)
  HD void constraint2(F c0, F c1) { acc[0] = acc[0] * alpha[0] + c0; acc[1] = acc[1] * alpha[1] + c1; }
  HD void transition(F c) { constraint(c * z_last); }
  HD void first_row(F c) { constraint(c * l_first); }
  HD void last_row(F c) { constraint(c * l_last); }
  // n consecutive constraints whose value is identically zero (acc <- acc * alpha^n)
  HD void zeros(int n) {
    for (int i = 0; i < n; i++) {
#pragma unroll
      for (int k = 0; k < SBN_MAX_CHALLENGES; k++) acc[k] = acc[k] * alpha[k];
    }
  }
};

#define BN254_LIMB_LIST {64839, 55420, 35862, 15392, 51853, 26737, 27281, 38785, 22621, 33153, 17846, 47184, 41001, 57649, 20082, 12388}
#define GL_INV_65536 18446462594437939201ULL   // reference src/modular/addcy.rs:13
#define AUX_COEFF_ABS_MAX_U (1ULL << 29)       // reference src/modular/modular.rs:28

HD u64 bn254_limb(int i) {
  const u32 m[16] = BN254_LIMB_LIST;
  return m[i];
}

// Description of the input polynomial of a modular gadget:
//   in(x) = sum_t scale[t] * A[t](x) * B[t](x)  -  E(x)  -  E2(x)          (t < nprod <= 4)
// with A, B, E, E2 16-limb columns of the local row already loaded, scale a small signed integer; or, when
// `direct` is set, the 31 coefficients are read from memory (direct[k * direct_stride]) -- used for the Fq12
// product, whose limb polynomial is produced by a separate kernel.  Field arithmetic is exact, so any
// association of the sums yields the same canonical element as the reference's `pol_*` call sequence.
struct ModInput {
  int nprod;
  const F* A[4]; const F* B[4]; int scale[4];
  const F* E; const F* E2;
  const u64* direct; size_t direct_stride;
};
HD ModInput mod_input1(const F* a, const F* b, int s, const F* e = nullptr, const F* e2 = nullptr) {
  ModInput in; in.nprod = 1; in.A[0] = a; in.B[0] = b; in.scale[0] = s; in.E = e; in.E2 = e2; in.direct = nullptr; in.direct_stride = 0;
  for (int t = 1; t < 4; t++) { in.A[t] = nullptr; in.B[t] = nullptr; in.scale[t] = 0; }
  return in;
}
HD ModInput mod_input2(const F* a0, const F* b0, int s0, const F* a1, const F* b1, int s1, const F* e = nullptr, const F* e2 = nullptr) {
  ModInput in = mod_input1(a0, b0, s0, e, e2); in.nprod = 2; in.A[1] = a1; in.B[1] = b1; in.scale[1] = s1; return in;
}
HD ModInput mod_input4(const F* a0, const F* b0, int s0, const F* a1, const F* b1, int s1, const F* a2, const F* b2, int s2, const F* a3, const F* b3, int s3) {
  ModInput in = mod_input2(a0, b0, s0, a1, b1, s1); in.nprod = 4; in.A[2] = a2; in.B[2] = b2; in.scale[2] = s2; in.A[3] = a3; in.B[3] = b3; in.scale[3] = s3; return in;
}
HD ModInput mod_input_direct(const u64* p, size_t stride) { ModInput in = mod_input1(nullptr, nullptr, 0); in.nprod = 0; in.direct = p; in.direct_stride = stride; return in; }
HD F mod_input_coeff(const ModInput& in, int k) {
  if (in.direct) return F(in.direct[(size_t)k * in.direct_stride]);
  F acc;
  int lo = k > 15 ? k - 15 : 0, hi = k < 15 ? k : 15;
  for (int t = 0; t < in.nprod; t++) {
    F p;
    for (int i = lo; i <= hi; i++) p = p + in.A[t][i] * in.B[t][k - i];
    int sc = in.scale[t];
    int mag = sc < 0 ? -sc : sc;
    if (mag != 1) p = p * F((u64)mag);
    acc = sc < 0 ? acc - p : acc + p;
  }
  if (in.E && k < 16) { acc = acc - in.E[k]; if (in.E2) acc = acc - in.E2[k]; }
  return acc;
}

// reference src/modular/modular.rs:215-230 `eval_modular_op` (incl. `modular_constr_poly` :102-153 and
// `eval_packed_generic_addcy` addcy.rs:16-58): 33 + 1 + 32 constraints.
// Columns: output at out_col (16); aux block at aux_col = out_aux_red(16) | quot_abs(17) | lo(31) | hi(31);
// sign at sign_col.
HD void eval_modular_op(QPoint& q, F filter, const ModInput& in, const F* output /*16 loaded*/, int aux_col, int sign_col) {
  const F overflow(1ULL << 16), overflow_inv(GL_INV_65536);
  // addcy(modulus, out_aux_red, output, is_less_than=[1,0,..,0])
  F cy;
  for (int i = 0; i < 16; i++) {
    F t = cy + F(bn254_limb(i)) + q.lv(aux_col + i) - output[i];
    q.constraint(filter * t * (overflow - t));
    cy = t * overflow_inv;
  }
  q.zeros(1);                                  // filter * given_cy[0] * (given_cy[0] - 1), given_cy[0] = 1
  q.constraint(filter * (cy - F(1)));
  q.zeros(15);                                 // filter * given_cy[i], i >= 1
  F sign = q.lv(sign_col);
  q.constraint(filter * (sign * sign - F(1)));
  // constr_poly[k] = sum_{i+j=k} (sign*quot_abs[i]) * m[j] + output[k] + (x - beta) * aux(x) [k] - input[k]
  F quot[17];
  for (int i = 0; i < 17; i++) quot[i] = sign * q.lv(aux_col + 16 + i);
  const F base(1ULL << 16), offset(AUX_COEFF_ABS_MAX_U);
  F aux_prev;  // aux_poly[k-1]
  for (int k = 0; k < 32; k++) {
    F c;
    int lo = k > 15 ? k - 15 : 0, hi = k < 16 ? k : 16;
    for (int i = lo; i <= hi; i++) c = c + quot[i] * F(bn254_limb(k - i));
    if (k < 16) c = c + output[k];
    F aux_k;
    if (k < 31) aux_k = q.lv(aux_col + 33 + k) - offset + base * q.lv(aux_col + 64 + k);
    // pol_adjoin_root: res[0] = -root*a[0]; res[k] = a[k-1] - root*a[k]
    c = c + (k == 0 ? -(base * aux_k) : aux_prev - base * aux_k);
    aux_prev = aux_k;
    if (k < 31) c = c - mod_input_coeff(in, k);
    q.constraint(filter * c);
  }
}

// reference src/modular/modular_zero.rs:82-120 `eval_modular_zero`: 1 + 32 constraints.
// aux block at aux_col = quot_abs(17) | lo(31) | hi(31).
HD void eval_modular_zero(QPoint& q, F filter, const ModInput& in, int aux_col, int sign_col) {
  F sign = q.lv(sign_col);
  q.constraint(filter * (sign * sign - F(1)));
  F quot[17];
  for (int i = 0; i < 17; i++) quot[i] = sign * q.lv(aux_col + i);
  const F base(1ULL << 16), offset(AUX_COEFF_ABS_MAX_U);
  F aux_prev;
  for (int k = 0; k < 32; k++) {
    F c;
    int lo = k > 15 ? k - 15 : 0, hi = k < 16 ? k : 16;
    for (int i = lo; i <= hi; i++) c = c + quot[i] * F(bn254_limb(k - i));
    F aux_k;
    if (k < 31) aux_k = q.lv(aux_col + 17 + k) - offset + base * q.lv(aux_col + 48 + k);
    c = c + (k == 0 ? -(base * aux_k) : aux_prev - base * aux_k);
    aux_prev = aux_k;
    if (k < 31) c = c - mod_input_coeff(in, k);
    q.constraint(filter * c);
  }
}

HD void load16(const QPoint& q, int col, F* out) { for (int i = 0; i < 16; i++) out[i] = q.lv(col + i); }
HD void load16n(const QPoint& q, int col, F* out) { for (int i = 0; i < 16; i++) out[i] = q.nv(col + i); }

// reference src/utils/lookup.rs:13-34 `eval_lookups`
HD void eval_lookups(QPoint& q, int col_in, int col_tab) {
  F local_in = q.lv(col_in), next_tab = q.nv(col_tab), next_in = q.nv(col_in);
  F d_prev = next_in - local_in, d_tab = next_in - next_tab;
  q.constraint(d_prev * d_tab);
  q.last_row(d_tab);
}

// reference src/utils/flags.rs:136-195 `eval_flags`: 26 constraints.
HD void eval_flags(QPoint& q, int sf) {
  const int is_final_c = sf, is_rotate_c = sf + 1, a = sf + 2, b = sf + 3, fbit = sf + 4, bit_c = sf + 5, sl = sf + 6, el = sl + 8;
  const F one(1);
  q.first_row(q.lv(a));
  q.first_row(q.lv(b) - one);
  F bit = q.lv(bit_c);
  q.constraint(bit * bit - bit);
  q.constraint(bit * q.lv(b) - q.lv(fbit));
  q.constraint(q.lv(is_rotate_c) * q.lv(a));
  q.constraint(q.lv(is_final_c) * q.lv(is_rotate_c));
  q.transition(q.lv(a) + q.nv(a) - one);
  q.transition(q.lv(b) + q.nv(b) - one);
  F first_limb = q.lv(sl), next_first_limb = q.nv(sl), next_bit = q.nv(bit_c), is_split = q.lv(a), is_final = q.lv(is_final_c);
  F is_not_final = one - is_final;
  q.transition(is_not_final * is_split * (first_limb - F(2) * next_first_limb - next_bit));
  F is_not_split = one - is_split, is_rotate = q.lv(is_rotate_c), nrf = one - is_rotate - is_final;
  q.transition(is_not_split * (next_bit - bit));
  q.transition(nrf * is_not_split * (first_limb - next_first_limb));
  for (int c = sl + 1; c < el; c++) q.transition(is_rotate * (q.nv(c - 1) - q.lv(c)));
  q.transition(is_rotate * q.nv(el - 1));
  for (int c = sl + 1; c < el; c++) q.transition(nrf * (q.nv(c) - q.lv(c)));
}

// reference src/utils/pulse.rs:146-170 `eval_periodic_pulse`: 5 constraints.
HD void eval_periodic_pulse(QPoint& q, int pulse_col, int start, int period, int first_pulse) {
  const F one(1);
  F counter = q.lv(start), witness = q.lv(start + 1), is_reset = q.lv(pulse_col), next_counter = q.nv(start);
  q.first_row(counter - F((u64)(period - first_pulse - 1)));
  q.transition((one - is_reset) * (next_counter - counter - one));
  q.transition(is_reset * next_counter);
  F delta = counter - F((u64)(period - 1));
  q.constraint(delta * witness + is_reset - one);
  q.constraint(delta * is_reset);
}

// reference src/utils/pulse.rs:45-63 `eval_pulse` with positions = [rows_per_io*i, rows_per_io*i + rows_per_io-1]
// (reference src/curves/g1/exp.rs:153-163 `get_pulse_positions`): 2 + 4*num_io constraints.
HD void eval_pulse(QPoint& q, int start, int num_io, int rows_per_io) {
  const F one(1);
  F counter = q.lv(start);
  q.first_row(counter);
  q.transition(q.nv(start) - counter - one);
  for (int i = 0; i < 2 * num_io; i++) {
    u64 pos = (u64)(i >> 1) * rows_per_io + ((i & 1) ? rows_per_io - 1 : 0);
    F cmp = counter - F(pos);
    F witness = q.lv(start + 1 + 2 * i), pulse = q.lv(start + 2 + 2 * i);
    q.constraint(cmp * witness + pulse - one);
    q.constraint(cmp * pulse);
  }
}

// reference src/utils/range_check.rs:49-68 `eval_u16_range_check`: 2*ntargets + 3 constraints.
HD void eval_u16_range_check(QPoint& q, int start, int ntargets) {
  for (int i = start + 1; i < start + 1 + 2 * ntargets; i += 2) eval_lookups(q, i, i + 1);
  F cur = q.lv(start), next = q.nv(start);
  q.first_row(cur);
  F incr = next - cur;
  q.transition(incr * incr - incr);
  q.last_row(cur - F((1 << 16) - 1));
}

// The same constraint list cut into item ranges for small evaluation domains (quotient.cu k_segment_chunked): run 0 = the
// `original - (lo + 256 hi)` constraints of targets [i0, i1), run 1 = their four lookup constraints each, run 2 = the three
// table-column constraints that close the list.
HD void eval_split_u16_range_check_run(QPoint& q, int main_col, int t0, int i0, int i1, int run) {
  if (run == 0) {
    for (int i = i0; i < i1; i++) {
      F original = q.lv(t0 + i), lo = q.lv(main_col + 1 + 6 * i), hi = q.lv(main_col + 4 + 6 * i);
      q.constraint(original - (lo + hi * F(1 << 8)));
    }
  } else if (run == 1) {
    for (int i = main_col + 1 + 6 * i0; i < main_col + 1 + 6 * i1; i += 6) { eval_lookups(q, i + 1, i + 2); eval_lookups(q, i + 4, i + 5); }
  } else {
    F cur = q.lv(main_col), next = q.nv(main_col);
    q.first_row(cur);
    F incr = next - cur;
    q.transition(incr * incr - incr);
    q.last_row(cur - F((1 << 8) - 1));
  }
}
// reference src/utils/range_check.rs:162-192 `eval_split_u16_range_check`: 5*ntargets + 3 constraints.
HD void eval_split_u16_range_check(QPoint& q, int main_col, int t0, int ntargets) {
  for (int i = 0; i < ntargets; i++) {
    F original = q.lv(t0 + i), lo = q.lv(main_col + 1 + 6 * i), hi = q.lv(main_col + 4 + 6 * i);
    q.constraint(original - (lo + hi * F(1 << 8)));
  }
  for (int i = main_col + 1; i < main_col + 1 + 6 * ntargets; i += 6) { eval_lookups(q, i + 1, i + 2); eval_lookups(q, i + 4, i + 5); }
  F cur = q.lv(main_col), next = q.nv(main_col);
  q.first_row(cur);
  F incr = next - cur;
  q.transition(incr * incr - incr);
  q.last_row(cur - F((1 << 8) - 1));
}

// ---- ModularStark (reference src/modular/modular.rs:440-482), the part after the range check ----
HD void eval_modular_stark_core(QPoint& q) {
  F in0[16], in1[16], out[16];
  load16(q, 0, in0); load16(q, 16, in1); load16(q, 32, out);
  ModInput in = mod_input1(in0, in1, 1);
  F filter = q.lv(48 + 95 + 1);
  eval_modular_op(q, filter, in, out, 48, 48 + 95);
}

// ---- G1 (reference src/curves/g1/muladd.rs) ----
// G1Output block at column o: lambda(16) new_x(16) new_y(16) aux_zero(79) aux_x(95) aux_y(95) sign_zero sign_x sign_y
#define G1O_LAMBDA 0
#define G1O_NEW_X 16
#define G1O_NEW_Y 32
#define G1O_AUX_ZERO 48
#define G1O_AUX_X 127
#define G1O_AUX_Y 222
#define G1O_SIGN_ZERO 317
#define G1O_SIGN_X 318
#define G1O_SIGN_Y 319

// shared tail of eval_g1_add / eval_g1_double (muladd.rs:199-229 / :317-341)
// `part`: 0 = the whole gadget; 1 / 2 / 3 = only its zero / new_x / new_y modular operation (consecutive runs of the constraint
// list: the quotient kernel launches them as separate instantiations so that none carries the others' registers).
HD void eval_g1_tail(QPoint& q, F filter, int o, const F* lambda, const F* x1, const F* x2, const F* y1, int part = 0) {
  F new_x[16];
  load16(q, o + G1O_NEW_X, new_x);
  if (part == 0 || part == 2) {
    // new_x_input = lambda^2 - (x1 + x2)
    ModInput inx = mod_input1(lambda, lambda, 1, x1, x2);
    eval_modular_op(q, filter, inx, new_x, o + G1O_AUX_X, o + G1O_SIGN_X);
  }
  if (part == 0 || part == 3) {
    // new_y_input = lambda * (x1 - new_x) - y1
    F d[16], new_y[16];
    load16(q, o + G1O_NEW_Y, new_y);
    for (int i = 0; i < 16; i++) d[i] = x1[i] - new_x[i];
    ModInput iny = mod_input1(lambda, d, 1, y1);
    eval_modular_op(q, filter, iny, new_y, o + G1O_AUX_Y, o + G1O_SIGN_Y);
  }
}
// reference src/curves/g1/muladd.rs:179-230 `eval_g1_add`: 33 + 66 + 66 constraints.  a at cols 0..31, b at 32..63.
HD void eval_g1_add(QPoint& q, F filter, int o, int part = 0) {
  F ax[16], ay[16], bx[16], lambda[16];
  load16(q, 0, ax); load16(q, 16, ay); load16(q, 32, bx); load16(q, o + G1O_LAMBDA, lambda);
  if (part == 0 || part == 1) {
    F by[16], dx[16], dy[16];
    load16(q, 48, by);
    for (int i = 0; i < 16; i++) { dx[i] = bx[i] - ax[i]; dy[i] = by[i] - ay[i]; }
    // zero_pol = lambda * delta_x - delta_y
    ModInput inz = mod_input1(lambda, dx, 1, dy);
    eval_modular_zero(q, filter, inz, o + G1O_AUX_ZERO, o + G1O_SIGN_ZERO);
  }
  if (part != 1) eval_g1_tail(q, filter, o, lambda, ax, bx, ay, part);
}
// reference src/curves/g1/muladd.rs:291-342 `eval_g1_double`: 165 constraints.
HD void eval_g1_double(QPoint& q, F filter, int o, int part = 0) {
  F x[16], y[16], lambda[16];
  load16(q, 0, x); load16(q, 16, y); load16(q, o + G1O_LAMBDA, lambda);
  if (part == 0 || part == 1) {
    // zero_pol = 2*lambda*y - 3*x*x
    ModInput inz = mod_input2(lambda, y, 2, x, x, -3);
    eval_modular_zero(q, filter, inz, o + G1O_AUX_ZERO, o + G1O_SIGN_ZERO);
  }
  if (part != 1) eval_g1_tail(q, filter, o, lambda, x, x, y, part);
}

// ---- public-input binding of the exponentiation AIRs, folded over the instances ----
// The reference emits, for every instance i, io_len constraints  pulse_{kind(u), i} * (PI_{i,u} - V_u)  (u < io_len;
// kind = input or output pulse; V_u an expression of the local row that does not depend on i), e.g.
// src/curves/g1/exp.rs:374-392.  Under the consumer's Horner fold (acc = acc * alpha + c) constraint (i, u) carries the
// weight a_i * alpha^(io_len - 1 - u) * alpha^(#later constraints) with a_i = alpha^((n - 1 - i) io_len), so the whole
// block equals io_len "virtual" constraints
//        W_u = U_u(x) - V_u(x) * S_kind(u)(x),   U_u = sum_i a_i PI_{i,u} pulse_{kind(u), i},   S_kind = sum_i a_i pulse_{kind, i}
// folded after multiplying the accumulator by alpha^((n - 1) io_len).  On the trace domain a pulse column is 1 at its
// row and 0 elsewhere, so U_u and S_kind are the low-degree extensions of SPARSE columns (a_i PI_{i,u} at the pulse
// rows): the prover builds those io_len + 2 columns per challenge (+ one shared column, the plain sum of the output
// pulses), transforms them with the NTT kernels and hands their values at this point in `q.pic`:
//   col 0: sum_i pulse_out_i;  per challenge c: base = 1 + c * (io_len + 2): S_in, S_out, U_0 .. U_{io_len-1}.
// Same field element as the reference's n * io_len constraints, at 1/n of the per-point work.
// emission index u -> (index inside the instance's public-input record, pulse kind 0 = input / 1 = output).
// group_count: G for the u32 cores (Fq 1, G1 2, G2 4); 0 = Fq12 (io record 584), -1 = Fq12 with u64 exponent (577).
HD void pi_map(int group_count, int u, int& pi, int& kind) {
  if (group_count > 0) {
    const int G = group_count;
    if (u < 16 * G) { pi = u; kind = 0; }
    else if (u < 24 * G) { pi = u + 8; kind = 1; }
    else { pi = u - 8 * G; kind = 0; }
  } else {
    const int out_off = group_count == 0 ? 392 : 385;
    if (u >= 576) { pi = 384 + (u - 576); kind = 0; return; }
    const int c = u / 48, r = u % 48;
    if (r < 16) { pi = 16 * c + r; kind = 0; }
    else if (r < 32) { pi = 192 + 16 * c + (r - 16); kind = 0; }
    else { pi = out_off + 16 * c + (r - 32); kind = 1; }
  }
}
HD int pi_io_len(int group_count) { return group_count > 0 ? (3 * group_count + 1) * 8 : (group_count == 0 ? 584 : 577); }
HD void pi_virtual(QPoint& q, int u, int kind, F v) {
  const int b0 = 1, b1 = 1 + q.pic_per_chal;
  q.constraint2(q.picol(b0 + 2 + u) - v * q.picol(b0 + kind), q.picol(b1 + 2 + u) - v * q.picol(b1 + kind));
}
HD void pi_begin(QPoint& q, F is_final) {
  q.constraint(is_final - q.picol(0));
#pragma unroll
  for (int k = 0; k < SBN_MAX_CHALLENGES; k++) q.acc[k] = q.acc[k] * q.pi_skip[k];
}

// ---- generic exponentiation-AIR core for the AIRs whose public inputs are u32 limbs (Fq, G1, G2) ----
// reference src/fields/fq/exp.rs:289-359, src/curves/g1/exp.rs:340-461, src/curves/g2/exp.rs:354-472:
// is_final constraint, public-input binding, a/b transitions.  The row starts with a (G groups of 16 limbs)
// and b (G groups); `newv` is the column of the G-group value the operation produces (Fq: output, G1/G2:
// new_x | new_y); per-io public inputs are x[G][8] offset[G][8] exp_val[8] output[G][8].
// 1 + (3G + 1) * 8 * num_io + 3 * 32 * G constraints.
template <int G> HD void eval_exp_core_u32(QPoint& q, int num_io, int newv, int sf) {
  const F one(1), base(1ULL << 16);
  F is_add = q.lv(sf + 4), is_double = q.lv(sf + 2), is_final = q.lv(sf), is_not_final = one - is_final;
  pi_begin(q, is_final);
  // emission order per instance: x groups (input pulse), offset groups (input), output groups (output pulse), exp_val (input)
  for (int g = 0; g < G; g++) for (int j = 0; j < 8; j++) pi_virtual(q, 8 * g + j, 0, q.lv(16 * g + 2 * j) + base * q.lv(16 * g + 2 * j + 1));
  for (int g = 0; g < G; g++) for (int j = 0; j < 8; j++) pi_virtual(q, 8 * (G + g) + j, 0, q.lv(16 * (G + g) + 2 * j) + base * q.lv(16 * (G + g) + 2 * j + 1));
  for (int g = 0; g < G; g++) for (int j = 0; j < 8; j++) pi_virtual(q, 8 * (2 * G + g) + j, 1, q.lv(16 * (G + g) + 2 * j) + base * q.lv(16 * (G + g) + 2 * j + 1));
  for (int j = 0; j < 8; j++) {
    F limb = q.lv(sf + 6 + j);
    if (j == 0) limb = limb * F(2) + is_add;
    pi_virtual(q, 24 * G + j, 0, limb);
  }
  const int W = 16 * G;
  F f1 = is_not_final * is_double, f2 = is_not_final * is_add, f3 = is_not_final * (one - is_double - is_add);
  for (int i = 0; i < W; i++) q.transition(f1 * (q.nv(i) - q.lv(newv + i)));
  for (int i = 0; i < W; i++) q.transition(f1 * (q.nv(W + i) - q.lv(W + i)));
  for (int i = 0; i < W; i++) q.transition(f2 * (q.nv(i) - q.lv(i)));
  for (int i = 0; i < W; i++) q.transition(f2 * (q.nv(W + i) - q.lv(newv + i)));
  for (int i = 0; i < 2 * W; i++) q.transition(f3 * (q.nv(i) - q.lv(i)));
}

// ---- Fq (reference src/fields/fq/mul.rs:69-87 `eval_fq_mul`): a at 0, b at 16, FqOutput at 32 = output16 | aux95 | sign ----
HD void eval_fq_mul(QPoint& q, F filter, bool square) {
  F a[16], b[16], out[16];
  load16(q, 0, a); load16(q, square ? 0 : 16, b); load16(q, 32, out);
  ModInput in = mod_input1(a, b, 1);
  eval_modular_op(q, filter, in, out, 48, 48 + 95);
}

// ---- G2 (reference src/curves/g2/muladd.rs) ----
// Row: a_x(c0,c1) a_y b_x b_y = columns 0..127; G2Output block at o = 128:
// lambda32 new_x32 new_y32 | aux_zero 2x79 | aux 4x95 | sign_zero x2 | sign x4   (muladd.rs:56-80)
#define G2O_LAMBDA 0
#define G2O_NEW_X 32
#define G2O_NEW_Y 64
#define G2O_AUX_ZERO 96
#define G2O_AUX 254
#define G2O_SIGN_ZERO 634
#define G2O_SIGN 636
// shared tail of eval_g2_add / eval_g2_double (muladd.rs:233-260 / :446-471): x1, y1 are the "a" operands
HD void eval_g2_tail(QPoint& q, F filter, int o, const F* l0, const F* l1, const F* x1c0, const F* x1c1, const F* x2c0, const F* x2c1, const F* y1c0, const F* y1c1,
                     int part = 0) {
  F nx0[16], nx1[16];
  load16(q, o + G2O_NEW_X, nx0); load16(q, o + G2O_NEW_X + 16, nx1);
  if (part == 0 || part == 2) {
    // new_x_input = lambda^2 - (x1 + x2):  c0 = l0 l0 - l1 l1, c1 = l0 l1 + l1 l0
    { ModInput in = mod_input2(l0, l0, 1, l1, l1, -1, x1c0, x2c0); eval_modular_op(q, filter, in, nx0, o + G2O_AUX, o + G2O_SIGN); }
    { ModInput in = mod_input1(l0, l1, 2, x1c1, x2c1); eval_modular_op(q, filter, in, nx1, o + G2O_AUX + 95, o + G2O_SIGN + 1); }
  }
  if (part == 0 || part == 3) {
    // new_y_input = lambda * (x1 - new_x) - y1
    F d0[16], d1[16], ny0[16], ny1[16];
    for (int i = 0; i < 16; i++) { d0[i] = x1c0[i] - nx0[i]; d1[i] = x1c1[i] - nx1[i]; }
    load16(q, o + G2O_NEW_Y, ny0); load16(q, o + G2O_NEW_Y + 16, ny1);
    { ModInput in = mod_input2(l0, d0, 1, l1, d1, -1, y1c0); eval_modular_op(q, filter, in, ny0, o + G2O_AUX + 190, o + G2O_SIGN + 2); }
    { ModInput in = mod_input2(l0, d1, 1, l1, d0, 1, y1c1); eval_modular_op(q, filter, in, ny1, o + G2O_AUX + 285, o + G2O_SIGN + 3); }
  }
}
// reference src/curves/g2/muladd.rs:416-472 `eval_g2_add`: 2*33 + 4*66 = 330 constraints
HD void eval_g2_add(QPoint& q, F filter, int o, int part = 0) {
  F ax0[16], ax1[16], ay0[16], ay1[16], bx0[16], bx1[16], l0[16], l1[16];
  load16(q, 0, ax0); load16(q, 16, ax1); load16(q, 32, ay0); load16(q, 48, ay1); load16(q, 64, bx0); load16(q, 80, bx1);
  load16(q, o + G2O_LAMBDA, l0); load16(q, o + G2O_LAMBDA + 16, l1);
  if (part == 0 || part == 1) {
    F dx0[16], dx1[16], dy0[16], dy1[16];
    for (int i = 0; i < 16; i++) { dx0[i] = bx0[i] - ax0[i]; dx1[i] = bx1[i] - ax1[i]; dy0[i] = q.lv(96 + i) - ay0[i]; dy1[i] = q.lv(112 + i) - ay1[i]; }
    // zero_pol = lambda * delta_x - delta_y
    { ModInput in = mod_input2(l0, dx0, 1, l1, dx1, -1, dy0); eval_modular_zero(q, filter, in, o + G2O_AUX_ZERO, o + G2O_SIGN_ZERO); }
    { ModInput in = mod_input2(l0, dx1, 1, l1, dx0, 1, dy1); eval_modular_zero(q, filter, in, o + G2O_AUX_ZERO + 79, o + G2O_SIGN_ZERO + 1); }
  }
  if (part != 1) eval_g2_tail(q, filter, o, l0, l1, ax0, ax1, bx0, bx1, ay0, ay1, part);
}
// reference src/curves/g2/muladd.rs:203-261 `eval_g2_double`: 330 constraints
HD void eval_g2_double(QPoint& q, F filter, int o, int part = 0) {
  F x0[16], x1[16], y0[16], y1[16], l0[16], l1[16];
  load16(q, 0, x0); load16(q, 16, x1); load16(q, 32, y0); load16(q, 48, y1);
  load16(q, o + G2O_LAMBDA, l0); load16(q, o + G2O_LAMBDA + 16, l1);
  if (part == 0 || part == 1) {
    // zero_pol = 2 * lambda * y - 3 * x * x
    { ModInput in = mod_input4(l0, y0, 2, l1, y1, -2, x0, x0, -3, x1, x1, 3); eval_modular_zero(q, filter, in, o + G2O_AUX_ZERO, o + G2O_SIGN_ZERO); }
    { ModInput in = mod_input4(l0, y1, 2, l1, y0, 2, x0, x1, -3, x1, x0, -3); eval_modular_zero(q, filter, in, o + G2O_AUX_ZERO + 79, o + G2O_SIGN_ZERO + 1); }
  }
  if (part != 1) eval_g2_tail(q, filter, o, l0, l1, x0, x1, x0, x1, y0, y1, part);
}

// ---- Fq12 (reference src/fields/fq12/mul.rs, src/fields/fq12/exp.rs, src/fields/fq12_u64) ----
// Row: a (12x16) at 0, b at 192, Fq12Output at 384 = output 12x16 | aux 12x95 | sign x12   (mul.rs:219-231)
#define FQ12O_AUX 192
#define FQ12O_SIGN (192 + 12 * 95)
// reference src/fields/fq12/mul.rs:254-275 `eval_fq12_mul`: 12 * 66 constraints.  `prod` holds the limb polynomial of
// pol_mul_fq12(x, y, 9) for this point: coefficient k of output i at prod[(i * 31 + k) * prod_stride].
HD void eval_fq12_mul(QPoint& q, F filter, const u64* prod, size_t prod_stride) {
  const int o = 384;
  for (int i = 0; i < 12; i++) {
    F out[16];
    load16(q, o + 16 * i, out);
    ModInput in = mod_input_direct(prod + (size_t)i * 31 * prod_stride, prod_stride);
    eval_modular_op(q, filter, in, out, o + FQ12O_AUX + 95 * i, o + FQ12O_SIGN + i);
  }
}
// pol_mul_fq12(x, y, 9) (reference src/fields/fq12/mul.rs:24-87): the 31 limbs of output coefficient `oi` (flat MyFq12
// order) for x at column xa and y at column ya of the local row, accumulated while the contributing (x_i, y_j)
// limb pairs stream through.
//   re[m] = sum_{i+j=m} (x_i y_j - x_{i+6} y_{j+6}),  im[m] = sum_{i+j=m} (x_i y_{j+6} + x_{i+6} y_j)
//   out[i] = re[i] + 9 re[i+6] - im[i+6],  out[i+6] = im[i] + re[i+6] + 9 im[i+6]   (i < 5);  out[5] = re[5], out[11] = im[5]
HD void fq12_product_acc(const QPoint& q, int xa, int ya, int oi, F* acc /*31*/) {
#pragma unroll
  for (int k = 0; k < 31; k++) acc[k] = F();
  const bool imag = oi >= 6; const int i0 = imag ? oi - 6 : oi;
  // terms: (m, part, weight): part 0 = re, 1 = im
  for (int term = 0; term < 3; term++) {
    int m, part, w;
    if (term == 0) { m = i0; part = imag ? 1 : 0; w = 1; }
    else if (i0 == 5) break;
    else if (term == 1) { m = i0 + 6; part = 0; w = imag ? 1 : 9; }
    else { m = i0 + 6; part = 1; w = imag ? 9 : -1; }
    for (int i = 0; i < 6; i++) {
      const int j = m - i;
      if (j < 0 || j > 5) continue;
      for (int half = 0; half < 2; half++) {
        // re: (x_i, y_j, +), (x_{i+6}, y_{j+6}, -);  im: (x_i, y_{j+6}, +), (x_{i+6}, y_j, +)
        const int xi = half ? i + 6 : i;
        const int yj = part == 0 ? (half ? j + 6 : j) : (half ? j : j + 6);
        const bool neg = (part == 0 && half == 1) != (w < 0);
        F x[16], y[16];
#pragma unroll
        for (int t = 0; t < 16; t++) { x[t] = q.lv(xa + 16 * xi + t); y[t] = q.lv(ya + 16 * yj + t); }
        if (w == 9 || w == -9) {
#pragma unroll
          for (int t = 0; t < 16; t++) x[t] = x[t] * F(9);
        }
        if (neg) {
#pragma unroll
          for (int t = 0; t < 16; t++) x[t] = -x[t];
        }
#ifdef __CUDA_ARCH__
        // one anti-diagonal at a time: up to 16 products summed unreduced in 160 bits (gl_acc), one reduction per coefficient
        // instead of one per product
#pragma unroll
        for (int k = 0; k < 31; k++) {
          gl_acc d = gl_acc_zero();
#pragma unroll
          for (int s = (k > 15 ? k - 15 : 0); s <= (k < 15 ? k : 15); s++) gl_acc_mac(d, x[s].v, y[k - s].v);
          acc[k] = acc[k] + F(gl_acc_reduce(d));
        }
#else
#pragma unroll
        for (int s = 0; s < 16; s++)
#pragma unroll
          for (int t = 0; t < 16; t++) acc[s + t] = acc[s + t] + x[s] * y[t];
#endif
      }
    }
  }
}
// reference src/fields/fq12/exp.rs:340-393 (u64 variant: src/fields/fq12_u64/exp_u64.rs:331-383): is_final, public inputs
// (u16 limbs: x[12][16] offset[12][16] exp output[12][16]), transitions.  1 + io_len * num_io + 6 * 192 constraints;
// the public-input block is folded over the instances as described above (emission order per instance: for each
// coefficient c: x[c] (input pulse), offset[c] (input), output[c] (output pulse); then the exponent (input)).
HD void eval_fq12_exp_core(QPoint& q, int num_io, int sf, bool u64_variant) {
  const int is_sq_c = u64_variant ? sf + 1 : sf + 2, is_mul_c = u64_variant ? sf + 3 : sf + 4;
  const F one(1);
  F is_mul = q.lv(is_mul_c), is_sq = q.lv(is_sq_c), is_final = q.lv(sf), is_not_final = one - is_final;
  pi_begin(q, is_final);
  for (int c = 0; c < 12; c++) {
    for (int j = 0; j < 16; j++) pi_virtual(q, 48 * c + j, 0, q.lv(16 * c + j));
    for (int j = 0; j < 16; j++) pi_virtual(q, 48 * c + 16 + j, 0, q.lv(192 + 16 * c + j));
    for (int j = 0; j < 16; j++) pi_virtual(q, 48 * c + 32 + j, 1, q.lv(192 + 16 * c + j));
  }
  if (u64_variant) {
    pi_virtual(q, 576, 0, q.lv(sf + 5) * F(2) + is_mul);
  } else {
    for (int j = 0; j < 8; j++) {
      F limb = q.lv(sf + 6 + j);
      if (j == 0) limb = limb * F(2) + is_mul;
      pi_virtual(q, 576 + j, 0, limb);
    }
  }
  F f1 = is_not_final * is_sq, f2 = is_not_final * is_mul, f3 = is_not_final * (one - is_sq - is_mul);
  for (int i = 0; i < 192; i++) q.transition(f1 * (q.nv(i) - q.lv(384 + i)));
  for (int i = 0; i < 192; i++) q.transition(f1 * (q.nv(192 + i) - q.lv(192 + i)));
  for (int i = 0; i < 192; i++) q.transition(f2 * (q.nv(i) - q.lv(i)));
  for (int i = 0; i < 192; i++) q.transition(f2 * (q.nv(192 + i) - q.lv(384 + i)));
  for (int i = 0; i < 384; i++) q.transition(f3 * (q.nv(i) - q.lv(i)));
}
// reference src/fields/fq12_u64/flags_u64.rs:96-139 `eval_flags_u64`: 9 constraints.  is_final a b filtered_bit bit val
HD void eval_flags_u64(QPoint& q, int sf) {
  const int a = sf + 1, b = sf + 2, fbit = sf + 3, bit_c = sf + 4, val = sf + 5;
  const F one(1);
  q.first_row(q.lv(a));
  q.first_row(q.lv(b) - one);
  F bit = q.lv(bit_c);
  q.constraint(bit * bit - bit);
  q.constraint(bit * q.lv(b) - q.lv(fbit));
  q.transition(q.lv(a) + q.nv(a) - one);
  q.transition(q.lv(b) + q.nv(b) - one);
  F first_limb = q.lv(val), next_first_limb = q.nv(val), next_bit = q.nv(bit_c), is_split = q.lv(a), is_final = q.lv(sf);
  F is_not_final = one - is_final;
  q.transition(is_not_final * is_split * (first_limb - F(2) * next_first_limb - next_bit));
  F is_not_split = one - is_split;
  q.transition(is_not_split * (next_bit - bit));
  q.transition(is_not_final * is_not_split * (first_limb - next_first_limb));
}
