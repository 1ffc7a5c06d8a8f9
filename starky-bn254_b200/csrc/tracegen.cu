// K1: batched trace generation on the device.  Replaces the reference's serial, BigInt-based
// `generate_trace` (reference src/curves/g1/exp.rs:255-318, src/modular/modular.rs:379-434) with:
//   chain kernel  (one thread per instance, Jacobian double-and-add, no inversions)
//   affine kernel (one thread per chain point)
//   row kernel    (one thread per trace row: slope, limb products, quotient and carry witnesses)
//   pulse / lookup kernels (closed forms, counting sort, and the reference's greedy table permutation).
// The trace is written column-major, so the threads of a warp (32 consecutive rows) store 256
// contiguous bytes per column.
#include "tracegen.cuh"
#include "witness.cuh"
#include "../../include/starky_bn254_b200.h"

struct ColWriter {
  u64* base; size_t stride;
  __device__ __forceinline__ void operator()(int col, u64 v) const { base[(size_t)col * stride] = v; }
};

// ---------------- lookups (reference src/utils/range_check.rs, src/utils/lookup.rs:60-111) ----------------
struct LookupDesc { int src_col, shift, sorted_col, perm_col; };

__global__ void k_table_col(u64* col, size_t N, u32 R) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r < N) col[r] = r < R ? r : R - 1;
}
// target t0 + i (grid.y = i) -> its low and high byte columns at main_col + 1 + 6 i and + 3 more (one launch for all targets:
// one launch per target was 1 332 launches and 5.4 ms of launch latency per Fq12 proof)
__global__ void k_split_cols(u64* cols, size_t N, int t0, int main_col) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  const int i = blockIdx.y, b = main_col + 1 + 6 * i;
  const u64 v = cols[(size_t)(t0 + i) * N + r];
  cols[(size_t)b * N + r] = v & 0xFF;
  cols[(size_t)(b + 3) * N + r] = (v >> 8) & 0xFF;
}
// histogram of (col >> shift) & (R-1); grid.y = lookup
__global__ void __launch_bounds__(256) k_lookup_hist(const u64* __restrict__ cols, size_t N, u32 R, const LookupDesc* __restrict__ descs, u32* __restrict__ cnt,
                                                     int* __restrict__ err) {
  const LookupDesc d = descs[blockIdx.y];
  const u64* src = cols + (size_t)d.src_col * N;
  u32* c = cnt + (size_t)blockIdx.y * R;
  __shared__ u32 sh[256];
  const bool small = R <= 256;
  if (small) { sh[threadIdx.x] = 0; __syncthreads(); }
  for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < N; r += (size_t)gridDim.x * blockDim.x) {
    u64 v = src[r] >> d.shift;
    if (d.shift == 0 && v >= R) { *err = 1; v = R - 1; }   // reference asserts every target value < range_max
    v &= (R - 1);
    if (small) atomicAdd(&sh[v], 1u); else atomicAdd(&c[v], 1u);
  }
  if (small) { __syncthreads(); if (threadIdx.x < R && sh[threadIdx.x]) atomicAdd(&c[threadIdx.x], sh[threadIdx.x]); }
}
// One warp per lookup: counting-sort output and the greedy permuted table of `permuted_cols`
// (lookup.rs:60-111) in histogram form: walking the table values v in order, a value absent from the inputs
// is pushed on the unused stack; the first input copy of v takes the table's v, each further copy pops the
// most recent unused value or, if none, is deferred; deferred positions finally take the leftover unused
// values bottom-first.  The table's last value R-1 occurs N-R+1 times (range_check.rs:33-35).
__global__ void __launch_bounds__(32) k_lookup_walk(const u32* __restrict__ cnt_all, u32 R, size_t N, u64* __restrict__ cols, const LookupDesc* __restrict__ descs,
                                                    u32* __restrict__ stack_all, u32* __restrict__ defer_all) {
  const int lane = threadIdx.x;
  const LookupDesc d = descs[blockIdx.x];
  const u32* cnt = cnt_all + (size_t)blockIdx.x * R;
  u32* stack = stack_all + (size_t)blockIdx.x * R;
  u32* defer = defer_all + (size_t)blockIdx.x * N;
  u64* sorted = cols + (size_t)d.sorted_col * N;
  u64* perm = cols + (size_t)d.perm_col * N;
  size_t off = 0, ndef = 0; u32 top = 0;
  for (u32 v0 = 0; v0 < R; v0 += 32) {
    const u32 cl = cnt[v0 + lane];
    for (int k = 0; k < 32; k++) {
      const u32 v = v0 + k;
      if (v == R - 1) break;
      const u32 c = __shfl_sync(0xffffffffu, cl, k);
      if (c == 0) {
        if (lane == 0) stack[top] = v;
        top++;
      } else {
        for (u32 t = lane; t < c; t += 32) sorted[off + t] = v;
        if (lane == 0) perm[off] = v;
        const u32 dd = c - 1, npop = dd < top ? dd : top;
        for (u32 t = lane; t < npop; t += 32) perm[off + 1 + t] = stack[top - 1 - t];
        top -= npop;
        const u32 nd = dd - npop;
        for (u32 t = lane; t < nd; t += 32) defer[ndef + t] = (u32)(off + 1 + npop + t);
        ndef += nd; off += c;
      }
      __syncwarp();
    }
  }
  {  // v = R-1: table holds tcount copies
    const size_t c = cnt[R - 1], tcount = N - R + 1, m = c < tcount ? c : tcount;
    for (size_t t = lane; t < c; t += 32) sorted[off + t] = R - 1;
    for (size_t t = lane; t < m; t += 32) perm[off + t] = R - 1;
    if (c > tcount) { for (size_t t = lane; t < c - tcount; t += 32) defer[ndef + t] = (u32)(off + tcount + t); ndef += c - tcount; }
  }
  __syncwarp();
  for (size_t k = lane; k < ndef; k += 32) perm[defer[k]] = k < top ? stack[k] : R - 1;
}

// Same algorithm, 32 table values per step: the warp loads 32 counts at once, replays their pushes and pops on
// BIT MASKS in registers (only lanes with surplus copies are visited; a pop takes the highest still-available in-chunk
// push, then entries of the global stack below the chunk), and then every lane writes its own run.  The sequential
// kernel above spends one dependent global-memory round trip per table value; this one spends one per 32 values.
// Long runs (> 32 copies of one value: default rows, skewed limbs) are written by the whole warp.  Bit-identical output.
__global__ void __launch_bounds__(32) k_lookup_walk_chunked(const u32* __restrict__ cnt_all, u32 R, size_t N, u64* __restrict__ cols, const LookupDesc* __restrict__ descs,
                                                            u32* __restrict__ stack_all, uint2* __restrict__ defer_all) {
  const int lane = threadIdx.x;
  const LookupDesc d = descs[blockIdx.x];
  const u32* cnt = cnt_all + (size_t)blockIdx.x * R;
  u32* stack = stack_all + (size_t)blockIdx.x * R;
  uint2* defr = defer_all + (size_t)blockIdx.x * R;
  u64* sorted = cols + (size_t)d.sorted_col * N;
  u64* perm = cols + (size_t)d.perm_col * N;
  u32 off = 0, top = 0, ndef = 0, last_off = 0;   // warp-uniform
  u32 cnext = cnt[lane];
  for (u32 v0 = 0; v0 < R; v0 += 32) {
    const u32 v = v0 + lane;
    const bool last = (v == R - 1);
    const u32 c = cnext;
    if (v0 + 32 < R) cnext = cnt[v + 32];          // prefetch the next chunk's counts
    const u32 cs = last ? 1u : c;                  // the table's last value neither pushes nor pops
    u32 incl = c;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, dd); if (lane >= dd) incl += t; }
    const u32 myoff = off + incl - c;
    const u32 total = __shfl_sync(0xffffffffu, incl, 31);
    const u32 zero_mask = __ballot_sync(0xffffffffu, cs == 0);
    u32 pm = __ballot_sync(0xffffffffu, cs >= 2);
    u32 consumed = 0, pre_top = top;
    u32 my_taken = 0, my_pre_start = 0, my_pre_cnt = 0, my_def = 0, my_def_slot = 0;
    while (pm) {
      const int i = __ffs(pm) - 1;
      pm &= pm - 1;
      u32 dsur = __shfl_sync(0xffffffffu, cs, i) - 1;
      u32 avail = zero_mask & ((1u << i) - 1) & ~consumed, taken = 0;
      while (dsur > 0 && avail) { const int b = 31 - __clz(avail); avail ^= 1u << b; taken |= 1u << b; dsur--; }
      consumed |= taken;
      const u32 pre_take = dsur < pre_top ? dsur : pre_top;
      if (lane == i) { my_taken = taken; my_pre_start = pre_top; my_pre_cnt = pre_take; my_def = dsur - pre_take; my_def_slot = ndef; }
      pre_top -= pre_take;
      if (dsur > pre_take) ndef++;
    }
    if (!last && c > 0) {
      if (c <= 32) for (u32 t = 0; t < c; t++) sorted[myoff + t] = v;
      perm[myoff] = v;
      u32 p = myoff + 1;
      for (u32 m = my_taken; m;) { const int b = 31 - __clz(m); m ^= 1u << b; perm[p++] = v0 + b; }
      if (my_pre_cnt <= 32) for (u32 t = 0; t < my_pre_cnt; t++) perm[p + t] = stack[my_pre_start - 1 - t];
      if (my_def) defr[my_def_slot] = make_uint2(p + my_pre_cnt, my_def);
    }
    u32 big = __ballot_sync(0xffffffffu, !last && (c > 32 || my_pre_cnt > 32));
    while (big) {
      const int i = __ffs(big) - 1;
      big &= big - 1;
      const u32 ci = __shfl_sync(0xffffffffu, c, i), oi = __shfl_sync(0xffffffffu, myoff, i);
      if (ci > 32) for (u32 t = lane; t < ci; t += 32) sorted[oi + t] = v0 + i;
      const u32 pc = __shfl_sync(0xffffffffu, my_pre_cnt, i), ps = __shfl_sync(0xffffffffu, my_pre_start, i);
      const u32 pp = oi + 1 + __popc(__shfl_sync(0xffffffffu, my_taken, i));
      if (pc > 32) for (u32 t = lane; t < pc; t += 32) perm[pp + t] = stack[ps - 1 - t];
    }
    __syncwarp();   // pops above read stack slots that the surviving pushes below may overwrite
    const u32 surv = zero_mask & ~consumed;
    if ((surv >> lane) & 1) stack[pre_top + __popc(surv & ((1u << lane) - 1))] = v;
    top = pre_top + __popc(surv);
    if (v0 + 32 >= R) last_off = __shfl_sync(0xffffffffu, myoff, 31);
    off += total;
    __syncwarp();
  }
  {  // v = R-1: the table holds tcount copies; surplus inputs are deferred, never popped (lookup.rs:79,103-104)
    const size_t c = cnt[R - 1], tcount = N - R + 1, m = c < tcount ? c : tcount;
    for (size_t t = lane; t < c; t += 32) sorted[last_off + t] = R - 1;
    for (size_t t = lane; t < m; t += 32) perm[last_off + t] = R - 1;
    if (c > tcount) { if (lane == 0) defr[ndef] = make_uint2((u32)(last_off + tcount), (u32)(c - tcount)); ndef++; }
  }
  __syncwarp();
  // deferred positions (ascending) take the leftover unused values bottom-first, then the table's last value
  u32 k = 0;
  for (u32 j = 0; j < ndef; j++) {
    const uint2 r = defr[j];
    for (u32 t = lane; t < r.y; t += 32) perm[r.x + t] = (k + t < top) ? stack[k + t] : R - 1;
    k += r.y;
  }
}

// Same result without the walk: one block per lookup, u16 table (R = 65 536).  With delta(v) = +1 for an absent value (a push),
// -(c - 1) for a value with c >= 2 copies (its pops) and S the prefix sums of delta, M = min(0, running minimum of S):
//   stack height after v  H(v) = S(v) - M(v),   positions deferred so far  D(v) = -M(v),
//   a pushed value u is popped by the first w > u with H(w) < H(u), as pop number H(w-1) - H(u) of that group; a push nobody
//   pops is entry H(u) - 1 of the leftover stack; the group at v pops min(c - 1, H(v-1)) values and defers the rest, whose
//   k-th overall takes leftover[k] (or R-1 past the leftover's end)  -- the pairing lookup.rs:60-111 produces with its
//   explicit stack.  Passes 1-2 are two block scans with 64 consecutive values per thread (heights into shared memory under a
//   4-ary minimum tree for the "next smaller" search; offsets and deferred counts into the scratch); passes 3-4 place every
//   table value independently, a warp taking 32 consecutive values per step so that its stores fall into one short run of
//   rows.  65 536 dependent steps of one warp become 64 per thread.
constexpr int LWP_THREADS = 1024, LWP_R = 1 << 16, LWP_VPT = LWP_R / LWP_THREADS, LWP_TREE = 21844, LWP_HUGE = 32, LWP_HUGE_MIN = 2048;
struct LwpScan { int sum, mn; u32 off; };
constexpr size_t LWP_SMEM = (size_t)LWP_R * 2 + (size_t)LWP_TREE * 2 + 32 * sizeof(LwpScan) + (2 + LWP_HUGE) * 4;
// level l >= 1 of the minimum tree (65536 >> 2l nodes, node i = min of nodes 4i .. 4i+3 one level down) starts at LWP_OFF[l]; level 0 = heights
__constant__ int LWP_OFF[8] = {0, 0, 16384, 20480, 21504, 21760, 21824, 21840};
// heights are stored at a swizzled index: bank-conflict free both for "64 consecutive values per lane" and "32 consecutive values per warp"
__device__ __forceinline__ u32 lwp_swz(u32 v) { return v ^ ((v >> 5) & 0x3eu); }
// First w > u with H(w) < x, or -1: up the tree past the siblings to the right, down into the first subtree whose minimum is below x.
// One node per iteration and a warp-wide vote as the loop condition, so the 32 searches of a warp advance together (with early
// returns the lanes leave the loops at different times and each finishes its search alone: measured 1.95 active lanes per
// instruction).  Every lane of the warp calls it; `active` = this lane searches.
__device__ __forceinline__ int lwp_next_smaller(const u16* H, int u, u32 x, bool active) {
  int pos = u + 1, l = 0, res = -1;
  int mode = active ? 1 : 0;   // 0 finished, 1 going up, 2 going down
  while (__any_sync(0xffffffffu, mode != 0)) {
    if (mode == 1) {   // a node that is the first child of its parent: the parent stands for it and its three siblings
      const int k = min((__ffs(pos) - 1) >> 1, 7 - l);
      pos >>= 2 * k; l += k;
      if ((pos & 3) == 0) mode = 0;   // level 7, node 4: past the end -- nobody pops u
    }
    if (mode != 0) {
      const u32 t = H[(l == 0 ? 0 : LWP_R + LWP_OFF[l]) + (pos ^ (l == 0 ? (pos >> 5) & 0x3e : 0))];   // the tree follows the heights in shared memory
      if (t >= x) pos++;
      else if (l == 0) { res = pos; mode = 0; }
      else { l--; pos <<= 2; mode = 2; }
    }
  }
  return res;
}
__global__ void __launch_bounds__(LWP_THREADS, 1) k_lookup_walk_parallel(const u32* __restrict__ cnt_all, size_t N, u64* __restrict__ cols, const LookupDesc* __restrict__ descs,
                                                                         u32* __restrict__ stack_all, u32* __restrict__ scratch_all, u32 huge_cap /* <= LWP_HUGE */) {
  extern __shared__ __align__(16) unsigned char lwp_smem[];
  u16* H = reinterpret_cast<u16*>(lwp_smem);
  u16* tree = H + LWP_R;
  LwpScan* wscan = reinterpret_cast<LwpScan*>(tree + LWP_TREE);
  u32* misc = reinterpret_cast<u32*>(wscan + 32);            // [0] positions deferred before the table's last value, [1] length of huge[]
  u32* huge = misc + 2;                                       // values with more than LWP_HUGE_MIN copies: written by the whole block
  constexpr u32 R = LWP_R, FULL = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const LookupDesc d = descs[blockIdx.x];
  const u32* cnt = cnt_all + (size_t)blockIdx.x * R;
  u32* leftover = stack_all + (size_t)blockIdx.x * R;
  u32* off0g = scratch_all + (size_t)blockIdx.x * 2 * R;     // exclusive prefix sums of the counts
  u32* dg = off0g + R;                                        // D(v)
  u64* sorted = cols + (size_t)d.sorted_col * N;
  u64* perm = cols + (size_t)d.perm_col * N;
  const u32 v0 = (u32)tid * LWP_VPT;
  const uint4* c4 = reinterpret_cast<const uint4*>(cnt + v0);
  if (tid == 0) misc[1] = 0;
  // ---- pass 1: per-thread totals, block scan ----
  LwpScan mine; mine.sum = 0; mine.mn = 0x3fffffff; mine.off = 0;
#pragma unroll 1
  for (int q = 0; q < LWP_VPT / 4; q++) {
    const uint4 x = c4[q];
    const u32 cc[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 c = cc[k], v = v0 + q * 4 + k;
      if (v != R - 1) mine.sum += c == 0 ? 1 : -(int)(c - 1);
      mine.mn = min(mine.mn, mine.sum);
      mine.off += c;
    }
  }
  LwpScan inc = mine;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    const int ps = __shfl_up_sync(FULL, inc.sum, dd), pm = __shfl_up_sync(FULL, inc.mn, dd);
    const u32 po = __shfl_up_sync(FULL, inc.off, dd);
    if (lane >= dd) { inc.mn = min(pm, ps + inc.mn); inc.sum += ps; inc.off += po; }
  }
  if (lane == 31) wscan[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    LwpScan w = wscan[lane];
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      const int ps = __shfl_up_sync(FULL, w.sum, dd), pm = __shfl_up_sync(FULL, w.mn, dd);
      const u32 po = __shfl_up_sync(FULL, w.off, dd);
      if (lane >= dd) { w.mn = min(pm, ps + w.mn); w.sum += ps; w.off += po; }
    }
    wscan[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  LwpScan base;        // everything before this thread's first value
  {
    LwpScan ex;
    ex.sum = __shfl_up_sync(FULL, inc.sum, 1); ex.mn = __shfl_up_sync(FULL, inc.mn, 1); ex.off = __shfl_up_sync(FULL, inc.off, 1);
    if (lane == 0) { ex.sum = 0; ex.mn = 0x3fffffff; ex.off = 0; }
    if (warp == 0) base = ex;
    else { const LwpScan wb = wscan[warp - 1]; base.sum = wb.sum + ex.sum; base.mn = min(wb.mn, wb.sum + ex.mn); base.off = wb.off + ex.off; }
  }
  // ---- pass 2: heights, the three lowest tree levels, offsets and deferred counts ----
  {
    int S = base.sum, M = min(0, base.mn); u32 off = base.off, m1 = 0xffffu, m2 = 0xffffu, m3 = 0xffffu;
#pragma unroll 1
    for (int q = 0; q < LWP_VPT / 4; q++) {
      const uint4 x = c4[q];
      const u32 cc[4] = {x.x, x.y, x.z, x.w};
      uint4 o, dq;
      u32* op = reinterpret_cast<u32*>(&o);
      u32* dp = reinterpret_cast<u32*>(&dq);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const u32 c = cc[k], v = v0 + q * 4 + k;
        if (v != R - 1) S += c == 0 ? 1 : -(int)(c - 1);
        M = min(M, S);
        const u32 h = (u32)(S - M);
        H[lwp_swz(v)] = (u16)h;
        m1 = min(m1, h);
        op[k] = off; off += c;
        dp[k] = (u32)(-M);
      }
      reinterpret_cast<uint4*>(off0g + v0)[q] = o;
      reinterpret_cast<uint4*>(dg + v0)[q] = dq;
      tree[(v0 >> 2) + q] = (u16)m1;
      m2 = min(m2, m1); m1 = 0xffffu;
      if ((q & 3) == 3) { tree[16384 + (v0 >> 4) + (q >> 2)] = (u16)m2; m3 = min(m3, m2); m2 = 0xffffu; }
    }
    tree[20480 + tid] = (u16)m3;
    if (tid == LWP_THREADS - 1) misc[0] = (u32)(-M);
  }
  __syncthreads();
#pragma unroll 1
  for (int l = 4; l <= 7; l++) {   // 256, 64, 16, 4 nodes
    const int n = LWP_R >> (2 * l);
    if (tid < n) { const u16* c = tree + LWP_OFF[l - 1] + 4 * tid; tree[LWP_OFF[l] + tid] = min(min(c[0], c[1]), min(c[2], c[3])); }
    __syncthreads();
  }
  const u32 top = H[lwp_swz(R - 1)];
  const u32 wbase = (u32)warp * (R / 32);
  // ---- pass 3: every push finds its place; first copies and sorted runs ----
  bool any_deferred = false;
#pragma unroll 1
  for (int j = 0; j < (int)(R / 32 / 32); j++) {
    const u32 v = wbase + j * 32 + lane;
    const u32 c = cnt[v], off = off0g[v];
    const bool live = v != R - 1, push = live && c == 0;
    const u32 hx = H[lwp_swz(v)];
    const int w = lwp_next_smaller(H, (int)v, hx, push);
    if (push) {
      if (w < 0) leftover[hx - 1] = v;
      else perm[off0g[w] + 1 + ((u32)H[lwp_swz((u32)w - 1)] - hx)] = v;
    } else if (live) {
      perm[off] = v;
      if (c <= 32) for (u32 t = 0; t < c; t++) sorted[off + t] = v;
      const u32 hp = v ? (u32)H[lwp_swz(v - 1)] : 0u;
      if (c - 1 > hp) any_deferred = true;
      if (c > LWP_HUGE_MIN) { const u32 slot = atomicAdd(&misc[1], 1u); if (slot < huge_cap) huge[slot] = v; }
    }
    u32 big = __ballot_sync(FULL, live && c > 32 && c <= LWP_HUGE_MIN);
    while (big) {
      const int i = __ffs(big) - 1;
      big &= big - 1;
      const u32 ci = __shfl_sync(FULL, c, i), oi = __shfl_sync(FULL, off, i), vi = wbase + j * 32 + i;
      for (u32 t = lane; t < ci; t += 32) sorted[oi + t] = vi;
    }
  }
  any_deferred = __syncthreads_or(any_deferred);   // leftover[] and huge[] complete
  const u32 nhuge = misc[1];                        // > huge_cap: the list overflowed, the values beyond it are written by their warps
  // ---- pass 4: deferred positions take the leftover stack bottom-first, then the table's last value ----
#pragma unroll 1
  for (int j = 0; j < (int)(R / 32 / 32); j++) {
    if (!any_deferred && nhuge <= huge_cap) break;
    const u32 v = wbase + j * 32 + lane;
    const u32 c = cnt[v];
    const bool live = v != R - 1;
    const u32 hp = v ? (u32)H[lwp_swz(v - 1)] : 0u;
    bool listed = false;
    if (live && c > LWP_HUGE_MIN) for (u32 e = 0; e < min(nhuge, huge_cap); e++) listed |= huge[e] == v;
    const u32 nd = (live && !listed && c >= 2 && c - 1 > hp) ? c - 1 - hp : 0u;
    const bool fill = live && !listed && c > LWP_HUGE_MIN;   // a long run that did not fit the list
    if (!__ballot_sync(FULL, nd > 0 || fill)) continue;
    const u32 off = off0g[v], pos = off + 1 + hp, D = v ? dg[v - 1] : 0u;
    if (nd <= 32) for (u32 t = 0; t < nd; t++) { const u32 k = D + t; perm[pos + t] = k < top ? leftover[k] : R - 1; }
    u32 big = __ballot_sync(FULL, nd > 32 || fill);
    while (big) {
      const int i = __ffs(big) - 1;
      big &= big - 1;
      const u32 ni = __shfl_sync(FULL, nd, i), pi = __shfl_sync(FULL, pos, i), Di = __shfl_sync(FULL, D, i);
      if (ni > 32) for (u32 t = lane; t < ni; t += 32) { const u32 k = Di + t; perm[pi + t] = k < top ? leftover[k] : R - 1; }
      if (__shfl_sync(FULL, (int)fill, i)) {
        const u32 ci = __shfl_sync(FULL, c, i), oi = __shfl_sync(FULL, off, i), vi = wbase + j * 32 + i;
        for (u32 t = lane; t < ci; t += 32) sorted[oi + t] = vi;
      }
    }
  }
  for (u32 e = 0; e < min(nhuge, huge_cap); e++) {   // the longest runs: sorted copies and deferred positions by the whole block
    const u32 v = huge[e], c = cnt[v], off = off0g[v], hp = v ? (u32)H[lwp_swz(v - 1)] : 0u, D = v ? dg[v - 1] : 0u;
    const u32 npop = min(c - 1, hp), nd = c - 1 - npop;
    for (u32 t = tid; t < c; t += LWP_THREADS) sorted[off + t] = v;
    for (u32 t = tid; t < nd; t += LWP_THREADS) { const u32 k = D + t; perm[off + 1 + npop + t] = k < top ? leftover[k] : R - 1; }
  }
  {  // the table's last value R-1 occurs tcount times: its surplus inputs are deferred, never popped
    const size_t c = cnt[R - 1], tcount = N - R + 1, m = c < tcount ? c : tcount, lo = off0g[R - 1];
    const u32 D = misc[0];
    for (size_t t = tid; t < c; t += LWP_THREADS) sorted[lo + t] = R - 1;
    for (size_t t = tid; t < m; t += LWP_THREADS) perm[lo + t] = R - 1;
    if (c > tcount) for (size_t t = tid; t < c - tcount; t += LWP_THREADS) { const size_t k = D + t; perm[lo + tcount + t] = k < top ? leftover[k] : R - 1; }
  }
}

static void run_lookups(sbn_ctx* ctx, u64* d_cols, size_t N, u32 R, const std::vector<LookupDesc>& descs) {
  SBN_REQUIRE(N >= R, "range-check table does not fit the trace (reference asserts rows >= range_max)");
  SBN_REQUIRE(N < (size_t(1) << 32), "trace too long");
  SBN_REQUIRE(R % 32 == 0, "lookup table size must be a multiple of 32");
  DevBuf<int> err(ctx, 1);
  CUDA_CHECK(cudaMemsetAsync(err, 0, 4, ctx->stream));
  // lookups are processed in groups so the scratch (counts, unused stack, deferred ranges per lookup) stays bounded
  size_t per = (size_t)R * (4 + 4 + 8);
  size_t group = std::max<size_t>(1, (size_t(512) << 20) / per);
  group = std::min(group, descs.size());
  DevBuf<LookupDesc> d_desc(ctx, descs.size());
  ctx->upload(d_desc, descs.data(), descs.size() * sizeof(LookupDesc));
  // SBN_LOOKUP_SEQUENTIAL=1 selects the one-value-per-step kernel (k_lookup_walk), kept as the in-library cross-check of the
  // chunked kernel (tests/test_gpu_parity.py compares the two on skewed and uniform columns).
  // SBN_LOOKUP_WALK=chunked selects the one-warp-per-lookup kernel for the u16 table too (it always serves the other table sizes).
  const bool sequential = getenv("SBN_LOOKUP_SEQUENTIAL") != nullptr;
  const char* walk_env = getenv("SBN_LOOKUP_WALK");
  const bool parallel = !sequential && R == (u32)LWP_R && N < (size_t(1) << 31) && !(walk_env && !strcmp(walk_env, "chunked"));
  // SBN_LOOKUP_HUGE_LIST=n (test switch): capacity of the kernel's list of very long runs, 0 exercises its overflow path
  const char* huge_env = getenv("SBN_LOOKUP_HUGE_LIST");
  const u32 huge_cap = huge_env ? (u32)std::min(std::max(atoi(huge_env), 0), (int)LWP_HUGE) : (u32)LWP_HUGE;
  if (parallel) CUDA_CHECK(cudaFuncSetAttribute(k_lookup_walk_parallel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LWP_SMEM));
  if (sequential) group = std::min(group, std::max<size_t>(1, (size_t(512) << 20) / ((N + 2 * (size_t)R) * 4)));
  DevBuf<u32> cnt(ctx, group * R), stack(ctx, group * R);
  DevBuf<uint2> defer(ctx, sequential ? 1 : group * R);
  DevBuf<u32> defer_seq(ctx, sequential ? group * N : 1);
  for (size_t g0 = 0; g0 < descs.size(); g0 += group) {
    size_t ng = std::min(group, descs.size() - g0);
    CUDA_CHECK(cudaMemsetAsync(cnt, 0, ng * R * 4, ctx->stream));
    unsigned bx = (unsigned)std::min<size_t>((N + 255) / 256, 64);
    { KScope ks(ctx, "lookup_hist");
    k_lookup_hist<<<dim3(bx, (unsigned)ng), 256, 0, ctx->stream>>>(d_cols, N, R, d_desc + g0, cnt, err);
    LAUNCH_CHECK(ctx); }
    KScope ks2(ctx, "lookup_walk");
    if (sequential) k_lookup_walk<<<(unsigned)ng, 32, 0, ctx->stream>>>(cnt, R, N, d_cols, d_desc + g0, stack, defer_seq);
    else if (parallel) k_lookup_walk_parallel<<<(unsigned)ng, LWP_THREADS, LWP_SMEM, ctx->stream>>>(cnt, N, d_cols, d_desc + g0, stack, reinterpret_cast<u32*>((uint2*)defer), huge_cap);
    else k_lookup_walk_chunked<<<(unsigned)ng, 32, 0, ctx->stream>>>(cnt, R, N, d_cols, d_desc + g0, stack, defer);
    LAUNCH_CHECK(ctx);
  }
  int h_err = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  SBN_REQUIRE(!h_err, "range-checked column holds a value >= 2^16");
}

// reference src/utils/range_check.rs:20-47: table | (sorted_i, permuted_table_i) per target column
void generate_u16_range_check_cols(sbn_ctx* ctx, u64* d_cols, size_t N, int t0, int ntargets, int start_lookups) {
  k_table_col<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_cols + (size_t)start_lookups * N, N, 1u << 16);
  LAUNCH_CHECK(ctx);
  std::vector<LookupDesc> descs;
  for (int i = 0; i < ntargets; i++) descs.push_back({t0 + i, 0, start_lookups + 1 + 2 * i, start_lookups + 2 + 2 * i});
  run_lookups(ctx, d_cols, N, 1u << 16, descs);
}
// reference src/utils/range_check.rs:116-160: table | (lo, sorted_lo, perm_lo, hi, sorted_hi, perm_hi) per target
void generate_split_u16_range_check_cols(sbn_ctx* ctx, u64* d_cols, size_t N, int t0, int ntargets, int main_col) {
  k_table_col<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_cols + (size_t)main_col * N, N, 1u << 8);
  LAUNCH_CHECK(ctx);
  std::vector<LookupDesc> descs;
  for (int i0 = 0; i0 < ntargets; i0 += 32768) {   // grid.y limit
    const int n = std::min(ntargets - i0, 32768);
    k_split_cols<<<dim3((unsigned)((N + 255) / 256), (unsigned)n), 256, 0, ctx->stream>>>(d_cols, N, t0 + i0, main_col + 6 * i0);
    LAUNCH_CHECK(ctx);
  }
  for (int i = 0; i < ntargets; i++) {
    int b = main_col + 1 + 6 * i;
    descs.push_back({b, 0, b + 1, b + 2});
    descs.push_back({b + 3, 0, b + 4, b + 5});
  }
  run_lookups(ctx, d_cols, N, 1u << 8, descs);
}

// ---------------- ModularStark ----------------
__global__ void __launch_bounds__(128) k_modular_rows(const u64* __restrict__ ios, u64* __restrict__ cols, size_t N, int* __restrict__ err) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  { u32 c[8]; u64x4_to_words(ios + r * 8, c); if (fq_geq_p(c)) *err = 1; u64x4_to_words(ios + r * 8 + 4, c); if (fq_geq_p(c)) *err = 1; }
  ColWriter w{cols + r, N};
  modular_stark_row(ios + r * 8, w);
}

// ---------------- G1ExpStark ----------------
struct G1Io { u64 x_x[4], x_y[4], off_x[4], off_y[4]; u32 exp[8]; u64 out_x[4], out_y[4]; };

// chain points: A[k] = 2^k * x, B[k] = offset + sum_{j<k, bit_j} A[j], k = 0..256 (Jacobian, Montgomery)
// Block = 64 threads for 32 instances: warp 0 runs the doubling chain A, warp 1 the addition chain B (which needs A[k] before it
// is doubled: handed over through a double-buffered shared slot, one barrier per bit).  Both chains are sequential in k; on
// separate warps their steps overlap instead of alternating in one thread (6.3 -> 3.9 ms for 128 instances).
__global__ void __launch_bounds__(64) k_g1_chain(const G1Io* __restrict__ ios, size_t num_io, G1Jac* __restrict__ jac /* [io][2][257] */, int* __restrict__ err) {
  __shared__ G1Jac hand[2][32];
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;   // 0: doublings, 1: additions
  const size_t i = blockIdx.x * (size_t)32 + lane;
  const bool live = i < num_io;
  const G1Io& io = ios[live ? i : 0];
  u32 w[8];
  G1Jac P;   // A on warp 0, B on warp 1
  u64x4_to_words(role ? io.off_x : io.x_x, w); if (live && fq_geq_p(w)) *err = 2;   // coordinates must be canonical residues
  P.x = fq_from_words(w);
  u64x4_to_words(role ? io.off_y : io.x_y, w); if (live && fq_geq_p(w)) *err = 2;
  P.y = fq_from_words(w); P.z = fq_one();
  G1Jac* out = jac + (live ? i : 0) * 2 * 257 + (role ? 257 : 0);
  if (live) out[0] = P;
  for (int k = 0; k < 256; k++) {
    if (role == 0) hand[k & 1][lane] = P;
    __syncthreads();
    if (role == 0) P = g1_jac_dbl(P);
    else if ((io.exp[k >> 5] >> (k & 31)) & 1) P = g1_jac_add(hand[k & 1][lane], P);
    if (live) out[k + 1] = P;
  }
}
// The same chains on more of the machine (round 2).  A[k] = 2^k x is inherently sequential (one thread per instance, 256
// doublings); B[k] = offset + sum_{j<k, bit_j} A[j] is a PREFIX SUM under point addition, which is associative: one block per
// instance, one thread per bit, a Kogge-Stone scan of the terms bit_j ? A[j] : identity through shared memory (8 rounds of one
// addition each instead of 256 dependent ones), then one addition of the offset.  The points differ from the sequential chain's
// only in their Jacobian representation; the affine coordinates the trace is built from are the same field elements.  The
// identity is Z = 0; a sum of two finite points that comes out with Z = 0 (equal or opposite points) raises the error flag the
// sequential chain raises through k_g1_affine.
__global__ void __launch_bounds__(32) k_g1_dbl_chain(const G1Io* __restrict__ ios, size_t num_io, G1Jac* __restrict__ jac /* [io][2][257] */, int* __restrict__ err) {
  const size_t i = blockIdx.x * (size_t)32 + threadIdx.x;
  if (i >= num_io) return;
  const G1Io& io = ios[i];
  u32 w[8];
  G1Jac P;
  u64x4_to_words(io.x_x, w); if (fq_geq_p(w)) *err = 2;
  P.x = fq_from_words(w);
  u64x4_to_words(io.x_y, w); if (fq_geq_p(w)) *err = 2;
  P.y = fq_from_words(w); P.z = fq_one();
  G1Jac* out = jac + i * 2 * 257;
  out[0] = P;
  for (int k = 0; k < 256; k++) { P = g1_jac_dbl(P); out[k + 1] = P; }
}
HD G1Jac g1_jac_add_id(const G1Jac& p, const G1Jac& q, int* err) {   // identity-aware
  if (fq_is_zero(p.z)) return q;
  if (fq_is_zero(q.z)) return p;
  G1Jac r = g1_jac_add(p, q);
  if (fq_is_zero(r.z)) *err = 1;
  return r;
}
__global__ void __launch_bounds__(256) k_g1_sum_scan(const G1Io* __restrict__ ios, G1Jac* __restrict__ jac /* [io][2][257] */, int* __restrict__ err) {
  __shared__ G1Jac buf[256];
  const size_t i = blockIdx.x;
  const int j = threadIdx.x;
  const G1Io& io = ios[i];
  G1Jac* A = jac + i * 2 * 257;
  G1Jac* B = A + 257;
  G1Jac v;
  if ((io.exp[j >> 5] >> (j & 31)) & 1) v = A[j];
  else { v.x = fq_one(); v.y = fq_one(); v.z = fq_zero(); }
#pragma unroll 1
  for (int d = 1; d < 256; d <<= 1) {
    buf[j] = v;
    __syncthreads();
    if (j >= d) v = g1_jac_add_id(buf[j - d], v, err);
    __syncthreads();
  }
  u32 w[8];
  G1Jac off;
  u64x4_to_words(io.off_x, w); if (fq_geq_p(w)) *err = 2;
  off.x = fq_from_words(w);
  u64x4_to_words(io.off_y, w); if (fq_geq_p(w)) *err = 2;
  off.y = fq_from_words(w); off.z = fq_one();
  if (j == 0) B[0] = off;
  B[j + 1] = g1_jac_add_id(off, v, err);
}
__global__ void __launch_bounds__(128) k_g1_affine(const G1Jac* __restrict__ jac, size_t npoints, u32* __restrict__ aff /* [point][16] */, int* __restrict__ err) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npoints) return;
  G1Jac p = jac[i];
  if (fq_is_zero(p.z)) { *err = 1; return; }
  Fq zi = fq_inv(p.z), zi2 = fq_sqr(zi);
  Fq x = fq_mul(p.x, zi2), y = fq_mul(p.y, fq_mul(zi2, zi));
  u32 w[8];
  fq_to_words(x, w);
#pragma unroll
  for (int k = 0; k < 8; k++) aff[i * 16 + k] = w[k];
  fq_to_words(y, w);
#pragma unroll
  for (int k = 0; k < 8; k++) aff[i * 16 + 8 + k] = w[k];
}
// Rows of the exponentiation AIRs alternate between the squaring / doubling step (even rows) and the conditional multiplication /
// addition step (odd rows).  A warp takes 32 rows of one parity (warp 2g: rows 64g, 64g+2, ...; warp 2g+1: the odd ones) so that
// its lanes run the same witness code: with 32 consecutive rows per warp the two operations were serialised (measured 12.1
// active lanes per instruction in k_g1_rows).  The two warps of a pair sit in the same block and fill each other's sectors.
__device__ __forceinline__ size_t exp_row_of_thread() {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  return ((t >> 6) << 6) + 2 * (t & 31) + ((t >> 5) & 1);
}
// main columns: a(32) b(32) G1Output(320) flags(14)   (reference src/curves/g1/exp.rs:165-230, flags.rs:46-134)
__global__ void __launch_bounds__(128) k_g1_rows(const G1Io* __restrict__ ios, const u32* __restrict__ aff, u64* __restrict__ cols, size_t N, int* __restrict__ err) {
  const size_t r = exp_row_of_thread();
  if (r >= N) return;
  const size_t inst = r >> 9; const int rr = (int)(r & 511), k = rr >> 1;
  u32 e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = ios[inst].exp[i];
  u64 fl[14];
  flags_row(e, rr, fl);
  const u32* pa = aff + ((inst * 2) * 257 + k) * 16;
  const u32* pb = aff + ((inst * 2 + 1) * 257 + k + (rr & 1)) * 16;
  u32 ax[8], ay[8], bx[8], by[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { ax[i] = pa[i]; ay[i] = pa[8 + i]; bx[i] = pb[i]; by[i] = pb[8 + i]; }
  const int op = fl[2] ? G1_OP_DOUBLE : (fl[4] ? G1_OP_ADD : G1_OP_NONE);   // a = is_double, filtered_bit = is_add
  ColWriter w{cols + r, N};
  if (!g1_row(ax, ay, bx, by, op, w)) *err = 1;
#pragma unroll
  for (int i = 0; i < 14; i++) w(384 + i, fl[i]);
}
struct Inv64 { u64 v[64]; };
// periodic pulse witness (reference src/utils/pulse.rs:100-144, period 64, first pulse 62): counter, inverse witness
__global__ void k_periodic_pulse(u64* counter_col, u64* witness_col, size_t N, Inv64 inv) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  u32 c = (u32)((r + 1) & 63);
  counter_col[r] = c; witness_col[r] = inv.v[c];
}
__global__ void k_inverse_table(u64* t, size_t N) {  // t[d] = d^-1, d in [1, N)
  size_t d = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (d >= N) return;
  t[d] = d ? gl_inv((u64)d) : 0;
}
// io pulses (reference src/utils/pulse.rs:20-43): counter | (witness_i, pulse_i) for positions 512k, 512k+511
__global__ void __launch_bounds__(256) k_io_pulses(u64* __restrict__ cols /* at start_io_pulses */, size_t N, int rows_per_io, const u64* __restrict__ invtab) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  int i = blockIdx.y;  // position index
  if (i == 0) cols[r] = r;
  size_t pos = (size_t)(i >> 1) * rows_per_io + ((i & 1) ? rows_per_io - 1 : 0);
  u64 wv = 0;
  if (r > pos) wv = invtab[r - pos]; else if (r < pos) wv = gl_neg(invtab[pos - r]);
  cols[(size_t)(1 + 2 * i) * N + r] = wv;
  cols[(size_t)(2 + 2 * i) * N + r] = (r == pos) ? 1 : 0;
}

static void generate_g1(sbn_ctx* ctx, const AirDesc& air, const void* ios, bool on_device, u64* d_cols, u64* h_results) {
  const size_t n = air.num_io, N = air.num_rows;
  static_assert(sizeof(G1Io) == sizeof(sbn_g1_exp_io), "io layout");
  DevBuf<G1Io> d_ios_buf;
  const G1Io* d_ios = (const G1Io*)ios;
  if (!on_device) {
    d_ios_buf = DevBuf<G1Io>(ctx, n);
    CUDA_CHECK(cudaMemcpyAsync(d_ios_buf, ios, n * sizeof(G1Io), cudaMemcpyHostToDevice, ctx->stream));
    d_ios = d_ios_buf;
  }
  DevBuf<int> err(ctx, 1);
  CUDA_CHECK(cudaMemsetAsync(err, 0, 4, ctx->stream));
  const size_t npoints = n * 2 * 257;
  DevBuf<G1Jac> jac(ctx, npoints);
  DevBuf<u32> aff(ctx, npoints * 16);
  // SBN_CHAIN=sequential: the two-warp kernel with 256 dependent additions (kept as the cross-check of the scan)
  const bool seq_chain = getenv("SBN_CHAIN") && !strcmp(getenv("SBN_CHAIN"), "sequential");
  if (seq_chain) { KScope ks(ctx, "g1_chain"); k_g1_chain<<<(unsigned)((n + 31) / 32), 64, 0, ctx->stream>>>(d_ios, n, jac, err); LAUNCH_CHECK(ctx); }
  else {
    KScope ks(ctx, "g1_chain");
    k_g1_dbl_chain<<<(unsigned)((n + 31) / 32), 32, 0, ctx->stream>>>(d_ios, n, jac, err); LAUNCH_CHECK(ctx);
    k_g1_sum_scan<<<(unsigned)n, 256, 0, ctx->stream>>>(d_ios, jac, err); LAUNCH_CHECK(ctx);
  }
  { KScope ks(ctx, "g1_affine"); k_g1_affine<<<(unsigned)((npoints + 127) / 128), 128, 0, ctx->stream>>>(jac, npoints, aff, err); LAUNCH_CHECK(ctx); }
  { KScope ks(ctx, "g1_rows"); k_g1_rows<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(d_ios, aff, d_cols, N, err); LAUNCH_CHECK(ctx); }
  // results: b on the last row of each block = B[256]
  std::vector<u32> res(n * 16);
  // one strided copy: instance i's result sits (2 * 257 * 16) words after instance i-1's
  CUDA_CHECK(cudaMemcpy2DAsync(res.data(), 64, aff + (size_t)(257 + 256) * 16, (size_t)2 * 257 * 16 * 4, 64, n, cudaMemcpyDeviceToHost, ctx->stream));
  const int sf = 384, pp = sf + 14, iop = pp + 2, lookups = iop + 1 + 4 * (int)n;
  Inv64 inv;
  for (int c = 0; c < 64; c++) inv.v[c] = c == 63 ? 0 : gl_inv(gl_sub((u64)c, 63));
  k_periodic_pulse<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_cols + (size_t)pp * N, d_cols + (size_t)(pp + 1) * N, N, inv); LAUNCH_CHECK(ctx);
  DevBuf<u64> invtab(ctx, N);
  k_inverse_table<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(invtab, N); LAUNCH_CHECK(ctx);
  { KScope ksp(ctx, "io_pulses");
    k_io_pulses<<<dim3((unsigned)((N + 255) / 256), (unsigned)(2 * n)), 256, 0, ctx->stream>>>(d_cols + (size_t)iop * N, N, 512, invtab); LAUNCH_CHECK(ctx); }
  generate_u16_range_check_cols(ctx, d_cols, N, 0, 24 * 16 - 3, lookups);
  int h_err = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  SBN_REQUIRE(h_err != 2, "G1 coordinate is not a canonical Fq residue");
  SBN_REQUIRE(!h_err, "degenerate G1 input: the chain hit the point at infinity or two points with equal x (the reference panics here)");
  if (h_results) for (size_t i = 0; i < n; i++) memcpy(h_results + i * 8, res.data() + i * 16, 64);
}


// Common tail of every exponentiation AIR (e.g. reference src/curves/g1/exp.rs:299-312): periodic-pulse witness,
// io pulses, range-check lookups.
struct ExpTail { int sf, nflags; bool periodic; int rows_per_io; int t0, ntargets; bool split; };
static void generate_exp_tail(sbn_ctx* ctx, const AirDesc& air, u64* d_cols, const ExpTail& t) {
  const size_t n = air.num_io, N = air.num_rows;
  int col = t.sf + t.nflags;
  if (t.periodic) {
    Inv64 inv;
    for (int c = 0; c < 64; c++) inv.v[c] = c == 63 ? 0 : gl_inv(gl_sub((u64)c, 63));
    k_periodic_pulse<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(d_cols + (size_t)col * N, d_cols + (size_t)(col + 1) * N, N, inv); LAUNCH_CHECK(ctx);
    col += 2;
  }
  DevBuf<u64> invtab(ctx, N);
  k_inverse_table<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(invtab, N); LAUNCH_CHECK(ctx);
  { KScope ksp(ctx, "io_pulses");
    k_io_pulses<<<dim3((unsigned)((N + 255) / 256), (unsigned)(2 * n)), 256, 0, ctx->stream>>>(d_cols + (size_t)col * N, N, t.rows_per_io, invtab); LAUNCH_CHECK(ctx); }
  const int lookups = col + 1 + 4 * (int)n;
  if (t.split) generate_split_u16_range_check_cols(ctx, d_cols, N, t.t0, t.ntargets, lookups);
  else generate_u16_range_check_cols(ctx, d_cols, N, t.t0, t.ntargets, lookups);
}
static void check_chain_error(sbn_ctx* ctx, int* d_err, const char* what) {
  int h_err = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  SBN_REQUIRE(h_err != 2, "input coordinate is not a canonical Fq residue");
  SBN_REQUIRE(!h_err, what);
}
template <class T> static const T* stage_ios(sbn_ctx* ctx, const void* ios, bool on_device, size_t n, DevBuf<T>& buf) {
  if (on_device) return (const T*)ios;
  buf = DevBuf<T>(ctx, n);
  CUDA_CHECK(cudaMemcpyAsync(buf, ios, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return buf;
}

// ---------------- FqExpStark (reference src/fields/fq/exp.rs) ----------------
// chain values as canonical words: A[k] = x^(2^k), B[k] = offset * prod_{j<k, bit_j} A[j], k = 0..256
__global__ void __launch_bounds__(32) k_fq_chain(const sbn_fq_exp_io* __restrict__ ios, size_t num_io, u32* __restrict__ chain /* [io][2][257][8] */, int* __restrict__ err) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= num_io) return;
  const sbn_fq_exp_io& io = ios[i];
  u32 w[8];
  u64x4_to_words((const u64*)io.x, w); if (fq_geq_p(w)) *err = 2; Fq A = fq_from_words(w);
  u64x4_to_words((const u64*)io.offset, w); if (fq_geq_p(w)) *err = 2; Fq B = fq_from_words(w);
  u32* ca = chain + i * 2 * 257 * 8; u32* cb = ca + 257 * 8;
  for (int k = 0; k <= 256; k++) {
    fq_to_words(A, ca + k * 8); fq_to_words(B, cb + k * 8);
    if (k == 256) break;
    if ((io.exp_val[k >> 5] >> (k & 31)) & 1) B = fq_mul(B, A);
    A = fq_sqr(A);
  }
}
// main columns: a16 b16 FqOutput(112) flags14   (reference src/fields/fq/exp.rs:128-178)
__global__ void __launch_bounds__(128) k_fq_rows(const sbn_fq_exp_io* __restrict__ ios, const u32* __restrict__ chain, u64* __restrict__ cols, size_t N) {
  const size_t r = exp_row_of_thread();
  if (r >= N) return;
  const size_t inst = r >> 9; const int rr = (int)(r & 511), k = rr >> 1;
  u32 e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = ios[inst].exp_val[i];
  u64 fl[14];
  flags_row(e, rr, fl);
  const u32* pa = chain + ((inst * 2) * 257 + k) * 8;
  const u32* pb = chain + ((inst * 2 + 1) * 257 + k + (rr & 1)) * 8;
  u32 a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = pa[i]; b[i] = pb[i]; }
  const int op = fl[2] ? EXP_OP_SQUARE : (fl[4] ? EXP_OP_MUL : EXP_OP_NONE);
  ColWriter w{cols + r, N};
  fq_exp_row(a, b, op, w);
#pragma unroll
  for (int i = 0; i < 14; i++) w(144 + i, fl[i]);
}
static void generate_fq(sbn_ctx* ctx, const AirDesc& air, const void* ios, bool on_device, u64* d_cols, u64* h_results) {
  const size_t n = air.num_io, N = air.num_rows;
  DevBuf<sbn_fq_exp_io> buf; const sbn_fq_exp_io* d_ios = stage_ios(ctx, ios, on_device, n, buf);
  DevBuf<int> err(ctx, 1);
  CUDA_CHECK(cudaMemsetAsync(err, 0, 4, ctx->stream));
  DevBuf<u32> chain(ctx, n * 2 * 257 * 8);
  { KScope ks(ctx, "fq_chain"); k_fq_chain<<<(unsigned)((n + 31) / 32), 32, 0, ctx->stream>>>(d_ios, n, chain, err); LAUNCH_CHECK(ctx); }
  { KScope ks(ctx, "fq_rows"); k_fq_rows<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(d_ios, chain, d_cols, N); LAUNCH_CHECK(ctx); }
  std::vector<u32> res(n * 8);
  CUDA_CHECK(cudaMemcpy2DAsync(res.data(), 32, chain + (size_t)(257 + 256) * 8, (size_t)2 * 257 * 8 * 4, 32, n, cudaMemcpyDeviceToHost, ctx->stream));
  generate_exp_tail(ctx, air, d_cols, {144, 14, true, 512, 0, 9 * 16 - 1, false});
  check_chain_error(ctx, err, "FqExpStark: internal error");
  if (h_results) memcpy(h_results, res.data(), n * 32);
}

// ---------------- G2ExpStark (reference src/curves/g2/exp.rs) ----------------
// As k_g1_chain: warp 0 doubles, warp 1 adds, A[k] handed over through shared memory.
__global__ void __launch_bounds__(64) k_g2_chain(const sbn_g2_exp_io* __restrict__ ios, size_t num_io, G2Jac* __restrict__ jac /* [io][2][257] */, int* __restrict__ err) {
  __shared__ G2Jac hand[2][32];
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  const size_t i = blockIdx.x * (size_t)32 + lane;
  const bool live = i < num_io;
  const sbn_g2_exp_io& io = ios[live ? i : 0];
  u32 w[16];
  G2Jac P;
  for (int t = 0; t < 4; t++) {
    u64x4_to_words((const u64*)(role ? io.offset : io.x) + 4 * t, w); if (live && fq_geq_p(w)) *err = 2;
    Fq v = fq_from_words(w); (t == 0 ? P.x.c0 : t == 1 ? P.x.c1 : t == 2 ? P.y.c0 : P.y.c1) = v;
  }
  P.z = fq2_one();
  G2Jac* out = jac + (live ? i : 0) * 2 * 257 + (role ? 257 : 0);
  if (live) out[0] = P;
  for (int k = 0; k < 256; k++) {
    if (role == 0) hand[k & 1][lane] = P;
    __syncthreads();
    if (role == 0) P = g2_jac_dbl(P);
    else if ((io.exp_val[k >> 5] >> (k & 31)) & 1) P = g2_jac_add(hand[k & 1][lane], P);
    if (live) out[k + 1] = P;
  }
}
// G2: doubling chain + prefix-sum scan, as k_g1_dbl_chain / k_g1_sum_scan.
HD void g2_load_point(const u64* src, G2Jac& P, bool& bad) {
  u32 w[16];
  for (int t = 0; t < 4; t++) {
    u64x4_to_words(src + 4 * t, w); if (fq_geq_p(w)) bad = true;
    Fq v = fq_from_words(w); (t == 0 ? P.x.c0 : t == 1 ? P.x.c1 : t == 2 ? P.y.c0 : P.y.c1) = v;
  }
  P.z = fq2_one();
}
__global__ void __launch_bounds__(32) k_g2_dbl_chain(const sbn_g2_exp_io* __restrict__ ios, size_t num_io, G2Jac* __restrict__ jac, int* __restrict__ err) {
  const size_t i = blockIdx.x * (size_t)32 + threadIdx.x;
  if (i >= num_io) return;
  G2Jac P; bool bad = false;
  g2_load_point((const u64*)ios[i].x, P, bad);
  if (bad) *err = 2;
  G2Jac* out = jac + i * 2 * 257;
  out[0] = P;
  for (int k = 0; k < 256; k++) { P = g2_jac_dbl(P); out[k + 1] = P; }
}
HD G2Jac g2_jac_add_id(const G2Jac& p, const G2Jac& q, int* err) {
  if (fq2_is_zero(p.z)) return q;
  if (fq2_is_zero(q.z)) return p;
  G2Jac r = g2_jac_add(p, q);
  if (fq2_is_zero(r.z)) *err = 1;
  return r;
}
__global__ void __launch_bounds__(256) k_g2_sum_scan(const sbn_g2_exp_io* __restrict__ ios, G2Jac* __restrict__ jac, int* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char g2_scan_smem[];
  G2Jac* buf = reinterpret_cast<G2Jac*>(g2_scan_smem);
  const size_t i = blockIdx.x;
  const int j = threadIdx.x;
  const sbn_g2_exp_io& io = ios[i];
  G2Jac* A = jac + i * 2 * 257;
  G2Jac* B = A + 257;
  G2Jac v;
  if ((io.exp_val[j >> 5] >> (j & 31)) & 1) v = A[j];
  else { v.x = fq2_one(); v.y = fq2_one(); v.z.c0 = fq_zero(); v.z.c1 = fq_zero(); }
#pragma unroll 1
  for (int d = 1; d < 256; d <<= 1) {
    buf[j] = v;
    __syncthreads();
    if (j >= d) v = g2_jac_add_id(buf[j - d], v, err);
    __syncthreads();
  }
  G2Jac off; bool bad = false;
  g2_load_point((const u64*)io.offset, off, bad);
  if (bad) *err = 2;
  if (j == 0) B[0] = off;
  B[j + 1] = g2_jac_add_id(off, v, err);
}
__global__ void __launch_bounds__(128) k_g2_affine(const G2Jac* __restrict__ jac, size_t npoints, u32* __restrict__ aff /* [point][32] */, int* __restrict__ err) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npoints) return;
  G2Jac p = jac[i];
  if (fq2_is_zero(p.z)) { *err = 1; return; }
  Fq2 zi = fq2_inv(p.z), zi2 = fq2_sqr(zi);
  Fq2 x = fq2_mul(p.x, zi2), y = fq2_mul(p.y, fq2_mul(zi2, zi));
  u32 w[16];
  fq2_to_words(x, w);
#pragma unroll
  for (int k = 0; k < 16; k++) aff[i * 32 + k] = w[k];
  fq2_to_words(y, w);
#pragma unroll
  for (int k = 0; k < 16; k++) aff[i * 32 + 16 + k] = w[k];
}
// main columns: a(64) b(64) G2Output(640) flags(14)   (reference src/curves/g2/exp.rs:180-245)
__global__ void __launch_bounds__(128) k_g2_rows(const sbn_g2_exp_io* __restrict__ ios, const u32* __restrict__ aff, u64* __restrict__ cols, size_t N, int* __restrict__ err) {
  const size_t r = exp_row_of_thread();
  if (r >= N) return;
  const size_t inst = r >> 9; const int rr = (int)(r & 511), k = rr >> 1;
  u32 e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = ios[inst].exp_val[i];
  u64 fl[14];
  flags_row(e, rr, fl);
  const u32* pa = aff + ((inst * 2) * 257 + k) * 32;
  const u32* pb = aff + ((inst * 2 + 1) * 257 + k + (rr & 1)) * 32;
  u32 ax[16], ay[16], bx[16], by[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { ax[i] = pa[i]; ay[i] = pa[16 + i]; bx[i] = pb[i]; by[i] = pb[16 + i]; }
  const int op = fl[2] ? EXP_OP_SQUARE : (fl[4] ? EXP_OP_MUL : EXP_OP_NONE);   // a = is_double, filtered_bit = is_add
  ColWriter w{cols + r, N};
  if (!g2_row(ax, ay, bx, by, op, w)) *err = 1;
#pragma unroll
  for (int i = 0; i < 14; i++) w(768 + i, fl[i]);
}
static void generate_g2(sbn_ctx* ctx, const AirDesc& air, const void* ios, bool on_device, u64* d_cols, u64* h_results) {
  const size_t n = air.num_io, N = air.num_rows;
  DevBuf<sbn_g2_exp_io> buf; const sbn_g2_exp_io* d_ios = stage_ios(ctx, ios, on_device, n, buf);
  DevBuf<int> err(ctx, 1);
  CUDA_CHECK(cudaMemsetAsync(err, 0, 4, ctx->stream));
  const size_t npoints = n * 2 * 257;
  DevBuf<G2Jac> jac(ctx, npoints);
  DevBuf<u32> aff(ctx, npoints * 32);
  const bool seq_chain = getenv("SBN_CHAIN") && !strcmp(getenv("SBN_CHAIN"), "sequential");
  if (seq_chain) { KScope ks(ctx, "g2_chain"); k_g2_chain<<<(unsigned)((n + 31) / 32), 64, 0, ctx->stream>>>(d_ios, n, jac, err); LAUNCH_CHECK(ctx); }
  else {
    KScope ks(ctx, "g2_chain");
    const int smem = (int)(256 * sizeof(G2Jac));
    CUDA_CHECK(cudaFuncSetAttribute(k_g2_sum_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_g2_dbl_chain<<<(unsigned)((n + 31) / 32), 32, 0, ctx->stream>>>(d_ios, n, jac, err); LAUNCH_CHECK(ctx);
    k_g2_sum_scan<<<(unsigned)n, 256, smem, ctx->stream>>>(d_ios, jac, err); LAUNCH_CHECK(ctx);
  }
  { KScope ks(ctx, "g2_affine"); k_g2_affine<<<(unsigned)((npoints + 127) / 128), 128, 0, ctx->stream>>>(jac, npoints, aff, err); LAUNCH_CHECK(ctx); }
  { KScope ks(ctx, "g2_rows"); k_g2_rows<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(d_ios, aff, d_cols, N, err); LAUNCH_CHECK(ctx); }
  std::vector<u32> res(n * 32);
  CUDA_CHECK(cudaMemcpy2DAsync(res.data(), 128, aff + (size_t)(257 + 256) * 32, (size_t)2 * 257 * 32 * 4, 128, n, cudaMemcpyDeviceToHost, ctx->stream));
  generate_exp_tail(ctx, air, d_cols, {768, 14, true, 512, 0, 48 * 16 - 6, false});
  check_chain_error(ctx, err, "degenerate G2 input: the chain hit the point at infinity or two points with equal x (the reference panics here)");
  if (h_results) memcpy(h_results, res.data(), n * 128);
}

// ---------------- Fq12ExpStark / Fq12ExpU64Stark (reference src/fields/fq12/exp.rs, src/fields/fq12_u64/exp_u64.rs) ----------------
struct Fq12Io { u64 x[48], offset[48]; };   // common prefix of sbn_fq12_exp_io and sbn_fq12_exp_u64_io
HD bool fq12_exp_bit(const void* io_base, size_t io_size, size_t inst, int k, bool u64_variant) {
  const unsigned char* p = (const unsigned char*)io_base + inst * io_size + 768;
  if (u64_variant) return (*(const u64*)p >> k) & 1;
  return (((const u32*)p)[k >> 5] >> (k & 31)) & 1;
}
// One block per instance, 288 threads.  Per bit: thread (which, i, j) forms one pairwise product (which = 0: A_i A_j for the
// squaring, 1: B_i A_j for the multiplication, skipped when the bit is clear); 24 threads sum the 144 products of their
// output coefficient while 24 threads of another warp convert the current A, B to canonical words and store them.
// chain[inst][2][nbits+1][12][8]: canonical words of A[k] = x^(2^k) and B[k] = offset * prod_{j<k, bit_j} A[j].
__global__ void __launch_bounds__(288) k_fq12_chain(const void* __restrict__ ios, size_t io_size, int nbits, bool u64_variant, u32* __restrict__ chain, int* __restrict__ err) {
  __shared__ Fq sa[12], sb[12], pr[2][144];
  const size_t inst = blockIdx.x;
  const int t = threadIdx.x, which = t / 144, ij = t % 144, pi = ij / 12, pj = ij % 12;
  const Fq12Io* io = (const Fq12Io*)((const unsigned char*)ios + inst * io_size);
  if (t < 24) {
    u32 w[8];
    u64x4_to_words((const u64*)(t < 12 ? io->x : io->offset) + 4 * (t % 12), w);
    if (fq_geq_p(w)) *err = 2;
    (t < 12 ? sa : sb)[t % 12] = fq_from_words(w);
  }
  __syncthreads();
  u32* ca = chain + inst * 2 * (size_t)(nbits + 1) * 96; u32* cb = ca + (size_t)(nbits + 1) * 96;
  for (int k = 0; k <= nbits; k++) {
    const bool bit = k < nbits && fq12_exp_bit(ios, io_size, inst, k, u64_variant);
    if (k < nbits && (which == 0 || bit)) pr[which][ij] = fq_mul((which ? sb : sa)[pi], sa[pj]);
    if (t >= 32 && t < 56) {   // canonical words of the current values (a separate warp from the summing threads)
      const int c = t - 32;
      fq_to_words((c < 12 ? sa : sb)[c % 12], (c < 12 ? ca : cb) + ((size_t)k * 12 + c % 12) * 8);
    }
    if (k == nbits) break;
    __syncthreads();
    Fq nv;
    if (t < 12) nv = fq12_sum_coeff(pr[0], t);
    else if (t < 24 && bit) nv = fq12_sum_coeff(pr[1], t - 12);
    __syncthreads();
    if (t < 12) sa[t] = nv; else if (t < 24 && bit) sb[t - 12] = nv;
    __syncthreads();
  }
}
// main columns: a(192) b(192) Fq12Output(1344) flags   (reference src/fields/fq12/exp.rs:142-214).  Block = 32 rows x 12
// coefficients; thread (row, oi) produces coefficient oi of the row's a, b and output block.
__global__ void __launch_bounds__(384) k_fq12_rows(const void* __restrict__ ios, size_t io_size, int nbits, bool u64_variant, const u32* __restrict__ chain,
                                                   u64* __restrict__ cols, size_t N) {
  __shared__ unsigned short sx[32][192], sy[32][192];
  const int lr = threadIdx.x & 31, oi = threadIdx.x >> 5;
  const size_t r = blockIdx.x * (size_t)32 + lr;
  const int rows_per_io = 2 * nbits;
  const size_t inst = r / rows_per_io; const int rr = (int)(r % rows_per_io), k = rr >> 1;
  const bool square = rr & 1;
  const bool bit = fq12_exp_bit(ios, io_size, inst, k, u64_variant);
  const int op = square ? EXP_OP_SQUARE : (bit ? EXP_OP_MUL : EXP_OP_NONE);
  const u32* ca = chain + inst * 2 * (size_t)(nbits + 1) * 96; const u32* cb = ca + (size_t)(nbits + 1) * 96;
  const u32* pa = ca + ((size_t)k * 12 + oi) * 8;                       // a = A[k]
  const u32* pb = cb + ((size_t)(k + (square ? 1 : 0)) * 12 + oi) * 8;  // b = B[k] on even rows, B[k+1] on odd rows
  u32 aw[8], bw[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { aw[i] = pa[i]; bw[i] = pb[i]; }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    sx[lr][16 * oi + 2 * i] = aw[i] & 0xFFFF; sx[lr][16 * oi + 2 * i + 1] = aw[i] >> 16;
    const u32 yv = square ? aw[i] : bw[i];
    sy[lr][16 * oi + 2 * i] = yv & 0xFFFF; sy[lr][16 * oi + 2 * i + 1] = yv >> 16;
  }
  __syncthreads();
  ColWriter w{cols + r, N};
  write_limbs16(w, 16 * oi, aw); write_limbs16(w, 192 + 16 * oi, bw);
  // output = a*a = A[k+1] on square rows, a*b = B[k+1] on multiplication rows
  const u32* po = (square ? ca : cb) + ((size_t)(k + 1) * 12 + oi) * 8;
  u32 ow[8];
#pragma unroll
  for (int i = 0; i < 8; i++) ow[i] = op == EXP_OP_NONE ? 0 : po[i];
  fq12_row_coeff(sx[lr], sy[lr], ow, oi, op, w);
  if (oi == 0) {
    const int sf = 108 * 16;
    const unsigned char* ep = (const unsigned char*)ios + inst * io_size + 768;
    if (u64_variant) {
      u64 fl[6]; flags_u64_row(*(const u64*)ep, rr, fl);
      for (int i = 0; i < 6; i++) w(sf + i, fl[i]);
    } else {
      u32 e[8]; for (int i = 0; i < 8; i++) e[i] = ((const u32*)ep)[i];
      u64 fl[14]; flags_row(e, rr, fl);
      for (int i = 0; i < 14; i++) w(sf + i, fl[i]);
    }
  }
}
static void generate_fq12(sbn_ctx* ctx, const AirDesc& air, const void* ios, bool on_device, u64* d_cols, u64* h_results) {
  const bool u64v = air.air_id == SBN_AIR_FQ12_EXP_U64;
  const int nbits = u64v ? 64 : 256;
  const size_t n = air.num_io, N = air.num_rows, io_size = air.io_size;
  static_assert(sizeof(sbn_fq12_exp_io) == 1184 && sizeof(sbn_fq12_exp_u64_io) == 1160, "io layout");
  DevBuf<unsigned char> buf; const void* d_ios = stage_ios<unsigned char>(ctx, ios, on_device, n * io_size, buf);
  if (u64v) {   // exp_val must be a canonical field element (it is a public input: exp_u64.rs:106)
    std::vector<unsigned char> h(n * io_size);
    if (on_device) { CUDA_CHECK(cudaMemcpyAsync(h.data(), ios, h.size(), cudaMemcpyDeviceToHost, ctx->stream)); ctx->sync(); }
    else memcpy(h.data(), ios, h.size());
    for (size_t i = 0; i < n; i++) { u64 e; memcpy(&e, h.data() + i * io_size + 768, 8); SBN_REQUIRE(e < GL_P, "Fq12ExpU64Stark: exp_val is not a canonical field element"); }
  }
  DevBuf<int> err(ctx, 1);
  CUDA_CHECK(cudaMemsetAsync(err, 0, 4, ctx->stream));
  const size_t per_inst = (size_t)2 * (nbits + 1) * 96;
  DevBuf<u32> chain(ctx, n * per_inst);
  { KScope ks(ctx, "fq12_chain"); k_fq12_chain<<<(unsigned)n, 288, 0, ctx->stream>>>(d_ios, io_size, nbits, u64v, chain, err); LAUNCH_CHECK(ctx); }
  { KScope ks(ctx, "fq12_rows"); k_fq12_rows<<<(unsigned)(N / 32), 384, 0, ctx->stream>>>(d_ios, io_size, nbits, u64v, chain, d_cols, N); LAUNCH_CHECK(ctx); }
  std::vector<u32> res(n * 96);
  CUDA_CHECK(cudaMemcpy2DAsync(res.data(), 384, chain + (size_t)(nbits + 1) * 96 + (size_t)nbits * 96, per_inst * 4, 384, n, cudaMemcpyDeviceToHost, ctx->stream));
  generate_exp_tail(ctx, air, d_cols, {108 * 16, u64v ? 6 : 14, !u64v, 2 * nbits, 24 * 16, 84 * 16 - 12, true});
  check_chain_error(ctx, err, "Fq12ExpStark: internal error");
  if (h_results) memcpy(h_results, res.data(), n * 384);
}

static void generate_modular(sbn_ctx* ctx, const AirDesc& air, const void* ios, bool on_device, u64* d_cols) {
  const size_t N = air.num_rows;
  DevBuf<u64> d_ios_buf; const u64* d_ios = (const u64*)ios;
  if (!on_device) {
    d_ios_buf = DevBuf<u64>(ctx, N * 8);
    CUDA_CHECK(cudaMemcpyAsync(d_ios_buf, ios, N * 64, cudaMemcpyHostToDevice, ctx->stream));
    d_ios = d_ios_buf;
  }
  DevBuf<int> err(ctx, 1);
  CUDA_CHECK(cudaMemsetAsync(err, 0, 4, ctx->stream));
  { KScope ks(ctx, "modular_rows"); k_modular_rows<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(d_ios, d_cols, N, err); LAUNCH_CHECK(ctx); }
  generate_split_u16_range_check_cols(ctx, d_cols, N, 32, 111, 145);
  int h_err = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  SBN_REQUIRE(!h_err, "ModularStark input is not a canonical Fq residue");
}

// ---------------- gadget test AIRs: G1Stark (one addition per row), Fq12Stark (one product per row) ----------------
// main columns: a(32) b(32) G1Output(320) is_add is_double   (reference src/curves/g1/muladd.rs:481-546)
__global__ void __launch_bounds__(128) k_g1_muladd_rows(const sbn_g1_muladd_io* __restrict__ ios, u64* __restrict__ cols, size_t N, int* __restrict__ err) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  u32 ax[8], ay[8], bx[8], by[8];
  u64x4_to_words((const u64*)ios[r].a_x, ax); u64x4_to_words((const u64*)ios[r].a_y, ay); u64x4_to_words((const u64*)ios[r].b_x, bx); u64x4_to_words((const u64*)ios[r].b_y, by);
  if (fq_geq_p(ax) || fq_geq_p(ay) || fq_geq_p(bx) || fq_geq_p(by)) { *err = 2; return; }
  ColWriter w{cols + r, N};
  if (!g1_row(ax, ay, bx, by, G1_OP_ADD, w)) *err = 1;
  w(384, 1); w(385, 0);
}
// main columns: x(192) y(192) Fq12Output(1344) filter   (reference src/fields/fq12/mul.rs:378-419).  Block = 32 rows x 12
// coefficients; thread (row, oi) produces coefficient oi of the row's product and its reduction witness.
__global__ void __launch_bounds__(384) k_fq12_mul_rows(const sbn_fq12_mul_io* __restrict__ ios, u64* __restrict__ cols, size_t N, int* __restrict__ err) {
  __shared__ unsigned short sx[32][192], sy[32][192];
  __shared__ Fq mx[32][12], my[32][12];
  const int lr = threadIdx.x & 31, oi = threadIdx.x >> 5;
  const size_t r = blockIdx.x * (size_t)32 + lr;   // N is a multiple of 32
  u32 xw[8], yw[8];
  u64x4_to_words((const u64*)ios[r].x + 4 * oi, xw); u64x4_to_words((const u64*)ios[r].y + 4 * oi, yw);
  if (fq_geq_p(xw) || fq_geq_p(yw)) *err = 2;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    sx[lr][16 * oi + 2 * i] = xw[i] & 0xFFFF; sx[lr][16 * oi + 2 * i + 1] = xw[i] >> 16;
    sy[lr][16 * oi + 2 * i] = yw[i] & 0xFFFF; sy[lr][16 * oi + 2 * i + 1] = yw[i] >> 16;
  }
  mx[lr][oi] = fq_from_words(xw); my[lr][oi] = fq_from_words(yw);
  __syncthreads();
  ColWriter w{cols + r, N};
  write_limbs16(w, 16 * oi, xw); write_limbs16(w, 192 + 16 * oi, yw);
  u32 ow[8];
  fq_to_words(fq12_mul_coeff(mx[lr], my[lr], oi), ow);
  fq12_row_coeff(sx[lr], sy[lr], ow, oi, EXP_OP_MUL, w);
  if (oi == 0) w(108 * 16, 1);
}
static void generate_gadget(sbn_ctx* ctx, const AirDesc& air, const void* ios, bool on_device, u64* d_cols) {
  const size_t N = air.num_rows;
  DevBuf<unsigned char> buf; const void* d_ios = ios;
  if (!on_device) {
    buf = DevBuf<unsigned char>(ctx, N * air.io_size);
    CUDA_CHECK(cudaMemcpyAsync(buf, ios, N * air.io_size, cudaMemcpyHostToDevice, ctx->stream));
    d_ios = buf;
  }
  DevBuf<int> err(ctx, 1);
  CUDA_CHECK(cudaMemsetAsync(err, 0, 4, ctx->stream));
  if (air.air_id == SBN_AIR_G1_MULADD) {
    { KScope ks(ctx, "g1_muladd_rows"); k_g1_muladd_rows<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>((const sbn_g1_muladd_io*)d_ios, d_cols, N, err); LAUNCH_CHECK(ctx); }
    generate_split_u16_range_check_cols(ctx, d_cols, N, 4 * 16, 20 * 16 - 4, 24 * 16 + 2);
  } else {
    { KScope ks(ctx, "fq12_mul_rows"); k_fq12_mul_rows<<<(unsigned)(N / 32), 384, 0, ctx->stream>>>((const sbn_fq12_mul_io*)d_ios, d_cols, N, err); LAUNCH_CHECK(ctx); }
    generate_split_u16_range_check_cols(ctx, d_cols, N, 24 * 16, 84 * 16 - 12, 108 * 16 + 1);
  }
  int h_err = 0;
  CUDA_CHECK(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  SBN_REQUIRE(h_err != 2, "gadget AIR input is not a canonical Fq residue");
  SBN_REQUIRE(h_err != 1, "G1Stark: the two points of a row have equal x (the addition gadget divides by b.x - a.x)");
}

void generate_trace(sbn_ctx* ctx, const AirDesc& air, const void* ios, bool ios_on_device, u64* d_cols, u64* h_results) {
  switch (air.air_id) {
    case SBN_AIR_G1_MULADD: case SBN_AIR_FQ12_MUL: generate_gadget(ctx, air, ios, ios_on_device, d_cols); break;
    case SBN_AIR_MODULAR: generate_modular(ctx, air, ios, ios_on_device, d_cols); break;
    case SBN_AIR_G1_EXP: generate_g1(ctx, air, ios, ios_on_device, d_cols, h_results); break;
    case SBN_AIR_FQ_EXP: generate_fq(ctx, air, ios, ios_on_device, d_cols, h_results); break;
    case SBN_AIR_G2_EXP: generate_g2(ctx, air, ios, ios_on_device, d_cols, h_results); break;
    case SBN_AIR_FQ12_EXP: case SBN_AIR_FQ12_EXP_U64: generate_fq12(ctx, air, ios, ios_on_device, d_cols, h_results); break;
    default: throw SbnError(-3, "unknown AIR identifier");
  }
}

void format_public_inputs(const AirDesc& air, const void* ios, u64* out) {
  switch (air.air_id) {
    case SBN_AIR_MODULAR: case SBN_AIR_G1_MULADD: case SBN_AIR_FQ12_MUL: break;
    case SBN_AIR_G1_EXP: {  // reference src/curves/g1/exp.rs:124-135: x.x x.y offset.x offset.y exp_val output.x output.y, 8 u32 limbs each
      const sbn_g1_exp_io* h = (const sbn_g1_exp_io*)ios;
      for (size_t i = 0; i < air.num_io; i++) {
        u64* o = out + 56 * i;
        auto put = [&](const uint64_t* v, int slot) { for (int k = 0; k < 8; k++) o[8 * slot + k] = (v[k >> 1] >> (32 * (k & 1))) & 0xFFFFFFFFULL; };
        put(h[i].x_x, 0); put(h[i].x_y, 1); put(h[i].offset_x, 2); put(h[i].offset_y, 3);
        for (int k = 0; k < 8; k++) o[32 + k] = h[i].exp_val[k];
        put(h[i].output_x, 5); put(h[i].output_y, 6);
      }
      break;
    }
    case SBN_AIR_FQ_EXP: {  // reference src/fields/fq/exp.rs:103-111: x offset exp_val output, 8 u32 limbs each
      const sbn_fq_exp_io* h = (const sbn_fq_exp_io*)ios;
      for (size_t i = 0; i < air.num_io; i++) {
        u64* o = out + 32 * i;
        auto put = [&](const uint64_t* v, int slot) { for (int k = 0; k < 8; k++) o[8 * slot + k] = (v[k >> 1] >> (32 * (k & 1))) & 0xFFFFFFFFULL; };
        put(h[i].x, 0); put(h[i].offset, 1);
        for (int k = 0; k < 8; k++) o[16 + k] = h[i].exp_val[k];
        put(h[i].output, 3);
      }
      break;
    }
    case SBN_AIR_G2_EXP: {  // reference src/curves/g2/exp.rs:139-156: x(4) offset(4) exp_val output(4), 8 u32 limbs each
      const sbn_g2_exp_io* h = (const sbn_g2_exp_io*)ios;
      for (size_t i = 0; i < air.num_io; i++) {
        u64* o = out + 104 * i;
        auto put = [&](const uint64_t* v, int slot) { for (int k = 0; k < 8; k++) o[8 * slot + k] = (v[k >> 1] >> (32 * (k & 1))) & 0xFFFFFFFFULL; };
        for (int t = 0; t < 4; t++) { put(h[i].x + 4 * t, t); put(h[i].offset + 4 * t, 4 + t); put(h[i].output + 4 * t, 9 + t); }
        for (int k = 0; k < 8; k++) o[64 + k] = h[i].exp_val[k];
      }
      break;
    }
    case SBN_AIR_FQ12_EXP: case SBN_AIR_FQ12_EXP_U64: {  // reference src/fields/fq12/exp.rs:107-124, fq12_u64/exp_u64.rs:102-122: u16 limbs
      const bool u64v = air.air_id == SBN_AIR_FQ12_EXP_U64;
      const size_t io_len = u64v ? 577 : 584;
      for (size_t i = 0; i < air.num_io; i++) {
        const unsigned char* rec = (const unsigned char*)ios + i * air.io_size;
        u64* o = out + io_len * i;
        auto put12 = [&](const unsigned char* src, u64* dst) { for (int k = 0; k < 192; k++) { uint16_t v; memcpy(&v, src + 2 * k, 2); dst[k] = v; } };   // little-endian host
        put12(rec, o); put12(rec + 384, o + 192);
        if (u64v) { u64 e; memcpy(&e, rec + 768, 8); SBN_REQUIRE(e < GL_P, "exp_val is not a canonical field element"); o[384] = e; put12(rec + 776, o + 385); }
        else { for (int k = 0; k < 8; k++) { uint32_t e; memcpy(&e, rec + 768 + 4 * k, 4); o[384 + k] = e; } put12(rec + 800, o + 392); }
      }
      break;
    }
    default: throw SbnError(-3, "unknown AIR identifier");
  }
}
