// Micro-benchmark (not product code): issue throughput of the integer instructions the Goldilocks / Poseidon kernels are
// made of, in warp-instructions per clock per SM (4 sub-partitions => 4.0 = one instruction per scheduler per clock).
// Usage: int_throughput   (prints one line per instruction mix)
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
#define ITER 4096
#define CH 8   // independent chains per thread

template <int MODE> __global__ void __launch_bounds__(256) k(u32* out, u32 seed) {
  u32 a[CH], b[CH]; u64 w[CH]; double f[CH], g[CH]; const double fc = (double)(seed | 41);
#pragma unroll
  for (int i = 0; i < CH; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i * 5 + 1; w[i] = ((u64)a[i] << 32) | b[i]; f[i] = (double)a[i]; g[i] = 1.0 / (double)(b[i] | 1) * 1e-3; }
  u32 c = seed | 41;
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(c));                    // IMAD.WIDE.U32 accumulate
      if (MODE == 1) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((u32)w[i]), "r"(c));                    // IMAD.WIDE.U32 no addend
      if (MODE == 2) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));                       // IMAD (32-bit)
      if (MODE == 3) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));                                      // IMAD.HI.U32
      if (MODE == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                                      // IADD3
      if (MODE == 5) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(c), "r"(seed));   // IADD3 + IADD3.X (64-bit add)
      if (MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));                   // LOP3
      if (MODE == 7) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));                            // SHF
      if (MODE == 8) asm volatile("prmt.b32 %0, %0, %1, 0x1032;" : "+r"(a[i]) : "r"(b[i]));                             // PRMT
      if (MODE == 9) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"((u32)w[i])); }   // 1 IMAD.WIDE : 1 IADD
      if (MODE == 10) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(c)); asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"((u32)w[i]), "r"((u32)(w[i] >> 32))); }  // 1 IMAD.WIDE : 2 IADD
      if (MODE == 11) { asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i])); }   // 1 IMAD : 1 IADD
      if (MODE == 12) asm volatile("{.reg .u32 t; mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;}" : "+r"(a[i]), "+r"(b[i]) : "r"(c), "r"(seed));  // mad.lo.cc + madc.hi
      if (MODE == 13) { u64 t = w[i] * (u64)c; w[i] = t + (w[i] >> 3); }   // 64-bit mul.lo (compiler's choice)
      if (MODE == 14) w[i] = __umul64hi(w[i], w[i] | 1) + w[i];            // 64x64 high + add
      if (MODE == 16) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));                    // IDP.4A
      if (MODE == 17) { asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c)); asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(a[i]), "r"(c)); }   // dp4a : IMAD 1:1
      if (MODE == 18) { asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(c)); }   // dp4a : LOP3 1:1
      if (MODE == 19) { if (i % 4 == 0) asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+r"(a[i]), "+r"(a[i + 1]), "+r"(a[i + 2]), "+r"(a[i + 3]) : "r"(b[i]), "r"(b[i + 1]), "r"(c)); }   // 2 IMMA.16816 per body
      if (MODE == 20) { if (i % 4 == 0) asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+r"(a[i]), "+r"(a[i + 1]), "+r"(a[i + 2]), "+r"(a[i + 3]) : "r"(b[i]), "r"(b[i + 1]), "r"(b[i + 2]), "r"(b[i + 3]), "r"(c), "r"(seed)); }   // 2 IMMA.16832 per body
      if (MODE == 21) { if (i % 4 == 0) asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+r"(a[i]), "+r"(a[i + 1]), "+r"(a[i + 2]), "+r"(a[i + 3]) : "r"(b[i]), "r"(b[i + 1]), "r"(c)); else asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(b[i]), "r"(c)); }   // 2 IMMA.16816 + 6 IMAD per body
      if (MODE == 22) { asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(c)); }   // IMAD : LOP3 1:1
      if (MODE == 23) { asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(b[i]), "r"(c)); { u32 wl = (u32)w[i]; asm volatile("add.u32 %0, %0, %1;" : "+r"(wl) : "r"(c)); w[i] = wl; } }
      if (MODE == 15) asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w[i]) : "r"(a[i]), "r"(c), "l"(w[(i + 1) % CH]));   // IMAD.WIDE with a different addend register
      if (MODE == 24) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(g[i]), "d"(fc));                       // DFMA
      if (MODE == 25) { asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(g[i]), "d"(fc)); asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c)); }   // DFMA : IDP 1:1
      if (MODE == 26) { asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(g[i]), "d"(fc)); asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(b[i]), "r"(c)); }   // DFMA : IDP : LOP3 1:1:1
      if (MODE == 27) { asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(g[i]), "d"(fc)); asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c)); asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(a[i]), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(b[i]), "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c)); }   // DFMA : IDP : ALU 1:2:2
      if (MODE == 28) { asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(g[i]), "d"(fc)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(b[i]), "r"(c)); }   // DFMA : LOP3 1:1
      if (MODE == 29) asm volatile("add.f64 %0, %0, %1;" : "+d"(f[i]) : "d"(fc));                                        // DADD
      if (MODE == 30) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));                 // IDP.2A
    }
  }
  u32 r = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) r ^= a[i] ^ b[i] ^ (u32)w[i] ^ (u32)(w[i] >> 32) ^ (u32)__double_as_longlong(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
static const char* NAMES[] = {"IMAD.WIDE.U32 acc (mad.wide)", "IMAD.WIDE.U32 (mul.wide)", "IMAD 32 (mad.lo)", "IMAD.HI.U32 (mul.hi)", "IADD3 (add)", "IADD3+IADD3.X (64-bit add)",
                              "LOP3", "SHF", "PRMT", "mul.wide + add (1:1)", "mul.wide + 64-bit add (1:2)", "mad.lo + add (1:1)", "mad.lo.cc + madc.hi", "64-bit mul.lo + add", "umul64hi + add",
                              "IMAD.WIDE acc, other addend", "IDP.4A (dp4a)", "dp4a + mad.lo (1:1)", "dp4a + lop3 (1:1)", "2 IMMA m16n8k16 u8 per body", "2 IMMA m16n8k32 u8 per body",
                              "2 IMMA.16816 + 6 IMAD per body", "mad.lo + lop3 (1:1)", "mad.lo + lop3 + add (1:1:1)", "DFMA", "DFMA + dp4a (1:1)", "DFMA + dp4a + lop3 (1:1:1)", "DFMA + 2 dp4a + lop3 + add", "DFMA + lop3 (1:1)", "DADD", "IDP.2A (dp2a)"};
static const int PTX_PER_ITER[] = {1, 1, 1, 1, 1, 2, 1, 1, 1, 2, 3, 2, 2, 0, 0, 1, 1, 2, 2, 0, 0, 0, 2, 3, 1, 2, 3, 5, 2, 1, 1};
template <int MODE> void run(u32* d, int sms, double mhz) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks = sms * 8;
  k<MODE><<<blocks, 256>>>(d, 3);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(d, 5);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_iters = (double)blocks * 8 * ITER * CH;   // warp-level loop bodies
  double clk = ms * 1e-3 * mhz * 1e6;
  double per_sm_clk = warp_iters / sms / clk;
  printf("%-34s %8.3f ms  %6.3f bodies/clk/SM", NAMES[MODE], ms, per_sm_clk);
  if (PTX_PER_ITER[MODE]) printf("  = %6.3f PTX-instr/clk/SM", per_sm_clk * PTX_PER_ITER[MODE]);
  printf("\n");
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; double mhz = p.clockRate / 1e3;
  printf("%s, %d SMs, %.0f MHz (nominal max; the run may clock differently)\n", p.name, sms, mhz);
  u32* d; cudaMalloc(&d, sms * 8 * 256 * 4);
  run<0>(d, sms, mhz); run<1>(d, sms, mhz); run<2>(d, sms, mhz); run<3>(d, sms, mhz); run<4>(d, sms, mhz); run<5>(d, sms, mhz); run<6>(d, sms, mhz); run<7>(d, sms, mhz);
  run<8>(d, sms, mhz); run<9>(d, sms, mhz); run<10>(d, sms, mhz); run<11>(d, sms, mhz); run<12>(d, sms, mhz); run<13>(d, sms, mhz); run<14>(d, sms, mhz); run<15>(d, sms, mhz);
  run<16>(d, sms, mhz); run<17>(d, sms, mhz); run<18>(d, sms, mhz); run<19>(d, sms, mhz); run<20>(d, sms, mhz); run<21>(d, sms, mhz); run<22>(d, sms, mhz); run<23>(d, sms, mhz);
  run<24>(d, sms, mhz); run<25>(d, sms, mhz); run<26>(d, sms, mhz); run<27>(d, sms, mhz); run<28>(d, sms, mhz); run<29>(d, sms, mhz); run<30>(d, sms, mhz);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
