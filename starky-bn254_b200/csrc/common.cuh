// Shared host-side plumbing of the B200 prover library: context, error reporting, a caching device
// allocator (buffers are reused across proofs so the steady state performs no cudaMalloc), and
// per-size twiddle tables.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <string>
#include <vector>
#include <stdexcept>
#include "gl.cuh"

struct SbnError : std::runtime_error { int code; SbnError(int c, const std::string& m) : std::runtime_error(m), code(c) {} };
#include "../../include/starky_bn254_b200.h"

#define CUDA_CHECK(x)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (x);                                                                               \
    if (e_ != cudaSuccess) {                                                                            \
      char b_[512]; snprintf(b_, sizeof b_, "%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      throw SbnError(SBN_ERR_CUDA, b_);                                                                 \
    }                                                                                                   \
  } while (0)
#define SBN_REQUIRE(c, msg) do { if (!(c)) throw SbnError(SBN_ERR_INVALID, msg); } while (0)

struct NttTables {  // for domain size N = 2^logn: powers of w_N and w_N^-1 (N entries each, device)
  int logn = 0; u64* w_fwd = nullptr; u64* w_inv = nullptr;
};

struct sbn_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  int num_sms = 148;
  std::string last_error;
  // caching allocator
  struct Block { void* p; size_t bytes; bool used; };
  std::vector<Block> blocks;
  size_t bytes_allocated = 0;
  std::map<int, NttTables> ntt_tables;
  std::map<std::tuple<int, bool, u64>, u64*> fourstep_tables;   // (logn, inverse, coset base) -> ntt.cu F table
  std::map<std::pair<u64, int>, u64*> pow_tables;  // (base, logn) -> base^i, i < 2^logn
  unsigned long long launches = 0;                 // kernels launched by this library (bench: gpu_launches)
  // optional per-kernel-family CUDA-event timing on ctx->stream (bench.py roofline + breakdown)
  bool ktime_enabled = false;
  struct KPending { std::string name; cudaEvent_t a, b; };
  std::vector<KPending> kpending;
  std::vector<cudaEvent_t> kpool;
  struct KStat { double ms = 0; unsigned long long count = 0; };
  std::map<std::string, KStat> kstats;
  cudaEvent_t kevent() { if (!kpool.empty()) { cudaEvent_t e = kpool.back(); kpool.pop_back(); return e; } cudaEvent_t e; cudaEventCreate(&e); return e; }
  void kresolve() {
    for (auto& p : kpending) {
      cudaEventSynchronize(p.b);
      float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b);
      auto& st = kstats[p.name]; st.ms += ms; st.count++;
      kpool.push_back(p.a); kpool.push_back(p.b);
    }
    kpending.clear();
  }

  void* alloc(size_t bytes) {
    if (bytes == 0) bytes = 8;
    bytes = (bytes + 255) & ~size_t(255);
    Block* best = nullptr;
    for (auto& b : blocks) if (!b.used && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 && (!best || b.bytes < best->bytes)) best = &b;
    if (best) { best->used = true; return best->p; }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      // release cached free blocks and retry once
      for (auto it = blocks.begin(); it != blocks.end();) { if (!it->used) { cudaFree(it->p); bytes_allocated -= it->bytes; it = blocks.erase(it); } else ++it; }
      cudaGetLastError();
      CUDA_CHECK(cudaMalloc(&p, bytes));
    }
    bytes_allocated += bytes;
    blocks.push_back({p, bytes, true});
    return p;
  }
  template <class T> T* alloc_n(size_t n) { return (T*)alloc(n * sizeof(T)); }
  void free(void* p) {
    if (!p) return;
    for (auto& b : blocks) if (b.p == p) { b.used = false; return; }
  }
  void release_all() {
    for (auto& b : blocks) cudaFree(b.p);
    blocks.clear(); bytes_allocated = 0;
  }
};

// RAII device buffer tied to the context's caching allocator.
template <class T> struct DevBuf {
  sbn_ctx* ctx = nullptr; T* p = nullptr; size_t n = 0;
  DevBuf() {}
  DevBuf(sbn_ctx* c, size_t n_) : ctx(c), p(c->alloc_n<T>(n_)), n(n_) {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : ctx(o.ctx), p(o.p), n(o.n) { o.p = nullptr; }
  DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { reset(); ctx = o.ctx; p = o.p; n = o.n; o.p = nullptr; } return *this; }
  void reset() { if (p && ctx) ctx->free(p); p = nullptr; n = 0; }
  ~DevBuf() { reset(); }
  T* get() const { return p; }
  operator T*() const { return p; }
};

// Times everything enqueued on ctx->stream during its lifetime under `name` (no-op unless enabled).
struct KScope {
  sbn_ctx* ctx; cudaEvent_t a; const char* name;
  KScope(sbn_ctx* c, const char* n) : ctx(c), a(nullptr), name(n) { if (c->ktime_enabled) { a = c->kevent(); cudaEventRecord(a, c->stream); } }
  ~KScope() { if (a) { cudaEvent_t b = ctx->kevent(); cudaEventRecord(b, ctx->stream); ctx->kpending.push_back({name, a, b}); } }
};

#define LAUNCH_CHECK(ctx) do { (ctx)->launches++; CUDA_CHECK(cudaGetLastError()); } while (0)

static inline int ilog2(size_t n) { int l = 0; while ((size_t(1) << l) < n) l++; return l; }
