// Shared host-side plumbing of the B200 prover library: context, error reporting, a caching device
// allocator (buffers are reused across proofs so the steady state performs no cudaMalloc), and
// per-size twiddle tables.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <chrono>
#include <thread>
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

// Read-only device tables (twiddles, four-step factors, power tables) of one GPU.  A context owns its own set unless it was
// created by a batch (sbn_batch_create), whose lanes share one: a table is built on the requesting context's stream, that
// stream is synchronised, and only then is the pointer published under the mutex, so any other stream may read it.
struct SharedTables {
  std::mutex mu;
  // U1: multiplicative generator = coset shift, and the generator of the 2^32 roots of unity derived
  // from it as g^((p - 1) / 2^32).  Default: plonky2_field pair (B) = (7, 1753635133440165772).  Tables depend on the pair.
  u64 mult_gen = GL_MULT_GENERATOR, pow2_gen = GL_POW2_GENERATOR;
  void select_generator(u64 g) {   // callers hold no table pointers across this (start of a prove / batch call)
    if (g == 0) g = GL_MULT_GENERATOR;
    std::lock_guard<std::mutex> lk(mu);
    if (g == mult_gen) return;
    const u64 r = gl_pow(g, 0xFFFFFFFFULL);   // (p - 1) / 2^32 = 2^32 - 1
    if (g >= GL_P || gl_exp_pow2(r, 31) != GL_P - 1) throw SbnError(SBN_ERR_INVALID, "coset_shift is not a generator of the multiplicative group (two-adic part of order < 2^32)");
    cudaDeviceSynchronize();
    for (auto& kv : pow_tables) cudaFree(kv.second);
    for (auto& kv : fourstep_tables) cudaFree(kv.second);
    pow_tables.clear(); fourstep_tables.clear(); ntt_tables.clear();
    mult_gen = g; pow2_gen = r;
  }
  std::map<int, NttTables> ntt_tables;
  std::map<std::tuple<int, bool, u64>, u64*> fourstep_tables;   // (logn, inverse, coset base) -> ntt.cu F table
  std::map<std::pair<u64, int>, u64*> pow_tables;               // (base, logn) -> base^i, i < 2^logn
  ~SharedTables() {
    for (auto& kv : pow_tables) cudaFree(kv.second);
    for (auto& kv : fourstep_tables) cudaFree(kv.second);
  }
};

struct sbn_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  int num_sms = 148;
  std::string last_error;
  std::shared_ptr<SharedTables> tables = std::make_shared<SharedTables>();
  unsigned ntt_attr_mask = 0;                      // sub-transform sizes whose kernels already have their shared-memory opt-in on this device
  int live_handles = 0;                            // sbn_trace objects still holding buffers of this context
  bool destroy_pending = false;                    // sbn_ctx_destroy was called with live handles: the last sbn_trace_free destroys
  // Host waits: poll the stream, yielding the core between polls, and after ~300 polls (a few hundred microseconds: longer than
  // the short kernels most waits are for) sleep between polls, 20 us growing to 100 us.  A blocking event wait measured ~1 ms per
  // wake-up on the bench box (170 vs 87 ms for one G1 proof); a pure spin / yield loop keeps one core per lane busy while a 25 ms
  // leaf-hash kernel runs, and eight ranks x six lanes on a 32-core box then slow each other down (8-GPU step 66.3 vs 63.2 ms).
  void sync() {
    for (int polls = 0;; polls++) {
      cudaError_t e = cudaStreamQuery(stream);
      if (e == cudaSuccess) return;
      if (e != cudaErrorNotReady) throw SbnError(-2, std::string("stream synchronisation: ") + cudaGetErrorString(e));
      if (polls < 300) std::this_thread::yield();
      else std::this_thread::sleep_for(std::chrono::microseconds(polls < 600 ? 20 : 100));
    }
  }
  // Small host -> device uploads (challenge-dependent tables, descriptors) go through a pinned bump arena: a cudaMemcpyAsync from
  // pageable memory first waits for everything queued on the stream.  The arena is rewound at the start of every entry point
  // (nothing of the previous call is still in flight: every entry point ends with a host synchronisation).
  uint8_t* pinned = nullptr; size_t pinned_cap = 0, pinned_used = 0;
  void upload(void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return;
    const size_t need = (bytes + 63) & ~size_t(63);
    if (!pinned) { pinned_cap = size_t(8) << 20; if (cudaMallocHost((void**)&pinned, pinned_cap) != cudaSuccess) { pinned = nullptr; pinned_cap = 0; cudaGetLastError(); } }
    if (pinned_used + need > pinned_cap) {   // too large (or arena missing): plain copy, then wait so that `src` may die
      cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
      if (e != cudaSuccess) throw SbnError(-2, std::string("cudaMemcpyAsync: ") + cudaGetErrorString(e));
      sync();
      return;
    }
    memcpy(pinned + pinned_used, src, bytes);
    cudaError_t e = cudaMemcpyAsync(dst, pinned + pinned_used, bytes, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) throw SbnError(-2, std::string("cudaMemcpyAsync: ") + cudaGetErrorString(e));
    pinned_used += need;
  }
  void begin_call() { pinned_used = 0; }
  u64 coset_shift() const { return tables->mult_gen; }
  u64 root_of_unity(int logn) const { return gl_exp_pow2(tables->pow2_gen, 32 - logn); }
  // caching allocator
  struct Block { void* p; size_t bytes; bool used; };
  std::vector<Block> blocks;
  size_t bytes_allocated = 0;
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
  void trim() {   // give the cached (unused) blocks back to the device; blocks in use stay
    for (auto it = blocks.begin(); it != blocks.end();) { if (!it->used) { cudaFree(it->p); bytes_allocated -= it->bytes; it = blocks.erase(it); } else ++it; }
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
