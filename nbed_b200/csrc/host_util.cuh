// Host-side plumbing of the C-ABI library: error type, device buffers, stage timers, the context struct.
#pragma once
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a profiler is attached
#include <cstdarg>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/nbed_b200.h"

namespace nbd {

struct Error {
  int code;
  std::string msg;
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error{code, buf};
}

#define NBD_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) ::nbd::fail(NBD_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#define NBD_SOLVER(call)                                                                            \
  do {                                                                                              \
    cusolverStatus_t s_ = (call);                                                                   \
    if (s_ != CUSOLVER_STATUS_SUCCESS) ::nbd::fail(NBD_ERR_CUDA, "%s:%d %s: cusolver status %d", __FILE__, __LINE__, #call, (int)s_); \
  } while (0)
#define NBD_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) ::nbd::fail(code, __VA_ARGS__); \
  } while (0)

// Grow-only device buffer.
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  T* ensure(size_t n) {
    if (n > cap) {
      if (p) cudaFree(p);
      p = nullptr;
      cap = 0;
      NBD_CUDA(cudaMalloc(&p, n * sizeof(T)));
      cap = n;
    }
    return p;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  ~DBuf() { release(); }
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
};

// Pinned host bounce buffer (grow-only).
struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  void* ensure(size_t bytes) {
    if (bytes > cap) {
      if (p) cudaFreeHost(p);
      p = nullptr;
      cap = 0;
      NBD_CUDA(cudaMallocHost(&p, bytes));
      cap = bytes;
    }
    return p;
  }
  ~PinnedBuf() {
    if (p) cudaFreeHost(p);
  }
};

// Per-stage device timers: event pairs recorded on the stream, resolved after a synchronize.
struct StageTimers {
  struct Rec {
    std::string key;
    cudaEvent_t a, b;
  };
  std::vector<cudaEvent_t> pool;
  std::vector<Rec> open;
  std::map<std::string, double> ms;
  bool enabled = true;
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    NBD_CUDA(cudaEventCreate(&e));
    return e;
  }
  void reset() { ms.clear(); }
  void resolve() {
    for (auto& r : open) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) ms[r.key] += t;
      pool.push_back(r.a);
      pool.push_back(r.b);
    }
    open.clear();
  }
  ~StageTimers() {
    for (auto& r : open) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    for (auto e : pool) cudaEventDestroy(e);
  }
};

struct StageScope {
  StageTimers* t;
  cudaStream_t st;
  size_t idx;
  bool on;
  StageScope(StageTimers& timers, cudaStream_t stream, const char* key) : t(&timers), st(stream), on(timers.enabled) {
    if (!on) return;
    nvtxRangePushA(key);  // same names as the nbd_timer_ms keys: `ncu --nvtx --nvtx-include "jk_x/"` selects a stage
    StageTimers::Rec r{key, t->get(), t->get()};
    cudaEventRecord(r.a, st);
    t->open.push_back(r);
    idx = t->open.size() - 1;
  }
  ~StageScope() {
    if (!on) return;
    cudaEventRecord(t->open[idx].b, st);
    nvtxRangePop();
  }
};

}  // namespace nbd
