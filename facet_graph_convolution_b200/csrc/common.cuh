// Shared helpers for the facetconv_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "facetconv_b200.h"

namespace fgc {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

// Opt-in per-kernel timing (fgc_profile_begin/_end): every launch site marks the stream with an
// event; consecutive marks on the one stream bracket each kernel.
void prof_set_stream(cudaStream_t st);
void prof_mark(const char* name);
extern bool g_prof_on;

inline cudaStream_t as_stream(void* s) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  if (g_prof_on) prof_set_stream(st);
  return st;
}

#define FGC_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::fgc::set_error(__VA_ARGS__);    \
      return FGC_ERR_ARG;               \
    }                                   \
  } while (0)

#define FGC_UNSUPPORTED(cond, ...)      \
  do {                                  \
    if (cond) {                         \
      ::fgc::set_error(__VA_ARGS__);    \
      return FGC_ERR_UNSUPPORTED;       \
    }                                   \
  } while (0)

#define FGC_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::fgc::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                       \
      return FGC_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

// after every kernel launch: surface launch errors, count the launch
#define FGC_LAUNCHED(name)                                                              \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      ::fgc::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));       \
      return FGC_ERR_CUDA;                                                              \
    }                                                                                   \
    ::fgc::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
    if (::fgc::g_prof_on) ::fgc::prof_mark(name);                                       \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Workspace {
  char* base;
  size_t size;
  size_t used;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t off = align_up(used, 256);
    size_t bytes = count * sizeof(T);
    if (base == nullptr || off + bytes > size) {
      used = size + 1;  // poison
      return nullptr;
    }
    used = off + bytes;
    return reinterpret_cast<T*>(base + off);
  }
  bool ok() const { return used <= size; }
};

inline size_t ws_bytes(size_t count, size_t elem) { return align_up(count * elem, 256) + 256; }

int num_sms();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float lrelu_f(float x, float alpha) {
  // relu(x) - alpha*relu(-x)  (reference Code/model.py:828-830)
  return fmaxf(x, 0.f) - alpha * fmaxf(-x, 0.f);
}

}  // namespace fgc
