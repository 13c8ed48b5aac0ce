// Shared helpers for libvoxactb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "../../include/voxactb.h"

namespace vxb {

void set_error(const char* fmt, ...);

#define VXB_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      vxb::set_error(__VA_ARGS__);               \
      return VXB_E_BADARG;                       \
    }                                            \
  } while (0)

#define VXB_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      vxb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                     __FILE__, __LINE__);                                           \
      return VXB_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

#define VXB_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      vxb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),    \
                     __FILE__, __LINE__);                                           \
      return VXB_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

#define VXB_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != 0) return _s;      \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Bump allocator over a caller-provided arena. With base == nullptr it only measures.
struct Arena {
  char* base;
  size_t off;
  size_t cap;
  bool ok;
  Arena(void* b, size_t c) : base((char*)b), off(0), cap(c), ok(true) {}
  template <typename T>
  T* get(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    size_t o = off;
    off += bytes;
    if (base == nullptr) return (T*)nullptr;
    if (off > cap) {
      ok = false;
      return (T*)nullptr;
    }
    return (T*)(base + o);
  }
};

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// counter-based dropout mask of the attention probabilities (train mode, perceiver_lang_io.py:127-128): element `idx` of stream
// `seed` is kept when its hashed 32-bit value >= p * 2^32.  idx = flat index into the [B*H*Nq, ld] probability matrix
// (ld = Nk rounded up to 4).  Never stored: the forward (flash_umma.cuh / bwd_ops.cuh) and the backward regenerate it.
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long idx, unsigned int thresh) {
  unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned int)(z >> 32) >= thresh;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace vxb
