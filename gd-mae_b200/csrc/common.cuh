// Shared helpers for the gdmae_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define GDMAE_OK 0
#define GDMAE_ERR_ARG (-1)
#define GDMAE_ERR_WORKSPACE (-2)
#define GDMAE_ERR_CUDA (-3)

extern "C" void gdmae_set_error(const char* msg);

#define GDMAE_CHECK_ARG(cond)                                                                   \
  do {                                                                                          \
    if (!(cond)) {                                                                              \
      char _b[256];                                                                             \
      snprintf(_b, sizeof(_b), "%s:%d invalid argument: %s", __FILE__, __LINE__, #cond);        \
      gdmae_set_error(_b);                                                                      \
      return GDMAE_ERR_ARG;                                                                     \
    }                                                                                           \
  } while (0)

#define GDMAE_CHECK_CUDA(expr)                                                                  \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      char _b[256];                                                                             \
      snprintf(_b, sizeof(_b), "%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      gdmae_set_error(_b);                                                                      \
      return GDMAE_ERR_CUDA;                                                                    \
    }                                                                                           \
  } while (0)

extern "C" void gdmae_count_launch(void);
// every hand-written kernel launch is followed by this macro: it also feeds gdmae_launch_count()
#define GDMAE_LAUNCH_CHECK()                 \
  do {                                       \
    gdmae_count_launch();                    \
    GDMAE_CHECK_CUDA(cudaGetLastError());    \
  } while (0)

static inline int gdmae_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t gdmae_align(size_t x) { return (x + 255) & ~(size_t)255; }

// B200: 148 SMs.  Grid-stride kernels are sized as a multiple of the SM count.
#define GDMAE_NUM_SMS 148
static inline int gdmae_grid(long long work_items, int threads, int max_ctas_per_sm = 16) {
  long long need = (work_items + threads - 1) / threads;
  long long cap = (long long)GDMAE_NUM_SMS * max_ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Bump allocator over the caller-provided workspace.
struct Workspace {
  char* base;
  size_t size, off;
  __host__ Workspace(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <typename T>
  __host__ T* take(size_t count) {
    size_t bytes = gdmae_align(count * sizeof(T));
    if (off + bytes > size) return nullptr;
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};

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
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
extern "C" int gdmae_timing_on(void);
extern "C" void gdmae_timing_push(int kind, int d, int64_t n, int64_t bytes, void* e0, void* e1);
// bench-only CUDA events around one launch sequence (no-op unless gdmae_timing_enable(1))
struct GdmaeSpan {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t st;
  explicit GdmaeSpan(cudaStream_t s) : st(s) {
    if (gdmae_timing_on()) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, st);
    }
  }
  void end(int kind, int d, int64_t n, int64_t bytes) {
    if (e0) {
      cudaEventRecord(e1, st);
      gdmae_timing_push(kind, d, n, bytes, e0, e1);
    }
  }
};


