// Library-wide C-ABI plumbing: last-error string, version, device probe.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

extern "C" void gdmae_set_error(const char* msg) {
  strncpy(g_err, msg, sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}
extern "C" const char* gdmae_last_error(void) { return g_err; }

static unsigned long long g_launches = 0;
extern "C" void gdmae_count_launch(void) { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
// number of hand-written kernels launched by this process so far (CUB primitives not counted)
extern "C" int64_t gdmae_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" int gdmae_version(void) { return 100; }

// 0 when a device of compute capability 10.x is current, negative otherwise (the library carries
// sm_100a SASS only; there is no fallback path).
extern "C" int gdmae_check_device(void) {
  int dev = 0;
  cudaDeviceProp prop;
  GDMAE_CHECK_CUDA(cudaGetDevice(&dev));
  GDMAE_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    gdmae_set_error("gdmae_b200 needs an sm_100a (B200) device");
    return GDMAE_ERR_ARG;
  }
  return GDMAE_OK;
}

// ---- optional kernel timing for bench.py: CUDA events around selected launches made inside the library
// (the encoder-layer executor issues its kernels from C, out of reach of Python-side events)
#include <mutex>
#include <vector>
struct TimedSpan { int kind, d; long long n, bytes; cudaEvent_t e0, e1; };
static std::vector<TimedSpan> g_spans;
static std::mutex g_span_mutex;
static int g_timing = 0;
extern "C" void gdmae_timing_enable(int on) { g_timing = on; }
extern "C" int gdmae_timing_on(void) { return g_timing; }
extern "C" void gdmae_timing_push(int kind, int d, int64_t n, int64_t bytes, void* e0, void* e1) {
  std::lock_guard<std::mutex> lock(g_span_mutex);
  g_spans.push_back(TimedSpan{kind, d, (long long)n, (long long)bytes, (cudaEvent_t)e0, (cudaEvent_t)e1});
}
// meta: 4 int64 per span (kind: 0 SRA forward, 1 SRA backward; d; N; algorithmic bytes); ms: elapsed per span.
// Waits for the spans' events, frees them and returns how many were written (at most cap).
extern "C" int gdmae_timing_drain(int64_t* meta, float* ms, int cap) {
  std::lock_guard<std::mutex> lock(g_span_mutex);
  int n = 0;
  for (auto& s : g_spans) {
    if (n < cap) {
      cudaEventSynchronize(s.e1);
      float t = 0.f;
      cudaEventElapsedTime(&t, s.e0, s.e1);
      meta[4 * n] = s.kind; meta[4 * n + 1] = s.d; meta[4 * n + 2] = s.n; meta[4 * n + 3] = s.bytes;
      ms[n] = t;
      ++n;
    }
    cudaEventDestroy(s.e0);
    cudaEventDestroy(s.e1);
  }
  g_spans.clear();
  return n;
}
