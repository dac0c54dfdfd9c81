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
