// Flat-bucket optimizer step: gradient-norm clip + decoupled weight decay + Adam with OneCycle
// hyper-parameters, one pass over one contiguous fp32 bucket (the same bucket NCCL all-reduces).
//
// Replaces (reference file:line, relative to /root/reference):
//   clip_grad_norm_(model.parameters(), 10)      tools/train_utils/train_utils.py:52
//   OptimWrapper.step (true_wd, bn_wd) + Adam    tools/train_utils/optimization/fastai_optim.py:135-152
#include "common.cuh"

__global__ void __launch_bounds__(256) sumsq_kernel(const float4* __restrict__ g, long long n4, const float* __restrict__ tail,
                                                    int ntail, double* __restrict__ out) {
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(g + i);
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < ntail) acc += (double)tail[threadIdx.x] * tail[threadIdx.x];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(out, s);
  }
}

// out (1, double, caller zeroes) += sum g^2
extern "C" int gdmae_grad_sumsq(const float* grads, int64_t n, double* out, void* stream_) {
  GDMAE_CHECK_ARG(n >= 0 && ((uintptr_t)grads % 16) == 0);
  if (n == 0) return GDMAE_OK;
  long long n4 = n / 4;
  sumsq_kernel<<<gdmae_grid(n4 > 0 ? n4 : 1, 256, 4), 256, 0, (cudaStream_t)stream_>>>((const float4*)grads, n4, grads + 4 * n4,
                                                                                     (int)(n - 4 * n4), out);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

struct AdamArgs {
  float clip, decay, mom, beta2, eps, step_size, bc2_sqrt, grad_scale;
};

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, const double* __restrict__ sumsq, AdamArgs a) {
  float total_norm = (float)sqrt(*sumsq) * a.grad_scale;
  float coef = fminf(a.clip / (total_norm + 1e-6f), 1.0f) * a.grad_scale;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gg = g[i] * coef;
    float pp = p[i] * a.decay;
    float mm = m[i] * a.mom + (1.f - a.mom) * gg;
    float vv = v[i] * a.beta2 + (1.f - a.beta2) * gg * gg;
    float denom = sqrtf(vv) / a.bc2_sqrt + a.eps;
    p[i] = pp - a.step_size * (mm / denom);
    m[i] = mm;
    v[i] = vv;
  }
}

// params/grads/exp_avg/exp_avg_sq: the first n_opt elements of the flat bucket (the parameters the
// reference's flatten_model() hands to Adam); sumsq covers the WHOLE bucket (clip norm counts the
// attention in-proj/tau gradients too).  grad_scale = 1/world_size when the bucket holds an
// all-reduce SUM.  decay = 1 - wd*lr, step_size = lr / (1 - mom^t), bc2_sqrt = sqrt(1 - beta2^t).
extern "C" int gdmae_adam_onecycle_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n_opt,
                                        const double* sumsq, float clip, float decay, float mom, float beta2, float eps,
                                        float step_size, float bc2_sqrt, float grad_scale, void* stream_) {
  GDMAE_CHECK_ARG(n_opt >= 0);
  if (n_opt == 0) return GDMAE_OK;
  AdamArgs a{clip, decay, mom, beta2, eps, step_size, bc2_sqrt, grad_scale};
  adam_kernel<<<gdmae_grid(n_opt, 256, 8), 256, 0, (cudaStream_t)stream_>>>(params, grads, exp_avg, exp_avg_sq, n_opt, sumsq, a);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
