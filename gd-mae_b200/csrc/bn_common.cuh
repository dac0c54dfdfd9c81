// Shared pieces of the BatchNorm kernels (batchnorm.cu, vfe_mlp.cu): the per-CTA partial layout
// [sum(C) | sumsq(C)] (forward) / [dbeta(C) | dgamma(C)] (backward) and its fp64 final combination.
#pragma once
#include "common.cuh"

#define BN_PART_BLOCKS (GDMAE_NUM_SMS * 4)

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per channel: lanes stride over the per-block partials
static __global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, double count, float eps,
                                                          float momentum, float* __restrict__ mean, float* __restrict__ rstd,
                                                          float* __restrict__ running_mean, float* __restrict__ running_var) {
  int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = lane; b < nblocks; b += 32) {
    s += (double)partial[(long long)b * 2 * C + c];
    q += (double)partial[(long long)b * 2 * C + C + c];
  }
  s = warp_sum_f64(s);
  q = warp_sum_f64(q);
  if (lane != 0) return;
  double m = s / count;
  double var = q / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

static __global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C,
                                                              const float* __restrict__ extra_dbeta, const float* __restrict__ extra_dgamma,
                                                              float* __restrict__ dbeta, float* __restrict__ dgamma) {
  int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = lane; b < nblocks; b += 32) {
    s += (double)partial[(long long)b * 2 * C + c];
    q += (double)partial[(long long)b * 2 * C + C + c];
  }
  s = warp_sum_f64(s);
  q = warp_sum_f64(q);
  if (lane != 0) return;
  if (extra_dbeta) s += (double)extra_dbeta[c];
  if (extra_dgamma) q += (double)extra_dgamma[c];
  dbeta[c] = (float)s;
  dgamma[c] = (float)q;
}

