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

// One CTA per 32 channels: lane = channel (coalesced 128-byte reads of the partial rows), the 8 warps take the rows
// b = warp, warp + 8, ... and meet in shared memory; fp64 accumulation like the forward statistics.  (The first version -
// one warp per channel, lanes striding over the rows - read 32 sectors per load and took ~22 us per call, 0.24 ms per step.)
// Launch with gdmae_div_up(C, 32) CTAs of 256 threads: CTA i owns channels [32 i, 32 i + 32).
static __global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C,
                                                              const float* __restrict__ extra_dbeta, const float* __restrict__ extra_dgamma,
                                                              float* __restrict__ dbeta, float* __restrict__ dgamma) {
  __shared__ double red[2][8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s = 0.0, q = 0.0;
  if (c < C) {
    // eight partial rows per trip: sixteen independent loads in flight per thread (r2 ncu: the one-row-per-trip loop cost
    // 40 us per call cold - 74 dependent trips of L2 latency - and 0.44 ms per step over its 11 calls)
    int b = wid;
    for (; b + 56 < nblocks; b += 64) {
      float vs[8], vq[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        vs[u] = __ldg(partial + (long long)(b + 8 * u) * 2 * C + c);
        vq[u] = __ldg(partial + (long long)(b + 8 * u) * 2 * C + C + c);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { s += (double)vs[u]; q += (double)vq[u]; }
    }
    for (; b < nblocks; b += 8) {
      s += (double)__ldg(partial + (long long)b * 2 * C + c);
      q += (double)__ldg(partial + (long long)b * 2 * C + C + c);
    }
  }
  red[0][wid][lane] = s;
  red[1][wid][lane] = q;
  __syncthreads();
  if (wid != 0 || c >= C) return;
  for (int w = 1; w < 8; ++w) { s += red[0][w][lane]; q += red[1][w][lane]; }
  if (extra_dbeta) s += (double)extra_dbeta[c];
  if (extra_dgamma) q += (double)extra_dgamma[c];
  dbeta[c] = (float)s;
  dgamma[c] = (float)q;
}

