// Pillar feature encoder MLP, forward and backward, as a handful of bandwidth-bound passes that
// materialise only what a later pass must read.
//
// Replaces (reference file:line, relative to /root/reference):
//   DynVFE.forward: dvfe_mlps (Linear -> BatchNorm1d(train) -> ReLU, twice) + torch_scatter.scatter_max
//                                                      pcdet/models/backbones_3d/vfe/dyn_vfe.py:105-111
//   make_fc_layers                                     pcdet/models/model_utils/network_utils.py:7-21
//
// The reference (and the op-by-op path) writes and re-reads four point-sized fp32 tensors per direction
// (y1, h1, y2, h2: 2 GB at 1.27 M points) - ~15 GB of traffic per step.  Here (K = 10 input columns,
// C1 = 64, C2 = 128):
//   forward   y1 = x W1^T is never stored: K is tiny, so the statistics pass and the apply pass both recompute
//             it from x; h1 = relu(bn1(y1)) is stored once (operand dtype), y2 = h1 W2^T comes from cuBLASLt
//             (operand dtype), bn2 + ReLU + the per-pillar max/argmax are ONE pass over y2 - h2 is never stored.
//   backward  the gradient of the max is non-zero only at the argmax rows, so the two BN2 sums are a sparse
//             pass over M x C2 entries; one dense pass writes dy2 (the mean terms reach every row); cuBLASLt
//             gives dW2 and dh1; BN1's sums and its apply pass recompute y1 / the ReLU mask from x, and the
//             apply pass accumulates dW1 = dy1^T x on the fly - dy1 is never stored.
// ~3.5 GB of traffic per step in the bf16 configuration.
#include "common.cuh"
#include <stdlib.h>
#include "bn_common.cuh"
#include "bulk_pipe.cuh"
#include <cuda_bf16.h>
#include "../../include/gdmae_b200.h"

#define V_C1 64
#define V_C2 128
#define V_KMAX 16
#define V_TILE 64       // point rows per shared-memory tile
#define V_THREADS 256   // 8 warps; a warp takes one row at a time, a lane two channels of C1

typedef __nv_bfloat16 vbf16;

template <typename T> struct VT;
template <> struct VT<float> {
  static __device__ __forceinline__ void store2(float* p, long long i2, float a, float b) { reinterpret_cast<float2*>(p)[i2] = make_float2(a, b); }
  static __device__ __forceinline__ float2 load2(const float* p, long long i2) { return __ldg(reinterpret_cast<const float2*>(p) + i2); }
  static __device__ __forceinline__ float4 load4(const float* p, long long i4) { return __ldg(reinterpret_cast<const float4*>(p) + i4); }
  static __device__ __forceinline__ void store4(float* p, long long i4, float4 v) { reinterpret_cast<float4*>(p)[i4] = v; }
};
template <> struct VT<vbf16> {
  static __device__ __forceinline__ void store2(vbf16* p, long long i2, float a, float b) {
    reinterpret_cast<__nv_bfloat162*>(p)[i2] = __floats2bfloat162_rn(a, b);
  }
  static __device__ __forceinline__ float2 load2(const vbf16* p, long long i2) {
    unsigned u = __ldg(reinterpret_cast<const unsigned*>(p) + i2);
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  }
  static __device__ __forceinline__ float4 load4(const vbf16* p, long long i4) {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(p) + i4);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
  }
  static __device__ __forceinline__ void store4(vbf16* p, long long i4, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<unsigned*>(&a);
    u.y = *reinterpret_cast<unsigned*>(&b);
    reinterpret_cast<uint2*>(p)[i4] = u;
  }
};

// ---------------------------------------------------------------------------------- layer 1 (recomputed from x)
struct L1Ctx {
  float w0[V_KMAX], w1[V_KMAX];   // rows 2*c2 and 2*c2+1 of W1
};

__device__ __forceinline__ void l1_load_w(L1Ctx& c, const float* __restrict__ W1, int K, int c2) {
#pragma unroll
  for (int k = 0; k < V_KMAX; ++k) {
    c.w0[k] = k < K ? __ldg(W1 + (2 * c2) * K + k) : 0.f;
    c.w1[k] = k < K ? __ldg(W1 + (2 * c2 + 1) * K + k) : 0.f;
  }
}

// tile of V_TILE rows of x -> shared memory (coalesced)
__device__ __forceinline__ void l1_load_tile(float* xs, const float* __restrict__ x, long long row0, long long Np, int K) {
  const long long base = row0 * K;
  const long long lim = (Np - row0 < V_TILE ? Np - row0 : V_TILE) * K;
  for (int i = threadIdx.x; i < V_TILE * K; i += V_THREADS) xs[i] = i < lim ? __ldg(x + base + i) : 0.f;
}

__device__ __forceinline__ void l1_y(const L1Ctx& c, const float* xr, int K, float& y0, float& y1) {
  y0 = 0.f; y1 = 0.f;
#pragma unroll
  for (int k = 0; k < V_KMAX; ++k) {
    if (k < K) {
      const float xv = xr[k];
      y0 = fmaf(xv, c.w0[k], y0);
      y1 = fmaf(xv, c.w1[k], y1);
    }
  }
}

// per-CTA combine of (s0, s1, q0, q1) over the 8 warps -> partial[block] = [first(C1) | second(C1)]
__device__ __forceinline__ void l1_block_partial(float s0, float s1, float q0, float q1, float* __restrict__ partial) {
  __shared__ float red[8][4][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  red[wid][0][lane] = s0; red[wid][1][lane] = s1; red[wid][2][lane] = q0; red[wid][3][lane] = q1;
  __syncthreads();
  if (wid == 0) {
    for (int w = 1; w < 8; ++w) { s0 += red[w][0][lane]; s1 += red[w][1][lane]; q0 += red[w][2][lane]; q1 += red[w][3][lane]; }
    float* dst = partial + (long long)blockIdx.x * 2 * V_C1;
    dst[2 * lane] = s0; dst[2 * lane + 1] = s1;
    dst[V_C1 + 2 * lane] = q0; dst[V_C1 + 2 * lane + 1] = q1;
  }
}

__global__ void __launch_bounds__(V_THREADS) vfe1_stats_kernel(const float* __restrict__ x, long long Np, int K,
                                                               const float* __restrict__ W1, float* __restrict__ partial) {
  __shared__ float xs[V_TILE * V_KMAX];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  L1Ctx c;
  l1_load_w(c, W1, K, lane);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const long long ntile = (Np + V_TILE - 1) / V_TILE;
  for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    __syncthreads();
    l1_load_tile(xs, x, tile * V_TILE, Np, K);
    __syncthreads();
    const int nrow = (int)min((long long)V_TILE, Np - tile * V_TILE);
    for (int r = wid; r < nrow; r += 8) {
      float y0, y1;
      l1_y(c, xs + r * K, K, y0, y1);
      s0 += y0; s1 += y1;
      q0 = fmaf(y0, y0, q0); q1 = fmaf(y1, y1, q1);
    }
  }
  l1_block_partial(s0, s1, q0, q1, partial);
}

template <typename T>
__global__ void __launch_bounds__(V_THREADS) vfe1_apply_kernel(const float* __restrict__ x, long long Np, int K,
                                                               const float* __restrict__ W1, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, T* __restrict__ h1) {
  __shared__ float xs[V_TILE * V_KMAX];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  L1Ctx c;
  l1_load_w(c, W1, K, lane);
  const float m0 = mean[2 * lane], m1 = mean[2 * lane + 1];
  const float a0 = rstd[2 * lane] * gamma[2 * lane], a1 = rstd[2 * lane + 1] * gamma[2 * lane + 1];
  const float b0 = beta[2 * lane], b1 = beta[2 * lane + 1];
  const long long ntile = (Np + V_TILE - 1) / V_TILE;
  for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    __syncthreads();
    l1_load_tile(xs, x, tile * V_TILE, Np, K);
    __syncthreads();
    const int nrow = (int)min((long long)V_TILE, Np - tile * V_TILE);
    for (int r = wid; r < nrow; r += 8) {
      float y0, y1;
      l1_y(c, xs + r * K, K, y0, y1);
      VT<T>::store2(h1, (tile * V_TILE + r) * (V_C1 / 2) + lane, fmaxf(fmaf(y0 - m0, a0, b0), 0.f), fmaxf(fmaf(y1 - m1, a1, b1), 0.f));
    }
  }
}

// backward sums of BN1: dbeta = sum g, dgamma = sum g * xhat, g = dh1 where relu(bn1(y1)) > 0
template <typename T>
__global__ void __launch_bounds__(V_THREADS) vfe1_bwd_stats_kernel(const float* __restrict__ x, long long Np, int K,
                                                                   const float* __restrict__ W1, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, const T* __restrict__ dh1,
                                                                   float* __restrict__ partial) {
  __shared__ float xs[V_TILE * V_KMAX];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  L1Ctx c;
  l1_load_w(c, W1, K, lane);
  const float m0 = mean[2 * lane], m1 = mean[2 * lane + 1], r0 = rstd[2 * lane], r1 = rstd[2 * lane + 1];
  const float a0 = r0 * gamma[2 * lane], a1 = r1 * gamma[2 * lane + 1];
  const float b0 = beta[2 * lane], b1 = beta[2 * lane + 1];
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const long long ntile = (Np + V_TILE - 1) / V_TILE;
  for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    __syncthreads();
    l1_load_tile(xs, x, tile * V_TILE, Np, K);
    __syncthreads();
    const int nrow = (int)min((long long)V_TILE, Np - tile * V_TILE);
    float2 gv[V_TILE / 8];
#pragma unroll
    for (int i = 0; i < V_TILE / 8; ++i) {
      const int r = wid + 8 * i;
      gv[i] = r < nrow ? VT<T>::load2(dh1, (tile * V_TILE + r) * (V_C1 / 2) + lane) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < V_TILE / 8; ++i) {
      const int r = wid + 8 * i;
      if (r < nrow) {
        float y0, y1;
        l1_y(c, xs + r * K, K, y0, y1);
        const float g0 = fmaf(y0 - m0, a0, b0) > 0.f ? gv[i].x : 0.f, g1 = fmaf(y1 - m1, a1, b1) > 0.f ? gv[i].y : 0.f;
        s0 += g0; s1 += g1;
        q0 = fmaf(g0, (y0 - m0) * r0, q0); q1 = fmaf(g1, (y1 - m1) * r1, q1);
      }
    }
  }
  l1_block_partial(s0, s1, q0, q1, partial);
}

// dy1 = gamma rstd (g - dbeta/n - xhat dgamma/n) is formed row by row and folded into dW1 = dy1^T x at once
template <typename T>
__global__ void __launch_bounds__(V_THREADS) vfe1_bwd_wgrad_kernel(const float* __restrict__ x, long long Np, int K,
                                                                   const float* __restrict__ W1, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, const T* __restrict__ dh1,
                                                                   const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                                                                   float inv_n, float* __restrict__ dW1) {
  __shared__ float xs[V_TILE * V_KMAX];
  __shared__ float red[8][V_C1][V_KMAX + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  L1Ctx c;
  l1_load_w(c, W1, K, lane);
  const float m0 = mean[2 * lane], m1 = mean[2 * lane + 1], r0 = rstd[2 * lane], r1 = rstd[2 * lane + 1];
  const float a0 = r0 * gamma[2 * lane], a1 = r1 * gamma[2 * lane + 1];
  const float b0 = beta[2 * lane], b1 = beta[2 * lane + 1];
  const float c10 = a0 * dbeta[2 * lane] * inv_n, c11 = a1 * dbeta[2 * lane + 1] * inv_n;
  const float c20 = a0 * dgamma[2 * lane] * inv_n, c21 = a1 * dgamma[2 * lane + 1] * inv_n;
  float acc0[V_KMAX], acc1[V_KMAX];
#pragma unroll
  for (int k = 0; k < V_KMAX; ++k) acc0[k] = acc1[k] = 0.f;
  const long long ntile = (Np + V_TILE - 1) / V_TILE;
  for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    __syncthreads();
    l1_load_tile(xs, x, tile * V_TILE, Np, K);
    __syncthreads();
    const int nrow = (int)min((long long)V_TILE, Np - tile * V_TILE);
    float2 gv[V_TILE / 8];
#pragma unroll
    for (int i = 0; i < V_TILE / 8; ++i) {       // the warp's 8 rows of the tile: all loads in flight before the math
      const int r = wid + 8 * i;
      gv[i] = r < nrow ? VT<T>::load2(dh1, (tile * V_TILE + r) * (V_C1 / 2) + lane) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < V_TILE / 8; ++i) {
      const int r = wid + 8 * i;
      if (r < nrow) {
        const float* xr = xs + r * K;
        float y0, y1;
        l1_y(c, xr, K, y0, y1);
        const float g0 = fmaf(y0 - m0, a0, b0) > 0.f ? gv[i].x : 0.f, g1 = fmaf(y1 - m1, a1, b1) > 0.f ? gv[i].y : 0.f;
        const float d0 = a0 * g0 - c10 - (y0 - m0) * r0 * c20, d1 = a1 * g1 - c11 - (y1 - m1) * r1 * c21;
#pragma unroll
        for (int k = 0; k < V_KMAX; ++k) {
          if (k < K) {
            acc0[k] = fmaf(d0, xr[k], acc0[k]);
            acc1[k] = fmaf(d1, xr[k], acc1[k]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < V_KMAX; ++k) { red[wid][2 * lane][k] = acc0[k]; red[wid][2 * lane + 1][k] = acc1[k]; }
  __syncthreads();
  for (int i = threadIdx.x; i < V_C1 * K; i += V_THREADS) {
    const int ch = i / K, k = i - ch * K;
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][ch][k];
    atomicAdd(dW1 + i, t);   // ~300 CTAs x 640 entries; the host wrapper zeroes dW1 when it does not accumulate
  }
}

// ---------------------------------------------------------------------------------- layer 1 through the input moments (r2)
// y1 = W1 x is linear in x, so every point-sized reduction of layer 1 that does not involve a ReLU mask follows from the
// first two moments of x alone:  S1 = sum_p x_p (K),  S2 = sum_p x_p x_p^T (K x K, upper triangle).
//   forward   mean(y1_c) = w_c . S1 / n,   E[y1_c^2] = w_c^T S2 w_c / n                       (replaces vfe1_stats_kernel)
//   backward  dW1 = sum_p dy1_p x_p^T with dy1 = a (g - dbeta/n - xhat dgamma/n), a = rstd gamma, g = dh1 [h1 > 0]:
//             dW1[c, k] = a_c G[c, k] - a_c dbeta_c / n S1[k] - a_c dgamma_c / n rstd_c (w_c . S2[:, k] - mean_c S1[k])
//             where only G = sum_p g_p x_p^T needs the data - and the same pass yields dbeta = sum g, dgamma = sum g xhat with
//             xhat = (h1 - beta) / gamma wherever the ReLU is active (g = 0 elsewhere): ONE pass over h1, dh1, x replaces
//             vfe1_bwd_stats_kernel + vfe1_bwd_wgrad_kernel, which each recomputed the K = 10 product per point and ran at
//             0.06-0.09 of the HBM rate (r2 ncu: 309 + 385 us, issue-bound).
// The moments are accumulated in fp32 per thread over a handful of rows and combined in fp64.
template <int K>
struct Mom {
  static constexpr int NM = K + K * (K + 1) / 2;
};

template <int K>
__global__ void __launch_bounds__(256) vfe1_moments_kernel(const float* __restrict__ x, long long Np, double* __restrict__ partial) {
  constexpr int NM = Mom<K>::NM;
  float acc[NM];
#pragma unroll
  for (int e = 0; e < NM; ++e) acc[e] = 0.f;
  for (long long row = blockIdx.x * 256ll + threadIdx.x; row < Np; row += gridDim.x * 256ll) {
    float v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = __ldg(x + row * K + k);
    int e = K;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      acc[i] += v[i];
#pragma unroll
      for (int j = i; j < K; ++j) {
        acc[e] = fmaf(v[i], v[j], acc[e]);
        ++e;
      }
    }
  }
  __shared__ double red[8][NM];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < NM; ++e) {
    double d = (double)acc[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) red[wid][e] = d;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NM; e += 256) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w][e];
    partial[(long long)blockIdx.x * NM + e] = t;
  }
}

// moments[0:K] = S1, moments[K:] = upper triangle of S2 (row-major, j >= i); then BatchNorm statistics of y1 = W1 x
template <int K>
__global__ void __launch_bounds__(256) vfe1_moments_finalize_kernel(const double* __restrict__ partial, int nblocks, const float* __restrict__ W1,
                                                                    double count, float eps, float momentum, double* __restrict__ moments,
                                                                    float* __restrict__ mean, float* __restrict__ rstd,
                                                                    float* __restrict__ running_mean, float* __restrict__ running_var) {
  constexpr int NM = Mom<K>::NM;
  __shared__ double mom[NM];
  for (int e = threadIdx.x; e < NM; e += 256) {
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += partial[(long long)b * NM + e];
    mom[e] = t;
    moments[e] = t;
  }
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= V_C1) return;
  double m = 0.0, sq = 0.0;
  int e = K;
  for (int i = 0; i < K; ++i) {
    const double wi = (double)W1[c * K + i];
    m += wi * mom[i];
    for (int j = i; j < K; ++j, ++e) sq += (i == j ? 1.0 : 2.0) * wi * (double)W1[c * K + j] * mom[e];
  }
  m /= count;
  double var = sq / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// one pass over h1, dh1, x: per CTA partial[block] = [ s (64) | q (64) | G (64 x K) ].  Thread = 4 channels of one row.
template <typename T, int K>
__global__ void __launch_bounds__(256) vfe1_bwd_pass_kernel(const float* __restrict__ x, const T* __restrict__ h1, const T* __restrict__ dh1,
                                                            long long Np, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ partial) {
  constexpr int W = 2 + K;                       // values per channel
  const int cg = threadIdx.x & 15, rsub = threadIdx.x >> 4;
  float ig[4], be[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float g = gamma[4 * cg + j];
    ig[j] = fabsf(g) > 1e-20f ? 1.f / g : 0.f;
    be[j] = beta[4 * cg + j];
  }
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f}, G[4][K];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) G[j][k] = 0.f;
  for (long long row = blockIdx.x * 16ll + rsub; row < Np; row += gridDim.x * 16ll) {
    const float4 h = VT<T>::load4(h1, row * (V_C1 / 4) + cg), d = VT<T>::load4(dh1, row * (V_C1 / 4) + cg);
    float xv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) xv[k] = __ldg(x + row * K + k);
    const float hv[4] = {h.x, h.y, h.z, h.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float g = hv[j] > 0.f ? dv[j] : 0.f;
      s[j] += g;
      q[j] = fmaf(g, (hv[j] - be[j]) * ig[j], q[j]);
#pragma unroll
      for (int k = 0; k < K; ++k) G[j][k] = fmaf(g, xv[k], G[j][k]);
    }
  }
  // the two rows of a warp meet by shuffle, the eight warps in shared memory
  __shared__ float red[8][V_C1 * W];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
    q[j] += __shfl_xor_sync(0xffffffffu, q[j], 16);
#pragma unroll
    for (int k = 0; k < K; ++k) G[j][k] += __shfl_xor_sync(0xffffffffu, G[j][k], 16);
  }
  if (lane < 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = 4 * cg + j;
      red[wid][ch] = s[j];
      red[wid][V_C1 + ch] = q[j];
#pragma unroll
      for (int k = 0; k < K; ++k) red[wid][2 * V_C1 + ch * K + k] = G[j][k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V_C1 * W; i += 256) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][i];
    partial[(long long)blockIdx.x * (V_C1 * W) + i] = t;
  }
}

// ---- the same pass and the forward apply pass with their rows staged by bulk copies (bf16 configuration) ----------------
// A tile = 64 consecutive point rows: h1 and dh1 (8 KB each) and x (64 K floats) are contiguous byte ranges, so one elected
// thread brings a tile with three cp.async.bulk transactions into a ring of VB_STAGES stages; the 256 threads consume a stage
// with the arithmetic of the kernels above and hand it back with one CTA barrier.  Bytes in flight per SM = CTAs x stages x
// 19 KB instead of "resident warps x one row": the register-staged form ran at 1.6 TB/s (235 us).  Only whole tiles are
// handled here; the host runs the generic kernel on the last Np % 64 rows.
#define VB_ROWS 64
#define VB_STAGES 4
template <int K> struct VbStage {
  static constexpr int H_BYTES = VB_ROWS * V_C1 * 2, X_BYTES = VB_ROWS * K * 4;
  static constexpr int BYTES = 2 * H_BYTES + ((X_BYTES + 127) / 128) * 128;
};

template <int K>
__global__ void __launch_bounds__(256, 2) vfe1_bwd_pass_bulk_kernel(const float* __restrict__ x, const vbf16* __restrict__ h1,
                                                                    const vbf16* __restrict__ dh1, long long ntile,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    float* __restrict__ partial) {
  using S = VbStage<K>;
  constexpr int W = 2 + K;
  extern __shared__ __align__(128) unsigned char vb_smem[];
  __shared__ unsigned long long full[VB_STAGES];
  const int tid = threadIdx.x, cg = tid & 15, rsub = tid >> 4;
  const long long my_tiles = ntile > blockIdx.x ? (ntile - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int st = 0; st < VB_STAGES; ++st) bp::mbar_init(&full[st], 1);
    bp::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](long long it) {            // thread 0: tile `it` of this CTA -> stage it % VB_STAGES
    const int st = (int)(it % VB_STAGES);
    const long long row0 = (blockIdx.x + it * gridDim.x) * VB_ROWS;
    unsigned char* base = vb_smem + st * S::BYTES;
    bp::mbar_expect_tx(&full[st], 2 * S::H_BYTES + S::X_BYTES);
    bp::g2s(base, h1 + row0 * V_C1, S::H_BYTES, &full[st]);
    bp::g2s(base + S::H_BYTES, dh1 + row0 * V_C1, S::H_BYTES, &full[st]);
    bp::g2s(base + 2 * S::H_BYTES, x + row0 * K, S::X_BYTES, &full[st]);
  };
  if (tid == 0)
    for (long long it = 0; it < VB_STAGES && it < my_tiles; ++it) issue(it);
  float ig[4], be[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float g = gamma[4 * cg + j];
    ig[j] = fabsf(g) > 1e-20f ? 1.f / g : 0.f;
    be[j] = beta[4 * cg + j];
  }
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f}, G[4][K];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) G[j][k] = 0.f;
  for (long long it = 0; it < my_tiles; ++it) {
    const int st = (int)(it % VB_STAGES);
    bp::mbar_wait(&full[st], (unsigned)((it / VB_STAGES) & 1));
    const unsigned char* base = vb_smem + st * S::BYTES;
    const uint2* hs = reinterpret_cast<const uint2*>(base);
    const uint2* ds = reinterpret_cast<const uint2*>(base + S::H_BYTES);
    const float* xs = reinterpret_cast<const float*>(base + 2 * S::H_BYTES);
#pragma unroll
    for (int p4 = 0; p4 < VB_ROWS / 16; ++p4) {
      const int r = p4 * 16 + rsub;
      const uint2 hu = hs[r * (V_C1 / 4) + cg], du = ds[r * (V_C1 / 4) + cg];
      const float hv[4] = {__uint_as_float(hu.x << 16), __uint_as_float(hu.x & 0xffff0000u), __uint_as_float(hu.y << 16),
                           __uint_as_float(hu.y & 0xffff0000u)};
      const float dv[4] = {__uint_as_float(du.x << 16), __uint_as_float(du.x & 0xffff0000u), __uint_as_float(du.y << 16),
                           __uint_as_float(du.y & 0xffff0000u)};
      float xv[K];
#pragma unroll
      for (int k = 0; k < K; ++k) xv[k] = xs[r * K + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g = hv[j] > 0.f ? dv[j] : 0.f;
        s[j] += g;
        q[j] = fmaf(g, (hv[j] - be[j]) * ig[j], q[j]);
#pragma unroll
        for (int k = 0; k < K; ++k) G[j][k] = fmaf(g, xv[k], G[j][k]);
      }
    }
    __syncthreads();                          // every thread is done with the stage
    if (tid == 0 && it + VB_STAGES < my_tiles) {
      bp::fence_async_smem();
      issue(it + VB_STAGES);
    }
  }
  // the two rows of a warp meet by shuffle, the eight warps in shared memory (the stage ring is free now)
  float (*red)[V_C1 * W] = reinterpret_cast<float (*)[V_C1 * W]>(vb_smem);
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
    q[j] += __shfl_xor_sync(0xffffffffu, q[j], 16);
#pragma unroll
    for (int k = 0; k < K; ++k) G[j][k] += __shfl_xor_sync(0xffffffffu, G[j][k], 16);
  }
  __syncthreads();
  if (lane < 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = 4 * cg + j;
      red[wid][ch] = s[j];
      red[wid][V_C1 + ch] = q[j];
#pragma unroll
      for (int k = 0; k < K; ++k) red[wid][2 * V_C1 + ch * K + k] = G[j][k];
    }
  }
  __syncthreads();
  for (int i = tid; i < V_C1 * W; i += 256) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][i];
    partial[(long long)blockIdx.x * (V_C1 * W) + i] = t;
  }
}

// forward: h1 = relu(bn1(x W1^T)) for whole 64-row tiles; x arrives by bulk copy, the 8 KB bf16 output tile is composed in
// shared memory and leaves as ONE bulk store (two output buffers: tile i's store reads its buffer while tile i + 1 is built)
template <int K>
__global__ void __launch_bounds__(256, 4) vfe1_apply_bulk_kernel(const float* __restrict__ x, long long ntile, const float* __restrict__ W1,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 vbf16* __restrict__ h1) {
  constexpr int X_BYTES = VB_ROWS * K * 4, X_STRIDE = ((X_BYTES + 127) / 128) * 128, O_BYTES = VB_ROWS * V_C1 * 2;
  __shared__ __align__(128) unsigned char xs_raw[VB_STAGES * X_STRIDE];
  __shared__ __align__(128) unsigned char os_raw[2 * O_BYTES];
  __shared__ unsigned long long full[VB_STAGES];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long my_tiles = ntile > blockIdx.x ? (ntile - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int st = 0; st < VB_STAGES; ++st) bp::mbar_init(&full[st], 1);
    bp::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](long long it) {
    const int st = (int)(it % VB_STAGES);
    const long long row0 = (blockIdx.x + it * gridDim.x) * VB_ROWS;
    bp::mbar_expect_tx(&full[st], X_BYTES);
    bp::g2s(xs_raw + st * X_STRIDE, x + row0 * K, X_BYTES, &full[st]);
  };
  if (tid == 0)
    for (long long it = 0; it < VB_STAGES && it < my_tiles; ++it) issue(it);
  float w0[K], w1[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { w0[k] = __ldg(W1 + (2 * lane) * K + k); w1[k] = __ldg(W1 + (2 * lane + 1) * K + k); }
  const float m0 = mean[2 * lane], m1 = mean[2 * lane + 1];
  const float a0 = rstd[2 * lane] * gamma[2 * lane], a1 = rstd[2 * lane + 1] * gamma[2 * lane + 1];
  const float b0 = beta[2 * lane], b1 = beta[2 * lane + 1];
  for (long long it = 0; it < my_tiles; ++it) {
    const int st = (int)(it % VB_STAGES);
    bp::mbar_wait(&full[st], (unsigned)((it / VB_STAGES) & 1));
    const float* xs = reinterpret_cast<const float*>(xs_raw + st * X_STRIDE);
    __nv_bfloat162* os = reinterpret_cast<__nv_bfloat162*>(os_raw + (it & 1) * O_BYTES);
#pragma unroll
    for (int p8 = 0; p8 < VB_ROWS / 8; ++p8) {
      const int r = p8 * 8 + wid;
      float y0 = 0.f, y1 = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float xv = xs[r * K + k];
        y0 = fmaf(xv, w0[k], y0);
        y1 = fmaf(xv, w1[k], y1);
      }
      os[r * (V_C1 / 2) + lane] = __floats2bfloat162_rn(fmaxf(fmaf(y0 - m0, a0, b0), 0.f), fmaxf(fmaf(y1 - m1, a1, b1), 0.f));
    }
    bp::fence_async_smem();                    // the tile just written (generic proxy) is read by the bulk store
    __syncthreads();                           // tile complete, x stage free
    if (tid == 0) {
      const long long row0 = (blockIdx.x + it * gridDim.x) * VB_ROWS;
      bp::s2g(h1 + row0 * V_C1, os, O_BYTES);
      bp::s2g_commit();
      if (it + VB_STAGES < my_tiles) issue(it + VB_STAGES);
      bp::s2g_wait_read<1>();                  // the store of tile it - 1 has read its buffer: tile it + 1 may overwrite it
    }
    __syncthreads();
  }
  if (tid == 0) bp::s2g_wait_all<0>();
}

// sums the per-CTA partials (fp64) and finishes: tmp_dbeta1, tmp_dgamma1 and dW1 (closed form above)
template <int K>
__global__ void __launch_bounds__(1024) vfe1_bwd_finish_kernel(const float* __restrict__ partial, int nblocks, const double* __restrict__ moments,
                                                               const float* __restrict__ W1, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, const float* __restrict__ gamma, double count,
                                                               int accumulate, float* __restrict__ dbeta, float* __restrict__ dgamma,
                                                               float* __restrict__ dW1) {
  constexpr int W = 2 + K, NV = V_C1 * W;
  __shared__ double tot[NV];
  for (int i = threadIdx.x; i < NV; i += 1024) {
    double t = 0.0;
    int b = 0;
    for (; b + 8 <= nblocks; b += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(partial + (long long)(b + u) * NV + i);
#pragma unroll
      for (int u = 0; u < 8; ++u) t += (double)v[u];
    }
    for (; b < nblocks; ++b) t += (double)__ldg(partial + (long long)b * NV + i);
    tot[i] = t;
  }
  __syncthreads();
  if (threadIdx.x < V_C1) {
    dbeta[threadIdx.x] = (float)tot[threadIdx.x];
    dgamma[threadIdx.x] = (float)tot[V_C1 + threadIdx.x];
  }
  for (int i = threadIdx.x; i < V_C1 * K; i += 1024) {
    const int c = i / K, k = i - c * K;
    const double a = (double)rstd[c] * (double)gamma[c];
    const double c1 = a * tot[c] / count, c2 = a * tot[V_C1 + c] / count;
    // w_c . S2[:, k] with S2 stored as the upper triangle
    double ws2 = 0.0;
    for (int j = 0; j < K; ++j) {
      const int lo = j < k ? j : k, hi = j < k ? k : j;
      const int e = K + lo * K - lo * (lo - 1) / 2 + (hi - lo);
      ws2 += (double)W1[c * K + j] * moments[e];
    }
    const double val = a * tot[2 * V_C1 + i] - c1 * moments[k] - c2 * (double)rstd[c] * (ws2 - (double)mean[c] * moments[k]);
    dW1[i] = accumulate ? dW1[i] + (float)val : (float)val;
  }
}

// ---------------------------------------------------------------------------------- layer 2
// statistics of y2 (Np, C2): thread = 8 channels, 16 rows per CTA pass
template <typename T>
__global__ void __launch_bounds__(256) vfe2_stats_kernel(const T* __restrict__ y, long long Np, float* __restrict__ partial) {
  constexpr int C4 = V_C2 / 4;   // 32 lanes per row
  const int c = threadIdx.x % C4, rsub = threadIdx.x / C4, rper = 256 / C4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  for (long long row = (long long)blockIdx.x * rper + rsub; row < Np; row += (long long)gridDim.x * rper) {
    const float4 v = VT<T>::load4(y, row * C4 + c);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
  }
  __shared__ float4 red[2][256];
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  if (rsub == 0) {
    for (int j = 1; j < rper; ++j) {
      const float4 a = red[0][j * C4 + c], b = red[1][j * C4 + c];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
    }
    float* dst = partial + (long long)blockIdx.x * 2 * V_C2;
    *reinterpret_cast<float4*>(dst + 4 * c) = s;
    *reinterpret_cast<float4*>(dst + V_C2 + 4 * c) = q;
  }
}

// bn2 + ReLU + per-pillar max / argmax in one pass over y2: one warp per pillar, lane = 4 channels.
// Ties keep the lowest point index (CSR rows ascend), like the stand-alone segment max.
// SORTED: the point rows are already in pillar (CSR) order (seg_pts == NULL, pillar m owns rows [seg_off[m], seg_off[m+1])):
// no index indirection, the rows of a pillar are independent loads (four in flight) and the next pillar's offsets are
// fetched while the current one is reduced.  Unsorted rows chain seg_pts[k] -> row and are latency bound (r1: 184 us for
// 475 MB; a four-rows-in-flight variant of the unsorted loop was slower still, 236 us).
template <typename T, bool SORTED>
__global__ void __launch_bounds__(256) vfe2_apply_max_kernel(const T* __restrict__ y, const int* __restrict__ seg_off,
                                                             const int* __restrict__ seg_pts, int M, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ out,
                                                             unsigned char* __restrict__ arg) {
  const int lane = threadIdx.x & 31;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + lane), rs = __ldg(reinterpret_cast<const float4*>(rstd) + lane);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + lane), be = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  const float4 a = make_float4(rs.x * ga.x, rs.y * ga.y, rs.z * ga.z, rs.w * ga.w);
  const int stride = (gridDim.x * blockDim.x) >> 5;
  int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int s = m < M ? __ldg(seg_off + m) : 0, e = m < M ? __ldg(seg_off + m + 1) : 0;
  while (m < M) {
    const int mn = m + stride;
    const int sn = mn < M ? __ldg(seg_off + mn) : 0, en = mn < M ? __ldg(seg_off + mn + 1) : 0;
    float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
    int4 bi = make_int4(0, 0, 0, 0);
    if (SORTED) {
      for (int k = s; k < e; k += 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (k + u < e) v[u] = VT<T>::load4(y, (long long)(k + u) * (V_C2 / 4) + lane);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (k + u < e) {
            const int pnt = k + u;
            float4 w;
            w.x = fmaxf(fmaf(v[u].x - mu.x, a.x, be.x), 0.f);
            w.y = fmaxf(fmaf(v[u].y - mu.y, a.y, be.y), 0.f);
            w.z = fmaxf(fmaf(v[u].z - mu.z, a.z, be.z), 0.f);
            w.w = fmaxf(fmaf(v[u].w - mu.w, a.w, be.w), 0.f);
            if (pnt == s || w.x > best.x) { best.x = w.x; bi.x = pnt - s; }
            if (pnt == s || w.y > best.y) { best.y = w.y; bi.y = pnt - s; }
            if (pnt == s || w.z > best.z) { best.z = w.z; bi.z = pnt - s; }
            if (pnt == s || w.w > best.w) { best.w = w.w; bi.w = pnt - s; }
          }
        }
      }
    } else {
      for (int k = s; k < e; ++k) {
        const int pnt = seg_pts[k];
        float4 v = VT<T>::load4(y, (long long)pnt * (V_C2 / 4) + lane);
        v.x = fmaxf(fmaf(v.x - mu.x, a.x, be.x), 0.f);
        v.y = fmaxf(fmaf(v.y - mu.y, a.y, be.y), 0.f);
        v.z = fmaxf(fmaf(v.z - mu.z, a.z, be.z), 0.f);
        v.w = fmaxf(fmaf(v.w - mu.w, a.w, be.w), 0.f);
        if (k == s || v.x > best.x) { best.x = v.x; bi.x = k - s; }
        if (k == s || v.y > best.y) { best.y = v.y; bi.y = k - s; }
        if (k == s || v.z > best.z) { best.z = v.z; bi.z = k - s; }
        if (k == s || v.w > best.w) { best.w = v.w; bi.w = k - s; }
      }
    }
    reinterpret_cast<float4*>(out)[(long long)m * (V_C2 / 4) + lane] = best;
    // the arg-max is kept as ONE BYTE per (pillar, channel): the position inside the pillar's CSR segment, saturated at 255
    // (r1 stored the int32 row: 108-126 MB of extra writes per step, 1.21x the algorithmic bytes of this kernel).  Pillars
    // with more than 255 points are rare (Waymo: max ~70); the backward pass recomputes their arg-max instead of reading it.
    reinterpret_cast<uchar4*>(arg)[(long long)m * (V_C2 / 4) + lane] =
        make_uchar4((unsigned char)min(bi.x, 255), (unsigned char)min(bi.y, 255), (unsigned char)min(bi.z, 255), (unsigned char)min(bi.w, 255));
    m = mn; s = sn; e = en;
  }
}

// bf16 rows in pillar order (the benched configuration), r2: relu(a (y - mu) + beta) is monotone in y for a fixed channel, so
// the per-pillar maximum follows from the pillar's largest y (a >= 0) or smallest y (a < 0).  The pass therefore only keeps a
// running max and min of the RAW bf16 values with packed bf16x2 compares (exact: no arithmetic on them) - 4 instructions per row
// and lane instead of ~28 for convert + normalise + ReLU + max + arg-max per element (r2 ncu: the pass was issue-bound at 53 %
// issue-active, 3.4 TB/s) - and applies BatchNorm + ReLU once per pillar.  No arg-max is stored: the backward pass recognises
// the arg-max row as the first row whose activation equals the stored maximum (same arithmetic, bit-equal).
// Two pillars in flight per warp: while pillar m is reduced, the first nine rows of pillar m + stride are already on their
// way and the segment bounds of m + 2 stride are being fetched - the one-pillar-at-a-time form below exposes a full DRAM
// round trip per 1.4 KB segment (r2 ncu: 41 % of the DRAM rate with 40 % of the warp slots active).
__global__ void __launch_bounds__(256, 4) vfe2_apply_max_piped_kernel(const vbf16* __restrict__ y, const int* __restrict__ seg_off, int M,
                                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                      float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + lane), rs = __ldg(reinterpret_cast<const float4*>(rstd) + lane);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + lane), be = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  const float4 a = make_float4(rs.x * ga.x, rs.y * ga.y, rs.z * ga.z, rs.w * ga.w);
  const uint2* rows = reinterpret_cast<const uint2*>(y) + lane;
  const int stride = (gridDim.x * blockDim.x) >> 5;
  int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  int s = __ldg(seg_off + m), e = __ldg(seg_off + m + 1);
  int sn = 0, en = 0;
  if (m + stride < M) { sn = __ldg(seg_off + m + stride); en = __ldg(seg_off + m + stride + 1); }
  uint2 v[9];
#pragma unroll
  for (int u = 0; u < 9; ++u) v[u] = __ldg(rows + (long long)(s + u < e ? s + u : s) * (V_C2 / 4));
  while (true) {
    const int mn = m + stride, mnn = mn + stride;
    uint2 w[9];
    if (mn < M) {
#pragma unroll
      for (int u = 0; u < 9; ++u) w[u] = __ldg(rows + (long long)(sn + u < en ? sn + u : sn) * (V_C2 / 4));
    }
    int snn = 0, enn = 0;
    if (mnn < M) { snn = __ldg(seg_off + mnn); enn = __ldg(seg_off + mnn + 1); }
    __nv_bfloat162 mx0 = *reinterpret_cast<__nv_bfloat162*>(&v[0].x), mx1 = *reinterpret_cast<__nv_bfloat162*>(&v[0].y);
    __nv_bfloat162 mi0 = mx0, mi1 = mx1;
#pragma unroll
    for (int u = 1; u < 9; ++u) {
      const __nv_bfloat162 p0 = *reinterpret_cast<__nv_bfloat162*>(&v[u].x), p1 = *reinterpret_cast<__nv_bfloat162*>(&v[u].y);
      mx0 = __hmax2(mx0, p0); mx1 = __hmax2(mx1, p1);
      mi0 = __hmin2(mi0, p0); mi1 = __hmin2(mi1, p1);
    }
    for (int k = s + 9; k < e; k += 8) {          // pillars above nine points: the rest eight rows at a time
      uint2 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = __ldg(rows + (long long)(k + u < e ? k + u : s) * (V_C2 / 4));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const __nv_bfloat162 p0 = *reinterpret_cast<__nv_bfloat162*>(&t[u].x), p1 = *reinterpret_cast<__nv_bfloat162*>(&t[u].y);
        mx0 = __hmax2(mx0, p0); mx1 = __hmax2(mx1, p1);
        mi0 = __hmin2(mi0, p0); mi1 = __hmin2(mi1, p1);
      }
    }
    const float2 hx0 = __bfloat1622float2(mx0), hx1 = __bfloat1622float2(mx1), lo0 = __bfloat1622float2(mi0), lo1 = __bfloat1622float2(mi1);
    float4 best;
    best.x = fmaxf(fmaf((a.x >= 0.f ? hx0.x : lo0.x) - mu.x, a.x, be.x), 0.f);
    best.y = fmaxf(fmaf((a.y >= 0.f ? hx0.y : lo0.y) - mu.y, a.y, be.y), 0.f);
    best.z = fmaxf(fmaf((a.z >= 0.f ? hx1.x : lo1.x) - mu.z, a.z, be.z), 0.f);
    best.w = fmaxf(fmaf((a.w >= 0.f ? hx1.y : lo1.y) - mu.w, a.w, be.w), 0.f);
    __stcs(reinterpret_cast<float4*>(out) + (long long)m * (V_C2 / 4) + lane, best);
    if (mn >= M) break;
    m = mn; s = sn; e = en; sn = snn; en = enn;
#pragma unroll
    for (int u = 0; u < 9; ++u) v[u] = w[u];
  }
}

__global__ void __launch_bounds__(256) vfe2_apply_max_packed_kernel(const vbf16* __restrict__ y, const int* __restrict__ seg_off, int M,
                                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + lane), rs = __ldg(reinterpret_cast<const float4*>(rstd) + lane);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + lane), be = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  const float4 a = make_float4(rs.x * ga.x, rs.y * ga.y, rs.z * ga.z, rs.w * ga.w);
  const uint2* rows = reinterpret_cast<const uint2*>(y);
  const int stride = (gridDim.x * blockDim.x) >> 5;
  int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int s = m < M ? __ldg(seg_off + m) : 0, e = m < M ? __ldg(seg_off + m + 1) : 0;
  while (m < M) {
    const int mn = m + stride;
    const int sn = mn < M ? __ldg(seg_off + mn) : 0, en = mn < M ? __ldg(seg_off + mn + 1) : 0;
    uint2 first = __ldg(rows + (long long)s * (V_C2 / 4) + lane);
    __nv_bfloat162 mx0 = *reinterpret_cast<__nv_bfloat162*>(&first.x), mx1 = *reinterpret_cast<__nv_bfloat162*>(&first.y);
    __nv_bfloat162 mi0 = mx0, mi1 = mx1;
    for (int k = s + 1; k < e; k += 8) {
      uint2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = k + u < e ? __ldg(rows + (long long)(k + u) * (V_C2 / 4) + lane) : first;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const __nv_bfloat162 p0 = *reinterpret_cast<__nv_bfloat162*>(&v[u].x), p1 = *reinterpret_cast<__nv_bfloat162*>(&v[u].y);
        mx0 = __hmax2(mx0, p0); mx1 = __hmax2(mx1, p1);
        mi0 = __hmin2(mi0, p0); mi1 = __hmin2(mi1, p1);
      }
    }
    const float2 hx0 = __bfloat1622float2(mx0), hx1 = __bfloat1622float2(mx1), lo0 = __bfloat1622float2(mi0), lo1 = __bfloat1622float2(mi1);
    float4 best;
    best.x = fmaxf(fmaf((a.x >= 0.f ? hx0.x : lo0.x) - mu.x, a.x, be.x), 0.f);
    best.y = fmaxf(fmaf((a.y >= 0.f ? hx0.y : lo0.y) - mu.y, a.y, be.y), 0.f);
    best.z = fmaxf(fmaf((a.z >= 0.f ? hx1.x : lo1.x) - mu.z, a.z, be.z), 0.f);
    best.w = fmaxf(fmaf((a.w >= 0.f ? hx1.y : lo1.y) - mu.w, a.w, be.w), 0.f);
    reinterpret_cast<float4*>(out)[(long long)m * (V_C2 / 4) + lane] = best;
    m = mn; s = sn; e = en;
  }
}

// r2, second form of the same pass: the rows are in pillar order, so a warp can simply STREAM a run of consecutive rows and
// cut it at the pillar boundaries instead of walking one short segment (5.6 rows on average) per trip: a warp takes 8
// consecutive pillars (their offsets sit in the lanes), issues the row loads eight at a time without looking at the pillar
// structure, and flushes a pillar whenever the running row index passes its end.  No predicated-off loads, one dependent
// load round trip per 8 rows instead of per pillar.  Same arithmetic as the kernel above (running max / min of the raw bf16
// values from -inf / +inf, BatchNorm + ReLU once per pillar) - bit-equal results.
#define V_STREAM_PILLARS 8
__global__ void __launch_bounds__(256) vfe2_apply_max_stream_kernel(const vbf16* __restrict__ y, const int* __restrict__ seg_off, int M,
                                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + lane), rs = __ldg(reinterpret_cast<const float4*>(rstd) + lane);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + lane), be = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  const float4 a = make_float4(rs.x * ga.x, rs.y * ga.y, rs.z * ga.z, rs.w * ga.w);
  const uint2* rows = reinterpret_cast<const uint2*>(y);
  const int ngroups = (M + V_STREAM_PILLARS - 1) / V_STREAM_PILLARS;
  const int stride = (gridDim.x * blockDim.x) >> 5;
  const unsigned NEG_INF2 = 0xff80ff80u, POS_INF2 = 0x7f807f80u;      // bf16x2 -inf / +inf
  for (int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp < ngroups; grp += stride) {
    const int m0 = grp * V_STREAM_PILLARS;
    const int np = min(V_STREAM_PILLARS, M - m0);
    const int so = __ldg(seg_off + min(m0 + lane, M));              // lanes 0 .. np hold the offsets of the group's pillars
    const int r1 = __shfl_sync(0xffffffffu, so, np);
    int m = 0;                                                        // pillar (inside the group) the running row belongs to
    int e = __shfl_sync(0xffffffffu, so, 1);
    unsigned mx0 = NEG_INF2, mx1 = NEG_INF2, mi0 = POS_INF2, mi1 = POS_INF2;
    auto flush = [&]() {
      const float2 hx0 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&mx0)), hx1 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&mx1));
      const float2 lo0 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&mi0)), lo1 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&mi1));
      float4 best;
      best.x = fmaxf(fmaf((a.x >= 0.f ? hx0.x : lo0.x) - mu.x, a.x, be.x), 0.f);
      best.y = fmaxf(fmaf((a.y >= 0.f ? hx0.y : lo0.y) - mu.y, a.y, be.y), 0.f);
      best.z = fmaxf(fmaf((a.z >= 0.f ? hx1.x : lo1.x) - mu.z, a.z, be.z), 0.f);
      best.w = fmaxf(fmaf((a.w >= 0.f ? hx1.y : lo1.y) - mu.w, a.w, be.w), 0.f);
      reinterpret_cast<float4*>(out)[(long long)(m0 + m) * (V_C2 / 4) + lane] = best;
      mx0 = mx1 = NEG_INF2;
      mi0 = mi1 = POS_INF2;
    };
    for (int r = __shfl_sync(0xffffffffu, so, 0); r < r1; r += 8) {
      uint2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(rows + (long long)min(r + u, r1 - 1) * (V_C2 / 4) + lane);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (r + u < r1) {                                             // warp-uniform
          while (r + u >= e) {                                        // the row starts the next pillar: the finished one goes out
            flush();
            ++m;
            e = __shfl_sync(0xffffffffu, so, m + 1);
          }
          const __nv_bfloat162 p0 = *reinterpret_cast<__nv_bfloat162*>(&v[u].x), p1 = *reinterpret_cast<__nv_bfloat162*>(&v[u].y);
          __nv_bfloat162 t0 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&mx0), p0), t1 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&mx1), p1);
          __nv_bfloat162 t2 = __hmin2(*reinterpret_cast<__nv_bfloat162*>(&mi0), p0), t3 = __hmin2(*reinterpret_cast<__nv_bfloat162*>(&mi1), p1);
          mx0 = *reinterpret_cast<unsigned*>(&t0); mx1 = *reinterpret_cast<unsigned*>(&t1);
          mi0 = *reinterpret_cast<unsigned*>(&t2); mi1 = *reinterpret_cast<unsigned*>(&t3);
        }
      }
    }
    // the last pillar(s) of the group (trailing pillars without rows cannot occur: every pillar holds at least one point)
    while (m < np) {
      flush();
      ++m;
    }
  }
}

// backward sums of BN2 over the argmax entries only (every other element of d h2 is zero).  xhat at the argmax is
// recovered from the stored maximum itself: out = xhat * gamma + beta wherever out > 0 (no gather of y2).
__global__ void __launch_bounds__(256) vfe2_bwd_stats_kernel(int M, const float* __restrict__ out, const float* __restrict__ dout,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float* __restrict__ partial) {
  // thread = four channels (one float4) of the pillar rows rsub, rsub + 8, ...; four rows of `out` and `dout` in flight per
  // trip (r2: the one-channel, one-row-per-trip form with the dout load behind the o > 0 test ran at 1.5 TB/s, 158 us)
  constexpr int C = V_C2, C4 = V_C2 / 4, RPER = 256 / C4;
  const int c4 = threadIdx.x % C4, rsub = threadIdx.x / C4;
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4), be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  const float4 ig = make_float4(fabsf(ga.x) > 1e-20f ? 1.f / ga.x : 0.f, fabsf(ga.y) > 1e-20f ? 1.f / ga.y : 0.f,
                                fabsf(ga.z) > 1e-20f ? 1.f / ga.z : 0.f, fabsf(ga.w) > 1e-20f ? 1.f / ga.w : 0.f);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  const float4* o4 = reinterpret_cast<const float4*>(out) + c4;
  const float4* g4 = reinterpret_cast<const float4*>(dout) + c4;
  const long long step = (long long)gridDim.x * RPER;
  for (long long m0 = (long long)blockIdx.x * RPER + rsub; m0 < M; m0 += 4 * step) {
    float4 o[4], g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long m = m0 + u * step;
      o[u] = make_float4(0.f, 0.f, 0.f, 0.f); g[u] = o[u];
      if (m < M) { o[u] = __ldg(o4 + m * C4); g[u] = __ldg(g4 + m * C4); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (o[u].x > 0.f) { s.x += g[u].x; q.x = fmaf(g[u].x, (o[u].x - be.x) * ig.x, q.x); }
      if (o[u].y > 0.f) { s.y += g[u].y; q.y = fmaf(g[u].y, (o[u].y - be.y) * ig.y, q.y); }
      if (o[u].z > 0.f) { s.z += g[u].z; q.z = fmaf(g[u].z, (o[u].z - be.z) * ig.z, q.z); }
      if (o[u].w > 0.f) { s.w += g[u].w; q.w = fmaf(g[u].w, (o[u].w - be.w) * ig.w, q.w); }
    }
  }
  __shared__ float4 red[2][256];
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  if (rsub == 0) {
    for (int j = 1; j < RPER; ++j) {
      const float4 a = red[0][j * C4 + c4], b = red[1][j * C4 + c4];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
    }
    float* dst = partial + (long long)blockIdx.x * 2 * C;
    *reinterpret_cast<float4*>(dst + 4 * c4) = s;
    *reinterpret_cast<float4*>(dst + C + 4 * c4) = q;
  }
}

// dy2[p] = gamma rstd (g[p] - dbeta/n - xhat[p] dgamma/n), g[p, c] = dout[m, c] iff p is pillar m's argmax for c
template <typename T, bool SORTED, bool BYEQ>
__global__ void __launch_bounds__(256) vfe2_bwd_apply_kernel(const T* __restrict__ y, const int* __restrict__ seg_off,
                                                             const int* __restrict__ seg_pts, int M, const float* __restrict__ out,
                                                             const unsigned char* __restrict__ arg, const float* __restrict__ dout,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                                                             float inv_n, T* __restrict__ dy) {
  const int lane = threadIdx.x & 31;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + lane), rs = __ldg(reinterpret_cast<const float4*>(rstd) + lane);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  const float4 db = __ldg(reinterpret_cast<const float4*>(dbeta) + lane), dg = __ldg(reinterpret_cast<const float4*>(dgamma) + lane);
  const float4 a0 = make_float4(rs.x * ga.x, rs.y * ga.y, rs.z * ga.z, rs.w * ga.w);
  const float4 a1 = make_float4(a0.x * db.x * inv_n, a0.y * db.y * inv_n, a0.z * db.z * inv_n, a0.w * db.w * inv_n);
  const float4 a2 = make_float4(a0.x * dg.x * inv_n, a0.y * dg.y * inv_n, a0.z * dg.z * inv_n, a0.w * dg.w * inv_n);
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += (gridDim.x * blockDim.x) >> 5) {
    const int s = seg_off[m], e = seg_off[m + 1];
    // (r1, sorted rows: recomputing the arg-max here instead of reading it - two sweeps over the pillar's rows - cost 377 us
    // against 259 us, and a four-rows-in-flight body 259 us against this loop's ~240 us: kept simple)
    const float4 o = __ldg(reinterpret_cast<const float4*>(out) + (long long)m * (V_C2 / 4) + lane);
    float4 g = __ldg(reinterpret_cast<const float4*>(dout) + (long long)m * (V_C2 / 4) + lane);
    int4 ai = make_int4(-1, -1, -1, -1);                   // position of the arg-max row inside the pillar's segment
    if (!BYEQ) {
      const uchar4 a8 = __ldg(reinterpret_cast<const uchar4*>(arg) + (long long)m * (V_C2 / 4) + lane);
      ai = make_int4(a8.x, a8.y, a8.z, a8.w);
    }
    if (!BYEQ && e - s > 255) {
      // a byte cannot address this pillar's rows: recompute the arg-max (first maximum of relu(bn(y)), as the forward did)
      float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = s; k < e; ++k) {
        const int pnt = SORTED ? k : seg_pts[k];
        const float4 v = VT<T>::load4(y, (long long)pnt * (V_C2 / 4) + lane);
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta) + lane);
        const float wx = fmaxf(fmaf(v.x - mu.x, a0.x, b4.x), 0.f), wy = fmaxf(fmaf(v.y - mu.y, a0.y, b4.y), 0.f);
        const float wz = fmaxf(fmaf(v.z - mu.z, a0.z, b4.z), 0.f), ww = fmaxf(fmaf(v.w - mu.w, a0.w, b4.w), 0.f);
        if (k == s || wx > best.x) { best.x = wx; ai.x = k - s; }
        if (k == s || wy > best.y) { best.y = wy; ai.y = k - s; }
        if (k == s || wz > best.z) { best.z = wz; ai.z = k - s; }
        if (k == s || ww > best.w) { best.w = ww; ai.w = k - s; }
      }
    }
    g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    for (int k = s; k < e; ++k) {
      const int pnt = SORTED ? k : seg_pts[k];
      const int off = k - s;
      const float4 v = VT<T>::load4(y, (long long)pnt * (V_C2 / 4) + lane);
      if (BYEQ) {
        // the arg-max row is the first one whose activation equals the stored maximum (bit-equal: same arithmetic as forward)
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta) + lane);
        if (ai.x < 0 && fmaxf(fmaf(v.x - mu.x, a0.x, b4.x), 0.f) == o.x) ai.x = off;
        if (ai.y < 0 && fmaxf(fmaf(v.y - mu.y, a0.y, b4.y), 0.f) == o.y) ai.y = off;
        if (ai.z < 0 && fmaxf(fmaf(v.z - mu.z, a0.z, b4.z), 0.f) == o.z) ai.z = off;
        if (ai.w < 0 && fmaxf(fmaf(v.w - mu.w, a0.w, b4.w), 0.f) == o.w) ai.w = off;
      }
      float4 r;
      r.x = a0.x * (ai.x == off ? g.x : 0.f) - a1.x - (v.x - mu.x) * rs.x * a2.x;
      r.y = a0.y * (ai.y == off ? g.y : 0.f) - a1.y - (v.y - mu.y) * rs.y * a2.y;
      r.z = a0.z * (ai.z == off ? g.z : 0.f) - a1.z - (v.z - mu.z) * rs.z * a2.z;
      r.w = a0.w * (ai.w == off ? g.w : 0.f) - a1.w - (v.w - mu.w) * rs.w * a2.w;
      VT<T>::store4(dy, (long long)pnt * (V_C2 / 4) + lane, r);
    }
  }
}

__global__ void vfe_bn_grads_kernel(const float* __restrict__ db1, const float* __restrict__ dg1, const float* __restrict__ db2,
                                    const float* __restrict__ dg2, float* __restrict__ o_b1, float* __restrict__ o_g1,
                                    float* __restrict__ o_b2, float* __restrict__ o_g2, int accumulate) {
  const int i = threadIdx.x;
  if (i < V_C1) {
    o_b1[i] = accumulate ? o_b1[i] + db1[i] : db1[i];
    o_g1[i] = accumulate ? o_g1[i] + dg1[i] : dg1[i];
  }
  if (i < V_C2) {
    o_b2[i] = accumulate ? o_b2[i] + db2[i] : db2[i];
    o_g2[i] = accumulate ? o_g2[i] + dg2[i] : dg2[i];
  }
}

// ---------------------------------------------------------------------------------- host side
static int vfe_check(const gdmae_vfe_mlp_args* a) {
  GDMAE_CHECK_ARG(a && a->Np >= 0 && a->M >= 0 && a->K >= 1 && a->K <= V_KMAX && a->C1 == V_C1 && a->C2 == V_C2);
  GDMAE_CHECK_ARG(a->gemm_mode >= 0 && a->gemm_mode <= 2);
  if (a->ws_bytes < gdmae_vfe_mlp_workspace_bytes(a->K)) { gdmae_set_error("vfe_mlp: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  return GDMAE_OK;
}

extern "C" size_t gdmae_vfe_mlp_workspace_bytes(int K) {
  size_t a = (size_t)BN_PART_BLOCKS * 2 * V_C2 * 4, b = (size_t)BN_PART_BLOCKS * V_C1 * ((K > 0 ? K : 1) + 2) * 4;
  size_t c = (size_t)BN_PART_BLOCKS * 160 * 8;        // fp64 moment partials
  a = a > b ? a : b;
  return (a > c ? a : c) + 256;
}

#define VFE_CALL(expr)       \
  do {                       \
    int _rc = (expr);        \
    if (_rc) return _rc;     \
  } while (0)

extern "C" int gdmae_vfe_mlp_fwd(const gdmae_vfe_mlp_args* a) {
  VFE_CALL(vfe_check(a));
  cudaStream_t st = (cudaStream_t)a->stream;
  const long long Np = a->Np;
  const int K = a->K;
  const bool bf = a->gemm_mode == 1;
  float* partial = (float*)a->ws;
  if (Np == 0 || a->M == 0) {
    gdmae_set_error("vfe_mlp: empty batch (training-mode BatchNorm needs at least one point)");
    return GDMAE_ERR_ARG;
  }
  const long long ntile = (Np + V_TILE - 1) / V_TILE;
  const int g1 = (int)min((long long)BN_PART_BLOCKS, ntile);
  if ((K == 10 || K == 11) && a->moments) {
    // BatchNorm-1 statistics from the first two moments of x (y1 = W1 x is linear): one light pass over x, kept for backward
    const int gm = (int)min((long long)BN_PART_BLOCKS, (Np + 255) / 256);
    if (K == 10) vfe1_moments_kernel<10><<<gm, 256, 0, st>>>(a->x, Np, (double*)a->ws);
    else vfe1_moments_kernel<11><<<gm, 256, 0, st>>>(a->x, Np, (double*)a->ws);
    GDMAE_LAUNCH_CHECK();
    if (K == 10)
      vfe1_moments_finalize_kernel<10><<<1, 256, 0, st>>>((const double*)a->ws, gm, a->W1, (double)Np, a->eps, a->momentum, a->moments, a->mean1,
                                                          a->rstd1, a->running_mean1, a->running_var1);
    else
      vfe1_moments_finalize_kernel<11><<<1, 256, 0, st>>>((const double*)a->ws, gm, a->W1, (double)Np, a->eps, a->momentum, a->moments, a->mean1,
                                                          a->rstd1, a->running_mean1, a->running_var1);
    GDMAE_LAUNCH_CHECK();
  } else {
    vfe1_stats_kernel<<<g1, V_THREADS, 0, st>>>(a->x, Np, K, a->W1, partial);
    GDMAE_LAUNCH_CHECK();
    bn_finalize_kernel<<<gdmae_div_up(V_C1 * 32, 256), 256, 0, st>>>(partial, g1, V_C1, (double)Np, a->eps, a->momentum, a->mean1, a->rstd1,
                                                                    a->running_mean1, a->running_var1);
    GDMAE_LAUNCH_CHECK();
  }
  const int g1a = (int)min((long long)GDMAE_NUM_SMS * 8, ntile);
  static const bool bulk = [] { const char* e = getenv("GDMAE_VFE_BULK"); return !(e && e[0] == '0'); }();     // =0: generic form (A/B)
  const long long ntb = Np / VB_ROWS, tailb = Np - ntb * VB_ROWS;
  if (bf && bulk && ntb > 0 && (K == 10 || K == 11)) {
    // whole 64-row tiles: x by bulk copy, the output tile by one bulk store; the last Np % 64 rows by one CTA of the generic kernel
    const int gb = (int)min((long long)GDMAE_NUM_SMS * 4, ntb);
    if (K == 10) vfe1_apply_bulk_kernel<10><<<gb, 256, 0, st>>>(a->x, ntb, a->W1, a->mean1, a->rstd1, a->g1, a->b1, (vbf16*)a->h1);
    else vfe1_apply_bulk_kernel<11><<<gb, 256, 0, st>>>(a->x, ntb, a->W1, a->mean1, a->rstd1, a->g1, a->b1, (vbf16*)a->h1);
    GDMAE_LAUNCH_CHECK();
    if (tailb > 0)
      vfe1_apply_kernel<vbf16><<<1, V_THREADS, 0, st>>>(a->x + ntb * VB_ROWS * K, tailb, K, a->W1, a->mean1, a->rstd1, a->g1, a->b1,
                                                        (vbf16*)a->h1 + ntb * VB_ROWS * V_C1);
  } else if (bf) vfe1_apply_kernel<vbf16><<<g1a, V_THREADS, 0, st>>>(a->x, Np, K, a->W1, a->mean1, a->rstd1, a->g1, a->b1, (vbf16*)a->h1);
  else vfe1_apply_kernel<float><<<g1a, V_THREADS, 0, st>>>(a->x, Np, K, a->W1, a->mean1, a->rstd1, a->g1, a->b1, (float*)a->h1);
  GDMAE_LAUNCH_CHECK();
  // y2 (Np, C2) = h1 (Np, C1) W2^T, operand dtype in and out
  // layer 2 (64 -> 128): own tcgen05 / TMA GEMM in the bf16 configuration, library GEMM in the fp32 parity configurations
  if (bf) VFE_CALL(gdmae_tc_gemm(0, 1, Np, V_C2, V_C1, a->h1, V_C1, a->W2_g, V_C1, a->y2, V_C2, 1, 0.f, 0, nullptr, a->stream));
  else VFE_CALL(gdmae_gemm(0, 1, Np, V_C2, V_C1, a->h1, V_C1, a->W2_g, V_C1, a->gemm_mode, a->y2, V_C2, 0, 0.f, a->stream));
  const int g2 = (int)min((long long)BN_PART_BLOCKS, (Np + 7) / 8);
  if (bf) vfe2_stats_kernel<vbf16><<<g2, 256, 0, st>>>((const vbf16*)a->y2, Np, partial);
  else vfe2_stats_kernel<float><<<g2, 256, 0, st>>>((const float*)a->y2, Np, partial);
  GDMAE_LAUNCH_CHECK();
  bn_finalize_kernel<<<gdmae_div_up(V_C2 * 32, 256), 256, 0, st>>>(partial, g2, V_C2, (double)Np, a->eps, a->momentum, a->mean2, a->rstd2,
                                                                  a->running_mean2, a->running_var2);
  GDMAE_LAUNCH_CHECK();
  const int g3 = gdmae_grid((long long)a->M * 32, 256, 16);
  GdmaeSpan span(st);
#define VFE_APPLY_MAX(T, S)                                                                                                        \
  vfe2_apply_max_kernel<T, S><<<g3, 256, 0, st>>>((const T*)a->y2, a->seg_offsets, a->seg_points, (int)a->M, a->mean2, a->rstd2, a->g2, \
                                                  a->b2, a->out, a->argmax)
  const bool sorted = a->seg_points == nullptr;      // point rows already in pillar order
  if (sorted && bf)      // no arg-max array in this configuration: the backward pass finds the row by equality
  {
    // GDMAE_VFE_STREAM=1 selects the row-streaming form (r2 A-B in the step: 159.6 us against 146.6 us for the
    // warp-per-pillar form - the streaming form does not help, the default stays)
    static const bool stream_form = [] { const char* e = getenv("GDMAE_VFE_STREAM"); return e && e[0] == '1'; }();
    if (stream_form) {
      const int gs = gdmae_grid(((long long)a->M + V_STREAM_PILLARS - 1) / V_STREAM_PILLARS * 32, 256, 8);
      vfe2_apply_max_stream_kernel<<<gs, 256, 0, st>>>((const vbf16*)a->y2, a->seg_offsets, (int)a->M, a->mean2, a->rstd2, a->g2, a->b2, a->out);
    } else {
      // GDMAE_VFE_PIPE=0 selects the one-pillar-at-a-time form (A/B)
      static const bool piped = [] { const char* e = getenv("GDMAE_VFE_PIPE"); return !(e && e[0] == '0'); }();
      if (piped) {
        const int gp = gdmae_grid((long long)a->M * 32 / 4, 256, 4);      // >= 4 pillars per warp, 4 CTAs per SM
        vfe2_apply_max_piped_kernel<<<gp, 256, 0, st>>>((const vbf16*)a->y2, a->seg_offsets, (int)a->M, a->mean2, a->rstd2, a->g2, a->b2, a->out);
      } else
        vfe2_apply_max_packed_kernel<<<g3, 256, 0, st>>>((const vbf16*)a->y2, a->seg_offsets, (int)a->M, a->mean2, a->rstd2, a->g2, a->b2, a->out);
    }
  }
  else if (sorted) VFE_APPLY_MAX(float, true);
  else if (bf) VFE_APPLY_MAX(vbf16, false);
  else VFE_APPLY_MAX(float, false);
#undef VFE_APPLY_MAX
  GDMAE_LAUNCH_CHECK();
  // pillar scatter-max, algorithmic bytes: point rows in, segment index in, pillar rows out
  span.end(2, V_C2, Np, Np * V_C2 * (bf ? 2 : 4) + Np * 4 + a->M * V_C2 * 4);
  return GDMAE_OK;
}

extern "C" int gdmae_vfe_mlp_bwd(const gdmae_vfe_mlp_args* a) {
  VFE_CALL(vfe_check(a));
  cudaStream_t st = (cudaStream_t)a->stream;
  const long long Np = a->Np;
  const int K = a->K, M = (int)a->M;
  const bool bf = a->gemm_mode == 1;
  const int acc = a->accumulate ? 1 : 0;
  float* partial = (float*)a->ws;
  GDMAE_CHECK_ARG(Np > 0 && M > 0);
  const float inv_n = (float)(1.0 / (double)Np);
  // ---- BN2 + max: sparse sums, then one dense pass for dy2
  const int gs = (int)min((long long)BN_PART_BLOCKS, (long long)(M + 7) / 8);      // 8 pillar rows per CTA trip
  vfe2_bwd_stats_kernel<<<gs, 256, 0, st>>>(M, a->out, a->dout, a->g2, a->b2, partial);
  GDMAE_LAUNCH_CHECK();
  // this step's sums go to tmp_* (the apply passes need them alone); they reach the parameter gradients at the end
  bn_bwd_finalize_kernel<<<gdmae_div_up(V_C2, 32), 256, 0, st>>>(partial, gs, V_C2, nullptr, nullptr, a->tmp_dbeta2, a->tmp_dgamma2);
  GDMAE_LAUNCH_CHECK();
  const int g3 = gdmae_grid((long long)M * 32, 256, 16);
#define VFE_BWD_APPLY(T, S)                                                                                                        \
  vfe2_bwd_apply_kernel<T, S, (S && sizeof(T) == 2)><<<g3, 256, 0, st>>>((const T*)a->y2, a->seg_offsets, a->seg_points, M, a->out, a->argmax, a->dout, a->mean2, \
                                                  a->rstd2, a->g2, a->b2, a->tmp_dbeta2, a->tmp_dgamma2, inv_n, (T*)a->dy2)
  const bool sorted = a->seg_points == nullptr;
  if (bf) { if (sorted) VFE_BWD_APPLY(vbf16, true); else VFE_BWD_APPLY(vbf16, false); }
  else { if (sorted) VFE_BWD_APPLY(float, true); else VFE_BWD_APPLY(float, false); }
#undef VFE_BWD_APPLY
  GDMAE_LAUNCH_CHECK();
  // ---- linear 2: dW2 (C2, C1) = dy2^T h1, dh1 (Np, C1) = dy2 W2
  if (bf) {
    VFE_CALL(gdmae_tc_gemm(1, 0, V_C2, V_C1, Np, a->dy2, V_C2, a->h1, V_C1, a->d_W2, V_C1, 0, acc ? 1.f : 0.f, 1, nullptr, a->stream));
    VFE_CALL(gdmae_tc_gemm(0, 0, Np, V_C1, V_C2, a->dy2, V_C2, a->W2_g, V_C1, a->dh1, V_C1, 1, 0.f, 0, nullptr, a->stream));
  } else {
    VFE_CALL(gdmae_gemm(1, 0, V_C2, V_C1, Np, a->dy2, V_C2, a->h1, V_C1, a->gemm_mode, a->d_W2, V_C1, 0, acc ? 1.f : 0.f, a->stream));
    VFE_CALL(gdmae_gemm(0, 0, Np, V_C1, V_C2, a->dy2, V_C2, a->W2_g, V_C1, a->gemm_mode, a->dh1, V_C1, 0, 0.f, a->stream));
  }
  // ---- BN1 + linear 1
  if ((K == 10 || K == 11) && a->moments) {
    // one pass over h1, dh1, x; everything else in closed form from the moments of x saved by the forward pass
    int gp = (int)min((long long)GDMAE_NUM_SMS * 3, (Np + 15) / 16);
#define VFE_BWD_PASS(T, KK) vfe1_bwd_pass_kernel<T, KK><<<gp, 256, 0, st>>>(a->x, (const T*)a->h1, (const T*)a->dh1, Np, a->g1, a->b1, partial)
    static const bool bulk = [] { const char* e = getenv("GDMAE_VFE_BULK"); return !(e && e[0] == '0'); }();   // =0: register-staged form (A/B)
    const long long nt = Np / VB_ROWS, tail = Np - nt * VB_ROWS;
    if (bf && bulk && nt > 0) {
      // whole 64-row tiles through the bulk-copy ring, the last Np % 64 rows by one CTA of the generic kernel (one more partial row)
      const int gb = (int)min((long long)GDMAE_NUM_SMS * 2, nt);
      if (K == 10) {
        static int attr10 = cudaFuncSetAttribute(vfe1_bwd_pass_bulk_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_STAGES * VbStage<10>::BYTES);
        (void)attr10;
        vfe1_bwd_pass_bulk_kernel<10><<<gb, 256, VB_STAGES * VbStage<10>::BYTES, st>>>(a->x, (const vbf16*)a->h1, (const vbf16*)a->dh1, nt, a->g1, a->b1, partial);
      } else {
        static int attr11 = cudaFuncSetAttribute(vfe1_bwd_pass_bulk_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_STAGES * VbStage<11>::BYTES);
        (void)attr11;
        vfe1_bwd_pass_bulk_kernel<11><<<gb, 256, VB_STAGES * VbStage<11>::BYTES, st>>>(a->x, (const vbf16*)a->h1, (const vbf16*)a->dh1, nt, a->g1, a->b1, partial);
      }
      GDMAE_LAUNCH_CHECK();
      gp = gb;
      if (tail > 0) {
        const long long r0 = nt * VB_ROWS;
        float* ptail = partial + (long long)gb * V_C1 * (2 + K);
        if (K == 10) vfe1_bwd_pass_kernel<vbf16, 10><<<1, 256, 0, st>>>(a->x + r0 * K, (const vbf16*)a->h1 + r0 * V_C1, (const vbf16*)a->dh1 + r0 * V_C1, tail, a->g1, a->b1, ptail);
        else vfe1_bwd_pass_kernel<vbf16, 11><<<1, 256, 0, st>>>(a->x + r0 * K, (const vbf16*)a->h1 + r0 * V_C1, (const vbf16*)a->dh1 + r0 * V_C1, tail, a->g1, a->b1, ptail);
        gp = gb + 1;
      }
    } else if (bf) { if (K == 10) VFE_BWD_PASS(vbf16, 10); else VFE_BWD_PASS(vbf16, 11); }
    else { if (K == 10) VFE_BWD_PASS(float, 10); else VFE_BWD_PASS(float, 11); }
#undef VFE_BWD_PASS
    GDMAE_LAUNCH_CHECK();
    if (K == 10)
      vfe1_bwd_finish_kernel<10><<<1, 1024, 0, st>>>(partial, gp, a->moments, a->W1, a->mean1, a->rstd1, a->g1, (double)Np, acc, a->tmp_dbeta1,
                                                     a->tmp_dgamma1, a->d_W1);
    else
      vfe1_bwd_finish_kernel<11><<<1, 1024, 0, st>>>(partial, gp, a->moments, a->W1, a->mean1, a->rstd1, a->g1, (double)Np, acc, a->tmp_dbeta1,
                                                     a->tmp_dgamma1, a->d_W1);
    GDMAE_LAUNCH_CHECK();
  } else {
    const long long ntile = (Np + V_TILE - 1) / V_TILE;
    const int g1 = (int)min((long long)BN_PART_BLOCKS, ntile);
    if (bf) vfe1_bwd_stats_kernel<vbf16><<<g1, V_THREADS, 0, st>>>(a->x, Np, K, a->W1, a->mean1, a->rstd1, a->g1, a->b1, (const vbf16*)a->dh1, partial);
    else vfe1_bwd_stats_kernel<float><<<g1, V_THREADS, 0, st>>>(a->x, Np, K, a->W1, a->mean1, a->rstd1, a->g1, a->b1, (const float*)a->dh1, partial);
    GDMAE_LAUNCH_CHECK();
    bn_bwd_finalize_kernel<<<gdmae_div_up(V_C1, 32), 256, 0, st>>>(partial, g1, V_C1, nullptr, nullptr, a->tmp_dbeta1, a->tmp_dgamma1);
    GDMAE_LAUNCH_CHECK();
    if (!acc) GDMAE_CHECK_CUDA(cudaMemsetAsync(a->d_W1, 0, (size_t)V_C1 * K * 4, st));
    const int gw = (int)min((long long)GDMAE_NUM_SMS * 4, ntile);
    if (bf)
      vfe1_bwd_wgrad_kernel<vbf16><<<gw, V_THREADS, 0, st>>>(a->x, Np, K, a->W1, a->mean1, a->rstd1, a->g1, a->b1, (const vbf16*)a->dh1,
                                                             a->tmp_dbeta1, a->tmp_dgamma1, inv_n, a->d_W1);
    else
      vfe1_bwd_wgrad_kernel<float><<<gw, V_THREADS, 0, st>>>(a->x, Np, K, a->W1, a->mean1, a->rstd1, a->g1, a->b1, (const float*)a->dh1,
                                                             a->tmp_dbeta1, a->tmp_dgamma1, inv_n, a->d_W1);
    GDMAE_LAUNCH_CHECK();
  }
  vfe_bn_grads_kernel<<<1, 256, 0, st>>>(a->tmp_dbeta1, a->tmp_dgamma1, a->tmp_dbeta2, a->tmp_dgamma2, a->d_b1, a->d_g1, a->d_b2, a->d_g2, acc);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
