// Training-mode BatchNorm1d + ReLU over the rows of an (N, C) fp32 matrix, forward and backward,
// as two bandwidth-bound passes each (statistics, then apply) with deterministic two-stage column
// reductions (per-CTA partials in fp32, final combination in fp64).
//
// Replaces (reference file:line, relative to /root/reference):
//   norm_fn(out_channels) + nn.ReLU() of post_act_block on sparse features   pcdet/utils/spconv_utils.py:50-54
//   BatchNorm1d(eps=1e-3, momentum=0.01) + ReLU of make_fc_layers             pcdet/models/model_utils/network_utils.py:7-21
// ATen's path for the same work is batch_norm_collect_statistics + transform_input + a separate
// ReLU forward, and threshold_backward + backward_reduce + backward_elemt backward.
#include "common.cuh"
#include <stdlib.h>
#include <stdint.h>
#include <cuda_bf16.h>
#include "bn_common.cuh"
#include "bulk_pipe.cuh"


// thread owns float4 column group (threadIdx.x % C4) and rows (threadIdx.x / C4) + k * rows_per_cta
__global__ void __launch_bounds__(256) bn_stats_kernel(const float4* __restrict__ y, long long N, int C4, float* __restrict__ partial) {
  int c = threadIdx.x % C4, rsub = threadIdx.x / C4, rper = blockDim.x / C4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  for (long long row = (long long)blockIdx.x * rper + rsub; row < N; row += (long long)gridDim.x * rper) {
    float4 v = __ldg(y + row * C4 + c);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
  }
  __shared__ float4 red[2][256];
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  if (rsub == 0) {
    for (int j = 1; j < rper; ++j) {
      float4 a = red[0][j * C4 + c], b = red[1][j * C4 + c];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
    }
    float* dst = partial + (long long)blockIdx.x * 8 * C4;       // [sum(C) | sumsq(C)]
    *reinterpret_cast<float4*>(dst + 4 * c) = s;
    *reinterpret_cast<float4*>(dst + 4 * C4 + 4 * c) = q;
  }
}

// mean, rstd from the partials (fp64 combine); count = number of rows that enter the statistics
// (may exceed the rows present: zero rows of a sparse->dense map).  Updates the running buffers.
__global__ void __launch_bounds__(256) bn_relu_apply_kernel(const float4* __restrict__ y, const float4* __restrict__ mean,
                                                            const float4* __restrict__ rstd, const float4* __restrict__ gamma,
                                                            const float4* __restrict__ beta, long long n4, int C4, int relu,
                                                            float4* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C4);
    float4 v = __ldg(y + i), m = __ldg(mean + c), r = __ldg(rstd + c), g = __ldg(gamma + c), b = __ldg(beta + c);
    float4 o = make_float4((v.x - m.x) * r.x * g.x + b.x, (v.y - m.y) * r.y * g.y + b.y, (v.z - m.z) * r.z * g.z + b.z,
                           (v.w - m.w) * r.w * g.w + b.w);
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    out[i] = o;
  }
}

// backward pass 1: partial column sums of g = dout * [out > 0] and g * xhat.  The ReLU mask is recomputed from y with the
// forward's expression instead of reading the saved output (one tensor less per pass).
__global__ void __launch_bounds__(256) bn_relu_bwd_stats_kernel(const float4* __restrict__ y, const float4* __restrict__ gamma,
                                                                const float4* __restrict__ beta, const float4* __restrict__ dout,
                                                                const float4* __restrict__ mean, const float4* __restrict__ rstd,
                                                                long long N, int C4, int relu, float* __restrict__ partial) {
  int c = threadIdx.x % C4, rsub = threadIdx.x / C4, rper = blockDim.x / C4;
  float4 m = __ldg(mean + c), r = __ldg(rstd + c), ga = __ldg(gamma + c), be = __ldg(beta + c);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  for (long long row = (long long)blockIdx.x * rper + rsub; row < N; row += (long long)gridDim.x * rper) {
    float4 v = __ldg(y + row * C4 + c), g = __ldg(dout + row * C4 + c);
    if (relu) {
      float4 o = make_float4((v.x - m.x) * r.x * ga.x + be.x, (v.y - m.y) * r.y * ga.y + be.y, (v.z - m.z) * r.z * ga.z + be.z,
                             (v.w - m.w) * r.w * ga.w + be.w);
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    q.x = fmaf(g.x, (v.x - m.x) * r.x, q.x); q.y = fmaf(g.y, (v.y - m.y) * r.y, q.y);
    q.z = fmaf(g.z, (v.z - m.z) * r.z, q.z); q.w = fmaf(g.w, (v.w - m.w) * r.w, q.w);
  }
  __shared__ float4 red[2][256];
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  if (rsub == 0) {
    for (int j = 1; j < rper; ++j) {
      float4 a = red[0][j * C4 + c], b = red[1][j * C4 + c];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
    }
    float* dst = partial + (long long)blockIdx.x * 8 * C4;       // [dbeta(C) | dgamma(C)]
    *reinterpret_cast<float4*>(dst + 4 * c) = s;
    *reinterpret_cast<float4*>(dst + 4 * C4 + 4 * c) = q;
  }
}

// extra_dbeta / extra_dgamma (nullable): contributions of rows that are not materialised (the
// constant background cells of the decoder map) to the two batch sums
// backward pass 2: dy = gamma * rstd * (g - dbeta / count - xhat * dgamma / count)
__global__ void __launch_bounds__(256) bn_relu_bwd_apply_kernel(const float4* __restrict__ y, const float4* __restrict__ beta,
                                                                const float4* __restrict__ dout, const float4* __restrict__ mean,
                                                                const float4* __restrict__ rstd, const float4* __restrict__ gamma,
                                                                const float4* __restrict__ dbeta, const float4* __restrict__ dgamma,
                                                                float inv_count, long long n4, int C4, int relu,
                                                                float4* __restrict__ dy, uint2* __restrict__ dy16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C4);
    float4 v = __ldg(y + i), g = __ldg(dout + i), m = __ldg(mean + c), r = __ldg(rstd + c), ga = __ldg(gamma + c);
    float4 db = __ldg(dbeta + c), dg = __ldg(dgamma + c);
    if (relu) {
      float4 be = __ldg(beta + c);
      float4 o = make_float4((v.x - m.x) * r.x * ga.x + be.x, (v.y - m.y) * r.y * ga.y + be.y, (v.z - m.z) * r.z * ga.z + be.z,
                             (v.w - m.w) * r.w * ga.w + be.w);
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    float4 d;
    d.x = ga.x * r.x * (g.x - db.x * inv_count - (v.x - m.x) * r.x * dg.x * inv_count);
    d.y = ga.y * r.y * (g.y - db.y * inv_count - (v.y - m.y) * r.y * dg.y * inv_count);
    d.z = ga.z * r.z * (g.z - db.z * inv_count - (v.z - m.z) * r.z * dg.z * inv_count);
    d.w = ga.w * r.w * (g.w - db.w * inv_count - (v.w - m.w) * r.w * dg.w * inv_count);
    if (dy) dy[i] = d;
    if (dy16) {        // bf16 copy = the operand of the GEMMs that consume the gradient (deblock rows in the bf16 configuration)
      __nv_bfloat162 lo = __floats2bfloat162_rn(d.x, d.y), hi = __floats2bfloat162_rn(d.z, d.w);
      dy16[i] = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
    }
  }
}

extern "C" size_t gdmae_batchnorm_workspace_bytes(int C) { return (size_t)BN_PART_BLOCKS * 2 * C * 4 + 256; }

static int bn_grid(long long N, int C4, int* threads) {
  int rper = 256 / C4;
  *threads = rper * C4;
  long long need = (N + rper - 1) / rper;
  return (int)(need < BN_PART_BLOCKS ? (need < 1 ? 1 : need) : BN_PART_BLOCKS);
}

// out (N,C) = relu?( (y - mean) * rstd * gamma + beta ) with batch statistics over `count` rows (count >= N;
// rows beyond N count as zeros).  mean/rstd (C) are written for the backward pass; running buffers (may be
// NULL) are updated with momentum and the unbiased variance like nn.BatchNorm1d in training mode.
extern "C" int gdmae_batchnorm_relu_fwd(const float* y, const float* gamma, const float* beta, int64_t N, int C, double count,
                                        float eps, float momentum, int relu, float* out, float* mean, float* rstd,
                                        float* running_mean, float* running_var, void* workspace, size_t ws_bytes, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && C > 0 && (C % 4) == 0 && (C / 4) <= 256 && count >= (double)N && count >= 1.0);
  if (ws_bytes < gdmae_batchnorm_workspace_bytes(C)) { gdmae_set_error("batchnorm: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream_;
  float* partial = (float*)workspace;
  int threads, C4 = C / 4;
  int grid = bn_grid(N, C4, &threads);
  if (N == 0) {
    GDMAE_CHECK_CUDA(cudaMemsetAsync(partial, 0, (size_t)2 * C * 4, st));
    grid = 1;
  } else {
    bn_stats_kernel<<<grid, threads, 0, st>>>((const float4*)y, N, C4, partial);
    GDMAE_LAUNCH_CHECK();
  }
  bn_finalize_kernel<<<gdmae_div_up(C * 32, 256), 256, 0, st>>>(partial, grid, C, count, eps, momentum, mean, rstd, running_mean, running_var);
  GDMAE_LAUNCH_CHECK();
  if (N == 0) return GDMAE_OK;
  bn_relu_apply_kernel<<<gdmae_grid(N * C4, 256, 16), 256, 0, st>>>((const float4*)y, (const float4*)mean, (const float4*)rstd,
                                                                  (const float4*)gamma, (const float4*)beta, N * C4, C4, relu,
                                                                  (float4*)out);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// dy (N,C), dgamma (C), dbeta (C) from dout; beta (C) with gamma / mean / rstd rebuilds the ReLU mask from y (the forward
// output is not read), count as in forward.
// extra_dbeta / extra_dgamma (C, nullable) are added to the batch sums (rows that exist only implicitly).
extern "C" int gdmae_batchnorm_relu_bwd(const float* y, const float* beta, const float* dout, const float* gamma, const float* mean,
                                        const float* rstd, int64_t N, int C, double count, int relu, const float* extra_dbeta,
                                        const float* extra_dgamma, float* dy, void* dy_bf16, float* dgamma, float* dbeta,
                                        void* workspace, size_t ws_bytes, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && C > 0 && (C % 4) == 0 && (C / 4) <= 256 && count >= (double)N && count >= 1.0 && (dy || dy_bf16));
  if (ws_bytes < gdmae_batchnorm_workspace_bytes(C)) { gdmae_set_error("batchnorm: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream_;
  float* partial = (float*)workspace;
  int threads, C4 = C / 4;
  int grid = bn_grid(N, C4, &threads);
  if (N == 0) {
    GDMAE_CHECK_CUDA(cudaMemsetAsync(partial, 0, (size_t)2 * C * 4, st));
    grid = 1;
  } else {
    bn_relu_bwd_stats_kernel<<<grid, threads, 0, st>>>((const float4*)y, (const float4*)gamma, (const float4*)beta, (const float4*)dout,
                                                       (const float4*)mean, (const float4*)rstd, N, C4, relu, partial);
    GDMAE_LAUNCH_CHECK();
  }
  bn_bwd_finalize_kernel<<<gdmae_div_up(C, 32), 256, 0, st>>>(partial, grid, C, extra_dbeta, extra_dgamma, dbeta, dgamma);
  GDMAE_LAUNCH_CHECK();
  if (N == 0) return GDMAE_OK;
  bn_relu_bwd_apply_kernel<<<gdmae_grid(N * C4, 256, 16), 256, 0, st>>>(
      (const float4*)y, (const float4*)beta, (const float4*)dout, (const float4*)mean, (const float4*)rstd, (const float4*)gamma,
      (const float4*)dbeta, (const float4*)dgamma, (float)(1.0 / count), N * C4, C4, relu, (float4*)dy, (uint2*)dy_bf16);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// =====================================================================================================
// Decoder tail: BatchNorm2d (training statistics over ALL B*Y*X cells) + ReLU of the dense decoder map,
// evaluated only where the MAE head reads it - at the pillar cells - instead of materialising a second
// dense map (reference: decoder_conv_out BN + ReLU, spt_backbone_mae.py:52-57, then the gather at
// :141-143).  Forward = one read of the conv output for the statistics + a gather of M rows.
// Backward = the two BN sums over the M pillar rows only (every other cell has zero upstream gradient)
// + ONE dense pass that writes d(conv output) for all cells (the mean terms reach every cell).
// y is the cuDNN conv output, NHWC, fp32 or bf16; C % 8 == 0, 256 % (C/8) == 0.
#include <cuda_bf16.h>
template <typename T> struct Row8;
template <> struct Row8<float> {
  static __device__ __forceinline__ void load(const float* p, long long i8, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p) + 2 * i8), b = __ldg(reinterpret_cast<const float4*>(p) + 2 * i8 + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, long long i8, const float (&v)[8]) {
    __stcs(reinterpret_cast<float4*>(p) + 2 * i8, make_float4(v[0], v[1], v[2], v[3]));
    __stcs(reinterpret_cast<float4*>(p) + 2 * i8 + 1, make_float4(v[4], v[5], v[6], v[7]));
  }
};
template <> struct Row8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, long long i8, float (&v)[8]) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + i8);
    const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, long long i8, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    u.x = *reinterpret_cast<unsigned*>(&a); u.y = *reinterpret_cast<unsigned*>(&b);
    u.z = *reinterpret_cast<unsigned*>(&c); u.w = *reinterpret_cast<unsigned*>(&d);
    __stcs(reinterpret_cast<uint4*>(p) + i8, u);
  }
};

// block-level combine of per-thread 8-channel sums s, q over the rows of the CTA -> partial[block] = [sum(C) | sumsq(C)]
__device__ __forceinline__ void tail_block_reduce(float (&s)[8], float (&q)[8], int C8, float* __restrict__ partial) {
  __shared__ float red[2][8][256];
  const int c = threadIdx.x % C8, rsub = threadIdx.x / C8, rper = blockDim.x / C8;
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[0][i][threadIdx.x] = s[i]; red[1][i][threadIdx.x] = q[i]; }
  __syncthreads();
  if (rsub == 0) {
    for (int j = 1; j < rper; ++j) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += red[0][i][j * C8 + c]; q[i] += red[1][i][j * C8 + c]; }
    }
    float* dst = partial + (long long)blockIdx.x * 16 * C8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { dst[8 * c + i] = s[i]; dst[8 * C8 + 8 * c + i] = q[i]; }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) tail_dense_stats_kernel(const T* __restrict__ y, long long n_cells, int C8, float* __restrict__ partial) {
  const int c = threadIdx.x % C8, rsub = threadIdx.x / C8, rper = blockDim.x / C8;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long row = (long long)blockIdx.x * rper + rsub; row < n_cells; row += (long long)gridDim.x * rper) {
    float v[8];
    Row8<T>::load(y, row * C8 + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += v[i]; q[i] = fmaf(v[i], v[i], q[i]); }
  }
  tail_block_reduce(s, q, C8, partial);
}

template <typename T>
__global__ void __launch_bounds__(256) tail_gather_apply_kernel(const T* __restrict__ y, const long long* __restrict__ vc, long long M,
                                                                int Y, int X, int C8, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < M * C8; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / C8;
    const int c = (int)(t - m * C8);
    const long long cell = (vc[4 * m] * Y + vc[4 * m + 2]) * X + vc[4 * m + 3];
    float v[8];
    Row8<T>::load(y, cell * C8 + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = 8 * c + i;
      v[i] = fmaxf(fmaf((v[i] - __ldg(mean + ch)) * __ldg(rstd + ch), __ldg(gamma + ch), __ldg(beta + ch)), 0.f);
    }
    float4* o = reinterpret_cast<float4*>(out + m * 8 * C8 + 8 * c);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// partial sums over the pillar rows: dbeta = sum g, dgamma = sum g * xhat, g = dout where out > 0
template <typename T>
__global__ void __launch_bounds__(256) tail_sparse_bwd_stats_kernel(const T* __restrict__ y, const long long* __restrict__ vc, long long M,
                                                                    int Y, int X, int C8, const float* __restrict__ out,
                                                                    const float* __restrict__ dout, const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd, float* __restrict__ partial) {
  const int c = threadIdx.x % C8, rsub = threadIdx.x / C8, rper = blockDim.x / C8;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float mu[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { mu[i] = __ldg(mean + 8 * c + i); rs[i] = __ldg(rstd + 8 * c + i); }
  for (long long m = (long long)blockIdx.x * rper + rsub; m < M; m += (long long)gridDim.x * rper) {
    const long long cell = (vc[4 * m] * Y + vc[4 * m + 2]) * X + vc[4 * m + 3];
    float v[8], o[8], g[8];
    Row8<T>::load(y, cell * C8 + c, v);
    Row8<float>::load(out, m * C8 + c, o);
    Row8<float>::load(dout, m * C8 + c, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gi = o[i] > 0.f ? g[i] : 0.f;
      s[i] += gi;
      q[i] = fmaf(gi, (v[i] - mu[i]) * rs[i], q[i]);
    }
  }
  tail_block_reduce(s, q, C8, partial);
}

// dy[cell] = gamma * rstd * (g[cell] - dbeta / n - xhat[cell] * dgamma / n) for every cell; g = 0 off the pillars
template <typename T>
__global__ void __launch_bounds__(256) tail_dense_bwd_apply_kernel(const T* __restrict__ y, const int* __restrict__ cell2pillar,
                                                                   long long n_cells, int C8, const float* __restrict__ out,
                                                                   const float* __restrict__ dout, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                   const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                                                                   float inv_n, T* __restrict__ dy) {
  const int c = threadIdx.x % C8, rsub = threadIdx.x / C8, rper = blockDim.x / C8;
  float mu[8], rs[8], a0[8], a1[8], a2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = 8 * c + i;
    mu[i] = __ldg(mean + ch);
    rs[i] = __ldg(rstd + ch);
    a0[i] = __ldg(gamma + ch) * rs[i];                 // dy = a0 * g - a1 - xhat * a2
    a1[i] = a0[i] * __ldg(dbeta + ch) * inv_n;
    a2[i] = a0[i] * __ldg(dgamma + ch) * inv_n;
  }
  for (long long cell = (long long)blockIdx.x * rper + rsub; cell < n_cells; cell += (long long)gridDim.x * rper) {
    float v[8], r[8];
    Row8<T>::load(y, cell * C8 + c, v);
    const int m = __ldg(cell2pillar + cell);
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = -a1[i] - (v[i] - mu[i]) * rs[i] * a2[i];
    if (m >= 0) {
      float o[8], g[8];
      Row8<float>::load(out, (long long)m * C8 + c, o);
      Row8<float>::load(dout, (long long)m * C8 + c, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = fmaf(a0[i], o[i] > 0.f ? g[i] : 0.f, r[i]);
    }
    Row8<T>::store(dy, cell * C8 + c, r);
  }
}

// The bf16, C = 128 case of tail_dense_bwd_apply_kernel with the dense map staged by bulk copies: a tile = 32 consecutive
// cells of y (8 KB) + their cell -> pillar indices (128 B) arrives in a ring of TDB_STAGES stages, the 8 KB tile of dy is
// composed in shared memory and leaves as one bulk store; the rows of out / dout of the ~13 % cells that carry a pillar are
// plain loads.  r2: the register-staged kernel moved its 900 MB at 3.2 TB/s (283 us).
#define TDB_ROWS 32
#define TDB_STAGES 4
#define TDB_C 128
#define TDB_TILE_BYTES (TDB_ROWS * TDB_C * 2)
#define TDB_STAGE_BYTES (TDB_TILE_BYTES + 128)
#define TDB_SMEM (TDB_STAGES * TDB_STAGE_BYTES + 2 * TDB_TILE_BYTES)
__global__ void __launch_bounds__(256, 3) tail_dense_bwd_apply_bulk_kernel(const __nv_bfloat16* __restrict__ y, const int* __restrict__ cell2pillar,
                                                                           long long ntile, const float* __restrict__ out,
                                                                           const float* __restrict__ dout, const float* __restrict__ mean,
                                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                           const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                                                                           float inv_n, __nv_bfloat16* __restrict__ dy) {
  extern __shared__ __align__(128) unsigned char tdb_smem[];
  __shared__ unsigned long long full[TDB_STAGES];
  constexpr int C8 = TDB_C / 8;
  const int tid = threadIdx.x, c = tid % C8, rsub = tid / C8;
  const long long my_tiles = ntile > blockIdx.x ? (ntile - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int st = 0; st < TDB_STAGES; ++st) bp::mbar_init(&full[st], 1);
    bp::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](long long it) {
    const int st = (int)(it % TDB_STAGES);
    const long long cell0 = (blockIdx.x + it * gridDim.x) * TDB_ROWS;
    unsigned char* base = tdb_smem + st * TDB_STAGE_BYTES;
    bp::mbar_expect_tx(&full[st], TDB_TILE_BYTES + TDB_ROWS * 4);
    bp::g2s(base, y + cell0 * TDB_C, TDB_TILE_BYTES, &full[st]);
    bp::g2s(base + TDB_TILE_BYTES, cell2pillar + cell0, TDB_ROWS * 4, &full[st]);
  };
  if (tid == 0)
    for (long long it = 0; it < TDB_STAGES && it < my_tiles; ++it) issue(it);
  float mu[8], a0[8], a1[8], c2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = 8 * c + i;
    const float rs = __ldg(rstd + ch);
    mu[i] = __ldg(mean + ch);
    a0[i] = __ldg(gamma + ch) * rs;                    // dy = a0 * g - a1 - (y - mean) * c2,  c2 = rstd * a0 * dgamma / n
    a1[i] = a0[i] * __ldg(dbeta + ch) * inv_n;
    c2[i] = rs * (a0[i] * __ldg(dgamma + ch) * inv_n);
  }
  unsigned char* obuf = tdb_smem + TDB_STAGES * TDB_STAGE_BYTES;
  for (long long it = 0; it < my_tiles; ++it) {
    const int st = (int)(it % TDB_STAGES);
    bp::mbar_wait(&full[st], (unsigned)((it / TDB_STAGES) & 1));
    const uint4* ys = reinterpret_cast<const uint4*>(tdb_smem + st * TDB_STAGE_BYTES);
    const int* ms = reinterpret_cast<const int*>(tdb_smem + st * TDB_STAGE_BYTES + TDB_TILE_BYTES);
    uint4* os = reinterpret_cast<uint4*>(obuf + (it & 1) * TDB_TILE_BYTES);
#pragma unroll
    for (int ps = 0; ps < TDB_ROWS / 16; ++ps) {
      const int row = ps * 16 + rsub;
      const uint4 u = ys[row * C8 + c];
      const int m = ms[row];
      const unsigned w[4] = {u.x, u.y, u.z, u.w};
      float r[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        r[2 * i] = -a1[2 * i] - (__uint_as_float(w[i] << 16) - mu[2 * i]) * c2[2 * i];
        r[2 * i + 1] = -a1[2 * i + 1] - (__uint_as_float(w[i] & 0xffff0000u) - mu[2 * i + 1]) * c2[2 * i + 1];
      }
      if (m >= 0) {
        float o[8], g[8];
        Row8<float>::load(out, (long long)m * C8 + c, o);
        Row8<float>::load(dout, (long long)m * C8 + c, g);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = fmaf(a0[i], o[i] > 0.f ? g[i] : 0.f, r[i]);
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(r[0], r[1]), p1 = __floats2bfloat162_rn(r[2], r[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(r[4], r[5]), p3 = __floats2bfloat162_rn(r[6], r[7]);
      os[row * C8 + c] = make_uint4(*reinterpret_cast<unsigned*>(&p0), *reinterpret_cast<unsigned*>(&p1), *reinterpret_cast<unsigned*>(&p2),
                                    *reinterpret_cast<unsigned*>(&p3));
    }
    bp::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      const long long cell0 = (blockIdx.x + it * gridDim.x) * TDB_ROWS;
      bp::s2g(dy + cell0 * TDB_C, os, TDB_TILE_BYTES);
      bp::s2g_commit();
      if (it + TDB_STAGES < my_tiles) issue(it + TDB_STAGES);
      bp::s2g_wait_read<1>();
    }
    __syncthreads();
  }
  if (tid == 0) bp::s2g_wait_all<0>();
}

static int tail_check(int B, int Y, int X, int C, int dtype, size_t ws_bytes) {
  GDMAE_CHECK_ARG(B >= 1 && Y >= 1 && X >= 1 && C > 0 && (C % 8) == 0 && 256 % (C / 8) == 0 && (dtype == 0 || dtype == 1));
  if (ws_bytes < gdmae_batchnorm_workspace_bytes(C)) { gdmae_set_error("decoder tail: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  return GDMAE_OK;
}

// out (M,C) fp32 = relu(BN(y))[pillar cells]; mean / rstd (C) kept for backward; running buffers updated (nullable).
extern "C" int gdmae_decoder_tail_fwd(const void* y, int dtype, int B, int Y, int X, int C, const int64_t* voxel_coords, int64_t M,
                                      const float* gamma, const float* beta, float eps, float momentum, float* out, float* mean,
                                      float* rstd, float* running_mean, float* running_var, void* workspace, size_t ws_bytes,
                                      void* stream_) {
  int rc = tail_check(B, Y, X, C, dtype, ws_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream_;
  float* partial = (float*)workspace;
  const long long n_cells = (long long)B * Y * X;
  const int C8 = C / 8, rper = 256 / C8;
  const int grid = (int)min((long long)BN_PART_BLOCKS, (n_cells + rper - 1) / rper);
  if (dtype == 0) tail_dense_stats_kernel<float><<<grid, 256, 0, st>>>((const float*)y, n_cells, C8, partial);
  else tail_dense_stats_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, n_cells, C8, partial);
  GDMAE_LAUNCH_CHECK();
  bn_finalize_kernel<<<gdmae_div_up(C * 32, 256), 256, 0, st>>>(partial, grid, C, (double)n_cells, eps, momentum, mean, rstd, running_mean,
                                                               running_var);
  GDMAE_LAUNCH_CHECK();
  if (M == 0) return GDMAE_OK;
  const int g2 = gdmae_grid(M * C8, 256, 16);
  if (dtype == 0)
    tail_gather_apply_kernel<float><<<g2, 256, 0, st>>>((const float*)y, (const long long*)voxel_coords, M, Y, X, C8, mean, rstd, gamma,
                                                        beta, out);
  else
    tail_gather_apply_kernel<__nv_bfloat16><<<g2, 256, 0, st>>>((const __nv_bfloat16*)y, (const long long*)voxel_coords, M, Y, X, C8,
                                                                mean, rstd, gamma, beta, out);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// dy (B,Y,X,C) in y's dtype, dgamma / dbeta (C).  cell2pillar (B*Y*X): pillar row of every cell or -1 (gdmae_dynvox).
extern "C" int gdmae_decoder_tail_bwd(const void* y, int dtype, int B, int Y, int X, int C, const int64_t* voxel_coords,
                                      const int32_t* cell2pillar, int64_t M, const float* out, const float* dout, const float* gamma,
                                      const float* mean, const float* rstd, void* dy, float* dgamma, float* dbeta, void* workspace,
                                      size_t ws_bytes, void* stream_) {
  int rc = tail_check(B, Y, X, C, dtype, ws_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream_;
  float* partial = (float*)workspace;
  const long long n_cells = (long long)B * Y * X;
  const int C8 = C / 8, rper = 256 / C8;
  int grid = (int)min((long long)BN_PART_BLOCKS, (long long)((M + rper - 1) / rper));
  if (M == 0) {
    GDMAE_CHECK_CUDA(cudaMemsetAsync(partial, 0, (size_t)2 * C * 4, st));
    grid = 1;
  } else {
    if (dtype == 0)
      tail_sparse_bwd_stats_kernel<float><<<grid, 256, 0, st>>>((const float*)y, (const long long*)voxel_coords, M, Y, X, C8, out, dout,
                                                                mean, rstd, partial);
    else
      tail_sparse_bwd_stats_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, (const long long*)voxel_coords, M, Y, X,
                                                                        C8, out, dout, mean, rstd, partial);
    GDMAE_LAUNCH_CHECK();
  }
  bn_bwd_finalize_kernel<<<gdmae_div_up(C, 32), 256, 0, st>>>(partial, grid, C, nullptr, nullptr, dbeta, dgamma);
  GDMAE_LAUNCH_CHECK();
  const int g2 = (int)min((long long)GDMAE_NUM_SMS * 16, (n_cells + rper - 1) / rper);
  const float inv_n = (float)(1.0 / (double)n_cells);
  if (dtype == 0)
    tail_dense_bwd_apply_kernel<float><<<g2, 256, 0, st>>>((const float*)y, cell2pillar, n_cells, C8, out, dout, mean, rstd, gamma, dbeta,
                                                           dgamma, inv_n, (float*)dy);
  else {
    static const bool bulk_on = [] { const char* e = getenv("GDMAE_BN_BULK"); return !(e && e[0] == '0'); }();      // =0: generic kernel (A/B)
    const long long nt = n_cells / TDB_ROWS, tail = n_cells - nt * TDB_ROWS;
    if (bulk_on && C == TDB_C && nt > 0 && (((uintptr_t)y | (uintptr_t)dy | (uintptr_t)cell2pillar) & 15) == 0) {
      static int attr_t = cudaFuncSetAttribute(tail_dense_bwd_apply_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TDB_SMEM);
      (void)attr_t;
      const int gb = (int)min((long long)GDMAE_NUM_SMS * 3, nt);
      tail_dense_bwd_apply_bulk_kernel<<<gb, 256, TDB_SMEM, st>>>((const __nv_bfloat16*)y, cell2pillar, nt, out, dout, mean, rstd, gamma, dbeta,
                                                                  dgamma, inv_n, (__nv_bfloat16*)dy);
      GDMAE_LAUNCH_CHECK();
      if (tail > 0)
        tail_dense_bwd_apply_kernel<__nv_bfloat16><<<1, 256, 0, st>>>((const __nv_bfloat16*)y + nt * TDB_ROWS * C, cell2pillar + nt * TDB_ROWS, tail,
                                                                      C8, out, dout, mean, rstd, gamma, dbeta, dgamma, inv_n,
                                                                      (__nv_bfloat16*)dy + nt * TDB_ROWS * C);
    } else {
      tail_dense_bwd_apply_kernel<__nv_bfloat16><<<g2, 256, 0, st>>>((const __nv_bfloat16*)y, cell2pillar, n_cells, C8, out, dout, mean, rstd,
                                                                     gamma, dbeta, dgamma, inv_n, (__nv_bfloat16*)dy);
    }
  }
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// =====================================================================================================
// Typed BatchNorm + ReLU over (N, C) rows: y / out / dout / dy each fp32 or bf16 (dtype 0 / 1), 8-channel packets
// (one 16-byte access per thread in bf16, two in fp32), per-thread channel constants in registers (every thread keeps its
// channel packet for the whole launch: the trip length is a multiple of C8).  Used by the decoder deblocks of the bf16
// configuration (spt_backbone_mae.py:31-44): their output rows only feed the bf16 dense map and their upstream gradient IS
// bf16 (rows gathered from d(map)), so keeping both as bf16 loses nothing and halves four passes.
template <typename TY, typename TO>
__global__ void __launch_bounds__(256) bn8_relu_apply_kernel(const TY* __restrict__ y, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, long long n8, int C8, int relu,
                                                             TO* __restrict__ out) {
  const int c = threadIdx.x % C8;      // blockDim.x = 256 is a multiple of C8
  float m[8], r[8], g[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = __ldg(mean + 8 * c + i); r[i] = __ldg(rstd + 8 * c + i); g[i] = __ldg(gamma + 8 * c + i); b[i] = __ldg(beta + 8 * c + i); }
  const long long step = (long long)gridDim.x * blockDim.x;
#pragma unroll 2
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += step) {
    float v[8];
    Row8<TY>::load(y, i, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = (v[k] - m[k]) * r[k] * g[k] + b[k];
      if (relu) v[k] = fmaxf(v[k], 0.f);
    }
    Row8<TO>::store(out, i, v);
  }
}

template <typename TY, typename TD>
__global__ void __launch_bounds__(256) bn8_relu_bwd_stats_kernel(const TY* __restrict__ y, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, const TD* __restrict__ dout,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 long long N, int C8, int relu, float* __restrict__ partial) {
  const int c = threadIdx.x % C8, rsub = threadIdx.x / C8, rper = blockDim.x / C8;
  float m[8], r[8], ga[8], be[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = __ldg(mean + 8 * c + i); r[i] = __ldg(rstd + 8 * c + i); ga[i] = __ldg(gamma + 8 * c + i); be[i] = __ldg(beta + 8 * c + i); }
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
  for (long long row = (long long)blockIdx.x * rper + rsub; row < N; row += (long long)gridDim.x * rper) {
    float v[8], d[8];
    Row8<TY>::load(y, row * C8 + c, v);
    Row8<TD>::load(dout, row * C8 + c, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (v[k] - m[k]) * r[k];
      if (relu && !((v[k] - m[k]) * r[k] * ga[k] + be[k] > 0.f)) d[k] = 0.f;
      s[k] += d[k];
      q[k] = fmaf(d[k], xh, q[k]);
    }
  }
  tail_block_reduce(s, q, C8, partial);        // [dbeta(C) | dgamma(C)] per CTA
}

template <typename TY, typename TD, typename TG>
__global__ void __launch_bounds__(256) bn8_relu_bwd_apply_kernel(const TY* __restrict__ y, const float* __restrict__ beta,
                                                                 const TD* __restrict__ dout, const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                 const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                                                                 float inv_count, long long n8, int C8, int relu, TG* __restrict__ dy) {
  const int c = threadIdx.x % C8;
  float m[8], r[8], ga[8], be[8], db[8], dg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    m[i] = __ldg(mean + 8 * c + i); r[i] = __ldg(rstd + 8 * c + i); ga[i] = __ldg(gamma + 8 * c + i); be[i] = __ldg(beta + 8 * c + i);
    db[i] = __ldg(dbeta + 8 * c + i) * inv_count; dg[i] = __ldg(dgamma + 8 * c + i);
  }
  const long long step = (long long)gridDim.x * blockDim.x;
#pragma unroll 2
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += step) {
    float v[8], d[8];
    Row8<TY>::load(y, i, v);
    Row8<TD>::load(dout, i, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (v[k] - m[k]) * r[k];
      if (relu && !((v[k] - m[k]) * r[k] * ga[k] + be[k] > 0.f)) d[k] = 0.f;
      d[k] = ga[k] * r[k] * (d[k] - db[k] - xh * dg[k] * inv_count);
    }
    Row8<TG>::store(dy, i, d);
  }
}

// ---- the bf16 / bf16 / bf16, C = 128 case of the two backward passes with their rows staged by bulk copies -----------------
// (the decoder deblocks of the bf16 configuration: 135 M elements per tensor and step).  A tile = 32 consecutive rows of y
// and of dout (8 KB each, contiguous), brought by one elected thread into a ring of BNB_STAGES stages (csrc/bulk_pipe.cuh);
// the apply pass composes its bf16 output tile in shared memory and sends it back with one bulk store.  Only whole tiles;
// the host runs the generic kernels above on the last N % 32 rows.  r2: the register-staged kernels ran at 3.0 (statistics)
// and 3.7 TB/s (apply).
#define BNB_STAGES 4
#define BNB_TILE_BYTES 8192                  // one tile of y: 32 rows of 128 channels or 16 rows of 256, bf16
template <typename TD, int C> struct BnbCfg {
  static constexpr int ROWS = BNB_TILE_BYTES / (C * 2);
  static constexpr int D_BYTES = ROWS * C * (int)sizeof(TD);
  static constexpr int STAGE = BNB_TILE_BYTES + D_BYTES;
  static constexpr int SMEM_STATS = BNB_STAGES * STAGE;
  static constexpr int SMEM_APPLY = BNB_STAGES * STAGE + 2 * BNB_TILE_BYTES;
};

__device__ __forceinline__ void bnb_unpack(const uint4 u, float (&v)[8]) {
  const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <typename TD> __device__ __forceinline__ void bnb_load_d(const unsigned char* tile, int i8, float (&d)[8]);
template <> __device__ __forceinline__ void bnb_load_d<__nv_bfloat16>(const unsigned char* tile, int i8, float (&d)[8]) {
  bnb_unpack(reinterpret_cast<const uint4*>(tile)[i8], d);
}
template <> __device__ __forceinline__ void bnb_load_d<float>(const unsigned char* tile, int i8, float (&d)[8]) {
  const float4 a = reinterpret_cast<const float4*>(tile)[2 * i8], b = reinterpret_cast<const float4*>(tile)[2 * i8 + 1];
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}

// y bf16, dout TD (bf16: deblock rows gathered from d(map); fp32: sparse-conv outputs feeding the fp32 residual stream), dy bf16
template <bool APPLY, typename TD, int C>
__global__ void __launch_bounds__(256, 2) bn8_relu_bwd_bulk_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, const TD* __restrict__ dout,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                   const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                                                                   float inv_count, long long ntile, int relu, float* __restrict__ partial,
                                                                   __nv_bfloat16* __restrict__ dy) {
  using Cfg = BnbCfg<TD, C>;
  extern __shared__ __align__(128) unsigned char bnb_smem[];
  __shared__ unsigned long long full[BNB_STAGES];
  constexpr int C8 = C / 8, RPP = 256 / C8;                          // rows per pass
  const int tid = threadIdx.x, c = tid % C8, rsub = tid / C8;
  const long long my_tiles = ntile > blockIdx.x ? (ntile - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int st = 0; st < BNB_STAGES; ++st) bp::mbar_init(&full[st], 1);
    bp::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](long long it) {
    const int st = (int)(it % BNB_STAGES);
    const long long row0 = (blockIdx.x + it * gridDim.x) * Cfg::ROWS;
    unsigned char* base = bnb_smem + st * Cfg::STAGE;
    bp::mbar_expect_tx(&full[st], BNB_TILE_BYTES + Cfg::D_BYTES);
    bp::g2s(base, y + row0 * C, BNB_TILE_BYTES, &full[st]);
    bp::g2s(base + BNB_TILE_BYTES, dout + row0 * C, Cfg::D_BYTES, &full[st]);
  };
  if (tid == 0)
    for (long long it = 0; it < BNB_STAGES && it < my_tiles; ++it) issue(it);
  float m[8], r[8], ga[8], be[8], db[8], dg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    m[i] = __ldg(mean + 8 * c + i); r[i] = __ldg(rstd + 8 * c + i); ga[i] = __ldg(gamma + 8 * c + i); be[i] = __ldg(beta + 8 * c + i);
    db[i] = APPLY ? __ldg(dbeta + 8 * c + i) * inv_count : 0.f;
    dg[i] = APPLY ? __ldg(dgamma + 8 * c + i) : 0.f;
  }
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  unsigned char* obuf = bnb_smem + BNB_STAGES * Cfg::STAGE;           // APPLY: two output tiles
  for (long long it = 0; it < my_tiles; ++it) {
    const int st = (int)(it % BNB_STAGES);
    bp::mbar_wait(&full[st], (unsigned)((it / BNB_STAGES) & 1));
    const unsigned char* base = bnb_smem + st * Cfg::STAGE;
    const uint4* ys = reinterpret_cast<const uint4*>(base);
    uint4* os = reinterpret_cast<uint4*>(obuf + (it & 1) * BNB_TILE_BYTES);
#pragma unroll
    for (int ps = 0; ps < Cfg::ROWS / RPP; ++ps) {
      const int row = ps * RPP + rsub;
      float v[8], d[8];
      bnb_unpack(ys[row * C8 + c], v);
      bnb_load_d<TD>(base + BNB_TILE_BYTES, row * C8 + c, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (v[k] - m[k]) * r[k];
        if (relu && !((v[k] - m[k]) * r[k] * ga[k] + be[k] > 0.f)) d[k] = 0.f;
        if (APPLY) d[k] = ga[k] * r[k] * (d[k] - db[k] - xh * dg[k] * inv_count);
        else { s[k] += d[k]; q[k] = fmaf(d[k], xh, q[k]); }
      }
      if (APPLY) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(d[0], d[1]), p1 = __floats2bfloat162_rn(d[2], d[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(d[4], d[5]), p3 = __floats2bfloat162_rn(d[6], d[7]);
        os[row * C8 + c] = make_uint4(*reinterpret_cast<unsigned*>(&p0), *reinterpret_cast<unsigned*>(&p1),
                                      *reinterpret_cast<unsigned*>(&p2), *reinterpret_cast<unsigned*>(&p3));
      }
    }
    if (APPLY) bp::fence_async_smem();         // the output tile (generic proxy) is read by the bulk store
    __syncthreads();                           // stage consumed / output tile complete
    if (tid == 0) {
      if (APPLY) {
        const long long row0 = (blockIdx.x + it * gridDim.x) * Cfg::ROWS;
        bp::s2g(dy + row0 * C, os, BNB_TILE_BYTES);
        bp::s2g_commit();
      }
      if (it + BNB_STAGES < my_tiles) {
        bp::fence_async_smem();
        issue(it + BNB_STAGES);
      }
      if (APPLY) bp::s2g_wait_read<1>();       // the store of tile it - 1 has read its buffer: tile it + 1 may overwrite it
    }
    if (APPLY) __syncthreads();
  }
  if (APPLY) {
    if (tid == 0) bp::s2g_wait_all<0>();
  } else {
    tail_block_reduce(s, q, C8, partial);      // [dbeta(C) | dgamma(C)] per CTA
  }
}

// forward apply on the same ring: y bf16 in, out TO (fp32: sparse-conv outputs, bf16: deblock rows) composed in shared memory
template <typename TO, int C>
__global__ void __launch_bounds__(256, 3) bn8_relu_apply_bulk_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, long long ntile, int relu,
                                                                     TO* __restrict__ out) {
  constexpr int ROWS = BNB_TILE_BYTES / (C * 2), O_BYTES = ROWS * C * (int)sizeof(TO);
  extern __shared__ __align__(128) unsigned char bnb_smem[];        // BNB_STAGES y tiles | two output tiles
  __shared__ unsigned long long full[BNB_STAGES];
  constexpr int C8 = C / 8, RPP = 256 / C8;
  const int tid = threadIdx.x, c = tid % C8, rsub = tid / C8;
  const long long my_tiles = ntile > blockIdx.x ? (ntile - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int st = 0; st < BNB_STAGES; ++st) bp::mbar_init(&full[st], 1);
    bp::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](long long it) {
    const int st = (int)(it % BNB_STAGES);
    const long long row0 = (blockIdx.x + it * gridDim.x) * ROWS;
    bp::mbar_expect_tx(&full[st], BNB_TILE_BYTES);
    bp::g2s(bnb_smem + st * BNB_TILE_BYTES, y + row0 * C, BNB_TILE_BYTES, &full[st]);
  };
  if (tid == 0)
    for (long long it = 0; it < BNB_STAGES && it < my_tiles; ++it) issue(it);
  float m[8], r[8], g[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = __ldg(mean + 8 * c + i); r[i] = __ldg(rstd + 8 * c + i); g[i] = __ldg(gamma + 8 * c + i); b[i] = __ldg(beta + 8 * c + i); }
  unsigned char* obuf = bnb_smem + BNB_STAGES * BNB_TILE_BYTES;
  for (long long it = 0; it < my_tiles; ++it) {
    const int st = (int)(it % BNB_STAGES);
    bp::mbar_wait(&full[st], (unsigned)((it / BNB_STAGES) & 1));
    const uint4* ys = reinterpret_cast<const uint4*>(bnb_smem + st * BNB_TILE_BYTES);
    unsigned char* os = obuf + (it & 1) * O_BYTES;
#pragma unroll
    for (int ps = 0; ps < ROWS / RPP; ++ps) {
      const int row = ps * RPP + rsub;
      float v[8];
      bnb_unpack(ys[row * C8 + c], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] = (v[k] - m[k]) * r[k] * g[k] + b[k];
        if (relu) v[k] = fmaxf(v[k], 0.f);
      }
      if (sizeof(TO) == 2) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
        reinterpret_cast<uint4*>(os)[row * C8 + c] = make_uint4(*reinterpret_cast<unsigned*>(&p0), *reinterpret_cast<unsigned*>(&p1),
                                                                *reinterpret_cast<unsigned*>(&p2), *reinterpret_cast<unsigned*>(&p3));
      } else {
        reinterpret_cast<float4*>(os)[2 * (row * C8 + c)] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(os)[2 * (row * C8 + c) + 1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    bp::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      const long long row0 = (blockIdx.x + it * gridDim.x) * ROWS;
      bp::s2g(out + row0 * C, os, O_BYTES);
      bp::s2g_commit();
      if (it + BNB_STAGES < my_tiles) issue(it + BNB_STAGES);
      bp::s2g_wait_read<1>();
    }
    __syncthreads();
  }
  if (tid == 0) bp::s2g_wait_all<0>();
}
template <typename TO, int C>
static long long bnb_apply_launch(const void* y, const float* mean, const float* rstd, const float* gamma, const float* beta, long long N, int relu,
                                  void* out, cudaStream_t st) {
  constexpr int ROWS = BNB_TILE_BYTES / (C * 2), SMEM = BNB_STAGES * BNB_TILE_BYTES + 2 * ROWS * C * (int)sizeof(TO);
  const long long nt = N / ROWS;
  if (nt == 0) return 0;
  static int attr = cudaFuncSetAttribute(bn8_relu_apply_bulk_kernel<TO, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  (void)attr;
  const int gb = (int)min((long long)GDMAE_NUM_SMS * 3, nt);
  bn8_relu_apply_bulk_kernel<TO, C><<<gb, 256, SMEM, st>>>((const __nv_bfloat16*)y, mean, rstd, gamma, beta, nt, relu, (TO*)out);
  return nt * ROWS;
}

// launches the bulk form over the whole tiles and returns the rows it covered (0: not applicable, the caller runs the generic form)
template <bool APPLY, typename TD, int C>
static long long bnb_launch(const void* y, const float* gamma, const float* beta, const void* dout, const float* mean, const float* rstd,
                            const float* dbeta, const float* dgamma, float inv, long long N, int relu, float* partial, void* dy, int* nblocks,
                            cudaStream_t st) {
  using Cfg = BnbCfg<TD, C>;
  const long long nt = N / Cfg::ROWS;
  if (nt == 0) return 0;
  constexpr int SMEM = APPLY ? Cfg::SMEM_APPLY : Cfg::SMEM_STATS;
  static int attr = cudaFuncSetAttribute(bn8_relu_bwd_bulk_kernel<APPLY, TD, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  (void)attr;
  const int gb = (int)min((long long)GDMAE_NUM_SMS * 2, nt);
  bn8_relu_bwd_bulk_kernel<APPLY, TD, C><<<gb, 256, SMEM, st>>>((const __nv_bfloat16*)y, gamma, beta, (const TD*)dout, mean, rstd, dbeta, dgamma, inv,
                                                                nt, relu, partial, (__nv_bfloat16*)dy);
  if (nblocks) *nblocks = gb;
  return nt * Cfg::ROWS;
}
template <bool APPLY>
static long long bnb_dispatch(int dout_dtype, int C, const void* y, const float* gamma, const float* beta, const void* dout, const float* mean,
                              const float* rstd, const float* dbeta, const float* dgamma, float inv, long long N, int relu, float* partial,
                              void* dy, int* nblocks, cudaStream_t st) {
  if (C == 128 && dout_dtype == 1) return bnb_launch<APPLY, __nv_bfloat16, 128>(y, gamma, beta, dout, mean, rstd, dbeta, dgamma, inv, N, relu, partial, dy, nblocks, st);
  if (C == 128 && dout_dtype == 0) return bnb_launch<APPLY, float, 128>(y, gamma, beta, dout, mean, rstd, dbeta, dgamma, inv, N, relu, partial, dy, nblocks, st);
  if (C == 256 && dout_dtype == 1) return bnb_launch<APPLY, __nv_bfloat16, 256>(y, gamma, beta, dout, mean, rstd, dbeta, dgamma, inv, N, relu, partial, dy, nblocks, st);
  if (C == 256 && dout_dtype == 0) return bnb_launch<APPLY, float, 256>(y, gamma, beta, dout, mean, rstd, dbeta, dgamma, inv, N, relu, partial, dy, nblocks, st);
  return 0;
}

#define BN8_DISPATCH2(A, B, CALL)                                                  \
  do {                                                                             \
    if ((A) == 0 && (B) == 0) { using T0 = float; using T1 = float; CALL; }        \
    else if ((A) == 0) { using T0 = float; using T1 = __nv_bfloat16; CALL; }       \
    else if ((B) == 0) { using T0 = __nv_bfloat16; using T1 = float; CALL; }       \
    else { using T0 = __nv_bfloat16; using T1 = __nv_bfloat16; CALL; }             \
  } while (0)

// as gdmae_batchnorm_relu_fwd, y and out typed (0 = fp32, 1 = bf16); C % 8 == 0 and 256 % (C / 8) == 0
extern "C" int gdmae_batchnorm_relu_fwd_t(const void* y, int y_dtype, const float* gamma, const float* beta, int64_t N, int C, double count,
                                          float eps, float momentum, int relu, void* out, int out_dtype, float* mean, float* rstd,
                                          float* running_mean, float* running_var, void* workspace, size_t ws_bytes, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && C > 0 && (C % 8) == 0 && 256 % (C / 8) == 0 && count >= (double)N && count >= 1.0);
  GDMAE_CHECK_ARG((y_dtype == 0 || y_dtype == 1) && (out_dtype == 0 || out_dtype == 1));
  if (ws_bytes < gdmae_batchnorm_workspace_bytes(C)) { gdmae_set_error("batchnorm: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream_;
  float* partial = (float*)workspace;
  const int C8 = C / 8, rper = 256 / C8;
  int grid = (int)min((long long)BN_PART_BLOCKS, (long long)((N + rper - 1) / rper));
  if (N == 0) {
    GDMAE_CHECK_CUDA(cudaMemsetAsync(partial, 0, (size_t)2 * C * 4, st));
    grid = 1;
  } else {
    if (y_dtype == 0) tail_dense_stats_kernel<float><<<grid, 256, 0, st>>>((const float*)y, N, C8, partial);
    else tail_dense_stats_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, N, C8, partial);
    GDMAE_LAUNCH_CHECK();
  }
  bn_finalize_kernel<<<gdmae_div_up(C * 32, 256), 256, 0, st>>>(partial, grid, C, count, eps, momentum, mean, rstd, running_mean, running_var);
  GDMAE_LAUNCH_CHECK();
  if (N == 0) return GDMAE_OK;
  static const bool bulk_on = [] { const char* e = getenv("GDMAE_BN_BULK"); return !(e && e[0] == '0'); }();      // =0: generic kernel (A/B)
  long long done = 0;
  if (bulk_on && y_dtype == 1 && (((uintptr_t)y | (uintptr_t)out) & 15) == 0) {
    if (C == 128 && out_dtype == 1) done = bnb_apply_launch<__nv_bfloat16, 128>(y, mean, rstd, gamma, beta, N, relu, out, st);
    else if (C == 128) done = bnb_apply_launch<float, 128>(y, mean, rstd, gamma, beta, N, relu, out, st);
    else if (C == 256 && out_dtype == 1) done = bnb_apply_launch<__nv_bfloat16, 256>(y, mean, rstd, gamma, beta, N, relu, out, st);
    else if (C == 256) done = bnb_apply_launch<float, 256>(y, mean, rstd, gamma, beta, N, relu, out, st);
    GDMAE_LAUNCH_CHECK();
  }
  if (N > done) {
    const long long n8 = (N - done) * C8;
    const int g2 = gdmae_grid(n8, 256, 16);
    const size_t yo = (size_t)done * C * (y_dtype ? 2 : 4), oo = (size_t)done * C * (out_dtype ? 2 : 4);
    BN8_DISPATCH2(y_dtype, out_dtype,
                  (bn8_relu_apply_kernel<T0, T1><<<g2, 256, 0, st>>>((const T0*)((const char*)y + yo), mean, rstd, gamma, beta, n8, C8, relu,
                                                                     (T1*)((char*)out + oo))));
    GDMAE_LAUNCH_CHECK();
  }
  return GDMAE_OK;
}

// as gdmae_batchnorm_relu_bwd, y / dout / dy typed (0 = fp32, 1 = bf16)
extern "C" int gdmae_batchnorm_relu_bwd_t(const void* y, int y_dtype, const float* beta, const void* dout, int dout_dtype, const float* gamma,
                                          const float* mean, const float* rstd, int64_t N, int C, double count, int relu,
                                          const float* extra_dbeta, const float* extra_dgamma, void* dy, int dy_dtype, float* dgamma,
                                          float* dbeta, void* workspace, size_t ws_bytes, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && C > 0 && (C % 8) == 0 && 256 % (C / 8) == 0 && count >= (double)N && count >= 1.0 && dy);
  GDMAE_CHECK_ARG((y_dtype == 0 || y_dtype == 1) && (dout_dtype == 0 || dout_dtype == 1) && (dy_dtype == 0 || dy_dtype == 1));
  if (ws_bytes < gdmae_batchnorm_workspace_bytes(C)) { gdmae_set_error("batchnorm: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream_;
  float* partial = (float*)workspace;
  const int C8 = C / 8, rper = 256 / C8;
  int grid = (int)min((long long)BN_PART_BLOCKS, (long long)((N + rper - 1) / rper));
  static const bool bulk_on = [] { const char* e = getenv("GDMAE_BN_BULK"); return !(e && e[0] == '0'); }();      // =0: generic kernels (A/B)
  const bool bulk = bulk_on && y_dtype == 1 && dy_dtype == 1 && (((uintptr_t)y | (uintptr_t)dout | (uintptr_t)dy) & 15) == 0;
  long long done = 0;                    // rows the bulk form covered (whole tiles); the generic kernels take the rest
  if (N == 0) {
    GDMAE_CHECK_CUDA(cudaMemsetAsync(partial, 0, (size_t)2 * C * 4, st));
    grid = 1;
  } else {
    int gb = 0;
    if (bulk) done = bnb_dispatch<false>(dout_dtype, C, y, gamma, beta, dout, mean, rstd, nullptr, nullptr, 0.f, N, relu, partial, nullptr, &gb, st);
    GDMAE_LAUNCH_CHECK();
    if (done > 0) {
      grid = gb;
      if (N > done) {                    // one CTA of the generic kernel on the last rows, one more partial row
        const size_t yo = (size_t)done * C * 2, dd = (size_t)done * C * (dout_dtype ? 2 : 4);
        BN8_DISPATCH2(1, dout_dtype,
                      (bn8_relu_bwd_stats_kernel<T0, T1><<<1, 256, 0, st>>>((const T0*)((const char*)y + yo), gamma, beta, (const T1*)((const char*)dout + dd),
                                                                            mean, rstd, N - done, C8, relu, partial + (long long)gb * 2 * C)));
        grid = gb + 1;
      }
    } else {
      BN8_DISPATCH2(y_dtype, dout_dtype,
                    (bn8_relu_bwd_stats_kernel<T0, T1><<<grid, 256, 0, st>>>((const T0*)y, gamma, beta, (const T1*)dout, mean, rstd, N, C8, relu, partial)));
    }
    GDMAE_LAUNCH_CHECK();
  }
  bn_bwd_finalize_kernel<<<gdmae_div_up(C, 32), 256, 0, st>>>(partial, grid, C, extra_dbeta, extra_dgamma, dbeta, dgamma);
  GDMAE_LAUNCH_CHECK();
  if (N == 0) return GDMAE_OK;
  const long long n8 = N * C8;
  const int g2 = gdmae_grid(n8, 256, 16);
  const float inv = (float)(1.0 / count);
  if (done > 0) {
    bnb_dispatch<true>(dout_dtype, C, y, gamma, beta, dout, mean, rstd, dbeta, dgamma, inv, N, relu, nullptr, dy, nullptr, st);
    GDMAE_LAUNCH_CHECK();
    if (N > done) {
      const size_t yo = (size_t)done * C * 2, dd = (size_t)done * C * (dout_dtype ? 2 : 4);
      const long long t8 = (N - done) * C8;
      BN8_DISPATCH2(1, dout_dtype,
                    (bn8_relu_bwd_apply_kernel<T0, T1, __nv_bfloat16><<<gdmae_grid(t8, 256, 16), 256, 0, st>>>(
                        (const T0*)((const char*)y + yo), beta, (const T1*)((const char*)dout + dd), mean, rstd, gamma, dbeta, dgamma, inv, t8, C8, relu,
                        (__nv_bfloat16*)((char*)dy + yo))));
    }
  } else if (dy_dtype == 0)
    BN8_DISPATCH2(y_dtype, dout_dtype,
                  (bn8_relu_bwd_apply_kernel<T0, T1, float><<<g2, 256, 0, st>>>((const T0*)y, beta, (const T1*)dout, mean, rstd, gamma, dbeta, dgamma,
                                                                                inv, n8, C8, relu, (float*)dy)));
  else
    BN8_DISPATCH2(y_dtype, dout_dtype,
                  (bn8_relu_bwd_apply_kernel<T0, T1, __nv_bfloat16><<<g2, 256, 0, st>>>((const T0*)y, beta, (const T1*)dout, mean, rstd, gamma, dbeta,
                                                                                        dgamma, inv, n8, C8, relu, (__nv_bfloat16*)dy)));
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
