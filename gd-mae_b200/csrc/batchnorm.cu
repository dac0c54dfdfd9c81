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

#define BN_PART_BLOCKS (GDMAE_NUM_SMS * 4)

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
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, double count, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    s += (double)partial[(long long)b * 2 * C + c];
    q += (double)partial[(long long)b * 2 * C + C + c];
  }
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

// backward pass 1: partial column sums of g = dout * [out > 0] and g * xhat
__global__ void __launch_bounds__(256) bn_relu_bwd_stats_kernel(const float4* __restrict__ y, const float4* __restrict__ out,
                                                                const float4* __restrict__ dout, const float4* __restrict__ mean,
                                                                const float4* __restrict__ rstd, long long N, int C4, int relu,
                                                                float* __restrict__ partial) {
  int c = threadIdx.x % C4, rsub = threadIdx.x / C4, rper = blockDim.x / C4;
  float4 m = __ldg(mean + c), r = __ldg(rstd + c);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  for (long long row = (long long)blockIdx.x * rper + rsub; row < N; row += (long long)gridDim.x * rper) {
    float4 v = __ldg(y + row * C4 + c), g = __ldg(dout + row * C4 + c);
    if (relu) {
      float4 o = __ldg(out + row * C4 + c);
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
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, const float* __restrict__ extra_dbeta,
                                       const float* __restrict__ extra_dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ dgamma) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = extra_dbeta ? (double)extra_dbeta[c] : 0.0, q = extra_dgamma ? (double)extra_dgamma[c] : 0.0;
  for (int b = 0; b < nblocks; ++b) {
    s += (double)partial[(long long)b * 2 * C + c];
    q += (double)partial[(long long)b * 2 * C + C + c];
  }
  dbeta[c] = (float)s;
  dgamma[c] = (float)q;
}

// backward pass 2: dy = gamma * rstd * (g - dbeta / count - xhat * dgamma / count)
__global__ void __launch_bounds__(256) bn_relu_bwd_apply_kernel(const float4* __restrict__ y, const float4* __restrict__ out,
                                                                const float4* __restrict__ dout, const float4* __restrict__ mean,
                                                                const float4* __restrict__ rstd, const float4* __restrict__ gamma,
                                                                const float4* __restrict__ dbeta, const float4* __restrict__ dgamma,
                                                                float inv_count, long long n4, int C4, int relu,
                                                                float4* __restrict__ dy) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C4);
    float4 v = __ldg(y + i), g = __ldg(dout + i), m = __ldg(mean + c), r = __ldg(rstd + c), ga = __ldg(gamma + c);
    float4 db = __ldg(dbeta + c), dg = __ldg(dgamma + c);
    if (relu) {
      float4 o = __ldg(out + i);
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    float4 d;
    d.x = ga.x * r.x * (g.x - db.x * inv_count - (v.x - m.x) * r.x * dg.x * inv_count);
    d.y = ga.y * r.y * (g.y - db.y * inv_count - (v.y - m.y) * r.y * dg.y * inv_count);
    d.z = ga.z * r.z * (g.z - db.z * inv_count - (v.z - m.z) * r.z * dg.z * inv_count);
    d.w = ga.w * r.w * (g.w - db.w * inv_count - (v.w - m.w) * r.w * dg.w * inv_count);
    dy[i] = d;
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
  bn_finalize_kernel<<<gdmae_div_up(C, 128), 128, 0, st>>>(partial, grid, C, count, eps, momentum, mean, rstd, running_mean, running_var);
  GDMAE_LAUNCH_CHECK();
  if (N == 0) return GDMAE_OK;
  bn_relu_apply_kernel<<<gdmae_grid(N * C4, 256, 16), 256, 0, st>>>((const float4*)y, (const float4*)mean, (const float4*)rstd,
                                                                  (const float4*)gamma, (const float4*)beta, N * C4, C4, relu,
                                                                  (float4*)out);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// dy (N,C), dgamma (C), dbeta (C) from dout; `out` is the forward output (ReLU mask), count as in forward.
// extra_dbeta / extra_dgamma (C, nullable) are added to the batch sums (rows that exist only implicitly).
extern "C" int gdmae_batchnorm_relu_bwd(const float* y, const float* out, const float* dout, const float* gamma, const float* mean,
                                        const float* rstd, int64_t N, int C, double count, int relu, const float* extra_dbeta,
                                        const float* extra_dgamma, float* dy, float* dgamma, float* dbeta, void* workspace,
                                        size_t ws_bytes, void* stream_) {
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
    bn_relu_bwd_stats_kernel<<<grid, threads, 0, st>>>((const float4*)y, (const float4*)out, (const float4*)dout, (const float4*)mean,
                                                       (const float4*)rstd, N, C4, relu, partial);
    GDMAE_LAUNCH_CHECK();
  }
  bn_bwd_finalize_kernel<<<gdmae_div_up(C, 128), 128, 0, st>>>(partial, grid, C, extra_dbeta, extra_dgamma, dbeta, dgamma);
  GDMAE_LAUNCH_CHECK();
  if (N == 0) return GDMAE_OK;
  bn_relu_bwd_apply_kernel<<<gdmae_grid(N * C4, 256, 16), 256, 0, st>>>(
      (const float4*)y, (const float4*)out, (const float4*)dout, (const float4*)mean, (const float4*)rstd, (const float4*)gamma,
      (const float4*)dbeta, (const float4*)dgamma, (float)(1.0 / count), N * C4, C4, relu, (float4*)dy);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
