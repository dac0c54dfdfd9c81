// Fused row-wise kernels of the SRA encoder layer: residual (+bias) + LayerNorm (fwd/bwd), bias +
// exact GELU (fwd/bwd), column sums (bias gradients) and the positional gather-add.  All HBM-bound,
// 128-bit accesses, one warp per token row, parameter gradients reduced deterministically in two
// stages (per-CTA partials, then a fixed-order final sum) instead of float atomics.  Every kernel
// can additionally (or instead) emit its result as bf16: those copies are the operands of the
// cuBLAS GEMMs in the bf16 configuration, so no separate cast pass exists.
//
// Replaces (reference file:line, relative to /root/reference):
//   src = norm1(src + src2); src = norm2(src + src2)        pcdet/models/model_utils/sst_basic_block.py:79-83
//   activation(linear1(src)) with activation = F.gelu (erf)   pcdet/models/model_utils/sst_basic_block.py:81,117-125
//   q = k = feat_3d + pos                                     pcdet/models/model_utils/sst_basic_block.py:39-46
#include "common.cuh"
#include <cuda_bf16.h>

#define EW_PART_BLOCKS (GDMAE_NUM_SMS * 2)

// Second stage of the column reductions, fused into the producing kernel: every CTA adds its column sums to the
// output with fp32 atomics (RED.ADD) - about 300 adds per column, spread over the kernel's tail.  The summation order
// is not fixed (last-bit run-to-run differences in bias / LayerNorm gradients, like the reference's atomics-based
// scatter ops); a single-CTA fixed-order tail was measured 20-30 us slower per call.  accumulate = 0: the host
// wrapper zeroes the output first.
__device__ __forceinline__ void ew_cta_atomic_add(const float* __restrict__ row, int ncols, float* __restrict__ out0, int n0,
                                                  float* __restrict__ out1, float* __restrict__ out2 = nullptr) {
  __syncthreads();
  for (int c = threadIdx.x; c < ncols; c += blockDim.x)
    atomicAdd(c < n0 ? out0 + c : (c < 2 * n0 ? out1 + (c - n0) : out2 + (c - 2 * n0)), row[c]);
}

__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* p, long long i4, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<unsigned int*>(&a);
  u.y = *reinterpret_cast<unsigned int*>(&b);
  reinterpret_cast<uint2*>(p)[i4] = u;
}
__device__ __forceinline__ float4 load_bf16x4(const __nv_bfloat16* p, long long i4) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p) + i4);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// ------------------------------------------------------------------ y = LayerNorm(x + res + bias)
// VEC = d / 128 float4 per lane (d = 128 -> 1, d = 256 -> 2); bias, y_bf16 nullable
template <int VEC, bool RESBF>
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(const float4* __restrict__ x, const void* __restrict__ res,
                                                         const float4* __restrict__ bias, const float4* __restrict__ gamma,
                                                         const float4* __restrict__ beta, long long N, float eps,
                                                         float4* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16,
                                                         float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  constexpr int D = VEC * 128;
  int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 g[VEC], b[VEC], bi[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    g[v] = __ldg(gamma + v * 32 + lane);
    b[v] = __ldg(beta + v * 32 + lane);
    bi[v] = bias ? __ldg(bias + v * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = warp; row < N; row += nwarps) {
    float4 z[VEC];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float4 a = __ldg(x + row * (D / 4) + v * 32 + lane);
      float4 r = RESBF ? load_bf16x4((const __nv_bfloat16*)res, row * (D / 4) + v * 32 + lane)
                       : __ldg((const float4*)res + row * (D / 4) + v * 32 + lane);
      z[v] = make_float4(a.x + r.x + bi[v].x, a.y + r.y + bi[v].y, a.z + r.z + bi[v].z, a.w + r.w + bi[v].w);
      s += z[v].x + z[v].y + z[v].z + z[v].w;
    }
    float mean = warp_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float dx = z[v].x - mean, dy = z[v].y - mean, dz = z[v].z - mean, dw = z[v].w - mean;
      q += dx * dx + dy * dy + dz * dz + dw * dw;
    }
    float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float4 o;
      o.x = (z[v].x - mean) * rstd * g[v].x + b[v].x;
      o.y = (z[v].y - mean) * rstd * g[v].y + b[v].y;
      o.z = (z[v].z - mean) * rstd * g[v].z + b[v].z;
      o.w = (z[v].w - mean) * rstd * g[v].w + b[v].w;
      y[row * (D / 4) + v * 32 + lane] = o;
      if (y_bf16) store_bf16x4(y_bf16, row * (D / 4) + v * 32 + lane, o);
    }
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  }
}

// dz = rstd * (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat));  partial dgamma/dbeta per CTA
template <int VEC, bool RESBF, bool DBIAS>
__global__ void __launch_bounds__(256) add_ln_bwd_kernel(const float4* __restrict__ x, const void* __restrict__ res,
                                                         const float4* __restrict__ bias, const float4* __restrict__ gamma,
                                                         const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                         const float4* __restrict__ dy, long long N, float4* __restrict__ dz,
                                                         __nv_bfloat16* __restrict__ dz_bf16, float* __restrict__ partial,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                         float* __restrict__ dbias, int accumulate) {
  constexpr int D = VEC * 128;
  constexpr int NACC = DBIAS ? 3 : 2;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 g[VEC], dg[VEC], db[VEC], bi[VEC], ds[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    g[v] = __ldg(gamma + v * 32 + lane);
    bi[v] = bias ? __ldg(bias + v * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    dg[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    ds[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = warp; row < N; row += nwarps) {
    float mean = mean_in[row], rstd = rstd_in[row];
    float4 xh[VEC], dxh[VEC];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float4 a = __ldg(x + row * (D / 4) + v * 32 + lane);
      float4 r = RESBF ? load_bf16x4((const __nv_bfloat16*)res, row * (D / 4) + v * 32 + lane)
                       : __ldg((const float4*)res + row * (D / 4) + v * 32 + lane);
      float4 d = __ldg(dy + row * (D / 4) + v * 32 + lane);
      xh[v] = make_float4((a.x + r.x + bi[v].x - mean) * rstd, (a.y + r.y + bi[v].y - mean) * rstd,
                          (a.z + r.z + bi[v].z - mean) * rstd, (a.w + r.w + bi[v].w - mean) * rstd);
      dxh[v] = make_float4(d.x * g[v].x, d.y * g[v].y, d.z * g[v].z, d.w * g[v].w);
      c1 += dxh[v].x + dxh[v].y + dxh[v].z + dxh[v].w;
      c2 += dxh[v].x * xh[v].x + dxh[v].y * xh[v].y + dxh[v].z * xh[v].z + dxh[v].w * xh[v].w;
      dg[v].x += d.x * xh[v].x; dg[v].y += d.y * xh[v].y; dg[v].z += d.z * xh[v].z; dg[v].w += d.w * xh[v].w;
      db[v].x += d.x; db[v].y += d.y; db[v].z += d.z; db[v].w += d.w;
    }
    c1 = warp_sum(c1) * (1.f / D);
    c2 = warp_sum(c2) * (1.f / D);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float4 o;
      o.x = rstd * (dxh[v].x - c1 - xh[v].x * c2);
      o.y = rstd * (dxh[v].y - c1 - xh[v].y * c2);
      o.z = rstd * (dxh[v].z - c1 - xh[v].z * c2);
      o.w = rstd * (dxh[v].w - c1 - xh[v].w * c2);
      dz[row * (D / 4) + v * 32 + lane] = o;
      if (dz_bf16) store_bf16x4(dz_bf16, row * (D / 4) + v * 32 + lane, o);
      if (DBIAS) { ds[v].x += o.x; ds[v].y += o.y; ds[v].z += o.z; ds[v].w += o.w; }
    }
  }
  // CTA reduction over its 8 warps (fixed order), then one partial row [dgamma(D) | dbeta(D) | dbias(D)] per CTA
  __shared__ float4 red[8][NACC * VEC][32];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    red[wid][v][lane] = dg[v];
    red[wid][VEC + v][lane] = db[v];
    if (DBIAS) red[wid][2 * VEC + v][lane] = ds[v];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int v = 0; v < NACC * VEC; ++v) {
      float4 acc = red[0][v][lane];
      for (int w = 1; w < 8; ++w) {
        float4 t = red[w][v][lane];
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      // v < VEC: dgamma columns [128 v + 4 lane ..]; then dbeta; then dbias
      int colbase = (v / VEC) * D + (v % VEC) * 128 + 4 * lane;
      *reinterpret_cast<float4*>(partial + (long long)blockIdx.x * NACC * D + colbase) = acc;
    }
  }
  ew_cta_atomic_add(partial + (long long)blockIdx.x * NACC * D, NACC * D, dgamma, D, dbeta, dbias);
}

extern "C" size_t gdmae_rowwise_workspace_bytes(int max_cols) { return (size_t)EW_PART_BLOCKS * 2 * max_cols * 4 + 256; }

static int ew_grid(long long rows) {
  long long need = (rows + 7) / 8;
  return (int)(need < EW_PART_BLOCKS ? (need < 1 ? 1 : need) : EW_PART_BLOCKS);
}

// y = LayerNorm(x + res + bias) * gamma + beta over rows of d in {128, 256}; bias (d) and y_bf16 (N,d) may be
// NULL; mean/rstd (N) saved for backward.  Internal form: res may be a bf16 tensor (a GEMM output of the bf16
// configuration).
int ew_add_layernorm_fwd(const float* x, const void* res, int res_bf16, const float* bias, const float* gamma, const float* beta,
                         int64_t N, int d, float eps, float* y, void* y_bf16, float* mean, float* rstd, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && (d == 128 || d == 256));
  if (N == 0) return GDMAE_OK;
  cudaStream_t st = (cudaStream_t)stream_;
  int grid = gdmae_grid(N * 32, 256, 8);
#define EW_LN_FWD(V, RB)                                                                                                        \
  add_ln_fwd_kernel<V, RB><<<grid, 256, 0, st>>>((const float4*)x, res, (const float4*)bias, (const float4*)gamma,              \
                                                 (const float4*)beta, N, eps, (float4*)y, (__nv_bfloat16*)y_bf16, mean, rstd)
  if (d == 128) { if (res_bf16) EW_LN_FWD(1, true); else EW_LN_FWD(1, false); }
  else { if (res_bf16) EW_LN_FWD(2, true); else EW_LN_FWD(2, false); }
#undef EW_LN_FWD
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

extern "C" int gdmae_add_layernorm_fwd(const float* x, const float* res, const float* bias, const float* gamma, const float* beta,
                                       int64_t N, int d, float eps, float* y, void* y_bf16, float* mean, float* rstd,
                                       void* stream_) {
  return ew_add_layernorm_fwd(x, res, 0, bias, gamma, beta, N, d, eps, y, y_bf16, mean, rstd, stream_);
}

// dz (N,d) = gradient w.r.t. (x + res + bias) (also emitted as bf16 when dz_bf16 != NULL); dgamma/dbeta (d)
// written (accumulate=0) or added to (accumulate=1).  Internal form: res may be bf16; dbias (d, nullable) receives
// the column sums of dz - the gradient of the bias inside the LayerNorm argument - from the same pass.
int ew_add_layernorm_bwd(const float* x, const void* res, int res_bf16, const float* bias, const float* gamma, const float* mean,
                         const float* rstd, const float* dy, int64_t N, int d, float* dz, void* dz_bf16, float* dgamma,
                         float* dbeta, float* dbias, int accumulate, void* workspace, size_t ws_bytes, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && (d == 128 || d == 256));
  if (ws_bytes < gdmae_rowwise_workspace_bytes(dbias ? 2 * d : d)) { gdmae_set_error("add_layernorm_bwd: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream_;
  if (!accumulate) {
    GDMAE_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)d * 4, st));
    GDMAE_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)d * 4, st));
    if (dbias) GDMAE_CHECK_CUDA(cudaMemsetAsync(dbias, 0, (size_t)d * 4, st));
  }
  if (N == 0) return GDMAE_OK;
  float* partial = (float*)workspace;
  int grid = ew_grid(N);
#define EW_LN_BWD(V, RB, DB)                                                                                                     \
  add_ln_bwd_kernel<V, RB, DB><<<grid, 256, 0, st>>>((const float4*)x, res, (const float4*)bias, (const float4*)gamma, mean, rstd, \
                                                     (const float4*)dy, N, (float4*)dz, (__nv_bfloat16*)dz_bf16, partial, dgamma, \
                                                     dbeta, dbias, accumulate)
#define EW_LN_BWD_V(V)                                          \
  do {                                                          \
    if (res_bf16) { if (dbias) EW_LN_BWD(V, true, true); else EW_LN_BWD(V, true, false); }   \
    else { if (dbias) EW_LN_BWD(V, false, true); else EW_LN_BWD(V, false, false); }          \
  } while (0)
  if (d == 128) EW_LN_BWD_V(1);
  else EW_LN_BWD_V(2);
#undef EW_LN_BWD_V
#undef EW_LN_BWD
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

extern "C" int gdmae_add_layernorm_bwd(const float* x, const float* res, const float* bias, const float* gamma, const float* mean,
                                       const float* rstd, const float* dy, int64_t N, int d, float* dz, void* dz_bf16,
                                       float* dgamma, float* dbeta, int accumulate, void* workspace, size_t ws_bytes,
                                       void* stream_) {
  return ew_add_layernorm_bwd(x, res, 0, bias, gamma, mean, rstd, dy, N, d, dz, dz_bf16, dgamma, dbeta, nullptr, accumulate, workspace,
                              ws_bytes, stream_);
}

// ------------------------------------------------------------------ g = gelu_erf(h + b)
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * __expf(-0.5f * x * x);
}

// bf16 configuration only: Phi(v) and phi(v) from ONE exp and ONE reciprocal (Abramowitz & Stegun 7.1.26, |erf error| <=
// 1.5e-7 - far below the bf16 rounding of the inputs and outputs there).  The erff-based passes were instruction bound
// (ncu: 0.19 / 0.26 of the HBM rate, issue active 71 % / 52 %).  The fp32 parity configuration keeps erff.
__device__ __forceinline__ void gelu_terms_fast(float v, float& cdf, float& pdf) {
  const float z = fabsf(v) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  const float e = __expf(-0.5f * v * v);                       // exp(-z^2)
  const float poly = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
  const float erf_abs = fmaf(-poly, e, 1.f);                   // erf(|z|)
  cdf = 0.5f * (1.f + copysignf(erf_abs, v));
  pdf = 0.39894228040143267794f * e;
}
template <bool FAST>
__device__ __forceinline__ float gelu_t(float v) {
  if (!FAST) return gelu_f(v);
  float cdf, pdf;
  gelu_terms_fast(v, cdf, pdf);
  return v * cdf;
}
template <bool FAST>
__device__ __forceinline__ float gelu_grad_t(float v) {
  if (!FAST) return gelu_grad_f(v);
  float cdf, pdf;
  gelu_terms_fast(v, cdf, pdf);
  return fmaf(v, pdf, cdf);
}

template <bool HBF>
__global__ void __launch_bounds__(256) bias_gelu_fwd_kernel(const void* __restrict__ h, const float4* __restrict__ bias,
                                                            long long n4, int C4, float4* __restrict__ out,
                                                            __nv_bfloat16* __restrict__ out_bf16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = HBF ? load_bf16x4((const __nv_bfloat16*)h, i) : __ldg((const float4*)h + i), b = __ldg(bias + (int)(i % C4));
    float4 o = make_float4(gelu_t<HBF>(a.x + b.x), gelu_t<HBF>(a.y + b.y), gelu_t<HBF>(a.z + b.z), gelu_t<HBF>(a.w + b.w));
    if (out) out[i] = o;
    if (out_bf16) store_bf16x4(out_bf16, i, o);
  }
}

// dh = dg * gelu'(h + b); partial column sums of dh (bias gradient).  Thread t owns float4 column (t % C4) of
// rows (t / C4), stepping whole CTAs.
template <bool HBF>
__global__ void __launch_bounds__(256) bias_gelu_bwd_kernel(const void* __restrict__ h, const float4* __restrict__ bias,
                                                            const void* __restrict__ dg, long long N, int C4,
                                                            float4* __restrict__ dh, __nv_bfloat16* __restrict__ dh_bf16,
                                                            float* __restrict__ partial, float* __restrict__ dbias, int accumulate) {
  int c = threadIdx.x % C4, rsub = threadIdx.x / C4, rper = blockDim.x / C4;
  float4 b = __ldg(bias + c);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = (long long)blockIdx.x * rper + rsub; row < N; row += (long long)gridDim.x * rper) {
    float4 a = HBF ? load_bf16x4((const __nv_bfloat16*)h, row * C4 + c) : __ldg((const float4*)h + row * C4 + c);
    float4 g = HBF ? load_bf16x4((const __nv_bfloat16*)dg, row * C4 + c) : __ldg((const float4*)dg + row * C4 + c);
    float4 o = make_float4(g.x * gelu_grad_t<HBF>(a.x + b.x), g.y * gelu_grad_t<HBF>(a.y + b.y), g.z * gelu_grad_t<HBF>(a.z + b.z),
                           g.w * gelu_grad_t<HBF>(a.w + b.w));
    if (dh) dh[row * C4 + c] = o;
    if (dh_bf16) store_bf16x4(dh_bf16, row * C4 + c, o);
    acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
  }
  __shared__ float4 red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  if (rsub == 0) {
    for (int j = 1; j < rper; ++j) {
      float4 t = red[j * C4 + c];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    *reinterpret_cast<float4*>(partial + (long long)blockIdx.x * 4 * C4 + 4 * c) = acc;
  }
  ew_cta_atomic_add(partial + (long long)blockIdx.x * 4 * C4, 4 * C4, dbias, 4 * C4, nullptr);
}

// out / out_bf16 (N,C): either may be NULL.  Internal form: h (and dg in the backward) may be bf16 tensors.
int ew_bias_gelu_fwd(const void* h, int h_bf16, const float* bias, int64_t N, int C, float* out, void* out_bf16, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && C > 0 && (C % 4) == 0 && (out || out_bf16));
  if (N == 0) return GDMAE_OK;
  const int grid = gdmae_grid(N * (C / 4), 256, 8);
  if (h_bf16)
    bias_gelu_fwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream_>>>(h, (const float4*)bias, N * (C / 4), C / 4, (float4*)out,
                                                                        (__nv_bfloat16*)out_bf16);
  else
    bias_gelu_fwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream_>>>(h, (const float4*)bias, N * (C / 4), C / 4, (float4*)out,
                                                                         (__nv_bfloat16*)out_bf16);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

extern "C" int gdmae_bias_gelu_fwd(const float* h, const float* bias, int64_t N, int C, float* out, void* out_bf16, void* stream_) {
  return ew_bias_gelu_fwd(h, 0, bias, N, C, out, out_bf16, stream_);
}

int ew_bias_gelu_bwd(const void* h, const void* dg, int hdg_bf16, const float* bias, int64_t N, int C, float* dh, void* dh_bf16,
                     float* dbias, int accumulate, void* workspace, size_t ws_bytes, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && C > 0 && (C % 4) == 0 && (C / 4) <= 256 && 256 % (C / 4) == 0 && (dh || dh_bf16));
  if (ws_bytes < gdmae_rowwise_workspace_bytes(C)) { gdmae_set_error("bias_gelu_bwd: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  if (N == 0) return GDMAE_OK;
  cudaStream_t st = (cudaStream_t)stream_;
  int rper = 256 / (C / 4);
  long long need = (N + rper - 1) / rper;
  int grid = (int)(need < EW_PART_BLOCKS ? need : EW_PART_BLOCKS);
  float* partial = (float*)workspace;
  if (!accumulate) GDMAE_CHECK_CUDA(cudaMemsetAsync(dbias, 0, (size_t)C * 4, st));
  if (hdg_bf16)
    bias_gelu_bwd_kernel<true><<<grid, 256, 0, st>>>(h, (const float4*)bias, dg, N, C / 4, (float4*)dh, (__nv_bfloat16*)dh_bf16, partial,
                                                     dbias, accumulate);
  else
    bias_gelu_bwd_kernel<false><<<grid, 256, 0, st>>>(h, (const float4*)bias, dg, N, C / 4, (float4*)dh, (__nv_bfloat16*)dh_bf16, partial,
                                                      dbias, accumulate);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

extern "C" int gdmae_bias_gelu_bwd(const float* h, const float* bias, const float* dg, int64_t N, int C, float* dh, void* dh_bf16,
                                   float* dbias, int accumulate, void* workspace, size_t ws_bytes, void* stream_) {
  return ew_bias_gelu_bwd(h, dg, 0, bias, N, C, dh, dh_bf16, dbias, accumulate, workspace, ws_bytes, stream_);
}

// ------------------------------------------------------------------ column sums (bias gradients of the GEMMs)
template <bool BF16>
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ x_, long long N, int ld4, int col4, int C4,
                                                     float* __restrict__ partial, float* __restrict__ out, int accumulate) {
  int c = threadIdx.x % C4, rsub = threadIdx.x / C4, rper = blockDim.x / C4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // eight rows in flight per thread: with one 8/16-byte load per iteration the kernel sat at 0.24 of the HBM rate, with four
  // at 0.35-0.40 (r2: 26 us per launch for 69 MB)
  const long long step = (long long)gridDim.x * rper;
  long long row = (long long)blockIdx.x * rper + rsub;
  for (; row + 7 * step < N; row += 8 * step) {
    float4 a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      a[u] = BF16 ? load_bf16x4((const __nv_bfloat16*)x_, (row + u * step) * ld4 + col4 + c)
                  : __ldg((const float4*)x_ + (row + u * step) * ld4 + col4 + c);
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc.x += a[u].x; acc.y += a[u].y; acc.z += a[u].z; acc.w += a[u].w; }
  }
  for (; row < N; row += step) {
    float4 a = BF16 ? load_bf16x4((const __nv_bfloat16*)x_, row * ld4 + col4 + c) : __ldg((const float4*)x_ + row * ld4 + col4 + c);
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
  }
  __shared__ float4 red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  if (rsub == 0) {
    for (int j = 1; j < rper; ++j) {
      float4 t = red[j * C4 + c];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    *reinterpret_cast<float4*>(partial + (long long)blockIdx.x * 4 * C4 + 4 * c) = acc;
  }
  ew_cta_atomic_add(partial + (long long)blockIdx.x * 4 * C4, 4 * C4, out, 4 * C4, nullptr);
}

// out (C) = (accumulate ? out : 0) + sum over rows of x (N, ld) columns [col0, col0 + C); x fp32 (dtype 0) or bf16 (1)
extern "C" int gdmae_colsum(const void* x, int dtype, int64_t N, int ld, int col0, int C, float* out, int accumulate,
                            void* workspace, size_t ws_bytes, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && C > 0 && (C % 4) == 0 && (ld % 4) == 0 && (col0 % 4) == 0 && col0 + C <= ld && (C / 4) <= 256);
  if (ws_bytes < gdmae_rowwise_workspace_bytes(C)) { gdmae_set_error("colsum: workspace too small"); return GDMAE_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream_;
  if (N == 0) {
    if (!accumulate) GDMAE_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)C * 4, st));
    return GDMAE_OK;
  }
  int C4 = C / 4;
  int rper = 256 / C4;
  long long need = (N + rper - 1) / rper;
  int grid = (int)(need < EW_PART_BLOCKS ? need : EW_PART_BLOCKS);
  float* partial = (float*)workspace;
  if (!accumulate) GDMAE_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)C * 4, st));
  if (dtype == 0) colsum_kernel<false><<<grid, C4 * rper, 0, st>>>(x, N, ld / 4, col0 / 4, C4, partial, out, accumulate);
  else colsum_kernel<true><<<grid, C4 * rper, 0, st>>>(x, N, ld / 4, col0 / 4, C4, partial, out, accumulate);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// ------------------------------------------------------------------ out[n] = x[n] + table[idx[n]]  (q = k = feat + pos)
__global__ void __launch_bounds__(256) gather_add_kernel(const float4* __restrict__ x, const float4* __restrict__ table,
                                                         const unsigned char* __restrict__ idx, long long n4, int C4,
                                                         float4* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    long long row = i / C4;
    int c = (int)(i % C4);
    float4 a = __ldg(x + i), t = __ldg(table + (int)idx[row] * C4 + c);
    float4 o = make_float4(a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w);
    if (out) out[i] = o;
    if (out_bf16) store_bf16x4(out_bf16, i, o);
  }
}

// out / out_bf16 (N,C) = x (N,C) + table (64,C)[idx (N) uint8]; either output may be NULL
extern "C" int gdmae_gather_add_rows(const float* x, const float* table, const uint8_t* idx, int64_t N, int C, float* out,
                                     void* out_bf16, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && C > 0 && (C % 4) == 0 && (out || out_bf16));
  if (N == 0) return GDMAE_OK;
  gather_add_kernel<<<gdmae_grid(N * (C / 4), 256, 8), 256, 0, (cudaStream_t)stream_>>>(
      (const float4*)x, (const float4*)table, idx, N * (C / 4), C / 4, (float4*)out, (__nv_bfloat16*)out_bf16);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
