// Feature movers of the sparse pyramid: sparse-conv row gather (im2col over the 3x3 neighbour
// map), its transposed gather for the backward pass, and the decoder's sparse->dense BEV fill.
//
// Replaces (reference file:line, relative to /root/reference):
//   spconv SubMConv2d / SparseConv2d gather-GEMM-scatter   pcdet/utils/spconv_utils.py:37-56 (third party spconv 2.x)
//   SparseConvTensor.dense() + ConvTranspose2d(k=s) + BatchNorm2d + ReLU + torch.cat
//                                                          pcdet/models/backbones_3d/spt_backbone_mae.py:125-132
//
// Decoder design (B200-first): ConvTranspose2d with kernel == stride maps every active site to
// its own k x k block and every empty cell to exactly 0 (no bias), so after BatchNorm+ReLU the
// dense 384-channel map is "one constant vector per scale" everywhere except at the cells
// covered by active sites.  The three deconvs, BNs, ReLUs and the concat therefore collapse into
// per-site GEMMs on the sparse rows plus ONE write-only pass over the dense NHWC map
// (gdmae_dense_fill); the reference makes ten dense passes for the same tensor.
#include "common.cuh"
#include <cuda_bf16.h>

// 4-channel packets in fp32 (16 B) or bf16 (8 B)
template <typename T> struct Pack4;
template <> struct Pack4<float> {
  typedef float4 type;
  static __device__ __forceinline__ float4 load(const float* p, long long i4) { return __ldg(reinterpret_cast<const float4*>(p) + i4); }
  static __device__ __forceinline__ void store(float* p, long long i4, float4 v) { reinterpret_cast<float4*>(p)[i4] = v; }
};
template <> struct Pack4<__nv_bfloat16> {
  typedef uint2 type;
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p, long long i4) {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(p) + i4);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, long long i4, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<unsigned int*>(&a);
    u.y = *reinterpret_cast<unsigned int*>(&b);
    reinterpret_cast<uint2*>(p)[i4] = u;
  }
};

// out[n, k*C + c] = src[map[n,k], c] (0 where map < 0).  One thread per float4.
template <typename T>
__global__ void gather_rows_kernel(const float4* __restrict__ src, const int* __restrict__ map, long long N, int K, int C4,
                                   T* __restrict__ out) {
  long long total = N * K * C4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int c = (int)(t % C4);
    long long nk = t / C4;
    int m = __ldg(map + nk);
    Pack4<T>::store(out, t, m >= 0 ? __ldg(src + (long long)m * C4 + c) : make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

// out (N, K*C) in fp32 (out_dtype 0) or bf16 (1): the GEMM operand of the sparse conv
extern "C" int gdmae_gather_rows(const float* src, const int32_t* map, int64_t N, int K, int C, void* out, int out_dtype,
                                 void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && K > 0 && C > 0 && (C % 4) == 0);
  if (N == 0) return GDMAE_OK;
  int g = gdmae_grid(N * K * (C / 4), 256, 32);
  cudaStream_t st = (cudaStream_t)stream_;
  if (out_dtype == 0) gather_rows_kernel<float><<<g, 256, 0, st>>>((const float4*)src, map, N, K, C / 4, (float*)out);
  else gather_rows_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const float4*)src, map, N, K, C / 4, (__nv_bfloat16*)out);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// dsrc[i, c] = sum_k dcol[tmap[i, mirror ? K-1-k : k], k*C + c]   (gather form of the scatter-add:
// deterministic, no atomics).  For SubM convs tmap is the forward map and mirror = 1
// (nbr[n,k] = m  <=>  nbr[m,8-k] = n); for strided convs tmap is the "up" map and mirror = 0.
template <typename T>
__global__ void gather_rows_t_kernel(const T* __restrict__ dcol, const int* __restrict__ tmap, long long N, int K, int C4,
                                     int mirror, float4* __restrict__ dsrc) {
  long long total = N * C4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int c = (int)(t % C4);
    long long i = t / C4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      int m = __ldg(tmap + i * K + (mirror ? K - 1 - k : k));
      if (m >= 0) {
        float4 v = Pack4<T>::load(dcol, ((long long)m * K + k) * C4 + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    dsrc[t] = acc;
  }
}

// dcol in fp32 (in_dtype 0) or bf16 (1); dsrc fp32
extern "C" int gdmae_gather_rows_transposed(const void* dcol, int in_dtype, const int32_t* tmap, int64_t N, int K, int C,
                                            int mirror, float* dsrc, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && K > 0 && C > 0 && (C % 4) == 0);
  if (N == 0) return GDMAE_OK;
  int g = gdmae_grid(N * (C / 4), 256, 32);
  cudaStream_t st = (cudaStream_t)stream_;
  if (in_dtype == 0) gather_rows_t_kernel<float><<<g, 256, 0, st>>>((const float*)dcol, tmap, N, K, C / 4, mirror, (float4*)dsrc);
  else gather_rows_t_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)dcol, tmap, N, K, C / 4, mirror, (float4*)dsrc);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// ------------------------------------------------------------------ decoder dense fill
// out (B, Y, X, 3*Cs) NHWC.  Scale s in {0,1,2} has stride k_s = 1,2,4, a rank grid on the
// (Y/k_s, X/k_s) lattice and rows a_s (N_s * k_s^2, Cs): row (rank*k_s^2 + (y%k_s)*k_s + x%k_s).
// Cells whose scale-s site is empty receive bg_s (Cs).
struct DenseFillArgs {
  const float* rows[3];
  const float* bg[3];
  const int* grid[3];
  int k[3];
  int H[3], W[3];
  int B, Y, X, Cs;
};

// 8-channel packets: fp32 = two 16-byte stores, bf16 = one 16-byte store
template <typename T> struct Pack8;
template <> struct Pack8<float> {
  static __device__ __forceinline__ void store(float* p, long long i8, float4 a, float4 b) {
    __stcs(reinterpret_cast<float4*>(p) + 2 * i8, a);
    __stcs(reinterpret_cast<float4*>(p) + 2 * i8 + 1, b);
  }
  static __device__ __forceinline__ void load(const float* p, long long i8, float4& a, float4& b) {
    a = __ldcs(reinterpret_cast<const float4*>(p) + 2 * i8);
    b = __ldcs(reinterpret_cast<const float4*>(p) + 2 * i8 + 1);
  }
};
template <> struct Pack8<__nv_bfloat16> {
  static __device__ __forceinline__ void store(__nv_bfloat16* p, long long i8, float4 a, float4 b) {
    __nv_bfloat162 x0 = __floats2bfloat162_rn(a.x, a.y), x1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 x2 = __floats2bfloat162_rn(b.x, b.y), x3 = __floats2bfloat162_rn(b.z, b.w);
    float4 u;
    u.x = __uint_as_float(*reinterpret_cast<unsigned*>(&x0)); u.y = __uint_as_float(*reinterpret_cast<unsigned*>(&x1));
    u.z = __uint_as_float(*reinterpret_cast<unsigned*>(&x2)); u.w = __uint_as_float(*reinterpret_cast<unsigned*>(&x3));
    __stcs(reinterpret_cast<float4*>(p) + i8, u);
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, long long i8, float4& a, float4& b) {
    float4 u = __ldcs(reinterpret_cast<const float4*>(p) + i8);
    unsigned w0 = __float_as_uint(u.x), w1 = __float_as_uint(u.y), w2 = __float_as_uint(u.z), w3 = __float_as_uint(u.w);
    float2 f0 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w0)), f1 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w1));
    float2 f2 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w2)), f3 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w3));
    a = make_float4(f0.x, f0.y, f1.x, f1.y);
    b = make_float4(f2.x, f2.y, f3.x, f3.y);
  }
};

// one thread = 8 channels of one (cell, scale): a streaming 16-byte store per thread in bf16.
// grid = (Y, B): a CTA walks one map row, 8 cells x 3 scales x C8 lanes per trip (the 8 cells' 6 KB are
// contiguous in the output), four trips in flight.  Coordinates come from the block index; the only
// per-trip index math is x += 8 and a shift (strides are powers of two).
template <typename T>
__global__ void __launch_bounds__(384) dense_fill_kernel(DenseFillArgs a, T* __restrict__ out) {
  const int C8 = a.Cs >> 3;                 // 16 for Cs = 128
  const int per_cell = 3 * C8;
  const int xl = threadIdx.x / per_cell, rem = threadIdx.x - xl * per_cell;
  const int s = rem / C8, c = rem - s * C8;
  const int y = blockIdx.x, b = blockIdx.y;
  const int sh = a.k[s] == 1 ? 0 : (a.k[s] == 2 ? 1 : 2);
  const int k = 1 << sh;
  const int gy = y >> sh;
  const bool row_ok = gy < a.H[s];
  const int Ws = a.W[s];
  const int* grow = a.grid[s] + ((long long)b * a.H[s] + gy) * Ws;
  const float4* rows = reinterpret_cast<const float4*>(a.rows[s]);
  const float4* bg = reinterpret_cast<const float4*>(a.bg[s]) + 2 * c;
  const int suby = (y - (gy << sh)) * k;
  const long long t0 = ((long long)b * a.Y + y) * a.X * per_cell + rem;
#pragma unroll 4
  for (int x = xl; x < a.X; x += 8) {
    const int gx = x >> sh;
    int rank = (row_ok && gx < Ws) ? __ldg(grow + gx) : -1;
    const float4* src = rank >= 0 ? rows + ((long long)rank * k * k + suby + (x - (gx << sh))) * (2 * C8) + 2 * c : bg;
    Pack8<T>::store(out, t0 + (long long)x * per_cell, __ldg(src), __ldg(src + 1));
  }
}

// backward: drows_s[row] = dout[cell, s*Cs : (s+1)*Cs] at covered cells (gather),
//           dbg_s[c]     = sum over uncovered cells of dout[cell, s*Cs + c].
template <typename T>
__global__ void dense_fill_bwd_rows_kernel(DenseFillArgs a, int s, const int* __restrict__ indices, long long Ns,
                                           const T* __restrict__ dout, float4* __restrict__ drows) {
  int C4 = a.Cs >> 2;
  int k = a.k[s], kk = k * k;
  long long total = Ns * kk * C4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int c = (int)(t % C4);
    long long row = t / C4;
    long long n = row / kk;
    int sub = (int)(row % kk);
    int b = indices[3 * n], y = indices[3 * n + 1] * k + sub / k, x = indices[3 * n + 2] * k + sub % k;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < a.Y && x < a.X) v = Pack4<T>::load(dout, ((((long long)b * a.Y + y) * a.X + x) * 3 + s) * C4 + c);
    drows[t] = v;
  }
}

// column sums of dout over the cells NOT covered at scale s.  grid = (Y*B / rows_per_cta, 3); block 256 =
// 16 cell-lanes x 16 channel-lanes (8 channels = one 16-byte streaming load in bf16); a CTA walks whole map rows,
// so the only per-cell index math is a shift.  Per-block partials are combined with float atomics into dbg
// (3*Cs, caller zeroes).
template <typename T>
__global__ void __launch_bounds__(256) dense_fill_bwd_bg_kernel(DenseFillArgs a, const T* __restrict__ dout,
                                                               float* __restrict__ dbg) {
  const int C8 = a.Cs >> 3;  // == 16 for Cs = 128
  const int s = blockIdx.y;
  const int lane = threadIdx.x % C8, sub = threadIdx.x / C8, nsub = blockDim.x / C8;
  const int sh = a.k[s] == 1 ? 0 : (a.k[s] == 2 ? 1 : 2);
  const int n_rows = a.B * a.Y;
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int b = row / a.Y, y = row - b * a.Y;
    const int gy = y >> sh;
    const bool row_ok = gy < a.H[s];
    const int* grow = a.grid[s] + ((long long)b * a.H[s] + gy) * a.W[s];
    const long long base = (long long)row * a.X;
    for (int x = sub; x < a.X; x += nsub) {
      int gx = x >> sh;
      int rank = (row_ok && gx < a.W[s]) ? __ldg(grow + gx) : -1;
      if (rank < 0) {
        float4 v0, v1;
        Pack8<T>::load(dout, ((base + x) * 3 + s) * C8 + lane, v0, v1);
        acc0.x += v0.x; acc0.y += v0.y; acc0.z += v0.z; acc0.w += v0.w;
        acc1.x += v1.x; acc1.y += v1.y; acc1.z += v1.z; acc1.w += v1.w;
      }
    }
  }
  __shared__ float4 red[2][256];
  red[0][threadIdx.x] = acc0;
  red[1][threadIdx.x] = acc1;
  __syncthreads();
  if (sub == 0) {
    for (int j = 1; j < nsub; ++j) {
      float4 v0 = red[0][j * C8 + lane], v1 = red[1][j * C8 + lane];
      acc0.x += v0.x; acc0.y += v0.y; acc0.z += v0.z; acc0.w += v0.w;
      acc1.x += v1.x; acc1.y += v1.y; acc1.z += v1.z; acc1.w += v1.w;
    }
    float* dst = dbg + s * a.Cs + 8 * lane;
    atomicAdd(dst, acc0.x); atomicAdd(dst + 1, acc0.y); atomicAdd(dst + 2, acc0.z); atomicAdd(dst + 3, acc0.w);
    atomicAdd(dst + 4, acc1.x); atomicAdd(dst + 5, acc1.y); atomicAdd(dst + 6, acc1.z); atomicAdd(dst + 7, acc1.w);
  }
}

static int fill_args(DenseFillArgs& a, const float* const* rows, const float* const* bg, const int32_t* const* grids,
                     const int* strides, int B, int Y, int X, int Cs) {
  GDMAE_CHECK_ARG(B >= 1 && Y >= 1 && X >= 1 && Cs > 0 && (Cs % 8) == 0 && 256 % (Cs / 8) == 0);
  for (int s = 0; s < 3; ++s) {
    GDMAE_CHECK_ARG(strides[s] >= 1);
    a.rows[s] = rows ? rows[s] : nullptr;
    a.bg[s] = bg ? bg[s] : nullptr;
    a.grid[s] = grids[s];
    a.k[s] = strides[s];
    // lattice of the scale-s sites: 468 -> 234 -> 117 (each level (H-1)/2+1)
    int h = Y, w = X;
    for (int kk = strides[s]; kk > 1; kk >>= 1) { h = (h - 1) / 2 + 1; w = (w - 1) / 2 + 1; }
    a.H[s] = h; a.W[s] = w;
  }
  a.B = B; a.Y = Y; a.X = X; a.Cs = Cs;
  return GDMAE_OK;
}

template <typename T>
static int dense_fill_impl(const float* const* rows, const float* const* bg, const int32_t* const* rank_grids, const int* strides,
                           int B, int Y, int X, int Cs, T* out, void* stream_) {
  DenseFillArgs a;
  int rc = fill_args(a, rows, bg, rank_grids, strides, B, Y, X, Cs);
  if (rc) return rc;
  GDMAE_CHECK_ARG(Cs == 128 && B < 65536);  // block = 8 cells x 3 scales x 16 lanes
  for (int s = 0; s < 3; ++s) GDMAE_CHECK_ARG(strides[s] == 1 || strides[s] == 2 || strides[s] == 4);
  dim3 grid(Y, B);
  dense_fill_kernel<T><<<grid, 384, 0, (cudaStream_t)stream_>>>(a, out);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

template <typename T>
static int dense_fill_bwd_impl(const T* dout, const int32_t* const* rank_grids, const int32_t* const* indices, const int64_t* n_sites,
                               const int* strides, int B, int Y, int X, int Cs, float* const* drows, float* dbg, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DenseFillArgs a;
  int rc = fill_args(a, nullptr, nullptr, rank_grids, strides, B, Y, X, Cs);
  if (rc) return rc;
  GDMAE_CHECK_CUDA(cudaMemsetAsync(dbg, 0, (size_t)3 * Cs * 4, st));
  for (int s = 0; s < 3; ++s) {
    long long total = n_sites[s] * strides[s] * strides[s] * (Cs / 4);
    if (total == 0) continue;
    dense_fill_bwd_rows_kernel<T><<<gdmae_grid(total, 256, 32), 256, 0, st>>>(a, s, indices[s], n_sites[s], dout, (float4*)drows[s]);
    GDMAE_LAUNCH_CHECK();
  }
  dim3 grid(GDMAE_NUM_SMS * 8, 3);
  dense_fill_bwd_bg_kernel<T><<<grid, 256, 0, st>>>(a, dout, dbg);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// out_dtype / dtype: 0 = fp32, 1 = bf16 (the dense BEV map feeds the cuDNN decoder conv)
extern "C" int gdmae_dense_fill(const float* const* rows, const float* const* bg, const int32_t* const* rank_grids,
                                const int* strides, int B, int Y, int X, int Cs, void* out, int out_dtype, void* stream_) {
  if (out_dtype == 0) return dense_fill_impl<float>(rows, bg, rank_grids, strides, B, Y, X, Cs, (float*)out, stream_);
  return dense_fill_impl<__nv_bfloat16>(rows, bg, rank_grids, strides, B, Y, X, Cs, (__nv_bfloat16*)out, stream_);
}

extern "C" int gdmae_dense_fill_bwd(const void* dout, int dtype, const int32_t* const* rank_grids, const int32_t* const* indices,
                                    const int64_t* n_sites, const int* strides, int B, int Y, int X, int Cs,
                                    float* const* drows, float* dbg /* (3*Cs) */, void* stream_) {
  if (dtype == 0) return dense_fill_bwd_impl<float>((const float*)dout, rank_grids, indices, n_sites, strides, B, Y, X, Cs, drows, dbg, stream_);
  return dense_fill_bwd_impl<__nv_bfloat16>((const __nv_bfloat16*)dout, rank_grids, indices, n_sites, strides, B, Y, X, Cs, drows, dbg, stream_);
}

// out[m, :] = src[b, y, x, :] for NHWC src (B,Y,X,C) at the M pillar cells (coalesced row gather);
// coords are the int64 (M,4) [b,z,y,x] voxel coords.  spt_backbone_mae.py:141-143.
template <typename T>
__global__ void gather_nhwc_kernel(const T* __restrict__ src, const long long* __restrict__ coords, long long M, int Y, int X,
                                   int C4, float4* __restrict__ out) {
  long long total = M * C4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int c = (int)(t % C4);
    long long m = t / C4;
    long long b = coords[4 * m], y = coords[4 * m + 2], x = coords[4 * m + 3];
    out[t] = Pack4<T>::load(src, ((b * Y + y) * X + x) * C4 + c);
  }
}
template <typename T>
__global__ void scatter_nhwc_kernel(const float4* __restrict__ dout, const long long* __restrict__ coords, long long M, int Y, int X,
                                    int C4, T* __restrict__ dsrc) {
  long long total = M * C4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int c = (int)(t % C4);
    long long m = t / C4;
    long long b = coords[4 * m], y = coords[4 * m + 2], x = coords[4 * m + 3];
    Pack4<T>::store(dsrc, ((b * Y + y) * X + x) * C4 + c, dout[t]);
  }
}

// src (B,Y,X,C) NHWC in fp32 (dtype 0) or bf16 (dtype 1); out (M, C) fp32
extern "C" int gdmae_gather_nhwc(const void* src, int dtype, const int64_t* voxel_coords, int64_t M, int Y, int X, int C, float* out,
                                 void* stream_) {
  GDMAE_CHECK_ARG(M >= 0 && C > 0 && (C % 4) == 0);
  if (M == 0) return GDMAE_OK;
  int g = gdmae_grid(M * (C / 4), 256, 32);
  cudaStream_t st = (cudaStream_t)stream_;
  if (dtype == 0) gather_nhwc_kernel<float><<<g, 256, 0, st>>>((const float*)src, (const long long*)voxel_coords, M, Y, X, C / 4, (float4*)out);
  else gather_nhwc_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (const long long*)voxel_coords, M, Y, X, C / 4, (float4*)out);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// dsrc must be zero-filled by the caller (pillar cells are unique, so plain stores suffice).
extern "C" int gdmae_scatter_nhwc(const float* dout, const int64_t* voxel_coords, int64_t M, int Y, int X, int C, void* dsrc,
                                  int dtype, void* stream_) {
  GDMAE_CHECK_ARG(M >= 0 && C > 0 && (C % 4) == 0);
  if (M == 0) return GDMAE_OK;
  int g = gdmae_grid(M * (C / 4), 256, 32);
  cudaStream_t st = (cudaStream_t)stream_;
  if (dtype == 0) scatter_nhwc_kernel<float><<<g, 256, 0, st>>>((const float4*)dout, (const long long*)voxel_coords, M, Y, X, C / 4, (float*)dsrc);
  else scatter_nhwc_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const float4*)dout, (const long long*)voxel_coords, M, Y, X, C / 4, (__nv_bfloat16*)dsrc);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
