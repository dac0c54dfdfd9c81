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
#include <algorithm>
#include <cstdint>

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

// bf16 im2col, C / 8 a power of two (C = 128, 256): thread = 8 channels (two float4 loads of the fp32 source row, one
// 16-byte store), index arithmetic in shifts, four (row, tap) pairs per trip with the map entries fetched first and the
// eight row loads issued together (r2: the generic kernel above - one float4 per thread, a 64-bit division per element,
// map -> row as a dependent chain - ran at 2.7 TB/s, 365 us per step over its five launches).
__global__ void __launch_bounds__(256) gather_rows_bf16_kernel(const float4* __restrict__ src, const int* __restrict__ map, int NK, int sh8,
                                                               uint4* __restrict__ out) {
  const int C8 = 1 << sh8;
  const int c = threadIdx.x & (C8 - 1);
  const int per = blockDim.x >> sh8;                 // (row, tap) pairs per CTA per unit
  const int step = gridDim.x * per;
  for (int nk0 = blockIdx.x * per + (threadIdx.x >> sh8); nk0 < NK; nk0 += 4 * step) {
    int m[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) m[u] = nk0 + u * step < NK ? __ldg(map + nk0 + u * step) : -1;
    float4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a[u] = make_float4(0.f, 0.f, 0.f, 0.f); b[u] = a[u];
      if (m[u] >= 0) {
        const float4* r = src + (((long long)m[u] << sh8) + c) * 2;
        a[u] = __ldg(r); b[u] = __ldg(r + 1);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int nk = nk0 + u * step;
      if (nk < NK) {
        __nv_bfloat162 x0 = __floats2bfloat162_rn(a[u].x, a[u].y), x1 = __floats2bfloat162_rn(a[u].z, a[u].w);
        __nv_bfloat162 x2 = __floats2bfloat162_rn(b[u].x, b[u].y), x3 = __floats2bfloat162_rn(b[u].z, b[u].w);
        out[((long long)nk << sh8) + c] = make_uint4(*reinterpret_cast<unsigned*>(&x0), *reinterpret_cast<unsigned*>(&x1),
                                                     *reinterpret_cast<unsigned*>(&x2), *reinterpret_cast<unsigned*>(&x3));
      }
    }
  }
}

// transposed gather of bf16 columns, K = 9, C / 8 a power of two: the nine map entries of an output row are fetched first,
// then the (up to) nine 16-byte pieces together; accumulation in tap order (deterministic, as the generic kernel)
__global__ void __launch_bounds__(256) gather_rows_t_bf16_k9_kernel(const uint4* __restrict__ dcol, const int* __restrict__ tmap, int N, int sh8,
                                                                    int mirror, float4* __restrict__ dsrc) {
  const int C8 = 1 << sh8;
  const int c = threadIdx.x & (C8 - 1);
  const int per = blockDim.x >> sh8;
  for (int i = blockIdx.x * per + (threadIdx.x >> sh8); i < N; i += gridDim.x * per) {
    int m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = __ldg(tmap + (long long)i * 9 + (mirror ? 8 - k : k));
    uint4 v[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      v[k] = make_uint4(0u, 0u, 0u, 0u);
      if (m[k] >= 0) v[k] = __ldg(dcol + ((((long long)m[k] * 9 + k) << sh8) + c));
    }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const unsigned w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[2 * j] += __uint_as_float(w[j] << 16); acc[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u); }
    }
    float4* d = dsrc + (((long long)i << sh8) + c) * 2;
    d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

static inline int pow2_shift(int v) {      // log2(v) if v is a power of two in [1, 256], else -1
  for (int s = 0; s <= 8; ++s) if ((1 << s) == v) return s;
  return -1;
}

// out (N, K*C) in fp32 (out_dtype 0) or bf16 (1): the GEMM operand of the sparse conv
extern "C" int gdmae_gather_rows(const float* src, const int32_t* map, int64_t N, int K, int C, void* out, int out_dtype,
                                 void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && K > 0 && C > 0 && (C % 4) == 0);
  if (N == 0) return GDMAE_OK;
  int g = gdmae_grid(N * K * (C / 4), 256, 32);
  cudaStream_t st = (cudaStream_t)stream_;
  const int sh8 = (C % 8) == 0 ? pow2_shift(C / 8) : -1;
  if (out_dtype == 1 && sh8 >= 0 && sh8 <= 8 && N * K < (1ll << 30) && ((uintptr_t)out & 15) == 0) {
    const long long NK = N * K;
    const int per = 256 >> sh8;
    const int gb = (int)std::min<long long>((long long)GDMAE_NUM_SMS * 8, (NK + 4 * per - 1) / (4 * per));
    gather_rows_bf16_kernel<<<gb, 256, 0, st>>>((const float4*)src, map, (int)NK, sh8, (uint4*)out);
  }
  else if (out_dtype == 0) gather_rows_kernel<float><<<g, 256, 0, st>>>((const float4*)src, map, N, K, C / 4, (float*)out);
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
  const int sh8 = (C % 8) == 0 ? pow2_shift(C / 8) : -1;
  if (in_dtype == 1 && K == 9 && sh8 >= 0 && N < (1ll << 27) && ((uintptr_t)dcol & 15) == 0) {
    const int per = 256 >> sh8;
    const int gb = (int)std::min<long long>((long long)GDMAE_NUM_SMS * 8, (N + per - 1) / per);
    gather_rows_t_bf16_k9_kernel<<<gb, 256, 0, st>>>((const uint4*)dcol, tmap, (int)N, sh8, mirror, (float4*)dsrc);
  }
  else if (in_dtype == 0) gather_rows_t_kernel<float><<<g, 256, 0, st>>>((const float*)dcol, tmap, N, K, C / 4, mirror, (float4*)dsrc);
  else gather_rows_t_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)dcol, tmap, N, K, C / 4, mirror, (float4*)dsrc);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// ------------------------------------------------------------------ decoder dense fill
// out (B, Y, X, 3*Cs) NHWC.  Scale s in {0,1,2} has stride k_s = 1,2,4, a rank grid on the
// (Y/k_s, X/k_s) lattice and rows a_s (N_s * k_s^2, Cs): row (rank*k_s^2 + (y%k_s)*k_s + x%k_s).
// Cells whose scale-s site is empty receive bg_s (Cs).
struct DenseFillArgs {
  const void* rows[3];       // fp32 or bf16 (rows_dtype of the entry points)
  const float* bg[3];
  const int* grid[3];
  int k[3];
  int H[3], W[3];
  int B, Y, X, Cs;
};

// 8 channels as they lie in memory: fp32 = two float4, bf16 = one uint4.  Same-type moves copy the bits.
template <typename T> struct P8;
template <> struct P8<float> {
  float4 a, b;
  static __device__ __forceinline__ P8 load_stream(const float* p, long long i8) {
    P8 r; r.a = __ldcs(reinterpret_cast<const float4*>(p) + 2 * i8); r.b = __ldcs(reinterpret_cast<const float4*>(p) + 2 * i8 + 1); return r;
  }
  static __device__ __forceinline__ P8 load_keep(const float* p, long long i8) {
    P8 r; r.a = __ldg(reinterpret_cast<const float4*>(p) + 2 * i8); r.b = __ldg(reinterpret_cast<const float4*>(p) + 2 * i8 + 1); return r;
  }
  __device__ __forceinline__ void store_stream(float* p, long long i8) const {
    __stcs(reinterpret_cast<float4*>(p) + 2 * i8, a); __stcs(reinterpret_cast<float4*>(p) + 2 * i8 + 1, b);
  }
  __device__ __forceinline__ void add_to(float4& s0, float4& s1) const {
    s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w; s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
  }
};
template <> struct P8<__nv_bfloat16> {
  uint4 u;
  static __device__ __forceinline__ P8 load_stream(const __nv_bfloat16* p, long long i8) {
    P8 r; r.u = __ldcs(reinterpret_cast<const uint4*>(p) + i8); return r;
  }
  static __device__ __forceinline__ P8 load_keep(const __nv_bfloat16* p, long long i8) {
    P8 r; r.u = __ldg(reinterpret_cast<const uint4*>(p) + i8); return r;
  }
  __device__ __forceinline__ void store_stream(__nv_bfloat16* p, long long i8) const { __stcs(reinterpret_cast<uint4*>(p) + i8, u); }
  __device__ __forceinline__ void add_to(float4& s0, float4& s1) const {
    s0.x += __uint_as_float(u.x << 16); s0.y += __uint_as_float(u.x & 0xffff0000u);
    s0.z += __uint_as_float(u.y << 16); s0.w += __uint_as_float(u.y & 0xffff0000u);
    s1.x += __uint_as_float(u.z << 16); s1.y += __uint_as_float(u.z & 0xffff0000u);
    s1.z += __uint_as_float(u.w << 16); s1.w += __uint_as_float(u.w & 0xffff0000u);
  }
};
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<unsigned*>(&v);
}
template <typename TO, typename TI> __device__ __forceinline__ P8<TO> p8_convert(const P8<TI>& v);
template <> __device__ __forceinline__ P8<float> p8_convert<float, float>(const P8<float>& v) { return v; }
template <> __device__ __forceinline__ P8<__nv_bfloat16> p8_convert<__nv_bfloat16, __nv_bfloat16>(const P8<__nv_bfloat16>& v) { return v; }
template <> __device__ __forceinline__ P8<__nv_bfloat16> p8_convert<__nv_bfloat16, float>(const P8<float>& v) {
  P8<__nv_bfloat16> r;
  r.u = make_uint4(pack_bf16x2(v.a.x, v.a.y), pack_bf16x2(v.a.z, v.a.w), pack_bf16x2(v.b.x, v.b.y), pack_bf16x2(v.b.z, v.b.w));
  return r;
}
template <> __device__ __forceinline__ P8<float> p8_convert<float, __nv_bfloat16>(const P8<__nv_bfloat16>& v) {
  P8<float> r;
  r.a = make_float4(__uint_as_float(v.u.x << 16), __uint_as_float(v.u.x & 0xffff0000u), __uint_as_float(v.u.y << 16), __uint_as_float(v.u.y & 0xffff0000u));
  r.b = make_float4(__uint_as_float(v.u.z << 16), __uint_as_float(v.u.z & 0xffff0000u), __uint_as_float(v.u.w << 16), __uint_as_float(v.u.w & 0xffff0000u));
  return r;
}

// Thread map of both fill kernels (Cs = 128: 16 packets per (cell, scale), 48 per cell): 12 warps = 4 cell pairs x 3 scales,
// a warp = 2 adjacent cells x 16 packets of ONE scale, so the scale - hence every shift and the rank-grid row - is
// warp-uniform and a k > 1 site is looked up once for both cells.  grid = map rows; a CTA walks a row 8 cells per trip,
// four trips in flight; the row's 8 x 768 B are contiguous.  (r2: the earlier map - 48 consecutive threads per cell - mixed
// scales inside a warp and spent ~80 instructions per 16-byte packet on index arithmetic: 360 us forward for a 1.35 GB
// bf16 map whichever type the rows had.)
#define FILL_C8 16
#define FILL_PER_CELL 48
struct FillLane {
  int c, sub, s, xp, sh, k, Ws, suby;
  bool row_ok;
  const int* grow;
  __device__ __forceinline__ void init(const DenseFillArgs& a, int b, int y) {
    c = threadIdx.x & 15; sub = (threadIdx.x >> 4) & 1;
    const int warp = threadIdx.x >> 5;
    s = warp % 3; xp = warp / 3;
    k = a.k[s];
    sh = k == 1 ? 0 : (k == 2 ? 1 : 2);
    const int gy = y >> sh;
    row_ok = gy < a.H[s];
    Ws = a.W[s];
    grow = a.grid[s] + ((long long)b * a.H[s] + gy) * Ws;
    suby = (y & (k - 1)) << sh;
  }
  // rank of the site covering cell x at this scale, or -1
  __device__ __forceinline__ int rank_of(int x, int X) const {
    const int gx = x >> sh;
    return (row_ok && x < X && gx < Ws) ? __ldg(grow + gx) : -1;
  }
  // packet index of (rank, cell x) inside the scale's rows
  __device__ __forceinline__ int row_packet(int rank, int x) const { return (((rank << (2 * sh)) + suby + (x & (k - 1))) << 4) + c; }
};

template <typename T, typename TR>
__global__ void __launch_bounds__(384) dense_fill_kernel(const __grid_constant__ DenseFillArgs a, T* __restrict__ out) {
  const int y = blockIdx.x, b = blockIdx.y;
  FillLane L;
  L.init(a, b, y);
  const TR* rows = reinterpret_cast<const TR*>(a.rows[L.s]);
  P8<float> bgf = P8<float>::load_keep(a.bg[L.s], L.c);
  const P8<T> bg = p8_convert<T, float>(bgf);
  T* orow = out + (((long long)b * a.Y + y) * a.X * FILL_PER_CELL + L.s * FILL_C8 + L.c) * 8;     // packet (x = 0, s, c)
  for (int x0 = 2 * L.xp + L.sub; x0 < a.X; x0 += 32) {
    int rank[4];
    P8<T> v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) rank[u] = L.rank_of(x0 + 8 * u, a.X);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = bg;
      if (rank[u] >= 0) v[u] = p8_convert<T, TR>(P8<TR>::load_keep(rows, L.row_packet(rank[u], x0 + 8 * u)));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (x0 + 8 * u < a.X) v[u].store_stream(orow, (x0 + 8 * u) * FILL_PER_CELL);
  }
}

// backward, ONE pass over dout (B, Y, X, 3*Cs) with the forward's walk: a covered (cell, scale) packet goes to its sparse
// row drows_s[rank*k*k + sub] (TR: bf16 rows are the map's own bits), an uncovered one is added to the thread's running sum
// for dbg_s.  dout is read exactly once, sequentially (r2 before this kernel: a gather kernel per scale plus a pass over
// the uncovered cells - 556 us for 1.35 GB).  grid-stride over the map rows; per-CTA sums meet in dbg by float atomics.
template <typename T, typename TR>
__global__ void __launch_bounds__(384, 3) dense_fill_bwd_kernel(const __grid_constant__ DenseFillArgs a, const T* __restrict__ dout, void* d0,
                                                             void* d1, void* d2, float* __restrict__ dbg) {
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
  const int n_rows = a.B * a.Y;
  FillLane L;
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int b = row / a.Y, y = row - b * a.Y;
    L.init(a, b, y);
    TR* drows = reinterpret_cast<TR*>(L.s == 0 ? d0 : (L.s == 1 ? d1 : d2));
    const T* irow = dout + ((long long)row * a.X * FILL_PER_CELL + L.s * FILL_C8 + L.c) * 8;
    for (int x0 = 2 * L.xp + L.sub; x0 < a.X; x0 += 32) {
      int rank[4];
      P8<T> v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        rank[u] = L.rank_of(x0 + 8 * u, a.X);
        if (x0 + 8 * u < a.X) v[u] = P8<T>::load_stream(irow, (x0 + 8 * u) * FILL_PER_CELL);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (rank[u] >= 0) p8_convert<TR, T>(v[u]).store_stream(drows, L.row_packet(rank[u], x0 + 8 * u));
        else if (x0 + 8 * u < a.X) v[u].add_to(acc0, acc1);
      }
    }
  }
  // the 8 lanes (4 cell pairs x 2 cells) that share (scale, packet) meet in shared memory
  __shared__ float4 red[2][384];
  red[0][threadIdx.x] = acc0;
  red[1][threadIdx.x] = acc1;
  __syncthreads();
  if (L.xp == 0 && L.sub == 0) {
    for (int j = 0; j < 4; ++j)
      for (int h = 0; h < 2; ++h) {
        if (j == 0 && h == 0) continue;
        const int t = (j * 3 + L.s) * 32 + h * 16 + L.c;
        const float4 v0 = red[0][t], v1 = red[1][t];
        acc0.x += v0.x; acc0.y += v0.y; acc0.z += v0.z; acc0.w += v0.w;
        acc1.x += v1.x; acc1.y += v1.y; acc1.z += v1.z; acc1.w += v1.w;
      }
    float* dst = dbg + L.s * a.Cs + 8 * L.c;
    atomicAdd(dst, acc0.x); atomicAdd(dst + 1, acc0.y); atomicAdd(dst + 2, acc0.z); atomicAdd(dst + 3, acc0.w);
    atomicAdd(dst + 4, acc1.x); atomicAdd(dst + 5, acc1.y); atomicAdd(dst + 6, acc1.z); atomicAdd(dst + 7, acc1.w);
  }
}

static int fill_args(DenseFillArgs& a, const void* const* rows, const float* const* bg, const int32_t* const* grids,
                     const int* strides, int B, int Y, int X, int Cs) {
  GDMAE_CHECK_ARG(B >= 1 && Y >= 1 && X >= 1 && Cs > 0 && (Cs % 8) == 0 && 256 % (Cs / 8) == 0);
  GDMAE_CHECK_ARG(Cs == 128 && B < 65536);  // block = 8 cells x 3 scales x 16 lanes
  for (int s = 0; s < 3; ++s) {
    GDMAE_CHECK_ARG(strides[s] == 1 || strides[s] == 2 || strides[s] == 4);
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

#define FILL_DISPATCH2(A, B, CALL)                                                 \
  do {                                                                             \
    if ((A) == 0 && (B) == 0) { using T0 = float; using T1 = float; CALL; }        \
    else if ((A) == 0) { using T0 = float; using T1 = __nv_bfloat16; CALL; }       \
    else if ((B) == 0) { using T0 = __nv_bfloat16; using T1 = float; CALL; }       \
    else { using T0 = __nv_bfloat16; using T1 = __nv_bfloat16; CALL; }             \
  } while (0)

// rows_dtype / out_dtype / dtype / drows_dtype: 0 = fp32, 1 = bf16 (the dense BEV map feeds the decoder conv)
extern "C" int gdmae_dense_fill(const void* const* rows, int rows_dtype, const float* const* bg, const int32_t* const* rank_grids,
                                const int* strides, int B, int Y, int X, int Cs, void* out, int out_dtype, void* stream_) {
  GDMAE_CHECK_ARG((rows_dtype == 0 || rows_dtype == 1) && (out_dtype == 0 || out_dtype == 1));
  DenseFillArgs a;
  int rc = fill_args(a, rows, bg, rank_grids, strides, B, Y, X, Cs);
  if (rc) return rc;
  dim3 grid(Y, B);
  FILL_DISPATCH2(out_dtype, rows_dtype, (dense_fill_kernel<T0, T1><<<grid, 384, 0, (cudaStream_t)stream_>>>(a, (T0*)out)));
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// drows[s] (n_sites[s] * k_s^2, Cs) in drows_dtype, dbg (3*Cs) fp32.  Sub-cells of boundary sites that fall outside the map
// (odd lattice sizes) are never visited by the pass over the map: those scales are zero-filled first.
extern "C" int gdmae_dense_fill_bwd(const void* dout, int dtype, const int32_t* const* rank_grids, const int64_t* n_sites,
                                    const int* strides, int B, int Y, int X, int Cs, void* const* drows, int drows_dtype,
                                    float* dbg /* (3*Cs) */, void* stream_) {
  GDMAE_CHECK_ARG((dtype == 0 || dtype == 1) && (drows_dtype == 0 || drows_dtype == 1));
  cudaStream_t st = (cudaStream_t)stream_;
  DenseFillArgs a;
  int rc = fill_args(a, nullptr, nullptr, rank_grids, strides, B, Y, X, Cs);
  if (rc) return rc;
  GDMAE_CHECK_CUDA(cudaMemsetAsync(dbg, 0, (size_t)3 * Cs * 4, st));
  for (int s = 0; s < 3; ++s)
    if ((a.H[s] * a.k[s] > Y || a.W[s] * a.k[s] > X) && n_sites[s] > 0)
      GDMAE_CHECK_CUDA(cudaMemsetAsync(drows[s], 0, (size_t)n_sites[s] * a.k[s] * a.k[s] * Cs * (drows_dtype ? 2 : 4), st));
  const int n_rows = B * Y;
  const int grid = n_rows < GDMAE_NUM_SMS * 5 ? n_rows : GDMAE_NUM_SMS * 5;      // 5 CTAs of 384 threads per SM
  FILL_DISPATCH2(dtype, drows_dtype,
                 (dense_fill_bwd_kernel<T0, T1><<<grid, 384, 0, st>>>(a, (const T0*)dout, drows[0], drows[1], drows[2], dbg)));
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
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
