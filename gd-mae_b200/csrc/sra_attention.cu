// Sparse Regional Attention core (cosine multi-head attention inside 8x8 pillar windows), forward
// and backward, on the FLAT token layout.
//
// Replaces (reference file:line, relative to /root/reference):
//   sst_utils.flat2window_v2 / window2flat_v2              pcdet/models/model_utils/sst_utils.py:107-181
//   WindowAttention.forward (per drop-level loop)          pcdet/models/model_utils/sst_basic_block.py:22-54
//   _scaled_cosine_attention                               pcdet/models/model_utils/cosine_msa.py:114-176
//   key-padding mask handling                              pcdet/models/model_utils/cosine_msa.py:404-420
//
// The reference pads every window to 16/32/64 tokens (fill 0.20 in MAE mode), materialises q, k,
// v, the (nWin*8, T, T) score tensor, its softmax and a head-averaged copy of it, per drop level.
// Here nothing is padded: tokens are visited in window (CSR) order, the positional embedding
// enters as a 64-row look-up table already multiplied by Wq / Wk (pos has only 64 distinct
// rows), softmax is computed online.  HBM traffic = qkv in + o out (+ the L2-resident LUT):
// N*d*(3+1)*4 + N*8 bytes, the algorithmic minimum of SURVEY.md 8d.
//
// Work decomposition (v4): the CSR row order is cut into bins of SRA_BIN rows; a CTA owns the
// windows that START inside its bin (<= BIN+63 rows because a window holds <= 64 tokens) and a
// 128-channel slice of the embedding (all 8 heads for d=128, 4 of 8 heads for d=256).
// The kernel is organised around its dependent-load chain (r1 ncu: long-scoreboard stalls
// dominated v3), which is now two round trips deep:
//   trip 1  one coalesced read of the 16-byte row_info records of rows [bin, bin+BIN+64) - a
//           superset of the CTA's rows, so the bin bounds need no separate lookup;
//   trip 2  cp.async (LDGSTS) copies of the raw k / v (or q / dO) rows straight into shared
//           memory, all in flight at once, while every thread loads its own operand rows
//           (q, or k and v) into registers;
//   then    in-place "+LUT, L2-normalise per head" pass over the staged rows, and the partner
//           loop: each (row, head) pair is one thread that streams its window partners from
//           shared memory (pitch SLICE+4 floats: rows of different windows hit different banks,
//           partners of the same window are a broadcast).  No global load inside the loop.
#include "common.cuh"
#include <cuda_bf16.h>

#define SRA_EPS 1e-12f
// tunables (overridable with -D for tools/bench_sra.py sweeps)
#ifndef SRA_BIN
#define SRA_BIN 64            // CSR rows per bin            (r1 sweep, tools/bench_sra.py: 64/64 is the best
#endif                        //                              of {16,32,64} x {64,128} on all three scales)
#ifndef SRA_SLICE
#define SRA_SLICE 64          // channels per CTA
#endif
#ifndef SRA_FWD_THREADS
#define SRA_FWD_THREADS 256
#endif
#ifndef SRA_BWD_THREADS
#define SRA_BWD_THREADS 128
#endif
#ifndef SRA_MIN_CTAS
#define SRA_MIN_CTAS 2
#endif
#define SRA_ROWS (SRA_BIN + 64)        // >= SRA_BIN + 63
#define SRA_C4 (SRA_SLICE / 4)         // float4 per staged row
#define SRA_PITCH (SRA_SLICE + 4)      // floats per staged row (+4: bank shift of 4 per row)
#define SRA_SMEM_BYTES (2 * SRA_ROWS * SRA_PITCH * 4 + SRA_ROWS * 16 + 2 * SRA_ROWS * 8 * 4 + 64)

struct SraArgs {
  const float* qkv;      // (N, 3d): x Wq^T | x Wk^T | x Wv^T + bv   (q, k without bias / pos term)
  const float* lut;      // (64, 2d): pos_table Wq^T + bq | pos_table Wk^T + bk
  const int4* row_info;  // (N) per CSR row: token, first row of its window, one past its last row, in-window cell
  const float* tau;      // (1) learnable temperature
  const float* bv;       // (d) value bias, nullable: o = sum_j p_j (v_j + bv) = sum_j p_j v_j + bv (qkv holds v WITHOUT bias)
  float tau_min;
  int N, d;
  int io_bf16;           // 1: the attention output o and dqkv are bf16 (GEMM operands of the bf16 configuration)
};

struct SraSmem {
  float* a;     // [SRA_ROWS][SRA_PITCH]  k-hat   (fwd, bwd_q)   | q-hat (bwd_kv)
  float* b;     // [SRA_ROWS][SRA_PITCH]  v       (fwd, bwd_q)   | dO    (bwd_kv)
  int4* info;   // [SRA_ROWS] raw row_info of rows bin .. bin+SRA_ROWS-1 (absolute row numbers)
  float* lse;   // [SRA_ROWS][8]  (bwd_kv)
  float* D;     // [SRA_ROWS][8]  (bwd_kv)
};

__device__ __forceinline__ SraSmem sra_carve(unsigned char* base) {
  SraSmem s;
  s.a = (float*)base;
  s.b = s.a + SRA_ROWS * SRA_PITCH;
  s.info = (int4*)(s.b + SRA_ROWS * SRA_PITCH);
  s.lse = (float*)(s.info + SRA_ROWS);
  s.D = s.lse + SRA_ROWS * 8;
  return s;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Trip 1: records of rows [bin, bin + SRA_ROWS) -> smem; returns the CTA's row range [row0, row0+R)
// (the windows that start inside the bin).  info is indexed with (absolute row - bin).
__device__ __forceinline__ void sra_bin_setup(const SraArgs& a, const SraSmem& s, int& row0, int& R) {
  const int bin = blockIdx.x * SRA_BIN;
  for (int r = threadIdx.x; r < SRA_ROWS; r += blockDim.x) {
    int p = bin + r;
    s.info[r] = p < a.N ? __ldg(a.row_info + p) : make_int4(0, a.N, a.N, 0);
  }
  __syncthreads();
  int4 f = s.info[0];                       // first window start >= bin
  row0 = (f.y == bin) ? bin : f.z;
  int p1 = bin + SRA_BIN;                   // first window start >= bin + BIN
  int row1 = a.N;
  if (p1 < a.N) {
    int4 l = s.info[SRA_BIN];
    row1 = (l.y == p1) ? p1 : l.z;
  }
  R = row1 - row0;
}

// Trip 2: raw 16-byte pieces of srcA[:, colA] and srcB[:, colB] for rows [row0, row0+R) -> smem (async).
__device__ __forceinline__ void sra_stage_async(const SraSmem& s, int row0, int R, const float* __restrict__ srcA,
                                                long long strideA, int colA, const float* __restrict__ srcB, long long strideB,
                                                int colB) {
  const int shift = row0 - blockIdx.x * SRA_BIN;
  for (int idx = threadIdx.x; idx < R * SRA_C4; idx += blockDim.x) {
    int r = idx / SRA_C4, c4 = idx % SRA_C4;
    long long tok = s.info[r + shift].x;
    cp_async16(s.a + r * SRA_PITCH + 4 * c4, srcA + tok * strideA + colA + 4 * c4);
    cp_async16(s.b + r * SRA_PITCH + 4 * c4, srcB + tok * strideB + colB + 4 * c4);
  }
}

// In place: a[r] = normalise_per_head(a[r] + lut[pos(r), lutcol : lutcol + SLICE])
template <int HD>
__device__ __forceinline__ void sra_normalise(const SraSmem& s, int row0, int R, const float* __restrict__ lut, int lut_stride,
                                              int lutcol) {
  constexpr int LPH = HD / 4;  // lanes per head
  const int shift = row0 - blockIdx.x * SRA_BIN;
  int steps = (R * SRA_C4 + blockDim.x - 1) / blockDim.x;
  for (int it = 0; it < steps; ++it) {
    int idx = it * blockDim.x + threadIdx.x;
    int r = idx / SRA_C4, c4 = idx % SRA_C4;
    bool valid = r < R;
    int rr = valid ? r : 0;
    float4 k = *reinterpret_cast<const float4*>(s.a + rr * SRA_PITCH + 4 * c4);
    float4 l = __ldg(reinterpret_cast<const float4*>(lut + s.info[rr + shift].w * lut_stride + lutcol) + c4);
    k.x += l.x; k.y += l.y; k.z += l.z; k.w += l.w;
    float ss = k.x * k.x + k.y * k.y + k.z * k.z + k.w * k.w;
#pragma unroll
    for (int o = 1; o < LPH; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    float sc = 1.f / fmaxf(sqrtf(ss), SRA_EPS);
    if (valid) *reinterpret_cast<float4*>(s.a + r * SRA_PITCH + 4 * c4) = make_float4(k.x * sc, k.y * sc, k.z * sc, k.w * sc);
  }
}

template <int HD>
__device__ __forceinline__ void load_head(const float* __restrict__ p, float* v) {
  const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) {
    float4 t = __ldg(p4 + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
template <int HD>
__device__ __forceinline__ void add_head(const float* __restrict__ p, float* v) {
  const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) {
    float4 t = __ldg(p4 + i);
    v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
  }
}
template <int HD>
__device__ __forceinline__ void store_head(float* __restrict__ p, const float* v) {
  float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) p4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// head row stored as fp32 or bf16 (elem = element offset of the head's first channel)
template <int HD>
__device__ __forceinline__ void store_head_t(void* base, int bf16, long long elem, const float* v) {
  if (!bf16) { store_head<HD>((float*)base + elem, v); return; }
  uint4* p = reinterpret_cast<uint4*>((__nv_bfloat16*)base + elem);
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]), b = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]), e = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
    uint4 u;
    u.x = *reinterpret_cast<unsigned*>(&a); u.y = *reinterpret_cast<unsigned*>(&b);
    u.z = *reinterpret_cast<unsigned*>(&c); u.w = *reinterpret_cast<unsigned*>(&e);
    p[i] = u;
  }
}
template <int HD>
__device__ __forceinline__ void load_head_t(const void* base, int bf16, long long elem, float* v) {
  if (!bf16) { load_head<HD>((const float*)base + elem, v); return; }
  const uint4* p = reinterpret_cast<const uint4*>((const __nv_bfloat16*)base + elem);
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    uint4 u = __ldg(p + i);
    float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.y));
    float2 c = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.z)), e = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.w));
    v[8 * i] = a.x; v[8 * i + 1] = a.y; v[8 * i + 2] = b.x; v[8 * i + 3] = b.y;
    v[8 * i + 4] = c.x; v[8 * i + 5] = c.y; v[8 * i + 6] = e.x; v[8 * i + 7] = e.y;
  }
}
template <int HD>
__device__ __forceinline__ float dot_smem(const float* __restrict__ sm, const float* r) {
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) {
    float4 t = *reinterpret_cast<const float4*>(sm + 4 * i);
    acc0 = fmaf(r[4 * i], t.x, acc0); acc1 = fmaf(r[4 * i + 1], t.y, acc1);
    acc0 = fmaf(r[4 * i + 2], t.z, acc0); acc1 = fmaf(r[4 * i + 3], t.w, acc1);
  }
  return acc0 + acc1;
}
template <int HD>
__device__ __forceinline__ void axpy_smem(float a, const float* __restrict__ sm, float* r) {
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) {
    float4 t = *reinterpret_cast<const float4*>(sm + 4 * i);
    r[4 * i] = fmaf(a, t.x, r[4 * i]); r[4 * i + 1] = fmaf(a, t.y, r[4 * i + 1]);
    r[4 * i + 2] = fmaf(a, t.z, r[4 * i + 2]); r[4 * i + 3] = fmaf(a, t.w, r[4 * i + 3]);
  }
}
template <int HD>
__device__ __forceinline__ float dot_reg(const float* a, const float* b) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < HD; ++i) acc = fmaf(a[i], b[i], acc);
  return acc;
}

// ------------------------------------------------------------------------------ forward
template <int HD>
__global__ void __launch_bounds__(SRA_FWD_THREADS, SRA_MIN_CTAS) sra_fwd_kernel(SraArgs a, void* __restrict__ out,
                                                                              float* __restrict__ lse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SraSmem s = sra_carve(smem_raw);
  constexpr int HS = SRA_SLICE / HD;  // heads per slice
  const int d = a.d, col = blockIdx.y * SRA_SLICE;
  int row0, R;
  sra_bin_setup(a, s, row0, R);
  if (R == 0) return;
  const int shift = row0 - blockIdx.x * SRA_BIN;
  sra_stage_async(s, row0, R, a.qkv, 3 * d, d + col, a.qkv, 3 * d, 2 * d + col);
  // this thread's first (row, head) pair: q row + LUT row are fetched while the copies fly
  const int npairs = R * HS;
  float q[HD];
  int4 inf = make_int4(0, 0, 0, 0);
  if ((int)threadIdx.x < npairs) {
    int r = threadIdx.x % R, h = threadIdx.x / R;
    inf = s.info[r + shift];
    load_head<HD>(a.qkv + (long long)inf.x * 3 * d + col + h * HD, q);
    add_head<HD>(a.lut + inf.w * 2 * d + col + h * HD, q);
  }
  const float inv_tau = 1.f / fmaxf(__ldg(a.tau), a.tau_min);
  cp_async_wait_all();
  __syncthreads();
  sra_normalise<HD>(s, row0, R, a.lut, 2 * d, d + col);
  __syncthreads();
  for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
    int r = p % R, h = p / R;
    if (p != (int)threadIdx.x) {  // later passes: fetch now
      inf = s.info[r + shift];
      load_head<HD>(a.qkv + (long long)inf.x * 3 * d + col + h * HD, q);
      add_head<HD>(a.lut + inf.w * 2 * d + col + h * HD, q);
    }
    float qs = inv_tau / fmaxf(sqrtf(dot_reg<HD>(q, q)), SRA_EPS);
#pragma unroll
    for (int i = 0; i < HD; ++i) q[i] *= qs;
    float m = -INFINITY, l = 0.f, o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;
    const int je = inf.z - row0;
    for (int j = inf.y - row0; j < je; ++j) {
      float sc = dot_smem<HD>(s.a + j * SRA_PITCH + h * HD, q);
      float mn = fmaxf(m, sc);
      float corr = __expf(m - mn);
      float pj = __expf(sc - mn);
      l = fmaf(l, corr, pj);
#pragma unroll
      for (int i = 0; i < HD; ++i) o[i] *= corr;
      axpy_smem<HD>(pj, s.b + j * SRA_PITCH + h * HD, o);
      m = mn;
    }
    float il = 1.f / l;
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] *= il;
    if (a.bv) add_head<HD>(a.bv + col + h * HD, o);
    store_head_t<HD>(out, a.io_bf16, (long long)inf.x * d + col + h * HD, o);
    lse[(long long)inf.x * 8 + blockIdx.y * HS + h] = m + __logf(l);
  }
}

// ------------------------------------------------------------------------------ backward, query side
// dq_t and D_t = dO_t . O_t ; accumulates sum_ij dS_ij S_ij for the temperature gradient.
template <int HD>
__global__ void __launch_bounds__(SRA_BWD_THREADS, SRA_MIN_CTAS) sra_bwd_q_kernel(SraArgs a, const void* __restrict__ out,
                                                                                const float* __restrict__ lse,
                                                                                const float* __restrict__ dout,
                                                                                void* __restrict__ dqkv, float* __restrict__ Dbuf,
                                                                                double* __restrict__ dtau_acc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SraSmem s = sra_carve(smem_raw);
  constexpr int HS = SRA_SLICE / HD;
  const int d = a.d, col = blockIdx.y * SRA_SLICE;
  int row0, R;
  sra_bin_setup(a, s, row0, R);
  if (R == 0) return;
  const int shift = row0 - blockIdx.x * SRA_BIN;
  sra_stage_async(s, row0, R, a.qkv, 3 * d, d + col, a.qkv, 3 * d, 2 * d + col);
  const float inv_tau = 1.f / fmaxf(__ldg(a.tau), a.tau_min);
  const int npairs = R * HS;
  float q[HD], dO[HD];
  float Dt = 0.f, ls = 0.f;
  int4 inf = make_int4(0, 0, 0, 0);
  auto fetch = [&](int p) {
    int r = p % R, h = p / R;
    inf = s.info[r + shift];
    load_head<HD>(a.qkv + (long long)inf.x * 3 * d + col + h * HD, q);
    add_head<HD>(a.lut + inf.w * 2 * d + col + h * HD, q);
    load_head<HD>(dout + (long long)inf.x * d + col + h * HD, dO);
    float o[HD];
    load_head_t<HD>(out, a.io_bf16, (long long)inf.x * d + col + h * HD, o);
    if (a.bv) {  // the staged v rows carry no bias: D must be taken against o - bv
      float b[HD];
      load_head<HD>(a.bv + col + h * HD, b);
#pragma unroll
      for (int i = 0; i < HD; ++i) o[i] -= b[i];
    }
    Dt = dot_reg<HD>(dO, o);
    ls = lse[(long long)inf.x * 8 + blockIdx.y * HS + h];
  };
  if ((int)threadIdx.x < npairs) fetch(threadIdx.x);
  cp_async_wait_all();
  __syncthreads();
  sra_normalise<HD>(s, row0, R, a.lut, 2 * d, d + col);
  __syncthreads();
  float tacc = 0.f;
  for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
    int h = p / R;
    if (p != (int)threadIdx.x) fetch(p);
    float iqn = 1.f / fmaxf(sqrtf(dot_reg<HD>(q, q)), SRA_EPS);
#pragma unroll
    for (int i = 0; i < HD; ++i) q[i] *= iqn;  // q-hat
    float dq[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) dq[i] = 0.f;
    const int je = inf.z - row0;
    for (int j = inf.y - row0; j < je; ++j) {
      const float* kj = s.a + j * SRA_PITCH + h * HD;
      float sc = dot_smem<HD>(kj, q) * inv_tau;
      float dp = dot_smem<HD>(s.b + j * SRA_PITCH + h * HD, dO);
      float pj = __expf(sc - ls);
      float ds = pj * (dp - Dt);
      tacc = fmaf(ds, sc, tacc);
      axpy_smem<HD>(ds * inv_tau, kj, dq);
    }
    float proj = dot_reg<HD>(q, dq);
#pragma unroll
    for (int i = 0; i < HD; ++i) dq[i] = (dq[i] - q[i] * proj) * iqn;
    store_head_t<HD>(dqkv, a.io_bf16, (long long)inf.x * 3 * d + col + h * HD, dq);
    Dbuf[(long long)inf.x * 8 + blockIdx.y * HS + h] = Dt;
  }
  tacc = warp_sum(tacc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    if (t != 0.f) atomicAdd(dtau_acc, (double)t);
  }
}

// ------------------------------------------------------------------------------ backward, key/value side
template <int HD>
__global__ void __launch_bounds__(SRA_BWD_THREADS, SRA_MIN_CTAS) sra_bwd_kv_kernel(SraArgs a, const float* __restrict__ lse,
                                                                                 const float* __restrict__ dout,
                                                                                 const float* __restrict__ Dbuf,
                                                                                 void* __restrict__ dqkv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SraSmem s = sra_carve(smem_raw);
  constexpr int HS = SRA_SLICE / HD;
  const int d = a.d, col = blockIdx.y * SRA_SLICE;
  int row0, R;
  sra_bin_setup(a, s, row0, R);
  if (R == 0) return;
  const int shift = row0 - blockIdx.x * SRA_BIN;
  // stationary operands of this pass: q (normalised below) and dO
  sra_stage_async(s, row0, R, a.qkv, 3 * d, col, dout, d, col);
  const int npairs = R * HS;
  for (int i = threadIdx.x; i < npairs; i += blockDim.x) {
    int r = i / HS, h = i % HS;
    long long g = (long long)s.info[r + shift].x * 8 + blockIdx.y * HS + h;
    s.lse[r * 8 + h] = lse[g];
    s.D[r * 8 + h] = Dbuf[g];
  }
  const float inv_tau = 1.f / fmaxf(__ldg(a.tau), a.tau_min);
  float k[HD], v[HD];
  int4 inf = make_int4(0, 0, 0, 0);
  auto fetch = [&](int p) {
    int r = p % R, h = p / R;
    inf = s.info[r + shift];
    load_head<HD>(a.qkv + (long long)inf.x * 3 * d + d + col + h * HD, k);
    add_head<HD>(a.lut + inf.w * 2 * d + d + col + h * HD, k);
    load_head<HD>(a.qkv + (long long)inf.x * 3 * d + 2 * d + col + h * HD, v);
  };
  if ((int)threadIdx.x < npairs) fetch(threadIdx.x);
  cp_async_wait_all();
  __syncthreads();
  sra_normalise<HD>(s, row0, R, a.lut, 2 * d, col);
  __syncthreads();
  for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
    int h = p / R;
    if (p != (int)threadIdx.x) fetch(p);
    float dk[HD], dv[HD];
    float ikn = 1.f / fmaxf(sqrtf(dot_reg<HD>(k, k)), SRA_EPS);
#pragma unroll
    for (int i = 0; i < HD; ++i) { k[i] *= ikn; dk[i] = 0.f; dv[i] = 0.f; }
    const int ie = inf.z - row0;
    for (int i = inf.y - row0; i < ie; ++i) {
      const float* qi = s.a + i * SRA_PITCH + h * HD;
      const float* doi = s.b + i * SRA_PITCH + h * HD;
      float sc = dot_smem<HD>(qi, k) * inv_tau;
      float dp = dot_smem<HD>(doi, v);
      float pj = __expf(sc - s.lse[i * 8 + h]);
      float ds = pj * (dp - s.D[i * 8 + h]);
      axpy_smem<HD>(ds * inv_tau, qi, dk);
      axpy_smem<HD>(pj, doi, dv);
    }
    float proj = dot_reg<HD>(k, dk);
#pragma unroll
    for (int i = 0; i < HD; ++i) dk[i] = (dk[i] - k[i] * proj) * ikn;
    store_head_t<HD>(dqkv, a.io_bf16, (long long)inf.x * 3 * d + d + col + h * HD, dk);
    store_head_t<HD>(dqkv, a.io_bf16, (long long)inf.x * 3 * d + 2 * d + col + h * HD, dv);
  }
}

static int sra_check(int64_t N, int d, int nhead, const void* row_info) {
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 27));
  GDMAE_CHECK_ARG(nhead == 8 && (d == 128 || d == 256));
  GDMAE_CHECK_ARG(((uintptr_t)row_info % 16) == 0);
  return GDMAE_OK;
}

// >48 KB of dynamic shared memory needs the opt-in attribute on every kernel instantiation
static int sra_smem_attrs() {
  static bool done = false;
  if (done) return GDMAE_OK;
#define SRA_ATTR(k) GDMAE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SRA_SMEM_BYTES))
  SRA_ATTR(sra_fwd_kernel<16>); SRA_ATTR(sra_fwd_kernel<32>);
  SRA_ATTR(sra_bwd_q_kernel<16>); SRA_ATTR(sra_bwd_q_kernel<32>);
  SRA_ATTR(sra_bwd_kv_kernel<16>); SRA_ATTR(sra_bwd_kv_kernel<32>);
#undef SRA_ATTR
  done = true;
  return GDMAE_OK;
}

// o (N,d) = softmax_j( cos(q_i, k_j) / max(tau, tau_min) ) v_j over the tokens j of i's window; lse (N,8).
// bv (d, nullable): value bias added to o.  io_bf16: out is bf16 instead of fp32.
extern "C" int gdmae_sra_attention_fwd(const float* qkv, const float* lut, const int32_t* row_info, int64_t N, int d, int nhead,
                                       const float* tau, float tau_min, const float* bv, int io_bf16, void* out, float* lse,
                                       void* stream_) {
  int rc = sra_check(N, d, nhead, row_info);
  if (rc) return rc;
  if (N == 0) return GDMAE_OK;
  SraArgs a{qkv, lut, (const int4*)row_info, tau, bv, tau_min, (int)N, d, io_bf16};
  cudaStream_t st = (cudaStream_t)stream_;
  dim3 grid(gdmae_div_up(N, SRA_BIN), d / SRA_SLICE);
  if ((rc = sra_smem_attrs())) return rc;
  if (d == 128) sra_fwd_kernel<16><<<grid, SRA_FWD_THREADS, SRA_SMEM_BYTES, st>>>(a, out, lse);
  else sra_fwd_kernel<32><<<grid, SRA_FWD_THREADS, SRA_SMEM_BYTES, st>>>(a, out, lse);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// dqkv (N,3d) = [dq | dk | dv]; dtau_sum (1, double, caller zeroes) accumulates sum dS*S
// (d loss / d tau = -dtau_sum / tau_c when tau >= tau_min, else 0); work_D (N,8) scratch.
// io_bf16: `out` (the forward output) and `dqkv` are bf16.
extern "C" int gdmae_sra_attention_bwd(const float* qkv, const float* lut, const int32_t* row_info, int64_t N, int d, int nhead,
                                       const float* tau, float tau_min, const float* bv, int io_bf16, const void* out,
                                       const float* lse, const float* dout, void* dqkv, double* dtau_sum, float* work_D,
                                       void* stream_) {
  int rc = sra_check(N, d, nhead, row_info);
  if (rc) return rc;
  if (N == 0) return GDMAE_OK;
  SraArgs a{qkv, lut, (const int4*)row_info, tau, bv, tau_min, (int)N, d, io_bf16};
  cudaStream_t st = (cudaStream_t)stream_;
  dim3 grid(gdmae_div_up(N, SRA_BIN), d / SRA_SLICE);
  if ((rc = sra_smem_attrs())) return rc;
  if (d == 128) {
    sra_bwd_q_kernel<16><<<grid, SRA_BWD_THREADS, SRA_SMEM_BYTES, st>>>(a, out, lse, dout, dqkv, work_D, dtau_sum);
    GDMAE_LAUNCH_CHECK();
    sra_bwd_kv_kernel<16><<<grid, SRA_BWD_THREADS, SRA_SMEM_BYTES, st>>>(a, lse, dout, work_D, dqkv);
  } else {
    sra_bwd_q_kernel<32><<<grid, SRA_BWD_THREADS, SRA_SMEM_BYTES, st>>>(a, out, lse, dout, dqkv, work_D, dtau_sum);
    GDMAE_LAUNCH_CHECK();
    sra_bwd_kv_kernel<32><<<grid, SRA_BWD_THREADS, SRA_SMEM_BYTES, st>>>(a, lse, dout, work_D, dqkv);
  }
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
