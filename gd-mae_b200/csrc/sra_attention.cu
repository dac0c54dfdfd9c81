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
// Work decomposition (v3): the CSR row order is cut into bins of 32 rows; a CTA owns the windows
// that START inside its bin (<= 95 rows because a window holds <= 64 tokens) and a 128-channel
// slice of the embedding (all 8 heads for d=128, 4 of 8 heads for d=256).  Everything a CTA must
// know about a row (token, window extent, in-window cell) is ONE 16-byte row_info record written
// by gdmae_window_table, so the bin bounds cost two dependent loads instead of a search.  The
// "stationary" operands of the bin (k-hat and v for the forward / dq pass, q-hat and dO for the
// dk,dv pass) are staged ONCE in shared memory with coalesced 128-bit loads (pitch 132 floats so
// that rows of different windows fall into different banks); every (row, head) pair is then one
// thread that keeps its own operand in registers and streams its window partners from shared
// memory - lanes of a warp are consecutive rows, so partners of the same window are a broadcast.
// No global load sits inside the partner loop.  107 KB of shared memory per CTA -> two CTAs per
// SM overlap one CTA's staging latency with the other's partner loop.
#include "common.cuh"

#define SRA_EPS 1e-12f
#define SRA_BIN 32
#define SRA_ROWS 96           // >= SRA_BIN + 63
#define SRA_SLICE 128         // channels per CTA
#define SRA_PITCH 132         // floats per staged row (128 + 4: bank shift of 4 per row)
#define SRA_SMEM_BYTES (2 * SRA_ROWS * SRA_PITCH * 4 + SRA_ROWS * 16 + 2 * SRA_ROWS * 8 * 4 + 64)

struct SraArgs {
  const float* qkv;      // (N, 3d): x Wq^T | x Wk^T | x Wv^T + bv   (q, k without bias / pos term)
  const float* lut;      // (64, 2d): pos_table Wq^T + bq | pos_table Wk^T + bk
  const int4* row_info;  // (N) per CSR row: token, first row of its window, one past its last row, in-window cell
  const float* tau;      // (1) learnable temperature
  float tau_min;
  int N, d;
};

struct SraSmem {
  float* a;     // [SRA_ROWS][SRA_PITCH]  k-hat   (fwd, bwd_q)   | q-hat (bwd_kv)
  float* b;     // [SRA_ROWS][SRA_PITCH]  v       (fwd, bwd_q)   | dO    (bwd_kv)
  int4* info;   // [SRA_ROWS] row_info with the window extent made bin-relative
  float* lse;   // [SRA_ROWS][8]  (bwd_kv)
  float* D;     // [SRA_ROWS][8]  (bwd_kv)
  int* hdr;     // [0] row0, [1] R
};

__device__ __forceinline__ SraSmem sra_carve(unsigned char* base) {
  SraSmem s;
  s.a = (float*)base;
  s.b = s.a + SRA_ROWS * SRA_PITCH;
  s.info = (int4*)(s.b + SRA_ROWS * SRA_PITCH);
  s.lse = (float*)(s.info + SRA_ROWS);
  s.D = s.lse + SRA_ROWS * 8;
  s.hdr = (int*)(s.D + SRA_ROWS * 8);
  return s;
}

// first window start >= row t (t < N): t itself if row t opens a window, else the end of its window
__device__ __forceinline__ int sra_first_start(const int4* __restrict__ info, int t, int N) {
  if (t >= N) return N;
  int4 r = __ldg(info + t);
  return r.y == t ? t : r.z;
}

// Rows of the bin + per-row window extents.  Returns R (0 -> nothing to do).  Ends with __syncthreads.
__device__ __forceinline__ int sra_bin_setup(const SraArgs& a, const SraSmem& s) {
  if (threadIdx.x < 2) {
    int t = (blockIdx.x + threadIdx.x) * SRA_BIN;
    s.hdr[threadIdx.x] = sra_first_start(a.row_info, t, a.N);
  }
  __syncthreads();
  int row0 = s.hdr[0], R = s.hdr[1] - row0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    int4 v = __ldg(a.row_info + row0 + r);
    v.y -= row0;
    v.z -= row0;
    s.info[r] = v;
  }
  __syncthreads();
  return R;
}

// Stage rows [0,R) x 128-channel slice:  dstA = normalise_per_head(srcA[:, colA] + lut[pos, lutcol]) ,
// dstB = srcB[:, colB].  One float4 per thread per step, a warp covers one row (32 float4 = 128 ch).
template <int HD>
__device__ __forceinline__ void sra_stage(const SraSmem& s, int R, const float* __restrict__ srcA, long long strideA, int colA,
                                          const float* __restrict__ lut, int lut_stride, int lutcol,
                                          const float* __restrict__ srcB, long long strideB, int colB) {
  constexpr int LPH = HD / 4;  // lanes per head
  int steps = (R * 32 + blockDim.x - 1) / blockDim.x;
  for (int it = 0; it < steps; ++it) {
    int idx = it * blockDim.x + threadIdx.x;
    int r = idx >> 5, c4 = idx & 31;
    bool valid = r < R;
    int4 inf = s.info[valid ? r : 0];
    float4 k = __ldg(reinterpret_cast<const float4*>(srcA + (long long)inf.x * strideA + colA) + c4);
    float4 l = __ldg(reinterpret_cast<const float4*>(lut + inf.w * lut_stride + lutcol) + c4);
    float4 v = __ldg(reinterpret_cast<const float4*>(srcB + (long long)inf.x * strideB + colB) + c4);
    k.x += l.x; k.y += l.y; k.z += l.z; k.w += l.w;
    float ss = k.x * k.x + k.y * k.y + k.z * k.z + k.w * k.w;
#pragma unroll
    for (int o = 1; o < LPH; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    float sc = 1.f / fmaxf(sqrtf(ss), SRA_EPS);
    if (valid) {
      *reinterpret_cast<float4*>(s.a + r * SRA_PITCH + 4 * c4) = make_float4(k.x * sc, k.y * sc, k.z * sc, k.w * sc);
      *reinterpret_cast<float4*>(s.b + r * SRA_PITCH + 4 * c4) = v;
    }
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
__device__ __forceinline__ void load_head_add(const float* __restrict__ p, const float* __restrict__ q, float* v) {
  const float4* p4 = reinterpret_cast<const float4*>(p);
  const float4* q4 = reinterpret_cast<const float4*>(q);
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) {
    float4 s = __ldg(p4 + i), t = __ldg(q4 + i);
    v[4 * i] = s.x + t.x; v[4 * i + 1] = s.y + t.y; v[4 * i + 2] = s.z + t.z; v[4 * i + 3] = s.w + t.w;
  }
}
template <int HD>
__device__ __forceinline__ void store_head(float* __restrict__ p, const float* v) {
  float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < HD / 4; ++i) p4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
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
__global__ void __launch_bounds__(256, 2) sra_fwd_kernel(SraArgs a, float* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SraSmem s = sra_carve(smem_raw);
  constexpr int HS = SRA_SLICE / HD;  // heads per slice
  const int d = a.d, col = blockIdx.y * SRA_SLICE;
  int R = sra_bin_setup(a, s);
  if (R == 0) return;
  sra_stage<HD>(s, R, a.qkv, 3 * d, d + col, a.lut, 2 * d, d + col, a.qkv, 3 * d, 2 * d + col);
  __syncthreads();
  float inv_tau = 1.f / fmaxf(__ldg(a.tau), a.tau_min);
  for (int p = threadIdx.x; p < R * HS; p += blockDim.x) {
    int r = p % R, h = p / R;
    int4 inf = s.info[r];
    float q[HD];
    load_head_add<HD>(a.qkv + (long long)inf.x * 3 * d + col + h * HD, a.lut + inf.w * 2 * d + col + h * HD, q);
    float qs = inv_tau / fmaxf(sqrtf(dot_reg<HD>(q, q)), SRA_EPS);
#pragma unroll
    for (int i = 0; i < HD; ++i) q[i] *= qs;
    float m = -INFINITY, l = 0.f, o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;
    for (int j = inf.y; j < inf.z; ++j) {
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
    store_head<HD>(out + (long long)inf.x * d + col + h * HD, o);
    lse[(long long)inf.x * 8 + blockIdx.y * HS + h] = m + __logf(l);
  }
}

// ------------------------------------------------------------------------------ backward, query side
// dq_t and D_t = dO_t . O_t ; accumulates sum_ij dS_ij S_ij for the temperature gradient.
template <int HD>
__global__ void __launch_bounds__(128, 2) sra_bwd_q_kernel(SraArgs a, const float* __restrict__ out, const float* __restrict__ lse,
                                                           const float* __restrict__ dout, float* __restrict__ dqkv,
                                                           float* __restrict__ Dbuf, double* __restrict__ dtau_acc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SraSmem s = sra_carve(smem_raw);
  constexpr int HS = SRA_SLICE / HD;
  const int d = a.d, col = blockIdx.y * SRA_SLICE;
  int R = sra_bin_setup(a, s);
  if (R == 0) return;
  sra_stage<HD>(s, R, a.qkv, 3 * d, d + col, a.lut, 2 * d, d + col, a.qkv, 3 * d, 2 * d + col);
  __syncthreads();
  float inv_tau = 1.f / fmaxf(__ldg(a.tau), a.tau_min);
  float tacc = 0.f;
  for (int p = threadIdx.x; p < R * HS; p += blockDim.x) {
    int r = p % R, h = p / R;
    int4 inf = s.info[r];
    int hg = blockIdx.y * HS + h;
    float q[HD], dO[HD], dq[HD];
    load_head_add<HD>(a.qkv + (long long)inf.x * 3 * d + col + h * HD, a.lut + inf.w * 2 * d + col + h * HD, q);
    float iqn = 1.f / fmaxf(sqrtf(dot_reg<HD>(q, q)), SRA_EPS);
#pragma unroll
    for (int i = 0; i < HD; ++i) q[i] *= iqn;  // q-hat
    load_head<HD>(dout + (long long)inf.x * d + col + h * HD, dO);
    float Dt;
    {
      float o[HD];
      load_head<HD>(out + (long long)inf.x * d + col + h * HD, o);
      Dt = dot_reg<HD>(dO, o);
    }
    float ls = lse[(long long)inf.x * 8 + hg];
#pragma unroll
    for (int i = 0; i < HD; ++i) dq[i] = 0.f;
    for (int j = inf.y; j < inf.z; ++j) {
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
    store_head<HD>(dqkv + (long long)inf.x * 3 * d + col + h * HD, dq);
    Dbuf[(long long)inf.x * 8 + hg] = Dt;
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
__global__ void __launch_bounds__(128, 2) sra_bwd_kv_kernel(SraArgs a, const float* __restrict__ lse, const float* __restrict__ dout,
                                                            const float* __restrict__ Dbuf, float* __restrict__ dqkv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SraSmem s = sra_carve(smem_raw);
  constexpr int HS = SRA_SLICE / HD;
  const int d = a.d, col = blockIdx.y * SRA_SLICE;
  int R = sra_bin_setup(a, s);
  if (R == 0) return;
  // stationary operands of this pass: q-hat (normalised, LUT added) and dO
  sra_stage<HD>(s, R, a.qkv, 3 * d, col, a.lut, 2 * d, col, dout, d, col);
  for (int i = threadIdx.x; i < R * HS; i += blockDim.x) {
    int r = i / HS, h = i % HS;
    long long g = (long long)s.info[r].x * 8 + blockIdx.y * HS + h;
    s.lse[r * 8 + h] = lse[g];
    s.D[r * 8 + h] = Dbuf[g];
  }
  __syncthreads();
  float inv_tau = 1.f / fmaxf(__ldg(a.tau), a.tau_min);
  for (int p = threadIdx.x; p < R * HS; p += blockDim.x) {
    int r = p % R, h = p / R;
    int4 inf = s.info[r];
    float k[HD], v[HD], dk[HD], dv[HD];
    load_head_add<HD>(a.qkv + (long long)inf.x * 3 * d + d + col + h * HD, a.lut + inf.w * 2 * d + d + col + h * HD, k);
    load_head<HD>(a.qkv + (long long)inf.x * 3 * d + 2 * d + col + h * HD, v);
    float ikn = 1.f / fmaxf(sqrtf(dot_reg<HD>(k, k)), SRA_EPS);
#pragma unroll
    for (int i = 0; i < HD; ++i) { k[i] *= ikn; dk[i] = 0.f; dv[i] = 0.f; }
    for (int i = inf.y; i < inf.z; ++i) {
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
    store_head<HD>(dqkv + (long long)inf.x * 3 * d + d + col + h * HD, dk);
    store_head<HD>(dqkv + (long long)inf.x * 3 * d + 2 * d + col + h * HD, dv);
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
extern "C" int gdmae_sra_attention_fwd(const float* qkv, const float* lut, const int32_t* row_info, int64_t N, int d, int nhead,
                                       const float* tau, float tau_min, float* out, float* lse, void* stream_) {
  int rc = sra_check(N, d, nhead, row_info);
  if (rc) return rc;
  if (N == 0) return GDMAE_OK;
  SraArgs a{qkv, lut, (const int4*)row_info, tau, tau_min, (int)N, d};
  cudaStream_t st = (cudaStream_t)stream_;
  dim3 grid(gdmae_div_up(N, SRA_BIN), d / SRA_SLICE);
  if ((rc = sra_smem_attrs())) return rc;
  if (d == 128) sra_fwd_kernel<16><<<grid, 256, SRA_SMEM_BYTES, st>>>(a, out, lse);
  else sra_fwd_kernel<32><<<grid, 256, SRA_SMEM_BYTES, st>>>(a, out, lse);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// dqkv (N,3d) = [dq | dk | dv]; dtau_sum (1, double, caller zeroes) accumulates sum dS*S
// (d loss / d tau = -dtau_sum / tau_c when tau >= tau_min, else 0); work_D (N,8) scratch.
extern "C" int gdmae_sra_attention_bwd(const float* qkv, const float* lut, const int32_t* row_info, int64_t N, int d, int nhead,
                                       const float* tau, float tau_min, const float* out, const float* lse, const float* dout,
                                       float* dqkv, double* dtau_sum, float* work_D, void* stream_) {
  int rc = sra_check(N, d, nhead, row_info);
  if (rc) return rc;
  if (N == 0) return GDMAE_OK;
  SraArgs a{qkv, lut, (const int4*)row_info, tau, tau_min, (int)N, d};
  cudaStream_t st = (cudaStream_t)stream_;
  dim3 grid(gdmae_div_up(N, SRA_BIN), d / SRA_SLICE);
  if ((rc = sra_smem_attrs())) return rc;
  if (d == 128) {
    sra_bwd_q_kernel<16><<<grid, 128, SRA_SMEM_BYTES, st>>>(a, out, lse, dout, dqkv, work_D, dtau_sum);
    GDMAE_LAUNCH_CHECK();
    sra_bwd_kv_kernel<16><<<grid, 128, SRA_SMEM_BYTES, st>>>(a, lse, dout, work_D, dqkv);
  } else {
    sra_bwd_q_kernel<32><<<grid, 128, SRA_SMEM_BYTES, st>>>(a, out, lse, dout, dqkv, work_D, dtau_sum);
    GDMAE_LAUNCH_CHECK();
    sra_bwd_kv_kernel<32><<<grid, 128, SRA_SMEM_BYTES, st>>>(a, lse, dout, work_D, dqkv);
  }
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
