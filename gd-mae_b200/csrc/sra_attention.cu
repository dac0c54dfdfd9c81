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
// Here nothing is padded: a token finds its window partners through the CSR window table
// (win_off / win_tok from gdmae_window_table), the positional embedding enters as a 64-row
// look-up table that was already multiplied by Wq / Wk (pos has only 64 distinct rows), and
// softmax is computed online.  HBM traffic = qkv in + o out (+ the L2-resident LUT):
// N*d*(3+1)*4 + N*8 bytes, the algorithmic minimum of SURVEY.md 8d.
//
// Thread mapping: 16 channels per thread, d/16 consecutive threads per token (8 for d=128, 16
// for d=256) so that a token's row is one contiguous 512 B / 1 KB segment across its threads;
// SUB = head_dim/16 threads share a head and combine dot products with one xor-shuffle.
#include "common.cuh"

#define SRA_EPS 1e-12f

struct SraArgs {
  const float* qkv;          // (N, 3d): x Wq^T | x Wk^T | x Wv^T + bv   (q, k without bias / pos term)
  const float* lut;          // (64, 2d): pos_table Wq^T + bq | pos_table Wk^T + bk
  const int* win_tok;        // (N) tokens grouped by window
  const int* win_of;         // (N) dense window id per token
  const unsigned char* pos_of;  // (N) in-window cell
  const int* win_off;        // (nW+1)
  const float* tau;          // (1) learnable temperature
  float tau_min;
  int N;
};

__device__ __forceinline__ void load16(const float* __restrict__ p, float* v) {
  const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 t = __ldg(p4 + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void load16_add(const float* __restrict__ a, const float* __restrict__ b, float* v) {
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 s = __ldg(a4 + i), t = __ldg(b4 + i);
    v[4 * i] = s.x + t.x; v[4 * i + 1] = s.y + t.y; v[4 * i + 2] = s.z + t.z; v[4 * i + 3] = s.w + t.w;
  }
}
__device__ __forceinline__ void store16(float* __restrict__ p, const float* v) {
  float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) p4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ float dot16(const float* a, const float* b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s = fmaf(a[i], b[i], s);
  return s;
}
template <int SUB>
__device__ __forceinline__ float head_sum(float v) {
  if (SUB == 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------ forward
template <int SUB>
__global__ void __launch_bounds__(256) sra_fwd_kernel(SraArgs a, float* __restrict__ out, float* __restrict__ lse) {
  constexpr int G = 8 * SUB;  // threads per token
  constexpr int D = 16 * G;
  long long g = blockIdx.x * 256ll + threadIdx.x;
  int p = (int)(g / G), c = (int)(g % G);
  bool live = p < a.N;
  int t = a.win_tok[live ? p : a.N - 1];
  int w = a.win_of[t];
  int s = a.win_off[w];
  int n = live ? a.win_off[w + 1] - s : 0;
  float inv_tau = 1.f / fmaxf(__ldg(a.tau), a.tau_min);

  float q[16];
  load16_add(a.qkv + (long long)t * 3 * D + 16 * c, a.lut + (int)a.pos_of[t] * 2 * D + 16 * c, q);
  float qn = fmaxf(sqrtf(head_sum<SUB>(dot16(q, q))), SRA_EPS);
  float qs = inv_tau / qn;
#pragma unroll
  for (int i = 0; i < 16; ++i) q[i] *= qs;

  float m = -INFINITY, l = 0.f, o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = 0.f;

  int nmax = warp_max_int(n);
  for (int j = 0; j < nmax; ++j) {
    bool valid = j < n;
    int u = a.win_tok[valid ? s + j : s];
    float k[16];
    load16_add(a.qkv + (long long)u * 3 * D + D + 16 * c, a.lut + (int)a.pos_of[u] * 2 * D + D + 16 * c, k);
    float kk = head_sum<SUB>(dot16(k, k));
    float qk = head_sum<SUB>(dot16(q, k));
    if (valid) {
      float sc = qk / fmaxf(sqrtf(kk), SRA_EPS);
      float mn = fmaxf(m, sc);
      float corr = expf(m - mn);  // m = -inf on the first key -> 0
      float pj = expf(sc - mn);
      l = l * corr + pj;
      float v[16];
      load16(a.qkv + (long long)u * 3 * D + 2 * D + 16 * c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = fmaf(pj, v[i], o[i] * corr);
      m = mn;
    }
  }
  if (live) {
    float il = 1.f / l;
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] *= il;
    store16(out + (long long)t * D + 16 * c, o);
    if ((c % SUB) == 0) lse[(long long)t * 8 + c / SUB] = m + logf(l);
  }
}

// ------------------------------------------------------------------------------ backward, query side
// dq_t and D_t = dO_t . O_t ; also accumulates sum_ij dS_ij S_ij for the temperature gradient.
template <int SUB>
__global__ void __launch_bounds__(256) sra_bwd_q_kernel(SraArgs a, const float* __restrict__ out, const float* __restrict__ lse,
                                                        const float* __restrict__ dout, float* __restrict__ dqkv,
                                                        float* __restrict__ Dbuf, double* __restrict__ dtau_acc) {
  constexpr int G = 8 * SUB;
  constexpr int D = 16 * G;
  long long g = blockIdx.x * 256ll + threadIdx.x;
  int p = (int)(g / G), c = (int)(g % G);
  bool live = p < a.N;
  int t = a.win_tok[live ? p : a.N - 1];
  int w = a.win_of[t];
  int s = a.win_off[w];
  int n = live ? a.win_off[w + 1] - s : 0;
  float inv_tau = 1.f / fmaxf(__ldg(a.tau), a.tau_min);

  float q[16], dO[16], o[16];
  load16_add(a.qkv + (long long)t * 3 * D + 16 * c, a.lut + (int)a.pos_of[t] * 2 * D + 16 * c, q);
  load16(dout + (long long)t * D + 16 * c, dO);
  load16(out + (long long)t * D + 16 * c, o);
  float qn = fmaxf(sqrtf(head_sum<SUB>(dot16(q, q))), SRA_EPS);
  float iqn = 1.f / qn;
#pragma unroll
  for (int i = 0; i < 16; ++i) q[i] *= iqn;  // q-hat
  float Dt = head_sum<SUB>(dot16(dO, o));
  float ls = lse[(long long)t * 8 + c / SUB];
  float dqh[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dqh[i] = 0.f;
  float tacc = 0.f;

  int nmax = warp_max_int(n);
  for (int j = 0; j < nmax; ++j) {
    bool valid = j < n;
    int u = a.win_tok[valid ? s + j : s];
    float k[16], v[16];
    load16_add(a.qkv + (long long)u * 3 * D + D + 16 * c, a.lut + (int)a.pos_of[u] * 2 * D + D + 16 * c, k);
    load16(a.qkv + (long long)u * 3 * D + 2 * D + 16 * c, v);
    float kk = head_sum<SUB>(dot16(k, k));
    float qk = head_sum<SUB>(dot16(q, k));
    float dp = head_sum<SUB>(dot16(dO, v));
    if (valid) {
      float ikn = 1.f / fmaxf(sqrtf(kk), SRA_EPS);
      float sc = qk * ikn * inv_tau;
      float pj = expf(sc - ls);
      float ds = pj * (dp - Dt);
      tacc = fmaf(ds, sc, tacc);
      float f = ds * inv_tau * ikn;
#pragma unroll
      for (int i = 0; i < 16; ++i) dqh[i] = fmaf(f, k[i], dqh[i]);
    }
  }
  // through the L2 normalisation: dq = (dqh - qh (qh . dqh)) / |q|
  float proj = head_sum<SUB>(dot16(q, dqh));
#pragma unroll
  for (int i = 0; i < 16; ++i) dqh[i] = (dqh[i] - q[i] * proj) * iqn;
  if (live) {
    store16(dqkv + (long long)t * 3 * D + 16 * c, dqh);
    if ((c % SUB) == 0) Dbuf[(long long)t * 8 + c / SUB] = Dt;
  }
  // one contribution per head: only the first thread of a head keeps its partial
  if (!live || (c % SUB) != 0) tacc = 0.f;
  tacc = warp_sum(tacc);
  if ((threadIdx.x & 31) == 0 && tacc != 0.f) atomicAdd(dtau_acc, (double)tacc);
}

// ------------------------------------------------------------------------------ backward, key/value side
template <int SUB>
__global__ void __launch_bounds__(256) sra_bwd_kv_kernel(SraArgs a, const float* __restrict__ lse, const float* __restrict__ dout,
                                                         const float* __restrict__ Dbuf, float* __restrict__ dqkv) {
  constexpr int G = 8 * SUB;
  constexpr int D = 16 * G;
  long long g = blockIdx.x * 256ll + threadIdx.x;
  int p = (int)(g / G), c = (int)(g % G);
  bool live = p < a.N;
  int t = a.win_tok[live ? p : a.N - 1];
  int w = a.win_of[t];
  int s = a.win_off[w];
  int n = live ? a.win_off[w + 1] - s : 0;
  float inv_tau = 1.f / fmaxf(__ldg(a.tau), a.tau_min);

  float k[16], v[16];
  load16_add(a.qkv + (long long)t * 3 * D + D + 16 * c, a.lut + (int)a.pos_of[t] * 2 * D + D + 16 * c, k);
  load16(a.qkv + (long long)t * 3 * D + 2 * D + 16 * c, v);
  float kn = fmaxf(sqrtf(head_sum<SUB>(dot16(k, k))), SRA_EPS);
  float ikn = 1.f / kn;
#pragma unroll
  for (int i = 0; i < 16; ++i) k[i] *= ikn;  // k-hat
  float dkh[16], dv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { dkh[i] = 0.f; dv[i] = 0.f; }

  int nmax = warp_max_int(n);
  for (int j = 0; j < nmax; ++j) {
    bool valid = j < n;
    int u = a.win_tok[valid ? s + j : s];
    float q[16], dO[16];
    load16_add(a.qkv + (long long)u * 3 * D + 16 * c, a.lut + (int)a.pos_of[u] * 2 * D + 16 * c, q);
    load16(dout + (long long)u * D + 16 * c, dO);
    float qq = head_sum<SUB>(dot16(q, q));
    float qk = head_sum<SUB>(dot16(q, k));
    float dp = head_sum<SUB>(dot16(dO, v));
    if (valid) {
      float iqn = 1.f / fmaxf(sqrtf(qq), SRA_EPS);
      float sc = qk * iqn * inv_tau;
      float pj = expf(sc - lse[(long long)u * 8 + c / SUB]);
      float ds = pj * (dp - Dbuf[(long long)u * 8 + c / SUB]);
      float f = ds * inv_tau * iqn;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        dkh[i] = fmaf(f, q[i], dkh[i]);
        dv[i] = fmaf(pj, dO[i], dv[i]);
      }
    }
  }
  float proj = head_sum<SUB>(dot16(k, dkh));
#pragma unroll
  for (int i = 0; i < 16; ++i) dkh[i] = (dkh[i] - k[i] * proj) * ikn;
  if (live) {
    store16(dqkv + (long long)t * 3 * D + D + 16 * c, dkh);
    store16(dqkv + (long long)t * 3 * D + 2 * D + 16 * c, dv);
  }
}

static int sra_check(int64_t N, int d, int nhead) {
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 27));
  GDMAE_CHECK_ARG(nhead == 8 && (d == 128 || d == 256));
  return GDMAE_OK;
}

// o (N,d) = softmax_j( cos(q_i, k_j) / max(tau, tau_min) ) v_j over the tokens j of i's window; lse (N,8).
extern "C" int gdmae_sra_attention_fwd(const float* qkv, const float* lut, const int32_t* win_tok, const int32_t* win_of_token,
                                       const uint8_t* pos_of_token, const int32_t* win_off, int64_t N, int d, int nhead,
                                       const float* tau, float tau_min, float* out, float* lse, void* stream_) {
  int rc = sra_check(N, d, nhead);
  if (rc) return rc;
  if (N == 0) return GDMAE_OK;
  SraArgs a{qkv, lut, win_tok, win_of_token, pos_of_token, win_off, tau, tau_min, (int)N};
  cudaStream_t st = (cudaStream_t)stream_;
  if (d == 128) sra_fwd_kernel<1><<<gdmae_div_up(N * 8, 256), 256, 0, st>>>(a, out, lse);
  else sra_fwd_kernel<2><<<gdmae_div_up(N * 16, 256), 256, 0, st>>>(a, out, lse);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// dqkv (N,3d) = [dq | dk | dv]; dtau_sum (1, double, caller zeroes) accumulates sum dS*S
// (d loss / d tau = -dtau_sum / tau_c when tau >= tau_min, else 0); work_D (N,8) scratch.
extern "C" int gdmae_sra_attention_bwd(const float* qkv, const float* lut, const int32_t* win_tok, const int32_t* win_of_token,
                                       const uint8_t* pos_of_token, const int32_t* win_off, int64_t N, int d, int nhead,
                                       const float* tau, float tau_min, const float* out, const float* lse, const float* dout,
                                       float* dqkv, double* dtau_sum, float* work_D, void* stream_) {
  int rc = sra_check(N, d, nhead);
  if (rc) return rc;
  if (N == 0) return GDMAE_OK;
  SraArgs a{qkv, lut, win_tok, win_of_token, pos_of_token, win_off, tau, tau_min, (int)N};
  cudaStream_t st = (cudaStream_t)stream_;
  if (d == 128) {
    int g = gdmae_div_up(N * 8, 256);
    sra_bwd_q_kernel<1><<<g, 256, 0, st>>>(a, out, lse, dout, dqkv, work_D, dtau_sum);
    GDMAE_LAUNCH_CHECK();
    sra_bwd_kv_kernel<1><<<g, 256, 0, st>>>(a, lse, dout, work_D, dqkv);
  } else {
    int g = gdmae_div_up(N * 16, 256);
    sra_bwd_q_kernel<2><<<g, 256, 0, st>>>(a, out, lse, dout, dqkv, work_D, dtau_sum);
    GDMAE_LAUNCH_CHECK();
    sra_bwd_kv_kernel<2><<<g, 256, 0, st>>>(a, lse, dout, work_D, dqkv);
  }
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
