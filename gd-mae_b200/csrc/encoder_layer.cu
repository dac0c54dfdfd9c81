// Encoder-layer executor: the whole forward (and the whole hand-written backward) of one SST
// EncoderLayer issued from ONE C call - 4 cuBLASLt GEMMs + 6 kernels forward, 10 GEMMs + 8 kernels
// backward, all on the caller's stream.
//
// Replaces (reference file:line, relative to /root/reference):
//   EncoderLayer.forward (post-norm)           pcdet/models/model_utils/sst_basic_block.py:60-92
//   WindowAttention.forward                     pcdet/models/model_utils/sst_basic_block.py:22-54
//   CosineMultiheadAttention in/out projection  pcdet/models/model_utils/cosine_msa.py:57-62, 380-431
//   and the autograd graph torch builds for them (about 60 nodes per layer in the reference).
//
// Why native: the step runs 12 of these layers; driven from Python through ctypes each layer cost
// ~0.4 ms (forward) + ~0.8 ms (backward) of host time (r1 profile), which made the whole step
// host-bound.  The executor only sequences the kernels of elementwise.cu / sra_attention*.cu /
// gemm.cu; the caller owns every buffer (activations saved for backward, gradients, workspace).
#include "common.cuh"
#include <cuda_bf16.h>
#include "../../include/gdmae_b200.h"

// lut[p, j] = b_in[j] + sum_c pos_table[p, c] * w_in[j, c]   (p < 64, j < 2d): positional term of q and k
// incl. their biases.  One warp per output, coalesced reads of both rows.
__global__ void __launch_bounds__(256) pos_lut_kernel(const float* __restrict__ pos, const float* __restrict__ w_in,
                                                      const float* __restrict__ b_in, int d, float* __restrict__ lut) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = 64 * 2 * d;
  if (warp >= total) return;
  const int p = warp / (2 * d), j = warp % (2 * d);
  const float* pr = pos + (long long)p * d;
  const float* wr = w_in + (long long)j * d;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) acc = fmaf(__ldg(pr + c), __ldg(wr + c), acc);
  acc = warp_sum(acc);
  if (lane == 0) lut[warp] = acc + __ldg(b_in + j);
}

// fp32 -> bf16 copy of the layer input when the previous kernel did not hand one over
__global__ void cast_bf16_kernel(const float4* __restrict__ x, long long n4, uint2* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(x + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<unsigned*>(&a);
    u.y = *reinterpret_cast<unsigned*>(&b);
    out[i] = u;
  }
}

// dtau (+)= tau >= tau_min ? -(sum dS*S) / tau : 0      (S = cos / clamp(tau): dS/dtau = -S / tau)
__global__ void dtau_kernel(const double* __restrict__ dtau_sum, const float* __restrict__ tau, float tau_min, int accumulate,
                            float* __restrict__ dtau) {
  float t = tau[0];
  float g = t >= tau_min ? -(float)dtau_sum[0] / fmaxf(t, tau_min) : 0.f;
  dtau[0] = accumulate ? dtau[0] + g : g;
}

// internal forms of the row kernels (elementwise.cu): bf16 GEMM outputs as inputs, bias gradient fused into the LayerNorm backward
int ew_add_layernorm_fwd(const float* x, const void* res, int res_bf16, const float* bias, const float* gamma, const float* beta,
                         int64_t N, int d, float eps, float* y, void* y_bf16, float* mean, float* rstd, void* stream_);
int ew_add_layernorm_bwd(const float* x, const void* res, int res_bf16, const float* bias, const float* gamma, const float* mean,
                         const float* rstd, const float* dy, int64_t N, int d, float* dz, void* dz_bf16, float* dgamma,
                         float* dbeta, float* dbias, int accumulate, void* workspace, size_t ws_bytes, void* stream_);
int ew_bias_gelu_fwd(const void* h, int h_bf16, const float* bias, int64_t N, int C, float* out, void* out_bf16, void* stream_);
int ew_bias_gelu_bwd(const void* h, const void* dg, int hdg_bf16, const float* bias, int64_t N, int C, float* dh, void* dh_bf16,
                     float* dbias, int accumulate, void* workspace, size_t ws_bytes, void* stream_);

#define EL_CALL(expr)        \
  do {                       \
    int _rc = (expr);        \
    if (_rc) return _rc;     \
  } while (0)

static int el_check(const gdmae_encoder_layer_args* a) {
  GDMAE_CHECK_ARG(a && a->N >= 0 && (a->d == 128 || a->d == 256) && a->dff > 0 && a->dff % 8 == 0 && a->nhead == 8);
  GDMAE_CHECK_ARG(a->gemm_mode >= 0 && a->gemm_mode <= 2);
  GDMAE_CHECK_ARG(!a->sra_tensor_cores || a->gemm_mode == 1);   // the tensor-core SRA kernels take bf16 q/k/v
  return GDMAE_OK;
}

// bf16 configuration: every contraction of the layer runs on the own tcgen05 / TMA kernel (csrc/tc_gemm.cu); the weight
// gradients (transposed A, K = tokens) split K across the SMs and reduce into the gradient bucket.  The fp32 parity
// configurations keep the library GEMM (strict fp32 / TF32 math has no tcgen05 kind::f16 equivalent).
static inline int el_gemm(const gdmae_encoder_layer_args* a, int ta, int tb, int64_t M, int64_t N, int64_t K, const void* A,
                          int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int c_bf16, float beta) {
  if (a->gemm_mode == 1)
    return gdmae_tc_gemm(ta, tb, M, N, K, A, lda, B, ldb, C, ldc, c_bf16, beta, ta ? 1 : 0, nullptr, a->stream);
  return gdmae_gemm(ta, tb, M, N, K, A, lda, B, ldb, a->gemm_mode, C, ldc, c_bf16, beta, a->stream);
}

extern "C" int gdmae_encoder_layer_fwd(const gdmae_encoder_layer_args* a) {
  EL_CALL(el_check(a));
  if (a->N == 0) return GDMAE_OK;
  const int64_t N = a->N;
  const int d = a->d, dff = a->dff;
  const bool bf = a->gemm_mode == 1;
  cudaStream_t st = (cudaStream_t)a->stream;
  const void* xg = a->xg_in;
  if (!xg) {
    if (bf) {
      long long n4 = N * d / 4;
      cast_bf16_kernel<<<gdmae_grid(n4, 256, 16), 256, 0, st>>>((const float4*)a->x, n4, (uint2*)a->xg);
      GDMAE_LAUNCH_CHECK();
      xg = a->xg;
    } else {
      xg = a->x;
    }
  }
  // in-projection without biases; positional term + q/k biases go through the 64-row LUT, the v bias to the output
  const int tc = a->sra_tensor_cores ? 1 : 0;
  pos_lut_kernel<<<gdmae_div_up(64ll * 2 * d * 32, 256), 256, 0, st>>>(a->pos_table, a->w_in, a->b_in, d, a->lut);
  GDMAE_LAUNCH_CHECK();
  if (tc) {
    // tensor-core kernels: the GEMM epilogue adds the LUT row, normalises q and k per head and writes q^, k^, v (bf16) in
    // CSR (window) order - the attention kernel then fetches whole bins of windows with TMA boxes
    gdmae_tc_epilogue e0 = {};
    e0.mode = 4; e0.tok_info = a->bin_units + gdmae_sra_tok_info_offset(N); e0.lut = a->lut; e0.tau = a->tau; e0.tau_min = a->tau_min;
    e0.lrr = a->lse;
    EL_CALL(gdmae_tc_gemm(0, 1, N, 3 * d, d, xg, d, a->w_in_g, d, a->qkv, 3 * d, 1, 0.f, 0, &e0, a->stream));
  } else {
    EL_CALL(el_gemm(a, 0, 1, N, 3 * d, d, xg, d, a->w_in_g, d, a->qkv, 3 * d, 0, 0.f));
  }
  GdmaeSpan span(st);
  if (tc)
    EL_CALL(gdmae_sra_fwd_win(a->qkv, a->bin_units, N, d, a->b_in + 2 * d, bf, a->o, a->lse, 1, a->stream));
  else
    EL_CALL(gdmae_sra_attention_fwd((const float*)a->qkv, a->lut, a->row_info, N, d, a->nhead, a->tau, a->tau_min, a->b_in + 2 * d, bf, a->o,
                                    a->lse, a->stream));
  // algorithmic bytes (SURVEY.md 8d, a18 minus projections): q, k, v in, o out, lse out
  span.end(0, d, N, N * d * (3 * (tc ? 2 : 4) + (bf ? 2 : 4)) + N * 32);
  // bf16 configuration: the GEMM outputs that only feed a row kernel (a, h, f) leave the GEMM epilogue as bf16
  const int ob = bf ? 1 : 0;
  if (bf) {
    // three GEMMs, no row kernels: residual + bias + LayerNorm and bias + GELU run in the GEMM epilogues on the
    // accumulator rows in tensor memory; a / f (bf16) are still written because the backward pass reads them, and h holds
    // gelu'(W1 x1 + b1) (bf16) - the only thing the backward pass needs of the pre-activation
    gdmae_tc_epilogue e1 = {};
    e1.mode = 2; e1.bias = a->b_o; e1.res = a->x; e1.gamma = a->g1; e1.beta_ln = a->be1; e1.eps = a->eps;
    e1.y32 = a->x1; e1.y16 = a->x1g; e1.mean = a->mean1; e1.rstd = a->rstd1;
    EL_CALL(gdmae_tc_gemm(0, 1, N, d, d, a->o, d, a->w_o_g, d, a->a, d, 1, 0.f, 0, &e1, a->stream));
    gdmae_tc_epilogue e2 = {};
    e2.mode = 1; e2.bias = a->b1; e2.c2 = a->g; e2.ldc2 = dff;
    EL_CALL(gdmae_tc_gemm(0, 1, N, dff, d, a->x1g, d, a->w1_g, d, a->h, dff, 1, 0.f, 0, &e2, a->stream));
    gdmae_tc_epilogue e3 = {};
    e3.mode = 2; e3.bias = a->b2; e3.res = a->x1; e3.gamma = a->g2; e3.beta_ln = a->be2; e3.eps = a->eps;
    e3.y32 = a->x2; e3.y16 = a->x2g; e3.mean = a->mean2; e3.rstd = a->rstd2;
    EL_CALL(gdmae_tc_gemm(0, 1, N, d, dff, a->g, dff, a->w2_g, dff, a->f, d, 1, 0.f, 0, &e3, a->stream));
    return GDMAE_OK;
  }
  EL_CALL(el_gemm(a, 0, 1, N, d, d, a->o, d, a->w_o_g, d, a->a, d, ob, 0.f));
  EL_CALL(ew_add_layernorm_fwd(a->x, a->a, ob, a->b_o, a->g1, a->be1, N, d, a->eps, a->x1, nullptr, a->mean1, a->rstd1, a->stream));
  EL_CALL(el_gemm(a, 0, 1, N, dff, d, (const void*)a->x1, d, a->w1_g, d, a->h, dff, ob, 0.f));
  EL_CALL(ew_bias_gelu_fwd(a->h, ob, a->b1, N, dff, (float*)a->g, nullptr, a->stream));
  EL_CALL(el_gemm(a, 0, 1, N, d, dff, a->g, dff, a->w2_g, dff, a->f, d, ob, 0.f));
  EL_CALL(ew_add_layernorm_fwd(a->x1, a->f, ob, a->b2, a->g2, a->be2, N, d, a->eps, a->x2, nullptr, a->mean2, a->rstd2, a->stream));
  return GDMAE_OK;
}

extern "C" size_t gdmae_encoder_layer_bwd_workspace_bytes(int64_t N, int d, int dff) {
  size_t n = (size_t)(N > 0 ? N : 1);
  // fp32: dz2, dgl, do; operand dtype (<= 4 bytes): dz2g, dh, dz1g, dqkv, xpos; work, dtau_sum; row-kernel scratch
  size_t floats = n * d + n * dff + n * d + n * d + n * dff + n * d + n * 3 * d + n * d + n * 8;
  return floats * 4 + 16 * 256 + gdmae_rowwise_workspace_bytes(1024);
}

extern "C" int gdmae_encoder_layer_bwd(const gdmae_encoder_layer_args* a) {
  EL_CALL(el_check(a));
  if (a->N == 0) return GDMAE_OK;
  const int64_t N = a->N;
  const int d = a->d, dff = a->dff;
  const bool bf = a->gemm_mode == 1;
  const int acc = a->accumulate ? 1 : 0;
  const float wbeta = acc ? 1.f : 0.f;
  cudaStream_t st = (cudaStream_t)a->stream;
  const void* xg = a->xg_in ? a->xg_in : (bf ? (const void*)a->xg : (const void*)a->x);
  Workspace ws(a->ws, a->ws_bytes);
  float* dz2 = ws.take<float>(N * d);
  float* dgl = ws.take<float>(N * dff);
  float* dout = ws.take<float>(N * d);
  float* dz2g = ws.take<float>(N * d);        // operand dtype buffers are sized for fp32
  float* dh = ws.take<float>(N * dff);
  float* dz1g = ws.take<float>(N * d);
  float* dqkv = ws.take<float>(N * 3 * d);
  float* xpos = ws.take<float>(N * d);
  float* work = ws.take<float>(N * 8);
  double* dtau_sum = ws.take<double>(1);
  const size_t rw_bytes = gdmae_rowwise_workspace_bytes(1024);
  void* rw = ws.take<char>(rw_bytes);
  GDMAE_CHECK_ARG(rw != nullptr && "workspace too small: gdmae_encoder_layer_bwd_workspace_bytes");
  float* dz1 = a->dx;                          // the residual gradient accumulates into the output
  const void* x1g = bf ? a->x1g : (const void*)a->x1;
  const size_t es = bf ? 2 : 4;

  // ---- LayerNorm 2 and the feed-forward
  const int ob = bf ? 1 : 0;   // a, h, f (forward) and dgl are bf16 GEMM outputs in the bf16 configuration
  EL_CALL(ew_add_layernorm_bwd(a->x1, a->f, ob, a->b2, a->g2, a->mean2, a->rstd2, a->dy, N, d, dz2, bf ? dz2g : nullptr, a->d_g2,
                               a->d_be2, a->d_b2, acc, rw, rw_bytes, a->stream));   // d_b2 = column sums of dz2, same pass
  const void* dz2_op = bf ? (const void*)dz2g : (const void*)dz2;
  EL_CALL(el_gemm(a, 1, 0, d, dff, N, dz2_op, d, a->g, dff, a->d_w2, dff, 0, wbeta));
  if (bf) {
    // dgl = dz2 W2 never leaves the GEMM: its epilogue multiplies by the saved gelu'(h + b1), writes dh (bf16) and adds the column
    // sums (the gradient of b1) to d_b1
    if (!acc) GDMAE_CHECK_CUDA(cudaMemsetAsync(a->d_b1, 0, (size_t)dff * sizeof(float), st));
    gdmae_tc_epilogue eg = {};
    eg.mode = 3; eg.bias = a->b1; eg.h16 = a->h; eg.ldh = dff; eg.colsum = a->d_b1;
    EL_CALL(gdmae_tc_gemm(0, 0, N, dff, d, dz2_op, d, a->w2_g, dff, dh, dff, 1, 0.f, 0, &eg, a->stream));
  } else {
    EL_CALL(el_gemm(a, 0, 0, N, dff, d, dz2_op, d, a->w2_g, dff, dgl, dff, ob, 0.f));
    EL_CALL(ew_bias_gelu_bwd(a->h, dgl, ob, a->b1, N, dff, dh, nullptr, a->d_b1, acc, rw, rw_bytes, a->stream));
  }
  EL_CALL(el_gemm(a, 1, 0, dff, d, N, dh, dff, x1g, d, a->d_w1, d, 0, wbeta));
  EL_CALL(el_gemm(a, 0, 0, N, d, dff, dh, dff, a->w1_g, d, dz2, d, 0, 1.f));   // dz2 := gradient w.r.t. x1
  // ---- LayerNorm 1 and the attention
  EL_CALL(ew_add_layernorm_bwd(a->x, a->a, ob, a->b_o, a->g1, a->mean1, a->rstd1, dz2, N, d, dz1, bf ? dz1g : nullptr, a->d_g1,
                               a->d_be1, a->d_b_o, acc, rw, rw_bytes, a->stream));  // d_b_o = column sums of dz1
  const void* dz1_op = bf ? (const void*)dz1g : (const void*)dz1;
  EL_CALL(el_gemm(a, 1, 0, d, d, N, dz1_op, d, a->o, d, a->d_w_o, d, 0, wbeta));
  const int tc = a->sra_tensor_cores ? 1 : 0;
  if (tc) {
    // dO = dz1 Wo leaves the GEMM as bf16 rows in CSR (window) order: the fourth tensor of the window-major array
    gdmae_tc_epilogue e5 = {};
    e5.mode = 5; e5.tok_info = a->bin_units + gdmae_sra_tok_info_offset(N); e5.plane0 = 3 * (d / 64);
    EL_CALL(gdmae_tc_gemm(0, 0, N, d, d, dz1_op, d, a->w_o_g, d, a->qkv, d, 1, 0.f, 0, &e5, a->stream));
  } else {
    EL_CALL(el_gemm(a, 0, 0, N, d, d, dz1_op, d, a->w_o_g, d, dout, d, 0, 0.f));
  }
  GDMAE_CHECK_CUDA(cudaMemsetAsync(dtau_sum, 0, sizeof(double), st));
  GdmaeSpan span(st);
  if (tc)
    EL_CALL(gdmae_sra_bwd_win(a->qkv, a->lse, a->bin_units, N, d, a->tau, a->tau_min, dqkv, dtau_sum, a->stream));
  else
    EL_CALL(gdmae_sra_attention_bwd((const float*)a->qkv, a->lut, a->row_info, N, d, a->nhead, a->tau, a->tau_min, a->b_in + 2 * d, bf,
                                    a->o, a->lse, dout, dqkv, dtau_sum, work, a->stream));
  // q, k, v, dO (+ o for the SIMT kernels) in, dq, dk, dv out, lse in
  span.end(1, d, N, tc ? N * d * (6 + 2 + 6) + N * 32 : N * d * (12 + (bf ? 2 : 4) + 4 + 3 * (bf ? 2 : 4)) + N * 32);
  dtau_kernel<<<1, 1, 0, st>>>(dtau_sum, a->tau, a->tau_min, acc, a->d_tau);
  GDMAE_LAUNCH_CHECK();
  // in-projection: q = (x + pos) Wq^T + bq, k likewise, v = x Wv^T + bv
  EL_CALL(gdmae_gather_add_rows(a->x, a->pos_table, a->pos_of_token, N, d, bf ? nullptr : xpos, bf ? xpos : nullptr, a->stream));
  EL_CALL(el_gemm(a, 1, 0, 2 * d, d, N, dqkv, 3 * d, xpos, d, a->d_w_in, d, 0, wbeta));
  EL_CALL(el_gemm(a, 1, 0, d, d, N, (const char*)dqkv + (size_t)2 * d * es, 3 * d, xg, d, a->d_w_in + (size_t)2 * d * d, d, 0, wbeta));
  EL_CALL(gdmae_colsum(dqkv, bf, N, 3 * d, 0, 3 * d, a->d_b_in, acc, rw, rw_bytes, a->stream));   // q, k and v biases in one pass
  EL_CALL(el_gemm(a, 0, 0, N, d, 3 * d, dqkv, 3 * d, a->w_in_g, d, dz1, d, 0, 1.f));   // dx = residual + through the projection
  return GDMAE_OK;
}
