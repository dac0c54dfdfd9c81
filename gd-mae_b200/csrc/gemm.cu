// Plain dense GEMMs of the path through cuBLASLt (library GEMM, not a hand-written kernel): row-major
//   C (M,N) fp32 = op(A) op(B) + beta * C,   A / B in fp32 (TF32 tensor cores) or bf16, fp32 accumulate.
// Used for the projections / FFN / sparse-conv GEMMs of the bf16 configuration, where the operands
// are the bf16 copies written by the hand-written kernels and the outputs stay fp32.
//
// Host cost matters here: the step issues ~150 GEMMs and the token count M changes every step.
// cublasGemmEx / torch.mm(out_dtype=) run the cublasLt heuristic on every call for mixed bf16->fp32
// problems (~200 us of host time each, measured r1), which made the step host-bound.  This file
// therefore calls cublasLtMatmul directly and caches the selected algorithm per
// (transposes, dtype, N, K and M rounded to 4096) - the heuristic runs once per bucket.
//
// Replaces torch.nn.functional.linear calls of the reference (cuBLAS SGEMM/TF32):
//   cosine_msa.py:57-62,431 (in/out projections), sst_basic_block.py:81 (FFN), spconv GEMMs.
#include "common.cuh"
#include <cublasLt.h>
#include <map>
#include <mutex>
#include <tuple>

#define GEMM_WS_BYTES (64ull << 20)

struct LtState {
  cublasLtHandle_t lt = nullptr;
  void* workspace = nullptr;  // library-owned scratch, allocated once per device on first use
  std::map<std::tuple<int, int, int, long long, long long, long long>, cublasLtMatmulAlgo_t> algos;
};
static LtState g_state[64];
static std::mutex g_mutex;

static int lt_fail(const char* what, int st) {
  char buf[160];
  snprintf(buf, sizeof(buf), "%s failed with cuBLAS status %d", what, st);
  gdmae_set_error(buf);
  return GDMAE_ERR_CUDA;
}

#define LT_CHECK(expr)                                             \
  do {                                                             \
    cublasStatus_t _s = (expr);                                    \
    if (_s != CUBLAS_STATUS_SUCCESS) return lt_fail(#expr, (int)_s); \
  } while (0)

// Row-major: C (M,N, ldc) = op(A) (M,K) * op(B) (K,N) + beta * C.  transa: A is stored (K,M) with leading
// dimension lda; transb: B is stored (N,K).  ab_dtype: 0 fp32 (TF32 math), 1 bf16, 2 fp32 (fp32 math).  c_dtype: 0 fp32, 1 bf16
// (bf16 operands only).
extern "C" int gdmae_gemm(int transa, int transb, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                          int64_t ldb, int ab_dtype, void* C, int64_t ldc, int c_dtype, float beta, void* stream_) {
  GDMAE_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && (ab_dtype >= 0 && ab_dtype <= 2) && (c_dtype == 0 || (c_dtype == 1 && ab_dtype == 1)));
  if (M == 0 || N == 0) return GDMAE_OK;
  int dev = 0;
  GDMAE_CHECK_CUDA(cudaGetDevice(&dev));
  GDMAE_CHECK_ARG(dev >= 0 && dev < 64);
  std::lock_guard<std::mutex> lock(g_mutex);
  LtState& S = g_state[dev];
  if (!S.lt) {
    LT_CHECK(cublasLtCreate(&S.lt));
    GDMAE_CHECK_CUDA(cudaMalloc(&S.workspace, GEMM_WS_BYTES));
  }
  cudaDataType_t ab = ab_dtype == 1 ? CUDA_R_16BF : CUDA_R_32F;
  cublasComputeType_t ct = ab_dtype == 0 ? CUBLAS_COMPUTE_32F_FAST_TF32 : CUBLAS_COMPUTE_32F;
  // column-major view of the row-major problem: C^T (N,M) = op(B)^T op(A)^T
  cublasOperation_t op1 = transb ? CUBLAS_OP_T : CUBLAS_OP_N;  // applies to B's memory
  cublasOperation_t op2 = transa ? CUBLAS_OP_T : CUBLAS_OP_N;  // applies to A's memory
  cublasLtMatmulDesc_t desc = nullptr;
  cublasLtMatrixLayout_t l1 = nullptr, l2 = nullptr, lc = nullptr;
  LT_CHECK(cublasLtMatmulDescCreate(&desc, ct, CUDA_R_32F));
  LT_CHECK(cublasLtMatmulDescSetAttribute(desc, CUBLASLT_MATMUL_DESC_TRANSA, &op1, sizeof(op1)));
  LT_CHECK(cublasLtMatmulDescSetAttribute(desc, CUBLASLT_MATMUL_DESC_TRANSB, &op2, sizeof(op2)));
  LT_CHECK(cublasLtMatrixLayoutCreate(&l1, ab, transb ? K : N, transb ? N : K, ldb));
  LT_CHECK(cublasLtMatrixLayoutCreate(&l2, ab, transa ? M : K, transa ? K : M, lda));
  LT_CHECK(cublasLtMatrixLayoutCreate(&lc, c_dtype == 0 ? CUDA_R_32F : CUDA_R_16BF, N, M, ldc));
  const float alpha = 1.f;
  auto bucket = [](long long v) { return v <= 4096 ? v : (v + 4095) / 4096 * 4096; };
  auto key = std::make_tuple(transa, transb, ab_dtype + 4 * (beta != 0.f) + 8 * c_dtype, bucket(M), bucket(N), bucket(K));
  int rc = GDMAE_OK;
  for (int attempt = 0; attempt < 2; ++attempt) {
    auto it = S.algos.find(key);
    if (it == S.algos.end()) {
      cublasLtMatmulPreference_t pref = nullptr;
      cublasLtMatmulHeuristicResult_t res;
      int found = 0;
      size_t wsb = GEMM_WS_BYTES;
      cublasLtMatmulPreferenceCreate(&pref);
      cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &wsb, sizeof(wsb));
      cublasStatus_t hs = cublasLtMatmulAlgoGetHeuristic(S.lt, desc, l1, l2, lc, lc, pref, 1, &res, &found);
      cublasLtMatmulPreferenceDestroy(pref);
      if (hs != CUBLAS_STATUS_SUCCESS || found == 0) { rc = lt_fail("cublasLtMatmulAlgoGetHeuristic", (int)hs); break; }
      it = S.algos.emplace(key, res.algo).first;
    }
    cublasStatus_t ms = cublasLtMatmul(S.lt, desc, &alpha, B, l1, A, l2, &beta, C, lc, C, lc, &it->second, S.workspace, GEMM_WS_BYTES,
                                       (cudaStream_t)stream_);
    if (ms == CUBLAS_STATUS_SUCCESS) break;
    S.algos.erase(key);  // cached algorithm does not fit this exact shape: pick again once
    if (attempt == 1) rc = lt_fail("cublasLtMatmul", (int)ms);
  }
  cublasLtMatrixLayoutDestroy(l1);
  cublasLtMatrixLayoutDestroy(l2);
  cublasLtMatrixLayoutDestroy(lc);
  cublasLtMatmulDescDestroy(desc);
  return rc;
}
