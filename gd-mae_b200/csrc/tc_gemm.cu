// Blackwell-native GEMM of the path: tcgen05.mma (accumulators in TMEM) fed by TMA, persistent and warp-specialised,
// with the row kernels of the encoder layer fused into the epilogue.
//
//   C (M,N) = op(A) (M,K) * op(B) (K,N),   A / B bf16, fp32 accumulation in tensor memory
//
// Replaces the dense contractions the reference runs through torch.nn.functional.linear / spconv's GEMM (cuBLAS):
//   cosine_msa.py:57-62,431   in/out projections        sst_basic_block.py:77-84   FFN + residual + LayerNorm
//   spconv_utils.py:37-56     sparse-conv GEMM          spt_backbone_mae.py:31-44  deblock GEMMs
//   dyn_vfe.py:107-108        VFE layer 2
// and, as epilogues, the row kernels that followed them in r1 (csrc/elementwise.cu): bias + GELU, residual + bias +
// LayerNorm.  The weight-gradient GEMMs (K = tokens) run split-K across the 148 SMs with fp32 reductions straight into
// the gradient bucket (red.global.add.f32) - no split-K workspace, no reduce kernel.
//
// Structure (one CTA per SM, 256 threads):
//   warp 0 (one lane)   TMA producer: cp.async.bulk.tensor.2d global -> 128B-swizzled shared stages, mbarrier expect_tx
//   warp 1 (one lane)   MMA issuer: tcgen05.mma.cta_group::1.kind::f16 128 x BN x 16, tcgen05.commit frees a stage /
//                       publishes an accumulator
//   warp 2              tcgen05.alloc / dealloc of the 512 TMEM columns (two accumulator stages)
//   warps 4-7           epilogue: tcgen05.ld 32x32b (one accumulator row per thread), fused math, then 32-row x 128-byte
//                       units staged in swizzled shared memory and written with TMA stores (cp.async.bulk.tensor, rows
//                       past M clipped by the tensor map) - or TMA reduce-add (cp.reduce.async.bulk.tensor .add) for
//                       the split-K partial tiles; the epilogue of tile i overlaps the main loop of tile i+1 through
//                       the second accumulator stage
// Operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows], the transposed operands of the
// gradient GEMMs): both are TMA tiles of 64 x 128-byte rows in SWIZZLE_128B layout; only the shared-memory descriptor
// (leading / stride byte offsets, K advance) and the instruction descriptor's major bits differ.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>
#include <stdlib.h>
#include "../../include/gdmae_b200.h"

namespace {

constexpr int BM = 128;        // UMMA M (one TMEM lane per accumulator row)
constexpr int BK = 64;         // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
#ifndef TCG_EPI_WARPS
#define TCG_EPI_WARPS 8
#endif
constexpr int NUM_EPI_WARPS = TCG_EPI_WARPS;   // NQ = NUM_EPI_WARPS / 4 per TMEM lane quarter: each takes 1 / NQ of the tile's columns
constexpr int NQ = NUM_EPI_WARPS / 4;
constexpr int NUM_THREADS = 128 + 32 * NUM_EPI_WARPS;
constexpr int UNIT_BYTES = 32 * 128;                  // one epilogue output unit: 32 rows x 128 bytes
constexpr int UNIT_BYTES_TOTAL = NUM_EPI_WARPS * UNIT_BYTES;
constexpr int RING_BYTES = NUM_EPI_WARPS <= 8 ? 192 * 1024 : 156 * 1024;   // + unit buffers + barriers <= 227 KB

struct TcgParams {
  int M, N, K;
  int m_tiles, n_tiles, k_splits, kb_per_split, kb_total;
  int a_mn, b_mn;
  // epilogue
  void* C; long long ldc; int c_bf16; int beta_one; int atomic;
  const float* bias; void* C2; long long ldc2;
  const float* res; const float* gamma; const float* beta_ln; float eps; float* y32; void* y16; float* mean; float* rstd;
  const void* h16; long long ldh; float* colsum;      // GELU-backward epilogue
  const int* tok_info; const float* lut; const float* tau; float tau_min; float* lrr; int plane0;   // window-major epilogues
};

__device__ unsigned int g_tcg_wait_timeouts;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait (about a second): a lost signal ends in a trapped launch (sticky CUDA error), never in a hung GPU.
__device__ __noinline__ void tcg_wait_timed_out() {
  atomicAdd(&g_tcg_wait_timeouts, 1u);
  __threadfence_system();
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (int spin = 0; spin < 100000; ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity), "r"(10000u) : "memory");
    if (ok) return;
  }
  tcg_wait_timed_out();
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (SWIZZLE_128B, sm_100 version bit); offsets in bytes
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive fp32 columns of this thread's accumulator row (TMEM lane = 32 * (warp % 4) + lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// ---- epilogue data path.  An epilogue thread owns one accumulator ROW (tcgen05.ld 32x32b), global memory wants a
// warp to touch whole 128-byte lines: every 32-row x 32-column unit therefore passes through a 4 KB shared buffer of the
// warp, written row-per-thread and read back line-per-quarter-warp, both without bank conflicts:
//   fp32 unit (128-byte rows): 16-byte chunk j of row r at r * 128 + ((j ^ (r & 7)) << 4)   (= TMA SWIZZLE_128B)
//   bf16 unit ( 64-byte rows): 16-byte chunk j of row r at r *  64 + ((j ^ ((r >> 1) & 3)) << 4)
// The split-K partial tiles leave the same fp32 buffer through TMA reduce-add (cp.reduce.async.bulk.tensor .add).
__device__ __forceinline__ uint32_t f32_slot(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t bf16_slot(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void stage_f32(uint8_t* buf, int lane, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(buf + f32_slot(lane, i)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void stage_bf16(uint8_t* buf, int lane, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<uint4*>(buf + bf16_slot(lane, i)) =
        make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]), pack_bf16(v[8 * i + 4], v[8 * i + 5]),
                   pack_bf16(v[8 * i + 6], v[8 * i + 7]));
}
// staged fp32 unit -> global rows r0.. (32 columns from col), 4 rows x 128 bytes per warp instruction
__device__ __forceinline__ void flush_f32(uint8_t* buf, int lane, float* base, long long ld, int r0, int col, int M) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (lane >> 3) + 4 * i, j = lane & 7;
    const float4 val = *reinterpret_cast<const float4*>(buf + f32_slot(r, j));
    if (r0 + r < M) *reinterpret_cast<float4*>(base + (long long)(r0 + r) * ld + col + j * 4) = val;
  }
  __syncwarp();
}
// staged bf16 unit -> global, 8 rows x 64 bytes per warp instruction
__device__ __forceinline__ void flush_bf16(uint8_t* buf, int lane, __nv_bfloat16* base, long long ld, int r0, int col, int M) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i, j = lane & 3;
    const uint4 val = *reinterpret_cast<const uint4*>(buf + bf16_slot(r, j));
    if (r0 + r < M) *reinterpret_cast<uint4*>(base + (long long)(r0 + r) * ld + col + j * 8) = val;
  }
  __syncwarp();
}
// 32 rows x 32 fp32 columns of a global matrix in the coalesced (line-per-quarter-warp) distribution: issued early, used
// after the accumulator chunk has arrived, so that the global-load latency overlaps the tensor-memory loads and staging
__device__ __forceinline__ void prefetch_f32(const float* base, long long ld, int r0, int col, int M, int lane, float4 (&pre)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (lane >> 3) + 4 * i, j = lane & 7;
    pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < M) pre[i] = *reinterpret_cast<const float4*>(base + (long long)(r0 + r) * ld + col + j * 4);
  }
}
// v (this thread's row) += the prefetched unit, transposed through the buffer
__device__ __forceinline__ void add_prefetched(uint8_t* buf, int lane, const float4 (&pre)[8], float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(buf + f32_slot((lane >> 3) + 4 * i, lane & 7)) = pre[i];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 val = *reinterpret_cast<const float4*>(buf + f32_slot(lane, i));
    v[4 * i] += val.x; v[4 * i + 1] += val.y; v[4 * i + 2] += val.z; v[4 * i + 3] += val.w;
  }
  __syncwarp();
}
// staged fp32 unit + prefetched old values -> global (C += acc, coalesced read-modify-write)
__device__ __forceinline__ void flush_f32_add(uint8_t* buf, int lane, float* base, long long ld, int r0, int col, int M, const float4 (&pre)[8]) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (lane >> 3) + 4 * i, j = lane & 7;
    float4 val = *reinterpret_cast<const float4*>(buf + f32_slot(r, j));
    val.x += pre[i].x; val.y += pre[i].y; val.z += pre[i].z; val.w += pre[i].w;
    if (r0 + r < M) *reinterpret_cast<float4*>(base + (long long)(r0 + r) * ld + col + j * 4) = val;
  }
  __syncwarp();
}
// staged fp32 unit -> C += unit, done by the L2 (TMA reduce-add); returns once the engine has read the buffer
__device__ __forceinline__ void reduce_f32(uint8_t* buf, int lane, const CUtensorMap* map, int c0, int r0) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the async proxy
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_u32(buf)), "r"(c0), "r"(r0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncwarp();
}

// GELU through Phi from one exp and one reciprocal (Abramowitz & Stegun 7.1.26, |erf error| <= 1.5e-7; the same
// approximation as elementwise.cu's bf16 configuration - inputs and outputs are bf16 here).
// GELU(v) and d GELU(v) / dv = Phi(v) + v phi(v) from ONE evaluation of the same approximation: exp(-v^2/2) is both the
// tail factor of the erf formula and sqrt(2 pi) phi(v)
__device__ __forceinline__ void gelu_both_fast(float v, float& gl, float& dgl) {
  const float z = fabsf(v) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  const float e = __expf(-0.5f * v * v);
  const float poly = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
  const float cdf = 0.5f * (1.f + copysignf(fmaf(-poly, e, 1.f), v));
  gl = v * cdf;
  dgl = fmaf(v, 0.39894228040143267794f * e, cdf);
}
// 32 rows x 32 bf16 columns of a global matrix, coalesced (8 rows x 64 bytes per warp instruction), requested early
__device__ __forceinline__ void prefetch_bf16(const __nv_bfloat16* base, long long ld, int r0, int col, int M, int lane, uint4 (&pre)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i, j = lane & 3;
    pre[i] = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + r < M) pre[i] = *reinterpret_cast<const uint4*>(base + (long long)(r0 + r) * ld + col + j * 8);
  }
}
// the prefetched unit, transposed through the buffer: this thread's row as 32 floats
__device__ __forceinline__ void take_prefetched_bf16(uint8_t* buf, int lane, const uint4 (&pre)[4], float (&out)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(buf + bf16_slot((lane >> 2) + 8 * i, lane & 3)) = pre[i];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = *reinterpret_cast<const uint4*>(buf + bf16_slot(lane, i));
    const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      out[8 * i + 2 * k] = __uint_as_float(w[k] << 16);
      out[8 * i + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
  }
  __syncwarp();
}
// column sums of a staged fp32 unit (32 rows): lane = column
__device__ __forceinline__ float unit_colsum(const uint8_t* buf, int lane) {
  float t = 0.f;
#pragma unroll
  for (int r = 0; r < 32; ++r) t += *reinterpret_cast<const float*>(buf + f32_slot(r, lane >> 2) + (lane & 3) * 4);
  return t;
}

// staged bf16 unit -> the window-major array [plane][M rows][64] (sra_attention_tc.cu): the 32 columns starting at `col` of
// row r go to row wrow(r) of plane plane0 + col / 64 - 64 contiguous bytes per row, 8 rows per warp instruction
__device__ __forceinline__ void flush_bf16_win(uint8_t* buf, int lane, __nv_bfloat16* base, int plane0, int col, long long M, int wrow, bool valid) {
  __syncwarp();
  __nv_bfloat16* dst0 = base + (long long)(plane0 + (col >> 6)) * M * 64 + (col & 63);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i, j = lane & 3;
    const uint4 val = *reinterpret_cast<const uint4*>(buf + bf16_slot(r, j));
    const int wr = __shfl_sync(0xffffffffu, wrow, r);
    const bool ok = __shfl_sync(0xffffffffu, (int)valid, r) != 0;
    if (ok) *reinterpret_cast<uint4*>(dst0 + (long long)wr * 64 + j * 8) = val;
  }
  __syncwarp();
}

enum { EPI_PLAIN = 0, EPI_GELU = 1, EPI_LN = 2, EPI_GELU_BWD = 3, EPI_QKV_WIN = 4, EPI_ROWS_WIN = 5 };

template <int BN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1) tcg_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const __grid_constant__ CUtensorMap tmC, const TcgParams p) {
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGES = RING_BYTES / STAGE_BYTES;
  constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;     // two accumulator stages (power of two: 128 / 256 / 512)
  extern __shared__ __align__(1024) uint8_t smem[];                  // SWIZZLE_128B stages need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* units = smem + STAGES * STAGE_BYTES;                      // one 4 KB unit buffer per epilogue warp
  float* ln_stat = reinterpret_cast<float*>(units + UNIT_BYTES_TOTAL); // LayerNorm partial sums of the NQ column parts: [2][NQ][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(ln_stat + 2 * NQ * 128);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull + i, 1);
      mbar_init(tempty + i, NUM_EPI_WARPS);      // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles = p.m_tiles * p.n_tiles;
  const int total = tiles * p.k_splits;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int ks = w / tiles, t = w % tiles, m_blk = t / p.n_tiles, n_blk = t % p.n_tiles;
        const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          mbar_expect_tx(full + stage, STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
          if (!p.a_mn) {
            tma_load_2d(sa, &tmA, full + stage, kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(sa + i * (BK * 128), &tmA, full + stage, m_blk * BM + i * 64, kb * BK);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &tmB, full + stage, kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * (BK * 128), &tmB, full + stage, n_blk * BN + i * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: fp32 accumulate, bf16 x bf16, majors, N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // K-major: 8-row groups 1024 bytes apart, 16 K elements = 32 bytes inside the swizzle row.
      // MN-major: 64-element MN groups BK * 128 bytes apart (leading), 8-row K groups 1024 bytes apart (stride),
      // 16 K elements = 16 rows of 128 bytes.
      const uint32_t a_lbo = p.a_mn ? BK * 128 : 16, a_sbo = 1024, a_kstep = p.a_mn ? UMMA_K * 128 : UMMA_K * 2;
      const uint32_t b_lbo = p.b_mn ? BK * 128 : 16, b_sbo = 1024, b_kstep = p.b_mn ? UMMA_K * 128 : UMMA_K * 2;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int ks = w / tiles;
        const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait(tempty + acc, acc_phase ^ 1);       // the epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full + stage, phase);             // TMA bytes of this stage have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo);
            const uint64_t bdesc = make_smem_desc(sb + k * b_kstep, b_lbo, b_sbo);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty + stage);                 // frees the stage once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull + acc);                     // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    // NP column parts per TMEM lane quarter (as many warps as the tile has 32-column chunks for, at most NQ); the
    // other epilogue warps of a narrow tile only keep the accumulator hand-shake going
    constexpr int NP = BN / 32 < NQ ? BN / 32 : NQ;
    const int half = (warp - 4) >> 2;                 // which part of the tile's columns it handles
    constexpr int HC = BN / NP;                       // columns per epilogue warp
    const bool idle = half >= NP;
    uint8_t* buf = units + (warp - 4) * UNIT_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int t = w % tiles, m_blk = t / p.n_tiles, n_blk = t % p.n_tiles;
      const int r0 = m_blk * BM + q * 32;             // first row of this warp's 32-row slab
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * HC);
      const int col0 = n_blk * BN + half * HC;
      float v[32];
      if (idle) {
        mbar_wait(tfull + acc, acc_phase);
      } else if (EPI == EPI_PLAIN && p.beta_one) {
        // C += acc (fp32): the old values of chunk c + 1 are requested while chunk c is processed
        float4 pre[8];
        prefetch_f32(reinterpret_cast<const float*>(p.C), p.ldc, r0, col0, p.M, lane, pre);
        mbar_wait(tfull + acc, acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HC / 32; ++c) {
          tmem_ld32(taddr + c * 32, v);
          stage_f32(buf, lane, v);
          flush_f32_add(buf, lane, reinterpret_cast<float*>(p.C), p.ldc, r0, col0 + c * 32, p.M, pre);
          if (c + 1 < HC / 32) prefetch_f32(reinterpret_cast<const float*>(p.C), p.ldc, r0, col0 + (c + 1) * 32, p.M, lane, pre);
        }
      } else if (EPI == EPI_PLAIN) {
        mbar_wait(tfull + acc, acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HC / 32; ++c) {
          tmem_ld32(taddr + c * 32, v);
          if (p.c_bf16) {
            stage_bf16(buf, lane, v);
            flush_bf16(buf, lane, reinterpret_cast<__nv_bfloat16*>(p.C), p.ldc, r0, col0 + c * 32, p.M);
          } else {
            stage_f32(buf, lane, v);
            if (p.atomic) reduce_f32(buf, lane, &tmC, col0 + c * 32, r0);     // split-K partial tile: C += tile in the L2
            else flush_f32(buf, lane, reinterpret_cast<float*>(p.C), p.ldc, r0, col0 + c * 32, p.M);
          }
        }
      } else if (EPI == EPI_QKV_WIN) {
        // in-projection of the attention (N = 3d, BN = d: the tile is 128 tokens of ONE of q, k, v).  q and k: + row of the
        // positional LUT (pos_table [Wq;Wk]^T + [bq;bk], row = the token's in-window cell), L2 norm per head (the 32
        // accumulator columns a thread holds are one head of 32 channels or two of 16), q also * log2(e) / tau; 1/|q|, 1/|k|
        // go to the per-row records of the backward.  Rows leave as bf16 in CSR (window) order, one plane per (tensor, slice).
        const int part = n_blk;
        const bool valid = r0 + lane < p.M;
        const int ti = valid ? __ldg(p.tok_info + r0 + lane) : 0;
        const int wrow = ti & 0x3ffffff, cell = (ti >> 26) & 63;
        const float qs = part == 0 ? 1.4426950408889634f / fmaxf(__ldg(p.tau), p.tau_min) : 1.f;
        const int colp0 = half * HC;                  // first column of this warp inside the tensor
        float4 pre[8];
        if (part < 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int cl = __shfl_sync(0xffffffffu, cell, (lane >> 3) + 4 * i);
            pre[i] = __ldg(reinterpret_cast<const float4*>(p.lut + (long long)cl * 2 * BN + part * BN + colp0) + (lane & 7));
          }
        }
        mbar_wait(tfull + acc, acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HC / 32; ++c) {
          tmem_ld32(taddr + c * 32, v);
          if (part < 2) {
            add_prefetched(buf, lane, pre, v);
            if (c + 1 < HC / 32) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int cl = __shfl_sync(0xffffffffu, cell, (lane >> 3) + 4 * i);
                pre[i] = __ldg(reinterpret_cast<const float4*>(p.lut + (long long)cl * 2 * BN + part * BN + colp0 + (c + 1) * 32) + (lane & 7));
              }
            }
            constexpr int HD = BN / 8, NH = 32 / HD;     // heads inside the 32 columns
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {
              float ss = 0.f;
#pragma unroll
              for (int j = 0; j < HD; ++j) ss = fmaf(v[hh * HD + j], v[hh * HD + j], ss);
              const float rn = rsqrtf(fmaxf(ss, 1e-24f));
              if (valid) p.lrr[(long long)wrow * 24 + 8 + 8 * part + (colp0 + c * 32) / HD + hh] = rn;
              const float f = rn * qs;
#pragma unroll
              for (int j = 0; j < HD; ++j) v[hh * HD + j] *= f;
            }
          }
          stage_bf16(buf, lane, v);
          flush_bf16_win(buf, lane, reinterpret_cast<__nv_bfloat16*>(p.C), part * (BN / 64), colp0 + c * 32, p.M, wrow, valid);
        }
      } else if (EPI == EPI_ROWS_WIN) {
        // plain bf16 output whose rows leave in CSR (window) order into planes plane0 + column / 64 of the window-major array
        // (dO of the attention: the out-projection's input gradient)
        const bool valid = r0 + lane < p.M;
        const int wrow = valid ? (__ldg(p.tok_info + r0 + lane) & 0x3ffffff) : 0;
        mbar_wait(tfull + acc, acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HC / 32; ++c) {
          tmem_ld32(taddr + c * 32, v);
          stage_bf16(buf, lane, v);
          flush_bf16_win(buf, lane, reinterpret_cast<__nv_bfloat16*>(p.C), p.plane0, col0 + c * 32, p.M, wrow, valid);
        }
      } else if (EPI == EPI_GELU) {
        // C2 = gelu(h + bias), C = gelu'(h + bias), both bf16, h = the accumulator.  The derivative - not h - is what the
        // backward pass keeps: it falls out of the forward's evaluation for one more FMA (from the fp32 accumulator, not
        // from a bf16-rounded h), and the backward epilogue (mode 3) becomes one multiply per element instead of a second
        // exp + reciprocal + polynomial (r2: 31 instructions per element on 8 epilogue warps, 61 us per launch, 0.29 of the
        // HBM roofline).
        float w[32];
        mbar_wait(tfull + acc, acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HC / 32; ++c) {
          tmem_ld32(taddr + c * 32, v);
          const float4* bp = reinterpret_cast<const float4*>(p.bias + col0 + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(bp + j);
            gelu_both_fast(v[4 * j] + b.x, v[4 * j], w[4 * j]); gelu_both_fast(v[4 * j + 1] + b.y, v[4 * j + 1], w[4 * j + 1]);
            gelu_both_fast(v[4 * j + 2] + b.z, v[4 * j + 2], w[4 * j + 2]); gelu_both_fast(v[4 * j + 3] + b.w, v[4 * j + 3], w[4 * j + 3]);
          }
          stage_bf16(buf, lane, w);
          flush_bf16(buf, lane, reinterpret_cast<__nv_bfloat16*>(p.C), p.ldc, r0, col0 + c * 32, p.M);
          stage_bf16(buf, lane, v);
          flush_bf16(buf, lane, reinterpret_cast<__nv_bfloat16*>(p.C2), p.ldc2, r0, col0 + c * 32, p.M);
        }
      } else if (EPI == EPI_GELU_BWD) {
        // acc = dg (gradient w.r.t. gelu(h + bias)); C = dh = dg * gelu'(h + bias) (bf16) with the derivative as mode 1 saved
        // it (h16), and the column sums of dh (the gradient of the bias) are added to p.colsum.  The saved rows of chunk
        // c + 1 are requested while chunk c is processed.
        uint4 pre[4];
        float hv[32];
        prefetch_bf16(reinterpret_cast<const __nv_bfloat16*>(p.h16), p.ldh, r0, col0, p.M, lane, pre);
        mbar_wait(tfull + acc, acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HC / 32; ++c) {
          tmem_ld32(taddr + c * 32, v);
          take_prefetched_bf16(buf, lane, pre, hv);
          if (c + 1 < HC / 32) prefetch_bf16(reinterpret_cast<const __nv_bfloat16*>(p.h16), p.ldh, r0, col0 + (c + 1) * 32, p.M, lane, pre);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= hv[j];
          stage_f32(buf, lane, v);               // rows past M hold zeros (their A rows were zero-filled by TMA)
          __syncwarp();
          atomicAdd(p.colsum + col0 + c * 32 + lane, unit_colsum(buf, lane));
          __syncwarp();
          stage_bf16(buf, lane, v);
          flush_bf16(buf, lane, reinterpret_cast<__nv_bfloat16*>(p.C), p.ldc, r0, col0 + c * 32, p.M);
        }
      } else {
        // z = acc + bias + residual row; y = LayerNorm(z) * gamma + beta.  N == BN: the tile holds whole rows, shared by
        // the two warps of a lane quarter (column halves), which exchange their partial sums through shared memory.  z
        // goes back into the accumulator columns (tcgen05.st): the two further passes read tensor memory only.
        float* st_sum = ln_stat;                 // [NP][128]
        float* st_sq = ln_stat + NQ * 128;       // [NP][128]
        const int rl = q * 32 + lane;       // row inside the tile
        float sum = 0.f;
        float4 pre[8];                      // residual rows of the next chunk, requested one chunk ahead
        prefetch_f32(p.res, BN, r0, col0, p.M, lane, pre);
        mbar_wait(tfull + acc, acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HC / 32; ++c) {
          tmem_ld32(taddr + c * 32, v);
          if (p.C) {                        // the raw GEMM output (bf16) is kept for the LayerNorm backward
            stage_bf16(buf, lane, v);
            flush_bf16(buf, lane, reinterpret_cast<__nv_bfloat16*>(p.C), p.ldc, r0, col0 + c * 32, p.M);
          }
          add_prefetched(buf, lane, pre, v);
          if (c + 1 < HC / 32) prefetch_f32(p.res, BN, r0, col0 + (c + 1) * 32, p.M, lane, pre);
          const float4* bp = reinterpret_cast<const float4*>(p.bias + col0 + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(bp + j);
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            sum += (v[4 * j] + v[4 * j + 1]) + (v[4 * j + 2] + v[4 * j + 3]);
          }
          tmem_st32(taddr + c * 32, v);
        }
        tmem_st_wait();
        st_sum[half * 128 + rl] = sum;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * NP) : "memory");       // the NP warps of this lane quarter
        float tot = 0.f;
#pragma unroll
        for (int pp = 0; pp < NP; ++pp) tot += st_sum[pp * 128 + rl];
        const float mean = tot * (1.f / BN);
        float sq = 0.f;
#pragma unroll 1
        for (int c = 0; c < HC / 32; ++c) {
          tmem_ld32(taddr + c * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) { const float dlt = v[j] - mean; sq = fmaf(dlt, dlt, sq); }
        }
        st_sq[half * 128 + rl] = sq;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * NP) : "memory");
        float totq = 0.f;
#pragma unroll
        for (int pp = 0; pp < NP; ++pp) totq += st_sq[pp * 128 + rl];
        const float rstd = rsqrtf(totq * (1.f / BN) + p.eps);
#pragma unroll 1
        for (int c = 0; c < HC / 32; ++c) {
          tmem_ld32(taddr + c * 32, v);
          const float4* gp = reinterpret_cast<const float4*>(p.gamma + col0 + c * 32);
          const float4* ep = reinterpret_cast<const float4*>(p.beta_ln + col0 + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g = __ldg(gp + j), e = __ldg(ep + j);
            v[4 * j] = fmaf((v[4 * j] - mean) * rstd, g.x, e.x); v[4 * j + 1] = fmaf((v[4 * j + 1] - mean) * rstd, g.y, e.y);
            v[4 * j + 2] = fmaf((v[4 * j + 2] - mean) * rstd, g.z, e.z); v[4 * j + 3] = fmaf((v[4 * j + 3] - mean) * rstd, g.w, e.w);
          }
          stage_f32(buf, lane, v);
          flush_f32(buf, lane, p.y32, BN, r0, col0 + c * 32, p.M);
          if (p.y16) {
            stage_bf16(buf, lane, v);
            flush_bf16(buf, lane, reinterpret_cast<__nv_bfloat16*>(p.y16), BN, r0, col0 + c * 32, p.M);
          }
        }
        if (half == 0 && r0 + lane < p.M) {
          p.mean[r0 + lane] = mean;
          p.rstd[r0 + lane] = rstd;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ==============================================================================================
// Weight gradient of the decoder's dense 3x3 convolution (spt_backbone_mae.py:45-49: Conv2d(384, 128, 3, padding=1) on the
// (B, Y, X, 384) NHWC map):  dW[co][ky][kx][ci] = sum_{b,y,x} dy[b,y,x,co] * in[b, y+ky-1, x+kx-1, ci].
// r1/r2 profile: cuDNN's wgrad read 10.8 GB of DRAM for 1.8 GB of operands (2.46 ms, 11 % of the step).  Here it is a
// GEMM with K = pixels whose operands are both "MN-major" (pixels are the rows of the NHWC arrays), so the TMA boxes of the
// own GEMM apply unchanged - the B operand of tap (ky, kx) is simply the box of the input map shifted by (ky-1, kx-1), halo
// and padding zero-filled by TMA's out-of-bounds handling.  No im2col, nothing transposed.
//   CTA role (ky, ci-chunk of 128): 9 roles; accumulators = the three kx taps x (128 co x 128 ci) fp32 = 384 TMEM columns.
//   A k-step = 64 pixels of one image row: A = dy (64 px x 128 co, 16 KB), B = three shifted input boxes (64 px x 128 ci
//   each): 12 tcgen05.mma (128 x 128 x 16) per k-step for 64 KB of operands - 98 FLOP per byte from L2.
//   The pixel rows are split across 16 groups of CTAs (9 x 16 = 144 CTAs, one wave); the CTAs of a group walk the same
//   rows at the same time, so the operands come from DRAM once and from the L2 nine times.  Every CTA adds its three
//   partial tiles to dW with TMA reduce-add.
//   SHIFT variant: the three kx taps read ONE box of 64 + 2 pixels (72 rows of 128 bytes); the tap's operand is that box
//   entered kx rows further down (shared-memory descriptor start + kx * 128 bytes - the 128B swizzle is a function of the
//   absolute shared-memory address, so a start inside a swizzle atom addresses the rows TMA wrote): 34 KB per k-step.
constexpr int CW_GROUPS = 16;
template <bool SHIFT> struct CwCfg {
  static constexpr int B_HALF = SHIFT ? 72 * 128 : 8192;                  // one 64-channel half of a B box
  static constexpr int STAGE_BYTES = 16384 + (SHIFT ? 2 * B_HALF : 3 * 16384);
  static constexpr int STAGES = SHIFT ? 5 : 3;
};

struct ConvWgradParams {
  int B, Y, X;
  int rows_total;        // B * Y
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// tmDy: (128 co, X, Y, B) bf16, box 64 x 64 x 1 x 1;  tmIn: (384 ci, X, Y, B) bf16, box 64 x 64 x 1 x 1;
// tmW: dW as (3456 = (ky, kx, ci), 128 co) fp32, box 32 x 32 (the reduce-add units of the GEMM epilogue)
constexpr int CW_THREADS = 128 + 32 * 8;      // eight epilogue warps, whatever the GEMM kernel uses
constexpr int CW_UNIT_BYTES_TOTAL = 8 * UNIT_BYTES;
template <bool SHIFT>
__global__ void __launch_bounds__(CW_THREADS, 1) conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy,
                                                                       const __grid_constant__ CUtensorMap tmIn,
                                                                       const __grid_constant__ CUtensorMap tmW, const ConvWgradParams p) {
  constexpr int CW_STAGE_BYTES = CwCfg<SHIFT>::STAGE_BYTES, CW_STAGES = CwCfg<SHIFT>::STAGES, B_HALF = CwCfg<SHIFT>::B_HALF;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* units = smem + CW_STAGES * CW_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(units + CW_UNIT_BYTES_TOTAL);
  uint64_t* empty = full + CW_STAGES;
  uint64_t* tfull = empty + CW_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int role = blockIdx.x % 9, grp = blockIdx.x / 9;
  const int ky = role / 3, cchunk = role % 3;
  // pixel rows (b, y) of this group, and the 64-pixel blocks of a row
  const int rows_per = (p.rows_total + CW_GROUPS - 1) / CW_GROUPS;
  const int row0 = grp * rows_per, row1 = min(row0 + rows_per, p.rows_total);
  const int xblocks = (p.X + 63) / 64;
  const int ksteps = (row1 > row0 ? row1 - row0 : 0) * xblocks;
  if (threadIdx.x == 0) {
    for (int i = 0; i < CW_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmDy) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmIn) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int r = row0; r < row1; ++r) {
        const int b = r / p.Y, y = r - b * p.Y;
        for (int xb = 0; xb < xblocks; ++xb) {
          mbar_wait(empty + stage, phase ^ 1);
          mbar_expect_tx(full + stage, CW_STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + stage * CW_STAGE_BYTES);
          // A: dy, 64 pixels x 128 co as two 64-channel boxes (MN-major: one pixel = one 128-byte row)
          tma_load_4d(sa, &tmDy, full + stage, 0, xb * 64, y, b);
          tma_load_4d(sa + 8192, &tmDy, full + stage, 64, xb * 64, y, b);
          // B: the input map shifted by (ky - 1, kx - 1); rows / columns outside the map arrive as zeros
          if (SHIFT) {
            const uint32_t sb = sa + 16384;          // pixels x0 - 1 .. x0 + 70 (box of 72 rows), both channel halves
            tma_load_4d(sb, &tmIn, full + stage, cchunk * 128, xb * 64 - 1, y + ky - 1, b);
            tma_load_4d(sb + B_HALF, &tmIn, full + stage, cchunk * 128 + 64, xb * 64 - 1, y + ky - 1, b);
          } else {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint32_t sb = sa + 16384 + kx * 16384;
              tma_load_4d(sb, &tmIn, full + stage, cchunk * 128, xb * 64 + kx - 1, y + ky - 1, b);
              tma_load_4d(sb + 8192, &tmIn, full + stage, cchunk * 128 + 64, xb * 64 + kx - 1, y + ky - 1, b);
            }
          }
          if (++stage == CW_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && ksteps > 0) {
      // both operands MN-major, N = 128 per tap
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(full + stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * CW_STAGE_BYTES);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t sb = sa + 16384 + (SHIFT ? kx * 128 : kx * 16384);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * (UMMA_K * 128), BK * 128, 1024);
            const uint64_t bdesc = make_smem_desc(sb + k * (UMMA_K * 128), B_HALF, 1024);
            umma_bf16(tmem_base + (uint32_t)(kx * 128), adesc, bdesc, idesc, (ks > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty + stage);
        if (++stage == CW_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull);
    }
    __syncwarp();
  } else if (warp >= 4 && ksteps > 0) {
    // epilogue: the three 128 x 128 fp32 partial tiles are added to dW by the L2 (TMA reduce-add), 32 x 32 units
    const int q = warp & 3, half = (warp - 4) >> 2;
    uint8_t* buf = units + (warp - 4) * UNIT_BYTES;
    mbar_wait(tfull, 0);
    tc_fence_after();
    float v[32];
#pragma unroll 1
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col = half * 64 + c * 32;                     // ci inside the chunk
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kx * 128 + col), v);
        stage_f32(buf, lane, v);
        reduce_f32(buf, lane, &tmW, (ky * 3 + kx) * 384 + cchunk * 128 + col, q * 32);
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;

EncodeTiledFn encode_fn() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      g_encode = (EncodeTiledFn)fn;
  });
  return g_encode;
}

// matrix stored row-major [outer, inner] with leading dimension ld (elements of `esize` bytes): box = 128 bytes of the inner
// dimension (SWIZZLE_128B) x box_outer rows; out-of-bounds elements read as zero and are not written
int make_map(CUtensorMap* m, const void* ptr, int esize, long long inner, long long outer, long long ld, int box_outer) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) { gdmae_set_error("cuTensorMapEncodeTiled is not available from the driver"); return GDMAE_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[200];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed (%d): ptr %p esize %d inner %lld outer %lld ld %lld box %d", (int)r, ptr, esize, inner,
             outer, ld, box_outer);
    gdmae_set_error(b);
    return GDMAE_ERR_CUDA;
  }
  return GDMAE_OK;
}

// NHWC activation (B, Y, X, C) bf16 as a 4-D tensor (C, X, Y, B): box = 64 channels x 64 pixels of one image row
int make_map_nhwc(CUtensorMap* m, const void* ptr, int C, int X, int Y, int B, int box_px = 64) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) { gdmae_set_error("cuTensorMapEncodeTiled is not available from the driver"); return GDMAE_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * X, (cuuint64_t)C * 2 * X * Y};
  cuuint32_t box[4] = {64, (cuuint32_t)box_px, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[160];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled (nhwc) failed (%d): ptr %p C %d X %d Y %d B %d", (int)r, ptr, C, X, Y, B);
    gdmae_set_error(b);
    return GDMAE_ERR_CUDA;
  }
  return GDMAE_OK;
}

template <int BN, int EPI>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const TcgParams& p, cudaStream_t st) {
  constexpr int STAGE_BYTES = BM * BK * 2 + BN * BK * 2;
  constexpr int STAGES = RING_BYTES / STAGE_BYTES;
  constexpr int SMEM = STAGES * STAGE_BYTES + UNIT_BYTES_TOTAL + 2 * NQ * 128 * 4 + (2 * STAGES + 4) * 8 + 16;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static bool configured[64] = {};
  int dev = 0;
  GDMAE_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(tcg_gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured[dev] = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.k_splits;
  const int grid = total < GDMAE_NUM_SMS ? total : GDMAE_NUM_SMS;
  tcg_gemm_kernel<BN, EPI><<<grid, NUM_THREADS, SMEM, st>>>(ta, tb, tc, p);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

}  // namespace

// dW (128, 3, 3, 384) fp32 += weight gradient of the 3 x 3 / padding 1 convolution 384 -> 128 on NHWC bf16 maps:
// dy (B, Y, X, 128), in (B, Y, X, 384).  accumulate = 0 zero-fills dW first (see include/gdmae_b200.h).
extern "C" int gdmae_conv3x3_wgrad(const void* dy_bf16, const void* in_bf16, int B, int Y, int X, int c_in, int c_out, float* dW,
                                   int accumulate, void* stream_) {
  GDMAE_CHECK_ARG(dy_bf16 && in_bf16 && dW && B > 0 && Y > 0 && X > 0 && c_in == 384 && c_out == 128);
  GDMAE_CHECK_ARG(((uintptr_t)dy_bf16 & 15) == 0 && ((uintptr_t)in_bf16 & 15) == 0 && ((uintptr_t)dW & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream_;
  if (!accumulate) GDMAE_CHECK_CUDA(cudaMemsetAsync(dW, 0, (size_t)c_out * 9 * c_in * sizeof(float), st));
  // GDMAE_WGRAD_SHIFT=0 selects the variant that loads one box per tap (development / A-B measurements)
  static const bool shift = [] { const char* e = getenv("GDMAE_WGRAD_SHIFT"); return !(e && e[0] == '0'); }();
  CUtensorMap tdy, tin, tw;
  int rc = make_map_nhwc(&tdy, dy_bf16, c_out, X, Y, B);
  if (rc) return rc;
  rc = make_map_nhwc(&tin, in_bf16, c_in, X, Y, B, shift ? 72 : 64);
  if (rc) return rc;
  rc = make_map(&tw, dW, 4, 9 * c_in, c_out, 9 * c_in, 32);
  if (rc) return rc;
  ConvWgradParams p = {B, Y, X, B * Y};
  int dev = 0;
  GDMAE_CHECK_CUDA(cudaGetDevice(&dev));
  static bool configured[2][64] = {};
  if (shift) {
    constexpr int SMEM = CwCfg<true>::STAGES * CwCfg<true>::STAGE_BYTES + CW_UNIT_BYTES_TOTAL + (2 * CwCfg<true>::STAGES + 1) * 8 + 16;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    if (dev >= 0 && dev < 64 && !configured[1][dev]) {
      GDMAE_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      configured[1][dev] = true;
    }
    conv3x3_wgrad_kernel<true><<<9 * CW_GROUPS, CW_THREADS, SMEM, st>>>(tdy, tin, tw, p);
  } else {
    constexpr int SMEM = CwCfg<false>::STAGES * CwCfg<false>::STAGE_BYTES + CW_UNIT_BYTES_TOTAL + (2 * CwCfg<false>::STAGES + 1) * 8 + 16;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    if (dev >= 0 && dev < 64 && !configured[0][dev]) {
      GDMAE_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      configured[0][dev] = true;
    }
    conv3x3_wgrad_kernel<false><<<9 * CW_GROUPS, CW_THREADS, SMEM, st>>>(tdy, tin, tw, p);
  }
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// cuTensorMapEncodeTiled from the driver (nullptr if unavailable) for the other TMA users of the library (sra_attention_tc.cu)
void* gdmae_tensor_map_encoder() { return (void*)encode_fn(); }

extern "C" int gdmae_tc_gemm_timeouts(int* out) {
  unsigned int v = 0;
  GDMAE_CHECK_CUDA(cudaMemcpyFromSymbol(&v, g_tcg_wait_timeouts, sizeof(v)));
  *out = (int)v;
  return GDMAE_OK;
}

// Row-major C (M,N) = op(A) (M,K) * op(B) (K,N), bf16 operands, fp32 accumulation (see include/gdmae_b200.h).
extern "C" int gdmae_tc_gemm(int transa, int transb, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                             int64_t ldb, void* C, int64_t ldc, int c_dtype, float beta, int split_k_atomic,
                             const gdmae_tc_epilogue* epi, void* stream_) {
  GDMAE_CHECK_ARG(M >= 0 && N > 0 && K > 0 && A && B && (c_dtype == 0 || c_dtype == 1));
  GDMAE_CHECK_ARG(beta == 0.f || (beta == 1.f && c_dtype == 0));
  if (M == 0) return GDMAE_OK;
  const int mode = epi ? epi->mode : 0;
  GDMAE_CHECK_ARG(mode >= 0 && mode <= 5);
  GDMAE_CHECK_ARG(N % 64 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0);
  GDMAE_CHECK_ARG(mode == EPI_LN || (C && ((uintptr_t)C & 15) == 0 && (ldc % 8 == 0 || mode >= EPI_QKV_WIN)));
  GDMAE_CHECK_ARG(!transa || M % 64 == 0);           // MN-major A is loaded as 64-wide boxes
  cudaStream_t st = (cudaStream_t)stream_;
  int BN = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 64);
  if (mode == EPI_LN) {
    GDMAE_CHECK_ARG((N == 128 || N == 256) && epi->res && epi->gamma && epi->beta_ln && epi->bias && epi->y32 && epi->mean && epi->rstd);
    GDMAE_CHECK_ARG(((uintptr_t)epi->y32 & 15) == 0 && ((uintptr_t)epi->y16 & 15) == 0 && ((uintptr_t)epi->res & 15) == 0);
    BN = (int)N;
  }
  if (mode == EPI_GELU_BWD)
    GDMAE_CHECK_ARG(epi->h16 && epi->colsum && c_dtype == 1 && N % 128 == 0 && epi->ldh % 8 == 0 && ((uintptr_t)epi->h16 & 15) == 0);
  if (mode == EPI_QKV_WIN) {
    // N = 3d with d in {128, 256}: one N tile per tensor
    GDMAE_CHECK_ARG((N == 384 || N == 768) && c_dtype == 1 && !split_k_atomic && epi->tok_info && epi->lut && epi->tau && epi->lrr);
    GDMAE_CHECK_ARG(((uintptr_t)epi->lut & 15) == 0 && M < (1ll << 26));
    BN = (int)(N / 3);
  }
  if (mode == EPI_ROWS_WIN) GDMAE_CHECK_ARG(N % 128 == 0 && c_dtype == 1 && !split_k_atomic && epi->tok_info && epi->plane0 >= 0 && M < (1ll << 26));
  if (mode == EPI_GELU) GDMAE_CHECK_ARG(epi->bias && epi->c2 && c_dtype == 1 && N % 128 == 0 && epi->ldc2 % 8 == 0 && ((uintptr_t)epi->c2 & 15) == 0);
  TcgParams p = {};
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.m_tiles = (int)((M + BM - 1) / BM);
  p.n_tiles = (int)(N / BN);
  p.kb_total = (int)((K + BK - 1) / BK);
  p.k_splits = 1;
  if (split_k_atomic) {
    GDMAE_CHECK_ARG(mode == 0 && c_dtype == 0);
    // one wave of work items: every CTA reduces exactly one partial tile into C
    const int tiles = p.m_tiles * p.n_tiles;
    int want = GDMAE_NUM_SMS / tiles > 0 ? GDMAE_NUM_SMS / tiles : 1;
    int max_splits = p.kb_total / 4 > 0 ? p.kb_total / 4 : 1;      // at least four K blocks per split
    p.k_splits = want < max_splits ? want : max_splits;
    if (beta == 0.f) {
      GDMAE_CHECK_ARG(ldc == N);
      GDMAE_CHECK_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st));
    }
  }
  p.kb_per_split = (p.kb_total + p.k_splits - 1) / p.k_splits;
  p.k_splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;   // no empty split
  p.a_mn = transa ? 1 : 0;
  p.b_mn = transb ? 0 : 1;
  p.C = C; p.ldc = ldc; p.c_bf16 = c_dtype; p.beta_one = (beta == 1.f && !split_k_atomic) ? 1 : 0; p.atomic = split_k_atomic ? 1 : 0;
  if (epi) {
    p.bias = epi->bias; p.C2 = epi->c2; p.ldc2 = epi->ldc2; p.res = epi->res; p.gamma = epi->gamma; p.beta_ln = epi->beta_ln;
    p.eps = epi->eps; p.y32 = epi->y32; p.y16 = epi->y16; p.mean = epi->mean; p.rstd = epi->rstd;
    p.h16 = epi->h16; p.ldh = epi->ldh; p.colsum = epi->colsum;
    if (mode >= EPI_QKV_WIN) {
      p.tok_info = epi->tok_info; p.lut = epi->lut; p.tau = epi->tau; p.tau_min = epi->tau_min; p.lrr = epi->lrr; p.plane0 = epi->plane0;
    }
  }
  CUtensorMap ta, tb, tc;
  int rc;
  // A: (M,K) row-major = K-major, box 64 x 128 rows;  or stored (K,M) = MN-major, boxes of 64 (M) x 64 (K)
  rc = !transa ? make_map(&ta, A, 2, K, M, lda, BM) : make_map(&ta, A, 2, M, K, lda, BK);
  if (rc) return rc;
  // B: stored (N,K) = K-major, box 64 x BN rows;  or stored (K,N) = MN-major, boxes of 64 (N) x 64 (K)
  rc = transb ? make_map(&tb, B, 2, K, N, ldb, BN) : make_map(&tb, B, 2, N, K, ldb, BK);
  if (rc) return rc;
  // split-K partial tiles are reduced into C by TMA in units of 32 rows x 32 fp32 columns
  if (split_k_atomic) {
    rc = make_map(&tc, C, 4, N, M, ldc, 32);
    if (rc) return rc;
  } else {
    tc = tb;
  }
  if (mode == EPI_LN && C) GDMAE_CHECK_ARG(c_dtype == 1 && ((uintptr_t)C & 15) == 0 && ldc % 8 == 0);
  // bench-only span (bench.py `kernels`: tc_gemm_bn<BN>_mode<m>).  Algorithmic bytes: both operands once + the result once,
  // + what the fused epilogue must move besides (old C of `C +=`; fp32 residual in, fp32 + bf16 row out of the LayerNorm
  // mode; the bf16 pre-activation out / in of the two GELU modes)
  GdmaeSpan span(st);
  const long long s_out = c_dtype ? 2 : 4;
  long long bytes = M * K * 2 + K * N * 2 + (split_k_atomic ? 0 : M * N * s_out) + (p.beta_one ? M * N * s_out : 0);
  if (mode == EPI_LN) bytes += M * N * (4 + 4 + 2);
  if (mode == EPI_GELU || mode == EPI_GELU_BWD) bytes += M * N * 2;
  auto dispatch = [&]() -> int {
    if (mode == EPI_PLAIN) {
      if (BN == 256) return launch<256, EPI_PLAIN>(ta, tb, tc, p, st);
      if (BN == 128) return launch<128, EPI_PLAIN>(ta, tb, tc, p, st);
      return launch<64, EPI_PLAIN>(ta, tb, tc, p, st);
    }
    if (mode == EPI_QKV_WIN) {
      if (BN == 256) return launch<256, EPI_QKV_WIN>(ta, tb, tc, p, st);
      return launch<128, EPI_QKV_WIN>(ta, tb, tc, p, st);
    }
    if (mode == EPI_ROWS_WIN) {
      if (BN == 256) return launch<256, EPI_ROWS_WIN>(ta, tb, tc, p, st);
      return launch<128, EPI_ROWS_WIN>(ta, tb, tc, p, st);
    }
    if (mode == EPI_GELU_BWD) {
      if (BN == 256) return launch<256, EPI_GELU_BWD>(ta, tb, tc, p, st);
      return launch<128, EPI_GELU_BWD>(ta, tb, tc, p, st);
    }
    if (mode == EPI_GELU) {
      if (BN == 256) return launch<256, EPI_GELU>(ta, tb, tc, p, st);
      return launch<128, EPI_GELU>(ta, tb, tc, p, st);
    }
    if (BN == 256) return launch<256, EPI_LN>(ta, tb, tc, p, st);
    return launch<128, EPI_LN>(ta, tb, tc, p, st);
  };
  rc = dispatch();
  span.end(3, BN * 10 + mode, M, bytes);
  return rc;
}
