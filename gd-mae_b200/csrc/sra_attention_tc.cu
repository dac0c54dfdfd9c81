// Sparse Regional Attention forward on tensor cores (TF32 mma.sync m16n8k8, fp32 accumulate).
//
// Same operator and data layout as sra_attention.cu (reference: cosine_msa.py:114-176,
// sst_basic_block.py:22-54, sst_utils.py:107-181): flat tokens, CSR windows, 64-row positional
// LUT, nothing padded in HBM.  The fp32 SIMT kernel is the parity path; at pyramid scales 2 and 3
// (13-28 tokens per window) it is FMA-bound, so the dense contractions inside each window -
// S = Q K^T and O = P V - move to the tensor cores here:
//   * a CTA owns the windows that start in a 32-row bin (<= 95 rows) and a 128-channel slice;
//     q-hat (scaled by 1/tau), k-hat and v of those rows are staged once in shared memory with
//     coalesced 128-bit loads, rounded to TF32 (pitch 132 floats: all fragment loads conflict free);
//   * one warp per (window, head): S tiles (16 x 8) accumulate in registers, the softmax runs on
//     the C fragments (quad shuffles), and P feeds the second MMA straight from registers - the
//     C->A fragment mismatch is absorbed by permuting the key index of the V fragment
//     (A col t <-> key 2t, A col t+4 <-> key 2t+1), so P never touches shared memory;
//   * rows past the window end are masked with -inf; 16 zero pad rows follow the bin.
// TF32 operands (10-bit mantissa) put this kernel in the "bf16/tf32" performance configuration;
// parity tests use the fp32 kernel and compare this one at 2e-3.
#include "common.cuh"

#define TC_EPS 1e-12f
#define TC_BIN 32
#define TC_ROWS 112           // 95 rows + 16 zero pad rows + 1
#define TC_SLICE 128
#define TC_PITCH 132
#define TC_MAXWIN 96
#define TC_SMEM_BYTES (3 * TC_ROWS * TC_PITCH * 4 + TC_MAXWIN * 16 + TC_MAXWIN * 4 + 64)

struct TcArgs {
  const float* qkv;
  const float* lut;
  const int4* row_info;
  const float* tau;
  float tau_min;
  int N, d;
};

__device__ __forceinline__ float to_tf32(float x) {
  unsigned int u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const float (&a)[4], float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                 "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

__device__ __forceinline__ int tc_first_start(const int4* __restrict__ info, int t, int N) {
  if (t >= N) return N;
  int4 r = __ldg(info + t);
  return r.y == t ? t : r.z;
}

template <int HD>
__global__ void __launch_bounds__(512, 1) sra_fwd_tc_kernel(TcArgs a, float* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = (float*)smem_raw;
  float* sk = sq + TC_ROWS * TC_PITCH;
  float* sv = sk + TC_ROWS * TC_PITCH;
  int4* sinfo = (int4*)(sv + TC_ROWS * TC_PITCH);
  int* swin = (int*)(sinfo + TC_MAXWIN);
  int* hdr = swin + TC_MAXWIN;  // [0] row0, [1] row1, [2] number of windows
  constexpr int HS = TC_SLICE / HD;
  constexpr int LPH = HD / 4;
  constexpr int KS = HD / 8;    // k-steps of Q K^T, n-tiles of the output
  const int d = a.d, col = blockIdx.y * TC_SLICE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid < 2) hdr[tid] = tc_first_start(a.row_info, (blockIdx.x + tid) * TC_BIN, a.N);
  __syncthreads();
  const int row0 = hdr[0], R = hdr[1] - row0;
  if (R == 0) return;
  for (int r = tid; r < R; r += blockDim.x) {
    int4 v = __ldg(a.row_info + row0 + r);
    v.y -= row0;
    v.z -= row0;
    sinfo[r] = v;
  }
  __syncthreads();
  // window starts of the bin (warp 0): row r opens a window iff its window's first row is r
  if (warp == 0) {
    int base = 0;
    for (int r0 = 0; r0 < R; r0 += 32) {
      int r = r0 + lane;
      bool st = r < R && sinfo[r].y == r;
      unsigned m = __ballot_sync(0xffffffffu, st);
      if (st) swin[base + __popc(m & ((1u << lane) - 1u))] = r;
      base += __popc(m);
    }
    if (lane == 0) hdr[2] = base;
  }
  // ---- stage q-hat/tau, k-hat, v (TF32 rounded); zero the 16 pad rows
  const float inv_tau = 1.f / fmaxf(__ldg(a.tau), a.tau_min);
  const int Rp = min(R + 16, TC_ROWS);
  const int steps = (Rp * 32 + blockDim.x - 1) / blockDim.x;
  for (int it = 0; it < steps; ++it) {
    int idx = it * blockDim.x + tid;
    int r = idx >> 5, c4 = idx & 31;
    bool valid = r < R;
    int4 inf = sinfo[valid ? r : 0];
    const float* base = a.qkv + (long long)inf.x * 3 * d + col;
    const float* lb = a.lut + inf.w * 2 * d + col;
    float4 q = __ldg(reinterpret_cast<const float4*>(base) + c4);
    float4 k = __ldg(reinterpret_cast<const float4*>(base + d) + c4);
    float4 v = __ldg(reinterpret_cast<const float4*>(base + 2 * d) + c4);
    float4 lq = __ldg(reinterpret_cast<const float4*>(lb) + c4);
    float4 lk = __ldg(reinterpret_cast<const float4*>(lb + d) + c4);
    q.x += lq.x; q.y += lq.y; q.z += lq.z; q.w += lq.w;
    k.x += lk.x; k.y += lk.y; k.z += lk.z; k.w += lk.w;
    float sq2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    float sk2 = k.x * k.x + k.y * k.y + k.z * k.z + k.w * k.w;
#pragma unroll
    for (int o = 1; o < LPH; o <<= 1) {
      sq2 += __shfl_xor_sync(0xffffffffu, sq2, o);
      sk2 += __shfl_xor_sync(0xffffffffu, sk2, o);
    }
    float fq = inv_tau / fmaxf(sqrtf(sq2), TC_EPS), fk = 1.f / fmaxf(sqrtf(sk2), TC_EPS);
    if (r < Rp) {
      float4 oq = make_float4(0.f, 0.f, 0.f, 0.f), ok = oq, ov = oq;
      if (valid) {
        oq = make_float4(to_tf32(q.x * fq), to_tf32(q.y * fq), to_tf32(q.z * fq), to_tf32(q.w * fq));
        ok = make_float4(to_tf32(k.x * fk), to_tf32(k.y * fk), to_tf32(k.z * fk), to_tf32(k.w * fk));
        ov = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      }
      *reinterpret_cast<float4*>(sq + r * TC_PITCH + 4 * c4) = oq;
      *reinterpret_cast<float4*>(sk + r * TC_PITCH + 4 * c4) = ok;
      *reinterpret_cast<float4*>(sv + r * TC_PITCH + 4 * c4) = ov;
    }
  }
  __syncthreads();

  // ---- one warp per (window, head)
  const int g = lane >> 2, t = lane & 3;
  const int units = hdr[2] * HS;
  const int nwarps = blockDim.x >> 5;
  for (int u = warp; u < units; u += nwarps) {
    const int h = u % HS;
    const int s = swin[u / HS];
    const int n = sinfo[s].z - s;
    const int NT = (n + 7) >> 3;
    const int ch = h * HD;
    for (int m0 = 0; m0 < n; m0 += 16) {
      // Q fragments of this 16-row tile
      float qa[KS][4];
      const float* qr0 = sq + (s + m0 + g) * TC_PITCH + ch + t;
      const float* qr1 = qr0 + 8 * TC_PITCH;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        qa[ks][0] = qr0[8 * ks];
        qa[ks][1] = qr1[8 * ks];
        qa[ks][2] = qr0[8 * ks + 4];
        qa[ks][3] = qr1[8 * ks + 4];
      }
      // S = Q K^T  (up to 8 key tiles of 8)
      float c[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
        if (nt < NT) {
          const float* kr = sk + (s + 8 * nt + g) * TC_PITCH + ch + t;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) mma_tf32(c[nt], qa[ks], kr[8 * ks], kr[8 * ks + 4]);
        }
      }
      // mask keys >= n, row max (rows g and g+8 of the tile)
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < NT) {
          int k0 = 8 * nt + 2 * t;
          if (k0 >= n) { c[nt][0] = -INFINITY; c[nt][2] = -INFINITY; }
          if (k0 + 1 >= n) { c[nt][1] = -INFINITY; c[nt][3] = -INFINITY; }
          mx0 = fmaxf(mx0, fmaxf(c[nt][0], c[nt][1]));
          mx1 = fmaxf(mx1, fmaxf(c[nt][2], c[nt][3]));
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < NT) {
          c[nt][0] = __expf(c[nt][0] - mx0); c[nt][1] = __expf(c[nt][1] - mx0);
          c[nt][2] = __expf(c[nt][2] - mx1); c[nt][3] = __expf(c[nt][3] - mx1);
          l0 += c[nt][0] + c[nt][1];
          l1 += c[nt][2] + c[nt][3];
        }
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      // O = P V : P from the C fragments (key permutation: A col t <-> key 2t, A col t+4 <-> key 2t+1)
      float o[KS][4];
#pragma unroll
      for (int nd = 0; nd < KS; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 8; ++kt) {
        if (kt < NT) {
          float pa[4] = {to_tf32(c[kt][0]), to_tf32(c[kt][2]), to_tf32(c[kt][1]), to_tf32(c[kt][3])};
          const float* v0 = sv + (s + 8 * kt + 2 * t) * TC_PITCH + ch + g;
          const float* v1 = v0 + TC_PITCH;
#pragma unroll
          for (int nd = 0; nd < KS; ++nd) mma_tf32(o[nd], pa, v0[8 * nd], v1[8 * nd]);
        }
      }
      // write rows g and g+8 of the tile
      const float il0 = 1.f / l0, il1 = 1.f / l1;
      const int rA = m0 + g, rB = m0 + g + 8;
      if (rA < n) {
        int tok = sinfo[s + rA].x;
        float* dst = out + (long long)tok * d + col + ch + 2 * t;
#pragma unroll
        for (int nd = 0; nd < KS; ++nd) *reinterpret_cast<float2*>(dst + 8 * nd) = make_float2(o[nd][0] * il0, o[nd][1] * il0);
        if (t == 0) lse[(long long)tok * 8 + blockIdx.y * HS + h] = mx0 + __logf(l0);
      }
      if (rB < n) {
        int tok = sinfo[s + rB].x;
        float* dst = out + (long long)tok * d + col + ch + 2 * t;
#pragma unroll
        for (int nd = 0; nd < KS; ++nd) *reinterpret_cast<float2*>(dst + 8 * nd) = make_float2(o[nd][2] * il1, o[nd][3] * il1);
        if (t == 0) lse[(long long)tok * 8 + blockIdx.y * HS + h] = mx1 + __logf(l1);
      }
    }
  }
}

// Tensor-core (TF32) variant of gdmae_sra_attention_fwd: same arguments, same outputs.
extern "C" int gdmae_sra_attention_fwd_tc(const float* qkv, const float* lut, const int32_t* row_info, int64_t N, int d, int nhead,
                                          const float* tau, float tau_min, float* out, float* lse, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 27) && nhead == 8 && (d == 128 || d == 256) && ((uintptr_t)row_info % 16) == 0);
  if (N == 0) return GDMAE_OK;
  static bool attr_done = false;
  if (!attr_done) {
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_done = true;
  }
  TcArgs a{qkv, lut, (const int4*)row_info, tau, tau_min, (int)N, d};
  dim3 grid(gdmae_div_up(N, TC_BIN), d / TC_SLICE);
  cudaStream_t st = (cudaStream_t)stream_;
  if (d == 128) sra_fwd_tc_kernel<16><<<grid, 512, TC_SMEM_BYTES, st>>>(a, out, lse);
  else sra_fwd_tc_kernel<32><<<grid, 512, TC_SMEM_BYTES, st>>>(a, out, lse);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
