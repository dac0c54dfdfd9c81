// Sparse Regional Attention forward on tensor cores (TF32 mma.sync m16n8k8, fp32 accumulate).
//
// Same operator and data layout as sra_attention.cu (reference: cosine_msa.py:114-176,
// sst_basic_block.py:22-54, sst_utils.py:107-181): flat tokens, CSR windows, 64-row positional
// LUT, nothing padded in HBM.  The fp32 SIMT kernel is the parity path; at pyramid scales 2 and 3
// (13-28 tokens per window) its partner loop is instruction bound (r1 ncu: 107 M warp
// instructions for 16 M query-key-head triples) and each of its CTAs pays the whole
// row_info -> q/k/v -> LUT dependent-load chain before any math.  This kernel
//   * is persistent: one CTA per SM owns a 64-channel slice (2 heads of 32 or 4 heads of 16) and
//     walks bins of 32 CSR rows (the windows that START in the bin, <= 95 rows).  The slice's 64 x 128
//     positional LUT lives in shared memory for the whole kernel;
//   * is software pipelined with cp.async: while bin k is normalised and multiplied, the raw q/k/v
//     rows of bin k+1 and the row_info records of bin k+2 are in flight (two data buffers, three
//     record buffers), so the dependent-load chain is off the critical path;
//   * normalises in place: +LUT, L2-normalise q and k per head, fold 1/tau into q, round all three
//     operands to TF32; 16 zero rows follow the bin so MMA tiles may overrun a window;
//   * runs the two contractions of a window - S = Q K^T and O = P V - on the tensor cores, one warp
//     per (window, head, 16-query tile).  S tiles stay in registers, the softmax runs on the C
//     fragments (quad shuffles), and P feeds the second MMA straight from registers: the C->A
//     fragment mismatch is absorbed by permuting the key index of the V fragment
//     (A col t <-> key 2t, A col t+4 <-> key 2t+1), so P never touches shared memory;
//   * pitch 68 floats: every fragment load is bank-conflict free.
// TF32 operands (10-bit mantissa) put this kernel in the bf16/tf32 performance configuration;
// parity tests use the fp32 kernel and hold this one to 4e-3.
#include "common.cuh"
#include <cuda_bf16.h>

#define TC_EPS 1e-12f
#define TC_BIN 32
#define TC_ROWS 112            // 95 rows + 16 zero pad rows + 1
#define TC_INFO (TC_BIN + 64)  // row_info records cached per bin
#define TC_SLICE 64
#define TC_C4 (TC_SLICE / 4)
#define TC_PITCH (TC_SLICE + 4)
#define TC_THREADS 512
#define TC_BUF_FLOATS (3 * TC_ROWS * TC_PITCH)
#define TC_UNITS 192           // >= heads x 16-query tiles of one bin
#define TC_SMEM_BYTES (64 * 2 * TC_SLICE * 4 + 2 * TC_BUF_FLOATS * 4 + 3 * TC_INFO * 16 + TC_UNITS * 4 + 64)

struct TcArgs {
  const float* qkv;
  const float* lut;
  const int4* row_info;
  const float* tau;
  const float* bv;   // (d) value bias added to the output, nullable
  float tau_min;
  int N, d;
  int out_bf16;
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

__device__ __forceinline__ void tc_cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void tc_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// rows [row0, row1) = the windows that start inside the bin; inf = records of rows bin .. bin+95
__device__ __forceinline__ void tc_bin_range(const int4* inf, int bin, int N, int& row0, int& R) {
  int4 f = inf[0];
  row0 = (f.y == bin) ? bin : f.z;
  int row1 = N;
  if (bin + TC_BIN < N) {
    int4 l = inf[TC_BIN];
    row1 = (l.y == bin + TC_BIN) ? bin + TC_BIN : l.z;
  }
  R = row1 - row0;
}

__device__ __forceinline__ void tc_issue_info(int4* inf, const int4* row_info, int bin, int N, int tid) {
  if (tid < TC_INFO && bin + tid < N) tc_cp_async16(inf + tid, row_info + bin + tid);
}

__device__ __forceinline__ void tc_issue_rows(float* buf, const int4* inf, const float* qkv, int d, int col, int row0, int bin, int R,
                                              int tid) {
  const int shift = row0 - bin;
  float* sq = buf;
  float* sk = sq + TC_ROWS * TC_PITCH;
  float* sv = sk + TC_ROWS * TC_PITCH;
  for (int idx = tid; idx < R * TC_C4; idx += TC_THREADS) {
    int r = idx / TC_C4, c4 = idx % TC_C4;
    const float* base = qkv + (long long)inf[r + shift].x * 3 * d + col + 4 * c4;
    tc_cp_async16(sq + r * TC_PITCH + 4 * c4, base);
    tc_cp_async16(sk + r * TC_PITCH + 4 * c4, base + d);
    tc_cp_async16(sv + r * TC_PITCH + 4 * c4, base + 2 * d);
  }
}

template <int HD>
__global__ void __launch_bounds__(TC_THREADS, 1) sra_fwd_tc_kernel(TcArgs a, void* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* slut = (float*)smem_raw;                            // [64][128]: q part | k part of this slice
  float* sbuf = slut + 64 * 2 * TC_SLICE;                    // two data buffers
  int4* sinfo_all = (int4*)(sbuf + 2 * TC_BUF_FLOATS);       // three record buffers
  int* sunit = (int*)(sinfo_all + 3 * TC_INFO);              // work units of the current bin: window start | n | head | tile
  int* hdr = sunit + TC_UNITS;                               // [0] number of units
  constexpr int HS = TC_SLICE / HD;
  constexpr int LPH = HD / 4;
  constexpr int KS = HD / 8;    // k-steps of Q K^T, n-tiles of the output
  const int d = a.d;
  const int nsl = d / TC_SLICE;
  const int sl = blockIdx.x % nsl, cta = blockIdx.x / nsl, ncta = gridDim.x / nsl;
  const int col = sl * TC_SLICE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbins = (a.N + TC_BIN - 1) / TC_BIN;
  const int nwarps = TC_THREADS >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (cta >= nbins) return;

  // ---- prologue: LUT slice, records of the first two bins, rows of the first bin
  for (int idx = tid; idx < 64 * 2 * TC_C4; idx += TC_THREADS) {
    int pos = idx / (2 * TC_C4), rem = idx % (2 * TC_C4);
    int part = rem / TC_C4, c4 = rem % TC_C4;
    tc_cp_async16(slut + pos * 2 * TC_SLICE + part * TC_SLICE + 4 * c4, a.lut + (long long)pos * 2 * d + part * d + col + 4 * c4);
  }
  tc_issue_info(sinfo_all, a.row_info, cta * TC_BIN, a.N, tid);
  if (cta + ncta < nbins) tc_issue_info(sinfo_all + TC_INFO, a.row_info, (cta + ncta) * TC_BIN, a.N, tid);
  tc_commit();
  tc_wait_all();
  __syncthreads();
  {
    int row0, R;
    tc_bin_range(sinfo_all, cta * TC_BIN, a.N, row0, R);
    tc_issue_rows(sbuf, sinfo_all, a.qkv, d, col, row0, cta * TC_BIN, R, tid);
    tc_commit();
  }
  const float inv_tau = 1.f / fmaxf(__ldg(a.tau), a.tau_min);

  for (int k = 0;; ++k) {
    const int bi = cta + k * ncta;
    if (bi >= nbins) break;
    const int bin = bi * TC_BIN;
    const int4* sinfo = sinfo_all + (k % 3) * TC_INFO;
    float* sq = sbuf + (k & 1) * TC_BUF_FLOATS;
    float* sk = sq + TC_ROWS * TC_PITCH;
    float* sv = sk + TC_ROWS * TC_PITCH;
    tc_wait_all();       // rows of bin k, records of bin k+1
    __syncthreads();     // ... visible to all; everyone is done with bin k-1 (its buffers are free)
    // ---- keep the pipe full: rows of bin k+1, records of bin k+2
    if (bi + ncta < nbins) {
      const int4* ninfo = sinfo_all + ((k + 1) % 3) * TC_INFO;
      int nrow0, nR;
      tc_bin_range(ninfo, (bi + ncta) * TC_BIN, a.N, nrow0, nR);
      tc_issue_rows(sbuf + ((k + 1) & 1) * TC_BUF_FLOATS, ninfo, a.qkv, d, col, nrow0, (bi + ncta) * TC_BIN, nR, tid);
      if (bi + 2 * ncta < nbins) tc_issue_info(sinfo_all + ((k + 2) % 3) * TC_INFO, a.row_info, (bi + 2 * ncta) * TC_BIN, a.N, tid);
    }
    tc_commit();

    int row0, R;
    tc_bin_range(sinfo, bin, a.N, row0, R);
    if (R == 0) continue;
    const int shift = row0 - bin;
    // ---- work units of the bin: (window, head, 16-query tile), built by one warp while the others normalise
    if (warp == nwarps - 1) {
      int base = 0;
      for (int r0 = 0; r0 < R; r0 += 32) {
        int r = r0 + lane;
        int4 rec = sinfo[min(r, R - 1) + shift];
        bool st = r < R && rec.y == row0 + r;
        int n = rec.z - rec.y;
        int cnt = st ? HS * ((n + 15) >> 4) : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        int off = base + incl - cnt;
        for (int j = 0; j < cnt; ++j) sunit[off + j] = r | (n << 8) | ((j % HS) << 16) | ((j / HS) << 20);
        base += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) hdr[0] = base;
    }
    // ---- in place: +LUT, normalise q (x 1/tau) and k per head, round q, k, v to TF32; zero 16 pad rows
    {
      const int Rp = min(R + 16, TC_ROWS);
      const int steps = (Rp * TC_C4 + TC_THREADS - 1) / TC_THREADS;
      for (int it = 0; it < steps; ++it) {
        int idx = it * TC_THREADS + tid;
        int r = idx / TC_C4, c4 = idx % TC_C4;
        bool valid = r < R;
        int rr = valid ? r : 0;
        const float* lb = slut + sinfo[rr + shift].w * 2 * TC_SLICE + 4 * c4;
        float4 q = *reinterpret_cast<const float4*>(sq + rr * TC_PITCH + 4 * c4);
        float4 kk = *reinterpret_cast<const float4*>(sk + rr * TC_PITCH + 4 * c4);
        float4 v = *reinterpret_cast<const float4*>(sv + rr * TC_PITCH + 4 * c4);
        float4 lq = *reinterpret_cast<const float4*>(lb);
        float4 lk = *reinterpret_cast<const float4*>(lb + TC_SLICE);
        q.x += lq.x; q.y += lq.y; q.z += lq.z; q.w += lq.w;
        kk.x += lk.x; kk.y += lk.y; kk.z += lk.z; kk.w += lk.w;
        float sq2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
        float sk2 = kk.x * kk.x + kk.y * kk.y + kk.z * kk.z + kk.w * kk.w;
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
            ok = make_float4(to_tf32(kk.x * fk), to_tf32(kk.y * fk), to_tf32(kk.z * fk), to_tf32(kk.w * fk));
            ov = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
          }
          *reinterpret_cast<float4*>(sq + r * TC_PITCH + 4 * c4) = oq;
          *reinterpret_cast<float4*>(sk + r * TC_PITCH + 4 * c4) = ok;
          *reinterpret_cast<float4*>(sv + r * TC_PITCH + 4 * c4) = ov;
        }
      }
    }
    __syncthreads();

    // ---- one warp per (window, head, 16-query tile), round-robin
    const int nunits = hdr[0];
    for (int u = warp; u < nunits; u += nwarps) {
      {
        const int code = sunit[u];
        const int s = code & 0xff, n = (code >> 8) & 0xff, h = (code >> 16) & 0xf, mt = code >> 20;
        const int NT = (n + 7) >> 3;
        const int ch = h * HD;
        const int m0 = mt * 16;
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
        const float il0 = 1.f / l0, il1 = 1.f / l1;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int rq = m0 + g + 8 * half;
          if (rq < n) {
            const int tok = sinfo[s + rq + shift].x;
            const float il = half ? il1 : il0;
            const long long e0 = (long long)tok * d + col + ch + 2 * t;
#pragma unroll
            for (int nd = 0; nd < KS; ++nd) {
              float x0 = o[nd][2 * half] * il, x1 = o[nd][2 * half + 1] * il;
              if (a.bv) { x0 += __ldg(a.bv + col + ch + 8 * nd + 2 * t); x1 += __ldg(a.bv + col + ch + 8 * nd + 2 * t + 1); }
              if (a.out_bf16) *reinterpret_cast<__nv_bfloat162*>((__nv_bfloat16*)out + e0 + 8 * nd) = __floats2bfloat162_rn(x0, x1);
              else *reinterpret_cast<float2*>((float*)out + e0 + 8 * nd) = make_float2(x0, x1);
            }
            if (t == 0) lse[(long long)tok * 8 + sl * HS + h] = (half ? mx1 : mx0) + __logf(half ? l1 : l0);
          }
        }
      }
    }
  }
  tc_wait_all();
}

// Tensor-core (TF32) variant of gdmae_sra_attention_fwd: same arguments, same outputs.
extern "C" int gdmae_sra_attention_fwd_tc(const float* qkv, const float* lut, const int32_t* row_info, int64_t N, int d, int nhead,
                                          const float* tau, float tau_min, const float* bv, int io_bf16, void* out, float* lse,
                                          void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 27) && nhead == 8 && (d == 128 || d == 256) && ((uintptr_t)row_info % 16) == 0 &&
                  ((uintptr_t)qkv % 16) == 0 && ((uintptr_t)lut % 16) == 0);
  if (N == 0) return GDMAE_OK;
  static bool attr_done = false;
  if (!attr_done) {
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_done = true;
  }
  TcArgs a{qkv, lut, (const int4*)row_info, tau, bv, tau_min, (int)N, d, io_bf16};
  cudaStream_t st = (cudaStream_t)stream_;
  // one CTA per SM; 148 is a multiple of the 2 (d = 128) and 4 (d = 256) channel slices
  if (d == 128) sra_fwd_tc_kernel<16><<<GDMAE_NUM_SMS, TC_THREADS, TC_SMEM_BYTES, st>>>(a, out, lse);
  else sra_fwd_tc_kernel<32><<<GDMAE_NUM_SMS, TC_THREADS, TC_SMEM_BYTES, st>>>(a, out, lse);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
