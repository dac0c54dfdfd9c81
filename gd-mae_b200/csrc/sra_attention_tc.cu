// Sparse Regional Attention on the tensor cores: bf16 mma.sync m16n8k16 with fp32 accumulation,
// operands fetched with ldmatrix - the kernel of the bf16 configuration.
//
// Same operator and data layout as sra_attention.cu (reference: cosine_msa.py:114-176,
// sst_basic_block.py:22-54, sst_utils.py:107-181): flat tokens, CSR windows, 64-row positional
// LUT, nothing padded in HBM; here q/k/v arrive as bf16 (the in-projection GEMM writes them so).
// The fp32 SIMT kernel is the parity path.  Its r1 profile: 107 M warp instructions per launch for
// 16 M query-key-head triples, issue bound; every CTA pays the row_info -> q/k/v -> LUT dependent
// load chain before any math.  This kernel removes both:
//   * persistent, one CTA per SM, each owning a 64-channel slice (2 heads of 32 or 4 heads of 16) and
//     walking bins of 64 CSR rows (the windows that START in the bin, <= 127 rows).  The slice of the
//     positional LUT lives in shared memory (bf16) for the whole kernel;
//   * warp specialised, three stage buffers: in iteration j the 8 "stager" warps take bin j (wait for
//     its cp.async rows, build the work units, normalise q and k in place) while the 8 "math" warps
//     run the MMAs of bin j-1 and the rows of bin j+1 / the row_info records of bins j+2, j+3 are in
//     flight; work units are handed out through a shared-memory counter and the stagers join the math
//     once their bin is staged; one CTA barrier per iteration;
//   * one pass over q and k only: + LUT, L2-normalise per head, fold log2(e)/tau into q, back to
//     bf16 in place.  v is used as it arrives;
//   * work units are PACKED: a unit is either a run of whole small windows totalling <= 16 rows, or a
//     16-row chunk of a large window; per-row key bounds (from row_info) give the block-diagonal
//     mask.  A 3-token window therefore costs 3/16 of an MMA tile instead of a whole one;
//   * one warp per (unit, head): S = Q K^T accumulates in registers (ldmatrix operands), the softmax
//     runs on the C fragments with quad shuffles, P goes back into the second MMA as the A operand
//     directly from registers, V comes in through ldmatrix.trans.  Row pitch 144 B makes every
//     ldmatrix phase conflict free.  The unit body is compiled for 16 / 32 / 48 / 64 keys.
// bf16 operands (8-bit mantissa): tests hold this kernel to 1e-2 against the fp32 kernel on the same
// (bf16-rounded) inputs.
#include "common.cuh"
#include <cuda_bf16.h>

#define MM_BIN 64
#define MM_ROWS 144            // 127 rows + 16 rows of MMA overrun + 1
#define MM_INFO 128            // row_info records cached per bin
#define MM_SLICE 64
#define MM_PITCH 72            // bf16 elements per smem row (144 B)
#define MM_THREADS 512
#define MM_GROUP 256           // threads per role
#define MM_STAGES 3
#define MM_NINFO 5
#define MM_UNITS 48
#define MM_ARR (MM_ROWS * MM_PITCH)
#define MM_STAGE_ELEMS (3 * MM_ARR)
#define MM_SMEM_BYTES (64 * 2 * MM_SLICE * 2 + MM_STAGES * MM_STAGE_ELEMS * 2 + MM_NINFO * MM_INFO * 16 + 2 * (MM_UNITS + 16) * 4)

typedef __nv_bfloat16 bf16;

struct MmArgs {
  const bf16* qkv;     // (N, 3d) bf16
  const float* lut;    // (64, 2d)
  const int4* row_info;
  const float* tau;
  const float* bv;     // (d) value bias added to the output, nullable
  float tau_min;
  int N, d;
  int out_bf16;
};

__device__ __forceinline__ void mm_cp16(void* smem_dst, const void* gmem_src) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void mm_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void mm_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }
__device__ __forceinline__ void mm_bar_stagers() { asm volatile("bar.sync 1, %0;" ::"n"(MM_GROUP) : "memory"); }

__device__ __forceinline__ void ldsm_x4(unsigned (&r)[4], const bf16* p) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(sa));
}
__device__ __forceinline__ void ldsm_x4_t(unsigned (&r)[4], const bf16* p) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(sa));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<unsigned*>(&v);
}

// rows [row0, row1) = the windows that start inside the bin; inf = records of rows bin .. bin+127
__device__ __forceinline__ void mm_bin_range(const int4* inf, int bin, int N, int& row0, int& R) {
  int4 f = inf[0];
  row0 = (f.y == bin) ? bin : f.z;
  int row1 = N;
  if (bin + MM_BIN < N) {
    int4 l = inf[MM_BIN];
    row1 = (l.y == bin + MM_BIN) ? bin + MM_BIN : l.z;
  }
  R = row1 - row0;
  if (R < 0) R = 0;
}

// the stager group (MM_GROUP threads, index gt) issues all copies
__device__ __forceinline__ void mm_issue_info(int4* inf, const int4* row_info, int bin, int N, int gt) {
  if (gt < MM_INFO && bin + gt < N) mm_cp16(inf + gt, row_info + bin + gt);
}

__device__ __forceinline__ void mm_issue_rows(bf16* stage, const int4* inf, const bf16* qkv, int d, int col, int bin, int N, int gt) {
  int row0, R;
  mm_bin_range(inf, bin, N, row0, R);
  const int shift = row0 - bin;
  for (int idx = gt; idx < R * 8; idx += MM_GROUP) {      // one 16-byte chunk of q, k and v each
    const int r = idx >> 3, c8 = idx & 7;
    const bf16* src = qkv + (long long)inf[r + shift].x * 3 * d + col + 8 * c8;
    bf16* dst = stage + r * MM_PITCH + 8 * c8;
    mm_cp16(dst, src);
    mm_cp16(dst + MM_ARR, src + d);
    mm_cp16(dst + 2 * MM_ARR, src + 2 * d);
  }
}

// bits [pos, pos+32) of the 128-bit mask (m1:m0)
__device__ __forceinline__ unsigned mm_bits(unsigned long long m0, unsigned long long m1, int pos) {
  unsigned long long v;
  if (pos >= 128) return 0u;
  if (pos >= 64) v = m1 >> (pos - 64);
  else v = (m0 >> pos) | (pos ? (m1 << (64 - pos)) : 0ull);
  return (unsigned)v;
}

// One (unit, head) on one warp: S = Q K^T, masked softmax, O = P V, for NT2 8-key tiles (NT2 even).
template <int HD, int NT2>
__device__ __forceinline__ void mm_unit(const MmArgs& a, const bf16* sq, const bf16* sk, const bf16* sv, int q0, int qn, int k0,
                                        int ch, int4 recA, int4 recB, int kbase, int lane, long long out_col, int lse_col,
                                        void* __restrict__ out, float* __restrict__ lse) {
  constexpr int KS = HD / 16, ND = HD / 8;
  const int g = lane >> 2, t = lane & 3;
  const int loA = recA.y - kbase, wA = recA.z - recA.y, loB = recB.y - kbase, wB = recB.z - recB.y;
  unsigned qa[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
    ldsm_x4(qa[ks], sq + (q0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * MM_PITCH + ch + 16 * ks + 8 * (lane >> 4));
  float c[NT2][4];
#pragma unroll
  for (int nt = 0; nt < NT2; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
#pragma unroll
  for (int np = 0; np < NT2 / 2; ++np) {
    unsigned kb[4];
    if (KS == 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        ldsm_x4(kb, sk + (k0 + 8 * (2 * np + u) + (lane & 7)) * MM_PITCH + ch + 8 * (lane >> 3));
        mma_bf16(c[2 * np + u], qa[0], kb[0], kb[1]);
        mma_bf16(c[2 * np + u], qa[KS - 1], kb[2], kb[3]);
      }
    } else {
      ldsm_x4(kb, sk + (k0 + 16 * np + 8 * (lane >> 4) + (lane & 7)) * MM_PITCH + ch + 8 * ((lane >> 3) & 1));
      mma_bf16(c[2 * np], qa[0], kb[0], kb[1]);
      mma_bf16(c[2 * np + 1], qa[0], kb[2], kb[3]);
    }
  }
  // block-diagonal mask + row max
  float mx0 = -INFINITY, mx1 = -INFINITY;
  const int ka = 2 * t - loA, kb_ = 2 * t - loB;
#pragma unroll
  for (int nt = 0; nt < NT2; ++nt) {
    if ((unsigned)(ka + 8 * nt) >= (unsigned)wA) c[nt][0] = -INFINITY;
    if ((unsigned)(ka + 8 * nt + 1) >= (unsigned)wA) c[nt][1] = -INFINITY;
    if ((unsigned)(kb_ + 8 * nt) >= (unsigned)wB) c[nt][2] = -INFINITY;
    if ((unsigned)(kb_ + 8 * nt + 1) >= (unsigned)wB) c[nt][3] = -INFINITY;
    mx0 = fmaxf(mx0, fmaxf(c[nt][0], c[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(c[nt][2], c[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT2; ++nt) {
    c[nt][0] = fast_exp2(c[nt][0] - mx0); c[nt][1] = fast_exp2(c[nt][1] - mx0);
    c[nt][2] = fast_exp2(c[nt][2] - mx1); c[nt][3] = fast_exp2(c[nt][3] - mx1);
    l0 += c[nt][0] + c[nt][1];
    l1 += c[nt][2] + c[nt][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  float o[ND][4];
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < NT2 / 2; ++kt) {
    const unsigned pa[4] = {pack_bf16(c[2 * kt][0], c[2 * kt][1]), pack_bf16(c[2 * kt][2], c[2 * kt][3]),
                            pack_bf16(c[2 * kt + 1][0], c[2 * kt + 1][1]), pack_bf16(c[2 * kt + 1][2], c[2 * kt + 1][3])};
#pragma unroll
    for (int np = 0; np < ND / 2; ++np) {
      unsigned vb[4];
      ldsm_x4_t(vb, sv + (k0 + 16 * kt + (lane & 7) + 8 * ((lane >> 3) & 1)) * MM_PITCH + ch + 16 * np + 8 * (lane >> 4));
      mma_bf16(o[2 * np], pa, vb[0], vb[1]);
      mma_bf16(o[2 * np + 1], pa, vb[2], vb[3]);
    }
  }
  float2 bias[ND];
#pragma unroll
  for (int nd = 0; nd < ND; ++nd)
    bias[nd] = a.bv ? __ldg(reinterpret_cast<const float2*>(a.bv + out_col + 8 * nd + 2 * t)) : make_float2(0.f, 0.f);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if (g + 8 * half < qn) {
      const int tok = half ? recB.x : recA.x;
      const float il = __fdividef(1.f, half ? l1 : l0);
      const long long e0 = (long long)tok * a.d + out_col + 2 * t;
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        const float x0 = fmaf(o[nd][2 * half], il, bias[nd].x), x1 = fmaf(o[nd][2 * half + 1], il, bias[nd].y);
        if (a.out_bf16) *reinterpret_cast<unsigned*>((bf16*)out + e0 + 8 * nd) = pack_bf16(x0, x1);
        else *reinterpret_cast<float2*>((float*)out + e0 + 8 * nd) = make_float2(x0, x1);
      }
      // natural-log lse of the scores S = cos / tau (the backward kernels expect it)
      if (t == 0) lse[(long long)tok * 8 + lse_col] = ((half ? mx1 : mx0) + __log2f(half ? l1 : l0)) * 0.6931471805599453f;
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(MM_THREADS, 1) sra_fwd_mma_kernel(MmArgs a, void* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* slut = (bf16*)smem_raw;                               // [64][128]: q part | k part of this slice
  bf16* sdata = slut + 64 * 2 * MM_SLICE;                     // MM_STAGES x {q, k, v} x [144][72]
  int4* sinfo_all = (int4*)(sdata + MM_STAGES * MM_STAGE_ELEMS);
  int* sunit_all = (int*)(sinfo_all + MM_NINFO * MM_INFO);    // 2 x { units: q0 | qn << 7 | k0 << 12 | kn << 19 ; [MM_UNITS] = count }
  constexpr int HS = MM_SLICE / HD;
  const int d = a.d;
  const int nsl = d / MM_SLICE;
  const int sl = blockIdx.x % nsl, cta = blockIdx.x / nsl, ncta = gridDim.x / nsl;
  const int col = sl * MM_SLICE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbins = (a.N + MM_BIN - 1) / MM_BIN;
  constexpr int GW = MM_GROUP >> 5;                           // warps per role
  const bool stager = warp < GW;
  const int gt = tid & (MM_GROUP - 1), gw = warp & (GW - 1);
  if (cta >= nbins) return;
  const int my_bins = (nbins - cta + ncta - 1) / ncta;        // bins cta, cta + ncta, ...

  // ---- prologue: zero the data buffers (overrun rows must hold finite values), LUT slice -> bf16, first records
  for (int i = tid; i < MM_STAGES * MM_STAGE_ELEMS / 8; i += MM_THREADS) reinterpret_cast<uint4*>(sdata)[i] = make_uint4(0, 0, 0, 0);
  for (int idx = tid; idx < 64 * MM_SLICE; idx += MM_THREADS) {      // one bf16 pair each
    int pos = idx >> 6, c2 = idx & 63;
    int part = c2 >> 5, cc = (c2 & 31) * 2;
    float2 v = __ldg(reinterpret_cast<const float2*>(a.lut + (long long)pos * 2 * d + part * d + col + cc));
    *reinterpret_cast<unsigned*>(slut + pos * 2 * MM_SLICE + part * MM_SLICE + cc) = pack_bf16(v.x, v.y);
  }
  if (stager) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (j < my_bins) mm_issue_info(sinfo_all + j * MM_INFO, a.row_info, (cta + j * ncta) * MM_BIN, a.N, gt);
    mm_commit();
    mm_wait<0>();
  }
  __syncthreads();
  if (stager) {
    // the copy group committed at the end of iteration j holds the rows of bin j+1 and the records of bin j+3
    mm_issue_rows(sdata, sinfo_all, a.qkv, d, col, cta * MM_BIN, a.N, gt);
    mm_commit();
  }
  const float qscale = 1.4426950408889634f / fmaxf(__ldg(a.tau), a.tau_min);   // log2(e) / tau: softmax in base 2

  for (int j = 0; j <= my_bins; ++j) {
    if (stager) {
      // ================= stagers: bin j
      if (j < my_bins) {
        const int bin = (cta + j * ncta) * MM_BIN;
        const int4* sinfo = sinfo_all + (j % MM_NINFO) * MM_INFO;
        bf16* sq = sdata + (j % MM_STAGES) * MM_STAGE_ELEMS;
        // rows of bin j+1 (needs the records of bin j+1: landed with the previous group) and records of bin j+3
        if (j + 1 < my_bins)
          mm_issue_rows(sdata + ((j + 1) % MM_STAGES) * MM_STAGE_ELEMS, sinfo_all + ((j + 1) % MM_NINFO) * MM_INFO, a.qkv, d, col,
                        (cta + (j + 1) * ncta) * MM_BIN, a.N, gt);
        if (j + 3 < my_bins)
          mm_issue_info(sinfo_all + ((j + 3) % MM_NINFO) * MM_INFO, a.row_info, (cta + (j + 3) * ncta) * MM_BIN, a.N, gt);
        mm_commit();
        mm_wait<1>();          // everything but the group just committed: rows of bin j, records of bin j+2
        mm_bar_stagers();      // ... from every stager thread
        int row0, R;
        mm_bin_range(sinfo, bin, a.N, row0, R);
        const int shift = row0 - bin;
        // ---- q and k in place: + LUT, L2-normalise per head, q also x log2(e)/tau; one task = 8 channels,
        // two tasks per thread in flight
        const int ntask = R * 16;
        for (int base = 0; base < ntask; base += 2 * MM_GROUP) {
          uint4 raw[2], lr[2];
          bf16* p[2];
          bool valid[2];
          int part[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int idx = base + u * MM_GROUP + gt;
            valid[u] = idx < ntask;
            const int r = valid[u] ? (idx >> 4) : 0;
            part[u] = (idx >> 3) & 1;
            const int c8 = idx & 7;
            p[u] = sq + part[u] * MM_ARR + r * MM_PITCH + 8 * c8;
            raw[u] = *reinterpret_cast<const uint4*>(p[u]);
            lr[u] = *reinterpret_cast<const uint4*>(slut + sinfo[r + shift].w * 2 * MM_SLICE + part[u] * MM_SLICE + 8 * c8);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float x[8];
            const unsigned rw[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w}, lw[4] = {lr[u].x, lr[u].y, lr[u].z, lr[u].w};
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              x[2 * i] = __uint_as_float(rw[i] << 16) + __uint_as_float(lw[i] << 16);
              x[2 * i + 1] = __uint_as_float(rw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
              ss = fmaf(x[2 * i], x[2 * i], ss);
              ss = fmaf(x[2 * i + 1], x[2 * i + 1], ss);
            }
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            if (HD == 32) ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            float f;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(fmaxf(ss, 1e-24f)));
            if (part[u] == 0) f *= qscale;
            if (valid[u]) {
              uint4 o;
              o.x = pack_bf16(x[0] * f, x[1] * f);
              o.y = pack_bf16(x[2] * f, x[3] * f);
              o.z = pack_bf16(x[4] * f, x[5] * f);
              o.w = pack_bf16(x[6] * f, x[7] * f);
              *reinterpret_cast<uint4*>(p[u]) = o;
            }
          }
        }
      }
    } else if (gw == 0 && j < my_bins) {
      // ---- work units of bin j, built by one math warp (the records of bin j landed an iteration ago)
      const int bin = (cta + j * ncta) * MM_BIN;
      const int4* sinfo = sinfo_all + (j % MM_NINFO) * MM_INFO;
      int* sunit = sunit_all + (j & 1) * (MM_UNITS + 16);
      int row0, R;
      mm_bin_range(sinfo, bin, a.N, row0, R);
      const int shift = row0 - bin;
      int nu = 0;
      if (R > 0) {
        unsigned long long m0 = 0, m1 = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          int r = 32 * w + lane;
          bool st = r < R && sinfo[min(r, R - 1) + shift].y == row0 + r;
          unsigned long long b = __ballot_sync(0xffffffffu, st);
          if (w < 2) m0 |= b << (32 * w);
          else m1 |= b << (32 * (w - 2));
        }
        if (R < 64) m0 |= 1ull << R;      // sentinel: one past the last row
        else m1 |= 1ull << (R - 64);
        int s = 0;
        while (s < R && nu < MM_UNITS) {
          unsigned w16 = mm_bits(m0, m1, s + 1) & 0xffffu;     // window starts at rows s+1 .. s+16
          if (w16) {                                           // run of whole windows with <= 16 rows in total
            int e = s + 32 - __clz(w16);
            if (lane == 0) sunit[nu] = s | ((e - s) << 7) | (s << 12) | ((e - s) << 19);
            ++nu;
            s = e;
          } else {                                             // a window of more than 16 rows: 16-row query chunks
            unsigned lo = mm_bits(m0, m1, s + 17), hi = mm_bits(m0, m1, s + 49);
            int n = lo ? 16 + __ffs(lo) : 48 + __ffs(hi);
            for (int m = 0; m < n && nu < MM_UNITS; m += 16) {
              if (lane == 0) sunit[nu] = (s + m) | (min(16, n - m) << 7) | (s << 12) | (n << 19);
              ++nu;
            }
            s += n;
          }
        }
      }
      if (lane == 0) { sunit[MM_UNITS] = nu; sunit[MM_UNITS + 1] = 0; }   // count, work counter
    }
    if (j > 0) {
      // ================= bin j-1: the math warps start at once, the stagers join when bin j is staged;
      // (unit, head) entries are handed out through a shared-memory counter
      const int jb = j - 1;
      const int bin = (cta + jb * ncta) * MM_BIN;
      const int4* sinfo = sinfo_all + (jb % MM_NINFO) * MM_INFO;
      const bf16* sq = sdata + (jb % MM_STAGES) * MM_STAGE_ELEMS;
      const bf16* sk = sq + MM_ARR;
      const bf16* sv = sk + MM_ARR;
      int* sunit = sunit_all + (jb & 1) * (MM_UNITS + 16);
      int row0, R;
      mm_bin_range(sinfo, bin, a.N, row0, R);
      const int shift = row0 - bin;
      const int nent = sunit[MM_UNITS] * HS;
      const int g = lane >> 2;
      for (;;) {
        int e = 0;
        if (lane == 0) e = atomicAdd(sunit + MM_UNITS + 1, 1);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= nent) break;
        const int code = sunit[e / HS];
        const int h = e % HS;
        const int q0 = code & 127, qn = (code >> 7) & 31, k0 = (code >> 12) & 127, kn = (code >> 19) & 127;
        const int ch = h * HD;
        const int4 recA = sinfo[min(q0 + g, R - 1) + shift], recB = sinfo[min(q0 + g + 8, R - 1) + shift];
        const int kbase = row0 + k0;
        const long long oc = col + ch;
        const int lc = sl * HS + h;
        if (kn <= 16) mm_unit<HD, 2>(a, sq, sk, sv, q0, qn, k0, ch, recA, recB, kbase, lane, oc, lc, out, lse);
        else if (kn <= 32) mm_unit<HD, 4>(a, sq, sk, sv, q0, qn, k0, ch, recA, recB, kbase, lane, oc, lc, out, lse);
        else if (kn <= 48) mm_unit<HD, 6>(a, sq, sk, sv, q0, qn, k0, ch, recA, recB, kbase, lane, oc, lc, out, lse);
        else mm_unit<HD, 8>(a, sq, sk, sv, q0, qn, k0, ch, recA, recB, kbase, lane, oc, lc, out, lse);
      }
    }
    __syncthreads();   // bin j is staged for the math warps; bin j-1's buffers are free for the copies of bin j+2
  }
  if (stager) mm_wait<0>();
}

static int mm_attrs() {
  static bool done = false;
  if (!done) {
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_BYTES));
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_BYTES));
    done = true;
  }
  return GDMAE_OK;
}

// Tensor-core variant of gdmae_sra_attention_fwd for bf16 q/k/v: qkv (N, 3d) bf16, otherwise the same arguments and outputs.
extern "C" int gdmae_sra_attention_fwd_tc(const void* qkv_bf16, const float* lut, const int32_t* row_info, int64_t N, int d, int nhead,
                                          const float* tau, float tau_min, const float* bv, int io_bf16, void* out, float* lse,
                                          void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 27) && nhead == 8 && (d == 128 || d == 256));
  GDMAE_CHECK_ARG(((uintptr_t)row_info % 16) == 0 && ((uintptr_t)qkv_bf16 % 16) == 0 && ((uintptr_t)lut % 8) == 0);
  GDMAE_CHECK_ARG(bv == nullptr || ((uintptr_t)bv % 8) == 0);
  if (N == 0) return GDMAE_OK;
  int rc = mm_attrs();
  if (rc) return rc;
  MmArgs a{(const bf16*)qkv_bf16, lut, (const int4*)row_info, tau, bv, tau_min, (int)N, d, io_bf16};
  cudaStream_t st = (cudaStream_t)stream_;
  // one CTA per SM; 148 is a multiple of the 2 (d = 128) and 4 (d = 256) channel slices
  if (d == 128) sra_fwd_mma_kernel<16><<<GDMAE_NUM_SMS, MM_THREADS, MM_SMEM_BYTES, st>>>(a, out, lse);
  else sra_fwd_mma_kernel<32><<<GDMAE_NUM_SMS, MM_THREADS, MM_SMEM_BYTES, st>>>(a, out, lse);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
