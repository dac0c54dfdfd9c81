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
//   * warp specialised and decoupled: "stager" warps copy the rows of a bin with cp.async (three stage buffers) and
//     normalise q and k in place, "math" warps run the MMAs; the two roles meet only through mbarriers (bin staged /
//     bin consumed) - no CTA barrier inside the bin loop, a math warp that runs out of work in bin j moves on to bin
//     j+1 on its own.  try_wait carries a suspend hint: a polling wait was measured to burn a third of the issue slots;
//   * the work units of every bin are built ONCE per window table (gdmae_sra_bin_units: a table serves two layers,
//     forward and backward) and arrive with the bin's row records; (unit, head) entries are handed out through a
//     shared-memory counter that arrives zeroed with them;
//   * one pass over q and k only: + LUT, L2-normalise per head, fold log2(e)/tau into q, back to
//     bf16 in place.  v is used as it arrives;
//   * work units are PACKED: a unit is either a run of whole small windows totalling <= 16 rows, or a
//     16-row chunk of a large window; per-row key bounds (from row_info) give the block-diagonal
//     mask.  A 3-token window therefore costs 3/16 of an MMA tile instead of a whole one;
//   * one warp per (unit, head) - per (unit, head pair) for 16-channel heads, the two heads interleaved as independent
//     instruction streams: S = Q K^T accumulates in registers (ldmatrix operands), the softmax
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
#ifndef MM_STAGER_WARPS
#define MM_STAGER_WARPS 6      // forward: warps that copy and stage bins; the other 16 - MM_STAGER_WARPS run the MMAs
#endif
// forward: heads one math warp runs interleaved per work entry.  Measured (tools/sweep_sra_tc.py, r1): two heads win for the
// 16-channel heads of d = 128 (28.7 vs 30.7 us), one head for the 32-channel heads of d = 256 (78 vs 89 us: the two-head body
// sits at the 128-register cap and halves the entries the math warps can share)
#ifdef MM_HPE
#define MM_HEADS_PER_ENTRY(HD) (MM_HPE)
#else
#define MM_HEADS_PER_ENTRY(HD) ((HD) == 16 ? 2 : 1)
#endif
#ifndef MM_STAGE_ILP
#define MM_STAGE_ILP 4         // forward: rows a stager thread keeps in flight
#endif
#ifndef MF_THREADS
#define MF_THREADS 512         // forward CTA size (the register cap follows: 65536 / MF_THREADS)
#endif
#define MM_GROUP (32 * MM_STAGER_WARPS)   // stager threads (forward)
#define MM_MATH_WARPS (MF_THREADS / 32 - MM_STAGER_WARPS)
#define MM_STAGES 3
#define MM_NINFO 5
#define MM_UNITS 48
#define MM_UNIT_STRIDE 64      // ints per bin in the unit table: [0, 48) units, [48] count, [49] work counter (0), pad
#define MM_UNIT_CHUNKS 13      // 16-byte chunks copied per bin (52 ints)
#define MM_ARR (MM_ROWS * MM_PITCH)
#define MM_STAGE_ELEMS (3 * MM_ARR)
#define MM_SMEM_BYTES (64 * 2 * MM_SLICE * 2 + MM_STAGES * MM_STAGE_ELEMS * 2 + MM_NINFO * MM_INFO * 16 + MM_NINFO * MM_UNIT_STRIDE * 4)

typedef __nv_bfloat16 bf16;

struct MmArgs {
  const bf16* qkv;     // (N, 3d) bf16
  const float* lut;    // (64, 2d)
  const int4* row_info;
  const int* bin_units; // (ceil(N/64), 64) work units per bin (gdmae_sra_bin_units)
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

// ---- mbarriers (shared::cta): producer/consumer hand-over of stage buffers without CTA-wide barriers
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// bounded waits that ran out (a logic error or a lost signal) are FATAL: the counter is bumped (gdmae_sra_wait_timeouts
// tells the host why) and the kernel traps, so the launch ends in a sticky CUDA error that the next stream
// synchronisation / loss read raises - a training run can never continue on unstaged buffers.
__device__ unsigned int g_sra_wait_timeouts;
__device__ __noinline__ void sra_wait_timed_out() {
  atomicAdd(&g_sra_wait_timeouts, 1u);
  __threadfence_system();
  __trap();
}

// waits for the completion of the phase with the given parity.  try_wait carries a suspend-time hint: the warp sleeps in
// hardware until the phase completes (a polling loop without it was measured to burn a third of the SM's issue slots and
// starve the producer warps).  Bounded (about a second) so that a logic error ends in a trapped launch, not in a hung GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  for (int spin = 0; spin < 50000; ++spin) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
    if (ok) return;
  }
  sra_wait_timed_out();
}

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
__device__ __forceinline__ void mm_issue_info(int4* inf, int* units, const int4* row_info, const int* bin_units, int bin, int N, int gt) {
  if (gt < MM_INFO) {
    if (bin + gt < N) mm_cp16(inf + gt, row_info + bin + gt);
  } else if (gt < MM_INFO + MM_UNIT_CHUNKS) {
    mm_cp16(units + 4 * (gt - MM_INFO), bin_units + (long long)(bin / MM_BIN) * MM_UNIT_STRIDE + 4 * (gt - MM_INFO));
  }
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

// Staging of U 8-channel groups (U rows, same channels) of q or k, in place: + LUT (packed bf16 add), L2 norm of the
// head in fp32 (the 8-channel groups of a head sit in adjacent lanes: 2 lanes for 16-channel heads, 4 for 32), scale,
// back to bf16.  All loads come first and all stores last so that the U dependency chains overlap.  rn[u] = 1/|x|.
template <int HD, int U>
__device__ __forceinline__ void mm_stage_tasks(bf16* const (&p)[U], const bf16* const (&lut_row)[U], float extra_scale,
                                               const bool (&valid)[U], float (&rn)[U]) {
  uint4 raw[U], lr[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    raw[u] = *reinterpret_cast<const uint4*>(p[u]);
    lr[u] = *reinterpret_cast<const uint4*>(lut_row[u]);
  }
  float x[U][8], ss[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
    const unsigned l[4] = {lr[u].x, lr[u].y, lr[u].z, lr[u].w};
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 sum = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]), *reinterpret_cast<const __nv_bfloat162*>(&l[i]));
      const unsigned v = *reinterpret_cast<unsigned*>(&sum);
      x[u][2 * i] = __uint_as_float(v << 16);
      x[u][2 * i + 1] = __uint_as_float(v & 0xffff0000u);
      s0 = fmaf(x[u][2 * i], x[u][2 * i], s0);
      s1 = fmaf(x[u][2 * i + 1], x[u][2 * i + 1], s1);
    }
    ss[u] = s0 + s1;
  }
#pragma unroll
  for (int u = 0; u < U; ++u) ss[u] += __shfl_xor_sync(0xffffffffu, ss[u], 1);
  if (HD == 32) {
#pragma unroll
    for (int u = 0; u < U; ++u) ss[u] += __shfl_xor_sync(0xffffffffu, ss[u], 2);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(ss[u], 1e-24f)));
    rn[u] = r;
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float f = rn[u] * extra_scale;
    if (valid[u]) {
      uint4 o;
      o.x = pack_bf16(x[u][0] * f, x[u][1] * f);
      o.y = pack_bf16(x[u][2] * f, x[u][3] * f);
      o.z = pack_bf16(x[u][4] * f, x[u][5] * f);
      o.w = pack_bf16(x[u][6] * f, x[u][7] * f);
      *reinterpret_cast<uint4*>(p[u]) = o;
    }
  }
}

// The staging pass of one bin: NT threads (index t) own the 8-channel group t & 7 of array (t >> 3) & 1 (q or k) and walk
// the rows U * NT / 16 at a time.  srq / srk (nullable): 1/|q|, 1/|k| per (row, head) for the backward.
template <int HD, int U, int NT>
__device__ __forceinline__ void mm_stage_bin(bf16* sq, const bf16* slut, const int4* inf, int R, float qscale, int t, float* srq,
                                             float* srk) {
  constexpr int RP = NT / 16;                                 // rows per sub-pass
  const int c8 = t & 7, part = (t >> 3) & 1, rsub = t >> 4;
  const float sc = part == 0 ? qscale : 1.f;
  bf16* base = sq + part * MM_ARR + 8 * c8;
  const bf16* lutc = slut + part * MM_SLICE + 8 * c8;
  float* srn = part == 0 ? srq : srk;
  for (int rb = 0; rb < R; rb += U * RP) {                    // uniform trip count: the tasks shuffle inside the warp
    bf16* p[U];
    const bf16* l[U];
    bool valid[U];
    float rn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + rsub + u * RP;
      valid[u] = r < R;
      const int c = valid[u] ? r : 0;
      p[u] = base + c * MM_PITCH;
      l[u] = lutc + inf[c].w * 2 * MM_SLICE;
    }
    mm_stage_tasks<HD, U>(p, l, sc, valid, rn);
    if (srn != nullptr && (c8 & (HD / 8 - 1)) == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (valid[u]) srn[(rb + rsub + u * RP) * 4 + c8 / (HD / 8)] = rn[u];
    }
  }
}

// One unit and NH adjacent heads on one warp: S = Q K^T, masked softmax, O = P V, for NT2 8-key tiles (NT2 even).  The NH
// heads are independent instruction streams over the same rows and masks; every step is written as a loop over the heads
// so that their dependency chains (ldmatrix -> mma -> shuffle -> exp2 -> mma) overlap in the one warp.
template <int HD, int NT2, int NH>
__device__ __forceinline__ void mm_unit(const MmArgs& a, const bf16* sq, const bf16* sk, const bf16* sv, int q0, int qn, int k0,
                                        int ch, int4 recA, int4 recB, int kbase, int lane, long long out_col, int lse_col,
                                        void* __restrict__ out, float* __restrict__ lse) {
  constexpr int KS = HD / 16, ND = HD / 8;
  const int g = lane >> 2, t = lane & 3;
  const int loA = recA.y - kbase, wA = recA.z - recA.y, loB = recB.y - kbase, wB = recB.z - recB.y;
  unsigned qa[NH][KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int hh = 0; hh < NH; ++hh)
      ldsm_x4(qa[hh][ks], sq + (q0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * MM_PITCH + ch + hh * HD + 16 * ks + 8 * (lane >> 4));
  float c[NH][NT2][4];
#pragma unroll
  for (int hh = 0; hh < NH; ++hh)
#pragma unroll
    for (int nt = 0; nt < NT2; ++nt) c[hh][nt][0] = c[hh][nt][1] = c[hh][nt][2] = c[hh][nt][3] = 0.f;
#pragma unroll
  for (int np = 0; np < NT2 / 2; ++np) {
    if (KS == 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        unsigned kb[NH][4];
#pragma unroll
        for (int hh = 0; hh < NH; ++hh)
          ldsm_x4(kb[hh], sk + (k0 + 8 * (2 * np + u) + (lane & 7)) * MM_PITCH + ch + hh * HD + 8 * (lane >> 3));
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) mma_bf16(c[hh][2 * np + u], qa[hh][0], kb[hh][0], kb[hh][1]);
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) mma_bf16(c[hh][2 * np + u], qa[hh][KS - 1], kb[hh][2], kb[hh][3]);
      }
    } else {
      unsigned kb[NH][4];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh)
        ldsm_x4(kb[hh], sk + (k0 + 16 * np + 8 * (lane >> 4) + (lane & 7)) * MM_PITCH + ch + hh * HD + 8 * ((lane >> 3) & 1));
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        mma_bf16(c[hh][2 * np], qa[hh][0], kb[hh][0], kb[hh][1]);
        mma_bf16(c[hh][2 * np + 1], qa[hh][0], kb[hh][2], kb[hh][3]);
      }
    }
  }
  // block-diagonal mask (shared by the heads) + row max
  float mx0[NH], mx1[NH];
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) mx0[hh] = mx1[hh] = -INFINITY;
  const int ka = 2 * t - loA, kb_ = 2 * t - loB;
#pragma unroll
  for (int nt = 0; nt < NT2; ++nt) {
    const bool m0 = (unsigned)(ka + 8 * nt) >= (unsigned)wA, m1 = (unsigned)(ka + 8 * nt + 1) >= (unsigned)wA;
    const bool m2 = (unsigned)(kb_ + 8 * nt) >= (unsigned)wB, m3 = (unsigned)(kb_ + 8 * nt + 1) >= (unsigned)wB;
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      if (m0) c[hh][nt][0] = -INFINITY;
      if (m1) c[hh][nt][1] = -INFINITY;
      if (m2) c[hh][nt][2] = -INFINITY;
      if (m3) c[hh][nt][3] = -INFINITY;
      mx0[hh] = fmaxf(mx0[hh], fmaxf(c[hh][nt][0], c[hh][nt][1]));
      mx1[hh] = fmaxf(mx1[hh], fmaxf(c[hh][nt][2], c[hh][nt][3]));
    }
  }
#pragma unroll
  for (int o_ = 1; o_ <= 2; o_ <<= 1)
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      mx0[hh] = fmaxf(mx0[hh], __shfl_xor_sync(0xffffffffu, mx0[hh], o_));
      mx1[hh] = fmaxf(mx1[hh], __shfl_xor_sync(0xffffffffu, mx1[hh], o_));
    }
  float l0[NH], l1[NH];
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) l0[hh] = l1[hh] = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT2; ++nt)
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      c[hh][nt][0] = fast_exp2(c[hh][nt][0] - mx0[hh]); c[hh][nt][1] = fast_exp2(c[hh][nt][1] - mx0[hh]);
      c[hh][nt][2] = fast_exp2(c[hh][nt][2] - mx1[hh]); c[hh][nt][3] = fast_exp2(c[hh][nt][3] - mx1[hh]);
      l0[hh] += c[hh][nt][0] + c[hh][nt][1];
      l1[hh] += c[hh][nt][2] + c[hh][nt][3];
    }
#pragma unroll
  for (int o_ = 1; o_ <= 2; o_ <<= 1)
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      l0[hh] += __shfl_xor_sync(0xffffffffu, l0[hh], o_);
      l1[hh] += __shfl_xor_sync(0xffffffffu, l1[hh], o_);
    }
  float o[NH][ND][4];
#pragma unroll
  for (int hh = 0; hh < NH; ++hh)
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) o[hh][nd][0] = o[hh][nd][1] = o[hh][nd][2] = o[hh][nd][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < NT2 / 2; ++kt) {
    unsigned pa[NH][4];
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      pa[hh][0] = pack_bf16(c[hh][2 * kt][0], c[hh][2 * kt][1]);
      pa[hh][1] = pack_bf16(c[hh][2 * kt][2], c[hh][2 * kt][3]);
      pa[hh][2] = pack_bf16(c[hh][2 * kt + 1][0], c[hh][2 * kt + 1][1]);
      pa[hh][3] = pack_bf16(c[hh][2 * kt + 1][2], c[hh][2 * kt + 1][3]);
    }
#pragma unroll
    for (int np = 0; np < ND / 2; ++np) {
      unsigned vb[NH][4];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh)
        ldsm_x4_t(vb[hh], sv + (k0 + 16 * kt + (lane & 7) + 8 * ((lane >> 3) & 1)) * MM_PITCH + ch + hh * HD + 16 * np + 8 * (lane >> 4));
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        mma_bf16(o[hh][2 * np], pa[hh], vb[hh][0], vb[hh][1]);
        mma_bf16(o[hh][2 * np + 1], pa[hh], vb[hh][2], vb[hh][3]);
      }
    }
  }
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    float2 bias[ND];
#pragma unroll
    for (int nd = 0; nd < ND; ++nd)
      bias[nd] = a.bv ? __ldg(reinterpret_cast<const float2*>(a.bv + out_col + hh * HD + 8 * nd + 2 * t)) : make_float2(0.f, 0.f);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (g + 8 * half < qn) {
        const int tok = half ? recB.x : recA.x;
        const float il = __fdividef(1.f, half ? l1[hh] : l0[hh]);
        const long long e0 = (long long)tok * a.d + out_col + hh * HD + 2 * t;
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) {
          const float x0 = fmaf(o[hh][nd][2 * half], il, bias[nd].x), x1 = fmaf(o[hh][nd][2 * half + 1], il, bias[nd].y);
          if (a.out_bf16) *reinterpret_cast<unsigned*>((bf16*)out + e0 + 8 * nd) = pack_bf16(x0, x1);
          else *reinterpret_cast<float2*>((float*)out + e0 + 8 * nd) = make_float2(x0, x1);
        }
        // natural-log lse of the scores S = cos / tau (the backward kernels expect it)
        if (t == 0)
          lse[(long long)tok * 8 + lse_col + hh] = ((half ? mx1[hh] : mx0[hh]) + __log2f(half ? l1[hh] : l0[hh])) * 0.6931471805599453f;
      }
    }
  }
}

#ifdef MM_PROFILE
// development build only (tools/sweep_sra_tc.py -DMM_PROFILE): cycle counters of one stager warp and one math warp per CTA
__device__ unsigned long long g_mm_prof[16];
#define MM_PROF_T(var) const long long var = clock64()
#define MM_PROF_ADD(slot, cycles) do { if (lane == 0) atomicAdd(&g_mm_prof[slot], (unsigned long long)(cycles)); } while (0)
extern "C" int gdmae_sra_prof_read(unsigned long long* out16, int reset) {
  cudaMemcpyFromSymbol(out16, g_mm_prof, sizeof(g_mm_prof));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_mm_prof, z, sizeof(z)); }
  return 0;
}
#else
#define MM_PROF_T(var)
#define MM_PROF_ADD(slot, cycles)
#endif

template <int HD>
__global__ void __launch_bounds__(MF_THREADS, 1) sra_fwd_mma_kernel(MmArgs a, void* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* slut = (bf16*)smem_raw;                               // [64][128]: q part | k part of this slice
  bf16* sdata = slut + 64 * 2 * MM_SLICE;                     // MM_STAGES x {q, k, v} x [144][72]
  int4* sinfo_all = (int4*)(sdata + MM_STAGES * MM_STAGE_ELEMS);
  int* sunit_all = (int*)(sinfo_all + MM_NINFO * MM_INFO);    // MM_NINFO x { units: q0 | qn << 7 | k0 << 12 | kn << 19 ; [MM_UNITS] = count, [+1] = work counter }
  __shared__ unsigned long long s_full[MM_STAGES], s_empty[MM_STAGES];   // bin staged / bin consumed
  static_assert(MM_GROUP >= MM_INFO + MM_UNIT_CHUNKS && MM_MATH_WARPS >= 1, "stager group too small for the record copies");
  constexpr int HS = MM_SLICE / HD;
  const int d = a.d;
  const int nsl = d / MM_SLICE;
  const int sl = blockIdx.x % nsl, cta = blockIdx.x / nsl, ncta = gridDim.x / nsl;
  const int col = sl * MM_SLICE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbins = (a.N + MM_BIN - 1) / MM_BIN;
  const bool stager = warp < MM_STAGER_WARPS;
  const int gt = tid;                                         // index inside the stager group (stagers are warps 0 ..)
  if (cta >= nbins) return;
  const int my_bins = (nbins - cta + ncta - 1) / ncta;        // bins cta, cta + ncta, ...
  MM_PROF_T(k0_);

  // ---- prologue: first records (their latency hides behind the rest), LUT slice -> bf16 (loads batched), zeroed data
  // buffers (overrun rows must hold finite values), barriers
  if (stager) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (j < my_bins)
        mm_issue_info(sinfo_all + j * MM_INFO, sunit_all + j * MM_UNIT_STRIDE, a.row_info, a.bin_units, (cta + j * ncta) * MM_BIN, a.N, gt);
    mm_commit();
  }
  {
    constexpr int NL = (64 * MM_SLICE + MF_THREADS - 1) / MF_THREADS;   // bf16 pairs per thread
    float2 v[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int idx = tid + i * MF_THREADS;
      const int pos = idx >> 6, c2 = idx & 63;
      const int part = c2 >> 5, cc = (c2 & 31) * 2;
      v[i] = idx < 64 * MM_SLICE ? __ldg(reinterpret_cast<const float2*>(a.lut + (long long)pos * 2 * d + part * d + col + cc))
                                 : make_float2(0.f, 0.f);
    }
    for (int i = tid; i < MM_STAGES * MM_STAGE_ELEMS / 8; i += MF_THREADS) reinterpret_cast<uint4*>(sdata)[i] = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int idx = tid + i * MF_THREADS;
      const int pos = idx >> 6, c2 = idx & 63;
      const int part = c2 >> 5, cc = (c2 & 31) * 2;
      if (idx < 64 * MM_SLICE) *reinterpret_cast<unsigned*>(slut + pos * 2 * MM_SLICE + part * MM_SLICE + cc) = pack_bf16(v[i].x, v[i].y);
    }
  }
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < MM_STAGES; ++st) {
      mbar_init(&s_full[st], MM_STAGER_WARPS);   // every stager warp arrives once its share of the bin is staged (one
                                                 // arrival per thread was measured at ~900 cycles per bin on the stagers)
      mbar_init(&s_empty[st], MM_MATH_WARPS);    // every math warp arrives once it has no more work in the bin
    }
  }
  if (stager) mm_wait<0>();
  __syncthreads();     // the only CTA-wide barrier: from here the two roles meet through the mbarriers
  MM_PROF_T(k1_);

  if (stager) {
    // ================= stagers: stage bin j in place and hand it over, then start the copies of bin j+2 (rows) and
    // bin j+4 (records + units) into the buffers the math warps release with bin j-1.  Staging never waits for the math
    // warps; only the copies do.  (cp.async per thread: 1-D TMA bulk copies of the 128-byte row pieces were measured at
    // ~60 cycles of issue per copy from one warp - 11.6 k cycles per bin - and dropped.)
    const float qscale = 1.4426950408889634f / fmaxf(__ldg(a.tau), a.tau_min);   // log2(e) / tau: softmax in base 2
    mm_issue_rows(sdata, sinfo_all, a.qkv, d, col, cta * MM_BIN, a.N, gt);
    mm_commit();                                                                  // group 0: rows of bin 0
    if (1 < my_bins)
      mm_issue_rows(sdata + MM_STAGE_ELEMS, sinfo_all + MM_INFO, a.qkv, d, col, (cta + ncta) * MM_BIN, a.N, gt);
    if (3 < my_bins)
      mm_issue_info(sinfo_all + 3 * MM_INFO, sunit_all + 3 * MM_UNIT_STRIDE, a.row_info, a.bin_units, (cta + 3 * ncta) * MM_BIN, a.N, gt);
    mm_commit();                                                                  // group 1: rows of bin 1, records of bin 3
    for (int j = 0; j < my_bins; ++j) {
      const int bin = (cta + j * ncta) * MM_BIN;
      const int4* sinfo = sinfo_all + (j % MM_NINFO) * MM_INFO;
      bf16* sq = sdata + (j % MM_STAGES) * MM_STAGE_ELEMS;
      MM_PROF_T(t0);
      mm_wait<1>();          // everything but the latest group: rows of bin j, records + units of bin j+2
      mm_bar_stagers();      // ... from every stager thread
      MM_PROF_T(t1);
      int row0, R;
      mm_bin_range(sinfo, bin, a.N, row0, R);
      const int shift = row0 - bin;
#ifndef MM_DEBUG_NO_STAGE
      mm_stage_bin<HD, MM_STAGE_ILP, MM_GROUP>(sq, slut, sinfo + shift, R, qscale, gt, nullptr, nullptr);
#endif
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full[j % MM_STAGES]);     // release: the warp's copies (waited above) and staged rows
      MM_PROF_T(t2);
      if (j + 2 < my_bins) {
        // the buffer of bin j+2 held bin j-1 (and the record slot bin j+4 takes is bin j-1's): math must be done with it
        if (j + 2 >= MM_STAGES) mbar_wait(&s_empty[(j + 2) % MM_STAGES], (((j + 2) / MM_STAGES) - 1) & 1);
      }
      MM_PROF_T(t3);
      if (j + 2 < my_bins)
        mm_issue_rows(sdata + ((j + 2) % MM_STAGES) * MM_STAGE_ELEMS, sinfo_all + ((j + 2) % MM_NINFO) * MM_INFO, a.qkv, d, col,
                      (cta + (j + 2) * ncta) * MM_BIN, a.N, gt);
      if (j + 4 < my_bins)
        mm_issue_info(sinfo_all + ((j + 4) % MM_NINFO) * MM_INFO, sunit_all + ((j + 4) % MM_NINFO) * MM_UNIT_STRIDE, a.row_info,
                      a.bin_units, (cta + (j + 4) * ncta) * MM_BIN, a.N, gt);
      mm_commit();
#ifdef MM_PROFILE
      if (warp == 0) {
        MM_PROF_T(t4);
        MM_PROF_ADD(0, t3 - t2); MM_PROF_ADD(1, t4 - t3); MM_PROF_ADD(2, t1 - t0); MM_PROF_ADD(4, t2 - t1); MM_PROF_ADD(5, 1);
      }
#endif
    }
    mm_wait<0>();
  } else {
    // ================= math warps: entries of bin j (a unit and one or two heads), handed out through the bin's
    // work counter; a warp that finds the bin exhausted signals and moves on to bin j+1 without waiting for the others
    const int g = lane >> 2;
    for (int j = 0; j < my_bins; ++j) {
      const int bin = (cta + j * ncta) * MM_BIN;
      const int4* sinfo = sinfo_all + (j % MM_NINFO) * MM_INFO;
      const bf16* sq = sdata + (j % MM_STAGES) * MM_STAGE_ELEMS;
      const bf16* sk = sq + MM_ARR;
      const bf16* sv = sk + MM_ARR;
      int* sunit = sunit_all + (j % MM_NINFO) * MM_UNIT_STRIDE;
      MM_PROF_T(m0);
      mbar_wait(&s_full[j % MM_STAGES], (j / MM_STAGES) & 1);
      MM_PROF_T(m1);
#ifdef MM_PROFILE
      int n_done = 0;
#endif
      int row0, R;
      mm_bin_range(sinfo, bin, a.N, row0, R);
      const int shift = row0 - bin;
      constexpr int HPE = MM_HEADS_PER_ENTRY(HD);
      constexpr int EPU = HS / HPE;                           // entries per unit
#ifdef MM_DEBUG_NO_MATH
      const int nent = 0;
#else
      const int nent = sunit[MM_UNITS] * EPU;
#endif
      for (;;) {
        int e = 0;
        if (lane == 0) e = atomicAdd(sunit + MM_UNITS + 1, 1);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= nent) break;
#ifdef MM_PROFILE
        ++n_done;
#endif
        const int code = sunit[e / EPU];
        const int h = (e % EPU) * HPE;
        const int q0 = code & 127, qn = (code >> 7) & 31, k0 = (code >> 12) & 127, kn = (code >> 19) & 127;
        const int ch = h * HD;
        const int4 recA = sinfo[min(q0 + g, R - 1) + shift], recB = sinfo[min(q0 + g + 8, R - 1) + shift];
        const int kbase = row0 + k0;
        const long long oc = col + ch;
        const int lc = sl * HS + h;
        if (kn <= 16) mm_unit<HD, 2, HPE>(a, sq, sk, sv, q0, qn, k0, ch, recA, recB, kbase, lane, oc, lc, out, lse);
        else if (kn <= 32) mm_unit<HD, 4, HPE>(a, sq, sk, sv, q0, qn, k0, ch, recA, recB, kbase, lane, oc, lc, out, lse);
        else {
          // large windows: one head at a time (the two-head body would spill at the register cap)
#pragma unroll 1
          for (int hh = 0; hh < HPE; ++hh) {
            if (kn <= 48) mm_unit<HD, 6, 1>(a, sq, sk, sv, q0, qn, k0, ch + hh * HD, recA, recB, kbase, lane, oc + hh * HD, lc + hh, out, lse);
            else mm_unit<HD, 8, 1>(a, sq, sk, sv, q0, qn, k0, ch + hh * HD, recA, recB, kbase, lane, oc + hh * HD, lc + hh, out, lse);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[j % MM_STAGES]);
#ifdef MM_PROFILE
      if (warp == MM_STAGER_WARPS || warp == MF_THREADS / 32 - 1) {
        MM_PROF_T(m2);
        const int o = warp == MM_STAGER_WARPS ? 6 : 10;
        MM_PROF_ADD(o, m1 - m0); MM_PROF_ADD(o + 1, m2 - m1); MM_PROF_ADD(o + 2, n_done);
        if (o == 6) MM_PROF_ADD(9, nent);
      }
#endif
    }
#ifdef MM_PROFILE
    if (warp == MF_THREADS / 32 - 1 && lane == 0) {
      const long long k2_ = clock64();
      atomicAdd(&g_mm_prof[13], (unsigned long long)(k2_ - k0_));
      atomicMax(&g_mm_prof[14], (unsigned long long)(k2_ - k0_));
      atomicAdd(&g_mm_prof[15], (unsigned long long)(k1_ - k0_));
    }
#endif
  }
}

// =====================================================================================================
// Backward.  Same bins, packing, operand staging and stager / math decoupling as the forward kernel; four staged arrays
// (Qs = q_hat * log2(e)/tau, K_hat, V, dO) in two stage buffers, scalars (lse, 1/|q|, 1/|k|, D) per stage:
//   stage    : + LUT, normalise q and k in place, keep 1/|q|, 1/|k| per (row, head); lse rows by cp.async
//   entries  : query side of every (unit, head), then key side, from one work queue; a key-side entry waits on a
//              shared-memory counter for the query-side entries of its window (they produce D), not on a barrier
//   phase 1  : one warp per (unit, head), query side.  One sweep over the key tiles computes S' = Qs K^T and
//              dP = dO V^T on the tensor cores, P = exp2(S' - lse), and the row sums D = sum P dP,
//              T1 = sum P dP S', T2 = sum P S' (d tau needs sum dS S = T1 - D T2); P stays in registers as
//              packed bf16, dP as fp32; then dS = P (dP - D) feeds G = dS K as the A operand and the
//              normalisation backward dq = (G - Qs (Qs.G) / qscale^2) / (tau |q|) is applied on the fragments.
//   phase 2  : key side, the same units with the roles swapped (attention inside a window is all-to-all, so the
//              queries of a key tile are exactly the unit's key range): per 16-query step S'^T = K Qs^T,
//              dP^T = V dO^T, P^T, dS^T = P^T (dP^T - D), then dV += P^T dO and H += dS^T Qs; nothing is held
//              across steps.  dk = ln2 (H - K (K.H)) / |k|.
// dq, dk, dv rows go straight from the fragments to dqkv (bf16); sum dS S is reduced per CTA into dtau_sum.
#define MB_STAGES 2
#define MB_NINFO 4
#define MB_STAGE_ELEMS (4 * MM_ARR)
#define MB_SCAL (MM_ROWS * 4)      // one fp32 per (row, head of the slice)
#ifndef MB_STAGER_WARPS
#define MB_STAGER_WARPS 5      // backward: warps that copy and stage bins (>= 5: 141 threads copy records + units)
#endif
#define MB_GROUP (32 * MB_STAGER_WARPS)
#define MB_MATH_WARPS (MM_THREADS / 32 - MB_STAGER_WARPS)
// LUT | stages x {q,k,v,dO} | records | per stage: lse, 1/|q|, 1/|k|, D | units | per stage: query-side completion counters
#define MB_SMEM_BYTES (64 * 2 * MM_SLICE * 2 + MB_STAGES * MB_STAGE_ELEMS * 2 + MB_NINFO * MM_INFO * 16 + MB_STAGES * 4 * MB_SCAL * 4 + MB_NINFO * MM_UNIT_STRIDE * 4 + MB_STAGES * 128 * 4 * 4)

struct MbArgs {
  const bf16* qkv;     // (N, 3d) bf16
  const float* lut;    // (64, 2d)
  const int4* row_info;
  const int* bin_units; // (ceil(N/64), 64) work units per bin (gdmae_sra_bin_units)
  const float* tau;
  const float* lse;    // (N, 8)
  const bf16* dout;    // (N, d) bf16
  bf16* dqkv;          // (N, 3d) bf16
  double* dtau_sum;
  float tau_min;
  int N, d;
};

template <int NT>
__device__ __forceinline__ void mb_issue_rows(bf16* stage, float* slse, const int4* inf, const MbArgs& a, int col, int hs, int lse_col,
                                              int bin, int tid) {
  int row0, R;
  mm_bin_range(inf, bin, a.N, row0, R);
  const int shift = row0 - bin;
  const int d = a.d;
  for (int idx = tid; idx < R * 8; idx += NT) {              // one 16-byte chunk of q, k, v and dO each
    const int r = idx >> 3, c8 = idx & 7;
    const long long tok = inf[r + shift].x;
    const bf16* src = a.qkv + tok * 3 * d + col + 8 * c8;
    bf16* dst = stage + r * MM_PITCH + 8 * c8;
    mm_cp16(dst, src);
    mm_cp16(dst + MM_ARR, src + d);
    mm_cp16(dst + 2 * MM_ARR, src + 2 * d);
    mm_cp16(dst + 3 * MM_ARR, a.dout + tok * d + col + 8 * c8);
  }
  for (int idx = tid; idx < R * hs; idx += NT) {             // lse of the slice's heads, 4 bytes each
    const int r = idx / hs, h = idx - r * hs;
    unsigned sa = (unsigned)__cvta_generic_to_shared(slse + r * 4 + h);
    const float* src = a.lse + (long long)inf[r + shift].x * 8 + lse_col + h;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(src) : "memory");
  }
}

__device__ __forceinline__ float2 unpack_bf16(unsigned u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }

// query side of one (unit, head): writes dq rows and sD, returns this lane's share of sum dS*S' (valid rows only)
template <int HD, int NT2>
__device__ __forceinline__ float mb_unit_q(const MbArgs& a, const bf16* sq, const bf16* sk, const bf16* sv, const bf16* sdo,
                                           const float* slse, const float* srq, float* sD, int q0, int qn, int k0, int h, int4 recA,
                                           int4 recB, int kbase, int lane, int col, float inv_tau, float inv_qs2) {
  constexpr int KS = HD / 16, ND = HD / 8;
  const int g = lane >> 2, t = lane & 3, ch = h * HD;
  const int loA = recA.y - kbase, wA = recA.z - recA.y, loB = recB.y - kbase, wB = recB.z - recB.y;
  const int ka = 2 * t - loA, kb_ = 2 * t - loB;
  unsigned qa[KS][4], da[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int off = (q0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * MM_PITCH + ch + 16 * ks + 8 * (lane >> 4);
    ldsm_x4(qa[ks], sq + off);
    ldsm_x4(da[ks], sdo + off);
  }
  const float lA = slse[min(q0 + g, MM_ROWS - 1) * 4 + h] * 1.4426950408889634f;
  const float lB = slse[min(q0 + g + 8, MM_ROWS - 1) * 4 + h] * 1.4426950408889634f;
  unsigned pp[NT2][2];   // P packed bf16: [nt][0] = row g, [nt][1] = row g+8
  float dp[NT2][4];
  float D0 = 0.f, D1 = 0.f, T10 = 0.f, T11 = 0.f, T20 = 0.f, T21 = 0.f;
#pragma unroll
  for (int np = 0; np < NT2 / 2; ++np) {
    float s[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      s[u][0] = s[u][1] = s[u][2] = s[u][3] = 0.f;
      dp[2 * np + u][0] = dp[2 * np + u][1] = dp[2 * np + u][2] = dp[2 * np + u][3] = 0.f;
    }
    unsigned kb[4], vb[4];
    if (KS == 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int off = (k0 + 8 * (2 * np + u) + (lane & 7)) * MM_PITCH + ch + 8 * (lane >> 3);
        ldsm_x4(kb, sk + off);
        ldsm_x4(vb, sv + off);
        mma_bf16(s[u], qa[0], kb[0], kb[1]);
        mma_bf16(s[u], qa[KS - 1], kb[2], kb[3]);
        mma_bf16(dp[2 * np + u], da[0], vb[0], vb[1]);
        mma_bf16(dp[2 * np + u], da[KS - 1], vb[2], vb[3]);
      }
    } else {
      const int off = (k0 + 16 * np + 8 * (lane >> 4) + (lane & 7)) * MM_PITCH + ch + 8 * ((lane >> 3) & 1);
      ldsm_x4(kb, sk + off);
      ldsm_x4(vb, sv + off);
      mma_bf16(s[0], qa[0], kb[0], kb[1]);
      mma_bf16(s[1], qa[0], kb[2], kb[3]);
      mma_bf16(dp[2 * np], da[0], vb[0], vb[1]);
      mma_bf16(dp[2 * np + 1], da[0], vb[2], vb[3]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int nt = 2 * np + u;
      float p[4];
      p[0] = (unsigned)(ka + 8 * nt) < (unsigned)wA ? fast_exp2(s[u][0] - lA) : 0.f;
      p[1] = (unsigned)(ka + 8 * nt + 1) < (unsigned)wA ? fast_exp2(s[u][1] - lA) : 0.f;
      p[2] = (unsigned)(kb_ + 8 * nt) < (unsigned)wB ? fast_exp2(s[u][2] - lB) : 0.f;
      p[3] = (unsigned)(kb_ + 8 * nt + 1) < (unsigned)wB ? fast_exp2(s[u][3] - lB) : 0.f;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float x0 = p[e] * dp[nt][e], x1 = p[2 + e] * dp[nt][2 + e];
        D0 += x0; D1 += x1;
        T10 = fmaf(x0, s[u][e], T10); T11 = fmaf(x1, s[u][2 + e], T11);
        T20 = fmaf(p[e], s[u][e], T20); T21 = fmaf(p[2 + e], s[u][2 + e], T21);
      }
      pp[nt][0] = pack_bf16(p[0], p[1]);
      pp[nt][1] = pack_bf16(p[2], p[3]);
    }
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    D0 += __shfl_xor_sync(0xffffffffu, D0, o); D1 += __shfl_xor_sync(0xffffffffu, D1, o);
    T10 += __shfl_xor_sync(0xffffffffu, T10, o); T11 += __shfl_xor_sync(0xffffffffu, T11, o);
    T20 += __shfl_xor_sync(0xffffffffu, T20, o); T21 += __shfl_xor_sync(0xffffffffu, T21, o);
  }
  float dtau_part = 0.f;   // the quad holds identical sums: count each row once
  if (t == 0) {
    if (g < qn) { dtau_part += T10 - D0 * T20; sD[(q0 + g) * 4 + h] = D0; }
    if (g + 8 < qn) { dtau_part += T11 - D1 * T21; sD[(q0 + g + 8) * 4 + h] = D1; }
  }
  // G = dS K_hat
  float acc[ND][4];
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < NT2 / 2; ++kt) {
    unsigned sa[4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int nt = 2 * kt + u;
      const float2 pA = unpack_bf16(pp[nt][0]), pB = unpack_bf16(pp[nt][1]);
      sa[2 * u] = pack_bf16(pA.x * (dp[nt][0] - D0), pA.y * (dp[nt][1] - D0));
      sa[2 * u + 1] = pack_bf16(pB.x * (dp[nt][2] - D1), pB.y * (dp[nt][3] - D1));
    }
#pragma unroll
    for (int np = 0; np < ND / 2; ++np) {
      unsigned kb[4];
      ldsm_x4_t(kb, sk + (k0 + 16 * kt + (lane & 7) + 8 * ((lane >> 3) & 1)) * MM_PITCH + ch + 16 * np + 8 * (lane >> 4));
      mma_bf16(acc[2 * np], sa, kb[0], kb[1]);
      mma_bf16(acc[2 * np + 1], sa, kb[2], kb[3]);
    }
  }
  // normalisation backward on the fragments: dq = (G - Qs (Qs.G) / qscale^2) / (tau |q|)
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = q0 + g + 8 * half;
    float2 qv[ND];
    float dot = 0.f;
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      qv[nd] = unpack_bf16(*reinterpret_cast<const unsigned*>(sq + row * MM_PITCH + ch + 8 * nd + 2 * t));
      dot = fmaf(qv[nd].x, acc[nd][2 * half], dot);
      dot = fmaf(qv[nd].y, acc[nd][2 * half + 1], dot);
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    if (g + 8 * half < qn) {
      const float f = srq[row * 4 + h] * inv_tau, c2 = dot * inv_qs2;
      bf16* dst = a.dqkv + (long long)(half ? recB.x : recA.x) * 3 * a.d + col + ch + 2 * t;
#pragma unroll
      for (int nd = 0; nd < ND; ++nd)
        *reinterpret_cast<unsigned*>(dst + 8 * nd) =
            pack_bf16(f * (acc[nd][2 * half] - qv[nd].x * c2), f * (acc[nd][2 * half + 1] - qv[nd].y * c2));
    }
  }
  return dtau_part;
}

// key side of one (unit, head): rows = the unit's 16-row tile as KEYS, columns = its key range as QUERIES
template <int HD, int NT2>
__device__ __forceinline__ void mb_unit_kv(const MbArgs& a, const bf16* sq, const bf16* sk, const bf16* sv, const bf16* sdo,
                                           const float* slse, const float* srk, const float* sD, int q0, int qn, int k0, int h,
                                           int4 recA, int4 recB, int kbase, int lane, int col) {
  constexpr int KS = HD / 16, ND = HD / 8;
  const int g = lane >> 2, t = lane & 3, ch = h * HD;
  const int loA = recA.y - kbase, wA = recA.z - recA.y, loB = recB.y - kbase, wB = recB.z - recB.y;
  const int ka = 2 * t - loA, kb_ = 2 * t - loB;
  unsigned ka_f[KS][4], va_f[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int off = (q0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * MM_PITCH + ch + 16 * ks + 8 * (lane >> 4);
    ldsm_x4(ka_f[ks], sk + off);
    ldsm_x4(va_f[ks], sv + off);
  }
  float dv[ND][4], hk[ND][4];
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) {
    dv[nd][0] = dv[nd][1] = dv[nd][2] = dv[nd][3] = 0.f;
    hk[nd][0] = hk[nd][1] = hk[nd][2] = hk[nd][3] = 0.f;
  }
#pragma unroll
  for (int np = 0; np < NT2 / 2; ++np) {
    float s[2][4], dp[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      s[u][0] = s[u][1] = s[u][2] = s[u][3] = 0.f;
      dp[u][0] = dp[u][1] = dp[u][2] = dp[u][3] = 0.f;
    }
    unsigned qb[4], ob[4];
    if (KS == 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int off = (k0 + 8 * (2 * np + u) + (lane & 7)) * MM_PITCH + ch + 8 * (lane >> 3);
        ldsm_x4(qb, sq + off);
        ldsm_x4(ob, sdo + off);
        mma_bf16(s[u], ka_f[0], qb[0], qb[1]);
        mma_bf16(s[u], ka_f[KS - 1], qb[2], qb[3]);
        mma_bf16(dp[u], va_f[0], ob[0], ob[1]);
        mma_bf16(dp[u], va_f[KS - 1], ob[2], ob[3]);
      }
    } else {
      const int off = (k0 + 16 * np + 8 * (lane >> 4) + (lane & 7)) * MM_PITCH + ch + 8 * ((lane >> 3) & 1);
      ldsm_x4(qb, sq + off);
      ldsm_x4(ob, sdo + off);
      mma_bf16(s[0], ka_f[0], qb[0], qb[1]);
      mma_bf16(s[1], ka_f[0], qb[2], qb[3]);
      mma_bf16(dp[0], va_f[0], ob[0], ob[1]);
      mma_bf16(dp[1], va_f[0], ob[2], ob[3]);
    }
    unsigned pa[4], sa[4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int nt = 2 * np + u;
      // this lane's two query columns of the tile
      const int qc = min(k0 + 8 * nt + 2 * t, MM_ROWS - 2);
      const float l0 = slse[qc * 4 + h] * 1.4426950408889634f, l1 = slse[(qc + 1) * 4 + h] * 1.4426950408889634f;
      const float d0 = sD[qc * 4 + h], d1 = sD[(qc + 1) * 4 + h];
      const float p0 = (unsigned)(ka + 8 * nt) < (unsigned)wA ? fast_exp2(s[u][0] - l0) : 0.f;
      const float p1 = (unsigned)(ka + 8 * nt + 1) < (unsigned)wA ? fast_exp2(s[u][1] - l1) : 0.f;
      const float p2 = (unsigned)(kb_ + 8 * nt) < (unsigned)wB ? fast_exp2(s[u][2] - l0) : 0.f;
      const float p3 = (unsigned)(kb_ + 8 * nt + 1) < (unsigned)wB ? fast_exp2(s[u][3] - l1) : 0.f;
      pa[2 * u] = pack_bf16(p0, p1);
      pa[2 * u + 1] = pack_bf16(p2, p3);
      sa[2 * u] = pack_bf16(p0 * (dp[u][0] - d0), p1 * (dp[u][1] - d1));
      sa[2 * u + 1] = pack_bf16(p2 * (dp[u][2] - d0), p3 * (dp[u][3] - d1));
    }
#pragma unroll
    for (int nq = 0; nq < ND / 2; ++nq) {
      unsigned ob2[4], qb2[4];
      const int off = (k0 + 16 * np + (lane & 7) + 8 * ((lane >> 3) & 1)) * MM_PITCH + ch + 16 * nq + 8 * (lane >> 4);
      ldsm_x4_t(ob2, sdo + off);
      ldsm_x4_t(qb2, sq + off);
      mma_bf16(dv[2 * nq], pa, ob2[0], ob2[1]);
      mma_bf16(dv[2 * nq + 1], pa, ob2[2], ob2[3]);
      mma_bf16(hk[2 * nq], sa, qb2[0], qb2[1]);
      mma_bf16(hk[2 * nq + 1], sa, qb2[2], qb2[3]);
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = q0 + g + 8 * half;
    float2 kv[ND];
    float dot = 0.f;
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      kv[nd] = unpack_bf16(*reinterpret_cast<const unsigned*>(sk + row * MM_PITCH + ch + 8 * nd + 2 * t));
      dot = fmaf(kv[nd].x, hk[nd][2 * half], dot);
      dot = fmaf(kv[nd].y, hk[nd][2 * half + 1], dot);
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    if (g + 8 * half < qn) {
      const float f = srk[row * 4 + h] * 0.6931471805599453f;
      bf16* dst = a.dqkv + (long long)(half ? recB.x : recA.x) * 3 * a.d + a.d + col + ch + 2 * t;
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        *reinterpret_cast<unsigned*>(dst + 8 * nd) =
            pack_bf16(f * (hk[nd][2 * half] - kv[nd].x * dot), f * (hk[nd][2 * half + 1] - kv[nd].y * dot));
        *reinterpret_cast<unsigned*>(dst + a.d + 8 * nd) = pack_bf16(dv[nd][2 * half], dv[nd][2 * half + 1]);
      }
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(MM_THREADS, 1) sra_bwd_mma_kernel(MbArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* slut = (bf16*)smem_raw;
  bf16* sdata = slut + 64 * 2 * MM_SLICE;                     // MB_STAGES x {q, k, v, dO} x [144][72]
  int4* sinfo_all = (int4*)(sdata + MB_STAGES * MB_STAGE_ELEMS);
  float* sscal_all = (float*)(sinfo_all + MB_NINFO * MM_INFO); // MB_STAGES x { lse, 1/|q|, 1/|k|, D } x [144][4]
  int* sunit_all = (int*)(sscal_all + MB_STAGES * 4 * MB_SCAL);
  int* sdone_all = sunit_all + MB_NINFO * MM_UNIT_STRIDE;     // MB_STAGES x [128 window start rows][4 heads]: query-side entries finished
  __shared__ float s_dtau[MM_THREADS / 32];
  __shared__ unsigned long long b_full[MB_STAGES], b_empty[MB_STAGES];   // bin staged / bin consumed
  static_assert(MB_GROUP >= MM_INFO + MM_UNIT_CHUNKS && MB_MATH_WARPS >= 1, "stager group too small for the record copies");
  constexpr int HS = MM_SLICE / HD;
  const int d = a.d;
  const int nsl = d / MM_SLICE;
  const int sl = blockIdx.x % nsl, cta = blockIdx.x / nsl, ncta = gridDim.x / nsl;
  const int col = sl * MM_SLICE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbins = (a.N + MM_BIN - 1) / MM_BIN;
  if (cta >= nbins) return;
  const int my_bins = (nbins - cta + ncta - 1) / ncta;
  const int lse_col = sl * HS;
  const bool stager = warp < MB_STAGER_WARPS;

  // ---- prologue: first records, LUT slice -> bf16 (loads batched), zeroed buffers (overrun rows and scalars must be finite)
  if (stager) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (j < my_bins)
        mm_issue_info(sinfo_all + j * MM_INFO, sunit_all + j * MM_UNIT_STRIDE, a.row_info, a.bin_units, (cta + j * ncta) * MM_BIN, a.N, tid);
    mm_commit();
  }
  {
    constexpr int NL = (64 * MM_SLICE + MM_THREADS - 1) / MM_THREADS;   // bf16 pairs per thread
    float2 v[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int idx = tid + i * MM_THREADS;
      const int pos = idx >> 6, c2 = idx & 63;
      const int part = c2 >> 5, cc = (c2 & 31) * 2;
      v[i] = idx < 64 * MM_SLICE ? __ldg(reinterpret_cast<const float2*>(a.lut + (long long)pos * 2 * d + part * d + col + cc))
                                 : make_float2(0.f, 0.f);
    }
    for (int i = tid; i < MB_STAGES * MB_STAGE_ELEMS / 8; i += MM_THREADS) reinterpret_cast<uint4*>(sdata)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < MB_STAGES * 4 * MB_SCAL; i += MM_THREADS) sscal_all[i] = 0.f;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int idx = tid + i * MM_THREADS;
      const int pos = idx >> 6, c2 = idx & 63;
      const int part = c2 >> 5, cc = (c2 & 31) * 2;
      if (idx < 64 * MM_SLICE) *reinterpret_cast<unsigned*>(slut + pos * 2 * MM_SLICE + part * MM_SLICE + cc) = pack_bf16(v[i].x, v[i].y);
    }
  }
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < MB_STAGES; ++st) {
      mbar_init(&b_full[st], MB_STAGER_WARPS);
      mbar_init(&b_empty[st], MB_MATH_WARPS);
    }
  }
  if (stager) mm_wait<0>();
  __syncthreads();     // from here the two roles meet through the mbarriers (and once more for the d tau sum)
  const float tau_c = fmaxf(__ldg(a.tau), a.tau_min);
  const float qscale = 1.4426950408889634f / tau_c, inv_tau = 1.f / tau_c, inv_qs2 = 1.f / (qscale * qscale);
  float dtau_acc = 0.f;

  if (stager) {
    // ================= stagers: stage bin k in place (+ LUT, normalise, keep 1/|q|, 1/|k| per (row, head)), hand it over, then
    // copy bin k+1 (rows, lse) and the records + units of bin k+3 into the buffers the math warps release with bin k-1
    mb_issue_rows<MB_GROUP>(sdata, sscal_all, sinfo_all, a, col, HS, lse_col, cta * MM_BIN, tid);
    mm_commit();
    for (int k = 0; k < my_bins; ++k) {
      const int bin = (cta + k * ncta) * MM_BIN;
      const int st = k % MB_STAGES;
      const int4* sinfo = sinfo_all + (k % MB_NINFO) * MM_INFO;
      bf16* sq = sdata + st * MB_STAGE_ELEMS;
      float* sscal = sscal_all + st * 4 * MB_SCAL;
      mm_wait<0>();        // rows of bin k, records + units of bin k+2
      asm volatile("bar.sync 1, %0;" ::"n"(MB_GROUP) : "memory");   // ... from every stager thread
      int row0, R;
      mm_bin_range(sinfo, bin, a.N, row0, R);
      const int shift = row0 - bin;
      for (int i = tid; i < 128 * 4; i += MB_GROUP) sdone_all[st * 512 + i] = 0;
      mm_stage_bin<HD, 4, MB_GROUP>(sq, slut, sinfo + shift, R, qscale, tid, sscal + MB_SCAL, sscal + 2 * MB_SCAL);
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[st]);
      if (k + 1 < my_bins) {
        const int sn = (k + 1) % MB_STAGES;
        if (k + 1 >= MB_STAGES) mbar_wait(&b_empty[sn], (((k + 1) / MB_STAGES) - 1) & 1);   // math is done with bin k-1
        mb_issue_rows<MB_GROUP>(sdata + sn * MB_STAGE_ELEMS, sscal_all + sn * 4 * MB_SCAL, sinfo_all + ((k + 1) % MB_NINFO) * MM_INFO, a, col, HS,
                                lse_col, (cta + (k + 1) * ncta) * MM_BIN, tid);
      }
      if (k + 3 < my_bins)
        mm_issue_info(sinfo_all + ((k + 3) % MB_NINFO) * MM_INFO, sunit_all + ((k + 3) % MB_NINFO) * MM_UNIT_STRIDE, a.row_info, a.bin_units,
                      (cta + (k + 3) * ncta) * MM_BIN, a.N, tid);
      mm_commit();
    }
    mm_wait<0>();
  } else {
    // ================= math warps.  Entries [0, nent): query side; [nent, 2 nent): key side, handed out in this order through the
    // work counter that arrived (zero) with the unit list.  A key-side entry needs D of every query row of its window: it waits
    // until the query-side entries covering the window (one for a packed unit, ceil(rows/16) chunks for a large window) have
    // signalled.  Every query-side entry is pulled before any key-side entry and never blocks, so the wait cannot deadlock.  A
    // warp that finds the bin exhausted signals and moves on to bin k+1 without waiting for the others.
    const int g = lane >> 2;
    for (int k = 0; k < my_bins; ++k) {
      const int bin = (cta + k * ncta) * MM_BIN;
      const int st = k % MB_STAGES;
      const int4* sinfo = sinfo_all + (k % MB_NINFO) * MM_INFO;
      const bf16* sq = sdata + st * MB_STAGE_ELEMS;
      const bf16* sk = sq + MM_ARR;
      const bf16* sv = sk + MM_ARR;
      const bf16* sdo = sv + MM_ARR;
      float* slse = sscal_all + st * 4 * MB_SCAL;
      float* srq = slse + MB_SCAL;
      float* srk = srq + MB_SCAL;
      float* sD = srk + MB_SCAL;
      int* sdone = sdone_all + st * 512;
      int* sunit = sunit_all + (k % MB_NINFO) * MM_UNIT_STRIDE;
      mbar_wait(&b_full[st], (k / MB_STAGES) & 1);
      int row0, R;
      mm_bin_range(sinfo, bin, a.N, row0, R);
      const int shift = row0 - bin;
      const int nent = sunit[MM_UNITS] * HS;
      for (;;) {
        int e = 0;
        if (lane == 0) e = atomicAdd(sunit + MM_UNITS + 1, 1);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= 2 * nent) break;
        const bool key_side = e >= nent;
        if (key_side) e -= nent;
        const int code = sunit[e / HS];
        const int h = e % HS;
        const int q0 = code & 127, qn = (code >> 7) & 31, k0 = (code >> 12) & 127, kn = (code >> 19) & 127;
        const int4 recA = sinfo[min(q0 + g, R - 1) + shift], recB = sinfo[min(q0 + g + 8, R - 1) + shift];
        const int kbase = row0 + k0;
        if (!key_side) {
          if (kn <= 16) dtau_acc += mb_unit_q<HD, 2>(a, sq, sk, sv, sdo, slse, srq, sD, q0, qn, k0, h, recA, recB, kbase, lane, col, inv_tau, inv_qs2);
          else if (kn <= 32) dtau_acc += mb_unit_q<HD, 4>(a, sq, sk, sv, sdo, slse, srq, sD, q0, qn, k0, h, recA, recB, kbase, lane, col, inv_tau, inv_qs2);
          else if (kn <= 48) dtau_acc += mb_unit_q<HD, 6>(a, sq, sk, sv, sdo, slse, srq, sD, q0, qn, k0, h, recA, recB, kbase, lane, col, inv_tau, inv_qs2);
          else dtau_acc += mb_unit_q<HD, 8>(a, sq, sk, sv, sdo, slse, srq, sD, q0, qn, k0, h, recA, recB, kbase, lane, col, inv_tau, inv_qs2);
          __threadfence_block();
          __syncwarp();
          if (lane == 0) atomicAdd(sdone + k0 * 4 + h, 1);
        } else {
          const int need = kn > 16 ? (kn + 15) >> 4 : 1;
          if (lane == 0) {
            const volatile int* flag = sdone + k0 * 4 + h;
            int spin = 0;
            for (; *flag < need && spin < (1 << 20); ++spin) __nanosleep(32);   // bounded: a logic error must not hang the GPU
            if (spin == (1 << 20)) sra_wait_timed_out();
          }
          __syncwarp();
          __threadfence_block();
          if (kn <= 16) mb_unit_kv<HD, 2>(a, sq, sk, sv, sdo, slse, srk, sD, q0, qn, k0, h, recA, recB, kbase, lane, col);
          else if (kn <= 32) mb_unit_kv<HD, 4>(a, sq, sk, sv, sdo, slse, srk, sD, q0, qn, k0, h, recA, recB, kbase, lane, col);
          else if (kn <= 48) mb_unit_kv<HD, 6>(a, sq, sk, sv, sdo, slse, srk, sD, q0, qn, k0, h, recA, recB, kbase, lane, col);
          else mb_unit_kv<HD, 8>(a, sq, sk, sv, sdo, slse, srk, sD, q0, qn, k0, h, recA, recB, kbase, lane, col);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_empty[st]);
    }
  }
  // sum dS*S of this CTA (natural-log scores: S = S' ln2)
  dtau_acc = warp_sum(dtau_acc);
  if (lane == 0) s_dtau[warp] = dtau_acc;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < MM_THREADS / 32; ++w) tot += s_dtau[w];
    atomicAdd(a.dtau_sum, (double)tot * 0.6931471805599453);
  }
}

// =====================================================================================================
// Work units of every 64-row bin, built once per window table (a table serves 2 encoder layers, forward and
// backward) instead of by one warp of every CTA for every bin of every launch.  One warp per bin.
// units[bin][0..47] = q0 | qn << 7 | k0 << 12 | kn << 19 (rows relative to the first window that starts in the bin),
// [48] = number of units, [49] = 0 (the kernels' work counter lands on it), rest padding.
__global__ void __launch_bounds__(256) sra_bin_units_kernel(const int4* __restrict__ row_info, int N, int* __restrict__ units) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nbins = (N + MM_BIN - 1) / MM_BIN;
  if (warp >= nbins) return;
  const int bin = warp * MM_BIN;
  int* out = units + (long long)warp * MM_UNIT_STRIDE;
  const int4 f = __ldg(row_info + bin);
  const int row0 = (f.y == bin) ? bin : f.z;
  int row1 = N;
  if (bin + MM_BIN < N) {
    const int4 l = __ldg(row_info + bin + MM_BIN);
    row1 = (l.y == bin + MM_BIN) ? bin + MM_BIN : l.z;
  }
  const int R = max(row1 - row0, 0);
  int nu = 0;
  if (R > 0) {
    unsigned long long m0 = 0, m1 = 0;     // bit r: row row0 + r starts a window
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int r = 32 * w + lane;
      const bool st = r < R && __ldg(row_info + row0 + min(r, R - 1)).y == row0 + r;
      const unsigned long long b = __ballot_sync(0xffffffffu, st);
      if (w < 2) m0 |= b << (32 * w);
      else m1 |= b << (32 * (w - 2));
    }
    if (R < 64) m0 |= 1ull << R;           // sentinel: one past the last row
    else m1 |= 1ull << (R - 64);
    int s = 0;
    while (s < R && nu < MM_UNITS) {
      const unsigned w16 = mm_bits(m0, m1, s + 1) & 0xffffu;     // window starts at rows s+1 .. s+16
      if (w16) {                                                 // run of whole windows with <= 16 rows in total
        const int e = s + 32 - __clz(w16);
        if (lane == 0) out[nu] = s | ((e - s) << 7) | (s << 12) | ((e - s) << 19);
        ++nu;
        s = e;
      } else {                                                   // a window of more than 16 rows: 16-row query chunks
        const unsigned lo = mm_bits(m0, m1, s + 17), hi = mm_bits(m0, m1, s + 49);
        const int n = lo ? 16 + __ffs(lo) : 48 + __ffs(hi);
        for (int m = 0; m < n && nu < MM_UNITS; m += 16) {
          if (lane == 0) out[nu] = (s + m) | (min(16, n - m) << 7) | (s << 12) | (n << 19);
          ++nu;
        }
        s += n;
      }
    }
  }
  if (lane < MM_UNIT_STRIDE - MM_UNITS) out[MM_UNITS + lane] = lane == 0 ? nu : 0;
}

extern "C" size_t gdmae_sra_bin_units_bytes(int64_t N) { return (size_t)((N + MM_BIN - 1) / MM_BIN + 1) * MM_UNIT_STRIDE * 4; }

// bin_units (gdmae_sra_bin_units_bytes(N), 16-byte aligned) <- work units of the CSR rows described by row_info (N,4)
extern "C" int gdmae_sra_bin_units(const int32_t* row_info, int64_t N, int32_t* bin_units, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 27) && ((uintptr_t)row_info % 16) == 0 && ((uintptr_t)bin_units % 16) == 0);
  if (N == 0) return GDMAE_OK;
  const long long nbins = (N + MM_BIN - 1) / MM_BIN;
  sra_bin_units_kernel<<<(unsigned)((nbins * 32 + 255) / 256), 256, 0, (cudaStream_t)stream_>>>((const int4*)row_info, (int)N, bin_units);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// number of bounded waits inside the tensor-core SRA kernels that timed out since the library was loaded (synchronises
// the device; 0 in a healthy run)
extern "C" int gdmae_sra_wait_timeouts(int* out) {
  unsigned int v = 0;
  GDMAE_CHECK_ARG(out != nullptr);
  GDMAE_CHECK_CUDA(cudaMemcpyFromSymbol(&v, g_sra_wait_timeouts, sizeof(v)));
  *out = (int)v;
  return GDMAE_OK;
}

static int mm_attrs() {
  static bool done = false;
  if (!done) {
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_BYTES));
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_BYTES));
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_bwd_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM_BYTES));
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_bwd_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM_BYTES));
    done = true;
  }
  return GDMAE_OK;
}

// Tensor-core variant of gdmae_sra_attention_fwd for bf16 q/k/v: qkv (N, 3d) bf16, otherwise the same arguments and outputs.
extern "C" int gdmae_sra_attention_fwd_tc(const void* qkv_bf16, const float* lut, const int32_t* row_info, const int32_t* bin_units,
                                          int64_t N, int d, int nhead, const float* tau, float tau_min, const float* bv, int io_bf16,
                                          void* out, float* lse, void* stream_) {
  GDMAE_CHECK_ARG(bin_units != nullptr && ((uintptr_t)bin_units % 16) == 0);
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 27) && nhead == 8 && (d == 128 || d == 256));
  GDMAE_CHECK_ARG(((uintptr_t)row_info % 16) == 0 && ((uintptr_t)qkv_bf16 % 16) == 0 && ((uintptr_t)lut % 8) == 0);
  GDMAE_CHECK_ARG(bv == nullptr || ((uintptr_t)bv % 8) == 0);
  if (N == 0) return GDMAE_OK;
  int rc = mm_attrs();
  if (rc) return rc;
  MmArgs a{(const bf16*)qkv_bf16, lut, (const int4*)row_info, bin_units, tau, bv, tau_min, (int)N, d, io_bf16};
  cudaStream_t st = (cudaStream_t)stream_;
  // one CTA per SM; 148 is a multiple of the 2 (d = 128) and 4 (d = 256) channel slices
  if (d == 128) sra_fwd_mma_kernel<16><<<GDMAE_NUM_SMS, MF_THREADS, MM_SMEM_BYTES, st>>>(a, out, lse);
  else sra_fwd_mma_kernel<32><<<GDMAE_NUM_SMS, MF_THREADS, MM_SMEM_BYTES, st>>>(a, out, lse);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// Tensor-core backward for bf16 tensors: qkv (N,3d), dout (N,d) and dqkv (N,3d) are bf16; lse (N,8) from the forward;
// dtau_sum (1, double, caller zeroes) accumulates sum dS*S as gdmae_sra_attention_bwd does.  The value bias and the
// forward output are not needed (sum_j P dP replaces dO.(o - bv)).
extern "C" int gdmae_sra_attention_bwd_tc(const void* qkv_bf16, const float* lut, const int32_t* row_info, const int32_t* bin_units,
                                          int64_t N, int d, int nhead, const float* tau, float tau_min, const float* lse,
                                          const void* dout_bf16, void* dqkv_bf16, double* dtau_sum, void* stream_) {
  GDMAE_CHECK_ARG(bin_units != nullptr && ((uintptr_t)bin_units % 16) == 0);
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 27) && nhead == 8 && (d == 128 || d == 256));
  GDMAE_CHECK_ARG(((uintptr_t)row_info % 16) == 0 && ((uintptr_t)qkv_bf16 % 16) == 0 && ((uintptr_t)lut % 8) == 0);
  GDMAE_CHECK_ARG(((uintptr_t)dout_bf16 % 16) == 0 && ((uintptr_t)dqkv_bf16 % 16) == 0 && ((uintptr_t)lse % 4) == 0);
  if (N == 0) return GDMAE_OK;
  int rc = mm_attrs();
  if (rc) return rc;
  MbArgs a{(const bf16*)qkv_bf16, lut, (const int4*)row_info, bin_units, tau, lse, (const bf16*)dout_bf16, (bf16*)dqkv_bf16, dtau_sum, tau_min,
           (int)N, d};
  cudaStream_t st = (cudaStream_t)stream_;
  if (d == 128) sra_bwd_mma_kernel<16><<<GDMAE_NUM_SMS, MM_THREADS, MB_SMEM_BYTES, st>>>(a);
  else sra_bwd_mma_kernel<32><<<GDMAE_NUM_SMS, MM_THREADS, MB_SMEM_BYTES, st>>>(a);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}
