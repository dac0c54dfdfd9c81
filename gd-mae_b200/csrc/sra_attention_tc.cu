// Sparse Regional Attention on the tensor cores: bf16 mma.sync m16n8k16 with fp32 accumulation, operands fetched with
// ldmatrix from 128B-swizzled shared-memory tiles that TMA fills - the kernels of the bf16 configuration.
//
// Same operator as sra_attention.cu (reference: cosine_msa.py:114-176, sst_basic_block.py:22-54, sst_utils.py:107-181):
// flat tokens, CSR windows, 64-row positional LUT, nothing padded in HBM.  The fp32 SIMT kernels are the parity path.
//
// r2 layout ("window-major"): the in-projection GEMM's epilogue (tc_gemm.cu, mode 4) adds the positional LUT row,
// L2-normalises q and k per head in fp32, folds log2(e)/tau into q and writes q^, k^, v as bf16 ROWS IN CSR (WINDOW)
// ORDER: qkvw[tensor][d/64 slices][row][64] (dO joins as a fourth tensor for the backward, mode 5).  The rows of a bin are
// then one contiguous rectangle per (tensor, slice), so a bin is fetched by TMA: ONE cp.async.bulk.tensor (4-D box: 64
// channels x 16 rows x 1 slice x all tensors) per 16 rows, issued by one thread - the r1 kernel gathered 128-byte row
// pieces with 1536 cp.async instructions per bin and re-did the normalisation in every launch (forward, and again in
// backward); a copy-only build of it ran at 1.9 TB/s.
//   * persistent, one CTA per SM, each owning a 64-channel slice (2 heads of 32 or 4 heads of 16) and walking bins of 64
//     CSR rows (the windows that START in the bin, <= 127 rows);
//   * warp 0 is the producer: per bin one mbarrier.expect_tx, ONE bulk copy of the bin's table block (work units + row
//     records; + the per-row scalars in the backward) and ceil(rows / 16) boxes; the other 15 warps run the MMAs.  The two
//     roles meet only through mbarriers (bin landed / bin consumed) - no CTA barrier inside the bin loop;
//   * staging memory is a RING of 16-row groups ({q, k, v[, dO]} boxes of 16 rows each, SWIZZLE_128B: 16-byte chunk c of
//     row r at chunk c ^ (r & 7), so every ldmatrix phase touches the 32 banks once and nothing is padded); a bin takes
//     ceil(rows / 16) consecutive groups, ~7 bins are in flight in the forward kernel;
//   * the work units of every bin are built ONCE per window table (gdmae_sra_bin_units: a table serves two layers,
//     forward and backward) and arrive with the bin's row records; (unit, head) entries are handed out through a
//     shared-memory counter that arrives zeroed with them;
//   * work units are PACKED: a unit is either a run of whole small windows totalling <= 16 rows, or a 16-row chunk of a
//     large window; per-row key bounds (from row_info) give the block-diagonal mask.  A 3-token window therefore costs
//     3/16 of an MMA tile instead of a whole one;
//   * one warp per (unit, head) - per (unit, head pair) for 16-channel heads, the two heads interleaved as independent
//     instruction streams: S = Q K^T accumulates in registers, the softmax runs on the C fragments with quad shuffles,
//     P goes back into the second MMA as the A operand directly from registers, V comes in through ldmatrix.trans.
//     The unit body is compiled for 16 / 32 / 48 / 64 keys.
// The entry points that take the flat (N, 3d) bf16 q|k|v of r1 (gdmae_sra_attention_fwd_tc / _bwd_tc) run a small
// re-layout kernel first (sra_prep_kernel) and exist for callers outside the fused encoder layer and for the tests.
// bf16 operands (8-bit mantissa): tests hold these kernels to 1e-2 against the fp32 kernels / the oracle.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

#define MM_BIN 64
#define MM_ROWS 144            // 127 rows + 16 rows of MMA overrun + 1
#define MM_INFO 128            // row_info records per bin
#define MM_SLICE 64
#define MM_BOX 16              // rows per TMA box (2 KB)
#define MM_GROUPS (MM_ROWS / MM_BOX)           // 16-row groups of a staged bin
#define MF_GROUP (3 * MM_BOX * MM_SLICE)      // forward: bf16 elements of one group = the q | k | v boxes of 16 rows (6 KB)
#define MB_GROUP (4 * MM_BOX * MM_SLICE)      // backward: q | k | v | dO (8 KB)
#ifndef MM_THREADS
#define MM_THREADS 512         // backward CTA size
#endif
#ifndef MF_THREADS
#define MF_THREADS 512         // forward CTA size (the register cap follows: 65536 / MF_THREADS)
#endif
// forward: heads one math warp runs interleaved per work entry.  Measured (tools/sweep_sra_tc.py, r1): two heads win for the
// 16-channel heads of d = 128 (28.7 vs 30.7 us), one head for the 32-channel heads of d = 256 (78 vs 89 us: the two-head body
// sits at the 128-register cap and halves the entries the math warps can share)
#ifdef MM_HPE
#define MM_HEADS_PER_ENTRY(HD) (MM_HPE)
#else
#define MM_HEADS_PER_ENTRY(HD) ((HD) == 16 ? 2 : 1)
#endif
#define MM_MATH_WARPS (MF_THREADS / 32 - 1)
// Staging memory is a RING of 16-row groups, not worst-case (144-row) stages: a bin takes ceil(rows / 16) consecutive groups
// (4.5 on average), so ~7 bins are in flight in the forward kernel instead of 3.  r2 measurement: with 3 stages the period
// per bin was (release -> issue -> landing -> slowest entry) / 3 = ~3000 cycles whatever the math per bin - a latency-bound
// pipeline; the 16 MMA-overrun rows behind a bin's last group may belong to a neighbour (masked, only finite values needed).
#ifndef MF_SLOTS
#define MF_SLOTS 8             // forward: bins in flight (table block + barriers per slot)
#endif
#ifndef MF_RING
#define MF_RING 33             // forward: groups of the ring (+ 1 never-written pad group behind it)
#endif
#define MM_UNITS 48
#define MM_UNIT_STRIDE 64      // ints of the unit list: [0, 48) units, [48] count, [49] work counter (0), [50] first row, [51] rows, pad
#define MM_BLOCK_INTS (MM_UNIT_STRIDE + 4 * MM_INFO)   // per-bin block of the table: unit list + the 128 row records from the bin's first row (2304 B)
#define MM_RR_SLOTS 256        // (first row, rows) of a CTA's first bins, preloaded into shared memory
#define MM_SMEM_BYTES ((MF_RING + 1) * MF_GROUP * 2 + MF_SLOTS * (MM_BLOCK_INTS * 4 + 8 + 16) + MM_RR_SLOTS * 8 + 1024)

typedef __nv_bfloat16 bf16;

struct MmArgs {
  const int* bin_units; // per-bin blocks (gdmae_sra_bin_units)
  const float* bv;      // (d) value bias added to the output, nullable
  int N, d;
  int out_bf16;
  int lse_by_row;       // lse goes to column h of the (N, 24) per-row record of the backward (CSR-row order) instead of (N, 8) by token
};

// ---- mbarriers (shared::cta): producer/consumer hand-over of stage buffers without CTA-wide barriers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
#ifndef MM_WAIT_HINT
#define MM_WAIT_HINT 20000u
#endif
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  for (int spin = 0; spin < 50000 * (20000u / MM_WAIT_HINT); ++spin) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity), "r"(MM_WAIT_HINT) : "memory");
    if (ok) return;
  }
  sra_wait_timed_out();
}

// ---- the producer's copies (async proxy): a box = 16 rows of every tensor (q, k, v[, dO]) of one channel slice of the
// (64 channels, N rows, d/64 slices, 3 or 4 tensors) window-major array - one TMA operation per 16-row group (a TMA
// operation costs ~80 cycles of the engine whatever its size: r2 measurement, tools/sweep_sra_tc.py) - and plain byte ranges
__device__ __forceinline__ void tma_box(void* dst, const CUtensorMap* map, unsigned long long* bar, int row, int slice) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"((unsigned long long)map), "r"(smem_u32(bar)), "r"(0), "r"(row), "r"(slice), "r"(0) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// element offset of (row, channel) of the first tensor inside a staged bin: groups of 16 rows, GE elements apart (the boxes of
// the other tensors follow at + 1024 elements each), SWIZZLE_128B inside a box (16-byte chunk c of row r at chunk c ^ (r & 7))
template <int GE>
__device__ __forceinline__ int sw_off(int row, int col) {
  return (row >> 4) * GE + ((row & 15) << 6) + ((((col >> 3) ^ row) & 7) << 3) + (col & 7);
}
#define SWF(row, col) sw_off<MF_GROUP>(row, col)
#define SWB(row, col) sw_off<MB_GROUP>(row, col)

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

// bits [pos, pos+32) of the 128-bit mask (m1:m0)
__device__ __forceinline__ unsigned mm_bits(unsigned long long m0, unsigned long long m1, int pos) {
  unsigned long long v;
  if (pos >= 128) return 0u;
  if (pos >= 64) v = m1 >> (pos - 64);
  else v = (m0 >> pos) | (pos ? (m1 << (64 - pos)) : 0ull);
  return (unsigned)v;
}

// One unit and NH adjacent heads on one warp: S = Q K^T, masked softmax, O = P V, for NT2 8-key tiles (NT2 even).  The NH
// heads are independent instruction streams over the same rows and masks; every step is written as a loop over the heads
// so that their dependency chains (ldmatrix -> mma -> shuffle -> exp2 -> mma) overlap in the one warp.
template <int HD, int NT2, int NH, int FUSED>
__device__ __forceinline__ void mm_unit(const MmArgs& a, const bf16* sq, const bf16* sk, const bf16* sv, int q0, int qn, int k0,
                                        int ch, int4 recA, int4 recB, int kbase, int rowbase, int lane, long long out_col, int lse_col,
                                        void* __restrict__ out, float* __restrict__ lse) {
  constexpr int KS = HD / 16, ND = HD / 8;
  const int g = lane >> 2, t = lane & 3;
  const int loA = recA.y - kbase, wA = recA.z - recA.y, loB = recB.y - kbase, wB = recB.z - recB.y;
  // Operand addresses: three swizzled base offsets per entry; every further tile is a compile-time offset from them.  Rows:
  // + 16 rows = + one group; + 8 rows alternates between + 512 elements and + (group - 512) depending on bit 3 of the first
  // row.  Channels: + 16 / + 32 channels toggles bit 1 / 2 of the 16-byte chunk index, which the bases leave clear
  // (ch is a multiple of 32 wherever a channel step is taken), so it is an XOR on the offset.  (r2 profile: the generic
  // form spent ~7 integer instructions on each of the 11-25 ldmatrix addresses of an entry.)
  const int qoff = SWF(q0 + (lane & 7) + 8 * ((lane >> 3) & 1), ch + 8 * (lane >> 4));
  const int krow = KS == 2 ? k0 + (lane & 7) : k0 + 8 * (lane >> 4) + (lane & 7);
  const int koff0 = SWF(krow, ch + (KS == 2 ? 8 * (lane >> 3) : 8 * ((lane >> 3) & 1)));
  const int koffA = koff0 + ((krow & 8) ? MF_GROUP - 512 : 512);          // KS == 2: the odd 8-key tiles
  const int voff = SWF(k0 + (lane & 7) + 8 * ((lane >> 3) & 1), ch + 8 * (lane >> 4));
  unsigned qa[NH][KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int hh = 0; hh < NH; ++hh)
      ldsm_x4(qa[hh][ks], sq + (qoff ^ (hh * HD + 16 * ks)));
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
          ldsm_x4(kb[hh], sk + (((u ? koffA : koff0) + np * MF_GROUP) ^ (hh * HD)));
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) mma_bf16(c[hh][2 * np + u], qa[hh][0], kb[hh][0], kb[hh][1]);
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) mma_bf16(c[hh][2 * np + u], qa[hh][KS - 1], kb[hh][2], kb[hh][3]);
      }
    } else {
      unsigned kb[NH][4];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh)
        ldsm_x4(kb[hh], sk + ((koff0 + np * MF_GROUP) ^ (hh * HD)));
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
        ldsm_x4_t(vb[hh], sv + ((voff + kt * MF_GROUP) ^ (hh * HD + 16 * np)));
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
          if (FUSED || a.out_bf16) *reinterpret_cast<unsigned*>((bf16*)out + e0 + 8 * nd) = pack_bf16(x0, x1);
          else *reinterpret_cast<float2*>((float*)out + e0 + 8 * nd) = make_float2(x0, x1);
        }
        // natural-log lse of the scores S = cos / tau (the backward kernels expect it)
        // (lse_by_row: indexed by CSR row - the layout the window-major backward reads with one bulk copy per bin)
        if (t == 0)
          lse[((FUSED || a.lse_by_row) ? (long long)(rowbase + g + 8 * half) * 24 : (long long)tok * 8) + lse_col + hh] =
              ((half ? mx1[hh] : mx0[hh]) + __log2f(half ? l1[hh] : l0[hh])) * 0.6931471805599453f;
      }
    }
  }
}
#ifdef MM_PROFILE
// development build only (tools/sweep_sra_tc.py -DMM_PROFILE): cycle counters of the producer and one math warp per CTA
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

// tmQ: (64 channels, N rows, d/64 slices, 3 tensors) bf16 view of the window-major q^ | k^ | v, box 64 x 16 x 1 x 3, SWIZZLE_128B
// FUSED = 1: bf16 output and lse into the per-row records (the encoder layer's call) as compile-time facts
template <int HD, int FUSED>
__global__ void __launch_bounds__(MF_THREADS, 1) sra_fwd_mma_kernel(const __grid_constant__ CUtensorMap tmQ, MmArgs a, void* __restrict__ out,
                                                                    float* __restrict__ lse) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // swizzled boxes: 1024-byte aligned
  bf16* sdata = (bf16*)smem_raw;                              // ring: (MF_RING + 1) groups x {q, k, v} x [16][64]
  int* sblock_all = (int*)(sdata + (MF_RING + 1) * MF_GROUP); // MF_SLOTS x { unit list (64 ints) | 128 row records }
  int2* s_rr = (int2*)(sblock_all + MF_SLOTS * MM_BLOCK_INTS);          // (first row, rows) of this CTA's bins
  unsigned long long* s_full = (unsigned long long*)(s_rr + MM_RR_SLOTS);   // bin landed
  unsigned long long* s_empty = s_full + MF_SLOTS;                          // bin consumed
  int* s_base = (int*)(s_empty + MF_SLOTS);                   // first ring group of the slot's bin
  int* s_alloc = s_base + MF_SLOTS;                           // groups the slot's bin holds (producer only)
  constexpr int HS = MM_SLICE / HD;
  const int d = a.d;
  const int nsl = d / MM_SLICE;
  const int sl = blockIdx.x % nsl, cta = blockIdx.x / nsl, ncta = gridDim.x / nsl;
  const int col = sl * MM_SLICE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbins = (a.N + MM_BIN - 1) / MM_BIN;
  if (cta >= nbins) return;
  const int my_bins = (nbins - cta + ncta - 1) / ncta;        // bins cta, cta + ncta, ...
  MM_PROF_T(k0_);

  // ---- prologue: (first row, rows) of the CTA's bins, zeroed data buffers (overrun rows must hold finite values), barriers
  if (tid < MM_RR_SLOTS && tid < my_bins)
    s_rr[tid] = __ldg(reinterpret_cast<const int2*>(a.bin_units + (long long)(cta + tid * ncta) * MM_BLOCK_INTS + MM_UNITS + 2));
  for (int i = tid; i < (MF_RING + 1) * MF_GROUP / 8; i += MF_THREADS) reinterpret_cast<uint4*>(sdata)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < MF_SLOTS; ++st) {
      mbar_init(&s_full[st], 1);                 // the producer's expect_tx arrival + the bytes of its copies
      mbar_init(&s_empty[st], MM_MATH_WARPS);    // every math warp arrives once it has no more work in the bin
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zero fill is ordered before the first TMA writes
  __syncthreads();     // the only CTA-wide barrier: from here the two roles meet through the mbarriers
  MM_PROF_T(k1_);

  if (warp == 0) {
    // ================= producer (one thread): bin j takes slot j % MF_SLOTS and ceil(rows / 16) consecutive groups of the
    // ring (allocated in FIFO order, never split across the end of the ring), as soon as the math warps have released
    // enough of the oldest bins.  One transaction per bin: the bin's block of the table (unit list incl. the zeroed work
    // counter + row records) and one box per 16 rows (q^, k^ and v together).
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmQ) : "memory");
      int head = 0, used = 0, tail = 0;          // next free group, groups held by the bins tail .. j-1, oldest unreclaimed bin
      for (int j = 0; j < my_bins; ++j) {
        const int bi = cta + j * ncta, st = j % MF_SLOTS;
        const int2 rr = j < MM_RR_SLOTS ? s_rr[j]
                                        : __ldg(reinterpret_cast<const int2*>(a.bin_units + (long long)bi * MM_BLOCK_INTS + MM_UNITS + 2));
        const int row0 = rr.x, nbox = (rr.y + MM_BOX - 1) / MM_BOX;
        MM_PROF_T(t0);
        int need;
        for (;;) {
          need = nbox + (head + nbox > MF_RING ? MF_RING - head : 0);   // the tail of the ring is skipped (and held) if the bin does not fit
          if (j - tail < MF_SLOTS && used + need <= MF_RING) break;
          mbar_wait(&s_empty[tail % MF_SLOTS], (tail / MF_SLOTS) & 1);
          used -= s_alloc[tail % MF_SLOTS];
          ++tail;
        }
        MM_PROF_T(t1);
        const int start = head + nbox > MF_RING ? 0 : head;
        head = start + nbox;
        used += need;
        s_alloc[st] = need;
        s_base[st] = start;
#ifdef MM_DEBUG_NO_COPY
        mbar_expect_tx(&s_full[st], (unsigned)((j < 8 ? nbox * MF_GROUP * 2 : 0) + MM_BLOCK_INTS * 4));
#else
        mbar_expect_tx(&s_full[st], (unsigned)(nbox * MF_GROUP * 2 + MM_BLOCK_INTS * 4));
#endif
        bulk_copy(sblock_all + st * MM_BLOCK_INTS, a.bin_units + (long long)bi * MM_BLOCK_INTS, MM_BLOCK_INTS * 4, &s_full[st]);
        bf16* tile = sdata + start * MF_GROUP;
#ifdef MM_DEBUG_NO_COPY
        if (j < 8)      // development build: math on whatever the first bins left in the ring
#endif
        for (int i = 0; i < nbox; ++i) tma_box(tile + i * MF_GROUP, &tmQ, &s_full[st], row0 + i * MM_BOX, sl);
#ifdef MM_PROFILE
        MM_PROF_T(t2);
        MM_PROF_ADD(0, t1 - t0); MM_PROF_ADD(1, t2 - t1); MM_PROF_ADD(5, 1);
#endif
      }
    }
  } else {
    // ================= math warps: entries of bin j (a unit and one or two heads), handed out through the bin's
    // work counter; a warp that finds the bin exhausted signals and moves on to bin j+1 without waiting for the others
    const int g = lane >> 2;
    for (int j = 0; j < my_bins; ++j) {
      const int st = j % MF_SLOTS;
      int* sunit = sblock_all + st * MM_BLOCK_INTS;
      const int4* sinfo = (const int4*)(sunit + MM_UNIT_STRIDE);      // records of rows row0 .. row0 + 127
      MM_PROF_T(m0);
      mbar_wait(&s_full[st], (j / MF_SLOTS) & 1);
      MM_PROF_T(m1);
      const bf16* sq = sdata + s_base[st] * MF_GROUP;
      const bf16* sk = sq + MM_BOX * MM_SLICE;
      const bf16* sv = sk + MM_BOX * MM_SLICE;
#ifdef MM_PROFILE
      int n_done = 0;
#endif
      const int row0 = sunit[MM_UNITS + 2], R = sunit[MM_UNITS + 3];
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
        const int4 recA = sinfo[min(q0 + g, R - 1)], recB = sinfo[min(q0 + g + 8, R - 1)];
        const int kbase = row0 + k0, rowbase = row0 + q0;
        const long long oc = col + ch;
        const int lc = sl * HS + h;
        if (kn <= 16) mm_unit<HD, 2, HPE, FUSED>(a, sq, sk, sv, q0, qn, k0, ch, recA, recB, kbase, rowbase, lane, oc, lc, out, lse);
        else if (kn <= 32) mm_unit<HD, 4, HPE, FUSED>(a, sq, sk, sv, q0, qn, k0, ch, recA, recB, kbase, rowbase, lane, oc, lc, out, lse);
        else {
          // large windows: one head at a time (the two-head body would spill at the register cap)
#pragma unroll 1
          for (int hh = 0; hh < HPE; ++hh) {
            if (kn <= 48) mm_unit<HD, 6, 1, FUSED>(a, sq, sk, sv, q0, qn, k0, ch + hh * HD, recA, recB, kbase, rowbase, lane, oc + hh * HD, lc + hh, out, lse);
            else mm_unit<HD, 8, 1, FUSED>(a, sq, sk, sv, q0, qn, k0, ch + hh * HD, recA, recB, kbase, rowbase, lane, oc + hh * HD, lc + hh, out, lse);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
#ifdef MM_PROFILE
      if (warp == 1 || warp == MF_THREADS / 32 - 1) {
        MM_PROF_T(m2);
        const int o = warp == 1 ? 6 : 10;
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
// Backward.  Same bins, packing, tiles and producer / math decoupling as the forward kernel; four staged tiles
// (Qs = q_hat * log2(e)/tau, K_hat, V, dO - all window-major, one TMA box per 16 rows) in the ring, the per-row records
// (lse, 1/|q|, 1/|k| as the forward pass left them, CSR-row order: one bulk copy per bin) and D per slot:
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
// dq, dk, dv rows go straight from the fragments to the flat (N, 3d) dqkv (bf16, token order: the operand of the
// in-projection's gradient GEMMs); sum dS S is reduced per CTA into dtau_sum.
#ifndef MB_SLOTS
#define MB_SLOTS 3             // backward: bins in flight (sweep r2: 3 slots + 19 groups 137 us, 4 + 16 147 us, 2 + 18 143 us at d = 256)
#endif
#ifndef MB_RING
#define MB_RING 19             // backward: groups of the ring (+ 1 never-written pad group)
#endif
#define MB_SCAL (MM_ROWS * 24)     // per-row record of 24 fp32: lse | 1/|q| | 1/|k|, each per head
#define MB_MATH_WARPS (MM_THREADS / 32 - 1)
// ring of {q,k,v,dO} groups | per slot: row records [144][24], D [144][4], table block, query-side completion counters, barriers
#define MB_SLOT_BYTES (MB_SCAL * 4 + MM_ROWS * 4 * 4 + MM_BLOCK_INTS * 4 + 128 * 4 * 4 + 8 + 16)
#define MB_SMEM_BYTES ((MB_RING + 1) * MB_GROUP * 2 + MB_SLOTS * MB_SLOT_BYTES + MM_RR_SLOTS * 8 + 1024)

struct MbArgs {
  const int* bin_units; // per-bin blocks (gdmae_sra_bin_units)
  const float* tau;
  const float* lrr;    // (N, 24) by CSR row: lse (forward kernel) | 1/|q| | 1/|k| (in-projection epilogue), 8 heads each
  bf16* dqkv;          // (N, 3d) bf16, token order
  double* dtau_sum;
  float tau_min;
  int N, d;
};

__device__ __forceinline__ float2 unpack_bf16(unsigned u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }

// query side of one (unit, head): writes dq rows and sD, returns this lane's share of sum dS*S' (valid rows only)
template <int HD, int NT2>
__device__ __forceinline__ float mb_unit_q(const MbArgs& a, const bf16* sq, const bf16* sk, const bf16* sv, const bf16* sdo,
                                           const float* slse, const float* srq, float* sD, int* sdone_slot, int q0, int qn, int k0, int h, int hg, int4 recA,
                                           int4 recB, int kbase, int lane, int col, float inv_tau, float inv_qs2) {
  constexpr int KS = HD / 16, ND = HD / 8;
  const int g = lane >> 2, t = lane & 3, ch = h * HD;
  const int loA = recA.y - kbase, wA = recA.z - recA.y, loB = recB.y - kbase, wB = recB.z - recB.y;
  const int ka = 2 * t - loA, kb_ = 2 * t - loB;
  // operand addresses as in the forward kernel: swizzled bases, compile-time group / half-group steps, XOR channel steps
  const int qoff = SWB(q0 + (lane & 7) + 8 * ((lane >> 3) & 1), ch + 8 * (lane >> 4));
  const int krow = KS == 2 ? k0 + (lane & 7) : k0 + 8 * (lane >> 4) + (lane & 7);
  const int koff0 = SWB(krow, ch + (KS == 2 ? 8 * (lane >> 3) : 8 * ((lane >> 3) & 1)));
  const int koffA = koff0 + ((krow & 8) ? MB_GROUP - 512 : 512);
  const int voff = SWB(k0 + (lane & 7) + 8 * ((lane >> 3) & 1), ch + 8 * (lane >> 4));
  unsigned qa[KS][4], da[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int off = qoff ^ (16 * ks);
    ldsm_x4(qa[ks], sq + off);
    ldsm_x4(da[ks], sdo + off);
  }
  const float lA = slse[min(q0 + g, MM_ROWS - 1) * 24 + hg] * 1.4426950408889634f;
  const float lB = slse[min(q0 + g + 8, MM_ROWS - 1) * 24 + hg] * 1.4426950408889634f;
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
        const int off = (u ? koffA : koff0) + np * MB_GROUP;
        ldsm_x4(kb, sk + off);
        ldsm_x4(vb, sv + off);
        mma_bf16(s[u], qa[0], kb[0], kb[1]);
        mma_bf16(s[u], qa[KS - 1], kb[2], kb[3]);
        mma_bf16(dp[2 * np + u], da[0], vb[0], vb[1]);
        mma_bf16(dp[2 * np + u], da[KS - 1], vb[2], vb[3]);
      }
    } else {
      const int off = koff0 + np * MB_GROUP;
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
  // D of these rows is published here, before the long tail of the entry: the key-side entries of the window wait for it
  // (the fence would otherwise also wait for this entry's global dq stores: r2 profile, ~4 % of the warp samples)
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    atomicAdd(sdone_slot, 1);
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
      ldsm_x4_t(kb, sk + ((voff + kt * MB_GROUP) ^ (16 * np)));
      mma_bf16(acc[2 * np], sa, kb[0], kb[1]);
      mma_bf16(acc[2 * np + 1], sa, kb[2], kb[3]);
    }
  }
  // normalisation backward on the fragments: dq = (G - Qs (Qs.G) / qscale^2) / (tau |q|)
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = q0 + g + 8 * half;
    const int noff = SWB(row, ch + 2 * t);
    float2 qv[ND];
    float dot = 0.f;
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      qv[nd] = unpack_bf16(*reinterpret_cast<const unsigned*>(sq + (noff ^ (8 * nd))));
      dot = fmaf(qv[nd].x, acc[nd][2 * half], dot);
      dot = fmaf(qv[nd].y, acc[nd][2 * half + 1], dot);
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    if (g + 8 * half < qn) {
      const float f = srq[row * 24 + hg] * inv_tau, c2 = dot * inv_qs2;
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
                                           const float* slse, const float* srk, const float* sD, int q0, int qn, int k0, int h, int hg,
                                           int4 recA, int4 recB, int kbase, int lane, int col) {
  constexpr int KS = HD / 16, ND = HD / 8;
  const int g = lane >> 2, t = lane & 3, ch = h * HD;
  const int loA = recA.y - kbase, wA = recA.z - recA.y, loB = recB.y - kbase, wB = recB.z - recB.y;
  const int ka = 2 * t - loA, kb_ = 2 * t - loB;
  const int aoff = SWB(q0 + (lane & 7) + 8 * ((lane >> 3) & 1), ch + 8 * (lane >> 4));
  const int krow = KS == 2 ? k0 + (lane & 7) : k0 + 8 * (lane >> 4) + (lane & 7);
  const int koff0 = SWB(krow, ch + (KS == 2 ? 8 * (lane >> 3) : 8 * ((lane >> 3) & 1)));
  const int koffA = koff0 + ((krow & 8) ? MB_GROUP - 512 : 512);
  const int voff = SWB(k0 + (lane & 7) + 8 * ((lane >> 3) & 1), ch + 8 * (lane >> 4));
  unsigned ka_f[KS][4], va_f[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int off = aoff ^ (16 * ks);
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
        const int off = (u ? koffA : koff0) + np * MB_GROUP;
        ldsm_x4(qb, sq + off);
        ldsm_x4(ob, sdo + off);
        mma_bf16(s[u], ka_f[0], qb[0], qb[1]);
        mma_bf16(s[u], ka_f[KS - 1], qb[2], qb[3]);
        mma_bf16(dp[u], va_f[0], ob[0], ob[1]);
        mma_bf16(dp[u], va_f[KS - 1], ob[2], ob[3]);
      }
    } else {
      const int off = koff0 + np * MB_GROUP;
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
      const float l0 = slse[qc * 24 + hg] * 1.4426950408889634f, l1 = slse[(qc + 1) * 24 + hg] * 1.4426950408889634f;
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
      const int off = (voff + np * MB_GROUP) ^ (16 * nq);
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
    const int noff = SWB(row, ch + 2 * t);
    float2 kv[ND];
    float dot = 0.f;
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      kv[nd] = unpack_bf16(*reinterpret_cast<const unsigned*>(sk + (noff ^ (8 * nd))));
      dot = fmaf(kv[nd].x, hk[nd][2 * half], dot);
      dot = fmaf(kv[nd].y, hk[nd][2 * half + 1], dot);
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    if (g + 8 * half < qn) {
      const float f = srk[row * 24 + hg] * 0.6931471805599453f;
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
// tmQ: (64 channels, N rows, d/64 slices, 4 tensors) view of the window-major q^ | k^ | v | dO, box 64 x 16 x 1 x 4
template <int HD>
__global__ void __launch_bounds__(MM_THREADS, 1) sra_bwd_mma_kernel(const __grid_constant__ CUtensorMap tmQ, MbArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  bf16* sdata = (bf16*)smem_raw;                               // ring: (MB_RING + 1) groups x {q, k, v, dO} x [16][64]
  float* sscal_all = (float*)(sdata + (MB_RING + 1) * MB_GROUP);            // MB_SLOTS x { records [144][24], D [144][4] }
  constexpr int SCAL_STAGE = MB_SCAL + MM_ROWS * 4;
  int* sblock_all = (int*)(sscal_all + MB_SLOTS * SCAL_STAGE);
  int* sdone_all = sblock_all + MB_SLOTS * MM_BLOCK_INTS;      // MB_SLOTS x [128 window start rows][4 heads]: query-side entries finished
  int2* s_rr = (int2*)(sdone_all + MB_SLOTS * 512);
  unsigned long long* b_full = (unsigned long long*)(s_rr + MM_RR_SLOTS);   // bin landed
  unsigned long long* b_empty = b_full + MB_SLOTS;                          // bin consumed
  int* s_base = (int*)(b_empty + MB_SLOTS);                    // first ring group of the slot's bin
  int* s_alloc = s_base + MB_SLOTS;                            // groups the slot's bin holds (producer only)
  __shared__ float s_dtau[MM_THREADS / 32];
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

  // ---- prologue: (first row, rows) of the CTA's bins, zeroed buffers (overrun rows and scalars must be finite), barriers
  if (tid < MM_RR_SLOTS && tid < my_bins)
    s_rr[tid] = __ldg(reinterpret_cast<const int2*>(a.bin_units + (long long)(cta + tid * ncta) * MM_BLOCK_INTS + MM_UNITS + 2));
  for (int i = tid; i < (MB_RING + 1) * MB_GROUP / 8; i += MM_THREADS) reinterpret_cast<uint4*>(sdata)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < MB_SLOTS * SCAL_STAGE; i += MM_THREADS) sscal_all[i] = 0.f;
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < MB_SLOTS; ++st) {
      mbar_init(&b_full[st], 1);
      mbar_init(&b_empty[st], MB_MATH_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();     // from here the two roles meet through the mbarriers (and once more for the d tau sum)
  const float tau_c = fmaxf(__ldg(a.tau), a.tau_min);
  const float qscale = 1.4426950408889634f / tau_c, inv_tau = 1.f / tau_c, inv_qs2 = 1.f / (qscale * qscale);
  float dtau_acc = 0.f;

  if (warp == 0) {
    // ================= producer warp: slot and ring groups are claimed as in the forward kernel (lane 0), the warp zeroes the
    // slot's completion counters, then lane 0 issues the bin's transaction: table block, row records and one box per 16 rows
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&tmQ) : "memory");
    int head = 0, used = 0, tail = 0;
    for (int k = 0; k < my_bins; ++k) {
      const int bi = cta + k * ncta, st = k % MB_SLOTS;
      int row0 = 0, R = 0, nbox = 0, start = 0;
      if (lane == 0) {
        const int2 rr = k < MM_RR_SLOTS ? s_rr[k]
                                        : __ldg(reinterpret_cast<const int2*>(a.bin_units + (long long)bi * MM_BLOCK_INTS + MM_UNITS + 2));
        row0 = rr.x; R = rr.y; nbox = (R + MM_BOX - 1) / MM_BOX;
        int need;
        for (;;) {
          need = nbox + (head + nbox > MB_RING ? MB_RING - head : 0);
          if (k - tail < MB_SLOTS && used + need <= MB_RING) break;
          mbar_wait(&b_empty[tail % MB_SLOTS], (tail / MB_SLOTS) & 1);
          used -= s_alloc[tail % MB_SLOTS];
          ++tail;
        }
        start = head + nbox > MB_RING ? 0 : head;
        head = start + nbox;
        used += need;
        s_alloc[st] = need;
        s_base[st] = start;
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<int4*>(sdone_all + st * 512)[lane + 32 * i] = make_int4(0, 0, 0, 0);
      __syncwarp();
      if (lane == 0) {
        mbar_expect_tx(&b_full[st], (unsigned)(nbox * MB_GROUP * 2 + MM_BLOCK_INTS * 4 + R * 96));
        bulk_copy(sblock_all + st * MM_BLOCK_INTS, a.bin_units + (long long)bi * MM_BLOCK_INTS, MM_BLOCK_INTS * 4, &b_full[st]);
        if (R > 0) bulk_copy(sscal_all + st * SCAL_STAGE, a.lrr + (long long)row0 * 24, R * 96, &b_full[st]);
        bf16* tile = sdata + start * MB_GROUP;
        for (int i = 0; i < nbox; ++i) tma_box(tile + i * MB_GROUP, &tmQ, &b_full[st], row0 + i * MM_BOX, sl);
      }
    }
  } else {
    // ================= math warps.  Entries [0, nent): query side; [nent, 2 nent): key side, handed out in this order through the
    // work counter that arrived (zero) with the unit list.  A key-side entry needs D of every query row of its window: it waits
    // until the query-side entries covering the window (one for a packed unit, ceil(rows/16) chunks for a large window) have
    // signalled.  Every query-side entry is pulled before any key-side entry and never blocks, so the wait cannot deadlock.  A
    // warp that finds the bin exhausted signals and moves on to bin k+1 without waiting for the others.
    const int g = lane >> 2;
    for (int k = 0; k < my_bins; ++k) {
      const int st = k % MB_SLOTS;
      int* sunit = sblock_all + st * MM_BLOCK_INTS;
      const int4* sinfo = (const int4*)(sunit + MM_UNIT_STRIDE);      // records of rows row0 .. row0 + 127
      float* slse = sscal_all + st * SCAL_STAGE;                      // [row][24]: lse | 1/|q| | 1/|k|
      float* srq = slse + 8;
      float* srk = slse + 16;
      float* sD = slse + MB_SCAL;
      int* sdone = sdone_all + st * 512;
      mbar_wait(&b_full[st], (k / MB_SLOTS) & 1);
      const bf16* sq = sdata + s_base[st] * MB_GROUP;
      const bf16* sk = sq + MM_BOX * MM_SLICE;
      const bf16* sv = sk + MM_BOX * MM_SLICE;
      const bf16* sdo = sv + MM_BOX * MM_SLICE;
      const int row0 = sunit[MM_UNITS + 2], R = sunit[MM_UNITS + 3];
      const int nent = sunit[MM_UNITS] * HS;
      for (;;) {
        int e = 0;
        if (lane == 0) e = atomicAdd(sunit + MM_UNITS + 1, 1);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= 2 * nent) break;
        const bool key_side = e >= nent;
        if (key_side) e -= nent;
        const int code = sunit[e / HS];
        const int h = e % HS, hg = lse_col + h;
        const int q0 = code & 127, qn = (code >> 7) & 31, k0 = (code >> 12) & 127, kn = (code >> 19) & 127;
        const int4 recA = sinfo[min(q0 + g, R - 1)], recB = sinfo[min(q0 + g + 8, R - 1)];
        const int kbase = row0 + k0;
        if (!key_side) {
          if (kn <= 16) dtau_acc += mb_unit_q<HD, 2>(a, sq, sk, sv, sdo, slse, srq, sD, sdone + k0 * 4 + h, q0, qn, k0, h, hg, recA, recB, kbase, lane, col, inv_tau, inv_qs2);
          else if (kn <= 32) dtau_acc += mb_unit_q<HD, 4>(a, sq, sk, sv, sdo, slse, srq, sD, sdone + k0 * 4 + h, q0, qn, k0, h, hg, recA, recB, kbase, lane, col, inv_tau, inv_qs2);
          else if (kn <= 48) dtau_acc += mb_unit_q<HD, 6>(a, sq, sk, sv, sdo, slse, srq, sD, sdone + k0 * 4 + h, q0, qn, k0, h, hg, recA, recB, kbase, lane, col, inv_tau, inv_qs2);
          else dtau_acc += mb_unit_q<HD, 8>(a, sq, sk, sv, sdo, slse, srq, sD, sdone + k0 * 4 + h, q0, qn, k0, h, hg, recA, recB, kbase, lane, col, inv_tau, inv_qs2);
        } else {
          const int need = kn > 16 ? (kn + 15) >> 4 : 1;
          if (lane == 0) {
            const unsigned flag = smem_u32(sdone + k0 * 4 + h);
            int spin = 0;
            for (; spin < (1 << 20); ++spin) {                                   // bounded: a logic error must not hang the GPU
              int v;
              asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(flag) : "memory");
              if (v >= need) break;
              __nanosleep(32);
            }
            if (spin == (1 << 20)) sra_wait_timed_out();
          }
          __syncwarp();
          if (kn <= 16) mb_unit_kv<HD, 2>(a, sq, sk, sv, sdo, slse, srk, sD, q0, qn, k0, h, hg, recA, recB, kbase, lane, col);
          else if (kn <= 32) mb_unit_kv<HD, 4>(a, sq, sk, sv, sdo, slse, srk, sD, q0, qn, k0, h, hg, recA, recB, kbase, lane, col);
          else if (kn <= 48) mb_unit_kv<HD, 6>(a, sq, sk, sv, sdo, slse, srk, sD, q0, qn, k0, h, hg, recA, recB, kbase, lane, col);
          else mb_unit_kv<HD, 8>(a, sq, sk, sv, sdo, slse, srk, sD, q0, qn, k0, h, hg, recA, recB, kbase, lane, col);
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
// The per-bin table, built once per window table (a table serves 2 encoder layers, forward and backward) instead of by
// one warp of every CTA for every bin of every launch.  One warp per 64-row bin writes the bin's BLOCK (576 ints):
// [0..47] work units q0 | qn << 7 | k0 << 12 | kn << 19 (rows relative to the first window that starts in the bin),
// [48] number of units, [49] 0 (the kernels' work counter lands on it), [50] first CSR row of the bin's windows, [51] their
// row count, [64..575] the 128 row records (token, window start, window end, cell) from that first row on - everything a
// CTA needs about a bin arrives with ONE bulk copy.
__global__ void __launch_bounds__(256) sra_bin_units_kernel(const int4* __restrict__ row_info, int N, int* __restrict__ units) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nbins = (N + MM_BIN - 1) / MM_BIN;
  if (warp >= nbins) return;
  const int bin = warp * MM_BIN;
  int* out = units + (long long)warp * MM_BLOCK_INTS;
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
      const int4 rec = __ldg(row_info + min(row0 + r, N - 1));
      reinterpret_cast<int4*>(out + MM_UNIT_STRIDE)[r] = row0 + r < N ? rec : make_int4(0, 0, 0, 0);
      const bool st = r < R && rec.y == row0 + r;
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
  } else {
#pragma unroll
    for (int w = 0; w < 4; ++w) reinterpret_cast<int4*>(out + MM_UNIT_STRIDE)[32 * w + lane] = make_int4(0, 0, 0, 0);
  }
  if (lane < MM_UNIT_STRIDE - MM_UNITS) out[MM_UNITS + lane] = lane == 0 ? nu : (lane == 2 ? row0 : (lane == 3 ? R : 0));
}

// tok_info[token] = CSR row | in-window cell << 26: what the in-projection's epilogue needs per token (destination row of
// the window-major layout, row of the positional LUT)
__global__ void __launch_bounds__(256) sra_tok_info_kernel(const int4* __restrict__ row_info, int N, int* __restrict__ tok_info) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
    const int4 f = __ldg(row_info + r);
    tok_info[f.x] = r | (f.w << 26);
  }
}

static inline long long sra_units_ints(long long N) { return ((N + MM_BIN - 1) / MM_BIN + 1) * MM_BLOCK_INTS; }

extern "C" size_t gdmae_sra_bin_units_bytes(int64_t N) { return (size_t)(sra_units_ints(N) + ((N + 3) & ~3ll)) * 4; }
// offset (in int32 elements) of tok_info (N) inside the bin_units buffer
extern "C" int64_t gdmae_sra_tok_info_offset(int64_t N) { return sra_units_ints(N); }

// bin_units (gdmae_sra_bin_units_bytes(N), 16-byte aligned) <- per-bin blocks of the CSR rows described by row_info (N,4),
// followed by tok_info (N)
extern "C" int gdmae_sra_bin_units(const int32_t* row_info, int64_t N, int32_t* bin_units, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 26) && ((uintptr_t)row_info % 16) == 0 && ((uintptr_t)bin_units % 16) == 0);
  if (N == 0) return GDMAE_OK;
  const long long nbins = (N + MM_BIN - 1) / MM_BIN;
  sra_bin_units_kernel<<<(unsigned)((nbins * 32 + 255) / 256), 256, 0, (cudaStream_t)stream_>>>((const int4*)row_info, (int)N, bin_units);
  GDMAE_LAUNCH_CHECK();
  sra_tok_info_kernel<<<gdmae_grid(N, 256), 256, 0, (cudaStream_t)stream_>>>((const int4*)row_info, (int)N, bin_units + sra_units_ints(N));
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

// ---- re-layout for callers that hold the flat (N, 3d) bf16 q | k | v (no biases, no positional term) of r1:
// one warp per CSR row adds the LUT row, normalises q and k per head (fp32), folds log2(e)/tau into q and writes the
// window-major tensors 0..2 + 1/|q|, 1/|k| into the row's record; optionally moves dO (N, d) into tensor 3 and lse (N, 8,
// token order) into the record.  In the fused encoder layer the GEMM epilogues do this (tc_gemm.cu modes 4 / 5).
// qkvdw[tensor][slice][row][64] bf16; lrr[row][24] fp32 = lse | 1/|q| | 1/|k|.
template <int D>
__global__ void __launch_bounds__(256) sra_prep_kernel(const bf16* __restrict__ qkv, const float* __restrict__ lut,
                                                       const int4* __restrict__ row_info, const float* __restrict__ tau, float tau_min,
                                                       int N, bf16* __restrict__ qkvdw, float* __restrict__ lrr,
                                                       const bf16* __restrict__ dout, const float* __restrict__ lse_tok) {
  constexpr int C = D / 32, NSL = D / 64;     // channels per lane: 4 lanes per head for both head sizes
  const int lane = threadIdx.x & 31;
  const int c0 = lane * C, sl = c0 >> 6, cw = c0 & 63;
  const float qscale = qkv ? 1.4426950408889634f / fmaxf(__ldg(tau), tau_min) : 0.f;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < N; row += (gridDim.x * blockDim.x) >> 5) {
    const int4 rec = __ldg(row_info + row);
    const long long tok = rec.x;
    if (qkv) {
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        float x[C];
        const bf16* src = qkv + tok * 3 * D + part * D + c0;
#pragma unroll
        for (int i = 0; i < C; ++i) x[i] = __bfloat162float(src[i]);
        if (part < 2) {
          const float* l = lut + (long long)rec.w * 2 * D + part * D + c0;
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < C; ++i) { x[i] += __ldg(l + i); ss = fmaf(x[i], x[i], ss); }
          ss += __shfl_xor_sync(0xffffffffu, ss, 1);
          ss += __shfl_xor_sync(0xffffffffu, ss, 2);
          const float rn = rsqrtf(fmaxf(ss, 1e-24f));
          if ((lane & 3) == 0) lrr[(long long)row * 24 + 8 + 8 * part + (lane >> 2)] = rn;
          const float f = part == 0 ? rn * qscale : rn;
#pragma unroll
          for (int i = 0; i < C; ++i) x[i] *= f;
        }
        bf16* dst = qkvdw + ((long long)(part * NSL + sl) * N + row) * 64 + cw;
#pragma unroll
        for (int i = 0; i < C; i += 2) *reinterpret_cast<unsigned*>(dst + i) = pack_bf16(x[i], x[i + 1]);
      }
    }
    if (dout) {
      const bf16* src = dout + tok * D + c0;
      bf16* dst = qkvdw + ((long long)(3 * NSL + sl) * N + row) * 64 + cw;
#pragma unroll
      for (int i = 0; i < C; i += 2) *reinterpret_cast<unsigned*>(dst + i) = *reinterpret_cast<const unsigned*>(src + i);
    }
    if (lse_tok && lane < 8) lrr[(long long)row * 24 + lane] = __ldg(lse_tok + tok * 8 + lane);
  }
}

// (64 channels, N rows, d/64 slices, `tensors`) bf16 array, box 64 x 16 x 1 x tensors, SWIZZLE_128B; rows past N read as zero
void* gdmae_tensor_map_encoder();   // tc_gemm.cu: cuTensorMapEncodeTiled from the driver, or nullptr
static int sra_make_map(CUtensorMap* m, const void* ptr, long long N, int nsl, int tensors) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = (EncodeTiledFn)gdmae_tensor_map_encoder();
  if (!enc) { gdmae_set_error("cuTensorMapEncodeTiled is not available from the driver"); return GDMAE_ERR_CUDA; }
  cuuint64_t dims[4] = {64, (cuuint64_t)N, (cuuint64_t)nsl, (cuuint64_t)tensors};
  cuuint64_t strides[3] = {128, (cuuint64_t)N * 128, (cuuint64_t)N * 128 * nsl};
  cuuint32_t box[4] = {64, MM_BOX, 1, (cuuint32_t)tensors};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[160];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled (sra) failed (%d): ptr %p N %lld slices %d tensors %d", (int)r, ptr, N, nsl, tensors);
    gdmae_set_error(b);
    return GDMAE_ERR_CUDA;
  }
  return GDMAE_OK;
}

static int mm_attrs() {
  static bool done = false;
  if (!done) {
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_mma_kernel<16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_BYTES));
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_mma_kernel<32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_BYTES));
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_mma_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_BYTES));
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_fwd_mma_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_BYTES));
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_bwd_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM_BYTES));
    GDMAE_CHECK_CUDA(cudaFuncSetAttribute(sra_bwd_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM_BYTES));
    done = true;
  }
  return GDMAE_OK;
}

// Forward on the window-major layout: qkvw[tensor][slice][row][64] bf16 (q^ * log2(e)/tau, k^, v; rows in CSR order, as
// tc_gemm mode 4 or the re-layout kernel writes them); out (N, d) token order, fp32 or bf16; lse: (N, 8) by token, or
// (lse_by_row) columns 0..7 of the (N, 24) per-row records the backward reads.
extern "C" int gdmae_sra_fwd_win(const void* qkvw, const int32_t* bin_units, int64_t N, int d, const float* bv, int out_bf16, void* out,
                                 float* lse, int lse_by_row, void* stream_) {
  GDMAE_CHECK_ARG(bin_units != nullptr && ((uintptr_t)bin_units % 16) == 0);
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 26) && (d == 128 || d == 256));
  GDMAE_CHECK_ARG(((uintptr_t)qkvw % 128) == 0 && (bv == nullptr || ((uintptr_t)bv % 8) == 0));
  if (N == 0) return GDMAE_OK;
  int rc = mm_attrs();
  if (rc) return rc;
  CUtensorMap tq;
  rc = sra_make_map(&tq, qkvw, N, d / MM_SLICE, 3);
  if (rc) return rc;
  MmArgs a{bin_units, bv, (int)N, d, out_bf16, lse_by_row};
  cudaStream_t st = (cudaStream_t)stream_;
  // one CTA per SM; 148 is a multiple of the 2 (d = 128) and 4 (d = 256) channel slices
  if (out_bf16 && lse_by_row) {
    if (d == 128) sra_fwd_mma_kernel<16, 1><<<GDMAE_NUM_SMS, MF_THREADS, MM_SMEM_BYTES, st>>>(tq, a, out, lse);
    else sra_fwd_mma_kernel<32, 1><<<GDMAE_NUM_SMS, MF_THREADS, MM_SMEM_BYTES, st>>>(tq, a, out, lse);
  } else {
    if (d == 128) sra_fwd_mma_kernel<16, 0><<<GDMAE_NUM_SMS, MF_THREADS, MM_SMEM_BYTES, st>>>(tq, a, out, lse);
    else sra_fwd_mma_kernel<32, 0><<<GDMAE_NUM_SMS, MF_THREADS, MM_SMEM_BYTES, st>>>(tq, a, out, lse);
  }
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// Backward on the window-major layout: qkvdw = the forward's three tensors + dO as the fourth (tc_gemm mode 5 / re-layout);
// lrr (N, 24) per-row records by CSR row: lse | 1/|q| | 1/|k|; dqkv (N, 3d) bf16 in token order; dtau_sum (1, double,
// caller zeroes) accumulates sum dS*S.
extern "C" int gdmae_sra_bwd_win(const void* qkvdw, const float* lrr, const int32_t* bin_units, int64_t N, int d, const float* tau,
                                 float tau_min, void* dqkv_bf16, double* dtau_sum, void* stream_) {
  GDMAE_CHECK_ARG(bin_units != nullptr && ((uintptr_t)bin_units % 16) == 0);
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 26) && (d == 128 || d == 256));
  GDMAE_CHECK_ARG(((uintptr_t)qkvdw % 128) == 0 && ((uintptr_t)lrr % 32) == 0 && ((uintptr_t)dqkv_bf16 % 16) == 0);
  if (N == 0) return GDMAE_OK;
  int rc = mm_attrs();
  if (rc) return rc;
  CUtensorMap tq;
  rc = sra_make_map(&tq, qkvdw, N, d / MM_SLICE, 4);
  if (rc) return rc;
  MbArgs a{bin_units, tau, lrr, (bf16*)dqkv_bf16, dtau_sum, tau_min, (int)N, d};
  cudaStream_t st = (cudaStream_t)stream_;
  if (d == 128) sra_bwd_mma_kernel<16><<<GDMAE_NUM_SMS, MM_THREADS, MB_SMEM_BYTES, st>>>(tq, a);
  else sra_bwd_mma_kernel<32><<<GDMAE_NUM_SMS, MM_THREADS, MB_SMEM_BYTES, st>>>(tq, a);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// workspace of the flat-layout entry points below: qkvdw (4, N, d) bf16 + lrr (N, 24) fp32
extern "C" size_t gdmae_sra_tc_workspace_bytes(int64_t N, int d) {
  const size_t n = (size_t)(N > 0 ? N : 1);
  return gdmae_align(n * 4 * d * 2) + gdmae_align(n * 24 * 4) + 1024;
}

struct SraWs {
  bf16* qkvdw;
  float* lrr;
};
static int sra_carve(void* workspace, size_t ws_bytes, int64_t N, int d, SraWs* w) {
  Workspace ws(workspace, ws_bytes);
  w->qkvdw = ws.take<bf16>(N * 4 * d);
  w->lrr = ws.take<float>(N * 24);
  if (!w->lrr || ((uintptr_t)w->qkvdw % 128) != 0) { gdmae_set_error("sra tc: workspace too small or not 256-byte aligned (gdmae_sra_tc_workspace_bytes)"); return GDMAE_ERR_WORKSPACE; }
  return GDMAE_OK;
}

// flat (N, 3d) bf16 q | k | v (nullable) -> tensors 0..2 of qkvdw + 1/|q|, 1/|k| in lrr; dout (N, d, nullable) -> tensor 3;
// lse_tok (N, 8, nullable) -> lrr.  The stand-alone form of what the fused layer's GEMM epilogues do.
extern "C" int gdmae_sra_relayout(const void* qkv_bf16, const float* lut, const int32_t* row_info, const float* tau, float tau_min,
                                  int64_t N, int d, const void* dout_bf16, const float* lse_tok, void* qkvdw, float* lrr, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 26) && (d == 128 || d == 256) && ((uintptr_t)row_info % 16) == 0 && qkvdw && lrr);
  GDMAE_CHECK_ARG(!qkv_bf16 || (lut && tau));
  if (N == 0) return GDMAE_OK;
  cudaStream_t st = (cudaStream_t)stream_;
  if (d == 128)
    sra_prep_kernel<128><<<gdmae_grid(N * 32, 256), 256, 0, st>>>((const bf16*)qkv_bf16, lut, (const int4*)row_info, tau, tau_min, (int)N, (bf16*)qkvdw,
                                                                  lrr, (const bf16*)dout_bf16, lse_tok);
  else
    sra_prep_kernel<256><<<gdmae_grid(N * 32, 256), 256, 0, st>>>((const bf16*)qkv_bf16, lut, (const int4*)row_info, tau, tau_min, (int)N, (bf16*)qkvdw,
                                                                  lrr, (const bf16*)dout_bf16, lse_tok);
  GDMAE_LAUNCH_CHECK();
  return GDMAE_OK;
}

// Tensor-core variant of gdmae_sra_attention_fwd for the flat bf16 q/k/v: qkv (N, 3d) bf16, otherwise the same arguments
// and outputs (lse by token); re-layout + window-major kernel.  workspace: gdmae_sra_tc_workspace_bytes(N, d).
extern "C" int gdmae_sra_attention_fwd_tc(const void* qkv_bf16, const float* lut, const int32_t* row_info, const int32_t* bin_units,
                                          int64_t N, int d, int nhead, const float* tau, float tau_min, const float* bv, int io_bf16,
                                          void* out, float* lse, void* workspace, size_t ws_bytes, void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 26) && nhead == 8 && (d == 128 || d == 256));
  GDMAE_CHECK_ARG(((uintptr_t)qkv_bf16 % 16) == 0 && ((uintptr_t)lut % 8) == 0);
  if (N == 0) return GDMAE_OK;
  SraWs w;
  int rc = sra_carve(workspace, ws_bytes, N, d, &w);
  if (rc) return rc;
  rc = gdmae_sra_relayout(qkv_bf16, lut, row_info, tau, tau_min, N, d, nullptr, nullptr, w.qkvdw, w.lrr, stream_);
  if (rc) return rc;
  return gdmae_sra_fwd_win(w.qkvdw, bin_units, N, d, bv, io_bf16, out, lse, 0, stream_);
}

// Tensor-core backward for the flat bf16 tensors: qkv (N,3d), dout (N,d) and dqkv (N,3d) are bf16; lse (N,8) by token from
// the forward; dtau_sum (1, double, caller zeroes) accumulates sum dS*S as gdmae_sra_attention_bwd does.  The value bias
// and the forward output are not needed (sum_j P dP replaces dO.(o - bv)).
extern "C" int gdmae_sra_attention_bwd_tc(const void* qkv_bf16, const float* lut, const int32_t* row_info, const int32_t* bin_units,
                                          int64_t N, int d, int nhead, const float* tau, float tau_min, const float* lse,
                                          const void* dout_bf16, void* dqkv_bf16, double* dtau_sum, void* workspace, size_t ws_bytes,
                                          void* stream_) {
  GDMAE_CHECK_ARG(N >= 0 && N < (1ll << 26) && nhead == 8 && (d == 128 || d == 256));
  GDMAE_CHECK_ARG(((uintptr_t)qkv_bf16 % 16) == 0 && ((uintptr_t)lut % 8) == 0 && ((uintptr_t)dout_bf16 % 16) == 0 && ((uintptr_t)lse % 4) == 0);
  if (N == 0) return GDMAE_OK;
  SraWs w;
  int rc = sra_carve(workspace, ws_bytes, N, d, &w);
  if (rc) return rc;
  rc = gdmae_sra_relayout(qkv_bf16, lut, row_info, tau, tau_min, N, d, dout_bf16, lse, w.qkvdw, w.lrr, stream_);
  if (rc) return rc;
  return gdmae_sra_bwd_win(w.qkvdw, w.lrr, bin_units, N, d, tau, tau_min, dqkv_bf16, dtau_sum, stream_);
}
