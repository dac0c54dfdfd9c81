// Bulk-copy pipeline primitives for the row-streaming kernels: contiguous row ranges travel global -> shared (and back) as
// cp.async.bulk transactions of the TMA engine, completion is counted by mbarriers, so the bytes in flight per SM are set by
// the stage ring in shared memory and not by registers or resident warps (sm_100a; see sra_attention_tc.cu for the tensor-map
// form).  Every wait is bounded and traps: a lost signal ends in a sticky CUDA error, never in a hang or in stale data.
#pragma once
#include "common.cuh"

namespace bp {
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  for (int spin = 0; spin < 50000; ++spin) {          // ~1 s with the 20 us suspend hint
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy accesses to shared memory before, async-proxy (bulk copy) accesses after
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// global -> shared, bytes % 16 == 0, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global (bulk group of the issuing thread)
__device__ __forceinline__ void s2g(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s2g_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of the thread's committed store groups may still be READING shared memory
template <int N> __device__ __forceinline__ void s2g_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void s2g_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
}  // namespace bp
