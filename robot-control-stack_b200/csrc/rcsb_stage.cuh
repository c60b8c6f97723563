// Model staging into shared memory with a TMA bulk copy, shared by every kernel variant.
#pragma once
#include <stddef.h>
// ------------------------------------------------------------------ TMA bulk copy helpers (sm_90+/sm_100a PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!ok);
}

#define RCSB_MODEL_BYTES ((sizeof(RcsbModel) + 15) & ~(size_t)15)                 // global-memory copy
#define RCSB_MODEL_HOT_BYTES ((offsetof(RcsbModel, cold_begin) + 15) & ~(size_t)15)  // part staged into shared memory
#define RCSB_SMEM_HEADER (RCSB_MODEL_HOT_BYTES + 16)
// warps per CTA are bounded by the per-warp shared-memory workspace (about 22 KB for the FR3 scenes),
// so the register budget per thread can be generous
#ifndef RCSB_MAX_WARPS
#define RCSB_MAX_WARPS 28
#endif

__device__ __forceinline__ const RcsbModel* stage_model(const RcsbModel* gm) {
  RcsbModel* sm = (RcsbModel*)rcsb_smem;
  uint64_t* bar = (uint64_t*)(rcsb_smem + RCSB_MODEL_HOT_BYTES);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (uint32_t)RCSB_MODEL_HOT_BYTES);
    tma_bulk_g2s(sm, gm, (uint32_t)RCSB_MODEL_HOT_BYTES, bar);
  }
  mbar_wait(bar, 0);
  return sm;
}


