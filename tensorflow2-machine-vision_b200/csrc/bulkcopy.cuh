// bulkcopy.cuh — mbarrier + 1-D bulk asynchronous copy (cp.async.bulk, SASS UBLKCP) helpers shared by the streaming kernels.
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t bc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bc_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bc_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "BC_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra BC_DONE;\n"
      "bra BC_WAIT_LOOP;\n"
      "BC_DONE:\n"
      "}\n" ::"r"(bc_smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion counted on `bar`
__device__ __forceinline__ void bc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bc_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(bc_smem_u32(bar))
               : "memory");
}
// orders earlier generic-proxy accesses of shared memory before later async-proxy (bulk copy) writes to it
__device__ __forceinline__ void bc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
