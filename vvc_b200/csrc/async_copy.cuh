// async_copy.cuh -- asynchronous global -> shared copies of sm_100a used by the persistent tile kernels:
// TMA 2-D boxes (cp.async.bulk.tensor, SASS UTMALDG) completing on an mbarrier, and cp.async (LDGSTS) for small records.
#pragma once

#include <cuda.h>
#include <cstdint>

namespace vtmgpu
{

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cpAsync16(void* smem, const void* gmem)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smemAddr(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpAsync8(void* smem, const void* gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smemAddr(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpAsync4(void* smem, const void* gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smemAddr(smem)), "l"(gmem) : "memory");
}
template <int N> __device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mbarInit(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "MBAR_WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra MBAR_DONE_%=;\n"
    "bra MBAR_WAIT_%=;\n"
    "MBAR_DONE_%=:\n"
    "}\n" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
// plain arrival of one thread (release): the split CTA barriers of k_alf -- every warp arrives when its part is done and waits
// (mbarWait) only where it needs the others' results
__device__ __forceinline__ void mbarArrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory"); }
// one 2-D box of a plane (element coordinates, may start outside: those elements arrive as zeros) into shared memory
__device__ __forceinline__ void tmaLoad2D(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smemAddr(dst)), "l"(map), "r"(x), "r"(y), "r"(smemAddr(bar)) : "memory");
}

// contiguous bytes (multiple of 16, both addresses 16-byte aligned) into shared memory, completing on the same kind of mbarrier (SASS UBLKCP)
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}

// position of a CTA in its round-robin walk over the work items of a batch of picture slots, and one step of the walk
// (= gridDim.x items) decomposed on the host so that the loop needs no division
struct TileStep
{
  int dItem, dSlot;      // gridDim.x = dSlot * itemsPerSlot + dItem
};

}   // namespace vtmgpu
