// packed16.cuh -- two 16-bit samples per 32-bit register: thin wrappers over the native packed integer instructions of
// sm_100a (VIADD.16x2, VIADDMNMX.S16x2.RELU, PRMT, SHF) used by all filter kernels of libvtmgpu.
#pragma once

#include "vtmgpu_dev.cuh"

namespace vtmgpu
{

// ---- packed 16x2 helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
  return d;
}
__device__ __forceinline__ uint32_t dup16(int v) { return (uint32_t)(v & 0xffff) * 0x10001u; }
// (hi lane of lo, lo lane of hi): the pair that starts one sample after lo
__device__ __forceinline__ uint32_t mid16(uint32_t lo, uint32_t hi) { return __funnelshift_r(lo, hi, 16); }
// max(min(a + b, c), 0) per signed 16-bit lane
__device__ __forceinline__ uint32_t addClamp0(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2_relu(a, b, c); }

}   // namespace vtmgpu
