// mb_occ.cu -- how the issue rate of the ALF tap stream (clip - cur, two VIADDMNMX clamps, sum, two IDP.2A; rows by LDS.128 from
// the vertical-pair copy) scales with the number of resident warps.  A 6-tap version of the block filter needs few enough
// registers for 4 CTAs of 256 threads per SM; occupancy is set through the dynamic shared memory size.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "alf_fast.cuh"
using namespace vtmgpu;
#define ITER 64
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void __launch_bounds__(256, 4) k(const AlfLumaEntry* __restrict__ tab, pel* __restrict__ out, long long* cyc, int* instr)
{
  extern __shared__ __align__(128) unsigned char sm[];
  uint32_t* V = reinterpret_cast<uint32_t*>(sm);
  const int tid = threadIdx.x, bi = tid >> 4, bj = tid & 15;
  for (int i = tid; i < AV_ROWS * AV_COLS; i += 256) V[i] = (hash32(i * 2654435761u + blockIdx.x) & 1023) | (hash32(i * 40503u + blockIdx.x) & 1023) << 16;
  __syncthreads();
  pel* o = out + (size_t)blockIdx.x * 64 * 64 + (4 * bi) * 64 + 4 * bj;
  const uint32_t* v = V + (4 * bi + 4) * AV_COLS + 4 * bj;
  const uint32_t maxvP = dup16(1023);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++)
  {
    const AlfLumaEntry* e = tab + hash32(tid * 977 + it * 131 + blockIdx.x) % 100;
    uint32_t coefB[6], clipP1[6], clip2[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { coefB[i] = __ldg(&e->coefB[i]); clipP1[i] = __ldg(&e->clipP1[i]); clip2[i] = __ldg(&e->clip2[i]); }
    const int bias = __ldg(&e->bias);
#pragma unroll 1
    for (int half = 0; half < 4; half++)          // 4 x (6 taps x 4 columns) = the tap count of one 4x4 block
    {
      const uint32_t* v0 = v + (half & 1) * 2 * AV_COLS;
      uint32_t C[12], P[12], M[12], ncur[4];
      int acc0[4], acc1[4];
#define LD12(W, PTR) { const uint4* p_ = reinterpret_cast<const uint4*>(PTR); const uint4 a_ = p_[0], b_ = p_[1], c_ = p_[2]; \
      W[0] = a_.x; W[1] = a_.y; W[2] = a_.z; W[3] = a_.w; W[4] = b_.x; W[5] = b_.y; W[6] = b_.z; W[7] = b_.w; W[8] = c_.x; W[9] = c_.y; W[10] = c_.z; W[11] = c_.w; }
      LD12(C, v0)
#pragma unroll
      for (int c = 0; c < 4; c++) { ncur[c] = ~C[4 + c]; acc0[c] = bias; acc1[c] = bias; }
#define TAP(K, DX, PP, MM) _Pragma("unroll") for (int c = 0; c < 4; c++) { \
        const uint32_t cb = __vadd2(clipP1[K], ncur[c]); \
        const uint32_t s = addAlu(addClamp0(PP[4 + c + (DX)], cb, clip2[K]), addClamp0(MM[4 + c - (DX)], cb, clip2[K])); \
        acc0[c] = __dp2a_lo((int)s, (int)coefB[K], acc0[c]); acc1[c] = __dp2a_hi((int)s, (int)coefB[K], acc1[c]); }
      TAP(0, 3, C, C) TAP(1, 1, C, C)
      LD12(P, v0 + AV_COLS) LD12(M, v0 - AV_COLS)
      TAP(2, 2, P, M) TAP(3, 0, P, M) TAP(4, -1, P, M) TAP(5, -2, P, M)
      uint32_t res[4];
#pragma unroll
      for (int c = 0; c < 4; c++) res[c] = addClamp0(C[4 + c], prmt((uint32_t)(acc0[c] >> 7), (uint32_t)(acc1[c] >> 7), 0x5410u), maxvP);
      pel* oo = out + (size_t)(half & 3) * 64;
      *reinterpret_cast<uint2*>(o + (half & 3) * 64) = make_uint2(prmt(res[0], res[1], 0x5410u) ^ prmt(res[0], res[1], 0x7632u), prmt(res[2], res[3], 0x5410u) ^ prmt(res[2], res[3], 0x7632u));
      (void)oo;
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  std::vector<AlfLumaEntry> tab(100);
  srand(5);
  const int clips[4] = { 1024, 181, 32, 6 };
  for (auto& e : tab) { int bias = 64; for (int k = 0; k < 12; k++) { const int co = rand() % 41 - 20, cl = clips[rand() & 3];
      e.coefB[k] = (uint32_t)(co & 0xff) * 0x01000001u; e.clipP1[k] = (uint32_t)((cl + 1) & 0xffff) * 0x10001u; e.clip2[k] = (uint32_t)((2 * cl) & 0xffff) * 0x10001u; bias -= co * 2 * cl; } e.bias = bias; }
  AlfLumaEntry* dtab; pel* dout; long long* dcyc;
  cudaMalloc(&dtab, sizeof(AlfLumaEntry) * 100);
  cudaMemcpy(dtab, tab.data(), sizeof(AlfLumaEntry) * 100, cudaMemcpyHostToDevice);
  cudaMalloc(&dout, (size_t)nsm * 8 * 4096 * 2);
  cudaMalloc(&dcyc, sizeof(long long) * nsm * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
  printf("registers per thread: %d\n", fa.numRegs);
  const int smems[5] = { 200 * 1024, 110 * 1024, 74 * 1024, 55 * 1024, 24 * 1024 };
  for (int s = 0; s < 5; s++)
  {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, smems[s]);
    const int grid = occ * nsm;
    for (int rep = 0; rep < 2; rep++) { k<<<grid, 256, smems[s]>>>(dtab, dout, dcyc, nullptr); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0; for (long long c : h) avg += (double)c; avg /= grid * ITER;
    // per iteration and thread: 4 x 24 tap-columns x 6 instructions
    printf("%d CTAs/SM (%2d warps/scheduler): %7.0f clocks per tile-iteration and CTA, %6.1f clocks per SM and tile, tap instructions IPC %.3f\n",
           occ, occ * 2, avg, avg / occ, 4.0 * 24 * 6 * 8 / 4 * occ / avg);
  }
  return 0;
}
