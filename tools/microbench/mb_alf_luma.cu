// mb_alf_luma.cu -- the 7x7 luma block filter of k_alf in isolation (sm_100a): round-1 routine on the row-major tile
// (alfLumaBlockFast, funnel shifts for odd offsets) against the vertical-pair routine (alfLumaBlockV), same samples, same
// filter table, launch shape of the real kernel (256 threads, 2 CTAs per SM, one 64x64 tile per CTA and iteration).
// Prints clocks per tile and CTA, the equivalent microseconds per 3840x2160 picture, and whether both routines agree.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../include -I../../vvc_b200/csrc mb_alf_luma.cu -o mb_alf_luma
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "alf_kernel.cuh"

using namespace vtmgpu;

#define ITER 64

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// VAR bit 0: no CTA barrier between iterations ; bit 1: the filter entry is loaded once, before the loop
template <int MODE, int VAR = 0> __global__ void __launch_bounds__(256, 2) k(const AlfLumaEntry* __restrict__ tab, pel* __restrict__ out, long long* cyc)
{
  extern __shared__ __align__(128) unsigned char sm[];
  pel* H = reinterpret_cast<pel*>(sm);                                       // [SA_H][SA_P], tile origin at (SA_HY, SA_HX)
  uint32_t* V = reinterpret_cast<uint32_t*>(sm + SA_H * SA_P * 2);           // [AV_ROWS][AV_COLS]
  const int tid = threadIdx.x, bi = tid >> 4, bj = tid & 15;
  for (int i = tid; i < SA_H * SA_P; i += 256) H[i] = (pel)(hash32(i * 2654435761u + blockIdx.x) & 1023);
  __syncthreads();
  for (int i = tid; i < AV_ROWS * AV_COLS; i += 256)
  {
    const int rr = i / AV_COLS, lx = i - rr * AV_COLS;                       // sample rows (rr - 4, rr - 3) of the tile, column lx - 4
    const uint32_t a = (uint16_t)H[rr * SA_P + lx + SA_HX - 4], b = (uint16_t)H[(rr + 1) * SA_P + lx + SA_HX - 4];
    V[i] = a | b << 16;
  }
  __syncthreads();
  pel* o = out + (size_t)blockIdx.x * 64 * 64 + (4 * bi) * 64 + 4 * bj;
  const pel* c0 = &H[(4 * bi + SA_HY) * SA_P + 4 * bj + SA_HX];
  const uint32_t* v = V + (4 * bi + 4) * AV_COLS + 4 * bj;
  const uint32_t maxvP = dup16(1023);
  const LumaCoef K0 = loadLumaCoef(tab + hash32(tid * 977 + blockIdx.x) % 100);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++)
  {
    const AlfLumaEntry* e = tab + hash32(tid * 977 + it * 131 + blockIdx.x) % 100;
    if (MODE == 0) alfLumaBlockFast(c0, o, 64, e, maxvP, 0);
    else if (VAR & 2) alfLumaBlockV<MODE - 1>(v, o + (it & 1) * 4096, 64, K0, maxvP);
    else           alfLumaBlockV<MODE - 1>(v, o, 64, loadLumaCoef(e), maxvP);
    if (!(VAR & 1)) __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount, grid = 2 * nsm;
  std::vector<AlfLumaEntry> tab(100);
  srand(5);
  const int clips[4] = { 1024, 181, 32, 6 };
  for (auto& e : tab)
  {
    int bias = 64;
    for (int k = 0; k < 12; k++)
    {
      const int co = rand() % 41 - 20, cl = clips[rand() & 3];
      e.coefB[k] = (uint32_t)(co & 0xff) * 0x01000001u;
      e.clipP1[k] = (uint32_t)((cl + 1) & 0xffff) * 0x10001u;
      e.clip2[k] = (uint32_t)((2 * cl) & 0xffff) * 0x10001u;
      bias -= co * 2 * cl;
    }
    e.bias = bias;
  }
  AlfLumaEntry* dtab; pel* dout[2]; long long* dcyc;
  cudaMalloc(&dtab, sizeof(AlfLumaEntry) * 100);
  cudaMemcpy(dtab, tab.data(), sizeof(AlfLumaEntry) * 100, cudaMemcpyHostToDevice);
  for (auto& d : dout) { cudaMalloc(&d, (size_t)(grid + 1) * 4096 * 2); cudaMemset(d, 0, (size_t)(grid + 1) * 4096 * 2); }
  cudaMalloc(&dcyc, sizeof(long long) * grid);
  const int smem = SA_H * SA_P * 2 + AV_BYTES;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<long long> h(grid);
  const char* names[5] = { "alfLumaBlockFast (row-major, round 1)", "alfLumaBlockV (vertical pairs), adds by ptxas", "alfLumaBlockV, tap sum on the ALU pipe",
                           "alfLumaBlockV, clip - cur on the ALU pipe", "alfLumaBlockV, both adds on the ALU pipe" };
  for (int mode = 0; mode < 5; mode++)
  {
    for (int rep = 0; rep < 2; rep++)
    {
      if (mode == 0) k<0><<<grid, 256, smem>>>(dtab, dout[0], dcyc);
      if (mode == 1) k<1><<<grid, 256, smem>>>(dtab, dout[1], dcyc);
      if (mode == 2) k<2><<<grid, 256, smem>>>(dtab, dout[1], dcyc);
      if (mode == 3) k<3><<<grid, 256, smem>>>(dtab, dout[1], dcyc);
      if (mode == 4) k<4><<<grid, 256, smem>>>(dtab, dout[1], dcyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h.data(), dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (long long c : h) avg += (double)c;
    avg /= grid * ITER;
    // one 4K picture = 60 x 34 tiles over 2 * nsm resident CTAs
    printf("%-48s %8.0f clocks per tile and CTA  -> %6.2f us per 3840x2160 picture at %.3f GHz (luma 7x7 alone)\n",
           names[mode], avg, avg * (60.0 * 34.0 / grid) / (p.clockRate * 1e-3), p.clockRate * 1e-6);
  }
  {
    auto runv = [&](const char* name, auto kern) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      for (int rep = 0; rep < 2; rep++) { kern<<<grid, 256, smem>>>(dtab, dout[1], dcyc); cudaDeviceSynchronize(); }
      cudaMemcpy(h.data(), dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
      double avg = 0;
      for (long long c : h) avg += (double)c;
      avg /= grid * ITER;
      printf("%-48s %8.0f clocks per tile and CTA  -> %6.2f us per 3840x2160 picture\n", name, avg, avg * (60.0 * 34.0 / grid) / (p.clockRate * 1e-3));
    };
    runv("alfLumaBlockV<1>, no barrier", k<2, 1>);
    runv("alfLumaBlockV<1>, entry loaded once", k<2, 2>);
    runv("alfLumaBlockV<1>, no barrier, entry loaded once", k<2, 3>);
    k<2, 0><<<grid, 256, smem>>>(dtab, dout[1], dcyc);      // leaves the reference result of the comparison below in dout[1]
    cudaDeviceSynchronize();
  }
  std::vector<pel> a((size_t)grid * 4096), b((size_t)grid * 4096);
  cudaMemcpy(a.data(), dout[0], a.size() * 2, cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), dout[1], b.size() * 2, cudaMemcpyDeviceToHost);
  size_t diff = 0, nz = 0;
  for (size_t i = 0; i < a.size(); i++) { diff += a[i] != b[i]; nz += a[i] != 0; }
  printf("outputs: %zu samples, %zu non-zero, %zu differ -> %s\n", a.size(), nz, diff, diff == 0 && nz > 0 ? "IDENTICAL" : "MISMATCH");
  return diff != 0;
}
