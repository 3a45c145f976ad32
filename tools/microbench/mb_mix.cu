// mb_mix.cu -- which instruction mixes of the ALF inner loop can a B200 SM sub-partition issue at more than the 0.68 IPC the
// 4 x VIADDMNMX + 4 x IDP.2A loop of mb_packed.cu reaches?  Every kernel runs ITER iterations of a loop body of independent
// dependency chains (one instruction per chain and iteration); 1024 threads per SM (8 warps per scheduler) unless stated.
// Output: warp-instructions per clock and scheduler (IPC) counting only the named instructions.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048

__device__ __forceinline__ unsigned A3(unsigned x, unsigned a, unsigned b) { return __viaddmin_s16x2_relu(x, a, b); }     // ALU, 3 register sources
__device__ __forceinline__ unsigned A2(unsigned x, unsigned a) { return __vadd2(x, a); }                                   // ALU, 2 register sources
__device__ __forceinline__ unsigned F3(unsigned x, unsigned a, unsigned b) { return (unsigned)__dp2a_lo((int)a, (int)b, (int)x); }   // FMA-heavy, 3 sources
__device__ __forceinline__ unsigned F2(unsigned x, unsigned a) { return x * 3u + a; }                                      // IMAD with immediate
__device__ __forceinline__ float    L3(float x, float a, float b) { return fmaf(x, a, b); }                                // FFMA (either FMA pipe)

template <int MODE> __global__ void __launch_bounds__(1024) k(unsigned* out, long long* cyc, unsigned a, unsigned b, float fa, float fb)
{
  unsigned v[12], w[16];
  float f[4];
#pragma unroll
  for (int i = 0; i < 12; i++) v[i] = threadIdx.x * 12 + i + a;
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = threadIdx.x * 7 + i * a + b;
#pragma unroll
  for (int i = 0; i < 4; i++) f[i] = (float)(threadIdx.x + i) * fa;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++)
  {
    if (MODE == 0) { // AAAA FFFF (3-source both) = mb_packed mix
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = A3(v[i], a, b);
#pragma unroll
      for (int i = 4; i < 8; i++) v[i] = F3(v[i], a, b);
    }
    if (MODE == 1) { // interleaved A F A F ...
#pragma unroll
      for (int i = 0; i < 4; i++) { v[i] = A3(v[i], a, b); v[4 + i] = F3(v[4 + i], a, b); }
    }
    if (MODE == 2) { // 2-source ALU + 3-source FMA
#pragma unroll
      for (int i = 0; i < 4; i++) { v[i] = A2(v[i], a); v[4 + i] = F3(v[4 + i], a, b); }
    }
    if (MODE == 3) { // 2-source ALU + immediate IMAD
#pragma unroll
      for (int i = 0; i < 4; i++) { v[i] = A2(v[i], a); v[4 + i] = F2(v[4 + i], a); }
    }
    if (MODE == 4) { // 4 ALU + 4 FMA + 4 FFMA
#pragma unroll
      for (int i = 0; i < 4; i++) { v[i] = A3(v[i], a, b); v[4 + i] = F3(v[4 + i], a, b); f[i] = L3(f[i], fa, fb); }
    }
    if (MODE == 5) { // FFMA alone
#pragma unroll
      for (int i = 0; i < 4; i++) { f[i] = L3(f[i], fa, fb); }
#pragma unroll
      for (int i = 0; i < 4; i++) { f[i] = L3(f[i], fb, fa); }
    }
    if (MODE == 6) { // 4 IDP + 4 FFMA
#pragma unroll
      for (int i = 0; i < 4; i++) { v[4 + i] = F3(v[4 + i], a, b); f[i] = L3(f[i], fa, fb); }
    }
    if (MODE == 7) { // 4 ALU + 4 FFMA
#pragma unroll
      for (int i = 0; i < 4; i++) { v[i] = A3(v[i], a, b); f[i] = L3(f[i], fa, fb); }
    }
    if (MODE == 8) { // the ALF tap: cb = A2 ; two clamps A3 ; add (left to ptxas) ; two IDP -- 2 independent taps
#pragma unroll
      for (int i = 0; i < 2; i++)
      {
        const unsigned cb = A2(v[8 + i], a);
        const unsigned s = A3(v[i], cb, b) + A3(v[2 + i], cb, b);
        v[4 + i] = F3(v[4 + i], s, b);
        v[6 + i] = (unsigned)__dp2a_hi((int)s, (int)b, (int)v[6 + i]);
      }
    }
    if (MODE == 11) { // AFAF with DISTINCT register operands in every instruction (register-file read bandwidth)
#pragma unroll
      for (int i = 0; i < 4; i++) { v[i] = A3(v[i], w[i], w[4 + i]); v[4 + i] = F3(v[4 + i], w[8 + i], w[12 + i]); }
    }
    if (MODE == 12) { // same, operands shared pairwise (what .reuse can catch)
#pragma unroll
      for (int i = 0; i < 4; i++) { v[i] = A3(v[i], w[i & 1], w[4]); v[4 + i] = F3(v[4 + i], w[8 + (i & 1)], w[12]); }
    }
    if (MODE == 13) { // tap stream with real dependencies: 4 columns of one tap, data rotating through registers
#pragma unroll
      for (int i = 0; i < 4; i++)
      {
        const unsigned cb = A2(w[i], w[4]);
        const unsigned s = __viaddmin_s16x2(A3(w[8 + i], cb, w[5]), A3(w[12 + i], cb, w[5]), 0x7fff7fffu);
        v[i] = F3(v[i], s, w[6]);
        v[4 + i] = (unsigned)__dp2a_hi((int)s, (int)w[6], (int)v[4 + i]);
      }
#pragma unroll
      for (int i = 0; i < 16; i++) w[i] ^= v[i & 7] >> 31;      // keeps w loop-variant at negligible cost? (counted below)
    }
    if (MODE == 9) { // 6 ALU : 2 FMA
#pragma unroll
      for (int i = 0; i < 6; i++) v[i] = A3(v[i], a, b);
#pragma unroll
      for (int i = 6; i < 8; i++) v[i] = F3(v[i], a, b);
    }
    if (MODE == 10) { // 2 ALU : 6 FMA
#pragma unroll
      for (int i = 0; i < 2; i++) v[i] = A3(v[i], a, b);
#pragma unroll
      for (int i = 2; i < 8; i++) v[i] = F3(v[i], a, b);
    }
  }
  const long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) s ^= v[i];
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= w[i];
#pragma unroll
  for (int i = 0; i < 4; i++) s ^= __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int instrPerIter, int threads, int nsm, unsigned* out, long long* cyc)
{
  for (int rep = 0; rep < 2; rep++) { k<MODE><<<nsm, threads>>>(out, cyc, 0x00030005u, 0x00070009u, 1.0001f, 0.5f); cudaDeviceSynchronize(); }
  long long h[256];
  cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < nsm; i++) avg += (double)h[i];
  avg /= nsm;
  printf("%-52s %4d thr/SM  IPC per scheduler %.3f\n", name, threads, (double)instrPerIter * ITER * (threads / 32) / 4.0 / avg);
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  unsigned* out; long long* cyc;
  cudaMalloc(&out, sizeof(unsigned) * 1024 * nsm);
  cudaMalloc(&cyc, sizeof(long long) * nsm);
  for (int threads : { 1024, 512 })
  {
    run<0>("AAAA FFFF  (VIADDMNMX / IDP.2A, 3 sources)", 8, threads, nsm, out, cyc);
    run<1>("AFAFAFAF   (same, interleaved)", 8, threads, nsm, out, cyc);
    run<2>("VIADD.16x2 (2 src) + IDP.2A", 8, threads, nsm, out, cyc);
    run<3>("VIADD.16x2 (2 src) + IMAD imm", 8, threads, nsm, out, cyc);
    run<4>("4 VIADDMNMX + 4 IDP + 4 FFMA", 12, threads, nsm, out, cyc);
    run<5>("8 FFMA", 8, threads, nsm, out, cyc);
    run<6>("4 IDP + 4 FFMA", 8, threads, nsm, out, cyc);
    run<7>("4 VIADDMNMX + 4 FFMA", 8, threads, nsm, out, cyc);
    run<8>("ALF tap x2 (VIADD, 2 VIADDMNMX, add, 2 IDP)", 12, threads, nsm, out, cyc);
    run<9>("6 VIADDMNMX + 2 IDP", 8, threads, nsm, out, cyc);
    run<10>("2 VIADDMNMX + 6 IDP", 8, threads, nsm, out, cyc);
    run<11>("AFAF, distinct register operands", 8, threads, nsm, out, cyc);
    run<12>("AFAF, operands shared pairwise", 8, threads, nsm, out, cyc);
    run<13>("tap x 4 columns (24 + 32 helper LOP/SHF)", 56, threads, nsm, out, cyc);
  }
  return 0;
}
