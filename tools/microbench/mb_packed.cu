// mb_packed.cu -- issue-rate microbenchmark of the packed 16-bit integer instructions the ALF/SAO/DBF kernels
// are built from (sm_100a): VIADD.16x2, VIMNMX.S16x2, VIADDMNMX.S16x2(.RELU), IDP.2A, PRMT, IMAD, IADD3, LDS.
// One CTA of 1024 threads per SM; every thread runs ITER iterations of 8 independent dependency chains of one
// instruction; cycles by clock64() inside the kernel.  Output: lane-ops per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

template <int OP> __device__ __forceinline__ unsigned op(unsigned x, unsigned a, unsigned b)
{
  if (OP == 0) return __vadd2(x, a);                         // VIADD.16x2
  if (OP == 1) return __vmins2(x, a);                        // VIMNMX.S16x2
  if (OP == 2) return __viaddmax_s16x2(x, a, b);             // VIADDMNMX.S16x2
  if (OP == 3) return __viaddmin_s16x2_relu(x, a, b);        // VIADDMNMX.S16x2.RELU
  if (OP == 4) return (unsigned)__dp2a_lo((int)a, (int)b, (int)x);   // IDP.2A
  if (OP == 5) return __byte_perm(x, a, 0x5432);             // PRMT
  if (OP == 6) return x * a + b;                             // IMAD
  if (OP == 7) return x + a;                                 // IADD3 / VIADD
  if (OP == 8) return min(max((int)x, (int)a), (int)b);      // 2 x VIMNMX
  if (OP == 9) return __viaddmax_s32((int)x, (int)a, (int)b);// VIADDMNMX
  if (OP == 10) return __vimax3_s16x2(x, a, b);              // VIMNMX3.S16x2
  if (OP == 11) return __funnelshift_r(x, a, 16);            // SHF.R.W
  if (OP == 12) return (unsigned)__dp2a_lo((int)x, (int)a, (int)b);  // IDP.2A with the chain through the multiplicand
  return x;
}

template <int OP> __global__ void __launch_bounds__(1024) k(unsigned* out, long long* cyc, unsigned a, unsigned b)
{
  unsigned v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 8 + i + a;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++)
  {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = op<OP>(v[i], a, b);
  }
  const long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mixed ALU + FMA pipe: VIADDMNMX chain interleaved with IDP.2A chain (the ALF inner loop mix)
__global__ void __launch_bounds__(1024) kmix(unsigned* out, long long* cyc, unsigned a, unsigned b)
{
  unsigned v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 8 + i + a;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++)
  {
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = __viaddmin_s16x2_relu(v[i], a, b);
#pragma unroll
    for (int i = 4; i < 8; i++) v[i] = (unsigned)__dp2a_lo((int)a, (int)b, (int)v[i]);
  }
  const long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared-memory load rate: LDS.32 / LDS.64 / LDS.128, conflict-free
template <int W> __global__ void __launch_bounds__(1024) klds(unsigned* out, long long* cyc, unsigned a)
{
  __shared__ __align__(16) unsigned sm[1024 * 4 + 64];
  for (int i = threadIdx.x; i < 1024 * 4 + 64; i += 1024) sm[i] = i * a;
  __syncthreads();
  unsigned s = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER / 4; it++)
  {
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
      const int idx = ((threadIdx.x + i * 32 + it) & 1023) * W;
      if (W == 1) s ^= sm[idx];
      if (W == 2) { uint2 q = *reinterpret_cast<uint2*>(&sm[idx]); s ^= q.x ^ q.y; }
      if (W == 4) { uint4 q = *reinterpret_cast<uint4*>(&sm[idx]); s ^= q.x ^ q.y ^ q.z ^ q.w; }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class F> void run(const char* name, F launch, double opsPerThread, int nsm, unsigned* out, long long* cyc)
{
  launch();
  cudaDeviceSynchronize();
  launch();
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < nsm; i++) avg += (double)h[i];
  avg /= nsm;
  printf("%-34s %8.1f lane-ops/clk/SM   (%.0f cycles) %s\n", name, 1024.0 * opsPerThread / avg, avg, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  printf("device %s, %d SMs\n", p.name, nsm);
  unsigned* out; long long* cyc;
  cudaMalloc(&out, sizeof(unsigned) * 1024 * nsm);
  cudaMalloc(&cyc, sizeof(long long) * nsm);
  const double n = 8.0 * ITER;
#define R(OP, NAME) run(NAME, [&] { k<OP><<<nsm, 1024>>>(out, cyc, 0x00030005u, 0x00070009u); }, n, nsm, out, cyc)
  R(0, "VIADD.16x2");
  R(1, "VIMNMX.S16x2");
  R(2, "VIADDMNMX.S16x2");
  R(3, "VIADDMNMX.S16x2.RELU");
  R(4, "IDP.2A (acc chain)");
  R(12, "IDP.2A (multiplicand chain)");
  R(5, "PRMT");
  R(6, "IMAD");
  R(7, "IADD");
  R(8, "VIMNMX x2 (clamp s32)");
  R(9, "VIADDMNMX s32");
  R(10, "VIMNMX3.S16x2");
  R(11, "SHF.R.W (funnelshift)");
  run("mix 4x VIADDMNMX.relu + 4x IDP.2A", [&] { kmix<<<nsm, 1024>>>(out, cyc, 0x00030005u, 0x00070009u); }, n, nsm, out, cyc);
  run("LDS.32  (loads/clk/SM x32 lanes)", [&] { klds<1><<<nsm, 1024>>>(out, cyc, 3); }, 8.0 * ITER / 4, nsm, out, cyc);
  run("LDS.64", [&] { klds<2><<<nsm, 1024>>>(out, cyc, 3); }, 8.0 * ITER / 4, nsm, out, cyc);
  run("LDS.128", [&] { klds<4><<<nsm, 1024>>>(out, cyc, 3); }, 8.0 * ITER / 4, nsm, out, cyc);
  return 0;
}
