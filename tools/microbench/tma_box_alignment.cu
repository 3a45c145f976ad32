// standalone check of a 2-D u32 TMA box load with the geometry of the deblocking record arrays
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../vvc_b200/csrc/async_copy.cuh"
using namespace vtmgpu;
__global__ void k(const CUtensorMap* map, uint32_t* out, int x, int y, int cols, int rows)
{
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8192);
  if (threadIdx.x == 0) { mbarInit(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) { mbarExpectTx(bar, cols * rows * 4); tmaLoad2D(sm, map, x, y, bar); }
  mbarWait(bar, 0);
  for (int i = threadIdx.x; i < cols * rows; i += blockDim.x) out[i] = reinterpret_cast<uint32_t*>(sm)[i];
}
int main(int argc, char** argv)
{
  const int cols = argc > 1 ? atoi(argv[1]) : 36, rows = 20;
  const int W = argc > 2 ? atoi(argv[2]) : 104, H = 60, P = argc > 3 ? atoi(argv[3]) : W;
  const int l2 = argc > 4 ? atoi(argv[4]) : 2;
  const int u16 = argc > 5 ? atoi(argv[5]) : 0;
  const int cx = argc > 6 ? atoi(argv[6]) : -1, cy = argc > 7 ? atoi(argv[7]) : -2;
  std::vector<uint32_t> h((size_t)P * H);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint32_t)i + 1;
  uint32_t *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 8192);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap m;
  const cuuint64_t dims[2] = { (cuuint64_t)W * (u16 ? 2 : 1), (cuuint64_t)H }, strides[1] = { (cuuint64_t)P * 4 };
  const cuuint32_t box[2] = { (cuuint32_t)cols * (u16 ? 2 : 1), (cuuint32_t)rows }, es[2] = { 1, 1 };
  CUresult r = ((EncodeFn)fn)(&m, u16 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d cols=%d W=%d P=%d l2=%d\n", (int)r, cols, W, P, l2);
  CUtensorMap* dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  k<<<1, 128, 16384>>>(dm, o, u16 ? 2 * cx : cx, cy, cols, rows);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<uint32_t> res(cols * rows);
  cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
  printf("row2: %u %u %u ... expect 0 1 2\n", res[2 * cols], res[2 * cols + 1], res[2 * cols + 2]);
  return 0;
}
