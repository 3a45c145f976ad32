// border_kernel.cuh -- reference-picture border extension on the device (sm_100a).
//
//   Picture::extendPicBorder   (CommonLib/Picture.cpp:737-772): every sample of the margin around the reconstructed picture takes the
//   value of the nearest picture sample (left / right margins row by row, then whole rows copied upwards and downwards).
//
// The filtered picture is written once more into a padded buffer -- picture + margins -- so that the download is ONE 2-D copy per
// plane straight into the decoder's picture buffer (whose allocation already has the margins).  Pure data movement: one thread per
// 8-sample group of the padded plane, 128-bit loads and stores where the group lies inside the picture.
#pragma once

#include "vtmgpu_dev.cuh"

namespace vtmgpu
{

struct ExtendArgs
{
  const pel* src[3];
  pel* dst[3];
  int w[3], h[3], srcPitch[3], dstPitch[3], xm[3], ym[3];      // dstPitch, xm multiples of 8 samples
  int rowStart[4];                                            // first padded row of each plane in the grid's y range
  int ncomp;
};

__global__ void __launch_bounds__(256) k_extend_border(const ExtendArgs a)
{
  const int gy = blockIdx.y;
  int k = 0;
  while (k + 1 < a.ncomp && gy >= a.rowStart[k + 1]) k++;
  const int py = gy - a.rowStart[k];                          // padded row
  const int groups = (a.w[k] + 2 * a.xm[k] + 7) >> 3;
  const int gx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gx >= groups) return;
  const int y = min(max(py - a.ym[k], 0), a.h[k] - 1), x0 = gx * 8 - a.xm[k];
  const pel* s = a.src[k] + (size_t)y * a.srcPitch[k];
  uint4 v;
  if (x0 >= 0 && x0 + 8 <= a.w[k]) v = *reinterpret_cast<const uint4*>(s + x0);
  else
  {
    uint32_t wv[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
      const int xa = min(max(x0 + 2 * i, 0), a.w[k] - 1), xb = min(max(x0 + 2 * i + 1, 0), a.w[k] - 1);
      wv[i] = (uint32_t)(uint16_t)s[xa] | (uint32_t)(uint16_t)s[xb] << 16;
    }
    v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
  }
  *reinterpret_cast<uint4*>(a.dst[k] + (size_t)py * a.dstPitch[k] + gx * 8) = v;
}

}   // namespace vtmgpu
