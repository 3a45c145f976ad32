// sao_device.cuh -- sample adaptive offset on 8-sample groups, two samples per register (sm_100a).
//
//   SampleAdaptiveOffset::offsetBlock   SampleAdaptiveOffset.cpp:293-547
//   deriveLoopFilterBoundaryAvailibility :668-729 (arrives as the 8 availability bits of the CTU)
//
// The input is the deblocked tile in shared memory (SAO reads the frozen deblocked picture, never its own output,
// SampleAdaptiveOffset.cpp:641); the output goes straight to global memory, so every sample is classified exactly once.
//   edge offset : sgn(a - v) + 1 = max(min(a + (1 - v), 2), 0) is ONE VIADDMNMX.S16x2.RELU; the category index picks the
//                 offset out of a 5-byte table with two PRMTs (sign-replicating selector nibble); clip = one more VIADDMNMX
//   band offset : band index by shift + mask on both lanes at once, the same byte table look-up
#pragma once

#include "packed16.cuh"
#include "vtmgpu_dev.cuh"

namespace vtmgpu
{

__device__ __forceinline__ bool saoCtuAvail(uint32_t avail, int rx, int ry)
{
  if ((rx | ry) == 0) return true;
  const int bit = ry == 0 ? (rx < 0 ? 0x01 : 0x02) : (ry < 0 ? (rx == 0 ? 0x04 : (rx < 0 ? 0x10 : 0x20)) : (rx == 0 ? 0x08 : (rx < 0 ? 0x40 : 0x80)));
  return (avail & bit) != 0;
}

// lanes of the 8-sample group at (x,y) that edge-offset class (dxa,dya) must leave untouched: a neighbour outside the
// picture or in a CTU that is not available (offsetBlock start/end and first/last line rules, SampleAdaptiveOffset.cpp
// :311-312,:340-341,:398-399,:444-445,:476-477,:514-515)
__device__ __noinline__ uint32_t saoSkipLanesSlow(int x, int y, int dxa, int dya, uint32_t avail, int w, int h, int cwLog, int chLog)
{
  const int cx = x >> cwLog, cy = y >> chLog, last = min(7, w - 1 - x);
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 2; k++)
  {
    const int ddx = k ? -dxa : dxa, ddy = k ? -dya : dya;
    const int ny = y + ddy;
    if (ny < 0 || ny >= h) { m = 0xff; continue; }
    const int ry = (ny >> chLog) - cy;
    uint32_t mm = saoCtuAvail(avail, 0, ry) ? 0u : 0xffu;
    if (ddx < 0)
    {
      const bool av = x > 0 && saoCtuAvail(avail, ((x - 1) >> cwLog) - cx, ry);
      mm = (mm & ~1u) | (av ? 0u : 1u);
    }
    if (ddx > 0)
    {
      const int nx = x + last + 1;
      const bool av = nx < w && saoCtuAvail(avail, (nx >> cwLog) - cx, ry);
      mm = (mm & ~(1u << last)) | (av ? 0u : (1u << last));
    }
    m |= mm;
  }
  return m;
}

__device__ __forceinline__ uint32_t saoSkipLanes(int x, int y, int dxa, int dya, uint32_t avail, int w, int h, int cwLog, int chLog)
{
  const int cwm = (1 << cwLog) - 1, chm = (1 << chLog) - 1;
  if ((y & chm) != 0 && ((y + 1) & chm) != 0 && y + 1 < h && (x & cwm) != 0 && ((x + 8) & cwm) != 0 && x + 8 < w) return 0;
  if (avail == 0xffu && y > 0 && y + 1 < h && x > 0 && x + 8 < w) return 0;       // all 8 neighbour CTUs usable, not on the picture border
  return saoSkipLanesSlow(x, y, dxa, dya, avail, w, h, cwLog, chLog);
}

// the 8 samples at horizontal offset DX from the group at base, as 4 packed registers
template <int DX> __device__ __forceinline__ uint4 saoNeighbours(const pel* base)
{
  const uint4 q = *reinterpret_cast<const uint4*>(base);
  if (DX == 0) return q;
  if (DX < 0)
  {
    const uint32_t l = *reinterpret_cast<const uint32_t*>(base - 2);
    return make_uint4(mid16(l, q.x), mid16(q.x, q.y), mid16(q.y, q.z), mid16(q.z, q.w));
  }
  const uint32_t r = *reinterpret_cast<const uint32_t*>(base + 8);
  return make_uint4(mid16(q.x, q.y), mid16(q.y, q.z), mid16(q.z, q.w), mid16(q.w, r));
}

// offset look-up for two lanes: k = packed indices 0..4, lut = 5 signed bytes (lutHi holds byte 4); returns packed s16
__device__ __forceinline__ uint32_t saoLut(uint32_t k, uint32_t lutLo, uint32_t lutHi)
{
  const uint32_t sel = k * 0x11u + 0x00800080u;             // per lane: low nibble k, high nibble k + 8 (= sign replicate)
  return prmt(lutLo, lutHi, prmt(sel, 0u, 0x4420u));
}

template <int DXA, int DYA> __device__ __forceinline__ uint4 saoEdge(const pel* ap, int pitch, uint4 v, uint32_t lutLo, uint32_t lutHi, uint32_t maxvP)
{
  const uint4 na = saoNeighbours<DXA>(ap + DYA * pitch), nb = saoNeighbours<-DXA>(ap - DYA * pitch);
  const uint32_t vv[4] = { v.x, v.y, v.z, v.w }, aa[4] = { na.x, na.y, na.z, na.w }, bb[4] = { nb.x, nb.y, nb.z, nb.w };
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; j++)
  {
    const uint32_t c1 = __vadd2(~vv[j], 0x00020002u);                           // 1 - v
    const uint32_t k = addClamp0(aa[j], c1, 0x00020002u) + addClamp0(bb[j], c1, 0x00020002u);   // 2 + sgn(a-v) + sgn(b-v)
    o[j] = addClamp0(vv[j], saoLut(k, lutLo, lutHi), maxvP);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ uint32_t laneMask2(uint32_t bits) { return ((bits & 1u) ? 0xffffu : 0u) | ((bits & 2u) ? 0xffff0000u : 0u); }

// lanes of the 8-sample group at (x,y) (component coordinates) that lie next to a signalled virtual boundary the edge class looks
// across (isProcessDisabled, SampleAdaptiveOffset.h:96-116: EO 0 tests the vertical boundaries only, EO 90 the horizontal ones only,
// the diagonal classes both; band offset is not affected).  type: 1 = EO 0, 2 = EO 90, 3 = EO 135, 4 = EO 45
__device__ __noinline__ uint32_t saoVbLanes(const VbDev* vb, int type, int x, int y, int csx, int csy)
{
  uint32_t m = 0;
  if (type != 2)
#pragma unroll 1
    for (int i = 0; i < vb->nv; i++)
    {
      const int d = (vb->x[i] >> csx) - x;               // lane of the first sample right of the boundary
      if (d >= 0 && d < 8) m |= 1u << d;
      if (d >= 1 && d < 9) m |= 1u << (d - 1);
    }
  if (type != 1)
#pragma unroll 1
    for (int i = 0; i < vb->nh; i++)
    {
      const int b = vb->y[i] >> csy;
      if (y == b || y == b - 1) m = 0xffu;
    }
  return m;
}

// SAO of a vertical strip of nrows 8-sample groups inside ONE CTU (one parameter set pq = the SaoDev as 4 words:
// .x = type | band << 8 | avail << 16, .y = off0 | off1 << 16, .z = off2 | off3 << 16, .w = off4).
//   a    deblocked samples of the first group in shared memory (pitch ap samples; the 8-aligned neighbours are resident)
//   out  destination of the first group in global memory (pitch op samples)
//   x,y  plane position of the first group; logs = cwLog | chLog << 8 | bitDepth << 16 (CTU size of this plane)
//   vb   signalled virtual boundaries (NULL: none); csxy = subsampling shifts of this plane, x | y << 8
// One copy of this code serves every component and tile shape.
__device__ __noinline__ void saoStrip(pel* out, int op, const pel* a, int ap, int nrows, int x, int y, uint4 pq, int w, int h, int logs, const VbDev* vb, int csxy)
{
  const int cwLog = logs & 0xff, chLog = (logs >> 8) & 0xff, bd = logs >> 16;
  const int type = pq.x & 0xff;
  if (type == 0)
  {
    for (int k = 0; k < nrows; k++) *reinterpret_cast<uint4*>(out + (size_t)k * op) = *reinterpret_cast<const uint4*>(a + k * ap);
    return;
  }
  const uint32_t maxvP = dup16((1 << bd) - 1);
  if (type == 5)
  {
    // band offset: k = (band - first band) & 31 ; bands k = 0..3 carry an offset (:528-541)
    const uint32_t lutLo = prmt(pq.y, pq.z, 0x6420u);
    const uint32_t nstart = dup16(-(int)((pq.x >> 8) & 0xff));
    const int sh = bd - 5;
    for (int k = 0; k < nrows; k++)
    {
      const uint4 v = *reinterpret_cast<const uint4*>(a + k * ap);
      uint32_t vv[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        const uint32_t band = (vv[j] >> sh) & 0x001f001fu;
        const uint32_t kk = __vminu2(__vadd2(band, nstart) & 0x001f001fu, 0x00040004u);
        vv[j] = addClamp0(vv[j], saoLut(kk, lutLo, 0u), maxvP);
      }
      *reinterpret_cast<uint4*>(out + (size_t)k * op) = make_uint4(vv[0], vv[1], vv[2], vv[3]);
    }
    return;
  }
  // edge offset: index k = 2 - edgeType  ->  the table holds the offsets in reverse order: off[4], off[3], off[2], off[1] | off[0]
  const uint32_t lutLo = (prmt(pq.y, pq.z, 0x2460u) & 0xffffff00u) | (pq.w & 0xffu);
  const uint32_t lutHi = pq.y & 0xffu;
  const uint32_t avail = (pq.x >> 16) & 0xff;
  const int dxa = type == 2 ? 0 : (type == 4 ? 1 : -1), dya = type == 1 ? 0 : -1;
  for (int k = 0; k < nrows; k++)
  {
    const pel* apk = a + k * ap;
    const uint4 v = *reinterpret_cast<const uint4*>(apk);
    uint4 o;
    if (type == 1)      o = saoEdge<-1, 0>(apk, ap, v, lutLo, lutHi, maxvP);
    else if (type == 2) o = saoEdge<0, -1>(apk, ap, v, lutLo, lutHi, maxvP);
    else if (type == 3) o = saoEdge<-1, -1>(apk, ap, v, lutLo, lutHi, maxvP);
    else                o = saoEdge<1, -1>(apk, ap, v, lutLo, lutHi, maxvP);
    uint32_t skip = saoSkipLanes(x, y + k, dxa, dya, avail, w, h, cwLog, chLog);
    if (vb) skip |= saoVbLanes(vb, type, x, y + k, csxy & 0xff, csxy >> 8);
    if (skip)
    {
      const uint32_t m0 = laneMask2(skip), m1 = laneMask2(skip >> 2), m2 = laneMask2(skip >> 4), m3 = laneMask2(skip >> 6);
      o.x = (o.x & ~m0) | (v.x & m0); o.y = (o.y & ~m1) | (v.y & m1); o.z = (o.z & ~m2) | (v.z & m2); o.w = (o.w & ~m3) | (v.w & m3);
    }
    *reinterpret_cast<uint4*>(out + (size_t)k * op) = o;
  }
}

}   // namespace vtmgpu
