// saoalf_kernel.cuh -- SAO + ALF + CC-ALF in ONE pass over the deblocked picture (sm_100a).
//
//   SampleAdaptiveOffset::SAOProcess  (SampleAdaptiveOffset.cpp:618, offsetBlock :293-547)
//   AdaptiveLoopFilter::ALFProcess    (AdaptiveLoopFilter.cpp:393; deriveClassificationBlk :873-1082,
//                                      filterBlk :1084-1324, filterBlkCcAlf :1327-1416)
//
// One CTA owns a 64x64 luma tile and the collocated chroma tiles.  The reference makes two whole-picture temp copies
// (SAO input, ALF input incl. 3-sample replicate border); here each plane is read ONCE from HBM into shared memory
// (tile + halo, 8-sample aligned 128-bit loads, coordinates clamped = replicate border), SAO is applied in shared
// memory (tile + 3), the 4x4 Laplacian classification and the diamond filters read that SAO output, CC-ALF reads the
// SAO-output luma tile that is still resident, and each plane is written ONCE.  With saoOn=0 / alfOn=0 the same
// kernel is the stand-alone ALF / SAO stage.
//
// Per luma pixel algorithmic HBM bytes at 4:2:0: read 3, write 3 (+ CTU params, negligible).
#pragma once

#include "vtmgpu_dev.cuh"
#include "vtmgpu.h"

namespace vtmgpu
{

constexpr int SA_T = 64;                    // luma tile edge
constexpr int SA_THREADS = 256;             // = (SA_T/4)^2 : one thread per 4x4 luma block
constexpr int SA_HX = 8, SA_HY = 4;         // halo loaded around a tile (x: one aligned group of 8)
constexpr int SA_W = SA_T + 2 * SA_HX;      // 80
constexpr int SA_H = SA_T + 2 * SA_HY;      // 72
constexpr int SA_P = SA_W + 8;              // smem pitch in samples (88 -> 176 B)
constexpr int SA_LAPN = SA_T / 2 + 2;       // 34 gradient positions per dimension
constexpr int SA_LAPP = SA_LAPN + 1;

__constant__ int8_t c_perm7[4][12] = { { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 }, { 9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6 },
                                       { 0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11 }, { 9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6 } };

struct SaoAlfSmem
{
  pel      a[SA_H * SA_P];                  // input tile of the current component
  pel      bl[SA_H * SA_P];                 // SAO output, luma (stays resident for CC-ALF)
  union
  {
    uint16_t lap[4][SA_LAPN][SA_LAPP];      // V, H, D0, D1 Laplacian pair sums
    pel      bc[SA_H * SA_P];               // SAO output, chroma component
  } u;
  SaoDev   sao[3][9];                       // 3x3 CTU neighbourhood per component
  short2   lumaSet[25][12];
  short2   chromaSet[2][6];
  int16_t  cc[2][8];
};

// loads rows y0-SA_HY .. , columns x0-SA_HX .. of a plane into s (tile w x h samples + halo), replicate border
__device__ __forceinline__ void saLoadTile(pel* s, const PlaneDev& pl, int x0, int y0, int tw, int th)
{
  const int groups = (tw + 2 * SA_HX) >> 3, rows = th + 2 * SA_HY;
  for (int i = threadIdx.x; i < groups * rows; i += SA_THREADS)
  {
    const int r = i / groups, gc = i - r * groups;
    const int y = min(max(y0 - SA_HY + r, 0), pl.h - 1), x = x0 - SA_HX + gc * 8;
    const pel* row = pl.p + (size_t)y * pl.pitch;
    int4 v;
    if (x >= 0 && x < pl.w) v = __ldg(reinterpret_cast<const int4*>(row + x));
    else
    {
      const uint32_t e = (uint16_t)row[x < 0 ? 0 : pl.w - 1];
      const int ee = (int)(e | (e << 16));
      v = make_int4(ee, ee, ee, ee);
    }
    *reinterpret_cast<int4*>(&s[r * SA_P + gc * 8]) = v;
  }
}

// SAO of the sample at plane position (x,y) (inside the picture); a = smem tile, (ax,ay) its tile coordinates
__device__ __forceinline__ int saoSample(const pel* a, int ax, int ay, int x, int y, int w, int h, int cw, int ch, int tcx, int tcy,
                                         const SaoDev* nb, int bd)
{
  const int v = a[ay * SA_P + ax];
  const int cx = x / cw, cy = y / ch;
  const SaoDev& P = nb[(cy - tcy + 1) * 3 + (cx - tcx + 1)];
  if (P.type == 0) return v;
  const int maxv = (1 << bd) - 1;
  if (P.type == 5)
  {
    const int k = ((v >> (bd - 5)) - P.band) & 31;
    return k < 4 ? clip3(0, maxv, v + P.off[k]) : v;
  }
  const int t = P.type - 1;                       // 0: 0deg, 1: 90deg, 2: 135deg, 3: 45deg
  const int dx = t == 1 ? 0 : (t == 3 ? 1 : -1);  // first neighbour (dx,dy); second is (-dx,-dy)
  const int dy = t == 0 ? 0 : -1;
  int e = 0;
#pragma unroll
  for (int k = 0; k < 2; k++)
  {
    const int sx_ = k ? -dx : dx, sy_ = k ? -dy : dy;
    const int nx = x + sx_, ny = y + sy_;
    if (nx < 0 || ny < 0 || nx >= w || ny >= h) return v;
    const int rx = nx / cw - cx, ry = ny / ch - cy;
    if (rx | ry)
    {
      // VTMGPU_AVAIL_* bit of the neighbouring CTU at (rx,ry)
      const int bit = ry == 0 ? (rx < 0 ? 0x01 : 0x02) : (ry < 0 ? (rx == 0 ? 0x04 : (rx < 0 ? 0x10 : 0x20)) : (rx == 0 ? 0x08 : (rx < 0 ? 0x40 : 0x80)));
      if (!(P.avail & bit)) return v;
    }
    const int n = a[(ay + sy_) * SA_P + ax + sx_];
    e += (v > n) - (v < n);
  }
  return clip3(0, maxv, v + P.off[2 + e]);
}

// SAO over tile + 3 (everything the ALF stage may read); positions outside the picture take the value of the
// clamped position (= UnitBuf::extendBorderPel of the SAO output, AdaptiveLoopFilter.cpp:411)
__device__ __forceinline__ void saSaoTile(pel* b, const pel* a, bool on, int x0, int y0, int tw, int th, int w, int h, int cw, int ch,
                                          int tcx, int tcy, const SaoDev* nb, int bd)
{
  const int cols = tw + 6, rows = th + 6;
  for (int i = threadIdx.x; i < cols * rows; i += SA_THREADS)
  {
    const int r = i / cols, c = i - r * cols;
    const int px = x0 - 3 + c, py = y0 - 3 + r;
    const int x = min(max(px, 0), w - 1), y = min(max(py, 0), h - 1);
    const int ax = x - (x0 - SA_HX), ay = y - (y0 - SA_HY);
    b[(py - (y0 - SA_HY)) * SA_P + px - (x0 - SA_HX)] = (pel)(on ? saoSample(a, ax, ay, x, y, w, h, cw, ch, tcx, tcy, nb, bd) : a[ay * SA_P + ax]);
  }
}

// one output sample of filterBlk; c = centre in smem, o1..o3 = row offsets (already limited by the virtual boundary)
__device__ __forceinline__ int alfTap(const pel* c, int off, int cur, short2 f)
{
  const int cl = f.y;
  return f.x * (clip3(-cl, cl, (int)c[off] - cur) + clip3(-cl, cl, (int)c[-off] - cur));
}

__device__ __forceinline__ void vbLimit(int yv, int vbPos, int span, int& lim, bool& nearVb)
{
  lim = 3; nearVb = false;
  if (yv < vbPos && yv >= vbPos - span) { lim = vbPos - 1 - yv; nearVb = yv == vbPos - 1; }
  else if (yv >= vbPos && yv <= vbPos + span - 1) { lim = yv - vbPos; nearVb = yv == vbPos; }
}

__global__ void __launch_bounds__(SA_THREADS) k_sao_alf(const SlotDev* __restrict__ slots, int firstSlot, int srcBuf, int dstBuf, Geom g, int tilesX,
                                                        int doSao, int doAlf)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  SaoAlfSmem& sm = *reinterpret_cast<SaoAlfSmem*>(smraw);
  const SlotDev& S = slots[firstSlot + blockIdx.y];
  const int tid = threadIdx.x;
  const int x0 = (blockIdx.x % tilesX) * SA_T, y0 = (blockIdx.x / tilesX) * SA_T;
  const int tcx = x0 >> g.ctuLog2, tcy = y0 >> g.ctuLog2, ctuIdx = tcy * g.wCtus + tcx;
  const int nCtus = g.wCtus * g.hCtus;
  const bool saoOn = doSao && S.saoOn, alfPic = doAlf && S.alfOn && (S.alf->enabled[0] | S.alf->enabled[1] | S.alf->enabled[2]);

  // ---- per-tile parameters ------------------------------------------------------------------------------
  if (saoOn && tid < 27)
  {
    const int c = tid / 9, k = tid - c * 9, cx = tcx + k % 3 - 1, cy = tcy + k / 3 - 1;
    SaoDev z = {};
    if (cx >= 0 && cy >= 0 && cx < g.wCtus && cy < g.hCtus && c < g.ncomp) z = S.sao[(cy * g.wCtus + cx) * 3 + c];
    sm.sao[c][k] = z;
  }
  bool alfY = false, alfC[2] = { false, false };
  int ccIdc[2] = { 0, 0 };
  if (alfPic)
  {
    alfY = S.alfCtu[0 * nCtus + ctuIdx] != 0;
    for (int c = 0; c < 2; c++)
    {
      alfC[c] = g.ncomp > 1 && S.alfCtu[(1 + c) * nCtus + ctuIdx] != 0;
      ccIdc[c] = (g.ncomp > 1 && S.alf->ccEnabled[c]) ? S.alfCtu[(5 + c) * nCtus + ctuIdx] : 0;
    }
    if (alfY)
    {
      const short2* set = &S.alf->luma[S.alfFilterIdx[ctuIdx]][0][0];
      for (int i = tid; i < 25 * 12; i += SA_THREADS) (&sm.lumaSet[0][0])[i] = set[i];
    }
    if (tid < 12 && alfC[tid / 6]) sm.chromaSet[tid / 6][tid % 6] = S.alf->chroma[S.alfCtu[(3 + tid / 6) * nCtus + ctuIdx]][tid % 6];
    if (tid < 16 && ccIdc[tid >> 3]) sm.cc[tid >> 3][tid & 7] = S.alf->cc[tid >> 3][ccIdc[tid >> 3] - 1][tid & 7];
  }

  // ---- luma ----------------------------------------------------------------------------------------------
  const PlaneDev srcY = S.buf[srcBuf][0], dstY = S.buf[dstBuf][0];
  saLoadTile(sm.a, srcY, x0, y0, SA_T, SA_T);
  __syncthreads();
  saSaoTile(sm.bl, sm.a, saoOn, x0, y0, SA_T, SA_T, g.w, g.h, g.ctu, g.ctu, tcx, tcy, sm.sao[0], g.bdL);
  __syncthreads();

  const int vbL = g.ctu - 4;
  if (alfY)
  {
    // Laplacian pair sums at the 2x2-subsampled positions (r,c) = (y0-2+2i, x0-2+2j)
    for (int i = tid; i < SA_LAPN * SA_LAPN; i += SA_THREADS)
    {
      const int li = i / SA_LAPN, lj = i - li * SA_LAPN;
      const int r = y0 - 2 + 2 * li;
      const pel* p0 = &sm.bl[(2 * li + SA_HY - 2) * SA_P + 2 * lj + SA_HX - 2];   // (r, c)
      int up = -SA_P, dn2 = 2 * SA_P;                                            // row r-1, row r+2
      const int rv = r & (g.ctu - 1);
      if (r > 0 && rv == vbL - 2) dn2 = SA_P;
      else if (r > 0 && rv == vbL) up = 0;
      const int y0v = p0[0] << 1, y1v = p0[SA_P + 1] << 1;
      sm.u.lap[0][li][lj] = (uint16_t)(iabs(y0v - p0[up] - p0[SA_P]) + iabs(y1v - p0[1] - p0[dn2 + 1]));
      sm.u.lap[1][li][lj] = (uint16_t)(iabs(y0v - p0[1] - p0[-1]) + iabs(y1v - p0[SA_P + 2] - p0[SA_P]));
      sm.u.lap[2][li][lj] = (uint16_t)(iabs(y0v - p0[up - 1] - p0[SA_P + 1]) + iabs(y1v - p0[0] - p0[dn2 + 2]));
      sm.u.lap[3][li][lj] = (uint16_t)(iabs(y0v - p0[SA_P - 1] - p0[up + 1]) + iabs(y1v - p0[dn2] - p0[2]));
    }
    __syncthreads();
  }
  {
    const int bi = tid >> 4, bj = tid & 15;
    const int bx = x0 + 4 * bj, by = y0 + 4 * bi;
    if (bx < g.w && by < g.h)
    {
      const pel* c0 = &sm.bl[(4 * bi + SA_HY) * SA_P + 4 * bj + SA_HX];
      pel* out = dstY.p + (size_t)by * dstY.pitch + bx;
      if (!alfY)
      {
#pragma unroll
        for (int r = 0; r < 4; r++) *reinterpret_cast<int2*>(out + (size_t)r * dstY.pitch) = *reinterpret_cast<const int2*>(c0 + r * SA_P);
      }
      else
      {
        // classification of this 4x4 block (deriveClassificationBlk)
        const int yb = by & (g.ctu - 1);
        const int i0 = (yb == vbL) ? 1 : 0, i1 = (yb == vbL - 4) ? 3 : 4;
        int sum[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int d = 0; d < 4; d++)
          for (int i = i0; i < i1; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) sum[d] += sm.u.lap[d][2 * bi + i][2 * bj + j];
        const int sumV = sum[0], sumH = sum[1], sumD0 = sum[2], sumD1 = sum[3];
        const int scale = (yb == vbL - 4 || yb == vbL) ? 96 : 64;
        const int act = clip3(0, 15, ((sumV + sumH) * scale) >> (g.bdL + 4));
        // th[] = {0,1,2,2,2,2,2,3,3,3,3,3,3,3,3,4}
        int cls = act == 0 ? 0 : (act == 1 ? 1 : (act < 7 ? 2 : (act < 15 ? 3 : 4)));
        int hv1, hv0, d1, d0, dirHV, dirD;
        if (sumV > sumH) { hv1 = sumV; hv0 = sumH; dirHV = 1; } else { hv1 = sumH; hv0 = sumV; dirHV = 3; }
        if (sumD0 > sumD1) { d1 = sumD0; d0 = sumD1; dirD = 0; } else { d1 = sumD1; d0 = sumD0; dirD = 2; }
        int hvd1, hvd0, mainDir, secDir;
        if ((uint32_t)d1 * (uint32_t)hv0 > (uint32_t)hv1 * (uint32_t)d0) { hvd1 = d1; hvd0 = d0; mainDir = dirD; secDir = dirHV; }
        else { hvd1 = hv1; hvd0 = hv0; mainDir = dirHV; secDir = dirD; }
        int strength = 0;
        if (hvd1 > 2 * hvd0) strength = 1;
        if (hvd1 * 2 > 9 * hvd0) strength = 2;
        if (strength) cls += (((mainDir & 1) << 1) + strength) * 5;
        // transposeTable = {0,1,0,2,2,3,1,3}[mainDir*2 + (secDir>>1)]
        const int tIdx = (0x31322010 >> (4 * (mainDir * 2 + (secDir >> 1)))) & 0xf;

        short2 f[12];
#pragma unroll
        for (int k = 0; k < 12; k++) f[k] = sm.lumaSet[cls][c_perm7[tIdx][k]];
        const int maxv = (1 << g.bdL) - 1;
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
          int lim; bool nearVb;
          vbLimit((by + r) & (g.ctu - 1), vbL, 4, lim, nearVb);
          const int o1 = min(1, lim) * SA_P, o2 = min(2, lim) * SA_P, o3 = lim * SA_P;   // lim <= 3
          pel res[4];
#pragma unroll
          for (int q = 0; q < 4; q++)
          {
            const pel* c = c0 + r * SA_P + q;
            const int cur = c[0];
            int s = alfTap(c, o3, cur, f[0]) + alfTap(c, o2 + 1, cur, f[1]) + alfTap(c, o2, cur, f[2]) + alfTap(c, o2 - 1, cur, f[3]) +
                    alfTap(c, o1 + 2, cur, f[4]) + alfTap(c, o1 + 1, cur, f[5]) + alfTap(c, o1, cur, f[6]) + alfTap(c, o1 - 1, cur, f[7]) +
                    alfTap(c, o1 - 2, cur, f[8]) + alfTap(c, 3, cur, f[9]) + alfTap(c, 2, cur, f[10]) + alfTap(c, 1, cur, f[11]);
            s = (s + 64) >> (nearVb ? 10 : 7);
            res[q] = (pel)clip3(0, maxv, cur + s);
          }
          *reinterpret_cast<int2*>(out + (size_t)r * dstY.pitch) = *reinterpret_cast<const int2*>(res);
        }
      }
    }
  }
  if (g.ncomp == 1) return;

  // ---- chroma (Cb then Cr reuse the same shared buffers) ---------------------------------------------------
  const int tw = SA_T >> g.sx, th = SA_T >> g.sy, cx0 = x0 >> g.sx, cy0 = y0 >> g.sy;
  const int cw = g.w >> g.sx, chh = g.h >> g.sy, ctuW = g.ctu >> g.sx, ctuH = g.ctu >> g.sy;
  const int vbC = ctuH - 2, maxc = (1 << g.bdC) - 1, half = (1 << g.bdC) >> 1;
  for (int c = 0; c < 2; c++)
  {
    const PlaneDev srcC = S.buf[srcBuf][1 + c], dstC = S.buf[dstBuf][1 + c];
    __syncthreads();                                  // previous users of sm.a / sm.u are done
    saLoadTile(sm.a, srcC, cx0, cy0, tw, th);
    __syncthreads();
    saSaoTile(sm.u.bc, sm.a, saoOn, cx0, cy0, tw, th, cw, chh, ctuW, ctuH, tcx, tcy, sm.sao[1 + c], g.bdC);
    __syncthreads();
    const bool fOn = alfC[c];
    const int idc = ccIdc[c];
    short2 f[6];
    if (fOn)
    {
#pragma unroll
      for (int k = 0; k < 6; k++) f[k] = sm.chromaSet[c][k];
    }
    int cc[7];
    if (idc)
    {
#pragma unroll
      for (int k = 0; k < 7; k++) cc[k] = sm.cc[c][k];
    }
    // one thread = 4 horizontally adjacent chroma samples
    const int quads = (tw >> 2) * th;
    for (int i = tid; i < quads; i += SA_THREADS)
    {
      const int r = i / (tw >> 2), qx = (i - r * (tw >> 2)) * 4;
      const int x = cx0 + qx, y = cy0 + r;
      if (x >= cw || y >= chh) continue;
      const pel* cb = &sm.u.bc[(r + SA_HY) * SA_P + qx + SA_HX];
      int lim; bool nearVb;
      vbLimit(y & (ctuH - 1), vbC, 2, lim, nearVb);
      const int o1 = min(1, lim) * SA_P, o2 = min(2, lim) * SA_P;
      // CC-ALF row offsets in the luma tile (filterBlkCcAlf :1376-1386)
      const int ly = (r << g.sy) + SA_HY, lpos = (y << g.sy) & (g.ctu - 1);
      int l1 = SA_P, l2 = -SA_P, l3 = 2 * SA_P;
      if (lpos == vbL - 2 || lpos == vbL + 1) l3 = SA_P;
      else if (lpos == vbL - 1 || lpos == vbL) l1 = l2 = l3 = 0;
      pel res[4];
#pragma unroll
      for (int q = 0; q < 4; q++)
      {
        const pel* p = cb + q;
        int v = p[0];
        if (fOn)
        {
          int s = alfTap(p, o2, v, f[0]) + alfTap(p, o1 + 1, v, f[1]) + alfTap(p, o1, v, f[2]) + alfTap(p, o1 - 1, v, f[3]) +
                  alfTap(p, 2, v, f[4]) + alfTap(p, 1, v, f[5]);
          s = (s + 64) >> (nearVb ? 10 : 7);
          v = clip3(0, maxc, v + s);
        }
        if (idc)
        {
          const pel* l = &sm.bl[ly * SA_P + ((qx + q) << g.sx) + SA_HX];
          const int cur = l[0];
          int s = cc[0] * (l[l2] - cur) + cc[1] * (l[-1] - cur) + cc[2] * (l[1] - cur) + cc[3] * (l[l1 - 1] - cur) + cc[4] * (l[l1] - cur) +
                  cc[5] * (l[l1 + 1] - cur) + cc[6] * (l[l3] - cur);
          s = (s + 64) >> 7;
          s = clip3(0, maxc, s + half) - half;
          v = clip3(0, maxc, v + s);
        }
        res[q] = (pel)v;
      }
      *reinterpret_cast<int2*>(dstC.p + (size_t)y * dstC.pitch + x) = *reinterpret_cast<const int2*>(res);
    }
  }
}

}   // namespace vtmgpu
