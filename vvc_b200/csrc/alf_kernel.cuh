// alf_kernel.cuh -- ALF + CC-ALF over the SAO output picture (sm_100a).
//
//   AdaptiveLoopFilter::ALFProcess    (AdaptiveLoopFilter.cpp:393; deriveClassificationBlk :873-1082,
//                                      filterBlk :1084-1324, filterBlkCcAlf :1327-1416)
//
// Persistent CTAs walk the 64x64 luma tiles (+ collocated chroma tiles) of a batch of pictures.  The reference copies the
// whole picture and pads it by 3 samples (:408-411); here the three planes of a tile arrive by TMA (tile + halo, zero
// filled outside the picture, border samples replicated in shared memory for picture-border tiles), double buffered so
// that the loads of tile i+1 fly while tile i is filtered; the 4x4 Laplacian classification and the diamond filters read
// shared memory, CC-ALF reads the luma tile that is still resident, and each plane is written ONCE.
//
// The chain is instruction-issue bound on B200, so all sample arithmetic runs TWO SAMPLES PER
// 32-BIT REGISTER on the native packed 16-bit integer instructions of sm_100a:
//   VIADD.16x2            packed add                      (__vadd2)
//   VIADDMNMX.S16x2.RELU  max(min(a+b, c), 0) per lane    (__viaddmin_s16x2_relu): one instruction = subtract + clamp
//   IDP.2A                s16 x s8 dot product into s32   (__dp2a_lo / __dp2a_hi): the ALF multiply-accumulate
//   PRMT / SHF            16-bit lane shuffles
// A filter tap pair costs 3 ALU-pipe + 3 FMA-pipe instructions per two pixels (scalar code: ~14 per two pixels).
// The 7x7 luma filter handles the rows next to the ALF virtual boundary (CTU height - 4) in the same packed loop (row offsets
// clamped at run time); their Laplacian cells, wide coefficients and non-4:2:0 CC-ALF use the generic scalar routines in this
// file, which follow the reference line by line.
//
// Sides the filter must not read across -- the picture border, slice / tile boundaries (one clip byte per CTU) and signalled
// virtual boundaries -- are all the same operation: the samples outside a clamp window are replaced by the nearest sample inside
// (saReplicateBorder).  A tile that a virtual boundary cuts, or that covers several CTUs (CTU size 32), is filtered part by part
// (instantiation k_alf<true>, only launched for such pictures): every part pads its own copy of the tile.
//
// Per luma pixel algorithmic HBM bytes at 4:2:0: read 3, write 3 (+ CTU params, negligible).
#pragma once

#include <cuda.h>

#include "async_copy.cuh"
#include "packed16.cuh"
#include "vtmgpu_dev.cuh"
#include "vtmgpu.h"
#include "alf_fast.cuh"

namespace vtmgpu
{

constexpr int SA_T = 64;                    // luma tile width
constexpr int SA_TH = 64;                   // luma tile height
constexpr int SA_THLOG = 6;
constexpr int SA_THREADS = (SA_T / 4) * (SA_TH / 4);   // 256: one thread per 4x4 luma block
#ifndef SA_CTAS_PER_SM
#define SA_CTAS_PER_SM 2
#endif
#ifndef ALF_WHATIF
#define ALF_WHATIF 0        // timing experiments only (results are wrong): 1 no chroma, 2 no classification, 4 no phase 1, 8 no luma filter
#endif
#ifndef ALF_BAL
#define ALF_BAL 2               // which adds of the luma tap run on the ALU pipe (alfLumaBlockV): bit 0 the tap sum, bit 1 clip - cur
#endif
constexpr int SA_HX = 8, SA_HY = 4;         // halo loaded around a tile (x: one aligned group of 8)
constexpr int SA_W = SA_T + 2 * SA_HX;      // 80
constexpr int SA_H = SA_TH + 2 * SA_HY;     // 72
constexpr int SA_P = SA_W + 8;              // smem pitch in samples (88 -> 176 B: rows shift by 12 banks)
constexpr int SA_CELLS = SA_T / 2 + 2;      // 34 Laplacian cells (2x2 samples) per row: tile + 2 samples each side
constexpr int SA_CELLR = SA_TH / 2 + 2;     // 34 cell rows
constexpr int SA_CELLP = SA_CELLS + 2;      // cell row pitch (uint2 units)

__constant__ int8_t c_perm7[4][12] = { { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 }, { 9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6 },
                                       { 0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11 }, { 9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6 } };

// Shared memory of one CTA (dynamic; offsets in bytes).  Input tiles are double buffered: while a CTA filters tile i the
// TMA copies of tile i+1 are in flight.
//   A[stage][comp]  input tile + halo (SAO output samples; replicate padded at the picture border by the CTA)
//   cell            Laplacian sums of the 2x2 cells: .x = V | H << 16, .y = D0 | D1 << 16
//   ctl[stage]      control record of the tile's CTU
//   bar[stage]      mbarriers the TMA loads complete on
struct SaLayout
{
  int pitchC, rowsC;               // chroma tile buffers: pitch in samples, rows
  int lumaBytes, chromaBytes, stageBytes, offSet, offSmall, offCell, offV, offPar, offBar, offDesc, total;
  __host__ __device__ int comp(int c) const { return c ? lumaBytes + (c - 1) * chromaBytes : 0; }
  __host__ __device__ int offA(int stage, int c) const { return stage * stageBytes + comp(c); }
};

constexpr int AT_SET_BYTES = 100 * (int)sizeof(AlfLumaEntry);      // one luma filter set: 25 classes x 4 transposes (17 600 bytes)

// all tile buffer sizes are multiples of 128 bytes (TMA destination alignment)
// tables = true (k_alf): every stage also holds the luma filter set of the tile's CTU and the chroma / CC-ALF operand tables of its
// picture, brought in by bulk copies next to the sample tiles
__host__ __device__ inline SaLayout saLayout(int sx, int sy, int ncomp, bool tables = false)
{
  SaLayout L;
  const int tw = SA_T >> sx, th = SA_TH >> sy;
  L.pitchC = tw + 2 * SA_HX;            // 48 samples = 24 words at 4:2:0: chroma rows r and r + 2 are 16 banks apart
  L.rowsC = th + 2 * SA_HY;
  L.lumaBytes = SA_H * SA_P * 2;
  L.chromaBytes = ncomp > 1 ? L.rowsC * L.pitchC * 2 : 0;
  L.offSet = L.lumaBytes + 2 * L.chromaBytes;                  // inside a stage
  L.offSmall = L.offSet + (tables ? (AT_SET_BYTES + 127) / 128 * 128 : 0);
  L.stageBytes = L.offSmall + (tables ? ALF_SMALL_BYTES : 0);
  L.offCell = 2 * L.stageBytes;
  L.offV = L.offCell + SA_CELLR * SA_CELLP * 8;                // vertical-pair copy of the luma tile (alf_fast.cuh), 16-byte aligned
  L.offPar = L.offV + AV_BYTES;
  L.offBar = L.offPar + 2 * 4 * (int)sizeof(CtuCtlDev);      // per stage: the control records of the (up to 2 x 2) CTUs under the tile
  L.offDesc = L.offBar + 32;                                 // per stage: tile descriptor written by the walking thread (k_alf); bars: 2 x tile loads, cells ready, tile done
  L.total = L.offDesc + 2 * 48;
  return L;
}

// TMA fills positions outside the picture with zeros; the filters want the border samples replicated
// (rows / columns of tile +- margin that lie outside take the value of the clamped position).  The valid window
// [xlo,xhi) x [ylo,yhi) is the picture, narrowed to the CTU on the sides the ALF must not read across (slice / tile
// boundaries, ALFProcess :452-477: copy of the CTU + extendBorderPel = the same clamping).
__device__ __noinline__ void saReplicateBorder(pel* s, int bx0, int by0, int xlo, int xhi, int ylo, int yhi, int pitch, int tw, int th, int margin)
{
  const int xl = bx0 + SA_HX - margin, yt = by0 + SA_HY - margin;
  const int cols = tw + 2 * margin, rows = th + 2 * margin;
  const int nl = max(0, xlo - xl), cr = min(cols, max(0, xhi - xl));    // columns [0,nl) and [cr,cols) of the region are outside
  const int nTop = min(rows, max(0, ylo - yt)), rBot = min(rows, max(nTop, yhi - yt));      // rows [0,nTop) and [rBot,rows) are outside
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows outside the window (a handful): the whole row from the clamped row, a warp per row
  const int nOut = nTop + rows - rBot;
  for (int q = warp; q < nOut; q += SA_THREADS / 32)
  {
    const int rr = q < nTop ? q : rBot + q - nTop;
    const int py = yt + rr, sy_ = min(max(py, ylo), yhi - 1);
    const pel* srow = s + (sy_ - by0) * pitch - bx0;
    pel* drow = s + (py - by0) * pitch - bx0;
    for (int cc = lane; cc < cols; cc += 32) { const int px = xl + cc; drow[px] = srow[min(max(px, xlo), xhi - 1)]; }
  }
  // rows inside: only the columns outside (none for a tile that touches the top / bottom side alone), a lane per column
  const int nCol = nl + cols - cr;
  if (nCol > 0)
    for (int rr = nTop + warp; rr < rBot; rr += SA_THREADS / 32)
    {
      pel* row = s + (yt + rr - by0) * pitch - bx0;
      for (int k = lane; k < nCol; k += 32)
      {
        const bool left = k < nl;
        row[xl + (left ? k : cr + k - nl)] = row[left ? xlo : xhi - 1];
      }
    }
}

// raster-scan slices (ALFProcess :478-488, padBorderPel): the CTU's top-left (bottom-right) neighbour is in another slice
// while the adjacent ones are not -- in the corner outside the CTU [cx0,cx1) x [cy0,cy1) every row takes the sample of the
// CTU's first (last) column of the same row.  Rows above / below the CTU hold real samples here (those sides are not clipped).
// pad = the pad flags of the part in flight (the caller keeps them only for the part that holds that corner of the CTU).
__device__ __forceinline__ void saPadCorners(pel* s, int bx0, int by0, int cx0, int cy0, int cx1, int cy1, int pitch, int margin, int pad)
{
  const int t = threadIdx.x;
  if (t >= 2 * margin * margin) return;
  const bool br = t >= margin * margin;
  const int k = br ? t - margin * margin : t, dy = k / margin, dx = k - dy * margin;
  if (!br)
  {
    if (!(pad & VTMGPU_ALF_PAD_TL)) return;
    pel* row = s + (cy0 - 1 - dy - by0) * pitch - bx0;
    row[cx0 - 1 - dx] = row[cx0];
  }
  else
  {
    if (!(pad & VTMGPU_ALF_PAD_BR)) return;
    pel* row = s + (cy1 + dy - by0) * pitch - bx0;
    row[cx1 + dx] = row[cx1 - 1];
  }
}

// Cuts of a tile (cold path, kept out of line).  A tile is filtered in parts when a signalled virtual boundary lies strictly inside it
// or when it covers several CTUs (CTU size 32: every CTU has its own control record and clip flags).  xs / ys receive the part
// boundaries in ascending order, first = tile origin, last = tile end (at most 5 entries each: virtual boundaries are at least one
// CTU apart).  Returns nx | ny << 4 | touch << 8, touch = a virtual boundary lies on or inside the tile.
__device__ __noinline__ int saTileCuts(const VbDev* v, int ctu, int x0, int y0, int tileW, int tileH, int* xs, int* ys)
{
  int touch = 0, n[2];
#pragma unroll 1
  for (int dir = 0; dir < 2; dir++)
  {
    int* o = dir ? ys : xs;
    const int lo = dir ? y0 : x0, hi = lo + (dir ? tileH : tileW), nb = dir ? v->nh : v->nv;
    int m = 0;
    o[m++] = lo;
    if (ctu < SA_T && lo + ctu < hi) o[m++] = lo + ctu;
#pragma unroll 1
    for (int i = 0; i < nb; i++)
    {
      const int b = dir ? v->y[i] : v->x[i];
      touch |= b >= lo && b <= hi;
      if (b <= lo || b >= hi) continue;
      bool dup = false;
      for (int k = 0; k < m; k++) dup |= o[k] == b;
      if (dup || m >= 4) continue;
      int k = m++;
      while (k > 0 && o[k - 1] > b) { o[k] = o[k - 1]; k--; }      // insertion keeps the list sorted
      o[k] = b;
    }
    o[m++] = hi;
    n[dir] = m - 1;
  }
  return n[0] | n[1] << 4 | touch << 8;
}

// a virtual boundary on an edge of the part (tile edge, CTU edge or the cut) closes that side of the clamp window win = {xlo, xhi,
// ylo, yhi}; the raster-slice corner pads are dropped when a virtual boundary clips that corner's sides
// (isCrossedByVirtualBoundaries :178-200).  Returns the pad flags that remain.
__device__ __noinline__ int saVbWindow(const VbDev* v, int4 part, int4 ctu, int pad, int* win)
{
  bool cL = false, cR = false, cT = false, cB = false;
#pragma unroll 1
  for (int i = 0; i < v->nv; i++)
  {
    const int b = v->x[i];
    if (b == part.x) win[0] = b;
    if (b == part.y) win[1] = b;
    cL |= b == ctu.x; cR |= b == ctu.y;
  }
#pragma unroll 1
  for (int i = 0; i < v->nh; i++)
  {
    const int b = v->y[i];
    if (b == part.z) win[2] = b;
    if (b == part.w) win[3] = b;
    cT |= b == ctu.z; cB |= b == ctu.w;
  }
  if (cL || cT) pad &= ~VTMGPU_ALF_PAD_TL;
  if (cR || cB) pad &= ~VTMGPU_ALF_PAD_BR;
  return pad;
}

// ---- ALF: generic scalar routines (virtual-boundary rows, halo cells, wide coefficients, non-4:2:0 CC-ALF) ----------
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

// Laplacian sums of the 2x2 cell whose top-left sample is p0 (picture row r): positions (r,c) and (r+1,c+1);
// rows beyond the virtual boundary are replaced (deriveClassificationBlk :906-915)
__device__ __noinline__ uint2 alfCellGeneric(const pel* p0, int r, int ctuMask, int vbL)
{
  int up = -SA_P, dn2 = 2 * SA_P;                                             // row r-1, row r+2
  const int rv = r & ctuMask;
  if (r > 0 && rv == vbL - 2) dn2 = SA_P;
  else if (r > 0 && rv == vbL) up = 0;
  const int y0v = p0[0] << 1, y1v = p0[SA_P + 1] << 1;
  const int v = iabs(y0v - p0[up] - p0[SA_P]) + iabs(y1v - p0[1] - p0[dn2 + 1]);
  const int h = iabs(y0v - p0[1] - p0[-1]) + iabs(y1v - p0[SA_P + 2] - p0[SA_P]);
  const int d0 = iabs(y0v - p0[up - 1] - p0[SA_P + 1]) + iabs(y1v - p0[0] - p0[dn2 + 2]);
  const int d1 = iabs(y0v - p0[SA_P - 1] - p0[up + 1]) + iabs(y1v - p0[dn2] - p0[2]);
  return make_uint2((uint32_t)v | (uint32_t)h << 16, (uint32_t)d0 | (uint32_t)d1 << 16);
}

// class index and transpose index of a 4x4 block from the four direction sums (deriveClassificationBlk :1012-1075)
__device__ __forceinline__ void alfClassify(int sumV, int sumH, int sumD0, int sumD1, int scale, int bd, int& cls, int& tIdx)
{
  const int act = clip3(0, 15, ((sumV + sumH) * scale) >> (bd + 4));
  cls = act == 0 ? 0 : (act == 1 ? 1 : (act < 7 ? 2 : (act < 15 ? 3 : 4)));      // th[] = {0,1,2,2,2,2,2,3,3,3,3,3,3,3,3,4}
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
  tIdx = (0x31322010 >> (4 * (mainDir * 2 + (secDir >> 1)))) & 0xf;              // transposeTable = {0,1,0,2,2,3,1,3}
}


// one 4x4 luma block, any row position: classification incl. the virtual-boundary rules and the 7x7 filter with row clamping
__device__ __noinline__ void alfLumaBlockGeneric(const uint2 (*cell)[SA_CELLP], const pel* c0, pel* out, int pitchOut, int bi, int bj, int by,
                                                 const short2* set, int ctuMask, int vbL, int bd)
{
  const int yb = by & ctuMask;
  const int i0 = (yb == vbL) ? 1 : 0, i1 = (yb == vbL - 4) ? 3 : 4;
  int sum[4] = { 0, 0, 0, 0 };
  for (int i = i0; i < i1; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
      const uint2 c = cell[2 * bi + i][2 * bj + j];
      sum[0] += c.x & 0xffff; sum[1] += c.x >> 16; sum[2] += c.y & 0xffff; sum[3] += c.y >> 16;
    }
  int cls, tIdx;
  alfClassify(sum[0], sum[1], sum[2], sum[3], (yb == vbL - 4 || yb == vbL) ? 96 : 64, bd, cls, tIdx);
  short2 f[12];
#pragma unroll
  for (int k = 0; k < 12; k++) f[k] = __ldg(&set[cls * 12 + c_perm7[tIdx][k]]);
  const int maxv = (1 << bd) - 1;
  for (int r = 0; r < 4; r++)
  {
    int lim; bool nearVb;
    vbLimit((by + r) & ctuMask, vbL, 4, lim, nearVb);
    const int o1 = min(1, lim) * SA_P, o2 = min(2, lim) * SA_P, o3 = lim * SA_P;   // lim <= 3
    int res[4];
#pragma unroll
    for (int q = 0; q < 4; q++)
    {
      const pel* c = c0 + r * SA_P + q;
      const int cur = c[0];
      int s = alfTap(c, o3, cur, f[0]) + alfTap(c, o2 + 1, cur, f[1]) + alfTap(c, o2, cur, f[2]) + alfTap(c, o2 - 1, cur, f[3]) +
              alfTap(c, o1 + 2, cur, f[4]) + alfTap(c, o1 + 1, cur, f[5]) + alfTap(c, o1, cur, f[6]) + alfTap(c, o1 - 1, cur, f[7]) +
              alfTap(c, o1 - 2, cur, f[8]) + alfTap(c, 3, cur, f[9]) + alfTap(c, 2, cur, f[10]) + alfTap(c, 1, cur, f[11]);
      s = (s + 64) >> (nearVb ? 10 : 7);
      res[q] = clip3(0, maxv, cur + s);
    }
    *reinterpret_cast<uint2*>(out + (size_t)r * pitchOut) = make_uint2((uint32_t)res[0] | (uint32_t)res[1] << 16, (uint32_t)res[2] | (uint32_t)res[3] << 16);
  }
}

// ---- ALF luma: packed fast path (block rows that do not touch a virtual boundary) -------------------------------------
// Laplacian cells of the block's own 4x4 samples (2x2 cells) from the 6x6 sample patch around it
__device__ __forceinline__ void alfOwnCells(uint2 (*cell)[SA_CELLP], const pel* c0, int bi, int bj)
{
  int p[6][6];                                               // p[r][c] = sample (r-1, c-1) relative to the block
#pragma unroll
  for (int r = 0; r < 6; r++)
  {
    const pel* row = c0 + (r - 1) * SA_P;
    const uint2 q0 = *reinterpret_cast<const uint2*>(row - 4), q1 = *reinterpret_cast<const uint2*>(row);
    const uint32_t w4 = *reinterpret_cast<const uint32_t*>(row + 4);
    p[r][0] = q0.y >> 16; p[r][1] = q1.x & 0xffff; p[r][2] = q1.x >> 16; p[r][3] = q1.y & 0xffff; p[r][4] = q1.y >> 16; p[r][5] = w4 & 0xffff;
  }
#pragma unroll
  for (int cy = 0; cy < 2; cy++)
#pragma unroll
    for (int cx = 0; cx < 2; cx++)
    {
      const int r = 1 + 2 * cy, c = 1 + 2 * cx;
      const int a2 = 2 * p[r][c], b2 = 2 * p[r + 1][c + 1];
      const int v = iabs(a2 - p[r - 1][c] - p[r + 1][c]) + iabs(b2 - p[r][c + 1] - p[r + 2][c + 1]);
      const int h = iabs(a2 - p[r][c - 1] - p[r][c + 1]) + iabs(b2 - p[r + 1][c] - p[r + 1][c + 2]);
      const int d0 = iabs(a2 - p[r - 1][c - 1] - p[r + 1][c + 1]) + iabs(b2 - p[r][c] - p[r + 2][c + 2]);
      const int d1 = iabs(a2 - p[r + 1][c - 1] - p[r - 1][c + 1]) + iabs(b2 - p[r + 2][c] - p[r][c + 2]);
      cell[2 * bi + 1 + cy][2 * bj + 1 + cx] = make_uint2((uint32_t)v | (uint32_t)h << 16, (uint32_t)d0 | (uint32_t)d1 << 16);
    }
}

// 7x7 diamond on one 4x4 block, two pixels per register.  e = pre-expanded {coefficient, clip} entry of the block's
// (filter set, class, transpose).  clamp(n - cur, -c, c) + c  ==  max(min(n + (c - cur), 2c), 0)  is ONE
// instruction; the excess sum(coef * 2c) is folded into e->bias together with the rounding offset 64 (filterBlk :1249-1297).
//
// The four output rows are a REAL loop (not unrolled): both kernels of the chain are bound by instruction fetch (the SM
// instruction cache misses into the GPC-level cache, profiles/: gcc instruction requests at 75 % of peak with the fully
// unrolled version), so the body is kept small (~200 instructions) and each row re-reads its 7 input rows from shared
// memory (68 LDS.64 per block instead of 30).  With the rows addressed at run time the virtual-boundary clamping of
// filterBlk :1227-1247 (row offset min(|dy|, lim), >> 10 on the row next to the boundary) needs no separate code:
// vb = 0 away from the boundary, 1 / 2 for the block directly above / below it.
__device__ __forceinline__ void alfLumaBlockFast(const pel* c0, pel* out, int pitchOut, const AlfLumaEntry* __restrict__ e, uint32_t maxvP, int vb)
{
#define ALF_LOAD_ENTRY(E)                                                                                               \
  {                                                                                                                     \
    const uint4* q = reinterpret_cast<const uint4*>(E);                                                                 \
    _Pragma("unroll") for (int i = 0; i < 3; i++)                                                                       \
    {                                                                                                                   \
      const uint4 a = q[i], b = q[3 + i], c = q[6 + i];                  /* plain loads: the entry may sit in shared memory */ \
      coefB[4 * i] = a.x; coefB[4 * i + 1] = a.y; coefB[4 * i + 2] = a.z; coefB[4 * i + 3] = a.w;                       \
      clipP1[4 * i] = b.x; clipP1[4 * i + 1] = b.y; clipP1[4 * i + 2] = b.z; clipP1[4 * i + 3] = b.w;                   \
      clip2[4 * i] = c.x; clip2[4 * i + 1] = c.y; clip2[4 * i + 2] = c.z; clip2[4 * i + 3] = c.w;                       \
    }                                                                                                                   \
    bias = (E)->bias;                                                                                                   \
  }
  uint32_t coefB[12], clipP1[12], clip2[12];
  int bias;
  ALF_LOAD_ENTRY(e)
#pragma unroll 1
  for (int orow = 0; orow < 4; orow++)
  {
    const int lim = vb == 0 ? 3 : (vb == 1 ? 3 - orow : orow);                 // rows available on this side of the virtual boundary
    const int sh = (vb == 1 && orow == 3) || (vb == 2 && orow == 0) ? 10 : 7;
    const int o1 = min(1, lim) * SA_P, o2 = min(2, lim) * SA_P, o3 = lim * SA_P;
    const pel* rc = c0 + orow * SA_P;
    // w*[j] = samples (4bj - 4 + 2j, +1) of a row ; o*[j] = the pair starting one sample later
    uint32_t wc[6], oc[5], wp1[6], op1[5], wm1[6], om1[5], wp2[6], op2[5], wm2[6], om2[5], wp3[2], wm3[2];
#define ALF_LOADROW(W, O, PTR)                                                                                          \
    {                                                                                                                   \
      const uint2* rp_ = reinterpret_cast<const uint2*>((PTR) - 4);                                                     \
      const uint2 q0 = rp_[0], q1 = rp_[1], q2 = rp_[2];                                                                \
      W[0] = q0.x; W[1] = q0.y; W[2] = q1.x; W[3] = q1.y; W[4] = q2.x; W[5] = q2.y;                                     \
      _Pragma("unroll") for (int j = 0; j < 5; j++) O[j] = mid16(W[j], W[j + 1]);                                       \
    }
    ALF_LOADROW(wc, oc, rc)
    ALF_LOADROW(wp1, op1, rc + o1) ALF_LOADROW(wm1, om1, rc - o1)
    ALF_LOADROW(wp2, op2, rc + o2) ALF_LOADROW(wm2, om2, rc - o2)
    { const uint2 q = *reinterpret_cast<const uint2*>(rc + o3); wp3[0] = q.x; wp3[1] = q.y; }
    { const uint2 q = *reinterpret_cast<const uint2*>(rc - o3); wm3[0] = q.x; wm3[1] = q.y; }
#undef ALF_LOADROW
    // pair of a row that starts at sample C relative to the block (C = 2 px + dx): even C -> W[(C+4)/2], odd C -> O[(C+3)/2]
#define ALF_PAIR(W, O, C) ((((C) & 1) != 0) ? O[((C) + 3) >> 1] : W[((C) + 4) >> 1])
#define ALF_TAP(K, DX, NP, NM)                                                                                          \
    {                                                                                                                   \
      const uint32_t cb = __vadd2(clipP1[K], ncur);                                                                     \
      const uint32_t s = addClamp0(NP, cb, clip2[K]) + addClamp0(NM, cb, clip2[K]);                                     \
      acc0 = __dp2a_lo((int)s, (int)coefB[K], acc0);                                                                    \
      acc1 = __dp2a_hi((int)s, (int)coefB[K], acc1);                                                                    \
    }
    uint32_t res[2];
#pragma unroll
    for (int px = 0; px < 2; px++)
    {
      const uint32_t cur = wc[2 + px], ncur = ~cur;                            // clipP1 + ~cur = clip - cur per lane
      int acc0 = bias, acc1 = bias;
      ALF_TAP(0, 0, wp3[px], wm3[px])
      ALF_TAP(1, 1, ALF_PAIR(wp2, op2, 2 * px + 1), ALF_PAIR(wm2, om2, 2 * px - 1))
      ALF_TAP(2, 0, ALF_PAIR(wp2, op2, 2 * px), ALF_PAIR(wm2, om2, 2 * px))
      ALF_TAP(3, -1, ALF_PAIR(wp2, op2, 2 * px - 1), ALF_PAIR(wm2, om2, 2 * px + 1))
      ALF_TAP(4, 2, ALF_PAIR(wp1, op1, 2 * px + 2), ALF_PAIR(wm1, om1, 2 * px - 2))
      ALF_TAP(5, 1, ALF_PAIR(wp1, op1, 2 * px + 1), ALF_PAIR(wm1, om1, 2 * px - 1))
      ALF_TAP(6, 0, ALF_PAIR(wp1, op1, 2 * px), ALF_PAIR(wm1, om1, 2 * px))
      ALF_TAP(7, -1, ALF_PAIR(wp1, op1, 2 * px - 1), ALF_PAIR(wm1, om1, 2 * px + 1))
      ALF_TAP(8, -2, ALF_PAIR(wp1, op1, 2 * px - 2), ALF_PAIR(wm1, om1, 2 * px + 2))
      ALF_TAP(9, 3, ALF_PAIR(wc, oc, 2 * px + 3), ALF_PAIR(wc, oc, 2 * px - 3))
      ALF_TAP(10, 2, ALF_PAIR(wc, oc, 2 * px + 2), ALF_PAIR(wc, oc, 2 * px - 2))
      ALF_TAP(11, 1, ALF_PAIR(wc, oc, 2 * px + 1), ALF_PAIR(wc, oc, 2 * px - 1))
      res[px] = addClamp0(cur, prmt((uint32_t)(acc0 >> sh), (uint32_t)(acc1 >> sh), 0x5410u), maxvP);
    }
    *reinterpret_cast<uint2*>(out + (size_t)orow * pitchOut) = make_uint2(res[0], res[1]);
#undef ALF_TAP
#undef ALF_PAIR
  }
#undef ALF_LOAD_ENTRY
}

// ---- chroma 5x5 + CC-ALF, packed (four horizontally adjacent chroma samples per thread) ----------------------------
struct ChromaCoef
{
  uint32_t coefB[6], clipP1[6], clip2[6];
  int bias;
};

__device__ __forceinline__ ChromaCoef chromaCoef(const AlfChromaEntry* __restrict__ e)
{
  ChromaCoef c;
  const uint4* q = reinterpret_cast<const uint4*>(e);
  const uint4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];          // plain loads: the table may sit in shared memory
  c.coefB[0] = q0.x; c.coefB[1] = q0.y; c.coefB[2] = q0.z; c.coefB[3] = q0.w; c.coefB[4] = q1.x; c.coefB[5] = q1.y;
  c.clipP1[0] = q1.z; c.clipP1[1] = q1.w; c.clipP1[2] = q2.x; c.clipP1[3] = q2.y; c.clipP1[4] = q2.z; c.clipP1[5] = q2.w;
  c.clip2[0] = q3.x; c.clip2[1] = q3.y; c.clip2[2] = q3.z; c.clip2[3] = q3.w; c.clip2[4] = q4.x; c.clip2[5] = q4.y;
  c.bias = (int)q4.z;
  return c;
}

// ALF of the 4 chroma samples at cb (smem, SAO output); o1/o2 = row offsets after virtual-boundary clamping
__device__ __forceinline__ uint2 alfChromaQuad(const pel* cb, int o1, int o2, bool nearVb, const ChromaCoef& C, uint32_t maxvP)
{
  // words of a row: wm2 = samples (-2,-1), w0 = (0,1), w2 = (2,3), w4 = (4,5)
  uint32_t res[2];
  const uint2 c0q = *reinterpret_cast<const uint2*>(cb - 4), c0r = *reinterpret_cast<const uint2*>(cb);
  const uint32_t c0w4 = *reinterpret_cast<const uint32_t*>(cb + 4);
  const uint32_t r0[4] = { c0q.y, c0r.x, c0r.y, c0w4 };                                     // centre row: wm2, w0, w2, w4
  const uint32_t r0o[3] = { mid16(r0[0], r0[1]), mid16(r0[1], r0[2]), mid16(r0[2], r0[3]) };    // pairs starting at -1, 1, 3
  uint32_t pe[2][2], po[2][3];                                                              // rows +o1 / -o1: even words w0,w2 ; odd pairs
#pragma unroll
  for (int s = 0; s < 2; s++)
  {
    const pel* rp = cb + (s ? -o1 : o1);
    const uint2 q = *reinterpret_cast<const uint2*>(rp - 4), r = *reinterpret_cast<const uint2*>(rp);
    const uint32_t w4 = *reinterpret_cast<const uint32_t*>(rp + 4);
    pe[s][0] = r.x; pe[s][1] = r.y;
    po[s][0] = mid16(q.y, r.x); po[s][1] = mid16(r.x, r.y); po[s][2] = mid16(r.y, w4);
  }
  const uint2 p2 = *reinterpret_cast<const uint2*>(cb + o2), m2 = *reinterpret_cast<const uint2*>(cb - o2);
  const uint32_t p2w[2] = { p2.x, p2.y }, m2w[2] = { m2.x, m2.y };
#pragma unroll
  for (int px = 0; px < 2; px++)
  {
    const uint32_t cur = r0[1 + px], ncur = ~cur;
    int acc0 = C.bias, acc1 = C.bias;
#define CH_TAP(K, NP, NM)                                                                       \
    {                                                                                           \
      const uint32_t cbv = __vadd2(C.clipP1[K], ncur);                                          \
      const uint32_t s = addClamp0(NP, cbv, C.clip2[K]) + addClamp0(NM, cbv, C.clip2[K]);       \
      acc0 = __dp2a_lo((int)s, (int)C.coefB[K], acc0);                                          \
      acc1 = __dp2a_hi((int)s, (int)C.coefB[K], acc1);                                          \
    }
    CH_TAP(0, p2w[px], m2w[px])                       // (0,+2) / (0,-2)
    CH_TAP(1, po[0][px + 1], po[1][px])               // (+1,+1) / (-1,-1)
    CH_TAP(2, pe[0][px], pe[1][px])                   // (0,+1) / (0,-1)
    CH_TAP(3, po[0][px], po[1][px + 1])               // (-1,+1) / (+1,-1)
    CH_TAP(4, r0[2 + px], r0[px])                     // (+2,0) / (-2,0)
    CH_TAP(5, r0o[px + 1], r0o[px])                   // (+1,0) / (-1,0)
#undef CH_TAP
    const int sh = nearVb ? 10 : 7;
    res[px] = addClamp0(cur, prmt((uint32_t)(acc0 >> sh), (uint32_t)(acc1 >> sh), 0x5410u), maxvP);
  }
  return make_uint2(res[0], res[1]);
}

// CC-ALF correction for 4 chroma samples in 4:2:0 / 4:2:2 (sx = 1): collocated luma column = 2 * chroma column.
// l = luma sample collocated with the first chroma sample (smem, SAO output); l1,l2,l3 = row offsets (filterBlkCcAlf
// :1376-1386).  Returns the packed corrections (already clipped to the chroma range around 0).
__device__ __forceinline__ uint2 ccAlfQuad420(const pel* l, int l1, int l2, int l3, const uint32_t* __restrict__ ccw, uint32_t maxcP, uint32_t halfP)
{
  // luma words of a row relative to l: W(-2), W(0), W(2), W(4), W(6)
  uint32_t ctr[2], up[2], lf[2], rt[2], dl[2], dm[2], dr[2], d2[2];
  {
    const uint2 q = *reinterpret_cast<const uint2*>(l - 4), r = *reinterpret_cast<const uint2*>(l), t = *reinterpret_cast<const uint2*>(l + 4);
    // even samples (0,2),(4,6) ; odd samples (-1,1),(3,5) and (1,3),(5,7)
    ctr[0] = prmt(r.x, r.y, 0x5410u); ctr[1] = prmt(t.x, t.y, 0x5410u);
    lf[0] = prmt(q.y, r.x, 0x7632u); lf[1] = prmt(r.y, t.x, 0x7632u);
    rt[0] = prmt(r.x, r.y, 0x7632u); rt[1] = prmt(t.x, t.y, 0x7632u);
  }
  {
    const uint2 r = *reinterpret_cast<const uint2*>(l + l2), t = *reinterpret_cast<const uint2*>(l + l2 + 4);
    up[0] = prmt(r.x, r.y, 0x5410u); up[1] = prmt(t.x, t.y, 0x5410u);
  }
  {
    const uint2 q = *reinterpret_cast<const uint2*>(l + l1 - 4), r = *reinterpret_cast<const uint2*>(l + l1), t = *reinterpret_cast<const uint2*>(l + l1 + 4);
    dm[0] = prmt(r.x, r.y, 0x5410u); dm[1] = prmt(t.x, t.y, 0x5410u);
    dl[0] = prmt(q.y, r.x, 0x7632u); dl[1] = prmt(r.y, t.x, 0x7632u);
    dr[0] = prmt(r.x, r.y, 0x7632u); dr[1] = prmt(t.x, t.y, 0x7632u);
  }
  {
    const uint2 r = *reinterpret_cast<const uint2*>(l + l3), t = *reinterpret_cast<const uint2*>(l + l3 + 4);
    d2[0] = prmt(r.x, r.y, 0x5410u); d2[1] = prmt(t.x, t.y, 0x5410u);
  }
  // coefficients as IDP.2A byte operands + their sum, expanded on the host (AlfDev::ccB)
  const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(ccw)), w1 = __ldg(reinterpret_cast<const uint4*>(ccw) + 1);
  const uint32_t cB[7] = { w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z };
  const int fsum = (int)w1.w;
  uint32_t res[2];
#pragma unroll
  for (int px = 0; px < 2; px++)
  {
    int a0 = 64 - (int)(ctr[px] & 0xffff) * fsum, a1 = 64 - (int)(ctr[px] >> 16) * fsum;
#define CC_TAP(K, N) a0 = __dp2a_lo((int)(N), (int)cB[K], a0); a1 = __dp2a_hi((int)(N), (int)cB[K], a1);
    CC_TAP(0, up[px]) CC_TAP(1, lf[px]) CC_TAP(2, rt[px]) CC_TAP(3, dl[px]) CC_TAP(4, dm[px]) CC_TAP(5, dr[px]) CC_TAP(6, d2[px])
#undef CC_TAP
    const uint32_t t = addClamp0(prmt((uint32_t)(a0 >> 7), (uint32_t)(a1 >> 7), 0x5410u), halfP, maxcP);     // ClipPel(sum + half)
    res[px] = __vadd2(t, ~halfP + 0x00010001u);                                                                // - half
  }
  return make_uint2(res[0], res[1]);
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
struct SaWalk            // position of a CTA in its round-robin walk over the tiles of a batch of picture slots
{
  int slot, ty, tx;
};

struct SaStep            // one step of the walk = gridDim.x tiles, decomposed on the host (no division in the loop)
{
  int dx, dy, ds;
};

__device__ __forceinline__ void saAdvance(SaWalk& p, const SaStep& st, int tilesX, int tilesY)
{
  p.tx += st.dx; if (p.tx >= tilesX) { p.tx -= tilesX; p.ty++; }
  p.ty += st.dy; if (p.ty >= tilesY) { p.ty -= tilesY; p.slot++; }
  p.slot += st.ds;
}

// issues the asynchronous loads of one tile into `stage`: three TMA boxes (one thread) + the control record of its CTU
__device__ __forceinline__ void saPrefetch(unsigned char* smraw, const SaLayout& L, int stage, const SlotDev& S, const CUtensorMap* maps, const SaWalk& p,
                                           const Geom& g, int ty0, bool ctlToShared)
{
  const int tid = threadIdx.x, x0 = p.tx * SA_T, y0 = (p.ty + ty0) * SA_TH;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smraw + L.offBar) + stage;
  if (tid == 0)
  {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic-proxy accesses to the stage buffers are ordered before the TMA writes
    mbarExpectTx(bar, (uint32_t)(L.lumaBytes + 2 * L.chromaBytes));
    tmaLoad2D(smraw + L.offA(stage, 0), maps, x0 - SA_HX, y0 - SA_HY, bar);
    if (g.ncomp > 1)
    {
      tmaLoad2D(smraw + L.offA(stage, 1), maps + 1, (x0 >> g.sx) - SA_HX, (y0 >> g.sy) - SA_HY, bar);
      tmaLoad2D(smraw + L.offA(stage, 2), maps + 2, (x0 >> g.sx) - SA_HX, (y0 >> g.sy) - SA_HY, bar);
    }
  }
  // control record(s) of the CTU(s) under the tile: one for CTU sizes >= 64, 2 x 2 for CTU size 32 (record k = CTU (k & 1, k >> 1))
  if (ctlToShared && tid >= 32 && tid < 36)
  {
    const int k = tid - 32, cx = (x0 >> g.ctuLog2) + (k & 1), cy = (y0 >> g.ctuLog2) + (k >> 1);
    if ((k == 0 || g.ctu < SA_T) && cx < g.wCtus && cy < g.hCtus)
      cpAsync16(reinterpret_cast<CtuCtlDev*>(smraw + L.offPar) + stage * 4 + k, &S.ctuCtl[cy * g.wCtus + cx]);
  }
}

// kVirtualBoundaries = false compiles the part loop of signalled virtual boundaries away (it costs the common path 10 % otherwise);
// the host launches the <true> instantiation only for batches that contain a picture with virtual boundaries.
// Persistent kernel: gridDim.x CTAs walk the tiles of slots [firstSlot, firstSlot + numSlots) round robin; while a CTA
// filters tile i, the TMA loads of tile i+1 are in flight (two stages).  maps = tensor maps of the source buffer of the
// first slot: [slot][3 buffers][3 planes].
__global__ void __launch_bounds__(SA_THREADS, SA_CTAS_PER_SM) k_alf_parts(const SlotDev* __restrict__ slots, const CUtensorMap* __restrict__ tmaps, int firstSlot, int numSlots,
                                                           int srcBuf, int dstBuf, Geom g, int tilesX, int tilesY, int ty0, SaStep step)
{
  constexpr bool kVirtualBoundaries = true;
  extern __shared__ __align__(128) unsigned char smraw[];
  const SaLayout L = saLayout(g.sx, g.sy, g.ncomp);
  const int tid = threadIdx.x;
  uint2 (*cell)[SA_CELLP] = reinterpret_cast<uint2 (*)[SA_CELLP]>(smraw + L.offCell);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + L.offBar);
  const int vbL = g.ctu - 4, ctuMask = g.ctu - 1;
  const int bi = tid >> 4, bj = tid & 15;
  const int tw = SA_T >> g.sx, th = SA_TH >> g.sy, thLogC = SA_THLOG - g.sy;
  const int ctuH = g.ctu >> g.sy;
  const int vbC = ctuH - 2, maxc = (1 << g.bdC) - 1, half = (1 << g.bdC) >> 1;
  const uint32_t maxcP = dup16(maxc), halfP = dup16(half);

  SaWalk cur;
  {
    const int tilesPerPic = tilesX * tilesY, t = blockIdx.x;
    cur.slot = t / tilesPerPic;
    const int rem = t - cur.slot * tilesPerPic;
    cur.ty = rem / tilesX;
    cur.tx = rem - cur.ty * tilesX;
  }
  if (cur.slot >= numSlots) return;
  if (tid == 0)
  {
    mbarInit(&bars[0], 1);
    mbarInit(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // k_alf<false>: one CTU per tile -- its control record travels through registers, loaded one tile ahead, and the tile needs no
  // CTA barrier between the arrival of its samples (every thread waits on the mbarrier itself) and the first phase
  uint4 ctlNext = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    const SlotDev& S = slots[firstSlot + cur.slot];
    saPrefetch(smraw, L, 0, S, tmaps + ((size_t)(firstSlot + cur.slot) * 3 + srcBuf) * 3, cur, g, ty0, kVirtualBoundaries);
    cpAsyncCommit();
    if (!kVirtualBoundaries) ctlNext = __ldg(reinterpret_cast<const uint4*>(&S.ctuCtl[(((cur.ty + ty0) * SA_TH) >> g.ctuLog2) * g.wCtus + ((cur.tx * SA_T) >> g.ctuLog2)]));
  }
  for (uint32_t it = 0; cur.slot < numSlots; it++)
  {
    const int stage = it & 1;
    const SlotDev& S = slots[firstSlot + cur.slot];
    SaWalk nxt = cur;
    saAdvance(nxt, step, tilesX, tilesY);
    if (nxt.slot < numSlots)
    {
      const SlotDev& Sn = slots[firstSlot + nxt.slot];
      saPrefetch(smraw, L, stage ^ 1, Sn, tmaps + ((size_t)(firstSlot + nxt.slot) * 3 + srcBuf) * 3, nxt, g, ty0, kVirtualBoundaries);
    }
    const uint4 ctlCur = ctlNext;
    if (kVirtualBoundaries)
    {
      cpAsyncCommit();
      cpAsyncWait<1>();
      mbarWait(&bars[stage], (it >> 1) & 1);
      __syncthreads();                                       // tile and its control records are in shared memory
    }
    else
    {
      if (nxt.slot < numSlots)
        ctlNext = __ldg(reinterpret_cast<const uint4*>(&slots[firstSlot + nxt.slot].ctuCtl[(((nxt.ty + ty0) * SA_TH) >> g.ctuLog2) * g.wCtus + ((nxt.tx * SA_T) >> g.ctuLog2)]));
      mbarWait(&bars[stage], (it >> 1) & 1);                 // the TMA writes of this tile are visible to this thread
    }

    const int x0 = cur.tx * SA_T, y0 = (cur.ty + ty0) * SA_TH;
    const bool alfOn = S.alfOn != 0;
    pel* const A0 = reinterpret_cast<pel*>(smraw + L.offA(stage, 0));
    pel* const A1 = reinterpret_cast<pel*>(smraw + L.offA(stage, 1));
    pel* const A2 = reinterpret_cast<pel*>(smraw + L.offA(stage, 2));
    // Parts of the tile.  Normally one: the whole tile.  Signalled virtual boundaries (vtmgpu_virtual_boundaries, multiples of 8
    // luma samples, at least a CTU apart) cut a tile into parts that the reference filters as separately padded blocks
    // (ALFProcess :452-490), and with CTU size 32 a tile covers 2 x 2 CTUs with their own control records and clip flags; each
    // part is then filtered from its own padded copy of the tile.
    const int tileW = min(SA_T, g.w - x0), tileH = min(SA_TH, g.h - y0);          // the last tile of a row / column may be partial
    const VbDev& pvb = S.vbAlf;
    const bool anyVb = kVirtualBoundaries && alfOn && (pvb.nv | pvb.nh) != 0;
    int xs[5], ys[5], nx = 1, ny = 1;
    bool vbTile = false;
    xs[0] = x0; xs[1] = x0 + tileW; ys[0] = y0; ys[1] = y0 + tileH;
    if (kVirtualBoundaries && (anyVb || g.ctu < SA_T))
    {
      const int r = saTileCuts(&pvb, g.ctu, x0, y0, tileW, tileH, xs, ys);
      nx = r & 15; ny = (r >> 4) & 15; vbTile = anyVb && (r >> 8) != 0;
    }
    const int nparts = nx * ny;
    const pel* const B0 = A0;
    const pel* const B1 = A1;
    const pel* const B2 = A2;
    const PlaneDev dstY = S.buf[dstBuf][0];
#pragma unroll 1
    for (int part = 0; part < nparts; part++)
    {
      const int pj = part / nx, pi = part - pj * nx;
      const int px0 = xs[pi], px1 = xs[pi + 1], py0 = ys[pj], py1 = ys[pj + 1];       // luma rectangle of this part
      // control record of the part's CTU
      const int cidx = (kVirtualBoundaries && g.ctu < SA_T) ? (((px0 - x0) >> g.ctuLog2) | ((py0 - y0) >> g.ctuLog2) << 1) : 0;
      CtuCtlDev ctl;
      if (kVirtualBoundaries) ctl = reinterpret_cast<const CtuCtlDev*>(smraw + L.offPar)[stage * 4 + cidx];
      else                    *reinterpret_cast<uint4*>(&ctl) = ctlCur;
      const bool ctuOn = alfOn && (ctl.flags & 1) != 0;          // the CTU's own slice runs ALF (ALFProcess :429)
      const bool alfY = ctuOn && ctl.enY != 0, alfCb = ctuOn && ctl.enCb != 0, alfCr = ctuOn && ctl.enCr != 0;
      const int ccCb = ctuOn ? ctl.ccCb : 0, ccCr = ctuOn ? ctl.ccCr : 0;
      const int clip = ctl.clip;
      const AlfDev* const alfG = S.alf + ctl.grp;                // chroma / CC-ALF data of the CTU's slice (the luma sets are S.alf[0]'s)
      const int cx0 = px0 & ~ctuMask, cy0 = py0 & ~ctuMask, cx1 = min(cx0 + g.ctu, g.w), cy1 = min(cy0 + g.ctu, g.h);
      if (nparts > 1)
      {
        // the padding of a part overwrites its neighbours' samples: the first part saves the tile, the others start from that copy
        uint4* stg = reinterpret_cast<uint4*>(smraw + L.offA(stage, 0));
        uint4* scr = reinterpret_cast<uint4*>(smraw + L.total);
        const int n16 = (L.lumaBytes + 2 * L.chromaBytes) >> 4;
        if (part == 0) { for (int i = tid; i < n16; i += SA_THREADS) scr[i] = stg[i]; }
        else           { for (int i = tid; i < n16; i += SA_THREADS) stg[i] = scr[i]; }
        __syncthreads();
      }
      // tiles on the picture border: replicate the border samples into the zero-filled outside
      // (= UnitBuf::extendBorderPel of the ALF input, AdaptiveLoopFilter.cpp:411); parts of a CTU with a slice / tile / virtual
      // boundary the filter must not read across: the same replication at the clipped sides (:452-490)
      const bool onBorder = x0 == 0 || y0 == 0 || x0 + SA_T + 8 > g.w || y0 + SA_TH + 8 > g.h || clip != 0 || vbTile;
      if (onBorder)
      {
        int win[4] = { (clip & VTMGPU_ALF_CLIP_LEFT) ? cx0 : 0, (clip & VTMGPU_ALF_CLIP_RIGHT) ? cx1 : g.w,
                       (clip & VTMGPU_ALF_CLIP_TOP) ? cy0 : 0, (clip & VTMGPU_ALF_CLIP_BOTTOM) ? cy1 : g.h };
        int pad = clip & (VTMGPU_ALF_PAD_TL | VTMGPU_ALF_PAD_BR);      // only for the part that holds that corner of the CTU
        if (px0 != cx0 || py0 != cy0) pad &= ~VTMGPU_ALF_PAD_TL;
        if (px1 != cx1 || py1 != cy1) pad &= ~VTMGPU_ALF_PAD_BR;
        if (anyVb) pad = saVbWindow(&pvb, make_int4(px0, px1, py0, py1), make_int4(cx0, cx1, cy0, cy1), pad, win);
        const int xlo = win[0], xhi = win[1], ylo = win[2], yhi = win[3];
        const int bxc = (x0 >> g.sx) - SA_HX, byc = (y0 >> g.sy) - SA_HY;
        if (alfY || ccCb || ccCr) saReplicateBorder(A0, x0 - SA_HX, y0 - SA_HY, xlo, xhi, ylo, yhi, SA_P, SA_T, SA_TH, 3);
        if (alfCb) saReplicateBorder(A1, bxc, byc, xlo >> g.sx, xhi >> g.sx, ylo >> g.sy, yhi >> g.sy, L.pitchC, tw, th, 3);
        if (alfCr) saReplicateBorder(A2, bxc, byc, xlo >> g.sx, xhi >> g.sx, ylo >> g.sy, yhi >> g.sy, L.pitchC, tw, th, 3);
        if (pad)
        {
          // no barrier needed in between: the corners lie outside the clamp window's replicated ranges (their sides are not clipped)
          if (alfY || ccCb || ccCr) saPadCorners(A0, x0 - SA_HX, y0 - SA_HY, cx0, cy0, cx1, cy1, SA_P, 3, pad);
          if (alfCb) saPadCorners(A1, bxc, byc, cx0 >> g.sx, cy0 >> g.sy, cx1 >> g.sx, cy1 >> g.sy, L.pitchC, 3, pad);
          if (alfCr) saPadCorners(A2, bxc, byc, cx0 >> g.sx, cy0 >> g.sy, cx1 >> g.sx, cy1 >> g.sy, L.pitchC, 3, pad);
        }
        __syncthreads();
      }

      // ---- phase 1: Laplacian cells (luma ALF only) -------------------------------------------------------------------
      const pel* lumaB = B0;
      const int bx = x0 + 4 * bj, by = y0 + 4 * bi;
      const pel* c0 = &lumaB[(4 * bi + SA_HY) * SA_P + 4 * bj + SA_HX];
      const int yb = by & ctuMask;
      const bool vbBlk = yb == vbL - 4 || yb == vbL;             // uniform per warp (two block rows) for CTU sizes >= 32
      if (alfY)
      {
        if (!vbBlk) alfOwnCells(cell, c0, bi, bj);
        else
        {
          for (int k = 0; k < 4; k++)
          {
            const int li = 2 * bi + 1 + (k >> 1), lj = 2 * bj + 1 + (k & 1);
            cell[li][lj] = alfCellGeneric(&lumaB[(2 * li + SA_HY - 2) * SA_P + 2 * lj + SA_HX - 2], y0 - 2 + 2 * li, ctuMask, vbL);
          }
        }
        // ring of halo cells: top and bottom cell rows, left and right cell columns
        for (int q = tid; q < 2 * SA_CELLS + 2 * (SA_CELLR - 2); q += SA_THREADS)
        {
          int li, lj;
          if (q < 2 * SA_CELLS) { li = q < SA_CELLS ? 0 : SA_CELLR - 1; lj = q < SA_CELLS ? q : q - SA_CELLS; }
          else { const int k = q - 2 * SA_CELLS; li = 1 + (k >> 1); lj = (k & 1) ? SA_CELLS - 1 : 0; }
          cell[li][lj] = alfCellGeneric(&lumaB[(2 * li + SA_HY - 2) * SA_P + 2 * lj + SA_HX - 2], y0 - 2 + 2 * li, ctuMask, vbL);
        }
        __syncthreads();
      }

      // ---- phase 2: filters; every plane is written once ----------------------------------------------------------------
      if (alfY)
      {
        if (bx >= px0 && bx < px1 && by >= py0 && by < py1)
        {
          pel* out = dstY.p + (size_t)by * dstY.pitch + bx;
          const int setIdx = ctl.setIdx;
          if (S.alfWide) alfLumaBlockGeneric(cell, c0, out, dstY.pitch, bi, bj, by, &S.alf->luma[setIdx][0][0], ctuMask, vbL, g.bdL);
          else
          {
            // window = cells (2bi .. 2bi+3) x (2bj .. 2bj+3); row sums stay below 2^16 per lane for any bit depth <= 12.
            // Blocks at the virtual boundary use 3 of the 4 cell rows and the scale 96 (deriveClassificationBlk :977-1010)
            const int vb = yb == vbL - 4 ? 1 : (yb == vbL ? 2 : 0);
            int sumV = 0, sumH = 0, sumD0 = 0, sumD1 = 0;
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
              if ((vb == 1 && i == 3) || (vb == 2 && i == 0)) continue;
              const uint4* rp = reinterpret_cast<const uint4*>(&cell[2 * bi + i][2 * bj]);
              const uint4 q0 = rp[0], q1 = rp[1];
              const uint32_t vh = q0.x + q0.z + q1.x + q1.z, dd = q0.y + q0.w + q1.y + q1.w;
              sumV += vh & 0xffff; sumH += vh >> 16; sumD0 += dd & 0xffff; sumD1 += dd >> 16;
            }
            int cls, tIdx;
            alfClassify(sumV, sumH, sumD0, sumD1, vb ? 96 : 64, g.bdL, cls, tIdx);
            const AlfLumaEntry* e = S.lumaTab + ((size_t)(setIdx * 25 + cls) * 4 + tIdx);
            const uint32_t maxvP = dup16((1 << g.bdL) - 1);
            alfLumaBlockFast(c0, out, dstY.pitch, e, maxvP, vb);
          }
        }
      }
      else
      {
        // no luma ALF in this CTU: copy (128-bit rows)
        for (int i = tid; i < SA_TH * (SA_T / 8); i += SA_THREADS)
        {
          const int r = i >> 3, gc = i & 7;
          const int y = y0 + r, x = x0 + 8 * gc;
          if (y >= py0 && y < py1 && x >= px0 && x < px1)
            *reinterpret_cast<int4*>(dstY.p + (size_t)y * dstY.pitch + x) = *reinterpret_cast<const int4*>(&lumaB[(r + SA_HY) * SA_P + 8 * gc + SA_HX]);
        }
      }
      if (g.ncomp > 1)
      {
        // chroma: one item = 4 horizontally adjacent samples
        const int tcx0 = x0 >> g.sx, tcy0 = y0 >> g.sy, qShift = 4 - g.sx, quads = (tw >> 2) << thLogC;       // tw / 4 = 1 << qShift
#pragma unroll 1
        for (int c = 0; c < 2; c++)
        {
          const PlaneDev dstC = S.buf[dstBuf][1 + c];
          const pel* Bc = c ? B2 : B1;
          const bool fOn = c ? alfCr : alfCb;
          const int idc = c ? ccCr : ccCb;
          ChromaCoef C;
          if (fOn) C = chromaCoef(&alfG->chromaTab[c ? ctl.altCr : ctl.altCb]);
          const int16_t* ccg = alfG->cc[c][idc ? idc - 1 : 0];
          for (int j = tid; j < quads; j += SA_THREADS)
          {
            const int r = j >> qShift, qx = (j & ((1 << qShift) - 1)) * 4;
            const int x = tcx0 + qx, y = tcy0 + r;
            if (x < (px0 >> g.sx) || x >= (px1 >> g.sx) || y < (py0 >> g.sy) || y >= (py1 >> g.sy)) continue;
            const pel* cb = &Bc[(r + SA_HY) * L.pitchC + qx + SA_HX];
            uint2 v = *reinterpret_cast<const uint2*>(cb);
            if (fOn)
            {
              int lim; bool nearVb;
              vbLimit(y & (ctuH - 1), vbC, 2, lim, nearVb);
              v = alfChromaQuad(cb, min(1, lim) * L.pitchC, min(2, lim) * L.pitchC, nearVb, C, maxcP);
            }
            if (idc)
            {
              // CC-ALF row offsets in the luma tile (filterBlkCcAlf :1376-1386)
              const int ly = (r << g.sy) + SA_HY, lpos = (y << g.sy) & ctuMask;
              int l1 = SA_P, l2 = -SA_P, l3 = 2 * SA_P;
              if (lpos == vbL - 2 || lpos == vbL + 1) l3 = SA_P;
              else if (lpos == vbL - 1 || lpos == vbL) l1 = l2 = l3 = 0;
              if (g.sx == 1)
              {
                const uint2 d = ccAlfQuad420(&lumaB[ly * SA_P + (qx << 1) + SA_HX], l1, l2, l3, alfG->ccB[c][idc - 1], maxcP, halfP);
                v.x = addClamp0(v.x, d.x, maxcP);
                v.y = addClamp0(v.y, d.y, maxcP);
              }
              else
              {
                int res[4] = { (int)(v.x & 0xffff), (int)(v.x >> 16), (int)(v.y & 0xffff), (int)(v.y >> 16) };
                int cc[7];
#pragma unroll
                for (int k = 0; k < 7; k++) cc[k] = __ldg(&ccg[k]);
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                  const pel* l = &lumaB[ly * SA_P + qx + q + SA_HX];
                  const int cur = l[0];
                  int s = cc[0] * (l[l2] - cur) + cc[1] * (l[-1] - cur) + cc[2] * (l[1] - cur) + cc[3] * (l[l1 - 1] - cur) + cc[4] * (l[l1] - cur) +
                          cc[5] * (l[l1 + 1] - cur) + cc[6] * (l[l3] - cur);
                  s = (s + 64) >> 7;
                  s = clip3(0, maxc, s + half) - half;
                  res[q] = clip3(0, maxc, res[q] + s);
                }
                v = make_uint2((uint32_t)res[0] | (uint32_t)res[1] << 16, (uint32_t)res[2] | (uint32_t)res[3] << 16);
              }
            }
            *reinterpret_cast<uint2*>(dstC.p + (size_t)y * dstC.pitch + x) = v;
          }
        }
      }
      if (part + 1 < nparts) __syncthreads();                  // the next part re-pads the tile and recomputes the cells
    }
    __syncthreads();                                         // all reads of stage buffers / B / cells are done before they are refilled
    cur = nxt;
  }
}


// ---- the common kernel: one CTU per tile, no signalled virtual boundary ------------------------------------------------------
// (pictures with signalled virtual boundaries or CTU size 32 take k_alf_parts above.)  Per tile:
//   wait      the loads of the tile (issued one tile ahead) have landed: three TMA boxes of samples + two bulk copies of operand
//             tables -- the luma filter set of the tile's CTU (25 classes x 4 transposes, 17.6 KB) and the chroma / CC-ALF
//             tables of its picture.  Round 2 measurement (tools/microbench/mb_alf_luma.cu): gathering the ten 128-bit words of
//             a block's filter entry from global memory -- every thread of a warp another entry -- cost a third of the luma
//             filter's time; from shared memory the same gather is ten LDS.128.
//   pad       only tiles on a picture border or with a clipped CTU side: replicate samples into the outside (CTA barrier)
//   phase 1   every thread: Laplacian cells of its 4x4 block (packed diagonal pairs) and its part of the vertical-pair copy of the
//             luma tile; spare threads: the ring of halo cells and the border of the copy                       (CTA barrier)
//   phase 2   class of the block from the 4x4 cell window, then the chroma 5x5 + CC-ALF quads of the thread, then the 7x7 luma
//             block from the vertical-pair copy                                                                    (CTA barrier)
// The control record of a tile's CTU is all the kernel reads from global memory besides the samples; it travels through
// registers, loaded two tiles ahead (the tile after next needs it when its table copy is issued).  Plane and table addresses
// are computed from the slot number (AlfAddr).  k420 = true compiles the chroma geometry of 4:2:0 in (one quad per thread and
// plane, no loop); the other formats take the run-time instantiation.
__device__ __forceinline__ uint4 alfLoadCtl(const AlfAddr& A, const Geom& g, int firstSlot, const SaWalk& p, int ty0)
{
  const CtuCtlDev* ctl = reinterpret_cast<const CtuCtlDev*>(A.side + (size_t)(firstSlot + p.slot) * A.sideStride + A.offCtl);
  return __ldg(reinterpret_cast<const uint4*>(&ctl[(((p.ty + ty0) * SA_TH) >> g.ctuLog2) * g.wCtus + ((p.tx * SA_T) >> g.ctuLog2)]));
}

// asynchronous loads of one tile into `stage` (one thread): sample boxes + operand tables
__device__ __forceinline__ void alfPrefetch(unsigned char* smraw, const SaLayout& L, int stage, const AlfAddr& A, const CUtensorMap* maps, int slotAbs, const SaWalk& p,
                                            const Geom& g, int ncomp, int sx, int sy, int ty0, uint4 ctlv)
{
  const int x0 = p.tx * SA_T, y0 = (p.ty + ty0) * SA_TH;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smraw + L.offBar) + stage;
  unsigned char* st = smraw + stage * L.stageBytes;
  const uint32_t flags = (ctlv.z >> 8) & 0xff, enY = ctlv.x & 0xff, setIdx = ctlv.y >> 24, grp = (ctlv.z >> 16) & 0xff;
  const bool small = (flags & 1) != 0, tab = small && enY != 0 && !(flags & 2);
  const unsigned char* side = A.side + (size_t)slotAbs * A.sideStride;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic-proxy accesses to the stage buffers are ordered before the async writes
  mbarExpectTx(bar, (uint32_t)(L.lumaBytes + 2 * L.chromaBytes + (tab ? AT_SET_BYTES : 0) + (small ? ALF_SMALL_BYTES : 0)));
  tmaLoad2D(st, maps, x0 - SA_HX, y0 - SA_HY, bar);
  if (ncomp > 1)
  {
    tmaLoad2D(st + L.lumaBytes, maps + 1, (x0 >> sx) - SA_HX, (y0 >> sy) - SA_HY, bar);
    tmaLoad2D(st + L.lumaBytes + L.chromaBytes, maps + 2, (x0 >> sx) - SA_HX, (y0 >> sy) - SA_HY, bar);
  }
  if (tab) bulkLoad(st + L.offSet, side + A.offTab + (size_t)setIdx * AT_SET_BYTES, AT_SET_BYTES, bar);
  if (small) bulkLoad(st + L.offSmall, side + A.offAlf + grp * sizeof(AlfDev), ALF_SMALL_BYTES, bar);
}

// peer band mode: a tile of the band's first (last) tile row reads the four rows above (below) the band, which the neighbour's
// k_dbf_sao stores into this rank's plane -- the walking thread waits for the neighbour's flag before it issues the tile's loads
__device__ __forceinline__ void alfBandWait(const BandDev& band, int ty, int tilesY)
{
  if (!band.myFlags) return;
  if (ty == 0 && band.peerPlanes[0]) while (ldAcquireSys(&band.myFlags[0]) < band.iter) { }
  if (ty == tilesY - 1 && band.peerPlanes[1]) while (ldAcquireSys(&band.myFlags[1]) < band.iter) { }
  asm volatile("fence.proxy.async.global;" ::: "memory");        // the rows were written through the generic proxy, the tile loads read through the async proxy
}

// what every thread needs to know about a tile: written to shared memory by the one thread that walks the tile sequence
struct alignas(16) AlfTileDesc
{
  int x0, y0, slotAbs, valid;
  uint4 ctl;                     // CtuCtlDev of the tile's CTU
  pel* dY;                       // destination buffer of the tile's slot (the 64-bit address arithmetic is done once per tile, not per thread)
  uint32_t border, pad;          // border: the tile touches the picture border or a clipped CTU side (samples outside must be replicated)
};
static_assert(sizeof(AlfTileDesc) == 48, "AlfTileDesc layout (SaLayout reserves 2 x 48 bytes)");

// the tile reads samples outside the picture or across a CTU side the filter must not cross (clip byte of the control record)
__device__ __forceinline__ uint32_t alfTileBorder(int x0, int y0, const Geom& g, uint4 ctlv)
{
  return x0 == 0 || y0 == 0 || x0 + SA_T + 8 > g.w || y0 + SA_TH + 8 > g.h || (ctlv.z & 0xff) != 0;
}

template <bool k420>
__global__ void __launch_bounds__(SA_THREADS, SA_CTAS_PER_SM) k_alf(const AlfAddr A, const CUtensorMap* __restrict__ tmaps, int firstSlot, int numSlots,
                                                                    int srcBuf, int dstBuf, Geom g, int tilesX, int tilesY, int ty0, SaStep step, const BandDev band)
{
  extern __shared__ __align__(128) unsigned char smraw[];
  const int sx = k420 ? 1 : g.sx, sy = k420 ? 1 : g.sy, ncomp = k420 ? 3 : g.ncomp;
  const SaLayout L = saLayout(sx, sy, ncomp, true);
  const int tid = threadIdx.x;
  uint2 (*cell)[SA_CELLP] = reinterpret_cast<uint2 (*)[SA_CELLP]>(smraw + L.offCell);
  uint32_t* const V = reinterpret_cast<uint32_t*>(smraw + L.offV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + L.offBar);
  AlfTileDesc* const desc = reinterpret_cast<AlfTileDesc*>(smraw + L.offDesc);
  const int vbL = g.ctu - 4, ctuMask = g.ctu - 1;
  const int bi = tid >> 4, bj = tid & 15;
  const int ctuH = g.ctu >> sy, vbC = ctuH - 2;
  const uint32_t maxcP = dup16((1 << g.bdC) - 1), halfP = dup16((1 << g.bdC) >> 1), maxvP = dup16((1 << g.bdL) - 1);
  const int wC = g.w >> sx, hC = g.h >> sy;
  const int pitchY = A.pitchY, pitchCh = A.pitchC;
  // thread constants: block origin in the row-major tile and in the vertical-pair copy
  const int hOff = (4 * bi + SA_HY) * SA_P + 4 * bj + SA_HX;
  uint32_t* const vBlk = V + (4 * bi + 4) * AV_COLS + 4 * bj;
  // 4:2:0: the thread's chroma quad.  The rows of a warp are visited in the order 0, 2, 1, 3 so that the two rows of a half warp
  // lie 16 banks apart in shared memory (chroma pitch 24 words).
  const int rC = ((tid >> 3) & ~3) | (((tid >> 3) & 1) << 1) | ((tid >> 4) & 1), qC = (tid & 7) * 4;
  const int cOff = (rC + SA_HY) * L.pitchC + qC + SA_HX, lOff = (2 * rC + SA_HY) * SA_P + 2 * qC + SA_HX;
  // phase-1 side jobs: border of the vertical-pair copy (threads 0..217), ring of halo cells (the last 132 threads)
  int cpH = -1, cpV = 0;
  if (tid < 90 + 128)
  {
    int rr, lx;
    if (tid < 90) { const int q = tid / 18; rr = q < 3 ? 1 + q : 65 + q; lx = (tid - q * 18) * 4; }    // row pairs 1..3, 68, 69 x 18 column quads
    else          { const int q = tid - 90; rr = 4 + (q >> 1); lx = (q & 1) ? 68 : 0; }                // columns 0..3 and 68..71 of row pairs 4..67
    cpH = rr * SA_P + lx + SA_HX - 4; cpV = rr * AV_COLS + lx;
  }
  constexpr int kRing = 2 * SA_CELLS + 2 * (SA_CELLR - 2);
  int ringI = -1, ringJ = 0;
  if (tid >= SA_THREADS - kRing)
  {
    const int q = tid - (SA_THREADS - kRing);
    if (q < 2 * SA_CELLS) { ringI = q < SA_CELLS ? 0 : SA_CELLR - 1; ringJ = q < SA_CELLS ? q : q - SA_CELLS; }
    else { const int k = q - 2 * SA_CELLS; ringI = 1 + (k >> 1); ringJ = (k & 1) ? SA_CELLS - 1 : 0; }
  }

  // ---- the tile walk lives in thread 0: it loads the control records two tiles ahead, issues the loads of the next tile and
  // leaves a descriptor per tile in shared memory (round 2: the walk replicated in every thread was 4 instructions per pixel)
  SaWalk nxt;
  uint4 ctlB = make_uint4(0, 0, 0, 0);
  {
    SaWalk cur;
    const int tilesPerPic = tilesX * tilesY, t = blockIdx.x;
    cur.slot = t / tilesPerPic;
    const int rem = t - cur.slot * tilesPerPic;
    cur.ty = rem / tilesX;
    cur.tx = rem - cur.ty * tilesX;
    if (cur.slot >= numSlots) return;
    nxt = cur;
    if (tid == 0)
    {
      mbarInit(&bars[0], 1);
      mbarInit(&bars[1], 1);
      mbarInit(&bars[2], SA_THREADS / 32);                     // "cells ready": one arrival per warp and tile
      mbarInit(&bars[3], SA_THREADS / 32);                     // "tile done"
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      saAdvance(nxt, step, tilesX, tilesY);
      const uint4 ctlA = alfLoadCtl(A, g, firstSlot, cur, ty0);
      if (nxt.slot < numSlots) ctlB = alfLoadCtl(A, g, firstSlot, nxt, ty0);
      desc[0].x0 = cur.tx * SA_T; desc[0].y0 = (cur.ty + ty0) * SA_TH; desc[0].slotAbs = firstSlot + cur.slot; desc[0].valid = 1; desc[0].ctl = ctlA;
      desc[0].dY = A.planes + (size_t)(firstSlot + cur.slot) * A.slotStride + (size_t)dstBuf * A.bufStride;
      desc[0].border = alfTileBorder(desc[0].x0, desc[0].y0, g, ctlA);
      alfBandWait(band, cur.ty, tilesY);
      alfPrefetch(smraw, L, 0, A, tmaps + ((size_t)(firstSlot + cur.slot) * 3 + srcBuf) * 3, firstSlot + cur.slot, cur, g, ncomp, sx, sy, ty0, ctlA);
    }
  }
  __syncthreads();
  for (uint32_t it = 0;; it++)
  {
    const int stage = it & 1;
    const int4 dsc = *reinterpret_cast<const int4*>(&desc[stage]);
    if (!dsc.w) break;
    const uint4 ctl0 = desc[stage].ctl;
    // Split barriers (mbarriers 2 and 3, one arrival per warp): a warp waits for the others only where it needs their results.
    //   gate   before a thread's first store of phase 1: the previous tile is done everywhere (its readers of the copy, the cells and
    //          the other stage are finished); the walking thread then describes the next tile and issues its loads
    //   cells  after the chroma quads, before the classification: the cells and the copy of this tile are complete
    // The loads and the arithmetic of phase 1 and the chroma quads run while slower warps are still in the previous phase.
    auto gate = [&]()
    {
      if (it > 0) mbarWait(&bars[3], (it - 1) & 1);
      if (tid == 0)
      {
        const bool more = nxt.slot < numSlots;
        AlfTileDesc& d = desc[stage ^ 1];
        d.x0 = nxt.tx * SA_T; d.y0 = (nxt.ty + ty0) * SA_TH; d.slotAbs = firstSlot + nxt.slot; d.valid = more; d.ctl = ctlB;
        d.dY = A.planes + (size_t)(firstSlot + nxt.slot) * A.slotStride + (size_t)dstBuf * A.bufStride;
        d.border = alfTileBorder(d.x0, d.y0, g, ctlB);
        if (more)
        {
          alfBandWait(band, nxt.ty, tilesY);
          alfPrefetch(smraw, L, stage ^ 1, A, tmaps + ((size_t)(firstSlot + nxt.slot) * 3 + srcBuf) * 3, firstSlot + nxt.slot, nxt, g, ncomp, sx, sy, ty0, ctlB);
          saAdvance(nxt, step, tilesX, tilesY);
          ctlB = nxt.slot < numSlots ? alfLoadCtl(A, g, firstSlot, nxt, ty0) : make_uint4(0, 0, 0, 0);
        }
      }
    };
    const int x0 = dsc.x, y0 = dsc.y;
    const uint4 dsc2 = *reinterpret_cast<const uint4*>(&desc[stage].dY);
    pel* const dY = reinterpret_cast<pel*>((uint64_t)dsc2.x | (uint64_t)dsc2.y << 32);
    CtuCtlDev ctl;
    *reinterpret_cast<uint4*>(&ctl) = ctl0;
    // the picture's AlfDev records in global memory are only read on the rare paths (wide coefficients, CC-ALF filters whose
    // coefficient sum does not fit a byte, non-4:2:0 CC-ALF): [0] holds the luma sets, [ctl.grp] the chroma / CC-ALF data of the CTU's slice
    auto alfDevAt = [&](int grp) { return reinterpret_cast<const AlfDev*>(A.side + (size_t)dsc.z * A.sideStride + A.offAlf) + grp; };
    const bool alfOn = (ctl.flags & 1) != 0, wide = (ctl.flags & 2) != 0;
    const bool alfY = alfOn && ctl.enY != 0, alfCb = alfOn && ctl.enCb != 0, alfCr = alfOn && ctl.enCr != 0;
    const int ccCb = alfOn ? ctl.ccCb : 0, ccCr = alfOn ? ctl.ccCr : 0;
    const int clip = ctl.clip;
    unsigned char* const st = smraw + stage * L.stageBytes;
    pel* const A0 = reinterpret_cast<pel*>(st);
    pel* const A1 = reinterpret_cast<pel*>(st + L.lumaBytes);
    pel* const A2 = reinterpret_cast<pel*>(st + L.lumaBytes + L.chromaBytes);
    const AlfDev* const small = reinterpret_cast<const AlfDev*>(st + L.offSmall);      // only chromaTab and ccK are there

    mbarWait(&bars[stage], (it >> 1) & 1);                   // the async writes of this tile are visible to this thread

    // tiles on the picture border: replicate the border samples into the zero-filled outside
    // (= UnitBuf::extendBorderPel of the ALF input, AdaptiveLoopFilter.cpp:411); CTU sides at a slice / tile boundary the
    // filter must not read across: the same replication at the clipped sides (:452-490)
    if (dsc2.z)
    {
      const int cx0 = x0 & ~ctuMask, cy0 = y0 & ~ctuMask, cx1 = min(cx0 + g.ctu, g.w), cy1 = min(cy0 + g.ctu, g.h);
      const int xlo = (clip & VTMGPU_ALF_CLIP_LEFT) ? cx0 : 0, xhi = (clip & VTMGPU_ALF_CLIP_RIGHT) ? cx1 : g.w;
      const int ylo = (clip & VTMGPU_ALF_CLIP_TOP) ? cy0 : 0, yhi = (clip & VTMGPU_ALF_CLIP_BOTTOM) ? cy1 : g.h;
      int pad = clip & (VTMGPU_ALF_PAD_TL | VTMGPU_ALF_PAD_BR);      // only for the tile that holds that corner of the CTU
      if (x0 != cx0 || y0 != cy0) pad &= ~VTMGPU_ALF_PAD_TL;
      if (min(x0 + SA_T, g.w) != cx1 || min(y0 + SA_TH, g.h) != cy1) pad &= ~VTMGPU_ALF_PAD_BR;
      const int bxc = (x0 >> sx) - SA_HX, byc = (y0 >> sy) - SA_HY, twc = SA_T >> sx, thc = SA_TH >> sy;
      if (alfY || ccCb || ccCr) saReplicateBorder(A0, x0 - SA_HX, y0 - SA_HY, xlo, xhi, ylo, yhi, SA_P, SA_T, SA_TH, 4);
      if (alfCb) saReplicateBorder(A1, bxc, byc, xlo >> sx, xhi >> sx, ylo >> sy, yhi >> sy, L.pitchC, twc, thc, 3);
      if (alfCr) saReplicateBorder(A2, bxc, byc, xlo >> sx, xhi >> sx, ylo >> sy, yhi >> sy, L.pitchC, twc, thc, 3);
      if (pad)
      {
        // no barrier needed in between: the corners lie outside the clamp window's replicated ranges (their sides are not clipped)
        if (alfY || ccCb || ccCr) saPadCorners(A0, x0 - SA_HX, y0 - SA_HY, cx0, cy0, cx1, cy1, SA_P, 4, pad);
        if (alfCb) saPadCorners(A1, bxc, byc, cx0 >> sx, cy0 >> sy, cx1 >> sx, cy1 >> sy, L.pitchC, 3, pad);
        if (alfCr) saPadCorners(A2, bxc, byc, cx0 >> sx, cy0 >> sy, cx1 >> sx, cy1 >> sy, L.pitchC, 3, pad);
      }
      __syncthreads();
    }

    const pel* const c0 = A0 + hOff;
    const int yb = (y0 + 4 * bi) & ctuMask;
    const int vb = yb == vbL - 4 ? 1 : (yb == vbL ? 2 : 0);        // non-zero for a whole warp (two block rows)
    const bool vbTile = ((y0 + 56) & ctuMask) == vbL - 4;          // the tile's block rows 14 and 15 lie at the ALF virtual boundary

    // ---- phase 1: Laplacian cells + vertical-pair copy (luma ALF only) -------------------------------------------------
    if (alfY && !(ALF_WHATIF & 4))
    {
      alfBlockCellsAndCopy<SA_P, SA_CELLP>(c0, cell, vBlk + 4, bi, bj, vb == 0, gate);
      if (vbTile && tid < 128)
      {
        // the 128 cells of the two block rows at the virtual boundary (rows replaced, alfCellAny), one per thread of the first four
        // warps: left to the warp that owns those rows (four cells per thread) they made it the straggler of every such tile
        const int li = 29 + 2 * (tid >> 6) + ((tid >> 1) & 1), lj = 2 * ((tid & 63) >> 2) + 1 + (tid & 1);
        cell[li][lj] = alfCellAny<SA_P>(&A0[(2 * li + SA_HY - 2) * SA_P + 2 * lj + SA_HX - 2], y0 - 2 + 2 * li, ctuMask, vbL);
      }
      if (cpH >= 0) alfCopyQuad<SA_P>(A0 + cpH, V + cpV);
      if (ringI >= 0) cell[ringI][ringJ] = alfCellAny<SA_P>(&A0[(2 * ringI + SA_HY - 2) * SA_P + 2 * ringJ + SA_HX - 2], y0 - 2 + 2 * ringI, ctuMask, vbL);
    }
    else gate();
    __syncwarp();
    if ((tid & 31) == 0) mbarArrive(&bars[2]);

    // ---- phase 2 ----------------------------------------------------------------------------------------------------
    const bool lumaBlk = alfY && x0 + 4 * bj < g.w && y0 + 4 * bi < g.h;         // the last tile of a row / column may be partial

    if (ncomp > 1 && !(ALF_WHATIF & 1))
    {
      const int qShift = 4 - sx, quads = k420 ? SA_THREADS : ((SA_T >> sx) >> 2) << (SA_THLOG - sy);
#pragma unroll 1
      for (int j = tid; j < quads; j += SA_THREADS)
      {
        // one item = 4 horizontally adjacent chroma samples, both planes (what depends only on the position is computed once)
        int r, qx, co, lo;
        if (k420) { r = rC; qx = qC; co = cOff; lo = lOff; }
        else
        {
          r = j >> qShift;
          if (qShift == 3) r = (r & ~3) | ((r & 1) << 1) | ((r >> 1) & 1);
          qx = (j & ((1 << qShift) - 1)) * 4;
          co = (r + SA_HY) * L.pitchC + qx + SA_HX; lo = ((r << sy) + SA_HY) * SA_P + (qx << sx) + SA_HX;
        }
        const int x = (x0 >> sx) + qx, y = (y0 >> sy) + r;
        if (x >= wC || y >= hC) continue;
        int lim; bool nearVb;
        vbLimit(y & (ctuH - 1), vbC, 2, lim, nearVb);
        const int o1 = min(1, lim) * L.pitchC, o2 = min(2, lim) * L.pitchC;
        // CC-ALF row offsets in the luma tile (filterBlkCcAlf :1376-1386)
        const int lpos = (y << sy) & ctuMask;
        int l1 = SA_P, l2 = -SA_P, l3 = 2 * SA_P;
        if (lpos == vbL - 2 || lpos == vbL + 1) l3 = SA_P;
        else if (lpos == vbL - 1 || lpos == vbL) l1 = l2 = l3 = 0;
        const pel* l = A0 + lo;
        const int dOff = y * pitchCh + x;
#pragma unroll 1
        for (int c = 0; c < 2; c++)
        {
          const bool fOn = c ? alfCr : alfCb;
          const int idc = c ? ccCr : ccCb;
          const pel* cb = (c ? A2 : A1) + co;
          uint2 v = *reinterpret_cast<const uint2*>(cb);
          if (fOn) v = alfChromaQuad(cb, o1, o2, nearVb, chromaCoef(&small->chromaTab[c ? ctl.altCr : ctl.altCb]), maxcP);
          if (idc)
          {
            if (sx == 1)
            {
              const uint4 ck = *reinterpret_cast<const uint4*>(small->ccK[c][idc - 1]);
              const uint2 d = ck.w ? ccAlfQuadDual(l, l1, l2, l3, ck.x, ck.y, ck.z, maxcP, halfP)
                                   : ccAlfQuad420(l, l1, l2, l3, alfDevAt(ctl.grp)->ccB[c][idc - 1], maxcP, halfP);
              v.x = addClamp0(v.x, d.x, maxcP);
              v.y = addClamp0(v.y, d.y, maxcP);
            }
            else
            {
              const int16_t* ccg = alfDevAt(ctl.grp)->cc[c][idc - 1];
              const int maxc = (1 << g.bdC) - 1, half = (1 << g.bdC) >> 1;
              int res[4] = { (int)(v.x & 0xffff), (int)(v.x >> 16), (int)(v.y & 0xffff), (int)(v.y >> 16) };
              int cc[7];
#pragma unroll
              for (int k = 0; k < 7; k++) cc[k] = __ldg(&ccg[k]);
#pragma unroll
              for (int q = 0; q < 4; q++)
              {
                const pel* lq = l + q;
                const int cu = lq[0];
                int s = cc[0] * (lq[l2] - cu) + cc[1] * (lq[-1] - cu) + cc[2] * (lq[1] - cu) + cc[3] * (lq[l1 - 1] - cu) + cc[4] * (lq[l1] - cu) +
                        cc[5] * (lq[l1 + 1] - cu) + cc[6] * (lq[l3] - cu);
                s = (s + 64) >> 7;
                s = clip3(0, maxc, s + half) - half;
                res[q] = clip3(0, maxc, res[q] + s);
              }
              v = make_uint2((uint32_t)res[0] | (uint32_t)res[1] << 16, (uint32_t)res[2] | (uint32_t)res[3] << 16);
            }
          }
          *reinterpret_cast<uint2*>(dY + (c ? A.compOff[2] : A.compOff[1]) + dOff) = v;
        }
      }
    }

    mbarWait(&bars[2], it & 1);                              // the cells and the copy of this tile are complete
    const AlfLumaEntry* e = nullptr;
    if ((ALF_WHATIF & 2) && lumaBlk && !wide) e = reinterpret_cast<const AlfLumaEntry*>(st + L.offSet) + (tid & 63);
    if (!(ALF_WHATIF & 2) && lumaBlk && !wide)
    {
      // window = cells (2bi .. 2bi+3) x (2bj .. 2bj+3).  Blocks at the virtual boundary use 3 of the 4 cell rows and the
      // scale 96 (deriveClassificationBlk :977-1010).  Packed 16-bit sums: a cell holds at most 4 * (2^bd - 1) per direction, so
      // up to 10 bits the whole window (16 cells) stays below 2^16 per lane; above, every cell row is unpacked on its own.
      int sumV, sumH, sumD0, sumD1;
      const uint4* rp = reinterpret_cast<const uint4*>(&cell[2 * bi][2 * bj]);
      if (g.bdL <= 10)
      {
        uint32_t vh = 0, dd = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
          if ((vb == 1 && i == 3) || (vb == 2 && i == 0)) continue;
          const uint4 q0 = rp[i * (SA_CELLP / 2)], q1 = rp[i * (SA_CELLP / 2) + 1];
          vh += q0.x + q0.z + q1.x + q1.z; dd += q0.y + q0.w + q1.y + q1.w;
        }
        sumV = vh & 0xffff; sumH = vh >> 16; sumD0 = dd & 0xffff; sumD1 = dd >> 16;
      }
      else
      {
        sumV = sumH = sumD0 = sumD1 = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
          if ((vb == 1 && i == 3) || (vb == 2 && i == 0)) continue;
          const uint4 q0 = rp[i * (SA_CELLP / 2)], q1 = rp[i * (SA_CELLP / 2) + 1];
          const uint32_t vh = q0.x + q0.z + q1.x + q1.z, dd = q0.y + q0.w + q1.y + q1.w;
          sumV += vh & 0xffff; sumH += vh >> 16; sumD0 += dd & 0xffff; sumD1 += dd >> 16;
        }
      }
      int cls, tIdx;
      alfClassify(sumV, sumH, sumD0, sumD1, vb ? 96 : 64, g.bdL, cls, tIdx);
      e = reinterpret_cast<const AlfLumaEntry*>(st + L.offSet) + (cls * 4 + tIdx);
    }


    if (alfY && !(ALF_WHATIF & 8))
    {
      if (lumaBlk)
      {
        pel* out = dY + (y0 + 4 * bi) * pitchY + x0 + 4 * bj;
        if (wide)    alfLumaBlockGeneric(cell, c0, out, pitchY, bi, bj, y0 + 4 * bi, &alfDevAt(0)->luma[ctl.setIdx][0][0], ctuMask, vbL, g.bdL);
        else if (vb) alfLumaBlockFast(c0, out, pitchY, e, maxvP, vb);
        else         alfLumaBlockV<ALF_BAL>(vBlk, out, pitchY, loadLumaCoef(e), maxvP);
      }
    }
    else
    {
      // no luma ALF in this CTU: copy (128-bit rows)
      for (int i = tid; i < SA_TH * (SA_T / 8); i += SA_THREADS)
      {
        const int r = i >> 3, gc = i & 7;
        const int y = y0 + r, x = x0 + 8 * gc;
        if (y < g.h && x < g.w)
          *reinterpret_cast<int4*>(dY + y * pitchY + x) = *reinterpret_cast<const int4*>(&A0[(r + SA_HY) * SA_P + 8 * gc + SA_HX]);
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbarArrive(&bars[3]);               // this warp's reads of the stage buffers, the copy and the cells are done
  }
  if (band.myFlags && tid == 0)
  {
    // peer band mode: the last CTA to finish tells both neighbours that this rank has read its halo rows of this iteration
    __threadfence_system();
    if (atomicAdd(&band.myFlags[5], 1u) == gridDim.x - 1)
    {
      band.myFlags[5] = 0;
      __threadfence_system();
      if (band.peerFlags[0]) stReleaseSys(&band.peerFlags[0][3], band.iter);        // above: "the neighbour below has finished its ALF"
      if (band.peerFlags[1]) stReleaseSys(&band.peerFlags[1][2], band.iter);        // below: "the neighbour above has finished its ALF"
    }
  }
}

}   // namespace vtmgpu
