// dbf_kernel.cuh -- deblocking filter + SAO in one pass over the picture (sm_100a): kernel k_dbf_sao, and k_dbf_queues.
//
//   LoopFilter::loopFilterPic (LoopFilter.cpp:145; xEdgeFilterLuma :971-1080, xEdgeFilterChroma :1246-1279, filters :1302-1555)
//   SampleAdaptiveOffset::SAOProcess (SampleAdaptiveOffset.cpp:618; offsetBlock :293-547) -- sao_device.cuh
//
// Persistent CTAs (3 per SM, 256 threads) walk the 128 x 64 tiles of the planes of a batch of picture slots round robin.  Per tile:
//   load     the tile + an 8-sample halo arrives by ONE TMA box (zero filled outside the picture) and the tile's two queues of ACTIVE
//            segments by bulk copies, all on one mbarrier, issued one tile ahead (two stage buffers)
//   pass 1   all vertical edges of tile + halo, in place in shared memory: the queue entries are handed out to the warps, a lane PAIR
//            decides a luma segment (lines 0 and 3), a QUAD filters it (one line per lane); chroma: a thread per segment
//   barrier
//   pass 2   all horizontal edges on the V-filtered samples, same scheme
//   barrier
//   epilogue every thread: SAO of its 8 x 4 strip of the tile's OWN samples from the deblocked tile (a 128-bit copy where SAO is off),
//            written straight to the OUTPUT plane -- each plane is read once and written once
// The edges on a tile border are evaluated by both neighbouring tiles (each keeps only its own side), so no inter-CTA ordering is
// needed and input / output planes are distinct.  Why the halo of 8 suffices: tile origins are multiples of 64, every block side
// >= 32 samples starts on a multiple of 16, hence an edge 4 samples outside the tile can modify at most 3 samples on the tile's side.
// The queues are built once per picture by k_dbf_queues (below) from the packed segment records of include/vtmgpu.h, whoever
// produced them (uploaded arrays, scattered lists, k_dbf_derive).  HBM traffic per plane: read (1 + halo) + write 1 samples,
// + the queue entries of the active segments.
#pragma once

#include "async_copy.cuh"
#include "sao_device.cuh"
#include "vtmgpu_dev.cuh"
#include "vtmgpu.h"

namespace vtmgpu
{

#ifndef DBF_TILE_W
#define DBF_TILE_W 128          // plane tile width: 128 (256 threads, 3 CTAs per SM) or 64 (128 threads, 5 CTAs per SM)
#endif
constexpr int DBF_TW = DBF_TILE_W, DBF_TH = 64, DBF_HALO = 8;
constexpr int DBF_SW = DBF_TW + 2 * DBF_HALO;           // 144
constexpr int DBF_SH = DBF_TH + 2 * DBF_HALO;           // 80
constexpr int DBF_PITCH = DBF_SW + 8;                   // 152 samples = 304 B: rows shift by 12 banks
constexpr int DBF_THREADS = DBF_TW * 2;                 // one thread per 8 x 4 strip of the tile
constexpr int DBF_CTAS_PER_SM = DBF_TW == 128 ? 3 : 5;
static_assert(DBF_TW == 128 || DBF_TW == 64, "tile width");

// x points at q0 of one line in shared memory; o = step across the edge; P(k) = x[-(k+1)*o], Q(k) = x[k*o]
// tc / beta tables of the standard (LoopFilter.cpp:66-74), only needed on the device with LADF
__constant__ uint16_t kDbfTcTable[66] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,3,4,4,4,4,5,5,5,5,7,7,8,9,10,10,11,13,14,15,17,19,21,24,25,29,33,36,
                                          41,45,51,57,64,71,80,89,100,112,125,141,157,177,198,222,250,280,314,352,395 };
__constant__ uint8_t kDbfBetaTable[64] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,
                                           42,44,46,48,50,52,54,56,58,60,62,64,66,68,70,72,74,76,78,80,82,84,86,88 };

#define PK(k) ((int)x[-((k) + 1) * o])
#define QK(k) ((int)x[(k) * o])

__device__ __forceinline__ bool dbfStrongShort(const pel* x, int o, int d, int beta, int tc, bool pOne)
{
  const int sp3 = pOne ? iabs(PK(1) - PK(0)) : iabs(PK(3) - PK(0));
  const int sq3 = iabs(QK(3) - QK(0));
  return (sp3 + sq3 < (beta >> 3)) && (d < (beta >> 2)) && (iabs(PK(0) - QK(0)) < ((tc * 5 + 1) >> 1));
}

__device__ __forceinline__ bool dbfStrongLong(const pel* x, int o, int d, int beta, int tc, bool largeP, bool largeQ, int lenP, int lenQ)
{
  int sp3 = iabs(PK(3) - PK(0));
  int sq3 = iabs(QK(3) - QK(0));
  if (largeP)
  {
    int far;
    if (lenP == 7) { far = PK(7); sp3 += iabs(PK(4) - PK(5) - PK(6) + far); }
    else           { far = PK(5); }
    sp3 = (sp3 + iabs(PK(3) - far) + 1) >> 1;
  }
  if (largeQ)
  {
    int far;
    if (lenQ == 7) { far = QK(7); sq3 += iabs(QK(4) - QK(5) - QK(6) + far); }
    else           { far = QK(5); }
    sq3 = (sq3 + iabs(far - QK(3)) + 1) >> 1;
  }
  return (sp3 + sq3 < ((beta * 3) >> 5)) && (d < (beta >> 4)) && (iabs(PK(0) - QK(0)) < ((tc * 5 + 1) >> 1));
}

// bilinear long filter of one line (xFilteringPandQ / xBilinearFilter); nP,nQ in {3,5,7}, not both 3
__device__ void dbfLongLine(pel* x, int o, int nP, int nQ, int tc, bool wP, bool wQ)
{
  int p[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { p[k] = PK(k); q[k] = QK(k); }
  const int refP = nP == 7 ? (p[6] + p[7] + 1) >> 1 : (nP == 5 ? (p[4] + p[5] + 1) >> 1 : (p[2] + p[3] + 1) >> 1);
  const int refQ = nQ == 7 ? (q[6] + q[7] + 1) >> 1 : (nQ == 5 ? (q[4] + q[5] + 1) >> 1 : (q[2] + q[3] + 1) >> 1);
  int mid;
  if (nP == nQ)
  {
    if (nP == 5) mid = (2 * (p[0] + q[0] + p[1] + q[1] + p[2] + q[2]) + p[3] + q[3] + p[4] + q[4] + 8) >> 4;
    else         mid = (2 * (p[0] + q[0]) + p[1] + q[1] + p[2] + q[2] + p[3] + q[3] + p[4] + q[4] + p[5] + q[5] + p[6] + q[6] + 8) >> 4;
  }
  else
  {
    const int nL = max(nP, nQ), nS = min(nP, nQ);
    if (nL == 7 && nS == 5)
      mid = (2 * (p[0] + q[0] + p[1] + q[1]) + p[2] + q[2] + p[3] + q[3] + p[4] + q[4] + p[5] + q[5] + 8) >> 4;
    else if (nL == 7 && nS == 3)
    {
      const bool pl = nP > nQ;
      const int L0 = pl ? p[0] : q[0], S0 = pl ? q[0] : p[0], S1 = pl ? q[1] : p[1], S2 = pl ? q[2] : p[2];
      int sumL = 0;
#pragma unroll
      for (int k = 1; k < 7; k++) sumL += pl ? p[k] : q[k];
      mid = (2 * (L0 + S0) + S0 + 2 * (S1 + S2) + S1 + sumL + 8) >> 4;
    }
    else
      mid = (p[0] + q[0] + p[1] + q[1] + p[2] + q[2] + p[3] + q[3] + 4) >> 3;
  }
  // dbCoeffs{7,5,3} are arithmetic progressions: 59-9k, 58-13k, 53-21k ; tc scale {6,5,4,3,2,1,1} or {6,4,2}
  if (wP)
  {
    const int c0 = nP == 7 ? 59 : (nP == 5 ? 58 : 53), cd = nP == 7 ? 9 : (nP == 5 ? 13 : 21);
#pragma unroll
    for (int k = 0; k < 7; k++)
      if (k < nP)
      {
        const int c = c0 - cd * k, t = nP == 3 ? 6 - 2 * k : max(6 - k, 1);
        const int cv = (tc * t) >> 1;
        x[-(k + 1) * o] = (pel)clip3(p[k] - cv, p[k] + cv, (mid * c + refP * (64 - c) + 32) >> 6);
      }
  }
  if (wQ)
  {
    const int c0 = nQ == 7 ? 59 : (nQ == 5 ? 58 : 53), cd = nQ == 7 ? 9 : (nQ == 5 ? 13 : 21);
#pragma unroll
    for (int k = 0; k < 7; k++)
      if (k < nQ)
      {
        const int c = c0 - cd * k, t = nQ == 3 ? 6 - 2 * k : max(6 - k, 1);
        const int cv = (tc * t) >> 1;
        x[k * o] = (pel)clip3(q[k] - cv, q[k] + cv, (mid * c + refQ * (64 - c) + 32) >> 6);
      }
  }
}

__device__ __forceinline__ void dbfLumaLine(pel* x, int o, int tc, bool strong, bool wP, bool wQ, bool secondP, bool secondQ, int maxv)
{
  const int p0 = PK(0), p1 = PK(1), p2 = PK(2), q0 = QK(0), q1 = QK(1), q2 = QK(2);
  if (strong)
  {
    const int p3 = PK(3), q3 = QK(3);
    if (wP)
    {
      x[-1 * o] = (pel)clip3(p0 - 3 * tc, p0 + 3 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      x[-2 * o] = (pel)clip3(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
      x[-3 * o] = (pel)clip3(p2 - tc, p2 + tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
    }
    if (wQ)
    {
      x[0]     = (pel)clip3(q0 - 3 * tc, q0 + 3 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      x[1 * o] = (pel)clip3(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
      x[2 * o] = (pel)clip3(q2 - tc, q2 + tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
    }
    return;
  }
  int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
  if (iabs(delta) >= tc * 10) return;
  delta = clip3(-tc, tc, delta);
  const int tc2 = tc >> 1;
  if (wP)
  {
    x[-1 * o] = (pel)clip3(0, maxv, p0 + delta);
    if (secondP) x[-2 * o] = (pel)clip3(0, maxv, p1 + clip3(-tc2, tc2, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1));
  }
  if (wQ)
  {
    x[0] = (pel)clip3(0, maxv, q0 - delta);
    if (secondQ) x[1 * o] = (pel)clip3(0, maxv, q1 + clip3(-tc2, tc2, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1));
  }
}

// LADF: a luma record that carries QPs (include/vtmgpu.h, vtmgpu_ladf) is turned into the usual {tc, beta} record.  The QP
// offset is selected by the mean of p0 / q0 of lines 0 and 3 of the segment in their current state (deriveLADFShift,
// LoopFilter.cpp:815-841), tc / beta come from the tables (:971-975).  Called by all 32 lanes, lane ln of a quad owns line ln;
// tc = 0 in the result means "nothing to filter" (the host drops such records when it derives tc itself).
__device__ __noinline__ uint32_t dbfLadfRecord(uint32_t rec, const pel* x0, int o, int s, int ln, int bd, const LadfDev* ladf)
{
  const pel* x = x0 + ln * s;
  const unsigned qb = threadIdx.x & 28u;
  const int own = PK(0) + QK(0);
  const int level = (__shfl_sync(0xffffffffu, own, qb) + __shfl_sync(0xffffffffu, own, qb + 3)) >> 2;
  int shift = ladf->off[0];
  for (int k = 1; k < ladf->n; k++)
  {
    if (level > ladf->lb[k]) shift = ladf->off[k];
    else break;
  }
  const int t = kDbfTcTable[clip3(0, 65, (int)(rec & 0x7ff) - VTMGPU_DBF_LADF_BIAS + shift)];
  const int tc = bd < 10 ? (t + 2) >> (10 - bd) : t << (bd - 10);
  const int beta = (int)kDbfBetaTable[clip3(0, 63, (int)((rec >> VTMGPU_DBF_L_BETA_SHIFT) & 0x7ff) - VTMGPU_DBF_LADF_BIAS + shift)] << (bd - 8);
  return (rec & ~0x3fffffu) | (uint32_t)tc | (uint32_t)beta << VTMGPU_DBF_L_BETA_SHIFT;
}

// One 4-line luma segment (x0 = q0 of line 0, o = step across the edge, s = step along it), spread over the 4 lanes of a quad:
// lane ln owns line ln of the segment, computes the
// gradients / strong-filter tests of its own line, lines 0 and 3 are broadcast with shuffles (xEdgeFilterLuma evaluates the
// first and the last line of a segment, LoopFilter.cpp:977-1043), then every lane filters its own line.  All 32 lanes of
// the warp must call this together; `valid` = false marks a quad without work (reads stay legal, nothing is written).
__device__ __forceinline__ void dbfLumaSegmentQuad(pel* x0, int o, int s, uint32_t rec, int maxv, int ln, bool valid)
{
  const int tc = rec & 0x7ff;
  const int beta = (rec >> VTMGPU_DBF_L_BETA_SHIFT) & 0x7ff;
  pel* xl = x0 + ln * s;
  const pel* x = xl;
  const int lenP = (rec >> VTMGPU_DBF_L_LENP_SHIFT) & 7, lenQ = (rec >> VTMGPU_DBF_L_LENQ_SHIFT) & 7;
  const bool wP = valid && !(rec & VTMGPU_DBF_L_PNOFILT), wQ = valid && !(rec & VTMGPU_DBF_L_QNOFILT);
  const bool largeP = lenP > 3 && !(rec & VTMGPU_DBF_L_CTUROW), largeQ = lenQ > 3;
  const int sideThr = (beta + (beta >> 1)) >> 3;
  const int dp = iabs(PK(2) - 2 * PK(1) + PK(0)), dq = iabs(QK(0) - 2 * QK(1) + QK(2));
  int dpL = dp, dqL = dq;
  if (largeP) dpL = (dp + iabs(PK(5) - 2 * PK(4) + PK(3)) + 1) >> 1;
  if (largeQ) dqL = (dq + iabs(QK(3) - 2 * QK(4) + QK(5)) + 1) >> 1;
  const bool anyLarge = largeP || largeQ;
  const bool sL = anyLarge && dbfStrongLong(x, o, 2 * (dpL + dqL), beta, tc, largeP, largeQ, lenP, lenQ);
  const bool sS = lenP > 2 && lenQ > 2 && dbfStrongShort(x, o, 2 * (dp + dq), beta, tc, false);
  // broadcast lines 0 and 3: gradients packed dp | dq << 16 (each < 2^13), flags in bits 0,1
  const unsigned lane = threadIdx.x & 31, qb = lane & ~3u;
  const uint32_t gS = (uint32_t)dp | (uint32_t)dq << 16, gL = (uint32_t)dpL | (uint32_t)dqL << 16, fl = (sL ? 1u : 0u) | (sS ? 2u : 0u);
  const uint32_t gS0 = __shfl_sync(0xffffffffu, gS, qb), gS3 = __shfl_sync(0xffffffffu, gS, qb + 3);
  const uint32_t gL0 = __shfl_sync(0xffffffffu, gL, qb), gL3 = __shfl_sync(0xffffffffu, gL, qb + 3);
  const uint32_t fl0 = __shfl_sync(0xffffffffu, fl, qb), fl3 = __shfl_sync(0xffffffffu, fl, qb + 3);
  if (anyLarge)
  {
    const int dL = (int)(gL0 & 0xffff) + (int)(gL0 >> 16) + (int)(gL3 & 0xffff) + (int)(gL3 >> 16);
    if (dL < beta && (fl0 & 1) && (fl3 & 1))
    {
      dbfLongLine(xl, o, largeP ? lenP : 3, largeQ ? lenQ : 3, tc, wP, wQ);
      return;
    }
  }
  const int dp0 = gS0 & 0xffff, dq0 = gS0 >> 16, dp3 = gS3 & 0xffff, dq3 = gS3 >> 16;
  if (dp0 + dq0 + dp3 + dq3 < beta)
  {
    bool secondP = false, secondQ = false;
    if (lenP > 1 && lenQ > 1)
    {
      secondP = (dp0 + dp3) < sideThr;
      secondQ = (dq0 + dq3) < sideThr;
    }
    const bool strong = (fl0 & 2) && (fl3 & 2);
    dbfLumaLine(xl, o, tc, strong, wP, wQ, secondP, secondQ, maxv);
  }
}

// The same in two steps for SIXTEEN segments per warp round.  The decisions of xEdgeFilterLuma only look at the first and the last
// line of a segment (LoopFilter.cpp:977-1043), so a PAIR of lanes decides a segment -- lane parity = line 0 / line 3, the partner's
// gradients and strong-filter tests come by one shuffle -- instead of a quad of which two lanes computed for nothing; then the
// sixteen segments are filtered in two sub-rounds of eight, one line per lane, with the decision fetched from the pair by shuffle.
//   decision word: bits 0-1 = 0 nothing / 1 long filter / 2 normal strong / 3 normal weak, bit 2 secondP, bit 3 secondQ
__device__ __forceinline__ uint32_t dbfLumaDecidePair(const pel* x0, int o, int s, uint32_t rec)
{
  const int tc = rec & 0x7ff;
  const int beta = (rec >> VTMGPU_DBF_L_BETA_SHIFT) & 0x7ff;
  const pel* x = x0 + ((threadIdx.x & 1) ? 3 * s : 0);
  const int lenP = (rec >> VTMGPU_DBF_L_LENP_SHIFT) & 7, lenQ = (rec >> VTMGPU_DBF_L_LENQ_SHIFT) & 7;
  const bool largeP = lenP > 3 && !(rec & VTMGPU_DBF_L_CTUROW), largeQ = lenQ > 3;
  const int sideThr = (beta + (beta >> 1)) >> 3;
  const int dp = iabs(PK(2) - 2 * PK(1) + PK(0)), dq = iabs(QK(0) - 2 * QK(1) + QK(2));
  int dpL = dp, dqL = dq;
  if (largeP) dpL = (dp + iabs(PK(5) - 2 * PK(4) + PK(3)) + 1) >> 1;
  if (largeQ) dqL = (dq + iabs(QK(3) - 2 * QK(4) + QK(5)) + 1) >> 1;
  const bool anyLarge = largeP || largeQ;
  const bool sL = anyLarge && dbfStrongLong(x, o, 2 * (dpL + dqL), beta, tc, largeP, largeQ, lenP, lenQ);
  const bool sS = lenP > 2 && lenQ > 2 && dbfStrongShort(x, o, 2 * (dp + dq), beta, tc, false);
  const uint32_t gS = (uint32_t)dp | (uint32_t)dq << 16, gL = (uint32_t)dpL | (uint32_t)dqL << 16, fl = (sL ? 1u : 0u) | (sS ? 2u : 0u);
  const uint32_t gSo = __shfl_xor_sync(0xffffffffu, gS, 1), gLo = __shfl_xor_sync(0xffffffffu, gL, 1), flo = __shfl_xor_sync(0xffffffffu, fl, 1);
  if (anyLarge)
  {
    const int dL = (int)(gL & 0xffff) + (int)(gL >> 16) + (int)(gLo & 0xffff) + (int)(gLo >> 16);
    if (dL < beta && (fl & flo & 1)) return 1u;
  }
  const int dpS = (int)(gS & 0xffff) + (int)(gSo & 0xffff), dqS = (int)(gS >> 16) + (int)(gSo >> 16);
  if (dpS + dqS >= beta) return 0u;
  uint32_t d = (fl & flo & 2) ? 2u : 3u;
  if (lenP > 1 && lenQ > 1)
  {
    if (dpS < sideThr) d |= 4u;
    if (dqS < sideThr) d |= 8u;
  }
  return d;
}

__device__ __forceinline__ void dbfLumaApplyLine(pel* xl, int o, uint32_t rec, uint32_t d, int maxv, bool valid)
{
  const int tc = rec & 0x7ff;
  const int lenP = (rec >> VTMGPU_DBF_L_LENP_SHIFT) & 7, lenQ = (rec >> VTMGPU_DBF_L_LENQ_SHIFT) & 7;
  const bool wP = valid && !(rec & VTMGPU_DBF_L_PNOFILT), wQ = valid && !(rec & VTMGPU_DBF_L_QNOFILT);
  const uint32_t type = d & 3;
  if (type == 1)
  {
    const bool largeP = lenP > 3 && !(rec & VTMGPU_DBF_L_CTUROW), largeQ = lenQ > 3;
    dbfLongLine(xl, o, largeP ? lenP : 3, largeQ ? lenQ : 3, tc, wP, wQ);
  }
  else if (type) dbfLumaLine(xl, o, tc, type == 2, wP, wQ, (d & 4) != 0, (d & 8) != 0, maxv);
}

// LADF for a lane pair (see dbfLadfRecord): the two lanes hold lines 0 and 3
__device__ __noinline__ uint32_t dbfLadfRecordPair(uint32_t rec, const pel* x0, int o, int s, int bd, const LadfDev* ladf)
{
  const pel* x = x0 + ((threadIdx.x & 1) ? 3 * s : 0);
  const int own = PK(0) + QK(0);
  const int level = (own + __shfl_xor_sync(0xffffffffu, own, 1)) >> 2;
  int shift = ladf->off[0];
  for (int k = 1; k < ladf->n; k++)
  {
    if (level > ladf->lb[k]) shift = ladf->off[k];
    else break;
  }
  const int t = kDbfTcTable[clip3(0, 65, (int)(rec & 0x7ff) - VTMGPU_DBF_LADF_BIAS + shift)];
  const int tc = bd < 10 ? (t + 2) >> (10 - bd) : t << (bd - 10);
  const int beta = (int)kDbfBetaTable[clip3(0, 63, (int)((rec >> VTMGPU_DBF_L_BETA_SHIFT) & 0x7ff) - VTMGPU_DBF_LADF_BIAS + shift)] << (bd - 8);
  return (rec & ~0x3fffffu) | (uint32_t)tc | (uint32_t)beta << VTMGPU_DBF_L_BETA_SHIFT;
}

__device__ __forceinline__ void dbfChromaLine(pel* x, int o, int tc, bool strong, bool ctb, bool wP, bool wQ, int maxv)
{
  const int p0 = PK(0), p1 = PK(1), q0 = QK(0), q1 = QK(1);
  if (strong)
  {
    const int p2 = PK(2), p3 = PK(3), q2 = QK(2), q3 = QK(3);
    if (ctb)
    {
      if (wP) x[-1 * o] = (pel)clip3(p0 - tc, p0 + tc, (3 * p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3);
      if (wQ) x[0]      = (pel)clip3(q0 - tc, q0 + tc, (2 * p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3);
    }
    else
    {
      if (wP)
      {
        x[-3 * o] = (pel)clip3(p2 - tc, p2 + tc, (3 * p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3);
        x[-2 * o] = (pel)clip3(p1 - tc, p1 + tc, (2 * p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3);
        x[-1 * o] = (pel)clip3(p0 - tc, p0 + tc, (p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3);
      }
      if (wQ) x[0] = (pel)clip3(q0 - tc, q0 + tc, (p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3);
    }
    if (wQ)
    {
      x[1 * o] = (pel)clip3(q1 - tc, q1 + tc, (p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3);
      x[2 * o] = (pel)clip3(q2 - tc, q2 + tc, (p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3);
    }
    return;
  }
  const int delta = clip3(-tc, tc, (((q0 - p0) << 2) + p1 - q1 + 4) >> 3);
  if (wP) x[-1 * o] = (pel)clip3(0, maxv, p0 + delta);
  if (wQ) x[0]      = (pel)clip3(0, maxv, q0 - delta);
}

// one chroma segment of n (2 or 4) lines of one component
__device__ void dbfChromaSegment(pel* x0, int o, int s, int n, int tc, int beta, bool large, bool ctb, bool wP, bool wQ, int maxv)
{
  bool strong = false;
  if (large)
  {
    const int l3 = (n == 2 ? 1 : 3) * s;
    const pel* x = x0;
    const int dp0 = ctb ? iabs(PK(0) - PK(1)) : iabs(PK(2) - 2 * PK(1) + PK(0));
    const int dq0 = iabs(QK(0) - 2 * QK(1) + QK(2));
    x = x0 + l3;
    const int dp3 = ctb ? iabs(PK(0) - PK(1)) : iabs(PK(2) - 2 * PK(1) + PK(0));
    const int dq3 = iabs(QK(0) - 2 * QK(1) + QK(2));
    const int d0 = dp0 + dq0, d3 = dp3 + dq3;
    if (d0 + d3 < beta)
      strong = dbfStrongShort(x0, o, 2 * d0, beta, tc, ctb) && dbfStrongShort(x0 + l3, o, 2 * d3, beta, tc, ctb);
  }
  for (int i = 0; i < n; i++) dbfChromaLine(x0 + i * s, o, tc, strong, ctb, wP, wQ, maxv);
}
#undef PK
#undef QK

struct DbfLaunch
{
  int tilesXL, tilesL;      // luma tile grid (of the rows being filtered)
  int tilesXC, tilesC;      // per chroma plane
  int ty0L, ty0C;           // first tile row (band mode: a CTU-row band of the picture; 0 for whole pictures)
  int tilesLFull, tilesCFull;   // tiles of the WHOLE plane (the queues of a slot are indexed by picture tile: luma, Cb, Cr)
};

// ---- the kernel ---------------------------------------------------------------------------------------------------
// Deblocking (both passes) of one tile followed by SAO of the tile's own samples, written straight to the output plane.
// SAO classifies against the deblocked neighbours one sample outside the tile, so both passes are evaluated 4 samples
// beyond the tile on every side (luma: the vertical edges x0-4 and x0+TW+4 and the horizontal-edge segments of the
// columns x0-4..x0-1 / x0+TW..x0+TW+3; chroma: one more horizontal-edge segment on each side).  The 8-sample halo still
// suffices: a block side that starts 4 samples off a multiple of 16 is shorter than 32, so those edges read at most 4
// samples on their far side.
//
// Persistent CTAs walk the plane tiles of a batch of picture slots round robin (per slot: luma tiles, Cb tiles, Cr tiles).
// While tile i is filtered, tile i+1 arrives: the samples (tile + 8 halo, zero filled outside the picture) by ONE TMA box,
// the queues of the active segments of both passes by bulk copies (cp.async.bulk).
constexpr int DBF_RECL = DBF_TW / 4 + 8;                // columns of the luma record boxes (40 at tile width 128)
constexpr int DBF_RECC = DBF_TW / 8 + 2;                // columns of the chroma vertical-edge record box (18)
constexpr int dbfMax(int a, int b) { return a > b ? a : b; }
constexpr int dbfUp16(int v) { return (v + 15) / 16 * 16; }
// Per-tile queues of the ACTIVE segments, built once per picture by k_dbf_queues when the records are set (round 2: the kernel used to
// load the record boxes of a tile and every CTA compacted them itself -- 31 % of its shared-memory fill, 8 of its 53 instructions per
// pixel and one CTA barrier per tile).  Entry = 64 bits: the record in the low 48 bits (luma: 32), the slot index inside the tile's record
// box in the top 16.  Capacities = every slot of a pass active (luma / chroma, whichever is larger), multiples of 16 entries.
constexpr int DBF_QCAP1 = dbfUp16(dbfMax((DBF_TW / 4 + 3) * (DBF_SH / 4), (DBF_TW / 8 + 1) * (DBF_SH / 2)));        // 704 at width 128
constexpr int DBF_QCAP2 = dbfUp16(dbfMax((DBF_TH / 4 + 1) * (DBF_TW / 4 + 2), (DBF_TH / 8 + 1) * (DBF_TW / 2 + 2)));    // 608
constexpr int DBF_QTILE = DBF_QCAP1 + DBF_QCAP2;       // entries per tile in the queue memory of a slot
constexpr int DBF_TILE_BYTES = DBF_SH * DBF_PITCH * 2;  // 24320 (multiple of 128: TMA destination)
static_assert(DBF_TILE_BYTES % 128 == 0, "TMA destination alignment");
constexpr int DBF_STAGE_BYTES = DBF_TILE_BYTES + DBF_QTILE * 8;
constexpr int DBF_SMEM_BYTES = 2 * DBF_STAGE_BYTES + 32;      // stages, 2 mbarriers + the queue lengths of the two stages (4 x uint16... as 4 ints)

// record boxes of one tile (columns x rows, in records) for the arrays lumaV, lumaH, chromaV, chromaH.  TMA wants the first
// element of a box row on a 16-byte boundary and the row length a multiple of 16 bytes, so the luma boxes start 4 records
// left of the tile (3 unused + the edge x0-4) and the chroma horizontal-edge box 2 records left (1 unused + the extra segment)
struct DbfRecBoxes
{
  int cols[4], rows[4];
};

__host__ __device__ inline DbfRecBoxes dbfRecBoxes(int sx, int sy)
{
  DbfRecBoxes B;
  const int nv = 4 >> sy, nh = 4 >> sx;
  B.cols[0] = DBF_RECL; B.rows[0] = DBF_SH / 4;              // vertical edges x0-4 .. x0+TW+4 (columns 3..37 used), segment rows of tile + halo
  B.cols[1] = DBF_RECL; B.rows[1] = DBF_TH / 4 + 1;          // horizontal edges y0 .. y0+TH, segments of columns x0-4 .. x0+TW+3 (columns 3..36 used)
  B.cols[2] = DBF_RECC; B.rows[2] = DBF_SH / nv;                   // chroma vertical edges x0 .. x0+TW step 8 (17 used)
  B.cols[3] = DBF_TW / nh + 4; B.rows[3] = DBF_TH / 8 + 1;   // chroma horizontal edges, one extra segment each side (columns 1 .. TW/nh+2 used)
  return B;
}

struct DbfTile
{
  int comp, x0, y0;
};

__device__ __forceinline__ DbfTile dbfDecodeTile(int item, const DbfLaunch& L)
{
  DbfTile T;
  T.comp = 0;
  if (item >= L.tilesL) { item -= L.tilesL; T.comp = 1; if (item >= L.tilesC) { item -= L.tilesC; T.comp = 2; } }
  const int tilesX = T.comp ? L.tilesXC : L.tilesXL, ty = item / tilesX;
  T.x0 = (item - ty * tilesX) * DBF_TW;
  T.y0 = (ty + (T.comp ? L.ty0C : L.ty0L)) * DBF_TH;
  return T;
}

// slot counts of the two passes of a tile (work items i = 0 .. n-1; the record of item i sits at rec[i])
struct DbfPassGeom
{
  int ne1, ns1, n1;      // pass 1: ne1 edges per row of segments, ns1 segment rows
  int ne2, ns2, n2;      // pass 2: ne2 edge rows, ns2 segments per edge row
  int nv, nh;            // chroma: samples along the edge per record (vertical edges / horizontal edges)
  int p1, p2;            // row pitch (records) of the two record boxes in shared memory
};

__device__ __forceinline__ DbfPassGeom dbfPassGeom(int comp, const Geom& g)
{
  DbfPassGeom P;
  if (comp == 0) { P.ne1 = DBF_TW / 4 + 3; P.ns1 = DBF_SH / 4; P.ne2 = DBF_TH / 4 + 1; P.ns2 = DBF_TW / 4 + 2; P.nv = P.nh = 4; P.p1 = P.p2 = DBF_RECL; }
  else
  {
    P.nv = 4 >> g.sy; P.nh = 4 >> g.sx;
    P.ne1 = DBF_TW / 8 + 1; P.ns1 = DBF_SH / P.nv; P.ne2 = DBF_TH / 8 + 1; P.ns2 = DBF_TW / P.nh + 2;
    P.p1 = DBF_RECC; P.p2 = P.ns2 + 2;
  }
  P.n1 = P.p1 * P.ns1; P.n2 = P.p2 * P.ne2;                   // slots scanned (incl. the padding column of pass 1)
  return P;
}

// picture tile number of a tile (luma tiles, then Cb, then Cr) = index of its queues
__device__ __forceinline__ int dbfQueueTile(const DbfTile& T, const DbfLaunch& L)
{
  const int tx = T.x0 / DBF_TW, ty = T.y0 / DBF_TH;
  if (T.comp == 0) return ty * L.tilesXL + tx;
  return L.tilesLFull + (T.comp - 1) * L.tilesCFull + ty * L.tilesXC + tx;
}

// Builds the queues of one slot from its record arrays (after any of the three producers: uploaded arrays, scattered lists, k_dbf_derive):
// one warp per (tile, pass) scans the record slots the tile evaluates -- the box the kernel used to load: tile + the halo the passes
// are evaluated on, slot index i = row * box pitch + column as before -- and appends the active ones in ascending order.
struct DbfQueueArgs
{
  const uint32_t* lumaRec[2];
  const uint64_t* chromaRec[2];
  int recW[4], recH[4], recP[4];
  uint64_t* q;              // [tiles][DBF_QTILE]
  uint32_t* cnt;            // [tiles][2]
  DbfLaunch L;              // whole picture
  int sx, sy, ncomp;
};

__global__ void __launch_bounds__(256) k_dbf_queues(const DbfQueueArgs A)
{
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int tiles = A.L.tilesLFull + (A.ncomp > 1 ? 2 * A.L.tilesCFull : 0);
  if (warp >= 2 * tiles) return;
  const int tile = warp >> 1, pass = warp & 1;
  int comp = 0, t = tile;
  if (t >= A.L.tilesLFull) { t -= A.L.tilesLFull; comp = 1; if (t >= A.L.tilesCFull) { t -= A.L.tilesCFull; comp = 2; } }
  const int tilesX = comp ? A.L.tilesXC : A.L.tilesXL, ty = t / tilesX, x0 = (t - ty * tilesX) * DBF_TW, y0 = ty * DBF_TH;
  // box of the pass in the record array: origin (c0, r0), pitch p, rows n, used columns [u0, u1)
  int a, c0, r0, p, n, u0, u1;
  if (comp == 0)
  {
    a = pass; p = DBF_RECL; c0 = (x0 >> 2) - 4;
    if (pass == 0) { r0 = (y0 - DBF_HALO) >> 2; n = DBF_SH / 4; u0 = 3; u1 = 3 + DBF_TW / 4 + 3; }
    else           { r0 = y0 >> 2; n = DBF_TH / 4 + 1; u0 = 3; u1 = 3 + DBF_TW / 4 + 2; }
  }
  else
  {
    const int nvLog = 2 - A.sy, nhLog = 2 - A.sx, nh = 4 >> A.sx;
    a = 2 + pass;
    if (pass == 0) { p = DBF_RECC; c0 = x0 >> 3; r0 = (y0 - DBF_HALO) >> nvLog; n = DBF_SH >> nvLog; u0 = 0; u1 = DBF_RECC - 1; }
    else           { const int ns2 = DBF_TW / nh + 2; p = ns2 + 2; c0 = (x0 >> nhLog) - 2; r0 = y0 >> 3; n = DBF_TH / 8 + 1; u0 = 1; u1 = 1 + ns2; }
  }
  const int tcShift = comp == 2 ? VTMGPU_DBF_C_TCCR_SHIFT : 0;
  uint64_t* q = A.q + (size_t)tile * DBF_QTILE + (pass ? DBF_QCAP1 : 0);
  int count = 0;
  const int slots = n * p;
  for (int i0 = 0; i0 < slots; i0 += 128)
  {
    // four chunks of 32 slots at a time: the loads are independent of each other (the kernel is all load latency)
    uint64_t rec[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
      const int i = i0 + 32 * u + lane, r = i / p, c = i - r * p;
      rec[u] = 0;
      if (i < slots && c >= u0 && c < u1)
      {
        const int row = r0 + r, col = c0 + c;
        if (row >= 0 && row < A.recH[a] && col >= 0 && col < A.recW[a])
        {
          if (comp == 0) rec[u] = A.lumaRec[pass][(size_t)row * A.recP[a] + col];
          else           rec[u] = A.chromaRec[pass][(size_t)row * A.recP[a] + col];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
    {
      const bool act = ((rec[u] >> tcShift) & 0x7ff) != 0;
      const unsigned m = __ballot_sync(0xffffffffu, act);
      if (act) q[count + __popc(m & ((1u << lane) - 1))] = (rec[u] & 0xffffffffffffull) | (uint64_t)(i0 + 32 * u + lane) << 48;
      count += __popc(m);
    }
  }
  if (lane == 0) A.cnt[tile * 2 + pass] = (uint32_t)count;
}

// issues the asynchronous loads of one tile into a stage (one thread): the samples (tile + 8 halo, zero filled outside the picture) by
// one TMA box and the two queues of its active segments by bulk copies (their lengths were read one tile earlier), all on one mbarrier
__device__ __forceinline__ void dbfPrefetch(unsigned char* stageMem, uint64_t* bar, int* stageCnt, const CUtensorMap* planeMap, const uint64_t* q, int n1, int n2, const DbfTile& T)
{
  stageCnt[0] = n1; stageCnt[1] = n2;                       // published by the release of the arrive below
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t b1 = (uint32_t)((n1 + 1) & ~1) * 8, b2 = (uint32_t)((n2 + 1) & ~1) * 8;      // whole 16-byte units
  mbarExpectTx(bar, DBF_TILE_BYTES + b1 + b2);
  tmaLoad2D(stageMem, planeMap, T.x0 - DBF_HALO, T.y0 - DBF_HALO, bar);
  if (b1) bulkLoad(stageMem + DBF_TILE_BYTES, q, b1, bar);
  if (b2) bulkLoad(stageMem + DBF_TILE_BYTES + DBF_QCAP1 * 8, q + DBF_QCAP1, b2, bar);
}

// maps = TMA descriptors of the plane buffers: [slot][3 buffers][3 planes], box = DBF_PITCH x DBF_SH samples
__global__ void __launch_bounds__(DBF_THREADS, DBF_CTAS_PER_SM) k_dbf_sao(const SlotDev* __restrict__ slots, const CUtensorMap* __restrict__ tmaps,
                                                             int firstSlot, int numSlots, int srcBuf, int dstBuf, Geom g, DbfLaunch L, TileStep step, int doDbf, int doSao,
                                                             const BandDev band)
{
  extern __shared__ __align__(128) unsigned char smraw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + 2 * DBF_STAGE_BYTES);
  const int tid = threadIdx.x, lane = tid & 31;
  int* stageCnt = reinterpret_cast<int*>(smraw + 2 * DBF_STAGE_BYTES + 16);        // [stage][2]: queue lengths of the tile in the stage
  const int itemsPerSlot = L.tilesL + 2 * L.tilesC;
  int slot = blockIdx.x / itemsPerSlot, item = blockIdx.x - slot * itemsPerSlot;
  if (slot >= numSlots) return;
  if (tid == 0)
  {
    mbarInit(&bars[0], 1);
    mbarInit(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  DbfTile T = dbfDecodeTile(item, L), Tn = T;
  // queue lengths of a tile (thread 0): loaded one tile before the tile's loads are issued, i.e. two tiles ahead of its filtering
  auto queueLengths = [&](int sl, const DbfTile& t) -> uint2 {
    const SlotDev& Sq = slots[firstSlot + sl];
    if (!(doDbf && Sq.dbfOn)) return make_uint2(0u, 0u);
    return __ldg(reinterpret_cast<const uint2*>(Sq.dbfQCnt) + dbfQueueTile(t, L));
  };
  uint2 cntN = make_uint2(0u, 0u);             // lengths for the tile whose loads the NEXT iteration issues
  if (tid == 0)
  {
    const SlotDev& S = slots[firstSlot + slot];
    const uint2 c0 = queueLengths(slot, T);
    dbfPrefetch(smraw, &bars[0], stageCnt, tmaps + ((size_t)(firstSlot + slot) * 3 + srcBuf) * 3 + T.comp, S.dbfQ + (size_t)dbfQueueTile(T, L) * DBF_QTILE, (int)c0.x, (int)c0.y, T);
    int s1 = slot + step.dSlot, i1 = item + step.dItem;
    if (i1 >= itemsPerSlot) { i1 -= itemsPerSlot; s1++; }
    if (s1 < numSlots) cntN = queueLengths(s1, dbfDecodeTile(i1, L));
  }
  for (uint32_t it = 0; slot < numSlots; it++)
  {
    const int stage = it & 1;
    unsigned char* stageMem = smraw + stage * DBF_STAGE_BYTES;
    const SlotDev& S = slots[firstSlot + slot];
    int nslot = slot + step.dSlot, nitem = item + step.dItem;
    if (nitem >= itemsPerSlot) { nitem -= itemsPerSlot; nslot++; }
    if (nslot < numSlots)
    {
      Tn = dbfDecodeTile(nitem, L);
      if (tid == 0)
      {
        const SlotDev& Sn = slots[firstSlot + nslot];
        dbfPrefetch(smraw + (stage ^ 1) * DBF_STAGE_BYTES, &bars[stage ^ 1], stageCnt + 2 * (stage ^ 1), tmaps + ((size_t)(firstSlot + nslot) * 3 + srcBuf) * 3 + Tn.comp,
                    Sn.dbfQ + (size_t)dbfQueueTile(Tn, L) * DBF_QTILE, (int)cntN.x, (int)cntN.y, Tn);
        int s2 = nslot + step.dSlot, i2 = nitem + step.dItem;
        if (i2 >= itemsPerSlot) { i2 -= itemsPerSlot; s2++; }
        cntN = s2 < numSlots ? queueLengths(s2, dbfDecodeTile(i2, L)) : make_uint2(0u, 0u);
      }
    }

    const int comp = T.comp, x0 = T.x0, y0 = T.y0;
    const PlaneDev dst = S.buf[dstBuf][comp];
    const int w = dst.w, h = dst.h;
    // SAO parameters of this thread's strip (one 8-sample group column x 4 rows; a 4-row strip never crosses a CTU boundary):
    // loaded now, consumed after the deblocking passes
    const int gcol = (tid & 7) | ((tid >> 4) & 8), rb = (tid >> 3) & 15;     // warp = 8 group columns x 4 row blocks: one CTU for every plane
    const int sx_ = x0 + 8 * gcol, sy_ = y0 + 4 * rb;
    const int cwLog = g.ctuLog2 - (comp ? g.sx : 0), chLog = g.ctuLog2 - (comp ? g.sy : 0);
    uint4 pq = make_uint4(0, 0, 0, 0);
    if (doSao && S.saoOn && sx_ < w && sy_ < h) pq = __ldg(reinterpret_cast<const uint4*>(&S.sao[((sy_ >> chLog) * g.wCtus + (sx_ >> cwLog)) * 3 + comp]));

    mbarWait(&bars[stage], (it >> 1) & 1);                   // samples and records of this tile are in shared memory
    pel* sm = reinterpret_cast<pel*>(stageMem);

    if (comp == 0 && srcBuf == 0 && S.lmcsOn)
    {
      // LMCS inverse mapping of the reconstruction (AreaBuf<Pel>::rspSignal, Buffer.cpp:380-393, called by executeLoopFilters before
      // the deblocking, DecLib.cpp:570-577), folded into the tile load: every luma sample of tile + halo goes through the table
      // (1 << bit depth entries, read through L1).  Samples outside the picture are never used.
      const uint16_t* __restrict__ lut = reinterpret_cast<const uint16_t*>(S.lmcs);
      uint4* t = reinterpret_cast<uint4*>(sm);
      for (int i = tid; i < DBF_SH * (DBF_PITCH / 8); i += DBF_THREADS)
      {
        uint4 v = t[i];
        uint32_t* w = &v.x;
#pragma unroll
        for (int k = 0; k < 4; k++) w[k] = (uint32_t)__ldg(&lut[w[k] & 0xffff]) | (uint32_t)__ldg(&lut[w[k] >> 16]) << 16;
        t[i] = v;
      }
      __syncthreads();
    }

    if (doDbf && S.dbfOn)
    {
      const int maxv = (1 << (comp ? g.bdC : g.bdL)) - 1;
      const DbfPassGeom P = dbfPassGeom(comp, g);
      const LadfDev* const ladf = S.ladf.n > 0 ? &S.ladf : nullptr;
      if (comp == 0)
      {
        const uint64_t* qa = reinterpret_cast<const uint64_t*>(stageMem + DBF_TILE_BYTES);
        const uint64_t* qb = qa + DBF_QCAP1;
        // 16 segments per warp round: decisions by lane pairs, filters by quads
        constexpr int NE = DBF_RECL;                            // pitch of the slot numbering (record box columns; edge e sits in column e + 3)
        constexpr int NSH = DBF_RECL;
        // pass 1: vertical edges x0-4 .. x0+TW+4 (step 4), all rows of tile + halo
        {
          const int cnt = stageCnt[2 * stage];
          // few segments (the usual case: ~60 per pass and tile): eight per warp round so that all warps take part; many: sixteen
          const int per = cnt > 8 * (DBF_THREADS / 32) ? 16 : 8;
          for (int k0 = (tid >> 5) * per; k0 < cnt; k0 += (DBF_THREADS / 32) * per)
          {
            // decisions: a lane pair per segment
            const int k = k0 + (lane >> 1);
            const bool valid = k < cnt && (lane >> 1) < per;
            const uint64_t ent = qa[valid ? k : k0];
            const int i = (int)(ent >> 48), sg = i / NE, e = i - sg * NE - 3;
            const int segOff = (4 * sg) * DBF_PITCH + DBF_HALO - 4 + 4 * e;
            uint32_t rec = (uint32_t)ent;
            if (ladf) rec = dbfLadfRecordPair(rec, &sm[segOff], 1, DBF_PITCH, g.bdL, ladf);
            uint32_t d = dbfLumaDecidePair(&sm[segOff], 1, DBF_PITCH, rec);
            if (!valid || !(rec & 0x7ff)) d = 0;
            // filters: a quad per segment, eight segments per sub-round
#pragma unroll 1
            for (int r = 0; r < (per >> 3); r++)
            {
              const int src = 2 * (r * 8 + (lane >> 2));
              const uint32_t dj = __shfl_sync(0xffffffffu, d, src), recj = __shfl_sync(0xffffffffu, rec, src);
              const int offj = __shfl_sync(0xffffffffu, segOff, src);
              dbfLumaApplyLine(&sm[offj + (lane & 3) * DBF_PITCH], 1, recj, dj, maxv, dj != 0);
            }
          }
        }
        __syncthreads();
        // pass 2: horizontal edges y0 .. y0+TH (step 4), columns x0-4 .. x0+TW+3
        {
          const int cnt = stageCnt[2 * stage + 1];
          const int per = cnt > 8 * (DBF_THREADS / 32) ? 16 : 8;
          for (int k0 = (tid >> 5) * per; k0 < cnt; k0 += (DBF_THREADS / 32) * per)
          {
            const int k = k0 + (lane >> 1);
            const bool valid = k < cnt && (lane >> 1) < per;
            const uint64_t ent = qb[valid ? k : k0];
            const int i = (int)(ent >> 48), e = i / NSH, sg = i - e * NSH - 3;
            const int segOff = (DBF_HALO + 4 * e) * DBF_PITCH + DBF_HALO - 4 + 4 * sg;
            uint32_t rec = (uint32_t)ent;
            if (ladf) rec = dbfLadfRecordPair(rec, &sm[segOff], DBF_PITCH, 1, g.bdL, ladf);
            uint32_t d = dbfLumaDecidePair(&sm[segOff], DBF_PITCH, 1, rec);
            if (!valid || !(rec & 0x7ff)) d = 0;
#pragma unroll 1
            for (int r = 0; r < (per >> 3); r++)
            {
              const int src = 2 * (r * 8 + (lane >> 2));
              const uint32_t dj = __shfl_sync(0xffffffffu, d, src), recj = __shfl_sync(0xffffffffu, rec, src);
              const int offj = __shfl_sync(0xffffffffu, segOff, src);
              dbfLumaApplyLine(&sm[offj + (lane & 3)], DBF_PITCH, recj, dj, maxv, dj != 0);
            }
          }
        }
      }
      else
      {
        const uint64_t* qa = reinterpret_cast<const uint64_t*>(stageMem + DBF_TILE_BYTES);
        const uint64_t* qb = qa + DBF_QCAP1;
        const int c = comp - 1;
        const int tcShift = c ? VTMGPU_DBF_C_TCCR_SHIFT : 0, betaShift = c ? VTMGPU_DBF_C_BETACR_SHIFT : VTMGPU_DBF_C_BETACB_SHIFT;
        const int p2 = P.p2;
        // pass 1: vertical edges on the 8-sample chroma grid; one item = the chroma rows of one 4-luma-row unit
        {
          const int cnt = stageCnt[2 * stage];
          for (int k = tid; k < cnt; k += DBF_THREADS)
          {
            const uint64_t rec = qa[k];
            const int i = (int)(rec >> 48), sg = i / DBF_RECC, e = i - sg * DBF_RECC;
            dbfChromaSegment(&sm[(P.nv * sg) * DBF_PITCH + DBF_HALO + 8 * e], 1, DBF_PITCH, P.nv, (int)(rec >> tcShift) & 0x7ff, (int)(rec >> betaShift) & 0x7ff,
                             (rec & VTMGPU_DBF_C_LARGE) != 0, (rec & VTMGPU_DBF_C_CTB) != 0, !(rec & VTMGPU_DBF_C_PNOFILT), !(rec & VTMGPU_DBF_C_QNOFILT), maxv);
          }
        }
        __syncthreads();
        // pass 2: horizontal edges, one more segment of columns on each side of the tile
        {
          const int cnt = stageCnt[2 * stage + 1];
          for (int k = tid; k < cnt; k += DBF_THREADS)
          {
            const uint64_t rec = qb[k];
            const int i = (int)(rec >> 48), e = i / p2, sg = i - e * p2 - 1;
            dbfChromaSegment(&sm[(DBF_HALO + 8 * e) * DBF_PITCH + DBF_HALO - P.nh + P.nh * sg], DBF_PITCH, 1, P.nh, (int)(rec >> tcShift) & 0x7ff,
                             (int)(rec >> betaShift) & 0x7ff, (rec & VTMGPU_DBF_C_LARGE) != 0, (rec & VTMGPU_DBF_C_CTB) != 0, !(rec & VTMGPU_DBF_C_PNOFILT),
                             !(rec & VTMGPU_DBF_C_QNOFILT), maxv);
          }
        }
      }
      __syncthreads();
    }

    // ---- epilogue: SAO of the own region straight into the output plane (a plain 128-bit copy where SAO is off) -------
    if (sx_ < w && sy_ < h)
    {
      const pel* a = &sm[(DBF_HALO + 4 * rb) * DBF_PITCH + DBF_HALO + 8 * gcol];
      pel* o = dst.p + (size_t)sy_ * dst.pitch + sx_;
      const int nrows = min(4, h - sy_);
      if ((pq.x & 0xff) == 0)
      {
        // SAO off for this CTU (the common case): 128-bit copy of the strip
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (k < nrows) *reinterpret_cast<uint4*>(o + (size_t)k * dst.pitch) = *reinterpret_cast<const uint4*>(a + k * DBF_PITCH);
      }
      else saoStrip(o, dst.pitch, a, DBF_PITCH, nrows, sx_, sy_, pq, w, h, cwLog | chLog << 8 | (comp ? g.bdC : g.bdL) << 16,
                    (S.vbSao.nv | S.vbSao.nh) ? &S.vbSao : nullptr, comp ? (g.sx | g.sy << 8) : 0);
      if (band.myFlags)
      {
        // peer band mode: the first / last four rows of the band also go into the neighbour's plane (same layout, peer-mapped)
        const int sh = comp ? g.sy : 0, bb = band.rowBegin >> sh, be = min(band.rowEnd >> sh, h);
#pragma unroll 1
        for (int side = 0; side < 2; side++)
        {
          if (!band.peerPlanes[side] || sy_ != (side ? be - 4 : bb)) continue;
          // the neighbour's ALF of the previous iteration must have finished reading the rows this store overwrites
          while (ldAcquireSys(&band.myFlags[2 + side]) + 1 < band.iter) { }
          pel* po = band.peerPlanes[side] + (o - band.myPlanes);
          if ((pq.x & 0xff) == 0)
          {
#pragma unroll
            for (int k = 0; k < 4; k++)
              if (k < nrows) *reinterpret_cast<uint4*>(po + (size_t)k * dst.pitch) = *reinterpret_cast<const uint4*>(a + k * DBF_PITCH);
          }
          else saoStrip(po, dst.pitch, a, DBF_PITCH, nrows, sx_, sy_, pq, w, h, cwLog | chLog << 8 | (comp ? g.bdC : g.bdL) << 16,
                        (S.vbSao.nv | S.vbSao.nh) ? &S.vbSao : nullptr, comp ? (g.sx | g.sy << 8) : 0);
          __threadfence_system();
        }
      }
    }
    __syncthreads();                                         // the stage is free for the load after next
    slot = nslot; item = nitem; T = Tn;
  }
  if (band.myFlags)
  {
    // the last CTA to get here tells both neighbours that their halo rows of this iteration have landed
    __syncthreads();
    if (tid == 0)
    {
      __threadfence_system();
      if (atomicAdd(&band.myFlags[4], 1u) == gridDim.x - 1)
      {
        band.myFlags[4] = 0;
        __threadfence_system();
        if (band.peerFlags[0]) stReleaseSys(&band.peerFlags[0][1], band.iter);      // above: "rows from below have landed"
        if (band.peerFlags[1]) stReleaseSys(&band.peerFlags[1][0], band.iter);      // below: "rows from above have landed"
      }
    }
  }
}

}   // namespace vtmgpu
