/*
 * oracle/vvc_filters_oracle.c -- TEST INFRASTRUCTURE.  NOT product code, never on the product path.
 *
 * Plain-C, single-threaded, sample-at-a-time RESTATEMENT of the arithmetic of the reference's
 * in-loop filter chain (VTM 7.3+ snapshot), consuming exactly the flattened side information of
 * include/vtmgpu.h.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this library, and only as the checker / CPU baseline.
 *
 * Pinning: this restatement is checked picture-by-picture and stage-by-stage against the real
 * reference classes (RefLoopFilter / RefSampleAdaptiveOffset / RefAdaptiveLoopFilter = the
 * reference's own LoopFilter.cpp / SampleAdaptiveOffset.cpp / AdaptiveLoopFilter.cpp compiled
 * unmodified, see oracle/Makefile) run in this container on reference-encoded streams; the
 * resulting stage outputs + decoded-picture MD5s are committed under tests/golden/ together with
 * the generating script (tools/make_golden.py).  tests/test_oracle_golden.py re-checks it on CPU.
 *
 * Reference locations restated here (paths under source/Lib/CommonLib/):
 *   luma edge filter      LoopFilter.cpp:971-1080 (decisions), :1302-1485 (filters), :1566-1667
 *   chroma edge filter    LoopFilter.cpp:1246-1279, :1497-1555
 *   SAO                   SampleAdaptiveOffset.cpp:148-171, :230-290 (params), :293-547 (offsetBlock)
 *   ALF coefficients      AdaptiveLoopFilter.cpp:651-713, :743-762, :792-807
 *   ALF classification    AdaptiveLoopFilter.cpp:873-1082
 *   ALF 7x7 / 5x5         AdaptiveLoopFilter.cpp:1084-1324
 *   CC-ALF                AdaptiveLoopFilter.cpp:1327-1416
 */
#include <stdlib.h>
#include <string.h>

#include "vtmgpu.h"
#include "vvc_alf_fixed_tables.h"

typedef int16_t pel;

static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int sgn(int v) { return (v > 0) - (v < 0); }

/* ------------------------------------------------------------------------------------------------
 * deblocking
 * ---------------------------------------------------------------------------------------------- */

/* x points at q0 of one line; o = step across the edge (towards Q); P(k) = x[-(k+1)*o], Q(k) = x[k*o] */
#define PK(k) ((int)x[-((k) + 1) * o])
#define QK(k) ((int)x[(k) * o])

static int dbf_strong_short(const pel* x, ptrdiff_t o, int d, int beta, int tc, int p_one_sample)
{
  /* LoopFilter.cpp:1566-1579,1649 ; p_one_sample = chroma horizontal CTB boundary variant (:1574) */
  int sp3 = p_one_sample ? iabs(PK(1) - PK(0)) : iabs(PK(3) - PK(0));
  int sq3 = iabs(QK(3) - QK(0));
  return (sp3 + sq3 < (beta >> 3)) && (d < (beta >> 2)) && (iabs(PK(0) - QK(0)) < ((tc * 5 + 1) >> 1));
}

static int dbf_strong_long(const pel* x, ptrdiff_t o, int d, int beta, int tc, int largeP, int largeQ, int lenP, int lenQ)
{
  /* LoopFilter.cpp:1581-1618 (JVET_Q0054 variant) */
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

static void dbf_long_filter(pel* x, ptrdiff_t o, int nP, int nQ, int tc, int writeP, int writeQ)
{
  /* LoopFilter.cpp:1302-1395 */
  static const int c7[7] = { 59, 50, 41, 32, 23, 14, 5 }, c5[5] = { 58, 45, 32, 19, 6 }, c3[3] = { 53, 32, 11 };
  static const int t7[7] = { 6, 5, 4, 3, 2, 1, 1 }, t3[3] = { 6, 4, 2 };
  int p[8], q[8];
  for (int k = 0; k < 8; k++) { p[k] = PK(k); q[k] = QK(k); }
  const int* cP = nP == 7 ? c7 : (nP == 5 ? c5 : c3);
  const int* cQ = nQ == 7 ? c7 : (nQ == 5 ? c5 : c3);
  const int* tP = nP == 3 ? t3 : t7;
  const int* tQ = nQ == 3 ? t3 : t7;
  int refP = (p[nP - 1] + p[nP] + 1) >> 1;
  int refQ = (q[nQ - 1] + q[nQ] + 1) >> 1;
  int mid;
  if (nP == nQ)
  {
    if (nP == 5)
      mid = (2 * (p[0] + q[0] + p[1] + q[1] + p[2] + q[2]) + p[3] + q[3] + p[4] + q[4] + 8) >> 4;
    else
      mid = (2 * (p[0] + q[0]) + p[1] + q[1] + p[2] + q[2] + p[3] + q[3] + p[4] + q[4] + p[5] + q[5] + p[6] + q[6] + 8) >> 4;
  }
  else
  {
    int nL = nP > nQ ? nP : nQ, nS = nP > nQ ? nQ : nP;
    const int* L = nP > nQ ? p : q;
    const int* S = nP > nQ ? q : p;
    if (nL == 7 && nS == 5)
      mid = (2 * (p[0] + q[0] + p[1] + q[1]) + p[2] + q[2] + p[3] + q[3] + p[4] + q[4] + p[5] + q[5] + 8) >> 4;
    else if (nL == 7 && nS == 3)
      mid = (2 * (L[0] + S[0]) + S[0] + 2 * (S[1] + S[2]) + L[1] + S[1] + L[2] + L[3] + L[4] + L[5] + L[6] + 8) >> 4;
    else
      mid = (p[0] + q[0] + p[1] + q[1] + p[2] + q[2] + p[3] + q[3] + 4) >> 3;
  }
  if (writeP)
    for (int k = 0; k < nP; k++)
    {
      int cv = (tc * tP[k]) >> 1;
      x[-(k + 1) * o] = (pel)clip3(p[k] - cv, p[k] + cv, (mid * cP[k] + refP * (64 - cP[k]) + 32) >> 6);
    }
  if (writeQ)
    for (int k = 0; k < nQ; k++)
    {
      int cv = (tc * tQ[k]) >> 1;
      x[k * o] = (pel)clip3(q[k] - cv, q[k] + cv, (mid * cQ[k] + refQ * (64 - cQ[k]) + 32) >> 6);
    }
}

static void dbf_luma_line(pel* x, ptrdiff_t o, int tc, int strong, int writeP, int writeQ, int secondP, int secondQ, int maxv)
{
  /* LoopFilter.cpp:1397-1456, short filters */
  const int p0 = PK(0), p1 = PK(1), p2 = PK(2), p3 = PK(3), q0 = QK(0), q1 = QK(1), q2 = QK(2), q3 = QK(3);
  if (strong)
  {
    if (writeP)
    {
      x[-1 * o] = (pel)clip3(p0 - 3 * tc, p0 + 3 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      x[-2 * o] = (pel)clip3(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
      x[-3 * o] = (pel)clip3(p2 - tc, p2 + tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
    }
    if (writeQ)
    {
      x[0]      = (pel)clip3(q0 - 3 * tc, q0 + 3 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      x[1 * o]  = (pel)clip3(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
      x[2 * o]  = (pel)clip3(q2 - tc, q2 + tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
    }
    return;
  }
  int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
  if (iabs(delta) >= tc * 10) return;
  delta = clip3(-tc, tc, delta);
  const int tc2 = tc >> 1;
  if (writeP)
  {
    x[-1 * o] = (pel)clip3(0, maxv, p0 + delta);
    if (secondP) x[-2 * o] = (pel)clip3(0, maxv, p1 + clip3(-tc2, tc2, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1));
  }
  if (writeQ)
  {
    x[0] = (pel)clip3(0, maxv, q0 - delta);
    if (secondQ) x[1 * o] = (pel)clip3(0, maxv, q1 + clip3(-tc2, tc2, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1));
  }
}

/* one 4-line luma segment; x = q0 of line 0, o = across, s = along */
/* tc / beta tables (LoopFilter.cpp:66-74) */
static const uint16_t tc_table[66] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,3,4,4,4,4,5,5,5,5,7,7,8,9,10,10,11,13,14,15,17,19,21,24,25,29,33,36,
                                       41,45,51,57,64,71,80,89,100,112,125,141,157,177,198,222,250,280,314,352,395 };
static const uint8_t beta_table[64] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,
                                        42,44,46,48,50,52,54,56,58,60,62,64,66,68,70,72,74,76,78,80,82,84,86,88 };

static void dbf_luma_segment(pel* x0, ptrdiff_t o, ptrdiff_t s, uint32_t rec, int bd, const vtmgpu_ladf* ladf)
{
  int tc = rec & 0x7ff;
  if (!tc) return;
  int beta = (rec >> VTMGPU_DBF_L_BETA_SHIFT) & 0x7ff;
  if (ladf)
  {
    /* deriveLADFShift (LoopFilter.cpp:815-841) on the picture in its current state, then :971-975 */
    const int level = (x0[0] + x0[3 * s] + x0[-o] + x0[3 * s - o]) >> 2;
    int shift = ladf->qp_offset[0];
    for (int k = 1; k < ladf->num_intervals; k++)
    {
      if (level > ladf->lower_bound[k]) shift = ladf->qp_offset[k];
      else break;
    }
    const int t = tc_table[clip3(0, 65, tc - VTMGPU_DBF_LADF_BIAS + shift)];
    tc = bd < 10 ? (t + 2) >> (10 - bd) : t << (bd - 10);
    beta = beta_table[clip3(0, 63, beta - VTMGPU_DBF_LADF_BIAS + shift)] << (bd - 8);
    if (!tc) return;
  }
  const int lenP = (rec >> VTMGPU_DBF_L_LENP_SHIFT) & 7, lenQ = (rec >> VTMGPU_DBF_L_LENQ_SHIFT) & 7;
  const int writeP = !(rec & VTMGPU_DBF_L_PNOFILT), writeQ = !(rec & VTMGPU_DBF_L_QNOFILT);
  const int largeP = lenP > 3 && !(rec & VTMGPU_DBF_L_CTUROW), largeQ = lenQ > 3;
  const int side_thr = (beta + (beta >> 1)) >> 3;
  const int maxv = (1 << bd) - 1;
  const pel* x; /* used by PK/QK */
  x = x0;
  const int dp0 = iabs(PK(2) - 2 * PK(1) + PK(0)), dq0 = iabs(QK(0) - 2 * QK(1) + QK(2));
  x = x0 + 3 * s;
  const int dp3 = iabs(PK(2) - 2 * PK(1) + PK(0)), dq3 = iabs(QK(0) - 2 * QK(1) + QK(2));

  if (largeP || largeQ)
  {
    int dp0L = dp0, dq0L = dq0, dp3L = dp3, dq3L = dq3;
    if (largeP)
    {
      x = x0;         dp0L = (dp0L + iabs(PK(5) - 2 * PK(4) + PK(3)) + 1) >> 1;
      x = x0 + 3 * s; dp3L = (dp3L + iabs(PK(5) - 2 * PK(4) + PK(3)) + 1) >> 1;
    }
    if (largeQ)
    {
      x = x0;         dq0L = (dq0L + iabs(QK(3) - 2 * QK(4) + QK(5)) + 1) >> 1;
      x = x0 + 3 * s; dq3L = (dq3L + iabs(QK(3) - 2 * QK(4) + QK(5)) + 1) >> 1;
    }
    const int d0L = dp0L + dq0L, d3L = dp3L + dq3L;
    if (d0L + d3L < beta)
    {
      if (dbf_strong_long(x0, o, 2 * d0L, beta, tc, largeP, largeQ, lenP, lenQ) &&
          dbf_strong_long(x0 + 3 * s, o, 2 * d3L, beta, tc, largeP, largeQ, lenP, lenQ))
      {
        for (int i = 0; i < 4; i++)
          dbf_long_filter(x0 + i * s, o, largeP ? lenP : 3, largeQ ? lenQ : 3, tc, writeP, writeQ);
        return;
      }
    }
  }
  const int d0 = dp0 + dq0, d3 = dp3 + dq3;
  if (d0 + d3 < beta)
  {
    int secondP = 0, secondQ = 0, strong = 0;
    if (lenP > 1 && lenQ > 1)
    {
      secondP = (dp0 + dp3) < side_thr;
      secondQ = (dq0 + dq3) < side_thr;
    }
    if (lenP > 2 && lenQ > 2)
      strong = dbf_strong_short(x0, o, 2 * d0, beta, tc, 0) && dbf_strong_short(x0 + 3 * s, o, 2 * d3, beta, tc, 0);
    for (int i = 0; i < 4; i++)
      dbf_luma_line(x0 + i * s, o, tc, strong, writeP, writeQ, secondP, secondQ, maxv);
  }
}

static void dbf_chroma_line(pel* x, ptrdiff_t o, int tc, int strong, int ctb, int large, int writeP, int writeQ, int maxv)
{
  /* LoopFilter.cpp:1497-1555 */
  const int p0 = PK(0), p1 = PK(1), p2 = PK(2), p3 = PK(3), q0 = QK(0), q1 = QK(1), q2 = QK(2), q3 = QK(3);
  (void)large;
  if (strong)
  {
    if (ctb)
    {
      if (writeP) x[-1 * o] = (pel)clip3(p0 - tc, p0 + tc, (3 * p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3);
      if (writeQ)
      {
        x[0]     = (pel)clip3(q0 - tc, q0 + tc, (2 * p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3);
        x[1 * o] = (pel)clip3(q1 - tc, q1 + tc, (p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3);
        x[2 * o] = (pel)clip3(q2 - tc, q2 + tc, (p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3);
      }
    }
    else
    {
      if (writeP)
      {
        x[-3 * o] = (pel)clip3(p2 - tc, p2 + tc, (3 * p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3);
        x[-2 * o] = (pel)clip3(p1 - tc, p1 + tc, (2 * p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3);
        x[-1 * o] = (pel)clip3(p0 - tc, p0 + tc, (p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3);
      }
      if (writeQ)
      {
        x[0]     = (pel)clip3(q0 - tc, q0 + tc, (p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3);
        x[1 * o] = (pel)clip3(q1 - tc, q1 + tc, (p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3);
        x[2 * o] = (pel)clip3(q2 - tc, q2 + tc, (p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3);
      }
    }
    return;
  }
  const int delta = clip3(-tc, tc, (((q0 - p0) << 2) + p1 - q1 + 4) >> 3);
  if (writeP) x[-1 * o] = (pel)clip3(0, maxv, p0 + delta);
  if (writeQ) x[0]      = (pel)clip3(0, maxv, q0 - delta);
}

/* one chroma segment of n (2 or 4) lines of one component */
static void dbf_chroma_segment(pel* x0, ptrdiff_t o, ptrdiff_t s, int n, int tc, int beta, int large, int ctb, int writeP, int writeQ, int bd)
{
  if (!tc) return;
  const int maxv = (1 << bd) - 1;
  int strong = 0;
  if (large)
  {
    /* LoopFilter.cpp:1246-1272 ; second decision line is +1 when the segment has 2 lines, else +3 */
    const ptrdiff_t l3 = (n == 2 ? 1 : 3) * s;
    const pel* x = x0;
    const int dp0 = ctb ? iabs(PK(1) - 2 * PK(1) + PK(0)) : iabs(PK(2) - 2 * PK(1) + PK(0));
    const int dq0 = iabs(QK(0) - 2 * QK(1) + QK(2));
    x = x0 + l3;
    const int dp3 = ctb ? iabs(PK(1) - 2 * PK(1) + PK(0)) : iabs(PK(2) - 2 * PK(1) + PK(0));
    const int dq3 = iabs(QK(0) - 2 * QK(1) + QK(2));
    const int d0 = dp0 + dq0, d3 = dp3 + dq3;
    if (d0 + d3 < beta)
    {
      strong = dbf_strong_short(x0, o, 2 * d0, beta, tc, ctb) && dbf_strong_short(x0 + l3, o, 2 * d3, beta, tc, ctb);
      for (int i = 0; i < n; i++) dbf_chroma_line(x0 + i * s, o, tc, strong, ctb, large, writeP, writeQ, maxv);
      return;
    }
  }
  for (int i = 0; i < n; i++) dbf_chroma_line(x0 + i * s, o, tc, 0, ctb, large, writeP, writeQ, maxv);
}
#undef PK
#undef QK

static void chroma_shifts(int cf, int* sx, int* sy)
{
  *sx = (cf == 1 || cf == 2) ? 1 : 0;
  *sy = (cf == 1) ? 1 : 0;
}

/* LoopFilter::loopFilterPic (LoopFilter.cpp:145): all vertical edges, then all horizontal edges, in place */
int vvco_deblock(int16_t* const plane[3], const ptrdiff_t stride[3], int width, int height, int chroma_format,
                 int bd_luma, int bd_chroma, const vtmgpu_deblock_params* p)
{
  int sx, sy;
  chroma_shifts(chroma_format, &sx, &sy);
  const int uw = width / 4, uh = height / 4;
  for (int dir = 0; dir < 2; dir++)
  {
    const uint32_t* rl = p->luma[dir];
    for (int uy = 0; uy < uh; uy++)
      for (int ux = 0; ux < uw; ux++)
      {
        pel* x = plane[0] + (ptrdiff_t)(uy * 4) * stride[0] + ux * 4;
        if (dir == 0) dbf_luma_segment(x, 1, stride[0], rl[uy * uw + ux], bd_luma, p->ladf);
        else          dbf_luma_segment(x, stride[0], 1, rl[uy * uw + ux], bd_luma, p->ladf);
      }
    if (chroma_format == 0 || !p->chroma[dir]) continue;
    const uint64_t* rc = p->chroma[dir];
    const int gx = 8 << sx, gy = 8 << sy;                       /* chroma grid pitch in luma samples */
    if (dir == 0)
    {
      const int cols = (width + gx - 1) / gx, n = 4 >> sy;
      for (int uy = 0; uy < uh; uy++)
        for (int e = 0; e < cols; e++)
        {
          const uint64_t r = rc[uy * cols + e];
          if (!r) continue;
          for (int c = 0; c < 2; c++)
          {
            const int tc = (int)((r >> (c ? VTMGPU_DBF_C_TCCR_SHIFT : 0)) & 0x7ff);
            const int beta = (int)((r >> (c ? VTMGPU_DBF_C_BETACR_SHIFT : VTMGPU_DBF_C_BETACB_SHIFT)) & 0x7ff);
            pel* x = plane[1 + c] + (ptrdiff_t)((uy * 4) >> sy) * stride[1 + c] + ((e * gx) >> sx);
            dbf_chroma_segment(x, 1, stride[1 + c], n, tc, beta, !!(r & VTMGPU_DBF_C_LARGE), !!(r & VTMGPU_DBF_C_CTB),
                               !(r & VTMGPU_DBF_C_PNOFILT), !(r & VTMGPU_DBF_C_QNOFILT), bd_chroma);
          }
        }
    }
    else
    {
      const int rows = (height + gy - 1) / gy, n = 4 >> sx;
      for (int e = 0; e < rows; e++)
        for (int ux = 0; ux < uw; ux++)
        {
          const uint64_t r = rc[e * uw + ux];
          if (!r) continue;
          for (int c = 0; c < 2; c++)
          {
            const int tc = (int)((r >> (c ? VTMGPU_DBF_C_TCCR_SHIFT : 0)) & 0x7ff);
            const int beta = (int)((r >> (c ? VTMGPU_DBF_C_BETACR_SHIFT : VTMGPU_DBF_C_BETACB_SHIFT)) & 0x7ff);
            pel* x = plane[1 + c] + (ptrdiff_t)((e * gy) >> sy) * stride[1 + c] + ((ux * 4) >> sx);
            dbf_chroma_segment(x, stride[1 + c], 1, n, tc, beta, !!(r & VTMGPU_DBF_C_LARGE), !!(r & VTMGPU_DBF_C_CTB),
                               !(r & VTMGPU_DBF_C_PNOFILT), !(r & VTMGPU_DBF_C_QNOFILT), bd_chroma);
          }
        }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * SAO
 * ---------------------------------------------------------------------------------------------- */

/* xReconstructBlkSAOParams (SampleAdaptiveOffset.cpp:266-290) with invertQuantOffsets (:148-171) */
int vvco_sao_reconstruct(vtmgpu_sao_ctu* ctu, int num_ctus, int width_in_ctus, int num_comps, int log2_scale_luma, int log2_scale_chroma)
{
  int mask = 0;
  for (int a = 0; a < num_ctus; a++)
  {
    for (int c = 0; c < num_comps; c++)
    {
      vtmgpu_sao_offset* o = &ctu[a].comp[c];
      if (o->mode == VTMGPU_SAO_MODE_OFF) continue;
      if (o->mode == VTMGPU_SAO_MODE_NEW)
      {
        const int sc = c == 0 ? log2_scale_luma : log2_scale_chroma;
        int16_t coded[32];
        memcpy(coded, o->offset, sizeof(coded));
        memset(o->offset, 0, sizeof(o->offset));
        if (o->type == VTMGPU_SAO_BO)
          for (int i = 0; i < 4; i++) o->offset[(o->aux + i) & 31] = (int16_t)(coded[(o->aux + i) & 31] * (1 << sc));
        else
        {
          for (int i = 0; i < 5; i++) o->offset[i] = (int16_t)(coded[i] * (1 << sc));
          if (o->offset[2] != 0) return -1;
        }
      }
      else if (o->mode == VTMGPU_SAO_MODE_MERGE)
      {
        int src;
        if (o->type == VTMGPU_SAO_MERGE_LEFT)       { if (!ctu[a].merge_left_ok || (a % width_in_ctus) == 0) return -2; src = a - 1; }
        else if (o->type == VTMGPU_SAO_MERGE_ABOVE) { if (!ctu[a].merge_above_ok || a < width_in_ctus) return -2; src = a - width_in_ctus; }
        else return -3;
        *o = ctu[src].comp[c];
      }
      else return -4;
    }
    for (int c = 0; c < num_comps; c++)
      if (ctu[a].comp[c].mode != VTMGPU_SAO_MODE_OFF) mask |= 1 << c;
  }
  return mask;
}

/* SAOProcess (SampleAdaptiveOffset.cpp:618) on reconstructed params: reads a frozen copy, writes in place.
 * offsetBlock (:293-547) restated per sample: an EO sample is modified iff both neighbours lie inside the
 * picture and inside the own CTU or a neighbouring CTU whose availability bit is set. */
int vvco_sao(int16_t* const plane[3], const ptrdiff_t stride[3], int width, int height, int chroma_format,
             int bd_luma, int bd_chroma, int ctu_size, const vtmgpu_sao_params* p)
{
  int sx, sy;
  chroma_shifts(chroma_format, &sx, &sy);
  const int ncomp = chroma_format == 0 ? 1 : 3;
  const int wctu = (width + ctu_size - 1) / ctu_size;
  static const int nb[4][4] = { { -1, 0, 1, 0 }, { 0, -1, 0, 1 }, { -1, -1, 1, 1 }, { 1, -1, -1, 1 } };
  /* availability bit of the neighbour CTU at (cx,cy) in {-1,0,1}^2 */
  static const int bit[3][3] = { { VTMGPU_AVAIL_ABOVE_LEFT, VTMGPU_AVAIL_ABOVE, VTMGPU_AVAIL_ABOVE_RIGHT },
                                 { VTMGPU_AVAIL_LEFT, 0x100, VTMGPU_AVAIL_RIGHT },
                                 { VTMGPU_AVAIL_BELOW_LEFT, VTMGPU_AVAIL_BELOW, VTMGPU_AVAIL_BELOW_RIGHT } };
  int any = 0;
  for (int a = 0; a < p->num_ctus; a++)
    for (int c = 0; c < ncomp; c++) any |= p->ctu[a].comp[c].mode != VTMGPU_SAO_MODE_OFF;
  if (!any) return 0;
  for (int c = 0; c < ncomp; c++)
  {
    const int csx = c ? sx : 0, csy = c ? sy : 0;
    const int w = width >> csx, h = height >> csy, cw = ctu_size >> csx, ch = ctu_size >> csy;
    const int bd = c ? bd_chroma : bd_luma, maxv = (1 << bd) - 1;
    pel* src = (pel*)malloc(sizeof(pel) * (size_t)w * h);
    if (!src) return -1;
    for (int y = 0; y < h; y++) memcpy(src + (size_t)y * w, plane[c] + y * stride[c], sizeof(pel) * w);
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++)
      {
        const int ctx = x / cw, cty = y / ch;
        const vtmgpu_sao_ctu* cp = &p->ctu[cty * wctu + ctx];
        const vtmgpu_sao_offset* o = &cp->comp[c];
        if (o->mode == VTMGPU_SAO_MODE_OFF) continue;
        const int v = src[(size_t)y * w + x];
        if (o->type == VTMGPU_SAO_BO)
        {
          plane[c][y * stride[c] + x] = (pel)clip3(0, maxv, v + o->offset[v >> (bd - 5)]);
          continue;
        }
        const int* d = nb[o->type];
        int ok = 1, e = 0;
        if (p->vb)
        {
          /* isProcessDisabled (SampleAdaptiveOffset.h:96-116): classes that look across a vertical / horizontal boundary leave the
           * two samples next to it untouched (EO 0: vertical boundaries only, :319; EO 90: horizontal only, :362; diagonals: both) */
          if (o->type != VTMGPU_SAO_EO_90)
            for (int i = 0; i < p->vb->num_ver; i++) { const int b = p->vb->pos_x[i] >> csx; if (x == b || x == b - 1) ok = 0; }
          if (o->type != VTMGPU_SAO_EO_0)
            for (int i = 0; i < p->vb->num_hor; i++) { const int b = p->vb->pos_y[i] >> csy; if (y == b || y == b - 1) ok = 0; }
        }
        for (int k = 0; k < 2 && ok; k++)
        {
          const int nx = x + d[2 * k], ny = y + d[2 * k + 1];
          if (nx < 0 || ny < 0 || nx >= w || ny >= h) { ok = 0; break; }
          const int rx = nx / cw - ctx, ry = ny / ch - cty;
          if ((rx || ry) && !(cp->avail & bit[ry + 1][rx + 1])) { ok = 0; break; }
          e += sgn(v - src[(size_t)ny * w + nx]);
        }
        if (!ok) continue;
        plane[c][y * stride[c] + x] = (pel)clip3(0, maxv, v + o->offset[2 + e]);
      }
    free(src);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * ALF / CC-ALF
 * ---------------------------------------------------------------------------------------------- */

/* final per-class tables of one luma filter set: coeff[25][12], clip[25][12]
 * set < 16: fixed set (AdaptiveLoopFilter.cpp:792-807), else APS set-16 (:651-713) */
int vvco_alf_luma_set(const vtmgpu_alf_params* p, int set, int bd_luma, int16_t coeff[25][12], int16_t clip[25][12])
{
  const int clipv[4] = { 1 << bd_luma, 1 << (bd_luma - 3), 1 << (bd_luma - 5), 1 << (bd_luma - 7) };
  if (set < VTMGPU_ALF_FIXED_SETS)
  {
    for (int c = 0; c < 25; c++)
      for (int k = 0; k < 12; k++)
      {
        coeff[c][k] = vvc_alf_fix_coeff[vvc_alf_class_to_filt[set * 25 + c] * 12 + k];
        clip[c][k] = (int16_t)clipv[0];
      }
    return 0;
  }
  set -= VTMGPU_ALF_FIXED_SETS;
  if (set >= p->num_luma_aps || !p->luma_aps) return -1;
  const vtmgpu_alf_luma_aps* a = &p->luma_aps[set];
  for (int c = 0; c < 25; c++)
  {
    const int f = a->delta_idx[c];
    if (f < 0 || f >= a->num_filters) return -2;
    for (int k = 0; k < 12; k++)
    {
      coeff[c][k] = a->coeff[f][k];
      const int ci = a->nonlinear ? a->clip_idx[f][k] : 0;
      if (ci < 0 || ci > 3) return -3;
      clip[c][k] = (int16_t)clipv[ci];
    }
  }
  return 0;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* tightly packed copy of a plane with replicate-border access.  The window [xlo,xhi] x [ylo,yhi] is the picture, narrowed per
 * CTU to the sides that must not be read across (ALFProcess :452-477: copy of the CTU + extendBorderPel = nearest sample
 * inside); tl / br: raster-slice corner padding (padBorderPel, Buffer.h:571-607): left of tlx and above tly the sample of
 * column tlx in the same row is used, right of brx and below bry the one of column brx. */
typedef struct { const pel* s; int w, h, xlo, xhi, ylo, yhi, tl, tlx, tly, br, brx, bry; } cplane;
static inline int at(const cplane* c, int x, int y)
{
  if (c->tl && x < c->tlx && y < c->tly) x = c->tlx;
  if (c->br && x > c->brx && y > c->bry) x = c->brx;
  return c->s[(size_t)clampi(y, c->ylo, c->yhi) * c->w + clampi(x, c->xlo, c->xhi)];
}

/* narrows the access window of component plane S to CTU [x0,x1) x [y0,y1) (component samples) on the clipped sides */
static void alf_window(cplane* S, int x0, int y0, int x1, int y1, int clip)
{
  S->xlo = (clip & VTMGPU_ALF_CLIP_LEFT) ? x0 : 0;       S->xhi = (clip & VTMGPU_ALF_CLIP_RIGHT) ? x1 - 1 : S->w - 1;
  S->ylo = (clip & VTMGPU_ALF_CLIP_TOP) ? y0 : 0;        S->yhi = (clip & VTMGPU_ALF_CLIP_BOTTOM) ? y1 - 1 : S->h - 1;
  S->tl = (clip & VTMGPU_ALF_PAD_TL) != 0; S->tlx = x0; S->tly = y0;
  S->br = (clip & VTMGPU_ALF_PAD_BR) != 0; S->brx = x1 - 1; S->bry = y1 - 1;
}

/* deriveClassificationBlk for the 4x4 block at (bx,by) (AdaptiveLoopFilter.cpp:873-1082) */
static void alf_classify(const cplane* L, int bx, int by, int bd, int ctu, int* cls, int* tr)
{
  static const int th[16] = { 0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4 };
  static const int ttab[8] = { 0, 1, 0, 2, 2, 3, 1, 3 };
  const int vb = ctu - 4, yb = by & (ctu - 1);
  int sum[4] = { 0, 0, 0, 0 }; /* V H D0 D1 */
  for (int i = 0; i < 8; i += 2)
  {
    if (yb == vb - 4 && i == 6) continue;       /* block directly above the virtual boundary: 3 row pairs */
    if (yb == vb && i == 0) continue;           /* block directly below                                   */
    const int r = by - 2 + i;
    int rm1 = r - 1, rp2 = r + 2;
    if (r > 0 && (r & (ctu - 1)) == vb - 2) rp2 = r + 1;
    else if (r > 0 && (r & (ctu - 1)) == vb) rm1 = r;
    for (int j = 0; j < 8; j += 2)
    {
      const int c = bx - 2 + j;
      const int y0 = at(L, c, r) << 1, y1 = at(L, c + 1, r + 1) << 1;
      sum[0] += iabs(y0 - at(L, c, rm1) - at(L, c, r + 1)) + iabs(y1 - at(L, c + 1, r) - at(L, c + 1, rp2));
      sum[1] += iabs(y0 - at(L, c + 1, r) - at(L, c - 1, r)) + iabs(y1 - at(L, c + 2, r + 1) - at(L, c, r + 1));
      sum[2] += iabs(y0 - at(L, c - 1, rm1) - at(L, c + 1, r + 1)) + iabs(y1 - at(L, c, r) - at(L, c + 2, rp2));
      sum[3] += iabs(y0 - at(L, c - 1, r + 1) - at(L, c + 1, rm1)) + iabs(y1 - at(L, c, rp2) - at(L, c + 2, r));
    }
  }
  const int sumV = sum[0], sumH = sum[1], sumD0 = sum[2], sumD1 = sum[3];
  const int scale = (yb == vb - 4 || yb == vb) ? 96 : 64;
  const int act = clip3(0, 15, ((sumV + sumH) * scale) >> (bd + 4));
  int ci = th[act];
  int hv1, hv0, d1, d0, dirHV, dirD, hvd1, hvd0, mainDir, secDir;
  if (sumV > sumH) { hv1 = sumV; hv0 = sumH; dirHV = 1; } else { hv1 = sumH; hv0 = sumV; dirHV = 3; }
  if (sumD0 > sumD1) { d1 = sumD0; d0 = sumD1; dirD = 0; } else { d1 = sumD1; d0 = sumD0; dirD = 2; }
  if ((uint32_t)d1 * (uint32_t)hv0 > (uint32_t)hv1 * (uint32_t)d0) { hvd1 = d1; hvd0 = d0; mainDir = dirD; secDir = dirHV; }
  else { hvd1 = hv1; hvd0 = hv0; mainDir = dirHV; secDir = dirD; }
  int strength = 0;
  if (hvd1 > 2 * hvd0) strength = 1;
  if (hvd1 * 2 > 9 * hvd0) strength = 2;
  if (strength) ci += (((mainDir & 1) << 1) + strength) * 5;
  *cls = ci;
  *tr = ttab[mainDir * 2 + (secDir >> 1)];
}

static const int8_t tap7[12][2] = { { 3, 0 }, { 2, 1 }, { 2, 0 }, { 2, -1 }, { 1, 2 }, { 1, 1 }, { 1, 0 }, { 1, -1 }, { 1, -2 }, { 0, 3 }, { 0, 2 }, { 0, 1 } };
static const int8_t tap5[6][2]  = { { 2, 0 }, { 1, 1 }, { 1, 0 }, { 1, -1 }, { 0, 2 }, { 0, 1 } };
static const int8_t perm7[4][12] = { { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 }, { 9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6 },
                                     { 0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11 }, { 9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6 } };

/* one sample of filterBlk (AdaptiveLoopFilter.cpp:1215-1300): ntap 12 (7x7) or 6 (5x5) */
static int alf_sample(const cplane* S, int x, int y, int ntap, const int8_t (*tap)[2], const int16_t* coeff, const int16_t* clip,
                      const int8_t* perm, int vb_h, int vb_pos, int maxv)
{
  const int yv = y & (vb_h - 1);
  const int span = ntap == 12 ? 4 : 2;
  int lim = 3, near = 0;
  if (yv < vb_pos && yv >= vb_pos - span) { lim = vb_pos - 1 - yv; near = (yv == vb_pos - 1); }
  else if (yv >= vb_pos && yv <= vb_pos + span - 1) { lim = yv - vb_pos; near = (yv == vb_pos); }
  const int cur = at(S, x, y);
  int sum = 0;
  for (int k = 0; k < ntap; k++)
  {
    int dy = tap[k][0], dx = tap[k][1];
    if (dy > lim) dy = lim;
    const int c = coeff[perm ? perm[k] : k], cl = clip[perm ? perm[k] : k];
    sum += c * (clip3(-cl, cl, at(S, x + dx, y + dy) - cur) + clip3(-cl, cl, at(S, x - dx, y - dy) - cur));
  }
  sum = (sum + 64) >> (near ? 10 : 7);
  return clip3(0, maxv, sum + cur);
}

/* ALFProcess (AdaptiveLoopFilter.cpp:393-618) incl. the clip / pad path at slice and tile boundaries (p->ctu_clip); signalled
 * virtual boundaries inside a CTU are not covered */
int vvco_alf(int16_t* const plane[3], const ptrdiff_t stride[3], int width, int height, int chroma_format,
             int bd_luma, int bd_chroma, int ctu_size, const vtmgpu_alf_params* p)
{
  int sx, sy;
  chroma_shifts(chroma_format, &sx, &sy);
  const int ncomp = chroma_format == 0 ? 1 : 3;
  if (!p->enabled[0] && !p->enabled[1] && !p->enabled[2]) return 0;
  const int wctu = (width + ctu_size - 1) / ctu_size, hctu = (height + ctu_size - 1) / ctu_size;
  cplane S[3];
  pel* copy[3] = { 0, 0, 0 };
  for (int c = 0; c < ncomp; c++)
  {
    const int w = width >> (c ? sx : 0), h = height >> (c ? sy : 0);
    copy[c] = (pel*)malloc(sizeof(pel) * (size_t)w * h);
    if (!copy[c]) return -1;
    for (int y = 0; y < h; y++) memcpy(copy[c] + (size_t)y * w, plane[c] + y * stride[c], sizeof(pel) * w);
    S[c].s = copy[c]; S[c].w = w; S[c].h = h;
    alf_window(&S[c], 0, 0, w, h, 0);
  }
  int rc = 0;
  int16_t lc[25][12], lk[25][12];
  int cur_set = -1;
  const int clipc[4] = { 1 << bd_chroma, 1 << (bd_chroma - 3), 1 << (bd_chroma - 5), 1 << (bd_chroma - 7) };
  for (int cy = 0; cy < hctu && !rc; cy++)
    for (int cx = 0; cx < wctu && !rc; cx++)
    {
      const int a = cy * wctu + cx;
      const int x0 = cx * ctu_size, y0 = cy * ctu_size;
      const int x1 = x0 + ctu_size < width ? x0 + ctu_size : width, y1 = y0 + ctu_size < height ? y0 + ctu_size : height;
      const int clip0 = p->ctu_clip ? p->ctu_clip[a] : 0;
      /* virtual boundaries strictly inside the CTU split it into up to 2 x 2 separately padded parts; one that coincides with a
       * CTU edge clips that side (isCrossedByVirtualBoundaries :86-120, sub-block loop :452-477) */
      int xs[3] = { x0, x1, x1 }, ys[3] = { y0, y1, y1 }, nxs = 1, nys = 1, clipv = 0;
      if (p->vb)
      {
        for (int i = 0; i < p->vb->num_ver; i++)
        {
          const int b = p->vb->pos_x[i];
          if (b == x0) clipv |= VTMGPU_ALF_CLIP_LEFT; else if (b == x1) clipv |= VTMGPU_ALF_CLIP_RIGHT;
          else if (b > x0 && b < x1) { xs[1] = b; xs[2] = x1; nxs = 2; }
        }
        for (int i = 0; i < p->vb->num_hor; i++)
        {
          const int b = p->vb->pos_y[i];
          if (b == y0) clipv |= VTMGPU_ALF_CLIP_TOP; else if (b == y1) clipv |= VTMGPU_ALF_CLIP_BOTTOM;
          else if (b > y0 && b < y1) { ys[1] = b; ys[2] = y1; nys = 2; }
        }
      }
      for (int pj = 0; pj < nys && !rc; pj++)
      for (int pi = 0; pi < nxs && !rc; pi++)
      {
      const int px0 = xs[pi], px1 = xs[pi + 1], py0 = ys[pj], py1 = ys[pj + 1];
      int clip = (clip0 | clipv) & 15;
      if (pi > 0) clip |= VTMGPU_ALF_CLIP_LEFT;
      if (pi < nxs - 1) clip |= VTMGPU_ALF_CLIP_RIGHT;
      if (pj > 0) clip |= VTMGPU_ALF_CLIP_TOP;
      if (pj < nys - 1) clip |= VTMGPU_ALF_CLIP_BOTTOM;
      /* raster-slice corner padding applies to the part that holds the CTU's corner (:478-488); the reference derives the pad
       * flags only when the corresponding sides are not clipped, virtual boundaries included (:178-200) */
      if ((clip0 & VTMGPU_ALF_PAD_TL) && pi == 0 && pj == 0 && !(clipv & (VTMGPU_ALF_CLIP_TOP | VTMGPU_ALF_CLIP_LEFT))) clip |= VTMGPU_ALF_PAD_TL;
      if ((clip0 & VTMGPU_ALF_PAD_BR) && pi == nxs - 1 && pj == nys - 1 && !(clipv & (VTMGPU_ALF_CLIP_BOTTOM | VTMGPU_ALF_CLIP_RIGHT))) clip |= VTMGPU_ALF_PAD_BR;
      for (int c = 0; c < ncomp; c++) alf_window(&S[c], px0 >> (c ? sx : 0), py0 >> (c ? sy : 0), px1 >> (c ? sx : 0), py1 >> (c ? sy : 0), clip);
      if (p->ctu_enable[0][a])
      {
        const int set = p->ctu_filter_idx[a];
        if (set != cur_set) { rc = vvco_alf_luma_set(p, set, bd_luma, lc, lk); cur_set = set; if (rc) break; }
        for (int by = py0; by < py1; by += 4)
          for (int bx = px0; bx < px1; bx += 4)
          {
            int cls, tr;
            alf_classify(&S[0], bx, by, bd_luma, ctu_size, &cls, &tr);
            for (int y = by; y < by + 4; y++)
              for (int x = bx; x < bx + 4; x++)
                plane[0][y * stride[0] + x] = (pel)alf_sample(&S[0], x, y, 12, tap7, lc[cls], lk[cls], perm7[tr], ctu_size, ctu_size - 4, (1 << bd_luma) - 1);
          }
      }
      for (int c = 1; c < ncomp; c++)
      {
        const int vbh = ctu_size >> sy, vbp = vbh - 2, maxv = (1 << bd_chroma) - 1;
        if (p->ctu_enable[c][a])
        {
          const int alt = p->ctu_alt[c - 1][a];
          if (!p->chroma_aps || alt >= p->chroma_aps->num_alts) { rc = -4; break; }
          int16_t cc[6], ck[6];
          for (int k = 0; k < 6; k++)
          {
            cc[k] = p->chroma_aps->coeff[alt][k];
            ck[k] = (int16_t)clipc[p->chroma_aps->nonlinear ? p->chroma_aps->clip_idx[alt][k] : 0];
          }
          for (int y = py0 >> sy; y < (py1 >> sy); y++)
            for (int x = px0 >> sx; x < (px1 >> sx); x++)
              plane[c][y * stride[c] + x] = (pel)alf_sample(&S[c], x, y, 6, tap5, cc, ck, 0, vbh, vbp, maxv);
        }
        if (p->ccalf_enabled[c - 1] && p->ccalf_idc[c - 1][a] != 0)
        {
          const int16_t* f = p->ccalf_coeff[c - 1][p->ccalf_idc[c - 1][a] - 1];
          const int half = (1 << bd_chroma) >> 1;
          for (int y = py0 >> sy; y < (py1 >> sy); y++)
            for (int x = px0 >> sx; x < (px1 >> sx); x++)
            {
              const int lx = x << sx, ly = y << sy, pos = ly & (ctu_size - 1), vb = ctu_size - 4;
              int o1 = 1, o2 = -1, o3 = 2;
              if (pos == vb - 2 || pos == vb + 1) o3 = 1;
              else if (pos == vb - 1 || pos == vb) o1 = o2 = o3 = 0;
              const int cur = at(&S[0], lx, ly);
              int sum = f[0] * (at(&S[0], lx, ly + o2) - cur) + f[1] * (at(&S[0], lx - 1, ly) - cur) + f[2] * (at(&S[0], lx + 1, ly) - cur)
                      + f[3] * (at(&S[0], lx - 1, ly + o1) - cur) + f[4] * (at(&S[0], lx, ly + o1) - cur) + f[5] * (at(&S[0], lx + 1, ly + o1) - cur)
                      + f[6] * (at(&S[0], lx, ly + o3) - cur);
              sum = (sum + 64) >> 7;
              sum = clip3(0, maxv, sum + half) - half;
              pel* d = &plane[c][y * stride[c] + x];
              *d = (pel)clip3(0, maxv, sum + *d);
            }
        }
      }
      }
    }
  for (int c = 0; c < 3; c++) free(copy[c]);
  return rc;
}

/* DecLib::executeLoopFilters order (DecLib.cpp:579-595): DBF -> SAO -> ALF; NULL params = stage off */
int vvco_filter_picture(int16_t* const plane[3], const ptrdiff_t stride[3], int width, int height, int chroma_format,
                        int bd_luma, int bd_chroma, int ctu_size, const vtmgpu_deblock_params* dbf,
                        const vtmgpu_sao_params* sao, const vtmgpu_alf_params* alf)
{
  int rc = 0;
  if (dbf) rc = vvco_deblock(plane, stride, width, height, chroma_format, bd_luma, bd_chroma, dbf);
  if (!rc && sao) rc = vvco_sao(plane, stride, width, height, chroma_format, bd_luma, bd_chroma, ctu_size, sao);
  if (!rc && alf) rc = vvco_alf(plane, stride, width, height, chroma_format, bd_luma, bd_chroma, ctu_size, alf);
  return rc;
}
