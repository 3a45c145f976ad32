/* vtmgpu_derive.h -- the deblocking DERIVATION of one 4x4 luma unit and one edge direction from the flattened block structure
 * (vtmgpu_deblock_units, include/vtmgpu.h): is there an edge, which boundary strength, which filter lengths, tc / beta -> the packed
 * luma and chroma segment records the deblocking kernel consumes.
 *
 *   LoopFilter::xDeblockCU                             LoopFilter.cpp:261-408   (edge maps, the CU walk)
 *   xSetLoopfilterParam / xSetEdgefilterMultiple       :627-672
 *   xSetMaxFilterLengthPQFromTransformSizes            :454-548
 *   xSetMaxFilterLengthPQForCodingSubBlocks            :550-625
 *   xGetBoundaryStrengthSingle                         :674-812
 *   xEdgeFilterLuma / xEdgeFilterChroma up to the sample reads   :844-977, :1087-1249
 *
 * The reference walks the CUs of a CTU and keeps per-CTU arrays that later CUs read; everything it stores for a unit depends only on
 * the coding unit / transform unit that cover the unit, their neighbours across the edge and the motion field, so here ONE FUNCTION
 * evaluates a unit from the flattened tables -- the kernel k_dbf_derive runs it for every unit and direction (vvc_b200/csrc), and the
 * test build of the shim runs the very same source on the host to check it against the CU walk on every picture.  Plain C++11, no CUDA
 * or reference types.
 */
#pragma once

#include <stdint.h>

#include "vtmgpu.h"

#ifdef __CUDACC__
#define VTMGPU_HD __host__ __device__ __forceinline__
#else
#define VTMGPU_HD inline
#endif

namespace vtmgpu_derive
{

struct Ctx
{
  const vtmgpu_dbf_cu*    cus;
  const vtmgpu_dbf_tu*    tus;
  const vtmgpu_dbf_slice* slices;
  const uint32_t* tuL;
  const uint32_t* tuC;
  const unsigned char* motion;
  int miBytes, miPitch, offMv0, offMv1, offRef0, offRef1;
  int w4, h4;                 /* picture size in 4x4 units */
  int sx, sy, chroma;         /* chroma shifts; chroma = 0 for 4:0:0 */
  int bdL, bdC, ctuLog2;
  int flags;                  /* VTMGPU_UNITS_* */
  int ladf;                   /* the luma records carry QPs (vtmgpu_ladf) */
  int nvb[2], vb[2][3];       /* signalled virtual boundaries: [0] x positions, [1] y positions */
  const uint16_t* tcTable;    /* 66 entries (LoopFilter.cpp:66) */
  const uint8_t*  betaTable;  /* 64 entries (:70) */
};

enum { VER = 0, HOR = 1 };

VTMGPU_HD int clip3i(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
VTMGPU_HD int absi(int v) { return v < 0 ? -v : v; }

VTMGPU_HD bool onVb(const Ctx& D, int dir, int pos)
{
  for (int i = 0; i < D.nvb[dir]; i++) if (D.vb[dir][i] == pos) return true;
  return false;
}

/* "m_transformEdge" of a component at a unit: a transform block of that component starts here and the edge may be filtered
 * (xSetMaxFilterLengthPQFromTransformSizes: at the CU border the left / top flag decides, inside the CU the internal flag) */
VTMGPU_HD bool tuEdgeLuma(const Ctx& D, int x, int y, int dir)
{
  const vtmgpu_dbf_tu& t = D.tus[D.tuL[y * D.w4 + x]];
  if (!t.w) return false;
  const vtmgpu_dbf_cu& c = D.cus[t.cu];
  const int pos = dir == VER ? 4 * x : 4 * y, ts = dir == VER ? t.x : t.y, cs = dir == VER ? c.x : c.y;
  if (pos != ts) return false;
  return (c.flags & (ts == cs ? (dir == VER ? VTMGPU_CU_EN_LEFT : VTMGPU_CU_EN_TOP) : VTMGPU_CU_EN_INT)) != 0;
}

/* the luma-channel TU on the P side of an edge: the reference looks up the sample next to the edge (posQ - 1), which in a unit made of
 * ISP sub-partitions 1 or 2 samples wide (high) is another TU than the one that holds the unit's first sample */
VTMGPU_HD const vtmgpu_dbf_tu& tuLumaP(const Ctx& D, int unit, int dir)
{
  const vtmgpu_dbf_tu& t = D.tus[D.tuL[unit]];
  const int n = dir == VER ? t.w : t.h;
  return (n && n < 4) ? D.tus[t.tail] : t;
}

VTMGPU_HD bool usable(const Ctx& D, const vtmgpu_dbf_cu& q, const vtmgpu_dbf_cu& p)
{
  /* isAvailableLeft / isAvailableAbove (LoopFilter.cpp:85-93) */
  return ((D.flags & VTMGPU_UNITS_ACROSS_SLICES) || D.slices[q.slice].independent_idx == D.slices[p.slice].independent_idx) &&
         ((D.flags & VTMGPU_UNITS_ACROSS_TILES) || q.tile == p.tile);
}

struct Mv2 { int h, v; };
VTMGPU_HD Mv2 loadMv(const unsigned char* mi, int off)
{
  const int* p = reinterpret_cast<const int*>(mi + off);
  Mv2 m; m.h = p[0]; m.v = p[1];
  return m;
}
VTMGPU_HD bool farMv(const Mv2& a, const Mv2& b) { return absi(a.h - b.h) >= 8 || absi(a.v - b.v) >= 8; }     /* half a luma sample in 1/16 units */

/* reference picture of a motion entry: -1 none, -2 the current picture (IBC), else the slice's picture id */
VTMGPU_HD int refOf(const Ctx& D, const vtmgpu_dbf_cu& cu, int list, int refIdx)
{
  if (cu.flags & VTMGPU_CU_IBC) return list == 0 ? -2 : -1;
  if (refIdx < 0 || refIdx > 15) return -1;
  return D.slices[cu.slice].ref_pic[list][refIdx];
}

/* xGetBoundaryStrengthSingle: packed Y | Cb << 2 | Cr << 4.  q / p = the CUs either side (channel of q), tq / tp = the TUs there,
 * code = 0 no transform edge / 1 transform edge / 3 transform edge that is also a PU or sub-block edge, uq / up = unit indices */
VTMGPU_HD unsigned strength(const Ctx& D, const vtmgpu_dbf_cu& q, const vtmgpu_dbf_cu& p, const vtmgpu_dbf_tu& tq, const vtmgpu_dbf_tu& tp, int code, int uq, int up)
{
  const bool intraP = (p.flags & VTMGPU_CU_INTRA) != 0, intraQ = (q.flags & VTMGPU_CU_INTRA) != 0;
  if (intraP || intraQ)
  {
    const unsigned y = (intraP && (p.flags & VTMGPU_CU_BDPCM) && intraQ && (q.flags & VTMGPU_CU_BDPCM)) ? 0 : 2;
    const unsigned c = (intraP && (p.flags & VTMGPU_CU_BDPCM_C) && intraQ && (q.flags & VTMGPU_CU_BDPCM_C)) ? 0 : 2;
    return y | (c << 2) | (c << 4);
  }
  const bool ciip = ((p.flags | q.flags) & VTMGPU_CU_CIIP) != 0;
  if (code && ciip) return 2 | (2 << 2) | (2 << 4);
  unsigned bs = 0;
  if (code)
  {
    const unsigned cb = tq.cbf | tp.cbf;
    if (cb & VTMGPU_TU_CBF_Y) bs |= 1;
    if (cb & (VTMGPU_TU_CBF_CB | VTMGPU_TU_JOINT)) bs |= 1 << 2;
    if (cb & (VTMGPU_TU_CBF_CR | VTMGPU_TU_JOINT)) bs |= 1 << 4;
  }
  if ((bs & 3) == 1) return bs;
  if (ciip) return 1;
  if (!(q.flags & VTMGPU_CU_HAS_LUMA)) return bs;
  if (code != 0 && code != 3) return bs;          /* pure transform edge: no motion test */
  if (!D.motion) return bs + 1;                   /* (not reached: pictures with inter CUs carry the motion field) */
  const unsigned char* miQ = D.motion + (size_t)((uq / D.w4) * D.miPitch + uq % D.w4) * D.miBytes;
  const unsigned char* miP = D.motion + (size_t)((up / D.w4) * D.miPitch + up % D.w4) * D.miBytes;
  const int rq0 = *reinterpret_cast<const int16_t*>(miQ + D.offRef0), rq1 = *reinterpret_cast<const int16_t*>(miQ + D.offRef1);
  const int rp0 = *reinterpret_cast<const int16_t*>(miP + D.offRef0), rp1 = *reinterpret_cast<const int16_t*>(miP + D.offRef1);
  if (D.slices[q.slice].inter_b || D.slices[p.slice].inter_b)
  {
    const int refP0 = refOf(D, p, 0, rp0), refP1 = refOf(D, p, 1, rp1), refQ0 = refOf(D, q, 0, rq0), refQ1 = refOf(D, q, 1, rq1);
    Mv2 z; z.h = z.v = 0;
    const Mv2 p0 = rp0 >= 0 ? loadMv(miP, D.offMv0) : z, p1 = rp1 >= 0 ? loadMv(miP, D.offMv1) : z;
    const Mv2 q0 = rq0 >= 0 ? loadMv(miQ, D.offMv0) : z, q1 = rq1 >= 0 ? loadMv(miQ, D.offMv1) : z;
    unsigned mvBs = 1;                         /* different reference pictures */
    if ((refP0 == refQ0 && refP1 == refQ1) || (refP0 == refQ1 && refP1 == refQ0))
    {
      if (refP0 != refP1) mvBs = refP0 == refQ0 ? (farMv(q0, p0) || farMv(q1, p1)) : (farMv(q1, p0) || farMv(q0, p1));
      else                mvBs = (farMv(q0, p0) || farMv(q1, p1)) && (farMv(q1, p0) || farMv(q0, p1));
    }
    return bs + mvBs;
  }
  /* P slices */
  if (refOf(D, p, 0, rp0) != refOf(D, q, 0, rq0)) return bs + 1;
  return farMv(loadMv(miQ, D.offMv0), loadMv(miP, D.offMv0)) ? bs + 1 : bs;
}

/* One unit (x, y in 4x4 units), one direction.  lumaRec: the record of the unit; chromaRec / chromaSlot: the record of the chroma
 * grid position the unit sits on (chromaSlot = false: the unit is not on the chroma edge grid and owns no record). */
VTMGPU_HD void deriveUnit(const Ctx& D, int x, int y, int dir, uint32_t& lumaRec, uint64_t& chromaRec, bool& chromaSlot)
{
  lumaRec = 0; chromaRec = 0;
  const int px = 4 * x, py = 4 * y, pos = dir == VER ? px : py;
  const int ctuMask4 = (1 << (D.ctuLog2 - 2)) - 1;
  const int unitC = dir == VER ? 4 >> D.sx : 4 >> D.sy;                 /* chroma samples per unit across the edge */
  chromaSlot = D.chroma && (((dir == VER ? x : y) & ctuMask4) % (8 / unitC)) == 0;
  if ((dir == VER ? x : y) == 0) return;
  const int u = y * D.w4 + x, up = dir == VER ? u - 1 : u - D.w4;
  const bool vbHit = onVb(D, dir, pos);

  /* ---- luma channel: the CU that holds the luma block here ------------------------------------------------------------------ */
  const vtmgpu_dbf_tu& tq = D.tus[D.tuL[u]];
  const vtmgpu_dbf_cu& cq = D.cus[tq.cu];
  const vtmgpu_dbf_tu& tp = tuLumaP(D, up, dir);
  const vtmgpu_dbf_cu& cp = D.cus[tp.cu];
  const int cuStart = dir == VER ? cq.x : cq.y;
  const bool border = pos == cuStart;
  const bool tuE = tq.w != 0 && pos == (dir == VER ? tq.x : tq.y);
  const bool subE = (cq.flags & VTMGPU_CU_MVSUB) && !border && ((pos - cuStart) & 7) == 0;
  const bool en = (cq.flags & (border ? (dir == VER ? VTMGPU_CU_EN_LEFT : VTMGPU_CU_EN_TOP) : VTMGPU_CU_EN_INT)) != 0;
  unsigned bs = 0;                                                        /* packed strengths of the luma-channel CU */
  if ((tuE || border || subE) && en && !vbHit)
  {
    const int code = tuE ? ((border || subE) ? 3 : 1) : 0;
    bs = strength(D, cq, cp, tq, tp, code, u, up);
  }
  if (bs & 3)
  {
    /* xEdgeFilterLuma up to the sample reads */
    if (!usable(D, cq, cp)) bs = 0;          /* also suppresses the chroma filtering of this unit (LoopFilter.cpp:918-933) */
    else
    {
      int lenP = 0, lenQ = 0;
      const bool te0 = tuEdgeLuma(D, x, y, dir);
      if (te0)
      {
        const int sizeQ = dir == VER ? tq.w : tq.h, sizeP = dir == VER ? tp.w : tp.h;
        const bool small = sizeP <= 4 || sizeQ <= 4;
        lenQ = small ? 1 : (sizeQ >= 32 ? 7 : 3);
        lenP = small ? 1 : (sizeP >= 32 ? 7 : 3);
      }
      if ((cq.flags & VTMGPU_CU_MVSUB) && ((pos - cuStart) & 7) == 0)
      {
        /* sub-block edges of affine / SbTMVP CUs: lengths limited by the distance to the next transform edge */
        const int d = pos - cuStart, across = 4 * (dir == VER ? cq.w4 : cq.h4);
        const int dx = dir == VER ? 1 : 0, dy = dir == VER ? 0 : 1;
        if (te0)
        {
          lenQ = lenQ < 5 ? lenQ : 5;
          if (d > 0) lenP = lenP < 5 ? lenP : 5;
        }
        else if (d > 0 && (tuEdgeLuma(D, x - dx, y - dy, dir) || d + 4 >= across || tuEdgeLuma(D, x + dx, y + dy, dir))) lenQ = lenP = 1;
        else if (d > 0 && (tuEdgeLuma(D, x - 2 * dx, y - 2 * dy, dir) || d + 8 >= across || tuEdgeLuma(D, x + 2 * dx, y + 2 * dy, dir))) lenQ = lenP = 2;
        else lenQ = lenP = 3;
      }
      const vtmgpu_dbf_slice& sl = D.slices[cq.slice];
      const int qp = (cp.qp + cq.qp + 1) >> 1, bsY = (int)(bs & 3);
      if (lenP > 5 && (cp.flags & VTMGPU_CU_AFFINE)) lenP = 5;
      const bool ctuRow = dir == HOR && (py & ((1 << D.ctuLog2) - 1)) == 0;
      unsigned tc, beta;
      if (D.ladf)
      {
        tc   = (unsigned)(VTMGPU_DBF_LADF_BIAS + qp + 2 * (bsY - 1) + sl.tc_offset);
        beta = (unsigned)(VTMGPU_DBF_LADF_BIAS + qp + sl.beta_offset);
      }
      else
      {
        const int iTc = clip3i(0, 63 + 2, qp + 2 * (bsY - 1) + sl.tc_offset), iB = clip3i(0, 63, qp + sl.beta_offset);
        tc   = D.bdL < 10 ? (unsigned)(D.tcTable[iTc] + 2) >> (10 - D.bdL) : (unsigned)D.tcTable[iTc] << (D.bdL - 10);
        beta = (unsigned)D.betaTable[iB] << (D.bdL - 8);
      }
      uint32_t rec = tc | (beta << VTMGPU_DBF_L_BETA_SHIFT) | ((uint32_t)lenP << VTMGPU_DBF_L_LENP_SHIFT) | ((uint32_t)lenQ << VTMGPU_DBF_L_LENQ_SHIFT);
      if (D.flags & VTMGPU_UNITS_PLT)
      {
        if (cp.flags & VTMGPU_CU_PLT) rec |= VTMGPU_DBF_L_PNOFILT;
        if (cq.flags & VTMGPU_CU_PLT) rec |= VTMGPU_DBF_L_QNOFILT;
      }
      if (ctuRow) rec |= VTMGPU_DBF_L_CTUROW;
      lumaRec = tc ? rec : 0;
    }
  }
  if (!chromaSlot) return;

  /* ---- chroma channel: the CU that holds the chroma block here ---------------------------------------------------------------- */
  const vtmgpu_dbf_tu& tqc = D.tus[D.tuC[u]];
  const vtmgpu_dbf_cu& cqc = D.cus[tqc.cu];
  if (!(cqc.flags & VTMGPU_CU_HAS_CHROMA)) return;
  const vtmgpu_dbf_tu& tpc = D.tus[D.tuC[up]];
  const vtmgpu_dbf_cu& cpc = D.cus[tpc.cu];                           /* xEdgeFilterChroma's P side: always the CU that holds the chroma there */
  const int cStart = dir == VER ? cqc.x : cqc.y;
  const bool cBorder = pos == cStart;
  unsigned bsC;                                                         /* packed strengths as the chroma-emitting CU computed them */
  const vtmgpu_dbf_tu* tuQ;                                             /* the TUs whose QPs the chroma edge takes (LoopFilter.cpp:1196-1217) */
  if (cqc.flags & VTMGPU_CU_HAS_LUMA)
  {
    /* single tree: the strengths of the luma-channel evaluation above (the same CU); with ISP only the CU border carries chroma edges */
    if ((cqc.flags & VTMGPU_CU_ISP) && !cBorder) return;
    bsC = bs;
    tuQ = &tq;
  }
  else
  {
    /* chroma-only CU (dual tree, local dual tree): the walk marks its border only, as a transform + PU edge */
    if (!cBorder || vbHit || !(cqc.flags & (dir == VER ? VTMGPU_CU_EN_LEFT : VTMGPU_CU_EN_TOP))) return;
    bsC = strength(D, cqc, cpc, tqc, tpc, 3, u, up);
    tuQ = &tqc;
  }
  const unsigned b2[2] = { (bsC >> 2) & 3, (bsC >> 4) & 3 };
  if (!b2[0] && !b2[1]) return;
  const vtmgpu_dbf_tu* tuP = (cpc.flags & VTMGPU_CU_HAS_LUMA) ? &tp : &tpc;
  /* filter lengths of the chroma components: set where a chroma transform block starts (chroma channel on both sides) */
  bool large = false;
  {
    const int cpos = dir == VER ? (int)tqc.cx << D.sx : (int)tqc.cy << D.sy;
    if (tqc.cw && cpos == pos)
    {
      const bool atBorder = cpos == cStart;
      if (cqc.flags & (atBorder ? (dir == VER ? VTMGPU_CU_EN_LEFT : VTMGPU_CU_EN_TOP) : VTMGPU_CU_EN_INT))
      {
        const int sizeQ = dir == VER ? tqc.cw : tqc.ch, sizeP = dir == VER ? tpc.cw : tpc.ch;
        large = sizeQ >= 8 && sizeP >= 8;
      }
    }
  }
  const bool ctb = dir == HOR && (py & ((1 << D.ctuLog2) - 1)) == 0;
  const vtmgpu_dbf_slice& sl = D.slices[cqc.slice];
  uint64_t rec = 0;
  for (int c = 0; c < 2; c++)
  {
    if (!(b2[c] == 2 || (large && b2[c] == 1))) continue;
    const int qp = ((c ? tuQ->qp_cr : tuQ->qp_cb) + (c ? tuP->qp_cr : tuP->qp_cb) + 1) >> 1;
    const int iTc = clip3i(0, 63 + 2, qp + 2 * ((int)b2[c] - 1) + sl.tc_offset);
    const uint64_t tc = D.bdC < 10 ? (uint64_t)((D.tcTable[iTc] + 2) >> (10 - D.bdC)) : (uint64_t)D.tcTable[iTc] << (D.bdC - 10);
    uint64_t beta = 0;
    if (large) beta = (uint64_t)D.betaTable[clip3i(0, 63, qp + sl.beta_offset)] << (D.bdC - 8);
    rec |= tc << (c ? VTMGPU_DBF_C_TCCR_SHIFT : 0);
    rec |= beta << (c ? VTMGPU_DBF_C_BETACR_SHIFT : VTMGPU_DBF_C_BETACB_SHIFT);
  }
  if (!(rec & 0x3fffff)) return;             /* both tc zero: nothing to do */
  if (large) rec |= VTMGPU_DBF_C_LARGE;
  if (ctb)   rec |= VTMGPU_DBF_C_CTB;
  if (D.flags & VTMGPU_UNITS_PLT)
  {
    if (cpc.flags & VTMGPU_CU_PLT) rec |= VTMGPU_DBF_C_PNOFILT;
    if (cqc.flags & VTMGPU_CU_PLT) rec |= VTMGPU_DBF_C_QNOFILT;
  }
  chromaRec = rec;
}

}   // namespace vtmgpu_derive
