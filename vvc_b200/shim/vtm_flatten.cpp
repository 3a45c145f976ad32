// vtm_flatten.cpp -- see vtm_flatten.h.  Host-side derivation of the per-picture side information.
//
// The deblocking part follows LoopFilter::xDeblockCU and its helpers (LoopFilter.cpp:261-812) decision by
// decision, but is organised around EMITTING packed segment records (vtmgpu.h) for the device kernels instead of
// calling the sample filters; the luma/chroma edge walks follow xEdgeFilterLuma :896-977 and
// xEdgeFilterChroma :1114-1249 up to (not including) the sample reads.
#include "vtm_flatten.h"

#include <algorithm>
#include <exception>
#include <atomic>
#include <thread>
#include <vector>

#include <algorithm>
#include <cstring>

#include "vtmgpu_derive.h"

#include "CodingStructure.h"
#include "Picture.h"
#include "Quant.h"
#include "Slice.h"
#include "UnitTools.h"
#include "AlfParameters.h"

namespace vtmshim
{

vtmgpu_virtual_boundaries flattenVirtualBoundaries(const CodingStructure& cs)
{
  vtmgpu_virtual_boundaries v{};
  const PicHeader* ph = cs.picHeader;
  if (!ph->getLoopFilterAcrossVirtualBoundariesDisabledFlag()) return v;
  v.num_ver = (int)ph->getNumVerVirtualBoundaries();
  v.num_hor = (int)ph->getNumHorVirtualBoundaries();
  CHECK(v.num_ver > 3 || v.num_hor > 3, "vtmgpu shim: more than 3 virtual boundaries per direction");
  for (int i = 0; i < v.num_ver; i++) v.pos_x[i] = (int)ph->getVirtualBoundariesPosX(i);
  for (int i = 0; i < v.num_hor; i++) v.pos_y[i] = (int)ph->getVirtualBoundariesPosY(i);
  return v;
}

vtmgpu_deblock_params FlatDeblock::view() const
{
  vtmgpu_deblock_params p;
  for (int d = 0; d < 2; d++)
  {
    p.luma[d]   = luma[d].data();
    p.chroma[d] = chroma[d].empty() ? nullptr : chroma[d].data();
  }
  p.ladf = hasLadf ? &ladf : nullptr;
  return p;
}

vtmgpu_deblock_sparse FlatDeblock::sparseView() const
{
  vtmgpu_deblock_sparse p;
  for (int d = 0; d < 2; d++)
  {
    p.luma[d]   = lists.luma[d].data();   p.luma_count[d]   = (uint32_t)lists.luma[d].size();
    p.chroma[d] = lists.chroma[d].data(); p.chroma_count[d] = (uint32_t)lists.chroma[d].size();
  }
  p.ladf = hasLadf ? &ladf : nullptr;
  return p;
}

bool FlatDeblock::listsMatchDense() const
{
  for (int d = 0; d < 2; d++)
  {
    size_t nl = 0, nc = 0;
    for (uint32_t r : luma[d]) nl += r != 0;
    for (uint64_t r : chroma[d]) nc += r != 0;
    if (nl != lists.luma[d].size() || nc != lists.chroma[d].size()) return false;
    for (const auto& e : lists.luma[d]) if (e.index >= luma[d].size() || luma[d][e.index] != e.rec) return false;
    for (const auto& e : lists.chroma[d]) if (e.index >= chroma[d].size() || chroma[d][e.index] != e.rec) return false;
  }
  return true;
}

vtmgpu_sao_params FlatSao::view() const
{
  vtmgpu_sao_params p;
  p.ctu      = ctu.data();
  p.num_ctus = (int)ctu.size();
  p.vb       = (vb.num_ver || vb.num_hor) ? &vb : nullptr;
  return p;
}

const vtmgpu_alf_params* FlatAlf::view()
{
  p.luma_aps   = lumaAps.empty() ? nullptr : lumaAps.data();
  p.chroma_aps = hasChromaAps ? &chromaAps : nullptr;
  for (int c = 0; c < 3; c++) p.ctu_enable[c] = ctuEnable[c].data();
  for (int c = 0; c < 2; c++) { p.ctu_alt[c] = ctuAlt[c].data(); p.ccalf_idc[c] = ccIdc[c].data(); }
  p.ctu_filter_idx = filterIdx.data();
  p.ctu_clip = ctuClip.empty() ? nullptr : ctuClip.data();
  p.vb = (vb.num_ver || vb.num_hor) ? &vb : nullptr;
  return &p;
}

const vtmgpu_alf_params* FlatAlf::slicesView()
{
  sliceViews.assign(1, *view());
  for (SliceSet& s : more)
  {
    s.p.luma_aps   = s.lumaAps.empty() ? nullptr : s.lumaAps.data();
    s.p.chroma_aps = s.hasChromaAps ? &s.chromaAps : nullptr;
    s.p.num_ctus   = p.num_ctus;
    sliceViews.push_back(s.p);
  }
  return sliceViews.data();
}

// ---------------------------------------------------------------------------------------------------------
// deblocking derivation
// ---------------------------------------------------------------------------------------------------------
namespace
{

// tc / beta tables of the standard (LoopFilter.cpp:66-74)
const uint16_t kTc[66] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,3,4,4,4,4,5,5,5,5,7,7,8,9,10,10,11,13,14,15,17,19,21,24,25,29,33,36,
                           41,45,51,57,64,71,80,89,100,112,125,141,157,177,198,222,250,280,314,352,395 };
const uint8_t  kBeta[64] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,
                             42,44,46,48,50,52,54,56,58,60,62,64,66,68,70,72,74,76,78,80,82,84,86,88 };

constexpr int kU = 32;   // 4x4 luma units per CTU side at CTU 128

enum { VER = 0, HOR = 1 };

// what the reference keeps per CTU and direction in m_aapucBS / m_aapbEdgeFilter / m_maxFilterLength{P,Q} /
// m_transformEdge; here everything is indexed in units of 4 luma samples
struct CtuState
{
  uint8_t code[kU * kU];   // before the bS step: 0 none / 1 TU edge / 3 TU+PU edge; afterwards packed bS Y|Cb<<2|Cr<<4
  bool    on[kU * kU];
  uint8_t lenP[3][kU][kU], lenQ[3][kU][kU];   // [comp][x unit][y unit]
  bool    tuEdge[3][kU][kU];
  void clear() { memset(this, 0, sizeof(*this)); }
};

class Deriver
{
public:
  Deriver(CodingStructure& cs, FlatDeblock& out) : m_cs(cs), m_pcv(*cs.pcv), m_out(out)
  {
    m_sx = getComponentScaleX(COMPONENT_Cb, m_pcv.chrFormat);
    m_sy = getComponentScaleY(COMPONENT_Cb, m_pcv.chrFormat);
  }
  void run();

private:
  CodingStructure&     m_cs;
  const PreCalcValues& m_pcv;
  FlatDeblock&         m_out;
  FlatDeblock::Lists*  m_lists = nullptr;      // this thread's share of the record lists
  vtmgpu_virtual_boundaries m_vb{};
  mutable const TransformUnit *m_tuQ = nullptr, *m_tuP = nullptr;    // memo of strength()
  // xDeriveEdgefilterParam (LoopFilter.cpp:435-452): an edge that lies on a signalled virtual boundary is not filtered
  bool onVb(int dir, const Area& a) const
  {
    if (dir == VER) { for (int i = 0; i < m_vb.num_ver; i++) if (m_vb.pos_x[i] == (int)a.x) return true; }
    else            { for (int i = 0; i < m_vb.num_hor; i++) if (m_vb.pos_y[i] == (int)a.y) return true; }
    return false;
  }
  CtuState             m_st;
  const Slice*         m_ctuSlice = nullptr;   // slice of the first CU of the CTU in flight (what the reference leaves in cs.slice)
  int  m_sx, m_sy, m_ctuX = 0, m_ctuY = 0, m_dir = VER;
  void work(std::atomic<int>& next);
  bool m_left = false, m_top = false, m_internal = false;

  int uidx(int x, int y) const { return ((y & (int)m_pcv.maxCUHeightMask) >> 2) * kU + ((x & (int)m_pcv.maxCUWidthMask) >> 2); }
  // unit coordinates inside the CTU of a position given in component samples
  int ux(int comp, int xComp) const { return ((comp ? (xComp << m_sx) : xComp) - m_ctuX) >> 2; }
  int uy(int comp, int yComp) const { return ((comp ? (yComp << m_sy) : yComp) - m_ctuY) >> 2; }

  void ctuPass(const UnitArea& ctuArea);
  void deriveCU(CodingUnit& cu);
  void edgeEnables(const CodingUnit& cu);
  void mark(int dir, const Area& a, bool value, bool puInternal);
  void lengthsFromTU(const CodingUnit& cu, const TransformUnit& tu);
  void lengthsFromSubBlocks(const PredictionUnit& pu, bool mvSub, int sub, const Area& ap);
  unsigned strength(const CodingUnit& cu, const Position& p) const;
  void emitLuma(const CodingUnit& cu, int edge);
  void emitChroma(const CodingUnit& cu, int edge);
  void putLuma(size_t idx, uint32_t rec)
  {
    uint32_t& slot = m_out.luma[m_dir][idx];
    if (slot) m_lists->twice = true;
    slot = rec;
    if (rec) m_lists->luma[m_dir].push_back(vtmgpu_dbf_luma_entry{ (uint32_t)idx, rec });
  }
  void putChroma(size_t idx, uint64_t rec)
  {
    uint64_t& slot = m_out.chroma[m_dir][idx];
    if (slot) m_lists->twice = true;
    slot = rec;
    if (rec) m_lists->chroma[m_dir].push_back(vtmgpu_dbf_chroma_entry{ rec, (uint32_t)idx, 0u });
  }
};

bool usable(const CodingUnit& q, const CodingUnit& p, const PPS& pps)
{
  // isAvailableLeft / isAvailableAbove (LoopFilter.cpp:85-93)
  return (pps.getLoopFilterAcrossSlicesEnabledFlag() || CU::isSameSlice(q, p)) &&
         (pps.getLoopFilterAcrossTilesEnabledFlag() || CU::isSameTile(q, p));
}

void Deriver::run()
{
  const int W = m_pcv.lumaWidth, H = m_pcv.lumaHeight;
  CHECK(m_pcv.maxCUWidth > 128 || m_pcv.minCUWidth != 4 || m_pcv.minCUHeight != 4, "vtmgpu shim: unsupported CTU / min CU size");
  m_vb = flattenVirtualBoundaries(m_cs);
  m_out.hasLadf = m_cs.sps->getLadfEnabled();
  if (m_out.hasLadf)
  {
    m_out.ladf = vtmgpu_ladf{};
    m_out.ladf.num_intervals = m_cs.sps->getLadfNumIntervals();
    for (int k = 0; k < m_out.ladf.num_intervals && k < 5; k++)
    {
      m_out.ladf.qp_offset[k] = m_cs.sps->getLadfQpOffset(k);
      m_out.ladf.lower_bound[k] = m_cs.sps->getLadfIntervalLowerBound(k);
    }
  }
  m_out.width = W; m_out.height = H; m_out.sx = m_sx; m_out.sy = m_sy;
  const bool chroma = m_pcv.chrFormat != CHROMA_400;
  const int gx = 8 << m_sx, gy = 8 << m_sy;
  const size_t nL = (size_t)(W / 4) * (H / 4), nC[2] = { chroma ? (size_t)((W + gx - 1) / gx) * (H / 4) : 0, chroma ? (size_t)((H + gy - 1) / gy) * (W / 4) : 0 };
  if (m_out.listsValid && m_out.luma[VER].size() == nL && m_out.luma[HOR].size() == nL && m_out.chroma[VER].size() == nC[0] && m_out.chroma[HOR].size() == nC[1])
  {
    // same geometry as the previous picture: un-set its records instead of clearing 0.75 B per luma pixel
    for (int d = 0; d < 2; d++)
    {
      for (const auto& e : m_out.lists.luma[d]) m_out.luma[d][e.index] = 0;
      for (const auto& e : m_out.lists.chroma[d]) m_out.chroma[d][e.index] = 0;
    }
  }
  else
  {
    m_out.luma[VER].assign(nL, 0);
    m_out.luma[HOR].assign(nL, 0);
    m_out.chroma[VER].assign(nC[0], 0);
    m_out.chroma[HOR].assign(nC[1], 0);
  }
  for (int d = 0; d < 2; d++) { m_out.lists.luma[d].clear(); m_out.lists.chroma[d].clear(); }
  m_out.lists.twice = false;
  m_out.listsValid = false;

  // two passes over the picture as in loopFilterPic (LoopFilter.cpp:165-240).  No record depends on samples
  // (with LADF the records carry QPs and the device resolves them), so both passes can be derived before any filtering happens
  // -- and CTUs are independent of each other: all state is CTU-local (LoopFilter.cpp:169-175), every record lies inside its
  // own CTU and the coding structure is only read.  The CTU passes are therefore handed out dynamically to a few host threads
  // (VTMGPU_SHIM_THREADS, default min(16, cores); SURVEY.md 8f n1 "parallel host").
  int nThreads = 1;
  if (const char* e = getenv("VTMGPU_SHIM_THREADS")) nThreads = atoi(e);
  else nThreads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  nThreads = std::max(1, std::min(nThreads, (int)m_pcv.sizeInCtus));
  std::atomic<int> next(0);
  if (nThreads == 1) { m_lists = &m_out.lists; work(next); }
  else
  {
    std::vector<std::thread> pool;
    std::vector<std::exception_ptr> err(nThreads);
    std::vector<FlatDeblock::Lists> part(nThreads);
    for (int t = 0; t < nThreads; t++)
      pool.emplace_back([this, t, &err, &part, &next] {
        try { Deriver d(m_cs, m_out); d.m_lists = &part[t]; d.m_vb = m_vb; d.work(next); }
        catch (...) { err[t] = std::current_exception(); }
      });
    for (auto& th : pool) th.join();
    for (auto& e : err) if (e) std::rethrow_exception(e);
    for (int d = 0; d < 2; d++)
    {
      size_t nl = 0, nc = 0;
      for (const auto& q : part) { nl += q.luma[d].size(); nc += q.chroma[d].size(); }
      m_out.lists.luma[d].reserve(nl);
      m_out.lists.chroma[d].reserve(nc);
      for (const auto& q : part)
      {
        m_out.lists.luma[d].insert(m_out.lists.luma[d].end(), q.luma[d].begin(), q.luma[d].end());
        m_out.lists.chroma[d].insert(m_out.lists.chroma[d].end(), q.chroma[d].begin(), q.chroma[d].end());
        m_out.lists.twice |= q.twice;
      }
    }
  }
  m_out.listsValid = !m_out.lists.twice;
  // side effect of the reference's CTU loop (LoopFilter.cpp:179,218): cs.slice ends up as the slice of the last CTU
  m_cs.slice = m_cs.getCU(Position((m_pcv.widthInCtus - 1) << m_pcv.maxCUWidthLog2, (m_pcv.heightInCtus - 1) << m_pcv.maxCUHeightLog2), CH_L)->slice;
}

// CTU passes are handed out one at a time (all vertical-edge passes, then all horizontal-edge passes): every pass is independent
// of every other one -- CTU-local state (LoopFilter.cpp:169-175), records inside the CTU, read-only coding structure
void Deriver::work(std::atomic<int>& next)
{
  const int w = (int)m_pcv.widthInCtus, n = w * (int)m_pcv.heightInCtus;
  for (int i; (i = next.fetch_add(1, std::memory_order_relaxed)) < 2 * n;)
  {
    m_dir = i < n ? VER : HOR;
    const int a = i < n ? i : i - n, x = a % w, y = a / w;
    const UnitArea ctuArea(m_pcv.chrFormat, Area(x << m_pcv.maxCUWidthLog2, y << m_pcv.maxCUHeightLog2, m_pcv.maxCUWidth, m_pcv.maxCUWidth));
    m_ctuX = x << m_pcv.maxCUWidthLog2;
    m_ctuY = y << m_pcv.maxCUHeightLog2;
    ctuPass(ctuArea);
  }
}

void Deriver::ctuPass(const UnitArea& ctuArea)
{
  m_st.clear();
  m_ctuSlice = m_cs.getCU(ctuArea.lumaPos(), CH_L)->slice;    // the reference assigns this to cs.slice (LoopFilter.cpp:179,218); see run()
  for (auto& cu : m_cs.traverseCUs(CS::getArea(m_cs, ctuArea, CH_L), CH_L)) deriveCU(cu);
  if (CS::isDualITree(m_cs))
  {
    m_st.clear();
    for (auto& cu : m_cs.traverseCUs(CS::getArea(m_cs, ctuArea, CH_C), CH_C)) deriveCU(cu);
  }
}

void Deriver::edgeEnables(const CodingUnit& cu)
{
  // xSetLoopfilterParam (LoopFilter.cpp:656-672)
  m_left = m_top = m_internal = false;
  if (cu.slice->getDeblockingFilterDisable()) return;
  const Position& pos = cu.blocks[cu.chType].pos();
  const PPS& pps = *cu.cs->pps;
  m_internal = true;
  m_left = pos.x > 0 && usable(cu, *cu.cs->getCU(pos.offset(-1, 0), cu.chType), pps);
  m_top  = pos.y > 0 && usable(cu, *cu.cs->getCU(pos.offset(0, -1), cu.chType), pps);
}

void Deriver::mark(int dir, const Area& a, bool value, bool puInternal)
{
  // xSetEdgefilterMultiple (LoopFilter.cpp:627-655) for the left column (VER) / top row (HOR) of units of area a.
  // The reference updates the arrays of both directions but only ever reads those of the pass direction.
  if (dir != m_dir) return;
  const int n = dir == VER ? a.height / 4 : a.width / 4;
  for (int i = 0; i < n; i++)
  {
    const int id = dir == VER ? uidx(a.x, a.y + 4 * i) : uidx(a.x + 4 * i, a.y);
    m_st.on[id] = value;
    if (m_st.code[id] && value) m_st.code[id] = 3;            // TU edge that is also a PU edge
    else if (!puInternal)       m_st.code[id] = value;
  }
}

void Deriver::lengthsFromTU(const CodingUnit& cu, const TransformUnit& tu)
{
  // xSetMaxFilterLengthPQFromTransformSizes (LoopFilter.cpp:454-548)
  for (int c = 0; c < MAX_NUM_COMPONENT; c++)
  {
    const ComponentID comp = ComponentID(c);
    const ChannelType ch   = toChannelType(comp);
    const CompArea&   tb   = tu.block(comp);
    if (!tb.valid()) continue;
    const bool atCuBorder = m_dir == HOR ? tb.y == cu.block(comp).y : tb.x == cu.block(comp).x;
    if (!(atCuBorder ? (m_dir == HOR ? m_top : m_left) : m_internal)) continue;
    const int step   = 4 >> (c ? (m_dir == HOR ? m_sx : m_sy) : 0);
    const int extent = m_dir == HOR ? (int)tu.blocks[c].width : (int)tu.blocks[c].height;
    const int sizeQ  = m_dir == HOR ? tb.height : tb.width;
    const TransformUnit* tuPp = nullptr;                     // neighbouring TU of the previous unit: usually still the one across the edge
    for (int k = 0; k < extent; k += step)
    {
      const Position posQ = m_dir == HOR ? Position(tu.blocks[ch].x + k, tu.blocks[ch].y) : Position(tu.blocks[ch].x, tu.blocks[ch].y + k);
      const Position posP = m_dir == HOR ? posQ.offset(0, -1) : posQ.offset(-1, 0);
      if (!tuPp || !tuPp->blocks[ch].valid() || !tuPp->blocks[ch].contains(posP)) tuPp = cu.cs->getTU(posP, ch);
      const TransformUnit& tuP = *tuPp;
      const int sizeP = m_dir == HOR ? tuP.block(comp).height : tuP.block(comp).width;
      const int x = m_dir == HOR ? ux(c, tb.x + k) : ux(c, tb.x);
      const int y = m_dir == HOR ? uy(c, tb.y) : uy(c, tb.y + k);
      m_st.tuEdge[c][x][y] = true;
      if (c == 0)
      {
        const bool small = sizeP <= 4 || sizeQ <= 4;
        m_st.lenQ[c][x][y] = small ? 1 : (sizeQ >= 32 ? 7 : 3);
        m_st.lenP[c][x][y] = small ? 1 : (sizeP >= 32 ? 7 : 3);
      }
      else
      {
        m_st.lenQ[c][x][y] = m_st.lenP[c][x][y] = (sizeQ >= 8 && sizeP >= 8) ? 3 : 1;
      }
    }
  }
}

void Deriver::lengthsFromSubBlocks(const PredictionUnit& pu, bool mvSub, int sub, const Area& ap)
{
  // xSetMaxFilterLengthPQForCodingSubBlocks (LoopFilter.cpp:550-625); luma only, a = along the edge, d = across
  if (!mvSub || !pu.Y().valid()) return;
  const int x0 = ux(0, pu.Y().x), y0 = uy(0, pu.Y().y);
  const int across = m_dir == HOR ? ap.height : ap.width, along = m_dir == HOR ? ap.width : ap.height;
  auto te   = [&](int d, int a) -> bool { return m_dir == HOR ? m_st.tuEdge[0][x0 + a / 4][y0 + d / 4] : m_st.tuEdge[0][x0 + d / 4][y0 + a / 4]; };
  auto lenP = [&](int d, int a) -> uint8_t& { return m_dir == HOR ? m_st.lenP[0][x0 + a / 4][y0 + d / 4] : m_st.lenP[0][x0 + d / 4][y0 + a / 4]; };
  auto lenQ = [&](int d, int a) -> uint8_t& { return m_dir == HOR ? m_st.lenQ[0][x0 + a / 4][y0 + d / 4] : m_st.lenQ[0][x0 + d / 4][y0 + a / 4]; };
  for (int d = 0; d < across; d += sub)
    for (int a = 0; a < along; a += 4)
    {
      if (te(d, a))
      {
        lenQ(d, a) = std::min<int>(lenQ(d, a), 5);
        if (d > 0) lenP(d, a) = std::min<int>(lenP(d, a), 5);
      }
      else if (d > 0 && (te(d - 4, a) || d + 4 >= across || te(d + 4, a)))   lenQ(d, a) = lenP(d, a) = 1;
      else if (d > 0 && (te(d - 8, a) || d + 8 >= across || te(d + 8, a)))   lenQ(d, a) = lenP(d, a) = 2;
      else                                                                   lenQ(d, a) = lenP(d, a) = 3;
    }
}

unsigned Deriver::strength(const CodingUnit& cu, const Position& lumaPos) const
{
  // xGetBoundaryStrengthSingle (LoopFilter.cpp:674-812); result packed Y | Cb<<2 | Cr<<4
  const bool hasLuma = cu.Y().valid();
  const int  sh = hasLuma ? 0 : getComponentScaleX(COMPONENT_Cb, cu.firstPU->chromaFormat);
  const int  sv = hasLuma ? 0 : getComponentScaleY(COMPONENT_Cb, cu.firstPU->chromaFormat);
  const Position posQ(lumaPos.x >> sh, lumaPos.y >> sv);
  const Position posP = m_dir == VER ? posQ.offset(-1, 0) : posQ.offset(0, -1);
  const CodingUnit& cuQ = cu;
  // internal edges (TU / PU / sub-block boundaries inside the CU): the P side is the CU itself, no look-up
  const CodingUnit& cuP = cu.blocks[cu.chType].contains(posP) ? cu : *cu.cs->getCU(posP, cu.chType);
  const bool intraP = cuP.predMode == MODE_INTRA, intraQ = cuQ.predMode == MODE_INTRA;
  if (intraP || intraQ)
  {
    const unsigned y = (intraP && cuP.bdpcmMode && intraQ && cuQ.bdpcmMode) ? 0 : 2;
    const unsigned c = (intraP && cuP.bdpcmModeChroma && intraQ && cuQ.bdpcmModeChroma) ? 0 : 2;
    return y | (c << 2) | (c << 4);
  }
  // consecutive units along an edge mostly stay inside the same two transform units: one-entry memo per side
  if (!m_tuQ || m_tuQ->cu != &cuQ || !m_tuQ->blocks[cuQ.chType].valid() || !m_tuQ->blocks[cuQ.chType].contains(posQ)) m_tuQ = cuQ.cs->getTU(posQ, cuQ.chType);
  if (!m_tuP || m_tuP->cu != &cuP || !m_tuP->blocks[cuQ.chType].valid() || !m_tuP->blocks[cuQ.chType].contains(posP)) m_tuP = cuP.cs->getTU(posP, cuQ.chType);
  const TransformUnit& tuQ = *m_tuQ;
  const TransformUnit& tuP = *m_tuP;
  const uint8_t code = m_st.code[uidx(lumaPos.x, lumaPos.y)];
  const bool ciip = cuP.firstPU->ciipFlag || cuQ.firstPU->ciipFlag;
  if (code && ciip) return 2 | (2 << 2) | (2 << 4);

  unsigned bs = 0;
  if (code)
  {
    if (TU::getCbf(tuQ, COMPONENT_Y) || TU::getCbf(tuP, COMPONENT_Y)) bs |= 1;
    if (TU::getCbf(tuQ, COMPONENT_Cb) || TU::getCbf(tuP, COMPONENT_Cb) || tuQ.jointCbCr || tuP.jointCbCr) bs |= 1 << 2;
    if (TU::getCbf(tuQ, COMPONENT_Cr) || TU::getCbf(tuP, COMPONENT_Cr) || tuQ.jointCbCr || tuP.jointCbCr) bs |= 1 << 4;
  }
  if ((bs & 3) == 1) return bs;
  if (ciip) return 1;
  if (!hasLuma) return bs;
  if (code != 0 && code != 3) return bs;          // pure TU edge: no motion test

  // motion test, on the UNREFINED motion field (setRefinedMotionField runs after deblocking, DecLib.cpp:579-580)
  const Position lumaPosP = m_dir == VER ? lumaPos.offset(-1, 0) : lumaPos.offset(0, -1);
  const MotionInfo& miQ = cuQ.cs->getMotionInfo(lumaPos);
  const MotionInfo& miP = cuP.cs->getMotionInfo(lumaPosP);
  const Slice& sliceQ = *cuQ.slice;
  const Slice& sliceP = *cuP.slice;
  const int thr = (1 << MV_FRACTIONAL_BITS_INTERNAL) >> 1;
  auto far = [thr](const Mv& a, const Mv& b) { return abs(a.getHor() - b.getHor()) >= thr || abs(a.getVer() - b.getVer()) >= thr; };

  if (sliceQ.isInterB() || sliceP.isInterB())
  {
    const Picture* refP0 = CU::isIBC(cuP) ? sliceP.getPic() : (miP.refIdx[0] < 0 ? nullptr : sliceP.getRefPic(REF_PIC_LIST_0, miP.refIdx[0]));
    const Picture* refP1 = CU::isIBC(cuP) ? nullptr         : (miP.refIdx[1] < 0 ? nullptr : sliceP.getRefPic(REF_PIC_LIST_1, miP.refIdx[1]));
    const Picture* refQ0 = CU::isIBC(cuQ) ? sliceQ.getPic() : (miQ.refIdx[0] < 0 ? nullptr : sliceQ.getRefPic(REF_PIC_LIST_0, miQ.refIdx[0]));
    const Picture* refQ1 = CU::isIBC(cuQ) ? nullptr         : (miQ.refIdx[1] < 0 ? nullptr : sliceQ.getRefPic(REF_PIC_LIST_1, miQ.refIdx[1]));
    Mv p0, p1, q0, q1;   // zero unless the list is used
    if (miP.refIdx[0] >= 0) p0 = miP.mv[0];
    if (miP.refIdx[1] >= 0) p1 = miP.mv[1];
    if (miQ.refIdx[0] >= 0) q0 = miQ.mv[0];
    if (miQ.refIdx[1] >= 0) q1 = miQ.mv[1];
    unsigned mvBs = 1;                         // different reference pictures
    if ((refP0 == refQ0 && refP1 == refQ1) || (refP0 == refQ1 && refP1 == refQ0))
    {
      if (refP0 != refP1)                      // two distinct references
        mvBs = refP0 == refQ0 ? (far(q0, p0) || far(q1, p1)) : (far(q1, p0) || far(q0, p1));
      else                                     // both lists point at the same picture
        mvBs = (far(q0, p0) || far(q1, p1)) && (far(q1, p0) || far(q0, p1));
    }
    return bs + mvBs;
  }
  // P slices
  const Picture* refP0 = CU::isIBC(cuP) ? sliceP.getPic() : sliceP.getRefPic(REF_PIC_LIST_0, miP.refIdx[0]);
  const Picture* refQ0 = CU::isIBC(cuQ) ? sliceQ.getPic() : sliceQ.getRefPic(REF_PIC_LIST_0, miQ.refIdx[0]);
  if (refP0 != refQ0) return bs + 1;
  return far(miQ.mv[0], miP.mv[0]) ? bs + 1 : bs;
}

void Deriver::emitLuma(const CodingUnit& cu, int edge)
{
  // xEdgeFilterLuma (LoopFilter.cpp:844-977): everything up to the sample reads
  const CompArea& la = cu.block(COMPONENT_Y);
  const SPS& sps = *cu.cs->sps;
  const PPS& pps = *cu.cs->pps;
  const Slice& slice = *cu.slice;
  const int bd = sps.getBitDepth(CHANNEL_TYPE_LUMA);
  const ClpRng& clp = m_ctuSlice->clpRng(COMPONENT_Y);
  CHECK(clp.min != 0 || clp.max != (1 << bd) - 1, "vtmgpu shim: non-default luma clipping range");
  const int n = m_dir == VER ? la.height / 4 : la.width / 4;
  const int tcOff = slice.getDeblockingFilterTcOffsetDiv2() * 2, betaOff = slice.getDeblockingFilterBetaOffsetDiv2() * 2;
  for (int i = 0; i < n; i++)
  {
    const Position pos = m_dir == VER ? Position(la.x + edge * 4, la.y + i * 4) : Position(la.x + i * 4, la.y + edge * 4);
    const int id = uidx(pos.x, pos.y);
    const unsigned bs = m_st.code[id] & 3;
    if (!bs) continue;
    const Position posP = m_dir == VER ? pos.offset(-4, 0) : pos.offset(0, -4);
    const CodingUnit& cuP = la.contains(posP) ? cu : *cu.cs->getCU(posP, cu.chType);
    if (!usable(cu, cuP, pps))
    {
      m_st.code[id] = 0;      // also suppresses the chroma filtering of this unit (LoopFilter.cpp:918-933)
      continue;
    }
    const int qp = (cuP.qp + cu.qp + 1) >> 1;
    int lp = m_st.lenP[0][(pos.x - m_ctuX) >> 2][(pos.y - m_ctuY) >> 2];
    const int lq = m_st.lenQ[0][(pos.x - m_ctuX) >> 2][(pos.y - m_ctuY) >> 2];
    if (lp > 5 && cuP.affine) lp = 5;
    const bool ctuRow = m_dir == HOR && pos.y % (int)slice.getSPS()->getCTUSize() == 0;
    const int iTc = Clip3(0, MAX_QP + 2, qp + 2 * (int(bs) - 1) + tcOff);
    const int iB  = Clip3(0, MAX_QP, qp + betaOff);
    unsigned tc   = bd < 10 ? (kTc[iTc] + 2) >> (10 - bd) : kTc[iTc] << (bd - 10);
    unsigned beta = kBeta[iB] << (bd - 8);
    if (sps.getLadfEnabled())
    {
      // LADF: the QP offset depends on reconstructed samples (deriveLADFShift, LoopFilter.cpp:815-841) -> the record carries
      // the QPs before the offset and the device derives tc / beta (include/vtmgpu.h, vtmgpu_ladf)
      tc   = VTMGPU_DBF_LADF_BIAS + qp + 2 * (int(bs) - 1) + tcOff;
      beta = VTMGPU_DBF_LADF_BIAS + qp + betaOff;
    }
    CHECK(tc > 0x7ff || beta > 0x7ff, "vtmgpu shim: tc/beta out of record range");
    uint32_t rec = tc | (beta << VTMGPU_DBF_L_BETA_SHIFT) | (uint32_t(lp) << VTMGPU_DBF_L_LENP_SHIFT) | (uint32_t(lq) << VTMGPU_DBF_L_LENQ_SHIFT);
    if (sps.getPLTMode())
    {
      if (CU::isPLT(cuP)) rec |= VTMGPU_DBF_L_PNOFILT;
      if (CU::isPLT(cu))  rec |= VTMGPU_DBF_L_QNOFILT;
    }
    if (ctuRow) rec |= VTMGPU_DBF_L_CTUROW;
    if (!tc) rec = 0;
    putLuma((size_t)(pos.y / 4) * (m_out.width / 4) + pos.x / 4, rec);
  }
}

void Deriver::emitChroma(const CodingUnit& cu, int edge)
{
  // xEdgeFilterChroma (LoopFilter.cpp:1087-1249): everything up to the sample reads
  const bool hasLuma = cu.Y().valid();
  const Position lumaPos = hasLuma ? cu.Y().pos() : recalcPosition(cu.chromaFormat, cu.chType, CHANNEL_TYPE_LUMA, cu.blocks[cu.chType].pos());
  const Size     lumaSize = hasLuma ? cu.Y().size() : recalcSize(cu.chromaFormat, cu.chType, CHANNEL_TYPE_LUMA, cu.blocks[cu.chType].size());
  const SPS& sps = *cu.cs->sps;
  const PPS& pps = *cu.cs->pps;
  const Slice& slice = *cu.slice;
  const int unitH = 4 >> m_sx, unitV = 4 >> m_sy;          // chroma samples per 4-luma unit
  // only edges on the 8x8 chroma-sample grid, measured from the CTU origin (LoopFilter.cpp:1115-1126)
  const int edgeInCtu = (m_dir == VER ? ((lumaPos.x & (int)m_pcv.maxCUWidthMask) >> 2) : ((lumaPos.y & (int)m_pcv.maxCUHeightMask) >> 2)) + edge;
  if (unitH < 8 && unitV < 8 && (edgeInCtu % (8 / (m_dir == VER ? unitH : unitV))) != 0) return;

  const int bd = sps.getBitDepth(CHANNEL_TYPE_CHROMA);
  const int n = m_dir == VER ? lumaSize.height / 4 : lumaSize.width / 4;
  const int tcOff = slice.getDeblockingFilterTcOffsetDiv2() * 2, betaOff = slice.getDeblockingFilterBetaOffsetDiv2() * 2;
  const int gx = 8 << m_sx, gy = 8 << m_sy;
  for (int i = 0; i < n; i++)
  {
    const Position pos = m_dir == VER ? Position(lumaPos.x + edge * 4, lumaPos.y + i * 4) : Position(lumaPos.x + i * 4, lumaPos.y + edge * 4);
    const unsigned packed = m_st.code[uidx(pos.x, pos.y)];
    const unsigned bsC[2] = { (packed >> 2) & 3, (packed >> 4) & 3 };
    if (!bsC[0] && !bsC[1]) continue;
    const Position lumaP = m_dir == VER ? pos.offset(-4, 0) : pos.offset(0, -4);
    const CodingUnit& cuP1 = *cu.cs->getCU(recalcPosition(cu.chromaFormat, CHANNEL_TYPE_LUMA, cu.chType, lumaP), cu.chType);
    const ChannelType chP  = cuP1.isSepTree() ? CHANNEL_TYPE_CHROMA : cu.chType;
    const CodingUnit& cuP  = *cu.cs->getCU(recalcPosition(cu.chromaFormat, CHANNEL_TYPE_LUMA, chP, lumaP), chP);
    CHECK(!usable(cu, cuP, pps), "Neighbour not available");

    const int x = (pos.x - m_ctuX) >> 2, y = (pos.y - m_ctuY) >> 2;
    const bool large = m_st.lenP[COMPONENT_Cb][x][y] >= 3 && m_st.lenQ[COMPONENT_Cb][x][y] >= 3;
    const bool ctb   = m_dir == HOR && pos.y % (int)cuP.slice->getSPS()->getCTUSize() == 0;
    uint64_t rec = 0;
    const TransformUnit *tuQp = nullptr, *tuPp = nullptr;      // the same for Cb and Cr: looked up once per unit, when needed
    for (int c = 0; c < 2; c++)
    {
      if (!(bsC[c] == 2 || (large && bsC[c] == 1))) continue;
      const ComponentID comp = ComponentID(c + 1);
      const ClpRng& clp = m_ctuSlice->clpRng(comp);
      CHECK(clp.min != 0 || clp.max != (1 << bd) - 1, "vtmgpu shim: non-default chroma clipping range");
      if (!tuQp)
      {
        const int shP = cuP.Y().valid() ? 0 : getComponentScaleX(COMPONENT_Cb, cuP.firstPU->chromaFormat);
        const int svP = cuP.Y().valid() ? 0 : getComponentScaleY(COMPONENT_Cb, cuP.firstPU->chromaFormat);
        const int shQ = hasLuma ? 0 : getComponentScaleX(COMPONENT_Cb, cu.firstPU->chromaFormat);
        const int svQ = hasLuma ? 0 : getComponentScaleY(COMPONENT_Cb, cu.firstPU->chromaFormat);
        const Position posQ(pos.x >> shQ, pos.y >> svQ);
        const Position posP1(pos.x >> shP, pos.y >> svP);
        const Position posP = m_dir == VER ? posP1.offset(-1, 0) : posP1.offset(0, -1);
        tuQp = cu.cs->getTU(posQ, cu.chType);
        tuPp = cuP.cs->getTU(posP, cuP.chType);
      }
      const TransformUnit& tuQ = *tuQp;
      const TransformUnit& tuP = *tuPp;
      const QpParam qpP(tuP, comp, -MAX_INT, false);
      const QpParam qpQ(tuQ, comp, -MAX_INT, false);
      const int bdOff = tuP.cs->sps->getQpBDOffset(toChannelType(comp));
      const int qp = ((qpQ.Qp(0) - bdOff) + (qpP.Qp(0) - bdOff) + 1) >> 1;
      const int iTc = Clip3<int>(0, MAX_QP + 2, qp + 2 * (int(bsC[c]) - 1) + tcOff);
      const uint64_t tc = bd < 10 ? (kTc[iTc] + 2) >> (10 - bd) : kTc[iTc] << (bd - 10);
      uint64_t beta = 0;
      if (large) beta = kBeta[Clip3<int>(0, MAX_QP, qp + betaOff)] << (bd - 8);
      CHECK(tc > 0x7ff || beta > 0x7ff, "vtmgpu shim: tc/beta out of record range");
      rec |= tc << (c ? VTMGPU_DBF_C_TCCR_SHIFT : 0);
      rec |= beta << (c ? VTMGPU_DBF_C_BETACR_SHIFT : VTMGPU_DBF_C_BETACB_SHIFT);
    }
    if (!(rec & 0x3fffff)) continue;           // both tc zero: nothing to do
    if (large) rec |= VTMGPU_DBF_C_LARGE;
    if (ctb)   rec |= VTMGPU_DBF_C_CTB;
    if (sps.getPLTMode())
    {
      if (CU::isPLT(cuP)) rec |= VTMGPU_DBF_C_PNOFILT;
      if (CU::isPLT(cu))  rec |= VTMGPU_DBF_C_QNOFILT;
    }
    if (m_dir == VER) putChroma((size_t)(pos.y / 4) * ((m_out.width + gx - 1) / gx) + pos.x / gx, rec);
    else              putChroma((size_t)(pos.y / gy) * (m_out.width / 4) + pos.x / 4, rec);
  }
}

void Deriver::deriveCU(CodingUnit& cu)
{
  // xDeblockCU (LoopFilter.cpp:261-408)
  const bool hasLuma = cu.Y().valid();
  const Area area = hasLuma ? Area(cu.Y())
                            : Area(recalcPosition(cu.chromaFormat, cu.chType, CHANNEL_TYPE_LUMA, cu.blocks[cu.chType].pos()),
                                   recalcSize(cu.chromaFormat, cu.chType, CHANNEL_TYPE_LUMA, cu.blocks[cu.chType].size()));
  edgeEnables(cu);
  std::vector<int> edges;
  const CompArea& own = cu.blocks[cu.chType];
  auto edgeOf = [&](const CompArea& b, int off) { return m_dir == HOR ? (b.y + off - own.y) / 4 : (b.x + off - own.x) / 4; };

  for (auto& tu : CU::traverseTUs(cu))
  {
    const Area at = hasLuma ? Area(tu.block(COMPONENT_Y)) : area;
    mark(VER, at, m_internal && !onVb(VER, at), false);
    mark(HOR, at, m_internal && !onVb(HOR, at), false);
    lengthsFromTU(cu, tu);
    edges.push_back(edgeOf(tu.blocks[cu.chType], 0));
  }

  bool mvSub = false;
  const int sub = 8;
  for (auto& pu : CU::traversePUs(cu))
  {
    const Area ap = hasLuma ? Area(pu.block(COMPONENT_Y)) : area;
    const bool xOff = pu.blocks[cu.chType].x != own.x, yOff = pu.blocks[cu.chType].y != own.y;
    mark(VER, ap, (xOff ? m_internal : m_left) && !onVb(VER, ap), xOff);
    mark(HOR, ap, (yOff ? m_internal : m_top) && !onVb(HOR, ap), yOff);
    edges.push_back(edgeOf(pu.blocks[cu.chType], 0));
    if ((pu.mergeFlag && pu.mergeType == MRG_TYPE_SUBPU_ATMVP) || cu.affine)
    {
      mvSub = true;
      const int extent = m_dir == HOR ? ap.height : ap.width;
      for (int off = sub; off < extent; off += sub)
      {
        const Area blk = m_dir == HOR ? Area(cu.Y().x, cu.Y().y + off, cu.Y().width, 4) : Area(cu.Y().x + off, cu.Y().y, 4, cu.Y().height);
        mark(m_dir, blk, m_internal && !onVb(m_dir, blk), true);
        edges.push_back(edgeOf(pu.blocks[cu.chType], off));
      }
    }
    lengthsFromSubBlocks(pu, mvSub, sub, ap);
  }

  for (int y = 0; y < (int)area.height; y += 4)
    for (int x = 0; x < (int)area.width; x += 4)
    {
      const int id = uidx(area.x + x, area.y + y);
      if (m_st.on[id]) m_st.code[id] = (uint8_t)strength(cu, Position(area.x + x, area.y + y));
    }

  std::sort(edges.begin(), edges.end());
  edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
  for (int e : edges)
  {
    if (cu.blocks[COMPONENT_Y].valid()) emitLuma(cu, e);
    if (cu.blocks[COMPONENT_Cb].valid() && m_pcv.chrFormat != CHROMA_400 && (!cu.ispMode || e == 0)) emitChroma(cu, e);
  }
}

}   // namespace

void deriveDeblockRecords(CodingStructure& cs, FlatDeblock& out)
{
  Deriver d(cs, out);
  d.run();
}

// ---------------------------------------------------------------------------------------------------------
// block structure for the device-side derivation
// ---------------------------------------------------------------------------------------------------------
FlatUnits::~FlatUnits() { free(block); }

const vtmgpu_deblock_units* FlatUnits::view()
{
  p.num_cus = (int)numCus; p.num_tus = (int)numTus; p.num_slices = (int)slices.size();
  p.cus = cus; p.tus = tus; p.slices = slices.data();
  p.tu_luma = tuLuma;
  p.tu_chroma = tuChroma;
  p.ladf = hasLadf ? &ladf : nullptr;
  p.vb = (vb.num_ver || vb.num_hor) ? &vb : nullptr;
  return &p;
}

namespace
{
// the CTUs are handed out to a few host threads (they only read the coding structure; every CTU writes its own CUs / TUs -- reserved with
// one atomic add per CTU -- and its own part of the unit maps)
struct UnitFlattener
{
  CodingStructure& cs;
  const PreCalcValues& pcv;
  FlatUnits& out;
  int sx, sy, w4, h4;
  bool chroma, dual;
  std::vector<const Slice*> sliceOf;
  std::atomic<size_t> nCus{ 0 }, nTus{ 0 };
  std::atomic<int> nextCtu{ 0 };
  std::atomic<bool> unsupported{ false }, anyInter{ false };
  UnitFlattener(CodingStructure& c, FlatUnits& o)
    : cs(c), pcv(*c.pcv), out(o), sx((int)getComponentScaleX(COMPONENT_Cb, c.pcv->chrFormat)), sy((int)getComponentScaleY(COMPONENT_Cb, c.pcv->chrFormat)),
      w4((int)c.pcv->lumaWidth / 4), h4((int)c.pcv->lumaHeight / 4), chroma(c.pcv->chrFormat != CHROMA_400), dual(CS::isDualITree(c)) {}

  void addSlice(const Slice* s, std::vector<const Picture*>& pics)
  {
    vtmgpu_dbf_slice d{};
    d.tc_offset = (int8_t)(s->getDeblockingFilterTcOffsetDiv2() * 2);
    d.beta_offset = (int8_t)(s->getDeblockingFilterBetaOffsetDiv2() * 2);
    d.inter_b = s->isInterB();
    d.independent_idx = (uint16_t)s->getIndependentSliceIdx();
    for (int l = 0; l < 2; l++)
      for (int r = 0; r < 16; r++)
      {
        d.ref_pic[l][r] = -1;
        if (s->isIntra() || r >= s->getNumRefIdx(RefPicList(l))) continue;
        const Picture* pic = s->getRefPic(RefPicList(l), r);
        size_t k = 0;
        while (k < pics.size() && pics[k] != pic) k++;
        if (k == pics.size()) pics.push_back(pic);
        d.ref_pic[l][r] = (int16_t)k;
      }
    sliceOf.push_back(s);
    out.slices.push_back(d);
  }
  int sliceIndex(const Slice* s) const
  {
    for (size_t i = 0; i < sliceOf.size(); i++) if (sliceOf[i] == s) return (int)i;
    THROW("vtmgpu shim: a CU refers to a slice that is not in the picture's slice list");
  }

  void fill(uint32_t* map, int x, int y, int w, int h, uint32_t id) const      // luma sample rectangle -> 4x4 units
  {
    // a unit belongs to the block that contains its top-left sample (what getTU(pos) of that sample returns; ISP sub-partitions can be narrower than a unit)
    const int x0 = (x + 3) >> 2, y0 = (y + 3) >> 2, x1 = std::min(w4, (x + w + 3) >> 2), y1 = std::min(h4, (y + h + 3) >> 2);
    for (int yy = y0; yy < y1; yy++) std::fill(map + (size_t)yy * w4 + x0, map + (size_t)yy * w4 + x1, id);
  }

  void addCU(const CodingUnit& cu, size_t& cuAt, size_t& tuAt)
  {
    const bool hasLuma = cu.Y().valid(), hasChroma = chroma && cu.blocks.size() > 1 && cu.blocks[COMPONENT_Cb].valid();
    const Area area = hasLuma ? Area(cu.Y())
                              : Area(recalcPosition(cu.chromaFormat, cu.chType, CHANNEL_TYPE_LUMA, cu.blocks[cu.chType].pos()),
                                     recalcSize(cu.chromaFormat, cu.chType, CHANNEL_TYPE_LUMA, cu.blocks[cu.chType].size()));
    vtmgpu_dbf_cu c{};
    c.x = (uint16_t)area.x; c.y = (uint16_t)area.y; c.w4 = (uint8_t)(area.width / 4); c.h4 = (uint8_t)(area.height / 4);
    c.qp = cu.qp;
    c.slice = (uint8_t)sliceIndex(cu.slice);
    c.tile = (uint16_t)cu.tileIdx;
    unsigned f = 0;
    if (cu.predMode == MODE_INTRA) f |= VTMGPU_CU_INTRA; else anyInter = true;
    if (CU::isIBC(cu)) f |= VTMGPU_CU_IBC;
    if (CU::isPLT(cu)) f |= VTMGPU_CU_PLT;
    if (cu.affine) f |= VTMGPU_CU_AFFINE;
    if (cu.bdpcmMode) f |= VTMGPU_CU_BDPCM;
    if (cu.bdpcmModeChroma) f |= VTMGPU_CU_BDPCM_C;
    if (cu.firstPU->ciipFlag) f |= VTMGPU_CU_CIIP;
    if (cu.ispMode) f |= VTMGPU_CU_ISP;
    if (hasLuma) f |= VTMGPU_CU_HAS_LUMA;
    if (hasChroma) f |= VTMGPU_CU_HAS_CHROMA;
    // the unit model describes CUs with one prediction unit that covers the CU (all VVC prediction modes); anything else takes the CU walk
    const PredictionUnit& pu = *cu.firstPU;
    if ((pu.mergeFlag && pu.mergeType == MRG_TYPE_SUBPU_ATMVP) || cu.affine) f |= VTMGPU_CU_MVSUB;
    if (pu.next && pu.next->cu == &cu) unsupported = true;
    if (pu.blocks[cu.chType].pos() != cu.blocks[cu.chType].pos() || pu.blocks[cu.chType].size() != cu.blocks[cu.chType].size()) unsupported = true;
    // xSetLoopfilterParam (LoopFilter.cpp:656-672)
    if (!cu.slice->getDeblockingFilterDisable())
    {
      const Position& pos = cu.blocks[cu.chType].pos();
      const PPS& pps = *cu.cs->pps;
      f |= VTMGPU_CU_EN_INT;
      if (pos.x > 0 && usable(cu, *cu.cs->getCU(pos.offset(-1, 0), cu.chType), pps)) f |= VTMGPU_CU_EN_LEFT;
      if (pos.y > 0 && usable(cu, *cu.cs->getCU(pos.offset(0, -1), cu.chType), pps)) f |= VTMGPU_CU_EN_TOP;
    }
    c.flags = (uint16_t)f;
    const uint32_t cuId = (uint32_t)cuAt;
    out.cus[cuAt++] = c;
    for (auto& tu : CU::traverseTUs(cu))
    {
      vtmgpu_dbf_tu t{};
      t.cu = cuId;
      const uint32_t tuId = (uint32_t)tuAt;
      t.tail = tuId;
      if (tu.blocks[COMPONENT_Y].valid())
      {
        const CompArea& b = tu.blocks[COMPONENT_Y];
        t.x = (uint16_t)b.x; t.y = (uint16_t)b.y; t.w = (uint8_t)b.width; t.h = (uint8_t)b.height;
        fill(out.tuLuma, b.x, b.y, b.width, b.height, tuId);
        // ISP sub-partitions 1 or 2 samples wide (high): they follow each other in the TU list, the one that holds the unit's last
        // column (row) is 3 / width (3 / height) entries further
        if (b.width < 4 && (b.x & 3) == 0)       t.tail = tuId + 3 / b.width;
        else if (b.height < 4 && (b.y & 3) == 0) t.tail = tuId + 3 / b.height;
      }
      if (chroma && tu.blocks.size() > 1 && tu.blocks[COMPONENT_Cb].valid())
      {
        const CompArea& b = tu.blocks[COMPONENT_Cb];
        t.cx = (uint16_t)b.x; t.cy = (uint16_t)b.y; t.cw = (uint8_t)b.width; t.ch = (uint8_t)b.height;
        fill(out.tuChroma, b.x << sx, b.y << sy, b.width << sx, b.height << sy, tuId);
      }
      if (TU::getCbf(tu, COMPONENT_Y)) t.cbf |= VTMGPU_TU_CBF_Y;
      if (chroma)
      {
        if (TU::getCbf(tu, COMPONENT_Cb)) t.cbf |= VTMGPU_TU_CBF_CB;
        if (TU::getCbf(tu, COMPONENT_Cr)) t.cbf |= VTMGPU_TU_CBF_CR;
        if (tu.jointCbCr) t.cbf |= VTMGPU_TU_JOINT;
        // chroma QPs as xEdgeFilterChroma takes them (LoopFilter.cpp:1213-1217)
        const int bdOff = tu.cs->sps->getQpBDOffset(CHANNEL_TYPE_CHROMA);
        t.qp_cb = (int8_t)(QpParam(tu, COMPONENT_Cb, -MAX_INT, false).Qp(0) - bdOff);
        t.qp_cr = (int8_t)(QpParam(tu, COMPONENT_Cr, -MAX_INT, false).Qp(0) - bdOff);
      }
      out.tus[tuAt++] = t;
    }
  }

  void ctu(int a)
  {
    const int cx = a % (int)pcv.widthInCtus, cy = a / (int)pcv.widthInCtus;
    const UnitArea ctuArea(pcv.chrFormat, Area(cx << pcv.maxCUWidthLog2, cy << pcv.maxCUHeightLog2, pcv.maxCUWidth, pcv.maxCUWidth));
    // count, reserve, fill
    size_t ncu = 0, ntu = 0;
    for (int tree = 0; tree < (dual ? 2 : 1); tree++)
      for (auto& cu : cs.traverseCUs(CS::getArea(cs, ctuArea, tree ? CH_C : CH_L), tree ? CH_C : CH_L))
      {
        ncu++;
        for (auto& tu : CU::traverseTUs(cu)) { (void)tu; ntu++; }
      }
    size_t cuAt = nCus.fetch_add(ncu), tuAt = nTus.fetch_add(ntu);
    if (cuAt + ncu > out.capCus || tuAt + ntu > out.capTus) { unsupported = true; return; }
    for (int tree = 0; tree < (dual ? 2 : 1); tree++)
      for (auto& cu : cs.traverseCUs(CS::getArea(cs, ctuArea, tree ? CH_C : CH_L), tree ? CH_C : CH_L)) addCU(cu, cuAt, tuAt);
  }

  void work()
  {
    const int n = (int)pcv.sizeInCtus;
    for (int a; (a = nextCtu.fetch_add(1, std::memory_order_relaxed)) < n;) ctu(a);
  }
};
}   // namespace

void flattenUnits(CodingStructure& cs, FlatUnits& out)
{
  const PreCalcValues& pcv = *cs.pcv;
  CHECK(pcv.maxCUWidth > 128 || pcv.minCUWidth != 4 || pcv.minCUHeight != 4, "vtmgpu shim: unsupported CTU / min CU size");
  const bool chroma = pcv.chrFormat != CHROMA_400;
  const size_t units = (size_t)(pcv.lumaWidth / 4) * (pcv.lumaHeight / 4);
  out.blockIsNew = false;
  if (units != out.units || !out.block)
  {
    // capacities: a picture rarely has more CUs / TUs than a third of its units; a picture that does takes the CU walk
    free(out.block);
    out.units = units;
    out.capCus = out.capTus = units;
    const size_t a = (sizeof(vtmgpu_dbf_cu) * out.capCus + 4095) & ~size_t(4095), b = (sizeof(vtmgpu_dbf_tu) * out.capTus + 4095) & ~size_t(4095), m = (units * 4 + 4095) & ~size_t(4095);
    out.blockBytes = a + b + 2 * m;
    out.block = nullptr;
    CHECK(posix_memalign(&out.block, 4096, out.blockBytes) != 0, "vtmgpu shim: out of memory");
    unsigned char* p = static_cast<unsigned char*>(out.block);
    out.cus = reinterpret_cast<vtmgpu_dbf_cu*>(p);
    out.tus = reinterpret_cast<vtmgpu_dbf_tu*>(p + a);
    out.tuLuma = reinterpret_cast<uint32_t*>(p + a + b);
    out.tuChroma = reinterpret_cast<uint32_t*>(p + a + b + m);
    out.blockIsNew = true;
  }
  out.slices.clear();
  out.supported = true;
  out.vb = flattenVirtualBoundaries(cs);
  out.hasLadf = cs.sps->getLadfEnabled();
  if (out.hasLadf)
  {
    out.ladf = vtmgpu_ladf{};
    out.ladf.num_intervals = cs.sps->getLadfNumIntervals();
    for (int k = 0; k < out.ladf.num_intervals && k < 5; k++)
    {
      out.ladf.qp_offset[k] = cs.sps->getLadfQpOffset(k);
      out.ladf.lower_bound[k] = cs.sps->getLadfIntervalLowerBound(k);
    }
  }
  UnitFlattener F(cs, out);
  {
    std::vector<const Picture*> pics;
    CHECK(cs.picture->slices.size() > 255, "vtmgpu shim: more than 255 slices in a picture");
    for (const Slice* s : cs.picture->slices) F.addSlice(s, pics);
  }
  int nThreads = 1;
  if (const char* e = getenv("VTMGPU_SHIM_THREADS")) nThreads = atoi(e);
  else nThreads = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
  nThreads = std::max(1, std::min(nThreads, (int)pcv.sizeInCtus));
  if (nThreads == 1) F.work();
  else
  {
    std::vector<std::thread> pool;
    std::vector<std::exception_ptr> err(nThreads);
    for (int t = 0; t < nThreads; t++) pool.emplace_back([&F, &err, t] { try { F.work(); } catch (...) { err[t] = std::current_exception(); } });
    for (auto& th : pool) th.join();
    for (auto& e : err) if (e) std::rethrow_exception(e);
  }
  out.numCus = std::min(F.nCus.load(), out.capCus);
  out.numTus = std::min(F.nTus.load(), out.capTus);
  out.supported = !F.unsupported;
  vtmgpu_deblock_units& p = out.p;
  p = vtmgpu_deblock_units{};
  p.flags = (cs.pps->getLoopFilterAcrossSlicesEnabledFlag() ? VTMGPU_UNITS_ACROSS_SLICES : 0) | (cs.pps->getLoopFilterAcrossTilesEnabledFlag() ? VTMGPU_UNITS_ACROSS_TILES : 0) |
            (cs.sps->getPLTMode() ? VTMGPU_UNITS_PLT : 0);
  if (F.anyInter)
  {
    // the motion field as the decoder keeps it, before setRefinedMotionField (DecLib.cpp:579-580)
    const CMotionBuf mb = const_cast<const CodingStructure&>(cs).getMotionBuf();
    p.motion = mb.buf;
    p.motion_elem_bytes = (int)sizeof(MotionInfo);
    p.motion_pitch = (int)mb.stride;
    p.off_mv0 = (int)offsetof(MotionInfo, mv);
    p.off_mv1 = (int)(offsetof(MotionInfo, mv) + sizeof(Mv));
    p.off_ref0 = (int)offsetof(MotionInfo, refIdx);
    p.off_ref1 = (int)(offsetof(MotionInfo, refIdx) + sizeof(int16_t));
    static_assert(sizeof(Mv) == 8 && offsetof(Mv, hor) == 0 && offsetof(Mv, ver) == 4, "Mv layout");
  }
}

void deriveFromUnits(const FlatUnits& in, int width, int height, int chromaFormat, int bdLuma, int bdChroma, int ctuSize, FlatDeblock& out)
{
  vtmgpu_derive::Ctx D{};
  D.cus = in.cus; D.tus = in.tus; D.slices = in.slices.data();
  D.tuL = in.tuLuma; D.tuC = in.tuChroma;
  D.motion = static_cast<const unsigned char*>(in.p.motion);
  D.miBytes = in.p.motion_elem_bytes; D.miPitch = in.p.motion_pitch;
  D.offMv0 = in.p.off_mv0; D.offMv1 = in.p.off_mv1; D.offRef0 = in.p.off_ref0; D.offRef1 = in.p.off_ref1;
  D.w4 = width / 4; D.h4 = height / 4;
  D.sx = (chromaFormat == 1 || chromaFormat == 2) ? 1 : 0; D.sy = chromaFormat == 1 ? 1 : 0; D.chroma = chromaFormat != 0;
  D.bdL = bdLuma; D.bdC = bdChroma;
  D.ctuLog2 = ctuSize == 128 ? 7 : (ctuSize == 64 ? 6 : 5);
  D.flags = in.p.flags;
  D.ladf = in.hasLadf;
  D.nvb[0] = in.vb.num_ver; D.nvb[1] = in.vb.num_hor;
  for (int i = 0; i < 3; i++) { D.vb[0][i] = in.vb.pos_x[i]; D.vb[1][i] = in.vb.pos_y[i]; }
  D.tcTable = kTc; D.betaTable = kBeta;
  const int gx = 8 << D.sx, gy = 8 << D.sy;
  out.width = width; out.height = height; out.sx = D.sx; out.sy = D.sy;
  const size_t nL = (size_t)D.w4 * D.h4;
  for (int d = 0; d < 2; d++) out.luma[d].assign(nL, 0);
  out.chroma[0].assign(D.chroma ? (size_t)((width + gx - 1) / gx) * D.h4 : 0, 0);
  out.chroma[1].assign(D.chroma ? (size_t)((height + gy - 1) / gy) * D.w4 : 0, 0);
  for (int dir = 0; dir < 2; dir++)
    for (int y = 0; y < D.h4; y++)
      for (int x = 0; x < D.w4; x++)
      {
        uint32_t lr; uint64_t cr; bool slot;
        vtmgpu_derive::deriveUnit(D, x, y, dir, lr, cr, slot);
        out.luma[dir][(size_t)y * D.w4 + x] = lr;
        if (slot)
        {
          if (dir == 0) out.chroma[0][(size_t)y * ((width + gx - 1) / gx) + (4 * x) / gx] = cr;
          else          out.chroma[1][(size_t)((4 * y) / gy) * D.w4 + x] = cr;
        }
      }
}

// ---------------------------------------------------------------------------------------------------------
// SAO
// ---------------------------------------------------------------------------------------------------------
void flattenSao(CodingStructure& cs, const SAOBlkParam* blk, int log2ScaleLuma, int log2ScaleChroma, FlatSao& out)
{
  const PreCalcValues& pcv = *cs.pcv;
  out.vb = flattenVirtualBoundaries(cs);
  out.widthInCtus = pcv.widthInCtus;
  out.numComps = getNumberValidComponents(pcv.chrFormat);
  out.log2ScaleLuma = log2ScaleLuma;
  out.log2ScaleChroma = log2ScaleChroma;
  out.ctu.assign(pcv.sizeInCtus, vtmgpu_sao_ctu{});
  const bool acrossSlices = cs.pps->getLoopFilterAcrossSlicesEnabledFlag(), acrossTiles = cs.pps->getLoopFilterAcrossTilesEnabledFlag();
  static const int nbDx[8] = { -1, 1, 0, 0, -1, 1, -1, 1 }, nbDy[8] = { 0, 0, -1, 1, -1, -1, 1, 1 };   // bit order of VTMGPU_AVAIL_*
  for (int a = 0; a < (int)pcv.sizeInCtus; a++)
  {
    vtmgpu_sao_ctu& o = out.ctu[a];
    const int cx = a % pcv.widthInCtus, cy = a / pcv.widthInCtus;
    const Position pos(cx * pcv.maxCUWidth, cy * pcv.maxCUHeight);
    for (int c = 0; c < out.numComps; c++)
    {
      const SAOOffset& s = blk[a][c];
      o.comp[c].mode = (int8_t)s.modeIdc;
      o.comp[c].type = (int8_t)s.typeIdc;
      o.comp[c].aux  = (int8_t)s.typeAuxInfo;
      for (int k = 0; k < 32; k++) o.comp[c].offset[k] = (int16_t)s.offset[k];
    }
    // deriveLoopFilterBoundaryAvailibility (SampleAdaptiveOffset.cpp:668-729)
    const CodingUnit* cur = cs.getCU(pos, CH_L);
    for (int k = 0; k < 8; k++)
    {
      const CodingUnit* nb = cs.getCU(pos.offset(nbDx[k] * (int)pcv.maxCUWidth, nbDy[k] * (int)pcv.maxCUHeight), CH_L);
      bool ok = nb != nullptr;
      if (ok && !acrossSlices) ok = CU::isSameSlice(*cur, *nb);
      if (ok && !acrossTiles)  ok = CU::isSameTile(*cur, *nb);
      if (ok) o.avail |= 1 << k;
    }
    // getMergeList (:173-227)
    o.merge_left_ok  = cx > 0 && cs.getCURestricted(pos.offset(-(int)pcv.maxCUWidth, 0), *cur, cur->chType) != nullptr;
    o.merge_above_ok = cy > 0 && cs.getCURestricted(pos.offset(0, -(int)pcv.maxCUHeight), *cur, cur->chType) != nullptr;
  }
}

void writeBackSao(const FlatSao& in, SAOBlkParam* blk)
{
  for (size_t a = 0; a < in.ctu.size(); a++)
    for (int c = 0; c < in.numComps; c++)
    {
      SAOOffset& s = blk[a][c];
      const vtmgpu_sao_offset& o = in.ctu[a].comp[c];
      s.modeIdc = SAOMode(o.mode);
      s.typeIdc = o.type;
      s.typeAuxInfo = o.aux;
      for (int k = 0; k < 32; k++) s.offset[k] = o.offset[k];
    }
}

// ---------------------------------------------------------------------------------------------------------
// ALF
// ---------------------------------------------------------------------------------------------------------
namespace
{
// what one slice signals for ALF: enable flags, the luma APS list, the chroma APS, the CC-ALF enable flags (reconstructCoeffAPSs :620-649)
void sliceAlf(Slice* sl, vtmgpu_alf_params& p, std::vector<vtmgpu_alf_luma_aps>& lumaAps, vtmgpu_alf_chroma_aps& chromaAps, bool& hasChromaAps)
{
  for (int c = 0; c < 3; c++) p.enabled[c] = sl->getTileGroupAlfEnabledFlag(ComponentID(c));
  for (int c = 0; c < 2; c++) p.ccalf_enabled[c] = sl->m_ccAlfFilterParam.ccAlfFilterEnabled[c];
  APS** apss = sl->getAlfAPSs();
  lumaAps.clear();
  hasChromaAps = false;
  p.num_luma_aps = 0;
  if (!(p.enabled[0] || p.enabled[1] || p.enabled[2])) return;
  const std::vector<int> ids = sl->getTileGroupApsIdLuma();
  p.num_luma_aps = sl->getTileGroupNumAps();
  for (int i = 0; i < p.num_luma_aps; i++)
  {
    APS* aps = apss[ids[i]];
    CHECK(aps == nullptr, "invalid APS");
    const AlfParam& ap = aps->getAlfAPSParam();
    vtmgpu_alf_luma_aps f{};
    f.num_filters = ap.numLumaFilters;
    f.nonlinear = ap.nonLinearFlag[CHANNEL_TYPE_LUMA];
    for (int k = 0; k < VTMGPU_ALF_CLASSES; k++)
    {
      f.delta_idx[k] = ap.filterCoeffDeltaIdx[k];
      for (int j = 0; j < VTMGPU_ALF_LUMA_COEFF; j++)
      {
        f.coeff[k][j]    = ap.lumaCoeff[k * MAX_NUM_ALF_LUMA_COEFF + j];
        f.clip_idx[k][j] = ap.lumaClipp[k * MAX_NUM_ALF_LUMA_COEFF + j];
      }
    }
    lumaAps.push_back(f);
  }
  if (p.enabled[1] || p.enabled[2])
  {
    APS* aps = apss[sl->getTileGroupApsIdChroma()];
    CHECK(aps == nullptr, "invalid chroma APS");
    const AlfParam& ap = aps->getAlfAPSParam();
    hasChromaAps = true;
    chromaAps = vtmgpu_alf_chroma_aps();
    chromaAps.num_alts = ap.numAlternativesChroma;
    chromaAps.nonlinear = ap.nonLinearFlag[CHANNEL_TYPE_CHROMA];
    for (int k = 0; k < VTMGPU_ALF_MAX_ALTS; k++)
      for (int j = 0; j < VTMGPU_ALF_CHROMA_COEFF; j++)
      {
        chromaAps.coeff[k][j]    = ap.chromaCoeff[k][j];
        chromaAps.clip_idx[k][j] = ap.chromaClipp[k][j];
      }
  }
}

bool sameSliceAlf(const vtmgpu_alf_params& a, const std::vector<vtmgpu_alf_luma_aps>& la, const vtmgpu_alf_chroma_aps& ca, bool hca,
                  const vtmgpu_alf_params& b, const std::vector<vtmgpu_alf_luma_aps>& lb, const vtmgpu_alf_chroma_aps& cb, bool hcb)
{
  if (memcmp(a.enabled, b.enabled, sizeof(a.enabled)) || memcmp(a.ccalf_enabled, b.ccalf_enabled, sizeof(a.ccalf_enabled))) return false;
  if (la.size() != lb.size() || hca != hcb) return false;
  if (!la.empty() && memcmp(la.data(), lb.data(), la.size() * sizeof(la[0]))) return false;
  return !hca || memcmp(&ca, &cb, sizeof(ca)) == 0;
}
}   // namespace

void flattenAlf(CodingStructure& cs, const CcAlfFilterParam& cc, uint8_t* const ccControl[2], FlatAlf& out)
{
  const PreCalcValues& pcv = *cs.pcv;
  const int n = pcv.sizeInCtus;
  out = FlatAlf();
  out.p.num_ctus = n;
  // The first slice's parameters go down as the picture's set; every further slice whose ALF data differ (ALFProcess reloads the
  // APSs at each slice change and tests the CTU's own slice, AdaptiveLoopFilter.cpp:429-441, :451, :532) becomes a further set with
  // a per-CTU slice index.  The CC-ALF COEFFICIENTS stay per picture: the reference's ALF object holds one copy (DecLib.cpp:589).
  Slice* first = cs.getCU(Position(0, 0), CH_L)->slice;
  sliceAlf(first, out.p, out.lumaAps, out.chromaAps, out.hasChromaAps);
  {
    std::vector<Slice*> seen(1, first);
    std::vector<int> setOf(1, 0);                      // slice object -> parameter set (0 = the first slice's)
    out.ctuSlice.assign(n, 0);
    for (int a = 0; a < n; a++)
    {
      const Position pos((a % pcv.widthInCtus) * pcv.maxCUWidth, (a / pcv.widthInCtus) * pcv.maxCUHeight);
      Slice* s = cs.getCU(pos, CH_L)->slice;
      size_t k = 0;
      while (k < seen.size() && seen[k] != s) k++;
      if (k == seen.size())
      {
        FlatAlf::SliceSet t;
        sliceAlf(s, t.p, t.lumaAps, t.chromaAps, t.hasChromaAps);
        int idx = -1;
        if (sameSliceAlf(out.p, out.lumaAps, out.chromaAps, out.hasChromaAps, t.p, t.lumaAps, t.chromaAps, t.hasChromaAps)) idx = 0;
        for (size_t m = 0; m < out.more.size() && idx < 0; m++)
          if (sameSliceAlf(out.more[m].p, out.more[m].lumaAps, out.more[m].chromaAps, out.more[m].hasChromaAps, t.p, t.lumaAps, t.chromaAps, t.hasChromaAps)) idx = (int)m + 1;
        if (idx < 0)
        {
          CHECK(out.more.size() >= 254, "vtmgpu shim: too many slices with different ALF parameters in one picture");
          out.more.push_back(t);
          idx = (int)out.more.size();
        }
        seen.push_back(s);
        setOf.push_back(idx);
      }
      out.ctuSlice[a] = (uint8_t)setOf[k];
    }
    if (out.more.empty()) out.ctuSlice.clear();
  }
  cs.slice = cs.getCU(Position(((n - 1) % pcv.widthInCtus) * pcv.maxCUWidth, ((n - 1) / pcv.widthInCtus) * pcv.maxCUHeight), CH_L)->slice;
  // the clip / pad path of ALFProcess (:452-555) is entered when a CTU touches a slice or tile boundary that must not be
  // crossed: the per-CTU flags of isCrossedByVirtualBoundaries (:79-202).  Signalled virtual boundaries go down as positions
  // (they may lie inside a CTU); one that coincides with a CTU edge suppresses the slice/tile test and the corner pad of that side.
  out.vb = flattenVirtualBoundaries(cs);
  const bool acrossSlices = cs.pps->getLoopFilterAcrossSlicesEnabledFlag(), acrossTiles = cs.pps->getLoopFilterAcrossTilesEnabledFlag();
  if (!acrossSlices || !acrossTiles)
  {
    const int wc = pcv.widthInCtus, hc = pcv.heightInCtus;
    auto ctuCU = [&](int cx, int cy) { return cs.getCU(Position(cx * pcv.maxCUWidth, cy * pcv.maxCUHeight), CH_L); };
    auto blocked = [&](const CodingUnit* a, const CodingUnit* b) {
      return (!acrossSlices && !CU::isSameSlice(*a, *b)) || (!acrossTiles && !CU::isSameTile(*a, *b));
    };
    bool any = false;
    out.ctuClip.assign(n, 0);
    for (int a = 0; a < n; a++)
    {
      const int cx = a % wc, cy = a / wc;
      const CodingUnit* cur = ctuCU(cx, cy);
      int f = 0;
      if (cy > 0 && blocked(cur, ctuCU(cx, cy - 1))) f |= VTMGPU_ALF_CLIP_TOP;
      if (cy + 1 < hc && blocked(cur, ctuCU(cx, cy + 1))) f |= VTMGPU_ALF_CLIP_BOTTOM;
      if (cx > 0 && blocked(cur, ctuCU(cx - 1, cy))) f |= VTMGPU_ALF_CLIP_LEFT;
      if (cx + 1 < wc && blocked(cur, ctuCU(cx + 1, cy))) f |= VTMGPU_ALF_CLIP_RIGHT;
      // raster-scan slices: the diagonal neighbour may belong to another slice although the two adjacent ones do not (:178-200)
      if (!(f & (VTMGPU_ALF_CLIP_TOP | VTMGPU_ALF_CLIP_LEFT)) && cx > 0 && cy > 0 && !acrossSlices && !CU::isSameSlice(*cur, *ctuCU(cx - 1, cy - 1)))
        f |= VTMGPU_ALF_PAD_TL;
      if (!(f & (VTMGPU_ALF_CLIP_BOTTOM | VTMGPU_ALF_CLIP_RIGHT)) && cx + 1 < wc && cy + 1 < hc && !acrossSlices && !CU::isSameSlice(*cur, *ctuCU(cx + 1, cy + 1)))
        f |= VTMGPU_ALF_PAD_BR;
      out.ctuClip[a] = (uint8_t)f;
      any |= f != 0;
    }
    if (!any) out.ctuClip.clear();
  }

  for (int c = 0; c < 3; c++) out.ctuEnable[c].assign(cs.picture->getAlfCtuEnableFlag(c), cs.picture->getAlfCtuEnableFlag(c) + n);
  out.filterIdx.assign(cs.picture->getAlfCtbFilterIndex(), cs.picture->getAlfCtbFilterIndex() + n);
  for (int c = 0; c < 2; c++)
  {
    out.ctuAlt[c].assign(cs.picture->getAlfCtuAlternativeData(c + 1), cs.picture->getAlfCtuAlternativeData(c + 1) + n);
    bool anyCc = out.p.ccalf_enabled[c] != 0;
    for (const FlatAlf::SliceSet& t : out.more) anyCc |= t.p.ccalf_enabled[c] != 0;
    out.ccIdc[c].assign(n, 0);
    if (anyCc && ccControl[c]) out.ccIdc[c].assign(ccControl[c], ccControl[c] + n);
    for (int f = 0; f < VTMGPU_CCALF_MAX_FILTERS; f++)
      for (int j = 0; j < VTMGPU_CCALF_COEFF; j++) out.p.ccalf_coeff[c][f][j] = cc.ccAlfCoeff[c][f][j];
  }
}

}   // namespace vtmshim
