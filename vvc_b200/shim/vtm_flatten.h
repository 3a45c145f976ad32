// vtm_flatten.h -- host side of the drop-in boundary: turns the decoder's block metadata
// (CodingStructure / CodingUnit / TransformUnit / MotionInfo, Slice / APS) into the POD side information of
// include/vtmgpu.h.  Compiled against the reference's headers; nothing here touches sample arithmetic.
#pragma once

#include <cstdint>
#include <vector>

#include "vtmgpu.h"

class CodingStructure;
struct SAOBlkParam;
struct CcAlfFilterParam;

namespace vtmshim
{

struct FlatDeblock
{
  int width = 0, height = 0, sx = 0, sy = 0;
  std::vector<uint32_t> luma[2];
  std::vector<uint64_t> chroma[2];
  bool hasLadf = false;             // SPS LADF: the luma records carry QPs, see vtmgpu_ladf
  vtmgpu_ladf ladf{};
  vtmgpu_deblock_params view() const;
  // the same records as lists of the active units (vtmgpu_deblock_sparse), appended by the CU walk; listsValid is false when a
  // unit was emitted twice (the dense arrays, where the later record wins, are authoritative then)
  struct Lists
  {
    std::vector<vtmgpu_dbf_luma_entry>   luma[2];
    std::vector<vtmgpu_dbf_chroma_entry> chroma[2];
    bool twice = false;
  };
  Lists lists;
  bool listsValid = false;
  vtmgpu_deblock_sparse sparseView() const;
  bool listsMatchDense() const;       // self-check used by the capture mode
};

// picture-header virtual boundaries (empty when the flag is off)
vtmgpu_virtual_boundaries flattenVirtualBoundaries(const CodingStructure& cs);

struct FlatSao
{
  std::vector<vtmgpu_sao_ctu> ctu;   // raw (as parsed) until reconstruct() is called
  int widthInCtus = 0, numComps = 3, log2ScaleLuma = 0, log2ScaleChroma = 0;
  int enabledMask = 0;               // m_picSAOEnabled after reconstruct
  vtmgpu_virtual_boundaries vb{};
  vtmgpu_sao_params view() const;
};

struct FlatAlf
{
  vtmgpu_alf_params p{};
  std::vector<vtmgpu_alf_luma_aps> lumaAps;
  vtmgpu_alf_chroma_aps chromaAps{};
  bool hasChromaAps = false;
  std::vector<uint8_t> ctuEnable[3], ctuAlt[2], ccIdc[2];
  vtmgpu_virtual_boundaries vb{};
  std::vector<uint8_t> ctuClip;      // VTMGPU_ALF_CLIP_* / PAD_* per CTU; empty = no partition boundary restricts the filter
  std::vector<int16_t> filterIdx;
  const vtmgpu_alf_params* view();
  // slices of the picture whose ALF data differ from the first slice's (ALFProcess reloads the APSs whenever the CTU's slice changes,
  // AdaptiveLoopFilter.cpp:429-441): `p` describes the first slice, more[k - 1] the k-th further one, ctuSlice the index per CTU.
  // Empty for pictures whose slices agree (every stream the reference encoder writes) -> the single-set entry point.
  struct SliceSet
  {
    vtmgpu_alf_params p{};
    std::vector<vtmgpu_alf_luma_aps> lumaAps;
    vtmgpu_alf_chroma_aps chromaAps{};
    bool hasChromaAps = false;
  };
  std::vector<SliceSet> more;
  std::vector<uint8_t> ctuSlice;
  std::vector<vtmgpu_alf_params> sliceViews;       // view() of p followed by the further sets, as vtmgpu_set_alf_slices reads them
  const vtmgpu_alf_params* slicesView();
};

// The block structure of a picture flattened for the DEVICE-side derivation of the deblocking records (SURVEY 8f n1,
// vtmgpu_deblock_units in include/vtmgpu.h): one entry per CU and TU, per 4x4 unit the transform unit that covers it in the luma and in
// the chroma channel, the slices' deblocking offsets / reference picture tables; the motion field is handed over as the decoder keeps it.
struct FlatUnits
{
  // one page-aligned block per picture geometry (the shim page-locks it): CU table | TU table | luma-channel map | chroma-channel map
  void*  block = nullptr;
  size_t blockBytes = 0;
  bool   blockIsNew = false;        // set when flattenUnits (re)allocated the block
  vtmgpu_dbf_cu* cus = nullptr;
  vtmgpu_dbf_tu* tus = nullptr;
  uint32_t *tuLuma = nullptr, *tuChroma = nullptr;
  size_t capCus = 0, capTus = 0, numCus = 0, numTus = 0, units = 0;
  std::vector<vtmgpu_dbf_slice> slices;
  ~FlatUnits();
  bool hasLadf = false;
  vtmgpu_ladf ladf{};
  vtmgpu_virtual_boundaries vb{};
  vtmgpu_deblock_units p{};
  bool supported = true;            // false: a structure the unit model does not describe (a CU with several PUs): the CU walk derives this picture
  const vtmgpu_deblock_units* view();
};
void flattenUnits(CodingStructure& cs, FlatUnits& out);
// the derivation of include/vtmgpu_derive.h run on the host over every unit, into dense record arrays (self check against the CU walk)
void deriveFromUnits(const FlatUnits& in, int width, int height, int chromaFormat, int bdLuma, int bdChroma, int ctuSize, FlatDeblock& out);

// LoopFilter::xDeblockCU (LoopFilter.cpp:261-408) and everything it calls except the sample filters,
// re-stated to EMIT one record per edge segment instead of filtering (records: include/vtmgpu.h).
void deriveDeblockRecords(CodingStructure& cs, FlatDeblock& out);

// SAOBlkParam[] (TypeDef.h:938-963) + neighbour availability (SampleAdaptiveOffset.cpp:668-729) + merge
// candidate availability (getMergeList, :173-227)
void flattenSao(CodingStructure& cs, const SAOBlkParam* blk, int log2ScaleLuma, int log2ScaleChroma, FlatSao& out);
// writes the reconstructed parameters back (SAOProcess side effect, SampleAdaptiveOffset.cpp:247,255)
void writeBackSao(const FlatSao& in, SAOBlkParam* blk);

// slice / APS / per-CTU ALF control data read by ALFProcess (AdaptiveLoopFilter.cpp:393-456, :620-649)
void flattenAlf(CodingStructure& cs, const CcAlfFilterParam& cc, uint8_t* const ccControl[2], FlatAlf& out);

}   // namespace vtmshim
