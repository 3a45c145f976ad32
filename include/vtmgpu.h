/*
 * vtmgpu.h -- C ABI of libvtmgpu: the B200-native (sm_100a CUDA) implementation of the VVC
 * decoder in-loop filter chain  deblocking -> SAO -> ALF / CC-ALF.
 *
 * This is the drop-in boundary of SURVEY.md section 8(b): plain pointers and sizes, no C++/VTM/torch
 * types.  The reference interfaces each entry point stands behind (paths relative to the reference
 * tree, VTM 7.3+ snapshot):
 *
 *   vtmgpu_create / vtmgpu_destroy      LoopFilter::create/destroy            CommonLib/LoopFilter.cpp:111,130
 *                                        SampleAdaptiveOffset::create/destroy  CommonLib/SampleAdaptiveOffset.cpp:127,143
 *                                        AdaptiveLoopFilter::create/destroy    CommonLib/AdaptiveLoopFilter.cpp:715,816
 *   vtmgpu_upload / vtmgpu_download     Picture::getRecoBuf() planes          CommonLib/Picture.cpp:317  (PelStorage, int16 Pel)
 *   vtmgpu_set_deblock[_sparse] + vtmgpu_deblock   LoopFilter::loopFilterPic  CommonLib/LoopFilter.cpp:145
 *                                        (edge filtering xEdgeFilterLuma :844, xEdgeFilterChroma :1087; the per-4x4
 *                                        bS / tc / beta / filter-length DERIVATION :261-812 stays on the host and arrives
 *                                        here as packed segment records -- picture-sized arrays or lists of the active units;
 *                                        with LADF, deriveLADFShift :815, the records carry QPs and tc / beta are derived here)
 *   vtmgpu_sao_reconstruct              SampleAdaptiveOffset::xReconstructBlkSAOParams   SampleAdaptiveOffset.cpp:266
 *   vtmgpu_set_sao + vtmgpu_sao         SampleAdaptiveOffset::SAOProcess      CommonLib/SampleAdaptiveOffset.cpp:618
 *   vtmgpu_set_alf + vtmgpu_alf         AdaptiveLoopFilter::ALFProcess        CommonLib/AdaptiveLoopFilter.cpp:393
 *                                        (incl. reconstructCoeffAPSs :620, deriveClassificationBlk :873, filterBlk :1084,
 *                                        filterBlkCcAlf :1327, the clip / pad path at slice, tile and virtual boundaries :452-555)
 *   vtmgpu_filter                       DecLib::executeLoopFilters            DecoderLib/DecLib.cpp:560  (whole chain, batched)
 *
 * Error convention: every call returns 0 on success, non-zero on failure; vtmgpu_last_error() gives the text
 * (the reference throws Exception via THROW/CHECK, TypeDef.h:1152 -- the C++ shim converts non-zero to THROW).
 * There is NO CPU fallback: without a CUDA device vtmgpu_create fails.
 *
 * Threading: one ctx per decoder (DecLib owns one filter object of each kind, DecLib.h:97-99); a ctx is not
 * thread-safe; calls are synchronous unless the name ends in _async.
 *
 * Pictures live in "slots" (0 .. capacity-1) so that a batch of independent pictures can be filtered by one
 * launch sequence (picture-parallel replay); the decoder drop-in uses slot 0.
 */
#ifndef VTMGPU_H
#define VTMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VTMGPU_ABI_VERSION 3

/* chroma_format values follow ChromaFormat (TypeDef.h): 0 = 4:0:0, 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4 */
typedef struct vtmgpu_seq_params
{
  int32_t width;             /* luma samples, multiple of 8 */
  int32_t height;            /* luma samples, multiple of 8 */
  int32_t chroma_format;
  int32_t bit_depth_luma;    /* 8..12 */
  int32_t bit_depth_chroma;  /* 8..12 */
  int32_t ctu_size;          /* 32, 64 or 128 (sps CTUSize) */
  int32_t capacity;          /* number of picture slots (>= 1) */
  int32_t device;            /* CUDA device ordinal */
} vtmgpu_seq_params;

/* ---------------------------------------------------------------------------------------------
 * Deblocking segment records.  One record describes one 4-luma-sample long
 * piece of one edge.  A record of 0 means "do not filter".
 *
 * luma  (uint32), one per 4x4 luma unit and direction; the record at unit (ux,uy) describes the edge on
 *        the LEFT side (dir 0 = EDGE_VER) or TOP side (dir 1 = EDGE_HOR) of that unit:
 *   bits  0-10  tc      (LoopFilter.cpp:974, already scaled to the bit depth; 0 = edge not filtered)
 *   bits 11-21  beta    (:975)
 *   bits 22-24  maxFilterLengthP (:948, after the cuP.affine clamp :953-960)   1,2,3,5,7
 *   bits 25-27  maxFilterLengthQ (:949)
 *   bit  28     bPartPNoFilter  (P side is palette coded, :1017)
 *   bit  29     bPartQNoFilter
 *   bit  30     horizontal edge on a CTU-row boundary: sidePisLarge forced false (:967-970)
 *   array: dbf_luma[dir][(y/4) * (width/4) + x/4]
 *
 * chroma (uint64), only for edges on the 8x8 chroma-sample grid (:1118-1126); one record covers the
 *        chroma samples belonging to one 4-luma-sample unit along the edge, both Cb and Cr:
 *   bits  0-10  tc  Cb   (0 = Cb not filtered)      bits 11-21  tc  Cr
 *   bits 22-32  beta Cb                              bits 33-43  beta Cr
 *   bit  44     largeBoundary (both filter lengths >= 3, :1202)
 *   bit  45     isChromaHorCTBBoundary (:1207)
 *   bit  46     bPartPNoFilter     bit 47  bPartQNoFilter
 *   dir 0 array: [(y/4) * cols0 + x/(8<<sx)],  cols0 = ceil(width  / (8<<sx))
 *   dir 1 array: [(y/(8<<sy)) * (width/4) + x/4], rows1 = ceil(height / (8<<sy))
 *   (x,y in luma samples; sx,sy = chroma subsampling shifts)
 * --------------------------------------------------------------------------------------------- */
#define VTMGPU_DBF_TC_BITS      11
#define VTMGPU_DBF_L_BETA_SHIFT 11
#define VTMGPU_DBF_L_LENP_SHIFT 22
#define VTMGPU_DBF_L_LENQ_SHIFT 25
#define VTMGPU_DBF_L_PNOFILT    (1u << 28)
#define VTMGPU_DBF_L_QNOFILT    (1u << 29)
#define VTMGPU_DBF_L_CTUROW     (1u << 30)
#define VTMGPU_DBF_C_TCCR_SHIFT   11
#define VTMGPU_DBF_C_BETACB_SHIFT 22
#define VTMGPU_DBF_C_BETACR_SHIFT 33
#define VTMGPU_DBF_C_LARGE      (1ull << 44)
#define VTMGPU_DBF_C_CTB        (1ull << 45)
#define VTMGPU_DBF_C_PNOFILT    (1ull << 46)
#define VTMGPU_DBF_C_QNOFILT    (1ull << 47)

/* LADF (luma adaptive deblocking filter QP offset, an SPS tool; deriveLADFShift, LoopFilter.cpp:815-841): the QP of a
 * luma edge segment gets an offset selected by the mean of four reconstructed samples next to the edge -- for horizontal
 * edges those are samples AFTER the vertical pass, which only the device has.  With ladf != NULL the luma records
 * therefore carry QPs instead of thresholds,
 *      tc   field = 128 + QP + 2 * (bS - 1) + 2 * slice_tc_offset_div2      (before the LADF offset, unclipped; :971)
 *      beta field = 128 + QP + 2 * slice_beta_offset_div2                    (:972)
 * and the kernel adds the offset and looks tc / beta up (tables :66-74, bit-depth scaling :974-975).  Chroma records are
 * unchanged (LADF is luma only). */
typedef struct vtmgpu_ladf
{
  int32_t num_intervals;       /* SPS getLadfNumIntervals(), 2..5                              */
  int32_t qp_offset[5];        /* getLadfQpOffset(k)                                           */
  int32_t lower_bound[5];      /* getLadfIntervalLowerBound(k), k >= 1 ([0] unused)            */
} vtmgpu_ladf;
#define VTMGPU_DBF_LADF_BIAS 128

typedef struct vtmgpu_deblock_params
{
  const uint32_t* luma[2];     /* [dir] width/4 * height/4 records                         */
  const uint64_t* chroma[2];   /* [dir] see above; NULL for 4:0:0                            */
  const vtmgpu_ladf* ladf;     /* NULL unless the SPS enables LADF                            */
} vtmgpu_deblock_params;

/* Sparse form of the same records: only the units that carry an edge to be filtered (typically ~10 % of the
 * units of an inter picture), as produced by a CU walk (LoopFilter::xDeblockCU, LoopFilter.cpp:261-408) that
 * appends instead of storing into a picture-sized array.  array a: 0 = luma dir 0, 1 = luma dir 1,
 * 2 = chroma dir 0, 3 = chroma dir 1; index = position in the corresponding dense array above.  Every index
 * may appear at most once per array; units that are not listed have no edge.  If the four lists sit in one buffer, in array
 * order, each starting on the next 16-byte boundary after the previous one, they are uploaded with a single copy. */
typedef struct vtmgpu_dbf_luma_entry   { uint32_t index; uint32_t rec; } vtmgpu_dbf_luma_entry;
typedef struct vtmgpu_dbf_chroma_entry { uint64_t rec; uint32_t index; uint32_t reserved; } vtmgpu_dbf_chroma_entry;
typedef struct vtmgpu_deblock_sparse
{
  const vtmgpu_dbf_luma_entry*   luma[2];
  const vtmgpu_dbf_chroma_entry* chroma[2];   /* ignored for 4:0:0 */
  uint32_t luma_count[2];
  uint32_t chroma_count[2];
  const vtmgpu_ladf* ladf;     /* as in vtmgpu_deblock_params */
} vtmgpu_deblock_sparse;

/* Third form: the block structure itself, flattened once per picture; the DEVICE then derives the records (edge
 * maps, filter lengths, boundary strengths, tc / beta of LoopFilter::xDeblockCU .. xGetBoundaryStrengthSingle, LoopFilter.cpp:261-812,
 * one thread per 4x4 unit and direction; include/vtmgpu_derive.h is the derivation, shared by the kernel and the host-side self check).
 * All positions in luma samples unless stated.  tu_luma / tu_chroma: for every 4x4 luma unit (raster order, width/4 per row) the index
 * of the transform unit that CodingStructure::getTU(pos, CHANNEL_TYPE_LUMA / CHROMA) returns for the unit's top-left sample. */
#define VTMGPU_CU_INTRA     0x0001   /* predMode == MODE_INTRA                                   */
#define VTMGPU_CU_IBC       0x0002   /* CU::isIBC                                                */
#define VTMGPU_CU_PLT       0x0004   /* CU::isPLT                                                */
#define VTMGPU_CU_AFFINE    0x0008
#define VTMGPU_CU_MVSUB     0x0010   /* affine or SbTMVP merge: sub-block edges every 8 samples  */
#define VTMGPU_CU_BDPCM     0x0020
#define VTMGPU_CU_BDPCM_C   0x0040
#define VTMGPU_CU_CIIP      0x0080
#define VTMGPU_CU_ISP       0x0100
#define VTMGPU_CU_HAS_LUMA  0x0400   /* blocks[COMPONENT_Y].valid()                              */
#define VTMGPU_CU_HAS_CHROMA 0x0800  /* blocks[COMPONENT_Cb].valid()                             */
#define VTMGPU_CU_EN_LEFT   0x1000   /* xSetLoopfilterParam (LoopFilter.cpp:656-672): left CU edge, top CU edge, internal edges */
#define VTMGPU_CU_EN_TOP    0x2000
#define VTMGPU_CU_EN_INT    0x4000
typedef struct vtmgpu_dbf_cu
{
  uint16_t x, y;          /* area of the CU (a chroma-only CU: its chroma block scaled to luma samples) */
  uint8_t  w4, h4;        /* size in units of 4 luma samples                                          */
  int8_t   qp;            /* CodingUnit::qp                                                            */
  uint8_t  slice;         /* index into slices[]                                                       */
  uint16_t tile;          /* CodingUnit::tileIdx                                                       */
  uint16_t flags;         /* VTMGPU_CU_*                                                               */
  uint32_t reserved;
} vtmgpu_dbf_cu;
#define VTMGPU_TU_CBF_Y   1
#define VTMGPU_TU_CBF_CB  2
#define VTMGPU_TU_CBF_CR  4
#define VTMGPU_TU_JOINT   8
typedef struct vtmgpu_dbf_tu
{
  uint16_t x, y;          /* luma block; w = 0: the TU has none                                        */
  uint8_t  w, h;          /* in samples (ISP sub-partitions may be 1 or 2 samples wide)                */
  uint8_t  cw, ch;        /* chroma block size in chroma samples; cw = 0: the TU has none              */
  uint16_t cx, cy;        /* chroma block position in chroma samples                                   */
  uint8_t  cbf;           /* VTMGPU_TU_*                                                               */
  int8_t   qp_cb, qp_cr;  /* QpParam(tu, comp).Qp(0) - qpBdOffset (LoopFilter.cpp:1213-1217)           */
  uint8_t  reserved;
  uint32_t cu;            /* index into cus[]                                                          */
  uint32_t tail;          /* luma blocks narrower (lower) than a unit -- ISP sub-partitions: the TU that holds the LAST column (row) of the
                             unit this one starts, i.e. what getTU returns for the sample next to the following edge; else the TU itself */
} vtmgpu_dbf_tu;
typedef struct vtmgpu_dbf_slice
{
  int8_t  tc_offset, beta_offset;      /* getDeblockingFilterTcOffsetDiv2() * 2, ...BetaOffsetDiv2() * 2 */
  uint8_t inter_b;                     /* isInterB()                                                      */
  uint8_t reserved;
  uint16_t independent_idx;            /* getIndependentSliceIdx() (CU::isSameSlice)                      */
  int16_t ref_pic[2][16];              /* [list][refIdx] -> picture id (any numbering that identifies the Picture object) */
} vtmgpu_dbf_slice;
#define VTMGPU_UNITS_ACROSS_SLICES 1   /* PPS loop_filter_across_slices_enabled_flag */
#define VTMGPU_UNITS_ACROSS_TILES  2
#define VTMGPU_UNITS_PLT           4   /* SPS getPLTMode() */
#define VTMGPU_UNITS_MOTION_PRELOADED 8  /* the motion field of this picture is already on the device (vtmgpu_upload_motion); `motion` is ignored */
typedef struct vtmgpu_deblock_units
{
  int32_t num_cus, num_tus, num_slices, flags;
  const vtmgpu_dbf_cu*    cus;
  const vtmgpu_dbf_tu*    tus;
  const vtmgpu_dbf_slice* slices;
  const uint32_t* tu_luma;       /* [units] */
  const uint32_t* tu_chroma;     /* [units]; NULL for 4:0:0 */
  /* the picture's motion field at 4x4 granularity as the decoder keeps it (MotionInfo, MotionInfo.h:101): base address, bytes per
   * element, elements per row, and the byte offsets of mv[0].hor (ver follows), mv[1].hor, refIdx[0], refIdx[1] (int16).  NULL for
   * pictures without inter CUs. */
  const void* motion;
  int32_t motion_elem_bytes, motion_pitch, off_mv0, off_mv1, off_ref0, off_ref1;
  const vtmgpu_ladf* ladf;                    /* as in vtmgpu_deblock_params */
  const struct vtmgpu_virtual_boundaries* vb; /* edges on a signalled virtual boundary are not filtered (xDeriveEdgefilterParam) */
} vtmgpu_deblock_units;

/* ---------------------------------------------------------------------------------------------
 * SAO (SAOOffset / SAOBlkParam, TypeDef.h:938-963; enums :706-748)
 * --------------------------------------------------------------------------------------------- */
enum { VTMGPU_SAO_MODE_OFF = 0, VTMGPU_SAO_MODE_NEW = 1, VTMGPU_SAO_MODE_MERGE = 2 };
enum { VTMGPU_SAO_EO_0 = 0, VTMGPU_SAO_EO_90 = 1, VTMGPU_SAO_EO_135 = 2, VTMGPU_SAO_EO_45 = 3, VTMGPU_SAO_BO = 4 };
enum { VTMGPU_SAO_MERGE_LEFT = 0, VTMGPU_SAO_MERGE_ABOVE = 1 };

typedef struct vtmgpu_sao_offset
{
  int8_t  mode;         /* VTMGPU_SAO_MODE_*                                            */
  int8_t  type;         /* NEW: VTMGPU_SAO_EO_* / BO;  MERGE: VTMGPU_SAO_MERGE_*         */
  int8_t  aux;          /* BO: first band (typeAuxInfo)                                   */
  int8_t  reserved;
  int16_t offset[32];   /* as parsed (CABACReader.cpp:318); after reconstruct: scaled    */
} vtmgpu_sao_offset;

/* neighbour-CTU availability bits for the EO classes (deriveLoopFilterBoundaryAvailibility, :668) */
#define VTMGPU_AVAIL_LEFT        0x01
#define VTMGPU_AVAIL_RIGHT       0x02
#define VTMGPU_AVAIL_ABOVE       0x04
#define VTMGPU_AVAIL_BELOW       0x08
#define VTMGPU_AVAIL_ABOVE_LEFT  0x10
#define VTMGPU_AVAIL_ABOVE_RIGHT 0x20
#define VTMGPU_AVAIL_BELOW_LEFT  0x40
#define VTMGPU_AVAIL_BELOW_RIGHT 0x80

typedef struct vtmgpu_sao_ctu
{
  vtmgpu_sao_offset comp[3];
  uint8_t avail;             /* VTMGPU_AVAIL_* of the 8 neighbouring CTUs                               */
  uint8_t merge_left_ok;     /* left  CTU usable as merge candidate (getMergeList, :173: same slice+tile) */
  uint8_t merge_above_ok;
  uint8_t reserved;
} vtmgpu_sao_ctu;

/* Virtual boundaries signalled in the picture header with loop_filter_across_virtual_boundaries_disabled_flag (a tool for
 * 360-degree video; positions are multiples of 8 luma samples, at least one CTU apart, strictly inside the picture).  Deblocking
 * drops the edges on them on the host (xDeriveEdgefilterParam, LoopFilter.cpp:435-452); SAO skips the edge-offset samples next
 * to a boundary its class looks across (isProcessDisabled, SampleAdaptiveOffset.h:96); ALF treats every part of a CTU between
 * boundaries as a separately padded block (ALFProcess :452-490). */
typedef struct vtmgpu_virtual_boundaries
{
  int32_t num_ver, num_hor;    /* 0..3 each                                                     */
  int32_t pos_x[3], pos_y[3];  /* luma samples                                                  */
} vtmgpu_virtual_boundaries;

typedef struct vtmgpu_sao_params
{
  const vtmgpu_sao_ctu* ctu;   /* [ctus], raster; ALREADY reconstructed (vtmgpu_sao_reconstruct)    */
  int32_t num_ctus;
  const vtmgpu_virtual_boundaries* vb;   /* NULL: none                                         */
} vtmgpu_sao_params;

/* ---------------------------------------------------------------------------------------------
 * ALF / CC-ALF (AlfParam / CcAlfFilterParam, AlfParameters.h:133-287)
 * --------------------------------------------------------------------------------------------- */
#define VTMGPU_ALF_CLASSES        25
#define VTMGPU_ALF_LUMA_COEFF     13
#define VTMGPU_ALF_CHROMA_COEFF   7
#define VTMGPU_ALF_MAX_APS        8
#define VTMGPU_ALF_MAX_ALTS       8
#define VTMGPU_ALF_FIXED_SETS     16
#define VTMGPU_CCALF_MAX_FILTERS  4
#define VTMGPU_CCALF_COEFF        8

typedef struct vtmgpu_alf_luma_aps   /* the luma part of one ALF APS as parsed (VLCReader.cpp:837) */
{
  int32_t num_filters;                                                  /* numLumaFilters          */
  int32_t nonlinear;                                                    /* nonLinearFlag[LUMA]     */
  int16_t delta_idx[VTMGPU_ALF_CLASSES];                                /* filterCoeffDeltaIdx     */
  int16_t coeff[VTMGPU_ALF_CLASSES][VTMGPU_ALF_LUMA_COEFF];             /* lumaCoeff               */
  int16_t clip_idx[VTMGPU_ALF_CLASSES][VTMGPU_ALF_LUMA_COEFF];          /* lumaClipp (index 0..3)  */
} vtmgpu_alf_luma_aps;

typedef struct vtmgpu_alf_chroma_aps
{
  int32_t num_alts;                                                     /* numAlternativesChroma   */
  int32_t nonlinear;                                                    /* nonLinearFlag[CHROMA]   */
  int16_t coeff[VTMGPU_ALF_MAX_ALTS][VTMGPU_ALF_CHROMA_COEFF];
  int16_t clip_idx[VTMGPU_ALF_MAX_ALTS][VTMGPU_ALF_CHROMA_COEFF];
} vtmgpu_alf_chroma_aps;

/* ALF at slice / tile boundaries with loop filtering across them disabled (ALFProcess :452-555): samples beyond a clipped
 * side of the CTU are replaced by the nearest sample inside (copy + extendBorderPel of the reference); PAD_TL / PAD_BR: the
 * top-left / bottom-right neighbour CTU belongs to another raster-scan slice although the top and left (bottom and right)
 * neighbours do not -- the corner is padded horizontally from the CTU's first / last column (padBorderPel, Buffer.h:571). */
#define VTMGPU_ALF_CLIP_TOP     1
#define VTMGPU_ALF_CLIP_BOTTOM  2
#define VTMGPU_ALF_CLIP_LEFT    4
#define VTMGPU_ALF_CLIP_RIGHT   8
#define VTMGPU_ALF_PAD_TL       16
#define VTMGPU_ALF_PAD_BR       32

typedef struct vtmgpu_alf_params
{
  int32_t enabled[3];                   /* slice getTileGroupAlfEnabledFlag(Y/Cb/Cr) (AdaptiveLoopFilter.cpp:429) */
  int32_t num_luma_aps;                 /* slice getTileGroupNumAps()                                          */
  const vtmgpu_alf_luma_aps*   luma_aps;     /* [num_luma_aps], in getTileGroupApsIdLuma() order               */
  const vtmgpu_alf_chroma_aps* chroma_aps;   /* APS getTileGroupApsIdChroma(), or NULL                         */
  const uint8_t* ctu_enable[3];         /* Picture::getAlfCtuEnableFlag(comp)  [ctus]                          */
  const int16_t* ctu_filter_idx;        /* Picture::getAlfCtbFilterIndex()     [ctus] (<16 fixed, else APS)    */
  const uint8_t* ctu_alt[2];            /* Picture::getAlfCtuAlternativeData(Cb/Cr) [ctus]                     */
  int32_t ccalf_enabled[2];             /* CcAlfFilterParam::ccAlfFilterEnabled                                */
  int16_t ccalf_coeff[2][VTMGPU_CCALF_MAX_FILTERS][VTMGPU_CCALF_COEFF];
  const uint8_t* ccalf_idc[2];          /* m_ccAlfFilterControl[comp-1] [ctus]; 0 = off, k = filter k-1        */
  int32_t num_ctus;
  const uint8_t* ctu_clip;              /* [ctus] VTMGPU_ALF_CLIP_* / PAD_*, or NULL: picture partition boundaries that the
                                           filter must not cross (isCrossedByVirtualBoundaries, AdaptiveLoopFilter.cpp:79-202) */
  const vtmgpu_virtual_boundaries* vb;  /* signalled virtual boundaries, or NULL                                             */
} vtmgpu_alf_params;

/* ---------------------------------------------------------------------------------------------
 * entry points
 * --------------------------------------------------------------------------------------------- */
typedef struct vtmgpu_ctx vtmgpu_ctx;

int          vtmgpu_abi_version(void);
int          vtmgpu_abi_sizeof(int which);   /* sizeof of ABI struct #which (0 seq, 1 deblock, 2 sao_offset, 3 sao_ctu, 4 sao_params,
                                                5 alf_luma_aps, 6 alf_chroma_aps, 7 alf_params, 8 deblock_sparse, 9 ladf, 10 virtual_boundaries) for binding self-checks */
const char*  vtmgpu_last_error(const vtmgpu_ctx* ctx);   /* ctx may be NULL: error of the last failed create */

int  vtmgpu_create(const vtmgpu_seq_params* seq, vtmgpu_ctx** out);
void vtmgpu_destroy(vtmgpu_ctx* ctx);

/* host planes (int16 samples, stride in samples) <-> device slot; plane[1], plane[2] ignored for 4:0:0 */
int vtmgpu_upload  (vtmgpu_ctx* ctx, int slot, const int16_t* const plane[3], const ptrdiff_t stride[3]);
int vtmgpu_download(vtmgpu_ctx* ctx, int slot, int16_t* const plane[3], const ptrdiff_t stride[3]);
/* the same, enqueued on the ctx stream without waiting (host memory must stay valid until vtmgpu_sync; truly
 * asynchronous only from/to page-locked host memory) */
int vtmgpu_upload_async  (vtmgpu_ctx* ctx, int slot, const int16_t* const plane[3], const ptrdiff_t stride[3]);
int vtmgpu_download_async(vtmgpu_ctx* ctx, int slot, int16_t* const plane[3], const ptrdiff_t stride[3]);
/* page-locks host memory the caller keeps transferring from / to (the decoder's picture buffers live as long as the sequence): uploads
 * and downloads then run as plain DMA instead of through the driver's staging buffers.  Thin wrappers over cudaHostRegister /
 * cudaHostUnregister so that a host program needs no CUDA headers; non-zero = not registered (the transfers still work, staged). */
int vtmgpu_host_register(void* ptr, size_t bytes);
int vtmgpu_host_unregister(void* ptr);
/* download + reference-picture border extension (Picture::extendPicBorder, CommonLib/Picture.cpp:737-772, no wrap-around): plane[k]
 * points at sample (0,0) of a host buffer with at least margin_luma >> (chroma shift) samples of room on every side of the picture;
 * the picture and its margins (every outside sample = the nearest picture sample) arrive with one copy per plane.  Synchronous. */
int vtmgpu_download_extended(vtmgpu_ctx* ctx, int slot, int16_t* const plane[3], const ptrdiff_t stride[3], int margin_luma);

/* per-picture side information (host pointers; copied before return, never retained) */
int vtmgpu_set_deblock(vtmgpu_ctx* ctx, int slot, const vtmgpu_deblock_params* p);   /* NULL = stage off */
/* the same without the staging copy: the record arrays are read by asynchronous copies on the ctx stream and must stay valid
 * (and should be page-locked) until vtmgpu_sync */
int vtmgpu_set_deblock_async(vtmgpu_ctx* ctx, int slot, const vtmgpu_deblock_params* p);
/* the records as lists (0.1-0.2 B of upload per luma pixel instead of 0.75): the lists are read by asynchronous copies on
 * the ctx stream and scattered into the record arrays on the device (the scatter is enqueued by the next stage call of the slot,
 * behind the upload of the SAO / ALF side information); page-locked lists must stay valid until vtmgpu_sync, pageable ones may
 * be reused on return */
int vtmgpu_set_deblock_sparse(vtmgpu_ctx* ctx, int slot, const vtmgpu_deblock_sparse* p);
/* the block structure; the tables and maps are read by asynchronous copies on the ctx stream (page-locked memory must stay valid until
 * vtmgpu_sync, pageable memory may be reused on return) and k_dbf_derive writes the record arrays on the device */
int vtmgpu_set_deblock_units(vtmgpu_ctx* ctx, int slot, const vtmgpu_deblock_units* p);
/* the motion field alone, ahead of the tables (it does not depend on the host's flattening, so a caller can overlap the two): same
 * parameters as in vtmgpu_deblock_units; enqueued on the ctx stream like an upload */
int vtmgpu_upload_motion(vtmgpu_ctx* ctx, int slot, const void* motion, int elem_bytes, int pitch, int off_mv0, int off_mv1, int off_ref0, int off_ref1);
/* test hook: the record arrays of a slot as the device holds them, laid out like the arrays of vtmgpu_deblock_params; the caller allocates them */
int vtmgpu_get_deblock_records(vtmgpu_ctx* ctx, int slot, uint32_t* const luma[2], uint64_t* const chroma[2]);
/* LMCS inverse luma mapping of the reconstruction (AreaBuf<Pel>::rspSignal with Reshape::getInvLUT(), CommonLib/Buffer.cpp:380-393,
 * called by DecLib::executeLoopFilters right before loopFilterPic, DecoderLib/DecLib.cpp:570-577): with a table set, the luma plane
 * uploaded to the slot is the RESHAPED-domain reconstruction and the first stage that reads it (deblocking and / or SAO) maps every
 * luma sample through inv_lut while it loads its tile -- no separate pass over the picture.  entries = 1 << bit_depth_luma.
 * NULL = off (the default).  The table stays with the slot until it is replaced or switched off. */
int vtmgpu_set_lmcs   (vtmgpu_ctx* ctx, int slot, const int16_t* inv_lut, int entries);
int vtmgpu_set_sao    (vtmgpu_ctx* ctx, int slot, const vtmgpu_sao_params* p);       /* NULL = stage off */
int vtmgpu_set_alf    (vtmgpu_ctx* ctx, int slot, const vtmgpu_alf_params* p);       /* NULL = stage off */
/* Pictures whose slices carry DIFFERENT ALF parameters: ALFProcess reloads the APS data whenever the CTU's slice changes
 * (AdaptiveLoopFilter.cpp:429-441) and tests the slice's own enable flags per CTU (:429, :451, :532).
 *   slices[k]      what slice k signals: enabled[], num_luma_aps / luma_aps, chroma_aps, ccalf_enabled[]
 *   ctu_slice[ctu] index of the slice the CTU belongs to
 * Everything that is per picture in the reference is taken from slices[0]: the per-CTU arrays, ctu_clip, vb, num_ctus and --
 * like the reference, whose ALF object holds ONE copy of them (DecLib.cpp:589, AdaptiveLoopFilter.cpp:537,610) --
 * ccalf_coeff.  ctu_filter_idx >= 16 counts in the CTU's own slice's luma_aps list.  Slices with equal parameters share their
 * tables; at most VTMGPU_ALF_MAX_SLICE_SETS distinct parameter sets and VTMGPU_ALF_MAX_APS distinct luma APSs per picture. */
#define VTMGPU_ALF_MAX_SLICE_SETS 8
int vtmgpu_set_alf_slices(vtmgpu_ctx* ctx, int slot, int num_slices, const vtmgpu_alf_params* slices, const uint8_t* ctu_slice);

/* host-only helper: resolves MERGE / scales NEW offsets in place, raster order
 * (xReconstructBlkSAOParams, SampleAdaptiveOffset.cpp:266-290).  Returns <0 on error, else a 3-bit mask
 * of the components that have any CTU with SAO on (m_picSAOEnabled). */
int vtmgpu_sao_reconstruct(vtmgpu_sao_ctu* ctu, int num_ctus, int width_in_ctus, int num_comps,
                           int log2_offset_scale_luma, int log2_offset_scale_chroma);

/* stages on slots [first, first+count): synchronous, in place from the caller's point of view */
int vtmgpu_deblock(vtmgpu_ctx* ctx, int first, int count);   /* loopFilterPic                              */
int vtmgpu_sao    (vtmgpu_ctx* ctx, int first, int count);   /* SAOProcess (no-op where SAO is off)        */
int vtmgpu_alf    (vtmgpu_ctx* ctx, int first, int count);   /* ALFProcess                                 */
int vtmgpu_deblock_sao(vtmgpu_ctx* ctx, int first, int count);   /* loopFilterPic + SAOProcess in ONE pass (one kernel)   */
int vtmgpu_sao_alf(vtmgpu_ctx* ctx, int first, int count);   /* SAOProcess + ALFProcess back to back, one sync */
/* whole chain DBF -> SAO -> ALF in two kernels (deblocking + SAO, ALF + CC-ALF); same result as the three calls above */
int vtmgpu_filter (vtmgpu_ctx* ctx, int first, int count);

/* ---------------------------------------------------------------------------------------------
 * band mode (one large picture split into CTU-row bands over several contexts / GPUs).
 * Every context is created for the FULL picture geometry, so all CTU / virtual-boundary / picture-border rules keep
 * their absolute positions; it uploads and filters only its band:
 *   vtmgpu_set_rows(ctx, y0, y1)          stage calls filter luma rows [y0, y1) only (multiples of 128; y1 may be the height)
 *   vtmgpu_upload_rows(.., y0-16, y1+16)  the band plus the 16 luma rows (8 chroma rows at 4:2:0) the deblocking of its border tiles
 *                                         reads: 8 for the filters + 8 so that the chroma rows are whole (vvc_b200/bands.py UPLOAD_MARGIN)
 *   vtmgpu_deblock / vtmgpu_sao           exact for the band (they read pre-filter samples only)
 *   vtmgpu_export_rows / import_rows      4 rows of SAO output on each side of a band border travel to the neighbour
 *                                         (dense device buffers; the caller moves them with NCCL send/recv over NVLink)
 *   vtmgpu_alf                            exact for the band (ALF reads 3 rows, the Laplacian classifier 3 rows across)
 *   vtmgpu_download_rows(.., y0, y1)
 * plane[k] / stride[k] in the *_rows calls describe the FULL host picture (row 0), as in vtmgpu_upload.
 * --------------------------------------------------------------------------------------------- */
int vtmgpu_set_rows(vtmgpu_ctx* ctx, int y_begin, int y_end);
/* Redirects all work of the ctx to the caller's CUDA stream (cudaStream_t; NULL = back to the ctx's own stream) so that it
 * orders with the caller's other work on that stream (e.g. NCCL send/recv of the halo rows) without host synchronisation.
 * async_stages != 0: stage calls and export/import only ENQUEUE; the caller synchronises (vtmgpu_sync). */
int vtmgpu_set_stream(vtmgpu_ctx* ctx, void* cuda_stream, int async_stages);
int vtmgpu_upload_rows  (vtmgpu_ctx* ctx, int slot, const int16_t* const plane[3], const ptrdiff_t stride[3], int y_begin, int y_end);
int vtmgpu_download_rows(vtmgpu_ctx* ctx, int slot, int16_t* const plane[3], const ptrdiff_t stride[3], int y_begin, int y_end);
/* rows [y0, y0+nrows) (in samples of component comp) of the slot's current state <-> dense device memory (width * nrows int16) */
int vtmgpu_export_rows(vtmgpu_ctx* ctx, int slot, int comp, int y0, int nrows, void* dev_dst);
int vtmgpu_import_rows(vtmgpu_ctx* ctx, int slot, int comp, int y0, int nrows, const void* dev_src);
/* the same for all components at once: nrows rows of every plane starting at row y[comp], packed luma, Cb, Cr
 * (one buffer = one message per band border and direction) */
int vtmgpu_export_halo(vtmgpu_ctx* ctx, int slot, const int y[3], int nrows, void* dev_dst);
int vtmgpu_import_halo(vtmgpu_ctx* ctx, int slot, const int y[3], int nrows, const void* dev_src);

/* ---------------------------------------------------------------------------------------------
 * decoded-picture hash ON THE DEVICE: the hash of the decoded picture hash SEI over the CURRENT state of the
 * slots' planes, so that a filtered picture can be verified (or its SEI written) without leaving HBM.
 *   VTMGPU_HASH_MD5       calcMD5       CommonLib/PicYuvMD5.cpp:188   16 bytes per component
 *   VTMGPU_HASH_CRC       calcCRC       CommonLib/PicYuvMD5.cpp:130    2 bytes per component
 *   VTMGPU_HASH_CHECKSUM  calcChecksum  CommonLib/PicYuvMD5.cpp:169    4 bytes per component
 * digest receives, per slot, the components' digests one after the other exactly as PictureHash::hash holds them
 * (48 bytes are reserved per slot); *bytes_per_component returns 16 / 2 / 4.  Called from DecLib.cpp:699-708 in the reference
 * (calcAndPrintHashStatus).  CRC and checksum are parallel reductions; MD5 is one serial chain per component (all components and
 * slots of the call run in parallel, so it pays for batches -- for a single picture the host is faster).
 * --------------------------------------------------------------------------------------------- */
#define VTMGPU_HASH_MD5      1
#define VTMGPU_HASH_CRC      2
#define VTMGPU_HASH_CHECKSUM 3
#define VTMGPU_HASH_SLOT_BYTES 48
int vtmgpu_hash(vtmgpu_ctx* ctx, int first, int count, int kind, uint8_t* digest, int* bytes_per_component);

/* ---------------------------------------------------------------------------------------------
 * band mode over PEER MEMORY (NVLink / NVSwitch): no copies, no NCCL, no host synchronisation inside an iteration.
 * Every rank (one process per GPU) exports a handle of its plane memory and flag block, the handles travel by any host channel
 * (vvc_b200/bands.py: torch.distributed all_gather), and every rank connects to the ranks that own the bands above and below.
 * vtmgpu_band_filter_async then enqueues the whole chain for the rank's band: the deblocking + SAO kernel stores the four rows on
 * each side of a band border ALSO into the neighbour's plane and releases a flag there, the ALF kernel acquires the flags before
 * it loads the tiles of the band's first / last tile row.  All ranks must issue the same sequence of
 * vtmgpu_band_filter_async calls (the flags count iterations).  Contexts of all ranks need the same geometry and capacity;
 * pictures that need the virtual-boundary / CTU-32 kernel are not supported in this mode (use the export / import calls above).
 * --------------------------------------------------------------------------------------------- */
#define VTMGPU_BAND_HANDLE_BYTES 160
typedef struct vtmgpu_band_handle { unsigned char bytes[VTMGPU_BAND_HANDLE_BYTES]; } vtmgpu_band_handle;
int vtmgpu_band_export(vtmgpu_ctx* ctx, vtmgpu_band_handle* handle);
/* above / below: handles of the ranks that own the neighbouring bands (NULL: picture border); call after vtmgpu_set_rows */
int vtmgpu_band_connect(vtmgpu_ctx* ctx, const vtmgpu_band_handle* above, const vtmgpu_band_handle* below);
int vtmgpu_band_filter_async(vtmgpu_ctx* ctx, int slot);
int vtmgpu_band_disconnect(vtmgpu_ctx* ctx);      /* every rank: after its last iteration has completed on ALL ranks (a host barrier) */

/* ---------------------------------------------------------------------------------------------
 * host batches: the whole boundary for a run of independent pictures that live in HOST memory, in ONE call.
 * A batch object owns `lanes` single-picture contexts (the buffers of the pictures in flight) and three CUDA streams -- uploads,
 * kernels, downloads; vtmgpu_batch_filter walks the pictures round robin over the lanes -- upload of the planes, record lists,
 * SAO / ALF parameters on the first stream, the chain on the second, the download on the third, ordered by events -- so that the
 * copies of one picture overlap the kernels and the opposite copies of others (use at least 3 lanes: the download of a picture is
 * then aligned with the upload after next, which the PCIe link rewards), all issued by the calling thread (one issuing
 * thread per GPU; a decoder that holds several reconstructed pictures of a GOP calls this once instead of ~10 entry points per
 * picture).
 * Page-locked buffers make every copy asynchronous; pageable buffers work but serialise.  Returns when every output has landed.
 * in / out may alias (in-place, like DecLib::executeLoopFilters on the picture's reco buffer).
 * --------------------------------------------------------------------------------------------- */
typedef struct vtmgpu_host_picture
{
  const int16_t* in[3];                  /* pre-filter planes (Picture::getRecoBuf), strides in samples */
  ptrdiff_t      in_stride[3];
  int16_t*       out[3];                 /* filtered planes */
  ptrdiff_t      out_stride[3];
  const vtmgpu_deblock_sparse* deblock;  /* NULL: deblocking off for this picture */
  const vtmgpu_sao_params*     sao;      /* reconstructed (vtmgpu_sao_reconstruct); NULL: SAO off */
  const vtmgpu_alf_params*     alf;      /* NULL: ALF off */
} vtmgpu_host_picture;

typedef struct vtmgpu_batch vtmgpu_batch;
int  vtmgpu_batch_create(const vtmgpu_seq_params* seq, int lanes, vtmgpu_batch** batch);   /* seq->capacity is ignored (one slot per lane) */
void vtmgpu_batch_destroy(vtmgpu_batch* batch);
int  vtmgpu_batch_filter(vtmgpu_batch* batch, const vtmgpu_host_picture* pictures, int count);
const char* vtmgpu_batch_last_error(const vtmgpu_batch* batch);
int64_t vtmgpu_batch_launch_count(const vtmgpu_batch* batch);

/* replay/benchmark support: enqueue the whole chain on the ctx stream without synchronising; timing by
 * CUDA events recorded on that same stream */
int vtmgpu_filter_async(vtmgpu_ctx* ctx, int first, int count);
int vtmgpu_sync(vtmgpu_ctx* ctx);
int vtmgpu_timer_start(vtmgpu_ctx* ctx);                 /* records an event on the ctx stream          */
int vtmgpu_timer_stop (vtmgpu_ctx* ctx, float* ms);      /* records + synchronises, elapsed ms          */
/* makes the pristine upload the current state of the slots again so that a replay can repeat the chain: the uploaded buffer is
 * never written by a kernel (the stages ping-pong between the two working buffers), so this only resets the slots' state -- no copy */
int vtmgpu_rewind(vtmgpu_ctx* ctx, int first, int count);
/* number of kernel launches issued by this ctx so far (bench.py "gpu_launches") */
int64_t vtmgpu_launch_count(const vtmgpu_ctx* ctx);
/* per-stage device time of the last vtmgpu_filter* call when profiling was enabled (ms; 0 = deblocking + SAO kernel, 1 = ALF kernel) */
int vtmgpu_set_profiling(vtmgpu_ctx* ctx, int on);
int vtmgpu_stage_ms(vtmgpu_ctx* ctx, float ms[4]);

#ifdef __cplusplus
}
#endif
#endif /* VTMGPU_H */
