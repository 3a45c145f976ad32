// vtmgpu_dev.cuh -- device-side data model shared by the kernels of libvtmgpu (sm_100a).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace vtmgpu
{

typedef int16_t pel;

// One sample plane in HBM: planar int16, no apron; pitch (in samples) is a multiple of 64 so every row starts on a
// 128-byte boundary and every aligned group of 8 samples (one 128-bit access) lies inside the row.
struct PlaneDev
{
  pel* p;
  int  pitch, w, h;
};

// SAO parameters of one CTU component, reconstructed and compacted (16 bytes): SampleAdaptiveOffset.cpp:528-541 keeps
// a 32-entry band table with 4 live bands; here the 4 band offsets are stored with the first band position.
struct SaoDev
{
  uint8_t type;      // 0 off, 1..4 = EO 0/90/135/45 degrees, 5 = BO
  uint8_t band;      // BO: first band
  uint8_t avail;     // VTMGPU_AVAIL_* bits of the 8 neighbouring CTUs
  uint8_t pad;
  int16_t off[5];    // EO: offsets for edgeType -2..2 ; BO: off[0..3] for bands band..band+3 (mod 32)
  int16_t pad2;
};

#define VTMGPU_MAX_LUMA_SETS 24

// One chroma filter alternative expanded for the packed 5x5 kernel (see AlfLumaEntry below)
struct AlfChromaEntry
{
  uint32_t coefB[6], clipP1[6], clip2[6];
  int32_t  bias;
  int32_t  pad;
};

// ALF data of one picture -- of one slice parameter set when the slices of a picture differ (a slot holds ALF_MAX_GROUPS of
// these; the luma sets of ALL groups live in the first one, the per-CTU set index is picture-global).  The chroma / CC-ALF
// operand tables come first: k_alf copies these ALF_SMALL_BYTES into shared memory with one bulk copy per tile.
#define ALF_SMALL_BYTES 768
#define ALF_MAX_GROUPS 8
struct AlfDev
{
  AlfChromaEntry chromaTab[8];                       // chroma alternatives expanded into the operands of the packed 5x5 kernel
  uint32_t ccK[2][4][4];                             // CC-ALF coefficients as two-tap operands of ccAlfQuadDual (alf_fast.cuh); [3] = 1 when the coefficient sum fits a byte
  int32_t enabled[3];
  int32_t ccEnabled[2];
  int32_t numSets;                                   // 16 fixed + APS sets
  int32_t wide;                                      // a luma coefficient does not fit the s8 operand of IDP.2A: generic path
  int32_t pad;
  uint32_t ccB[2][4][8];                             // CC-ALF coefficients as IDP.2A byte operands (bytes 0 and 3); [7] = sum of the 7 (16-byte aligned rows)
  short2  luma[VTMGPU_MAX_LUMA_SETS][25][12];        // {coeff, clip} per set, class, tap (transpose 0 order)
  short2  chroma[8][6];                              // {coeff, clip} per alternative, tap
  int16_t cc[2][4][8];                               // CC-ALF coefficients (7 used)
};

// One luma filter of a (filter set, class, transpose index), expanded on the host for the packed 7x7 kernel:
// tap k (after the transpose permutation, AdaptiveLoopFilter.cpp:1170-1189) as the three register operands it needs.
struct AlfLumaEntry
{
  uint32_t coefB[12];    // coefficient as s8 in byte 0 and byte 3 (IDP.2A.LO uses bytes 0-1, .HI bytes 2-3)
  uint32_t clipP1[12];   // (clip + 1) in both 16-bit lanes
  uint32_t clip2[12];    // (2 * clip) in both 16-bit lanes
  int32_t  bias;         // 64 - sum(coef * 2 * clip)
  int32_t  pad[7];       // 176 bytes = 44 words: entries e and e + 1 start 12 banks apart when a filter set sits in shared memory
};

// Per-CTU control record (16 bytes, one 128-bit load per tile): ALF control (Picture::getAlfCtuEnableFlag /
// getAlfCtbFilterIndex / getAlfCtuAlternativeData, m_ccAlfFilterControl) written by vtmgpu_set_alf.
struct alignas(16) CtuCtlDev
{
  uint8_t enY, enCb, enCr;       // ALF CTU enable flags
  uint8_t altCb, altCr;          // chroma filter alternative
  uint8_t ccCb, ccCr;            // CC-ALF filter idc (0 = off)
  uint8_t setIdx;                // luma filter set: < 16 fixed, else APS
  uint8_t clip;                  // VTMGPU_ALF_CLIP_* / PAD_*: partition boundaries the filter must not read across
  uint8_t flags;                 // bit 0: the CTU's slice runs ALF, bit 1: SlotDev::alfWide -- so that k_alf needs nothing but this record
  uint8_t grp;                   // which AlfDev of the slot holds the chroma / CC-ALF data of the CTU's slice (vtmgpu_set_alf_slices; else 0)
  uint8_t pad[5];
};

struct LadfDev                   // vtmgpu_ladf; n = 0: off (the luma records carry tc / beta)
{
  int32_t n, off[5], lb[5], pad;
};

struct VbDev                     // vtmgpu_virtual_boundaries (luma sample positions); nv = nh = 0: none
{
  int32_t nv, nh, x[3], y[3];
};

struct SlotDev
{
  PlaneDev buf[3][3];            // [buffer][component]; buffer 0 = pristine upload, 1/2 = working
  const uint32_t* dbfL[2];
  const uint64_t* dbfC[2];
  const uint64_t* dbfQ;          // per picture tile (luma, Cb, Cr) the queues of the active segments of both passes (k_dbf_queues, dbf_kernel.cuh)
  const uint32_t* dbfQCnt;       // [tiles][2] their lengths
  const SaoDev*   sao;           // [ctus][3]                     (NULL: stage off)
  const AlfDev*   alf;           // [ALF_MAX_GROUPS]              (NULL: stage off)
  const CtuCtlDev* ctuCtl;       // [ctus]
  const AlfLumaEntry* lumaTab;   // [sets][25 classes][4 transposes]
  const int16_t* lmcs;           // LMCS inverse table (vtmgpu_set_lmcs)
  int32_t lmcsOn;                // the luma plane of buffer 0 is in the reshaped domain: the stage that reads it maps it through lmcs
  int32_t dbfOn, saoOn, alfOn;   // alfOn: parameters set AND the slice enables ALF for at least one component
  int32_t alfWide;               // a luma coefficient does not fit the s8 operand of IDP.2A: generic path
  LadfDev ladf;
  VbDev vbSao, vbAlf;            // as given to vtmgpu_set_sao / vtmgpu_set_alf
};

// where the planes and the ALF side information of a slot live: all slots of a context share one allocation each, so k_alf
// computes its addresses from the slot number (round 1 loaded five pointers per tile from the slot table)
struct AlfAddr
{
  pel* planes;                   // slot s, buffer b, component k: planes + s * slotStride + b * bufStride + compOff[k]
  size_t slotStride, bufStride, compOff[3];
  int pitchY, pitchC;            // plane pitches in samples
  const unsigned char* side;     // slot s: side + s * sideStride ; + offTab = AlfLumaEntry[sets][100], + offAlf = AlfDev, + offCtl = CtuCtlDev[ctus]
  size_t sideStride, offTab, offAlf, offCtl;
};

// Band mode over peer memory (one picture split into CTU-row bands, one GPU per band): the neighbours' plane
// allocations are mapped into this GPU's address space (CUDA IPC over NVLink).  k_dbf_sao stores the four rows of SAO output on
// each side of a band border ALSO into the neighbour's plane -- the rows its ALF reads across the border -- and the last CTA to
// finish releases a flag in the neighbour's memory; k_alf's walking thread acquires the flag before it loads a tile of the band's
// first / last tile row.  A second pair of flags travels the other way (ALF finished reading) so that the next iteration's stores
// do not overtake the neighbour's reads.  No host synchronisation and no copy kernels inside an iteration.
//   flags (uint32, in the memory of the rank that WAITS on them): [0] rows from above have landed, [1] rows from below,
//   [2] the neighbour above has finished its ALF, [3] the neighbour below, [4] [5] exit counters of this rank's two kernels
struct BandDev
{
  pel* myPlanes;                 // base of this context's plane allocation (the neighbours' have the same layout)
  pel* peerPlanes[2];            // the neighbour above / below, peer-mapped; nullptr = no neighbour on that side
  uint32_t* peerFlags[2];
  uint32_t* myFlags;             // nullptr: not in peer band mode
  uint32_t iter;                 // 1, 2, ... : the value the flags carry
  int rowBegin, rowEnd;          // luma rows of this band
};

__device__ __forceinline__ uint32_t ldAcquireSys(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stReleaseSys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

struct Geom
{
  int w, h;             // luma
  int sx, sy;           // chroma shifts
  int ncomp;
  int bdL, bdC;
  int ctu, ctuLog2;
  int wCtus, hCtus;
};

__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }
__device__ __forceinline__ int clip3(int lo, int hi, int v) { return min(max(v, lo), hi); }

}   // namespace vtmgpu
