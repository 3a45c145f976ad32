// vtmgpu.cu -- libvtmgpu: host side of the C ABI (include/vtmgpu.h) + kernel launches.  sm_100a only, no CPU path.
//
// HBM layout per picture slot (sized for 180 GB: a 3840x2160 4:2:0 slot is 3 x 24.9 MB planes + 6.7 MB side info; all slots
// of a context share one plane allocation and one side-info allocation at constant strides):
//   buf[0]  pristine upload (never written by a kernel, so a replay can rewind without another H2D)
//   buf[1], buf[2]  working buffers: every stage kernel reads the slot's current buffer and writes the other working one
//           (k_dbf_sao: deblocking + SAO, k_alf / k_alf_parts: ALF + CC-ALF; vtmgpu_filter = buf[0] -> buf[1] -> buf[2])
//   side    one contiguous block: luma / chroma segment records of both directions (ABI indexing, TMA row pitch), the
//           expanded luma filter tables (16 fixed sets + APS sets), CtuCtlDev[ctus], AlfDev, SaoDev[ctus][3] -- mirrored in
//           pinned host memory: vtmgpu_set_* pack into the mirror, the next stage call uploads the changed range.
// All work of a ctx is enqueued on its stream; stage calls synchronise unless named *_async.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <cuda.h>
#include <string>
#include <vector>

#include "dbf_kernel.cuh"
#include "alf_kernel.cuh"
#include "hash_kernel.cuh"
#include "border_kernel.cuh"
#include "vtmgpu.h"
#include "vtmgpu_derive.h"
#include "vvc_alf_fixed_tables.h"

using namespace vtmgpu;

namespace
{
std::string g_createError;

struct SideLayout     // byte offsets inside a slot's side-info block
{
  size_t dbfL[2], dbfC[2], dbfEnd, sao, lmcs, alfTab, ctuCtl, alf, total;
  size_t nL, nC[2];          // records of the ABI arrays (dense)
  int recW[4], recH[4], recP[4];   // device record arrays lumaV, lumaH, chromaV, chromaH: width, height, row pitch (records; pitch * size is a multiple of 16 B for TMA)
};

size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

// tap order of the 7x7 diamond for transposeIdx 0..3 (AdaptiveLoopFilter.cpp:1170-1189)
const int8_t kPerm7[4][12] = { { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 }, { 9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6 },
                               { 0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11 }, { 9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6 } };

// expands one luma filter set ({coeff, clip} per class and tap) into the per-(class, transpose) operand table of the packed
// kernel; returns true when a coefficient does not fit the signed 8-bit operand of IDP.2A
bool expandLumaSet(const short2 (*set)[12], AlfLumaEntry* out)
{
  bool wide = false;
  for (int cl = 0; cl < 25; cl++)
    for (int t = 0; t < 4; t++)
    {
      AlfLumaEntry& e = out[cl * 4 + t];
      int bias = 64;
      for (int k = 0; k < 12; k++)
      {
        const int co = set[cl][kPerm7[t][k]].x, clp = set[cl][kPerm7[t][k]].y;
        wide |= co < -128 || co > 127;
        e.coefB[k] = (uint32_t)(co & 0xff) * 0x01000001u;
        e.clipP1[k] = (uint32_t)((clp + 1) & 0xffff) * 0x10001u;
        e.clip2[k] = (uint32_t)((2 * clp) & 0xffff) * 0x10001u;
        bias -= co * 2 * clp;
      }
      e.bias = bias;
      e.pad[0] = e.pad[1] = e.pad[2] = 0;
    }
  return wide;
}
}   // namespace

// arguments of k_dbf_scatter (record lists -> record arrays of a slot)
struct ScatterArgs
{
  const void* list[4];
  void* dense[4];
  uint32_t count[4], first[5];          // first[a] = number of entries before array a
  int recW[4], recH[4], recP[4];
};

struct vtmgpu_ctx
{
  vtmgpu_seq_params seq{};
  Geom g{};
  int nCtus = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t ownStream = nullptr;    // created by vtmgpu_create; `stream` may be redirected to a caller's stream (vtmgpu_set_stream)
  bool asyncStages = false;            // vtmgpu_set_async: stage / row-copy calls only enqueue
  cudaEvent_t ev[2] = { nullptr, nullptr };
  cudaEvent_t stageEv[3] = { nullptr, nullptr, nullptr };
  bool profiling = false, stageValid = false;
  std::string err;
  SideLayout lay{};
  std::vector<pel*> planeMem;          // per slot (3 buffers x ncomp planes), all slots in ONE allocation at a constant stride: k_alf computes
  std::vector<unsigned char*> sideDev; // per slot                                 plane and table addresses instead of loading them
  pel* planeAll = nullptr;
  unsigned char* sideAll = nullptr;
  size_t planeAllBytes = 0;
  AlfAddr alfAddr{};
  // peer band mode (vtmgpu_band_*): flag block of this rank, the neighbours' mapped memory, iteration counter
  uint32_t* bandFlags = nullptr;
  void* bandPeerMem[2][2] = { { nullptr, nullptr }, { nullptr, nullptr } };     // [above / below][planes, flags] as opened (for cudaIpcCloseMemHandle)
  BandDev band{};
  bool bandOn = false, bandCall = false;   // connected ; inside vtmgpu_band_filter_async (only those launches carry the flags)
  std::vector<unsigned char*> unitsDev;  // per slot, allocated by the first vtmgpu_set_deblock_units: landing area of the block structure
  std::vector<size_t> unitsBytes;
  struct MotionDev { unsigned char* dev = nullptr; size_t bytes = 0; int elemBytes = 0, offMv0 = 0, offMv1 = 0, offRef0 = 0, offRef1 = 0; bool loaded = false; };
  std::vector<MotionDev> motionDev;      // per slot: the motion field of the picture in the slot (vtmgpu_upload_motion / vtmgpu_set_deblock_units)
  pel* extendBuf = nullptr;              // picture + margins of one slot (vtmgpu_download_extended), allocated by the first call
  size_t extendElems = 0;
  std::vector<unsigned char*> sparseDev; // per slot, allocated by the first vtmgpu_set_deblock_sparse: landing area of the record lists
  // vtmgpu_set_deblock_sparse only enqueues the copies of the lists; the kernels that turn them into record arrays and tile queues
  // run with the next stage call, BEHIND the uploads of the small side information (flush).  The copy engine takes the H2D
  // operations of all streams in issue order: a copy that waits for a kernel of its own stream holds up the plane uploads of the
  // pictures behind it (measured: 10 of 44 ms per 64-picture batch of vtmgpu_batch_filter).
  std::vector<ScatterArgs> sparseArgs;
  std::vector<char> sparsePending;
  unsigned char* sidePinned = nullptr; // capacity * lay.total
  SlotDev* slotsPinned = nullptr;      // capacity entries (pinned mirror)
  SlotDev* slotsDev = nullptr;
  std::vector<int> cur;                // buffer index holding the current state of each slot
  int64_t launches = 0;
  int numSms = 0;
  int rowBegin = 0, rowEnd = 0;         // luma rows the stage calls filter (band mode); whole picture by default
  CUtensorMap* tmapsDev = nullptr;     // [capacity][3 buffers][3 planes]: TMA descriptors of the plane buffers, box = smem tile of k_alf
  CUtensorMap* tmapsDbfDev = nullptr;  // the same planes, box = smem tile of k_dbf_sao
  unsigned char* dbfQueueAll = nullptr;  // per slot: [tiles][DBF_QTILE] queue entries + [tiles][2] lengths (k_dbf_queues)
  size_t dbfQueueStride = 0;
  DbfLaunch dbfFull{};                 // tile grid of the whole picture

  int fail(const char* fmt, ...)
  {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return -1;
  }
  int cuda(cudaError_t e, const char* what)
  {
    if (e == cudaSuccess) return 0;
    return fail("%s: %s", what, cudaGetErrorString(e));
  }
  bool slotOk(int first, int count) { return first >= 0 && count >= 1 && first + count <= seq.capacity; }
  unsigned char* pinnedSide(int slot) { return sidePinned + (size_t)slot * lay.total; }
  // Small per-picture side information is uploaded lazily: vtmgpu_set_* fill the pinned mirror and mark it, the stage calls
  // push one range per slot and one run of slot descriptors (every H2D operation costs several microseconds of copy-engine
  // time whatever its size, which is what bounds the end-to-end rate once the planes themselves move at PCIe speed).
  std::vector<size_t> dirtyLo, dirtyHi;     // per slot: byte range of the side block to upload (lo >= hi: clean)
  std::vector<char> slotDirty;
  // the uploads read the pinned mirror asynchronously: an event per slot marks the last upload that read it, and every vtmgpu_set_*
  // waits for it before it packs into the mirror again (a caller that pipelines pictures through one ctx with the *_async calls)
  std::vector<cudaEvent_t> mirrorEv;
  std::vector<char> mirrorBusy;
  void mirrorRead(int slot)
  {
    if (cudaEventRecord(mirrorEv[slot], stream) == cudaSuccess) mirrorBusy[slot] = 1;
  }
  void mirrorWrite(int slot)
  {
    if (mirrorBusy[slot]) { cudaEventSynchronize(mirrorEv[slot]); mirrorBusy[slot] = 0; }
  }
  int pushSlot(int slot) { slotDirty[slot] = 1; return 0; }
  void markSide(int slot, size_t off, size_t bytes)
  {
    dirtyLo[slot] = dirtyLo[slot] < dirtyHi[slot] ? std::min(dirtyLo[slot], off) : off;
    dirtyHi[slot] = std::max(dirtyHi[slot], off + bytes);
  }
  int flush(int first, int count)
  {
    for (int s = first; s < first + count; s++)
      if (dirtyLo[s] < dirtyHi[s])
      {
        if (pushSide(s, dirtyLo[s], dirtyHi[s] - dirtyLo[s])) return -1;
        dirtyLo[s] = dirtyHi[s] = 0;
        mirrorRead(s);
      }
    for (int s = first; s < first + count; s++)
    {
      if (!slotDirty[s]) continue;
      int e = s;
      while (e < first + count && slotDirty[e]) slotDirty[e++] = 0;
      if (cuda(cudaMemcpyAsync(slotsDev + s, slotsPinned + s, sizeof(SlotDev) * (e - s), cudaMemcpyHostToDevice, stream), "slot table upload")) return -1;
      for (int k = s; k < e; k++) mirrorRead(k);
      s = e - 1;
    }
    return 0;
  }
  int pushSide(int slot, size_t off, size_t bytes)
  {
    return cuda(cudaMemcpyAsync(sideDev[slot] + off, pinnedSide(slot) + off, bytes, cudaMemcpyHostToDevice, stream), "side info upload");
  }
};

static int runPendingSparse(vtmgpu_ctx* c, int first, int count);

// ------------------------------------------------------------------------------------------------------------
extern "C" int vtmgpu_abi_version(void) { return VTMGPU_ABI_VERSION; }

extern "C" const char* vtmgpu_last_error(const vtmgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

// sizes of the ABI structures, so that language bindings can verify their mirrors
extern "C" int vtmgpu_abi_sizeof(int which)
{
  switch (which)
  {
  case 0: return (int)sizeof(vtmgpu_seq_params);
  case 1: return (int)sizeof(vtmgpu_deblock_params);
  case 2: return (int)sizeof(vtmgpu_sao_offset);
  case 3: return (int)sizeof(vtmgpu_sao_ctu);
  case 4: return (int)sizeof(vtmgpu_sao_params);
  case 5: return (int)sizeof(vtmgpu_alf_luma_aps);
  case 6: return (int)sizeof(vtmgpu_alf_chroma_aps);
  case 7: return (int)sizeof(vtmgpu_alf_params);
  case 8: return (int)sizeof(vtmgpu_deblock_sparse);
  case 9: return (int)sizeof(vtmgpu_ladf);
  case 10: return (int)sizeof(vtmgpu_virtual_boundaries);
  case 11: return (int)sizeof(vtmgpu_host_picture);
  default: return -1;
  }
}

extern "C" int vtmgpu_band_disconnect(vtmgpu_ctx* c);

extern "C" void vtmgpu_destroy(vtmgpu_ctx* c)
{
  if (!c) return;
  cudaSetDevice(c->seq.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  vtmgpu_band_disconnect(c);
  if (c->bandFlags) cudaFree(c->bandFlags);
  if (c->planeAll) cudaFree(c->planeAll);
  if (c->sideAll) cudaFree(c->sideAll);
  for (unsigned char* p : c->sparseDev) if (p) cudaFree(p);
  if (c->extendBuf) cudaFree(c->extendBuf);
  for (unsigned char* p : c->unitsDev) if (p) cudaFree(p);
  for (auto& m : c->motionDev) if (m.dev) cudaFree(m.dev);
  if (c->sidePinned) cudaFreeHost(c->sidePinned);
  if (c->slotsPinned) cudaFreeHost(c->slotsPinned);
  if (c->slotsDev) cudaFree(c->slotsDev);
  if (c->tmapsDev) cudaFree(c->tmapsDev);
  if (c->tmapsDbfDev) cudaFree(c->tmapsDbfDev);
  if (c->dbfQueueAll) cudaFree(c->dbfQueueAll);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  for (auto& e : c->stageEv) if (e) cudaEventDestroy(e);
  for (auto& e : c->mirrorEv) if (e) cudaEventDestroy(e);
  if (c->ownStream) cudaStreamDestroy(c->ownStream);
  delete c;
}

extern "C" int vtmgpu_create(const vtmgpu_seq_params* seq, vtmgpu_ctx** out)
{
  if (!seq || !out) { g_createError = "vtmgpu_create: null argument"; return -1; }
  *out = nullptr;
  const vtmgpu_seq_params& s = *seq;
  auto bad = [](const char* m) { g_createError = std::string("vtmgpu_create: ") + m; return -1; };
  if (s.width <= 0 || s.height <= 0 || (s.width & 7) || (s.height & 7)) return bad("width/height must be positive multiples of 8");
  if (s.chroma_format < 0 || s.chroma_format > 3) return bad("bad chroma_format");
  if (s.bit_depth_luma < 8 || s.bit_depth_luma > 12 || s.bit_depth_chroma < 8 || s.bit_depth_chroma > 12) return bad("bit depth must be 8..12");
  if (s.ctu_size != 32 && s.ctu_size != 64 && s.ctu_size != 128) return bad("ctu_size must be 32, 64 or 128");
  if (s.capacity < 1) return bad("capacity must be >= 1");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return bad("no CUDA device available (libvtmgpu has no CPU fallback)");
  if (s.device < 0 || s.device >= ndev) return bad("bad device ordinal");
  if ((e = cudaSetDevice(s.device)) != cudaSuccess) return bad(cudaGetErrorString(e));

  vtmgpu_ctx* c = new vtmgpu_ctx();
  c->seq = s;
  Geom& g = c->g;
  g.w = s.width; g.h = s.height;
  g.sx = (s.chroma_format == 1 || s.chroma_format == 2) ? 1 : 0;
  g.sy = (s.chroma_format == 1) ? 1 : 0;
  g.ncomp = s.chroma_format == 0 ? 1 : 3;
  g.bdL = s.bit_depth_luma; g.bdC = s.bit_depth_chroma;
  g.ctu = s.ctu_size; g.ctuLog2 = s.ctu_size == 128 ? 7 : (s.ctu_size == 64 ? 6 : 5);
  g.wCtus = (g.w + g.ctu - 1) / g.ctu; g.hCtus = (g.h + g.ctu - 1) / g.ctu;
  c->nCtus = g.wCtus * g.hCtus;
  c->rowBegin = 0; c->rowEnd = g.h;

  SideLayout& L = c->lay;
  L.nL = (size_t)(g.w / 4) * (g.h / 4);
  L.nC[0] = g.ncomp > 1 ? (size_t)((g.w + (8 << g.sx) - 1) / (8 << g.sx)) * (g.h / 4) : 0;
  L.nC[1] = g.ncomp > 1 ? (size_t)((g.h + (8 << g.sy) - 1) / (8 << g.sy)) * (g.w / 4) : 0;
  size_t off = 0;
  L.recW[0] = L.recW[1] = g.w / 4; L.recH[0] = L.recH[1] = g.h / 4;
  L.recW[2] = g.ncomp > 1 ? (g.w + (8 << g.sx) - 1) / (8 << g.sx) : 0; L.recH[2] = g.ncomp > 1 ? g.h / 4 : 0;
  L.recW[3] = g.ncomp > 1 ? g.w / 4 : 0;                                L.recH[3] = g.ncomp > 1 ? (g.h + (8 << g.sy) - 1) / (8 << g.sy) : 0;
  for (int a = 0; a < 4; a++) L.recP[a] = (int)alignUp(L.recW[a], a < 2 ? 4 : 2);
  for (int d = 0; d < 2; d++) { L.dbfL[d] = off; off = alignUp(off + (size_t)L.recP[d] * L.recH[d] * 4, 256); }
  for (int d = 0; d < 2; d++) { L.dbfC[d] = off; off = alignUp(off + (size_t)L.recP[2 + d] * L.recH[2 + d] * 8, 256); }
  L.dbfEnd = off;
  // the per-picture parts (APS filter tables, CTU control, ALF parameters, SAO parameters) are adjacent: one upload per picture
  L.alfTab = off; off = alignUp(off + sizeof(AlfLumaEntry) * VTMGPU_MAX_LUMA_SETS * 25 * 4, 256);
  L.ctuCtl = off; off = alignUp(off + (size_t)c->nCtus * sizeof(CtuCtlDev), 256);
  L.sao = off;    off = alignUp(off + (size_t)c->nCtus * 3 * sizeof(SaoDev), 256);
  L.lmcs = off;   off = alignUp(off + sizeof(int16_t) * 4096, 256);            // LMCS inverse table (<= 12 bit)
  L.alf = off;    off = alignUp(off + sizeof(AlfDev) * ALF_MAX_GROUPS, 256);   // last: a picture uploads only the groups it uses
  L.total = off;

#define CK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_createError = std::string("vtmgpu_create: ") + what + ": " + cudaGetErrorString(e_); vtmgpu_destroy(c); return -1; } } while (0)
  CK(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking), "stream");
  c->stream = c->ownStream;
  for (auto& ev : c->ev) CK(cudaEventCreate(&ev), "event");
  for (auto& ev : c->stageEv) CK(cudaEventCreate(&ev), "event");
  CK(cudaHostAlloc((void**)&c->sidePinned, L.total * s.capacity, cudaHostAllocDefault), "pinned side info");
  CK(cudaHostAlloc((void**)&c->slotsPinned, sizeof(SlotDev) * s.capacity, cudaHostAllocDefault), "pinned slot table");
  CK(cudaMalloc((void**)&c->slotsDev, sizeof(SlotDev) * s.capacity), "slot table");
  memset(c->sidePinned, 0, L.total * s.capacity);
  memset(c->slotsPinned, 0, sizeof(SlotDev) * s.capacity);
  c->cur.assign(s.capacity, 0);
  c->dirtyLo.assign(s.capacity, 0);
  c->dirtyHi.assign(s.capacity, 0);
  c->slotDirty.assign(s.capacity, 1);
  c->mirrorEv.assign(s.capacity, nullptr);
  c->mirrorBusy.assign(s.capacity, 0);
  for (auto& ev : c->mirrorEv) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event");

  // plane geometry: pitch multiple of 64 samples (128 B)
  int pw[3], ph[3], pitch[3];
  size_t planeElems[3], slotElems = 0;
  for (int k = 0; k < g.ncomp; k++)
  {
    pw[k] = k ? g.w >> g.sx : g.w;
    ph[k] = k ? g.h >> g.sy : g.h;
    pitch[k] = (int)alignUp(pw[k], 64);
    planeElems[k] = alignUp((size_t)pitch[k] * ph[k], 128);
    slotElems += planeElems[k];
  }
  c->planeAllBytes = slotElems * 3 * sizeof(pel) * s.capacity;
  CK(cudaMalloc((void**)&c->planeAll, slotElems * 3 * sizeof(pel) * s.capacity), "plane memory");
  CK(cudaMemsetAsync(c->planeAll, 0, slotElems * 3 * sizeof(pel) * s.capacity, c->stream), "memset");
  CK(cudaMalloc((void**)&c->sideAll, L.total * s.capacity), "side info memory");
  CK(cudaMemsetAsync(c->sideAll, 0, L.total * s.capacity, c->stream), "memset");
  {
    AlfAddr& a = c->alfAddr;
    a.planes = c->planeAll; a.slotStride = slotElems * 3; a.bufStride = slotElems;
    a.compOff[0] = 0; a.compOff[1] = planeElems[0]; a.compOff[2] = g.ncomp > 1 ? planeElems[0] + planeElems[1] : 0;
    a.pitchY = pitch[0]; a.pitchC = g.ncomp > 1 ? pitch[1] : 0;
    a.side = c->sideAll; a.sideStride = L.total; a.offTab = L.alfTab; a.offAlf = L.alf; a.offCtl = L.ctuCtl;
  }
  {
    // per-tile queues of the active deblocking segments (k_dbf_queues): tile grid of the whole picture
    DbfLaunch& F = c->dbfFull;
    F.tilesXL = (g.w + DBF_TW - 1) / DBF_TW; F.ty0L = 0;
    F.tilesL = F.tilesLFull = F.tilesXL * ((g.h + DBF_TH - 1) / DBF_TH);
    F.tilesXC = g.ncomp > 1 ? ((g.w >> g.sx) + DBF_TW - 1) / DBF_TW : 0; F.ty0C = 0;
    F.tilesC = F.tilesCFull = g.ncomp > 1 ? F.tilesXC * (((g.h >> g.sy) + DBF_TH - 1) / DBF_TH) : 0;
    const size_t tiles = (size_t)F.tilesLFull + 2 * F.tilesCFull;
    c->dbfQueueStride = alignUp(tiles * DBF_QTILE * 8 + tiles * 2 * 4, 256);
    CK(cudaMalloc((void**)&c->dbfQueueAll, c->dbfQueueStride * s.capacity), "deblocking queue memory");
    CK(cudaMemsetAsync(c->dbfQueueAll, 0, c->dbfQueueStride * s.capacity, c->stream), "memset");
  }
  for (int sl = 0; sl < s.capacity; sl++)
  {
    pel* mem = c->planeAll + (size_t)sl * slotElems * 3;
    unsigned char* side = c->sideAll + (size_t)sl * L.total;
    c->planeMem.push_back(mem);
    c->sideDev.push_back(side);
    SlotDev& sd = c->slotsPinned[sl];
    pel* p = mem;
    for (int b = 0; b < 3; b++)
      for (int k = 0; k < g.ncomp; k++)
      {
        sd.buf[b][k] = PlaneDev{ p, pitch[k], pw[k], ph[k] };
        p += planeElems[k];
      }
    for (int d = 0; d < 2; d++)
    {
      sd.dbfL[d] = reinterpret_cast<const uint32_t*>(side + L.dbfL[d]);
      sd.dbfC[d] = reinterpret_cast<const uint64_t*>(side + L.dbfC[d]);
    }
    sd.sao = reinterpret_cast<const SaoDev*>(side + L.sao);
    sd.alf = reinterpret_cast<const AlfDev*>(side + L.alf);
    sd.ctuCtl = reinterpret_cast<const CtuCtlDev*>(side + L.ctuCtl);
    sd.lumaTab = reinterpret_cast<const AlfLumaEntry*>(side + L.alfTab);
    sd.lmcs = reinterpret_cast<const int16_t*>(side + L.lmcs);
    sd.dbfQ = reinterpret_cast<const uint64_t*>(c->dbfQueueAll + (size_t)sl * c->dbfQueueStride);
    sd.dbfQCnt = reinterpret_cast<const uint32_t*>(c->dbfQueueAll + (size_t)sl * c->dbfQueueStride + ((size_t)c->dbfFull.tilesLFull + 2 * c->dbfFull.tilesCFull) * DBF_QTILE * 8);
    sd.dbfOn = sd.saoOn = sd.alfOn = sd.lmcsOn = 0;
  }
  {
    // TMA descriptors: one per plane buffer; box = the shared-memory tile of k_sao_alf (row pitch incl. padding), zero fill outside
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres), "cuTensorMapEncodeTiled lookup");
    if (!fn || qres != cudaDriverEntryPointSuccess) { g_createError = "vtmgpu_create: driver has no cuTensorMapEncodeTiled"; vtmgpu_destroy(c); return -1; }
    const SaLayout SL = saLayout(g.sx, g.sy, g.ncomp);
    for (int which = 0; which < 2; which++)
    {
      std::vector<CUtensorMap> maps((size_t)s.capacity * 9);
      memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
      for (int sl = 0; sl < s.capacity; sl++)
        for (int b = 0; b < 3; b++)
          for (int k = 0; k < g.ncomp; k++)
          {
            const PlaneDev& pd = c->slotsPinned[sl].buf[b][k];
            const cuuint64_t dims[2] = { (cuuint64_t)pd.w, (cuuint64_t)pd.h }, strides[1] = { (cuuint64_t)pd.pitch * 2 };
            const cuuint32_t boxAlf[2] = { (cuuint32_t)(k ? SL.pitchC : SA_P), (cuuint32_t)(k ? SL.rowsC : SA_H) };
            const cuuint32_t boxDbf[2] = { (cuuint32_t)DBF_PITCH, (cuuint32_t)DBF_SH }, es[2] = { 1, 1 };
            const CUresult r = reinterpret_cast<EncodeFn>(fn)(&maps[((size_t)sl * 3 + b) * 3 + k], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, pd.p, dims, strides,
                                                              which ? boxDbf : boxAlf, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { g_createError = "vtmgpu_create: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; vtmgpu_destroy(c); return -1; }
          }
      CUtensorMap** dev = which ? &c->tmapsDbfDev : &c->tmapsDev;
      CK(cudaMalloc((void**)dev, maps.size() * sizeof(CUtensorMap)), "tensor maps");
      CK(cudaMemcpy(*dev, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice), "tensor maps upload");
    }
  }
  {
    // the 16 fixed luma filter sets (AdaptiveLoopFilter.cpp:204-289, clip = 1 << bitDepth :500-509) never change: expand and upload once
    std::vector<short2> fixedSet(25 * 12);
    for (int fs = 0; fs < VTMGPU_ALF_FIXED_SETS; fs++)
    {
      for (int cl = 0; cl < 25; cl++)
        for (int k = 0; k < 12; k++) fixedSet[cl * 12 + k] = make_short2(vvc_alf_fix_coeff[vvc_alf_class_to_filt[fs * 25 + cl] * 12 + k], (short)(1 << g.bdL));
      expandLumaSet(reinterpret_cast<const short2(*)[12]>(fixedSet.data()), reinterpret_cast<AlfLumaEntry*>(c->pinnedSide(0) + L.alfTab) + fs * 100);
    }
    const size_t fixedBytes = sizeof(AlfLumaEntry) * VTMGPU_ALF_FIXED_SETS * 100;
    for (int sl = 0; sl < s.capacity; sl++)
    {
      if (sl) memcpy(c->pinnedSide(sl) + L.alfTab, c->pinnedSide(0) + L.alfTab, fixedBytes);
      CK(cudaMemcpyAsync(c->sideDev[sl] + L.alfTab, c->pinnedSide(sl) + L.alfTab, fixedBytes, cudaMemcpyHostToDevice, c->stream), "fixed filter table upload");
    }
  }
  CK(cudaMemcpyAsync(c->slotsDev, c->slotsPinned, sizeof(SlotDev) * s.capacity, cudaMemcpyHostToDevice, c->stream), "slot table upload");
  CK(cudaFuncSetAttribute(k_dbf_sao, cudaFuncAttributeMaxDynamicSharedMemorySize, DBF_SMEM_BYTES), "smem attribute");
  {
    const SaLayout SL = saLayout(g.sx, g.sy, g.ncomp), ST = saLayout(g.sx, g.sy, g.ncomp, true);      // parts: + one more tile: scratch copy for tiles cut by a virtual boundary
    CK(cudaFuncSetAttribute(k_alf<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST.total), "smem attribute");
    CK(cudaFuncSetAttribute(k_alf<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST.total), "smem attribute");
    CK(cudaFuncSetAttribute(k_alf_parts, cudaFuncAttributeMaxDynamicSharedMemorySize, SL.total + SL.lumaBytes + 2 * SL.chromaBytes), "smem attribute");
  }
  CK(cudaDeviceGetAttribute(&c->numSms, cudaDevAttrMultiProcessorCount, s.device), "device attribute");
  CK(cudaStreamSynchronize(c->stream), "sync");
#undef CK
  *out = c;
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// planes
// ------------------------------------------------------------------------------------------------------------
extern "C" int vtmgpu_upload_async(vtmgpu_ctx* c, int slot, const int16_t* const plane[3], const ptrdiff_t stride[3])
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("upload: bad slot %d", slot);
  cudaSetDevice(c->seq.device);
  const SlotDev& sd = c->slotsPinned[slot];
  for (int k = 0; k < c->g.ncomp; k++)
  {
    if (!plane[k]) return c->fail("upload: plane %d is NULL", k);
    const PlaneDev& d = sd.buf[0][k];
    // rows that are contiguous on both sides (4K / 8K / 1080p: the width is a multiple of the 64-sample pitch unit) go as ONE linear
    // copy: the copy engine moves a pitched 2D copy row by row, measurably below the PCIe rate of a linear one
    const bool linear = d.pitch == d.w && stride[k] == (ptrdiff_t)d.w;
    if (c->cuda(linear ? cudaMemcpyAsync(d.p, plane[k], (size_t)d.w * d.h * 2, cudaMemcpyHostToDevice, c->stream)
                       : cudaMemcpy2DAsync(d.p, (size_t)d.pitch * 2, plane[k], (size_t)stride[k] * 2, (size_t)d.w * 2, d.h, cudaMemcpyHostToDevice, c->stream), "upload")) return -1;
  }
  c->cur[slot] = 0;
  return 0;
}

extern "C" int vtmgpu_download_async(vtmgpu_ctx* c, int slot, int16_t* const plane[3], const ptrdiff_t stride[3])
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("download: bad slot %d", slot);
  cudaSetDevice(c->seq.device);
  const SlotDev& sd = c->slotsPinned[slot];
  for (int k = 0; k < c->g.ncomp; k++)
  {
    if (!plane[k]) return c->fail("download: plane %d is NULL", k);
    const PlaneDev& d = sd.buf[c->cur[slot]][k];
    const bool linear = d.pitch == d.w && stride[k] == (ptrdiff_t)d.w;
    if (c->cuda(linear ? cudaMemcpyAsync(plane[k], d.p, (size_t)d.w * d.h * 2, cudaMemcpyDeviceToHost, c->stream)
                       : cudaMemcpy2DAsync(plane[k], (size_t)stride[k] * 2, d.p, (size_t)d.pitch * 2, (size_t)d.w * 2, d.h, cudaMemcpyDeviceToHost, c->stream), "download")) return -1;
  }
  return 0;
}

extern "C" int vtmgpu_sync(vtmgpu_ctx* c)
{
  if (!c) return -1;
  cudaSetDevice(c->seq.device);
  return c->cuda(cudaStreamSynchronize(c->stream), "sync");
}

extern "C" int vtmgpu_upload(vtmgpu_ctx* c, int slot, const int16_t* const plane[3], const ptrdiff_t stride[3])
{
  return vtmgpu_upload_async(c, slot, plane, stride) ? -1 : vtmgpu_sync(c);
}

extern "C" int vtmgpu_download(vtmgpu_ctx* c, int slot, int16_t* const plane[3], const ptrdiff_t stride[3])
{
  return vtmgpu_download_async(c, slot, plane, stride) ? -1 : vtmgpu_sync(c);
}

extern "C" int vtmgpu_host_register(void* ptr, size_t bytes)
{
  if (!ptr || !bytes) return -1;
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) { cudaGetLastError(); return -1; }
  return 0;
}

extern "C" int vtmgpu_host_unregister(void* ptr)
{
  if (!ptr) return -1;
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return -1; }
  return 0;
}

extern "C" int vtmgpu_download_extended(vtmgpu_ctx* c, int slot, int16_t* const plane[3], const ptrdiff_t stride[3], int margin_luma)
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("download_extended: bad slot %d", slot);
  if (margin_luma < 16 || margin_luma > 1024 || (margin_luma & 15)) return c->fail("download_extended: margin must be a multiple of 16 in 16..1024");
  cudaSetDevice(c->seq.device);
  const Geom& g = c->g;
  const SlotDev& sd = c->slotsPinned[slot];
  ExtendArgs a{};
  a.ncomp = g.ncomp;
  size_t elems = 0, off[3];
  int rows = 0, maxGroups = 0;
  for (int k = 0; k < g.ncomp; k++)
  {
    if (!plane[k]) return c->fail("download_extended: plane %d is NULL", k);
    const PlaneDev& d = sd.buf[c->cur[slot]][k];
    a.src[k] = d.p; a.w[k] = d.w; a.h[k] = d.h; a.srcPitch[k] = d.pitch;
    a.xm[k] = margin_luma >> (k ? g.sx : 0); a.ym[k] = margin_luma >> (k ? g.sy : 0);
    a.dstPitch[k] = (int)alignUp(d.w + 2 * a.xm[k], 8);
    a.rowStart[k] = rows;
    rows += d.h + 2 * a.ym[k];
    maxGroups = std::max(maxGroups, a.dstPitch[k] / 8);
    off[k] = elems;
    elems += alignUp((size_t)a.dstPitch[k] * (d.h + 2 * a.ym[k]), 64);
  }
  a.rowStart[g.ncomp] = rows;
  if (elems > c->extendElems)
  {
    // picture + margins of one slot, allocated by the first call
    if (c->extendBuf) { cudaStreamSynchronize(c->stream); cudaFree(c->extendBuf); c->extendBuf = nullptr; c->extendElems = 0; }
    if (c->cuda(cudaMalloc((void**)&c->extendBuf, elems * sizeof(pel)), "padded picture allocation")) return -1;
    c->extendElems = elems;
  }
  for (int k = 0; k < g.ncomp; k++) a.dst[k] = c->extendBuf + off[k];
  k_extend_border<<<dim3((maxGroups + 255) / 256, rows), 256, 0, c->stream>>>(a);
  c->launches++;
  if (c->cuda(cudaGetLastError(), "k_extend_border launch")) return -1;
  for (int k = 0; k < g.ncomp; k++)
    if (c->cuda(cudaMemcpy2DAsync(plane[k] - (ptrdiff_t)a.ym[k] * stride[k] - a.xm[k], (size_t)stride[k] * 2, a.dst[k], (size_t)a.dstPitch[k] * 2,
                                  (size_t)(a.w[k] + 2 * a.xm[k]) * 2, a.h[k] + 2 * a.ym[k], cudaMemcpyDeviceToHost, c->stream), "download_extended")) return -1;
  return vtmgpu_sync(c);
}

extern "C" int vtmgpu_set_lmcs(vtmgpu_ctx* c, int slot, const int16_t* inv_lut, int entries)
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("set_lmcs: bad slot %d", slot);
  cudaSetDevice(c->seq.device);
  c->mirrorWrite(slot);
  SlotDev& sd = c->slotsPinned[slot];
  sd.lmcsOn = 0;
  if (inv_lut)
  {
    if (entries != 1 << c->g.bdL) { c->pushSlot(slot); return c->fail("set_lmcs: expected %d table entries, got %d", 1 << c->g.bdL, entries); }
    memcpy(c->pinnedSide(slot) + c->lay.lmcs, inv_lut, (size_t)entries * sizeof(int16_t));
    c->markSide(slot, c->lay.lmcs, (size_t)entries * sizeof(int16_t));
    sd.lmcsOn = 1;
  }
  return c->pushSlot(slot);
}

// ------------------------------------------------------------------------------------------------------------
// band mode: one picture split into CTU-row bands over several contexts (GPUs)
// ------------------------------------------------------------------------------------------------------------
extern "C" int vtmgpu_set_stream(vtmgpu_ctx* c, void* cuda_stream, int async_stages)
{
  if (!c) return -1;
  cudaSetDevice(c->seq.device);
  if (c->cuda(cudaStreamSynchronize(c->stream), "set_stream")) return -1;      // drain the stream being left
  c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->ownStream;
  c->asyncStages = async_stages != 0;
  return 0;
}

extern "C" int vtmgpu_set_rows(vtmgpu_ctx* c, int y_begin, int y_end)
{
  if (!c) return -1;
  const int h = c->g.h;
  if (y_begin < 0 || y_end > h || y_begin >= y_end || (y_begin & 127) || ((y_end & 127) && y_end != h))
    return c->fail("set_rows: [%d,%d) must be a non-empty range of multiples of 128 luma rows (the end may be the picture height %d)", y_begin, y_end, h);
  c->rowBegin = y_begin; c->rowEnd = y_end;
  return 0;
}

namespace
{
// rows [y0, y0+n) of one plane of a slot <-> host or device memory (pitch in samples), on the ctx stream
int copyRows(vtmgpu_ctx* c, int slot, int comp, int buf, int y0, int n, void* mem, ptrdiff_t stride, cudaMemcpyKind kind, bool toSlot, const char* what)
{
  if (!c->slotOk(slot, 1)) return c->fail("%s: bad slot %d", what, slot);
  if (comp < 0 || comp >= c->g.ncomp) return c->fail("%s: bad component %d", what, comp);
  const PlaneDev& d = c->slotsPinned[slot].buf[buf][comp];
  if (y0 < 0 || n < 0 || y0 + n > d.h) return c->fail("%s: rows [%d,%d) outside the plane (height %d)", what, y0, y0 + n, d.h);
  if (!mem && n) return c->fail("%s: NULL buffer", what);
  if (n == 0) return 0;
  cudaSetDevice(c->seq.device);
  pel* dev = d.p + (size_t)y0 * d.pitch;
  return c->cuda(toSlot ? cudaMemcpy2DAsync(dev, (size_t)d.pitch * 2, mem, (size_t)stride * 2, (size_t)d.w * 2, n, kind, c->stream)
                       : cudaMemcpy2DAsync(mem, (size_t)stride * 2, dev, (size_t)d.pitch * 2, (size_t)d.w * 2, n, kind, c->stream), what);
}
}   // namespace

extern "C" int vtmgpu_upload_rows(vtmgpu_ctx* c, int slot, const int16_t* const plane[3], const ptrdiff_t stride[3], int y_begin, int y_end)
{
  if (!c) return -1;
  if (y_begin < 0 || y_end > c->g.h || y_begin > y_end || ((y_begin | y_end) & 1)) return c->fail("upload_rows: bad row range [%d,%d)", y_begin, y_end);
  for (int k = 0; k < c->g.ncomp; k++)
  {
    const int sy = k ? c->g.sy : 0;
    if (!plane[k]) return c->fail("upload_rows: plane %d is NULL", k);
    // plane[k] points at row 0 of the FULL picture plane; only the requested rows are read
    if (copyRows(c, slot, k, 0, y_begin >> sy, (y_end - y_begin) >> sy, const_cast<int16_t*>(plane[k]) + (size_t)(y_begin >> sy) * stride[k], stride[k],
                 cudaMemcpyHostToDevice, true, "upload_rows")) return -1;
  }
  c->cur[slot] = 0;
  return c->cuda(cudaStreamSynchronize(c->stream), "upload_rows");
}

extern "C" int vtmgpu_download_rows(vtmgpu_ctx* c, int slot, int16_t* const plane[3], const ptrdiff_t stride[3], int y_begin, int y_end)
{
  if (!c) return -1;
  if (y_begin < 0 || y_end > c->g.h || y_begin > y_end || ((y_begin | y_end) & 1)) return c->fail("download_rows: bad row range [%d,%d)", y_begin, y_end);
  for (int k = 0; k < c->g.ncomp; k++)
  {
    const int sy = k ? c->g.sy : 0;
    if (!plane[k]) return c->fail("download_rows: plane %d is NULL", k);
    if (copyRows(c, slot, k, c->cur[slot], y_begin >> sy, (y_end - y_begin) >> sy, plane[k] + (size_t)(y_begin >> sy) * stride[k], stride[k],
                 cudaMemcpyDeviceToHost, false, "download_rows")) return -1;
  }
  return c->cuda(cudaStreamSynchronize(c->stream), "download_rows");
}

// halo exchange between neighbouring bands: rows of the CURRENT state of a plane to / from a dense device buffer
// (width x nrows int16) that the caller sends over NVLink (NCCL send/recv, peer copy)
extern "C" int vtmgpu_export_rows(vtmgpu_ctx* c, int slot, int comp, int y0, int nrows, void* dev_dst)
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("export_rows: bad slot %d", slot);
  if (comp < 0 || comp >= c->g.ncomp) return c->fail("export_rows: bad component %d", comp);
  if (copyRows(c, slot, comp, c->cur[slot], y0, nrows, dev_dst, c->slotsPinned[slot].buf[0][comp].w, cudaMemcpyDeviceToDevice, false, "export_rows")) return -1;
  return c->asyncStages ? 0 : c->cuda(cudaStreamSynchronize(c->stream), "export_rows");
}

extern "C" int vtmgpu_import_rows(vtmgpu_ctx* c, int slot, int comp, int y0, int nrows, const void* dev_src)
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("import_rows: bad slot %d", slot);
  if (comp < 0 || comp >= c->g.ncomp) return c->fail("import_rows: bad component %d", comp);
  if (copyRows(c, slot, comp, c->cur[slot], y0, nrows, const_cast<void*>(dev_src), c->slotsPinned[slot].buf[0][comp].w, cudaMemcpyDeviceToDevice, true, "import_rows")) return -1;
  return c->asyncStages ? 0 : c->cuda(cudaStreamSynchronize(c->stream), "import_rows");
}

// all components at once: nrows rows of every plane, starting at row y[comp] of component comp, packed one plane after the other
// (luma, Cb, Cr; each width x nrows int16) -- one buffer and one message per band border and direction
extern "C" int vtmgpu_export_halo(vtmgpu_ctx* c, int slot, const int y[3], int nrows, void* dev_dst)
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("export_halo: bad slot %d", slot);
  if (!y || (!dev_dst && nrows)) return c->fail("export_halo: NULL argument");
  int16_t* d = static_cast<int16_t*>(dev_dst);
  for (int k = 0; k < c->g.ncomp; k++)
  {
    const int w = c->slotsPinned[slot].buf[0][k].w;
    if (copyRows(c, slot, k, c->cur[slot], y[k], nrows, d, w, cudaMemcpyDeviceToDevice, false, "export_halo")) return -1;
    d += (size_t)w * nrows;
  }
  return c->asyncStages ? 0 : c->cuda(cudaStreamSynchronize(c->stream), "export_halo");
}

extern "C" int vtmgpu_import_halo(vtmgpu_ctx* c, int slot, const int y[3], int nrows, const void* dev_src)
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("import_halo: bad slot %d", slot);
  if (!y || (!dev_src && nrows)) return c->fail("import_halo: NULL argument");
  int16_t* d = static_cast<int16_t*>(const_cast<void*>(dev_src));
  for (int k = 0; k < c->g.ncomp; k++)
  {
    const int w = c->slotsPinned[slot].buf[0][k].w;
    if (copyRows(c, slot, k, c->cur[slot], y[k], nrows, d, w, cudaMemcpyDeviceToDevice, true, "import_halo")) return -1;
    d += (size_t)w * nrows;
  }
  return c->asyncStages ? 0 : c->cuda(cudaStreamSynchronize(c->stream), "import_halo");
}

extern "C" int vtmgpu_rewind(vtmgpu_ctx* c, int first, int count)
{
  if (!c) return -1;
  if (!c->slotOk(first, count)) return c->fail("rewind: bad slot range");
  for (int s = first; s < first + count; s++) c->cur[s] = 0;   // buf[0] is never written by a kernel
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// side information
// ------------------------------------------------------------------------------------------------------------
namespace
{
// the records of the edges ON the picture border (column 0 of the vertical-edge arrays, row 0 of the horizontal-edge arrays)
// are never filtered (no neighbour, LoopFilter.cpp:918-933): cleared on the device so that the kernel needs no position test
int clearBorderRecords(vtmgpu_ctx* c, int slot)
{
  const SideLayout& L = c->lay;
  unsigned char* d = c->sideDev[slot];
  if (c->cuda(cudaMemset2DAsync(d + L.dbfL[0], (size_t)L.recP[0] * 4, 0, 4, L.recH[0], c->stream), "record clear")) return -1;
  if (c->cuda(cudaMemsetAsync(d + L.dbfL[1], 0, (size_t)L.recW[1] * 4, c->stream), "record clear")) return -1;
  if (L.recW[2])
  {
    if (c->cuda(cudaMemset2DAsync(d + L.dbfC[0], (size_t)L.recP[2] * 8, 0, 8, L.recH[2], c->stream), "record clear")) return -1;
    if (c->cuda(cudaMemsetAsync(d + L.dbfC[1], 0, (size_t)L.recW[3] * 8, c->stream), "record clear")) return -1;
  }
  return 0;
}

// device arrays keep the ABI indexing but with a row pitch that is a multiple of 16 bytes (TMA).  direct = copy straight
// from the caller's arrays (asynchronous when they are page-locked), else through the context's pinned staging block.
// the per-tile queues of a slot from its record arrays (enqueued behind whatever wrote the arrays)
int buildDbfQueues(vtmgpu_ctx* c, int slot)
{
  const SideLayout& L = c->lay;
  const SlotDev& sd = c->slotsPinned[slot];
  DbfQueueArgs A{};
  for (int d = 0; d < 2; d++) { A.lumaRec[d] = sd.dbfL[d]; A.chromaRec[d] = sd.dbfC[d]; }
  for (int a = 0; a < 4; a++) { A.recW[a] = L.recW[a]; A.recH[a] = L.recH[a]; A.recP[a] = L.recP[a]; }
  A.q = const_cast<uint64_t*>(sd.dbfQ); A.cnt = const_cast<uint32_t*>(sd.dbfQCnt);
  A.L = c->dbfFull; A.sx = c->g.sx; A.sy = c->g.sy; A.ncomp = c->g.ncomp;
  const int warps = 2 * (A.L.tilesLFull + (A.ncomp > 1 ? 2 * A.L.tilesCFull : 0));
  k_dbf_queues<<<(warps * 32 + 255) / 256, 256, 0, c->stream>>>(A);
  c->launches++;
  return c->cuda(cudaGetLastError(), "k_dbf_queues launch");
}

int setLadf(vtmgpu_ctx* c, int slot, const vtmgpu_ladf* l)
{
  LadfDev& d = c->slotsPinned[slot].ladf;
  d = LadfDev{};
  if (!l) return 0;
  if (l->num_intervals < 2 || l->num_intervals > 5) return c->fail("set_deblock: LADF needs 2..5 intervals, got %d", l->num_intervals);
  d.n = l->num_intervals;
  for (int k = 0; k < d.n; k++) { d.off[k] = l->qp_offset[k]; d.lb[k] = l->lower_bound[k]; }
  return 0;
}

int setDeblock(vtmgpu_ctx* c, int slot, const vtmgpu_deblock_params* p, bool direct)
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("set_deblock: bad slot %d", slot);
  cudaSetDevice(c->seq.device);
  c->mirrorWrite(slot);
  SlotDev& sd = c->slotsPinned[slot];
  sd.dbfOn = p != nullptr;
  c->pushSlot(slot);
  if (!c->sparsePending.empty()) c->sparsePending[slot] = 0;      // record lists given earlier are superseded
  if (setLadf(c, slot, p ? p->ladf : nullptr)) return -1;
  if (p)
  {
    const SideLayout& L = c->lay;
    for (int a = 0; a < 4; a++)
    {
      if (!L.recW[a]) continue;
      const int d = a & 1, es = a < 2 ? 4 : 8;
      const void* src = a < 2 ? (const void*)p->luma[d] : (const void*)p->chroma[d];
      const size_t off = a < 2 ? L.dbfL[d] : L.dbfC[d];
      if (a < 2 && !src) return c->fail("set_deblock: luma[%d] is NULL", d);
      if (!src)
      {
        if (c->cuda(cudaMemsetAsync(c->sideDev[slot] + off, 0, (size_t)L.recP[a] * L.recH[a] * es, c->stream), "set_deblock")) return -1;
        continue;
      }
      if (direct)
      {
        if (c->cuda(cudaMemcpy2DAsync(c->sideDev[slot] + off, (size_t)L.recP[a] * es, src, (size_t)L.recW[a] * es, (size_t)L.recW[a] * es, L.recH[a],
                                      cudaMemcpyHostToDevice, c->stream), "set_deblock")) return -1;
      }
      else
      {
        unsigned char* stage = c->pinnedSide(slot) + off;
        if (L.recP[a] == L.recW[a]) memcpy(stage, src, (size_t)L.recW[a] * L.recH[a] * es);
        else for (int r = 0; r < L.recH[a]; r++) memcpy(stage + (size_t)r * L.recP[a] * es, (const unsigned char*)src + (size_t)r * L.recW[a] * es, (size_t)L.recW[a] * es);
        if (c->pushSide(slot, off, (size_t)L.recP[a] * L.recH[a] * es)) return -1;
      }
    }
    if (clearBorderRecords(c, slot)) return -1;
    if (buildDbfQueues(c, slot)) return -1;
  }
  return c->pushSlot(slot);
}
}   // namespace

// scatter of the record lists into the (zeroed) record arrays of a slot: one thread per list entry.  Entries on the picture
// border (column 0 of the vertical-edge arrays, row 0 of the horizontal-edge arrays) are dropped like clearBorderRecords does.

__global__ void __launch_bounds__(256) k_clear16(uint4* __restrict__ p, size_t n)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(256) k_dbf_scatter(ScatterArgs A)
{
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < A.first[4]; t += gridDim.x * blockDim.x)
  {
    const int a = t >= A.first[2] ? (t >= A.first[3] ? 3 : 2) : (t >= A.first[1] ? 1 : 0);
    const uint32_t k = t - A.first[a];
    uint32_t index;
    uint64_t rec;
    if (a < 2)
    {
      const uint2 e = reinterpret_cast<const uint2*>(A.list[a])[k];
      index = e.x; rec = e.y;
    }
    else
    {
      const uint4 e = reinterpret_cast<const uint4*>(A.list[a])[k];
      rec = (uint64_t)e.x | (uint64_t)e.y << 32; index = e.z;
    }
    const uint32_t row = index / (uint32_t)A.recW[a], col = index - row * (uint32_t)A.recW[a];
    if (row >= (uint32_t)A.recH[a] || ((a & 1) ? row == 0 : col == 0)) continue;
    const size_t o = (size_t)row * A.recP[a] + col;
    if (a < 2) reinterpret_cast<uint32_t*>(A.dense[a])[o] = (uint32_t)rec;
    else       reinterpret_cast<uint64_t*>(A.dense[a])[o] = rec;
  }
}

// Device-side derivation of the deblocking records: one thread per 4x4 unit and edge direction evaluates
// include/vtmgpu_derive.h on the flattened block structure and stores the unit's luma record and, where the unit sits on the chroma
// edge grid, the chroma record -- every position of the four record arrays is written, no clearing pass.
struct DeriveArgs
{
  vtmgpu_derive::Ctx D;
  uint32_t* luma[2];
  uint64_t* chroma[2];
  int recP[4];
};

__global__ void __launch_bounds__(256) k_dbf_derive(const DeriveArgs A)
{
  vtmgpu_derive::Ctx D = A.D;
  D.tcTable = kDbfTcTable;
  D.betaTable = kDbfBetaTable;
  const int n = D.w4 * D.h4, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * n) return;
  const int dir = t >= n, u = t - dir * n, y = u / D.w4, x = u - y * D.w4;
  uint32_t lr; uint64_t cr; bool slot;
  vtmgpu_derive::deriveUnit(D, x, y, dir, lr, cr, slot);
  A.luma[dir][(size_t)y * A.recP[dir] + x] = lr;
  if (slot)
  {
    if (dir == 0) A.chroma[0][(size_t)y * A.recP[2] + ((4 * x) >> (3 + D.sx))] = cr;
    else          A.chroma[1][(size_t)((4 * y) >> (3 + D.sy)) * A.recP[3] + x] = cr;
  }
}

namespace
{
int uploadMotion(vtmgpu_ctx* c, int slot, const void* motion, int elemBytes, int pitch, int offMv0, int offMv1, int offRef0, int offRef1, const char* who)
{
  const Geom& g = c->g;
  if (!motion || elemBytes < 20 || pitch < g.w / 4 || (elemBytes & 3) || ((offMv0 | offMv1) & 3) || ((offRef0 | offRef1) & 1) || offMv0 < 0 || offMv1 < 0 || offRef0 < 0 || offRef1 < 0 ||
      offMv0 + 8 > elemBytes || offMv1 + 8 > elemBytes || offRef0 + 2 > elemBytes || offRef1 + 2 > elemBytes)
    return c->fail("%s: bad motion field layout", who);
  if (c->motionDev.empty()) c->motionDev.resize(c->seq.capacity);
  vtmgpu_ctx::MotionDev& m = c->motionDev[slot];
  const size_t bytes = (size_t)(g.w / 4) * (g.h / 4) * elemBytes;
  if (bytes > m.bytes)
  {
    if (m.dev) { cudaStreamSynchronize(c->stream); cudaFree(m.dev); m.dev = nullptr; m.bytes = 0; }
    if (c->cuda(cudaMalloc(&m.dev, bytes), "motion field allocation")) return -1;
    m.bytes = bytes;
  }
  // rows of the motion field, without the pitch padding
  if (c->cuda(cudaMemcpy2DAsync(m.dev, (size_t)(g.w / 4) * elemBytes, motion, (size_t)pitch * elemBytes, (size_t)(g.w / 4) * elemBytes, g.h / 4, cudaMemcpyHostToDevice, c->stream), who)) return -1;
  m.elemBytes = elemBytes; m.offMv0 = offMv0; m.offMv1 = offMv1; m.offRef0 = offRef0; m.offRef1 = offRef1;
  m.loaded = true;
  return 0;
}
}   // namespace

extern "C" int vtmgpu_upload_motion(vtmgpu_ctx* c, int slot, const void* motion, int elem_bytes, int pitch, int off_mv0, int off_mv1, int off_ref0, int off_ref1)
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("upload_motion: bad slot %d", slot);
  cudaSetDevice(c->seq.device);
  return uploadMotion(c, slot, motion, elem_bytes, pitch, off_mv0, off_mv1, off_ref0, off_ref1, "upload_motion");
}

extern "C" int vtmgpu_set_deblock_units(vtmgpu_ctx* c, int slot, const vtmgpu_deblock_units* p)
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("set_deblock_units: bad slot %d", slot);
  if (!p) return setDeblock(c, slot, nullptr, false);
  const Geom& g = c->g;
  const size_t units = (size_t)(g.w / 4) * (g.h / 4);
  if (p->num_cus < 1 || p->num_tus < 1 || p->num_slices < 1 || p->num_slices > 255 || !p->cus || !p->tus || !p->slices || !p->tu_luma) return c->fail("set_deblock_units: missing tables");
  if (g.ncomp > 1 && !p->tu_chroma) return c->fail("set_deblock_units: tu_chroma is NULL");
  const bool preloaded = (p->flags & VTMGPU_UNITS_MOTION_PRELOADED) != 0;
  if (preloaded && (c->motionDev.empty() || !c->motionDev[slot].loaded)) return c->fail("set_deblock_units: no motion field was uploaded to slot %d (vtmgpu_upload_motion)", slot);
  if (p->vb && (p->vb->num_ver < 0 || p->vb->num_ver > 3 || p->vb->num_hor < 0 || p->vb->num_hor > 3)) return c->fail("set_deblock_units: bad virtual boundaries");
  cudaSetDevice(c->seq.device);
  c->mirrorWrite(slot);
  SlotDev& sd = c->slotsPinned[slot];
  sd.dbfOn = 1;
  if (!c->sparsePending.empty()) c->sparsePending[slot] = 0;      // record lists given earlier are superseded
  if (setLadf(c, slot, p->ladf)) { sd.dbfOn = 0; c->pushSlot(slot); return -1; }
  // landing area of the tables on the device (per slot, grown on demand)
  const size_t szCu = alignUp(sizeof(vtmgpu_dbf_cu) * p->num_cus, 256), szTu = alignUp(sizeof(vtmgpu_dbf_tu) * p->num_tus, 256), szSl = alignUp(sizeof(vtmgpu_dbf_slice) * p->num_slices, 256);
  const size_t szMap = alignUp(units * 4, 256);
  const size_t need = szCu + szTu + szSl + szMap * (g.ncomp > 1 ? 2 : 1);
  if (c->unitsDev.empty()) { c->unitsDev.assign(c->seq.capacity, nullptr); c->unitsBytes.assign(c->seq.capacity, 0); }
  if (need > c->unitsBytes[slot])
  {
    if (c->unitsDev[slot]) { cudaStreamSynchronize(c->stream); cudaFree(c->unitsDev[slot]); c->unitsDev[slot] = nullptr; c->unitsBytes[slot] = 0; }
    const size_t cap = need + need / 4;
    if (c->cuda(cudaMalloc(&c->unitsDev[slot], cap), "block structure allocation")) { sd.dbfOn = 0; c->pushSlot(slot); return -1; }
    c->unitsBytes[slot] = cap;
  }
  unsigned char* d = c->unitsDev[slot];
  DeriveArgs A{};
  vtmgpu_derive::Ctx& D = A.D;
  size_t o = 0;
#define UP(dst, src, bytes) do { if (c->cuda(cudaMemcpyAsync(d + o, src, bytes, cudaMemcpyHostToDevice, c->stream), "set_deblock_units")) { sd.dbfOn = 0; c->pushSlot(slot); return -1; } dst = reinterpret_cast<decltype(dst)>(d + o); } while (0)
  UP(D.cus, p->cus, sizeof(vtmgpu_dbf_cu) * p->num_cus); o += szCu;
  UP(D.tus, p->tus, sizeof(vtmgpu_dbf_tu) * p->num_tus); o += szTu;
  UP(D.slices, p->slices, sizeof(vtmgpu_dbf_slice) * p->num_slices); o += szSl;
  UP(D.tuL, p->tu_luma, units * 4); o += szMap;
  if (g.ncomp > 1) { UP(D.tuC, p->tu_chroma, units * 4); o += szMap; }
#undef UP
  if (!preloaded && p->motion &&
      uploadMotion(c, slot, p->motion, p->motion_elem_bytes, p->motion_pitch, p->off_mv0, p->off_mv1, p->off_ref0, p->off_ref1, "set_deblock_units")) { sd.dbfOn = 0; c->pushSlot(slot); return -1; }
  if (preloaded || p->motion)
  {
    const vtmgpu_ctx::MotionDev& m = c->motionDev[slot];
    D.motion = m.dev;
    D.miBytes = m.elemBytes; D.miPitch = g.w / 4;
    D.offMv0 = m.offMv0; D.offMv1 = m.offMv1; D.offRef0 = m.offRef0; D.offRef1 = m.offRef1;
  }
  if (!c->motionDev.empty()) c->motionDev[slot].loaded = false;       // one picture's field: the next picture uploads its own
  D.w4 = g.w / 4; D.h4 = g.h / 4; D.sx = g.sx; D.sy = g.sy; D.chroma = g.ncomp > 1;
  D.bdL = g.bdL; D.bdC = g.bdC; D.ctuLog2 = g.ctuLog2;
  D.flags = p->flags;
  D.ladf = p->ladf != nullptr;
  if (p->vb)
  {
    D.nvb[0] = p->vb->num_ver; D.nvb[1] = p->vb->num_hor;
    for (int i = 0; i < 3; i++) { D.vb[0][i] = p->vb->pos_x[i]; D.vb[1][i] = p->vb->pos_y[i]; }
  }
  const SideLayout& L = c->lay;
  for (int a = 0; a < 2; a++)
  {
    A.luma[a] = reinterpret_cast<uint32_t*>(c->sideDev[slot] + L.dbfL[a]);
    A.chroma[a] = reinterpret_cast<uint64_t*>(c->sideDev[slot] + L.dbfC[a]);
  }
  for (int a = 0; a < 4; a++) A.recP[a] = L.recP[a];
  const int threads = 2 * (int)units;
  k_dbf_derive<<<(threads + 255) / 256, 256, 0, c->stream>>>(A);
  c->launches++;
  if (c->cuda(cudaGetLastError(), "k_dbf_derive launch") || buildDbfQueues(c, slot)) { sd.dbfOn = 0; c->pushSlot(slot); return -1; }
  return c->pushSlot(slot);
}

extern "C" int vtmgpu_get_deblock_records(vtmgpu_ctx* c, int slot, uint32_t* const luma[2], uint64_t* const chroma[2])
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("get_deblock_records: bad slot %d", slot);
  cudaSetDevice(c->seq.device);
  if (runPendingSparse(c, slot, 1)) return -1;
  const SideLayout& L = c->lay;
  for (int a = 0; a < 4; a++)
  {
    if (!L.recW[a]) continue;
    const int d = a & 1, es = a < 2 ? 4 : 8;
    void* dst = a < 2 ? (void*)luma[d] : (void*)chroma[d];
    if (!dst) return c->fail("get_deblock_records: NULL array");
    const size_t off = a < 2 ? L.dbfL[d] : L.dbfC[d];
    if (c->cuda(cudaMemcpy2DAsync(dst, (size_t)L.recW[a] * es, c->sideDev[slot] + off, (size_t)L.recP[a] * es, (size_t)L.recW[a] * es, L.recH[a], cudaMemcpyDeviceToHost, c->stream), "get_deblock_records")) return -1;
  }
  return vtmgpu_sync(c);
}

extern "C" int vtmgpu_set_deblock_sparse(vtmgpu_ctx* c, int slot, const vtmgpu_deblock_sparse* p)
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("set_deblock_sparse: bad slot %d", slot);
  if (!p) return setDeblock(c, slot, nullptr, true);
  cudaSetDevice(c->seq.device);
  c->mirrorWrite(slot);
  const SideLayout& L = c->lay;
  size_t offs[5] = { 0, 0, 0, 0, 0 };    // landing area: worst case = every unit listed
  for (int a = 0; a < 4; a++) offs[a + 1] = alignUp(offs[a] + (size_t)L.recW[a] * L.recH[a] * (a < 2 ? sizeof(vtmgpu_dbf_luma_entry) : sizeof(vtmgpu_dbf_chroma_entry)), 256);
  if (c->sparseDev.empty()) c->sparseDev.assign(c->seq.capacity, nullptr);
  if (!c->sparseDev[slot] && c->cuda(cudaMalloc(&c->sparseDev[slot], offs[4]), "record list allocation")) return -1;
  ScatterArgs A{};
  // the lists land packed (16-byte aligned, in array order); a producer that keeps them in ONE buffer with this very layout
  // gets a single copy instead of four
  const void* src[4];
  size_t esz[4], poff[5] = { 0, 0, 0, 0, 0 };
  for (int a = 0; a < 4; a++)
  {
    const int d = a & 1;
    src[a] = a < 2 ? (const void*)p->luma[d] : (const void*)p->chroma[d];
    esz[a] = a < 2 ? sizeof(vtmgpu_dbf_luma_entry) : sizeof(vtmgpu_dbf_chroma_entry);
    uint32_t n = a < 2 ? p->luma_count[d] : p->chroma_count[d];
    if (!L.recW[a]) n = 0;
    if (n > (uint32_t)L.recW[a] * (uint32_t)L.recH[a]) return c->fail("set_deblock_sparse: list %d has %u entries for %d units", a, n, L.recW[a] * L.recH[a]);
    if (n && !src[a]) return c->fail("set_deblock_sparse: list %d is NULL", a);
    poff[a + 1] = alignUp(poff[a] + (size_t)n * esz[a], 16);
    A.list[a] = c->sparseDev[slot] + poff[a];
    A.dense[a] = c->sideDev[slot] + (a < 2 ? L.dbfL[d] : L.dbfC[d]);
    A.count[a] = n; A.first[a + 1] = A.first[a] + n;
    A.recW[a] = L.recW[a] ? L.recW[a] : 1; A.recH[a] = L.recH[a]; A.recP[a] = L.recP[a];
  }
  int firstList = -1, lastList = -1, nLists = 0;
  bool packed = true;
  for (int a = 0; a < 4; a++)
  {
    if (!A.count[a]) continue;
    if (firstList < 0) firstList = a;
    lastList = a; nLists++;
    packed = packed && (const unsigned char*)src[a] == (const unsigned char*)src[firstList] + (poff[a] - poff[firstList]);
  }
  if (nLists >= 2 && packed)
  {
    // one copy from the start of the first list to the END OF THE LAST ENTRY (the caller owes no padding after the last list)
    const size_t bytes = poff[lastList] + (size_t)A.count[lastList] * esz[lastList] - poff[firstList];
    if (c->cuda(cudaMemcpyAsync(c->sparseDev[slot] + poff[firstList], src[firstList], bytes, cudaMemcpyHostToDevice, c->stream), "set_deblock_sparse")) return -1;
  }
  else
    for (int a = 0; a < 4; a++)
      if (A.count[a] && c->cuda(cudaMemcpyAsync(c->sparseDev[slot] + poff[a], src[a], (size_t)A.count[a] * esz[a], cudaMemcpyHostToDevice, c->stream), "set_deblock_sparse")) return -1;
  if (c->sparseArgs.empty()) { c->sparseArgs.assign(c->seq.capacity, ScatterArgs{}); c->sparsePending.assign(c->seq.capacity, 0); }
  c->sparseArgs[slot] = A;               // the kernels follow with the next stage call (runPendingSparse)
  c->sparsePending[slot] = 1;
  c->slotsPinned[slot].dbfOn = 1;
  if (setLadf(c, slot, p->ladf)) return -1;
  return c->pushSlot(slot);
}

// the device side of vtmgpu_set_deblock_sparse for the slots [first, first + count): clear the record arrays, scatter the lists,
// build the tile queues
static int runPendingSparse(vtmgpu_ctx* c, int first, int count)
{
  if (c->sparsePending.empty()) return 0;
  const SideLayout& L = c->lay;
  for (int slot = first; slot < first + count; slot++)
  {
    if (!c->sparsePending[slot]) continue;
    c->sparsePending[slot] = 0;
    const ScatterArgs& A = c->sparseArgs[slot];
    // cleared by a kernel, not by cudaMemsetAsync: the driver hands a memset of this size to a copy engine, where it sits in line
    // with the plane copies of the other pictures (measured: 5.6 of 42 ms per 64-picture batch of vtmgpu_batch_filter)
    {
      uint4* const z = reinterpret_cast<uint4*>(c->sideDev[slot] + L.dbfL[0]);
      const size_t n16 = (L.dbfEnd - L.dbfL[0]) / 16;
      k_clear16<<<(int)std::min<size_t>((n16 + 255) / 256, 8u * c->numSms), 256, 0, c->stream>>>(z, n16);
      if (c->cuda(cudaGetLastError(), "k_clear16 launch")) return -1;
      c->launches++;
    }
    if (A.first[4])
    {
      const int grid = (int)std::min<uint32_t>((A.first[4] + 255) / 256, 4u * c->numSms);
      k_dbf_scatter<<<grid, 256, 0, c->stream>>>(A);
      if (c->cuda(cudaGetLastError(), "k_dbf_scatter launch")) return -1;
      c->launches++;
    }
    if (buildDbfQueues(c, slot)) return -1;
  }
  return 0;
}

extern "C" int vtmgpu_set_deblock(vtmgpu_ctx* c, int slot, const vtmgpu_deblock_params* p) { return setDeblock(c, slot, p, false); }
extern "C" int vtmgpu_set_deblock_async(vtmgpu_ctx* c, int slot, const vtmgpu_deblock_params* p) { return setDeblock(c, slot, p, true); }

extern "C" int vtmgpu_sao_reconstruct(vtmgpu_sao_ctu* ctu, int num_ctus, int width_in_ctus, int num_comps, int log2_scale_luma, int log2_scale_chroma)
{
  // xReconstructBlkSAOParams (SampleAdaptiveOffset.cpp:266-290): raster order so that a MERGE candidate is already final;
  // NEW: invertQuantOffsets (:148-171); MERGE: copy of the left / above CTU's component parameters (:250-257)
  if (!ctu || num_ctus < 0 || width_in_ctus < 1 || num_comps < 1 || num_comps > 3) return -1;
  int mask = 0;
  for (int a = 0; a < num_ctus; a++)
    for (int k = 0; k < num_comps; k++)
    {
      vtmgpu_sao_offset& o = ctu[a].comp[k];
      switch (o.mode)
      {
      case VTMGPU_SAO_MODE_OFF: continue;
      case VTMGPU_SAO_MODE_NEW:
      {
        const int mul = 1 << (k == 0 ? log2_scale_luma : log2_scale_chroma);
        if (o.type == VTMGPU_SAO_BO)
        {
          int16_t keep[4];
          for (int i = 0; i < 4; i++) keep[i] = (int16_t)(o.offset[(o.aux + i) & 31] * mul);
          memset(o.offset, 0, sizeof(o.offset));
          for (int i = 0; i < 4; i++) o.offset[(o.aux + i) & 31] = keep[i];
        }
        else if (o.type >= VTMGPU_SAO_EO_0 && o.type <= VTMGPU_SAO_EO_45)
        {
          for (int i = 0; i < 32; i++) o.offset[i] = i < 5 ? (int16_t)(o.offset[i] * mul) : 0;
          if (o.offset[2] != 0) return -2;       // "EO offset is not '0'"
        }
        else return -3;
        break;
      }
      case VTMGPU_SAO_MODE_MERGE:
      {
        int from;
        if (o.type == VTMGPU_SAO_MERGE_LEFT && ctu[a].merge_left_ok && (a % width_in_ctus) > 0) from = a - 1;
        else if (o.type == VTMGPU_SAO_MERGE_ABOVE && ctu[a].merge_above_ok && a >= width_in_ctus) from = a - width_in_ctus;
        else return -4;                          // "Merge target does not exist"
        o = ctu[from].comp[k];
        break;
      }
      default: return -5;
      }
      if (o.mode != VTMGPU_SAO_MODE_OFF) mask |= 1 << k;
    }
  return mask;
}

namespace
{
// validates and copies signalled virtual boundaries (multiples of 8 luma samples strictly inside the picture, at least one CTU
// apart in each direction -- the kernels rely on at most one cut per tile and direction)
int setVb(vtmgpu_ctx* c, VbDev& d, const vtmgpu_virtual_boundaries* v, const char* who)
{
  d = VbDev{};
  if (!v) return 0;
  if (v->num_ver < 0 || v->num_ver > 3 || v->num_hor < 0 || v->num_hor > 3) return c->fail("%s: 0..3 virtual boundaries per direction", who);
  for (int dir = 0; dir < 2; dir++)
  {
    const int n = dir ? v->num_hor : v->num_ver, lim = dir ? c->g.h : c->g.w;
    const int32_t* pos = dir ? v->pos_y : v->pos_x;
    for (int i = 0; i < n; i++)
    {
      if (pos[i] <= 0 || pos[i] >= lim || (pos[i] & 7)) return c->fail("%s: virtual boundary %d is not a multiple of 8 inside the picture", who, pos[i]);
      for (int j = 0; j < i; j++)
        if (std::abs(pos[i] - pos[j]) < c->g.ctu) return c->fail("%s: virtual boundaries %d and %d are less than a CTU apart", who, pos[j], pos[i]);
      (dir ? d.y : d.x)[i] = pos[i];
    }
  }
  d.nv = v->num_ver; d.nh = v->num_hor;
  return 0;
}
}   // namespace

extern "C" int vtmgpu_set_sao(vtmgpu_ctx* c, int slot, const vtmgpu_sao_params* p)
{
  if (!c) return -1;
  if (!c->slotOk(slot, 1)) return c->fail("set_sao: bad slot %d", slot);
  cudaSetDevice(c->seq.device);
  c->mirrorWrite(slot);
  SlotDev& sd = c->slotsPinned[slot];
  sd.saoOn = 0;
  c->pushSlot(slot);                     // also on the error returns below: the device must not keep a stale "on" with half-updated parameters
  if (setVb(c, sd.vbSao, p ? p->vb : nullptr, "set_sao")) return -1;
  if (p)
  {
    if (!p->ctu || p->num_ctus != c->nCtus) return c->fail("set_sao: expected %d CTUs, got %d", c->nCtus, p->num_ctus);
    SaoDev* d = reinterpret_cast<SaoDev*>(c->pinnedSide(slot) + c->lay.sao);
    int any = 0;
    for (int a = 0; a < c->nCtus; a++)
      for (int k = 0; k < 3; k++)
      {
        SaoDev z{};
        const vtmgpu_sao_offset& o = p->ctu[a].comp[k];
        if (k < c->g.ncomp && o.mode != VTMGPU_SAO_MODE_OFF)
        {
          if (o.mode != VTMGPU_SAO_MODE_NEW) return c->fail("set_sao: CTU %d comp %d is not reconstructed (mode %d)", a, k, o.mode);
          if (o.type == VTMGPU_SAO_BO)
          {
            z.type = 5;
            z.band = (uint8_t)(o.aux & 31);
            for (int i = 0; i < 4; i++) z.off[i] = o.offset[(o.aux + i) & 31];
          }
          else if (o.type >= 0 && o.type <= 3)
          {
            z.type = (uint8_t)(1 + o.type);
            for (int i = 0; i < 5; i++) z.off[i] = o.offset[i];
          }
          else return c->fail("set_sao: CTU %d comp %d bad type %d", a, k, o.type);
          z.avail = p->ctu[a].avail;
          any = 1;
        }
        d[a * 3 + k] = z;
      }
    sd.saoOn = any;     // SAOProcess returns early when no CTU has SAO on (SampleAdaptiveOffset.cpp:626-637)
    c->markSide(slot, c->lay.sao, (size_t)c->nCtus * 3 * sizeof(SaoDev));
  }
  return c->pushSlot(slot);
}

namespace
{

// what makes two slices of a picture share their ALF tables (vtmgpu_set_alf_slices): everything a slice signals
bool sameLumaAps(const vtmgpu_alf_luma_aps& a, const vtmgpu_alf_luma_aps& b)
{
  if (a.num_filters != b.num_filters || (a.nonlinear != 0) != (b.nonlinear != 0)) return false;
  if (memcmp(a.delta_idx, b.delta_idx, sizeof(a.delta_idx)) || memcmp(a.coeff, b.coeff, sizeof(a.coeff))) return false;
  return !a.nonlinear || memcmp(a.clip_idx, b.clip_idx, sizeof(a.clip_idx)) == 0;
}

bool sameSliceAlf(const vtmgpu_alf_params& a, const vtmgpu_alf_params& b, bool chroma)
{
  for (int k = 0; k < 3; k++) if ((a.enabled[k] != 0) != (b.enabled[k] != 0)) return false;
  if (a.num_luma_aps != b.num_luma_aps) return false;
  for (int i = 0; i < a.num_luma_aps; i++) if (!sameLumaAps(a.luma_aps[i], b.luma_aps[i])) return false;
  if ((a.chroma_aps != nullptr) != (b.chroma_aps != nullptr)) return false;
  if (a.chroma_aps && memcmp(a.chroma_aps, b.chroma_aps, sizeof(*a.chroma_aps))) return false;
  for (int k = 0; k < 2 && chroma; k++) if ((a.ccalf_enabled[k] != 0) != (b.ccalf_enabled[k] != 0)) return false;
  return true;
}

// ALF side information of one picture.  slices[0 .. ns) = the parameter sets of its slices, ctuSlice = slice index per CTU (NULL
// with one slice); per-picture data (CTU arrays, clip flags, virtual boundaries, CC-ALF coefficients) come from slices[0].
int setAlf(vtmgpu_ctx* c, int slot, int ns, const vtmgpu_alf_params* slices, const uint8_t* ctuSlice, const char* who)
{
  if (!c->slotOk(slot, 1)) return c->fail("%s: bad slot %d", who, slot);
  cudaSetDevice(c->seq.device);
  c->mirrorWrite(slot);
  SlotDev& sd = c->slotsPinned[slot];
  sd.alfOn = 0;
  c->pushSlot(slot);                     // also on the error returns below
  const vtmgpu_alf_params* p = slices;
  if (setVb(c, sd.vbAlf, p ? p->vb : nullptr, who)) return -1;
  if (!p)
  {
    // ALF off for this picture: k_alf reads nothing but the control records, so their flags are cleared
    memset(c->pinnedSide(slot) + c->lay.ctuCtl, 0, (size_t)c->nCtus * sizeof(CtuCtlDev));
    c->markSide(slot, c->lay.ctuCtl, (size_t)c->nCtus * sizeof(CtuCtlDev));
    return c->pushSlot(slot);
  }
  const int n = c->nCtus;
  if (p->num_ctus != n) return c->fail("%s: expected %d CTUs, got %d", who, n, p->num_ctus);
  if (ns > 1 && !ctuSlice) return c->fail("%s: ctu_slice is NULL", who);
  const bool chroma = c->g.ncomp > 1;
  // slices -> groups of equal parameters ; luma APSs of all groups -> one list of distinct sets
  std::vector<int> grpOf(ns), grpRep;
  for (int s = 0; s < ns; s++)
  {
    const vtmgpu_alf_params& q = slices[s];
    if (q.num_luma_aps < 0 || q.num_luma_aps > VTMGPU_ALF_MAX_APS) return c->fail("%s: bad num_luma_aps", who);
    if (q.num_luma_aps && !q.luma_aps) return c->fail("%s: luma_aps is NULL", who);
    int gi = -1;
    for (size_t k = 0; k < grpRep.size() && gi < 0; k++) if (sameSliceAlf(slices[grpRep[k]], q, chroma)) gi = (int)k;
    if (gi < 0)
    {
      if ((int)grpRep.size() == ALF_MAX_GROUPS) return c->fail("%s: more than %d distinct ALF parameter sets in one picture", who, ALF_MAX_GROUPS);
      gi = (int)grpRep.size();
      grpRep.push_back(s);
    }
    grpOf[s] = gi;
  }
  const int ng = (int)grpRep.size();
  std::vector<const vtmgpu_alf_luma_aps*> sets;                       // distinct luma APSs of the picture
  int setOf[ALF_MAX_GROUPS][VTMGPU_ALF_MAX_APS];                      // group, position in its list -> picture-global APS set
  for (int gi = 0; gi < ng; gi++)
  {
    const vtmgpu_alf_params& q = slices[grpRep[gi]];
    for (int i = 0; i < q.num_luma_aps; i++)
    {
      int si = -1;
      for (size_t k = 0; k < sets.size() && si < 0; k++) if (sameLumaAps(*sets[k], q.luma_aps[i])) si = (int)k;
      if (si < 0)
      {
        if ((int)sets.size() == VTMGPU_ALF_MAX_APS) return c->fail("%s: more than %d distinct luma APSs in one picture", who, VTMGPU_ALF_MAX_APS);
        si = (int)sets.size();
        sets.push_back(&q.luma_aps[i]);
      }
      setOf[gi][i] = si;
    }
  }
  AlfDev* const AG = reinterpret_cast<AlfDev*>(c->pinnedSide(slot) + c->lay.alf);
  CtuCtlDev* ctl = reinterpret_cast<CtuCtlDev*>(c->pinnedSide(slot) + c->lay.ctuCtl);
  memset(AG, 0, sizeof(AlfDev) * ng);
  // coefficient tables: reconstructCoeff (AdaptiveLoopFilter.cpp:651-713), clip values :743-762, fixed sets :792-807
  const int bdL = c->g.bdL, bdC = c->g.bdC;
  const int clipL[4] = { 1 << bdL, 1 << (bdL - 3), 1 << (bdL - 5), 1 << (bdL - 7) };
  const int clipC[4] = { 1 << bdC, 1 << (bdC - 3), 1 << (bdC - 5), 1 << (bdC - 7) };
  AlfDev& A0 = AG[0];                    // the luma sets of the picture live in the first group's record
  A0.numSets = VTMGPU_ALF_FIXED_SETS + (int)sets.size();
  for (int s = 0; s < VTMGPU_ALF_FIXED_SETS; s++)
    for (int cl = 0; cl < 25; cl++)
      for (int k = 0; k < 12; k++) A0.luma[s][cl][k] = make_short2(vvc_alf_fix_coeff[vvc_alf_class_to_filt[s * 25 + cl] * 12 + k], (short)clipL[0]);
  for (size_t s = 0; s < sets.size(); s++)
  {
    const vtmgpu_alf_luma_aps& a = *sets[s];
    for (int cl = 0; cl < 25; cl++)
    {
      const int f = a.delta_idx[cl];
      if (f < 0 || f >= a.num_filters || f >= 25) return c->fail("%s: bad coeff delta idx in a luma APS", who);
      for (int k = 0; k < 12; k++)
      {
        const int ci = a.nonlinear ? a.clip_idx[f][k] : 0;
        if (ci < 0 || ci > 3) return c->fail("%s: bad clip idx in a luma APS", who);
        A0.luma[VTMGPU_ALF_FIXED_SETS + s][cl][k] = make_short2(a.coeff[f][k], (short)clipL[ci]);
      }
    }
  }
  AlfLumaEntry* tab = reinterpret_cast<AlfLumaEntry*>(c->pinnedSide(slot) + c->lay.alfTab);
  bool wide = false;
  for (size_t s = 0; s < sets.size(); s++) wide |= expandLumaSet(A0.luma[VTMGPU_ALF_FIXED_SETS + s], tab + (VTMGPU_ALF_FIXED_SETS + s) * 100);
  int numAlts[ALF_MAX_GROUPS];
  bool anyOn = false;
  for (int gi = 0; gi < ng; gi++)
  {
    const vtmgpu_alf_params& q = slices[grpRep[gi]];
    AlfDev& A = AG[gi];
    for (int k = 0; k < 3; k++) A.enabled[k] = q.enabled[k] != 0;
    anyOn |= (A.enabled[0] | A.enabled[1] | A.enabled[2]) != 0;
    A.numSets = VTMGPU_ALF_FIXED_SETS + (gi ? q.num_luma_aps : (int)sets.size());
    A.wide = wide;
    numAlts[gi] = 0;
    if (q.chroma_aps)
    {
      const int na = numAlts[gi] = q.chroma_aps->num_alts;
      if (na < 1 || na > VTMGPU_ALF_MAX_ALTS) return c->fail("%s: bad number of chroma alternatives", who);
      for (int a = 0; a < na; a++)
      {
        AlfChromaEntry& e = A.chromaTab[a];
        int bias = 64;
        for (int k = 0; k < 6; k++)
        {
          const int ci = q.chroma_aps->nonlinear ? q.chroma_aps->clip_idx[a][k] : 0;
          if (ci < 0 || ci > 3) return c->fail("%s: bad chroma clip idx", who);
          const int co = q.chroma_aps->coeff[a][k], clp = clipC[ci];      // AlfCoeffC is restricted to -127..127 by the parser (VLCReader.cpp:3840)
          if (co < -128 || co > 127) return c->fail("%s: chroma coefficient %d out of range", who, co);
          A.chroma[a][k] = make_short2((short)co, (short)clp);
          e.coefB[k] = (uint32_t)(co & 0xff) * 0x01000001u;
          e.clipP1[k] = (uint32_t)((clp + 1) & 0xffff) * 0x10001u;
          e.clip2[k] = (uint32_t)((2 * clp) & 0xffff) * 0x10001u;
          bias -= co * 2 * clp;
        }
        e.bias = bias;
      }
    }
    for (int k = 0; k < 2; k++)
    {
      A.ccEnabled[k] = chroma && q.ccalf_enabled[k];
      memcpy(A.cc[k], p->ccalf_coeff[k], sizeof(A.cc[k]));               // the coefficients are per picture (see vtmgpu.h)
      for (int f = 0; f < VTMGPU_CCALF_MAX_FILTERS; f++)
      {
        int sum = 0;
        for (int t = 0; t < 7; t++)
        {
          const int co = p->ccalf_coeff[k][f][t];
          if (A.ccEnabled[k] && (co < -128 || co > 127)) return c->fail("%s: CC-ALF coefficient %d out of range", who, co);
          A.ccB[k][f][t] = (uint32_t)(co & 0xff) * 0x01000001u;
          sum += co;
        }
        A.ccB[k][f][7] = (uint32_t)sum;
        // two taps per IDP.2A (ccAlfQuadDual): (f0, 0, f6, 0), (0, f1, 0, f3), (-sum, f2, f4, f5)
        const int16_t* co = p->ccalf_coeff[k][f];
        auto b = [](int v) { return (uint32_t)(v & 0xff); };
        A.ccK[k][f][0] = b(co[0]) | b(co[6]) << 16;
        A.ccK[k][f][1] = b(co[1]) << 8 | b(co[3]) << 24;
        A.ccK[k][f][2] = b(-sum) | b(co[2]) << 8 | b(co[4]) << 16 | b(co[5]) << 24;
        A.ccK[k][f][3] = -sum >= -128 && -sum <= 127;
      }
    }
  }
  // per-CTU control: a CTU is skipped when its slice runs no ALF at all (ALFProcess :429), the CC-ALF test reads the slice's own
  // enable flag (:451, :532), the filter set index counts in the slice's own APS list (:436-441)
  for (int a = 0; a < n; a++)
  {
    int gi = 0;
    if (ns > 1)
    {
      if (ctuSlice[a] >= ns) return c->fail("%s: CTU %d: bad slice index %d", who, a, ctuSlice[a]);
      gi = grpOf[ctuSlice[a]];
    }
    const AlfDev& A = AG[gi];
    const vtmgpu_alf_params& q = slices[grpRep[gi]];
    CtuCtlDev& r = ctl[a];
    memset(&r, 0, sizeof(r));
    r.flags = (uint8_t)(((A.enabled[0] | A.enabled[1] | A.enabled[2]) != 0 ? 1 : 0) | (wide ? 2 : 0));
    r.grp = (uint8_t)gi;
    r.clip = p->ctu_clip ? (uint8_t)(p->ctu_clip[a] & 63) : 0;
    r.enY = p->ctu_enable[0] && p->ctu_enable[0][a];
    if (r.enY && (r.flags & 1))
    {
      const int fi = p->ctu_filter_idx ? p->ctu_filter_idx[a] : -1;
      if (fi < 0 || fi >= VTMGPU_ALF_FIXED_SETS + q.num_luma_aps) return c->fail("%s: CTU %d: bad filter set index", who, a);
      r.setIdx = (uint8_t)(fi < VTMGPU_ALF_FIXED_SETS ? fi : VTMGPU_ALF_FIXED_SETS + setOf[gi][fi - VTMGPU_ALF_FIXED_SETS]);
    }
    for (int k = 0; k < 2 && chroma; k++)
    {
      const bool on = p->ctu_enable[1 + k] && p->ctu_enable[1 + k][a];
      (k ? r.enCr : r.enCb) = on;
      if (on && (r.flags & 1))
      {
        const int alt = p->ctu_alt[k] ? p->ctu_alt[k][a] : 0;
        if (alt >= numAlts[gi]) return c->fail("%s: CTU %d: chroma alternative %d not in the APS", who, a, alt);
        (k ? r.altCr : r.altCb) = (uint8_t)alt;
      }
      if (A.ccEnabled[k])
      {
        const int idc = p->ccalf_idc[k] ? p->ccalf_idc[k][a] : 0;
        if (idc > VTMGPU_CCALF_MAX_FILTERS) return c->fail("%s: CTU %d: bad CC-ALF idc", who, a);
        (k ? r.ccCr : r.ccCb) = (uint8_t)idc;
      }
    }
  }
  sd.alfOn = anyOn;                      // ALFProcess filters nothing otherwise (AdaptiveLoopFilter.cpp:429)
  sd.alfWide = wide;
  if (!sets.empty()) c->markSide(slot, c->lay.alfTab + sizeof(AlfLumaEntry) * VTMGPU_ALF_FIXED_SETS * 100, sizeof(AlfLumaEntry) * 100 * sets.size());
  c->markSide(slot, c->lay.ctuCtl, (size_t)c->nCtus * sizeof(CtuCtlDev));
  c->markSide(slot, c->lay.alf, sizeof(AlfDev) * ng);
  return c->pushSlot(slot);
}
}   // namespace

extern "C" int vtmgpu_set_alf(vtmgpu_ctx* c, int slot, const vtmgpu_alf_params* p)
{
  if (!c) return -1;
  return setAlf(c, slot, 1, p, nullptr, "set_alf");
}

extern "C" int vtmgpu_set_alf_slices(vtmgpu_ctx* c, int slot, int num_slices, const vtmgpu_alf_params* slices, const uint8_t* ctu_slice)
{
  if (!c) return -1;
  if (num_slices < 1 || num_slices > 255 || !slices) return c->fail("set_alf_slices: bad argument");
  return setAlf(c, slot, num_slices, slices, ctu_slice, "set_alf_slices");
}

// ------------------------------------------------------------------------------------------------------------
// stages
// ------------------------------------------------------------------------------------------------------------
namespace
{

enum Stage { ST_DBF = 1, ST_SAO = 2, ST_ALF = 4 };

// runs one kernel over maximal runs of slots that share the same current buffer
template <class F> int forRuns(vtmgpu_ctx* c, int first, int count, F f)
{
  int s = first;
  while (s < first + count)
  {
    int e = s + 1;
    while (e < first + count && c->cur[e] == c->cur[s]) e++;
    if (f(s, e - s, c->cur[s])) return -1;
    s = e;
  }
  return 0;
}

// deblocking and / or SAO: one pass over the picture (k_dbf_sao); stages that are off for a slot are skipped inside the kernel
int launchDbfSao(vtmgpu_ctx* c, int first, int count, int doDbf, int doSao)
{
  bool any = false;
  for (int s = first; s < first + count; s++) any |= (doDbf && c->slotsPinned[s].dbfOn) || (doSao && c->slotsPinned[s].saoOn);
  if (!any && !c->bandCall) return 0;
  const Geom& g = c->g;
  DbfLaunch L;
  // rows [rowBegin, rowEnd) of the picture; vtmgpu_set_rows keeps both on multiples of 128 luma rows (or the picture end), so whole tile rows
  L.tilesXL = (g.w + DBF_TW - 1) / DBF_TW;
  L.ty0L = c->rowBegin / DBF_TH;
  L.tilesL = L.tilesXL * ((c->rowEnd + DBF_TH - 1) / DBF_TH - L.ty0L);
  L.tilesXC = g.ncomp > 1 ? ((g.w >> g.sx) + DBF_TW - 1) / DBF_TW : 0;
  L.ty0C = (c->rowBegin >> g.sy) / DBF_TH;
  L.tilesC = g.ncomp > 1 ? L.tilesXC * (((c->rowEnd >> g.sy) + DBF_TH - 1) / DBF_TH - L.ty0C) : 0;
  L.tilesLFull = c->dbfFull.tilesLFull; L.tilesCFull = c->dbfFull.tilesCFull;
  return forRuns(c, first, count, [&](int s, int n, int src) {
    const int dst = src == 1 ? 2 : 1;
    // persistent CTAs: three per SM, each walks the plane tiles round robin with double-buffered TMA loads
    const int items = L.tilesL + 2 * L.tilesC, grid = std::min(items * n, DBF_CTAS_PER_SM * c->numSms);
    TileStep st;
    st.dSlot = grid / items;
    st.dItem = grid % items;
    k_dbf_sao<<<grid, DBF_THREADS, DBF_SMEM_BYTES, c->stream>>>(c->slotsDev, c->tmapsDbfDev, s, n, src, dst, g, L, st, doDbf, doSao, c->bandCall ? c->band : BandDev{});
    c->launches++;
    for (int i = s; i < s + n; i++) c->cur[i] = dst;
    return c->cuda(cudaGetLastError(), "k_dbf_sao launch");
  });
}

// ALF + CC-ALF (k_alf)
int launchAlf(vtmgpu_ctx* c, int first, int count)
{
  bool any = false;
  for (int s = first; s < first + count; s++) any |= c->slotsPinned[s].alfOn != 0;
  if (!any && !c->bandCall) return 0;
  const Geom& g = c->g;
  const int tilesX = (g.w + SA_T - 1) / SA_T, ty0 = c->rowBegin / SA_TH, tilesY = (c->rowEnd + SA_TH - 1) / SA_TH - ty0;
  const SaLayout SL = saLayout(g.sx, g.sy, g.ncomp);
  return forRuns(c, first, count, [&](int s, int n, int src) {
    const int dst = src == 1 ? 2 : 1;
    bool vb = g.ctu < SA_T;                                      // tiles filtered in parts (virtual boundaries, several CTUs per tile) need a scratch copy of the tile
    for (int i = s; i < s + n; i++) vb |= (c->slotsPinned[i].vbAlf.nv | c->slotsPinned[i].vbAlf.nh) != 0;
    const int smem = vb ? SL.total + SL.lumaBytes + 2 * SL.chromaBytes : saLayout(g.sx, g.sy, g.ncomp, true).total;
    // persistent CTAs: two per SM (123 registers, 108 KB of shared memory each), each walks the tiles round robin with double-buffered TMA loads
    const int grid = std::min(tilesX * tilesY * n, SA_CTAS_PER_SM * c->numSms);
    SaStep st;
    st.dx = grid % tilesX;
    st.dy = (grid / tilesX) % tilesY;
    st.ds = (grid / tilesX) / tilesY;
    if (vb && c->bandCall) return c->fail("band_filter: pictures with virtual boundaries / CTU size 32 are not supported in peer band mode");
    if (vb) k_alf_parts<<<grid, SA_THREADS, smem, c->stream>>>(c->slotsDev, c->tmapsDev, s, n, src, dst, g, tilesX, tilesY, ty0, st);
    else if (g.ncomp == 3 && g.sx == 1 && g.sy == 1) k_alf<true><<<grid, SA_THREADS, smem, c->stream>>>(c->alfAddr, c->tmapsDev, s, n, src, dst, g, tilesX, tilesY, ty0, st, c->bandCall ? c->band : BandDev{});
    else                                             k_alf<false><<<grid, SA_THREADS, smem, c->stream>>>(c->alfAddr, c->tmapsDev, s, n, src, dst, g, tilesX, tilesY, ty0, st, c->bandCall ? c->band : BandDev{});
    c->launches++;
    for (int i = s; i < s + n; i++) c->cur[i] = dst;
    return c->cuda(cudaGetLastError(), "k_alf launch");
  });
}

int runStages(vtmgpu_ctx* c, int first, int count, int stages, bool sync, const char* what)
{
  if (!c) return -1;
  if (!c->slotOk(first, count)) return c->fail("%s: bad slot range [%d,%d)", what, first, first + count);
  cudaSetDevice(c->seq.device);
  c->stageValid = false;
  if (!(stages & (ST_DBF | ST_SAO)))
    for (int sl = first; sl < first + count; sl++)
      if (c->slotsPinned[sl].lmcsOn && c->cur[sl] == 0)
        return c->fail("%s: slot %d holds a reshaped-domain picture (vtmgpu_set_lmcs): the inverse mapping is part of the deblocking / SAO pass", what, sl);
  if (c->flush(first, count)) return -1;
  if (runPendingSparse(c, first, count)) return -1;
  if (c->profiling) cudaEventRecord(c->stageEv[0], c->stream);
  if ((stages & (ST_DBF | ST_SAO)) && launchDbfSao(c, first, count, (stages & ST_DBF) != 0, (stages & ST_SAO) != 0)) return -1;
  if (c->profiling) cudaEventRecord(c->stageEv[1], c->stream);
  if ((stages & ST_ALF) && launchAlf(c, first, count)) return -1;
  if (c->profiling) { cudaEventRecord(c->stageEv[2], c->stream); c->stageValid = true; }
  if (sync && !c->asyncStages && c->cuda(cudaStreamSynchronize(c->stream), what)) return -1;
  return 0;
}

}   // namespace

extern "C" int vtmgpu_deblock(vtmgpu_ctx* c, int first, int count) { return runStages(c, first, count, ST_DBF, true, "deblock"); }
extern "C" int vtmgpu_sao(vtmgpu_ctx* c, int first, int count) { return runStages(c, first, count, ST_SAO, true, "sao"); }
extern "C" int vtmgpu_alf(vtmgpu_ctx* c, int first, int count) { return runStages(c, first, count, ST_ALF, true, "alf"); }
extern "C" int vtmgpu_deblock_sao(vtmgpu_ctx* c, int first, int count) { return runStages(c, first, count, ST_DBF | ST_SAO, true, "deblock_sao"); }
extern "C" int vtmgpu_sao_alf(vtmgpu_ctx* c, int first, int count) { return runStages(c, first, count, ST_SAO | ST_ALF, true, "sao_alf"); }
extern "C" int vtmgpu_filter(vtmgpu_ctx* c, int first, int count) { return runStages(c, first, count, ST_DBF | ST_SAO | ST_ALF, true, "filter"); }
extern "C" int vtmgpu_filter_async(vtmgpu_ctx* c, int first, int count) { return runStages(c, first, count, ST_DBF | ST_SAO | ST_ALF, false, "filter_async"); }

// ------------------------------------------------------------------------------------------------------------
// decoded-picture hash on the device
// ------------------------------------------------------------------------------------------------------------
extern "C" int vtmgpu_hash(vtmgpu_ctx* c, int first, int count, int kind, uint8_t* digest, int* bytes_per_component)
{
  if (!c) return -1;
  if (!c->slotOk(first, count)) return c->fail("hash: bad slot range [%d,%d)", first, first + count);
  if (!digest) return c->fail("hash: NULL digest");
  if (kind != VTMGPU_HASH_MD5 && kind != VTMGPU_HASH_CRC && kind != VTMGPU_HASH_CHECKSUM) return c->fail("hash: unknown kind %d", kind);
  cudaSetDevice(c->seq.device);
  const int nc = c->g.ncomp, n = count * nc, words = kind == VTMGPU_HASH_MD5 ? 4 : 1;
  std::vector<HashPlane> hp(n);
  unsigned char* dev = nullptr;
  const size_t planeBytes = sizeof(HashPlane) * n, outBytes = sizeof(uint32_t) * words * n;
  if (c->cuda(cudaMalloc((void**)&dev, planeBytes + outBytes), "hash scratch")) return -1;
  uint32_t* out = reinterpret_cast<uint32_t*>(dev + planeBytes);
  for (int s = 0; s < count; s++)
    for (int k = 0; k < nc; k++)
    {
      const PlaneDev& d = c->slotsPinned[first + s].buf[c->cur[first + s]][k];
      hp[s * nc + k] = HashPlane{ d.p, d.pitch, d.w, d.h, k ? c->g.bdC : c->g.bdL, out + (size_t)(s * nc + k) * words };
    }
  int rc = 0;
  std::vector<uint32_t> res((size_t)words * n);
  if (c->cuda(cudaMemcpyAsync(dev, hp.data(), planeBytes, cudaMemcpyHostToDevice, c->stream), "hash upload") ||
      c->cuda(cudaMemsetAsync(out, 0, outBytes, c->stream), "hash clear")) rc = -1;
  if (!rc)
  {
    if (kind == VTMGPU_HASH_MD5)           k_hash_md5<<<n, 32, 0, c->stream>>>(reinterpret_cast<const HashPlane*>(dev), n);
    else if (kind == VTMGPU_HASH_CRC)      k_hash_crc<<<dim3(64, n), 128, 0, c->stream>>>(reinterpret_cast<const HashPlane*>(dev), n);
    else                                   k_hash_checksum<<<dim3(64, n), 256, 0, c->stream>>>(reinterpret_cast<const HashPlane*>(dev), n);
    c->launches++;
    if (c->cuda(cudaGetLastError(), "hash launch") || c->cuda(cudaMemcpyAsync(res.data(), out, outBytes, cudaMemcpyDeviceToHost, c->stream), "hash download") ||
        c->cuda(cudaStreamSynchronize(c->stream), "hash")) rc = -1;
  }
  cudaFree(dev);
  if (rc) return rc;
  const int len = kind == VTMGPU_HASH_MD5 ? 16 : (kind == VTMGPU_HASH_CRC ? 2 : 4);
  memset(digest, 0, (size_t)count * VTMGPU_HASH_SLOT_BYTES);
  for (int s = 0; s < count; s++)
    for (int k = 0; k < nc; k++)
    {
      uint8_t* d = digest + (size_t)s * VTMGPU_HASH_SLOT_BYTES + k * len;
      const uint32_t* r = &res[(size_t)(s * nc + k) * words];
      if (kind == VTMGPU_HASH_MD5) for (int i = 0; i < 16; i++) d[i] = (uint8_t)(r[i >> 2] >> (8 * (i & 3)));         // state words, little endian
      else if (kind == VTMGPU_HASH_CRC) { d[0] = (uint8_t)(r[0] >> 8); d[1] = (uint8_t)r[0]; }                          // :127-128
      else for (int i = 0; i < 4; i++) d[i] = (uint8_t)(r[0] >> (24 - 8 * i));                                          // :180-183
    }
  if (bytes_per_component) *bytes_per_component = len;
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// band mode over peer memory
// ------------------------------------------------------------------------------------------------------------
namespace
{
struct BandHandle      // the bytes of vtmgpu_band_handle
{
  cudaIpcMemHandle_t planes, flags;
  int32_t width, height, chroma_format, capacity, device;
  uint64_t planeBytes;
};
static_assert(sizeof(BandHandle) <= VTMGPU_BAND_HANDLE_BYTES, "vtmgpu_band_handle too small");
}   // namespace

extern "C" int vtmgpu_band_export(vtmgpu_ctx* c, vtmgpu_band_handle* out)
{
  if (!c) return -1;
  if (!out) return c->fail("band_export: NULL handle");
  cudaSetDevice(c->seq.device);
  if (!c->bandFlags)
  {
    if (c->cuda(cudaMalloc((void**)&c->bandFlags, 64), "band flags")) return -1;
    if (c->cuda(cudaMemset(c->bandFlags, 0, 64), "band flags")) return -1;
  }
  BandHandle h{};
  if (c->cuda(cudaIpcGetMemHandle(&h.planes, c->planeAll), "cudaIpcGetMemHandle (planes)")) return -1;
  if (c->cuda(cudaIpcGetMemHandle(&h.flags, c->bandFlags), "cudaIpcGetMemHandle (flags)")) return -1;
  h.width = c->seq.width; h.height = c->seq.height; h.chroma_format = c->seq.chroma_format; h.capacity = c->seq.capacity; h.device = c->seq.device;
  h.planeBytes = c->planeAllBytes;
  memset(out, 0, sizeof(*out));
  memcpy(out->bytes, &h, sizeof(h));
  return 0;
}

extern "C" int vtmgpu_band_disconnect(vtmgpu_ctx* c)
{
  if (!c) return -1;
  if (!c->bandOn && !c->bandPeerMem[0][0] && !c->bandPeerMem[1][0]) return 0;
  cudaSetDevice(c->seq.device);
  cudaStreamSynchronize(c->stream);
  for (auto& side : c->bandPeerMem)
    for (void*& p : side)
      if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
  c->bandOn = false;
  c->band = BandDev{};
  return 0;
}

extern "C" int vtmgpu_band_connect(vtmgpu_ctx* c, const vtmgpu_band_handle* above, const vtmgpu_band_handle* below)
{
  if (!c) return -1;
  if (!c->bandFlags) return c->fail("band_connect: call vtmgpu_band_export first (it creates this rank's flag block)");
  if (vtmgpu_band_disconnect(c)) return -1;
  cudaSetDevice(c->seq.device);
  if (c->cuda(cudaMemset(c->bandFlags, 0, 64), "band flags")) return -1;
  BandDev b{};
  b.myPlanes = c->planeAll; b.myFlags = c->bandFlags; b.iter = 0; b.rowBegin = c->rowBegin; b.rowEnd = c->rowEnd;
  const vtmgpu_band_handle* hs[2] = { above, below };
  for (int side = 0; side < 2; side++)
  {
    if (!hs[side]) continue;
    BandHandle h;
    memcpy(&h, hs[side]->bytes, sizeof(h));
    if (h.width != c->seq.width || h.height != c->seq.height || h.chroma_format != c->seq.chroma_format || h.capacity != c->seq.capacity || h.planeBytes != c->planeAllBytes)
      return c->fail("band_connect: the neighbour's context has another geometry or capacity");
    if (c->cuda(cudaIpcOpenMemHandle(&c->bandPeerMem[side][0], h.planes, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (planes)")) return -1;
    if (c->cuda(cudaIpcOpenMemHandle(&c->bandPeerMem[side][1], h.flags, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (flags)")) return -1;
    b.peerPlanes[side] = static_cast<pel*>(c->bandPeerMem[side][0]);
    b.peerFlags[side] = static_cast<uint32_t*>(c->bandPeerMem[side][1]);
  }
  c->band = b;
  c->bandOn = true;
  return c->cuda(cudaDeviceSynchronize(), "band_connect");
}

extern "C" int vtmgpu_band_filter_async(vtmgpu_ctx* c, int slot)
{
  if (!c) return -1;
  if (!c->bandOn) return c->fail("band_filter: not connected (vtmgpu_band_connect)");
  c->band.iter++;
  c->band.rowBegin = c->rowBegin; c->band.rowEnd = c->rowEnd;
  c->bandCall = true;            // both kernels are launched even when their stages are off for the picture: the neighbours wait for the flags
  const int rc = runStages(c, slot, 1, ST_DBF | ST_SAO | ST_ALF, false, "band_filter");
  c->bandCall = false;
  return rc;
}

// ------------------------------------------------------------------------------------------------------------
// host batches
// ------------------------------------------------------------------------------------------------------------
// A batch moves pictures through THREE streams -- uploads, kernels, downloads -- with events between them, over `lanes`
// single-picture contexts that only serve as buffers:
//   up    planes, record lists, SAO / ALF side information of picture i          (waits for the download of picture i - lanes)
//   run   k_clear16, k_dbf_scatter, k_dbf_queues, k_dbf_sao, k_alf of picture i  (waits for its upload)
//   down  the filtered planes                                                    (waits for its kernels)
// Each copy engine sees ONE queue of back-to-back copies that never wait for a kernel, whatever the number of hardware queues the
// device maps streams to, and the download of a picture is held back until the upload after it has finished so that the two
// directions start their copies together.  Measured on B200, 64 4K pictures per call (tools/microbench/e2e_variants.py, event time
// stamps per picture): the uplink is busy without a gap and bounds the call.  Downloads released as soon as their kernels are done
// begin in the middle of an upload, which then runs at 40 GB/s beside the 48 GB/s download: 42.3 ms (the earlier form, one stream
// per lane with upload, kernels and download in stream order, measured the same 42.6 for 2, 4, 8 or 16 lanes).  Aligned starts:
// 37.3 ms, against 35.4 for the same bytes as two plain copy streams (47 + 45 GB/s) and 36.7 for this pipeline without kernels.
struct vtmgpu_batch
{
  std::vector<vtmgpu_ctx*> lane;
  std::vector<cudaEvent_t> evUp, evRun, evDown;      // per lane: upload / kernels / download of its current picture finished
  std::vector<char> used;                            // per lane: evDown has been recorded
  cudaStream_t up = nullptr, run = nullptr, down = nullptr;
  int device = 0;
  std::string err;
};

extern "C" const char* vtmgpu_batch_last_error(const vtmgpu_batch* b) { return b ? b->err.c_str() : g_createError.c_str(); }

extern "C" void vtmgpu_batch_destroy(vtmgpu_batch* b)
{
  if (!b) return;
  cudaSetDevice(b->device);
  for (cudaStream_t st : { b->up, b->run, b->down }) if (st) { cudaStreamSynchronize(st); }
  for (vtmgpu_ctx* c : b->lane) { c->stream = c->ownStream; vtmgpu_destroy(c); }
  for (auto* v : { &b->evUp, &b->evRun, &b->evDown }) for (cudaEvent_t e : *v) if (e) cudaEventDestroy(e);
  for (cudaStream_t st : { b->up, b->run, b->down }) if (st) cudaStreamDestroy(st);
  delete b;
}

extern "C" int vtmgpu_batch_create(const vtmgpu_seq_params* seq, int lanes, vtmgpu_batch** out)
{
  if (!seq || !out || lanes < 1 || lanes > 64) { g_createError = "vtmgpu_batch_create: bad argument (1..64 lanes)"; return -1; }
  *out = nullptr;
  vtmgpu_batch* b = new vtmgpu_batch();
  vtmgpu_seq_params s = *seq;
  s.capacity = 1;
  b->device = s.device;
  for (int k = 0; k < lanes; k++)
  {
    vtmgpu_ctx* c = nullptr;
    if (vtmgpu_create(&s, &c)) { vtmgpu_batch_destroy(b); return -1; }      // g_createError holds the reason
    b->lane.push_back(c);
  }
  cudaSetDevice(s.device);
  bool ok = cudaStreamCreateWithFlags(&b->up, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&b->run, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&b->down, cudaStreamNonBlocking) == cudaSuccess;
  b->evUp.assign(lanes, nullptr); b->evRun.assign(lanes, nullptr); b->evDown.assign(lanes, nullptr); b->used.assign(lanes, 0);
  for (int k = 0; k < lanes && ok; k++)
    ok = cudaEventCreateWithFlags(&b->evUp[k], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&b->evRun[k], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&b->evDown[k], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { g_createError = std::string("vtmgpu_batch_create: ") + cudaGetErrorString(cudaGetLastError()); vtmgpu_batch_destroy(b); return -1; }
  *out = b;
  return 0;
}

extern "C" int64_t vtmgpu_batch_launch_count(const vtmgpu_batch* b)
{
  int64_t n = 0;
  if (b) for (const vtmgpu_ctx* c : b->lane) n += c->launches;
  return b ? n : -1;
}

extern "C" int vtmgpu_batch_filter(vtmgpu_batch* b, const vtmgpu_host_picture* pics, int count)
{
  if (!b) return -1;
  if (count < 0 || (count && !pics)) { b->err = "batch_filter: bad argument"; return -1; }
  auto lastError = [&](vtmgpu_ctx* c, int i) { b->err = "batch_filter: picture " + std::to_string(i) + ": " + c->err; return -1; };
  cudaSetDevice(b->device);
  const int lanes = (int)b->lane.size();
#ifdef VTMGPU_BATCH_KNOBS
  const int skip = getenv("VTMGPU_BATCH_SKIP") ? atoi(getenv("VTMGPU_BATCH_SKIP")) : 0;     // experiment build (tools/microbench/e2e_variants.py)
#else
  const int skip = 0;
#endif
  // The download of picture i is enqueued behind the upload of picture i + 1 and waits for it (and for its own kernels): a
  // download then starts together with the upload after next instead of somewhere inside one.  Measured (e2e_variants.py): a
  // download that begins in the middle of an upload costs the uplink 15 % of its rate for the rest of that copy (43.0 -> 37.3-39 ms
  // per 64 pictures).  With one or two lanes the wait would serialise the lanes: the download follows its kernels directly.
  const bool defer = lanes >= 3;
  auto download = [&](int j, int next) -> int
  {
    const int kj = j % lanes;
    vtmgpu_ctx* c = b->lane[kj];
    c->stream = b->down;
    if ((next >= 0 && c->cuda(cudaStreamWaitEvent(b->down, b->evUp[next % lanes], 0), "batch_filter")) ||
        c->cuda(cudaStreamWaitEvent(b->down, b->evRun[kj], 0), "batch_filter") || (!(skip & 16) && vtmgpu_download_async(c, 0, pics[j].out, pics[j].out_stride)) ||
        c->cuda(cudaEventRecord(b->evDown[kj], b->down), "batch_filter"))
      return lastError(c, j);
    b->used[kj] = 1;
    return 0;
  };
  int rc = 0, issued = 0;                // issued = pictures whose upload and kernels are enqueued
  for (int i = 0; i < count && !rc; i++)
  {
    const int k = i % lanes;
    vtmgpu_ctx* c = b->lane[k];
    const vtmgpu_host_picture& p = pics[i];
    // everything below only enqueues; the vtmgpu_set_* calls wait (mirror events) until the side-information upload of the lane's
    // previous picture has left the pinned mirror
    c->stream = b->up;
    if (b->used[k] && c->cuda(cudaStreamWaitEvent(b->up, b->evDown[k], 0), "batch_filter")) { rc = lastError(c, i); break; }   // the lane's buffers are free again
    if ((!(skip & 32) && vtmgpu_upload_async(c, 0, p.in, p.in_stride)) || (!(skip & 1) && vtmgpu_set_deblock_sparse(c, 0, p.deblock)) || vtmgpu_set_sao(c, 0, p.sao) ||
        vtmgpu_set_alf(c, 0, p.alf) || c->flush(0, 1) || c->cuda(cudaEventRecord(b->evUp[k], b->up), "batch_filter"))
    { rc = lastError(c, i); break; }
    c->stream = b->run;
    if (c->cuda(cudaStreamWaitEvent(b->run, b->evUp[k], 0), "batch_filter") || (!(skip & 8) && vtmgpu_filter_async(c, 0, 1)) || c->cuda(cudaEventRecord(b->evRun[k], b->run), "batch_filter"))
    { rc = lastError(c, i); break; }
    issued = i + 1;
    if (!defer) rc = download(i, -1);
    else if (i > 0) rc = download(i - 1, i);
  }
  if (defer && issued > 0 && !rc) rc = download(issued - 1, -1);
  // every output has landed when the three streams are drained (also on the error path: nothing may still read the caller's buffers)
  for (cudaStream_t st : { b->up, b->run, b->down })
    if (st && cudaStreamSynchronize(st) != cudaSuccess && !rc) { b->err = std::string("batch_filter: ") + cudaGetErrorString(cudaGetLastError()); rc = -1; }
  for (vtmgpu_ctx* c : b->lane) c->stream = c->ownStream;
  return rc;
}

extern "C" int vtmgpu_timer_start(vtmgpu_ctx* c)
{
  if (!c) return -1;
  cudaSetDevice(c->seq.device);
  return c->cuda(cudaEventRecord(c->ev[0], c->stream), "timer_start");
}

extern "C" int vtmgpu_timer_stop(vtmgpu_ctx* c, float* ms)
{
  if (!c || !ms) return -1;
  cudaSetDevice(c->seq.device);
  if (c->cuda(cudaEventRecord(c->ev[1], c->stream), "timer_stop")) return -1;
  if (c->cuda(cudaEventSynchronize(c->ev[1]), "timer_stop")) return -1;
  return c->cuda(cudaEventElapsedTime(ms, c->ev[0], c->ev[1]), "timer_stop");
}

extern "C" int64_t vtmgpu_launch_count(const vtmgpu_ctx* c) { return c ? c->launches : -1; }

extern "C" int vtmgpu_set_profiling(vtmgpu_ctx* c, int on)
{
  if (!c) return -1;
  c->profiling = on != 0;
  c->stageValid = false;
  return 0;
}

extern "C" int vtmgpu_stage_ms(vtmgpu_ctx* c, float ms[4])
{
  if (!c || !ms) return -1;
  if (!c->stageValid) return c->fail("stage_ms: no profiled run recorded");
  cudaSetDevice(c->seq.device);
  if (c->cuda(cudaEventSynchronize(c->stageEv[2]), "stage_ms")) return -1;
  ms[0] = ms[1] = ms[2] = ms[3] = 0.f;
  cudaEventElapsedTime(&ms[0], c->stageEv[0], c->stageEv[1]);
  cudaEventElapsedTime(&ms[1], c->stageEv[1], c->stageEv[2]);
  return 0;
}
