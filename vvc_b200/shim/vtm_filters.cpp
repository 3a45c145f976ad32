// vtm_filters.cpp -- the drop-in: OUR definitions of the entry points DecLib::executeLoopFilters needs from
// LoopFilter / SampleAdaptiveOffset / AdaptiveLoopFilter (the 17 symbols of SURVEY.md 8b), compiled against the
// reference's unmodified headers (class layouts stay identical, DecLib embeds the objects by value) and linked
// INSTEAD of the reference's LoopFilter.cpp / SampleAdaptiveOffset.cpp / AdaptiveLoopFilter.cpp (+ the ALF SIMD stubs).
//
//   loopFilterPic : derive segment records on the host -> upload planes (+ records); the device pass is deferred while a
//                   later stage follows (nothing reads the reconstruction between the three calls of executeLoopFilters)
//   SAOProcess    : reconstruct SAO params; deferred when the SPS enables ALF, else deblocking+SAO kernel -> download
//   ALFProcess    : flatten slice/APS/CTU data -> vtmgpu_filter = (deblocking + SAO) kernel, (ALF + CC-ALF) kernel -> download
//
// No sample is filtered on the CPU here.  If libvtmgpu cannot be loaded or any call fails the shim THROWs
// (reference error convention, TypeDef.h:1152) -- there is no fallback.
//
// environment:
//   VTMGPU_LIB=<path>           libvtmgpu.so to dlopen (default: libvtmgpu.so via rpath / LD_LIBRARY_PATH)
//   VTMGPU_DEVICE=<n>           CUDA device ordinal (default 0)
//   VTMGPU_STAGED=1             run and download stage by stage (no SAO+ALF fusion)
//   VTMGPU_CAPTURE_DIR=<dir>    write one VTMGCAP1 file per picture (capture_format.h); implies staged
//   VTMGPU_SHIM_BACKEND=ref     test binaries only: use the linked alternative backend (shim_backend.h)
#include <dlfcn.h>

#include <chrono>
#include <exception>
#include <mutex>
#include <thread>

#include <cstdio>
#include <cstdlib>
#include <string>

#include "AdaptiveLoopFilter.h"
#include "CodingStructure.h"
#include "LoopFilter.h"
#include "Picture.h"
#include "SampleAdaptiveOffset.h"

#include "capture_format.h"
#ifdef VTMGPU_SHIM_TEST_BACKEND
#include "shim_backend.h"               // test binaries only (oracle/_ref/DecoderApp_cap): the reference's own classes as a second backend
#define SHIM_ALT(CALL) vtmgpu_shim_alt_backend()->CALL
#else
#define SHIM_ALT(CALL) ((void)0)        // product object: the hook is compiled out, `useRef` is a constant false
#endif
#include "vtm_flatten.h"
#include "vtmgpu.h"

namespace
{

using namespace vtmshim;

struct GpuApi
{
  void* so = nullptr;
  decltype(&vtmgpu_abi_version) abi_version = nullptr;
  decltype(&vtmgpu_last_error) last_error = nullptr;
  decltype(&vtmgpu_create) create = nullptr;
  decltype(&vtmgpu_destroy) destroy = nullptr;
  decltype(&vtmgpu_upload) upload = nullptr;
  decltype(&vtmgpu_download) download = nullptr;
  decltype(&vtmgpu_set_deblock) set_deblock = nullptr;
  decltype(&vtmgpu_set_deblock_sparse) set_deblock_sparse = nullptr;
  decltype(&vtmgpu_set_deblock_units) set_deblock_units = nullptr;
  decltype(&vtmgpu_get_deblock_records) get_deblock_records = nullptr;
  decltype(&vtmgpu_upload_motion) upload_motion = nullptr;
  decltype(&vtmgpu_sync) sync = nullptr;
  decltype(&vtmgpu_set_sao) set_sao = nullptr;
  decltype(&vtmgpu_set_alf) set_alf = nullptr;
  decltype(&vtmgpu_set_alf_slices) set_alf_slices = nullptr;
  decltype(&vtmgpu_host_register) host_register = nullptr;
  decltype(&vtmgpu_host_unregister) host_unregister = nullptr;
  decltype(&vtmgpu_set_lmcs) set_lmcs = nullptr;
  decltype(&vtmgpu_download_extended) download_extended = nullptr;
  decltype(&vtmgpu_sao_reconstruct) sao_reconstruct = nullptr;
  decltype(&vtmgpu_deblock) deblock = nullptr;
  decltype(&vtmgpu_sao) sao = nullptr;
  decltype(&vtmgpu_alf) alf = nullptr;
  decltype(&vtmgpu_sao_alf) sao_alf = nullptr;
  decltype(&vtmgpu_deblock_sao) deblock_sao = nullptr;
  decltype(&vtmgpu_filter) filter = nullptr;

  void load()
  {
    if (so) return;
    const char* path = getenv("VTMGPU_LIB");
    so = dlopen(path ? path : "libvtmgpu.so", RTLD_NOW | RTLD_LOCAL);
    if (!so) THROW("vtmgpu shim: cannot load libvtmgpu.so (" << dlerror() << ") -- no CPU fallback");
#define SYM(n) n = reinterpret_cast<decltype(n)>(dlsym(so, "vtmgpu_" #n)); if (!n) THROW("vtmgpu shim: missing symbol vtmgpu_" #n)
    SYM(abi_version); SYM(last_error); SYM(create); SYM(destroy); SYM(upload); SYM(download); SYM(set_deblock); SYM(set_deblock_sparse); SYM(set_deblock_units); SYM(get_deblock_records); SYM(upload_motion); SYM(sync); SYM(set_sao);
    SYM(set_alf); SYM(set_alf_slices); SYM(set_lmcs); SYM(download_extended); SYM(host_register); SYM(host_unregister); SYM(sao_reconstruct); SYM(deblock); SYM(sao); SYM(alf); SYM(sao_alf); SYM(deblock_sao); SYM(filter);
#undef SYM
    if (abi_version() != VTMGPU_ABI_VERSION) THROW("vtmgpu shim: ABI version mismatch");
  }
};

struct Shim
{
  GpuApi api;
  vtmgpu_ctx* ctx = nullptr;
  vtmgpu_seq_params seq{};
#ifdef VTMGPU_SHIM_TEST_BACKEND
  bool useRef = false;
#else
  static constexpr bool useRef = false;
#endif
  bool staged = false, saoPending = false, dbfPending = false, denseRecords = false;
  // the two whole-picture host passes either side of the chain move to the device --
  //   LMCS inverse mapping: the host call is recorded (see AreaBuf<Pel>::rspSignal below) and applied by the deblocking kernel's tile load
  //   border extension:     the last download of a picture brings the margins along (vtmgpu_download_extended)
  // Both are off for captures and the reference backend (they see / produce host-side samples), VTMGPU_SHIM_HOST_LMCS=1 / VTMGPU_SHIM_EXTEND=0 switch them off.
  bool deferLmcs = false, extendOnDevice = false, pinOn = true;
  bool lmcsPending = false, lmcsActive = false;
  const Pel* lmcsBuf = nullptr;
  std::vector<Pel> lmcsLut;
  int lmcsPics = 0, extendedPics = 0, lmcsHostPics = 0;
  bool capturePreOnly = false;   // VTMGPU_CAPTURE_PRE_ONLY=1: captures hold the pre-filter planes and the side information only (a quarter of the size)
  std::string captureDir;
  int picCount = 0;
  // per-picture
  FlatDeblock dbf, dbfFromUnits;
  FlatUnits units;
  // the block structure goes down (flattenUnits) and the DEVICE derives the deblocking records (VTMGPU_SHIM_DEVICE_DERIVE=0: the CU walk);
  // VTMGPU_SHIM_CHECK_UNITS=1 -- self check: the CU walk and the unit derivation (the kernel's own source run on the host, and with a device
  // the kernel's output) must produce identical records for every picture
  bool deviceDerive = false, checkUnits = false;
  int unitPics = 0, unitFallbackPics = 0, checkedPics = 0, deviceCheckedPics = 0;
  double flattenSec = 0;
  FlatSao sao;
  FlatAlf alf;
  CaptureWriter cap;
  int saoLog2Scale[2] = { 0, 0 };

  Shim()
  {
    const char* b = getenv("VTMGPU_SHIM_BACKEND");
#ifdef VTMGPU_SHIM_TEST_BACKEND
    useRef = b && std::string(b) == "ref";
    if (useRef && (!vtmgpu_shim_alt_backend || !vtmgpu_shim_alt_backend())) THROW("vtmgpu shim: no alternative backend linked into this binary");
#else
    if (b && std::string(b) == "ref") THROW("vtmgpu shim: this binary has no alternative backend (VTMGPU_SHIM_BACKEND=ref is a test-binary switch)");
#endif
    if (const char* d = getenv("VTMGPU_CAPTURE_DIR")) captureDir = d;
    capturePreOnly = getenv("VTMGPU_CAPTURE_PRE_ONLY") && atoi(getenv("VTMGPU_CAPTURE_PRE_ONLY"));
    staged = !captureDir.empty() || (getenv("VTMGPU_STAGED") && atoi(getenv("VTMGPU_STAGED")));
    denseRecords = getenv("VTMGPU_DENSE_RECORDS") && atoi(getenv("VTMGPU_DENSE_RECORDS"));
    timing = getenv("VTMGPU_SHIM_TIMING") && atoi(getenv("VTMGPU_SHIM_TIMING"));
#ifdef VTMGPU_SHIM_ENCODER
    // encoder-side reuse: only the deblocking of the final reconstruction runs on the device, the encoder's SAO / ALF
    // parameter searches read the deblocked picture on the host right afterwards (EncGOP.cpp:2794-2830) -> download after the stage
    staged = true;
    const bool product = false;
#else
    const bool product = !useRef && captureDir.empty();
#endif
    deferLmcs = product && !(getenv("VTMGPU_SHIM_HOST_LMCS") && atoi(getenv("VTMGPU_SHIM_HOST_LMCS")));
    extendOnDevice = product && !(getenv("VTMGPU_SHIM_EXTEND") && !atoi(getenv("VTMGPU_SHIM_EXTEND")));
    pinOn = !(getenv("VTMGPU_SHIM_PIN") && !atoi(getenv("VTMGPU_SHIM_PIN")));
    deviceDerive = product && !(getenv("VTMGPU_SHIM_DEVICE_DERIVE") && !atoi(getenv("VTMGPU_SHIM_DEVICE_DERIVE")));      // default in the product decoder; =0: the CU walk on host threads
    checkUnits = getenv("VTMGPU_SHIM_CHECK_UNITS") && atoi(getenv("VTMGPU_SHIM_CHECK_UNITS"));
  }
  ~Shim()
  {
    if (timing)
      printf("vtmgpu-shim-timing: pictures=%d luma_pixels=%lld filter_s=%.6f dbf_s=%.6f sao_s=%.6f alf_s=%.6f derive_s=%.6f backend=%s "
             "record_lists=%d record_bytes=%lld lmcs_on_device=%d border_on_device=%d lmcs_on_host=%d pinned_planes=%d device_derived=%d walk_derived_instead=%d units_checked=%d kernel_records_checked=%d flatten_s=%.6f\n", picCount, lumaPixels, stageSec[0] + stageSec[1] + stageSec[2], stageSec[0], stageSec[1], stageSec[2], deriveSec,
             useRef ? "ref" : "gpu", listPics, recordBytes, lmcsPics, extendedPics, lmcsHostPics, pinnedPlanes, unitPics, unitFallbackPics, checkedPics, deviceCheckedPics, flattenSec);
    if (ctx)
    {
      for (const Range& r : ranges) if (r.live) api.host_unregister(reinterpret_cast<void*>(r.lo));
      api.destroy(ctx);
    }
  }

  // VTMGPU_SHIM_TIMING=1: steady_clock around the backend's stage calls only (for the reference backend that is exactly
  // loopFilterPic / SAOProcess / ALFProcess of the reference classes -- BASELINE.md section 4); host derivation separately
  bool timing = false;
  int listPics = 0;              // pictures whose deblocking records went up as lists
  long long recordBytes = 0;     // bytes of deblocking records handed to the library
  double stageSec[3] = { 0, 0, 0 }, deriveSec = 0;
  long long lumaPixels = 0;
  std::chrono::steady_clock::time_point t0;
  void tic() { if (timing) t0 = std::chrono::steady_clock::now(); }
  void toc(double& acc) { if (timing) acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }

  void check(int rc, const char* what) { if (rc) THROW("vtmgpu shim: " << what << " failed: " << (api.last_error ? api.last_error(ctx) : "?")); }

  void ensureCtx(const CodingStructure& cs)
  {
    vtmgpu_seq_params s{};
    s.width = cs.pcv->lumaWidth;
    s.height = cs.pcv->lumaHeight;
    s.chroma_format = (int)cs.pcv->chrFormat;
    s.bit_depth_luma = cs.sps->getBitDepth(CHANNEL_TYPE_LUMA);
    s.bit_depth_chroma = cs.sps->getBitDepth(CHANNEL_TYPE_CHROMA);
    s.ctu_size = cs.pcv->maxCUWidth;
    s.capacity = 1;
    s.device = getenv("VTMGPU_DEVICE") ? atoi(getenv("VTMGPU_DEVICE")) : 0;
    const bool same = s.width == seq.width && s.height == seq.height && s.chroma_format == seq.chroma_format && s.bit_depth_luma == seq.bit_depth_luma &&
                      s.bit_depth_chroma == seq.bit_depth_chroma && s.ctu_size == seq.ctu_size;
    seq = s;
    if (useRef || (ctx && same)) return;
    api.load();
    if (ctx)
    {
      // new sequence geometry: the decoder re-creates its picture buffers -- drop every page lock taken for the old ones
      {
        std::lock_guard<std::mutex> lock(pinMutex);
        for (const Range& r : ranges) if (r.live) api.host_unregister(reinterpret_cast<void*>(r.lo));
        ranges.clear();
        seenPlanes.clear();
        pinnedPlanes = 0;
      }
      api.destroy(ctx); ctx = nullptr;
    }
    if (api.create(&seq, &ctx)) THROW("vtmgpu shim: vtmgpu_create failed: " << api.last_error(nullptr));
    lmcsActive = false;
  }

  // The decoder's picture buffers are pageable (PelStorage::create, Buffer.cpp:726) and live as long as the sequence (the DPB recycles
  // them): each one is page-locked the first time it is transferred (vtmgpu_host_register), so uploads and downloads are plain DMA
  // instead of two staged copies at a fraction of the PCIe rate.  A buffer that cannot be registered simply stays pageable.
  // VTMGPU_SHIM_PIN=0 switches this off.
  struct Range { uintptr_t lo, hi; bool live; };       // page-aligned; live = registered, else given up (stays pageable)
  std::vector<Range> ranges;
  std::vector<const void*> seenPlanes;
  int pinnedPlanes = 0;
  std::mutex pinMutex;                 // the upload thread and the decoder thread both transfer (and therefore pin)
  void pin(const CodingStructure& cs, const Pel* buf, ptrdiff_t stride, int height, int comp)
  {
    std::lock_guard<std::mutex> lock(pinMutex);
    for (const void* e : seenPlanes) if (e == buf) return;
    seenPlanes.push_back(buf);
    if (!pinOn) return;
    const int xm = (int)cs.picture->margin >> getComponentScaleX(ComponentID(comp), cs.pcv->chrFormat), ym = (int)cs.picture->margin >> getComponentScaleY(ComponentID(comp), cs.pcv->chrFormat);
    // the allocation behind the plane: margins on every side (Picture::create, Picture.cpp:199-206), rows `stride` apart
    pinRange(reinterpret_cast<uintptr_t>(buf - (ptrdiff_t)ym * stride - xm), reinterpret_cast<uintptr_t>(buf + (ptrdiff_t)(height + ym - 1) * stride + (stride - xm)));
  }
  void pinOnce(const void* p, size_t bytes)           // other long-lived host buffers the shim transfers from (motion fields, the block-structure tables)
  {
    std::lock_guard<std::mutex> lock(pinMutex);
    for (const void* e : seenPlanes) if (e == p) return;
    seenPlanes.push_back(p);
    if (pinOn) pinRange(reinterpret_cast<uintptr_t>(p), reinterpret_cast<uintptr_t>(p) + bytes);
  }
  void unpin(const void* p)
  {
    std::lock_guard<std::mutex> lock(pinMutex);
    for (size_t i = 0; i < seenPlanes.size(); i++) if (seenPlanes[i] == p) { seenPlanes.erase(seenPlanes.begin() + i); break; }
    const uintptr_t a = reinterpret_cast<uintptr_t>(p) & ~uintptr_t(4095);
    for (size_t i = 0; i < ranges.size(); i++)
      if (ranges[i].lo == a) { if (ranges[i].live) api.host_unregister(reinterpret_cast<void*>(a)); ranges.erase(ranges.begin() + i); break; }
  }
  void pinRange(uintptr_t lo, uintptr_t hi)
  {
    lo &= ~uintptr_t(4095); hi = (hi + 4095) & ~uintptr_t(4095);
    // Small pictures come from the heap and neighbouring allocations share pages: a page cannot be registered twice, and a transfer whose
    // host range is only partly page-locked is refused by the driver -- so ranges that touch are merged into one registration, and a
    // range that cannot be registered is remembered and left alone (its planes stay pageable as a whole).
    for (;;)
    {
      bool merged = false;
      for (size_t i = 0; i < ranges.size(); i++)
      {
        Range& r = ranges[i];
        if (r.lo <= lo && hi <= r.hi) { if (r.live) pinnedPlanes++; return; }          // already covered (or given up)
        if (r.lo < hi && lo < r.hi)
        {
          if (r.live) api.host_unregister(reinterpret_cast<void*>(r.lo));
          lo = std::min(lo, r.lo); hi = std::max(hi, r.hi);
          const bool wasDead = !r.live;
          ranges.erase(ranges.begin() + i);
          if (wasDead) { ranges.push_back(Range{ lo, hi, false }); return; }
          merged = true;
          break;
        }
      }
      if (!merged) break;
    }
    const bool ok = api.host_register(reinterpret_cast<void*>(lo), hi - lo) == 0;
    ranges.push_back(Range{ lo, hi, ok });
    if (ok) pinnedPlanes++;
  }
  void planes(CodingStructure& cs, int16_t* p[3], ptrdiff_t st[3], int w[3], int h[3])
  {
    PelUnitBuf rec = cs.getRecoBuf();
    for (int c = 0; c < 3; c++)
    {
      p[c] = nullptr; st[c] = 0; w[c] = h[c] = 0;
      if (c >= (int)getNumberValidComponents(cs.pcv->chrFormat)) continue;
      PelBuf& b = rec.get(ComponentID(c));
      p[c] = b.buf; st[c] = b.stride; w[c] = b.width; h[c] = b.height;
      if (!useRef && ctx) pin(cs, b.buf, b.stride, b.height, c);
    }
  }
  void upload(CodingStructure& cs)   { int16_t* p[3]; ptrdiff_t st[3]; int w[3], h[3]; planes(cs, p, st, w, h); check(api.upload(ctx, 0, p, st), "upload"); }
  // last = the picture is final: its reference-picture margins come along (Picture::extendPicBorder, Picture.cpp:737-772, would otherwise
  // run on the host when the picture is first used as a reference, Slice.cpp:456); the wrap-around variant stays on the host
  void download(CodingStructure& cs, bool last)
  {
    int16_t* p[3]; ptrdiff_t st[3]; int w[3], h[3];
    planes(cs, p, st, w, h);
    if (last && extendOnDevice && !cs.sps->getWrapAroundEnabledFlag() && (cs.picture->margin & 15) == 0 && cs.picture->margin >= 16 && cs.picture->margin <= 1024)
    {
      check(api.download_extended(ctx, 0, p, st, (int)cs.picture->margin), "download_extended");
      cs.picture->setBorderExtension(true);
      extendedPics++;
    }
    else check(api.download(ctx, 0, p, st), "download");
  }

  void capturePlanes(CodingStructure& cs, const char* stage)
  {
    if (captureDir.empty() || (capturePreOnly && std::string(stage) != "pre")) return;
    int16_t* p[3]; ptrdiff_t st[3]; int w[3], h[3];
    planes(cs, p, st, w, h);
    for (int c = 0; c < 3; c++)
      if (p[c]) cap.addPlane((std::string(stage) + "_" + std::to_string(c)).c_str(), p[c], st[c], w[c], h[c]);
  }
  void finishPicture(CodingStructure& cs, int stagesMask)
  {
    if (!captureDir.empty())
    {
      const int32_t hdr[8] = { seq.width, seq.height, seq.chroma_format, seq.bit_depth_luma, seq.bit_depth_chroma, seq.ctu_size, cs.slice->getPOC(), stagesMask };
      cap.add("seq", hdr, sizeof(hdr));
      char name[64];
      snprintf(name, sizeof(name), "/pic%04d_poc%03d.cap", picCount, cs.slice->getPOC());
      if (!cap.write(captureDir + name)) THROW("vtmgpu shim: cannot write capture file in " << captureDir);
      cap.clear();
    }
    picCount++;
  }
};

Shim& shim()
{
  static Shim s;
  return s;
}

bool lastStage(const CodingStructure& cs, int stage)   // 0 dbf, 1 sao, 2 alf
{
  const bool sao = cs.sps->getSAOEnabledFlag(), alf = cs.sps->getALFEnabledFlag();
  return stage == 2 || (stage == 1 && !alf) || (stage == 0 && !sao && !alf);
}

}   // namespace

// ---------------------------------------------------------------------------------------------------------
// LMCS table mapping of a sample block (declared in Buffer.h, defined in the reference's Buffer.cpp:380-393).  This definition takes
// the reference's place at link time (the shim objects precede libCommonNoFilters.a, -Wl,--allow-multiple-definition; a maintainer
// would guard the one call in DecLib.cpp instead).  The decoder calls it per CU with the forward table
// (DecCu.cpp:698,742,765) -- done here exactly as the reference does -- and ONCE per picture with the inverse table on the whole luma
// reconstruction right before loopFilterPic (DecLib.cpp:570-577): a block larger than any CU is that call, and the product shim only
// records it; k_dbf_sao maps the samples while it loads its tiles (vtmgpu_set_lmcs).
// ---------------------------------------------------------------------------------------------------------
#ifndef VTMGPU_SHIM_ENCODER
template<>
void AreaBuf<Pel>::rspSignal(std::vector<Pel>& pLUT)
{
  Shim& s = shim();
  if (s.deferLmcs && (width > MAX_CU_SIZE || height > MAX_CU_SIZE))
  {
    s.lmcsLut = pLUT;
    s.lmcsBuf = buf;
    s.lmcsPending = true;
    return;
  }
  if (width > MAX_CU_SIZE || height > MAX_CU_SIZE) s.lmcsHostPics++;
  const Pel* lut = pLUT.data();
  for (unsigned y = 0; y < height; y++)
  {
    Pel* row = buf + (ptrdiff_t)y * stride;
    for (unsigned x = 0; x < width; x++) row[x] = lut[row[x]];
  }
}

// ---------------------------------------------------------------------------------------------------------
// SAOOffset / SAOBlkParam (declared in TypeDef.h:938-963, defined in the reference's SampleAdaptiveOffset.cpp)
// ---------------------------------------------------------------------------------------------------------
SAOOffset::SAOOffset() { reset(); }
SAOOffset::~SAOOffset() {}
void SAOOffset::reset()
{
  modeIdc = SAO_MODE_OFF;
  typeIdc = typeAuxInfo = -1;
  std::fill(offset, offset + MAX_NUM_SAO_CLASSES, 0);
}
const SAOOffset& SAOOffset::operator=(const SAOOffset& src)
{
  modeIdc = src.modeIdc;
  typeIdc = src.typeIdc;
  typeAuxInfo = src.typeAuxInfo;
  std::copy(src.offset, src.offset + MAX_NUM_SAO_CLASSES, offset);
  return *this;
}
SAOBlkParam::SAOBlkParam() { reset(); }
SAOBlkParam::~SAOBlkParam() {}
void SAOBlkParam::reset() { for (auto& o : offsetParam) o.reset(); }
const SAOBlkParam& SAOBlkParam::operator=(const SAOBlkParam& src)
{
  for (int c = 0; c < MAX_NUM_COMPONENT; c++) offsetParam[c] = src.offsetParam[c];
  return *this;
}

#endif   // !VTMGPU_SHIM_ENCODER

// ---------------------------------------------------------------------------------------------------------
// LoopFilter
// ---------------------------------------------------------------------------------------------------------
#ifdef VTMGPU_SHIM_ENCODER
// Encoder-side reuse (EncGOP.cpp:2794, :3436 call the same LoopFilter::loopFilterPic): the encoder links this object for
// its LoopFilter and keeps the reference's SampleAdaptiveOffset / AdaptiveLoopFilter objects (its SAO / ALF parameter searches live in
// classes derived from them).  Besides the five symbols the decoder needs, EncoderLib references three more:
//   sm_betaTable         read by the encoder's deblocking-parameter heuristics through the inline getBeta() (LoopFilter.h:128)
//   initEncPicYuvBuffer  per-picture allocation of the buffer of the deblocking-aware RD optimisation (EncGOP.cpp:2761; the buffer is only
//                        touched with EncDbOpt = 1, off in the CTC configurations): nothing to allocate here
//   xDeblockCU           called by that optimisation per CU -- it needs filtered samples of a partial picture on the host: refused loudly
const uint8_t LoopFilter::sm_betaTable[MAX_QP + 1] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,
                                                       42,44,46,48,50,52,54,56,58,60,62,64,66,68,70,72,74,76,78,80,82,84,86,88 };
void LoopFilter::initEncPicYuvBuffer(ChromaFormat, int, int) {}
void LoopFilter::xDeblockCU(CodingUnit&, const DeblockEdgeDir) { THROW("vtmgpu shim: EncDbOpt (deblocking inside the RD search) is not available with the GPU deblocking filter"); }
#endif
LoopFilter::LoopFilter() {}
LoopFilter::~LoopFilter() {}
void LoopFilter::create(const unsigned uiMaxCUDepth)
{
  m_enc = false;
  if (shim().useRef) SHIM_ALT(lfCreate(uiMaxCUDepth));
}
void LoopFilter::destroy() {}

void LoopFilter::loopFilterPic(CodingStructure& cs)
{
  Shim& s = shim();
  s.ensureCtx(cs);
  s.saoPending = s.dbfPending = false;
  s.lumaPixels += (long long)cs.pcv->lumaWidth * cs.pcv->lumaHeight;
  if (s.lmcsPending)
  {
    // executeLoopFilters asked for the inverse mapping of THIS picture a moment ago (DecLib.cpp:574): the device applies it
    CHECK(s.lmcsBuf != cs.getRecoBuf().Y().buf, "vtmgpu shim: deferred LMCS mapping belongs to another buffer");
    s.check(s.api.set_lmcs(s.ctx, 0, s.lmcsLut.data(), (int)s.lmcsLut.size()), "set_lmcs");
    s.lmcsPending = false;
    s.lmcsActive = true;
    s.lmcsPics++;
  }
  else if (s.lmcsActive)
  {
    s.check(s.api.set_lmcs(s.ctx, 0, nullptr, 0), "set_lmcs");
    s.lmcsActive = false;
  }
  // the upload of the reconstruction (a staged copy out of pageable decoder memory) runs beside the host derivation
  std::thread uploader;
  std::exception_ptr uploadError;
  // device-side derivation: the motion field (read before setRefinedMotionField changes it, DecLib.cpp:579-580) goes down behind the planes
  // while this thread flattens the block structure
  bool motionPreloaded = false;
  const void* miBuf = nullptr;
  int miPitch = 0;
  if (s.deviceDerive)
  {
    bool inter = cs.sps->getIBCFlag();
    for (const Slice* sl : cs.picture->slices) inter |= !sl->isIntra();
    if (inter)
    {
      const CMotionBuf mb = const_cast<const CodingStructure&>(cs).getMotionBuf();
      miBuf = mb.buf; miPitch = (int)mb.stride;
      s.pinOnce(miBuf, (size_t)miPitch * (cs.pcv->lumaHeight / 4) * sizeof(MotionInfo));
      motionPreloaded = true;
    }
  }
  if (!s.useRef) uploader = std::thread([&] {
    try
    {
      s.upload(cs);
      if (motionPreloaded)
      {
        s.check(s.api.upload_motion(s.ctx, 0, miBuf, (int)sizeof(MotionInfo), miPitch, (int)offsetof(MotionInfo, mv), (int)(offsetof(MotionInfo, mv) + sizeof(Mv)),
                                    (int)offsetof(MotionInfo, refIdx), (int)(offsetof(MotionInfo, refIdx) + sizeof(int16_t))), "upload_motion");
        s.check(s.api.sync(s.ctx), "sync");
      }
    }
    catch (...) { uploadError = std::current_exception(); }
  });
  bool useUnits = false;
  try
  {
    if (s.deviceDerive || s.checkUnits)
    {
      s.tic();
      const size_t unitsNow = (size_t)(cs.pcv->lumaWidth / 4) * (cs.pcv->lumaHeight / 4);
      if (s.units.block && s.units.units != unitsNow && !s.useRef) s.unpin(s.units.block);      // the tables are about to be reallocated
      flattenUnits(cs, s.units);
      s.toc(s.flattenSec);
      if (s.deviceDerive && s.ctx)
      {
        if (s.units.blockIsNew) s.pinOnce(s.units.block, s.units.blockBytes);
        if (s.units.p.motion) s.pinOnce(s.units.p.motion, (size_t)s.units.p.motion_pitch * (cs.pcv->lumaHeight / 4) * s.units.p.motion_elem_bytes);
      }
      useUnits = s.deviceDerive && s.units.supported;
      if (s.deviceDerive) (useUnits ? s.unitPics : s.unitFallbackPics)++;
    }
    if ((!useUnits || s.checkUnits) && (!s.useRef || !s.captureDir.empty() || !s.timing || s.checkUnits))   // the reference backend derives its own parameters; skip ours when only timing it
    {
      s.tic();
      deriveDeblockRecords(cs, s.dbf);
      s.toc(s.deriveSec);
    }
    if (s.checkUnits && s.units.supported)
    {
      // the unit derivation (the kernel's source, run here on the host) against the CU walk, record by record
      deriveFromUnits(s.units, cs.pcv->lumaWidth, cs.pcv->lumaHeight, (int)cs.pcv->chrFormat, cs.sps->getBitDepth(CHANNEL_TYPE_LUMA), cs.sps->getBitDepth(CHANNEL_TYPE_CHROMA),
                      cs.pcv->maxCUWidth, s.dbfFromUnits);
      for (int d = 0; d < 2; d++)
      {
        for (size_t i = 0; i < s.dbf.luma[d].size(); i++)
          if (s.dbf.luma[d][i] != s.dbfFromUnits.luma[d][i])
            THROW("vtmgpu shim: unit derivation differs from the CU walk: POC " << cs.slice->getPOC() << " luma dir " << d << " unit (" << (i % (cs.pcv->lumaWidth / 4)) << "," << (i / (cs.pcv->lumaWidth / 4))
                  << ") walk " << s.dbf.luma[d][i] << " units " << s.dbfFromUnits.luma[d][i] << " | tuQ " << s.units.tus[s.units.tuLuma[i]].x << "," << s.units.tus[s.units.tuLuma[i]].y << " "
                  << (int)s.units.tus[s.units.tuLuma[i]].w << "x" << (int)s.units.tus[s.units.tuLuma[i]].h << " cu " << s.units.cus[s.units.tus[s.units.tuLuma[i]].cu].x << "," << s.units.cus[s.units.tus[s.units.tuLuma[i]].cu].y
                  << " " << (int)s.units.cus[s.units.tus[s.units.tuLuma[i]].cu].w4 << "x" << (int)s.units.cus[s.units.tus[s.units.tuLuma[i]].cu].h4 << " flags " << s.units.cus[s.units.tus[s.units.tuLuma[i]].cu].flags);
        for (size_t i = 0; i < s.dbf.chroma[d].size(); i++)
          if (s.dbf.chroma[d][i] != s.dbfFromUnits.chroma[d][i])
            THROW("vtmgpu shim: unit derivation differs from the CU walk: POC " << cs.slice->getPOC() << " chroma dir " << d << " index " << i
                  << " walk " << s.dbf.chroma[d][i] << " units " << s.dbfFromUnits.chroma[d][i]);
      }
      s.checkedPics++;
    }
  }
  catch (...) { if (uploader.joinable()) uploader.join(); throw; }
  if (uploader.joinable()) uploader.join();
  if (uploadError) std::rethrow_exception(uploadError);
  if (!s.captureDir.empty())
  {
    s.cap.clear();
    CHECK(s.dbf.listsValid && !s.dbf.listsMatchDense(), "vtmgpu shim: record lists and record arrays disagree");
    s.capturePlanes(cs, "pre");
    s.cap.add("dbfrec_l0", s.dbf.luma[0].data(), s.dbf.luma[0].size() * 4);
    s.cap.add("dbfrec_l1", s.dbf.luma[1].data(), s.dbf.luma[1].size() * 4);
    s.cap.add("dbfrec_c0", s.dbf.chroma[0].data(), s.dbf.chroma[0].size() * 8);
    s.cap.add("dbfrec_c1", s.dbf.chroma[1].data(), s.dbf.chroma[1].size() * 8);
    if (s.dbf.hasLadf) s.cap.add("dbf_ladf", &s.dbf.ladf, sizeof(s.dbf.ladf));
    const vtmgpu_virtual_boundaries vb = flattenVirtualBoundaries(cs);
    if (vb.num_ver || vb.num_hor) s.cap.add("vb", &vb, sizeof(vb));
  }
  s.tic();
  if (s.useRef)
  {
    SHIM_ALT(lfRun(cs));
  }
  else
  {
    if (useUnits)
    {
      // the block structure goes down and k_dbf_derive writes the records; the motion field is read before setRefinedMotionField
      // changes it (DecLib.cpp:579-580): the copies are complete when this call returns
      vtmgpu_deblock_units up = *s.units.view();
      if (motionPreloaded) { up.flags |= VTMGPU_UNITS_MOTION_PRELOADED; up.motion = nullptr; }
      const bool motionInCall = up.motion != nullptr;      // the decoder changes its motion field right after this call returns
      s.check(s.api.set_deblock_units(s.ctx, 0, &up), "set_deblock_units");
      // the tables live in the shim's own page-locked block, which is not written again before this picture's last download has
      // completed: no need to wait here -- the copies and k_dbf_derive run while the host packs the SAO / ALF parameters
      if (motionInCall || s.checkUnits) s.check(s.api.sync(s.ctx), "sync");
      if (s.checkUnits)
      {
        // the kernel's records against the CU walk's
        FlatDeblock& k = s.dbfFromUnits;
        uint32_t* l[2] = { k.luma[0].data(), k.luma[1].data() };
        uint64_t* cch[2] = { k.chroma[0].empty() ? nullptr : k.chroma[0].data(), k.chroma[1].empty() ? nullptr : k.chroma[1].data() };
        for (int d = 0; d < 2; d++) { std::fill(k.luma[d].begin(), k.luma[d].end(), 0xffffffffu); std::fill(k.chroma[d].begin(), k.chroma[d].end(), ~0ull); }
        s.check(s.api.get_deblock_records(s.ctx, 0, l, cch), "get_deblock_records");
        for (int d = 0; d < 2; d++)
        {
          // (records on the picture border are never filtered; the CU walk leaves them zero as well)
          CHECK(k.luma[d] != s.dbf.luma[d], "vtmgpu shim: k_dbf_derive differs from the CU walk (luma records, POC " << cs.slice->getPOC() << ")");
          CHECK(k.chroma[d] != s.dbf.chroma[d], "vtmgpu shim: k_dbf_derive differs from the CU walk (chroma records, POC " << cs.slice->getPOC() << ")");
        }
        s.deviceCheckedPics++;
      }
    }
    else if (s.dbf.listsValid && !s.denseRecords)
    {
      const vtmgpu_deblock_sparse p = s.dbf.sparseView();      // only the active units cross the bus
      s.listPics++;
      for (int d = 0; d < 2; d++) s.recordBytes += (long long)p.luma_count[d] * sizeof(vtmgpu_dbf_luma_entry) + (long long)p.chroma_count[d] * sizeof(vtmgpu_dbf_chroma_entry);
      s.check(s.api.set_deblock_sparse(s.ctx, 0, &p), "set_deblock_sparse");
    }
    else
    {
      const vtmgpu_deblock_params p = s.dbf.view();
      for (int d = 0; d < 2; d++) s.recordBytes += (long long)s.dbf.luma[d].size() * 4 + (long long)s.dbf.chroma[d].size() * 8;
      s.check(s.api.set_deblock(s.ctx, 0, &p), "set_deblock");
    }
    s.check(s.api.set_sao(s.ctx, 0, nullptr), "set_sao");
    s.check(s.api.set_alf(s.ctx, 0, nullptr), "set_alf");
    s.dbfPending = !(s.staged || lastStage(cs, 0));
    if (!s.dbfPending)
    {
      s.check(s.api.deblock(s.ctx, 0, 1), "deblock");
      s.download(cs, lastStage(cs, 0));
    }
  }
  s.toc(s.stageSec[0]);
  s.capturePlanes(cs, "dbf");
  if (lastStage(cs, 0)) s.finishPicture(cs, 1);
}

#ifndef VTMGPU_SHIM_ENCODER
// ---------------------------------------------------------------------------------------------------------
// SampleAdaptiveOffset
// ---------------------------------------------------------------------------------------------------------
SampleAdaptiveOffset::SampleAdaptiveOffset() { m_numberOfComponents = 0; m_pcReshape = nullptr; }
SampleAdaptiveOffset::~SampleAdaptiveOffset() { destroy(); }
void SampleAdaptiveOffset::create(int picWidth, int picHeight, ChromaFormat format, uint32_t maxCUWidth, uint32_t maxCUHeight, uint32_t maxCUDepth,
                                  uint32_t lumaBitShift, uint32_t chromaBitShift)
{
  for (int c = 0; c < MAX_NUM_COMPONENT; c++) m_offsetStepLog2[c] = isLuma(ComponentID(c)) ? lumaBitShift : chromaBitShift;
  m_numberOfComponents = getNumberValidComponents(format);
  shim().saoLog2Scale[0] = lumaBitShift;
  shim().saoLog2Scale[1] = chromaBitShift;
  if (shim().useRef) SHIM_ALT(saoCreate(picWidth, picHeight, (int)format, maxCUWidth, maxCUHeight, maxCUDepth, lumaBitShift, chromaBitShift));
}
void SampleAdaptiveOffset::destroy() {}

void SampleAdaptiveOffset::SAOProcess(CodingStructure& cs, SAOBlkParam* saoBlkParams)
{
  CHECK(!saoBlkParams, "No parameters present");
  Shim& s = shim();
  flattenSao(cs, saoBlkParams, s.saoLog2Scale[0], s.saoLog2Scale[1], s.sao);
  if (!s.captureDir.empty())
  {
    s.cap.add("sao_raw", s.sao.ctu.data(), s.sao.ctu.size() * sizeof(vtmgpu_sao_ctu));
    s.cap.add("sao_scale", s.saoLog2Scale, sizeof(s.saoLog2Scale));
  }
  s.tic();
  if (s.useRef)
  {
    SHIM_ALT(saoRun(cs, saoBlkParams));
  }
  else
  {
    const int mask = s.api.sao_reconstruct(s.sao.ctu.data(), (int)s.sao.ctu.size(), s.sao.widthInCtus, s.sao.numComps, s.saoLog2Scale[0], s.saoLog2Scale[1]);
    if (mask < 0) THROW("vtmgpu shim: invalid SAO parameters (" << mask << ")");
    writeBackSao(s.sao, saoBlkParams);
    const vtmgpu_sao_params p = s.sao.view();
    s.check(s.api.set_sao(s.ctx, 0, &p), "set_sao");
    if (s.staged || lastStage(cs, 1))
    {
      if (s.dbfPending) s.check(s.api.deblock_sao(s.ctx, 0, 1), "deblock_sao");     // one kernel: deblocking with the SAO epilogue
      else              s.check(s.api.sao(s.ctx, 0, 1), "sao");
      s.dbfPending = false;
      s.download(cs, lastStage(cs, 1));
    }
    else
    {
      s.saoPending = true;     // fused into the ALF pass
    }
  }
  s.toc(s.stageSec[1]);
  s.capturePlanes(cs, "sao");
  if (lastStage(cs, 1)) s.finishPicture(cs, 3);
}

// ---------------------------------------------------------------------------------------------------------
// AdaptiveLoopFilter
// ---------------------------------------------------------------------------------------------------------
AdaptiveLoopFilter::AdaptiveLoopFilter() : m_classifier(nullptr)
{
  for (int c = 0; c < MAX_NUM_COMPONENT; c++) { m_ctuEnableFlag[c] = nullptr; m_ctuAlternative[c] = nullptr; }
  m_ccAlfFilterControl[0] = m_ccAlfFilterControl[1] = nullptr;
  m_deriveClassificationBlk = nullptr;
  m_filterCcAlf = nullptr;
  m_filter5x5Blk = m_filter7x7Blk = nullptr;
}

void AdaptiveLoopFilter::create(const int picWidth, const int picHeight, const ChromaFormat format, const int maxCUWidth, const int maxCUHeight,
                                const int maxCUDepth, const int inputBitDepth[MAX_NUM_CHANNEL_TYPE])
{
  // called again for every picture (DecLib.cpp:1151): idempotent while the geometry is unchanged
  const int ctus = ((picWidth + maxCUWidth - 1) / maxCUWidth) * ((picHeight + maxCUHeight - 1) / maxCUHeight);
  if (m_created && (picWidth != m_picWidth || picHeight != m_picHeight || maxCUWidth != m_maxCUWidth || ctus != m_numCTUsInPic)) destroy();
  m_inputBitDepth[0] = inputBitDepth[0];
  m_inputBitDepth[1] = inputBitDepth[1];
  m_picWidth = picWidth; m_picHeight = picHeight; m_maxCUWidth = maxCUWidth; m_maxCUHeight = maxCUHeight; m_maxCUDepth = maxCUDepth;
  m_chromaFormat = format;
  m_numCTUsInPic = ctus;
  if (shim().useRef) SHIM_ALT(alfCreate(picWidth, picHeight, (int)format, maxCUWidth, maxCUHeight, maxCUDepth, inputBitDepth));
  if (m_created) return;
  m_ccAlfFilterControl[0] = new uint8_t[ctus]();     // CABACReader writes the per-CTU CC-ALF idc here (DecLib.cpp:1154-1155)
  m_ccAlfFilterControl[1] = new uint8_t[ctus]();
  m_created = true;
}

void AdaptiveLoopFilter::destroy()
{
  if (!m_created) return;
  delete[] m_ccAlfFilterControl[0];
  delete[] m_ccAlfFilterControl[1];
  m_ccAlfFilterControl[0] = m_ccAlfFilterControl[1] = nullptr;
  m_created = false;
}

void AdaptiveLoopFilter::ALFProcess(CodingStructure& cs)
{
  Shim& s = shim();
  flattenAlf(cs, m_ccAlfFilterParam, m_ccAlfFilterControl, s.alf);
  const vtmgpu_alf_params* p = s.alf.view();
  const bool perSlice = !s.alf.more.empty();       // slices with different ALF data: not a case the capture format or the reference-encoded streams hold
  CHECK(perSlice && !s.captureDir.empty(), "vtmgpu shim: the capture format does not cover pictures whose slices carry different ALF parameters");
  if (!s.captureDir.empty())
  {
    const int32_t hdr[8] = { p->enabled[0], p->enabled[1], p->enabled[2], p->num_luma_aps, s.alf.hasChromaAps, p->ccalf_enabled[0], p->ccalf_enabled[1], p->num_ctus };
    s.cap.add("alf_hdr", hdr, sizeof(hdr));
    s.cap.add("alf_luma_aps", s.alf.lumaAps.data(), s.alf.lumaAps.size() * sizeof(vtmgpu_alf_luma_aps));
    s.cap.add("alf_chroma_aps", &s.alf.chromaAps, sizeof(vtmgpu_alf_chroma_aps));
    for (int c = 0; c < 3; c++) s.cap.add(("alf_en" + std::to_string(c)).c_str(), s.alf.ctuEnable[c].data(), s.alf.ctuEnable[c].size());
    s.cap.add("alf_fidx", s.alf.filterIdx.data(), s.alf.filterIdx.size() * 2);
    for (int c = 0; c < 2; c++) s.cap.add(("alf_alt" + std::to_string(c)).c_str(), s.alf.ctuAlt[c].data(), s.alf.ctuAlt[c].size());
    s.cap.add("alf_cccoef", p->ccalf_coeff, sizeof(p->ccalf_coeff));
    if (!s.alf.ctuClip.empty()) s.cap.add("alf_clip", s.alf.ctuClip.data(), s.alf.ctuClip.size());
    for (int c = 0; c < 2; c++) s.cap.add(("alf_ccidc" + std::to_string(c)).c_str(), s.alf.ccIdc[c].data(), s.alf.ccIdc[c].size());
  }
  s.tic();
  if (s.useRef)
  {
    SHIM_ALT(alfRun(cs, m_ccAlfFilterParam, m_ccAlfFilterControl, p->num_ctus));
  }
  else
  {
    if (perSlice) s.check(s.api.set_alf_slices(s.ctx, 0, 1 + (int)s.alf.more.size(), s.alf.slicesView(), s.alf.ctuSlice.data()), "set_alf_slices");
    else          s.check(s.api.set_alf(s.ctx, 0, p), "set_alf");
    if (s.dbfPending)      s.check(s.api.filter(s.ctx, 0, 1), "filter");        // whole chain: two kernels, one synchronisation
    else if (s.saoPending) s.check(s.api.sao_alf(s.ctx, 0, 1), "sao_alf");
    else                   s.check(s.api.alf(s.ctx, 0, 1), "alf");
    s.saoPending = s.dbfPending = false;
    s.download(cs, true);
  }
  s.toc(s.stageSec[2]);
  s.capturePlanes(cs, "alf");
  s.finishPicture(cs, cs.sps->getSAOEnabledFlag() ? 7 : 5);
}

// InitX86.cpp (kept in the host build for the non-filter SIMD) still references the per-ISA ALF initialisers that
// lived in the removed x86/*/AdaptiveLoopFilter_*.cpp stubs; they are never called (our constructor does not
// dispatch), so empty definitions satisfy the linker.
#ifdef TARGET_SIMD_X86
template<X86_VEXT vext> void AdaptiveLoopFilter::_initAdaptiveLoopFilterX86() {}
template void AdaptiveLoopFilter::_initAdaptiveLoopFilterX86<SSE41>();
template void AdaptiveLoopFilter::_initAdaptiveLoopFilterX86<AVX>();
template void AdaptiveLoopFilter::_initAdaptiveLoopFilterX86<AVX2>();
#endif
#endif   // !VTMGPU_SHIM_ENCODER
