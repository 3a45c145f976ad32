// oracle/ref_backend.cpp -- TEST INFRASTRUCTURE.  Alternative shim backend that runs the REFERENCE's own filter
// classes: LoopFilter.cpp / SampleAdaptiveOffset.cpp / AdaptiveLoopFilter.cpp compiled unmodified with the class
// names macro-renamed to Ref* (oracle/Makefile, libRefFilters.a).  Linked only into oracle/_ref/DecoderApp_cap,
// which is used to (a) produce the captures / golden fixtures and (b) A/B the host-side derivation of the shim.
// This translation unit is compiled with the same rename macros, so it only ever sees the Ref* names.
#include "AdaptiveLoopFilter.h"
#include "CodingStructure.h"
#include "LoopFilter.h"
#include "SampleAdaptiveOffset.h"

#include <cstdlib>
#include <cstring>

#include "shim_backend.h"

namespace
{
LoopFilter           g_lf;     // = RefLoopFilter
SampleAdaptiveOffset g_sao;    // = RefSampleAdaptiveOffset
AdaptiveLoopFilter   g_alf;    // = RefAdaptiveLoopFilter

void lfCreate(unsigned depth) { g_lf.create(depth); }
void lfRun(CodingStructure& cs) { g_lf.loopFilterPic(cs); }
void saoCreate(int w, int h, int cf, uint32_t cw, uint32_t ch, uint32_t depth, uint32_t ls, uint32_t cs_) { g_sao.create(w, h, ChromaFormat(cf), cw, ch, depth, ls, cs_); }
void saoRun(CodingStructure& cs, void* blk) { g_sao.SAOProcess(cs, reinterpret_cast<SAOBlkParam*>(blk)); }
void alfCreate(int w, int h, int cf, int cw, int ch, int depth, const int bd[2]) { g_alf.create(w, h, ChromaFormat(cf), cw, ch, depth, bd); }
void alfRun(CodingStructure& cs, const CcAlfFilterParam& cc, uint8_t* const ctl[2], int numCtus)
{
  g_alf.getCcAlfFilterParam() = cc;
  for (int c = 0; c < 2; c++) memcpy(g_alf.getCcAlfControlIdc(ComponentID(c + 1)), ctl[c], numCtus);
  g_alf.ALFProcess(cs);
}
const VtmgpuShimAltBackend g_backend = { lfCreate, lfRun, saoCreate, saoRun, alfCreate, alfRun };
}   // namespace

extern "C" const VtmgpuShimAltBackend* vtmgpu_shim_alt_backend() { return &g_backend; }

// The reference's ISA dispatch for ALF lives in x86/InitX86.cpp, which is linked un-renamed for the rest of the
// decoder; the renamed class needs its own copy of the (trivial) dispatch: pick the widest ALF SIMD the CPU offers,
// exactly what the stock decoder does (InitX86.cpp:147-168) -- this is what makes the Ref* path the "x86-SIMD" one.
#ifdef TARGET_SIMD_X86
void AdaptiveLoopFilter::initAdaptiveLoopFilterX86()
{
  // VTMGPU_REF_SIMD=SCALAR: leave the C++ filter routines installed (what the stock decoder does with --SIMD=SCALAR) -- bench.py's
  // cpu_baseline reports that figure beside the SIMD one
  const char* e = getenv("VTMGPU_REF_SIMD");
  if (e && !strcmp(e, "SCALAR")) return;
  const X86_VEXT v = read_x86_extension_flags();
  if (v >= AVX2)       _initAdaptiveLoopFilterX86<AVX2>();
  else if (v == AVX)   _initAdaptiveLoopFilterX86<AVX>();
  else if (v >= SSE41) _initAdaptiveLoopFilterX86<SSE41>();
}
#endif
