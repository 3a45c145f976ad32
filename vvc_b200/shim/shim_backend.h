// shim_backend.h -- optional alternative backend for the shim (test infrastructure hook).
//
// The product shim runs the three stages on libvtmgpu.  A test binary may additionally link an object that
// defines vtmgpu_shim_alt_backend(); with VTMGPU_SHIM_BACKEND=ref the shim then hands each stage to it instead
// (oracle/ref_backend.cpp runs the reference's own, macro-renamed filter classes).  The product binary does not
// define the symbol (it is weak), so the hook is inert there.
#pragma once

#include <cstdint>

class CodingStructure;
struct CcAlfFilterParam;

struct VtmgpuShimAltBackend
{
  void (*lfCreate)(unsigned maxCUDepth);
  void (*lfRun)(CodingStructure& cs);
  void (*saoCreate)(int picWidth, int picHeight, int chromaFormat, uint32_t maxCUWidth, uint32_t maxCUHeight, uint32_t maxCUDepth,
                    uint32_t lumaBitShift, uint32_t chromaBitShift);
  void (*saoRun)(CodingStructure& cs, void* saoBlkParams);
  void (*alfCreate)(int picWidth, int picHeight, int chromaFormat, int maxCUWidth, int maxCUHeight, int maxCUDepth, const int bitDepth[2]);
  void (*alfRun)(CodingStructure& cs, const CcAlfFilterParam& cc, uint8_t* const ccControl[2], int numCtus);
};

extern "C" const VtmgpuShimAltBackend* vtmgpu_shim_alt_backend() __attribute__((weak));
