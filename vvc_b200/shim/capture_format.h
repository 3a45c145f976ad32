// capture_format.h -- "VTMGCAP1" container: everything that crosses the drop-in boundary for ONE picture
// (pre-filter planes + flattened side information) and, optionally, the planes after each stage.
// Written by the shim (VTMGPU_CAPTURE_DIR=...), read by vvc_b200/capture.py for replay, tests and bench.
//
//   file    := "VTMGCAP1" u32 nsections { section }
//   section := char name[24] (NUL padded)  u64 nbytes  payload  (padded to a multiple of 8 bytes)
//
// sections: seq (int32[8]: width height chroma_format bd_luma bd_chroma ctu_size poc stages_mask)
//           pre_0..2 | dbf_0..2 | sao_0..2 | alf_0..2   int16 planes, tightly packed
//           dbfrec_l0 dbfrec_l1 (u32)  dbfrec_c0 dbfrec_c1 (u64)
//           sao_raw (vtmgpu_sao_ctu[], as parsed)  sao_scale (int32[2])
//           alf_hdr (int32[8]: enabled[3] num_luma_aps has_chroma_aps ccalf_enabled[2] num_ctus)
//           alf_luma_aps alf_chroma_aps (raw structs)  alf_en0..2 alf_fidx alf_alt0..1 alf_cccoef alf_ccidc0..1
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace vtmshim
{

class CaptureWriter
{
public:
  void add(const char* name, const void* data, size_t nbytes)
  {
    Sec s;
    memset(s.name, 0, sizeof(s.name));
    strncpy(s.name, name, sizeof(s.name) - 1);
    s.data.assign((const uint8_t*)data, (const uint8_t*)data + nbytes);
    m_secs.push_back(std::move(s));
  }
  // tightly packs a strided int16 plane
  void addPlane(const char* name, const int16_t* p, ptrdiff_t stride, int w, int h)
  {
    std::vector<int16_t> t((size_t)w * h);
    for (int y = 0; y < h; y++) memcpy(&t[(size_t)y * w], p + y * stride, sizeof(int16_t) * w);
    add(name, t.data(), t.size() * sizeof(int16_t));
  }
  bool write(const std::string& path) const
  {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const uint32_t n = (uint32_t)m_secs.size();
    fwrite("VTMGCAP1", 1, 8, f);
    fwrite(&n, 4, 1, f);
    static const uint8_t zero[8] = { 0 };
    for (const Sec& s : m_secs)
    {
      const uint64_t nb = s.data.size();
      fwrite(s.name, 1, sizeof(s.name), f);
      fwrite(&nb, 8, 1, f);
      fwrite(s.data.data(), 1, s.data.size(), f);
      fwrite(zero, 1, (8 - nb % 8) % 8, f);
    }
    return fclose(f) == 0;
  }
  void clear() { m_secs.clear(); }
  bool empty() const { return m_secs.empty(); }

private:
  struct Sec { char name[24]; std::vector<uint8_t> data; };
  std::vector<Sec> m_secs;
};

}   // namespace vtmshim
