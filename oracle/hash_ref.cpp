// oracle/hash_ref.cpp -- TEST INFRASTRUCTURE.  The reference's OWN decoded-picture hash functions (calcMD5 / calcCRC /
// calcChecksum, CommonLib/PicYuvMD5.cpp:130,169,188) on raw planes, so that the device hashes of libvtmgpu (vtmgpu_hash)
// can be pinned against them on arbitrary pictures.  Built by oracle/Makefile into oracle/_ref/hash_ref from the reference
// sources where they lie; only tests execute it.
//
//   hash_ref FILE     FILE = int32 header {width, height, chroma_format, bit_depth_luma, bit_depth_chroma} + int16 planes, dense
//   prints three lines: "md5 <hex>", "crc <hex>", "checksum <hex>"  (digests of all components concatenated, as in the SEI)
#include <cstdio>
#include <vector>

#include "CommonLib/CommonDef.h"
#include "CommonLib/Buffer.h"
#include "CommonLib/Picture.h"
#include "CommonLib/Unit.h"

// defined in CommonLib/PicYuvMD5.cpp (no header declares them; EncoderLib/SEIEncoder.cpp:39-41 declares them the same way)
uint32_t calcMD5(const CPelUnitBuf& pic, PictureHash& digest, const BitDepths& bitDepths);
uint32_t calcCRC(const CPelUnitBuf& pic, PictureHash& digest, const BitDepths& bitDepths);
uint32_t calcChecksum(const CPelUnitBuf& pic, PictureHash& digest, const BitDepths& bitDepths);

int main(int argc, char** argv)
{
  if (argc < 2) { fprintf(stderr, "usage: hash_ref FILE\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  int32_t hdr[5];
  if (fread(hdr, sizeof(hdr), 1, f) != 1) return 2;
  const int w = hdr[0], h = hdr[1];
  const ChromaFormat cf = ChromaFormat(hdr[2]);
  const UnitArea area(cf, Area(0, 0, w, h));
  std::vector<std::vector<Pel>> mem(area.blocks.size());
  PelUnitBuf pic;
  pic.chromaFormat = cf;
  for (size_t c = 0; c < area.blocks.size(); c++)
  {
    const CompArea& b = area.blocks[c];
    mem[c].resize((size_t)b.width * b.height);
    if (fread(mem[c].data(), sizeof(Pel), mem[c].size(), f) != mem[c].size()) { fprintf(stderr, "short file\n"); return 2; }
    pic.bufs.push_back(PelBuf(mem[c].data(), b.width, b.width, b.height));
  }
  fclose(f);
  BitDepths bd;
  bd.recon[CHANNEL_TYPE_LUMA] = hdr[3];
  bd.recon[CHANNEL_TYPE_CHROMA] = hdr[4];
  const char* names[3] = { "md5", "crc", "checksum" };
  for (int k = 0; k < 3; k++)
  {
    PictureHash d;
    if (k == 0) calcMD5(pic, d, bd); else if (k == 1) calcCRC(pic, d, bd); else calcChecksum(pic, d, bd);
    printf("%s ", names[k]);
    for (uint8_t b : d.hash) printf("%02x", b);
    printf("\n");
  }
  return 0;
}
