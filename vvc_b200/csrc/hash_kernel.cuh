// hash_kernel.cuh -- decoded-picture hashes on the device: MD5, CRC and checksum of the planes of a picture
// slot as the decoded picture hash SEI defines them, so that a picture can be verified without leaving HBM.
//
//   calcMD5 / md5_plane   CommonLib/PicYuvMD5.cpp:130-212,66-86   per component, samples in raster order, little-endian,
//                                                                   one byte per sample up to 8 bits, two bytes above
//   calcCRC / compCRC     CommonLib/PicYuvMD5.cpp:88-142           CRC-16 (x^16 + x^12 + x^5 + 1), initial value 0xffff, per sample the
//                                                                   low byte then (bit depth > 8) the high byte, MSB first, 16 flush bits
//   calcChecksum          CommonLib/PicYuvMD5.cpp:144-186          sum of (byte ^ mask(x, y)) mod 2^32
//
// Checksum and CRC are reductions: the checksum is an integer sum; the CRC is linear over GF(2), so every thread takes the
// remainder of its own run of samples and multiplies it by x^(bits that follow) mod the polynomial, and the results are xored.
// MD5 is a serial chain per component -- one lane per (slot, component) runs the rounds while the whole warp stages the next
// 2 KB of the component's byte stream in shared memory; the three components (and all slots of a batch) run in parallel.
#pragma once

#include "vtmgpu_dev.cuh"

namespace vtmgpu
{

struct HashPlane
{
  const pel* p;
  int pitch, w, h, bd;
  uint32_t* out;                 // MD5: 4 words ; CRC: 1 word (xor-accumulated) ; checksum: 1 word (sum)
};

// ---- checksum -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_hash_checksum(const HashPlane* __restrict__ planes, int nPlanes)
{
  const HashPlane P = planes[blockIdx.y];
  uint32_t sum = 0;
  const int groups = P.w >> 3;                                    // widths are multiples of 8 (4 for chroma of odd multiples: handled by the tail loop)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < groups * P.h; i += gridDim.x * blockDim.x)
  {
    const int y = i / groups, x0 = (i - y * groups) * 8;
    const uint4 q = *reinterpret_cast<const uint4*>(P.p + (size_t)y * P.pitch + x0);
    const uint32_t wv[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
      const uint32_t v = (wv[k >> 1] >> (16 * (k & 1))) & 0xffff, x = x0 + k;
      const uint32_t m = ((x & 0xff) ^ (y & 0xff) ^ (x >> 8) ^ (y >> 8)) & 0xff;           // xor_mask is a uint8_t (:150,156)
      sum += (v & 0xff) ^ m;
      if (P.bd > 8) sum += (v >> 8) ^ m;
    }
  }
  const int tail = P.w & 7;
  if (tail)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tail * P.h; i += gridDim.x * blockDim.x)
    {
      const int y = i / tail, x = (P.w & ~7) + (i - y * tail);
      const uint32_t v = (uint16_t)P.p[(size_t)y * P.pitch + x];
      const uint32_t m = ((x & 0xff) ^ (y & 0xff) ^ (x >> 8) ^ (y >> 8)) & 0xff;
      sum += (v & 0xff) ^ m;
      if (P.bd > 8) sum += (v >> 8) ^ m;
    }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(P.out, sum);
}

// ---- CRC --------------------------------------------------------------------------------------------------------------------
// polynomials over GF(2) modulo x^16 + 0x1021
__device__ __forceinline__ uint32_t crcMulMod(uint32_t a, uint32_t b)
{
  uint32_t r = 0;
#pragma unroll
  for (int i = 15; i >= 0; i--)
  {
    r = ((r << 1) & 0xffff) ^ ((r >> 15) & 1 ? 0x1021u : 0u);
    if ((b >> i) & 1) r ^= a;
  }
  return r;
}

__device__ __forceinline__ uint32_t crcXPow(uint64_t e)           // x^e mod P
{
  uint32_t r = 1, base = 2;
  while (e)
  {
    if (e & 1) r = crcMulMod(r, base);
    base = crcMulMod(base, base);
    e >>= 1;
  }
  return r;
}

constexpr int CRC_RUN = 64;                                       // samples per thread

__global__ void __launch_bounds__(128) k_hash_crc(const HashPlane* __restrict__ planes, int nPlanes)
{
  const HashPlane P = planes[blockIdx.y];
  const int runsPerRow = (P.w + CRC_RUN - 1) / CRC_RUN, bitsPerSample = P.bd > 8 ? 16 : 8;
  const uint64_t totalBits = (uint64_t)P.w * P.h * bitsPerSample;
  uint32_t acc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < runsPerRow * P.h; i += gridDim.x * blockDim.x)
  {
    const int y = i / runsPerRow, x0 = (i - y * runsPerRow) * CRC_RUN, n = min(CRC_RUN, P.w - x0);
    const pel* s = P.p + (size_t)y * P.pitch + x0;
    uint32_t crc = 0;                                             // remainder of this run's bits alone
    for (int k = 0; k < n; k++)
    {
      const uint32_t v = (uint16_t)s[k];
      const uint32_t bits = P.bd > 8 ? ((v & 0xff) << 8 | v >> 8) : (v & 0xff);       // low byte first, then the high byte (:103-118)
      for (int b = bitsPerSample - 1; b >= 0; b--)
      {
        const uint32_t msb = (crc >> 15) & 1;
        crc = (((crc << 1) | ((bits >> b) & 1)) & 0xffff) ^ (msb ? 0x1021u : 0u);
      }
    }
    // bits that follow this run in the picture, plus the 16 flush bits (:121-125)
    const uint64_t after = totalBits - ((uint64_t)y * P.w + x0 + n) * bitsPerSample + 16;
    acc ^= crcMulMod(crc, crcXPow(after));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) acc ^= crcMulMod(0xffffu, crcXPow(totalBits + 16));      // the initial value travels through all bits
  for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicXor(P.out, acc);
}

// ---- MD5 --------------------------------------------------------------------------------------------------------------------
__constant__ uint32_t c_md5K[64] = {
  0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193,
  0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed,
  0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa,
  0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
  0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391 };

__device__ __forceinline__ void md5Block(uint32_t st[4], const uint32_t* __restrict__ X)
{
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
#define MD5_STEP(F, G, S, I)                                                          \
  {                                                                                   \
    const uint32_t t = a + (F) + c_md5K[I] + X[G];                                    \
    a = d; d = c; c = b; b = b + __funnelshift_l(t, t, S);                            \
  }
#pragma unroll
  for (int i = 0; i < 16; i++) MD5_STEP((b & c) | (~b & d), i, ((i & 3) == 0 ? 7 : (i & 3) == 1 ? 12 : (i & 3) == 2 ? 17 : 22), i)
#pragma unroll
  for (int i = 16; i < 32; i++) MD5_STEP((d & b) | (~d & c), (5 * i + 1) & 15, ((i & 3) == 0 ? 5 : (i & 3) == 1 ? 9 : (i & 3) == 2 ? 14 : 20), i)
#pragma unroll
  for (int i = 32; i < 48; i++) MD5_STEP(b ^ c ^ d, (3 * i + 5) & 15, ((i & 3) == 0 ? 4 : (i & 3) == 1 ? 11 : (i & 3) == 2 ? 16 : 23), i)
#pragma unroll
  for (int i = 48; i < 64; i++) MD5_STEP(c ^ (b | ~d), (7 * i) & 15, ((i & 3) == 0 ? 6 : (i & 3) == 1 ? 10 : (i & 3) == 2 ? 15 : 21), i)
#undef MD5_STEP
  st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}

// one warp per plane: lanes stage 32 blocks of 64 bytes of the component's byte stream, lane 0 runs the chain
__global__ void __launch_bounds__(32) k_hash_md5(const HashPlane* __restrict__ planes, int nPlanes)
{
  __shared__ __align__(16) uint32_t buf[32 * 16];
  const HashPlane P = planes[blockIdx.x];
  const int lane = threadIdx.x, bps = P.bd > 8 ? 2 : 1;
  const uint64_t total = (uint64_t)P.w * P.h * bps;               // bytes of the message
  const uint64_t padded = (total + 9 + 63) / 64 * 64;             // + 0x80 + 64-bit length, whole blocks
  uint32_t st[4] = { 0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u };
  for (uint64_t base = 0; base < padded; base += 32 * 64)
  {
    // lane l fills block l of this batch: bytes [base + 64 l, base + 64 l + 64) of the padded message
    const uint64_t b0 = base + (uint64_t)lane * 64;
    if (b0 < padded)
    {
      uint8_t* dst = reinterpret_cast<uint8_t*>(&buf[lane * 16]);
      if (b0 + 64 <= total && bps == 2 && (P.w & 31) == 0)
      {
        // the common case: 32 whole samples of one row (the row length is a multiple of 32 samples)
        const uint64_t s0 = b0 >> 1;
        const int y = (int)(s0 / P.w), x = (int)(s0 - (uint64_t)y * P.w);
        const uint4* src = reinterpret_cast<const uint4*>(P.p + (size_t)y * P.pitch + x);
#pragma unroll
        for (int k = 0; k < 4; k++) reinterpret_cast<uint4*>(dst)[k] = src[k];
      }
      else
      {
        for (int k = 0; k < 64; k++)
        {
          const uint64_t o = b0 + k;
          uint8_t v = 0;
          if (o < total)
          {
            const uint64_t s = o / bps;
            const int y = (int)(s / P.w), x = (int)(s - (uint64_t)y * P.w);
            const uint32_t sv = (uint16_t)P.p[(size_t)y * P.pitch + x];
            v = (uint8_t)(bps == 2 && (o & 1) ? sv >> 8 : sv);
          }
          else if (o == total) v = 0x80;
          else if (o >= padded - 8) v = (uint8_t)((total * 8) >> (8 * (o - (padded - 8))));
          dst[k] = v;
        }
      }
    }
    __syncwarp();
    if (lane == 0)
    {
      const int nb = (int)min((uint64_t)32, (padded - base) / 64);
      for (int k = 0; k < nb; k++) md5Block(st, &buf[k * 16]);
    }
    __syncwarp();
  }
  if (lane == 0) { P.out[0] = st[0]; P.out[1] = st[1]; P.out[2] = st[2]; P.out[3] = st[3]; }
}

}   // namespace vtmgpu
