// alf_fast.cuh -- the packed fast-path routines of k_alf (sm_100a): Laplacian cells, luma 7x7 and CC-ALF.
//
//   AdaptiveLoopFilter.cpp: deriveClassificationBlk :873-1082, filterBlk :1084-1324, filterBlkCcAlf :1327-1416
//
// LUMA LAYOUT.  TMA delivers the luma tile row-major (two horizontally adjacent samples per 32-bit word).  The 7x7 diamond
// needs every horizontal offset -3..3 of every row, and with that packing each odd offset costs a funnel shift per word
// (round 1: 100 SHF + 68 LDS.64 per 4x4 block, a quarter of the filter's instructions).  The fast path therefore filters
// from a second copy of the tile in VERTICAL pairs: word V[r][x] = (S[r][x], S[r+1][x]) for EVERY row r (even and odd), so
//   * a horizontal offset is an address offset: one LDS.128 delivers four columns, no shifts;
//   * a vertical offset is another row of the copy: V[r + dy] holds (S[r+dy][x], S[r+1+dy][x]) -- no re-pairing either;
//   * one thread filters a 4x4 block as two output row pairs x four columns: 34 LDS.128 and no shift per block.
// The copy is written by the Laplacian phase from the registers it has loaded anyway (16 PRMT + 4 STS.128 per block).
#pragma once

#include "packed16.cuh"
#include "vtmgpu_dev.cuh"

namespace vtmgpu
{

constexpr int AV_COLS = 72;                 // words per row of the vertical-pair copy: luma columns x0-4 .. x0+67
constexpr int AV_ROWS = 70;                 // rows rr = 0..69 <-> sample row pair (y0 - 4 + rr, y0 - 3 + rr)
constexpr int AV_BYTES = AV_ROWS * AV_COLS * 4;

// 7x7 diamond on one 4x4 block from the vertical-pair copy.  v = &V[row pair of the block's first sample row][block column - 4]
// (16-byte aligned).  e = pre-expanded {coefficient, clip} entry of the block's (filter set, class, transpose).
//   clamp(n - cur, -c, c) + c  ==  max(min(n + (c - cur), 2c), 0)  is ONE VIADDMNMX.S16x2.RELU ; the excess sum(coef * 2c)
//   is folded into e->bias together with the rounding offset 64 (filterBlk :1249-1297).
// Tap k of the diamond (transpose already applied by the table): 0 = (0,+3) ; 1,2,3 = (+1,+2) (0,+2) (-1,+2) ;
// 4..8 = (+2,+1) (+1,+1) (0,+1) (-1,+1) (-2,+1) ; 9,10,11 = (+3,0) (+2,0) (+1,0)   [(dx,dy), each with its point mirror].
// Not for the block rows next to the ALF virtual boundary (the caller sends those to the row-clamping routine).
// packed 16x2 add on the ALU pipe: VIADD.16x2 and IMAD.IADD (what ptxas picks for a + b) both execute on the FMA pipe together
// with IDP.2A (tools/microbench/mb_mix.cu); VIADDMNMX with an unreachable bound is the same add on the other pipe
__device__ __forceinline__ uint32_t addAlu(uint32_t a, uint32_t b) { return __viaddmin_s16x2(a, b, 0x7fff7fffu); }

// the operands of one filter entry in registers (nine 128-bit loads + the bias)
struct LumaCoef
{
  uint32_t coefB[12], clipP1[12], clip2[12];
  int bias;
};

__device__ __forceinline__ LumaCoef loadLumaCoef(const AlfLumaEntry* __restrict__ e)
{
  LumaCoef K;
  const uint4* q = reinterpret_cast<const uint4*>(e);
#pragma unroll
  for (int i = 0; i < 3; i++)
  {
    const uint4 a = q[i], b = q[3 + i], c = q[6 + i];            // plain loads: k_alf keeps the tile's filter set in shared memory
    K.coefB[4 * i] = a.x; K.coefB[4 * i + 1] = a.y; K.coefB[4 * i + 2] = a.z; K.coefB[4 * i + 3] = a.w;
    K.clipP1[4 * i] = b.x; K.clipP1[4 * i + 1] = b.y; K.clipP1[4 * i + 2] = b.z; K.clipP1[4 * i + 3] = b.w;
    K.clip2[4 * i] = c.x; K.clip2[4 * i + 1] = c.y; K.clip2[4 * i + 2] = c.z; K.clip2[4 * i + 3] = c.w;
  }
  K.bias = e->bias;
  return K;
}

template <int BAL = 0>
__device__ __forceinline__ void alfLumaBlockV(const uint32_t* __restrict__ v, pel* __restrict__ out, int pitchOut, const LumaCoef& K, uint32_t maxvP)
{
  const uint32_t* coefB = K.coefB; const uint32_t* clipP1 = K.clipP1; const uint32_t* clip2 = K.clip2;
  const int bias = K.bias;
#define AV_LOAD12(W, PTR)                                                                             \
  {                                                                                                   \
    const uint4* p_ = reinterpret_cast<const uint4*>(PTR);                                            \
    const uint4 a_ = p_[0], b_ = p_[1], c_ = p_[2];                                                   \
    W[0] = a_.x; W[1] = a_.y; W[2] = a_.z; W[3] = a_.w; W[4] = b_.x; W[5] = b_.y; W[6] = b_.z; W[7] = b_.w; \
    W[8] = c_.x; W[9] = c_.y; W[10] = c_.z; W[11] = c_.w;                                             \
  }
  // tap K at horizontal offset DX: P = row +d, column +DX ; M = row -d, column -DX ; all four columns of the block
#define AV_TAP(K, DX, P, M)                                                                           \
  _Pragma("unroll") for (int c = 0; c < 4; c++)                                                       \
  {                                                                                                   \
    const uint32_t cb = (BAL & 2) ? addAlu(clipP1[K], ncur[c]) : __vadd2(clipP1[K], ncur[c]);         \
    const uint32_t tp_ = addClamp0(P[4 + c + (DX)], cb, clip2[K]), tm_ = addClamp0(M[4 + c - (DX)], cb, clip2[K]); \
    const uint32_t s = (BAL & 1) ? addAlu(tp_, tm_) : tp_ + tm_;                                      \
    acc0[c] = __dp2a_lo((int)s, (int)coefB[K], acc0[c]);                                              \
    acc1[c] = __dp2a_hi((int)s, (int)coefB[K], acc1[c]);                                              \
  }
#pragma unroll 1
  for (int half = 0; half < 2; half++)                       // output row pairs (y, y+1) and (y+2, y+3)
  {
    const uint32_t* v0 = v + half * 2 * AV_COLS;
    uint32_t cur[4], ncur[4];
    int acc0[4], acc1[4];
    {
      uint32_t C[12];
      AV_LOAD12(C, v0)
#pragma unroll
      for (int c = 0; c < 4; c++) { cur[c] = C[4 + c]; ncur[c] = ~cur[c]; acc0[c] = bias; acc1[c] = bias; }    // clipP1 + ~cur = clip - cur per lane
      AV_TAP(9, 3, C, C) AV_TAP(10, 2, C, C) AV_TAP(11, 1, C, C)
    }
    {
      uint32_t P[12], M[12];
      AV_LOAD12(P, v0 + AV_COLS) AV_LOAD12(M, v0 - AV_COLS)
      AV_TAP(4, 2, P, M) AV_TAP(5, 1, P, M) AV_TAP(6, 0, P, M) AV_TAP(7, -1, P, M) AV_TAP(8, -2, P, M)
    }
    {
      uint32_t P[12], M[12];
      AV_LOAD12(P, v0 + 2 * AV_COLS) AV_LOAD12(M, v0 - 2 * AV_COLS)
      AV_TAP(1, 1, P, M) AV_TAP(2, 0, P, M) AV_TAP(3, -1, P, M)
    }
    {
      uint32_t P[12], M[12];
      { const uint4 a_ = *reinterpret_cast<const uint4*>(v0 + 3 * AV_COLS + 4); P[4] = a_.x; P[5] = a_.y; P[6] = a_.z; P[7] = a_.w; }
      { const uint4 a_ = *reinterpret_cast<const uint4*>(v0 - 3 * AV_COLS + 4); M[4] = a_.x; M[5] = a_.y; M[6] = a_.z; M[7] = a_.w; }
      AV_TAP(0, 0, P, M)
    }
    uint32_t res[4];
#pragma unroll
    for (int c = 0; c < 4; c++) res[c] = addClamp0(cur[c], prmt((uint32_t)(acc0[c] >> 7), (uint32_t)(acc1[c] >> 7), 0x5410u), maxvP);
    pel* o = out + (size_t)(2 * half) * pitchOut;
    *reinterpret_cast<uint2*>(o) = make_uint2(prmt(res[0], res[1], 0x5410u), prmt(res[2], res[3], 0x5410u));
    *reinterpret_cast<uint2*>(o + pitchOut) = make_uint2(prmt(res[0], res[1], 0x7632u), prmt(res[2], res[3], 0x7632u));
  }
#undef AV_TAP
#undef AV_LOAD12
}


// ---- Laplacian cells of one 4x4 block + its part of the vertical-pair copy -------------------------------------------------
// deriveClassificationBlk :917-975 sub-samples the Laplacians: of every 2x2 cell only the positions (r, c) and (r+1, c+1) are
// evaluated.  The two positions of a cell go into the two 16-bit lanes of one register ("diagonal pair"
// D(r,c) = (S[r][c], S[r+1][c+1]), one PRMT from the row-major words), so every neighbour sum, comparison and the final
// lane sum work on both positions at once:
//   |2a - b - c| = max(b + c, 2a) - min(b + c, 2a)       VIADDMNMX.S16x2 twice (add fused with max / min)
//   sum over the two lanes and the subtraction           IDP.2A with byte operands (1,1) and (-1,-1)
// c0 = first sample of the block in the row-major tile (pitch P samples); W[r][j] = word j (columns x-4+2j, +1) of row y-1+r.
// cells = false (block rows next to the ALF virtual boundary: the caller computes their cells with the row-replacing routine)
// only writes the copy.
// gate() is called once after the loads and before the first store: the caller waits there until the previous tile's readers of the
// copy and the cells are done (k_alf: split barrier), so the loads and the arithmetic of this phase overlap the stragglers.
template <int P, int CELLP, class Gate>
__device__ __forceinline__ void alfBlockCellsAndCopy(const pel* __restrict__ c0, uint2 (*cell)[CELLP], uint32_t* __restrict__ vOwn, int bi, int bj, bool cells, Gate gate)
{
  uint32_t W[6][5];                                                    // words 1..4 of each row are used (index 0 is never read)
#pragma unroll
  for (int r = 0; r < 6; r++)
  {
    const pel* row = c0 + (r - 1) * P;
    const uint2 q = *reinterpret_cast<const uint2*>(row);
    W[r][1] = *reinterpret_cast<const uint32_t*>(row - 2); W[r][2] = q.x; W[r][3] = q.y; W[r][4] = *reinterpret_cast<const uint32_t*>(row + 4);
  }
  // vertical pairs of the block's own columns: row pairs (y+k, y+k+1), k = 0..3
  uint4 vp[4];
#pragma unroll
  for (int k = 0; k < 4; k++)
    vp[k] = make_uint4(prmt(W[k + 1][2], W[k + 2][2], 0x5410u), prmt(W[k + 1][2], W[k + 2][2], 0x7632u),
                       prmt(W[k + 1][3], W[k + 2][3], 0x5410u), prmt(W[k + 1][3], W[k + 2][3], 0x7632u));
  if (!cells)
  {
    gate();
#pragma unroll
    for (int k = 0; k < 4; k++) *reinterpret_cast<uint4*>(vOwn + k * AV_COLS) = vp[k];
    return;
  }
  // D[ri][ci] = (S[y-1+ri][x-1+ci], S[y+ri][x+ci]), ri, ci = 0..4
  uint32_t D[5][5];
#pragma unroll
  for (int ri = 0; ri < 5; ri++)
  {
    D[ri][0] = prmt(W[ri][1], W[ri + 1][2], 0x5432u);
    D[ri][1] = prmt(W[ri][2], W[ri + 1][2], 0x7610u);
    D[ri][2] = prmt(W[ri][2], W[ri + 1][3], 0x5432u);
    D[ri][3] = prmt(W[ri][3], W[ri + 1][3], 0x7610u);
    D[ri][4] = prmt(W[ri][3], W[ri + 1][4], 0x5432u);
  }
  uint2 cv[2][2];
#pragma unroll
  for (int cy = 0; cy < 2; cy++)
#pragma unroll
    for (int cx = 0; cx < 2; cx++)
    {
      const int r = 1 + 2 * cy, c = 1 + 2 * cx;
      const uint32_t c2 = __vadd2(D[r][c], D[r][c]);
#define AV_LAP(A, B) __dp2a_lo((int)__viaddmin_s16x2(A, B, c2), (int)0xffffu, __dp2a_lo((int)__viaddmax_s16x2(A, B, c2), (int)0x0101u, 0))
      const int v = AV_LAP(D[r - 1][c], D[r + 1][c]), h = AV_LAP(D[r][c - 1], D[r][c + 1]);
      const int d0 = AV_LAP(D[r - 1][c - 1], D[r + 1][c + 1]), d1 = AV_LAP(D[r + 1][c - 1], D[r - 1][c + 1]);
#undef AV_LAP
      cv[cy][cx] = make_uint2((uint32_t)v | (uint32_t)h << 16, (uint32_t)d0 | (uint32_t)d1 << 16);
    }
  gate();
#pragma unroll
  for (int k = 0; k < 4; k++) *reinterpret_cast<uint4*>(vOwn + k * AV_COLS) = vp[k];
  // the two cells of a cell row are adjacent: one 128-bit store per row (cell index 2 bj + 1 is odd: 8-byte aligned only -> two 64-bit stores)
#pragma unroll
  for (int cy = 0; cy < 2; cy++)
  {
    cell[2 * bi + 1 + cy][2 * bj + 1] = cv[cy][0];
    cell[2 * bi + 1 + cy][2 * bj + 2] = cv[cy][1];
  }
}

// One Laplacian cell anywhere in the tile (ring of halo cells, block rows next to the ALF virtual boundary), same packed scheme.
// p0 = top-left sample of the 2x2 cell (even column), r = its picture row: rows beyond the virtual boundary are replaced
// (deriveClassificationBlk :906-915), i.e. the diagonal pairs are taken from the row pairs (rm, r), (r, r+1), (r+1, r2).
template <int P>
__device__ __forceinline__ uint2 alfCellAny(const pel* __restrict__ p0, int r, int ctuMask, int vbL)
{
  int up = -P, dn2 = 2 * P;                                                   // row r-1, row r+2
  const int rv = r & ctuMask;
  if (r > 0 && rv == vbL - 2) dn2 = P;
  else if (r > 0 && rv == vbL) up = 0;
  uint32_t W[4][3];                                                           // rows rm, r, r+1, r2 ; words at columns -2, 0, +2
  const pel* rows[4] = { p0 + up, p0, p0 + P, p0 + dn2 };
#pragma unroll
  for (int k = 0; k < 4; k++)
  {
    W[k][0] = *reinterpret_cast<const uint32_t*>(rows[k] - 2); W[k][1] = *reinterpret_cast<const uint32_t*>(rows[k]);
    W[k][2] = *reinterpret_cast<const uint32_t*>(rows[k] + 2);
  }
  uint32_t D[3][3];                                                           // D[k][ci] = (S[row k][c-1+ci], S[row k+1][c+ci])
#pragma unroll
  for (int k = 0; k < 3; k++)
  {
    D[k][0] = prmt(W[k][0], W[k + 1][1], 0x5432u);
    D[k][1] = prmt(W[k][1], W[k + 1][1], 0x7610u);
    D[k][2] = prmt(W[k][1], W[k + 1][2], 0x5432u);
  }
  const uint32_t c2 = __vadd2(D[1][1], D[1][1]);
#define AV_LAP(A, B) __dp2a_lo((int)__viaddmin_s16x2(A, B, c2), (int)0xffffu, __dp2a_lo((int)__viaddmax_s16x2(A, B, c2), (int)0x0101u, 0))
  const int v = AV_LAP(D[0][1], D[2][1]), h = AV_LAP(D[1][0], D[1][2]);
  const int d0 = AV_LAP(D[0][0], D[2][2]), d1 = AV_LAP(D[2][0], D[0][2]);
#undef AV_LAP
  return make_uint2((uint32_t)v | (uint32_t)h << 16, (uint32_t)d0 | (uint32_t)d1 << 16);
}

// one item of the border of the vertical-pair copy (the rows above / below and the columns left / right of the blocks' own
// part): four columns of one row pair from the row-major tile.  h = sample (row pair's first row, first column) in the tile.
template <int P>
__device__ __forceinline__ void alfCopyQuad(const pel* __restrict__ h, uint32_t* __restrict__ v)
{
  const uint2 a = *reinterpret_cast<const uint2*>(h), b = *reinterpret_cast<const uint2*>(h + P);
  *reinterpret_cast<uint4*>(v) = make_uint4(prmt(a.x, b.x, 0x5410u), prmt(a.x, b.x, 0x7632u), prmt(a.y, b.y, 0x5410u), prmt(a.y, b.y, 0x7632u));
}

// ---- CC-ALF, 4:2:0 / 4:2:2 (collocated luma column = 2 * chroma column) ------------------------------------------------------
// filterBlkCcAlf :1376-1386 is linear: sum_k f[k] * (L[pos_k] - L[cur]).  The collocated luma column is even, so every word of
// the row-major tile holds (L[2i], L[2i+1]) and IDP.2A applies TWO taps of one chroma sample per instruction -- 6 IDP.2A per
// sample and no lane shuffles (round 1: 14 IDP.2A + 14 PRMT per four samples and a multiply per sample):
//   row l2 (-1): word i   x (f0, 0)        row l3 (+2): word i   x (f6, 0)
//   row 0      : word i-1 x (0, f1)                     word i   x (-sum f, f2)
//   row l1 (+1): word i-1 x (0, f3)                     word i   x (f4, f5)
// k[0] = bytes (f0, 0, f6, 0), k[1] = (0, f1, 0, f3), k[2] = (-sum f, f2, f4, f5)  (expanded on the host, AlfDev::ccK; filters whose
// coefficient sum does not fit a signed byte take the general routine).  l = luma sample collocated with the first of the
// four chroma samples; l1, l2, l3 = row offsets after the virtual-boundary rule.  Returns the corrections, clipped.
__device__ __forceinline__ uint2 ccAlfQuadDual(const pel* __restrict__ l, int l1, int l2, int l3, uint32_t kA, uint32_t kB, uint32_t kC, uint32_t maxcP, uint32_t halfP)
{
  const uint4 r0 = *reinterpret_cast<const uint4*>(l), r1 = *reinterpret_cast<const uint4*>(l + l1);
  const uint4 r2 = *reinterpret_cast<const uint4*>(l + l2), r3 = *reinterpret_cast<const uint4*>(l + l3);
  const uint32_t m0 = *reinterpret_cast<const uint32_t*>(l - 2), m1 = *reinterpret_cast<const uint32_t*>(l + l1 - 2);
  const uint32_t w0[5] = { m0, r0.x, r0.y, r0.z, r0.w }, w1[5] = { m1, r1.x, r1.y, r1.z, r1.w };
  const uint32_t w2[4] = { r2.x, r2.y, r2.z, r2.w }, w3[4] = { r3.x, r3.y, r3.z, r3.w };
  int a[4];
#pragma unroll
  for (int i = 0; i < 4; i++)
  {
    int s = __dp2a_lo((int)w2[i], (int)kA, 64);
    s = __dp2a_hi((int)w3[i], (int)kA, s);
    s = __dp2a_lo((int)w0[i], (int)kB, s);
    s = __dp2a_hi((int)w1[i], (int)kB, s);
    s = __dp2a_lo((int)w0[i + 1], (int)kC, s);
    s = __dp2a_hi((int)w1[i + 1], (int)kC, s);
    a[i] = s >> 7;
  }
  uint2 res;
  // ClipPel(sum + half) - half
  res.x = __vadd2(addClamp0(prmt((uint32_t)a[0], (uint32_t)a[1], 0x5410u), halfP, maxcP), ~halfP + 0x00010001u);
  res.y = __vadd2(addClamp0(prmt((uint32_t)a[2], (uint32_t)a[3], 0x5410u), halfP, maxcP), ~halfP + 0x00010001u);
  return res;
}

}   // namespace vtmgpu
