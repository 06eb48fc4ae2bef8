// drfe ORB front end for sm_100a: ComputePyramid, per-cell FAST-9 + quadtree distribution,
// IC_Angle, 7x7 Gaussian and steered rBRIEF — the work of
// Planar_SLAM::ORBextractor::operator() (reference src/ORBextractor.cc:1043-1105),
// batched over independent frames.  All arithmetic is integer / fixed point except
// fastAtan2 and the descriptor steering, which use explicit round-to-nearest float ops
// (no FMA contraction) so results match the CPU oracle bit for bit.
//
// Kernels (one launch each per batch unless noted):
//   k_pyr_level0   gray -> level 0 + 19 px reflect-101 frame   (ORBextractor.cc:1125-1129)
//   k_pyr_resize   level l-1 -> level l, OpenCV fixed-point INTER_LINEAR + frame; one
//                  launch per level because level l reads level l-1   (:1118-1123)
//   k_fast_strips  one CTA per row of 30 px grid cells: rows staged in smem by TMA bulk copies,
//                  FAST-9/16 score for every pixel with packed u16x2 min/max (VIMNMX3), per-cell
//                  NMS, iniThFAST -> minThFAST fallback                      (:789-829)
//   k_quadtree     one CTA per (frame, level): DistributeOctTree replayed with prefix scans,
//                  order-free (node membership is geometric; ties resolved by the
//                  reference's candidate order encoded in a key)           (:539-763)
//   k_blur         separable 8.8 fixed-point Gaussian 7x7, sigma 2          (:1085-1086)
//   k_orient_describe  one warp per keypoint: IC_Angle (:77-104) + rBRIEF-256 (:107-147)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)

#include "drfe_internal.h"

namespace drfe {

static const int kEdge = 19;     // EDGE_THRESHOLD (ORBextractor.cc:74)
static const int kXOff = 32;     // byte offset of ROI column 0 inside a bordered row
static const int kHalfPatch = 15;

__constant__ int c_umax[16];
static const int8_t h_pattern[1024] = {
#include "../../include/drfe_orb_pattern.inc"
};

struct LevelDev {
  int w, h, pitch, rows;            // bordered buffer: rows x pitch bytes, ROI at (kEdge, kXOff)
  long long img_off, img_fstride;   // bytes
  int bpitch;
  long long blur_off, blur_fstride;
  int regW, regH, nCols, nRows, wCell, hCell;
  int nfeat, nIni;
  int blur_nq;                      // 4-pixel columns of the blur kernel, ceil(w/4)
  uint32_t blur_magic;              // ceil(2^32 / blur_nq)
  uint32_t wcell_magic;             // ceil(2^32 / wCell): interior x -> cell column by __umulhi
  float hX;
  int root_x[5];                    // root boundaries int(hX*i), i = 0..nIni (nIni <= 4)
  int cand_cap;
  long long cand_off;               // element offset inside one frame's candidate arena
  int node_cap;
  int kp_off;                       // offset inside one frame's per-level keypoint arena
  float scale, size;
  long long xtab_off, ytab_off;     // element offsets into the resize tables (level >= 1)
  int ptile_off, ptile_cnt, ptile_gx;   // this level's k_pyr_resize tiles inside OrbDev::ptiles
  int pcol_off, pcol_groups, ytab2_off; // this level's k_pyr_stream column records / row table inside OrbDev::pcols / ytab2
};

struct PyrTile;
// FAST strip: interior rows [y0, y0+nrows) of one cell row, interior columns [x0, x0+xw).  Levels wider than
// kFastSplitWidth are cut into segments of 16 cell columns (x0 is then a multiple of both the cell width and 16, which
// keeps the TMA source and the 16-pixel compass groups 16-byte aligned), so that a CTA's shared memory stays near 70 KB
// (3 CTAs per SM) whatever the image width; per-cell NMS makes segments independent.
struct StripDev { short level, y0, nrows, x0, xw, pad; uint32_t groups_magic; };   // groups_magic = ceil(2^32 / groups)
static const int kFastSplitWidth = 704, kFastSegCells = 16;
static const int kMaxStripCells = 160;

struct OrbDev {
  int nlevels, B, ini_th, min_th;
  LevelDev lv[DRFE_MAX_LEVELS];
  uint8_t* pyr;
  uint8_t* blur;
  const uint2* rtab;                // resize tables {idx0 | idx1<<16, c0 | c1<<16}
  const StripDev* strips;
  const struct PyrTile* ptiles;     // k_pyr_resize tiles of all levels
  const struct PyrCol* pcols;       // k_pyr_stream per-column-group records of all levels (null: generic kernel only)
  const uint2* ytab2;               // k_pyr_stream row table {r0 | b0 << 16, b1} per destination row
  int pyr_src_bytes;                // smem bytes reserved for a tile's source window
  uint32_t* cand;                   // [B][cand_total] packed x | y<<12 | q<<24 (region coords)
  long long cand_fstride;
  int* cand_cnt;                    // [B][nlevels]
  uint16_t* node_of;                // [B][cand_total] scratch for the quadtree
  uint32_t* lkp;                    // [B][lkp_total] packed x | y<<12 | q<<24 (level coords)
  int lkp_fstride;
  int* lkp_cnt;                     // [B][nlevels]
  drfe_keypoint* out_kp;            // [B][kp_cap]
  uint8_t* out_desc;                // [B][kp_cap][32]
  int* out_cnt;                     // [B]
  int kp_cap;
  int* status;                      // bit 0: candidate overflow, bit 1: node overflow
  int fast_tp, fast_rows, fast_list_off, fast_list_cap;
  int fast_tma2d;                   // the FAST tile comes by one tensor-map TMA load per strip
  int blur_blk_off[DRFE_MAX_LEVELS + 1];   // first k_blur block of each level
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
__device__ __forceinline__ uint8_t* roi_ptr(const OrbDev& P, const LevelDev& L, int f) {
  return P.pyr + L.img_off + (long long)f * L.img_fstride + (long long)kEdge * L.pitch + kXOff;
}

// ------------------------------------------------------------------ K1 pyramid
// Each thread writes 16 horizontally adjacent bytes (one aligned uint4) of the bordered level
// image.  Threads are numbered so that whole warps do the same kind of work: first every
// (row, 16-byte chunk inside the image) pair — one 16-byte load, one 16-byte store — then the
// chunks that touch the 19 px frame, which gather their bytes from the reflected interior
// coordinate (identical bytes to copyMakeBorder, no second pass).  ci: chunks inside the image per
// row (0 when the source is not 16-byte aligned), cb: frame chunks per row, *_magic = ceil(2^32 / n).
__global__ void __launch_bounds__(256) k_pyr_level0(const OrbDev* __restrict__ Pp,
                                                     const uint8_t* __restrict__ src,
                                                     long long row_stride, long long frame_stride, int f0,
                                                     int ci, uint32_t ci_magic, int cb, uint32_t cb_magic) {
  DRFE_GRID_DEP();
  const OrbDev& P = *Pp;
  const LevelDev& L = P.lv[0];
  const int f = blockIdx.y + f0;
  const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
  const uint32_t nA = (uint32_t)L.rows * (uint32_t)ci;
  uint8_t* dbase = P.pyr + L.img_off + (long long)f * L.img_fstride;
  const uint8_t* sbase = src + (long long)f * frame_stride;
  if (idx < nA) {
    const int by = (int)__umulhi(idx, ci_magic), c = (int)idx - by * ci;
    const uint4 o = __ldg(reinterpret_cast<const uint4*>(sbase + (long long)reflect101(by - kEdge, L.h) * row_stride) + c);
    *reinterpret_cast<uint4*>(dbase + (long long)by * L.pitch + kXOff + 16 * c) = o;
    return;
  }
  const uint32_t j = idx - nA;
  if (j >= (uint32_t)L.rows * (uint32_t)cb) return;
  const int by = (int)__umulhi(j, cb_magic), q = (int)j - by * cb;
  const int c = q < 2 ? q : q + ci;                              // chunk index from the row start (ROI starts at chunk 2)
  const uint8_t* s = sbase + (long long)reflect101(by - kEdge, L.h) * row_stride;
  const int x0 = 16 * c - kXOff;                                 // image column of the chunk's first byte
  uint32_t w[4];
#pragma unroll
  for (int qq = 0; qq < 4; ++qq) {
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int x = x0 + 4 * qq + k;
      x = min(max(x, -(L.w - 1)), 2 * (L.w - 1));                // bytes outside the frame are never read back; keep the index valid
      v |= (uint32_t)__ldg(s + reflect101(x, L.w)) << (8 * k);
    }
    w[qq] = v;
  }
  *reinterpret_cast<uint4*>(dbase + (long long)by * L.pitch + 16 * c) = make_uint4(w[0], w[1], w[2], w[3]);
}

// cv::resize INTER_LINEAR 8UC1: 11-bit coefficient fixed point.
// One CTA = a 128 x 32 tile of the bordered destination level (border pixels are computed
// like any other pixel, from their reflected coordinate).  The source window of the tile is
// staged in shared memory with coalesced word loads; the horizontal pass runs once per
// source row (thread = 4 fixed destination columns, its 4 table entries live in registers)
// and leaves T >> 4 as u16 in shared memory; the vertical pass combines two such rows per
// destination row.  One launch per level: level l reads level l-1.
static const int kPyrTW = 128, kPyrTH = 32;
struct PyrTile { short g0, by0, sx0, sw4, sy0, nsy, pad0, pad1; };   // first 4-px group, first bordered row, source window

__global__ void __launch_bounds__(256) k_pyr_resize(const OrbDev* __restrict__ Pp, int level, int f0) {
  DRFE_GRID_DEP();
  extern __shared__ __align__(128) uint8_t smem[];
  const OrbDev& P = *Pp;
  const LevelDev& L = P.lv[level];
  const LevelDev& S = P.lv[level - 1];
  const PyrTile t = P.ptiles[L.ptile_off + blockIdx.x];
  const int f = blockIdx.y + f0;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int SW = t.sw4 * 4;                                   // source window pitch (bytes)
  uint8_t* s_src = smem;                                      // [nsy][SW]
  uint2* s_T = reinterpret_cast<uint2*>(smem + P.pyr_src_bytes);   // [nsy][32] 4 x u16 per thread column
  // ---- source window -> smem
  {
    const uint8_t* src = roi_ptr(P, S, f) + (long long)t.sy0 * S.pitch + t.sx0;   // sx0 is a multiple of 4
    const int nwords = t.nsy * t.sw4;
    uint32_t* d = reinterpret_cast<uint32_t*>(s_src);
    for (int i = threadIdx.x; i < nwords; i += 256) {
      const int r = i / t.sw4, c = i - r * t.sw4;
      d[i] = __ldg(reinterpret_cast<const uint32_t*>(src + (long long)r * S.pitch) + c);
    }
  }
  // ---- this thread's 4 destination columns
  const int groups = (L.w + 40 + 3) >> 2;
  const int g = t.g0 + tx;
  const bool col_ok = g < groups;
  int i0[4], i1[4], c0[4], c1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint2 e = __ldg(P.rtab + L.xtab_off + reflect101(4 * g - 20 + k, L.w));
    i0[k] = (int)(e.x & 0xFFFF) - t.sx0; i1[k] = (int)(e.x >> 16) - t.sx0;
    c0[k] = (int)(e.y & 0xFFFF); c1[k] = (int)(e.y >> 16);
    if (!col_ok) { i0[k] = 0; i1[k] = 0; }
  }
  __syncthreads();
  // ---- horizontal pass: one source row at a time
  for (int r = ty; r < t.nsy; r += 8) {
    const uint8_t* row = s_src + r * SW;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (uint32_t)(row[i0[k]] * c0[k] + row[i1[k]] * c1[k]) >> 4;
    s_T[r * 32 + tx] = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
  }
  __syncthreads();
  // ---- vertical pass
  if (!col_ok) return;
  uint8_t* dbase = P.pyr + L.img_off + (long long)f * L.img_fstride + (kXOff - 20) + 4 * g;
#pragma unroll
  for (int rr = 0; rr < kPyrTH / 8; ++rr) {
    const int by = t.by0 + ty + 8 * rr;
    if (by >= L.rows) break;
    const uint2 e = __ldg(P.rtab + L.ytab_off + reflect101(by - kEdge, L.h));
    const int r0 = (int)(e.x & 0xFFFF) - t.sy0, r1 = (int)(e.x >> 16) - t.sy0;
    const int b0 = (int)(e.y & 0xFFFF), b1 = (int)(e.y >> 16);
    const uint2 T0 = s_T[r0 * 32 + tx], T1 = s_T[r1 * 32 + tx];
    const uint32_t o0 = (((b0 * (int)(T0.x & 0xFFFF)) >> 16) + ((b1 * (int)(T1.x & 0xFFFF)) >> 16) + 2) >> 2;
    const uint32_t o1 = (((b0 * (int)(T0.x >> 16)) >> 16) + ((b1 * (int)(T1.x >> 16)) >> 16) + 2) >> 2;
    const uint32_t o2 = (((b0 * (int)(T0.y & 0xFFFF)) >> 16) + ((b1 * (int)(T1.y & 0xFFFF)) >> 16) + 2) >> 2;
    const uint32_t o3 = (((b0 * (int)(T0.y >> 16)) >> 16) + ((b1 * (int)(T1.y >> 16)) >> 16) + 2) >> 2;
    *reinterpret_cast<uint32_t*>(dbase + (long long)by * L.pitch) = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24);
  }
}


// ---- streaming resize (the fast path; k_pyr_resize above stays as the generic fallback)
// One thread = 4 adjacent columns of the bordered level x kPyrR consecutive interior rows.  Per
// source row the thread loads the 12 aligned bytes that hold its <= 6 source pixels straight from
// global memory (neighbouring threads share the lines in L1), realigns them with two funnel
// shifts and gets each horizontal sum s0*c0 + s1*c1 from ONE funnel shift + ONE 2-way dot product
// (IDP.2A: 16-bit coefficients x 8-bit pixels).  Source rows are consumed in order — r0(y) is
// monotone — so each is filtered once and only two rows of sums live in registers; the vertical
// pass is two high-multiplies per pixel: ((T >> 4) * b) >> 16 == umulhi(T & ~15, b << 12).
// Rows 1..19 and h-20..h-2 are also stored to their BORDER_REFLECT_101 mirror rows; mirrored
// columns come for free from the per-column records (a border column is just another column
// whose source index is the reflected one).
static const int kPyrR = 16;
struct PyrCol { uint32_t base, sh, coef[4], pad0, pad1; };   // window base (byte offset in the source ROI row), 8*d_k per byte, c0 | c1 << 16

struct PyrRow { uint32_t w0, w1, w2; };
__device__ __forceinline__ PyrRow pyr_load(const uint8_t* __restrict__ p) {
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
  PyrRow r;
  r.w0 = __ldg(q); r.w1 = __ldg(q + 1); r.w2 = __ldg(q + 2);
  return r;
}
__device__ __forceinline__ void pyr_hfilter(const PyrRow& r, int al, uint32_t sh, const uint32_t (&coef)[4], uint32_t (&T)[4]) {
  const uint32_t lo = __funnelshift_r(r.w0, r.w1, al), hi = __funnelshift_r(r.w1, r.w2, al);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t v = __funnelshift_rc(lo, hi, (sh >> (8 * k)) & 0xFFu);      // bytes: s0, s1, ...
    T[k] = __dp2a_lo(coef[k], v, 0u) & ~15u;
  }
}

__global__ void __launch_bounds__(128, 10) k_pyr_stream(const OrbDev* __restrict__ Pp, int level, int f0, int R) {
  DRFE_GRID_DEP();
  const OrbDev& P = *Pp;
  const LevelDev& L = P.lv[level];
  const LevelDev& S = P.lv[level - 1];
  const int g = blockIdx.x * 32 + threadIdx.x;
  const int y0 = (blockIdx.y * 4 + threadIdx.y) * R;
  const int f = blockIdx.z + f0;
  if (g >= L.pcol_groups || y0 >= L.h) return;
  const int y1 = min(y0 + R, L.h), h = L.h;
  const PyrCol pc = P.pcols[L.pcol_off + g];
  const int al = 8 * (int)(pc.base & 3u);
  const long long sp = S.pitch, dp = L.pitch;
  const uint2* __restrict__ yt = P.ytab2 + L.ytab2_off;
  int a = (int)(yt[y0].x & 0xFFFFu);
  const uint8_t* pn = roi_ptr(P, S, f) + (pc.base & ~3u) + (long long)a * sp;   // next source row to fetch
  uint8_t* dcol = P.pyr + L.img_off + (long long)f * L.img_fstride + (kXOff - 20) + 4 * g;
  uint8_t* dptr = dcol + (long long)(y0 + kEdge) * dp;
  uint32_t Ta[4], Tb[4];
  {
    const PyrRow ra = pyr_load(pn), rb = pyr_load(pn + sp);
    pn += 2 * sp;
    pyr_hfilter(ra, al, pc.sh, pc.coef, Ta);
    pyr_hfilter(rb, al, pc.sh, pc.coef, Tb);
  }
  PyrRow nx = pyr_load(pn), nx2 = pyr_load(pn + sp);   // rows a + 2 and a + 3, fetched two steps ahead of their use
  pn += sp;
#pragma unroll 1
  for (int y = y0; y < y1; ++y) {
    const uint2 e = yt[y];                       // {r0 | b0 << 16, b1}
    int adv = (int)(e.x & 0xFFFFu) - a;          // 0, 1 or 2 source rows to move on; warp-uniform (a warp shares y)
    a += adv;
#pragma unroll 1
    while (adv > 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) Ta[k] = Tb[k];
      pyr_hfilter(nx, al, pc.sh, pc.coef, Tb);
      nx = nx2;
      pn += sp;
      nx2 = pyr_load(pn);
      --adv;
    }
    const uint32_t B0 = (e.x >> 16) << 12, B1 = e.y << 12;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = (__umulhi(Ta[k], B0) + __umulhi(Tb[k], B1) + 2u) >> 2;
    const uint32_t word = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
    *reinterpret_cast<uint32_t*>(dptr) = word;
    dptr += dp;
    if (y <= kEdge) { if (y >= 1) *reinterpret_cast<uint32_t*>(dcol + (long long)(kEdge - y) * dp) = word; }
    if (y >= h - 1 - kEdge) { if (y <= h - 2) *reinterpret_cast<uint32_t*>(dcol + (long long)(kEdge + 2 * (h - 1) - y) * dp) = word; }
  }
}

// ------------------------------------------------------------------ K2 FAST
// Threshold-free FAST-9/16 score.  With ring values r[0..15] around centre v,
//   A = min over the 16 nine-pixel arcs of max(r over the arc)
//   B = max over the 16 nine-pixel arcs of min(r over the arc)
//   q = max(v - A, B - v) - 1          (= OpenCV's cornerScore; corner at threshold t <=> q >= t)
// Two pixels are processed per 32-bit register as u16 lanes (VIMNMX3.U16x2 on sm_100a) whose HIGH byte is
// the pixel and whose low byte is arbitrary (the neighbouring image byte): min/max of such lanes carry the
// right high byte, so the ring needs no byte extraction at all — every packed min/max, PRMT, SHF and LOP3
// issues on the half-rate ALU pipe (tools/ubench/pipes.cu), which is what bounds this kernel.
// Arcs k and k+1 (k even) share the 8 pixels k+1..k+8:  max(min arc k, min arc k+1) =
// min(r[k+1..k+8], max(r[k], r[k+9])); with pair minima p[j] = min(r[j], r[j+1]) (j odd) and
// pp[j] = min(p[j], p[j+2]) that is min3(pp[k+1], pp[k+5], max(r[k], r[k+9])): 36 operations per polarity.
// Returns e = max(q + 1 - minTh, 0) in the high byte of each lane (low bytes 0); kmin8 = minTh << 8 per lane.
__device__ __forceinline__ uint32_t fast_e_lanes(const uint32_t (&r)[16], uint32_t v, uint32_t kmin8) {
  uint32_t pmn[8], pmx[8];                                    // pairs (1,2) (3,4) ... (15,0)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    pmn[j] = __vminu2(r[2 * j + 1], r[(2 * j + 2) & 15]);
    pmx[j] = __vmaxu2(r[2 * j + 1], r[(2 * j + 2) & 15]);
  }
  uint32_t qmn[8], qmx[8];                                    // 4 consecutive pixels 2j+1 .. 2j+4
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    qmn[j] = __vminu2(pmn[j], pmn[(j + 1) & 7]);
    qmx[j] = __vmaxu2(pmx[j], pmx[(j + 1) & 7]);
  }
  uint32_t bv[8], av[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {                               // k = 2i: pixels k+1..k+8 = q[i], q[i+2]
    const uint32_t hi = __vmaxu2(r[2 * i], r[(2 * i + 9) & 15]), lo = __vminu2(r[2 * i], r[(2 * i + 9) & 15]);
    bv[i] = __vimin3_u16x2(qmn[i], qmn[(i + 2) & 7], hi);
    av[i] = __vimax3_u16x2(qmx[i], qmx[(i + 2) & 7], lo);
  }
  uint32_t B = __vimax3_u16x2(bv[0], bv[1], bv[2]), A = __vimin3_u16x2(av[0], av[1], av[2]);
  B = __vimax3_u16x2(B, bv[3], bv[4]); A = __vimin3_u16x2(A, av[3], av[4]);
  B = __vimax3_u16x2(B, bv[5], bv[6]); A = __vimin3_u16x2(A, av[5], av[6]);
  B = __vmaxu2(B, bv[7]); A = __vminu2(A, av[7]);
  const uint32_t vh = v & 0xFF00FF00u, Ah = A & 0xFF00FF00u, Bh = B & 0xFF00FF00u;
  const uint32_t d1 = vh - __vminu2(vh, Ah), d2 = Bh - __vminu2(Bh, vh);     // max(v - A, 0), max(B - v, 0), << 8
  const uint32_t E = __vmaxu2(d1, d2);                                         // (q + 1) << 8 when positive
  return E - __vminu2(E, kmin8);
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// one 3-D tensor-map TMA load (cp.async.bulk.tensor, SASS UTMALDG): box of the level image {columns as 32-bit words, rows, 1 frame}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
               : "memory");
}
// the per-level tensor maps of k_fast_strips: level image as {pitch / 4 words, rows, frames}, box {fast_tp / 4, hCell + 6, 1}
struct FastMaps { CUtensorMap m[DRFE_MAX_LEVELS]; };

// Full threshold-free score of the 4 pixels of quad g in interior row ry: returns the packed bytes
// e = max(q + 1 - minTh, 0) of pixels 0..3.  tile32: strip rows as words, TP4 words per row.
__device__ __forceinline__ uint32_t fast_eval_quad(const uint32_t* __restrict__ tile32, int TP4, int ry, int g, uint32_t kmin8) {
  // window rows ry..ry+6 (level rows y-3..y+3), bytes 16+4g .. 16+4g+11 (level columns x-3..x+8)
  uint32_t w[7][3];
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t* p = tile32 + (ry + r) * TP4 + 4 + g;
    w[r][0] = p[0]; w[r][1] = p[1]; w[r][2] = p[2];
  }
  // 4 adjacent bytes starting at byte offset o (0..8) of row r
  auto quad = [&](int r, int o) -> uint32_t {
    return (o & 3) == 0 ? w[r][o >> 2] : __funnelshift_r(w[r][o >> 2], w[r][(o >> 2) + 1], 8 * (o & 3));
  };
  // Lanes with the pixel in the high byte: the word of 4 adjacent bytes is that for pixels 1,3 as it stands
  // and for pixels 0,2 after a shift by one byte (an IMAD.SHL on the FMA pipe, not the ALU pipe).
  const uint32_t vq = quad(3, 3);                          // centres of the 4 pixels
  uint32_t re[16], ro[16];
#define DRFE_RING(k, dx, dy)                                   \
  {                                                            \
    ro[k] = quad(3 + (dy), 3 + (dx));                          \
    re[k] = ro[k] * 256u;                                      \
  }
  DRFE_RING(0, 0, 3) DRFE_RING(1, 1, 3) DRFE_RING(2, 2, 2) DRFE_RING(3, 3, 1)
  DRFE_RING(4, 3, 0) DRFE_RING(5, 3, -1) DRFE_RING(6, 2, -2) DRFE_RING(7, 1, -3)
  DRFE_RING(8, 0, -3) DRFE_RING(9, -1, -3) DRFE_RING(10, -2, -2) DRFE_RING(11, -3, -1)
  DRFE_RING(12, -3, 0) DRFE_RING(13, -3, 1) DRFE_RING(14, -2, 2) DRFE_RING(15, -1, 3)
#undef DRFE_RING
  const uint32_t ee = fast_e_lanes(re, vq * 256u, kmin8);  // pixels 0,2 in the high bytes
  const uint32_t eo = fast_e_lanes(ro, vq, kmin8);         // pixels 1,3
  return (ee >> 8) | eo;                                   // bytes: pixel 0,1,2,3
}

// Necessary condition for "corner at threshold t" on the 4 pixels of a quad: every 9-pixel arc of the
// ring contains two adjacent compass points (ring 0,4,8,12), and the adjacent compass pairs are exactly
// {r0,r8} x {r4,r12}, so a corner needs (|r0-v| > t or |r8-v| > t) and (|r4-v| > t or |r12-v| > t).
// (Dropping the "same sign" part of the condition lets 4 pixels go through one VABSDIFF4 per compass
// point; on textured frames it passes 41.6 % of the quads instead of 39.5 %.)
// One call tests the 4 quads 4G..4G+3 of interior row ry (16 pixels; their window starts at a 16-byte
// boundary of the strip row): bit k of the result = quad 4G+k may hold a corner.
// cadd = (127 - min(t,127)) * 0x01010101: a > t  <=>  bit 7 of (a | ((a & 0x7f) + 127 - t)).
__device__ __forceinline__ uint32_t fast_compass_group(const uint32_t* __restrict__ tile32, int TP4, int ry, int G, uint32_t cadd) {
  const uint32_t* pu = tile32 + ry * TP4 + 4 + 4 * G;        // level row y-3: bytes x-3.. of quad 4G
  const uint32_t* pm = pu + 3 * TP4;                         // row y
  const uint32_t* pd = pu + 6 * TP4;                         // row y+3
  const uint4 ua = *reinterpret_cast<const uint4*>(pu), ma = *reinterpret_cast<const uint4*>(pm), da = *reinterpret_cast<const uint4*>(pd);
  const uint2 mb = *reinterpret_cast<const uint2*>(pm + 4);
  const uint32_t U[5] = {ua.x, ua.y, ua.z, ua.w, pu[4]}, D[5] = {da.x, da.y, da.z, da.w, pd[4]};
  const uint32_t M[6] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y};
  uint32_t flags = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t v = __funnelshift_r(M[k], M[k + 1], 24);                              // centres: byte offset 3
    const uint32_t a8 = __vabsdiffu4(__funnelshift_r(U[k], U[k + 1], 24), v);            // (0,-3)
    const uint32_t a0 = __vabsdiffu4(__funnelshift_r(D[k], D[k + 1], 24), v);            // (0,+3)
    const uint32_t a4 = __vabsdiffu4(__funnelshift_r(M[k + 1], M[k + 2], 16), v);        // (+3,0): offset 6
    const uint32_t a12 = __vabsdiffu4(M[k], v);                                           // (-3,0): offset 0
    const uint32_t h8 = (a8 & 0x7F7F7F7Fu) + cadd, h0 = (a0 & 0x7F7F7F7Fu) + cadd;
    const uint32_t h4 = (a4 & 0x7F7F7F7Fu) + cadd, h12 = (a12 & 0x7F7F7F7Fu) + cadd;
    const uint32_t vert = a8 | h8 | a0 | h0, horz = a4 | h4 | a12 | h12;
    flags |= ((vert & horz & 0x80808080u) ? 1u : 0u) << k;
  }
  return flags;
}

// ring offsets (dx,dy), k = 0..15: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)
// (-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)  
//
// One CTA per (cell row of one level, frame).  The cells' interior bands tile the region
// [19, w-19) x [19, h-19) exactly (App. A.3b), so a strip = the interior rows of one cell row
// over the full width; pixel -> cell is (x-19)/wCell.  Stages:
//   1. rows Y0-3 .. Y1+2 of the level image -> smem by TMA bulk copies (one per row)
//   2. score e = max(q + 1 - minTh, 0) for every interior pixel, 4 pixels per task
//   3. per-cell NMS (neighbours in another cell count as 0) -> survivor list in smem, and a
//      per-cell count of survivors with q >= iniTh
//   4. emit survivors with q >= iniTh, or all of them for cells where FAST(iniTh) found
//      nothing (the minThFAST fallback, ORBextractor.cc:812-816), in arbitrary order (the
//      quadtree orders by a key)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_fast_strips(const OrbDev* __restrict__ Pp, int f0, const __grid_constant__ FastMaps maps, int strip0) {
  DRFE_GRID_DEP();
  extern __shared__ __align__(128) uint8_t smem[];
  const OrbDev& P = *Pp;
  const StripDev strip = P.strips[blockIdx.x + strip0];
  const int f = blockIdx.y + f0;
  const LevelDev& L = P.lv[strip.level];
  const int TP = P.fast_tp;                  // smem row pitch (multiple of 16)
  const int nrows = strip.nrows;             // interior rows of this strip
  uint8_t* tile = smem;                      // (nrows + 6) x TP: level rows Y0-3.., ROI columns 0..
  uint8_t* score = smem + (size_t)TP * (P.fast_rows + 6);   // (nrows + 2) x TP with zero guards
  uint32_t* list = reinterpret_cast<uint32_t*>(smem + P.fast_list_off);   // kept maxima
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_n, s_ini[kMaxStripCells];
  uint4* s_mask = reinterpret_cast<uint4*>(score + (size_t)TP * (P.fast_rows + 2));   // [quads]
  unsigned short* queue = reinterpret_cast<unsigned short*>(s_mask + (TP >> 2));       // [ntask] quads that pass the compass test
  __shared__ int s_nq;
  const int tid = threadIdx.x;
  if (tid == 0) { s_n = 0; s_nq = 0; mbar_init(&s_bar, 1); }
  for (int i = tid; i < kMaxStripCells; i += THREADS) s_ini[i] = 0;
  __syncthreads();
  // ---- 1. TMA: the strip's rows Y0-3 .. Y1+2 as ONE tensor-map load (box = {row pitch of the tile, cell height + 6 rows, this frame};
  // rows past a short last strip and columns past the image are loaded or zero-filled and never looked at), issued by one thread.
  // Without tensor maps (driver entry point missing): one bulk copy per row.
  if (P.fast_tma2d) {
    if (tid == 0) {
      mbar_expect_tx(&s_bar, (uint32_t)TP * (uint32_t)(L.hCell + 6));
      tma_load_3d(tile, &maps.m[strip.level], (kXOff + strip.x0) >> 2, kEdge + strip.y0 - 3, f, &s_bar);
    }
  } else {
    const uint32_t row_bytes = (uint32_t)((min(strip.xw + 2 * kEdge, L.w - strip.x0) + 15) & ~15);
    const uint8_t* src = roi_ptr(P, L, f) + (long long)(strip.y0 - 3) * L.pitch + strip.x0;
    if (tid == 0) mbar_expect_tx(&s_bar, row_bytes * (uint32_t)(nrows + 6));
    for (int r = tid; r < nrows + 6; r += THREADS) bulk_g2s(tile + (size_t)r * TP, src + (long long)r * L.pitch, row_bytes, &s_bar);
  }
  // ---- 2. scores.  Score map: (nrows + 2) x TP4 words, one word = 4 pixels; word column 0 and
  // quads + 1 and rows 0 and nrows + 1 are zero guards (neighbours outside the strip belong to
  // other cells and count as 0).
  const int iw = strip.xw;                   // interior width of this strip (segment)
  const int quads = (iw + 3) >> 2;
  const int TP4 = TP >> 2;
  const int wCell = L.wCell;
  const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(tile);
  uint32_t* score32 = reinterpret_cast<uint32_t*>(score);
  // per-quad cell-boundary masks, one u16 lane per pixel {left even, left odd, right even, right odd}:
  // a pixel in the first (last) column of its cell has no left (right) neighbours
  for (int g = tid; g < quads; g += THREADS) {
    uint32_t ml = 0, mr = 0;
    int xm = (4 * g) % wCell;                                 // x mod wCell, stepped
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (xm != 0) ml |= 1u << k;
      if (++xm == wCell) xm = 0;
      if (xm != 0) mr |= 1u << k;
    }
    auto lanes = [](uint32_t m, int lo, int hi) { return ((m >> lo) & 1u ? 0xFFFFu : 0u) | ((m >> hi) & 1u ? 0xFFFF0000u : 0u); };
    s_mask[g] = make_uint4(lanes(ml, 0, 2), lanes(ml, 1, 3), lanes(mr, 0, 2), lanes(mr, 1, 3));
  }
  const uint32_t kmin8 = (uint32_t)(P.min_th << 8) * 0x00010001u;
  const int lane = tid & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  // the whole score map starts at zero: pixels the compass test rules out are never written
  {
    uint4* s4 = reinterpret_cast<uint4*>(score32);
    for (int i = tid; i < (nrows + 2) * (TP4 >> 2); i += THREADS) s4[i] = make_uint4(0, 0, 0, 0);
  }
  mbar_wait(&s_bar, 0);                      // the strip's rows have landed
  // Work queue of quads (entries ry << 10 | g): the compass test is a necessary condition, and every later stage
  // (score, NMS) touches only queued quads.  (Queueing pixel pairs instead was measured slower: when one pair of a
  // quad passes the other nearly always does, and two pair evaluations cost more than one quad evaluation.)
  const uint32_t kini = (uint32_t)(P.ini_th - P.min_th);                 // e > kini <=> q >= iniThFAST
  const int ncells_x = (iw + wCell - 1) / wCell;
  const int list_cap = P.fast_list_cap;
  const int groups = (quads + 3) >> 2;                                    // 16-pixel groups per row
  const int ngtask = nrows * groups;
  // queue every quad of the rows/groups selected by `want` that passes the compass test at threshold t
  auto build_queue = [&](int t, auto want) {
    const uint32_t cadd = (uint32_t)(127 - min(t, 127)) * 0x01010101u;
    for (int task0 = tid - lane; task0 < ngtask; task0 += THREADS) {
      const int task = task0 + lane;
      uint32_t fl = 0;
      int ry = 0, G = 0;
      if (task < ngtask) {
        ry = (int)__umulhi((uint32_t)task, strip.groups_magic);
        G = task - ry * groups;
        fl = fast_compass_group(tile32, TP4, ry, G, cadd);
        const int rem = quads - 4 * G;                          // quads of this group inside the row
        if (rem < 4) fl &= (1u << rem) - 1u;
        fl = want(G, fl);
      }
      const unsigned any = __ballot_sync(0xFFFFFFFFu, fl != 0);
      if (any == 0) continue;
      const unsigned b0 = __ballot_sync(0xFFFFFFFFu, fl & 1u), b1 = __ballot_sync(0xFFFFFFFFu, fl & 2u),
                     b2 = __ballot_sync(0xFFFFFFFFu, fl & 4u), b3 = __ballot_sync(0xFFFFFFFFu, fl & 8u);
      const int n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3);
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_nq, n0 + n1 + n2 + n3);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      const uint32_t ent = (uint32_t)(ry << 10) | (uint32_t)(4 * G);
      if (fl & 1u) queue[base + __popc(b0 & lt_mask)] = (unsigned short)ent;
      if (fl & 2u) queue[base + n0 + __popc(b1 & lt_mask)] = (unsigned short)(ent + 1);
      if (fl & 4u) queue[base + n0 + n1 + __popc(b2 & lt_mask)] = (unsigned short)(ent + 2);
      if (fl & 8u) queue[base + n0 + n1 + n2 + __popc(b3 & lt_mask)] = (unsigned short)(ent + 3);
    }
  };
  // scores of the queued quads -> score map, all lanes busy.  Every warp owns a contiguous chunk of the queue and
  // compacts it in place while it goes: only quads holding a score above the pass threshold (e > kth in some byte)
  // stay, so that the NMS pass below visits about half as many quads (a quad passing the compass test holds a
  // FAST(iniTh) corner in 47 % of the cases).  Returns the number of entries this warp kept.
  const int wid = tid >> 5;
  auto eval_queue = [&](uint32_t kth) -> int {
    const int nq = s_nq;
    const int chunk = (((nq + (THREADS >> 5) - 1) / (THREADS >> 5)) + 31) & ~31;
    const int q0 = wid * chunk, q1 = min(q0 + chunk, nq);
    const uint32_t cadd = (127u - min(kth, 127u)) * 0x01010101u;   // byte > kth  <=>  bit 7 of (a | ((a & 0x7f) + 127 - kth))
    int wp = q0;
    for (int qb = q0; qb < q1; qb += 32) {
      const int qi = qb + lane;
      uint32_t packed = 0;
      int task = 0;
      if (qi < q1) {
        task = queue[qi];
        const int ry = task >> 10, g = task & 1023;
        packed = fast_eval_quad(tile32, TP4, ry, g, kmin8);
        const int rem = iw - 4 * g;                              // pixels of this quad inside the interior
        if (rem < 4) packed &= 0xFFFFFFFFu >> (8 * (4 - rem));
        score32[(ry + 1) * TP4 + g + 1] = packed;
      }
      const bool stay = ((packed | ((packed & 0x7F7F7F7Fu) + cadd)) & 0x80808080u) != 0;
      const unsigned bal = __ballot_sync(0xFFFFFFFFu, stay);     // (every lane has read its entry: the ballot orders it)
      if (stay) queue[wp + __popc(bal & lt_mask)] = (unsigned short)task;
      wp += __popc(bal);
    }
    return wp - q0;
  };
  // ---- 3. per-cell non-maximum suppression: strict maximum over the 8 neighbours, branch-free.  Scores are
  // compared as u16 lanes whose HIGH byte is the pixel of interest and whose low byte is whatever byte precedes
  // it in the score row: min/max of such lanes have the right high byte, and "c > n" is tested as
  // (c & 0xff00) > (n | 0x00ff), so no byte is ever extracted.  Lanes of a funnel shift of the score words
  // (x0 | x1 | x2 = quads g-1, g, g+1), pixel numbers relative to quad g:
  //   x1: pixels 1,3    f24 = (x0:x1)>>24: pixels 0,2    f16 = (x0:x1)>>16: pixels -1,1    f8 = (x1:x2)>>8: pixels 2,4
  // Returns bit k = pixel k of the quad is a strict maximum with e > kth.
  auto nms_quad = [&](int ry, int g, uint32_t kth) -> uint32_t {
    const uint32_t* sp = score32 + (ry + 1) * TP4 + g + 1;
    const uint32_t u0 = sp[-TP4 - 1], u1 = sp[-TP4], u2 = sp[-TP4 + 1];
    const uint32_t m0 = sp[-1], m1 = sp[0], m2 = sp[1];
    const uint32_t d0 = sp[TP4 - 1], d1 = sp[TP4], d2 = sp[TP4 + 1];
    if (m1 == 0) return 0u;
    const uint4 mk = s_mask[g];
    const uint32_t u24 = __funnelshift_r(u0, u1, 24), u16 = __funnelshift_r(u0, u1, 16), u8 = __funnelshift_r(u1, u2, 8);
    const uint32_t m24 = __funnelshift_r(m0, m1, 24), m16 = __funnelshift_r(m0, m1, 16), m8 = __funnelshift_r(m1, m2, 8);
    const uint32_t d24 = __funnelshift_r(d0, d1, 24), d16 = __funnelshift_r(d0, d1, 16), d8 = __funnelshift_r(d1, d2, 8);
    const uint32_t th = (kth << 8 | 0xFFu) * 0x00010001u;
    // even pixels 0,2: centre m24, left column f16, right column x1, above/below f24
    const uint32_t nb_e = __vimax3_u16x2(__vimax3_u16x2(u16, m16, d16) & mk.x, __vimax3_u16x2(u1, m1, d1) & mk.z, __vmaxu2(u24, d24));
    // odd pixels 1,3: centre x1, left column f24, right column f8, above/below x1
    const uint32_t nb_o = __vimax3_u16x2(__vimax3_u16x2(u24, m24, d24) & mk.y, __vimax3_u16x2(u8, m8, d8) & mk.w, __vmaxu2(u1, d1));
    const uint32_t ce = m24 & 0xFF00FF00u, co = m1 & 0xFF00FF00u;
    const uint32_t te = ce - __vminu2(ce, __vmaxu2(nb_e | 0x00FF00FFu, th));      // per lane: > 0 <=> kept
    const uint32_t to = co - __vminu2(co, __vmaxu2(nb_o | 0x00FF00FFu, th));
    return ((te & 0xFFFFu) ? 1u : 0u) | ((to & 0xFFFFu) ? 2u : 0u) | ((te >> 16) ? 4u : 0u) | ((to >> 16) ? 8u : 0u);
  };
  // Kept maxima -> list.  FAST(iniThFAST) maxima mark their cell; FAST(minThFAST) maxima are taken only in cells
  // that have none (the fallback applies to the pixel's own cell, ORBextractor.cc:812-816).  List slots come from
  // one ballot + one shared atomic per warp and round.
  auto nms_queue = [&](uint32_t kth, bool fallback, int kept) {
    const int nq = s_nq;
    const int chunk = (((nq + (THREADS >> 5) - 1) / (THREADS >> 5)) + 31) & ~31;
    const int q0 = wid * chunk, q1 = q0 + kept;                 // this warp's compacted entries
    for (int qb = q0; qb < q1; qb += 32) {
      const int qi = qb + lane;
      uint32_t surv = 0, c0 = 0;
      int ry = 0, g = 0;
      if (qi < q1) {
        const int task = queue[qi];
        ry = task >> 10; g = task & 1023;
        surv = nms_quad(ry, g, kth);
        if (surv) {
          c0 = score32[(ry + 1) * TP4 + g + 1];
          const int ca = (int)__umulhi((uint32_t)(4 * g), L.wcell_magic), cb = (int)__umulhi((uint32_t)min(4 * g + 3, iw - 1), L.wcell_magic);
          if (fallback) {
            // cells that had FAST(iniThFAST) corners keep those only
            if (ca == cb) { if (s_ini[ca] != 0) surv = 0; }
            else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if ((surv >> k & 1u) && s_ini[(int)__umulhi((uint32_t)(4 * g + k), L.wcell_magic)] != 0) surv &= ~(1u << k);
            }
          } else {
            if (ca == cb) s_ini[ca] = 1;                         // benign race: everybody writes 1
            else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (surv >> k & 1u) s_ini[(int)__umulhi((uint32_t)(4 * g + k), L.wcell_magic)] = 1;
            }
          }
        }
      }
      // usually one round; a quad straddling a cell edge can hold up to three maxima
      unsigned bal = __ballot_sync(0xFFFFFFFFu, surv != 0);
      while (bal) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_n, __popc(bal));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (surv) {
          const int k = __ffs(surv) - 1, slot = base + __popc(bal & lt_mask);
          const uint32_t sc = (c0 >> (8 * k)) & 0xFFu;
          if (slot < list_cap) list[slot] = (uint32_t)(4 * g + k) | ((uint32_t)ry << 12) | (sc << 24);
          else {
            // the shared list holds one maximum per 8 px; denser strips (strict 3x3 maxima can reach one per 4 px) put the
            // rest straight into the level's arena, which is sized for the densest possible image
            const int pos = atomicAdd(P.cand_cnt + f * P.nlevels + strip.level, 1);
            if (pos < L.cand_cap)
              (P.cand + (long long)f * P.cand_fstride + L.cand_off)[pos] =
                  (uint32_t)(4 * g + k + 3 + strip.x0) | ((uint32_t)(ry + strip.y0 - kEdge + 3) << 12) | ((sc + P.min_th - 1) << 24);
            else atomicOr(P.status, 1);
          }
          surv &= surv - 1u;
        }
        bal = __ballot_sync(0xFFFFFFFFu, surv != 0);
      }
    }
  };
  // ---- pass 0: what FAST(iniThFAST) returns
  build_queue(P.ini_th, [](int, uint32_t fl) { return fl; });
  __syncthreads();
  const int kept0 = eval_queue(kini);
  __syncthreads();
  nms_queue(kini, false, kept0);
  __syncthreads();
  // ---- pass 1: cells where FAST(iniThFAST) found nothing get the maxima of FAST(minThFAST)
  {
    int missing = 0;
    for (int c = tid; c < ncells_x; c += THREADS) missing |= (s_ini[c] == 0);
    if (__syncthreads_or(missing)) {
      // only quads that touch a cell without corners
      auto needs = [&](int G, uint32_t fl) -> uint32_t {
        if (fl == 0) return 0u;
        const int x0 = 16 * G;
        const int ca = (int)__umulhi((uint32_t)x0, L.wcell_magic), cb = (int)__umulhi((uint32_t)min(x0 + 15, iw - 1), L.wcell_magic);
        if (ca == cb) return s_ini[ca] == 0 ? fl : 0u;
        uint32_t keep = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int xa = min(x0 + 4 * k, iw - 1), xb = min(x0 + 4 * k + 3, iw - 1);
          if (s_ini[(int)__umulhi((uint32_t)xa, L.wcell_magic)] == 0 || s_ini[(int)__umulhi((uint32_t)xb, L.wcell_magic)] == 0) keep |= 1u << k;
        }
        return fl & keep;
      };
      if (tid == 0) s_nq = 0;
      __syncthreads();
      build_queue(P.min_th, needs);
      __syncthreads();
      const int kept1 = eval_queue(0u);
      __syncthreads();
      nms_queue(0u, true, kept1);
    }
  }
  __syncthreads();
  // ---- 4. emit (region coordinates: origin (16,16) => interior x + 3), order-free
  const int n = min(s_n, list_cap);
  if (n == 0) return;
  __shared__ int s_base;
  if (tid == 0) s_base = atomicAdd(P.cand_cnt + f * P.nlevels + strip.level, n);
  __syncthreads();
  uint32_t* out = P.cand + (long long)f * P.cand_fstride + L.cand_off;
  const int oy = strip.y0 - kEdge + 3;
  const int base = s_base;
  for (int i = tid; i < n; i += THREADS) {
    const uint32_t ent = list[i];
    const uint32_t q = (ent >> 24) + P.min_th - 1;
    const int pos = base + i;
    if (pos < L.cand_cap)
      out[pos] = ((ent & 0xFFF) + 3 + strip.x0) | ((((ent >> 12) & 0xFFF) + oy) << 12) | (q << 24);
    else
      atomicOr(P.status, 1);
  }
}

// ------------------------------------------------------------------ K3 quadtree
// In-place exclusive scan of a[0..n) by the whole block; returns the total to every thread.
template <int THREADS>
__device__ int block_excl_scan(int* a, int n, int* warp_tot) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int chunk = (n + THREADS - 1) / THREADS;
  const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += a[i];
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  int base = 0, total = 0;
#pragma unroll
  for (int wI = 0; wI < THREADS / 32; ++wI) {
    const int t = warp_tot[wI];
    if (wI < wid) base += t;
    total += t;
  }
  int run = base + inc - sum;
  for (int i = lo; i < hi; ++i) { const int v = a[i]; a[i] = run; run += v; }
  __syncthreads();
  return total;
}

struct QBox { short x0, y0, x1, y1; };

__device__ __forceinline__ void child_mid(const QBox& b, int& mx, int& my) {
  // halfX = ceil(float(x1-x0)/2) (ORBextractor.cc:483-484); exact in integers for |w| < 2^24
  const int wdt = b.x1 - b.x0, hgt = b.y1 - b.y0;
  const int hx = (wdt >= 0) ? (wdt + 1) / 2 : -((-wdt) / 2);
  const int hy = (hgt >= 0) ? (hgt + 1) / 2 : -((-hgt) / 2);
  mx = b.x0 + hx; my = b.y0 + hy;
}

static const int kQuadStage = 6144;       // candidate keys of one (frame, level) list held in shared memory by k_quadtree
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_quadtree(const OrbDev* __restrict__ Pp, int f0) {
  DRFE_GRID_DEP();
  extern __shared__ __align__(128) uint8_t smem[];
  const OrbDev& P = *Pp;
  const int level = blockIdx.x, f = blockIdx.y + f0;
  const LevelDev& L = P.lv[level];
  const int NC = L.node_cap, N = L.nfeat;
  const int tid = threadIdx.x;
  // ---- shared layout
  unsigned long long* best = reinterpret_cast<unsigned long long*>(smem);
  QBox* boxA = reinterpret_cast<QBox*>(best + NC);
  QBox* boxB = boxA + NC;
  int* cntA = reinterpret_cast<int*>(boxB + NC);
  int* cntB = cntA + NC;
  int* ccnt = cntB + NC;        // [NC*4] children counts of split candidates
  int* a_rank = ccnt + 4 * NC;  // per pos: rank among candidates or -1
  int* a_order = a_rank + NC;   // rank -> pos
  int* a_off = a_order + NC;    // per rank: creation offset
  int* a_keep = a_off + NC;     // per pos: index among kept nodes
  int* a_tmp = a_keep + NC;
  unsigned short* remap = reinterpret_cast<unsigned short*>(a_tmp + NC);  // [NC*4]
  unsigned short* remap_keep = remap + 4 * NC;                            // [NC]
  __shared__ int warp_tot[THREADS / 32];
  __shared__ int s_nsplit, s_expand;

  const int n = min(P.cand_cnt[f * P.nlevels + level], L.cand_cap);
  const uint32_t* keys = P.cand + (long long)f * P.cand_fstride + L.cand_off;
  uint16_t* node_of = P.node_of + (long long)f * P.cand_fstride + L.cand_off;
  int* out_cnt = P.lkp_cnt + f * P.nlevels + level;
  if (n == 0) { if (tid == 0) *out_cnt = 0; return; }
  // every pass walks all keys (packed position + response, and the node each key sits in): up to kQuadStage of them live in shared
  // memory for the whole kernel — from global memory each pass paid an L2 round trip per key (node_of is rewritten every pass,
  // so the L1 never holds it); the kernel is a chain of such passes (58 -> ~25 us for one 640x480 level-0 list)
  if (n <= kQuadStage) {
    uint32_t* s_keys = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(remap_keep + NC) + 15) & ~(uintptr_t)15);
    for (int k0 = tid; k0 < n; k0 += 4 * THREADS) {
      uint32_t e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) e[u] = (k0 + u * THREADS < n) ? keys[k0 + u * THREADS] : 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u) if (k0 + u * THREADS < n) s_keys[k0 + u * THREADS] = e[u];
    }
    keys = s_keys;
    node_of = reinterpret_cast<uint16_t*>(s_keys + kQuadStage);
    __syncthreads();
  }

  QBox* box = boxA; QBox* nbox = boxB;
  int* cnt = cntA; int* ncnt = cntB;

  // ---- roots (ORBextractor.cc:543-592): key -> root by int(x / hX); empty roots dropped
  for (int i = tid; i < NC; i += THREADS) { cnt[i] = 0; }
  __syncthreads();
  for (int k0 = tid; k0 < n; k0 += 4 * THREADS) {
    int r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = (k0 + u * THREADS < n) ? (int)(keys[k0 + u * THREADS] & 0xFFF) : 0;
      r[u] = min((int)__fdiv_rn((float)x, L.hX), L.nIni - 1);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (k0 + u * THREADS < n) { node_of[k0 + u * THREADS] = (uint16_t)r[u]; atomicAdd(&cnt[r[u]], 1); }
  }
  __syncthreads();
  int S = 0;
  {
    // compact non-empty roots (nIni <= 4: done redundantly by every thread)
    int newpos[4];
    for (int i = 0; i < L.nIni; ++i) { newpos[i] = S; if (cnt[i] > 0) ++S; }
    int c[4];
    for (int i = 0; i < L.nIni; ++i) c[i] = cnt[i];
    __syncthreads();
    if (tid == 0)
      for (int i = 0; i < L.nIni; ++i)
        if (c[i] > 0) {
          cnt[newpos[i]] = c[i];
          box[newpos[i]] = QBox{(short)L.root_x[i], 0, (short)L.root_x[i + 1], (short)L.regH};
        }
    for (int k = tid; k < n; k += THREADS) node_of[k] = (uint16_t)newpos[node_of[k]];
    __syncthreads();
  }

  bool fine = false;
  int Cprev = 0;
  for (int iter = 0; iter < 64; ++iter) {
    const int prevS = S;
    // 1. split candidates: coarse = every multi-key node; fine = multi-key nodes created by
    //    the previous pass (they sit at the list front)
    for (int p = tid; p < S; p += THREADS) {
      const bool cand = cnt[p] > 1 && (!fine || p < Cprev);
      a_tmp[p] = cand ? 1 : 0;
      ccnt[4 * p] = ccnt[4 * p + 1] = ccnt[4 * p + 2] = ccnt[4 * p + 3] = 0;
    }
    if (tid == 0) { s_nsplit = 0x7FFFFFFF; s_expand = 0; }
    __syncthreads();
    // compact index of candidates in list order
    for (int p = tid; p < S; p += THREADS) a_rank[p] = a_tmp[p];
    __syncthreads();
    const int V = block_excl_scan<THREADS>(a_rank, S, warp_tot);   // a_rank[p] = compact idx
    if (V == 0) break;                                             // nothing to split: size unchanged
    // 2. classify the keys of candidate nodes into their 4 children (DivideNode :515-531)
    for (int k0 = tid; k0 < n; k0 += 4 * THREADS) {              // four keys per trip: the loads of all four go out first
      int nd[4];
      uint32_t e[4];
      bool act[4];
      QBox bx[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * THREADS;
        nd[u] = (k < n) ? (node_of[k] & 0x3FFF) : 0;
        e[u] = (k < n) ? keys[k] : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { act[u] = (k0 + u * THREADS < n) && a_tmp[nd[u]]; bx[u] = box[nd[u]]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!act[u]) continue;
        const int x = e[u] & 0xFFF, y = (e[u] >> 12) & 0xFFF;
        int mx, my;
        child_mid(bx[u], mx, my);
        const int c = (x < mx) ? ((y < my) ? 0 : 2) : ((y < my) ? 1 : 3);
        atomicAdd(&ccnt[4 * nd[u] + c], 1);
        node_of[k0 + u * THREADS] = (uint16_t)(nd[u] | (c << 14));
      }
    }
    // 3. split order.  coarse: list order.  fine: sort by (nKeys, creation seq) ascending and
    //    walk from the back (:684-686) == (nKeys desc, list position asc).
    for (int p = tid; p < S; p += THREADS)
      if (a_tmp[p]) a_order[a_rank[p]] = p;       // compact list (coarse: final order)
    __syncthreads();
    if (fine) {
      for (int i = tid; i < V; i += THREADS) a_off[i] = a_order[i];  // stash compact list
      __syncthreads();
      for (int i = tid; i < V; i += THREADS) {
        const int p = a_off[i], cp = cnt[p];
        int r = 0;
        for (int j = 0; j < V; ++j) {
          const int pj = a_off[j], cj = cnt[pj];
          r += (cj > cp) || (cj == cp && pj < p);
        }
        a_keep[i] = r;                                            // rank of compact entry i
      }
      __syncthreads();
      for (int i = tid; i < V; i += THREADS) { a_order[a_keep[i]] = a_off[i]; }
      __syncthreads();
      for (int r = tid; r < V; r += THREADS) a_rank[a_order[r]] = r;
      __syncthreads();
      // stop splitting as soon as the list reaches N nodes (:730-731)
      for (int r = tid; r < V; r += THREADS) {
        const int p = a_order[r];
        const int nch = (ccnt[4 * p] > 0) + (ccnt[4 * p + 1] > 0) + (ccnt[4 * p + 2] > 0) + (ccnt[4 * p + 3] > 0);
        a_off[r] = nch - 1;
      }
      __syncthreads();
      // inclusive gains
      {
        // exclusive scan then add own gain
        for (int r = tid; r < V; r += THREADS) a_keep[r] = a_off[r];
        __syncthreads();
        block_excl_scan<THREADS>(a_keep, V, warp_tot);
        for (int r = tid; r < V; r += THREADS)
          if (S + a_keep[r] + a_off[r] >= N) atomicMin(&s_nsplit, r + 1);
        __syncthreads();
      }
    }
    const int nsplit = min(s_nsplit, V);
    // 4. creation offsets in split order
    for (int r = tid; r < V; r += THREADS) {
      const int p = a_order[r];
      const int nch = (ccnt[4 * p] > 0) + (ccnt[4 * p + 1] > 0) + (ccnt[4 * p + 2] > 0) + (ccnt[4 * p + 3] > 0);
      a_off[r] = (r < nsplit) ? nch : 0;
    }
    for (int p = tid; p < S; p += THREADS) a_keep[p] = (a_tmp[p] && a_rank[p] < nsplit) ? 0 : 1;
    __syncthreads();
    const int C = block_excl_scan<THREADS>(a_off, V, warp_tot);
    const int K = block_excl_scan<THREADS>(a_keep, S, warp_tot);
    const int newS = C + K;
    if (newS > NC) { if (tid == 0) atomicOr(P.status, 2); break; }
    // 5. build the new list: children are pushed to the front one by one (n1..n4, :636-673),
    //    so the list starts with the created nodes in reverse creation order, followed by
    //    the surviving nodes in their old order.
    for (int r = tid; r < nsplit; r += THREADS) {
      const int p = a_order[r];
      const QBox b = box[p];
      int mx, my;
      child_mid(b, mx, my);
      int ci = a_off[r], nexp = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int cc = ccnt[4 * p + c];
        if (cc == 0) continue;
        const int np = C - 1 - ci;
        ++ci;
        QBox nb;
        nb.x0 = (c & 1) ? (short)mx : b.x0; nb.x1 = (c & 1) ? b.x1 : (short)mx;
        nb.y0 = (c & 2) ? (short)my : b.y0; nb.y1 = (c & 2) ? b.y1 : (short)my;
        nbox[np] = nb; ncnt[np] = cc;
        remap[4 * p + c] = (unsigned short)np;
        nexp += (cc > 1);
      }
      if (nexp) atomicAdd(&s_expand, nexp);
    }
    for (int p = tid; p < S; p += THREADS) {
      if (a_tmp[p] && a_rank[p] < nsplit) continue;
      const int np = C + a_keep[p];
      nbox[np] = box[p]; ncnt[np] = cnt[p];
      remap_keep[p] = (unsigned short)np;
    }
    __syncthreads();
    for (int k0 = tid; k0 < n; k0 += 4 * THREADS) {
      int v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (k0 + u * THREADS < n) ? node_of[k0 + u * THREADS] : 0;
      bool sp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int nd = v[u] & 0x3FFF; sp[u] = a_tmp[nd] && a_rank[nd] < nsplit; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int nd = v[u] & 0x3FFF;
        if (k0 + u * THREADS < n) node_of[k0 + u * THREADS] = sp[u] ? remap[4 * nd + (v[u] >> 14)] : remap_keep[nd];
      }
    }
    const int nexpand = s_expand;
    __syncthreads();
    { QBox* t = box; box = nbox; nbox = t; int* u = cnt; cnt = ncnt; ncnt = u; }
    S = newS; Cprev = C;
    // 6. termination (:677-681, :735-736)
    if (S >= N || S == prevS) break;
    if (!fine && S + 3 * nexpand > N) fine = true;
  }
  __syncthreads();
  // ---- keep the best key per node: max response, first in the reference's candidate order
  // (cell row-major, then pixel row-major) on ties (:741-760).
  for (int p = tid; p < S; p += THREADS) best[p] = 0ull;
  __syncthreads();
  for (int k0 = tid; k0 < n; k0 += 4 * THREADS) {
    unsigned long long val[4];
    int nd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = min(k0 + u * THREADS, n - 1);
      const uint32_t e = keys[k];
      nd[u] = node_of[k] & 0x3FFF;
      const int x = e & 0xFFF, y = (e >> 12) & 0xFFF;
      const unsigned cellid = (unsigned)(((y - 3) / L.hCell) * L.nCols + (x - 3) / L.wCell);
      const unsigned long long ord = ((unsigned long long)cellid << 24) | ((unsigned long long)y << 12) | x;
      val[u] = ((unsigned long long)(e >> 24) << 40) | (0xFFFFFFFFFFull - ord);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (k0 + u * THREADS < n) atomicMax(&best[nd[u]], val[u]);
  }
  __syncthreads();
  uint32_t* out = P.lkp + (long long)f * P.lkp_fstride + L.kp_off;
  for (int p = tid; p < S; p += THREADS) {
    const unsigned long long v = best[p];
    const unsigned long long ord = 0xFFFFFFFFFFull - (v & 0xFFFFFFFFFFull);
    const uint32_t x = (uint32_t)(ord & 0xFFF) + (kEdge - 3), y = (uint32_t)((ord >> 12) & 0xFFF) + (kEdge - 3);
    out[p] = x | (y << 12) | ((uint32_t)(v >> 40) << 24);
  }
  if (tid == 0) *out_cnt = S;
}

// ------------------------------------------------------------------ K5 Gaussian
// 7x7 sigma=2 separable, 8.8 fixed point {18,34,48,56,48,34,18}, rounding (v + 2^15) >> 16
// (SURVEY App. A.5).  Reads the bordered level image, whose 19 px reflect-101 frame is
// exactly the BORDER_REFLECT_101 extension GaussianBlur applies to the cloned level.
// One thread = a 4-pixel-wide column of kBlurRows output rows: per input row three aligned
// words, the horizontal pass as two u8 dot products (IDP.4A) per pixel, the vertical pass over
// a 7-row register window.  No shared memory: neighbouring threads hit the same L1 lines.
// rows per thread: measured 0.286 / 0.220 / 0.246 / 0.268 ms per 256 frames for 4 / 8 / 16 / 32 — the kernel is bound by
// load latency (long-scoreboard stalls), so more, shorter threads win over fewer redundant row filters
static const int kBlurRows = 8, kBlurAhead = 2;
struct BlurRow { uint32_t a, b, c; };                                   // columns x-4 .. x+7 of one source row
__device__ __forceinline__ BlurRow blur_load(const uint8_t* __restrict__ row) {
  const uint32_t* p = reinterpret_cast<const uint32_t*>(row);
  BlurRow w;
  w.a = __ldg(p); w.b = __ldg(p + 1); w.c = __ldg(p + 2);
  return w;
}
__device__ __forceinline__ void blur_hfilter(const BlurRow& w, int (&h)[4]) {
  const uint32_t k0 = 18u | (34u << 8) | (48u << 16) | (56u << 24);    // taps -3..0
  const uint32_t k1 = 48u | (34u << 8) | (18u << 16);                  // taps +1..+3
  h[0] = __dp4a(__funnelshift_r(w.b, w.c, 8), k1, __dp4a(__funnelshift_r(w.a, w.b, 8), k0, 0u));
  h[1] = __dp4a(__funnelshift_r(w.b, w.c, 16), k1, __dp4a(__funnelshift_r(w.a, w.b, 16), k0, 0u));
  h[2] = __dp4a(__funnelshift_r(w.b, w.c, 24), k1, __dp4a(__funnelshift_r(w.a, w.b, 24), k0, 0u));
  h[3] = __dp4a(w.c, k1, __dp4a(w.b, k0, 0u));
}
__device__ __forceinline__ void blur_hrow(const uint8_t* __restrict__ row, int (&h)[4]) { blur_hfilter(blur_load(row), h); }

__global__ void __launch_bounds__(256) k_blur(const OrbDev* __restrict__ Pp, int f0, int check_counts) {
  DRFE_GRID_DEP();
  const OrbDev& P = *Pp;
  int level = 0;
  while (level + 1 < P.nlevels && (int)blockIdx.x >= P.blur_blk_off[level + 1]) ++level;
  const LevelDev& L = P.lv[level];
  const int f = blockIdx.y + f0;
  // the reference skips levels without keypoints (:1081); when the blur runs beside FAST and the quadtree the counts are not there yet
  // and every level is blurred (a blurred level without keypoints is never read)
  if (check_counts && P.lkp_cnt[f * P.nlevels + level] == 0) return;
  // flattened (row group, 4-pixel column) index, columns fastest
  const int t = ((int)blockIdx.x - P.blur_blk_off[level]) * 256 + threadIdx.x;
  const int rg = (int)__umulhi((uint32_t)t, L.blur_magic);
  const int q = t - rg * L.blur_nq;
  const int y0 = rg * kBlurRows;
  if (y0 >= L.h) return;
  const int nrows = min(kBlurRows, L.h - y0);
  const uint8_t* src = roi_ptr(P, L, f) + (long long)(y0 - 3) * L.pitch + 4 * q - 4;
  uint8_t* dst = P.blur + L.blur_off + (long long)f * L.blur_fstride + (long long)y0 * L.bpitch + 4 * q;
  int h[7][4];
#pragma unroll
  for (int r = 0; r < 6; ++r) blur_hrow(src + (long long)r * L.pitch, h[r]);
  // the rows to come are loaded kBlurAhead steps before they are filtered (the kernel was waiting on its loads: 9.6
  // warps per issue slot stalled on the long scoreboard).  Rows past the strip are inside the 19-px border frame or the
  // spare row of the level buffer, so the prefetch needs no bounds test.
  BlurRow nx[kBlurAhead];
#pragma unroll
  for (int k = 0; k < kBlurAhead; ++k) nx[k] = blur_load(src + (long long)(6 + k) * L.pitch);
#pragma unroll
  for (int r = 0; r < kBlurRows; ++r) {
    if (r < nrows) {
      blur_hfilter(nx[r % kBlurAhead], h[(r + 6) % 7]);
      if (r + kBlurAhead < nrows) nx[r % kBlurAhead] = blur_load(src + (long long)(r + 6 + kBlurAhead) * L.pitch);
      uint32_t o = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int v = 18 * (h[r % 7][j] + h[(r + 6) % 7][j]) + 34 * (h[(r + 1) % 7][j] + h[(r + 5) % 7][j]) +
                      48 * (h[(r + 2) % 7][j] + h[(r + 4) % 7][j]) + 56 * h[(r + 3) % 7][j] + 32768;
        o |= (uint32_t)(v >> 16) << (8 * j);                            // <= 255 by construction
      }
      *reinterpret_cast<uint32_t*>(dst + (long long)r * L.bpitch) = o;
    }
  }
}

// ------------------------------------------------------------------ K4+K6 orientation + descriptor
// cv::fastAtan2 (degrees), plain float32 with explicit rounding of every op
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float s = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s,
              p7 = -0.04432655554792128f * s;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// A warp owns kOdG consecutive keypoints of one (frame, level) and runs three phases, so that the scalar parts of
// the work are spread over the lanes instead of being paid once per keypoint by a whole warp, and so that every
// global access is a coalesced asynchronous copy (the kernel is bound by L1 wavefronts, not by issue slots: with one
// lane per image row in the loads, 9 loads touched 279 lines per keypoint and the L1 ran at 95 %):
//   1. IC_Angle moments, one keypoint at a time.  The 31 x 36-byte window of the level image (aligned words around the
//      radius-15 disc) is staged in shared memory by 4-byte cp.async, consecutive lanes = consecutive words, three
//      keypoints ahead; then lane l = disc row v = l - 15 reads its 9 words (pitch 9 words: conflict-free), realigns
//      them by funnel shifts, masks the bytes outside the disc (|u| <= umax[|v|]) and gets sum(u * I) and sum(I) as
//      8 + 8 four-way dot products (IDP.4A) instead of 31 byte loads and multiplies; m01 = v * sum(I).  Lane k keeps
//      the reduced moments of keypoint k.
//   2. every lane finishes its own keypoint in parallel: fastAtan2, the correctly rounded sine / cosine (double),
//      the cv::KeyPoint record.  (A warp per keypoint spent a quarter of its instructions here with one lane active.)
//   3. rBRIEF, one keypoint at a time: 39 rows x 64 bytes of the blurred level (|rotated pattern coordinate| <= 18
//      < EDGE_THRESHOLD, plus the slack of a 16-byte aligned start) are staged by 16-byte cp.async into a two-slot ring,
//      keypoint k+1's copy in flight while keypoint k's 512 steered samples are gathered; lane j computes descriptor
//      byte j, its 8 tests' pattern points live in 8 registers as packed bytes; no I2F / F2I anywhere (see below).
static const int kOdWarps = 4, kOdG = 16;
static const int kRawRows = 2 * kHalfPatch + 1, kRawW4 = 9, kRawWords = kRawRows * kRawW4, kRawSlots = 4;   // 31 rows x 36 bytes
static const int kRawIters = (kRawWords + 31) / 32;
static const int kPatchRows = 2 * kEdge + 1, kPatchPitch = 80, kPatchChunks = kPatchRows * 4;   // 39 rows, 64 bytes used of an 80-byte pitch (bank spread)
static const int kPatchIters = (kPatchChunks + 31) / 32;
static const int kOdWarpBytes = 2 * kPatchRows * kPatchPitch;                                    // per warp: the larger of the two rings
static_assert(kRawSlots * kRawWords * 4 <= kOdWarpBytes, "the raw ring shares the patch ring's memory");
__device__ uint32_t g_pattern_s8[256];   // entry t * 32 + j = test t of descriptor byte j: x0 | y0 << 8 | x1 << 16 | y1 << 24, each coordinate + 32 as a byte

__device__ __forceinline__ int dp4a_su(uint32_t w_s8, uint32_t d_u8, int acc) {   // signed weights x unsigned bytes
  int r;
  asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(w_s8), "r"(d_u8), "r"(acc));
  return r;
}

__global__ void __launch_bounds__(kOdWarps * 32, 8) k_orient_describe(const OrbDev* __restrict__ Pp, int f0) {
  DRFE_GRID_DEP();
  __shared__ __align__(16) uint8_t s_ring[kOdWarps][kOdWarpBytes];
  const OrbDev& P = *Pp;
  const int level = blockIdx.y, f = blockIdx.z + f0;
  const LevelDev& L = P.lv[level];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int* cnts = P.lkp_cnt + f * P.nlevels;
  const int nk = cnts[level];
  if (blockIdx.x == 0 && level == 0 && threadIdx.x == 0) {
    int tot = 0;
    for (int l = 0; l < P.nlevels; ++l) tot += cnts[l];
    P.out_cnt[f] = min(tot, P.kp_cap);
  }
  const int k0 = ((int)blockIdx.x * kOdWarps + wid) * kOdG;     // this warp's first keypoint of the level
  if (k0 >= nk) return;
  const int ng = min(kOdG, nk - k0);
  int base = k0;
  for (int l = 0; l < level; ++l) base += cnts[l];              // output slot of keypoint k0
  const uint32_t e_mine = lane < ng ? P.lkp[(long long)f * P.lkp_fstride + L.kp_off + k0 + lane] : 0u;   // x | y << 12 | response << 24
  const uint32_t ring = smem_u32(s_ring[wid]);

  // ---- 1. moments
  // staging of keypoint k's window into raw slot k & 3: word idx = lane + 32 * it -> row idx / 9, word idx % 9
  const uint8_t* img = roi_ptr(P, L, f);
  const int pitch = L.pitch;
  const int rr0 = lane / kRawW4, rw0 = lane - rr0 * kRawW4;
  auto stage_raw = [&](int k) {
    if (k < ng) {
      const uint32_t e = __shfl_sync(0xFFFFFFFFu, e_mine, k);
      const int cx = e & 0xFFF, cy = (e >> 12) & 0xFFF;
      const uint8_t* src = img + (long long)(cy - kHalfPatch) * pitch + ((cx - kHalfPatch) & ~3);   // ROI rows are 4-byte aligned at column 0
      const uint32_t dst = ring + (uint32_t)((k & (kRawSlots - 1)) * kRawWords * 4) + 4u * (uint32_t)lane;
      int r = rr0, wc = rw0;
#pragma unroll
      for (int it = 0; it < kRawIters; ++it) {
        if (it < kRawIters - 1 || lane + 32 * it < kRawWords)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 128u * (uint32_t)it), "l"(src + r * pitch + 4 * wc) : "memory");
        r += 3; wc += 32 - 3 * kRawW4;                           // idx += 32 = 3 * 9 + 5
        if (wc >= kRawW4) { wc -= kRawW4; ++r; }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");        // (an empty group when k >= ng keeps the wait counts uniform)
  };
  stage_raw(0); stage_raw(1); stage_raw(2);
  // byte b of the realigned row is column u = b - 15; bytes outside the disc are masked to 0
  const int v = lane - kHalfPatch;
  uint32_t dmask[8];
  {
    const int m = lane < kRawRows ? c_umax[abs(v)] : -1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t w = 0;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (abs(4 * j + t - kHalfPatch) <= m) w |= 0xFFu << (8 * t);
      dmask[j] = w;
    }
  }
  int my10 = 0, my01 = 0;
  for (int k = 0; k < ng; ++k) {
    stage_raw(k + 3);                                           // slot (k + 3) & 3 was read by keypoint k - 1: every lane is past its reduction
    asm volatile("cp.async.wait_group 3;" ::: "memory");
    __syncwarp();
    const uint32_t e = __shfl_sync(0xFFFFFFFFu, e_mine, k);
    const int sh = 8 * (((int)(e & 0xFFF) - kHalfPatch) & 3);
    int su = 0, s1 = 0;
    if (lane < kRawRows) {
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(s_ring[wid]) + (k & (kRawSlots - 1)) * kRawWords + lane * kRawW4;
      uint32_t w[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) w[j] = rw[j];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t d = __funnelshift_r(w[j], w[j + 1], sh) & dmask[j];
        const int u0 = 4 * j - kHalfPatch;                      // weights u0 .. u0 + 3 as signed bytes
        const uint32_t wu = (uint32_t)(u0 & 0xFF) | ((uint32_t)((u0 + 1) & 0xFF) << 8) | ((uint32_t)((u0 + 2) & 0xFF) << 16) | ((uint32_t)((u0 + 3) & 0xFF) << 24);
        su = dp4a_su(wu, d, su);
        s1 = (int)__dp4a(d, 0x01010101u, (uint32_t)s1);
      }
    }
    int m10 = su, m01 = v * s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o);
      m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o);
    }
    if (lane == k) { my10 = m10; my01 = m01; }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");         // (only empty groups are left)
  __syncwarp();                                                 // the ring changes hands: raw windows -> blurred patches

  // ---- staging of keypoint k's blurred patch into patch slot k & 1: 16-byte chunk idx = lane + 32 * it -> row idx / 4
  const uint8_t* blur = P.blur + L.blur_off + (long long)f * L.blur_fstride;
  const int bpitch = L.bpitch;
  auto stage_patch = [&](int k) {
    const uint32_t e = __shfl_sync(0xFFFFFFFFu, e_mine, k);
    const int cx = e & 0xFFF, cy = (e >> 12) & 0xFFF;
    const uint8_t* src = blur + (long long)(cy - kEdge + (lane >> 2)) * bpitch + ((cx - kEdge) & ~15) + 16 * (lane & 3);
    const uint32_t dst = ring + (uint32_t)((k & 1) * kPatchRows * kPatchPitch) + (uint32_t)((lane >> 2) * kPatchPitch + 16 * (lane & 3));
#pragma unroll
    for (int it = 0; it < kPatchIters; ++it) {
      if (it < kPatchIters - 1 || lane + 32 * it < kPatchChunks)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(8 * it * kPatchPitch)), "l"(src + (long long)(8 * it) * bpitch) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_patch(0);                                               // in flight during phase 2

  // ---- 2. angle, sine / cosine and the keypoint record of this lane's keypoint
  float a = 0.f, b = 0.f;
  if (lane < ng) {
    const float angle = fast_atan2_deg((float)my01, (float)my10);
    const float rad = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.f));
    double sd, cd;
    sincos((double)rad, &sd, &cd);   // correctly rounded cosf/sinf: double result rounded to float
    a = (float)cd;
    b = (float)sd;
    const int o = base + lane;
    if (o < P.kp_cap) {
      const int cx = e_mine & 0xFFF, cy = (e_mine >> 12) & 0xFFF;   // cvRound of integral coords
      drfe_keypoint kp;
      // pt *= mvScaleFactor[level] for level > 0 (ORBextractor.cc:1095-1101)
      kp.x = level ? __fmul_rn((float)cx, L.scale) : (float)cx;
      kp.y = level ? __fmul_rn((float)cy, L.scale) : (float)cy;
      kp.size = L.size;
      kp.angle = angle;
      kp.response = (float)(e_mine >> 24);
      kp.octave = level;
      kp.class_id = -1;
      P.out_kp[(long long)f * P.kp_cap + o] = kp;
    }
  }
  // ---- 3. rBRIEF: lane j computes descriptor byte j (8 tests) of one keypoint at a time
  uint32_t pat[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) pat[j] = g_pattern_s8[j * 32 + lane];
  for (int k = 0; k < ng; ++k) {
    if (k + 1 < ng) {
      stage_patch(k + 1);                                       // slot (k + 1) & 1 was read by keypoint k - 1: all lanes are past it
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const uint32_t e = __shfl_sync(0xFFFFFFFFu, e_mine, k);
    const float ak = __shfl_sync(0xFFFFFFFFu, a, k), bk = __shfl_sync(0xFFFFFFFFu, b, k);
    const int ox = ((int)(e & 0xFFF) - kEdge) & 15;
    // No conversion instructions (I2F / F2I run on the quarter-rate XU pipe, which was 77 % busy and bound the kernel):
    //  * a pattern coordinate v is stored as the byte v + 32; PRMT puts it under the exponent of 2^23 (0x4B000000 | byte
    //    = 8388608 + byte as a float) and one exact subtraction leaves float(v);
    //  * cvRound of a steered coordinate t is t + 1.5 * 2^23 (the sum is rounded to an integer, half to even, like
    //    cvRound), whose bit pattern is 0x4B400000 + round(t): the row * pitch + column of the sample is computed
    //    straight from the two bit patterns in wrapping 32-bit arithmetic, the constants folded into `corner`.
    const uint32_t corner = ring + (uint32_t)((k & 1) * kPatchRows * kPatchPitch + kEdge * kPatchPitch + ox + kEdge) -
                            0x4B400000u * (uint32_t)(kPatchPitch + 1);
    auto sample = [&](float x, float y) -> uint32_t {          // I[cy + round(x*b + y*a)][cx + round(x*a - y*b)] of the blurred level
      const float rm = __fadd_rn(__fadd_rn(__fmul_rn(x, bk), __fmul_rn(y, ak)), 12582912.f);
      const float cm = __fadd_rn(__fsub_rn(__fmul_rn(x, ak), __fmul_rn(y, bk)), 12582912.f);
      uint32_t v;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(__float_as_uint(rm) * (uint32_t)kPatchPitch + __float_as_uint(cm) + corner));
      return v;
    };
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t pw = pat[j];
      const float x0 = __fsub_rn(__uint_as_float(__byte_perm(pw, 0x4B000000u, 0x7440)), 8388640.f),
                  y0 = __fsub_rn(__uint_as_float(__byte_perm(pw, 0x4B000000u, 0x7441)), 8388640.f),
                  x1 = __fsub_rn(__uint_as_float(__byte_perm(pw, 0x4B000000u, 0x7442)), 8388640.f),
                  y1 = __fsub_rn(__uint_as_float(__byte_perm(pw, 0x4B000000u, 0x7443)), 8388640.f);
      val |= (sample(x0, y0) < sample(x1, y1)) << j;
    }
    const int o = base + k;
    if (o < P.kp_cap) P.out_desc[((long long)f * P.kp_cap + o) * 32 + lane] = (uint8_t)val;
    __syncwarp();                                               // everybody is done with slot k & 1 before it is refilled
  }
}


// ------------------------------------------------------------------ per-frame steps after extraction
// Frame::UndistortKeyPoints (Frame.cc:835-861; cv::undistortPoints with R = I, P = K: five fixed-point
// iterations in double), ComputeStereoFromRGBD (:893-911) and AssignFeaturesToGrid (:224-237) on the keypoints
// the batch just produced.  One CTA per frame; the grid lists come out in the reference's order (cells x-major,
// ascending keypoint index inside a cell) through a count / scan / rank pass over shared memory.
static const int kGridCells = DRFE_FRAME_GRID_COLS * DRFE_FRAME_GRID_ROWS;
struct PostDev {
  drfe_frame_params prm;
  const float* depth; long long depth_rs, depth_fs;
  const uint16_t* depth16; float depth_factor;                 // raw sensor depth instead: (float)u16 * factor (Frame.cc:113-115)
  drfe_keypoint* keys_un; float* u_right; float* kp_depth;   // [B][kp_cap]
  uint16_t* grid_count;                                         // [B][kGridCells]
  uint16_t* grid_index;                                         // [B][kp_cap]
};

__device__ __forceinline__ void undistort_point(const drfe_frame_params& p, float u, float v, float& ou, float& ov) {
  const double fx = p.fx, fy = p.fy, cx = p.cx, cy = p.cy;
  const double k1 = p.dist[0], k2 = p.dist[1], p1 = p.dist[2], p2 = p.dist[3], k3 = p.dist[4];
  const double ifx = 1. / fx, ify = 1. / fy;
  double x = ((double)u - cx) * ifx, y = ((double)v - cy) * ify;
  const double x0 = x, y0 = y;
#pragma unroll 1
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
    const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
    const double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  ou = (float)(fx * x + cx);
  ov = (float)(fy * y + cy);
}

__global__ void __launch_bounds__(256) k_frame_post(const OrbDev* __restrict__ Pp, PostDev Q) {
  __shared__ __align__(4) unsigned short s_cnt[kGridCells];     // keypoints per cell, then exclusive offsets
  extern __shared__ unsigned short s_cell[];       // [kp_cap] cell of each keypoint (0xFFFF: outside the grid)
  __shared__ int s_warp[8];
  const OrbDev& P = *Pp;
  const int f = blockIdx.x, tid = threadIdx.x;
  const int n = min(P.out_cnt[f], P.kp_cap);
  const drfe_frame_params& prm = Q.prm;
  for (int c = tid; c < kGridCells; c += 256) s_cnt[c] = 0;
  __syncthreads();
  const float inv_w = __fdiv_rn((float)DRFE_FRAME_GRID_COLS, __fsub_rn(prm.max_x, prm.min_x));
  const float inv_h = __fdiv_rn((float)DRFE_FRAME_GRID_ROWS, __fsub_rn(prm.max_y, prm.min_y));
  const long long o = (long long)f * P.kp_cap;
  const float* depth = Q.depth + (long long)f * Q.depth_fs;
  for (int i = tid; i < n; i += 256) {
    const drfe_keypoint k = P.out_kp[o + i];
    drfe_keypoint ku = k;
    if (prm.dist[0] != 0.0f) undistort_point(prm, k.x, k.y, ku.x, ku.y);
    Q.keys_un[o + i] = ku;
    const long long dpos = (long long)(int)k.y * Q.depth_rs + (int)k.x;      // imDepth.at<float>(v, u): floats truncate
    const float d = Q.depth16 ? __fmul_rn((float)Q.depth16[(long long)f * Q.depth_fs + dpos], Q.depth_factor) : depth[dpos];
    Q.kp_depth[o + i] = d > 0 ? d : -1.f;
    Q.u_right[o + i] = d > 0 ? __fsub_rn(ku.x, __fdiv_rn(prm.bf, d)) : -1.f;
    const int px = (int)roundf(__fmul_rn(__fsub_rn(ku.x, prm.min_x), inv_w));   // PosInGrid (:816-825)
    const int py = (int)roundf(__fmul_rn(__fsub_rn(ku.y, prm.min_y), inv_h));
    unsigned short cell = 0xFFFF;
    if (px >= 0 && px < DRFE_FRAME_GRID_COLS && py >= 0 && py < DRFE_FRAME_GRID_ROWS) {
      cell = (unsigned short)(px * DRFE_FRAME_GRID_ROWS + py);
      atomicAdd(reinterpret_cast<unsigned int*>(s_cnt) + (cell >> 1), (cell & 1) ? 0x10000u : 1u);   // counts < 65536: halves never carry
    }
    s_cell[i] = cell;
  }
  __syncthreads();
  uint16_t* gc = Q.grid_count + (long long)f * kGridCells;
  // counts out, then exclusive scan of the 3072 counts: 12 cells per thread
  int loc[12], sum = 0;
#pragma unroll
  for (int k = 0; k < 12; ++k) { loc[k] = s_cnt[tid * 12 + k]; gc[tid * 12 + k] = (uint16_t)loc[k]; sum += loc[k]; }
  int inc = sum;
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  int base = inc - sum;
  for (int w = 0; w < wid; ++w) base += s_warp[w];
#pragma unroll
  for (int k = 0; k < 12; ++k) { s_cnt[tid * 12 + k] = (unsigned short)base; base += loc[k]; }
  __syncthreads();
  // stable placement: rank of keypoint i inside its cell = number of earlier keypoints of the same cell
  uint16_t* gi = Q.grid_index + o;
  for (int i = tid; i < n; i += 256) {
    const unsigned short cell = s_cell[i];
    if (cell == 0xFFFF) continue;
    int rank = 0;
    for (int j = 0; j < i; ++j) rank += (s_cell[j] == cell);
    gi[s_cnt[cell] + rank] = (uint16_t)i;
  }
}

// ------------------------------------------------------------------ descriptor search by projection
// The data-parallel core of ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) (ORBmatcher.cc:69-116) on
// the device-resident results of drfe_orb_frame_post: for every query (a projected map point) the keypoints
// Frame::GetFeaturesInArea returns (Frame.cc:730-779: grid cells x-major, push_back order inside a cell, level and
// |dx|,|dy| < r filters), minus occupied ones and those failing the right-coordinate test, ranked by
// ORBmatcher::DescriptorDistance (:1712-1728) into best / second best with the reference's strict '<' updates.
// The ratio test and the in-order assignment F.mvpMapPoints[bestIdx] = pMP stay with the caller.
// One CTA per frame: the cell offsets (prefix sum of the cell counts) go to shared memory, then a thread per query.
struct SearchDev {
  const drfe_proj_query* q; const uint8_t* qdesc; const uint8_t* occupied; const int* nq; drfe_proj_match* out; int qcap;
};

// exclusive prefix sum of the 3072 mGrid cell counts of one frame into s_off[0 .. 3072] (all 256 threads of the CTA)
__device__ __forceinline__ void grid_offsets(const uint16_t* __restrict__ gc, unsigned short* s_off, int* s_part) {
  const int tid = threadIdx.x;                 // the first 256 threads work, every thread of the CTA takes the barriers
  constexpr int PER = kGridCells / 256;
  int loc[PER], sum = 0;
  if (tid < 256) {
#pragma unroll
    for (int k = 0; k < PER; ++k) { loc[k] = sum; sum += gc[tid * PER + k]; }
    s_part[tid] = sum;
  }
  __syncthreads();
  if (tid == 0) { int run = 0; for (int i = 0; i < 256; ++i) { const int t = s_part[i]; s_part[i] = run; run += t; } s_off[kGridCells] = (unsigned short)run; }
  __syncthreads();
  if (tid < 256) {
#pragma unroll
    for (int k = 0; k < PER; ++k) s_off[tid * PER + k] = (unsigned short)(s_part[tid] + loc[k]);
  }
  __syncthreads();
}

// Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) (Frame.cc:730-779): visit(idx, kp) for every index the reference
// pushes into vIndices, in its order (cells x-major, push_back order inside a cell)
template <class Visit>
__device__ __forceinline__ void features_in_area(const drfe_frame_params& prm, const unsigned short* s_off, const uint16_t* __restrict__ gi,
                                                 const drfe_keypoint* __restrict__ ku, float x, float y, float r, int min_level, int max_level,
                                                 Visit&& visit) {
  const float inv_w = __fdiv_rn((float)DRFE_FRAME_GRID_COLS, __fsub_rn(prm.max_x, prm.min_x));
  const float inv_h = __fdiv_rn((float)DRFE_FRAME_GRID_ROWS, __fsub_rn(prm.max_y, prm.min_y));
  const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, prm.min_x), r), inv_w)));
  const int cx1 = min(DRFE_FRAME_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, prm.min_x), r), inv_w)));
  const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, prm.min_y), r), inv_h)));
  const int cy1 = min(DRFE_FRAME_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, prm.min_y), r), inv_h)));
  if (!(cx0 < DRFE_FRAME_GRID_COLS && cx1 >= 0 && cy0 < DRFE_FRAME_GRID_ROWS && cy1 >= 0)) return;
  const bool check_levels = (min_level > 0) || (max_level >= 0);
  for (int ix = cx0; ix <= cx1; ++ix)
    for (int iy = cy0; iy <= cy1; ++iy) {
      const int cell = ix * DRFE_FRAME_GRID_ROWS + iy;
      for (int j = s_off[cell]; j < s_off[cell + 1]; ++j) {
        const int idx = gi[j];
        const drfe_keypoint kp = ku[idx];
        if (check_levels) {
          if (kp.octave < min_level) continue;
          if (max_level >= 0 && kp.octave > max_level) continue;
        }
        if (!(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) continue;
        visit(idx, kp);
      }
    }
}

__device__ __forceinline__ int descriptor_distance(const uint32_t (&d)[8], const uint32_t* __restrict__ kd) {   // ORBmatcher.cc:1712-1728
  const uint4 a = *reinterpret_cast<const uint4*>(kd), b = *reinterpret_cast<const uint4*>(kd + 4);
  return __popc(d[0] ^ a.x) + __popc(d[1] ^ a.y) + __popc(d[2] ^ a.z) + __popc(d[3] ^ a.w) + __popc(d[4] ^ b.x) + __popc(d[5] ^ b.y) +
         __popc(d[6] ^ b.z) + __popc(d[7] ^ b.w);
}

__global__ void __launch_bounds__(256) k_search_projection(const OrbDev* __restrict__ Pp, PostDev Q, SearchDev S) {
  __shared__ unsigned short s_off[kGridCells + 1];
  __shared__ int s_part[256];
  const OrbDev& P = *Pp;
  const int f = blockIdx.x, tid = threadIdx.x;
  grid_offsets(Q.grid_count + (long long)f * kGridCells, s_off, s_part);
  const drfe_keypoint* ku = Q.keys_un + (long long)f * P.kp_cap;
  const float* ur = Q.u_right + (long long)f * P.kp_cap;
  const uint16_t* gi = Q.grid_index + (long long)f * P.kp_cap;
  const uint32_t* desc = reinterpret_cast<const uint32_t*>(P.out_desc + (long long)f * P.kp_cap * 32);
  const uint8_t* occ = S.occupied ? S.occupied + (long long)f * P.kp_cap : nullptr;
  const int nq = min(S.nq[f], S.qcap);
  for (int qi = tid; qi < nq; qi += 256) {
    const drfe_proj_query q = S.q[(long long)f * S.qcap + qi];
    const uint32_t* qd = reinterpret_cast<const uint32_t*>(S.qdesc + ((long long)f * S.qcap + qi) * 32);
    uint32_t d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = qd[k];
    drfe_proj_match m;
    m.best_dist = 256; m.best_idx = -1; m.best_level = -1; m.best_dist2 = 256; m.best_level2 = -1;
    features_in_area(Q.prm, s_off, gi, ku, q.x, q.y, q.r, q.min_level, q.max_level, [&](int idx, const drfe_keypoint& kp) {
      // ORBmatcher.cc:88-101
      if (occ && occ[idx]) return;
      const float u = ur[idx];
      if (u > 0.f && fabsf(__fsub_rn(q.xr, u)) > q.r) return;
      const int dist = descriptor_distance(d, desc + idx * 8);
      if (dist < m.best_dist) { m.best_dist2 = m.best_dist; m.best_dist = dist; m.best_level2 = m.best_level; m.best_level = kp.octave; m.best_idx = idx; }
      else if (dist < m.best_dist2) { m.best_level2 = kp.octave; m.best_dist2 = dist; }
    });
    S.out[(long long)f * S.qcap + qi] = m;
  }
}

// ------------------------------------------------------------------ frame-to-frame search by projection
// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (ORBmatcher.cc:1396-1535), whole.
// One CTA per current frame.  The reference visits the last frame's points in order and a match made by point j (if its
// map point has observations) hides that keypoint from every later point (:1471-1473).  Here all points search in
// parallel against owner[idx] = the lowest-numbered observed point that chose keypoint idx in the previous sweep (-1: held
// on entry), skipping idx when owner[idx] < i, until a sweep changes no choice.  That fixed point satisfies the
// reference's recurrence point by point, and the recurrence has one solution (induction on i), so it is the reference's
// result; point i is final after at most i + 1 sweeps, in practice after 2-3.  Then rotation histogram / three maxima.
// threads per CTA (= per frame): one point per thread for up to 1024 last-frame points; 64 registers per thread
constexpr int kTrackThreads = 1024;
struct TrackDev {
  const drfe_track_params* tp; const drfe_last_point* pts; const uint8_t* pdesc; const uint8_t* occupied; const int* np;
  int32_t* match_key; int32_t* match_dist; int32_t* key_point; int* nmatches; int* sweeps; int pcap;
};
__global__ void __launch_bounds__(kTrackThreads) k_search_last_frame(const OrbDev* __restrict__ Pp, PostDev Q, TrackDev S) {
  extern __shared__ int s_dyn[];
  __shared__ unsigned short s_off[kGridCells + 1];
  __shared__ int s_part[256];
  __shared__ int s_hist[DRFE_HISTO_LENGTH];
  __shared__ int s_keep[3];
  __shared__ int s_cnt[2];
  const OrbDev& P = *Pp;
  const int f = blockIdx.x, tid = threadIdx.x, cap = P.kp_cap;
  int* owner = s_dyn;                 // [cap]
  int* choice = s_dyn + cap;          // [pcap] matched keypoint of point i or -1
  int* cdist = choice + S.pcap;       // [pcap] bestDist
  int* bidx = cdist + S.pcap;         // [pcap] the keypoint that gave bestDist (also when bestDist > TH_HIGH), -1: nothing compared
  unsigned char* skipped = reinterpret_cast<unsigned char*>(bidx + S.pcap);   // [pcap] the last search of point i skipped a taken keypoint
  grid_offsets(Q.grid_count + (long long)f * kGridCells, s_off, s_part);
  const drfe_frame_params prm = Q.prm;
  const drfe_track_params tp = S.tp[f];
  const drfe_keypoint* ku = Q.keys_un + (long long)f * cap;
  const float* ur = Q.u_right + (long long)f * cap;
  const uint16_t* gi = Q.grid_index + (long long)f * cap;
  const uint32_t* desc = reinterpret_cast<const uint32_t*>(P.out_desc + (long long)f * cap * 32);
  const uint8_t* occ = S.occupied ? S.occupied + (long long)f * cap : nullptr;
  const drfe_last_point* pts = S.pts + (long long)f * S.pcap;
  const int np = min(S.np[f], S.pcap);
  const int nkeys = P.out_cnt[f];
  for (int i = tid; i < np; i += kTrackThreads) choice[i] = -1;
  int sweeps = 0;
  for (;;) {
    // owner[] from the previous sweep's choices
    for (int k = tid; k < cap; k += kTrackThreads) owner[k] = (occ && k < nkeys && occ[k]) ? -1 : 0x7fffffff;
    __syncthreads();
    for (int i = tid; i < np; i += kTrackThreads)
      if (choice[i] >= 0 && (pts[i].flags & DRFE_LP_OBSERVED)) atomicMin(&owner[choice[i]], i);
    __syncthreads();
    int changed = 0;
    for (int i = tid; i < np; i += kTrackThreads) {
      // a point whose last search skipped nothing (so it saw its unconstrained first minimum) keeps that result as long as
      // that keypoint is not taken by a lower-numbered point: hiding other candidates cannot change a first minimum
      if (sweeps > 0 && !skipped[i] && (bidx[i] < 0 || owner[bidx[i]] >= i)) continue;
      const drfe_last_point lp = pts[i];
      int best = 256, best_idx = -1;
      bool skip = false;
      if (lp.flags & DRFE_LP_VALID) {
        // :1426-1442  x3Dc = Rcw*x3Dw + tcw: cv::gemm's 3x3 float path (products and sums in float, left to right, then + c)
        const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(tp.Tcw[0], lp.X), __fmul_rn(tp.Tcw[1], lp.Y)), __fmul_rn(tp.Tcw[2], lp.Z)), tp.Tcw[3]);
        const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(tp.Tcw[4], lp.X), __fmul_rn(tp.Tcw[5], lp.Y)), __fmul_rn(tp.Tcw[6], lp.Z)), tp.Tcw[7]);
        const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(tp.Tcw[8], lp.X), __fmul_rn(tp.Tcw[9], lp.Y)), __fmul_rn(tp.Tcw[10], lp.Z)), tp.Tcw[11]);
        const float invzc = (float)(1.0 / (double)zc);
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(prm.fx, xc), invzc), prm.cx);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(prm.fy, yc), invzc), prm.cy);
        const bool in = !(invzc < 0.f) && !(u < prm.min_x || u > prm.max_x) && !(v < prm.min_y || v > prm.max_y) && u == u && v == v;
        if (in) {
          const int oct = lp.octave;
          const float radius = __fmul_rn(tp.th, P.lv[oct].scale);
          const int lo = tp.mode == 1 ? oct : tp.mode == 2 ? 0 : oct - 1;
          const int hi = tp.mode == 1 ? -1 : tp.mode == 2 ? oct : oct + 1;
          const float urp = __fsub_rn(u, __fmul_rn(prm.bf, invzc));
          const uint32_t* qd = reinterpret_cast<const uint32_t*>(S.pdesc + ((long long)f * S.pcap + i) * 32);
          uint32_t d[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) d[k] = qd[k];
          features_in_area(prm, s_off, gi, ku, u, v, radius, lo, hi, [&](int idx, const drfe_keypoint&) {
            if (owner[idx] < i) { skip = true; return; }                       // :1471-1473
            const float r2 = ur[idx];
            if (r2 > 0.f && fabsf(__fsub_rn(urp, r2)) > radius) return;        // :1475-1481
            const int dist = descriptor_distance(d, desc + idx * 8);
            if (dist < best) { best = dist; best_idx = idx; }                  // :1487-1491
          });
        }
      }
      const int c = best <= DRFE_TH_HIGH ? best_idx : -1;                      // :1494 (256 when nothing was compared)
      changed |= (c != choice[i]);
      choice[i] = c; cdist[i] = best; bidx[i] = best_idx; skipped[i] = skip;
    }
    ++sweeps;
    if (!__syncthreads_or(changed)) break;
  }
  // rotation consistency (:1499-1532)
  if (tid < DRFE_HISTO_LENGTH) s_hist[tid] = 0;
  if (tid < 2) s_cnt[tid] = 0;
  for (int k = tid; k < cap; k += kTrackThreads) owner[k] = -1;                          // becomes key_point
  __syncthreads();
  const float factor = 1.0f / DRFE_HISTO_LENGTH;
  auto rot_bin = [&](int i) {
    float rot = __fsub_rn(pts[i].angle, ku[choice[i]].angle);
    if (rot < 0.f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == DRFE_HISTO_LENGTH) bin = 0;
    return bin;
  };
  int mine = 0;
  for (int i = tid; i < np; i += kTrackThreads)
    if (choice[i] >= 0) {
      ++mine;
      if (tp.check_orientation) { const int b = rot_bin(i); if (b >= 0 && b < DRFE_HISTO_LENGTH) atomicAdd(&s_hist[b], 1); }
    }
  if (mine) atomicAdd(&s_cnt[0], mine);
  __syncthreads();
  if (tid == 0) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    if (tp.check_orientation) {                                               // ComputeThreeMaxima (:1666-1707)
      int max1 = 0, max2 = 0, max3 = 0;
      for (int i = 0; i < DRFE_HISTO_LENGTH; ++i) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
    }
    s_keep[0] = ind1; s_keep[1] = ind2; s_keep[2] = ind3;
  }
  __syncthreads();
  // the in-order assignment mvpMapPoints[bestIdx2] = pMP leaves the highest-numbered point that chose the keypoint
  for (int i = tid; i < np; i += kTrackThreads)
    if (choice[i] >= 0) atomicMax(&owner[choice[i]], i);
  __syncthreads();
  if (tp.check_orientation) {
    int removed = 0;
    for (int i = tid; i < np; i += kTrackThreads)
      if (choice[i] >= 0) {
        const int b = rot_bin(i);
        if (b != s_keep[0] && b != s_keep[1] && b != s_keep[2]) { owner[choice[i]] = -2; ++removed; }   // set to NULL, nmatches-- (:1527-1528)
      }
    if (removed) atomicAdd(&s_cnt[1], removed);
  }
  __syncthreads();
  if (S.key_point) for (int k = tid; k < cap; k += kTrackThreads) S.key_point[(long long)f * cap + k] = owner[k];
  for (int i = tid; i < S.pcap; i += kTrackThreads) {                          // slots past npoints[f]: no match
    S.match_key[(long long)f * S.pcap + i] = i < np ? choice[i] : -1;
    S.match_dist[(long long)f * S.pcap + i] = i < np ? cdist[i] : 256;
  }
  if (tid == 0) { S.nmatches[f] = s_cnt[0] - s_cnt[1]; S.sweeps[f] = sweeps; }
}

// cv::cvtColor(..., CV_RGB2GRAY and friends) of Tracking::GrabImageRGBD (Tracking.cc:194-207) in cv's fixed point: a thread
// converts 4 pixels (12 or 16 bytes in, one 32-bit word out)
template <int CH>
__global__ void __launch_bounds__(256) k_color_to_gray(const uint8_t* __restrict__ src, long long rs, long long fs, uint8_t* __restrict__ dst, int W, int H,
                                                       int cr, int cg, int cb, int shift, int swap_rb) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, f = blockIdx.z;
  if (x4 >= W) return;
  const uint8_t* p = src + f * fs + y * rs + (long long)x4 * CH;
  uint8_t* q = dst + ((long long)f * H + y) * W + x4;
  const int half = 1 << (shift - 1);
  uint32_t out = 0;
  const int n = min(4, W - x4);
  for (int i = 0; i < n; ++i) {
    const int c0 = p[i * CH], c1 = p[i * CH + 1], c2 = p[i * CH + 2];
    const int r = swap_rb ? c2 : c0, b = swap_rb ? c0 : c2;
    out |= (uint32_t)((r * cr + c1 * cg + b * cb + half) >> shift) << (8 * i);
  }
  if (n == 4 && (W & 3) == 0) *reinterpret_cast<uint32_t*>(q) = out;
  else for (int i = 0; i < n; ++i) q[i] = (uint8_t)(out >> (8 * i));
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (ORBmatcher.cc:46-130) whole: k_search_projection's
// search + ratio test + in-order assignment, the latter by parallel sweeps to the fixed point as in k_search_last_frame.
// Here a hidden candidate can change the second best and with it the ratio test, so every sweep searches every point again.
struct LocalDev {
  const drfe_proj_query* q; const uint8_t* qdesc; const uint8_t* qflags; const uint8_t* occupied; const int* nq;
  drfe_proj_match* out; int32_t* assigned; int32_t* key_point; int* nmatches; int qcap; float nnratio;
};
__global__ void __launch_bounds__(kTrackThreads) k_search_local_points(const OrbDev* __restrict__ Pp, PostDev Q, LocalDev S) {
  extern __shared__ int s_dyn[];
  __shared__ unsigned short s_off[kGridCells + 1];
  __shared__ int s_part[256];
  __shared__ int s_cnt;
  const OrbDev& P = *Pp;
  const int f = blockIdx.x, tid = threadIdx.x, cap = P.kp_cap;
  int* owner = s_dyn;            // [cap]
  int* choice = s_dyn + cap;     // [qcap]
  grid_offsets(Q.grid_count + (long long)f * kGridCells, s_off, s_part);
  const drfe_frame_params prm = Q.prm;
  const drfe_keypoint* ku = Q.keys_un + (long long)f * cap;
  const float* ur = Q.u_right + (long long)f * cap;
  const uint16_t* gi = Q.grid_index + (long long)f * cap;
  const uint32_t* desc = reinterpret_cast<const uint32_t*>(P.out_desc + (long long)f * cap * 32);
  const uint8_t* occ = S.occupied ? S.occupied + (long long)f * cap : nullptr;
  const uint8_t* qfl = S.qflags + (long long)f * S.qcap;
  const int nq = min(S.nq[f], S.qcap), nkeys = P.out_cnt[f];
  for (int i = tid; i < nq; i += kTrackThreads) choice[i] = -1;
  if (tid == 0) s_cnt = 0;
  for (;;) {
    for (int k = tid; k < cap; k += kTrackThreads) owner[k] = (occ && k < nkeys && occ[k]) ? -1 : 0x7fffffff;
    __syncthreads();
    for (int i = tid; i < nq; i += kTrackThreads)
      if (choice[i] >= 0 && (qfl[i] & DRFE_LP_OBSERVED)) atomicMin(&owner[choice[i]], i);
    __syncthreads();
    int changed = 0;
    for (int i = tid; i < nq; i += kTrackThreads) {
      drfe_proj_match m;
      m.best_dist = 256; m.best_idx = -1; m.best_level = -1; m.best_dist2 = 256; m.best_level2 = -1;
      int c = -1;
      if (qfl[i] & DRFE_LP_VALID) {
        const drfe_proj_query q = S.q[(long long)f * S.qcap + i];
        const uint32_t* qd = reinterpret_cast<const uint32_t*>(S.qdesc + ((long long)f * S.qcap + i) * 32);
        uint32_t d[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = qd[k];
        features_in_area(prm, s_off, gi, ku, q.x, q.y, q.r, q.min_level, q.max_level, [&](int idx, const drfe_keypoint& kp) {
          if (owner[idx] < i) return;                                                          // :88-90
          const float u = ur[idx];
          if (u > 0.f && fabsf(__fsub_rn(q.xr, u)) > q.r) return;                              // :92-97
          const int dist = descriptor_distance(d, desc + idx * 8);
          if (dist < m.best_dist) { m.best_dist2 = m.best_dist; m.best_dist = dist; m.best_level2 = m.best_level; m.best_level = kp.octave; m.best_idx = idx; }
          else if (dist < m.best_dist2) { m.best_level2 = kp.octave; m.best_dist2 = dist; }
        });
        if (m.best_dist <= DRFE_TH_HIGH &&                                                     // :117-126
            !(m.best_level == m.best_level2 && (float)m.best_dist > __fmul_rn(S.nnratio, (float)m.best_dist2)))
          c = m.best_idx;
      }
      changed |= (c != choice[i]);
      choice[i] = c;
      S.out[(long long)f * S.qcap + i] = m;
    }
    if (!__syncthreads_or(changed)) break;
  }
  for (int k = tid; k < cap; k += kTrackThreads) owner[k] = -1;
  __syncthreads();
  int mine = 0;
  for (int i = tid; i < S.qcap; i += kTrackThreads) {
    if (i >= nq) {                                                        // slots past nqueries[f]: the empty record
      drfe_proj_match m;
      m.best_dist = 256; m.best_idx = -1; m.best_level = -1; m.best_dist2 = 256; m.best_level2 = -1;
      S.out[(long long)f * S.qcap + i] = m;
      S.assigned[(long long)f * S.qcap + i] = -1;
      continue;
    }
    S.assigned[(long long)f * S.qcap + i] = choice[i];
    if (choice[i] >= 0) { ++mine; atomicMax(&owner[choice[i]], i); }     // the last writer of F.mvpMapPoints[bestIdx] stays
  }
  if (mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  for (int k = tid; k < cap; k += kTrackThreads) S.key_point[(long long)f * cap + k] = owner[k];
  if (tid == 0) S.nmatches[f] = s_cnt;
}

__global__ void k_zero_counts(const OrbDev* __restrict__ Pp, int f0, int nframes) {
  DRFE_GRID_DEP();
  const OrbDev& P = *Pp;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nframes * P.nlevels) { P.cand_cnt[f0 * P.nlevels + i] = 0; P.lkp_cnt[f0 * P.nlevels + i] = 0; }
  if (i < nframes) P.out_cnt[f0 + i] = 0;
}

}  // namespace drfe

// ====================================================================== host side
using namespace drfe;

static const int kHiPrioMinFrames = 128;   // launches of at least this many frames run on the handle's high-priority stream
struct drfe_orb {
  int device = 0, width = 0, height = 0, max_batch = 0;
  drfe_orb_params prm{};
  double scaleFactorD = 1.2;
  std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
  std::vector<int> per_level;
  OrbDev hd{};                 // host copy of the device descriptor
  OrbDev* dd = nullptr;        // device copy
  cudaStream_t stream = nullptr;
  uint8_t* d_gray = nullptr;   // staging for host inputs [B][H][W]
  cudaEvent_t ev_shared = nullptr;   // drfe_orb_frame_post_shared_depth
  static const int kMaxSplit = 4;
  int split = 1;                     // parts a long batch's kernel chains are cut into (orb_launch)
  cudaStream_t aux[kMaxSplit - 1] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxSplit - 1] = {nullptr, nullptr, nullptr};
  cudaStream_t hi = nullptr;             // one priority level above `stream`, for launches of many frames
  cudaStream_t blur_stream = nullptr;    // small launches: the blur beside FAST and the quadtree
  cudaEvent_t ev_blur_fork = nullptr, ev_blur_join = nullptr;
  cudaStream_t fast0_stream = nullptr;   // small launches: FAST on level 0 beside the resizes of levels 1..
  cudaEvent_t ev_l0 = nullptr, ev_fast0 = nullptr;
  int nstrips_l0 = 0;
  cudaEvent_t ev_hi = nullptr;
  int hi_min_frames = 0;
  uint8_t* d_color = nullptr; size_t color_bytes = 0; bool gray_valid = false;   // drfe_orb_enqueue_color
  void* d_rtab = nullptr; void* d_strips = nullptr;
  FastMaps fast_maps{};
  int nstrips = 0, blur_blocks = 0, max_node_cap = 0, max_lkp = 0;
  size_t fast_smem = 0, quad_smem = 0, pyr_smem = 0;
  int last_frames = 0;
  bool pending = false;
  bool post_valid = false;       // drfe_orb_frame_post ran on the batch that is on the device now
  StageTimer timer;
  ChunkPipe pipe;
  PostDev post{};                // buffers of drfe_orb_frame_post, allocated on first use
  float* d_post_depth = nullptr;
  drfe_proj_query* d_sq = nullptr; uint8_t* d_sdesc = nullptr; uint8_t* d_socc = nullptr; int* d_snq = nullptr;   // drfe_orb_search_by_projection
  drfe_track_params* d_ttp = nullptr; drfe_last_point* d_tpts = nullptr; uint8_t* d_tdesc = nullptr; int32_t* d_tout = nullptr; int32_t* d_tkey = nullptr; int* d_tcnt = nullptr; int track_pcap = 0;   // drfe_orb_search_last_frame
  drfe_proj_match* d_sout = nullptr; int search_qcap = 0;
  int* batch_counts = nullptr;   // host destination of the running batch call
  int batch_cap = 0;
  std::vector<void*> allocs;
};

static inline int cv_round_f(float v) { return (int)lrintf(v); }

static void make_resize_tab(int ssize, int dsize, std::vector<uint2>& out) {
  // OpenCV resize.cpp (INTER_LINEAR, 8U): scale = 1/(dsize/ssize); fx = (dx+0.5)*scale-0.5 (float)
  const double scale = 1.0 / ((double)dsize / ssize);
  for (int d = 0; d < dsize; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    const int c0 = cv_round_f((1.f - f) * 2048.f), c1 = cv_round_f(f * 2048.f);
    const int s1 = std::min(s + 1, ssize - 1);
    out.push_back(make_uint2((unsigned)s | ((unsigned)s1 << 16), (unsigned)c0 | ((unsigned)c1 << 16)));
  }
}

template <typename T>
static int dev_alloc(drfe_orb* h, T** p, size_t count) {
  void* q = nullptr;
  DRFE_CUDA(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  h->allocs.push_back(q);
  *p = (T*)q;
  return DRFE_OK;
}

static int orb_build(drfe_orb* h) {
  const drfe_orb_params& pr = h->prm;
  const int nl = pr.nlevels;
  // ---- ORBextractor::ORBextractor (ORBextractor.cc:410-446)
  h->scaleFactorD = (double)pr.scale_factor;
  h->scale.assign(nl, 1.f); h->sigma2.assign(nl, 1.f);
  for (int i = 1; i < nl; ++i) {
    h->scale[i] = (float)(h->scale[i - 1] * h->scaleFactorD);
    h->sigma2[i] = h->scale[i] * h->scale[i];
  }
  h->inv_scale.resize(nl); h->inv_sigma2.resize(nl);
  for (int i = 0; i < nl; ++i) { h->inv_scale[i] = 1.0f / h->scale[i]; h->inv_sigma2[i] = 1.0f / h->sigma2[i]; }
  h->per_level.resize(nl);
  {
    const float factor = (float)(1.0f / h->scaleFactorD);
    float want = pr.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; ++l) { h->per_level[l] = cv_round_f(want); sum += h->per_level[l]; want *= factor; }
    h->per_level[nl - 1] = std::max(pr.nfeatures - sum, 0);
  }
  int umax[16];
  {
    const int vmax = (int)floor(kHalfPatch * sqrt(2.f) / 2 + 1), vmin = (int)ceil(kHalfPatch * sqrt(2.f) / 2);
    for (int v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt((double)kHalfPatch * kHalfPatch - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0; ++v0;
    }
  }
  // ---- per-level geometry (ComputePyramid :1111-1113, ComputeKeyPointsOctTree :773-787)
  OrbDev& D = h->hd;
  memset(&D, 0, sizeof(D));
  D.nlevels = nl; D.B = h->max_batch; D.ini_th = pr.ini_th_fast; D.min_th = pr.min_th_fast;
  std::vector<uint2> rtab;
  std::vector<PyrTile> ptiles;
  std::vector<PyrCol> pcols;
  std::vector<uint2> ytab2;
  bool stream_ok = true;
  int max_pyr_src = 0, max_pyr_nsy = 0;
  std::vector<StripDev> strips;
  long long img_total = 0, blur_total = 0, cand_total = 0;
  int lkp_total = 0, kp_cap = 0, max_strip_rows = 0, max_strip_tp = 0;
  for (int l = 0; l < nl; ++l) {
    LevelDev& L = D.lv[l];
    L.w = cv_round_f((float)h->width * h->inv_scale[l]);
    L.h = cv_round_f((float)h->height * h->inv_scale[l]);
    L.pitch = (kXOff + L.w + 32 + 127) / 128 * 128;
    L.rows = L.h + 2 * kEdge;
    L.img_fstride = (long long)L.pitch * (L.rows + 1);
    L.img_off = img_total;
    img_total += L.img_fstride * h->max_batch;
    L.bpitch = (L.w + 127) / 128 * 128;
    L.blur_fstride = (long long)L.bpitch * L.h;
    L.blur_off = blur_total;
    blur_total += L.blur_fstride * h->max_batch;
    L.regW = L.w - 2 * (kEdge - 3); L.regH = L.h - 2 * (kEdge - 3);
    if (L.regW < 30 || L.regH < 30) { set_error("level %d (%dx%d) is too small for the 30 px FAST grid", l, L.w, L.h); return DRFE_ERR_ARG; }
    const float width = (float)L.regW, height = (float)L.regH;
    L.nCols = (int)(width / 30.f); L.nRows = (int)(height / 30.f);
    L.wCell = (int)ceilf(width / L.nCols); L.hCell = (int)ceilf(height / L.nRows);
    L.nfeat = h->per_level[l];
    L.nIni = (int)roundf(width / height);
    if (L.nIni < 1 || L.nIni > 4) { set_error("unsupported aspect ratio (quadtree roots = %d)", L.nIni); return DRFE_ERR_ARG; }
    L.hX = width / L.nIni;
    for (int i = 0; i <= L.nIni; ++i) L.root_x[i] = (int)(L.hX * (float)i);
    // FAST candidates a level can hold: the most that per-cell strict 3x3 non-maximum suppression can leave, i.e.
    // ceil(w/2) * ceil(h/2) per cell interior of w x h pixels — no image can overflow the arena.  (Measured on the
    // synthetic sequences: at most one per 67 px on level 0 and one per 20 px on level 7.)
    {
      const int iw = L.w - 2 * kEdge, ih = L.h - 2 * kEdge;
      long long cap = 0;
      for (int i = 0; i < L.nRows; ++i) {
        const int hc = std::min((i + 1) * L.hCell, ih) - i * L.hCell;
        if (hc <= 0) continue;
        for (int j = 0; j < L.nCols; ++j) {
          const int wc = std::min((j + 1) * L.wCell, iw) - j * L.wCell;
          if (wc > 0) cap += (long long)((wc + 1) / 2) * ((hc + 1) / 2);
        }
      }
      L.cand_cap = (int)std::min<long long>(1 << 21, std::max<long long>(1024, cap));
    }
    L.cand_off = cand_total; cand_total += L.cand_cap;
    L.node_cap = std::max(L.nfeat + 3, 4 * L.nIni) + 1;
    if (L.node_cap > 0x3FFF) { set_error("nfeatures too large"); return DRFE_ERR_ARG; }
    h->max_node_cap = std::max(h->max_node_cap, L.node_cap);
    L.kp_off = lkp_total; lkp_total += L.node_cap;
    kp_cap += L.node_cap;
    L.scale = h->scale[l];
    L.size = (float)(int)(31 * h->scale[l]);
    if (l > 0) {
      L.xtab_off = (long long)rtab.size(); make_resize_tab(D.lv[l - 1].w, L.w, rtab);
      L.ytab_off = (long long)rtab.size(); make_resize_tab(D.lv[l - 1].h, L.h, rtab);
      // k_pyr_resize tiles and their source windows
      const int groups = (L.w + 40 + 3) / 4;
      L.ptile_off = (int)ptiles.size();
      L.ptile_gx = (groups + 31) / 32;
      auto refl = [](int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * (n - 1) - i; return i; };
      for (int by0 = 0; by0 < L.rows; by0 += kPyrTH)
        for (int g0 = 0; g0 < groups; g0 += kPyrTW / 4) {
          int sx0 = 1 << 30, sx1 = -1, sy0 = 1 << 30, sy1 = -1;
          for (int g = g0; g < std::min(g0 + kPyrTW / 4, groups); ++g)
            for (int k = 0; k < 4; ++k) {
              const uint2 e = rtab[L.xtab_off + refl(4 * g - 20 + k, L.w)];
              sx0 = std::min(sx0, (int)(e.x & 0xFFFF)); sx1 = std::max(sx1, (int)(e.x >> 16));
            }
          for (int by = by0; by < std::min(by0 + kPyrTH, L.rows); ++by) {
            const uint2 e = rtab[L.ytab_off + refl(by - kEdge, L.h)];
            sy0 = std::min(sy0, (int)(e.x & 0xFFFF)); sy1 = std::max(sy1, (int)(e.x >> 16));
          }
          sx0 &= ~3;
          const int sw4 = (sx1 - sx0) / 4 + 1, nsy = sy1 - sy0 + 1;
          ptiles.push_back(PyrTile{(short)g0, (short)by0, (short)sx0, (short)sw4, (short)sy0, (short)nsy, 0, 0});
          max_pyr_src = std::max(max_pyr_src, sw4 * 4 * nsy);
          max_pyr_nsy = std::max(max_pyr_nsy, nsy);
        }
      L.ptile_cnt = (int)ptiles.size() - L.ptile_off;
      // k_pyr_stream records: per 4-column group the byte window base, the pixel offsets inside it and the
      // coefficient pairs; per destination row the first source row and the two vertical coefficients
      L.pcol_off = (int)pcols.size(); L.pcol_groups = groups;
      for (int g = 0; g < groups; ++g) {
        PyrCol pc{};
        uint32_t i0[4];
        uint32_t base = 0xFFFFFFFFu;
        for (int k = 0; k < 4; ++k) {
          const uint2 e = rtab[L.xtab_off + refl(4 * g - 20 + k, L.w)];
          i0[k] = e.x & 0xFFFF;
          const uint32_t i1 = e.x >> 16, c1 = e.y >> 16;
          if (i1 != i0[k] + 1 && c1 != 0) stream_ok = false;
          pc.coef[k] = e.y;
          base = std::min(base, i0[k]);
        }
        pc.base = base;
        for (int k = 0; k < 4; ++k) {
          if (i0[k] - base > 4) stream_ok = false;
          pc.sh |= (8u * (i0[k] - base)) << (8 * k);
        }
        pcols.push_back(pc);
      }
      L.ytab2_off = (int)ytab2.size();
      for (int y = 0; y < L.h; ++y) {
        const uint2 e = rtab[L.ytab_off + y];
        const uint32_t r0 = e.x & 0xFFFF, r1 = e.x >> 16, b0 = e.y & 0xFFFF, b1 = e.y >> 16;
        if (r1 != r0 + 1 && b1 != 0) stream_ok = false;
        if (y > 0 && r0 < (rtab[L.ytab_off + y - 1].x & 0xFFFF)) stream_ok = false;
        ytab2.push_back(make_uint2(r0 | (b0 << 16), b1));
      }
    }
    // FAST strips: the interior rows [19 + i*hCell, min(19 + (i+1)*hCell, h-19)) of cell row i.
    // Cell (i, j) of the reference's grid (:789-829) evaluates FAST exactly on the pixels
    // [19 + j*wCell, ..) x [19 + i*hCell, ..) clipped to [19, w-19) x [19, h-19) (App. A.3b).
    {
      const int iw = L.w - 2 * kEdge;
      const int seg_cells = iw > kFastSplitWidth ? kFastSegCells : L.nCols;
      if (seg_cells > kMaxStripCells) { set_error("image too wide (%d FAST cell columns)", L.nCols); return DRFE_ERR_ARG; }
      L.wcell_magic = (uint32_t)(((1ull << 32) + L.wCell - 1) / L.wCell);
      for (int i = 0; i < L.nRows; ++i) {
        const int y0 = kEdge + i * L.hCell, y1 = std::min(y0 + L.hCell, L.h - kEdge);
        if (y1 <= y0) continue;
        for (int j0 = 0; j0 < L.nCols; j0 += seg_cells) {
          const int x0 = j0 * L.wCell, x1 = std::min((j0 + seg_cells) * L.wCell, iw);
          if (x1 <= x0) continue;
          const unsigned quads = (unsigned)((x1 - x0 + 3) / 4), groups = (quads + 3) / 4;
          if (quads > 1023) { set_error("image too wide for the FAST strip kernel (%u quads per row)", quads); return DRFE_ERR_ARG; }
          strips.push_back(StripDev{(short)l, (short)y0, (short)(y1 - y0), (short)x0, (short)(x1 - x0), 0,
                                    (uint32_t)(((1ull << 32) + groups - 1) / groups)});
          max_strip_rows = std::max(max_strip_rows, y1 - y0);
          max_strip_tp = std::max(max_strip_tp, (std::min(x1 - x0 + 2 * kEdge, L.w - x0) + 15) / 16 * 16);
        }
      }
    }
    L.blur_nq = (L.w + 3) / 4;
    L.blur_magic = (uint32_t)(((1ull << 32) + L.blur_nq - 1) / L.blur_nq);
    D.blur_blk_off[l] = h->blur_blocks;
    h->blur_blocks += (L.blur_nq * ((L.h + kBlurRows - 1) / kBlurRows) + 255) / 256;
    D.blur_blk_off[l + 1] = h->blur_blocks;
    if (L.w > 4000 || L.h > 4000) { set_error("image too large (12-bit packed coordinates)"); return DRFE_ERR_ARG; }
  }
  h->nstrips = (int)strips.size();
  h->nstrips_l0 = 0;
  for (const StripDev& sd : strips) h->nstrips_l0 += sd.level == 0;
  D.cand_fstride = cand_total; D.lkp_fstride = lkp_total; D.kp_cap = kp_cap; h->max_lkp = 0;
  for (int l = 0; l < nl; ++l) h->max_lkp = std::max(h->max_lkp, D.lv[l].node_cap);
  D.fast_tp = max_strip_tp;                       // smem row pitch of a FAST strip (TMA rows are 16 B multiples)
  D.fast_rows = max_strip_rows;
  if (max_strip_rows > 63) { set_error("FAST strip of %d rows (queue entries hold 6 row bits)", max_strip_rows); return DRFE_ERR_ARG; }
  {
    // tile | score map (+2 guard rows) | per-quad masks | compass queue (u16 per quad) | list of kept maxima
    const size_t tp = D.fast_tp, rows = D.fast_rows, max_tasks = (tp / 4) * rows;
    size_t off = tp * (rows + 6) + tp * (rows + 2) + tp * 4;
    // queue (u16 per quad), then the list of kept maxima: max_tasks/2 entries, i.e. one maximum per
    // 8 pixels; denser input overflows loudly (DRFE_ERR_CAPACITY), like the per-level candidate arena
    off += (max_tasks * 2 + 15) / 16 * 16;
    D.fast_list_off = (int)off;
    D.fast_list_cap = (int)(max_tasks / 2);
    h->fast_smem = off + (size_t)D.fast_list_cap * 4;
  }
  if (h->fast_smem > 220 * 1024) { set_error("image too wide for the FAST strip kernel (%zu B of shared memory)", h->fast_smem); return DRFE_ERR_ARG; }
  const int NC = h->max_node_cap;
  h->quad_smem = (size_t)NC * (8 + 8 + 8 + 4 + 4 + 16 + 5 * 4 + 8 + 2) + 64 + 16 + (size_t)kQuadStage * 6;

  // ---- device memory
  const int B = h->max_batch;
  // The handle's stream has the default priority; launches of at least kHiPrioMinFrames frames are forked onto `hi`, a stream one
  // priority level up, and joined back.  The ORB chain is the longer of the two a frame batch needs (1.48 ms against 0.58 ms for
  // the planes at 256 frames): with the higher priority the block scheduler places its CTAs first and the plane kernels — shorter,
  // two of them latency-bound — run in what is left: 1.87 -> 1.83 ms per two-stream step (the reverse: 1.94 ms).  Small launches
  // stay on the default priority: there the chains are latency-bound, and a prioritised ORB chain that also launches dependents early
  // (DRFE_LAUNCH_PDL) keeps the plane kernels off the SMs (32 frames: 0.326 ms -> 0.370 ms; 64: 0.532 -> 0.613).
  // DRFE_ORB_PRIO / DRFE_CAPE_PRIO set the handle streams' own priority, DRFE_ORB_HI_MIN_FRAMES the threshold (0: never).
  { const char* e = getenv("DRFE_ORB_PRIO"); DRFE_CUDA(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, e ? atoi(e) : 0)); }
  if (!(getenv("DRFE_ORB_NO_BLUR_FORK") && getenv("DRFE_ORB_NO_BLUR_FORK")[0] == '1')) {
    DRFE_CUDA(cudaStreamCreateWithFlags(&h->blur_stream, cudaStreamNonBlocking));
    DRFE_CUDA(cudaEventCreateWithFlags(&h->ev_blur_fork, cudaEventDisableTiming));
    DRFE_CUDA(cudaEventCreateWithFlags(&h->ev_blur_join, cudaEventDisableTiming));
    DRFE_CUDA(cudaStreamCreateWithFlags(&h->fast0_stream, cudaStreamNonBlocking));
    DRFE_CUDA(cudaEventCreateWithFlags(&h->ev_l0, cudaEventDisableTiming));
    DRFE_CUDA(cudaEventCreateWithFlags(&h->ev_fast0, cudaEventDisableTiming));
  }
  {
    const char* e = getenv("DRFE_ORB_HI_MIN_FRAMES");
    h->hi_min_frames = e ? atoi(e) : kHiPrioMinFrames;
    if (h->hi_min_frames > 0) {
      DRFE_CUDA(cudaStreamCreateWithPriority(&h->hi, cudaStreamNonBlocking, -1));
      DRFE_CUDA(cudaEventCreateWithFlags(&h->ev_hi, cudaEventDisableTiming));
    }
  }
  {
    const char* e = getenv("DRFE_ORB_SPLIT");
    h->split = e ? std::min(std::max(atoi(e), 1), (int)drfe_orb::kMaxSplit) : 1;
    DRFE_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    for (int p = 0; p + 1 < h->split; ++p) {
      DRFE_CUDA(cudaStreamCreateWithFlags(&h->aux[p], cudaStreamNonBlocking));
      DRFE_CUDA(cudaEventCreateWithFlags(&h->ev_join[p], cudaEventDisableTiming));
    }
  }
  if (dev_alloc(h, &D.pyr, (size_t)img_total + 256)) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &D.blur, (size_t)blur_total + 256)) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &D.cand, (size_t)cand_total * B)) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &D.node_of, (size_t)cand_total * B)) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &D.cand_cnt, (size_t)B * nl)) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &D.lkp, (size_t)lkp_total * B)) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &D.lkp_cnt, (size_t)B * nl)) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &D.out_kp, (size_t)kp_cap * B)) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &D.out_desc, (size_t)kp_cap * B * 32)) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &D.out_cnt, (size_t)B)) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &D.status, 1)) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &h->d_gray, (size_t)B * h->width * h->height)) return DRFE_ERR_CUDA;
  uint2* d_rtab; StripDev* d_strips; PyrTile* d_ptiles;
  if (dev_alloc(h, &d_rtab, rtab.size())) return DRFE_ERR_CUDA;
  if (dev_alloc(h, &d_ptiles, ptiles.size())) return DRFE_ERR_CUDA;
  DRFE_CUDA(cudaMemcpy(d_ptiles, ptiles.data(), ptiles.size() * sizeof(PyrTile), cudaMemcpyHostToDevice));
  D.ptiles = d_ptiles;
  // DRFE_PYR_GENERIC=1 forces the generic tile kernel (kept for scale factors the streaming kernel cannot take;
  // the environment switch exists so that the tests can exercise it)
  if (stream_ok && !pcols.empty() && getenv("DRFE_PYR_GENERIC") == nullptr) {
    PyrCol* d_pcols; uint2* d_ytab2;
    if (dev_alloc(h, &d_pcols, pcols.size())) return DRFE_ERR_CUDA;
    if (dev_alloc(h, &d_ytab2, ytab2.size())) return DRFE_ERR_CUDA;
    DRFE_CUDA(cudaMemcpy(d_pcols, pcols.data(), pcols.size() * sizeof(PyrCol), cudaMemcpyHostToDevice));
    DRFE_CUDA(cudaMemcpy(d_ytab2, ytab2.data(), ytab2.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    D.pcols = d_pcols; D.ytab2 = d_ytab2;
  }
  D.pyr_src_bytes = (max_pyr_src + 15) / 16 * 16;
  h->pyr_smem = (size_t)D.pyr_src_bytes + (size_t)max_pyr_nsy * 32 * sizeof(uint2);
  if (dev_alloc(h, &d_strips, strips.size())) return DRFE_ERR_CUDA;
  DRFE_CUDA(cudaMemcpy(d_rtab, rtab.data(), rtab.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  DRFE_CUDA(cudaMemcpy(d_strips, strips.data(), strips.size() * sizeof(StripDev), cudaMemcpyHostToDevice));
  D.rtab = d_rtab; D.strips = d_strips;
  DRFE_CUDA(cudaMemset(D.pyr, 0, (size_t)img_total + 256));
  DRFE_CUDA(cudaMemset(D.status, 0, sizeof(int)));
  // tensor maps of the level images for k_fast_strips (32-bit elements: a box may be at most 256 elements wide)
  {
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    bool ok = getenv("DRFE_FAST_NO_TMA2D") == nullptr && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
              fn != nullptr && qres == cudaDriverEntryPointSuccess && (D.fast_tp & 15) == 0 && D.fast_tp / 4 <= 256;
    for (int l = 0; ok && l < nl; ++l) {
      const LevelDev& L = D.lv[l];
      const cuuint64_t dims[3] = {(cuuint64_t)(L.pitch / 4), (cuuint64_t)(L.rows + 1), (cuuint64_t)B};
      const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)L.img_fstride};
      const cuuint32_t box[3] = {(cuuint32_t)(D.fast_tp / 4), (cuuint32_t)(L.hCell + 6), 1u};
      const cuuint32_t estr[3] = {1u, 1u, 1u};
      ok = L.hCell + 6 <= 256 && L.hCell <= D.fast_rows &&
           ((EncodeTiled)fn)(&h->fast_maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, D.pyr + L.img_off, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    D.fast_tma2d = ok ? 1 : 0;
  }
  if (dev_alloc(h, &h->dd, 1)) return DRFE_ERR_CUDA;
  DRFE_CUDA(cudaMemcpy(h->dd, &D, sizeof(D), cudaMemcpyHostToDevice));
  DRFE_CUDA(cudaMemcpyToSymbol(c_umax, umax, sizeof(umax)));
  {
    std::vector<uint32_t> ps(256);
    for (int t = 0; t < 256; ++t)
      ps[(t & 7) * 32 + (t >> 3)] = (uint32_t)(h_pattern[4 * t] + 32) | ((uint32_t)(h_pattern[4 * t + 1] + 32) << 8) |
                                    ((uint32_t)(h_pattern[4 * t + 2] + 32) << 16) | ((uint32_t)(h_pattern[4 * t + 3] + 32) << 24);
    DRFE_CUDA(cudaMemcpyToSymbol(g_pattern_s8, ps.data(), sizeof(uint32_t) * 256));
  }
  DRFE_CUDA(raise_dyn_smem((k_fast_strips<256>), h->device, (size_t)(h->fast_smem)));
  DRFE_CUDA(raise_dyn_smem((k_quadtree<256>), h->device, (size_t)(h->quad_smem)));
  if (h->timer.create()) return DRFE_ERR_CUDA;
  return DRFE_OK;
}

int drfe::orb_batch_view(drfe_orb* h, OrbBatchView* v) {
  if (!h || !v) return DRFE_ERR_ARG;
  v->device = h->device; v->nframes = h->last_frames; v->cap = h->hd.kp_cap; v->pending = h->pending; v->stream = h->stream;
  v->desc = h->hd.out_desc; v->cnt = h->hd.out_cnt; v->kp = h->hd.out_kp;
  return DRFE_OK;
}

extern "C" {

int drfe_orb_create(const drfe_orb_params* params, int width, int height, int max_batch, int device,
                    drfe_orb** out) {
  if (!params || !out) { set_error("drfe_orb_create: null argument"); return DRFE_ERR_ARG; }
  *out = nullptr;
  if (params->nlevels < 1 || params->nlevels > DRFE_MAX_LEVELS || params->nfeatures < 1 ||
      !(params->scale_factor > 1.0f) || params->min_th_fast < 1 || params->ini_th_fast < params->min_th_fast ||
      params->ini_th_fast > 254 || max_batch < 1 || width < 64 || height < 64) {
    set_error("drfe_orb_create: invalid parameters");
    return DRFE_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("drfe_orb_create: no CUDA device available (there is no CPU fallback)");
    return DRFE_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { set_error("drfe_orb_create: bad device %d", device); return DRFE_ERR_ARG; }
  DeviceScope ds(device);
  if (!ds.ok) { set_error("cudaSetDevice(%d) failed", device); return DRFE_ERR_CUDA; }
  drfe_orb* h = new drfe_orb();
  h->device = device; h->width = width; h->height = height; h->max_batch = max_batch; h->prm = *params;
  const int rc = orb_build(h);
  if (rc != DRFE_OK) { drfe_orb_destroy(h); return rc; }
  *out = h;
  return DRFE_OK;
}

int drfe_orb_destroy(drfe_orb* h) {
  if (!h) return DRFE_OK;
  DeviceScope ds(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->pipe.destroy();
  for (void* p : h->allocs) cudaFree(p);
  h->timer.destroy();
  if (h->ev_shared) cudaEventDestroy(h->ev_shared);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_hi) cudaEventDestroy(h->ev_hi);
  if (h->ev_blur_fork) cudaEventDestroy(h->ev_blur_fork);
  if (h->ev_blur_join) cudaEventDestroy(h->ev_blur_join);
  if (h->blur_stream) { cudaStreamSynchronize(h->blur_stream); cudaStreamDestroy(h->blur_stream); }
  if (h->ev_l0) cudaEventDestroy(h->ev_l0);
  if (h->ev_fast0) cudaEventDestroy(h->ev_fast0);
  if (h->fast0_stream) { cudaStreamSynchronize(h->fast0_stream); cudaStreamDestroy(h->fast0_stream); }
  if (h->hi) { cudaStreamSynchronize(h->hi); cudaStreamDestroy(h->hi); }
  for (int p = 0; p < drfe_orb::kMaxSplit - 1; ++p) {
    if (h->ev_join[p]) cudaEventDestroy(h->ev_join[p]);
    if (h->aux[p]) { cudaStreamSynchronize(h->aux[p]); cudaStreamDestroy(h->aux[p]); }
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return DRFE_OK;
}

int drfe_orb_get_levels(const drfe_orb* h) { return h ? h->prm.nlevels : 0; }
float drfe_orb_get_scale_factor(const drfe_orb* h) { return h ? (float)h->scaleFactorD : 0.f; }
int drfe_orb_get_scale_factors(const drfe_orb* h, float* s, float* is, float* s2, float* is2) {
  if (!h) return DRFE_ERR_ARG;
  for (int i = 0; i < h->prm.nlevels; ++i) {
    if (s) s[i] = h->scale[i];
    if (is) is[i] = h->inv_scale[i];
    if (s2) s2[i] = h->sigma2[i];
    if (is2) is2[i] = h->inv_sigma2[i];
  }
  return DRFE_OK;
}
int drfe_orb_features_per_level(const drfe_orb* h, int level) {
  return (h && level >= 0 && level < h->prm.nlevels) ? h->per_level[level] : DRFE_ERR_ARG;
}
int drfe_orb_max_keypoints(const drfe_orb* h) { return h ? h->hd.kp_cap : 0; }
void* drfe_orb_stream(drfe_orb* h) { return h ? (void*)h->stream : nullptr; }
int drfe_orb_level_size(const drfe_orb* h, int level, int* w, int* hgt) {
  if (!h || level < 0 || level >= h->prm.nlevels) { set_error("bad level"); return DRFE_ERR_ARG; }
  if (w) *w = h->hd.lv[level].w;
  if (hgt) *hgt = h->hd.lv[level].h;
  return DRFE_OK;
}
int drfe_orb_set_profiling(drfe_orb* h, int on) { if (!h) return DRFE_ERR_ARG; h->timer.reset(on != 0); return DRFE_OK; }
int drfe_orb_stage_times(drfe_orb* h, float* ms, const char** names, int cap, int* nstages) {
  if (!h || !ms || !nstages) return DRFE_ERR_ARG;
  DeviceScope ds(h->device);
  return h->timer.read(ms, names, cap, nstages);
}

// all kernels of frames [f0, f0 + n) on the handle's stream; src/rs/fs address frame 0 of the batch
static int orb_launch_on(drfe_orb* h, cudaStream_t st, int f0, int n, const uint8_t* src, long long rs, long long fs, bool timed) {
  NvtxRange nvtx_("orb_launch");
  const bool drfe_pdl_ = pdl_enabled() && n <= pdl_max_frames();
  h->post_valid = false;         // the keypoints drfe_orb_frame_post worked on are being replaced
  const OrbDev& D = h->hd;
  const int nl = D.nlevels;
  DRFE_LAUNCH_PDL(k_zero_counts, (n * nl + 255) / 256, 256, 0, st, h->dd, f0, n);
  {
    const LevelDev& L = D.lv[0];
    const int chunks = (kXOff + L.w + 20 + 15) / 16;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)rs | (uintptr_t)fs) & 15) == 0;
    const int ci = vec_ok ? L.w / 16 : 0, cb = chunks - ci;
    const uint32_t ci_magic = ci ? (uint32_t)(((1ull << 32) + ci - 1) / ci) : 0, cb_magic = (uint32_t)(((1ull << 32) + cb - 1) / cb);
    const long long threads = (long long)L.rows * (ci + cb);
    DRFE_LAUNCH_PDL(k_pyr_level0, dim3((unsigned)((threads + 255) / 256), n), 256, 0, st, h->dd, src, rs, fs, f0, ci, ci_magic, cb, cb_magic);
  }
  // small launches: FAST on level 0 (a third of its work) needs level 0 only and runs on another stream beside the seven resizes
  // (only for drfe_orb_enqueue — `timed` — with the stage timers off: the chunks of the pipelined batch calls are bound by the host link, and
  // the extra event calls per chunk cost the issuing thread more than the kernels gain: e2e 50.1 k -> 49.3 k frames/s with the forks there)
  const bool fork_small = timed && !h->timer.enabled && n <= kPdlMaxFrames && h->blur_stream != nullptr;
  const bool fork_fast0 = fork_small && nl > 1 && h->nstrips_l0 > 0 && h->nstrips_l0 < h->nstrips;
  if (fork_fast0) {
    DRFE_CUDA(cudaEventRecord(h->ev_l0, st));
    DRFE_CUDA(cudaStreamWaitEvent(h->fast0_stream, h->ev_l0, 0));
    DRFE_LAUNCH(k_fast_strips<256>, dim3(h->nstrips_l0, n), 256, h->fast_smem, h->fast0_stream, h->dd, f0, h->fast_maps, 0);
    DRFE_CUDA(cudaEventRecord(h->ev_fast0, h->fast0_stream));
  }
  for (int l = 1; l < nl; ++l) {
    const LevelDev& L = D.lv[l];
    if (D.pcols) {
      // rows per thread: long strips amortise the two priming rows on the big levels; the small levels are
      // latency-bound (few warps), so they get short strips = more threads and shorter dependent chains
      const long long px = (long long)L.w * L.h * n;
      const int R = px >= (24 << 20) ? kPyrR : (px >= (8 << 20) ? 8 : 4);
      const int strips = (L.h + R - 1) / R;
      DRFE_LAUNCH_PDL(k_pyr_stream, dim3((L.pcol_groups + 31) / 32, (strips + 3) / 4, n), dim3(32, 4), 0, st, h->dd, l, f0, R);
    } else {
      DRFE_LAUNCH_PDL(k_pyr_resize, dim3(L.ptile_cnt, n), 256, h->pyr_smem, st, h->dd, l, f0);
    }
  }
  if (timed) h->timer.mark("pyramid", st);
  // small launches are a chain of latencies: the blur needs the pyramid only, so it runs on a second stream beside FAST and the quadtree
  // and is joined in front of the descriptors (32 frames: ORB chain 0.29 -> 0.26 ms).  Large launches fill the GPU kernel by kernel
  // (there the fork gained nothing), timed launches keep the stages one after the other.
  const bool fork_blur = fork_small;
  if (fork_blur) {
    DRFE_CUDA(cudaEventRecord(h->ev_blur_fork, st));
    DRFE_CUDA(cudaStreamWaitEvent(h->blur_stream, h->ev_blur_fork, 0));
    DRFE_LAUNCH(k_blur, dim3(h->blur_blocks, n), 256, 0, h->blur_stream, h->dd, f0, 0);
    DRFE_CUDA(cudaEventRecord(h->ev_blur_join, h->blur_stream));
  }
  if (fork_fast0) {
    DRFE_LAUNCH_PDL(k_fast_strips<256>, dim3(h->nstrips - h->nstrips_l0, n), 256, h->fast_smem, st, h->dd, f0, h->fast_maps, h->nstrips_l0);
    DRFE_CUDA(cudaStreamWaitEvent(st, h->ev_fast0, 0));
  } else {
    DRFE_LAUNCH_PDL(k_fast_strips<256>, dim3(h->nstrips, n), 256, h->fast_smem, st, h->dd, f0, h->fast_maps, 0);
  }
  if (timed) h->timer.mark("fast", st);
  DRFE_LAUNCH_PDL(k_quadtree<256>, dim3(nl, n), 256, h->quad_smem, st, h->dd, f0);
  if (timed) h->timer.mark("quadtree", st);
  if (fork_blur) DRFE_CUDA(cudaStreamWaitEvent(st, h->ev_blur_join, 0));
  else DRFE_LAUNCH_PDL(k_blur, dim3(h->blur_blocks, n), 256, 0, st, h->dd, f0, 1);
  if (timed) h->timer.mark("blur", st);
  DRFE_LAUNCH_PDL(k_orient_describe, dim3((h->max_lkp + kOdWarps * kOdG - 1) / (kOdWarps * kOdG), nl, n), kOdWarps * 32, 0, st, h->dd, f0);
  if (timed) h->timer.mark("orient_describe", st);
  return DRFE_OK;
}

// A long batch is cut into `split` parts whose kernel chains run on separate streams (forked from and joined back into the
// handle's stream): the latency-bound stages of one part (the quadtree's serial passes, 45 % of the SMs busy) then overlap the
// issue-bound stages of the others.  DRFE_ORB_SPLIT sets the number of parts (default in orb_build).
static int orb_launch(drfe_orb* h, int f0, int n, const uint8_t* src, long long rs, long long fs, bool timed) {
  const int parts = std::min(h->split, n / 32);
  if (parts <= 1) {
    if (!h->hi || n < h->hi_min_frames) return orb_launch_on(h, h->stream, f0, n, src, rs, fs, timed);
    DRFE_CUDA(cudaEventRecord(h->ev_fork, h->stream));
    DRFE_CUDA(cudaStreamWaitEvent(h->hi, h->ev_fork, 0));
    const int rc = orb_launch_on(h, h->hi, f0, n, src, rs, fs, timed);
    if (rc != DRFE_OK) return rc;
    DRFE_CUDA(cudaEventRecord(h->ev_hi, h->hi));
    DRFE_CUDA(cudaStreamWaitEvent(h->stream, h->ev_hi, 0));
    return DRFE_OK;
  }
  DRFE_CUDA(cudaEventRecord(h->ev_fork, h->stream));
  const int base = n / parts, extra = n % parts;
  int first = f0 + base + (extra > 0 ? 1 : 0);
  for (int p = 1; p < parts; ++p) {
    const int np = base + (p < extra ? 1 : 0);
    DRFE_CUDA(cudaStreamWaitEvent(h->aux[p - 1], h->ev_fork, 0));
    const int rc = orb_launch_on(h, h->aux[p - 1], first, np, src, rs, fs, false);
    if (rc != DRFE_OK) return rc;
    DRFE_CUDA(cudaEventRecord(h->ev_join[p - 1], h->aux[p - 1]));
    first += np;
  }
  const int rc = orb_launch_on(h, h->stream, f0, base + (extra > 0 ? 1 : 0), src, rs, fs, timed);
  if (rc != DRFE_OK) return rc;
  for (int p = 1; p < parts; ++p) DRFE_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join[p - 1], 0));
  return DRFE_OK;
}

int drfe_orb_enqueue(drfe_orb* h, int nframes, const uint8_t* gray, size_t row_stride, size_t frame_stride,
                     int mem_kind) {
  NvtxRange nvtx_("drfe_orb_enqueue");
  if (!h || !gray) { set_error("drfe_orb_enqueue: null argument"); return DRFE_ERR_ARG; }
  if (nframes < 1 || nframes > h->max_batch) { set_error("drfe_orb_enqueue: nframes %d outside [1,%d]", nframes, h->max_batch); return DRFE_ERR_ARG; }
  if (row_stride < (size_t)h->width) { set_error("drfe_orb_enqueue: row_stride < width"); return DRFE_ERR_ARG; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  const uint8_t* src = gray;
  long long rs = (long long)row_stride, fs = (long long)frame_stride;
  h->timer.begin(st);
  if (mem_kind == DRFE_MEM_HOST) {
    if (row_stride == (size_t)h->width && (frame_stride == (size_t)h->width * h->height || nframes == 1)) {
      DRFE_CUDA(cudaMemcpyAsync(h->d_gray, gray, (size_t)nframes * h->width * h->height, cudaMemcpyHostToDevice, st));
    } else {
      for (int f = 0; f < nframes; ++f)
        DRFE_CUDA(cudaMemcpy2DAsync(h->d_gray + (size_t)f * h->width * h->height, h->width, gray + f * frame_stride,
                                    row_stride, h->width, h->height, cudaMemcpyHostToDevice, st));
    }
    src = h->d_gray; rs = h->width; fs = (long long)h->width * h->height;
    h->timer.mark("h2d", st);
  } else if (mem_kind != DRFE_MEM_DEVICE) {
    set_error("drfe_orb_enqueue: bad mem_kind"); return DRFE_ERR_ARG;
  }
  const int rc = orb_launch(h, 0, nframes, src, rs, fs, true);
  if (rc != DRFE_OK) return rc;
  h->last_frames = nframes;
  h->pending = true;
  h->gray_valid = (mem_kind == DRFE_MEM_HOST);
  return DRFE_OK;
}

int drfe_orb_enqueue_color(drfe_orb* h, int nframes, const uint8_t* pixels, int channels, int rgb_order, int coeffs, size_t row_stride,
                           size_t frame_stride, int mem_kind) {
  NvtxRange nvtx_("drfe_orb_enqueue_color");
  if (!h || !pixels) { set_error("drfe_orb_enqueue_color: null argument"); return DRFE_ERR_ARG; }
  if (nframes < 1 || nframes > h->max_batch) { set_error("drfe_orb_enqueue_color: nframes %d outside [1,%d]", nframes, h->max_batch); return DRFE_ERR_ARG; }
  if ((channels != 3 && channels != 4) || (coeffs != DRFE_GRAY_Q15 && coeffs != DRFE_GRAY_Q14)) { set_error("drfe_orb_enqueue_color: channels %d / coeffs %d", channels, coeffs); return DRFE_ERR_ARG; }
  const int W = h->width, H = h->height;
  if (row_stride < (size_t)W * channels || (nframes > 1 && frame_stride < row_stride * H)) { set_error("drfe_orb_enqueue_color: bad strides"); return DRFE_ERR_ARG; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  const uint8_t* src = pixels;
  long long rs = (long long)row_stride, fs = (long long)frame_stride;
  h->timer.begin(st);
  if (mem_kind == DRFE_MEM_HOST) {
    const size_t line = (size_t)W * channels;
    if (!h->d_color || h->color_bytes < line * H * h->max_batch) {
      if (dev_alloc(h, &h->d_color, line * H * h->max_batch)) return DRFE_ERR_CUDA;
      h->color_bytes = line * H * h->max_batch;
    }
    for (int f = 0; f < nframes; ++f)
      DRFE_CUDA(cudaMemcpy2DAsync(h->d_color + (size_t)f * line * H, line, pixels + f * frame_stride, row_stride, line, H, cudaMemcpyHostToDevice, st));
    src = h->d_color; rs = (long long)line; fs = (long long)(line * H);
    h->timer.mark("h2d", st);
  } else if (mem_kind != DRFE_MEM_DEVICE) {
    set_error("drfe_orb_enqueue_color: bad mem_kind"); return DRFE_ERR_ARG;
  }
  const bool q15 = coeffs == DRFE_GRAY_Q15;
  const int cr = q15 ? 9798 : 4899, cg = q15 ? 19235 : 9617, cb = q15 ? 3735 : 1868, shift = q15 ? 15 : 14;
  const dim3 grid(((W + 3) / 4 + 255) / 256, H, nframes);
  if (channels == 3) DRFE_LAUNCH(k_color_to_gray<3>, grid, 256, 0, st, src, rs, fs, h->d_gray, W, H, cr, cg, cb, shift, rgb_order ? 0 : 1);
  else DRFE_LAUNCH(k_color_to_gray<4>, grid, 256, 0, st, src, rs, fs, h->d_gray, W, H, cr, cg, cb, shift, rgb_order ? 0 : 1);
  const int rc = orb_launch(h, 0, nframes, h->d_gray, W, (long long)W * H, true);
  if (rc != DRFE_OK) return rc;
  h->last_frames = nframes;
  h->pending = true;
  h->gray_valid = true;
  return DRFE_OK;
}

int drfe_orb_get_gray(drfe_orb* h, int frame, uint8_t* dst) {
  if (!h || !dst) { set_error("drfe_orb_get_gray: null argument"); return DRFE_ERR_ARG; }
  if (!h->pending || !h->gray_valid || frame < 0 || frame >= h->last_frames) { set_error("drfe_orb_get_gray: no converted frame %d", frame); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  DRFE_CUDA(cudaMemcpyAsync(dst, h->d_gray + (size_t)frame * h->width * h->height, (size_t)h->width * h->height, cudaMemcpyDeviceToHost, h->stream));
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  return DRFE_OK;
}

int drfe_orb_sync(drfe_orb* h) {
  if (!h) return DRFE_ERR_ARG;
  DeviceScope ds(h->device);
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  return DRFE_OK;
}

static int orb_check_status(drfe_orb* h) {
  int st = 0;
  DRFE_CUDA(cudaMemcpyAsync(&st, h->hd.status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  if (st) {
    DRFE_CUDA(cudaMemsetAsync(h->hd.status, 0, sizeof(int), h->stream));
    set_error("device buffer overflow (status %d: 1 = FAST candidates, 2 = quadtree nodes)", st);
    return DRFE_ERR_CAPACITY;
  }
  return DRFE_OK;
}

int drfe_orb_download(drfe_orb* h, drfe_keypoint* kps, uint8_t* desc, int cap_per_frame, int* counts) {
  NvtxRange nvtx_("drfe_orb_download");
  if (!h || !counts) { set_error("drfe_orb_download: null argument"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("drfe_orb_download: nothing enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  cudaStream_t st = h->stream;
  const int nf = h->last_frames, cap = h->hd.kp_cap;
  DRFE_CUDA(cudaMemcpyAsync(counts, h->hd.out_cnt, nf * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (kps) {
    if (cap_per_frame >= cap)
      DRFE_CUDA(cudaMemcpy2DAsync(kps, (size_t)cap_per_frame * sizeof(drfe_keypoint), h->hd.out_kp, (size_t)cap * sizeof(drfe_keypoint),
                                  (size_t)cap * sizeof(drfe_keypoint), nf, cudaMemcpyDeviceToHost, st));
    else
      DRFE_CUDA(cudaMemcpy2DAsync(kps, (size_t)cap_per_frame * sizeof(drfe_keypoint), h->hd.out_kp, (size_t)cap * sizeof(drfe_keypoint),
                                  (size_t)cap_per_frame * sizeof(drfe_keypoint), nf, cudaMemcpyDeviceToHost, st));
  }
  if (desc) {
    const size_t wbytes = (size_t)std::min(cap, cap_per_frame) * 32;
    DRFE_CUDA(cudaMemcpy2DAsync(desc, (size_t)cap_per_frame * 32, h->hd.out_desc, (size_t)cap * 32, wbytes, nf,
                                cudaMemcpyDeviceToHost, st));
  }
  const int rc = orb_check_status(h);
  if (rc != DRFE_OK) return rc;
  for (int f = 0; f < nf; ++f)
    if (counts[f] > cap_per_frame && (kps || desc)) {
      set_error("drfe_orb_download: frame %d has %d keypoints, cap_per_frame is %d", f, counts[f], cap_per_frame);
      return DRFE_ERR_CAPACITY;
    }
  return DRFE_OK;
}

int drfe_orb_extract(drfe_orb* h, const uint8_t* gray, int width, int height, size_t row_stride, drfe_keypoint* kps,
                     uint8_t* desc, int cap, int* n) {
  NvtxRange nvtx_("drfe_orb_extract");
  if (!h) { set_error("drfe_orb_extract: null handle"); return DRFE_ERR_ARG; }
  if (!n) { set_error("drfe_orb_extract: null count pointer"); return DRFE_ERR_ARG; }
  if (!gray || width == 0 || height == 0) { *n = 0; return DRFE_OK; }  // empty image: silent return (ORBextractor.cc:1046); the caller's containers stay as they are
  if (width != h->width || height != h->height) { set_error("drfe_orb_extract: image is %dx%d, handle was created for %dx%d", width, height, h->width, h->height); return DRFE_ERR_ARG; }
  int rc = drfe_orb_enqueue(h, 1, gray, row_stride, row_stride * height, DRFE_MEM_HOST);
  if (rc != DRFE_OK) return rc;
  return drfe_orb_download(h, kps, desc, cap, n);
}


int drfe_orb_extract_batch(drfe_orb* h, int nframes, const uint8_t* gray, size_t row_stride, size_t frame_stride,
                           drfe_keypoint* kps, uint8_t* desc, int cap_per_frame, int* counts, int chunk_frames) {
  NvtxRange nvtx_("drfe_orb_extract_batch");
  if (!h || !gray || !counts) { set_error("drfe_orb_extract_batch: null argument"); return DRFE_ERR_ARG; }
  if (nframes < 1 || nframes > h->max_batch) { set_error("drfe_orb_extract_batch: nframes %d outside [1,%d]", nframes, h->max_batch); return DRFE_ERR_ARG; }
  if (row_stride < (size_t)h->width || (nframes > 1 && frame_stride < row_stride * h->height)) { set_error("drfe_orb_extract_batch: bad strides"); return DRFE_ERR_ARG; }
  if (h->pipe.active) { set_error("drfe_orb_extract_batch: the previous batch was not finished (drfe_orb_finish_batch)"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  if (h->pipe.create() != DRFE_OK) return DRFE_ERR_CUDA;
  ChunkPipe& pp = h->pipe;
  cudaStream_t st = h->stream;
  const int W = h->width, H = h->height, cap = h->hd.kp_cap;
  const size_t fbytes = (size_t)W * H;
  int cstart[ChunkPipe::kMaxChunks + 1];
  const int nchunks = ChunkPipe::schedule(nframes, chunk_frames, cstart);
  // the copy streams start after whatever the handle's stream was doing with the staging / result buffers
  DRFE_CUDA(cudaEventRecord(pp.ev_start, st));
  DRFE_CUDA(cudaStreamWaitEvent(pp.h2d, pp.ev_start, 0));
  DRFE_CUDA(cudaStreamWaitEvent(pp.d2h, pp.ev_start, 0));
  const bool dense = row_stride == (size_t)W && frame_stride == fbytes;
  for (int k = 0; k < nchunks; ++k) {
    const int f0 = cstart[k], n = cstart[k + 1] - f0;
    if (dense || n == 1)
      DRFE_CUDA(cudaMemcpy2DAsync(h->d_gray + f0 * fbytes, W, gray + (size_t)f0 * frame_stride, row_stride, W, (size_t)H * (dense ? n : 1),
                                  cudaMemcpyHostToDevice, pp.h2d));
    else
      for (int f = f0; f < f0 + n; ++f)
        DRFE_CUDA(cudaMemcpy2DAsync(h->d_gray + f * fbytes, W, gray + (size_t)f * frame_stride, row_stride, W, H, cudaMemcpyHostToDevice, pp.h2d));
    DRFE_CUDA(cudaEventRecord(pp.ev_in[k], pp.h2d));
    DRFE_CUDA(cudaStreamWaitEvent(st, pp.ev_in[k], 0));
    const int rc = orb_launch(h, f0, n, h->d_gray, W, (long long)fbytes, false);
    if (rc != DRFE_OK) { cudaStreamSynchronize(pp.h2d); cudaStreamSynchronize(st); cudaStreamSynchronize(pp.d2h); return rc; }   // nothing stays queued on the caller's buffers
    DRFE_CUDA(cudaEventRecord(pp.ev_done[k], st));
    DRFE_CUDA(cudaStreamWaitEvent(pp.d2h, pp.ev_done[k], 0));
    DRFE_CUDA(cudaMemcpyAsync(counts + f0, h->hd.out_cnt + f0, n * sizeof(int), cudaMemcpyDeviceToHost, pp.d2h));
    const int wk = std::min(cap, cap_per_frame);
    if (kps)
      DRFE_CUDA(cudaMemcpy2DAsync(kps + (size_t)f0 * cap_per_frame, (size_t)cap_per_frame * sizeof(drfe_keypoint), h->hd.out_kp + (size_t)f0 * cap,
                                  (size_t)cap * sizeof(drfe_keypoint), (size_t)wk * sizeof(drfe_keypoint), n, cudaMemcpyDeviceToHost, pp.d2h));
    if (desc)
      DRFE_CUDA(cudaMemcpy2DAsync(desc + (size_t)f0 * cap_per_frame * 32, (size_t)cap_per_frame * 32, h->hd.out_desc + (size_t)f0 * cap * 32,
                                  (size_t)cap * 32, (size_t)wk * 32, n, cudaMemcpyDeviceToHost, pp.d2h));
  }
  DRFE_CUDA(cudaMemcpyAsync(pp.h_status, h->hd.status, sizeof(int), cudaMemcpyDeviceToHost, pp.d2h));
  pp.active = true;
  h->batch_counts = counts; h->batch_cap = (kps || desc) ? cap_per_frame : 0x7FFFFFFF;
  h->last_frames = nframes;
  h->pending = true;
  return DRFE_OK;
}

int drfe_orb_finish_batch(drfe_orb* h) {
  NvtxRange nvtx_("drfe_orb_finish_batch");
  if (!h) { set_error("drfe_orb_finish_batch: null handle"); return DRFE_ERR_ARG; }
  if (!h->pipe.active) { set_error("drfe_orb_finish_batch: no batch in flight"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  DRFE_CUDA(cudaStreamSynchronize(h->pipe.d2h));
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  h->pipe.active = false;
  if (*h->pipe.h_status) {
    const int stv = *h->pipe.h_status;
    *h->pipe.h_status = 0;
    DRFE_CUDA(cudaMemsetAsync(h->hd.status, 0, sizeof(int), h->stream));
    set_error("device buffer overflow (status %d: 1 = FAST candidates, 2 = quadtree nodes)", stv);
    return DRFE_ERR_CAPACITY;
  }
  for (int f = 0; f < h->last_frames; ++f)
    if (h->batch_counts[f] > h->batch_cap) {
      set_error("drfe_orb_finish_batch: frame %d has %d keypoints, cap_per_frame is %d", f, h->batch_counts[f], h->batch_cap);
      return DRFE_ERR_CAPACITY;
    }
  return DRFE_OK;
}


// Frame::ComputeImageBounds (Frame.cc:863-891)
int drfe_frame_image_bounds(drfe_frame_params* p, int width, int height) {
  if (!p || width < 1 || height < 1) { set_error("drfe_frame_image_bounds: bad argument"); return DRFE_ERR_ARG; }
  if (p->dist[0] != 0.0f) {
    const float cu[4] = {0.f, (float)width, 0.f, (float)width}, cv[4] = {0.f, 0.f, (float)height, (float)height};
    float x[4], y[4];
    for (int i = 0; i < 4; ++i) {   // the same five iterations as undistort_point, on the host
      const double fx = p->fx, fy = p->fy, cx = p->cx, cy = p->cy;
      const double k1 = p->dist[0], k2 = p->dist[1], p1 = p->dist[2], p2 = p->dist[3], k3 = p->dist[4];
      const double ifx = 1. / fx, ify = 1. / fy;
      double xx = ((double)cu[i] - cx) * ifx, yy = ((double)cv[i] - cy) * ify;
      const double x0 = xx, y0 = yy;
      for (int j = 0; j < 5; ++j) {
        const double r2 = xx * xx + yy * yy;
        const double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
        const double dx = 2 * p1 * xx * yy + p2 * (r2 + 2 * xx * xx);
        const double dy = p1 * (r2 + 2 * yy * yy) + 2 * p2 * xx * yy;
        xx = (x0 - dx) * icdist;
        yy = (y0 - dy) * icdist;
      }
      x[i] = (float)(fx * xx + cx); y[i] = (float)(fy * yy + cy);
    }
    p->min_x = std::min(x[0], x[2]); p->max_x = std::max(x[1], x[3]);
    p->min_y = std::min(y[0], y[1]); p->max_y = std::max(y[2], y[3]);
  } else {
    p->min_x = 0.f; p->max_x = (float)width; p->min_y = 0.f; p->max_y = (float)height;
  }
  return DRFE_OK;
}

// launch k_frame_post on the depth Q names and copy the requested outputs out
static int frame_post_run(drfe_orb* h, drfe_keypoint* keys_un, float* u_right, float* kp_depth, uint16_t* grid_count, uint16_t* grid_index,
                          int cap_per_frame) {
  NvtxRange nvtx_("frame_post_run");
  cudaStream_t st = h->stream;
  const int nf = h->last_frames, cap = h->hd.kp_cap;
  PostDev& Q = h->post;
  DRFE_LAUNCH(k_frame_post, nf, 256, cap * sizeof(unsigned short), st, h->dd, Q);
  h->post_valid = true;
  const int wk = std::min(cap, cap_per_frame);
  if (keys_un)
    DRFE_CUDA(cudaMemcpy2DAsync(keys_un, (size_t)cap_per_frame * sizeof(drfe_keypoint), Q.keys_un, (size_t)cap * sizeof(drfe_keypoint),
                                (size_t)wk * sizeof(drfe_keypoint), nf, cudaMemcpyDeviceToHost, st));
  if (u_right)
    DRFE_CUDA(cudaMemcpy2DAsync(u_right, (size_t)cap_per_frame * 4, Q.u_right, (size_t)cap * 4, (size_t)wk * 4, nf, cudaMemcpyDeviceToHost, st));
  if (kp_depth)
    DRFE_CUDA(cudaMemcpy2DAsync(kp_depth, (size_t)cap_per_frame * 4, Q.kp_depth, (size_t)cap * 4, (size_t)wk * 4, nf, cudaMemcpyDeviceToHost, st));
  if (grid_count) DRFE_CUDA(cudaMemcpyAsync(grid_count, Q.grid_count, (size_t)kGridCells * nf * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
  if (grid_index)
    DRFE_CUDA(cudaMemcpy2DAsync(grid_index, (size_t)cap_per_frame * 2, Q.grid_index, (size_t)cap * 2, (size_t)wk * 2, nf, cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaStreamSynchronize(st));
  return DRFE_OK;
}

static int frame_post_buffers(drfe_orb* h, const drfe_frame_params* p) {
  const int cap = h->hd.kp_cap, B = h->max_batch;
  PostDev& Q = h->post;
  if (cap > 65535) { set_error("drfe_orb_frame_post: %d keypoints per frame do not fit the 16-bit grid indices", cap); return DRFE_ERR_CAPACITY; }
  if (!Q.keys_un) {
    if (dev_alloc(h, &Q.keys_un, (size_t)cap * B) || dev_alloc(h, &Q.u_right, (size_t)cap * B) || dev_alloc(h, &Q.kp_depth, (size_t)cap * B) ||
        dev_alloc(h, &Q.grid_count, (size_t)kGridCells * B) || dev_alloc(h, &Q.grid_index, (size_t)cap * B)) return DRFE_ERR_CUDA;
    DRFE_CUDA(raise_dyn_smem(k_frame_post, h->device, (size_t)((cap * sizeof(unsigned short)))));
  }
  Q.prm = *p;
  Q.depth = nullptr; Q.depth16 = nullptr; Q.depth_factor = 1.f;
  return DRFE_OK;
}

int drfe_orb_frame_post(drfe_orb* h, const drfe_frame_params* p, const float* depth, size_t row_stride, size_t frame_stride,
                        int mem_kind, drfe_keypoint* keys_un, float* u_right, float* kp_depth, uint16_t* grid_count,
                        uint16_t* grid_index, int cap_per_frame) {
  if (!h || !p || !depth) { set_error("drfe_orb_frame_post: null argument"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("drfe_orb_frame_post: nothing enqueued"); return DRFE_ERR_STATE; }
  if (!(p->max_x > p->min_x) || !(p->max_y > p->min_y)) { set_error("drfe_orb_frame_post: image bounds not set (drfe_frame_image_bounds)"); return DRFE_ERR_ARG; }
  if (row_stride < (size_t)h->width) { set_error("drfe_orb_frame_post: row_stride < width"); return DRFE_ERR_ARG; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  const int nf = h->last_frames, B = h->max_batch;
  const size_t N = (size_t)h->width * h->height;
  if (frame_post_buffers(h, p)) return DRFE_ERR_CUDA;
  PostDev& Q = h->post;
  if (mem_kind == DRFE_MEM_HOST) {
    if (!h->d_post_depth && dev_alloc(h, &h->d_post_depth, N * B)) return DRFE_ERR_CUDA;
    for (int f = 0; f < nf; ++f)
      DRFE_CUDA(cudaMemcpy2DAsync(h->d_post_depth + f * N, h->width * sizeof(float), depth + f * frame_stride, row_stride * sizeof(float),
                                  h->width * sizeof(float), h->height, cudaMemcpyHostToDevice, st));
    Q.depth = h->d_post_depth; Q.depth_rs = h->width; Q.depth_fs = (long long)N;
  } else if (mem_kind == DRFE_MEM_DEVICE) {
    Q.depth = depth; Q.depth_rs = (long long)row_stride; Q.depth_fs = (long long)frame_stride;
  } else { set_error("drfe_orb_frame_post: bad mem_kind"); return DRFE_ERR_ARG; }
  return frame_post_run(h, keys_un, u_right, kp_depth, grid_count, grid_index, cap_per_frame);
}

int drfe_orb_frame_post_shared_depth(drfe_orb* h, drfe_cape* cape, const drfe_frame_params* p, drfe_keypoint* keys_un, float* u_right,
                                     float* kp_depth, uint16_t* grid_count, uint16_t* grid_index, int cap_per_frame) {
  CapeDepthView V;
  if (!h || !p || cape_depth_view(cape, &V)) { set_error("drfe_orb_frame_post_shared_depth: null argument"); return DRFE_ERR_ARG; }
  if (!h->pending || !V.pending) { set_error("drfe_orb_frame_post_shared_depth: nothing enqueued"); return DRFE_ERR_STATE; }
  if (!V.depth && !V.depth16) { set_error("drfe_orb_frame_post_shared_depth: the CAPE handle was fed a cloud, not depth images"); return DRFE_ERR_STATE; }
  if (V.device != h->device || V.width != h->width || V.height != h->height || V.nframes < h->last_frames) {
    set_error("drfe_orb_frame_post_shared_depth: the CAPE handle holds %d frames of %dx%d on device %d, the extractor %d of %dx%d on device %d",
              V.nframes, V.width, V.height, V.device, h->last_frames, h->width, h->height, h->device);
    return DRFE_ERR_ARG;
  }
  if (!(p->max_x > p->min_x) || !(p->max_y > p->min_y)) { set_error("drfe_orb_frame_post_shared_depth: image bounds not set (drfe_frame_image_bounds)"); return DRFE_ERR_ARG; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  if (frame_post_buffers(h, p)) return DRFE_ERR_CUDA;
  // the depth copy was queued on the CAPE handle's stream: this handle's stream waits for it (no host synchronisation)
  if (!h->ev_shared) DRFE_CUDA(cudaEventCreateWithFlags(&h->ev_shared, cudaEventDisableTiming));
  DRFE_CUDA(cudaEventRecord(h->ev_shared, V.stream));
  DRFE_CUDA(cudaStreamWaitEvent(h->stream, h->ev_shared, 0));
  PostDev& Q = h->post;
  Q.depth = V.depth; Q.depth16 = V.depth16; Q.depth_factor = V.factor; Q.depth_rs = V.row_stride; Q.depth_fs = V.frame_stride;
  return frame_post_run(h, keys_un, u_right, kp_depth, grid_count, grid_index, cap_per_frame);
}

int drfe_orb_search_by_projection(drfe_orb* h, const int* nqueries, const drfe_proj_query* queries, const uint8_t* qdesc,
                                  const uint8_t* occupied, int qcap, drfe_proj_match* out) {
  NvtxRange nvtx_("drfe_orb_search_by_projection");
  if (!h || !nqueries || !queries || !qdesc || !out || qcap < 1) { set_error("drfe_orb_search_by_projection: bad argument"); return DRFE_ERR_ARG; }
  if (!h->pending || !h->post.keys_un || !h->post_valid) { set_error("drfe_orb_search_by_projection: run drfe_orb_frame_post on this batch first"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  const int nf = h->last_frames, cap = h->hd.kp_cap, B = h->max_batch;
  if (qcap > h->search_qcap) {            // (re)allocate the query-side buffers; the old ones stay in the handle's arena until destroy
    if (dev_alloc(h, &h->d_sq, (size_t)qcap * B) || dev_alloc(h, &h->d_sdesc, (size_t)qcap * B * 32) || dev_alloc(h, &h->d_sout, (size_t)qcap * B)) return DRFE_ERR_CUDA;
    if (!h->d_socc && (dev_alloc(h, &h->d_socc, (size_t)cap * B) || dev_alloc(h, &h->d_snq, (size_t)B))) return DRFE_ERR_CUDA;
    h->search_qcap = qcap;
  }
  for (int f = 0; f < nf; ++f)
    if (nqueries[f] < 0 || nqueries[f] > qcap) { set_error("drfe_orb_search_by_projection: frame %d has %d queries, qcap is %d", f, nqueries[f], qcap); return DRFE_ERR_ARG; }
  DRFE_CUDA(cudaMemcpyAsync(h->d_snq, nqueries, (size_t)nf * sizeof(int), cudaMemcpyHostToDevice, st));
  DRFE_CUDA(cudaMemcpyAsync(h->d_sq, queries, (size_t)nf * qcap * sizeof(drfe_proj_query), cudaMemcpyHostToDevice, st));
  DRFE_CUDA(cudaMemcpyAsync(h->d_sdesc, qdesc, (size_t)nf * qcap * 32, cudaMemcpyHostToDevice, st));
  if (occupied) DRFE_CUDA(cudaMemcpyAsync(h->d_socc, occupied, (size_t)nf * cap, cudaMemcpyHostToDevice, st));
  SearchDev S;
  S.q = h->d_sq; S.qdesc = h->d_sdesc; S.occupied = occupied ? h->d_socc : nullptr; S.nq = h->d_snq; S.out = h->d_sout; S.qcap = qcap;
  DRFE_LAUNCH(k_search_projection, nf, 256, 0, st, h->dd, h->post, S);
  DRFE_CUDA(cudaMemcpyAsync(out, h->d_sout, (size_t)nf * qcap * sizeof(drfe_proj_match), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaStreamSynchronize(st));
  return DRFE_OK;
}

int drfe_orb_search_local_points(drfe_orb* h, const int* nqueries, const drfe_proj_query* queries, const uint8_t* qdesc, const uint8_t* qflags,
                                 const uint8_t* occupied, int qcap, float nnratio, drfe_proj_match* out, int32_t* assigned, int32_t* key_point,
                                 int* nmatches) {
  NvtxRange nvtx_("drfe_orb_search_local_points");
  if (!h || !nqueries || !queries || !qdesc || !qflags || qcap < 1) { set_error("drfe_orb_search_local_points: bad argument"); return DRFE_ERR_ARG; }
  if (!h->pending || !h->post.keys_un || !h->post_valid) { set_error("drfe_orb_search_local_points: run drfe_orb_frame_post on this batch first"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  const int nf = h->last_frames, cap = h->hd.kp_cap;
  const size_t smem = ((size_t)cap + qcap) * sizeof(int);
  if (smem > 160 * 1024) { set_error("drfe_orb_search_local_points: qcap %d too large", qcap); return DRFE_ERR_CAPACITY; }
  for (int f = 0; f < nf; ++f)
    if (nqueries[f] < 0 || nqueries[f] > qcap) { set_error("drfe_orb_search_local_points: frame %d has %d queries, qcap is %d", f, nqueries[f], qcap); return DRFE_ERR_ARG; }
  // one stream-ordered scratch block per call
  const size_t nq = (size_t)nf * qcap, nc = (size_t)nf * cap;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_q = take(nq * sizeof(drfe_proj_query)), o_d = take(nq * 32), o_fl = take(nq), o_occ = take(nc), o_n = take((size_t)nf * 4),
               o_out = take(nq * sizeof(drfe_proj_match)), o_as = take(nq * 4), o_kp = take(nc * 4), o_nm = take((size_t)nf * 4);
  char* d = nullptr;
  DRFE_CUDA(cudaMallocAsync((void**)&d, off, st));
  cudaError_t e = cudaSuccess;
  auto acc = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  auto up = [&](size_t o, const void* src, size_t bytes) { acc(cudaMemcpyAsync(d + o, src, bytes, cudaMemcpyHostToDevice, st)); };
  up(o_q, queries, nq * sizeof(drfe_proj_query)); up(o_d, qdesc, nq * 32); up(o_fl, qflags, nq); up(o_n, nqueries, (size_t)nf * 4);
  if (occupied) up(o_occ, occupied, nc);
  LocalDev S;
  S.q = (const drfe_proj_query*)(d + o_q); S.qdesc = (const uint8_t*)(d + o_d); S.qflags = (const uint8_t*)(d + o_fl);
  S.occupied = occupied ? (const uint8_t*)(d + o_occ) : nullptr; S.nq = (const int*)(d + o_n); S.out = (drfe_proj_match*)(d + o_out);
  S.assigned = (int32_t*)(d + o_as); S.key_point = (int32_t*)(d + o_kp); S.nmatches = (int*)(d + o_nm); S.qcap = qcap; S.nnratio = nnratio;
  acc(raise_dyn_smem(k_search_local_points, h->device, (size_t)(smem)));
  if (e == cudaSuccess) {
    k_search_local_points<<<nf, kTrackThreads, smem, st>>>(h->dd, h->post, S);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    acc(cudaGetLastError());
  }
  auto down = [&](void* dst, size_t o, size_t bytes) { if (dst) acc(cudaMemcpyAsync(dst, d + o, bytes, cudaMemcpyDeviceToHost, st)); };
  down(out, o_out, nq * sizeof(drfe_proj_match)); down(assigned, o_as, nq * 4); down(key_point, o_kp, nc * 4); down(nmatches, o_nm, (size_t)nf * 4);
  cudaFreeAsync(d, st);
  acc(cudaStreamSynchronize(st));
  if (e != cudaSuccess) { set_error("drfe_orb_search_local_points: %s", cudaGetErrorString(e)); return DRFE_ERR_CUDA; }
  return DRFE_OK;
}

int drfe_orb_search_last_frame(drfe_orb* h, const drfe_track_params* tp, const int* npoints, const drfe_last_point* points,
                               const uint8_t* pdesc, const uint8_t* occupied, int pcap, int32_t* match_key, int32_t* match_dist,
                               int32_t* key_point, int* nmatches, int* sweeps) {
  NvtxRange nvtx_("drfe_orb_search_last_frame");
  if (!h || !tp || !npoints || !points || !pdesc || pcap < 1) { set_error("drfe_orb_search_last_frame: bad argument"); return DRFE_ERR_ARG; }
  if (!h->pending || !h->post.keys_un || !h->post_valid) { set_error("drfe_orb_search_last_frame: run drfe_orb_frame_post on this batch first"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  const int nf = h->last_frames, cap = h->hd.kp_cap, B = h->max_batch;
  const size_t smem = ((size_t)cap + 3 * (size_t)pcap) * sizeof(int) + (size_t)pcap;
  if (smem > 160 * 1024) { set_error("drfe_orb_search_last_frame: pcap %d too large", pcap); return DRFE_ERR_CAPACITY; }
  for (int f = 0; f < nf; ++f) {
    if (npoints[f] < 0 || npoints[f] > pcap) { set_error("drfe_orb_search_last_frame: frame %d has %d points, pcap is %d", f, npoints[f], pcap); return DRFE_ERR_ARG; }
    if (tp[f].mode < 0 || tp[f].mode > 2) { set_error("drfe_orb_search_last_frame: frame %d: mode %d", f, tp[f].mode); return DRFE_ERR_ARG; }
    for (int i = 0; i < npoints[f]; ++i) {
      const drfe_last_point& lp = points[(size_t)f * pcap + i];
      if ((lp.flags & DRFE_LP_VALID) && (lp.octave < 0 || lp.octave >= h->prm.nlevels)) {
        set_error("drfe_orb_search_last_frame: frame %d point %d: octave %d", f, i, lp.octave); return DRFE_ERR_ARG;
      }
    }
  }
  if (pcap > h->track_pcap) {
    if (dev_alloc(h, &h->d_tpts, (size_t)pcap * B) || dev_alloc(h, &h->d_tdesc, (size_t)pcap * B * 32) || dev_alloc(h, &h->d_tout, (size_t)pcap * B * 2)) return DRFE_ERR_CUDA;
    if (!h->d_ttp && (dev_alloc(h, &h->d_ttp, (size_t)B) || dev_alloc(h, &h->d_tkey, (size_t)cap * B) || dev_alloc(h, &h->d_tcnt, (size_t)B * 3))) return DRFE_ERR_CUDA;
    if (!h->d_socc && (dev_alloc(h, &h->d_socc, (size_t)cap * B) || dev_alloc(h, &h->d_snq, (size_t)B))) return DRFE_ERR_CUDA;
    DRFE_CUDA(raise_dyn_smem(k_search_last_frame, h->device, (size_t)(smem)));
    h->track_pcap = pcap;
  }
  DRFE_CUDA(cudaMemcpyAsync(h->d_tcnt, npoints, (size_t)nf * sizeof(int), cudaMemcpyHostToDevice, st));
  DRFE_CUDA(cudaMemcpyAsync(h->d_ttp, tp, (size_t)nf * sizeof(drfe_track_params), cudaMemcpyHostToDevice, st));
  DRFE_CUDA(cudaMemcpyAsync(h->d_tpts, points, (size_t)nf * pcap * sizeof(drfe_last_point), cudaMemcpyHostToDevice, st));
  DRFE_CUDA(cudaMemcpyAsync(h->d_tdesc, pdesc, (size_t)nf * pcap * 32, cudaMemcpyHostToDevice, st));
  if (occupied) DRFE_CUDA(cudaMemcpyAsync(h->d_socc, occupied, (size_t)nf * cap, cudaMemcpyHostToDevice, st));
  TrackDev S;
  S.tp = h->d_ttp; S.pts = h->d_tpts; S.pdesc = h->d_tdesc; S.occupied = occupied ? h->d_socc : nullptr; S.np = h->d_tcnt;
  S.match_key = h->d_tout; S.match_dist = h->d_tout + (size_t)pcap * B; S.key_point = h->d_tkey; S.nmatches = h->d_tcnt + B; S.sweeps = h->d_tcnt + 2 * B;
  S.pcap = pcap;
  DRFE_LAUNCH(k_search_last_frame, nf, kTrackThreads, smem, st, h->dd, h->post, S);
  if (match_key) DRFE_CUDA(cudaMemcpyAsync(match_key, S.match_key, (size_t)nf * pcap * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (match_dist) DRFE_CUDA(cudaMemcpyAsync(match_dist, S.match_dist, (size_t)nf * pcap * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (key_point) DRFE_CUDA(cudaMemcpyAsync(key_point, S.key_point, (size_t)nf * cap * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (nmatches) DRFE_CUDA(cudaMemcpyAsync(nmatches, S.nmatches, (size_t)nf * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (sweeps) DRFE_CUDA(cudaMemcpyAsync(sweeps, S.sweeps, (size_t)nf * sizeof(int), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaStreamSynchronize(st));
  return DRFE_OK;
}

// ---------------------------------------------------------------- intermediates
static int orb_frame_level_ok(drfe_orb* h, int frame, int level) {
  if (!h) { set_error("null handle"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("nothing enqueued"); return DRFE_ERR_STATE; }
  if (frame < 0 || frame >= h->last_frames || level < 0 || level >= h->prm.nlevels) { set_error("bad frame/level"); return DRFE_ERR_ARG; }
  return DRFE_OK;
}

int drfe_orb_get_pyramid(drfe_orb* h, int frame, int level, int bordered, uint8_t* dst) {
  int rc = orb_frame_level_ok(h, frame, level);
  if (rc) return rc;
  if (!dst) return DRFE_ERR_ARG;
  DeviceScope ds(h->device);
  const LevelDev& L = h->hd.lv[level];
  const uint8_t* base = h->hd.pyr + L.img_off + (long long)frame * L.img_fstride;
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  if (bordered)
    DRFE_CUDA(cudaMemcpy2D(dst, L.w + 2 * kEdge, base + (kXOff - kEdge), L.pitch, L.w + 2 * kEdge, L.rows, cudaMemcpyDeviceToHost));
  else
    DRFE_CUDA(cudaMemcpy2D(dst, L.w, base + (long long)kEdge * L.pitch + kXOff, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
  return DRFE_OK;
}

int drfe_orb_get_blurred(drfe_orb* h, int frame, int level, uint8_t* dst) {
  int rc = orb_frame_level_ok(h, frame, level);
  if (rc) return rc;
  if (!dst) return DRFE_ERR_ARG;
  DeviceScope ds(h->device);
  const LevelDev& L = h->hd.lv[level];
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  DRFE_CUDA(cudaMemcpy2D(dst, L.w, h->hd.blur + L.blur_off + (long long)frame * L.blur_fstride, L.bpitch, L.w, L.h, cudaMemcpyDeviceToHost));
  return DRFE_OK;
}

int drfe_orb_get_candidates(drfe_orb* h, int frame, int level, float* xyr, int cap, int* n) {
  int rc = orb_frame_level_ok(h, frame, level);
  if (rc) return rc;
  if (!n) return DRFE_ERR_ARG;
  DeviceScope ds(h->device);
  const LevelDev& L = h->hd.lv[level];
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  int cnt = 0;
  DRFE_CUDA(cudaMemcpy(&cnt, h->hd.cand_cnt + frame * h->prm.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
  *n = cnt;
  const int m = std::min(std::min(cnt, cap), L.cand_cap);
  if (xyr && m > 0) {
    std::vector<uint32_t> tmp(m);
    DRFE_CUDA(cudaMemcpy(tmp.data(), h->hd.cand + (long long)frame * h->hd.cand_fstride + L.cand_off, (size_t)m * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; ++i) {
      xyr[3 * i] = (float)(tmp[i] & 0xFFF); xyr[3 * i + 1] = (float)((tmp[i] >> 12) & 0xFFF); xyr[3 * i + 2] = (float)(tmp[i] >> 24);
    }
  }
  if (cnt > L.cand_cap) { set_error("candidate buffer overflow on level %d (%d > %d)", level, cnt, L.cand_cap); return DRFE_ERR_CAPACITY; }
  return DRFE_OK;
}

int drfe_orb_get_level_keypoints(drfe_orb* h, int frame, int level, drfe_keypoint* dst, int cap, int* n) {
  int rc = orb_frame_level_ok(h, frame, level);
  if (rc) return rc;
  if (!n) return DRFE_ERR_ARG;
  DeviceScope ds(h->device);
  const int nl = h->prm.nlevels, kcap = h->hd.kp_cap;
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  std::vector<int> cnts(nl);
  DRFE_CUDA(cudaMemcpy(cnts.data(), h->hd.lkp_cnt + frame * nl, nl * sizeof(int), cudaMemcpyDeviceToHost));
  int base = 0;
  for (int l = 0; l < level; ++l) base += cnts[l];
  const int cnt = cnts[level];
  *n = cnt;
  const int m = std::min(cnt, cap);
  if (dst && m > 0) {
    DRFE_CUDA(cudaMemcpy(dst, h->hd.out_kp + (long long)frame * kcap + base, (size_t)m * sizeof(drfe_keypoint), cudaMemcpyDeviceToHost));
    // out_kp holds level-0-scaled coordinates; undo with the exact packed level coordinates
    std::vector<uint32_t> pk(m);
    DRFE_CUDA(cudaMemcpy(pk.data(), h->hd.lkp + (long long)frame * h->hd.lkp_fstride + h->hd.lv[level].kp_off, (size_t)m * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; ++i) { dst[i].x = (float)(pk[i] & 0xFFF); dst[i].y = (float)((pk[i] >> 12) & 0xFFF); }
  }
  return DRFE_OK;
}

}  // extern "C"
