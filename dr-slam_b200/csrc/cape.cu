// drfe CAPE plane extraction for sm_100a — the work of PlaneDetection_CAPE::runPlaneDetection
// (reference src/PlaneExtractor.cpp:111-191) and CAPE::process (src/CAPE/CAPE.cpp:47-457),
// batched over independent frames.  Built with --fmad=false: every float/double operation
// is individually rounded exactly like the CPU oracle.
//
// Kernels:
//   k_cape_cells   half-warp per grid cell: depth -> XYZ (double math, PlaneExtractor.cpp:
//                  117-127) -> cell-major cloud, the 9 float moment sums in the declared
//                  16-lane tree, missing-data / depth-jump tests and fitPlane with a 3x3
//                  Jacobi eigen-solve in registers (PlaneSeg.cpp:8-142)
//   k_cape_grid    one CTA per frame: normal histogram, seeded region growing as frontier
//                  propagation, plane merging, erode/dilate cell masks (CAPE.cpp:82-291)
//   k_cape_refine  one warp per cell: per-pixel boundary refinement + seg_output in image
//                  layout (CAPE.cpp:294-319, 395-432)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "drfe_internal.h"

namespace drfe {

static const int kMaxPlanes = 255;       // labels are uchar (CAPE.cpp:286)
static const int kHistBins = 20;

struct CellSums;
// MSE, normal-histogram bin (-1: not planar), planar flag and bits 0..3 = the cell can be activated from its left / right /
// upper / lower neighbour
struct CellMeta { float mse; short bin; uint8_t edge, planar; };
struct CapeDev {
  int H, W, cw, ch, ncx, ncy, ncells, npc, B;
  float min_cos, max_merge_dist;
  float fx, fy, cx, cy;
  const float* depth; long long depth_rs, depth_fs;   // input depth (may be null: cloud given)
  const uint16_t* depth16; float depth_factor;        // raw sensor depth: z = (float)u16 * depth_factor (Frame.cc:113-115)
  float* cloud;                 // [B][3][H*W] cell-major
  drfe_plane* cells;            // [B][ncells]
  float* tols;                  // [B][ncells]
  struct CellMeta* cell_meta;   // [B][ncells] what the grid stage reads of every cell (k_cape_edges)
  int* plane_map;               // [B][ncells]
  uint8_t* eroded_map;          // [B][ncells]
  uint32_t* border_vec;         // [B][kMaxPlanes+1][ceil(ncells/32)] bit c of row p: cell c is in mask_diff of final plane p (1-based)
  double* jobacc;               // [B][max_jobs][10] accumulated sums + point count of each grown region
  drfe_plane* jobseg;           // [B][max_jobs] fitted region of each job
  int max_jobs;                 // regions of >= 4 cells a frame can have: ncells/4 + 1
  long long* dbg;               // [B][16] k_cape_grid counters/cycles (diagnostics, may be null)
  int grid_sums_smem;           // k_cape_grid keeps the cells' moment sums in shared memory
  drfe_plane* segs;             // [B][kMaxPlanes+1] scratch: plane_segments
  drfe_plane* planes;           // [B][kMaxPlanes]   plane_segments_final
  float4* plane_eq;             // [B][kMaxPlanes+1] (nx,ny,nz,d) float of final planes (1-based)
  float* plane_maxd;            // [B][kMaxPlanes+1] 9*MSE
  int* nplanes;                 // [B]
  uint8_t* seg;                 // [B][H*W]
  uint8_t* cell_label;          // [B][ncells] label a cell is painted with as a whole (k_cape_refine_plan -> k_cape_paint)
  int2* border_list;            // [B * ncells] {batch-wide cell index, mask of its first 32 planes} of the launch's border cells (k_cape_refine_plan -> k_cape_refine_border)
  int* border_count;            // [1]
  int* status;
  struct CellSums* sums;        // [B][ncells] per-cell moment sums (k_cape_sums -> k_cape_fit)
  // ---- cylinder detection (CylinderSeg.cpp, CAPE.cpp:179-216, 323-393); all null / 0 when it is off
  int cyl;                      // CAPE(..., cylinder_detection, ...)
  float cylK1, cylK2;           // RANSAC iteration bounds, float K of CylinderSeg.cpp:84 / :164
  const uint32_t* rand_tab;     // declared rand() stream (glibc TYPE_3, seed 1), restarted per frame (App. B.9)
  int rand_n;
  double* cyl_scratch;          // [B][6][ncells] projected normals / means of the region being fitted
  struct CylSub* subs;          // [B][max_sub] RANSAC sub-segments in creation order
  int max_sub;
  int* cyl_map;                 // [B][ncells] grid_cylinder_seg_map
  uint8_t* cyl_eroded_map;      // [B][ncells] 50 + k
  struct CylEq* cyl_eq;         // [B][kMaxPlanes+1] refinement parameters of final cylinder k (1-based)
  drfe_cylinder* cyls;          // [B][max_sub] cylinder_segments_final (all cylinders found)
  int* ncyl_found;              // [B]
  int* ncyl_final;              // [B]
  int border_rows;              // rows of border_vec per frame: 256 (planes) or 512 (+ cylinders)
};

// one RANSAC sub-segment of an extruded region: either re-fitted as a plane or kept as a cylinder
struct CylSub {
  drfe_plane ps;                // plane through the inlier cells (CAPE.cpp:186-193)
  double center[3], axis[3], mse;
  float radius, p1[3], p2[3], n12;
  int is_cyl, label;            // label: plane number or cylinder number (1-based), set when numbering
};
struct CylEq { float p2[3], dir[3]; double n12, radius; float maxd; int pad; };

// ---- 3x3 symmetric eigen-solve (cyclic Jacobi).  Mirrors eig3_sym() of the oracle
// operation for operation; only + - * / sqrt fabs, no FMA.
__device__ void eig3_sym(const double in[6], double w[3], double v[3][3]) {
  double a[3][3] = {{in[0], in[1], in[2]}, {in[1], in[3], in[4]}, {in[2], in[4], in[5]}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 24; ++sweep) {
    if (a[0][1] == 0.0 && a[0][2] == 0.0 && a[1][2] == 0.0) break;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int p = (k == 2) ? 1 : 0, q = (k == 0) ? 1 : 2, r = (k == 0) ? 2 : ((k == 1) ? 1 : 0);
      const double apq = a[p][q];
      if (apq == 0.0) continue;
      const double app = a[p][p], aqq = a[q][q];
      const double g = 100.0 * fabs(apq);
      if (sweep > 3 && fabs(app) + g == fabs(app) && fabs(aqq) + g == fabs(aqq)) {
        a[p][q] = a[q][p] = 0.0;
        continue;
      }
      const double h = aqq - app;
      double t;
      if (fabs(h) + g == fabs(h)) {
        t = apq / h;
      } else {
        const double theta = 0.5 * h / apq;
        t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
        if (theta < 0.0) t = -t;
      }
      const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
      a[p][p] = app - t * apq;
      a[q][q] = aqq + t * apq;
      a[p][q] = a[q][p] = 0.0;
      const double arp = a[r][p], arq = a[r][q];
      a[r][p] = a[p][r] = c * arp - s * arq;
      a[r][q] = a[q][r] = s * arp + c * arq;
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const double vp = v[m][p], vq = v[m][q];
        v[m][p] = c * vp - s * vq;
        v[m][q] = s * vp + c * vq;
      }
    }
  }
  int i0 = 0, i1 = 1, i2 = 2;
  const double d[3] = {a[0][0], a[1][1], a[2][2]};
  if (d[i1] < d[i0]) { int t = i0; i0 = i1; i1 = t; }
  if (d[i2] < d[i1]) { int t = i1; i1 = i2; i2 = t; }
  if (d[i1] < d[i0]) { int t = i0; i0 = i1; i1 = t; }
  double vv[3][3];
  const int idx[3] = {i0, i1, i2};
  for (int i = 0; i < 3; ++i) {
    w[i] = d[idx[i]];
    for (int m = 0; m < 3; ++m) vv[m][i] = v[m][idx[i]];
  }
  for (int i = 0; i < 3; ++i)
    for (int m = 0; m < 3; ++m) v[m][i] = vv[m][i];
}

// PlaneSeg::fitPlane (PlaneSeg.cpp:111-142)
__device__ void fit_plane(drfe_plane& s) {
  const double n = (double)s.nr_pts;
  s.mean[0] = s.x_acc / n; s.mean[1] = s.y_acc / n; s.mean[2] = s.z_acc / n;
  const double cov[6] = {s.xx_acc - s.x_acc * s.x_acc / n, s.xy_acc - s.x_acc * s.y_acc / n,
                         s.xz_acc - s.x_acc * s.z_acc / n, s.yy_acc - s.y_acc * s.y_acc / n,
                         s.yz_acc - s.y_acc * s.z_acc / n, s.zz_acc - s.z_acc * s.z_acc / n};
  double w[3], v[3][3];
  eig3_sym(cov, w, v);
  const double v0 = v[0][0], v1 = v[1][0], v2 = v[2][0];
  double d = -(v0 * s.mean[0] + v1 * s.mean[1] + v2 * s.mean[2]);
  if (d > 0) { s.normal[0] = v0; s.normal[1] = v1; s.normal[2] = v2; }
  else { s.normal[0] = -v0; s.normal[1] = -v1; s.normal[2] = -v2; d = -d; }
  s.d = d;
  s.MSE = (float)(w[0] / n);
  s.score = (float)(w[1] / w[0]);
}

__device__ __forceinline__ void expand_seg(drfe_plane& a, const drfe_plane& b) {
  a.x_acc += b.x_acc; a.y_acc += b.y_acc; a.z_acc += b.z_acc;
  a.xx_acc += b.xx_acc; a.yy_acc += b.yy_acc; a.zz_acc += b.zz_acc;
  a.xy_acc += b.xy_acc; a.xz_acc += b.xz_acc; a.yz_acc += b.yz_acc;
  a.nr_pts += b.nr_pts;
}

// ------------------------------------------------------------------ cells
// k_cape_sums: 16 lanes per cell.  Lane l owns the declared accumulator l of each of the 9 sums:
// elements l, l+16, l+32, ... in ascending order (Eigen's two-packet AVX redux, App. B.1),
// then lanes l and l+8 are added, then the 8 -> 4 -> 2 -> 1 halving tree.  The depth values
// of a chunk are loaded before any of them is used so the loads overlap; z is kept in shared
// memory for the two depth-jump scans (PlaneSeg.cpp:36-76), which two lanes run side by side.
// Output per cell: 9 float sums, the valid-point count and the planarity flags so far
// (CellSums); k_cape_fit turns that into the PlaneSeg with one thread per cell.
struct CellSums { float s[9]; int cnt; int planar; float diam; };   // diam: distance between the cell's first and last point

__device__ __forceinline__ float tree16(float v, bool has_extra, unsigned mask, float extra8) {
  float p = v + __shfl_down_sync(mask, v, 8, 16);        // lane[l] + lane[l+8]
  if (has_extra) p = p + extra8;                           // one more aligned packet (lanes 0..7)
  p = p + __shfl_down_sync(mask, p, 4, 16);
  p = p + __shfl_down_sync(mask, p, 2, 16);
  p = p + __shfl_down_sync(mask, p, 1, 16);
  return p;  // valid in lane 0 of the group
}

// ---- depth -> X, Y (PlaneExtractor.cpp:117-127): x = ((double)j - cx) * z / fx in double, stored as float.
// (float)(t / den) without the fp64 division: q' = z * (dcol * RN(1/fx)) carries three roundings, so it is within
// 4 ulps of the correctly rounded quotient q = RN(dcol * z / fx) (the product dcol * z is exact: a small half-integer
// times a float), and float(q') == float(q) unless a float rounding midpoint (double mantissa bits 28..0 ==
// 0x10000000) lies within a few ulps of q' — or the result leaves the float normal range, which a depth inside
// [2^-20, 2^20) and sane intrinsics (XyCtx::sane) rule out.  Those rare cases take the exact division, so every
// kernel that converts a depth (k_cape_sums, k_cape_refine, k_cape_cloud) produces the reference's bits.
struct XyCtx { double fx, fy, rfx, rfy, cx, cy; bool sane; };
__device__ __forceinline__ XyCtx xy_ctx(const CapeDev& P) {
  XyCtx C;
  C.fx = (double)P.fx; C.fy = (double)P.fy; C.cx = (double)P.cx; C.cy = (double)P.cy;
  C.rfx = 1.0 / C.fx; C.rfy = 1.0 / C.fy;
  C.sane = P.fx >= 0x1p-10f && P.fx <= 0x1p20f && P.fy >= 0x1p-10f && P.fy <= 0x1p20f && fabsf(P.cx) <= 32768.f && fabsf(P.cy) <= 32768.f &&
           P.W <= 32768 && P.H <= 32768;
  return C;
}
// true when float(qx) or float(qy) might differ from the float of the exact quotient: within 8 double ulps of a float
// rounding midpoint, or z outside {0} u [2^-20, 2^20)
__device__ __forceinline__ bool xy_needs_exact(float z, double qx, double qy) {
  const uint32_t a = ((uint32_t)__double2loint(qx) & 0x1FFFFFFFu) - 0x0FFFFFF8u;
  const uint32_t b = ((uint32_t)__double2loint(qy) & 0x1FFFFFFFu) - 0x0FFFFFF8u;
  const uint32_t zb = __float_as_uint(z);
  return min(a, b) <= 16u || ((zb - 0x35800000u) >= 0x14000000u && zb != 0u);
}
__device__ __noinline__ void xy_exact(float z, double dcol, double drow, double fx, double fy, float& x, float& y) {
  const double zd = (double)z;
  x = (float)(dcol * zd / fx);
  y = (float)(drow * zd / fy);
}
// one pixel, column offset dcol = (double)j - cx and row offset drow = (double)i - cy (both exact)
__device__ __forceinline__ void xy_from_depth(const XyCtx& C, float z, double dcol, double drow, float& x, float& y) {
  const double zd = (double)z;
  const double qx = zd * (dcol * C.rfx), qy = zd * (drow * C.rfy);
  if (!C.sane || xy_needs_exact(z, qx, qy)) xy_exact(z, dcol, drow, C.fx, C.fy, x, y);
  else { x = (float)qx; y = (float)qy; }
}

static const int kSumsThreads = 128, kSumsChunk = 5;
static const int kSumsCells = 4;     // consecutive cells per 16-lane group (depth of cell c+2 is in flight while c is summed)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// MODE 0: the cell-major cloud is given; 1: float depth image; 2: raw 16-bit depth image scaled by depth_factor
// CELL: compile-time cell edge (square cells of 20 or 10 px, the two sizes DR-SLAM's yaml files use) so that
// the per-element index arithmetic and bounds tests fold away; 0 = any cell size, read from the descriptor.
// A 16-lane group walks kSumsCells consecutive cells; with a float depth image whose cell rows are 16-byte
// aligned, the depth of the next two cells is copied into a two-slot shared-memory ring by cp.async (LDGSTS)
// while the current cell is summed, so only the group's first load latency is exposed.
// The cloud itself is NOT written (it was 900 MB per 256-frame step, two thirds of the kernel's DRAM traffic, for a
// refinement stage that reads a third of it): k_cape_refine converts the depth of its border cells again, and
// k_cape_cloud materialises the cell-major cloud when a caller asks for it (drfe_cape_get_cloud / plane_points).
template <int MODE, int CELL>
__global__ void __launch_bounds__(kSumsThreads, 6) k_cape_sums(const CapeDev* __restrict__ Pp, int f0, int nframes) {
  DRFE_GRID_DEP();
  constexpr bool FROM_DEPTH = MODE != 0;
  extern __shared__ __align__(16) float s_zall[];           // [8 groups][2 slots][npc]
  const CapeDev& P = *Pp;
  const int grp = (blockIdx.x * kSumsThreads + threadIdx.x) >> 4;
  const int l = threadIdx.x & 15;
  // full-warp mask: both 16-lane groups of a warp run the same shuffles (width 16); a group that
  // has returned is simply absent.  (A per-group runtime mask makes the compiler serialise them.)
  const unsigned mask = 0xFFFFFFFFu;
  const int ncells = P.ncells;
  const int total = nframes * ncells;
  const int first = grp * kSumsCells;                       // first cell of this group within the launch
  if (first >= total) return;                               // whole 16-lane groups exit together
  const int ncell_here = min(kSumsCells, total - first);
  const int npc = CELL ? CELL * CELL : P.npc, cw = CELL ? CELL : P.cw, ch = CELL ? CELL : P.ch;
  float* s_ring = s_zall + (threadIdx.x >> 4) * 2 * npc;
  const long long N = (long long)P.H * P.W;
  const int body = (npc / 16) * 16;
  const bool has_extra = npc - body >= 8;
  const int full8 = has_extra ? body + 8 : body;
  const int drs = (int)P.depth_rs;
  // every cell row starts on a 16-byte boundary: rows can be moved as float4 (MODE 1) / 4 x u16 (MODE 2)
  const bool vec1 = MODE == 1 && (cw & 3) == 0 && (drs & 3) == 0 && (P.depth_fs & 3) == 0 && ((reinterpret_cast<uintptr_t>(P.depth) & 15) == 0);
  const bool vec2 = MODE == 2 && (cw & 3) == 0 && (drs & 3) == 0 && (P.depth_fs & 3) == 0 && ((reinterpret_cast<uintptr_t>(P.depth16) & 7) == 0);
  // asynchronous copy of cell (first + c)'s depth into ring slot c & 1 (MODE 1, aligned)
  auto prefetch = [&](int c) {
    const int gid = first + c + f0 * ncells;
    const int f = gid / ncells, cell = gid - f * ncells;
    const int cr = cell / P.ncx, cc = cell - cr * P.ncx;
    const float* __restrict__ dsrc = P.depth + (long long)f * P.depth_fs + (long long)(cr * ch) * P.depth_rs + cc * cw;
    float* dst = s_ring + (c & 1) * npc;
    const int q4 = cw >> 2, n4 = npc >> 2;                   // float4 per cell row / per cell
    for (int j = l; j < n4; j += 16) {
      const int r = j / q4, c4 = j - r * q4;
      cp_async16(dst + 4 * j, dsrc + r * drs + 4 * c4);      // cell-local index = r*cw + 4*c4 = 4*j
    }
    cp_async_commit();
  };
  if (vec1) { prefetch(0); if (ncell_here > 1) prefetch(1); }
  const XyCtx C = xy_ctx(P);
  const int step_r = 16 / cw, step_c = 16 - step_r * cw;    // element i + 16
  const double dstep_c = (double)step_c, dstep_r = (double)step_r, dcw = (double)cw;
  const int last_lane = (npc - 1) & 15;                      // the lane that owns the cell's last element
  for (int c = 0; c < ncell_here; ++c) {
    const int gid = first + c + f0 * ncells;                  // global cell index over the batch
    const int f = gid / ncells, cell = gid - f * ncells;
    float* s_z = s_ring + (c & 1) * npc;
    const float* __restrict__ CX = MODE == 0 ? P.cloud + (long long)f * 3 * N + (long long)cell * npc : nullptr;
    const float* __restrict__ CY = CX + N;
    const float* __restrict__ CZ = CY + N;
    float ax = 0, ay = 0, az = 0, axx = 0, ayy = 0, azz = 0, axy = 0, axz = 0, ayz = 0;
    float ex = 0, ey = 0, ez = 0, exx = 0, eyy = 0, ezz = 0, exy = 0, exz = 0, eyz = 0;  // extra packet
    float x_first = 0, y_first = 0, z_first = 0, x_last = 0, y_last = 0, z_last = 0;     // elements 0 and npc - 1 (the cell's diameter)
    int cnt = 0;
    const int cr = cell / P.ncx, cc = cell - cr * P.ncx;
    // ---- stage z in shared memory (it is also what the depth-jump scans read)
    if (MODE == 1) {
      if (vec1) {
        if (c + 1 < ncell_here) cp_async_wait<1>(); else cp_async_wait<0>();
      } else {
        const float* __restrict__ dsrc = P.depth + (long long)f * P.depth_fs + (long long)(cr * ch) * P.depth_rs + cc * cw;
        for (int i = l; i < npc; i += 16) {
          const int r = i / cw, cl = i - r * cw;
          s_z[i] = __ldg(dsrc + r * drs + cl);
        }
      }
    } else if (MODE == 2) {
      // imDepth.convertTo(imDepth, CV_32F, mDepthMapFactor) (Frame.cc:113-115): float(u16) * float(factor)
      const float fac = P.depth_factor;
      const uint16_t* __restrict__ dsrc = P.depth16 + (long long)f * P.depth_fs + (long long)(cr * ch) * P.depth_rs + cc * cw;
      if (vec2) {
        const int q4 = cw >> 2, n4 = npc >> 2;
        for (int j = l; j < n4; j += 16) {
          const int r = j / q4, c4 = j - r * q4;
          const uint2 v = __ldg(reinterpret_cast<const uint2*>(dsrc + r * drs) + c4);
          *reinterpret_cast<float4*>(s_z + 4 * j) = make_float4((float)(v.x & 0xFFFFu) * fac, (float)(v.x >> 16) * fac,
                                                                (float)(v.y & 0xFFFFu) * fac, (float)(v.y >> 16) * fac);
        }
      } else {
        for (int i = l; i < npc; i += 16) {
          const int r = i / cw, cl = i - r * cw;
          s_z[i] = (float)__ldg(dsrc + r * drs + cl) * fac;
        }
      }
    } else {
      for (int i = l; i < npc; i += 16) s_z[i] = CZ[i];
    }
    __syncwarp(mask);
    // (double)j - cx and (double)i - cy of the cell's first pixel: differences of small (half-)integers, exact
    const double col0 = (double)(cc * cw) - C.cx, row0 = (double)(cr * ch) - C.cy;
    // the moment sums of one element: FIRST = the lane's first element (it starts the accumulators), CHECKED = bounds tests needed
    auto accumulate = [&](int i, float x, float y, float z, auto first_tag, auto checked_tag) {
      constexpr bool FIRST = decltype(first_tag)::value, CHECKED = decltype(checked_tag)::value;
      cnt += (z > 0.f);
      if (!CHECKED || i < body) {
        if (CHECKED ? i < 16 : FIRST) { ax = x; ay = y; az = z; axx = x * x; ayy = y * y; azz = z * z; axy = x * y; axz = x * z; ayz = y * z; }
        else { ax = ax + x; ay = ay + y; az = az + z; axx = axx + x * x; ayy = ayy + y * y; azz = azz + z * z;
               axy = axy + x * y; axz = axz + x * z; ayz = ayz + y * z; }
      } else if (i < full8) {
        ex = x; ey = y; ez = z; exx = x * x; eyy = y * y; ezz = z * z; exy = x * y; exz = x * z; eyz = y * z;
      }
    };
    if (CELL == 20 && FROM_DEPTH) {
      // 400 = 5 chunks x 5 x 16 elements and 5 x 16 elements = 4 whole cell rows: element l + 16 u + 80 k of a lane sits in
      // column lc[u] (the same for every chunk k) and row lr[u] + 4 k, so the column factors are five loop invariants and
      // the row offsets advance by an exact 4.0 per chunk; every lane owns exactly 25 elements, no bounds tests
      double kx[kSumsChunk], dr[kSumsChunk];
#pragma unroll
      for (int u = 0; u < kSumsChunk; ++u) {
        const int i = l + 16 * u, lr = i / 20, lc = i - 20 * lr;
        kx[u] = (col0 + (double)lc) * C.rfx;
        dr[u] = row0 + (double)lr;
      }
      auto chunk = [&](int i0, auto first_tag) {
#pragma unroll
        for (int u = 0; u < kSumsChunk; ++u) {
          const int i = i0 + 16 * u;
          const float z = s_z[i];
          const double zd = (double)z;
          const double qx = zd * kx[u], qy = zd * (dr[u] * C.rfy);
          float x, y;
          if (!C.sane || xy_needs_exact(z, qx, qy)) {             // rare, out of line
            const int lr = i / 20, lc = i - 20 * lr;
            xy_exact(z, col0 + (double)lc, row0 + (double)lr, C.fx, C.fy, x, y);
          } else { x = (float)qx; y = (float)qy; }
          dr[u] += 4.0;
          if (decltype(first_tag)::value && u == 0) accumulate(i, x, y, z, std::true_type(), std::false_type());
          else accumulate(i, x, y, z, std::false_type(), std::false_type());
        }
      };
      chunk(l, std::true_type());
#pragma unroll 1
      for (int i0 = l + 16 * kSumsChunk; i0 < 400; i0 += 16 * kSumsChunk) chunk(i0, std::false_type());
      if (l == 0) {                                             // the cell's first and last point once more, for its diameter
        z_first = s_z[0]; z_last = s_z[399];
        xy_from_depth(C, z_first, col0, row0, x_first, y_first);
        xy_from_depth(C, z_last, col0 + 19.0, row0 + 19.0, x_last, y_last);
      }
    } else {
      int lr = l / cw, lc = l - lr * cw;                        // (row, column) of element i inside the cell
      // (double)j - cx and (double)i - cy of the element, stepped along with (lr, lc): sums of small half-integers, exact
      double dcol = col0 + (double)lc, drow = row0 + (double)lr;
      auto element = [&](int i, auto first_tag, auto checked_tag) {
        constexpr bool CHECKED = decltype(checked_tag)::value;
        if (!CHECKED || i < npc) {
          float x, y;
          const float z = s_z[i];
          if (FROM_DEPTH) xy_from_depth(C, z, dcol, drow, x, y);
          else { x = CX[i]; y = CY[i]; }
          if (i == 0) { x_first = x; y_first = y; z_first = z; }
          if (i == npc - 1) { x_last = x; y_last = y; z_last = z; }
          accumulate(i, x, y, z, first_tag, checked_tag);
        }
        lc += step_c; lr += step_r; dcol += dstep_c; drow += dstep_r;
        if (lc >= cw) { lc -= cw; ++lr; dcol -= dcw; drow += 1.0; }
      };
      if (CELL == 20) {
        element(l, std::true_type(), std::false_type());
#pragma unroll
        for (int u = 1; u < kSumsChunk; ++u) element(l + 16 * u, std::false_type(), std::false_type());
#pragma unroll 1
        for (int i0 = l + 16 * kSumsChunk; i0 < 400; i0 += 16 * kSumsChunk) {
#pragma unroll
          for (int u = 0; u < kSumsChunk; ++u) element(i0 + 16 * u, std::false_type(), std::false_type());
        }
      } else {
        for (int i0 = l; i0 < npc; i0 += 16 * kSumsChunk) {
#pragma unroll
          for (int u = 0; u < kSumsChunk; ++u) element(i0 + 16 * u, std::false_type(), std::true_type());
        }
      }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) cnt += __shfl_down_sync(mask, cnt, o, 16);
    // (body is a multiple of 16, so element body+j was handled by lane j: the extra packet is in place)
    float sx = tree16(ax, has_extra, mask, ex), sy = tree16(ay, has_extra, mask, ey), sz = tree16(az, has_extra, mask, ez),
          sxx = tree16(axx, has_extra, mask, exx), syy = tree16(ayy, has_extra, mask, eyy),
          szz = tree16(azz, has_extra, mask, ezz), sxy = tree16(axy, has_extra, mask, exy),
          sxz = tree16(axz, has_extra, mask, exz), syz = tree16(ayz, has_extra, mask, eyz);
    // the cell's last element, for its diameter (CAPE.cpp:69-73 reads rows 0 and npc - 1 of the cell's block)
    if (!(CELL == 20 && FROM_DEPTH)) {
      x_last = __shfl_sync(mask, x_last, last_lane, 16); y_last = __shfl_sync(mask, y_last, last_lane, 16); z_last = __shfl_sync(mask, z_last, last_lane, 16);
    }
    __syncwarp(mask);                                         // s_z is complete
    // depth-jump scans through the middle row and the middle column (PlaneSeg.cpp:36-76): z_last follows the valid
    // depths as long as consecutive valid ones differ by < 100; a valid depth further away is a jump and leaves z_last
    // alone.  Without jumps z_last is simply the previous valid depth, so the 16 lanes first test every element
    // against its previous valid one (found in a ballot mask); only a group that sees a violation replays the scan
    // sequentially (lane 0: row, lane 1: column) to count the jumps.
    int jumps = 0;
    {
      const int chh = npc / cw;
      const int shift = threadIdx.x & 16;                     // this group's half of a ballot
      bool fail = false;
      if (cw <= 32 && chh <= 32) {
#pragma unroll
        for (int sc = 0; sc < 2; ++sc) {
          const int base = sc == 0 ? cw * (chh / 2) : cw / 2, step = sc == 0 ? 1 : cw, n = sc == 0 ? cw : chh;
          const float za = l < n ? s_z[base + l * step] : 0.f, zb = l + 16 < n ? s_z[base + (l + 16) * step] : 0.f;
          const uint32_t va = (__ballot_sync(mask, za > 0.f) >> shift) & 0xFFFFu, vb = (__ballot_sync(mask, zb > 0.f) >> shift) & 0xFFFFu;
          const uint32_t valid = va | (vb << 16);
          const float z_init = fmaxf(s_z[base], s_z[base + step]);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int k = l + 16 * h2;
            const float z = h2 ? zb : za;
            if (k >= 1 && z > 0.f) {
              const uint32_t before = valid & ((1u << k) - 2u);   // valid elements 1 .. k-1
              const float zp = before ? s_z[base + (31 - __clz(before)) * step] : z_init;
              fail |= !(fabsf(z - zp) < 100.0f);                   // == (double)|dz| < 100.0: 100 is a float
            }
          }
        }
      } else {
        fail = true;
      }
      const bool group_fail = ((__ballot_sync(mask, fail) >> shift) & 0xFFFFu) != 0;
      if (group_fail && l < 2) {
        int i, j, step;
        float z_prev;
        if (l == 0) { i = cw * (chh / 2); j = i + cw; step = 1; z_prev = fmaxf(s_z[i], s_z[i + 1]); }
        else { i = cw / 2; j = npc - i; step = cw; z_prev = fmaxf(s_z[i], s_z[i + cw]); }
        i += step;
        while (i < j) {
          const float z = s_z[i];
          if (z > 0 && fabsf(z - z_prev) < 100.0f) z_prev = z;
          else if (z > 0) ++jumps;
          i += step;
        }
      }
    }
    const int jumps_v = __shfl_down_sync(mask, jumps, 1, 16);
    if (l == 0) {
      // scalar tail (Eigen's unaligned end), sequential
      for (int i = full8; i < npc; ++i) {
        float x, y;
        const float z = s_z[i];
        if (FROM_DEPTH) { const int lr = i / cw, lc = i - lr * cw; xy_from_depth(C, z, col0 + (double)lc, row0 + (double)lr, x, y); }
        else { x = CX[i]; y = CY[i]; }
        sx = sx + x; sy = sy + y; sz = sz + z; sxx = sxx + x * x; syy = syy + y * y; szz = szz + z * z;
        sxy = sxy + x * y; sxz = sxz + x * z; syz = syz + y * z;
      }
      CellSums o;
      o.s[0] = sx; o.s[1] = sy; o.s[2] = sz; o.s[3] = sxx; o.s[4] = syy; o.s[5] = szz; o.s[6] = sxy; o.s[7] = sxz; o.s[8] = syz;
      o.cnt = cnt;
      o.planar = (cnt >= npc / 2 && jumps <= 1 && jumps_v <= 1) ? 1 : 0;
      // cloud_array.block(cell)[npc - 1] - [0], the cell's diameter (CAPE.cpp:70-72)
      const float dx = x_last - x_first, dy = y_last - y_first, dz = z_last - z_first;
      o.diam = sqrtf(dx * dx + dy * dy + dz * dz);
      float4* dst = reinterpret_cast<float4*>(P.sums + (long long)gid);
      const float4* srcv = reinterpret_cast<const float4*>(&o);
      dst[0] = srcv[0]; dst[1] = srcv[1]; dst[2] = srcv[2];
    }
    // the scans are done with this slot (the shuffle above is after them in every lane): refill it
    __syncwarp(mask);
    if (vec1 && c + 2 < ncell_here) prefetch(c + 2);
  }
}

// The cell-major cloud of organizePointCloudByCell (PlaneExtractor.cpp:80-99, 112-127), on demand: thread = 4 pixels
// of an image row (cells are at least 2 wide; a quad may straddle cells, every pixel is placed on its own).
template <int MODE>
__global__ void __launch_bounds__(256) k_cape_cloud(const CapeDev* __restrict__ Pp, int nframes) {
  const CapeDev& P = *Pp;
  const XyCtx C = xy_ctx(P);
  const int W = P.W, H = P.H, q = (W + 3) >> 2;
  const long long N = (long long)H * W;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)nframes * H * q) return;
  const int f = (int)(t / ((long long)H * q));
  const int rem = (int)(t - (long long)f * H * q);
  const int r = rem / q, c0 = 4 * (rem - r * q);
  float* CX = P.cloud + (long long)f * 3 * N;
  if (r >= P.ncy * P.ch) return;                                // below the last full cell row: not part of any cell
  const int cell_r = r / P.ch, lr = r - cell_r * P.ch;
  const double drow = (double)r - C.cy;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + k;
    if (c >= P.ncx * P.cw) break;
    float z;
    if (MODE == 1) z = __ldg(P.depth + (long long)f * P.depth_fs + (long long)r * P.depth_rs + c);
    else z = (float)__ldg(P.depth16 + (long long)f * P.depth_fs + (long long)r * P.depth_rs + c) * P.depth_factor;
    float x, y;
    xy_from_depth(C, z, (double)c - C.cx, drow, x, y);
    const int cell_c = c / P.cw, lc = c - cell_c * P.cw;
    const long long idx = (long long)(cell_r * P.ncx + cell_c) * P.npc + lr * P.cw + lc;
    CX[idx] = x; CX[idx + N] = y; CX[idx + 2 * N] = z;
  }
}

// k_cape_fit: one thread per cell — the rest of PlaneSeg::PlaneSeg (PlaneSeg.cpp:78-94): sums
// widened to double, fitPlane, the depth-dependent MSE test, and the cell's merge tolerance
// (CAPE.cpp:69-73).
__global__ void __launch_bounds__(128) k_cape_fit(const CapeDev* __restrict__ Pp, int f0, int nframes) {
  DRFE_GRID_DEP();
  // the 152-byte PlaneSeg records of the block's 128 cells leave through shared memory as full 16-byte words of
  // consecutive addresses (a struct store per thread touched 32 lines per instruction)
  __shared__ __align__(16) drfe_plane s_out[128];
  static_assert(sizeof(drfe_plane) % 8 == 0, "drfe_plane is moved as 8-byte words");
  const CapeDev& P = *Pp;
  const int gid0 = blockIdx.x * 128 + threadIdx.x;
  const int total = nframes * P.ncells;
  if (gid0 < total) {
    const int gid = gid0 + f0 * P.ncells;
    
    const int npc = P.npc;
    CellSums in;
    {
      const float4* srcv = reinterpret_cast<const float4*>(P.sums + (long long)gid);
      float4* d = reinterpret_cast<float4*>(&in);
      d[0] = srcv[0]; d[1] = srcv[1]; d[2] = srcv[2];
    }
    drfe_plane s;
    memset(&s, 0, sizeof(s));                        // zero-filled PlaneSeg storage (App. B.2)
    s.min_nr_pts = npc / 2;
    s.nr_pts = in.cnt;
    s.planar = in.planar;
    float tol = 0.f;
    if (s.planar) {
      s.x_acc = in.s[0]; s.y_acc = in.s[1]; s.z_acc = in.s[2]; s.xx_acc = in.s[3]; s.yy_acc = in.s[4]; s.zz_acc = in.s[5];
      s.xy_acc = in.s[6]; s.xz_acc = in.s[7]; s.yz_acc = in.s[8];
      fit_plane(s);
      const double lim = 0.000001425 * s.mean[2] * s.mean[2] + 10.0;   // Params.h:6-7
      if ((double)s.MSE > lim * lim) s.planar = 0;
      if (s.planar) {  // cell_distance_tols (CAPE.cpp:69-73); the cell's diameter comes with the sums
        const float diam = in.diam;
        const float sin_merge = (float)sqrt(1.0 - (double)P.min_cos * (double)P.min_cos);
        const float t = fminf(fmaxf(diam * sin_merge, 20.0f), P.max_merge_dist);
        tol = t * t;
      }
    }
    s_out[threadIdx.x] = s;
    P.tols[gid] = tol;
  }
  __syncthreads();
  {
    const int nvalid = min(128, total - (int)blockIdx.x * 128);
    const int nwords = nvalid * (int)(sizeof(drfe_plane) / 8);
    const uint2* src = reinterpret_cast<const uint2*>(s_out);
    uint2* dst = reinterpret_cast<uint2*>(P.cells + ((long long)blockIdx.x * 128 + (long long)f0 * P.ncells));
    for (int i = threadIdx.x; i < nwords; i += 128) dst[i] = src[i];
  }
}

// k_cape_edges: one thread per cell, after k_cape_fit — what the grid stage needs of every cell besides its plane:
// the bin of the normal histogram (CAPE.cpp:82-101, Histogram.cpp:15-43: double acos / atan2) and whether
// RegionGrowing (CAPE.cpp:485-506) would activate the cell from each of its four neighbours.  Both depend on the
// cell and its neighbours only, so they are computed here by 196 k threads instead of by the grid stage's 128.
__global__ void __launch_bounds__(128) k_cape_edges(const CapeDev* __restrict__ Pp, int f0, int nframes) {
  DRFE_GRID_DEP();
  const CapeDev& P = *Pp;
  const int gid0 = blockIdx.x * 128 + threadIdx.x;
  if (gid0 >= nframes * P.ncells) return;
  const int gid = gid0 + f0 * P.ncells;
  const int f = gid / P.ncells, c = gid - f * P.ncells;
  const int ncx = P.ncx, ncy = P.ncy;
  const int y = c / ncx, x = c - y * ncx;
  const drfe_plane* cells = P.cells + (long long)f * P.ncells;
  const drfe_plane& g = cells[c];
  int b = -1;
  unsigned e = 0;
  if (g.planar != 0) {
    const double min_cos = (double)P.min_cos;
    const double nx = g.normal[0], ny = g.normal[1], nz = g.normal[2];
    const double mx = g.mean[0], my = g.mean[1], mz = g.mean[2];
    const double pn = sqrt(nx * nx + ny * ny);
    const double polar = acos(-nz);
    const int xq = (int)((kHistBins - 1) * (polar - 0.0) / (3.14 - 0.0));
    int yq = 0;
    if (xq > 0) yq = (int)((kHistBins - 1) * (atan2(nx / pn, ny / pn) - (-3.14)) / (3.14 - (-3.14)));
    b = yq * kHistBins + xq;
    // can this cell be activated from neighbour p (RegionGrowing called with p's normal and d)?
    const double tol = (double)P.tols[gid];
    auto edge = [&](int p) -> bool {
      const drfe_plane& a = cells[p];
      const double dist = a.normal[0] * mx + a.normal[1] * my + a.normal[2] * mz + a.d;
      return !(a.normal[0] * nx + a.normal[1] * ny + a.normal[2] * nz < min_cos || dist * dist > tol);
    };
    if (x > 0 && edge(c - 1)) e |= 1u;
    if (x + 1 < ncx && edge(c + 1)) e |= 2u;
    if (y > 0 && edge(c - ncx)) e |= 4u;
    if (y + 1 < ncy && edge(c + ncx)) e |= 8u;
  }
  CellMeta m;
  m.mse = g.MSE; m.bin = (short)b; m.edge = (uint8_t)e; m.planar = g.planar != 0 ? 1 : 0;
  *reinterpret_cast<uint2*>(P.cell_meta + gid) = *reinterpret_cast<const uint2*>(&m);
}

// ------------------------------------------------------------------ grid stage
// CAPE::process between the per-cell fits and the per-pixel refinement (CAPE.cpp:82-291), one
// CTA per frame.  The cell grid is handled as bit vectors (bit c = cell c, 32 cells per word):
//  * RegionGrowing's acceptance test of cell c against an activated 4-neighbour p
//    (CAPE.cpp:485-506) depends only on the pair (p, c), so the four directed edge masks
//    FL/FR/FU/FD are computed once per frame; growing a region is then reachability over those
//    masks: A |= ((A<<1)&FL | (A>>1)&FR | (A<<ncx)&FU | (A>>ncx)&FD) & U until A stops changing —
//    the same set the recursion activates (order-free), done by one warp on a few words.
//  * erode (3x3 cross) / dilate (3x3 square) of the per-plane cell masks (:270, :282) are
//    shifts and ANDs / ORs of the mask vector.
// The seed loop itself is sequential and runs on warp 0; the other warps join for the
// per-cell set-up and the per-plane mask stage.
__device__ __forceinline__ uint32_t bv_get(const uint32_t* V, int nw, int w) { return (w >= 0 && w < nw) ? V[w] : 0u; }
// word w of V shifted up by k cells (bit c of the result = bit c-k of V), zeros shifted in
__device__ __forceinline__ uint32_t bv_shl(const uint32_t* V, int nw, int w, int k) {
  const int q = k >> 5, r = k & 31;
  const uint32_t lo = bv_get(V, nw, w - q);
  return r ? ((lo << r) | (bv_get(V, nw, w - q - 1) >> (32 - r))) : lo;
}
// word w of V shifted down by k cells (bit c of the result = bit c+k of V)
__device__ __forceinline__ uint32_t bv_shr(const uint32_t* V, int nw, int w, int k) {
  const int q = k >> 5, r = k & 31;
  const uint32_t hi = bv_get(V, nw, w + q);
  return r ? ((hi >> r) | (bv_get(V, nw, w + q + 1) << (32 - r))) : hi;
}


// ------------------------------------------------------------------ cylinders
// CylinderSeg::CylinderSeg (CylinderSeg.cpp:7-247) for one grown region, run by one warp.  The
// reference is a chain of sequential decisions (RANSAC draws from one rand() stream, MSAC sums
// in ascending cell order); per-cell work (projection, hypothesis distances) is spread over the
// lanes, every ordered double sum is done by one lane in the reference's order, and the small
// 3-vector algebra is replicated on all lanes.  The operation order of every sum is the reference's (ascending cell order).
struct CylCtx {
  const drfe_plane* cells;      // Grid of this frame
  const float* sums;            // smem [nc][9] or null
  const int* npts;              // smem [nc]
  const unsigned short* jobid;  // smem [nc]
  int nc;
  int* l2g;                     // smem [nc] local2global_map
  int* ids_left;                // smem [nc]
  int* inl;                     // smem [nc] inlier list of the accepted hypothesis
  unsigned char* flag;          // smem [nc] bit0: ids_left_mask, bit1: I, bit2: I_final
  double* D;                    // smem [nc]
  double* sN;                   // global [3][nc] projected, normalised normals
  double* sP;                   // global [3][nc] projected means
  const uint32_t* rand_tab; int rand_n;
  CylSub* subs; int max_sub;
  unsigned short* subid;        // smem [nc] sub-segment of each cell (0xFFFF: none)
  float K1, K2;
  int* status;
};
__device__ __forceinline__ double dot3(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

// (forced inline: a real call inside k_cape_grid slows the whole kernel — measured 3.4x on the seed loop —
// presumably because the shared-memory pointers then travel as generic addresses through the ABI)
__device__ __forceinline__ void cyl_job(const CylCtx& X, int j, int m, const drfe_plane& seedcell, int& rpos, int& nsub) {
  const unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31, nc = X.nc;
  const double thr = 0.0225;                                 // cylinder_RANSAC_sqr_max_dist (Params.h:9)
  // local2global_map: the region's cells in ascending order
  {
    int base = 0;
    for (int c0 = 0; c0 < nc; c0 += 32) {
      const int c = c0 + lane;
      const bool is = c < nc && X.jobid[c] == j;
      const unsigned bal = __ballot_sync(FULL, is);
      if (is) X.l2g[base + __popc(bal & ((1u << lane) - 1))] = c;
      base += __popc(bal);
    }
  }
  __syncwarp();
  for (int q = lane; q < m; q += 32) {
    const drfe_plane& g = X.cells[X.l2g[q]];
#pragma unroll
    for (int k = 0; k < 3; ++k) { X.sN[k * nc + q] = g.normal[k]; X.sP[k * nc + q] = g.mean[k]; }
  }
  __syncwarp();
  // cov = [N -N][N -N]^T / (2m - 1): six ordered sums, one per lane
  double c6[6];
  {
    double sacc = 0;
    if (lane < 6) {
      const int a = lane < 3 ? 0 : (lane < 5 ? 1 : 2), b = lane < 3 ? lane : (lane < 5 ? lane - 2 : 2);
      for (int q = 0; q < m; ++q) sacc += X.sN[a * nc + q] * X.sN[b * nc + q];
      sacc = (2.0 * sacc) / (double)(2 * m - 1);
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) c6[e] = __shfl_sync(FULL, sacc, e);
  }
  double S[3], V[3][3];
  eig3_sym(c6, S, V);
  if (S[2] / S[0] < 100.0) return;                           // Checkpoint 1 (CylinderSeg.cpp:53)
  const double vec[3] = {V[0][0], V[1][0], V[2][0]};
  for (int q = lane; q < m; q += 32) {
    double Pq[3] = {X.sP[q], X.sP[nc + q], X.sP[2 * nc + q]}, Nq[3] = {X.sN[q], X.sN[nc + q], X.sN[2 * nc + q]};
    const double pd = dot3(vec, Pq), nd = dot3(vec, Nq);
#pragma unroll
    for (int k = 0; k < 3; ++k) { Pq[k] = Pq[k] - pd * vec[k]; Nq[k] = Nq[k] - nd * vec[k]; }
    const double nn = sqrt((Nq[0] * Nq[0] + Nq[1] * Nq[1]) + Nq[2] * Nq[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) { X.sP[k * nc + q] = Pq[k]; X.sN[k * nc + q] = Nq[k] / nn; }
    X.ids_left[q] = q;
    X.flag[q] = 1;
  }
  __syncwarp();
  float K = X.K1;
  int m_left = m;
  while (m_left > 5 && (double)m_left > 0.1 * (double)m) {    // sequential RANSAC (:94)
    double min_hyp = thr * (double)m_left;
    const int accepted = (int)(0.9 * (double)m_left);
    int max_inl = 0;
    for (int q = lane; q < m; q += 32) X.flag[q] &= 1;
    __syncwarp();
    // The K hypotheses of a run are independent up to the accept / early-break logic and their rand() draws
    // are known in advance (hypothesis k uses draws 3k .. 3k+2), so 32 of them are evaluated at once, one per
    // lane: every lane walks the remaining cells in ascending order and accumulates ITS hypothesis's MSAC sum in
    // the reference's order (:137-148; all lanes read the same cell at the same time, the loads broadcast).
    // The accept test is then replayed over the lanes in k order; hypotheses after an early break were
    // speculative and their draws are not consumed.
    double best_r = 0.0, best_c[3] = {0.0, 0.0, 0.0};
    bool found = false, stop = false;
    int k = 0;
    while (!stop && (float)k < K) {
      int nh = 0;
      while (nh < 32 && (float)(k + nh) < K) ++nh;
      if (rpos + 3 * nh > X.rand_n) { if (lane == 0) atomicOr(X.status, 8); return; }
      const int lk = lane < nh ? lane : 0;                    // idle lanes shadow hypothesis 0
      const int id1 = X.ids_left[(int)(X.rand_tab[rpos + 3 * lk] % (unsigned)m_left)];
      const int id2 = X.ids_left[(int)(X.rand_tab[rpos + 3 * lk + 1] % (unsigned)m_left)];
      const int id3 = X.ids_left[(int)(X.rand_tab[rpos + 3 * lk + 2] % (unsigned)m_left)];
      double e1[3], e2[3], t[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double n1 = X.sN[c * nc + id1], n2 = X.sN[c * nc + id2], n3 = X.sN[c * nc + id3];
        const double p1 = X.sP[c * nc + id1], p2 = X.sP[c * nc + id2], p3 = X.sP[c * nc + id3];
        e1[c] = (n1 + n2) + n3;
        e2[c] = (p1 + p2) + p3;
        t[c] = (n1 * p1 + n2 * p2) + n3 * p3;
      }
      const double a = 1.0 - dot3(e1, e1) / 9.0;
      const double b = ((t[0] + t[1]) + t[2]) / 3.0 - dot3(e1, e2) / 9.0;
      const double r = b / a;
      double center[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) center[c] = (e2[c] - r * e1[c]) / 3.0;
      const double rr = r * r;
      double dist = 0.0;
      int inl = 0;
      for (int tq = 0; tq < m_left; ++tq) {
        const int i = X.ids_left[tq];
        const double x = (X.sP[i] - r * X.sN[i]) - center[0], y = (X.sP[nc + i] - r * X.sN[nc + i]) - center[1],
                     z = (X.sP[2 * nc + i] - r * X.sN[2 * nc + i]) - center[2];
        const double d = ((x * x + y * y) + z * z) / rr;
        if (d < thr) { ++inl; dist += d; }
        else dist += thr;
      }
      int used = nh;
      for (int j = 0; j < nh; ++j) {
        const double dj = __shfl_sync(FULL, dist, j);
        if (dj < min_hyp) {
          min_hyp = dj;
          max_inl = __shfl_sync(FULL, inl, j);
          best_r = __shfl_sync(FULL, r, j);
#pragma unroll
          for (int c = 0; c < 3; ++c) best_c[c] = __shfl_sync(FULL, center[c], j);
          found = true;
          if (max_inl > accepted) { used = j + 1; stop = true; break; }
        }
      }
      rpos += 3 * used;
      k += used;
    }
    if (found) {
      // I_final of the accepted hypothesis (:150-156): the same distances, recomputed
      const double rr = best_r * best_r;
      for (int tq = lane; tq < m_left; tq += 32) {
        const int i = X.ids_left[tq];
        const double x = (X.sP[i] - best_r * X.sN[i]) - best_c[0], y = (X.sP[nc + i] - best_r * X.sN[nc + i]) - best_c[1],
                     z = (X.sP[2 * nc + i] - best_r * X.sN[2 * nc + i]) - best_c[2];
        const double d = ((x * x + y * y) + z * z) / rr;
        if (d < thr) X.flag[i] |= 4;
      }
    }
    __syncwarp();
    if (max_inl < 6) break;                                   // Checkpoint 2 (:160)
    K = X.K2;
    // inlier list; remove the inliers from the remaining cells (:167-176)
    {
      int nin = 0, nleft = 0;
      for (int q0 = 0; q0 < m; q0 += 32) {
        const int q = q0 + lane;
        const unsigned char f = q < m ? X.flag[q] : 0;
        const bool fin = (f & 4) != 0, lf = (f & 1) && !fin;
        const unsigned bi = __ballot_sync(FULL, fin), bl = __ballot_sync(FULL, lf);
        if (fin) X.inl[nin + __popc(bi & ((1u << lane) - 1))] = q;
        if (lf) X.ids_left[nleft + __popc(bl & ((1u << lane) - 1))] = q;
        if (q < m) X.flag[q] = (unsigned char)((lf ? 1 : 0) | (f & 4));
        nin += __popc(bi);
        nleft += __popc(bl);
      }
      m_left = nleft;
    }
    __syncwarp();
    // LLS over all inliers (:178-199): seven ordered sums, one per lane
    double e1[3], e2[3], bsum;
    {
      double acc = 0.0;
      if (lane < 7) {
        for (int tq = 0; tq < max_inl; ++tq) {
          const int i = X.inl[tq];
          double v;
          if (lane < 3) v = X.sN[lane * nc + i];
          else if (lane < 6) v = X.sP[(lane - 3) * nc + i];
          else v = (X.sN[i] * X.sP[i] + X.sN[nc + i] * X.sP[nc + i]) + X.sN[2 * nc + i] * X.sP[2 * nc + i];
          acc += v;
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) { e1[c] = __shfl_sync(FULL, acc, c); e2[c] = __shfl_sync(FULL, acc, 3 + c); }
      bsum = __shfl_sync(FULL, acc, 6);
    }
    const double n2 = (double)(max_inl * max_inl);
    const double a = 1.0 - dot3(e1, e1) / n2;
    bsum = bsum / (double)max_inl;
    bsum = bsum - dot3(e1, e2) / n2;
    double r = bsum / a;
    double center[3], P2d[3], dir[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) center[c] = (e2[c] - r * e1[c]) / (double)max_inl;
    if (r < 0) r = -r;
#pragma unroll
    for (int c = 0; c < 3; ++c) { P2d[c] = center[c] + vec[c]; dir[c] = P2d[c] - center[c]; }
    const double n12 = sqrt(dir[0] * dir[0] + (dir[1] * dir[1] + dir[2] * dir[2]));
    // MSE of the inliers' point-to-axis distances (:206-226)
    for (int tq = lane; tq < max_inl; tq += 32) {
      const int i = X.inl[tq];
      const drfe_plane& g = X.cells[X.l2g[i]];
      const double q0 = g.mean[0] - P2d[0], q1 = g.mean[1] - P2d[1], q2 = g.mean[2] - P2d[2];
      const double cx = dir[1] * q2 - dir[2] * q1, cy = dir[2] * q0 - dir[0] * q2, cz = dir[0] * q1 - dir[1] * q0;
      const double dd = sqrt(cx * cx + (cy * cy + cz * cz)) / n12 - r;
      X.D[i] = dd * dd;
    }
    __syncwarp();
    double mse = 0.0;
    if (lane == 0) {
      for (int tq = 0; tq < max_inl; ++tq) mse += X.D[X.inl[tq]];
      mse = mse / (double)max_inl;
    }
    mse = __shfl_sync(FULL, mse, 0);
    // plane through the same cells: clearPoints + expandSegment in ascending order + fitPlane
    // (CAPE.cpp:186-193); ten ordered sums, one per lane
    double pacc = 0.0;
    int pn = 0;
    if (lane < 9) {
      for (int tq = 0; tq < max_inl; ++tq) {
        const int c = X.l2g[X.inl[tq]];
        pacc += X.sums ? (double)X.sums[c * 9 + lane] : (&X.cells[c].x_acc)[lane];
      }
    } else if (lane == 9) {
      for (int tq = 0; tq < max_inl; ++tq) pn += X.npts[X.l2g[X.inl[tq]]];
    }
    double ps9[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) ps9[k] = __shfl_sync(FULL, pacc, k);
    pn = __shfl_sync(FULL, pn, 9);
    if (nsub >= X.max_sub) { if (lane == 0) atomicOr(X.status, 16); return; }
    int is_cyl = 0;
    if (lane == 0) {
      CylSub& o = X.subs[nsub];
      o.ps = seedcell;
      o.ps.x_acc = ps9[0]; o.ps.y_acc = ps9[1]; o.ps.z_acc = ps9[2]; o.ps.xx_acc = ps9[3]; o.ps.yy_acc = ps9[4];
      o.ps.zz_acc = ps9[5]; o.ps.xy_acc = ps9[6]; o.ps.xz_acc = ps9[7]; o.ps.yz_acc = ps9[8];
      o.ps.nr_pts = pn;
      fit_plane(o.ps);
      is_cyl = ((double)o.ps.MSE < mse) ? 0 : 1;             // model selection (CAPE.cpp:195)
      o.is_cyl = is_cyl;
      o.label = 0;
      o.mse = mse;
      o.radius = (float)r;
      o.n12 = (float)n12;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        o.center[c] = center[c]; o.axis[c] = vec[c];
        o.p1[c] = (float)center[c]; o.p2[c] = (float)P2d[c];
      }
    }
    for (int tq = lane; tq < max_inl; tq += 32) X.subid[X.l2g[X.inl[tq]]] = (unsigned short)nsub;
    ++nsub;
    __syncwarp();
  }
}

enum { BV_FL = 0, BV_FR, BV_FU, BV_FD, BV_U, BV_A, BV_B, BV_M, BV_H, BV_C0, BV_CL, BV_R0, BV_RL, BV_VALID, BV_COUNT };

// CYL: compiled with the cylinder stages (cylinder_detection); the plain instantiation carries none of that code
template <int THREADS, bool CYL>
__global__ void __launch_bounds__(THREADS) k_cape_grid(const CapeDev* __restrict__ Pp, int f0) {
  DRFE_GRID_DEP();
  extern __shared__ __align__(16) uint8_t smem[];
  const CapeDev& P = *Pp;
  const int f = blockIdx.x + f0, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nc = P.ncells, ncx = P.ncx, ncy = P.ncy, nw = (nc + 31) >> 5;
  // ---- shared layout
  uint32_t* bv = reinterpret_cast<uint32_t*>(smem);               // [BV_COUNT][nw]
  int* hist = reinterpret_cast<int*>(bv + BV_COUNT * nw);         // [400]
  uint32_t* assoc = reinterpret_cast<uint32_t*>(hist + kHistBins * kHistBins);   // [256][8] plane adjacency bits
  int* list = reinterpret_cast<int*>(assoc + 256 * 8);            // [nc] candidate cells of the chosen bin
  int* npts = list + nc;                                          // [nc] nr_pts of each cell
  float* mse = reinterpret_cast<float*>(npts + nc);               // [nc]
  float* sums = mse + nc;                                         // [nc][9] cell moment sums (floats, exact) if they fit
  short* bin = reinterpret_cast<short*>(sums + (P.grid_sums_smem ? 9 * nc : 0));   // [nc] histogram bin (-1: none)
  unsigned short* jobid = reinterpret_cast<unsigned short*>(bin + nc);   // [nc] job of each grown cell (0xFFFF: none)
  short* job_seed = reinterpret_cast<short*>(jobid + nc);         // [max_jobs] seed cell of each job
  uint8_t* pmap = reinterpret_cast<uint8_t*>(job_seed + P.max_jobs);   // [nc] grid_plane_seg_map (labels 1..255)
  uint8_t* job_label = pmap + nc;                                 // [max_jobs] 0 / plane label of the job
  unsigned short* job_nact = reinterpret_cast<unsigned short*>((reinterpret_cast<uintptr_t>(job_label + P.max_jobs) + 1) & ~(uintptr_t)1);   // [max_jobs] cells of the job
  // cylinder detection only: [max_jobs+1] first sub-segment of each job, RANSAC flags, distances, inlier list
  unsigned short* job_sub0 = job_nact + P.max_jobs;
  unsigned char* cflag = reinterpret_cast<unsigned char*>(job_sub0 + P.max_jobs + 1);
  double* cylD = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(cflag + nc) + 7) & ~(uintptr_t)7);
  int* cinl = reinterpret_cast<int*>(cylD + nc);
  __shared__ int merge[kMaxPlanes + 1];
  __shared__ int s_np, s_njobs, s_ncyl;
  uint32_t* FL = bv + BV_FL * nw; uint32_t* FR = bv + BV_FR * nw; uint32_t* FU = bv + BV_FU * nw; uint32_t* FD = bv + BV_FD * nw;
  uint32_t* U = bv + BV_U * nw; uint32_t* A = bv + BV_A * nw; uint32_t* Bv = bv + BV_B * nw;
  uint32_t* M = bv + BV_M * nw; uint32_t* Hh = bv + BV_H * nw;
  uint32_t* C0 = bv + BV_C0 * nw; uint32_t* CL = bv + BV_CL * nw; uint32_t* R0 = bv + BV_R0 * nw; uint32_t* RL = bv + BV_RL * nw;
  uint32_t* VALID = bv + BV_VALID * nw;

  const drfe_plane* cells = P.cells + (long long)f * nc;
  const float* tols = P.tols + (long long)f * nc;
  drfe_plane* segs = P.segs + (long long)f * (kMaxPlanes + 1);
  // first / last cell that ever fell into each bin: cells only leave a bin, so the candidate scan of a seed can stay
  // inside [first, last] (a plane's cells are spatially compact) and stop once it has found the histogram's count
  __shared__ int s_bfirst[kHistBins * kHistBins], s_blast[kHistBins * kHistBins];
  for (int i = tid; i < kHistBins * kHistBins; i += THREADS) { hist[i] = 0; s_bfirst[i] = 0x7FFFFFFF; s_blast[i] = -1; }
  for (int i = tid; i < 256 * 8; i += THREADS) assoc[i] = 0;
  if (tid == 0) { s_np = 0; s_njobs = 0; }
  __syncthreads();
  // ---- per-cell set-up: histogram bin (CAPE.cpp:82-101, Histogram.cpp:15-43), edge masks
  const double min_cos = (double)P.min_cos;
  for (int c0 = wid * 32; c0 < nw * 32; c0 += THREADS) {
    const int c = c0 + lane;
    const bool valid = c < nc;
    bool planar = false, fl = false, fr = false, fu = false, fd = false;
    int y = 0, x = 0;
    if (valid) {
      y = c / ncx; x = c - y * ncx;
      // everything the loop below needs of a cell comes from two compact records: CellMeta (k_cape_edges) and the
      // moment sums of k_cape_sums, which are the PlaneSeg sums of every planar cell (floats, widened exactly)
      CellMeta meta;
      *reinterpret_cast<uint2*>(&meta) = *reinterpret_cast<const uint2*>(P.cell_meta + (long long)f * nc + c);
      const float4* sv = reinterpret_cast<const float4*>(P.sums + (long long)f * nc + c);
      const float4 s0 = sv[0], s1 = sv[1], s2 = sv[2];          // s[0..8], cnt, planar-so-far, pad
      planar = meta.planar != 0;
      mse[c] = meta.mse;
      npts[c] = __float_as_int(s2.y);
      pmap[c] = 0;
      jobid[c] = 0xFFFF;
      if (P.grid_sums_smem) {
        float* d = sums + c * 9;
        d[0] = s0.x; d[1] = s0.y; d[2] = s0.z; d[3] = s0.w; d[4] = s1.x; d[5] = s1.y; d[6] = s1.z; d[7] = s1.w; d[8] = s2.x;
      }
      const int b = meta.bin;
      if (planar) {
        const unsigned e = meta.edge;
        atomicAdd(&hist[b], 1);
        atomicMin(&s_bfirst[b], c);
        atomicMax(&s_blast[b], c);
        fl = e & 1u; fr = e & 2u; fu = e & 4u; fd = e & 8u;
      }
      bin[c] = (short)b;
    }
    const int w = c0 >> 5;
    const unsigned b_fl = __ballot_sync(0xFFFFFFFFu, fl), b_fr = __ballot_sync(0xFFFFFFFFu, fr),
                   b_fu = __ballot_sync(0xFFFFFFFFu, fu), b_fd = __ballot_sync(0xFFFFFFFFu, fd),
                   b_u = __ballot_sync(0xFFFFFFFFu, planar), b_v = __ballot_sync(0xFFFFFFFFu, valid),
                   b_c0 = __ballot_sync(0xFFFFFFFFu, valid && x == 0), b_cl = __ballot_sync(0xFFFFFFFFu, valid && x == ncx - 1),
                   b_r0 = __ballot_sync(0xFFFFFFFFu, valid && y == 0), b_rl = __ballot_sync(0xFFFFFFFFu, valid && y == ncy - 1);
    if (lane == 0) {
      FL[w] = b_fl; FR[w] = b_fr; FU[w] = b_fu; FD[w] = b_fd; U[w] = b_u; VALID[w] = b_v;
      C0[w] = b_c0; CL[w] = b_cl; R0[w] = b_r0; RL[w] = b_rl; A[w] = 0; Bv[w] = 0;
    }
  }
  __syncthreads();

  // ---- seeded region growing (CAPE.cpp:114-218).  Warp 0 runs the sequential part: pick the
  // seed, grow the region, take its cells out of the histogram.  What happens to a grown region
  // afterwards (accumulate its cells' sums, fit, score > 100 ?) does not influence the following
  // seeds, so it is recorded as a "job" (jobid[c] = job of cell c) and all jobs are
  // accumulated and fitted in parallel after the loop; labels are then handed out in job order.
  long long t_setup = clock64();
  if (wid == 0) {
    long long c_arg = 0, c_scan = 0, c_bfs = 0, c_acc = 0, n_seeds = 0, n_bfs = 0, s_ncand = 0, s_nact = 0, tt = clock64();
#define DRFE_TICK(acc) { const long long now = clock64(); acc += now - tt; tt = now; }
    int remaining = 0;
    for (int w = lane; w < nw; w += 32) remaining += __popc(U[w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) remaining += __shfl_xor_sync(0xFFFFFFFFu, remaining, o);
    int njobs = 0;
    const bool one_word = nw <= 32;          // the whole grid fits one word per lane: grow in registers
    const int qsh = ncx >> 5, rsh = ncx & 31;
    const uint32_t r_fl = (one_word && lane < nw) ? FL[lane] : 0u, r_fr = (one_word && lane < nw) ? FR[lane] : 0u,
                   r_fu = (one_word && lane < nw) ? FU[lane] : 0u, r_fd = (one_word && lane < nw) ? FD[lane] : 0u;
    for (int guard = 0; remaining > 0 && guard <= nc; ++guard) {
      __syncwarp();   // the loop body is warp-collective throughout: start every iteration converged
      // most frequent bin, first maximum wins (Histogram.cpp:49-55)
      unsigned best = 0;
      for (int b = lane; b < kHistBins * kHistBins; b += 32) {
        const int hcnt = hist[b];
        if (hcnt > 0) best = max(best, ((unsigned)hcnt << 16) | (unsigned)(0xFFFF - b));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
      if (best == 0) break;
      const int best_bin = 0xFFFF - (int)(best & 0xFFFFu);
      const int ncand = (int)(best >> 16);
      if (ncand < 5) break;                                  // Checkpoint 1 (:120)
      // candidate cells in ascending order
      {
        // four words of cells per round: the bin loads are issued together, ahead of the ballots and list stores
        int base = 0;
        const unsigned lt = (1u << lane) - 1u;
        const int c_end = s_blast[best_bin] + 1;
        for (int c0 = s_bfirst[best_bin] & ~31; c0 < c_end && base < ncand; c0 += 128) {
          short bv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) { const int c = c0 + 32 * k + lane; bv[k] = c < nc ? bin[c] : (short)-2; }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool is = bv[k] == best_bin;
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, is);
            if (is) list[base + __popc(bal & lt)] = c0 + 32 * k + lane;
            base += __popc(bal);
          }
        }
      }
      __syncwarp();
      DRFE_TICK(c_arg) ++n_seeds; s_ncand += ncand;
      // seed = candidate with the smallest MSE, with the reference's stray index (:125-132):
      // the running minimum is refreshed from Grid[i] (loop counter), not Grid[candidate].
      // Both MSE values of a step are independent of the running state, so they are fetched
      // ahead (4 steps at a time); only the compare/select is sequential.
      int seed = 0;
      if (lane == 0) {
        seed = list[0];
        float min_mse = (float)2147483647;
        int i = 0;
        for (; i + 4 <= ncand; i += 4) {
          const int c0 = list[i], c1 = list[i + 1], c2 = list[i + 2], c3 = list[i + 3];
          const float a0 = mse[c0], a1 = mse[c1], a2 = mse[c2], a3 = mse[c3];
          const float b0 = mse[i], b1 = mse[i + 1], b2 = mse[i + 2], b3 = mse[i + 3];
          if (a0 < min_mse) { seed = c0; min_mse = b0; }
          if (a1 < min_mse) { seed = c1; min_mse = b1; }
          if (a2 < min_mse) { seed = c2; min_mse = b2; }
          if (a3 < min_mse) { seed = c3; min_mse = b3; }
        }
        for (; i < ncand; ++i) {
          const int c = list[i];
          if (mse[c] < min_mse) { seed = c; min_mse = mse[i]; }
        }
      }
      seed = __shfl_sync(0xFFFFFFFFu, seed, 0);
      DRFE_TICK(c_scan)
      // RegionGrowing from the seed with its own plane (:142)
      const drfe_plane& sd = cells[seed];
      bool seed_ok;
      {
        const double dist = sd.normal[0] * sd.mean[0] + sd.normal[1] * sd.mean[1] + sd.normal[2] * sd.mean[2] + sd.d;
        seed_ok = ((U[seed >> 5] >> (seed & 31)) & 1u) &&
                  !(sd.normal[0] * sd.normal[0] + sd.normal[1] * sd.normal[1] + sd.normal[2] * sd.normal[2] < min_cos ||
                    dist * dist > (double)tols[seed]);
      }
      if (!seed_ok) { if (lane == 0) atomicOr(P.status, 2); break; }   // the reference would never terminate here
      uint32_t* cur = A;
      if (one_word) {
        // one word per lane: neighbours' words by shuffle, everything else in registers
        const uint32_t u = lane < nw ? U[lane] : 0u;
        uint32_t a = (lane == (seed >> 5)) ? (1u << (seed & 31)) : 0u;
        for (;;) {
          const uint32_t up1 = __shfl_up_sync(0xFFFFFFFFu, a, 1), dn1 = __shfl_down_sync(0xFFFFFFFFu, a, 1);
          const uint32_t p1 = lane >= 1 ? up1 : 0u, n1 = lane + 1 < 32 ? dn1 : 0u;
          // A << ncx and A >> ncx, word granularity qsh, bit granularity rsh
          uint32_t sl = __shfl_up_sync(0xFFFFFFFFu, a, qsh), sl2 = __shfl_up_sync(0xFFFFFFFFu, a, qsh + 1);
          uint32_t sr = __shfl_down_sync(0xFFFFFFFFu, a, qsh), sr2 = __shfl_down_sync(0xFFFFFFFFu, a, qsh + 1);
          if (lane < qsh) sl = 0u;
          if (lane < qsh + 1) sl2 = 0u;
          if (lane + qsh >= 32) sr = 0u;
          if (lane + qsh + 1 >= 32) sr2 = 0u;
          if (qsh == 0) { sl = a; sr = a; }
          const uint32_t shl_n = rsh ? ((sl << rsh) | (sl2 >> (32 - rsh))) : sl;
          const uint32_t shr_n = rsh ? ((sr >> rsh) | (sr2 << (32 - rsh))) : sr;
          uint32_t a2 = a | (((((a << 1) | (p1 >> 31)) & r_fl) | (((a >> 1) | (n1 << 31)) & r_fr) | (shl_n & r_fu) | (shr_n & r_fd)) & u);
#pragma unroll
          for (int it = 0; it < 4; ++it) a2 |= (((a2 << 1) & r_fl) | ((a2 >> 1) & r_fr)) & u;   // in-word row closure
          const bool changed = a2 != a;
          a = a2;
          ++n_bfs;
          if (!__any_sync(0xFFFFFFFFu, changed)) break;
        }
        if (lane < nw) cur[lane] = a;
        __syncwarp();
      } else {
        uint32_t* nxt = Bv;
        for (int w = lane; w < nw; w += 32) cur[w] = (w == (seed >> 5)) ? (1u << (seed & 31)) : 0u;
        __syncwarp();
        for (;;) {
          bool changed = false;
          for (int w = lane; w < nw; w += 32) {
            const uint32_t a = cur[w], u = U[w], l = FL[w], r = FR[w];
            uint32_t a2 = a | (((bv_shl(cur, nw, w, 1) & l) | (bv_shr(cur, nw, w, 1) & r) | (bv_shl(cur, nw, w, ncx) & FU[w]) |
                                (bv_shr(cur, nw, w, ncx) & FD[w])) & u);
#pragma unroll
            for (int it = 0; it < 4; ++it) a2 |= (((a2 << 1) & l) | ((a2 >> 1) & r)) & u;   // in-word row closure
            nxt[w] = a2;
            changed |= (a2 != a);
          }
          __syncwarp();
          { uint32_t* t = cur; cur = nxt; nxt = t; }
          ++n_bfs;
          if (!__any_sync(0xFFFFFFFFu, changed)) break;
        }
      }
      DRFE_TICK(c_bfs)
      int nact = 0;
      for (int w = lane; w < nw; w += 32) nact += __popc(cur[w]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nact += __shfl_xor_sync(0xFFFFFFFFu, nact, o);
      // remove the activated cells from the histogram and the unassigned mask (:150-153); regions
      // of at least 4 cells (Checkpoint 2, :157) become a job
      const bool is_job = nact >= 4;
      if (is_job && njobs >= P.max_jobs) { if (lane == 0) atomicOr(P.status, 4); break; }
      for (int w = lane; w < nw; w += 32) {
        uint32_t bits = cur[w];
        U[w] &= ~bits;
        while (bits) {
          const int c = (w << 5) + __ffs(bits) - 1;
          bits &= bits - 1;
          atomicSub(&hist[bin[c]], 1);
          bin[c] = -1;
          if (is_job) jobid[c] = (unsigned short)njobs;
        }
      }
      if (is_job) { if (lane == 0) { job_seed[njobs] = seed; job_nact[njobs] = (unsigned short)nact; } ++njobs; }
      remaining -= nact;
      __syncwarp();
      DRFE_TICK(c_acc) s_nact += nact;
    }
    if (lane == 0) s_njobs = njobs;
    if (lane == 0 && P.dbg) {
      long long* d = P.dbg + (long long)f * 16;
      d[0] = n_seeds; d[1] = n_bfs; d[2] = s_ncand; d[3] = s_nact; d[4] = c_arg; d[5] = c_scan; d[6] = c_bfs; d[7] = c_acc; d[8] = 0;
      d[9] = t_setup; d[11] = clock64();
    }
#undef DRFE_TICK
  }
  __syncthreads();
  const int njobs = s_njobs;
  double* jobacc = P.jobacc + (long long)f * P.max_jobs * 10;
  drfe_plane* jobseg = P.jobseg + (long long)f * P.max_jobs;
  const bool sums_smem = P.grid_sums_smem != 0;
  // The cells of every job, ascending, one job after the other (counting sort by job: job_nact[] are the counts), so
  // that a job's sums visit its own cells only.  list[] and mse[] are free between the seed loop and the cylinder stage.
  int* jcell = list;                                          // [sum of job_nact]
  int* job_off = reinterpret_cast<int*>(mse);                 // [njobs + 1]
  int* job_cur = job_off + njobs + 1;                         // [njobs] fill positions
  if (wid == 0) {
    const unsigned lt = (1u << lane) - 1u;
    int run = 0;
    for (int j0 = 0; j0 < njobs; j0 += 32) {
      const int j = j0 + lane;
      const int v = j < njobs ? (int)job_nact[j] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
      if (j < njobs) { job_off[j] = run + inc - v; job_cur[j] = run + inc - v; }
      run += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (lane == 0) job_off[njobs] = run;
    __syncwarp();
    for (int c0 = 0; c0 < nc; c0 += 32) {
      const int c = c0 + lane;
      const int j = c < nc ? (int)jobid[c] : 0xFFFF;
      const bool has = j != 0xFFFF;
      const unsigned act = __ballot_sync(0xFFFFFFFFu, has);
      if (has) {
        const unsigned m = __match_any_sync(act, j);          // lanes of this word in the same job
        const int leader = __ffs(m) - 1;
        int pos = 0;
        if (lane == leader) { pos = job_cur[j]; job_cur[j] = pos + __popc(m); }
        pos = __shfl_sync(m, pos, leader);
        jcell[pos + __popc(m & lt)] = c;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // ---- jobs: new_ps = *Grid[seed], then expandSegment(Grid[i]) for every activated i in ascending
  // order (the seed is counted twice, :134,146-149).  One thread per (job, sum).
  for (int t = tid; t < njobs * 10; t += THREADS) {
    const int j = t / 10, k = t - j * 10;
    const drfe_plane& sd = cells[job_seed[j]];
    const int i0 = job_off[j], i1 = job_off[j + 1];
    if (k < 9) {
      double acc = (&sd.x_acc)[k];
      for (int i = i0; i < i1; ++i) { const int c = jcell[i]; acc += sums_smem ? (double)sums[c * 9 + k] : (&cells[c].x_acc)[k]; }
      jobacc[j * 10 + k] = acc;
    } else {
      int acc = sd.nr_pts;
      for (int i = i0; i < i1; ++i) acc += npts[jcell[i]];
      jobacc[j * 10 + 9] = (double)acc;
    }
  }
  __syncthreads();
  for (int j = tid; j < njobs; j += THREADS) {
    drfe_plane ps = cells[job_seed[j]];
    const double* a = jobacc + j * 10;
    ps.x_acc = a[0]; ps.y_acc = a[1]; ps.z_acc = a[2]; ps.xx_acc = a[3]; ps.yy_acc = a[4];
    ps.zz_acc = a[5]; ps.xy_acc = a[6]; ps.xz_acc = a[7]; ps.yz_acc = a[8];
    ps.nr_pts = (int)a[9];
    fit_plane(ps);
    jobseg[j] = ps;
    job_label[j] = ps.score > 100 ? 1 : 0;                   // it is a plane (:163)
  }
  __syncthreads();
  if (tid == 0 && P.dbg) P.dbg[(long long)f * 16 + 12] = clock64();   // after accumulate + fit of the jobs
  // ---- extruded regions (cylinder_detection, CAPE.cpp:179-216): a region of more than 5 cells that is not
  // a plane goes through CylinderSeg; regions are taken in job order because they share one rand() stream.
  CylSub* subs = CYL ? P.subs + (long long)f * P.max_sub : nullptr;
  unsigned short* subid = reinterpret_cast<unsigned short*>(bin);   // bin[] is free after the seed loop
  if (CYL) {
    for (int c = tid; c < nc; c += THREADS) subid[c] = 0xFFFF;
    __syncthreads();
    if (wid == 0) {
      CylCtx X;
      X.cells = cells; X.sums = sums_smem ? sums : nullptr; X.npts = npts; X.jobid = jobid; X.nc = nc;
      X.l2g = list; X.ids_left = reinterpret_cast<int*>(mse); X.inl = cinl; X.flag = cflag; X.D = cylD;
      X.sN = P.cyl_scratch + (long long)f * 6 * nc; X.sP = X.sN + 3 * nc;
      X.rand_tab = P.rand_tab; X.rand_n = P.rand_n; X.subs = subs; X.max_sub = P.max_sub; X.subid = subid;
      X.K1 = P.cylK1; X.K2 = P.cylK2; X.status = P.status;
      int rpos = 0, nsub = 0;
      for (int j = 0; j < njobs; ++j) {
        if (lane == 0) job_sub0[j] = (unsigned short)nsub;
        if (!job_label[j] && job_nact[j] > 5) cyl_job(X, j, job_nact[j], cells[job_seed[j]], rpos, nsub);
      }
      if (lane == 0) job_sub0[njobs] = (unsigned short)nsub;
    }
    __syncthreads();
  }
  if (tid == 0 && P.dbg) P.dbg[(long long)f * 16 + 13] = clock64();   // after the cylinder jobs
  // ---- labels in the reference's push order: job by job; an extruded job contributes its sub-segments
  if (tid == 0) {
    int np = 0, ncyl = 0;
    auto next_plane = [&]() -> int {
      if (np < kMaxPlanes) return ++np;
      atomicOr(P.status, 1);
      return 0;
    };
    for (int j = 0; j < njobs; ++j) {
      if (job_label[j]) { job_label[j] = (uint8_t)next_plane(); continue; }
      for (int s = CYL ? job_sub0[j] : 0, s_end = CYL ? job_sub0[j + 1] : 0; s < s_end; ++s) {
        if (!subs[s].is_cyl) {
          const int l = next_plane();
          subs[s].label = l;
          if (l) segs[l - 1] = subs[s].ps;
        } else {
          subs[s].label = ++ncyl;                              // cylinder2region_map (CAPE.cpp:203-204)
          drfe_cylinder cyo;
          cyo.radius = subs[s].radius;
          for (int k = 0; k < 3; ++k) { cyo.center[k] = subs[s].center[k]; cyo.axis[k] = subs[s].axis[k]; }
          P.cyls[(long long)f * P.max_sub + ncyl - 1] = cyo;   // cylinder_segments_final (:434-445)
        }
      }
    }
    s_np = np;
    s_ncyl = ncyl;
  }
  __syncthreads();
  for (int j = tid; j < njobs; j += THREADS)
    if (job_label[j]) segs[job_label[j] - 1] = jobseg[j];
  for (int c = tid; c < nc; c += THREADS) {
    const int j = jobid[c];
    int pl = (j == 0xFFFF) ? 0 : job_label[j], cl = 0;
    if (CYL && subid[c] != 0xFFFF) {
      const CylSub& sb = subs[subid[c]];
      if (sb.is_cyl) cl = sb.label; else pl = sb.label;
    }
    pmap[c] = (uint8_t)pl;
    if (CYL) P.cyl_map[(long long)f * nc + c] = cl;
  }
  __syncthreads();
  if (tid == 0 && P.dbg) P.dbg[(long long)f * 16 + 14] = clock64();   // after labelling
  // ---- plane merging (CAPE.cpp:220-252; getConnectedComponents :459-481)
  const int np = s_np;
  for (int c = tid; c < nc; c += THREADS) {
    const int r = c / ncx, x = c - r * ncx;
    if (r >= ncy - 1 || x >= ncx - 1) continue;
    const int v = pmap[c];
    if (v <= 0) continue;
    const int right = pmap[c + 1], below = pmap[c + ncx];
    if (right > 0 && v != right) atomicOr(&assoc[(v - 1) * 8 + ((right - 1) >> 5)], 1u << ((right - 1) & 31));
    if (below > 0 && v != below) atomicOr(&assoc[(v - 1) * 8 + ((below - 1) >> 5)], 1u << ((below - 1) & 31));
  }
  __syncthreads();
  if (tid == 0) {
    auto get = [&](int r, int c) { return (assoc[r * 8 + (c >> 5)] >> (c & 31)) & 1u; };
    for (int i = 0; i < np; ++i) merge[i] = i;
    for (int r = 0; r < np; ++r) {
      const int pid = merge[r];
      bool expanded = false;
      for (int c = r + 1; c < np; ++c) {
        if (!(get(r, c) || get(c, r))) continue;
        const drfe_plane& Ap = segs[pid];
        const drfe_plane& Cc = segs[c];
        const double cosang = Ap.normal[0] * Cc.normal[0] + Ap.normal[1] * Cc.normal[1] + Ap.normal[2] * Cc.normal[2];
        // sic: the x term uses plane r, the others plane_id (:238-240)
        const double dd = segs[r].normal[0] * Cc.mean[0] + Ap.normal[1] * Cc.mean[1] + Ap.normal[2] * Cc.mean[2] + Ap.d;
        if (cosang > (double)P.min_cos && dd * dd < (double)P.max_merge_dist) {
          expand_seg(segs[pid], segs[c]);
          merge[c] = pid;
          expanded = true;
        }
      }
      if (expanded) fit_plane(segs[pid]);
    }
  }
  __syncthreads();
  // ---- per final plane: cell mask, erode (cross), dilate (square) (CAPE.cpp:254-291)
  uint8_t* eroded_map = P.eroded_map + (long long)f * nc;
  uint32_t* border = P.border_vec + (long long)f * P.border_rows * nw;
  for (int c = tid; c < nc; c += THREADS) eroded_map[c] = 0;
  int nfinal = 0;
  for (int i = 0; i < np; ++i) {
    if (merge[i] != i) continue;
    for (int w = tid; w < nw; w += THREADS) {
      uint32_t m = 0;
      const int cend = min(32, nc - (w << 5));
      for (int b = 0; b < cend; ++b) {
        const int v = pmap[(w << 5) + b];
        if (v > i && merge[v - 1] == i) m |= 1u << b;          // j >= i with merge label i
      }
      M[w] = m;
    }
    __syncthreads();
    int any = 0;
    uint32_t er_w[4];                                          // this thread's words of the eroded mask
    for (int w = tid, k = 0; w < nw; w += THREADS, ++k) {
      const uint32_t m = M[w];
      const uint32_t l1 = bv_shl(M, nw, w, 1), r1 = bv_shr(M, nw, w, 1);
      const uint32_t e = m & (l1 | C0[w]) & (r1 | CL[w]) & (bv_shl(M, nw, w, ncx) | R0[w]) & (bv_shr(M, nw, w, ncx) | RL[w]);
      Hh[w] = m | (l1 & ~C0[w]) | (r1 & ~CL[w]);
      if (k < 4) er_w[k] = e;
      any |= (e != 0);
    }
    const int keep = __syncthreads_or(any);                    // completely eroded planes are ignored (:275)
    if (keep) {
      const int plane_nr = ++nfinal;
      if (tid == 0) {
        const drfe_plane& ps = segs[i];
        P.planes[(long long)f * kMaxPlanes + plane_nr - 1] = ps;
        P.plane_eq[(long long)f * (kMaxPlanes + 1) + plane_nr] =
            make_float4((float)ps.normal[0], (float)ps.normal[1], (float)ps.normal[2], (float)ps.d);
        P.plane_maxd[(long long)f * (kMaxPlanes + 1) + plane_nr] = 9 * ps.MSE;
      }
      for (int w = tid, k = 0; w < nw; w += THREADS, ++k) {
        const uint32_t e = er_w[k & 3];
        const uint32_t d = (Hh[w] | bv_shl(Hh, nw, w, ncx) | bv_shr(Hh, nw, w, ncx)) & VALID[w];
        border[(long long)plane_nr * nw + w] = d & ~e;         // mask_diff = dilated - eroded (:284)
        uint32_t bits = e;
        while (bits) { const int c = (w << 5) + __ffs(bits) - 1; bits &= bits - 1; eroded_map[c] = (uint8_t)plane_nr; }
      }
    }
    __syncthreads();
  }
  // ---- cylinders: same erode / dilate per cylinder region (CAPE.cpp:323-357); border rows follow the planes'
  if (CYL) {
    uint8_t* cyl_eroded = P.cyl_eroded_map + (long long)f * nc;
    const int* cyl_map = P.cyl_map + (long long)f * nc;
    for (int c = tid; c < nc; c += THREADS) cyl_eroded[c] = 0;
    const int ncyl = s_ncyl;
    int ncf = 0;
    for (int i = 0; i < ncyl; ++i) {
      for (int w = tid; w < nw; w += THREADS) {
        uint32_t m = 0;
        const int cend = min(32, nc - (w << 5));
        for (int b = 0; b < cend; ++b)
          if (cyl_map[(w << 5) + b] == i + 1) m |= 1u << b;
        M[w] = m;
      }
      __syncthreads();
      int any = 0;
      uint32_t er_w[4];
      for (int w = tid, k = 0; w < nw; w += THREADS, ++k) {
        const uint32_t m = M[w];
        const uint32_t l1 = bv_shl(M, nw, w, 1), r1 = bv_shr(M, nw, w, 1);
        const uint32_t e = m & (l1 | C0[w]) & (r1 | CL[w]) & (bv_shl(M, nw, w, ncx) | R0[w]) & (bv_shr(M, nw, w, ncx) | RL[w]);
        Hh[w] = m | (l1 & ~C0[w]) | (r1 & ~CL[w]);
        if (k < 4) er_w[k] = e;
        any |= (e != 0);
      }
      const int keep = __syncthreads_or(any);
      if (keep && ncf + 1 + 50 > 255) { if (tid == 0) atomicOr(P.status, 32); __syncthreads(); continue; }
      if (keep) {
        const int cyl_nr = ++ncf;
        if (tid == 0) {
          // the sub-segment with cylinder number i + 1
          int s = 0;
          for (int q = 0; q < job_sub0[njobs]; ++q)
            if (subs[q].is_cyl && subs[q].label == i + 1) s = q;
          const CylSub& sb = subs[s];
          CylEq eqo;
          for (int k = 0; k < 3; ++k) { eqo.p2[k] = sb.p2[k]; eqo.dir[k] = sb.p2[k] - sb.p1[k]; }
          eqo.n12 = (double)sb.n12; eqo.radius = (double)sb.radius;
          eqo.maxd = (float)(9.0 * sb.mse); eqo.pad = 0;
          P.cyl_eq[(long long)f * (kMaxPlanes + 1) + cyl_nr] = eqo;
        }
        for (int w = tid, k = 0; w < nw; w += THREADS, ++k) {
          const uint32_t e = er_w[k & 3];
          const uint32_t d = (Hh[w] | bv_shl(Hh, nw, w, ncx) | bv_shr(Hh, nw, w, ncx)) & VALID[w];
          border[(long long)(kMaxPlanes + 1 + cyl_nr) * nw + w] = d & ~e;
          uint32_t bits = e;
          while (bits) { const int c = (w << 5) + __ffs(bits) - 1; bits &= bits - 1; cyl_eroded[c] = (uint8_t)(50 + cyl_nr); }
        }
      }
      __syncthreads();
    }
    if (tid == 0) { P.ncyl_final[f] = ncf; P.ncyl_found[f] = ncyl; }
  }
  int* plane_map = P.plane_map + (long long)f * nc;
  for (int c = tid; c < nc; c += THREADS) plane_map[c] = pmap[c];
  if (tid == 0) P.nplanes[f] = nfinal;
  if (tid == 0 && P.dbg) { long long* d = P.dbg + (long long)f * 16; d[10] = clock64(); }
}

// ------------------------------------------------------------------ refinement + output
// Label per pixel = argmin over final planes (in order) of the squared float distance, subject to < 9*MSE, strict '<'
// against the running minimum which starts at the bit pattern memset(...,100,...) leaves (0x64646464, CAPE.cpp:60);
// cells inside an eroded mask are painted whole (:410-412).  Three launches:
//   k_cape_refine_plan    thread per cell: the label a cell is painted with as a whole (eroded plane mask, then eroded
//                         cylinder mask, else 0) and — for the cells inside some plane's / cylinder's dilated-minus-eroded
//                         mask — an entry in the launch's list of border cells;
//   k_cape_paint          image space, one word (4 pixels) per thread: every pixel gets its cell's label, the margin
//                         outside the last full cell row / column gets 0 (fully coalesced stores of seg_output);
//   k_cape_refine_border  one warp per LISTED cell, warps looping over the list: the per-pixel argmin.  Only the cells
//                         that need it (29 % on the synthetic sequence) occupy warps: when the same kernel also painted,
//                         a block kept its slots until its one or two border cells were done and the SMs ran at 15 %
//                         occupancy.  With a depth image the points of a border cell are converted again here (the same
//                         arithmetic, hence the same bits) instead of being read back from a cloud that k_cape_sums no
//                         longer writes.
template <bool CYL>
__global__ void __launch_bounds__(256) k_cape_refine_plan(const CapeDev* __restrict__ Pp, int f0, int nframes) {
  DRFE_GRID_DEP();
  const CapeDev& P = *Pp;
  const int gid0 = blockIdx.x * 256 + threadIdx.x;
  if (gid0 >= nframes * P.ncells) return;
  const int gid = gid0 + f0 * P.ncells;
  const int f = gid / P.ncells, cell = gid - f * P.ncells;
  const int er = P.eroded_map[gid];
  const int cer = CYL ? P.cyl_eroded_map[gid] : 0;
  const int label = er > 0 ? er : (cer > 0 ? cer : 0);
  P.cell_label[gid] = (uint8_t)label;
  if (label) return;
  const int nw = (P.ncells + 31) >> 5;
  const uint32_t* bvec = P.border_vec + (long long)f * P.border_rows * nw + (cell >> 5);
  const uint32_t bit = 1u << (cell & 31);
  bool border = false;
  uint32_t first32 = 0;                                          // bit b: final plane b + 1 has this cell in its border mask
  const int npl = P.nplanes[f];
  for (int p = 1; p <= npl; ++p)
    if (bvec[(long long)p * nw] & bit) { border = true; if (p <= 32) first32 |= 1u << (p - 1); }
  if (CYL) {
    const int ncf = P.ncyl_final[f];
    const uint32_t* cvec = bvec + (long long)(kMaxPlanes + 1) * nw;
    for (int p = 1; p <= ncf && !border; ++p) border = (cvec[(long long)p * nw] & bit) != 0;
  }
  if (border) P.border_list[atomicAdd(P.border_count, 1)] = make_int2(gid, (int)first32);
}

// grid (words of a row / 64, rows / (4 * kPaintRows), frames), block (64, 4): a thread owns one word column (4 pixels) of
// kPaintRows consecutive rows and looks its cells' labels up again only when the cell row changes; magic_* = ceil(2^32 / d)
// for d = cell width, cell height
static const int kPaintRows = 5;
__global__ void __launch_bounds__(256) k_cape_paint(const CapeDev* __restrict__ Pp, int f0, uint32_t magic_cw, uint32_t magic_ch) {
  DRFE_GRID_DEP();
  const CapeDev& P = *Pp;
  const int W = P.W, H = P.H;
  const int c0 = 4 * (blockIdx.x * 64 + threadIdx.x), r0 = (blockIdx.y * 4 + threadIdx.y) * kPaintRows;
  if (c0 >= W || r0 >= H) return;
  const int f = blockIdx.z + f0;
  const uint8_t* lab = P.cell_label + (long long)f * P.ncells;
  const int ca = (int)__umulhi((uint32_t)c0, magic_cw), cb = (int)__umulhi((uint32_t)(c0 + 3), magic_cw);
  uint8_t* out = P.seg + ((long long)f * H + r0) * W + c0;
  const bool whole = (W & 3) == 0;
  int last_cell_r = -1;
  uint32_t word = 0;
#pragma unroll
  for (int k = 0; k < kPaintRows; ++k) {
    const int r = r0 + k;
    if (r >= H) break;
    const int cell_r = (int)__umulhi((uint32_t)r, magic_ch);
    if (cell_r != last_cell_r) {
      last_cell_r = cell_r;
      word = 0;
      if (cell_r < P.ncy) {
        const uint8_t* row = lab + cell_r * P.ncx;
        if (ca == cb) word = ca < P.ncx ? (uint32_t)row[ca] * 0x01010101u : 0u;      // the usual case: the 4 pixels share a cell
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cell_c = (int)__umulhi((uint32_t)(c0 + j), magic_cw);
            if (cell_c < P.ncx) word |= (uint32_t)row[cell_c] << (8 * j);
          }
        }
      }
    }
    if (whole) *reinterpret_cast<uint32_t*>(out + (long long)k * W) = word;
    else
      for (int j = 0; j < 4 && c0 + j < W; ++j) out[(long long)k * W + j] = (uint8_t)(word >> (8 * j));
  }
}

static const int kBorderWarps = 4;     // warps per block of k_cape_refine_border
template <bool CYL, int MODE>
__global__ void __launch_bounds__(kBorderWarps * 32, CYL ? 4 : 5) k_cape_refine_border(const CapeDev* __restrict__ Pp) {
  DRFE_GRID_DEP();
  const CapeDev& P = *Pp;
  const int lane = threadIdx.x & 31;
  const int nlist = *P.border_count;
  const int npc = P.npc, cw = P.cw;
  const long long N = (long long)P.H * P.W;
  // a cell row is cw bytes; with cw and W multiples of 4 everything below moves 4 pixels per lane and access
  const bool vec4 = ((cw | P.W) & 3) == 0;
  const int q4 = cw >> 2;
  const uint32_t q4_magic = (65536u + (uint32_t)q4 - 1u) / (uint32_t)max(q4, 1);   // j / q4 == (j * magic) >> 16 for j < 2^16 / q4
  const XyCtx C = xy_ctx(P);
  for (int li = blockIdx.x * kBorderWarps + (threadIdx.x >> 5); li < nlist; li += gridDim.x * kBorderWarps) {
    const int2 ent = P.border_list[li];
    const int gw = ent.x;
    const int f = gw / P.ncells, cell = gw - f * P.ncells;
    const int cr = cell / P.ncx, cc = cell - cr * P.ncx;
    uint8_t* out = P.seg + (long long)f * N + (long long)(cr * P.ch) * P.W + cc * cw;
    // planes whose dilated-minus-eroded mask contains this cell: bit b of bits[m] = final plane 32*m + b + 1.  The first 32
    // come with the list entry; frames with more planes gather the rest from the border vectors
    const int nw = (P.ncells + 31) >> 5, npl = P.nplanes[f];
    const uint32_t* bvec = P.border_vec + (long long)f * P.border_rows * nw + (cell >> 5);
    uint32_t bits[8], cbits[8];
    uint32_t anyc = 0;
    bits[0] = (uint32_t)ent.y;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      cbits[m] = 0;
      if (m > 0) {
        bits[m] = 0;
        if (32 * m < npl) {                                        // warp-uniform
          const int p = 32 * m + lane + 1;
          const bool in = p <= npl && ((bvec[(long long)p * nw] >> (cell & 31)) & 1u);
          bits[m] = __ballot_sync(0xFFFFFFFFu, in);
        }
      }
    }
    if (CYL) {
      const int ncf = P.ncyl_final[f];
      const uint32_t* cvec = bvec + (long long)(kMaxPlanes + 1) * nw;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        if (32 * m < ncf) {
          const int p = 32 * m + lane + 1;
          const bool in = p <= ncf && ((cvec[(long long)p * nw] >> (cell & 31)) & 1u);
          cbits[m] = __ballot_sync(0xFFFFFFFFu, in);
          anyc |= cbits[m];
        }
      }
    }
    const float* CX = MODE == 0 ? P.cloud + (long long)f * 3 * N + (long long)cell * npc : nullptr;
    const float* CY = CX + N;
    const float* CZ = CY + N;
    const double col0 = (double)(cc * cw) - C.cx, row0 = (double)(cr * P.ch) - C.cy;
    const long long dbase = (long long)f * P.depth_fs + (long long)(cr * P.ch) * P.depth_rs + cc * cw;   // the cell's first depth element
    const float4* eq = P.plane_eq + (long long)f * (kMaxPlanes + 1);
    const float* maxd = P.plane_maxd + (long long)f * (kMaxPlanes + 1);
    if (!CYL && vec4 && (npc & 3) == 0 && npc < 8192 &&
        (MODE == 0 || (((P.depth_rs | P.depth_fs) & 3) == 0 && (reinterpret_cast<uintptr_t>(MODE == 1 ? (const void*)P.depth : (const void*)P.depth16) & (MODE == 1 ? 15 : 7)) == 0))) {
      // planes only: 4 pixels per lane (float4 loads of the cell-major cloud, or one aligned 4-pixel depth load), one word store
      const float4* X4 = reinterpret_cast<const float4*>(CX);
      const float4* Y4 = reinterpret_cast<const float4*>(CY);
      const float4* Z4 = reinterpret_cast<const float4*>(CZ);
      // the depth of every quad this lane will handle is requested before any of it is used (a border cell is a few
      // dependent iterations per lane: one load latency per iteration was a quarter of the kernel's stall samples)
      constexpr int kAhead = 4;                                   // 4 x 32 quads = cells of up to 512 pixels in one go
      const int nq = npc >> 2;
      float4 zq[kAhead];
      auto load_depth = [&](int j) -> float4 {
        const int lr = (int)(((uint32_t)j * q4_magic) >> 16), c4 = j - lr * q4;
        if (MODE == 1) return __ldg(reinterpret_cast<const float4*>(P.depth + dbase + (long long)lr * P.depth_rs + 4 * c4));
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(P.depth16 + dbase + (long long)lr * P.depth_rs + 4 * c4));
        const float fac = P.depth_factor;
        return make_float4((float)(v.x & 0xFFFFu) * fac, (float)(v.x >> 16) * fac, (float)(v.y & 0xFFFFu) * fac, (float)(v.y >> 16) * fac);
      };
      if (MODE != 0) {
  #pragma unroll
        for (int a = 0; a < kAhead; ++a) { const int j = lane + 32 * a; zq[a] = j < nq ? load_depth(j) : make_float4(0.f, 0.f, 0.f, 0.f); }
      }
      // the cell's first kLocal planes (ascending plane number; a border cell rarely sees more than two) come into
      // registers with independent loads before the pixel loop; whatever is left stays in `bits` for the generic loop
      constexpr int kLocal = 4;
      float4 peq[kLocal];
      float pmd[kLocal];
      int pid[kLocal], nloc = 0;
      uint32_t any_rest = 0;
  #pragma unroll
      for (int k = 0; k < 8; ++k) {
        while (bits[k] && nloc < kLocal) {
          const int p = k * 32 + __ffs(bits[k]);
          bits[k] &= bits[k] - 1;
  #pragma unroll
          for (int i = 0; i < kLocal; ++i)
            if (i == nloc) { pid[i] = p; peq[i] = eq[p]; pmd[i] = maxd[p]; }
          ++nloc;
        }
        any_rest |= bits[k];
      }
  #pragma unroll 1
      for (int j0 = lane; j0 < nq; j0 += 32 * kAhead) {
  #pragma unroll
       for (int a = 0; a < kAhead; ++a) {
        const int j = j0 + 32 * a;
        if (j >= nq) break;
        const int lr = (int)(((uint32_t)j * q4_magic) >> 16), c4 = j - lr * q4;
        float xs[4], ys[4], zs[4];
        if (MODE == 0) {
          const float4 xv = X4[j], yv = Y4[j], zv = Z4[j];
          xs[0] = xv.x; xs[1] = xv.y; xs[2] = xv.z; xs[3] = xv.w; ys[0] = yv.x; ys[1] = yv.y; ys[2] = yv.z; ys[3] = yv.w;
          zs[0] = zv.x; zs[1] = zv.y; zs[2] = zv.z; zs[3] = zv.w;
        } else {
          const float4 zv = j0 == lane ? zq[a] : load_depth(j);   // cells of more than 512 pixels: later rounds load on the spot
          zs[0] = zv.x; zs[1] = zv.y; zs[2] = zv.z; zs[3] = zv.w;
          const double drow = row0 + (double)lr, dcol = col0 + (double)(4 * c4);
          const double ky = drow * C.rfy;
          double qx[4], qy[4];
          bool exact = !C.sane;
  #pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double zd = (double)zs[u];
            qx[u] = zd * ((dcol + (double)u) * C.rfx); qy[u] = zd * ky;
            exact |= xy_needs_exact(zs[u], qx[u], qy[u]);
          }
          if (exact) {                                              // rare: all four the slow way
  #pragma unroll
            for (int u = 0; u < 4; ++u) xy_exact(zs[u], dcol + (double)u, drow, C.fx, C.fy, xs[u], ys[u]);
          } else {
  #pragma unroll
            for (int u = 0; u < 4; ++u) { xs[u] = (float)qx[u]; ys[u] = (float)qy[u]; }
          }
        }
        float best[4];
        int labs[4];
  #pragma unroll
        for (int u = 0; u < 4; ++u) { best[u] = __uint_as_float(0x64646464u); labs[u] = 0; }
  #pragma unroll
        for (int i = 0; i < kLocal; ++i) {
          if (i < nloc) {
            const float4 e = peq[i];
  #pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float v = xs[u] * e.x + ys[u] * e.y + zs[u] * e.z + e.w;
              const float dist = v * v;
              if (dist < fminf(pmd[i], best[u])) { best[u] = dist; labs[u] = pid[i]; }
            }
          }
        }
        if (any_rest) {
  #pragma unroll
          for (int k = 0; k < 8; ++k) {
            uint32_t b = bits[k];
            while (b) {
              const int p = k * 32 + __ffs(b);
              b &= b - 1;
              const float4 e = eq[p];
              const float md = maxd[p];
  #pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float v = xs[u] * e.x + ys[u] * e.y + zs[u] * e.z + e.w;
                const float dist = v * v;
                if (dist < fminf(md, best[u])) { best[u] = dist; labs[u] = p; }
              }
            }
          }
        }
        *reinterpret_cast<uint32_t*>(out + (long long)lr * P.W + 4 * c4) =
            (uint32_t)labs[0] | ((uint32_t)labs[1] << 8) | ((uint32_t)labs[2] << 16) | ((uint32_t)labs[3] << 24);
       }
      }
      continue;
    }
    for (int i = lane; i < npc; i += 32) {
      const int lr = i / cw, lc = i - lr * cw;
      float x, y, z;
      if (MODE == 0) { x = CX[i]; y = CY[i]; z = CZ[i]; }
      else {
        if (MODE == 1) z = __ldg(P.depth + dbase + (long long)lr * P.depth_rs + lc);
        else z = (float)__ldg(P.depth16 + dbase + (long long)lr * P.depth_rs + lc) * P.depth_factor;
        xy_from_depth(C, z, col0 + (double)lc, row0 + (double)lr, x, y);
      }
      float best = __uint_as_float(0x64646464u);
      int lab = 0;
  #pragma unroll
      for (int k = 0; k < 8; ++k) {
        uint32_t b = bits[k];
        while (b) {
          const int p = k * 32 + __ffs(b);
          b &= b - 1;
          const float4 e = eq[p];
          const float v = x * e.x + y * e.y + z * e.z + e.w;
          const float dist = v * v;
          if (dist < maxd[p] && dist < best) { best = dist; lab = p; }
        }
      }
      if (CYL && anyc && z > 0.f) {
        // point-to-axis distance minus radius (CAPE.cpp:375-385): float cross / norm, double divide
        const CylEq* ceq = P.cyl_eq + (long long)f * (kMaxPlanes + 1);
  #pragma unroll
        for (int k = 0; k < 8; ++k) {
          uint32_t b = cbits[k];
          while (b) {
            const int p = k * 32 + __ffs(b);
            b &= b - 1;
            const CylEq& e = ceq[p];
            const float q0 = x - e.p2[0], q1 = y - e.p2[1], q2 = z - e.p2[2];
            const float c0 = e.dir[1] * q2 - e.dir[2] * q1, c1 = e.dir[2] * q0 - e.dir[0] * q2, c2 = e.dir[0] * q1 - e.dir[1] * q0;
            const float nrm = sqrtf(c0 * c0 + (c1 * c1 + c2 * c2));
            float dist = (float)((double)nrm / e.n12 - e.radius);
            dist = dist * dist;
            if (dist < e.maxd && dist < best) { best = dist; lab = 50 + p; }
          }
        }
      }
      out[(long long)lr * P.W + lc] = (uint8_t)lab;
    }
  }
}


}  // namespace drfe

// ====================================================================== host side
using namespace drfe;

// ------------------------------------------------------------------ per-plane point lists
// PlaneDetection_CAPE::runPlaneDetection after CAPE::process (PlaneExtractor.cpp:165-190): every pixel whose
// seg_output code is > 0 appends its (x, y, z) to plane_cloud[code - 1], pixels visited in row-major order.
// One CTA per frame, each of its 32 warps owns a contiguous run of pixels: pass 1 counts the run's pixels per
// label (one shared-memory add per group of equal labels in a 32-pixel step, via match.any), an exclusive scan
// over (label, warp) turns the counts into output positions, pass 2 re-reads the labels and writes the points
// (stable: lower pixel index first inside a label).  Codes above nr_planes (cylinder labels 51+) are skipped —
// the reference would index plane_cloud out of range there; DR-SLAM runs CAPE with cylinder detection off.
static const int kPtsWarps = 32;
static __global__ void __launch_bounds__(kPtsWarps * 32) k_cape_plane_points(const CapeDev* __restrict__ Pp, int f0, float* __restrict__ out,
                                                                     int* __restrict__ offsets, uint32_t magic_w, uint32_t magic_cw, uint32_t magic_ch) {
  extern __shared__ int s_pos[];                              // [kPtsWarps][np + 1] counts, then running output positions
  __shared__ int s_start[kMaxPlanes + 2];
  const CapeDev& P = *Pp;
  const int f = blockIdx.x + f0;
  const int np = min(P.nplanes[f], kMaxPlanes);
  const int N = P.H * P.W;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  int* offs = offsets + (long long)f * (kMaxPlanes + 1);
  if (np == 0) { if (tid == 0) offs[0] = 0; return; }
  const int stride = np + 1;
  for (int i = tid; i < kPtsWarps * stride; i += kPtsWarps * 32) s_pos[i] = 0;
  __syncthreads();
  const uint8_t* __restrict__ seg = P.seg + (long long)f * N;
  const int run = ((N + kPtsWarps - 1) / kPtsWarps + 31) & ~31;  // pixels per warp, a multiple of 32
  const int p0 = w * run, p1 = min(p0 + run, N);
  int* mine = s_pos + w * stride;
  for (int p = p0 + lane; p - lane < p1; p += 32) {
    int key = p < p1 ? seg[p] : 0;
    if (key > np) key = 0;
    const unsigned m = __match_any_sync(0xFFFFFFFFu, key);
    if (key && (m & lt) == 0) mine[key] += __popc(m);         // the group's lowest lane adds its size
    __syncwarp();                                              // the next trip's leader for this key may be another lane
  }
  __syncthreads();
  if (tid < np) {                                              // total of label tid + 1
    int t = 0;
    for (int k = 0; k < kPtsWarps; ++k) t += s_pos[k * stride + tid + 1];
    s_start[tid + 1] = t;
  }
  __syncthreads();
  if (tid == 0) {
    int run_tot = 0;
    for (int L = 1; L <= np; ++L) { const int t = s_start[L]; s_start[L] = run_tot; offs[L - 1] = run_tot; run_tot += t; }
    offs[np] = run_tot;
  }
  __syncthreads();
  if (tid < np) {
    int at = s_start[tid + 1];
    for (int k = 0; k < kPtsWarps; ++k) { const int t = s_pos[k * stride + tid + 1]; s_pos[k * stride + tid + 1] = at; at += t; }
  }
  __syncthreads();
  const float* __restrict__ CX = P.cloud + (long long)f * 3 * N;
  float* __restrict__ dst = out + (long long)f * 3 * N;
  for (int p = p0 + lane; p - lane < p1; p += 32) {
    int key = p < p1 ? seg[p] : 0;
    if (key > np) key = 0;
    const unsigned m = __match_any_sync(0xFFFFFFFFu, key);
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (key && lane == leader) { base = mine[key]; mine[key] = base + __popc(m); }
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (key) {
      const int r = (int)__umulhi((uint32_t)p, magic_w), c = p - r * P.W;
      const int cell_r = (int)__umulhi((uint32_t)r, magic_ch), cell_c = (int)__umulhi((uint32_t)c, magic_cw);
      const int idx = (cell_r * P.ncx + cell_c) * P.npc + (r - cell_r * P.ch) * P.cw + (c - cell_c * P.cw);
      float* o = dst + 3ll * (base + __popc(m & lt));
      o[0] = CX[idx]; o[1] = CX[idx + N]; o[2] = CX[idx + 2 * N];
    }
    __syncwarp();
  }
}


// ------------------------------------------------------------------ 5 cm voxel filter of the per-plane point lists
// pcl::VoxelGrid<PointT>::applyFilter (PCL 1.9, pcl/filters/impl/voxel_grid.hpp) as Frame::ComputePlanes_CAPE runs it on
// every plane_cloud[i] (reference src/Frame.cc:1121-1125: setLeafSize(0.05, 0.05, 0.05), default downsample_all_data): bounding
// box, leaf index per point, points ordered by leaf, one centroid per leaf (float sum / float count, AccumulatorXYZ), leaves in
// ascending index order.  Declared: a leaf's points are summed in ascending input order (PCL's std::sort leaves that order to
// the library).  One segment = one plane of one frame; a CTA works on one segment at a time:
//   k_voxel_sort       bounds (min / max reduction), leaf index of every point, stable LSD radix sort of (leaf, point) by 8-bit
//                      digits — only as many passes as the box's leaf count needs —, then the leaf heads are counted and listed;
//   k_voxel_centroids  one thread per leaf adds its points in order and writes the centroid at the plane's offset in the frame.
struct VoxSeg { int nvox, unfiltered, sorted_in_b, pad; };
static const int kVoxThreads = 1024;

__global__ void __launch_bounds__(kVoxThreads) k_voxel_sort(const float* __restrict__ pts, const int* __restrict__ offs, const int* __restrict__ nplanes, int N,
                                                            float inv_leaf, uint32_t* __restrict__ keyA, uint32_t* __restrict__ valA,
                                                            uint32_t* __restrict__ keyB, uint32_t* __restrict__ valB, VoxSeg* __restrict__ segs) {
  __shared__ uint32_t s_cnt[32][256];                           // per warp and digit: count, then output position
  __shared__ uint32_t s_base[256];
  __shared__ float s_red[6][32];
  __shared__ int s_i[8];
  const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const int np = min(nplanes[f], kMaxPlanes);
  const int* fo = offs + (long long)f * (kMaxPlanes + 1);
  for (int p = blockIdx.x; p < np; p += gridDim.x) {
    const int s0 = fo[p], n = fo[p + 1] - s0;
    VoxSeg* seg = segs + (long long)f * kMaxPlanes + p;
    const float* P3 = pts + ((long long)f * N + s0) * 3;
    uint32_t* kA = keyA + (long long)f * N + s0; uint32_t* vA = valA + (long long)f * N + s0;
    uint32_t* kB = keyB + (long long)f * N + s0; uint32_t* vB = valB + (long long)f * N + s0;
    if (n <= 0) { if (tid == 0) { seg->nvox = 0; seg->unfiltered = 0; seg->sorted_in_b = 0; } continue; }
    // ---- getMinMax3D
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = tid; i < n; i += kVoxThreads)
#pragma unroll
      for (int k = 0; k < 3; ++k) { const float v = P3[3 * i + k]; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { mn[k] = fminf(mn[k], __shfl_xor_sync(0xFFFFFFFFu, mn[k], o)); mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xFFFFFFFFu, mx[k], o)); }
      if (lane == 0) { s_red[k][wid] = mn[k]; s_red[3 + k][wid] = mx[k]; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      mn[k] = s_red[k][0]; mx[k] = s_red[3 + k][0];
      for (int w = 1; w < 32; ++w) { mn[k] = fminf(mn[k], s_red[k][w]); mx[k] = fmaxf(mx[k], s_red[3 + k][w]); }
    }
    // "Leaf size is too small for the input dataset": more than INT32_MAX leaves in the box -> the input is returned as it is
    const long long d0 = (long long)(__fmul_rn(__fsub_rn(mx[0], mn[0]), inv_leaf)) + 1, d1 = (long long)(__fmul_rn(__fsub_rn(mx[1], mn[1]), inv_leaf)) + 1,
                    d2 = (long long)(__fmul_rn(__fsub_rn(mx[2], mn[2]), inv_leaf)) + 1;
    const bool unfiltered = (double)d0 * (double)d1 * (double)d2 > 2147483647.0;
    if (unfiltered) {
      if (tid == 0) { seg->nvox = n; seg->unfiltered = 1; seg->sorted_in_b = 0; }
      __syncthreads();
      continue;
    }
    int min_b[3], mul[3];
    {
      int div_b[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        min_b[k] = (int)floorf(__fmul_rn(mn[k], inv_leaf));
        div_b[k] = (int)floorf(__fmul_rn(mx[k], inv_leaf)) - min_b[k] + 1;
      }
      mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
      const uint32_t total = (uint32_t)(div_b[0] * div_b[1] * div_b[2]);
      if (tid == 0) s_i[0] = total > 1 ? 32 - __clz(total - 1) : 0;   // bits a leaf index needs
    }
    // ---- leaf index of every point
    for (int i = tid; i < n; i += kVoxThreads) {
      int idx = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) idx += (int)__fsub_rn(floorf(__fmul_rn(P3[3 * i + k], inv_leaf)), (float)min_b[k]) * mul[k];
      kA[i] = (uint32_t)idx; vA[i] = (uint32_t)i;
    }
    __syncthreads();
    const int passes = (s_i[0] + 7) >> 3;
    uint32_t* ki = kA; uint32_t* vi = vA; uint32_t* ko = kB; uint32_t* vo = vB;
    for (int pass = 0; pass < passes; ++pass) {
      const int shift = 8 * pass;
      if (tid < 256) s_base[tid] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += kVoxThreads) atomicAdd(&s_base[(ki[i] >> shift) & 255u], 1u);
      __syncthreads();
      if (wid == 0) {                                            // exclusive scan of the 256 digit counts
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = s_base[lane * 8 + j]; sum += c[j]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
        uint32_t run = inc - sum;
#pragma unroll
        for (int j = 0; j < 8; ++j) { s_base[lane * 8 + j] = run; run += c[j]; }
      }
      __syncthreads();
      for (int t0 = 0; t0 < n; t0 += kVoxThreads) {
        for (int j = tid; j < 32 * 256; j += kVoxThreads) (&s_cnt[0][0])[j] = 0;
        __syncthreads();
        const int i = t0 + tid;
        const bool act = i < n;
        const uint32_t key = act ? ki[i] : 0u, val = act ? vi[i] : 0u;
        const uint32_t dgt = act ? ((key >> shift) & 255u) : 256u;
        const unsigned m = __match_any_sync(0xFFFFFFFFu, dgt);
        const int lower = __popc(m & lt);
        if (act && lower == 0) s_cnt[wid][dgt] = (uint32_t)__popc(m);
        __syncthreads();
        if (tid < 256) {                                         // digit tid: positions of the warps' groups, in warp order (stable)
          uint32_t run = s_base[tid];
          for (int w = 0; w < 32; ++w) { const uint32_t c = s_cnt[w][tid]; s_cnt[w][tid] = run; run += c; }
          s_base[tid] = run;
        }
        __syncthreads();
        if (act) { const uint32_t pos = s_cnt[wid][dgt] + (uint32_t)lower; ko[pos] = key; vo[pos] = val; }
        __syncthreads();
      }
      uint32_t* t = ki; ki = ko; ko = t; t = vi; vi = vo; vo = t;
    }
    // ---- leaf heads: count them and list their positions in the other key buffer
    int nv = 0;
    for (int t0 = 0; t0 < n; t0 += kVoxThreads) {
      const int i = t0 + tid;
      const bool head = i < n && (i == 0 || ki[i] != ki[i - 1]);
      const unsigned b = __ballot_sync(0xFFFFFFFFu, head);
      if (lane == 0) s_cnt[0][wid] = (uint32_t)__popc(b);
      __syncthreads();
      int before = 0, tot = 0;
      for (int w = 0; w < 32; ++w) { const int c = (int)s_cnt[0][w]; if (w < wid) before += c; tot += c; }
      if (head) ko[nv + before + __popc(b & lt)] = (uint32_t)i;
      nv += tot;
      __syncthreads();
    }
    if (tid == 0) { seg->nvox = nv; seg->unfiltered = 0; seg->sorted_in_b = (ki == kB) ? 1 : 0; }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_voxel_centroids(const float* __restrict__ pts, const int* __restrict__ offs, const int* __restrict__ nplanes, int N,
                                                         const uint32_t* __restrict__ keyA, const uint32_t* __restrict__ valA, const uint32_t* __restrict__ keyB,
                                                         const uint32_t* __restrict__ valB, const VoxSeg* __restrict__ segs, float* __restrict__ out,
                                                         int* __restrict__ out_offs) {
  const int f = blockIdx.y, tid = threadIdx.x;
  const int np = min(nplanes[f], kMaxPlanes);
  const int* fo = offs + (long long)f * (kMaxPlanes + 1);
  const VoxSeg* fs = segs + (long long)f * kMaxPlanes;
  int* oo = out_offs + (long long)f * (kMaxPlanes + 1);
  for (int p = blockIdx.x; p < np; p += gridDim.x) {
    int start = 0;
    for (int q = 0; q < p; ++q) start += fs[q].nvox;             // (a few planes per frame)
    const VoxSeg sg = fs[p];
    if (tid == 0) { oo[p] = start; if (p == np - 1) oo[np] = start + sg.nvox; }
    const int s0 = fo[p], n = fo[p + 1] - s0;
    const float* P3 = pts + ((long long)f * N + s0) * 3;
    float* O = out + ((long long)f * N + start) * 3;
    if (sg.unfiltered) {
      for (int i = tid; i < 3 * n; i += 256) O[i] = P3[i];
      continue;
    }
    const uint32_t* vals = (sg.sorted_in_b ? valB : valA) + (long long)f * N + s0;
    const uint32_t* heads = (sg.sorted_in_b ? keyA : keyB) + (long long)f * N + s0;
    for (int j = tid; j < sg.nvox; j += 256) {
      const int a = (int)heads[j], b = j + 1 < sg.nvox ? (int)heads[j + 1] : n;
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (int i = a; i < b; ++i) { const float* q = P3 + 3 * vals[i]; sx = sx + q[0]; sy = sy + q[1]; sz = sz + q[2]; }
      const float cnt = (float)(b - a);
      O[3 * j] = __fdiv_rn(sx, cnt); O[3 * j + 1] = __fdiv_rn(sy, cnt); O[3 * j + 2] = __fdiv_rn(sz, cnt);
    }
  }
  if (np == 0 && blockIdx.x == 0 && tid == 0) oo[0] = 0;
}

// ------------------------------------------------------------------ the 1/3-resolution cloud of Frame::ComputePlanes_CAPE
// (reference src/Frame.cc:1153-1172): every third pixel of every third row, z = d > max_point_dist ? 0 : d,
// x = (n - cx) * z / fx in float arithmetic; out [frames][ceil(H/3)][ceil(W/3)][3]
template <int MODE>
__global__ void __launch_bounds__(256) k_cape_third_cloud(const CapeDev* __restrict__ Pp, int nframes, float max_point_dist, float* __restrict__ out) {
  const CapeDev& P = *Pp;
  const int w3 = (P.W + 2) / 3, h3 = (P.H + 2) / 3;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)nframes * w3 * h3) return;
  const int f = (int)(t / (w3 * h3));
  const int rem = (int)(t - (long long)f * w3 * h3);
  const int r = rem / w3, c = rem - r * w3;
  const int m = 3 * r, n = 3 * c;
  float d;
  if (MODE == 1) d = __ldg(P.depth + (long long)f * P.depth_fs + (long long)m * P.depth_rs + n);
  else d = (float)__ldg(P.depth16 + (long long)f * P.depth_fs + (long long)m * P.depth_rs + n) * P.depth_factor;
  const float z = d > max_point_dist ? 0.f : d;
  float* o = out + 3 * t;
  o[0] = __fdiv_rn(__fmul_rn(__fsub_rn((float)n, P.cx), z), P.fx);
  o[1] = __fdiv_rn(__fmul_rn(__fsub_rn((float)m, P.cy), z), P.fy);
  o[2] = z;
}

struct drfe_cape {
  int device = 0, max_batch = 0;
  drfe_cape_params prm{};
  CapeDev hd{};
  CapeDev* dd = nullptr;
  cudaStream_t stream = nullptr;
  float* d_depth = nullptr;   // staging for host depth
  size_t grid_smem = 0;
  int last_frames = 0;
  bool pending = false;
  int sm_count = 148;
  bool cloud_valid = false;      // hd.cloud holds the cell-major cloud of the last enqueue (given by the caller, or materialised on demand)
  StageTimer timer;
  ChunkPipe pipe;
  int* batch_nplanes = nullptr;  // host destination of the running batch call
  float* d_plane_pts = nullptr;  // [B][H*W][3] per-plane point lists (allocated by the first drfe_cape_plane_points)
  int* d_plane_offs = nullptr;   // [B][kMaxPlanes+1]
  std::vector<int> h_plane_offs;
  VoxelScratch vox;                      // drfe_cape_plane_points_voxel: sort buffers, per (frame, plane) records, centroids, offsets
  float* d_third = nullptr;      // drfe_cape_third_cloud
  NormalsScratch nrm;            // drfe_cape_third_cloud_normals
  int batch_plane_cap = 0;
  std::vector<void*> allocs;
};

template <typename T>
static int cape_alloc(drfe_cape* h, T** p, size_t count) {
  void* q = nullptr;
  DRFE_CUDA(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  h->allocs.push_back(q);
  *p = (T*)q;
  return DRFE_OK;
}

// glibc rand(): TYPE_3 additive feedback generator, srand(seed) then n outputs (stdlib/random_r.c).
// CylinderSeg draws its RANSAC triplets from the process-global rand() (CylinderSeg.cpp:111-113);
// the declared stream is this one, seed 1, restarted for every frame.
static void glibc_rand_table(uint32_t seed, int n, std::vector<uint32_t>& out) {
  std::vector<uint32_t> r(344 + (size_t)n);
  int32_t w = (int32_t)(seed ? seed : 1);
  r[0] = (uint32_t)w;
  for (int i = 1; i < 31; ++i) {
    const int32_t hi = w / 127773, lo = w % 127773;
    w = 16807 * lo - 2836 * hi;
    if (w < 0) w += 2147483647;
    r[i] = (uint32_t)w;
  }
  for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
  for (size_t i = 34; i < r.size(); ++i) r[i] = r[i - 31] + r[i - 3];
  out.resize(n);
  for (int i = 0; i < n; ++i) out[i] = r[344 + i] >> 1;
}

int drfe::cape_depth_view(drfe_cape* h, CapeDepthView* v) {
  if (!h || !v) return DRFE_ERR_ARG;
  v->device = h->device; v->nframes = h->last_frames; v->width = h->hd.W; v->height = h->hd.H; v->pending = h->pending; v->stream = h->stream;
  v->depth = h->hd.depth; v->depth16 = h->hd.depth16; v->factor = h->hd.depth_factor; v->row_stride = h->hd.depth_rs; v->frame_stride = h->hd.depth_fs;
  return DRFE_OK;
}

// ---- the voxel filter for handles that hold per-plane point lists on the device (drfe_internal.h)
int drfe::voxel_scratch_alloc(VoxelScratch& s, size_t B, size_t N) {
  void** slots[7] = {(void**)&s.key[0], (void**)&s.key[1], (void**)&s.val[0], (void**)&s.val[1], &s.seg, (void**)&s.out, (void**)&s.out_offs};
  const size_t bytes[7] = {B * N * 4, B * N * 4, B * N * 4, B * N * 4, B * kMaxPlanes * sizeof(VoxSeg), B * N * 3 * sizeof(float), B * (kMaxPlanes + 1) * sizeof(int)};
  for (int i = 0; i < 7; ++i)
    if (cudaMalloc(slots[i], bytes[i]) != cudaSuccess) { voxel_scratch_free(s); return 1; }
  return 0;
}
void drfe::voxel_scratch_free(VoxelScratch& s) {
  void** slots[7] = {(void**)&s.key[0], (void**)&s.key[1], (void**)&s.val[0], (void**)&s.val[1], &s.seg, (void**)&s.out, (void**)&s.out_offs};
  for (int i = 0; i < 7; ++i) { if (*slots[i]) cudaFree(*slots[i]); *slots[i] = nullptr; }
}
int drfe::voxel_filter_launch(cudaStream_t st, int nf, const float* pts, const int* offs, const int* nplanes, int N, float leaf_size, const VoxelScratch& s) {
  const float inv_leaf = 1.0f / leaf_size;                        // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
  const dim3 grid(8, nf);                                         // up to 8 CTAs share a frame's planes
  VoxSeg* seg = static_cast<VoxSeg*>(s.seg);
  DRFE_LAUNCH(k_voxel_sort, grid, kVoxThreads, 0, st, pts, offs, nplanes, N, inv_leaf, s.key[0], s.val[0], s.key[1], s.val[1], seg);
  DRFE_LAUNCH(k_voxel_centroids, grid, 256, 0, st, pts, offs, nplanes, N, s.key[0], s.val[0], s.key[1], s.val[1], seg, s.out, s.out_offs);
  return DRFE_OK;
}


extern "C" {

int drfe_cape_create(const drfe_cape_params* pr, int max_batch, int device, drfe_cape** out) {
  if (!pr || !out) { set_error("drfe_cape_create: null argument"); return DRFE_ERR_ARG; }
  *out = nullptr;
  if (pr->depth_height < 1 || pr->depth_width < 1 || pr->cell_width < 2 || pr->cell_height < 2 || max_batch < 1 ||
      pr->cell_width * pr->cell_height < 16 || pr->depth_width / pr->cell_width < 2 || pr->depth_height / pr->cell_height < 2) {
    set_error("drfe_cape_create: invalid parameters");
    return DRFE_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("drfe_cape_create: no CUDA device available (there is no CPU fallback)");
    return DRFE_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { set_error("drfe_cape_create: bad device %d", device); return DRFE_ERR_ARG; }
  DeviceScope ds(device);
  if (!ds.ok) { set_error("cudaSetDevice(%d) failed", device); return DRFE_ERR_CUDA; }
  drfe_cape* h = new drfe_cape();
  h->device = device; h->max_batch = max_batch; h->prm = *pr;
  CapeDev& D = h->hd;
  memset(&D, 0, sizeof(D));
  D.H = pr->depth_height; D.W = pr->depth_width; D.cw = pr->cell_width; D.ch = pr->cell_height;
  D.ncx = D.W / D.cw; D.ncy = D.H / D.ch; D.ncells = D.ncx * D.ncy; D.npc = D.cw * D.ch; D.B = max_batch;
  D.min_cos = pr->min_cos_angle_4_merge; D.max_merge_dist = pr->max_merge_dist;
  { cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount; }
  const size_t N = (size_t)D.H * D.W, B = max_batch, nc = D.ncells;
  int rc = DRFE_OK;
  auto fail = [&](int code) { drfe_cape_destroy(h); return code; };
  { const char* e = getenv("DRFE_CAPE_PRIO"); if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, e ? atoi(e) : 0) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(DRFE_ERR_CUDA); } }
  rc |= cape_alloc(h, &D.cells, nc * B);
  rc |= cape_alloc(h, &D.tols, nc * B);
  rc |= cape_alloc(h, &D.cell_meta, nc * B);
  rc |= cape_alloc(h, &D.sums, nc * B);
  rc |= cape_alloc(h, &D.dbg, 16 * B);
  D.max_jobs = (int)nc / 4 + 1;
  rc |= cape_alloc(h, &D.jobacc, (size_t)D.max_jobs * 10 * B);
  rc |= cape_alloc(h, &D.jobseg, (size_t)D.max_jobs * B);
  rc |= cape_alloc(h, &D.plane_map, nc * B);
  rc |= cape_alloc(h, &D.eroded_map, nc * B);
  D.cyl = pr->cylinder_detection ? 1 : 0;
  D.border_rows = (kMaxPlanes + 1) * (D.cyl ? 2 : 1);
  rc |= cape_alloc(h, &D.border_vec, (size_t)D.border_rows * ((nc + 31) / 32) * B);
  if (D.cyl) {
    // CylinderSeg.cpp:82-84,164: float K = log(1 - p_success) / log(1 - pow(w, 3)), p_success = 0.8f, w = 0.33f then 0.5
    D.cylK1 = (float)((double)logf(1.0f - 0.8f) / log(1.0 - pow((double)0.33f, 3)));
    D.cylK2 = (float)((double)logf(1.0f - 0.8f) / log(1.0 - pow(0.5, 3)));
    D.max_sub = (int)nc / 6 + 1;
    D.rand_n = 1 << 16;
    std::vector<uint32_t> tab;
    glibc_rand_table(1u, D.rand_n, tab);
    uint32_t* d_tab = nullptr;
    rc |= cape_alloc(h, &d_tab, (size_t)D.rand_n);
    if (!rc && cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) rc = DRFE_ERR_CUDA;
    D.rand_tab = d_tab;
    rc |= cape_alloc(h, &D.cyl_scratch, (size_t)6 * nc * B);
    rc |= cape_alloc(h, &D.subs, (size_t)D.max_sub * B);
    rc |= cape_alloc(h, &D.cyl_map, nc * B);
    rc |= cape_alloc(h, &D.cyl_eroded_map, nc * B);
    rc |= cape_alloc(h, &D.cyl_eq, (size_t)(kMaxPlanes + 1) * B);
    rc |= cape_alloc(h, &D.cyls, (size_t)D.max_sub * B);
    rc |= cape_alloc(h, &D.ncyl_found, B);
    rc |= cape_alloc(h, &D.ncyl_final, B);
  }
  rc |= cape_alloc(h, &D.segs, (size_t)(kMaxPlanes + 1) * B);
  rc |= cape_alloc(h, &D.planes, (size_t)kMaxPlanes * B);
  rc |= cape_alloc(h, &D.plane_eq, (size_t)(kMaxPlanes + 1) * B);
  rc |= cape_alloc(h, &D.plane_maxd, (size_t)(kMaxPlanes + 1) * B);
  rc |= cape_alloc(h, &D.nplanes, B);
  rc |= cape_alloc(h, &D.seg, N * B);
  rc |= cape_alloc(h, &D.cell_label, nc * B);
  rc |= cape_alloc(h, &D.border_list, nc * B);
  rc |= cape_alloc(h, &D.border_count, 1);
  rc |= cape_alloc(h, &D.status, 1);
  rc |= cape_alloc(h, &h->d_depth, N * B);
  rc |= cape_alloc(h, &h->dd, 1);
  if (rc) return fail(DRFE_ERR_CUDA);
  if (cudaMemset(D.status, 0, sizeof(int)) != cudaSuccess) {
    set_error("cudaMemset failed"); return fail(DRFE_ERR_CUDA);
  }
  {
    const size_t nw = (nc + 31) / 32;
    const size_t mj = nc / 4 + 1;
    const size_t base = (size_t)BV_COUNT * nw * 4 + kHistBins * kHistBins * 4 + 256 * 8 * 4 + nc * (4 + 4 + 4 + 2 + 2 + 1) + mj * 3 + 2 +
                        mj * 2 + (D.cyl ? (mj + 1) * 2 + nc * (1 + 8 + 4) + 16 : 0) + 64;
    D.grid_sums_smem = (base + nc * 36 <= 200 * 1024) ? 1 : 0;
    h->grid_smem = base + (D.grid_sums_smem ? nc * 36 : 0);
    if (h->grid_smem > 200 * 1024 || nw > 4 * 128) { set_error("drfe_cape_create: too many cells (%zu) for the grid stage", nc); return fail(DRFE_ERR_ARG); }
  }
  if (raise_dyn_smem((k_cape_grid<128, false>), h->device, (size_t)(h->grid_smem)) != cudaSuccess ||
      raise_dyn_smem((k_cape_grid<128, true>), h->device, (size_t)(h->grid_smem)) != cudaSuccess) {
    set_error("cudaFuncSetAttribute failed"); return fail(DRFE_ERR_CUDA);
  }
  {
    const size_t sums_smem = (size_t)(kSumsThreads / 16) * 2 * D.npc * sizeof(float);
    if (sums_smem > 200 * 1024) { set_error("drfe_cape_create: cells of %d points are too large", D.npc); return fail(DRFE_ERR_ARG); }
    cudaError_t e = cudaSuccess;
#define DRFE_SUMS_ATTR(M, C) if (e == cudaSuccess) e = raise_dyn_smem((k_cape_sums<M, C>), h->device, sums_smem)
    DRFE_SUMS_ATTR(0, 0); DRFE_SUMS_ATTR(1, 0); DRFE_SUMS_ATTR(2, 0);
    DRFE_SUMS_ATTR(0, 20); DRFE_SUMS_ATTR(1, 20); DRFE_SUMS_ATTR(2, 20);
    DRFE_SUMS_ATTR(0, 10); DRFE_SUMS_ATTR(1, 10); DRFE_SUMS_ATTR(2, 10);
#undef DRFE_SUMS_ATTR
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute failed"); return fail(DRFE_ERR_CUDA); }
  }
  if (h->timer.create()) return fail(DRFE_ERR_CUDA);
  *out = h;
  return DRFE_OK;
}

int drfe_cape_destroy(drfe_cape* h) {
  if (!h) return DRFE_OK;
  DeviceScope ds(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->pipe.destroy();
  for (void* p : h->allocs) cudaFree(p);
  voxel_scratch_free(h->vox);
  normals_scratch_free(h->nrm);
  h->timer.destroy();
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return DRFE_OK;
}

void* drfe_cape_stream(drfe_cape* h) { return h ? (void*)h->stream : nullptr; }
int drfe_cape_num_cells(const drfe_cape* h, int* cx, int* cy) {
  if (!h) return DRFE_ERR_ARG;
  if (cx) *cx = h->hd.ncx;
  if (cy) *cy = h->hd.ncy;
  return DRFE_OK;
}
int drfe_cape_set_profiling(drfe_cape* h, int on) { if (!h) return DRFE_ERR_ARG; h->timer.reset(on != 0); return DRFE_OK; }
int drfe_cape_stage_times(drfe_cape* h, float* ms, const char** names, int cap, int* nstages) {
  if (!h || !ms || !nstages) return DRFE_ERR_ARG;
  DeviceScope ds(h->device);
  return h->timer.read(ms, names, cap, nstages);
}

// all kernels of frames [f0, f0 + n) on the handle's stream (the device descriptor must be current)
static int cape_launch(drfe_cape* h, int f0, int n, bool timed) {
  NvtxRange nvtx_("cape_launch");
  const bool drfe_pdl_ = pdl_enabled() && n <= pdl_max_frames();
  cudaStream_t st = h->stream;
  const int ncell_total = n * h->hd.ncells;
  const size_t sums_smem = (size_t)(kSumsThreads / 16) * 2 * h->hd.npc * sizeof(float);
  const int sums_groups = (ncell_total + kSumsCells - 1) / kSumsCells;
  const int sums_grid = (sums_groups * 16 + kSumsThreads - 1) / kSumsThreads;
  const int mode = h->hd.depth16 ? 2 : (h->hd.depth ? 1 : 0);
  const int cell = (h->hd.cw == h->hd.ch && (h->hd.cw == 20 || h->hd.cw == 10)) ? h->hd.cw : 0;
#define DRFE_SUMS(M, C) DRFE_LAUNCH_PDL((k_cape_sums<M, C>), sums_grid, kSumsThreads, sums_smem, st, h->dd, f0, n)
#define DRFE_SUMS_MODE(C) do { if (mode == 2) DRFE_SUMS(2, C); else if (mode == 1) DRFE_SUMS(1, C); else DRFE_SUMS(0, C); } while (0)
  if (cell == 20) DRFE_SUMS_MODE(20); else if (cell == 10) DRFE_SUMS_MODE(10); else DRFE_SUMS_MODE(0);
#undef DRFE_SUMS_MODE
#undef DRFE_SUMS
  if (timed) h->timer.mark("cells", st);
  DRFE_LAUNCH_PDL(k_cape_fit, (ncell_total + 127) / 128, 128, 0, st, h->dd, f0, n);
  DRFE_LAUNCH_PDL(k_cape_edges, (ncell_total + 127) / 128, 128, 0, st, h->dd, f0, n);
  if (timed) h->timer.mark("fit", st);
  if (h->hd.cyl) DRFE_LAUNCH_PDL((k_cape_grid<128, true>), n, 128, h->grid_smem, st, h->dd, f0);
  else DRFE_LAUNCH_PDL((k_cape_grid<128, false>), n, 128, h->grid_smem, st, h->dd, f0);
  if (timed) h->timer.mark("grid", st);
  {
    DRFE_CUDA(cudaMemsetAsync(h->hd.border_count, 0, sizeof(int), st));
    if (h->hd.cyl) DRFE_LAUNCH_PDL(k_cape_refine_plan<true>, (ncell_total + 255) / 256, 256, 0, st, h->dd, f0, n);
    else DRFE_LAUNCH_PDL(k_cape_refine_plan<false>, (ncell_total + 255) / 256, 256, 0, st, h->dd, f0, n);
    auto magic = [](unsigned d) { return (uint32_t)(((1ull << 32) + d - 1) / d); };
    const unsigned q = (unsigned)(h->hd.W + 3) / 4;
    DRFE_LAUNCH_PDL(k_cape_paint, dim3((q + 63) / 64, (unsigned)(h->hd.H + 4 * kPaintRows - 1) / (4 * kPaintRows), (unsigned)n), dim3(64, 4), 0, st, h->dd, f0, magic((unsigned)h->hd.cw),
                magic((unsigned)h->hd.ch));
    // warps loop over the list of border cells (its length is only known on the device): enough blocks to fill the GPU
    const int blocks = std::min((ncell_total + kBorderWarps - 1) / kBorderWarps, h->sm_count * 10);
#define DRFE_REFINE(CYL, M) DRFE_LAUNCH_PDL((k_cape_refine_border<CYL, M>), blocks, kBorderWarps * 32, 0, st, h->dd)
#define DRFE_REFINE_MODE(CYL) do { if (mode == 2) DRFE_REFINE(CYL, 2); else if (mode == 1) DRFE_REFINE(CYL, 1); else DRFE_REFINE(CYL, 0); } while (0)
    if (h->hd.cyl) DRFE_REFINE_MODE(true); else DRFE_REFINE_MODE(false);
#undef DRFE_REFINE_MODE
#undef DRFE_REFINE
  }
  if (timed) h->timer.mark("refine", st);
  return DRFE_OK;
}

// The cell-major cloud buffer (3.7 MB per 640x480 frame) exists only for callers that hand a cloud in
// (drfe_cape_enqueue_cloud) or ask for it (drfe_cape_get_cloud, drfe_cape_plane_points).
static int cape_cloud_buffer(drfe_cape* h) {
  if (h->hd.cloud) return DRFE_OK;
  const size_t n = (size_t)3 * h->hd.H * h->hd.W * h->max_batch;
  if (cape_alloc(h, &h->hd.cloud, n)) return DRFE_ERR_CUDA;
  DRFE_CUDA(cudaMemsetAsync(h->hd.cloud, 0, n * sizeof(float), h->stream));
  return DRFE_OK;
}
// organizePointCloudByCell on demand: converts the depth of the last enqueue (which must still be where the caller put it
// when it was given as a device pointer) into the cell-major cloud, once
static int cape_materialize_cloud(drfe_cape* h) {
  if (h->cloud_valid) return DRFE_OK;
  if (!h->hd.depth && !h->hd.depth16) { set_error("no depth image to build the cloud from"); return DRFE_ERR_STATE; }
  int rc = cape_cloud_buffer(h);
  if (rc != DRFE_OK) return rc;
  cudaStream_t st = h->stream;
  DRFE_CUDA(cudaMemcpyAsync(h->dd, &h->hd, sizeof(CapeDev), cudaMemcpyHostToDevice, st));
  const long long threads = (long long)h->last_frames * h->hd.H * ((h->hd.W + 3) / 4);
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  if (h->hd.depth16) DRFE_LAUNCH(k_cape_cloud<2>, blocks, 256, 0, st, h->dd, h->last_frames);
  else DRFE_LAUNCH(k_cape_cloud<1>, blocks, 256, 0, st, h->dd, h->last_frames);
  h->cloud_valid = true;
  return DRFE_OK;
}

static int cape_run(drfe_cape* h, int nframes) {
  DRFE_CUDA(cudaMemcpyAsync(h->dd, &h->hd, sizeof(CapeDev), cudaMemcpyHostToDevice, h->stream));
  const int rc = cape_launch(h, 0, nframes, true);
  if (rc != DRFE_OK) return rc;
  h->last_frames = nframes;
  h->pending = true;
  return DRFE_OK;
}

int drfe_cape_enqueue_cloud(drfe_cape* h, int nframes, const float* cloud, size_t frame_stride, int mem_kind) {
  NvtxRange nvtx_("drfe_cape_enqueue_cloud");
  if (!h || !cloud) { set_error("drfe_cape_enqueue_cloud: null argument"); return DRFE_ERR_ARG; }
  if (nframes < 1 || nframes > h->max_batch) { set_error("drfe_cape_enqueue_cloud: nframes %d outside [1,%d]", nframes, h->max_batch); return DRFE_ERR_ARG; }
  const size_t N3 = (size_t)3 * h->hd.H * h->hd.W;
  if (frame_stride < N3 && nframes > 1) { set_error("drfe_cape_enqueue_cloud: frame_stride too small"); return DRFE_ERR_ARG; }
  DeviceScope ds(h->device);
  cudaStream_t st = h->stream;
  if (cape_cloud_buffer(h) != DRFE_OK) return DRFE_ERR_CUDA;
  h->timer.begin(st);
  const cudaMemcpyKind kind = mem_kind == DRFE_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  if (mem_kind != DRFE_MEM_HOST && mem_kind != DRFE_MEM_DEVICE) { set_error("bad mem_kind"); return DRFE_ERR_ARG; }
  if (frame_stride == N3 || nframes == 1)
    DRFE_CUDA(cudaMemcpyAsync(h->hd.cloud, cloud, N3 * nframes * sizeof(float), kind, st));
  else
    DRFE_CUDA(cudaMemcpy2DAsync(h->hd.cloud, N3 * sizeof(float), cloud, frame_stride * sizeof(float), N3 * sizeof(float), nframes, kind, st));
  h->timer.mark("copy_in", st);
  h->hd.depth = nullptr; h->hd.depth16 = nullptr;
  const int rc = cape_run(h, nframes);
  h->cloud_valid = rc == DRFE_OK;
  return rc;
}

int drfe_cape_enqueue_depth(drfe_cape* h, int nframes, const float* depth, size_t row_stride, size_t frame_stride,
                            int mem_kind, float fx, float fy, float cx, float cy) {
  NvtxRange nvtx_("drfe_cape_enqueue_depth");
  if (!h || !depth) { set_error("drfe_cape_enqueue_depth: null argument"); return DRFE_ERR_ARG; }
  if (nframes < 1 || nframes > h->max_batch) { set_error("drfe_cape_enqueue_depth: nframes %d outside [1,%d]", nframes, h->max_batch); return DRFE_ERR_ARG; }
  const int W = h->hd.W, H = h->hd.H;
  if (row_stride < (size_t)W) { set_error("drfe_cape_enqueue_depth: row_stride < width"); return DRFE_ERR_ARG; }
  DeviceScope ds(h->device);
  cudaStream_t st = h->stream;
  h->timer.begin(st);
  if (mem_kind == DRFE_MEM_HOST) {
    if (row_stride == (size_t)W && (frame_stride == (size_t)W * H || nframes == 1))
      DRFE_CUDA(cudaMemcpyAsync(h->d_depth, depth, (size_t)nframes * W * H * sizeof(float), cudaMemcpyHostToDevice, st));
    else
      for (int f = 0; f < nframes; ++f)
        DRFE_CUDA(cudaMemcpy2DAsync(h->d_depth + (size_t)f * W * H, W * sizeof(float), depth + f * frame_stride,
                                    row_stride * sizeof(float), W * sizeof(float), H, cudaMemcpyHostToDevice, st));
    h->hd.depth = h->d_depth; h->hd.depth_rs = W; h->hd.depth_fs = (long long)W * H;
    h->timer.mark("h2d", st);
  } else if (mem_kind == DRFE_MEM_DEVICE) {
    h->hd.depth = depth; h->hd.depth_rs = (long long)row_stride; h->hd.depth_fs = (long long)frame_stride;
  } else { set_error("bad mem_kind"); return DRFE_ERR_ARG; }
  h->hd.fx = fx; h->hd.fy = fy; h->hd.cx = cx; h->hd.cy = cy; h->hd.depth16 = nullptr;
  h->cloud_valid = false;
  return cape_run(h, nframes);
}


int drfe_cape_enqueue_depth_u16(drfe_cape* h, int nframes, const uint16_t* depth, size_t row_stride, size_t frame_stride,
                                int mem_kind, float depth_factor, float fx, float fy, float cx, float cy) {
  NvtxRange nvtx_("drfe_cape_enqueue_depth_u16");
  if (!h || !depth) { set_error("drfe_cape_enqueue_depth_u16: null argument"); return DRFE_ERR_ARG; }
  if (nframes < 1 || nframes > h->max_batch) { set_error("drfe_cape_enqueue_depth_u16: nframes %d outside [1,%d]", nframes, h->max_batch); return DRFE_ERR_ARG; }
  const int W = h->hd.W, H = h->hd.H;
  if (row_stride < (size_t)W) { set_error("drfe_cape_enqueue_depth_u16: row_stride < width"); return DRFE_ERR_ARG; }
  DeviceScope ds(h->device);
  cudaStream_t st = h->stream;
  h->timer.begin(st);
  if (mem_kind == DRFE_MEM_HOST) {
    uint16_t* stage = reinterpret_cast<uint16_t*>(h->d_depth);
    for (int f = 0; f < nframes; ++f)
      DRFE_CUDA(cudaMemcpy2DAsync(stage + (size_t)f * W * H, W * sizeof(uint16_t), depth + f * frame_stride, row_stride * sizeof(uint16_t),
                                  W * sizeof(uint16_t), H, cudaMemcpyHostToDevice, st));
    h->hd.depth16 = stage; h->hd.depth_rs = W; h->hd.depth_fs = (long long)W * H;
    h->timer.mark("h2d", st);
  } else if (mem_kind == DRFE_MEM_DEVICE) {
    h->hd.depth16 = depth; h->hd.depth_rs = (long long)row_stride; h->hd.depth_fs = (long long)frame_stride;
  } else { set_error("bad mem_kind"); return DRFE_ERR_ARG; }
  h->hd.depth = nullptr; h->hd.depth_factor = depth_factor;
  h->hd.fx = fx; h->hd.fy = fy; h->hd.cx = cx; h->hd.cy = cy;
  h->cloud_valid = false;
  return cape_run(h, nframes);
}

int drfe_cape_process_depth_batch(drfe_cape* h, int nframes, const void* depth, int depth_is_u16, float depth_factor,
                                  size_t row_stride, size_t frame_stride, float fx, float fy, float cx, float cy,
                                  uint8_t* seg_out, drfe_plane* planes, int plane_cap, int* nr_planes,
                                  drfe_cylinder* cylinders, int cyl_cap, int* nr_cylinders, int chunk_frames) {
  NvtxRange nvtx_("drfe_cape_process_depth_batch");
  if (!h || !depth || !nr_planes) { set_error("drfe_cape_process_depth_batch: null argument"); return DRFE_ERR_ARG; }
  if (nframes < 1 || nframes > h->max_batch) { set_error("drfe_cape_process_depth_batch: nframes %d outside [1,%d]", nframes, h->max_batch); return DRFE_ERR_ARG; }
  const int W = h->hd.W, H = h->hd.H;
  if (row_stride < (size_t)W || (nframes > 1 && frame_stride < row_stride * H)) { set_error("drfe_cape_process_depth_batch: bad strides"); return DRFE_ERR_ARG; }
  if (h->pipe.active) { set_error("drfe_cape_process_depth_batch: the previous batch was not finished (drfe_cape_finish_batch)"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  if (h->pipe.create() != DRFE_OK) return DRFE_ERR_CUDA;
  ChunkPipe& pp = h->pipe;
  cudaStream_t st = h->stream;
  const size_t N = (size_t)W * H, esz = depth_is_u16 ? sizeof(uint16_t) : sizeof(float);
  int cstart[ChunkPipe::kMaxChunks + 1];
  const int nchunks = ChunkPipe::schedule(nframes, chunk_frames, cstart);
  if (depth_is_u16) { h->hd.depth16 = reinterpret_cast<const uint16_t*>(h->d_depth); h->hd.depth = nullptr; h->hd.depth_factor = depth_factor; }
  else { h->hd.depth = h->d_depth; h->hd.depth16 = nullptr; }
  h->hd.depth_rs = W; h->hd.depth_fs = (long long)N;
  h->hd.fx = fx; h->hd.fy = fy; h->hd.cx = cx; h->hd.cy = cy;
  h->cloud_valid = false;
  DRFE_CUDA(cudaMemcpyAsync(h->dd, &h->hd, sizeof(CapeDev), cudaMemcpyHostToDevice, st));
  DRFE_CUDA(cudaEventRecord(pp.ev_start, st));
  DRFE_CUDA(cudaStreamWaitEvent(pp.h2d, pp.ev_start, 0));
  DRFE_CUDA(cudaStreamWaitEvent(pp.d2h, pp.ev_start, 0));
  const bool dense = row_stride == (size_t)W && frame_stride == N;
  uint8_t* stage = reinterpret_cast<uint8_t*>(h->d_depth);
  const uint8_t* src = reinterpret_cast<const uint8_t*>(depth);
  for (int k = 0; k < nchunks; ++k) {
    const int f0 = cstart[k], n = cstart[k + 1] - f0;
    if (dense)
      DRFE_CUDA(cudaMemcpyAsync(stage + f0 * N * esz, src + (size_t)f0 * frame_stride * esz, (size_t)n * N * esz, cudaMemcpyHostToDevice, pp.h2d));
    else
      for (int f = f0; f < f0 + n; ++f)
        DRFE_CUDA(cudaMemcpy2DAsync(stage + f * N * esz, W * esz, src + (size_t)f * frame_stride * esz, row_stride * esz, W * esz, H,
                                    cudaMemcpyHostToDevice, pp.h2d));
    DRFE_CUDA(cudaEventRecord(pp.ev_in[k], pp.h2d));
    DRFE_CUDA(cudaStreamWaitEvent(st, pp.ev_in[k], 0));
    const int rc = cape_launch(h, f0, n, false);
    if (rc != DRFE_OK) { cudaStreamSynchronize(pp.h2d); cudaStreamSynchronize(st); cudaStreamSynchronize(pp.d2h); return rc; }   // nothing stays queued on the caller's buffers
    DRFE_CUDA(cudaEventRecord(pp.ev_done[k], st));
    DRFE_CUDA(cudaStreamWaitEvent(pp.d2h, pp.ev_done[k], 0));
    DRFE_CUDA(cudaMemcpyAsync(nr_planes + f0, h->hd.nplanes + f0, n * sizeof(int), cudaMemcpyDeviceToHost, pp.d2h));
    if (seg_out) DRFE_CUDA(cudaMemcpyAsync(seg_out + f0 * N, h->hd.seg + f0 * N, N * n, cudaMemcpyDeviceToHost, pp.d2h));
    if (planes && plane_cap > 0)
      DRFE_CUDA(cudaMemcpy2DAsync(planes + (size_t)f0 * plane_cap, (size_t)plane_cap * sizeof(drfe_plane), h->hd.planes + (size_t)f0 * kMaxPlanes,
                                  (size_t)kMaxPlanes * sizeof(drfe_plane), (size_t)std::min(plane_cap, kMaxPlanes) * sizeof(drfe_plane), n,
                                  cudaMemcpyDeviceToHost, pp.d2h));
    if (h->hd.cyl) {
      if (nr_cylinders) DRFE_CUDA(cudaMemcpyAsync(nr_cylinders + f0, h->hd.ncyl_final + f0, n * sizeof(int), cudaMemcpyDeviceToHost, pp.d2h));
      if (cylinders && cyl_cap > 0)
        DRFE_CUDA(cudaMemcpy2DAsync(cylinders + (size_t)f0 * cyl_cap, (size_t)cyl_cap * sizeof(drfe_cylinder), h->hd.cyls + (size_t)f0 * h->hd.max_sub,
                                    (size_t)h->hd.max_sub * sizeof(drfe_cylinder), (size_t)std::min(cyl_cap, h->hd.max_sub) * sizeof(drfe_cylinder), n,
                                    cudaMemcpyDeviceToHost, pp.d2h));
    }
  }
  DRFE_CUDA(cudaMemcpyAsync(pp.h_status, h->hd.status, sizeof(int), cudaMemcpyDeviceToHost, pp.d2h));
  if (!h->hd.cyl && nr_cylinders) for (int f = 0; f < nframes; ++f) nr_cylinders[f] = 0;
  pp.active = true;
  h->batch_nplanes = nr_planes; h->batch_plane_cap = planes ? plane_cap : 0x7FFFFFFF;
  h->last_frames = nframes;
  h->pending = true;
  return DRFE_OK;
}

int drfe_cape_finish_batch(drfe_cape* h) {
  NvtxRange nvtx_("drfe_cape_finish_batch");
  if (!h) { set_error("drfe_cape_finish_batch: null handle"); return DRFE_ERR_ARG; }
  if (!h->pipe.active) { set_error("drfe_cape_finish_batch: no batch in flight"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  DRFE_CUDA(cudaStreamSynchronize(h->pipe.d2h));
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  h->pipe.active = false;
  if (*h->pipe.h_status) {
    const int stv = *h->pipe.h_status;
    *h->pipe.h_status = 0;
    DRFE_CUDA(cudaMemsetAsync(h->hd.status, 0, sizeof(int), h->stream));
    set_error("CAPE: device-side capacity exceeded (status %d)", stv);
    return DRFE_ERR_CAPACITY;
  }
  for (int f = 0; f < h->last_frames; ++f)
    if (h->batch_nplanes[f] > h->batch_plane_cap) {
      set_error("drfe_cape_finish_batch: frame %d has %d planes, plane_cap is %d", f, h->batch_nplanes[f], h->batch_plane_cap);
      return DRFE_ERR_CAPACITY;
    }
  return DRFE_OK;
}

int drfe_cape_sync(drfe_cape* h) {
  if (!h) return DRFE_ERR_ARG;
  DeviceScope ds(h->device);
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  return DRFE_OK;
}

int drfe_cape_download(drfe_cape* h, uint8_t* seg_out, drfe_plane* planes, int plane_cap, int* nr_planes,
                       drfe_cylinder* cylinders, int cyl_cap, int* nr_cylinders) {
  NvtxRange nvtx_("drfe_cape_download");
  if (!h || !nr_planes) { set_error("drfe_cape_download: null argument"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("drfe_cape_download: nothing enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  cudaStream_t st = h->stream;
  const int nf = h->last_frames;
  const size_t N = (size_t)h->hd.H * h->hd.W;
  int status = 0;
  DRFE_CUDA(cudaMemcpyAsync(nr_planes, h->hd.nplanes, nf * sizeof(int), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaMemcpyAsync(&status, h->hd.status, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (seg_out) DRFE_CUDA(cudaMemcpyAsync(seg_out, h->hd.seg, N * nf, cudaMemcpyDeviceToHost, st));
  if (planes && plane_cap > 0) {
    const size_t w = (size_t)std::min(plane_cap, kMaxPlanes) * sizeof(drfe_plane);
    DRFE_CUDA(cudaMemcpy2DAsync(planes, (size_t)plane_cap * sizeof(drfe_plane), h->hd.planes, (size_t)kMaxPlanes * sizeof(drfe_plane),
                                w, nf, cudaMemcpyDeviceToHost, st));
  }
  if (h->hd.cyl) {
    if (nr_cylinders) DRFE_CUDA(cudaMemcpyAsync(nr_cylinders, h->hd.ncyl_final, nf * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (cylinders && cyl_cap > 0) {
      const size_t w = (size_t)std::min(cyl_cap, h->hd.max_sub) * sizeof(drfe_cylinder);
      DRFE_CUDA(cudaMemcpy2DAsync(cylinders, (size_t)cyl_cap * sizeof(drfe_cylinder), h->hd.cyls, (size_t)h->hd.max_sub * sizeof(drfe_cylinder),
                                  w, nf, cudaMemcpyDeviceToHost, st));
    }
  }
  DRFE_CUDA(cudaStreamSynchronize(st));
  if (!h->hd.cyl && nr_cylinders) for (int f = 0; f < nf; ++f) nr_cylinders[f] = 0;
  if (status) {
    cudaMemsetAsync(h->hd.status, 0, sizeof(int), st);
    set_error("CAPE: device-side capacity exceeded (status %d: 1 = more than %d planes, 2 = seed rejected, 4 = regions, "
              "8 = rand() table, 16 = cylinder sub-segments, 32 = cylinder labels)", status, kMaxPlanes);
    return DRFE_ERR_CAPACITY;
  }
  if (planes)
    for (int f = 0; f < nf; ++f)
      if (nr_planes[f] > plane_cap) { set_error("drfe_cape_download: frame %d has %d planes, plane_cap is %d", f, nr_planes[f], plane_cap); return DRFE_ERR_CAPACITY; }
  return DRFE_OK;
}

// the per-plane point lists of the last batch on the device (h->d_plane_pts / d_plane_offs), no copy out
static int cape_plane_points_device(drfe_cape* h) {
  cudaStream_t st = h->stream;
  const int nf = h->last_frames;
  const size_t N = (size_t)h->hd.H * h->hd.W;
  if (h->pipe.active) { set_error("a batch call is in flight (drfe_cape_finish_batch first)"); return DRFE_ERR_STATE; }
  { const int rc = cape_materialize_cloud(h); if (rc != DRFE_OK) return rc; }
  if (!h->d_plane_pts) {
    if (cape_alloc(h, &h->d_plane_pts, (size_t)h->max_batch * N * 3)) return DRFE_ERR_CUDA;
    if (cape_alloc(h, &h->d_plane_offs, (size_t)h->max_batch * (kMaxPlanes + 1))) return DRFE_ERR_CUDA;
    h->h_plane_offs.resize((size_t)h->max_batch * (kMaxPlanes + 1));
    DRFE_CUDA(raise_dyn_smem(k_cape_plane_points, h->device, (size_t)(kPtsWarps * (kMaxPlanes + 1) * (int)sizeof(int))));
  }
  auto magic = [](unsigned d) { return (uint32_t)(((1ull << 32) + d - 1) / d); };
  DRFE_LAUNCH(k_cape_plane_points, nf, kPtsWarps * 32, kPtsWarps * (kMaxPlanes + 1) * sizeof(int), st, h->dd, 0, h->d_plane_pts, h->d_plane_offs,
              magic((unsigned)h->hd.W), magic((unsigned)h->hd.cw), magic((unsigned)h->hd.ch));
  return DRFE_OK;
}

// offsets [nf][256] and points [nf][N][3] on the device -> the caller's arrays
static int cape_points_out(drfe_cape* h, const char* who, const float* d_pts, const int* d_offs, float* points, size_t cap_per_frame, int* offsets, int plane_cap) {
  cudaStream_t st = h->stream;
  const int nf = h->last_frames;
  const size_t N = (size_t)h->hd.H * h->hd.W;
  int* ho = h->h_plane_offs.data();
  DRFE_CUDA(cudaMemcpyAsync(ho, d_offs, (size_t)nf * (kMaxPlanes + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  std::vector<int> np(nf);
  DRFE_CUDA(cudaMemcpyAsync(np.data(), h->hd.nplanes, nf * sizeof(int), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaStreamSynchronize(st));
  for (int f = 0; f < nf; ++f) {
    const int n = std::min(np[f], kMaxPlanes);
    if (n > plane_cap) { set_error("%s: frame %d has %d planes, plane_cap is %d", who, f, n, plane_cap); return DRFE_ERR_CAPACITY; }
    const int* src = ho + (size_t)f * (kMaxPlanes + 1);
    int* dst = offsets + (size_t)f * (plane_cap + 1);
    for (int i = 0; i <= n; ++i) dst[i] = src[i];
    for (int i = n + 1; i <= plane_cap; ++i) dst[i] = src[n];
    if ((size_t)src[n] > cap_per_frame) { set_error("%s: frame %d has %d plane points, cap_per_frame is %zu", who, f, src[n], cap_per_frame); return DRFE_ERR_CAPACITY; }
    if (src[n] > 0)
      DRFE_CUDA(cudaMemcpyAsync(points + (size_t)f * cap_per_frame * 3, d_pts + (size_t)f * N * 3, (size_t)src[n] * 3 * sizeof(float),
                                cudaMemcpyDeviceToHost, st));
  }
  DRFE_CUDA(cudaStreamSynchronize(st));
  return DRFE_OK;
}

int drfe_cape_plane_points(drfe_cape* h, float* points, size_t cap_per_frame, int* offsets, int plane_cap) {
  NvtxRange nvtx_("drfe_cape_plane_points");
  if (!h || !points || !offsets || plane_cap < 1) { set_error("drfe_cape_plane_points: bad argument"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("drfe_cape_plane_points: nothing enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  const int rc = cape_plane_points_device(h);
  if (rc != DRFE_OK) return rc;
  return cape_points_out(h, "drfe_cape_plane_points", h->d_plane_pts, h->d_plane_offs, points, cap_per_frame, offsets, plane_cap);
}

int drfe_cape_plane_points_voxel(drfe_cape* h, float leaf_size, float* points, size_t cap_per_frame, int* offsets, int plane_cap) {
  NvtxRange nvtx_("drfe_cape_plane_points_voxel");
  if (!h || !points || !offsets || plane_cap < 1 || !(leaf_size > 0.f)) { set_error("drfe_cape_plane_points_voxel: bad argument"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("drfe_cape_plane_points_voxel: nothing enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  int rc = cape_plane_points_device(h);
  if (rc != DRFE_OK) return rc;
  const size_t N = (size_t)h->hd.H * h->hd.W;
  if (!h->vox.out) {
    if (voxel_scratch_alloc(h->vox, (size_t)h->max_batch, N)) { set_error("drfe_cape_plane_points_voxel: cudaMalloc failed"); return DRFE_ERR_CUDA; }
  }
  rc = voxel_filter_launch(h->stream, h->last_frames, h->d_plane_pts, h->d_plane_offs, h->hd.nplanes, (int)N, leaf_size, h->vox);
  if (rc != DRFE_OK) return rc;
  return cape_points_out(h, "drfe_cape_plane_points_voxel", h->vox.out, h->vox.out_offs, points, cap_per_frame, offsets, plane_cap);
}

int drfe_cape_third_cloud(drfe_cape* h, float max_point_dist, float* cloud) {
  NvtxRange nvtx_("drfe_cape_third_cloud");
  if (!h || !cloud) { set_error("drfe_cape_third_cloud: null argument"); return DRFE_ERR_ARG; }
  if (!h->pending || (!h->hd.depth && !h->hd.depth16)) { set_error("drfe_cape_third_cloud: no depth image enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  const size_t w3 = (h->hd.W + 2) / 3, h3 = (h->hd.H + 2) / 3, per = w3 * h3 * 3;
  if (!h->d_third && cape_alloc(h, &h->d_third, per * h->max_batch)) return DRFE_ERR_CUDA;
  cudaStream_t st = h->stream;
  const int nf = h->last_frames;
  DRFE_CUDA(cudaMemcpyAsync(h->dd, &h->hd, sizeof(CapeDev), cudaMemcpyHostToDevice, st));
  const unsigned blocks = (unsigned)((w3 * h3 * nf + 255) / 256);
  if (h->hd.depth16) DRFE_LAUNCH(k_cape_third_cloud<2>, blocks, 256, 0, st, h->dd, nf, max_point_dist, h->d_third);
  else DRFE_LAUNCH(k_cape_third_cloud<1>, blocks, 256, 0, st, h->dd, nf, max_point_dist, h->d_third);
  DRFE_CUDA(cudaMemcpyAsync(cloud, h->d_third, per * nf * sizeof(float), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaStreamSynchronize(st));
  return DRFE_OK;
}

// the surface normals Frame::ComputePlanes(_CAPE) takes from pcl::IntegralImageNormalEstimation on that cloud (Frame.cc:1174-1216)
int drfe_cape_third_cloud_normals(drfe_cape* h, float max_point_dist, float max_depth_change_factor, float normal_smoothing_size, float* cloud, float* normals) {
  NvtxRange nvtx_("drfe_cape_third_cloud_normals");
  if (!h || !normals || !(normal_smoothing_size >= 1.f)) { set_error("drfe_cape_third_cloud_normals: bad argument"); return DRFE_ERR_ARG; }
  if (!h->pending || (!h->hd.depth && !h->hd.depth16)) { set_error("drfe_cape_third_cloud_normals: no depth image enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  const int w3 = (h->hd.W + 2) / 3, h3 = (h->hd.H + 2) / 3;
  const size_t per = (size_t)w3 * h3 * 3;
  if (2 * (int)normal_smoothing_size >= std::min(w3, h3)) { set_error("drfe_cape_third_cloud_normals: smoothing size %g leaves no interior", (double)normal_smoothing_size); return DRFE_ERR_ARG; }
  if (!h->d_third && cape_alloc(h, &h->d_third, per * h->max_batch)) return DRFE_ERR_CUDA;
  if (!h->nrm.normals && normals_scratch_alloc(h->nrm, (size_t)h->max_batch, w3, h3)) { set_error("drfe_cape_third_cloud_normals: cudaMalloc failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  const int nf = h->last_frames;
  DRFE_CUDA(cudaMemcpyAsync(h->dd, &h->hd, sizeof(CapeDev), cudaMemcpyHostToDevice, st));
  const unsigned blocks = (unsigned)(((size_t)w3 * h3 * nf + 255) / 256);
  if (h->hd.depth16) DRFE_LAUNCH(k_cape_third_cloud<2>, blocks, 256, 0, st, h->dd, nf, max_point_dist, h->d_third);
  else DRFE_LAUNCH(k_cape_third_cloud<1>, blocks, 256, 0, st, h->dd, nf, max_point_dist, h->d_third);
  int rc = normals_launch(st, h->device, nf, h->d_third, w3, h3, max_depth_change_factor, normal_smoothing_size, h->nrm);
  if (rc != DRFE_OK) return rc;
  if (cloud) DRFE_CUDA(cudaMemcpyAsync(cloud, h->d_third, per * nf * sizeof(float), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaMemcpyAsync(normals, h->nrm.normals, per * nf * sizeof(float), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaStreamSynchronize(st));
  return DRFE_OK;
}

int drfe_cape_process(drfe_cape* h, const float* cloud, uint8_t* seg_out, drfe_plane* planes, int plane_cap,
                      int* nr_planes, drfe_cylinder* cylinders, int cyl_cap, int* nr_cylinders) {
  if (!h) { set_error("null handle"); return DRFE_ERR_ARG; }
  int rc = drfe_cape_enqueue_cloud(h, 1, cloud, (size_t)3 * h->hd.H * h->hd.W, DRFE_MEM_HOST);
  if (rc) return rc;
  return drfe_cape_download(h, seg_out, planes, plane_cap, nr_planes, cylinders, cyl_cap, nr_cylinders);
}

int drfe_cape_process_depth(drfe_cape* h, const float* depth, size_t row_stride, float fx, float fy, float cx, float cy,
                            uint8_t* seg_out, drfe_plane* planes, int plane_cap, int* nr_planes,
                            drfe_cylinder* cylinders, int cyl_cap, int* nr_cylinders) {
  if (!h) { set_error("null handle"); return DRFE_ERR_ARG; }
  int rc = drfe_cape_enqueue_depth(h, 1, depth, row_stride, row_stride * h->hd.H, DRFE_MEM_HOST, fx, fy, cx, cy);
  if (rc) return rc;
  return drfe_cape_download(h, seg_out, planes, plane_cap, nr_planes, cylinders, cyl_cap, nr_cylinders);
}

static int cape_frame_ok(drfe_cape* h, int frame) {
  if (!h) { set_error("null handle"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("nothing enqueued"); return DRFE_ERR_STATE; }
  if (frame < 0 || frame >= h->last_frames) { set_error("bad frame"); return DRFE_ERR_ARG; }
  return DRFE_OK;
}

int drfe_cape_get_cloud(drfe_cape* h, int frame, float* cloud) {
  int rc = cape_frame_ok(h, frame);
  if (rc) return rc;
  DeviceScope ds(h->device);
  const size_t N3 = (size_t)3 * h->hd.H * h->hd.W;
  rc = cape_materialize_cloud(h);
  if (rc) return rc;
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  DRFE_CUDA(cudaMemcpy(cloud, h->hd.cloud + N3 * frame, N3 * sizeof(float), cudaMemcpyDeviceToHost));
  return DRFE_OK;
}
int drfe_cape_get_cells(drfe_cape* h, int frame, drfe_plane* cells) {
  int rc = cape_frame_ok(h, frame);
  if (rc) return rc;
  DeviceScope ds(h->device);
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  DRFE_CUDA(cudaMemcpy(cells, h->hd.cells + (size_t)h->hd.ncells * frame, (size_t)h->hd.ncells * sizeof(drfe_plane), cudaMemcpyDeviceToHost));
  return DRFE_OK;
}
int drfe_cape_debug_counters(drfe_cape* h, int frame, long long* out16) {
  int rc = cape_frame_ok(h, frame);
  if (rc) return rc;
  DeviceScope ds(h->device);
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  DRFE_CUDA(cudaMemcpy(out16, h->hd.dbg + 16 * frame, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return DRFE_OK;
}
int drfe_cape_cylinders_found(drfe_cape* h, int* counts) {
  if (!h || !counts) { set_error("drfe_cape_cylinders_found: null argument"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("drfe_cape_cylinders_found: nothing enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!h->hd.cyl) { for (int f = 0; f < h->last_frames; ++f) counts[f] = 0; return DRFE_OK; }
  DRFE_CUDA(cudaMemcpyAsync(counts, h->hd.ncyl_found, h->last_frames * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  return DRFE_OK;
}
int drfe_cape_get_cyl_maps(drfe_cape* h, int frame, int32_t* cyl_map, uint8_t* cyl_eroded_map) {
  int rc = cape_frame_ok(h, frame);
  if (rc) return rc;
  if (!h->hd.cyl) { set_error("drfe_cape_get_cyl_maps: cylinder detection is off"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  const size_t nc = h->hd.ncells;
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  if (cyl_map) DRFE_CUDA(cudaMemcpy(cyl_map, h->hd.cyl_map + nc * frame, nc * sizeof(int), cudaMemcpyDeviceToHost));
  if (cyl_eroded_map) DRFE_CUDA(cudaMemcpy(cyl_eroded_map, h->hd.cyl_eroded_map + nc * frame, nc, cudaMemcpyDeviceToHost));
  return DRFE_OK;
}
int drfe_cape_get_grid_maps(drfe_cape* h, int frame, int32_t* plane_map, uint8_t* eroded_map) {
  int rc = cape_frame_ok(h, frame);
  if (rc) return rc;
  DeviceScope ds(h->device);
  const size_t nc = h->hd.ncells;
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  if (plane_map) DRFE_CUDA(cudaMemcpy(plane_map, h->hd.plane_map + nc * frame, nc * sizeof(int), cudaMemcpyDeviceToHost));
  if (eroded_map) DRFE_CUDA(cudaMemcpy(eroded_map, h->hd.eroded_map + nc * frame, nc, cudaMemcpyDeviceToHost));
  return DRFE_OK;
}

}  // extern "C"
