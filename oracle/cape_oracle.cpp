// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.
//
// CPU restatement of DR-SLAM's CAPE plane extractor:
//   PlaneDetection_CAPE::runPlaneDetection  src/PlaneExtractor.cpp:111-191
//   CAPE::CAPE / CAPE::process              src/CAPE/CAPE.cpp:8-457
//   CAPE::getConnectedComponents / RegionGrowing  CAPE.cpp:459-506
//   PlaneSeg                                src/CAPE/PlaneSeg.cpp:8-142
//   Histogram                               src/CAPE/Histogram.cpp:8-70
// Dependency-free C++17.  Eigen and OpenCV (third-party, not vendored in the
// reference; README.md:27-33) own some of the arithmetic; their behaviour is
// restated under the following DECLARED rules:
//   B.1  MatrixXf::sum() float reductions: 16 lane accumulators (2 AVX packets,
//        stride 16), lanes i and i+8 added, then an 8->4->2->1 halving tree;
//        products rounded to float before accumulation (no FMA).
//   B.2  PlaneSeg storage is zero-filled, so the stray read of Grid[i]->MSE at
//        CAPE.cpp:130 sees 0 for cells that returned before fitPlane().
//   B.11 SelfAdjointEigenSolver<Matrix3d> is replaced by a cyclic Jacobi solver
//        (eig3_sym below) — same eigenpairs to ~1e-15; the GPU runs the same
//        operation sequence so segmentation decisions are bit-identical.
//   cv::erode / cv::dilate 3x3 with the default (ignore-outside) border.
//   B.9  CylinderSeg (src/CAPE/CylinderSeg.cpp:7-247) draws its RANSAC triplets from the
//        process-global rand(); declared: glibc TYPE_3 rand() seeded with 1 and restarted
//        at every process() call (GlibcRand below).  Eigen reductions in it are restated
//        as: dynamic 3-vectors (x0+x1)+x2, fixed Vector3 x0+(x1+x2) (Eigen 3.3 Redux.h),
//        N*N^T as 2*(sequential sum over the activated cells), no FMA.  The PCA axis sign
//        is solver-defined (compare up to sign); nothing downstream depends on it.
// PARITY STATUS: "parity unpinned" by the reference (no tests / fixtures, cannot
// be built here: needs Eigen + OpenCV headers).
#include <algorithm>
#include <array>
#include <climits>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "drfe_oracle.h"

namespace {

const double kDepthSigmaCoeff = 0.000001425;  // Params.h:6
const double kDepthSigmaMargin = 10;          // Params.h:7

// Symmetric 3x3 eigen-decomposition, cyclic Jacobi, plain IEEE double ops only
// (+,-,*,/,sqrt,fabs) so that the CUDA mirror reproduces it bit for bit.
// in: a = {xx,xy,xz,yy,yz,zz}.  out: w ascending, v[k][i] = component k of evec i.
void eig3_sym(const double in[6], double w[3], double v[3][3]) {
  double a[3][3] = {{in[0], in[1], in[2]}, {in[1], in[3], in[4]}, {in[2], in[4], in[5]}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  static const int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2}, R[3] = {2, 1, 0};
  for (int sweep = 0; sweep < 24; ++sweep) {
    if (a[0][1] == 0.0 && a[0][2] == 0.0 && a[1][2] == 0.0) break;
    for (int k = 0; k < 3; ++k) {
      const int p = P[k], q = Q[k], r = R[k];
      const double apq = a[p][q];
      if (apq == 0.0) continue;
      const double app = a[p][p], aqq = a[q][q];
      const double g = 100.0 * std::fabs(apq);
      if (sweep > 3 && std::fabs(app) + g == std::fabs(app) && std::fabs(aqq) + g == std::fabs(aqq)) {
        a[p][q] = a[q][p] = 0.0;
        continue;
      }
      const double h = aqq - app;
      double t;
      if (std::fabs(h) + g == std::fabs(h)) {
        t = apq / h;
      } else {
        const double theta = 0.5 * h / apq;
        t = 1.0 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
        if (theta < 0.0) t = -t;
      }
      const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
      a[p][p] = app - t * apq;
      a[q][q] = aqq + t * apq;
      a[p][q] = a[q][p] = 0.0;
      const double arp = a[r][p], arq = a[r][q];
      a[r][p] = a[p][r] = c * arp - s * arq;
      a[r][q] = a[q][r] = s * arp + c * arq;
      for (int m = 0; m < 3; ++m) {
        const double vp = v[m][p], vq = v[m][q];
        v[m][p] = c * vp - s * vq;
        v[m][q] = s * vp + c * vq;
      }
    }
  }
  int idx[3] = {0, 1, 2};
  double d[3] = {a[0][0], a[1][1], a[2][2]};
  // stable 3-element sort, ascending
  if (d[idx[1]] < d[idx[0]]) std::swap(idx[0], idx[1]);
  if (d[idx[2]] < d[idx[1]]) std::swap(idx[1], idx[2]);
  if (d[idx[1]] < d[idx[0]]) std::swap(idx[0], idx[1]);
  double vv[3][3];
  for (int i = 0; i < 3; ++i) {
    w[i] = d[idx[i]];
    for (int m = 0; m < 3; ++m) vv[m][i] = v[m][idx[i]];
  }
  memcpy(v, vv, sizeof(vv));
}

// Declared float reduction tree of Eigen::MatrixXf::sum() (B.1) over n values
// produced by f(i); n must be a multiple of 16 for the pure packet path, the
// remainder (Eigen's scalar tail) is added sequentially afterwards.
template <typename F>
float eigen_sum_f32(int n, F f) {
  float lane[16];
  const int body = (n / 16) * 16;
  if (body == 0) {
    float s = 0.f;  // Eigen: res = coeff(0); then += ; (0 + x == x exactly)
    for (int i = 0; i < n; ++i) s = (i == 0) ? f(0) : s + f(i);
    return s;
  }
  for (int l = 0; l < 16; ++l) lane[l] = f(l);
  for (int i = 16; i < body; i += 16)
    for (int l = 0; l < 16; ++l) lane[l] = lane[l] + f(i + l);
  float p8[8], p4[4], p2[2];
  for (int l = 0; l < 8; ++l) p8[l] = lane[l] + lane[l + 8];
  int rest = body;
  if (n - body >= 8) {  // one more aligned packet
    for (int l = 0; l < 8; ++l) p8[l] = p8[l] + f(body + l);
    rest = body + 8;
  }
  for (int l = 0; l < 4; ++l) p4[l] = p8[l] + p8[l + 4];
  for (int l = 0; l < 2; ++l) p2[l] = p4[l] + p4[l + 2];
  float s = p2[0] + p2[1];
  for (int i = rest; i < n; ++i) s = s + f(i);
  return s;
}

struct Seg : drfe_plane {
  Seg() { memset(static_cast<drfe_plane*>(this), 0, sizeof(drfe_plane)); }

  // PlaneSeg::fitPlane (PlaneSeg.cpp:111-142)
  void fit() {
    mean[0] = x_acc / nr_pts; mean[1] = y_acc / nr_pts; mean[2] = z_acc / nr_pts;
    double cov[6] = {xx_acc - x_acc * x_acc / nr_pts, xy_acc - x_acc * y_acc / nr_pts,
                     xz_acc - x_acc * z_acc / nr_pts, yy_acc - y_acc * y_acc / nr_pts,
                     yz_acc - y_acc * z_acc / nr_pts, zz_acc - z_acc * z_acc / nr_pts};
    double w[3], v[3][3];
    eig3_sym(cov, w, v);
    double v0 = v[0][0], v1 = v[1][0], v2 = v[2][0];
    d = -(v0 * mean[0] + v1 * mean[1] + v2 * mean[2]);
    if (d > 0) { normal[0] = v0; normal[1] = v1; normal[2] = v2; }
    else { normal[0] = -v0; normal[1] = -v1; normal[2] = -v2; d = -d; }
    MSE = (float)(w[0] / nr_pts);
    score = (float)(w[1] / w[0]);
  }
  // PlaneSeg::expandSegment (PlaneSeg.cpp:96-101)
  void expand(const Seg& o) {
    x_acc += o.x_acc; y_acc += o.y_acc; z_acc += o.z_acc;
    xx_acc += o.xx_acc; yy_acc += o.yy_acc; zz_acc += o.zz_acc;
    xy_acc += o.xy_acc; xz_acc += o.xz_acc; yz_acc += o.yz_acc;
    nr_pts += o.nr_pts;
  }
};

// PlaneSeg::PlaneSeg (PlaneSeg.cpp:8-94) on one cell of the cell-major cloud.
Seg fit_cell(const float* X, const float* Y, const float* Z, int npts, int cell_w) {
  Seg s;
  s.min_nr_pts = npts / 2;
  const int cell_h = npts / cell_w;
  const double max_diff = 100;
  s.planar = 1;
  int cnt = 0;
  for (int i = 0; i < npts; ++i) cnt += (Z[i] > 0);
  s.nr_pts = cnt;
  if (s.nr_pts < s.min_nr_pts) { s.planar = 0; return s; }
  // horizontal scan through the middle row
  {
    int jumps = 0;
    int i = cell_w * (cell_h / 2), j = i + cell_w;
    float z_last = std::max(Z[i], Z[i + 1]);
    ++i;
    while (i < j) {
      float z = Z[i];
      if (z > 0 && std::fabs(z - z_last) < max_diff) z_last = z;
      else if (z > 0) ++jumps;
      ++i;
    }
    if (jumps > 1) { s.planar = 0; return s; }
  }
  // vertical scan through the middle column
  {
    int jumps = 0;
    int i = cell_w / 2, j = npts - i;
    float z_last = std::max(Z[i], Z[i + cell_w]);
    i += cell_w;
    while (i < j) {
      float z = Z[i];
      if (z > 0 && std::fabs(z - z_last) < max_diff) z_last = z;
      else if (z > 0) ++jumps;
      i += cell_w;
    }
    if (jumps > 1) { s.planar = 0; return s; }
  }
  s.x_acc = eigen_sum_f32(npts, [&](int i) { return X[i]; });
  s.y_acc = eigen_sum_f32(npts, [&](int i) { return Y[i]; });
  s.z_acc = eigen_sum_f32(npts, [&](int i) { return Z[i]; });
  s.xx_acc = eigen_sum_f32(npts, [&](int i) { return X[i] * X[i]; });
  s.yy_acc = eigen_sum_f32(npts, [&](int i) { return Y[i] * Y[i]; });
  s.zz_acc = eigen_sum_f32(npts, [&](int i) { return Z[i] * Z[i]; });
  s.xy_acc = eigen_sum_f32(npts, [&](int i) { return X[i] * Y[i]; });
  s.xz_acc = eigen_sum_f32(npts, [&](int i) { return X[i] * Z[i]; });
  s.yz_acc = eigen_sum_f32(npts, [&](int i) { return Y[i] * Z[i]; });
  s.fit();
  const double lim = kDepthSigmaCoeff * s.mean[2] * s.mean[2] + kDepthSigmaMargin;
  if ((double)s.MSE > lim * lim) s.planar = 0;
  return s;
}

// glibc rand(): TYPE_3 additive feedback generator (r[i] = r[i-31] + r[i-3], output >> 1),
// seeded like srand(seed) (stdlib/random_r.c).  Checked against libc in tests/.
struct GlibcRand {
  uint32_t r[34];
  int pos = 0;
  explicit GlibcRand(uint32_t seed = 1) {
    std::vector<uint32_t> t(344);
    int32_t w = (int32_t)(seed ? seed : 1);
    t[0] = (uint32_t)w;
    for (int i = 1; i < 31; ++i) {
      const int32_t hi = w / 127773, lo = w % 127773;
      w = 16807 * lo - 2836 * hi;
      if (w < 0) w += 2147483647;
      t[i] = (uint32_t)w;
    }
    for (int i = 31; i < 34; ++i) t[i] = t[i - 31];
    for (int i = 34; i < 344; ++i) t[i] = t[i - 31] + t[i - 3];
    for (int i = 0; i < 34; ++i) r[i] = t[310 + i];
  }
  int next() {  // r[] is a ring of the last 34 values, pos = oldest
    const uint32_t v = r[(pos + 3) % 34] + r[(pos + 31) % 34];
    r[pos] = v;
    pos = (pos + 1) % 34;
    return (int)(v >> 1);
  }
};

const double kCylScoreMin = 100;            // Params.h:8
const double kCylSqrMaxDist = 0.0225;       // Params.h:9

// CylinderSeg::CylinderSeg (CylinderSeg.cpp:7-247)
struct CylSeg {
  int nr_segments = 0;
  double axis[3] = {0, 0, 0};
  std::vector<float> radii;
  std::vector<std::array<double, 3>> centers;
  std::vector<std::vector<uint8_t>> inliers;   // [segment][local cell]
  std::vector<double> MSEs;
  std::vector<int> local2global;
  std::vector<uint8_t> cylindrical;
  std::vector<std::array<float, 3>> P1, P2;
  std::vector<float> P1P2_norm;

  static double dot3(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

  CylSeg(const std::vector<Seg>& grid, const std::vector<uint8_t>& act, int m, GlibcRand& rng) {
    const int ns = (int)grid.size();
    std::vector<double> N(3 * m), P(3 * m);   // column j at [3*j .. 3*j+2]
    local2global.resize(m);
    int j = 0;
    for (int i = 0; i < ns; ++i)
      if (act[i]) {
        for (int k = 0; k < 3; ++k) { N[3 * j + k] = grid[i].normal[k]; P[3 * j + k] = grid[i].mean[k]; }
        local2global[j++] = i;
      }
    // cov = [N -N][N -N]^T / (2m - 1)   (:43)
    double c6[6];
    {
      static const int A[6] = {0, 0, 0, 1, 1, 2}, B[6] = {0, 1, 2, 1, 2, 2};
      for (int e = 0; e < 6; ++e) {
        double s = 0;
        for (int q = 0; q < m; ++q) s += N[3 * q + A[e]] * N[3 * q + B[e]];
        c6[e] = (2.0 * s) / (double)(2 * m - 1);
      }
    }
    double S[3], V[3][3];
    eig3_sym(c6, S, V);
    const double score = S[2] / S[0];
    if (score < kCylScoreMin) return;          // Checkpoint 1 (:53)
    const double vec[3] = {V[0][0], V[1][0], V[2][0]};
    axis[0] = vec[0]; axis[1] = vec[1]; axis[2] = vec[2];
    // projection onto the plane orthogonal to the axis, normalised normals (:59-79)
    std::vector<double> Pp(3 * m);
    for (int q = 0; q < m; ++q) {
      const double pd = dot3(vec, &P[3 * q]);
      for (int k = 0; k < 3; ++k) Pp[3 * q + k] = P[3 * q + k] - pd * vec[k];
      const double nd = dot3(vec, &N[3 * q]);
      for (int k = 0; k < 3; ++k) N[3 * q + k] = N[3 * q + k] - nd * vec[k];
      const double nn = std::sqrt((N[3 * q] * N[3 * q] + N[3 * q + 1] * N[3 * q + 1]) + N[3 * q + 2] * N[3 * q + 2]);
      for (int k = 0; k < 3; ++k) N[3 * q + k] = N[3 * q + k] / nn;
    }
    // float K = log(1-p_success)/log(1-pow(w,3)) = 43.97.. ; later with w = 0.5: 12.05.. (:82-84,164)
    float K = (float)((double)std::log(1.0f - 0.8f) / std::log(1.0 - std::pow((double)0.33f, 3)));
    int m_left = m;
    std::vector<int> ids_left(m);
    std::vector<uint8_t> left_mask(m, 1);
    for (int i = 0; i < m; ++i) ids_left[i] = i;
    std::vector<double> D(m);
    std::vector<uint8_t> I(m), I_final(m, 0);
    while (m_left > 5 && m_left > 0.1 * m) {    // sequential RANSAC (:94)
      double min_hyp = kCylSqrMaxDist * m_left;
      const int accepted = (int)(0.9 * m_left);
      int max_inl = 0;
      std::fill(I_final.begin(), I_final.end(), 0);   // (reference: uninitialised; only read after a hypothesis wrote it)
      for (int k = 0; k < K; ++k) {
        const int id1 = ids_left[rng.next() % m_left];
        const int id2 = ids_left[rng.next() % m_left];
        const int id3 = ids_left[rng.next() % m_left];
        const double *n1 = &N[3 * id1], *n2 = &N[3 * id2], *n3 = &N[3 * id3];
        const double *p1 = &Pp[3 * id1], *p2 = &Pp[3 * id2], *p3 = &Pp[3 * id3];
        double e1[3], e2[3], t[3];
        for (int c = 0; c < 3; ++c) {
          e1[c] = (n1[c] + n2[c]) + n3[c];
          e2[c] = (p1[c] + p2[c]) + p3[c];
          t[c] = (n1[c] * p1[c] + n2[c] * p2[c]) + n3[c] * p3[c];
        }
        const double a = 1 - dot3(e1, e1) / 9;
        const double b = ((t[0] + t[1]) + t[2]) / 3 - dot3(e1, e2) / 9;
        double r = b / a;
        double center[3];
        for (int c = 0; c < 3; ++c) center[c] = (e2[c] - r * e1[c]) / 3;
        const double rr = r * r;
        for (int i = 0; i < m; ++i) {
          const double x = (Pp[3 * i] - r * N[3 * i]) - center[0], y = (Pp[3 * i + 1] - r * N[3 * i + 1]) - center[1],
                       z = (Pp[3 * i + 2] - r * N[3 * i + 2]) - center[2];
          D[i] = ((x * x + y * y) + z * z) / rr;
          I[i] = D[i] < kCylSqrMaxDist;
        }
        double dist = 0;                          // MSAC truncated distance (:137-148)
        int inl = 0;
        for (int i = 0; i < m; ++i)
          if (left_mask[i]) {
            if (I[i]) { ++inl; dist += D[i]; }
            else dist += kCylSqrMaxDist;
          }
        if (dist < min_hyp) {
          min_hyp = dist;
          max_inl = inl;
          for (int i = 0; i < m; ++i) I_final[i] = left_mask[i] ? I[i] : 0;
          if (inl > accepted) break;
        }
      }
      if (max_inl < 6) break;                     // Checkpoint 2 (:160)
      K = (float)((double)std::log(1.0f - 0.8f) / std::log(1.0 - std::pow(0.5, 3)));
      ids_left.clear();
      for (int i = 0; i < m; ++i) {
        if (I_final[i]) { left_mask[i] = 0; --m_left; }
        else if (left_mask[i]) ids_left.push_back(i);
      }
      // LLS over all inliers (:178-199)
      double e1[3] = {0, 0, 0}, e2[3] = {0, 0, 0}, b = 0;
      for (int i = 0; i < m; ++i)
        if (I_final[i]) {
          for (int c = 0; c < 3; ++c) { e1[c] += N[3 * i + c]; e2[c] += Pp[3 * i + c]; }
          b += (N[3 * i] * Pp[3 * i] + N[3 * i + 1] * Pp[3 * i + 1]) + N[3 * i + 2] * Pp[3 * i + 2];
        }
      const double n2 = (double)(max_inl * max_inl);
      const double a = 1 - dot3(e1, e1) / n2;
      b /= max_inl;
      b -= dot3(e1, e2) / n2;
      double r = b / a;
      std::array<double, 3> center;
      for (int c = 0; c < 3; ++c) center[c] = (e2[c] - r * e1[c]) / max_inl;
      if (r < 0) r = -r;
      ++nr_segments;
      radii.push_back((float)r);
      centers.push_back(center);
      inliers.push_back(I_final);
      // point-to-axis distances of the inliers' means (:206-226); fixed-size Vector3d reductions
      double P2d[3], dir[3];
      for (int c = 0; c < 3; ++c) { P2d[c] = center[c] + vec[c]; dir[c] = P2d[c] - center[c]; }
      const double P1P2d = std::sqrt(dir[0] * dir[0] + (dir[1] * dir[1] + dir[2] * dir[2]));
      double mse = 0;
      for (int i = 0; i < m; ++i)
        if (I_final[i]) {
          const double q[3] = {P[3 * i] - P2d[0], P[3 * i + 1] - P2d[1], P[3 * i + 2] - P2d[2]};
          const double cx = dir[1] * q[2] - dir[2] * q[1], cy = dir[2] * q[0] - dir[0] * q[2], cz = dir[0] * q[1] - dir[1] * q[0];
          const double dd = std::sqrt(cx * cx + (cy * cy + cz * cz)) / P1P2d - r;
          mse += dd * dd;
        }
      mse = mse / max_inl;
      MSEs.push_back(mse);
      P1.push_back({(float)center[0], (float)center[1], (float)center[2]});
      P2.push_back({(float)P2d[0], (float)P2d[1], (float)P2d[2]});
      P1P2_norm.push_back((float)P1P2d);
      cylindrical.push_back(1);
    }
  }
};

struct CapeOracle {
  int H, W, cw, ch, cyl;
  float max_merge_dist, min_cos;
  int ncx, ncy, ncells, npc;
  std::vector<Seg> grid;
  std::vector<int32_t> plane_map, cyl_map;
  std::vector<uint8_t> eroded_map, cyl_eroded_map;
  int nr_cylinders_found = 0;   // size of cylinder_segments_final (CAPE.cpp:434-445)

  CapeOracle(int h, int w, int cw_, int ch_, int cyl_, float mc, float mmd)
      : H(h), W(w), cw(cw_), ch(ch_), cyl(cyl_), max_merge_dist(mmd), min_cos(mc) {
    ncx = W / cw; ncy = H / ch; ncells = ncx * ncy; npc = cw * ch;
  }

  // RegionGrowing (CAPE.cpp:485-506), literal recursion.
  void grow(const std::vector<uint8_t>& input, std::vector<uint8_t>& output,
            const std::vector<float>& tols, int x, int y, const double* n1, double d1) {
    const int idx = x + ncx * y;
    if (!input[idx] || output[idx]) return;
    const double* n2 = grid[idx].normal;
    const double* m = grid[idx].mean;
    const double d2 = grid[idx].d;
    const double dist = n1[0] * m[0] + n1[1] * m[1] + n1[2] * m[2] + d1;
    if (n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2] < (double)min_cos ||
        dist * dist > (double)tols[idx])
      return;
    output[idx] = 1;
    if (x > 0) grow(input, output, tols, x - 1, y, n2, d2);
    if (x < ncx - 1) grow(input, output, tols, x + 1, y, n2, d2);
    if (y > 0) grow(input, output, tols, x, y - 1, n2, d2);
    if (y < ncy - 1) grow(input, output, tols, x, y + 1, n2, d2);
  }

  // 3x3 morphology with OpenCV's default border (outside pixels never win).
  void morph(const std::vector<uint8_t>& in, std::vector<uint8_t>& out, bool cross, bool erode) {
    out.assign(ncells, 0);
    for (int r = 0; r < ncy; ++r)
      for (int c = 0; c < ncx; ++c) {
        int acc = erode ? 255 : 0;
        for (int dr = -1; dr <= 1; ++dr)
          for (int dc = -1; dc <= 1; ++dc) {
            if (cross && dr != 0 && dc != 0) continue;
            int rr = r + dr, cc = c + dc;
            if (rr < 0 || rr >= ncy || cc < 0 || cc >= ncx) continue;
            int v = in[rr * ncx + cc];
            acc = erode ? std::min(acc, v) : std::max(acc, v);
          }
        out[r * ncx + c] = (uint8_t)acc;
      }
  }

  int process(const float* cloud, uint8_t* seg_out, drfe_plane* planes_out, int plane_cap,
              int* nr_planes_final, drfe_cylinder* cyls_out, int cyl_cap, int* nr_cylinders_final) {
    const size_t N = (size_t)H * W;
    const float* CX = cloud;
    const float* CY = cloud + N;
    const float* CZ = cloud + 2 * N;
    plane_map.assign(ncells, 0);
    eroded_map.assign(ncells, 0);
    cyl_map.assign(ncells, 0);
    cyl_eroded_map.assign(ncells, 0);
    const int cylinder_code_offset = 50;      // CAPE.cpp:53
    GlibcRand rng(1);                          // declared: restarted per frame (B.9)
    std::vector<CylSeg> cylinder_segments;
    std::vector<std::pair<int, int>> cylinder2region;
    int nr_cylinders = 0;
    std::vector<uint8_t> seg_stack(N, 0);
    std::vector<float> dist_stack(N);
    memset(dist_stack.data(), 100, N * sizeof(float));  // CAPE.cpp:60 (0x64646464)

    // ---- planar cell fitting (CAPE.cpp:62-80)
    grid.assign(ncells, Seg());
    std::vector<float> tols(ncells, 0.f);
    const float sin_merge = (float)std::sqrt(1 - (double)min_cos * (double)min_cos);
    for (int id = 0; id < ncells; ++id) {
      const size_t o = (size_t)id * npc;
      grid[id] = fit_cell(CX + o, CY + o, CZ + o, npc, cw);
      if (grid[id].planar) {
        const float dx = CX[o + npc - 1] - CX[o], dy = CY[o + npc - 1] - CY[o],
                    dz = CZ[o + npc - 1] - CZ[o];
        const float diam = std::sqrt(dx * dx + dy * dy + dz * dz);
        const float t = std::min(std::max(diam * sin_merge, 20.0f), max_merge_dist);
        tols[id] = t * t;
      }
    }
    // ---- histogram of normals in spherical coordinates (CAPE.cpp:82-101, Histogram.cpp)
    const int nb = 20;
    std::vector<int> Hh(nb * nb, 0), B(ncells, -1);
    std::vector<uint8_t> unassigned(ncells, 0);
    int remaining = 0;
    for (int id = 0; id < ncells; ++id) {
      if (!grid[id].planar) continue;
      const double nx = grid[id].normal[0], ny = grid[id].normal[1], nz = grid[id].normal[2];
      const double pn = std::sqrt(nx * nx + ny * ny);
      const double polar = std::acos(-nz);
      const double azim = std::atan2(nx / pn, ny / pn);
      int xq = (int)((nb - 1) * (polar - 0.0) / (3.14 - 0.0));
      int yq = 0;
      if (xq > 0) yq = (int)((nb - 1) * (azim - (-3.14)) / (3.14 - (-3.14)));
      const int bin = yq * nb + xq;
      B[id] = bin;
      Hh[bin]++;
      unassigned[id] = 1;
      ++remaining;
    }
    // ---- region growing with embedded model fitting (CAPE.cpp:114-218)
    std::vector<Seg> segs;
    std::vector<uint8_t> activation(ncells);
    while (remaining > 0) {
      int best_bin = -1, best_cnt = 0;
      for (int b = 0; b < nb * nb; ++b)
        if (Hh[b] > best_cnt) { best_bin = b; best_cnt = Hh[b]; }
      std::vector<int> cand;
      if (best_cnt > 0)
        for (int id = 0; id < ncells; ++id)
          if (B[id] == best_bin) cand.push_back(id);
      if (cand.size() < 5) break;
      int seed = cand[0];  // (reference: uninitialised if no candidate passes; cannot happen for finite MSE)
      float min_mse = (float)INT_MAX;
      for (size_t i = 0; i < cand.size(); ++i)
        if (grid[cand[i]].MSE < min_mse) {
          seed = cand[i];
          min_mse = grid[i].MSE;  // sic: loop counter, not candidate id (CAPE.cpp:130)
        }
      Seg acc = grid[seed];
      std::fill(activation.begin(), activation.end(), 0);
      grow(unassigned, activation, tols, seed % ncx, seed / ncx, acc.normal, acc.d);
      int activated = 0;
      for (int id = 0; id < ncells; ++id)
        if (activation[id]) {
          acc.expand(grid[id]);  // the seed is counted twice (CAPE.cpp:134 + :148)
          ++activated;
          Hh[B[id]]--; B[id] = -1;
          unassigned[id] = 0;
          --remaining;
        }
      if (activated < 4) continue;
      acc.fit();
      if (acc.score > 100) {
        segs.push_back(acc);
        const int label = (int)segs.size();
        for (int id = 0; id < ncells; ++id)
          if (activation[id]) plane_map[id] = label;
      }
      else if (cyl && activated > 5) {
        // it is an extrusion (CAPE.cpp:179-216)
        cylinder_segments.emplace_back(grid, activation, activated, rng);
        CylSeg& cy = cylinder_segments.back();
        for (int sid = 0; sid < cy.nr_segments; ++sid) {
          acc.x_acc = acc.y_acc = acc.z_acc = acc.xx_acc = acc.yy_acc = acc.zz_acc = acc.xy_acc = acc.xz_acc = acc.yz_acc = 0;
          acc.nr_pts = 0;                      // clearPoints
          for (int c = 0; c < activated; ++c)
            if (cy.inliers[sid][c]) acc.expand(grid[cy.local2global[c]]);
          acc.fit();
          if ((double)acc.MSE < cy.MSEs[sid]) {  // model selection (:195)
            segs.push_back(acc);
            for (int c = 0; c < activated; ++c)
              if (cy.inliers[sid][c]) plane_map[cy.local2global[c]] = (int)segs.size();
            cy.cylindrical[sid] = 0;
          } else {
            ++nr_cylinders;
            cylinder2region.push_back({(int)cylinder_segments.size() - 1, sid});
            for (int c = 0; c < activated; ++c)
              if (cy.inliers[sid][c]) cyl_map[cy.local2global[c]] = nr_cylinders;
            cy.cylindrical[sid] = 1;
          }
        }
      }
    }
    // ---- plane merging (CAPE.cpp:220-252, getConnectedComponents :459-481)
    const int np = (int)segs.size();
    std::vector<uint8_t> assoc((size_t)np * np, 0);
    for (int r = 0; r < ncy - 1; ++r)
      for (int c = 0; c < ncx - 1; ++c) {
        const int v = plane_map[r * ncx + c];
        if (v <= 0) continue;
        const int right = plane_map[r * ncx + c + 1], below = plane_map[(r + 1) * ncx + c];
        if (right > 0 && v != right) assoc[(size_t)(v - 1) * np + right - 1] = 1;
        if (below > 0 && v != below) assoc[(size_t)(v - 1) * np + below - 1] = 1;
      }
    for (int r = 0; r < np; ++r)
      for (int c = r + 1; c < np; ++c)
        assoc[(size_t)r * np + c] = assoc[(size_t)r * np + c] || assoc[(size_t)c * np + r];
    std::vector<int> merge(np);
    for (int i = 0; i < np; ++i) merge[i] = i;
    for (int r = 0; r < np; ++r) {
      const int pid = merge[r];
      bool expanded = false;
      for (int c = r + 1; c < np; ++c) {
        if (!assoc[(size_t)r * np + c]) continue;
        const double cosang = segs[pid].normal[0] * segs[c].normal[0] +
                              segs[pid].normal[1] * segs[c].normal[1] +
                              segs[pid].normal[2] * segs[c].normal[2];
        // sic: first term uses plane r, the rest plane_id (CAPE.cpp:238-240)
        const double dd = segs[r].normal[0] * segs[c].mean[0] + segs[pid].normal[1] * segs[c].mean[1] +
                          segs[pid].normal[2] * segs[c].mean[2] + segs[pid].d;
        const double dist = dd * dd;
        if (cosang > (double)min_cos && dist < (double)max_merge_dist) {
          segs[pid].expand(segs[c]);
          merge[c] = pid;
          expanded = true;
        } else {
          assoc[(size_t)r * np + c] = 0;
        }
      }
      if (expanded) segs[pid].fit();
    }
    // ---- boundary refinement (CAPE.cpp:254-321)
    std::vector<uint8_t> mask(ncells), er, di;
    int nfinal = 0;
    for (int i = 0; i < np; ++i) {
      if (i != merge[i]) continue;
      std::fill(mask.begin(), mask.end(), 0);
      for (int j = i; j < np; ++j)
        if (merge[j] == merge[i])
          for (int id = 0; id < ncells; ++id)
            if (plane_map[id] == j + 1) mask[id] = 1;
      morph(mask, er, /*cross=*/true, /*erode=*/true);
      if (*std::max_element(er.begin(), er.end()) == 0) continue;
      if (nfinal < plane_cap) planes_out[nfinal] = segs[i];
      ++nfinal;
      morph(mask, di, /*cross=*/false, /*erode=*/false);
      const uint8_t plane_nr = (uint8_t)nfinal;
      const float nx = (float)segs[i].normal[0], ny = (float)segs[i].normal[1],
                  nz = (float)segs[i].normal[2], d = (float)segs[i].d;
      for (int id = 0; id < ncells; ++id)
        if (er[id] > 0) eroded_map[id] = plane_nr;
      for (int id = 0; id < ncells; ++id) {
        if ((uint8_t)(di[id] - er[id]) == 0) continue;
        const float max_dist = 9 * segs[i].MSE;
        const size_t o = (size_t)id * npc;
        for (int j = 0; j < npc; ++j) {
          const float v = CX[o + j] * nx + CY[o + j] * ny + CZ[o + j] * nz + d;
          const float dist = v * v;
          if (dist < max_dist && dist < dist_stack[o + j]) {
            dist_stack[o + j] = dist;
            seg_stack[o + j] = plane_nr;
          }
        }
      }
    }
    *nr_planes_final = nfinal;
    // ---- cylinder boundary refinement (CAPE.cpp:323-393)
    int ncyl_final = 0;
    if (cyl) {
      for (int i = 0; i < nr_cylinders; ++i) {
        const CylSeg& cy = cylinder_segments[cylinder2region[i].first];
        const int sid = cylinder2region[i].second;
        for (int id = 0; id < ncells; ++id) mask[id] = cyl_map[id] == i + 1;
        morph(mask, er, true, true);
        if (*std::max_element(er.begin(), er.end()) == 0) continue;
        ++ncyl_final;
        morph(mask, di, false, false);
        const uint8_t label = (uint8_t)(cylinder_code_offset + ncyl_final);
        for (int id = 0; id < ncells; ++id)
          if (er[id] > 0) cyl_eroded_map[id] = label;
        const float P2[3] = {cy.P2[sid][0], cy.P2[sid][1], cy.P2[sid][2]};
        const float P1P2[3] = {P2[0] - cy.P1[sid][0], P2[1] - cy.P1[sid][1], P2[2] - cy.P1[sid][2]};
        const double norm12 = cy.P1P2_norm[sid], radius = cy.radii[sid];
        const float max_dist = (float)(9 * cy.MSEs[sid]);
        for (int id = 0; id < ncells; ++id) {
          if ((uint8_t)(di[id] - er[id]) == 0) continue;
          const size_t o = (size_t)id * npc;
          for (int j = 0; j < npc; ++j) {
            const float z = CZ[o + j];
            if (!(z > 0)) continue;
            const float q[3] = {CX[o + j] - P2[0], CY[o + j] - P2[1], z - P2[2]};
            const float c0 = P1P2[1] * q[2] - P1P2[2] * q[1], c1 = P1P2[2] * q[0] - P1P2[0] * q[2],
                        c2 = P1P2[0] * q[1] - P1P2[1] * q[0];
            const float nrm = std::sqrt(c0 * c0 + (c1 * c1 + c2 * c2));   // fixed-size Vector3f redux
            float dist = (float)((double)nrm / norm12 - radius);
            dist *= dist;
            if (dist < max_dist && dist < dist_stack[o + j]) {
              dist_stack[o + j] = dist;
              seg_stack[o + j] = label;
            }
          }
        }
      }
    }
    if (nr_cylinders_final) *nr_cylinders_final = ncyl_final;
    nr_cylinders_found = nr_cylinders;
    for (int i = 0; i < nr_cylinders && i < cyl_cap && cyls_out; ++i) {   // CAPE.cpp:434-445
      const CylSeg& cy = cylinder_segments[cylinder2region[i].first];
      const int sid = cylinder2region[i].second;
      cyls_out[i].radius = cy.radii[sid];
      for (int c = 0; c < 3; ++c) { cyls_out[i].center[c] = cy.centers[sid][c]; cyls_out[i].axis[c] = cy.axis[c]; }
    }
    // ---- write seg_output in image layout (CAPE.cpp:395-432)
    for (int cr = 0; cr < ncy; ++cr)
      for (int cc = 0; cc < ncx; ++cc) {
        const int id = cr * ncx + cc;
        const uint8_t* st = &seg_stack[(size_t)id * npc];
        for (int r = 0; r < ch; ++r) {
          uint8_t* row = seg_out + (size_t)(cr * ch + r) * W + cc * cw;
          for (int c = 0; c < cw; ++c) {
            if (eroded_map[id] > 0) row[c] = eroded_map[id];
            else if (cyl_eroded_map[id] > 0) row[c] = cyl_eroded_map[id];
            else if (st[r * cw + c] > 0) row[c] = st[r * cw + c];
          }
        }
      }
    return 0;
  }
};

}  // namespace

extern "C" {

void* orc_cape_create(int depth_height, int depth_width, int cell_width, int cell_height,
                      int cylinder_detection, float min_cos_angle_4_merge, float max_merge_dist) {
  return new CapeOracle(depth_height, depth_width, cell_width, cell_height, cylinder_detection,
                        min_cos_angle_4_merge, max_merge_dist);
}
void orc_cape_destroy(void* h) { delete (CapeOracle*)h; }

void orc_cape_depth_to_cloud(void* h, const float* depth, int row_stride, float fx, float fy,
                             float cx, float cy, float* cloud) {
  CapeOracle* o = (CapeOracle*)h;
  const size_t N = (size_t)o->H * o->W;
  for (int i = 0; i < o->H; ++i)
    for (int j = 0; j < o->W; ++j) {
      const double z = (double)depth[(size_t)i * row_stride + j];
      const double x = ((double)j - (double)cx) * z / (double)fx;
      const double y = ((double)i - (double)cy) * z / (double)fy;
      // cell_map (PlaneExtractor.cpp:135-148); pixels beyond the last full cell are dropped
      const int cr = i / o->ch, lr = i % o->ch, cc = j / o->cw, lc = j % o->cw;
      if (cr >= o->ncy || cc >= o->ncx) continue;
      const size_t id = (size_t)(cr * o->ncx + cc) * o->npc + lr * o->cw + lc;
      cloud[id] = (float)x;
      cloud[N + id] = (float)y;
      cloud[2 * N + id] = (float)z;
    }
}

int orc_cape_process(void* h, const float* cloud, uint8_t* seg_out, drfe_plane* planes,
                     int plane_cap, int* nr_planes, drfe_cylinder* cyls, int cyl_cap,
                     int* nr_cylinders) {
  return ((CapeOracle*)h)->process(cloud, seg_out, planes, plane_cap, nr_planes, cyls, cyl_cap, nr_cylinders);
}
int orc_cape_cylinders_found(void* h) { return ((CapeOracle*)h)->nr_cylinders_found; }
int orc_cape_get_cyl_maps(void* h, int32_t* cyl_map, uint8_t* cyl_eroded_map) {
  CapeOracle* o = (CapeOracle*)h;
  memcpy(cyl_map, o->cyl_map.data(), o->ncells * sizeof(int32_t));
  memcpy(cyl_eroded_map, o->cyl_eroded_map.data(), o->ncells);
  return o->ncells;
}
// ---- pcl::VoxelGrid<PointT>::applyFilter on one plane_cloud (reference src/Frame.cc:1121-1125: leaf 0.05, default
// downsample_all_data, min_points_per_voxel 0).  PCL is a third-party dependency that is not vendored in the reference
// (CMakeLists.txt:59 asks for PCL 1.9); this restates the published algorithm of pcl/filters/impl/voxel_grid.hpp (1.9):
//   getMinMax3D -> min_b / max_b = floor(min|max * inverse_leaf) per axis, div_b = max_b - min_b + 1;
//   if the box has more than INT32_MAX leaves the input is returned unfiltered ("Leaf size is too small");
//   leaf index of a point = sum_k (int)(floor(p_k * inverse_leaf_k) - (float)min_b_k) * divb_mul_k;
//   points sorted by leaf index; per leaf the centroid = float sum of its points / (float)count (AccumulatorXYZ),
//   leaves in ascending index order.
// Declared: PCL sorts with std::sort, which leaves the order of the points INSIDE a leaf (and so the last bits of the
// float sum) to the library's introsort; here a leaf's points are summed in ascending input index (a stable sort).
// Returns the number of output points (n when unfiltered, *unfiltered = 1).
int orc_voxel_grid(const float* xyz, int n, float leaf, float* out, int* unfiltered) {
  *unfiltered = 0;
  if (n <= 0) return 0;
  const float inv = 1.0f / leaf;                                  // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], xyz[3 * i + k]); mx[k] = std::max(mx[k], xyz[3 * i + k]); }
  long long d[3];
  for (int k = 0; k < 3; ++k) d[k] = (long long)((mx[k] - mn[k]) * inv) + 1;
  if (d[0] * d[1] * d[2] > (long long)INT32_MAX) {
    std::memcpy(out, xyz, (size_t)n * 3 * sizeof(float));
    *unfiltered = 1;
    return n;
  }
  int min_b[3], div_b[3];
  for (int k = 0; k < 3; ++k) {
    min_b[k] = (int)std::floor(mn[k] * inv);
    div_b[k] = (int)std::floor(mx[k] * inv) - min_b[k] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<std::pair<unsigned, int>> iv((size_t)n);
  for (int i = 0; i < n; ++i) {
    int idx = 0;
    for (int k = 0; k < 3; ++k) idx += (int)(std::floor(xyz[3 * i + k] * inv) - (float)min_b[k]) * mul[k];
    iv[i] = std::make_pair((unsigned)idx, i);
  }
  std::stable_sort(iv.begin(), iv.end(), [](const std::pair<unsigned, int>& a, const std::pair<unsigned, int>& b) { return a.first < b.first; });
  int total = 0;
  for (size_t a = 0; a < iv.size();) {
    size_t b = a + 1;
    while (b < iv.size() && iv[b].first == iv[a].first) ++b;
    float sx = 0.f, sy = 0.f, sz = 0.f;                            // AccumulatorXYZ: xyz starts at zero, += every point
    for (size_t i = a; i < b; ++i) { sx += xyz[3 * iv[i].second]; sy += xyz[3 * iv[i].second + 1]; sz += xyz[3 * iv[i].second + 2]; }
    const float cnt = (float)(b - a);
    out[3 * total] = sx / cnt; out[3 * total + 1] = sy / cnt; out[3 * total + 2] = sz / cnt;
    ++total;
    a = b;
  }
  return total;
}

// ---- the 1/3-resolution cloud Frame::ComputePlanes_CAPE builds for the surface normals (reference src/Frame.cc:1153-1172):
// every third pixel of every third row, p.z = d > max_point_dist ? 0 : d, p.x = (n - cx) * p.z / fx in FLOAT arithmetic
// (int n converted to float; cx, fx are the Frame's float members).  out: ceil(H/3) x ceil(W/3) x 3.
void orc_third_cloud(const float* depth, int width, int height, int row_stride, float fx, float fy, float cx, float cy, float max_point_dist,
                     float* out) {
  size_t o = 0;
  for (int m = 0; m < height; m += 3)
    for (int n = 0; n < width; n += 3) {
      const float dd = depth[(size_t)m * row_stride + n];
      const float z = dd > max_point_dist ? 0.f : dd;
      out[o++] = ((float)n - cx) * z / fx;
      out[o++] = ((float)m - cy) * z / fy;
      out[o++] = z;
    }
}

int orc_glibc_rand(uint32_t seed, int n, int32_t* out) {
  GlibcRand g(seed);
  for (int i = 0; i < n; ++i) out[i] = g.next();
  return n;
}
int orc_cape_get_cells(void* h, drfe_plane* cells) {
  CapeOracle* o = (CapeOracle*)h;
  for (int i = 0; i < o->ncells; ++i) cells[i] = o->grid[i];
  return o->ncells;
}
int orc_cape_get_grid_maps(void* h, int32_t* plane_map, uint8_t* eroded_map) {
  CapeOracle* o = (CapeOracle*)h;
  memcpy(plane_map, o->plane_map.data(), o->ncells * sizeof(int32_t));
  memcpy(eroded_map, o->eroded_map.data(), o->ncells);
  return o->ncells;
}

}  // extern "C"
