// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline legs may load this library.
//
// CPU restatement of the plane extractor that is live in DR-SLAM's Frame constructor (Frame.cc:126, :937-949):
//   PlaneDetection::readDepthImage / runPlaneDetection      src/PlaneExtractor.cpp:28-63
//   ahc::PlaneFitter<ImagePointCloud>::run                  include/peac/AHCPlaneFitter.hpp:211-259
//     initGraph :804-965, ahCluster :976-1190, refineDetails :298-382, findBlockMembership :494-600, floodFill :434-488
//   ahc::PlaneSeg (+ Stats)                                  include/peac/AHCPlaneSeg.hpp:57-409
//   ahc::ParamSet                                            include/peac/AHCParamSet.hpp:46-147
//   DisjointSet                                              include/peac/DisjointSet.hpp:31-95
// Dependency-free C++17, every loop in the reference's order.  DECLARED where the reference's own result is left to a
// library, an allocator or the compiler:
//   P.1  LA::eig33sym = Eigen::SelfAdjointEigenSolver<Matrix3d> (eig33sym.hpp:63-68) is replaced by the cyclic Jacobi solver
//        of the CAPE oracle (same eigenpairs to ~1e-15; the GPU runs the same operation sequence, so every mse comparison of
//        the clustering comes out the same).
//   P.2  std::priority_queue<shared_ptr<PlaneSeg>, ..., PlaneSegMinMSECmp> compares mse only; among equal mse the pop order
//        is the heap's.  Declared: smaller creation sequence number first.  Stale entries (nouse) are skipped like :1005-1008.
//   P.3  PlaneSeg::nbs is a std::set<PlaneSeg*>, iterated in ADDRESS order when the merge candidates are tried (:1030-1051).
//        Declared: ascending creation sequence number.  (It matters only when two candidates give exactly the same mse.)
//   P.4  std::sort of extractedPlanes by N (:1186-1188) is not stable; declared: stable (libstdc++'s std::sort is an
//        insertion sort, hence stable, up to 16 elements anyway).
//   P.5  no FMA contraction in Stats::push / compute (the reference's -O3 -march=native build may contract).
// PARITY STATUS: "parity unpinned" by the reference (no tests / fixtures; it cannot be built here: the headers include OpenCV).
// Pinned by: oracle/peac_py.py (an independent Python restatement: equal value for value, and decision for decision with LAPACK as the
// solver), numpy.linalg.eigh for the plane fit, tests/golden/peac_640x480.npz against drift (tests/test_peac.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <queue>
#include <set>
#include <vector>

#include "drfe_oracle.h"

namespace {

// cyclic Jacobi (cape_oracle.cpp's eig3_sym, except that a pivot too small to change the diagonal is zeroed from the fourth sweep on
// instead of the fifth: three rotations less per fit, the same eigenpairs; tests/test_peac.py pins them to LAPACK): in a = {xx,xy,xz,yy,yz,zz}; w ascending, v[k][i] = component k of evec i
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
      if (sweep > 2 && std::fabs(app) + g == std::fabs(app) && std::fabs(aqq) + g == std::fabs(aqq)) {
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
  int i0 = 0, i1 = 1, i2 = 2;
  const double d[3] = {a[0][0], a[1][1], a[2][2]};
  if (d[i1] < d[i0]) std::swap(i0, i1);
  if (d[i2] < d[i1]) std::swap(i1, i2);
  if (d[i1] < d[i0]) std::swap(i0, i1);
  double vv[3][3];
  const int idx[3] = {i0, i1, i2};
  for (int i = 0; i < 3; ++i) {
    w[i] = d[idx[i]];
    for (int m = 0; m < 3; ++m) vv[m][i] = v[m][idx[i]];
  }
  for (int i = 0; i < 3; ++i)
    for (int m = 0; m < 3; ++m) v[m][i] = vv[m][i];
}

struct Params {   // ahc::ParamSet, AHCParamSet.hpp:46-78, + the PlaneFitter members of AHCPlaneFitter.hpp:110-118
  double depthSigma, stdTol_init, stdTol_merge, z_near, z_far, angle_near, angle_far, similarityTh_merge, similarityTh_refine, depthAlpha,
      depthChangeTol;
  int minSupport, windowWidth, windowHeight;
  double T_mse_init(double z) const { const double t = depthSigma * z * z + stdTol_init; return t * t; }      // std::pow(., 2)
  double T_mse_merge(double z) const { const double t = depthSigma * z * z + stdTol_merge; return t * t; }
  double T_ang_init(double z) const {
    double cz = z;
    cz = std::max(cz, z_near);
    cz = std::min(cz, z_far);
    const double factor = (angle_far - angle_near) / (z_far - z_near);
    return std::cos(factor * cz + angle_near - factor * z_near);
  }
  double T_dz(double z) const { return depthAlpha * std::fabs(z) + depthChangeTol; }
};

struct Stats {
  double sx = 0, sy = 0, sz = 0, sxx = 0, syy = 0, szz = 0, sxy = 0, syz = 0, sxz = 0;
  int N = 0;
  void push(double x, double y, double z) {
    sx += x; sy += y; sz += z;
    sxx += x * x; syy += y * y; szz += z * z;
    sxy += x * y; syz += y * z; sxz += x * z;
    ++N;
  }
  void clear() { *this = Stats(); }
  static Stats merged(const Stats& a, const Stats& b) {
    Stats s;
    s.sx = a.sx + b.sx; s.sy = a.sy + b.sy; s.sz = a.sz + b.sz;
    s.sxx = a.sxx + b.sxx; s.syy = a.syy + b.syy; s.szz = a.szz + b.szz;
    s.sxy = a.sxy + b.sxy; s.syz = a.syz + b.syz; s.sxz = a.sxz + b.sxz;
    s.N = a.N + b.N;
    return s;
  }
  // AHCPlaneSeg.hpp:128-162
  void compute(double center[3], double normal[3], double& mse, double& curvature) const {
    const double sc = 1.0 / N;
    center[0] = sx * sc; center[1] = sy * sc; center[2] = sz * sc;
    const double K[6] = {sxx - sx * sx * sc, sxy - sx * sy * sc, sxz - sx * sz * sc, syy - sy * sy * sc, syz - sy * sz * sc, szz - sz * sz * sc};
    double sv[3], V[3][3];
    eig3_sym(K, sv, V);
    if (V[0][0] * center[0] + V[1][0] * center[1] + V[2][0] * center[2] <= 0) {
      normal[0] = V[0][0]; normal[1] = V[1][0]; normal[2] = V[2][0];
    } else {
      normal[0] = -V[0][0]; normal[1] = -V[1][0]; normal[2] = -V[2][0];
    }
    mse = sv[0] * sc;
    curvature = sv[0] / (sv[0] + sv[1] + sv[2]);
  }
};

struct Seg {
  Stats stats;
  int rid = 0, N = 0;
  double mse = 0, center[3] = {0, 0, 0}, normal[3] = {0, 0, 0}, curvature = 0;
  bool nouse = false;
  std::set<int> nbs;   // ids (creation sequence numbers), P.3
  double normalSimilarity(const Seg& p) const { return std::abs(normal[0] * p.normal[0] + normal[1] * p.normal[1] + normal[2] * p.normal[2]); }
  double signedDist(const double pt[3]) const {
    return normal[0] * (pt[0] - center[0]) + normal[1] * (pt[1] - center[1]) + normal[2] * (pt[2] - center[2]);
  }
};

struct DisjointSet {
  std::vector<int> parent, size;
  explicit DisjointSet(int n) : parent(n), size(n, 1) { for (int i = 0; i < n; ++i) parent[i] = i; }
  int Find(int x) { if (parent[x] != x) parent[x] = Find(parent[x]); return parent[x]; }
  int getSetSize(int x) { return size[Find(x)]; }
  int Union(int x, int y) {
    const int xr = Find(x), yr = Find(y);
    if (xr == yr) return xr;
    if (size[xr] < size[yr]) { parent[xr] = yr; size[yr] += size[xr]; return yr; }
    parent[yr] = xr; size[xr] += size[yr]; return xr;
  }
};

struct Fitter {
  Params prm;
  int width = 0, height = 0;
  const double* cloud = nullptr;   // [H*W][3] (ImagePointCloud::vertices)
  std::vector<Seg> segs;           // by creation sequence number
  DisjointSet* ds = nullptr;
  std::vector<int> extracted;      // extractedPlanes, ids
  std::vector<int> membership;     // membershipImg
  std::vector<int> blkMap;
  std::vector<std::pair<int, int>> rfQueue;

  bool get(int i, int j, double& x, double& y, double& z) const {   // ImagePointCloud::get, PlaneExtractor.h:50-58
    const double* p = cloud + 3 * ((size_t)i * width + j);
    z = p[2];
    if (z == 0 || std::isnan(z)) return false;
    x = p[0]; y = p[1];
    return true;
  }
  bool depthDisContinuous(double d0, double d1) const { return std::fabs(d0 - d1) > prm.T_dz(d0); }

  typedef std::pair<double, int> QE;   // (mse, id)
  struct QCmp { bool operator()(const QE& a, const QE& b) const { return b.first < a.first || (b.first == a.first && b.second < a.second); } };   // P.2
  typedef std::priority_queue<QE, std::vector<QE>, QCmp> MinQ;

  // PlaneSeg(points, root_block_id, seed_row, seed_col, ...), AHCPlaneSeg.hpp:213-290 (INIT_STRICT)
  Seg initSeg(int rid, int seed_row, int seed_col) {
    Seg s;
    s.rid = rid;
    bool windowValid = true;
    const int winH = prm.windowHeight, winW = prm.windowWidth;
    for (int i = seed_row, icnt = 0; icnt < winH && i < height; ++i, ++icnt) {
      for (int j = seed_col, jcnt = 0; jcnt < winW && j < width; ++j, ++jcnt) {
        double x = 0, y = 0, z = 10000;
        if (!get(i, j, x, y, z)) { windowValid = false; break; }
        double xn = 0, yn = 0, zn = 10000;
        if (j + 1 < width && (get(i, j + 1, xn, yn, zn) && depthDisContinuous(z, zn))) { windowValid = false; break; }
        if (i + 1 < height && (get(i + 1, j, xn, yn, zn) && depthDisContinuous(z, zn))) { windowValid = false; break; }
        s.stats.push(x, y, z);
      }
      if (!windowValid) break;
    }
    if (windowValid) { s.nouse = false; s.N = s.stats.N; }
    else { s.N = 0; s.stats.clear(); s.nouse = true; }
    if (s.N < 4) s.mse = s.curvature = std::numeric_limits<double>::quiet_NaN();
    else s.stats.compute(s.center, s.normal, s.mse, s.curvature);
    return s;
  }
  void connect(int a, int b) { segs[a].nbs.insert(b); segs[b].nbs.insert(a); }
  void disconnectAllNbs(int a) {
    for (int nb : segs[a].nbs) segs[nb].nbs.erase(a);
    segs[a].nbs.clear();
  }

  // AHCPlaneFitter.hpp:804-965
  void initGraph(MinQ& minQ) {
    const int Nh = height / prm.windowHeight, Nw = width / prm.windowWidth;
    std::vector<int> G(Nh * Nw, -1);
    segs.clear();
    segs.reserve(2 * Nh * Nw);
    for (int i = 0; i < Nh; ++i)
      for (int j = 0; j < Nw; ++j) {
        Seg p = initSeg(i * Nw + j, i * prm.windowHeight, j * prm.windowWidth);
        if (p.mse < prm.T_mse_init(p.center[2]) && !p.nouse) {
          segs.push_back(p);
          G[i * Nw + j] = (int)segs.size() - 1;
          minQ.push(QE(p.mse, (int)segs.size() - 1));
        }
      }
    auto sim = [&](int a, int b) { return segs[a].normalSimilarity(segs[b]); };
    for (int i = 0; i < Nh; ++i) {
      for (int j = 1; j < Nw; j += 2) {
        const int cidx = i * Nw + j;
        if (G[cidx - 1] < 0) { --j; continue; }
        if (G[cidx] < 0) continue;
        if (j < Nw - 1 && G[cidx + 1] < 0) { ++j; continue; }
        const double th = prm.T_ang_init(segs[G[cidx]].center[2]);
        if ((j < Nw - 1 && sim(G[cidx - 1], G[cidx + 1]) >= th) || (j == Nw - 1 && sim(G[cidx], G[cidx - 1]) >= th)) {
          connect(G[cidx], G[cidx - 1]);
          if (j < Nw - 1) connect(G[cidx], G[cidx + 1]);
        } else {
          --j;
        }
      }
    }
    for (int j = 0; j < Nw; ++j) {
      for (int i = 1; i < Nh; i += 2) {
        const int cidx = i * Nw + j;
        if (G[cidx - Nw] < 0) { --i; continue; }
        if (G[cidx] < 0) continue;
        if (i < Nh - 1 && G[cidx + Nw] < 0) { ++i; continue; }
        const double th = prm.T_ang_init(segs[G[cidx]].center[2]);
        if ((i < Nh - 1 && sim(G[cidx - Nw], G[cidx + Nw]) >= th) || (i == Nh - 1 && sim(G[cidx], G[cidx - Nw]) >= th)) {
          connect(G[cidx], G[cidx - Nw]);
          if (i < Nh - 1) connect(G[cidx], G[cidx + Nw]);
        } else {
          --i;
        }
      }
    }
  }

  // AHCPlaneFitter.hpp:976-1190 (maxStep = 100000 is never reached: at most one step per node)
  int ahCluster(MinQ& minQ) {
    int step = 0;
    while (!minQ.empty()) {
      const int p = minQ.top().second;
      minQ.pop();
      if (segs[p].nouse) continue;
      int cand_nb = -1;
      Seg cand_merge;
      bool have = false;
      for (int nb : segs[p].nbs) {                               // P.3: ascending creation sequence
        if (segs[p].normalSimilarity(segs[nb]) < prm.similarityTh_merge) continue;
        Seg merge;                                                // PlaneSeg(pa, pb), AHCPlaneSeg.hpp:298-320
        merge.stats = Stats::merged(segs[p].stats, segs[nb].stats);
        merge.nouse = false;
        merge.rid = segs[p].N >= segs[nb].N ? segs[p].rid : segs[nb].rid;
        merge.N = merge.stats.N;
        merge.stats.compute(merge.center, merge.normal, merge.mse, merge.curvature);
        if (!have || cand_merge.mse > merge.mse || (cand_merge.mse == merge.mse && cand_merge.N < merge.mse)) {
          cand_merge = merge; cand_nb = nb; have = true;
        }
      }
      if (have && cand_merge.mse < prm.T_mse_merge(cand_merge.center[2])) {
        segs.push_back(cand_merge);
        const int m = (int)segs.size() - 1;
        minQ.push(QE(segs[m].mse, m));
        // mergeNbsFrom(pa, pb, ds), AHCPlaneSeg.hpp:378-407
        ds->Union(segs[p].rid, segs[cand_nb].rid);
        segs[m].nbs.insert(segs[p].nbs.begin(), segs[p].nbs.end());
        segs[m].nbs.insert(segs[cand_nb].nbs.begin(), segs[cand_nb].nbs.end());
        segs[m].nbs.erase(p);
        segs[m].nbs.erase(cand_nb);
        disconnectAllNbs(p);
        disconnectAllNbs(cand_nb);
        for (int nb : segs[m].nbs) segs[nb].nbs.insert(m);
        segs[p].nouse = segs[cand_nb].nouse = true;
      } else {
        if (segs[p].N >= prm.minSupport) extracted.push_back(p);
        disconnectAllNbs(p);
      }
      ++step;
    }
    std::stable_sort(extracted.begin(), extracted.end(), [&](int a, int b) { return segs[b].N < segs[a].N; });   // P.4
    return step;
  }

  static int getValid4Neighbor(int i, int j, int H, int W, int nbs[4]) {
    const int id = i * W + j;
    int cnt = 0;
    if (j > 0) nbs[cnt++] = id - 1;
    if (j < W - 1) nbs[cnt++] = id + 1;
    if (i > 0) nbs[cnt++] = id - W;
    if (i < H - 1) nbs[cnt++] = id + W;
    return cnt;
  }
  int getBlockIdx(int pixX, int pixY) const {
    const int Nw = width / prm.windowWidth, Nh = height / prm.windowHeight;
    const int by = pixY / prm.windowHeight, bx = pixX / prm.windowWidth;
    return (by < Nh && bx < Nw) ? (by * Nw + bx) : -1;
  }

  // AHCPlaneFitter.hpp:494-600 (erodeType = ERODE_ALL_BORDER)
  void findBlockMembership(std::vector<char>& isValid) {
    std::map<int, int> rid2plid;
    for (int plid = 0; plid < (int)extracted.size(); ++plid) rid2plid.insert(std::make_pair(segs[extracted[plid]].rid, plid));
    const int Nh = height / prm.windowHeight, Nw = width / prm.windowWidth, winH = prm.windowHeight, winW = prm.windowWidth;
    const int NptsPerBlk = winH * winW;
    membership.assign((size_t)height * width, -1);
    blkMap.assign(Nh * Nw, 0);
    isValid.assign(extracted.size(), 0);
    for (int i = 0, blkid = 0; i < Nh; ++i) {
      for (int j = 0; j < Nw; ++j, ++blkid) {
        const int setid = ds->Find(blkid);
        const int setSize = ds->getSetSize(setid) * NptsPerBlk;
        if (setSize >= prm.minSupport) {
          int nbs[4] = {-1};
          const int nNbs = getValid4Neighbor(i, j, Nh, Nw, nbs);
          bool same = true;
          for (int k = 0; k < nNbs; ++k)
            if (ds->Find(nbs[k]) != setid) { same = false; break; }   // ERODE_ALL_BORDER
          const int plid = rid2plid[setid];
          if (same) {
            blkMap[blkid] = plid;
            for (int y = i * winH; y < (i + 1) * winH; ++y)
              for (int x = j * winW; x < (j + 1) * winW; ++x) membership[(size_t)y * width + x] = plid;
            isValid[plid] = 1;
          } else {
            blkMap[blkid] = -1;
          }
        } else {
          blkMap[blkid] = -1;
        }
        if (blkMap[blkid] < 0) {
          if (i > 0) {
            const int u = blkid - Nw;
            if (blkMap[u] >= 0) {
              const int spix = (i * winH - 1) * width + j * winW;
              for (int k = 1; k < winW; ++k) rfQueue.push_back(std::make_pair(spix + k, blkMap[u]));
            }
          }
          if (j > 0) {
            const int l = blkid - 1;
            if (blkMap[l] >= 0) {
              const int spix = (i * winH) * width + j * winW - 1;
              for (int k = 0; k < winH - 1; ++k) rfQueue.push_back(std::make_pair(spix + k * width, blkMap[l]));
            }
          }
        } else {
          const int plid = blkMap[blkid];
          if (i > 0) {
            const int u = blkid - Nw;
            if (blkMap[u] != plid) {
              const int spix = (i * winH) * width + j * winW;
              for (int k = 0; k < winW - 1; ++k) rfQueue.push_back(std::make_pair(spix + k, plid));
            }
          }
          if (j > 0) {
            const int l = blkid - 1;
            if (blkMap[l] != plid) {
              const int spix = (i * winH) * width + j * winW;
              for (int k = 1; k < winH; ++k) rfQueue.push_back(std::make_pair(spix + k * width, plid));
            }
          }
        }
      }
    }
  }

  // AHCPlaneFitter.hpp:434-488
  void floodFill() {
    std::vector<float> distMap((size_t)height * width, std::numeric_limits<float>::max());
    for (int k = 0; k < (int)rfQueue.size(); ++k) {
      const int sIdx = rfQueue[k].first;
      const int seedy = sIdx / width, seedx = sIdx - seedy * width;
      const int plid = rfQueue[k].second;
      const Seg& pl = segs[extracted[plid]];
      int nbs[4] = {-1};
      const int Nnbs = getValid4Neighbor(seedy, seedx, height, width, nbs);
      for (int itr = 0; itr < Nnbs; ++itr) {
        const int cIdx = nbs[itr];
        int& trail = membership[cIdx];
        if (trail <= -6) continue;
        if (trail >= 0 && trail == plid) continue;
        const int cy = cIdx / width, cx = cIdx - cy * width;
        const int blkid = getBlockIdx(cx, cy);
        if (blkid >= 0 && blkMap[blkid] >= 0) continue;
        double pt[3] = {0, 0, 0};
        float cdist = -1;
        bool ok = get(cy, cx, pt[0], pt[1], pt[2]);
        if (ok) {
          cdist = (float)std::abs(pl.signedDist(pt));
          ok = (double)cdist * (double)cdist < 9 * pl.mse + 1e-5;   // std::pow(float, 2) = the exact double square
        }
        if (ok) {
          if (trail >= 0) {
            Seg& n_pl = segs[extracted[trail]];
            if (pl.normalSimilarity(n_pl) >= prm.similarityTh_refine) connect(extracted[trail], extracted[plid]);
          }
          float& old_dist = distMap[cIdx];
          if (cdist < old_dist) {
            trail = plid;
            old_dist = cdist;
            rfQueue.push_back(std::make_pair(cIdx, plid));
          } else if (trail < 0) {
            trail -= 1;
          }
        } else {
          if (trail < 0) trail -= 1;
        }
      }
    }
  }
};

}  // namespace

extern "C" {

// Stats::compute on nine sums {sx, sy, sz, sxx, syy, szz, sxy, syz, sxz} and N -> out = center[3], normal[3], mse, curvature (for the tests
// that pin the solver to LAPACK)
void orc_peac_fit(const double* s9, int N, double* out8) {
  Stats st;
  st.sx = s9[0]; st.sy = s9[1]; st.sz = s9[2]; st.sxx = s9[3]; st.syy = s9[4]; st.szz = s9[5]; st.sxy = s9[6]; st.syz = s9[7]; st.sxz = s9[8];
  st.N = N;
  st.compute(out8, out8 + 3, out8[6], out8[7]);
}

// PlaneDetection::readDepthImage (PlaneExtractor.cpp:28-55): cloud [H*W][3] doubles from a 16-bit depth image
void orc_peac_cloud(const uint16_t* depth, int width, int height, int row_stride, float depth_factor, float fx, float fy, float cx, float cy,
                    double* cloud) {
  size_t v = 0;
  for (int i = 0; i < height; ++i)
    for (int j = 0; j < width; ++j, ++v) {
      const double z = (double)depth[(size_t)i * row_stride + j] * depth_factor;
      if (std::isnan(z)) { cloud[3 * v] = 0; cloud[3 * v + 1] = 0; cloud[3 * v + 2] = z; continue; }
      if (z > 5.0) { cloud[3 * v] = cloud[3 * v + 1] = cloud[3 * v + 2] = 0; continue; }
      cloud[3 * v] = ((double)j - cx) * z / fx;
      cloud[3 * v + 1] = ((double)i - cy) * z / fy;
      cloud[3 * v + 2] = z;
    }
}

// ahc::PlaneFitter::run with doRefine = true on a cloud of doubles.  prm: the 11 doubles of ParamSet in declaration order, then
// minSupport, windowWidth, windowHeight.  Out: seg_out [H*W] u8 (plid + 1, 0 elsewhere), planes [nplanes][12] = normal[3],
// center[3], mse, curvature, N, rid, 0, 0; member_offsets [nplanes + 1] into member_idx (pixel indices in scan order).
// Returns the number of planes (or -1 when plane_cap / member_cap are too small).
int orc_peac_run(const double* cloud, int width, int height, const double* prm11, int minSupport, int windowWidth, int windowHeight, uint8_t* seg_out,
                 double* planes, int plane_cap, int* member_offsets, int* member_idx, int member_cap, int* steps_out) {
  Fitter F;
  Params& P = F.prm;
  P.depthSigma = prm11[0]; P.stdTol_init = prm11[1]; P.stdTol_merge = prm11[2]; P.z_near = prm11[3]; P.z_far = prm11[4];
  P.angle_near = prm11[5]; P.angle_far = prm11[6]; P.similarityTh_merge = prm11[7]; P.similarityTh_refine = prm11[8];
  P.depthAlpha = prm11[9]; P.depthChangeTol = prm11[10];
  P.minSupport = minSupport; P.windowWidth = windowWidth; P.windowHeight = windowHeight;
  F.width = width; F.height = height; F.cloud = cloud;
  DisjointSet ds((height / windowHeight) * (width / windowWidth));
  F.ds = &ds;
  Fitter::MinQ minQ;
  F.initGraph(minQ);
  int steps = F.ahCluster(minQ);
  // refineDetails, AHCPlaneFitter.hpp:298-382
  std::vector<char> isValid;
  F.findBlockMembership(isValid);
  F.floodFill();
  std::vector<int> old;
  old.swap(F.extracted);
  Fitter::MinQ minQ2;
  for (int i = 0; i < (int)old.size(); ++i)
    if (isValid[i]) minQ2.push(Fitter::QE(F.segs[old[i]].mse, old[i]));
  steps += F.ahCluster(minQ2);
  if (steps_out) *steps_out = steps;
  std::vector<int> plidmap(old.size(), -1);
  const int nFinal = (int)F.extracted.size();
  for (int i = 0; i < (int)old.size(); ++i) {
    if (!isValid[i]) continue;
    const int np_rid = ds.Find(F.segs[old[i]].rid);
    for (int j = 0; j < nFinal; ++j)
      if (np_rid == F.segs[F.extracted[j]].rid) { plidmap[i] = j; break; }
  }
  if (nFinal > plane_cap) return -1;
  std::vector<std::vector<int>> mem(nFinal);
  const int nPixels = width * height;
  std::memset(seg_out, 0, (size_t)nPixels);
  for (int i = 0; i < nPixels; ++i) {
    int& plid = F.membership[i];
    if (plid >= 0 && plidmap[plid] >= 0) {
      plid = plidmap[plid];
      seg_out[i] = (uint8_t)(plid + 1);
      mem[plid].push_back(i);
    }
  }
  int off = 0;
  for (int p = 0; p < nFinal; ++p) {
    const Seg& s = F.segs[F.extracted[p]];
    double* o = planes + 12 * p;
    for (int k = 0; k < 3; ++k) { o[k] = s.normal[k]; o[3 + k] = s.center[k]; }
    o[6] = s.mse; o[7] = s.curvature; o[8] = s.N; o[9] = s.rid; o[10] = o[11] = 0;
    member_offsets[p] = off;
    if (off + (int)mem[p].size() > member_cap) return -1;
    std::memcpy(member_idx + off, mem[p].data(), mem[p].size() * sizeof(int));
    off += (int)mem[p].size();
  }
  member_offsets[nFinal] = off;
  return nFinal;
}

}  // extern "C"
