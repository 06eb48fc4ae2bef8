// drfe PEAC-AHC plane extraction for sm_100a — the plane extractor that is live in DR-SLAM's Frame constructor:
//   PlaneDetection::readDepthImage / runPlaneDetection          reference src/PlaneExtractor.cpp:28-63
//   ahc::PlaneFitter<ImagePointCloud>::run (doRefine)           include/peac/AHCPlaneFitter.hpp:211-259
//     initGraph :804-965, ahCluster :976-1190, refineDetails :298-382, findBlockMembership :494-600, floodFill :434-488
//   ahc::PlaneSeg / Stats                                       include/peac/AHCPlaneSeg.hpp:57-409
// batched over independent frames.  Built with --fmad=false: every double operation is rounded on its own, like the CPU oracle
// (the test oracle's PEAC restatement; its declared orders P.1 - P.5).
//
// Kernels:
//   k_peac_blocks     thread per 10x10 window: the window's 100 points in the reference's order (X, Y by true double division),
//                     missing-data / depth-discontinuity rejection, nine double sums, PCA (Jacobi) -> node record
//   k_peac_frame      one CTA per frame:
//       edges         the two passes of initGraph (a thread per block row, then per block column: the loops carry state)
//       cluster       the pop-min / best-neighbour / merge loop.  The queue is an argmin over (mse, creation number) of the queued
//                     nodes (a total order, so any priority structure pops the same sequence); a merged node takes over the slot
//                     of the popped one, whose neighbours are its neighbours anyway, adjacency is a bit matrix over slots with
//                     dead slots skipped; the merge candidates of a step are fitted in parallel, one thread each, the best one is
//                     found by shared-memory atomics on the ordered bit patterns of the mse, and the thread that fitted it — the fit
//                     is still in its registers — writes the merged node.  Every thread caches the minimum of its own queue slots
//       membership    block erosion (ERODE_ALL_BORDER) and the seed pixels of the region growing, in the reference's order by a scan
//       floodFill     the FIFO region grower, 256 queue entries at a time (two per thread).  Its result depends on the visiting order
//                     only through the state of the visited pixel, so the 1024 visits of a chunk run in rounds: in each round every
//                     pixel takes the earliest of its pending visits (atomicMin of the visit number), which keeps the per-pixel order
//                     of the sequential loop; what a visit computes without that state (vertex, distance, 3-sigma test) is done before
//                     the rounds; the pixels claimed by a chunk are appended to the queue in visit order by a scan.  A chunk ends at
//                     the queue length it started with
//       cluster again on the planes the region growing connected, plane numbering, seg_output
//   k_peac_members    plane_vertices_ and the member points as Frame::ComputePlanes reads them (optionally without the points
//                     beyond mMax_point_dist: the lists drfe_peac_plane_points_voxel filters)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "drfe_internal.h"

namespace drfe {

static const int kPeacThreads = 128;
static const int kFloodPer = 2;            // queue entries a thread takes per chunk of the region growing
static const int kPeacMaxPlanes = 255;     // seg_output is uchar (plid + 1)
static const int kPeacCand = 1024;         // merge candidates of one step

struct PeacNode {
  double s[9];                               // sx sy sz sxx syy szz sxy syz sxz (Stats)
  double mse, curvature, center[3], normal[3];
  double th_init;                            // T_ang(P_INIT, center z) of an initial node
  int N, rid, seq, ok;                       // ok: passed the init test (in the graph)
};

struct PeacDev {
  int W, H, winW, winH, Nw, Nh, NB, nwords, B, nslots, min_support, qcap;
  int flood_chunk;          // queue entries the region growing takes at a time (kFloodPer * kPeacThreads; 1 = the sequential loop, for checks)
  double depthSigma, stdTol_init, stdTol_merge, z_near, z_far, angle_near, angle_far, sim_merge, sim_refine, depthAlpha, depthChangeTol;
  float max_depth, depth_factor, fx, fy, cx, cy;
  const uint16_t* depth; long long depth_rs, depth_fs;
  PeacNode* nodes;          // [B][NB]
  uint32_t* adj;            // [nslots][NB][nwords] adjacency bit matrix over node slots
  int* trail;               // [nslots][H*W] membershipImg
  float* dist;              // [nslots][H*W] distMap
  int* first;               // [nslots][H*W] earliest pending visit of a pixel in the running chunk
  int2* queue;              // [nslots][qcap] rfQueue {pixel, plid}
  uint8_t* seg;             // [B][H*W]
  drfe_peac_plane* planes;  // [B][kPeacMaxPlanes]
  int* nplanes;             // [B]
  int* counters;            // [B][12] cluster steps, queue length, first-pass planes, -, then kilocycles at the end of each stage
  int* status;              // 1: more than 255 planes, 2: flood-fill queue overflow, 4: too many merge candidates
};

// ---- cyclic Jacobi, operation for operation the oracle's eig3_sym (peac_oracle.cpp; a pivot that no longer changes the
// diagonal is zeroed from the fourth sweep on)
__device__ void peac_eig3(const double in[6], double w[3], double v[3][3]) {
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
      if (sweep > 2 && fabs(app) + g == fabs(app) && fabs(aqq) + g == fabs(aqq)) {
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

// Stats::compute (AHCPlaneSeg.hpp:128-162)
__device__ void peac_compute(const double s[9], int N, double center[3], double normal[3], double& mse, double& curvature) {
  const double sc = 1.0 / (double)N;
  center[0] = s[0] * sc; center[1] = s[1] * sc; center[2] = s[2] * sc;
  const double K[6] = {s[3] - s[0] * s[0] * sc, s[6] - s[0] * s[1] * sc, s[8] - s[0] * s[2] * sc,
                       s[4] - s[1] * s[1] * sc, s[7] - s[1] * s[2] * sc, s[5] - s[2] * s[2] * sc};
  double sv[3], V[3][3];
  peac_eig3(K, sv, V);
  if (V[0][0] * center[0] + V[1][0] * center[1] + V[2][0] * center[2] <= 0) { normal[0] = V[0][0]; normal[1] = V[1][0]; normal[2] = V[2][0]; }
  else { normal[0] = -V[0][0]; normal[1] = -V[1][0]; normal[2] = -V[2][0]; }
  mse = sv[0] * sc;
  curvature = sv[0] / (sv[0] + sv[1] + sv[2]);
}

// depth of pixel (i, j) as readDepthImage stores it (0 when culled); false when ImagePointCloud::get would fail
__device__ __forceinline__ bool peac_z(const PeacDev& P, int f, int i, int j, double& z) {
  const uint16_t d = __ldg(P.depth + (long long)f * P.depth_fs + (long long)i * P.depth_rs + j);
  z = (double)d * (double)P.depth_factor;
  if (z > (double)P.max_depth) z = 0.0;
  return z != 0.0;
}
__device__ __forceinline__ bool peac_get(const PeacDev& P, int f, int i, int j, double& x, double& y, double& z) {
  if (!peac_z(P, f, i, j, z)) return false;
  x = ((double)j - (double)P.cx) * z / (double)P.fx;
  y = ((double)i - (double)P.cy) * z / (double)P.fy;
  return true;
}

// ------------------------------------------------------------------ initial nodes (PlaneSeg ctor, AHCPlaneSeg.hpp:213-290, INIT_STRICT)
__global__ void __launch_bounds__(128) k_peac_blocks(const PeacDev* __restrict__ Pp, int nframes) {
  const PeacDev& P = *Pp;
  const int gid = blockIdx.x * 128 + threadIdx.x;
  if (gid >= nframes * P.NB) return;
  const int f = gid / P.NB, blk = gid - f * P.NB;
  const int bi = blk / P.Nw, bj = blk - bi * P.Nw;
  const int seed_row = bi * P.winH, seed_col = bj * P.winW;
  double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int N = 0;
  bool valid = true;
  for (int i = seed_row, ic = 0; ic < P.winH && i < P.H && valid; ++i, ++ic) {
    for (int j = seed_col, jc = 0; jc < P.winW && j < P.W; ++j, ++jc) {
      double x, y, z, zn;
      if (!peac_get(P, f, i, j, x, y, z)) { valid = false; break; }
      if (j + 1 < P.W && peac_z(P, f, i, j + 1, zn) && fabs(z - zn) > P.depthAlpha * fabs(z) + P.depthChangeTol) { valid = false; break; }
      if (i + 1 < P.H && peac_z(P, f, i + 1, j, zn) && fabs(z - zn) > P.depthAlpha * fabs(z) + P.depthChangeTol) { valid = false; break; }
      s[0] += x; s[1] += y; s[2] += z;
      s[3] += x * x; s[4] += y * y; s[5] += z * z;
      s[6] += x * y; s[7] += y * z; s[8] += x * z;
      ++N;
    }
  }
  PeacNode nd;
  if (!valid) { N = 0; for (int k = 0; k < 9; ++k) s[k] = 0; }
  for (int k = 0; k < 9; ++k) nd.s[k] = s[k];
  nd.N = N; nd.rid = blk; nd.seq = blk; nd.ok = 0;
  nd.mse = nd.curvature = nd.th_init = 0;
  nd.center[0] = nd.center[1] = nd.center[2] = nd.normal[0] = nd.normal[1] = nd.normal[2] = 0;
  if (valid && N >= 4) {
    peac_compute(s, N, nd.center, nd.normal, nd.mse, nd.curvature);
    const double t = P.depthSigma * nd.center[2] * nd.center[2] + P.stdTol_init;      // T_mse(P_INIT)
    if (nd.mse < t * t) {
      nd.ok = 1;
      double cz = nd.center[2];                                                         // T_ang(P_INIT), AHCParamSet.hpp:110-118
      cz = fmax(cz, P.z_near);
      cz = fmin(cz, P.z_far);
      const double factor = (P.angle_far - P.angle_near) / (P.z_far - P.z_near);
      nd.th_init = cos(factor * cz + P.angle_near - factor * P.z_near);
    }
  }
  P.nodes[(long long)f * P.NB + blk] = nd;
}

__device__ __forceinline__ double peac_sim(const PeacNode& a, const PeacNode& b) {
  return fabs(a.normal[0] * b.normal[0] + a.normal[1] * b.normal[1] + a.normal[2] * b.normal[2]);
}

// what the shared memory of k_peac_frame holds
struct PeacShared {
  double* qmse;      // [NB] mse of the queued node in a slot, +inf when the slot is not queued
  int* qseq;         // [NB] creation number of the node in a slot
  int* parent;       // [NB] DisjointSet
  int* dsize;        // [NB]
  uint32_t* alive;   // [nwords] slot takes part in the graph
  int* cand;         // [kPeacCand] neighbour slots of the popped node
  double* cmse;      // [kPeacCand] mse of the merge with that neighbour (+inf: not similar enough)
  int* extracted;    // [kPeacMaxPlanes + 1]
  uint8_t* cpass;    // [kPeacCand] the merge with that neighbour passes T_mse(P_MERGING)
};

__device__ __forceinline__ int ds_find(const int* parent, int x) {      // no path compression: the root is the same
  while (parent[x] != x) x = parent[x];
  return x;
}
__device__ __forceinline__ void ds_union(int* parent, int* dsize, int x, int y) {   // DisjointSet::Union, DisjointSet.hpp:60-81
  const int xr = ds_find(parent, x), yr = ds_find(parent, y);
  if (xr == yr) return;
  if (dsize[xr] < dsize[yr]) { parent[xr] = yr; dsize[yr] += dsize[xr]; }
  else { parent[yr] = xr; dsize[xr] += dsize[yr]; }
}

// a double's place in the order of doubles as a signed 64-bit integer (NaNs sort after +inf: never popped, like `nan < x`)
__device__ __forceinline__ long long peac_key(double m) {
  const long long b = __double_as_longlong(m + 0.0);
  return b ^ ((b >> 63) & 0x7FFFFFFFFFFFFFFFLL);
}

// The clustering loop (ahCluster, AHCPlaneFitter.hpp:976-1190) over the slots that are queued in S.qmse.  Whole CTA.
// Returns the number of steps; extracted planes are appended to S.extracted (n_ext) in extraction order.
__device__ int peac_cluster(const PeacDev& P, PeacNode* nodes, uint32_t* adj, PeacShared& S, int& next_seq, int& n_ext, int* s_tmp, double* s_dtmp) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int NB = P.NB, nw = P.nwords;
  int steps = 0;
  bool dirty = true;
  const long long kInfKey = 0x7FF0000000000000LL;
  long long own_k = kInfKey;
  int own_s = -1, own_q = 0x7FFFFFFF;
  long long* s_ktmp = reinterpret_cast<long long*>(s_dtmp);
  long long* s_minkey = s_ktmp + 6;                             // smallest candidate key of the running step
#ifdef PEAC_PHASE_PROFILE
  long long ph[6] = {0, 0, 0, 0, 0, 0}, tph = clock64();
#define PH(i) { const long long t_ = clock64(); ph[i] += t_ - tph; tph = t_; }
#else
#define PH(i)
#endif
  for (;;) {
    // ---- pop: argmin of (mse, creation number) over the queued slots.  A thread keeps the minimum of its own slots
    // (i = tid mod kPeacThreads) and scans them again only after one of them changed.
    // The order of doubles is taken on their bit patterns (signed 64-bit after folding the negatives; -0.0 is made +0.0 first):
    // integer compares instead of the FP64 pipe's.
    if (dirty) {
      own_k = kInfKey; own_s = -1; own_q = 0x7FFFFFFF;
#pragma unroll 8
      for (int i = tid; i < NB; i += kPeacThreads) {              // branch-free, so that the loads of an unrolled group go out together
        const long long k = peac_key(S.qmse[i]);
        const int sq = S.qseq[i];
        const bool better = k < own_k || (k == own_k && sq < own_q);
        own_k = better ? k : own_k; own_s = better ? i : own_s; own_q = better ? sq : own_q;
      }
      if (own_k >= kInfKey) { own_k = kInfKey; own_s = -1; own_q = 0x7FFFFFFF; }
      dirty = false;
    }
    // the warp's minimum of (key, creation number) by three redux.sync steps: high word, low word among the lanes that hold the
    // minimal high word, creation number among those that hold the minimal key (creation numbers are unique)
    long long bk = own_k;
    int bs = own_s, bq = own_q;
    {
      const uint32_t hi = (uint32_t)((unsigned long long)own_k >> 32) ^ 0x80000000u;     // signed order as unsigned order
      const uint32_t mhi = __reduce_min_sync(0xFFFFFFFFu, hi);
      const uint32_t lo = hi == mhi ? (uint32_t)own_k : 0xFFFFFFFFu;
      const uint32_t mlo = __reduce_min_sync(0xFFFFFFFFu, lo);
      const bool tie = hi == mhi && (uint32_t)own_k == mlo;
      const uint32_t mq = __reduce_min_sync(0xFFFFFFFFu, tie ? (uint32_t)own_q : 0xFFFFFFFFu);
      const unsigned win = __ballot_sync(0xFFFFFFFFu, tie && (uint32_t)own_q == mq);
      const int src = __ffs(win) - 1;
      bk = __shfl_sync(0xFFFFFFFFu, own_k, src);
      bs = __shfl_sync(0xFFFFFFFFu, own_s, src);
      bq = __shfl_sync(0xFFFFFFFFu, own_q, src);
    }
    if (lane == 0) { s_ktmp[wid] = bk; s_tmp[wid] = bs; s_tmp[4 + wid] = bq; }
    __syncthreads();
    bk = s_ktmp[0]; bs = s_tmp[0]; bq = s_tmp[4];
    for (int w = 1; w < kPeacThreads / 32; ++w) {
      const long long ok = s_ktmp[w];
      const int os = s_tmp[w], oq = s_tmp[4 + w];
      if (os >= 0 && (bs < 0 || ok < bk || (ok == bk && oq < bq))) { bk = ok; bs = os; bq = oq; }
    }
    __syncthreads();
    if (bs < 0) break;                                           // queue empty
    PH(0)
    const int p = bs;
    if (tid == p % kPeacThreads) { S.qmse[p] = INFINITY; dirty = true; }   // popped
    if (tid == 0) { *s_minkey = kInfKey; s_tmp[16] = 0x7FFFFFFF; s_tmp[18] = 0; }
    // ---- the popped node's live neighbours, in ascending slot order
    uint32_t* row_p = adj + (long long)p * nw;
    int mycnt = 0;
    uint32_t myword = 0;
    if (tid < nw) { myword = row_p[tid] & S.alive[tid]; mycnt = __popc(myword); }
    // exclusive scan of the per-word counts over the CTA (nw <= 128 words handled by the first nw threads)
    int inc = mycnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_tmp[8 + wid] = inc;
    __syncthreads();
    int basec = 0, ncand = 0;
    for (int w = 0; w < kPeacThreads / 32; ++w) { if (w < wid) basec += s_tmp[8 + w]; ncand += s_tmp[8 + w]; }
    {
      int at = basec + inc - mycnt;
      while (myword) {
        const int b = __ffs(myword) - 1;
        myword &= myword - 1;
        if (at < kPeacCand) S.cand[at] = tid * 32 + b;
        ++at;
      }
    }
    if (ncand > kPeacCand) { if (tid == 0) atomicOr(P.status, 4); ncand = kPeacCand; }
#ifdef PEAC_PHASE_PROFILE
    ph[5] += ncand;
#endif
    __syncthreads();
    PH(1)
    // ---- fit the merge with every similar neighbour, one thread per candidate; the fit stays in the thread's registers, the
    // thread that fitted the chosen candidate finishes the step (a fit is ~8 k cycles of dependent FP64 operations).  Each fit leaves
    // its mse, whether it passes T_mse(P_MERGING), and takes part in a shared-memory minimum over the ordered bit patterns.
    const PeacNode& np = nodes[p];
    int my_c = -1, my_N = 0;
    double my_s[9], my_ctr[3], my_nrm[3], my_curv = 0, my_m = 0;
    for (int c = tid; c < ncand; c += kPeacThreads) {
      const PeacNode& nb = nodes[S.cand[c]];
      double m = INFINITY;
      bool pass = false;
      if (!(peac_sim(np, nb) < P.sim_merge)) {
#pragma unroll
        for (int k = 0; k < 9; ++k) my_s[k] = np.s[k] + nb.s[k];
        my_N = np.N + nb.N;
        peac_compute(my_s, my_N, my_ctr, my_nrm, my_m, my_curv);
        my_c = c;
        m = my_m;
        const double t = P.depthSigma * my_ctr[2] * my_ctr[2] + P.stdTol_merge;         // T_mse(P_MERGING)
        pass = my_m < t * t;
        if (!(m < INFINITY)) m = DBL_MAX;                         // (a NaN would never be chosen after another candidate; keep it last)
        atomicMin(s_minkey, peac_key(m));
      }
      S.cmse[c] = m;
      S.cpass[c] = pass ? 1 : 0;
    }
    __syncthreads();
    PH(2)
    // ---- the candidate the reference's scan keeps (ascending creation number; :1040-1051): the minimum mse; among exact ties
    // the scan `cand == 0 || cand.mse > m.mse || (cand.mse == m.mse && cand.N < m.mse)` depends on the visiting order and is
    // replayed literally over the tied candidates (earlier non-tied ones cannot survive a tie with the minimum, later ones cannot
    // replace it).
    {
      const long long mk = *s_minkey;
      if (mk < kInfKey)
        for (int c = tid; c < ncand; c += kPeacThreads)
          if (S.cmse[c] < INFINITY && peac_key(S.cmse[c]) == mk) { atomicAdd(&s_tmp[18], 1); atomicMin(&s_tmp[16], c); }
    }
    __syncthreads();
    int best = s_tmp[18] ? s_tmp[16] : -1;
    if (s_tmp[18] > 1) {                                          // exact ties (uniform branch): thread 0 replays the scan
      __syncthreads();
      if (tid == 0) {
        const double mn = S.cmse[best];
        const int tot = s_tmp[18];
        int bst = -1, best_N = 0, last_q = -1;
        double best_m = 0;
        for (int round = 0; round < tot; ++round) {
          int c_next = -1, q_next = 0x7FFFFFFF;
          for (int c = 0; c < ncand; ++c)
            if (S.cmse[c] == mn) { const int q = S.qseq[S.cand[c]]; if (q > last_q && q < q_next) { q_next = q; c_next = c; } }
          last_q = q_next;
          const int Nm = np.N + nodes[S.cand[c_next]].N;
          if (bst < 0 || best_m > mn || (best_m == mn && (double)best_N < mn)) { bst = c_next; best_m = mn; best_N = Nm; }
        }
        s_tmp[16] = bst;
      }
      __syncthreads();
      best = s_tmp[16];
    }
    PH(3)
    bool merged = false;
    if (best >= 0) {
      const int q = S.cand[best];
      const PeacNode& nb = nodes[q];
      const bool owner = tid == best % kPeacThreads;
      merged = S.cpass[best] != 0;
      if (merged) {
        const int rid_p = np.rid, rid_q = nb.rid;
        const int new_rid = np.N >= nb.N ? rid_p : rid_q;
        uint32_t* row_q = adj + (long long)q * nw;
        // the merged node's neighbours: (row p | row q) minus the two; every neighbour of q learns about slot p
        if (tid < nw) {
          const uint32_t wq = row_q[tid] & S.alive[tid];
          uint32_t w = row_p[tid] | wq;
          if ((p >> 5) == tid) w &= ~(1u << (p & 31));
          if ((q >> 5) == tid) w &= ~(1u << (q & 31));
          row_p[tid] = w;
          uint32_t it = wq;
          if ((p >> 5) == tid) it &= ~(1u << (p & 31));
          while (it) {
            const int b = __ffs(it) - 1;
            it &= it - 1;
            atomicOr(adj + (long long)(tid * 32 + b) * nw + (p >> 5), 1u << (p & 31));
          }
        }
        __syncthreads();                                          // every thread has read nodes[p] / nodes[q] / alive
        if (owner) {
          if (my_c != best) {                                     // more than kPeacThreads candidates and a later fit took the registers
#pragma unroll
            for (int k = 0; k < 9; ++k) my_s[k] = np.s[k] + nb.s[k];
            my_N = np.N + nb.N;
            peac_compute(my_s, my_N, my_ctr, my_nrm, my_m, my_curv);
          }
          PeacNode nn;
          for (int k = 0; k < 9; ++k) nn.s[k] = my_s[k];
          nn.mse = my_m; nn.curvature = my_curv; nn.th_init = 0;
          for (int k = 0; k < 3; ++k) { nn.center[k] = my_ctr[k]; nn.normal[k] = my_nrm[k]; }
          nn.N = my_N; nn.rid = new_rid; nn.seq = next_seq; nn.ok = 1;
          nodes[p] = nn;
          ds_union(S.parent, S.dsize, rid_p, rid_q);
          S.alive[q >> 5] &= ~(1u << (q & 31));
          S.qmse[q] = INFINITY;
          S.qmse[p] = my_m;
          S.qseq[p] = next_seq;
        }
        if (tid == p % kPeacThreads || tid == q % kPeacThreads) dirty = true;
        ++next_seq;
      }
    }
    if (!merged) {
      if (tid == 0) {
        if (np.N >= P.min_support) {
          if (n_ext < kPeacMaxPlanes) S.extracted[n_ext] = p; else atomicOr(P.status, 1);
        }
        S.alive[p >> 5] &= ~(1u << (p & 31));
      }
      if (np.N >= P.min_support && n_ext < kPeacMaxPlanes) ++n_ext;
    }
    ++steps;
    __syncthreads();
    PH(4)
  }
#ifdef PEAC_PHASE_PROFILE
  if (tid == 0 && steps > 100) { for (int i = 0; i < 5; ++i) P.counters[12 * blockIdx.x + 4 + i] = (int)(ph[i] >> 10); P.counters[12 * blockIdx.x + 11] = (int)(ph[5]); }
#endif
  // ---- std::sort(extractedPlanes, N decreasing), declared stable (P.4): insertion sort by thread 0
  if (tid == 0) {
    for (int i = 1; i < n_ext; ++i) {
      const int e = S.extracted[i], Ne = nodes[e].N;
      int j = i - 1;
      while (j >= 0 && nodes[S.extracted[j]].N < Ne) { S.extracted[j + 1] = S.extracted[j]; --j; }
      S.extracted[j + 1] = e;
    }
  }
  __syncthreads();
  return steps;
}

__global__ void __launch_bounds__(kPeacThreads) k_peac_frame(const PeacDev* __restrict__ Pp, int nframes) {
  extern __shared__ __align__(16) uint8_t smem[];
  const PeacDev& P = *Pp;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int NB = P.NB, nw = P.nwords, Nw = P.Nw, Nh = P.Nh, W = P.W, H = P.H, winW = P.winW, winH = P.winH;
  PeacShared S;
  {
    uint8_t* q = smem;
    S.qmse = (double*)q; q += sizeof(double) * NB;
    S.cmse = (double*)q; q += sizeof(double) * kPeacCand;
    S.qseq = (int*)q; q += sizeof(int) * NB;
    S.parent = (int*)q; q += sizeof(int) * NB;
    S.dsize = (int*)q; q += sizeof(int) * NB;
    S.cand = (int*)q; q += sizeof(int) * kPeacCand;
    S.alive = (uint32_t*)q; q += sizeof(uint32_t) * nw;
    S.extracted = (int*)q; q += sizeof(int) * (kPeacMaxPlanes + 1);
    S.cpass = q; q += kPeacCand;
  }
  __shared__ int s_tmp[32];
  __shared__ double s_dtmp[8];
  __shared__ int s_old[kPeacMaxPlanes + 1], s_plidmap[kPeacMaxPlanes + 1], s_valid[kPeacMaxPlanes + 1];
  __shared__ uint32_t s_padj[kPeacMaxPlanes + 1][8];            // plane adjacency found by the region growing (bit = plid)
  __shared__ int s_scan[kPeacThreads];
  __shared__ double s_pl[kPeacMaxPlanes + 1][7];            // normal, centre, mse of the first-pass planes (region growing)
  const int slot = blockIdx.x;
  uint32_t* adj = P.adj + (long long)slot * NB * nw;
  int* trail = P.trail + (long long)slot * W * H;
  float* dist = P.dist + (long long)slot * W * H;
  int* first = P.first + (long long)slot * W * H;
  int2* queue = P.queue + (long long)slot * P.qcap;
  for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
    PeacNode* nodes = P.nodes + (long long)f * NB;
    const long long t_begin = clock64();
    #ifdef PEAC_PHASE_PROFILE
    auto stamp = [&](int i) { if (tid == 0 && i >= 5) P.counters[12 * f + 4 + i] = (int)((clock64() - t_begin) >> 10); };
#else
    auto stamp = [&](int i) { if (tid == 0) P.counters[12 * f + 4 + i] = (int)((clock64() - t_begin) >> 10); };
#endif
    // ---- reset
    for (long long i = tid; i < (long long)NB * nw; i += kPeacThreads) adj[i] = 0;
    for (int i = tid; i < NB; i += kPeacThreads) {
      const bool ok = nodes[i].ok != 0;
      S.qmse[i] = ok ? nodes[i].mse : INFINITY;
      S.qseq[i] = i;
      S.parent[i] = i; S.dsize[i] = 1;
    }
    for (int i = tid; i < nw; i += kPeacThreads) S.alive[i] = 0;
    for (int i = tid; i < W * H; i += kPeacThreads) { trail[i] = -1; dist[i] = FLT_MAX; first[i] = 0x7FFFFFFF; }
    for (int i = tid; i < (kPeacMaxPlanes + 1) * 8; i += kPeacThreads) (&s_padj[0][0])[i] = 0;
    __syncthreads();
    for (int i = tid; i < NB; i += kPeacThreads)
      if (nodes[i].ok) atomicOr(&S.alive[i >> 5], 1u << (i & 31));
    __syncthreads();
    stamp(0);
    // ---- edges (AHCPlaneFitter.hpp:901-965): the loops skip and step back, so one thread walks one block row / column
    auto G = [&](int idx) { return nodes[idx].ok != 0; };
    auto connect = [&](int a, int b) {
      atomicOr(adj + (long long)a * nw + (b >> 5), 1u << (b & 31));
      atomicOr(adj + (long long)b * nw + (a >> 5), 1u << (a & 31));
    };
    for (int i = tid; i < Nh; i += kPeacThreads) {
      for (int j = 1; j < Nw; j += 2) {
        const int c = i * Nw + j;
        if (!G(c - 1)) { --j; continue; }
        if (!G(c)) continue;
        if (j < Nw - 1 && !G(c + 1)) { ++j; continue; }
        const double th = nodes[c].th_init;
        if ((j < Nw - 1 && peac_sim(nodes[c - 1], nodes[c + 1]) >= th) || (j == Nw - 1 && peac_sim(nodes[c], nodes[c - 1]) >= th)) {
          connect(c, c - 1);
          if (j < Nw - 1) connect(c, c + 1);
        } else {
          --j;
        }
      }
    }
    __syncthreads();
    for (int j = tid; j < Nw; j += kPeacThreads) {
      for (int i = 1; i < Nh; i += 2) {
        const int c = i * Nw + j;
        if (!G(c - Nw)) { --i; continue; }
        if (!G(c)) continue;
        if (i < Nh - 1 && !G(c + Nw)) { ++i; continue; }
        const double th = nodes[c].th_init;
        if ((i < Nh - 1 && peac_sim(nodes[c - Nw], nodes[c + Nw]) >= th) || (i == Nh - 1 && peac_sim(nodes[c], nodes[c - Nw]) >= th)) {
          connect(c, c - Nw);
          if (i < Nh - 1) connect(c, c + Nw);
        } else {
          --i;
        }
      }
    }
    __syncthreads();
    stamp(1);
    // ---- first clustering
    int next_seq = NB, n_ext = 0;
    int steps = peac_cluster(P, nodes, adj, S, next_seq, n_ext, s_tmp, s_dtmp);
    stamp(2);
    const int n_old = n_ext;
    // ---- findBlockMembership (:494-600, ERODE_ALL_BORDER): rid2plid as an array in S.qseq (free now), blkMap in S.cand? no: NB entries -> reuse S.qmse as ints
    int* rid2plid = S.qseq;
    int* blkMap = reinterpret_cast<int*>(S.qmse);                // NB ints fit in NB doubles
    for (int i = tid; i < NB; i += kPeacThreads) rid2plid[i] = -1;
    for (int i = tid; i <= kPeacMaxPlanes; i += kPeacThreads) { s_valid[i] = 0; s_old[i] = i < n_old ? S.extracted[i] : -1; }
    __syncthreads();
    for (int i = tid; i < n_old; i += kPeacThreads) rid2plid[nodes[s_old[i]].rid] = i;
    __syncthreads();
    const int NptsPerBlk = winH * winW;
    for (int b = tid; b < NB; b += kPeacThreads) {
      const int i = b / Nw, j = b - i * Nw;
      const int setid = ds_find(S.parent, b);
      int res = -1;
      if (S.dsize[setid] * NptsPerBlk >= P.min_support) {
        bool same = true;
        if (j > 0 && ds_find(S.parent, b - 1) != setid) same = false;
        if (same && j < Nw - 1 && ds_find(S.parent, b + 1) != setid) same = false;
        if (same && i > 0 && ds_find(S.parent, b - Nw) != setid) same = false;
        if (same && i < Nh - 1 && ds_find(S.parent, b + Nw) != setid) same = false;
        if (same) { res = rid2plid[setid]; if (res >= 0) s_valid[res] = 1; }
      }
      blkMap[b] = res;
    }
    __syncthreads();
    // membershipImg of the kept blocks
    for (int idx = tid; idx < NB * NptsPerBlk; idx += kPeacThreads) {
      const int b = idx / NptsPerBlk, r = idx - b * NptsPerBlk;
      const int plid = blkMap[b];
      if (plid >= 0) {
        const int bi = b / Nw, bj = b - bi * Nw;
        trail[(bi * winH + r / winW) * W + bj * winW + r % winW] = plid;
      }
    }
    // seeds of the region growing in the reference's order: per block a count, a scan over the blocks, then the writes
    int qlen = 0;
    {
      auto seeds_of = [&](int b, int2* out) -> int {
        const int i = b / Nw, j = b - i * Nw;
        int n = 0;
        if (blkMap[b] < 0) {
          if (i > 0 && blkMap[b - Nw] >= 0) {
            const int sp = (i * winH - 1) * W + j * winW;
            for (int k = 1; k < winW; ++k, ++n) if (out) out[n] = make_int2(sp + k, blkMap[b - Nw]);
          }
          if (j > 0 && blkMap[b - 1] >= 0) {
            const int sp = (i * winH) * W + j * winW - 1;
            for (int k = 0; k < winH - 1; ++k, ++n) if (out) out[n] = make_int2(sp + k * W, blkMap[b - 1]);
          }
        } else {
          const int plid = blkMap[b];
          if (i > 0 && blkMap[b - Nw] != plid) {
            const int sp = (i * winH) * W + j * winW;
            for (int k = 0; k < winW - 1; ++k, ++n) if (out) out[n] = make_int2(sp + k, plid);
          }
          if (j > 0 && blkMap[b - 1] != plid) {
            const int sp = (i * winH) * W + j * winW;
            for (int k = 1; k < winH; ++k, ++n) if (out) out[n] = make_int2(sp + k * W, plid);
          }
        }
        return n;
      };
      // contiguous chunk of blocks per thread keeps the order
      const int per = (NB + kPeacThreads - 1) / kPeacThreads;
      const int b0 = min(tid * per, NB), b1 = min(b0 + per, NB);
      int cnt = 0;
      for (int b = b0; b < b1; ++b) cnt += seeds_of(b, nullptr);
      s_scan[tid] = cnt;
      __syncthreads();
      int off = 0;
      for (int t = 0; t < kPeacThreads; ++t) { if (t < tid) off += s_scan[t]; qlen += s_scan[t]; }
      if (qlen > P.qcap) { if (tid == 0) atomicOr(P.status, 2); qlen = 0; }
      else
        for (int b = b0; b < b1; ++b) off += seeds_of(b, queue + off);
    }
    __syncthreads();
    stamp(3);
    // ---- floodFill (:434-488), kFloodPer * kPeacThreads queue entries per chunk.  Everything about a visit that does not depend
    // on mutable state is computed before the rounds (the vertex of the visited pixel, its distance to the visiting plane, the
    // 3-sigma test); a round then costs one trip to the L2 (first / trail / dist of every pending visit are requested together).
    for (int i = tid; i < n_old; i += kPeacThreads) {
      const PeacNode& pl = nodes[s_old[i]];
      s_pl[i][0] = pl.normal[0]; s_pl[i][1] = pl.normal[1]; s_pl[i][2] = pl.normal[2];
      s_pl[i][3] = pl.center[0]; s_pl[i][4] = pl.center[1]; s_pl[i][5] = pl.center[2];
      s_pl[i][6] = pl.mse;
    }
    __syncthreads();
    const double sim_refine = P.sim_refine;
    const int chunk = min(P.flood_chunk, kFloodPer * kPeacThreads);
    for (int q0 = 0, q1; q0 < qlen; q0 = q1) {
      q1 = min(q0 + chunk, qlen);                                   // the entries queued when the chunk starts
      int tgt[kFloodPer][4], plid[kFloodPer];
      float cd[kFloodPer][4];
      uint32_t pending = 0, pushed = 0, okv = 0;                    // bit 4 * j + it
#pragma unroll
      for (int j = 0; j < kFloodPer; ++j) {
        const int k = q0 + j * kPeacThreads + tid;
        plid[j] = 0;
#pragma unroll
        for (int it = 0; it < 4; ++it) { tgt[j][it] = -1; cd[j][it] = -1.f; }
        if (k >= q1) continue;
        const int2 e = queue[k];
        plid[j] = e.y;
        const int sy = e.x / W, sx = e.x - sy * W;
        // getValid4Neighbor order: left, right, up, down; only pixels outside the kept blocks change (inside, trail stays the block's plane)
        const int nbp[4] = {sx > 0 ? e.x - 1 : -1, sx < W - 1 ? e.x + 1 : -1, sy > 0 ? e.x - W : -1, sy < H - 1 ? e.x + W : -1};
        double z[4];
        bool has[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int c = nbp[it];
          has[it] = false; z[it] = 0;
          if (c < 0) continue;
          const int cy = c / W, cx = c - cy * W;
          const int by = cy / winH, bx = cx / winW;
          const int blk = (by < Nh && bx < Nw) ? by * Nw + bx : -1;
          if (blk >= 0 && blkMap[blk] >= 0) continue;
          tgt[j][it] = c;
          pending |= 1u << (4 * j + it);
          has[it] = peac_z(P, f, cy, cx, z[it]);
        }
        const double* pl = s_pl[plid[j]];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          if (!has[it]) continue;
          const int c = tgt[j][it];
          const int cy = c / W, cx = c - cy * W;
          const double x = ((double)cx - (double)P.cx) * z[it] / (double)P.fx;
          const double y = ((double)cy - (double)P.cy) * z[it] / (double)P.fy;
          const double sd = pl[0] * (x - pl[3]) + pl[1] * (y - pl[4]) + pl[2] * (z[it] - pl[5]);
          const float cdist = (float)fabs(sd);
          cd[j][it] = cdist;
          if ((double)cdist * (double)cdist < 9 * pl[6] + 1e-5) okv |= 1u << (4 * j + it);
        }
      }
      // rounds: every pixel serves the earliest of its pending visits (visit number = 4 * entry of the chunk + neighbour slot)
      for (;;) {
#pragma unroll
        for (int j = 0; j < kFloodPer; ++j)
#pragma unroll
          for (int it = 0; it < 4; ++it)
            if (pending >> (4 * j + it) & 1u) atomicMin(&first[tgt[j][it]], 4 * (j * kPeacThreads + tid) + it);
        const int any = __syncthreads_or(pending != 0);
        if (!any) break;
        // (first / trail / dist are written by other threads of the CTA, first by atomics that live in L2: read and write them
        // past the L1 with ld.cg / st.cg.  trail and dist are read before it is known whether the visit is served this round:
        // the values are used by the winner only, and nobody else writes that pixel in this round.)
        int fv[kFloodPer][4], trv[kFloodPer][4];
        float dv[kFloodPer][4];
#pragma unroll
        for (int j = 0; j < kFloodPer; ++j)
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            fv[j][it] = -1; trv[j][it] = 0; dv[j][it] = 0.f;
            if (pending >> (4 * j + it) & 1u) {
              const int c = tgt[j][it];
              fv[j][it] = __ldcg(&first[c]); trv[j][it] = __ldcg(&trail[c]); dv[j][it] = __ldcg(&dist[c]);
            }
          }
#pragma unroll
        for (int j = 0; j < kFloodPer; ++j)
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const uint32_t bit = 1u << (4 * j + it);
            if (!(pending & bit)) continue;
            if (fv[j][it] != 4 * (j * kPeacThreads + tid) + it) continue;
            const int c = tgt[j][it];
            pending &= ~bit;
            __stcg(&first[c], 0x7FFFFFFF);
            const int tr = trv[j][it];
            if (tr <= -6) continue;
            if (tr >= 0 && tr == plid[j]) continue;
            if (okv & bit) {
              if (tr >= 0) {
                const double* pl = s_pl[plid[j]];
                const double* npl = s_pl[tr];
                const double sim = fabs(pl[0] * npl[0] + pl[1] * npl[1] + pl[2] * npl[2]);
                if (sim >= sim_refine) {
                  atomicOr(&s_padj[tr][plid[j] >> 5], 1u << (plid[j] & 31));
                  atomicOr(&s_padj[plid[j]][tr >> 5], 1u << (tr & 31));
                }
              }
              if (cd[j][it] < dv[j][it]) {
                __stcg(&trail[c], plid[j]);
                __stcg(&dist[c], cd[j][it]);
                pushed |= bit;
              } else if (tr < 0) {
                __stcg(&trail[c], tr - 1);
              }
            } else {
              if (tr < 0) __stcg(&trail[c], tr - 1);
            }
          }
        __syncthreads();
      }
      // append the claimed pixels in visit order: all entries j = 0 of the chunk, then j = 1 (counts packed 16 bits each)
      uint32_t mine = 0;
#pragma unroll
      for (int j = 0; j < kFloodPer; ++j) mine |= (uint32_t)__popc((pushed >> (4 * j)) & 15u) << (16 * j);
      uint32_t inc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
      if (lane == 31) s_tmp[20 + wid] = (int)inc;
      __syncthreads();
      uint32_t basep = 0, totp = 0;
      for (int w = 0; w < kPeacThreads / 32; ++w) { if (w < wid) basep += (uint32_t)s_tmp[20 + w]; totp += (uint32_t)s_tmp[20 + w]; }
      int tot = 0;
#pragma unroll
      for (int j = 0; j < kFloodPer; ++j) tot += (int)((totp >> (16 * j)) & 0xFFFFu);
      if (qlen + tot > P.qcap) { if (tid == 0) atomicOr(P.status, 2); tot = 0; }
      else {
        int before = 0;                                             // pushes of the earlier j
#pragma unroll
        for (int j = 0; j < kFloodPer; ++j) {
          int at = qlen + before + (int)(((basep + inc - mine) >> (16 * j)) & 0xFFFFu);
#pragma unroll
          for (int it = 0; it < 4; ++it) if (pushed >> (4 * j + it) & 1u) queue[at++] = make_int2(tgt[j][it], plid[j]);
          before += (int)((totp >> (16 * j)) & 0xFFFFu);
        }
      }
      qlen += tot;
      __syncthreads();
    }
    // ---- one more clustering over the planes the region growing connected (:318-327)
    // planes keep their slots; adjacency rows of those slots are rebuilt from s_padj; only valid planes are queued
    __syncthreads();
    stamp(4);
    for (int i = tid; i < NB; i += kPeacThreads) { S.qmse[i] = INFINITY; S.qseq[i] = nodes[i].seq; }   // (blkMap / rid2plid are done)
    for (int i = tid; i < nw; i += kPeacThreads) S.alive[i] = 0;
    __syncthreads();
    for (int i = tid; i < n_old; i += kPeacThreads) {
      const int sl = s_old[i];
      for (int w = 0; w < nw; ++w) adj[(long long)sl * nw + w] = 0;
    }
    __syncthreads();
    for (int i = tid; i < n_old; i += kPeacThreads) {
      if (!s_valid[i]) continue;
      const int sl = s_old[i];
      atomicOr(&S.alive[sl >> 5], 1u << (sl & 31));
      S.qmse[sl] = nodes[sl].mse;
      for (int j = 0; j < n_old; ++j)
        if ((s_padj[i][j >> 5] >> (j & 31)) & 1u) { const int sj = s_old[j]; atomicOr(adj + (long long)sl * nw + (sj >> 5), 1u << (sj & 31)); }
    }
    __syncthreads();
    n_ext = 0;
    steps += peac_cluster(P, nodes, adj, S, next_seq, n_ext, s_tmp, s_dtmp);
    stamp(5);
    // ---- plane numbering (:329-343) and the outputs
    for (int i = tid; i <= kPeacMaxPlanes; i += kPeacThreads) {
      int m = -1;
      if (i < n_old && s_valid[i]) {
        const int np_rid = ds_find(S.parent, nodes[s_old[i]].rid);
        for (int j = 0; j < n_ext; ++j)
          if (nodes[S.extracted[j]].rid == np_rid) { m = j; break; }
      }
      s_plidmap[i] = m;
    }
    __syncthreads();
    for (int j = tid; j < n_ext; j += kPeacThreads) {
      const PeacNode& nd = nodes[S.extracted[j]];
      drfe_peac_plane o;
      for (int k = 0; k < 3; ++k) { o.normal[k] = nd.normal[k]; o.center[k] = nd.center[k]; }
      o.mse = nd.mse; o.curvature = nd.curvature; o.N = nd.N; o.rid = nd.rid;
      P.planes[(long long)f * kPeacMaxPlanes + j] = o;
    }
    if (tid == 0) {
      P.nplanes[f] = n_ext;
      P.counters[12 * f] = steps; P.counters[12 * f + 1] = qlen; P.counters[12 * f + 2] = n_old; P.counters[12 * f + 3] = 0;
    }
    uint8_t* seg = P.seg + (long long)f * W * H;
    for (int i = tid; i < (W * H) / 4; i += kPeacThreads) {
      const int4 t = __ldcg(reinterpret_cast<const int4*>(trail) + i);
      const int v[4] = {t.x, t.y, t.z, t.w};
      uint32_t word = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (v[k] >= 0 && s_plidmap[v[k]] >= 0) word |= (uint32_t)(s_plidmap[v[k]] + 1) << (8 * k);
      reinterpret_cast<uint32_t*>(seg)[i] = word;
    }
    for (int i = ((W * H) / 4) * 4 + tid; i < W * H; i += kPeacThreads) {
      const int v = __ldcg(&trail[i]);
      seg[i] = (v >= 0 && s_plidmap[v] >= 0) ? (uint8_t)(s_plidmap[v] + 1) : 0;
    }
    __syncthreads();
    stamp(6);
  }
}

// the plane a pixel is listed under: its label, or none when the caller culls far points (Frame::ComputePlanes drops the points
// with (float) z > mMax_point_dist before the voxel filter, Frame.cc:960-963)
__device__ __forceinline__ int member_key(const PeacDev& P, int f, const uint8_t* __restrict__ seg, int p, float max_z) {
  const int key = seg[p];
  if (key == 0 || max_z == FLT_MAX) return key;
  const int i = p / P.W, j = p - i * P.W;
  double z;
  peac_z(P, f, i, j, z);
  return (float)z > max_z ? 0 : key;
}

// plane_vertices_ and the member points, like k_cape_plane_points: count per (plane, warp run), scan, stable scatter
static const int kMemWarps = 32;
__global__ void __launch_bounds__(kMemWarps * 32) k_peac_members(const PeacDev* __restrict__ Pp, int* __restrict__ out_idx, float* __restrict__ out_pts,
                                                                  int* __restrict__ offsets, float max_z) {
  extern __shared__ int s_pos[];                              // [kMemWarps][np + 1]
  __shared__ int s_start[kPeacMaxPlanes + 2];
  const PeacDev& P = *Pp;
  const int f = blockIdx.x;
  const int np = min(P.nplanes[f], kPeacMaxPlanes);
  const int N = P.H * P.W;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  int* offs = offsets + (long long)f * (kPeacMaxPlanes + 1);
  if (np == 0) { if (tid == 0) offs[0] = 0; return; }
  const int stride = np + 1;
  for (int i = tid; i < kMemWarps * stride; i += kMemWarps * 32) s_pos[i] = 0;
  __syncthreads();
  const uint8_t* __restrict__ seg = P.seg + (long long)f * N;
  const int run = ((N + kMemWarps - 1) / kMemWarps + 31) & ~31;
  const int p0 = w * run, p1 = min(p0 + run, N);
  int* mine = s_pos + w * stride;
  for (int p = p0 + lane; p - lane < p1; p += 32) {
    const int key = p < p1 ? member_key(P, f, seg, p, max_z) : 0;
    const unsigned m = __match_any_sync(0xFFFFFFFFu, key);
    if (key && (m & lt) == 0) mine[key] += __popc(m);
    __syncwarp();                                                  // the next trip's leader for this key may be another lane
  }
  __syncthreads();
  if (tid < np) {
    int t = 0;
    for (int k = 0; k < kMemWarps; ++k) t += s_pos[k * stride + tid + 1];
    s_start[tid + 1] = t;
  }
  __syncthreads();
  if (tid == 0) {
    int tot = 0;
    for (int L = 1; L <= np; ++L) { const int t = s_start[L]; s_start[L] = tot; offs[L - 1] = tot; tot += t; }
    offs[np] = tot;
  }
  __syncthreads();
  if (tid < np) {
    int at = s_start[tid + 1];
    for (int k = 0; k < kMemWarps; ++k) { const int t = s_pos[k * stride + tid + 1]; s_pos[k * stride + tid + 1] = at; at += t; }
  }
  __syncthreads();
  int* oi = out_idx ? out_idx + (long long)f * N : nullptr;
  float* op = out_pts ? out_pts + (long long)f * N * 3 : nullptr;
  for (int p = p0 + lane; p - lane < p1; p += 32) {
    const int key = p < p1 ? member_key(P, f, seg, p, max_z) : 0;
    const unsigned m = __match_any_sync(0xFFFFFFFFu, key);
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (key && lane == leader) { base = mine[key]; mine[key] = base + __popc(m); }
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (key) {
      const int at = base + __popc(m & lt);
      if (oi) oi[at] = p;
      if (op) {
        const int i = p / P.W, j = p - i * P.W;
        double x = 0, y = 0, z = 0;
        if (!peac_get(P, f, i, j, x, y, z)) { x = 0; y = 0; }     // a culled vertex is (0, 0, 0)
        op[3 * at] = (float)x; op[3 * at + 1] = (float)y; op[3 * at + 2] = (float)z;
      }
    }
    __syncwarp();
  }
}

}  // namespace drfe

using namespace drfe;

struct drfe_peac {
  int device = 0, max_batch = 0, width = 0, height = 0;
  drfe_peac_params prm{};
  PeacDev hd{};
  PeacDev* dd = nullptr;
  cudaStream_t stream = nullptr;
  uint16_t* d_depth = nullptr;
  int* d_mem_idx = nullptr; float* d_mem_pts = nullptr; int* d_mem_offs = nullptr;
  VoxelScratch vox;
  float* d_third = nullptr;
  NormalsScratch nrm;
  std::vector<int> h_offs;
  size_t frame_smem = 0;
  int last_frames = 0;
  bool pending = false;
  std::vector<void*> allocs;
};

template <typename T>
static int peac_alloc(drfe_peac* h, T** p, size_t count) {
  void* q = nullptr;
  DRFE_CUDA(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  h->allocs.push_back(q);
  *p = (T*)q;
  return DRFE_OK;
}

extern "C" {

int drfe_peac_default_params(drfe_peac_params* p) {
  if (!p) return DRFE_ERR_ARG;
  // ahc::ParamSet::ParamSet (AHCParamSet.hpp:68-78) and PlaneFitter::PlaneFitter (AHCPlaneFitter.hpp:149-154)
  p->depthSigma = 1.6e-6; p->stdTol_init = 5; p->stdTol_merge = 8;
  p->z_near = 500; p->z_far = 4000;
  p->angle_near = 15.0 * M_PI / 180.0; p->angle_far = 90.0 * M_PI / 180.0;
  p->similarityTh_merge = std::cos(60.0 * M_PI / 180.0);
  p->similarityTh_refine = std::cos(30.0 * M_PI / 180.0);
  p->depthAlpha = 0.04; p->depthChangeTol = 0.02;
  p->min_support = 3000; p->window_width = 10; p->window_height = 10;
  p->max_depth = 5.0f;
  return DRFE_OK;
}

int drfe_peac_create(int width, int height, const drfe_peac_params* params, int max_batch, int device, drfe_peac** out) {
  if (!out) { set_error("drfe_peac_create: null argument"); return DRFE_ERR_ARG; }
  *out = nullptr;
  drfe_peac_params prm;
  if (params) prm = *params; else drfe_peac_default_params(&prm);
  if (width < 2 * prm.window_width || height < 2 * prm.window_height || prm.window_width < 2 || prm.window_height < 2 || max_batch < 1 ||
      prm.min_support < 1 || !(prm.z_far > prm.z_near) || (width & 3)) {
    set_error("drfe_peac_create: invalid parameters"); return DRFE_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("drfe_peac_create: no CUDA device available (there is no CPU fallback)"); return DRFE_ERR_CUDA; }
  if (device < 0 || device >= ndev) { set_error("drfe_peac_create: bad device %d", device); return DRFE_ERR_ARG; }
  DeviceScope ds(device);
  if (!ds.ok) { set_error("cudaSetDevice(%d) failed", device); return DRFE_ERR_CUDA; }
  drfe_peac* h = new drfe_peac();
  h->device = device; h->max_batch = max_batch; h->width = width; h->height = height; h->prm = prm;
  PeacDev& D = h->hd;
  memset(&D, 0, sizeof(D));
  D.W = width; D.H = height; D.winW = prm.window_width; D.winH = prm.window_height;
  D.Nw = width / D.winW; D.Nh = height / D.winH; D.NB = D.Nw * D.Nh; D.nwords = (D.NB + 31) / 32; D.B = max_batch;
  D.min_support = prm.min_support;
  { const char* e = getenv("DRFE_PEAC_FLOOD_CHUNK"); D.flood_chunk = e ? std::min(std::max(atoi(e), 1), kFloodPer * kPeacThreads) : kFloodPer * kPeacThreads; }
  D.depthSigma = prm.depthSigma; D.stdTol_init = prm.stdTol_init; D.stdTol_merge = prm.stdTol_merge; D.z_near = prm.z_near; D.z_far = prm.z_far;
  D.angle_near = prm.angle_near; D.angle_far = prm.angle_far; D.sim_merge = prm.similarityTh_merge; D.sim_refine = prm.similarityTh_refine;
  D.depthAlpha = prm.depthAlpha; D.depthChangeTol = prm.depthChangeTol; D.max_depth = prm.max_depth;
  auto fail = [&](int code) { drfe_peac_destroy(h); return code; };
  if (D.nwords > kPeacThreads) { set_error("drfe_peac_create: %d windows are too many for the clustering kernel (at most %d)", D.NB, 32 * kPeacThreads); return fail(DRFE_ERR_ARG); }
  h->frame_smem = sizeof(double) * D.NB + sizeof(double) * kPeacCand + sizeof(int) * 3 * D.NB + sizeof(int) * kPeacCand + sizeof(uint32_t) * D.nwords +
                  sizeof(int) * (kPeacMaxPlanes + 1) + kPeacCand + 64;
  if (h->frame_smem > 200 * 1024) { set_error("drfe_peac_create: %d windows need %zu bytes of shared memory", D.NB, h->frame_smem); return fail(DRFE_ERR_ARG); }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  D.nslots = std::min(max_batch, 2 * sms);
  D.qcap = 3 * width * height;
  const size_t N = (size_t)width * height, B = max_batch, S = D.nslots;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(DRFE_ERR_CUDA); }
  int rc = DRFE_OK;
  rc |= peac_alloc(h, &D.nodes, (size_t)D.NB * B);
  rc |= peac_alloc(h, &D.adj, (size_t)D.NB * D.nwords * S);
  rc |= peac_alloc(h, &D.trail, N * S);
  rc |= peac_alloc(h, &D.dist, N * S);
  rc |= peac_alloc(h, &D.first, N * S);
  rc |= peac_alloc(h, &D.queue, (size_t)D.qcap * S);
  rc |= peac_alloc(h, &D.seg, N * B);
  rc |= peac_alloc(h, &D.planes, (size_t)kPeacMaxPlanes * B);
  rc |= peac_alloc(h, &D.nplanes, B);
  rc |= peac_alloc(h, &D.counters, 12 * B);
  rc |= peac_alloc(h, &D.status, 1);
  rc |= peac_alloc(h, &h->d_depth, N * B);
  rc |= peac_alloc(h, &h->dd, 1);
  if (rc) return fail(DRFE_ERR_CUDA);
  if (cudaMemset(D.status, 0, sizeof(int)) != cudaSuccess) { set_error("cudaMemset failed"); return fail(DRFE_ERR_CUDA); }
  if (raise_dyn_smem(k_peac_frame, device, h->frame_smem) != cudaSuccess) { set_error("cudaFuncSetAttribute failed"); return fail(DRFE_ERR_CUDA); }
  *out = h;
  return DRFE_OK;
}

int drfe_peac_destroy(drfe_peac* h) {
  if (!h) return DRFE_OK;
  DeviceScope ds(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  voxel_scratch_free(h->vox);
  normals_scratch_free(h->nrm);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return DRFE_OK;
}

void* drfe_peac_stream(drfe_peac* h) { return h ? (void*)h->stream : nullptr; }
int drfe_peac_sync(drfe_peac* h) {
  if (!h) return DRFE_ERR_ARG;
  DeviceScope ds(h->device);
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  return DRFE_OK;
}

int drfe_peac_enqueue_depth_u16(drfe_peac* h, int nframes, const uint16_t* depth, size_t row_stride, size_t frame_stride, int mem_kind,
                                float depth_factor, float fx, float fy, float cx, float cy) {
  NvtxRange nvtx_("drfe_peac_enqueue_depth_u16");
  if (!h || !depth) { set_error("drfe_peac_enqueue_depth_u16: null argument"); return DRFE_ERR_ARG; }
  if (nframes < 1 || nframes > h->max_batch) { set_error("drfe_peac_enqueue_depth_u16: nframes %d outside [1,%d]", nframes, h->max_batch); return DRFE_ERR_ARG; }
  const int W = h->width, H = h->height;
  if (row_stride < (size_t)W || (nframes > 1 && frame_stride < row_stride * H)) { set_error("drfe_peac_enqueue_depth_u16: bad strides"); return DRFE_ERR_ARG; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  if (mem_kind == DRFE_MEM_HOST) {
    for (int f = 0; f < nframes; ++f)
      DRFE_CUDA(cudaMemcpy2DAsync(h->d_depth + (size_t)f * W * H, W * sizeof(uint16_t), depth + f * frame_stride, row_stride * sizeof(uint16_t),
                                  W * sizeof(uint16_t), H, cudaMemcpyHostToDevice, st));
    h->hd.depth = h->d_depth; h->hd.depth_rs = W; h->hd.depth_fs = (long long)W * H;
  } else if (mem_kind == DRFE_MEM_DEVICE) {
    h->hd.depth = depth; h->hd.depth_rs = (long long)row_stride; h->hd.depth_fs = (long long)frame_stride;
  } else { set_error("drfe_peac_enqueue_depth_u16: bad mem_kind"); return DRFE_ERR_ARG; }
  h->hd.depth_factor = depth_factor; h->hd.fx = fx; h->hd.fy = fy; h->hd.cx = cx; h->hd.cy = cy;
  DRFE_CUDA(cudaMemcpyAsync(h->dd, &h->hd, sizeof(PeacDev), cudaMemcpyHostToDevice, st));
  DRFE_LAUNCH(k_peac_blocks, (nframes * h->hd.NB + 127) / 128, 128, 0, st, h->dd, nframes);
  DRFE_LAUNCH(k_peac_frame, std::min(nframes, h->hd.nslots), kPeacThreads, h->frame_smem, st, h->dd, nframes);
  h->last_frames = nframes;
  h->pending = true;
  return DRFE_OK;
}

static int peac_status(drfe_peac* h, int status) {
  if (!status) return DRFE_OK;
  cudaMemsetAsync(h->hd.status, 0, sizeof(int), h->stream);
  set_error("PEAC: device-side capacity exceeded (status %d: 1 = more than %d planes, 2 = region-growing queue, 4 = merge candidates)", status, kPeacMaxPlanes);
  return DRFE_ERR_CAPACITY;
}

int drfe_peac_download(drfe_peac* h, uint8_t* seg_out, drfe_peac_plane* planes, int plane_cap, int* nr_planes) {
  NvtxRange nvtx_("drfe_peac_download");
  if (!h || !nr_planes) { set_error("drfe_peac_download: null argument"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("drfe_peac_download: nothing enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  cudaStream_t st = h->stream;
  const int nf = h->last_frames;
  const size_t N = (size_t)h->width * h->height;
  int status = 0;
  DRFE_CUDA(cudaMemcpyAsync(nr_planes, h->hd.nplanes, nf * sizeof(int), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaMemcpyAsync(&status, h->hd.status, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (seg_out) DRFE_CUDA(cudaMemcpyAsync(seg_out, h->hd.seg, N * nf, cudaMemcpyDeviceToHost, st));
  if (planes && plane_cap > 0)
    DRFE_CUDA(cudaMemcpy2DAsync(planes, (size_t)plane_cap * sizeof(drfe_peac_plane), h->hd.planes, (size_t)kPeacMaxPlanes * sizeof(drfe_peac_plane),
                                (size_t)std::min(plane_cap, kPeacMaxPlanes) * sizeof(drfe_peac_plane), nf, cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaStreamSynchronize(st));
  const int rc = peac_status(h, status);
  if (rc != DRFE_OK) return rc;
  if (planes)
    for (int f = 0; f < nf; ++f)
      if (nr_planes[f] > plane_cap) { set_error("drfe_peac_download: frame %d has %d planes, plane_cap is %d", f, nr_planes[f], plane_cap); return DRFE_ERR_CAPACITY; }
  return DRFE_OK;
}

int drfe_peac_plane_vertices(drfe_peac* h, int32_t* indices, float* points, size_t cap_per_frame, int* offsets, int plane_cap) {
  NvtxRange nvtx_("drfe_peac_plane_vertices");
  if (!h || !offsets || plane_cap < 1) { set_error("drfe_peac_plane_vertices: bad argument"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("drfe_peac_plane_vertices: nothing enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  const int nf = h->last_frames;
  const size_t N = (size_t)h->width * h->height;
  if (!h->d_mem_offs) {
    if (peac_alloc(h, &h->d_mem_idx, (size_t)h->max_batch * N) || peac_alloc(h, &h->d_mem_pts, (size_t)h->max_batch * N * 3) ||
        peac_alloc(h, &h->d_mem_offs, (size_t)h->max_batch * (kPeacMaxPlanes + 1))) return DRFE_ERR_CUDA;
    h->h_offs.resize((size_t)h->max_batch * (kPeacMaxPlanes + 1));
    DRFE_CUDA(raise_dyn_smem(k_peac_members, h->device, (size_t)kMemWarps * (kPeacMaxPlanes + 1) * sizeof(int)));
  }
  DRFE_LAUNCH(k_peac_members, nf, kMemWarps * 32, kMemWarps * (kPeacMaxPlanes + 1) * sizeof(int), st, h->dd, indices ? h->d_mem_idx : nullptr,
              points ? h->d_mem_pts : nullptr, h->d_mem_offs, FLT_MAX);
  int* ho = h->h_offs.data();
  DRFE_CUDA(cudaMemcpyAsync(ho, h->d_mem_offs, (size_t)nf * (kPeacMaxPlanes + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  std::vector<int> np(nf);
  DRFE_CUDA(cudaMemcpyAsync(np.data(), h->hd.nplanes, nf * sizeof(int), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaStreamSynchronize(st));
  for (int f = 0; f < nf; ++f) {
    const int n = std::min(np[f], kPeacMaxPlanes);
    if (n > plane_cap) { set_error("drfe_peac_plane_vertices: frame %d has %d planes, plane_cap is %d", f, n, plane_cap); return DRFE_ERR_CAPACITY; }
    const int* src = ho + (size_t)f * (kPeacMaxPlanes + 1);
    int* dst = offsets + (size_t)f * (plane_cap + 1);
    for (int i = 0; i <= n; ++i) dst[i] = src[i];
    for (int i = n + 1; i <= plane_cap; ++i) dst[i] = src[n];
    if ((size_t)src[n] > cap_per_frame) { set_error("drfe_peac_plane_vertices: frame %d has %d member pixels, cap_per_frame is %zu", f, src[n], cap_per_frame); return DRFE_ERR_CAPACITY; }
    if (src[n] > 0) {
      if (indices) DRFE_CUDA(cudaMemcpyAsync(indices + (size_t)f * cap_per_frame, h->d_mem_idx + (size_t)f * N, (size_t)src[n] * sizeof(int), cudaMemcpyDeviceToHost, st));
      if (points) DRFE_CUDA(cudaMemcpyAsync(points + (size_t)f * cap_per_frame * 3, h->d_mem_pts + (size_t)f * N * 3, (size_t)src[n] * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
  }
  DRFE_CUDA(cudaStreamSynchronize(st));
  return DRFE_OK;
}

// Frame::ComputePlanes' per-plane clouds (Frame.cc:954-990): the vertices of every plane with (float) z <= max_point_dist, through
// pcl::VoxelGrid with a cubic leaf — the lists never leave the device, the centroids do
int drfe_peac_plane_points_voxel(drfe_peac* h, float max_point_dist, float leaf_size, float* points, size_t cap_per_frame, int* offsets, int plane_cap) {
  NvtxRange nvtx_("drfe_peac_plane_points_voxel");
  if (!h || !points || !offsets || plane_cap < 1 || !(leaf_size > 0.f)) { set_error("drfe_peac_plane_points_voxel: bad argument"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("drfe_peac_plane_points_voxel: nothing enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  const int nf = h->last_frames;
  const size_t N = (size_t)h->width * h->height;
  if (!h->d_mem_offs) {
    if (peac_alloc(h, &h->d_mem_idx, (size_t)h->max_batch * N) || peac_alloc(h, &h->d_mem_pts, (size_t)h->max_batch * N * 3) ||
        peac_alloc(h, &h->d_mem_offs, (size_t)h->max_batch * (kPeacMaxPlanes + 1))) return DRFE_ERR_CUDA;
    h->h_offs.resize((size_t)h->max_batch * (kPeacMaxPlanes + 1));
    DRFE_CUDA(raise_dyn_smem(k_peac_members, h->device, (size_t)kMemWarps * (kPeacMaxPlanes + 1) * sizeof(int)));
  }
  if (!h->vox.out && voxel_scratch_alloc(h->vox, (size_t)h->max_batch, N)) { set_error("drfe_peac_plane_points_voxel: cudaMalloc failed"); return DRFE_ERR_CUDA; }
  DRFE_LAUNCH(k_peac_members, nf, kMemWarps * 32, kMemWarps * (kPeacMaxPlanes + 1) * sizeof(int), st, h->dd, (int*)nullptr, h->d_mem_pts, h->d_mem_offs,
              max_point_dist);
  int rc = voxel_filter_launch(st, nf, h->d_mem_pts, h->d_mem_offs, h->hd.nplanes, (int)N, leaf_size, h->vox);
  if (rc != DRFE_OK) return rc;
  int* ho = h->h_offs.data();
  DRFE_CUDA(cudaMemcpyAsync(ho, h->vox.out_offs, (size_t)nf * (kPeacMaxPlanes + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  std::vector<int> np(nf);
  DRFE_CUDA(cudaMemcpyAsync(np.data(), h->hd.nplanes, nf * sizeof(int), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaStreamSynchronize(st));
  for (int f = 0; f < nf; ++f) {
    const int n = std::min(np[f], kPeacMaxPlanes);
    if (n > plane_cap) { set_error("drfe_peac_plane_points_voxel: frame %d has %d planes, plane_cap is %d", f, n, plane_cap); return DRFE_ERR_CAPACITY; }
    const int* src = ho + (size_t)f * (kPeacMaxPlanes + 1);
    int* dst = offsets + (size_t)f * (plane_cap + 1);
    for (int i = 0; i <= n; ++i) dst[i] = src[i];
    for (int i = n + 1; i <= plane_cap; ++i) dst[i] = src[n];
    if ((size_t)src[n] > cap_per_frame) { set_error("drfe_peac_plane_points_voxel: frame %d has %d centroids, cap_per_frame is %zu", f, src[n], cap_per_frame); return DRFE_ERR_CAPACITY; }
    if (src[n] > 0)
      DRFE_CUDA(cudaMemcpyAsync(points + (size_t)f * cap_per_frame * 3, h->vox.out + (size_t)f * N * 3, (size_t)src[n] * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  DRFE_CUDA(cudaStreamSynchronize(st));
  return DRFE_OK;
}

// the 1/3-resolution cloud of Frame::ComputePlanes (Frame.cc:1044-1066) and PCL's integral-image normals on it (:1068-1100), from the
// depth batch the handle was given (imDepth = float(raw) * depth factor)
int drfe_peac_third_cloud_normals(drfe_peac* h, float max_point_dist, float max_depth_change_factor, float normal_smoothing_size, float* cloud, float* normals) {
  NvtxRange nvtx_("drfe_peac_third_cloud_normals");
  if (!h || !normals || !(normal_smoothing_size >= 1.f)) { set_error("drfe_peac_third_cloud_normals: bad argument"); return DRFE_ERR_ARG; }
  if (!h->pending) { set_error("drfe_peac_third_cloud_normals: nothing enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  const int w3 = (h->width + 2) / 3, h3 = (h->height + 2) / 3;
  const size_t per = (size_t)w3 * h3 * 3;
  if (2 * (int)normal_smoothing_size >= std::min(w3, h3)) { set_error("drfe_peac_third_cloud_normals: smoothing size %g leaves no interior", (double)normal_smoothing_size); return DRFE_ERR_ARG; }
  if (!h->d_third && peac_alloc(h, &h->d_third, per * h->max_batch)) return DRFE_ERR_CUDA;
  if (!h->nrm.normals && normals_scratch_alloc(h->nrm, (size_t)h->max_batch, w3, h3)) { set_error("drfe_peac_third_cloud_normals: cudaMalloc failed"); return DRFE_ERR_CUDA; }
  cudaStream_t st = h->stream;
  const int nf = h->last_frames;
  int rc = third_cloud_u16_launch(st, nf, h->hd.depth, h->hd.depth_rs, h->hd.depth_fs, h->hd.depth_factor, h->hd.fx, h->hd.fy, h->hd.cx, h->hd.cy, h->width, h->height,
                                  max_point_dist, h->d_third);
  if (rc != DRFE_OK) return rc;
  rc = normals_launch(st, h->device, nf, h->d_third, w3, h3, max_depth_change_factor, normal_smoothing_size, h->nrm);
  if (rc != DRFE_OK) return rc;
  if (cloud) DRFE_CUDA(cudaMemcpyAsync(cloud, h->d_third, per * nf * sizeof(float), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaMemcpyAsync(normals, h->nrm.normals, per * nf * sizeof(float), cudaMemcpyDeviceToHost, st));
  DRFE_CUDA(cudaStreamSynchronize(st));
  return DRFE_OK;
}

int drfe_peac_debug_counters(drfe_peac* h, int frame, int32_t* out12) {
  if (!h || !out12) { set_error("drfe_peac_debug_counters: null argument"); return DRFE_ERR_ARG; }
  if (!h->pending || frame < 0 || frame >= h->last_frames) { set_error("drfe_peac_debug_counters: bad frame"); return DRFE_ERR_STATE; }
  DeviceScope ds(h->device);
  DRFE_CUDA(cudaStreamSynchronize(h->stream));
  DRFE_CUDA(cudaMemcpy(out12, h->hd.counters + 12 * frame, 12 * sizeof(int), cudaMemcpyDeviceToHost));
  return DRFE_OK;
}

}  // extern "C"
