// ORACLE — test infrastructure only (never linked into libdrfe.so).  Second, independent restatement (C++, std::vector /
// std::map as the reference uses them) of the matchers whose first restatement is the Python of oracle/oracle.py:
//   ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)            reference src/ORBmatcher.cc:46-130
//   ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, ..) reference src/ORBmatcher.cc:1396-1535
//   ORBmatcher::ComputeThreeMaxima / DescriptorDistance                             reference src/ORBmatcher.cc:1666-1728
//   Frame::GetFeaturesInArea                                                        reference src/Frame.cc:730-779
// tests/test_match_oracle.py requires the two restatements to agree value for value.  Float expressions are written as
// the reference writes them; the library is built with -ffp-contract=off (declared: no FMA contraction).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "drfe_oracle.h"

namespace {

const int TH_HIGH = 100, HISTO_LENGTH = 30;

struct FrameView {   // what the matchers read of a Frame
  const drfe_frame_params* p;
  const drfe_keypoint* mvKeysUn;
  const float* mvuRight;
  const uint8_t* mDescriptors;
  int N;
  std::vector<std::vector<size_t>> mGrid;   // [64 * 48], x-major
  float mfGridElementWidthInv, mfGridElementHeightInv;
};

FrameView make_frame(const drfe_frame_params* p, const drfe_keypoint* ku, const float* ur, const uint8_t* desc, int n, const uint16_t* grid_count,
                     const uint16_t* grid_index) {
  FrameView F;
  F.p = p; F.mvKeysUn = ku; F.mvuRight = ur; F.mDescriptors = desc; F.N = n;
  F.mGrid.resize(DRFE_FRAME_GRID_COLS * DRFE_FRAME_GRID_ROWS);
  int o = 0;
  for (size_t c = 0; c < F.mGrid.size(); ++c)
    for (int k = 0; k < grid_count[c]; ++k) F.mGrid[c].push_back(grid_index[o++]);
  F.mfGridElementWidthInv = static_cast<float>(DRFE_FRAME_GRID_COLS) / static_cast<float>(p->max_x - p->min_x);   // Frame.cc:176-177
  F.mfGridElementHeightInv = static_cast<float>(DRFE_FRAME_GRID_ROWS) / static_cast<float>(p->max_y - p->min_y);
  return F;
}

// Frame::GetFeaturesInArea (Frame.cc:730-779)
std::vector<size_t> GetFeaturesInArea(const FrameView& F, const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) {
  std::vector<size_t> vIndices;
  const float mnMinX = F.p->min_x, mnMinY = F.p->min_y;
  const int nMinCellX = std::max(0, (int)floor((x - mnMinX - r) * F.mfGridElementWidthInv));
  if (nMinCellX >= DRFE_FRAME_GRID_COLS) return vIndices;
  const int nMaxCellX = std::min((int)DRFE_FRAME_GRID_COLS - 1, (int)ceil((x - mnMinX + r) * F.mfGridElementWidthInv));
  if (nMaxCellX < 0) return vIndices;
  const int nMinCellY = std::max(0, (int)floor((y - mnMinY - r) * F.mfGridElementHeightInv));
  if (nMinCellY >= DRFE_FRAME_GRID_ROWS) return vIndices;
  const int nMaxCellY = std::min((int)DRFE_FRAME_GRID_ROWS - 1, (int)ceil((y - mnMinY + r) * F.mfGridElementHeightInv));
  if (nMaxCellY < 0) return vIndices;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
      const std::vector<size_t>& vCell = F.mGrid[ix * DRFE_FRAME_GRID_ROWS + iy];
      for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
        const drfe_keypoint& kpUn = F.mvKeysUn[vCell[j]];
        if (bCheckLevels) {
          if (kpUn.octave < minLevel) continue;
          if (maxLevel >= 0 && kpUn.octave > maxLevel) continue;
        }
        const float distx = kpUn.x - x, disty = kpUn.y - y;
        if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
      }
    }
  return vIndices;
}

// ORBmatcher::DescriptorDistance (ORBmatcher.cc:1712-1728)
int DescriptorDistance(const uint8_t* a, const uint8_t* b) {
  int32_t pa[8], pb[8];
  memcpy(pa, a, 32); memcpy(pb, b, 32);
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    unsigned int v = pa[i] ^ pb[i];
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:1666-1707)
void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) ind3 = -1;
}

}  // namespace

extern "C" {

// SearchByProjection(CurrentFrame, LastFrame, th, bMono): holder[idx] models CurrentFrame.mvpMapPoints[idx] (-1 untouched,
// -2 set to NULL by the rotation check, else the last-frame index), observed[idx] = that map point's Observations() > 0
int orc_search_last_frame(const drfe_frame_params* p, const float* scale_factors, const drfe_keypoint* keys_un, const float* u_right, int n,
                          const uint16_t* grid_count, const uint16_t* grid_index, const uint8_t* desc, const float* Tcw, float th, int mode,
                          int check_orientation, const drfe_last_point* points, const uint8_t* pdesc, int npoints, const uint8_t* occupied,
                          int32_t* match_key, int32_t* match_dist, int32_t* holder) {
  FrameView CurrentFrame = make_frame(p, keys_un, u_right, desc, n, grid_count, grid_index);
  std::vector<char> observed(n, 0);
  for (int i = 0; i < n; ++i) { holder[i] = -1; observed[i] = occupied ? (occupied[i] != 0) : 0; }
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  const bool bForward = mode == 1, bBackward = mode == 2;
  const float fx = p->fx, fy = p->fy, cx = p->cx, cy = p->cy, mbf = p->bf;
  for (int i = 0; i < npoints; i++) {
    match_key[i] = -1; match_dist[i] = 256;
    if (!(points[i].flags & DRFE_LP_VALID)) continue;
    // cv::Mat x3Dc = Rcw*x3Dw+tcw: cv::gemm's 3x3 float path — products and sums in float, then + tcw through double
    float x3Dc[3];
    for (int r = 0; r < 3; ++r) {
      const float t0 = Tcw[4 * r] * points[i].X + Tcw[4 * r + 1] * points[i].Y + Tcw[4 * r + 2] * points[i].Z;
      x3Dc[r] = (float)((double)t0 + (double)Tcw[4 * r + 3]);
    }
    const float xc = x3Dc[0], yc = x3Dc[1];
    const float invzc = 1.0 / x3Dc[2];
    if (invzc < 0) continue;
    float u = fx * xc * invzc + cx;
    float v = fy * yc * invzc + cy;
    if (u < p->min_x || u > p->max_x) continue;
    if (v < p->min_y || v > p->max_y) continue;
    if (u != u || v != v) continue;   // declared: NaN image coordinates (z == 0 and x == 0) give no match
    int nLastOctave = points[i].octave;
    float radius = th * scale_factors[nLastOctave];
    std::vector<size_t> vIndices2;
    if (bForward) vIndices2 = GetFeaturesInArea(CurrentFrame, u, v, radius, nLastOctave);
    else if (bBackward) vIndices2 = GetFeaturesInArea(CurrentFrame, u, v, radius, 0, nLastOctave);
    else vIndices2 = GetFeaturesInArea(CurrentFrame, u, v, radius, nLastOctave - 1, nLastOctave + 1);
    if (vIndices2.empty()) continue;
    const uint8_t* dMP = pdesc + (size_t)i * 32;
    int bestDist = 256, bestIdx2 = -1;
    for (std::vector<size_t>::const_iterator vit = vIndices2.begin(), vend = vIndices2.end(); vit != vend; vit++) {
      const size_t i2 = *vit;
      if (observed[i2]) continue;
      if (CurrentFrame.mvuRight[i2] > 0) {
        const float ur = u - mbf * invzc;
        const float er = fabs(ur - CurrentFrame.mvuRight[i2]);
        if (er > radius) continue;
      }
      const int dist = DescriptorDistance(dMP, CurrentFrame.mDescriptors + i2 * 32);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    match_dist[i] = bestDist;
    if (bestDist <= TH_HIGH) {
      holder[bestIdx2] = i;
      observed[bestIdx2] = (points[i].flags & DRFE_LP_OBSERVED) != 0;
      match_key[i] = bestIdx2;
      nmatches++;
      if (check_orientation) {
        float rot = points[i].angle - CurrentFrame.mvKeysUn[bestIdx2].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(bestIdx2);
      }
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++)
      if (i != ind1 && i != ind2 && i != ind3)
        for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) { holder[rotHist[i][j]] = -2; nmatches--; }
  }
  return nmatches;
}

// SearchByProjection(F, vpMapPoints, th) whole; queries carry mTrackProjX/Y, r * mvScaleFactors[level], mTrackProjXR, levels
int orc_search_local_points(const drfe_frame_params* p, const drfe_keypoint* keys_un, const float* u_right, int n, const uint16_t* grid_count,
                            const uint16_t* grid_index, const uint8_t* desc, const drfe_proj_query* queries, const uint8_t* qdesc,
                            const uint8_t* qflags, int nq, float mfNNratio, const uint8_t* occupied, drfe_proj_match* out, int32_t* assigned,
                            int32_t* holder) {
  FrameView F = make_frame(p, keys_un, u_right, desc, n, grid_count, grid_index);
  std::vector<char> observed(n, 0);
  for (int i = 0; i < n; ++i) { holder[i] = -1; observed[i] = occupied ? (occupied[i] != 0) : 0; }
  int nmatches = 0;
  for (int iMP = 0; iMP < nq; iMP++) {
    drfe_proj_match& m = out[iMP];
    m.best_dist = 256; m.best_idx = -1; m.best_level = -1; m.best_dist2 = 256; m.best_level2 = -1;
    assigned[iMP] = -1;
    if (!(qflags[iMP] & DRFE_LP_VALID)) continue;
    const drfe_proj_query& q = queries[iMP];
    const std::vector<size_t> vIndices = GetFeaturesInArea(F, q.x, q.y, q.r, q.min_level, q.max_level);
    if (vIndices.empty()) continue;
    const uint8_t* MPdescriptor = qdesc + (size_t)iMP * 32;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (std::vector<size_t>::const_iterator vit = vIndices.begin(), vend = vIndices.end(); vit != vend; vit++) {
      const size_t idx = *vit;
      if (observed[idx]) continue;
      if (F.mvuRight[idx] > 0) {
        const float er = fabs(q.xr - F.mvuRight[idx]);
        if (er > q.r) continue;
      }
      const int dist = DescriptorDistance(MPdescriptor, F.mDescriptors + idx * 32);
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F.mvKeysUn[idx].octave; bestIdx = idx; }
      else if (dist < bestDist2) { bestLevel2 = F.mvKeysUn[idx].octave; bestDist2 = dist; }
    }
    m.best_dist = bestDist; m.best_idx = bestIdx; m.best_level = bestLevel; m.best_dist2 = bestDist2; m.best_level2 = bestLevel2;
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > mfNNratio * bestDist2) continue;
      holder[bestIdx] = iMP;
      observed[bestIdx] = (qflags[iMP] & DRFE_LP_OBSERVED) != 0;
      assigned[iMP] = bestIdx;
      nmatches++;
    }
  }
  return nmatches;
}

}  // extern "C"
