// ORACLE — test infrastructure only (never linked into libdrfe.so).  Second, independent restatement (C++, std::vector /
// std::map as the reference uses them) of the matchers whose first restatement is the Python of oracle/oracle.py:
//   ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)            reference src/ORBmatcher.cc:46-130
//   ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, ..) reference src/ORBmatcher.cc:1396-1535
//   ORBmatcher::ComputeThreeMaxima / DescriptorDistance                             reference src/ORBmatcher.cc:1666-1728
//   Frame::GetFeaturesInArea                                                        reference src/Frame.cc:730-779
//   (further down) DBoW2 transform with BowVector / FeatureVector, ORBmatcher::SearchByBoW
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

// ---------------------------------------------------------------- DBoW2 transform and SearchByBoW, with the reference's std::maps
//   TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)  Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1258
//   BowVector::addWeight / addIfNotExist / normalize                                Thirdparty/DBoW2/DBoW2/BowVector.cpp:34-84
//   FeatureVector::addFeature                                                       Thirdparty/DBoW2/DBoW2/FeatureVector.cpp:31-45
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)                  src/ORBmatcher.cc:160-292
#include <map>

namespace {

struct Node {   // TemplatedVocabulary::Node (TemplatedVocabulary.h:297-330)
  unsigned id = 0;
  double weight = 0;
  std::vector<unsigned> children;
  unsigned parent = 0;
  uint8_t descriptor[32] = {0};
  unsigned word_id = 0;
  bool isLeaf() const { return children.empty(); }
};

typedef std::map<unsigned, double> BowVector;
typedef std::map<unsigned, std::vector<unsigned>> FeatureVector;

void addWeight(BowVector& v, unsigned id, double w) {   // BowVector.cpp:34-46
  BowVector::iterator vit = v.lower_bound(id);
  if (vit != v.end() && !(v.key_comp()(id, vit->first))) vit->second += w;
  else v.insert(vit, BowVector::value_type(id, w));
}
void addIfNotExist(BowVector& v, unsigned id, double w) {   // :50-58
  BowVector::iterator vit = v.lower_bound(id);
  if (vit == v.end() || (v.key_comp()(id, vit->first))) v.insert(vit, BowVector::value_type(id, w));
}
void addFeature(FeatureVector& fv, unsigned id, unsigned i_feature) {   // FeatureVector.cpp:31-45
  FeatureVector::iterator vit = fv.lower_bound(id);
  if (vit != fv.end() && vit->first == id) vit->second.push_back(i_feature);
  else { vit = fv.insert(vit, FeatureVector::value_type(id, std::vector<unsigned>())); vit->second.push_back(i_feature); }
}

}  // namespace

extern "C" {

// returns the BowVector size; bow_key / bow_value and fv_node / fv_start / fv_feat receive the maps in iteration order
int orc_bow_transform(int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* descriptors,
                      const double* weights, const uint8_t* features, int nfeatures, int levelsup, int32_t* word_of, int32_t* node_of,
                      int32_t* bow_key, double* bow_value, int* fv_n, int32_t* fv_node, int32_t* fv_start, int32_t* fv_feat) {
  // loadFromTextFile (:1376-1418)
  std::vector<Node> m_nodes(1);
  unsigned nwords = 0;
  for (int i = 0; i < nnodes; ++i) {
    const unsigned nid = m_nodes.size();
    m_nodes.resize(m_nodes.size() + 1);
    m_nodes[nid].id = nid;
    m_nodes[nid].parent = parent[i];
    m_nodes[parent[i]].children.push_back(nid);
    memcpy(m_nodes[nid].descriptor, descriptors + (size_t)i * 32, 32);
    m_nodes[nid].weight = weights[i];
    if (is_leaf[i] > 0) m_nodes[nid].word_id = nwords++;
  }
  BowVector v;
  FeatureVector fv;
  const bool must = scoring != 5;               // ScoringObject.h:73-89
  const bool l2 = scoring == 1;
  for (int i_feature = 0; i_feature < nfeatures; ++i_feature) {
    const uint8_t* feature = features + (size_t)i_feature * 32;
    // transform(feature, word_id, weight, nid, levelsup) (:1217-1258)
    const int nid_level = L - levelsup;
    unsigned nid = 0;
    bool nid_set = nid_level <= 0;
    unsigned final_id = 0;
    int current_level = 0;
    do {
      ++current_level;
      const std::vector<unsigned>& nodes = m_nodes[final_id].children;
      final_id = nodes[0];
      double best_d = DescriptorDistance(feature, m_nodes[final_id].descriptor);
      for (std::vector<unsigned>::const_iterator nit = nodes.begin() + 1; nit != nodes.end(); ++nit) {
        const unsigned id = *nit;
        const double d = DescriptorDistance(feature, m_nodes[id].descriptor);
        if (d < best_d) { best_d = d; final_id = id; }
      }
      if (current_level == nid_level) { nid = final_id; nid_set = true; }
    } while (!m_nodes[final_id].isLeaf());
    if (!nid_set) nid = final_id;               // declared (the reference leaves *nid uninitialised)
    const unsigned id = m_nodes[final_id].word_id;
    const double w = m_nodes[final_id].weight;
    word_of[i_feature] = -1; node_of[i_feature] = -1;
    if (w > 0) {
      if (weighting <= 1) addWeight(v, id, w); else addIfNotExist(v, id, w);
      addFeature(fv, nid, i_feature);
      word_of[i_feature] = (int32_t)id; node_of[i_feature] = (int32_t)nid;
    }
  }
  if (weighting <= 1 && !v.empty() && !must) {  // :1164-1170
    const double nd = v.size();
    for (BowVector::iterator vit = v.begin(); vit != v.end(); vit++) vit->second /= nd;
  }
  if (must) {                                   // BowVector::normalize (:62-84)
    double norm = 0.0;
    if (!l2) { for (BowVector::iterator it = v.begin(); it != v.end(); ++it) norm += fabs(it->second); }
    else { for (BowVector::iterator it = v.begin(); it != v.end(); ++it) norm += it->second * it->second; norm = sqrt(norm); }
    if (norm > 0.0) for (BowVector::iterator it = v.begin(); it != v.end(); ++it) it->second /= norm;
  }
  int j = 0;
  for (BowVector::const_iterator it = v.begin(); it != v.end(); ++it, ++j) { bow_key[j] = (int32_t)it->first; bow_value[j] = it->second; }
  int k = 0, o = 0;
  for (FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++k) {
    fv_node[k] = (int32_t)it->first; fv_start[k] = o;
    for (size_t t = 0; t < it->second.size(); ++t) fv_feat[o++] = (int32_t)it->second[t];
  }
  fv_start[k] = o;
  *fv_n = k;
  return j;
}

// SearchByBoW: FeatureVectors given flat (node, start, feat); f_match models vpMapPointMatches (keyframe feature index or -1)
int orc_search_by_bow(const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid, int nk, int kf_fv_n, const int32_t* kf_fv_node,
                      const int32_t* kf_fv_start, const int32_t* kf_fv_feat, const uint8_t* f_desc, const float* f_angle, int nf, int f_fv_n,
                      const int32_t* f_fv_node, const int32_t* f_fv_start, const int32_t* f_fv_feat, float mfNNratio, int check_orientation,
                      int32_t* kf_match, int32_t* f_match) {
  const int TH_LOW = 50;
  FeatureVector vFeatVecKF, FFeatVec;
  for (int j = 0; j < kf_fv_n; ++j) vFeatVecKF[kf_fv_node[j]] = std::vector<unsigned>(kf_fv_feat + kf_fv_start[j], kf_fv_feat + kf_fv_start[j + 1]);
  for (int j = 0; j < f_fv_n; ++j) FFeatVec[f_fv_node[j]] = std::vector<unsigned>(f_fv_feat + f_fv_start[j], f_fv_feat + f_fv_start[j + 1]);
  for (int i = 0; i < nf; ++i) f_match[i] = -1;
  for (int i = 0; i < nk; ++i) kf_match[i] = -1;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  FeatureVector::const_iterator KFit = vFeatVecKF.begin(), Fit = FFeatVec.begin(), KFend = vFeatVecKF.end(), Fend = FFeatVec.end();
  while (KFit != KFend && Fit != Fend) {
    if (KFit->first == Fit->first) {
      const std::vector<unsigned> vIndicesKF = KFit->second, vIndicesF = Fit->second;
      for (size_t iKF = 0; iKF < vIndicesKF.size(); iKF++) {
        const unsigned realIdxKF = vIndicesKF[iKF];
        if (!kf_valid[realIdxKF]) continue;
        const uint8_t* dKF = kf_desc + (size_t)realIdxKF * 32;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (size_t iF = 0; iF < vIndicesF.size(); iF++) {
          const unsigned realIdxF = vIndicesF[iF];
          if (f_match[realIdxF] >= 0) continue;
          const int dist = DescriptorDistance(dKF, f_desc + (size_t)realIdxF * 32);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 <= TH_LOW) {
          if (static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {
            f_match[bestIdxF] = realIdxKF;
            kf_match[realIdxKF] = bestIdxF;
            if (check_orientation) {
              float rot = kf_angle[realIdxKF] - f_angle[bestIdxF];
              if (rot < 0.0) rot += 360.0f;
              int bin = round(rot * factor);
              if (bin == HISTO_LENGTH) bin = 0;
              rotHist[bin].push_back(bestIdxF);
            }
            nmatches++;
          }
        }
      }
      KFit++; Fit++;
    } else if (KFit->first < Fit->first) KFit = vFeatVecKF.lower_bound(Fit->first);
    else Fit = FFeatVec.lower_bound(KFit->first);
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) { f_match[rotHist[i][j]] = -1; nmatches--; }
    }
  }
  return nmatches;
}

}  // extern "C"
