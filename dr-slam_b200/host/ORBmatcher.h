// The three per-frame matchers of Planar_SLAM::ORBmatcher (reference include/ORBmatcher.h, src/ORBmatcher.cc) that run whole
// on the device, on the drfe C ABI: SearchByProjection(F, vpMapPoints, th) (:46-130), SearchByProjection(CurrentFrame,
// LastFrame, th, bMono) (:1396-1535) and SearchByBoW(pKF, F, vpMapPointMatches) (:160-292).  The current frame is the extractor's last call (keypoints, descriptors, grid on the
// device after FramePost); map points are referred to by their index in the other frame, the caller maps indices back to
// MapPoint* (INTEGRATION.md).
#pragma once
#include "ORBVocabulary.h"

namespace Planar_SLAM {

// what Frame::Frame does with the keypoints right after ExtractORB (Frame.cc:160-180): UndistortKeyPoints,
// ComputeStereoFromRGBD, AssignFeaturesToGrid — leaves mvKeysUn / mvuRight / mGrid on the device for the matchers
inline void FramePost(ORBextractor& ex, const drfe_frame_params& fp, const float* depth, int width, int height,
                      std::vector<ORBextractor::KeyPointT>* mvKeysUn = nullptr, std::vector<float>* mvuRight = nullptr,
                      std::vector<float>* mvDepth = nullptr) {
  const int cap = ex.max_keypoints();
  if (mvKeysUn) mvKeysUn->resize(cap);
  if (mvuRight) mvuRight->resize(cap);
  if (mvDepth) mvDepth->resize(cap);
  if (drfe_orb_frame_post(ex.handle(), &fp, depth, (size_t)width, (size_t)width * height, DRFE_MEM_HOST,
                          mvKeysUn ? reinterpret_cast<drfe_keypoint*>(mvKeysUn->data()) : nullptr, mvuRight ? mvuRight->data() : nullptr,
                          mvDepth ? mvDepth->data() : nullptr, nullptr, nullptr, cap) != DRFE_OK)
    throw std::runtime_error(std::string("FramePost: ") + drfe_last_error());
}

class ORBmatcher {
 public:
  static const int TH_LOW = DRFE_TH_LOW, TH_HIGH = DRFE_TH_HIGH, HISTO_LENGTH = DRFE_HISTO_LENGTH;   // ORBmatcher.cc:38-40
  ORBmatcher(float nnratio = 0.6f, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

  // SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono).  Tcw: rows 0..2 of CurrentFrame.mTcw; mode
  // 0 / 1 (bForward) / 2 (bBackward) from :1406-1414; last[i] = position, angle, octave and flags of LastFrame's keypoint i,
  // lastDesc = its map point's descriptor.  On return mvpMapPoints[idx] = the index i whose map point
  // CurrentFrame.mvpMapPoints[idx] holds, -1 = untouched, -2 = NULLed by the rotation check.
  int SearchByProjection(ORBextractor& cur, const float Tcw[12], const std::vector<drfe_last_point>& last, const std::vector<uint8_t>& lastDesc,
                         float th, int mode, std::vector<int32_t>& mvpMapPoints) {
    drfe_track_params tp{};
    std::memcpy(tp.Tcw, Tcw, sizeof(tp.Tcw));
    tp.th = th; tp.mode = mode; tp.check_orientation = mbCheckOrientation;
    const int n = (int)last.size();
    mvpMapPoints.assign(cur.max_keypoints(), -1);
    int nmatches = 0;
    if (n == 0) return 0;
    if (drfe_orb_search_last_frame(cur.handle(), &tp, &n, last.data(), lastDesc.data(), nullptr, n, nullptr, nullptr, mvpMapPoints.data(), &nmatches,
                                   nullptr) != DRFE_OK)
      throw std::runtime_error(std::string("ORBmatcher::SearchByProjection: ") + drfe_last_error());
    return nmatches;
  }

  // SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th) (:46-130), the matcher of SearchLocalPoints:
  // queries[i] = {mTrackProjX, mTrackProjY, r * F.mvScaleFactors[level] (r = RadiusByViewingCos(mTrackViewCos) [* th]),
  // mTrackProjXR, level - 1, level}, flags[i] = DRFE_LP_VALID (mbTrackInView && !isBad()) | DRFE_LP_OBSERVED, descs = the map
  // points' descriptors, occupied[idx] != 0 where F.mvpMapPoints[idx] already holds an observed point.  On return
  // mvpMapPoints[idx] = index of the map point assigned to keypoint idx, or -1 (untouched).
  int SearchByProjection(ORBextractor& F, const std::vector<drfe_proj_query>& queries, const std::vector<uint8_t>& descs,
                         const std::vector<uint8_t>& flags, const std::vector<uint8_t>& occupied, std::vector<int32_t>& mvpMapPoints) {
    const int n = (int)queries.size();
    mvpMapPoints.assign(F.max_keypoints(), -1);
    int nmatches = 0;
    if (n == 0) return 0;
    if (drfe_orb_search_local_points(F.handle(), &n, queries.data(), descs.data(), flags.data(), occupied.empty() ? nullptr : occupied.data(), n,
                                     mfNNratio, nullptr, nullptr, mvpMapPoints.data(), &nmatches) != DRFE_OK)
      throw std::runtime_error(std::string("ORBmatcher::SearchByProjection: ") + drfe_last_error());
    return nmatches;
  }

  // SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches): kfDesc / kfAngle / kfGood per keyframe
  // feature (pKF->mDescriptors, mvKeysUn[i].angle, map point present and not bad), the two FeatureVectors; on return
  // vpMapPointMatches[idx] = keyframe feature index or -1
  int SearchByBoW(const std::vector<uint8_t>& kfDesc, const std::vector<float>& kfAngle, const std::vector<uint8_t>& kfGood,
                  const DBoW2::FeatureVector& kfFeatVec, ORBextractor& F, const DBoW2::FeatureVector& fFeatVec, std::vector<int32_t>& vpMapPointMatches) {
    const int nk = (int)kfAngle.size(), cap = F.max_keypoints(), kcap = nk > 0 ? nk : 1;
    std::vector<int32_t> knode(kcap), kstart(kcap + 1), kfeat(kcap), fnode(cap), fstart(cap + 1), ffeat(cap);
    const int kn = pack(kfFeatVec, knode, kstart, kfeat), fn = pack(fFeatVec, fnode, fstart, ffeat);
    vpMapPointMatches.assign(cap, -1);
    int nmatches = 0;
    if (nk == 0) return 0;
    if (drfe_orb_search_by_bow(F.handle(), kcap, &nk, kfDesc.data(), kfAngle.data(), kfGood.data(), &kn, knode.data(), kstart.data(), kfeat.data(), &fn,
                               fnode.data(), fstart.data(), ffeat.data(), mfNNratio, mbCheckOrientation, nullptr, vpMapPointMatches.data(),
                               &nmatches) != DRFE_OK)
      throw std::runtime_error(std::string("ORBmatcher::SearchByBoW: ") + drfe_last_error());
    return nmatches;
  }

 protected:
  static int pack(const DBoW2::FeatureVector& fv, std::vector<int32_t>& node, std::vector<int32_t>& start, std::vector<int32_t>& feat) {
    int j = 0, o = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end() && j < (int)node.size(); ++it, ++j) {
      node[j] = (int32_t)it->first; start[j] = o;
      for (size_t t = 0; t < it->second.size() && o < (int)feat.size(); ++t) feat[o++] = (int32_t)it->second[t];
    }
    start[j] = o;
    return j;
  }
  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace Planar_SLAM
