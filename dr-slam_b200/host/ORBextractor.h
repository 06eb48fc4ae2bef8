// Planar_SLAM::ORBextractor on the drfe C ABI — drop-in for reference include/ORBextractor.h
// (class at :44-116) / src/ORBextractor.cc.  Same constructor, same operator(), same getters,
// same public mvImagePyramid; the work runs on the GPU through libdrfe.so (no CPU fallback:
// construction throws std::runtime_error if no CUDA device is usable).
//
// Differences a maintainer should know:
//  * the device handle is created on the first operator() call, when the image size is known
//    (the reference constructor does not take a size); a different size re-creates it;
//  * mvImagePyramid is only filled when keep_pyramid is set (nothing in DR-SLAM reads it).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/drfe.h"
#ifdef DRFE_WITH_OPENCV
#include <opencv2/core/core.hpp>
#else
#include "drfe_compat.h"
#endif

namespace Planar_SLAM {

class ORBextractor {
 public:
#ifdef DRFE_WITH_OPENCV
  typedef cv::KeyPoint KeyPointT;
#else
  typedef drfe_compat::KeyPoint KeyPointT;
  typedef drfe_compat::Mat8u ImageT;
#endif
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0)
      : device_(device) {
    prm_.nfeatures = nfeatures; prm_.scale_factor = scaleFactor; prm_.nlevels = nlevels;
    prm_.ini_th_fast = iniThFAST; prm_.min_th_fast = minThFAST;
    // the tables of ORBextractor::ORBextractor (ORBextractor.cc:415-431), exactly as the reference computes them
    mvScaleFactor.assign(nlevels, 1.f); mvLevelSigma2.assign(nlevels, 1.f);
    const double sf = (double)scaleFactor;
    for (int i = 1; i < nlevels; ++i) {
      mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * sf);
      mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; ++i) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
    int n = 0;
    if (drfe_device_count(&n) != DRFE_OK || n == 0)
      throw std::runtime_error(std::string("ORBextractor: no CUDA device (") + drfe_last_error() + ")");
  }
  ~ORBextractor() { drfe_orb_destroy(h_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

#ifdef DRFE_WITH_OPENCV
  // Mask is ignored in the reference too (ORBextractor.h:58).
  void operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray _descriptors) {
    if (_image.empty()) return;                                   // ORBextractor.cc:1046
    cv::Mat image = _image.getMat();
    CV_Assert(image.type() == CV_8UC1);                           // :1050
    int n = 0;
    run(image.data, image.cols, image.rows, image.step, n);
    keypoints.clear();
    if (n == 0) { _descriptors.release(); return; }               // :1064-1065
    _descriptors.create(n, 32, CV_8U);
    cv::Mat d = _descriptors.getMat();
    keypoints.resize(n);
    std::memcpy(keypoints.data(), kps_.data(), (size_t)n * sizeof(cv::KeyPoint));
    for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), desc_.data() + (size_t)i * 32, 32);
    if (keep_pyramid) fetch_pyramid();
  }
  std::vector<cv::Mat> mvImagePyramid;
#else
  void operator()(const ImageT& image, const ImageT& /*mask*/, std::vector<KeyPointT>& keypoints, ImageT& descriptors) {
    if (image.empty()) return;                                    // ORBextractor.cc:1046
    int n = 0;
    run(image.data, image.cols, image.rows, image.step, n);
    keypoints.clear();
    if (n == 0) { descriptors.release(); return; }                // :1064-1065
    descriptors.create(n, 32);
    keypoints.resize(n);
    std::memcpy(keypoints.data(), kps_.data(), (size_t)n * sizeof(KeyPointT));
    std::memcpy(descriptors.data, desc_.data(), (size_t)n * 32);
    if (keep_pyramid) fetch_pyramid();
  }
  std::vector<ImageT> mvImagePyramid;
#endif

  int GetLevels() { return prm_.nlevels; }
  float GetScaleFactor() { return prm_.scale_factor; }
  std::vector<float> GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  bool keep_pyramid = false;
  // the device handle (keypoints, descriptors and pyramid of the last call stay on the device): what ORBVocabulary.h and
  // ORBmatcher.h run on
  drfe_orb* handle() const { return h_; }
  int max_keypoints() const { return cap_; }

 protected:
  void run(const uint8_t* data, int w, int h, size_t step, int& n) {
    if (!h_ || w != w_ || h != hgt_) {
      drfe_orb_destroy(h_);
      h_ = nullptr;
      if (drfe_orb_create(&prm_, w, h, 1, device_, &h_) != DRFE_OK)
        throw std::runtime_error(std::string("ORBextractor: ") + drfe_last_error());
      w_ = w; hgt_ = h;
      cap_ = drfe_orb_max_keypoints(h_);
      kps_.resize(cap_); desc_.resize((size_t)cap_ * 32);
    }
    if (drfe_orb_extract(h_, data, w, h, step, kps_.data(), desc_.data(), cap_, &n) != DRFE_OK)
      throw std::runtime_error(std::string("ORBextractor: ") + drfe_last_error());
  }
  void fetch_pyramid() {
    mvImagePyramid.resize(prm_.nlevels);
    for (int l = 0; l < prm_.nlevels; ++l) {
      int lw = 0, lh = 0;
      drfe_orb_level_size(h_, l, &lw, &lh);
#ifdef DRFE_WITH_OPENCV
      mvImagePyramid[l].create(lh, lw, CV_8U);
#else
      mvImagePyramid[l].create(lh, lw);
#endif
      drfe_orb_get_pyramid(h_, 0, l, 0, mvImagePyramid[l].data);
    }
  }

  drfe_orb_params prm_{};
  drfe_orb* h_ = nullptr;
  int device_ = 0, w_ = 0, hgt_ = 0, cap_ = 0;
  std::vector<drfe_keypoint> kps_;
  std::vector<uint8_t> desc_;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

}  // namespace Planar_SLAM
