// CAPE / PlaneSeg / CylinderSeg / PlaneDetection_CAPE on the drfe C ABI — drop-in for the
// reference's src/CAPE/CAPE.h:19-53, PlaneSeg.h:15-35, CylinderSeg.h and
// include/PlaneExtractor.h:84-115 (src/PlaneExtractor.cpp:66-191).  Same constructor arguments,
// same process(...) argument meaning, same public result fields.  The work runs on the GPU
// through libdrfe.so; there is no CPU fallback (constructors throw if no device is usable).
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/drfe.h"
#ifdef DRFE_WITH_OPENCV
#include <opencv2/core/core.hpp>
#endif
#ifdef DRFE_WITH_EIGEN
#include <Eigen/Dense>
#endif
#include "drfe_compat.h"

// Public data members of the reference PlaneSeg (PlaneSeg.h:17-28); bool planar is stored as int
// in the ABI struct and converted here.
class PlaneSeg {
 public:
  int nr_pts = 0, min_nr_pts = 0;
  double x_acc = 0, y_acc = 0, z_acc = 0, xx_acc = 0, yy_acc = 0, zz_acc = 0, xy_acc = 0, xz_acc = 0, yz_acc = 0;
  float score = 0, MSE = 0;
  bool planar = false;
  double mean[3] = {0, 0, 0};
  double normal[3] = {0, 0, 0};
  double d = 0;
  PlaneSeg() = default;
  explicit PlaneSeg(const drfe_plane& p)
      : nr_pts(p.nr_pts), min_nr_pts(p.min_nr_pts), x_acc(p.x_acc), y_acc(p.y_acc), z_acc(p.z_acc), xx_acc(p.xx_acc),
        yy_acc(p.yy_acc), zz_acc(p.zz_acc), xy_acc(p.xy_acc), xz_acc(p.xz_acc), yz_acc(p.yz_acc), score(p.score),
        MSE(p.MSE), planar(p.planar != 0), d(p.d) {
    for (int i = 0; i < 3; ++i) { mean[i] = p.mean[i]; normal[i] = p.normal[i]; }
  }
};

// What CAPE::process hands back per cylinder (CAPE.cpp:434-445).
class CylinderSeg {
 public:
  std::vector<float> radii;
  std::vector<double> centers;   // 3 per segment
  double axis[3] = {0, 0, 0};
};

class CAPE {
 public:
#ifdef DRFE_WITH_OPENCV
  typedef cv::Mat SegImageT;
#else
  typedef drfe_compat::Mat8u SegImageT;
#endif
  CAPE(int depth_height, int depth_width, int cell_width, int cell_height, bool cylinder_detection,
       float min_cos_angle_4_merge = 0.97814f, float max_merge_dist = 900.f, int device = 0) {
    prm_.depth_height = depth_height; prm_.depth_width = depth_width; prm_.cell_width = cell_width; prm_.cell_height = cell_height;
    prm_.cylinder_detection = cylinder_detection ? 1 : 0;
    prm_.min_cos_angle_4_merge = min_cos_angle_4_merge; prm_.max_merge_dist = max_merge_dist;
    if (drfe_cape_create(&prm_, 1, device, &h_) != DRFE_OK) throw std::runtime_error(std::string("CAPE: ") + drfe_last_error());
    seg_.resize((size_t)depth_height * depth_width);
    planes_.resize(kPlaneCap);
    if (cylinder_detection) cyls_.resize((size_t)(depth_height / cell_height) * (depth_width / cell_width) / 6 + 1);
  }
  ~CAPE() { drfe_cape_destroy(h_); }
  CAPE(const CAPE&) = delete;
  CAPE& operator=(const CAPE&) = delete;

  // cloud_array: cell-major organised cloud, N x 3 column-major (what organizePointCloudByCell
  // produces, PlaneExtractor.cpp:80-99).  Works with Eigen::MatrixXf or drfe_compat::MatrixXf.
  // Like the reference, only labelled pixels of seg_output are written (the caller zeroes it,
  // PlaneExtractor.cpp:129) and plane_segments_final is appended to (CAPE.cpp:279).
  template <class MatrixT>
  void process(MatrixT& cloud_array, int& nr_planes, int& nr_cylinders, SegImageT& seg_output,
               std::vector<PlaneSeg>& plane_segments_final, std::vector<CylinderSeg>& cylinder_segments_final) {
    if (drfe_cape_process(h_, cloud_array.data(), seg_.data(), planes_.data(), kPlaneCap, &nr_planes, cyls_.data(), (int)cyls_.size(),
                          &nr_cylinders) != DRFE_OK)
      throw std::runtime_error(std::string("CAPE::process: ") + drfe_last_error());
    finish(nr_planes, seg_output, plane_segments_final);
    finish_cylinders(cylinder_segments_final);
  }
  // fused PlaneDetection_CAPE path: depth image -> cloud (double math) -> cell-major -> process
  void processDepth(const float* depth, size_t row_stride_elems, float fx, float fy, float cx, float cy, int& nr_planes,
                    int& nr_cylinders, SegImageT& seg_output, std::vector<PlaneSeg>& plane_segments_final,
                    std::vector<CylinderSeg>* cylinder_segments_final = nullptr) {
    if (drfe_cape_process_depth(h_, depth, row_stride_elems, fx, fy, cx, cy, seg_.data(), planes_.data(), kPlaneCap, &nr_planes,
                                cyls_.data(), (int)cyls_.size(), &nr_cylinders) != DRFE_OK)
      throw std::runtime_error(std::string("CAPE::processDepth: ") + drfe_last_error());
    finish(nr_planes, seg_output, plane_segments_final);
    if (cylinder_segments_final) finish_cylinders(*cylinder_segments_final);
  }
  // plane_cloud of PlaneDetection_CAPE (PlaneExtractor.cpp:165-190) for the frame just processed: xyz of plane p are
  // points[3*offsets[p]] .. points[3*offsets[p+1]), gathered on the device (drfe_cape_plane_points)
  void planePoints(std::vector<float>& points, std::vector<int>& offsets) {
    const size_t N = (size_t)prm_.depth_height * prm_.depth_width;
    points.resize(N * 3);
    offsets.assign(kPlaneCap + 1, 0);
    if (drfe_cape_plane_points(h_, points.data(), N, offsets.data(), kPlaneCap) != DRFE_OK)
      throw std::runtime_error(std::string("CAPE::planePoints: ") + drfe_last_error());
  }
  drfe_cape* handle() { return h_; }

 private:
  static const int kPlaneCap = 255;
  void finish(int nr_planes, SegImageT& seg_output, std::vector<PlaneSeg>& out) {
    for (int i = 0; i < nr_planes; ++i) out.push_back(PlaneSeg(planes_[i]));
    const int H = prm_.depth_height, W = prm_.depth_width;
    for (int r = 0; r < H; ++r) {
      uint8_t* dst = seg_output.ptr(r);
      const uint8_t* src = seg_.data() + (size_t)r * W;
      for (int c = 0; c < W; ++c)
        if (src[c] > 0) dst[c] = src[c];                           // CAPE.cpp:423-425
    }
  }
  // cylinder_segments_final: one CylinderSeg per cylinder found, erosion survivors or not (CAPE.cpp:434-445)
  void finish_cylinders(std::vector<CylinderSeg>& out) {
    int found = 0;
    if (cyls_.empty() || drfe_cape_cylinders_found(h_, &found) != DRFE_OK) return;
    for (int i = 0; i < found && i < (int)cyls_.size(); ++i) {
      CylinderSeg cy;
      cy.radii.push_back(cyls_[i].radius);
      for (int k = 0; k < 3; ++k) { cy.centers.push_back(cyls_[i].center[k]); cy.axis[k] = cyls_[i].axis[k]; }
      out.push_back(cy);
    }
  }
  drfe_cape_params prm_{};
  drfe_cape* h_ = nullptr;
  std::vector<uint8_t> seg_;
  std::vector<drfe_plane> planes_;
  std::vector<drfe_cylinder> cyls_;
};

namespace Planar_SLAM {

// PlaneDetection_CAPE (PlaneExtractor.h:84-115): readColorImage / readDepthImage / runPlaneDetection and the public
// result fields.  plane_cloud (PCL) is replaced by per-plane xyz lists so that no PCL is needed.  With
// -DDRFE_WITH_OPENCV the members are the reference's cv::Mat's (seg_output is CV_8U like its cv::Mat_<uchar>,
// K_ the 3x3 CV_32F camera matrix); otherwise the stand-ins of drfe_compat.h.
class PlaneDetection_CAPE {
 public:
  struct PointT { float x, y, z; };
  typedef std::vector<PointT> PointCloud;

  PlaneDetection_CAPE() = default;
  ~PlaneDetection_CAPE() { delete plane_detector; }
  PlaneDetection_CAPE(const PlaneDetection_CAPE&) = delete;
  PlaneDetection_CAPE& operator=(const PlaneDetection_CAPE&) = delete;

#ifdef DRFE_WITH_OPENCV
  bool readColorImage(cv::Mat RGBImg) {                            // PlaneExtractor.cpp:71-78 (the colour image is not used by the path)
    color_img_ = RGBImg;
    return !(color_img_.empty() || color_img_.depth() != CV_8U);
  }
  bool readDepthImage(cv::Mat depthImg, cv::Mat& K) {              // PlaneExtractor.cpp:101-109
    depth_img = depthImg;
    K_ = K;
    return !(depth_img.empty() || depth_img.depth() != CV_32F);
  }
  cv::Mat seg_output, color_img_, depth_img, K_;
#else
  bool readDepthImage(const drfe_compat::Mat32f& depthImg, const float K[9]) {
    depth_img = depthImg;
    for (int i = 0; i < 9; ++i) K_[i] = K[i];
    return !depth_img.empty();
  }
  drfe_compat::Mat8u seg_output;
  drfe_compat::Mat32f depth_img;
  float K_[9] = {0};
#endif

  void runPlaneDetection() {
    const int rows = depth_img.rows, cols = depth_img.cols;
#ifdef DRFE_WITH_OPENCV
    seg_output.create(rows, cols, CV_8U);                          // cv::Mat_<uchar>(rows, cols, uchar(0)), PlaneExtractor.cpp:129
    seg_output.setTo(cv::Scalar(0));
    const float* depth = depth_img.ptr<float>(0);
    const float fx = K_.at<float>(0, 0), fy = K_.at<float>(1, 1), cx = K_.at<float>(0, 2), cy = K_.at<float>(1, 2);   // :113-116
#else
    seg_output.create(rows, cols);
    seg_output.setTo(0);
    const float* depth = depth_img.data;
    const float fx = K_[0], fy = K_[4], cx = K_[2], cy = K_[5];
#endif
    // CAPE::process appends (CAPE.cpp:279) and the reference uses one PlaneDetection_CAPE per Frame; this object may
    // be run again, so its results start empty: plane_params[i], plane_cloud[i] and label i + 1 of seg_output line up
    plane_params.clear();
    cylinder_params.clear();
    if (!plane_detector || rows != det_rows_ || cols != det_cols_) {   // the reference news one per frame and leaks it (:149)
      delete plane_detector;
      plane_detector = new CAPE(rows, cols, PATCH_SIZE, PATCH_SIZE, cylinder_detection, COS_ANGLE_MAX, MAX_MERGE_DIST);
      det_rows_ = rows; det_cols_ = cols;
    }
    plane_detector->processDepth(depth, (size_t)depth_img.step / sizeof(float), fx, fy, cx, cy, nr_planes, nr_cylinders, seg_output,
                                 plane_params, &cylinder_params);
    // per-plane point lists (PlaneExtractor.cpp:165-190), gathered on the device from seg_output and the cloud
    plane_detector->planePoints(pts_, offs_);
    plane_cloud.assign(nr_planes, PointCloud());
    for (int p = 0; p < nr_planes; ++p) {
      const PointT* first = reinterpret_cast<const PointT*>(pts_.data()) + offs_[p];
      plane_cloud[p].assign(first, first + (offs_[p + 1] - offs_[p]));
    }
  }

  // What Frame::ComputePlanes_CAPE computes next on the depth image (Frame.cc:1153-1216): the 1/3-resolution cloud (far points zeroed)
  // and PCL's integral-image surface normals on it (AVERAGE_3D_GRADIENT, depth-change factor 0.05, smoothing 10).  cloud3 / normals3:
  // [h3][w3][3] floats, inputCloud->at(n, m) = cloud3[(m * w3 + n) * 3 ..], NaN where PCL leaves NaN.  Call after runPlaneDetection().
  void thirdCloudNormals(float max_point_dist, std::vector<float>& cloud3, std::vector<float>& normals3, int& w3, int& h3) {
    if (!plane_detector) throw std::runtime_error("PlaneDetection_CAPE::thirdCloudNormals: runPlaneDetection first");
    w3 = (depth_img.cols + 2) / 3; h3 = (depth_img.rows + 2) / 3;
    cloud3.resize((size_t)w3 * h3 * 3); normals3.resize((size_t)w3 * h3 * 3);
    if (drfe_cape_third_cloud_normals(plane_detector->handle(), max_point_dist, 0.05f, 10.0f, cloud3.data(), normals3.data()) != DRFE_OK)
      throw std::runtime_error(std::string("drfe_cape_third_cloud_normals: ") + drfe_last_error());
  }

  std::vector<PointCloud> plane_cloud;
  std::vector<PlaneSeg> plane_params;
  std::vector<CylinderSeg> cylinder_params;
  int nr_planes = 0, nr_cylinders = 0;
  int PATCH_SIZE = 20;
  float COS_ANGLE_MAX = (float)std::cos(M_PI / 12);
  float MAX_MERGE_DIST = 50.f;
  bool cylinder_detection = false;
  CAPE* plane_detector = nullptr;

 private:
  int det_rows_ = 0, det_cols_ = 0;
  std::vector<float> pts_;
  std::vector<int> offs_;
};

}  // namespace Planar_SLAM
