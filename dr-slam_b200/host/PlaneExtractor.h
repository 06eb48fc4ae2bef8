// Planar_SLAM::PlaneDetection on the drfe C ABI — drop-in for the reference's include/PlaneExtractor.h:40-81 and
// src/PlaneExtractor.cpp:7-63, the plane extractor Frame::Frame starts on its own thread (Frame.cc:126, ComputePlanes
// :937-1040): readDepthImage(Depth, K, depthFactor) + runPlaneDetection() = ahc::PlaneFitter<ImagePointCloud>::run.
// Same member names and meaning for everything Frame::ComputePlanes reads:
//   plane_num_, plane_vertices_[i] (pixel indices, scan order), cloud.vertices[j] for the pixels of those lists,
//   plane_filter.extractedPlanes[i]->normal / center (and mse, N, curvature), seg_output.
// The work runs on the GPU through libdrfe.so (drfe_peac_*); there is no CPU fallback (the constructor throws if no device is
// usable).  cloud.vertices holds the vertices of the plane members only — (double) of the float the reference's own
// (float) cloud.vertices[j][k] (Frame.cc:960-962) gives, so that cast comes out the same; all other entries are 0.
#pragma once
#include <array>
#include <memory>
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

namespace ahc {
// the public fields of ahc::PlaneSeg (include/peac/AHCPlaneSeg.hpp:57-120) that DR-SLAM reads
struct PlaneSeg {
  int rid = 0, N = 0;
  double mse = 0, center[3] = {0, 0, 0}, normal[3] = {0, 0, 0}, curvature = 0;
  typedef std::shared_ptr<PlaneSeg> shared_ptr;
};
}  // namespace ahc

namespace Planar_SLAM {
#ifdef DRFE_WITH_EIGEN
typedef Eigen::Vector3d VertexType;
#else
typedef std::array<double, 3> VertexType;
#endif

const int kDepthWidth = 640;
const int kDepthHeight = 480;

// the organized cloud as the reference exposes it (include/PlaneExtractor.h:40-58): row-major vertices, get() refuses points without depth
struct ImagePointCloud {
  std::vector<VertexType> vertices;
  int w = 0, h = 0;
  int width() const { return w; }
  int height() const { return h; }
  bool get(const int row, const int col, double& x, double& y, double& z) const {
    const VertexType& v = vertices[(size_t)row * w + col];
    z = v[2];
    const bool has_depth = z != 0 && z == z;                       // 0 or NaN: not a point
    if (has_depth) { x = v[0]; y = v[1]; }
    return has_depth;
  }
};

class PlaneDetection {
 public:
  // what is left of ahc::PlaneFitter on the host: its parameters before the first run, its extractedPlanes after one
  struct Fitter {
    drfe_peac_params params;
    std::vector<ahc::PlaneSeg::shared_ptr> extractedPlanes;
    Fitter() { drfe_peac_default_params(&params); }
  };

  ImagePointCloud cloud;
  Fitter plane_filter;
  std::vector<std::vector<int>> plane_vertices_;
  int plane_num_ = 0;
#ifdef DRFE_WITH_OPENCV
  cv::Mat seg_output, color_img_;
#else
  drfe_compat::Mat8u seg_output;
#endif

  explicit PlaneDetection(int device = 0) : device_(device) {      // PlaneExtractor.cpp:7-11
    cloud.vertices.resize((size_t)kDepthHeight * kDepthWidth);
    cloud.w = kDepthWidth;
    cloud.h = kDepthHeight;
  }
  ~PlaneDetection() { if (h_) drfe_peac_destroy(h_); }
  PlaneDetection(const PlaneDetection&) = delete;
  PlaneDetection& operator=(const PlaneDetection&) = delete;

#ifdef DRFE_WITH_OPENCV
  bool readColorImage(cv::Mat RGBImg) {                            // PlaneExtractor.cpp:19-26 (not used by the path)
    color_img_ = RGBImg;
    return !(color_img_.empty() || color_img_.depth() != CV_8U);
  }
  bool readDepthImage(cv::Mat depthImg, cv::Mat& K, const float depthfactor) {   // PlaneExtractor.cpp:28-55
    if (depthImg.empty() || depthImg.depth() != CV_16U) return false;
    return enqueue(depthImg.ptr<unsigned short>(0), depthImg.rows, depthImg.cols, (size_t)depthImg.step / sizeof(unsigned short), depthfactor,
                   K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2));
  }
#else
  bool readDepthImage(const drfe_compat::Mat16u& depthImg, const float K[9], const float depthfactor) {
    if (depthImg.empty()) return false;
    return enqueue(depthImg.data, depthImg.rows, depthImg.cols, depthImg.step / sizeof(uint16_t), depthfactor, K[0], K[4], K[2], K[5]);
  }
#endif

  void runPlaneDetection() {                                       // PlaneExtractor.cpp:57-63
    if (!h_ || !enqueued_) throw std::runtime_error("PlaneDetection::runPlaneDetection: readDepthImage first");
    enqueued_ = false;
    const int W = cloud.w, H = cloud.h;
#ifdef DRFE_WITH_OPENCV
    seg_output.create(H, W, CV_8UC1);
#else
    seg_output.create(H, W);
#endif
    std::vector<drfe_peac_plane> planes(255);
    seg_buf_.resize((size_t)W * H);
    check(drfe_peac_download(h_, seg_buf_.data(), planes.data(), 255, &plane_num_), "drfe_peac_download");
    for (int r = 0; r < H; ++r) std::memcpy(seg_output.ptr(r), seg_buf_.data() + (size_t)r * W, (size_t)W);
    idx_.resize((size_t)W * H);
    pts_.resize((size_t)W * H * 3);
    offs_.assign(256, 0);
    check(drfe_peac_plane_vertices(h_, idx_.data(), pts_.data(), (size_t)W * H, offs_.data(), 255), "drfe_peac_plane_vertices");
    for (size_t j : touched_) cloud.vertices[j] = VertexType{0.0, 0.0, 0.0};
    touched_.clear();
    plane_vertices_.assign(plane_num_, std::vector<int>());
    plane_filter.extractedPlanes.clear();
    for (int p = 0; p < plane_num_; ++p) {
      plane_vertices_[p].assign(idx_.begin() + offs_[p], idx_.begin() + offs_[p + 1]);
      for (int k = offs_[p]; k < offs_[p + 1]; ++k) {
        cloud.vertices[idx_[k]] = VertexType{(double)pts_[3 * (size_t)k], (double)pts_[3 * (size_t)k + 1], (double)pts_[3 * (size_t)k + 2]};
        touched_.push_back((size_t)idx_[k]);
      }
      auto seg = std::make_shared<ahc::PlaneSeg>();
      seg->rid = planes[p].rid; seg->N = planes[p].N; seg->mse = planes[p].mse; seg->curvature = planes[p].curvature;
      for (int k = 0; k < 3; ++k) { seg->center[k] = planes[p].center[k]; seg->normal[k] = planes[p].normal[k]; }
      plane_filter.extractedPlanes.push_back(seg);
    }
  }

  // What Frame::ComputePlanes does with every plane next (Frame.cc:954-990): the vertices with (float) z <= max_point_dist through
  // pcl::VoxelGrid(leaf, leaf, leaf), on the device — coarse[i] is `coarseCloud` of plane i (x, y, z triples).  Call after
  // runPlaneDetection(); the per-plane lists do not have to come to the host for it.
  void planeCloudsVoxel(float max_point_dist, float leaf, std::vector<std::vector<std::array<float, 3>>>& coarse) {
    if (!h_) throw std::runtime_error("PlaneDetection::planeCloudsVoxel: runPlaneDetection first");
    const size_t N = (size_t)cloud.w * cloud.h;
    pts_.resize(N * 3);
    offs_.assign(256, 0);
    check(drfe_peac_plane_points_voxel(h_, max_point_dist, leaf, pts_.data(), N, offs_.data(), 255), "drfe_peac_plane_points_voxel");
    coarse.assign(plane_num_, std::vector<std::array<float, 3>>());
    for (int p = 0; p < plane_num_; ++p) {
      const std::array<float, 3>* first = reinterpret_cast<const std::array<float, 3>*>(pts_.data()) + offs_[p];
      coarse[p].assign(first, first + (offs_[p + 1] - offs_[p]));
    }
  }

  // ... and on the depth image (Frame.cc:1044-1100): the 1/3-resolution cloud (far points zeroed) and PCL's integral-image surface
  // normals on it.  cloud3 / normals3: [h3][w3][3] floats, inputCloud->at(n, m) = cloud3[(m * w3 + n) * 3 ..], NaN where PCL leaves NaN
  void thirdCloudNormals(float max_point_dist, std::vector<float>& cloud3, std::vector<float>& normals3, int& w3, int& h3) {
    if (!h_) throw std::runtime_error("PlaneDetection::thirdCloudNormals: runPlaneDetection first");
    w3 = (cloud.w + 2) / 3; h3 = (cloud.h + 2) / 3;
    cloud3.resize((size_t)w3 * h3 * 3); normals3.resize((size_t)w3 * h3 * 3);
    check(drfe_peac_third_cloud_normals(h_, max_point_dist, 0.05f, 10.0f, cloud3.data(), normals3.data()), "drfe_peac_third_cloud_normals");
  }

 private:
  static void check(int rc, const char* what) {
    if (rc != DRFE_OK) throw std::runtime_error(std::string(what) + ": " + drfe_last_error());
  }
  bool enqueue(const uint16_t* depth, int rows, int cols, size_t row_stride, float factor, float fx, float fy, float cx, float cy) {
    if (!h_ || rows != cloud.h || cols != cloud.w) {
      if (h_) { drfe_peac_destroy(h_); h_ = nullptr; }
      check(drfe_peac_create(cols, rows, &plane_filter.params, 1, device_, &h_), "drfe_peac_create");
      cloud.w = cols; cloud.h = rows;
      cloud.vertices.assign((size_t)rows * cols, VertexType{0.0, 0.0, 0.0});
      touched_.clear();
    }
    check(drfe_peac_enqueue_depth_u16(h_, 1, depth, row_stride, row_stride * (size_t)rows, DRFE_MEM_HOST, factor, fx, fy, cx, cy),
          "drfe_peac_enqueue_depth_u16");
    enqueued_ = true;
    return true;
  }

  int device_ = 0;
  drfe_peac* h_ = nullptr;
  bool enqueued_ = false;
  std::vector<uint8_t> seg_buf_;
  std::vector<int32_t> idx_;
  std::vector<float> pts_;
  std::vector<int> offs_;
  std::vector<size_t> touched_;
};

}  // namespace Planar_SLAM
