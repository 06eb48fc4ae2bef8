// Minimal stand-ins for the OpenCV / Eigen types that appear in the reference signatures, used
// when the adapters are compiled without those libraries (this container has neither).  With
// -DDRFE_WITH_OPENCV / -DDRFE_WITH_EIGEN the adapters use the real cv:: / Eigen:: types
// instead and these are not needed.  Layouts match: compat::KeyPoint == cv::KeyPoint (28 B).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace drfe_compat {

struct Point2f { float x, y; };

struct KeyPoint {               // cv::KeyPoint: pt, size, angle, response, octave, class_id
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

// 8-bit single-channel image (the role of cv::Mat CV_8UC1 / cv::Mat_<uchar>): view or owner.
struct Mat8u {
  int rows = 0, cols = 0;
  size_t step = 0;
  uint8_t* data = nullptr;
  std::vector<uint8_t> store;
  Mat8u() = default;
  Mat8u(int r, int c, uint8_t* d, size_t s) : rows(r), cols(c), step(s), data(d) {}
  void create(int r, int c) {
    if (r == rows && c == cols && !store.empty()) return;
    rows = r; cols = c; step = (size_t)c;
    store.assign((size_t)r * c, 0);
    data = store.data();
  }
  void release() { rows = cols = 0; step = 0; data = nullptr; store.clear(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  void setTo(uint8_t v) { for (int r = 0; r < rows; ++r) std::memset(data + r * step, v, (size_t)cols); }
  uint8_t* ptr(int r) { return data + (size_t)r * step; }
  const uint8_t* ptr(int r) const { return data + (size_t)r * step; }
};

// float image (cv::Mat CV_32FC1), view only
struct Mat32f {
  int rows = 0, cols = 0;
  size_t step = 0;              // bytes
  const float* data = nullptr;
  Mat32f() = default;
  Mat32f(int r, int c, const float* d, size_t step_bytes) : rows(r), cols(c), step(step_bytes), data(d) {}
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
};

// 16-bit image (cv::Mat CV_16UC1: the raw depth map), view only
struct Mat16u {
  int rows = 0, cols = 0;
  size_t step = 0;              // bytes
  const uint16_t* data = nullptr;
  Mat16u() = default;
  Mat16u(int r, int c, const uint16_t* d, size_t step_bytes) : rows(r), cols(c), step(step_bytes), data(d) {}
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
};

// column-major N x 3 float matrix (the role of Eigen::MatrixXf in CAPE::process)
struct MatrixXf {
  int nrows = 0, ncols = 0;
  std::vector<float> store;
  MatrixXf() = default;
  MatrixXf(int r, int c) : nrows(r), ncols(c), store((size_t)r * c, 0.f) {}
  float* data() { return store.data(); }
  const float* data() const { return store.data(); }
  int rows() const { return nrows; }
  int cols() const { return ncols; }
  float& operator()(int r, int c) { return store[(size_t)c * nrows + r]; }
};

}  // namespace drfe_compat
