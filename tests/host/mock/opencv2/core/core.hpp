// MOCK of the slice of <opencv2/core/core.hpp> that the drfe host adapters touch — test infrastructure only.
// This container has no OpenCV headers; compiling the adapters' DRFE_WITH_OPENCV branches against these functional
// stand-ins (same names, same member signatures as OpenCV 3.4 for everything used) type-checks the real drop-in code
// and lets tests/host/adapters_opencv.cpp run it.  Nothing here is shipped or used by the product.
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_Assert(expr) assert(expr)

typedef unsigned char uchar;

namespace cv {

struct Point2f { float x, y; };
struct Scalar { double val[4]; Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) : val{v0, v1, v2, v3} {} };

class KeyPoint {
 public:
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct MatStep {
  size_t p = 0;
  operator size_t() const { return p; }
};

class Mat {
 public:
  int rows = 0, cols = 0;
  uchar* data = nullptr;
  MatStep step;
  Mat() = default;
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* d, size_t s = 0) : rows(r), cols(c), data((uchar*)d), type_(type) { step.p = s ? s : (size_t)c * elemSize(); }
  void create(int r, int c, int type) {
    if (r == rows && c == cols && type == type_ && store_) return;
    rows = r; cols = c; type_ = type;
    step.p = (size_t)c * elemSize();
    store_ = std::make_shared<std::vector<uchar>>(step.p * (size_t)r);
    data = store_->data();
  }
  void release() { rows = cols = 0; data = nullptr; store_.reset(); step.p = 0; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize() const { return (size_t)channels() * (depth() == CV_32F ? 4 : depth() == CV_16U ? 2 : 1); }
  uchar* ptr(int r = 0) { return data + (size_t)r * step.p; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step.p; }
  template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.p); }
  template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.p); }
  template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step.p))[c]; }
  template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step.p))[c]; }
  Mat& setTo(const Scalar& s) {
    assert(depth() == CV_8U);
    for (int r = 0; r < rows; ++r) std::memset(ptr(r), (int)s.val[0], (size_t)cols * elemSize());
    return *this;
  }

 private:
  int type_ = 0;
  std::shared_ptr<std::vector<uchar>> store_;
};

class _InputArray {
 public:
  _InputArray() = default;
  _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
  Mat getMat() const { return m_ ? *m_ : Mat(); }
  bool empty() const { return !m_ || m_->empty(); }

 protected:
  Mat* m_ = nullptr;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray(Mat& m) { m_ = &m; }
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void release() const { m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

}  // namespace cv
