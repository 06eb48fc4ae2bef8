// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product path (libdrfe.so) never links or calls it.
//
// CPU restatement of DR-SLAM's ORB front end (reference: src/ORBextractor.cc).
// Dependency-free C++17: the OpenCV-owned primitives the reference calls
// (cv::resize INTER_LINEAR 8U, copyMakeBorder REFLECT_101, cv::FAST 9/16 with
// NMS, GaussianBlur 7x7 sigma 2 8U, fastAtan2, cvRound) are restated here as
// integer / float32 models of OpenCV's published algorithms (OpenCV is a
// third-party dependency that is NOT vendored under /root/reference; the
// reference links OpenCV 3.4, README.md:27-29).  Each primitive is pinned
// bit-exact against cv2 4.13 in tests/test_oracle_vs_cv2.py and against the
// committed fixtures in tests/golden/ (made by tests/golden/make_golden.py).
//
// PARITY STATUS: "parity unpinned" by the reference itself — the reference
// ships no tests, golden vectors or fixtures for this path and
// cannot be compiled here (needs OpenCV 3.4 + Eigen headers).  What pins this
// oracle instead: (1) cv2 4.13 for every OpenCV-owned primitive, (2) an
// independent literal Python restatement (oracle/py_ref.py) of the
// reference-owned logic, compared on seeded frames.
//
// Declared deviations from "whatever the author's binary did" (all are
// allocator / compiler defined in the reference):
//   * octree fine-phase sort tie: reference sorts pair<int, ExtractorNode*>
//     (ORBextractor.cc:684) i.e. ties by heap address; we tie-break by node
//     creation sequence number (ascending).
//   * descriptor steering float ops are evaluated without FMA contraction.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

#include "drfe_oracle.h"

namespace {

const int kPatch = 31;      // ORBextractor.cc:72
const int kHalfPatch = 15;  // ORBextractor.cc:73
const int kEdge = 19;       // ORBextractor.cc:74

const int8_t kPattern[1024] = {
#include "../include/drfe_orb_pattern.inc"
};

// cvRound: SSE cvtss2si / cvtsd2si under default MXCSR = round-half-even.
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }

// ---------------------------------------------------------------- images
struct Image {
  int w = 0, h = 0, stride = 0;
  std::vector<uint8_t> buf;
  void alloc(int w_, int h_) { w = w_; h = h_; stride = w_; buf.assign((size_t)w_ * h_, 0); }
  uint8_t* row(int y) { return buf.data() + (size_t)y * stride; }
  const uint8_t* row(int y) const { return buf.data() + (size_t)y * stride; }
};

// BORDER_REFLECT_101 index map (OpenCV borderInterpolate): -1 -> 1, n -> n-2.
inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
  return i;
}

// cv::resize(..., INTER_LINEAR) for 8UC1: fixed point, 11-bit coefficients
// (OpenCV imgproc resize.cpp: HResizeLinear<uchar,int,short> + VResizeLinear
// <uchar,int,short,FixedPtCast<int,uchar,22>>, INTER_RESIZE_COEF_BITS = 11).
struct ResizeTab {
  std::vector<int> idx;
  std::vector<short> c0, c1;
};
ResizeTab make_resize_tab(int ssize, int dsize) {
  ResizeTab t;
  t.idx.resize(dsize); t.c0.resize(dsize); t.c1.resize(dsize);
  double scale = 1.0 / ((double)dsize / ssize);  // inv_scale = dsize/ssize; scale = 1/inv_scale
  for (int d = 0; d < dsize; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    t.idx[d] = s;
    // saturate_cast<short>(float) = cvRound then clamp
    t.c0[d] = (short)cv_round((1.f - f) * 2048.f);
    t.c1[d] = (short)cv_round(f * 2048.f);
  }
  return t;
}
void resize_linear_u8(const Image& src, Image& dst, int dw, int dh) {
  dst.alloc(dw, dh);
  ResizeTab tx = make_resize_tab(src.w, dw), ty = make_resize_tab(src.h, dh);
  std::vector<int> r0(dw), r1(dw);
  int cached0 = -1, cached1 = -1;
  auto hrow = [&](int sy, std::vector<int>& out) {
    const uint8_t* s = src.row(sy);
    for (int x = 0; x < dw; ++x) {
      int sx = tx.idx[x];
      int sx1 = std::min(sx + 1, src.w - 1);
      out[x] = s[sx] * tx.c0[x] + s[sx1] * tx.c1[x];
    }
  };
  for (int y = 0; y < dh; ++y) {
    int sy0 = ty.idx[y], sy1 = std::min(sy0 + 1, src.h - 1);
    if (sy0 == cached1) { std::swap(r0, r1); std::swap(cached0, cached1); }
    if (sy0 != cached0) { hrow(sy0, r0); cached0 = sy0; }
    if (sy1 == cached0) { r1 = r0; cached1 = sy1; }
    else if (sy1 != cached1) { hrow(sy1, r1); cached1 = sy1; }
    int b0 = ty.c0[y], b1 = ty.c1[y];
    uint8_t* d = dst.row(y);
    for (int x = 0; x < dw; ++x) {
      int v = (((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2;
      d[x] = (uint8_t)v;  // always within 0..255 for valid coefficients
    }
  }
}

// copyMakeBorder(src, dst, b,b,b,b, BORDER_REFLECT_101).
void add_border(const Image& src, Image& dst, int b) {
  dst.alloc(src.w + 2 * b, src.h + 2 * b);
  for (int y = 0; y < dst.h; ++y) {
    const uint8_t* s = src.row(reflect101(y - b, src.h));
    uint8_t* d = dst.row(y);
    for (int x = 0; x < dst.w; ++x) d[x] = s[reflect101(x - b, src.w)];
  }
}

// ---------------------------------------------------------------- FAST 9/16
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// OpenCV cornerScore<16>(ptr, pixel, threshold) (features2d fast_score.cpp).
int corner_score16(const uint8_t* p, const int* off, int threshold) {
  int d[25];
  int v = p[0];
  for (int k = 0; k < 25; ++k) d[k] = v - p[off[k & 15]];
  int a0 = threshold;
  for (int k = 0; k < 16; k += 2) {
    int a = std::min(d[k + 1], d[k + 2]);
    a = std::min(a, d[k + 3]);
    if (a <= a0) continue;
    a = std::min(a, d[k + 4]); a = std::min(a, d[k + 5]);
    a = std::min(a, d[k + 6]); a = std::min(a, d[k + 7]);
    a = std::min(a, d[k + 8]);
    a0 = std::max(a0, std::min(a, d[k]));
    a0 = std::max(a0, std::min(a, d[k + 9]));
  }
  int b0 = -a0;
  for (int k = 0; k < 16; k += 2) {
    int b = std::max(d[k + 1], d[k + 2]);
    b = std::max(b, d[k + 3]); b = std::max(b, d[k + 4]); b = std::max(b, d[k + 5]);
    if (b >= b0) continue;
    b = std::max(b, d[k + 6]); b = std::max(b, d[k + 7]); b = std::max(b, d[k + 8]);
    b0 = std::min(b0, std::max(b, d[k]));
    b0 = std::min(b0, std::max(b, d[k + 9]));
  }
  return -b0 - 1;
}

struct Cand { float x, y, response; };

// cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16, on the
// w x h sub-image starting at `base` (row stride `stride`).  Keypoints come
// out row-major, pt = (x, y) sub-image relative, response = cornerScore.
void fast9_nms(const uint8_t* base, int stride, int w, int h, int threshold,
               std::vector<Cand>& out) {
  if (w < 7 || h < 7) return;
  int off[16];
  for (int k = 0; k < 16; ++k) off[k] = kRingDy[k] * stride + kRingDx[k];
  // score rows: 0 for non-corners.  Keep the whole map (cells are ~36x36).
  std::vector<int> score((size_t)w * h, 0);
  for (int y = 3; y < h - 3; ++y) {
    const uint8_t* r = base + (size_t)y * stride;
    for (int x = 3; x < w - 3; ++x) {
      const uint8_t* p = r + x;
      int v = p[0], hi = v + threshold, lo = v - threshold;
      // 9 contiguous of 16 all brighter than hi or all darker than lo.
      unsigned br = 0, dk = 0;
      for (int k = 0; k < 16; ++k) {
        int q = p[off[k]];
        br |= (unsigned)(q > hi) << k;
        dk |= (unsigned)(q < lo) << k;
      }
      auto has_arc9 = [](unsigned m) {
        m |= m << 16;  // unroll the circle
        unsigned r2 = m & (m >> 1);
        unsigned r4 = r2 & (r2 >> 2);
        unsigned r8 = r4 & (r4 >> 4);
        unsigned r9 = r8 & (m >> 8);
        return (r9 & 0xFFFFu) != 0;
      };
      if (has_arc9(br) || has_arc9(dk)) score[(size_t)y * w + x] = corner_score16(p, off, threshold);
    }
  }
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      int s = score[(size_t)y * w + x];
      if (s == 0) continue;  // corner scores are >= threshold >= 1
      const int* c = &score[(size_t)y * w + x];
      if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] &&
          s > c[w - 1] && s > c[w] && s > c[w + 1])
        out.push_back({(float)x, (float)y, (float)s});
    }
}

// ---------------------------------------------------------------- fastAtan2
// OpenCV core mathfuncs_core: 7th-order odd polynomial, plain float32.
float fast_atan2_deg(float y, float x) {
  const float s = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s,
              p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ---------------------------------------------------------------- Gaussian
// cv::GaussianBlur(7x7, sigma 2) on 8U: fixed-point 8.8 separable kernel
// {18,34,48,56,48,34,18}/256, BORDER_REFLECT_101 (OpenCV smooth: fixedSmooth
// <uint8_t, ufixedpoint16>).
void gaussian7_u8(const Image& src, Image& dst) {
  static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
  dst.alloc(src.w, src.h);
  std::vector<uint16_t> tmp((size_t)src.w * src.h);
  for (int y = 0; y < src.h; ++y) {
    const uint8_t* s = src.row(y);
    for (int x = 0; x < src.w; ++x) {
      int acc = 0;
      for (int i = 0; i < 7; ++i) acc += k[i] * s[reflect101(x + i - 3, src.w)];
      tmp[(size_t)y * src.w + x] = (uint16_t)acc;  // <= 256*255 fits u16
    }
  }
  for (int y = 0; y < src.h; ++y) {
    const uint16_t* r[7];
    for (int i = 0; i < 7; ++i) r[i] = &tmp[(size_t)reflect101(y + i - 3, src.h) * src.w];
    uint8_t* d = dst.row(y);
    for (int x = 0; x < src.w; ++x) {
      uint32_t acc = 0;
      for (int i = 0; i < 7; ++i) acc += (uint32_t)k[i] * r[i][x];
      uint32_t v = (acc + 32768u) >> 16;
      d[x] = (uint8_t)std::min(v, 255u);
    }
  }
}

// ---------------------------------------------------------------- quadtree
struct QKey { float x, y, response; };
struct QNode {
  int x0, y0, x1, y1;          // UL.x, UL.y, UR.x(=BR.x), BL.y(=BR.y)
  std::vector<QKey> keys;
  bool frozen = false;         // bNoMore
  long seq = 0;                // creation sequence number (declared tie rule)
  std::list<QNode>::iterator self;
};

// ExtractorNode::DivideNode (ORBextractor.cc:481-537)
void quarter(const QNode& n, QNode ch[4]) {
  const int hx = (int)std::ceil((float)(n.x1 - n.x0) / 2);
  const int hy = (int)std::ceil((float)(n.y1 - n.y0) / 2);
  const int mx = n.x0 + hx, my = n.y0 + hy;
  ch[0] = QNode{n.x0, n.y0, mx, my};
  ch[1] = QNode{mx, n.y0, n.x1, my};
  ch[2] = QNode{n.x0, my, mx, n.y1};
  ch[3] = QNode{mx, my, n.x1, n.y1};
  for (const QKey& k : n.keys) {
    int c = (k.x < (float)mx) ? ((k.y < (float)my) ? 0 : 2) : ((k.y < (float)my) ? 1 : 3);
    ch[c].keys.push_back(k);
  }
  for (int c = 0; c < 4; ++c) ch[c].frozen = (ch[c].keys.size() == 1);
}

// ORBextractor::DistributeOctTree (ORBextractor.cc:539-763).
// `tie_flag` (optional) is set when the fine-phase "size >= N" break cuts
// through a group of equal-size nodes, i.e. where the reference's result
// would depend on heap addresses.
std::vector<QKey> distribute_quadtree(const std::vector<QKey>& in, int minX, int maxX, int minY,
                                      int maxY, int N, int* tie_flag) {
  const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
  const float hX = (float)(maxX - minX) / nIni;
  std::list<QNode> nodes;
  long seq = 0;
  std::vector<QNode*> roots(nIni);
  for (int i = 0; i < nIni; ++i) {
    QNode r{(int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), maxY - minY};
    r.seq = seq++;
    nodes.push_back(r);
    roots[i] = &nodes.back();
  }
  for (const QKey& k : in) roots[(int)(k.x / hX)]->keys.push_back(k);
  for (auto it = nodes.begin(); it != nodes.end();) {
    if (it->keys.size() == 1) { it->frozen = true; ++it; }
    else if (it->keys.empty()) it = nodes.erase(it);
    else ++it;
  }
  typedef std::pair<int, QNode*> SizedNode;
  auto by_size_then_seq = [](const SizedNode& a, const SizedNode& b) {
    return a.first != b.first ? a.first < b.first : a.second->seq < b.second->seq;
  };
  // split one node: children to the list front in order 0..3, parent removed by caller
  auto push_children = [&](QNode& parent, std::vector<SizedNode>& expandable, int& nExpand) {
    QNode ch[4];
    quarter(parent, ch);
    for (int c = 0; c < 4; ++c) {
      if (ch[c].keys.empty()) continue;
      ch[c].seq = seq++;
      nodes.push_front(ch[c]);
      nodes.front().self = nodes.begin();
      if (nodes.front().keys.size() > 1) {
        ++nExpand;
        expandable.push_back(SizedNode((int)nodes.front().keys.size(), &nodes.front()));
      }
    }
  };
  bool done = false;
  std::vector<SizedNode> expandable;
  while (!done) {
    const int before = (int)nodes.size();
    int nExpand = 0;
    expandable.clear();
    for (auto it = nodes.begin(); it != nodes.end();) {
      if (it->frozen) { ++it; continue; }
      push_children(*it, expandable, nExpand);
      it = nodes.erase(it);
    }
    if ((int)nodes.size() >= N || (int)nodes.size() == before) {
      done = true;
    } else if ((int)nodes.size() + nExpand * 3 > N) {
      while (!done) {
        const int before2 = (int)nodes.size();
        std::vector<SizedNode> prev = expandable;
        expandable.clear();
        std::sort(prev.begin(), prev.end(), by_size_then_seq);
        for (int j = (int)prev.size() - 1; j >= 0; --j) {
          int dummy = 0;
          push_children(*prev[j].second, expandable, dummy);
          nodes.erase(prev[j].second->self);
          if ((int)nodes.size() >= N) {
            if (tie_flag && j > 0 && prev[j - 1].first == prev[j].first) *tie_flag = 1;
            break;
          }
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == before2) done = true;
      }
    }
  }
  std::vector<QKey> best;
  best.reserve(nodes.size());
  for (const QNode& n : nodes) {
    const QKey* b = &n.keys[0];
    for (size_t k = 1; k < n.keys.size(); ++k)
      if (n.keys[k].response > b->response) b = &n.keys[k];
    best.push_back(*b);
  }
  return best;
}

// ---------------------------------------------------------------- extractor
struct LevelData {
  Image bordered;   // (w+38) x (h+38), ROI at (19,19) == mvImagePyramid[level]
  Image plain;      // w x h copy of the ROI (what .clone() yields)
  Image blurred;
  std::vector<Cand> cands;      // region coords (origin = minBorder 16,16)
  std::vector<drfe_keypoint> kps;  // level coords, angle set, pt NOT yet scaled
  int tie = 0;
};

struct OrbOracle {
  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> perLevel, umax;
  std::vector<LevelData> lv;
  std::vector<drfe_keypoint> out_kps;
  std::vector<uint8_t> out_desc;

  // ORBextractor::ORBextractor (ORBextractor.cc:410-470)
  OrbOracle(int nf, float sf, int nl, int ini, int mn)
      : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
    scale.assign(nl, 1.f); sigma2.assign(nl, 1.f);
    for (int i = 1; i < nl; ++i) {
      scale[i] = (float)(scale[i - 1] * scaleFactor);
      sigma2[i] = scale[i] * scale[i];
    }
    invScale.resize(nl); invSigma2.resize(nl);
    for (int i = 0; i < nl; ++i) { invScale[i] = 1.0f / scale[i]; invSigma2[i] = 1.0f / sigma2[i]; }
    perLevel.resize(nl);
    float factor = (float)(1.0f / scaleFactor);
    float want = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nl - 1; ++l) {
      perLevel[l] = cv_round(want);
      sum += perLevel[l];
      want *= factor;
    }
    perLevel[nl - 1] = std::max(nfeatures - sum, 0);
    umax.assign(kHalfPatch + 1, 0);
    int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
    lv.resize(nl);
  }

  // ORBextractor::ComputePyramid (ORBextractor.cc:1107-1132)
  void pyramid(const uint8_t* gray, int w, int h, int stride) {
    for (int l = 0; l < nlevels; ++l) {
      float s = invScale[l];
      int lw = cv_round((float)w * s), lh = cv_round((float)h * s);
      if (l == 0) {
        lv[0].plain.alloc(lw, lh);
        for (int y = 0; y < h; ++y) memcpy(lv[0].plain.row(y), gray + (size_t)y * stride, w);
      } else {
        resize_linear_u8(lv[l - 1].plain, lv[l].plain, lw, lh);
      }
      add_border(lv[l].plain, lv[l].bordered, kEdge);
    }
  }

  // ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cc:765-853), per level,
  // up to (not including) orientation.
  void detect(int l) {
    LevelData& L = lv[l];
    const Image& im = L.plain;
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = im.w - kEdge + 3, maxBY = im.h - kEdge + 3;
    const float W = 30;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    L.cands.clear();
    std::vector<Cand> cell;
    for (int i = 0; i < nRows; ++i) {
      const float iniY = (float)(minBY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < nCols; ++j) {
        const float iniX = (float)(minBX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, chh = (int)maxY - y0;
        cell.clear();
        fast9_nms(im.row(y0) + x0, im.stride, cw, chh, iniTh, cell);
        if (cell.empty()) fast9_nms(im.row(y0) + x0, im.stride, cw, chh, minTh, cell);
        for (const Cand& c : cell)
          L.cands.push_back({c.x + (float)(j * wCell), c.y + (float)(i * hCell), c.response});
      }
    }
    std::vector<QKey> keys(L.cands.size());
    for (size_t k = 0; k < keys.size(); ++k) keys[k] = {L.cands[k].x, L.cands[k].y, L.cands[k].response};
    L.tie = 0;
    std::vector<QKey> kept;
    if (!keys.empty()) kept = distribute_quadtree(keys, minBX, maxBX, minBY, maxBY, perLevel[l], &L.tie);
    const int scaledPatch = (int)(kPatch * scale[l]);
    L.kps.clear();
    for (const QKey& k : kept) {
      drfe_keypoint kp;
      kp.x = k.x + (float)minBX;
      kp.y = k.y + (float)minBY;
      kp.size = (float)scaledPatch;
      kp.angle = -1.f;
      kp.response = k.response;
      kp.octave = l;
      kp.class_id = -1;
      L.kps.push_back(kp);
    }
  }

  // IC_Angle (ORBextractor.cc:77-104) on the un-blurred level image.
  float ic_angle(const Image& b, float px, float py) const {
    // `b` is the bordered image; ROI origin at (kEdge,kEdge)
    const int st = b.stride;
    const uint8_t* c = b.row(cv_round(py) + kEdge) + cv_round(px) + kEdge;
    int m01 = 0, m10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
      int vs = 0, d = umax[v];
      for (int u = -d; u <= d; ++u) {
        int p = c[u + v * st], m = c[u - v * st];
        vs += p - m;
        m10 += u * (p + m);
      }
      m01 += v * vs;
    }
    return fast_atan2_deg((float)m01, (float)m10);
  }

  // computeOrbDescriptor (ORBextractor.cc:107-147) on the blurred level image.
  static void describe(const Image& im, const drfe_keypoint& kp, uint8_t* desc) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    float angle = kp.angle * factorPI;
    float a = cosf(angle), b = sinf(angle);
    const int st = im.stride;
    const uint8_t* c = im.row(cv_round(kp.y)) + cv_round(kp.x);
    const int8_t* p = kPattern;
    for (int i = 0; i < 32; ++i, p += 32) {
      int val = 0;
      for (int j = 0; j < 8; ++j) {
        const float x0 = (float)p[4 * j], y0 = (float)p[4 * j + 1];
        const float x1 = (float)p[4 * j + 2], y1 = (float)p[4 * j + 3];
        volatile float r0a = x0 * b, r0b = y0 * a, c0a = x0 * a, c0b = y0 * b;  // no FMA contraction
        volatile float r1a = x1 * b, r1b = y1 * a, c1a = x1 * a, c1b = y1 * b;
        int t0 = c[cv_round(r0a + r0b) * st + cv_round(c0a - c0b)];
        int t1 = c[cv_round(r1a + r1b) * st + cv_round(c1a - c1b)];
        val |= (t0 < t1) << j;
      }
      desc[i] = (uint8_t)val;
    }
  }

  // ORBextractor::operator() (ORBextractor.cc:1043-1105)
  int run(const uint8_t* gray, int w, int h, int stride) {
    out_kps.clear(); out_desc.clear();
    if (!gray || w <= 0 || h <= 0) return 0;
    pyramid(gray, w, h, stride);
    for (int l = 0; l < nlevels; ++l) detect(l);
    for (int l = 0; l < nlevels; ++l)
      for (drfe_keypoint& kp : lv[l].kps) kp.angle = ic_angle(lv[l].bordered, kp.x, kp.y);
    for (int l = 0; l < nlevels; ++l) {
      LevelData& L = lv[l];
      if (L.kps.empty()) { L.blurred.alloc(0, 0); continue; }
      gaussian7_u8(L.plain, L.blurred);
      size_t base = out_desc.size();
      out_desc.resize(base + 32 * L.kps.size());
      for (size_t i = 0; i < L.kps.size(); ++i) describe(L.blurred, L.kps[i], &out_desc[base + 32 * i]);
      for (const drfe_keypoint& kp : L.kps) {
        drfe_keypoint o = kp;
        if (l != 0) { o.x *= scale[l]; o.y *= scale[l]; }
        out_kps.push_back(o);
      }
    }
    return (int)out_kps.size();
  }
};

}  // namespace

// ------------------------------------------------------------------ C API
extern "C" {

void* orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
  return new OrbOracle(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
void orc_orb_destroy(void* h) { delete (OrbOracle*)h; }
int orc_orb_run(void* h, const uint8_t* gray, int w, int hgt, int stride) {
  return ((OrbOracle*)h)->run(gray, w, hgt, stride);
}
int orc_orb_features_per_level(void* h, int level) { return ((OrbOracle*)h)->perLevel[level]; }
float orc_orb_scale_factor(void* h, int level) { return ((OrbOracle*)h)->scale[level]; }
int orc_orb_umax(void* h, int v) { return ((OrbOracle*)h)->umax[v]; }
int orc_orb_level_size(void* h, int level, int* w, int* hgt) {
  OrbOracle* o = (OrbOracle*)h;
  *w = o->lv[level].plain.w; *hgt = o->lv[level].plain.h;
  return 0;
}
// bordered != 0: (w+38)x(h+38) image incl. the 19-px reflect border
int orc_orb_get_level(void* h, int level, int bordered, uint8_t* dst) {
  OrbOracle* o = (OrbOracle*)h;
  const Image& im = bordered ? o->lv[level].bordered : o->lv[level].plain;
  memcpy(dst, im.buf.data(), im.buf.size());
  return 0;
}
int orc_orb_get_blurred(void* h, int level, uint8_t* dst) {
  OrbOracle* o = (OrbOracle*)h;
  memcpy(dst, o->lv[level].blurred.buf.data(), o->lv[level].blurred.buf.size());
  return (int)o->lv[level].blurred.buf.size();
}
// FAST candidates of a level in reference order; xyr = [x, y, response]* (region coords)
int orc_orb_get_candidates(void* h, int level, float* xyr, int cap) {
  OrbOracle* o = (OrbOracle*)h;
  int n = (int)o->lv[level].cands.size();
  for (int i = 0; i < n && i < cap; ++i) {
    xyr[3 * i] = o->lv[level].cands[i].x; xyr[3 * i + 1] = o->lv[level].cands[i].y;
    xyr[3 * i + 2] = o->lv[level].cands[i].response;
  }
  return n;
}
// octree-retained keypoints of a level (level coords, list order), angle filled
int orc_orb_get_level_keypoints(void* h, int level, drfe_keypoint* dst, int cap) {
  OrbOracle* o = (OrbOracle*)h;
  int n = (int)o->lv[level].kps.size();
  for (int i = 0; i < n && i < cap; ++i) dst[i] = o->lv[level].kps[i];
  return n;
}
int orc_orb_level_tie(void* h, int level) { return ((OrbOracle*)h)->lv[level].tie; }
int orc_orb_get_result(void* h, drfe_keypoint* kps, uint8_t* desc, int cap) {
  OrbOracle* o = (OrbOracle*)h;
  int n = (int)o->out_kps.size();
  for (int i = 0; i < n && i < cap; ++i) {
    kps[i] = o->out_kps[i];
    memcpy(desc + 32 * i, &o->out_desc[32 * (size_t)i], 32);
  }
  return n;
}

// ---- primitive-level entry points (pinned against cv2 in tests) ----
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  Image s, d;
  s.alloc(sw, sh); memcpy(s.buf.data(), src, (size_t)sw * sh);
  resize_linear_u8(s, d, dw, dh);
  memcpy(dst, d.buf.data(), (size_t)dw * dh);
}
void orc_border101_u8(const uint8_t* src, int w, int h, int b, uint8_t* dst) {
  Image s, d;
  s.alloc(w, h); memcpy(s.buf.data(), src, (size_t)w * h);
  add_border(s, d, b);
  memcpy(dst, d.buf.data(), d.buf.size());
}
int orc_fast9_nms(const uint8_t* img, int stride, int w, int h, int threshold, float* xyr, int cap) {
  std::vector<Cand> c;
  fast9_nms(img, stride, w, h, threshold, c);
  for (int i = 0; i < (int)c.size() && i < cap; ++i) {
    xyr[3 * i] = c[i].x; xyr[3 * i + 1] = c[i].y; xyr[3 * i + 2] = c[i].response;
  }
  return (int)c.size();
}
void orc_gaussian7_u8(const uint8_t* src, int w, int h, uint8_t* dst) {
  Image s, d;
  s.alloc(w, h); memcpy(s.buf.data(), src, (size_t)w * h);
  gaussian7_u8(s, d);
  memcpy(dst, d.buf.data(), (size_t)w * h);
}
float orc_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }
// standalone quadtree: xyr in (region coords), returns kept keys in list order
int orc_distribute_quadtree(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N,
                            float* out_xyr, int cap, int* tie) {
  std::vector<QKey> in(n);
  for (int i = 0; i < n; ++i) in[i] = {xyr[3 * i], xyr[3 * i + 1], xyr[3 * i + 2]};
  int t = 0;
  std::vector<QKey> r;
  if (n > 0) r = distribute_quadtree(in, minX, maxX, minY, maxY, N, &t);
  if (tie) *tie = t;
  for (int i = 0; i < (int)r.size() && i < cap; ++i) {
    out_xyr[3 * i] = r[i].x; out_xyr[3 * i + 1] = r[i].y; out_xyr[3 * i + 2] = r[i].response;
  }
  return (int)r.size();
}

// ---- per-frame steps after extraction: UndistortKeyPoints (Frame.cc:835-861; the
// arithmetic is cv::undistortPoints with R = I, P = K: 5 fixed-point iterations in double, OpenCV 3.4
// undistort.cpp cvUndistortPointsInternal), ComputeStereoFromRGBD (:893-911), AssignFeaturesToGrid
// (:224-237) with PosInGrid (:816-825).
void orc_undistort_point(const drfe_frame_params* p, float u, float v, float* ou, float* ov) {
  const double fx = p->fx, fy = p->fy, cx = p->cx, cy = p->cy;
  const double k1 = p->dist[0], k2 = p->dist[1], p1 = p->dist[2], p2 = p->dist[3], k3 = p->dist[4];
  const double ifx = 1. / fx, ify = 1. / fy;
  double x = ((double)u - cx) * ifx, y = ((double)v - cy) * ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
    const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
    const double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  // P = K, R = I: xx = fx*x + 0*y + cx, ww = 1 / (0*x + 0*y + 1)
  const double xx = fx * x + cx, yy = fy * y + cy;
  *ou = (float)xx; *ov = (float)yy;
}
int orc_frame_image_bounds(drfe_frame_params* p, int width, int height) {
  if (p->dist[0] != 0.0f) {
    float x[4], y[4];
    const float cu[4] = {0.f, (float)width, 0.f, (float)width}, cv[4] = {0.f, 0.f, (float)height, (float)height};
    for (int i = 0; i < 4; ++i) orc_undistort_point(p, cu[i], cv[i], &x[i], &y[i]);
    p->min_x = std::min(x[0], x[2]); p->max_x = std::max(x[1], x[3]);
    p->min_y = std::min(y[0], y[1]); p->max_y = std::max(y[2], y[3]);
  } else {
    p->min_x = 0.f; p->max_x = (float)width; p->min_y = 0.f; p->max_y = (float)height;
  }
  return 0;
}
int orc_frame_post(const drfe_frame_params* p, const drfe_keypoint* keys, int n, const float* depth, int row_stride,
                   drfe_keypoint* keys_un, float* u_right, float* kp_depth, uint16_t* grid_count, uint16_t* grid_index) {
  const int GC = DRFE_FRAME_GRID_COLS, GR = DRFE_FRAME_GRID_ROWS;
  const float inv_w = (float)GC / (float)(p->max_x - p->min_x), inv_h = (float)GR / (float)(p->max_y - p->min_y);
  std::vector<std::vector<uint16_t>> grid(GC * GR);
  for (int i = 0; i < n; ++i) {
    drfe_keypoint ku = keys[i];
    if (p->dist[0] != 0.0f) orc_undistort_point(p, keys[i].x, keys[i].y, &ku.x, &ku.y);
    keys_un[i] = ku;
    const float d = depth[(size_t)(int)keys[i].y * row_stride + (int)keys[i].x];   // imDepth.at<float>(v, u): floats truncate
    u_right[i] = -1.f; kp_depth[i] = -1.f;
    if (d > 0) { kp_depth[i] = d; u_right[i] = ku.x - p->bf / d; }
    const int px = (int)roundf((ku.x - p->min_x) * inv_w), py = (int)roundf((ku.y - p->min_y) * inv_h);
    if (px < 0 || px >= GC || py < 0 || py >= GR) continue;
    grid[px * GR + py].push_back((uint16_t)i);
  }
  int o = 0;
  for (int c = 0; c < GC * GR; ++c) {
    grid_count[c] = (uint16_t)grid[c].size();
    for (uint16_t i : grid[c]) grid_index[o++] = i;
  }
  return o;
}

}  // extern "C"
