// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
// legs may load this library.
//
// The surface normals DR-SLAM takes from PCL for the 1/3-resolution cloud (reference src/Frame.cc:1174-1216 and :1057-1100):
//     pcl::IntegralImageNormalEstimation<PointT, pcl::Normal> ne;
//     ne.setNormalEstimationMethod(ne.AVERAGE_3D_GRADIENT); ne.setMaxDepthChangeFactor(0.05f); ne.setNormalSmoothingSize(10.0f);
//     ne.setInputCloud(inputCloud); ne.compute(*cloud_normals);
// PCL is a dependency of the reference that is NOT vendored in it (CMakeLists.txt:59 `find_package(PCL 1.9 REQUIRED)`) and is absent
// from this container.  This file restates the PUBLISHED algorithm of PCL 1.9.x
//     features/include/pcl/features/impl/integral_image_normal.hpp   (computeFeature, computeFeatureFull, initAverage3DGradientMethod,
//                                                                     computePointNormal / AVERAGE_3D_GRADIENT, flipNormalTowardsViewpoint)
//     features/include/pcl/features/impl/integral_image2D.hpp        (IntegralImage2D<float, 3>: double first-order sums + finite counts)
// as its author remembers it, with the defaults DR-SLAM leaves in place: BORDER_POLICY_IGNORE, use_depth_dependent_smoothing_ = false,
// viewpoint (0, 0, 0), rectangular organized input.  The steps:
//   1. depth-change map: a pixel and its right (lower) neighbour are marked when |z - z_r| > f * (|z| + 1) * 2 (float) or either is
//      not finite (rows 0 .. H-2, columns 0 .. W-2);
//   2. distance map: 0 at marked pixels, W + H elsewhere, then the two raster passes of the 3-4 chamfer transform with weights 1.0f /
//      1.4f.  As in PCL the passes index one element past the row (previous_row[ci + 1] at the last column is the first element of the
//      current row; next_row[ci - 1] at column 0 is the last element of the current row); before the first element of the image the
//      second pass reads nothing because it ends at row 0 with next_row = row 1;
//   3. the image border of int(smoothing_size) pixels gets NaN normals; inside, a non-finite depth gives NaN; the window size is
//      s = int(min(distance, smoothing_size)), NaN if min(...) <= 2;
//   4. AVERAGE_3D_GRADIENT: DX(r, c) = P(r, c+1) - P(r, c-1), DY(r, c) = P(r+1, c) - P(r-1, c) per component in float on the interior,
//      0 on the image border; integral images of both in DOUBLE by I(r+1, c+1) = I(r, c+1) + I(r+1, c) - I(r, c) (+ the element if its
//      x + y + z is finite, which also counts it); the window [c - s/2, r - s/2] of s x s elements: NaN when either finite count is 0;
//      n = gy x gx (double), NaN when |n|^2 == 0, n / sqrt(|n|^2), cast to float, flipped so that (0 - P) . n >= 0; curvature is NaN.
// Details a maintainer with PCL 1.9 at hand should check first (they decide last bits and the NaN outline, not the normals of smooth
// surfaces): U.1 the bounds of the two raster passes (rows 1..H-1 / cols 1..W-1, then rows H-2..0 / cols W-2..0) and so which reads fall
// one element outside the row; U.2 the threshold's trailing `* 2.0f`; U.3 the window origin `pos - rect/2` with rect = int(smoothing)
// for both axes; U.4 the association (previous[c+1] + current[c]) - previous[c] in the integral-image recurrence.
// PARITY STATUS: "parity unpinned" — no PCL here to run, the reference has no fixture; the restatement is pinned only by its own
// properties (tests/test_normals.py: exact normals on synthetic planes, NaN pattern, the window rule).  The GPU path
// (drfe_cape_third_cloud_normals) is bit-identical to THIS restatement.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "drfe_oracle.h"

extern "C" {

// cloud: [h][w][3] float (x, y, z), organized; normals out: [h][w][3] float (NaN where PCL leaves NaN); distance_map out (optional): [h][w]
void orc_integral_normals(const float* cloud, int w, int h, float max_depth_change_factor, float smoothing_size, float* normals, float* distance_map) {
  const float nan = std::numeric_limits<float>::quiet_NaN();
  const size_t n = (size_t)w * h;
  std::vector<unsigned char> change(n, 255);
  auto Z = [&](size_t i) { return cloud[3 * i + 2]; };
  for (int ri = 0; ri < h - 1; ++ri)
    for (int ci = 0; ci < w - 1; ++ci) {
      const size_t index = (size_t)ri * w + ci;
      const float depth = Z(index), depthR = Z(index + 1), depthD = Z(index + w);
      const float th = (max_depth_change_factor * (fabsf(depth) + 1.0f) * 2.0f);
      if (std::fabs(depth - depthR) > th || !std::isfinite(depth) || !std::isfinite(depthR)) { change[index] = 0; change[index + 1] = 0; }
      if (std::fabs(depth - depthD) > th || !std::isfinite(depth) || !std::isfinite(depthD)) { change[index] = 0; change[index + w] = 0; }
    }
  std::vector<float> dist(n + 1);                                 // (+1: the first pass reads one element past the last row's end)
  for (size_t i = 0; i < n; ++i) dist[i] = change[i] == 0 ? 0.0f : (float)(w + h);
  dist[n] = 0.0f;
  {
    float* previous_row = dist.data();
    float* current_row = previous_row + w;
    for (int ri = 1; ri < h; ++ri) {
      for (int ci = 1; ci < w; ++ci) {
        const float upLeft = previous_row[ci - 1] + 1.4f, up = previous_row[ci] + 1.0f, upRight = previous_row[ci + 1] + 1.4f;
        const float left = current_row[ci - 1] + 1.0f, center = current_row[ci];
        const float minValue = std::min(std::min(upLeft, up), std::min(left, upRight));
        if (minValue < center) current_row[ci] = minValue;
      }
      previous_row = current_row;
      current_row += w;
    }
    float* next_row = dist.data() + (size_t)w * (h - 1);
    current_row = next_row - w;
    for (int ri = h - 2; ri >= 0; --ri) {
      for (int ci = w - 2; ci >= 0; --ci) {
        const float lowerLeft = next_row[ci - 1] + 1.4f, lower = next_row[ci] + 1.0f, lowerRight = next_row[ci + 1] + 1.4f;
        const float right = current_row[ci + 1] + 1.0f, center = current_row[ci];
        const float minValue = std::min(std::min(lowerLeft, lower), std::min(right, lowerRight));
        if (minValue < center) current_row[ci] = minValue;
      }
      next_row = current_row;
      current_row -= w;
    }
  }
  if (distance_map) memcpy(distance_map, dist.data(), n * sizeof(float));
  // central differences (interior only) and their integral images
  std::vector<float> dx(n * 3, 0.f), dy(n * 3, 0.f);
  for (int ri = 1; ri < h - 1; ++ri)
    for (int ci = 1; ci < w - 1; ++ci) {
      const size_t i = (size_t)ri * w + ci;
      for (int k = 0; k < 3; ++k) {
        dx[3 * i + k] = cloud[3 * (i + 1) + k] - cloud[3 * (i - 1) + k];
        dy[3 * i + k] = cloud[3 * (i + w) + k] - cloud[3 * (i - w) + k];
      }
    }
  const int W1 = w + 1;
  std::vector<double> ix((size_t)W1 * (h + 1) * 3, 0.0), iy((size_t)W1 * (h + 1) * 3, 0.0);
  std::vector<unsigned> cx((size_t)W1 * (h + 1), 0u), cy((size_t)W1 * (h + 1), 0u);
  auto integrate = [&](const std::vector<float>& d, std::vector<double>& I, std::vector<unsigned>& Cn) {
    for (int r = 0; r < h; ++r) {
      double* prev = &I[(size_t)r * W1 * 3];
      double* cur = prev + (size_t)W1 * 3;
      unsigned* cprev = &Cn[(size_t)r * W1];
      unsigned* ccur = cprev + W1;
      cur[0] = cur[1] = cur[2] = 0.0;
      ccur[0] = 0;
      for (int c = 0; c < w; ++c) {
        for (int k = 0; k < 3; ++k) cur[3 * (c + 1) + k] = prev[3 * (c + 1) + k] + cur[3 * c + k] - prev[3 * c + k];
        ccur[c + 1] = cprev[c + 1] + ccur[c] - cprev[c];
        const float* e = &d[((size_t)r * w + c) * 3];
        if (std::isfinite(e[0] + e[1] + e[2])) {
          for (int k = 0; k < 3; ++k) cur[3 * (c + 1) + k] += (double)e[k];
          ++ccur[c + 1];
        }
      }
    }
  };
  integrate(dx, ix, cx);
  integrate(dy, iy, cy);
  auto rect_sum = [&](const std::vector<double>& I, int sx, int sy, int rw, int rh, double out[3]) {
    const size_t ul = (size_t)sy * W1 + sx, ur = ul + rw, ll = (size_t)(sy + rh) * W1 + sx, lr = ll + rw;
    for (int k = 0; k < 3; ++k) out[k] = I[3 * lr + k] + I[3 * ul + k] - I[3 * ur + k] - I[3 * ll + k];
  };
  auto rect_cnt = [&](const std::vector<unsigned>& Cn, int sx, int sy, int rw, int rh) {
    const size_t ul = (size_t)sy * W1 + sx, ur = ul + rw, ll = (size_t)(sy + rh) * W1 + sx, lr = ll + rw;
    return Cn[lr] + Cn[ul] - Cn[ur] - Cn[ll];
  };
  for (size_t i = 0; i < n * 3; ++i) normals[i] = nan;
  const int border = (int)smoothing_size;
  for (int ri = border; ri < h - border; ++ri)
    for (int ci = border; ci < w - border; ++ci) {
      const size_t index = (size_t)ri * w + ci;
      const float depth = Z(index);
      if (!std::isfinite(depth)) continue;
      const float smoothing = std::min(dist[index], smoothing_size);
      if (!(smoothing > 2.0f)) continue;
      const int rw = (int)smoothing, rh = (int)smoothing, rw2 = rw / 2, rh2 = rh / 2;
      const unsigned count_x = rect_cnt(cx, ci - rw2, ri - rh2, rw, rh), count_y = rect_cnt(cy, ci - rw2, ri - rh2, rw, rh);
      if (count_x == 0 || count_y == 0) continue;
      double gx[3], gy[3];
      rect_sum(ix, ci - rw2, ri - rh2, rw, rh, gx);
      rect_sum(iy, ci - rw2, ri - rh2, rw, rh, gy);
      double nv[3] = {gy[1] * gx[2] - gy[2] * gx[1], gy[2] * gx[0] - gy[0] * gx[2], gy[0] * gx[1] - gy[1] * gx[0]};   // gradient_y.cross(gradient_x)
      const double len2 = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
      if (len2 == 0.0) continue;
      const double len = std::sqrt(len2);
      float nx = (float)(nv[0] / len), ny = (float)(nv[1] / len), nz = (float)(nv[2] / len);
      const float vx = 0.f - cloud[3 * index], vy = 0.f - cloud[3 * index + 1], vz = 0.f - cloud[3 * index + 2];   // flipNormalTowardsViewpoint
      const float cos_theta = (vx * nx + vy * ny + vz * nz);
      if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
      normals[3 * index] = nx; normals[3 * index + 1] = ny; normals[3 * index + 2] = nz;
    }
}

}  // extern "C"
