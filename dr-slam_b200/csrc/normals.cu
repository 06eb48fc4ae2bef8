// pcl::IntegralImageNormalEstimation (AVERAGE_3D_GRADIENT, BORDER_POLICY_IGNORE, fixed smoothing size) on organized clouds that are on
// the device — what DR-SLAM computes on the 1/3-resolution cloud of every frame (reference src/Frame.cc:1174-1216, :1057-1100):
//     ne.setNormalEstimationMethod(ne.AVERAGE_3D_GRADIENT); ne.setMaxDepthChangeFactor(0.05f); ne.setNormalSmoothingSize(10.0f);
// PCL is not vendored in the reference; this is the published algorithm of PCL 1.9 (features/impl/integral_image_normal.hpp,
// integral_image2D.hpp) as restated in the oracle's normals_oracle.cpp, whose header lists the steps.  Two of them are sequential in PCL
// and are kept sequential here, because their floating-point results depend on the order: the two raster passes of the chamfer
// distance transform (float additions of 1.0f / 1.4f along the scan) and the integral-image recurrence
// I(r+1, c+1) = I(r, c+1) + I(r+1, c) - I(r, c) + e in double.  A frame is 214 x 160 points, the passes are short, and frames run side by
// side — one CTA per frame; this is an optional post-processing call, not part of the timed step.
//   k_normals_distance   depth-change map (gather form), distance-map initialisation, the two raster passes: per row the three
//                        upper (lower) neighbours are combined by all threads, the left-to-right (right-to-left) chain by one thread
//   k_normals_integral   central differences on the fly, the two integral images (3 doubles each) and their finite counts: six + two
//                        threads walk a row's recurrences, the previous row stays in shared memory
//   k_normals_estimate   one thread per point: window from the distance map, four corners of both integral images, cross product,
//                        normalisation in double, flip towards the origin
#include <cfloat>
#include <cmath>

#include "drfe_internal.h"

namespace drfe {

static const int kNormThreads = 256;

__global__ void __launch_bounds__(kNormThreads) k_normals_distance(const float* __restrict__ cloud, int w, int h, float factor, float* __restrict__ dist_out) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* prev = reinterpret_cast<float*>(smem);           // [w + 1] the row above (below); [w] = the wrap element PCL reads
  float* cur = prev + (w + 1);                            // [w + 1]
  float* m3 = cur + (w + 1);                              // [w] min of the three neighbours in the other row
  const int f = blockIdx.x, tid = threadIdx.x;
  const float* C = cloud + (size_t)f * w * h * 3;
  float* D = dist_out + (size_t)f * w * h;
  // depth-change map in gather form: a point is marked by its own two tests, by the right-neighbour test of the point to its left and by
  // the lower-neighbour test of the point above (each test uses the threshold of ITS origin point); origins are rows 0..h-2, cols 0..w-2
  auto origin_marks = [&](int r, int c, bool right) {
    const size_t i = (size_t)r * w + c;
    const float depth = C[3 * i + 2], other = C[3 * (right ? i + 1 : i + w) + 2];
    const float th = __fmul_rn(__fmul_rn(factor, __fadd_rn(fabsf(depth), 1.0f)), 2.0f);
    return fabs((double)__fsub_rn(depth, other)) > (double)th || !isfinite(depth) || !isfinite(other);
  };
  for (int i = tid; i < w * h; i += kNormThreads) {
    const int r = i / w, c = i - r * w;
    bool mark = false;
    if (r < h - 1 && c < w - 1) mark = origin_marks(r, c, true) || origin_marks(r, c, false);
    if (!mark && c > 0 && r < h - 1) mark = origin_marks(r, c - 1, true);
    if (!mark && r > 0 && c < w - 1) mark = origin_marks(r - 1, c, false);
    D[i] = mark ? 0.0f : (float)(w + h);
  }
  __syncthreads();
  // ---- first pass: rows 1 .. h-1, columns 1 .. w-1; previous_row[w] is the first element of the current row
  for (int c = tid; c < w; c += kNormThreads) prev[c] = D[c];
  __syncthreads();
  for (int r = 1; r < h; ++r) {
    for (int c = tid; c < w; c += kNormThreads) cur[c] = D[(size_t)r * w + c];
    __syncthreads();
    if (tid == 0) prev[w] = cur[0];
    __syncthreads();
    for (int c = 1 + tid; c < w; c += kNormThreads)
      m3[c] = fminf(fminf(__fadd_rn(prev[c - 1], 1.4f), __fadd_rn(prev[c], 1.0f)), __fadd_rn(prev[c + 1], 1.4f));
    __syncthreads();
    if (tid == 0) {
      float left = cur[0];
      for (int c = 1; c < w; ++c) {
        const float center = cur[c];
        const float mv = fminf(m3[c], __fadd_rn(left, 1.0f));
        left = mv < center ? mv : center;
        cur[c] = left;
      }
    }
    __syncthreads();
    for (int c = tid; c < w; c += kNormThreads) { D[(size_t)r * w + c] = cur[c]; prev[c] = cur[c]; }
    __syncthreads();
  }
  // ---- second pass: rows h-2 .. 0, columns w-2 .. 0; next_row[-1] is the last element of the current row.  prev = next_row here,
  // stored shifted by one so that index -1 exists: prev[c + 1] = next_row[c], prev[0] = next_row[-1]
  for (int c = tid; c < w; c += kNormThreads) prev[c + 1] = D[(size_t)(h - 1) * w + c];
  __syncthreads();
  for (int r = h - 2; r >= 0; --r) {
    for (int c = tid; c < w; c += kNormThreads) cur[c] = D[(size_t)r * w + c];
    __syncthreads();
    if (tid == 0) prev[0] = cur[w - 1];
    __syncthreads();
    for (int c = tid; c < w - 1; c += kNormThreads)
      m3[c] = fminf(fminf(__fadd_rn(prev[c], 1.4f), __fadd_rn(prev[c + 1], 1.0f)), __fadd_rn(prev[c + 2], 1.4f));
    __syncthreads();
    if (tid == 0) {
      float right = cur[w - 1];
      for (int c = w - 2; c >= 0; --c) {
        const float center = cur[c];
        const float mv = fminf(m3[c], __fadd_rn(right, 1.0f));
        right = mv < center ? mv : center;
        cur[c] = right;
      }
    }
    __syncthreads();
    for (int c = tid; c < w; c += kNormThreads) { D[(size_t)r * w + c] = cur[c]; prev[c + 1] = cur[c]; }
    __syncthreads();
  }
}

// the element of DX (which == 0) or DY (which == 1) at (r, c): central difference on the interior, 0 on the image border
__device__ __forceinline__ float diff_elem(const float* __restrict__ C, int w, int h, int r, int c, int which, int k) {
  if (r < 1 || r >= h - 1 || c < 1 || c >= w - 1) return 0.f;
  const size_t i = (size_t)r * w + c;
  return which == 0 ? __fsub_rn(C[3 * (i + 1) + k], C[3 * (i - 1) + k]) : __fsub_rn(C[3 * (i + w) + k], C[3 * (i - w) + k]);
}

__global__ void __launch_bounds__(kNormThreads) k_normals_integral(const float* __restrict__ cloud, int w, int h, double* __restrict__ I, unsigned* __restrict__ Cn) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int W1 = w + 1;
  double* prev = reinterpret_cast<double*>(smem);         // [2][W1][3]
  double* cur = prev + (size_t)2 * W1 * 3;                // [2][W1][3]
  unsigned* cprev = reinterpret_cast<unsigned*>(cur + (size_t)2 * W1 * 3);   // [2][W1]
  unsigned* ccur = cprev + 2 * W1;
  float* el = reinterpret_cast<float*>(ccur + 2 * W1);    // [2][w][3] the row's elements
  const int f = blockIdx.x, tid = threadIdx.x;
  const float* C = cloud + (size_t)f * w * h * 3;
  double* If = I + (size_t)f * 2 * W1 * (h + 1) * 3;      // [which][h + 1][W1][3]
  unsigned* Cf = Cn + (size_t)f * 2 * W1 * (h + 1);
  for (int i = tid; i < 2 * W1 * 3; i += kNormThreads) prev[i] = 0.0;
  for (int i = tid; i < 2 * W1; i += kNormThreads) cprev[i] = 0u;
  for (int which = 0; which < 2; ++which)
    for (int i = tid; i < W1 * 3; i += kNormThreads) If[(size_t)which * W1 * (h + 1) * 3 + i] = 0.0;
  for (int which = 0; which < 2; ++which)
    for (int i = tid; i < W1; i += kNormThreads) Cf[(size_t)which * W1 * (h + 1) + i] = 0u;
  __syncthreads();
  for (int r = 0; r < h; ++r) {
    for (int i = tid; i < 2 * w * 3; i += kNormThreads) {
      const int which = i / (w * 3), rem = i - which * w * 3, c = rem / 3, k = rem - 3 * c;
      el[i] = diff_elem(C, w, h, r, c, which, k);
    }
    __syncthreads();
    if (tid < 6) {                                        // one thread per (image, component): the recurrence of a row
      const int which = tid / 3, k = tid - 3 * which;
      const double* pr = prev + (size_t)which * W1 * 3;
      double* cu = cur + (size_t)which * W1 * 3;
      const float* e = el + (size_t)which * w * 3;
      double left = 0.0;
      cu[k] = 0.0;
      for (int c = 0; c < w; ++c) {
        double v = __dsub_rn(__dadd_rn(pr[3 * (c + 1) + k], left), pr[3 * c + k]);
        const float s = __fadd_rn(__fadd_rn(e[3 * c], e[3 * c + 1]), e[3 * c + 2]);      // element->sum()
        if (isfinite(s)) v = __dadd_rn(v, (double)e[3 * c + k]);
        cu[3 * (c + 1) + k] = v;
        left = v;
      }
    } else if (tid >= 32 && tid < 34) {                   // the finite counts, on another warp
      const int which = tid - 32;
      const unsigned* pr = cprev + which * W1;
      unsigned* cu = ccur + which * W1;
      const float* e = el + (size_t)which * w * 3;
      unsigned left = 0u;
      cu[0] = 0u;
      for (int c = 0; c < w; ++c) {
        unsigned v = pr[c + 1] + left - pr[c];
        if (isfinite(__fadd_rn(__fadd_rn(e[3 * c], e[3 * c + 1]), e[3 * c + 2]))) ++v;
        cu[c + 1] = v;
        left = v;
      }
    }
    __syncthreads();
    for (int i = tid; i < 2 * W1 * 3; i += kNormThreads) {
      const int which = i / (W1 * 3), rem = i - which * W1 * 3;
      const double v = cur[i];
      If[((size_t)which * (h + 1) + (r + 1)) * W1 * 3 + rem] = v;
      prev[i] = v;
    }
    for (int i = tid; i < 2 * W1; i += kNormThreads) {
      const int which = i / W1, rem = i - which * W1;
      const unsigned v = ccur[i];
      Cf[((size_t)which * (h + 1) + (r + 1)) * W1 + rem] = v;
      cprev[i] = v;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_normals_estimate(const float* __restrict__ cloud, int w, int h, int nframes, float smoothing_size, const float* __restrict__ dist,
                                                          const double* __restrict__ I, const unsigned* __restrict__ Cn, float* __restrict__ normals) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)nframes * w * h) return;
  const int f = (int)(t / (w * h)), index = (int)(t - (long long)f * w * h);
  const int ri = index / w, ci = index - ri * w;
  const int W1 = w + 1;
  const float nanv = __int_as_float(0x7FC00000);
  float* out = normals + 3 * t;
  out[0] = out[1] = out[2] = nanv;
  const int border = (int)smoothing_size;
  if (ri < border || ri >= h - border || ci < border || ci >= w - border) return;
  const float* P = cloud + 3 * t;
  if (!isfinite(P[2])) return;
  const float smoothing = fminf(dist[t], smoothing_size);
  if (!(smoothing > 2.0f)) return;
  const int rw = (int)smoothing, rw2 = rw / 2;
  const int sx = ci - rw2, sy = ri - rw2;
  const size_t ul = (size_t)sy * W1 + sx, ur = ul + rw, ll = (size_t)(sy + rw) * W1 + sx, lr = ll + rw;
  const unsigned* Cf = Cn + (size_t)f * 2 * W1 * (h + 1);
  const unsigned* Cx = Cf;
  const unsigned* Cy = Cf + (size_t)W1 * (h + 1);
  if (Cx[lr] + Cx[ul] - Cx[ur] - Cx[ll] == 0u || Cy[lr] + Cy[ul] - Cy[ur] - Cy[ll] == 0u) return;
  const double* Ix = I + (size_t)f * 2 * W1 * (h + 1) * 3;
  const double* Iy = Ix + (size_t)W1 * (h + 1) * 3;
  double gx[3], gy[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    gx[k] = __dsub_rn(__dsub_rn(__dadd_rn(Ix[3 * lr + k], Ix[3 * ul + k]), Ix[3 * ur + k]), Ix[3 * ll + k]);
    gy[k] = __dsub_rn(__dsub_rn(__dadd_rn(Iy[3 * lr + k], Iy[3 * ul + k]), Iy[3 * ur + k]), Iy[3 * ll + k]);
  }
  const double nv0 = __dsub_rn(__dmul_rn(gy[1], gx[2]), __dmul_rn(gy[2], gx[1]));
  const double nv1 = __dsub_rn(__dmul_rn(gy[2], gx[0]), __dmul_rn(gy[0], gx[2]));
  const double nv2 = __dsub_rn(__dmul_rn(gy[0], gx[1]), __dmul_rn(gy[1], gx[0]));
  const double len2 = __dadd_rn(__dadd_rn(__dmul_rn(nv0, nv0), __dmul_rn(nv1, nv1)), __dmul_rn(nv2, nv2));
  if (len2 == 0.0) return;
  const double len = sqrt(len2);
  float nx = (float)(nv0 / len), ny = (float)(nv1 / len), nz = (float)(nv2 / len);
  const float vx = __fsub_rn(0.f, P[0]), vy = __fsub_rn(0.f, P[1]), vz = __fsub_rn(0.f, P[2]);
  const float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(vx, nx), __fmul_rn(vy, ny)), __fmul_rn(vz, nz));
  if (cos_theta < 0) { nx = -nx; ny = -ny; nz = -nz; }
  out[0] = nx; out[1] = ny; out[2] = nz;
}

// the 1/3-resolution cloud of Frame::ComputePlanes (Frame.cc:1044-1066) from a 16-bit depth batch: d = float(raw) * factor (imDepth),
// z = d > max_point_dist ? 0 : d, x = (n - cx) * z / fx in float arithmetic
__global__ void __launch_bounds__(256) k_third_cloud_u16(const uint16_t* __restrict__ depth, long long rs, long long fs, float factor, float fx, float fy, float cx, float cy,
                                                         int W, int H, int nframes, float max_point_dist, float* __restrict__ out) {
  const int w3 = (W + 2) / 3, h3 = (H + 2) / 3;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)nframes * w3 * h3) return;
  const int f = (int)(t / (w3 * h3));
  const int rem = (int)(t - (long long)f * w3 * h3);
  const int r = rem / w3, c = rem - r * w3;
  const int m = 3 * r, n = 3 * c;
  const float d = __fmul_rn((float)__ldg(depth + f * fs + m * rs + n), factor);
  const float z = d > max_point_dist ? 0.f : d;
  float* o = out + 3 * t;
  o[0] = __fdiv_rn(__fmul_rn(__fsub_rn((float)n, cx), z), fx);
  o[1] = __fdiv_rn(__fmul_rn(__fsub_rn((float)m, cy), z), fy);
  o[2] = z;
}
int third_cloud_u16_launch(cudaStream_t st, int nf, const uint16_t* depth, long long rs, long long fs, float factor, float fx, float fy, float cx, float cy, int W, int H,
                           float max_point_dist, float* out) {
  const unsigned blocks = (unsigned)(((size_t)((W + 2) / 3) * ((H + 2) / 3) * nf + 255) / 256);
  DRFE_LAUNCH(k_third_cloud_u16, blocks, 256, 0, st, depth, rs, fs, factor, fx, fy, cx, cy, W, H, nf, max_point_dist, out);
  return DRFE_OK;
}

int normals_scratch_alloc(NormalsScratch& s, size_t B, int w, int h) {
  const size_t W1 = (size_t)w + 1;
  if (cudaMalloc((void**)&s.dist, B * w * h * sizeof(float)) != cudaSuccess || cudaMalloc((void**)&s.I, B * 2 * W1 * (h + 1) * 3 * sizeof(double)) != cudaSuccess ||
      cudaMalloc((void**)&s.Cn, B * 2 * W1 * (h + 1) * sizeof(unsigned)) != cudaSuccess || cudaMalloc((void**)&s.normals, B * w * h * 3 * sizeof(float)) != cudaSuccess) {
    normals_scratch_free(s);
    return 1;
  }
  return 0;
}
void normals_scratch_free(NormalsScratch& s) {
  if (s.dist) cudaFree(s.dist);
  if (s.I) cudaFree(s.I);
  if (s.Cn) cudaFree(s.Cn);
  if (s.normals) cudaFree(s.normals);
  s = NormalsScratch();
}
int normals_launch(cudaStream_t st, int device, int nf, const float* cloud, int w, int h, float max_depth_change_factor, float smoothing_size, const NormalsScratch& s) {
  const size_t smem_d = (size_t)(3 * w + 2) * sizeof(float);
  const size_t smem_i = (size_t)4 * (w + 1) * 3 * sizeof(double) + (size_t)4 * (w + 1) * sizeof(unsigned) + (size_t)2 * w * 3 * sizeof(float);
  if (smem_i > 200 * 1024) { set_error("normals: a cloud row of %d points is too wide for the integral-image kernel", w); return DRFE_ERR_ARG; }
  if (raise_dyn_smem(k_normals_integral, device, smem_i) != cudaSuccess || raise_dyn_smem(k_normals_distance, device, smem_d) != cudaSuccess) {
    set_error("cudaFuncSetAttribute failed");
    return DRFE_ERR_CUDA;
  }
  DRFE_LAUNCH(k_normals_distance, nf, kNormThreads, smem_d, st, cloud, w, h, max_depth_change_factor, s.dist);
  DRFE_LAUNCH(k_normals_integral, nf, kNormThreads, smem_i, st, cloud, w, h, s.I, s.Cn);
  const unsigned blocks = (unsigned)(((size_t)nf * w * h + 255) / 256);
  DRFE_LAUNCH(k_normals_estimate, blocks, 256, 0, st, cloud, w, h, nf, smoothing_size, s.dist, s.I, s.Cn, s.normals);
  return DRFE_OK;
}

}  // namespace drfe
