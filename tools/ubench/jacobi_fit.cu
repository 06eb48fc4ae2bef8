// Micro-benchmark: one PEAC plane fit (Stats::compute: scatter matrix, cyclic Jacobi, normal, mse) per lane of one warp alone on an
// SM, classic rotation (theta / t / c chain) against the rotation from r = sqrt(h^2 + 4 apq^2).  Prints cycles per fit and the
// number of rotations.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o jacobi_fit jacobi_fit.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

template <int V>
__device__ int eig3(const double in[6], double w[3], double v[3][3]) {
  double a[3][3] = {{in[0], in[1], in[2]}, {in[1], in[3], in[4]}, {in[2], in[4], in[5]}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  int rot = 0;
  for (int sweep = 0; sweep < 24; ++sweep) {
    if (a[0][1] == 0.0 && a[0][2] == 0.0 && a[1][2] == 0.0) break;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int p = (k == 2) ? 1 : 0, q = (k == 0) ? 1 : 2, r = (k == 0) ? 2 : ((k == 1) ? 1 : 0);
      const double apq = a[p][q];
      if (apq == 0.0) continue;
      const double app = a[p][p], aqq = a[q][q];
      const double g = 100.0 * fabs(apq);
      if (sweep > 2 && fabs(app) + g == fabs(app) && fabs(aqq) + g == fabs(aqq)) { a[p][q] = a[q][p] = 0.0; continue; }
      ++rot;
      const double h = aqq - app;
      double t, c;
      if (fabs(h) + g == fabs(h)) { t = apq / h; c = 1.0 / sqrt(1.0 + t * t); }
      else if (V == 0) {
        const double theta = 0.5 * h / apq;
        t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
        if (theta < 0.0) t = -t;
        c = 1.0 / sqrt(1.0 + t * t);
      } else {
        const double rr = sqrt(h * h + 4.0 * apq * apq);
        t = 2.0 * apq / (h >= 0.0 ? h + rr : h - rr);
        c = sqrt(0.5 + 0.5 * (fabs(h) / rr));
      }
      const double s = t * c;
      a[p][p] = app - t * apq; a[q][q] = aqq + t * apq; a[p][q] = a[q][p] = 0.0;
      const double arp = a[r][p], arq = a[r][q];
      a[r][p] = a[p][r] = c * arp - s * arq;
      a[r][q] = a[q][r] = s * arp + c * arq;
#pragma unroll
      for (int m = 0; m < 3; ++m) { const double vp = v[m][p], vq = v[m][q]; v[m][p] = c * vp - s * vq; v[m][q] = s * vp + c * vq; }
    }
  }
  w[0] = a[0][0]; w[1] = a[1][1]; w[2] = a[2][2];
  return rot;
}

template <int V>
__global__ void k(const double* K6, double* out, long long* cyc, int* rots) {
  double in[6], w[3], v[3][3];
  for (int i = 0; i < 6; ++i) in[i] = K6[i] * (1.0 + 1e-3 * threadIdx.x);
  const long long t0 = clock64();
  const int r = eig3<V>(in, w, v);
  const long long t1 = clock64();
  out[threadIdx.x] = w[0] + w[1] + w[2] + v[0][0] + v[1][1] + v[2][2];
  if (threadIdx.x == 0) { cyc[V] = t1 - t0; rots[V] = r; }
}

int main() {
  // scatter matrix of 100 points of a tilted plane patch: extents 3 cm x 3 cm, noise 2 mm
  double h[6], *d, *out; long long* cyc; int* rots;
  double pts[100][3]; unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) / 16777216.0 - 0.5; };
  double sx[3] = {0, 0, 0}, sxx[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 100; ++i) {
    const double u = 0.03 * rnd(), vv = 0.03 * rnd(), n = 0.002 * rnd();
    pts[i][0] = u + 0.3 * n; pts[i][1] = vv - 0.2 * n; pts[i][2] = 2.0 + 0.4 * u - 0.3 * vv + n;
    for (int a = 0; a < 3; ++a) sx[a] += pts[i][a];
    sxx[0] += pts[i][0] * pts[i][0]; sxx[1] += pts[i][0] * pts[i][1]; sxx[2] += pts[i][0] * pts[i][2];
    sxx[3] += pts[i][1] * pts[i][1]; sxx[4] += pts[i][1] * pts[i][2]; sxx[5] += pts[i][2] * pts[i][2];
  }
  h[0] = sxx[0] - sx[0] * sx[0] / 100; h[1] = sxx[1] - sx[0] * sx[1] / 100; h[2] = sxx[2] - sx[0] * sx[2] / 100;
  h[3] = sxx[3] - sx[1] * sx[1] / 100; h[4] = sxx[4] - sx[1] * sx[2] / 100; h[5] = sxx[5] - sx[2] * sx[2] / 100;
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 16); cudaMallocManaged(&rots, 8);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; ++rep) { k<0><<<1, 32>>>(d, out, cyc, rots); k<1><<<1, 32>>>(d, out, cyc, rots); cudaDeviceSynchronize(); }
  printf("classic rotation: %lld cycles per fit, %d rotations (%.0f cycles each)\n", cyc[0], rots[0], (double)cyc[0] / rots[0]);
  printf("sqrt-first rotation: %lld cycles per fit, %d rotations (%.0f cycles each)\n", cyc[1], rots[1], (double)cyc[1] / rots[1]);
  return 0;
}
