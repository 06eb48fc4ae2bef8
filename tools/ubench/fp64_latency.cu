// Micro-benchmark: latency of the FP64 operations a Jacobi rotation is a chain of, on sm_100a, one warp alone on an SM
// (the situation of PEAC's candidate fits): dependent DADD / DMUL / DFMA, IEEE double division, IEEE double sqrt, and
// one whole 3x3 Jacobi fit as peac.cu runs it.  Prints cycles per operation (clock64 around a chain of 256).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int T>
__global__ void k(double* out, long long* cyc, double a, double b) {
  double x = a + threadIdx.x * 1e-9, y = b, z0 = x + 1, z1 = x + 2, z2 = x + 3;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (T == 0) x = x + y;
      if (T == 1) x = x * y;
      if (T == 2) x = __fma_rn(x, y, y);
      if (T == 3) x = y / x + 1.0;          // the + 1.0 keeps the value in range; its cost is measured by T == 0
      if (T == 4) x = sqrt(x) + 1.0;
      if (T == 5) x = 1.0 / sqrt(1.0 + x * x);
    }
    if (T == 6) {                            // four INDEPENDENT chains: what one warp can issue, not what one chain waits for
      x = __fma_rn(x, y, y); z0 = __fma_rn(z0, y, y); z1 = __fma_rn(z1, y, y); z2 = __fma_rn(z2, y, y);
      x = __fma_rn(x, y, y); z0 = __fma_rn(z0, y, y); z1 = __fma_rn(z1, y, y); z2 = __fma_rn(z2, y, y);
      x = __fma_rn(x, y, y); z0 = __fma_rn(z0, y, y); z1 = __fma_rn(z1, y, y); z2 = __fma_rn(z2, y, y);
      x = __fma_rn(x, y, y); z0 = __fma_rn(z0, y, y); z1 = __fma_rn(z1, y, y); z2 = __fma_rn(z2, y, y);
    }
  }
  x += z0 + z1 + z2;
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[T] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * sizeof(double)); cudaMallocManaged(&cyc, 8 * sizeof(long long));
  k<0><<<1, 32>>>(out, cyc, 1.0, 1e-3); k<1><<<1, 32>>>(out, cyc, 1.0, 1.0000001); k<2><<<1, 32>>>(out, cyc, 1.0, 0.5);
  k<3><<<1, 32>>>(out, cyc, 1.5, 0.7); k<4><<<1, 32>>>(out, cyc, 1.5, 0.7); k<5><<<1, 32>>>(out, cyc, 1.5, 0.7); k<6><<<1, 32>>>(out, cyc, 1.0, 0.5);
  cudaDeviceSynchronize();
  const char* names[6] = {"DADD", "DMUL", "DFMA", "div + add", "sqrt + add", "1/sqrt(1+x*x)"};
  for (int t = 0; t < 6; ++t) printf("%-14s %6.1f cycles per dependent operation\n", names[t], cyc[t] / 256.0);
  printf("4 independent DFMA chains, one warp: %.1f cycles per DFMA issued (1024 DFMA)\n", cyc[6] / 1024.0);
  return 0;
}
