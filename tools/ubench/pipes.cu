// Micro-benchmark: which issue pipe do the packed min/max candidates of the FAST score network use on sm_100a?
// Each test runs 8 independent dependency chains per thread, 1024 iterations, 148*8 CTAs of 256 threads.
// Prints warp-instructions per clock per SM sub-partition (1.0 = the issue limit, 0.5 = a half-rate pipe).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 1024

template <int T>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
  uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { a[i] = seed * (threadIdx.x + i + 1); b[i] = seed ^ (0x64646464u + i); }
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (T == 0) a[i] = __vimax3_u16x2(a[i], b[i], b[(i + 1) % CHAINS]);
      if (T == 1) { __half2 x = *reinterpret_cast<__half2*>(&a[i]), y = *reinterpret_cast<__half2*>(&b[i]); x = __hmax2(x, y); a[i] = *reinterpret_cast<uint32_t*>(&x); b[i] += 1; }
      if (T == 2) { if (i & 1) { __half2 x = *reinterpret_cast<__half2*>(&a[i]), y = *reinterpret_cast<__half2*>(&b[i]); x = __hmax2(x, y); a[i] = *reinterpret_cast<uint32_t*>(&x); } else a[i] = __vimax3_u16x2(a[i], b[i], b[(i + 1) % CHAINS]); }
      if (T == 3) a[i] = __vabsdiffu4(a[i], b[i]) + 0;
      if (T == 4) a[i] = __funnelshift_r(a[i], b[i], 8);
      if (T == 5) a[i] = __byte_perm(a[i], b[i], 0x4240 + i);
      if (T == 6) a[i] = a[i] * 3 + b[i];
      if (T == 7) { if (i & 1) a[i] = a[i] * 3 + b[i]; else a[i] = __vimax3_u16x2(a[i], b[i], b[(i + 1) % CHAINS]); }
      if (T == 8) a[i] = __vmaxu2(a[i], b[i]) ^ 1;
      if (T == 9) { if (i & 1) a[i] = __vabsdiffu4(a[i], b[i]); else a[i] = __vimax3_u16x2(a[i], b[i], b[(i + 1) % CHAINS]); }
      if (T == 10) { float x = __uint_as_float(a[i]), y = __uint_as_float(b[i]); a[i] = __float_as_uint(fmaxf(x, y)); b[i] ^= a[i]; }
      if (T == 11) { asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }
      if (T == 12) { if (i & 1) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); else a[i] = __vimax3_u16x2(a[i], b[i], b[(i + 1) % CHAINS]); }
      if (T == 13) { asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }
      if (T == 14) a[i] = __dp4a(a[i], b[i], a[i]);
      if (T == 15) a[i] = __dp2a_lo(a[i], b[i], a[i]);
      if (T == 16) a[i] = __umulhi(a[i], b[i]) + 1u;
      if (T == 17) { if (i & 1) a[i] = __dp4a(a[i], b[i], a[i]); else a[i] = __vimax3_u16x2(a[i], b[i], b[(i + 1) % CHAINS]); }
      if (T == 18) { if (i & 1) a[i] = __dp4a(a[i], b[i], a[i]); else a[i] = a[i] * 3 + b[i]; }
      if (T == 19) { float x = __uint_as_float(a[i]), y = __uint_as_float(b[i]); a[i] = __float_as_uint(__fmaf_rn(x, y, x)); }
      if (T == 20) { if (i & 1) { float x = __uint_as_float(a[i]), y = __uint_as_float(b[i]); a[i] = __float_as_uint(__fmaf_rn(x, y, x)); } else a[i] = __vimax3_u16x2(a[i], b[i], b[(i + 1) % CHAINS]); }
      if (T == 21) a[i] = (a[i] << 3) + b[i];            // LEA
      if (T == 22) a[i] = __popc(a[i]) + b[i];
      if (T == 23) a[i] = (uint32_t)__float2int_rn(__uint_as_float(a[i])) + b[i];
      if (T == 24) { if (i & 1) { float x = __uint_as_float(a[i]), y = __uint_as_float(b[i]); a[i] = __float_as_uint(__fmaf_rn(x, y, x)); } else a[i] = a[i] * 3 + b[i]; }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s ^= a[i] ^ b[i];
  if (s == 0x12345u) out[0] = s;
}

template <int T>
static void run(const char* name, int ops_per_chain_iter) {
  uint32_t* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<T><<<148 * 8, 256>>>(d, 12345u); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<T><<<148 * 8, 256>>>(d, 12345u);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double winst = 148.0 * 8 * 8 * CHAINS * ITERS * ops_per_chain_iter;   // warp instructions
  const double cyc = ms * 1e-3 * clk * 1e3;
  printf("%-34s %8.3f ms  %.3f warp-inst/clk/SMSP (counting %d op per chain-iteration, nominal clock %d kHz)\n", name, ms, winst / cyc / (148 * 4), ops_per_chain_iter, clk);
  cudaFree(d);
}

int main() {
  run<0>("VIMNMX3.U16x2", 1);
  run<1>("HMNMX2 (+IADD)", 2);
  run<2>("VIMNMX3 / HMNMX2 alternating", 1);
  run<3>("VABSDIFF4", 1);
  run<4>("SHF funnel", 1);
  run<5>("PRMT", 1);
  run<6>("IMAD", 1);
  run<7>("VIMNMX3 / IMAD alternating", 1);
  run<8>("VIMNMX.U16x2 + LOP", 2);
  run<9>("VIMNMX3 / VABSDIFF4 alternating", 1);
  run<10>("FMNMX + LOP", 2);
  run<11>("max.f16x2 asm", 1);
  run<12>("VIMNMX3 / max.f16x2 alternating", 1);
  run<13>("max.bf16x2 asm", 1);
  run<14>("IDP.4A", 1);
  run<15>("IDP.2A", 1);
  run<16>("IMAD.HI (+IADD)", 2);
  run<17>("VIMNMX3 / IDP.4A alternating", 1);
  run<18>("IMAD / IDP.4A alternating", 1);
  run<19>("FFMA", 1);
  run<20>("VIMNMX3 / FFMA alternating", 1);
  run<21>("LEA (shift-add)", 1);
  run<22>("POPC + IADD", 2);
  run<23>("F2I + IADD", 2);
  run<24>("IMAD / FFMA alternating", 1);
  return 0;
}
