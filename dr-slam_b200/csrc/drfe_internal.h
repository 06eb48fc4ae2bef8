// Internal helpers shared by the drfe CUDA translation units (not installed).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/drfe.h"

namespace drfe {

void set_error(const char* fmt, ...);          // thread-local message for drfe_last_error()
extern std::atomic<long long> g_launches;      // counted by DRFE_LAUNCH

#define DRFE_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      drfe::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,  \
                      __LINE__);                                                          \
      return DRFE_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

// kernel<<<grid, block, smem, stream>>>(args...) + launch counting + launch-error check
#define DRFE_LAUNCH(kernel, grid, block, smem, stream, ...)                               \
  do {                                                                                    \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                           \
    drfe::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      drfe::set_error("launch of %s failed: %s (%s:%d)", #kernel, cudaGetErrorString(e__), \
                      __FILE__, __LINE__);                                                \
      return DRFE_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

// RAII "make this device current for the scope" (handles are callable from any thread)
struct DeviceScope {
  int prev = -1;
  bool ok = true;
  explicit DeviceScope(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceScope() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

// per-stage device timers (cudaEvent pairs on the handle's stream)
struct StageTimer {
  static const int kMax = 16;
  cudaEvent_t ev[kMax + 1];
  const char* names[kMax];
  int n = 0;
  bool enabled = false, created = false, valid = false;
  int create() {
    for (int i = 0; i <= kMax; ++i) DRFE_CUDA(cudaEventCreate(&ev[i]));
    created = true;
    return DRFE_OK;
  }
  void destroy() {
    if (created)
      for (int i = 0; i <= kMax; ++i) cudaEventDestroy(ev[i]);
    created = false;
  }
  void begin(cudaStream_t s) {
    n = 0; valid = false;
    if (enabled) cudaEventRecord(ev[0], s);
  }
  void mark(const char* name, cudaStream_t s) {
    if (!enabled || n >= kMax) return;
    names[n] = name;
    cudaEventRecord(ev[n + 1], s);
    ++n;
    valid = true;
  }
  int read(float* ms, const char** out_names, int cap, int* nstages) {
    *nstages = 0;
    if (!enabled || !valid) return DRFE_OK;
    DRFE_CUDA(cudaEventSynchronize(ev[n]));
    for (int i = 0; i < n && i < cap; ++i) {
      DRFE_CUDA(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
      if (out_names) out_names[i] = names[i];
    }
    *nstages = n < cap ? n : cap;
    return DRFE_OK;
  }
};

}  // namespace drfe
