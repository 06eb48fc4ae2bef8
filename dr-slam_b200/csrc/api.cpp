// drfe C ABI: process-wide pieces (error text, version, launch counter).
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>

#include "drfe_internal.h"

namespace drfe {

cudaError_t raise_dyn_smem_impl(const void* func, int device, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> high;
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = high[std::make_pair(func, device)];
  if (bytes <= cur) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

static thread_local char t_err[512] = "";
std::atomic<long long> g_launches{0};
int pdl_max_frames() {
  static const int v = [] { const char* e = getenv("DRFE_PDL_MAX_FRAMES"); return e ? atoi(e) : kPdlMaxFrames; }();
  return v;
}
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("DRFE_NO_PDL"); return !(e && e[0] == '1'); }();
  return on;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

}  // namespace drfe

extern "C" {

const char* drfe_last_error(void) { return drfe::t_err; }
const char* drfe_version(void) { return "drfe 0.1 (sm_100a)"; }
int64_t drfe_kernel_launch_count(void) { return (int64_t)drfe::g_launches.load(); }
int drfe_device_count(int* count) {
  if (!count) return DRFE_ERR_ARG;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    drfe::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return DRFE_ERR_CUDA;
  }
  *count = n;
  return DRFE_OK;
}


// ---- device timers for callers that drive several handles (bench): CUDA events on the
// handles' own streams.
int drfe_event_create(void** ev) {
  if (!ev) return DRFE_ERR_ARG;
  cudaEvent_t e;
  DRFE_CUDA(cudaEventCreate(&e));
  *ev = (void*)e;
  return DRFE_OK;
}
int drfe_event_destroy(void* ev) {
  if (ev) DRFE_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return DRFE_OK;
}
int drfe_event_record(void* ev, void* stream) {
  DRFE_CUDA(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream));
  return DRFE_OK;
}
int drfe_stream_wait_event(void* stream, void* ev) {
  DRFE_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0));
  return DRFE_OK;
}
int drfe_event_elapsed_ms(void* start, void* stop, float* ms) {
  if (!ms) return DRFE_ERR_ARG;
  DRFE_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
  DRFE_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return DRFE_OK;
}

}  // extern "C"
