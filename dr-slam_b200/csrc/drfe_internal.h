// Internal helpers shared by the drfe CUDA translation units (not installed).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only (the tools inject the implementation); a no-op without a profiler attached

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/drfe.h"

namespace drfe {

void set_error(const char* fmt, ...);          // thread-local message for drfe_last_error()
extern std::atomic<long long> g_launches;      // counted by DRFE_LAUNCH
static const int kPdlMaxFrames = 64;
bool pdl_enabled();
int pdl_max_frames();                          // kPdlMaxFrames, or DRFE_PDL_MAX_FRAMES                            // programmatic dependent launch in the batch chains (off with DRFE_NO_PDL=1)

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

// The same for the kernels of the per-batch chains (orb_launch_on, cape_launch), with programmatic dependent launch: the
// kernel may be scheduled while its predecessor in the stream is still running — its CTAs take the SMs the predecessor's
// last wave leaves idle and wait in DRFE_GRID_DEP() (first statement of every kernel launched this way) until the predecessor
// has completed and its writes are visible.  That hides the launch latency and the ramp-up of each of the chain's 13 + 7
// kernel boundaries.  Measured: pyramid (8 launches) 0.087 -> 0.072 ms at 32 frames, 0.052 -> 0.033 ms at one frame; at 256
// frames the waiting CTAs take SM slots from the OTHER handle's stream and the two-stream step gets 1.5 % slower — so the
// caller sets `drfe_pdl_` (a local the macro reads) only for launches of at most kPdlMaxFrames frames (DRFE_PDL_MAX_FRAMES overrides;
// with it on at 256 frames and the ORB stream prioritised: 1.84 -> 1.97 ms).  DRFE_NO_PDL=1: never.
#define DRFE_LAUNCH_PDL(kernel, grid_, block_, smem_, strm_, ...)                           \
  do {                                                                                    \
    cudaLaunchConfig_t cfg__ = {};                                                        \
    cfg__.gridDim = dim3(grid_); cfg__.blockDim = dim3(block_);                             \
    cfg__.dynamicSmemBytes = (smem_); cfg__.stream = (strm_);                             \
    cudaLaunchAttribute at__[1];                                                          \
    at__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                      \
    at__[0].val.programmaticStreamSerializationAllowed = drfe_pdl_ ? 1 : 0;               \
    cfg__.attrs = at__; cfg__.numAttrs = 1;                                               \
    cudaError_t e__ = cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__);                    \
    drfe::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
    if (e__ == cudaSuccess) e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                             \
      drfe::set_error("launch of %s failed: %s (%s:%d)", #kernel, cudaGetErrorString(e__), \
                      __FILE__, __LINE__);                                                \
      return DRFE_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)
#ifdef __CUDACC__
#define DRFE_GRID_DEP()                                                  \
  do {                                                                   \
    asm volatile("griddepcontrol.launch_dependents;");                   \
    asm volatile("griddepcontrol.wait;" ::: "memory");                   \
  } while (0)
#endif

// pcl::VoxelGrid::applyFilter on per-(frame, plane) point lists that are already on the device (k_voxel_sort / k_voxel_centroids in
// cape.cu), for the handles that produce such lists (CAPE's plane_cloud, PEAC's plane vertices): pts [nf][N][3], offs [nf][256]
// (first point of plane p; offs[np] = total), nplanes [nf].  The scratch holds the centroids (out, [nf][N][3]) and their offsets.
struct VoxelScratch {
  uint32_t* key[2] = {nullptr, nullptr};
  uint32_t* val[2] = {nullptr, nullptr};
  void* seg = nullptr;
  float* out = nullptr;
  int* out_offs = nullptr;
};
int voxel_scratch_alloc(VoxelScratch& s, size_t max_frames, size_t N);    // cudaMalloc on the current device; 0 on success
void voxel_scratch_free(VoxelScratch& s);
int voxel_filter_launch(cudaStream_t st, int nf, const float* pts, const int* offs, const int* nplanes, int N, float leaf_size, const VoxelScratch& s);

// pcl::IntegralImageNormalEstimation (AVERAGE_3D_GRADIENT) on organized device clouds [nf][h][w][3] (normals.cu)
struct NormalsScratch {
  float* dist = nullptr;        // [nf][h][w] distance map
  double* I = nullptr;          // [nf][2][h + 1][w + 1][3] integral images of the x- and y-differences
  unsigned* Cn = nullptr;       // [nf][2][h + 1][w + 1] finite counts
  float* normals = nullptr;     // [nf][h][w][3]
};
int normals_scratch_alloc(NormalsScratch& s, size_t max_frames, int w, int h);
void normals_scratch_free(NormalsScratch& s);
int third_cloud_u16_launch(cudaStream_t st, int nf, const uint16_t* depth, long long rs, long long fs, float factor, float fx, float fy, float cx, float cy, int W, int H,
                           float max_point_dist, float* out);
int normals_launch(cudaStream_t st, int device, int nf, const float* cloud, int w, int h, float max_depth_change_factor, float smoothing_size, const NormalsScratch& s);

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to (kernel, device), not to a handle, and the last setter
// wins: two live handles of different geometry would shrink each other's limit.  raise_dyn_smem keeps a
// process-wide high-water mark per (kernel, device) under a mutex and only ever raises the attribute.  The
// current device must be `device`.
cudaError_t raise_dyn_smem_impl(const void* func, int device, size_t bytes);
template <typename K>
inline cudaError_t raise_dyn_smem(K* kernel, int device, size_t bytes) { return raise_dyn_smem_impl((const void*)kernel, device, bytes); }

// NVTX range for the scope: host-side spans of the entry points and of every stage's launches, so that an Nsight Systems
// timeline of an application shows which drfe call / stage a kernel belongs to
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

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

// per-stage device timers: cudaEvent pairs on the handle's stream, kept for the last
// kSlots enqueues so a bench can read per-kernel averages AFTER its timed region without
// synchronising inside it.
struct StageTimer {
  static const int kMax = 12, kSlots = 64;
  cudaEvent_t ev[kSlots][kMax + 1];
  const char* names[kMax];
  int n = 0, slot = -1, filled = 0;
  bool enabled = false, created = false;
  int create() {
    for (int s = 0; s < kSlots; ++s)
      for (int i = 0; i <= kMax; ++i) DRFE_CUDA(cudaEventCreate(&ev[s][i]));
    created = true;
    return DRFE_OK;
  }
  void destroy() {
    if (created)
      for (int s = 0; s < kSlots; ++s)
        for (int i = 0; i <= kMax; ++i) cudaEventDestroy(ev[s][i]);
    created = false;
  }
  void reset(bool on) { enabled = on; slot = -1; filled = 0; n = 0; }
  void begin(cudaStream_t s) {
    if (!enabled) return;
    slot = (slot + 1) % kSlots;
    if (filled < kSlots) ++filled;
    n = 0;
    cudaEventRecord(ev[slot][0], s);
  }
  void mark(const char* name, cudaStream_t s) {
    if (!enabled || slot < 0 || n >= kMax) return;
    names[n] = name;
    cudaEventRecord(ev[slot][n + 1], s);
    ++n;
  }
  // average over the recorded enqueues (all must have run the same stage sequence)
  int read(float* ms, const char** out_names, int cap, int* nstages) {
    *nstages = 0;
    if (!enabled || filled == 0) return DRFE_OK;
    const int m = n < cap ? n : cap;
    for (int i = 0; i < m; ++i) ms[i] = 0.f;
    for (int s = 0; s < filled; ++s) {
      DRFE_CUDA(cudaEventSynchronize(ev[s][n]));
      for (int i = 0; i < m; ++i) {
        float t = 0.f;
        DRFE_CUDA(cudaEventElapsedTime(&t, ev[s][i], ev[s][i + 1]));
        ms[i] += t;
      }
    }
    for (int i = 0; i < m; ++i) {
      ms[i] /= (float)filled;
      if (out_names) out_names[i] = names[i];
    }
    *nstages = m;
    return DRFE_OK;
  }
};

// Streams and events of the chunk-pipelined batch calls (drfe_orb_extract_batch,
// drfe_cape_process_depth_batch): chunk k's host->device copy runs on `h2d`, its kernels on the
// handle's stream, its device->host copies on `d2h`, so that the three overlap across chunks.
struct ChunkPipe {
  static const int kMaxChunks = 64;
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t ev_in[kMaxChunks], ev_done[kMaxChunks], ev_start = nullptr;
  int* h_status = nullptr;   // pinned copy of the handle's device status word
  bool created = false, active = false;
  int create() {
    if (created) return DRFE_OK;
    DRFE_CUDA(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
    DRFE_CUDA(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
    DRFE_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    for (int i = 0; i < kMaxChunks; ++i) {
      DRFE_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
      DRFE_CUDA(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
    }
    DRFE_CUDA(cudaHostAlloc((void**)&h_status, sizeof(int), cudaHostAllocDefault));
    *h_status = 0;
    created = true;
    return DRFE_OK;
  }
  void destroy() {
    if (!created) return;
    cudaStreamSynchronize(h2d); cudaStreamSynchronize(d2h);
    for (int i = 0; i < kMaxChunks; ++i) { cudaEventDestroy(ev_in[i]); cudaEventDestroy(ev_done[i]); }
    cudaEventDestroy(ev_start);
    cudaStreamDestroy(h2d); cudaStreamDestroy(d2h);
    cudaFreeHost(h_status);
    created = false;
  }
  // chunk boundaries start[0 .. n]: the caller's wish (> 0: uniform chunks of that many frames, never more than kMaxChunks), or
  // (<= 0) 32-frame chunks with a ramp at both ends (8, 16, 32 ... 32, 16, 8) when the batch is long enough: the first kernels
  // start after a quarter of a chunk's H2D copy and the last chunk's kernels + D2H copy (the part nothing overlaps) are short
  static int schedule(int nframes, int wish, int* start) {
    int n = 0, f = 0;
    if (wish <= 0 && nframes >= 128) {
      const int head[2] = {8, 16};
      for (int i = 0; i < 2; ++i) { start[n++] = f; f += head[i]; }
      const int mid_end = nframes - 24;
      int mid = 32;                                              // larger middle chunks on very long batches: at most kMaxChunks chunks
      while ((mid_end - 24 + mid - 1) / mid > kMaxChunks - 4) ++mid;
      while (f < mid_end) { start[n++] = f; f += (mid_end - f < mid) ? mid_end - f : mid; }
      start[n++] = f; f += 16;
      start[n++] = f; f += 8;
    } else {
      int c = wish > 0 ? wish : 32;
      if (c > nframes) c = nframes;
      while ((nframes + c - 1) / c > kMaxChunks) ++c;
      for (; f < nframes; f += c) start[n++] = f;
    }
    start[n] = nframes;
    return n;
  }
};

// what bow.cu needs of an ORB handle's last batch (defined in orb.cu, where struct drfe_orb lives)
struct OrbBatchView {
  int device, nframes, cap;
  bool pending;
  cudaStream_t stream;
  const uint8_t* desc;  // [B][cap][32]
  const int* cnt;       // [B]
  const drfe_keypoint* kp;  // [B][cap] mvKeys
};
int orb_batch_view(drfe_orb* h, OrbBatchView* v);

// the depth images a CAPE handle's last enqueue / batch call left on the device (defined in cape.cu)
struct CapeDepthView {
  int device, nframes, width, height;
  bool pending;
  cudaStream_t stream;
  const float* depth;        // float metres, or
  const uint16_t* depth16;   // raw sensor depth, z = (float)u16 * factor
  float factor;
  long long row_stride, frame_stride;   // in elements
};
int cape_depth_view(drfe_cape* h, CapeDepthView* v);

}  // namespace drfe
