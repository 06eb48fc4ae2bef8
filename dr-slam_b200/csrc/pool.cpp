// drfe multi-device front end: the frames of one batch are independent (ORBextractor keeps no
// state across frames, CAPE is created per frame — Frame.cc:124-134, PlaneExtractor.cpp:149), so a batch is cut into
// contiguous blocks, one per device; every device has its own ORB + CAPE handle pair (own streams, own staging
// arenas) driven by its own persistent host thread through the chunk-pipelined batch calls, and the results land
// in the caller's arrays by frame index.  No collective, no peer traffic: the only shared resource is the host link.
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "drfe_internal.h"

namespace {

struct PoolJob {                   // one drfe_pool_extract_batch call, as every worker sees it
  int nframes = 0;
  const uint8_t* gray = nullptr; size_t gray_rs = 0, gray_fs = 0;
  const void* depth = nullptr; int depth_is_u16 = 0; float depth_factor = 1.f; size_t depth_rs = 0, depth_fs = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0;
  drfe_keypoint* kps = nullptr; uint8_t* desc = nullptr; int cap_per_frame = 0; int* counts = nullptr;
  uint8_t* seg = nullptr; drfe_plane* planes = nullptr; int plane_cap = 0; int* nr_planes = nullptr;
  drfe_cylinder* cyls = nullptr; int cyl_cap = 0; int* nr_cyls = nullptr;
  int chunk_frames = 0;
};

struct PoolWorker {
  int device = 0, index = 0;
  drfe_orb* orb = nullptr;
  drfe_cape* cape = nullptr;
  std::thread thread;
  int f0 = 0, f1 = 0;              // this worker's block of the current job
  int rc = DRFE_OK;
  std::string err;
  float ms = 0.f;                  // device time of the block (first H2D to last D2H), CUDA events
};

}  // namespace

struct drfe_pool {
  drfe_pool_params prm{};
  std::vector<PoolWorker> workers;
  int frames_per_device = 0;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  long long generation = 0;        // bumped per job
  int remaining = 0;
  bool quit = false;
  PoolJob job;
};

namespace {

void worker_run(drfe_pool* p, PoolWorker* w) {
  const PoolJob& J = p->job;
  const int n = w->f1 - w->f0;
  w->rc = DRFE_OK; w->err.clear(); w->ms = 0.f;
  if (n <= 0) return;
  drfe::NvtxRange nvtx_("drfe_pool worker block");
  const size_t f0 = (size_t)w->f0;
  const size_t desz = J.depth_is_u16 ? sizeof(uint16_t) : sizeof(float);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaSetDevice(w->device);
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, (cudaStream_t)drfe_orb_stream(w->orb));
  int rc = drfe_orb_extract_batch(w->orb, n, J.gray + f0 * J.gray_fs, J.gray_rs, J.gray_fs, J.kps ? J.kps + f0 * J.cap_per_frame : nullptr,
                                  J.desc ? J.desc + f0 * J.cap_per_frame * 32 : nullptr, J.cap_per_frame, J.counts + f0, J.chunk_frames);
  bool orb_live = rc == DRFE_OK;
  bool cape_live = false;
  if (rc == DRFE_OK) {
    rc = drfe_cape_process_depth_batch(w->cape, n, (const uint8_t*)J.depth + f0 * J.depth_fs * desz, J.depth_is_u16, J.depth_factor, J.depth_rs, J.depth_fs,
                                       J.fx, J.fy, J.cx, J.cy, J.seg ? J.seg + f0 * (size_t)p->prm.width * p->prm.height : nullptr,
                                       J.planes ? J.planes + f0 * J.plane_cap : nullptr, J.plane_cap, J.nr_planes + f0,
                                       J.cyls ? J.cyls + f0 * J.cyl_cap : nullptr, J.cyl_cap, J.nr_cyls ? J.nr_cyls + f0 : nullptr, J.chunk_frames);
    cape_live = rc == DRFE_OK;
  }
  if (rc != DRFE_OK) w->err = drfe_last_error();
  // both batches are always finished, so that nothing stays queued on the caller's buffers
  if (orb_live) { const int r = drfe_orb_finish_batch(w->orb); if (r != DRFE_OK && rc == DRFE_OK) { rc = r; w->err = drfe_last_error(); } }
  if (cape_live) { const int r = drfe_cape_finish_batch(w->cape); if (r != DRFE_OK && rc == DRFE_OK) { rc = r; w->err = drfe_last_error(); } }
  cudaEventRecord(e1, (cudaStream_t)drfe_orb_stream(w->orb));
  if (cudaEventSynchronize(e1) == cudaSuccess) cudaEventElapsedTime(&w->ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  w->rc = rc;
}

void worker_main(drfe_pool* p, PoolWorker* w) {
  long long seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lock(p->mu);
      p->cv_go.wait(lock, [&] { return p->quit || p->generation != seen; });
      if (p->quit) return;
      seen = p->generation;
    }
    worker_run(p, w);
    {
      std::lock_guard<std::mutex> lock(p->mu);
      if (--p->remaining == 0) p->cv_done.notify_all();
    }
  }
}

}  // namespace

extern "C" {

int drfe_pool_create(const drfe_pool_params* prm, const int* devices, int ndevices, drfe_pool** out) {
  if (!prm || !devices || !out || ndevices < 1 || ndevices > 64 || prm->max_batch < 1) { drfe::set_error("drfe_pool_create: bad argument"); return DRFE_ERR_ARG; }
  *out = nullptr;
  drfe_pool* p = new drfe_pool();
  p->prm = *prm;
  p->frames_per_device = (prm->max_batch + ndevices - 1) / ndevices;
  p->workers.resize(ndevices);
  drfe_cape_params cp = prm->cape;
  cp.depth_width = prm->width; cp.depth_height = prm->height;
  for (int i = 0; i < ndevices; ++i) {
    PoolWorker& w = p->workers[i];
    w.device = devices[i]; w.index = i;
    int rc = drfe_orb_create(&prm->orb, prm->width, prm->height, p->frames_per_device, w.device, &w.orb);
    if (rc == DRFE_OK) rc = drfe_cape_create(&cp, p->frames_per_device, w.device, &w.cape);
    if (rc != DRFE_OK) { drfe_pool_destroy(p); return rc; }
  }
  for (int i = 0; i < ndevices; ++i) p->workers[i].thread = std::thread(worker_main, p, &p->workers[i]);
  *out = p;
  return DRFE_OK;
}

int drfe_pool_destroy(drfe_pool* p) {
  if (!p) return DRFE_OK;
  {
    std::lock_guard<std::mutex> lock(p->mu);
    p->quit = true;
  }
  p->cv_go.notify_all();
  for (PoolWorker& w : p->workers) {
    if (w.thread.joinable()) w.thread.join();
    drfe_orb_destroy(w.orb);
    drfe_cape_destroy(w.cape);
  }
  delete p;
  return DRFE_OK;
}

int drfe_pool_num_devices(const drfe_pool* p) { return p ? (int)p->workers.size() : 0; }
int drfe_pool_max_keypoints(const drfe_pool* p) { return (p && !p->workers.empty()) ? drfe_orb_max_keypoints(p->workers[0].orb) : 0; }

int drfe_pool_extract_batch(drfe_pool* p, int nframes, const uint8_t* gray, size_t gray_row_stride, size_t gray_frame_stride, const void* depth,
                            int depth_is_u16, float depth_factor, size_t depth_row_stride, size_t depth_frame_stride, float fx, float fy, float cx,
                            float cy, drfe_keypoint* kps, uint8_t* desc, int cap_per_frame, int* counts, uint8_t* seg_out, drfe_plane* planes,
                            int plane_cap, int* nr_planes, drfe_cylinder* cylinders, int cyl_cap, int* nr_cylinders) {
  drfe::NvtxRange nvtx_("drfe_pool_extract_batch");
  if (!p || !gray || !depth || !counts || !nr_planes) { drfe::set_error("drfe_pool_extract_batch: null argument"); return DRFE_ERR_ARG; }
  if (nframes < 1 || nframes > p->prm.max_batch) { drfe::set_error("drfe_pool_extract_batch: nframes %d outside [1,%d]", nframes, p->prm.max_batch); return DRFE_ERR_ARG; }
  const int nd = (int)p->workers.size();
  {
    std::lock_guard<std::mutex> lock(p->mu);
    PoolJob& J = p->job;
    J.nframes = nframes;
    J.gray = gray; J.gray_rs = gray_row_stride; J.gray_fs = gray_frame_stride;
    J.depth = depth; J.depth_is_u16 = depth_is_u16; J.depth_factor = depth_factor; J.depth_rs = depth_row_stride; J.depth_fs = depth_frame_stride;
    J.fx = fx; J.fy = fy; J.cx = cx; J.cy = cy;
    J.kps = kps; J.desc = desc; J.cap_per_frame = cap_per_frame; J.counts = counts;
    J.seg = seg_out; J.planes = planes; J.plane_cap = plane_cap; J.nr_planes = nr_planes;
    J.cyls = cylinders; J.cyl_cap = cyl_cap; J.nr_cyls = nr_cylinders;
    J.chunk_frames = p->prm.chunk_frames;
    // contiguous blocks that differ by at most one frame, the first (nframes % nd) devices take the longer ones
    // (the same rule as dr-slam_b200/shard.py frame_block, which the torchrun ranks of bench.py use)
    const int base = nframes / nd, extra = nframes % nd;
    for (int i = 0; i < nd; ++i) {
      p->workers[i].f0 = i * base + std::min(i, extra);
      p->workers[i].f1 = p->workers[i].f0 + base + (i < extra ? 1 : 0);
    }
    p->remaining = nd;
    ++p->generation;
  }
  p->cv_go.notify_all();
  {
    std::unique_lock<std::mutex> lock(p->mu);
    p->cv_done.wait(lock, [&] { return p->remaining == 0; });
  }
  for (const PoolWorker& w : p->workers)
    if (w.rc != DRFE_OK) { drfe::set_error("device %d (frames %d..%d): %s", w.device, w.f0, w.f1 - 1, w.err.c_str()); return w.rc; }
  return DRFE_OK;
}

int drfe_pool_device_times(const drfe_pool* p, float* ms, int cap) {
  if (!p || !ms) return DRFE_ERR_ARG;
  for (int i = 0; i < cap && i < (int)p->workers.size(); ++i) ms[i] = p->workers[i].ms;
  return DRFE_OK;
}

// ---- pinned host memory for the callers of the batch calls (a pageable buffer makes every H2D copy synchronous and
// staged by the driver; DR-SLAM's cv::Mat data can be registered in place)
int drfe_host_alloc(void** ptr, size_t bytes, int write_combined) {
  if (!ptr) return DRFE_ERR_ARG;
  DRFE_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
  return DRFE_OK;
}
int drfe_host_free(void* ptr) {
  if (ptr) DRFE_CUDA(cudaFreeHost(ptr));
  return DRFE_OK;
}
int drfe_host_register(void* ptr, size_t bytes) {
  if (!ptr) return DRFE_ERR_ARG;
  DRFE_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return DRFE_OK;
}
int drfe_host_unregister(void* ptr) {
  if (ptr) DRFE_CUDA(cudaHostUnregister(ptr));
  return DRFE_OK;
}

}  // extern "C"
