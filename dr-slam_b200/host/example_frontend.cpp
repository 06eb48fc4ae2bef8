// Host-side usage of the adapters, written the way Frame::Frame drives the extractors
// (reference src/Frame.cc:124-134: one std::thread per extractor, then join): ORB and the CAPE
// plane detection run concurrently on one synthetic RGB-D frame; prints a digest that
// tests/test_gpu_adapters.py compares with the CPU oracle.
//   build:  g++ -std=c++17 -O2 example_frontend.cpp -o example_frontend -L.. -ldrfe -Wl,-rpath,'$ORIGIN/..' -lpthread
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "CAPE.h"
#include "ORBextractor.h"

static uint64_t fnv1a(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
  const uint8_t* b = (const uint8_t*)p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  const int W = 640, H = 480;
  const int scene = argc > 1 ? atoi(argv[1]) : 1;
  const uint32_t seed = argc > 2 ? (uint32_t)atoll(argv[2]) : 20260042u;
  std::vector<uint8_t> gray((size_t)W * H);
  std::vector<float> depth((size_t)W * H);
  float fx, fy, cx, cy;
  if (drfe_synth_frame(W, H, scene, seed, 1.0f, gray.data(), depth.data(), &fx, &fy, &cx, &cy) != DRFE_OK) return 2;

  try {
    Planar_SLAM::ORBextractor orb(1000, 1.2f, 8, 20, 7);          // Tracking.cc:120-126
    Planar_SLAM::PlaneDetection_CAPE planes;                        // Frame.cc:1096-1104
    planes.PATCH_SIZE = 20; planes.MAX_MERGE_DIST = 50.f;

    std::vector<drfe_compat::KeyPoint> keys;
    drfe_compat::Mat8u desc, mask;
    drfe_compat::Mat8u im(H, W, gray.data(), (size_t)W);
    const float K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};

    std::thread threadORB([&] { orb(im, mask, keys, desc); });      // Frame::ExtractORB (Frame.cc:473-478)
    std::thread threadPlanes([&] {                                  // Frame::ComputePlanes_CAPE
      planes.readDepthImage(drfe_compat::Mat32f(H, W, depth.data(), (size_t)W * sizeof(float)), K);
      planes.runPlaneDetection();
    });
    threadORB.join();
    threadPlanes.join();

    size_t npts = 0;
    for (auto& pc : planes.plane_cloud) npts += pc.size();
    printf("keypoints %zu kp_hash %016llx desc_hash %016llx planes %d seg_hash %016llx plane_points %zu levels %d\n", keys.size(),
           (unsigned long long)fnv1a(keys.data(), keys.size() * sizeof(keys[0])),
           (unsigned long long)fnv1a(desc.data, (size_t)desc.rows * 32), planes.nr_planes,
           (unsigned long long)fnv1a(planes.seg_output.data, (size_t)W * H), npts, orb.GetLevels());
    for (int i = 0; i < planes.nr_planes; ++i)
      printf("plane %d n %.9f %.9f %.9f d %.9f\n", i, planes.plane_params[i].normal[0], planes.plane_params[i].normal[1],
             planes.plane_params[i].normal[2], planes.plane_params[i].d);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
