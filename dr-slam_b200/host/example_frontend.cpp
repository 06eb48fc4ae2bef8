// Host-side usage of the adapters, written the way Frame::Frame drives the extractors
// (reference src/Frame.cc:124-134: one std::thread per extractor, then join): ORB and the CAPE
// plane detection run concurrently on one synthetic RGB-D frame; prints a digest that
// tests/test_gpu_adapters.py compares with the CPU oracle.  With a vocabulary file (argv[3], ORBvoc.txt format) it goes on
// like Tracking does: the per-frame steps on the keypoints, Frame::ComputeBoW, and the two whole-function matchers with
// the frame matched against itself (identity pose / itself as the keyframe).
//   build:  g++ -std=c++17 -O2 example_frontend.cpp -o example_frontend -L.. -ldrfe -Wl,-rpath,'$ORIGIN/..' -lpthread
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "CAPE.h"
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "PlaneExtractor.h"
#include "drfe_synth.h"   // tools/synth: the tests' input generator, compiled into this example, not part of libdrfe

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
  if (drfe_synth_frame(W, H, scene, seed, 1.0f, gray.data(), depth.data(), &fx, &fy, &cx, &cy) != 0) return 2;

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
    if (argc > 3) {
      drfe_frame_params fp = {fx, fy, cx, cy, {0.1f, -0.05f, 0.001f, 0.0005f, 0.f}, 40.f, 0, 0, 0, 0};
      drfe_frame_image_bounds(&fp, W, H);                              // Frame::ComputeImageBounds, first frame only
      std::vector<drfe_compat::KeyPoint> keysUn;
      std::vector<float> uRight, kpDepth;
      Planar_SLAM::FramePost(orb, fp, depth.data(), W, H, &keysUn, &uRight, &kpDepth);
      Planar_SLAM::ORBVocabulary voc;
      if (!voc.loadFromTextFile(argv[3])) { fprintf(stderr, "cannot load %s\n", argv[3]); return 3; }
      DBoW2::BowVector bow;
      DBoW2::FeatureVector fv;
      voc.transform(orb, bow, fv, 4);                                  // Frame::ComputeBoW (Frame.cc:828-833)
      uint64_t hb = 1469598103934665603ull, hf = hb;
      for (auto& kv : bow) { hb = fnv1a(&kv.first, 4, hb); hb = fnv1a(&kv.second, 8, hb); }
      for (auto& kv : fv) { hf = fnv1a(&kv.first, 4, hf); hf = fnv1a(kv.second.data(), kv.second.size() * 4, hf); }
      // TrackWithMotionModel against itself: every keypoint with depth is a last-frame point at its own back-projection
      const int n = (int)keys.size();
      std::vector<drfe_last_point> last(n);
      std::vector<uint8_t> lastDesc(desc.data, desc.data + (size_t)n * 32), good(n, 1);
      std::vector<float> ang(n);
      for (int i = 0; i < n; ++i) {
        const float z = kpDepth[i] > 0 ? kpDepth[i] : 2.f;
        last[i] = {(keysUn[i].pt.x - cx) * z / fx, (keysUn[i].pt.y - cy) * z / fy, z, keysUn[i].angle, keysUn[i].octave,
                   DRFE_LP_VALID | DRFE_LP_OBSERVED};
        ang[i] = keysUn[i].angle;
      }
      const float Tcw[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
      Planar_SLAM::ORBmatcher matcher(0.9f, true);                     // Tracking.cc: ORBmatcher matcher(0.9,true)
      std::vector<int32_t> mp, mb;
      const int nproj = matcher.SearchByProjection(orb, Tcw, last, lastDesc, 15.f, 0, mp);
      Planar_SLAM::ORBmatcher matcher2(0.7f, true);                    // TrackReferenceKeyFrame: ORBmatcher matcher(0.7,true)
      const int nbow = matcher2.SearchByBoW(lastDesc, ang, good, fv, orb, fv, mb);
      printf("words %u bow %zu bow_hash %016llx fv %zu fv_hash %016llx proj %d proj_hash %016llx bowmatch %d bowmatch_hash %016llx\n", voc.size(),
             bow.size(), (unsigned long long)hb, fv.size(), (unsigned long long)hf, nproj, (unsigned long long)fnv1a(mp.data(), (size_t)n * 4), nbow,
             (unsigned long long)fnv1a(mb.data(), (size_t)n * 4));
    }
    // the plane extractor Frame::Frame runs (Frame.cc:126 -> ComputePlanes :937-949) on the 16-bit depth map; the synthetic
    // sensor's isolated dropouts are filled from the left first (INIT_STRICT drops every window with a missing pixel)
    {
      std::vector<uint16_t> d16((size_t)W * H);
      for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
          uint16_t v = (uint16_t)lrintf(depth[(size_t)r * W + c] * 5000.f);
          if (v == 0 && c > 0) v = d16[(size_t)r * W + c - 1];
          d16[(size_t)r * W + c] = v;
        }
      Planar_SLAM::PlaneDetection planeDetector;
      planeDetector.readDepthImage(drfe_compat::Mat16u(H, W, d16.data(), (size_t)W * sizeof(uint16_t)), K, 1.0f / 5000.0f);
      planeDetector.runPlaneDetection();
      uint64_t hv = 1469598103934665603ull, hp = hv;
      size_t nv = 0;
      for (int i = 0; i < planeDetector.plane_num_; ++i) {
        auto& indices = planeDetector.plane_vertices_[i];
        nv += indices.size();
        hv = fnv1a(indices.data(), indices.size() * sizeof(int), hv);
        for (int j : indices) {                                          // Frame.cc:958-963
          const float p[3] = {(float)planeDetector.cloud.vertices[j][0], (float)planeDetector.cloud.vertices[j][1], (float)planeDetector.cloud.vertices[j][2]};
          hp = fnv1a(p, sizeof(p), hp);
        }
      }
      printf("peac planes %d seg_hash %016llx vertices %zu index_hash %016llx point_hash %016llx", planeDetector.plane_num_,
             (unsigned long long)fnv1a(planeDetector.seg_output.data, (size_t)W * H), nv, (unsigned long long)hv, (unsigned long long)hp);
      for (int i = 0; i < planeDetector.plane_num_; ++i) {
        auto pl = planeDetector.plane_filter.extractedPlanes[i];
        printf(" | %.17g %.17g %.17g %.17g", pl->normal[0], pl->normal[1], pl->normal[2],
               -(pl->normal[0] * pl->center[0] + pl->normal[1] * pl->center[1] + pl->normal[2] * pl->center[2]));
      }
      std::vector<std::vector<std::array<float, 3>>> coarse;          // voxel.filter(*coarseCloud) of every plane, Frame.cc:981-985
      planeDetector.planeCloudsVoxel(3.0f, 0.05f, coarse);
      size_t nc = 0;
      uint64_t hc = 1469598103934665603ull;
      for (auto& c : coarse) { nc += c.size(); hc = fnv1a(c.data(), c.size() * sizeof(c[0]), hc); }
      printf(" | coarse %zu %016llx", nc, (unsigned long long)hc);
      std::vector<float> cloud3, normals3;                            // Frame.cc:1044-1100: the 1/3 cloud and PCL's normals on it
      int w3 = 0, h3 = 0;
      planeDetector.thirdCloudNormals(10.0f, cloud3, normals3, w3, h3);
      size_t nn = 0;
      uint64_t hn = 1469598103934665603ull;
      for (size_t i = 0; i < normals3.size(); i += 3)
        if (normals3[i] == normals3[i]) { ++nn; hn = fnv1a(&normals3[i], 3 * sizeof(float), hn); }
      printf(" | normals %d %d %zu %016llx", w3, h3, nn, (unsigned long long)hn);
      printf("\n");
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
