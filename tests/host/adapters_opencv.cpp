// The adapters' DRFE_WITH_OPENCV / DRFE_WITH_EIGEN branches — the code a DR-SLAM maintainer actually compiles — built
// against the mock headers of tests/host/mock and driven with the REFERENCE's call shapes:
//   (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors)                          Frame.cc:473-478
//   planeDetector.readDepthImage(imDepth, K); planeDetector.runPlaneDetection()         Frame.cc:1096-1104
//   plane_detector->process(cloud_array_organized, nr_planes, nr_cylinders, seg_output, plane_params, cylinder_params)
//                                                                                        PlaneExtractor.cpp:149-154
// Prints the same digest line as dr-slam_b200/host/example_frontend (tests/test_gpu_adapters.py compares the two and
// the oracle).  Test infrastructure; build: see tests/test_host_adapters_opencv.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "CAPE.h"
#include "ORBextractor.h"
#include "PlaneExtractor.h"
#include "drfe_synth.h"

static uint64_t fnv1a(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
  const uint8_t* b = (const uint8_t*)p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  const int W = 640, H = 480;
  const int scene = argc > 1 ? atoi(argv[1]) : 1;
  const uint32_t seed = argc > 2 ? (uint32_t)atoll(argv[2]) : 20260042u;
  cv::Mat im(H, W, CV_8UC1), imDepth(H, W, CV_32FC1), K(3, 3, CV_32FC1);
  float fx, fy, cx, cy;
  if (drfe_synth_frame(W, H, scene, seed, 1.0f, im.data, imDepth.ptr<float>(0), &fx, &fy, &cx, &cy) != 0) return 2;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) K.at<float>(r, c) = 0.f;
  K.at<float>(0, 0) = fx; K.at<float>(1, 1) = fy; K.at<float>(0, 2) = cx; K.at<float>(1, 2) = cy; K.at<float>(2, 2) = 1.f;
  try {
    Planar_SLAM::ORBextractor orb(1000, 1.2f, 8, 20, 7);
    Planar_SLAM::PlaneDetection_CAPE planes;
    planes.PATCH_SIZE = 20; planes.MAX_MERGE_DIST = 50.f;
    std::vector<cv::KeyPoint> mvKeys;
    cv::Mat mDescriptors;
    std::thread threadORB([&] { orb(im, cv::Mat(), mvKeys, mDescriptors); });
    std::thread threadPlanes([&] {
      planes.readDepthImage(imDepth, K);
      planes.runPlaneDetection();
    });
    threadORB.join();
    threadPlanes.join();
    size_t npts = 0;
    for (auto& pc : planes.plane_cloud) npts += pc.size();
    // descriptors row by row: a cv::Mat need not be continuous
    uint64_t dh = 1469598103934665603ull;
    for (int i = 0; i < mDescriptors.rows; ++i) dh = fnv1a(mDescriptors.ptr(i), 32, dh);
    uint64_t sh = 1469598103934665603ull;
    for (int r = 0; r < H; ++r) sh = fnv1a(planes.seg_output.ptr(r), (size_t)W, sh);
    printf("keypoints %zu kp_hash %016llx desc_hash %016llx planes %d seg_hash %016llx plane_points %zu levels %d\n", mvKeys.size(),
           (unsigned long long)fnv1a(mvKeys.data(), mvKeys.size() * sizeof(mvKeys[0])), (unsigned long long)dh, planes.nr_planes,
           (unsigned long long)sh, npts, orb.GetLevels());
    for (int i = 0; i < planes.nr_planes; ++i)
      printf("plane %d n %.9f %.9f %.9f d %.9f\n", i, planes.plane_params[i].normal[0], planes.plane_params[i].normal[1],
             planes.plane_params[i].normal[2], planes.plane_params[i].d);
    // CAPE::process with the reference's argument list on an Eigen::MatrixXf cloud (cell-major, N x 3 column-major) built
    // the way PlaneExtractor.cpp:112-148 builds it
    const int cw = 20, ch = 20, ncx = W / cw;
    Eigen::MatrixXf cloud(W * H, 3);
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) {
        const double z = (double)imDepth.at<float>(r, c);
        const int id = ((r / ch) * ncx + c / cw) * cw * ch + (r % ch) * cw + (c % cw);
        cloud(id, 0) = (float)(((double)c - (double)cx) * z / (double)fx);
        cloud(id, 1) = (float)(((double)r - (double)cy) * z / (double)fy);
        cloud(id, 2) = (float)z;
      }
    CAPE detector(H, W, cw, ch, false, planes.COS_ANGLE_MAX, 50.f);
    int nr_planes = 0, nr_cylinders = 0;
    cv::Mat seg(H, W, CV_8U);
    seg.setTo(cv::Scalar(0));
    std::vector<PlaneSeg> plane_params;
    std::vector<CylinderSeg> cylinder_params;
    detector.process(cloud, nr_planes, nr_cylinders, seg, plane_params, cylinder_params);
    uint64_t sh2 = 1469598103934665603ull;
    for (int r = 0; r < H; ++r) sh2 = fnv1a(seg.ptr(r), (size_t)W, sh2);
    printf("process planes %d cylinders %d seg_hash %016llx appended %zu\n", nr_planes, nr_cylinders, (unsigned long long)sh2, plane_params.size());
    // an empty image leaves the caller's containers untouched (ORBextractor.cc:1046-1047)
    std::vector<cv::KeyPoint> untouched(3);
    cv::Mat d2;
    orb(cv::Mat(), cv::Mat(), untouched, d2);
    printf("empty_image keeps %zu\n", untouched.size());
    // planeDetector.readDepthImage(Depth, K, depthFactor); planeDetector.runPlaneDetection()   Frame.cc:940-942
    {
      cv::Mat Depth(H, W, CV_16UC1);
      for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
          unsigned short v = (unsigned short)lrintf(imDepth.at<float>(r, c) * 5000.f);
          if (v == 0 && c > 0) v = Depth.at<unsigned short>(r, c - 1);
          Depth.at<unsigned short>(r, c) = v;
        }
      Planar_SLAM::PlaneDetection planeDetector;
      planeDetector.readColorImage(im);
      planeDetector.readDepthImage(Depth, K, 1.0f / 5000.0f);
      planeDetector.runPlaneDetection();
      uint64_t hv = 1469598103934665603ull, hp = hv, hs = hv;
      size_t nv = 0;
      for (int i = 0; i < planeDetector.plane_num_; ++i) {
        auto& indices = planeDetector.plane_vertices_[i];
        nv += indices.size();
        hv = fnv1a(indices.data(), indices.size() * sizeof(int), hv);
        for (int j : indices) {
          const float p[3] = {(float)planeDetector.cloud.vertices[j][0], (float)planeDetector.cloud.vertices[j][1], (float)planeDetector.cloud.vertices[j][2]};
          hp = fnv1a(p, sizeof(p), hp);
        }
      }
      for (int r = 0; r < H; ++r) hs = fnv1a(planeDetector.seg_output.ptr(r), (size_t)W, hs);
      printf("peac planes %d seg_hash %016llx vertices %zu index_hash %016llx point_hash %016llx", planeDetector.plane_num_, (unsigned long long)hs, nv,
             (unsigned long long)hv, (unsigned long long)hp);
      for (int i = 0; i < planeDetector.plane_num_; ++i) {
        auto extractedPlane = planeDetector.plane_filter.extractedPlanes[i];
        printf(" | %.17g %.17g %.17g %.17g", extractedPlane->normal[0], extractedPlane->normal[1], extractedPlane->normal[2],
               -(extractedPlane->normal[0] * extractedPlane->center[0] + extractedPlane->normal[1] * extractedPlane->center[1] +
                 extractedPlane->normal[2] * extractedPlane->center[2]));
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
