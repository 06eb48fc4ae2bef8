/* ORACLE — TEST INFRASTRUCTURE ONLY (see header of orb_oracle.cpp / cape_oracle.cpp).
 * C entry points of the CPU restatement; loaded through ctypes by tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  Never linked into libdrfe.so. */
#ifndef DRFE_ORACLE_H_
#define DRFE_ORACLE_H_
#include "../include/drfe.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- ORB (orb_oracle.cpp) ---- */
void* orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void orc_orb_destroy(void* h);
int orc_orb_run(void* h, const uint8_t* gray, int w, int hgt, int stride);
int orc_orb_features_per_level(void* h, int level);
float orc_orb_scale_factor(void* h, int level);
int orc_orb_umax(void* h, int v);
int orc_orb_level_size(void* h, int level, int* w, int* hgt);
int orc_orb_get_level(void* h, int level, int bordered, uint8_t* dst);
int orc_orb_get_blurred(void* h, int level, uint8_t* dst);
int orc_orb_get_candidates(void* h, int level, float* xyr, int cap);
int orc_orb_get_level_keypoints(void* h, int level, drfe_keypoint* dst, int cap);
int orc_orb_level_tie(void* h, int level);
int orc_orb_get_result(void* h, drfe_keypoint* kps, uint8_t* desc, int cap);
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
void orc_border101_u8(const uint8_t* src, int w, int h, int b, uint8_t* dst);
int orc_fast9_nms(const uint8_t* img, int stride, int w, int h, int threshold, float* xyr, int cap);
void orc_gaussian7_u8(const uint8_t* src, int w, int h, uint8_t* dst);
float orc_fast_atan2(float y, float x);
int orc_distribute_quadtree(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N,
                            float* out_xyr, int cap, int* tie);

/* per-frame steps after extraction: Frame::UndistortKeyPoints / ComputeImageBounds / ComputeStereoFromRGBD /
 * AssignFeaturesToGrid (Frame.cc:835-911, 224-237); returns the number of keypoints placed in the grid */
void orc_undistort_point(const drfe_frame_params* p, float u, float v, float* ou, float* ov);
int orc_frame_image_bounds(drfe_frame_params* p, int width, int height);
int orc_frame_post(const drfe_frame_params* p, const drfe_keypoint* keys, int n, const float* depth, int row_stride,
                   drfe_keypoint* keys_un, float* u_right, float* kp_depth, uint16_t* grid_count, uint16_t* grid_index);

/* ---- matchers (match_oracle.cpp): the C++ restatement beside the Python one of oracle.py ---- */
int orc_search_last_frame(const drfe_frame_params* p, const float* scale_factors, const drfe_keypoint* keys_un, const float* u_right, int n,
                          const uint16_t* grid_count, const uint16_t* grid_index, const uint8_t* desc, const float* Tcw, float th, int mode,
                          int check_orientation, const drfe_last_point* points, const uint8_t* pdesc, int npoints, const uint8_t* occupied,
                          int32_t* match_key, int32_t* match_dist, int32_t* holder);
int orc_search_local_points(const drfe_frame_params* p, const drfe_keypoint* keys_un, const float* u_right, int n, const uint16_t* grid_count,
                            const uint16_t* grid_index, const uint8_t* desc, const drfe_proj_query* queries, const uint8_t* qdesc,
                            const uint8_t* qflags, int nq, float mfNNratio, const uint8_t* occupied, drfe_proj_match* out, int32_t* assigned,
                            int32_t* holder);

/* ---- CAPE (cape_oracle.cpp) ---- */
void* orc_cape_create(int depth_height, int depth_width, int cell_width, int cell_height,
                      int cylinder_detection, float min_cos_angle_4_merge, float max_merge_dist);
void orc_cape_destroy(void* h);
/* PlaneDetection_CAPE::runPlaneDetection cloud + organize (PlaneExtractor.cpp:112-152) */
void orc_cape_depth_to_cloud(void* h, const float* depth, int row_stride, float fx, float fy,
                             float cx, float cy, float* cloud_cellmajor);
/* CAPE::process; seg_out must be zeroed by the caller like the reference */
int orc_cape_process(void* h, const float* cloud_cellmajor, uint8_t* seg_out, drfe_plane* planes,
                     int plane_cap, int* nr_planes, drfe_cylinder* cyls, int cyl_cap,
                     int* nr_cylinders);
int orc_cape_get_cells(void* h, drfe_plane* cells);
int orc_cape_get_grid_maps(void* h, int32_t* plane_map, uint8_t* eroded_map);
/* cylinder_detection = 1: size of cylinder_segments_final (CAPE.cpp:434-445; nr_cylinders of
 * orc_cape_process is nr_cylinders_final, the ones that survive erosion), cylinder cell maps */
int orc_cape_cylinders_found(void* h);
int orc_cape_get_cyl_maps(void* h, int32_t* cyl_map, uint8_t* cyl_eroded_map);
/* the declared rand() stream of CylinderSeg: glibc TYPE_3, srand(seed) */
int orc_glibc_rand(uint32_t seed, int n, int32_t* out);
/* PEAC (ahc::PlaneFitter, include/peac/): PlaneDetection::readDepthImage and PlaneFitter::run with doRefine (peac_oracle.cpp) */
void orc_integral_normals(const float* cloud, int w, int h, float max_depth_change_factor, float smoothing_size, float* normals, float* distance_map);
void orc_peac_fit(const double* s9, int N, double* out8);
void orc_peac_cloud(const uint16_t* depth, int width, int height, int row_stride, float depth_factor, float fx, float fy, float cx, float cy,
                    double* cloud);
int orc_peac_run(const double* cloud, int width, int height, const double* prm11, int minSupport, int windowWidth, int windowHeight, uint8_t* seg_out,
                 double* planes, int plane_cap, int* member_offsets, int* member_idx, int member_cap, int* steps_out);
/* pcl::VoxelGrid (Frame.cc:1121-1125) on one point list; the 1/3-resolution cloud of Frame.cc:1153-1172 */
int orc_voxel_grid(const float* xyz, int n, float leaf, float* out, int* unfiltered);
void orc_third_cloud(const float* depth, int width, int height, int row_stride, float fx, float fy, float cx, float cy, float max_point_dist,
                     float* out);

#ifdef __cplusplus
}
#endif
#endif
