/* drfe — B200-native RGB-D feature front end: the C ABI.
 *
 * This header is the drop-in boundary for DR-SLAM's per-frame feature front end
 * (the work Frame::Frame spawns before Tracking::Track, reference src/Frame.cc:124-134):
 *
 *   Planar_SLAM::ORBextractor::operator()(image, mask, keypoints, descriptors)
 *        reference include/ORBextractor.h:51-61, src/ORBextractor.cc:1043-1105
 *   CAPE::process(cloud, nr_planes, nr_cylinders, seg_output, planes, cylinders)
 *        reference src/CAPE/CAPE.h:47-48, src/CAPE/CAPE.cpp:47-457
 *   PlaneDetection_CAPE::runPlaneDetection()  (depth -> organized cloud -> CAPE)
 *        reference src/PlaneExtractor.cpp:111-191
 *
 * and, on the results those leave on the device, the data-parallel per-frame steps either side of them
 * ("next" rows of SURVEY.md 8f), each entry point citing what it replaces:
 *   cvtColor to gray (Tracking.cc:194-207), depth scaling (Frame.cc:113-115)
 *   UndistortKeyPoints / ComputeStereoFromRGBD / AssignFeaturesToGrid (Frame.cc:835-911, 224-237)
 *   per-plane point lists (PlaneExtractor.cpp:165-190)
 *   Frame::ComputeBoW = DBoW2 transform (Frame.cc:828-833, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1258)
 *   ORBmatcher::SearchByProjection (map points :46-130, last frame :1396-1535), SearchByBoW (:160-292)
 *
 * Plain C, POD only (no OpenCV / Eigen / torch types).  Every function returns an
 * int status: 0 = DRFE_OK, negative = error (drfe_last_error() has the text).  There
 * is NO CPU fallback: if no CUDA device is usable every create call fails loudly.
 *
 * Threading: a handle may be driven from any host thread, one call at a time per
 * handle (the reference calls its extractors from a fresh std::thread each frame,
 * Frame.cc:124-127).  Each handle owns a non-blocking CUDA stream; nothing runs on the
 * legacy default stream.
 *
 * The header-only C++ adapters in dr-slam_b200/host/ re-create the reference's class
 * signatures on top of these entry points; INTEGRATION.md shows the binding a DR-SLAM
 * maintainer would add.
 */
#ifndef DRFE_H_
#define DRFE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRFE_OK 0
#define DRFE_ERR_ARG (-1)      /* bad argument (null, size mismatch, unsupported config) */
#define DRFE_ERR_CUDA (-2)     /* CUDA runtime error or no usable device                 */
#define DRFE_ERR_CAPACITY (-3) /* a device-side buffer overflowed (result incomplete)     */
#define DRFE_ERR_STATE (-4)    /* call order violated (e.g. download before enqueue)      */

#define DRFE_MEM_HOST 0   /* pointer is host memory (pinned => the copy is asynchronous) */
#define DRFE_MEM_DEVICE 1 /* pointer is device memory on the handle's device             */

#define DRFE_MAX_LEVELS 16

/* Same layout as cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave,
 * class_id.  ORBextractor fills class_id = -1, size = int(31*scale[octave]), angle in
 * degrees [0,360) (ORBextractor.cc:837-847, :77-104). */
typedef struct drfe_keypoint {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} drfe_keypoint;

/* Constructor arguments of ORBextractor (ORBextractor.h:51-52). */
typedef struct drfe_orb_params {
  int32_t nfeatures;
  float scale_factor;
  int32_t nlevels;
  int32_t ini_th_fast;
  int32_t min_th_fast;
} drfe_orb_params;

/* Mirror of the public data members of PlaneSeg (src/CAPE/PlaneSeg.h:15-28).  (The 4 padding bytes after `planar` are not written:
 * compare records field by field, not with memcmp.) */
typedef struct drfe_plane {
  int32_t nr_pts, min_nr_pts;
  double x_acc, y_acc, z_acc, xx_acc, yy_acc, zz_acc, xy_acc, xz_acc, yz_acc;
  float score, MSE;
  int32_t planar;
  double mean[3], normal[3], d;
} drfe_plane;

/* Mirror of what CAPE::process copies out per cylinder (CAPE.cpp:434-445). */
typedef struct drfe_cylinder {
  float radius;
  double center[3];
  double axis[3];
} drfe_cylinder;

/* Constructor arguments of CAPE (CAPE.h:47) + camera intrinsics used by the wrapper
 * (PlaneExtractor.cpp:117-127). */
typedef struct drfe_cape_params {
  int32_t depth_height, depth_width;
  int32_t cell_width, cell_height;
  int32_t cylinder_detection;
  float min_cos_angle_4_merge; /* reference default 0.97814; DR-SLAM passes cos(pi/12) */
  float max_merge_dist;        /* reference default 900;     DR-SLAM passes Plane.MAX_MERGE_DIST */
} drfe_cape_params;

typedef struct drfe_orb drfe_orb;
typedef struct drfe_cape drfe_cape;

/* ------------------------------------------------------------------ general */
const char* drfe_last_error(void);   /* thread-local text of the last failure            */
const char* drfe_version(void);
int drfe_device_count(int* count);   /* number of CUDA devices visible                    */
/* total number of drfe kernel launches issued by this process (bench "gpu_launches") */
int64_t drfe_kernel_launch_count(void);

/* CUDA-event timers on handle streams (for benches driving several handles): events are
 * opaque; `stream` is what drfe_orb_stream / drfe_cape_stream return. */
int drfe_event_create(void** ev);
int drfe_event_destroy(void* ev);
int drfe_event_record(void* ev, void* stream);
int drfe_stream_wait_event(void* stream, void* ev);
int drfe_event_elapsed_ms(void* start, void* stop, float* ms); /* waits for `stop` */

/* ------------------------------------------------------------------ ORB
 * One handle = one ORBextractor instance bound to a device, an image size and a maximum
 * frame batch (frames of a batch are independent; batch=1 reproduces operator()). */
int drfe_orb_create(const drfe_orb_params* params, int width, int height, int max_batch,
                    int device, drfe_orb** out);
int drfe_orb_destroy(drfe_orb* h);

/* Getters mirroring ORBextractor::GetLevels / GetScaleFactor(s) / GetInverseScaleFactors /
 * GetScaleSigmaSquares / GetInverseScaleSigmaSquares (ORBextractor.h:63-83); arrays hold
 * nlevels floats. */
int drfe_orb_get_levels(const drfe_orb* h);
float drfe_orb_get_scale_factor(const drfe_orb* h);
int drfe_orb_get_scale_factors(const drfe_orb* h, float* scale, float* inv_scale,
                               float* sigma2, float* inv_sigma2);
int drfe_orb_features_per_level(const drfe_orb* h, int level); /* mnFeaturesPerLevel */
/* upper bound on keypoints one frame can return (>= nfeatures; the quadtree may overshoot
 * each level's quota by up to 3, ORBextractor.cc:730) */
int drfe_orb_max_keypoints(const drfe_orb* h);

/* operator()(image, mask, keypoints, descriptors) for ONE 8-bit gray frame in host memory.
 * `mask` is ignored by the reference (ORBextractor.h:58) and has no parameter here.
 * gray == NULL or w/h == 0 mirrors the reference's silent return on an empty image
 * (ORBextractor.cc:1046): *n is left untouched and DRFE_OK is returned.
 * desc receives 32 bytes per keypoint, row-aligned with kps. */
int drfe_orb_extract(drfe_orb* h, const uint8_t* gray, int width, int height, size_t row_stride,
                     drfe_keypoint* kps, uint8_t* desc, int cap, int* n);

/* Batched, asynchronous form.  enqueue: (copy nframes images, H2D if mem_kind is host,) and
 * launch the whole pipeline on the handle's stream; returns without waiting.
 * download: wait, then copy results to host: kps[f*cap_per_frame + i], desc[(f*cap+i)*32],
 * counts[f].  Any of kps/desc may be NULL to skip that copy. */
int drfe_orb_enqueue(drfe_orb* h, int nframes, const uint8_t* gray, size_t row_stride,
                     size_t frame_stride, int mem_kind);
int drfe_orb_download(drfe_orb* h, drfe_keypoint* kps, uint8_t* desc, int cap_per_frame,
                      int* counts);
int drfe_orb_sync(drfe_orb* h);
void* drfe_orb_stream(drfe_orb* h); /* the handle's cudaStream_t */

/* The colour conversion Tracking::GrabImageRGBD does before the frame is built (Tracking.cc:194-207:
 * cvtColor(mImGray, mImGray, CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY)) fused in
 * front of drfe_orb_enqueue: pixels = interleaved 8-bit colour frames (channels 3 or 4, rgb_order != 0: R first as in
 * CV_RGB2GRAY, 0: B first), row_stride / frame_stride in BYTES.  coeffs selects cv::cvtColor's fixed-point arithmetic:
 * DRFE_GRAY_Q15 = (R*9798 + G*19235 + B*3735 + 2^14) >> 15 — OpenCV 4.x and late 3.4.x; bit-identical to cv2 4.13 on
 * all 2^24 colours (tests/test_gpu_color.py); DRFE_GRAY_Q14 = (R*4899 + G*9617 + B*1868 + 2^13) >> 14 — OpenCV 2.4 ..
 * 3.4.x before the bit-exact rewrite (published constants R2Y / G2Y / B2Y, yuv_shift 14; no library here to pin it
 * against).  drfe_orb_get_gray copies the converted frame (mImGray) back. */
#define DRFE_GRAY_Q15 0
#define DRFE_GRAY_Q14 1
int drfe_orb_enqueue_color(drfe_orb* h, int nframes, const uint8_t* pixels, int channels, int rgb_order,
                           int coeffs, size_t row_stride, size_t frame_stride, int mem_kind);
int drfe_orb_get_gray(drfe_orb* h, int frame, uint8_t* dst);

/* A whole batch of HOST images in, HOST results out, in one asynchronous call: the batch is cut
 * into chunks of chunk_frames frames (<= 0: 32-frame chunks with 8- and 16-frame chunks at both
 * ends, so that the first kernels start early and the last copy out is short) and chunk k's host->device copy, kernels and
 * device->host copies run on three streams, so PCIe in both directions overlaps the kernels
 * (pinned host memory is needed for the overlap; pageable memory works, serialised).  The call
 * returns after queueing; the buffers must stay valid until drfe_orb_finish_batch(), which waits
 * and reports capacity errors.  Results are those of drfe_orb_enqueue + drfe_orb_download. */
int drfe_orb_extract_batch(drfe_orb* h, int nframes, const uint8_t* gray, size_t row_stride,
                           size_t frame_stride, drfe_keypoint* kps, uint8_t* desc,
                           int cap_per_frame, int* counts, int chunk_frames);
int drfe_orb_finish_batch(drfe_orb* h);

/* ------------------------------------------------------------------ per-frame steps after extraction
 * ("next" row of SURVEY.md 8f): what Frame::Frame does with the keypoints right after ExtractORB —
 * UndistortKeyPoints (Frame.cc:835-861), ComputeStereoFromRGBD (:893-911) and AssignFeaturesToGrid
 * (:224-237, PosInGrid :816-825) — on the keypoints of the handle's last batch, which are already
 * on the device. */
#define DRFE_FRAME_GRID_COLS 64 /* FRAME_GRID_COLS / FRAME_GRID_ROWS of include/Frame.h */
#define DRFE_FRAME_GRID_ROWS 48
typedef struct drfe_frame_params {
  float fx, fy, cx, cy;               /* mK */
  float dist[5];                      /* mDistCoef: k1 k2 p1 p2 k3 (dist[0] == 0: keypoints are copied) */
  float bf;                           /* mbf = baseline * fx */
  float min_x, max_x, min_y, max_y;   /* mnMinX .. mnMaxY (drfe_frame_image_bounds) */
} drfe_frame_params;
/* Frame::ComputeImageBounds (Frame.cc:863-891): fills min_x .. max_y of *p from the other fields. */
int drfe_frame_image_bounds(drfe_frame_params* p, int width, int height);
/* depth: the float depth image(s) of the same frames (host or device).  Outputs on the host, any may
 * be NULL: keys_un[f*cap + i] (mvKeysUn), u_right / kp_depth[f*cap + i] (mvuRight / mvDepth, -1 where
 * the depth is not positive), grid_count[f*3072 + x*48 + y] = mGrid[x][y].size(), grid_index[f*cap ..]
 * = the keypoint indices of all cells concatenated in x-major cell order, ascending inside a cell
 * (what the push_back loop produces). */
int drfe_orb_frame_post(drfe_orb* h, const drfe_frame_params* p, const float* depth, size_t row_stride,
                        size_t frame_stride, int mem_kind, drfe_keypoint* keys_un, float* u_right,
                        float* kp_depth, uint16_t* grid_count, uint16_t* grid_index, int cap_per_frame);

/* The same with the depth images a CAPE handle already holds on the device (its last enqueue_depth / enqueue_depth_u16 /
 * process_depth_batch of the same frames: float metres, or the sensor's raw 16-bit depth, converted per gathered pixel
 * as (float)u16 * factor like imDepth.convertTo(CV_32F, mDepthMapFactor), Frame.cc:113-115): the depth crosses PCIe once
 * for both extractors, as one imDepth serves both in Frame::Frame.  The handle's stream waits for the CAPE stream on the
 * device.  In a batch call the CAPE handle's drfe_cape_finish_batch must have returned (its copies run on their own stream). */
int drfe_orb_frame_post_shared_depth(drfe_orb* h, drfe_cape* cape, const drfe_frame_params* p,
                                     drfe_keypoint* keys_un, float* u_right, float* kp_depth,
                                     uint16_t* grid_count, uint16_t* grid_index, int cap_per_frame);

/* The data-parallel core of ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)
 * (ORBmatcher.cc:46-130; first step of SURVEY.md 8f next-3) on the device-resident results of the last
 * drfe_orb_frame_post: per query, the keypoints Frame::GetFeaturesInArea(x, y, r, min_level, max_level)
 * returns (Frame.cc:730-779) — minus those flagged in occupied[] (F.mvpMapPoints[idx] with observations,
 * :88-90) and those whose mvuRight is > 0 and further than r from xr (:92-97) — ranked by
 * ORBmatcher::DescriptorDistance (:1712-1728) into best / second best exactly as lines 103-115 do.
 * The caller passes r = RadiusByViewingCos(...) * th * mvScaleFactors[level], applies the ratio test
 * (:119-122) and does the in-order assignment (:124), which is sequential Tracking logic. */
typedef struct drfe_proj_query {
  float x, y, r, xr;                 /* mTrackProjX, mTrackProjY, window radius, mTrackProjXR      */
  int32_t min_level, max_level;      /* nPredictedLevel - 1, nPredictedLevel                       */
} drfe_proj_query;
typedef struct drfe_proj_match {
  int32_t best_dist, best_idx, best_level, best_dist2, best_level2; /* 256 / -1 when there is none */
} drfe_proj_match;
/* queries[f*qcap + i], qdesc[(f*qcap + i)*32 ..] for i < nqueries[f]; occupied[f*max_keypoints + idx]
 * (may be NULL); out[f*qcap + i].  All pointers are host memory; frames = the handle's last batch. */
int drfe_orb_search_by_projection(drfe_orb* h, const int* nqueries, const drfe_proj_query* queries,
                                  const uint8_t* qdesc, const uint8_t* occupied, int qcap,
                                  drfe_proj_match* out);

/* ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th) (ORBmatcher.cc:46-130) WHOLE, the matcher
 * of Tracking::SearchLocalPoints / TrackLocalMap: the loops of drfe_orb_search_by_projection plus what that call leaves to
 * the caller — the ratio test (:119-122) and the in-order assignment F.mvpMapPoints[bestIdx] = pMP (:124), whose effect on
 * later map points (:88-90) the device reproduces with the parallel sweeps of drfe_orb_search_last_frame.
 * qflags[f*qcap + i]: DRFE_LP_VALID = pMP->mbTrackInView && !pMP->isBad() (:55-59), DRFE_LP_OBSERVED = pMP->Observations() > 0;
 * occupied as in drfe_orb_search_by_projection (F.mvpMapPoints[idx] holds an observed point on entry: the matches of the
 * previous tracking stage).  Outputs (host, any may be NULL): out[f*qcap + i] = best / second best as the reference's loop
 * leaves them; assigned[f*qcap + i] = bestIdx if map point i was assigned, else -1; key_point[f*max_keypoints + idx] = the
 * map point index F.mvpMapPoints[idx] holds on return or -1 (untouched); nmatches[f] = the return value. */
int drfe_orb_search_local_points(drfe_orb* h, const int* nqueries, const drfe_proj_query* queries,
                                 const uint8_t* qdesc, const uint8_t* qflags, const uint8_t* occupied, int qcap,
                                 float nnratio, drfe_proj_match* out, int32_t* assigned, int32_t* key_point,
                                 int* nmatches);

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (ORBmatcher.cc:1396-1535),
 * the matcher TrackWithMotionModel runs on every frame — whole function, on the device-resident results of the last
 * drfe_orb_frame_post (the current frames) — SURVEY.md 8f next-3.  Per last-frame point i with a map point that is not
 * an outlier (:1417-1423): x3Dc = Rcw*x3Dw + tcw (cv::Mat float product, :1426-1427), the projection (:1429-1442),
 * radius = th * mvScaleFactors[octave] (:1447), Frame::GetFeaturesInArea with the level range bForward / bBackward
 * select (:1451-1456), the skips of keypoints that already hold a map point with observations (:1471-1473) and of
 * keypoints failing the right-coordinate test (:1475-1481), DescriptorDistance, first minimum (:1487-1491), match if
 * bestDist <= TH_HIGH (:1494).  The reference's loop is sequential because a match made by point j makes later
 * points skip that keypoint; the device reaches the same result as the fixed point of "every point searches in
 * parallel, skipping the keypoints that lower-numbered points with observations chose in the previous sweep"
 * (point 0 is final after one sweep, point i after at most i + 1; two or three sweeps in practice).  Then the rotation
 * histogram (:1499-1509), ComputeThreeMaxima (:1666-1707) and the removal of the other bins (:1512-1532). */
#define DRFE_TH_HIGH 100     /* ORBmatcher::TH_HIGH (ORBmatcher.cc:38) */
#define DRFE_HISTO_LENGTH 30 /* ORBmatcher::HISTO_LENGTH (:40) */
typedef struct drfe_last_point {
  float X, Y, Z;   /* LastFrame.mvpMapPoints[i]->GetWorldPos()                                              */
  float angle;     /* LastFrame.mvKeysUn[i].angle                                                           */
  int32_t octave;  /* LastFrame.mvKeys[i].octave                                                            */
  int32_t flags;   /* DRFE_LP_VALID: map point present and !mvbOutlier[i]; DRFE_LP_OBSERVED: Observations() > 0 */
} drfe_last_point;
#define DRFE_LP_VALID 1
#define DRFE_LP_OBSERVED 2
typedef struct drfe_track_params {
  float Tcw[12];             /* rows 0..2 of CurrentFrame.mTcw, row-major 3x4: [Rcw | tcw]                  */
  float th;                  /* window factor (15 / 7 in TrackWithMotionModel, doubled on the retry)        */
  int32_t mode;              /* 0: levels octave-1 .. octave+1; 1: bForward (>= octave); 2: bBackward (0 .. octave) (:1413-1414, :1451-1456) */
  int32_t check_orientation; /* mbCheckOrientation                                                          */
} drfe_track_params;
/* tp[f], npoints[f], points[f*pcap + i], pdesc[(f*pcap + i)*32 ..] (pMP->GetDescriptor()); occupied[f*max_keypoints + idx]
 * != 0: CurrentFrame.mvpMapPoints[idx] holds an observed map point on entry (NULL: none, what TrackWithMotionModel's
 * fill(NULL) gives).  Outputs (host, any may be NULL): match_key[f*pcap + i] = bestIdx2 of point i or -1 and
 * match_dist[f*pcap + i] = its bestDist (256 if no candidate), both BEFORE the rotation check; key_point[f*max_keypoints
 * + idx] = the point i that CurrentFrame.mvpMapPoints[idx] holds on return, or -1; nmatches[f] = the return value;
 * sweeps[f] = parallel sweeps used. */
int drfe_orb_search_last_frame(drfe_orb* h, const drfe_track_params* tp, const int* npoints,
                               const drfe_last_point* points, const uint8_t* pdesc, const uint8_t* occupied,
                               int pcap, int32_t* match_key, int32_t* match_dist, int32_t* key_point,
                               int* nmatches, int* sweeps);

/* ------------------------------------------------------------------ Frame::ComputeBoW
 * Frame::ComputeBoW (Frame.cc:828-833) = ORBVocabulary::transform(descriptors, mBowVec, mFeatVec, 4)
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1194): every descriptor descends the vocabulary tree
 * (:1217-1258: at each node the child at the smallest FORB::distance, first minimum wins, FORB.cpp:81-101) to a word;
 * BowVector::addWeight sums the word's weight once per descriptor (BowVector.cpp:34-46), BowVector::normalize divides by
 * the L1 / L2 norm accumulated in word order (:62-84); FeatureVector::addFeature lists the descriptor indices per node at
 * level L - levelsup (FeatureVector.cpp:31-45).  The vocabulary is given as the arrays
 * TemplatedVocabulary::loadFromTextFile parses (TemplatedVocabulary.h:1338-1422): node i + 1 of the file is row i. */
typedef struct drfe_vocab drfe_vocab;
/* k, L, scoring, weighting: the header line of the vocabulary file (m_k, m_L, ScoringType, WeightingType; the ORB
 * vocabulary of ORB-SLAM2 is 10 6 0 0 = L1_NORM, TF_IDF).  Rows i < nnodes: parent[i] (node id of the parent; 0 = root),
 * is_leaf[i], descriptors[32*i ..], weights[i].  Word ids are assigned to the leaves in file order (:1408-1415).
 * All six scorings are handled as ScoringObject.h:73-89 defines them (L1 / L2 normalisation, none for DOT_PRODUCT) and the
 * four weightings as transform does (TF_IDF / TF: addWeight, IDF / BINARY: addIfNotExist). */
int drfe_vocab_create(int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent,
                      const uint8_t* is_leaf, const uint8_t* descriptors, const double* weights, int device,
                      drfe_vocab** out);
void drfe_vocab_destroy(drfe_vocab* v);
int drfe_vocab_words(const drfe_vocab* v); /* m_words.size() */
/* transform() of the device-resident descriptors of the handle's last batch (same device as the vocabulary).
 * Outputs on the host, any may be NULL, cap = the handle's max_keypoints:
 *   word_id / node_id [f*cap + i]  the word and the level-(L - levelsup) node of descriptor i (-1 = word stopped, weight 0)
 *   bow_n[f], bow_word / bow_value [f*cap + j]   mBowVec in key order (std::map iteration order)
 *   fv_n[f], fv_node [f*cap + j], fv_start [f*(cap+1) + j], fv_feat [f*cap + ..]   mFeatVec in key order: the descriptor
 *   indices of node fv_node[j] are fv_feat[fv_start[j] .. fv_start[j+1]) (ascending, the push_back order). */
int drfe_orb_compute_bow(drfe_orb* h, const drfe_vocab* v, int levelsup, int32_t* word_id, int32_t* node_id,
                         int* bow_n, int32_t* bow_word, double* bow_value, int* fv_n, int32_t* fv_node,
                         int32_t* fv_start, int32_t* fv_feat);

/* ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (ORBmatcher.cc:160-292), whole:
 * frame f of the handle's last batch is F, matched against one keyframe per frame.  For every vocabulary node both
 * FeatureVectors hold (:186-261, a sorted intersection), every keyframe feature with a good map point looks for its
 * best / second best descriptor among F's features of that node that are not matched yet (:200-232), accepts with
 * bestDist1 <= TH_LOW and the ratio test (:234-238), then the rotation histogram (:244-254, 271-289).  The reference's
 * "not matched yet" makes the loop sequential inside a node; the device reaches the same result as the fixed point of
 * parallel sweeps (a warp per node), as drfe_orb_search_last_frame does.  Nodes are independent of each other.
 * Keyframe side (host): kf_n[f] features, kf_desc[(f*kcap + i)*32 ..] (pKF->mDescriptors), kf_angle[f*kcap + i]
 * (pKF->mvKeysUn[i].angle), kf_valid[f*kcap + i] != 0: vpMapPointsKF[i] && !isBad(); its FeatureVector kf_fv_n[f],
 * kf_fv_node[f*kcap + j], kf_fv_start[f*(kcap+1) + j], kf_fv_feat[f*kcap + ..].  Frame side: its FeatureVector in the
 * layout drfe_orb_compute_bow returns (cap = max_keypoints); descriptors and angles are on the device.
 * Outputs (host, any may be NULL): kf_match[f*kcap + i] = bestIdxF of keyframe feature i or -1, BEFORE the rotation check;
 * f_match[f*cap + idx] = the keyframe feature whose map point vpMapPointMatches[idx] holds on return, or -1 (NULL);
 * nmatches[f] = the return value. */
#define DRFE_TH_LOW 50 /* ORBmatcher::TH_LOW (ORBmatcher.cc:39) */
int drfe_orb_search_by_bow(drfe_orb* h, int kcap, const int* kf_n, const uint8_t* kf_desc, const float* kf_angle,
                           const uint8_t* kf_valid, const int* kf_fv_n, const int32_t* kf_fv_node,
                           const int32_t* kf_fv_start, const int32_t* kf_fv_feat, const int* f_fv_n,
                           const int32_t* f_fv_node, const int32_t* f_fv_start, const int32_t* f_fv_feat,
                           float nnratio, int check_orientation, int32_t* kf_match, int32_t* f_match,
                           int* nmatches);

/* mvImagePyramid access (ORBextractor.h:85) and per-stage intermediates, copied to host.
 * bordered != 0 returns the (w+38)x(h+38) buffer including the 19-px BORDER_REFLECT_101
 * frame that ComputePyramid builds (ORBextractor.cc:1107-1132). */
int drfe_orb_level_size(const drfe_orb* h, int level, int* width, int* height);
int drfe_orb_get_pyramid(drfe_orb* h, int frame, int level, int bordered, uint8_t* dst);
int drfe_orb_get_blurred(drfe_orb* h, int frame, int level, uint8_t* dst);
/* FAST candidates of one level: xyr[3*i..] = x, y (region coords, origin (16,16)) and
 * response, in no particular order (compare as a set); returns the count in *n. */
int drfe_orb_get_candidates(drfe_orb* h, int frame, int level, float* xyr, int cap, int* n);
/* quadtree-retained keypoints of one level in the reference's list order (level
 * coordinates, angle filled, not yet scaled to level 0). */
int drfe_orb_get_level_keypoints(drfe_orb* h, int frame, int level, drfe_keypoint* dst, int cap,
                                 int* n);
/* per-stage device time (ms) averaged over the enqueues (at most the last 64) issued since
 * drfe_orb_set_profiling(h, 1); names[i] are static strings.  Reading waits for the stream;
 * recording itself never synchronises. */
int drfe_orb_set_profiling(drfe_orb* h, int on);
int drfe_orb_stage_times(drfe_orb* h, float* ms, const char** names, int cap, int* nstages);

/* ------------------------------------------------------------------ CAPE
 * One handle = one CAPE instance (CAPE.h:47) for a fixed depth size, on one device, able
 * to process up to max_batch independent frames per call. */
int drfe_cape_create(const drfe_cape_params* params, int max_batch, int device, drfe_cape** out);
int drfe_cape_destroy(drfe_cape* h);

/* CAPE::process on cell-major organized clouds (the layout organizePointCloudByCell
 * produces, PlaneExtractor.cpp:80-99): per frame 3*H*W floats, column-major N x 3
 * (all X, then all Y, then all Z), point index = cell_id*cell_w*cell_h + local_r*cell_w +
 * local_c. */
int drfe_cape_enqueue_cloud(drfe_cape* h, int nframes, const float* cloud, size_t frame_stride,
                            int mem_kind);
/* PlaneDetection_CAPE::runPlaneDetection: depth (float, same unit the thresholds are
 * meant in) -> cloud in double -> float, cell-major scatter, then CAPE::process. */
int drfe_cape_enqueue_depth(drfe_cape* h, int nframes, const float* depth, size_t row_stride,
                            size_t frame_stride, int mem_kind, float fx, float fy, float cx,
                            float cy);
/* The same from the sensor's raw 16-bit depth image: z = (float)d * depth_factor on the device,
 * i.e. imDepth.convertTo(imDepth, CV_32F, mDepthMapFactor) of Frame.cc:113-115 (TUM: factor =
 * 1/5000) fused into the cloud kernel; halves the host->device bytes of a depth frame. */
int drfe_cape_enqueue_depth_u16(drfe_cape* h, int nframes, const uint16_t* depth, size_t row_stride,
                                size_t frame_stride, int mem_kind, float depth_factor, float fx,
                                float fy, float cx, float cy);
/* Chunk-pipelined HOST-in / HOST-out batch call (see drfe_orb_extract_batch): depth is float
 * (depth_is_u16 == 0) or raw uint16_t scaled by depth_factor (depth_is_u16 != 0); results as
 * drfe_cape_download delivers them.  Asynchronous; drfe_cape_finish_batch() waits. */
int drfe_cape_process_depth_batch(drfe_cape* h, int nframes, const void* depth, int depth_is_u16,
                                  float depth_factor, size_t row_stride, size_t frame_stride,
                                  float fx, float fy, float cx, float cy, uint8_t* seg_out,
                                  drfe_plane* planes, int plane_cap, int* nr_planes,
                                  drfe_cylinder* cylinders, int cyl_cap, int* nr_cylinders,
                                  int chunk_frames);
int drfe_cape_finish_batch(drfe_cape* h);
/* wait + copy out.  seg_out: nframes * H*W labels (0 = none, 1..n planes, 51.. cylinders),
 * fully written (the reference only writes labelled pixels of a caller-zeroed image).
 * planes[f*plane_cap + i]; cylinders may be NULL when cylinder detection is off.
 * nr_cylinders[f] = nr_cylinders_final (cylinders that survive erosion, CAPE.cpp:340-346);
 * cylinders[f*cyl_cap + i] is cylinder_segments_final (CAPE.cpp:434-445), which holds EVERY
 * cylinder found — drfe_cape_cylinders_found gives that list's length per frame. */
int drfe_cape_download(drfe_cape* h, uint8_t* seg_out, drfe_plane* planes, int plane_cap,
                       int* nr_planes, drfe_cylinder* cylinders, int cyl_cap, int* nr_cylinders);
/* The per-plane point lists PlaneDetection_CAPE::runPlaneDetection builds after CAPE::process
 * (plane_cloud, PlaneExtractor.cpp:165-190; "next" row of SURVEY.md 8f): for every frame of the last
 * batch, the cloud points (x, y, z) of the pixels whose seg_output code is 1..nr_planes, grouped by
 * plane and in row-major pixel order inside a plane, gathered on the device so that the 3.7 MB cloud
 * of a frame never has to cross PCIe.  points[(f*cap_per_frame + k)*3 ..]: plane p of frame f is
 * k in [offsets[f*(plane_cap+1) + p], offsets[f*(plane_cap+1) + p + 1]).  Codes above nr_planes
 * (cylinder labels) are not gathered: the reference indexes plane_cloud out of range for them. */
int drfe_cape_plane_points(drfe_cape* h, float* points, size_t cap_per_frame, int* offsets,
                           int plane_cap);
/* pcl::VoxelGrid<PointT> voxel; voxel.setLeafSize(l, l, l); voxel.filter(*coarseCloud) on every plane_cloud[i], as
 * Frame::ComputePlanes_CAPE does with l = 0.05 right after the plane extraction (reference src/Frame.cc:1121-1125; PCL 1.9
 * voxel_grid.hpp: one centroid per occupied leaf, leaves in ascending index order; a leaf's points are summed in input order —
 * PCL's std::sort leaves that order open).  Runs on the point lists drfe_cape_plane_points builds, without moving them to the
 * host: a 640x480 frame's 1.2 MB of plane points become a few thousand centroids.  Same output layout as drfe_cape_plane_points.
 * A plane whose bounding box has more than INT32_MAX leaves comes back unfiltered, as in PCL. */
int drfe_cape_plane_points_voxel(drfe_cape* h, float leaf_size, float* points, size_t cap_per_frame, int* offsets, int plane_cap);
/* The 1/3-resolution cloud Frame::ComputePlanes_CAPE builds for its surface normals (reference src/Frame.cc:1153-1172) from the
 * depth of the last batch: cloud [nframes][ceil(H/3)][ceil(W/3)][3], z = d > max_point_dist ? 0 : d, x = (n - cx) * z / fx in
 * float.  (The normal estimation itself — pcl::IntegralImageNormalEstimation — is not part of this library.) */
int drfe_cape_third_cloud(drfe_cape* h, float max_point_dist, float* cloud);
/* ... and the surface normals DR-SLAM takes from PCL on that cloud (Frame.cc:1174-1216, :1057-1100: pcl::IntegralImageNormalEstimation,
 * AVERAGE_3D_GRADIENT, setMaxDepthChangeFactor(0.05f), setNormalSmoothingSize(10.0f), border policy and viewpoint at their defaults):
 * normals [nframes][ceil(H/3)][ceil(W/3)][3], NaN where PCL leaves NaN (the border of int(smoothing) points, depth discontinuities,
 * empty windows); cloud as drfe_cape_third_cloud, may be NULL.  PCL is not vendored in the reference: this is its published algorithm
 * (PCL 1.9 integral_image_normal.hpp / integral_image2D.hpp) as restated in oracle/normals_oracle.cpp — parity unpinned by PCL itself. */
int drfe_cape_third_cloud_normals(drfe_cape* h, float max_point_dist, float max_depth_change_factor, float normal_smoothing_size, float* cloud, float* normals);
int drfe_cape_cylinders_found(drfe_cape* h, int* counts); /* counts[f] = cylinder_segments_final.size() */
int drfe_cape_sync(drfe_cape* h);
void* drfe_cape_stream(drfe_cape* h);

/* Single-frame conveniences with the reference's argument meaning. */
int drfe_cape_process(drfe_cape* h, const float* cloud_cellmajor, uint8_t* seg_out,
                      drfe_plane* planes, int plane_cap, int* nr_planes, drfe_cylinder* cylinders,
                      int cyl_cap, int* nr_cylinders);
int drfe_cape_process_depth(drfe_cape* h, const float* depth, size_t row_stride, float fx, float fy,
                            float cx, float cy, uint8_t* seg_out, drfe_plane* planes, int plane_cap,
                            int* nr_planes, drfe_cylinder* cylinders, int cyl_cap,
                            int* nr_cylinders);

/* intermediates for parity tests, copied to host */
int drfe_cape_num_cells(const drfe_cape* h, int* cells_x, int* cells_y);
int drfe_cape_get_cloud(drfe_cape* h, int frame, float* cloud_cellmajor);
/* per-cell PlaneSeg after stage 1 (CAPE.cpp:62-80): cells[cell_id] */
int drfe_cape_get_cells(drfe_cape* h, int frame, drfe_plane* cells);
/* grid_plane_seg_map after region growing + the eroded map used for painting */
int drfe_cape_get_grid_maps(drfe_cape* h, int frame, int32_t* plane_map, uint8_t* eroded_map);
/* cylinder detection on: grid_cylinder_seg_map (cylinder numbers) and its eroded map (labels 50 + k) */
int drfe_cape_get_cyl_maps(drfe_cape* h, int frame, int32_t* cyl_map, uint8_t* cyl_eroded_map);
/* diagnostics of the grid stage of one frame: [0] seeds, [1] growth sweeps, [2] sum of candidates,
 * [3] sum of activated cells, [4..8] cycles in bin search+list / seed scan / growth / accumulate /
 * fit+label, [9] clock after set-up, [10] clock at the end, [11] after the seed loop, [12] after the jobs'
 * accumulate + fit, [13] after the cylinder jobs, [14] after labelling */
int drfe_cape_debug_counters(drfe_cape* h, int frame, long long* out16);
int drfe_cape_set_profiling(drfe_cape* h, int on);
int drfe_cape_stage_times(drfe_cape* h, float* ms, const char** names, int cap, int* nstages);

/* ------------------------------------------------------------------ PEAC-AHC plane extraction (next-1 of SURVEY.md 8f)
 * The plane extractor that is live in Frame::Frame (reference src/Frame.cc:126, :937-949): PlaneDetection::readDepthImage
 * (src/PlaneExtractor.cpp:28-55: 16-bit depth * factor in double, z > 5 culled, X / Y by true double division) and
 * ahc::PlaneFitter<ImagePointCloud>::run with its defaults (include/peac/AHCPlaneFitter.hpp:211-259: 10x10 windows, minSupport
 * 3000, doRefine, ERODE_ALL_BORDER, INIT_STRICT) — block statistics and PCA, graph edges, agglomerative clustering by minimum
 * MSE, block erosion, pixel-level region growing (floodFill), one more clustering pass over the grown planes, relabelling.
 * Declared orders where the reference leaves them to a library (oracle/peac_oracle.cpp P.1 - P.5): equal MSE in the priority
 * queue and equal merge candidates go by creation order, equal plane sizes keep extraction order, Jacobi eigen-solver.
 * One handle = one PlaneDetection object bound to a device, an image size and a maximum batch (frames are independent). */
typedef struct drfe_peac_params {   /* ahc::ParamSet (AHCParamSet.hpp:46-78) + the PlaneFitter members DR-SLAM leaves at their defaults */
  double depthSigma, stdTol_init, stdTol_merge;        /* T_mse = (depthSigma * z^2 + stdTol)^2                                  */
  double z_near, z_far, angle_near, angle_far;         /* T_ang(P_INIT): linear map of the clipped depth to an angle, its cosine */
  double similarityTh_merge, similarityTh_refine;      /* cos 60 deg, cos 30 deg                                                 */
  double depthAlpha, depthChangeTol;                   /* T_dz = depthAlpha * |z| + depthChangeTol                               */
  int32_t min_support, window_width, window_height;    /* 3000, 10, 10                                                           */
  float max_depth;                                     /* readDepthImage culls z > 5.0 (PlaneExtractor.cpp:44)                   */
} drfe_peac_params;
typedef struct drfe_peac_plane {    /* what Frame::ComputePlanes reads of plane_filter.extractedPlanes[i] (Frame.cc:971-979), and its statistics */
  double normal[3], center[3], mse, curvature;
  int32_t N, rid;
} drfe_peac_plane;
typedef struct drfe_peac drfe_peac;
int drfe_peac_default_params(drfe_peac_params* p);
int drfe_peac_create(int width, int height, const drfe_peac_params* params, int max_batch, int device, drfe_peac** out);
int drfe_peac_destroy(drfe_peac* h);
void* drfe_peac_stream(drfe_peac* h);
int drfe_peac_sync(drfe_peac* h);
/* planeDetector.readDepthImage(Depth, K, depthFactor); planeDetector.runPlaneDetection(); for every frame of a batch
 * (asynchronous; depth: [nframes][H][W] uint16, host or device) */
int drfe_peac_enqueue_depth_u16(drfe_peac* h, int nframes, const uint16_t* depth, size_t row_stride, size_t frame_stride, int mem_kind,
                                float depth_factor, float fx, float fy, float cx, float cy);
/* seg_output (plid + 1 per pixel, 0 elsewhere; [nframes][H][W]), extractedPlanes ([nframes][plane_cap]) and plane_num_ */
int drfe_peac_download(drfe_peac* h, uint8_t* seg_out, drfe_peac_plane* planes, int plane_cap, int* nr_planes);
/* plane_vertices_ (pixel indices of every plane, in scan order) and the points Frame::ComputePlanes reads for them
 * (Frame.cc:954-963: (float) of cloud.vertices[j]); layout as drfe_cape_plane_points; indices or points may be null */
int drfe_peac_plane_vertices(drfe_peac* h, int32_t* indices, float* points, size_t cap_per_frame, int* offsets, int plane_cap);
/* What Frame::ComputePlanes does with every plane right after the extraction (Frame.cc:954-990): the vertices with
 * (float) z <= max_point_dist (mMax_point_dist; FLT_MAX keeps all), through pcl::VoxelGrid with setLeafSize(leaf, leaf, leaf) —
 * the declared summation order of drfe_cape_plane_points_voxel applies.  The lists stay on the device, the centroids come back:
 * points [nframes][cap_per_frame][3], offsets as drfe_peac_plane_vertices */
int drfe_peac_plane_points_voxel(drfe_peac* h, float max_point_dist, float leaf_size, float* points, size_t cap_per_frame, int* offsets, int plane_cap);
/* the 1/3-resolution cloud Frame::ComputePlanes builds from imDepth (Frame.cc:1044-1066) and the surface normals it takes from
 * pcl::IntegralImageNormalEstimation on it (:1068-1100) — as drfe_cape_third_cloud / drfe_cape_third_cloud_normals, from the depth batch
 * this handle was given; cloud may be NULL */
int drfe_peac_third_cloud_normals(drfe_peac* h, float max_point_dist, float max_depth_change_factor, float normal_smoothing_size, float* cloud, float* normals);
/* diagnostics of a frame: [0] clustering steps (both passes), [1] region-growing queue length, [2] planes of the first pass,
 * [4..10] kilocycles from the frame's start to the end of: reset, edges, first clustering, block membership + seeds, region
 * growing, second clustering, outputs */
int drfe_peac_debug_counters(drfe_peac* h, int frame, int32_t* out12);

/* ------------------------------------------------------------------ input resize (next-4 of SURVEY.md 8f)
 * cv::resize(im, IM, Size(640,480)) and cv::resize(depthmap, Depthmap, Size(640,480)) of System::TrackRGBD (reference
 * src/System.cc:325-329; default INTER_LINEAR), batched: OpenCV's own arithmetic — 11-bit fixed point for 8U (1, 3 or 4
 * channels), separately rounded float products for 16U (saturate_cast of cvRound) and 32F.  Source and destination may each be
 * host or device memory; with a device destination the call is asynchronous on the resizer's stream (drfe_resizer_sync, or
 * drfe_stream_wait_event on drfe_resizer_stream, before another handle consumes it — e.g. drfe_orb_enqueue_color or
 * drfe_cape_enqueue_depth_u16 with DRFE_MEM_DEVICE).  Strides are in BYTES. */
#define DRFE_PIX_U8 0
#define DRFE_PIX_U16 2
#define DRFE_PIX_F32 5
typedef struct drfe_resizer drfe_resizer;
int drfe_resizer_create(int src_width, int src_height, int dst_width, int dst_height, int max_batch, int device, drfe_resizer** out);
int drfe_resizer_destroy(drfe_resizer* h);
void* drfe_resizer_stream(drfe_resizer* h);
int drfe_resizer_sync(drfe_resizer* h);
int drfe_resize(drfe_resizer* h, int nframes, const void* src, int pixel_type, int channels, size_t src_row_stride, size_t src_frame_stride,
                int src_mem_kind, void* dst, size_t dst_row_stride, size_t dst_frame_stride, int dst_mem_kind);

/* ------------------------------------------------------------------ several devices
 * Frames are independent (ORBextractor keeps no state across frames, Frame.cc:124-134 builds a fresh
 * PlaneDetection per frame), so a pool cuts a batch of host frames into contiguous blocks, one per device; every
 * device has its own ORB + CAPE handle pair, its own streams and staging arenas, and its own persistent host thread
 * that drives the chunk-pipelined batch calls; results land in the caller's arrays BY FRAME INDEX.  No collective
 * and no peer traffic.  The same device may be listed more than once (two workers sharing a GPU).
 * What a caller with G GPUs replaces: the per-frame std::thread pair of Frame::Frame (Frame.cc:124-134) run over a
 * recorded sequence (Examples/RGB-D/rgbd_tum.cc:76-115 feeds frames one by one). */
typedef struct drfe_pool drfe_pool;
typedef struct drfe_pool_params {
  drfe_orb_params orb;
  drfe_cape_params cape;       /* depth_width / depth_height are taken from width / height below */
  int32_t width, height;       /* image size, the same for gray and depth */
  int32_t max_batch;           /* most frames one drfe_pool_extract_batch call may carry, over all devices */
  int32_t chunk_frames;        /* frames per pipelined chunk on a device, 0 = library default */
} drfe_pool_params;
int drfe_pool_create(const drfe_pool_params* params, const int* devices, int ndevices, drfe_pool** out);
int drfe_pool_destroy(drfe_pool* p);
int drfe_pool_num_devices(const drfe_pool* p);
int drfe_pool_max_keypoints(const drfe_pool* p);     /* = drfe_orb_max_keypoints of every device's handle */
/* One batch, host in / host out; arguments as drfe_orb_extract_batch + drfe_cape_process_depth_batch, every array
 * indexed by frame (kps [nframes][cap_per_frame], desc [nframes][cap_per_frame][32], seg_out [nframes][H][W],
 * planes [nframes][plane_cap], cylinders [nframes][cyl_cap]; any output except counts / nr_planes may be null).
 * Returns when every device has delivered its block; the first failing device's status and text otherwise. */
int drfe_pool_extract_batch(drfe_pool* p, int nframes, const uint8_t* gray, size_t gray_row_stride, size_t gray_frame_stride,
                            const void* depth, int depth_is_u16, float depth_factor, size_t depth_row_stride,
                            size_t depth_frame_stride, float fx, float fy, float cx, float cy, drfe_keypoint* kps, uint8_t* desc,
                            int cap_per_frame, int* counts, uint8_t* seg_out, drfe_plane* planes, int plane_cap, int* nr_planes,
                            drfe_cylinder* cylinders, int cyl_cap, int* nr_cylinders);
/* device time (CUDA events, first H2D to last D2H) each device spent on its block of the last batch */
int drfe_pool_device_times(const drfe_pool* p, float* ms, int cap);
/* Pinned host memory for the batch calls (pageable buffers make every copy synchronous): allocate, or register a
 * buffer the application already owns (e.g. the cv::Mat data of a recorded sequence). */
int drfe_host_alloc(void** ptr, size_t bytes, int write_combined);
int drfe_host_free(void* ptr);
int drfe_host_register(void* ptr, size_t bytes);
int drfe_host_unregister(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* DRFE_H_ */
