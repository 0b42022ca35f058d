// TEST INFRASTRUCTURE — CPU oracle, not the product.
//
// One C API, exported twice:
//   * oracle/_ref/liborb_ref.so   (prefix ref_)  — the reference's own translation units, compiled
//     where they lie under /root/reference against oracle/cvshim (see oracle/Makefile);
//   * oracle/_build/liborb_oracle.so (prefix orc_) — the stand-alone restatement in orb_oracle.cc,
//     which is what travels to machines without /root/reference.
// tests/ cross-check the two wherever both exist, then check the CUDA path against the restatement.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load these libraries.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifndef ORACLE_PREFIX
#error "define ORACLE_PREFIX (ref_ or orc_)"
#endif
#define ORACLE_CAT2(a, b) a##b
#define ORACLE_CAT(a, b) ORACLE_CAT2(a, b)
#define ORACLE_FN(name) ORACLE_CAT(ORACLE_PREFIX, name)

#ifdef __cplusplus
extern "C" {
#endif

// mirrors cv::KeyPoint field for field (28 bytes)
typedef struct oracle_keypoint
{
	float x, y, size, angle, response;
	int32_t octave, class_id;
} oracle_keypoint;

// packed FAST candidate / selected keypoint in level coordinates
typedef struct oracle_cand
{
	int32_t x, y, response;
} oracle_cand;

typedef struct oracle_camera
{
	float fx, fy, cx, cy, bf, baseline;
} oracle_camera;

// ---- extractor object (include/ORBextractor.h:34-80) ----
void* ORACLE_FN(extractor_create)(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void ORACLE_FN(extractor_destroy)(void* ex);
// Extract (src/ORBextractor.cc:743-820). Returns the keypoint count N (<= cap, else -needed),
// or -1 when the input contract is violated (the reference throws cv::Exception / divides by zero).
int ORACLE_FN(extractor_extract)(void* ex, const uint8_t* img, int w, int h, size_t pitch,
                                  oracle_keypoint* kps, uint8_t* desc, int cap);
int ORACLE_FN(extractor_level_size)(void* ex, int level, int* w, int* h);
int ORACLE_FN(extractor_level_copy)(void* ex, int level, uint8_t* dst, size_t pitch);
// scale tables, 4 x nlevels floats: scaleFactors, invScaleFactors, sigmaSq, invSigmaSq
void ORACLE_FN(extractor_tables)(void* ex, float* scale, float* inv_scale, float* sigma_sq, float* inv_sigma_sq);
// per-level quotas (ComputeNumFeaturesPerScale, src/ORBextractor.cc:472-487)
void ORACLE_FN(feature_quotas)(int nfeatures, float scaleFactor, int nlevels, int* out);

// ---- stages, stateless ----
// DetectFAST (src/ORBextractor.cc:489-540) with the roi Extract uses (16 px border, :755,760).
int ORACLE_FN(detect_fast)(const uint8_t* img, int w, int h, size_t pitch, int iniTh, int minTh,
                            oracle_cand* out, int cap);
// QuadTreeSuppression (src/ORBextractor.cc:542-693); roi as above from (w,h).
int ORACLE_FN(quadtree)(const oracle_cand* in, int n, int w, int h, int nfeatures, oracle_cand* out, int cap);
// IC_Angle (src/ORBextractor.cc:74-101)
float ORACLE_FN(ic_angle)(const uint8_t* img, int w, int h, size_t pitch, int x, int y);
// ComputeOrbDescriptor (src/ORBextractor.cc:103-140) on an already blurred level
void ORACLE_FN(descriptor)(const uint8_t* blurred, int w, int h, size_t pitch, int x, int y, float angle_deg,
                           uint8_t* desc32);

// ---- matchers ----
// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1449-1457)
int ORACLE_FN(descriptor_distance)(const uint8_t* a, const uint8_t* b);
// ComputeStereoMatches (src/ORBmatcher.cc:72-247). Pyramid levels are given as parallel arrays.
// Returns 0, or -1 when the reference would index an empty vector (no surviving match, :232-233);
// uright/depth are then all -1.
int ORACLE_FN(stereo_matches)(const oracle_keypoint* kpL, int nL, const uint8_t* descL,
                               const uint8_t* const* pyrL, const oracle_keypoint* kpR, int nR, const uint8_t* descR,
                               const uint8_t* const* pyrR, const int* level_w, const int* level_h,
                               const size_t* level_pitch, int nlevels, const float* scale, const float* inv_scale,
                               const oracle_camera* cam, float* uright, float* depth);
// best/second scan of SearchByBoW (src/ORBmatcher.cc:477-507): restated in both builds (the
// reference has no stand-alone entry point for it). match[q] = idx if accepted by
// best <= th_low && (float)best < nnratio*(float)second, else -1. Uses `threads` host threads.
void ORACLE_FN(knn2)(const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt, int th_low, float nnratio,
                     int32_t* idx, uint16_t* best, uint16_t* second, int32_t* match, int threads);

// (float)cos / (float)sin of angle * factorPI exactly as ComputeOrbDescriptor forms them (src/ORBextractor.cc:105-107), for the n float angles with
// bit patterns first_bits, first_bits + 1, ... — the host side of the exhaustive device-vs-glibc sweep
void ORACLE_FN(cos_sin_range)(uint32_t first_bits, int64_t n, float* c, float* s, int threads);

// ---- rows "next" of SURVEY §8(f) ----
// ConvertToGray (src/System.cc:122-137): channels 3 or 4, rgb != 0 when the first channel is R
void ORACLE_FN(convert_to_gray)(const uint8_t* src, int w, int h, size_t pitch, int channels, int rgb, uint8_t* dst, size_t dst_pitch);
// ComputeStereoFromRGBD (src/System.cc:197-219): depth map is float32, pitch in bytes
void ORACLE_FN(stereo_from_rgbd)(const oracle_keypoint* kps, const oracle_keypoint* kps_un, int n, const float* depth_map, int w, int h,
                                 size_t pitch, const oracle_camera* cam, float* uright, float* depth);
// UndistortKeyPoints (src/System.cc:153-174): cv::undistortPoints with P = K; dist has ndist (4, 5, 8, 12) coefficients; kps_un = kps when
// dist[0] == 0
void ORACLE_FN(undistort_keypoints)(const oracle_keypoint* kps, int n, const oracle_camera* cam, const float* dist, int ndist, oracle_keypoint* kps_un);
// the distance matrix + least-median selection of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:286-314): index of the
// descriptor with the least median distance to the others, first wins ties
int ORACLE_FN(distinctive_index)(const uint8_t* desc, int n);

// ---- guided matchers, SURVEY §8(f) #1 ----
typedef struct oracle_bounds { float minx, maxx, miny, maxy; } oracle_bounds;
// a Frame as the guided matchers see it (include/Frame.h:83-168): undistorted keypoints, descriptors, right coordinates (NULL = all -1),
// image bounds, pyramid scale factors
typedef struct oracle_frame_view
{
	int32_t n;
	const oracle_keypoint* kps_un;
	const uint8_t* desc;
	const float* uright;
	oracle_bounds bounds;
	int32_t nlevels;
	const float* scale_factors;
} oracle_frame_view;
typedef struct oracle_pose { float R[9]; float t[3]; } oracle_pose;   // CameraPose (include/CameraPose.h), R row-major
// one candidate map point of SearchByProjection(Frame&, mappoints, th): the track* scratch (include/MapPoint.h:92-97);
// flags bit0 = trackInView && !isBad(), bit1 = Observations() > 0
typedef struct oracle_track_point { float proj_x, proj_y, proj_xr, view_cos; int32_t scale_level; int32_t flags; } oracle_track_point;
// one keypoint of the last frame for SearchByProjection(currFrame, lastFrame, ...): world position of its map point, octave of
// lastFrame.keypoints[i], angle of lastFrame.keypointsUn[i]; flags bit0 = has a map point && !outlier, bit1 = Observations() > 0
typedef struct oracle_last_point { float xw[3]; int32_t octave; float angle; int32_t flags; } oracle_last_point;

// FeaturesGrid (src/Frame.cc:65-145)
void* ORACLE_FN(grid_create)(const oracle_keypoint* kps, int n, const oracle_bounds* bounds, int nlevels);
void ORACLE_FN(grid_destroy)(void* grid);
// GetFeaturesInArea; returns the count (indices beyond cap are dropped)
int ORACLE_FN(grid_query)(void* grid, float x, float y, float r, int min_level, int max_level, int32_t* out, int cap);
// The three matchers. frame_mp (n entries, in/out) is frame.mappoints: -1 = null, >= 0 = index of a point of `pts`, -2 = some other
// map point with Observations() > 0, -3 = some other map point without observations. All return nmatches.
int ORACLE_FN(search_local_map)(const oracle_frame_view* frame, int32_t* frame_mp, const oracle_track_point* pts, const uint8_t* pt_desc,
                                int npts, float th, float nnratio);
int ORACLE_FN(search_last_frame)(const oracle_frame_view* cur, const oracle_camera* cam, const oracle_pose* cur_pose,
                                 const oracle_pose* last_pose, int32_t* frame_mp, const oracle_last_point* pts, const uint8_t* pt_desc,
                                 int npts, float th, int monocular, float nnratio, int check_orientation);
// one map point of a key frame for SearchByProjection(Frame&, KeyFrame*, alreadyFound, th, ORBdist) (src/ORBmatcher.cc:1364-1447): world
// position, the members minDistance_ / maxDistance_ behind Get{Min,Max}DistanceInvariance and PredictScale, the angle of the key frame's
// keypoint (CheckOrientation); flags bit0 = map point present && !isBad() && not in alreadyFound
typedef struct oracle_kf_point { float xw[3]; float min_distance, max_distance; float angle; int32_t flags; } oracle_kf_point;
// frame_mp as for the other searches (any non-null entry closes its keypoint here, :1412-1413); log_scale_factor = pyramid.logScaleFactor
int ORACLE_FN(search_keyframe_projection)(const oracle_frame_view* frame, const oracle_camera* cam, const oracle_pose* pose, float log_scale_factor,
                                          int32_t* frame_mp, const oracle_kf_point* pts, const uint8_t* pt_desc, int npts, float th, int orb_dist,
                                          int check_orientation);
// one candidate map point of SearchByProjection(keyframe, Scw, mappoints, matched, th) (src/ORBmatcher.cc:518-612, loop closing): world
// position, viewing normal (GetNormal), minDistance_ / maxDistance_; flags bit0 = !isBad() && not already in `matched`
typedef struct oracle_sim3_point { float xw[3]; float normal[3]; float min_distance, max_distance; int32_t flags; } oracle_sim3_point;
typedef struct oracle_sim3 { float R[9]; float t[3]; float s; } oracle_sim3;
// matched (kf->n entries, in/out): -1 = null, -2 = some map point found before the call, >= 0 = index of the point stored by this call
int ORACLE_FN(search_sim3_projection)(const oracle_frame_view* kf, const oracle_camera* cam, const oracle_sim3* Scw, float log_scale_factor,
                                      int32_t* matched, const oracle_sim3_point* pts, const uint8_t* pt_desc, int npts, int th);
// DBoW2::FeatureVector (Thirdparty/DBoW2/DBoW2/FeatureVector.h) as CSR: node ids ascending, the feature indices of node k are
// indices[start[k] .. start[k+1])
typedef struct oracle_feature_vector { int32_t nnodes; const uint32_t* node_ids; const int32_t* start; const uint32_t* indices; } oracle_feature_vector;
// SearchByBoW: KeyFrame vs Frame (src/ORBmatcher.cc:452-516) when valid2 == NULL, KeyFrame vs KeyFrame (:696-766) otherwise.
// valid1 / valid2: per keypoint, map point present && !isBad(). match2 (f2->n entries, out): the keypoint of frame 1 matched to
// keypoint idx2 of frame 2, or -1. Returns nmatches.
int ORACLE_FN(search_by_bow)(const oracle_frame_view* f1, const oracle_feature_vector* fv1, const uint8_t* valid1, const oracle_frame_view* f2,
                             const oracle_feature_vector* fv2, const uint8_t* valid2, float nnratio, int check_orientation, int32_t* match2);
// CPU-baseline timing of the two tracking searches: the Frame (grid included) and the map points are built once, outside the timed
// region; each repetition restores frame.mappoints and times the matcher call alone. Returns mean seconds per call.
double ORACLE_FN(time_search_local_map)(const oracle_frame_view* frame, const int32_t* frame_mp, const oracle_track_point* pts,
                                        const uint8_t* pt_desc, int npts, float th, float nnratio, int reps);
double ORACLE_FN(time_search_last_frame)(const oracle_frame_view* cur, const oracle_camera* cam, const oracle_pose* cur_pose,
                                         const oracle_pose* last_pose, const int32_t* frame_mp, const oracle_last_point* pts,
                                         const uint8_t* pt_desc, int npts, float th, int monocular, float nnratio, int check_orientation,
                                         int reps);
// prev_matched: n1 x 2 floats in/out; matches12: n1 out
int ORACLE_FN(search_for_initialization)(const oracle_frame_view* f1, const oracle_frame_view* f2, float* prev_matched, int32_t* matches12,
                                         int window, float nnratio, int check_orientation);

// ---- the matchers of local mapping / loop closing whose per-point search is independent (second half) ----
// One-key-frame model of the map state Fuse mutates (see ref_guided_decl.h): kf_mp[N] = point index held by a keypoint or -1 (in/out),
// nobs / bad / in_kf per point (in/out). log (cap ints, out): the mutations in order as triples (1, point, other point) = Replace,
// (2, point, keypoint) = AddObservation + AddMapPoint; *nlog = ints written. Returns nfused.
int ORACLE_FN(fuse)(const oracle_frame_view* kf, const oracle_camera* cam, const oracle_pose* pose, float log_scale_factor, const float* inv_sigma_sq,
                    const oracle_sim3_point* pts, const uint8_t* pt_desc, int npts, float th, int32_t* kf_mp, int32_t* nobs, uint8_t* bad,
                    uint8_t* in_kf, int32_t* log, int cap, int* nlog);
// Fuse(keyframe, Scw, mappoints, th, replacePoints): replace[npts] out = point index held by the matched keypoint or -1
int ORACLE_FN(fuse_sim3)(const oracle_frame_view* kf, const oracle_camera* cam, const oracle_sim3* Scw, float log_scale_factor,
                         const oracle_sim3_point* pts, const uint8_t* pt_desc, int npts, float th, int32_t* kf_mp, int32_t* nobs, uint8_t* bad,
                         int32_t* replace, int32_t* log, int cap, int* nlog);
// SearchBySim3: pts1 / pts2 one per keypoint (flags bit0 = map point present && !bad; bit1 = already in matches12 on entry, with
// GetIndexInKeyFrame(kf2) = its own index... see the wrapper). matches12[n1] out = keypoint of kf2 or -1. Returns nfound.
int ORACLE_FN(search_by_sim3)(const oracle_frame_view* kf1, const oracle_camera* cam1, const oracle_pose* pose1, float lsf1,
                              const oracle_frame_view* kf2, const oracle_camera* cam2, const oracle_pose* pose2, float lsf2, const oracle_sim3* S12,
                              float th, const oracle_kf_point* pts1, const uint8_t* desc1, const oracle_kf_point* pts2, const uint8_t* desc2,
                              int32_t* matches12);
// SearchForTriangulation: has1 / has2 per keypoint = GetMapPoint(idx) != NULL; F12 row-major; the epipole is computed by the reference
// text from pose2 / cam2 / Ow1 when ep2 == NULL (ref library only), else taken as given. matches12[n1] out. Returns nmatches.
int ORACLE_FN(search_for_triangulation)(const oracle_frame_view* kf1, const oracle_feature_vector* fv1, const uint8_t* has1,
                                        const oracle_frame_view* kf2, const oracle_feature_vector* fv2, const uint8_t* has2, const float* F12,
                                        const float* ep2, const float* sigma_sq2, int only_stereo, int check_orientation, int32_t* matches12);

// ---- bag-of-words transform: ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> ----
// loadFromTextFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.cpp:21-90); NULL when the file is rejected
void* ORACLE_FN(voc_load_text)(const char* path);
void ORACLE_FN(voc_destroy)(void* voc);
// transform(features, BowVector, FeatureVector, levelsup) (TemplatedVocabulary.h:1129-1197). Outputs have room for n entries
// (fv_start: n + 1): the BowVector as ascending (word id, value), the FeatureVector as CSR (see oracle_feature_vector). Returns the number of words.
int ORACLE_FN(bow_transform)(void* voc, const uint8_t* desc, int n, int levelsup, int32_t* word_ids, double* word_vals, uint32_t* fv_nodes,
                             int32_t* fv_start, uint32_t* fv_items, int32_t* n_fv_nodes);
// TemplatedVocabulary::score = the scoring object's score (L1Scoring::score, ScoringObject.cpp:24-58, for the reference's vocabulary)
double ORACLE_FN(bow_score)(void* voc, const int32_t* ida, const double* va, int na, const int32_t* idb, const double* vb, int nb);
// mean seconds per transform call over reps repetitions (CPU baseline)
double ORACLE_FN(time_bow_transform)(void* voc, const uint8_t* desc, int n, int levelsup, int reps);

// ---- pinned third-party primitives (same code in both libraries; checked against cv2 4.13.0) ----
void ORACLE_FN(cv_resize)(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep);
int ORACLE_FN(cv_fast)(const uint8_t* img, int w, int h, size_t step, int th, int nms, oracle_cand* out, int cap);
void ORACLE_FN(cv_gaussian7)(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep);
// cv::remap(INTER_LINEAR) of an 8-bit image through two float maps (Examples/Stereo/stereo_euroc.cc:100-101); map pitch in bytes
void ORACLE_FN(cv_remap)(const uint8_t* src, int sw, int sh, size_t sstep, const float* mapx, const float* mapy, size_t mstep, uint8_t* dst,
                         int dw, int dh, size_t dstep);
float ORACLE_FN(cv_fast_atan2)(float y, float x);
int ORACLE_FN(cv_round_f)(float v);
int ORACLE_FN(cv_round_d)(double v);

#ifdef __cplusplus
}
#endif
