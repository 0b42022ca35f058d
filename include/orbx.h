/* orbx.h — C ABI of the B200-native ORB front-end (liborbx_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path this repository accelerates: the reference's
 * ORB feature extraction and Hamming matching. The reference has no FFI layer; its
 * boundary is a handful of C++ symbols called from src/System.cc. Each entry point below names the
 * reference interface it replaces. The C++ mirror of those symbols (same class/method names) lives
 * in include/orbx/ORBextractor.h and include/orbx/ORBmatcher.h and forwards to this ABI; see
 * INTEGRATION.md for the three-line change in src/System.cc.
 *
 * Conventions: plain pointers and sizes only; every function returns an orbx_status; nothing throws
 * across the boundary. All work runs on hand-written sm_100a CUDA kernels — there is no CPU
 * fallback: on a machine without a usable GPU every call returns ORBX_ERR_CUDA.
 * "host" pointers are ordinary (ideally pinned) host memory; "_device" variants take device
 * pointers valid on the handle's device and enqueue on the handle's stream without copies.
 */
#ifndef ORBX_H
#define ORBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum orbx_status
{
	ORBX_OK = 0,
	ORBX_ERR_INVALID = 1,    /* argument outside the reference's own input contract (it would throw / divide by zero) */
	ORBX_ERR_CUDA = 2,       /* CUDA runtime error or no sm_100 device; orbx_last_error() has the text */
	ORBX_ERR_CAPACITY = 3,   /* caller's output buffer too small; required size is reported */
	ORBX_ERR_STATE = 4       /* call order violated (e.g. stereo before extract) */
} orbx_status;

/* ORBextractor::Parameters — include/ORBextractor.h:38-47 (defaults 2000 / 1.2f / 8 / 20 / 7) */
typedef struct orbx_params
{
	int32_t nfeatures;
	float scale_factor;
	int32_t nlevels;
	int32_t ini_th_fast;
	int32_t min_th_fast;
} orbx_params;

/* cv::KeyPoint, field for field (pt.x, pt.y, size, angle, response, octave, class_id) — 28 bytes.
 * A std::vector<cv::KeyPoint> buffer can be handed in directly. */
typedef struct orbx_keypoint
{
	float x, y, size, angle, response;
	int32_t octave, class_id;
} orbx_keypoint;

/* CameraParams — include/CameraParameters.h:29-40 */
typedef struct orbx_camera
{
	float fx, fy, cx, cy, bf, baseline;
} orbx_camera;

typedef struct orbx_extractor* orbx_handle;

const char* orbx_last_error(void);
/* number of CUDA devices with compute capability 10.x; 0 when none (or no driver) */
int orbx_device_count(void);

/* ---- ORBextractor (include/ORBextractor.h:34-80, src/ORBextractor.cc:695-835) ---- */

/* ORBextractor::ORBextractor(const Parameters&) + Init() — src/ORBextractor.cc:695-741.
 * `device` is the CUDA ordinal; one handle owns one stream and one scratch arena, so — like the
 * reference instance (include/ORBextractor.h:74-76) — a handle is not re-entrant, while different
 * handles may run concurrently (src/System.cc:449-452). */
orbx_status orbx_create(const orbx_params* params, int device, orbx_handle* out);
orbx_status orbx_destroy(orbx_handle h);

/* GetLevels / GetScaleFactor / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares — include/ORBextractor.h:57-62. Each array holds nlevels floats; any may be NULL. */
orbx_status orbx_get_params(orbx_handle h, orbx_params* out);
orbx_status orbx_scale_tables(orbx_handle h, float* scale, float* inv_scale, float* sigma_sq, float* inv_sigma_sq);
/* per-level keypoint quotas — ComputeNumFeaturesPerScale, src/ORBextractor.cc:472-487 */
orbx_status orbx_feature_quotas(orbx_handle h, int32_t* quotas);

/* ORBextractor::Extract(image, keypoints, descriptors) — src/ORBextractor.cc:743-820, one frame,
 * host buffers. `kps` holds `cap` entries, `desc` cap*32 bytes (the N x 32 CV_8U matrix, row i = keypoint i).
 * *n = keypoint count; entries of kps/desc at and beyond *n are unspecified. When N == 0 the reference releases
 * `descriptors` and leaves `keypoints` untouched (:778-782); here *n = 0. ORBX_ERR_CAPACITY sets *n to the needed count. */
orbx_status orbx_extract(orbx_handle h, const uint8_t* image, int width, int height, size_t pitch,
                         orbx_keypoint* kps, uint8_t* desc, int cap, int* n);

/* The same for a batch of `frames` equally sized images (frame f at images + f*frame_stride). Outputs are
 * frame-major: kps + f*cap, desc + f*cap*32, n[f]. This is the throughput entry point: the batch is cut into chunks
 * that flow through two CUDA streams, so the upload of one chunk and the download of another overlap the kernels of a
 * third; inside a chunk every kernel is launched once per pyramid level (or once) for all its frames. Use pinned host
 * memory for the overlap to materialise. */
orbx_status orbx_extract_batch(orbx_handle h, const uint8_t* images, int frames, int width, int height, size_t pitch,
                               size_t frame_stride, orbx_keypoint* kps, uint8_t* desc, int cap, int* n);

/* Device-resident variant: inputs already in HBM, outputs stay in HBM (d_kps: frames*cap keypoints, d_desc:
 * frames*cap*32 bytes, d_n: frames int32). Asynchronous on the handle's stream; call orbx_synchronize (or use
 * orbx_stream) before reading. cap must be >= orbx_max_keypoints(h). When d_images, pitch and frame_stride are multiples of
 * 16 bytes the frames are read in place and ARE level 0 of the pyramid until the next extract on this handle
 * (orbx_pyramid_level(_device) of level 0 and orbx_stereo_match read them): keep them unchanged that long. Other layouts
 * are copied into an internal buffer first (ComputePyramid's own copyTo, src/ORBextractor.cc:462). */
orbx_status orbx_extract_batch_device(orbx_handle h, const uint8_t* d_images, int frames, int width, int height,
                                      size_t pitch, size_t frame_stride, orbx_keypoint* d_kps, uint8_t* d_desc,
                                      int cap, int32_t* d_n);
/* upper bound of keypoints per frame: sum over levels of (quota + 3) (the quadtree overshoots by at most 3). Before the
 * handle has seen an image size the bound assumes the widest aspect ratio; orbx_plan makes it exact. */
int orbx_max_keypoints(orbx_handle h);
/* Builds (or re-uses) the plan for `frames` images of width x height without extracting: level sizes, tables, device
 * buffers. What every extract call does first; call it to size output buffers exactly (ORBextractor's constructor plus the
 * first image's size, src/ORBextractor.cc:695-741, :455-470). */
orbx_status orbx_plan(orbx_handle h, int width, int height, int frames);
/* Shape of the device-resident result of the last extract call on this handle: number of frames and the row capacity
 * (keypoints per frame) of its keypoint / descriptor / uright / depth arrays. ORBX_ERR_STATE before the first extract. */
orbx_status orbx_last_result_shape(orbx_handle h, int* frames, int* cap);
orbx_status orbx_synchronize(orbx_handle h);
/* the handle's cudaStream_t, as void* */
void* orbx_stream(orbx_handle h);

/* Per-stage device time, for roofline accounting (bench.py). While enabled, every extract call brackets its stages with
 * CUDA events on the handle's stream; orbx_stage_times waits for the stream and returns the SUM over the calls since the
 * last query, in milliseconds: [0] pyramid (levels 1..n-1), [1] FAST cells, [2] quadtree selection, [3] blur, [4] orientation +
 * descriptor; *calls = number of extract calls summed. */
orbx_status orbx_enable_stage_timing(orbx_handle h, int enable);
orbx_status orbx_stage_times(orbx_handle h, float ms_sum[5], int* calls);

/* ORBextractor::GetImagePyramid() — include/ORBextractor.h:63. The pyramid of the last extract call stays on
 * the device; a level is downloaded only on request. Valid until the next extract on this handle. */
orbx_status orbx_level_size(orbx_handle h, int level, int* width, int* height);
orbx_status orbx_pyramid_level(orbx_handle h, int frame, int level, uint8_t* dst, size_t dst_pitch);
/* device view of the same: base pointer of `level` of `frame`, and its pitch in bytes */
orbx_status orbx_pyramid_level_device(orbx_handle h, int frame, int level, const uint8_t** d_ptr, size_t* pitch);

/* Stage probes for parity tests (same stage boundaries as the oracle): results of the last extract call.
 * candidates: DetectFAST output order (src/ORBextractor.cc:489-540); selected: QuadTreeSuppression output order
 * (:542-693); both as (x, y, response) int32 triples in level coordinates. blurred: cv::GaussianBlur result (:799). */
orbx_status orbx_debug_candidates(orbx_handle h, int frame, int level, int32_t* xyr, int cap, int* n);
orbx_status orbx_debug_selected(orbx_handle h, int frame, int level, int32_t* xyr, int cap, int* n);
orbx_status orbx_debug_blurred_level(orbx_handle h, int frame, int level, uint8_t* dst, size_t dst_pitch);
/* The rotation terms of ComputeOrbDescriptor — src/ORBextractor.cc:105-107: (float)cos / (float)sin of angle * factorPI — as the
 * descriptor kernel forms them, for the n float angles (degrees) whose bit patterns are first_bits, first_bits + 1, ...; host buffers.
 * Exists so that a test can compare EVERY float angle in [0, 360) with the host's libm. */
orbx_status orbx_debug_cos_sin(int device, uint32_t first_bits, int64_t n, float* cos_out, float* sin_out);

/* ---- Matching (include/ORBmatcher.h, src/ORBmatcher.cc) ---- */

/* ORBmatcher::DescriptorDistance — src/ORBmatcher.cc:1449-1457, for n descriptor pairs (a[i], b[i]), 32 bytes
 * each, host buffers. The single-pair call of the reference is n = 1. */
orbx_status orbx_descriptor_distance(int device, const uint8_t* a, const uint8_t* b, int64_t n, int32_t* dist);

/* ComputeStereoMatches — include/ORBmatcher.h:41-45, src/ORBmatcher.cc:72-247 — on the device-resident results
 * of the last extract of `left` and `right` (same batch size, same image size, same device). uright/depth: host,
 * frames*cap floats laid out like the keypoints of that extract (frames, cap = orbx_last_result_shape(left)); entries
 * past n[f] are not written. No keypoint/descriptor/pyramid traffic leaves the GPU. */
orbx_status orbx_stereo_match(orbx_handle left, orbx_handle right, const orbx_camera* camera, float* uright, float* depth);
orbx_status orbx_stereo_match_device(orbx_handle left, orbx_handle right, const orbx_camera* camera, float* d_uright,
                                     float* d_depth);
/* The same from host-side data, exactly the reference's argument list (one stereo pair): keypoints, descriptors
 * and 8-bit pyramids of both images. Levels are passed as arrays of base pointers / widths / heights / pitches. */
orbx_status orbx_stereo_match_host(int device, const orbx_keypoint* kps_l, int n_l, const uint8_t* desc_l,
                                   const uint8_t* const* pyr_l, const orbx_keypoint* kps_r, int n_r,
                                   const uint8_t* desc_r, const uint8_t* const* pyr_r, const int* level_w,
                                   const int* level_h, const size_t* level_pitch, int nlevels, const float* scale,
                                   const float* inv_scale, const orbx_camera* camera, float* uright, float* depth);

/* Brute-force best / second-best scan with the reference's inner-loop semantics (SearchByBoW,
 * src/ORBmatcher.cc:477-507): train rows are visited in ascending index order; idx = lowest index of the
 * minimum distance (or -1 when no row is closer than 256), best = that distance, second = the smallest distance
 * among the other rows (256 when none). match[q] = idx when best <= th_low && (float)best < nnratio*(float)second,
 * else -1 (match may be NULL). Host buffers. */
orbx_status orbx_knn2(int device, const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt, int th_low,
                      float nnratio, int32_t* idx, uint16_t* best, uint16_t* second, int32_t* match);
/* Device-resident variant on `stream` (a cudaStream_t as void*, NULL = default stream); asynchronous. */
orbx_status orbx_knn2_device(const uint8_t* d_query, int64_t nq, const uint8_t* d_train, int64_t nt, int th_low,
                             float nnratio, int32_t* d_idx, uint16_t* d_best, uint16_t* d_second, int32_t* d_match,
                             void* stream);

/* Train-sharded kNN (BASELINE.json config 5): each rank scans its contiguous slice of the train set and emits one
 * packed 64-bit partial per query; the partials of all ranks are all-gathered (NCCL) rank-major into
 * d_gathered[rank*nq + q] and merged. The merge equals one ascending scan over the whole train set.
 * index_base = global index of the shard's first row. */
orbx_status orbx_knn2_partial_device(const uint8_t* d_query, int64_t nq, const uint8_t* d_train_shard, int64_t nt_shard,
                                     int64_t index_base, uint64_t* d_partial, void* stream);
orbx_status orbx_knn2_merge_device(const uint64_t* d_gathered, int ranks, int64_t nq, int th_low, float nnratio,
                                   int32_t* d_idx, uint16_t* d_best, uint16_t* d_second, int32_t* d_match, void* stream);
/* The three steps in one call for a C/C++ host that owns an NCCL communicator: scan this rank's shard, ncclAllGather the
 * 8-byte-per-query partials of all `nranks` ranks over `comm` (an ncclComm_t, passed as void*), merge. Everything is enqueued on `stream`
 * (a cudaStream_t as void*); results are identical on every rank and equal one ascending scan over the whole train set. The library does
 * not link NCCL: ncclAllGather is resolved at the first call from the NCCL already loaded in the process (the one `comm` was created with),
 * else from libnccl.so.2; ORBX_ERR_STATE when neither exists. d_work: device scratch of (nranks + orbx_knn2_work_parts(nq, nt_shard)) * nq
 * uint64, or NULL to let the call allocate and free it stream-ordered. */
orbx_status orbx_knn2_sharded(void* comm, int rank, int nranks, const uint8_t* d_query, int64_t nq, const uint8_t* d_train_shard,
                              int64_t nt_shard, int64_t index_base, int th_low, float nnratio, int32_t* d_idx, uint16_t* d_best,
                              uint16_t* d_second, int32_t* d_match, uint64_t* d_work, void* stream);
int orbx_knn2_work_parts(int64_t nq, int64_t nt_shard);

/* ---- rows "next" of the hot path: the steps either side of Extract and the N x N Hamming site ---- */

/* ConvertToGray — src/System.cc:122-137 (cv::cvtColor {RGB,BGR,RGBA,BGRA}2GRAY, OpenCV 8-bit fixed point). Host buffers;
 * channels = 3 or 4, rgb != 0 when the first channel is red (the reference's RGB_ flag). */
orbx_status orbx_convert_to_gray(int device, const uint8_t* src, int width, int height, size_t pitch, int channels, int rgb,
                                 uint8_t* dst, size_t dst_pitch);
/* ConvertToGray fused into Extract (src/System.cc:445-450, 506-509, 560-567): interleaved colour frames in, the gray image is
 * written straight into level 0 of the pyramid on the device. channels = 1 behaves like orbx_extract_batch. */
orbx_status orbx_extract_batch_color(orbx_handle h, const uint8_t* images, int frames, int width, int height, size_t pitch,
                                     size_t frame_stride, int channels, int rgb, orbx_keypoint* kps, uint8_t* desc, int cap, int* n);
/* The rectification step in front of TrackStereo — Examples/Stereo/stereo_euroc.cc:100-101: cv::remap(im, imRect, M1, M2, INTER_LINEAR)
 * with the CV_32F maps of cv::initUndistortRectifyMap (:88-89); OpenCV's 1/32-pixel fixed point, constant border 0. Host buffers;
 * map_pitch in bytes; dst is width x height, the maps' size. */
orbx_status orbx_remap(int device, const uint8_t* src, int src_width, int src_height, size_t src_pitch, const float* map1, const float* map2,
                       size_t map_pitch, uint8_t* dst, int width, int height, size_t dst_pitch);
/* The same fused into Extract: the maps are converted and uploaded once per extractor (one per camera, as the reference keeps M1l/M2l
 * and M1r/M2r), then raw frames go in and the rectified image is written straight into level 0 of the pyramid on the device.
 * GetImagePyramid()[0] of such an extract is the rectified image. */
orbx_status orbx_set_rectification(orbx_handle h, const float* map1, const float* map2, size_t map_pitch, int width, int height, int src_width,
                                   int src_height);
/* Device-resident remap of a batch with the handle's maps, asynchronous on the handle's stream: d_raw holds `frames` raw frames of the
 * size given to orbx_set_rectification; d_dst receives the rectified frames (dst_pitch a multiple of 4, >= width; the kernel writes
 * whole 4-byte words, so a row may be written up to 3 bytes past `width`). */
orbx_status orbx_rectify_batch_device(orbx_handle h, const uint8_t* d_raw, int frames, size_t pitch, size_t frame_stride, uint8_t* d_dst,
                                      size_t dst_pitch, size_t dst_stride);
orbx_status orbx_extract_batch_rectified(orbx_handle h, const uint8_t* images, int frames, int src_width, int src_height, size_t pitch,
                                         size_t frame_stride, orbx_keypoint* kps, uint8_t* desc, int cap, int* n);
/* UndistortKeyPoints — src/System.cc:153-174: cv::undistortPoints(pts, pts, K, distCoeffs, noArray(), K) on every keypoint position
 * (OpenCV's 5-round double-precision iteration); `dist` holds ndist (0..14) OpenCV distortion coefficients k1 k2 p1 p2 [k3 [k4 k5 k6
 * [s1..s4]]]. kps_un = kps when dist[0] == 0, as the reference does. Host buffers. */
orbx_status orbx_undistort_keypoints(int device, const orbx_keypoint* kps, int n, const orbx_camera* camera, const float* dist, int ndist,
                                     orbx_keypoint* kps_un);
/* Device-resident variant, asynchronous on `stream` (a cudaStream_t as void*, NULL = default stream); in place when d_kps_un == d_kps. */
orbx_status orbx_undistort_keypoints_device(const orbx_keypoint* d_kps, int n, const orbx_camera* camera, const float* dist, int ndist,
                                            orbx_keypoint* d_kps_un, void* stream);
/* ComputeStereoFromRGBD — src/System.cc:197-219: depth_map is width x height float32 (pitch in bytes). Host buffers. */
orbx_status orbx_stereo_from_rgbd(int device, const orbx_keypoint* kps, const orbx_keypoint* kps_un, int n, const float* depth_map,
                                  int width, int height, size_t pitch, const orbx_camera* camera, float* uright, float* depth);
/* The distance matrix + least-median selection of MapPoint::ComputeDistinctiveDescriptors — src/MapPoint.cc:286-314, for a
 * batch of descriptor sets: set s is rows [offsets[s], offsets[s+1]) of desc (32 bytes each). best[s] = index inside the set of
 * the descriptor with the least median distance to the set (first wins ties), -1 for an empty set. Host buffers. */
orbx_status orbx_distinctive_descriptors(int device, const uint8_t* desc, const int64_t* offsets, int nsets, int32_t* best);

/* ---- guided matchers: FeaturesGrid + the window searches of Tracking ---- */

typedef struct orbx_bounds { float minx, maxx, miny, maxy; } orbx_bounds;   /* ImageBounds, include/Frame.h:40-48 */
typedef struct orbx_pose { float R[9]; float t[3]; } orbx_pose;             /* CameraPose (include/CameraPose.h:32-90), R row-major */

/* A Frame as the matchers see it (include/Frame.h:83-168): keypointsUn, descriptors (n x 32), uright (NULL = every entry -1),
 * imageBounds, pyramid.nlevels (<= 16) and pyramid.scaleFactors. Host pointers. */
typedef struct orbx_frame_view
{
	int32_t n;
	const orbx_keypoint* kps_un;
	const uint8_t* desc;
	const float* uright;
	orbx_bounds bounds;
	int32_t nlevels;
	const float* scale_factors;
} orbx_frame_view;

/* One candidate map point of SearchByProjection(Frame&, mappoints, th): the track* members set by Tracking::SearchLocalPoints
 * (include/MapPoint.h:92-97). flags bit 0 = trackInView && !isBad(), bit 1 = Observations() > 0. */
typedef struct orbx_track_point { float proj_x, proj_y, proj_xr, view_cos; int32_t scale_level; int32_t flags; } orbx_track_point;
/* One keypoint of the last frame for SearchByProjection(currFrame, lastFrame, th, monocular): world position of its map point,
 * lastFrame.keypoints[i].octave, lastFrame.keypointsUn[i].angle. flags bit 0 = has a map point && !outlier, bit 1 = Observations() > 0. */
typedef struct orbx_last_point { float xw[3]; int32_t octave; float angle; int32_t flags; } orbx_last_point;

/* A frame resident on one GPU: keypoints, descriptors, right coordinates and its FeaturesGrid (FeaturesGrid::AssignFeatures,
 * src/Frame.cc:70-100, built on the device). Not re-entrant, like a Frame that is being matched. */
typedef struct orbx_frame_s* orbx_frame;
orbx_status orbx_frame_create(const orbx_frame_view* view, int device, orbx_frame* out);
/* Re-uses the device buffers of `f` for another frame (Tracking builds one Frame per image): uploads the view, rebuilds the grid. */
orbx_status orbx_frame_assign(orbx_frame f, const orbx_frame_view* view);
/* The same from device-resident data — e.g. straight from orbx_extract_batch_device's outputs (through orbx_undistort_keypoints_device
 * where the camera has distortion) and orbx_stereo_match_device's uright: Extract -> Frame -> SearchByProjection without the keypoints
 * and descriptors ever leaving the GPU. The frame copies what it is given (device to device); d_uright may be NULL (monocular). The
 * caller must have ordered the producing work before this call (it runs on the frame's own stream). Octaves must be < 16. */
orbx_status orbx_frame_assign_device(orbx_frame f, const orbx_keypoint* d_kps_un, const uint8_t* d_desc, const float* d_uright, int n,
                                     const orbx_bounds* bounds, int nlevels, const float* scale_factors);
orbx_status orbx_frame_destroy(orbx_frame f);
/* The grid as CSR: cell_start has 64*48+1 entries, cell = cx*48 + cy (grid_[cx][cy], include/Frame.h:72-79); items holds the keypoint
 * indices of every cell in push_back order. *n_items = number of keypoints inside the grid. */
orbx_status orbx_frame_grid(orbx_frame f, int32_t* cell_start, int32_t* items, int cap, int* n_items);
/* FeaturesGrid::GetFeaturesInArea (src/Frame.cc:102-145) for nq windows: xyr = (x, y, r) triples, levels = (minLevel, maxLevel) pairs.
 * offsets gets nq+1 entries; the indices of window q are indices[offsets[q] .. offsets[q+1]) in the reference's output order.
 * ORBX_ERR_CAPACITY when cap < offsets[nq] (offsets are valid then). */
orbx_status orbx_frame_features_in_area(orbx_frame f, const float* xyr, const int32_t* levels, int nq, int32_t* offsets, int32_t* indices,
                                        int cap);

/* frame_mp (f->n entries, in/out) stands for frame.mappoints: -1 = null, >= 0 = index into `pts` of the map point stored there,
 * -2 = some other map point with Observations() > 0, -3 = some other map point without observations. */

/* ORBmatcher::SearchByProjection(Frame&, const std::vector<MapPoint*>&, float th) — src/ORBmatcher.cc:315-382 (local-map tracking).
 * pt_desc = mappoint->GetDescriptor() of every point, 32 bytes each. nnratio = the matcher's fNNRatio_. */
orbx_status orbx_search_by_projection_local_map(orbx_frame f, int32_t* frame_mp, const orbx_track_point* pts, const uint8_t* pt_desc,
                                                int npts, float th, float nnratio, int* nmatches);
/* ORBmatcher::SearchByProjection(Frame& currFrame, const Frame& lastFrame, float th, bool monocular) — src/ORBmatcher.cc:1279-1362
 * (motion-model tracking) with CheckOrientation (:249-309) when check_orientation != 0. camera/cur_pose belong to currFrame. */
orbx_status orbx_search_by_projection_last_frame(orbx_frame cur, const orbx_camera* camera, const orbx_pose* cur_pose,
                                                 const orbx_pose* last_pose, int32_t* frame_mp, const orbx_last_point* pts,
                                                 const uint8_t* pt_desc, int npts, float th, int monocular, int check_orientation,
                                                 int* nmatches);
/* ORBmatcher::SearchForInitialization — src/ORBmatcher.cc:614-694. prev_matched: f1->n (x, y) pairs, in/out; matches12: f1->n, out. */
orbx_status orbx_search_for_initialization(orbx_frame f1, orbx_frame f2, float* prev_matched, int32_t* matches12, int window_size,
                                           float nnratio, int check_orientation, int* nmatches);
/* One map point of a key frame for the relocalisation search: world position, the members minDistance_ / maxDistance_ behind
 * Get{Min,Max}DistanceInvariance and PredictScale (src/MapPoint.cc:382-414), the angle of the key frame's keypoint (CheckOrientation).
 * flags bit 0 = map point present && !isBad() && not in alreadyFound. */
typedef struct orbx_keyframe_point { float xw[3]; float min_distance, max_distance; float angle; int32_t flags; } orbx_keyframe_point;
/* ORBmatcher::SearchByProjection(Frame& frame, KeyFrame* keyframe, const std::set<MapPoint*>& alreadyFound, float th, int ORBdist) —
 * src/ORBmatcher.cc:1364-1447 (relocalisation). camera / pose belong to `frame`; log_scale_factor = frame.pyramid.logScaleFactor
 * (src/System.cc:143). Any non-null frame.mappoints entry closes its keypoint (:1412-1413). The per-point geometry (projection,
 * cv::norm, PredictScale's log) is evaluated on the host in the reference's operation order; the window search runs on the device. */
orbx_status orbx_search_by_projection_keyframe(orbx_frame f, const orbx_camera* camera, const orbx_pose* pose, float log_scale_factor,
                                              int32_t* frame_mp, const orbx_keyframe_point* pts, const uint8_t* pt_desc, int npts, float th,
                                              int orb_dist, int check_orientation, int* nmatches);
/* ORBmatcher::SearchByProjection(const KeyFrame* keyframe, const Sim3& Scw, const std::vector<MapPoint*>& mappoints,
 * std::vector<MapPoint*>& matched, int th) — src/ORBmatcher.cc:518-612 (loop closing). `f` holds the key frame; matched (f->n entries,
 * in/out): -1 = null, -2 = a map point found before the call, >= 0 = index of the point stored by this call. flags bit 0 of a point =
 * !isBad() && not already in `matched` (:536-537). Host geometry as above (incl. the 60-degree viewing test), device window search. */
typedef struct orbx_sim3 { float R[9]; float t[3]; float s; } orbx_sim3;                    /* include/Sim3.h */
typedef struct orbx_sim3_point { float xw[3]; float normal[3]; float min_distance, max_distance; int32_t flags; } orbx_sim3_point;
orbx_status orbx_search_by_projection_sim3(orbx_frame f, const orbx_camera* camera, const orbx_sim3* Scw, float log_scale_factor, int32_t* matched,
                                          const orbx_sim3_point* pts, const uint8_t* pt_desc, int npts, int th, int* nmatches);
/* The window search underneath it, for callers that evaluate the geometry themselves (the C++ mirror does, with the reference's own
 * MapPoint::PredictScale and camera classes): window i is centred on (u, v) with half-size radius and admits the octaves
 * [min_level, max_level] with GetFeaturesInArea's level semantics; flags bit 0 = search this point; angle = the orientation compared
 * with the keypoint's by CheckOrientation. Loop semantics of src/ORBmatcher.cc:1410-1432: any non-null frame.mappoints entry closes
 * its keypoint, best distance wins (first on ties), accepted when best <= max_dist; every match closes its keypoint for later points. */
typedef struct orbx_window { float u, v, radius, angle; int32_t min_level, max_level; int32_t flags; } orbx_window;
orbx_status orbx_search_windows(orbx_frame f, int32_t* frame_mp, const orbx_window* windows, const uint8_t* pt_desc, int npts, int max_dist,
                                int check_orientation, int* nmatches);
/* DBoW2::FeatureVector (Thirdparty/DBoW2/DBoW2/FeatureVector.h: std::map<NodeId, std::vector<unsigned>>) flattened: node ids ascending,
 * the features of node k are indices[start[k] .. start[k+1]). */
typedef struct orbx_feature_vector { int32_t nnodes; const uint32_t* node_ids; const int32_t* start; const uint32_t* indices; } orbx_feature_vector;
/* ORBmatcher::SearchByBoW — src/ORBmatcher.cc:452-516 (KeyFrame vs Frame: valid2 == NULL) and :696-766 (KeyFrame vs KeyFrame), with
 * FeatureVectorIterator (:406-450) and CheckOrientation. f1 is the (first) key frame, f2 the frame / second key frame, both resident.
 * valid1 / valid2: per keypoint, map point present && !isBad(). match2 (f2->n entries, out) = the keypoint of f1 matched to keypoint
 * idx2 of f2, or -1: `matches[idx2] = mappoints1[match2[idx2]]` for the first variant, `matches12[match2[idx2]] = mappoints2[idx2]`
 * for the second. The feature vectors come from orbx_bow_transform below (or from DBoW2 on the host). */
orbx_status orbx_search_by_bow(orbx_frame f1, const orbx_feature_vector* fv1, const uint8_t* valid1, orbx_frame f2, const orbx_feature_vector* fv2,
                               const uint8_t* valid2, float nnratio, int check_orientation, int32_t* match2, int* nmatches);
/* ---- Matchers of local mapping and loop closing whose per-point search is independent of the other points. --------------------------
 * The window search underneath them: for every window the keypoint of `f` with the smallest descriptor distance among
 * GetFeaturesInArea(u, v, radius) with octave in [min_level, max_level] — the first one in GetFeaturesInArea's order on ties
 * (`dist < bestDist`) — and that distance; (-1, 256) when there is none. flags bit 0 = search this window; bit 1 = also apply the
 * chi-square gate of Fuse (src/ORBmatcher.cc:934-945: stereo keypoints against (u, v, ur) at 7.8, monocular ones against (u, v) at 5.99,
 * scaled by inv_sigma_sq[octave]). */
typedef struct orbx_best_window { float u, v, radius, ur; int32_t min_level, max_level; int32_t flags; } orbx_best_window;
orbx_status orbx_search_best_in_windows(orbx_frame f, const orbx_best_window* windows, const uint8_t* pt_desc, int npts, const float* inv_sigma_sq,
                                        int32_t* best_idx, int32_t* best_dist);
/* ORBmatcher::Fuse(KeyFrame* keyframe, const std::vector<MapPoint*>& mappoints, float th) — src/ORBmatcher.cc:868-980, the search half:
 * projection, image / distance / 60-degree gates, PredictScale, window, octave and chi-square gates, best distance (:879-954), host
 * geometry in the reference's operation order as for orbx_search_by_projection_sim3. flags bit 0 of a point = non-null && !isBad().
 * The map mutation (:956-976: Replace / AddObservation / AddMapPoint when best_dist <= TH_LOW) and the IsInKeyFrame test of :876 depend
 * on the points before and stay with the caller, who replays them in order over (best_idx, best_dist): the search itself reads nothing
 * they change. inv_sigma_sq = keyframe->pyramid.invSigmaSq. */
orbx_status orbx_fuse(orbx_frame f, const orbx_camera* camera, const orbx_pose* pose, float log_scale_factor, const float* inv_sigma_sq,
                      const orbx_sim3_point* pts, const uint8_t* pt_desc, int npts, float th, int32_t* best_idx, int32_t* best_dist);
/* ORBmatcher::Fuse(KeyFrame*, const Sim3& Scw, mappoints, th, replacePoints) — src/ORBmatcher.cc:982-1088, the search half (:987-1067).
 * flags bit 0 = !isBad() && not in keyframe->GetMapPoints() (:1002-1003). The caller applies :1069-1084 over (best_idx, best_dist <= TH_LOW). */
orbx_status orbx_fuse_sim3(orbx_frame f, const orbx_camera* camera, const orbx_sim3* Scw, float log_scale_factor, const orbx_sim3_point* pts,
                           const uint8_t* pt_desc, int npts, float th, int32_t* best_idx, int32_t* best_dist);
/* ORBmatcher::SearchBySim3(kf1, kf2, matches12, S12, th) — src/ORBmatcher.cc:1090-1277. pts1 / pts2: one entry per keypoint of f1 / f2
 * (GetMapPointMatches), flags bit 0 = map point present && !alreadyMatched && !isBad() (:1130, :1197; `angle` unused); desc1 / desc2 the
 * map points' descriptors. Outputs: match1 / match2 (optional) = the two directed searches (:1127-1258), matches12[i1] = the keypoint of
 * f2 on which both agree or -1 (:1260-1274; the caller stores mappoints2[idx2]), *nfound. */
orbx_status orbx_search_by_sim3(orbx_frame f1, const orbx_camera* camera1, const orbx_pose* pose1, float log_scale_factor1, orbx_frame f2,
                                const orbx_camera* camera2, const orbx_pose* pose2, float log_scale_factor2, const orbx_sim3* S12, float th,
                                const orbx_keyframe_point* pts1, const uint8_t* desc1, const orbx_keyframe_point* pts2, const uint8_t* desc2,
                                int32_t* match1, int32_t* match2, int32_t* matches12, int* nfound);
/* ORBmatcher::SearchForTriangulation(kf1, kf2, F12, matchIds, onlyStereo) — src/ORBmatcher.cc:768-866 with CheckDistEpipolarLine
 * (:384-404), FeatureVectorIterator and CheckOrientation. has_mp1 / has_mp2: per keypoint, GetMapPoint(idx) != NULL. F12 row-major 3x3,
 * epipole2 = proj2.WorldToImage(keyframe1->GetCameraCenter()) (:772-773), sigma_sq2 = keyframe2->pyramid.sigmaSq. matches12 (f1->n
 * entries, out) = the keypoint of f2 matched to keypoint idx1 or -1; matchIds of the reference = the pairs (idx1, matches12[idx1] >= 0)
 * in ascending idx1. This fork never sets matched2 (:780, :814), so two keypoints of f1 may share a keypoint of f2 — reproduced. */
orbx_status orbx_search_for_triangulation(orbx_frame f1, const orbx_feature_vector* fv1, const uint8_t* has_mp1, orbx_frame f2,
                                          const orbx_feature_vector* fv2, const uint8_t* has_mp2, const float* F12, const float* epipole2,
                                          const float* sigma_sq2, int only_stereo, int check_orientation, int32_t* matches12, int* nmatches);

/* Diagnostics of the last search on `f`: rounds needed to reach the sequential result (>= 1), the kernel's duration (CUDA events) and
 * the microseconds its phases took (enumeration, -, -, distances, rounds, finalisation, and the part of `rounds` spent staging state into
 * shared memory; %globaltimer). Any pointer may be NULL; phase_us needs room for 7 floats. */
orbx_status orbx_frame_last_stats(orbx_frame f, int* rounds, float* kernel_ms, float* phase_us);

/* ---- Bag-of-words transform: Frame::ComputeBoW / KeyFrame::ComputeBoW — src/Frame.cc:208-214, src/KeyFrame.cc:66-74,
 * i.e. ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h). ---------------- */
typedef struct orbx_vocabulary_s* orbx_vocabulary;
/* The tree as loadFromTextFile reads it: entry i describes node id i + 1 (node 0 is the root), in file order — a parent precedes its
 * children, Node::children is filled in that order and word ids are handed out in that order to the entries with is_leaf != 0.
 * scoring / weighting = DBoW2::ScoringType / WeightingType (BowVector.h:36-55); the reference's vocabulary is k 10, L 6, L1_NORM, TF_IDF. */
typedef struct orbx_vocabulary_desc
{
	int32_t k, L, scoring, weighting;
	int64_t nnodes;                  /* without the root */
	const int32_t* parent;           /* [nnodes] node id of the parent (0 = root) */
	const uint8_t* is_leaf;          /* [nnodes] */
	const uint8_t* descriptors;      /* [nnodes][32] */
	const double* weights;           /* [nnodes] Node::weight */
} orbx_vocabulary_desc;
orbx_status orbx_vocabulary_create(const orbx_vocabulary_desc* desc, int device, orbx_vocabulary* out);
/* TemplatedVocabulary<FORB::TDescriptor, FORB>::loadFromTextFile — Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.cpp:21-90, with this fork's
 * tokenizer (strtok + atoi, weight read as int, :67). A file the reference rejects (:37) returns ORBX_ERR_INVALID. */
orbx_status orbx_vocabulary_load_text(const char* path, int device, orbx_vocabulary* out);
orbx_status orbx_vocabulary_info(orbx_vocabulary v, int* k, int* L, int* scoring, int* weighting, int64_t* nodes, int64_t* words);
orbx_status orbx_vocabulary_destroy(orbx_vocabulary v);
/* TemplatedVocabulary::transform(features, BowVector& v, FeatureVector& fv, levelsup) — TemplatedVocabulary.h:1129-1197 (per feature
 * :1220-1262, FORB::distance FORB.cpp:79-98, BowVector.cpp:32-87, FeatureVector.cpp:30-44). desc = n x 32 descriptor rows (what
 * Converter::toDescriptorVector splits). Outputs, each with room for n entries (fv_start: n + 1): the BowVector as (word_ids ascending,
 * word_vals) and the FeatureVector in the orbx_feature_vector layout. feat_word / feat_node (optional, n entries): the word and the
 * levelsup-ancestor of every feature, stopped or not. A leaf above level L - levelsup leaves the node id unset in the reference
 * (uninitialised NodeId, :1156); it is reported as node 0 here. n <= 16384. */
orbx_status orbx_bow_transform(orbx_vocabulary v, const uint8_t* desc, int n, int levelsup, int32_t* word_ids, double* word_vals, int32_t* n_words,
                               uint32_t* fv_nodes, int32_t* fv_start, uint32_t* fv_items, int32_t* n_fv_nodes, int32_t* feat_word, int32_t* feat_node);
/* The same for a batch of frames resident on the vocabulary's device, e.g. the outputs of orbx_extract_batch_device: d_desc
 * [frames][cap][32], d_n [frames]. Outputs [frames][cap] (d_fv_start [frames][cap + 1], d_counts [frames][2] = n_words, n_fv_nodes;
 * d_feat_word / d_feat_node optional). Enqueued on `stream` (a cudaStream_t; NULL = the vocabulary's own stream), no copies, no sync. */
orbx_status orbx_bow_transform_batch_device(orbx_vocabulary v, const uint8_t* d_desc, const int32_t* d_n, int frames, int cap, int levelsup,
                                            int32_t* d_word_ids, double* d_word_vals, uint32_t* d_fv_nodes, int32_t* d_fv_start, uint32_t* d_fv_items,
                                            int32_t* d_counts, int32_t* d_feat_word, int32_t* d_feat_node, void* stream);
/* L1Scoring::score(v1, v2) — Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:24-58, what ORBVocabulary::score evaluates for the reference's
 * vocabulary (src/KeyFrameDatabase.cc:117, 216; src/LoopClosing.cc:176) — for npairs pairs of BowVectors stored back to back: vector i is
 * (ids, vals)[offsets[i] .. offsets[i+1]), ids ascending. */
orbx_status orbx_bow_score_l1(orbx_vocabulary v, const int32_t* ids, const double* vals, const int32_t* offsets, const int32_t* pair_a,
                              const int32_t* pair_b, int npairs, double* scores);

/* Integer-pipe microbenchmark used as the roofline denominator of the matcher: sustained POPC.32 per second on
 * `device` (all SMs, register operands). */
orbx_status orbx_measure_popc_peak(int device, double* popc_per_second);

#ifdef __cplusplus
}
#endif

#endif /* ORBX_H */
