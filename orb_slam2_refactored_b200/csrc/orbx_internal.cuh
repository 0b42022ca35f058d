// Internal declarations shared by the sm_100a kernels and the C-ABI host layer (not installed).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orbx.h"

#define ORBX_MAX_LEVELS 12
#define ORBX_MAX_ROOTS 16
#define ORBX_BORDER 16          // EDGE_THRESHOLD - 3, src/ORBextractor.cc:755
#define ORBX_EDGE 19            // EDGE_THRESHOLD, :70
#define ORBX_CELL 30            // CELL_SIZE, :491
#define ORBX_HALF_PATCH 15      // HALF_PATCH_SIZE, :69
#define ORBX_PATCH 31           // PATCH_SIZE, :68

// Packed FAST candidate: x[0:12) | y[12:24) | response[24:32). Level coordinates; limits w,h <= 4096.
__host__ __device__ inline uint32_t orbx_pack(int x, int y, int r) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)r << 24); }
__host__ __device__ inline int orbx_px(uint32_t v) { return (int)(v & 0xfffu); }
__host__ __device__ inline int orbx_py(uint32_t v) { return (int)((v >> 12) & 0xfffu); }
__host__ __device__ inline int orbx_pr(uint32_t v) { return (int)(v >> 24); }

// One pyramid level of the plan. All offsets are per frame.
struct OrbxLevel
{
	int w, h, pitch;             // pixels, bytes
	int64_t offset;              // byte offset of the level inside a frame slab
	// FAST cell grid (src/ORBextractor.cc:489-540)
	int minx, miny, maxx, maxy;  // roi
	int ncx, ncy, cellw, cellh;  // cells actually visited, nominal cell size
	int cell_base;               // first cell (of the frame's cell array) belonging to this level
	int cell_cap;                // candidate slots per cell: ceil(cellw/2)*ceil(cellh/2), exact worst case of strict NMS
	int cand_base, cand_cap;     // slot range of the level in the frame's candidate arrays
	// quadtree (:542-693)
	int quota, n_roots;
	int root_base;               // offset of the n_roots+1 strip bounds in d_root_x
	int rootlut_base;            // offset of the per-x root id table in d_root_lut
	int sel_base, sel_cap;       // range of the level in the frame's selected-keypoint array
	// resize tables of this level as destination (level >= 1)
	int xtab_base, ytab_base;
	float scale;                 // scaleFactors_[s]
	int py_smem;                 // dynamic shared memory of the cp.async resize kernel producing this level
	int py_bw[2], py_bh[2];      // TMA box (bytes x rows of level s - 1) of the strip resize kernel producing this level, [0] throughput tiles, [1] the
	                             // 8-row tiles of small batches; 0: not usable (scale factor too large for one box)
};

struct OrbxPlanDev
{
	int nlevels, frames;
	int frame0;                  // index of this launch's first frame inside the planned batch (tensor maps address the whole batch)
	int ini_th, min_th;
	int cells_per_frame, cand_per_frame, sel_per_frame;
	int node_cap;                // max list length of the quadtree over all levels (+ margin)
	int out_cap;                 // keypoint slots per frame in the output arrays
	OrbxLevel lv[ORBX_MAX_LEVELS];
	// level 0 may alias the caller's device buffer
	const uint8_t* l0; int64_t l0_pitch, l0_stride;
	uint8_t* pyr; uint8_t* blur; int64_t slab;         // frame slabs (levels >= 1 of pyr; all levels of blur)
	uint8_t* fmap_ini; uint8_t* fmap_min;   // FAST bound bitmaps, [frames][slab / 8]: bit x of row y of level s at (offset_s + y * pitch_s) / 8 + x / 8
	uint32_t* cand;              // [frames][cand_per_frame] per-cell slots
	int* cell_count;             // [frames][cells_per_frame]
	const int4* cell_tab;        // [cells_per_frame] x0 | y0 << 16, view w | h << 16, level, cell index inside the level
	uint32_t* qbuf0; uint32_t* qbuf1;   // [frames][cand_per_frame] quadtree ping-pong segments
	int* cand_count;             // [frames][nlevels] DetectFAST totals (debug/probes)
	uint32_t* sel;               // [frames][sel_per_frame] selected keypoints, list order
	int* sel_count;              // [frames][nlevels]
	const int* root_x; const uint8_t* root_lut;
	const int* xofs; const short2* xcoef; const int* yofs; const short2* ycoef;   // resize tables
	int* pyr_done;               // [frames of the plan][ORBX_MAX_LEVELS] tiles of a level finished in the current one-launch ComputePyramid
	uint32_t* ovf_list; int* ovf_count;   // cells of this launch whose flagged pixels did not fit the cell kernel's list (frame * cells_per_frame + cell)
};

__device__ __forceinline__ const uint8_t* orbx_level_ptr(const OrbxPlanDev& P, int frame, int level)
{
	if (level == 0)
		return P.l0 + (int64_t)frame * P.l0_stride;
	return P.pyr + (int64_t)frame * P.slab + P.lv[level].offset;
}
__device__ __forceinline__ int64_t orbx_level_pitch(const OrbxPlanDev& P, int level)
{
	return level == 0 ? P.l0_pitch : (int64_t)P.lv[level].pitch;
}

// TMA descriptors of the pyramid levels as seen by the FAST kernel: 3-D u8 tensors (x, y, frame), box FT_TS x box_h x 1
struct OrbxTmaMaps
{
	CUtensorMap level[ORBX_MAX_LEVELS];
	int box_h[ORBX_MAX_LEVELS];
};
int orbx_fast_tile_stride();
int orbx_fast_tile_rows();
// TMA descriptors of the strip kernels: the levels as (pitch, h, frames) u8 tensors with the strip box (blur + dense FAST bound), and
// level s - 1 with the source box of the resize tile that produces level s
struct OrbxStripMaps { CUtensorMap level[ORBX_MAX_LEVELS]; };
struct OrbxPyrMaps { CUtensorMap src[ORBX_MAX_LEVELS]; };
// Tile rows of the strip kernels. which = 0: throughput tiles (ORBX_STRIP_TH / ORBX_PYR_TH, default 32); which = 2: the blur's throughput tiles
// (default 64); which = 1: the 8-row tiles used when
// a launch covers at most ORBX_SMALL_BATCH frames (Tracking extracts one frame at a time: more, shorter warps cut the launch's latency)
#define ORBX_SMALL_BATCH 16
int orbx_strip_rows(int which);
int orbx_strip_box_w();
int orbx_pyramid_strip_rows(int which);
cudaError_t orbx_kernels_init();  // per-device function attributes (dynamic shared memory limits); called once per orbx_create

// kernel launchers (orbx_extract.cu)
void orbx_launch_gray(const uint8_t* src, int64_t spitch, int64_t sstride, int channels, int rgb, uint8_t* dst, int64_t dpitch, int64_t dstride,
                      int w, int h, int frames, cudaStream_t st);
void orbx_launch_remap(const uint8_t* src, int64_t spitch, int64_t sstride, int sw, int sh, const int2* tab, uint8_t* dst, int64_t dpitch,
                       int64_t dstride, int w, int h, int frames, cudaStream_t st);   // tab[y*w + x] = (ix & 0xffff | iy << 16, fx | fy << 5)
void orbx_launch_pyramid(const OrbxPlanDev& P, const OrbxPyrMaps pmaps[2], int level, cudaStream_t st);
// levels [s0, s1) of every frame, one launch where the plan allows; the launchers below take level ranges too (s1 <= 0: up to the last level)
cudaError_t orbx_launch_pyramid_all(const OrbxPlanDev& P, const OrbxPyrMaps pmaps[2], cudaStream_t st, int s0 = 1, int s1 = 0);
void orbx_launch_fast(const OrbxPlanDev& P, const OrbxTmaMaps& maps, const OrbxStripMaps smaps[3], cudaStream_t st, int part = 0, int s0 = 0, int s1 = 0);   // part: 0 both kernels, 1 dense bound, 2 cells
void orbx_launch_quadtree(const OrbxPlanDev& P, int* cell_off, cudaStream_t st, int s0 = 0, int s1 = 0);   // cell_off: scratch, [frames][cells_per_frame]
void orbx_launch_blur(const OrbxPlanDev& P, const OrbxStripMaps smaps[3], cudaStream_t st);
void orbx_launch_describe(const OrbxPlanDev& P, orbx_keypoint* d_kps, uint8_t* d_desc, int32_t* d_n, cudaStream_t st);
void orbx_launch_stamp(unsigned long long* slot, cudaStream_t st);
void orbx_launch_debug_cos_sin(uint32_t first_bits, int64_t n, float* d_cos, float* d_sin, cudaStream_t st);
size_t orbx_quadtree_smem(int node_cap, bool big);
int orbx_pyramid_tile_rows();
int orbx_pyramid_max_src_rows();
int orbx_pyramid_tile_cols();
int orbx_pyramid_max_src_bytes();
cudaError_t orbx_upload_pattern();

// matcher launchers (orbx_match.cu)
void orbx_launch_hamming_pairs(const uint8_t* a, const uint8_t* b, int64_t n, int32_t* out, cudaStream_t st);
void orbx_launch_knn2_partial(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int64_t base, uint64_t* partial,
                              cudaStream_t st);
void orbx_launch_knn2_fold(const uint64_t* parts, int nparts, int64_t nq, uint64_t* packed, cudaStream_t st);
void orbx_launch_knn2_merge(const uint64_t* gathered, int ranks, int64_t nq, int th_low, float nnratio, int32_t* idx,
                            uint16_t* best, uint16_t* second, int32_t* match, cudaStream_t st);

struct OrbxStereoArgs
{
	int frames, cap, nlevels;
	const orbx_keypoint* kl; const uint8_t* dl; const int32_t* nl;   // [frames][cap]
	const orbx_keypoint* kr; const uint8_t* dr; const int32_t* nr;
	// pyramids: level pointers per frame are base + frame*stride + offset[level]
	const uint8_t* pl0; int64_t pl0_pitch, pl0_stride; const uint8_t* pl; int64_t pl_slab;
	const uint8_t* pr0; int64_t pr0_pitch, pr0_stride; const uint8_t* pr; int64_t pr_slab;
	int lw[ORBX_MAX_LEVELS], lh[ORBX_MAX_LEVELS], lpitch[ORBX_MAX_LEVELS]; int64_t loff[ORBX_MAX_LEVELS];
	float scale[ORBX_MAX_LEVELS], inv_scale[ORBX_MAX_LEVELS];
	float bf, baseline;
	float* uright; float* depth;      // [frames][cap]
	int* sad;                         // [frames][cap] scratch: SAD of kept matches, -1 otherwise
	// rowIndices of src/ORBmatcher.cc:84-100 as CSR per frame: row_start[frames][rows + 1], row_items[frames][items_cap] = (x bits, iR | octave << 16)
	int rows, items_cap;
	int* row_start; uint2* row_items;
};
int orbx_stereo_items_per_keypoint(float max_scale);   // upper bound of the rows one right keypoint is listed in
void orbx_launch_stereo(const OrbxStereoArgs& A, cudaStream_t st);
void orbx_launch_stereo_from_rgbd(const orbx_keypoint* kps, const orbx_keypoint* kps_un, int n, const uint8_t* depth_map, int64_t pitch, float bf,
                                  float* uright, float* depth, cudaStream_t st);
void orbx_launch_undistort(const orbx_keypoint* src, orbx_keypoint* dst, int n, const float* cam4, const float* dist, int ndist, cudaStream_t st);
void orbx_launch_distinctive(const uint8_t* desc, const int64_t* offsets, int nsets, int32_t* best, cudaStream_t st);
double orbx_popc_probe(int device);

// shared by the host layers (orbx_api.cu owns the thread-local error text)
orbx_status orbx_fail(orbx_status s, const char* msg);
bool orbx_device_usable(int device, const char** why);
