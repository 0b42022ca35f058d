// Guided matchers on sm_100a: FeaturesGrid and the window searches of Tracking, bit-exact against the reference's
// sequential loops.
//
//   FeaturesGrid::AssignFeatures / GetFeaturesInArea            src/Frame.cc:70-145            k_grid_build, k_features_in_area
//   ORBmatcher::SearchByProjection(Frame&, mappoints, th)       src/ORBmatcher.cc:315-382      k_guided_search, mode LOCAL_MAP
//   ORBmatcher::SearchByProjection(currFrame, lastFrame, ...)   src/ORBmatcher.cc:1279-1362    k_guided_search, mode LAST_FRAME
//   ORBmatcher::SearchForInitialization                         src/ORBmatcher.cc:614-694      k_guided_search, mode INITIALIZATION
//   CheckOrientation                                            src/ORBmatcher.cc:249-309      inside k_guided_search
//
// The reference loops are sequential: a map point skips keypoints that an EARLIER map point (with observations) has taken, and
// SearchForInitialization skips a keypoint that an earlier one matched at a smaller-or-equal distance. Both are restated as a
// fixpoint over rounds (oracle/guided_oracle.cc carries the argument): every point recomputes its choice against the claims of the
// lower-numbered points of the previous round until nothing changes; the unique fixpoint is the sequential result. Everything that
// does not depend on that state — window enumeration in the reference's output order, the uright gate, all Hamming distances — is done
// once, in parallel, into a per-point candidate list; a round only re-scans those short lists.
//
// One CTA per search (a frame holds 1-8 k keypoints: the work is latency-, not bandwidth-bound); phases are separated by block barriers.
#include "orbx_internal.cuh"

#include <cooperative_groups.h>
#include <math.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "orbx_sort.cuh"

namespace {

constexpr int GRID_COLS = 64, GRID_ROWS = 48, GRID_CELLS = GRID_COLS * GRID_ROWS;   // include/Frame.h:72-73
constexpr int G_THREADS = 1024;
constexpr int G_SMEM_INTS = 200 * 1024 / 4;   // dynamic shared memory of the search kernel, in ints
constexpr int G_CLUSTER = 8;          // CTAs (= SMs) of one search: a thread-block cluster, the portable maximum
constexpr int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;                        // src/ORBmatcher.cc:41-43
constexpr int MODE_LOCAL_MAP = 0, MODE_LAST_FRAME = 1, MODE_INIT = 2, MODE_BOW = 3, MODE_WINDOWS = 4;

// a search window computed on the host (MODE_WINDOWS): the geometry of the variants whose projection involves log / sqrt stays in
// host arithmetic, identical to the reference's; the device does the search
struct GuidedWindow { float u, v, radius, angle; int levels; int flags; };   // levels = minLevel | maxLevel << 16; flags bit0 = active

// candidate entry: keypoint index [0:19) | octave [19:23) | distance [23:32) (511 = removed by the uright gate)
constexpr int E_IDX_BITS = 19, E_LVL_BITS = 4;
constexpr uint32_t E_IDX_MASK = (1u << E_IDX_BITS) - 1, E_SKIP = 511;

struct GridDev
{
	int n, nlevels;
	orbx_bounds b;
	float invW, invH;
	const int* cell_start;   // [GRID_CELLS + 1]
	const int4* rec;         // per grid item, cell-major: x bits, y bits, keypoint index, octave
};

struct GuidedArgs
{
	GridDev G;
	const orbx_keypoint* kps2;   // keypointsUn of the searched frame
	const uint8_t* desc2;
	const float* uright2;
	float sf[16];
	int mode, npts;
	const orbx_track_point* tp;
	const orbx_last_point* lp;
	const orbx_keypoint* kps1;   // INIT: frame 1
	float* prev;                 // INIT: prevMatched, out
	const uint8_t* pt_desc;
	float th, nnratio, radius;
	float fx, fy, cx, cy, bf;
	float R[9], t[3];
	int forward, backward, check_ori;
	// scratch, per point
	float* pu; float* pv; float* pur; float* prad;
	int* plevels;                // minLevel | maxLevel << 16 (as int16), -32768 in the low half = inactive
	int* off;                    // start of the point's candidate list
	int* len;                    // its length
	int* cursor;                 // [2] list allocation cursors, used alternately (the idle one is zeroed for the next launch)
	int parity, lanes;           // cursor in use; threads per point in the enumeration (power of two <= 32)
	int* choice;                 // keypoint taken by the point
	int* aux0; int* aux1;        // INIT: best distance of the accepted match, ping-pong; LAST_FRAME: histogram bin of the match
	int* next;                   // INIT: claim lists
	// scratch, per keypoint of the searched frame
	int* owner;
	// candidate lists
	uint32_t* list; int* entry_pt; int cap;
	const int32_t* mp_in;        // frame.mappoints on entry (LOCAL_MAP, LAST_FRAME)
	const float* prev_in;        // INIT: prevMatched on entry
	int32_t* frame_mp;           // out: frame.mappoints (LOCAL_MAP, LAST_FRAME); INIT: matches12 [npts]
	int* result;                 // [0] nmatches, [1] entries, [2] rounds, [3] overflow
	const GuidedWindow* win;     // WINDOWS
	int max_dist;                // WINDOWS: accept best <= max_dist (ORBdist, :1424)
	// BOW: point i = (keypoint of frame 1, first position and count of its vocabulary node's features in bow_idx2)
	const int* bow_pt; const uint32_t* bow_idx2; const uint8_t* bow_valid2;
	int bow_strict, ori_swap;    // best < TH_LOW instead of <= (:750 vs :501); CheckOrientation's argument order (:763 vs :512)
	unsigned long long* stamps;  // [8] %globaltimer at the phase boundaries (diagnostics)
	int use_smem;                // cell offsets, owner, its initial value and choice in dynamic shared memory
	int smem_ints;               // dynamic shared memory of this launch, in ints
};

__device__ __forceinline__ int hamming256(const uint8_t* a, const uint8_t* b)
{
	const uint4 a0 = *reinterpret_cast<const uint4*>(a), a1 = *reinterpret_cast<const uint4*>(a + 16);
	const uint4 b0 = *reinterpret_cast<const uint4*>(b), b1 = *reinterpret_cast<const uint4*>(b + 16);
	return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
	       __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// GetFeaturesInArea (src/Frame.cc:102-145): f(index, octave) for every hit, in the reference's output order. A column of cells
// (fixed cx, cy ascending) is one contiguous run of the cell-major record array.
template <class F>
__device__ __forceinline__ void for_window(const GridDev& G, float x, float y, float r, int minLevel, int maxLevel, F f)
{
	const int mincx = max(__float2int_rd(G.invW * (x - r - G.b.minx)), 0);
	const int maxcx = min(__float2int_ru(G.invW * (x + r - G.b.minx)), GRID_COLS - 1);
	const int mincy = max(__float2int_rd(G.invH * (y - r - G.b.miny)), 0);
	const int maxcy = min(__float2int_ru(G.invH * (y + r - G.b.miny)), GRID_ROWS - 1);
	if (mincx >= GRID_COLS || maxcx < 0 || mincy >= GRID_ROWS || maxcy < 0) return;
	const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
	if (maxLevel < 0) maxLevel = G.nlevels;
	for (int cx = mincx; cx <= maxcx; cx++)
	{
		const int a = G.cell_start[cx * GRID_ROWS + mincy], b = G.cell_start[cx * GRID_ROWS + maxcy + 1];
		for (int p = a; p < b; p++)
		{
			const int4 rec = G.rec[p];
			if (checkLevels && (rec.w < minLevel || rec.w > maxLevel)) continue;
			if (fabsf(__int_as_float(rec.x) - x) < r && fabsf(__int_as_float(rec.y) - y) < r) f(rec.z, rec.w);
		}
	}
}

// The same walk restricted to the cell columns [c0, c1) of an already clipped window (mincy..maxcy); used when several lanes share a window.
template <class F>
__device__ __forceinline__ void for_columns(const GridDev& G, float x, float y, float r, int minLevel, int maxLevel, int c0, int c1, int mincy,
                                            int maxcy, F f)
{
	const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
	if (maxLevel < 0) maxLevel = G.nlevels;
	for (int cx = c0; cx < c1; cx++)
	{
		const int a = G.cell_start[cx * GRID_ROWS + mincy], b = G.cell_start[cx * GRID_ROWS + maxcy + 1];
		for (int p = a; p < b; p++)
		{
			const int4 rec = G.rec[p];
			if (checkLevels && (rec.w < minLevel || rec.w > maxLevel)) continue;
			if (fabsf(__int_as_float(rec.x) - x) < r && fabsf(__int_as_float(rec.y) - y) < r) f(rec.z, rec.w);
		}
	}
}

// exclusive scan of a[0..m) in place by the whole CTA; returns the total. s_w: G_THREADS/32 + 1 ints of shared memory.
__device__ int block_exscan_inplace(int* a, int m, int* s_w)
{
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	int base = 0;
	for (int i0 = 0; i0 < m; i0 += G_THREADS)
	{
		const int i = i0 + tid;
		const int v = i < m ? a[i] : 0;
		int inc = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const int o = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += o;
		}
		if (lane == 31) s_w[warp] = inc;
		__syncthreads();
		int wbase = 0, tot = 0;
		for (int w = 0; w < G_THREADS / 32; w++)
		{
			const int t = s_w[w];
			if (w < warp) wbase += t;
			tot += t;
		}
		if (i < m) a[i] = base + wbase + inc - v;
		base += tot;
		__syncthreads();
	}
	return base;
}

// ---------------------------------------------------------------------------------------------------------------------------
// FeaturesGrid::AssignFeatures (src/Frame.cc:70-100): counting sort of the keypoints into grid_[cx][cy], push_back order kept
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(G_THREADS) k_grid_build(const orbx_keypoint* __restrict__ kps, int n, orbx_bounds b, float invW, float invH,
                                                          int* __restrict__ cell_start, int4* __restrict__ rec, int* __restrict__ cell_of)
{
	__shared__ int s_cnt[GRID_CELLS];
	__shared__ int s_beg[GRID_CELLS];
	__shared__ int s_w[G_THREADS / 32 + 1];
	const int tid = threadIdx.x;
	for (int c = tid; c < GRID_CELLS; c += G_THREADS) s_cnt[c] = 0;
	__syncthreads();
	for (int i = tid; i < n; i += G_THREADS)
	{
		const int cx = (int)roundf(invW * (kps[i].x - b.minx));   // Round() = std::round, :32, :91-92
		const int cy = (int)roundf(invH * (kps[i].y - b.miny));
		const int cell = (cx < 0 || cx >= GRID_COLS || cy < 0 || cy >= GRID_ROWS) ? -1 : cx * GRID_ROWS + cy;   // :95-96
		cell_of[i] = cell;
		if (cell >= 0) atomicAdd(&s_cnt[cell], 1);
	}
	__syncthreads();
	for (int c = tid; c < GRID_CELLS; c += G_THREADS) s_beg[c] = s_cnt[c];
	__syncthreads();
	const int total = block_exscan_inplace(s_beg, GRID_CELLS, s_w);
	for (int c = tid; c < GRID_CELLS; c += G_THREADS) { cell_start[c] = s_beg[c]; s_cnt[c] = s_beg[c]; }
	if (tid == 0) cell_start[GRID_CELLS] = total;
	__syncthreads();
	for (int i = tid; i < n; i += G_THREADS)
	{
		const int cell = cell_of[i];
		if (cell < 0) continue;
		const int p = atomicAdd(&s_cnt[cell], 1);
		rec[p] = make_int4(__float_as_int(kps[i].x), __float_as_int(kps[i].y), i, kps[i].octave);
	}
	__syncthreads();
	// the atomics filled each cell in arbitrary order; a cell holds a handful of keypoints: sort it by index (= push_back order)
	for (int c = tid; c < GRID_CELLS; c += G_THREADS)
	{
		const int a = s_beg[c], e = s_cnt[c];
		for (int i = a + 1; i < e; i++)
		{
			const int4 v = rec[i];
			int j = i - 1;
			while (j >= a && rec[j].z > v.z) { rec[j + 1] = rec[j]; --j; }
			rec[j + 1] = v;
		}
	}
}

// GetFeaturesInArea for a batch of windows: count, scan, fill
__global__ void __launch_bounds__(G_THREADS) k_features_in_area(GridDev G, const float* __restrict__ xyr, const int* __restrict__ levels, int nq,
                                                               int* __restrict__ offsets, int* __restrict__ indices, int cap)
{
	__shared__ int s_w[G_THREADS / 32 + 1];
	const int tid = threadIdx.x;
	for (int q = tid; q < nq; q += G_THREADS)
	{
		int c = 0;
		for_window(G, xyr[3 * q], xyr[3 * q + 1], xyr[3 * q + 2], levels[2 * q], levels[2 * q + 1], [&](int, int) { c++; });
		offsets[q] = c;
	}
	__syncthreads();
	const int total = block_exscan_inplace(offsets, nq, s_w);
	if (tid == 0) offsets[nq] = total;
	if (total > cap) return;
	for (int q = tid; q < nq; q += G_THREADS)
	{
		int p = offsets[q];
		for_window(G, xyr[3 * q], xyr[3 * q + 1], xyr[3 * q + 2], levels[2 * q], levels[2 * q + 1], [&](int idx, int) { indices[p++] = idx; });
	}
}

// ---------------------------------------------------------------------------------------------------------------------------
// The three window searches
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pack_levels(int lo, int hi) { return (lo & 0xffff) | (hi << 16); }
__device__ __forceinline__ int level_lo(int v) { return (int)(short)(v & 0xffff); }
__device__ __forceinline__ int level_hi(int v) { return v >> 16; }
constexpr int INACTIVE = -32768;

// CheckOrientation (:249-309) over the matches (i, choice[i]); bins were stored by the caller. Invalidates through `erase(i)`,
// returns matches - reduction. s_hist: HISTO_LENGTH ints, zeroed by the caller before the bins were counted; s_misc: 2 ints.
template <class Erase>
__device__ int check_orientation(const GuidedArgs& A, const int* choice, const int* bin_of, int* s_hist, uint64_t* s_items, int* s_misc, Erase erase)
{
	const int tid = threadIdx.x;
	if (tid == 0)
	{
		int total = 0;
		for (int b = 0; b < HISTO_LENGTH; b++)
		{
			s_items[b] = ((uint64_t)(uint32_t)s_hist[b] << 32) | (uint32_t)b;
			total += s_hist[b];
		}
		qs_sort_serial(s_items, HISTO_LENGTH);   // std::sort by size, descending: unstable, so replayed (orbx_sort.cuh)
		const double max1 = (double)(uint32_t)(s_items[0] >> 32), max2 = (double)(uint32_t)(s_items[1] >> 32),
		             max3 = (double)(uint32_t)(s_items[2] >> 32);
		int eraseBin = 3;
		if (max2 < 0.1 * max1) eraseBin = 1;
		else if (max3 < 0.1 * max1) eraseBin = 2;
		uint32_t mask = 0;
		int reduction = 0;
		for (int b = eraseBin; b < HISTO_LENGTH; b++)
		{
			mask |= 1u << (uint32_t)(s_items[b] & 31u);
			reduction += (int)(s_items[b] >> 32);
		}
		s_misc[0] = (int)mask;
		s_misc[1] = total - reduction;
	}
	__syncthreads();
	const uint32_t mask = (uint32_t)s_misc[0];
	for (int i = tid; i < A.npts; i += G_THREADS)
		if (choice[i] >= 0 && ((mask >> bin_of[i]) & 1u)) erase(i);
	return s_misc[1];
}

__device__ __forceinline__ int orientation_bin(float a1, float a2)
{
	float diff = a1 - a2;
	if (diff < 0) diff += 360;
	int bin = __float2int_rn((1.f / HISTO_LENGTH) * diff);   // cvRound
	if (bin == HISTO_LENGTH) bin = 0;
	// the reference asserts 0 <= bin < 30 (:279) and would throw on angles outside [0, 360); keep the shared histogram in bounds
	return min(max(bin, 0), HISTO_LENGTH - 1);
}

__device__ __forceinline__ void stamp(const GuidedArgs& A, int k)
{
	if (threadIdx.x == 0)
	{
		unsigned long long t;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
		A.stamps[k] = t;
	}
}

// One search = one cluster of G_CLUSTER CTAs. The state-free phases (windows, enumeration, distances) run on all of them, separated by
// cluster barriers (release/acquire at cluster scope, so the global arrays written before are visible after); the rounds and the
// finalisation run in CTA 0, whose shared memory holds the state.
__global__ void __cluster_dims__(G_CLUSTER, 1, 1) __launch_bounds__(G_THREADS, 1) k_guided_search(const GuidedArgs A)
{
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
	const int rank = (int)cluster.block_rank();
	const int gid = rank * G_THREADS + (int)threadIdx.x, gstride = G_CLUSTER * G_THREADS;
	__shared__ int s_w[G_THREADS / 32 + 1];
	__shared__ int s_hist[HISTO_LENGTH];
	__shared__ uint64_t s_items[HISTO_LENGTH];
	__shared__ int s_misc[2];
	__shared__ int s_count;
	extern __shared__ int s_dyn[];
	const int tid = threadIdx.x;
	GridDev G = A.G;
	const int npts = A.npts, n2 = G.n;
	// the arrays every phase chases pointers through live in shared memory when they fit (the host decides): the grid's cell
	// offsets, the per-keypoint owner / claim heads, the per-point list offsets and picks
	int* owner = A.owner;
	int* off = A.off;                  // written by every CTA: global
	int* choice = A.choice;
	{
		int* cs = s_dyn;               // every CTA keeps its own copy of the cell offsets
		for (int c = tid; c <= GRID_CELLS; c += G_THREADS) cs[c] = A.G.cell_start[c];
		G.cell_start = cs;
		if (A.use_smem)
		{
			owner = cs + GRID_CELLS + 1;
			choice = owner + 2 * n2;     // owner's initial value sits between them (see the rounds)
		}
		__syncthreads();
	}

	if (rank == 0) stamp(A, 0);
	if (rank == 0)
	{
		for (int i = tid; i < npts; i += G_THREADS) choice[i] = -1;
		if (tid == 0) A.cursor[A.parity ^ 1] = 0;
	}
	// ---- the window of every point and its candidate list in GetFeaturesInArea order. `lanes` consecutive threads share a point:
	//      each takes a contiguous block of the window's cell columns, counts its hits, the group's leader reserves the list with one
	//      atomicAdd, and a second walk (now out of L1) writes the entries at the lane's offset.
	const int L = A.lanes, lig = tid & (L - 1);
	const unsigned gmask = L == 32 ? 0xffffffffu : (((1u << L) - 1u) << ((tid & 31) & ~(L - 1)));
	int* const cursor = A.cursor + A.parity;
	if (A.mode == MODE_BOW)
	{
		// SearchByBoW: the candidates of a point are the features of frame 2 in the same vocabulary node, in the node's order (:477)
		for (int i = gid / L; i < npts; i += gstride / L)
		{
			const int s0 = A.bow_pt[3 * i + 1], len = A.bow_pt[3 * i + 2];
			const int q = (len + L - 1) / L;
			const int c0 = s0 + min(lig * q, len), c1 = s0 + min(lig * q + q, len);
			int c = 0;
			for (int p2 = c0; p2 < c1; p2++) c += !A.bow_valid2 || A.bow_valid2[A.bow_idx2[p2]];     // :731-732
			int inc = c;
			for (int d = 1; d < L; d <<= 1)
			{
				const int o = __shfl_up_sync(gmask, inc, d, L);
				if (lig >= d) inc += o;
			}
			const int tot = __shfl_sync(gmask, inc, L - 1, L);
			int base = 0;
			if (lig == 0)
			{
				if (tot > 0) base = atomicAdd(cursor, tot);
				off[i] = base;
				A.len[i] = tot;
			}
			base = __shfl_sync(gmask, base, 0, L);
			int p = base + inc - c;
			if (base + tot <= A.cap)
				for (int p2 = c0; p2 < c1; p2++)
				{
					const uint32_t idx2 = A.bow_idx2[p2];
					if (A.bow_valid2 && !A.bow_valid2[idx2]) continue;
					A.list[p] = idx2;
					A.entry_pt[p] = i;
					p++;
				}
		}
	}
	else
	for (int i = gid / L; i < npts; i += gstride / L)
	{
		float u = 0.f, v = 0.f, ur = 0.f, radius = 0.f;
		int lv = pack_levels(INACTIVE, 0);
		if (A.mode == MODE_LOCAL_MAP)
		{
			const orbx_track_point p = A.tp[i];
			if (p.flags & 1)                                                   // :321-322
			{
				const int ps = p.scale_level;
				const float r = p.view_cos > 0.998 ? 2.5f : 4.f;                  // RadiusByViewingCos, :53 (float against a double literal)
				radius = A.th * r * A.sf[ps];                                     // :327
				u = p.proj_x; v = p.proj_y; ur = p.proj_xr;
				lv = pack_levels(ps - 1, ps);                                     // :332
			}
		}
		else if (A.mode == MODE_LAST_FRAME)
		{
			const orbx_last_point p = A.lp[i];
			if (p.flags & 1)                                                   // :1295-1297
			{
				// CameraProjection::WorldToCamera: cv::Matx product (s = 0; s += R(r,k) * Xw(k), k ascending), then + tcw
				float xc[3];
#pragma unroll
				for (int r = 0; r < 3; r++)
				{
					float s = 0.f;
#pragma unroll
					for (int k = 0; k < 3; k++) s = __fadd_rn(s, __fmul_rn(A.R[r * 3 + k], p.xw[k]));
					xc[r] = __fadd_rn(s, A.t[r]);
				}
				if (!(xc[2] < 0.f))                                            // :1302-1303
				{
					const float invZ = __fdiv_rn(1.f, xc[2]);                  // CameraToImage
					u = __fadd_rn(__fmul_rn(__fmul_rn(invZ, A.fx), xc[0]), A.cx);
					v = __fadd_rn(__fmul_rn(__fmul_rn(invZ, A.fy), xc[1]), A.cy);
					ur = __fsub_rn(u, __fdiv_rn(A.bf, xc[2]));                 // :1308
					if (u >= G.b.minx && u < G.b.maxx && v >= G.b.miny && v < G.b.maxy)   // ImageBounds::Contains, :1310
					{
						const int oct = p.octave;
						radius = A.th * A.sf[oct];                               // :1316
						lv = pack_levels(A.forward ? oct : (A.backward ? 0 : oct - 1), A.forward ? -1 : (A.backward ? oct : oct + 1));   // :1318-1319
					}
				}
			}
		}
		else if (A.mode == MODE_WINDOWS)
		{
			const GuidedWindow wd = A.win[i];
			if (wd.flags & 1)
			{
				u = wd.u; v = wd.v; radius = wd.radius;
				ur = __int_as_float(0x7fc00000);                                   // NaN: no stereo gate in this variant
				lv = wd.levels;
			}
		}
		else
		{
			if (!(A.kps1[i].octave > 0))                                       // :629-631
			{
				u = A.prev_in[2 * i]; v = A.prev_in[2 * i + 1];
				radius = A.radius;
				lv = pack_levels(0, 0);                                           // :637
			}
		}
		if (lig == 0)
		{
			A.pu[i] = u; A.pv[i] = v; A.pur[i] = ur; A.prad[i] = radius; A.plevels[i] = lv;
			if (A.mode == MODE_INIT) { A.aux0[i] = INT_MAX; A.aux1[i] = INT_MAX; }
		}
		// the window's cells (src/Frame.cc:112-118)
		const int mincx = max(__float2int_rd(G.invW * (u - radius - G.b.minx)), 0);
		const int maxcx = min(__float2int_ru(G.invW * (u + radius - G.b.minx)), GRID_COLS - 1);
		const int mincy = max(__float2int_rd(G.invH * (v - radius - G.b.miny)), 0);
		const int maxcy = min(__float2int_ru(G.invH * (v + radius - G.b.miny)), GRID_ROWS - 1);
		const bool live = level_lo(lv) != INACTIVE && !(mincx >= GRID_COLS || maxcx < 0 || mincy >= GRID_ROWS || maxcy < 0);
		const int ncols = live ? max(maxcx - mincx + 1, 0) : 0;
		const int q = (ncols + L - 1) / L;
		const int c0 = mincx + lig * q, c1 = min(c0 + q, mincx + ncols);
		int c = 0;
		for_columns(G, u, v, radius, level_lo(lv), level_hi(lv), c0, c1, mincy, maxcy, [&](int, int) { c++; });
		int inc = c;
		for (int d = 1; d < L; d <<= 1)
		{
			const int o = __shfl_up_sync(gmask, inc, d, L);
			if (lig >= d) inc += o;
		}
		const int tot = __shfl_sync(gmask, inc, L - 1, L);
		int base = 0;
		if (lig == 0)
		{
			if (tot > 0) base = atomicAdd(cursor, tot);
			off[i] = base;
			A.len[i] = tot;
		}
		base = __shfl_sync(gmask, base, 0, L);
		int p = base + inc - c;
		if (base + tot <= A.cap)       // otherwise the launch is repeated with larger arrays
			for_columns(G, u, v, radius, level_lo(lv), level_hi(lv), c0, c1, mincy, maxcy, [&](int idx, int oct) {
				A.list[p] = (uint32_t)idx | ((uint32_t)oct << E_IDX_BITS);
				A.entry_pt[p] = i;
				p++;
			});
	}
	cluster.sync();
	const int total = *cursor;
	if (total > A.cap)
	{
		if (gid == 0) { A.result[1] = total; A.result[3] = 1; }
		return;
	}
	if (gid == 0) { A.result[1] = total; A.result[3] = 0; }
	if (rank == 0) { stamp(A, 1); stamp(A, 2); }
	cluster.sync();
	if (rank == 0) stamp(A, 3);

	// ---- every distance, one thread per candidate (DescriptorDistance, :1449-1457), and the stereo gate (:342-343, :1333-1334)
	for (int e = gid; e < total; e += gstride)
	{
		const uint32_t ent = A.list[e];
		const int idx = (int)(ent & E_IDX_MASK), i = A.entry_pt[e];
		uint32_t d;
		const float ur2 = A.uright2[idx];
		if (A.mode == MODE_BOW) d = (uint32_t)hamming256(A.pt_desc + (size_t)A.bow_pt[3 * i] * 32, A.desc2 + (size_t)idx * 32);
		else if (A.mode != MODE_INIT && ur2 > 0 && fabsf(A.pur[i] - ur2) > A.prad[i]) d = E_SKIP;
		else d = (uint32_t)hamming256(A.pt_desc + (size_t)i * 32, A.desc2 + (size_t)idx * 32);
		A.list[e] = ent | (d << (E_IDX_BITS + E_LVL_BITS));
	}
	cluster.sync();
	if (rank != 0) return;
	stamp(A, 4);

	// the rounds re-scan every candidate list: keep the lists (and their bounds) in shared memory when they fit beside the state
	const uint32_t* lst = A.list;
	const int* lstart = off;
	const int* llen = A.len;
	if (A.use_smem && GRID_CELLS + 1 + 2 * n2 + 4 * npts + total <= A.smem_ints)
	{
		int* st = choice + 2 * npts;
		int* ln = st + npts;
		uint32_t* ls = reinterpret_cast<uint32_t*>(ln + npts);
		for (int i = tid; i < npts; i += G_THREADS) { st[i] = off[i]; ln[i] = A.len[i]; }
		for (int e = tid; e < total; e += G_THREADS) ls[e] = A.list[e];
		lst = ls; lstart = st; llen = ln;
		__syncthreads();
	}

	int rounds = 0;
	if (A.mode != MODE_INIT)
	{
		// ---- rounds: owner[c] = lowest point index with observations that takes keypoint c (-1: closed on entry)
		// Everything a round reads is staged once: owner's value before any claim (own0) and Observations() > 0 per point (obs),
		// in shared memory when the state fits there.
		const bool have_own0 = A.use_smem != 0;
		int* own0 = owner + n2;
		int* obs = have_own0 ? choice + npts : A.aux0;
		auto closed_on_entry = [&](int c) {
			if (A.mode == MODE_BOW) return false;                               // matches starts all null (:456, :709)
			const int m = A.mp_in[c];
			if (A.mode == MODE_WINDOWS) return m != -1;                         // any stored map point closes the keypoint (:1412-1413)
			return m == -2 || (m >= 0 && ((A.mode == MODE_LOCAL_MAP ? A.tp[m].flags : A.lp[m].flags) & 2));
		};
		if (have_own0)
			for (int c = tid; c < n2; c += G_THREADS) own0[c] = closed_on_entry(c) ? -1 : INT_MAX;
		for (int i = tid; i < npts; i += G_THREADS)
			obs[i] = (A.mode == MODE_BOW || A.mode == MODE_WINDOWS) ? 1 : (((A.mode == MODE_LOCAL_MAP ? A.tp[i].flags : A.lp[i].flags) & 2) ? 1 : 0);   // BoW: any match closes the keypoint (:477-478)
		__syncthreads();
		stamp(A, 7);
		for (;;)
		{
			for (int c = tid; c < n2; c += G_THREADS) owner[c] = have_own0 ? own0[c] : (closed_on_entry(c) ? -1 : INT_MAX);
			__syncthreads();
			for (int i = tid; i < npts; i += G_THREADS)
			{
				const int c = choice[i];
				if (c >= 0 && obs[i]) atomicMin(&owner[c], i);
			}
			__syncthreads();
			int changed = 0;
			for (int i = tid; i < npts; i += G_THREADS)
			{
				int best = 256, second = 256, bestLevel = -1, secondLevel = -1, bestIdx = -1;
				for (int e = lstart[i], e1 = lstart[i] + llen[i]; e < e1; e++)
				{
					const uint32_t ent = lst[e];
					const int d = (int)(ent >> (E_IDX_BITS + E_LVL_BITS));
					if (d == (int)E_SKIP) continue;
					const int idx = (int)(ent & E_IDX_MASK);
					if (owner[idx] < i) continue;                            // :339-340, :1330-1331
					const int lvl = (int)((ent >> E_IDX_BITS) & ((1u << E_LVL_BITS) - 1));
					if (d < best) { second = best; best = d; secondLevel = bestLevel; bestLevel = lvl; bestIdx = idx; }
					else if (d < second) { secondLevel = lvl; second = d; }
				}
				bool ok = best <= TH_HIGH;                                        // :370, :1349
				if (ok && A.mode == MODE_LOCAL_MAP && bestLevel == secondLevel && (float)best > A.nnratio * (float)second) ok = false;   // :372-373
				if (A.mode == MODE_WINDOWS) ok = bestIdx >= 0 && best <= A.max_dist;                                                  // :1424
				if (A.mode == MODE_BOW)
					ok = (A.bow_strict ? best < TH_LOW : best <= TH_LOW) && (float)best < A.nnratio * (float)second;                  // :501, :750
				const int c = ok ? bestIdx : -1;
				if (c != choice[i]) { choice[i] = c; changed = 1; }
			}
			rounds++;
			if (!__syncthreads_or(changed)) break;
		}

		stamp(A, 5);
		// ---- frame.mappoints: the last point (in loop order) that took a keypoint stays there
		if (tid == 0) s_count = 0;
		for (int b = tid; b < HISTO_LENGTH; b += G_THREADS) s_hist[b] = 0;
		for (int c = tid; c < n2; c += G_THREADS) owner[c] = -1;
		__syncthreads();
		int mine = 0;
		for (int i = tid; i < npts; i += G_THREADS)
		{
			const int c = choice[i];
			if (c < 0) continue;
			mine++;
			atomicMax(&owner[c], i);
			if (A.mode == MODE_BOW && A.check_ori)
			{
				const float a1 = A.kps1[A.bow_pt[3 * i]].angle, a2 = A.kps2[c].angle;
				const int bin = A.ori_swap ? orientation_bin(a2, a1) : orientation_bin(a1, a2);      // :763 / :512
				A.aux0[i] = bin;
				atomicAdd(&s_hist[bin], 1);
			}
			if (A.mode == MODE_WINDOWS && A.check_ori)
			{
				const int bin = orientation_bin(A.win[i].angle, A.kps2[c].angle);  // keypoints1 = keyframe->keypointsUn, :1444
				A.aux0[i] = bin;
				atomicAdd(&s_hist[bin], 1);
			}
			if (A.mode == MODE_LAST_FRAME && A.check_ori)
			{
				const int bin = orientation_bin(A.lp[i].angle, A.kps2[c].angle);   // keypoints1 = lastFrame.keypointsUn, :1358
				A.aux0[i] = bin;
				atomicAdd(&s_hist[bin], 1);
			}
		}
		if (mine) atomicAdd(&s_count, mine);
		__syncthreads();
		if (A.mode == MODE_BOW)
			for (int c = tid; c < n2; c += G_THREADS) A.frame_mp[c] = owner[c] >= 0 ? A.bow_pt[3 * owner[c]] : -1;   // the keypoint of frame 1
		else
			for (int c = tid; c < n2; c += G_THREADS) A.frame_mp[c] = owner[c] >= 0 ? owner[c] : A.mp_in[c];
		__syncthreads();
		int nmatches = s_count;
		if ((A.mode == MODE_LAST_FRAME || A.mode == MODE_BOW || A.mode == MODE_WINDOWS) && A.check_ori)
			nmatches = check_orientation(A, choice, A.aux0, s_hist, s_items, s_misc, [&](int i) { A.frame_mp[choice[i]] = -1; });   // :298-304
		if (tid == 0) { A.result[0] = nmatches; A.result[2] = rounds; }
		stamp(A, 6);
		return;
	}

	// ---- SearchForInitialization: matchedDistance[c] as seen by point i = the smallest accepted distance among points j < i on c
	int* bd_old = A.aux0;
	int* bd_new = A.aux1;
	for (;;)
	{
		for (int c = tid; c < n2; c += G_THREADS) owner[c] = -1;         // heads of the claim lists
		__syncthreads();
		for (int i = tid; i < npts; i += G_THREADS)
			if (choice[i] >= 0) A.next[i] = atomicExch(&owner[choice[i]], i);
		__syncthreads();
		int changed = 0;
		for (int i = tid; i < npts; i += G_THREADS)
		{
			int best = INT_MAX, second = INT_MAX, bestIdx = -1;
			for (int e = lstart[i], e1 = lstart[i] + llen[i]; e < e1; e++)
			{
				const uint32_t ent = lst[e];
				const int d = (int)(ent >> (E_IDX_BITS + E_LVL_BITS));
				const int idx = (int)(ent & E_IDX_MASK);
				int md = INT_MAX;
				for (int j = owner[idx]; j >= 0; j = A.next[j])
					if (j < i) md = min(md, bd_old[j]);
				if (md <= d) continue;                                         // :651-652
				if (d < best) { second = best; best = d; bestIdx = idx; }
				else if (d < second) second = d;
			}
			const bool ok = bestIdx >= 0 && best <= TH_LOW && (float)best < (float)second * A.nnratio;   // :666
			const int c = ok ? bestIdx : -1;
			const int bd = ok ? best : INT_MAX;
			if (c != choice[i] || bd != bd_old[i]) changed = 1;
			choice[i] = c;
			bd_new[i] = bd;
		}
		rounds++;
		int* t = bd_old; bd_old = bd_new; bd_new = t;
		if (!__syncthreads_or(changed)) break;
	}
	stamp(A, 5);
	// matches21[c] = the last accepted point on c; earlier ones were revoked (:668-672)
	if (tid == 0) s_count = 0;
	for (int b = tid; b < HISTO_LENGTH; b += G_THREADS) s_hist[b] = 0;
	for (int c = tid; c < n2; c += G_THREADS) owner[c] = -1;
	__syncthreads();
	int accepted = 0;
	for (int i = tid; i < npts; i += G_THREADS)
	{
		const int c = choice[i];
		if (c < 0) continue;
		accepted++;
		atomicMax(&owner[c], i);
		if (A.check_ori)
		{
			const int bin = orientation_bin(A.kps2[c].angle, A.kps1[i].angle);     // keypoints1 = frame2.keypointsUn, :686
			bd_new[i] = bin;
			atomicAdd(&s_hist[bin], 1);
		}
	}
	__syncthreads();
	int distinct = 0;
	for (int i = tid; i < npts; i += G_THREADS)
	{
		const int c = choice[i];
		const bool keep = c >= 0 && owner[c] == i;
		A.frame_mp[i] = keep ? c : -1;
		distinct += keep;
	}
	atomicAdd(&s_count, A.check_ori ? accepted : distinct);   // with CheckOrientation the count restarts from matchIds.size() (:686)
	__syncthreads();
	int nmatches = s_count;
	if (A.check_ori) nmatches = check_orientation(A, choice, bd_new, s_hist, s_items, s_misc, [&](int i) { A.frame_mp[i] = -1; });
	__syncthreads();
	for (int i = tid; i < npts; i += G_THREADS)
	{
		const int c = A.frame_mp[i];
		A.prev[2 * i] = c >= 0 ? A.kps2[c].x : A.prev_in[2 * i];                          // :689-691
		A.prev[2 * i + 1] = c >= 0 ? A.kps2[c].y : A.prev_in[2 * i + 1];
	}
	if (tid == 0) { A.result[0] = nmatches; A.result[2] = rounds; }
	stamp(A, 6);
}

template <class T> struct GBuf
{
	T* p = nullptr;
	size_t n = 0;
	cudaError_t ensure(size_t count)
	{
		if (count <= n) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; n = 0;
		const size_t want = count + count / 2 + 64;
		cudaError_t e = cudaMalloc(&p, want * sizeof(T));
		if (e == cudaSuccess) n = want;
		return e;
	}
	~GBuf() { if (p) cudaFree(p); }
	GBuf() = default;
	GBuf(const GBuf&) = delete;
	GBuf& operator=(const GBuf&) = delete;
};

#define GCU(call)                                                                                             \
	do {                                                                                                      \
		cudaError_t e_ = (call);                                                                              \
		if (e_ != cudaSuccess)                                                                                \
			return orbx_fail(ORBX_ERR_CUDA, (std::string(#call) + ": " + cudaGetErrorString(e_)).c_str());     \
	} while (0)

}  // namespace

struct orbx_frame_s
{
	int device = 0;
	int n = 0, nlevels = 0;
	orbx_bounds b{};
	float invW = 0.f, invH = 0.f;
	float sf[16] = {};
	cudaStream_t st = nullptr;
	int last_rounds = 0;
	float last_kernel_ms = 0.f;
	unsigned long long last_stamps[8] = {};
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	// one pinned staging buffer each way: a search is one H2D copy, one kernel, one D2H copy
	uint8_t* h_in = nullptr; size_t h_in_cap = 0;
	uint8_t* h_out = nullptr; size_t h_out_cap = 0;
	GBuf<uint8_t> d_in, d_out;
	GBuf<orbx_keypoint> kps;
	GBuf<uint8_t> desc;
	GBuf<float> uright;
	GBuf<int> cell_start, cell_of;
	GBuf<int4> rec;
	// per-search scratch
	GBuf<float> pf;
	GBuf<int> pi, owner, entry_pt, cursor;
	int launches = 0, last_entries = 0;
	GBuf<uint32_t> list;
	GBuf<float> qf;
	GBuf<int> qi, qoff, qidx;
	std::vector<float> h_angle;     // keypoint angles on the host (CheckOrientation of the host-side matchers); empty = not fetched yet

	cudaError_t ensure_staging(size_t in_bytes, size_t out_bytes)
	{
		cudaError_t e;
		if (in_bytes > h_in_cap)
		{
			if (h_in) cudaFreeHost(h_in);
			h_in = nullptr; h_in_cap = 0;
			if ((e = cudaMallocHost(&h_in, in_bytes * 2)) != cudaSuccess) return e;
			h_in_cap = in_bytes * 2;
		}
		if (out_bytes > h_out_cap)
		{
			if (h_out) cudaFreeHost(h_out);
			h_out = nullptr; h_out_cap = 0;
			if ((e = cudaMallocHost(&h_out, out_bytes * 2)) != cudaSuccess) return e;
			h_out_cap = out_bytes * 2;
		}
		if ((e = d_in.ensure(in_bytes)) != cudaSuccess) return e;
		return d_out.ensure(out_bytes);
	}

	GridDev grid() const
	{
		GridDev G;
		G.n = n; G.nlevels = nlevels; G.b = b; G.invW = invW; G.invH = invH;
		G.cell_start = cell_start.p; G.rec = rec.p;
		return G;
	}
	~orbx_frame_s()
	{
		if (h_in) cudaFreeHost(h_in);
		if (h_out) cudaFreeHost(h_out);
		if (ev0) cudaEventDestroy(ev0);
		if (ev1) cudaEventDestroy(ev1);
		if (st) cudaStreamDestroy(st);
	}
};

namespace {

inline size_t up16(size_t v) { return (v + 15) & ~(size_t)15; }

// Layout of one search in the staging buffers. Input: [points | point descriptors | frame_mp or prevMatched]; output:
// [result (4 ints) | frame_mp / matches12 | prevMatched]. One copy each way.
struct Staging
{
	size_t in_pts = 0, in_desc = 0, in_state = 0, in_bytes = 0;
	size_t out_res = 0, out_mp = 0, out_prev = 0, out_bytes = 0;
};

// fills the frame part and the scratch pointers of the kernel arguments; npts-sized scratch is carved from f->pf / f->pi
orbx_status prepare(orbx_frame_s* f, int npts, size_t pts_bytes, size_t desc_bytes, size_t state_bytes, size_t mp_entries, size_t prev_bytes,
                    GuidedArgs& A, Staging& S)
{
	GCU(cudaSetDevice(f->device));
	GCU(f->pf.ensure((size_t)npts * 4 + 4));
	GCU(f->pi.ensure((size_t)npts * 7 + 8));
	GCU(f->owner.ensure((size_t)f->n + 1));
	const size_t want = (size_t)npts * 48 + 4096;
	if (f->list.n < want)
	{
		GCU(f->list.ensure(want));
		GCU(f->entry_pt.ensure(want));
	}
	S.in_pts = 0; S.in_desc = up16(pts_bytes); S.in_state = S.in_desc + up16(desc_bytes); S.in_bytes = S.in_state + up16(state_bytes) + 16;
	S.out_res = 0; S.out_mp = 16 + 8 * sizeof(unsigned long long); S.out_prev = S.out_mp + up16(mp_entries * sizeof(int)); S.out_bytes = S.out_prev + up16(prev_bytes) + 16;
	GCU(f->ensure_staging(S.in_bytes, S.out_bytes));
	A.G = f->grid();
	A.kps2 = f->kps.p; A.desc2 = f->desc.p; A.uright2 = f->uright.p;
	for (int i = 0; i < 16; i++) A.sf[i] = f->sf[i];
	A.npts = npts;
	A.pu = f->pf.p; A.pv = A.pu + npts; A.pur = A.pv + npts; A.prad = A.pur + npts;
	A.plevels = f->pi.p; A.off = A.plevels + npts; A.choice = A.off + npts + 1; A.aux0 = A.choice + npts; A.aux1 = A.aux0 + npts;
	A.next = A.aux1 + npts; A.len = A.next + npts;
	A.cursor = f->cursor.p;
	{
		int lanes = 1;
		while (lanes < 16 && (size_t)npts * lanes * 2 <= (size_t)G_CLUSTER * G_THREADS) lanes *= 2;
		A.lanes = lanes;
	}
	A.owner = f->owner.p;
	A.list = f->list.p; A.entry_pt = f->entry_pt.p; A.cap = (int)std::min(f->list.n, f->entry_pt.n);
	A.result = reinterpret_cast<int*>(f->d_out.p + S.out_res);
	A.stamps = reinterpret_cast<unsigned long long*>(f->d_out.p + S.out_res + 16);
	A.frame_mp = reinterpret_cast<int32_t*>(f->d_out.p + S.out_mp);
	A.tp = nullptr; A.lp = nullptr; A.kps1 = nullptr; A.prev = nullptr; A.pt_desc = nullptr; A.mp_in = nullptr; A.prev_in = nullptr;
	A.win = nullptr; A.max_dist = 0;
	A.bow_pt = nullptr; A.bow_idx2 = nullptr; A.bow_valid2 = nullptr; A.bow_strict = 0; A.ori_swap = 0;
	A.th = 0.f; A.nnratio = 0.f; A.radius = 0.f; A.fx = A.fy = A.cx = A.cy = A.bf = 0.f;
	for (int i = 0; i < 9; i++) A.R[i] = 0.f;
	for (int i = 0; i < 3; i++) A.t[i] = 0.f;
	A.forward = A.backward = A.check_ori = 0;
	return ORBX_OK;
}

// One H2D copy of the staged inputs, the search, one D2H copy of the staged outputs. Grows the candidate arrays and repeats when they
// were too small (the kernel has not touched its outputs then).
orbx_status run_search(orbx_frame_s* f, GuidedArgs& A, const Staging& S, int* nmatches)
{
	for (int attempt = 0; attempt < 2; attempt++)
	{
		if (attempt == 0) GCU(cudaMemcpyAsync(f->d_in.p, f->h_in, S.in_bytes, cudaMemcpyHostToDevice, f->st));
		GCU(cudaEventRecord(f->ev0, f->st));
		A.parity = f->launches++ & 1;
		const size_t state = (size_t)GRID_CELLS + 1 + 2 * (size_t)A.G.n + 2 * (size_t)A.npts;
		A.use_smem = state <= (size_t)G_SMEM_INTS;
		// room for the candidate lists too when they are likely to fit (sized from the last search on this frame); asking for all of
		// the SM's shared memory regardless would shrink L1 under the enumeration
		size_t want = A.use_smem ? state + 2 * (size_t)A.npts + std::max<size_t>((size_t)f->last_entries * 5 / 4, (size_t)A.npts * 8) : (size_t)GRID_CELLS + 1;
		if (want > (size_t)G_SMEM_INTS) want = A.use_smem ? state : (size_t)GRID_CELLS + 1;
		A.smem_ints = (int)want;
		k_guided_search<<<G_CLUSTER, G_THREADS, sizeof(int) * want, f->st>>>(A);
		GCU(cudaGetLastError());
		GCU(cudaEventRecord(f->ev1, f->st));
		GCU(cudaMemcpyAsync(f->h_out, f->d_out.p, S.out_bytes, cudaMemcpyDeviceToHost, f->st));
		GCU(cudaStreamSynchronize(f->st));
		const int* res = reinterpret_cast<const int*>(f->h_out + S.out_res);
		if (!res[3]) break;
		if (attempt == 1) return orbx_fail(ORBX_ERR_CUDA, "candidate list overflow after regrowth");
		GCU(f->list.ensure((size_t)res[1] + 1024));
		GCU(f->entry_pt.ensure((size_t)res[1] + 1024));
		A.list = f->list.p; A.entry_pt = f->entry_pt.p; A.cap = (int)std::min(f->list.n, f->entry_pt.n);
	}
	const int* res = reinterpret_cast<const int*>(f->h_out + S.out_res);
	f->last_rounds = res[2];
	f->last_entries = res[1];
	memcpy(f->last_stamps, f->h_out + S.out_res + 16, sizeof(f->last_stamps));
	cudaEventElapsedTime(&f->last_kernel_ms, f->ev0, f->ev1);
	if (nmatches) *nmatches = res[0];
	return ORBX_OK;
}

// uploads a frame view into `f` (buffers grow as needed) and rebuilds the grid
orbx_status assign_frame(orbx_frame_s* f, const orbx_frame_view* v)
{
	GCU(cudaSetDevice(f->device));
	f->n = v->n; f->nlevels = v->nlevels; f->b = v->bounds;
	f->invW = GRID_COLS / (v->bounds.maxx - v->bounds.minx);   // src/Frame.cc:73-74
	f->invH = GRID_ROWS / (v->bounds.maxy - v->bounds.miny);
	for (int i = 0; i < 16; i++) f->sf[i] = i < v->nlevels ? v->scale_factors[i] : 0.f;
	const size_t n = (size_t)std::max(v->n, 1);
	GCU(f->kps.ensure(n));
	GCU(f->desc.ensure(n * 32));
	GCU(f->uright.ensure(n));
	GCU(f->cell_start.ensure(GRID_CELLS + 1));
	GCU(f->cell_of.ensure(n));
	GCU(f->rec.ensure(n));
	// one staged copy: [keypoints | descriptors | uright]
	const size_t o_desc = up16((size_t)v->n * sizeof(orbx_keypoint)), o_ur = o_desc + up16((size_t)v->n * 32);
	const size_t bytes = o_ur + up16((size_t)v->n * sizeof(float)) + 16;
	GCU(f->ensure_staging(bytes, 64));
	if (v->n > 0)
	{
		memcpy(f->h_in, v->kps_un, (size_t)v->n * sizeof(orbx_keypoint));
		f->h_angle.resize((size_t)v->n);
		for (int i = 0; i < v->n; i++) f->h_angle[i] = v->kps_un[i].angle;
		memcpy(f->h_in + o_desc, v->desc, (size_t)v->n * 32);
		float* ur = reinterpret_cast<float*>(f->h_in + o_ur);
		if (v->uright) memcpy(ur, v->uright, (size_t)v->n * sizeof(float));
		else for (int i = 0; i < v->n; i++) ur[i] = -1.f;   // a monocular Frame: uright = -1 everywhere (src/Frame.cc, monocular constructor)
		GCU(cudaMemcpyAsync(f->d_in.p, f->h_in, bytes, cudaMemcpyHostToDevice, f->st));
		GCU(cudaMemcpyAsync(f->kps.p, f->d_in.p, (size_t)v->n * sizeof(orbx_keypoint), cudaMemcpyDeviceToDevice, f->st));
		GCU(cudaMemcpyAsync(f->desc.p, f->d_in.p + o_desc, (size_t)v->n * 32, cudaMemcpyDeviceToDevice, f->st));
		GCU(cudaMemcpyAsync(f->uright.p, f->d_in.p + o_ur, (size_t)v->n * sizeof(float), cudaMemcpyDeviceToDevice, f->st));
	}
	k_grid_build<<<1, G_THREADS, 0, f->st>>>(f->kps.p, f->n, f->b, f->invW, f->invH, f->cell_start.p, f->rec.p, f->cell_of.p);
	GCU(cudaGetLastError());
	GCU(cudaStreamSynchronize(f->st));
	return ORBX_OK;
}

__global__ void k_fill_float(float* p, int n, float v)
{
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i < n) p[i] = v;
}

orbx_status check_view(const orbx_frame_view* v)
{
	if (!v) return orbx_fail(ORBX_ERR_INVALID, "null argument");
	if (v->n < 0 || (v->n > 0 && (!v->kps_un || !v->desc))) return orbx_fail(ORBX_ERR_INVALID, "frame without keypoints/descriptors");
	if (v->n >= (1 << E_IDX_BITS)) return orbx_fail(ORBX_ERR_INVALID, "more than 524287 keypoints in a frame");
	if (v->nlevels < 1 || v->nlevels > 16 || !v->scale_factors) return orbx_fail(ORBX_ERR_INVALID, "nlevels must be in [1, 16]");
	if (!(v->bounds.maxx > v->bounds.minx) || !(v->bounds.maxy > v->bounds.miny))
		return orbx_fail(ORBX_ERR_INVALID, "empty image bounds (the reference divides by zero, src/Frame.cc:73-74)");
	for (int i = 0; i < v->n; i++)
		if (v->kps_un[i].octave < 0 || v->kps_un[i].octave >= 16) return orbx_fail(ORBX_ERR_INVALID, "keypoint octave outside [0, 16)");
	return ORBX_OK;
}

}  // namespace

extern "C" {

orbx_status orbx_frame_create(const orbx_frame_view* v, int device, orbx_frame* out)
{
	if (!out) return orbx_fail(ORBX_ERR_INVALID, "null argument");
	if (orbx_status s = check_view(v)) return s;
	const char* why = nullptr;
	if (!orbx_device_usable(device, &why)) return orbx_fail(ORBX_ERR_CUDA, why);
	GCU(cudaSetDevice(device));
	orbx_frame_s* f = new orbx_frame_s;
	f->device = device;
	cudaError_t e;
	if ((e = cudaStreamCreateWithFlags(&f->st, cudaStreamNonBlocking)) != cudaSuccess || (e = cudaEventCreate(&f->ev0)) != cudaSuccess ||
	    (e = cudaEventCreate(&f->ev1)) != cudaSuccess)
	{
		delete f;
		return orbx_fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
	}
	// function attributes are per device: set on every create (cheap), after cudaSetDevice above
	if ((e = cudaFuncSetAttribute(k_guided_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(int) * G_SMEM_INTS)) != cudaSuccess)
	{
		delete f;
		return orbx_fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
	}
	if ((e = f->cursor.ensure(2)) != cudaSuccess || (e = cudaMemset(f->cursor.p, 0, 2 * sizeof(int))) != cudaSuccess)
	{
		delete f;
		return orbx_fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
	}
	if (orbx_status s = assign_frame(f, v))
	{
		delete f;
		return s;
	}
	*out = f;
	return ORBX_OK;
}

orbx_status orbx_frame_assign(orbx_frame f, const orbx_frame_view* v)
{
	if (!f) return orbx_fail(ORBX_ERR_INVALID, "null handle");
	if (orbx_status s = check_view(v)) return s;
	return assign_frame(f, v);
}

orbx_status orbx_frame_assign_device(orbx_frame f, const orbx_keypoint* d_kps_un, const uint8_t* d_desc, const float* d_uright, int n,
                                     const orbx_bounds* bounds, int nlevels, const float* scale_factors)
{
	if (!f || !bounds || !scale_factors || n < 0 || (n > 0 && (!d_kps_un || !d_desc))) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	if (n >= (1 << E_IDX_BITS)) return orbx_fail(ORBX_ERR_INVALID, "more than 524287 keypoints in a frame");
	if (nlevels < 1 || nlevels > 16) return orbx_fail(ORBX_ERR_INVALID, "nlevels must be in [1, 16]");
	if (!(bounds->maxx > bounds->minx) || !(bounds->maxy > bounds->miny))
		return orbx_fail(ORBX_ERR_INVALID, "empty image bounds (the reference divides by zero, src/Frame.cc:73-74)");
	GCU(cudaSetDevice(f->device));
	f->n = n; f->nlevels = nlevels; f->b = *bounds;
	f->h_angle.clear();
	f->invW = GRID_COLS / (bounds->maxx - bounds->minx);
	f->invH = GRID_ROWS / (bounds->maxy - bounds->miny);
	for (int i = 0; i < 16; i++) f->sf[i] = i < nlevels ? scale_factors[i] : 0.f;
	const size_t cap = (size_t)std::max(n, 1);
	GCU(f->kps.ensure(cap)); GCU(f->desc.ensure(cap * 32)); GCU(f->uright.ensure(cap));
	GCU(f->cell_start.ensure(GRID_CELLS + 1)); GCU(f->cell_of.ensure(cap)); GCU(f->rec.ensure(cap));
	if (n > 0)
	{
		// the frame keeps its own copy: the extractor's output buffers are overwritten by the next Extract
		GCU(cudaMemcpyAsync(f->kps.p, d_kps_un, (size_t)n * sizeof(orbx_keypoint), cudaMemcpyDeviceToDevice, f->st));
		GCU(cudaMemcpyAsync(f->desc.p, d_desc, (size_t)n * 32, cudaMemcpyDeviceToDevice, f->st));
		if (d_uright) GCU(cudaMemcpyAsync(f->uright.p, d_uright, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, f->st));
		else k_fill_float<<<(n + 255) / 256, 256, 0, f->st>>>(f->uright.p, n, -1.f);
	}
	k_grid_build<<<1, G_THREADS, 0, f->st>>>(f->kps.p, f->n, f->b, f->invW, f->invH, f->cell_start.p, f->rec.p, f->cell_of.p);
	GCU(cudaGetLastError());
	GCU(cudaStreamSynchronize(f->st));
	return ORBX_OK;
}

orbx_status orbx_frame_destroy(orbx_frame f)
{
	if (!f) return ORBX_OK;
	cudaSetDevice(f->device);
	delete f;
	return ORBX_OK;
}

orbx_status orbx_frame_grid(orbx_frame f, int32_t* cell_start, int32_t* items, int cap, int* n_items)
{
	if (!f || !cell_start || !n_items) return orbx_fail(ORBX_ERR_INVALID, "null argument");
	GCU(cudaSetDevice(f->device));
	GCU(cudaMemcpy(cell_start, f->cell_start.p, (GRID_CELLS + 1) * sizeof(int), cudaMemcpyDeviceToHost));
	*n_items = cell_start[GRID_CELLS];
	if (*n_items > cap || (!items && *n_items > 0)) return orbx_fail(ORBX_ERR_CAPACITY, "items buffer too small");
	std::vector<int4> rec((size_t)*n_items);
	if (*n_items) GCU(cudaMemcpy(rec.data(), f->rec.p, rec.size() * sizeof(int4), cudaMemcpyDeviceToHost));
	for (int i = 0; i < *n_items; i++) items[i] = rec[i].z;
	return ORBX_OK;
}

orbx_status orbx_frame_features_in_area(orbx_frame f, const float* xyr, const int32_t* levels, int nq, int32_t* offsets, int32_t* indices, int cap)
{
	if (!f || !xyr || !levels || !offsets || nq < 0 || cap < 0 || (cap > 0 && !indices)) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	GCU(cudaSetDevice(f->device));
	GCU(f->qf.ensure((size_t)nq * 3 + 1));
	GCU(f->qi.ensure((size_t)nq * 2 + 1));
	GCU(f->qoff.ensure((size_t)nq + 1));
	GCU(f->qidx.ensure((size_t)cap + 1));
	if (nq)
	{
		GCU(cudaMemcpyAsync(f->qf.p, xyr, (size_t)nq * 3 * sizeof(float), cudaMemcpyHostToDevice, f->st));
		GCU(cudaMemcpyAsync(f->qi.p, levels, (size_t)nq * 2 * sizeof(int), cudaMemcpyHostToDevice, f->st));
	}
	k_features_in_area<<<1, G_THREADS, 0, f->st>>>(f->grid(), f->qf.p, f->qi.p, nq, f->qoff.p, f->qidx.p, cap);
	GCU(cudaGetLastError());
	GCU(cudaMemcpyAsync(offsets, f->qoff.p, (size_t)(nq + 1) * sizeof(int), cudaMemcpyDeviceToHost, f->st));
	GCU(cudaStreamSynchronize(f->st));
	if (offsets[nq] > cap) return orbx_fail(ORBX_ERR_CAPACITY, "indices buffer too small; offsets[nq] holds the needed size");
	if (offsets[nq]) GCU(cudaMemcpy(indices, f->qidx.p, (size_t)offsets[nq] * sizeof(int), cudaMemcpyDeviceToHost));
	return ORBX_OK;
}

orbx_status orbx_search_by_projection_local_map(orbx_frame f, int32_t* frame_mp, const orbx_track_point* pts, const uint8_t* pt_desc, int npts,
                                                float th, float nnratio, int* nmatches)
{
	if (!f || !frame_mp || npts < 0 || (npts > 0 && (!pts || !pt_desc))) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	for (int i = 0; i < npts; i++)
		if ((pts[i].flags & 1) && (pts[i].scale_level < 0 || pts[i].scale_level >= f->nlevels))
			return orbx_fail(ORBX_ERR_INVALID, "trackScaleLevel outside the pyramid (the reference indexes scaleFactors out of range)");
	for (int c = 0; c < f->n; c++)
		if (frame_mp[c] < -3 || frame_mp[c] >= npts) return orbx_fail(ORBX_ERR_INVALID, "frame_mp entry is not -3..-1 or a point index");
	GuidedArgs A;
	Staging S;
	const size_t pts_bytes = (size_t)npts * sizeof(orbx_track_point), mp_bytes = (size_t)f->n * sizeof(int);
	if (orbx_status s = prepare(f, npts, pts_bytes, (size_t)npts * 32, mp_bytes, (size_t)f->n, 0, A, S)) return s;
	if (npts)
	{
		memcpy(f->h_in + S.in_pts, pts, pts_bytes);
		memcpy(f->h_in + S.in_desc, pt_desc, (size_t)npts * 32);
	}
	if (f->n) memcpy(f->h_in + S.in_state, frame_mp, mp_bytes);
	A.mode = MODE_LOCAL_MAP;
	A.tp = reinterpret_cast<const orbx_track_point*>(f->d_in.p + S.in_pts);
	A.pt_desc = f->d_in.p + S.in_desc;
	A.mp_in = reinterpret_cast<const int32_t*>(f->d_in.p + S.in_state);
	A.th = th; A.nnratio = nnratio;
	if (orbx_status s = run_search(f, A, S, nmatches)) return s;
	if (f->n) memcpy(frame_mp, f->h_out + S.out_mp, mp_bytes);
	return ORBX_OK;
}

orbx_status orbx_search_by_projection_last_frame(orbx_frame f, const orbx_camera* cam, const orbx_pose* cp, const orbx_pose* lp, int32_t* frame_mp,
                                                 const orbx_last_point* pts, const uint8_t* pt_desc, int npts, float th, int monocular,
                                                 int check_orientation, int* nmatches)
{
	if (!f || !cam || !cp || !lp || !frame_mp || npts < 0 || (npts > 0 && (!pts || !pt_desc))) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	for (int i = 0; i < npts; i++)
		if ((pts[i].flags & 1) && (pts[i].octave < 0 || pts[i].octave >= f->nlevels))
			return orbx_fail(ORBX_ERR_INVALID, "octave outside the pyramid (the reference indexes scaleFactors out of range)");
	for (int c = 0; c < f->n; c++)
		if (frame_mp[c] < -3 || frame_mp[c] >= npts) return orbx_fail(ORBX_ERR_INVALID, "frame_mp entry is not -3..-1 or a point index");
	GuidedArgs A;
	Staging S;
	const size_t pts_bytes = (size_t)npts * sizeof(orbx_last_point), mp_bytes = (size_t)f->n * sizeof(int);
	if (orbx_status s = prepare(f, npts, pts_bytes, (size_t)npts * 32, mp_bytes, (size_t)f->n, 0, A, S)) return s;
	if (npts)
	{
		memcpy(f->h_in + S.in_pts, pts, pts_bytes);
		memcpy(f->h_in + S.in_desc, pt_desc, (size_t)npts * 32);
	}
	if (f->n) memcpy(f->h_in + S.in_state, frame_mp, mp_bytes);
	A.mode = MODE_LAST_FRAME;
	A.lp = reinterpret_cast<const orbx_last_point*>(f->d_in.p + S.in_pts);
	A.pt_desc = f->d_in.p + S.in_desc;
	A.mp_in = reinterpret_cast<const int32_t*>(f->d_in.p + S.in_state);
	A.th = th;
	A.fx = cam->fx; A.fy = cam->fy; A.cx = cam->cx; A.cy = cam->cy; A.bf = cam->bf;
	for (int i = 0; i < 9; i++) A.R[i] = cp->R[i];
	for (int i = 0; i < 3; i++) A.t[i] = cp->t[i];
	// tlc = Rlw * (-Rcw^T * tcw) + tlw (src/ORBmatcher.cc:1286); cv::Matx products accumulate from 0 in k order. Host float math,
	// compiled without contraction (build.py passes -ffp-contract=off).
	float twc[3], tlc2 = 0.f;
	for (int i = 0; i < 3; i++)
	{
		float s = 0.f;
		for (int k = 0; k < 3; k++) s += (cp->R[k * 3 + i] * -1) * cp->t[k];
		twc[i] = s;
	}
	{
		float s = 0.f;
		for (int k = 0; k < 3; k++) s += lp->R[2 * 3 + k] * twc[k];
		tlc2 = s + lp->t[2];
	}
	A.forward = tlc2 > cam->baseline && !monocular;      // :1287-1288
	A.backward = -tlc2 > cam->baseline && !monocular;
	A.check_ori = check_orientation != 0;
	if (orbx_status s = run_search(f, A, S, nmatches)) return s;
	if (f->n) memcpy(frame_mp, f->h_out + S.out_mp, mp_bytes);
	return ORBX_OK;
}

orbx_status orbx_search_for_initialization(orbx_frame f1, orbx_frame f2, float* prev_matched, int32_t* matches12, int window_size, float nnratio,
                                           int check_orientation, int* nmatches)
{
	if (!f1 || !f2 || !prev_matched || !matches12) return orbx_fail(ORBX_ERR_INVALID, "null argument");
	if (f1->device != f2->device) return orbx_fail(ORBX_ERR_INVALID, "frames live on different devices");
	const int npts = f1->n;
	GuidedArgs A;
	Staging S;
	const size_t prev_bytes = (size_t)npts * 2 * sizeof(float);
	if (orbx_status s = prepare(f2, npts, 0, 0, prev_bytes, (size_t)npts, prev_bytes, A, S)) return s;
	if (npts) memcpy(f2->h_in + S.in_state, prev_matched, prev_bytes);
	A.mode = MODE_INIT;
	A.kps1 = f1->kps.p;
	A.pt_desc = f1->desc.p;
	A.prev_in = reinterpret_cast<const float*>(f2->d_in.p + S.in_state);
	A.prev = reinterpret_cast<float*>(f2->d_out.p + S.out_prev);
	A.radius = (float)window_size;                          // :626
	A.nnratio = nnratio;
	A.check_ori = check_orientation != 0;
	if (orbx_status s = run_search(f2, A, S, nmatches)) return s;
	if (npts)
	{
		memcpy(matches12, f2->h_out + S.out_mp, (size_t)npts * sizeof(int));
		memcpy(prev_matched, f2->h_out + S.out_prev, prev_bytes);
	}
	return ORBX_OK;
}

namespace {
// the windows are already in the staging buffer: stage descriptors and frame.mappoints, search, copy frame.mappoints back
orbx_status finish_windows(orbx_frame_s* f, GuidedArgs& A, const Staging& S, int32_t* frame_mp, const uint8_t* pt_desc, int npts, int max_dist,
                           int check_orientation, int* nmatches)
{
	const size_t mp_bytes = (size_t)f->n * sizeof(int);
	if (npts) memcpy(f->h_in + S.in_desc, pt_desc, (size_t)npts * 32);
	if (f->n) memcpy(f->h_in + S.in_state, frame_mp, mp_bytes);
	A.mode = MODE_WINDOWS;
	A.win = reinterpret_cast<const GuidedWindow*>(f->d_in.p + S.in_pts);
	A.pt_desc = f->d_in.p + S.in_desc;
	A.mp_in = reinterpret_cast<const int32_t*>(f->d_in.p + S.in_state);
	A.max_dist = max_dist;
	A.check_ori = check_orientation != 0;
	if (orbx_status s = run_search(f, A, S, nmatches)) return s;
	if (f->n) memcpy(frame_mp, f->h_out + S.out_mp, mp_bytes);
	return ORBX_OK;
}
}  // namespace

orbx_status orbx_search_by_projection_keyframe(orbx_frame f, const orbx_camera* cam, const orbx_pose* pose, float log_scale_factor,
                                              int32_t* frame_mp, const orbx_keyframe_point* pts, const uint8_t* pt_desc, int npts, float th,
                                              int orb_dist, int check_orientation, int* nmatches)
{
	if (!f || !cam || !pose || !frame_mp || npts < 0 || (npts > 0 && (!pts || !pt_desc))) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	if (!(log_scale_factor > 0.f)) return orbx_fail(ORBX_ERR_INVALID, "logScaleFactor must be positive");
	for (int c = 0; c < f->n; c++)
		if (frame_mp[c] < -3 || frame_mp[c] >= npts) return orbx_fail(ORBX_ERR_INVALID, "frame_mp entry is not -3..-1 or a point index");
	GuidedArgs A;
	Staging S;
	const size_t win_bytes = (size_t)npts * sizeof(GuidedWindow), mp_bytes = (size_t)f->n * sizeof(int);
	if (orbx_status s = prepare(f, npts, win_bytes, (size_t)npts * 32, mp_bytes, (size_t)f->n, 0, A, S)) return s;
	// The geometry of src/ORBmatcher.cc:1376-1406 per map point, in the reference's operation order (cv::Matx products accumulate from
	// 0 in k order; cv::norm squares in double; MapPoint::PredictScale takes log in double). Host arithmetic, no contraction.
	GuidedWindow* win = reinterpret_cast<GuidedWindow*>(f->h_in + S.in_pts);
	float Ow[3];                                   // frame.GetCameraCenter() = -R^T * t (src/Frame.cc:203-206)
	for (int i = 0; i < 3; i++)
	{
		float s = 0.f;
		for (int k = 0; k < 3; k++) s += (pose->R[k * 3 + i] * -1) * pose->t[k];
		Ow[i] = s;
	}
	for (int i = 0; i < npts; i++)
	{
		GuidedWindow& w = win[i];
		w.u = w.v = w.radius = 0.f; w.angle = pts[i].angle; w.levels = 0; w.flags = 0;
		if (!(pts[i].flags & 1)) continue;         // :1379-1381
		float xc[3];
		for (int r = 0; r < 3; r++)
		{
			float s = 0.f;
			for (int k = 0; k < 3; k++) s += pose->R[r * 3 + k] * pts[i].xw[k];
			xc[r] = s + pose->t[r];
		}
		const float invZ = 1.f / xc[2];            // WorldToImage: this variant has no depth test (:1385)
		const float u = invZ * cam->fx * xc[0] + cam->cx, v = invZ * cam->fy * xc[1] + cam->cy;
		if (!(u >= f->b.minx && u < f->b.maxx && v >= f->b.miny && v < f->b.maxy)) continue;   // :1389
		double ss = 0;
		for (int k = 0; k < 3; k++) { const float d = pts[i].xw[k] - Ow[k]; ss += (double)d * (double)d; }
		const float dist3D = (float)std::sqrt(ss);                                             // cv::norm(PO), :1394
		const float maxDistance = 1.2f * pts[i].max_distance, minDistance = 0.8f * pts[i].min_distance;   // src/MapPoint.cc:382-392
		if (dist3D < minDistance || dist3D > maxDistance) continue;                            // :1400-1401
		const float ratio = pts[i].max_distance / dist3D;                                      // PredictScale, src/MapPoint.cc:405-414
		const int scale = (int)std::ceil(std::log((double)ratio) / log_scale_factor);
		const int ps = std::max(0, std::min(scale, f->nlevels - 1));
		w.u = u; w.v = v;
		w.radius = th * f->sf[ps];                 // :1406
		w.levels = ((ps - 1) & 0xffff) | ((ps + 1) << 16);   // :1408
		w.flags = 1;
	}
	return finish_windows(f, A, S, frame_mp, pt_desc, npts, orb_dist, check_orientation, nmatches);
}

orbx_status orbx_search_by_projection_sim3(orbx_frame f, const orbx_camera* cam, const orbx_sim3* Scw, float log_scale_factor, int32_t* matched,
                                          const orbx_sim3_point* pts, const uint8_t* pt_desc, int npts, int th, int* nmatches)
{
	if (!f || !cam || !Scw || !matched || npts < 0 || (npts > 0 && (!pts || !pt_desc))) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	if (!(log_scale_factor > 0.f)) return orbx_fail(ORBX_ERR_INVALID, "logScaleFactor must be positive");
	for (int c = 0; c < f->n; c++)
		if (matched[c] < -3 || matched[c] >= npts) return orbx_fail(ORBX_ERR_INVALID, "matched entry is not -3..-1 or a point index");
	GuidedArgs A;
	Staging S;
	if (orbx_status s = prepare(f, npts, (size_t)npts * sizeof(GuidedWindow), (size_t)npts * 32, (size_t)f->n * sizeof(int), (size_t)f->n, 0, A, S)) return s;
	// The geometry of src/ORBmatcher.cc:521-574 per map point in the reference's operation order: pose(Scw.R(), Scw.Invs() * Scw.t()),
	// cv::Matx products accumulating from 0, cv::norm in double, Matx::dot in float, PredictScale's log in double.
	GuidedWindow* win = reinterpret_cast<GuidedWindow*>(f->h_in + S.in_pts);
	const float invs = 1.f / Scw->s;
	float t[3], Ow[3];
	for (int i = 0; i < 3; i++) t[i] = Scw->t[i] * invs;
	for (int i = 0; i < 3; i++)
	{
		float s = 0.f;
		for (int k = 0; k < 3; k++) s += (Scw->R[k * 3 + i] * -1) * t[k];
		Ow[i] = s;
	}
	for (int i = 0; i < npts; i++)
	{
		GuidedWindow& w = win[i];
		w.u = w.v = w.radius = w.angle = 0.f; w.levels = 0; w.flags = 0;
		if (!(pts[i].flags & 1)) continue;         // :536-537
		float xc[3];
		for (int r = 0; r < 3; r++)
		{
			float s = 0.f;
			for (int k = 0; k < 3; k++) s += Scw->R[r * 3 + k] * pts[i].xw[k];
			xc[r] = s + t[r];
		}
		if (xc[2] < 0.f) continue;                 // :546-547
		const float invZ = 1.f / xc[2];
		const float u = invZ * cam->fx * xc[0] + cam->cx, v = invZ * cam->fy * xc[1] + cam->cy;
		if (!(u >= f->b.minx && u < f->b.maxx && v >= f->b.miny && v < f->b.maxy)) continue;   // IsInImage, :555-556
		float PO[3];
		double ss = 0;
		for (int k = 0; k < 3; k++) { PO[k] = pts[i].xw[k] - Ow[k]; ss += (double)PO[k] * (double)PO[k]; }
		const float dist = (float)std::sqrt(ss);
		const float maxDistance = 1.2f * pts[i].max_distance, minDistance = 0.8f * pts[i].min_distance;
		if (dist < minDistance || dist > maxDistance) continue;                                // :563-564
		float dot = 0.f;
		for (int k = 0; k < 3; k++) dot += PO[k] * pts[i].normal[k];
		if (dot < 0.5 * dist) continue;            // viewing angle below 60 degrees, :568-569 (compared in double)
		const float ratio = pts[i].max_distance / dist;                                        // PredictScale, src/MapPoint.cc:394-403
		const int scale = (int)std::ceil(std::log((double)ratio) / log_scale_factor);
		const int ps = std::max(0, std::min(scale, f->nlevels - 1));
		w.u = u; w.v = v;
		w.radius = th * f->sf[ps];                 // :574
		w.levels = ((ps - 1) & 0xffff) | (ps << 16);   // octave in [predictedScale - 1, predictedScale], :590-591
		w.flags = 1;
	}
	return finish_windows(f, A, S, matched, pt_desc, npts, TH_LOW, 0, nmatches);   // bestDist <= TH_LOW, :603
}

orbx_status orbx_search_windows(orbx_frame f, int32_t* frame_mp, const orbx_window* windows, const uint8_t* pt_desc, int npts, int max_dist,
                                int check_orientation, int* nmatches)
{
	if (!f || !frame_mp || npts < 0 || (npts > 0 && (!windows || !pt_desc))) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	for (int c = 0; c < f->n; c++)
		if (frame_mp[c] < -3 || frame_mp[c] >= npts) return orbx_fail(ORBX_ERR_INVALID, "frame_mp entry is not -3..-1 or a point index");
	GuidedArgs A;
	Staging S;
	if (orbx_status s = prepare(f, npts, (size_t)npts * sizeof(GuidedWindow), (size_t)npts * 32, (size_t)f->n * sizeof(int), (size_t)f->n, 0, A, S)) return s;
	GuidedWindow* win = reinterpret_cast<GuidedWindow*>(f->h_in + S.in_pts);
	for (int i = 0; i < npts; i++)
	{
		const orbx_window& w = windows[i];
		if (w.min_level < -32767 || w.min_level > 32767 || w.max_level < -32767 || w.max_level > 32767) return orbx_fail(ORBX_ERR_INVALID, "level out of range");
		win[i].u = w.u; win[i].v = w.v; win[i].radius = w.radius; win[i].angle = w.angle;
		win[i].levels = (w.min_level & 0xffff) | (w.max_level << 16);
		win[i].flags = w.flags & 1;
	}
	return finish_windows(f, A, S, frame_mp, pt_desc, npts, max_dist, check_orientation, nmatches);
}

orbx_status orbx_search_by_bow(orbx_frame f1, const orbx_feature_vector* fv1, const uint8_t* valid1, orbx_frame f2, const orbx_feature_vector* fv2,
                               const uint8_t* valid2, float nnratio, int check_orientation, int32_t* match2, int* nmatches)
{
	if (!f1 || !f2 || !fv1 || !fv2 || !valid1 || !match2) return orbx_fail(ORBX_ERR_INVALID, "null argument");
	if (f1->device != f2->device) return orbx_fail(ORBX_ERR_INVALID, "frames live on different devices");
	if (fv1->nnodes < 0 || fv2->nnodes < 0 || (fv1->nnodes && (!fv1->node_ids || !fv1->start || !fv1->indices)) ||
	    (fv2->nnodes && (!fv2->node_ids || !fv2->start || !fv2->indices)))
		return orbx_fail(ORBX_ERR_INVALID, "malformed feature vector");
	const int n_idx2 = fv2->nnodes ? fv2->start[fv2->nnodes] : 0;
	for (int p = 0; p < n_idx2; p++)
		if (fv2->indices[p] >= (uint32_t)f2->n) return orbx_fail(ORBX_ERR_INVALID, "feature index outside frame 2");
	// FeatureVectorIterator (src/ORBmatcher.cc:406-450): the nodes both vectors hold, ascending; the points are frame 1's valid
	// features of those nodes in the nodes' order
	std::vector<int> pts;
	for (int a = 0, b = 0; a < fv1->nnodes && b < fv2->nnodes;)
	{
		if (fv1->node_ids[a] < fv2->node_ids[b]) { a++; continue; }
		if (fv2->node_ids[b] < fv1->node_ids[a]) { b++; continue; }
		for (int p1 = fv1->start[a]; p1 < fv1->start[a + 1]; p1++)
		{
			const uint32_t idx1 = fv1->indices[p1];
			if (idx1 >= (uint32_t)f1->n) return orbx_fail(ORBX_ERR_INVALID, "feature index outside frame 1");
			if (!valid1[idx1]) continue;                                               // :471-472, :719-720
			pts.push_back((int)idx1); pts.push_back(fv2->start[b]); pts.push_back(fv2->start[b + 1] - fv2->start[b]);
		}
		a++; b++;
	}
	const int npts = (int)pts.size() / 3;
	GuidedArgs A;
	Staging S;
	const size_t pts_bytes = pts.size() * sizeof(int), idx_bytes = (size_t)n_idx2 * sizeof(uint32_t), val_bytes = valid2 ? (size_t)f2->n : 0;
	if (orbx_status s = prepare(f2, npts, pts_bytes, idx_bytes, val_bytes, (size_t)f2->n, 0, A, S)) return s;
	if (npts) memcpy(f2->h_in + S.in_pts, pts.data(), pts_bytes);
	if (n_idx2) memcpy(f2->h_in + S.in_desc, fv2->indices, idx_bytes);
	if (valid2 && f2->n) memcpy(f2->h_in + S.in_state, valid2, val_bytes);
	A.mode = MODE_BOW;
	A.bow_pt = reinterpret_cast<const int*>(f2->d_in.p + S.in_pts);
	A.bow_idx2 = reinterpret_cast<const uint32_t*>(f2->d_in.p + S.in_desc);
	A.bow_valid2 = valid2 ? f2->d_in.p + S.in_state : nullptr;
	A.bow_strict = valid2 != nullptr;                      // the KeyFrame-KeyFrame variant tests best < TH_LOW (:750)
	A.ori_swap = valid2 != nullptr;                        // and hands CheckOrientation (keypoints2, keypoints1) (:763)
	A.kps1 = f1->kps.p;
	A.pt_desc = f1->desc.p;
	A.nnratio = nnratio;
	A.check_ori = check_orientation != 0;
	if (orbx_status s = run_search(f2, A, S, nmatches)) return s;
	if (f2->n) memcpy(match2, f2->h_out + S.out_mp, (size_t)f2->n * sizeof(int));
	return ORBX_OK;
}

orbx_status orbx_frame_last_stats(orbx_frame f, int* rounds, float* kernel_ms, float* phase_us)
{
	if (!f) return orbx_fail(ORBX_ERR_INVALID, "null handle");
	if (rounds) *rounds = f->last_rounds;
	if (kernel_ms) *kernel_ms = f->last_kernel_ms;
	if (phase_us)
	{
		for (int k = 0; k < 6; k++) phase_us[k] = (float)(f->last_stamps[k + 1] - f->last_stamps[k]) * 1e-3f;
		phase_us[6] = f->last_stamps[7] > f->last_stamps[4] ? (float)(f->last_stamps[7] - f->last_stamps[4]) * 1e-3f : 0.f;   // staging before the rounds
	}
	return ORBX_OK;
}


// =====================================================================================================================================
// Matchers whose per-point search does not depend on the other points (local mapping / loop closing):
//   ORBmatcher::Fuse(keyframe, mappoints, th)                       src/ORBmatcher.cc:868-980
//   ORBmatcher::Fuse(keyframe, Scw, mappoints, th, replacePoints)   src/ORBmatcher.cc:982-1088
//   ORBmatcher::SearchBySim3(kf1, kf2, matches12, S12, th)          src/ORBmatcher.cc:1090-1277
//   ORBmatcher::SearchForTriangulation(kf1, kf2, F12, ids, stereo)  src/ORBmatcher.cc:768-866 (+ CheckDistEpipolarLine :384-404)
// In Fuse the map is mutated between points (Replace / AddObservation / AddMapPoint, :956-976), but what a point searches — projection,
// window, scale gate, chi-square gate, best distance, first keypoint on ties — reads only the key frame's keypoints and the point itself,
// so the search of all points runs at once and the caller replays :874-877 and :956-976 in order over (best_idx, best_dist).
// SearchForTriangulation never sets matched2 in this fork (:780, :814), so every keypoint of key frame 1 is independent as well.
// =====================================================================================================================================
extern "C++" {
namespace {

struct BestWindow { float u, v, radius, ur; int min_level, max_level; int flags; int pad; };   // flags bit0 = search, bit1 = Fuse's chi-square gate

// 8 lanes per point. A candidate's rank in GetFeaturesInArea's output order is (cell column, position in the column's record run), so
// "smallest distance, first in that order" is the minimum of distance << 22 | column << 16 | position.
__global__ void __launch_bounds__(256) k_best_in_windows(const GridDev G, const orbx_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc,
                                                         const float* __restrict__ uright, const BestWindow* __restrict__ win,
                                                         const uint8_t* __restrict__ pt_desc, const int npts, const float* __restrict__ inv_sigma_sq,
                                                         int2* __restrict__ out)
{
	const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
	if (i >= npts) return;
	const unsigned gmask = 0xffu << (threadIdx.x & 24);
	const BestWindow w = win[i];
	uint32_t best = 0xffffffffu;
	int besti = -1;
	if (w.flags & 1)
	{
		const float x = w.u, y = w.v, r = w.radius;
		const int mincx = max(__float2int_rd(G.invW * (x - r - G.b.minx)), 0);
		const int maxcx = min(__float2int_ru(G.invW * (x + r - G.b.minx)), GRID_COLS - 1);
		const int mincy = max(__float2int_rd(G.invH * (y - r - G.b.miny)), 0);
		const int maxcy = min(__float2int_ru(G.invH * (y + r - G.b.miny)), GRID_ROWS - 1);
		if (!(mincx >= GRID_COLS || maxcx < 0 || mincy >= GRID_ROWS || maxcy < 0))
		{
			const uint8_t* d1 = pt_desc + (size_t)i * 32;
			for (int cx = mincx; cx <= maxcx; cx++)
			{
				const int a = G.cell_start[cx * GRID_ROWS + mincy], b = G.cell_start[cx * GRID_ROWS + maxcy + 1];
				for (int p = a + sub; p < b; p += 8)
				{
					const int4 rec = G.rec[p];
					if (!(fabsf(__int_as_float(rec.x) - x) < r && fabsf(__int_as_float(rec.y) - y) < r)) continue;   // src/Frame.cc:136-139
					if (rec.w < w.min_level || rec.w > w.max_level) continue;                                       // :930-931, :1057-1058, :1175
					if (w.flags & 2)
					{
						// chi-square gate of Fuse (:934-945): float products and sums in source order, compared in double
						const float dx = __fsub_rn(x, __int_as_float(rec.x)), dy = __fsub_rn(y, __int_as_float(rec.y));
						const float ur2 = uright[rec.z];
						const float e2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
						const float is = inv_sigma_sq[rec.w];
						if (ur2 >= 0.f)
						{
							const float dz = __fsub_rn(w.ur, ur2);
							if ((double)__fmul_rn(__fadd_rn(e2, __fmul_rn(dz, dz)), is) > 7.8) continue;
						}
						else if ((double)__fmul_rn(e2, is) > 5.99) continue;
					}
					const uint32_t key = ((uint32_t)hamming256(d1, desc + (size_t)rec.z * 32) << 22) | ((uint32_t)(cx - mincx) << 16) | (uint32_t)min(p - a, 65535);
					if (key < best) { best = key; besti = rec.z; }
				}
			}
		}
	}
#pragma unroll
	for (int d = 1; d < 8; d <<= 1)
	{
		const uint32_t ob = __shfl_xor_sync(gmask, best, d);
		const int oi = __shfl_xor_sync(gmask, besti, d);
		if (ob < best) { best = ob; besti = oi; }
	}
	if (sub == 0) out[i] = make_int2(besti, besti >= 0 ? (int)(best >> 22) : 256);
}

// SearchForTriangulation: one thread per keypoint of key frame 1 that has a vocabulary node in common with key frame 2; its candidates
// are that node's features of key frame 2, scanned in order with the reference's `dist > TH_LOW || dist > bestDist` (a later candidate
// at the same distance replaces an earlier one).
struct TriArgs
{
	const orbx_keypoint* kps1; const uint8_t* desc1; const float* uright1;
	const orbx_keypoint* kps2; const uint8_t* desc2; const float* uright2;
	const int* item;             // [nitems][3]: idx1, first position in idx2 list, count
	const uint32_t* idx2;
	const uint8_t* has_mp2;
	int nitems, only_stereo;
	float F[9], ex, ey;
	float sf2[16], sigma_sq2[16];
	int* match;                  // [nitems] best idx2 or -1
};
__global__ void __launch_bounds__(128) k_triangulation_search(const TriArgs A)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= A.nitems) return;
	const int idx1 = A.item[3 * t], first = A.item[3 * t + 1], cnt = A.item[3 * t + 2];
	const bool stereo1 = A.uright1[idx1] >= 0.f;
	if (A.only_stereo && !stereo1) { A.match[t] = -1; return; }         // :800-802
	const orbx_keypoint k1 = A.kps1[idx1];
	// epipolar line in the second image (:391-393), products and sums in source order
	const float a = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, A.F[0]), __fmul_rn(k1.y, A.F[3])), A.F[6]);
	const float b = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, A.F[1]), __fmul_rn(k1.y, A.F[4])), A.F[7]);
	const float c = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, A.F[2]), __fmul_rn(k1.y, A.F[5])), A.F[8]);
	const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
	int bestDist = TH_LOW, bestIdx2 = -1;
	for (int j = 0; j < cnt; j++)
	{
		const int idx2 = (int)A.idx2[first + j];
		if (A.has_mp2[idx2]) continue;                                   // :814 (matched2 is never set)
		const bool stereo2 = A.uright2[idx2] >= 0.f;
		if (A.only_stereo && !stereo2) continue;
		const int dist = hamming256(A.desc1 + (size_t)idx1 * 32, A.desc2 + (size_t)idx2 * 32);
		if (dist > TH_LOW || dist > bestDist) continue;                  // :823
		const orbx_keypoint k2 = A.kps2[idx2];
		if (!stereo1 && !stereo2)
		{
			const float dx = __fsub_rn(A.ex, k2.x), dy = __fsub_rn(A.ey, k2.y);
			if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, A.sf2[k2.octave])) continue;   // :828-833
		}
		// CheckDistEpipolarLine (:395-403)
		const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, k2.x), __fmul_rn(b, k2.y)), c);
		if (den == 0.f) continue;
		const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
		if ((double)dsqr < __dmul_rn(3.84, (double)A.sigma_sq2[k2.octave])) { bestIdx2 = idx2; bestDist = dist; }
	}
	A.match[t] = bestIdx2;
}

// CheckOrientation (src/ORBmatcher.cc:249-309) on the host, for the matchers whose matches come back to the host anyway. std::sort is
// libstdc++'s own here (the host side is compiled by the GCC the reference's behaviour is pinned to), applied to 30 items with the
// reference's comparator: the same comparisons, hence the same order among equal sizes. pairs = (i1, i2): angle1[i1] - angle2[i2].
int orbx_check_orientation_host(const float* angle1, const float* angle2, const std::vector<std::pair<int, int>>& pairs, int32_t* status, int n)
{
	const float factor = 1.f / HISTO_LENGTH;
	std::vector<int> hist[HISTO_LENGTH];
	for (const auto& m : pairs)
	{
		float diff = angle1[m.first] - angle2[m.second];
		if (diff < 0) diff += 360;
		int bin = (int)std::nearbyint(factor * diff);                   // cvRound(float): round half to even
		if (bin == HISTO_LENGTH) bin = 0;
		if (bin < 0 || bin >= HISTO_LENGTH) continue;                   // CV_Assert in the reference
		hist[bin].push_back(m.second);
	}
	std::sort(std::begin(hist), std::end(hist), [](const std::vector<int>& lhs, const std::vector<int>& rhs) { return lhs.size() > rhs.size(); });
	const size_t max1 = hist[0].size(), max2 = hist[1].size(), max3 = hist[2].size();
	int eraseBin = 3;
	if (max2 < 0.1 * max1) eraseBin = 1;
	else if (max3 < 0.1 * max1) eraseBin = 2;
	int reduction = 0;
	for (int bin = eraseBin; bin < HISTO_LENGTH; bin++)
		for (int i2 : hist[bin]) { status[i2] = -1; reduction++; }
	return (int)pairs.size() - reduction;
}

// H2D (windows + descriptors [+ invSigmaSq]) -> kernel -> D2H on f's stream, through f's staging buffers
orbx_status best_in_windows(orbx_frame_s* f, const std::vector<BestWindow>& win, const uint8_t* pt_desc, const float* inv_sigma_sq, int32_t* best_idx,
                            int32_t* best_dist)
{
	const int npts = (int)win.size();
	if (npts == 0) return ORBX_OK;
	GCU(cudaSetDevice(f->device));
	const size_t o_desc = up16((size_t)npts * sizeof(BestWindow)), o_sig = o_desc + up16((size_t)npts * 32), in_bytes = o_sig + 64;
	const size_t out_bytes = (size_t)npts * sizeof(int2);
	GCU(f->ensure_staging(in_bytes, out_bytes));
	memcpy(f->h_in, win.data(), (size_t)npts * sizeof(BestWindow));
	memcpy(f->h_in + o_desc, pt_desc, (size_t)npts * 32);
	float sig[16] = {};
	if (inv_sigma_sq) for (int l = 0; l < f->nlevels && l < 16; l++) sig[l] = inv_sigma_sq[l];
	memcpy(f->h_in + o_sig, sig, sizeof(sig));
	GCU(cudaMemcpyAsync(f->d_in.p, f->h_in, in_bytes, cudaMemcpyHostToDevice, f->st));
	GCU(cudaEventRecord(f->ev0, f->st));
	k_best_in_windows<<<(npts * 8 + 255) / 256, 256, 0, f->st>>>(f->grid(), f->kps.p, f->desc.p, f->uright.p, reinterpret_cast<const BestWindow*>(f->d_in.p),
	                                                             f->d_in.p + o_desc, npts, reinterpret_cast<const float*>(f->d_in.p + o_sig),
	                                                             reinterpret_cast<int2*>(f->d_out.p));
	GCU(cudaGetLastError());
	GCU(cudaEventRecord(f->ev1, f->st));
	GCU(cudaMemcpyAsync(f->h_out, f->d_out.p, out_bytes, cudaMemcpyDeviceToHost, f->st));
	GCU(cudaStreamSynchronize(f->st));
	cudaEventElapsedTime(&f->last_kernel_ms, f->ev0, f->ev1);
	const int2* r = reinterpret_cast<const int2*>(f->h_out);
	for (int i = 0; i < npts; i++) { best_idx[i] = r[i].x; best_dist[i] = r[i].y; }
	return ORBX_OK;
}

// cv::Matx arithmetic in the reference's operation order (products accumulate from 0 in k order; no contraction on the host)
inline void matx_mul(const float* R, const float* x, float* y) { for (int r = 0; r < 3; r++) { float s = 0.f; for (int k = 0; k < 3; k++) s += R[r * 3 + k] * x[k]; y[r] = s; } }
inline void pose_invt(const float* R, const float* t, float* o) { for (int i = 0; i < 3; i++) { float s = 0.f; for (int k = 0; k < 3; k++) s += (R[k * 3 + i] * -1) * t[k]; o[i] = s; } }   // -R.t() * t
inline float norm3(const float* v) { double ss = 0; for (int k = 0; k < 3; k++) ss += (double)v[k] * (double)v[k]; return (float)std::sqrt(ss); }   // cv::norm
inline int predict_scale(float max_distance, float dist, float log_scale_factor, int nlevels)     // src/MapPoint.cc:394-403
{
	const float ratio = max_distance / dist;
	const int scale = (int)std::ceil(std::log((double)ratio) / log_scale_factor);
	return std::max(0, std::min(scale, nlevels - 1));
}
inline bool in_image(const orbx_frame_s* f, float u, float v) { return u >= f->b.minx && u < f->b.maxx && v >= f->b.miny && v < f->b.maxy; }

// the projection / distance / viewing-angle / scale part shared by the two Fuse variants (:879-919, :1005-1046); xc = camera coordinates
inline void fuse_window(const orbx_frame_s* f, const orbx_camera* cam, const float* xc, const float* Ow, const orbx_sim3_point& p, float log_scale_factor,
                        float th, bool gate, BestWindow& w)
{
	w = BestWindow{ 0.f, 0.f, 0.f, 0.f, 0, 0, 0, 0 };
	if (xc[2] < 0.f) return;
	const float invZ = 1.f / xc[2];
	const float u = invZ * cam->fx * xc[0] + cam->cx, v = invZ * cam->fy * xc[1] + cam->cy;
	if (!in_image(f, u, v)) return;
	const float ur = u - cam->bf / xc[2];                               // DepthToDisparity
	const float maxDistance = 1.2f * p.max_distance, minDistance = 0.8f * p.min_distance;   // src/MapPoint.cc:382-392
	float PO[3];
	for (int k = 0; k < 3; k++) PO[k] = p.xw[k] - Ow[k];
	const float dist3D = norm3(PO);
	if (dist3D < minDistance || dist3D > maxDistance) return;
	float dot = 0.f;
	for (int k = 0; k < 3; k++) dot += PO[k] * p.normal[k];
	if (dot < 0.5 * dist3D) return;                                     // compared in double
	const int ps = predict_scale(p.max_distance, dist3D, log_scale_factor, f->nlevels);
	w.u = u; w.v = v; w.ur = ur;
	w.radius = th * f->sf[ps];
	w.min_level = ps - 1; w.max_level = ps;
	w.flags = gate ? 3 : 1;
}

}  // namespace
}  // extern "C++"

orbx_status orbx_search_best_in_windows(orbx_frame f, const orbx_best_window* windows, const uint8_t* pt_desc, int npts, const float* inv_sigma_sq,
                                        int32_t* best_idx, int32_t* best_dist)
{
	if (!f || npts < 0 || (npts > 0 && (!windows || !pt_desc || !best_idx || !best_dist))) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	std::vector<BestWindow> win((size_t)npts);
	for (int i = 0; i < npts; i++)
	{
		const orbx_best_window& s = windows[i];
		if ((s.flags & 2) && !inv_sigma_sq) return orbx_fail(ORBX_ERR_INVALID, "the chi-square gate needs invSigmaSq");
		win[i] = BestWindow{ s.u, s.v, s.radius, s.ur, s.min_level, s.max_level, s.flags & 3, 0 };
	}
	return best_in_windows(f, win, pt_desc, inv_sigma_sq, best_idx, best_dist);
}

orbx_status orbx_fuse(orbx_frame f, const orbx_camera* cam, const orbx_pose* pose, float log_scale_factor, const float* inv_sigma_sq,
                      const orbx_sim3_point* pts, const uint8_t* pt_desc, int npts, float th, int32_t* best_idx, int32_t* best_dist)
{
	if (!f || !cam || !pose || !inv_sigma_sq || npts < 0 || (npts > 0 && (!pts || !pt_desc || !best_idx || !best_dist))) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	if (!(log_scale_factor > 0.f)) return orbx_fail(ORBX_ERR_INVALID, "logScaleFactor must be positive");
	float Ow[3];
	pose_invt(pose->R, pose->t, Ow);                                    // keyframe->GetCameraCenter()
	std::vector<BestWindow> win((size_t)npts);
	for (int i = 0; i < npts; i++)
	{
		win[i] = BestWindow{ 0.f, 0.f, 0.f, 0.f, 0, 0, 0, 0 };
		if (!(pts[i].flags & 1)) continue;                              // null or bad (:876-877); IsInKeyFrame is the caller's, at replay time
		float xc[3];
		matx_mul(pose->R, pts[i].xw, xc);
		for (int k = 0; k < 3; k++) xc[k] += pose->t[k];
		fuse_window(f, cam, xc, Ow, pts[i], log_scale_factor, th, true, win[i]);
	}
	return best_in_windows(f, win, pt_desc, inv_sigma_sq, best_idx, best_dist);
}

orbx_status orbx_fuse_sim3(orbx_frame f, const orbx_camera* cam, const orbx_sim3* Scw, float log_scale_factor, const orbx_sim3_point* pts,
                           const uint8_t* pt_desc, int npts, float th, int32_t* best_idx, int32_t* best_dist)
{
	if (!f || !cam || !Scw || npts < 0 || (npts > 0 && (!pts || !pt_desc || !best_idx || !best_dist))) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	if (!(log_scale_factor > 0.f)) return orbx_fail(ORBX_ERR_INVALID, "logScaleFactor must be positive");
	const float invs = 1.f / Scw->s;                                    // pose(Scw.R(), Scw.Invs() * Scw.t()), :987
	float t[3], Ow[3];
	for (int i = 0; i < 3; i++) t[i] = Scw->t[i] * invs;
	pose_invt(Scw->R, t, Ow);
	std::vector<BestWindow> win((size_t)npts);
	for (int i = 0; i < npts; i++)
	{
		win[i] = BestWindow{ 0.f, 0.f, 0.f, 0.f, 0, 0, 0, 0 };
		if (!(pts[i].flags & 1)) continue;                              // bad or already in the key frame (:1002-1003)
		float xc[3];
		matx_mul(Scw->R, pts[i].xw, xc);
		for (int k = 0; k < 3; k++) xc[k] += t[k];
		fuse_window(f, cam, xc, Ow, pts[i], log_scale_factor, th, false, win[i]);
	}
	return best_in_windows(f, win, pt_desc, nullptr, best_idx, best_dist);
}

orbx_status orbx_search_by_sim3(orbx_frame f1, const orbx_camera* cam1, const orbx_pose* pose1, float log_scale_factor1, orbx_frame f2,
                                const orbx_camera* cam2, const orbx_pose* pose2, float log_scale_factor2, const orbx_sim3* S12, float th,
                                const orbx_keyframe_point* pts1, const uint8_t* desc1, const orbx_keyframe_point* pts2, const uint8_t* desc2,
                                int32_t* match1, int32_t* match2, int32_t* matches12, int* nfound)
{
	if (!f1 || !f2 || !cam1 || !cam2 || !pose1 || !pose2 || !S12 || !matches12) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	if ((f1->n > 0 && (!pts1 || !desc1)) || (f2->n > 0 && (!pts2 || !desc2))) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	if (!(log_scale_factor1 > 0.f) || !(log_scale_factor2 > 0.f)) return orbx_fail(ORBX_ERR_INVALID, "logScaleFactor must be positive");
	// S21 = S12.Inverse() = Sim3(R^T, -(1/s) * R^T * t, 1/s) (include/Sim3.h:42-47); Map(x) = (s * R) * x + t
	float R21[9], t21[3];
	const float is12 = 1.f / S12->s, nis = -is12;
	for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R21[r * 3 + c] = S12->R[c * 3 + r];
	for (int i = 0; i < 3; i++) { float s = 0.f; for (int k = 0; k < 3; k++) s += (R21[i * 3 + k] * nis) * S12->t[k]; t21[i] = s; }
	float sR12[9], sR21[9];
	for (int k = 0; k < 9; k++) { sR12[k] = S12->R[k] * S12->s; sR21[k] = R21[k] * is12; }
	auto direction = [&](orbx_frame_s* from, const orbx_pose* pfrom, const orbx_keyframe_point* pts, const uint8_t* desc, const float* sR, const float* t,
	                     orbx_frame_s* to, const orbx_camera* cam_to, float lsf_to, std::vector<int32_t>& match) -> orbx_status {
		const int n = from->n;
		std::vector<BestWindow> win((size_t)n);
		for (int i = 0; i < n; i++)
		{
			BestWindow& w = win[i];
			w = BestWindow{ 0.f, 0.f, 0.f, 0.f, 0, 0, 0, 0 };
			if (!(pts[i].flags & 1)) continue;                          // null, already matched or bad (:1130, :1197)
			float xa[3], xb[3];
			matx_mul(pfrom->R, pts[i].xw, xa);
			for (int k = 0; k < 3; k++) xa[k] += pfrom->t[k];           // proj.WorldToCamera
			matx_mul(sR, xa, xb);
			for (int k = 0; k < 3; k++) xb[k] += t[k];                  // S.Map
			if (xb[2] < 0.f) continue;
			const float invZ = 1.f / xb[2];
			const float u = invZ * cam_to->fx * xb[0] + cam_to->cx, v = invZ * cam_to->fy * xb[1] + cam_to->cy;
			if (!in_image(to, u, v)) continue;
			const float maxDistance = 1.2f * pts[i].max_distance, minDistance = 0.8f * pts[i].min_distance;
			const float dist3D = norm3(xb);
			if (dist3D < minDistance || dist3D > maxDistance) continue;
			const int ps = predict_scale(pts[i].max_distance, dist3D, lsf_to, to->nlevels);
			w.u = u; w.v = v; w.radius = th * to->sf[ps];
			w.min_level = ps - 1; w.max_level = ps; w.flags = 1;
		}
		std::vector<int32_t> dist((size_t)n);
		match.assign((size_t)n, -1);
		if (orbx_status s = best_in_windows(to, win, desc, nullptr, match.data(), dist.data())) return s;
		for (int i = 0; i < n; i++)
			if (!(dist[i] <= TH_HIGH)) match[i] = -1;                   // :1187, :1254
		return ORBX_OK;
	};
	std::vector<int32_t> m1, m2;
	if (orbx_status s = direction(f1, pose1, pts1, desc1, sR21, t21, f2, cam2, log_scale_factor2, m1)) return s;
	if (orbx_status s = direction(f2, pose2, pts2, desc2, sR12, S12->t, f1, cam1, log_scale_factor1, m2)) return s;
	int found = 0;
	for (int i1 = 0; i1 < f1->n; i1++)                                  // agreement, :1260-1274
	{
		matches12[i1] = -1;
		const int idx2 = m1[i1];
		if (idx2 >= 0 && m2[idx2] == i1) { matches12[i1] = idx2; found++; }
	}
	if (match1) memcpy(match1, m1.data(), (size_t)f1->n * 4);
	if (match2) memcpy(match2, m2.data(), (size_t)f2->n * 4);
	if (nfound) *nfound = found;
	return ORBX_OK;
}

orbx_status orbx_search_for_triangulation(orbx_frame f1, const orbx_feature_vector* fv1, const uint8_t* has_mp1, orbx_frame f2,
                                          const orbx_feature_vector* fv2, const uint8_t* has_mp2, const float* F12, const float* epipole2,
                                          const float* sigma_sq2, int only_stereo, int check_orientation, int32_t* matches12, int* nmatches)
{
	if (!f1 || !f2 || !fv1 || !fv2 || !F12 || !epipole2 || !sigma_sq2 || !matches12) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	if ((f1->n > 0 && !has_mp1) || (f2->n > 0 && !has_mp2)) return orbx_fail(ORBX_ERR_INVALID, "bad argument");
	GCU(cudaSetDevice(f1->device));
	// FeatureVectorIterator (:406-450): nodes present in both vectors, ascending; items in the order the reference visits idx1
	std::vector<int> items;
	for (int a = 0, b = 0; a < fv1->nnodes && b < fv2->nnodes;)
	{
		if (fv1->node_ids[a] == fv2->node_ids[b])
		{
			for (int j = fv1->start[a]; j < fv1->start[a + 1]; j++)
			{
				const uint32_t idx1 = fv1->indices[j];
				if (idx1 >= (uint32_t)f1->n) return orbx_fail(ORBX_ERR_INVALID, "feature index outside key frame 1");
				if (has_mp1[idx1]) continue;                            // :796-797
				items.push_back((int)idx1); items.push_back(fv2->start[b]); items.push_back(fv2->start[b + 1] - fv2->start[b]);
			}
			a++; b++;
		}
		else if (fv1->node_ids[a] < fv2->node_ids[b]) a++;
		else b++;
	}
	const int nitems = (int)items.size() / 3, n2items = fv2->nnodes ? fv2->start[fv2->nnodes] : 0;
	for (int j = 0; j < n2items; j++)
		if (fv2->indices[j] >= (uint32_t)f2->n) return orbx_fail(ORBX_ERR_INVALID, "feature index outside key frame 2");
	for (int i = 0; i < f1->n; i++) matches12[i] = -1;
	if (nmatches) *nmatches = 0;
	if (nitems == 0) return ORBX_OK;
	const size_t o_idx2 = up16((size_t)nitems * 12), o_mp2 = o_idx2 + up16((size_t)n2items * 4), in_bytes = o_mp2 + up16((size_t)f2->n);
	GCU(f1->ensure_staging(in_bytes, (size_t)nitems * 4));
	memcpy(f1->h_in, items.data(), (size_t)nitems * 12);
	memcpy(f1->h_in + o_idx2, fv2->indices, (size_t)n2items * 4);
	memcpy(f1->h_in + o_mp2, has_mp2, (size_t)f2->n);
	TriArgs A;
	A.kps1 = f1->kps.p; A.desc1 = f1->desc.p; A.uright1 = f1->uright.p;
	A.kps2 = f2->kps.p; A.desc2 = f2->desc.p; A.uright2 = f2->uright.p;
	A.item = reinterpret_cast<const int*>(f1->d_in.p); A.idx2 = reinterpret_cast<const uint32_t*>(f1->d_in.p + o_idx2); A.has_mp2 = f1->d_in.p + o_mp2;
	A.nitems = nitems; A.only_stereo = only_stereo != 0;
	for (int k = 0; k < 9; k++) A.F[k] = F12[k];
	A.ex = epipole2[0]; A.ey = epipole2[1];
	for (int l = 0; l < 16; l++) { A.sf2[l] = l < f2->nlevels ? f2->sf[l] : 0.f; A.sigma_sq2[l] = l < f2->nlevels ? sigma_sq2[l] : 0.f; }
	A.match = reinterpret_cast<int*>(f1->d_out.p);
	GCU(cudaMemcpyAsync(f1->d_in.p, f1->h_in, in_bytes, cudaMemcpyHostToDevice, f1->st));
	GCU(cudaEventRecord(f1->ev0, f1->st));
	k_triangulation_search<<<(nitems + 127) / 128, 128, 0, f1->st>>>(A);
	GCU(cudaGetLastError());
	GCU(cudaEventRecord(f1->ev1, f1->st));
	GCU(cudaMemcpyAsync(f1->h_out, f1->d_out.p, (size_t)nitems * 4, cudaMemcpyDeviceToHost, f1->st));
	GCU(cudaStreamSynchronize(f1->st));
	cudaEventElapsedTime(&f1->last_kernel_ms, f1->ev0, f1->ev1);
	const int* res = reinterpret_cast<const int*>(f1->h_out);
	int n = 0;
	std::vector<std::pair<int, int>> tmp;                               // (bestIdx2, idx1) in visiting order, :847-849
	for (int t = 0; t < nitems; t++)
	{
		const int idx1 = items[3 * t];
		if (res[t] < 0) continue;
		matches12[idx1] = res[t];                                       // a keypoint listed under one node only: FeatureVector keys are disjoint
		tmp.emplace_back(res[t], idx1);
		n++;
	}
	if (check_orientation)
	{
		// CheckOrientation(keyframe2->keypointsUn, keyframe1->keypointsUn, tmpMatchIds, matches12) (:853-854, :249-309) on the host copy of
		// the angles: 30-bin histogram of angle2 - angle1, the bins past the three largest (after std::sort by size) are erased
		for (orbx_frame_s* f : { f1, f2 })
			if ((int)f->h_angle.size() != f->n)      // a frame assigned from device pointers: fetch the angles once
			{
				std::vector<orbx_keypoint> k((size_t)f->n);
				GCU(cudaMemcpy(k.data(), f->kps.p, (size_t)f->n * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost));
				f->h_angle.resize((size_t)f->n);
				for (int i = 0; i < f->n; i++) f->h_angle[i] = k[i].angle;
			}
		const std::vector<float>& ang1 = f1->h_angle; const std::vector<float>& ang2 = f2->h_angle;
		n = orbx_check_orientation_host(ang2.data(), ang1.data(), tmp, matches12, n);
	}
	if (nmatches) *nmatches = n;
	return ORBX_OK;
}

}  // extern "C"
