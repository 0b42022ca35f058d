// Host side of the C ABI (include/orbx.h): plan construction (all float/double-derived tables are computed here
// with the reference's exact operation order; kernels are integer-only on pixels), buffer management, launches.
// No CPU compute path exists: without a usable sm_100 device every entry point fails with ORBX_ERR_CUDA.
#include "orbx_internal.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

int orbx_knn2_splits(int64_t nq, int64_t nt);

namespace {

thread_local std::string g_err;

orbx_status fail(orbx_status s, const std::string& msg)
{
	g_err = msg;
	return s;
}

#define CU(call)                                                                                         \
	do {                                                                                                 \
		cudaError_t e_ = (call);                                                                         \
		if (e_ != cudaSuccess)                                                                           \
			return fail(ORBX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
	} while (0)

inline int cv_round(double v) { return (int)lrint(v); }
inline int cv_round(float v) { return (int)lrintf(v); }
inline short sat_s16(int v) { return (short)std::min(32767, std::max(-32768, v)); }
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// cv::remap's own conversion of float maps to fixed point (modules/imgproc/src/imgwarp.cpp, INTER_BITS = 5): sx = cvRound(mapx * 32),
// integer part saturate_cast<short>(sx >> 5), fraction sx & 31
void pack_remap_table(const float* map1, const float* map2, size_t map_pitch, int w, int h, std::vector<int2>& tab)
{
	tab.resize((size_t)w * h);
	for (int y = 0; y < h; y++)
	{
		const float* mx = reinterpret_cast<const float*>(reinterpret_cast<const char*>(map1) + (size_t)y * map_pitch);
		const float* my = reinterpret_cast<const float*>(reinterpret_cast<const char*>(map2) + (size_t)y * map_pitch);
		for (int x = 0; x < w; x++)
		{
			const int sx = cv_round(mx[x] * 32), sy = cv_round(my[x] * 32);
			const int ix = sat_s16(sx >> 5), iy = sat_s16(sy >> 5);
			tab[(size_t)y * w + x] = make_int2((ix & 0xffff) | (int)((unsigned)iy << 16), (sx & 31) | ((sy & 31) << 5));
		}
	}
}

template <class T> struct DevBuf
{
	T* p = nullptr;
	size_t n = 0;
	cudaError_t ensure(size_t count)
	{
		if (count <= n) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; n = 0;
		cudaError_t e = cudaMalloc(&p, count * sizeof(T));
		if (e == cudaSuccess) n = count;
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
	~DevBuf() { release(); }
	DevBuf() = default;
	DevBuf(const DevBuf&) = delete;
	DevBuf& operator=(const DevBuf&) = delete;
};

bool device_ok(int device, std::string& why)
{
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess) { why = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e); return false; }
	if (device < 0 || device >= count) { why = "no such CUDA device"; return false; }
	cudaDeviceProp p;
	e = cudaGetDeviceProperties(&p, device);
	if (e != cudaSuccess) { why = cudaGetErrorString(e); return false; }
	if (p.major != 10) { why = "device is not compute capability 10.x (kernels are built for sm_100a only)"; return false; }
	return true;
}

}  // namespace

orbx_status orbx_fail(orbx_status s, const char* msg) { return fail(s, msg); }
bool orbx_device_usable(int device, const char** why)
{
	static thread_local std::string text;
	const bool ok = device_ok(device, text);
	*why = text.c_str();
	return ok;
}

struct orbx_extractor
{
	orbx_params prm;
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t stream2 = nullptr;     // second lane of the chunk pipeline of the host-buffer API
	cudaStream_t copy_in = nullptr;     // uploads of the chunk pipeline: every chunk has its own region of the level-0 buffer, so the uploads of a call
	std::vector<cudaEvent_t> ev_in;     // queue back to back here and the host-to-device engine never waits for a lane's kernels; one event per chunk
	cudaEvent_t done = nullptr;
	cudaEvent_t fork = nullptr, join = nullptr;   // order the second lane inside the handle's stream for the device-resident API
	cudaStream_t side[2] = { nullptr, nullptr };  // per lane: the blur runs here, beside the quadtree (latency-bound, leaves the SMs half empty)
	cudaEvent_t side_fork[2] = { nullptr, nullptr }, side_join[2] = { nullptr, nullptr };
	// a frame at a time (latency path): the levels run as three groups on their own streams, see enqueue_extract
	cudaStream_t grp[2][2] = { { nullptr, nullptr }, { nullptr, nullptr } };
	cudaEvent_t grp_ev[2][4] = { { nullptr, nullptr, nullptr, nullptr }, { nullptr, nullptr, nullptr, nullptr } };   // pyramid of group B done, pyramid done, group B done, group C done
	std::vector<float> scale, inv_scale, sigma_sq, inv_sigma_sq;
	std::vector<int> quota;

	// plan (depends on image size)
	int pw = 0, ph = 0, frames_cap = 0;
	OrbxPlanDev P;
	OrbxTmaMaps maps;
	OrbxStripMaps smaps[3];             // strip kernels (blur, dense FAST bound): every level with the strip box; [1]: the 8-row tiles of small batches, [2]: the blur's 64-row tiles
	OrbxPyrMaps pmaps[2];               // strip resize kernel: level s - 1 with the source box of a tile of level s
	DevBuf<uint8_t> fmap_ini, fmap_min; // FAST bound bitmaps (1 bit per level pixel each)
	const uint8_t* l0_map_base = nullptr;   // what the level-0 descriptors currently point at
	int64_t l0_map_pitch = 0, l0_map_stride = 0;
	int l0_map_frames = 0;
	DevBuf<uint8_t> color;              // interleaved colour frames / unrectified frames of the current batch (orbx_extract_batch_color / _rectified)
	DevBuf<int2> rect_tab;              // rectification table (orbx_set_rectification): per output pixel (ix | iy << 16, fx | fy << 5)
	int rect_w = 0, rect_h = 0, rect_sw = 0, rect_sh = 0;
	DevBuf<uint8_t> pyr, blur, l0buf;   // l0buf: level 0 of every frame, back to back (host-buffer API uploads land here)
	int64_t l0_pitch = 0, l0_stride = 0;
	uint8_t* l0base = nullptr;          // l0buf.p + 256: kernels may read up to 16 bytes in front of a row (aligned 16-byte tile copies)
	DevBuf<uint32_t> cand, qbuf0, qbuf1, sel;
	DevBuf<int> cell_count, cand_count, sel_count, pyr_done, ovf_count;
	DevBuf<uint32_t> ovf_list;          // cells left to the overflow launch of the FAST cell kernel, one region per frame range
	DevBuf<int> root_x, xofs, yofs;
	DevBuf<uint8_t> root_lut;
	DevBuf<short2> xcoef, ycoef;
	DevBuf<int4> cell_tab;
	// outputs owned by the handle (host-buffer API) and the state of the last extract (for stereo / probes)
	DevBuf<orbx_keypoint> out_kps;
	DevBuf<uint8_t> out_desc;
	DevBuf<int32_t> out_n;
	int32_t* h_counts = nullptr;        // pinned staging for the per-frame counts (a pageable target would serialise the pipeline)
	size_t h_counts_n = 0;
	// a frame at a time through the host-buffer API: the grouped launches + the three result copies of enqueue_extract as ONE graph launch.
	// Every pointer in it is the handle's own (level-0 buffer, result buffers, pinned staging); the key holds what it was captured with.
	struct SmallKey { uint64_t gen; const void* l0; const void* kps; const void* desc; const void* n; const void* hs; const void* counts; int frames, ocap; };
	SmallKey small_key = {};
	cudaGraphExec_t small_exec = nullptr;
	unsigned long long* trace = nullptr;   // ORBX_TRACE=1: 16 device slots of %globaltimer stamps along a one-frame call (debug)
	uint64_t gen = 0;                   // bumped whenever a tensor map is (re-)encoded or the plan is rebuilt
	uint8_t* h_small = nullptr;         // pinned staging for the keypoints + descriptors of a small batch (a frame at a time): three queued copies and
	size_t h_small_bytes = 0;           // one synchronisation instead of two blocking copies into the caller's pageable arrays
	DevBuf<float> st_uright, st_depth;
	DevBuf<int> st_sad, st_rows;
	DevBuf<uint2> st_items;
	bool stage_timing = false;
	std::vector<cudaEvent_t> ev_pool;   // 10 events per timed extract call
	size_t ev_used = 0;
	bool have_result = false;
	int last_frames = 0, last_cap = 0;
	const orbx_keypoint* last_kps = nullptr;
	const uint8_t* last_desc = nullptr;
	const int32_t* last_n = nullptr;
};

namespace {

// ORBextractor::Init (src/ORBextractor.cc:720-740) and ComputeNumFeaturesPerScale (:472-487)
void build_tables(orbx_extractor* h)
{
	const int nl = h->prm.nlevels;
	h->scale.resize(nl); h->inv_scale.resize(nl); h->sigma_sq.resize(nl); h->inv_sigma_sq.resize(nl); h->quota.resize(nl);
	float s = 1.f;
	for (int i = 0; i < nl; i++)
	{
		h->scale[i] = s;
		h->inv_scale[i] = 1.f / s;
		h->sigma_sq[i] = s * s;
		h->inv_sigma_sq[i] = 1.f / (s * s);
		s *= h->prm.scale_factor;
	}
	const double factor = 1 / h->prm.scale_factor;
	double nf = h->prm.nfeatures * (1 - factor) / (1 - std::pow(factor, nl));
	int sum = 0;
	for (int i = 0; i < nl - 1; i++)
	{
		h->quota[i] = cv_round(nf);
		sum += h->quota[i];
		nf *= factor;
	}
	h->quota[nl - 1] = std::max(h->prm.nfeatures - sum, 0);
}

// cv::resize coefficient tables
void resize_tables(int dn, int sn, int* ofs, short2* coef)
{
	const double scale = 1.0 / ((double)dn / sn);
	for (int d = 0; d < dn; d++)
	{
		float f = (float)((d + 0.5) * scale - 0.5);
		int s = (int)std::floor(f);
		f -= (float)s;
		if (s < 0) { s = 0; f = 0.f; }
		if (s + 1 >= sn) { s = sn - 1; f = 0.f; }
		ofs[d] = s;
		coef[d].x = sat_s16(cv_round((1.f - f) * 2048.f));
		coef[d].y = sat_s16(cv_round(f * 2048.f));
	}
}

// TMA descriptors of level s seen as a (dimx, h, frames) u8 tensor with row pitch `pitch` and frame stride `stride` at `base`: the cell
// view box (FAST cell kernel), the strip boxes (blur, dense FAST bound; throughput and small-batch tile heights) and, as the SOURCE of level
// s + 1, that level's resize boxes. base must be 16-byte aligned, pitch and stride multiples of 16 (cuTensorMapEncodeTiled checks).
orbx_status encode_level_maps(orbx_extractor* h, int s, const void* base, int64_t dimx, int64_t pitch, int64_t stride, int frames)
{
	typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
	                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static const EncodeFn fn = []() -> EncodeFn {          // resolved once (thread-safe static initialisation)
		void* p = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
		return (EncodeFn)p;
	}();
	if (!fn) return fail(ORBX_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
	h->gen++;
	const OrbxPlanDev& P = h->P;
	const OrbxLevel& L = P.lv[s];
	const cuuint64_t dims[3] = { (cuuint64_t)dimx, (cuuint64_t)L.h, (cuuint64_t)frames };
	const cuuint64_t strides[2] = { (cuuint64_t)pitch, (cuuint64_t)stride };
	const cuuint32_t estr[3] = { 1, 1, 1 };
	auto enc = [&](CUtensorMap* m, int bw, int bh, const char* what) -> orbx_status {
		const cuuint32_t box[3] = { (cuuint32_t)bw, (cuuint32_t)bh, 1 };
		const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
		                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS) return fail(ORBX_ERR_CUDA, std::string("cuTensorMapEncodeTiled (") + what + ") failed with CUresult " + std::to_string((int)r));
		return ORBX_OK;
	};
	orbx_status st = enc(&h->maps.level[s], orbx_fast_tile_stride(), h->maps.box_h[s], "cell view");
	if (st == ORBX_OK) st = enc(&h->smaps[2].level[s], orbx_strip_box_w(), orbx_strip_rows(2) + 6, "blur strip box");
	for (int which = 0; which < 2 && st == ORBX_OK; which++)
	{
		st = enc(&h->smaps[which].level[s], orbx_strip_box_w(), orbx_strip_rows(which) + 6, "strip box");
		if (st == ORBX_OK && s + 1 < P.nlevels && P.lv[s + 1].py_bw[which] > 0)
			st = enc(&h->pmaps[which].src[s + 1], P.lv[s + 1].py_bw[which], P.lv[s + 1].py_bh[which], "resize box");
	}
	return st;
}

// level 0 lives either in the handle's padded buffer or, for the device-resident API, in the caller's own buffer (no repack): its
// descriptors follow whichever the current call uses
orbx_status ensure_level0_maps(orbx_extractor* h, const uint8_t* base, int64_t dimx, int64_t pitch, int64_t stride, int frames)
{
	if (h->l0_map_base == base && h->l0_map_pitch == pitch && h->l0_map_stride == stride && h->l0_map_frames >= frames) return ORBX_OK;
	const orbx_status st = encode_level_maps(h, 0, base, dimx, pitch, stride, frames);
	if (st != ORBX_OK) { h->l0_map_base = nullptr; return st; }
	h->l0_map_base = base; h->l0_map_pitch = pitch; h->l0_map_stride = stride; h->l0_map_frames = frames;
	return ORBX_OK;
}

orbx_status build_plan(orbx_extractor* h, int w, int hgt, int frames)
{
	const int nl = h->prm.nlevels;
	if (w == h->pw && hgt == h->ph && frames <= h->frames_cap)
	{
		h->P.frames = frames;
		return ORBX_OK;
	}
	// from here on the old plan is gone: a failed rebuild must not leave a size cache that matches a zeroed plan
	h->pw = h->ph = h->frames_cap = 0;
	h->have_result = false;
	OrbxPlanDev& P = h->P;
	std::memset(&P, 0, sizeof(P));
	if (w > 4096 || hgt > 4096)
		return fail(ORBX_ERR_INVALID, "image larger than 4096 px per side (candidate packing limit)");

	P.nlevels = nl; P.frames = frames;
	P.ini_th = h->prm.ini_th_fast; P.min_th = h->prm.min_th_fast;

	std::vector<int> root_x, xofs, yofs;
	std::vector<uint8_t> root_lut;
	std::vector<short2> xcoef, ycoef;
	std::vector<int4> cell_tab;
	int64_t slab = 0;
	int cells = 0, cands = 0, sels = 0, node_cap = 0;
	for (int s = 0; s < nl; s++)
	{
		OrbxLevel& L = P.lv[s];
		// ComputePyramid (:455-470): sizes always from the original dims
		L.w = s == 0 ? w : cv_round(h->inv_scale[s] * (float)w);
		L.h = s == 0 ? hgt : cv_round(h->inv_scale[s] * (float)hgt);
		L.pitch = (int)align_up(L.w, 128);
		L.offset = slab;
		slab += (int64_t)L.pitch * L.h;
		slab = align_up(slab, 256);
		L.scale = h->scale[s];

		// DetectFAST grid (:489-523)
		const int rw = L.w - 2 * ORBX_BORDER, rh = L.h - 2 * ORBX_BORDER;
		if (rw < ORBX_CELL || rh < ORBX_CELL)
			return fail(ORBX_ERR_INVALID, "a pyramid level is smaller than 62 px: the reference divides by zero (src/ORBextractor.cc:508-511)");
		L.minx = ORBX_BORDER; L.miny = ORBX_BORDER; L.maxx = L.minx + rw; L.maxy = L.miny + rh;
		const int gridw = rw / ORBX_CELL, gridh = rh / ORBX_CELL;
		L.cellw = (int)std::ceil(1. * rw / gridw);
		L.cellh = (int)std::ceil(1. * rh / gridh);
		L.ncx = 0;
		for (int cx = 0, x0 = L.minx; cx < gridw && x0 + 6 < L.maxx; cx++, x0 += L.cellw) L.ncx++;
		L.ncy = 0;
		for (int cy = 0, y0 = L.miny; cy < gridh && y0 + 6 < L.maxy; cy++, y0 += L.cellh) L.ncy++;
		L.cell_base = cells;
		cells += L.ncx * L.ncy;
		for (int cy = 0; cy < L.ncy; cy++)
			for (int cx = 0; cx < L.ncx; cx++)
			{
				// cell view [x0, x1) x [y0, y1) (:521-524)
				const int x0 = L.minx + cx * L.cellw, y0 = L.miny + cy * L.cellh;
				const int x1 = std::min(x0 + L.cellw + 6, L.maxx), y1 = std::min(y0 + L.cellh + 6, L.maxy);
				cell_tab.push_back(make_int4(x0 | (y0 << 16), (x1 - x0) | ((y1 - y0) << 16), s, cy * L.ncx + cx));
			}
		L.cell_cap = ((L.cellw + 1) / 2) * ((L.cellh + 1) / 2);
		L.cand_base = cands;
		L.cand_cap = L.ncx * L.ncy * L.cell_cap;
		cands += L.cand_cap;

		// QuadTreeSuppression roots (:547-581)
		const int n0 = cv_round(1. * rw / rh);
		if (n0 < 1 || n0 > ORBX_MAX_ROOTS)
			return fail(ORBX_ERR_INVALID, "aspect ratio outside [1, 16]: the reference divides by zero for portrait images (src/ORBextractor.cc:547-548)");
		const double hx = 1. * rw / n0;
		L.n_roots = n0;
		L.root_base = (int)root_x.size();
		for (int i = 0; i <= n0; i++) root_x.push_back((int)(L.minx + hx * i));
		L.rootlut_base = (int)root_lut.size();
		for (int x = 0; x < L.w; x++)
		{
			int id = (int)(((float)x - L.minx) / hx);
			root_lut.push_back((uint8_t)std::min(std::max(id, 0), n0 - 1));
		}
		L.quota = h->quota[s];
		L.sel_cap = std::max(L.quota + 3, 4 * n0);
		L.sel_base = sels;
		sels += L.sel_cap;
		node_cap = std::max(node_cap, L.sel_cap);

		if (s > 0)
		{
			L.xtab_base = (int)xofs.size();
			xofs.resize(xofs.size() + L.w); xcoef.resize(xcoef.size() + L.w);
			resize_tables(L.w, P.lv[s - 1].w, xofs.data() + L.xtab_base, xcoef.data() + L.xtab_base);
			L.ytab_base = (int)yofs.size();
			yofs.resize(yofs.size() + L.h); ycoef.resize(ycoef.size() + L.h);
			resize_tables(L.h, P.lv[s - 1].h, yofs.data() + L.ytab_base, ycoef.data() + L.ytab_base);
			// the resize kernel stages the source rows/columns of one output tile (128 x 128) in shared memory
			const int th = orbx_pyramid_tile_rows(), tw = orbx_pyramid_tile_cols();
			int max_rows = 0, max_bytes = 0;
			for (int y0 = 0; y0 < L.h; y0 += th)
			{
				const int y1 = std::min(y0 + th, L.h) - 1;
				max_rows = std::max(max_rows, std::min(yofs[L.ytab_base + y1] + 1, P.lv[s - 1].h - 1) - yofs[L.ytab_base + y0] + 1);
			}
			for (int x0 = 0; x0 < L.w; x0 += tw)
			{
				const int x1 = std::min(x0 + tw, L.w) - 1;
				const int xa = xofs[L.xtab_base + x0] & ~15, xb = std::min(xofs[L.xtab_base + x1] + 1, P.lv[s - 1].w - 1);
				max_bytes = std::max(max_bytes, ((xb - xa + 16) >> 4) * 16);
			}
			if (max_rows > orbx_pyramid_max_src_rows() || max_bytes > orbx_pyramid_max_src_bytes())
				return fail(ORBX_ERR_INVALID, "scaleFactor too large for the resize kernel (max about 2.2)");
			L.py_smem = max_rows * max_bytes;
			// strip resize kernel (one warp per 128 x TH tile, source rectangle by one TMA box): box rows/bytes over all tiles. A lane reads
			// three aligned words from its first source column on and picks its four (s[x], s[x+1]) pairs out of 8 bytes.
			for (int which = 0; which < 2; which++)
			{
				const int sth = orbx_pyramid_strip_rows(which);
				int brows = 0, bbytes = 0, emax = 0;
				for (int y0 = 0; y0 < L.h; y0 += sth)
				{
					const int y1 = std::min(y0 + sth, L.h) - 1;
					brows = std::max(brows, std::min(yofs[L.ytab_base + y1] + 1, P.lv[s - 1].h - 1) - yofs[L.ytab_base + y0] + 1);
				}
				for (int x0 = 0; x0 < L.w; x0 += tw)
				{
					const int xa = xofs[L.xtab_base + x0] & ~15;
					for (int x = x0; x < std::min(x0 + tw, L.w); x += 4)
					{
						const int xl = std::min(x + 3, L.w - 1);
						emax = std::max(emax, xofs[L.xtab_base + xl] - xofs[L.xtab_base + x]);
						bbytes = std::max(bbytes, ((xofs[L.xtab_base + x] - xa) & ~3) + 12);
					}
				}
				bbytes = (int)align_up(bbytes, 16);
				if (emax <= 6 && bbytes <= 256 && brows <= 256 && bbytes * brows <= 96 * 1024) { L.py_bw[which] = bbytes; L.py_bh[which] = brows; }
				else { L.py_bw[which] = 0; L.py_bh[which] = 0; }      // scale factor too large for one TMA box: the cp.async kernel produces this level
			}
		}
	}
	P.slab = slab;
	P.cells_per_frame = cells; P.cand_per_frame = cands; P.sel_per_frame = sels;
	P.node_cap = node_cap + 8;
	P.out_cap = sels;
	if (sels >= 65536)
		return fail(ORBX_ERR_INVALID, "more than 65535 keypoints per frame");
	if (orbx_quadtree_smem(P.node_cap, true) > 200 * 1024)
		return fail(ORBX_ERR_INVALID, "nfeatures too large for the quadtree kernel's shared memory");

	CU(cudaSetDevice(h->device));
	const size_t F = (size_t)frames;
	CU(h->pyr.ensure(F * slab + 512)); CU(h->blur.ensure(F * slab + 512));
	h->l0_pitch = P.lv[0].pitch; h->l0_stride = (int64_t)P.lv[0].pitch * P.lv[0].h;   // frames back to back: one strided copy uploads a whole chunk
	CU(h->l0buf.ensure(F * h->l0_stride + 512));
	h->l0base = h->l0buf.p + 256;
	CU(h->cand.ensure(F * cands)); CU(h->qbuf0.ensure(F * cands)); CU(h->qbuf1.ensure(F * cands));
	CU(h->fmap_ini.ensure(F * (slab >> 3) + 64)); CU(h->fmap_min.ensure(F * (slab >> 3) + 64));
	CU(h->cell_count.ensure(2 * F * cells));      // counts, then offsets
	CU(h->cand_count.ensure(F * nl)); CU(h->sel_count.ensure(F * nl)); CU(h->pyr_done.ensure(F * ORBX_MAX_LEVELS));
	CU(h->ovf_list.ensure(F * cells)); CU(h->ovf_count.ensure(2));
	CU(h->sel.ensure(F * sels));
	CU(h->root_x.ensure(root_x.size())); CU(h->root_lut.ensure(root_lut.size()));
	CU(h->xofs.ensure(std::max<size_t>(xofs.size(), 1))); CU(h->xcoef.ensure(std::max<size_t>(xcoef.size(), 1)));
	CU(h->yofs.ensure(std::max<size_t>(yofs.size(), 1))); CU(h->ycoef.ensure(std::max<size_t>(ycoef.size(), 1)));
	CU(h->cell_tab.ensure(cell_tab.size()));
	CU(cudaMemcpyAsync(h->cell_tab.p, cell_tab.data(), cell_tab.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
	CU(h->out_kps.ensure(F * sels)); CU(h->out_desc.ensure(F * sels * 32)); CU(h->out_n.ensure(F));
	CU(cudaMemcpyAsync(h->root_x.p, root_x.data(), root_x.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync(h->root_lut.p, root_lut.data(), root_lut.size(), cudaMemcpyHostToDevice, h->stream));
	if (!xofs.empty())
	{
		CU(cudaMemcpyAsync(h->xofs.p, xofs.data(), xofs.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
		CU(cudaMemcpyAsync(h->xcoef.p, xcoef.data(), xcoef.size() * sizeof(short2), cudaMemcpyHostToDevice, h->stream));
		CU(cudaMemcpyAsync(h->yofs.p, yofs.data(), yofs.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
		CU(cudaMemcpyAsync(h->ycoef.p, ycoef.data(), ycoef.size() * sizeof(short2), cudaMemcpyHostToDevice, h->stream));
	}
	CU(cudaStreamSynchronize(h->stream));    // the host vectors above die with this scope

	P.pyr = h->pyr.p + 256; P.blur = h->blur.p;
	P.fmap_ini = h->fmap_ini.p; P.fmap_min = h->fmap_min.p;
	P.cand = h->cand.p; P.cell_count = h->cell_count.p; P.qbuf0 = h->qbuf0.p; P.qbuf1 = h->qbuf1.p;
	P.cand_count = h->cand_count.p; P.sel = h->sel.p; P.sel_count = h->sel_count.p; P.pyr_done = h->pyr_done.p;
	P.ovf_list = h->ovf_list.p; P.ovf_count = h->ovf_count.p;
	P.root_x = h->root_x.p; P.root_lut = h->root_lut.p; P.cell_tab = h->cell_tab.p;
	P.xofs = h->xofs.p; P.xcoef = h->xcoef.p; P.yofs = h->yofs.p; P.ycoef = h->ycoef.p;
	// TMA descriptors of every level as a (pitch, h, frames) u8 tensor, one per box shape (encode_level_maps)
	{
		std::memset(&h->maps, 0, sizeof(h->maps));
		int max_view_w = 0;
		for (int s = 0; s < nl; s++)
			for (int cx = 0, x0 = P.lv[s].minx; cx < P.lv[s].ncx; cx++, x0 += P.lv[s].cellw)
				max_view_w = std::max(max_view_w, std::min(x0 + P.lv[s].cellw + 6, P.lv[s].maxx) - x0);
		if (max_view_w + 15 > orbx_fast_tile_stride()) return fail(ORBX_ERR_INVALID, "cell wider than the FAST tile");
		std::memset(h->smaps, 0, sizeof(h->smaps));
		std::memset(h->pmaps, 0, sizeof(h->pmaps));
		for (int s = 0; s < nl; s++)
		{
			const OrbxLevel& L = P.lv[s];
			int box_h = 0;
			for (int cy = 0, y0 = L.miny; cy < L.ncy; cy++, y0 += L.cellh) box_h = std::max(box_h, std::min(y0 + L.cellh + 6, L.maxy) - y0);
			if (box_h > orbx_fast_tile_rows()) return fail(ORBX_ERR_INVALID, "cell taller than the FAST tile");
			h->maps.box_h[s] = box_h;
			const orbx_status st = s == 0 ? encode_level_maps(h, 0, h->l0base, L.pitch, L.pitch, h->l0_stride, frames)
			                              : encode_level_maps(h, s, P.pyr + L.offset, L.pitch, L.pitch, P.slab, frames);
			if (st != ORBX_OK) return st;
		}
		h->l0_map_base = h->l0base; h->l0_map_pitch = h->l0_pitch; h->l0_map_stride = h->l0_stride; h->l0_map_frames = frames;
	}
	h->pw = w; h->ph = hgt; h->frames_cap = frames;
	h->have_result = false;
	return ORBX_OK;
}

// enqueue the extraction of frames [fb, fb + fc) of the planned batch on stream st; level 0 of frame f is at
// l0 + f*l0_stride. All per-frame buffers are indexed by frame, so a chunk is the same launch with offset bases.
// paired: the call is one of the two lanes of a device-resident batch (see orbx_extract_batch_device) and takes that lane's stage order.
// grouped: a small batch whose launches are being captured into a graph: the levels run as three groups on parallel streams (below).
orbx_status enqueue_extract(orbx_extractor* h, int fb, int fc, cudaStream_t st, const uint8_t* l0, int64_t l0_pitch, int64_t l0_stride,
                            orbx_keypoint* d_kps, uint8_t* d_desc, int32_t* d_n, int cap, bool paired = false, bool grouped = false)
{
	OrbxPlanDev& P0 = h->P;
	P0.l0 = l0; P0.l0_pitch = l0_pitch; P0.l0_stride = l0_stride;
	P0.out_cap = cap;      // stride of the output arrays for this call (>= sel_per_frame)
	OrbxPlanDev P = P0;
	int* cell_off = P0.cell_count + (int64_t)h->frames_cap * P0.cells_per_frame + (int64_t)fb * P0.cells_per_frame;
	P.frames = fc;
	P.frame0 = fb;
	P.l0 += (int64_t)fb * l0_stride;
	P.pyr += (int64_t)fb * P.slab; P.blur += (int64_t)fb * P.slab;
	P.cand += (int64_t)fb * P.cand_per_frame; P.qbuf0 += (int64_t)fb * P.cand_per_frame; P.qbuf1 += (int64_t)fb * P.cand_per_frame;
	P.cell_count += (int64_t)fb * P.cells_per_frame;
	P.ovf_list += (int64_t)fb * P.cells_per_frame; P.ovf_count += st == h->stream2 ? 1 : 0;      // the two lanes run at the same time
	P.cand_count += (int64_t)fb * P.nlevels; P.sel_count += (int64_t)fb * P.nlevels;
	P.sel += (int64_t)fb * P.sel_per_frame;
	d_kps += (int64_t)fb * cap; d_desc += (int64_t)fb * cap * 32; d_n += fb;

	// 10 events per timed call: (start, end) of pyramid, FAST, quadtree, blur, describe on the stream each stage runs on
	cudaEvent_t* ev = nullptr;
	if (h->stage_timing)
	{
		while (h->ev_pool.size() < h->ev_used + 10)
		{
			cudaEvent_t e;
			CU(cudaEventCreate(&e));
			h->ev_pool.push_back(e);
		}
		ev = h->ev_pool.data() + h->ev_used;
		h->ev_used += 10;
	}
	// The blur on a side stream beside the quadtree was measured: +1.2 % device-resident (178.3 k vs 176.1 k frames/s), -3 % end to end
	// (147.6 k vs 151.9 k: two more streams per handle in the chunk pipeline), no change without the stage events. Default: in line.
	// Small batches (a frame per call) are latency-bound: the blur then runs as a parallel branch beside FAST + quadtree.
	static const bool blur_side_env = getenv("ORBX_BLUR_SIDE") != nullptr;
	const bool blur_inline = !(blur_side_env || (fc <= ORBX_SMALL_BATCH && !h->stage_timing));
	const int lane = st == h->stream2 ? 1 : 0;
	cudaStream_t side = blur_inline ? st : h->side[lane];
	// ---- a frame at a time: what the caller waits for is the longest dependent chain, and that chain is level 0's (most candidates, the
	// longest quadtree CTA). Level 0 needs nothing but the input, so FAST and the quadtree of level 0 start at once on the caller's stream
	// while the pyramid is still being built; levels 1-2 and levels 3.. follow as two more groups on their own streams as soon as their part
	// of the pyramid exists, the blur runs beside them, and the descriptor stage joins everything. Three chains of about equal length
	// instead of pyramid -> FAST -> quadtree of all levels in a row. Tuning knob ORBX_GROUPS=0: the single chain. Only where the launches are
	// replayed as a graph (the host-buffer API): issued call by call the 14 launches and 10 event operations are host-bound and slower than
	// the single chain (0.188 vs 0.173 ms per call), so orbx_extract_batch_device keeps the chain for small batches.
	static const bool groups_on = !(getenv("ORBX_GROUPS") && atoi(getenv("ORBX_GROUPS")) == 0);
	if (groups_on && grouped && !blur_inline && !ev && P.nlevels >= 4)
	{
		// first level of the third group (tuning knob ORBX_GROUP_C, 2 .. nlevels - 1)
		static const int group_c = getenv("ORBX_GROUP_C") ? atoi(getenv("ORBX_GROUP_C")) : 3;
		const int n = P.nlevels, sB = 1, sC = std::max(2, std::min(n - 1, group_c));
		cudaStream_t gB = h->grp[lane][0], gC = h->grp[lane][1];
		cudaEvent_t* ge = h->grp_ev[lane];
		static const bool trace_on = getenv("ORBX_TRACE") != nullptr;
		auto stamp = [&](int slot, cudaStream_t s_) { if (trace_on && h->trace) orbx_launch_stamp(h->trace + slot, s_); };
		stamp(0, st);
		CU(cudaEventRecord(h->side_fork[lane], st));                     // the input (and whatever the caller queued before) is in place
		// group B: levels 1-2
		CU(cudaStreamWaitEvent(gB, h->side_fork[lane], 0));
		CU(orbx_launch_pyramid_all(P, h->pmaps, gB, sB, sC));
		stamp(1, gB);
		CU(cudaEventRecord(ge[0], gB));
		// group A: level 0
		orbx_launch_fast(P, h->maps, h->smaps, st, 0, 0, 1);
		stamp(2, st);
		orbx_launch_quadtree(P, cell_off, st, 0, 1);
		stamp(3, st);
		// group C: levels 3 ..
		CU(cudaStreamWaitEvent(gC, ge[0], 0));
		CU(orbx_launch_pyramid_all(P, h->pmaps, gC, sC, n));
		stamp(4, gC);
		CU(cudaEventRecord(ge[1], gC));
		orbx_launch_fast(P, h->maps, h->smaps, gB, 0, sB, sC);
		stamp(5, gB);
		orbx_launch_quadtree(P, cell_off, gB, sB, sC);
		stamp(6, gB);
		CU(cudaEventRecord(ge[2], gB));
		orbx_launch_fast(P, h->maps, h->smaps, gC, 0, sC, n);
		stamp(7, gC);
		orbx_launch_quadtree(P, cell_off, gC, sC, n);
		stamp(8, gC);
		CU(cudaEventRecord(ge[3], gC));
		// blur of all levels beside the groups
		CU(cudaStreamWaitEvent(side, ge[1], 0));
		orbx_launch_blur(P, h->smaps, side);
		stamp(9, side);
		CU(cudaEventRecord(h->side_join[lane], side));
		CU(cudaStreamWaitEvent(st, ge[2], 0));
		CU(cudaStreamWaitEvent(st, ge[3], 0));
		CU(cudaStreamWaitEvent(st, h->side_join[lane], 0));
		orbx_launch_describe(P, d_kps, d_desc, d_n, st);
		stamp(10, st);
		CU(cudaGetLastError());
		return ORBX_OK;
	}
	// The two lanes of a batch are released by the same event, and which lane's first kernel the GPU takes first is then arbitrary (it
	// differed from one B200 box to the next, and decided whether a stage order helped or hurt). Two memsets of a counter the lane owns
	// put lane 1 a few microseconds behind lane 0, always. Tuning knob ORBX_LEAD: 0 = leave it to the hardware, 1 = lane 0 leads, 2 = lane 1.
	static const int lead = getenv("ORBX_LEAD") ? atoi(getenv("ORBX_LEAD")) : 1;
	if (paired && !ev && ((lead == 1 && lane == 1) || (lead == 2 && lane == 0)))
		for (int i = 0; i < 2; i++) CU(cudaMemsetAsync(P.ovf_count, 0, sizeof(int), st));
	if (ev) CU(cudaEventRecord(ev[0], st));
	CU(orbx_launch_pyramid_all(P, h->pmaps, st));
	if (ev) CU(cudaEventRecord(ev[1], st));
	// Where the in-line blur sits, one digit per lane (tuning knob ORBX_BLUR_POS): 0 = behind the pyramid, 1 = behind the dense FAST bound,
	// 2 = behind the FAST cell kernel, 3 = behind the quadtree. Nothing but the descriptor stage reads the blurred levels. With "20" the
	// leading lane runs pyramid, FAST, blur, quadtree and the other lane pyramid, blur, FAST, quadtree, so the lanes are never in the same
	// stage at the same time: the blur (dot-product pipe) runs beside the dense FAST bound (ALU pipe), the quadtree (latency) beside either.
	// 512 frames at C1, three B200 boxes: 2.058 -> 1.925 ms per step (248.8 k -> 266.0 k frames/s); tools/skew_probe.py has the grid.
	static const char* blur_pos_env = getenv("ORBX_BLUR_POS");
	const char* blur_pos_str = blur_pos_env && strlen(blur_pos_env) >= 2 ? blur_pos_env : "20";
	int blur_pos = 0;
	if (paired && blur_inline && !ev) blur_pos = std::max(0, std::min(3, blur_pos_str[lane] - '0'));
	if (blur_inline)
	{
		if (blur_pos == 0)
		{
			if (ev) CU(cudaEventRecord(ev[6], st));
			orbx_launch_blur(P, h->smaps, st);
			if (ev) CU(cudaEventRecord(ev[7], st));
		}
	}
	else
	{
		// The blur only depends on the pyramid and only the descriptor stage reads it: a parallel branch beside FAST + quadtree (the quadtree
		// leaves most of the GPU idle on a small batch; in a captured graph this fork / join becomes two independent paths)
		CU(cudaEventRecord(h->side_fork[lane], st));
		CU(cudaStreamWaitEvent(side, h->side_fork[lane], 0));
		if (ev) CU(cudaEventRecord(ev[6], side));
		orbx_launch_blur(P, h->smaps, side);
		if (ev) CU(cudaEventRecord(ev[7], side));
		CU(cudaEventRecord(h->side_join[lane], side));
	}
	if (ev) CU(cudaEventRecord(ev[2], st));
	if (blur_pos == 1)
	{
		orbx_launch_fast(P, h->maps, h->smaps, st, 1);
		orbx_launch_blur(P, h->smaps, st);
		orbx_launch_fast(P, h->maps, h->smaps, st, 2);
	}
	else orbx_launch_fast(P, h->maps, h->smaps, st);
	if (blur_pos == 2) orbx_launch_blur(P, h->smaps, st);
	if (ev) CU(cudaEventRecord(ev[3], st));
	if (ev) CU(cudaEventRecord(ev[4], st));
	orbx_launch_quadtree(P, cell_off, st);
	if (ev) CU(cudaEventRecord(ev[5], st));
	if (blur_pos == 3) orbx_launch_blur(P, h->smaps, st);
	if (!blur_inline) CU(cudaStreamWaitEvent(st, h->side_join[lane], 0));
	if (ev) CU(cudaEventRecord(ev[8], st));
	orbx_launch_describe(P, d_kps, d_desc, d_n, st);
	if (ev) CU(cudaEventRecord(ev[9], st));
	CU(cudaGetLastError());
	return ORBX_OK;
}

void note_result(orbx_extractor* h, int frames, int cap, const orbx_keypoint* d_kps, const uint8_t* d_desc, const int32_t* d_n)
{
	h->have_result = true;
	h->last_frames = frames; h->last_cap = cap;
	h->last_kps = d_kps; h->last_desc = d_desc; h->last_n = d_n;
	h->P.frames = frames;
}

}  // namespace

extern "C" {

const char* orbx_last_error(void) { return g_err.c_str(); }

int orbx_device_count(void)
{
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
	int ok = 0;
	for (int i = 0; i < count; i++)
	{
		cudaDeviceProp p;
		if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
	}
	return ok;
}

orbx_status orbx_create(const orbx_params* params, int device, orbx_handle* out)
{
	if (!params || !out) return fail(ORBX_ERR_INVALID, "null argument");
	if (params->nlevels < 1 || params->nlevels > ORBX_MAX_LEVELS) return fail(ORBX_ERR_INVALID, "nlevels must be in [1, 12]");
	if (params->nfeatures < 1) return fail(ORBX_ERR_INVALID, "nfeatures must be positive");
	if (!(params->scale_factor > 1.f)) return fail(ORBX_ERR_INVALID, "scaleFactor must be > 1");
	if (params->min_th_fast < 1 || params->ini_th_fast < params->min_th_fast || params->ini_th_fast > 254)
		return fail(ORBX_ERR_INVALID, "need 1 <= minThFAST <= iniThFAST <= 254 (a zero response makes the reference dereference null, src/ORBextractor.cc:681-691)");
	std::string why;
	if (!device_ok(device, why)) return fail(ORBX_ERR_CUDA, why);
	CU(cudaSetDevice(device));
	orbx_extractor* h = new orbx_extractor();
	h->prm = *params;
	h->device = device;
	build_tables(h);
	cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->done, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming);
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->join, cudaEventDisableTiming);
	for (int l = 0; l < 2; l++)
	{
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->side[l], cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->side_fork[l], cudaEventDisableTiming);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->side_join[l], cudaEventDisableTiming);
		for (int g = 0; g < 2; g++)
			if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->grp[l][g], cudaStreamNonBlocking);
		for (int g = 0; g < 4; g++)
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->grp_ev[l][g], cudaEventDisableTiming);
	}
	if (e == cudaSuccess && getenv("ORBX_TRACE")) e = cudaMalloc(&h->trace, 16 * sizeof(unsigned long long));
	if (e == cudaSuccess && h->trace) e = cudaMemset(h->trace, 0, 16 * sizeof(unsigned long long));
	if (e == cudaSuccess) e = orbx_upload_pattern();
	if (e == cudaSuccess) e = orbx_kernels_init();
	if (e != cudaSuccess)
	{
		delete h;
		return fail(ORBX_ERR_CUDA, cudaGetErrorString(e));
	}
	*out = h;
	return ORBX_OK;
}

orbx_status orbx_destroy(orbx_handle h)
{
	if (!h) return ORBX_OK;
	cudaSetDevice(h->device);
	if (h->stream) cudaStreamSynchronize(h->stream);
	if (h->stream2) cudaStreamSynchronize(h->stream2);
	h->l0buf.release(); h->color.release();
	h->pyr.release(); h->blur.release(); h->cand.release(); h->qbuf0.release(); h->qbuf1.release(); h->sel.release();
	h->cell_count.release(); h->cand_count.release(); h->sel_count.release(); h->pyr_done.release(); h->ovf_list.release(); h->ovf_count.release();
	h->cell_tab.release(); h->fmap_ini.release(); h->fmap_min.release();
	h->root_x.release(); h->xofs.release(); h->yofs.release(); h->root_lut.release(); h->xcoef.release(); h->ycoef.release();
	h->out_kps.release(); h->out_desc.release(); h->out_n.release();
	h->st_uright.release(); h->st_depth.release(); h->st_sad.release(); h->st_rows.release(); h->st_items.release();
	if (h->h_counts) cudaFreeHost(h->h_counts);
	if (h->small_exec) cudaGraphExecDestroy(h->small_exec);
	if (h->trace) cudaFree(h->trace);
	if (h->h_small) cudaFreeHost(h->h_small);
	for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
	if (h->done) cudaEventDestroy(h->done);
	if (h->fork) cudaEventDestroy(h->fork);
	if (h->join) cudaEventDestroy(h->join);
	if (h->stream) cudaStreamDestroy(h->stream);
	if (h->stream2) cudaStreamDestroy(h->stream2);
	if (h->copy_in) { cudaStreamSynchronize(h->copy_in); cudaStreamDestroy(h->copy_in); }
	for (cudaEvent_t e : h->ev_in) cudaEventDestroy(e);
	for (int l = 0; l < 2; l++)
	{
		if (h->side[l]) { cudaStreamSynchronize(h->side[l]); cudaStreamDestroy(h->side[l]); }
		if (h->side_fork[l]) cudaEventDestroy(h->side_fork[l]);
		if (h->side_join[l]) cudaEventDestroy(h->side_join[l]);
		for (int g = 0; g < 2; g++)
			if (h->grp[l][g]) { cudaStreamSynchronize(h->grp[l][g]); cudaStreamDestroy(h->grp[l][g]); }
		for (int g = 0; g < 4; g++)
			if (h->grp_ev[l][g]) cudaEventDestroy(h->grp_ev[l][g]);
	}
	delete h;
	return ORBX_OK;
}

orbx_status orbx_get_params(orbx_handle h, orbx_params* out)
{
	if (!h || !out) return fail(ORBX_ERR_INVALID, "null argument");
	*out = h->prm;
	return ORBX_OK;
}

orbx_status orbx_scale_tables(orbx_handle h, float* scale, float* inv_scale, float* sigma_sq, float* inv_sigma_sq)
{
	if (!h) return fail(ORBX_ERR_INVALID, "null handle");
	const size_t b = sizeof(float) * h->prm.nlevels;
	if (scale) std::memcpy(scale, h->scale.data(), b);
	if (inv_scale) std::memcpy(inv_scale, h->inv_scale.data(), b);
	if (sigma_sq) std::memcpy(sigma_sq, h->sigma_sq.data(), b);
	if (inv_sigma_sq) std::memcpy(inv_sigma_sq, h->inv_sigma_sq.data(), b);
	return ORBX_OK;
}

orbx_status orbx_feature_quotas(orbx_handle h, int32_t* quotas)
{
	if (!h || !quotas) return fail(ORBX_ERR_INVALID, "null argument");
	std::memcpy(quotas, h->quota.data(), sizeof(int) * h->prm.nlevels);
	return ORBX_OK;
}

int orbx_max_keypoints(orbx_handle h)
{
	if (!h) return 0;
	// sum over levels of max(quota + 3, 4 * roots); roots <= 16 is only known with the image size, so assume the
	// worst case before the first extract
	if (h->pw) return h->P.sel_per_frame;
	int s = 0;
	for (int q : h->quota) s += std::max(q + 3, 4 * ORBX_MAX_ROOTS);
	return s;
}

orbx_status orbx_plan(orbx_handle h, int width, int height, int frames)
{
	if (!h) return fail(ORBX_ERR_INVALID, "null handle");
	if (frames < 1 || width < 1 || height < 1) return fail(ORBX_ERR_INVALID, "bad image geometry");
	CU(cudaSetDevice(h->device));
	return build_plan(h, width, height, frames);
}

orbx_status orbx_last_result_shape(orbx_handle h, int* frames, int* cap)
{
	if (!h) return fail(ORBX_ERR_INVALID, "null handle");
	if (!h->have_result) return fail(ORBX_ERR_STATE, "no extract has run on this handle");
	if (frames) *frames = h->last_frames;
	if (cap) *cap = h->last_cap;
	return ORBX_OK;
}

orbx_status orbx_synchronize(orbx_handle h)
{
	if (!h) return fail(ORBX_ERR_INVALID, "null handle");
	CU(cudaSetDevice(h->device));
	CU(cudaStreamSynchronize(h->stream));
	return ORBX_OK;
}

void* orbx_stream(orbx_handle h) { return h ? (void*)h->stream : nullptr; }

orbx_status orbx_enable_stage_timing(orbx_handle h, int enable)
{
	if (!h) return fail(ORBX_ERR_INVALID, "null handle");
	h->stage_timing = enable != 0;
	return ORBX_OK;
}

orbx_status orbx_stage_times(orbx_handle h, float ms_sum[5], int* calls)
{
	if (!h || !ms_sum) return fail(ORBX_ERR_INVALID, "null argument");
	CU(cudaSetDevice(h->device));
	CU(cudaStreamSynchronize(h->stream));
	for (int i = 0; i < 5; i++) ms_sum[i] = 0.f;
	const size_t n = h->ev_used / 10;
	for (size_t c = 0; c < n; c++)
		for (int i = 0; i < 5; i++)
		{
			float ms = 0.f;
			CU(cudaEventElapsedTime(&ms, h->ev_pool[10 * c + 2 * i], h->ev_pool[10 * c + 2 * i + 1]));
			static const int slot[5] = { 0, 1, 2, 3, 4 };   // event pairs are pyramid, FAST, quadtree, blur, describe = the order of ms_sum
			ms_sum[slot[i]] += ms;
		}
	if (calls) *calls = (int)n;
	h->ev_used = 0;
	return ORBX_OK;
}

orbx_status orbx_extract_batch_device(orbx_handle h, const uint8_t* d_images, int frames, int width, int height,
                                      size_t pitch, size_t frame_stride, orbx_keypoint* d_kps, uint8_t* d_desc,
                                      int cap, int32_t* d_n)
{
	if (!h || !d_images || !d_kps || !d_desc || !d_n) return fail(ORBX_ERR_INVALID, "null argument");
	if (frames < 1 || width < 1 || height < 1 || pitch < (size_t)width) return fail(ORBX_ERR_INVALID, "bad image geometry");
	CU(cudaSetDevice(h->device));
	orbx_status st = build_plan(h, width, height, frames);
	if (st != ORBX_OK) return st;
	if (cap < h->P.sel_per_frame || cap >= 65536)
		return fail(ORBX_ERR_CAPACITY, "cap must be >= orbx_max_keypoints() (and < 65536)");
	// Level 0: every kernel that reads a level stages it by TMA (rows and columns outside the tensor arrive as zeros) or by 16-byte copies
	// from a 16-byte aligned row start, so a caller's buffer whose base, pitch and frame stride are multiples of 16 is read IN PLACE: no
	// repack, 2 x 300 KiB of HBM traffic per VGA frame less. The buffer then IS level 0 of the pyramid until the next extract on this handle
	// (GetImagePyramid()[0], the stereo matcher's SAD window read it). Anything else (odd pitch such as 1241, or a plan whose level 1 comes
	// from the cp.async resize kernel, which reads a few bytes outside a row) is copied into the padded level-0 buffer, which is
	// ComputePyramid's own copyTo (:462).
	static const bool allow_direct = getenv("ORBX_NO_DIRECT") == nullptr;
	const bool direct = allow_direct && ((uintptr_t)d_images % 16 == 0) && pitch % 16 == 0 && frame_stride % 16 == 0 &&
	                    (h->prm.nlevels == 1 || (h->P.lv[1].py_bw[0] > 0 && h->P.lv[1].py_bw[1] > 0));
	const uint8_t* l0 = direct ? d_images : h->l0base;
	const int64_t l0_pitch = direct ? (int64_t)pitch : h->l0_pitch, l0_stride = direct ? (int64_t)frame_stride : h->l0_stride;
	if (!direct)
	{
		if (frame_stride == pitch * (size_t)height)
			CU(cudaMemcpy2DAsync(h->l0base, h->l0_pitch, d_images, pitch, width, (size_t)height * frames, cudaMemcpyDeviceToDevice, h->stream));
		else
			for (int f = 0; f < frames; f++)
				CU(cudaMemcpy2DAsync(h->l0base + (int64_t)f * h->l0_stride, h->l0_pitch, d_images + (size_t)f * frame_stride, pitch,
				                     width, height, cudaMemcpyDeviceToDevice, h->stream));
	}
	st = ensure_level0_maps(h, l0, direct ? width : h->l0_pitch, l0_pitch, l0_stride, direct ? frames : h->frames_cap);
	if (st != ORBX_OK) return st;
	// The stages are bound by different things (FAST by instruction issue, pyramid/descriptor by memory latency), so the two
	// halves of a large batch run on two streams and fill each other's stalls. The second lane is forked from and joined back
	// into the handle's stream, so the call stays stream-ordered for the caller.
	int lanes = (frames >= 64 && !h->stage_timing) ? 2 : 1;
	if (const char* e = getenv("ORBX_LANES")) lanes = std::max(1, std::min(2, atoi(e)));
	// Sub-batches keep a chunk's pyramid + blur slabs (2.1 MB per VGA frame) inside the 126 MB L2 between the stages.
	int chunk = frames;
	if (const char* e = getenv("ORBX_DEV_CHUNK")) chunk = std::max(1, atoi(e));
	orbx_status st2 = ORBX_OK;
	if (lanes == 2)
	{
		CU(cudaEventRecord(h->fork, h->stream));
		CU(cudaStreamWaitEvent(h->stream2, h->fork, 0));
	}
	int ci = 0;
	for (int fb = 0; fb < frames && st2 == ORBX_OK; fb += chunk, ci++)
	{
		const int fc = std::min(chunk, frames - fb);
		if (lanes == 1)
			st2 = enqueue_extract(h, fb, fc, h->stream, l0, l0_pitch, l0_stride, d_kps, d_desc, d_n, cap);
		else
		{
			const int half = (fc + 1) / 2;
			st2 = enqueue_extract(h, fb, half, h->stream, l0, l0_pitch, l0_stride, d_kps, d_desc, d_n, cap, true);
			if (st2 == ORBX_OK && fc > half) st2 = enqueue_extract(h, fb + half, fc - half, h->stream2, l0, l0_pitch, l0_stride, d_kps, d_desc, d_n, cap, true);
		}
	}
	if (lanes == 2)
	{
		CU(cudaEventRecord(h->join, h->stream2));
		CU(cudaStreamWaitEvent(h->stream, h->join, 0));
	}
	if (st2 == ORBX_OK) note_result(h, frames, cap, d_kps, d_desc, d_n);
	return st2;
}

static orbx_status extract_batch_impl(orbx_handle h, const uint8_t* images, int frames, int width, int height, size_t pitch,
                                      size_t frame_stride, int channels, int rgb, orbx_keypoint* kps, uint8_t* desc, int cap, int* n,
                                      bool rectify = false)
{
	if (!h || !images || !n) return fail(ORBX_ERR_INVALID, "null argument");
	timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
	// rectify: width x height is the RAW frame; Extract runs on the rectified image, whose size is the table's
	const int raw_w = width, raw_h = height;
	if (rectify)
	{
		if (!h->rect_tab.p) return fail(ORBX_ERR_STATE, "orbx_set_rectification has not been called on this handle");
		if (width != h->rect_sw || height != h->rect_sh) return fail(ORBX_ERR_INVALID, "frame size differs from the one the rectification maps were set for");
		if (channels != 1) return fail(ORBX_ERR_INVALID, "rectification takes single-channel frames");
	}
	if (channels != 1 && channels != 3 && channels != 4)
		return fail(ORBX_ERR_INVALID, "CV_Assert(ch == 1 || ch == 3 || ch == 4) (src/System.cc:127)");
	if (frames < 1 || width < 1 || height < 1 || pitch < (size_t)width * channels) return fail(ORBX_ERR_INVALID, "bad image geometry");
	CU(cudaSetDevice(h->device));
	if (rectify) { width = h->rect_w; height = h->rect_h; }
	orbx_status st = build_plan(h, width, height, frames);
	if (st != ORBX_OK) return st;
	st = ensure_level0_maps(h, h->l0base, h->l0_pitch, h->l0_pitch, h->l0_stride, h->frames_cap);
	if (st != ORBX_OK) return st;
	const OrbxPlanDev& P = h->P;
	const int ocap = P.sel_per_frame;
	if (channels != 1) CU(h->color.ensure((size_t)frames * width * channels * height));
	if (rectify) CU(h->color.ensure((size_t)frames * raw_w * raw_h));
	// Chunk pipeline over two streams: upload of chunk c+1 and download of chunk c-1 overlap the kernels of chunk c.
	// Level 0 is uploaded straight into the level-0 buffer (ComputePyramid's copyTo, :462).
	int chunk = frames <= 32 ? frames : std::max(32, std::min(64, (frames + 3) / 4));
	if (const char* e = getenv("ORBX_CHUNK")) chunk = std::max(1, atoi(e));   // tuning knob
	const int ccap = std::min(cap, ocap);
	if (h->h_counts_n < (size_t)frames)
	{
		if (h->h_counts) cudaFreeHost(h->h_counts);
		h->h_counts = nullptr; h->h_counts_n = 0;
		CU(cudaMallocHost(&h->h_counts, sizeof(int32_t) * (size_t)frames));
		h->h_counts_n = (size_t)frames;
	}
	int32_t* counts = h->h_counts;
	// a frame at a time: results go through pinned staging (the caller's arrays are usually pageable, and a device-to-pageable copy blocks)
	const bool staged = frames <= ORBX_SMALL_BATCH && frames <= chunk && kps && desc && ccap > 0;
	if (staged)
	{
		const size_t need_bytes = (size_t)frames * ocap * (sizeof(orbx_keypoint) + 32);
		if (h->h_small_bytes < need_bytes)
		{
			if (h->h_small) cudaFreeHost(h->h_small);
			h->h_small = nullptr; h->h_small_bytes = 0;
			CU(cudaMallocHost(&h->h_small, need_bytes));
			h->h_small_bytes = need_bytes;
		}
	}
	// the per-frame buffers are shared with whatever an earlier asynchronous call (orbx_extract_batch_device, stereo) left pending on the
	// handle's stream: the second lane starts behind it
	const bool two_streams = frames > chunk;          // a single chunk (a frame at a time) never touches the second stream
	static const bool split_upload = getenv("ORBX_NO_COPY_STREAM") == nullptr;      // tuning knob
	const bool own_upload = two_streams && split_upload;
	if (two_streams)
	{
		CU(cudaEventRecord(h->fork, h->stream));
		CU(cudaStreamWaitEvent(h->stream2, h->fork, 0));
		if (own_upload)
		{
			CU(cudaStreamWaitEvent(h->copy_in, h->fork, 0));
			const size_t nchunks = (size_t)((frames + chunk - 1) / chunk);
			while (h->ev_in.size() < nchunks)
			{
				cudaEvent_t e;
				CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
				h->ev_in.push_back(e);
			}
		}
	}
	int ci = 0;
	for (int fb = 0; fb < frames; fb += chunk, ci++)
	{
		const int fc = std::min(chunk, frames - fb);
		const cudaStream_t lane_st = (ci & 1) ? h->stream2 : h->stream;
		cudaStream_t st = own_upload ? h->copy_in : lane_st;        // the uploads; the conversion kernels behind them stay on the lane
		if (rectify)
		{
			// the rectification remap (Examples/Stereo/stereo_euroc.cc:100-101) fused into the upload: raw frames land in a staging
			// buffer, the remap kernel writes level 0 of the pyramid
			const size_t rstride = (size_t)raw_w * raw_h;
			for (int f = fb; f < fb + fc; f++)
				CU(cudaMemcpy2DAsync(h->color.p + (size_t)f * rstride, raw_w, images + (size_t)f * frame_stride, pitch, raw_w, raw_h,
				                     cudaMemcpyHostToDevice, st));
			if (own_upload) { CU(cudaEventRecord(h->ev_in[ci], st)); CU(cudaStreamWaitEvent(lane_st, h->ev_in[ci], 0)); }
			orbx_launch_remap(h->color.p + (size_t)fb * rstride, raw_w, (int64_t)rstride, raw_w, raw_h, h->rect_tab.p,
			                  h->l0base + (int64_t)fb * h->l0_stride, h->l0_pitch, h->l0_stride, width, height, fc, lane_st);
		}
		else if (channels == 1)
		{
			if (frame_stride == pitch * (size_t)height)
				CU(cudaMemcpy2DAsync(h->l0base + (int64_t)fb * h->l0_stride, h->l0_pitch, images + (size_t)fb * frame_stride, pitch, width,
				                     (size_t)height * fc, cudaMemcpyHostToDevice, st));
			else
				for (int f = fb; f < fb + fc; f++)
					CU(cudaMemcpy2DAsync(h->l0base + (int64_t)f * h->l0_stride, h->l0_pitch, images + (size_t)f * frame_stride, pitch, width,
					                     height, cudaMemcpyHostToDevice, st));
			if (own_upload) { CU(cudaEventRecord(h->ev_in[ci], st)); CU(cudaStreamWaitEvent(lane_st, h->ev_in[ci], 0)); }
		}
		else
		{
			// ConvertToGray (src/System.cc:122-137) fused into the upload: colour frames land in a staging buffer and the
			// conversion kernel writes level 0 of the pyramid directly
			const size_t cpitch = (size_t)width * channels, cstride = cpitch * height;
			for (int f = fb; f < fb + fc; f++)
				CU(cudaMemcpy2DAsync(h->color.p + (size_t)f * cstride, cpitch, images + (size_t)f * frame_stride, pitch, cpitch, height,
				                     cudaMemcpyHostToDevice, st));
			if (own_upload) { CU(cudaEventRecord(h->ev_in[ci], st)); CU(cudaStreamWaitEvent(lane_st, h->ev_in[ci], 0)); }
			orbx_launch_gray(h->color.p + (size_t)fb * cstride, (int64_t)cpitch, (int64_t)cstride, channels, rgb, h->l0base + (int64_t)fb * h->l0_stride,
			                 h->l0_pitch, h->l0_stride, width, height, fc, lane_st);
		}
		st = lane_st;
		// (Replaying the ~12 launches of a one-frame call as a CUDA graph was measured: 0.212 vs 0.203 ms per call — the call is bound by the
		// GPU's dependent chain (quadtree 54 us, 7 pyramid levels 33 us), not by launch overhead, which the host hides behind it.)
		// A frame at a time: the grouped launches of enqueue_extract (three level groups + blur on four streams, ~26 API calls) and the three
		// result copies are captured once per plan and replayed as one graph launch; every pointer in them is the handle's own. (For the
		// single chain of round 1 a graph bought nothing: the GPU's dependent chain was longer than the host's launch sequence. With the
		// chain cut into parallel groups the host became the limit.) Tuning knob ORBX_GRAPH=0: plain launches.
		static const bool graph_on = !(getenv("ORBX_GRAPH") && atoi(getenv("ORBX_GRAPH")) == 0) && !(getenv("ORBX_GROUPS") && atoi(getenv("ORBX_GROUPS")) == 0);
		if (staged && graph_on && !h->stage_timing && h->P.nlevels >= 4)
		{
			const orbx_extractor::SmallKey key = { h->gen, h->l0base, h->out_kps.p, h->out_desc.p, h->out_n.p, h->h_small, counts, frames, ocap };
			if (!h->small_exec || std::memcmp(&key, &h->small_key, sizeof(key)) != 0)
			{
				if (h->small_exec) { cudaGraphExecDestroy(h->small_exec); h->small_exec = nullptr; }
				cudaGraph_t g = nullptr;
				CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
				orbx_status e = enqueue_extract(h, fb, fc, st, h->l0base, h->l0_pitch, h->l0_stride, h->out_kps.p, h->out_desc.p, h->out_n.p, ocap, false, true);
				cudaError_t ce = cudaSuccess;
				if (e == ORBX_OK) ce = cudaMemcpyAsync(counts + fb, h->out_n.p + fb, sizeof(int32_t) * fc, cudaMemcpyDeviceToHost, st);
				if (e == ORBX_OK && ce == cudaSuccess) ce = cudaMemcpyAsync(h->h_small, h->out_kps.p, sizeof(orbx_keypoint) * (size_t)ocap * frames, cudaMemcpyDeviceToHost, st);
				if (e == ORBX_OK && ce == cudaSuccess)
					ce = cudaMemcpyAsync(h->h_small + sizeof(orbx_keypoint) * (size_t)ocap * frames, h->out_desc.p, (size_t)32 * ocap * frames, cudaMemcpyDeviceToHost, st);
				const cudaError_t ee = cudaStreamEndCapture(st, &g);          // always: the stream must leave capture mode
				if (e != ORBX_OK) { if (g) cudaGraphDestroy(g); return e; }
				if (ce != cudaSuccess || ee != cudaSuccess) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return fail(ORBX_ERR_CUDA, "capturing the one-frame graph failed"); }
				ce = cudaGraphInstantiate(&h->small_exec, g, 0);
				cudaGraphDestroy(g);
				if (ce != cudaSuccess) { h->small_exec = nullptr; return fail(ORBX_ERR_CUDA, cudaGetErrorString(ce)); }
				h->small_key = key;
			}
			CU(cudaGraphLaunch(h->small_exec, st));
			continue;
		}
		orbx_status e = enqueue_extract(h, fb, fc, st, h->l0base, h->l0_pitch, h->l0_stride, h->out_kps.p, h->out_desc.p, h->out_n.p, ocap);
		if (e != ORBX_OK) return e;
		CU(cudaMemcpyAsync(counts + fb, h->out_n.p + fb, sizeof(int32_t) * fc, cudaMemcpyDeviceToHost, st));
		if (staged)
		{
			// whole capacity rows of this (only) chunk, contiguous: keypoints, then descriptors
			CU(cudaMemcpyAsync(h->h_small, h->out_kps.p, sizeof(orbx_keypoint) * (size_t)ocap * frames, cudaMemcpyDeviceToHost, st));
			CU(cudaMemcpyAsync(h->h_small + sizeof(orbx_keypoint) * (size_t)ocap * frames, h->out_desc.p, (size_t)32 * ocap * frames, cudaMemcpyDeviceToHost, st));
		}
		else if (kps && desc && ccap > 0)
		{
			CU(cudaMemcpy2DAsync(kps + (size_t)fb * cap, sizeof(orbx_keypoint) * (size_t)cap, h->out_kps.p + (size_t)fb * ocap,
			                     sizeof(orbx_keypoint) * (size_t)ocap, sizeof(orbx_keypoint) * (size_t)ccap, fc, cudaMemcpyDeviceToHost, st));
			CU(cudaMemcpy2DAsync(desc + (size_t)fb * cap * 32, (size_t)32 * cap, h->out_desc.p + (size_t)fb * ocap * 32, (size_t)32 * ocap,
			                     (size_t)32 * ccap, fc, cudaMemcpyDeviceToHost, st));
		}
	}
	static const bool trace_on = getenv("ORBX_TRACE") != nullptr;
	timespec ts1; if (trace_on) clock_gettime(CLOCK_MONOTONIC, &ts1);
	CU(cudaStreamSynchronize(h->stream));
	if (two_streams) CU(cudaStreamSynchronize(h->stream2));
	if (trace_on && h->trace)
	{
		timespec ts2; clock_gettime(CLOCK_MONOTONIC, &ts2);
		static int calls = 0;
		if (++calls % 64 == 0)
		{
			unsigned long long t[16];
			cudaMemcpy(t, h->trace, sizeof(t), cudaMemcpyDeviceToHost);
			const char* names[11] = { "start", "pyrB", "fastA", "qtA", "pyrC", "fastB", "qtB", "fastC", "qtC", "blur", "describe" };
			fprintf(stderr, "[trace us] enqueue %.1f sync-wait %.1f |", (ts1.tv_sec - ts0.tv_sec) * 1e6 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-3,
			        (ts2.tv_sec - ts1.tv_sec) * 1e6 + (ts2.tv_nsec - ts1.tv_nsec) * 1e-3);
			for (int i = 1; i < 11; i++) fprintf(stderr, " %s %.1f", names[i], ((double)t[i] - (double)t[0]) * 1e-3);
			fprintf(stderr, "\n");
		}
	}
	note_result(h, frames, ocap, h->out_kps.p, h->out_desc.p, h->out_n.p);
	int need = 0;
	for (int f = 0; f < frames; f++) { n[f] = counts[f]; need = std::max(need, counts[f]); }
	if (staged)
	{
		const uint8_t* sk = h->h_small; const uint8_t* sd = h->h_small + sizeof(orbx_keypoint) * (size_t)ocap * frames;
		for (int f = 0; f < frames; f++)
		{
			const size_t rows = (size_t)std::min(counts[f], ccap);
			std::memcpy(kps + (size_t)f * cap, sk + sizeof(orbx_keypoint) * (size_t)ocap * f, sizeof(orbx_keypoint) * rows);
			std::memcpy(desc + (size_t)f * cap * 32, sd + (size_t)32 * ocap * f, 32 * rows);
		}
	}
	if (need > cap || (need > 0 && (!kps || !desc)))
		return fail(ORBX_ERR_CAPACITY, "output buffers hold fewer keypoints than were found");
	return ORBX_OK;
}

orbx_status orbx_extract_batch(orbx_handle h, const uint8_t* images, int frames, int width, int height, size_t pitch,
                               size_t frame_stride, orbx_keypoint* kps, uint8_t* desc, int cap, int* n)
{
	return extract_batch_impl(h, images, frames, width, height, pitch, frame_stride, 1, 0, kps, desc, cap, n);
}

orbx_status orbx_extract_batch_color(orbx_handle h, const uint8_t* images, int frames, int width, int height, size_t pitch,
                                     size_t frame_stride, int channels, int rgb, orbx_keypoint* kps, uint8_t* desc, int cap, int* n)
{
	return extract_batch_impl(h, images, frames, width, height, pitch, frame_stride, channels, rgb, kps, desc, cap, n);
}

orbx_status orbx_extract(orbx_handle h, const uint8_t* image, int width, int height, size_t pitch,
                         orbx_keypoint* kps, uint8_t* desc, int cap, int* n)
{
	return orbx_extract_batch(h, image, 1, width, height, pitch, pitch * (size_t)height, kps, desc, cap, n);
}

orbx_status orbx_level_size(orbx_handle h, int level, int* width, int* height)
{
	if (!h || !h->pw) return fail(ORBX_ERR_STATE, "no extract has run on this handle");
	if (level < 0 || level >= h->prm.nlevels) return fail(ORBX_ERR_INVALID, "level out of range");
	if (width) *width = h->P.lv[level].w;
	if (height) *height = h->P.lv[level].h;
	return ORBX_OK;
}

orbx_status orbx_pyramid_level_device(orbx_handle h, int frame, int level, const uint8_t** d_ptr, size_t* pitch)
{
	if (!h || !h->have_result) return fail(ORBX_ERR_STATE, "no extract has run on this handle");
	if (level < 0 || level >= h->prm.nlevels || frame < 0 || frame >= h->last_frames) return fail(ORBX_ERR_INVALID, "frame/level out of range");
	const OrbxPlanDev& P = h->P;
	if (level == 0) { *d_ptr = P.l0 + (int64_t)frame * P.l0_stride; *pitch = (size_t)P.l0_pitch; }
	else { *d_ptr = P.pyr + (int64_t)frame * P.slab + P.lv[level].offset; *pitch = (size_t)P.lv[level].pitch; }
	return ORBX_OK;
}

orbx_status orbx_pyramid_level(orbx_handle h, int frame, int level, uint8_t* dst, size_t dst_pitch)
{
	const uint8_t* p; size_t pitch;
	orbx_status st = orbx_pyramid_level_device(h, frame, level, &p, &pitch);
	if (st != ORBX_OK) return st;
	CU(cudaSetDevice(h->device));
	CU(cudaMemcpy2DAsync(dst, dst_pitch, p, pitch, h->P.lv[level].w, h->P.lv[level].h, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return ORBX_OK;
}

orbx_status orbx_debug_blurred_level(orbx_handle h, int frame, int level, uint8_t* dst, size_t dst_pitch)
{
	if (!h || !h->have_result) return fail(ORBX_ERR_STATE, "no extract has run on this handle");
	if (level < 0 || level >= h->prm.nlevels || frame < 0 || frame >= h->last_frames) return fail(ORBX_ERR_INVALID, "frame/level out of range");
	const OrbxPlanDev& P = h->P;
	CU(cudaSetDevice(h->device));
	CU(cudaMemcpy2DAsync(dst, dst_pitch, P.blur + (int64_t)frame * P.slab + P.lv[level].offset, P.lv[level].pitch, P.lv[level].w,
	                     P.lv[level].h, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return ORBX_OK;
}

orbx_status orbx_debug_candidates(orbx_handle h, int frame, int level, int32_t* xyr, int cap, int* n)
{
	if (!h || !h->have_result) return fail(ORBX_ERR_STATE, "no extract has run on this handle");
	if (level < 0 || level >= h->prm.nlevels || frame < 0 || frame >= h->last_frames) return fail(ORBX_ERR_INVALID, "frame/level out of range");
	const OrbxPlanDev& P = h->P;
	const OrbxLevel& L = P.lv[level];
	CU(cudaSetDevice(h->device));
	const int ncell = L.ncx * L.ncy;
	std::vector<int> counts(ncell);
	CU(cudaMemcpy(counts.data(), P.cell_count + (int64_t)frame * P.cells_per_frame + L.cell_base, sizeof(int) * ncell, cudaMemcpyDeviceToHost));
	std::vector<uint32_t> slots((size_t)L.cand_cap);
	CU(cudaMemcpy(slots.data(), P.cand + (int64_t)frame * P.cand_per_frame + L.cand_base, sizeof(uint32_t) * slots.size(), cudaMemcpyDeviceToHost));
	int total = 0;
	for (int c = 0; c < ncell; c++)
		for (int k = 0; k < counts[c]; k++, total++)
			if (total < cap)
			{
				const uint32_t v = slots[(size_t)c * L.cell_cap + k];
				xyr[3 * total] = orbx_px(v); xyr[3 * total + 1] = orbx_py(v); xyr[3 * total + 2] = orbx_pr(v);
			}
	*n = total;
	return total > cap ? fail(ORBX_ERR_CAPACITY, "candidate buffer too small") : ORBX_OK;
}

orbx_status orbx_debug_selected(orbx_handle h, int frame, int level, int32_t* xyr, int cap, int* n)
{
	if (!h || !h->have_result) return fail(ORBX_ERR_STATE, "no extract has run on this handle");
	if (level < 0 || level >= h->prm.nlevels || frame < 0 || frame >= h->last_frames) return fail(ORBX_ERR_INVALID, "frame/level out of range");
	const OrbxPlanDev& P = h->P;
	const OrbxLevel& L = P.lv[level];
	CU(cudaSetDevice(h->device));
	int cnt = 0;
	CU(cudaMemcpy(&cnt, P.sel_count + (int64_t)frame * P.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
	*n = cnt;
	if (cnt > cap) return fail(ORBX_ERR_CAPACITY, "selected buffer too small");
	std::vector<uint32_t> v(std::max(cnt, 1));
	CU(cudaMemcpy(v.data(), P.sel + (int64_t)frame * P.sel_per_frame + L.sel_base, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost));
	for (int i = 0; i < cnt; i++) { xyr[3 * i] = orbx_px(v[i]); xyr[3 * i + 1] = orbx_py(v[i]); xyr[3 * i + 2] = orbx_pr(v[i]); }
	return ORBX_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// matching
// ---------------------------------------------------------------------------------------------------------------
orbx_status orbx_descriptor_distance(int device, const uint8_t* a, const uint8_t* b, int64_t n, int32_t* dist)
{
	if (!a || !b || !dist || n < 0) return fail(ORBX_ERR_INVALID, "bad argument");
	std::string why;
	if (!device_ok(device, why)) return fail(ORBX_ERR_CUDA, why);
	if (n == 0) return ORBX_OK;
	CU(cudaSetDevice(device));
	DevBuf<uint8_t> da, db; DevBuf<int32_t> dd;
	CU(da.ensure(32 * n)); CU(db.ensure(32 * n)); CU(dd.ensure(n));
	CU(cudaMemcpy(da.p, a, 32 * n, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(db.p, b, 32 * n, cudaMemcpyHostToDevice));
	orbx_launch_hamming_pairs(da.p, db.p, n, dd.p, 0);
	CU(cudaGetLastError());
	CU(cudaMemcpy(dist, dd.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
	da.release(); db.release(); dd.release();
	return ORBX_OK;
}

orbx_status orbx_knn2_partial_device(const uint8_t* d_query, int64_t nq, const uint8_t* d_train_shard, int64_t nt_shard,
                                     int64_t index_base, uint64_t* d_partial, void* stream)
{
	if (!d_query || !d_train_shard || !d_partial || nq < 1 || nt_shard < 0) return fail(ORBX_ERR_INVALID, "bad argument");
	if (index_base + nt_shard > 0xfffffffell) return fail(ORBX_ERR_INVALID, "train index does not fit 32 bits");
	cudaStream_t st = (cudaStream_t)stream;
	const int splits = orbx_knn2_splits(nq, nt_shard);
	if (splits == 1)
	{
		orbx_launch_knn2_partial(d_query, nq, d_train_shard, nt_shard, index_base, d_partial, st);
	}
	else
	{
		// few queries: the shard is scanned in `splits` slices to fill the GPU, then folded to one partial per query
		uint64_t* tmp = nullptr;
		CU(cudaMallocAsync(&tmp, sizeof(uint64_t) * (size_t)splits * nq, st));
		orbx_launch_knn2_partial(d_query, nq, d_train_shard, nt_shard, index_base, tmp, st);
		orbx_launch_knn2_fold(tmp, splits, nq, d_partial, st);
		CU(cudaFreeAsync(tmp, st));
	}
	CU(cudaGetLastError());
	return ORBX_OK;
}

orbx_status orbx_knn2_merge_device(const uint64_t* d_gathered, int ranks, int64_t nq, int th_low, float nnratio,
                                   int32_t* d_idx, uint16_t* d_best, uint16_t* d_second, int32_t* d_match, void* stream)
{
	if (!d_gathered || ranks < 1 || nq < 1) return fail(ORBX_ERR_INVALID, "bad argument");
	orbx_launch_knn2_merge(d_gathered, ranks, nq, th_low, nnratio, d_idx, d_best, d_second, d_match, (cudaStream_t)stream);
	CU(cudaGetLastError());
	return ORBX_OK;
}

int orbx_knn2_work_parts(int64_t nq, int64_t nt_shard) { return orbx_knn2_splits(nq, nt_shard) > 1 ? orbx_knn2_splits(nq, nt_shard) + 1 : 1; }

// ncclAllGather(sendbuff, recvbuff, sendcount, datatype, comm, stream); ncclUint64 = 5 in every NCCL 2.x (nccl.h: ncclInt8 0, ncclUint8 1,
// ncclInt32 2, ncclUint32 3, ncclInt64 4, ncclUint64 5). The library itself does not link NCCL.
typedef int (*orbx_nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
static orbx_nccl_allgather_fn resolve_nccl_allgather()
{
	static orbx_nccl_allgather_fn fn = []() -> orbx_nccl_allgather_fn {
		void* sym = dlsym(RTLD_DEFAULT, "ncclAllGather");        // the NCCL the caller's communicator came from
		if (!sym)
			for (const char* name : { "libnccl.so.2", "libnccl.so" })
				if (void* lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL)) { sym = dlsym(lib, "ncclAllGather"); if (sym) break; }
		return reinterpret_cast<orbx_nccl_allgather_fn>(sym);
	}();
	return fn;
}

orbx_status orbx_knn2_sharded(void* comm, int rank, int nranks, const uint8_t* d_query, int64_t nq, const uint8_t* d_train_shard,
                              int64_t nt_shard, int64_t index_base, int th_low, float nnratio, int32_t* d_idx, uint16_t* d_best,
                              uint16_t* d_second, int32_t* d_match, uint64_t* d_work, void* stream)
{
	if (!d_query || !d_train_shard || nq < 1 || nt_shard < 0 || nranks < 1 || rank < 0 || rank >= nranks) return fail(ORBX_ERR_INVALID, "bad argument");
	if (nranks > 1 && !comm) return fail(ORBX_ERR_INVALID, "an NCCL communicator is required for more than one rank");
	cudaStream_t st = (cudaStream_t)stream;
	const int parts = orbx_knn2_work_parts(nq, nt_shard);
	uint64_t* work = d_work;
	if (!work) CU(cudaMallocAsync(&work, sizeof(uint64_t) * (size_t)(nranks + parts) * nq, st));
	uint64_t* gathered = work;                                   // [nranks][nq]
	uint64_t* mine = work + (size_t)nranks * nq;                 // this rank's partial (and its split scratch behind it)
	orbx_status rc = ORBX_OK;
	if (parts == 1)
		orbx_launch_knn2_partial(d_query, nq, d_train_shard, nt_shard, index_base, mine, st);
	else
	{
		orbx_launch_knn2_partial(d_query, nq, d_train_shard, nt_shard, index_base, mine + nq, st);
		orbx_launch_knn2_fold(mine + nq, parts - 1, nq, mine, st);
	}
	if (nranks == 1)
		orbx_launch_knn2_merge(mine, 1, nq, th_low, nnratio, d_idx, d_best, d_second, d_match, st);
	else
	{
		orbx_nccl_allgather_fn allgather = resolve_nccl_allgather();
		if (!allgather) rc = fail(ORBX_ERR_STATE, "ncclAllGather not found: no NCCL is loaded in this process and libnccl.so.2 cannot be opened");
		else
		{
			const int r = allgather(mine, gathered, (size_t)nq, 5 /* ncclUint64 */, comm, st);
			if (r != 0) rc = fail(ORBX_ERR_CUDA, "ncclAllGather failed with ncclResult_t " + std::to_string(r));
			else orbx_launch_knn2_merge(gathered, nranks, nq, th_low, nnratio, d_idx, d_best, d_second, d_match, st);
		}
	}
	if (!d_work) cudaFreeAsync(work, st);
	if (rc != ORBX_OK) return rc;
	CU(cudaGetLastError());
	return ORBX_OK;
}

orbx_status orbx_knn2_device(const uint8_t* d_query, int64_t nq, const uint8_t* d_train, int64_t nt, int th_low,
                             float nnratio, int32_t* d_idx, uint16_t* d_best, uint16_t* d_second, int32_t* d_match,
                             void* stream)
{
	if (!d_query || !d_train || nq < 1 || nt < 0) return fail(ORBX_ERR_INVALID, "bad argument");
	if (nt > 0xfffffffell) return fail(ORBX_ERR_INVALID, "train index does not fit 32 bits");
	cudaStream_t st = (cudaStream_t)stream;
	const int splits = orbx_knn2_splits(nq, nt);
	uint64_t* partial = nullptr;
	CU(cudaMallocAsync(&partial, sizeof(uint64_t) * (size_t)splits * nq, st));
	orbx_launch_knn2_partial(d_query, nq, d_train, nt, 0, partial, st);
	orbx_launch_knn2_merge(partial, splits, nq, th_low, nnratio, d_idx, d_best, d_second, d_match, st);
	CU(cudaGetLastError());
	CU(cudaFreeAsync(partial, st));
	return ORBX_OK;
}

orbx_status orbx_knn2(int device, const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt, int th_low,
                      float nnratio, int32_t* idx, uint16_t* best, uint16_t* second, int32_t* match)
{
	if (!query || !train || nq < 1 || nt < 0) return fail(ORBX_ERR_INVALID, "bad argument");
	std::string why;
	if (!device_ok(device, why)) return fail(ORBX_ERR_CUDA, why);
	CU(cudaSetDevice(device));
	DevBuf<uint8_t> dq, dt; DevBuf<int32_t> di, dm; DevBuf<uint16_t> db, ds;
	CU(dq.ensure(32 * nq)); CU(dt.ensure(std::max<int64_t>(32 * nt, 32)));
	CU(di.ensure(nq)); CU(dm.ensure(nq)); CU(db.ensure(nq)); CU(ds.ensure(nq));
	CU(cudaMemcpy(dq.p, query, 32 * nq, cudaMemcpyHostToDevice));
	if (nt) CU(cudaMemcpy(dt.p, train, 32 * nt, cudaMemcpyHostToDevice));
	orbx_status st = orbx_knn2_device(dq.p, nq, dt.p, nt, th_low, nnratio, di.p, db.p, ds.p, dm.p, nullptr);
	if (st == ORBX_OK)
	{
		CU(cudaDeviceSynchronize());
		if (idx) CU(cudaMemcpy(idx, di.p, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost));
		if (best) CU(cudaMemcpy(best, db.p, sizeof(uint16_t) * nq, cudaMemcpyDeviceToHost));
		if (second) CU(cudaMemcpy(second, ds.p, sizeof(uint16_t) * nq, cudaMemcpyDeviceToHost));
		if (match) CU(cudaMemcpy(match, dm.p, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost));
	}
	dq.release(); dt.release(); di.release(); dm.release(); db.release(); ds.release();
	return st;
}

orbx_status orbx_stereo_match_device(orbx_handle left, orbx_handle right, const orbx_camera* camera, float* d_uright, float* d_depth)
{
	if (!left || !right || !camera || !d_uright || !d_depth) return fail(ORBX_ERR_INVALID, "null argument");
	if (!left->have_result || !right->have_result) return fail(ORBX_ERR_STATE, "both extractors must have run Extract first");
	if (left->device != right->device) return fail(ORBX_ERR_INVALID, "left and right extractors live on different devices");
	if (left->last_frames != right->last_frames || left->pw != right->pw || left->ph != right->ph ||
	    left->prm.nlevels != right->prm.nlevels || left->last_cap != right->last_cap)
		return fail(ORBX_ERR_INVALID, "left and right extractions differ in batch size, image size or parameters");
	CU(cudaSetDevice(left->device));
	const OrbxPlanDev& PL = left->P; const OrbxPlanDev& PR = right->P;
	OrbxStereoArgs A;
	std::memset(&A, 0, sizeof(A));
	A.frames = left->last_frames; A.cap = left->last_cap; A.nlevels = PL.nlevels;
	A.kl = left->last_kps; A.dl = left->last_desc; A.nl = left->last_n;
	A.kr = right->last_kps; A.dr = right->last_desc; A.nr = right->last_n;
	A.pl0 = PL.l0; A.pl0_pitch = PL.l0_pitch; A.pl0_stride = PL.l0_stride; A.pl = PL.pyr; A.pl_slab = PL.slab;
	A.pr0 = PR.l0; A.pr0_pitch = PR.l0_pitch; A.pr0_stride = PR.l0_stride; A.pr = PR.pyr; A.pr_slab = PR.slab;
	for (int s = 0; s < PL.nlevels; s++)
	{
		A.lw[s] = PL.lv[s].w; A.lh[s] = PL.lv[s].h; A.lpitch[s] = PL.lv[s].pitch; A.loff[s] = PL.lv[s].offset;
		A.scale[s] = left->scale[s]; A.inv_scale[s] = left->inv_scale[s];
	}
	A.bf = camera->bf; A.baseline = camera->baseline;
	A.uright = d_uright; A.depth = d_depth;
	CU(left->st_sad.ensure((size_t)A.frames * A.cap));
	A.sad = left->st_sad.p;
	A.rows = PL.lv[0].h;
	A.items_cap = A.cap * orbx_stereo_items_per_keypoint(left->scale[PL.nlevels - 1]);
	CU(left->st_rows.ensure((size_t)A.frames * (A.rows + 1)));
	CU(left->st_items.ensure((size_t)A.frames * A.items_cap));
	A.row_start = left->st_rows.p; A.row_items = left->st_items.p;
	// the right extraction runs on its own stream: order it before the match
	CU(cudaEventRecord(right->done, right->stream));
	CU(cudaStreamWaitEvent(left->stream, right->done, 0));
	orbx_launch_stereo(A, left->stream);
	CU(cudaGetLastError());
	return ORBX_OK;
}

orbx_status orbx_stereo_match(orbx_handle left, orbx_handle right, const orbx_camera* camera, float* uright, float* depth)
{
	if (!left || !uright || !depth) return fail(ORBX_ERR_INVALID, "null argument");
	if (!left->have_result) return fail(ORBX_ERR_STATE, "both extractors must have run Extract first");
	CU(cudaSetDevice(left->device));
	const size_t count = (size_t)left->last_frames * left->last_cap;
	CU(left->st_uright.ensure(count)); CU(left->st_depth.ensure(count));
	orbx_status st = orbx_stereo_match_device(left, right, camera, left->st_uright.p, left->st_depth.p);
	if (st != ORBX_OK) return st;
	CU(cudaMemcpyAsync(uright, left->st_uright.p, sizeof(float) * count, cudaMemcpyDeviceToHost, left->stream));
	CU(cudaMemcpyAsync(depth, left->st_depth.p, sizeof(float) * count, cudaMemcpyDeviceToHost, left->stream));
	CU(cudaStreamSynchronize(left->stream));
	return ORBX_OK;
}

orbx_status orbx_stereo_match_host(int device, const orbx_keypoint* kps_l, int n_l, const uint8_t* desc_l,
                                   const uint8_t* const* pyr_l, const orbx_keypoint* kps_r, int n_r,
                                   const uint8_t* desc_r, const uint8_t* const* pyr_r, const int* level_w,
                                   const int* level_h, const size_t* level_pitch, int nlevels, const float* scale,
                                   const float* inv_scale, const orbx_camera* camera, float* uright, float* depth)
{
	if (!kps_l || !kps_r || !desc_l || !desc_r || !pyr_l || !pyr_r || !camera || !uright || !depth || n_l < 0 || n_r < 0)
		return fail(ORBX_ERR_INVALID, "bad argument");
	if (nlevels < 1 || nlevels > ORBX_MAX_LEVELS) return fail(ORBX_ERR_INVALID, "nlevels must be in [1, 12]");
	if (n_l >= 65536 || n_r >= 65536) return fail(ORBX_ERR_INVALID, "more than 65535 keypoints");
	if (level_h[0] > 4096) return fail(ORBX_ERR_INVALID, "image taller than 4096 rows");
	std::string why;
	if (!device_ok(device, why)) return fail(ORBX_ERR_CUDA, why);
	if (n_l == 0) return ORBX_OK;
	CU(cudaSetDevice(device));
	const int cap = std::max(n_l, std::max(n_r, 1));
	OrbxStereoArgs A;
	std::memset(&A, 0, sizeof(A));
	A.frames = 1; A.cap = cap; A.nlevels = nlevels;
	int64_t slab = 0;
	for (int s = 0; s < nlevels; s++)
	{
		A.lw[s] = level_w[s]; A.lh[s] = level_h[s];
		A.lpitch[s] = (int)align_up(level_w[s], 16);
		A.loff[s] = slab;
		slab += (int64_t)A.lpitch[s] * level_h[s];
		A.scale[s] = scale[s]; A.inv_scale[s] = inv_scale[s];
	}
	DevBuf<uint8_t> pl, pr, dl, dr; DevBuf<orbx_keypoint> kl, kr; DevBuf<int32_t> cnt; DevBuf<float> du, dd; DevBuf<int> sad, rowbuf;
	DevBuf<uint2> itembuf;
	float max_scale = 1.f;
	for (int s = 0; s < nlevels; s++) max_scale = std::max(max_scale, scale[s]);
	A.rows = level_h[0];
	A.items_cap = cap * orbx_stereo_items_per_keypoint(max_scale);
	CU(rowbuf.ensure((size_t)A.rows + 1)); CU(itembuf.ensure((size_t)A.items_cap));
	A.row_start = rowbuf.p; A.row_items = itembuf.p;
	CU(pl.ensure(slab)); CU(pr.ensure(slab)); CU(dl.ensure((size_t)cap * 32)); CU(dr.ensure((size_t)cap * 32));
	CU(kl.ensure(cap)); CU(kr.ensure(cap)); CU(cnt.ensure(2)); CU(du.ensure(cap)); CU(dd.ensure(cap)); CU(sad.ensure(cap));
	for (int s = 0; s < nlevels; s++)
	{
		CU(cudaMemcpy2D(pl.p + A.loff[s], A.lpitch[s], pyr_l[s], level_pitch[s], level_w[s], level_h[s], cudaMemcpyHostToDevice));
		CU(cudaMemcpy2D(pr.p + A.loff[s], A.lpitch[s], pyr_r[s], level_pitch[s], level_w[s], level_h[s], cudaMemcpyHostToDevice));
	}
	CU(cudaMemcpy(kl.p, kps_l, sizeof(orbx_keypoint) * n_l, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(dl.p, desc_l, (size_t)32 * n_l, cudaMemcpyHostToDevice));
	if (n_r)
	{
		CU(cudaMemcpy(kr.p, kps_r, sizeof(orbx_keypoint) * n_r, cudaMemcpyHostToDevice));
		CU(cudaMemcpy(dr.p, desc_r, (size_t)32 * n_r, cudaMemcpyHostToDevice));
	}
	const int32_t counts[2] = { n_l, n_r };
	CU(cudaMemcpy(cnt.p, counts, sizeof(counts), cudaMemcpyHostToDevice));
	A.kl = kl.p; A.dl = dl.p; A.nl = cnt.p; A.kr = kr.p; A.dr = dr.p; A.nr = cnt.p + 1;
	// level 0 is addressed through the l0 fields, the rest through the slab
	A.pl0 = pl.p; A.pl0_pitch = A.lpitch[0]; A.pl0_stride = slab; A.pl = pl.p; A.pl_slab = slab;
	A.pr0 = pr.p; A.pr0_pitch = A.lpitch[0]; A.pr0_stride = slab; A.pr = pr.p; A.pr_slab = slab;
	A.bf = camera->bf; A.baseline = camera->baseline;
	A.uright = du.p; A.depth = dd.p; A.sad = sad.p;
	orbx_launch_stereo(A, 0);
	CU(cudaGetLastError());
	CU(cudaMemcpy(uright, du.p, sizeof(float) * n_l, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(depth, dd.p, sizeof(float) * n_l, cudaMemcpyDeviceToHost));
	pl.release(); pr.release(); dl.release(); dr.release(); kl.release(); kr.release(); cnt.release(); du.release(); dd.release(); sad.release();
	return ORBX_OK;
}

orbx_status orbx_convert_to_gray(int device, const uint8_t* src, int width, int height, size_t pitch, int channels, int rgb,
                                 uint8_t* dst, size_t dst_pitch)
{
	if (!src || !dst || width < 1 || height < 1) return fail(ORBX_ERR_INVALID, "bad argument");
	if (channels != 3 && channels != 4) return fail(ORBX_ERR_INVALID, "channels must be 3 or 4");
	std::string why;
	if (!device_ok(device, why)) return fail(ORBX_ERR_CUDA, why);
	CU(cudaSetDevice(device));
	const size_t cpitch = (size_t)width * channels;
	const int64_t dpitch = align_up(width, 128);
	DevBuf<uint8_t> ds, dd;
	CU(ds.ensure(cpitch * height)); CU(dd.ensure((size_t)dpitch * height));
	CU(cudaMemcpy2D(ds.p, cpitch, src, pitch, cpitch, height, cudaMemcpyHostToDevice));
	orbx_launch_gray(ds.p, (int64_t)cpitch, 0, channels, rgb, dd.p, dpitch, 0, width, height, 1, 0);
	CU(cudaGetLastError());
	CU(cudaMemcpy2D(dst, dst_pitch, dd.p, dpitch, width, height, cudaMemcpyDeviceToHost));
	return ORBX_OK;
}

orbx_status orbx_remap(int device, const uint8_t* src, int src_width, int src_height, size_t src_pitch, const float* map1, const float* map2,
                       size_t map_pitch, uint8_t* dst, int width, int height, size_t dst_pitch)
{
	if (!src || !map1 || !map2 || !dst || src_width < 1 || src_height < 1 || width < 1 || height < 1) return fail(ORBX_ERR_INVALID, "bad argument");
	std::string why;
	if (!device_ok(device, why)) return fail(ORBX_ERR_CUDA, why);
	CU(cudaSetDevice(device));
	std::vector<int2> tab;
	pack_remap_table(map1, map2, map_pitch, width, height, tab);
	const int64_t dpitch = align_up(width, 128);
	DevBuf<uint8_t> ds, dd; DevBuf<int2> dt;
	CU(ds.ensure((size_t)src_width * src_height)); CU(dd.ensure((size_t)dpitch * height)); CU(dt.ensure(tab.size()));
	CU(cudaMemcpy2D(ds.p, src_width, src, src_pitch, src_width, src_height, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(dt.p, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice));
	orbx_launch_remap(ds.p, src_width, 0, src_width, src_height, dt.p, dd.p, dpitch, 0, width, height, 1, 0);
	CU(cudaGetLastError());
	CU(cudaMemcpy2D(dst, dst_pitch, dd.p, dpitch, width, height, cudaMemcpyDeviceToHost));
	return ORBX_OK;
}

orbx_status orbx_set_rectification(orbx_handle h, const float* map1, const float* map2, size_t map_pitch, int width, int height, int src_width,
                                   int src_height)
{
	if (!h || !map1 || !map2 || width < 1 || height < 1 || src_width < 1 || src_height < 1) return fail(ORBX_ERR_INVALID, "bad argument");
	CU(cudaSetDevice(h->device));
	std::vector<int2> tab;
	pack_remap_table(map1, map2, map_pitch, width, height, tab);
	CU(cudaStreamSynchronize(h->stream)); CU(cudaStreamSynchronize(h->stream2));
	CU(h->rect_tab.ensure(tab.size()));
	CU(cudaMemcpy(h->rect_tab.p, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice));
	h->rect_w = width; h->rect_h = height; h->rect_sw = src_width; h->rect_sh = src_height;
	return ORBX_OK;
}

orbx_status orbx_rectify_batch_device(orbx_handle h, const uint8_t* d_raw, int frames, size_t pitch, size_t frame_stride, uint8_t* d_dst,
                                      size_t dst_pitch, size_t dst_stride)
{
	if (!h || !d_raw || !d_dst || frames < 1) return fail(ORBX_ERR_INVALID, "bad argument");
	if (!h->rect_tab.p) return fail(ORBX_ERR_STATE, "orbx_set_rectification has not been called on this handle");
	if (dst_pitch % 4 != 0 || dst_pitch < (size_t)h->rect_w) return fail(ORBX_ERR_INVALID, "dst_pitch must be a multiple of 4 and >= the rectified width");
	CU(cudaSetDevice(h->device));
	orbx_launch_remap(d_raw, (int64_t)pitch, (int64_t)frame_stride, h->rect_sw, h->rect_sh, h->rect_tab.p, d_dst, (int64_t)dst_pitch,
	                  (int64_t)dst_stride, h->rect_w, h->rect_h, frames, h->stream);
	CU(cudaGetLastError());
	return ORBX_OK;
}

orbx_status orbx_extract_batch_rectified(orbx_handle h, const uint8_t* images, int frames, int src_width, int src_height, size_t pitch,
                                         size_t frame_stride, orbx_keypoint* kps, uint8_t* desc, int cap, int* n)
{
	return extract_batch_impl(h, images, frames, src_width, src_height, pitch, frame_stride, 1, 0, kps, desc, cap, n, true);
}

orbx_status orbx_undistort_keypoints(int device, const orbx_keypoint* kps, int n, const orbx_camera* camera, const float* dist, int ndist,
                                     orbx_keypoint* kps_un)
{
	if (!kps || !kps_un || !camera || n < 0 || ndist < 0 || ndist > 14 || (ndist > 0 && !dist)) return fail(ORBX_ERR_INVALID, "bad argument");
	if (ndist == 0 || dist[0] == 0.f)            // src/System.cc:155-159: dst = src
	{
		if (kps_un != kps) std::memcpy(kps_un, kps, sizeof(orbx_keypoint) * (size_t)n);
		return ORBX_OK;
	}
	std::string why;
	if (!device_ok(device, why)) return fail(ORBX_ERR_CUDA, why);
	if (n == 0) return ORBX_OK;
	CU(cudaSetDevice(device));
	DevBuf<orbx_keypoint> ds, dd;
	CU(ds.ensure(n)); CU(dd.ensure(n));
	CU(cudaMemcpy(ds.p, kps, sizeof(orbx_keypoint) * n, cudaMemcpyHostToDevice));
	const float cam4[4] = { camera->fx, camera->fy, camera->cx, camera->cy };
	orbx_launch_undistort(ds.p, dd.p, n, cam4, dist, ndist, 0);
	CU(cudaGetLastError());
	CU(cudaMemcpy(kps_un, dd.p, sizeof(orbx_keypoint) * n, cudaMemcpyDeviceToHost));
	return ORBX_OK;
}

orbx_status orbx_undistort_keypoints_device(const orbx_keypoint* d_kps, int n, const orbx_camera* camera, const float* dist, int ndist,
                                            orbx_keypoint* d_kps_un, void* stream)
{
	if (!d_kps || !d_kps_un || !camera || n < 0 || ndist < 0 || ndist > 14 || (ndist > 0 && !dist)) return fail(ORBX_ERR_INVALID, "bad argument");
	if (n == 0) return ORBX_OK;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	if (ndist == 0 || dist[0] == 0.f)            // src/System.cc:155-159: dst = src
	{
		if (d_kps_un != d_kps) CU(cudaMemcpyAsync(d_kps_un, d_kps, sizeof(orbx_keypoint) * (size_t)n, cudaMemcpyDeviceToDevice, st));
		return ORBX_OK;
	}
	const float cam4[4] = { camera->fx, camera->fy, camera->cx, camera->cy };
	orbx_launch_undistort(d_kps, d_kps_un, n, cam4, dist, ndist, st);
	CU(cudaGetLastError());
	return ORBX_OK;
}

orbx_status orbx_stereo_from_rgbd(int device, const orbx_keypoint* kps, const orbx_keypoint* kps_un, int n, const float* depth_map,
                                  int width, int height, size_t pitch, const orbx_camera* camera, float* uright, float* depth)
{
	if (!kps || !kps_un || !depth_map || !camera || !uright || !depth || n < 0 || width < 1 || height < 1) return fail(ORBX_ERR_INVALID, "bad argument");
	std::string why;
	if (!device_ok(device, why)) return fail(ORBX_ERR_CUDA, why);
	if (n == 0) return ORBX_OK;
	CU(cudaSetDevice(device));
	DevBuf<orbx_keypoint> dk, du; DevBuf<uint8_t> dm; DevBuf<float> dr, dz;
	const size_t dpitch = (size_t)width * 4;
	CU(dk.ensure(n)); CU(du.ensure(n)); CU(dm.ensure(dpitch * height)); CU(dr.ensure(n)); CU(dz.ensure(n));
	CU(cudaMemcpy(dk.p, kps, sizeof(orbx_keypoint) * n, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(du.p, kps_un, sizeof(orbx_keypoint) * n, cudaMemcpyHostToDevice));
	CU(cudaMemcpy2D(dm.p, dpitch, depth_map, pitch, dpitch, height, cudaMemcpyHostToDevice));
	orbx_launch_stereo_from_rgbd(dk.p, du.p, n, dm.p, (int64_t)dpitch, camera->bf, dr.p, dz.p, 0);
	CU(cudaGetLastError());
	CU(cudaMemcpy(uright, dr.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(depth, dz.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
	return ORBX_OK;
}

orbx_status orbx_distinctive_descriptors(int device, const uint8_t* desc, const int64_t* offsets, int nsets, int32_t* best)
{
	if (!desc || !offsets || !best || nsets < 0) return fail(ORBX_ERR_INVALID, "bad argument");
	std::string why;
	if (!device_ok(device, why)) return fail(ORBX_ERR_CUDA, why);
	if (nsets == 0) return ORBX_OK;
	for (int s = 0; s < nsets; s++)
		if (offsets[s + 1] < offsets[s]) return fail(ORBX_ERR_INVALID, "offsets must be non-decreasing");
	CU(cudaSetDevice(device));
	const int64_t rows = offsets[nsets];
	DevBuf<uint8_t> dd; DevBuf<int64_t> dof; DevBuf<int32_t> db;
	CU(dd.ensure(std::max<int64_t>(32 * rows, 32))); CU(dof.ensure(nsets + 1)); CU(db.ensure(nsets));
	if (rows) CU(cudaMemcpy(dd.p, desc, 32 * rows, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(dof.p, offsets, sizeof(int64_t) * (nsets + 1), cudaMemcpyHostToDevice));
	orbx_launch_distinctive(dd.p, dof.p, nsets, db.p, 0);
	CU(cudaGetLastError());
	CU(cudaMemcpy(best, db.p, sizeof(int32_t) * nsets, cudaMemcpyDeviceToHost));
	return ORBX_OK;
}

orbx_status orbx_debug_cos_sin(int device, uint32_t first_bits, int64_t n, float* cos_out, float* sin_out)
{
	if (!cos_out || !sin_out || n < 0) return fail(ORBX_ERR_INVALID, "bad argument");
	std::string why;
	if (!device_ok(device, why)) return fail(ORBX_ERR_CUDA, why);
	if (n == 0) return ORBX_OK;
	CU(cudaSetDevice(device));
	DevBuf<float> dc, ds;
	CU(dc.ensure((size_t)n)); CU(ds.ensure((size_t)n));
	orbx_launch_debug_cos_sin(first_bits, n, dc.p, ds.p, 0);
	CU(cudaGetLastError());
	CU(cudaMemcpy(cos_out, dc.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(sin_out, ds.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
	return ORBX_OK;
}

orbx_status orbx_measure_popc_peak(int device, double* popc_per_second)
{
	if (!popc_per_second) return fail(ORBX_ERR_INVALID, "null argument");
	std::string why;
	if (!device_ok(device, why)) return fail(ORBX_ERR_CUDA, why);
	*popc_per_second = orbx_popc_probe(device);
	CU(cudaGetLastError());
	return ORBX_OK;
}

}  // extern "C"
