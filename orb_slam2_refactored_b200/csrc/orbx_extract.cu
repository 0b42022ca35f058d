// sm_100a kernels of the ORB extractor: pyramid, FAST cells, quadtree selection, blur, orientation + rBRIEF.
// Reference behaviour reproduced (bit-exact): src/ORBextractor.cc:74-140, :402-693, :743-820 and the OpenCV
// primitives it calls. Integer math only on pixels; every float op that the reference
// performs is issued with an explicit round-to-nearest intrinsic so that nvcc cannot contract it into an FMA.
#include "orbx_internal.cuh"

#include <math.h>
#include <algorithm>
#include <stdlib.h>
#include <stdio.h>

#include "orbx_sort.cuh"

namespace {

__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

// 16-byte asynchronous global -> shared copies (LDGSTS): a tile is staged with one or two instructions per thread and
// all of them in flight at once, which is what hides DRAM latency for these small tiles. Both addresses 16-byte aligned.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all()
{
	asm volatile("cp.async.commit_group;\n" ::);
	asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier: one thread issues a whole 2-D tile load, the copy engine writes shared memory and
// completes the transaction count on the barrier; the CTA's threads only wait. SASS: UTMALDG / SYNCS.
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
	asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
	             ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2)
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
	const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
	unsigned done = 0;
	for (int spin = 0; spin < (1 << 24) && !done; spin++)
		asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
	if (!done) __trap();      // a lost transaction must not hang the device
}

__constant__ uint32_t c_inv20[72];   // c_inv20[n] = (1 << 20) / n + 1: floor(i / n) == (i * c_inv20[n]) >> 20 for n <= 64, i < 4096 (exhaustively checked)

// =====================================================================================================
// K0  gray_u8 — ConvertToGray (src/System.cc:122-137) = cv::cvtColor {RGB,BGR,RGBA,BGRA}2GRAY, 8-bit fixed point of
//     OpenCV 4.13.0: (R*9798 + G*19235 + B*3735 + 16384) >> 15. Writes level 0 of the pyramid directly (4 pixels per thread).
// =====================================================================================================
__global__ void __launch_bounds__(256) k_gray_to_l0(const uint8_t* __restrict__ src, int64_t spitch, int64_t sstride, int channels, int r_index,
                                                   uint8_t* __restrict__ dst, int64_t dpitch, int64_t dstride, int w, int h)
{
	const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, y = blockIdx.y * 4 + (threadIdx.x >> 6), f = blockIdx.z;
	if (x0 >= w || y >= h) return;
	const uint8_t* __restrict__ s = src + (int64_t)f * sstride + (int64_t)y * spitch + (int64_t)x0 * channels;
	uint32_t out = 0;
#pragma unroll
	for (int j = 0; j < 4; j++)
		if (x0 + j < w)
		{
			const uint8_t* p = s + j * channels;
			const int v = ((int)__ldg(p + r_index) * 9798 + (int)__ldg(p + 1) * 19235 + (int)__ldg(p + 2 - r_index) * 3735 + 16384) >> 15;
			out |= (uint32_t)v << (8 * j);
		}
	*reinterpret_cast<uint32_t*>(dst + (int64_t)f * dstride + (int64_t)y * dpitch + x0) = out;   // dpitch is a multiple of 128
}

// =====================================================================================================
// K0b remap_u8 — cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) of the rectification step in front of Extract
//     (Examples/Stereo/stereo_euroc.cc:100-101). The float maps are converted once on the host, exactly as OpenCV converts them per
//     call (cvRound(map * 32): integer position as two shorts, 5 + 5 fraction bits); the kernel is integer only: four taps, weights
//     (32-fx)(32-fy)*32 ... summing to 2^15, (v + 2^14) >> 15. Writes level 0 of the pyramid directly, 4 pixels per thread.
//     Traffic per output pixel: 8 B of table, ~1 B of source (the gathers of a warp hit neighbouring lines), 1 B written.
// =====================================================================================================
__global__ void __launch_bounds__(256) k_remap_to_l0(const uint8_t* __restrict__ src, int64_t spitch, int64_t sstride, int sw, int sh,
                                                    const int2* __restrict__ tab, uint8_t* __restrict__ dst, int64_t dpitch, int64_t dstride,
                                                    int w, int h)
{
	const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, y = blockIdx.y * 4 + (threadIdx.x >> 6), f = blockIdx.z;
	if (x0 >= w || y >= h) return;
	const uint8_t* __restrict__ s = src + (int64_t)f * sstride;
	const int2* __restrict__ t = tab + (int64_t)y * w + x0;
	uint32_t out = 0;
#pragma unroll
	for (int j = 0; j < 4; j++)
		if (x0 + j < w)
		{
			const int2 e = __ldg(t + j);
			const int ix = (int)(short)(e.x & 0xffff), iy = e.x >> 16, fx = e.y & 31, fy = e.y >> 5;
			int v00 = 0, v01 = 0, v10 = 0, v11 = 0;
			if ((unsigned)ix < (unsigned)(sw - 1) && (unsigned)iy < (unsigned)(sh - 1))
			{
				const uint8_t* p = s + (int64_t)iy * spitch + ix;
				v00 = __ldg(p); v01 = __ldg(p + 1); v10 = __ldg(p + spitch); v11 = __ldg(p + spitch + 1);
			}
			else
			{
				const bool x0in = (unsigned)ix < (unsigned)sw, x1in = (unsigned)(ix + 1) < (unsigned)sw;
				const bool y0in = (unsigned)iy < (unsigned)sh, y1in = (unsigned)(iy + 1) < (unsigned)sh;
				if (y0in && x0in) v00 = __ldg(s + (int64_t)iy * spitch + ix);
				if (y0in && x1in) v01 = __ldg(s + (int64_t)iy * spitch + ix + 1);
				if (y1in && x0in) v10 = __ldg(s + (int64_t)(iy + 1) * spitch + ix);
				if (y1in && x1in) v11 = __ldg(s + (int64_t)(iy + 1) * spitch + ix + 1);
			}
			const int v = (v00 * ((32 - fx) * (32 - fy) * 32) + v01 * (fx * (32 - fy) * 32) + v10 * ((32 - fx) * fy * 32) + v11 * (fx * fy * 32) +
			               (1 << 14)) >> 15;
			out |= (uint32_t)v << (8 * j);
		}
	*reinterpret_cast<uint32_t*>(dst + (int64_t)f * dstride + (int64_t)y * dpitch + x0) = out;   // dpitch is a multiple of 4
}

// =====================================================================================================
// K1  pyramid_resize_u8 — cv::resize INTER_LINEAR 8UC1 in OpenCV's fixed point for
//     ComputePyramid (src/ORBextractor.cc:455-470). Coefficient tables are built on the host with the
//     exact float/double operation order; the kernel is integer only. One thread = 4 output pixels.
// =====================================================================================================
#define PY_TW 128     // output tile: 128 columns x 128 rows per CTA; warp w owns rows [16w, 16w + 16), lane l owns columns [4l, 4l + 4)
#define PY_TH 128
#define PY_RW 16      // output rows per warp
#define PY_SRC 288    // source rows a tile may touch: 127 * scaleFactor + 2; scale factors up to 2.2 (checked on the host)
#define PY_SW 304     // source bytes per staged row: 15 (alignment) + 128 * scaleFactor + 2, rounded up to 16
// RW = output rows per warp (tile height 8 * RW): 16 for throughput, 4 for small batches, where a level has too few 128-row tiles to
// fill the GPU and the per-warp row walk is the launch's latency.
template <int RW>
__global__ void __launch_bounds__(256) k_pyramid_resize(const OrbxPlanDev P, const int level)
{
	constexpr int TH = 8 * RW;
	// The source rows/columns an output tile needs are staged in shared memory with 16-byte async copies (level buffers are padded:
	// pitch a multiple of 128, 256 spare bytes in front and behind). A thread then walks DOWN its 4 columns: the horizontal pass of a
	// source row (2 taps x 4 columns) is computed once and kept in registers while the 1-2 output rows that need it are produced, so
	// the column offsets and coefficients are loaded once per thread and a source row is interpolated ~1.25 instead of 2 times per
	// output row.
	extern __shared__ __align__(16) uint8_t stile[];
	const OrbxLevel& D = P.lv[level];
	const int sw = P.lv[level - 1].w, sh = P.lv[level - 1].h;
	const int f = blockIdx.z, tid = threadIdx.x;
	const int dx0 = blockIdx.x * PY_TW, dy0 = blockIdx.y * TH;
	const int dy_last = min(dy0 + TH, D.h) - 1, dx_last = min(dx0 + PY_TW, D.w) - 1;
	const uint8_t* __restrict__ src = orbx_level_ptr(P, f, level - 1);
	const int64_t sp = orbx_level_pitch(P, level - 1);
	uint8_t* __restrict__ dst = P.pyr + (int64_t)f * P.slab + D.offset;
	const int* __restrict__ yofs = P.yofs + D.ytab_base;
	const short2* __restrict__ ycoef = P.ycoef + D.ytab_base;
	const int* __restrict__ xofs = P.xofs + D.xtab_base;

	const int s_lo = __ldg(yofs + dy0);
	const int s_hi = min(__ldg(yofs + dy_last) + 1, sh - 1);
	const int nsrc = s_hi - s_lo + 1;                       // <= PY_SRC (host-checked)
	const int xa = __ldg(xofs + dx0) & ~15;                 // first staged source column
	const int xb = min(__ldg(xofs + dx_last) + 1, sw - 1);  // last needed source column
	const int nchunk = (xb - xa + 16) >> 4;                 // 16-byte chunks per row, <= PY_SW / 16 (host-checked)
	const int rs = nchunk * 16;                             // staged row stride
	{
		const uint8_t* __restrict__ g0 = src + (int64_t)s_lo * sp + xa;
		for (int i = tid; i < nsrc * nchunk; i += 256)
		{
			const int r = (int)(((uint32_t)i * c_inv20[nchunk]) >> 20), c = i - r * nchunk;
			cp_async16(stile + r * rs + c * 16, g0 + (int64_t)r * sp + c * 16);
		}
	}
	const int lane = tid & 31, warp = tid >> 5;
	int x0r[4], x1r[4], a0[4], a1[4];
#pragma unroll
	for (int j = 0; j < 4; j++)
	{
		const int dx = min(dx0 + 4 * lane + j, D.w - 1);   // columns past the edge repeat the last one; they land in row padding
		const int sx = __ldg(xofs + dx);
		const short2 a = __ldg(P.xcoef + D.xtab_base + dx);
		x0r[j] = sx - xa; x1r[j] = min(sx + 1, sw - 1) - xa;
		a0[j] = a.x; a1[j] = a.y;
	}
	// the warp's rows: lane k holds source row and coefficients of row k, broadcast by shuffle in the loop
	const int wy0 = dy0 + warp * RW;
	int my_sy = 0, my_b = 0;
	if (lane < RW)
	{
		const int dy = min(wy0 + lane, D.h - 1);
		my_sy = __ldg(yofs + dy);
		const short2 b = __ldg(ycoef + dy);
		my_b = (int)(uint16_t)b.x | ((int)b.y << 16);
	}
	cp_async_wait_all();
	__syncthreads();
	if (wy0 >= D.h) return;
	const bool store = dx0 + 4 * lane < D.w;

	// horizontal pass of staged row r for this thread's 4 columns: (s[x0] * a0 + s[x1] * a1) >> 4
	auto hrow = [&](int r, int (&h)[4]) {
		const uint8_t* __restrict__ row = stile + (r - s_lo) * rs;
#pragma unroll
		for (int j = 0; j < 4; j++) h[j] = ((int)row[x0r[j]] * a0[j] + (int)row[x1r[j]] * a1[j]) >> 4;
	};
	int rc = -2, h0[4], h1[4];          // h0 = row rc, h1 = row min(rc + 1, sh - 1)
	const int nrows = min(RW, D.h - wy0);
	for (int k = 0; k < nrows; k++)
	{
		const int r = __shfl_sync(0xffffffffu, my_sy, k), bw = __shfl_sync(0xffffffffu, my_b, k);
		if (r != rc)
		{
			if (r == rc + 1)
			{
#pragma unroll
				for (int j = 0; j < 4; j++) h0[j] = h1[j];
			}
			else hrow(r, h0);
			hrow(min(r + 1, sh - 1), h1);      // at the last source row both taps are that row, as in the reference table
			rc = r;
		}
		const int b0 = (int)(short)(bw & 0xffff), b1 = bw >> 16;
		uint32_t out = 0;
#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			const int v = (((b0 * h0[j]) >> 16) + ((b1 * h1[j]) >> 16) + 2) >> 2;   // coefficients sum to 2048: v in [0, 255]
			out |= (uint32_t)v << (8 * j);
		}
		if (store) *reinterpret_cast<uint32_t*>(dst + (int64_t)(wy0 + k) * D.pitch + dx0 + 4 * lane) = out;   // pitch is a multiple of 128: in-row padding absorbs the tail
	}
}

// =====================================================================================================
// K2  fast9_cell — DetectFAST (src/ORBextractor.cc:489-540) with cv::FAST(..., nms = true) semantics
//     (SURVEY App. A.4). One warp per ~30 px cell: the cell view (cell + 6 px) is staged in shared memory,
//     the threshold-independent arc score S is computed for the pixels that pass a cheap 4-pair rejection
//     (every arc of 9 contains one pixel of each opposite pair (k, k+8), so min_arc(ring) <= max(ring_k, ring_k+8):
//     S_bright <= min_k max(ring_k, ring_k+8) - centre, S_dark <= centre - max_k min(ring_k, ring_k+8), k = 0, 2, 4, 6),
//     local maxima are found once, and the iniTh -> minTh retry is a second pass over the pixels in between.
//     Candidates are emitted row-major into the cell's private slot range, so DetectFAST's cell-major /
//     row-major push_back order is reproduced without atomics on global memory.
// =====================================================================================================
// Cell view staging. A cell is < 60 px per side (CELL_SIZE 30: ceil(roi / floor(roi / 30)) < 60), its view 6 px more, and the TMA box
// starts 16-byte aligned: 65 + 15 = 80 bytes per row always suffice. 80 bytes = 20 words also puts 8 consecutive rows on 8 different
// bank phases (64 would give 2, 96 gives 4): the exact-score pass gathers ring pixels of arbitrary (row, column) per lane, and ncu
// showed the cell kernel bound by shared-memory wavefronts (45 % of them bank conflicts at a 64-byte stride).
#define FT_TS 80
#define FT_TH 66

//@phase exact arc score (16 ring loads, packed min/max network)
template <int TS>
__device__ __forceinline__ int arc_score_packed_t(const uint8_t* __restrict__ c)
{
	// Ring offsets inside the shared tile, OpenCV order; compile-time so every load is [base + imm].
	constexpr int R[16] = { 3 * TS, 3 * TS + 1, 2 * TS + 2, TS + 3, 3, -TS + 3, -2 * TS + 2, -3 * TS + 1,
	                        -3 * TS, -3 * TS - 1, -2 * TS - 2, -TS - 3, -3, TS - 3, 2 * TS - 2, 3 * TS - 1 };
	// v[k] = ring_k | (255 - ring_k) << 16 (one IMAD); max over an arc of 9 in both halves, then min over the 16 arcs:
	// lo = min_arcs max_arc ring, hi = 255 - max_arcs min_arc ring. VIMNMX3.U16x2 is a 3-input packed max/min.
	uint32_t v[16];
#pragma unroll
	for (int k = 0; k < 16; k++)
		v[k] = (uint32_t)c[R[k]] * 0xFFFF0001u + 0x00FF0000u;
	uint32_t m3[16];
#pragma unroll
	for (int k = 0; k < 16; k++)
		m3[k] = __vimax3_u16x2(v[k], v[(k + 1) & 15], v[(k + 2) & 15]);
	uint32_t m9[16];
#pragma unroll
	for (int k = 0; k < 16; k++)
		m9[k] = __vimax3_u16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
	uint32_t a = __vimin3_u16x2(m9[0], m9[1], m9[2]);
	uint32_t b = __vimin3_u16x2(m9[3], m9[4], m9[5]);
	uint32_t d = __vimin3_u16x2(m9[6], m9[7], m9[8]);
	uint32_t e = __vimin3_u16x2(m9[9], m9[10], m9[11]);
	uint32_t g = __vimin3_u16x2(m9[12], m9[13], m9[14]);
	a = __vimin3_u16x2(a, b, d);
	e = __vimin3_u16x2(e, g, m9[15]);
	a = __vminu2(a, e);
	const int centre = c[0];
	const int dark = centre - (int)(a & 0xffffu);            // max_arcs min_arc (centre - ring)
	const int bright = (255 - (int)(a >> 16)) - centre;      // max_arcs min_arc (ring - centre)
	return max(dark, bright);
}
__device__ __forceinline__ int arc_score_packed(const uint8_t* __restrict__ c) { return arc_score_packed_t<FT_TS>(c); }

//@end
#include "orbx_strip.cuh"

// =====================================================================================================
// K3+K4  quadtree_select — QuadTreeSuppression + QTreeNode::divide (src/ORBextractor.cc:402-453, :542-693)
//     in pass form (executable spec: oracle/orb_oracle.cc quadtree()). One CTA per
//     (level, frame). A node owns a contiguous segment of packed candidates; dividing a node is a stable
//     4-way partition of its segment into the other ping-pong buffer (one warp per node). The std::list is
//     an array rebuilt every pass: [children of this pass in reverse push order] ++ [old list minus divided
//     nodes]. Phase 2's std::sort (unstable, libstdc++ introsort) is replayed step for step by one thread,
//     because its order among equal sizes decides which nodes are split before the quota break (:666-667).
// =====================================================================================================
#define QT_THREADS 256
#define QT_NS qt256
#include "orbx_quadtree.cuh"
#undef QT_THREADS
#undef QT_NS
#define QT_THREADS 128
#define QT_NS qt128
#include "orbx_quadtree.cuh"
#undef QT_THREADS
#undef QT_NS
#define QT_THREADS 512
#define QT_NS qt512
#include "orbx_quadtree.cuh"
#undef QT_THREADS
#undef QT_NS
#define QT_THREADS 96
#define QT_NS qt96
#include "orbx_quadtree.cuh"
#undef QT_THREADS
#undef QT_NS
// =====================================================================================================
// K5+K7  orient_describe — IC_Angle (src/ORBextractor.cc:74-101) on the un-blurred level, then
//     ComputeOrbDescriptor (:103-140) on the blurred level, then the keypoint record of Extract (:768-773,
//     :811-815). The kernel is k_orient_describe2 (orbx_describe.cuh): eight keypoints per warp. Float path pinned per SURVEY H2 / App. A.6-A.7.
// =====================================================================================================
// Lookup tables of the orientation/descriptor kernel live in global memory (L1-resident): lanes read different entries, which
// would serialise on the constant cache (measured: 28 % of the kernel) but is one coalesced request here.
__device__ float4 g_patf[256];                  // pair 8*byte + bit -> (x0,y0,x1,y1) as floats, stored at [bit][byte]: a warp reads 512 contiguous bytes
__device__ uint2 g_mom[8 * 16];                 // [k][|v|]: .x = ones mask, .y = column offsets u (s8) of window word k of disc row v
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
	const float R2D = (float)(180.0 / 3.14159265358979323846);
	const float p1 = __fmul_rn(0.9997878412794807f, R2D), p3 = __fmul_rn(-0.3258083974640975f, R2D);
	const float p5 = __fmul_rn(0.1555786518463281f, R2D), p7 = __fmul_rn(-0.04432655554792128f, R2D);
	const float eps = (float)2.2204460492503131e-16;
	const float ax = fabsf(x), ay = fabsf(y);
	float a, c, c2;
	if (ax >= ay)
	{
		c = __fdiv_rn(ay, __fadd_rn(ax, eps));
		c2 = __fmul_rn(c, c);
		a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
	}
	else
	{
		c = __fdiv_rn(ax, __fadd_rn(ay, eps));
		c2 = __fmul_rn(c, c);
		a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
	}
	if (x < 0.f) a = __fsub_rn(180.f, a);
	if (y < 0.f) a = __fsub_rn(360.f, a);
	return a;
}

// The rotation terms of ComputeOrbDescriptor (src/ORBextractor.cc:105-107): angle * factorPI in float, then the DOUBLE cos / sin rounded to
// float. tests/test_gpu_cos_sin_sweep.py evaluates this for every float angle in [0, 360) against glibc on the host.
__device__ __forceinline__ void orb_cos_sin(float angle_deg, float& a, float& b)
{
	const float factorPI = (float)(3.1415926535897932384626433832795 / (double)180.f);
	const float arad = __fmul_rn(angle_deg, factorPI);
	a = __double2float_rn(cos((double)arad));
	b = __double2float_rn(sin((double)arad));
}
__global__ void __launch_bounds__(256) k_debug_cos_sin(uint32_t first_bits, int64_t n, float* __restrict__ c, float* __restrict__ s)
{
	const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	float a, b;
	orb_cos_sin(__uint_as_float(first_bits + (uint32_t)i), a, b);
	c[i] = a; s[i] = b;
}

__device__ __forceinline__ int dp4a_u8_s8(uint32_t a, uint32_t b, int c)
{
	int d;
	asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}

#include "orbx_describe.cuh"

}  // namespace

// =====================================================================================================
// launchers
// =====================================================================================================
cudaError_t orbx_upload_pattern()
{
	// umax_ of ORBextractor::Init (src/ORBextractor.cc:705-718)
	int umax[ORBX_HALF_PATCH + 1];
	const int vmax = (int)floor(ORBX_HALF_PATCH * sqrt(2.) / 2 + 1);
	const int vmin = (int)ceil(ORBX_HALF_PATCH * sqrt(2.) / 2);
	for (int v = 0; v <= vmax; ++v)
		umax[v] = (int)lrint(sqrt((double)(ORBX_HALF_PATCH * ORBX_HALF_PATCH - v * v)));
	for (int v = ORBX_HALF_PATCH, v0 = 0; v >= vmin; --v)
	{
		while (umax[v0] == umax[v0 + 1]) ++v0;
		umax[v] = v0;
		++v0;
	}
	static const signed char pattern[1024] = {
#include "orb_pattern.inc"
	};
	float4 patf[256];
	for (int p = 0; p < 256; p++)
		patf[(p & 7) * 32 + (p >> 3)] = make_float4((float)pattern[4 * p], (float)pattern[4 * p + 1], (float)pattern[4 * p + 2], (float)pattern[4 * p + 3]);
	uint2 mom[8 * 16];
	for (int k = 0; k < 8; k++)
		for (int av = 0; av < 16; av++)
		{
			// window word k holds columns u = 4k - 16 .. 4k - 13 of a disc row; inside the disc iff |u| <= umax[|v|]
			uint32_t ones = 0, us = 0;
			for (int j = 0; j < 4; j++)
			{
				const int u = 4 * k + j - 16;
				if (abs(u) <= umax[av]) { ones |= 1u << (8 * j); us |= (uint32_t)(uint8_t)(signed char)u << (8 * j); }
			}
			mom[k * 16 + av] = make_uint2(ones, us);
		}
	uint32_t inv20[72];
	inv20[0] = 0;
	for (int n = 1; n < 72; n++) inv20[n] = (1u << 20) / (uint32_t)n + 1u;
	cudaError_t e;
	if ((e = cudaMemcpyToSymbol(c_inv20, inv20, sizeof(inv20))) != cudaSuccess) return e;
	if ((e = cudaMemcpyToSymbol(g_patf, patf, sizeof(patf))) != cudaSuccess) return e;
	if ((e = cudaMemcpyToSymbol(g_mom, mom, sizeof(mom))) != cudaSuccess) return e;
	return cudaSuccess;
}

// debug: one thread writes %globaltimer into a slot (ORBX_TRACE=1, tools/latency_trace.py): a timeline of the streams of a one-frame call
__global__ void k_stamp(unsigned long long* slot)
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	*slot = t;
}
void orbx_launch_stamp(unsigned long long* slot, cudaStream_t st) { k_stamp<<<1, 1, 0, st>>>(slot); }

void orbx_launch_debug_cos_sin(uint32_t first_bits, int64_t n, float* d_cos, float* d_sin, cudaStream_t st)
{
	k_debug_cos_sin<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(first_bits, n, d_cos, d_sin);
}

void orbx_launch_gray(const uint8_t* src, int64_t spitch, int64_t sstride, int channels, int rgb, uint8_t* dst, int64_t dpitch, int64_t dstride,
                      int w, int h, int frames, cudaStream_t st)
{
	dim3 grid((w + 255) / 256, (h + 3) / 4, frames);
	k_gray_to_l0<<<grid, 256, 0, st>>>(src, spitch, sstride, channels, rgb ? 0 : 2, dst, dpitch, dstride, w, h);
}

void orbx_launch_remap(const uint8_t* src, int64_t spitch, int64_t sstride, int sw, int sh, const int2* tab, uint8_t* dst, int64_t dpitch,
                       int64_t dstride, int w, int h, int frames, cudaStream_t st)
{
	dim3 grid((w + 255) / 256, (h + 3) / 4, frames);
	k_remap_to_l0<<<grid, 256, 0, st>>>(src, spitch, sstride, sw, sh, tab, dst, dpitch, dstride, w, h);
}

// ---- tuning knobs, read once per process
static int env_int(const char* name, int dflt)
{
	const char* e = getenv(name);
	return e ? atoi(e) : dflt;
}
int orbx_strip_rows(int which)
{
	static const int v = env_int("ORBX_STRIP_TH", 32) == 64 ? 64 : env_int("ORBX_STRIP_TH", 32) == 16 ? 16 : 32;
	// which = 2: the blur's throughput tiles. 64 rows: one tile header and one 6-row warm-up per 64 instead of per 32 rows (blur 0.254 ->
	// 0.244 ms per 512 frames); the FAST bound is no faster with them (0.909 vs 0.903 ms: its tiles take 64 registers and 11 KB each)
	static const int vb = env_int("ORBX_BLUR_TH", 64) == 32 ? 32 : 64;
	return which == 1 ? 8 : which == 2 ? std::max(v, vb) : v;
}
int orbx_strip_box_w() { return ST_BW; }
int orbx_pyramid_strip_rows(int which)
{
	static const int v = env_int("ORBX_PYR_TH", 32) == 16 ? 16 : 32;
	static const int vs = env_int("ORBX_PYR_TH_SMALL", 4) == 8 ? 8 : 4;     // tile rows of one-frame launches: the levels are a dependent chain, and a 4-row tile is done sooner (pyramid 30 -> 26 us, 0.141 -> 0.138 ms per call at C1)
	return which ? vs : v;
}

#define QT_SMEM_MAX (200 * 1024)
// Function attributes are per device and cost a driver call each: set once per orbx_create, not per launch.
cudaError_t orbx_kernels_init()
{
	cudaError_t e = cudaSuccess;
	auto set = [&](auto fn, int bytes) {
		if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	};
	set(k_pyramid_resize<PY_RW>, PY_SRC * PY_SW);
	set(k_pyramid_resize<4>, PY_SRC * PY_SW);
	set(k_pyramid_strip<32>, 100 * 1024);
	set(k_pyramid_strip<16>, 100 * 1024);
	set(k_pyramid_strip<8>, 100 * 1024);
	set(k_pyramid_all<32>, 100 * 1024); set(k_pyramid_all<16>, 100 * 1024); set(k_pyramid_all<8>, 100 * 1024); set(k_pyramid_all<4>, 100 * 1024);
	set(k_pyramid_strip<4>, 100 * 1024);
	set(k_fast_cells2, 64 * 1024); set(k_fast_cells2_overflow, 64 * 1024);
	set(k_level_strip<8, true, false>, 64 * 1024); set(k_level_strip<8, false, true>, 64 * 1024);
	set(k_level_strip<16, true, false>, 64 * 1024); set(k_level_strip<16, false, true>, 64 * 1024);
	set(k_level_strip<32, true, false>, 64 * 1024); set(k_level_strip<32, false, true>, 64 * 1024);
	set(k_level_strip<64, true, false>, 64 * 1024); set(k_level_strip<64, false, true>, 64 * 1024);
	set(qt128::k_quadtree<false>, QT_SMEM_MAX); set(qt256::k_quadtree<false>, QT_SMEM_MAX); set(qt96::k_quadtree<false>, QT_SMEM_MAX);
	set(qt256::k_quadtree<true>, QT_SMEM_MAX); set(qt512::k_quadtree<true>, QT_SMEM_MAX);
	set(k_orient_describe2<8>, OD2_SMEM); set(k_orient_describe2<2>, OD2_SMEM); set(k_orient_describe2<16>, OD2_SMEM);
	return e;
}

void orbx_launch_pyramid(const OrbxPlanDev& P, const OrbxPyrMaps pmaps[2], int level, cudaStream_t st)
{
	const OrbxLevel& D = P.lv[level];
	const bool small_batch = P.frames <= ORBX_SMALL_BATCH;
	const int which = small_batch ? 1 : 0;
	if (D.py_bw[which] > 0)
	{
		// strip kernel: one warp per 128 x TH output tile, source rectangle by one TMA box
		const int th = orbx_pyramid_strip_rows(which), bw = D.py_bw[which], bh = D.py_bh[which];
		dim3 grid((D.w + ST_TW - 1) / ST_TW, (D.h + th - 1) / th, P.frames);
		const int smem = ((bw * bh + 127) & ~127) + 16;
		if (th == 4) k_pyramid_strip<4><<<grid, 32, smem, st>>>(P, pmaps[which], level, bw, bh);
		else if (th == 8) k_pyramid_strip<8><<<grid, 32, smem, st>>>(P, pmaps[which], level, bw, bh);
		else if (th == 16) k_pyramid_strip<16><<<grid, 32, smem, st>>>(P, pmaps[which], level, bw, bh);
		else k_pyramid_strip<32><<<grid, 32, smem, st>>>(P, pmaps[which], level, bw, bh);
		return;
	}
	const int th = small_batch ? 32 : PY_TH;
	dim3 grid((D.w + PY_TW - 1) / PY_TW, (D.h + th - 1) / th, P.frames);
	// dynamic shared memory: staged source rows; sized per level by the host (P.lv[level].py_smem), up to PY_SRC * PY_SW = 86 KB.
	// py_smem is sized for the 128-row tile: an upper bound for the 32-row one
	if (small_batch) k_pyramid_resize<4><<<grid, 256, P.lv[level].py_smem, st>>>(P, level);
	else k_pyramid_resize<PY_RW><<<grid, 256, P.lv[level].py_smem, st>>>(P, level);
}

// levels [s0, s1) (s0 >= 1; s1 <= 0: up to the last level)
cudaError_t orbx_launch_pyramid_all(const OrbxPlanDev& P, const OrbxPyrMaps pmaps[2], cudaStream_t st, int s0, int s1)
{
	if (s1 <= 0 || s1 > P.nlevels) s1 = P.nlevels;
	s0 = std::max(s0, 1);
	if (s0 >= s1) return cudaSuccess;
	// One launch for a frame at a time (C1: 0.201 -> 0.193 ms per Extract call). For batches it was measured both ways: the pyramid stage
	// alone gains 3 % (0.309 -> 0.300 ms per 512 frames), but the step as callers run it — two half-batches on two streams — loses 2 %
	// (256.0 k -> 251.3 k frames/s): seven short launches per lane interleave better with the other lane's kernels than one long one.
	static const int merged = env_int("ORBX_PYR_ONE", -1);     // tuning knob: 0 = one launch per level, 1 = always one launch
	const int which = P.frames <= ORBX_SMALL_BATCH ? 1 : 0;
	bool ok = (merged < 0 ? which == 1 : merged != 0) && P.nlevels > 2 && P.pyr_done;
	for (int s = 1; s < P.nlevels && ok; s++) ok = P.lv[s].py_bw[which] > 0;
	if (!ok)
	{
		for (int s = s0; s < s1; s++) orbx_launch_pyramid(P, pmaps, s, st);
		return cudaSuccess;
	}
	const int th = orbx_pyramid_strip_rows(which);
	OrbxPyrTiles T = {};
	T.which = which;
	int smem = 0;
	for (int s = 1; s < P.nlevels; s++)
	{
		const OrbxLevel& D = P.lv[s];
		T.tx[s] = (D.w + ST_TW - 1) / ST_TW;
		T.base[s + 1] = T.base[s] + T.tx[s] * ((D.h + th - 1) / th);
		smem = std::max(smem, ((D.py_bw[which] * D.py_bh[which] + 127) & ~127) + 16);
	}
	for (int s = P.nlevels; s < ORBX_MAX_LEVELS; s++) T.base[s + 1] = T.base[s];
	// the completion counts are reset by the launch that starts the pyramid; a later launch over the next levels finds its source level's
	// count complete (stream order)
	if (s0 == 1)
	{
		const cudaError_t e = cudaMemsetAsync(P.pyr_done + (int64_t)P.frame0 * ORBX_MAX_LEVELS, 0, sizeof(int) * ORBX_MAX_LEVELS * (size_t)P.frames, st);
		if (e != cudaSuccess) return e;
	}
	T.first = T.base[s0];
	const unsigned grid = (unsigned)(T.base[s1] - T.base[s0]) * (unsigned)P.frames;
	if (th == 4) k_pyramid_all<4><<<grid, 32, smem, st>>>(P, pmaps[which], T, P.pyr_done);
	else if (th == 8) k_pyramid_all<8><<<grid, 32, smem, st>>>(P, pmaps[which], T, P.pyr_done);
	else if (th == 16) k_pyramid_all<16><<<grid, 32, smem, st>>>(P, pmaps[which], T, P.pyr_done);
	else k_pyramid_all<32><<<grid, 32, smem, st>>>(P, pmaps[which], T, P.pyr_done);
	return cudaSuccess;
}

int orbx_fast_tile_stride() { return FT_TS; }
int orbx_fast_tile_rows() { return FT_TH; }

static OrbxStripTiles strip_tiles(const OrbxPlanDev& P, int th, bool fast)
{
	OrbxStripTiles T = {};
	T.one = 1;
	T.base[0] = 0;
	// the FAST bound is only read inside the cells' interiors [minx + 3, maxx - 3) x [miny + 3, maxy - 3) (minx = miny = 16 on every level):
	// its tile grid starts there (columns: at the 16-byte aligned column below) instead of at the image corner: 17 % fewer tiles at VGA
	T.xorg = fast ? ORBX_BORDER : 0; T.yorg = fast ? ORBX_BORDER + 3 : 0;
	for (int s = 0; s < P.nlevels; s++)
	{
		const OrbxLevel& L = P.lv[s];
		const int cols = fast ? L.maxx - 3 - T.xorg : L.w, rows = fast ? L.maxy - 3 - T.yorg : L.h;
		T.tx[s] = (cols + ST_TW - 1) / ST_TW;
		T.base[s + 1] = T.base[s] + T.tx[s] * ((rows + th - 1) / th);
	}
	for (int s = P.nlevels; s < ORBX_MAX_LEVELS; s++) { T.tx[s] = 1; T.base[s + 1] = T.base[s]; }
	return T;
}

// one launch over the tiles of all levels: blur (mode 1) or dense FAST bound (mode 2). A fused form was measured: 110 registers, and both
// halves are issue-bound by themselves (ncu: 85 % / 80 % issue), so one kernel doing both was slower than the two launches (1.25-1.29 vs
// 0.98 + 0.25 ms per 512 frames).
static void launch_strip(const OrbxPlanDev& P, const OrbxStripMaps smaps_both[3], int mode, cudaStream_t st, int s0 = 0, int s1 = 0)
{
	if (s1 <= 0 || s1 > P.nlevels) s1 = P.nlevels;
	if (s0 >= s1) return;
	const int which = P.frames <= ORBX_SMALL_BATCH ? 1 : mode == 1 ? 2 : 0;
	const OrbxStripMaps& smaps = smaps_both[which];
	const int th = orbx_strip_rows(which);
	OrbxStripTiles T = strip_tiles(P, th, mode == 2);
	T.first = T.base[s0];
	dim3 grid(T.base[s1] - T.base[s0], P.frames);
	if (grid.x == 0) return;
	const int smem = st_tile_bytes(th) + 16;
#define ORBX_STRIP_CASE(TH_)                                                                                      \
	if (mode == 1) k_level_strip<TH_, true, false><<<grid, 32, smem, st>>>(P, smaps, T);                           \
	else k_level_strip<TH_, false, true><<<grid, 32, smem, st>>>(P, smaps, T);
	if (th == 8) { ORBX_STRIP_CASE(8) }
	else if (th == 16) { ORBX_STRIP_CASE(16) }
	else if (th == 64) { ORBX_STRIP_CASE(64) }
	else { ORBX_STRIP_CASE(32) }
#undef ORBX_STRIP_CASE
}

// per-warp shared memory of the cell kernels, sized by the largest cell of the plan
static void cell_extents(const OrbxPlanDev& P, const OrbxTmaMaps& maps, int& rows, int& maxrw, int& maxrh)
{
	rows = 0; maxrw = 0; maxrh = 0;
	for (int s = 0; s < P.nlevels; s++)
	{
		rows = std::max(rows, maps.box_h[s]);
		maxrw = std::max(maxrw, P.lv[s].cellw);
		maxrh = std::max(maxrh, P.lv[s].cellh);
	}
}

static void launch_cells2(const OrbxPlanDev& P, const OrbxTmaMaps& maps, cudaStream_t st, int s0 = 0, int s1 = 0)
{
	if (s1 <= 0 || s1 > P.nlevels) s1 = P.nlevels;
	if (s0 >= s1) return;
	const int cell0 = P.lv[s0].cell_base, cell1 = s1 < P.nlevels ? P.lv[s1].cell_base : P.cells_per_frame;
	if (cell1 <= cell0) return;
	int rows, maxrw, maxrh;
	cell_extents(P, maps, rows, maxrw, maxrh);
	// tile | score (1 px zero border) | list of pixels to score | survivor bitmap | mbarrier
	const int ts = FT_TS;
	auto layout = [&](int list_cap) {
		OrbxCellLayout Y;
		Y.score_stride = (maxrw + 2 + 7) & ~7;
		Y.off_score = (rows * ts + 15) & ~15;
		Y.off_list = (Y.off_score + (maxrh + 2) * Y.score_stride + 15) & ~15;
		Y.list_cap = list_cap;
		Y.off_bm = (Y.off_list + list_cap * 2 + 15) & ~15;
		Y.off_bar = Y.off_bm + 8 * maxrh;
		Y.warp_bytes = (Y.off_bar + 8 + 127) & ~127;
		return Y;
	};
	const int full = maxrw * maxrh;
	dim3 grid(cell1 - cell0, P.frames);
	// Throughput launches run with the short list (32 instead of 26 resident warps) and leave the rare cell that flags more pixels to a
	// second, tiny launch; a frame at a time keeps the full list and the single launch (the GPU is far from full there).
	static const int short_list = env_int("ORBX_SHORT_LIST", 1);      // tuning knob
	if (short_list && P.frames > ORBX_SMALL_BATCH && full > FT_LIST_CAP && P.ovf_list)
	{
		cudaMemsetAsync(P.ovf_count, 0, sizeof(int), st);
		const OrbxCellLayout Y = layout(FT_LIST_CAP);
		k_fast_cells2<<<grid, 32, Y.warp_bytes, st>>>(P, maps, Y, cell0);
		const OrbxCellLayout Yf = layout(full);
		k_fast_cells2_overflow<<<148, 32, Yf.warp_bytes, st>>>(P, maps, Yf);
		return;
	}
	const OrbxCellLayout Y = layout(full);
	k_fast_cells2<<<grid, 32, Y.warp_bytes, st>>>(P, maps, Y, cell0);
}

void orbx_launch_fast(const OrbxPlanDev& P, const OrbxTmaMaps& maps, const OrbxStripMaps smaps[3], cudaStream_t st, int part, int s0, int s1)
{
	// (Two cells per warp with one merged candidate list — 15 % fewer instructions per cell — was measured: 1.13 vs 0.96 ms per 512
	// frames. The doubled shared memory per warp halves the resident warps, and this kernel lives on latency hiding.)
	if (part != 2) launch_strip(P, smaps, 2, st, s0, s1);
	if (part != 1) launch_cells2(P, maps, st, s0, s1);
}

int orbx_pyramid_tile_rows() { return PY_TH; }
int orbx_pyramid_max_src_rows() { return PY_SRC; }
int orbx_pyramid_tile_cols() { return PY_TW; }
int orbx_pyramid_max_src_bytes() { return PY_SW; }

size_t orbx_quadtree_smem(int node_cap, bool big)
{
	// listA, listB (16 B), childcnt (16 B, its first half doubles as the sort items), proc, pbase (4 B each), gone (1 B); the CTA-parallel
	// sort of the BIG variant adds its leaf ranges (4 B x cap/2) and 2 segment queues (12 B x cap/16). 59 B per node: the 1748 nodes of a 4K
	// level 0 take 104 KB, so TWO 512-thread CTAs share an SM (at 65 B they did not)
	const size_t base = (size_t)node_cap * (16 + 16 + 16 + 4 + 4 + 1) + 64;
	return big ? base + 4 * (((size_t)node_cap / 2 + 2) & ~(size_t)1) + 2 * 12 * ((size_t)node_cap / 16 + 4) + 16 : base;
}

void orbx_launch_quadtree(const OrbxPlanDev& P, int* cell_off, cudaStream_t st, int s0, int s1)
{
	if (s1 <= 0 || s1 > P.nlevels) s1 = P.nlevels;
	if (s0 >= s1) return;
	const size_t smem = orbx_quadtree_smem(P.node_cap, true), smem_plain = orbx_quadtree_smem(P.node_cap, false);
	dim3 grid(P.frames, s1 - s0);
	// Two variants: BIG replays std::sort with the whole CTA and partitions large nodes block-wide. It costs registers (61 vs 40),
	// so it pays where one CTA's latency is what the launch waits for: levels that keep more than ~1000 keypoints (4K plans), and
	// small batches (a frame at a time, as Tracking calls Extract), where the GPU is far from full anyway.
	static const int force_big = getenv("ORBX_QT_BIG") ? atoi(getenv("ORBX_QT_BIG")) : -1;   // tuning knob
	const bool large_plan = P.node_cap > 1024, small_batch = P.frames <= 16;
	static unsigned long long* dbg = nullptr;
	static const bool want_dbg = getenv("ORBX_QT_STAMPS") != nullptr;
	if (want_dbg && !dbg) { cudaMalloc(&dbg, 64 * 8); }
	if (want_dbg) cudaMemsetAsync(dbg, 0, 64 * 8, st);
	if (force_big >= 0 ? force_big != 0 : (large_plan || small_batch))
	{
		// 4K-class plans keep > 1000 keypoints per level: the node lists need > 100 KB of shared memory, one CTA per SM, so a bigger CTA
		// costs no occupancy and its block-wide partition and parallel sort rounds use every warp
		// (measured at 3840x2160 / 8000 kp, 32 frames: 256 thr 63.8 us per frame, 512 thr 53.7, 1024 thr 66.2)
		static const int big_threads = getenv("ORBX_QT_THREADS") ? atoi(getenv("ORBX_QT_THREADS")) : 512;   // tuning knob: 256 or 512
		// a frame at a time with more than ~300 keypoints per level (KITTI: 2000 per frame): the divide passes have more nodes than a
		// 256-thread CTA has warps (one Extract call 0.329 -> 0.317 ms; at 1000 keypoints per frame no difference)
		static const int small_threads_env = getenv("ORBX_QT_SMALL_THREADS") ? atoi(getenv("ORBX_QT_SMALL_THREADS")) : 0;   // tuning knob: 256 or 512
		const int small_threads = small_threads_env ? small_threads_env : (P.node_cap > 320 ? 512 : 256);
		// a frame at a time: room for the level's candidates (both ping-pong segments) in shared memory, see the kernel. 6144 per segment
		// hold level 0 of a VGA or KITTI frame of ordinary texture; a level with more stays in global memory. Tuning knob ORBX_QT_SMEM_CAND.
		static const int smem_cand_env = getenv("ORBX_QT_SMEM_CAND") ? atoi(getenv("ORBX_QT_SMEM_CAND")) : 6144;
		int smem_cand = small_batch ? std::max(0, smem_cand_env) & ~3 : 0;
		while (smem_cand > 0 && smem + 16 + 8 * (size_t)smem_cand > QT_SMEM_MAX) smem_cand /= 2;
		smem_cand &= ~3;
		const size_t smem_all = smem + (smem_cand ? 16 + 8 * (size_t)smem_cand : 0);
		if ((large_plan && !small_batch && big_threads == 512) || (small_batch && small_threads == 512))
		{
			qt512::k_quadtree<true><<<grid, 512, smem_all, st>>>(P, cell_off, small_batch ? 1024 : 2048, small_batch ? 48 : 512, want_dbg ? dbg : nullptr, s0, smem_cand);
		}
		else
		{
			qt256::k_quadtree<true><<<grid, 256, smem_all, st>>>(P, cell_off, small_batch ? 1024 : 2048, small_batch ? 48 : 512, want_dbg ? dbg : nullptr, s0, smem_cand);
		}
	}
	else
	{
		// launches of >= 256 frames (>= 2048 CTAs): small CTAs, so that more of them share an SM while others are in their serial phases;
		// below that the SMs are not full and the CTA's own speed counts. Measured frames/s with 128 / 256 threads, one launch per
		// stage: 32 frames 83.2 k / 89.0 k, 128 frames 136.8 k / 142.3 k, 256 frames 165.8 k / 162.9 k, 512 frames 177.4 k / 174.6 k.
		// 96 threads from 512 frames (>= 4096 CTAs, 1.7 waves): 40 registers x 96 threads and 57 B of shared memory per node allow 16 CTAs
		// per SM where 128 threads get 12 (quadtree stage 0.268 -> 0.255 ms per 512 frames; at 256 frames per launch 128 threads are as
		// fast or faster: 0.203 vs 0.221 ms at the EuRoC shape).
		static const int small_env = getenv("ORBX_QT_SMALL") ? atoi(getenv("ORBX_QT_SMALL")) : 0;   // tuning knob: 96, 128 or 256
		// sort replays of more items than this run CTA-parallel (warp-parallel partition steps) in the plain variant too. Tuning knob.
		static const int plain_sort_min = getenv("ORBX_QT_PLAIN_SORT") ? atoi(getenv("ORBX_QT_PLAIN_SORT")) : 40;
		// As one lane of a two-lane batch (what callers get), 128 threads already pay from 128 frames per launch: C1 256-frame batches
		// 1.078 -> 1.034 ms, C3 1.276 -> 1.242 ms; lanes of 64 frames stay faster with 256 threads (tools/skew_probe.py).
		const int small_threads = small_env ? small_env : (P.frames >= 512 ? 96 : P.frames >= 128 ? 128 : 256);
		if (small_threads == 256)
		{
			qt256::k_quadtree<false><<<grid, 256, smem_plain, st>>>(P, cell_off, 1 << 30, plain_sort_min, want_dbg ? dbg : nullptr, s0, 0);
		}
		else if (small_threads == 96)
		{
			qt96::k_quadtree<false><<<grid, 96, smem_plain, st>>>(P, cell_off, 1 << 30, plain_sort_min, want_dbg ? dbg : nullptr, s0, 0);
		}
		else
		{
			qt128::k_quadtree<false><<<grid, 128, smem_plain, st>>>(P, cell_off, 1 << 30, plain_sort_min, want_dbg ? dbg : nullptr, s0, 0);
		}
	}
	if (want_dbg)
	{
		unsigned long long h[64];
		cudaStreamSynchronize(st);
		cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
		fprintf(stderr, "[quadtree stamps, level 0 frame 0, us]");
		for (unsigned long long k = 2; k <= h[0] && k < 64; k++) fprintf(stderr, " %.1f", (double)(h[k] - h[k - 1]) * 1e-3);
		fprintf(stderr, "\n");
	}
}

void orbx_launch_blur(const OrbxPlanDev& P, const OrbxStripMaps smaps[3], cudaStream_t st)
{
	launch_strip(P, smaps, 1, st);
}

void orbx_launch_describe(const OrbxPlanDev& P, orbx_keypoint* d_kps, uint8_t* d_desc, int32_t* d_n, cudaStream_t st)
{
	if (P.frames <= ORBX_SMALL_BATCH)
	{
		dim3 grid((P.out_cap + 1) / 2, P.frames);
		k_orient_describe2<2><<<grid, 32, OD2_SMEM, st>>>(P, d_kps, d_desc, d_n);
	}
	else
	{
		// 16 keypoints per warp (8: 0.355 -> 0.351 ms per 512 frames; the per-group constants and the angle / cos / sin step amortise
		// further) — except on 4K-class frames, whose patches come out of DRAM rather than L2: there the longer serial chain per warp
		// costs more than it saves (0.645 -> 0.789 ms per 96 frames), so they keep 8
		static const int g_env = env_int("ORBX_DESC_G", 0);    // tuning knob: 8 or 16
		const int g = g_env ? g_env : ((int64_t)P.lv[0].w * P.lv[0].h > (1 << 21) ? 8 : 16);
		if (g == 8)
		{
			dim3 grid((P.out_cap + 7) / 8, P.frames);
			k_orient_describe2<8><<<grid, 32, OD2_SMEM, st>>>(P, d_kps, d_desc, d_n);
		}
		else
		{
			dim3 grid((P.out_cap + 15) / 16, P.frames);
			k_orient_describe2<16><<<grid, 32, OD2_SMEM, st>>>(P, d_kps, d_desc, d_n);
		}
	}
}
