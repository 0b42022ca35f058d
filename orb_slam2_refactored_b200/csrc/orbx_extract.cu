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
#define FT_TS 96            // tile row stride in bytes (view <= 66 px + up to 15 px alignment slack, 16-byte chunks)
#define FT_TH 66
#define FT_SS 64            // score row stride (region <= 60 px + 1 px zero border each side)
#define FT_MAXR 60

__device__ __forceinline__ int arc_score_packed(const uint8_t* __restrict__ c)
{
	// Ring offsets inside the shared tile, OpenCV order; compile-time so every load is [base + imm].
	constexpr int R[16] = { 3 * FT_TS, 3 * FT_TS + 1, 2 * FT_TS + 2, FT_TS + 3, 3, -FT_TS + 3, -2 * FT_TS + 2, -3 * FT_TS + 1,
	                        -3 * FT_TS, -3 * FT_TS - 1, -2 * FT_TS - 2, -FT_TS - 3, -3, FT_TS - 3, 2 * FT_TS - 2, 3 * FT_TS - 1 };
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

// ---- warp-per-cell FAST with the bound pass inside the cell (ORBX_LEGACY=1; the default is the dense strip pass + k_fast_cells2
// of orbx_strip.cuh). ONE warp owns a cell end to end, so the phases need only __syncwarp, and the bound pass works on
// 4 horizontally adjacent pixels per lane: aligned 32-bit words of the tile are turned into packed u16x2 operands with PRMT
// (a pixel sits in the HIGH byte of a 16-bit lane, the low byte is a neighbour pixel and never decides a min/max), so the 8 ring
// loads + 8 packs + 7 min/max per pixel of the byte-wise version become 11 word loads + 13 PRMT + 24 VIMNMX per FOUR pixels.
#define FW_WARPS 1           // warps (= cells) per CTA; 1: a finished warp frees its shared memory at once (measured best, see orbx_launch_fast)
#define FW_TSW (FT_TS / 4)

struct OrbxFastLayout
{
	int score_stride;            // bytes per score row (region width + 2, rounded up to 8)
	int off_score, off_list, off_bm, off_bar;   // bm: the survivor bitmap, 64 bits per region row; bar: the warp's TMA mbarrier
	int warp_bytes;              // multiple of 128: every warp's tile is a TMA destination
};

// Bound pass for the 4 pixels of tile word q[0] (centre row). Returns one byte: bits (0,1,4,5) = U > minTh for pixels 0..3,
// bits (2,3,6,7) = U > iniTh, where U >= S is the 4-pair upper bound described above. kini/kmin = (0x8000 - 1 - t) in both
// halves fold the "+255" of the lane arithmetic and the threshold into one constant: bit 15 of a lane <=> U > t.
__device__ __forceinline__ uint32_t fast_bound_flags4(const uint32_t* __restrict__ q, const uint32_t kini, const uint32_t kdelta)
{
	const uint32_t c0 = q[-1], c1 = q[0], c2 = q[1];
	const uint32_t p3 = q[3 * FW_TSW], m3 = q[-3 * FW_TSW];
	const uint32_t pa = q[2 * FW_TSW - 1], pb = q[2 * FW_TSW], pc = q[2 * FW_TSW + 1];
	const uint32_t ma = q[-2 * FW_TSW - 1], mb = q[-2 * FW_TSW], mc = q[-2 * FW_TSW + 1];
	uint32_t f[2];
#pragma unroll
	for (int par = 0; par < 2; par++)
	{
		// par 0: pixels 1 and 3 (operand = bytes s..s+3 of the row, s = ring dx); par 1: pixels 0 and 2 (bytes s-1..s+2)
		const uint32_t a1 = par == 0 ? p3 : p3 << 8;                                              // ( 0, +3)
		const uint32_t a2 = par == 0 ? m3 : m3 << 8;                                              // ( 0, -3)
		const uint32_t b1 = par == 0 ? __byte_perm(pb, pc, 0x5432) : __byte_perm(pb, pc, 0x4321); // (+2, +2)
		const uint32_t b2 = par == 0 ? __byte_perm(ma, mb, 0x5432) : __byte_perm(ma, mb, 0x4321); // (-2, -2)
		const uint32_t d1 = par == 0 ? __byte_perm(c1, c2, 0x6543) : __byte_perm(c1, c2, 0x5432); // (+3,  0)
		const uint32_t d2 = par == 0 ? __byte_perm(c0, c1, 0x4321) : c0;                          // (-3,  0)
		const uint32_t e1 = par == 0 ? __byte_perm(mb, mc, 0x5432) : __byte_perm(mb, mc, 0x4321); // (+2, -2)
		const uint32_t e2 = par == 0 ? __byte_perm(pa, pb, 0x5432) : __byte_perm(pa, pb, 0x4321); // (-2, +2)
		const uint32_t hi = __vminu2(__vimin3_u16x2(__vmaxu2(a1, a2), __vmaxu2(b1, b2), __vmaxu2(d1, d2)), __vmaxu2(e1, e2));   // min_k max(pair): bright side
		const uint32_t lo = __vmaxu2(__vimax3_u16x2(__vminu2(a1, a2), __vminu2(b1, b2), __vminu2(d1, d2)), __vminu2(e1, e2));   // max_k min(pair): dark side
		// high bytes -> clean 16-bit lanes (PRMT against a zero register)
		const uint32_t H = __byte_perm(hi, 0, 0x4341), Lo = __byte_perm(lo, 0, 0x4341);
		const uint32_t C = par == 0 ? __byte_perm(c1, 0, 0x4341) : __byte_perm(c1, 0, 0x4240);
		// lanes: (H - c) + K and (c - Lo) + K with K = 0x7fff - t >= 255: no lane ever borrows or carries
		const uint32_t sb = H + kini - C, sd = C + kini - Lo;
		f[par] = __vmaxu2(sb, sd);
	}
	// bit 15 / 31 of f: U > iniTh; of f + kdelta (kdelta = iniTh - minTh per lane): U > minTh
	const uint32_t g0 = f[0] + kdelta, g1 = f[1] + kdelta;
	uint32_t z = f[0] & 0x80008000u;
	z |= (f[1] >> 1) & 0x40004000u;
	z |= (g0 >> 2) & 0x20002000u;
	z |= (g1 >> 3) & 0x10001000u;
	const uint32_t y = z >> 12;
	return (y | (y >> 12)) & 0xffu;
}

// 4 flag bytes (16 pixels) -> 16 consecutive bits; shift 2 selects the iniTh flags, 0 the minTh flags
__device__ __forceinline__ uint32_t fast_gather16(uint32_t w, int shift)
{
	uint32_t x = (w >> shift) & 0x33333333u;
	x = (x | (x >> 2)) & 0x0f0f0f0fu;
	x = (x | (x >> 4)) & 0x00ff00ffu;
	return (x | (x >> 8)) & 0xffffu;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) k_fast_cells(const OrbxPlanDev P, const __grid_constant__ OrbxTmaMaps maps, const OrbxFastLayout Y)
{
	extern __shared__ __align__(128) uint8_t fw_smem[];    // no static shared memory: 25 single-warp CTAs of 8 KB + 1 KB reserved fit an SM

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int cell = blockIdx.x * NW + warp, f = blockIdx.y;
	if (cell >= P.cells_per_frame) return;           // whole warps leave; there is no block barrier below
	uint8_t* const base = fw_smem + (size_t)warp * Y.warp_bytes;
	uint8_t* const tile = base;
	uint8_t* const score = base + Y.off_score;
	uint16_t* const list = reinterpret_cast<uint16_t*>(base + Y.off_list);
	uint8_t* const nib = base + Y.off_list;          // flag bytes of the bound pass, dead before the list is built
	uint32_t* const bm_sel = reinterpret_cast<uint32_t*>(base + Y.off_bm);   // [row][2]: 64 bits per region row
	uint64_t* const tma_bar = reinterpret_cast<uint64_t*>(base + Y.off_bar);
	const int SS = Y.score_stride;

	const int4 ct = __ldg(P.cell_tab + cell);
	const int x0 = ct.x & 0xffff, y0 = ct.x >> 16, vw = ct.y & 0xffff, vh = ct.y >> 16, lvl = ct.z, c = ct.w;
	const OrbxLevel& L = P.lv[lvl];
	const int rw = vw - 6, rh = vh - 6;
	const int sh = x0 & 15;                           // the TMA box starts 16-byte aligned
	if (lane == 0)
	{
		mbar_init(tma_bar, 1);
		mbar_expect_tx(tma_bar, (unsigned)(FT_TS * maps.box_h[lvl]));
		tma_load_3d(tile, &maps.level[lvl], x0 - sh, y0, P.frame0 + f, tma_bar);
	}
	for (int i = lane; i < (rh + 2) * (SS / 8); i += 32)
		reinterpret_cast<uint2*>(score)[i] = make_uint2(0, 0);
	for (int i = lane; i < 2 * rh; i += 32) bm_sel[i] = 0;
	__syncwarp();
	mbar_wait(tma_bar, 0);

	const int tmin = P.min_th, tini = P.ini_th;
	const uint8_t* __restrict__ t0 = tile + 3 * FT_TS + sh + 3;

	// ---- A: bound pass, 4 pixels per lane. Groups are the aligned words of a tile row that overlap the region: the first one
	//      starts `a` pixels left of it. Pixels outside the region produce flags that the row assembly below shifts/masks away.
	const int a = (sh + 3) & 3, G = (a + rw + 3) >> 2, ngroups = G * rh;
	{
		const uint32_t kini = (uint32_t)(0x7fff - tini) * 0x00010001u, kdelta = (uint32_t)(tini - tmin) * 0x00010001u;
		const uint32_t* __restrict__ q0 = reinterpret_cast<const uint32_t*>(tile) + 3 * FW_TSW + ((sh + 3) >> 2);
		const uint32_t invG = c_inv20[G];
		for (int g = lane; g < ngroups; g += 32)
		{
			const int ry = (int)(((uint32_t)g * invG) >> 20), cg = g - ry * G;
			nib[ry * 16 + cg] = (uint8_t)fast_bound_flags4(q0 + ry * FW_TSW + cg, kini, kdelta);
		}
	}
	__syncwarp();
	// rows -> bitmaps, kept in registers: lane r assembles region rows r and r + 32 (wa = U > iniTh, wb = minTh < U <= iniTh)
	uint32_t wa[4], wb[4];                           // this lane's rows: [row slot][lo, hi]
	{
		const uint64_t rowmask = rw >= 64 ? ~0ull : ((1ull << rw) - 1ull);
#pragma unroll
		for (int k = 0; k < 2; k++)
		{
			const int r = lane + 32 * k;
			uint64_t va = 0, vb = 0;
			if (r < rh)
			{
				const uint4 n = *reinterpret_cast<const uint4*>(nib + r * 16);
				va = (uint64_t)(fast_gather16(n.x, 2) | (fast_gather16(n.y, 2) << 16)) | ((uint64_t)(fast_gather16(n.z, 2) | (fast_gather16(n.w, 2) << 16)) << 32);
				vb = (uint64_t)(fast_gather16(n.x, 0) | (fast_gather16(n.y, 0) << 16)) | ((uint64_t)(fast_gather16(n.z, 0) | (fast_gather16(n.w, 0) << 16)) << 32);
				va = (va >> a) & rowmask; vb = (vb >> a) & rowmask;
			}
			wa[2 * k] = (uint32_t)va; wa[2 * k + 1] = (uint32_t)(va >> 32);
			wb[2 * k] = (uint32_t)vb & ~wa[2 * k]; wb[2 * k + 1] = (uint32_t)(vb >> 32) & ~wa[2 * k + 1];   // minTh < U <= iniTh
		}
	}
	__syncwarp();                                    // nib is dead: the list may overwrite it

	// exclusive warp scan of (c0, c1) in "all first rows, then all second rows" order = row-major; returns offsets, total in `total`
	auto scan2 = [&](int c0, int c1, int& o0, int& o1, int& total) {
		int i0 = c0, i1 = c1;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const int u0 = __shfl_up_sync(0xffffffffu, i0, d), u1 = __shfl_up_sync(0xffffffffu, i1, d);
			if (lane >= d) { i0 += u0; i1 += u1; }
		}
		const int t0s = __shfl_sync(0xffffffffu, i0, 31), t1s = __shfl_sync(0xffffffffu, i1, 31);
		o0 = i0 - c0; o1 = t0s + i1 - c1; total = t0s + t1s;
	};
	// append the pixels of this lane's row words to the list as ry << 6 | rx; returns how many the warp appended
	auto expand = [&](const uint32_t* w, int at) {
		int o0, o1, total;
		scan2(__popc(w[0]) + __popc(w[1]), __popc(w[2]) + __popc(w[3]), o0, o1, total);
#pragma unroll
		for (int k = 0; k < 4; k++)
		{
			uint32_t x = w[k];
			int pos = at + (k < 2 ? o0 : o1);
			if (k == 1) pos += __popc(w[0]);
			if (k == 3) pos += __popc(w[2]);
			const int tag = ((lane + 32 * (k >> 1)) << 6) | (32 * (k & 1));
			while (x)
			{
				list[pos++] = (uint16_t)(tag + __ffs(x) - 1);
				x &= x - 1;
			}
		}
		return total;
	};
	auto evaluate = [&](int from, int to) {
		for (int j = from + lane; j < to; j += 32)
		{
			const int e = list[j], ry = e >> 6, rx = e & 63;
			const int s = arc_score_packed(t0 + ry * FT_TS + rx);
			score[(ry + 1) * SS + rx + 1] = (uint8_t)max(s, 0);
		}
	};
	auto select = [&](int to, int t) {
		bool found = false;
		for (int j = lane; j < to; j += 32)
		{
			const int e = list[j], ry = e >> 6, rx = e & 63;
			const uint8_t* sp = score + (ry + 1) * SS + rx + 1;
			const int s = sp[0];
			if (s > t)
			{
				const int m = max(max(max((int)sp[-SS - 1], (int)sp[-SS]), max((int)sp[-SS + 1], (int)sp[-1])),
				                  max(max((int)sp[1], (int)sp[SS - 1]), max((int)sp[SS], (int)sp[SS + 1])));
				if (s > m) { atomicOr(&bm_sel[2 * ry + (rx >> 5)], 1u << (rx & 31)); found = true; }
			}
		}
		return found;
	};

	// ---- B + C at iniTh; retry at minTh if the cell has no corner (:526-530)
	const int n1 = expand(wa, 0);
	__syncwarp();
	evaluate(0, n1);
	__syncwarp();
	if (!__any_sync(0xffffffffu, select(n1, tini)))
	{
		const int n2 = expand(wb, n1);
		__syncwarp();
		evaluate(n1, n1 + n2);
		__syncwarp();
		select(n1 + n2, tmin);
	}
	__syncwarp();

	// ---- D: ordered emit (rows ascending, x ascending = cv::FAST's order inside the view)
	uint32_t ws[4];
#pragma unroll
	for (int k = 0; k < 2; k++)
	{
		const int r = lane + 32 * k;
		ws[2 * k] = r < rh ? bm_sel[2 * r] : 0u;
		ws[2 * k + 1] = r < rh ? bm_sel[2 * r + 1] : 0u;
	}
	int o0, o1, total;
	scan2(__popc(ws[0]) + __popc(ws[1]), __popc(ws[2]) + __popc(ws[3]), o0, o1, total);
	uint32_t* __restrict__ out = P.cand + (int64_t)f * P.cand_per_frame + L.cand_base + (int64_t)c * L.cell_cap;
#pragma unroll
	for (int k = 0; k < 4; k++)
	{
		uint32_t x = ws[k];
		int pos = (k < 2 ? o0 : o1);
		if (k == 1) pos += __popc(ws[0]);
		if (k == 3) pos += __popc(ws[2]);
		const int ry = lane + 32 * (k >> 1);
		while (x)
		{
			const int rx = 32 * (k & 1) + __ffs(x) - 1;
			x &= x - 1;
			const int s = score[(ry + 1) * SS + rx + 1];
			out[pos++] = orbx_pack(x0 + 3 + rx, y0 + 3 + ry, s - 1);
		}
	}
	if (lane == 0)
		P.cell_count[(int64_t)f * P.cells_per_frame + cell] = total;
}

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
// =====================================================================================================
// K6  gauss7x7_u8 — cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) in OpenCV's 8.8 fixed point
//     (src/ORBextractor.cc:799). Tile 128 x 32 with a 3 px halo in shared memory.
// =====================================================================================================
#define GB_TW 128                 // output tile
#define GB_TH 32
#define GB_RROWS (GB_TH + 8)      // raw rows y0-4 .. y0+35 (the row-pair grid needs an even start)
#define GB_RWORDS 40              // raw cols x0-16 .. x0+143 as ten 16-byte chunks
#define GB_LEFT 16                // raw byte of image column x0

__device__ __forceinline__ int reflect101(int i, int n)
{
	if (i < 0) i = -i;
	if (i >= n) i = 2 * n - 2 - i;
	return i;
}

// Horizontal pass with IDP.4A on packed bytes (two 4-tap dot products per output), results stored as 16-bit row pairs
// (h[2p][x] | h[2p+1][x] << 16); vertical pass with IDP.2A on those pairs (four 2-tap dot products per output).
// h <= 255*256 fits 16 bits; v = sum K*h < 2^24; out = (v + 32768) >> 16.
// One launch blurs every level: blockIdx.x runs over the tiles of all levels (the levels do not depend on each other).
struct BlurTiles { int base[ORBX_MAX_LEVELS + 1]; int tx[ORBX_MAX_LEVELS]; };
// ALL = false: one level per launch (grid = tiles_x, tiles_y, frames), the level is a kernel parameter.
template <bool ALL>
__global__ void __launch_bounds__(256) k_gauss7(const OrbxPlanDev P, const BlurTiles T, const int one_level)
{
	__shared__ __align__(16) uint32_t raw[GB_RROWS * GB_RWORDS];
	__shared__ __align__(16) uint32_t hv[(GB_RROWS / 2) * GB_TW];
	int level = one_level, tile_x = blockIdx.x, tile_y = blockIdx.y;
	if (ALL)
	{
		// unrolled so that every index into T is a compile-time constant (a dynamic index would copy the struct to local memory)
		int base = 0, tx = T.tx[0];
		level = 0;
#pragma unroll
		for (int s = 1; s < ORBX_MAX_LEVELS; s++)
			if ((int)blockIdx.x >= T.base[s] && T.base[s + 1] > T.base[s]) { level = s; base = T.base[s]; tx = T.tx[s]; }
		const int tile = (int)blockIdx.x - base;
		tile_y = tile / tx; tile_x = tile - tile_y * tx;
	}
	const OrbxLevel& L = P.lv[level];
	const int f = blockIdx.z, tid = threadIdx.x;
	const int x0 = tile_x * GB_TW, y0 = tile_y * GB_TH;
	const uint8_t* __restrict__ src = orbx_level_ptr(P, f, level);
	const int64_t sp = orbx_level_pitch(P, level);
	uint8_t* __restrict__ dst = P.blur + (int64_t)f * P.slab + L.offset;
	const int w = L.w, h = L.h;

	// ---- stage the raw tile with 16-byte async copies (rows beyond the image come from the reflected row; the padded level
	//      buffers make columns -16..-1 and w.. legal to read), then patch the 3 + 3 border columns in place with
	//      REFLECT_101 (-1 -> 1, w -> w-2): the mirrored pixels are inside the same staged row.
	uint8_t* rawb = reinterpret_cast<uint8_t*>(raw);
	for (int i = tid; i < GB_RROWS * (GB_RWORDS / 4); i += 256)
	{
		const int r = i / (GB_RWORDS / 4), c = i - r * (GB_RWORDS / 4);
		const int gy = reflect101(min(y0 - 4 + r, h + 2), h);
		cp_async16(rawb + r * (GB_RWORDS * 4) + c * 16, src + (int64_t)gy * sp + (x0 - GB_LEFT) + c * 16);
	}
	cp_async_wait_all();
	__syncthreads();
	if (tid < GB_RROWS)
	{
		uint8_t* row = rawb + tid * (GB_RWORDS * 4) + GB_LEFT - x0;     // row[x] = image column x
		if (x0 == 0) { row[-1] = row[1]; row[-2] = row[2]; row[-3] = row[3]; }
		if (w < x0 + GB_TW + 3)                                         // columns w .. w+2 are inside this tile's window
		{
#pragma unroll
			for (int k = 0; k < 3; k++)
				if (w + k < x0 + GB_TW + 3 && w - 2 - k >= x0 - GB_LEFT) row[w + k] = row[w - 2 - k];
		}
	}
	__syncthreads();

	// ---- horizontal pass: item = (row pair, column quad); output x = x0 + 4q + k taps image columns x-3 .. x+3
	const uint32_t KA = 18u | (34u << 8) | (48u << 16) | (56u << 24), KB = 48u | (34u << 8) | (18u << 16);
	for (int i = tid; i < (GB_RROWS / 2) * (GB_TW / 4); i += 256)
	{
		const int pr = i >> 5, q = i & 31;
		const uint32_t* r0 = raw + (2 * pr) * GB_RWORDS + q + 3;     // words holding image columns x0+4q-4 .. x0+4q+7
		const uint32_t* r1 = r0 + GB_RWORDS;
		const uint32_t a0 = r0[0], a1 = r0[1], a2 = r0[2], b0 = r1[0], b1 = r1[1], b2 = r1[2];
		uint4 o;
		uint32_t he, ho;
		he = __dp4a(__funnelshift_r(a0, a1, 8), KA, __dp4a(__funnelshift_r(a1, a2, 8), KB, 0u));
		ho = __dp4a(__funnelshift_r(b0, b1, 8), KA, __dp4a(__funnelshift_r(b1, b2, 8), KB, 0u));
		o.x = he | (ho << 16);
		he = __dp4a(__funnelshift_r(a0, a1, 16), KA, __dp4a(__funnelshift_r(a1, a2, 16), KB, 0u));
		ho = __dp4a(__funnelshift_r(b0, b1, 16), KA, __dp4a(__funnelshift_r(b1, b2, 16), KB, 0u));
		o.y = he | (ho << 16);
		he = __dp4a(__funnelshift_r(a0, a1, 24), KA, __dp4a(__funnelshift_r(a1, a2, 24), KB, 0u));
		ho = __dp4a(__funnelshift_r(b0, b1, 24), KA, __dp4a(__funnelshift_r(b1, b2, 24), KB, 0u));
		o.z = he | (ho << 16);
		he = __dp4a(a1, KA, __dp4a(a2, KB, 0u));
		ho = __dp4a(b1, KA, __dp4a(b2, KB, 0u));
		o.w = he | (ho << 16);
		reinterpret_cast<uint4*>(hv)[i] = o;      // hv[pr][4q .. 4q+3]
	}
	__syncthreads();

	// ---- vertical pass: output row y0 + r taps grid rows r+1 .. r+7 (grid row g = image row y0-4+g, pair = g >> 1)
	// r even: (0,K0)(K1,K2)(K3,K4)(K5,K6) on pairs r/2 .. r/2+3;  r odd: (K0,K1)(K2,K3)(K4,K5)(K6,0) on pairs (r+1)/2 ..
	for (int i = tid; i < GB_TH * (GB_TW / 4); i += 256)
	{
		const int r = i >> 5, q = i & 31;
		if (y0 + r >= h || x0 + 4 * q >= w) continue;
		const bool odd = r & 1;
		const uint32_t c01 = odd ? (18u | (34u << 8) | (48u << 16) | (56u << 24)) : ((18u << 8) | (34u << 16) | (48u << 24));
		const uint32_t c23 = odd ? (48u | (34u << 8) | (18u << 16)) : (56u | (48u << 8) | (34u << 16) | (18u << 24));
		const uint4* pp = reinterpret_cast<const uint4*>(hv) + ((r + 1) >> 1) * (GB_TW / 4) + q;
		const uint4 p0 = pp[0], p1 = pp[GB_TW / 4], p2 = pp[2 * (GB_TW / 4)], p3 = pp[3 * (GB_TW / 4)];
		uint32_t v0 = __dp2a_lo(p0.x, c01, 32768u), v1 = __dp2a_lo(p0.y, c01, 32768u), v2 = __dp2a_lo(p0.z, c01, 32768u), v3 = __dp2a_lo(p0.w, c01, 32768u);
		v0 = __dp2a_hi(p1.x, c01, v0); v1 = __dp2a_hi(p1.y, c01, v1); v2 = __dp2a_hi(p1.z, c01, v2); v3 = __dp2a_hi(p1.w, c01, v3);
		v0 = __dp2a_lo(p2.x, c23, v0); v1 = __dp2a_lo(p2.y, c23, v1); v2 = __dp2a_lo(p2.z, c23, v2); v3 = __dp2a_lo(p2.w, c23, v3);
		v0 = __dp2a_hi(p3.x, c23, v0); v1 = __dp2a_hi(p3.y, c23, v1); v2 = __dp2a_hi(p3.z, c23, v2); v3 = __dp2a_hi(p3.w, c23, v3);
		const uint32_t out = (v0 >> 16) | ((v1 >> 16) << 8) | ((v2 >> 16) << 16) | ((v3 >> 16) << 24);
		*reinterpret_cast<uint32_t*>(dst + (int64_t)(y0 + r) * L.pitch + x0 + 4 * q) = out;
	}
}

// =====================================================================================================
// K5+K7  orient_describe — IC_Angle (src/ORBextractor.cc:74-101) on the un-blurred level, then
//     ComputeOrbDescriptor (:103-140) on the blurred level, then the keypoint record of Extract (:768-773,
//     :811-815). One warp per output keypoint. Float path pinned per SURVEY H2 / App. A.6-A.7.
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

#define OD_WARPS 8
#define OD_PS 80                                     // patch row stride in shared memory (16-byte chunks; 20 words spreads rows over banks)
#define OD_IMG_ROWS 31                               // un-blurred patch rows y-15 .. y+15, 48 bytes each from (x-16) & ~15
#define OD_BLR_ROWS 37                               // blurred patch rows y-18 .. y+18, 64 bytes each from (x-18) & ~15
#define OD_WARP_BYTES ((OD_IMG_ROWS + OD_BLR_ROWS) * OD_PS)
#define OD_SMEM (OD_WARPS * OD_WARP_BYTES)

__global__ void __launch_bounds__(OD_WARPS * 32) k_orient_describe(const OrbxPlanDev P, orbx_keypoint* __restrict__ d_kps,
                                                                  uint8_t* __restrict__ d_desc, int32_t* __restrict__ d_n)
{
	// One warp per keypoint. Both patches the keypoint touches (31x31 of the level for IC_Angle, 37x37 of the blurred level for
	// the 512 rBRIEF samples) are staged in shared memory with 16-byte async copies — 241 coalesced chunks instead of ~570
	// scattered byte gathers — and everything after that reads shared memory.
	extern __shared__ __align__(16) uint8_t od_smem[];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int f = blockIdx.y;

	// which level does output slot `slot` belong to (levels are concatenated in order, :792-819): lane l holds level l's count
	const int slot = blockIdx.x * OD_WARPS + warp;
	const int cnt = (lane < P.nlevels) ? P.sel_count[(int64_t)f * P.nlevels + lane] : 0;
	int incl = cnt;
#pragma unroll
	for (int d = 1; d < 16; d <<= 1)
	{
		const int t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += t;
	}
	const int total = __shfl_sync(0xffffffffu, incl, ORBX_MAX_LEVELS - 1);
	const unsigned inside = __ballot_sync(0xffffffffu, lane < P.nlevels && slot < incl);
	if (slot == 0 && lane == 0) d_n[f] = total;
	const bool live = inside != 0 && slot < P.out_cap;
	int lvl = 0, x = 0, y = 0, resp = 0;
	uint8_t* pimg = od_smem + warp * OD_WARP_BYTES;
	uint8_t* pblr = pimg + OD_IMG_ROWS * OD_PS;
	int xi = 0, xb = 0;
	if (live)
	{
		lvl = __ffs(inside) - 1;
		const int start = __shfl_sync(0xffffffffu, incl - cnt, lvl);
		const OrbxLevel& L = P.lv[lvl];
		const uint32_t kp = P.sel[(int64_t)f * P.sel_per_frame + L.sel_base + (slot - start)];
		x = orbx_px(kp); y = orbx_py(kp); resp = orbx_pr(kp);
		const int64_t ip = orbx_level_pitch(P, lvl);
		xi = (x - 16) & ~15; xb = (x - 18) & ~15;
		const uint8_t* __restrict__ gi = orbx_level_ptr(P, f, lvl) + (int64_t)(y - 15) * ip + xi;
		const uint8_t* __restrict__ gb = P.blur + (int64_t)f * P.slab + L.offset + (int64_t)(y - 18) * L.pitch + xb;
		for (int i = lane; i < OD_IMG_ROWS * 3; i += 32)
		{
			const int r = i / 3, c = i - r * 3;
			cp_async16(pimg + r * OD_PS + c * 16, gi + (int64_t)r * ip + c * 16);
		}
		for (int i = lane; i < OD_BLR_ROWS * 4; i += 32)
		{
			const int r = i >> 2, c = i & 3;
			cp_async16(pblr + r * OD_PS + c * 16, gb + (int64_t)r * L.pitch + c * 16);
		}
	}
	if (!live)
		return;
	cp_async_wait_all();
	__syncwarp();               // this warp's two patches have landed; warps never wait for each other (no block barrier in this kernel)
	const OrbxLevel& L = P.lv[lvl];

	// ---- intensity centroid over the radius-15 disc (IC_Angle, :74-101): lane = disc row v in [-15, 15]. The row's 32 bytes
	//      [x-16, x+15] are 9 shared-memory words, re-aligned with funnel shifts and reduced with IDP.4A against the per-row
	//      coefficient words: row sum (m01 = sum v * I) and sum of u * I (m10). Integer arithmetic, so any order is exact.
	int m10 = 0, m01 = 0;
	if (lane < 31)
	{
		const int v = lane - ORBX_HALF_PATCH;
		const int o = x - 16 - xi;                                  // byte offset of column x-16 inside the staged row, 0..15
		const uint32_t* wp = reinterpret_cast<const uint32_t*>(pimg + lane * OD_PS) + (o >> 2);
		const int shb = (o & 3) * 8;
		uint32_t w[9];
#pragma unroll
		for (int k = 0; k < 9; k++) w[k] = wp[k];
		const uint2* __restrict__ tab = g_mom + abs(v);
		int rowsum = 0;
#pragma unroll
		for (int k = 0; k < 8; k++)
		{
			const uint32_t win = __funnelshift_r(w[k], w[k + 1], shb);
			const uint2 cf = __ldg(tab + k * 16);
			rowsum = (int)__dp4a(win, cf.x, (uint32_t)rowsum);
			m10 = dp4a_u8_s8(win, cf.y, m10);
		}
		m01 = v * rowsum;
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
	{
		m10 += __shfl_xor_sync(0xffffffffu, m10, d);
		m01 += __shfl_xor_sync(0xffffffffu, m01, d);
	}
	const float angle = fast_atan2_deg((float)m01, (float)m10);

	// ---- steered BRIEF (ComputeOrbDescriptor, :103-140): lane = descriptor byte, 8 pairs each, samples from the staged patch
	float ca, sb;
	orb_cos_sin(angle, ca, sb);
	const uint8_t* bl = pblr + 18 * OD_PS + (x - xb);
	uint32_t byte = 0;
#pragma unroll
	for (int bit = 0; bit < 8; bit++)
	{
		const float4 pt = __ldg(g_patf + bit * 32 + lane);
		const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(pt.x, sb), __fmul_rn(pt.y, ca)));
		const int q0 = __float2int_rn(__fsub_rn(__fmul_rn(pt.x, ca), __fmul_rn(pt.y, sb)));
		const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(pt.z, sb), __fmul_rn(pt.w, ca)));
		const int q1 = __float2int_rn(__fsub_rn(__fmul_rn(pt.z, ca), __fmul_rn(pt.w, sb)));
		const int t0 = bl[r0 * OD_PS + q0], t1 = bl[r1 * OD_PS + q1];
		byte |= (uint32_t)(t0 < t1) << bit;
	}
	d_desc[((int64_t)f * P.out_cap + slot) * 32 + lane] = (uint8_t)byte;

	if (lane == 0)
	{
		orbx_keypoint o;
		o.x = (float)x; o.y = (float)y;
		if (lvl > 0) { o.x = __fmul_rn(o.x, L.scale); o.y = __fmul_rn(o.y, L.scale); }   // :811-815
		o.size = __fmul_rn(L.scale, (float)ORBX_PATCH);                                    // :771
		o.angle = angle;
		o.response = (float)resp;
		o.octave = lvl;
		o.class_id = -1;
		d_kps[(int64_t)f * P.out_cap + slot] = o;
	}
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
static bool legacy_kernels() { static const bool v = env_int("ORBX_LEGACY", 0) != 0; return v; }     // A/B: the round-1 kernels
bool orbx_fused_blur_fast() { static const bool v = env_int("ORBX_FUSE", 0) != 0 && !legacy_kernels(); return v; }
int orbx_strip_rows(int which)
{
	static const int v = env_int("ORBX_STRIP_TH", 32) == 64 ? 64 : env_int("ORBX_STRIP_TH", 32) == 16 ? 16 : 32;
	return which ? 8 : v;
}
int orbx_strip_box_w() { return ST_BW; }
int orbx_pyramid_strip_rows(int which)
{
	static const int v = env_int("ORBX_PYR_TH", 32) == 16 ? 16 : 32;
	return which ? 8 : v;
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
	set(k_fast_cells<1>, 64 * 1024);
	set(k_fast_cells2, 64 * 1024);
	set(k_level_strip<8, true, false>, 64 * 1024); set(k_level_strip<8, false, true>, 64 * 1024); set(k_level_strip<8, true, true>, 64 * 1024);
	set(k_level_strip<16, true, false>, 64 * 1024); set(k_level_strip<16, false, true>, 64 * 1024); set(k_level_strip<16, true, true>, 64 * 1024);
	set(k_level_strip<32, true, false>, 64 * 1024); set(k_level_strip<32, false, true>, 64 * 1024); set(k_level_strip<32, true, true>, 64 * 1024);
	set(k_level_strip<64, true, false>, 64 * 1024); set(k_level_strip<64, false, true>, 64 * 1024); set(k_level_strip<64, true, true>, 64 * 1024);
	set(qt128::k_quadtree<false>, QT_SMEM_MAX); set(qt256::k_quadtree<false>, QT_SMEM_MAX);
	set(qt256::k_quadtree<true>, QT_SMEM_MAX); set(qt512::k_quadtree<true>, QT_SMEM_MAX);
	set(k_orient_describe, OD_SMEM);
	set(k_orient_describe2<8>, OD2_SMEM); set(k_orient_describe2<2>, OD2_SMEM);
	return e;
}

void orbx_launch_pyramid(const OrbxPlanDev& P, const OrbxPyrMaps pmaps[2], int level, cudaStream_t st)
{
	const OrbxLevel& D = P.lv[level];
	const bool small_batch = P.frames <= ORBX_SMALL_BATCH;
	const int which = small_batch ? 1 : 0;
	if (D.py_bw[which] > 0 && !legacy_kernels())
	{
		// strip kernel: one warp per 128 x TH output tile, source rectangle by one TMA box
		const int th = orbx_pyramid_strip_rows(which), bw = D.py_bw[which], bh = D.py_bh[which];
		dim3 grid((D.w + ST_TW - 1) / ST_TW, (D.h + th - 1) / th, P.frames);
		const int smem = ((bw * bh + 127) & ~127) + 16;
		if (th == 8) k_pyramid_strip<8><<<grid, 32, smem, st>>>(P, pmaps[which], level, bw, bh);
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

int orbx_fast_tile_stride() { return FT_TS; }
int orbx_fast_tile_rows() { return FT_TH; }

static OrbxStripTiles strip_tiles(const OrbxPlanDev& P, int th)
{
	OrbxStripTiles T = {};
	T.one = 1;
	T.base[0] = 0;
	for (int s = 0; s < P.nlevels; s++)
	{
		const OrbxLevel& L = P.lv[s];
		T.tx[s] = (L.w + ST_TW - 1) / ST_TW;
		T.base[s + 1] = T.base[s] + T.tx[s] * ((L.h + th - 1) / th);
	}
	for (int s = P.nlevels; s < ORBX_MAX_LEVELS; s++) { T.tx[s] = 1; T.base[s + 1] = T.base[s]; }
	return T;
}

// one launch over the tiles of all levels: blur (mode 1), dense FAST bound (mode 2) or both (mode 3)
static void launch_strip(const OrbxPlanDev& P, const OrbxStripMaps smaps_both[2], int mode, cudaStream_t st)
{
	const int which = P.frames <= ORBX_SMALL_BATCH ? 1 : 0;
	const OrbxStripMaps& smaps = smaps_both[which];
	const int th = orbx_strip_rows(which);
	const OrbxStripTiles T = strip_tiles(P, th);
	dim3 grid(T.base[P.nlevels], P.frames);
	const int smem = st_tile_bytes(th) + 16;
#define ORBX_STRIP_CASE(TH_)                                                                                      \
	if (mode == 1) k_level_strip<TH_, true, false><<<grid, 32, smem, st>>>(P, smaps, T);                           \
	else if (mode == 2) k_level_strip<TH_, false, true><<<grid, 32, smem, st>>>(P, smaps, T);                      \
	else k_level_strip<TH_, true, true><<<grid, 32, smem, st>>>(P, smaps, T);
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

static void launch_cells2(const OrbxPlanDev& P, const OrbxTmaMaps& maps, cudaStream_t st)
{
	int rows, maxrw, maxrh;
	cell_extents(P, maps, rows, maxrw, maxrh);
	// tile | score (1 px zero border) | list of pixels to score | survivor bitmap | mbarrier
	OrbxCellLayout Y;
	Y.score_stride = (maxrw + 2 + 7) & ~7;
	Y.off_score = (rows * FT_TS + 15) & ~15;
	Y.off_list = (Y.off_score + (maxrh + 2) * Y.score_stride + 15) & ~15;
	Y.off_bm = (Y.off_list + maxrw * maxrh * 2 + 15) & ~15;
	Y.off_bar = Y.off_bm + 8 * maxrh;
	Y.warp_bytes = (Y.off_bar + 8 + 127) & ~127;
	dim3 grid(P.cells_per_frame, P.frames);
	k_fast_cells2<<<grid, 32, Y.warp_bytes, st>>>(P, maps, Y);
}

void orbx_launch_fast(const OrbxPlanDev& P, const OrbxTmaMaps& maps, const OrbxStripMaps smaps[2], cudaStream_t st)
{
	if (!legacy_kernels())
	{
		launch_strip(P, smaps, 2, st);
		launch_cells2(P, maps, st);
		return;
	}
	// per-warp shared memory, sized by the largest cell of the plan: tile | score (1 px zero border) | list (aliases the flag bytes) | survivor bitmap | mbarrier
	int rows, maxrw, maxrh;
	cell_extents(P, maps, rows, maxrw, maxrh);
	OrbxFastLayout Y;
	Y.score_stride = (maxrw + 2 + 7) & ~7;
	Y.off_score = (rows * FT_TS + 15) & ~15;
	Y.off_list = (Y.off_score + (maxrh + 2) * Y.score_stride + 15) & ~15;          // 16-byte aligned: the flag bytes are read as uint4
	Y.off_bm = (Y.off_list + std::max(maxrw * maxrh * 2, maxrh * 16) + 15) & ~15;
	Y.off_bar = Y.off_bm + 8 * maxrh;
	Y.warp_bytes = (Y.off_bar + 8 + 127) & ~127;
	// One warp per CTA: a CTA's shared memory is released when its LAST warp retires, and cells differ in cost (retry, corner count); with
	// 2 or 4 warps per CTA resident warp slots sat idle behind a straggler (measured 0.492 / 0.503 vs 0.471 ms per 256 frames).
	dim3 grid(P.cells_per_frame, P.frames);
	k_fast_cells<1><<<grid, 32, Y.warp_bytes, st>>>(P, maps, Y);
}

void orbx_launch_blur_fast(const OrbxPlanDev& P, const OrbxTmaMaps& maps, const OrbxStripMaps smaps[2], cudaStream_t st)
{
	launch_strip(P, smaps, 3, st);
	launch_cells2(P, maps, st);
}

int orbx_pyramid_tile_rows() { return PY_TH; }
int orbx_pyramid_max_src_rows() { return PY_SRC; }
int orbx_pyramid_tile_cols() { return PY_TW; }
int orbx_pyramid_max_src_bytes() { return PY_SW; }

size_t orbx_quadtree_smem(int node_cap)
{
	// listA, listB (16 B), items (8 B), childcnt (16 B), proc, pbase (4 B each), leaf (8 B), gone (1 B), 2 segment queues (12 B x cap/16)
	return (size_t)node_cap * (16 + 16 + 8 + 16 + 4 + 4 + 8 + 1) + 2 * 12 * ((size_t)node_cap / 16 + 4) + 64;
}

void orbx_launch_quadtree(const OrbxPlanDev& P, int* cell_off, cudaStream_t st)
{
	const size_t smem = orbx_quadtree_smem(P.node_cap);
	dim3 grid(P.frames, P.nlevels);
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
		if (large_plan && !small_batch && big_threads == 512)
		{
			qt512::k_quadtree<true><<<grid, 512, smem, st>>>(P, cell_off, small_batch ? 1024 : 2048, small_batch ? 48 : 512, want_dbg ? dbg : nullptr);
		}
		else
		{
			qt256::k_quadtree<true><<<grid, 256, smem, st>>>(P, cell_off, small_batch ? 1024 : 2048, small_batch ? 48 : 512, want_dbg ? dbg : nullptr);
		}
	}
	else
	{
		// launches of >= 256 frames (>= 2048 CTAs): 128-thread CTAs, so that more of them share an SM while others are in their serial
		// phases; below that the SMs are not full and the CTA's own speed counts. Measured frames/s with 128 / 256 threads, one launch per
		// stage: 32 frames 83.2 k / 89.0 k, 128 frames 136.8 k / 142.3 k, 256 frames 165.8 k / 162.9 k, 512 frames 177.4 k / 174.6 k.
		static const int small_env = getenv("ORBX_QT_SMALL") ? atoi(getenv("ORBX_QT_SMALL")) : 0;   // tuning knob: 128 or 256
		const int small_threads = small_env ? small_env : (P.frames >= 256 ? 128 : 256);
		if (small_threads == 256)
		{
			qt256::k_quadtree<false><<<grid, 256, smem, st>>>(P, cell_off, 1 << 30, 1 << 30, want_dbg ? dbg : nullptr);
		}
		else
		{
			qt128::k_quadtree<false><<<grid, 128, smem, st>>>(P, cell_off, 1 << 30, 1 << 30, want_dbg ? dbg : nullptr);
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

void orbx_launch_blur(const OrbxPlanDev& P, const OrbxStripMaps smaps[2], cudaStream_t st)
{
	if (!legacy_kernels())
	{
		launch_strip(P, smaps, 1, st);
		return;
	}
	// Small batches (a frame at a time): one launch over the tiles of all levels, 8 launches fewer on the critical path (single-frame
	// blur 36 -> 9 us). Large batches: one launch per level, which measures ~6 % faster there.
	const bool one_launch = P.frames <= 16;
	BlurTiles T = {};
	if (one_launch)
	{
		T.base[0] = 0;
		for (int s = 0; s < P.nlevels; s++)
		{
			const OrbxLevel& L = P.lv[s];
			T.tx[s] = (L.w + GB_TW - 1) / GB_TW;
			T.base[s + 1] = T.base[s] + T.tx[s] * ((L.h + GB_TH - 1) / GB_TH);
		}
		for (int s = P.nlevels; s < ORBX_MAX_LEVELS; s++) { T.tx[s] = 1; T.base[s + 1] = T.base[s]; }
		dim3 grid(T.base[P.nlevels], 1, P.frames);
		k_gauss7<true><<<grid, 256, 0, st>>>(P, T, 0);
		return;
	}
	for (int s = 0; s < P.nlevels; s++)
	{
		const OrbxLevel& L = P.lv[s];
		dim3 grid((L.w + GB_TW - 1) / GB_TW, (L.h + GB_TH - 1) / GB_TH, P.frames);
		k_gauss7<false><<<grid, 256, 0, st>>>(P, T, s);
	}
}

void orbx_launch_describe(const OrbxPlanDev& P, orbx_keypoint* d_kps, uint8_t* d_desc, int32_t* d_n, cudaStream_t st)
{
	if (!legacy_kernels())
	{
		if (P.frames <= ORBX_SMALL_BATCH)
		{
			dim3 grid((P.out_cap + 1) / 2, P.frames);
			k_orient_describe2<2><<<grid, 32, OD2_SMEM, st>>>(P, d_kps, d_desc, d_n);
		}
		else
		{
			dim3 grid((P.out_cap + 7) / 8, P.frames);
			k_orient_describe2<8><<<grid, 32, OD2_SMEM, st>>>(P, d_kps, d_desc, d_n);
		}
		return;
	}
	dim3 grid((P.out_cap + OD_WARPS - 1) / OD_WARPS, P.frames);
	k_orient_describe<<<grid, OD_WARPS * 32, OD_SMEM, st>>>(P, d_kps, d_desc, d_n);
}
