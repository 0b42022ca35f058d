// Strip kernels of the extractor (included by orbx_extract.cu inside its anonymous namespace).
//
// One warp owns one tile of a pyramid level end to end: ONE TMA tile load (cp.async.bulk.tensor + a per-warp mbarrier) puts the
// tile and its halo in shared memory, then every lane walks DOWN its four pixel columns with the rows it needs held in registers
// (a sliding window), so a staged byte is read from shared memory once, there is no block barrier, and the per-row cost is the
// arithmetic itself. A single-warp CTA frees its shared memory the moment it retires. Three users:
//   k_level_strip<TH, BLUR, FAST>   7x7 Gaussian (cv::GaussianBlur, src/ORBextractor.cc:799) and/or the dense 4-pair FAST upper
//                                   bound (the rejection stage of DetectFAST, :489-540) of a 128 x TH tile — both need the same
//                                   3-pixel halo; they bind different pipes (IDP vs VIMNMX/PRMT), so the fused form interleaves well
//   k_pyramid_strip<TH>             cv::resize INTER_LINEAR of a 128 x TH output tile (ComputePyramid, :455-470)
//   k_fast_cells2                   per-cell exact scoring, non-maximum suppression, iniTh -> minTh retry and ordered emit, reading
//                                   the bound bitmaps the dense pass wrote

#define ST_TW 128                 // tile width in pixels: lane l owns columns [4l, 4l + 4)
#define ST_BW 160                 // TMA box width: 16 bytes of left halo (u8 boxes start 16-byte aligned), the tile, 16 bytes right
#define ST_BWW (ST_BW / 4)
#define ST_HALO 3

// one = 1, a value the compiler cannot fold (see FastRowOps); (xorg, yorg) = origin of the tile grid: (0, 0) for the blur, the 16-byte
// aligned column / the first row of the cells' interiors for the FAST bound, whose tiles then cover only what DetectFAST looks at
struct OrbxStripTiles { int base[ORBX_MAX_LEVELS + 1]; int tx[ORBX_MAX_LEVELS]; int one, xorg, yorg; int first; };   // first: tile index of blockIdx.x = 0 (launches over a level range)

__host__ __device__ constexpr int st_tile_bytes(int th) { return ((ST_BW * (th + 2 * ST_HALO) + 127) / 128) * 128; }

//@phase blur arithmetic (horizontal IDP.4A, vertical IDP.2A, pack)
// ---- 7-tap Gaussian, OpenCV 8.8 fixed point: K = {18, 34, 48, 56, 48, 34, 18} / 256, h = sum K s (16 bits, exact),
//      v = sum K h (< 2^24), out = (v + 2^15) >> 16. Nothing is rounded in between, so any evaluation order is exact.
// horizontal pass of one row for the lane's 4 pixels; W0 W1 W2 = columns x-4..x-1, x..x+3, x+4..x+7
__device__ __forceinline__ void blur_hrow(uint32_t W0, uint32_t W1, uint32_t W2, uint32_t (&h)[4])
{
	const uint32_t KA = 18u | (34u << 8) | (48u << 16) | (56u << 24), KB = 48u | (34u << 8) | (18u << 16);
	h[0] = __dp4a(__funnelshift_r(W0, W1, 8), KA, __dp4a(__funnelshift_r(W1, W2, 8), KB, 0u));
	h[1] = __dp4a(__funnelshift_r(W0, W1, 16), KA, __dp4a(__funnelshift_r(W1, W2, 16), KB, 0u));
	h[2] = __dp4a(__funnelshift_r(W0, W1, 24), KA, __dp4a(__funnelshift_r(W1, W2, 24), KB, 0u));
	h[3] = __dp4a(W1, KA, __dp4a(W2, KB, 0u));
}
// vertical pass over four row pairs p0..p3 (16-bit lanes: row 2k | row 2k+1 << 16). FIRST: the output row whose window starts on
// the low half of p0 (rows 0..6 of the 8), else the one that starts on its high half (rows 1..7).
template <bool FIRST>
__device__ __forceinline__ uint32_t blur_vrow(const uint32_t (&p0)[4], const uint32_t (&p1)[4], const uint32_t (&p2)[4], const uint32_t (&p3)[4])
{
	const uint32_t c01 = FIRST ? (18u | (34u << 8) | (48u << 16) | (56u << 24)) : ((18u << 8) | (34u << 16) | (48u << 24));
	const uint32_t c23 = FIRST ? (48u | (34u << 8) | (18u << 16)) : (56u | (48u << 8) | (34u << 16) | (18u << 24));
	uint32_t v[4];
#pragma unroll
	for (int j = 0; j < 4; j++)
	{
		uint32_t a = __dp2a_lo(p0[j], c01, 32768u);
		a = __dp2a_hi(p1[j], c01, a);
		a = __dp2a_lo(p2[j], c23, a);
		v[j] = __dp2a_hi(p3[j], c23, a);
	}
	// byte 2 of each v
	return __byte_perm(__byte_perm(v[0], v[1], 0x6262), __byte_perm(v[2], v[3], 0x6262), 0x5410);
}

//@phase FAST bound arithmetic (operands, pair max/min, thresholds, flag gather)
// ---- dense FAST bound for the lane's 4 pixels of one row (same arithmetic as the in-cell bound pass it replaces): rows y-3 (m3),
//      y-2, y, y+2, y+3 (p3). Returns 8 bits: bits 0..3 = U > iniTh for pixels 0..3, bits 4..7 = U > minTh.
//      The kernel is bound by the ALU pipe (VIMNMX, PRMT, LOP3, SHF, IADD3 all issue there; ncu: 91 % ALU, 6 % FMA), so everything
//      that has an integer multiply-add form is written as one: `one` / `mone` are 1 and -1 the compiler cannot see, which turns
//      a + b and a - b into IMADs (FMA pipe). In particular min(a, b) = a + b - max(a, b): exact as a 32-bit word for packed 16-bit
//      lanes (the final lanes are in range, modular arithmetic absorbs the borrows in between), one VIMNMX per pair instead of two.
struct FastRowOps
{
	uint32_t one, mone, c256, kini, kdelta;
	__device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) const { return a * one + b; }
	__device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) const { return b * mone + a; }      // a - b
};
// operands of one row in its three roles. D[0..3]: the row's bytes moved two columns left / right, both parities, used when the row
// is two above or two below the centre: (x+2: par 0, par 1), (x-2: par 0, par 1)
__device__ __forceinline__ void fast_row_diag(uint32_t wa, uint32_t wb, uint32_t wc, uint32_t (&D)[4])
{
	D[0] = __byte_perm(wb, wc, 0x5432); D[1] = __byte_perm(wb, wc, 0x4321);
	D[2] = __byte_perm(wa, wb, 0x5432); D[3] = __byte_perm(wa, wb, 0x4321);
}
// raw form: the iniTh flags in bits 28..31, the minTh flags in bits 24..27, partial products below
__device__ __forceinline__ uint32_t fast_bound_row4_raw(const FastRowOps& K, uint32_t m3, const uint32_t (&Dm)[4], uint32_t c0, uint32_t c1, uint32_t c2,
                                                        const uint32_t (&Dp)[4], uint32_t p3)
{
	uint32_t f[2];
#pragma unroll
	for (int par = 0; par < 2; par++)
	{
		// par 0: pixels 1 and 3 sit in the high bytes of the two 16-bit lanes; par 1: pixels 0 and 2 (operands one byte further left)
		const uint32_t a1 = par == 0 ? p3 : p3 * K.c256;                                            // ( 0, +3)
		const uint32_t a2 = par == 0 ? m3 : m3 * K.c256;                                            // ( 0, -3)
		const uint32_t b1 = Dp[par], b2 = Dm[2 + par];                                            // (+2, +2), (-2, -2)
		const uint32_t d1 = par == 0 ? __byte_perm(c1, c2, 0x6543) : __byte_perm(c1, c2, 0x5432); // (+3,  0)
		const uint32_t d2 = par == 0 ? __byte_perm(c0, c1, 0x4321) : c0;                          // (-3,  0)
		const uint32_t e1 = Dm[par], e2 = Dp[2 + par];                                            // (+2, -2), (-2, +2)
		const uint32_t xa = __vmaxu2(a1, a2), xb = __vmaxu2(b1, b2), xd = __vmaxu2(d1, d2), xe = __vmaxu2(e1, e2);
		const uint32_t na = K.sub(K.add(a1, a2), xa), nb = K.sub(K.add(b1, b2), xb), nd = K.sub(K.add(d1, d2), xd), ne = K.sub(K.add(e1, e2), xe);
		const uint32_t hi = __vminu2(__vimin3_u16x2(xa, xb, xd), xe);      // min_k max(pair)
		const uint32_t lo = __vmaxu2(__vimax3_u16x2(na, nb, nd), ne);      // max_k min(pair)
		const uint32_t H = __byte_perm(hi, 0, 0x4341), Lo = __byte_perm(lo, 0, 0x4341);
		const uint32_t C = par == 0 ? __byte_perm(c1, 0, 0x4341) : __byte_perm(c1, 0, 0x4240);
		// lanes: (H - c) + K and (c - Lo) + K with K = 0x7fff - iniTh: bit 15 of a lane <=> U > iniTh; the final lanes neither borrow nor carry
		f[par] = __vmaxu2(K.sub(K.add(H, K.kini), C), K.sub(K.add(C, K.kini), Lo));
	}
	const uint32_t g0 = K.add(f[0], K.kdelta), g1 = K.add(f[1], K.kdelta);     // kdelta = iniTh - minTh per lane: bit 15 <=> U > minTh
	// pixel order 0,1,2,3 = f[1].lo, f[0].lo, f[1].hi, f[0].hi: the sign bits land in bits 7,15,23,31 of one word (iniTh) and, moved
	// down by 4, in bits 3,11,19,27 (minTh); ONE multiply gathers all eight: bit 7 + 8j -> 24 + j, bit 3 + 8j -> 20 + j ... the partial
	// products fall on distinct bits, so nothing carries
	const uint32_t fi = __byte_perm(f[1], f[0], 0x7351) & 0x80808080u, gi = (__byte_perm(g1, g0, 0x7351) >> 4) & 0x08080808u;
	return (fi | gi) * 0x00204081u;                  // bit 7 + 8j -> 28 + j (iniTh), bit 3 + 8j -> 24 + j (minTh)
}
__device__ __forceinline__ uint32_t fast_bound_row4(const FastRowOps& K, uint32_t m3, const uint32_t (&Dm)[4], uint32_t c0, uint32_t c1, uint32_t c2,
                                                    const uint32_t (&Dp)[4], uint32_t p3)
{
	return fast_bound_row4_raw(K, m3, Dm, c0, c1, c2, Dp, p3) >> 24;      // low nibble = minTh flags, high nibble = iniTh flags
}

//@phase row walk: loads, window rotation, stores
// Row walk of one tile. CHECK = false: every output row of the tile exists and (FAST) lies inside the cells' rows, so the loop has a
// static trip count and no per-row predicate; CHECK = true: the last tile row of a level / the tiles that straddle the first or last
// cell row.
template <int TH, bool DO_BLUR, bool DO_FAST, bool CHECK>
__device__ __forceinline__ void strip_rows(const uint32_t* __restrict__ tw, uint8_t* __restrict__ bdst, const int64_t pitch, const bool bstore,
                                           uint8_t* __restrict__ fdst, const int64_t pitch8, const FastRowOps K, const uint32_t selx,
                                           const int lane, const int nrows, const int f0, const int f1)
{
	//@phase row walk: loads, window rotation, stores
	uint32_t R[8][3];            // FAST: raw words of the last 8 rows (left / right words are dead once the row has been the centre)
	uint32_t Dg[8][4];           // FAST: their two-column shifts (fast_row_diag), made when a row enters, dead once it is two above the centre
	uint32_t Pp[4][4];           // blur: horizontal sums of the last 4 row pairs
	auto load_pair = [&](const uint32_t* q, const int rs, const int slot) {      // box rows at q, q + ST_BWW into row slots rs, rs + 1 and pair slot `slot`
		uint32_t he[4], ho[4];
		{
			const uint32_t W0 = q[0], W1 = q[1], W2 = q[2];
			if (DO_FAST) { R[rs][0] = W0; R[rs][1] = W1; R[rs][2] = W2; fast_row_diag(W0, W1, W2, Dg[rs]); }
			if (DO_BLUR) blur_hrow(W0, W1, W2, he);
		}
		{
			const uint32_t W0 = q[ST_BWW], W1 = q[ST_BWW + 1], W2 = q[ST_BWW + 2];
			if (DO_FAST) { R[rs + 1][0] = W0; R[rs + 1][1] = W1; R[rs + 1][2] = W2; fast_row_diag(W0, W1, W2, Dg[rs + 1]); }
			if (DO_BLUR) blur_hrow(W0, W1, W2, ho);
		}
		if (DO_BLUR)
		{
#pragma unroll
			for (int j = 0; j < 4; j++) Pp[slot][j] = __byte_perm(he[j], ho[j], 0x5410);
		}
	};
	// FAST flags of the output row whose centre is window slot c
	auto fast_row = [&](const int c, uint8_t* dst) {
		const uint32_t v = fast_bound_row4(K, R[(c + 5) & 7][1], Dg[(c + 6) & 7], R[c][0], R[c][1], R[c][2], Dg[(c + 2) & 7], R[(c + 3) & 7][1]);
		const uint32_t o = __shfl_xor_sync(0xffffffffu, v, 1);
		// X = flags of the pair's lower 4 pixels | flags of its upper 4 pixels << 8; even lanes keep the iniTh (high) nibbles, odd lanes the minTh ones
		uint32_t X = __byte_perm(v, o, selx);
		X >>= 4 * ((lane & 1) ^ 1);
		*dst = (uint8_t)((X & 0xfu) | ((X >> 4) & 0xf0u));
	};

	// warm-up: box rows 0..5
	load_pair(tw, 0, 0); load_pair(tw + 2 * ST_BWW, 2, 1); load_pair(tw + 4 * ST_BWW, 4, 2);
	tw += 6 * ST_BWW;
#pragma unroll 1
	for (int it = 0; it < TH / 8; it++)
	{
		if (CHECK && 8 * it >= nrows) break;
#pragma unroll
		for (int kk = 0; kk < 4; kk++)
		{
			// newest pair: box rows 6 + 8 it + 2 kk (+1); it completes the windows of output rows t = 8 it + 2 kk and t + 1
			const int t = 8 * it + 2 * kk;
			load_pair(tw + 2 * kk * ST_BWW, (6 + 2 * kk) & 7, (3 + kk) & 3);
			if (DO_BLUR)
			{
				const uint32_t o0 = blur_vrow<true>(Pp[kk & 3], Pp[(kk + 1) & 3], Pp[(kk + 2) & 3], Pp[(kk + 3) & 3]);
				const uint32_t o1 = blur_vrow<false>(Pp[kk & 3], Pp[(kk + 1) & 3], Pp[(kk + 2) & 3], Pp[(kk + 3) & 3]);
				if (bstore)
				{
					if (!CHECK || t < nrows) *reinterpret_cast<uint32_t*>(bdst) = o0;          // pitch is a multiple of 128: in-row padding absorbs the tail
					if (!CHECK || t + 1 < nrows) *reinterpret_cast<uint32_t*>(bdst + pitch) = o1;
				}
				bdst += 2 * pitch;
			}
			if (DO_FAST)
			{
				// centre of output row t is box row t + 3 -> window slot (2 kk + 3) & 7
				if (!CHECK || (t >= f0 && t < f1)) fast_row((2 * kk + 3) & 7, fdst);
				if (!CHECK || (t + 1 >= f0 && t + 1 < f1)) fast_row((2 * kk + 4) & 7, fdst + pitch8);
				fdst += 2 * pitch8;
			}
		}
		tw += 8 * ST_BWW;
	}
}

template <int TH, bool DO_BLUR, bool DO_FAST>
__global__ void __launch_bounds__(32) k_level_strip(const OrbxPlanDev P, const __grid_constant__ OrbxStripMaps maps, const OrbxStripTiles T)
{
	//@phase tile header: level lookup, TMA issue, pointers, border patch
	static_assert(TH % 8 == 0, "the row window rotates with period 8");
	extern __shared__ __align__(128) uint8_t st_smem[];
	uint64_t* const bar = reinterpret_cast<uint64_t*>(st_smem + st_tile_bytes(TH));
	const int lane = threadIdx.x, f = blockIdx.y;

	// which level does this tile belong to (unrolled: every index into T is a compile-time constant)
	const int gtile = (int)blockIdx.x + T.first;
	int level = 0, base = 0, tx = T.tx[0];
#pragma unroll
	for (int s = 1; s < ORBX_MAX_LEVELS; s++)
		if (gtile >= T.base[s] && T.base[s + 1] > T.base[s]) { level = s; base = T.base[s]; tx = T.tx[s]; }
	const int tile = gtile - base;
	const int tile_y = tile / tx, tile_x = tile - tile_y * tx;
	const OrbxLevel& L = P.lv[level];
	const int w = L.w, h = L.h;
	const int x0 = tile_x * ST_TW + T.xorg, y0 = tile_y * TH + T.yorg;
	if (lane == 0)
	{
		mbar_init(bar, 1);
		mbar_expect_tx(bar, (unsigned)(ST_BW * (TH + 2 * ST_HALO)));
		tma_load_3d(st_smem, &maps.level[level], x0 - 16, y0 - ST_HALO, P.frame0 + f, bar);   // rows/columns outside the level arrive as zeros
	}
	const int x = x0 + 4 * lane;
	const int64_t pitch = L.pitch, pitch8 = L.pitch >> 3;
	uint8_t* __restrict__ bdst = nullptr;
	if (DO_BLUR) bdst = P.blur + (int64_t)f * P.slab + L.offset + (int64_t)y0 * pitch + x;
	// bound bitmaps: bit x of row y, one byte per lane pair; even lanes write the iniTh map, odd lanes the minTh map
	uint8_t* __restrict__ fdst = nullptr;
	FastRowOps K = {};
	uint32_t selx = 0;
	int f0 = 0, f1 = 0;
	if (DO_FAST)
	{
		fdst = ((lane & 1) ? P.fmap_min : P.fmap_ini) + (int64_t)f * (P.slab >> 3) + (L.offset >> 3) + (int64_t)y0 * pitch8 + (x0 >> 3) + (lane >> 1);
		K.one = (uint32_t)T.one; K.mone = 0u - K.one; K.c256 = 256u * K.one;
		K.kini = (uint32_t)(0x7fff - P.ini_th) * 0x00010001u; K.kdelta = (uint32_t)(P.ini_th - P.min_th) * 0x00010001u;
		selx = (lane & 1) ? 0x1104u : 0x1140u;
		// tile rows that belong to some cell's interior (cv::FAST skips a 3-pixel border of the cell view)
		f0 = L.miny + 3 - y0; f1 = L.maxy - 3 - y0;
	}
	const int nrows = min(TH, h - y0);            // output rows of this tile
	__syncwarp();
	mbar_wait(bar, 0);

	if (DO_BLUR)
	{
		// BORDER_REFLECT_101 is made physical in the staged tile, so the row walk below never looks at a border.
		// Rows: box row i holds image row y0 - 3 + i; rows -3..-1 are rows 3..1, rows h..h+2 are rows h-2..h-4 (all inside the box).
		if (y0 == 0 && lane < 30)
		{
			const int r = lane / 10, c = lane - 10 * r;     // rows -1-r <- 1+r: box row 2 - r <- box row 4 + r; ten 16-byte chunks per row
			reinterpret_cast<uint4*>(st_smem + (2 - r) * ST_BW)[c] = reinterpret_cast<const uint4*>(st_smem + (4 + r) * ST_BW)[c];
		}
		if (y0 + TH + ST_HALO > h && lane < 30)
		{
			const int r = lane / 10, c = lane - 10 * r;     // row h + r <- row h - 2 - r
			const int bd = h + r - (y0 - ST_HALO), bs = h - 2 - r - (y0 - ST_HALO);
			if (bd < TH + 2 * ST_HALO) reinterpret_cast<uint4*>(st_smem + bd * ST_BW)[c] = reinterpret_cast<const uint4*>(st_smem + bs * ST_BW)[c];
		}
		__syncwarp();
		// Columns: -1 -> 1 ... and w -> w-2 ...; the mirrored pixels are inside the same staged row
		if (x0 == 0 || w < x0 + ST_TW + ST_HALO)
		{
			for (int i = lane; i < TH + 2 * ST_HALO; i += 32)
			{
				uint8_t* row = st_smem + i * ST_BW + 16 - x0;        // row[c] = image column c
				if (x0 == 0) { row[-1] = row[1]; row[-2] = row[2]; row[-3] = row[3]; }
				if (w < x0 + ST_TW + ST_HALO)
				{
#pragma unroll
					for (int k = 0; k < 3; k++) row[w + k] = row[w - 2 - k];     // w + 2 < x0 + 144: inside the box
				}
			}
			__syncwarp();
		}
	}

	const uint32_t* __restrict__ tw = reinterpret_cast<const uint32_t*>(st_smem) + 3 + lane;   // word of columns x-4..x-1 of box row 0
	const bool full = nrows == TH && (!DO_FAST || (f0 <= 0 && f1 >= TH));
	if (full) strip_rows<TH, DO_BLUR, DO_FAST, false>(tw, bdst, pitch, x < w, fdst, pitch8, K, selx, lane, nrows, f0, f1);
	else strip_rows<TH, DO_BLUR, DO_FAST, true>(tw, bdst, pitch, x < w, fdst, pitch8, K, selx, lane, nrows, f0, f1);
}

//@end
// =====================================================================================================
// k_pyramid_strip — cv::resize INTER_LINEAR 8UC1 in OpenCV's 11-bit fixed point for one 128 x TH output tile.
// The source rectangle of the tile is one TMA box. A lane produces 4 adjacent output pixels per row and walks down the tile:
// per source row 3 aligned words -> 2 funnel shifts put the lane's first source column at byte 0 -> one PRMT per column picks the
// (s[x], s[x+1]) byte pair -> one IDP.2A against the packed (a0, a1) coefficients. A source row's horizontal pass is kept while the
// 1-2 output rows that use it are produced. Box width <= 256 bytes limits this kernel to scale factors <= ~1.8; beyond that the
// cp.async kernel (k_pyramid_resize) runs.
// =====================================================================================================
// one 128 x TH output tile of `level` of frame f; wait_src: spin (lane 0) until the source level is complete, see k_pyramid_all
template <int TH>
__device__ __forceinline__ void pyramid_tile(const OrbxPlanDev& P, const OrbxPyrMaps& maps, uint8_t* py_smem, const int level, const int bw, const int bh,
                                             const int tile_x, const int tile_y, const int f, const int* wait_src, const int wait_for)
{
	static_assert(TH <= 32, "lane k holds the table entry of tile row k");
	const OrbxLevel& D = P.lv[level];
	uint64_t* const bar = reinterpret_cast<uint64_t*>(py_smem + ((bw * bh + 127) & ~127));
	const int lane = threadIdx.x;
	const int sh = P.lv[level - 1].h;
	const int dx0 = tile_x * ST_TW, dy0 = tile_y * TH;
	const int* __restrict__ yofs = P.yofs + D.ytab_base;
	const int* __restrict__ xofs = P.xofs + D.xtab_base;
	const int s_lo = __ldg(yofs + dy0);
	const int xa = __ldg(xofs + dx0) & ~15;
	if (lane == 0)
	{
		mbar_init(bar, 1);
		mbar_expect_tx(bar, (unsigned)(bw * bh));
		if (wait_src)
		{
			// the source level is written by other CTAs of this launch: acquire its completion count, then order the copy engine's reads
			// (async proxy) behind it. A lost count must not hang the device.
			int seen, spins = 0;
			do
			{
				asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(seen) : "l"(wait_src) : "memory");
				if (seen < wait_for && ++spins > (1 << 22)) __trap();
			} while (seen < wait_for);
			asm volatile("fence.proxy.async.global;\n" ::: "memory");
		}
		tma_load_3d(py_smem, &maps.src[level], xa, s_lo, P.frame0 + f, bar);
	}
	// the lane's 4 columns: byte offset of the first source column, PRMT selectors of the (s[x], s[x+1]) pairs, packed coefficients
	int sx[4];
	uint32_t coef[4], sel[4];
#pragma unroll
	for (int j = 0; j < 4; j++)
	{
		const int dx = min(dx0 + 4 * lane + j, D.w - 1);      // columns past the edge repeat the last one; they land in row padding
		sx[j] = __ldg(xofs + dx) - xa;
		const short2 a = __ldg(P.xcoef + D.xtab_base + dx);
		coef[j] = (uint32_t)(uint16_t)a.x | ((uint32_t)(uint16_t)a.y << 16);
	}
#pragma unroll
	for (int j = 0; j < 4; j++)
	{
		const int e = sx[j] - sx[0];                          // 0..6 (host-checked); where s[x+1] is clamped to the last column its coefficient is 0
		sel[j] = (uint32_t)e | ((uint32_t)(e + 1) << 4) | 0x4400u;
	}
	const int wofs = sx[0] >> 2, shb = (sx[0] & 3) * 8;
	const int nsrc = min(bh, sh - s_lo);                      // source rows of the box that exist
	int my_r = 0, my_b = 0;
	if (lane < TH)
	{
		const int dy = min(dy0 + lane, D.h - 1);
		my_r = __ldg(yofs + dy) - s_lo;
		const short2 b = __ldg(P.ycoef + D.ytab_base + dy);
		my_b = (int)(uint16_t)b.x | ((int)b.y << 16);
	}
	uint8_t* __restrict__ dst = P.pyr + (int64_t)f * P.slab + D.offset + (int64_t)dy0 * D.pitch + dx0 + 4 * lane;
	const bool store = dx0 + 4 * lane < D.w;
	const int nrows = min(TH, D.h - dy0);
	const int bww = bw >> 2;
	const int64_t dpitch = D.pitch;
	__syncwarp();
	mbar_wait(bar, 0);

	const uint32_t* __restrict__ tw = reinterpret_cast<const uint32_t*>(py_smem) + wofs;
	auto hrow = [&](int r, int (&hh)[4]) {
		const uint32_t* q = tw + r * bww;
		const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
		const uint32_t u0 = __funnelshift_r(w0, w1, shb), u1 = __funnelshift_r(w1, w2, shb);
#pragma unroll
		for (int j = 0; j < 4; j++) hh[j] = (int)__dp2a_lo(coef[j], __byte_perm(u0, u1, sel[j]), 0u) >> 4;
	};
	int rc = -2, h0[4], h1[4];          // h0 = source row rc, h1 = row min(rc + 1, last)
#pragma unroll 1
	for (int k = 0; k < nrows; k++)
	{
		const int r = __shfl_sync(0xffffffffu, my_r, k), bwd = __shfl_sync(0xffffffffu, my_b, k);
		if (r != rc)
		{
			if (r == rc + 1)
			{
#pragma unroll
				for (int j = 0; j < 4; j++) h0[j] = h1[j];
			}
			else hrow(r, h0);
			hrow(min(r + 1, nsrc - 1), h1);     // at the last source row both taps are that row (its second coefficient is 0)
			rc = r;
		}
		const int b0 = (int)(short)(bwd & 0xffff), b1 = bwd >> 16;
		int s[4];
#pragma unroll
		for (int j = 0; j < 4; j++) s[j] = ((b0 * h0[j] + (2 << 16)) >> 16) + ((b1 * h1[j]) >> 16);   // = (b0 h0 >> 16) + (b1 h1 >> 16) + 2, in [0, 1023]
		const uint32_t q01 = __byte_perm((uint32_t)s[0], (uint32_t)s[1], 0x5410) >> 2, q23 = __byte_perm((uint32_t)s[2], (uint32_t)s[3], 0x5410) >> 2;
		if (store) *reinterpret_cast<uint32_t*>(dst) = __byte_perm(q01, q23, 0x6420);   // pitch is a multiple of 128: in-row padding absorbs the tail
		dst += dpitch;
	}
}

template <int TH>
__global__ void __launch_bounds__(32) k_pyramid_strip(const OrbxPlanDev P, const __grid_constant__ OrbxPyrMaps maps, const int level, const int bw, const int bh)
{
	extern __shared__ __align__(128) uint8_t py_smem[];
	pyramid_tile<TH>(P, maps, py_smem, level, bw, bh, blockIdx.x, blockIdx.y, blockIdx.z, nullptr, 0);
}

// ComputePyramid as ONE launch: the tiles of levels 1 .. n-1 of every frame, level-major (every frame's level 1, then every frame's
// level 2 ...). A tile of level s >= 2 needs level s - 1 of its own frame: the CTAs that produced it count themselves into
// done[frame][s - 1] (release), the consumer's lane 0 spins on that count (acquire) before it issues its TMA load. CTAs are dispatched in
// blockIdx order, so the producers of a tile are resident or finished before the tile starts: with 512 frames the count is there long
// before it is needed, and for one frame the whole grid is resident at once. Seven dependent launches (and their tails: the three
// smallest levels took 11 us each for 4 us of work) become one; a one-frame call saves six launch latencies.
struct OrbxPyrTiles { int base[ORBX_MAX_LEVELS + 1]; int tx[ORBX_MAX_LEVELS]; int which; int first; };   // first: tiles per frame of the levels before this launch's first level   // base[s]: tiles per frame of the levels before s (base[1] = 0)
template <int TH>
__global__ void __launch_bounds__(32) k_pyramid_all(const OrbxPlanDev P, const __grid_constant__ OrbxPyrMaps maps, const OrbxPyrTiles T, int* __restrict__ done)
{
	extern __shared__ __align__(128) uint8_t py_smem[];
	const unsigned F = (unsigned)P.frames, id = blockIdx.x + (unsigned)T.first * F;
	int level = 1;
#pragma unroll
	for (int s = 2; s < ORBX_MAX_LEVELS; s++)
		if (s < P.nlevels && id >= (unsigned)T.base[s] * F) level = s;
	const unsigned nt = (unsigned)(T.base[level + 1] - T.base[level]);
	const unsigned rem = id - (unsigned)T.base[level] * F;
	const unsigned f = rem / nt, tile = rem - f * nt;
	const int tile_y = (int)(tile / (unsigned)T.tx[level]), tile_x = (int)tile - tile_y * T.tx[level];
	int* const cnt = done + (int64_t)(P.frame0 + (int)f) * ORBX_MAX_LEVELS;
	const OrbxLevel& D = P.lv[level];
	pyramid_tile<TH>(P, maps, py_smem, level, D.py_bw[T.which], D.py_bh[T.which], tile_x, tile_y, (int)f, level > 1 ? cnt + level - 1 : nullptr,
	                 T.base[level] - T.base[level - 1]);
	if (level + 1 < P.nlevels)
	{
		__syncwarp();                                   // the warp's stores happen before lane 0's release
		if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;\n" ::"l"(cnt + level) : "memory");
	}
}

// =====================================================================================================
// k_fast_cells2 — DetectFAST (src/ORBextractor.cc:489-540) per cell, behind the dense bound pass: one warp per cell reads the
// rows of the two bound bitmaps that cover its region (wa: U > iniTh, wb: minTh < U <= iniTh), computes the exact arc score only
// for those pixels, finds strict 8-neighbour maxima above iniTh, retries at minTh when the cell has none (:526-530), and emits
// in cv::FAST's row-major order. The cell view comes in by one TMA tile load, as before.
// =====================================================================================================
#define FT_LIST_CAP 512          // entries of the pixel list in throughput launches. A cell of the synthetic frames flags 80 (U > iniTh) to 400 (U > minTh,
                                 // weak texture) pixels; the worst case, every pixel of a 37 x 34 cell, would take 2.5 KB per warp: 26 instead of 32
                                 // resident warps. A cell with more (noise images) records itself and is redone by k_fast_cells2_overflow.
struct OrbxCellLayout
{
	int score_stride;            // bytes per score row (region width + 2, rounded up to 8)
	int off_score, off_list, off_bm, off_bar;
	int list_cap;                // entries the list holds
	int warp_bytes;
};

// One cell. REDO = false: the usual launch; a cell whose flagged pixels do not fit the list appends itself to P.ovf_list and returns
// before it has written anything. REDO = true: the overflow launch, whose list holds a whole cell.
template <bool REDO>
__device__ __forceinline__ void fast_cell(const OrbxPlanDev& P, const OrbxTmaMaps& maps, const OrbxCellLayout& Y, uint8_t* fw_smem, const int cell, const int f,
                                          const unsigned parity)
{
	//@phase cell table, TMA issue
	const int lane = threadIdx.x;
	uint8_t* const tile = fw_smem;
	uint8_t* const score = fw_smem + Y.off_score;
	uint16_t* const list = reinterpret_cast<uint16_t*>(fw_smem + Y.off_list);
	uint32_t* const bm_sel = reinterpret_cast<uint32_t*>(fw_smem + Y.off_bm);   // [row][2]: 64 bits per region row
	uint64_t* const tma_bar = reinterpret_cast<uint64_t*>(fw_smem + Y.off_bar);
	const int SS = Y.score_stride;

	const int4 ct = __ldg(P.cell_tab + cell);
	const int x0 = ct.x & 0xffff, y0 = ct.x >> 16, vw = ct.y & 0xffff, vh = ct.y >> 16, lvl = ct.z, c = ct.w;
	const OrbxLevel& L = P.lv[lvl];
	const int rw = vw - 6, rh = vh - 6;
	const int sh = x0 & 15;                           // the TMA box starts 16-byte aligned
	if (lane == 0)
	{
		if (!REDO) mbar_init(tma_bar, 1);             // the overflow kernel walks a list of cells: it initialises the barrier once
		mbar_expect_tx(tma_bar, (unsigned)(FT_TS * maps.box_h[lvl]));
		tma_load_3d(tile, &maps.level[lvl], x0 - sh, y0, P.frame0 + f, tma_bar);
	}
	//@phase bound bitmaps of the region (loads, funnel shifts)
	// the region's rows of the bound bitmaps: lane r holds rows r and r + 32 as 64-bit masks (bit i = region column i)
	uint32_t wa[4], wb[4];
	{
		const int X0 = x0 + 3, bsh = X0 & 31;
		const int64_t fbase = (int64_t)f * (P.slab >> 3) + (L.offset >> 3);
		const uint32_t* __restrict__ mi = reinterpret_cast<const uint32_t*>(P.fmap_ini + fbase) + (X0 >> 5);
		const uint32_t* __restrict__ mm = reinterpret_cast<const uint32_t*>(P.fmap_min + fbase) + (X0 >> 5);
		const int p32 = L.pitch >> 5;
		const uint64_t rowmask = rw >= 64 ? ~0ull : ((1ull << rw) - 1ull);
#pragma unroll
		for (int k = 0; k < 2; k++)
		{
			const int r = lane + 32 * k;
			uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 0;
			if (r < rh)
			{
				const uint32_t* pi = mi + (int64_t)(y0 + 3 + r) * p32;
				const uint32_t* pm = mm + (int64_t)(y0 + 3 + r) * p32;
				const uint32_t i0 = __ldg(pi), i1 = __ldg(pi + 1), i2 = __ldg(pi + 2);     // the maps carry 16 spare bytes behind the last row
				const uint32_t m0 = __ldg(pm), m1 = __ldg(pm + 1), m2 = __ldg(pm + 2);
				a0 = __funnelshift_r(i0, i1, bsh) & (uint32_t)rowmask; a1 = __funnelshift_r(i1, i2, bsh) & (uint32_t)(rowmask >> 32);
				b0 = __funnelshift_r(m0, m1, bsh) & (uint32_t)rowmask; b1 = __funnelshift_r(m1, m2, bsh) & (uint32_t)(rowmask >> 32);
			}
			wa[2 * k] = a0; wa[2 * k + 1] = a1;
			wb[2 * k] = b0 & ~a0; wb[2 * k + 1] = b1 & ~a1;        // minTh < U <= iniTh
		}
	}
	//@phase zero scores and survivor bitmap, wait for the tile
	{
		const int n8 = (rh + 2) * (SS >> 3);              // score rows -1 .. rh, 8 bytes at a time (off_score is 16-byte aligned, SS a multiple of 8)
#pragma unroll 1
		for (int i = lane; i < n8; i += 32) reinterpret_cast<uint2*>(score)[i] = make_uint2(0, 0);
		if (lane < rh) reinterpret_cast<uint2*>(bm_sel)[lane] = make_uint2(0, 0);
		if (lane + 32 < rh) reinterpret_cast<uint2*>(bm_sel)[lane + 32] = make_uint2(0, 0);
	}
	__syncwarp();
	mbar_wait(tma_bar, parity);

	const int tmin = P.min_th, tini = P.ini_th;
	const uint8_t* __restrict__ t0 = tile + 3 * FT_TS + sh + 3;

	//@phase warp scans
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
	auto scan1 = [&](int c0, int& o0, int& total) {
		int i0 = c0;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const int u0 = __shfl_up_sync(0xffffffffu, i0, d);
			if (lane >= d) i0 += u0;
		}
		total = __shfl_sync(0xffffffffu, i0, 31);
		o0 = i0 - c0;
	};
	//@phase candidate list build (expand)
	const bool tall = rh > 32;                        // warp-uniform: most plans have no cell taller than 32 rows
	// append the pixels of this lane's row words to the list as ry << 6 | rx; returns how many the warp appended
	auto expand = [&](const uint32_t* w, int at) {
		int o0, o1 = 0, total;
		if (tall) scan2(__popc(w[0]) + __popc(w[1]), __popc(w[2]) + __popc(w[3]), o0, o1, total);
		else scan1(__popc(w[0]) + __popc(w[1]), o0, total);
		if (!REDO && at + total > Y.list_cap) return -1;           // does not fit: nothing written
#pragma unroll
		for (int k = 0; k < 4; k++)
		{
			if (k >= 2 && !tall) break;
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
	//@phase exact-score loop around the network (evaluate)
	auto evaluate = [&](int from, int to) {
		for (int j = from + lane; j < to; j += 32)
		{
			const int e = list[j], ry = e >> 6, rx = e & 63;
			const int s = arc_score_packed(t0 + ry * FT_TS + rx);
			score[(ry + 1) * SS + rx + 1] = (uint8_t)max(s, 0);
		}
	};
	//@phase strict 8-neighbour maxima (select)
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

	//@phase iniTh pass, retry decision, minTh pass (control flow)
	// exact scores + maxima at iniTh; retry at minTh if the cell has no corner (:526-530)
	auto overflow = [&]() {                            // leave the cell to the overflow launch (nothing of it has been written yet)
		if (lane == 0) P.ovf_list[atomicAdd(P.ovf_count, 1)] = (uint32_t)f * (uint32_t)P.cells_per_frame + (uint32_t)cell;
	};
	const int n1 = expand(wa, 0);
	if (!REDO && n1 < 0) { overflow(); return; }
	__syncwarp();
	evaluate(0, n1);
	__syncwarp();
	if (!__any_sync(0xffffffffu, select(n1, tini)))
	{
		const int n2 = expand(wb, n1);
		if (!REDO && n2 < 0) { overflow(); return; }
		__syncwarp();
		evaluate(n1, n1 + n2);
		__syncwarp();
		select(n1 + n2, tmin);
	}
	__syncwarp();

	//@phase ordered emit
	// ordered emit (rows ascending, x ascending = cv::FAST's order inside the view)
	uint32_t ws[4];
#pragma unroll
	for (int k = 0; k < 2; k++)
	{
		const int r = lane + 32 * k;
		ws[2 * k] = r < rh ? bm_sel[2 * r] : 0u;
		ws[2 * k + 1] = r < rh ? bm_sel[2 * r + 1] : 0u;
	}
	int o0, o1 = 0, total;
	if (tall) scan2(__popc(ws[0]) + __popc(ws[1]), __popc(ws[2]) + __popc(ws[3]), o0, o1, total);
	else scan1(__popc(ws[0]) + __popc(ws[1]), o0, total);
	uint32_t* __restrict__ out = P.cand + (int64_t)f * P.cand_per_frame + L.cand_base + (int64_t)c * L.cell_cap;
#pragma unroll
	for (int k = 0; k < 4; k++)
	{
		if (k >= 2 && !tall) break;
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

__global__ void __launch_bounds__(32) k_fast_cells2(const OrbxPlanDev P, const __grid_constant__ OrbxTmaMaps maps, const OrbxCellLayout Y, const int cell0)
{
	extern __shared__ __align__(128) uint8_t fw_smem[];
	fast_cell<false>(P, maps, Y, fw_smem, (int)blockIdx.x + cell0, blockIdx.y, 0u);      // cell0: first cell of the launch's level range
}

// the cells the launch above left over: a fixed small grid walks the list (empty on ordinary images)
__global__ void __launch_bounds__(32) k_fast_cells2_overflow(const OrbxPlanDev P, const __grid_constant__ OrbxTmaMaps maps, const OrbxCellLayout Y)
{
	extern __shared__ __align__(128) uint8_t fw_smem[];
	const int n = *P.ovf_count;
	if ((int)blockIdx.x >= n) return;
	if (threadIdx.x == 0) mbar_init(reinterpret_cast<uint64_t*>(fw_smem + Y.off_bar), 1);
	__syncwarp();
	unsigned parity = 0;
	for (int i = blockIdx.x; i < n; i += gridDim.x, parity ^= 1u)
	{
		const uint32_t e = P.ovf_list[i];
		const int f = (int)(e / (uint32_t)P.cells_per_frame);
		fast_cell<true>(P, maps, Y, fw_smem, (int)(e - (uint32_t)f * (uint32_t)P.cells_per_frame), f, parity);
		__syncwarp();             // every lane is done with the tile before the next cell's load is issued
	}
}
//@end
