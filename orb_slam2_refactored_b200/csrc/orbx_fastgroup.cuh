// k_fast_groups — DetectFAST (src/ORBextractor.cc:489-540) for a GROUP of up to four horizontally adjacent cells in ONE warp, end to
// end (included by orbx_extract.cu inside its anonymous namespace, after orbx_strip.cuh).
//
// Round 2 ran DetectFAST as two launches: a dense bound pass over 128 x 32 tiles of the whole level (bitmaps to global memory) and a
// warp per ~30 px cell that read them back. 53 % of the cell kernel's instructions were per-cell bookkeeping (table, second TMA load,
// bitmap reload, zeroing, scans, list build, emit) and 17 % of the dense tiles lay outside any cell. Here the tile IS a run of cells:
// the interiors of the cells of one cell row tile the level without gaps (cell views overlap by 6 px, cv::FAST skips a 3-px border of
// its view), so a warp takes the interior of up to four neighbouring cells (<= 128 columns, <= 40 rows), stages it with one TMA load,
// walks the dense 4-pair bound down it (strip_rows, flags to shared memory), and then does everything the cell kernel did — candidate
// list, exact arc score, strict 8-neighbour maxima, iniTh -> minTh retry PER CELL (:526-530), ordered emit PER CELL — on the staged
// tile. Non-maximum suppression must not see across a cell boundary (each cell is its own cv::FAST call): a pixel in the first or last
// column of its cell drops the neighbours on that side. Nothing goes through global memory between the stages, the
// per-group fixed cost is shared by four cells, and the last list round is 1/11 instead of 1/3 of the exact-score work.

#define FG_TH 40                 // most interior rows of a cell the kernel takes; plans with taller cells use the two-launch path
#define FG_SS ST_BW              // bytes per score row = the tile's row stride: one offset addresses a pixel and its score
#define FG_LIST_HALF 864         // a list chunk is closed once it holds this many entries; a lane adds at most 4 x 40 more
#define FG_LIST_CAP (FG_LIST_HALF + 4 * FG_TH)

struct OrbxGroupLayout
{
	int off_bits, off_sel, off_score, off_list, off_bar;   // byte offsets behind the staged tile
	int bytes;
};

//@phase row walk: loads, window rotation, stores
// Dense 4-pair bound of the group's rows (the arithmetic of strip_rows, orbx_strip.cuh). A lane keeps the flags of its four columns:
// after every 8 rows one word per threshold goes to shared memory (bit 4 * (row & 7) + j = column 4 * lane + j). The group's first
// column is any column: every row word is funnel-shifted into place (shb = 8 * byte offset). Rows come in pairs; a pair past the last
// row ends the walk (an odd last row computes one row of garbage flags, masked by the caller).
__device__ __forceinline__ void fast_walk(const uint32_t* __restrict__ tw, const int shb, uint2* __restrict__ cb, const FastRowOps K, const int nrows)
{
	uint32_t R[8][3], Dg[8][4];
	auto load_row = [&](const uint32_t* q, const int rs) {
		const uint32_t q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
		const uint32_t W0 = __funnelshift_r(q0, q1, shb), W1 = __funnelshift_r(q1, q2, shb), W2 = __funnelshift_r(q2, q3, shb);
		R[rs][0] = W0; R[rs][1] = W1; R[rs][2] = W2;
		fast_row_diag(W0, W1, W2, Dg[rs]);
	};
	uint32_t ai = 0, am = 0;
	auto row = [&](const int c) {      // window slot c is the centre
		const uint32_t m = fast_bound_row4_raw(K, R[(c + 5) & 7][1], Dg[(c + 6) & 7], R[c][0], R[c][1], R[c][2], Dg[(c + 2) & 7], R[(c + 3) & 7][1]);
		ai = (ai >> 4) | (m & 0xf0000000u);
		am = (am >> 4) | ((m << 4) & 0xf0000000u);
	};
#pragma unroll
	for (int i = 0; i < 6; i++) load_row(tw + i * ST_BWW, i);
	tw += 6 * ST_BWW;
	int it = 0, rem = 8;
#pragma unroll 1
	for (;; it++)
	{
#pragma unroll
		for (int kk = 0; kk < 4; kk++)
		{
			if (8 * it + 2 * kk >= nrows) { rem = 8 - 2 * kk; goto done; }
			load_row(tw + 2 * kk * ST_BWW, (6 + 2 * kk) & 7);
			load_row(tw + (2 * kk + 1) * ST_BWW, (7 + 2 * kk) & 7);
			row((2 * kk + 3) & 7);
			row((2 * kk + 4) & 7);
		}
		cb[32 * it] = make_uint2(ai, am);
		tw += 8 * ST_BWW;
	}
done:
	if (rem < 8) cb[32 * it] = make_uint2(ai >> (4 * rem), am >> (4 * rem));
}

__global__ void __launch_bounds__(32) k_fast_groups(const OrbxPlanDev P, const __grid_constant__ OrbxGroupMaps maps, const OrbxGroupLayout Y, const int one)
{
	//@phase group table, TMA issue
	extern __shared__ __align__(128) uint8_t fg_smem[];
	const int lane = threadIdx.x, f = blockIdx.y;
	uint8_t* const tile = fg_smem;
	uint2* const colbits = reinterpret_cast<uint2*>(fg_smem + Y.off_bits);        // [rows / 8][lane]: (U > iniTh, U > minTh) of the lane's 4 columns x 8 rows
	uint32_t* const bm_sel = reinterpret_cast<uint32_t*>(fg_smem + Y.off_sel);    // [row][4]: survivors, bit b = column X0 + b
	uint8_t* const score = fg_smem + Y.off_score;
	uint16_t* const list = reinterpret_cast<uint16_t*>(fg_smem + Y.off_list);
	uint64_t* const bar = reinterpret_cast<uint64_t*>(fg_smem + Y.off_bar);

	const int4 g = __ldg(P.group_tab + blockIdx.x);
	const int X0 = g.x & 0xffff, Y0 = g.x >> 16, rw = g.y & 0xffff, rh = g.y >> 16;
	const int level = g.z & 0xff, ncell = (g.z >> 8) & 0xff, cw = g.z >> 16, c0 = g.w;
	const OrbxLevel& L = P.lv[level];
	const int bx = (X0 - 4) & ~15;                    // u8 boxes start 16-byte aligned; lane l owns columns [X0 + 4l, X0 + 4l + 4)
	if (lane == 0)
	{
		mbar_init(bar, 1);
		mbar_expect_tx(bar, (unsigned)(ST_BW * maps.box_h));
		tma_load_3d(tile, &maps.level[level], bx, Y0 - ST_HALO, P.frame0 + f, bar);
	}
	const uint32_t inv = c_inv20[cw];
	FastRowOps K;
	K.one = (uint32_t)one; K.mone = 0u - K.one; K.c256 = 256u * K.one;
	K.kini = (uint32_t)(0x7fff - P.ini_th) * 0x00010001u; K.kdelta = (uint32_t)(P.ini_th - P.min_th) * 0x00010001u;

	//@phase zero scores and survivor bitmap, wait for the tile
	{
		const int n16 = ((rh + 2) * FG_SS + 15) >> 4;          // score rows -1 .. rh
#pragma unroll 1
		for (int i = lane; i < n16; i += 32) reinterpret_cast<uint4*>(score)[i] = make_uint4(0, 0, 0, 0);
		reinterpret_cast<uint4*>(bm_sel)[lane] = make_uint4(0, 0, 0, 0);
		if (lane + 32 < FG_TH + 2) reinterpret_cast<uint4*>(bm_sel)[lane + 32] = make_uint4(0, 0, 0, 0);
	}
	__syncwarp();
	mbar_wait(bar, 0);

	{
		const int off = X0 - 4 - bx;                  // byte offset of column X0 - 4 inside a box row
		fast_walk(reinterpret_cast<const uint32_t*>(tile) + (off >> 2) + lane, (off & 3) * 8, colbits + lane, K, rh);
	}

	//@phase bound bitmaps of the region (loads, masks)
	const bool tall = rh > 32;                        // warp-uniform
	uint32_t Wi[5], Wm[5];                            // rows 8k .. 8k + 7 of the lane's columns: U > iniTh; minTh < U <= iniTh
	{
		const int cols = rw - 4 * lane;               // the lane's columns inside the group
		const uint32_t cm = (cols >= 4 ? 0xfu : cols <= 0 ? 0u : (1u << cols) - 1u) * 0x11111111u;
#pragma unroll
		for (int k = 0; k < 5; k++)
		{
			const int nr = rh - 8 * k;                // rows of this word that exist
			uint2 w = make_uint2(0, 0);
			if (nr > 0 && (k < 4 || tall)) w = colbits[32 * k + lane];
			const uint32_t rmk = cm & (nr >= 8 ? ~0u : nr <= 0 ? 0u : (1u << (4 * nr)) - 1u);
			Wi[k] = w.x & rmk; Wm[k] = w.y & rmk & ~w.x;
		}
	}

	const int tmin = P.min_th, tini = P.ini_th;
	const uint8_t* __restrict__ t0 = tile + ST_HALO * ST_BW + (X0 - bx);     // pixel (row 0, column X0)
	uint8_t* const s0 = score + FG_SS + 1;                                    // its score (the map has the tile's row stride and a zero border)

	// A list entry: owner lane << 8 | word << 5 | bit, i.e. row = (e >> 2) & 63, column = ((e >> 8) << 2) | (e & 3)
	//@phase exact-score loop around the network (evaluate)
	// exact scores of list[0, n) into the score map; the entries scoring above tk are compacted to the front of the list (a round
	// writes below what it has read). Returns how many were kept.
	auto evaluate = [&](const int n, const int tk) {
		int kept = 0;
#pragma unroll 1
		for (int base = 0; base < n; base += 32)
		{
			const int j = base + lane;
			const bool on = j < n;
			const int e = on ? list[j] : 0;
			const int a = ((e >> 2) & 63) * ST_BW + (((e >> 6) & 0x7c) | (e & 3));
			int s = 0;
			if (on)
			{
				s = max(arc_score_packed_t<ST_BW>(t0 + a), 0);
				s0[a] = (uint8_t)s;
			}
			const unsigned bal = __ballot_sync(0xffffffffu, s > tk);
			if (s > tk) list[kept + __popc(bal & lanemask_lt())] = (uint16_t)e;
			kept += __popc(bal);
		}
		return kept;
	};
	//@phase strict 8-neighbour maxima (select)
	// Each cell is its own cv::FAST call: a pixel in the first (last) column of its cell has no left (right) neighbours.
	// Returns the cells (bit ci) in which this lane found a corner.
	auto select = [&](const int n, const int t) {
		uint32_t fnd = 0;
#pragma unroll 1
		for (int j = lane; j < n; j += 32)
		{
			const int e = list[j], ry = (e >> 2) & 63, b = ((e >> 6) & 0x7c) | (e & 3);
			const uint8_t* sp = s0 + ry * FG_SS + b;
			const int s = sp[0];
			if (s > t)
			{
				const int ci = (int)(((uint32_t)b * inv) >> 20), first = ci * cw;
				int ml = max(max((int)sp[-FG_SS - 1], (int)sp[-1]), (int)sp[FG_SS - 1]);
				int mr = max(max((int)sp[-FG_SS + 1], (int)sp[1]), (int)sp[FG_SS + 1]);
				const int mc = max((int)sp[-FG_SS], (int)sp[FG_SS]);
				if (b == first) ml = 0;
				if (b == first + cw - 1) mr = 0;          // the group's last column has the zero border behind it
				if (s > max(max(ml, mr), mc)) { atomicOr(&bm_sel[4 * ry + (b >> 5)], 1u << (b & 31)); fnd |= 1u << ci; }
			}
		}
		return fnd;
	};

	//@phase iniTh pass, retry decision, minTh pass (control flow)
	// Four passes, one copy of the code: 0 scores of the pixels with U > iniTh (those above iniTh stay listed), 1 strict maxima among
	// them; then, if a cell has no corner (:526-530), 2 scores of its pixels with minTh < U <= iniTh (those above minTh stay listed) and
	// 3 its pixels with U > iniTh join the list, maxima above minTh. A pass whose pixels do not fit one list chunk (noise images) runs
	// chunk by chunk and the following maxima pass lists everything again.
	uint32_t found = 0, rmk = 0;
	int kept = -1;                                    // the previous pass left its survivors in list[0, kept)
#pragma unroll 1
	for (int p = 0; p < 4; p++)
	{
		if (p == 2)
		{
			found = __reduce_or_sync(0xffffffffu, found);
			const uint32_t retry = ~found & ((1u << ncell) - 1u);
			if (!retry) break;
			// the lane's columns that lie in a cell without a corner
#pragma unroll
			for (int j = 0; j < 4; j++)
			{
				const int col = 4 * lane + j;
				const int ci = (int)(((uint32_t)col * inv) >> 20);
				if (col < rw && (retry >> ci & 1)) rmk |= 1u << j;
			}
			rmk *= 0x11111111u;
			kept = -1;
		}
		if (p == 1 && kept >= 0) { found |= select(kept, tini); continue; }
		//@phase candidate list build (expand)
		// pass 3 behind a one-chunk pass 2: only the iniTh pixels are missing from the list
		const bool append = p == 3 && kept >= 0;
		const int at = append ? kept : 0;
		uint32_t W[5];
#pragma unroll
		for (int k = 0; k < 5; k++) W[k] = p < 2 ? Wi[k] : p == 2 ? (Wm[k] & rmk) : append ? (Wi[k] & rmk) : ((Wi[k] | Wm[k]) & rmk);
		int c = __popc(W[0]) + __popc(W[1]) + __popc(W[2]) + __popc(W[3]);
		if (tall) c += __popc(W[4]);
		int o = c;                                    // offsets of the lanes' entries: lane-major
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const int u = __shfl_up_sync(0xffffffffu, o, d);
			if (lane >= d) o += u;
		}
		const int total = __shfl_sync(0xffffffffu, o, 31);
		o -= c;
		if (append && at + total > FG_LIST_CAP)
		{
			// does not fit behind the kept ones: pass 3 again, listing everything (never seen outside noise images)
			p = 2; kept = -1;
			continue;
		}
		const bool single = at + total <= FG_LIST_HALF;
		int done = 0, last = 0;
#pragma unroll 1
		while (done < total || (append && done == 0))
		{
			// a chunk: the lanes whose entries start in [done, done + FG_LIST_HALF)
			const bool in = c > 0 && o >= done && (single || append || o < done + FG_LIST_HALF);
			const int end = (single || append) ? total : __reduce_max_sync(0xffffffffu, in ? o + c : 0);
			if (in)
			{
				uint16_t* q = list + at + o - done;
#pragma unroll
				for (int k = 0; k < 5; k++)
				{
					if (k == 4 && !tall) break;
					uint32_t xr = __brev(W[k]);
					const int tag = (lane << 8) | (k << 5);
					while (xr)
					{
						const int i = __clz(xr);
						*q++ = (uint16_t)(tag + i);
						xr &= ~(0x80000000u >> i);
					}
				}
			}
			__syncwarp();
			const int n = at + end - done;
			if (p & 1) found |= select(n, p == 1 ? tini : tmin);
			else last = evaluate(n, p == 0 ? tini : tmin);
			__syncwarp();
			done = end;
			if (append) break;
		}
		kept = (single && !(p & 1)) ? last : -1;
	}
	__syncwarp();

	//@phase ordered emit
	// per cell, rows ascending, x ascending = cv::FAST's order inside the cell's view. A lane takes survivor row `lane`, or rows 2 lane
	// and 2 lane + 1 of a group taller than 32 rows, so one scan orders them. The counts of the four cells travel through the scan as
	// two packed pairs of 16-bit lanes (a cell holds at most cell_cap < 65536 corners).
	const bool wide = cw > 32;                        // warp-uniform: a cell's row needs a second word
	const int r0 = tall ? 2 * lane : lane;
	uint32_t cnt01 = 0, cnt23 = 0;
	auto cell_row = [&](const int r, const int j, uint32_t& lo, uint32_t& hi) {
		const int o = j * cw, wd = min((j + 1) * cw, rw) - o;               // first bit, width (1..63)
		const uint32_t* q = bm_sel + 4 * r + (o >> 5);                        // bm_sel carries two spare rows: q[2] exists
		const uint32_t q0 = q[0], q1 = q[1];
		lo = __funnelshift_r(q0, q1, o & 31);
		if (wd < 32) lo &= (1u << wd) - 1u;
		hi = 0;
		if (wide) { hi = __funnelshift_r(q1, q[2], o & 31); hi = wd > 32 ? hi & ((1u << (wd - 32)) - 1u) : 0u; }
	};
#pragma unroll
	for (int k = 0; k < 2; k++)
	{
		if (k == 1 && !tall) break;
		const int r = r0 + k;
#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			uint32_t lo = 0, hi = 0;
			if (r < rh && j < ncell) cell_row(r, j, lo, hi);
			const uint32_t cc = (uint32_t)(__popc(lo) + (wide ? __popc(hi) : 0));
			if (j < 2) cnt01 += cc << (16 * j); else cnt23 += cc << (16 * (j - 2));
		}
	}
	uint32_t off01, off23, tot01, tot23;
	{
		uint32_t i01 = cnt01, i23 = cnt23;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const uint32_t u01 = __shfl_up_sync(0xffffffffu, i01, d), u23 = __shfl_up_sync(0xffffffffu, i23, d);
			if (lane >= d) { i01 += u01; i23 += u23; }
		}
		tot01 = __shfl_sync(0xffffffffu, i01, 31); tot23 = __shfl_sync(0xffffffffu, i23, 31);
		off01 = i01 - cnt01; off23 = i23 - cnt23;
	}
	uint32_t* __restrict__ out = P.cand + (int64_t)f * P.cand_per_frame + L.cand_base + (int64_t)c0 * L.cell_cap;
#pragma unroll 1
	for (int j = 0; j < ncell; j++)
	{
		int pos = (int)(((j < 2 ? off01 : off23) >> (16 * (j & 1))) & 0xffffu);
#pragma unroll
		for (int k = 0; k < 2; k++)
		{
			if (k == 1 && !tall) break;
			const int ry = r0 + k;
			if (ry < rh)
			{
				uint32_t lo, hi;
				cell_row(ry, j, lo, hi);
				const uint8_t* srow = s0 + ry * FG_SS + j * cw;
				const uint32_t xy = (uint32_t)(X0 + j * cw) | ((uint32_t)(Y0 + ry) << 12);
				while (lo)
				{
					const int rx = __ffs(lo) - 1;
					lo &= lo - 1;
					out[pos++] = xy + (uint32_t)rx + ((uint32_t)((int)srow[rx] - 1) << 24);
				}
				while (hi)
				{
					const int rx = 32 + __ffs(hi) - 1;
					hi &= hi - 1;
					out[pos++] = xy + (uint32_t)rx + ((uint32_t)((int)srow[rx] - 1) << 24);
				}
			}
		}
		out += L.cell_cap;
	}
	if (lane < ncell)
	{
		const uint32_t t = lane < 2 ? tot01 : tot23;
		P.cell_count[(int64_t)f * P.cells_per_frame + L.cell_base + c0 + lane] = (int)((t >> (16 * (lane & 1))) & 0xffffu);
	}
}
//@end
