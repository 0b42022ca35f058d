// Quadtree selection kernel body, included by orbx_extract.cu once per CTA size (the includer defines QT_THREADS and QT_NS):
// 256 threads for small batches and 4K-class plans, where one CTA's latency is what the launch waits for and the CTA-parallel sort /
// block-wide partition use every thread; 128 threads for large batches, where more co-resident CTAs fill the SM while others sit in
// their serial phases (measured per 512 C1 frames: 256 thr 0.355 ms, 128 thr 0.301 ms, 64 thr 0.390 ms); 512 threads for batches of
// 4K-class plans, whose node lists take > 100 KB of shared memory (one CTA per SM either way).
#define QT_WARPS (QT_THREADS / 32)
namespace QT_NS {


struct QNode
{
	uint16_t x0, y0, x1, y1;
	uint32_t beg;
	uint32_t cnt;     // bit 31: which ping-pong buffer holds the segment
};
#undef QN_CNT
#undef QN_BUF
#define QN_CNT(n) ((n).cnt & 0x7fffffffu)
#define QN_BUF(n) ((n).cnt >> 31)

template <class Pred, class Emit>
__device__ __forceinline__ int block_ordered(int n, int* s_w, Pred pred, Emit emit)
{
	// calls emit(i, rank) for every i in [0,n) with pred(i), rank = number of earlier true elements. Uniform result.
	const int tid = threadIdx.x, warp = tid >> 5;
	int base = 0;
	for (int c0 = 0; c0 < n; c0 += QT_THREADS)
	{
		const int i = c0 + tid;
		const bool p = (i < n) && pred(i);
		const unsigned bal = __ballot_sync(0xffffffffu, p);
		if ((tid & 31) == 0) s_w[warp] = __popc(bal);
		__syncthreads();
		int wbase = 0, tot = 0;
#pragma unroll
		for (int w = 0; w < QT_WARPS; w++)
		{
			const int v = s_w[w];
			if (w < warp) wbase += v;
			tot += v;
		}
		if (p) emit(i, base + wbase + __popc(bal & lanemask_lt()));
		base += tot;
		__syncthreads();
	}
	return base;
}

template <class Get, class Put>
__device__ __forceinline__ int block_exscan(int n, int* s_w, Get get, Put put)
{
	// put(i, exclusive prefix of get) for i in [0,n); returns the total. Uniform result.
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	int base = 0;
	for (int c0 = 0; c0 < n; c0 += QT_THREADS)
	{
		const int i = c0 + tid;
		const int v = (i < n) ? get(i) : 0;
		int inc = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const int t = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += t;
		}
		if (lane == 31) s_w[warp] = inc;
		__syncthreads();
		int wbase = 0, tot = 0;
#pragma unroll
		for (int w = 0; w < QT_WARPS; w++)
		{
			const int t = s_w[w];
			if (w < warp) wbase += t;
			tot += t;
		}
		if (i < n) put(i, base + wbase + inc - v);
		base += tot;
		__syncthreads();
	}
	return base;
}

// std::sort replayed by the whole CTA with the SAME result as the serial algorithm. Two facts make that possible:
//  * after a partition step the two sub-ranges are never touched together again, so __introsort_loop's "recurse right,
//    iterate left" can run both sides at the same time: every round, each pending range (> 16 elements) is partitioned by
//    one thread; a range keeps the depth budget it would have had (both sides inherit depth - 1; budget 0 = heapsort);
//  * the partition leaves every element of a left range "not after" every element of the range to its right (sizes >= pivot
//    on the left, <= pivot on the right), so __final_insertion_sort never moves an element across a range boundary: it is
//    a stable insertion sort of every final range (<= 16 elements) on its own.
// Time is ~2n serial steps (n + n/2 + n/4 ...) instead of ~1.4 n log2 n.
struct QSeg { int first, last, depth; };

__device__ __noinline__ void qs_sort_block(uint64_t* a, int n, QSeg* segq, int segcap, ushort2* leaf, int* s_cnt /* [3]: nseg[2], nleaf */, uint16_t* rs /* [n] scratch */)
{
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	if (tid == 0)
	{
		int lg = 0;
		for (int m = n; m > 1; m >>= 1) ++lg;
		s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0;
		if (n > 16) { segq[0] = { 0, n, 2 * lg }; s_cnt[0] = 1; }
		else if (n > 1) { leaf[0] = make_ushort2(0, (unsigned short)n); s_cnt[2] = 1; }
	}
	__syncthreads();
	int cur = 0;
	for (;;)
	{
		const int ns = s_cnt[cur];
		if (ns == 0) break;
		__syncthreads();
		if (tid == 0) s_cnt[cur ^ 1] = 0;
		__syncthreads();
		QSeg* in = segq + cur * segcap;
		QSeg* out = segq + (cur ^ 1) * segcap;
		// one WARP per pending range: the partition step itself is parallel (qs_partition_warp), so the first rounds, where one or two long
		// ranges are all there is, no longer run at the pace of a single thread
		for (int si = warp; si < ns; si += QT_WARPS)
		{
			const QSeg sg = in[si];
			if (sg.depth == 0) { if (lane == 0) qs_heapsort(a, sg.first, sg.last); __syncwarp(); continue; }     // fully sorted, no insertion pass needed
			const int cut = qs_partition_warp(a, sg.first, sg.last, rs);
			if (lane == 0)
			{
#pragma unroll
				for (int side = 0; side < 2; side++)
				{
					const int f0 = side ? cut : sg.first, l0 = side ? sg.last : cut;
					if (l0 - f0 > 16) out[atomicAdd(&s_cnt[cur ^ 1], 1)] = { f0, l0, sg.depth - 1 };
					else if (l0 - f0 > 1) leaf[atomicAdd(&s_cnt[2], 1)] = make_ushort2((unsigned short)f0, (unsigned short)l0);
				}
			}
		}
		__syncthreads();
		cur ^= 1;
	}
	const int nleaf = s_cnt[2];
	for (int li = tid; li < nleaf; li += QT_THREADS) qs_insertion(a, leaf[li].x, leaf[li].y);
	__syncthreads();
}

__device__ __forceinline__ int quadrant_of(uint32_t v, int xm, int ym)
{
	const int x = orbx_px(v), y = orbx_py(v);
	return x < xm ? (y < ym ? 0 : 2) : (y < ym ? 1 : 3);
}

// Stable 4-way partition of one big node by the whole CTA (see the call site).
__device__ __noinline__ void qt_divide_big(const QNode nd, uint32_t* buf0, uint32_t* buf1, uint32_t* cc, int (*s_bigc)[4])
{
	// Stable 4-way partition of one big node by the whole CTA. Every warp owns a CONTIGUOUS run of the node's elements (a multiple of 32
	// long), so after one exchange of the per-warp quadrant counts each warp knows where its elements of every quadrant start and scatters
	// them with ballots alone: no barrier and no shared-memory traffic per chunk (the chunk-synchronous version spent 1.6 us per 512
	// elements on them: 242 us per pass over the 78 k candidates of a 4K level 0).
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int cnt = (int)QN_CNT(nd);
	const uint32_t* src = (QN_BUF(nd) ? buf1 : buf0) + nd.beg;
	uint32_t* dst = (QN_BUF(nd) ? buf0 : buf1) + nd.beg;
	const int xm = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1), ym = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
	const int seg = ((cnt + QT_WARPS - 1) / QT_WARPS + 31) & ~31;
	const int s0 = min(warp * seg, cnt), s1 = min(s0 + seg, cnt);
	int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 4
	for (int i = s0 + lane; i < s1; i += 32)
	{
		const int q = quadrant_of(src[i], xm, ym);
		c0 += q == 0; c1 += q == 1; c2 += q == 2; c3 += q == 3;
	}
	// one REDUX per counter instead of five shuffle + add steps (the divide of a node is a dependent chain: latency counts here)
	c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1);
	c2 = __reduce_add_sync(0xffffffffu, c2); c3 = __reduce_add_sync(0xffffffffu, c3);
	if (lane == 0) { s_bigc[warp][0] = c0; s_bigc[warp][1] = c1; s_bigc[warp][2] = c2; s_bigc[warp][3] = c3; }
	__syncthreads();
	int tot[4], run[4];
#pragma unroll
	for (int q = 0; q < 4; q++)
	{
		tot[q] = 0; run[q] = 0;
#pragma unroll
		for (int w = 0; w < QT_WARPS; w++)
		{
			const int x = s_bigc[w][q];
			tot[q] += x;
			if (w < warp) run[q] += x;                   // elements of quadrant q held by the warps before this one
		}
	}
	run[1] += tot[0]; run[2] += tot[0] + tot[1]; run[3] += tot[0] + tot[1] + tot[2];
	for (int i0 = s0; i0 < s1; i0 += 128)              // four chunks' loads in flight before the ballots
	{
		uint32_t vv[4];
#pragma unroll
		for (int u = 0; u < 4; u++)
		{
			const int i = i0 + 32 * u + lane;
			vv[u] = i < s1 ? src[i] : 0u;
		}
#pragma unroll
		for (int u = 0; u < 4; u++)
		{
			if (i0 + 32 * u >= s1) break;                // uniform
			const int i = i0 + 32 * u + lane;
			const bool ok = i < s1;
			const uint32_t v = vv[u];
			const int q = ok ? quadrant_of(v, xm, ym) : -1;
			unsigned bq[4];
#pragma unroll
			for (int k = 0; k < 4; k++) bq[k] = __ballot_sync(0xffffffffu, q == k);
			int mine = 0;
#pragma unroll
			for (int k = 0; k < 4; k++)
			{
				if (k == q) mine = run[k] + __popc(bq[k] & lanemask_lt());
				run[k] += __popc(bq[k]);
			}
			if (ok) dst[mine] = v;
		}
	}
	if (tid == 0) { cc[0] = tot[0]; cc[1] = tot[1]; cc[2] = tot[2]; cc[3] = tot[3]; }
	__syncthreads();
}

// BIG selects the variant with the CTA-parallel sort and big-node partition (4K-class levels); the plain variant keeps the
// register count at 40 for VGA-class levels, where occupancy matters more than the serial tails.
template <bool BIG>
__global__ void __launch_bounds__(QT_THREADS) k_quadtree(const OrbxPlanDev P, int* __restrict__ cell_off, const int big_node_min, const int par_sort_min, unsigned long long* dbg, const int lvl0, const int smem_cand)
{
	int dbgk = 0;
#define QT_STAMP() do { if (dbg && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && dbgk < 63) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[1 + dbgk++] = t_; dbg[0] = dbgk; } } while (0)
	QT_STAMP();
	extern __shared__ __align__(16) uint8_t qsm[];
	const int M = P.node_cap;
	QNode* listA = reinterpret_cast<QNode*>(qsm);
	QNode* listB = listA + M;
	uint32_t* childcnt = reinterpret_cast<uint32_t*>(listB + M);       // [M][4]
	uint64_t* items = reinterpret_cast<uint64_t*>(childcnt);           // phase-2 sort items: dead (copied to proc) before the divide step writes childcnt
	uint32_t* proc = childcnt + 4 * M;                                 // positions (old list) to divide, processing order
	uint32_t* pbase = proc + M;                                        // exclusive scan of non-empty child counts
	// the CTA-parallel sort's work lists exist only in the BIG variant: the plain one runs more CTAs per SM on the shared memory they would take
	ushort2* leaf = reinterpret_cast<ushort2*>(pbase + M);             // ranges (2 .. 16 items, so at most M / 2 of them; M < 65536) left for the insertion pass of the sort
	const int segcap = BIG ? M / 16 + 4 : 0;
	QSeg* segq = reinterpret_cast<QSeg*>(leaf + (BIG ? (M / 2 + 2) & ~1 : 0));   // [2][segcap] ranges still to be partitioned
	uint8_t* gone = reinterpret_cast<uint8_t*>(segq + 2 * segcap);     // old-list positions removed by this pass
	__shared__ int s_w[QT_WARPS];
	__shared__ int s_K;
	__shared__ int s_sort[3];
	__shared__ int s_rootcnt[ORBX_MAX_ROOTS];
	__shared__ int s_bigc[QT_WARPS][4];
	__shared__ int s_rc[QT_WARPS][ORBX_MAX_ROOTS];

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int lvl = (int)blockIdx.y + lvl0, f = blockIdx.x;      // lvl0: first level of the launch's level range      // x = frame: all level-0 CTAs (the longest) are scheduled first
	const OrbxLevel& L = P.lv[lvl];
	const int ncell = L.ncx * L.ncy;
	const int* __restrict__ ccount = P.cell_count + (int64_t)f * P.cells_per_frame + L.cell_base;
	int* __restrict__ coff = cell_off + (int64_t)f * P.cells_per_frame + L.cell_base;
	const uint32_t* __restrict__ slots = P.cand + (int64_t)f * P.cand_per_frame + L.cand_base;
	uint32_t* buf0 = P.qbuf0 + (int64_t)f * P.cand_per_frame + L.cand_base;
	uint32_t* buf1 = P.qbuf1 + (int64_t)f * P.cand_per_frame + L.cand_base;

	// ---- K3: cell-major compaction of the candidate slots (DetectFAST's push_back order, :532-537)
	const int n = block_exscan(ncell, s_w, [&](int i) { return ccount[i]; }, [&](int i, int off) { coff[i] = off; });
	if (tid == 0) P.cand_count[(int64_t)f * P.nlevels + lvl] = n;
	__syncthreads();
	// A frame at a time, the launch waits for its slowest CTA, and every pass of a CTA is a chain of dependent reads and writes of the
	// candidate segments. When the level's candidates fit the space the launcher reserved behind the node lists, both ping-pong buffers
	// live in shared memory instead of global memory (L2): n is uniform, so is the choice.
	if (n <= smem_cand)
	{
		buf0 = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(gone + M) + 15) & ~(uintptr_t)15);
		buf1 = buf0 + smem_cand;
	}
	const int nroots = L.n_roots;
	uint32_t* gathered = (nroots == 1) ? buf0 : buf1;
	// 8 lanes per cell (a cell holds ~15 candidates): four cells per warp in flight, so the chain of dependent loads
	// (count, offset, slots) is walked ncell / 32 times per warp instead of ncell / 8 times
	// (two cells per group and trip, counts / offsets and then the first 16 candidates of both loaded before anything is stored: the
	// dependent chain count -> slots is walked once for both)
	for (int c0 = tid >> 3; c0 < ncell; c0 += QT_THREADS / 4)
	{
		int cnt[2], off[2];
		const uint32_t* srcp[2];
#pragma unroll
		for (int u = 0; u < 2; u++)
		{
			const int cidx = c0 + u * (QT_THREADS / 8);
			const bool live = cidx < ncell;
			cnt[u] = live ? ccount[cidx] : 0;
			off[u] = live ? coff[cidx] : 0;
			srcp[u] = slots + (int64_t)(live ? cidx : 0) * L.cell_cap;
		}
		uint32_t v[2][2];
#pragma unroll
		for (int u = 0; u < 2; u++)
#pragma unroll
			for (int j = 0; j < 2; j++)
			{
				const int k = (tid & 7) + 8 * j;
				v[u][j] = k < cnt[u] ? srcp[u][k] : 0u;
			}
#pragma unroll
		for (int u = 0; u < 2; u++)
		{
#pragma unroll
			for (int j = 0; j < 2; j++)
			{
				const int k = (tid & 7) + 8 * j;
				if (k < cnt[u]) gathered[off[u] + k] = v[u][j];
			}
			for (int k = (tid & 7) + 16; k < cnt[u]; k += 8) gathered[off[u] + k] = srcp[u][k];
		}
	}
	__syncthreads();
	QT_STAMP();
	if (n == 0)
	{
		if (tid == 0) P.sel_count[(int64_t)f * P.nlevels + lvl] = 0;    // early return of :544-545 (dst == src stays empty)
		return;
	}

	// ---- roots (:547-581): vertical strips; candidate -> strip by the host-built table (double arithmetic there)
	const int* __restrict__ rootx = P.root_x + L.root_base;
	const uint8_t* __restrict__ rlut = P.root_lut + L.rootlut_base;
	if (nroots == 1)
	{
		if (tid == 0) s_rootcnt[0] = n;
	}
	else
	{
		// stable partition of the n candidates into the root strips, same scheme as qt_divide_big: a contiguous run per warp, one exchange
		// of per-warp counts, then ballots only. Lane r of a warp keeps the count / write cursor of strip r (nroots <= 16).
		const int seg = ((n + QT_WARPS - 1) / QT_WARPS + 31) & ~31;
		const int s0 = min(warp * seg, n), s1 = min(s0 + seg, n);
		int c = 0;
		for (int i0 = s0; i0 < s1; i0 += 32)
		{
			const int i = i0 + lane;
			const int key = i < s1 ? (int)rlut[orbx_px(buf1[i])] : -1;
			for (int r = 0; r < nroots; r++)
			{
				const unsigned b = __ballot_sync(0xffffffffu, key == r);
				if (lane == r) c += __popc(b);
			}
		}
		if (lane < ORBX_MAX_ROOTS) s_rc[warp][lane] = lane < nroots ? c : 0;
		__syncthreads();
		int run = 0;                                   // lane r: where this warp's elements of strip r start in buf0
		if (lane < nroots)
		{
			for (int r = 0; r < lane; r++)
				for (int w = 0; w < QT_WARPS; w++) run += s_rc[w][r];
			int tot = 0;
			for (int w = 0; w < QT_WARPS; w++)
			{
				const int x = s_rc[w][lane];
				if (w < warp) run += x;
				tot += x;
			}
			if (warp == 0) s_rootcnt[lane] = tot;
		}
		for (int i0 = s0; i0 < s1; i0 += 128)          // four chunks (candidate, then its strip) loaded before the ballots
		{
			uint32_t vv[4];
			int kk[4];
#pragma unroll
			for (int u = 0; u < 4; u++)
			{
				const int i = i0 + 32 * u + lane;
				vv[u] = i < s1 ? buf1[i] : 0u;
			}
#pragma unroll
			for (int u = 0; u < 4; u++)
			{
				const int i = i0 + 32 * u + lane;
				kk[u] = i < s1 ? (int)rlut[orbx_px(vv[u])] : -1;
			}
#pragma unroll
			for (int u = 0; u < 4; u++)
			{
				if (i0 + 32 * u >= s1) break;            // uniform
				const int key = kk[u];
				int pos = 0;
				for (int r = 0; r < nroots; r++)
				{
					const unsigned b = __ballot_sync(0xffffffffu, key == r);
					const int base = __shfl_sync(0xffffffffu, run, r);
					if (key == r) pos = base + __popc(b & lanemask_lt());
					if (lane == r) run += __popc(b);
				}
				if (key >= 0) buf0[pos] = vv[u];
			}
		}
	}
	__syncthreads();
	int listLen = 0;
	{
		int at = 0;
		for (int r = 0; r < nroots; r++)
		{
			const int c = s_rootcnt[r];
			if (c > 0)
			{
				if (tid == 0)
				{
					QNode nd;
					nd.x0 = (uint16_t)rootx[r]; nd.y0 = (uint16_t)L.miny; nd.x1 = (uint16_t)rootx[r + 1]; nd.y1 = (uint16_t)L.maxy;
					nd.beg = (uint32_t)at; nd.cnt = (uint32_t)c;     // buffer 0
					listA[listLen] = nd;
				}
				listLen++;
			}
			at += c;
		}
	}
	__syncthreads();

	QT_STAMP();
	QNode* cur = listA;
	QNode* nxt = listB;
	const int quota = L.quota;
	int phase = 1, lastP = 0;
	for (;;)
	{
		// ---- which nodes does this pass divide, and in which order
		int np;
		if (phase == 1)
		{
			// every divisible node, list order (:588-631)
			np = block_ordered(listLen, s_w, [&](int i) { return QN_CNT(cur[i]) > 1; }, [&](int i, int rank) { proc[rank] = i; });
		}
		else
		{
			// children of the previous pass with > 1 point, in push order (= back to front of the first lastP list
			// entries), sorted by size descending with libstdc++'s tie order (:635-643)
			np = block_ordered(lastP, s_w, [&](int g) { return QN_CNT(cur[lastP - 1 - g]) > 1; },
			                   [&](int g, int rank) { const int pos = lastP - 1 - g; items[rank] = ((uint64_t)QN_CNT(cur[pos]) << 32) | (uint32_t)pos; });
			__syncthreads();
			// proc and pbase (8 M bytes, contiguous) are written only after the sort: the plain variant, which reserves no work lists for the
			// CTA-parallel sort, borrows them: stopper scratch (2 M bytes), leaf ranges (2 M + 8), the two range queues (1.5 M + 96)
			const int segcap_p = M / 16 + 4;
			const size_t off_leaf = ((size_t)2 * M + 3) & ~(size_t)3, off_seg = (off_leaf + 4 * ((size_t)M / 2 + 2) + 15) & ~(size_t)15;
			const bool fits = (size_t)8 * M >= off_seg + (size_t)2 * segcap_p * sizeof(QSeg);
			if (BIG && np > par_sort_min)
				qs_sort_block(items, np, segq, segcap, leaf, s_sort, reinterpret_cast<uint16_t*>(proc));
			else if (!BIG && fits && np > par_sort_min)
			{
				uint8_t* sc = reinterpret_cast<uint8_t*>(proc);
				qs_sort_block(items, np, reinterpret_cast<QSeg*>(sc + off_seg), segcap_p, reinterpret_cast<ushort2*>(sc + off_leaf), s_sort,
				              reinterpret_cast<uint16_t*>(sc));
			}
			else
			{
				if (tid == 0) qs_sort_serial(items, np);
				__syncthreads();
			}
			for (int i = tid; i < np; i += QT_THREADS) proc[i] = (uint32_t)(items[i] & 0xffffffffu);
		}
		for (int i = tid; i < listLen; i += QT_THREADS) gone[i] = 0;
		__syncthreads();
		QT_STAMP();

		// ---- big nodes (the first passes of a large level: one node can hold tens of thousands of candidates) are divided by
		//      the whole CTA: every warp takes a contiguous run of the node, position = quadrant base + the quadrant's elements in earlier
		//      warps + in earlier chunks of this warp + in earlier lanes, i.e. the same stable 4-way partition as the warp version below.
		if (BIG && np <= 64)
			for (int t = 0; t < np; t++)
			{
				const QNode nd = cur[proc[t]];
				if ((int)QN_CNT(nd) >= big_node_min)            // uniform: every thread sees the same node
					qt_divide_big(nd, buf0, buf1, childcnt + 4 * t, s_bigc);
			}

		// ---- divide (speculatively all of them; Phase 2 may stop early, parents stay intact in their buffer)
		for (int t = warp; t < np; t += QT_WARPS)
		{
			const QNode nd = cur[proc[t]];
			const int cnt = (int)QN_CNT(nd);
			if (BIG && np <= 64 && cnt >= big_node_min) continue;  // done by the whole CTA above
			const uint32_t* src = (QN_BUF(nd) ? buf1 : buf0) + nd.beg;
			uint32_t* dst = (QN_BUF(nd) ? buf0 : buf1) + nd.beg;
			const int xm = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1), ym = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);   // ceil(0.5*d), :408-409
			if (cnt <= 32)
			{
				// most nodes of the later passes: one chunk. One load serves the counts (ballots) and the scatter, instead of a counting
				// trip and a scattering trip through the candidates (two dependent L2 round trips where they live in global memory)
				const bool ok = lane < cnt;
				const uint32_t v = ok ? src[lane] : 0u;
				const int q = ok ? quadrant_of(v, xm, ym) : -1;
				const unsigned b0 = __ballot_sync(0xffffffffu, q == 0), b1 = __ballot_sync(0xffffffffu, q == 1);
				const unsigned b2 = __ballot_sync(0xffffffffu, q == 2), b3 = __ballot_sync(0xffffffffu, q == 3);
				const unsigned lt = lanemask_lt();
				const int n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2);
				if (q == 0) dst[__popc(b0 & lt)] = v;
				else if (q == 1) dst[n0 + __popc(b1 & lt)] = v;
				else if (q == 2) dst[n0 + n1 + __popc(b2 & lt)] = v;
				else if (q == 3) dst[n0 + n1 + n2 + __popc(b3 & lt)] = v;
				if (lane == 0)
				{
					childcnt[4 * t + 0] = n0; childcnt[4 * t + 1] = n1; childcnt[4 * t + 2] = n2; childcnt[4 * t + 3] = __popc(b3);
				}
				continue;
			}
			// The candidates sit in global memory (L2): one chunk per trip costs a full L2 round trip. The count loop keeps per-lane
			// counters (no ballot between its loads, so they overlap); the scatter loop loads four chunks before it touches them.
			int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 4
			for (int i = lane; i < cnt; i += 32)
			{
				const int q = quadrant_of(src[i], xm, ym);
				c0 += q == 0; c1 += q == 1; c2 += q == 2; c3 += q == 3;
			}
			// one REDUX per counter instead of five shuffle + add steps (the divide of a node is a dependent chain: latency counts here)
			c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1);
			c2 = __reduce_add_sync(0xffffffffu, c2); c3 = __reduce_add_sync(0xffffffffu, c3);
			int a0 = 0, a1 = c0, a2 = c0 + c1, a3 = c0 + c1 + c2;
			for (int i0 = 0; i0 < cnt; i0 += 128)
			{
				uint32_t vv[4];
#pragma unroll
				for (int u = 0; u < 4; u++)
				{
					const int i = i0 + 32 * u + lane;
					vv[u] = i < cnt ? src[i] : 0u;
				}
#pragma unroll
				for (int u = 0; u < 4; u++)
				{
					if (i0 + 32 * u >= cnt) break;               // uniform
					const int i = i0 + 32 * u + lane;
					const bool ok = i < cnt;
					const uint32_t v = vv[u];
					const int q = ok ? quadrant_of(v, xm, ym) : -1;
					const unsigned b0 = __ballot_sync(0xffffffffu, q == 0), b1 = __ballot_sync(0xffffffffu, q == 1);
					const unsigned b2 = __ballot_sync(0xffffffffu, q == 2), b3 = __ballot_sync(0xffffffffu, q == 3);
					const unsigned lt = lanemask_lt();
					if (q == 0) dst[a0 + __popc(b0 & lt)] = v;
					else if (q == 1) dst[a1 + __popc(b1 & lt)] = v;
					else if (q == 2) dst[a2 + __popc(b2 & lt)] = v;
					else if (q == 3) dst[a3 + __popc(b3 & lt)] = v;
					a0 += __popc(b0); a1 += __popc(b1); a2 += __popc(b2); a3 += __popc(b3);
				}
			}
			if (lane == 0)
			{
				childcnt[4 * t + 0] = c0; childcnt[4 * t + 1] = c1; childcnt[4 * t + 2] = c2; childcnt[4 * t + 3] = c3;
			}
		}
		__syncthreads();

		QT_STAMP();
		// ---- how many of them are really processed: Phase 2 breaks once the list reaches the quota (:666-667)
		auto nkids = [&](int t) { return (int)(childcnt[4 * t] > 0) + (int)(childcnt[4 * t + 1] > 0) + (int)(childcnt[4 * t + 2] > 0) + (int)(childcnt[4 * t + 3] > 0); };
		const int totalPush = block_exscan(np, s_w, nkids, [&](int t, int ex) { pbase[t] = ex; });
		int K = np, Ppush = totalPush;
		if (phase == 2)
		{
			if (tid == 0) s_K = np;
			__syncthreads();
			// list length after processing t+1 items = listLen + pbase[t] + nkids(t) - (t+1); non-decreasing in t
			for (int t = tid; t < np; t += QT_THREADS)
			{
				const int after = listLen + (int)pbase[t] + nkids(t) - (t + 1);
				const int before = listLen + (int)pbase[t] - t;
				if (after >= quota && before < quota) s_K = t + 1;
			}
			__syncthreads();
			K = s_K;
			Ppush = (K < np) ? (int)pbase[K] : totalPush;
			__syncthreads();
		}

		QT_STAMP();
		// ---- rebuild the list: children in reverse push order, then the surviving old nodes in old order
		for (int t = tid; t < K; t += QT_THREADS)
		{
			const int pos = proc[t];
			gone[pos] = 1;
			const QNode nd = cur[pos];
			const int xm = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1), ym = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
			const uint32_t nb = (QN_BUF(nd) ^ 1u) << 31;
			uint32_t at = nd.beg;
			int g = (int)pbase[t];
#pragma unroll
			for (int q = 0; q < 4; q++)
			{
				const uint32_t c = childcnt[4 * t + q];
				if (c > 0)
				{
					QNode ch;
					ch.x0 = (q & 1) ? (uint16_t)xm : nd.x0; ch.x1 = (q & 1) ? nd.x1 : (uint16_t)xm;
					ch.y0 = (q & 2) ? (uint16_t)ym : nd.y0; ch.y1 = (q & 2) ? nd.y1 : (uint16_t)ym;
					ch.beg = at; ch.cnt = c | nb;
					nxt[Ppush - 1 - g] = ch;
					g++;
				}
				at += c;
			}
		}
		__syncthreads();
		const int kept = block_ordered(listLen, s_w, [&](int i) { return gone[i] == 0; }, [&](int i, int rank) { nxt[Ppush + rank] = cur[i]; });
		const int newLen = Ppush + kept;
		{ QNode* t = cur; cur = nxt; nxt = t; }
		const int prevLen = listLen;
		listLen = newLen;
		lastP = Ppush;
		__syncthreads();
		if (listLen >= quota || listLen == prevLen)
			break;
		if (phase == 1)
		{
			// toExpand = children of this pass with > 1 point (:617-622); switch to largest-first near the quota (:633-634)
			const int ndiv = block_ordered(lastP, s_w, [&](int i) { return QN_CNT(cur[i]) > 1; }, [&](int, int) {});
			if (listLen + 3 * ndiv > quota) phase = 2;
		}
	}

	QT_STAMP();
	// ---- keep the best response of every node, first wins ties, list order (:677-692)
	uint32_t* __restrict__ sel = P.sel + (int64_t)f * P.sel_per_frame + L.sel_base;
	// 8 lanes per node (final nodes hold ~10 candidates): four nodes per warp in flight. The loop bound is rounded up so that every
	// lane of a warp takes part in the shuffles.
	// Two nodes per 8-lane group and trip, their candidates loaded before either reduction starts, and the winner's packed value
	// travels through the shuffles with its response: one L2 round trip per trip instead of four.
	for (int i0 = 0; i0 < listLen; i0 += QT_THREADS / 4)
	{
		int bestr[2], besti[2], cntu[2];
		uint32_t bestv[2];
#pragma unroll
		for (int u = 0; u < 2; u++)
		{
			const int i = i0 + u * (QT_THREADS / 8) + (tid >> 3);
			const bool live = i < listLen;
			const QNode nd = cur[live ? i : 0];
			cntu[u] = live ? (int)QN_CNT(nd) : 0;
			const uint32_t* src = (QN_BUF(nd) ? buf1 : buf0) + nd.beg;
			bestr[u] = 0; besti[u] = 0x7fffffff; bestv[u] = 0u;
			for (int k = tid & 7; k < cntu[u]; k += 8)
			{
				const uint32_t v = src[k];
				const int r = orbx_pr(v);
				if (r > bestr[u]) { bestr[u] = r; besti[u] = k; bestv[u] = v; }
			}
		}
#pragma unroll
		for (int u = 0; u < 2; u++)
		{
#pragma unroll
			for (int d = 4; d > 0; d >>= 1)
			{
				const int orr = __shfl_xor_sync(0xffffffffu, bestr[u], d), oi = __shfl_xor_sync(0xffffffffu, besti[u], d);
				const uint32_t ov = __shfl_xor_sync(0xffffffffu, bestv[u], d);
				if (orr > bestr[u] || (orr == bestr[u] && oi < besti[u])) { bestr[u] = orr; besti[u] = oi; bestv[u] = ov; }
			}
			const int i = i0 + u * (QT_THREADS / 8) + (tid >> 3);
			if (i < listLen && (tid & 7) == 0) sel[i] = bestv[u];
		}
	}
	if (tid == 0) P.sel_count[(int64_t)f * P.nlevels + lvl] = listLen;
	QT_STAMP();
#undef QT_STAMP
}

}  // namespace QT_NS
#undef QT_WARPS
