// sm_100a kernels of the Hamming matchers: DescriptorDistance, brute-force best/second scan (knn2) with the
// train-sharded merge, and ComputeStereoMatches. Reference: src/ORBmatcher.cc:41-52, :60-247, :477-507, :1449-1457.
// Integer / popcount work: tensor cores are deliberately not used (BASELINE.json north_star).
#include "orbx_internal.cuh"

#include <math.h>

namespace {

__device__ __forceinline__ int hamming8(const uint32_t* a, const uint32_t* b)
{
	int d = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) d += __popc(a[i] ^ b[i]);
	return d;
}

// ---------------------------------------------------------------------------------------------------
// M1  DescriptorDistance for n independent pairs (src/ORBmatcher.cc:1449-1457)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_hamming_pairs(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int64_t n,
                                                       int32_t* __restrict__ out)
{
	const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	const uint4* pa = reinterpret_cast<const uint4*>(a + 32 * i);
	const uint4* pb = reinterpret_cast<const uint4*>(b + 32 * i);
	const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 1), b0 = __ldg(pb), b1 = __ldg(pb + 1);
	out[i] = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
	         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---------------------------------------------------------------------------------------------------
// M2  hamming_knn2 — best / second-best scan (SearchByBoW inner loop, src/ORBmatcher.cc:477-507).
//     Queries live in registers (KQ per thread), train rows stream through a double-buffered shared tile
//     and are read as warp-wide broadcasts. Per pair: 8 LOP3(xor) + 6 LOP3(carry-save) + 5 POPC + ~4 add/shift + 1 key + 3 min/max.
//     key = dist << 22 | row-in-chunk is unique inside a chunk (<= 4 Mi rows), so
//         k2 = min(k2, max(k1, key)); k1 = min(k1, key)
//     tracks the two smallest keys: k1 = (best, lowest index), k2 >> 22 = second-best distance. A row at
//     distance 256 yields key >= NONE and is never taken, exactly like the reference's `dist < 256` start.
// ---------------------------------------------------------------------------------------------------
#define KN_THREADS 128
#define KN_KQ 4                       // queries per thread
#define KN_QB (KN_THREADS * KN_KQ)    // queries per block
#define KN_TILE 128                   // train rows per shared tile
#define KN_CHUNK (1 << 22)
#define KN_NONE 0x40000000u           // 256 << 22

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// 256-bit Hamming distance with 5 POPC instead of 8. Measured on this B200 (tools/pipe_probe.cu): POPC 15, LOP3 59
// lane-ops/clk/SM, so the plain 8 x (XOR, POPC) form is POPC-bound at 0.53 clk/pair. Three carry-save adders (2 LOP3 each:
// sum = a^b^c, carry = majority) fold seven of the eight XOR words into one "ones" word and three "twos" words:
//   d = popc(ones) + popc(x7) + 2 * (popc(c1) + popc(c2) + popc(c3)).
// That is 14 LOP3 + 5 POPC per pair, which balances the two pipes (a fourth adder would save one more POPC but makes the
// LOP3 pipe the bottleneck: measured 711 vs 529 Gpairs/s for the plain form).
__device__ __forceinline__ int hamming256_csa(const uint32_t* q, const uint4 lo, const uint4 hi)
{
	const uint32_t x0 = q[0] ^ lo.x, x1 = q[1] ^ lo.y, x2 = q[2] ^ lo.z, x3 = q[3] ^ lo.w;
	const uint32_t x4 = q[4] ^ hi.x, x5 = q[5] ^ hi.y, x6 = q[6] ^ hi.z, x7 = q[7] ^ hi.w;
	const uint32_t s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 ^ x1));
	const uint32_t s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 ^ x4));
	const uint32_t s3 = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 ^ s2));
	return __popc(s3) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

__device__ __forceinline__ uint64_t knn_pack(uint32_t best, uint32_t second, uint32_t idx)
{
	return ((uint64_t)best << 48) | ((uint64_t)second << 32) | (uint64_t)idx;
}

// grid: (query blocks, train splits). Split s scans rows [s*rows_per_split, ...) and writes partial[s*nq + q].
__global__ void __launch_bounds__(KN_THREADS) k_knn2_partial(const uint8_t* __restrict__ query, int64_t nq,
                                                             const uint8_t* __restrict__ train, int64_t nt, int64_t rows_per_split,
                                                             int64_t index_base, uint64_t* __restrict__ partial)
{
	__shared__ __align__(16) uint4 tile[2][KN_TILE * 2];
	const int tid = threadIdx.x;
	const int64_t q0 = (int64_t)blockIdx.x * KN_QB;
	const int64_t t_begin = (int64_t)blockIdx.y * rows_per_split;
	const int64_t t_end = min(nt, t_begin + rows_per_split);

	uint32_t q[KN_KQ][8];
#pragma unroll
	for (int k = 0; k < KN_KQ; k++)
	{
		const int64_t qi = q0 + (int64_t)k * KN_THREADS + tid;
		if (qi < nq)
		{
			const uint4* p = reinterpret_cast<const uint4*>(query + 32 * qi);
			const uint4 lo = __ldg(p), hi = __ldg(p + 1);
			q[k][0] = lo.x; q[k][1] = lo.y; q[k][2] = lo.z; q[k][3] = lo.w;
			q[k][4] = hi.x; q[k][5] = hi.y; q[k][6] = hi.z; q[k][7] = hi.w;
		}
		else
		{
#pragma unroll
			for (int i = 0; i < 8; i++) q[k][i] = 0;
		}
	}
	// running result over chunks
	uint32_t B[KN_KQ], S[KN_KQ], I[KN_KQ];
#pragma unroll
	for (int k = 0; k < KN_KQ; k++) { B[k] = 256; S[k] = 256; I[k] = 0xffffffffu; }

	for (int64_t c_begin = t_begin; c_begin < t_end; c_begin += KN_CHUNK)
	{
		const int64_t c_end = min(t_end, c_begin + (int64_t)KN_CHUNK);
		const int ntiles = (int)((c_end - c_begin + KN_TILE - 1) / KN_TILE);
		uint32_t k1[KN_KQ], k2[KN_KQ];
#pragma unroll
		for (int k = 0; k < KN_KQ; k++) { k1[k] = KN_NONE; k2[k] = KN_NONE; }

		auto load_tile = [&](int t, int buf) {
			// 128 rows x 32 B = 256 x 16 B; 2 per thread. Rows past the end are zero-filled and masked by key below.
			const int64_t row0 = c_begin + (int64_t)t * KN_TILE;
#pragma unroll
			for (int j = 0; j < 2; j++)
			{
				const int v = tid + j * KN_THREADS;
				const int64_t row = row0 + (v >> 1);
				if (row < c_end) cp_async16(&tile[buf][v], train + 32 * row + 16 * (v & 1));
				else tile[buf][v] = make_uint4(0, 0, 0, 0);
			}
			cp_async_commit();
		};
		load_tile(0, 0);
		for (int t = 0; t < ntiles; t++)
		{
			if (t + 1 < ntiles) { load_tile(t + 1, (t + 1) & 1); cp_async_wait<1>(); }
			else cp_async_wait<0>();
			__syncthreads();
			const uint4* tb = tile[t & 1];
			const int rows = (int)min((int64_t)KN_TILE, c_end - (c_begin + (int64_t)t * KN_TILE));
			const uint32_t jbase = (uint32_t)(t * KN_TILE);
			if (rows == KN_TILE)
			{
#pragma unroll 4
				for (int j = 0; j < KN_TILE; j++)
				{
					const uint4 lo = tb[2 * j], hi = tb[2 * j + 1];
					const uint32_t jj = jbase + j;
#pragma unroll
					for (int k = 0; k < KN_KQ; k++)
					{
						const int d = hamming256_csa(q[k], lo, hi);
						const uint32_t key = ((uint32_t)d << 22) + jj;
						k2[k] = min(k2[k], max(k1[k], key));
						k1[k] = min(k1[k], key);
					}
				}
			}
			else
			{
				for (int j = 0; j < rows; j++)
				{
					const uint4 lo = tb[2 * j], hi = tb[2 * j + 1];
					const uint32_t jj = jbase + j;
#pragma unroll
					for (int k = 0; k < KN_KQ; k++)
					{
						const int d = hamming256_csa(q[k], lo, hi);
						const uint32_t key = ((uint32_t)d << 22) + jj;
						k2[k] = min(k2[k], max(k1[k], key));
						k1[k] = min(k1[k], key);
					}
				}
			}
			__syncthreads();
		}
		// fold the chunk into the running result: chunks are visited in ascending index order
#pragma unroll
		for (int k = 0; k < KN_KQ; k++)
		{
			const uint32_t b = k1[k] >> 22, s = k2[k] >> 22;
			if (b < B[k])
			{
				S[k] = B[k];
				B[k] = b;
				I[k] = (uint32_t)(index_base + c_begin + (int64_t)(k1[k] & (KN_CHUNK - 1)));
			}
			else if (b < S[k]) S[k] = b;
			if (s < S[k]) S[k] = s;
		}
	}
#pragma unroll
	for (int k = 0; k < KN_KQ; k++)
	{
		const int64_t qi = q0 + (int64_t)k * KN_THREADS + tid;
		if (qi < nq) partial[(int64_t)blockIdx.y * nq + qi] = knn_pack(B[k], S[k], I[k]);
	}
}

// K9  knn2_merge — fold R partials per query (train splits of one GPU, or the all-gathered ranks). Equals one
// ascending scan over the whole train set: best = lexicographic min of (best, index); second = min(second of the
// winner, best of every other part).
__global__ void __launch_bounds__(256) k_knn2_merge(const uint64_t* __restrict__ parts, int R, int64_t nq, int th_low, float nnratio,
                                                    int32_t* __restrict__ idx, uint16_t* __restrict__ best, uint16_t* __restrict__ second,
                                                    int32_t* __restrict__ match, uint64_t* __restrict__ packed)
{
	const int64_t qi = (int64_t)blockIdx.x * 256 + threadIdx.x;
	if (qi >= nq) return;
	uint32_t B = 256, S = 256, I = 0xffffffffu;
	for (int r = 0; r < R; r++)
	{
		const uint64_t p = __ldg(parts + (int64_t)r * nq + qi);
		const uint32_t b = (uint32_t)(p >> 48), s = (uint32_t)((p >> 32) & 0xffffu), i = (uint32_t)p;
		if (b < B || (b == B && b < 256 && i < I))
		{
			if (B < S) S = B;       // the previous winner becomes a runner-up
			B = b; I = i;
		}
		else if (b < S) S = b;
		if (s < S) S = s;
	}
	if (packed) packed[qi] = knn_pack(B, S, I);
	if (idx) idx[qi] = (int32_t)I;
	if (best) best[qi] = (uint16_t)B;
	if (second) second[qi] = (uint16_t)S;
	if (match) match[qi] = ((int)B <= th_low && (float)B < __fmul_rn(nnratio, (float)S)) ? (int32_t)I : -1;
}

// ---------------------------------------------------------------------------------------------------
// M3/M4  stereo_match — ComputeStereoMatches (src/ORBmatcher.cc:72-247). One warp per left keypoint:
//   row-band candidate scan over the right keypoints in ascending index order (the reference's per-row lists hold
//   ascending indices, so "first strict minimum" = lexicographic min of (distance, index)), 11x11 SAD over 11
//   shifts (PatchDistance :60-68), parabola, disparity test. A second kernel applies the median cut (:231-246).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ const uint8_t* lvl_ptr(const uint8_t* l0, int64_t l0_stride, const uint8_t* slab, int64_t slab_stride,
                                                  const int64_t* off, int f, int o)
{
	return o == 0 ? l0 + (int64_t)f * l0_stride : slab + (int64_t)f * slab_stride + off[o];
}

// rowIndices (src/ORBmatcher.cc:84-100): every right keypoint is listed in the rows [floor(y - r), ceil(y + r)], r = 2 * scale[octave].
// One CTA per frame builds the lists as CSR with shared-memory counters. The order inside a row is arbitrary here; the match below
// reduces with min over (distance, iR), which is what the reference's ascending scan with a strict '<' computes.
constexpr int ST_MAX_ROWS = 4096;
__global__ void __launch_bounds__(1024) k_stereo_rows(const OrbxStereoArgs A)
{
	__shared__ int s_cnt[ST_MAX_ROWS];
	__shared__ int s_beg[ST_MAX_ROWS];
	__shared__ int s_w[33];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int f = blockIdx.x, nR = A.nr[f], rows = A.rows;
	const orbx_keypoint* __restrict__ kr = A.kr + (int64_t)f * A.cap;
	int* __restrict__ rs = A.row_start + (int64_t)f * (rows + 1);
	uint2* __restrict__ items = A.row_items + (int64_t)f * A.items_cap;
	for (int y = tid; y < rows; y += 1024) s_cnt[y] = 0;
	__syncthreads();
	for (int iR = tid; iR < nR; iR += 1024)
	{
		const orbx_keypoint k = kr[iR];
		const float r = __fmul_rn(2.f, A.scale[k.octave]);
		const int miny = max((int)floorf(__fsub_rn(k.y, r)), 0), maxy = min((int)ceilf(__fadd_rn(k.y, r)), rows - 1);
		for (int y = miny; y <= maxy; y++) atomicAdd(&s_cnt[y], 1);
	}
	__syncthreads();
	// exclusive scan of the row counts
	int base = 0;
	for (int y0 = 0; y0 < rows; y0 += 1024)
	{
		const int y = y0 + tid;
		const int v = y < rows ? s_cnt[y] : 0;
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
		for (int w = 0; w < 32; w++)
		{
			const int t = s_w[w];
			if (w < warp) wbase += t;
			tot += t;
		}
		if (y < rows) { s_beg[y] = base + wbase + inc - v; rs[y] = s_beg[y]; }
		base += tot;
		__syncthreads();
	}
	if (tid == 0) rs[rows] = min(base, A.items_cap);
	for (int iR = tid; iR < nR; iR += 1024)
	{
		const orbx_keypoint k = kr[iR];
		const float r = __fmul_rn(2.f, A.scale[k.octave]);
		const int miny = max((int)floorf(__fsub_rn(k.y, r)), 0), maxy = min((int)ceilf(__fadd_rn(k.y, r)), rows - 1);
		const uint2 it = make_uint2(__float_as_uint(k.x), (uint32_t)iR | ((uint32_t)k.octave << 16));
		for (int y = miny; y <= maxy; y++)
		{
			const int p = atomicAdd(&s_beg[y], 1);
			if (p < A.items_cap) items[p] = it;
		}
	}
}

__global__ void __launch_bounds__(256) k_stereo_match(const OrbxStereoArgs A)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int f = blockIdx.y;
	const int iL = blockIdx.x * 8 + warp;
	const int nL = A.nl[f], nR = A.nr[f];
	if (iL >= nL) return;
	const int64_t o_idx = (int64_t)f * A.cap + iL;
	const orbx_keypoint kl = A.kl[o_idx];
	const orbx_keypoint* __restrict__ kr = A.kr + (int64_t)f * A.cap;
	const uint8_t* __restrict__ dr = A.dr + (int64_t)f * A.cap * 32;

	float out_u = -1.f, out_d = -1.f;
	int out_sad = -1;

	const int TH_HIGH = 100, TH_LOW = 50, TH_ORB = (TH_HIGH + TH_LOW) / 2, R = 5, SR = 5;
	const float maxd = __fdiv_rn(A.bf, A.baseline);       // :102-104, minZ = baseline
	const float minu = __fsub_rn(kl.x, maxd), maxu = kl.x;
	const int row = (int)kl.y;                            // :122

	uint32_t dq[8];
	{
		const uint4* p = reinterpret_cast<const uint4*>(A.dl + o_idx * 32);
		const uint4 lo = __ldg(p), hi = __ldg(p + 1);
		dq[0] = lo.x; dq[1] = lo.y; dq[2] = lo.z; dq[3] = lo.w; dq[4] = hi.x; dq[5] = hi.y; dq[6] = hi.z; dq[7] = hi.w;
	}
	// key = dist << 16 | iR; strict "<" against TH_HIGH with lowest index winning
	uint32_t bestKey = ((uint32_t)TH_HIGH << 16);
	bool any = false;
	// candidates = rowIndices[(int)vL] (:122): one CSR range; lanes stride over it
	if (row >= 0 && row < A.rows)
	{
		const int* __restrict__ rs = A.row_start + (int64_t)f * (A.rows + 1);
		const uint2* __restrict__ items = A.row_items + (int64_t)f * A.items_cap;
		const int a = rs[row], b = min(rs[row + 1], A.items_cap);
		any = b > a;
		for (int p = a + lane; p < b; p += 32)
		{
			const uint2 it = items[p];
			const int octR = (int)(it.y >> 16), iR = (int)(it.y & 0xffffu);
			const float uR = __uint_as_float(it.x);
			if (!(octR < kl.octave - 1 || octR > kl.octave + 1) && uR >= minu && uR <= maxu)
			{
				const uint4* p4 = reinterpret_cast<const uint4*>(dr + (int64_t)iR * 32);
				const uint4 lo = __ldg(p4), hi = __ldg(p4 + 1);
				const int d = __popc(dq[0] ^ lo.x) + __popc(dq[1] ^ lo.y) + __popc(dq[2] ^ lo.z) + __popc(dq[3] ^ lo.w) +
				              __popc(dq[4] ^ hi.x) + __popc(dq[5] ^ hi.y) + __popc(dq[6] ^ hi.z) + __popc(dq[7] ^ hi.w);
				if (d < TH_HIGH) bestKey = min(bestKey, ((uint32_t)d << 16) | (uint32_t)iR);
			}
		}
	}
	any = __any_sync(0xffffffffu, any);
	bestKey = __reduce_min_sync(0xffffffffu, bestKey);          // one REDUX instead of five shuffle + min steps
	const int bestDist = (int)(bestKey >> 16), bestR = (int)(bestKey & 0xffffu);

	if (any && maxu >= 0.f && bestDist < TH_ORB)
	{
		const int o = kl.octave;
		const float sf = A.inv_scale[o];
		const int suL = (int)roundf(__fmul_rn(sf, kl.x)), svL = (int)roundf(__fmul_rn(sf, kl.y));
		const int suR = (int)roundf(__fmul_rn(sf, kr[bestR].x));
		if (!(suR + SR - R < 0 || suR + SR + R + 1 >= A.lw[o]))
		{
			const int64_t pL = o == 0 ? A.pl0_pitch : A.lpitch[o], pR = o == 0 ? A.pr0_pitch : A.lpitch[o];
			const uint8_t* __restrict__ IL = lvl_ptr(A.pl0, A.pl0_stride, A.pl, A.pl_slab, A.loff, f, o) + (int64_t)(svL - R) * pL + (suL - R);
			const uint8_t* __restrict__ IR = lvl_ptr(A.pr0, A.pr0_stride, A.pr, A.pr_slab, A.loff, f, o) + (int64_t)(svL - R) * pR + (suR - SR - R);
			const int cL = __ldg(IL + R * pL + R);
			// lane handles patch pixels lane, lane+32, lane+64, lane+96 (< 121)
			int lv[4], ly[4], lx[4];
#pragma unroll
			for (int j = 0; j < 4; j++)
			{
				const int p = lane + 32 * j;
				ly[j] = p / 11; lx[j] = p - ly[j] * 11;
				lv[j] = p < 121 ? (int)__ldg(IL + ly[j] * pL + lx[j]) : 0;
			}
			int bestSad = 0x7fffffff, bestDx = 0, dist[11];
#pragma unroll
			for (int s = 0; s < 11; s++)   // dxR = s - 5
			{
				const int sub = cL - (int)__ldg(IR + R * pR + R + s);
				int sum = 0;
#pragma unroll
				for (int j = 0; j < 4; j++)
					if (lane + 32 * j < 121) sum += abs(lv[j] - (int)__ldg(IR + ly[j] * pR + lx[j] + s) - sub);
				sum = __reduce_add_sync(0xffffffffu, sum);
				dist[s] = sum;
				if (sum < bestSad) { bestSad = sum; bestDx = s - SR; }
			}
			if (!(bestDx == -SR || bestDx == SR))
			{
				int d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
				for (int s = 0; s < 11; s++)
				{
					if (s == SR + bestDx - 1) d1 = dist[s];
					if (s == SR + bestDx) d2 = dist[s];
					if (s == SR + bestDx + 1) d3 = dist[s];
				}
				const float deltaR = __fdiv_rn((float)(d1 - d3), __fmul_rn(2.f, __fsub_rn((float)(d1 + d3), __fmul_rn(2.f, (float)d2))));
				if (!(deltaR < -1.f || deltaR > 1.f))
				{
					float bestuR = __fmul_rn(A.scale[o], __fadd_rn((float)(suR + bestDx), deltaR));
					float disparity = __fsub_rn(kl.x, bestuR);
					if (disparity >= 0.f && disparity < maxd)
					{
						if (disparity <= 0.f) { disparity = 0.01f; bestuR = __fsub_rn(kl.x, 0.01f); }
						out_d = __fdiv_rn(A.bf, disparity);
						out_u = bestuR;
						out_sad = bestSad;
					}
				}
			}
		}
	}
	if (lane == 0)
	{
		A.uright[o_idx] = out_u;
		A.depth[o_idx] = out_d;
		A.sad[o_idx] = out_sad;
	}
}

// median cut (:231-246): among kept matches sort SAD descending, median = element max(n/2-1,0), drop SAD >= 2.1*median.
__global__ void __launch_bounds__(256) k_stereo_median_cut(const OrbxStereoArgs A)
{
	__shared__ int s_cnt[8];
	const int f = blockIdx.x, tid = threadIdx.x;
	const int nL = A.nl[f];
	const int* __restrict__ sad = A.sad + (int64_t)f * A.cap;
	auto count_ge = [&](int v) {
		int c = 0;
		for (int i = tid; i < nL; i += 256) c += (sad[i] >= v);
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
		__syncthreads();
		if ((tid & 31) == 0) s_cnt[tid >> 5] = c;
		__syncthreads();
		int t = 0;
#pragma unroll
		for (int w = 0; w < 8; w++) t += s_cnt[w];
		return t;
	};
	const int kept = count_ge(0);
	if (kept == 0) return;      // the reference reads distIndices[0] of an empty vector here (:232-233); nothing to do
	const int m = max(kept / 2 - 1, 0);
	// median = largest v with count(sad >= v) >= m + 1
	int lo = 0, hi = 1 << 17;   // SAD <= 121 * 510 < 2^17
	while (hi - lo > 1)
	{
		const int mid = (lo + hi) >> 1;
		if (count_ge(mid) >= m + 1) lo = mid; else hi = mid;
	}
	const float th = __fmul_rn(1.5f * 1.4f, (float)lo);
	for (int i = tid; i < nL; i += 256)
	{
		const int s = sad[i];
		if (s >= 0 && !((float)s < th))
		{
			A.uright[(int64_t)f * A.cap + i] = -1.f;
			A.depth[(int64_t)f * A.cap + i] = -1.f;
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// ComputeStereoFromRGBD (src/System.cc:197-219): one thread per keypoint
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stereo_from_rgbd(const orbx_keypoint* __restrict__ kps, const orbx_keypoint* __restrict__ kps_un, int n,
                                                          const uint8_t* __restrict__ depth_map, int64_t pitch, float bf,
                                                          float* __restrict__ uright, float* __restrict__ depth)
{
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	const int v = (int)kps[i].y, u = (int)kps[i].x;       // truncation, :210-211
	const float d = __ldg(reinterpret_cast<const float*>(depth_map + (int64_t)v * pitch) + u);
	float ur = -1.f, dp = -1.f;
	if (d > 0.f)
	{
		dp = d;
		ur = __fsub_rn(kps_un[i].x, __fdiv_rn(bf, d));
	}
	uright[i] = ur;
	depth[i] = dp;
}

// ---------------------------------------------------------------------------------------------------
// The N x N Hamming site of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:286-314): for every descriptor the
// median (sorted position (N-1)/2, the zero self-distance included) of its distances to the set; the result is the first
// descriptor with the least median. One CTA per set, one warp per row: 257-bin histogram in shared memory + prefix scan.
// ---------------------------------------------------------------------------------------------------
#define DD_WARPS 4
__global__ void __launch_bounds__(DD_WARPS * 32) k_distinctive(const uint8_t* __restrict__ desc, const int64_t* __restrict__ offsets,
                                                              int32_t* __restrict__ best)
{
	__shared__ int hist[DD_WARPS][288];
	__shared__ unsigned long long s_best;       // median << 32 | row: the minimum is the first row with the least median
	const int set = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int64_t b = offsets[set];
	const int n = (int)(offsets[set + 1] - b);
	if (threadIdx.x == 0) s_best = ~0ull;
	__syncthreads();
	const uint4* __restrict__ d = reinterpret_cast<const uint4*>(desc + 32 * b);
	const int k = (n - 1) / 2;
	for (int i = warp; i < n; i += DD_WARPS)
	{
		for (int t = lane; t < 288; t += 32) hist[warp][t] = 0;
		__syncwarp();
		const uint4 lo = __ldg(d + 2 * i), hi = __ldg(d + 2 * i + 1);
		for (int j = lane; j < n; j += 32)
		{
			const uint4 a = __ldg(d + 2 * j), c = __ldg(d + 2 * j + 1);
			const int dist = __popc(lo.x ^ a.x) + __popc(lo.y ^ a.y) + __popc(lo.z ^ a.z) + __popc(lo.w ^ a.w) +
			                 __popc(hi.x ^ c.x) + __popc(hi.y ^ c.y) + __popc(hi.z ^ c.z) + __popc(hi.w ^ c.w);
			atomicAdd(&hist[warp][dist], 1);
		}
		__syncwarp();
		// smallest value v with count(dist <= v) >= k + 1: lane l owns bins 9l .. 9l+8
		int c[9], sum = 0;
#pragma unroll
		for (int t = 0; t < 9; t++) { c[t] = hist[warp][9 * lane + t]; sum += c[t]; }
		int incl = sum;
#pragma unroll
		for (int dd = 1; dd < 32; dd <<= 1)
		{
			const int t = __shfl_up_sync(0xffffffffu, incl, dd);
			if (lane >= dd) incl += t;
		}
		int run = incl - sum, median = 0x7fffffff;
#pragma unroll
		for (int t = 0; t < 9; t++)
		{
			run += c[t];
			if (median == 0x7fffffff && run >= k + 1) median = 9 * lane + t;
		}
#pragma unroll
		for (int dd = 16; dd > 0; dd >>= 1) median = min(median, __shfl_xor_sync(0xffffffffu, median, dd));
		if (lane == 0) atomicMin(&s_best, ((unsigned long long)(unsigned)median << 32) | (unsigned)i);
		__syncwarp();
	}
	__syncthreads();
	if (threadIdx.x == 0) best[set] = n > 0 ? (int32_t)(s_best & 0xffffffffu) : -1;
}

// ---------------------------------------------------------------------------------------------------
// POPC-pipe probe: dependent-free popcounts on registers, all SMs; the matcher's roofline denominator.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_popc_probe(uint32_t* out, int iters)
{
	uint32_t x0 = threadIdx.x * 2654435761u + blockIdx.x, x1 = x0 ^ 0x9e3779b9u, x2 = x0 + 0x7f4a7c15u, x3 = ~x0;
	uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
	for (int i = 0; i < iters; i++)
	{
#pragma unroll
		for (int j = 0; j < 16; j++)
		{
			a0 += __popc(x0); a1 += __popc(x1); a2 += __popc(x2); a3 += __popc(x3);
			x0 += a3; x1 += a0; x2 += a1; x3 += a2;
		}
	}
	if (a0 + a1 + a2 + a3 == 0xdeadbeefu) out[0] = a0;
}

}  // namespace

void orbx_launch_hamming_pairs(const uint8_t* a, const uint8_t* b, int64_t n, int32_t* out, cudaStream_t st)
{
	if (n <= 0) return;
	k_hamming_pairs<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, b, n, out);
}

// number of train splits so that the grid fills the GPU at small nq (~4 CTAs of 128 threads per SM)
static int knn_splits(int64_t nq, int64_t nt)
{
	const int64_t qblocks = (nq + KN_QB - 1) / KN_QB;
	int64_t want = (148 * 4 + qblocks - 1) / qblocks;
	const int64_t max_splits = (nt + 4 * KN_TILE - 1) / (4 * KN_TILE);
	if (want > max_splits) want = max_splits;
	if (want < 1) want = 1;
	if (want > 64) want = 64;
	return (int)want;
}

int orbx_knn2_splits(int64_t nq, int64_t nt) { return knn_splits(nq, nt); }

void orbx_launch_knn2_partial(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int64_t base, uint64_t* partial,
                              cudaStream_t st)
{
	// `partial` must hold knn_splits(nq, nt) * nq entries
	const int splits = knn_splits(nq, nt);
	int64_t rows = (nt + splits - 1) / splits;
	rows = (rows + KN_TILE - 1) / KN_TILE * KN_TILE;
	dim3 grid((unsigned)((nq + KN_QB - 1) / KN_QB), splits);
	k_knn2_partial<<<grid, KN_THREADS, 0, st>>>(q, nq, t, nt, rows, base, partial);
}

void orbx_launch_knn2_merge(const uint64_t* gathered, int ranks, int64_t nq, int th_low, float nnratio, int32_t* idx,
                            uint16_t* best, uint16_t* second, int32_t* match, cudaStream_t st)
{
	if (nq <= 0) return;
	k_knn2_merge<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(gathered, ranks, nq, th_low, nnratio, idx, best, second, match, nullptr);
}

void orbx_launch_knn2_fold(const uint64_t* parts, int nparts, int64_t nq, uint64_t* packed, cudaStream_t st)
{
	if (nq <= 0) return;
	k_knn2_merge<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(parts, nparts, nq, 0, 0.f, nullptr, nullptr, nullptr, nullptr, packed);
}

void orbx_launch_stereo_from_rgbd(const orbx_keypoint* kps, const orbx_keypoint* kps_un, int n, const uint8_t* depth_map, int64_t pitch, float bf,
                                  float* uright, float* depth, cudaStream_t st)
{
	if (n <= 0) return;
	k_stereo_from_rgbd<<<(n + 255) / 256, 256, 0, st>>>(kps, kps_un, n, depth_map, pitch, bf, uright, depth);
}

void orbx_launch_distinctive(const uint8_t* desc, const int64_t* offsets, int nsets, int32_t* best, cudaStream_t st)
{
	if (nsets <= 0) return;
	k_distinctive<<<nsets, DD_WARPS * 32, 0, st>>>(desc, offsets, best);
}

// UndistortKeyPoints (src/System.cc:153-174) = cv::undistortPoints(pts, pts, K, distCoeffs, noArray(), K): OpenCV's double-precision
// fixed-point iteration (5 rounds, no epsilon test), then the camera matrix again, rounded to float. FP64 on the device, no contraction.
struct UndistortArgs { double fx, fy, cx, cy, k[14]; };
__global__ void __launch_bounds__(256) k_undistort_keypoints(const orbx_keypoint* __restrict__ src, orbx_keypoint* __restrict__ dst, int n, const UndistortArgs A)
{
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	orbx_keypoint kp = src[i];
	const double ifx = __ddiv_rn(1.0, A.fx), ify = __ddiv_rn(1.0, A.fy);
	const double u = (double)kp.x, v = (double)kp.y;
	double x = __dmul_rn(__dsub_rn(u, A.cx), ifx), y = __dmul_rn(__dsub_rn(v, A.cy), ify);
	const double x0 = x, y0 = y;
	const double* k = A.k;
	for (int j = 0; j < 5; j++)
	{
		const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
		const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[7], r2), k[6]), r2), k[5]), r2));
		const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[4], r2), k[1]), r2), k[0]), r2));
		const double icdist = __ddiv_rn(num, den);
		if (icdist < 0) { x = __dmul_rn(__dsub_rn(u, A.cx), ifx); y = __dmul_rn(__dsub_rn(v, A.cy), ify); break; }
		// deltaX = 2*k[2]*x*y + k[3]*(r2 + 2*x*x) + k[8]*r2 + k[9]*r2*r2, left to right
		const double dX = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, k[2]), x), y),
		                                                __dmul_rn(k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)))),
		                                      __dmul_rn(k[8], r2)), __dmul_rn(__dmul_rn(k[9], r2), r2));
		// deltaY = k[2]*(r2 + 2*y*y) + 2*k[3]*x*y + k[10]*r2 + k[11]*r2*r2
		const double dY = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
		                                                __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[3]), x), y)),
		                                      __dmul_rn(k[10], r2)), __dmul_rn(__dmul_rn(k[11], r2), r2));
		x = __dmul_rn(__dsub_rn(x0, dX), icdist);
		y = __dmul_rn(__dsub_rn(y0, dY), icdist);
	}
	// RR = K * I: xx = fx*x + 0*y + cx, yy = 0*x + fy*y + cy, ww = 1 / (0*x + 0*y + 1)
	const double xx = __dadd_rn(__dadd_rn(__dmul_rn(A.fx, x), __dmul_rn(0.0, y)), A.cx);
	const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(A.fy, y)), A.cy);
	const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
	kp.x = __double2float_rn(__dmul_rn(xx, ww));
	kp.y = __double2float_rn(__dmul_rn(yy, ww));
	dst[i] = kp;
}

void orbx_launch_undistort(const orbx_keypoint* src, orbx_keypoint* dst, int n, const float* cam4, const float* dist, int ndist, cudaStream_t st)
{
	UndistortArgs A;
	A.fx = cam4[0]; A.fy = cam4[1]; A.cx = cam4[2]; A.cy = cam4[3];
	for (int i = 0; i < 14; i++) A.k[i] = i < ndist ? (double)dist[i] : 0.0;
	k_undistort_keypoints<<<(n + 255) / 256, 256, 0, st>>>(src, dst, n, A);
}

int orbx_stereo_items_per_keypoint(float max_scale) { return 2 * (int)ceilf(2.f * max_scale) + 3; }

void orbx_launch_stereo(const OrbxStereoArgs& A, cudaStream_t st)
{
	dim3 grid((A.cap + 7) / 8, A.frames);
	k_stereo_rows<<<A.frames, 1024, 0, st>>>(A);
	k_stereo_match<<<grid, 256, 0, st>>>(A);
	k_stereo_median_cut<<<A.frames, 256, 0, st>>>(A);
}

double orbx_popc_probe(int device)
{
	cudaSetDevice(device);
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 0.0;
	uint32_t* d = nullptr;
	if (cudaMalloc(&d, 4) != cudaSuccess) return 0.0;
	const int blocks = prop.multiProcessorCount * 8, iters = 4096;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_popc_probe<<<blocks, 256>>>(d, 64);
	cudaDeviceSynchronize();
	double best = 0.0;
	for (int rep = 0; rep < 3; rep++)
	{
		cudaEventRecord(e0);
		k_popc_probe<<<blocks, 256>>>(d, iters);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		const double popc = (double)blocks * 256.0 * iters * 64.0;
		best = fmax(best, popc / (ms * 1e-3));
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	cudaFree(d);
	return best;
}
