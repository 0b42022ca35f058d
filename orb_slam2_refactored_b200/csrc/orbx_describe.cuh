// Grouped orientation + descriptor kernel (included by orbx_extract.cu inside its anonymous namespace, after the lookup tables).
// =====================================================================================================
// k_orient_describe2 — IC_Angle (src/ORBextractor.cc:74-101), ComputeOrbDescriptor (:103-140) and the keypoint record of Extract
// (:768-773, :811-815) for a GROUP of 8 consecutive output keypoints per warp. What that buys over one keypoint per warp:
//   * the patches are double-buffered: while keypoint k is reduced / sampled, the cp.async copies of keypoint k + 2 are in flight;
//   * fastAtan2 and the FP64 cos / sin run ONCE per group with one keypoint per lane instead of 32 identical copies per keypoint;
//   * the lane's 8 BRIEF pairs and its disc-row coefficient words live in registers for the whole group;
//   * cvRound is one FADD with 1.5 * 2^23 (the rounded integer appears in the low mantissa bits, ties to even like cvRound's
//     lrint), so the quarter-rate conversion pipe is not used and the bias folds into the patch base address.
// Float semantics as before: every reference float op is an explicit round-to-nearest intrinsic.
// (Blackwell's packed FP32 pairs — mul/fma.rn.f32x2, SASS FMUL2 / FFMA2 — would halve the rotation's float instructions, but ptxas 12.9
// contracts mul.rn.f32x2 into a following add / fma(p, 1.0, q) even with -fmad=false, dropping a rounding the reference performs: one
// descriptor bit in 300 000 differed on a noise image (tests/test_gpu_extract.py::test_noise_images_overflow_the_cell_list). Scalar
// __fmul_rn / __fadd_rn are never contracted.)
// =====================================================================================================
#define OD2_IPS 48                   // un-blurred patch: 31 rows x 48 bytes from (x - 16) & ~15 (three 16-byte chunks: a copy instruction then spans ~11 rows,
                                     // 4-byte copies with one row per lane were measured 1.5x slower overall: 31 cache lines per instruction)
#define OD2_IBUF (31 * OD2_IPS)      // 1488 bytes; a ring of four, so four keypoints' patches are in flight
#define OD2_BPS 80                   // blurred patch: 37 rows x 64 bytes from (x - 18) & ~15, stride 80 spreads the rows over the banks
#define OD2_BUF (37 * OD2_BPS)       // 2960 bytes; a ring of two that re-uses the memory of the (by then dead) un-blurred ring
#define OD2_SMEM (4 * OD2_IBUF > 2 * OD2_BUF ? 4 * OD2_IBUF : 2 * OD2_BUF)
#define OD2_MAGIC 12582912.f         // 1.5 * 2^23
#define OD2_MAGIC_BITS 0x4B400000

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem)
{
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr)
{
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(saddr));
	return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// G = keypoints per warp: 16 (or 8) for throughput; 2 when a frame at a time is extracted (the launch then has four times the warps and a
// quarter of the serial chain per warp)
template <int G>
__global__ void __launch_bounds__(32, 32) k_orient_describe2(const OrbxPlanDev P, orbx_keypoint* __restrict__ d_kps, uint8_t* __restrict__ d_desc,
                                                        int32_t* __restrict__ d_n)
{
	extern __shared__ __align__(16) uint8_t od2_smem[];
	const int lane = threadIdx.x, f = blockIdx.y;
	const int slot0 = blockIdx.x * G;

	// levels are concatenated in order (:792-819): lane l holds level l's count and the inclusive prefix
	const int cnt = (lane < P.nlevels) ? P.sel_count[(int64_t)f * P.nlevels + lane] : 0;
	int incl = cnt;
#pragma unroll
	for (int d = 1; d < 16; d <<= 1)
	{
		const int t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += t;
	}
	const int total = __shfl_sync(0xffffffffu, incl, ORBX_MAX_LEVELS - 1);
	if (slot0 == 0 && lane == 0) d_n[f] = total;
	const int nk = min(min(total, P.out_cap) - slot0, G);
	if (nk <= 0) return;

	// lane k < nk owns keypoint slot0 + k: its level, position, and the origins of its two patches
	int lvl = 0;
#pragma unroll
	for (int s = 0; s < ORBX_MAX_LEVELS; s++)
	{
		const int v = __shfl_sync(0xffffffffu, incl, s);
		if (s < P.nlevels - 1 && slot0 + lane >= v) lvl = s + 1;
	}
	const int start = __shfl_sync(0xffffffffu, incl - cnt, lvl);
	int x = 0, y = 0, resp = 0;
	const uint8_t* gi = nullptr; const uint8_t* gb = nullptr;
	int ipitch = 0, bpitch = 0;
	if (lane < nk)
	{
		const OrbxLevel& L = P.lv[lvl];
		const uint32_t kp = P.sel[(int64_t)f * P.sel_per_frame + L.sel_base + (slot0 + lane - start)];
		x = orbx_px(kp); y = orbx_py(kp); resp = orbx_pr(kp);
		ipitch = (int)orbx_level_pitch(P, lvl); bpitch = L.pitch;
		gi = orbx_level_ptr(P, f, lvl) + (int64_t)(y - 15) * ipitch + ((x - 16) & ~15);
		gb = P.blur + (int64_t)f * P.slab + L.offset + (int64_t)(y - 18) * bpitch + ((x - 18) & ~15);
	}
	// the lane's constants for the whole group: disc-row coefficient words (lane = row v = lane - 15) and its 8 BRIEF pairs
	uint32_t mones[8], mus[8];
	{
		const int av = abs(lane - ORBX_HALF_PATCH) & 15;
#pragma unroll
		for (int k = 0; k < 8; k++) { const uint2 cf = __ldg(g_mom + k * 16 + av); mones[k] = cf.x; mus[k] = cf.y; }
	}
	auto bcast_ptr = [&](const uint8_t* p, int k) {
		const unsigned long long v = (unsigned long long)p;
		const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, k), hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), k);
		return (const uint8_t*)(((unsigned long long)hi << 32) | lo);
	};
	// ---- phase 1: intensity centroids; the patches of four keypoints are in flight at a time
	// chunk i = lane + 32 j (j = 0, 1, 2; 93 chunks) is chunk i % 3 of row i / 3: the same for every keypoint
	int crow[3], ccol[3];
#pragma unroll
	for (int j = 0; j < 3; j++) { const int i = lane + 32 * j; crow[j] = i / 3; ccol[j] = (i - 3 * crow[j]) * 16; }
	auto stage_img = [&](int k) {
		const uint8_t* g = bcast_ptr(gi, k);
		const int pitch = __shfl_sync(0xffffffffu, ipitch, k);
		uint8_t* dst = od2_smem + (k & 3) * OD2_IBUF;
#pragma unroll
		for (int j = 0; j < 3; j++)
			if (j < 2 || lane < 29) cp_async16(dst + crow[j] * OD2_IPS + ccol[j], g + (int64_t)crow[j] * pitch + ccol[j]);
		cp_async_commit();
	};
	int my01 = 0, my10 = 0;
#pragma unroll
	for (int k = 0; k < 4; k++)
		if (k < nk) stage_img(k);
	for (int k = 0; k < nk; k++)
	{
		// groups are committed in keypoint order: keypoint k has landed once at most min(3, nk - 1 - k) newer groups are pending
		const int newer = min(3, nk - 1 - k);
		if (newer == 3) cp_async_wait<3>(); else if (newer == 2) cp_async_wait<2>(); else if (newer == 1) cp_async_wait<1>(); else cp_async_wait<0>();
		__syncwarp();
		const int o = __shfl_sync(0xffffffffu, x - 16, k) & 15;      // byte of column x - 16 inside the staged row
		const int shb = (o & 3) * 8;
		int m10 = 0, m01 = 0;
		if (lane < 31)
		{
			const uint32_t* wp = reinterpret_cast<const uint32_t*>(od2_smem + (k & 3) * OD2_IBUF + lane * OD2_IPS) + (o >> 2);
			uint32_t w[9];
#pragma unroll
			for (int j = 0; j < 9; j++) w[j] = wp[j];
			int rowsum = 0;
#pragma unroll
			for (int j = 0; j < 8; j++)
			{
				const uint32_t win = __funnelshift_r(w[j], w[j + 1], shb);
				rowsum = (int)__dp4a(win, mones[j], (uint32_t)rowsum);
				m10 = dp4a_u8_s8(win, mus[j], m10);
			}
			m01 = (lane - ORBX_HALF_PATCH) * rowsum;
		}
		m10 = __reduce_add_sync(0xffffffffu, m10);          // REDUX: one instruction per sum instead of five shuffle + add steps
		m01 = __reduce_add_sync(0xffffffffu, m01);
		if (lane == k) { my01 = m01; my10 = m10; }
		__syncwarp();                       // every lane is done with this buffer before the copies of keypoint k + 4 land in it
		if (k + 4 < nk) stage_img(k + 4);
	}

	// ---- phase 2: the blurred patches of the first two keypoints start flying, then angle, cos, sin with one keypoint per lane
	auto stage_blr = [&](int k) {
		const uint8_t* g = bcast_ptr(gb, k);
		const int pitch = __shfl_sync(0xffffffffu, bpitch, k);
		const uint8_t* src = g + (int64_t)(lane >> 2) * pitch + (lane & 3) * 16;
		uint8_t* dst = od2_smem + (k & 1) * OD2_BUF + (lane >> 2) * OD2_BPS + (lane & 3) * 16;
#pragma unroll
		for (int i = 0; i < 5; i++)
			if (i < 4 || lane < 20) cp_async16(dst + i * 8 * OD2_BPS, src + (int64_t)(i * 8) * pitch);      // rows (lane >> 2) + 8 i < 37
		cp_async_commit();
	};
	stage_blr(0);
	if (nk > 1) stage_blr(1);
	// the lane's 8 BRIEF pairs, loaded only now: phase 1's coefficient words are dead, so the two sets of constants never hold registers
	// at the same time (80 -> 64 registers, 32 instead of 24 resident warps); the loads fly during the angle / cos / sin below
	float4 pat[8];
#pragma unroll
	for (int bit = 0; bit < 8; bit++) pat[bit] = __ldg(g_patf + bit * 32 + lane);
	float angle = 0.f, ca = 0.f, sb = 0.f;
	if (lane < nk)
	{
		angle = fast_atan2_deg((float)my01, (float)my10);
		orb_cos_sin(angle, ca, sb);
	}

	// ---- phase 3: steered BRIEF (:103-140): lane = descriptor byte, 8 pairs each, samples from the staged patch
	for (int k = 0; k < nk; k++)
	{
		if (k + 1 < nk) cp_async_wait<1>(); else cp_async_wait<0>();
		__syncwarp();
		const float a = __shfl_sync(0xffffffffu, ca, k), b = __shfl_sync(0xffffffffu, sb, k);
		const int xo = (__shfl_sync(0xffffffffu, x - 18, k) & 15) + 18;                  // column of the keypoint inside the staged rows
		// shared address of sample (r, q) = base + r * BPS + q with r, q taken as the raw bits of the magic-rounded floats: the bias
		// 0x4B400000 * (BPS + 1) is folded into the base, all in 32-bit modular arithmetic (shared addresses are 32 bits)
		const uint32_t bl = (uint32_t)__cvta_generic_to_shared(od2_smem + (k & 1) * OD2_BUF + 18 * OD2_BPS + xo) - (uint32_t)OD2_MAGIC_BITS * (uint32_t)(OD2_BPS + 1);
		uint32_t byte = 0;
#pragma unroll
		for (int bit = 0; bit < 8; bit++)
		{
			const float4 pt = pat[bit];
			const uint32_t r0 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(pt.x, b), __fmul_rn(pt.y, a)), OD2_MAGIC));
			const uint32_t q0 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(pt.x, a), __fmul_rn(pt.y, b)), OD2_MAGIC));
			const uint32_t r1 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(pt.z, b), __fmul_rn(pt.w, a)), OD2_MAGIC));
			const uint32_t q1 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(pt.z, a), __fmul_rn(pt.w, b)), OD2_MAGIC));
			const uint32_t t0 = lds_u8(bl + r0 * OD2_BPS + q0), t1 = lds_u8(bl + r1 * OD2_BPS + q1);
			byte |= (uint32_t)(t0 < t1) << bit;
		}
		d_desc[((int64_t)f * P.out_cap + slot0 + k) * 32 + lane] = (uint8_t)byte;
		__syncwarp();
		if (k + 2 < nk) stage_blr(k + 2);
	}

	if (lane < nk)
	{
		const float scale = P.lv[lvl].scale;
		orbx_keypoint o;
		o.x = (float)x; o.y = (float)y;
		if (lvl > 0) { o.x = __fmul_rn(o.x, scale); o.y = __fmul_rn(o.y, scale); }   // :811-815
		o.size = __fmul_rn(scale, (float)ORBX_PATCH);                                // :771
		o.angle = angle;
		o.response = (float)resp;
		o.octave = lvl;
		o.class_id = -1;
		d_kps[(int64_t)f * P.out_cap + slot0 + lane] = o;
	}
}
