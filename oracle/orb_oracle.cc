// TEST INFRASTRUCTURE — CPU oracle, not the product.
//
// Stand-alone restatement of the reference's ORB front-end hot path, written
// against plain arrays so that it builds anywhere gcc does. It is NOT a copy of the reference's
// code: the quadtree in particular is restated in the pass-structured, array-based form the CUDA
// kernels use (segments of a permutation array instead of std::list<QTreeNode> with per-node
// vectors), which makes this file the executable spec of the device algorithm. It is pinned by
// tests/test_oracle_vs_ref.py against oracle/_ref (the reference's own TUs) stage by stage and end
// to end, and by tests/test_oracle_golden.py against committed golden outputs of oracle/_ref.
//
// Every function cites the reference lines it follows.
#define ORACLE_PREFIX orc_
#include "oracle_api.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "cv_primitives.h"
#include "stdsort_replay.h"

namespace {

const int kPatch = 31;        // PATCH_SIZE       src/ORBextractor.cc:68
const int kHalfPatch = 15;    // HALF_PATCH_SIZE  :69
const int kEdge = 19;         // EDGE_THRESHOLD   :70
const int kBorder = kEdge - 3;   // :755
const int kCell = 30;         // CELL_SIZE        :491

const signed char kPattern[1024] = {
#include "orb_pattern.inc"
};

// umax_ of ORBextractor::Init (:705-718): half-widths of the rows of the radius-15 disc
struct UMax
{
	int v[kHalfPatch + 1];
	UMax()
	{
		const int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.) / 2 + 1);
		const int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.) / 2);
		for (int i = 0; i <= vmax; i++)
			v[i] = cvp::round_rne(std::sqrt((double)(kHalfPatch * kHalfPatch - i * i)));
		for (int i = kHalfPatch, i0 = 0; i >= vmin; --i)
		{
			while (v[i0] == v[i0 + 1]) ++i0;
			v[i] = i0;
			++i0;
		}
	}
};
const UMax kUMax;

struct Image
{
	int w = 0, h = 0;
	std::vector<uint8_t> px;
	void alloc(int w_, int h_) { w = w_; h = h_; px.assign((size_t)w * h, 0); }
	const uint8_t* row(int y) const { return px.data() + (size_t)y * w; }
};

// ---------------------------------------------------------------------------------------------
// E2  DetectFAST (:489-540) in the "one score map per level" form (validated):
// cell view [x0,x1) x [y0,y1); detection region = view shrunk by 3; cell-local strict NMS; retry
// at minTh iff the NMS'd result at iniTh is empty (:529); emission cell-major, row-major inside.
// ---------------------------------------------------------------------------------------------
void detect_fast(const uint8_t* img, int w, int h, size_t pitch, int iniTh, int minTh, std::vector<oracle_cand>& out)
{
	out.clear();
	const int rw = w - 2 * kBorder, rh = h - 2 * kBorder;   // roi of Extract (:760)
	const int minx = kBorder, miny = kBorder, maxx = minx + rw, maxy = miny + rh;
	const int gridw = rw / kCell, gridh = rh / kCell;
	const int cellw = (int)std::ceil(1. * rw / gridw), cellh = (int)std::ceil(1. * rh / gridh);

	// arc score of every pixel that any cell can test: [19, w-19) x [19, h-19); clamp to >= 0
	std::vector<uint8_t> S((size_t)w * h, 0);
	for (int y = miny + 3; y < maxy - 3; y++)
		for (int x = minx + 3; x < maxx - 3; x++)
		{
			const int s = cvp::fast9_arc_score(img + (size_t)y * pitch + x, pitch);
			S[(size_t)y * w + x] = (uint8_t)std::max(0, std::min(255, s));
		}

	std::vector<oracle_cand> cell;
	for (int cy = 0, y0 = miny; cy < gridh && y0 + 6 < maxy; cy++, y0 += cellh)
		for (int cx = 0, x0 = minx; cx < gridw && x0 + 6 < maxx; cx++, x0 += cellw)
		{
			const int y1 = std::min(y0 + cellh + 6, maxy), x1 = std::min(x0 + cellw + 6, maxx);
			const int rx0 = x0 + 3, rx1 = x1 - 3, ry0 = y0 + 3, ry1 = y1 - 3;
			auto M = [&](int x, int y, int t) -> int {
				if (x < rx0 || x >= rx1 || y < ry0 || y >= ry1) return 0;
				const int s = S[(size_t)y * w + x];
				return s > t ? s - 1 : 0;
			};
			for (int pass = 0; pass < 2; pass++)
			{
				const int t = pass == 0 ? iniTh : minTh;
				cell.clear();
				for (int y = ry0; y < ry1; y++)
					for (int x = rx0; x < rx1; x++)
					{
						const int m = M(x, y, t);
						if (m <= 0) continue;
						if (m > M(x - 1, y - 1, t) && m > M(x, y - 1, t) && m > M(x + 1, y - 1, t) &&
						    m > M(x - 1, y, t) && m > M(x + 1, y, t) &&
						    m > M(x - 1, y + 1, t) && m > M(x, y + 1, t) && m > M(x + 1, y + 1, t))
							cell.push_back({ x, y, m });
					}
				if (!cell.empty()) break;
			}
			out.insert(out.end(), cell.begin(), cell.end());
		}
}

using stdsort::SortItem;
using stdsort::libstdcxx_sort_desc;

// ---------------------------------------------------------------------------------------------
// E3  QuadTreeSuppression + QTreeNode::divide (:402-453, :542-693), array form.
// A node owns the segment perm[beg, beg+cnt) of candidate indices, in the order the reference's
// per-node vector would hold them. `order` is the std::list front->back.
// ---------------------------------------------------------------------------------------------
struct Node { int x0, y0, x1, y1, beg, cnt; };

void quadtree(const std::vector<oracle_cand>& cand, int w, int h, int quota, std::vector<oracle_cand>& out)
{
	const int n = (int)cand.size();
	const int rx = kBorder, ry = kBorder, rw = w - 2 * kBorder, rh = h - 2 * kBorder;
	if (n == 0 || rw <= 0 || rh <= 0) { out = cand; return; }   // early return leaves dst (== src) untouched (:544-545)

	std::vector<int> perm(n), tmp(n);
	std::vector<Node> nodes;      // arena
	std::vector<int> order;       // list order, ids into the arena

	// roots (:547-581): n0 vertical strips; empty ones dropped
	const int n0 = cvp::round_rne(1. * rw / rh);
	const double hx = 1. * rw / n0;
	{
		std::vector<int> cnt(n0, 0), at(n0, 0);
		std::vector<int> which(n);
		for (int i = 0; i < n; i++)
		{
			which[i] = (int)(((float)cand[i].x - rx) / hx);
			cnt[which[i]]++;
		}
		for (int r = 1; r < n0; r++) at[r] = at[r - 1] + cnt[r - 1];
		std::vector<int> fill = at;
		for (int i = 0; i < n; i++) perm[fill[which[i]]++] = i;
		for (int r = 0; r < n0; r++)
		{
			if (cnt[r] == 0) continue;
			nodes.push_back({ (int)(rx + hx * r), ry, (int)(rx + hx * (r + 1)), ry + rh, at[r], cnt[r] });
			order.push_back((int)nodes.size() - 1);
		}
	}

	// divide one node (:406-447): stable 4-way split of its segment; returns ids of non-empty children in
	// TL,TR,BL,BR order
	auto divide = [&](int id, int child_ids[4]) -> int {
		const Node p = nodes[id];
		const int hx2 = (int)std::ceil(0.5 * (p.x1 - p.x0)), hy2 = (int)std::ceil(0.5 * (p.y1 - p.y0));
		const int xm = p.x0 + hx2, ym = p.y0 + hy2;
		int c[4] = { 0, 0, 0, 0 };
		auto quadrant = [&](int i) {
			const float x = (float)cand[i].x, y = (float)cand[i].y;
			return x < xm ? (y < ym ? 0 : 2) : (y < ym ? 1 : 3);
		};
		for (int k = 0; k < p.cnt; k++) c[quadrant(perm[p.beg + k])]++;
		int at[4] = { p.beg, p.beg + c[0], p.beg + c[0] + c[1], p.beg + c[0] + c[1] + c[2] };
		int fill[4] = { at[0], at[1], at[2], at[3] };
		for (int k = 0; k < p.cnt; k++) { const int i = perm[p.beg + k]; tmp[fill[quadrant(i)]++] = i; }
		std::copy(tmp.begin() + p.beg, tmp.begin() + p.beg + p.cnt, perm.begin() + p.beg);
		const Node kids[4] = {
			{ p.x0, p.y0, xm, ym, at[0], c[0] }, { xm, p.y0, p.x1, ym, at[1], c[1] },
			{ p.x0, ym, xm, p.y1, at[2], c[2] }, { xm, ym, p.x1, p.y1, at[3], c[3] } };
		int m = 0;
		for (int q = 0; q < 4; q++)
			if (c[q] > 0) { nodes.push_back(kids[q]); child_ids[m++] = (int)nodes.size() - 1; }
		return m;
	};

	// One pass = divide `todo` in the given order (stopping early once the list would reach `stop_at`
	// nodes, Phase 2's break at :666-667; pass a huge value for Phase 1), then rebuild the list:
	// children of this pass in reverse push order, followed by the old list minus the divided nodes.
	std::vector<int> divisibles;   // children with > 1 point, in push order (:617-622 / :657-662)
	auto run_pass = [&](const std::vector<int>& todo, size_t stop_at) {
		std::vector<int> pushed;
		std::vector<char> gone(nodes.size() + 4 * todo.size() + 4, 0);
		size_t live = order.size();
		divisibles.clear();
		for (int id : todo)
		{
			int kids[4];
			const int m = divide(id, kids);
			for (int k = 0; k < m; k++)
			{
				pushed.push_back(kids[k]);
				if (nodes[kids[k]].cnt > 1) divisibles.push_back(kids[k]);
			}
			gone[id] = 1;
			live += m - 1;
			if (live >= stop_at) break;
		}
		std::vector<int> next(pushed.rbegin(), pushed.rend());
		for (int id : order) if (!gone[id]) next.push_back(id);
		order.swap(next);
	};

	const size_t never = (size_t)-1;
	bool finish = false;
	while (!finish)
	{
		// Phase 1 pass (:588-631): every divisible node currently in the list, front to back
		const size_t prev = order.size();
		std::vector<int> todo;
		for (int id : order) if (nodes[id].cnt > 1) todo.push_back(id);
		run_pass(todo, never);
		if (order.size() >= (size_t)quota || order.size() == prev) break;

		if (order.size() + 3 * divisibles.size() > (size_t)quota)
		{
			// Phase 2 (:635-672): largest first, libstdc++ tie order, stop at the quota
			while (!finish)
			{
				const size_t prev2 = order.size();
				std::vector<SortItem> items;
				for (int id : divisibles) items.push_back({ nodes[id].cnt, id });
				libstdcxx_sort_desc(items);
				std::vector<int> todo2;
				for (const SortItem& it : items) todo2.push_back(it.node);
				run_pass(todo2, (size_t)quota);
				if (order.size() >= (size_t)quota || order.size() == prev2) finish = true;
			}
		}
	}

	// best response per node, first wins ties, list order (:677-692)
	out.clear();
	for (int id : order)
	{
		const Node& nd = nodes[id];
		int best = -1, best_resp = 0;
		for (int k = 0; k < nd.cnt; k++)
		{
			const int i = perm[nd.beg + k];
			if (cand[i].response > best_resp) { best_resp = cand[i].response; best = i; }
		}
		out.push_back(cand[best]);
	}
}

// ---------------------------------------------------------------------------------------------
// E4  IC_Angle (:74-101): integer moments over the disc, then fastAtan2
// ---------------------------------------------------------------------------------------------
float ic_angle(const uint8_t* img, size_t pitch, int x, int y)
{
	const uint8_t* c = img + (size_t)y * pitch + x;
	int m01 = 0, m10 = 0;
	for (int u = -kHalfPatch; u <= kHalfPatch; u++) m10 += u * c[u];
	for (int v = 1; v <= kHalfPatch; v++)
	{
		const int d = kUMax.v[v];
		int vsum = 0;
		const uint8_t* lo = c + (ptrdiff_t)v * (ptrdiff_t)pitch;
		const uint8_t* hi = c - (ptrdiff_t)v * (ptrdiff_t)pitch;
		for (int u = -d; u <= d; u++)
		{
			const int a = lo[u], b = hi[u];
			vsum += a - b;
			m10 += u * (a + b);
		}
		m01 += v * vsum;
	}
	return cvp::fast_atan2_deg((float)m01, (float)m10);
}

// ---------------------------------------------------------------------------------------------
// E6  ComputeOrbDescriptor (:103-140). Float semantics pinned per SURVEY H2: cos/sin evaluated in
// double on the float32 angle and rounded to float32; products and sums in float32, no contraction;
// round-half-even.
// ---------------------------------------------------------------------------------------------
void descriptor(const uint8_t* blurred, size_t pitch, int x, int y, float angle_deg, uint8_t* desc)
{
	const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
	const float angle = angle_deg * factorPI;
	const float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
	const uint8_t* c = blurred + (size_t)y * pitch + x;
	auto sample = [&](int i) -> int {
		const float px = (float)kPattern[2 * i], py = (float)kPattern[2 * i + 1];
		const float fr = px * b + py * a;    // two rounded products, one rounded sum (-ffp-contract=off)
		const float fc = px * a - py * b;
		return c[(ptrdiff_t)cvp::round_rne(fr) * (ptrdiff_t)pitch + cvp::round_rne(fc)];
	};
	for (int byte = 0; byte < 32; byte++)
	{
		int v = 0;
		for (int bit = 0; bit < 8; bit++)
		{
			const int pair = byte * 8 + bit;
			v |= (sample(2 * pair) < sample(2 * pair + 1)) << bit;
		}
		desc[byte] = (uint8_t)v;
	}
}

// ---------------------------------------------------------------------------------------------
// E0/E1/E7  ORBextractor: ctor/Init tables (:695-741), ComputePyramid (:455-470), Extract (:743-820)
// ---------------------------------------------------------------------------------------------
void quotas(int total, float scaleFactor, int nlevels, int* out)
{
	// ComputeNumFeaturesPerScale (:472-487), double arithmetic
	const double factor = 1 / scaleFactor;
	double nf = total * (1 - factor) / (1 - std::pow(factor, nlevels));
	int sum = 0;
	for (int s = 0; s < nlevels - 1; s++)
	{
		out[s] = cvp::round_rne(nf);
		sum += out[s];
		nf *= factor;
	}
	out[nlevels - 1] = std::max(total - sum, 0);
}

struct Extractor
{
	int nfeatures, nlevels, iniTh, minTh;
	float scaleFactor;
	std::vector<float> scale, inv, sig, invsig;
	std::vector<int> quota;
	std::vector<Image> pyr;

	Extractor(int nf, float sf, int nl, int ini, int mn) : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf)
	{
		scale.resize(nl); inv.resize(nl); sig.resize(nl); invsig.resize(nl); quota.resize(nl);
		float s = 1.f;
		for (int i = 0; i < nl; i++)
		{
			scale[i] = s; inv[i] = 1.f / s; sig[i] = s * s; invsig[i] = 1.f / (s * s);
			s *= sf;
		}
		quotas(nf, sf, nl, quota.data());
	}

	int extract(const uint8_t* img, int w, int h, size_t pitch, oracle_keypoint* kps, uint8_t* desc, int cap)
	{
		// input contract: every level must keep a >= 30 px roi and the image must be landscape
		// enough for cvRound(w/h) >= 1; the reference divides by zero otherwise
		pyr.resize(nlevels);
		pyr[0].alloc(w, h);
		for (int y = 0; y < h; y++) std::memcpy(&pyr[0].px[(size_t)y * w], img + (size_t)y * pitch, (size_t)w);
		for (int s = 1; s < nlevels; s++)
		{
			const int lh = cvp::round_rne(inv[s] * (float)h), lw = cvp::round_rne(inv[s] * (float)w);
			pyr[s].alloc(lw, lh);
			cvp::resize_linear_u8(pyr[s - 1].px.data(), pyr[s - 1].w, pyr[s - 1].h, (size_t)pyr[s - 1].w,
			                      pyr[s].px.data(), lw, lh, (size_t)lw);
		}
		for (int s = 0; s < nlevels; s++)
		{
			const int rw = pyr[s].w - 2 * kBorder, rh = pyr[s].h - 2 * kBorder;
			if (rw < kCell || rh < kCell || cvp::round_rne(1. * rw / rh) < 1) return -1;
		}

		std::vector<std::vector<oracle_cand>> sel(nlevels);
		std::vector<std::vector<float>> ang(nlevels);
		std::vector<oracle_cand> cand;
		int total = 0;
		for (int s = 0; s < nlevels; s++)
		{
			const Image& L = pyr[s];
			detect_fast(L.px.data(), L.w, L.h, (size_t)L.w, iniTh, minTh, cand);
			quadtree(cand, L.w, L.h, quota[s], sel[s]);
			for (const oracle_cand& k : sel[s]) ang[s].push_back(ic_angle(L.px.data(), (size_t)L.w, k.x, k.y));
			total += (int)sel[s].size();
		}
		if (total == 0) return 0;          // :778-782 (the caller's keypoint vector is left as it was)
		if (total > cap) return -total;

		Image blur;
		int at = 0;
		for (int s = 0; s < nlevels; s++)
		{
			if (sel[s].empty()) continue;
			const Image& L = pyr[s];
			blur.alloc(L.w, L.h);
			cvp::gauss7x7_u8(L.px.data(), L.w, L.h, (size_t)L.w, blur.px.data(), (size_t)L.w);
			for (size_t i = 0; i < sel[s].size(); i++, at++)
			{
				const oracle_cand& k = sel[s][i];
				descriptor(blur.px.data(), (size_t)L.w, k.x, k.y, ang[s][i], desc + 32 * (size_t)at);
				oracle_keypoint& o = kps[at];
				o.x = (float)k.x; o.y = (float)k.y;
				if (s > 0) { o.x *= scale[s]; o.y *= scale[s]; }   // :811-815, after the descriptor
				o.size = scale[s] * kPatch;                        // :771
				o.angle = ang[s][i];
				o.response = (float)k.response;
				o.octave = s;
				o.class_id = -1;
			}
		}
		return total;
	}
};

// ---------------------------------------------------------------------------------------------
// M1  DescriptorDistance (src/ORBmatcher.cc:1449-1457)
// ---------------------------------------------------------------------------------------------
inline int hamming256(const uint8_t* a, const uint8_t* b)
{
	uint32_t wa[8], wb[8];
	std::memcpy(wa, a, 32);
	std::memcpy(wb, b, 32);
	int d = 0;
	for (int i = 0; i < 8; i++) d += __builtin_popcount(wa[i] ^ wb[i]);
	return d;
}

// ---------------------------------------------------------------------------------------------
// M3/M4  PatchDistance (:60-68) and ComputeStereoMatches (:72-247)
// ---------------------------------------------------------------------------------------------
int stereo(const oracle_keypoint* kpL, int nL, const uint8_t* descL, const uint8_t* const* pyrL,
           const oracle_keypoint* kpR, int nR, const uint8_t* descR, const uint8_t* const* pyrR,
           const int* lw, const int* /*lh*/, const size_t* lp, int /*nlevels*/, const float* scale, const float* inv,
           const oracle_camera* cam, float* uright, float* depth)
{
	const int TH_HIGH = 100, TH_LOW = 50, R = 5, PS = 11, SR = 5;   // :41-47
	for (int i = 0; i < nL; i++) uright[i] = depth[i] = -1.f;
	if (nL == 0) return -1;

	// band of rows each right keypoint is filed under (:89-99)
	std::vector<int> rmin(nR), rmax(nR);
	for (int i = 0; i < nR; i++)
	{
		const float r = 2.f * scale[kpR[i].octave];
		rmin[i] = (int)std::floor(kpR[i].y - r);
		rmax[i] = (int)std::ceil(kpR[i].y + r);
	}
	const float minZ = cam->baseline, mind = 0.f, maxd = cam->bf / minZ;   // :102-104
	const int TH_ORB = (TH_HIGH + TH_LOW) / 2;
	const float eps = 0.01f;

	std::vector<std::pair<int, int>> kept;   // (sad, iL)
	for (int iL = 0; iL < nL; iL++)
	{
		const oracle_keypoint& kl = kpL[iL];
		const int row = (int)kl.y;              // :122 truncation
		const float minu = kl.x - maxd, maxu = kl.x - mind;
		if (maxu < 0) continue;

		// the row list holds right indices in ascending order, so "first strict minimum" = lowest index
		int bestDist = TH_HIGH, bestR = 0;
		bool any = false;
		for (int iR = 0; iR < nR; iR++)
		{
			if (row < rmin[iR] || row > rmax[iR]) continue;
			any = true;
			const oracle_keypoint& kr = kpR[iR];
			if (kr.octave < kl.octave - 1 || kr.octave > kl.octave + 1) continue;
			if (!(kr.x >= minu && kr.x <= maxu)) continue;
			const int d = hamming256(descL + 32 * (size_t)iL, descR + 32 * (size_t)iR);
			if (d < bestDist) { bestDist = d; bestR = iR; }
		}
		if (!any || bestDist >= TH_ORB) continue;

		// 11x11 SAD over 11 horizontal shifts on the left keypoint's pyramid level (:163-197)
		const int o = kl.octave;
		const float sf = inv[o];
		const int suL = (int)std::round(sf * kl.x), svL = (int)std::round(sf * kl.y);
		const int suR = (int)std::round(sf * kpR[bestR].x);
		if (suR + SR - R < 0 || suR + SR + R + 1 >= lw[o]) continue;
		const uint8_t* IL = pyrL[o] + (size_t)(svL - R) * lp[o] + (suL - R);
		int dist[2 * 5 + 1];
		int bestSad = 0x7fffffff, bestDx = 0;
		for (int dx = -SR; dx <= SR; dx++)
		{
			const uint8_t* IR = pyrR[o] + (size_t)(svL - R) * lp[o] + (suR + dx - R);
			const int sub = IL[(size_t)R * lp[o] + R] - IR[(size_t)R * lp[o] + R];
			int sum = 0;
			for (int y = 0; y < PS; y++)
				for (int x = 0; x < PS; x++)
					sum += std::abs((int)IL[(size_t)y * lp[o] + x] - (int)IR[(size_t)y * lp[o] + x] - sub);
			if (sum < bestSad) { bestSad = sum; bestDx = dx; }
			dist[SR + dx] = sum;
		}
		if (bestDx == -SR || bestDx == SR) continue;

		// parabola through the three SADs around the minimum (:203-210); float32, unfused
		const int d1 = dist[SR + bestDx - 1], d2 = dist[SR + bestDx], d3 = dist[SR + bestDx + 1];
		const float deltaR = (float)(d1 - d3) / (2.f * ((float)(d1 + d3) - 2.f * (float)d2));
		if (deltaR < -1 || deltaR > 1) continue;

		float bestuR = scale[o] * ((float)(suR + bestDx) + deltaR);
		float disparity = kl.x - bestuR;
		if (disparity >= mind && disparity < maxd)
		{
			if (disparity <= 0) { disparity = eps; bestuR = kl.x - eps; }
			depth[iL] = cam->bf / disparity;
			uright[iL] = bestuR;
			kept.push_back({ bestSad, iL });
		}
	}

	// outlier cut (:231-246): descending by (sad, iL); median at size/2-1; drop everything >= 2.1*median
	if (kept.empty()) return 0;
	std::sort(kept.begin(), kept.end(), std::greater<std::pair<int, int>>());
	const int m = std::max((int)kept.size() / 2 - 1, 0);
	const float th = 1.5f * 1.4f * kept[m].first;
	for (const auto& k : kept)
	{
		if (k.first < th) break;
		uright[k.second] = -1;
		depth[k.second] = -1;
	}
	return 0;
}

}  // namespace

extern "C" {

void* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
{
	return new Extractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
void orc_extractor_destroy(void* ex) { delete static_cast<Extractor*>(ex); }

int orc_extractor_extract(void* ex, const uint8_t* img, int w, int h, size_t pitch, oracle_keypoint* kps, uint8_t* desc, int cap)
{
	return static_cast<Extractor*>(ex)->extract(img, w, h, pitch, kps, desc, cap);
}
int orc_extractor_level_size(void* ex, int level, int* w, int* h)
{
	Extractor* e = static_cast<Extractor*>(ex);
	if (level < 0 || level >= (int)e->pyr.size()) return -1;
	*w = e->pyr[level].w; *h = e->pyr[level].h;
	return 0;
}
int orc_extractor_level_copy(void* ex, int level, uint8_t* dst, size_t pitch)
{
	Extractor* e = static_cast<Extractor*>(ex);
	if (level < 0 || level >= (int)e->pyr.size()) return -1;
	const Image& L = e->pyr[level];
	for (int y = 0; y < L.h; y++) std::memcpy(dst + (size_t)y * pitch, L.row(y), (size_t)L.w);
	return 0;
}
void orc_extractor_tables(void* ex, float* scale, float* inv_scale, float* sigma_sq, float* inv_sigma_sq)
{
	Extractor* e = static_cast<Extractor*>(ex);
	std::copy(e->scale.begin(), e->scale.end(), scale);
	std::copy(e->inv.begin(), e->inv.end(), inv_scale);
	std::copy(e->sig.begin(), e->sig.end(), sigma_sq);
	std::copy(e->invsig.begin(), e->invsig.end(), inv_sigma_sq);
}
void orc_feature_quotas(int nfeatures, float scaleFactor, int nlevels, int* out) { quotas(nfeatures, scaleFactor, nlevels, out); }

int orc_detect_fast(const uint8_t* img, int w, int h, size_t pitch, int iniTh, int minTh, oracle_cand* out, int cap)
{
	std::vector<oracle_cand> v;
	detect_fast(img, w, h, pitch, iniTh, minTh, v);
	if ((int)v.size() > cap) return -(int)v.size();
	std::copy(v.begin(), v.end(), out);
	return (int)v.size();
}
int orc_quadtree(const oracle_cand* in, int n, int w, int h, int nfeatures, oracle_cand* out, int cap)
{
	std::vector<oracle_cand> c(in, in + n), o;
	quadtree(c, w, h, nfeatures, o);
	if ((int)o.size() > cap) return -(int)o.size();
	std::copy(o.begin(), o.end(), out);
	return (int)o.size();
}
float orc_ic_angle(const uint8_t* img, int /*w*/, int /*h*/, size_t pitch, int x, int y) { return ic_angle(img, pitch, x, y); }
void orc_descriptor(const uint8_t* blurred, int /*w*/, int /*h*/, size_t pitch, int x, int y, float angle_deg, uint8_t* desc32)
{
	descriptor(blurred, pitch, x, y, angle_deg, desc32);
}
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }

int orc_stereo_matches(const oracle_keypoint* kpL, int nL, const uint8_t* descL, const uint8_t* const* pyrL,
                       const oracle_keypoint* kpR, int nR, const uint8_t* descR, const uint8_t* const* pyrR,
                       const int* level_w, const int* level_h, const size_t* level_pitch, int nlevels,
                       const float* scale, const float* inv_scale, const oracle_camera* cam, float* uright, float* depth)
{
	return stereo(kpL, nL, descL, pyrL, kpR, nR, descR, pyrR, level_w, level_h, level_pitch, nlevels, scale, inv_scale,
	              cam, uright, depth);
}

// best/second scan of SearchByBoW (src/ORBmatcher.cc:477-507)
void orc_knn2(const uint8_t* query, int64_t nq, const uint8_t* train, int64_t nt, int th_low, float nnratio,
              int32_t* idx, uint16_t* best, uint16_t* second, int32_t* match, int threads)
{
	auto work = [&](int64_t q0, int64_t q1) {
		for (int64_t q = q0; q < q1; q++)
		{
			int b = 256, s = 256, bi = -1;
			const uint8_t* dq = query + 32 * q;
			for (int64_t t = 0; t < nt; t++)
			{
				const int d = hamming256(dq, train + 32 * t);
				if (d < b) { s = b; b = d; bi = (int)t; }
				else if (d < s) s = d;
			}
			idx[q] = bi; best[q] = (uint16_t)b; second[q] = (uint16_t)s;
			if (match) match[q] = (b <= th_low && (float)b < nnratio * (float)s) ? bi : -1;
		}
	};
	if (threads <= 1) { work(0, nq); return; }
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; t++) pool.emplace_back(work, nq * t / threads, nq * (t + 1) / threads);
	for (auto& t : pool) t.join();
}

// the rotation terms of ComputeOrbDescriptor (src/ORBextractor.cc:105-107) for the n float angles whose bit patterns are first_bits,
// first_bits + 1, ...: factorPI = (float)(CV_PI / 180.f); float angle = kpt.angle * factorPI; a = (float)cos(angle), b = (float)sin(angle)
// (the double overloads, glibc). Lets a test sweep EVERY float angle in [0, 360) against the device.
void orc_cos_sin_range(uint32_t first_bits, int64_t n, float* c, float* s, int threads)
{
	auto work = [&](int64_t i0, int64_t i1) {
		const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
		for (int64_t i = i0; i < i1; i++)
		{
			const uint32_t bits = first_bits + (uint32_t)i;
			float kp_angle;
			std::memcpy(&kp_angle, &bits, 4);
			float angle = (float)kp_angle * factorPI;
			c[i] = (float)cos(angle); s[i] = (float)sin(angle);
		}
	};
	if (threads <= 1) { work(0, n); return; }
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; t++) pool.emplace_back(work, n * t / threads, n * (t + 1) / threads);
	for (auto& t : pool) t.join();
}

// ---- rows "next" of SURVEY §8(f) ----
// ConvertToGray (src/System.cc:122-137) = cv::cvtColor with the code picked from (channels, RGB flag)
void orc_convert_to_gray(const uint8_t* src, int w, int h, size_t pitch, int channels, int rgb, uint8_t* dst, size_t dst_pitch)
{
	cvp::cvt_gray_u8(src, w, h, pitch, channels, rgb != 0, dst, dst_pitch);
}

// ComputeStereoFromRGBD (src/System.cc:197-219)
void orc_stereo_from_rgbd(const oracle_keypoint* kps, const oracle_keypoint* kps_un, int n, const float* depth_map, int /*w*/, int /*h*/,
                          size_t pitch, const oracle_camera* cam, float* uright, float* depth)
{
	for (int i = 0; i < n; i++)
	{
		uright[i] = depth[i] = -1.f;
		const int v = (int)kps[i].y, u = (int)kps[i].x;      // truncation, :210-211
		const float d = *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(depth_map) + (size_t)v * pitch + 4 * (size_t)u);
		if (d > 0)
		{
			depth[i] = d;
			uright[i] = kps_un[i].x - cam->bf / d;
		}
	}
}

// distance matrix + least median of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:286-314)
int orc_distinctive_index(const uint8_t* desc, int n)
{
	if (n <= 0) return -1;
	int best_median = 0x7fffffff, best = 0;
	std::vector<int> row(n);
	for (int i = 0; i < n; i++)
	{
		for (int j = 0; j < n; j++) row[j] = i == j ? 0 : hamming256(desc + 32 * (size_t)i, desc + 32 * (size_t)j);
		std::nth_element(row.begin(), row.begin() + (n - 1) / 2, row.end());     // value at sorted position (N-1)/2
		const int median = row[(n - 1) / 2];
		if (median < best_median) { best_median = median; best = i; }
	}
	return best;
}

void orc_cv_resize(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep)
{
	cvp::resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}
int orc_cv_fast(const uint8_t* img, int w, int h, size_t step, int th, int nms, oracle_cand* out, int cap)
{
	std::vector<cvp::FastPoint> pts;
	cvp::fast9_16(img, w, h, step, th, nms != 0, pts);
	if ((int)pts.size() > cap) return -(int)pts.size();
	for (size_t i = 0; i < pts.size(); i++) out[i] = { pts[i].x, pts[i].y, pts[i].score };
	return (int)pts.size();
}
void orc_cv_gaussian7(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep)
{
	cvp::gauss7x7_u8(src, w, h, sstep, dst, dstep);
}
void orc_undistort_keypoints(const oracle_keypoint* kps, int n, const oracle_camera* cam, const float* dist, int ndist, oracle_keypoint* kps_un)
{
	// UndistortKeyPoints (src/System.cc:153-174): untouched when distCoeffs(0) == 0, else every pt through cv::undistortPoints(.., K, dist, noArray(), K)
	for (int i = 0; i < n; i++) kps_un[i] = kps[i];
	if (ndist < 1 || dist[0] == 0.f) return;
	std::vector<float> in((size_t)n * 2), out((size_t)n * 2);
	for (int i = 0; i < n; i++) { in[2 * i] = kps[i].x; in[2 * i + 1] = kps[i].y; }
	cvp::undistort_points(in.data(), n, cam->fx, cam->fy, cam->cx, cam->cy, dist, ndist, out.data());
	for (int i = 0; i < n; i++) { kps_un[i].x = out[2 * i]; kps_un[i].y = out[2 * i + 1]; }
}
void orc_cv_remap(const uint8_t* src, int sw, int sh, size_t sstep, const float* mapx, const float* mapy, size_t mstep, uint8_t* dst, int dw, int dh,
                  size_t dstep)
{
	cvp::remap_linear_u8(src, sw, sh, sstep, mapx, mapy, mstep, dst, dw, dh, dstep);
}
float orc_cv_fast_atan2(float y, float x) { return cvp::fast_atan2_deg(y, x); }
int orc_cv_round_f(float v) { return cvp::round_rne(v); }
int orc_cv_round_d(double v) { return cvp::round_rne(v); }

}  // extern "C"
