// TEST INFRASTRUCTURE — CPU oracle, not the product.
//
// Stand-alone restatement of the guided matchers in the array / pass form the CUDA path uses, so that it is at the
// same time the executable specification of orbx_guided.cu. Checked against the reference text itself (oracle/_ref, rule guided_gen.cc)
// by tests/test_oracle_vs_ref.py; travels to machines without /root/reference.
//
//   FeaturesGrid::AssignFeatures / GetFeaturesInArea          src/Frame.cc:70-145
//   CheckOrientation                                          src/ORBmatcher.cc:249-309
//   ORBmatcher::SearchByProjection(Frame&, mappoints, th)     src/ORBmatcher.cc:315-382
//   ORBmatcher::SearchForInitialization                       src/ORBmatcher.cc:614-694
//   ORBmatcher::SearchByProjection(currFrame, lastFrame, ..)  src/ORBmatcher.cc:1279-1362
//
// The two SearchByProjection loops are sequential in the reference: a keypoint taken by an earlier map point with observations is
// skipped by every later one. Here they are restated as a FIXPOINT: choice[i] = best candidate of point i among the keypoints not
// taken by any point j < i; iterate from "nothing taken" until no choice changes. By induction on i the unique fixpoint is the
// sequential result (choice[0] never depends on anyone; choice[i] only on choices of j < i), and the GPU runs exactly these rounds.
#define ORACLE_PREFIX orc_
#include "oracle_api.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

#include "stdsort_replay.h"

namespace {

int g_stat[2];   // [0] rounds of the last run_rounds, [1] matches revoked by the last search_for_initialization (test coverage only)

const int COLS = 64, ROWS = 48;        // include/Frame.h:72-73
const int TH_HIGH = 100, TH_LOW = 50;  // src/ORBmatcher.cc:41-42
const int HISTO_LENGTH = 30;           // :43

inline int hamming256(const uint8_t* a, const uint8_t* b)   // DescriptorDistance, :1449-1457
{
	int d = 0;
	for (int i = 0; i < 8; i++)
	{
		uint32_t x, y;
		memcpy(&x, a + 4 * i, 4);
		memcpy(&y, b + 4 * i, 4);
		d += __builtin_popcount(x ^ y);
	}
	return d;
}

// grid_[cx][cy] as one CSR array: cell = cx * ROWS + cy, items ascending (the push_back order of :87-99)
struct Grid
{
	std::vector<oracle_keypoint> kps;
	oracle_bounds b;
	int nlevels;
	float invW, invH;
	std::vector<int> start, items;

	void assign(const oracle_keypoint* k, int n, const oracle_bounds& bounds, int nl)
	{
		kps.assign(k, k + n);
		b = bounds;
		nlevels = nl;
		invW = COLS / (b.maxx - b.minx);   // :73-74, int / float
		invH = ROWS / (b.maxy - b.miny);
		std::vector<int> cell((size_t)n);
		start.assign(COLS * ROWS + 1, 0);
		for (int i = 0; i < n; i++)
		{
			const int cx = (int)std::round(invW * (k[i].x - b.minx));   // :91-92, half away from zero
			const int cy = (int)std::round(invH * (k[i].y - b.miny));
			cell[i] = (cx < 0 || cx >= COLS || cy < 0 || cy >= ROWS) ? -1 : cx * ROWS + cy;
			if (cell[i] >= 0) start[cell[i] + 1]++;
		}
		for (int c = 0; c < COLS * ROWS; c++) start[c + 1] += start[c];
		items.assign((size_t)start[COLS * ROWS], 0);
		std::vector<int> fill(start.begin(), start.end() - 1);
		for (int i = 0; i < n; i++)
			if (cell[i] >= 0) items[fill[cell[i]]++] = i;
	}

	// GetFeaturesInArea (:102-145): calls f(idx) in the reference's output order
	template <class F> void query(float x, float y, float r, int minLevel, int maxLevel, F f) const
	{
		const int mincx = std::max((int)std::floor(invW * (x - r - b.minx)), 0);
		const int maxcx = std::min((int)std::ceil(invW * (x + r - b.minx)), COLS - 1);
		const int mincy = std::max((int)std::floor(invH * (y - r - b.miny)), 0);
		const int maxcy = std::min((int)std::ceil(invH * (y + r - b.miny)), ROWS - 1);
		if (mincx >= COLS || maxcx < 0 || mincy >= ROWS || maxcy < 0) return;
		const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
		if (maxLevel < 0) maxLevel = nlevels;
		for (int cx = mincx; cx <= maxcx; cx++)
			for (int cy = mincy; cy <= maxcy; cy++)
				for (int p = start[cx * ROWS + cy]; p < start[cx * ROWS + cy + 1]; p++)
				{
					const int idx = items[p];
					const oracle_keypoint& kp = kps[idx];
					if (checkLevels && (kp.octave < minLevel || kp.octave > maxLevel)) continue;
					if (std::fabs(kp.x - x) < r && std::fabs(kp.y - y) < r) f(idx);
				}
	}
};

Grid make_grid(const oracle_frame_view* v)
{
	Grid g;
	g.assign(v->kps_un, v->n, v->bounds, v->nlevels);
	return g;
}

// CheckOrientation (:249-309). matches are (index into angle1, index into angle2 == index into status), in push order.
// erased[k] is set for every match whose status entry the reference invalidates; returns matches - reduction.
int check_orientation(const std::vector<std::pair<int, int>>& matches, const float* angle1, size_t stride1, const float* angle2, size_t stride2,
                      std::vector<int>& erased_i2)
{
	const float factor = 1.f / HISTO_LENGTH;
	std::vector<int> hist[HISTO_LENGTH];
	for (const auto& m : matches)
	{
		float diff = *(const float*)((const char*)angle1 + stride1 * m.first) - *(const float*)((const char*)angle2 + stride2 * m.second);
		if (diff < 0) diff += 360;
		int bin = (int)lrintf(factor * diff);   // cvRound(float)
		if (bin == HISTO_LENGTH) bin = 0;
		hist[bin].push_back(m.second);
	}
	// std::sort of the 30 bins by size, descending: unstable, so replayed (stdsort_replay.h)
	std::vector<stdsort::SortItem> order(HISTO_LENGTH);
	for (int i = 0; i < HISTO_LENGTH; i++) order[i] = { (int)hist[i].size(), i };
	stdsort::libstdcxx_sort_desc(order);
	const size_t max1 = (size_t)order[0].size, max2 = (size_t)order[1].size, max3 = (size_t)order[2].size;
	int eraseBin = 3;
	if (max2 < 0.1 * max1) eraseBin = 1;
	else if (max3 < 0.1 * max1) eraseBin = 2;
	int reduction = 0;
	erased_i2.clear();
	for (int b = eraseBin; b < HISTO_LENGTH; b++)
		for (int i2 : hist[order[b].node])
		{
			erased_i2.push_back(i2);
			reduction++;
		}
	return (int)matches.size() - reduction;
}

// what one point asks of the frame: a window, a level range, the right-image coordinate, its descriptor
struct Probe
{
	bool active;
	float u, v, ur, radius;
	int minLevel, maxLevel;
	bool obs;              // Observations() > 0: the keypoint it takes is closed to later points
};

struct Pick { int idx, best, second, bestLevel, secondLevel; };

// best / second scan of :337-367 (and, ignoring the second-best fields, of :1327-1347) over the window of one probe.
// `taken(idx)` is the only state-dependent test.
template <class Taken> Pick scan_window(const Grid& g, const oracle_frame_view* f, const Probe& p, const uint8_t* desc, Taken taken)
{
	Pick k = { -1, 256, 256, -1, -1 };
	g.query(p.u, p.v, p.radius, p.minLevel, p.maxLevel, [&](int idx) {
		if (taken(idx)) return;
		const float ur2 = f->uright ? f->uright[idx] : -1.f;
		if (ur2 > 0 && std::fabs(p.ur - ur2) > p.radius) return;
		const int d = hamming256(desc, f->desc + (size_t)idx * 32);
		if (d < k.best)
		{
			k.second = k.best; k.best = d;
			k.secondLevel = k.bestLevel; k.bestLevel = g.kps[idx].octave;
			k.idx = idx;
		}
		else if (d < k.second)
		{
			k.secondLevel = g.kps[idx].octave;
			k.second = d;
		}
	});
	return k;
}

// The fixpoint described in the header. accept(pick) says whether point i takes pick.idx. On return choice[i] is the keypoint taken
// by point i (or -1) and frame_mp is updated to the reference's final frame.mappoints.
template <class Accept>
int run_rounds(const Grid& g, const oracle_frame_view* f, const std::vector<Probe>& probes, const uint8_t* pt_desc, int32_t* frame_mp,
               std::vector<int>& choice, Accept accept, bool any_map_point_closes = false)
{
	const int np = (int)probes.size(), n = f->n;
	choice.assign((size_t)np, -1);
	// keypoints closed from the start: frame.mappoints[idx] && frame.mappoints[idx]->Observations() > 0 on entry
	std::vector<char> closed((size_t)n);
	for (int c = 0; c < n; c++)
		closed[c] = any_map_point_closes ? frame_mp[c] != -1 : (frame_mp[c] == -2 || (frame_mp[c] >= 0 && probes[(size_t)frame_mp[c]].obs));
	std::vector<int> owner((size_t)n);   // lowest point index with observations that takes the keypoint; INT_MAX = nobody
	for (int round = 0;; round++)
	{
		for (int c = 0; c < n; c++) owner[c] = INT_MAX;
		for (int i = 0; i < np; i++)
			if (choice[i] >= 0 && probes[i].obs && i < owner[choice[i]]) owner[choice[i]] = i;
		bool changed = false;
		for (int i = 0; i < np; i++)
		{
			if (!probes[i].active) continue;
			const Pick k = scan_window(g, f, probes[i], pt_desc + (size_t)i * 32,
			                           [&](int idx) { return closed[idx] || owner[idx] < i; });
			const int c = accept(k) ? k.idx : -1;
			if (c != choice[i]) { choice[i] = c; changed = true; }
		}
		g_stat[0] = round + 1;
		if (!changed) break;
	}
	int nmatches = 0;
	for (int i = 0; i < np; i++)
		if (choice[i] >= 0) { frame_mp[choice[i]] = i; nmatches++; }   // ascending i: the last writer wins, as in the reference loop
	return nmatches;
}

}  // namespace

extern "C" {

// port only: coverage counters of the last call (see g_stat)
int orc_guided_stat(int which) { return g_stat[which & 1]; }

void* orc_grid_create(const oracle_keypoint* kps, int n, const oracle_bounds* b, int nlevels)
{
	Grid* g = new Grid;
	g->assign(kps, n, *b, nlevels);
	return g;
}

void orc_grid_destroy(void* g) { delete static_cast<Grid*>(g); }

int orc_grid_query(void* g, float x, float y, float r, int min_level, int max_level, int32_t* out, int cap)
{
	int n = 0;
	static_cast<Grid*>(g)->query(x, y, r, min_level, max_level, [&](int idx) { if (n < cap) out[n] = idx; n++; });
	return n;
}

static int search_local_map(const Grid& g, const oracle_frame_view* f, int32_t* frame_mp, const oracle_track_point* pts, const uint8_t* pt_desc,
                            int npts, float th, float nnratio)
{
	std::vector<Probe> probes((size_t)npts);
	for (int i = 0; i < npts; i++)
	{
		Probe& p = probes[i];
		p.active = (pts[i].flags & 1) != 0;
		p.obs = (pts[i].flags & 2) != 0;
		if (!p.active) continue;
		const int ps = pts[i].scale_level;
		const float r = pts[i].view_cos > 0.998 ? 2.5f : 4.f;   // RadiusByViewingCos (:53): float against a double literal
		p.radius = th * r * f->scale_factors[ps];             // :326-327
		p.u = pts[i].proj_x; p.v = pts[i].proj_y; p.ur = pts[i].proj_xr;
		p.minLevel = ps - 1; p.maxLevel = ps;                   // :332
	}
	std::vector<int> choice;
	return run_rounds(g, f, probes, pt_desc, frame_mp, choice, [&](const Pick& k) {
		if (!(k.best <= TH_HIGH)) return false;                                               // :370
		if (k.bestLevel == k.secondLevel && k.best > nnratio * k.second) return false;        // :372-373
		return true;
	});
}

static int search_last_frame(const Grid& g, const oracle_frame_view* f, const oracle_camera* cam, const oracle_pose* cp, const oracle_pose* lp,
                             int32_t* frame_mp, const oracle_last_point* pts, const uint8_t* pt_desc, int npts, float th, int monocular,
                             int check_ori)
{
	// tlc = Rlw * (-Rcw^T * tcw) + tlw (:1286); every product accumulates from 0 in k order like cv::Matx
	float twc[3], tlc[3];
	for (int i = 0; i < 3; i++)
	{
		float s = 0;
		for (int k = 0; k < 3; k++) s += (cp->R[k * 3 + i] * -1) * cp->t[k];
		twc[i] = s;
	}
	for (int i = 0; i < 3; i++)
	{
		float s = 0;
		for (int k = 0; k < 3; k++) s += lp->R[i * 3 + k] * twc[k];
		tlc[i] = s + lp->t[i];
	}
	const bool forward = tlc[2] > cam->baseline && !monocular;    // :1287-1288
	const bool backward = -tlc[2] > cam->baseline && !monocular;

	std::vector<Probe> probes((size_t)npts);
	for (int i = 0; i < npts; i++)
	{
		Probe& p = probes[i];
		p.active = false;
		p.obs = (pts[i].flags & 2) != 0;
		if (!(pts[i].flags & 1)) continue;                         // :1296-1297
		float xc[3];
		for (int r = 0; r < 3; r++)
		{
			float s = 0;
			for (int k = 0; k < 3; k++) s += cp->R[r * 3 + k] * pts[i].xw[k];
			xc[r] = s + cp->t[r];
		}
		if (xc[2] < 0.f) continue;                                 // :1302-1303
		const float invZ = 1.f / xc[2];                            // CameraProjection::CameraToImage
		p.u = invZ * cam->fx * xc[0] + cam->cx;
		p.v = invZ * cam->fy * xc[1] + cam->cy;
		p.ur = p.u - cam->bf / xc[2];                              // :1308
		if (!(p.u >= f->bounds.minx && p.u < f->bounds.maxx && p.v >= f->bounds.miny && p.v < f->bounds.maxy)) continue;   // :1310
		const int oct = pts[i].octave;
		p.radius = th * f->scale_factors[oct];                     // :1316
		p.minLevel = forward ? oct : (backward ? 0 : oct - 1);     // :1318-1319
		p.maxLevel = forward ? -1 : (backward ? oct : oct + 1);
		p.active = true;
	}
	std::vector<int> choice;
	const int nmatches = run_rounds(g, f, probes, pt_desc, frame_mp, choice, [&](const Pick& k) { return k.best <= TH_HIGH; });
	if (!check_ori) return nmatches;
	std::vector<std::pair<int, int>> matches;
	for (int i = 0; i < npts; i++)
		if (choice[i] >= 0) matches.push_back({ i, choice[i] });
	std::vector<int> erased;
	const int kept = check_orientation(matches, &pts[0].angle, sizeof(oracle_last_point), &f->kps_un[0].angle, sizeof(oracle_keypoint), erased);
	for (int i2 : erased) frame_mp[i2] = -1;
	return kept;
}

int orc_search_local_map(const oracle_frame_view* f, int32_t* frame_mp, const oracle_track_point* pts, const uint8_t* pt_desc, int npts, float th,
                         float nnratio)
{
	return search_local_map(make_grid(f), f, frame_mp, pts, pt_desc, npts, th, nnratio);
}

int orc_search_last_frame(const oracle_frame_view* f, const oracle_camera* cam, const oracle_pose* cp, const oracle_pose* lp, int32_t* frame_mp,
                          const oracle_last_point* pts, const uint8_t* pt_desc, int npts, float th, int monocular, float nnratio,
                          int check_ori)
{
	(void)nnratio;
	return search_last_frame(make_grid(f), f, cam, cp, lp, frame_mp, pts, pt_desc, npts, th, monocular, check_ori);
}

// timing of the restatement (the grid is built once, outside the timed region); see oracle_api.h
double orc_time_search_local_map(const oracle_frame_view* f, const int32_t* frame_mp, const oracle_track_point* pts, const uint8_t* pt_desc,
                                 int npts, float th, float nnratio, int reps)
{
	const Grid g = make_grid(f);
	std::vector<int32_t> mp((size_t)f->n);
	double total = 0;
	for (int r = 0; r < reps; r++)
	{
		std::copy(frame_mp, frame_mp + f->n, mp.begin());
		const auto t0 = std::chrono::steady_clock::now();
		search_local_map(g, f, mp.data(), pts, pt_desc, npts, th, nnratio);
		total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	}
	return total / (reps > 0 ? reps : 1);
}

double orc_time_search_last_frame(const oracle_frame_view* f, const oracle_camera* cam, const oracle_pose* cp, const oracle_pose* lp,
                                  const int32_t* frame_mp, const oracle_last_point* pts, const uint8_t* pt_desc, int npts, float th, int monocular,
                                  float nnratio, int check_ori, int reps)
{
	(void)nnratio;
	const Grid g = make_grid(f);
	std::vector<int32_t> mp((size_t)f->n);
	double total = 0;
	for (int r = 0; r < reps; r++)
	{
		std::copy(frame_mp, frame_mp + f->n, mp.begin());
		const auto t0 = std::chrono::steady_clock::now();
		search_last_frame(g, f, cam, cp, lp, mp.data(), pts, pt_desc, npts, th, monocular, check_ori);
		total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	}
	return total / (reps > 0 ? reps : 1);
}

// SearchByProjection(Frame&, KeyFrame*, alreadyFound, th, ORBdist) (:1364-1447, relocalisation). The geometry per map point is host
// arithmetic in the reference's operation order; the window search runs through the same rounds as the other two.
int orc_search_keyframe_projection(const oracle_frame_view* f, const oracle_camera* cam, const oracle_pose* pose, float log_scale_factor,
                                   int32_t* frame_mp, const oracle_kf_point* pts, const uint8_t* pt_desc, int npts, float th, int orb_dist,
                                   int check_ori)
{
	const Grid g = make_grid(f);
	float Ow[3];                                   // frame.GetCameraCenter() = pose.Invt() = -R^T * t (src/Frame.cc:203-206)
	for (int i = 0; i < 3; i++)
	{
		float s = 0;
		for (int k = 0; k < 3; k++) s += (pose->R[k * 3 + i] * -1) * pose->t[k];
		Ow[i] = s;
	}
	std::vector<Probe> probes((size_t)npts);
	for (int i = 0; i < npts; i++)
	{
		Probe& p = probes[i];
		p.active = false;
		p.obs = true;                              // any stored map point closes its keypoint (:1412-1413)
		if (!(pts[i].flags & 1)) continue;         // :1379-1381
		float xc[3];
		for (int r = 0; r < 3; r++)
		{
			float s = 0;
			for (int k = 0; k < 3; k++) s += pose->R[r * 3 + k] * pts[i].xw[k];
			xc[r] = s + pose->t[r];
		}
		const float invZ = 1.f / xc[2];            // WorldToImage: no depth test in this variant (:1385)
		p.u = invZ * cam->fx * xc[0] + cam->cx;
		p.v = invZ * cam->fy * xc[1] + cam->cy;
		p.ur = NAN;                                // no stereo gate
		if (!(p.u >= f->bounds.minx && p.u < f->bounds.maxx && p.v >= f->bounds.miny && p.v < f->bounds.maxy)) continue;   // :1389
		double ss = 0;                             // cv::norm(PO): squares accumulated in double
		for (int k = 0; k < 3; k++) { const float d = pts[i].xw[k] - Ow[k]; ss += (double)d * (double)d; }
		const float dist3D = (float)std::sqrt(ss);
		const float maxDistance = 1.2f * pts[i].max_distance, minDistance = 0.8f * pts[i].min_distance;   // src/MapPoint.cc:382-392
		if (dist3D < minDistance || dist3D > maxDistance) continue;                                       // :1400-1401
		const float ratio = pts[i].max_distance / dist3D;                                                 // PredictScale, src/MapPoint.cc:405-414
		const int scale = (int)std::ceil(std::log((double)ratio) / log_scale_factor);
		const int ps = std::max(0, std::min(scale, f->nlevels - 1));
		p.radius = th * f->scale_factors[ps];      // :1406
		p.minLevel = ps - 1; p.maxLevel = ps + 1;  // :1408
		p.active = true;
	}
	std::vector<int> choice;
	const int nmatches = run_rounds(g, f, probes, pt_desc, frame_mp, choice, [&](const Pick& k) { return k.idx >= 0 && k.best <= orb_dist; }, true);
	if (!check_ori) return nmatches;
	std::vector<std::pair<int, int>> matches;
	for (int i = 0; i < npts; i++)
		if (choice[i] >= 0) matches.push_back({ i, choice[i] });
	std::vector<int> erased;
	const int kept = check_orientation(matches, &pts[0].angle, sizeof(oracle_kf_point), &f->kps_un[0].angle, sizeof(oracle_keypoint), erased);
	for (int i2 : erased) frame_mp[i2] = -1;
	return kept;
}

// SearchByProjection(keyframe, Scw, mappoints, matched, th) (:518-612, loop closing): geometry through the similarity transform in the
// reference's operation order, then the same window search with octaves [predictedScale - 1, predictedScale] (:589-591).
int orc_search_sim3_projection(const oracle_frame_view* f, const oracle_camera* cam, const oracle_sim3* S, float log_scale_factor, int32_t* matched,
                               const oracle_sim3_point* pts, const uint8_t* pt_desc, int npts, int th)
{
	const Grid g = make_grid(f);
	const float invs = 1.f / S->s;                 // Sim3::Invs
	float t[3], Ow[3];
	for (int i = 0; i < 3; i++) t[i] = S->t[i] * invs;   // pose(Scw.R(), Scw.Invs() * Scw.t()) (:523)
	for (int i = 0; i < 3; i++)
	{
		float s = 0;
		for (int k = 0; k < 3; k++) s += (S->R[k * 3 + i] * -1) * t[k];
		Ow[i] = s;                                 // pose.Invt() (:525)
	}
	std::vector<Probe> probes((size_t)npts);
	for (int i = 0; i < npts; i++)
	{
		Probe& p = probes[i];
		p.active = false;
		p.obs = true;
		if (!(pts[i].flags & 1)) continue;         // :536-537
		float xc[3];
		for (int r = 0; r < 3; r++)
		{
			float s = 0;
			for (int k = 0; k < 3; k++) s += S->R[r * 3 + k] * pts[i].xw[k];
			xc[r] = s + t[r];
		}
		if (xc[2] < 0.f) continue;                 // :546-547
		const float invZ = 1.f / xc[2];
		p.u = invZ * cam->fx * xc[0] + cam->cx;
		p.v = invZ * cam->fy * xc[1] + cam->cy;
		p.ur = NAN;
		if (!(p.u >= f->bounds.minx && p.u < f->bounds.maxx && p.v >= f->bounds.miny && p.v < f->bounds.maxy)) continue;   // :555-556
		float PO[3];
		double ss = 0;
		for (int k = 0; k < 3; k++) { PO[k] = pts[i].xw[k] - Ow[k]; ss += (double)PO[k] * (double)PO[k]; }
		const float dist = (float)std::sqrt(ss);
		const float maxDistance = 1.2f * pts[i].max_distance, minDistance = 0.8f * pts[i].min_distance;
		if (dist < minDistance || dist > maxDistance) continue;                                            // :563-564
		float dot = 0;                             // Matx::dot accumulates in the element type
		for (int k = 0; k < 3; k++) dot += PO[k] * pts[i].normal[k];
		if (dot < 0.5 * dist) continue;            // :568-569, compared in double
		const float ratio = pts[i].max_distance / dist;
		const int scale = (int)std::ceil(std::log((double)ratio) / log_scale_factor);
		const int ps = std::max(0, std::min(scale, f->nlevels - 1));
		p.radius = th * f->scale_factors[ps];      // :574
		p.minLevel = ps - 1; p.maxLevel = ps;      // :590-591
		p.active = true;
	}
	std::vector<int> choice;
	return run_rounds(g, f, probes, pt_desc, matched, choice, [&](const Pick& k) { return k.idx >= 0 && k.best <= TH_LOW; }, true);
}

// SearchByBoW (:452-516 KeyFrame vs Frame when valid2 == NULL, :696-766 KeyFrame vs KeyFrame otherwise). FeatureVectorIterator
// (:406-450) walks the nodes both feature vectors share, in ascending node id: a merge join of the two sorted id arrays.
int orc_search_by_bow(const oracle_frame_view* f1, const oracle_feature_vector* fv1, const uint8_t* valid1, const oracle_frame_view* f2,
                      const oracle_feature_vector* fv2, const uint8_t* valid2, float nnratio, int check_ori, int32_t* match2)
{
	for (int i = 0; i < f2->n; i++) match2[i] = -1;
	std::vector<std::pair<int, int>> matchIds;      // (idx1, idx2) in acceptance order
	int nmatches = 0;
	int a = 0, b = 0;
	while (a < fv1->nnodes && b < fv2->nnodes)
	{
		if (fv1->node_ids[a] < fv2->node_ids[b]) { a++; continue; }
		if (fv2->node_ids[b] < fv1->node_ids[a]) { b++; continue; }
		for (int p1 = fv1->start[a]; p1 < fv1->start[a + 1]; p1++)
		{
			const int idx1 = (int)fv1->indices[p1];
			if (!valid1[idx1]) continue;                                            // :471-472, :719-720
			int best = 256, second = 256, bestIdx2 = -1;
			for (int p2 = fv2->start[b]; p2 < fv2->start[b + 1]; p2++)
			{
				const int idx2 = (int)fv2->indices[p2];
				if (match2[idx2] >= 0) continue;                                    // matches[idx2] / matched2[idx2]
				if (valid2 && !valid2[idx2]) continue;                             // :731-732
				const int d = hamming256(f1->desc + (size_t)idx1 * 32, f2->desc + (size_t)idx2 * 32);
				if (d < best) { second = best; best = d; bestIdx2 = idx2; }
				else if (d < second) second = d;
			}
			const bool low = valid2 ? best < TH_LOW : best <= TH_LOW;               // :750 vs :501
			if (low && best < nnratio * second)
			{
				match2[bestIdx2] = idx1;
				nmatches++;
				matchIds.push_back({ idx1, bestIdx2 });
			}
		}
		a++; b++;
	}
	if (!check_ori) return nmatches;
	std::vector<int> erased;
	if (!valid2)
	{
		// CheckOrientation(keyframe->keypointsUn, frame.keypointsUn, (idx1, bestIdx2), matches) (:512): status indexed by idx2
		nmatches = check_orientation(matchIds, &f1->kps_un[0].angle, sizeof(oracle_keypoint), &f2->kps_un[0].angle, sizeof(oracle_keypoint), erased);
		for (int i2 : erased) match2[i2] = -1;
	}
	else
	{
		// CheckOrientation(keypoints2, keypoints1, (bestIdx2, idx1), matches12) (:763): status indexed by idx1
		std::vector<std::pair<int, int>> swapped;
		std::vector<int> of1((size_t)f1->n, -1);
		for (const auto& m : matchIds) { swapped.push_back({ m.second, m.first }); of1[m.first] = m.second; }
		nmatches = check_orientation(swapped, &f2->kps_un[0].angle, sizeof(oracle_keypoint), &f1->kps_un[0].angle, sizeof(oracle_keypoint), erased);
		for (int i1 : erased) match2[of1[i1]] = -1;
	}
	return nmatches;
}

int orc_search_for_initialization(const oracle_frame_view* f1, const oracle_frame_view* f2, float* prev, int32_t* matches12, int window,
                                  float nnratio, int check_ori)
{
	const Grid g2 = make_grid(f2);
	int nmatches = 0;
	g_stat[1] = 0;
	for (int i = 0; i < f1->n; i++) matches12[i] = -1;
	std::vector<int> matchedDistance((size_t)f2->n, INT_MAX), matches21((size_t)f2->n, -1);
	std::vector<std::pair<int, int>> matchIds;
	const float radius = (float)window;
	for (int i1 = 0; i1 < f1->n; i1++)
	{
		if (f1->kps_un[i1].octave > 0) continue;                    // :629-631
		int best = INT_MAX, second = INT_MAX, bestIdx = -1;
		g2.query(prev[2 * i1], prev[2 * i1 + 1], radius, 0, 0, [&](int i2) {
			const int d = hamming256(f1->desc + (size_t)i1 * 32, f2->desc + (size_t)i2 * 32);
			if (matchedDistance[i2] <= d) return;                    // :651-652
			if (d < best) { second = best; best = d; bestIdx = i2; }
			else if (d < second) second = d;
		});
		if (bestIdx < 0) continue;
		if (best <= TH_LOW && best < second * nnratio)              // :666
		{
			if (matches21[bestIdx] >= 0)
			{
				matches12[matches21[bestIdx]] = -1;
				nmatches--;
				g_stat[1]++;
			}
			matches12[i1] = bestIdx;
			matches21[bestIdx] = i1;
			matchedDistance[bestIdx] = best;
			nmatches++;
			if (check_ori) matchIds.push_back({ bestIdx, i1 });
		}
	}
	if (check_ori)
	{
		std::vector<int> erased;
		nmatches = check_orientation(matchIds, &f2->kps_un[0].angle, sizeof(oracle_keypoint), &f1->kps_un[0].angle, sizeof(oracle_keypoint), erased);
		for (int i1 : erased) matches12[i1] = -1;
	}
	for (int i1 = 0; i1 < f1->n; i1++)
		if (matches12[i1] >= 0)
		{
			prev[2 * i1] = f2->kps_un[matches12[i1]].x;
			prev[2 * i1 + 1] = f2->kps_un[matches12[i1]].y;
		}
	return nmatches;
}


// ---- Fuse x2, SearchBySim3, SearchForTriangulation (src/ORBmatcher.cc:868-980, 982-1088, 1090-1277, 768-866) ----------------------------
// Restated in two halves, the split the CUDA library uses: (1) a per-point search that reads only the key frame's keypoints and the point
// (projection, gates, window, best distance, first keypoint on ties), (2) the sequential replay of the map mutation over the search results.
// The map model is the one of oracle/ref_guided_decl.h (one key frame; AddObservation / Replace as in src/MapPoint.cc:103-121, 193-240).
namespace {

struct Best { int idx, dist; };

// best keypoint of a window: octave in [lo, hi], optional chi-square gate of Fuse (:934-945); `dist < bestDist` keeps the first minimum
Best best_in_window(const Grid& g, const oracle_frame_view* f, float u, float v, float radius, int lo, int hi, const uint8_t* desc, bool gate,
                    float ur, const float* inv_sigma_sq)
{
	Best b = { -1, 256 };
	g.query(u, v, radius, -1, -1, [&](int idx) {
		const oracle_keypoint& kp = f->kps_un[idx];
		const int scale = kp.octave;
		if (scale < lo || scale > hi) return;
		if (gate)
		{
			const float dx = u - kp.x, dy = v - kp.y;
			const float ur2 = f->uright ? f->uright[idx] : -1.f;
			if (ur2 >= 0)
			{
				const float dz = ur - ur2;
				if ((dx * dx + dy * dy + dz * dz) * inv_sigma_sq[scale] > 7.8) return;
			}
			else if ((dx * dx + dy * dy) * inv_sigma_sq[scale] > 5.99) return;
		}
		const int d = hamming256(desc, f->desc + (size_t)idx * 32);
		if (d < b.dist) { b.dist = d; b.idx = idx; }
	});
	return b;
}

void mul3(const float* R, const float* x, float* y) { for (int r = 0; r < 3; r++) { float s = 0; for (int k = 0; k < 3; k++) s += R[r * 3 + k] * x[k]; y[r] = s; } }
void invt(const float* R, const float* t, float* o) { for (int i = 0; i < 3; i++) { float s = 0; for (int k = 0; k < 3; k++) s += (R[k * 3 + i] * -1) * t[k]; o[i] = s; } }
float norm3(const float* v) { double ss = 0; for (int k = 0; k < 3; k++) ss += (double)v[k] * (double)v[k]; return (float)std::sqrt(ss); }
int predict_scale(float max_distance, float dist, float lsf, int nlevels)
{
	const float ratio = max_distance / dist;
	const int scale = (int)std::ceil(std::log((double)ratio) / lsf);
	return std::max(0, std::min(scale, nlevels - 1));
}
bool in_image(const oracle_frame_view* f, float u, float v) { return u >= f->bounds.minx && u < f->bounds.maxx && v >= f->bounds.miny && v < f->bounds.maxy; }

// :879-954 / :1002-1066 for one point given its camera coordinates
Best fuse_search(const Grid& g, const oracle_frame_view* f, const oracle_camera* cam, const float* xc, const float* Ow, const oracle_sim3_point& p,
                 const uint8_t* desc, float lsf, float th, bool gate, const float* inv_sigma_sq)
{
	const Best none = { -1, 256 };
	if (xc[2] < 0.f) return none;
	const float invZ = 1.f / xc[2];
	const float u = invZ * cam->fx * xc[0] + cam->cx, v = invZ * cam->fy * xc[1] + cam->cy;
	if (!in_image(f, u, v)) return none;
	const float ur = u - cam->bf / xc[2];
	const float maxDistance = 1.2f * p.max_distance, minDistance = 0.8f * p.min_distance;
	float PO[3];
	for (int k = 0; k < 3; k++) PO[k] = p.xw[k] - Ow[k];
	const float dist3D = norm3(PO);
	if (dist3D < minDistance || dist3D > maxDistance) return none;
	float dot = 0;
	for (int k = 0; k < 3; k++) dot += PO[k] * p.normal[k];
	if (dot < 0.5 * dist3D) return none;
	const int ps = predict_scale(p.max_distance, dist3D, lsf, f->nlevels);
	return best_in_window(g, f, u, v, th * f->scale_factors[ps], ps - 1, ps, desc, gate, ur, inv_sigma_sq);
}

struct MapModel
{
	int32_t* kf_mp; int32_t* nobs; uint8_t* bad; std::vector<int> where;   // where[p] = keypoint holding p in the key frame, -1 = not in it
	std::vector<int> log;
	const oracle_frame_view* f;
	void add_observation(int p, int idx)
	{
		log.push_back(2); log.push_back(p); log.push_back(idx);
		where[p] = idx;
		nobs[p] += (f->uright && f->uright[idx] >= 0) ? 2 : 1;
		kf_mp[idx] = p;                                    // keyframe->AddMapPoint
	}
	void replace(int p, int other)                         // p->Replace(other)
	{
		log.push_back(1); log.push_back(p); log.push_back(other);
		if (p == other) return;
		bad[p] = 1;
		if (where[p] >= 0)
		{
			if (where[other] < 0) { kf_mp[where[p]] = other; where[other] = where[p]; nobs[other] += nobs[p]; }
			else kf_mp[where[p]] = -1;
		}
	}
};

}  // namespace

int orc_fuse(const oracle_frame_view* f, const oracle_camera* cam, const oracle_pose* pose, float lsf, const float* inv_sigma_sq,
             const oracle_sim3_point* pts, const uint8_t* pt_desc, int npts, float th, int32_t* kf_mp, int32_t* nobs, uint8_t* bad, uint8_t* in_kf,
             int32_t* log, int cap, int* nlog)
{
	const Grid g = make_grid(f);
	float Ow[3];
	invt(pose->R, pose->t, Ow);
	// (1) search, independent per point
	std::vector<Best> best((size_t)npts, Best{ -1, 256 });
	for (int i = 0; i < npts; i++)
	{
		if (!(pts[i].flags & 1)) continue;
		float xc[3];
		mul3(pose->R, pts[i].xw, xc);
		for (int k = 0; k < 3; k++) xc[k] += pose->t[k];
		best[i] = fuse_search(g, f, cam, xc, Ow, pts[i], pt_desc + (size_t)i * 32, lsf, th, true, inv_sigma_sq);
	}
	// (2) replay of :876 and :956-976 in order
	MapModel M{ kf_mp, nobs, bad, std::vector<int>((size_t)npts, -1), {}, f };
	for (int c = 0; c < f->n; c++)
		if (kf_mp[c] >= 0 && in_kf[kf_mp[c]]) M.where[kf_mp[c]] = c;
	int nfused = 0;
	for (int i = 0; i < npts; i++)
	{
		if (!(pts[i].flags & 1) || bad[i] || M.where[i] >= 0) continue;
		if (best[i].dist > TH_LOW) continue;
		const int held = kf_mp[best[i].idx];
		if (held >= 0)
		{
			if (!bad[held])
			{
				if (nobs[held] > nobs[i]) M.replace(i, held);
				else M.replace(held, i);
			}
		}
		else M.add_observation(i, best[i].idx);
		nfused++;
	}
	for (int i = 0; i < npts; i++) in_kf[i] = M.where[i] >= 0;
	*nlog = (int)std::min<size_t>(M.log.size(), (size_t)cap);
	for (int i = 0; i < *nlog; i++) log[i] = M.log[i];
	return nfused;
}

int orc_fuse_sim3(const oracle_frame_view* f, const oracle_camera* cam, const oracle_sim3* S, float lsf, const oracle_sim3_point* pts,
                  const uint8_t* pt_desc, int npts, float th, int32_t* kf_mp, int32_t* nobs, uint8_t* bad, int32_t* replace, int32_t* log, int cap,
                  int* nlog)
{
	const Grid g = make_grid(f);
	const float invs = 1.f / S->s;
	float t[3], Ow[3];
	for (int i = 0; i < 3; i++) t[i] = S->t[i] * invs;
	invt(S->R, t, Ow);
	// alreadyFound = keyframe->GetMapPoints() taken once, before the loop (:992)
	std::vector<uint8_t> found((size_t)npts, 0);
	for (int c = 0; c < f->n; c++)
		if (kf_mp[c] >= 0 && !bad[kf_mp[c]]) found[kf_mp[c]] = 1;
	std::vector<Best> best((size_t)npts, Best{ -1, 256 });
	for (int i = 0; i < npts; i++)
	{
		float xc[3];
		mul3(S->R, pts[i].xw, xc);
		for (int k = 0; k < 3; k++) xc[k] += t[k];
		best[i] = fuse_search(g, f, cam, xc, Ow, pts[i], pt_desc + (size_t)i * 32, lsf, th, false, nullptr);
	}
	MapModel M{ kf_mp, nobs, bad, std::vector<int>((size_t)npts, -1), {}, f };
	for (int c = 0; c < f->n; c++)
		if (kf_mp[c] >= 0) M.where[kf_mp[c]] = c;
	int nfused = 0;
	for (int i = 0; i < npts; i++)
	{
		replace[i] = -1;
		if (bad[i] || found[i]) continue;                  // :1002-1003: isBad() is read at this point of the loop, alreadyFound is the snapshot
		if (best[i].dist > TH_LOW) continue;
		const int held = kf_mp[best[i].idx];
		if (held >= 0) { if (!bad[held]) replace[i] = held; }
		else M.add_observation(i, best[i].idx);
		nfused++;
	}
	*nlog = (int)std::min<size_t>(M.log.size(), (size_t)cap);
	for (int i = 0; i < *nlog; i++) log[i] = M.log[i];
	return nfused;
}

int orc_search_by_sim3(const oracle_frame_view* f1, const oracle_camera* cam1, const oracle_pose* pose1, float lsf1, const oracle_frame_view* f2,
                       const oracle_camera* cam2, const oracle_pose* pose2, float lsf2, const oracle_sim3* S12, float th, const oracle_kf_point* pts1,
                       const uint8_t* desc1, const oracle_kf_point* pts2, const uint8_t* desc2, int32_t* matches12)
{
	const Grid g1 = make_grid(f1), g2 = make_grid(f2);
	const int N1 = f1->n, N2 = f2->n;
	// S21 = S12.Inverse() (include/Sim3.h:42-47)
	float R21[9], t21[3], sR12[9], sR21[9];
	const float is12 = 1.f / S12->s, nis = -is12;
	for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R21[r * 3 + c] = S12->R[c * 3 + r];
	for (int i = 0; i < 3; i++) { float s = 0; for (int k = 0; k < 3; k++) s += (R21[i * 3 + k] * nis) * S12->t[k]; t21[i] = s; }
	for (int k = 0; k < 9; k++) { sR12[k] = S12->R[k] * S12->s; sR21[k] = R21[k] * is12; }
	// alreadyMatched (:1109-1121): flags bit1 of a kf1 point = matches12 holds kf2's point (i1 * 7) % N2 on entry
	std::vector<uint8_t> am1((size_t)N1, 0), am2((size_t)N2, 0);
	for (int i = 0; i < N1; i++)
		if (pts1[i].flags & 2) { am1[i] = 1; am2[(size_t)((i * 7) % N2)] = 1; }
	auto direction = [&](const oracle_frame_view* from, const oracle_pose* pf, const oracle_kf_point* pts, const uint8_t* desc, const std::vector<uint8_t>& am,
	                     const float* sR, const float* t, const Grid& gt, const oracle_frame_view* to, const oracle_camera* cam, float lsf) {
		std::vector<int> match((size_t)from->n, -1);
		for (int i = 0; i < from->n; i++)
		{
			if (!(pts[i].flags & (1 | 4)) || am[i] || (pts[i].flags & 4)) continue;     // null, already matched, bad
			float xa[3], xb[3];
			mul3(pf->R, pts[i].xw, xa);
			for (int k = 0; k < 3; k++) xa[k] += pf->t[k];
			mul3(sR, xa, xb);
			for (int k = 0; k < 3; k++) xb[k] += t[k];
			if (xb[2] < 0.f) continue;
			const float invZ = 1.f / xb[2];
			const float u = invZ * cam->fx * xb[0] + cam->cx, v = invZ * cam->fy * xb[1] + cam->cy;
			if (!in_image(to, u, v)) continue;
			const float maxDistance = 1.2f * pts[i].max_distance, minDistance = 0.8f * pts[i].min_distance;
			const float dist3D = norm3(xb);
			if (dist3D < minDistance || dist3D > maxDistance) continue;
			const int ps = predict_scale(pts[i].max_distance, dist3D, lsf, to->nlevels);
			const Best b = best_in_window(gt, to, u, v, th * to->scale_factors[ps], ps - 1, ps, desc + (size_t)i * 32, false, 0.f, nullptr);
			if (b.dist <= TH_HIGH) match[i] = b.idx;
		}
		return match;
	};
	const std::vector<int> m1 = direction(f1, pose1, pts1, desc1, am1, sR21, t21, g2, f2, cam2, lsf2);
	const std::vector<int> m2 = direction(f2, pose2, pts2, desc2, am2, sR12, S12->t, g1, f1, cam1, lsf1);
	int nfound = 0;
	for (int i = 0; i < N1; i++)
	{
		matches12[i] = am1[i] ? (i * 7) % N2 : -1;
		const int idx2 = m1[i];
		if (idx2 >= 0 && m2[idx2] == i) { matches12[i] = idx2; nfound++; }
	}
	return nfound;
}

int orc_search_for_triangulation(const oracle_frame_view* f1, const oracle_feature_vector* fv1, const uint8_t* has1, const oracle_frame_view* f2,
                                 const oracle_feature_vector* fv2, const uint8_t* has2, const float* F, const float* ep2, const float* sigma_sq2,
                                 int only_stereo, int check_ori, int32_t* matches12)
{
	for (int i = 0; i < f1->n; i++) matches12[i] = -1;
	int nmatches = 0;
	std::vector<std::pair<int, int>> tmp;
	for (int a = 0, b = 0; a < fv1->nnodes && b < fv2->nnodes;)
	{
		if (fv1->node_ids[a] < fv2->node_ids[b]) { a++; continue; }
		if (fv1->node_ids[a] > fv2->node_ids[b]) { b++; continue; }
		for (int j = fv1->start[a]; j < fv1->start[a + 1]; j++)
		{
			const int idx1 = (int)fv1->indices[j];
			if (has1[idx1]) continue;
			const bool stereo1 = f1->uright && f1->uright[idx1] >= 0;
			if (only_stereo && !stereo1) continue;
			const oracle_keypoint& k1 = f1->kps_un[idx1];
			const float la = k1.x * F[0] + k1.y * F[3] + F[6];
			const float lb = k1.x * F[1] + k1.y * F[4] + F[7];
			const float lc = k1.x * F[2] + k1.y * F[5] + F[8];
			int bestDist = TH_LOW, bestIdx2 = -1;
			for (int q = fv2->start[b]; q < fv2->start[b + 1]; q++)
			{
				const int idx2 = (int)fv2->indices[q];
				if (has2[idx2]) continue;                  // matched2 is never set in this fork
				const bool stereo2 = f2->uright && f2->uright[idx2] >= 0;
				if (only_stereo && !stereo2) continue;
				const int dist = hamming256(f1->desc + (size_t)idx1 * 32, f2->desc + (size_t)idx2 * 32);
				if (dist > TH_LOW || dist > bestDist) continue;
				const oracle_keypoint& k2 = f2->kps_un[idx2];
				if (!stereo1 && !stereo2)
				{
					const float dx = ep2[0] - k2.x, dy = ep2[1] - k2.y;
					if (dx * dx + dy * dy < 100 * f2->scale_factors[k2.octave]) continue;
				}
				const float num = la * k2.x + lb * k2.y + lc;
				const float den = la * la + lb * lb;
				if (den == 0) continue;
				const float dsqr = num * num / den;
				if (dsqr < 3.84 * sigma_sq2[k2.octave]) { bestIdx2 = idx2; bestDist = dist; }
			}
			if (bestIdx2 >= 0)
			{
				matches12[idx1] = bestIdx2;
				nmatches++;
				if (check_ori) tmp.emplace_back(bestIdx2, idx1);
			}
		}
		a++; b++;
	}
	if (check_ori)
	{
		std::vector<int> erased;
		nmatches = check_orientation(tmp, &f2->kps_un[0].angle, sizeof(oracle_keypoint), &f1->kps_un[0].angle, sizeof(oracle_keypoint), erased);
		for (int i2 : erased) matches12[i2] = -1;
	}
	return nmatches;
}

}  // extern "C"
