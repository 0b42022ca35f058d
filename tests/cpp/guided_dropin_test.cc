// C++ drop-in check of include/orbx/GuidedMatcher.h: Frame / MapPoint types with the reference's member names (include/Frame.h:83-168,
// include/MapPoint.h:49-97) go through the templates; results must equal the CPU oracle's. Prints "OK ..." or a diagnostic.
// Run by tests/test_cpp_dropin.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <set>
#include <vector>

#include "orbx/GuidedMatcher.h"

#define ORACLE_PREFIX orc_
#include "oracle_api.h"

namespace {

struct ImageBounds { float minx, maxx, miny, maxy; };
struct ScalePyramidInfo { std::vector<float> scaleFactors; float logScaleFactor = 0.f; int nlevels = 8; };
struct CameraParams { float fx, fy, cx, cy, bf, baseline; };
struct CameraPose
{
	cv::Matx33f R_; cv::Matx31f t_;
	const cv::Matx33f& R() const { return R_; }
	const cv::Matx31f& t() const { return t_; }
};

struct Frame;
struct MapPoint
{
	float minDistance_ = 0, maxDistance_ = 0;                          // include/MapPoint.h; accessors as src/MapPoint.cc:382-414
	float GetMinDistanceInvariance() const { return 0.8f * minDistance_; }
	float GetMaxDistanceInvariance() const { return 1.2f * maxDistance_; }
	inline int PredictScale(float dist, const Frame* frame) const;
	float trackProjX = 0, trackProjY = 0, trackProjXR = 0, trackViewCos = 0;
	bool trackInView = false;
	int trackScaleLevel = 0;
	cv::Matx31f worldPos;
	int nobs = 0;
	unsigned char desc[32];
	bool isBad() const { return false; }
	int Observations() const { return nobs; }
	cv::Mat GetDescriptor() const { return cv::Mat(1, 32, CV_8U, (void*)desc, 32); }
	cv::Matx31f GetWorldPos() const { return worldPos; }
};

using FeatureVector = std::map<unsigned, std::vector<unsigned>>;   // DBoW2::FeatureVector's base class

struct Frame
{
	FeatureVector featureVector;
	CameraParams camera;
	int N = 0;
	std::vector<cv::KeyPoint> keypoints, keypointsUn;
	std::vector<float> uright;
	cv::Mat descriptors;
	std::vector<MapPoint*> mappoints;
	std::vector<bool> outlier;
	CameraPose pose;
	ScalePyramidInfo pyramid;
	ImageBounds imageBounds;
	cv::Matx31f GetCameraCenter() const { return -pose.R_.t() * pose.t_; }   // src/Frame.cc:203-206, include/CameraPose.h:47
};

inline int MapPoint::PredictScale(float dist, const Frame* frame) const
{
	const float ratio = maxDistance_ / dist;
	const int scale = static_cast<int>(ceil(log(ratio) / frame->pyramid.logScaleFactor));
	return std::max(0, std::min(scale, frame->pyramid.nlevels - 1));
}

std::mt19937 rng(12345);
float uni(float a, float b) { return std::uniform_real_distribution<float>(a, b)(rng); }
int irand(int n) { return (int)(rng() % (unsigned)n); }

void make_frame(Frame& f, int n, std::vector<unsigned char>& desc)
{
	f.N = n;
	f.pyramid.scaleFactors.resize(8);
	f.pyramid.logScaleFactor = (float)log((double)1.2f);
	f.pyramid.scaleFactors[0] = 1.f;
	for (int i = 1; i < 8; i++) f.pyramid.scaleFactors[i] = f.pyramid.scaleFactors[i - 1] * 1.2f;
	f.imageBounds = { 0.f, 640.f, 0.f, 480.f };
	f.camera = { 517.3f, 516.5f, 318.6f, 255.3f, 40.f, 40.f / 517.3f };
	f.keypointsUn.resize(n);
	f.uright.resize(n);
	desc.resize((size_t)n * 32);
	for (int i = 0; i < n; i++)
	{
		const int oct = irand(10) < 5 ? 0 : irand(8);
		f.keypointsUn[i] = cv::KeyPoint(uni(16, 624), uni(16, 464), 31.f * f.pyramid.scaleFactors[oct], uni(0, 359.9f), 20.f, oct);
		f.uright[i] = irand(10) < 7 ? f.keypointsUn[i].pt.x - uni(1, 60) : -1.f;
		for (int b = 0; b < 32; b++) desc[(size_t)i * 32 + b] = (unsigned char)rng();
	}
	f.keypoints = f.keypointsUn;
	f.descriptors = cv::Mat(n, 32, CV_8U, desc.data(), 32);
	f.mappoints.assign(n, nullptr);
	f.outlier.assign(n, false);
}

void noisy_copy(unsigned char* dst, const unsigned char* src, int flips)
{
	memcpy(dst, src, 32);
	for (int k = 0; k < flips; k++) { const int b = irand(256); dst[b >> 3] ^= (unsigned char)(1 << (b & 7)); }
}

struct KeyFrame
{
	std::vector<cv::KeyPoint> keypointsUn;
	cv::Mat descriptors;
	ImageBounds imageBounds;
	ScalePyramidInfo pyramid;
	std::vector<float> uright;
	FeatureVector featureVector;
	std::vector<MapPoint*> mappoints;
	std::vector<MapPoint*> GetMapPointMatches() const { return mappoints; }
};

struct FlatFv
{
	std::vector<uint32_t> ids, idx; std::vector<int32_t> start;
	explicit FlatFv(const FeatureVector& fv)
	{
		start.push_back(0);
		for (const auto& n : fv) { ids.push_back(n.first); for (unsigned i : n.second) idx.push_back(i); start.push_back((int32_t)idx.size()); }
	}
	oracle_feature_vector view() const { return oracle_feature_vector{ (int32_t)ids.size(), ids.data(), start.data(), idx.data() }; }
};

oracle_frame_view view_of(const Frame& f)
{
	oracle_frame_view v;
	v.n = f.N; v.kps_un = reinterpret_cast<const oracle_keypoint*>(f.keypointsUn.data()); v.desc = f.descriptors.data;
	v.uright = f.uright.data();
	v.bounds = { f.imageBounds.minx, f.imageBounds.maxx, f.imageBounds.miny, f.imageBounds.maxy };
	v.nlevels = 8; v.scale_factors = f.pyramid.scaleFactors.data();
	return v;
}

}  // namespace

int main()
{
	try
	{
		const int n = 1200, npts = 1500;
		Frame frame;
		std::vector<unsigned char> fdesc;
		make_frame(frame, n, fdesc);
		ORB_SLAM2::b200::DeviceFrame dev(frame);

		// ---- GetFeaturesInArea
		const oracle_frame_view fv = view_of(frame);
		void* grid = orc_grid_create(fv.kps_un, n, &fv.bounds, 8);
		std::vector<int32_t> want(n);
		for (int q = 0; q < 50; q++)
		{
			const float x = uni(-20, 660), y = uni(-20, 500), r = uni(1, 60);
			const int lo = irand(4) - 1, hi = irand(9) - 1;
			const int cnt = orc_grid_query(grid, x, y, r, lo, hi, want.data(), n);
			const std::vector<size_t> got = dev.GetFeaturesInArea(x, y, r, lo, hi);
			if ((int)got.size() != cnt) { printf("GetFeaturesInArea count %zu vs %d\n", got.size(), cnt); return 1; }
			for (int i = 0; i < cnt; i++) if ((int)got[i] != want[i]) { printf("GetFeaturesInArea order\n"); return 1; }
		}
		orc_grid_destroy(grid);

		// ---- SearchByProjection(Frame&, mappoints, th)
		std::vector<MapPoint> store(npts);
		std::vector<MapPoint*> list(npts);
		std::vector<oracle_track_point> opts(npts);
		std::vector<unsigned char> odesc((size_t)npts * 32);
		for (int i = 0; i < npts; i++)
		{
			MapPoint& m = store[i];
			const int t = irand(n);
			const cv::KeyPoint& kp = frame.keypointsUn[t];
			m.trackScaleLevel = std::min(7, kp.octave + irand(2));
			m.trackProjX = kp.pt.x + uni(-3, 3); m.trackProjY = kp.pt.y + uni(-3, 3);
			m.trackProjXR = frame.uright[t] + uni(-2, 2);
			m.trackViewCos = uni(0.99f, 1.f);
			m.trackInView = irand(10) < 9;
			m.nobs = irand(10) < 8;
			noisy_copy(m.desc, &fdesc[(size_t)t * 32], irand(60));
			list[i] = &m;
			opts[i] = { m.trackProjX, m.trackProjY, m.trackProjXR, m.trackViewCos, m.trackScaleLevel, (m.trackInView ? 1 : 0) | (m.nobs ? 2 : 0) };
			memcpy(&odesc[(size_t)i * 32], m.desc, 32);
		}
		MapPoint seen, unseen;
		seen.nobs = 3;
		std::vector<int32_t> code(n, -1);
		for (int c = 0; c < n; c++)
		{
			const int u = irand(100);
			if (u < 5) { frame.mappoints[c] = &seen; code[c] = -2; }
			else if (u < 8) { frame.mappoints[c] = &unseen; code[c] = -3; }
		}
		const int want_n = orc_search_local_map(&fv, code.data(), opts.data(), odesc.data(), npts, 3.f, 0.8f);
		const int got_n = ORB_SLAM2::b200::GuidedMatcher(0.8f, true).SearchByProjection(frame, dev, list, 3.f);
		if (want_n != got_n) { printf("local map nmatches %d vs %d\n", got_n, want_n); return 1; }
		for (int c = 0; c < n; c++)
		{
			const MapPoint* w = code[c] >= 0 ? &store[code[c]] : code[c] == -2 ? &seen : code[c] == -3 ? &unseen : nullptr;
			if (frame.mappoints[c] != w) { printf("local map frame.mappoints[%d]\n", c); return 1; }
		}

		// ---- SearchByProjection(currFrame, lastFrame, th, monocular)
		Frame last;
		last.N = npts;
		last.keypoints.resize(npts); last.keypointsUn.resize(npts);
		last.mappoints.assign(npts, nullptr); last.outlier.assign(npts, false);
		frame.pose.R_ = cv::Matx33f(1, 0, 0, 0, 1, 0, 0, 0, 1); frame.pose.t_ = cv::Matx31f(0.02f, -0.01f, 0.05f);
		last.pose.R_ = frame.pose.R_; last.pose.t_ = cv::Matx31f(0.f, 0.f, 0.3f);
		std::vector<oracle_last_point> lpts(npts);
		for (int i = 0; i < npts; i++)
		{
			MapPoint& m = store[i];
			const int t = irand(n);
			const cv::KeyPoint& kp = frame.keypointsUn[t];
			const float z = uni(2, 30), u = kp.pt.x + uni(-4, 4), v = kp.pt.y + uni(-4, 4);
			const cv::Matx31f Xc((u - frame.camera.cx) / frame.camera.fx * z, (v - frame.camera.cy) / frame.camera.fy * z, z);
			m.worldPos = Xc - frame.pose.t_;
			noisy_copy(m.desc, &fdesc[(size_t)t * 32], irand(60));
			last.keypoints[i].octave = last.keypointsUn[i].octave = std::max(0, kp.octave - irand(2));
			last.keypointsUn[i].angle = std::fmod(kp.angle + 20.f + uni(-5, 5) + 360.f, 360.f);
			const bool has = irand(10) < 9;
			last.mappoints[i] = has ? &m : nullptr;
			last.outlier[i] = irand(20) == 0;
			lpts[i] = { { m.worldPos(0), m.worldPos(1), m.worldPos(2) }, last.keypoints[i].octave, last.keypointsUn[i].angle,
			            ((has && !last.outlier[i]) ? 1 : 0) | (m.nobs ? 2 : 0) };
			memcpy(&odesc[(size_t)i * 32], m.desc, 32);
		}
		for (int c = 0; c < n; c++) { frame.mappoints[c] = nullptr; code[c] = -1; }
		const oracle_camera ocam = { frame.camera.fx, frame.camera.fy, frame.camera.cx, frame.camera.cy, frame.camera.bf, frame.camera.baseline };
		oracle_pose cp, lp;
		for (int i = 0; i < 9; i++) { cp.R[i] = frame.pose.R_.val[i]; lp.R[i] = last.pose.R_.val[i]; }
		for (int i = 0; i < 3; i++) { cp.t[i] = frame.pose.t_.val[i]; lp.t[i] = last.pose.t_.val[i]; }
		int total = 0;
		for (int mono = 0; mono < 2; mono++)
		{
			for (int c = 0; c < n; c++) { frame.mappoints[c] = nullptr; code[c] = -1; }
			const int w2 = orc_search_last_frame(&fv, &ocam, &cp, &lp, code.data(), lpts.data(), odesc.data(), npts, mono ? 15.f : 7.f, mono, 0.9f, 1);
			const int g2 = ORB_SLAM2::b200::GuidedMatcher(0.9f, true).SearchByProjection(frame, dev, last, mono ? 15.f : 7.f, mono != 0);
			if (w2 != g2) { printf("last frame nmatches %d vs %d\n", g2, w2); return 1; }
			for (int c = 0; c < n; c++)
				if (frame.mappoints[c] != (code[c] >= 0 ? &store[code[c]] : nullptr)) { printf("last frame frame.mappoints[%d]\n", c); return 1; }
			total += g2;
		}

		// ---- SearchForInitialization
		Frame f2;
		std::vector<unsigned char> d2;
		make_frame(f2, n, d2);
		for (int i = 0; i < n; i += 2)
		{
			const int j = irand(n);
			f2.keypointsUn[j] = frame.keypointsUn[i];
			f2.keypointsUn[j].pt.x += uni(-5, 5); f2.keypointsUn[j].pt.y += uni(-5, 5);
			noisy_copy(&d2[(size_t)j * 32], &fdesc[(size_t)i * 32], irand(50));
		}
		ORB_SLAM2::b200::DeviceFrame dev2(f2);
		std::vector<cv::Point2f> prev(n), prev_want(n);
		for (int i = 0; i < n; i++) prev[i] = prev_want[i] = frame.keypointsUn[i].pt;
		std::vector<int> m12, m12_want(n);
		const oracle_frame_view fv2 = view_of(f2);
		const int w3 = orc_search_for_initialization(&fv, &fv2, reinterpret_cast<float*>(prev_want.data()), m12_want.data(), 100, 0.9f, 1);
		const int g3 = ORB_SLAM2::b200::GuidedMatcher(0.9f, true).SearchForInitialization(frame, dev, f2, dev2, prev, m12, 100);
		if (w3 != g3 || m12 != m12_want) { printf("initialization %d vs %d\n", g3, w3); return 1; }
		if (memcmp(prev.data(), prev_want.data(), sizeof(cv::Point2f) * n) != 0) { printf("prevMatched differs\n"); return 1; }

		// ---- SearchByProjection(frame, keyframe, alreadyFound, th, ORBdist): reuse `last`'s world points as the key frame's map points
		{
			KeyFrame kf;
			kf.keypointsUn = last.keypointsUn;
			kf.mappoints.assign(npts, nullptr);
			std::set<MapPoint*> found;
			std::vector<oracle_kf_point> kpts(npts);
			const cv::Matx31f Ow = frame.GetCameraCenter();
			for (int i = 0; i < npts; i++)
			{
				MapPoint& m = store[i];
				const cv::Matx31f d = m.worldPos - Ow;
				const float dist = (float)cv::norm(d);
				m.maxDistance_ = dist * frame.pyramid.scaleFactors[last.keypoints[i].octave] * uni(0.9f, 1.1f);
				m.minDistance_ = m.maxDistance_ / frame.pyramid.scaleFactors[7];
				const bool has = irand(10) < 9;
				if (has) kf.mappoints[i] = &m;
				if (has && irand(10) == 0) found.insert(&m);
				kpts[i] = { { m.worldPos(0), m.worldPos(1), m.worldPos(2) }, m.minDistance_, m.maxDistance_, kf.keypointsUn[i].angle,
				            (has && !found.count(&m)) ? 1 : 0 };
			}
			MapPoint other;
			for (int c = 0; c < n; c++)
			{
				const bool occ = irand(20) == 0;
				frame.mappoints[c] = occ ? &other : nullptr; code[c] = occ ? -3 : -1;
			}
			const int w6 = orc_search_keyframe_projection(&fv, &ocam, &cp, frame.pyramid.logScaleFactor, code.data(), kpts.data(), odesc.data(), npts, 10.f,
			                                              100, 1);
			const int g6 = ORB_SLAM2::b200::GuidedMatcher(0.9f, true).SearchByProjection(frame, dev, &kf, found, 10.f, 100);
			if (w6 != g6) { printf("relocalisation nmatches %d vs %d\n", g6, w6); return 1; }
			for (int c = 0; c < n; c++)
			{
				const MapPoint* w = code[c] >= 0 ? &store[code[c]] : code[c] == -3 ? &other : nullptr;
				if (frame.mappoints[c] != w) { printf("relocalisation frame.mappoints[%d]\n", c); return 1; }
			}
			total += g6;
		}

		// ---- SearchByBoW, both variants: frame -> key frame 1, f2 -> frame / key frame 2
		KeyFrame kf1, kf2;
		kf1.keypointsUn = frame.keypointsUn; kf1.descriptors = frame.descriptors; kf1.imageBounds = frame.imageBounds; kf1.pyramid = frame.pyramid;
		kf2.keypointsUn = f2.keypointsUn; kf2.descriptors = f2.descriptors; kf2.imageBounds = f2.imageBounds; kf2.pyramid = f2.pyramid;
		std::vector<MapPoint> mp1(n), mp2(n);
		std::vector<uint8_t> va1(n), va2(n);
		kf1.mappoints.assign(n, nullptr); kf2.mappoints.assign(n, nullptr);
		for (int i = 0; i < n; i++)
		{
			va1[i] = irand(10) < 8; va2[i] = irand(10) < 8;
			if (va1[i]) kf1.mappoints[i] = &mp1[i];
			if (va2[i]) kf2.mappoints[i] = &mp2[i];
			kf1.featureVector[1000 + 7 * (unsigned)irand(60)].push_back((unsigned)i);
		}
		for (int i = 0; i < n; i++) kf2.featureVector[1000 + 7 * (unsigned)irand(60)].push_back((unsigned)i);
		f2.featureVector = kf2.featureVector;
		const FlatFv a1(kf1.featureVector), a2(kf2.featureVector);
		const oracle_feature_vector o1 = a1.view(), o2 = a2.view();
		std::vector<int32_t> want_m2(n);
		ORB_SLAM2::b200::GuidedMatcher bow(0.9f, true);
		{
			const int w4 = orc_search_by_bow(&fv, &o1, va1.data(), &fv2, &o2, nullptr, 0.9f, 1, want_m2.data());
			std::vector<MapPoint*> matches;
			const int g4 = bow.SearchByBoW(&kf1, dev, f2, dev2, matches);
			if (w4 != g4) { printf("bow kf-frame nmatches %d vs %d\n", g4, w4); return 1; }
			for (int c = 0; c < n; c++)
				if (matches[c] != (want_m2[c] >= 0 ? &mp1[want_m2[c]] : nullptr)) { printf("bow kf-frame matches[%d]\n", c); return 1; }
			total += g4;
		}
		{
			const int w5 = orc_search_by_bow(&fv, &o1, va1.data(), &fv2, &o2, va2.data(), 0.9f, 1, want_m2.data());
			std::vector<MapPoint*> matches12;
			const int g5 = bow.SearchByBoW(&kf1, dev, &kf2, dev2, matches12);
			if (w5 != g5) { printf("bow kf-kf nmatches %d vs %d\n", g5, w5); return 1; }
			std::vector<MapPoint*> want12(n, nullptr);
			for (int c = 0; c < n; c++) if (want_m2[c] >= 0) want12[want_m2[c]] = &mp2[c];
			if (matches12 != want12) { printf("bow kf-kf matches12\n"); return 1; }
			total += g5;
		}

		printf("OK local %d, last+bow %d, init %d\n", got_n, total, g3);
		return 0;
	}
	catch (const cv::Exception& e)
	{
		printf("exception: %s\n", e.what());
		return 3;
	}
}
