// C++ drop-in check of the local-mapping / loop-closing templates of include/orbx/GuidedMatcher.h — Fuse x2, SearchBySim3,
// SearchForTriangulation — with KeyFrame / MapPoint / Sim3 types that carry the reference's member names (include/KeyFrame.h,
// include/MapPoint.h, include/Sim3.h). The map model is the one-key-frame model of the oracle (oracle/ref_guided_decl.h); results and
// the order of the map mutations must equal the CPU oracle's. Prints "OK ..." or a diagnostic. Run by tests/test_cpp_dropin.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <set>
#include <vector>

#include "orbx/GuidedMatcher.h"
#include "orbx/ORBVocabulary.h"

#define ORACLE_PREFIX orc_
#include "oracle_api.h"

namespace {

struct ImageBounds { float minx, maxx, miny, maxy; bool Contains(float x, float y) const { return x >= minx && x < maxx && y >= miny && y < maxy; } };
struct ScalePyramidInfo { std::vector<float> scaleFactors, sigmaSq, invSigmaSq; float logScaleFactor = 0.f; int nlevels = 8; };
struct CameraParams { float fx, fy, cx, cy, bf, baseline; };
struct CameraPose
{
	cv::Matx33f R_; cv::Matx31f t_;
	const cv::Matx33f& R() const { return R_; }
	const cv::Matx31f& t() const { return t_; }
	cv::Matx31f Invt() const { return -R_.t() * t_; }
};
struct Sim3 : CameraPose                                                   // include/Sim3.h:31-47
{
	float s_ = 1.f;
	Sim3() {}
	Sim3(const cv::Matx33f& R, const cv::Matx31f& t, float s) { R_ = R; t_ = t; s_ = s; }
	cv::Matx31f Map(const cv::Matx31f& x) const { return s_ * R_ * x + t_; }
	cv::Matx33f InvR() const { return R_.t(); }
	cv::Matx31f Invt() const { return -(1.f / s_) * R_.t() * t_; }
	float Invs() const { return 1.f / s_; }
	Sim3 Inverse() const { return Sim3(InvR(), Invt(), Invs()); }
};
using FeatureVector = std::map<unsigned, std::vector<unsigned>>;

struct KeyFrame;
struct MapPoint
{
	int id = -1;
	cv::Matx31f worldPos, normal;
	float minDistance_ = 0, maxDistance_ = 0;
	int nobs = 0;
	bool bad = false;
	unsigned char desc[32];
	const KeyFrame* inKeyFrame = nullptr;
	int indexInKeyFrame = -1;
	std::vector<int>* actions = nullptr;
	bool isBad() const { return bad; }
	int Observations() const { return nobs; }
	cv::Matx31f GetWorldPos() const { return worldPos; }
	cv::Matx31f GetNormal() const { return normal; }
	cv::Mat GetDescriptor() const { return cv::Mat(1, 32, CV_8U, (void*)desc, 32); }
	float GetMinDistanceInvariance() const { return 0.8f * minDistance_; }
	float GetMaxDistanceInvariance() const { return 1.2f * maxDistance_; }
	inline int PredictScale(float dist, const KeyFrame* kf) const;
	bool IsInKeyFrame(const KeyFrame* kf) const { return inKeyFrame == kf; }
	int GetIndexInKeyFrame(const KeyFrame* kf) const { return inKeyFrame == kf ? indexInKeyFrame : -1; }
	inline void AddObservation(KeyFrame* kf, size_t idx);
	inline void Replace(MapPoint* other);
};

struct KeyFrame
{
	int N = 0;
	std::vector<cv::KeyPoint> keypointsUn;
	std::vector<float> uright;
	cv::Mat descriptors;
	std::vector<MapPoint*> mappoints;
	CameraParams camera;
	ScalePyramidInfo pyramid;
	ImageBounds imageBounds;
	CameraPose pose;
	FeatureVector featureVector;
	CameraPose GetPose() const { return pose; }
	cv::Matx31f GetCameraCenter() const { return pose.Invt(); }
	bool IsInImage(float x, float y) const { return imageBounds.Contains(x, y); }
	MapPoint* GetMapPoint(size_t idx) const { return mappoints[idx]; }
	void AddMapPoint(MapPoint* mp, size_t idx) { mappoints[idx] = mp; }
	std::vector<MapPoint*> GetMapPointMatches() const { return mappoints; }
	std::set<MapPoint*> GetMapPoints() const
	{
		std::set<MapPoint*> s;
		for (MapPoint* mp : mappoints) if (mp && !mp->isBad()) s.insert(mp);
		return s;
	}
};

inline int MapPoint::PredictScale(float dist, const KeyFrame* kf) const
{
	const float ratio = maxDistance_ / dist;
	const int scale = static_cast<int>(ceil(log(ratio) / kf->pyramid.logScaleFactor));
	return std::max(0, std::min(scale, kf->pyramid.nlevels - 1));
}
inline void MapPoint::AddObservation(KeyFrame* kf, size_t idx)
{
	if (actions) { actions->push_back(2); actions->push_back(id); actions->push_back((int)idx); }
	inKeyFrame = kf; indexInKeyFrame = (int)idx;
	nobs += kf->uright[idx] >= 0 ? 2 : 1;
}
inline void MapPoint::Replace(MapPoint* other)
{
	if (actions) { actions->push_back(1); actions->push_back(id); actions->push_back(other->id); }
	if (other == this) return;
	bad = true;
	if (inKeyFrame)
	{
		KeyFrame* kf = const_cast<KeyFrame*>(inKeyFrame);
		if (!other->IsInKeyFrame(kf))
		{
			kf->mappoints[(size_t)indexInKeyFrame] = other;
			other->inKeyFrame = kf; other->indexInKeyFrame = indexInKeyFrame;
			other->nobs += nobs;
		}
		else kf->mappoints[(size_t)indexInKeyFrame] = nullptr;
	}
}

std::mt19937 rng(4242);
float uni(float a, float b) { return std::uniform_real_distribution<float>(a, b)(rng); }
int irand(int n) { return (int)(rng() % (unsigned)n); }

void make_keyframe(KeyFrame& f, int n, std::vector<unsigned char>& desc, float stereo_share)
{
	f.N = n;
	f.pyramid.scaleFactors.resize(8); f.pyramid.sigmaSq.resize(8); f.pyramid.invSigmaSq.resize(8);
	f.pyramid.logScaleFactor = (float)log((double)1.2f);
	f.pyramid.scaleFactors[0] = 1.f;
	for (int i = 1; i < 8; i++) f.pyramid.scaleFactors[i] = f.pyramid.scaleFactors[i - 1] * 1.2f;
	for (int i = 0; i < 8; i++) { f.pyramid.sigmaSq[i] = f.pyramid.scaleFactors[i] * f.pyramid.scaleFactors[i]; f.pyramid.invSigmaSq[i] = 1.f / f.pyramid.sigmaSq[i]; }
	f.imageBounds = { 0.f, 640.f, 0.f, 480.f };
	f.camera = { 517.3f, 516.5f, 318.6f, 255.3f, 40.f, 40.f / 517.3f };
	f.keypointsUn.resize(n); f.uright.resize(n);
	desc.resize((size_t)n * 32);
	for (int i = 0; i < n; i++)
	{
		const int oct = irand(10) < 5 ? 0 : irand(8);
		f.keypointsUn[i] = cv::KeyPoint(uni(16, 624), uni(16, 464), 31.f * f.pyramid.scaleFactors[oct], uni(0, 359.9f), 20.f, oct);
		f.uright[i] = uni(0, 1) < stereo_share ? f.keypointsUn[i].pt.x - uni(1, 60) : -1.f;
		for (int b = 0; b < 32; b++) desc[(size_t)i * 32 + b] = (unsigned char)rng();
	}
	f.descriptors = cv::Mat(n, 32, CV_8U, desc.data(), 32);
	f.mappoints.assign(n, nullptr);
	const float a = 0.1f;
	f.pose.R_ = cv::Matx33f(std::cos(a), -std::sin(a), 0.f, std::sin(a), std::cos(a), 0.f, 0.f, 0.f, 1.f);
	f.pose.t_ = cv::Matx31f(0.1f, -0.05f, 0.2f);
}

void noisy_copy(unsigned char* dst, const unsigned char* src, int flips)
{
	memcpy(dst, src, 32);
	for (int k = 0; k < flips; k++) { const int b = irand(256); dst[b >> 3] ^= (unsigned char)(1 << (b & 7)); }
}

oracle_frame_view view_of(const KeyFrame& f)
{
	oracle_frame_view v;
	v.n = f.N; v.kps_un = reinterpret_cast<const oracle_keypoint*>(f.keypointsUn.data()); v.desc = f.descriptors.data;
	v.uright = f.uright.data();
	v.bounds = { f.imageBounds.minx, f.imageBounds.maxx, f.imageBounds.miny, f.imageBounds.maxy };
	v.nlevels = 8; v.scale_factors = f.pyramid.scaleFactors.data();
	return v;
}
oracle_pose pose_of(const CameraPose& p)
{
	oracle_pose o;
	for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) o.R[i * 3 + j] = p.R_(i, j); o.t[i] = p.t_(i); }
	return o;
}

// a world point that the camera (R, t) of `kf` sees near keypoint j at depth z
cv::Matx31f world_point_at(const KeyFrame& kf, const CameraPose& pose, int j, float z, float jitter)
{
	const float u = kf.keypointsUn[j].pt.x + uni(-jitter, jitter), v = kf.keypointsUn[j].pt.y + uni(-jitter, jitter);
	const cv::Matx31f Xc((u - kf.camera.cx) / kf.camera.fx * z, (v - kf.camera.cy) / kf.camera.fy * z, z);
	return pose.R_.t() * (Xc - pose.t_);
}

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

}  // namespace

int main(int argc, char** argv)
{
	try
	{
		const int n = 1000, npts = 1400;
		ORB_SLAM2::b200::GuidedMatcher matcher(0.6f, true);
		int summary[4] = { 0, 0, 0, 0 };

		// ---- Fuse(keyframe, mappoints, th) and Fuse(keyframe, Scw, mappoints, th, replacePoints)
		for (int variant = 0; variant < 2; variant++)
		{
			KeyFrame kf;
			std::vector<unsigned char> kdesc;
			make_keyframe(kf, n, kdesc, 0.3f);
			ORB_SLAM2::b200::DeviceFrame dev(kf);
			const float s = variant ? 1.08f : 1.f;
			const Sim3 Scw(kf.pose.R_, s * kf.pose.t_, s);           // maps the s-times larger world onto the key frame
			std::vector<MapPoint> store(npts);
			std::vector<MapPoint*> list(npts);
			std::vector<oracle_sim3_point> opts(npts);
			std::vector<unsigned char> odesc((size_t)npts * 32);
			std::vector<int> log_got;
			const cv::Matx31f Ow = kf.GetCameraCenter();
			for (int i = 0; i < npts; i++)
			{
				MapPoint& m = store[i];
				m.id = i; m.actions = &log_got;
				const int j = irand(n);
				const cv::Matx31f X = world_point_at(kf, kf.pose, j, uni(1.f, 9.f), 1.5f);
				const cv::Matx31f d = X - Ow;
				const float dist = (float)cv::norm(d);
				m.worldPos = s * X;                                     // Scw.Map(s X) lands where pose maps X (camera coordinates scaled by s)
				m.maxDistance_ = s * dist * kf.pyramid.scaleFactors[std::min(7, kf.keypointsUn[j].octave + (irand(4) == 0))] * uni(0.95f, 1.05f);
				m.minDistance_ = m.maxDistance_ / kf.pyramid.scaleFactors[7];
				cv::Matx31f nrm = (-1.f / dist) * d + cv::Matx31f(uni(-0.4f, 0.4f), uni(-0.4f, 0.4f), uni(-0.4f, 0.4f));
				if (irand(8) == 0) nrm = -1.f * nrm;
				m.normal = -1.f * ((1.f / (float)cv::norm(nrm)) * nrm);   // from the camera towards the point, as PO
				m.nobs = irand(9);
				m.bad = irand(20) == 0;
				noisy_copy(m.desc, &kdesc[(size_t)j * 32], irand(70));
				memcpy(&odesc[(size_t)i * 32], m.desc, 32);
				list[i] = (variant == 0 && irand(15) == 0) ? nullptr : &m;
				opts[i] = { { m.worldPos(0), m.worldPos(1), m.worldPos(2) }, { m.normal(0), m.normal(1), m.normal(2) }, m.minDistance_, m.maxDistance_,
				            list[i] ? 1 : 0 };
			}
			std::vector<int32_t> kf_mp(n, -1), nobs(npts);
			std::vector<uint8_t> bad(npts), in_kf(npts, 0);
			for (int c = 0; c < n; c += 2)                              // half of the keypoints hold a map point of the list
			{
				int p;
				do p = irand(npts); while (in_kf[p]);
				kf_mp[c] = p; in_kf[p] = 1;
				kf.mappoints[c] = &store[p]; store[p].inKeyFrame = &kf; store[p].indexInKeyFrame = c;
			}
			for (int i = 0; i < npts; i++) { nobs[i] = store[i].nobs; bad[i] = store[i].bad; }
			const oracle_frame_view fv = view_of(kf);
			const oracle_camera ocam = { kf.camera.fx, kf.camera.fy, kf.camera.cx, kf.camera.cy, kf.camera.bf, kf.camera.baseline };
			std::vector<int32_t> log_want(3 * (size_t)npts + 3), replace_want(npts, -1);
			int nlog = 0, want_n, got_n;
			std::vector<MapPoint*> replace_got(npts, nullptr);
			if (variant == 0)
			{
				const oracle_pose op = pose_of(kf.pose);
				want_n = orc_fuse(&fv, &ocam, &op, kf.pyramid.logScaleFactor, kf.pyramid.invSigmaSq.data(), opts.data(), odesc.data(), npts, 3.f, kf_mp.data(),
				                  nobs.data(), bad.data(), in_kf.data(), log_want.data(), (int)log_want.size(), &nlog);
				got_n = matcher.Fuse(&kf, dev, list, 3.f);
			}
			else
			{
				oracle_sim3 os;
				for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) os.R[i * 3 + j] = Scw.R_(i, j); os.t[i] = Scw.t_(i); }
				os.s = Scw.s_;
				want_n = orc_fuse_sim3(&fv, &ocam, &os, kf.pyramid.logScaleFactor, opts.data(), odesc.data(), npts, 4.f, kf_mp.data(), nobs.data(), bad.data(),
				                       replace_want.data(), log_want.data(), (int)log_want.size(), &nlog);
				got_n = matcher.Fuse(&kf, dev, Scw, list, 4.f, replace_got);
			}
			if (want_n != got_n) { printf("Fuse variant %d: nfused %d vs %d\n", variant, got_n, want_n); return 1; }
			if ((int)log_got.size() != nlog || memcmp(log_got.data(), log_want.data(), sizeof(int) * (size_t)nlog) != 0) { printf("Fuse variant %d: mutations differ (%zu vs %d ints)\n", variant, log_got.size(), nlog); return 1; }
			for (int c = 0; c < n; c++)
				if ((kf.mappoints[c] ? kf.mappoints[c]->id : -1) != kf_mp[c]) { printf("Fuse variant %d: keyframe slot %d\n", variant, c); return 1; }
			for (int i = 0; i < npts; i++)
			{
				if (store[i].nobs != nobs[i] || store[i].bad != (bad[i] != 0)) { printf("Fuse variant %d: point %d state\n", variant, i); return 1; }
				if (variant == 1 && (replace_got[i] ? replace_got[i]->id : -1) != replace_want[i]) { printf("Fuse sim3: replacePoints[%d]\n", i); return 1; }
			}
			if (got_n < 50 || nlog < 60) { printf("Fuse variant %d: scene too thin (%d fused, %d log ints)\n", variant, got_n, nlog); return 1; }
			summary[variant] = got_n;
		}

		// ---- SearchBySim3
		{
			KeyFrame kf1, kf2;
			std::vector<unsigned char> d1, d2;
			make_keyframe(kf1, n, d1, 0.f);
			make_keyframe(kf2, n, d2, 0.f);
			kf2.pose.t_ = cv::Matx31f(-0.2f, 0.1f, 0.05f);
			ORB_SLAM2::b200::DeviceFrame dev1(kf1), dev2(kf2);
			const float a = 0.03f;
			const Sim3 S12(cv::Matx33f(1.f, 0.f, 0.f, 0.f, std::cos(a), -std::sin(a), 0.f, std::sin(a), std::cos(a)), cv::Matx31f(0.02f, -0.03f, 0.04f), 1.07f);
			const Sim3 S21 = S12.Inverse();
			std::vector<MapPoint> mp1(n), mp2(n);
			std::vector<oracle_kf_point> p1(n), p2(n);
			std::vector<unsigned char> od1((size_t)n * 32), od2((size_t)n * 32);
			std::vector<int> partner(n);
			for (int i = 0; i < n; i++) partner[i] = i;
			std::shuffle(partner.begin(), partner.end(), rng);
			for (int dir = 0; dir < 2; dir++)
			{
				KeyFrame& from = dir ? kf2 : kf1; KeyFrame& to = dir ? kf1 : kf2;
				std::vector<MapPoint>& mp = dir ? mp2 : mp1;
				std::vector<oracle_kf_point>& op = dir ? p2 : p1;
				std::vector<unsigned char>& od = dir ? od2 : od1;
				const std::vector<unsigned char>& dto = dir ? d1 : d2;
				const Sim3& Sback = dir ? S21 : S12;                  // target camera coordinates -> source camera coordinates
				for (int i = 0; i < n; i++)
				{
					int j = dir ? (int)(std::find(partner.begin(), partner.end(), i) - partner.begin()) : partner[i];
					if (irand(4) == 0) j = irand(n);                    // a quarter of the pairs disagree
					const float z = uni(1.5f, 8.f);
					const float u = to.keypointsUn[j].pt.x + uni(-1.5f, 1.5f), v = to.keypointsUn[j].pt.y + uni(-1.5f, 1.5f);
					const cv::Matx31f Xt((u - to.camera.cx) / to.camera.fx * z, (v - to.camera.cy) / to.camera.fy * z, z);
					const cv::Matx31f Xs = Sback.Map(Xt);
					MapPoint& m = mp[i];
					m.id = i;
					m.worldPos = from.pose.R_.t() * (Xs - from.pose.t_);
					m.maxDistance_ = (float)cv::norm(Xt) * to.pyramid.scaleFactors[std::min(7, to.keypointsUn[j].octave + (irand(5) == 0))] * uni(0.96f, 1.04f);
					m.minDistance_ = m.maxDistance_ / to.pyramid.scaleFactors[7];
					m.bad = irand(25) == 0;
					m.inKeyFrame = &from; m.indexInKeyFrame = i;
					noisy_copy(m.desc, &dto[(size_t)j * 32], irand(60));
					memcpy(&od[(size_t)i * 32], m.desc, 32);
					const bool present = irand(12) != 0;
					from.mappoints[i] = present ? &m : nullptr;
					op[i] = { { m.worldPos(0), m.worldPos(1), m.worldPos(2) }, m.minDistance_, m.maxDistance_, 0.f, present ? (m.bad ? 4 : 1) : 0 };
				}
			}
			std::vector<MapPoint*> matches12(n, nullptr);
			for (int i = 0; i < n; i++)
				if (irand(20) == 0) { matches12[i] = &mp2[(size_t)((i * 7) % n)]; p1[i].flags |= 2; }
			const oracle_frame_view v1 = view_of(kf1), v2 = view_of(kf2);
			const oracle_camera oc = { kf1.camera.fx, kf1.camera.fy, kf1.camera.cx, kf1.camera.cy, kf1.camera.bf, kf1.camera.baseline };
			const oracle_pose o1 = pose_of(kf1.pose), o2 = pose_of(kf2.pose);
			oracle_sim3 os;
			for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) os.R[i * 3 + j] = S12.R_(i, j); os.t[i] = S12.t_(i); }
			os.s = S12.s_;
			std::vector<int32_t> want12(n);
			const int want = orc_search_by_sim3(&v1, &oc, &o1, kf1.pyramid.logScaleFactor, &v2, &oc, &o2, kf2.pyramid.logScaleFactor, &os, 7.5f, p1.data(), od1.data(),
			                                    p2.data(), od2.data(), want12.data());
			const int got = matcher.SearchBySim3(&kf1, dev1, &kf2, dev2, matches12, S12, 7.5f);
			if (want != got) { printf("SearchBySim3 nfound %d vs %d\n", got, want); return 1; }
			for (int i = 0; i < n; i++)
				if ((matches12[i] ? matches12[i]->id : -1) != want12[i]) { printf("SearchBySim3 matches12[%d]\n", i); return 1; }
			if (got < 200) { printf("SearchBySim3: scene too thin (%d)\n", got); return 1; }
			summary[2] = got;
		}

		// ---- SearchForTriangulation
		{
			KeyFrame kf1, kf2;
			std::vector<unsigned char> d1, d2;
			make_keyframe(kf1, n, d1, 0.3f);
			make_keyframe(kf2, n, d2, 0.3f);
			kf1.pose.t_ = cv::Matx31f(-0.4f, 0.3f, 1.5f);               // its centre projects inside key frame 2's image
			MapPoint some;
			for (int i = 0; i < n; i += 2)
			{
				const int j = irand(n);
				kf2.keypointsUn[j] = kf1.keypointsUn[i];
				kf2.keypointsUn[j].pt.x += uni(-40, 40);
				kf2.keypointsUn[j].pt.y += uni(-2.5f, 2.5f) * kf2.pyramid.scaleFactors[kf2.keypointsUn[j].octave];
				kf2.keypointsUn[j].angle = std::fmod(kf1.keypointsUn[i].angle + 20.f + uni(-4, 4) + (irand(8) == 0 ? uni(0, 360) : 0.f), 360.f);
				noisy_copy(&d2[(size_t)j * 32], &d1[(size_t)i * 32], irand(55));
				const unsigned node = 1000 + 7 * (unsigned)irand(40);
				kf1.featureVector[node].push_back((unsigned)i);
				if (irand(10) != 0) kf2.featureVector[node].push_back((unsigned)j);
			}
			for (auto& nd : kf1.featureVector) std::sort(nd.second.begin(), nd.second.end());
			for (auto& nd : kf2.featureVector) { std::sort(nd.second.begin(), nd.second.end()); nd.second.erase(std::unique(nd.second.begin(), nd.second.end()), nd.second.end()); }
			std::vector<uint8_t> has1(n), has2(n);
			for (int i = 0; i < n; i++)
			{
				has1[i] = irand(5) == 0; has2[i] = irand(5) == 0;
				kf1.mappoints[i] = has1[i] ? &some : nullptr; kf2.mappoints[i] = has2[i] ? &some : nullptr;
			}
			ORB_SLAM2::b200::DeviceFrame dev1(kf1), dev2(kf2);
			cv::Mat F12(3, 3, CV_32F);
			const float Fv[9] = { 1e-8f, -3e-8f, 2e-5f, 2e-8f, 1e-8f, -1.f / 516.5f, -1e-5f, 1.f / 516.5f, 3e-4f };
			for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) F12.at<float>(r, c) = Fv[r * 3 + c];
			const cv::Matx31f Xc = kf2.pose.R_ * kf1.GetCameraCenter() + kf2.pose.t_;
			const float invZ = 1.f / Xc(2);
			const float ep[2] = { invZ * kf2.camera.fx * Xc(0) + kf2.camera.cx, invZ * kf2.camera.fy * Xc(1) + kf2.camera.cy };
			const FlatFv a1(kf1.featureVector), a2(kf2.featureVector);
			const oracle_feature_vector o1 = a1.view(), o2 = a2.view();
			const oracle_frame_view v1 = view_of(kf1), v2 = view_of(kf2);
			for (int only_stereo = 0; only_stereo < 2; only_stereo++)
			{
				std::vector<int32_t> want12(n);
				const int want = orc_search_for_triangulation(&v1, &o1, has1.data(), &v2, &o2, has2.data(), Fv, ep, kf2.pyramid.sigmaSq.data(), only_stereo, 1,
				                                              want12.data());
				std::vector<std::pair<size_t, size_t>> ids;
				const int got = matcher.SearchForTriangulation(&kf1, dev1, &kf2, dev2, F12, ids, only_stereo != 0);
				if (want != got) { printf("SearchForTriangulation(%d) nmatches %d vs %d\n", only_stereo, got, want); return 1; }
				size_t k = 0;
				for (int i = 0; i < n; i++)
					if (want12[i] >= 0)
					{
						if (k >= ids.size() || ids[k].first != (size_t)i || ids[k].second != (size_t)want12[i]) { printf("SearchForTriangulation(%d) matchIds[%zu]\n", only_stereo, k); return 1; }
						k++;
					}
				if (k != ids.size()) { printf("SearchForTriangulation(%d) matchIds size\n", only_stereo); return 1; }
				if (!only_stereo && got < 40) { printf("SearchForTriangulation: scene too thin (%d)\n", got); return 1; }
				if (!only_stereo) summary[3] = got;
			}
		}

		// ---- ORBVocabulary::loadFromTextFile / transform / score through std::map types shaped like DBoW2::BowVector / FeatureVector
		int bow_words = -1;
		if (argc > 1)
		{
			ORB_SLAM2::b200::ORBVocabulary voc;
			if (voc.loadFromTextFile("/nonexistent/voc.txt")) { printf("loadFromTextFile accepted a missing file\n"); return 1; }
			if (!voc.loadFromTextFile(argv[1])) { printf("loadFromTextFile rejected %s\n", argv[1]); return 1; }
			void* ov = orc_voc_load_text(argv[1]);
			if (!ov) { printf("oracle rejected the vocabulary\n"); return 1; }
			const int nf = 900;
			std::vector<unsigned char> fd((size_t)nf * 32);
			for (auto& b : fd) b = (unsigned char)rng();
			std::vector<cv::Mat> features;
			for (int i = 0; i < nf; i++) features.push_back(cv::Mat(1, 32, CV_8U, &fd[(size_t)i * 32], 32));
			std::map<unsigned, double> bv, bv2;
			std::map<unsigned, std::vector<unsigned>> fvec;
			voc.transform(features, bv, fvec, 2);
			std::vector<int32_t> wi(nf + 1), fs(nf + 2); std::vector<double> wv(nf + 1); std::vector<uint32_t> fnodes(nf + 1), fi(nf + 1);
			int32_t nfv = 0;
			const int nw = orc_bow_transform(ov, fd.data(), nf, 2, wi.data(), wv.data(), fnodes.data(), fs.data(), fi.data(), &nfv);
			if ((int)bv.size() != nw || (int)fvec.size() != nfv) { printf("BoW sizes %zu/%zu vs %d/%d\n", bv.size(), fvec.size(), nw, nfv); return 1; }
			int k = 0;
			for (const auto& e : bv) { if ((int)e.first != wi[k] || memcmp(&e.second, &wv[k], 8) != 0) { printf("BowVector entry %d\n", k); return 1; } k++; }
			k = 0;
			for (const auto& e : fvec)
			{
				if (e.first != fnodes[k] || (int)e.second.size() != fs[k + 1] - fs[k] || memcmp(e.second.data(), &fi[fs[k]], 4 * e.second.size()) != 0) { printf("FeatureVector node %d\n", k); return 1; }
				k++;
			}
			std::vector<cv::Mat> half(features.begin(), features.begin() + nf / 2);
			std::map<unsigned, std::vector<unsigned>> fv2;
			voc.transform(half, bv2, fv2, 2);
			std::vector<int32_t> wi2(nf + 1); std::vector<double> wv2(nf + 1);
			const int nw2 = orc_bow_transform(ov, fd.data(), nf / 2, 2, wi2.data(), wv2.data(), fnodes.data(), fs.data(), fi.data(), &nfv);
			const double want_s = orc_bow_score(ov, wi.data(), wv.data(), nw, wi2.data(), wv2.data(), nw2), got_s = voc.score(bv, bv2);
			if (memcmp(&want_s, &got_s, 8) != 0) { printf("score %.17g vs %.17g\n", got_s, want_s); return 1; }
			orc_voc_destroy(ov);
			bow_words = nw;
			if (nw < 100) { printf("vocabulary too thin (%d words)\n", nw); return 1; }
		}

		printf("OK fuse %d, fuse sim3 %d, sim3 %d, triangulation %d, bow words %d\n", summary[0], summary[1], summary[2], summary[3], bow_words);
		return 0;
	}
	catch (const cv::Exception& e)
	{
		printf("exception: %s\n", e.what());
		return 3;
	}
}
