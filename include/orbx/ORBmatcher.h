// Drop-in definitions for the two hot-path symbols of the reference's src/ORBmatcher.cc (header-only helpers; the
// out-of-line definitions with the reference's exact signatures are in orb_slam2_refactored_b200/csrc/host/orbx_matcher_dropin.cc):
//
//   void ORB_SLAM2::ComputeStereoMatches(...)            include/ORBmatcher.h:41-45, src/ORBmatcher.cc:72-247
//   int  ORB_SLAM2::ORBmatcher::DescriptorDistance(a, b) include/ORBmatcher.h:54,    src/ORBmatcher.cc:1449-1457
//
// plus the brute-force best/second scan that the reference only has as an inner loop (src/ORBmatcher.cc:477-507).
// Everything forwards to the C ABI (include/orbx.h); no CPU fallback.
#ifndef ORBX_ORBMATCHER_H
#define ORBX_ORBMATCHER_H

#include <cstdint>
#include <string>
#include <cstring>
#include <vector>

#include <opencv2/core.hpp>

#include "../orbx.h"
#include "ORBextractor.h"

namespace ORB_SLAM2
{
namespace b200
{

using Pyramid = std::vector<cv::Mat>;   // include/ORBmatcher.h:39

inline void Check(orbx_status st, const char* where)
{
	if (st != ORBX_OK)
		throw cv::Exception(cv::Error::StsError, orbx_last_error(), where, __FILE__, __LINE__);
}

// The reference's argument list, host data (one stereo pair). CameraT needs fx, fy, cx, cy, bf, baseline (include/CameraParameters.h:29-40).
template <class CameraT>
inline void ComputeStereoMatches(
	const KeyPoints& keypointsL, const cv::Mat& descriptorsL, const Pyramid& pyramidL,
	const KeyPoints& keypointsR, const cv::Mat& descriptorsR, const Pyramid& pyramidR,
	const std::vector<float>& scaleFactors, const std::vector<float>& invScaleFactors, const CameraT& camera,
	std::vector<float>& uright, std::vector<float>& depth, int device = 0)
{
	const int nL = static_cast<int>(keypointsL.size()), nR = static_cast<int>(keypointsR.size());
	uright.assign(nL, -1.f);
	depth.assign(nL, -1.f);
	if (nL == 0)
		return;     // the reference dereferences an empty vector here (src/ORBmatcher.cc:232-233)
	const int nlevels = static_cast<int>(pyramidL.size());
	std::vector<const uint8_t*> pl(nlevels), pr(nlevels);
	std::vector<int> lw(nlevels), lh(nlevels);
	std::vector<size_t> lp(nlevels);
	std::vector<cv::Mat> tmpR(nlevels);
	for (int s = 0; s < nlevels; s++)
	{
		pl[s] = pyramidL[s].data; lw[s] = pyramidL[s].cols; lh[s] = pyramidL[s].rows; lp[s] = pyramidL[s].step;
		if (pyramidR[s].step == pyramidL[s].step) pr[s] = pyramidR[s].data;
		else { tmpR[s].create(lh[s], lw[s], CV_8U); pyramidR[s].copyTo(tmpR[s]); pr[s] = tmpR[s].data; lp[s] = tmpR[s].step; pl[s] = pyramidL[s].data; }
	}
	// descriptors must be continuous N x 32 (Extract creates them so, :785)
	cv::Mat dl = descriptorsL, dr = descriptorsR;
	if (dl.step != 32) { cv::Mat c; dl.copyTo(c); dl = c; }
	if (dr.step != 32) { cv::Mat c; dr.copyTo(c); dr = c; }
	const orbx_camera cam = { camera.fx, camera.fy, camera.cx, camera.cy, camera.bf, camera.baseline };
	Check(orbx_stereo_match_host(device, reinterpret_cast<const orbx_keypoint*>(keypointsL.data()), nL, dl.data, pl.data(),
		reinterpret_cast<const orbx_keypoint*>(keypointsR.data()), nR, dr.data, pr.data(), lw.data(), lh.data(), lp.data(), nlevels,
		scaleFactors.data(), invScaleFactors.data(), &cam, uright.data(), depth.data()), "ComputeStereoMatches");
}

// Device-resident variant for the call site at src/System.cc:458-461: both extractors just ran Extract (:449-452), so
// keypoints, descriptors and pyramids are already in HBM and nothing but uright/depth crosses PCIe.
template <class CameraT>
inline void ComputeStereoMatches(const ORBextractor& extractorL, const ORBextractor& extractorR, size_t nkeypointsL, const CameraT& camera,
	std::vector<float>& uright, std::vector<float>& depth)
{
	const orbx_camera cam = { camera.fx, camera.fy, camera.cx, camera.cy, camera.bf, camera.baseline };
	int frames = 0, cap = 0;     // the result arrays are laid out like the keypoints of the last extract: frames x cap
	Check(orbx_last_result_shape(extractorL.Handle(), &frames, &cap), "ComputeStereoMatches");
	std::vector<float> u(static_cast<size_t>(frames) * cap), d(u.size());
	Check(orbx_stereo_match(extractorL.Handle(), extractorR.Handle(), &cam, u.data(), d.data()), "ComputeStereoMatches");
	if (nkeypointsL > static_cast<size_t>(cap)) nkeypointsL = static_cast<size_t>(cap);
	uright.assign(u.begin(), u.begin() + nkeypointsL);
	depth.assign(d.begin(), d.begin() + nkeypointsL);
}

// The N x N Hamming site of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:286-314) for MANY map points in one call: sets[p] are
// the observation descriptors of map point p (1 x 32 CV_8U rows); the result is, per map point, the index of the descriptor with the least
// median distance to the others (first on ties, as the reference's strict `<`), -1 for an empty set. This — not a DescriptorDistance call
// per pair, which would be a GPU round trip per pair — is what replaces the loop at :289-299.
inline std::vector<int32_t> ComputeDistinctiveDescriptors(const std::vector<std::vector<cv::Mat>>& sets, int device = 0)
{
	std::vector<int64_t> offsets(sets.size() + 1, 0);
	for (size_t p = 0; p < sets.size(); p++) offsets[p + 1] = offsets[p] + static_cast<int64_t>(sets[p].size());
	std::vector<uint8_t> rows(static_cast<size_t>(offsets.back()) * 32 + 32);
	for (size_t p = 0; p < sets.size(); p++)
		for (size_t i = 0; i < sets[p].size(); i++)
			std::memcpy(rows.data() + (static_cast<size_t>(offsets[p]) + i) * 32, sets[p][i].data, 32);
	std::vector<int32_t> best(sets.size(), -1);
	if (!sets.empty())
		Check(orbx_distinctive_descriptors(device, rows.data(), offsets.data(), static_cast<int>(sets.size()), best.data()), "ComputeDistinctiveDescriptors");
	return best;
}
// one map point: the reference's own shape (descriptors of its observations in, index of the representative out)
inline size_t ComputeDistinctiveDescriptors(const std::vector<cv::Mat>& descriptors, int device = 0)
{
	const std::vector<int32_t> best = ComputeDistinctiveDescriptors(std::vector<std::vector<cv::Mat>>(1, descriptors), device);
	return best[0] < 0 ? 0 : static_cast<size_t>(best[0]);
}

// ORBmatcher::DescriptorDistance: 1 x 32 CV_8U row headers. One pair per call costs a host-device round trip; loops over pairs
// (MapPoint::ComputeDistinctiveDescriptors, :289-299) use the batched forms above / orbx_descriptor_distance with n pairs.
inline int DescriptorDistance(const cv::Mat& a, const cv::Mat& b, int device = 0)
{
	int32_t dist = 0;
	Check(orbx_descriptor_distance(device, a.data, b.data, 1, &dist), "DescriptorDistance");
	return dist;
}

struct Knn2Result { std::vector<int32_t> idx, match; std::vector<uint16_t> best, second; };

// Every row of `query` (N x 32) against every row of `train` (M x 32): the inner loop of SearchByBoW (src/ORBmatcher.cc:477-507)
// with its acceptance test best <= thLow && best < nnratio * second. thLow defaults to TH_LOW (:42).
inline Knn2Result BruteForceKnn2(const cv::Mat& query, const cv::Mat& train, float nnratio = 0.6f, int thLow = 50, int device = 0)
{
	Knn2Result r;
	const int64_t nq = query.rows, nt = train.rows;
	r.idx.resize(nq); r.match.resize(nq); r.best.resize(nq); r.second.resize(nq);
	if (nq == 0) return r;
	Check(orbx_knn2(device, query.data, nq, train.data, nt, thLow, nnratio, r.idx.data(), r.best.data(), r.second.data(), r.match.data()), "BruteForceKnn2");
	return r;
}

} // namespace b200
} // namespace ORB_SLAM2

#endif
