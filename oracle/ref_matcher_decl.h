// TEST INFRASTRUCTURE — CPU oracle, not the product.
// Declarations that stand in for include/ORBmatcher.h:39-54, include/Point.h:32 and
// include/CameraParameters.h:29-40 when the matcher hot-path text is compiled by line range (see
// oracle/Makefile, rule matcher_gen.cc). Namespace `refm` keeps the matcher's PATCH_SIZE/RoundUp
// apart from the extractor's identically named file-statics.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <limits>
#include <utility>
#include <vector>

#include <opencv2/core.hpp>

#define popcnt32 __builtin_popcount
#define popcnt64 __builtin_popcountll

namespace refm {

using KeyPoints = std::vector<cv::KeyPoint>;
using Pyramid = std::vector<cv::Mat>;

struct CameraParams { float fx, fy, cx, cy, bf, baseline; };

struct ORBmatcher
{
	static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
};

void ComputeStereoMatches(
	const KeyPoints& keypointsL, const cv::Mat& descriptorsL, const Pyramid& pyramidL,
	const KeyPoints& keypointsR, const cv::Mat& descriptorsR, const Pyramid& pyramidR,
	const std::vector<float>& scaleFactors, const std::vector<float>& invScaleFactors, const CameraParams& camera,
	std::vector<float>& uright, std::vector<float>& depth);

size_t DistinctiveIndex(const std::vector<cv::Mat>& descriptors);
void ConvertToGrayRef(const cv::Mat& src, cv::Mat& dst, bool RGB);
void UndistortKeyPointsRef(const KeyPoints& src, KeyPoints& dst, const cv::Mat& K, const cv::Mat1f& distCoeffs);
void ComputeStereoFromRGBDRef(const KeyPoints& keypoints, const KeyPoints& keypointsUn, const cv::Mat& depthImage,
	const CameraParams& camera, std::vector<float>& uright, std::vector<float>& depth);

// The reference indexes distIndices[0] even when nothing matched (src/ORBmatcher.cc:232-233). With
// nL > 0 that is a harmless read of reserved storage; with nL == 0 it is a null dereference, so that
// one case is refused here.
inline int ComputeStereoMatchesGuarded(
	const KeyPoints& keypointsL, const cv::Mat& descriptorsL, const Pyramid& pyramidL,
	const KeyPoints& keypointsR, const cv::Mat& descriptorsR, const Pyramid& pyramidR,
	const std::vector<float>& scaleFactors, const std::vector<float>& invScaleFactors, const CameraParams& camera,
	std::vector<float>& uright, std::vector<float>& depth)
{
	if (keypointsL.empty())
	{
		uright.clear();
		depth.clear();
		return -1;
	}
	ComputeStereoMatches(keypointsL, descriptorsL, pyramidL, keypointsR, descriptorsR, pyramidR,
	                     scaleFactors, invScaleFactors, camera, uright, depth);
	return 0;
}

}  // namespace refm
