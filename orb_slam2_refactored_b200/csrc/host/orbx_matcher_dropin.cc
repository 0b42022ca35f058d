// Out-of-line definitions with the reference's exact signatures, to be compiled INSIDE the reference tree in place of the
// two definitions at src/ORBmatcher.cc:72-247 and :1449-1457. It includes the reference's own
// include/ORBmatcher.h, so it is not built in this repository; tests/cpp/dropin_test.cc exercises the same code through
// include/orbx/ORBmatcher.h against the OpenCV shim.
#include "ORBmatcher.h"              // the reference's header: declares ComputeStereoMatches and class ORBmatcher
#include "orbx/ORBmatcher.h"

namespace ORB_SLAM2
{

void ComputeStereoMatches(
	const KeyPoints& keypointsL, const cv::Mat& descriptorsL, const Pyramid& pyramidL,
	const KeyPoints& keypointsR, const cv::Mat& descriptorsR, const Pyramid& pyramidR,
	const std::vector<float>& scaleFactors, const std::vector<float>& invScaleFactors, const CameraParams& camera,
	std::vector<float>& uright, std::vector<float>& depth)
{
	b200::ComputeStereoMatches(keypointsL, descriptorsL, pyramidL, keypointsR, descriptorsR, pyramidR, scaleFactors, invScaleFactors,
		camera, uright, depth);
}

int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b)
{
	return b200::DescriptorDistance(a, b);
}

} // namespace ORB_SLAM2
