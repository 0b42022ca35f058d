// TEST INFRASTRUCTURE — CPU oracle, not the product.
// cv:: entry points of the shim, forwarding to the pinned restatements in cv_primitives.cc.
#include <opencv2/core.hpp>

#include "cv_primitives.h"

namespace cv {

void resize(const Mat& src, Mat& dst, Size dsize)
{
	CV_Assert(!src.empty() && dsize.width > 0 && dsize.height > 0);
	Mat out;
	out.create(dsize.height, dsize.width, CV_8U);
	cvp::resize_linear_u8(src.data, src.cols, src.rows, src.step, out.data, out.cols, out.rows, out.step);
	dst = out;
}

void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression)
{
	std::vector<cvp::FastPoint> pts;
	cvp::fast9_16(image.data, image.cols, image.rows, image.step, threshold, nonmaxSuppression, pts);
	keypoints.clear();
	keypoints.reserve(pts.size());
	for (const cvp::FastPoint& p : pts)
		keypoints.push_back(KeyPoint((float)p.x, (float)p.y, 7.f, -1.f, (float)p.score));
}

void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int borderType)
{
	// the only configuration the reference uses (ORBextractor.cc:799)
	CV_Assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2.0 && (sigmaY == 2.0 || sigmaY == 0.0));
	CV_Assert(borderType == BORDER_REFLECT_101);
	Mat out;
	out.create(src.rows, src.cols, CV_8U);
	cvp::gauss7x7_u8(src.data, src.cols, src.rows, src.step, out.data, out.step);
	dst = out;
}

float fastAtan2(float y, float x) { return cvp::fast_atan2_deg(y, x); }

void cvtColor(const Mat& src, Mat& dst, int code)
{
	CV_Assert(code == COLOR_BGR2GRAY || code == COLOR_RGB2GRAY || code == COLOR_BGRA2GRAY || code == COLOR_RGBA2GRAY);
	const int ch = (code == COLOR_BGR2GRAY || code == COLOR_RGB2GRAY) ? 3 : 4;
	CV_Assert(src.channels() == ch);
	Mat out;
	out.create(src.rows, src.cols, CV_8U);
	cvp::cvt_gray_u8(src.data, src.cols, src.rows, src.step, ch, code == COLOR_RGB2GRAY || code == COLOR_RGBA2GRAY, out.data, out.step);
	dst = out;
}

void undistortPoints(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Mat& K, const Mat& distCoeffs, const Mat& R, const Mat& P)
{
	CV_Assert(K.type() == CV_32F && K.rows == 3 && K.cols == 3 && distCoeffs.type() == CV_32F && R.empty());
	CV_Assert(P.data == K.data || (P.rows == 3 && P.cols == 3));
	const int nd = distCoeffs.rows * distCoeffs.cols;
	std::vector<Point2f> out(src.size());
	cvp::undistort_points(reinterpret_cast<const float*>(src.data()), (int)src.size(), K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2),
	                      K.at<float>(1, 2), distCoeffs.ptr<float>(), nd, reinterpret_cast<float*>(out.data()));
	dst = out;
}

void remap(const Mat& src, Mat& dst, const Mat& map1, const Mat& map2, int interpolation)
{
	CV_Assert(interpolation == INTER_LINEAR && src.type() == CV_8U && map1.type() == CV_32F && map2.type() == CV_32F);
	CV_Assert(map1.rows == map2.rows && map1.cols == map2.cols && map1.step == map2.step);
	Mat out;
	out.create(map1.rows, map1.cols, CV_8U);
	cvp::remap_linear_u8(src.data, src.cols, src.rows, src.step, map1.ptr<float>(), map2.ptr<float>(), map1.step, out.data, out.cols, out.rows, out.step);
	dst = out;
}

}  // namespace cv
