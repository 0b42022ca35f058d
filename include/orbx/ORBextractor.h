// Drop-in replacement for the reference's include/ORBextractor.h + src/ORBextractor.cc (header-only).
//
// Same namespace, class name, nested Parameters struct, constructor, Init(), Extract() and getters as
// /root/reference/include/ORBextractor.h:34-80, so src/System.cc:378-385, :449-450, :509, :567 and :139-148 compile
// unchanged. Everything is forwarded to the C ABI (include/orbx.h -> liborbx_b200.so, hand-written sm_100a kernels).
// Builds against real OpenCV (<opencv2/core.hpp>) or against oracle/cvshim where OpenCV is absent.
//
// Error behaviour mirrors the reference: Extract throws cv::Exception where the reference's CV_Assert fires
// (src/ORBextractor.cc:457) and — instead of the reference's undefined behaviour — also for images outside its
// input contract (level smaller than 62 px, portrait aspect) and for CUDA failures. There is no CPU fallback.
#ifndef ORBX_ORBEXTRACTOR_H
#define ORBX_ORBEXTRACTOR_H

#include <string>
#include <vector>

#include <opencv2/core.hpp>

#include "../orbx.h"

namespace ORB_SLAM2
{

#ifndef POINT_H
using KeyPoints = std::vector<cv::KeyPoint>;   // include/Point.h:32
#endif

class ORBextractor
{
public:

	struct Parameters   // include/ORBextractor.h:38-47
	{
		int nfeatures;
		float scaleFactor;
		int nlevels;
		int iniThFAST;
		int minThFAST;

		Parameters(int nfeatures = 2000, float scaleFactor = 1.2f, int nlevels = 8, int iniThFAST = 20, int minThFAST = 7)
			: nfeatures(nfeatures), scaleFactor(scaleFactor), nlevels(nlevels), iniThFAST(iniThFAST), minThFAST(minThFAST) {}
	};

	explicit ORBextractor(const Parameters& param, int device = 0) : param_(param), device_(device), handle_(nullptr) { Init(); }
	~ORBextractor() { if (handle_) orbx_destroy(handle_); }
	ORBextractor(const ORBextractor&) = delete;
	ORBextractor& operator=(const ORBextractor&) = delete;

	// src/ORBextractor.cc:697-741
	void Init()
	{
		if (handle_) { orbx_destroy(handle_); handle_ = nullptr; }
		const orbx_params p = { param_.nfeatures, param_.scaleFactor, param_.nlevels, param_.iniThFAST, param_.minThFAST };
		Check(orbx_create(&p, device_, &handle_), "ORBextractor::Init");
		const size_t n = static_cast<size_t>(param_.nlevels);
		scaleFactors_.resize(n); invScaleFactors_.resize(n); sigmaSq_.resize(n); invSigmaSq_.resize(n);
		Check(orbx_scale_tables(handle_, scaleFactors_.data(), invScaleFactors_.data(), sigmaSq_.data(), invSigmaSq_.data()), "scale tables");
	}

	// Compute the ORB features and descriptors on an image (src/ORBextractor.cc:743-820).
	void Extract(const cv::Mat& image, KeyPoints& keypoints, cv::Mat& descriptors)
	{
		if (image.type() != CV_8U)
			throw cv::Exception(cv::Error::StsAssert, "image.type() == CV_8U", "ORB_SLAM2::b200::ORBextractor::Extract", __FILE__, __LINE__);
		static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint must be 7 x 4 bytes");
		const int cap = orbx_max_keypoints(handle_);
		kpScratch_.resize(static_cast<size_t>(cap));
		descScratch_.create(cap, 32, CV_8U);
		int n = 0;
		Check(orbx_extract(handle_, image.data, image.cols, image.rows, image.step,
			reinterpret_cast<orbx_keypoint*>(kpScratch_.data()), descScratch_.data, cap, &n), "ORBextractor::Extract");
		pyramidValid_ = false;
		if (n == 0)
		{
			descriptors.release();   // keypoints deliberately left untouched, as the reference does (:778-782)
			return;
		}
		keypoints.assign(kpScratch_.begin(), kpScratch_.begin() + n);
		descriptors.create(n, 32, CV_8U);
		for (int i = 0; i < n; i++)
			std::memcpy(descriptors.ptr(i), descScratch_.ptr(i), 32);
	}

	// cv::ORB-style call operator of upstream ORB_SLAM2 (ORBextractor::operator()(image, mask, keypoints, descriptors)); this fork
	// renamed it to Extract and dropped the mask, which upstream ignores as well.
	template <class MaskT>
	void operator()(const cv::Mat& image, const MaskT& /*mask*/, KeyPoints& keypoints, cv::Mat& descriptors) { Extract(image, keypoints, descriptors); }

	int GetLevels() const { return param_.nlevels; }
	float GetScaleFactor() const { return param_.scaleFactor; }
	const std::vector<float>& GetScaleFactors() const { return scaleFactors_; }
	const std::vector<float>& GetInverseScaleFactors() const { return invScaleFactors_; }
	const std::vector<float>& GetScaleSigmaSquares() const { return sigmaSq_; }
	const std::vector<float>& GetInverseScaleSigmaSquares() const { return invSigmaSq_; }

	// The pyramid of the last Extract lives on the GPU; it is downloaded on the first request only (the reference's one
	// consumer is ComputeStereoMatches, which has a device-resident overload in orbx/ORBmatcher.h).
	const std::vector<cv::Mat>& GetImagePyramid() const
	{
		if (!pyramidValid_)
		{
			images_.resize(static_cast<size_t>(param_.nlevels));
			for (int s = 0; s < param_.nlevels; s++)
			{
				int w = 0, h = 0;
				Check(orbx_level_size(handle_, s, &w, &h), "GetImagePyramid");
				images_[s].create(h, w, CV_8U);
				Check(orbx_pyramid_level(handle_, 0, s, images_[s].data, images_[s].step), "GetImagePyramid");
			}
			pyramidValid_ = true;
		}
		return images_;
	}

	orbx_handle Handle() const { return handle_; }   // for the device-resident stereo matcher

private:

	static void Check(orbx_status st, const char* where)
	{
		if (st != ORBX_OK)
			throw cv::Exception(cv::Error::StsError, orbx_last_error(), where, __FILE__, __LINE__);
	}

	std::vector<float> scaleFactors_, invScaleFactors_, sigmaSq_, invSigmaSq_;
	mutable std::vector<cv::Mat> images_;
	mutable bool pyramidValid_ = false;
	KeyPoints kpScratch_;
	cv::Mat descScratch_;
	Parameters param_;
	int device_;
	orbx_handle handle_;
};

} // namespace ORB_SLAM2

#endif
