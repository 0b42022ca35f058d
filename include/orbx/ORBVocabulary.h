// Drop-in host side for the part of ORB_SLAM2::ORBVocabulary (include/ORBVocabulary.h: DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>)
// that the front-end calls, forwarding to the C ABI (include/orbx.h) — the tree lives on the GPU:
//
//   bool loadFromTextFile(const std::string& filename)                    Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.cpp:21-90
//   void transform(const std::vector<cv::Mat>& features, BowVector& v, FeatureVector& fv, int levelsup) const
//                                                                          Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1129-1197
//   double score(const BowVector& a, const BowVector& b) const            Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1201-1206 (L1Scoring)
//   bool empty() const, unsigned size() const                             :1003-1015
//
// transform and score are templates over the reference's own DBoW2::BowVector (a std::map<WordId, WordValue>) and DBoW2::FeatureVector
// (a std::map<NodeId, std::vector<unsigned>>), so that Frame::ComputeBoW (src/Frame.cc:208-214) keeps its one line:
//     voc->transform(Converter::toDescriptorVector(descriptors), bowVector, featureVector, 4);
// Word values are bit-equal to DBoW2's (FP64, accumulated and normalised in its order). No CPU fallback.
#ifndef ORBX_ORBVOCABULARY_H
#define ORBX_ORBVOCABULARY_H

#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include <opencv2/core.hpp>

#include "../orbx.h"
#include "ORBmatcher.h"

namespace ORB_SLAM2
{
namespace b200
{

class ORBVocabulary
{
public:
	explicit ORBVocabulary(int device = 0) : device_(device), h_(nullptr) {}
	~ORBVocabulary() { orbx_vocabulary_destroy(h_); }
	ORBVocabulary(const ORBVocabulary&) = delete;
	ORBVocabulary& operator=(const ORBVocabulary&) = delete;

	// false where the reference returns false (unreadable file, header outside its accepted ranges)
	bool loadFromTextFile(const std::string& filename)
	{
		orbx_vocabulary_destroy(h_);
		h_ = nullptr;
		const orbx_status st = orbx_vocabulary_load_text(filename.c_str(), device_, &h_);
		if (st == ORBX_ERR_INVALID) return false;
		Check(st, "ORBVocabulary::loadFromTextFile");
		return true;
	}

	unsigned int size() const
	{
		int64_t words = 0;
		if (h_) Check(orbx_vocabulary_info(h_, nullptr, nullptr, nullptr, nullptr, nullptr, &words), "ORBVocabulary::size");
		return static_cast<unsigned int>(words);
	}
	bool empty() const { return size() == 0; }

	// features: 1 x 32 CV_8U rows, what Converter::toDescriptorVector produces
	template <class BowVectorT, class FeatureVectorT>
	void transform(const std::vector<cv::Mat>& features, BowVectorT& v, FeatureVectorT& fv, int levelsup) const
	{
		v.clear();
		fv.clear();
		const int n = static_cast<int>(features.size());
		if (!h_ || n == 0) return;
		std::vector<uint8_t> desc((size_t)n * 32);
		for (int i = 0; i < n; i++) std::memcpy(&desc[(size_t)i * 32], features[i].data, 32);
		std::vector<int32_t> ids((size_t)n), start((size_t)n + 1);
		std::vector<double> vals((size_t)n);
		std::vector<uint32_t> nodes((size_t)n), items((size_t)n);
		int32_t nw = 0, nn = 0;
		Check(orbx_bow_transform(h_, desc.data(), n, levelsup, ids.data(), vals.data(), &nw, nodes.data(), start.data(), items.data(), &nn, nullptr,
			nullptr), "ORBVocabulary::transform");
		for (int k = 0; k < nw; k++)
			v.insert(v.end(), typename BowVectorT::value_type(static_cast<typename BowVectorT::key_type>(ids[k]), vals[k]));
		for (int k = 0; k < nn; k++)
			fv.insert(fv.end(), typename FeatureVectorT::value_type(static_cast<typename FeatureVectorT::key_type>(nodes[k]),
				typename FeatureVectorT::mapped_type(items.begin() + start[k], items.begin() + start[k + 1])));
	}

	// L1Scoring::score (ScoringObject.cpp:24-58), the scoring of the reference's vocabulary
	template <class BowVectorT>
	double score(const BowVectorT& a, const BowVectorT& b) const
	{
		std::vector<int32_t> ids;
		std::vector<double> vals;
		int32_t off[3] = { 0, 0, 0 };
		for (const auto& e : a) { ids.push_back((int32_t)e.first); vals.push_back(e.second); }
		off[1] = (int32_t)ids.size();
		for (const auto& e : b) { ids.push_back((int32_t)e.first); vals.push_back(e.second); }
		off[2] = (int32_t)ids.size();
		if (ids.empty()) { ids.push_back(0); vals.push_back(0.0); }
		const int32_t pa = 0, pb = 1;
		double s = 0.0;
		Check(orbx_bow_score_l1(h_, ids.data(), vals.data(), off, &pa, &pb, 1, &s), "ORBVocabulary::score");
		return s;
	}

	orbx_vocabulary Handle() const { return h_; }

private:
	int device_;
	orbx_vocabulary h_;
};

} // namespace b200
} // namespace ORB_SLAM2

#endif
