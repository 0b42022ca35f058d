// C++ drop-in check: the reference-shaped classes of include/orbx/*.h (built here against the OpenCV shim) must give the
// same result as the CPU oracle. Usage: dropin_test <w> <h> <nfeatures>; reads a raw u8 image from stdin twice the size
// (left then right), prints "OK <n>" or a diagnostic. Run by tests/test_cpp_dropin.py.
#include <cstdio>
#include <cstring>
#include <vector>

#include "orbx/ORBextractor.h"
#include "orbx/ORBmatcher.h"

#define ORACLE_PREFIX orc_
#include "oracle_api.h"

struct Camera { float fx, fy, cx, cy, bf, baseline; };

int main(int argc, char** argv)
{
	if (argc < 4) return 2;
	const int w = atoi(argv[1]), h = atoi(argv[2]), nf = atoi(argv[3]);
	std::vector<unsigned char> buf((size_t)w * h * 2);
	if (fread(buf.data(), 1, buf.size(), stdin) != buf.size()) { printf("short read\n"); return 2; }
	cv::Mat left(h, w, CV_8U, buf.data()), right(h, w, CV_8U, buf.data() + (size_t)w * h);
	try
	{
		const ORB_SLAM2::ORBextractor::Parameters params(nf);
		ORB_SLAM2::ORBextractor exL(params), exR(params);
		ORB_SLAM2::KeyPoints kl, kr;
		cv::Mat dl, dr;
		exL.Extract(left, kl, dl);
		exR.Extract(right, kr, dr);

		// oracle
		void* o = orc_extractor_create(nf, 1.2f, 8, 20, 7);
		std::vector<oracle_keypoint> ok(nf + 512);
		std::vector<unsigned char> od((nf + 512) * 32);
		const int n = orc_extractor_extract(o, left.data, w, h, left.step, ok.data(), od.data(), nf + 512);
		if (n != (int)kl.size()) { printf("count %d vs %zu\n", n, kl.size()); return 1; }
		if (memcmp(ok.data(), kl.data(), sizeof(oracle_keypoint) * n) != 0) { printf("keypoints differ\n"); return 1; }
		for (int i = 0; i < n; i++)
			if (memcmp(od.data() + 32 * i, dl.ptr(i), 32) != 0) { printf("descriptor %d differs\n", i); return 1; }
		if (exL.GetLevels() != 8 || exL.GetScaleFactors().size() != 8 || exL.GetImagePyramid().size() != 8) { printf("getters\n"); return 1; }
		if (exL.GetImagePyramid()[3].cols != (int)lrintf(exL.GetInverseScaleFactors()[3] * w)) { printf("pyramid size\n"); return 1; }

		// stereo, both entry points
		const Camera cam = { 435.2047f, 435.2047f, 367.4517f, 252.2009f, 47.90639f, 47.90639f / 435.2047f };
		std::vector<float> u1, d1, u2, d2;
		ORB_SLAM2::b200::ComputeStereoMatches(kl, dl, exL.GetImagePyramid(), kr, dr, exR.GetImagePyramid(), exL.GetScaleFactors(),
			exL.GetInverseScaleFactors(), cam, u1, d1);
		ORB_SLAM2::b200::ComputeStereoMatches(exL, exR, kl.size(), cam, u2, d2);
		if (u1 != u2 || d1 != d2) { printf("stereo paths differ\n"); return 1; }
		int matched = 0;
		for (float v : d1) matched += v > 0;

		// DescriptorDistance and the brute-force scan
		const int dd = ORB_SLAM2::b200::DescriptorDistance(dl.row(0), dr.row(0));
		if (dd != orc_descriptor_distance(dl.ptr(0), dr.ptr(0))) { printf("distance\n"); return 1; }
		ORB_SLAM2::b200::Knn2Result r = ORB_SLAM2::b200::BruteForceKnn2(dl, dr, 0.8f);
		std::vector<int32_t> idx(n), match(n); std::vector<uint16_t> best(n), second(n);
		orc_knn2(dl.data, n, dr.data, dr.rows, 50, 0.8f, idx.data(), best.data(), second.data(), match.data(), 4);
		if (idx != r.idx || best != r.best || second != r.second || match != r.match) { printf("knn2 differs\n"); return 1; }

		// MapPoint::ComputeDistinctiveDescriptors: observation sets of 1..9 descriptors cut from the extracted rows, one call for all sets
		{
			std::vector<std::vector<cv::Mat>> sets;
			for (int p = 0, at = 0; at + 9 < n && p < 40; p++) { std::vector<cv::Mat> s; for (int i = 0; i < 1 + p % 9; i++) s.push_back(dl.row(at++)); sets.push_back(s); }
			const std::vector<int32_t> best = ORB_SLAM2::b200::ComputeDistinctiveDescriptors(sets);
			for (size_t p = 0; p < sets.size(); p++)
			{
				std::vector<uint8_t> rows(sets[p].size() * 32);
				for (size_t i = 0; i < sets[p].size(); i++) memcpy(rows.data() + 32 * i, sets[p][i].data, 32);
				if (best[p] != orc_distinctive_index(rows.data(), (int)sets[p].size())) { printf("distinctive descriptor of set %d differs\n", (int)p); return 1; }
			}
			if (ORB_SLAM2::b200::ComputeDistinctiveDescriptors(sets[8]) != (size_t)best[8]) { printf("single-set form differs\n"); return 1; }
		}

		// the cv::ORB-style call operator of upstream ORB_SLAM2 (mask ignored)
		{
			ORB_SLAM2::KeyPoints k2; cv::Mat d2m;
			exL(left, cv::Mat(), k2, d2m);
			if ((int)k2.size() != n || memcmp(d2m.data, od.data(), 32 * (size_t)n) != 0) { printf("operator() differs from Extract\n"); return 1; }
		}

		// error behaviour: too small an image is refused with an exception, not a crash; the handle survives it (plan cache)
		bool threw = false;
		try { cv::Mat tiny(90, 120, CV_8U); tiny.setTo(0); ORB_SLAM2::KeyPoints kt; cv::Mat dt; exL.Extract(tiny, kt, dt); } catch (const cv::Exception&) { threw = true; }
		if (!threw) { printf("no exception for undersized image\n"); return 1; }
		{
			ORB_SLAM2::KeyPoints k3; cv::Mat d3;
			exL.Extract(left, k3, d3);
			if ((int)k3.size() != n || memcmp(d3.data, od.data(), 32 * (size_t)n) != 0) { printf("extract after a refused size differs\n"); return 1; }
		}
		printf("OK %d keypoints, %d stereo matches\n", n, matched);
		return 0;
	}
	catch (const cv::Exception& e)
	{
		printf("exception: %s\n", e.what());
		return 3;
	}
}
