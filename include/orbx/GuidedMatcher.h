// Drop-in host side for the guided matchers of the reference, forwarding to the C ABI (include/orbx.h):
//
//   FeaturesGrid::AssignFeatures / GetFeaturesInArea                       include/Frame.h:61-81,    src/Frame.cc:63-145
//   int ORBmatcher::SearchByProjection(Frame&, const std::vector<MapPoint*>&, float th)
//                                                                          include/ORBmatcher.h:58,  src/ORBmatcher.cc:315-382
//   int ORBmatcher::SearchByProjection(Frame& currFrame, const Frame& lastFrame, float th, bool monocular)
//                                                                          include/ORBmatcher.h:62,  src/ORBmatcher.cc:1279-1362
//   int ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const std::set<MapPoint*>& alreadyFound, float th, int ORBdist)
//                                                                          include/ORBmatcher.h:66,  src/ORBmatcher.cc:1364-1447
//   int ORBmatcher::SearchByBoW(KeyFrame*, Frame&, std::vector<MapPoint*>&) / (KeyFrame*, KeyFrame*, std::vector<MapPoint*>&)
//                                                                          include/ORBmatcher.h:72-73, src/ORBmatcher.cc:406-516, 696-766
//   int ORBmatcher::SearchForInitialization(Frame&, Frame&, std::vector<cv::Point2f>&, std::vector<int>&, int windowSize)
//                                                                          include/ORBmatcher.h:80,  src/ORBmatcher.cc:614-694
//   int ORBmatcher::Fuse(KeyFrame*, const std::vector<MapPoint*>&, float th) / Fuse(KeyFrame*, const Sim3&, ..., replacePoints)
//                                                                          src/ORBmatcher.cc:868-980, 982-1088
//   int ORBmatcher::SearchBySim3(KeyFrame*, KeyFrame*, std::vector<MapPoint*>&, const Sim3&, float th)   src/ORBmatcher.cc:1090-1277
//   int ORBmatcher::SearchForTriangulation(const KeyFrame*, const KeyFrame*, const cv::Mat& F12, matchIds, bool onlyStereo)
//                                                                          src/ORBmatcher.cc:768-866
//
// The methods are templates over the reference's own Frame and MapPoint types: they read exactly the members the reference code
// reads (frame.keypointsUn, .descriptors, .uright, .mappoints, .outlier, .keypoints, .imageBounds, .pyramid.scaleFactors, .camera,
// .pose; mappoint->trackInView, trackScaleLevel, trackViewCos, trackProjX/Y/XR, isBad(), Observations(), GetDescriptor(),
// GetWorldPos()) and write exactly what it writes (frame.mappoints, prevMatched, matches12). The sequential "already matched" state
// of the loops is reproduced on the GPU, so the results are identical, not approximately equal. No CPU fallback.
#ifndef ORBX_GUIDEDMATCHER_H
#define ORBX_GUIDEDMATCHER_H

#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include <opencv2/core.hpp>

#include "../orbx.h"
#include "ORBmatcher.h"

namespace ORB_SLAM2
{
namespace b200
{

// A Frame's matcher-visible data on the GPU, with its FeaturesGrid. Construct it once per Frame (or Assign() a long-lived one to each
// new Frame: the device buffers are reused) and pass it next to the Frame.
class DeviceFrame
{
public:
	template <class FrameT> explicit DeviceFrame(const FrameT& frame, int device = 0) : h_(nullptr)
	{
		Storage s;
		const orbx_frame_view v = View(frame, s);
		Check(orbx_frame_create(&v, device, &h_), "DeviceFrame");
	}
	template <class FrameT> void Assign(const FrameT& frame)
	{
		Storage s;
		const orbx_frame_view v = View(frame, s);
		Check(orbx_frame_assign(h_, &v), "DeviceFrame::Assign");
	}
	~DeviceFrame() { orbx_frame_destroy(h_); }
	DeviceFrame(const DeviceFrame&) = delete;
	DeviceFrame& operator=(const DeviceFrame&) = delete;

	// FeaturesGrid::GetFeaturesInArea (src/Frame.cc:102-145), same output order. One window per call is a kernel launch; the
	// matchers below never call it — they enumerate all their windows on the device.
	std::vector<size_t> GetFeaturesInArea(float x, float y, float r, int minLevel = -1, int maxLevel = -1) const
	{
		const float xyr[3] = { x, y, r };
		const int32_t lv[2] = { minLevel, maxLevel };
		int32_t off[2] = { 0, 0 };
		std::vector<int32_t> idx(256);
		orbx_status st = orbx_frame_features_in_area(h_, xyr, lv, 1, off, idx.data(), (int)idx.size());
		if (st == ORBX_ERR_CAPACITY)
		{
			idx.resize((size_t)off[1]);
			st = orbx_frame_features_in_area(h_, xyr, lv, 1, off, idx.data(), (int)idx.size());
		}
		Check(st, "GetFeaturesInArea");
		return std::vector<size_t>(idx.begin(), idx.begin() + off[1]);
	}

	orbx_frame Handle() const { return h_; }

private:
	struct Storage { cv::Mat desc; };
	template <class FrameT> static orbx_frame_view View(const FrameT& frame, Storage& s)
	{
		static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint layout");
		orbx_frame_view v;
		v.n = static_cast<int32_t>(frame.keypointsUn.size());
		v.kps_un = reinterpret_cast<const orbx_keypoint*>(frame.keypointsUn.data());
		s.desc = frame.descriptors;
		if (v.n > 0 && s.desc.step != 32) { cv::Mat c; s.desc.copyTo(c); s.desc = c; }
		v.desc = s.desc.data;
		v.uright = frame.uright.size() == frame.keypointsUn.size() && v.n > 0 ? frame.uright.data() : nullptr;
		v.bounds.minx = frame.imageBounds.minx; v.bounds.maxx = frame.imageBounds.maxx;
		v.bounds.miny = frame.imageBounds.miny; v.bounds.maxy = frame.imageBounds.maxy;
		v.nlevels = static_cast<int32_t>(frame.pyramid.scaleFactors.size());
		v.scale_factors = frame.pyramid.scaleFactors.data();
		return v;
	}
	orbx_frame h_;
};

class GuidedMatcher
{
public:
	GuidedMatcher(float nnratio = 0.6f, bool checkOri = true) : fNNRatio_(nnratio), checkOrientation_(checkOri) {}

	// src/ORBmatcher.cc:315-382
	template <class FrameT, class MapPointT>
	int SearchByProjection(FrameT& frame, DeviceFrame& dev, const std::vector<MapPointT*>& mappoints, float th = 3.f) const
	{
		const int npts = static_cast<int>(mappoints.size());
		std::vector<orbx_track_point> pts((size_t)npts);
		std::vector<uint8_t> desc((size_t)npts * 32);
		for (int i = 0; i < npts; i++)
		{
			const MapPointT* m = mappoints[i];
			orbx_track_point& p = pts[i];
			const bool active = m->trackInView && !m->isBad();                       // :321-322
			p.proj_x = m->trackProjX; p.proj_y = m->trackProjY; p.proj_xr = m->trackProjXR;
			p.view_cos = m->trackViewCos; p.scale_level = m->trackScaleLevel;
			p.flags = (active ? 1 : 0) | (m->Observations() > 0 ? 2 : 0);
			if (active) std::memcpy(&desc[(size_t)i * 32], m->GetDescriptor().data, 32);
		}
		std::vector<int32_t> state = Encode(frame);
		int nmatches = 0;
		Check(orbx_search_by_projection_local_map(dev.Handle(), state.data(), pts.data(), desc.data(), npts, th, fNNRatio_, &nmatches),
			"SearchByProjection");
		for (size_t c = 0; c < state.size(); c++)
			if (state[c] >= 0) frame.mappoints[c] = mappoints[(size_t)state[c]];           // :375
		return nmatches;
	}

	// src/ORBmatcher.cc:1279-1362
	template <class FrameT>
	int SearchByProjection(FrameT& currFrame, DeviceFrame& devCurr, const FrameT& lastFrame, float th, bool monocular) const
	{
		const int npts = lastFrame.N;
		std::vector<orbx_last_point> pts((size_t)npts);
		std::vector<uint8_t> desc((size_t)npts * 32);
		for (int i = 0; i < npts; i++)
		{
			orbx_last_point& p = pts[i];
			const auto* m = lastFrame.mappoints[i];
			p.flags = 0;
			p.xw[0] = p.xw[1] = p.xw[2] = 0.f;
			p.octave = lastFrame.keypoints[i].octave;                                 // :1313
			p.angle = lastFrame.keypointsUn[i].angle;                                 // CheckOrientation, :273
			if (m && !lastFrame.outlier[i])                                           // :1295-1297
			{
				const auto Xw = m->GetWorldPos();
				p.xw[0] = Xw(0); p.xw[1] = Xw(1); p.xw[2] = Xw(2);
				p.flags = 1 | (m->Observations() > 0 ? 2 : 0);
				std::memcpy(&desc[(size_t)i * 32], m->GetDescriptor().data, 32);
			}
		}
		const auto& cam = currFrame.camera;
		const orbx_camera ocam = { cam.fx, cam.fy, cam.cx, cam.cy, cam.bf, cam.baseline };
		const orbx_pose cp = Pose(currFrame.pose), lp = Pose(lastFrame.pose);
		std::vector<int32_t> state = Encode(currFrame);
		int nmatches = 0;
		Check(orbx_search_by_projection_last_frame(devCurr.Handle(), &ocam, &cp, &lp, state.data(), pts.data(), desc.data(), npts, th,
			monocular ? 1 : 0, checkOrientation_ ? 1 : 0, &nmatches), "SearchByProjection");
		for (size_t c = 0; c < state.size(); c++)
		{
			if (state[c] >= 0) currFrame.mappoints[c] = lastFrame.mappoints[(size_t)state[c]];   // :1351
			else if (state[c] == -1) currFrame.mappoints[c] = nullptr;                           // CheckOrientation, :301
		}
		return nmatches;
	}

	// src/ORBmatcher.cc:1364-1447 (relocalisation). The geometry is evaluated here with the reference's own classes — cv::Matx
	// arithmetic, cv::norm, mappoint->Get{Min,Max}DistanceInvariance(), mappoint->PredictScale(dist, &frame) — so it is the
	// reference's arithmetic by construction; the window search, its "already matched" state and CheckOrientation run on the GPU.
	template <class FrameT, class KeyFrameT, class SetT>
	int SearchByProjection(FrameT& frame, DeviceFrame& dev, KeyFrameT* keyframe, const SetT& alreadyFound, float th, int ORBdist) const
	{
		const auto mappoints = keyframe->GetMapPointMatches();
		const int npts = static_cast<int>(mappoints.size());
		std::vector<orbx_window> win((size_t)npts);
		std::vector<uint8_t> desc((size_t)npts * 32);
		const cv::Matx33f Rcw = frame.pose.R();
		const cv::Matx31f tcw = frame.pose.t();
		const cv::Matx31f Ow = frame.GetCameraCenter();
		for (int i = 0; i < npts; i++)
		{
			orbx_window& w = win[i];
			w.u = w.v = w.radius = 0.f; w.min_level = w.max_level = 0; w.flags = 0;
			w.angle = keyframe->keypointsUn[i].angle;
			auto* mappoint = mappoints[i];
			if (!mappoint || mappoint->isBad() || alreadyFound.count(mappoint)) continue;       // :1379-1381
			const cv::Matx31f Xw = mappoint->GetWorldPos();
			const cv::Matx31f Xc = Rcw * Xw + tcw;                                              // CameraProjection::WorldToImage
			const float invZ = 1.f / Xc(2);
			const float u = invZ * frame.camera.fx * Xc(0) + frame.camera.cx, v = invZ * frame.camera.fy * Xc(1) + frame.camera.cy;
			if (!(u >= frame.imageBounds.minx && u < frame.imageBounds.maxx && v >= frame.imageBounds.miny && v < frame.imageBounds.maxy)) continue;
			const cv::Matx31f PO = Xw - Ow;
			const float dist3D = static_cast<float>(cv::norm(PO));
			if (dist3D < mappoint->GetMinDistanceInvariance() || dist3D > mappoint->GetMaxDistanceInvariance()) continue;   // :1400-1401
			const int predictedScale = mappoint->PredictScale(dist3D, &frame);
			w.u = u; w.v = v;
			w.radius = th * frame.pyramid.scaleFactors[predictedScale];                          // :1406
			w.min_level = predictedScale - 1; w.max_level = predictedScale + 1;                  // :1408
			w.flags = 1;
			std::memcpy(&desc[(size_t)i * 32], mappoint->GetDescriptor().data, 32);
		}
		std::vector<int32_t> state(frame.mappoints.size());
		for (size_t c = 0; c < state.size(); c++) state[c] = frame.mappoints[c] ? -2 : -1;       // :1412-1413: any map point closes
		int nmatches = 0;
		Check(orbx_search_windows(dev.Handle(), state.data(), win.data(), desc.data(), npts, ORBdist, checkOrientation_ ? 1 : 0, &nmatches),
			"SearchByProjection");
		for (size_t c = 0; c < state.size(); c++)
		{
			if (state[c] >= 0) frame.mappoints[c] = mappoints[(size_t)state[c]];                 // :1426
			else if (state[c] == -1) frame.mappoints[c] = nullptr;                               // CheckOrientation, :301
		}
		return nmatches;
	}

	// src/ORBmatcher.cc:614-694
	template <class FrameT>
	int SearchForInitialization(FrameT& frame1, DeviceFrame& dev1, FrameT& frame2, DeviceFrame& dev2, std::vector<cv::Point2f>& prevMatched,
		std::vector<int>& matches12, int windowSize = 10) const
	{
		(void)frame2;
		const size_t n1 = frame1.keypointsUn.size();
		static_assert(sizeof(cv::Point2f) == 8, "cv::Point2f layout");
		matches12.assign(n1, -1);
		if (n1 == 0) return 0;
		int nmatches = 0;
		Check(orbx_search_for_initialization(dev1.Handle(), dev2.Handle(), reinterpret_cast<float*>(prevMatched.data()), matches12.data(), windowSize,
			fNNRatio_, checkOrientation_ ? 1 : 0, &nmatches), "SearchForInitialization");
		return nmatches;
	}

	// src/ORBmatcher.cc:452-516 — KeyFrameT needs featureVector (DBoW2::FeatureVector, i.e. std::map<NodeId, std::vector<unsigned>>),
	// GetMapPointMatches(); FrameT needs featureVector. dev1 / dev2 hold keyframe's and frame's keypointsUn + descriptors.
	template <class KeyFrameT, class FrameT, class MapPointT>
	int SearchByBoW(KeyFrameT* keyframe, DeviceFrame& dev1, FrameT& frame, DeviceFrame& dev2, std::vector<MapPointT*>& matches) const
	{
		const std::vector<MapPointT*> mappoints1 = keyframe->GetMapPointMatches();
		const size_t n2 = frame.keypointsUn.size();
		matches.assign(n2, nullptr);                                                          // :456
		std::vector<uint8_t> valid1(mappoints1.size());
		for (size_t i = 0; i < valid1.size(); i++) valid1[i] = mappoints1[i] && !mappoints1[i]->isBad();
		std::vector<int32_t> match2(n2 ? n2 : 1);
		const int nmatches = Bow(keyframe->featureVector, valid1, dev1, frame.featureVector, nullptr, dev2, match2);
		for (size_t c = 0; c < n2; c++)
			if (match2[c] >= 0) matches[c] = mappoints1[(size_t)match2[c]];                     // :503
		return nmatches;
	}

	// src/ORBmatcher.cc:696-766
	template <class KeyFrameT, class MapPointT>
	int SearchByBoW(KeyFrameT* keyframe1, DeviceFrame& dev1, KeyFrameT* keyframe2, DeviceFrame& dev2, std::vector<MapPointT*>& matches12) const
	{
		const std::vector<MapPointT*> mappoints1 = keyframe1->GetMapPointMatches(), mappoints2 = keyframe2->GetMapPointMatches();
		matches12.assign(mappoints1.size(), nullptr);                                         // :707
		std::vector<uint8_t> valid1(mappoints1.size()), valid2(mappoints2.size());
		for (size_t i = 0; i < valid1.size(); i++) valid1[i] = mappoints1[i] && !mappoints1[i]->isBad();
		for (size_t i = 0; i < valid2.size(); i++) valid2[i] = mappoints2[i] && !mappoints2[i]->isBad();
		std::vector<int32_t> match2(valid2.size() ? valid2.size() : 1);
		const int nmatches = Bow(keyframe1->featureVector, valid1, dev1, keyframe2->featureVector, &valid2, dev2, match2);
		for (size_t c = 0; c < valid2.size(); c++)
			if (match2[c] >= 0) matches12[(size_t)match2[c]] = mappoints2[c];                   // :752
		return nmatches;
	}

	// ---- Matchers of local mapping and loop closing whose per-point search reads nothing the loop changes -----------------------------
	// src/ORBmatcher.cc:868-980. The geometry of every point is evaluated here with the reference's own classes (:879-919), all window
	// searches with their octave and chi-square gates run in one kernel, and the loop of the reference is then replayed from :876 and
	// :956-976 over the results — the map calls (Replace, AddObservation, AddMapPoint) are the reference's own, in its order.
	template <class KeyFrameT, class MapPointT>
	int Fuse(KeyFrameT* keyframe, DeviceFrame& dev, const std::vector<MapPointT*>& mappoints, float th = 3.f) const
	{
		const int npts = static_cast<int>(mappoints.size());
		std::vector<orbx_best_window> win((size_t)npts);
		std::vector<uint8_t> desc((size_t)npts * 32);
		const auto pose = keyframe->GetPose();
		const cv::Matx33f Rcw = pose.R();
		const cv::Matx31f tcw = pose.t();
		const cv::Matx31f Ow = keyframe->GetCameraCenter();
		for (int i = 0; i < npts; i++)
		{
			orbx_best_window& w = win[i];
			w = orbx_best_window{ 0.f, 0.f, 0.f, 0.f, 0, 0, 0 };
			MapPointT* mappoint = mappoints[i];
			if (!mappoint || mappoint->isBad()) continue;
			const cv::Matx31f Xw = mappoint->GetWorldPos();
			const cv::Matx31f Xc = Rcw * Xw + tcw;
			if (Xc(2) < 0.f) continue;
			const float invZ = 1.f / Xc(2);
			const float u = invZ * keyframe->camera.fx * Xc(0) + keyframe->camera.cx, v = invZ * keyframe->camera.fy * Xc(1) + keyframe->camera.cy;
			if (!keyframe->IsInImage(u, v)) continue;
			const float ur = u - keyframe->camera.bf / Xc(2);
			const cv::Matx31f PO = Xw - Ow;
			const float dist3D = static_cast<float>(cv::norm(PO));
			if (dist3D < mappoint->GetMinDistanceInvariance() || dist3D > mappoint->GetMaxDistanceInvariance()) continue;
			const cv::Matx31f Pn = mappoint->GetNormal();
			if (PO.dot(Pn) < 0.5 * dist3D) continue;
			const int predictedScale = mappoint->PredictScale(dist3D, keyframe);
			w.u = u; w.v = v; w.ur = ur;
			w.radius = th * keyframe->pyramid.scaleFactors[predictedScale];
			w.min_level = predictedScale - 1; w.max_level = predictedScale;              // :930
			w.flags = 3;                                                                 // with the chi-square gate of :934-945
			std::memcpy(&desc[(size_t)i * 32], mappoint->GetDescriptor().data, 32);
		}
		std::vector<int32_t> bestIdx((size_t)npts + 1), bestDist((size_t)npts + 1);
		Check(orbx_search_best_in_windows(dev.Handle(), win.data(), desc.data(), npts, keyframe->pyramid.invSigmaSq.data(), bestIdx.data(),
			bestDist.data()), "Fuse");
		int nfused = 0;
		for (int i = 0; i < npts; i++)
		{
			MapPointT* mappoint = mappoints[i];
			if (!mappoint || mappoint->isBad() || mappoint->IsInKeyFrame(keyframe)) continue;    // :876, with the state the loop has produced so far
			if (!(win[i].flags & 1) || bestDist[i] > 50) continue;                               // TH_LOW, :957
			MapPointT* MPInKF = keyframe->GetMapPoint(bestIdx[i]);
			if (MPInKF)
			{
				if (!MPInKF->isBad())
				{
					if (MPInKF->Observations() > mappoint->Observations()) mappoint->Replace(MPInKF);
					else MPInKF->Replace(mappoint);
				}
			}
			else
			{
				mappoint->AddObservation(keyframe, bestIdx[i]);
				keyframe->AddMapPoint(mappoint, bestIdx[i]);
			}
			nfused++;
		}
		return nfused;
	}

	// src/ORBmatcher.cc:982-1088 (loop closing)
	template <class KeyFrameT, class Sim3T, class MapPointT>
	int Fuse(KeyFrameT* keyframe, DeviceFrame& dev, const Sim3T& Scw, const std::vector<MapPointT*>& mappoints, float th,
		std::vector<MapPointT*>& replacePoints) const
	{
		const int npts = static_cast<int>(mappoints.size());
		std::vector<orbx_best_window> win((size_t)npts);
		std::vector<uint8_t> desc((size_t)npts * 32);
		const cv::Matx33f Rcw = Scw.R();
		const cv::Matx31f tcw = Scw.Invs() * Scw.t();                                    // pose(Scw.R(), Scw.Invs() * Scw.t()), :987
		const cv::Matx31f Ow = -Rcw.t() * tcw;                                           // pose.Invt()
		const auto alreadyFound = keyframe->GetMapPoints();                              // :992
		for (int i = 0; i < npts; i++)
		{
			orbx_best_window& w = win[i];
			w = orbx_best_window{ 0.f, 0.f, 0.f, 0.f, 0, 0, 0 };
			MapPointT* mappoint = mappoints[i];
			if (mappoint->isBad() || alreadyFound.count(mappoint)) continue;
			const cv::Matx31f Xw = mappoint->GetWorldPos();
			const cv::Matx31f Xc = Rcw * Xw + tcw;
			if (Xc(2) < 0.f) continue;
			const float invZ = 1.f / Xc(2);
			const float u = invZ * keyframe->camera.fx * Xc(0) + keyframe->camera.cx, v = invZ * keyframe->camera.fy * Xc(1) + keyframe->camera.cy;
			if (!keyframe->IsInImage(u, v)) continue;
			const cv::Matx31f PO = Xw - Ow;
			const float dist3D = static_cast<float>(cv::norm(PO));
			if (dist3D < mappoint->GetMinDistanceInvariance() || dist3D > mappoint->GetMaxDistanceInvariance()) continue;
			const cv::Matx31f Pn = mappoint->GetNormal();
			if (PO.dot(Pn) < 0.5 * dist3D) continue;
			const int predictedScale = mappoint->PredictScale(dist3D, keyframe);
			w.u = u; w.v = v;
			w.radius = th * keyframe->pyramid.scaleFactors[predictedScale];
			w.min_level = predictedScale - 1; w.max_level = predictedScale;              // :1057
			w.flags = 1;
			std::memcpy(&desc[(size_t)i * 32], mappoint->GetDescriptor().data, 32);
		}
		std::vector<int32_t> bestIdx((size_t)npts + 1), bestDist((size_t)npts + 1);
		Check(orbx_search_best_in_windows(dev.Handle(), win.data(), desc.data(), npts, nullptr, bestIdx.data(), bestDist.data()), "Fuse");
		int nfused = 0;
		for (int i = 0; i < npts; i++)
		{
			MapPointT* mappoint = mappoints[i];
			if (mappoint->isBad() || alreadyFound.count(mappoint)) continue;             // :1002 (nothing in this loop makes a point bad)
			if (!(win[i].flags & 1) || bestDist[i] > 50) continue;                       // TH_LOW, :1070
			MapPointT* MPInKF = keyframe->GetMapPoint(bestIdx[i]);
			if (MPInKF)
			{
				if (!MPInKF->isBad()) replacePoints[i] = MPInKF;
			}
			else
			{
				mappoint->AddObservation(keyframe, bestIdx[i]);
				keyframe->AddMapPoint(mappoint, bestIdx[i]);
			}
			nfused++;
		}
		return nfused;
	}

	// src/ORBmatcher.cc:1090-1277 (loop closing): both directed searches on the GPU, the agreement test here
	template <class KeyFrameT, class Sim3T, class MapPointT>
	int SearchBySim3(KeyFrameT* keyframe1, DeviceFrame& dev1, KeyFrameT* keyframe2, DeviceFrame& dev2, std::vector<MapPointT*>& matches12,
		const Sim3T& S12, float th) const
	{
		const Sim3T S21 = S12.Inverse();
		const auto mappoints1 = keyframe1->GetMapPointMatches();
		const auto mappoints2 = keyframe2->GetMapPointMatches();
		const int N1 = static_cast<int>(mappoints1.size()), N2 = static_cast<int>(mappoints2.size());
		std::vector<bool> alreadyMatched1((size_t)N1, false), alreadyMatched2((size_t)N2, false);
		for (int i = 0; i < N1; i++)                                                     // :1109-1121
		{
			MapPointT* mappoint = matches12[i];
			if (mappoint)
			{
				alreadyMatched1[i] = true;
				const int idx2 = mappoint->GetIndexInKeyFrame(keyframe2);
				if (idx2 >= 0 && idx2 < N2) alreadyMatched2[idx2] = true;
			}
		}
		auto direction = [&](KeyFrameT* from, const std::vector<MapPointT*>& mappoints, const std::vector<bool>& already, const Sim3T& S,
			KeyFrameT* to, DeviceFrame& devTo, std::vector<int32_t>& match)
		{
			const int n = static_cast<int>(mappoints.size());
			std::vector<orbx_best_window> win((size_t)n);
			std::vector<uint8_t> desc((size_t)n * 32);
			const auto pose = from->GetPose();
			const cv::Matx33f R = pose.R();
			const cv::Matx31f t = pose.t();
			for (int i = 0; i < n; i++)
			{
				orbx_best_window& w = win[i];
				w = orbx_best_window{ 0.f, 0.f, 0.f, 0.f, 0, 0, 0 };
				MapPointT* mappoint = mappoints[i];
				if (!mappoint || already[i] || mappoint->isBad()) continue;              // :1130, :1197
				const cv::Matx31f Xw = mappoint->GetWorldPos();
				const cv::Matx31f Xa = R * Xw + t;
				const cv::Matx31f Xb = S.Map(Xa);
				if (Xb(2) < 0.f) continue;
				const float invZ = 1.f / Xb(2);
				const float u = invZ * to->camera.fx * Xb(0) + to->camera.cx, v = invZ * to->camera.fy * Xb(1) + to->camera.cy;
				if (!to->IsInImage(u, v)) continue;
				const float dist3D = static_cast<float>(cv::norm(Xb));
				if (dist3D < mappoint->GetMinDistanceInvariance() || dist3D > mappoint->GetMaxDistanceInvariance()) continue;
				const int predictedScale = mappoint->PredictScale(dist3D, to);
				w.u = u; w.v = v;
				w.radius = th * to->pyramid.scaleFactors[predictedScale];
				w.min_level = predictedScale - 1; w.max_level = predictedScale;          // :1175, :1242
				w.flags = 1;
				std::memcpy(&desc[(size_t)i * 32], mappoint->GetDescriptor().data, 32);
			}
			std::vector<int32_t> bestDist((size_t)n + 1);
			match.assign((size_t)n + 1, -1);
			Check(orbx_search_best_in_windows(devTo.Handle(), win.data(), desc.data(), n, nullptr, match.data(), bestDist.data()), "SearchBySim3");
			for (int i = 0; i < n; i++)
				if (bestDist[i] > 100) match[i] = -1;                                    // TH_HIGH, :1187, :1254
		};
		std::vector<int32_t> match1, match2;
		direction(keyframe1, mappoints1, alreadyMatched1, S21, keyframe2, dev2, match1);
		direction(keyframe2, mappoints2, alreadyMatched2, S12, keyframe1, dev1, match2);
		int nfound = 0;
		for (int i1 = 0; i1 < N1; i1++)                                                  // :1260-1274
		{
			const int idx2 = match1[i1];
			if (idx2 >= 0 && match2[idx2] == i1)
			{
				matches12[i1] = mappoints2[idx2];
				nfound++;
			}
		}
		return nfound;
	}

	// src/ORBmatcher.cc:768-866 (local mapping). F12 is the 3x3 CV_32F matrix the reference passes; the epipole is computed as it does.
	template <class KeyFrameT>
	int SearchForTriangulation(const KeyFrameT* keyframe1, DeviceFrame& dev1, const KeyFrameT* keyframe2, DeviceFrame& dev2, const cv::Mat& F12,
		std::vector<std::pair<size_t, size_t>>& matchIds, bool onlyStereo) const
	{
		const auto pose2 = keyframe2->GetPose();
		const cv::Matx31f Xc = pose2.R() * keyframe1->GetCameraCenter() + pose2.t();     // proj2.WorldToImage(keyframe1->GetCameraCenter()), :772-773
		const float invZ = 1.f / Xc(2);
		const float ep[2] = { invZ * keyframe2->camera.fx * Xc(0) + keyframe2->camera.cx, invZ * keyframe2->camera.fy * Xc(1) + keyframe2->camera.cy };
		float F[9];
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++) F[r * 3 + c] = F12.at<float>(r, c);
		std::vector<uint8_t> has1((size_t)keyframe1->N), has2((size_t)keyframe2->N);
		for (int i = 0; i < keyframe1->N; i++) has1[i] = keyframe1->GetMapPoint(i) != nullptr;
		for (int i = 0; i < keyframe2->N; i++) has2[i] = keyframe2->GetMapPoint(i) != nullptr;
		const FlatFeatureVector a = Flatten(keyframe1->featureVector), b = Flatten(keyframe2->featureVector);
		const orbx_feature_vector va = a.view(), vb = b.view();
		std::vector<int32_t> matches12((size_t)keyframe1->N + 1);
		int nmatches = 0;
		Check(orbx_search_for_triangulation(dev1.Handle(), &va, has1.data(), dev2.Handle(), &vb, has2.data(), F, ep, keyframe2->pyramid.sigmaSq.data(),
			onlyStereo ? 1 : 0, checkOrientation_ ? 1 : 0, matches12.data(), &nmatches), "SearchForTriangulation");
		matchIds.clear();
		matchIds.reserve((size_t)nmatches);
		for (int idx1 = 0; idx1 < keyframe1->N; idx1++)                                  // :859-863
			if (matches12[idx1] >= 0) matchIds.push_back(std::make_pair((size_t)idx1, (size_t)matches12[idx1]));
		return nmatches;
	}

private:
	struct FlatFeatureVector
	{
		std::vector<uint32_t> ids, indices;
		std::vector<int32_t> start;
		orbx_feature_vector view() const { return orbx_feature_vector{ (int32_t)ids.size(), ids.data(), start.data(), indices.data() }; }
	};
	template <class FeatureVectorT> static FlatFeatureVector Flatten(const FeatureVectorT& fv)
	{
		FlatFeatureVector f;
		f.start.push_back(0);
		for (const auto& node : fv)      // std::map: ascending node id
		{
			f.ids.push_back((uint32_t)node.first);
			for (auto i : node.second) f.indices.push_back((uint32_t)i);
			f.start.push_back((int32_t)f.indices.size());
		}
		return f;
	}
	template <class FeatureVectorT>
	int Bow(const FeatureVectorT& fv1, const std::vector<uint8_t>& valid1, DeviceFrame& dev1, const FeatureVectorT& fv2,
		const std::vector<uint8_t>* valid2, DeviceFrame& dev2, std::vector<int32_t>& match2) const
	{
		const FlatFeatureVector a = Flatten(fv1), b = Flatten(fv2);
		const orbx_feature_vector va = a.view(), vb = b.view();
		int nmatches = 0;
		Check(orbx_search_by_bow(dev1.Handle(), &va, valid1.data(), dev2.Handle(), &vb, valid2 ? valid2->data() : nullptr, fNNRatio_,
			checkOrientation_ ? 1 : 0, match2.data(), &nmatches), "SearchByBoW");
		return nmatches;
	}

	// frame.mappoints as the C ABI's codes: -1 null, -2 / -3 a map point with / without observations
	template <class FrameT> static std::vector<int32_t> Encode(const FrameT& frame)
	{
		std::vector<int32_t> code(frame.mappoints.size());
		for (size_t c = 0; c < code.size(); c++)
			code[c] = !frame.mappoints[c] ? -1 : (frame.mappoints[c]->Observations() > 0 ? -2 : -3);
		return code;
	}
	template <class PoseT> static orbx_pose Pose(const PoseT& pose)
	{
		orbx_pose p;
		for (int i = 0; i < 3; i++)
		{
			for (int j = 0; j < 3; j++) p.R[i * 3 + j] = pose.R()(i, j);
			p.t[i] = pose.t()(i);
		}
		return p;
	}

	float fNNRatio_;
	bool checkOrientation_;
};

} // namespace b200
} // namespace ORB_SLAM2

#endif
