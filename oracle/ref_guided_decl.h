// TEST INFRASTRUCTURE — CPU oracle, not the product.
//
// Stand-ins for the object graph the guided matchers walk, so that the reference text of
//   src/Frame.cc:32-34, 39-59, 63-145          (Round*, ImageBounds, FeaturesGrid)
//   include/Frame.h:40-81                      (ImageBounds, ScalePyramidInfo, FeaturesGrid declarations)
//   src/ORBmatcher.cc:37-58, 249-309           (constants, CheckOrientation)
//   src/ORBmatcher.cc:315-382                  SearchByProjection(Frame&, mappoints, th)       — local-map tracking
//   src/ORBmatcher.cc:406-516, 696-766         FeatureVectorIterator, SearchByBoW x2            — reference keyframe / relocalisation / loop
//   src/ORBmatcher.cc:518-612                  SearchByProjection(keyframe, Scw, points, matched) — loop closing
//   src/ORBmatcher.cc:614-694                  SearchForInitialization                          — monocular initialisation
//   src/ORBmatcher.cc:1279-1362                SearchByProjection(currFrame, lastFrame, th, m)  — motion-model tracking
//   src/ORBmatcher.cc:1364-1447                SearchByProjection(frame, keyframe, found, th, d) — relocalisation
//   src/ORBmatcher.cc:384-404, 768-866         CheckDistEpipolarLine, SearchForTriangulation      — local mapping
//   src/ORBmatcher.cc:868-980, 982-1088        Fuse x2                                            — local mapping, loop closing
//   src/ORBmatcher.cc:1090-1277                SearchBySim3                                       — loop closing
//   src/MapPoint.cc:382-392, 405-414; src/Frame.cc:203-206   distance invariance, PredictScale, GetCameraCenter
// compiles by line range (oracle/Makefile, rule guided_gen.cc) with the reference's own include/Point.h,
// include/CameraParameters.h, include/CameraPose.h and include/CameraProjection.h. The real Frame/MapPoint/KeyFrame drag
// in DBoW2, the vocabulary and the map; only the members those three functions touch exist here, under the same names
// (include/Frame.h:83-168, include/MapPoint.h:49-97).
//
// This header is included INSIDE namespace ORB_SLAM2 of the generated TU, after the Frame.h declarations.
#pragma once

// src/MapPoint.cc:29 takes a std::mutex here; the stand-in map points are touched by one thread
#define LOCK_MUTEX_POSITION()

struct Frame;
struct KeyFrame;

struct MapPoint
{
	// tracking scratch (include/MapPoint.h:92-97)
	float trackProjX = 0.f, trackProjY = 0.f, trackProjXR = 0.f;
	bool trackInView = false;
	int trackScaleLevel = 0;
	float trackViewCos = 0.f;

	// what the accessors return (include/MapPoint.h:49,55,64,75)
	Point3D worldPos;
	int nobs = 0;
	bool bad = false;
	cv::Mat descriptor;     // 1 x 32 CV_8U row header

	// scale-invariance distances (include/MapPoint.h; bodies from src/MapPoint.cc:382-392, 405-414 by line range)
	float minDistance_ = 0.f, maxDistance_ = 0.f;
	float GetMinDistanceInvariance() const;
	float GetMaxDistanceInvariance() const;
	int PredictScale(float dist, const Frame* frame) const;
	int PredictScale(float dist, const KeyFrame* keyframe) const;   // src/MapPoint.cc:394-403
	Vec3D normal;
	Vec3D GetNormal() const { return normal; }                      // src/MapPoint.cc:97-101

	// what Fuse and SearchBySim3 call besides (include/MapPoint.h:52-60, 70-72): a one-key-frame model of the observation graph, enough
	// to drive every branch of src/ORBmatcher.cc:876, 956-976, 1069-1084, 1109-1121; each mutation is appended to `actions` as
	// (kind, this point, other point or keypoint index): 1 = this->Replace(other), 2 = this->AddObservation(kf, idx)
	int id = -1;
	const KeyFrame* inKeyFrame = nullptr;
	int indexInKeyFrame = -1;
	std::vector<int>* actions = nullptr;
	bool IsInKeyFrame(const KeyFrame* kf) const { return inKeyFrame == kf; }
	int GetIndexInKeyFrame(const KeyFrame* kf) const { return inKeyFrame == kf ? indexInKeyFrame : -1; }
	void AddObservation(KeyFrame* kf, size_t idx);
	void Replace(MapPoint* other);

	Point3D GetWorldPos() const { return worldPos; }
	int Observations() const { return nobs; }
	bool isBad() const { return bad; }
	cv::Mat GetDescriptor() const { return descriptor; }
};


struct Frame
{
	CameraParams camera;
	int N = 0;
	KeyPoints keypoints;
	KeyPoints keypointsUn;
	std::vector<float> uright;
	cv::Mat descriptors;
	std::vector<MapPoint*> mappoints;
	std::vector<bool> outlier;
	FeaturesGrid grid;
	CameraPose pose;
	ScalePyramidInfo pyramid;
	ImageBounds imageBounds;
	DBoW2::FeatureVector featureVector;

	Point3D GetCameraCenter() const;   // src/Frame.cc:203-206 by line range

	// src/Frame.cc:216-219
	std::vector<size_t> GetFeaturesInArea(float x, float y, float r, int minLevel = -1, int maxLevel = -1) const
	{
		return grid.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
	}
};

// include/KeyFrame.h:78, 110, 141-148: what SearchByBoW reads of a KeyFrame
struct KeyFrame
{
	int N = 0;
	KeyPoints keypointsUn;
	cv::Mat descriptorsL;
	DBoW2::FeatureVector featureVector;
	std::vector<MapPoint*> mappoints;
	std::vector<MapPoint*> GetMapPointMatches() const { return mappoints; }
	// what the Sim3 projection search reads (include/KeyFrame.h:83-87, 134, 154-157; src/KeyFrame.cc:491-499)
	CameraParams camera;
	ScalePyramidInfo pyramid;
	ImageBounds imageBounds;
	FeaturesGrid grid;
	std::vector<size_t> GetFeaturesInArea(float x, float y, float r) const { return grid.GetFeaturesInArea(x, y, r); }
	bool IsInImage(float x, float y) const { return imageBounds.Contains(x, y); }
	// what Fuse x2, SearchBySim3 and SearchForTriangulation read or change besides (include/KeyFrame.h:60-62, 93-100, 131; src/KeyFrame.cc:76-92, 188-196)
	CameraPose pose;
	std::vector<float> uright;
	CameraPose GetPose() const { return pose; }
	Point3D GetCameraCenter() const { return pose.Invt(); }
	MapPoint* GetMapPoint(size_t idx) const { return mappoints[idx]; }
	void AddMapPoint(MapPoint* mp, size_t idx) { mappoints[idx] = mp; }
	std::set<MapPoint*> GetMapPoints() const
	{
		std::set<MapPoint*> s;
		for (MapPoint* mp : mappoints)
			if (mp && !mp->isBad()) s.insert(mp);
		return s;
	}
};

// src/MapPoint.cc:103-121 (AddObservation: +2 observations for a stereo keypoint, +1 otherwise) and :193-240 (Replace: this point goes bad,
// its observation moves to `other` unless `other` is already in that key frame, in which case the key frame's slot is cleared), for the
// one key frame of the model
inline void MapPoint::AddObservation(KeyFrame* kf, size_t idx)
{
	if (actions) { actions->push_back(2); actions->push_back(id); actions->push_back((int)idx); }
	inKeyFrame = kf; indexInKeyFrame = (int)idx;
	nobs += kf->uright.size() > idx && kf->uright[idx] >= 0 ? 2 : 1;
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
		else
			kf->mappoints[(size_t)indexInKeyFrame] = nullptr;
	}
}

// include/ORBmatcher.h:47-104, the members compiled here
class ORBmatcher
{
public:
	ORBmatcher(float nnratio = 0.6, bool checkOri = true) : fNNRatio_(nnratio), checkOrientation_(checkOri) {}
	static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
	int SearchByProjection(Frame& frame, const std::vector<MapPoint*>& mappoints, float th = 3);
	int SearchByProjection(Frame& currFrame, const Frame& lastFrame, float th, bool monocular);
	int SearchByProjection(Frame& frame, KeyFrame* keyframe, const std::set<MapPoint*>& alreadyFound, float th, int ORBdist);
	int SearchByProjection(const KeyFrame* keyframe, const Sim3& Scw, const std::vector<MapPoint*>& mappoints, std::vector<MapPoint*>& matched, int th);
	int SearchByBoW(KeyFrame* keyframe, Frame& frame, std::vector<MapPoint*>& matches);
	int SearchByBoW(KeyFrame* keyframe1, KeyFrame* keyframe2, std::vector<MapPoint*>& matches12);
	int SearchForInitialization(Frame& frame1, Frame& frame2, std::vector<cv::Point2f>& prevMatched, std::vector<int>& matches12,
		int windowSize = 10);
	int SearchForTriangulation(const KeyFrame* keyframe1, const KeyFrame* keyframe2, const cv::Mat& F12,
		std::vector<std::pair<size_t, size_t>>& matchIds, bool onlyStereo);
	int Fuse(KeyFrame* keyframe, const std::vector<MapPoint*>& mappoints, float th = 3.f);
	int Fuse(KeyFrame* keyframe, const Sim3& Scw, const std::vector<MapPoint*>& mappoints, float th, std::vector<MapPoint*>& replacePoints);
	int SearchBySim3(KeyFrame* keyframe1, KeyFrame* keyframe2, std::vector<MapPoint*>& matches12, const Sim3& S12, float th);

private:
	float fNNRatio_;
	bool checkOrientation_;
};
