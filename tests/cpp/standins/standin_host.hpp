// TEST INFRASTRUCTURE — host helpers of the stand-in MapPoint / KeyFrame types of oracle/ref_match_main.cpp for the build that drives the
// drop-in matcher classes (REF_MATCH_USE_SHIM: nothing of the reference is compiled, so the members the drop-in calls on the
// reference's objects need a definition here).  In production these are the reference's own methods; each one is a restatement of
// a few lines, cited.  Included after the stand-in classes, outside namespace ORB_SLAM2.
#ifndef HVO_TESTS_STANDIN_HOST_HPP
#define HVO_TESTS_STANDIN_HOST_HPP

#include <cmath>

namespace ORB_SLAM2 {

// src/MapPoint.cc:371-381
inline float MapPoint::GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
inline float MapPoint::GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }

// src/MapPoint.cc:383-415: float ratio, float logarithm (std::log(float)), ceil, clamp to the pyramid
template <class Owner>
static inline int standin_predict_scale(float mfMaxDistance, float currentDist, const Owner* o) {
    const float ratio = mfMaxDistance / currentDist;
    int nScale = (int)std::ceil(std::log(ratio) / o->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= o->mnScaleLevels) nScale = o->mnScaleLevels - 1;
    return nScale;
}
inline int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) { return standin_predict_scale(mfMaxDistance, currentDist, pKF); }
inline int MapPoint::PredictScale(const float& currentDist, Frame* pF) { return standin_predict_scale(mfMaxDistance, currentDist, pF); }

// src/MapPoint.cc:313-320
inline int MapPoint::GetIndexInKeyFrame(KeyFrame* pKF) {
    const auto it = mObservations.find(pKF);
    return it == mObservations.end() ? -1 : (int)it->second;
}

// src/KeyFrame.cc:780-783 (the key frame's integer bounds)
inline bool KeyFrame::IsInImage(const float& x, const float& y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }

// src/KeyFrame.cc:254-267
inline std::set<MapPoint*> KeyFrame::GetMapPoints() {
    std::set<MapPoint*> s;
    for (size_t i = 0; i < mvpMapPoints.size(); i++) {
        MapPoint* pMP = mvpMapPoints[i];
        if (pMP && !pMP->isBad()) s.insert(pMP);
    }
    return s;
}

}  // namespace ORB_SLAM2

#endif
