// TEST INFRASTRUCTURE — driver for the reference's own windowed matchers, compiled from the sources where they lie.  The
// functions are pulled out at build time by oracle/extract_ref.py (oracle/_ref/gen/*.inc, never committed):
//   src/Frame.cc:832-872, 1502-1555, 1557-1631, 1680-1690   AssignFeaturesToGrid(+ForLine), GetFeaturesInArea(+ForLine), PosInGrid
//   src/ORBmatcher.cc:37-43, 45-140, 1353-1497, 1630-1692   SearchByProjection(F, MapPoints, th), RadiusByViewingCos,
//                                                            SearchByProjection(Cur, Last, th, mono), ComputeThreeMaxima, DescriptorDistance
//   src/LSDmatcher.cpp:12-34, 561-664, 709-801, 1137-1153, 1436-1442   the two line SearchByProjection + helpers
//   src/Frame.cc:1371-1499                                   isInFrustum(MapPoint*, float), isInFrustum(MapLine*, float)  (as isInFrustumRef)
//   src/MapPoint.cc:371-381, 400-415; src/MapLine.cpp:537-558  Get{Min,Max}DistanceInvariance, PredictScale
//   src/ORBmatcher.cc:412-529                                SearchForInitialization, whole
//   include/auxiliar.h:40-45, src/LSDmatcher.cpp:968-1108    sort_descriptor_by_queryIdx, FrameBFMatchNew, mutualOverlap
//   src/ORBmatcher.cc:162-293                                SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), whole (with Thirdparty/DBoW2's
//                                                            own FeatureVector, FeatureVector.cpp compiled unmodified)
//   src/ORBmatcher.cc:143-160, 668-836                       CheckDistEpipolarLine, SearchForTriangulation, whole
//   src/ORBmatcher.cc:838-994                                Fuse(KeyFrame*, vpMapPoints, th), whole; src/KeyFrame.cc:627-666, 780-783
//                                                            GetFeaturesInArea, IsInImage; src/MapPoint.cc:383-398 PredictScale(dist, KeyFrame*)
//   src/MapPoint.cc:240-305, src/MapLine.cpp:331-396         ComputeDistinctiveDescriptors x2
//   src/ORBmatcher.cc:295-410, 531-666, 996-1121, 1123-1351, 1499-1628   SearchByProjection(KeyFrame*, Scw, ...), SearchByBoW(pKF1, pKF2, ...),
//                                                            Fuse(KeyFrame*, Scw, ...), SearchBySim3, SearchByProjection(Cur, KeyFrame*, found, th, ORBdist),
//                                                            whole; src/KeyFrame.cc:254-267 GetMapPoints, src/MapPoint.cc:313-320 GetIndexInKeyFrame
//   include/auxiliar.h:26-38, src/LSDmatcher.cpp:522-559, 803-966, 1110-1135, src/Frame.cc:1331-1355   SearchByDescriptor, matchNNR, match,
//                                                            SearchDouble x2, FrameBFMatch, lineDescriptorMAD x2 (their std::threads included)
//   src/LSDmatcher.cpp:1155-1231                             SearchForTriangulation(pKF1, pKF2, vector<pair>&) and (pKF1, pKF2, vector<int>&, isDouble)
//   src/lineIterator.cpp                                     whole file, unmodified
// and compiled against stand-in Frame / MapPoint / MapLine classes that carry exactly the members those functions touch
// (declared below with the reference header line each one mirrors) plus the OpenCV / Eigen stand-ins.
// One substitution: Frame::isInFrustum(MapLine*, float) (Frame.cc:1438-1499, not on the pinned path) returns the map line's
// precomputed mbTrackInView; the projection fields it would fill are given as input.
//
//   ref_match <in.bin> <out.bin>      (formats: see oracle/__init__.py ref_match_*)
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <thread>
#include <unordered_set>
#include <vector>

// REF_MATCH_USE_SHIM (tests/cpp/shim_match_main.cpp): the same driver and stand-in types, but ORBmatcher / LSDmatcher are the drop-in
// class templates of shim/ORBmatcher.h / shim/LSDmatcher.h (C ABI -> CUDA) instead of the reference's functions; nothing of the
// reference is compiled, the grid / candidate-list dumps are skipped (the drop-in builds its grids on the device).
#include "DBoW2/FeatureVector.h"   // Thirdparty/DBoW2 (header + FeatureVector.cpp, unmodified): std::map<NodeId, std::vector<unsigned int>>
#ifdef REF_MATCH_USE_SHIM
#include "keyline_standin.hpp"
#include <Eigen/Core>
#else
#include "precomp_custom.hpp"   // vendored line_descriptor umbrella (KeyLine)
#include <Eigen/Core>
#include "lineIterator.h"
#endif

using namespace std;
using namespace cv;
using namespace cv::line_descriptor;
using namespace Eigen;
typedef Matrix<double, 6, 1> Vector6d;   // include/auxiliar.h:23

#define FRAME_GRID_ROWS 48               // include/Frame.h:59-60
#define FRAME_GRID_COLS 64

namespace ORB_SLAM2 {
class KeyFrame;
class Frame;
class MapPoint {                          // include/MapPoint.h:49-102, 150-154
public:
    cv::Mat GetWorldPos() { return pos.clone(); }
    cv::Mat GetNormal() { return normal.clone(); }
    float GetMinDistanceInvariance();
    float GetMaxDistanceInvariance();
    int PredictScale(const float& currentDist, Frame* pF);
    MapPoint() {}
    MapPoint(const MapPoint& o) { *this = o; }
    MapPoint& operator=(const MapPoint& o) {   // (std::mutex is not copyable; the harness keeps map points in vectors)
        mTrackProjX = o.mTrackProjX; mTrackProjY = o.mTrackProjY; mTrackProjXR = o.mTrackProjXR; mbTrackInView = o.mbTrackInView;
        mnTrackScaleLevel = o.mnTrackScaleLevel; mTrackViewCos = o.mTrackViewCos; pos = o.pos; desc = o.desc; normal = o.normal;
        nobs = o.nobs; bad = o.bad; id = o.id; mfMinDistance = o.mfMinDistance; mfMaxDistance = o.mfMaxDistance;
        mObservations = o.mObservations;
        return *this;
    }
    void ComputeDistinctiveDescriptors();
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame* pKF);
    // bookkeeping stand-ins: what Fuse decides is logged as (map point id, key-frame keypoint or -1, action)
    void Replace(MapPoint* pMP);
    void AddObservation(KeyFrame* pKF, size_t idx);
    std::map<KeyFrame*, size_t> mObservations;      // include/MapPoint.h:129-155
    cv::Mat mDescriptor;
    bool mbBad = false;
    std::mutex mMutexFeatures;
    cv::Mat normal;
    float mfMinDistance = 0.f, mfMaxDistance = 0.f;
    std::mutex mMutexPos;
    int Observations() { return nobs; }
    bool isBad() { return bad; }
    cv::Mat GetDescriptor() { return desc.clone(); }
    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;
    cv::Mat pos, desc;
    int nobs = 0;
    bool bad = false;
    int id = -1;
};
class MapLine {                           // include/MapLine.h:40-110, 157-161
public:
    Vector6d GetWorldPos() { return wpos; }
    Vector3d GetWorldVector() { return wvec; }
    Vector3d GetNormal() { return wnormal; }
    float GetMinDistanceInvariance();
    float GetMaxDistanceInvariance();
    int PredictScale(const float& currentDist, const float& logScaleFactor);
    MapLine() {}
    MapLine(const MapLine& o) { *this = o; }
    MapLine& operator=(const MapLine& o) {
        mTrackProjX1 = o.mTrackProjX1; mTrackProjY1 = o.mTrackProjY1; mTrackProjX2 = o.mTrackProjX2; mTrackProjY2 = o.mTrackProjY2;
        mnTrackScaleLevel = o.mnTrackScaleLevel; mTrackViewCos = o.mTrackViewCos; mbTrackInView = o.mbTrackInView; wpos = o.wpos; wvec = o.wvec;
        wnormal = o.wnormal; desc = o.desc; nobs = o.nobs; bad = o.bad; id = o.id; mfMinDistance = o.mfMinDistance; mfMaxDistance = o.mfMaxDistance;
        mObservations = o.mObservations; mLDescriptor = o.mLDescriptor; mbBad = o.mbBad;
        return *this;
    }
    void ComputeDistinctiveDescriptors();
    std::map<KeyFrame*, size_t> mObservations;      // include/MapLine.h:138-162
    Mat mLDescriptor;
    bool mbBad = false;
    std::mutex mMutexFeatures;
    Vector3d wnormal;
    float mfMinDistance = 0.f, mfMaxDistance = 0.f;
    std::mutex mMutexPos;
    int Observations() { return nobs; }
    bool isBad() { return bad; }
    Mat GetDescriptor() { return desc.clone(); }
    float mTrackProjX1, mTrackProjY1, mTrackProjX2, mTrackProjY2;
    int mnTrackScaleLevel;
    float mTrackViewCos;
    bool mbTrackInView;
    Vector6d wpos;
    Vector3d wvec;
    Mat desc;
    int nobs = 0;
    bool bad = false;
    int id = -1;
};
class KeyFrame {                          // include/KeyFrame.h: the members SearchByBoW / ComputeDistinctiveDescriptors touch
public:
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    std::set<MapPoint*> GetMapPoints();
    std::vector<MapLine*> GetMapLineMatches() { return mvpMapLines; }
    MapLine* GetMapLine(const size_t& idx) { return mvpMapLines[idx]; }
    int NL = 0;
    std::vector<MapLine*> mvpMapLines;
    std::mutex mMutexFeatures;
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    bool isBad() { return bad; }
    bool IsInImage(const float& x, const float& y) const;
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    int N = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    int mnScaleLevels = 8;
    float mfLogScaleFactor = 0.f;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;   // include/KeyFrame.h:172-175, 249-252 (ints)
    float mfGridElementWidthInv = 0.f, mfGridElementHeightInv = 0.f;
    std::vector<std::vector<std::vector<size_t>>> mGrid;
    cv::Mat Ow, Rcw, tcw;
    DBoW2::FeatureVector mFeatVec;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors, mLineDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    bool bad = false;
};
class Frame {                             // include/Frame.h:129-131, 224-353
public:
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const;
    vector<size_t> GetFeaturesInAreaForLine(const float& x1, const float& y1, const float& x2, const float& y2, const float& r, const int minLevel = -1,
                                            const int maxLevel = -1, const float TH = 0.998) const;
    void AssignFeaturesToGrid();
    void AssignFeaturesToGridForLine();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    bool isInFrustum(MapLine* pML, float) { return pML->mbTrackInView; }
    bool isInFrustumRef(MapPoint* pMP, float viewingCosLimit);   // = the reference's isInFrustum, see the include below
    bool isInFrustumRef(MapLine* pML, float viewingCosLimit);
    void lineDescriptorMAD(vector<vector<DMatch>> matches, double& nn_mad, double& nn12_mad) const;   // include/Frame.h:98
    DBoW2::FeatureVector mFeatVec;           // include/Frame.h:268
    cv::Mat mRcw, mtcw, mOw;                 // include/Frame.h:408-411
    int mnScaleLevels = 8;                   // :330-332
    float mfLogScaleFactor = 0.f;
    static float fx, fy, cx, cy;
    float mb = 0.f, mbf = 0.f;
    int N = 0, NL = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors, mLdesc;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    vector<KeyLine> mvKeylinesUn;
    std::vector<std::pair<Eigen::Vector3d, Eigen::Vector3d>> mvLines3D;
    vector<Vector3d> mvKeyLineFunctions;
    vector<bool> mvbLineOutlier;
    std::vector<MapLine*> mvpMapLines;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridForLine[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    cv::Mat mTcw;
    vector<float> mvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};
float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;

#ifndef REF_MATCH_USE_SHIM
class ORBmatcher {                        // include/ORBmatcher.h:38-104
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<pair<size_t, size_t>>& vMatchedPairs, const bool bOnlyStereo);
    int Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th = 3.0);
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12, const float th);
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint);
    static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
protected:
    bool CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF);
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};
class LSDmatcher {                        // include/LSDmatcher.h:20-70
public:
    LSDmatcher(float nnratio = 0.6, bool checkOri = true);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th);
    int SearchByProjection(Frame& F, const std::vector<MapLine*>& vpMapLines, const bool eval_orient, const float th = 3);
    static int DescriptorDistance(const Mat& a, const Mat& b);
    void FrameBFMatchNew(cv::Mat ldesc1, cv::Mat ldesc2, vector<int>& LineMatches, vector<KeyLine> kls1, vector<KeyLine> kls2,
                         vector<Eigen::Vector3d> kls2func, cv::Mat F, float TH);
    float mutualOverlap(const std::vector<cv::Mat>& collinear_points);
    int SearchByDescriptor(KeyFrame* pKF, Frame& currentF, vector<MapLine*>& vpMapLineMatches);
    int matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12);
    int match(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12);
    int SearchDouble(Frame& InitialFrame, Frame& CurrentFrame, vector<int>& LineMatches);
    int SearchDouble(KeyFrame* KF, Frame& CurrentFrame);
    void FrameBFMatch(cv::Mat ldesc1, cv::Mat ldesc2, vector<int>& LineMatches, float TH);
    void lineDescriptorMAD(vector<vector<DMatch>> line_matches, double& nn_mad, double& nn12_mad) const;
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, vector<pair<size_t, size_t>>& vMatchedPairs);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, vector<int>& vMatchedPairs, bool isDouble = false);
    static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
protected:
    double computeAngle2D(const cv::Mat& vector1, const cv::Mat& vector2);
    float RadiusByViewingCos(const float& viewCos);
    float mfNNratio;
    bool mbCheckOrientation;
};
#include "gen/frame_grid.inc"
#include "gen/orbmatcher.inc"
#include "gen/orb_init.inc"
#include "gen/orb_bow.inc"
#include "gen/orb_tri.inc"
#include "gen/orb_fuse.inc"
#include "gen/keyframe_area.inc"
#include "gen/mappoint_scale_kf.inc"
#include "gen/orb_kf_scw.inc"
#include "gen/orb_bow_kf.inc"
#include "gen/orb_fuse_scw.inc"
#include "gen/orb_sim3.inc"
#include "gen/orb_reloc.inc"
#include "gen/keyframe_mps.inc"
#include "gen/mappoint_index.inc"
#define isInFrustum isInFrustumRef
#include "gen/frame_frustum.inc"
#undef isInFrustum
#include "gen/map_scale.inc"
}  // namespace ORB_SLAM2
namespace ORB_SLAM2 {
#include "gen/lsdmatcher.inc"
#include "gen/lsd_bfnew.inc"
#include "gen/lsd_bf.inc"
#include "gen/lsd_tri.inc"
#include "gen/distinctive.inc"
}  // namespace ORB_SLAM2
#else
}  // namespace ORB_SLAM2
#include "standin_host.hpp"   // tests/cpp/standins: KeyFrame::GetMapPoints / IsInImage, MapPoint::PredictScale ... for the stand-in types
#include "ORBmatcher.h"
#include "LSDmatcher.h"
#include "FrustumGPU.h"
namespace ORB_SLAM2 {
typedef hvo_shim::ORBmatcherT<Frame, MapPoint> ORBmatcher;
typedef hvo_shim::LSDmatcherT<Frame, MapLine> LSDmatcher;
}  // namespace ORB_SLAM2
#endif

using namespace ORB_SLAM2;

struct FuseEvent { int32_t mp, idx, action; };   // action 0: AddObservation + AddMapPoint, 1: pMP->Replace(pMPinKF), 2: pMPinKF->Replace(pMP)
static std::vector<FuseEvent> g_fuse_log;
static KeyFrame* g_fuse_kf = nullptr;
void ORB_SLAM2::MapPoint::AddObservation(KeyFrame*, size_t idx) { g_fuse_log.push_back(FuseEvent{id, (int32_t)idx, 0}); nobs++; }
void ORB_SLAM2::MapPoint::Replace(MapPoint* pMP) {
    // `this` is replaced by pMP.  One of the two sits in a slot of the key frame (pMPinKF), the other is the incoming map point.
    int32_t idx = -1;
    bool this_in_kf = false;
    for (size_t i = 0; i < g_fuse_kf->mvpMapPoints.size() && idx < 0; ++i) {
        if (g_fuse_kf->mvpMapPoints[i] == this) { idx = (int32_t)i; this_in_kf = true; }
        else if (g_fuse_kf->mvpMapPoints[i] == pMP) idx = (int32_t)i;
    }
    if (this_in_kf) {   // pMPinKF->Replace(pMP): the incoming point takes the slot (MapPoint::Replace re-points the key frame, src/MapPoint.cc:185-238)
        g_fuse_log.push_back(FuseEvent{pMP->id, idx, 2});
        g_fuse_kf->mvpMapPoints[idx] = pMP;
    } else {            // pMP->Replace(pMPinKF): the incoming point dissolves into the one the key frame holds
        g_fuse_log.push_back(FuseEvent{id, idx, 1});
    }
}


// ---- binary I/O ----
static FILE *g_in, *g_out;
template <class T> static T get() { T v; if (std::fread(&v, sizeof(T), 1, g_in) != 1) { std::fprintf(stderr, "ref_match: short input\n"); std::exit(4); } return v; }
template <class T> static void get_n(T* p, size_t n) { if (n && std::fread(p, sizeof(T), n, g_in) != n) { std::fprintf(stderr, "ref_match: short input\n"); std::exit(4); } }
template <class T> static void put(const T& v) { std::fwrite(&v, sizeof(T), 1, g_out); }
static cv::Mat get_desc_rows(int n) {
    cv::Mat m(n > 0 ? n : 1, 32, CV_8UC1);
    get_n(m.data, (size_t)n * 32);
    return m;
}
static void set_bounds() {
    float b[4]; get_n(b, 4);
    Frame::mnMinX = b[0]; Frame::mnMinY = b[1]; Frame::mnMaxX = b[2]; Frame::mnMaxY = b[3];
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);    // Frame.cc:198-199
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
}
static cv::Mat get_pose() { cv::Mat T(4, 4, CV_32FC1); get_n((float*)T.data, 16); return T; }

// point frame: N, keysUn[N], uright[N], desc[N][32], claimed[N] (holds a map point with observations)
static void read_point_frame(Frame& F, std::vector<MapPoint>& claimed_pool) {
    F.N = get<int32_t>();
    F.mvKeysUn.resize(F.N); get_n(F.mvKeysUn.data(), F.N);
    F.mvKeys = F.mvKeysUn;
    F.mvuRight.resize(F.N); get_n(F.mvuRight.data(), F.N);
    F.mDescriptors = get_desc_rows(F.N);
    std::vector<uint8_t> cl(F.N); get_n(cl.data(), F.N);
    claimed_pool.resize(F.N);
    F.mvpMapPoints.assign(F.N, nullptr);
    for (int i = 0; i < F.N; ++i) if (cl[i]) { claimed_pool[i].nobs = 1; claimed_pool[i].id = -2; F.mvpMapPoints[i] = &claimed_pool[i]; }
    F.mvScaleFactors.resize(8); get_n(F.mvScaleFactors.data(), 8);
#ifndef REF_MATCH_USE_SHIM
    F.AssignFeaturesToGrid();
#endif
}
// line frame: NL, keylinesUn[NL], line functions [NL][3], desc, lines3D [NL][6], claimed[NL]
static void read_line_frame(Frame& F, std::vector<MapLine>& claimed_pool) {
    F.NL = get<int32_t>();
    F.mvKeylinesUn.resize(F.NL); get_n(F.mvKeylinesUn.data(), F.NL);
    F.mvKeyLineFunctions.resize(F.NL);
    for (int i = 0; i < F.NL; ++i) get_n(F.mvKeyLineFunctions[i].data(), 3);
    F.mLdesc = get_desc_rows(F.NL);
    F.mvLines3D.resize(F.NL);
    for (int i = 0; i < F.NL; ++i) { get_n(F.mvLines3D[i].first.data(), 3); get_n(F.mvLines3D[i].second.data(), 3); }
    std::vector<uint8_t> cl(F.NL); get_n(cl.data(), F.NL);
    claimed_pool.resize(F.NL);
    F.mvpMapLines.assign(F.NL, nullptr);
    for (int i = 0; i < F.NL; ++i) if (cl[i]) { claimed_pool[i].nobs = 1; claimed_pool[i].id = -2; F.mvpMapLines[i] = &claimed_pool[i]; }
#ifndef REF_MATCH_USE_SHIM
    F.AssignFeaturesToGridForLine();
#endif
}
static void put_grid(const std::vector<std::size_t> (*grid)[FRAME_GRID_ROWS]) {
    for (int ix = 0; ix < FRAME_GRID_COLS; ++ix)
        for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) {
            put<int32_t>((int32_t)grid[ix][iy].size());
            for (size_t v : grid[ix][iy]) put<int32_t>((int32_t)v);
        }
}

// key frame = a point frame read through the Frame reader (grid included) + what KeyFrame's constructor copies from it
// (src/KeyFrame.cc:33-60: the grid, the integer image bounds, the inverse cell sizes)
static void keyframe_from_frame(const Frame& F, KeyFrame& K) {
    K.N = F.N; K.mvKeysUn = F.mvKeysUn; K.mvuRight = F.mvuRight; K.mDescriptors = F.mDescriptors; K.mvScaleFactors = F.mvScaleFactors;
    K.mvpMapPoints.assign(K.N, nullptr);
#ifndef REF_MATCH_USE_SHIM
    K.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
    for (int ix = 0; ix < FRAME_GRID_COLS; ++ix) for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) K.mGrid[ix][iy] = F.mGrid[ix][iy];
#endif
    K.mnMinX = (int)Frame::mnMinX; K.mnMinY = (int)Frame::mnMinY; K.mnMaxX = (int)Frame::mnMaxX; K.mnMaxY = (int)Frame::mnMaxY;
    K.mfGridElementWidthInv = Frame::mfGridElementWidthInv; K.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
}
// fx fy cx cy bf, Rcw, tcw, Ow, mfLogScaleFactor, mnScaleLevels
static void read_kf_camera(KeyFrame& K) {
    float cam[5]; get_n(cam, 5);
    K.fx = cam[0]; K.fy = cam[1]; K.cx = cam[2]; K.cy = cam[3]; K.mbf = cam[4];
    K.Rcw = cv::Mat(3, 3, CV_32FC1); get_n((float*)K.Rcw.data, 9);
    K.tcw = cv::Mat(3, 1, CV_32FC1); get_n((float*)K.tcw.data, 3);
    K.Ow = cv::Mat(3, 1, CV_32FC1); get_n((float*)K.Ow.data, 3);
    K.mfLogScaleFactor = get<float>(); K.mnScaleLevels = get<int32_t>();
}
// map point record: pos[3], normal[3], min, max distance (hvo_map_point), descriptor[32], 4 flag bytes, int32 slot
static void read_map_point(MapPoint& m, uint8_t fl[4], int32_t& slot) {
    m.pos = cv::Mat(3, 1, CV_32FC1); get_n((float*)m.pos.data, 3);
    m.normal = cv::Mat(3, 1, CV_32FC1); get_n((float*)m.normal.data, 3);
    m.mfMinDistance = get<float>(); m.mfMaxDistance = get<float>();
    m.desc = get_desc_rows(1);
    get_n(fl, 4);
    slot = get<int32_t>();
}
static void read_featvec(DBoW2::FeatureVector& fv) {
    const int nn = get<int32_t>();
    for (int k = 0; k < nn; ++k) {
        const int node = get<int32_t>(), cnt = get<int32_t>();
        for (int j = 0; j < cnt; ++j) fv.addFeature((DBoW2::NodeId)node, (unsigned int)get<int32_t>());
    }
}
// the map points a key frame holds (ops 14, 15): per keypoint has / bad / found / - flags (slot unused), ids = keypoint index
static void read_kf_map_points(KeyFrame& K, std::vector<MapPoint>& mps, std::vector<uint8_t>* found) {
    mps.resize(K.N);
    if (found) found->assign(K.N, 0);
    for (int i = 0; i < K.N; ++i) {
        uint8_t fl[4]; int32_t slot;
        read_map_point(mps[i], fl, slot);
        mps[i].id = i; mps[i].bad = fl[1] != 0; mps[i].nobs = 1;
        if (found) (*found)[i] = fl[2];
        if (fl[0]) { K.mvpMapPoints[i] = &mps[i]; mps[i].mObservations[&K] = (size_t)i; }
    }
}

int main(int argc, char** argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: ref_match in.bin out.bin\n"); return 2; }
    g_in = std::fopen(argv[1], "rb");
    g_out = std::fopen(argv[2], "wb");
    if (!g_in || !g_out) return 2;
    if (get<int32_t>() != 0x4d544348) return 3;
    const int op = get<int32_t>();
    set_bounds();
    if (op == 0) {          // ORBmatcher::SearchByProjection(F, vpMapPoints, th)
        Frame F; std::vector<MapPoint> pool;
        read_point_frame(F, pool);
        const int M = get<int32_t>();
        const float th = get<float>(), nnratio = get<float>();
        std::vector<MapPoint> mps(M);
        std::vector<MapPoint*> vp(M);
        for (int i = 0; i < M; ++i) {
            MapPoint& m = mps[i];
            m.mTrackProjX = get<float>(); m.mTrackProjY = get<float>(); m.mTrackProjXR = get<float>();
            m.mnTrackScaleLevel = get<int32_t>(); m.mTrackViewCos = get<float>();
            m.mbTrackInView = get<uint8_t>() != 0; m.bad = get<uint8_t>() != 0; m.nobs = get<uint8_t>(); (void)get<uint8_t>();
            m.desc = get_desc_rows(1); m.id = i; vp[i] = &m;
        }
        // window queries answered by the reference's GetFeaturesInArea, for the grid / candidate-order pin
        const int nq = get<int32_t>();
        std::vector<float> wq((size_t)nq * 5); get_n(wq.data(), wq.size());
#ifndef REF_MATCH_USE_SHIM
        put_grid(F.mGrid);
        for (int i = 0; i < nq; ++i) {
            const vector<size_t> v = F.GetFeaturesInArea(wq[5 * i], wq[5 * i + 1], wq[5 * i + 2], (int)wq[5 * i + 3], (int)wq[5 * i + 4]);
            put<int32_t>((int32_t)v.size());
            for (size_t x : v) put<int32_t>((int32_t)x);
        }
#endif
        ORBmatcher matcher(nnratio, true);
        const int nm = matcher.SearchByProjection(F, vp, th);
        put<int32_t>(nm);
        for (int i = 0; i < F.N; ++i) put<int32_t>(F.mvpMapPoints[i] ? F.mvpMapPoints[i]->id : -1);
    } else if (op == 1) {   // ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono)
        Frame C, L; std::vector<MapPoint> pool;
        read_point_frame(C, pool);
        float cam[6]; get_n(cam, 6);
        Frame::fx = cam[0]; Frame::fy = cam[1]; Frame::cx = cam[2]; Frame::cy = cam[3]; C.mb = cam[4]; C.mbf = cam[5];
        C.mTcw = get_pose(); L.mTcw = get_pose();
        const float th = get<float>();
        const int mono = get<int32_t>(), check_ori = get<int32_t>();
        L.N = get<int32_t>();
        L.mvKeys.resize(L.N); get_n(L.mvKeys.data(), L.N);
        L.mvKeysUn = L.mvKeys;
        std::vector<MapPoint> mps(L.N);
        L.mvpMapPoints.assign(L.N, nullptr); L.mvbOutlier.assign(L.N, false);
        for (int i = 0; i < L.N; ++i) {
            const int has = get<uint8_t>(); L.mvbOutlier[i] = get<uint8_t>() != 0; mps[i].nobs = get<uint8_t>(); (void)get<uint8_t>();
            mps[i].pos = cv::Mat(3, 1, CV_32FC1); get_n((float*)mps[i].pos.data, 3);
            mps[i].desc = get_desc_rows(1); mps[i].id = i;
            if (has) L.mvpMapPoints[i] = &mps[i];
        }
        ORBmatcher matcher(0.9f, check_ori != 0);
        const int nm = matcher.SearchByProjection(C, L, th, mono != 0);
        put<int32_t>(nm);
        for (int i = 0; i < C.N; ++i) put<int32_t>(C.mvpMapPoints[i] ? C.mvpMapPoints[i]->id : -1);
    } else if (op == 2) {   // LSDmatcher::SearchByProjection(F, vpMapLines, eval_orient, th)
        Frame F; std::vector<MapLine> pool;
        read_line_frame(F, pool);
        const int M = get<int32_t>();
        const float th = get<float>(), nnratio = get<float>();
        std::vector<MapLine> mls(M);
        std::vector<MapLine*> vp(M);
        for (int i = 0; i < M; ++i) {
            MapLine& m = mls[i];
            m.mTrackProjX1 = get<float>(); m.mTrackProjY1 = get<float>(); m.mTrackProjX2 = get<float>(); m.mTrackProjY2 = get<float>();
            m.mnTrackScaleLevel = get<int32_t>(); m.mTrackViewCos = get<float>();
            m.mbTrackInView = get<uint8_t>() != 0; m.bad = get<uint8_t>() != 0; m.nobs = get<uint8_t>(); (void)get<uint8_t>();
            get_n(m.wvec.data(), 3);
            m.desc = get_desc_rows(1); m.id = i; vp[i] = &m;
        }
        const int nq = get<int32_t>();
        std::vector<float> wq((size_t)nq * 6); get_n(wq.data(), wq.size());
#ifndef REF_MATCH_USE_SHIM
        put_grid(F.mGridForLine);
        for (int i = 0; i < nq; ++i) {
            const vector<size_t> v = F.GetFeaturesInAreaForLine(wq[6 * i], wq[6 * i + 1], wq[6 * i + 2], wq[6 * i + 3], wq[6 * i + 4], -1, -1, wq[6 * i + 5]);
            put<int32_t>((int32_t)v.size());
            for (size_t x : v) put<int32_t>((int32_t)x);
        }
#endif
        LSDmatcher matcher(nnratio, true);
        const int nm = matcher.SearchByProjection(F, vp, true, th);
        put<int32_t>(nm);
        for (int i = 0; i < F.NL; ++i) put<int32_t>(F.mvpMapLines[i] ? F.mvpMapLines[i]->id : -1);
    } else if (op == 3) {   // LSDmatcher::SearchByProjection(CurrentFrame, LastFrame, th)
        Frame C, L; std::vector<MapLine> pool;
        read_line_frame(C, pool);
        C.mTcw = get_pose(); L.mTcw = get_pose();
        const float th = get<float>();
        L.NL = get<int32_t>();
        L.mvKeylinesUn.resize(L.NL); get_n(L.mvKeylinesUn.data(), L.NL);
        std::vector<MapLine> mls(L.NL);
        L.mvpMapLines.assign(L.NL, nullptr); L.mvbLineOutlier.assign(L.NL, false);
        for (int i = 0; i < L.NL; ++i) {
            MapLine& m = mls[i];
            const int has = get<uint8_t>(); L.mvbLineOutlier[i] = get<uint8_t>() != 0; m.nobs = get<uint8_t>(); m.mbTrackInView = get<uint8_t>() != 0;
            m.mTrackProjX1 = get<float>(); m.mTrackProjY1 = get<float>(); m.mTrackProjX2 = get<float>(); m.mTrackProjY2 = get<float>();
            m.mnTrackScaleLevel = get<int32_t>();
            m.desc = get_desc_rows(1); m.id = i;
            if (has) L.mvpMapLines[i] = &m;
        }
        LSDmatcher matcher(0.95f, true);
        const int nm = matcher.SearchByProjection(C, L, th);
        put<int32_t>(nm);
        for (int i = 0; i < C.NL; ++i) put<int32_t>(C.mvpMapLines[i] ? C.mvpMapLines[i]->id : -1);
    } else if (op == 4 || op == 5) {   // Frame::isInFrustum(MapPoint*, limit) / (MapLine*, limit) over a batch
        Frame F;
        float cam[5]; get_n(cam, 5);
        Frame::fx = cam[0]; Frame::fy = cam[1]; Frame::cx = cam[2]; Frame::cy = cam[3]; F.mbf = cam[4];
        F.mRcw = cv::Mat(3, 3, CV_32FC1); get_n((float*)F.mRcw.data, 9);
        F.mtcw = cv::Mat(3, 1, CV_32FC1); get_n((float*)F.mtcw.data, 3);
        F.mOw = cv::Mat(3, 1, CV_32FC1); get_n((float*)F.mOw.data, 3);
        F.mfLogScaleFactor = get<float>(); F.mnScaleLevels = get<int32_t>();
        const float limit = get<float>();
        const int M = get<int32_t>();
        if (op == 4) {
            std::vector<MapPoint> ms(M);
            for (int i = 0; i < M; ++i) {
                MapPoint& m = ms[i];
                m.pos = cv::Mat(3, 1, CV_32FC1); get_n((float*)m.pos.data, 3);
                m.normal = cv::Mat(3, 1, CV_32FC1); get_n((float*)m.normal.data, 3);
                m.mfMinDistance = get<float>(); m.mfMaxDistance = get<float>();
                m.mTrackProjX = m.mTrackProjY = m.mTrackProjXR = m.mTrackViewCos = 0.f; m.mnTrackScaleLevel = 0;
            }
            std::vector<char> in(M, 0);
#ifdef REF_MATCH_USE_SHIM
            std::vector<MapPoint*> vp(M);
            for (int i = 0; i < M; ++i) vp[i] = &ms[i];
            hvo_shim::FrustumCullerT<Frame, MapPoint, MapLine> culler;
            culler.isInFrustum(F, vp, limit, in);
#else
            for (int i = 0; i < M; ++i) in[i] = F.isInFrustumRef(&ms[i], limit);
#endif
            for (int i = 0; i < M; ++i) {
                const MapPoint& m = ms[i];
                put<float>(m.mTrackProjX); put<float>(m.mTrackProjY); put<float>(m.mTrackProjXR); put<int32_t>(m.mnTrackScaleLevel);
                put<float>(m.mTrackViewCos); put<int32_t>(in[i] && m.mbTrackInView ? 1 : 0);
            }
        } else {
            std::vector<MapLine> ms(M);
            for (int i = 0; i < M; ++i) {
                MapLine& m = ms[i];
                get_n(m.wpos.data(), 6); get_n(m.wnormal.data(), 3); get_n(m.wvec.data(), 3);
                m.mfMinDistance = get<float>(); m.mfMaxDistance = get<float>();
                m.mTrackProjX1 = m.mTrackProjY1 = m.mTrackProjX2 = m.mTrackProjY2 = m.mTrackViewCos = 0.f; m.mnTrackScaleLevel = 0;
            }
            std::vector<char> in(M, 0);
#ifdef REF_MATCH_USE_SHIM
            std::vector<MapLine*> vp(M);
            for (int i = 0; i < M; ++i) vp[i] = &ms[i];
            hvo_shim::FrustumCullerT<Frame, MapPoint, MapLine> culler;
            culler.isInFrustum(F, vp, limit, in);
#else
            for (int i = 0; i < M; ++i) in[i] = F.isInFrustumRef(&ms[i], limit);
#endif
            for (int i = 0; i < M; ++i) {
                const MapLine& m = ms[i];
                put<float>(m.mTrackProjX1); put<float>(m.mTrackProjY1); put<float>(m.mTrackProjX2); put<float>(m.mTrackProjY2);
                put<int32_t>(m.mnTrackScaleLevel); put<float>(m.mTrackViewCos); put<int32_t>(in[i] && m.mbTrackInView ? 1 : 0);
            }
        }
    } else if (op == 6) {   // ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)
        Frame F1, F2; std::vector<MapPoint> pool;
        read_point_frame(F2, pool);
        F1.N = get<int32_t>();
        F1.mvKeysUn.resize(F1.N); get_n(F1.mvKeysUn.data(), F1.N);
        F1.mDescriptors = get_desc_rows(F1.N);
        std::vector<cv::Point2f> prev(F1.N);
        for (int i = 0; i < F1.N; ++i) { prev[i].x = get<float>(); prev[i].y = get<float>(); }
        const int window = get<int32_t>();
        const float nnratio = get<float>();
        const int check_ori = get<int32_t>();
        std::vector<int> m12;
        ORBmatcher matcher(nnratio, check_ori != 0);
        const int nm = matcher.SearchForInitialization(F1, F2, prev, m12, window);
        put<int32_t>(nm);
        for (int i = 0; i < F1.N; ++i) put<int32_t>(m12[i]);
        for (int i = 0; i < F1.N; ++i) { put<float>(prev[i].x); put<float>(prev[i].y); }
    } else if (op == 7) {   // LSDmatcher::FrameBFMatchNew(ldesc1, ldesc2, LineMatches, kls1, kls2, kls2func, F, TH)
        const int n1 = get<int32_t>();
        vector<KeyLine> k1(n1); get_n(k1.data(), n1);
        cv::Mat d1 = get_desc_rows(n1);
        const int n2 = get<int32_t>();
        vector<KeyLine> k2(n2); get_n(k2.data(), n2);
        cv::Mat d2 = get_desc_rows(n2);
        vector<Eigen::Vector3d> f2(n2);
        for (int i = 0; i < n2; ++i) get_n(f2[i].data(), 3);
        cv::Mat Fm(3, 3, CV_32FC1); get_n((float*)Fm.data, 9);
        const float TH = get<float>(), nnratio = get<float>();
        if (d1.rows != n1) d1 = d1.rowRange(0, n1);
        if (d2.rows != n2) d2 = d2.rowRange(0, n2);
        vector<int> lm;
        LSDmatcher matcher(nnratio, true);
        matcher.FrameBFMatchNew(d1, d2, lm, k1, k2, f2, Fm, TH);
        for (int i = 0; i < n1; ++i) put<int32_t>(lm[i]);
    } else if (op == 8) {   // MapPoint / MapLine::ComputeDistinctiveDescriptors over a batch of map elements
        const int lines = get<int32_t>(), ngroups = get<int32_t>();
        std::vector<int32_t> off(ngroups + 1); get_n(off.data(), ngroups + 1);
        const int total = off[ngroups];
        std::vector<KeyFrame> kfs(total > 0 ? total : 1);   // one key frame per observation, ascending addresses = std::map order
        for (int i = 0; i < total; ++i) {
            cv::Mat d = get_desc_rows(1);
            kfs[i].mDescriptors = d; kfs[i].mLineDescriptors = d;
        }
#ifndef REF_MATCH_USE_SHIM
        for (int g = 0; g < ngroups; ++g) {
            cv::Mat best;
            if (lines) {
                MapLine m;
                for (int i = off[g]; i < off[g + 1]; ++i) m.mObservations[&kfs[i]] = 0;
                m.ComputeDistinctiveDescriptors();
                best = m.mLDescriptor;
            } else {
                MapPoint m;
                for (int i = off[g]; i < off[g + 1]; ++i) m.mObservations[&kfs[i]] = 0;
                m.ComputeDistinctiveDescriptors();
                best = m.mDescriptor;
            }
            int idx = -1;    // which observation's descriptor was kept (first identical row)
            if (!best.empty())
                for (int i = off[g]; i < off[g + 1] && idx < 0; ++i) if (std::memcmp(best.data, kfs[i].mDescriptors.data, 32) == 0) idx = i - off[g];
            put<int32_t>(idx);
            uint8_t zero[32] = {0};
            std::fwrite(best.empty() ? zero : best.data, 1, 32, g_out);
        }
#else
        return 6;   // the drop-in for this call is hvo_match_distinctive (C ABI), exercised from Python
#endif
    } else if (op == 9) {   // ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches)
        KeyFrame KF; Frame F;
        const int n1 = get<int32_t>();
        KF.mvKeysUn.resize(n1); get_n(KF.mvKeysUn.data(), n1);
        KF.mDescriptors = get_desc_rows(n1);
        std::vector<uint8_t> has(n1), bad(n1); get_n(has.data(), n1); get_n(bad.data(), n1);
        std::vector<MapPoint> mps(n1);
        KF.mvpMapPoints.assign(n1, nullptr);
        for (int i = 0; i < n1; ++i) { mps[i].id = i; mps[i].bad = bad[i] != 0; if (has[i]) KF.mvpMapPoints[i] = &mps[i]; }
        auto read_fv = [&](DBoW2::FeatureVector& fv) {
            const int nn = get<int32_t>();
            for (int k = 0; k < nn; ++k) {
                const int node = get<int32_t>(), cnt = get<int32_t>();
                for (int j = 0; j < cnt; ++j) fv.addFeature((DBoW2::NodeId)node, (unsigned int)get<int32_t>());
            }
        };
        read_fv(KF.mFeatVec);
        F.N = get<int32_t>();
        F.mvKeys.resize(F.N); get_n(F.mvKeys.data(), F.N);
        F.mvKeysUn = F.mvKeys;
        F.mDescriptors = get_desc_rows(F.N);
        read_fv(F.mFeatVec);
        const float nnratio = get<float>();
        const int check_ori = get<int32_t>();
        std::vector<MapPoint*> matches;
        ORBmatcher matcher(nnratio, check_ori != 0);
        const int nm = matcher.SearchByBoW(&KF, F, matches);
        put<int32_t>(nm);
        for (int i = 0; i < F.N; ++i) put<int32_t>(matches[i] ? matches[i]->id : -1);
    } else if (op == 10) {   // ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)
        KeyFrame K1, K2;
        std::vector<MapPoint> pool1, pool2;
        auto read_kf = [&](KeyFrame& K, std::vector<MapPoint>& pool) {
            K.N = get<int32_t>();
            K.mvKeysUn.resize(K.N); get_n(K.mvKeysUn.data(), K.N);
            K.mvuRight.resize(K.N); get_n(K.mvuRight.data(), K.N);
            K.mDescriptors = get_desc_rows(K.N);
            std::vector<uint8_t> has(K.N); get_n(has.data(), K.N);
            pool.resize(K.N);
            K.mvpMapPoints.assign(K.N, nullptr);
            for (int i = 0; i < K.N; ++i) if (has[i]) K.mvpMapPoints[i] = &pool[i];
            const int nn = get<int32_t>();
            for (int k = 0; k < nn; ++k) {
                const int node = get<int32_t>(), cnt = get<int32_t>();
                for (int j = 0; j < cnt; ++j) K.mFeatVec.addFeature((DBoW2::NodeId)node, (unsigned int)get<int32_t>());
            }
        };
        read_kf(K1, pool1);
        K1.Ow = cv::Mat(3, 1, CV_32FC1); get_n((float*)K1.Ow.data, 3);
        read_kf(K2, pool2);
        K2.Rcw = cv::Mat(3, 3, CV_32FC1); get_n((float*)K2.Rcw.data, 9);
        K2.tcw = cv::Mat(3, 1, CV_32FC1); get_n((float*)K2.tcw.data, 3);
        float cam[4]; get_n(cam, 4);
        K2.fx = cam[0]; K2.fy = cam[1]; K2.cx = cam[2]; K2.cy = cam[3];
        K2.mvScaleFactors.resize(8); get_n(K2.mvScaleFactors.data(), 8);
        K2.mvLevelSigma2.resize(8); get_n(K2.mvLevelSigma2.data(), 8);
        cv::Mat F12(3, 3, CV_32FC1); get_n((float*)F12.data, 9);
        const int only_stereo = get<int32_t>(), check_ori = get<int32_t>();
        const float nnratio = get<float>();
        std::vector<pair<size_t, size_t>> pairs;
        ORBmatcher matcher(nnratio, check_ori != 0);
        const int nm = matcher.SearchForTriangulation(&K1, &K2, F12, pairs, only_stereo != 0);
        put<int32_t>(nm);
        put<int32_t>((int32_t)pairs.size());
        for (const auto& pr : pairs) { put<int32_t>((int32_t)pr.first); put<int32_t>((int32_t)pr.second); }
    } else if (op == 11) {   // ORBmatcher::Fuse(pKF, vpMapPoints, th)
        Frame F; std::vector<MapPoint> pool;
        read_point_frame(F, pool);               // key frame's features through the Frame reader (grid included); claimed = holds a map point
        KeyFrame K;
        K.N = F.N; K.mvKeysUn = F.mvKeysUn; K.mvuRight = F.mvuRight; K.mDescriptors = F.mDescriptors; K.mvScaleFactors = F.mvScaleFactors;
        std::vector<uint8_t> kfobs(K.N); get_n(kfobs.data(), K.N);      // Observations() of the map point a keypoint holds
        std::vector<MapPoint> held(K.N);
        K.mvpMapPoints.assign(K.N, nullptr);
        for (int i = 0; i < K.N; ++i) if (F.mvpMapPoints[i]) { held[i].id = -100 - i; held[i].nobs = kfobs[i]; K.mvpMapPoints[i] = &held[i]; }
#ifndef REF_MATCH_USE_SHIM
        K.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
        for (int ix = 0; ix < FRAME_GRID_COLS; ++ix) for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) K.mGrid[ix][iy] = F.mGrid[ix][iy];   // KeyFrame ctor: mGrid = F.mGrid
#endif
        K.mnMinX = (int)Frame::mnMinX; K.mnMinY = (int)Frame::mnMinY; K.mnMaxX = (int)Frame::mnMaxX; K.mnMaxY = (int)Frame::mnMaxY;
        K.mfGridElementWidthInv = Frame::mfGridElementWidthInv; K.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
        float cam[5]; get_n(cam, 5);
        K.fx = cam[0]; K.fy = cam[1]; K.cx = cam[2]; K.cy = cam[3]; K.mbf = cam[4];
        K.Rcw = cv::Mat(3, 3, CV_32FC1); get_n((float*)K.Rcw.data, 9);
        K.tcw = cv::Mat(3, 1, CV_32FC1); get_n((float*)K.tcw.data, 3);
        K.Ow = cv::Mat(3, 1, CV_32FC1); get_n((float*)K.Ow.data, 3);
        K.mvInvLevelSigma2.resize(8); get_n(K.mvInvLevelSigma2.data(), 8);
        K.mfLogScaleFactor = get<float>(); K.mnScaleLevels = get<int32_t>();
        const float th = get<float>();
        const int M = get<int32_t>();
        std::vector<MapPoint> mps(M);
        std::vector<MapPoint*> vp(M, nullptr);
        for (int i = 0; i < M; ++i) {
            MapPoint& m = mps[i];
            m.pos = cv::Mat(3, 1, CV_32FC1); get_n((float*)m.pos.data, 3);
            m.normal = cv::Mat(3, 1, CV_32FC1); get_n((float*)m.normal.data, 3);
            m.mfMinDistance = get<float>(); m.mfMaxDistance = get<float>();
            m.desc = get_desc_rows(1);
            const int present = get<uint8_t>(); m.bad = get<uint8_t>() != 0; m.nobs = get<uint8_t>(); const int inkf = get<uint8_t>();
            m.id = i;
            if (inkf) m.mObservations[&K] = 0;
            if (present) vp[i] = &m;
        }
        g_fuse_kf = &K;
        g_fuse_log.clear();
        ORBmatcher matcher(0.6f, true);
        const int nf = matcher.Fuse(&K, vp, th);
        put<int32_t>(nf);
        put<int32_t>((int32_t)g_fuse_log.size());
        for (const FuseEvent& e : g_fuse_log) { put<int32_t>(e.mp); put<int32_t>(e.idx); put<int32_t>(e.action); }
    } else if (op == 12) {   // ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)   (loop detection)
        Frame F; std::vector<MapPoint> pool;
        read_point_frame(F, pool);               // claimed[i] = vpMatched[i] is set at call time
        KeyFrame K; keyframe_from_frame(F, K);
        read_kf_camera(K);
        cv::Mat Scw = get_pose();
        const int th = get<int32_t>();
        const int M = get<int32_t>();
        std::vector<MapPoint*> vpMatched(K.N, nullptr);
        for (int i = 0; i < K.N; ++i) if (F.mvpMapPoints[i]) { pool[i].id = -100 - i; vpMatched[i] = &pool[i]; }
        std::vector<MapPoint> mps(M);
        std::vector<MapPoint*> vp(M);
        for (int i = 0; i < M; ++i) {
            uint8_t fl[4]; int32_t slot;
            read_map_point(mps[i], fl, slot);    // flags: -, bad, -, already in vpMatched (at `slot`)
            mps[i].id = i; mps[i].bad = fl[1] != 0; vp[i] = &mps[i];
            if (fl[3] && slot >= 0 && slot < K.N) vpMatched[slot] = &mps[i];
        }
        ORBmatcher matcher(0.75f, true);
        const int nm = matcher.SearchByProjection(&K, Scw, vp, vpMatched, th);
        put<int32_t>(nm);
        for (int i = 0; i < K.N; ++i) put<int32_t>(vpMatched[i] ? vpMatched[i]->id : -1);
    } else if (op == 13) {   // ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)   (loop closing)
        Frame F; std::vector<MapPoint> pool;
        read_point_frame(F, pool);               // claimed[i] = the key frame's keypoint holds a map point
        KeyFrame K; keyframe_from_frame(F, K);
        std::vector<uint8_t> kfbad(K.N); get_n(kfbad.data(), K.N);
        for (int i = 0; i < K.N; ++i) if (F.mvpMapPoints[i]) { pool[i].id = -100 - i; pool[i].bad = kfbad[i] != 0; K.mvpMapPoints[i] = &pool[i]; }
        read_kf_camera(K);
        cv::Mat Scw = get_pose();
        const float th = get<float>();
        const int M = get<int32_t>();
        std::vector<MapPoint> mps(M);
        std::vector<MapPoint*> vp(M);
        for (int i = 0; i < M; ++i) {
            uint8_t fl[4]; int32_t slot;
            read_map_point(mps[i], fl, slot);    // flags: -, bad, -, held by the key frame (at `slot`)
            mps[i].id = i; mps[i].bad = fl[1] != 0; vp[i] = &mps[i];
            if (fl[3] && slot >= 0 && slot < K.N) K.mvpMapPoints[slot] = &mps[i];
        }
        std::vector<MapPoint*> vpReplace(M, nullptr);
        g_fuse_kf = &K;
        g_fuse_log.clear();
        ORBmatcher matcher(0.8f, true);
        const int nf = matcher.Fuse(&K, Scw, vp, th, vpReplace);
        put<int32_t>(nf);
        for (int i = 0; i < M; ++i) put<int32_t>(vpReplace[i] ? vpReplace[i]->id : -1);
        put<int32_t>((int32_t)g_fuse_log.size());
        for (const FuseEvent& e : g_fuse_log) { put<int32_t>(e.mp); put<int32_t>(e.idx); }
        for (int i = 0; i < K.N; ++i) put<int32_t>(K.mvpMapPoints[i] ? K.mvpMapPoints[i]->id : -1);
    } else if (op == 14) {   // ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)
        Frame F1, F2; std::vector<MapPoint> pool1, pool2, mps1, mps2;
        KeyFrame K1, K2;
        read_point_frame(F1, pool1); keyframe_from_frame(F1, K1); read_kf_camera(K1); read_kf_map_points(K1, mps1, nullptr);
        read_point_frame(F2, pool2); keyframe_from_frame(F2, K2); read_kf_camera(K2); read_kf_map_points(K2, mps2, nullptr);
        std::vector<int32_t> m12(K1.N); get_n(m12.data(), K1.N);
        std::vector<MapPoint*> vpMatches12(K1.N, nullptr);
        for (int i = 0; i < K1.N; ++i) if (m12[i] >= 0) vpMatches12[i] = &mps2[m12[i]];
        const float s12 = get<float>();
        cv::Mat R12(3, 3, CV_32FC1); get_n((float*)R12.data, 9);
        cv::Mat t12(3, 1, CV_32FC1); get_n((float*)t12.data, 3);
        const float th = get<float>();
        ORBmatcher matcher(0.75f, true);
        const int nf = matcher.SearchBySim3(&K1, &K2, vpMatches12, s12, R12, t12, th);
        put<int32_t>(nf);
        for (int i = 0; i < K1.N; ++i) put<int32_t>(vpMatches12[i] ? vpMatches12[i]->id : -1);
    } else if (op == 15) {   // ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)   (relocalisation)
        Frame C; std::vector<MapPoint> pool, mps;
        read_point_frame(C, pool);               // claimed[i] = the frame's keypoint holds a map point
        float cam[4]; get_n(cam, 4);
        Frame::fx = cam[0]; Frame::fy = cam[1]; Frame::cx = cam[2]; Frame::cy = cam[3];
        C.mTcw = get_pose();
        C.mfLogScaleFactor = get<float>(); C.mnScaleLevels = get<int32_t>();
        const float th = get<float>();
        const int ORBdist = get<int32_t>(), check_ori = get<int32_t>();
        KeyFrame K;
        K.N = get<int32_t>();
        K.mvKeysUn.resize(K.N); get_n(K.mvKeysUn.data(), K.N);
        K.mvpMapPoints.assign(K.N, nullptr);
        std::vector<uint8_t> found;
        read_kf_map_points(K, mps, &found);
        std::set<MapPoint*> sFound;
        for (int i = 0; i < K.N; ++i) if (found[i]) sFound.insert(&mps[i]);
        ORBmatcher matcher(0.9f, check_ori != 0);
        const int nm = matcher.SearchByProjection(C, &K, sFound, th, ORBdist);
        put<int32_t>(nm);
        for (int i = 0; i < C.N; ++i) put<int32_t>(C.mvpMapPoints[i] ? C.mvpMapPoints[i]->id : -1);
    } else if (op == 16) {   // ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12)
        KeyFrame K1, K2; std::vector<MapPoint> mps1, mps2;
        auto read_kf = [&](KeyFrame& K, std::vector<MapPoint>& mps) {
            K.N = get<int32_t>();
            K.mvKeysUn.resize(K.N); get_n(K.mvKeysUn.data(), K.N);
            K.mDescriptors = get_desc_rows(K.N);
            std::vector<uint8_t> has(K.N), bad(K.N); get_n(has.data(), K.N); get_n(bad.data(), K.N);
            mps.resize(K.N);
            K.mvpMapPoints.assign(K.N, nullptr);
            for (int i = 0; i < K.N; ++i) { mps[i].id = i; mps[i].bad = bad[i] != 0; if (has[i]) K.mvpMapPoints[i] = &mps[i]; }
            read_featvec(K.mFeatVec);
        };
        read_kf(K1, mps1); read_kf(K2, mps2);
        const float nnratio = get<float>();
        const int check_ori = get<int32_t>();
        std::vector<MapPoint*> m12;
        ORBmatcher matcher(nnratio, check_ori != 0);
        const int nm = matcher.SearchByBoW(&K1, &K2, m12);
        put<int32_t>(nm);
        for (int i = 0; i < K1.N; ++i) put<int32_t>(m12[i] ? m12[i]->id : -1);
    } else if (op == 17) {   // LSDmatcher::FrameBFMatch, match / matchNNR, SearchDouble(InitialFrame, CurrentFrame, LineMatches)
        Frame F1, F2;
        F1.NL = get<int32_t>(); F1.mLdesc = get_desc_rows(F1.NL);
        F2.NL = get<int32_t>(); F2.mLdesc = get_desc_rows(F2.NL);
        if (F1.mLdesc.rows != F1.NL) F1.mLdesc = F1.mLdesc.rowRange(0, F1.NL);
        if (F2.mLdesc.rows != F2.NL) F2.mLdesc = F2.mLdesc.rowRange(0, F2.NL);
        const float TH = get<float>(), nnratio = get<float>(), nnr = get<float>();
        LSDmatcher matcher(nnratio, true);
        std::vector<int> lm, m12, dbl;
        matcher.FrameBFMatch(F1.mLdesc, F2.mLdesc, lm, TH);
        const int n_nnr = matcher.match(F1.mLdesc, F2.mLdesc, nnr, m12);
        const int n_dbl = matcher.SearchDouble(F1, F2, dbl);
        for (int i = 0; i < F1.NL; ++i) put<int32_t>(lm[i]);
        put<int32_t>(n_nnr);
        for (int i = 0; i < F1.NL; ++i) put<int32_t>(m12[i]);
        put<int32_t>(n_dbl);
        for (int i = 0; i < F1.NL; ++i) put<int32_t>(dbl[i]);
    } else if (op == 18) {   // LSDmatcher::SearchByDescriptor(pKF, currentF, vpMapLineMatches), SearchDouble(KF, CurrentFrame)
        KeyFrame K; Frame F;
        K.NL = get<int32_t>(); K.mLineDescriptors = get_desc_rows(K.NL);
        if (K.mLineDescriptors.rows != K.NL) K.mLineDescriptors = K.mLineDescriptors.rowRange(0, K.NL);
        std::vector<uint8_t> has(K.NL); get_n(has.data(), K.NL);
        std::vector<MapLine> mls(K.NL);
        K.mvpMapLines.assign(K.NL, nullptr);
        for (int i = 0; i < K.NL; ++i) { mls[i].id = i; if (has[i]) K.mvpMapLines[i] = &mls[i]; }
        F.NL = get<int32_t>(); F.mLdesc = get_desc_rows(F.NL);
        if (F.mLdesc.rows != F.NL) F.mLdesc = F.mLdesc.rowRange(0, F.NL);
        F.mvpMapLines.assign(F.NL, nullptr);
        const float nnratio = get<float>();
        LSDmatcher matcher(nnratio, true);
        std::vector<MapLine*> byDesc;
        const int n1 = matcher.SearchByDescriptor(&K, F, byDesc);
        const int n2 = matcher.SearchDouble(&K, F);
        put<int32_t>(n1);
        for (int i = 0; i < F.NL; ++i) put<int32_t>(byDesc[i] ? byDesc[i]->id : -1);
        put<int32_t>(n2);
        for (int i = 0; i < F.NL; ++i) put<int32_t>(F.mvpMapLines[i] ? F.mvpMapLines[i]->id : -1);
    } else if (op == 19) {   // LSDmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs) and (pKF1, pKF2, vMatchedPairs, isDouble) x2
        KeyFrame K1, K2;
        std::vector<MapLine> ml1, ml2;
        auto read_kf = [&](KeyFrame& K, std::vector<MapLine>& mls) {
            K.NL = get<int32_t>(); K.mLineDescriptors = get_desc_rows(K.NL);
            if (K.mLineDescriptors.rows != K.NL) K.mLineDescriptors = K.mLineDescriptors.rowRange(0, K.NL);
            std::vector<uint8_t> has(K.NL); get_n(has.data(), K.NL);
            mls.resize(K.NL);
            K.mvpMapLines.assign(K.NL, nullptr);
            for (int i = 0; i < K.NL; ++i) { mls[i].id = i; if (has[i]) K.mvpMapLines[i] = &mls[i]; }
        };
        read_kf(K1, ml1); read_kf(K2, ml2);
        const float nnratio = get<float>();
        LSDmatcher matcher(nnratio, true);
        std::vector<pair<size_t, size_t>> pairs;
        const int n0 = matcher.SearchForTriangulation(&K1, &K2, pairs);
        put<int32_t>(n0);
        put<int32_t>((int32_t)pairs.size());
        for (const auto& pr : pairs) { put<int32_t>((int32_t)pr.first); put<int32_t>((int32_t)pr.second); }
        for (int dbl = 0; dbl < 2; ++dbl) {
            std::vector<int> m;
            const int n = matcher.SearchForTriangulation(&K1, &K2, m, dbl != 0);
            put<int32_t>(n);
            for (int i = 0; i < K1.NL; ++i) put<int32_t>(m[i]);
        }
    } else {
        return 5;
    }
    std::fclose(g_in);
    std::fclose(g_out);
    return 0;
}
