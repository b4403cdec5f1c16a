// Drop-in ORB_SLAM2::ORBmatcher for the projection searches (reference include/ORBmatcher.h:38-77, src/ORBmatcher.cc), as a class
// template over the reference's own Frame / MapPoint types so that it compiles unchanged against them (and against the stand-ins of
// the tests):
//
//     #include "shim/ORBmatcher.h"
//     namespace ORB_SLAM2 { typedef hvo_shim::ORBmatcherT<Frame, MapPoint> ORBmatcher; }      // instead of include/ORBmatcher.h
//
// Same constructor, constants, method names, argument meaning and side effects (F.mvpMapPoints is written in place, the return value
// is nmatches).  The members it reads are the reference's: F.N, mvKeysUn, mvKeys, mvuRight, mDescriptors, mvpMapPoints, mvbOutlier,
// mvScaleFactors, mTcw, mb, mbf, Frame::fx/fy/cx/cy/mnMinX..; pMP->mbTrackInView, mTrackProjX/Y/XR, mnTrackScaleLevel, mTrackViewCos,
// isBad(), Observations(), GetDescriptor(), GetWorldPos().  The candidate loops run on the GPU through the C ABI (hvo_proj_*); what
// stays here is what the reference also does on the host around them: building one query per map point, the pose algebra of the
// last-frame search (cv::Mat expressions as written in the reference, so they evaluate identically), applying the assignment in
// query order and the rotation histogram.
//
//   SearchByProjection(F, vpMapPoints, th)                         src/ORBmatcher.cc:45-132
//   SearchByProjection(CurrentFrame, LastFrame, th, bMono)         src/ORBmatcher.cc:1353-1497
//   SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, w) src/ORBmatcher.cc:412-529
//   SearchByBoW(pKF, F, vpMapPointMatches)                         src/ORBmatcher.cc:162-293    (KeyFrame type deduced)
//   SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)  src/ORBmatcher.cc:1499-1628  (relocalisation)
//   SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)          src/ORBmatcher.cc:295-410    (loop detection)
//   SearchByBoW(pKF1, pKF2, vpMatches12)                           src/ORBmatcher.cc:531-666
//   SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, stereo) src/ORBmatcher.cc:668-836    (+ CheckDistEpipolarLine :143-160 on the device)
//   SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)       src/ORBmatcher.cc:1123-1351
//   Fuse(pKF, vpMapPoints, th)                                     src/ORBmatcher.cc:838-994
//   Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)                   src/ORBmatcher.cc:996-1121
//   DescriptorDistance                                             src/ORBmatcher.cc:1676-1692
// i.e. every public method of include/ORBmatcher.h:38-77.  The KeyFrame type is deduced; what is read of it is the reference's:
// N, fx, fy, cx, cy, mbf, mvKeysUn, mvuRight, mDescriptors, mFeatVec, mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2, mnMinX / mnMinY
// (the INTEGER origin KeyFrame::GetFeaturesInArea uses, hvo_proj_set_window_origin), GetMapPointMatches(), GetMapPoint(), GetMapPoints(),
// GetRotation(), GetTranslation(), GetCameraCenter(), IsInImage(), AddMapPoint(); of a map point GetWorldPos(), GetNormal(),
// Get{Min,Max}DistanceInvariance(), PredictScale(), IsInKeyFrame(), GetIndexInKeyFrame(), Replace(), AddObservation().  The key frame's
// grid is rebuilt on the device from its keypoints and the Frame's (static) float bounds, which is what the KeyFrame copied.
#ifndef HVO_SHIM_ORBMATCHER_H
#define HVO_SHIM_ORBMATCHER_H

#include <cmath>
#include <cstdio>
#include <climits>
#include <cstring>
#include <set>
#include <utility>
#include <vector>

#include "hvo_capi.h"

namespace hvo_shim {

template <class Frame, class MapPoint>
class ORBmatcherT {
public:
    static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;   // src/ORBmatcher.cc:37-39

    ORBmatcherT(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri), h_(nullptr) {
        if (hvo_proj_create(device, &h_) != HVO_OK) std::fprintf(stderr, "ORBmatcher: %s\n", hvo_last_error());
    }
    ~ORBmatcherT() { hvo_proj_destroy(h_); }
    ORBmatcherT(const ORBmatcherT&) = delete;
    ORBmatcherT& operator=(const ORBmatcherT&) = delete;

    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return hvo_hamming_distance(a.template ptr<uint8_t>(), b.template ptr<uint8_t>()); }

    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3) {
        if (!setFrame(F)) return 0;
        const bool bFactor = th != 1.0;
        std::vector<hvo_proj_query> q;
        std::vector<uint8_t> qdesc;
        std::vector<size_t> who;
        for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
            MapPoint* pMP = vpMapPoints[iMP];
            if (!pMP->mbTrackInView || pMP->isBad()) continue;
            const int level = pMP->mnTrackScaleLevel;
            float r = RadiusByViewingCos(pMP->mTrackViewCos);
            if (bFactor) r *= th;
            hvo_proj_query e;
            e.u = pMP->mTrackProjX; e.v = pMP->mTrackProjY; e.r = r * F.mvScaleFactors[level];
            e.min_level = level - 1; e.max_level = level; e.ur = pMP->mTrackProjXR;
            e.claims = pMP->Observations() > 0; e.reserved = 0;
            q.push_back(e);
            appendDescriptor(qdesc, pMP->GetDescriptor());
            who.push_back(iMP);
        }
        std::vector<int32_t> idx(q.size(), -1);
        int n = 0;
        const std::vector<uint8_t> claimed = claimedOf(F);
        if (!q.empty() && !ok(hvo_proj_search(h_, q.data(), qdesc.data(), (int)q.size(), claimed.data(), 0, TH_HIGH, mfNNratio, idx.data(), nullptr, &n)))
            return 0;
        for (size_t k = 0; k < q.size(); ++k)
            if (idx[k] >= 0) F.mvpMapPoints[idx[k]] = vpMapPoints[who[k]];
        return n;
    }

    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
        if (!setFrame(CurrentFrame)) return 0;
        const float factor = 1.0f / HISTO_LENGTH;
        // pose algebra exactly as the reference writes it (:1364-1375), so the cv::Mat expressions evaluate the same way
        const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
        const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
        const cv::Mat twc = -Rcw.t() * tcw;
        const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
        const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
        const cv::Mat tlc = Rlw * twc + tlw;
        const bool bForward = tlc.template at<float>(2) > CurrentFrame.mb && !bMono;
        const bool bBackward = -tlc.template at<float>(2) > CurrentFrame.mb && !bMono;

        std::vector<hvo_proj_query> q;
        std::vector<uint8_t> qdesc;
        std::vector<int> who;
        for (int i = 0; i < LastFrame.N; i++) {
            MapPoint* pMP = LastFrame.mvpMapPoints[i];
            if (!pMP || LastFrame.mvbOutlier[i]) continue;
            cv::Mat x3Dw = pMP->GetWorldPos();
            cv::Mat x3Dc = Rcw * x3Dw + tcw;
            const float xc = x3Dc.template at<float>(0), yc = x3Dc.template at<float>(1);
            const float invzc = 1.0 / x3Dc.template at<float>(2);
            if (invzc < 0) continue;
            const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
            const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
            if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
            if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
            const int nLastOctave = LastFrame.mvKeys[i].octave;
            hvo_proj_query e;
            e.u = u; e.v = v; e.r = th * CurrentFrame.mvScaleFactors[nLastOctave];
            if (bForward) { e.min_level = nLastOctave; e.max_level = -1; }
            else if (bBackward) { e.min_level = 0; e.max_level = nLastOctave; }
            else { e.min_level = nLastOctave - 1; e.max_level = nLastOctave + 1; }
            e.ur = u - CurrentFrame.mbf * invzc;
            e.claims = pMP->Observations() > 0; e.reserved = 0;
            q.push_back(e);
            appendDescriptor(qdesc, pMP->GetDescriptor());
            who.push_back(i);
        }
        std::vector<int32_t> idx(q.size(), -1);
        int nmatches = 0;
        const std::vector<uint8_t> claimed = claimedOf(CurrentFrame);
        if (!q.empty() && !ok(hvo_proj_search(h_, q.data(), qdesc.data(), (int)q.size(), claimed.data(), 1, TH_HIGH, mfNNratio, idx.data(), nullptr, &nmatches)))
            return 0;
        std::vector<int> rotHist[HISTO_LENGTH];
        for (size_t k = 0; k < q.size(); ++k) {
            if (idx[k] < 0) continue;
            CurrentFrame.mvpMapPoints[idx[k]] = LastFrame.mvpMapPoints[who[k]];
            if (mbCheckOrientation) rotHist[rotationBin(LastFrame.mvKeysUn[who[k]].angle - CurrentFrame.mvKeysUn[idx[k]].angle, factor)].push_back(idx[k]);
        }
        if (mbCheckOrientation) {   // :1458-1494
            int ind1 = -1, ind2 = -1, ind3 = -1;
            ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++)
                if (i != ind1 && i != ind2 && i != ind3)
                    for (size_t j = 0; j < rotHist[i].size(); j++) {
                        CurrentFrame.mvpMapPoints[rotHist[i][j]] = static_cast<MapPoint*>(NULL);
                        nmatches--;
                    }
        }
        return nmatches;
    }

    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10) {
        const int n1 = (int)F1.mvKeysUn.size();
        vnMatches12 = std::vector<int>(n1, -1);
        if (n1 == 0 || !setFrame(F2)) return 0;
        std::vector<float> prev((size_t)n1 * 2);
        std::vector<int32_t> octave(n1), m12(n1, -1), accepted(n1, -1);
        std::vector<uint8_t> d1;
        for (int i = 0; i < n1; ++i) {
            prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y;
            octave[i] = F1.mvKeysUn[i].octave;
            appendDescriptor(d1, F1.mDescriptors.row(i));
        }
        int nmatches = 0;
        if (!ok(hvo_proj_search_initialization(h_, prev.data(), octave.data(), d1.data(), n1, windowSize, TH_LOW, mfNNratio, m12.data(), accepted.data(),
                                               &nmatches)))
            return 0;
        for (int i = 0; i < n1; ++i) vnMatches12[i] = m12[i];
        if (mbCheckOrientation) {   // :499-523: every acceptance was pushed, taken over or not
            const float factor = 1.0f / HISTO_LENGTH;
            std::vector<int> rotHist[HISTO_LENGTH];
            for (int i1 = 0; i1 < n1; ++i1)
                if (accepted[i1] >= 0) rotHist[rotationBin(F1.mvKeysUn[i1].angle - F2.mvKeysUn[accepted[i1]].angle, factor)].push_back(i1);
            int ind1 = -1, ind2 = -1, ind3 = -1;
            ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++)
                if (i != ind1 && i != ind2 && i != ind3)
                    for (size_t j = 0; j < rotHist[i].size(); j++) {
                        const int idx1 = rotHist[i][j];
                        if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
                    }
        }
        for (size_t i1 = 0; i1 < vnMatches12.size(); i1++)
            if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2.mvKeysUn[vnMatches12[i1]].pt;
        return nmatches;
    }

    // src/ORBmatcher.cc:162-293.  KeyFrame is the reference's own type (deduced): GetMapPointMatches(), mFeatVec, mvKeysUn, mDescriptors;
    // F.mFeatVec, F.mvKeys, F.mDescriptors.  The walk over the two DBoW2::FeatureVectors (std::map, equal node ids) stays here and builds
    // the queries in the reference's visiting order; the greedy candidate loop runs on the device (hvo_proj_search_candidates).
    template <class KeyFrame>
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
        const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
        vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
        const auto& vFeatVecKF = pKF->mFeatVec;
        std::vector<uint8_t> qdesc, tdesc;
        std::vector<int32_t> offsets(1, 0), cand, who;
        auto KFit = vFeatVecKF.begin(), KFend = vFeatVecKF.end();
        auto Fit = F.mFeatVec.begin(), Fend = F.mFeatVec.end();
        while (KFit != KFend && Fit != Fend) {
            if (KFit->first == Fit->first) {
                for (size_t iKF = 0; iKF < KFit->second.size(); iKF++) {
                    const unsigned int realIdxKF = KFit->second[iKF];
                    MapPoint* pMP = vpMapPointsKF[realIdxKF];
                    if (!pMP || pMP->isBad()) continue;
                    appendDescriptor(qdesc, pKF->mDescriptors.row(realIdxKF));
                    for (size_t iF = 0; iF < Fit->second.size(); iF++) cand.push_back((int32_t)Fit->second[iF]);
                    offsets.push_back((int32_t)cand.size());
                    who.push_back((int32_t)realIdxKF);
                }
                KFit++; Fit++;
            } else if (KFit->first < Fit->first) KFit = vFeatVecKF.lower_bound(Fit->first);
            else Fit = F.mFeatVec.lower_bound(KFit->first);
        }
        const int nq = (int)who.size();
        if (nq == 0 || F.N == 0 || !h_) return 0;
        for (int i = 0; i < F.N; ++i) appendDescriptor(tdesc, F.mDescriptors.row(i));
        std::vector<int32_t> idx(nq, -1);
        int nmatches = 0;
        if (!ok(hvo_proj_search_candidates(h_, qdesc.data(), nq, tdesc.data(), F.N, offsets.data(), cand.data(), TH_LOW, mfNNratio, idx.data(), nullptr,
                                           &nmatches)))
            return 0;
        const float factor = 1.0f / HISTO_LENGTH;
        std::vector<int> rotHist[HISTO_LENGTH];
        for (int k = 0; k < nq; ++k) {
            if (idx[k] < 0) continue;
            vpMapPointMatches[idx[k]] = vpMapPointsKF[who[k]];
            if (mbCheckOrientation) rotHist[rotationBin(pKF->mvKeysUn[who[k]].angle - F.mvKeys[idx[k]].angle, factor)].push_back(idx[k]);
        }
        if (mbCheckOrientation) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++)
                if (i != ind1 && i != ind2 && i != ind3)
                    for (size_t j = 0; j < rotHist[i].size(); j++) { vpMapPointMatches[rotHist[i][j]] = static_cast<MapPoint*>(NULL); nmatches--; }
        }
        return nmatches;
    }

    // src/ORBmatcher.cc:1499-1628 (relocalisation).  The projection of the key frame's map points is the reference's own cv::Mat algebra;
    // the windowed search (levels [l-1, l+1], keypoints that hold any map point are skipped and every match claims its keypoint, best
    // distance <= ORBdist, no right-coordinate check) runs on the device; the rotation histogram is applied here.
    template <class KeyFrame>
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) {
        if (!setFrame(CurrentFrame, false)) return 0;
        const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
        const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
        const cv::Mat Ow = -Rcw.t() * tcw;
        const float factor = 1.0f / HISTO_LENGTH;
        const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
        std::vector<hvo_proj_query> q;
        std::vector<uint8_t> qdesc;
        std::vector<size_t> who;
        for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {
            MapPoint* pMP = vpMPs[i];
            if (!pMP) continue;
            if (pMP->isBad() || sAlreadyFound.count(pMP)) continue;
            cv::Mat x3Dw = pMP->GetWorldPos();
            cv::Mat x3Dc = Rcw * x3Dw + tcw;
            const float xc = x3Dc.template at<float>(0), yc = x3Dc.template at<float>(1);
            const float invzc = 1.0 / x3Dc.template at<float>(2);
            const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
            const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
            if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX) continue;
            if (v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
            cv::Mat PO = x3Dw - Ow;
            float dist3D = cv::norm(PO);
            const float maxDistance = pMP->GetMaxDistanceInvariance();
            const float minDistance = pMP->GetMinDistanceInvariance();
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            int nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame);
            hvo_proj_query e;
            e.u = u; e.v = v; e.r = th * CurrentFrame.mvScaleFactors[nPredictedLevel];
            e.min_level = nPredictedLevel - 1; e.max_level = nPredictedLevel + 1;
            e.ur = -1.f; e.claims = 1; e.reserved = 0;
            q.push_back(e);
            appendDescriptor(qdesc, pMP->GetDescriptor());
            who.push_back(i);
        }
        std::vector<uint8_t> claimed(CurrentFrame.mvpMapPoints.size() ? CurrentFrame.mvpMapPoints.size() : 1, 0);
        for (size_t i = 0; i < CurrentFrame.mvpMapPoints.size(); ++i) claimed[i] = CurrentFrame.mvpMapPoints[i] != NULL;   // :1568-1569
        std::vector<int32_t> idx(q.size(), -1);
        int nmatches = 0;
        if (!q.empty() && !ok(hvo_proj_search(h_, q.data(), qdesc.data(), (int)q.size(), claimed.data(), 1, ORBdist, mfNNratio, idx.data(), nullptr, &nmatches)))
            return 0;
        std::vector<int> rotHist[HISTO_LENGTH];
        for (size_t k = 0; k < q.size(); ++k) {
            if (idx[k] < 0) continue;
            CurrentFrame.mvpMapPoints[idx[k]] = vpMPs[who[k]];
            if (mbCheckOrientation) rotHist[rotationBin(pKF->mvKeysUn[who[k]].angle - CurrentFrame.mvKeysUn[idx[k]].angle, factor)].push_back(idx[k]);
        }
        if (mbCheckOrientation) {   // :1594-1622
            int ind1 = -1, ind2 = -1, ind3 = -1;
            ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++)
                if (i != ind1 && i != ind2 && i != ind3)
                    for (size_t j = 0; j < rotHist[i].size(); j++) {
                        CurrentFrame.mvpMapPoints[rotHist[i][j]] = static_cast<MapPoint*>(NULL);
                        nmatches--;
                    }
        }
        return nmatches;
    }

    // src/ORBmatcher.cc:295-410 (loop detection): candidates projected with the Sim3 Scw; key-frame features already in vpMatched are skipped
    // and every match enters it (so later candidates skip it too); best distance <= TH_LOW.
    template <class KeyFrame>
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th) {
        if (!setKeyFrame(pKF, false)) return 0;
        Sim3Camera<KeyFrame> cam(pKF, Scw);
        std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
        spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
        std::vector<hvo_proj_query> q;
        std::vector<uint8_t> qdesc;
        std::vector<int> who;
        for (int iMP = 0, iendMP = (int)vpPoints.size(); iMP < iendMP; iMP++) {
            MapPoint* pMP = vpPoints[iMP];
            if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
            hvo_proj_query e;
            if (!cam.project(pMP, (float)th, true, e)) continue;
            e.claims = 1;
            q.push_back(e);
            appendDescriptor(qdesc, pMP->GetDescriptor());
            who.push_back(iMP);
        }
        std::vector<uint8_t> claimed(vpMatched.size() ? vpMatched.size() : 1, 0);
        for (size_t i = 0; i < vpMatched.size(); ++i) claimed[i] = vpMatched[i] != NULL;
        std::vector<int32_t> idx(q.size(), -1);
        int nmatches = 0;
        if (!q.empty() && !ok(hvo_proj_search(h_, q.data(), qdesc.data(), (int)q.size(), claimed.data(), 1, TH_LOW, mfNNratio, idx.data(), nullptr, &nmatches)))
            return 0;
        for (size_t k = 0; k < q.size(); ++k)
            if (idx[k] >= 0) vpMatched[idx[k]] = vpPoints[who[k]];
        return nmatches;
    }

    // src/ORBmatcher.cc:531-666: like SearchByBoW(pKF, F, ...) between two key frames; the candidates of a vocabulary node are pKF2's
    // features that hold a good map point, the acceptance is bestDist1 < TH_LOW (strict).
    template <class KeyFrame>
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
        const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
        const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
        vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
        const auto& vFeatVec1 = pKF1->mFeatVec;
        const auto& vFeatVec2 = pKF2->mFeatVec;
        std::vector<uint8_t> good2(vpMapPoints2.size() ? vpMapPoints2.size() : 1, 0);
        for (size_t i = 0; i < vpMapPoints2.size(); ++i) good2[i] = vpMapPoints2[i] && !vpMapPoints2[i]->isBad();
        std::vector<uint8_t> qdesc, tdesc;
        std::vector<int32_t> offsets(1, 0), cand, who;
        auto f1it = vFeatVec1.begin(), f1end = vFeatVec1.end();
        auto f2it = vFeatVec2.begin(), f2end = vFeatVec2.end();
        while (f1it != f1end && f2it != f2end) {
            if (f1it->first == f2it->first) {
                for (size_t i1 = 0; i1 < f1it->second.size(); i1++) {
                    const size_t idx1 = f1it->second[i1];
                    MapPoint* pMP1 = vpMapPoints1[idx1];
                    if (!pMP1 || pMP1->isBad()) continue;
                    appendDescriptor(qdesc, pKF1->mDescriptors.row((int)idx1));
                    for (size_t i2 = 0; i2 < f2it->second.size(); i2++)
                        if (good2[f2it->second[i2]]) cand.push_back((int32_t)f2it->second[i2]);
                    offsets.push_back((int32_t)cand.size());
                    who.push_back((int32_t)idx1);
                }
                f1it++; f2it++;
            } else if (f1it->first < f2it->first) f1it = vFeatVec1.lower_bound(f2it->first);
            else f2it = vFeatVec2.lower_bound(f1it->first);
        }
        const int nq = (int)who.size(), n2 = (int)vpMapPoints2.size();
        if (nq == 0 || n2 == 0 || !h_) return 0;
        for (int i = 0; i < n2; ++i) appendDescriptor(tdesc, pKF2->mDescriptors.row(i));
        std::vector<int32_t> idx(nq, -1);
        int nmatches = 0;
        if (!ok(hvo_proj_search_candidates(h_, qdesc.data(), nq, tdesc.data(), n2, offsets.data(), cand.data(), TH_LOW - 1, mfNNratio, idx.data(), nullptr,
                                           &nmatches)))
            return 0;
        const float factor = 1.0f / HISTO_LENGTH;
        std::vector<int> rotHist[HISTO_LENGTH];
        for (int k = 0; k < nq; ++k) {
            if (idx[k] < 0) continue;
            vpMatches12[who[k]] = vpMapPoints2[idx[k]];
            if (mbCheckOrientation) rotHist[rotationBin(pKF1->mvKeysUn[who[k]].angle - pKF2->mvKeysUn[idx[k]].angle, factor)].push_back(who[k]);
        }
        if (mbCheckOrientation) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++)
                if (i != ind1 && i != ind2 && i != ind3)
                    for (size_t j = 0; j < rotHist[i].size(); j++) { vpMatches12[rotHist[i][j]] = static_cast<MapPoint*>(NULL); nmatches--; }
        }
        return nmatches;
    }

    // src/ORBmatcher.cc:668-836.  The epipole is the reference's cv::Mat algebra (:676-683); queries = pKF1's features without a map point in
    // (vocabulary node, index list) order, candidates = pKF2's features of the same node; the descriptor / epipole-distance / epipolar-line
    // gates (CheckDistEpipolarLine :143-160) run on the device; rotation histogram and vMatchedPairs here.
    template <class KeyFrame>
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
        const auto& vFeatVec1 = pKF1->mFeatVec;
        const auto& vFeatVec2 = pKF2->mFeatVec;
        cv::Mat Cw = pKF1->GetCameraCenter();
        cv::Mat R2w = pKF2->GetRotation();
        cv::Mat t2w = pKF2->GetTranslation();
        cv::Mat C2 = R2w * Cw + t2w;
        const float invz = 1.0f / C2.template at<float>(2);
        const float ex = pKF2->fx * C2.template at<float>(0) * invz + pKF2->cx;
        const float ey = pKF2->fy * C2.template at<float>(1) * invz + pKF2->cy;
        vMatchedPairs.clear();
        std::vector<uint8_t> qdesc, qstereo, tdesc, tflags;
        std::vector<hvo_keypoint> qkeys;
        std::vector<int32_t> offsets(1, 0), cand, who;
        auto f1it = vFeatVec1.begin(), f1end = vFeatVec1.end();
        auto f2it = vFeatVec2.begin(), f2end = vFeatVec2.end();
        while (f1it != f1end && f2it != f2end) {
            if (f1it->first == f2it->first) {
                for (size_t i1 = 0; i1 < f1it->second.size(); i1++) {
                    const size_t idx1 = f1it->second[i1];
                    if (pKF1->GetMapPoint(idx1)) continue;
                    const bool bStereo1 = pKF1->mvuRight[idx1] >= 0;
                    if (bOnlyStereo && !bStereo1) continue;
                    appendDescriptor(qdesc, pKF1->mDescriptors.row((int)idx1));
                    qkeys.push_back(reinterpret_cast<const hvo_keypoint&>(pKF1->mvKeysUn[idx1]));
                    qstereo.push_back(bStereo1);
                    for (size_t i2 = 0; i2 < f2it->second.size(); i2++) cand.push_back((int32_t)f2it->second[i2]);
                    offsets.push_back((int32_t)cand.size());
                    who.push_back((int32_t)idx1);
                }
                f1it++; f2it++;
            } else if (f1it->first < f2it->first) f1it = vFeatVec1.lower_bound(f2it->first);
            else f2it = vFeatVec2.lower_bound(f1it->first);
        }
        const int nq = (int)who.size(), n2 = (int)pKF2->mvKeysUn.size();
        if (nq == 0 || n2 == 0 || !h_) return 0;
        for (int i = 0; i < n2; ++i) {
            appendDescriptor(tdesc, pKF2->mDescriptors.row(i));
            tflags.push_back((uint8_t)((pKF2->GetMapPoint(i) ? 1 : 0) | (pKF2->mvuRight[i] >= 0 ? 2 : 0)));
        }
        float f12[9];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) f12[3 * r + c] = F12.template at<float>(r, c);
        std::vector<int32_t> idx(nq, -1);
        int nmatches = 0;
        if (!ok(hvo_proj_search_triangulation(h_, qdesc.data(), qkeys.data(), qstereo.data(), nq, tdesc.data(),
                                              reinterpret_cast<const hvo_keypoint*>(pKF2->mvKeysUn.data()), tflags.data(), n2, offsets.data(), cand.data(), f12,
                                              ex, ey, pKF2->mvScaleFactors.data(), pKF2->mvLevelSigma2.data(), (int)pKF2->mvScaleFactors.size(),
                                              bOnlyStereo ? 1 : 0, TH_LOW, idx.data(), nullptr, &nmatches)))
            return 0;
        std::vector<int> vMatches12(pKF1->mvKeysUn.size(), -1);
        const float factor = 1.0f / HISTO_LENGTH;
        std::vector<int> rotHist[HISTO_LENGTH];
        for (int k = 0; k < nq; ++k) {
            if (idx[k] < 0) continue;
            vMatches12[who[k]] = idx[k];
            if (mbCheckOrientation) rotHist[rotationBin(pKF1->mvKeysUn[who[k]].angle - pKF2->mvKeysUn[idx[k]].angle, factor)].push_back(who[k]);
        }
        if (mbCheckOrientation) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++)
                if (i != ind1 && i != ind2 && i != ind3)
                    for (size_t j = 0; j < rotHist[i].size(); j++) { vMatches12[rotHist[i][j]] = -1; nmatches--; }
        }
        vMatchedPairs.reserve(nmatches);
        for (size_t i = 0, iend = vMatches12.size(); i < iend; i++)
            if (vMatches12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)vMatches12[i]));
        return nmatches;
    }

    // src/ORBmatcher.cc:1123-1351: the map points of each key frame projected into the other with the Sim3 (reference's cv::Mat algebra),
    // each matched on the device to the best descriptor of its window at levels [l-1, l] when <= TH_HIGH; kept where both directions agree.
    template <class KeyFrame>
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12,
                     const float th) {
        const float &fx = pKF1->fx, &fy = pKF1->fy, &cx = pKF1->cx, &cy = pKF1->cy;
        cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation();
        cv::Mat R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
        cv::Mat sR12 = s12 * R12;
        cv::Mat sR21 = (1.0 / s12) * R12.t();
        cv::Mat t21 = -sR21 * t12;
        const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
        const int N1 = (int)vpMapPoints1.size();
        const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
        const int N2 = (int)vpMapPoints2.size();
        std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
        for (int i = 0; i < N1; i++) {
            MapPoint* pMP = vpMatches12[i];
            if (pMP) {
                vbAlreadyMatched1[i] = true;
                int idx2 = pMP->GetIndexInKeyFrame(pKF2);
                if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
            }
        }
        // one direction: the map points of `from` (world -> its camera with Raw / taw, then sRba / tba into the other camera `into`)
        auto direction = [&](const std::vector<MapPoint*>& vpFrom, const std::vector<bool>& already, const cv::Mat& Raw, const cv::Mat& taw, const cv::Mat& sRba,
                             const cv::Mat& tba, KeyFrame* into, std::vector<int>& vnMatch) -> bool {
            std::vector<hvo_proj_query> q;
            std::vector<uint8_t> qdesc;
            std::vector<int> who;
            for (int i = 0; i < (int)vpFrom.size(); i++) {
                MapPoint* pMP = vpFrom[i];
                if (!pMP || already[i]) continue;
                if (pMP->isBad()) continue;
                cv::Mat p3Dw = pMP->GetWorldPos();
                cv::Mat p3Dca = Raw * p3Dw + taw;
                cv::Mat p3Dcb = sRba * p3Dca + tba;
                if (p3Dcb.template at<float>(2) < 0.0) continue;
                const float invz = 1.0 / p3Dcb.template at<float>(2);
                const float x = p3Dcb.template at<float>(0) * invz;
                const float y = p3Dcb.template at<float>(1) * invz;
                const float u = fx * x + cx;
                const float v = fy * y + cy;
                if (!into->IsInImage(u, v)) continue;
                const float maxDistance = pMP->GetMaxDistanceInvariance();
                const float minDistance = pMP->GetMinDistanceInvariance();
                const float dist3D = cv::norm(p3Dcb);
                if (dist3D < minDistance || dist3D > maxDistance) continue;
                const int nPredictedLevel = pMP->PredictScale(dist3D, into);
                hvo_proj_query e;
                e.u = u; e.v = v; e.r = th * into->mvScaleFactors[nPredictedLevel];
                e.min_level = nPredictedLevel - 1; e.max_level = nPredictedLevel; e.ur = -1.f; e.claims = 0; e.reserved = 0;
                q.push_back(e);
                appendDescriptor(qdesc, pMP->GetDescriptor());
                who.push_back(i);
            }
            if (q.empty()) return true;
            if (!setKeyFrame(into, false)) return false;
            std::vector<int32_t> idx(q.size(), -1);
            int n = 0;
            if (!ok(hvo_proj_search(h_, q.data(), qdesc.data(), (int)q.size(), nullptr, 1, TH_HIGH, mfNNratio, idx.data(), nullptr, &n))) return false;
            for (size_t k = 0; k < q.size(); ++k) vnMatch[who[k]] = idx[k];
            return true;
        };
        std::vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1);
        if (!direction(vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, vnMatch1)) return 0;
        if (!direction(vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12, pKF1, vnMatch2)) return 0;
        int nFound = 0;
        for (int i1 = 0; i1 < N1; i1++) {
            int idx2 = vnMatch1[i1];
            if (idx2 >= 0) {
                int idx1 = vnMatch2[idx2];
                if (idx1 == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
            }
        }
        return nFound;
    }

    // src/ORBmatcher.cc:838-994 (local mapping).  Projection tests = the reference's cv::Mat algebra; the window search with the chi-square
    // reprojection gate runs on the device (mode 2: queries are independent); Replace / AddObservation / AddMapPoint are applied here in
    // the reference's order, re-testing isBad() / IsInKeyFrame() at a point's turn because an earlier Replace can change them.
    template <class KeyFrame>
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0) {
        cv::Mat Rcw = pKF->GetRotation();
        cv::Mat tcw = pKF->GetTranslation();
        const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy, &bf = pKF->mbf;
        cv::Mat Ow = pKF->GetCameraCenter();
        std::vector<hvo_proj_query> q;
        std::vector<uint8_t> qdesc;
        std::vector<int> who;
        const int nMPs = (int)vpMapPoints.size();
        for (int i = 0; i < nMPs; i++) {
            MapPoint* pMP = vpMapPoints[i];
            if (!pMP) continue;
            if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
            cv::Mat p3Dw = pMP->GetWorldPos();
            cv::Mat p3Dc = Rcw * p3Dw + tcw;
            if (p3Dc.template at<float>(2) < 0.0f) continue;
            const float invz = 1 / p3Dc.template at<float>(2);
            const float x = p3Dc.template at<float>(0) * invz;
            const float y = p3Dc.template at<float>(1) * invz;
            const float u = fx * x + cx;
            const float v = fy * y + cy;
            if (!pKF->IsInImage(u, v)) continue;
            const float ur = u - bf * invz;
            const float maxDistance = pMP->GetMaxDistanceInvariance();
            const float minDistance = pMP->GetMinDistanceInvariance();
            cv::Mat PO = p3Dw - Ow;
            const float dist3D = cv::norm(PO);
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            cv::Mat Pn = pMP->GetNormal();
            if (PO.dot(Pn) < 0.5 * dist3D) continue;
            int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
            hvo_proj_query e;
            e.u = u; e.v = v; e.r = th * pKF->mvScaleFactors[nPredictedLevel];
            e.min_level = nPredictedLevel - 1; e.max_level = nPredictedLevel; e.ur = ur; e.claims = 0; e.reserved = 0;
            q.push_back(e);
            appendDescriptor(qdesc, pMP->GetDescriptor());
            who.push_back(i);
        }
        if (q.empty()) return 0;
        if (!setKeyFrame(pKF, true)) return 0;
        if (!ok(hvo_proj_set_level_sigma(h_, pKF->mvInvLevelSigma2.data(), (int)pKF->mvInvLevelSigma2.size()))) return 0;
        std::vector<int32_t> idx(q.size(), -1);
        int n = 0;
        if (!ok(hvo_proj_search(h_, q.data(), qdesc.data(), (int)q.size(), nullptr, 2, TH_LOW, mfNNratio, idx.data(), nullptr, &n))) return 0;
        int nFused = 0;
        for (size_t k = 0; k < q.size(); ++k) {
            MapPoint* pMP = vpMapPoints[who[k]];
            if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;   // an earlier Replace of this call may have changed it (:856)
            if (idx[k] < 0) continue;
            const int bestIdx = idx[k];
            MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) {
                    if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                    else pMPinKF->Replace(pMP);
                }
            } else {
                pMP->AddObservation(pKF, bestIdx);
                pKF->AddMapPoint(pMP, bestIdx);
            }
            nFused++;
        }
        return nFused;
    }

    // src/ORBmatcher.cc:996-1121 (loop closing): Sim3 projection, best descriptor of the window at levels [l-1, l] when <= TH_LOW, nothing is
    // claimed; vpReplacePoint / AddObservation / AddMapPoint applied here in the reference's order.
    template <class KeyFrame>
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
        Sim3Camera<KeyFrame> cam(pKF, Scw);
        const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
        std::vector<hvo_proj_query> q;
        std::vector<uint8_t> qdesc;
        std::vector<int> who;
        const int nPoints = (int)vpPoints.size();
        for (int iMP = 0; iMP < nPoints; iMP++) {
            MapPoint* pMP = vpPoints[iMP];
            if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
            hvo_proj_query e;
            if (!cam.project(pMP, th, true, e)) continue;
            q.push_back(e);
            appendDescriptor(qdesc, pMP->GetDescriptor());
            who.push_back(iMP);
        }
        if (q.empty()) return 0;
        if (!setKeyFrame(pKF, false)) return 0;
        std::vector<int32_t> idx(q.size(), -1);
        int n = 0;
        if (!ok(hvo_proj_search(h_, q.data(), qdesc.data(), (int)q.size(), nullptr, 1, TH_LOW, mfNNratio, idx.data(), nullptr, &n))) return 0;
        int nFused = 0;
        for (size_t k = 0; k < q.size(); ++k) {
            if (idx[k] < 0) continue;
            MapPoint* pMP = vpPoints[who[k]];
            const int bestIdx = idx[k];
            MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) vpReplacePoint[who[k]] = pMPinKF;
            } else {
                pMP->AddObservation(pKF, bestIdx);
                pKF->AddMapPoint(pMP, bestIdx);
            }
            nFused++;
        }
        return nFused;
    }

protected:
    // The Sim3 camera of the two loop-closing searches (:303-308, :1004-1009) and their common projection tests (:321-360, :1026-1063),
    // written with the reference's cv::Mat expressions so that they evaluate identically.
    template <class KeyFrame>
    struct Sim3Camera {
        KeyFrame* pKF;
        cv::Mat Rcw, tcw, Ow;
        Sim3Camera(KeyFrame* kf, const cv::Mat& Scw) : pKF(kf) {
            cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
            const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
            Rcw = sRcw / scw;
            tcw = Scw.rowRange(0, 3).col(3) / scw;
            Ow = -Rcw.t() * tcw;
        }
        bool project(MapPoint* pMP, float th, bool viewing_angle, hvo_proj_query& e) const {
            const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy;
            cv::Mat p3Dw = pMP->GetWorldPos();
            cv::Mat p3Dc = Rcw * p3Dw + tcw;
            if (p3Dc.template at<float>(2) < 0.0) return false;
            const float invz = 1 / p3Dc.template at<float>(2);
            const float x = p3Dc.template at<float>(0) * invz;
            const float y = p3Dc.template at<float>(1) * invz;
            const float u = fx * x + cx;
            const float v = fy * y + cy;
            if (!pKF->IsInImage(u, v)) return false;
            const float maxDistance = pMP->GetMaxDistanceInvariance();
            const float minDistance = pMP->GetMinDistanceInvariance();
            cv::Mat PO = p3Dw - Ow;
            const float dist = cv::norm(PO);
            if (dist < minDistance || dist > maxDistance) return false;
            if (viewing_angle) {
                cv::Mat Pn = pMP->GetNormal();
                if (PO.dot(Pn) < 0.5 * dist) return false;
            }
            int nPredictedLevel = pMP->PredictScale(dist, pKF);
            e.u = u; e.v = v; e.r = th * pKF->mvScaleFactors[nPredictedLevel];
            e.min_level = nPredictedLevel - 1; e.max_level = nPredictedLevel; e.ur = -1.f; e.claims = 0; e.reserved = 0;
            return true;
        }
    };

    float RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }   // :134-140

    static int rotationBin(float rot, float factor) {
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        return bin;
    }
    // :1630-1671
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3) {
        int max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < L; i++) {
            const int s = (int)histo[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
    }

    // with_uright = false for the searches that do not test the right coordinate (relocalisation, the Sim3 searches)
    bool setFrame(const Frame& F, bool with_uright = true) {
        const int n = (int)F.mvKeysUn.size();
        std::vector<uint8_t> desc;
        desc.reserve((size_t)n * 32);
        for (int i = 0; i < n; ++i) appendDescriptor(desc, F.mDescriptors.row(i));
        static_assert(sizeof(cv::KeyPoint) == sizeof(hvo_keypoint), "cv::KeyPoint must be the 28-byte POD the C ABI expects");
        return h_ && ok(hvo_proj_set_frame(h_, reinterpret_cast<const hvo_keypoint*>(F.mvKeysUn.data()),
                                           !with_uright || F.mvuRight.empty() ? nullptr : F.mvuRight.data(), desc.data(), n, Frame::mnMinX, Frame::mnMinY,
                                           Frame::mnMaxX, Frame::mnMaxY));
    }
    // A key frame's grid is the copy of its Frame's (src/KeyFrame.cc:44-60: cells assigned with the static float bounds of Frame), but
    // KeyFrame::GetFeaturesInArea (:627-666) locates a window from the key frame's INTEGER mnMinX / mnMinY.
    template <class KeyFrame>
    bool setKeyFrame(KeyFrame* pKF, bool with_uright) {
        const int n = (int)pKF->mvKeysUn.size();
        std::vector<uint8_t> desc;
        desc.reserve((size_t)n * 32);
        for (int i = 0; i < n; ++i) appendDescriptor(desc, pKF->mDescriptors.row(i));
        return h_ && ok(hvo_proj_set_frame(h_, reinterpret_cast<const hvo_keypoint*>(pKF->mvKeysUn.data()),
                                           !with_uright || pKF->mvuRight.empty() ? nullptr : pKF->mvuRight.data(), desc.data(), n, Frame::mnMinX,
                                           Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY)) &&
               ok(hvo_proj_set_window_origin(h_, (float)pKF->mnMinX, (float)pKF->mnMinY));
    }
    static std::vector<uint8_t> claimedOf(const Frame& F) {   // keypoints holding a map point with observations at call time (:88-90)
        std::vector<uint8_t> c(F.mvpMapPoints.size() ? F.mvpMapPoints.size() : 1, 0);
        for (size_t i = 0; i < F.mvpMapPoints.size(); ++i) c[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;
        return c;
    }
    static void appendDescriptor(std::vector<uint8_t>& out, const cv::Mat& row) {
        const uint8_t* p = row.template ptr<uint8_t>();
        out.insert(out.end(), p, p + 32);
    }
    static bool ok(int st) { if (st != HVO_OK) std::fprintf(stderr, "ORBmatcher: %s\n", hvo_last_error()); return st == HVO_OK; }

    float mfNNratio;
    bool mbCheckOrientation;
    hvo_proj* h_;
};

}  // namespace hvo_shim

#endif
