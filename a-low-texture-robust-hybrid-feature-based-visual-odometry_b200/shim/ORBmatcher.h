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
//   DescriptorDistance                                             src/ORBmatcher.cc:1676-1692
// (Fuse / SearchForTriangulation / SearchBySim3 need the KeyFrame pose / observation bookkeeping: their device loops are reached
//  through shim/WindowedMatcherGPU.h's PointWindowMatcher, see INTEGRATION.md.)
#ifndef HVO_SHIM_ORBMATCHER_H
#define HVO_SHIM_ORBMATCHER_H

#include <cmath>
#include <cstdio>
#include <cstring>
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

protected:
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

    bool setFrame(const Frame& F) {
        const int n = (int)F.mvKeysUn.size();
        std::vector<uint8_t> desc;
        desc.reserve((size_t)n * 32);
        for (int i = 0; i < n; ++i) appendDescriptor(desc, F.mDescriptors.row(i));
        static_assert(sizeof(cv::KeyPoint) == sizeof(hvo_keypoint), "cv::KeyPoint must be the 28-byte POD the C ABI expects");
        return h_ && ok(hvo_proj_set_frame(h_, reinterpret_cast<const hvo_keypoint*>(F.mvKeysUn.data()), F.mvuRight.empty() ? nullptr : F.mvuRight.data(),
                                           desc.data(), n, Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY));
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
