// Drop-in ORB_SLAM2::LSDmatcher for the searches on the tracking path (reference include/LSDmatcher.h:20-75, src/LSDmatcher.cpp), as a
// class template over the reference's own Frame / MapLine types:
//
//     #include "shim/LSDmatcher.h"
//     namespace ORB_SLAM2 { typedef hvo_shim::LSDmatcherT<Frame, MapLine> LSDmatcher; }        // instead of include/LSDmatcher.h
//
// Same constructor, constants, method names, argument meaning and side effects (F.mvpMapLines written in place, nmatches returned).
// Members read: F.NL, mvKeylinesUn, mvKeyLineFunctions, mLdesc, mvLines3D, mvpMapLines, mvbLineOutlier, Frame::mnMinX..;
// pML->mbTrackInView, mTrackProjX1/Y1/X2/Y2, mnTrackScaleLevel, mTrackViewCos, isBad(), Observations(), GetDescriptor(),
// GetWorldVector(); CurrentFrame.isInFrustum(pML, 0.5) is the caller's (Frame's) own, as in the reference.
//
//   SearchByProjection(F, vpMapLines, eval_orient, th)     src/LSDmatcher.cpp:709-801
//   SearchByProjection(CurrentFrame, LastFrame, th)        src/LSDmatcher.cpp:561-664
//   matchNNR / match                                       src/LSDmatcher.cpp:803-863
//   FrameBFMatch (knn-2 + MAD gates)                       src/LSDmatcher.cpp:942-966, lineDescriptorMAD :1110-1135
//   FrameBFMatchNew (knn-2 + epipolar overlap)             src/LSDmatcher.cpp:968-1031
//   SearchByDescriptor(pKF, currentF, vpMapLineMatches)    src/LSDmatcher.cpp:522-559   (KeyFrame type deduced: mLineDescriptors, GetMapLineMatches())
//   SearchDouble(InitialFrame, CurrentFrame, LineMatches)  src/LSDmatcher.cpp:903-940   (FrameBFMatch both ways + cross-check)
//   SearchDouble(KF, CurrentFrame)                         src/LSDmatcher.cpp:865-901   (KF->NL, mLineDescriptors, GetMapLine())
//   SearchForTriangulation(pKF1, pKF2, vector<pair>&)      src/LSDmatcher.cpp:1155-1193 (FrameBFMatch both ways at TH_LOW, cross-check, no MapLine)
//   SearchForTriangulation(pKF1, pKF2, vector<int>&, dbl)  src/LSDmatcher.cpp:1195-1231 (TH_HIGH, optional cross-check)
//   DescriptorDistance                                     src/LSDmatcher.cpp:1137-1153
#ifndef HVO_SHIM_LSDMATCHER_H
#define HVO_SHIM_LSDMATCHER_H

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <utility>
#include <vector>

#include "hvo_capi.h"

namespace hvo_shim {

template <class Frame, class MapLine>
class LSDmatcherT {
public:
    static const int TH_HIGH = 80, TH_LOW = 50, HISTO_LENGTH = 30;   // src/LSDmatcher.cpp:12-14

    LSDmatcherT(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri), h_(nullptr), m_(nullptr) {
        if (hvo_lproj_create(device, &h_) != HVO_OK || hvo_matcher_create(device, &m_) != HVO_OK) std::fprintf(stderr, "LSDmatcher: %s\n", hvo_last_error());
    }
    ~LSDmatcherT() { hvo_lproj_destroy(h_); hvo_matcher_destroy(m_); }
    LSDmatcherT(const LSDmatcherT&) = delete;
    LSDmatcherT& operator=(const LSDmatcherT&) = delete;

    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return hvo_hamming_distance(a.template ptr<uint8_t>(), b.template ptr<uint8_t>()); }

    int SearchByProjection(Frame& F, const std::vector<MapLine*>& vpMapLines, const bool eval_orient, const float th = 3) {
        (void)eval_orient;   // unused by the reference's body as well
        if (!setFrame(F, true)) return 0;
        const bool bFactor = th != 1.0;
        std::vector<hvo_lproj_query> q;
        std::vector<uint8_t> qdesc;
        std::vector<size_t> who;
        for (size_t iML = 0; iML < vpMapLines.size(); iML++) {
            MapLine* pML = vpMapLines[iML];
            if (!pML->mbTrackInView || pML->isBad()) continue;
            float r = RadiusByViewingCos(pML->mTrackViewCos);
            if (bFactor) r *= th;
            hvo_lproj_query e;
            std::memset(&e, 0, sizeof(e));
            e.x1 = pML->mTrackProjX1; e.y1 = pML->mTrackProjY1; e.x2 = pML->mTrackProjX2; e.y2 = pML->mTrackProjY2;
            e.r = r; e.cos_th = 0.998f;
            const auto wv = pML->GetWorldVector();
            e.dir[0] = wv(0); e.dir[1] = wv(1); e.dir[2] = wv(2);
            e.claims = pML->Observations() > 0;
            q.push_back(e);
            appendDescriptor(qdesc, pML->GetDescriptor());
            who.push_back(iML);
        }
        return run(F, q, qdesc, 0, [&](size_t k) { return vpMapLines[who[k]]; });
    }

    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th) {
        if (!setFrame(CurrentFrame, false)) return 0;
        std::vector<hvo_lproj_query> q;
        std::vector<uint8_t> qdesc;
        std::vector<int> who;
        for (int i = 0; i < LastFrame.NL; i++) {
            MapLine* pML = LastFrame.mvpMapLines[i];
            if (!pML || LastFrame.mvbLineOutlier[i]) continue;
            if (!CurrentFrame.isInFrustum(pML, 0.5)) continue;
            hvo_lproj_query e;
            std::memset(&e, 0, sizeof(e));
            e.x1 = pML->mTrackProjX1; e.y1 = pML->mTrackProjY1; e.x2 = pML->mTrackProjX2; e.y2 = pML->mTrackProjY2;
            e.r = th; e.cos_th = 0.96f;
            const auto& kl = LastFrame.mvKeylinesUn[i];
            e.dir[0] = kl.ePointInOctaveX - kl.sPointInOctaveX;     // float difference widened to double, as Mat_<double> << (float - float) does
            e.dir[1] = kl.ePointInOctaveY - kl.sPointInOctaveY;
            e.length = kl.lineLength;
            e.claims = pML->Observations() > 0;
            q.push_back(e);
            appendDescriptor(qdesc, pML->GetDescriptor());
            who.push_back(i);
        }
        return run(CurrentFrame, q, qdesc, 1, [&](size_t k) { return LastFrame.mvpMapLines[who[k]]; });
    }

    int matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12) {
        matches_12.resize(desc1.rows, -1);
        std::vector<int32_t> idx, dist;
        if (!knn2(desc1, desc2, idx, dist)) return 0;
        int matches = 0;
        for (int i = 0; i < desc1.rows; ++i)
            if (idx[2 * i + 1] >= 0 && (float)dist[2 * i] < (float)dist[2 * i + 1] * nnr) { matches_12[i] = idx[2 * i]; matches++; }
        return matches;
    }
    int match(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12) { return matchNNR(desc1, desc2, nnr, matches_12); }

    void FrameBFMatch(cv::Mat ldesc1, cv::Mat ldesc2, std::vector<int>& LineMatches, float TH) {
        LineMatches = std::vector<int>(ldesc1.rows, -1);
        std::vector<int32_t> idx, dist;
        const int n = ldesc1.rows;
        if (n == 0 || !knn2(ldesc1, ldesc2, idx, dist)) return;
        // lineDescriptorMAD (:1110-1135): median absolute deviation of the NN12 gaps; the two sorts run on float distances
        std::vector<float> gap(n);
        for (int i = 0; i < n; ++i) gap[i] = (float)dist[2 * i + 1] - (float)dist[2 * i];
        std::vector<float> s(gap);
        std::sort(s.begin(), s.end(), [](float a, float b) { return a > b; });
        const double median = s[n / 2];
        std::vector<float> dev(n);
        for (int i = 0; i < n; ++i) dev[i] = (float)std::fabs((double)gap[i] - median);
        std::sort(dev.begin(), dev.end());
        const double nn12_dist_th = 1.4826 * (double)dev[n / 2] * 0.5;
        for (int i = 0; i < n; ++i) {
            const float d0 = (float)dist[2 * i], d1 = (float)dist[2 * i + 1];
            if ((double)gap[i] > nn12_dist_th && d0 < TH && d0 < mfNNratio * d1) LineMatches[i] = idx[2 * i];
        }
    }

    // src/LSDmatcher.cpp:522-559: knn-2 of the key frame's line descriptors in the frame's, ratio d0 / d1 < 1 / 1.5; a frame line receives the
    // MapLine of the key-frame line (later key-frame lines overwrite earlier ones, key-frame lines without a MapLine are skipped).
    // (The reference also calls currentF.lineDescriptorMAD here; its results are not used.)
    template <class KeyFrame>
    int SearchByDescriptor(KeyFrame* pKF, Frame& currentF, std::vector<MapLine*>& vpMapLineMatches) {
        const std::vector<MapLine*> vpMapLinesKF = pKF->GetMapLineMatches();
        vpMapLineMatches = std::vector<MapLine*>(currentF.NL, static_cast<MapLine*>(NULL));
        int nmatches = 0;
        std::vector<int32_t> idx, dist;
        const int n = pKF->mLineDescriptors.rows;
        if (n == 0 || currentF.mLdesc.rows < 2 || !knn2(pKF->mLineDescriptors, currentF.mLdesc, idx, dist)) return 0;
        const float minRatio = 1.0f / 1.5f;
        for (int qdx = 0; qdx < n; ++qdx) {
            const int tdx = idx[2 * qdx];
            const double dist_12 = (float)dist[2 * qdx] / (float)dist[2 * qdx + 1];
            if (dist_12 < minRatio) {
                MapLine* mapLine = vpMapLinesKF[qdx];
                if (mapLine) { vpMapLineMatches[tdx] = mapLine; nmatches++; }
            }
        }
        return nmatches;
    }

    // src/LSDmatcher.cpp:903-940 (the reference runs the two FrameBFMatch calls in two threads; here they are two device calls)
    int SearchDouble(Frame& InitialFrame, Frame& CurrentFrame, std::vector<int>& LineMatches) {
        LineMatches = std::vector<int>(InitialFrame.NL, -1);
        std::vector<int> tempMatches1, tempMatches2;
        cv::Mat ldesc1 = InitialFrame.mLdesc, ldesc2 = CurrentFrame.mLdesc;
        if (ldesc1.rows == 0 || ldesc2.rows == 0) return 0;
        int nmatches = 0;
        FrameBFMatch(ldesc1, ldesc2, tempMatches1, TH_LOW);
        FrameBFMatch(ldesc2, ldesc1, tempMatches2, TH_LOW);
        for (size_t i = 0; i < tempMatches1.size(); i++) {
            int j = tempMatches1[i];
            if (j >= 0) {
                if (tempMatches2[j] != (int)i) tempMatches1[i] = -1;
                else nmatches++;
            }
        }
        LineMatches = tempMatches1;
        return nmatches;
    }

    // src/LSDmatcher.cpp:865-901
    template <class KeyFrame>
    int SearchDouble(KeyFrame* KF, Frame& CurrentFrame) {
        std::vector<int> tempMatches1(KF->NL, -1), tempMatches2(CurrentFrame.NL, -1);
        cv::Mat ldesc1 = KF->mLineDescriptors, ldesc2 = CurrentFrame.mLdesc;
        if (ldesc1.rows == 0 || ldesc2.rows == 0) return 0;
        int nmatches = 0;
        FrameBFMatch(ldesc1, ldesc2, tempMatches1, TH_LOW);
        FrameBFMatch(ldesc2, ldesc1, tempMatches2, TH_LOW);
        for (size_t i = 0; i < tempMatches2.size(); i++) {
            int j = tempMatches2[i];
            if (j >= 0 && tempMatches1[j] == (int)i) {
                MapLine* pML = KF->GetMapLine(j);
                if (!pML) continue;
                CurrentFrame.mvpMapLines[i] = pML;
                nmatches++;
            }
        }
        return nmatches;
    }

    // src/LSDmatcher.cpp:1155-1193
    template <class KeyFrame>
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<std::pair<size_t, size_t> >& vMatchedPairs) {
        vMatchedPairs.clear();
        std::vector<int> tempMatches1, tempMatches2;
        cv::Mat ldesc1 = pKF1->mLineDescriptors, ldesc2 = pKF2->mLineDescriptors;
        if (ldesc1.rows == 0 || ldesc2.rows == 0) return 0;
        int nmatches = 0;
        FrameBFMatch(ldesc1, ldesc2, tempMatches1, TH_LOW);
        FrameBFMatch(ldesc2, ldesc1, tempMatches2, TH_LOW);
        for (size_t i = 0; i < tempMatches1.size(); i++) {
            int j = tempMatches1[i];
            if (j >= 0 && tempMatches2[j] == (int)i) {
                if (pKF1->GetMapLine(i) || pKF2->GetMapLine(j)) continue;
                vMatchedPairs.push_back(std::make_pair(i, (size_t)j));
                nmatches++;
            }
        }
        return nmatches;
    }

    // src/LSDmatcher.cpp:1195-1231
    template <class KeyFrame>
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<int>& vMatchedPairs, bool isDouble = false) {
        vMatchedPairs.clear();
        vMatchedPairs.resize(pKF1->NL, -1);
        std::vector<int> tempMatches1, tempMatches2;
        cv::Mat ldesc1 = pKF1->mLineDescriptors, ldesc2 = pKF2->mLineDescriptors;
        if (ldesc1.rows == 0 || ldesc2.rows == 0) return 0;
        int nmatches = 0;
        FrameBFMatch(ldesc1, ldesc2, tempMatches1, TH_HIGH);
        FrameBFMatch(ldesc2, ldesc1, tempMatches2, TH_HIGH);
        for (size_t i = 0; i < tempMatches1.size(); i++) {
            int j = tempMatches1[i];
            if (j >= 0) {
                if (isDouble && tempMatches2[j] != (int)i) continue;
                if (pKF1->GetMapLine(i) || pKF2->GetMapLine(j)) continue;
                vMatchedPairs[i] = j;
                nmatches++;
            }
        }
        return nmatches;
    }

    template <class KeyLine, class Vec3>
    void FrameBFMatchNew(cv::Mat ldesc1, cv::Mat ldesc2, std::vector<int>& LineMatches, std::vector<KeyLine> kls1, std::vector<KeyLine> kls2,
                         std::vector<Vec3> kls2func, cv::Mat F, float TH) {
        static_assert(sizeof(KeyLine) == sizeof(hvo_keyline), "KeyLine must be the 68-byte POD the C ABI expects");
        LineMatches = std::vector<int>(ldesc1.rows, -1);
        const int n1 = ldesc1.rows, n2 = ldesc2.rows;
        if (n1 == 0 || !m_) return;
        std::vector<uint8_t> d1, d2;
        for (int i = 0; i < n1; ++i) appendDescriptor(d1, ldesc1.row(i));
        for (int i = 0; i < n2; ++i) appendDescriptor(d2, ldesc2.row(i));
        std::vector<double> f2((size_t)n2 * 3);
        for (int i = 0; i < n2; ++i) for (int k = 0; k < 3; ++k) f2[3 * i + k] = kls2func[i](k);
        float Fm[9];
        for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) Fm[3 * i + k] = F.template at<float>(i, k);
        std::vector<int32_t> out(n1, -1);
        if (hvo_match_lines_epipolar(m_, d1.data(), reinterpret_cast<const hvo_keyline*>(kls1.data()), n1, d2.data(),
                                     reinterpret_cast<const hvo_keyline*>(kls2.data()), f2.data(), n2, Fm, TH, mfNNratio, out.data()) != HVO_OK) {
            std::fprintf(stderr, "LSDmatcher: %s\n", hvo_last_error());
            return;
        }
        for (int i = 0; i < n1; ++i) LineMatches[i] = out[i];
    }

protected:
    float RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 5.0f : 8.0f; }   // :1436-1442

    template <class Pick>
    int run(Frame& F, const std::vector<hvo_lproj_query>& q, const std::vector<uint8_t>& qdesc, int mode, Pick pick) {
        std::vector<int32_t> idx(q.size(), -1);
        int n = 0;
        std::vector<uint8_t> claimed(F.mvpMapLines.size() ? F.mvpMapLines.size() : 1, 0);
        for (size_t i = 0; i < F.mvpMapLines.size(); ++i) claimed[i] = F.mvpMapLines[i] && F.mvpMapLines[i]->Observations() > 0;
        if (!q.empty() && hvo_lproj_search(h_, q.data(), qdesc.data(), (int)q.size(), claimed.data(), mode, mfNNratio, idx.data(), nullptr, &n) != HVO_OK) {
            std::fprintf(stderr, "LSDmatcher: %s\n", hvo_last_error());
            return 0;
        }
        for (size_t k = 0; k < q.size(); ++k)
            if (idx[k] >= 0) F.mvpMapLines[idx[k]] = pick(k);
        return n;
    }
    bool setFrame(const Frame& F, bool need3d) {
        const int n = (int)F.mvKeylinesUn.size();
        std::vector<uint8_t> desc;
        std::vector<double> func((size_t)n * 3), l3d;
        for (int i = 0; i < n; ++i) {
            appendDescriptor(desc, F.mLdesc.row(i));
            for (int k = 0; k < 3; ++k) func[3 * i + k] = F.mvKeyLineFunctions[i](k);
        }
        if (need3d) {
            l3d.resize((size_t)n * 6);
            for (int i = 0; i < n; ++i)
                for (int k = 0; k < 3; ++k) { l3d[6 * i + k] = F.mvLines3D[i].first(k); l3d[6 * i + 3 + k] = F.mvLines3D[i].second(k); }
        }
        static_assert(sizeof(F.mvKeylinesUn[0]) == sizeof(hvo_keyline), "KeyLine must be the 68-byte POD the C ABI expects");
        const int st = h_ ? hvo_lproj_set_frame(h_, reinterpret_cast<const hvo_keyline*>(F.mvKeylinesUn.data()), func.data(), desc.data(),
                                                need3d ? l3d.data() : nullptr, n, Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY)
                          : HVO_ERR_STATE;
        if (st != HVO_OK) std::fprintf(stderr, "LSDmatcher: %s\n", hvo_last_error());
        return st == HVO_OK;
    }
    bool knn2(const cv::Mat& desc1, const cv::Mat& desc2, std::vector<int32_t>& idx, std::vector<int32_t>& dist) {
        const int nq = desc1.rows, nt = desc2.rows;
        idx.assign((size_t)std::max(nq, 1) * 2, -1); dist.assign((size_t)std::max(nq, 1) * 2, -1);
        if (!m_ || nq == 0 || nt == 0) return false;
        std::vector<uint8_t> q, t;
        for (int i = 0; i < nq; ++i) appendDescriptor(q, desc1.row(i));
        for (int i = 0; i < nt; ++i) appendDescriptor(t, desc2.row(i));
        if (hvo_match_knn2(m_, q.data(), nq, t.data(), nt, idx.data(), dist.data()) != HVO_OK) {
            std::fprintf(stderr, "LSDmatcher: %s\n", hvo_last_error());
            return false;
        }
        return true;
    }
    static void appendDescriptor(std::vector<uint8_t>& out, const cv::Mat& row) {
        const uint8_t* p = row.template ptr<uint8_t>();
        out.insert(out.end(), p, p + 32);
    }

    float mfNNratio;
    bool mbCheckOrientation;
    hvo_lproj* h_;
    hvo_matcher* m_;
};

}  // namespace hvo_shim

#endif
