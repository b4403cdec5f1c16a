// GPU bodies for the windowed (projection) searches of ORB_SLAM2::ORBmatcher / LSDmatcher and for the normal extraction of
// ORB_SLAM2::Manhattan.  The projection of map points / map lines (isInFrustum, pose algebra) stays in the reference's own
// code; the maintainer replaces the candidate loops with one call per search:
//
//   ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)        src/ORBmatcher.cc:45-132     PointWindowMatcher::search(mode 0)
//   ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, mono)           src/ORBmatcher.cc:1353-1497  PointWindowMatcher::search(mode 1)
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches)               src/ORBmatcher.cc:162-293    PointWindowMatcher::searchCandidates
//   ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th)                   src/ORBmatcher.cc:838-990    PointWindowMatcher::setLevelSigma + search(mode 2)
//   ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, pairs, bOnlyStereo)     src/ORBmatcher.cc:668-836    PointWindowMatcher::searchTriangulation
//   LSDmatcher::SearchByProjection(Frame&, const vector<MapLine*>&, bool, th)   src/LSDmatcher.cpp:709-801   LineWindowMatcher::search(mode 0)
//   LSDmatcher::SearchByProjection(CurrentFrame, LastFrame, th)                 src/LSDmatcher.cpp:561-664   LineWindowMatcher::search(mode 1)
//   Manhattan::computeNormalsLPVO(depth, K, pt_normals, depth_normals)          src/Manhattan.cpp:237-393    LpvoNormals::compute
//
// Every search returns, per query in the reference's visiting order, the index of the frame feature it is assigned to (or
// -1): applying `F.mvpMapPoints[idx[i]] = pMP_i` (resp. mvpMapLines) for i ascending reproduces the reference's final state.
#ifndef HVO_SHIM_WINDOWED_MATCHER_GPU_H
#define HVO_SHIM_WINDOWED_MATCHER_GPU_H

#include <cstdio>
#include <vector>

#include "hvo_capi.h"

namespace hvo_shim {

class PointWindowMatcher {
public:
    explicit PointWindowMatcher(int device = 0) : h_(nullptr) {
        if (hvo_proj_create(device, &h_) != HVO_OK) std::fprintf(stderr, "PointWindowMatcher: %s\n", hvo_last_error());
    }
    ~PointWindowMatcher() { hvo_proj_destroy(h_); }
    PointWindowMatcher(const PointWindowMatcher&) = delete;
    PointWindowMatcher& operator=(const PointWindowMatcher&) = delete;

    // F.mvKeysUn (cv::KeyPoint is the 28-byte hvo_keypoint), F.mvuRight (or null), F.mDescriptors rows, Frame::mnMinX.. bounds
    bool setFrame(const hvo_keypoint* keysUn, const float* uRight, const uint8_t* desc, int n, float minX, float minY, float maxX, float maxY) {
        return ok(hvo_proj_set_frame(h_, keysUn, uRight, desc, n, minX, minY, maxX, maxY));
    }
    // claimed[i] != 0: F.mvpMapPoints[i] holds a map point with Observations() > 0 at call time.  mode 0: TH_HIGH + level ratio
    // (ORBmatcher.cc:121-124); mode 1: best <= th only (ORBmatcher.cc:1440).  Returns nmatches (before any rotation-histogram culling).
    int search(const std::vector<hvo_proj_query>& q, const std::vector<uint8_t>& qdesc, const uint8_t* claimed, int mode, int th, float nnratio,
               std::vector<int32_t>& idx) {
        idx.assign(q.size(), -1);
        int n = 0;
        if (!q.empty() && !ok(hvo_proj_search(h_, q.data(), qdesc.data(), (int)q.size(), claimed, mode, th, nnratio, idx.data(), nullptr, &n))) return 0;
        return n;
    }
    // SearchByBoW: queries = key-frame features with a good map point in (node, list) order; cand lists = frame features of the node
    int searchCandidates(const std::vector<uint8_t>& qdesc, const std::vector<uint8_t>& tdesc, const std::vector<int32_t>& offsets,
                         const std::vector<int32_t>& cand, int th, float nnratio, std::vector<int32_t>& idx) {
        const int nq = (int)offsets.size() - 1;
        idx.assign(nq > 0 ? nq : 0, -1);
        int n = 0;
        if (nq > 0 && !ok(hvo_proj_search_candidates(h_, qdesc.data(), nq, tdesc.data(), (int)(tdesc.size() / 32), offsets.data(), cand.data(), th,
                                                     nnratio, idx.data(), nullptr, &n))) return 0;
        return n;
    }

    // Fuse: mvInvLevelSigma2 of the key frame searched in, then search(mode 2, TH_LOW)
    bool setLevelSigma(const std::vector<float>& invLevelSigma2) { return ok(hvo_proj_set_level_sigma(h_, invLevelSigma2.data(), (int)invLevelSigma2.size())); }
    // SearchForTriangulation: queries = pKF1 features without a map point in (node, list) order; tflags bit 0 = pKF2 feature holds a
    // map point, bit 1 = it has a right coordinate; F12 row-major; (ex, ey) epipole in pKF2
    int searchTriangulation(const std::vector<uint8_t>& qdesc, const std::vector<hvo_keypoint>& qkeys, const std::vector<uint8_t>& qstereo,
                            const std::vector<uint8_t>& tdesc, const std::vector<hvo_keypoint>& tkeys, const std::vector<uint8_t>& tflags,
                            const std::vector<int32_t>& offsets, const std::vector<int32_t>& cand, const float F12[9], float ex, float ey,
                            const std::vector<float>& scaleFactors, const std::vector<float>& levelSigma2, bool onlyStereo, int thLow,
                            std::vector<int32_t>& idx) {
        const int nq = (int)qkeys.size();
        idx.assign(nq, -1);
        int n = 0;
        if (nq > 0 && !ok(hvo_proj_search_triangulation(h_, qdesc.data(), qkeys.data(), qstereo.data(), nq, tdesc.data(), tkeys.data(), tflags.data(),
                                                        (int)tkeys.size(), offsets.data(), cand.data(), F12, ex, ey, scaleFactors.data(),
                                                        levelSigma2.data(), (int)scaleFactors.size(), onlyStereo ? 1 : 0, thLow, idx.data(), nullptr,
                                                        &n))) return 0;
        return n;
    }

private:
    static bool ok(int st) { if (st != HVO_OK) std::fprintf(stderr, "PointWindowMatcher: %s\n", hvo_last_error()); return st == HVO_OK; }
    hvo_proj* h_;
};

class LineWindowMatcher {
public:
    explicit LineWindowMatcher(int device = 0) : h_(nullptr) {
        if (hvo_lproj_create(device, &h_) != HVO_OK) std::fprintf(stderr, "LineWindowMatcher: %s\n", hvo_last_error());
    }
    ~LineWindowMatcher() { hvo_lproj_destroy(h_); }
    LineWindowMatcher(const LineWindowMatcher&) = delete;
    LineWindowMatcher& operator=(const LineWindowMatcher&) = delete;

    // F.mvKeylinesUn (KeyLine is the 68-byte hvo_keyline), F.mvKeyLineFunctions (3 doubles each), F.mLdesc rows,
    // F.mvLines3D as first.xyz, second.xyz (null when only the last-frame search is used)
    bool setFrame(const hvo_keyline* keylinesUn, const double* lineFunctions, const uint8_t* ldesc, const double* lines3D, int n, float minX,
                  float minY, float maxX, float maxY) {
        return ok(hvo_lproj_set_frame(h_, keylinesUn, lineFunctions, ldesc, lines3D, n, minX, minY, maxX, maxY));
    }
    int search(const std::vector<hvo_lproj_query>& q, const std::vector<uint8_t>& qdesc, const uint8_t* claimed, int mode, float nnratio,
               std::vector<int32_t>& idx) {
        idx.assign(q.size(), -1);
        int n = 0;
        if (!q.empty() && !ok(hvo_lproj_search(h_, q.data(), qdesc.data(), (int)q.size(), claimed, mode, nnratio, idx.data(), nullptr, &n))) return 0;
        return n;
    }

private:
    static bool ok(int st) { if (st != HVO_OK) std::fprintf(stderr, "LineWindowMatcher: %s\n", hvo_last_error()); return st == HVO_OK; }
    hvo_lproj* h_;
};

// Manhattan::computeNormalsLPVO on the raw 16-bit depth image (z = raw * factor, the float image the reference intends to read)
class LpvoNormals {
public:
    LpvoNormals(float fx, float fy, float cx, float cy, float depthFactor, int width, int height, int device = 0) : h_(nullptr), cap_(0) {
        hvo_plane_params p{fx, fy, cx, cy, depthFactor};
        if (hvo_lpvo_create(&p, width, height, 1, device, &h_) != HVO_OK) std::fprintf(stderr, "LpvoNormals: %s\n", hvo_last_error());
        cap_ = hvo_lpvo_capacity(h_);
    }
    ~LpvoNormals() { hvo_lpvo_destroy(h_); }
    LpvoNormals(const LpvoNormals&) = delete;
    LpvoNormals& operator=(const LpvoNormals&) = delete;

    // pt_normals as 3 doubles each, depth_normals, pixel (u, v) of every sample; returns the number of normals
    int compute(const uint16_t* depth16, std::vector<double>& ptNormals, std::vector<float>& depthNormals, std::vector<int32_t>& pixels) {
        ptNormals.resize((size_t)cap_ * 3); depthNormals.resize(cap_); pixels.resize((size_t)cap_ * 2);
        int32_t n = 0;
        if (!h_ || hvo_lpvo_compute_batch(h_, depth16, 1, ptNormals.data(), depthNormals.data(), pixels.data(), &n) != HVO_OK) {
            std::fprintf(stderr, "LpvoNormals: %s\n", hvo_last_error());
            n = 0;
        }
        ptNormals.resize((size_t)n * 3); depthNormals.resize(n); pixels.resize((size_t)n * 2);
        return n;
    }

private:
    hvo_lpvo* h_;
    int cap_;
};

}  // namespace hvo_shim

#endif
