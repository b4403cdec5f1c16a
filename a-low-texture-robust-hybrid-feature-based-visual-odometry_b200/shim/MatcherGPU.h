// GPU bodies for the brute-force parts of ORB_SLAM2::LSDmatcher / ORBmatcher (reference include/LSDmatcher.h:23-60,
// include/ORBmatcher.h:38-77).  The matcher classes keep their many geometric helpers on the host; the maintainer replaces
// the cv::BFMatcher::knnMatch call sites (src/LSDmatcher.cpp:811-812, :948-949, :529-534) with hvo_shim::knnMatch2 and keeps
// the ratio / MAD / cross-check logic around them untouched (it consumes the same (index, distance) pairs).
#ifndef HVO_SHIM_MATCHER_GPU_H
#define HVO_SHIM_MATCHER_GPU_H

#include <cstdio>
#include <cstring>
#include <vector>

#include <opencv2/core/core.hpp>

#include "hvo_capi.h"

namespace hvo_shim {

// int ORBmatcher::DescriptorDistance(const cv::Mat&, const cv::Mat&)  (src/ORBmatcher.cc:1676-1692); LSDmatcher has the same body
static inline int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return hvo_hamming_distance(a.ptr(0), b.ptr(0)); }

struct DMatchLite { int queryIdx, trainIdx; float distance; };  // cv::DMatch fields the reference reads

class BruteForceMatcher {
public:
    explicit BruteForceMatcher(int device = 0) : m_(nullptr) {
        if (hvo_matcher_create(device, &m_) != HVO_OK) std::fprintf(stderr, "BFMatcher: %s\n", hvo_last_error());
    }
    ~BruteForceMatcher() { hvo_matcher_destroy(m_); }
    BruteForceMatcher(const BruteForceMatcher&) = delete;
    BruteForceMatcher& operator=(const BruteForceMatcher&) = delete;

    // cv::BFMatcher(NORM_HAMMING, false).knnMatch(desc1, desc2, matches, 2): descriptors are N x 32 CV_8U rows
    void knnMatch2(const cv::Mat& desc1, const cv::Mat& desc2, std::vector<std::vector<DMatchLite>>& matches) {
        matches.clear();
        const int nq = desc1.rows, nt = desc2.rows;
        if (!m_ || nq == 0 || nt == 0) return;
        q_.resize((size_t)nq * 32);
        t_.resize((size_t)nt * 32);
        for (int i = 0; i < nq; ++i) std::memcpy(&q_[(size_t)i * 32], desc1.ptr(i), 32);
        for (int i = 0; i < nt; ++i) std::memcpy(&t_[(size_t)i * 32], desc2.ptr(i), 32);
        idx_.resize((size_t)nq * 2);
        dist_.resize((size_t)nq * 2);
        if (hvo_match_knn2(m_, q_.data(), nq, t_.data(), nt, idx_.data(), dist_.data()) != HVO_OK) {
            std::fprintf(stderr, "BFMatcher: %s\n", hvo_last_error());
            return;
        }
        matches.resize(nq);
        for (int i = 0; i < nq; ++i)
            for (int k = 0; k < 2; ++k)
                if (idx_[2 * i + k] >= 0) matches[i].push_back(DMatchLite{i, idx_[2 * i + k], (float)dist_[2 * i + k]});
    }

    // LSDmatcher::matchNNR (src/LSDmatcher.cpp:803-826): matches_12[i] = train index or -1; returns the number of matches
    int matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12) {
        std::vector<std::vector<DMatchLite>> m;
        knnMatch2(desc1, desc2, m);
        matches_12.assign(desc1.rows, -1);
        int matches = 0;
        for (int i = 0; i < (int)m.size(); ++i)
            if (m[i].size() == 2 && m[i][0].distance < m[i][1].distance * nnr) { matches_12[i] = m[i][0].trainIdx; ++matches; }
        return matches;
    }

private:
    hvo_matcher* m_;
    std::vector<uint8_t> q_, t_;
    std::vector<int32_t> idx_, dist_;
};

}  // namespace hvo_shim

#endif
