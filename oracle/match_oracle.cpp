// TEST INFRASTRUCTURE — CPU oracle for descriptor matching, never on the product path.
//
// Restates (paths relative to /root/reference):
//   ORBmatcher::DescriptorDistance                 src/ORBmatcher.cc:1676-1692  (= LSDmatcher.cpp:1137-1153)
//   cv::BFMatcher(NORM_HAMMING).knnMatch(k=2)      un-vendored OpenCV; pinned by tests/golden/prims_cv2.npz (cv2 4.13.0)
//   LSDmatcher::matchNNR                           src/LSDmatcher.cpp:803-826
//   LSDmatcher::FrameBFMatch + lineDescriptorMAD   src/LSDmatcher.cpp:942-966, 1110-1135
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

static int swar_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        std::memcpy(&x, a + 4 * i, 4);
        std::memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

static void knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx2, int32_t* dist2) {
    for (int i = 0; i < nq; ++i) {
        int d0 = 257, i0 = -1, d1 = 257, i1 = -1;
        for (int j = 0; j < nt; ++j) {  // stable: ties keep the lower train index
            const int d = swar_distance(q + 32 * i, t + 32 * j);
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
            else if (d < d1) { d1 = d; i1 = j; }
        }
        idx2[2 * i] = i0; idx2[2 * i + 1] = i1;
        dist2[2 * i] = i0 >= 0 ? d0 : -1; dist2[2 * i + 1] = i1 >= 0 ? d1 : -1;
    }
}

extern "C" {

int orc_hamming(const uint8_t* a, const uint8_t* b) { return swar_distance(a, b); }

void orc_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx2, int32_t* dist2) { knn2(q, nq, t, nt, idx2, dist2); }

// LSDmatcher::matchNNR: matches12[i] = train index or -1; returns the number of matches.  Requires nt >= 2
// (the reference indexes matches_[idx][1] unconditionally).
int orc_match_nnr(const uint8_t* q, int nq, const uint8_t* t, int nt, float nnr, int32_t* matches12) {
    std::vector<int32_t> idx(2 * (size_t)nq), dist(2 * (size_t)nq);
    knn2(q, nq, t, nt, idx.data(), dist.data());
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        matches12[i] = -1;
        if ((float)dist[2 * i] < (float)dist[2 * i + 1] * nnr) { matches12[i] = idx[2 * i]; ++n; }
    }
    return n;
}

// LSDmatcher::FrameBFMatch with lineDescriptorMAD (nn12 threshold = 0.5 * 1.4826 * MAD of the NN-gap)
void orc_frame_bf_match(const uint8_t* q, int nq, const uint8_t* t, int nt, float nnratio, float TH, int32_t* line_matches) {
    std::vector<int32_t> idx(2 * (size_t)nq), dist(2 * (size_t)nq);
    knn2(q, nq, t, nt, idx.data(), dist.data());
    for (int i = 0; i < nq; ++i) line_matches[i] = -1;
    if (nq == 0) return;
    // MAD of the gap d1 - d0 (the reference sorts descending by gap, takes element n/2, then |gap - median|)
    std::vector<float> gap(nq);
    for (int i = 0; i < nq; ++i) gap[i] = (float)dist[2 * i + 1] - (float)dist[2 * i];
    std::vector<float> s = gap;
    std::sort(s.begin(), s.end(), [](float a, float b) { return a > b; });
    const double med = s[nq / 2];
    std::vector<float> dev(nq);
    for (int i = 0; i < nq; ++i) dev[i] = std::fabs((float)(gap[i] - med));
    std::sort(dev.begin(), dev.end());
    double nn12_th = 1.4826 * dev[nq / 2];
    nn12_th *= 0.5;
    for (int i = 0; i < nq; ++i) {
        const double d12 = (double)((float)dist[2 * i + 1] - (float)dist[2 * i]);
        if (d12 > nn12_th && (float)dist[2 * i] < TH && (float)dist[2 * i] < nnratio * (float)dist[2 * i + 1]) line_matches[i] = idx[2 * i];
    }
}

// MapPoint / MapLine::ComputeDistinctiveDescriptors (src/MapPoint.cc:240-300, src/MapLine.cpp:331-400) for groups of descriptors
void orc_distinctive(const uint8_t* desc, const int32_t* off, int ngroups, int32_t* best_idx, int32_t* best_median) {
    for (int g = 0; g < ngroups; ++g) {
        const int b = off[g], N = off[g + 1] - b;
        best_idx[g] = -1; best_median[g] = -1;
        if (N <= 0) continue;
        int BestMedian = INT_MAX, BestIdx = 0;
        std::vector<int> v(N);
        for (int i = 0; i < N; ++i) {
            for (int j = 0; j < N; ++j) {
                int d = 0;
                for (int k = 0; k < 32; ++k) d += __builtin_popcount(desc[32 * (size_t)(b + i) + k] ^ desc[32 * (size_t)(b + j) + k]);
                v[j] = d;
            }
            std::sort(v.begin(), v.end());
            const int median = v[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best_idx[g] = BestIdx; best_median[g] = BestMedian;
    }
}

}  // extern "C"
