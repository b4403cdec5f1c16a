// TEST INFRASTRUCTURE — CPU restatement of the windowed (projection) matchers and of the frame grid they search.
// Not on the product path.
//
//   Frame::AssignFeaturesToGrid / PosInGrid        reference src/Frame.cc:832-847, 1680-1690   (64 x 48 cells)
//   Frame::GetFeaturesInArea                       src/Frame.cc:1502-1555  (cell range by floor / ceil, ix outer, iy inner,
//                                                  level filter, |dx| < r && |dy| < r; this order is the candidate order)
//   ORBmatcher::SearchByProjection(F, MapPoints)   src/ORBmatcher.cc:45-132   -> mode 0: best + second with their levels, accept
//                                                  best <= TH and not (bestLevel == bestLevel2 && best > ratio * second)
//   ORBmatcher::SearchByProjection(Cur, Last, ...) src/ORBmatcher.cc:1353-1497 (and the reloc variant :1499-1628)
//                                                  -> mode 1: best only, accept best <= TH
// Both walk their queries in order and skip keypoints already holding a map point with observations (:88-90,
// :1425-1428), i.e. claimed at call time or by an earlier query of the same call: a greedy, order-dependent assignment.
// The projection itself (isInFrustum / pose) stays with the caller: a query is the projected position, the search radius,
// the level range, the predicted right coordinate and whether the assigned map point counts as an observation.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace projo {

struct KeyPoint { float x, y, size, angle, response; int octave, class_id; };  // cv::KeyPoint, 28 bytes
struct Query { float u, v, r; int min_level, max_level; float ur; int claims; int pad; };  // hvo_proj_query, 32 bytes

static const int kCols = 64, kRows = 48;

// KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:627-666) looks windows up with the key frame's integer mnMinX / mnMinY in cells its Frame
// assigned with the float bounds; orc_window_origin(1, x, y) selects that origin for the grids built afterwards on this thread.
static thread_local bool g_win_on = false;
static thread_local float g_win_x = 0.f, g_win_y = 0.f;

struct Grid {
    float min_x, min_y, inv_w, inv_h;
    std::vector<int> cell[kCols][kRows];
    void build(const KeyPoint* k, int n, float mnx, float mny, float mxx, float mxy) {
        inv_w = (float)kCols / (mxx - mnx);
        inv_h = (float)kRows / (mxy - mny);
        for (int i = 0; i < n; ++i) {
            const int px = (int)std::round((k[i].x - mnx) * inv_w), py = (int)std::round((k[i].y - mny) * inv_h);
            if (px < 0 || px >= kCols || py < 0 || py >= kRows) continue;
            cell[px][py].push_back(i);
        }
        min_x = g_win_on ? g_win_x : mnx; min_y = g_win_on ? g_win_y : mny;   // origin of the window lookups below
    }
    void area(const KeyPoint* k, float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
        out.clear();
        const int x0 = std::max(0, (int)std::floor((x - min_x - r) * inv_w));
        if (x0 >= kCols) return;
        const int x1 = std::min(kCols - 1, (int)std::ceil((x - min_x + r) * inv_w));
        if (x1 < 0) return;
        const int y0 = std::max(0, (int)std::floor((y - min_y - r) * inv_h));
        if (y0 >= kRows) return;
        const int y1 = std::min(kRows - 1, (int)std::ceil((y - min_y + r) * inv_h));
        if (y1 < 0) return;
        const bool check = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = x0; ix <= x1; ++ix)
            for (int iy = y0; iy <= y1; ++iy)
                for (int id : cell[ix][iy]) {
                    const KeyPoint& kp = k[id];
                    if (check) {
                        if (kp.octave < minLevel) continue;
                        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                    }
                    const float dx = kp.x - x, dy = kp.y - y;
                    if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(id);
                }
    }
};

static int hamming(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i += 4) {
        uint32_t x, y;
        std::memcpy(&x, a + i, 4);
        std::memcpy(&y, b + i, 4);
        d += __builtin_popcount(x ^ y);
    }
    return d;
}

}  // namespace projo

extern "C" {

// grid inspection: cell_count [64*48] (index ix * 48 + iy), cell_items in the same order (n entries at most)
void orc_window_origin(int on, float x, float y) { projo::g_win_on = on != 0; projo::g_win_x = x; projo::g_win_y = y; }
void orc_grid_build(const void* keys, int n, float min_x, float min_y, float max_x, float max_y, int32_t* cell_count, int32_t* cell_items) {
    projo::Grid g;
    g.build((const projo::KeyPoint*)keys, n, min_x, min_y, max_x, max_y);
    int o = 0;
    for (int ix = 0; ix < projo::kCols; ++ix)
        for (int iy = 0; iy < projo::kRows; ++iy) {
            cell_count[ix * projo::kRows + iy] = (int)g.cell[ix][iy].size();
            for (int id : g.cell[ix][iy]) cell_items[o++] = id;
        }
}

int orc_features_in_area(const void* keys, int n, float min_x, float min_y, float max_x, float max_y, float x, float y, float r, int minLevel,
                         int maxLevel, int32_t* out, int cap) {
    projo::Grid g;
    g.build((const projo::KeyPoint*)keys, n, min_x, min_y, max_x, max_y);
    std::vector<int> v;
    g.area((const projo::KeyPoint*)keys, x, y, r, minLevel, maxLevel, v);
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}

// mode 0: SearchByProjection(F, vpMapPoints, th)  (best / second / levels / ratio);  mode 1: best only.
// claimed [n] (or null): keypoint holds a map point with observations at call time.  match_idx / match_dist [nq].
int orc_search_projection(const void* keys, const float* uright, const uint8_t* desc, int n, float min_x, float min_y, float max_x,
                          float max_y, const void* queries, const uint8_t* qdesc, int nq, const uint8_t* claimed_in, int mode,
                          int th_dist, float nnratio, int32_t* match_idx, int32_t* match_dist) {
    using namespace projo;
    const KeyPoint* K = (const KeyPoint*)keys;
    const Query* Q = (const Query*)queries;
    Grid g;
    g.build(K, n, min_x, min_y, max_x, max_y);
    std::vector<uint8_t> claimed(n > 0 ? n : 1, 0);
    if (claimed_in) std::memcpy(claimed.data(), claimed_in, n);
    std::vector<int> cand;
    int nmatches = 0;
    for (int k = 0; k < nq; ++k) {
        match_idx[k] = -1; match_dist[k] = 256;
        const Query& q = Q[k];
        g.area(K, q.u, q.v, q.r, q.min_level, q.max_level, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : cand) {
            if (claimed[idx]) continue;
            if (uright && uright[idx] > 0) {
                const float er = std::fabs(q.ur - uright[idx]);
                if (er > q.r) continue;
            }
            const int dist = hamming(qdesc + 32 * (size_t)k, desc + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = K[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = K[idx].octave; bestDist2 = dist;
            }
        }
        if (bestDist <= th_dist) {
            if (mode == 0 && bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
            match_idx[k] = bestIdx; match_dist[k] = bestDist;
            if (q.claims) claimed[bestIdx] = 1;
            ++nmatches;
        }
    }
    return nmatches;
}

// the candidate loop of ORBmatcher::Fuse(pKF, vpMapPoints, th) (src/ORBmatcher.cc:896-990): window at levels [min, max], chi-square gate
// on the reprojection error (5.99 mono / 7.8 with a right coordinate), best only, best <= th_dist; queries are independent
int orc_search_fuse(const void* keys, const float* uright, const uint8_t* desc, int n, float min_x, float min_y, float max_x, float max_y,
                    const void* queries, const uint8_t* qdesc, int nq, const float* inv_sigma2, int th_dist, int32_t* match_idx,
                    int32_t* match_dist) {
    using namespace projo;
    const KeyPoint* K = (const KeyPoint*)keys;
    const Query* Q = (const Query*)queries;
    Grid g;
    g.build(K, n, min_x, min_y, max_x, max_y);
    std::vector<int> cand;
    int nm = 0;
    for (int k = 0; k < nq; ++k) {
        match_idx[k] = -1; match_dist[k] = 256;
        const Query& q = Q[k];
        g.area(K, q.u, q.v, q.r, -1, -1, cand);   // pKF->GetFeaturesInArea(u, v, radius): every level
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            const KeyPoint& kp = K[idx];
            if (kp.octave < q.min_level || kp.octave > q.max_level) continue;
            const float ur = uright ? uright[idx] : -1.f;
            if (ur >= 0) {
                const float ex = q.u - kp.x, ey = q.v - kp.y, er = q.ur - ur;
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * inv_sigma2[kp.octave] > 7.8) continue;
            } else {
                const float ex = q.u - kp.x, ey = q.v - kp.y;
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv_sigma2[kp.octave] > 5.99) continue;
            }
            const int dist = hamming(qdesc + 32 * (size_t)k, desc + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= th_dist) { match_idx[k] = bestIdx; match_dist[k] = bestDist; ++nm; }
    }
    return nm;
}

// the candidate loop of ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:668-836) with CheckDistEpipolarLine (:143-160)
int orc_search_triangulation(const uint8_t* qdesc, const void* qkeys, const uint8_t* qstereo, int nq, const uint8_t* tdesc, const void* tkeys,
                             const uint8_t* tflags, const int32_t* off, const int32_t* cand, const float* F12, float ex, float ey,
                             const float* scale_factors, const float* level_sigma2, int only_stereo, int th_low, int32_t* match_idx,
                             int32_t* match_dist) {
    using namespace projo;
    const KeyPoint* K1 = (const KeyPoint*)qkeys;
    const KeyPoint* K2 = (const KeyPoint*)tkeys;
    int nm = 0;
    for (int k = 0; k < nq; ++k) {
        match_idx[k] = -1; match_dist[k] = 256;
        const bool bStereo1 = qstereo[k] != 0;
        if (only_stereo && !bStereo1) continue;
        const KeyPoint& kp1 = K1[k];
        int bestDist = th_low, bestIdx2 = -1;
        for (int c = off[k]; c < off[k + 1]; ++c) {
            const int idx2 = cand[c];
            if (tflags[idx2] & 1) continue;
            const bool bStereo2 = (tflags[idx2] & 2) != 0;
            if (only_stereo && !bStereo2) continue;
            const int dist = hamming(qdesc + 32 * (size_t)k, tdesc + 32 * (size_t)idx2);
            if (dist > th_low || dist > bestDist) continue;
            const KeyPoint& kp2 = K2[idx2];
            if (!bStereo1 && !bStereo2) {
                const float distex = ex - kp2.x, distey = ey - kp2.y;
                if (distex * distex + distey * distey < 100 * scale_factors[kp2.octave]) continue;
            }
            const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
            const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
            const float cc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
            const float num = a * kp2.x + b * kp2.y + cc;
            const float den = a * a + b * b;
            if (den == 0) continue;
            const float dsqr = num * num / den;
            if (dsqr < 3.84 * level_sigma2[kp2.octave]) { bestIdx2 = idx2; bestDist = dist; }
        }
        if (bestIdx2 >= 0) { match_idx[k] = bestIdx2; match_dist[k] = bestDist; ++nm; }
    }
    return nm;
}

// generic candidate lists (e.g. the per-vocabulary-node buckets of SearchByBoW, src/ORBmatcher.cc:162-293): for query i the
// train indices cand[off[i] .. off[i+1]) in the caller's order; best4[i] = {idx0, dist0, idx1, dist1}, strict '<' updates.
void orc_match_candidates(const uint8_t* q, int nq, const uint8_t* t, const int32_t* off, const int32_t* cand, int32_t* best4) {
    for (int i = 0; i < nq; ++i) {
        int d0 = 256, i0 = -1, d1 = 256, i1 = -1;
        for (int c = off[i]; c < off[i + 1]; ++c) {
            const int d = projo::hamming(q + 32 * (size_t)i, t + 32 * (size_t)cand[c]);
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = cand[c]; }
            else if (d < d1) { d1 = d; i1 = cand[c]; }
        }
        best4[4 * i] = i0; best4[4 * i + 1] = d0; best4[4 * i + 2] = i1; best4[4 * i + 3] = d1;
    }
}

// the greedy loop of ORBmatcher::SearchByBoW (src/ORBmatcher.cc:197-251) over the same kind of candidate lists: a train row
// assigned to an earlier query is skipped (:216-217); accept best <= th_dist && (float)best < nnratio * (float)second (:235-239)
int orc_search_candidates(const uint8_t* q, int nq, const uint8_t* t, int nt, const int32_t* off, const int32_t* cand, int th_dist, float nnratio,
                          int32_t* match_idx, int32_t* match_dist) {
    std::vector<char> taken(nt > 0 ? nt : 1, 0);
    int nm = 0;
    for (int i = 0; i < nq; ++i) {
        int d0 = 256, i0 = -1, d1 = 256;
        for (int c = off[i]; c < off[i + 1]; ++c) {
            if (taken[cand[c]]) continue;
            const int d = projo::hamming(q + 32 * (size_t)i, t + 32 * (size_t)cand[c]);
            if (d < d0) { d1 = d0; d0 = d; i0 = cand[c]; }
            else if (d < d1) d1 = d;
        }
        match_idx[i] = -1; match_dist[i] = 256;
        if (i0 >= 0 && d0 <= th_dist && (float)d0 < nnratio * (float)d1) { match_idx[i] = i0; match_dist[i] = d0; taken[i0] = 1; ++nm; }
    }
    return nm;
}

}  // extern "C"
