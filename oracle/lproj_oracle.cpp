// TEST INFRASTRUCTURE — CPU restatement of the windowed (projection) line matchers and of the line grid they search.
// Not on the product path.
//
//   ORB_SLAM2::LineIterator (Bresenham on doubles)      reference src/lineIterator.cpp:35-79
//   Frame::AssignFeaturesToGridForLine                  src/Frame.cc:849-872    (64 x 48 cells; endpoints * gridInv, no mnMin)
//   Frame::GetFeaturesInAreaForLine                     src/Frame.cc:1557-1631  (3 sample points, |cos| >= TH, point-line
//                                                       distance < r; first acceptance fixes the position in the list)
//   LSDmatcher::SearchByProjection(F, MapLines, ...)    src/LSDmatcher.cpp:709-801  -> mode 0: 3-D direction gate (cos 15 deg),
//                                                       best + second with their octaves, accept best <= 95 unless
//                                                       (bestLevel == bestLevel2 && best > nnratio * second)
//   LSDmatcher::SearchByProjection(Cur, Last, th)       src/LSDmatcher.cpp:561-664  -> mode 1: 2-D direction gate (cos 10 deg),
//                                                       length ratio >= 0.75, best only, accept best <= 95
// Both walk their queries in order and skip lines already holding a map line with observations (:755-757, :614-616): a
// greedy, order-dependent assignment.  Projection / isInFrustum stay with the caller.  No reference execution is possible
// here (LSDmatcher needs OpenCV, Eigen, the whole Frame): PARITY UNPINNED by execution.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace lprojo {

// cv::line_descriptor::KeyLine, 68 bytes (descriptor_custom.hpp:105-144)
struct KeyLine {
    float angle; int class_id, octave; float pt_x, pt_y, response, size, startPointX, startPointY, endPointX, endPointY, sPointInOctaveX,
        sPointInOctaveY, ePointInOctaveX, ePointInOctaveY, lineLength; int numOfPixels;
};
static_assert(sizeof(KeyLine) == 68, "KeyLine layout");

// hvo_lproj_query, 64 bytes
struct Query {
    float x1, y1, x2, y2, r, cos_th;  // GetFeaturesInAreaForLine(x1, y1, x2, y2, r, ., ., TH)
    double dir[3];                    // mode 0: world direction of the map line; mode 1: last keyline's InOctave end - start (x, y, -)
    float length;                     // mode 1: lineLength of the last frame's keyline
    int claims;                       // != 0: the assigned line counts as taken for later queries
    int pad[2];
};
static_assert(sizeof(Query) == 64, "query layout");

static const int kCols = 64, kRows = 48;

struct LineIt {  // src/lineIterator.cpp
    bool steep;
    double x1, y1, x2, y2, dx, dy, error;
    int maxX, ystep, y, x;
    LineIt(double x1_, double y1_, double x2_, double y2_) : steep(std::abs(y2_ - y1_) > std::abs(x2_ - x1_)), x1(x1_), y1(y1_), x2(x2_), y2(y2_) {
        if (steep) { std::swap(x1, y1); std::swap(x2, y2); }
        if (x1 > x2) { std::swap(x1, x2); std::swap(y1, y2); }
        dx = x2 - x1;
        dy = std::abs(y2 - y1);
        error = dx / 2.0;
        ystep = (y1 < y2) ? 1 : -1;
        x = static_cast<int>(x1);
        y = static_cast<int>(y1);
        maxX = static_cast<int>(x2);
    }
    bool next(int& px, int& py) {
        if (x > maxX) return false;
        if (steep) { px = y; py = x; } else { px = x; py = y; }
        error -= dy;
        if (error < 0) { y += ystep; error += dx; }
        x++;
        return true;
    }
};

struct Grid {
    float min_x, min_y, inv_w, inv_h;
    std::vector<int> cell[kCols][kRows];
    void build(const KeyLine* kl, int n, float mnx, float mny, float mxx, float mxy) {
        min_x = mnx; min_y = mny;
        inv_w = (float)kCols / (mxx - mnx);
        inv_h = (float)kRows / (mxy - mny);
        for (int i = 0; i < n; ++i) {
            LineIt it(kl[i].startPointX * inv_w, kl[i].startPointY * inv_h, kl[i].endPointX * inv_w, kl[i].endPointY * inv_h);
            int px, py;
            while (it.next(px, py))
                if (px >= 0 && px < kCols && py >= 0 && py < kRows) cell[px][py].push_back(i);
        }
    }
    void area(const KeyLine* kl, const double* func3, float x1, float y1, float x2, float y2, float r, float TH, std::vector<int>& out) const {
        out.clear();
        std::vector<char> in_set;
        float x[3] = {x1, (float)((x1 + x2) / 2.0), x2};
        float y[3] = {y1, (float)((y1 + y2) / 2.0), y2};
        float d1x = x1 - x2, d1y = y1 - y2;
        const float n1 = std::sqrt(d1x * d1x + d1y * d1y);
        d1x /= n1; d1y /= n1;
        for (int i = 0; i < 3; ++i) {
            const int cx0 = std::max(0, (int)std::floor((x[i] - min_x - r) * inv_w));
            if (cx0 >= kCols) continue;
            const int cx1 = std::min(kCols - 1, (int)std::ceil((x[i] - min_x + r) * inv_w));
            if (cx1 < 0) continue;
            const int cy0 = std::max(0, (int)std::floor((y[i] - min_y - r) * inv_h));
            if (cy0 >= kRows) continue;
            const int cy1 = std::min(kRows - 1, (int)std::ceil((y[i] - min_y + r) * inv_h));
            if (cy1 < 0) continue;
            for (int ix = cx0; ix <= cx1; ++ix)
                for (int iy = cy0; iy <= cy1; ++iy)
                    for (int id : cell[ix][iy]) {
                        if ((int)in_set.size() > id && in_set[id]) continue;
                        const KeyLine& k = kl[id];
                        float d2x = k.startPointX - k.endPointX, d2y = k.startPointY - k.endPointY;
                        const float n2 = std::sqrt(d2x * d2x + d2y * d2y);
                        d2x /= n2; d2y /= n2;
                        const float cs = std::fabs(d1x * d2x + d1y * d2y);
                        if (cs < TH) continue;
                        const double* L = func3 + 3 * (size_t)id;
                        const float dist = (float)(L[0] * x[i] + L[1] * y[i] + L[2]);
                        if (std::fabs(dist) < r) {
                            out.push_back(id);
                            if ((int)in_set.size() <= id) in_set.resize(id + 1, 0);
                            in_set[id] = 1;
                        }
                    }
        }
    }
};

static inline int hamming(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i += 8) {
        uint64_t x, y;
        std::memcpy(&x, a + i, 8); std::memcpy(&y, b + i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

}  // namespace lprojo

using namespace lprojo;

extern "C" {

// cell_count [64*48] (cell = ix * 48 + iy), cell_items = cells concatenated (caller sizes it from a first call with cell_items = null)
int orc_line_grid_build(const void* keylines, int n, float min_x, float min_y, float max_x, float max_y, int32_t* cell_count, int32_t* cell_items) {
    Grid g;
    g.build((const KeyLine*)keylines, n, min_x, min_y, max_x, max_y);
    int m = 0;
    for (int ix = 0; ix < kCols; ++ix)
        for (int iy = 0; iy < kRows; ++iy) {
            cell_count[ix * kRows + iy] = (int)g.cell[ix][iy].size();
            for (int id : g.cell[ix][iy]) { if (cell_items) cell_items[m] = id; ++m; }
        }
    return m;
}

int orc_line_features_in_area(const void* keylines, const double* func3, int n, float min_x, float min_y, float max_x, float max_y, float x1, float y1,
                              float x2, float y2, float r, float TH, int32_t* out, int cap) {
    Grid g;
    g.build((const KeyLine*)keylines, n, min_x, min_y, max_x, max_y);
    std::vector<int> v;
    g.area((const KeyLine*)keylines, func3, x1, y1, x2, y2, r, TH, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}

// lines3d [n][6] (mvLines3D first.xyz, second.xyz; mode 0 only).  claimed [n] or null.  match_idx / match_dist [nq].
int orc_line_search_projection(const void* keylines, const double* func3, const uint8_t* desc, const double* lines3d, int n, float min_x, float min_y,
                               float max_x, float max_y, const void* queries, const uint8_t* qdesc, int nq, const uint8_t* claimed, int mode,
                               float nnratio, int32_t* match_idx, int32_t* match_dist) {
    const KeyLine* kl = (const KeyLine*)keylines;
    const Query* qs = (const Query*)queries;
    Grid g;
    g.build(kl, n, min_x, min_y, max_x, max_y);
    std::vector<char> taken(n, 0);
    if (claimed) for (int i = 0; i < n; ++i) taken[i] = claimed[i] != 0;
    const double th_normal = std::cos(15.0 / 180.0 * M_PI), cos_th_angle = std::cos(10.0 / 180.0 * M_PI);
    std::vector<int> cand;
    int nm = 0;
    for (int k = 0; k < nq; ++k) {
        const Query& q = qs[k];
        match_idx[k] = -1; match_dist[k] = 256;
        g.area(kl, func3, q.x1, q.y1, q.x2, q.y2, q.r, q.cos_th, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int id : cand) {
            if (taken[id]) continue;
            if (mode == 0) {
                const double* P = lines3d + 6 * (size_t)id;
                const double w[3] = {P[0] - P[3], P[1] - P[4], P[2] - P[5]};
                const float dot = (float)(w[0] * q.dir[0] + (w[1] * q.dir[1] + w[2] * q.dir[2]));  // Eigen's unrolled 3-term reduction: x0 + (x1 + x2)
                const float mag_f = (float)std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
                const float mag_ml = (float)std::sqrt(q.dir[0] * q.dir[0] + q.dir[1] * q.dir[1] + q.dir[2] * q.dir[2]);
                const float angle = std::fabs(dot / (mag_f * mag_ml));
                if ((double)angle < th_normal) continue;
                const int dist = hamming(qdesc + 32 * (size_t)k, desc + 32 * (size_t)id);
                if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kl[id].octave; bestIdx = id; }
                else if (dist < bestDist2) { bestLevel2 = kl[id].octave; bestDist2 = dist; }
            } else {
                const double cx = (double)(kl[id].ePointInOctaveX - kl[id].sPointInOctaveX), cy = (double)(kl[id].ePointInOctaveY - kl[id].sPointInOctaveY);
                const double dotp = cx * q.dir[0] + cy * q.dir[1];
                const double magA = std::sqrt(cx * cx + cy * cy), magB = std::sqrt(q.dir[0] * q.dir[0] + q.dir[1] * q.dir[1]);
                const double angle = std::fabs(dotp / (magA * magB));
                if (angle < cos_th_angle) continue;
                const int dist = hamming(qdesc + 32 * (size_t)k, desc + 32 * (size_t)id);
                const float mx = std::max(q.length, kl[id].lineLength), mn = std::min(q.length, kl[id].lineLength);
                if (mn / mx < 0.75) continue;
                if (dist < bestDist) { bestDist = dist; bestIdx = id; }
            }
        }
        if (bestDist <= 95) {
            if (mode == 0 && bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
            match_idx[k] = bestIdx; match_dist[k] = bestDist;
            if (q.claims) taken[bestIdx] = 1; else taken[bestIdx] = 0;
            ++nm;
        }
    }
    return nm;
}

}  // extern "C"
