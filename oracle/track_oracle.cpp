// TEST INFRASTRUCTURE — CPU restatement of the tracking-time projection and of two matcher loops.  Not on the product path.
//
//   Frame::isInFrustum(MapPoint*, float)            reference src/Frame.cc:1371-1436
//   Frame::isInFrustum(MapLine*, float)             src/Frame.cc:1438-1499
//   MapPoint::PredictScale(dist, Frame*)            src/MapPoint.cc:400-415     (clamped to [0, nlevels))
//   MapLine::PredictScale(dist, logScaleFactor)     src/MapLine.cpp:549-558     (not clamped)
//   Get{Min,Max}DistanceInvariance                  src/MapPoint.cc:371-381, src/MapLine.cpp:537-547
//   ORBmatcher::SearchForInitialization             src/ORBmatcher.cc:412-497   (the loop; the rotation histogram :499-523 is
//                                                   restated in Python next to the mirror)
//   LSDmatcher::FrameBFMatchNew, mutualOverlap      src/LSDmatcher.cpp:968-1108
//
// The cv::Mat expressions are written out with OpenCV's evaluation rules for CV_32F (checked against cv2 4.13.0 where Python
// exposes the operation: gemm, norm, addWeighted): A * x (+ c) = float products summed in k order (the + c in double, which
// for two floats equals the float sum); cv::norm and Mat::dot accumulate in double, in order; Mat::cross in float;
// `m /= s` multiplies by (float)(1. / s).  Pinned by executing the reference's own functions: oracle/_ref/ref_match ops 4-7.
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace tracko {

struct Cam { float R[9], t[3], O[3], fx, fy, cx, cy, bf, min_x, min_y, max_x, max_y, log_scale_factor; int32_t n_levels; };  // hvo_frustum_cam
struct MapPoint { float pos[3], normal[3], min_distance, max_distance; };                                                    // hvo_map_point
struct TrackPoint { float u, v, ur; int32_t level; float view_cos; int32_t in_view; };                                       // hvo_track_point
struct MapLine { double pos[6], normal[3], dir[3]; float min_distance, max_distance; };                                      // hvo_map_line
struct TrackLine { float x1, y1, x2, y2; int32_t level; float view_cos; int32_t in_view; };                                  // hvo_track_line

static void to_camera(const Cam& c, const float* p, float* out) {   // mRcw * P + mtcw
    for (int i = 0; i < 3; ++i) {
        float s = 0.f;
        for (int k = 0; k < 3; ++k) s += c.R[3 * i + k] * p[k];
        out[i] = (float)((double)s + (double)c.t[i]);
    }
}
static double norm3(const float* v) {
    double s = 0;
    for (int i = 0; i < 3; ++i) s += (double)v[i] * (double)v[i];
    return std::sqrt(s);
}
static double dot3(const float* a, const float* b) {
    double s = 0;
    for (int i = 0; i < 3; ++i) s += (double)a[i] * (double)b[i];
    return s;
}

static TrackPoint frustum_point(const Cam& c, const MapPoint& m, float limit) {
    TrackPoint o{0.f, 0.f, 0.f, 0, 0.f, 0};
    float Pc[3];
    to_camera(c, m.pos, Pc);
    if (Pc[2] < 0.0f) return o;
    const float invz = 1.0f / Pc[2];
    const float u = c.fx * Pc[0] * invz + c.cx;
    const float v = c.fy * Pc[1] * invz + c.cy;
    if (u < c.min_x || u > c.max_x) return o;
    if (v < c.min_y || v > c.max_y) return o;
    const float maxDistance = 1.2f * m.max_distance, minDistance = 0.8f * m.min_distance;
    const float PO[3] = {m.pos[0] - c.O[0], m.pos[1] - c.O[1], m.pos[2] - c.O[2]};
    const float dist = (float)norm3(PO);
    if (dist < minDistance || dist > maxDistance) return o;
    const float viewCos = (float)(dot3(PO, m.normal) / dist);
    if (viewCos < limit) return o;
    const float ratio = m.max_distance / dist;
    int nScale = (int)std::ceil(std::log(ratio) / c.log_scale_factor);   // float log, float division (using namespace std; float argument)
    if (nScale < 0) nScale = 0;
    else if (nScale >= c.n_levels) nScale = c.n_levels - 1;
    o.u = u; o.v = v; o.ur = u - c.bf * invz; o.level = nScale; o.view_cos = viewCos; o.in_view = 1;
    return o;
}

static TrackLine frustum_line(const Cam& c, const MapLine& m, float limit) {
    TrackLine o{0.f, 0.f, 0.f, 0.f, 0, 0.f, 0};
    const float SP[3] = {(float)m.pos[0], (float)m.pos[1], (float)m.pos[2]}, EP[3] = {(float)m.pos[3], (float)m.pos[4], (float)m.pos[5]};
    float S[3], E[3];
    to_camera(c, SP, S);
    to_camera(c, EP, E);
    if (S[2] < 0.0f || E[2] < 0.0f) return o;
    const float invz1 = 1.0f / S[2];
    const float u1 = c.fx * S[0] * invz1 + c.cx, v1 = c.fy * S[1] * invz1 + c.cy;
    if (u1 < c.min_x || u1 > c.max_x) return o;
    if (v1 < c.min_y || v1 > c.max_y) return o;
    const float invz2 = 1.0f / E[2];
    const float u2 = c.fx * E[0] * invz2 + c.cx, v2 = c.fy * E[1] * invz2 + c.cy;
    if (u2 < c.min_x || u2 > c.max_x) return o;
    if (v2 < c.min_y || v2 > c.max_y) return o;
    const float maxDistance = 1.2f * m.max_distance, minDistance = 0.8f * m.min_distance;
    float OM[3];
    for (int i = 0; i < 3; ++i) OM[i] = (SP[i] * 0.5f + EP[i] * 0.5f) - c.O[i];   // addWeighted(SP, 0.5, EP, 0.5) - mOw
    const float dist = (float)norm3(OM);
    if (dist < minDistance || dist > maxDistance) return o;
    const float pn[3] = {(float)m.normal[0], (float)m.normal[1], (float)m.normal[2]};
    const float viewCos = (float)(dot3(OM, pn) / dist);
    if (viewCos < limit) return o;
    const float ratio = m.max_distance / dist;
    o.x1 = u1; o.y1 = v1; o.x2 = u2; o.y2 = v2; o.level = (int)std::ceil(std::log(ratio) / c.log_scale_factor); o.view_cos = viewCos; o.in_view = 1;
    return o;
}

// ---- the frame grid of F2 (as in proj_oracle.cpp; restated here so the file stands alone) ----
struct KeyPoint { float x, y, size, angle, response; int octave, class_id; };
static const int kCols = 64, kRows = 48;
struct Grid {
    float min_x, min_y, inv_w, inv_h;
    std::vector<int> cell[kCols][kRows];
    void build(const KeyPoint* k, int n, const float* b) {
        min_x = b[0]; min_y = b[1];
        inv_w = (float)kCols / (b[2] - b[0]);
        inv_h = (float)kRows / (b[3] - b[1]);
        for (int i = 0; i < n; ++i) {
            const int px = (int)std::round((k[i].x - min_x) * inv_w), py = (int)std::round((k[i].y - min_y) * inv_h);
            if (px < 0 || px >= kCols || py < 0 || py >= kRows) continue;
            cell[px][py].push_back(i);
        }
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
                    if (check) {
                        if (k[id].octave < minLevel) continue;
                        if (maxLevel >= 0 && k[id].octave > maxLevel) continue;
                    }
                    if (std::fabs(k[id].x - x) < r && std::fabs(k[id].y - y) < r) out.push_back(id);
                }
    }
};
static int hamming(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

struct KeyLine {  // cv::line_descriptor::KeyLine, 68 bytes
    float angle; int class_id, octave; float pt_x, pt_y, response, size, sx, sy, ex, ey, sox, soy, eox, eoy, length; int npix;
};
static float dist3(const float* a, const float* b) {
    const float d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    return (float)norm3(d);
}
static float mutual_overlap(const float pt[4][3]) {
    float max_dist = 0.0f;
    int outer1 = 0, outer2 = 3, inner1, inner2;
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 4; ++j) {
            const float d = dist3(pt[i], pt[j]);
            if (d > max_dist) { max_dist = d; outer1 = i; outer2 = j; }
        }
    if (max_dist < 1.0f) return 0.0f;
    if (outer1 == 0) {
        if (outer2 == 1) { inner1 = 2; inner2 = 3; }
        else if (outer2 == 2) { inner1 = 1; inner2 = 3; }
        else { inner1 = 1; inner2 = 2; }
    } else if (outer1 == 1) {
        inner1 = 0;
        inner2 = outer2 == 2 ? 3 : 2;
    } else { inner1 = 0; inner2 = 1; }
    const float d[3] = {pt[inner1][0] - pt[inner2][0], pt[inner1][1] - pt[inner2][1], pt[inner1][2] - pt[inner2][2]};
    return (float)(norm3(d) / max_dist);
}

}  // namespace tracko

using namespace tracko;

extern "C" {

void orc_frustum_points(const void* cam, const void* pts, int n, float limit, void* out) {
    const Cam& c = *(const Cam*)cam;
    for (int i = 0; i < n; ++i) ((TrackPoint*)out)[i] = frustum_point(c, ((const MapPoint*)pts)[i], limit);
}
void orc_frustum_lines(const void* cam, const void* lines, int n, float limit, void* out) {
    const Cam& c = *(const Cam*)cam;
    for (int i = 0; i < n; ++i) ((TrackLine*)out)[i] = frustum_line(c, ((const MapLine*)lines)[i], limit);
}

// ORBmatcher::SearchForInitialization, the loop (:424-496).  accepted12[i1] = the keypoint i1 took when visited (rotHist entries).
int orc_search_initialization(const void* keys2, const uint8_t* desc2, int n2, const float* bounds, const float* prev_xy, const int32_t* octave1,
                              const uint8_t* desc1, int n1, int window, int th_low, float nnratio, int32_t* matches12, int32_t* accepted12) {
    const KeyPoint* k2 = (const KeyPoint*)keys2;
    Grid* g = new Grid();
    g->build(k2, n2, bounds);
    int nmatches = 0;
    std::vector<int> vMatchedDistance(n2, INT_MAX), vnMatches21(n2, -1), cand;
    for (int i = 0; i < n1; ++i) { matches12[i] = -1; accepted12[i] = -1; }
    for (int i1 = 0; i1 < n1; ++i1) {
        if (octave1[i1] > 0) continue;
        g->area(k2, prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)window, 0, 0, cand);
        if (cand.empty()) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : cand) {
            const int dist = hamming(desc1 + 32 * (size_t)i1, desc2 + 32 * (size_t)i2);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= th_low && bestDist < (float)bestDist2 * nnratio) {
            if (vnMatches21[bestIdx2] >= 0) { matches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
            matches12[i1] = bestIdx2;
            accepted12[i1] = bestIdx2;
            vnMatches21[bestIdx2] = i1;
            vMatchedDistance[bestIdx2] = bestDist;
            nmatches++;
        }
    }
    delete g;
    return nmatches;
}

void orc_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx2, int32_t* dist2);

// LSDmatcher::FrameBFMatchNew (:968-1031): only j = 0 is visited (j < size() - 1 with k = 2)
void orc_lines_epipolar(const uint8_t* d1, const void* kls1, int n1, const uint8_t* d2, const void* kls2, const double* func2, int n2, const float* F,
                        float TH, float nnratio, int32_t* line_matches) {
    for (int i = 0; i < n1; ++i) line_matches[i] = -1;
    if (n1 <= 0 || n2 < 2) return;
    const KeyLine *k1 = (const KeyLine*)kls1, *k2 = (const KeyLine*)kls2;
    std::vector<int32_t> idx((size_t)n1 * 2), dist((size_t)n1 * 2);
    orc_knn2(d1, n1, d2, n2, idx.data(), dist.data());
    for (int q = 0; q < n1; ++q) {
        const int t = idx[2 * q];
        const float p1[3] = {k1[q].sx, k1[q].sy, 1.0f}, p2[3] = {k1[q].ex, k1[q].ey, 1.0f};
        float e1[3], e2[3], pt[4][3];
        for (int i = 0; i < 3; ++i) {
            float s = 0.f, u = 0.f;
            for (int k = 0; k < 3; ++k) { s += F[3 * i + k] * p1[k]; u += F[3 * i + k] * p2[k]; }
            e1[i] = s; e2[i] = u;
        }
        const float l2[3] = {(float)func2[3 * t], (float)func2[3 * t + 1], (float)func2[3 * t + 2]};
        const float* es[2] = {e1, e2};
        for (int s = 0; s < 2; ++s) {   // l2.cross(epi)
            const float* b = es[s];
            pt[s][0] = l2[1] * b[2] - l2[2] * b[1];
            pt[s][1] = l2[2] * b[0] - l2[0] * b[2];
            pt[s][2] = l2[0] * b[1] - l2[1] * b[0];
        }
        if (!(std::fabs(pt[0][2]) > 1e-12 && std::fabs(pt[1][2]) > 1e-12)) continue;
        for (int s = 0; s < 2; ++s) {
            const float f = (float)(1. / (double)pt[s][2]);
            for (int i = 0; i < 3; ++i) pt[s][i] = pt[s][i] * f;
        }
        pt[2][0] = k2[t].sx; pt[2][1] = k2[t].sy; pt[2][2] = 1.0f;
        pt[3][0] = k2[t].ex; pt[3][1] = k2[t].ey; pt[3][2] = 1.0f;
        const float score = mutual_overlap(pt);
        const float d0 = (float)dist[2 * q], dn = (float)dist[2 * q + 1];
        if (d0 < TH) {
            if (score > 0.8 && d0 < nnratio * dn) line_matches[q] = t;
        }
    }
}

}  // extern "C"
