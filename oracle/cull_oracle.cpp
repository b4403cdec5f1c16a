// TEST INFRASTRUCTURE — CPU restatement of Frame::cullingLine (reference src/Frame.cc:952-1116) with its helpers
// PointLineDistance (:1117-1126), TwoLineAngle (:1127-1140) and MergeTwoLines (:1141-1203).  Not on the product path.
//
// What the reference does after LINEextractor::operator() on every frame (Frame::ExtractLSD, src/Frame.cc:895-947):
//   1. pair test, i ascending, j > i, both not yet tagged: midpoint-to-line distance (one of the two < dis), |cos| of the
//      angle between the two normalised line functions > cos(angle deg), and, where the x (y) extents do not overlap,
//      a gap of at most endpoint_dis between the inner endpoints.  Line j joins group i; i and j are tagged.  Quirks
//      kept: the second midpoint is (start + end) / 2 + start (:975); TwoLineAngle divides by the third component
//      twice (:1131-1135).
//   2. every group is folded into one segment with MergeTwoLines (float endpoints, double arithmetic, atan/sin/cos),
//      ungrouped untagged lines are kept, in index order.
//   3. KeyLines are rebuilt from the new segments (:1061-1086; numOfPixels = cv::LineIterator(...).count on the
//      cvRound'ed endpoints, after cv::clipLine), sorted by response (descending, std::sort: ties in libstdc++'s order; the
//      oracle keeps creation order, like LineExtractor.cpp:353), class_id renumbered.
//   4. LBD descriptors are computed again on the new KeyLines (:1094-1096) and the line functions rebuilt (:1097-1108).
// cv::clipLine is un-vendored OpenCV (imgproc/src/drawing.cpp); its restatement below is pinned to cv2.clipLine in
// tests/test_cull.py.  KeyLine::angle: see the note at its assignment (libm's atan2f is not pinned by the reference).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "cvprims.hpp"

namespace cullo {

struct KeyLine {  // cv::line_descriptor::KeyLine, 68 bytes (descriptor_custom.hpp:105-144)
    float angle;
    int class_id, octave;
    float pt_x, pt_y, response, size;
    float startPointX, startPointY, endPointX, endPointY;
    float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength;
    int numOfPixels;
};
static_assert(sizeof(KeyLine) == 68, "KeyLine layout");

static double point_line_distance(const double line[4], float px, float py) {  // Frame.cc:1117-1126
    const double x0 = (double)px, y0 = (double)py;
    const double x1 = line[0], y1 = line[1], x2 = line[2], y2 = line[3];
    return std::fabs((y2 - y1) * x0 + (x1 - x2) * y0 + ((x2 * y1) - (x1 * y2))) /
           std::sqrt((y2 - y1) * (y2 - y1) + (x1 - x2) * (x1 - x2));
}

static double two_line_angle(const double l1[3], const double l2[3]) {  // Frame.cc:1127-1140
    double v1[3] = {l1[0], l1[1], l1[2]}, v2[3] = {l2[0], l2[1], l2[2]};
    v1[0] /= v1[2]; v1[1] /= v1[2];
    v2[0] /= v2[2]; v2[1] /= v2[2];
    const double a0 = v1[0] / v1[2], a1 = v1[1] / v1[2], b0 = v2[0] / v2[2], b1 = v2[1] / v2[2];
    const double a = a0 * b0 + a1 * b1;  // cv::Mat::dot of two 2x1 doubles
    const double b = std::sqrt(a0 * a0 + a1 * a1);
    const double c = std::sqrt(b0 * b0 + b1 * b1);
    return std::fabs(a / (b * c));
}

static void merge_two_lines(const float l1[4], const float l2[4], float out[4]) {  // Frame.cc:1141-1203
    const float ax = l1[0], ay = l1[1], bx = l1[2], by = l1[3];
    const float cx = l2[0], cy = l2[1], dx = l2[2], dy = l2[3];
    const float dlix = bx - ax, dliy = by - ay, dljx = dx - cx, dljy = dy - cy;
    const double li = std::sqrt((double)(dlix * dlix) + (double)(dliy * dliy));
    const double lj = std::sqrt((double)(dljx * dljx) + (double)(dljy * dljy));
    const double xg = (li * (double)(ax + bx) + lj * (double)(cx + dx)) / (double)(2.0 * (li + lj));
    const double yg = (li * (double)(ay + by) + lj * (double)(cy + dy)) / (double)(2.0 * (li + lj));
    const double kPi = 3.1415926535897932384626433832795;
    double thi, thj, thr;
    if (dlix == 0.0f) thi = kPi / 2.0;
    // atan(float) binds to the float overload in the reference (Frame.cc:33 'using namespace std', :1170): a float-precision
    // angle.  The oracle fixes the correctly rounded float; the host libm's atanf is within 1 ulp of it (checked against the
    // executed reference in tests/test_ref_lines.py).
    else thi = (double)(float)std::atan((double)(dliy / dlix));
    if (dljx == 0.0f) thj = kPi / 2.0;
    else thj = (double)(float)std::atan((double)(dljy / dljx));
    if (std::fabs(thi - thj) <= kPi / 2.0) {
        thr = (li * thi + lj * thj) / (li + lj);
    } else {
        const double tmp = thj - kPi * (thj / std::fabs(thj));
        thr = li * thi + lj * tmp;
        thr /= (li + lj);
    }
    const double s = std::sin(thr), c = std::cos(thr);
    const double axg = ((double)ay - yg) * s + ((double)ax - xg) * c;
    const double bxg = ((double)by - yg) * s + ((double)bx - xg) * c;
    const double cxg = ((double)cy - yg) * s + ((double)cx - xg) * c;
    const double dxg = ((double)dy - yg) * s + ((double)dx - xg) * c;
    const double d1 = std::min(axg, std::min(bxg, std::min(cxg, dxg)));
    const double d2 = std::max(axg, std::max(bxg, std::max(cxg, dxg)));
    out[0] = (float)(d1 * c + xg);
    out[1] = (float)(d1 * s + yg);
    out[2] = (float)(d2 * c + xg);
    out[3] = (float)(d2 * s + yg);
}

// cv::clipLine(Size2l, Point2l&, Point2l&) (OpenCV imgproc/src/drawing.cpp), un-vendored
static bool clip_line(long long width, long long height, long long& x1, long long& y1, long long& x2, long long& y2) {
    if (width <= 0 || height <= 0) return false;
    const long long right = width - 1, bottom = height - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (x2 - x1) / (y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (x2 - x1) / (y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (y2 - y1) / (x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (y2 - y1) / (x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// cv::LineIterator(img, Point(pt1), Point(pt2)).count, 8-connected (Point2f -> Point is cvRound)
static int line_iterator_count(int w, int h, float fx1, float fy1, float fx2, float fy2) {
    long long x1 = cvp::cv_round(fx1), y1 = cvp::cv_round(fy1), x2 = cvp::cv_round(fx2), y2 = cvp::cv_round(fy2);
    if ((unsigned long long)x1 >= (unsigned long long)w || (unsigned long long)x2 >= (unsigned long long)w ||
        (unsigned long long)y1 >= (unsigned long long)h || (unsigned long long)y2 >= (unsigned long long)h) {
        if (!clip_line(w, h, x1, y1, x2, y2)) return 0;
    }
    const long long dx = x2 > x1 ? x2 - x1 : x1 - x2, dy = y2 > y1 ? y2 - y1 : y1 - y2;
    return (int)std::max(dx, dy) + 1;
}

}  // namespace cullo

extern "C" {

int orc_clip_line(int w, int h, long long* pts4) {
    return cullo::clip_line(w, h, pts4[0], pts4[1], pts4[2], pts4[3]) ? 1 : 0;
}

// Steps 1-3: keylines_in [n] + linefunc_in [n][3] -> keylines_out [<= n] (sorted, renumbered).  group_of_out[i] (optional,
// n entries) receives for every input line the index of the group leader it was merged into, or -1.  Returns the count.
int orc_cull_lines(const void* keylines_in, const double* linefunc_in, int n, int w, int h, double dis, double angle_deg,
                   double endpoint_dis, void* keylines_out, int32_t* group_of_out) {
    using cullo::KeyLine;
    const KeyLine* kin = (const KeyLine*)keylines_in;
    std::vector<char> tag(n > 0 ? n : 1, 0);
    std::vector<std::vector<int>> robust(n);
    if (group_of_out) std::fill(group_of_out, group_of_out + n, -1);
    const double cos_th = std::cos(angle_deg * 0.0174533);
    for (int i = 0; i < n; ++i) {
        if (tag[i]) continue;
        const KeyLine& l1 = kin[i];
        const double* f1 = linefunc_in + 3 * (size_t)i;
        const double v1[4] = {l1.startPointX, l1.startPointY, l1.endPointX, l1.endPointY};
        for (int j = i + 1; j < n; ++j) {
            if (tag[j]) continue;
            const KeyLine& l2 = kin[j];
            const double* f2 = linefunc_in + 3 * (size_t)j;
            const double v2[4] = {l2.startPointX, l2.startPointY, l2.endPointX, l2.endPointY};
            const float m12x = (l1.startPointX + l1.endPointX) * 0.5f, m12y = (l1.startPointY + l1.endPointY) * 0.5f;
            float m21x = (l2.endPointX + l2.startPointX) * 0.5f, m21y = (l2.endPointY + l2.startPointY) * 0.5f;
            m21x += l2.startPointX; m21y += l2.startPointY;  // sic (:975)
            const double dis12 = cullo::point_line_distance(v2, m12x, m12y);
            const double dis21 = cullo::point_line_distance(v1, m21x, m21y);
            if (!(dis12 < dis || dis21 < dis)) continue;
            const double x11 = l1.startPointX, x12 = l1.endPointX, y11 = l1.startPointY, y12 = l1.endPointY;
            const double x21 = l2.startPointX, x22 = l2.endPointX, y21 = l2.startPointY, y22 = l2.endPointY;
            const double ang = cullo::two_line_angle(f1, f2);
            if (std::fabs(ang) > cos_th) {
                double bx[4] = {x11, x12, x21, x22}, by[4] = {y11, y12, y21, y22};
                std::sort(bx, bx + 4);
                std::sort(by, by + 4);
                const double dx = bx[3] - bx[0], dy = by[3] - by[0];
                const double dx1 = std::fabs(x11 - x12), dx2 = std::fabs(x21 - x22);
                const double dy1 = std::fabs(y11 - y12), dy2 = std::fabs(y21 - y22);
                if (dx > dx1 + dx2) { if (bx[2] - bx[1] > endpoint_dis) continue; }
                if (dy > dy1 + dy2) { if (by[2] - by[1] > endpoint_dis) continue; }
                robust[i].push_back(j);
                tag[i] = 1;
                tag[j] = 1;
                if (group_of_out) { group_of_out[j] = i; group_of_out[i] = i; }
            }
        }
    }
    std::fill(tag.begin(), tag.end(), 0);
    std::vector<std::array<float, 4>> lines;
    for (int i = 0; i < n; ++i) {
        float cur[4] = {kin[i].startPointX, kin[i].startPointY, kin[i].endPointX, kin[i].endPointY};
        for (int j : robust[i]) {
            const float y1[4] = {kin[j].startPointX, kin[j].startPointY, kin[j].endPointX, kin[j].endPointY};
            float m[4];
            cullo::merge_two_lines(cur, y1, m);
            std::memcpy(cur, m, sizeof(cur));
            tag[j] = 1;
            tag[i] = 1;
        }
        if (!robust[i].empty()) lines.push_back({cur[0], cur[1], cur[2], cur[3]});
        if (robust[i].empty() && !tag[i]) lines.push_back({cur[0], cur[1], cur[2], cur[3]});
    }
    const int m = (int)lines.size();
    std::vector<KeyLine> kl(m);
    for (int i = 0; i < m; ++i) {
        const std::array<float, 4>& L = lines[i];
        KeyLine k;
        k.startPointX = L[0]; k.startPointY = L[1]; k.endPointX = L[2]; k.endPointY = L[3];
        k.sPointInOctaveX = L[0]; k.sPointInOctaveY = L[1]; k.ePointInOctaveX = L[2]; k.ePointInOctaveY = L[3];
        const float ddx = L[0] - L[2], ddy = L[1] - L[3];
        k.lineLength = (float)std::sqrt((double)ddx * (double)ddx + (double)ddy * (double)ddy);  // pow(float, 2) promotes to double
        k.octave = 0;
        // the reference calls atan2 on float arguments (float overload, Frame.cc:1076): whatever the host libm's atan2f
        // returns (glibc 2.39: within 1 ulp, not the rounded value in ~16 % of the cases; glibc >= 2.41: correctly rounded).
        // The oracle fixes the correctly rounded result, which is also what the CUDA path produces.
        k.angle = (float)std::atan2((double)(k.endPointY - k.startPointY), (double)(k.endPointX - k.startPointX));
        k.size = (k.endPointX - k.startPointX) * (k.endPointY - k.startPointY);
        k.pt_x = (k.endPointX + k.startPointX) / 2;
        k.pt_y = (k.endPointY + k.startPointY) / 2;
        k.numOfPixels = cullo::line_iterator_count(w, h, L[0], L[1], L[2], L[3]);
        k.response = k.lineLength / (float)std::max(w, h);
        k.class_id = -1;  // KeyLine() default; renumbered below
        kl[i] = k;
    }
    std::vector<int> idx(m);
    std::iota(idx.begin(), idx.end(), 0);
    // std::sort with sort_lines_by_response (Frame.cc:1087): unstable; the order of equal responses is libstdc++'s (pinned by
    // executing the reference: tests/test_ref_lines.py).  Sorting indices gives the same permutation as sorting the structs.
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return kl[a].response > kl[b].response; });
    KeyLine* out = (KeyLine*)keylines_out;
    for (int i = 0; i < m; ++i) { out[i] = kl[idx[i]]; out[i].class_id = i; }
    return m;
}

}  // extern "C"
