// TEST INFRASTRUCTURE — CPU oracle for the line segment detector, never on the product path.
//
// The reference detects segments with cv::createLineSegmentDetector()->detect(octave image)
// (/root/reference/Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:149,158, reached from
// src/LineExtractor.cpp:340-342).  cv::LineSegmentDetector lives in OpenCV imgproc, an un-vendored dependency
// (version unpinned by CMakeLists.txt:32-38); this file restates its published algorithm (von Gioi et al. LSD as
// ported in OpenCV imgproc lsd.cpp) with the default parameters the reference uses: LSD_REFINE_STD, scale 0.8,
// sigma_scale 0.6, quant 2.0, ang_th 22.5, log_eps 0, density_th 0.7, n_bins 1024.
//
// PIN: cv2 4.13.0 in the build container.  The output equals cv2.createLineSegmentDetector().detect() segment for
// segment, bit for bit (tests/test_lsd.py: live when cv2 is importable, and the committed golden vectors
// tests/golden/lsd_cv2.npz).  Seed order: gradient-magnitude bin descending, scan order (y, then x) inside a bin —
// the order cv2 4.13.0 realises (a std::sort order would differ; checked).  Float semantics as everywhere in the
// oracle: no FMA contraction; cos/sin of the float angle are libm cosf/sinf, as in cv2.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "cvprims.hpp"

namespace lsdo {

static const double kPi = 3.1415926535897932384626433832795;  // CV_PI
static const double kNotDef = -1024.0;
static const double k32Pi = (3 * kPi) / 2, k2Pi = 2 * kPi;
static const double kDegToRads = kPi / 180;

struct Params {
    double scale = 0.8, sigma_scale = 0.6, quant = 2.0, ang_th = 22.5, density_th = 0.7;
    int n_bins = 1024;
};

// cv::GaussianBlur(8U, ksize 7, sigma 0.6/0.8 = 0.75): bit-exact Q8 kernel of cv2 4.13.0 (found by search,
// verified in tests/test_lsd.py) — the outer taps round to zero.
static const int kGauss7s075[7] = {0, 4, 56, 136, 56, 4, 0};

// cv::resize(..., Size(), fx, fy, INTER_LINEAR_EXACT) for 8UC1: 8.8 fixed-point coefficients
// (ufixedpoint16 = rint(frac * 256)), horizontal pass exact in 16 bits, vertical pass (v + 2^15) >> 16;
// destination pixels whose source position falls left of 0 / right of n-1 copy the border pixel.
struct LinCoef { int ofs; int c1; int mode; };  // mode 0 interior, 1 low border, 2 high border
static void exact_coeffs(int src, int dst, double inv_scale, std::vector<LinCoef>& out) {
    const double scale = 1.0 / inv_scale;
    out.resize(dst);
    for (int v = 0; v < dst; ++v) {
        double f = scale * ((double)v + 0.5) - 0.5;
        int i = cvp::cv_floor(f);
        LinCoef c{0, 0, 1};
        if (i >= 0 && src > 1) {
            if (i < src - 1) { c.ofs = i; c.c1 = cvp::cv_round((f - (double)i) * 256.0); c.mode = 0; }
            else { c.ofs = src - 1; c.mode = 2; }
        }
        out[v] = c;
    }
}
void resize_linear_exact_u8(const uint8_t* src, int sw, int sh, double fx, double fy, std::vector<uint8_t>& dst, int& dw, int& dh) {
    dw = cvp::cv_round((double)sw * fx);
    dh = cvp::cv_round((double)sh * fy);
    std::vector<LinCoef> cx, cy;
    exact_coeffs(sw, dw, fx, cx);
    exact_coeffs(sh, dh, fy, cy);
    std::vector<uint16_t> H((size_t)sh * dw);
    for (int y = 0; y < sh; ++y) {
        const uint8_t* S = src + (size_t)y * sw;
        uint16_t* Hr = &H[(size_t)y * dw];
        for (int x = 0; x < dw; ++x) {
            const LinCoef& c = cx[x];
            if (c.mode == 1) Hr[x] = (uint16_t)(S[0] << 8);
            else if (c.mode == 2) Hr[x] = (uint16_t)(S[sw - 1] << 8);
            else Hr[x] = (uint16_t)(S[c.ofs] * (256 - c.c1) + S[c.ofs + 1] * c.c1);
        }
    }
    dst.resize((size_t)dw * dh);
    for (int y = 0; y < dh; ++y) {
        const LinCoef& c = cy[y];
        uint8_t* D = &dst[(size_t)y * dw];
        if (c.mode != 0) {
            const uint16_t* Hr = &H[(size_t)(c.mode == 1 ? 0 : sh - 1) * dw];
            for (int x = 0; x < dw; ++x) D[x] = (uint8_t)((Hr[x] + 128) >> 8);
        } else {
            const uint16_t *H0 = &H[(size_t)c.ofs * dw], *H1 = &H[(size_t)(c.ofs + 1) * dw];
            const uint32_t w1 = (uint32_t)c.c1, w0 = 256u - w1;
            for (int x = 0; x < dw; ++x) D[x] = (uint8_t)((H0[x] * w0 + H1[x] * w1 + 32768u) >> 16);
        }
    }
}

struct RegionPoint { int x, y; double angle, modgrad; };
struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };
struct NormPoint { int x, y, norm; };

static inline double dist_sq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
static inline double dist(double x1, double y1, double x2, double y2) { return std::sqrt(dist_sq(x1, y1, x2, y2)); }
static inline double angle_diff_signed(double a, double b) {
    double d = a - b;
    while (d <= -kPi) d += k2Pi;
    while (d > kPi) d -= k2Pi;
    return d;
}
static inline double angle_diff(double a, double b) {
    double d = angle_diff_signed(a, b);
    return d < 0 ? -d : d;
}

class Detector {
public:
    explicit Detector(const Params& p) : P(p) {}

    // stage outputs kept for inspection by the tests
    std::vector<uint8_t> scaled;      // blurred + resized image
    int w = 0, h = 0;                 // its size
    std::vector<double> angles, modgrad;
    std::vector<NormPoint> ordered;
    double max_grad = -1;
    std::vector<float> segments;      // x1,y1,x2,y2 per segment
    std::vector<double> widths;

    void detect(const uint8_t* img, int iw, int ih) {
        const double prec = kPi * P.ang_th / 180;
        const double p = P.ang_th / 180;
        const double rho = P.quant / std::sin(prec);
        if (P.scale != 1) {
            std::vector<uint8_t> blurred((size_t)iw * ih);
            cvp::gaussian_blur_u8_q8(img, iw, ih, iw, blurred.data(), iw, kGauss7s075, 7);
            resize_linear_exact_u8(blurred.data(), iw, ih, P.scale, P.scale, scaled, w, h);
        } else {
            scaled.assign(img, img + (size_t)iw * ih);
            w = iw; h = ih;
        }
        ll_angle(rho);
        const double log_nt = 5 * (std::log10((double)w) + std::log10((double)h)) / 2 + std::log10(11.0);
        const size_t min_reg_size = (size_t)(-log_nt / std::log10(p));
        used.assign((size_t)w * h, 0);
        segments.clear();
        widths.clear();
        std::vector<RegionPoint> reg;
        for (size_t i = 0; i < ordered.size(); ++i) {
            const int sx = ordered[i].x, sy = ordered[i].y;
            const size_t si = (size_t)sy * w + sx;
            if (used[si] || angles[si] == kNotDef) continue;
            double reg_angle;
            region_grow(sx, sy, reg, reg_angle, prec);
            if (reg.size() < min_reg_size) continue;
            Rect rec;
            region2rect(reg, reg_angle, prec, p, rec);
            if (!refine(reg, reg_angle, prec, p, rec, P.density_th)) continue;
            rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
            if (P.scale != 1) { rec.x1 /= P.scale; rec.y1 /= P.scale; rec.x2 /= P.scale; rec.y2 /= P.scale; rec.width /= P.scale; }
            segments.push_back((float)rec.x1); segments.push_back((float)rec.y1);
            segments.push_back((float)rec.x2); segments.push_back((float)rec.y2);
            widths.push_back(rec.width);
        }
    }

private:
    Params P;
    std::vector<uint8_t> used;

    void ll_angle(double threshold) {
        angles.assign((size_t)w * h, kNotDef);
        modgrad.assign((size_t)w * h, 0.0);
        max_grad = -1;
        for (int y = 0; y < h - 1; ++y) {
            const uint8_t *r0 = &scaled[(size_t)y * w], *r1 = r0 + w;
            for (int x = 0; x < w - 1; ++x) {
                int DA = r1[x + 1] - r0[x], BC = r0[x + 1] - r1[x];
                int gx = DA + BC, gy = DA - BC;
                double norm = std::sqrt((gx * gx + gy * gy) / 4.0);
                modgrad[(size_t)y * w + x] = norm;
                if (norm <= threshold) continue;
                angles[(size_t)y * w + x] = cvp::fast_atan2_deg((float)gx, (float)-gy) * kDegToRads;
                if (norm > max_grad) max_grad = norm;
            }
        }
        const double bin_coef = (max_grad > 0) ? double(P.n_bins - 1) / max_grad : 0;
        ordered.clear();
        ordered.reserve((size_t)(w - 1) * (h - 1));
        for (int y = 0; y < h - 1; ++y)
            for (int x = 0; x < w - 1; ++x) ordered.push_back({x, y, (int)(modgrad[(size_t)y * w + x] * bin_coef)});
        auto cmp = [](const NormPoint& a, const NormPoint& b) { return a.norm > b.norm; };
        std::stable_sort(ordered.begin(), ordered.end(), cmp);
    }

    bool is_aligned(int x, int y, double theta, double prec) const {
        if (x < 0 || y < 0 || x >= w || y >= h) return false;
        const double a = angles[(size_t)y * w + x];
        if (a == kNotDef) return false;
        double n = theta - a;
        if (n < 0) n = -n;
        if (n > k32Pi) { n -= k2Pi; if (n < 0) n = -n; }
        return n <= prec;
    }

    void region_grow(int sx, int sy, std::vector<RegionPoint>& reg, double& reg_angle, double prec) {
        reg.clear();
        const size_t si = (size_t)sy * w + sx;
        reg_angle = angles[si];
        reg.push_back({sx, sy, reg_angle, modgrad[si]});
        float sumdx = (float)std::cos(reg_angle), sumdy = (float)std::sin(reg_angle);
        used[si] = 1;
        for (size_t i = 0; i < reg.size(); ++i) {
            const int px = reg[i].x, py = reg[i].y;
            const int x0 = std::max(px - 1, 0), x1 = std::min(px + 1, w - 1);
            const int y0 = std::max(py - 1, 0), y1 = std::min(py + 1, h - 1);
            for (int yy = y0; yy <= y1; ++yy)
                for (int xx = x0; xx <= x1; ++xx) {
                    const size_t ni = (size_t)yy * w + xx;
                    if (used[ni] || !is_aligned(xx, yy, reg_angle, prec)) continue;
                    const double a = angles[ni];
                    used[ni] = 1;
                    reg.push_back({xx, yy, a, modgrad[ni]});
                    sumdx += std::cos((float)a);   // float overloads (cosf/sinf), float accumulators
                    sumdy += std::sin((float)a);
                    reg_angle = cvp::fast_atan2_deg(sumdy, sumdx) * kDegToRads;
                }
        }
    }

    double get_theta(const std::vector<RegionPoint>& reg, double x, double y, double reg_angle, double prec) const {
        double Ixx = 0, Iyy = 0, Ixy = 0;
        for (const RegionPoint& r : reg) {
            double dx = (double)r.x - x, dy = (double)r.y - y;
            Ixx += dy * dy * r.modgrad;
            Iyy += dx * dx * r.modgrad;
            Ixy -= dx * dy * r.modgrad;
        }
        double lambda = 0.5 * (Ixx + Iyy - std::sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
        double theta = (std::fabs(Ixx) > std::fabs(Iyy)) ? (double)cvp::fast_atan2_deg((float)(lambda - Ixx), (float)Ixy)
                                                         : (double)cvp::fast_atan2_deg((float)Ixy, (float)(lambda - Iyy));
        theta *= kDegToRads;
        if (angle_diff(theta, reg_angle) > prec) theta += kPi;
        return theta;
    }

    void region2rect(const std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec) const {
        double x = 0, y = 0, sum = 0;
        for (const RegionPoint& r : reg) { x += (double)r.x * r.modgrad; y += (double)r.y * r.modgrad; sum += r.modgrad; }
        x /= sum; y /= sum;
        const double theta = get_theta(reg, x, y, reg_angle, prec);
        const double dx = std::cos(theta), dy = std::sin(theta);
        double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
        for (const RegionPoint& r : reg) {
            double rx = (double)r.x - x, ry = (double)r.y - y;
            double l = rx * dx + ry * dy, ww = -rx * dy + ry * dx;
            if (l > l_max) l_max = l; else if (l < l_min) l_min = l;
            if (ww > w_max) w_max = ww; else if (ww < w_min) w_min = ww;
        }
        rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy; rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
        rec.width = w_max - w_min; rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
        if (rec.width < 1.0) rec.width = 1.0;
    }

    bool refine(std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec, double density_th) {
        double density = (double)reg.size() / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density >= density_th) return true;
        const double xc = (double)reg[0].x, yc = (double)reg[0].y, ang_c = reg[0].angle;
        double sum = 0, s_sum = 0;
        int n = 0;
        for (const RegionPoint& r : reg) {
            used[(size_t)r.y * w + r.x] = 0;
            if (dist(xc, yc, (double)r.x, (double)r.y) < rec.width) {
                double d = angle_diff_signed(r.angle, ang_c);
                sum += d; s_sum += d * d; ++n;
            }
        }
        const double mean_angle = sum / (double)n;
        const double tau = 2.0 * std::sqrt((s_sum - 2.0 * mean_angle * sum) / (double)n + mean_angle * mean_angle);
        const int sx = reg[0].x, sy = reg[0].y;
        region_grow(sx, sy, reg, reg_angle, tau);
        if (reg.size() < 2) return false;
        region2rect(reg, reg_angle, prec, p, rec);
        density = (double)reg.size() / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density < density_th) return reduce_region_radius(reg, reg_angle, prec, p, rec, density, density_th);
        return true;
    }

    bool reduce_region_radius(std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec, double density,
                              double density_th) {
        const double xc = (double)reg[0].x, yc = (double)reg[0].y;
        const double r1 = dist_sq(xc, yc, rec.x1, rec.y1), r2 = dist_sq(xc, yc, rec.x2, rec.y2);
        double rad_sq = r1 > r2 ? r1 : r2;
        while (density < density_th) {
            rad_sq *= 0.75 * 0.75;
            for (size_t i = 0; i < reg.size(); ++i) {
                if (dist_sq(xc, yc, (double)reg[i].x, (double)reg[i].y) > rad_sq) {
                    used[(size_t)reg[i].y * w + reg[i].x] = 0;
                    std::swap(reg[i], reg[reg.size() - 1]);
                    reg.pop_back();
                    --i;
                }
            }
            if (reg.size() < 2) return false;
            region2rect(reg, reg_angle, prec, p, rec);
            density = (double)reg.size() / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        }
        return true;
    }
};

}  // namespace lsdo

extern "C" {

// segments4: [cap][4] float32 (x1,y1,x2,y2 in input-image coordinates); returns the segment count (may exceed cap).
// scaled_out (optional): the blurred + resized image, *sw x *sh.
int orc_lsd_detect(const uint8_t* gray, int w, int h, float* segments4, int cap, uint8_t* scaled_out, int* sw, int* sh) {
    lsdo::Params p;
    lsdo::Detector d(p);
    d.detect(gray, w, h);
    int n = (int)(d.segments.size() / 4);
    int m = n < cap ? n : cap;
    if (m) std::memcpy(segments4, d.segments.data(), (size_t)m * 16);
    if (sw) *sw = d.w;
    if (sh) *sh = d.h;
    if (scaled_out) std::memcpy(scaled_out, d.scaled.data(), d.scaled.size());
    return n;
}

void orc_lsd_scaled_size(int w, int h, int* sw, int* sh) {
    *sw = cvp::cv_round((double)w * 0.8);
    *sh = cvp::cv_round((double)h * 0.8);
}

}  // extern "C"
