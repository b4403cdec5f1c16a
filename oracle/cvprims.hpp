// TEST INFRASTRUCTURE — CPU oracle, never on the product path.
//
// Restatement of the third-party OpenCV primitives the reference's front-end calls. OpenCV's C++
// library is not vendored in /root/reference and is absent from this image; these functions restate the
// published algorithms of OpenCV 4.x imgproc/features2d and are PINNED bit-exactly against the Python
// cv2 4.13.0 wheel that IS in this image (tests/test_oracle_prims.py live, tests/golden/*.npz offline).
//
// Reference call sites (relative to /root/reference):
//   cv::resize INTER_LINEAR 8UC1      src/ORBextractor.cc:1118
//   cv::GaussianBlur 7x7 s=2          src/ORBextractor.cc:1084
//   cv::GaussianBlur 5x5 s=1          Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:358
//   cv::Sobel 3x3 -> CV_16S           Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:395-396
//   cv::FAST(.., thr, true)           src/ORBextractor.cc:807,812
//   cv::fastAtan2                     src/ORBextractor.cc:101
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace cvp {

// cvRound: round-half-to-even (x86 cvtss2si / lrint under the default rounding mode).
static inline int cv_round(float v) { return (int)lrintf(v); }
static inline int cv_round(double v) { return (int)lrint(v); }
static inline int cv_floor(float v) { return (int)floorf(v); }
static inline int cv_floor(double v) { return (int)floor(v); }
static inline int cv_ceil(float v) { return (int)ceilf(v); }
static inline int cv_ceil(double v) { return (int)ceil(v); }

static inline int reflect101(int p, int n) {
    // BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

// ---- cv::resize, 8UC1, INTER_LINEAR (fixed-point 11-bit coefficients) -------------------------------
static inline short sat_short_round(float v) {
    int i = cv_round(v);
    return (short)std::min(32767, std::max(-32768, i));
}

inline void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                             uint8_t* dst, int dw, int dh, size_t dstride) {
    const double scale_x = 1.0 / ((double)dw / sw);
    const double scale_y = 1.0 / ((double)dh / sh);
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> xa(2 * dw), ya(2 * dh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        xa[2 * dx] = sat_short_round((1.f - fx) * 2048.f);
        xa[2 * dx + 1] = sat_short_round(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        ya[2 * dy] = sat_short_round((1.f - fy) * 2048.f);
        ya[2 * dy + 1] = sat_short_round(fy * 2048.f);
    }
    std::vector<int> row0(dw), row1(dw);
    int have0 = -1000000, have1 = -1000000;  // which source rows are cached
    auto hrow = [&](int sy, std::vector<int>& out) {
        const uint8_t* S = src + (size_t)sy * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx];
            int s1 = sx + 1 < sw ? sx + 1 : sx;  // weight is 0 whenever this clamps
            out[dx] = S[sx] * xa[2 * dx] + S[s1] * xa[2 * dx + 1];
        }
    };
    for (int dy = 0; dy < dh; ++dy) {
        int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
        int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
        if (have1 == sy0) { row0.swap(row1); std::swap(have0, have1); }
        if (have0 != sy0) { hrow(sy0, row0); have0 = sy0; }
        if (have1 != sy1) { hrow(sy1, row1); have1 = sy1; }
        const int b0 = ya[2 * dy], b1 = ya[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)std::min(255, std::max(0, v));
        }
    }
}

// ---- cv::GaussianBlur, 8UC1, BORDER_REFLECT_101, fixed-point Q8 separable kernel -------------------
// K taps, weights sum to 256; H pass exact in 16 bits, V pass 32 bits, round (v + 2^15) >> 16.
inline void gaussian_blur_u8_q8(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst,
                                size_t dstride, const int* k, int ksize) {
    const int r = ksize / 2;
    std::vector<uint16_t> H((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* S = src + (size_t)y * sstride;
        uint16_t* Hr = &H[(size_t)y * w];
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            if (x >= r && x + r < w) {
                for (int i = 0; i < ksize; ++i) acc += S[x + i - r] * k[i];
            } else {
                for (int i = 0; i < ksize; ++i) acc += S[reflect101(x + i - r, w)] * k[i];
            }
            Hr[x] = (uint16_t)acc;
        }
    }
    std::vector<const uint16_t*> rows(ksize);
    for (int y = 0; y < h; ++y) {
        for (int i = 0; i < ksize; ++i) rows[i] = &H[(size_t)reflect101(y + i - r, h) * w];
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int i = 0; i < ksize; ++i) acc += (uint32_t)rows[i][x] * (uint32_t)k[i];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}
static const int kGauss7s2[7] = {18, 34, 48, 56, 48, 34, 18};  // 7x7, sigma 2   (ORB)
static const int kGauss5s1[5] = {14, 62, 104, 62, 14};         // 5x5, sigma 1   (LBD)

inline void gaussian_blur7_s2(const uint8_t* src, int w, int h, size_t ss, uint8_t* dst, size_t ds) {
    gaussian_blur_u8_q8(src, w, h, ss, dst, ds, kGauss7s2, 7);
}
inline void gaussian_blur5_s1(const uint8_t* src, int w, int h, size_t ss, uint8_t* dst, size_t ds) {
    gaussian_blur_u8_q8(src, w, h, ss, dst, ds, kGauss5s1, 5);
}

// ---- cv::Sobel 3x3 -> CV_16S, BORDER_REFLECT_101 ---------------------------------------------------
inline void sobel3_s16(const uint8_t* src, int w, int h, size_t ss, int16_t* dx, int16_t* dy) {
    for (int y = 0; y < h; ++y) {
        const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * ss;
        const uint8_t* r1 = src + (size_t)y * ss;
        const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * ss;
        for (int x = 0; x < w; ++x) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            int gx = (r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]);
            int gy = (r2[xm] - r0[xm]) + 2 * (r2[x] - r0[x]) + (r2[xp] - r0[xp]);
            dx[(size_t)y * w + x] = (int16_t)gx;
            dy[(size_t)y * w + x] = (int16_t)gy;
        }
    }
}

// ---- cv::FAST, TYPE_9_16 ---------------------------------------------------------------------------
static const int kFastDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kFastDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// Threshold-independent corner strength: S = max(A, B) - 1 with
//   A = max over the 16 contiguous 9-arcs of min(v - ring), B = same for (ring - v).
// A pixel is a FAST-9/16 corner at threshold t  <=>  S >= t, and for such pixels S equals the value
// OpenCV's cornerScore<16>() ladder returns (its accumulator starts at t and can only grow).
static inline int fast_strength(const uint8_t* p, const int* ofs) {
    int d[25];
    const int v = *p;
    for (int k = 0; k < 16; ++k) d[k] = v - p[ofs[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int A = -256, B = -256;
    for (int s = 0; s < 16; ++s) {
        int mn = d[s], mx = d[s];
        for (int i = 1; i < 9; ++i) { mn = std::min(mn, d[s + i]); mx = std::max(mx, d[s + i]); }
        A = std::max(A, mn);
        B = std::max(B, -mx);
    }
    return std::max(A, B) - 1;
}

struct FastKp { int x, y, score; };

// FAST with non-max suppression on a (sub)image: 3-px dead border, NMS strict '>' over the 8
// neighbours (non-corners and border pixels count as 0), output row-major.
inline void fast9_nms(const uint8_t* img, int w, int h, size_t stride, int thr, std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    int ofs[16];
    for (int k = 0; k < 16; ++k) ofs[k] = kFastDy[k] * (int)stride + kFastDx[k];
    std::vector<int> sc((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t* row = img + (size_t)y * stride;
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* p = row + x;
            // quick reject: every 9-arc contains ring[0] or ring[8], and ring[4] or ring[12]
            int v = *p;
            int d0 = v - p[ofs[0]], d8 = v - p[ofs[8]];
            bool pos = d0 > thr || d8 > thr, neg = d0 < -thr || d8 < -thr;
            if (!pos && !neg) continue;
            int d4 = v - p[ofs[4]], d12 = v - p[ofs[12]];
            pos = pos && (d4 > thr || d12 > thr);
            neg = neg && (d4 < -thr || d12 < -thr);
            if (!pos && !neg) continue;
            int s = fast_strength(p, ofs);
            if (s >= thr) sc[(size_t)y * w + x] = s;
        }
    }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = sc[(size_t)y * w + x];
            if (s == 0 && thr > 0) continue;
            if (s < thr) continue;
            const int* c = &sc[(size_t)y * w + x];
            if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] && s > c[w - 1] &&
                s > c[w] && s > c[w + 1])
                out.push_back({x, y, s});
        }
}

// ---- cv::fastAtan2 (degrees) -----------------------------------------------------------------------
static inline float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

}  // namespace cvp
