// TEST INFRASTRUCTURE — CPU oracle for Manhattan::computeNormalsLPVO, never on the product path.
//
// Restates /root/reference/src/Manhattan.cpp:237-393 (intrinsics :12-18): vertex map for 0.2 < z < 7 (:245-257), tangent
// vectors by +-1 central differences where the five depths lie in [0.2, 7] (:272-303), seven cv::integral images
// (CV_32F -> CV_64F) with row 0 / column 0 removed (:309-328), and every 15 px from (10, 10) the 10x10 box mean of the
// tangents, normal = v x u, cv::normalize (:335-392).
//
// cv::integral is un-vendored OpenCV: its accumulation order (per row a running double sum s, sum[y][x] = sum[y-1][x] + s)
// is restated in orc_integral_f32 and PINNED to cv2 4.13.0 (tests/golden/prims_cv2.npz + live when cv2 is importable).
// cv::normalize of the 3x1 double vector is pinned the same way (orc_normalize3).  The function itself is PINNED BY EXECUTION:
// oracle/_ref/ref_lpvo compiles Manhattan.cpp:237-393 + removeMatRow / removeMatCol (cv::Rect body) against the OpenCV stand-in and
// this file reproduces its output bit for bit (tests/test_lpvo.py, fixture tests/golden/lpvo_ref.npz).  The memcpy body of removeMatRow /
// removeMatCol that Manhattan.cpp is actually built with (USE_CV_RECT is defined in Frame.cc only) sizes CV_64F rows with sizeof(float) and
// turns most normals into NaN; it is executed too (ref_lpvo_asbuilt) but not followed.
//
// Reference bug: the live caller hands the raw CV_16U Mat to this function, which reads it with
// .at<float>.  Oracle and GPU implement the intended behaviour: the float depth image imDepth.convertTo(CV_32F, factor).
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

extern "C" {

// cv::integral(src CV_32F, sum CV_64F) without the leading zero row / column: out[y][x] = sum of src[0..y][0..x].
void orc_integral_f32(const float* src, int W, int H, double* out) {
    for (int y = 0; y < H; ++y) {
        double s = 0;
        for (int x = 0; x < W; ++x) {
            s += (double)src[(size_t)y * W + x];
            out[(size_t)y * W + x] = (y > 0 ? out[(size_t)(y - 1) * W + x] : 0.0) + s;
        }
    }
}

// cv::normalize(v, out) for a 3x1 CV_64F vector (NORM_L2, alpha 1) as cv2 4.13.0 computes it.
void orc_normalize3(const double* v, double* out) {
    const double s = std::fma(v[2], v[2], std::fma(v[1], v[1], v[0] * v[0]));
    const double norm = std::sqrt(s), scale = norm > DBL_EPSILON ? 1.0 / norm : 0.0;
    for (int k = 0; k < 3; ++k) out[k] = v[k] * scale;
}

// normals3: [cap][3] doubles, depth: [cap] floats (vertexMap z at the sample), pix2: [cap][2] ints (u, v).  Returns the count.
int orc_lpvo_normals(const uint16_t* depth16, int W, int H, float depth_factor, float fx, float fy, float cx, float cy,
                     double* normals3, float* depth, int32_t* pix2, int cap) {
    const int cell = 10, density = 15;
    const float inv_fx = 1.0f / fx, inv_fy = 1.0f / fy;
    const size_t px = (size_t)W * H;
    std::vector<float> Z(px), V(px * 3, 0.f), mask(px, 0.f), T(px * 6, 0.f);
    for (size_t i = 0; i < px; ++i) Z[i] = (float)depth16[i] * depth_factor;
    for (int v = 0; v < H; ++v)
        for (int u = 0; u < W; ++u) {
            const float z = Z[(size_t)v * W + u];
            if (z > 0.2f && z < 7.0f) {
                float* p = &V[((size_t)v * W + u) * 3];
                p[0] = (u - cx) * z * inv_fx;
                p[1] = (v - cy) * z * inv_fy;
                p[2] = z;
            }
        }
    auto bad = [](float z) { return z < 0.2f || z > 7.0f; };
    for (int v = 1; v < H - 1; ++v)
        for (int u = 1; u < W - 1; ++u) {
            const size_t i = (size_t)v * W + u;
            if (bad(Z[i]) || bad(Z[i - 1]) || bad(Z[i + 1]) || bad(Z[i - W]) || bad(Z[i + W])) continue;
            mask[i] = 1.0f;
            for (int k = 0; k < 3; ++k) {
                T[(size_t)k * px + i] = V[(i + 1) * 3 + k] - V[(i - 1) * 3 + k];
                T[(size_t)(3 + k) * px + i] = V[(i + W) * 3 + k] - V[(i - W) * 3 + k];
            }
        }
    std::vector<double> I(px * 7);
    for (int k = 0; k < 6; ++k) orc_integral_f32(&T[(size_t)k * px], W, H, &I[(size_t)k * px]);
    orc_integral_f32(mask.data(), W, H, &I[6 * px]);
    auto box = [&](int k, int v, int u) {
        const double* J = &I[(size_t)k * px];
        return J[(size_t)v * W + u] - J[(size_t)(v - cell) * W + u] - J[(size_t)v * W + u - cell] + J[(size_t)(v - cell) * W + u - cell];
    };
    int n = 0;
    for (int v = cell; v < H - 1; v += density)
        for (int u = cell; u < W - 1; u += density) {
            if (mask[(size_t)v * W + u] != 1) continue;
            const int num = (int)box(6, v, u);
            double uv[3], vv[3];
            for (int k = 0; k < 3; ++k) { uv[k] = box(k, v, u) / num; vv[k] = box(3 + k, v, u) / num; }
            const double nv[3] = {vv[1] * uv[2] - vv[2] * uv[1], vv[2] * uv[0] - vv[0] * uv[2], vv[0] * uv[1] - vv[1] * uv[0]};
            if (n < cap) {
                orc_normalize3(nv, normals3 + (size_t)n * 3);
                depth[n] = V[((size_t)v * W + u) * 3 + 2];
                pix2[2 * n] = u; pix2[2 * n + 1] = v;
            }
            ++n;
        }
    return n;
}

}  // extern "C"
