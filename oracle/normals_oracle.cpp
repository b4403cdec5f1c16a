// TEST INFRASTRUCTURE — CPU oracle for the surface normals of Frame::ComputePlanes, never on the product path.
//
// Restates /root/reference/src/Frame.cc:2155-2212: every 3rd pixel of the float depth image is back-projected into an
// organised cloud (invalid depth 0 becomes the finite point (0,0,0): SURVEY App. B #15), then
// pcl::IntegralImageNormalEstimation<PointXYZRGB, Normal> with AVERAGE_3D_GRADIENT, setMaxDepthChangeFactor(0.05),
// setNormalSmoothingSize(10), depth-dependent smoothing off, BORDER_POLICY_IGNORE, viewpoint at the origin; normals at
// odd (row, col) of the cloud are kept.
//
// PCL is an un-vendored dependency (find_package(PCL 1.7), CMakeLists.txt:51) that is absent from this image: the
// algorithm below restates PCL's published integral_image_normal.hpp / integral_image2D.hpp (1.7-1.12 are identical
// for this method): central 3-D differences, double-precision summed-area tables, depth-change mask, two-pass 3-4
// chamfer distance map (float, costs 1.0 / 1.4), window = min(distance, smoothing) when > 2, normal = gy x gx
// normalised in double, flipped towards the viewpoint.  PARITY UNPINNED by execution (no PCL here).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {
struct P3 { float x, y, z; };
}

extern "C" {

// depth16: H x W; cloud: ceil(H/3) x ceil(W/3).  out: per kept pixel 8 floats {nx, ny, nz, px, py, pz, fx_pos, fy_pos}.
// Returns the number of entries ((ch/2) * (cw/2)).
int orc_surface_normals(const uint16_t* depth16, int W, int H, float depth_factor, float fx, float fy, float cx, float cy,
                        float max_depth_change_factor, float smoothing_size, float* out8, float* dist_map_out) {
    const int cw = (int)std::ceil(W / 3.0), ch = (int)std::ceil(H / 3.0);
    std::vector<P3> pts((size_t)cw * ch);
    for (int m = 0, r = 0; m < H; m += 3, ++r)
        for (int n = 0, c = 0; n < W; n += 3, ++c) {
            const float d = (float)depth16[(size_t)m * W + n] * depth_factor;  // imDepth = convertTo(CV_32F, factor)
            P3 p;
            p.z = d;
            p.x = (n - cx) * p.z / fx;
            p.y = (m - cy) * p.z / fy;
            pts[(size_t)r * cw + c] = p;
        }
    // 3-D central differences (interior pixels only; borders stay 0)
    std::vector<float> dx((size_t)cw * ch * 3, 0.f), dy((size_t)cw * ch * 3, 0.f);
    for (int r = 1; r < ch - 1; ++r)
        for (int c = 1; c < cw - 1; ++c) {
            const P3 &L = pts[(size_t)r * cw + c - 1], &R = pts[(size_t)r * cw + c + 1], &U = pts[(size_t)(r - 1) * cw + c], &D = pts[(size_t)(r + 1) * cw + c];
            float* a = &dx[((size_t)r * cw + c) * 3];
            float* b = &dy[((size_t)r * cw + c) * 3];
            a[0] = R.x - L.x; a[1] = R.y - L.y; a[2] = R.z - L.z;
            b[0] = D.x - U.x; b[1] = D.y - U.y; b[2] = D.z - U.z;
        }
    // summed-area tables, (cw+1) x (ch+1), double
    const int iw = cw + 1;
    std::vector<double> IX((size_t)iw * (ch + 1) * 3, 0.0), IY((size_t)iw * (ch + 1) * 3, 0.0);
    for (int r = 0; r < ch; ++r)
        for (int c = 0; c < cw; ++c)
            for (int k = 0; k < 3; ++k) {
                const size_t cur = ((size_t)(r + 1) * iw + c + 1) * 3 + k, up = ((size_t)r * iw + c + 1) * 3 + k,
                             lf = ((size_t)(r + 1) * iw + c) * 3 + k, ul = ((size_t)r * iw + c) * 3 + k;
                IX[cur] = IX[up] + IX[lf] - IX[ul];
                IY[cur] = IY[up] + IY[lf] - IY[ul];
                IX[cur] += (double)dx[((size_t)r * cw + c) * 3 + k];
                IY[cur] += (double)dy[((size_t)r * cw + c) * 3 + k];
            }
    // depth change mask -> distance map
    std::vector<float> dist((size_t)cw * ch, (float)(cw + ch));
    for (int r = 0; r < ch - 1; ++r)
        for (int c = 0; c < cw - 1; ++c) {
            const size_t i = (size_t)r * cw + c;
            const float depth = pts[i].z, depthR = pts[i + 1].z, depthD = pts[i + cw].z;
            const float th = max_depth_change_factor * (std::fabs(depth) + 1.0f) * 2.0f;
            if (std::fabs(depth - depthR) > th || !std::isfinite(depth) || !std::isfinite(depthR)) { dist[i] = 0; dist[i + 1] = 0; }
            if (std::fabs(depth - depthD) > th || !std::isfinite(depth) || !std::isfinite(depthD)) { dist[i] = 0; dist[i + cw] = 0; }
        }
    for (int r = 1; r < ch; ++r)
        for (int c = 1; c < cw; ++c) {
            const float* prev = &dist[(size_t)(r - 1) * cw];
            float* cur = &dist[(size_t)r * cw];
            // as in PCL the upper-right neighbour of the last column is the first element of the current row
            const float upLeft = prev[c - 1] + 1.4f, up = prev[c] + 1.0f, upRight = prev[c + 1] + 1.4f, left = cur[c - 1] + 1.0f;
            const float mn = std::min(std::min(upLeft, up), std::min(left, upRight));
            if (mn < cur[c]) cur[c] = mn;
        }
    for (int r = ch - 2; r >= 0; --r)
        for (int c = cw - 2; c >= 0; --c) {
            const float* next = &dist[(size_t)(r + 1) * cw];
            float* cur = &dist[(size_t)r * cw];
            // as in PCL the lower-left neighbour of column 0 is the last element of the current row
            const float lowerLeft = next[c - 1] + 1.4f, lower = next[c] + 1.0f, lowerRight = next[c + 1] + 1.4f, right = cur[c + 1] + 1.0f;
            const float mn = std::min(std::min(lowerLeft, lower), std::min(right, lowerRight));
            if (mn < cur[c]) cur[c] = mn;
        }
    if (dist_map_out) std::memcpy(dist_map_out, dist.data(), dist.size() * sizeof(float));
    const float nan = std::numeric_limits<float>::quiet_NaN();
    const int border = (int)smoothing_size;
    int n_out = 0;
    for (int m = 0; m < ch; ++m) {
        if (m % 2 == 0) continue;
        for (int n = 0; n < cw; ++n) {
            if (n % 2 == 0) continue;
            float nx = nan, ny = nan, nz = nan;
            const size_t idx = (size_t)m * cw + n;
            if (m >= border && m < ch - border && n >= border && n < cw - border && std::isfinite(pts[idx].z)) {
                const float smoothing = std::min(dist[idx], smoothing_size);
                if (smoothing > 2.0f) {
                    const int rw = (int)smoothing, rh = (int)smoothing;
                    const int sx = n - rw / 2, sy = m - rh / 2;
                    double gx[3], gy[3];
                    for (int k = 0; k < 3; ++k) {
                        const size_t ul = ((size_t)sy * iw + sx) * 3 + k, ur = ul + (size_t)rw * 3, ll = ((size_t)(sy + rh) * iw + sx) * 3 + k, lr = ll + (size_t)rw * 3;
                        gx[k] = IX[lr] + IX[ul] - IX[ur] - IX[ll];
                        gy[k] = IY[lr] + IY[ul] - IY[ur] - IY[ll];
                    }
                    double v[3] = {gy[1] * gx[2] - gy[2] * gx[1], gy[2] * gx[0] - gy[0] * gx[2], gy[0] * gx[1] - gy[1] * gx[0]};
                    const double len2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
                    if (len2 != 0.0) {
                        const double l = std::sqrt(len2);
                        nx = (float)(v[0] / l); ny = (float)(v[1] / l); nz = (float)(v[2] / l);
                        const P3& p = pts[idx];
                        const float cos_theta = (0.f - p.x) * nx + (0.f - p.y) * ny + (0.f - p.z) * nz;
                        if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
                    }
                }
            }
            float* o = out8 + 8 * (size_t)n_out++;
            o[0] = nx; o[1] = ny; o[2] = nz;
            o[3] = pts[idx].x; o[4] = pts[idx].y; o[5] = pts[idx].z;
            o[6] = (float)(n * 3); o[7] = (float)(m * 3);
        }
    }
    return n_out;
}

}  // extern "C"
