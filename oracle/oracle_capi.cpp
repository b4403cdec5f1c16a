// TEST INFRASTRUCTURE — C entry points of the CPU oracle for ctypes (tests/, bench.py cpu_baseline only).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

#include "cvprims.hpp"
#include "orb_oracle.hpp"

extern "C" {

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
    cvp::resize_linear_u8(src, sw, sh, sw, dst, dw, dh, dw);
}
void orc_blur7(const uint8_t* src, int w, int h, uint8_t* dst) { cvp::gaussian_blur7_s2(src, w, h, w, dst, w); }
void orc_blur5(const uint8_t* src, int w, int h, uint8_t* dst) { cvp::gaussian_blur5_s1(src, w, h, w, dst, w); }
void orc_sobel3(const uint8_t* src, int w, int h, int16_t* dx, int16_t* dy) { cvp::sobel3_s16(src, w, h, w, dx, dy); }
int orc_fast9(const uint8_t* src, int w, int h, int stride, int thr, int32_t* out_xys, int cap) {
    std::vector<cvp::FastKp> r;
    cvp::fast9_nms(src, w, h, stride, thr, r);
    int n = (int)r.size();
    for (int i = 0; i < n && i < cap; ++i) { out_xys[3 * i] = r[i].x; out_xys[3 * i + 1] = r[i].y; out_xys[3 * i + 2] = r[i].score; }
    return n;
}
void orc_fast_atan2(const float* y, const float* x, float* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = cvp::fast_atan2_deg(y[i], x[i]);
}

// ---- ORB --------------------------------------------------------------------------------------------
void* orc_orb_create(int nfeatures, float scale, int nlevels, int ini_th, int min_th) {
    orbo::Params p;
    p.nfeatures = nfeatures; p.scale_factor = scale; p.nlevels = nlevels; p.ini_th = ini_th; p.min_th = min_th;
    return new orbo::Extractor(p);
}
void orc_orb_destroy(void* h) { delete (orbo::Extractor*)h; }
int orc_orb_extract(void* h, const uint8_t* gray, int w, int hh, int stride, orbo::KeyPoint* kps, uint8_t* desc, int cap) {
    std::vector<orbo::KeyPoint> k;
    std::vector<uint8_t> d;
    ((orbo::Extractor*)h)->extract(gray, w, hh, (size_t)stride, k, d);
    int n = (int)k.size();
    int m = n < cap ? n : cap;
    if (m) { std::memcpy(kps, k.data(), (size_t)m * sizeof(orbo::KeyPoint)); std::memcpy(desc, d.data(), (size_t)m * 32); }
    return n;
}
void orc_orb_tables(void* h, float* sf, float* isf, int* nfeat, int* umax) {
    orbo::Extractor* e = (orbo::Extractor*)h;
    for (size_t i = 0; i < e->scale_factors().size(); ++i) { sf[i] = e->scale_factors()[i]; isf[i] = e->inv_scale_factors()[i]; nfeat[i] = e->features_per_level()[i]; }
    for (int i = 0; i < 16; ++i) umax[i] = e->umax()[i];
}
void orc_orb_level_info(void* h, int l, int* out4) {
    const orbo::Level& L = ((orbo::Extractor*)h)->levels()[l];
    out4[0] = L.w; out4[1] = L.h; out4[2] = (int)L.cand.size(); out4[3] = (int)L.kps.size();
}
void orc_orb_level_image(void* h, int l, uint8_t* out, int blurred) {
    const orbo::Level& L = ((orbo::Extractor*)h)->levels()[l];
    const std::vector<uint8_t>& v = blurred ? L.blurred : L.img;
    if (!v.empty()) std::memcpy(out, v.data(), v.size());
}
void orc_orb_level_cand(void* h, int l, float* out3) {
    const orbo::Level& L = ((orbo::Extractor*)h)->levels()[l];
    for (size_t i = 0; i < L.cand.size(); ++i) { out3[3 * i] = L.cand[i].x; out3[3 * i + 1] = L.cand[i].y; out3[3 * i + 2] = L.cand[i].response; }
}
void orc_orb_level_kps(void* h, int l, orbo::KeyPoint* out) {
    const orbo::Level& L = ((orbo::Extractor*)h)->levels()[l];
    if (!L.kps.empty()) std::memcpy(out, L.kps.data(), L.kps.size() * sizeof(orbo::KeyPoint));
}
int orc_orb_distribute(const float* cand3, int n, int minX, int maxX, int minY, int maxY, int N, float* out3) {
    std::vector<orbo::Candidate> in(n);
    for (int i = 0; i < n; ++i) in[i] = {cand3[3 * i], cand3[3 * i + 1], cand3[3 * i + 2]};
    std::vector<orbo::Candidate> r = orbo::Extractor::distribute(in, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < r.size(); ++i) { out3[3 * i] = r[i].x; out3[3 * i + 1] = r[i].y; out3[3 * i + 2] = r[i].response; }
    return (int)r.size();
}
// Frame-parallel batch (CPU baseline): frames are independent, one Extractor per thread.
// counts[f] = number of keypoints of frame f; kps/desc (may be null) hold cap slots per frame.
void orc_orb_extract_batch(int nfeatures, float scale, int nlevels, int ini_th, int min_th, const uint8_t* frames,
                           int nframes, int w, int h, int nthreads, int32_t* counts, orbo::KeyPoint* kps,
                           uint8_t* desc, int cap) {
    orbo::Params p;
    p.nfeatures = nfeatures; p.scale_factor = scale; p.nlevels = nlevels; p.ini_th = ini_th; p.min_th = min_th;
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([=]() {
            orbo::Extractor ex(p);
            std::vector<orbo::KeyPoint> k;
            std::vector<uint8_t> d;
            for (int f = t; f < nframes; f += nthreads) {
                ex.extract(frames + (size_t)f * w * h, w, h, (size_t)w, k, d);
                counts[f] = (int32_t)k.size();
                int m = (int)k.size() < cap ? (int)k.size() : cap;
                if (kps && m) std::memcpy(kps + (size_t)f * cap, k.data(), (size_t)m * sizeof(orbo::KeyPoint));
                if (desc && m) std::memcpy(desc + (size_t)f * cap * 32, d.data(), (size_t)m * 32);
            }
        });
    for (auto& t : th) t.join();
}

// ---- whole front-end, frame-parallel (CPU baseline of bench.py) ---------------------------------------------
// Per frame, what Frame::Frame runs on its three threads (src/Frame.cc:205-233) as far as the oracle restates it:
//   ORBextractor::operator()                                    (orb_oracle.cpp)
//   LINEextractor::operator(): LSD -> KeyLines -> top-N -> LBD  (lsd_oracle.cpp, lbd_oracle.cpp)
//   PlaneDetection::readDepthImage + runPlaneDetection          (plane_oracle.cpp)
//   surface normals of Frame::ComputePlanes                     (normals_oracle.cpp)
int orc_lsd_detect(const uint8_t* gray, int w, int h, float* segments4, int cap, uint8_t* scaled_out, int* sw, int* sh);
void orc_keylines_from_segments(const float* seg4, int n, int w, int h, void* keylines_out);
void orc_lbd_compute(const uint8_t* gray, int w, int h, const void* keylines, int n, uint8_t* desc, float* fdesc);
int orc_plane_detect(const uint16_t* depth, int w, int h, float factor, float fx, float fy, float cx, float cy, double* planes7,
                     int max_planes, int32_t* membership);
int orc_surface_normals(const uint16_t* depth16, int W, int H, float depth_factor, float fx, float fy, float cx, float cy,
                        float max_depth_change_factor, float smoothing_size, float* out8, float* dist_map_out);

// LINEextractor::operator() (src/LineExtractor.cpp:329-380) on one frame; keylines: cap x 68 B, desc: cap x 32 B.
int orc_cull_lines(const void* keylines_in, const double* linefunc_in, int n, int w, int h, double dis, double angle_deg,
                   double endpoint_dis, void* keylines_out, int32_t* group_of_out);

// std::sort(first, last, sort_lines_by_response()) on indices: idx_out[k] = input position of the k-th line after the sort
void orc_sort_by_response(const float* resp, int n, int32_t* idx_out) {
    std::iota(idx_out, idx_out + n, 0);
    std::sort(idx_out, idx_out + n, [&](int a, int b) { return resp[a] > resp[b]; });
}

// cull != 0: followed by Frame::cullingLine(im, 5, 2.5, 15, 30) (src/Frame.cc:939), second LBD pass included
int orc_line_extract(const uint8_t* gray, int w, int h, int nfeat, void* keylines, uint8_t* desc, int cap, int cull) {
    std::vector<float> seg((size_t)4 * 16384);
    int n = orc_lsd_detect(gray, w, h, seg.data(), 16384, nullptr, nullptr, nullptr);
    if (n > 16384) n = 16384;
    std::vector<uint8_t> kl((size_t)n * 68);
    orc_keylines_from_segments(seg.data(), n, w, h, kl.data());
    std::vector<uint8_t> sel;
    if (n > nfeat) {  // sort_lines_by_response, truncate, renumber class_id (:351-360); std::sort: ties in libstdc++'s order
        std::vector<int> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        auto resp = [&](int i) { float r; std::memcpy(&r, &kl[(size_t)i * 68 + 20], 4); return r; };
        std::sort(idx.begin(), idx.end(), [&](int a, int b) { return resp(a) > resp(b); });
        sel.resize((size_t)nfeat * 68);
        for (int i = 0; i < nfeat; ++i) {
            std::memcpy(&sel[(size_t)i * 68], &kl[(size_t)idx[i] * 68], 68);
            const int32_t cid = i;
            std::memcpy(&sel[(size_t)i * 68 + 4], &cid, 4);
        }
        kl.swap(sel);
        n = nfeat;
    }
    std::vector<uint8_t> d((size_t)(n > 0 ? n : 1) * 32);
    if (n > 0) orc_lbd_compute(gray, w, h, kl.data(), n, d.data(), nullptr);
    if (cull && n > 0) {
        std::vector<double> lv((size_t)n * 3);
        for (int i = 0; i < n; ++i) {  // mvKeyLineFunctions (LineExtractor.cpp:365-377)
            float e[4];
            std::memcpy(e, &kl[(size_t)i * 68 + 28], 16);
            const double sx = e[0], sy = e[1], ex = e[2], ey = e[3];
            const double l0 = sy - ey, l1 = ex - sx, l2 = sx * ey - sy * ex, nn = std::sqrt(l0 * l0 + l1 * l1);
            lv[3 * i] = l0 / nn; lv[3 * i + 1] = l1 / nn; lv[3 * i + 2] = l2 / nn;
        }
        std::vector<uint8_t> kl2((size_t)n * 68);
        n = orc_cull_lines(kl.data(), lv.data(), n, w, h, 5.0, 2.5, 15.0, kl2.data(), nullptr);
        kl.swap(kl2);
        if (n > 0) orc_lbd_compute(gray, w, h, kl.data(), n, d.data(), nullptr);
    }
    const int m = n < cap ? n : cap;
    if (m > 0) {
        std::memcpy(keylines, kl.data(), (size_t)m * 68);
        std::memcpy(desc, d.data(), (size_t)m * 32);
    }
    return n;
}

// stages: bit 0 ORB, bit 1 lines, bit 2 planes, bit 3 normals, bit 4 cullingLine after the line extractor.  counts4[f] = {keypoints, lines, planes, normals}.
void orc_frontend_batch(const uint8_t* gray, const uint16_t* depth, int nframes, int w, int h, int nthreads, int stages,
                        int nfeatures, float scale, int nlevels, int ini_th, int min_th, int nlines, float depth_factor, float fx,
                        float fy, float cx, float cy, int32_t* counts4) {
    orbo::Params p;
    p.nfeatures = nfeatures; p.scale_factor = scale; p.nlevels = nlevels; p.ini_th = ini_th; p.min_th = min_th;
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([=]() {
            orbo::Extractor ex(p);
            std::vector<orbo::KeyPoint> k;
            std::vector<uint8_t> d;
            std::vector<uint8_t> kl((size_t)nlines * 68), ld((size_t)nlines * 32);
            std::vector<double> planes(64 * 7);
            std::vector<int32_t> mem((size_t)w * h);
            const int cw = (int)std::ceil(w / 3.0), ch = (int)std::ceil(h / 3.0);
            std::vector<float> nrm((size_t)(cw / 2) * (ch / 2) * 8), dist((size_t)cw * ch);
            for (int f = t; f < nframes; f += nthreads) {
                const uint8_t* g = gray + (size_t)f * w * h;
                const uint16_t* dp = depth + (size_t)f * w * h;
                int32_t* c = counts4 + 4 * (size_t)f;
                c[0] = c[1] = c[2] = c[3] = 0;
                if (stages & 1) { ex.extract(g, w, h, (size_t)w, k, d); c[0] = (int32_t)k.size(); }
                if (stages & 2) c[1] = orc_line_extract(g, w, h, nlines, kl.data(), ld.data(), nlines, stages & 16);
                if (stages & 4) c[2] = orc_plane_detect(dp, w, h, depth_factor, fx, fy, cx, cy, planes.data(), 64, mem.data());
                if (stages & 8) c[3] = orc_surface_normals(dp, w, h, depth_factor, fx, fy, cx, cy, 0.05f, 10.0f, nrm.data(), dist.data());
            }
        });
    for (auto& t : th) t.join();
}

}  // extern "C"
