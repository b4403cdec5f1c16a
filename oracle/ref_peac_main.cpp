// TEST INFRASTRUCTURE — driver for the reference's own plane extractor: src/PlaneExtractor.cpp and include/peac/*.hpp
// are compiled UNMODIFIED from /root/reference against the OpenCV stand-in (oracle/cvshim) and the Eigen stand-in
// (oracle/eigenshim).  Built into oracle/_ref/ref_peac by oracle/Makefile.
//
//   ref_peac <in.bin> <out.bin> [eigen perturbation, only in the ref_peac_perturb build]
//   in : int32 {magic 0x50454143, w, h, nframes}, float {fx, fy, cx, cy, factor}, frames (w*h u16 each)
//   out: per frame
//          int32 nblocks; per 10x10 block {int32 N, nouse; double center[3], normal[3], mse, curvature}   (AHCPlaneSeg.hpp:211-284)
//          int32 nplanes; per plane {int32 N, nvertices; double normal[3], center[3], mse}               (extractedPlanes, plane_vertices_)
//          int32 membership[h*w]                                                                         (plane_filter.membershipImg)
//
// Determinism: ahc::PlaneSeg::nbs is a std::set<PlaneSeg*> (AHCPlaneSeg.hpp:188), iterated in ADDRESS order by ahCluster
// (AHCPlaneFitter.hpp:1031-1050).  This driver serves `new PlaneSeg` from a monotonic arena so that address order ==
// creation order, the rule the oracle restatement documents.  Everything else uses malloc.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>

#include "PlaneExtractor.h"

static char* g_arena = nullptr;
static size_t g_arena_cap = 0, g_arena_off = 0;

void* operator new(size_t n) {
    if (n == sizeof(ahc::PlaneSeg) && g_arena) {
        const size_t need = (n + 15) & ~(size_t)15;
        if (g_arena_off + need > g_arena_cap) { std::fprintf(stderr, "ref_peac: PlaneSeg arena exhausted\n"); std::abort(); }
        void* p = g_arena + g_arena_off;
        g_arena_off += need;
        return p;
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void* p) noexcept {
    if (g_arena && (char*)p >= g_arena && (char*)p < g_arena + g_arena_cap) return;
    std::free(p);
}
void operator delete(void* p, size_t) noexcept { operator delete(p); }

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: ref_peac in.bin out.bin\n"); return 2; }
#ifdef EIGENSHIM_PERTURB
    if (argc > 3) Eigen::SelfAdjointEigenSolver<Eigen::Matrix3d>::perturbation() = std::atof(argv[3]);
#endif
    FILE* fi = std::fopen(argv[1], "rb");
    FILE* fo = std::fopen(argv[2], "wb");
    if (!fi || !fo) { std::fprintf(stderr, "ref_peac: cannot open files\n"); return 2; }
    int32_t hdr[4];
    float cam[5];
    if (std::fread(hdr, 4, 4, fi) != 4 || hdr[0] != 0x50454143 || std::fread(cam, 4, 5, fi) != 5) return 3;
    const int w = hdr[1], h = hdr[2], nframes = hdr[3];
    g_arena_cap = (size_t)1 << 30;
    g_arena = (char*)std::malloc(g_arena_cap);
    cv::Mat K(3, 3, CV_32FC1);
    for (int i = 0; i < 9; ++i) K.at<float>(i / 3, i % 3) = 0.f;
    K.at<float>(0, 0) = cam[0]; K.at<float>(1, 1) = cam[1]; K.at<float>(0, 2) = cam[2]; K.at<float>(1, 2) = cam[3]; K.at<float>(2, 2) = 1.f;
    std::vector<uint16_t> frame((size_t)w * h);
    for (int f = 0; f < nframes; ++f) {
        if (std::fread(frame.data(), 2, frame.size(), fi) != frame.size()) return 4;
        g_arena_off = 0;
        cv::Mat depth(h, w, CV_16UC1, frame.data(), (size_t)w * 2);
        {
            PlaneDetection pd;  // one per Frame, as in the reference (Frame.h:372)
            if (!pd.readDepthImage(depth, K, cam[4])) return 5;                 // Frame.cc:2107
            // the initial blocks exactly as PlaneFitter::initGraph builds them (AHCPlaneFitter.hpp:798-805)
            const int Nh = h / pd.plane_filter.windowHeight, Nw = w / pd.plane_filter.windowWidth;
            int32_t nb = Nh * Nw;
            std::fwrite(&nb, 4, 1, fo);
            for (int i = 0; i < Nh; ++i)
                for (int j = 0; j < Nw; ++j) {
                    ahc::PlaneSeg p(pd.cloud, i * Nw + j, i * pd.plane_filter.windowHeight, j * pd.plane_filter.windowWidth, w, h,
                                    pd.plane_filter.windowWidth, pd.plane_filter.windowHeight, pd.plane_filter.params);
                    int32_t iv[2] = {p.N, p.nouse ? 1 : 0};
                    double dv[8] = {p.center[0], p.center[1], p.center[2], p.normal[0], p.normal[1], p.normal[2], p.mse, p.curvature};
                    if (p.N < 4) for (int k = 0; k < 6; ++k) dv[k] = 0.0;  // centre / normal are uninitialised in the reference then
                    std::fwrite(iv, 4, 2, fo);
                    std::fwrite(dv, 8, 8, fo);
                }
            pd.runPlaneDetection(h, w);                                         // Frame.cc:2108
            int32_t np = pd.plane_num_;
            std::fwrite(&np, 4, 1, fo);
            for (int i = 0; i < np; ++i) {
                const ahc::PlaneSeg& p = *pd.plane_filter.extractedPlanes[i];
                int32_t iv[2] = {p.N, (int32_t)pd.plane_vertices_[i].size()};
                double dv[7] = {p.normal[0], p.normal[1], p.normal[2], p.center[0], p.center[1], p.center[2], p.mse};
                std::fwrite(iv, 4, 2, fo);
                std::fwrite(dv, 8, 7, fo);
            }
            for (int y = 0; y < h; ++y) std::fwrite(pd.plane_filter.membershipImg.ptr<int>(y), 4, (size_t)w, fo);
        }
    }
    std::fclose(fi);
    std::fclose(fo);
    return 0;
}
