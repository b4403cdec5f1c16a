// TEST INFRASTRUCTURE — driver for the reference's own ORBextractor.cc (compiled unmodified from
// /root/reference/src against oracle/cvshim).  Built into oracle/_ref/ref_orb by oracle/Makefile.
//
//   ref_orb <in.bin> <out.bin>
//   in : int32 {magic 0x4f524231, w, h, nframes, nfeatures, nlevels, iniTh, minTh}, float scale, frames (w*h u8 each)
//   out: per frame int32 n, n x 28-byte cv::KeyPoint, n x 32-byte descriptors
//
// Determinism: the reference sorts (size, ExtractorNode*) pairs (ORBextractor.cc:682), so ties between
// equally-sized nodes follow heap addresses.  This driver replaces the global operator new so that
// std::list<ExtractorNode> nodes come from a monotonic arena: "later created" == "higher address",
// which is the tie rule the oracle restatement documents.  All other allocations use malloc.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <list>
#include <new>
#include <vector>

#include "ORBextractor.h"

static const size_t kNodeBytes = sizeof(std::_List_node<ORB_SLAM2::ExtractorNode>);
static char* g_arena = nullptr;
static size_t g_arena_cap = 0, g_arena_off = 0;

void* operator new(size_t n) {
    if (n == kNodeBytes && g_arena && g_arena_off + n <= g_arena_cap) {
        void* p = g_arena + g_arena_off;
        g_arena_off += (n + 15) & ~(size_t)15;
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
    if (argc != 3) { std::fprintf(stderr, "usage: ref_orb in.bin out.bin\n"); return 2; }
    FILE* fi = std::fopen(argv[1], "rb");
    FILE* fo = std::fopen(argv[2], "wb");
    if (!fi || !fo) { std::fprintf(stderr, "ref_orb: cannot open files\n"); return 2; }
    int32_t hdr[8];
    float scale;
    if (std::fread(hdr, 4, 8, fi) != 8 || hdr[0] != 0x4f524231 || std::fread(&scale, 4, 1, fi) != 1) return 3;
    const int w = hdr[1], h = hdr[2], nframes = hdr[3];
    g_arena_cap = (size_t)256 << 20;
    g_arena = (char*)std::malloc(g_arena_cap);
    ORB_SLAM2::ORBextractor ex(hdr[4], scale, hdr[5], hdr[6], hdr[7]);
    std::vector<uint8_t> frame((size_t)w * h);
    for (int f = 0; f < nframes; ++f) {
        if (std::fread(frame.data(), 1, frame.size(), fi) != frame.size()) return 4;
        g_arena_off = 0;  // no list node outlives DistributeOctTree
        cv::Mat img(h, w, CV_8UC1, frame.data(), (size_t)w), desc;
        std::vector<cv::KeyPoint> kps;
        ex(img, cv::Mat(), kps, desc);
        int32_t n = (int32_t)kps.size();
        std::fwrite(&n, 4, 1, fo);
        if (n) {
            std::fwrite(kps.data(), sizeof(cv::KeyPoint), kps.size(), fo);
            for (int i = 0; i < n; ++i) std::fwrite(desc.ptr(i), 1, 32, fo);
        }
    }
    std::fclose(fi);
    std::fclose(fo);
    return 0;
}
