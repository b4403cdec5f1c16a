// Test driver: the reference-facing C++ shim (shim/ORBextractor.h -> C ABI -> CUDA), same file protocol as
// oracle/ref_orb_main.cpp so the two binaries can be diffed byte for byte.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "ORBextractor.h"

int main(int argc, char** argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: shim_orb in.bin out.bin\n"); return 2; }
    FILE* fi = std::fopen(argv[1], "rb");
    FILE* fo = std::fopen(argv[2], "wb");
    if (!fi || !fo) return 2;
    int32_t hdr[8];
    float scale;
    if (std::fread(hdr, 4, 8, fi) != 8 || hdr[0] != 0x4f524231 || std::fread(&scale, 4, 1, fi) != 1) return 3;
    const int w = hdr[1], h = hdr[2], nframes = hdr[3];
    ORB_SLAM2::ORBextractor ex(hdr[4], scale, hdr[5], hdr[6], hdr[7]);
    std::vector<uint8_t> frame((size_t)w * h);
    for (int f = 0; f < nframes; ++f) {
        if (std::fread(frame.data(), 1, frame.size(), fi) != frame.size()) return 4;
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
