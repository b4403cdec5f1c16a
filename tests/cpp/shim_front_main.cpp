// Test driver: the reference-facing C++ bridges for lines, planes and brute-force matching
// (shim/LineExtractorGPU.h, shim/PlaneExtractor.h, shim/MatcherGPU.h -> C ABI -> CUDA), called the way
// src/LineExtractor.cpp:329-380, src/Frame.cc:1094-1096, src/Frame.cc:2104-2130 and src/LSDmatcher.cpp:803-826 do.
//   in.bin : int32 magic, w, h ; gray w*h u8 ; depth w*h u16 ; float fx, fy, cx, cy, factor
//   out.bin: int32 nl ; nl x 68 B keylines ; nl x 32 B LBD ; nl x 3 doubles ; nl x 32 B LBD (recomputed through computeLBD) ;
//            int32 np ; np x {3 normal, 3 center doubles, int32 N, int32 n_vertices, first-vertex xyz (3 doubles)} ; w*h int32 membership ;
//            nl int32 matchNNR(desc, desc reversed rows, 0.95) ;
//            nl int32 LineWindowMatcher::search(mode 1) of the frame's own lines against itself (radius 15, TH 0.96) ;
//            int32 nn ; nn x 3 doubles LPVO normals ; nn floats ; nn x 2 int32 pixels
#include <cstdint>
#include <cstdio>
#include <vector>

#include "keyline_standin.hpp"
#include "LineExtractorGPU.h"
#include "MatcherGPU.h"
#include "PlaneExtractor.h"
#include "WindowedMatcherGPU.h"

int main(int argc, char** argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: shim_front in.bin out.bin\n"); return 2; }
    FILE* fi = std::fopen(argv[1], "rb");
    FILE* fo = std::fopen(argv[2], "wb");
    if (!fi || !fo) return 2;
    int32_t hdr[3];
    if (std::fread(hdr, 4, 3, fi) != 3 || hdr[0] != 0x46524e54) return 3;
    const int w = hdr[1], h = hdr[2];
    std::vector<uint8_t> gray((size_t)w * h);
    std::vector<uint16_t> depth((size_t)w * h);
    float cam[5];
    if (std::fread(gray.data(), 1, gray.size(), fi) != gray.size() || std::fread(depth.data(), 2, depth.size(), fi) != depth.size() ||
        std::fread(cam, 4, 5, fi) != 5) return 4;
    std::fclose(fi);

    // ---- lines: LINEextractor::operator() and the LBD recomputation of Frame::cullingLine ----
    cv::Mat img(h, w, CV_8UC1, gray.data(), (size_t)w), ldesc, ldesc2;
    std::vector<cv::line_descriptor::KeyLine> keylines;
    std::vector<Eigen::Vector3d> lineVec2d;
    hvo_shim::LineFrontEnd lines(1, 1.2f, 200, 0.125);
    lines.extract(img, keylines, ldesc, lineVec2d);
    lines.computeLBD(img, keylines, ldesc2);
    int32_t nl = (int32_t)keylines.size();
    std::fwrite(&nl, 4, 1, fo);
    if (nl) {
        std::fwrite(keylines.data(), 68, nl, fo);
        for (int i = 0; i < nl; ++i) std::fwrite(ldesc.ptr(i), 1, 32, fo);
        for (int i = 0; i < nl; ++i) std::fwrite(lineVec2d[i].v, 8, 3, fo);
        for (int i = 0; i < nl; ++i) std::fwrite(ldesc2.ptr(i), 1, 32, fo);
    }

    // ---- planes: Frame::ComputePlanes' use of PlaneDetection ----
    cv::Mat dimg(h, w, CV_16U, depth.data(), (size_t)w * 2), K(3, 3, CV_32F);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) K.at<float>(i, j) = 0.f;
    K.at<float>(0, 0) = cam[0]; K.at<float>(1, 1) = cam[1]; K.at<float>(0, 2) = cam[2]; K.at<float>(1, 2) = cam[3]; K.at<float>(2, 2) = 1.f;
    PlaneDetection planeDetector;
    planeDetector.readColorImage(img);
    if (!planeDetector.readDepthImage(dimg, K, cam[4])) return 5;
    planeDetector.runPlaneDetection(h, w);
    int32_t np = planeDetector.plane_num_;
    std::fwrite(&np, 4, 1, fo);
    for (int i = 0; i < np; ++i) {
        auto pl = planeDetector.plane_filter.extractedPlanes[i];
        std::fwrite(pl->normal, 8, 3, fo);
        std::fwrite(pl->center, 8, 3, fo);
        int32_t N = pl->N, nv = (int32_t)planeDetector.plane_vertices_[i].size();
        std::fwrite(&N, 4, 1, fo);
        std::fwrite(&nv, 4, 1, fo);
        const int j = planeDetector.plane_vertices_[i][0];
        double xyz[3] = {planeDetector.cloud.vertices[j][0], planeDetector.cloud.vertices[j][1], planeDetector.cloud.vertices[j][2]};
        std::fwrite(xyz, 8, 3, fo);
    }
    std::fwrite(planeDetector.membership().data(), 4, planeDetector.membership().size(), fo);

    // ---- brute-force matching: LSDmatcher::matchNNR ----
    if (nl) {
        cv::Mat rev(nl, 32, CV_8U);
        for (int i = 0; i < nl; ++i) std::memcpy(rev.ptr(i), ldesc.ptr(nl - 1 - i), 32);
        hvo_shim::BruteForceMatcher bf;
        std::vector<int> m12;
        bf.matchNNR(ldesc, rev, 0.95f, m12);
        std::fwrite(m12.data(), 4, m12.size(), fo);
        if (hvo_shim::DescriptorDistance(ldesc, ldesc) != 0) return 6;
    }
    // ---- windowed line matcher: LSDmatcher::SearchByProjection(CurrentFrame, LastFrame, th) with the frame matched against itself ----
    if (nl) {
        hvo_shim::LineWindowMatcher lw;
        std::vector<uint8_t> d((size_t)nl * 32);
        std::vector<double> fn((size_t)nl * 3);
        std::vector<hvo_lproj_query> q(nl);
        for (int i = 0; i < nl; ++i) {
            std::memcpy(&d[(size_t)i * 32], ldesc.ptr(i), 32);
            for (int k = 0; k < 3; ++k) fn[(size_t)i * 3 + k] = lineVec2d[i].v[k];
            const cv::line_descriptor::KeyLine& k = keylines[i];
            hvo_lproj_query& qi = q[i];
            std::memset(&qi, 0, sizeof(qi));
            qi.x1 = k.startPointX; qi.y1 = k.startPointY; qi.x2 = k.endPointX; qi.y2 = k.endPointY;
            qi.r = 15.f; qi.cos_th = 0.96f;
            qi.dir[0] = (double)(k.ePointInOctaveX - k.sPointInOctaveX); qi.dir[1] = (double)(k.ePointInOctaveY - k.sPointInOctaveY);
            qi.length = k.lineLength; qi.claims = 1;
        }
        if (!lw.setFrame(reinterpret_cast<const hvo_keyline*>(keylines.data()), fn.data(), d.data(), nullptr, nl, 0.f, 0.f, (float)w, (float)h)) return 7;
        std::vector<int32_t> idx;
        lw.search(q, d, nullptr, 1, 0.95f, idx);
        std::fwrite(idx.data(), 4, idx.size(), fo);
    }

    // ---- Manhattan::computeNormalsLPVO ----
    {
        hvo_shim::LpvoNormals lpvo(cam[0], cam[1], cam[2], cam[3], cam[4], w, h);
        std::vector<double> nrm; std::vector<float> dn; std::vector<int32_t> px;
        int32_t nn = lpvo.compute(depth.data(), nrm, dn, px);
        std::fwrite(&nn, 4, 1, fo);
        std::fwrite(nrm.data(), 8, nrm.size(), fo);
        std::fwrite(dn.data(), 4, dn.size(), fo);
        std::fwrite(px.data(), 4, px.size(), fo);
    }
    std::fclose(fo);
    return 0;
}
