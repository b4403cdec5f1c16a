// TEST INFRASTRUCTURE — driver for the reference's own line front-end, compiled from the sources where they lie:
//   Thirdparty/line_descriptor/src/LSDDetector_custom.cpp             whole file, unmodified (LSD wrapper: KeyLine fields)
//   Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp       the LBD functions, extracted at build time by
//                                                                     oracle/extract_ref.py into oracle/_ref/gen/lbd.inc
//                                                                     (:42-116, 206-259, 302-412, 523-687, 1026-1372; the rest
//                                                                     of that file is the EDLines detector, never called)
//   src/LineExtractor.cpp:329-380   LINEextractor::operator()         extracted (gen/line_extractor.inc)
//   include/auxiliar.h:47-52        sort_lines_by_response            extracted (gen/auxiliar.inc)
//   src/Frame.cc:952-1203           Frame::cullingLine, PointLineDistance, TwoLineAngle, MergeTwoLines   extracted (gen/frame_cull.inc)
// against stand-ins for OpenCV (oracle/cvshim: GaussianBlur / Sobel / LineSegmentDetector / clipLine forward to restatements
// that are pinned bit-exactly to cv2 4.13.0) and Eigen (oracle/eigenshim).  The reference links the un-vendored opencv_contrib
// line_descriptor module (src/LineExtractor.cpp:2, 340, 361); its in-tree vendored copy (Thirdparty/line_descriptor, *_custom)
// stands in for it here: LSDDetectorC for LSDDetector, the vendored BinaryDescriptor for the contrib one.
//
//   ref_lines <in.bin> <out.bin>
//   in : int32 {magic 0x4c494e45, w, h, nframes, nfeatures, cull}, frames (w*h u8 each)
//   out: per frame: int32 n1; n1 x KeyLine (68 B); n1 x 32 B LBD; n1 x 72 float LBD (before binarisation); n1 x 3 double line functions
//                   (= LINEextractor::operator(), LineExtractor.cpp:329-380)
//        if cull:   int32 n2; n2 x KeyLine; n2 x 32 B LBD; n2 x 3 double      (= Frame::cullingLine(im, 5, 2.5, 15, 30), Frame.cc:936)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "precomp_custom.hpp"   // the vendored line_descriptor's own umbrella header (includes descriptor_custom.hpp)
#include <Eigen/Core>

// ---- stubs for the members of the vendored classes that are declared but not on the path (EDLines detector, file I/O) ----
namespace cv { namespace line_descriptor {
BinaryDescriptor::EDLineDetector::EDLineDetector() {}
BinaryDescriptor::EDLineDetector::~EDLineDetector() {}
void BinaryDescriptor::Params::read(const FileNode&) {}
void BinaryDescriptor::Params::write(FileStorage&) const {}
void BinaryDescriptor::operator()(InputArray, InputArray, std::vector<KeyLine>&, OutputArray, bool, bool) const { std::abort(); }
void BinaryDescriptor::detectImpl(const Mat&, std::vector<KeyLine>&, const Mat&) const { std::abort(); }
// the contrib names the reference's own code uses
struct LSDDetector : public LSDDetectorC {
    static Ptr<LSDDetector> createLSDDetector() { return Ptr<LSDDetector>(new LSDDetector()); }
};
}}  // namespace cv::line_descriptor

using namespace std;
using namespace cv;
using namespace cv::line_descriptor;
using namespace Eigen;

#include "gen/auxiliar.inc"

// ---- stand-ins for the two reference classes: exactly the members the extracted functions touch ----
namespace ORB_SLAM2 {
class LINEextractor {   // include/LineExtractor.h:186-262
public:
    LINEextractor(int _numOctaves, float _scale, unsigned int _nLSDFeature, double _min_line_length)
        : numOctaves(_numOctaves), scale(_scale), nLSDFeature(_nLSDFeature), min_line_length(_min_line_length) {}
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<line_descriptor::KeyLine>& keylines, cv::OutputArray descriptors,
                    std::vector<Eigen::Vector3d>& lineVec2d);
    int numOctaves;
    float scale;
    unsigned int nLSDFeature;
    double min_line_length;
};
class Frame {           // include/Frame.h:164-167, 176, 274-276, 306
public:
    void cullingLine(const cv::Mat& imGray, const double dis, const double angle, const double endpoint_dis, const double min_len_pow);
    double PointLineDistance(Eigen::Vector4d line, cv::Point2f point);
    double TwoLineAngle(Eigen::Vector3d line1, Eigen::Vector3d line2);
    Eigen::Vector4f MergeTwoLines(const Eigen::Vector4f& line1, const Eigen::Vector4f& line2);
    vector<vector<int>> robust_Line;
    Mat mLdesc;
    vector<KeyLine> mvKeylinesUn;
    vector<Vector3d> mvKeyLineFunctions;
};
#include "gen/line_extractor.inc"
#include "gen/frame_cull.inc"
}  // namespace ORB_SLAM2

static_assert(sizeof(cv::line_descriptor::KeyLine) == 68, "KeyLine layout");

static void put_lines(FILE* fo, const vector<KeyLine>& kl, const Mat& desc, const Mat* fdesc, const vector<Vector3d>& lv) {
    int32_t n = (int32_t)kl.size();
    std::fwrite(&n, 4, 1, fo);
    if (!n) return;
    std::fwrite(kl.data(), sizeof(KeyLine), kl.size(), fo);
    for (int i = 0; i < n; ++i) std::fwrite(desc.ptr(i), 1, 32, fo);
    if (fdesc) for (int i = 0; i < n; ++i) std::fwrite(fdesc->ptr(i), 4, 72, fo);
    for (int i = 0; i < n; ++i) std::fwrite(lv[i].data(), 8, 3, fo);
}

int main(int argc, char** argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: ref_lines in.bin out.bin\n"); return 2; }
    FILE* fi = std::fopen(argv[1], "rb");
    FILE* fo = std::fopen(argv[2], "wb");
    if (!fi || !fo) { std::fprintf(stderr, "ref_lines: cannot open files\n"); return 2; }
    int32_t hdr[6];
    if (std::fread(hdr, 4, 6, fi) != 6 || hdr[0] != 0x4c494e45) return 3;
    const int w = hdr[1], h = hdr[2], nframes = hdr[3], nfeat = hdr[4], cull = hdr[5];
    // Tracking.cc:132: new LINEextractor(nLevels = 1, fScaleFactor = 1.2, nFeatures, min_line_length)
    ORB_SLAM2::LINEextractor ex(1, 1.2f, (unsigned)nfeat, 0.125);
    std::vector<uint8_t> frame((size_t)w * h);
    for (int f = 0; f < nframes; ++f) {
        if (std::fread(frame.data(), 1, frame.size(), fi) != frame.size()) return 4;
        cv::Mat img(h, w, CV_8UC1, frame.data(), (size_t)w), mask, desc, fdesc;
        ORB_SLAM2::Frame F;
        ex(img, mask, F.mvKeylinesUn, F.mLdesc, F.mvKeyLineFunctions);                               // Frame.cc:903
        if (!F.mvKeylinesUn.empty()) {
            std::vector<KeyLine> tmp = F.mvKeylinesUn;
            BinaryDescriptor::createBinaryDescriptor()->compute(img, tmp, fdesc, true);              // the float LBD, for diagnosis
        }
        put_lines(fo, F.mvKeylinesUn, F.mLdesc, &fdesc, F.mvKeyLineFunctions);
        if (cull) {
            if (!F.mvKeylinesUn.empty()) F.cullingLine(img, 5, 2.5, 15, 30);                          // Frame.cc:936
            put_lines(fo, F.mvKeylinesUn, F.mLdesc, nullptr, F.mvKeyLineFunctions);
        }
    }
    std::fclose(fi);
    std::fclose(fo);
    return 0;
}
