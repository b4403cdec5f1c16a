// TEST INFRASTRUCTURE — driver for the reference's own Manhattan::computeNormalsLPVO, compiled from the source where it lies: the
// function and removeMatRow / removeMatCol (src/Manhattan.cpp:237-393, 395-493) are pulled out at build time by oracle/extract_ref.py
// (oracle/_ref/gen/manhattan_lpvo.inc, never committed) and compiled against the OpenCV stand-in, with a stand-in Manhattan class that
// carries the four members the function reads (include/Manhattan.h:86-92; set as the constructor does, src/Manhattan.cpp:10-18).
//
// Two things are fixed, both documented in DESIGN.md §2 and SURVEY App. B (D6 is outside the parity gate for the first one):
//  * the reference's live caller (src/Frame.cc:222) hands the raw CV_16U depth to this function, which reads it with at<float>; the
//    driver passes what the function's signature asks for, imDepth.convertTo(CV_32F, mDepthMapFactor);
//  * removeMatRow / removeMatCol have two bodies: a cv::Rect one (#ifdef USE_CV_RECT) and a memcpy one that sizes its rows with
//    sizeof(float) although the integral images are CV_64F (it moves half of every row).  USE_CV_RECT is defined in src/Frame.cc:32 only,
//    so Manhattan.cpp as built takes the memcpy body; the driver compiles the cv::Rect body (-DUSE_CV_RECT), which does what the
//    function's own comment says ("Delete row and column 0").  `ref_lpvo ... asbuilt` runs the memcpy body instead, to show the difference.
// cv::integral / cv::normalize forward to the oracle's restatements, which are pinned to cv2 4.13.0 (tests/test_lpvo.py).
//
//   ref_lpvo <in.bin> <out.bin>     in: int32 W, H; float factor, fx, fy, cx, cy; uint16 depth[H*W]
//                                   out: int32 n; double normals[n][3]; float depth[n]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <opencv2/core/core.hpp>

extern "C" void orc_integral_f32(const float* src, int W, int H, double* out);
extern "C" void orc_normalize3(const double* v, double* out);

#ifndef CV_32FC3
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#endif
namespace cv {
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;
typedef Vec<int, 2> Vec2i;
template <typename T, int N>
static inline Vec<T, N> operator-(const Vec<T, N>& a, const Vec<T, N>& b) {
    Vec<T, N> r;
    for (int i = 0; i < N; ++i) r[i] = a[i] - b[i];
    return r;
}
// cv::integral(src CV_32F, sum): (H + 1) x (W + 1) CV_64F with a zero first row and column
static inline void integral(const Mat& src, Mat& sum) {
    assert(src.type() == CV_32FC1 && src.isContinuous());
    std::vector<double> body((size_t)src.rows * src.cols);
    orc_integral_f32(src.ptr<float>(), src.cols, src.rows, body.data());
    sum = Mat::zeros(src.rows + 1, src.cols + 1, CV_64FC1);
    for (int y = 0; y < src.rows; ++y) std::memcpy(&sum.at<double>(y + 1, 1), &body[(size_t)y * src.cols], (size_t)src.cols * sizeof(double));
}
// cv::normalize(src, dst) of a 3 x 1 CV_64F vector (NORM_L2, alpha = 1)
static inline void normalize(const Mat& src, Mat& dst) {
    assert(src.type() == CV_64FC1 && src.rows * src.cols == 3);
    Mat out(src.rows, src.cols, CV_64FC1);
    orc_normalize3(src.ptr<double>(), out.ptr<double>());
    dst = out;
}
}  // namespace cv

using namespace std;
using namespace cv;

namespace ORB_SLAM2 {
class Manhattan {   // include/Manhattan.h:31, 56-57, 86-92
public:
    void computeNormalsLPVO(const cv::Mat& im_depth_resized, const cv::Mat& K, std::vector<cv::Mat>& pt_normals, std::vector<float>& depth_normals);
    void removeMatCol(cv::Mat& matIn, int col);
    void removeMatRow(cv::Mat& matIn, int row);
    float mFx, mFy, mCx, mCy, mInvFx, mInvFy;
};
#include "gen/manhattan_lpvo.inc"
}  // namespace ORB_SLAM2

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: ref_lpvo in.bin out.bin\n"); return 2; }
    FILE* fi = std::fopen(argv[1], "rb");
    FILE* fo = std::fopen(argv[2], "wb");
    if (!fi || !fo) return 2;
    int32_t wh[2]; float p[5];
    if (std::fread(wh, 4, 2, fi) != 2 || std::fread(p, 4, 5, fi) != 5) return 3;
    const int W = wh[0], H = wh[1];
    std::vector<uint16_t> d16((size_t)W * H);
    if (std::fread(d16.data(), 2, d16.size(), fi) != d16.size()) return 3;
    cv::Mat depth(H, W, CV_32FC1);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) depth.at<float>(y, x) = (float)d16[(size_t)y * W + x] * p[0];   // convertTo(CV_32F, factor): a float product
    ORB_SLAM2::Manhattan m;
    m.mFx = p[1]; m.mFy = p[2]; m.mCx = p[3]; m.mCy = p[4];
    m.mInvFx = 1.0f / m.mFx; m.mInvFy = 1.0f / m.mFy;                                              // src/Manhattan.cpp:13-18
    cv::Mat K;
    std::vector<cv::Mat> normals;
    std::vector<float> z;
    m.computeNormalsLPVO(depth, K, normals, z);
    const int32_t n = (int32_t)normals.size();
    std::fwrite(&n, 4, 1, fo);
    for (const cv::Mat& v : normals) std::fwrite(v.ptr<double>(), 8, 3, fo);
    std::fwrite(z.data(), 4, z.size(), fo);
    std::fclose(fi);
    std::fclose(fo);
    return 0;
}
