// TEST INFRASTRUCTURE — minimal stand-in for the slice of the OpenCV C++ API that the reference sources compiled into
// oracle/_ref/ use (src/ORBextractor.cc, src/PlaneExtractor.cpp + include/peac, Thirdparty/line_descriptor, and the
// functions of src/Frame.cc / src/LineExtractor.cpp / src/*matcher* that oracle/extract_ref.py pulls out at build time), so
// those sources can be compiled UNMODIFIED from /root/reference (OpenCV's C++ headers/libs are absent from this image).
// The image primitives forward to oracle/cvprims.hpp and to the oracle's LSD / clipLine restatements, each of which is pinned
// bit-exactly against cv2 4.13.0 (tests/test_oracle_prims.py, tests/test_lsd.py, tests/test_cull.py).  Not a general OpenCV
// replacement.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../cvprims.hpp"

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
// OpenCV's type code: depth in the low 3 bits, (channels - 1) << 3 above
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
static inline size_t cvshim_elem_size(int type) {
    static const size_t depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
    return depth_bytes[type & 7] * (size_t)((type >> 3) + 1);
}

static inline int cvRound(double v) { return cvp::cv_round(v); }
static inline int cvRound(float v) { return cvp::cv_round(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { return cvp::cv_floor(v); }
static inline int cvFloor(float v) { return cvp::cv_floor(v); }
static inline int cvCeil(double v) { return cvp::cv_ceil(v); }
static inline int cvCeil(float v) { return cvp::cv_ceil(v); }

#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_OUT
#define CV_IN_OUT
#define CV_WRAP
#define CV_GRAY2BGR 8
#define CV_BGR2GRAY 6

// restatements living in oracle/_build/liboracle.so (pinned to cv2 4.13.0)
extern "C" int orc_lsd_detect(const uint8_t* gray, int w, int h, float* segments4, int cap, uint8_t* scaled_out, int* sw, int* sh);
extern "C" int orc_clip_line(int w, int h, long long* pts4);
extern "C" void orc_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx2, int32_t* dist2);

namespace cv {

using std::max;
using std::min;
using std::abs;
using std::sqrt;
using std::pow;
using std::exp;
using std::log;
using std::swap;

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16, BORDER_DEFAULT = 4 };
enum { COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8 };
enum { NORM_HAMMING = 6 };
enum { INTER_LINEAR = 1 };

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
};
template <typename T>
static inline Point_<T>& operator*=(Point_<T>& a, float b) {
    a.x = (T)(a.x * b);
    a.y = (T)(a.y * b);
    return a;
}
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> static inline Point_<T>& operator+=(Point_<T>& a, const Point_<T>& b) { a.x += b.x; a.y += b.y; return a; }
// cv::Point_<T> * double: saturate_cast<T>(a.x * b), computed in double
template <typename T> static inline Point_<T> operator*(const Point_<T>& a, double b) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
static inline bool operator==(const Size& a, const Size& b) { return a.width == b.width && a.height == b.height; }
static inline bool operator!=(const Size& a, const Size& b) { return !(a == b); }
struct Rect {
    int x, y, width, height;
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

struct Range {
    int start, end;
    Range(int s, int e) : start(s), end(e) {}
};
template <typename T, int N>
struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    explicit Vec(const T* p) { for (int i = 0; i < N; ++i) val[i] = p[i]; }
    Vec(T a, T b) { static_assert(N == 2, "Vec2"); val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { static_assert(N == 3, "Vec3"); val[0] = a; val[1] = b; val[2] = c; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<double, 2> Vec2d;
typedef Vec<float, 4> Vec4f;
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
};
struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4028235e38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};
// cv::FileStorage / cv::FileNode: never opened on the path (the harnesses build their data in memory); the members exist so that the
// YAML save / load members of DBoW2::TemplatedVocabulary (virtual, hence instantiated) compile.  isOpened() is false: they throw.
struct FileNode {
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator float() const { return 0.f; }
    operator double() const { return 0.; }
    operator std::string() const { return std::string(); }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const char*, int) {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <typename T> static inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
class Algorithm {
public:
    virtual ~Algorithm() {}
    virtual void read(const FileNode&) {}
    virtual void write(FileStorage&) const {}
};
template <typename T>
struct Ptr : public std::shared_ptr<T> {
    Ptr() {}
    Ptr(T* p) : std::shared_ptr<T>(p) {}
    template <typename U> Ptr(const Ptr<U>& o) : std::shared_ptr<T>(o) {}
    operator T*() const { return this->get(); }
};
template <typename T> struct cvshim_type;
template <> struct cvshim_type<uchar> { enum { value = CV_8UC1 }; };
template <> struct cvshim_type<int> { enum { value = CV_32SC1 }; };
template <> struct cvshim_type<float> { enum { value = CV_32FC1 }; };
template <> struct cvshim_type<double> { enum { value = CV_64FC1 }; };
static inline long long getTickCount() { return 0; }
static inline double getTickFrequency() { return 1.0; }

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
             int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct MatStep {
    size_t v;
    MatStep() : v(0) {}
    operator size_t() const { return v; }
};

struct MatZerosExpr { int rows, cols, type; };
struct MatOnesExpr { int rows, cols, type; };

struct MatTExpr;
class Mat {
public:
    int rows, cols;
    uchar* data;
    MatStep step;
    Mat() : rows(0), cols(0), data(nullptr) {}
    Mat(Size s, int type) : rows(0), cols(0), data(nullptr) { create(s.height, s.width, type); }
    Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr) { create(r, c, type); }
    // external (non-owning) buffer
    Mat(int r, int c, int type, void* ext, size_t st) : rows(r), cols(c), data((uchar*)ext), type_(type) {
        step.v = st;
    }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_) return;
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c * cvshim_elem_size(type));
        rows = r;
        cols = c;
        type_ = type;
        data = buf_->data();
        step.v = (size_t)c * cvshim_elem_size(type);
    }
    void release() { buf_.reset(); rows = cols = 0; data = nullptr; step.v = 0; }
    Mat(const MatZerosExpr& z) : rows(0), cols(0), data(nullptr) { create(z.rows, z.cols, z.type); }   // cv::Mat m = cv::Mat::zeros(...): fresh, zero-filled
    Mat& operator=(const MatZerosExpr& z) {
        create(z.rows, z.cols, z.type);  // no-op for a view of matching size: zeros are written in place
        for (int y = 0; y < rows; ++y) std::memset(data + (size_t)y * step.v, 0, (size_t)cols);
        return *this;
    }
    static MatZerosExpr zeros(int r, int c, int type) { return MatZerosExpr{r, c, type}; }
    static MatOnesExpr ones(int r, int c, int type) { return MatOnesExpr{r, c, type}; }
    Mat& operator=(const MatOnesExpr& z) {  // 8U only (the one use: ahc::PlaneFitter::getGraph)
        assert(z.type == CV_8UC1);
        create(z.rows, z.cols, z.type);
        for (int y = 0; y < rows; ++y) std::memset(data + (size_t)y * step.v, 1, (size_t)cols);
        return *this;
    }
    Mat operator()(const Range& rr, const Range& cr) const { return (*this)(Rect(cr.start, rr.start, cr.end - cr.start, rr.end - rr.start)); }
    Size size() const { return Size(cols, rows); }
    Mat col(int x) const { return (*this)(Rect(x, 0, 1, rows)); }
    MatTExpr t() const;  // CV_32F only (pose blocks); an expression, as in OpenCV: see MatTExpr below
    Mat row(int y) const { return (*this)(Rect(0, y, cols, 1)); }
    void copyTo(Mat& dst) const { dst = clone(); }
    void copyTo(Mat&& view) const {   // src(rect).copyTo(dst(rect)): the destination is a view of the same size and type, written in place
        assert(view.rows == rows && view.cols == cols && view.type_ == type_);
        for (int y = 0; y < rows; ++y) std::memmove(view.data + (size_t)y * view.step.v, data + (size_t)y * step.v, (size_t)cols * cvshim_elem_size(type_));
    }
    void copyTo(const class _OutputArray& dst) const;
    bool isContinuous() const { return step.v == (size_t)cols * cvshim_elem_size(type_); }
    Mat cross(const Mat& o) const;   // CV_32F 3-vectors (LSDmatcher::FrameBFMatchNew)
    double dot(const Mat& o) const {  // vectors (Frame::TwoLineAngle 64F, Frame::isInFrustum 32F); cv::Mat::dot accumulates in double, in order
        if (type_ == CV_32FC1) {
            assert(o.type_ == CV_32FC1 && rows * cols == o.rows * o.cols);
            double r32 = 0;
            for (int i = 0; i < rows * cols; ++i) r32 += (double)at<float>(i) * (double)o.at<float>(i);
            return r32;
        }
        assert(type_ == CV_64FC1 && o.type_ == CV_64FC1 && rows * cols == o.rows * o.cols);
        double r = 0;
        for (int i = 0; i < rows * cols; ++i) r += at<double>(i) * o.at<double>(i);
        return r;
    }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    // setTo for the element types the reference writes: int labels / 8U masks, and Vec3b colours
    Mat& setTo(int v) {
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                if (type_ == CV_32SC1) at<int>(y, x) = v;
                else if (type_ == CV_8UC1) at<uchar>(y, x) = (uchar)v;
                else { assert(!"cvshim: setTo(int) on an unsupported type"); }
            }
        return *this;
    }
    Mat& setTo(const Vec3b& v) {
        assert(type_ == CV_8UC3);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) at<Vec3b>(y, x) = v;
        return *this;
    }
    // linear element index of a continuous matrix (cv::Mat::at<T>(int i0))
    template <typename T> T& at(int i) { return *(T*)(data + (size_t)(i / cols) * step.v + (size_t)(i % cols) * sizeof(T)); }
    template <typename T> const T& at(int i) const { return *(const T*)(data + (size_t)(i / cols) * step.v + (size_t)(i % cols) * sizeof(T)); }
    Mat operator()(const Rect& r) const {
        Mat m(*this);
        m.data = data + (size_t)r.y * step.v + r.x * cvshim_elem_size(type_);
        m.rows = r.height;
        m.cols = r.width;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step.v, data + (size_t)y * step.v, (size_t)cols * cvshim_elem_size(type_));
        return m;
    }
    int type() const { return type_; }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step.v); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step.v); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t step1() const { return step.v; }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step.v + x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step.v + x * sizeof(T)); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step.v; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step.v; }

private:
    int type_ = CV_8UC1;
    std::shared_ptr<std::vector<uchar>> buf_;
};

// cv::Mat_<T>(r, c) << a, b, ... (comma initialiser)
template <typename T>
class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, cvshim_type<T>::value) {}
    struct Init {
        Mat_ m;
        int i;
        Init& operator,(T v) { m.template at<T>(i++) = v; return *this; }
        operator Mat() const { return m; }
        operator Mat_() const { return m; }
    };
    Init operator<<(T v) { Init it{*this, 0}; it.m.template at<T>(it.i++) = v; return it; }
    T& operator()(int y, int x) { return this->template at<T>(y, x); }
    const T& operator()(int y, int x) const { return this->template at<T>(y, x); }
};

// The three cv::MatExpr forms the reference's matchers use on small CV_32F pose blocks: A * B, A + B, -A.
// cv::gemm on CV_32F accumulates in float, in k order (checked against cv2.gemm 4.13.0: 0 mismatches in 2000 random 3x3 * 3x1).
static inline Mat operator*(const Mat& a, const Mat& b) {
    assert(a.type() == CV_32FC1 && b.type() == CV_32FC1 && a.cols == b.rows);
    Mat r(a.rows, b.cols, CV_32FC1);
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < b.cols; ++j) {
            float s = 0.f;
            for (int k = 0; k < a.cols; ++k) s += a.at<float>(i, k) * b.at<float>(k, j);
            r.at<float>(i, j) = s;
        }
    return r;
}
static inline Mat operator+(const Mat& a, const Mat& b) {
    assert(a.type() == CV_32FC1 && b.type() == CV_32FC1 && a.rows == b.rows && a.cols == b.cols);
    Mat r(a.rows, a.cols, CV_32FC1);
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) r.at<float>(i, j) = a.at<float>(i, j) + b.at<float>(i, j);
    return r;
}
static inline Mat operator-(const Mat& a) {
    assert(a.type() == CV_32FC1);
    Mat r(a.rows, a.cols, CV_32FC1);
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) r.at<float>(i, j) = -a.at<float>(i, j);
    return r;
}

class _InputArray {
public:
    _InputArray(const Mat& m) : m_(&m) {}
    bool empty() const { return m_->empty(); }
    Mat getMat() const { return *m_; }
private:
    const Mat* m_;
};
class _OutputArray {
public:
    _OutputArray(Mat& m) : m_(&m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void release() const { m_->release(); }
    Mat getMat() const { return *m_; }
private:
    Mat* m_;
};
inline void Mat::copyTo(const _OutputArray& dst) const {
    Mat c = clone();
    dst.create(rows, cols, type_);
    Mat d = dst.getMat();
    for (int y = 0; y < rows; ++y) std::memcpy(d.data + (size_t)y * d.step.v, c.data + (size_t)y * c.step.v, (size_t)cols * cvshim_elem_size(type_));
}
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

static inline float fastAtan2(float y, float x) { return cvp::fast_atan2_deg(y, x); }

static inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nms) {
    assert(nms);
    (void)nms;
    std::vector<cvp::FastKp> r;
    cvp::fast9_nms(img.data, img.cols, img.rows, img.step.v, threshold, r);
    kps.clear();
    for (const cvp::FastKp& k : r) kps.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1, (float)k.score));
}

static inline void resize(const Mat& src, Mat& dst, Size sz, double, double, int interp) {
    assert(interp == INTER_LINEAR);
    (void)interp;
    dst.create(sz.height, sz.width, CV_8UC1);
    cvp::resize_linear_u8(src.data, src.cols, src.rows, src.step.v, dst.data, dst.cols, dst.rows, dst.step.v);
}

static inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int type) {
    assert((type & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    (void)type;
    dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    // works in place when src is the interior ROI of dst: interior pixels map to themselves
    for (int y = 0; y < dst.rows; ++y) {
        const uchar* s = src.data + (size_t)cvp::reflect101(y - top, src.rows) * src.step.v;
        uchar* d = dst.data + (size_t)y * dst.step.v;
        for (int x = 0; x < dst.cols; ++x) {
            if (y >= top && y < top + src.rows && x >= left && x < left + src.cols && d + x == s + (x - left)) continue;
            d[x] = s[cvp::reflect101(x - left, src.cols)];
        }
    }
}

static inline void GaussianBlur(const Mat& src, Mat& dst, Size k, double sx, double sy = 0, int border = BORDER_DEFAULT) {
    // the two calls the reference makes: 7x7 sigma 2 (ORBextractor.cc:1084) and 5x5 sigma 1 (binary_descriptor_custom.cpp:358)
    const bool is7 = k.width == 7 && k.height == 7 && sx == 2 && (sy == 2 || sy == 0);
    const bool is5 = k.width == 5 && k.height == 5 && sx == 1 && (sy == 1 || sy == 0);
    assert((is7 || is5) && border == BORDER_REFLECT_101 && src.type() == CV_8UC1);
    (void)border; (void)is5;
    Mat tmp(src.rows, src.cols, CV_8UC1);
    if (is7) cvp::gaussian_blur7_s2(src.data, src.cols, src.rows, src.step.v, tmp.data, tmp.step.v);
    else cvp::gaussian_blur5_s1(src.data, src.cols, src.rows, src.step.v, tmp.data, tmp.step.v);
    dst.create(src.rows, src.cols, CV_8UC1);
    for (int y = 0; y < tmp.rows; ++y) std::memcpy(dst.data + (size_t)y * dst.step.v, tmp.data + (size_t)y * tmp.step.v, (size_t)tmp.cols);
}

// cv::Sobel(src 8U, dst, CV_16S, dx, dy, 3): exactly one of (1,0) / (0,1)
static inline void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize) {
    assert(src.type() == CV_8UC1 && ddepth == CV_16SC1 && ksize == 3 && dx + dy == 1 && src.isContinuous());
    (void)ddepth; (void)ksize;
    std::vector<int16_t> gx((size_t)src.rows * src.cols), gy(gx.size());
    cvp::sobel3_s16(src.data, src.cols, src.rows, src.step.v, gx.data(), gy.data());
    dst.create(src.rows, src.cols, CV_16SC1);
    const std::vector<int16_t>& g = dx ? gx : gy;
    for (int y = 0; y < src.rows; ++y) std::memcpy(dst.data + (size_t)y * dst.step.v, g.data() + (size_t)y * src.cols, (size_t)src.cols * 2);
}

// never reached on the reference's path (single octave, grey input); they only have to compile
static inline void pyrDown(const Mat&, Mat&, Size) { std::fprintf(stderr, "cvshim: pyrDown is not on the path (numOctaves == 1)\n"); std::abort(); }
static inline void cvtColor(const Mat& src, Mat& dst, int code) {
    if (code == COLOR_GRAY2BGR) { (void)src; dst = Mat(src.rows, src.cols, CV_8UC3); return; }   // debug drawing target only
    std::fprintf(stderr, "cvshim: cvtColor(BGR2GRAY) is not on the path (grey input)\n"); std::abort();
}
template <typename P> static inline void line(Mat&, P, P, const Scalar&, double = 1, int = 8, int = 0) {}  // debug drawing: no-op

// cv::LineSegmentDetector with the default parameters (what LSDDetector_custom.cpp:149 constructs): forwards to the oracle's
// restatement, which is bit-identical to cv2.createLineSegmentDetector().detect (tests/test_lsd.py, golden + live)
class LineSegmentDetector : public Algorithm {
public:
    void detect(const Mat& img, std::vector<Vec4f>& lines) {
        assert(img.type() == CV_8UC1 && img.isContinuous());
        std::vector<float> seg((size_t)4 * 65536);
        const int n = orc_lsd_detect(img.data, img.cols, img.rows, seg.data(), 65536, nullptr, nullptr, nullptr);
        assert(n <= 65536);
        lines.clear();
        for (int i = 0; i < n; ++i) lines.push_back(Vec4f(&seg[4 * i]));
    }
};
static inline Ptr<LineSegmentDetector> createLineSegmentDetector() { return Ptr<LineSegmentDetector>(new LineSegmentDetector()); }
static inline Ptr<LineSegmentDetector> createLineSegmentDetector(int, double, double, double, double, double, double, int) {
    std::fprintf(stderr, "cvshim: only the default LineSegmentDetector is on the path\n"); std::abort();
}

// cv::LineIterator(img, pt1, pt2): only .count is read (8-connected, after cv::clipLine; Point2f -> Point is cvRound).
// clipLine restatement pinned to cv2.clipLine (tests/test_cull.py).
class LineIterator {
public:
    int count;
    LineIterator(const Mat& img, Point2f p1, Point2f p2) { init(img, cvRound(p1.x), cvRound(p1.y), cvRound(p2.x), cvRound(p2.y)); }
    LineIterator(const Mat& img, Point p1, Point p2) { init(img, p1.x, p1.y, p2.x, p2.y); }
private:
    void init(const Mat& img, long long x1, long long y1, long long x2, long long y2) {
        const long long w = img.cols, h = img.rows;
        if ((unsigned long long)x1 >= (unsigned long long)w || (unsigned long long)x2 >= (unsigned long long)w ||
            (unsigned long long)y1 >= (unsigned long long)h || (unsigned long long)y2 >= (unsigned long long)h) {
            long long pts[4] = {x1, y1, x2, y2};
            if (!orc_clip_line((int)w, (int)h, pts)) { count = 0; return; }
            x1 = pts[0]; y1 = pts[1]; x2 = pts[2]; y2 = pts[3];
        }
        const long long dx = x2 > x1 ? x2 - x1 : x1 - x2, dy = y2 > y1 ? y2 - y1 : y1 - y2;
        count = (int)std::max(dx, dy) + 1;
    }
};


// ---- small CV_32F matrix algebra the tracking-time functions use (Frame::isInFrustum, LSDmatcher::FrameBFMatchNew) ----
// Semantics checked against cv2 4.13.0 where Python exposes the operation (gemm, norm, addWeighted: tests/test_track.py); the
// A.t() is an expression in OpenCV (MatOp_T) and the reference's `-Rcw.t()*tcw` (ORBmatcher.cc:308, 1009, 1366, 1505; LSDmatcher.cpp:577)
// becomes ONE cv::gemm(Rcw, tcw, -1, noArray(), 0, GEMM_1_T).  With a transpose flag cv::gemm leaves its small-matrix path and runs
// GEMMSingleMul<float, double>: products and sums in double, (float)(sum * alpha) at the end (checked against cv2.gemm 4.13.0 with
// GEMM_1_T: 0 mismatches in 5000 random 3x3^T * 3x1, whereas float accumulation differs in 80 % of them;
// tests/test_track.py::test_opencv_float_matrix_rules).  `s * A.t()` assigned to a Mat is transpose + convertTo(alpha): a float product.
struct MatTExpr {
    Mat a;
    double alpha;
    operator Mat() const {
        assert(a.type() == CV_32FC1);
        Mat r(a.cols, a.rows, CV_32FC1);
        const float f = (float)alpha;
        for (int y = 0; y < a.rows; ++y)
            for (int x = 0; x < a.cols; ++x) r.at<float>(x, y) = alpha == 1. ? a.at<float>(y, x) : a.at<float>(y, x) * f;
        return r;
    }
};
inline MatTExpr Mat::t() const { return MatTExpr{*this, 1.}; }
static inline MatTExpr operator-(MatTExpr e) { e.alpha = -e.alpha; return e; }
static inline MatTExpr operator*(double s, MatTExpr e) { e.alpha *= s; return e; }
static inline Mat operator*(const MatTExpr& e, const Mat& b) {
    assert(e.a.type() == CV_32FC1 && b.type() == CV_32FC1 && e.a.rows == b.rows);
    Mat r(e.a.cols, b.cols, CV_32FC1);
    for (int i = 0; i < e.a.cols; ++i)
        for (int j = 0; j < b.cols; ++j) {
            double s = 0.;
            for (int k = 0; k < e.a.rows; ++k) s += (double)e.a.at<float>(k, i) * (double)b.at<float>(k, j);
            r.at<float>(i, j) = (float)(s * e.alpha);
        }
    return r;
}
// A / s is convertTo(A, -1, 1. / s): a float multiplication by (float)(1. / s) (ORBmatcher.cc:306-307, 1007-1008: sRcw / scw)
static inline Mat operator/(const Mat& a, double s) {
    assert(a.type() == CV_32FC1);
    Mat r(a.rows, a.cols, CV_32FC1);
    const float f = (float)(1. / s);
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) r.at<float>(i, j) = a.at<float>(i, j) * f;
    return r;
}

// rest follows OpenCV's sources: Mat::dot / cv::norm accumulate in double in element order, Mat::cross works in float,
// `m /= s` is convertTo(m, -1, 1. / s) i.e. a float multiplication by (float)(1. / s), `s * (A + B)` is addWeighted(A, s, B, s).
// (A * B + C is one cv::gemm call in OpenCV, (float)(double(t) + double(c)): for two floats that equals the float sum.)
static inline Mat operator-(const Mat& a, const Mat& b) {
    assert(a.type() == CV_32FC1 && b.type() == CV_32FC1 && a.rows == b.rows && a.cols == b.cols);
    Mat r(a.rows, a.cols, CV_32FC1);
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) r.at<float>(i, j) = a.at<float>(i, j) - b.at<float>(i, j);
    return r;
}
static inline Mat operator*(double s, const Mat& a) {
    assert(a.type() == CV_32FC1);
    Mat r(a.rows, a.cols, CV_32FC1);
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) r.at<float>(i, j) = a.at<float>(i, j) * (float)s;
    return r;
}
static inline Mat& operator/=(Mat& a, double s) {
    assert(a.type() == CV_32FC1);
    const float f = (float)(1. / s);
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) a.at<float>(i, j) = a.at<float>(i, j) * f;
    return a;
}
static inline double norm(const Mat& a) {
    assert(a.type() == CV_32FC1);
    double s = 0;
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) s += (double)a.at<float>(i, j) * (double)a.at<float>(i, j);
    return std::sqrt(s);
}
static inline double norm(const Mat& a, const Mat& b, int normType) {   // NORM_HAMMING on CV_8U rows (MapLine::ComputeDistinctiveDescriptors)
    assert(normType == NORM_HAMMING && a.type() == CV_8UC1 && b.type() == CV_8UC1 && a.rows * a.cols == b.rows * b.cols);
    (void)normType;
    int d = 0;
    for (int i = 0; i < a.rows * a.cols; ++i) d += __builtin_popcount((unsigned)(a.at<uchar>(i) ^ b.at<uchar>(i)));
    return (double)d;
}
static inline double cvshim_dot32f(const Mat& a, const Mat& b) {
    assert(a.rows * a.cols == b.rows * b.cols);
    double r = 0;
    for (int i = 0; i < a.rows * a.cols; ++i) r += (double)a.at<float>(i) * (double)b.at<float>(i);
    return r;
}
static inline Mat cvshim_cross32f(const Mat& a, const Mat& b) {
    assert(a.type() == CV_32FC1 && b.type() == CV_32FC1 && a.rows * a.cols == 3 && b.rows * b.cols == 3);
    Mat c(a.rows, a.cols, CV_32FC1);
    const float a0 = a.at<float>(0), a1 = a.at<float>(1), a2 = a.at<float>(2), b0 = b.at<float>(0), b1 = b.at<float>(1), b2 = b.at<float>(2);
    c.at<float>(0) = a1 * b2 - a2 * b1;
    c.at<float>(1) = a2 * b0 - a0 * b2;
    c.at<float>(2) = a0 * b1 - a1 * b0;
    return c;
}

inline Mat Mat::cross(const Mat& o) const { return cvshim_cross32f(*this, o); }

// cv::BFMatcher(NORM_HAMMING, false).knnMatch(q, t, matches, k = 2): forwards to the oracle's brute force, which is pinned to
// cv2.BFMatcher (ties: lower train index first; tests/test_match.py)
class BFMatcher {
public:
    BFMatcher(int normType, bool crossCheck = false) { assert(normType == NORM_HAMMING && !crossCheck); (void)normType; (void)crossCheck; }
    static Ptr<BFMatcher> create(int normType, bool crossCheck = false) { return Ptr<BFMatcher>(new BFMatcher(normType, crossCheck)); }
    void knnMatch(const Mat& q, const Mat& t, std::vector<std::vector<DMatch>>& matches, int k) const {
        assert(k == 2 && q.type() == CV_8UC1 && t.type() == CV_8UC1 && q.cols == 32 && t.cols == 32 && q.isContinuous() && t.isContinuous());
        (void)k;
        std::vector<int32_t> idx((size_t)q.rows * 2), dist((size_t)q.rows * 2);
        orc_knn2(q.data, q.rows, t.data, t.rows, idx.data(), dist.data());
        matches.assign(q.rows, std::vector<DMatch>());
        for (int i = 0; i < q.rows; ++i)
            for (int j = 0; j < 2; ++j)
                if (idx[2 * i + j] >= 0) matches[i].push_back(DMatch(i, idx[2 * i + j], (float)dist[2 * i + j]));
    }
};

struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>&, int) {
        std::fprintf(stderr, "cvshim: KeyPointsFilter::retainBest is only reached from dead code\n");
        std::abort();
    }
};

}  // namespace cv
