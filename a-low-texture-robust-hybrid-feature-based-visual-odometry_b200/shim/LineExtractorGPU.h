// GPU body of ORB_SLAM2::LINEextractor::operator() and of the LBD recomputation in Frame::cullingLine.
//
// The reference's LINEextractor (include/LineExtractor.h:187-262) carries many host-only helpers (3-D line fitting,
// Mahalanobis tests, ...) that stay as they are, so this is not a header swap: the maintainer keeps LineExtractor.h and
// replaces two bodies in src/LineExtractor.cpp / src/Frame.cc with calls into the bridge below:
//
//   void LINEextractor::operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<KeyLine>& _keylines,
//                                  cv::OutputArray _descriptors, std::vector<Eigen::Vector3d>& _lineVec2d) {   // LineExtractor.cpp:329
//       if (_image.empty()) return;
//       gpu_.extract(_image.getMat(), _keylines, _descriptors, _lineVec2d);   // hvo_shim::LineFrontEnd gpu_{numOctaves, scale, nLSDFeature, min_line_length};
//   }
//   ...
//   lbd->compute(im, mvKeylinesUn, mLdesc);        // Frame.cc:1094-1096   ->   mpLSDextractorLeft->gpu().computeLBD(im, mvKeylinesUn, mLdesc);
//
// KeyLineT is cv::line_descriptor::KeyLine (Thirdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp:105-144,
// a 68-byte POD identical to hvo_keyline); Vec3T is Eigen::Vector3d (anything with operator[] works).
#ifndef HVO_SHIM_LINEEXTRACTOR_GPU_H
#define HVO_SHIM_LINEEXTRACTOR_GPU_H

#include <cstdio>
#include <cstring>
#include <vector>

#include <opencv2/core/core.hpp>

#include "hvo_capi.h"

namespace hvo_shim {

class LineFrontEnd {
public:
    LineFrontEnd(int numOctaves, float scale, unsigned int nLSDFeature, double min_line_length, int device = 0)
        : h_(nullptr), lbd_(nullptr), w_(0), hgt_(0), lw_(0), lh_(0), lcap_(0), device_(device) {
        p_.n_octaves = numOctaves;
        p_.scale = scale;
        p_.n_features = (int)nLSDFeature;
        p_.min_line_length = min_line_length;
    }
    ~LineFrontEnd() {
        hvo_line_destroy(h_);
        hvo_lbd_destroy(lbd_);
    }
    LineFrontEnd(const LineFrontEnd&) = delete;
    LineFrontEnd& operator=(const LineFrontEnd&) = delete;

    // LINEextractor::operator()  (src/LineExtractor.cpp:329-380).  The mask is ignored: the reference always passes an empty one.
    template <class KeyLineT, class Vec3T>
    void extract(const cv::Mat& image, std::vector<KeyLineT>& keylines, cv::OutputArray descriptors, std::vector<Vec3T>& lineVec2d) {
        static_assert(sizeof(KeyLineT) == sizeof(hvo_keyline), "KeyLine layout (descriptor_custom.hpp:105-144)");
        keylines.clear();
        lineVec2d.clear();
        if (image.empty()) return;  // LineExtractor.cpp:331-332
        if (image.type() != CV_8UC1) { std::fprintf(stderr, "LINEextractor: image must be CV_8UC1\n"); descriptors.release(); return; }
        if (!h_ || image.cols != w_ || image.rows != hgt_) {
            hvo_line_destroy(h_);
            h_ = nullptr;
            if (hvo_line_create(&p_, image.cols, image.rows, 1, device_, &h_) != HVO_OK) {
                std::fprintf(stderr, "LINEextractor: %s\n", hvo_last_error());
                descriptors.release();
                return;  // no CPU fallback: failure maps to "no lines"
            }
            w_ = image.cols; hgt_ = image.rows;
        }
        const int cap = hvo_line_max_lines(h_);
        kl_.resize(cap);
        desc_.resize((size_t)cap * 32);
        lv_.resize((size_t)cap * 3);
        int n = 0;
        if (hvo_line_extract(h_, image.data, (size_t)image.step, kl_.data(), desc_.data(), lv_.data(), cap, &n) != HVO_OK) {
            std::fprintf(stderr, "LINEextractor: %s\n", hvo_last_error());
            n = 0;
        }
        if (n == 0) { descriptors.release(); return; }
        keylines.resize(n);
        std::memcpy((void*)keylines.data(), kl_.data(), (size_t)n * sizeof(hvo_keyline));
        descriptors.create(n, 32, CV_8U);
        cv::Mat d = descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), &desc_[(size_t)i * 32], 32);
        lineVec2d.resize(n);
        for (int i = 0; i < n; ++i) { lineVec2d[i][0] = lv_[3 * i]; lineVec2d[i][1] = lv_[3 * i + 1]; lineVec2d[i][2] = lv_[3 * i + 2]; }
    }

    // cv::line_descriptor::BinaryDescriptor::compute(image, keylines, descriptors) as called at src/Frame.cc:1094-1096
    template <class KeyLineT>
    void computeLBD(const cv::Mat& image, std::vector<KeyLineT>& keylines, cv::OutputArray descriptors) {
        static_assert(sizeof(KeyLineT) == sizeof(hvo_keyline), "KeyLine layout");
        const int n = (int)keylines.size();
        if (n == 0 || image.empty()) { descriptors.release(); return; }
        if (!lbd_ || image.cols != lw_ || image.rows != lh_ || n > lcap_) {
            hvo_lbd_destroy(lbd_);
            lbd_ = nullptr;
            lcap_ = n > 256 ? n : 256;
            if (hvo_lbd_create(image.cols, image.rows, 1, lcap_, device_, &lbd_) != HVO_OK) {
                std::fprintf(stderr, "BinaryDescriptor: %s\n", hvo_last_error());
                descriptors.release();
                return;
            }
            lw_ = image.cols; lh_ = image.rows;
        }
        desc_.resize((size_t)n * 32);
        if (hvo_lbd_compute(lbd_, image.data, (size_t)image.step, reinterpret_cast<const hvo_keyline*>(keylines.data()), n, desc_.data()) != HVO_OK) {
            std::fprintf(stderr, "BinaryDescriptor: %s\n", hvo_last_error());
            descriptors.release();
            return;
        }
        descriptors.create(n, 32, CV_8U);
        cv::Mat d = descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), &desc_[(size_t)i * 32], 32);
    }

    // Frame::cullingLine(imGray, 5, 2.5, 15, 30) (src/Frame.cc:952-1116) in one call: the body of that function becomes
    //   mpLSDextractorLeft->gpu().cullingLine(imGray, mvKeylinesUn, mLdesc, mvKeyLineFunctions);
    // keylines / lineVec2d are replaced by the merged, response-sorted set, descriptors by their LBD.
    template <class KeyLineT, class Vec3T>
    void cullingLine(const cv::Mat& image, std::vector<KeyLineT>& keylines, cv::OutputArray descriptors, std::vector<Vec3T>& lineVec2d) {
        static_assert(sizeof(KeyLineT) == sizeof(hvo_keyline), "KeyLine layout");
        const int n = (int)keylines.size();
        if (n == 0 || image.empty() || !h_ || image.cols != w_ || image.rows != hgt_ || n > hvo_line_max_lines(h_)) return;
        const int cap = hvo_line_max_lines(h_);
        kl_.resize(cap);
        desc_.resize((size_t)cap * 32);
        lv_.resize((size_t)cap * 3);
        std::memcpy(kl_.data(), (const void*)keylines.data(), (size_t)n * sizeof(hvo_keyline));
        for (int i = 0; i < n; ++i) { lv_[3 * i] = lineVec2d[i][0]; lv_[3 * i + 1] = lineVec2d[i][1]; lv_[3 * i + 2] = lineVec2d[i][2]; }
        int m = 0;
        if (hvo_line_cull(h_, image.data, (size_t)image.step, kl_.data(), lv_.data(), n, desc_.data(), &m) != HVO_OK) {
            std::fprintf(stderr, "cullingLine: %s\n", hvo_last_error());
            return;
        }
        keylines.resize(m);
        std::memcpy((void*)keylines.data(), kl_.data(), (size_t)m * sizeof(hvo_keyline));
        descriptors.create(m, 32, CV_8U);
        cv::Mat d = descriptors.getMat();
        for (int i = 0; i < m; ++i) std::memcpy(d.ptr(i), &desc_[(size_t)i * 32], 32);
        lineVec2d.resize(m);
        for (int i = 0; i < m; ++i) { lineVec2d[i][0] = lv_[3 * i]; lineVec2d[i][1] = lv_[3 * i + 1]; lineVec2d[i][2] = lv_[3 * i + 2]; }
    }

private:
    hvo_line* h_;
    hvo_lbd* lbd_;
    hvo_line_params p_;
    int w_, hgt_, lw_, lh_, lcap_, device_;
    std::vector<hvo_keyline> kl_;
    std::vector<uint8_t> desc_;
    std::vector<double> lv_;
};

}  // namespace hvo_shim

#endif
