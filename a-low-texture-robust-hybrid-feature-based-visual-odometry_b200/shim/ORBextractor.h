// Drop-in replacement for the reference's include/ORBextractor.h: same namespace, class name, constructor,
// operator() and getters (reference include/ORBextractor.h:46-110), forwarding to the C ABI of libhvofront.so
// (include/hvo_capi.h).  Tracking.cc / Frame.cc compile against this header unchanged:
//
//   mpORBextractorLeft = new ORBextractor(nFeatures, fScaleFactor, nLevels, fIniThFAST, fMinThFAST);  // Tracking.cc:124
//   (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors);                                      // Frame.cc:876
//
// Needs <opencv2/core/core.hpp> (cv::Mat, cv::KeyPoint, cv::InputArray/OutputArray); in this repository's
// tests the OpenCV stand-in under oracle/cvshim provides those names.
#ifndef HVO_SHIM_ORBEXTRACTOR_H
#define HVO_SHIM_ORBEXTRACTOR_H

#include <cstdio>
#include <cstring>
#include <vector>

#include <opencv2/core/core.hpp>

#include "hvo_capi.h"

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
        : h_(nullptr), w_(0), hgt_(0), nlevels_(nlevels), scaleFactor_(scaleFactor) {
        p_.nfeatures = nfeatures;
        p_.scale_factor = scaleFactor;
        p_.nlevels = nlevels;
        p_.ini_th_fast = iniThFAST;
        p_.min_th_fast = minThFAST;
        // scale tables do not depend on the image size (ORBextractor.cc:413-429)
        mvScaleFactor.assign(nlevels, 1.0f);
        mvLevelSigma2.assign(nlevels, 1.0f);
        for (int i = 1; i < nlevels; i++) {
            mvScaleFactor[i] = (float)((double)mvScaleFactor[i - 1] * scaleFactor_);
            mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
        }
        mvInvScaleFactor.resize(nlevels);
        mvInvLevelSigma2.resize(nlevels);
        for (int i = 0; i < nlevels; i++) {
            mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
            mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
        }
        mvImagePyramid.resize(nlevels);
    }
    ~ORBextractor() { hvo_orb_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image; the mask is ignored, as in the reference.
    void operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& _keypoints,
                    cv::OutputArray _descriptors) {
        if (_image.empty()) return;  // ORBextractor.cc:1044-1045
        cv::Mat image = _image.getMat();
        if (image.type() != CV_8UC1) {  // the reference asserts (ORBextractor.cc:1048)
            std::fprintf(stderr, "ORBextractor: image must be CV_8UC1\n");
            _keypoints.clear();
            _descriptors.release();
            return;
        }
        if (!h_ || image.cols != w_ || image.rows != hgt_) {
            hvo_orb_destroy(h_);
            h_ = nullptr;
            if (hvo_orb_create(&p_, image.cols, image.rows, 1, 0, &h_) != HVO_OK) {
                std::fprintf(stderr, "ORBextractor: %s\n", hvo_last_error());
                _keypoints.clear();
                _descriptors.release();
                return;  // failure maps to "empty outputs" (no CPU fallback exists)
            }
            w_ = image.cols;
            hgt_ = image.rows;
        }
        const int cap = hvo_orb_capacity(h_);
        kps_.resize(cap);
        desc_.resize((size_t)cap * 32);
        int n = 0;
        if (hvo_orb_extract(h_, image.data, (size_t)image.step, kps_.data(), desc_.data(), cap, &n) != HVO_OK) {
            std::fprintf(stderr, "ORBextractor: %s\n", hvo_last_error());
            n = 0;
        }
        _keypoints.clear();
        if (n == 0) {
            _descriptors.release();
            return;
        }
        _descriptors.create(n, 32, CV_8U);
        cv::Mat descriptors = _descriptors.getMat();
        _keypoints.resize(n);
        static_assert(sizeof(cv::KeyPoint) == sizeof(hvo_keypoint), "cv::KeyPoint layout");
        std::memcpy((void*)_keypoints.data(), kps_.data(), (size_t)n * sizeof(hvo_keypoint));
        for (int i = 0; i < n; ++i) std::memcpy(descriptors.ptr(i), &desc_[(size_t)i * 32], 32);
        pyramid_valid_ = false;
    }

    int inline GetLevels() { return nlevels_; }
    float inline GetScaleFactor() { return (float)scaleFactor_; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // The reference exposes the pyramid as a public member read only by the stereo matcher
    // (Frame.cc:1770,1860,1877).  It is filled on demand to keep the RGB-D path free of device->host copies.
    std::vector<cv::Mat> mvImagePyramid;
    void FetchImagePyramid() {
        if (!h_ || pyramid_valid_) return;
        for (int l = 0; l < nlevels_; ++l) {
            int lw = 0, lh = 0;
            hvo_orb_level_size(h_, l, &lw, &lh);
            mvImagePyramid[l].create(lh, lw, CV_8UC1);
            hvo_orb_get_pyramid_level(h_, 0, l, mvImagePyramid[l].data, (size_t)mvImagePyramid[l].step);
        }
        pyramid_valid_ = true;
    }

protected:
    hvo_orb* h_;
    hvo_orb_params p_;
    int w_, hgt_, nlevels_;
    double scaleFactor_;
    bool pyramid_valid_ = false;
    std::vector<hvo_keypoint> kps_;
    std::vector<uint8_t> desc_;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

}  // namespace ORB_SLAM2

#endif
