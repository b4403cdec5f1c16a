// TEST INFRASTRUCTURE — CPU oracle for the LBD line descriptor, never on the product path.
//
// Restates BinaryDescriptor::compute(image, keylines, descriptors) of opencv_contrib line_descriptor as vendored
// (unbuilt) in /root/reference/Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:
//   constructor weights gaussCoefL_/gaussCoefG_ :217-259 (integer-division quirks kept: sigma_L = 7, u_L = 10, u_G = sigma_G = 31)
//   computeGaussianPyramid + computeSobel      :350-398  (5x5 sigma-1 blur, Sobel 3x3 -> CV_16S; cv2-pinned prims)
//   computeImpl                                 :539-687  (single octave; rows indexed by class_id)
//   computeLBD                                  :1026-1372
//   binaryConversion over `combinations`        :74-107, :401-412
// PARITY UNPINNED by execution: the contrib module is not buildable here and cv2 has no line_descriptor; the pin
// is the vendored source text.  Float semantics fixed by the oracle: float32, no FMA contraction, sequential
// accumulation in the reference's loop order, cos/sin evaluated in double and rounded to float.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "cvprims.hpp"

namespace lbdo {

struct KeyLine {  // cv::line_descriptor::KeyLine, 68 bytes (descriptor_custom.hpp:105-144)
    float angle;
    int class_id, octave;
    float pt_x, pt_y, response, size;
    float startPointX, startPointY, endPointX, endPointY;
    float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength;
    int numOfPixels;
};
static_assert(sizeof(KeyLine) == 68, "KeyLine layout");

static const int kBands = 9, kBandW = 7, kH = kBands * kBandW;  // 63 rows
static const int kComb[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
                                 {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
                                 {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

void weights(double* L21, double* G63) {
    double u = (kBandW * 3 - 1) / 2;          // integer division: 10
    double sigma = (kBandW * 2 + 1) / 2;      // integer division: 7
    double inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < kBandW * 3; ++i) { double d = i - u; L21[i] = std::exp(d * d * inv); }
    u = (kBands * kBandW - 1) / 2;            // 31
    sigma = u;
    inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < kH; ++i) { double d = i - u; G63[i] = std::exp(d * d * inv); }
}

// One line: 72 floats (9 bands x {mean pgdL, ngdL, pgdO, ngdO, std pgdL, ngdL, pgdO, ngdO}), normalised.
void lbd_float(const int16_t* dxI, const int16_t* dyI, int w, int h, const KeyLine& kl, const double* L21, const double* G63, float* des) {
    float sums[kBands][8];
    std::memset(sums, 0, sizeof(sums));
    const short imageWidth = (short)(w - 1), imageHeight = (short)(h - 1);
    const short len = (short)kl.numOfPixels;
    const short halfWidth = (short)((len - 1) / 2), halfHeight = (short)((kH - 1) / 2);
    const float midX = (float)(0.5 * (kl.sPointInOctaveX + kl.ePointInOctaveX));
    const float midY = (float)(0.5 * (kl.sPointInOctaveY + kl.ePointInOctaveY));
    const float dL0 = (float)std::cos((double)kl.angle), dL1 = (float)std::sin((double)kl.angle);
    const float dO0 = -dL1, dO1 = dL0;
    float sCorX0 = -dL0 * halfWidth + dL1 * halfHeight + midX;
    float sCorY0 = -dL1 * halfWidth - dL0 * halfHeight + midY;
    for (short hID = 0; hID < kH; ++hID) {
        float sCorX = sCorX0, sCorY = sCorY0;
        float pL = 0, nL = 0, pO = 0, nO = 0;
        for (short wID = 0; wID < len; ++wID) {
            short t = (short)std::round(sCorX);
            const short xCor = t < 0 ? 0 : (t > imageWidth ? imageWidth : t);
            t = (short)std::round(sCorY);
            const short yCor = t < 0 ? 0 : (t > imageHeight ? imageHeight : t);
            const short dx = dxI[yCor * w + xCor], dy = dyI[yCor * w + xCor];
            const float gDL = dx * dL0 + dy * dL1;
            const float gDO = dx * dO0 + dy * dO1;
            if (gDL > 0) pL += gDL; else nL -= gDL;
            if (gDO > 0) pO += gDO; else nO -= gDO;
            sCorX += dL0;
            sCorY += dL1;
        }
        sCorX0 -= dL1;
        sCorY0 += dL0;
        float c = (float)G63[hID];
        pL = c * pL; nL = c * nL; pO = c * pO; nO = c * nO;
        const float pL2 = pL * pL, nL2 = nL * nL, pO2 = pO * pO, nO2 = nO * nO;
        const int band = hID / kBandW;
        const int tgt[3] = {band, band - 1, band + 1};
        const int off[3] = {kBandW, 2 * kBandW, 0};
        for (int k = 0; k < 3; ++k) {
            const int b = tgt[k];
            if (b < 0 || b >= kBands) continue;
            c = (float)L21[hID % kBandW + off[k]];
            sums[b][0] += c * pL;      sums[b][1] += c * nL;
            sums[b][4] += c * c * pL2; sums[b][5] += c * c * nL2;
            sums[b][2] += c * pO;      sums[b][3] += c * nO;
            sums[b][6] += c * c * pO2; sums[b][7] += c * c * nO2;
        }
    }
    const float invN2 = (float)(1.0 / (kBandW * 2.0)), invN3 = (float)(1.0 / (kBandW * 3.0));
    for (int b = 0; b < kBands; ++b) {
        const float invN = (b == 0 || b == kBands - 1) ? invN2 : invN3;
        for (int q = 0; q < 4; ++q) {
            const float temp = sums[b][q] * invN;
            des[8 * b + q] = temp;
            des[8 * b + 4 + q] = std::sqrt(sums[b][4 + q] * invN - temp * temp);
        }
    }
    float tempM = 0, tempS = 0;
    for (int b = 0; b < kBands; ++b) {
        for (int q = 0; q < 4; ++q) tempM += des[8 * b + q] * des[8 * b + q];
        for (int q = 4; q < 8; ++q) tempS += des[8 * b + q] * des[8 * b + q];
    }
    tempM = 1 / std::sqrt(tempM);
    tempS = 1 / std::sqrt(tempS);
    for (int b = 0; b < kBands; ++b) {
        for (int q = 0; q < 4; ++q) des[8 * b + q] = des[8 * b + q] * tempM;
        for (int q = 4; q < 8; ++q) des[8 * b + q] = des[8 * b + q] * tempS;
    }
    for (int i = 0; i < 72; ++i)
        if (des[i] > 0.4) des[i] = (float)0.4;
    float temp = 0;
    for (int i = 0; i < 72; ++i) temp += des[i] * des[i];
    temp = 1 / std::sqrt(temp);
    for (int i = 0; i < 72; ++i) des[i] = des[i] * temp;
}

}  // namespace lbdo

extern "C" {

// gray -> (5x5 sigma-1 blur) -> Sobel dx, dy (CV_16S); the LBD preprocessing of the reference
void orc_lbd_gradients(const uint8_t* gray, int w, int h, int16_t* dx, int16_t* dy) {
    std::vector<uint8_t> bl((size_t)w * h);
    cvp::gaussian_blur5_s1(gray, w, h, w, bl.data(), w);
    cvp::sobel3_s16(bl.data(), w, h, w, dx, dy);
}

// keylines: n x 68-byte KeyLine (single octave).  desc: n x 32 bytes; fdesc (optional): n x 72 floats.
void orc_lbd_compute(const uint8_t* gray, int w, int h, const void* keylines, int n, uint8_t* desc, float* fdesc) {
    std::vector<int16_t> dx((size_t)w * h), dy((size_t)w * h);
    orc_lbd_gradients(gray, w, h, dx.data(), dy.data());
    double L21[21], G63[63];
    lbdo::weights(L21, G63);
    const lbdo::KeyLine* kl = (const lbdo::KeyLine*)keylines;
    for (int i = 0; i < n; ++i) {
        float des[72];
        lbdo::lbd_float(dx.data(), dy.data(), w, h, kl[i], L21, G63, des);
        if (fdesc) std::memcpy(fdesc + 72 * (size_t)i, des, sizeof(des));
        for (int c = 0; c < 32; ++c) {
            const float* f1 = &des[8 * lbdo::kComb[c][0]];
            const float* f2 = &des[8 * lbdo::kComb[c][1]];
            unsigned v = 0;
            for (int b = 0; b < 8; ++b)
                if (f1[b] > f2[b]) v += 1u << b;
            desc[32 * (size_t)i + c] = (uint8_t)v;
        }
    }
}

}  // extern "C"

// ---- LSD wrapper: KeyLine fields from raw segments -----------------------------------------------------------
// Restates LSDDetectorC::detectImpl's keyline loop for one octave (Thirdparty/line_descriptor/src/
// LSDDetector_custom.cpp:76-103 checkLineExtremes, :160-198) and cv::LineIterator::count (8-connected:
// max(|dx|,|dy|)+1 on cvRound'ed endpoints; endpoints are already clamped inside the image).
extern "C" void orc_keylines_from_segments(const float* seg4, int n, int w, int h, void* keylines_out) {
    lbdo::KeyLine* out = (lbdo::KeyLine*)keylines_out;
    for (int k = 0; k < n; ++k) {
        float e[4] = {seg4[4 * k], seg4[4 * k + 1], seg4[4 * k + 2], seg4[4 * k + 3]};
        if (e[0] < 0) e[0] = 0;
        if (e[0] >= w) e[0] = (float)w - 1.0f;
        if (e[2] < 0) e[2] = 0;
        if (e[2] >= w) e[2] = (float)w - 1.0f;
        if (e[1] < 0) e[1] = 0;
        if (e[1] >= h) e[1] = (float)h - 1.0f;
        if (e[3] < 0) e[3] = 0;
        if (e[3] >= h) e[3] = (float)h - 1.0f;
        lbdo::KeyLine kl;
        const float octaveScale = 1.0f;  // pow((float)scale, 0)
        kl.startPointX = e[0] * octaveScale; kl.startPointY = e[1] * octaveScale;
        kl.endPointX = e[2] * octaveScale; kl.endPointY = e[3] * octaveScale;
        kl.sPointInOctaveX = e[0]; kl.sPointInOctaveY = e[1]; kl.ePointInOctaveX = e[2]; kl.ePointInOctaveY = e[3];
        kl.lineLength = (float)std::sqrt(std::pow((double)(e[0] - e[2]), 2) + std::pow((double)(e[1] - e[3]), 2));
        const int x0 = cvp::cv_round(e[0]), y0 = cvp::cv_round(e[1]), x1 = cvp::cv_round(e[2]), y1 = cvp::cv_round(e[3]);
        kl.numOfPixels = std::max(std::abs(x1 - x0), std::abs(y1 - y0)) + 1;
        kl.angle = (float)std::atan2((double)(kl.endPointY - kl.startPointY), (double)(kl.endPointX - kl.startPointX));
        kl.class_id = k;
        kl.octave = 0;
        kl.size = (kl.endPointX - kl.startPointX) * (kl.endPointY - kl.startPointY);
        kl.response = kl.lineLength / (float)std::max(w, h);
        kl.pt_x = (kl.endPointX + kl.startPointX) / 2;
        kl.pt_y = (kl.endPointY + kl.startPointY) / 2;
        out[k] = kl;
    }
}
