// TEST INFRASTRUCTURE — field-for-field stand-in for cv::line_descriptor::KeyLine
// (reference Thirdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp:105-144); the real header needs OpenCV.
#pragma once
#include <opencv2/core/core.hpp>
namespace cv { namespace line_descriptor {
struct KeyLine {
    float angle;
    int class_id;
    int octave;
    Point2f pt;
    float response;
    float size;
    float startPointX, startPointY, endPointX, endPointY;
    float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength;
    int numOfPixels;
};
static_assert(sizeof(KeyLine) == 68, "KeyLine layout");
}}  // namespace cv::line_descriptor
