// TEST INFRASTRUCTURE — CPU oracle for the ORB extractor, never on the product path.
//
// Restates /root/reference/src/ORBextractor.cc (ctor :408-468, ComputePyramid :1105-1130,
// ComputeKeyPointsOctTree :763-851, DistributeOctTree :537-761, DivideNode :479-535, IC_Angle :75-102,
// computeOrbDescriptor :106-144, operator() :1041-1103) on top of the cv2-pinned primitives in
// cvprims.hpp.  First-party logic is additionally pinned by oracle/_ref (the reference's own
// ORBextractor.cc compiled against the cvshim stand-in, see oracle/Makefile).
//
// Deterministic choices: octree size ties are broken by creation order, later
// created node first (= "higher heap address" under a monotonic allocator); float32 without FMA
// contraction; cos/sin evaluated in double and rounded to float; cvRound = round-half-even.
#pragma once
#include <cstdint>
#include <vector>

namespace orbo {

struct Params {
    int nfeatures = 1000;
    float scale_factor = 1.2f;
    int nlevels = 8;
    int ini_th = 20;
    int min_th = 7;
};

struct KeyPoint {  // cv::KeyPoint layout, 28 bytes
    float x, y, size, angle, response;
    int octave, class_id;
};

struct Candidate { float x, y; float response; };  // level coords relative to the 16-px border

struct Level {
    int w = 0, h = 0;
    std::vector<uint8_t> img;       // unpadded level image
    std::vector<uint8_t> blurred;   // 7x7 sigma-2 blur (only filled when the level has keypoints)
    std::vector<Candidate> cand;    // pre-octree list, reference order
    std::vector<KeyPoint> kps;      // post-octree, oriented, level coordinates
};

class Extractor {
public:
    explicit Extractor(const Params& p);
    // gray: h x w, row stride `stride` bytes
    void extract(const uint8_t* gray, int w, int h, size_t stride, std::vector<KeyPoint>& kps,
                 std::vector<uint8_t>& desc);
    const std::vector<Level>& levels() const { return lv_; }
    const std::vector<float>& scale_factors() const { return sf_; }
    const std::vector<float>& inv_scale_factors() const { return isf_; }
    const std::vector<int>& features_per_level() const { return nfeat_; }
    const std::vector<int>& umax() const { return umax_; }

    static std::vector<Candidate> distribute(const std::vector<Candidate>& in, int minX, int maxX, int minY,
                                             int maxY, int N);

private:
    void pyramid(const uint8_t* gray, int w, int h, size_t stride);
    void detect();
    Params p_;
    std::vector<float> sf_, isf_;
    std::vector<int> nfeat_, umax_;
    std::vector<Level> lv_;
};

}  // namespace orbo
