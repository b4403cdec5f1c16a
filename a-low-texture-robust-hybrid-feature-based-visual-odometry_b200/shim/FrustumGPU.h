// Batched Frame::isInFrustum for the local map (reference src/Frame.cc:1371-1499), called where Tracking::SearchLocalPoints /
// SearchLocalLines loop over mvpLocalMapPoints / mvpLocalMapLines (src/Tracking.cc:3251-3264, 3315-3340):
//
//     hvo_shim::FrustumCullerT<Frame, MapPoint, MapLine> culler;                     // long-lived, next to the extractors
//     culler.isInFrustum(mCurrentFrame, candidates, 0.5f, inView);                   // instead of the per-point loop
//
// Writes exactly what the reference's isInFrustum writes into every map point / map line (mbTrackInView, mTrackProjX, mTrackProjY,
// mTrackProjXR, mnTrackScaleLevel, mTrackViewCos resp. mTrackProjX1..Y2) and reports the return value per element.  It reads
// mfMinDistance / mfMaxDistance of the map element directly (PredictScale uses mfMaxDistance, for which the reference has no public
// getter): the maintainer adds `friend class hvo_shim::FrustumCullerT<...>` (or two getters) to MapPoint.h / MapLine.h.
#ifndef HVO_SHIM_FRUSTUM_GPU_H
#define HVO_SHIM_FRUSTUM_GPU_H

#include <cstdio>
#include <vector>

#include "hvo_capi.h"

namespace hvo_shim {

template <class Frame, class MapPoint, class MapLine>
class FrustumCullerT {
public:
    explicit FrustumCullerT(int device = 0) : hp_(nullptr), hl_(nullptr) {
        if (hvo_proj_create(device, &hp_) != HVO_OK || hvo_lproj_create(device, &hl_) != HVO_OK) std::fprintf(stderr, "FrustumCuller: %s\n", hvo_last_error());
    }
    ~FrustumCullerT() { hvo_proj_destroy(hp_); hvo_lproj_destroy(hl_); }
    FrustumCullerT(const FrustumCullerT&) = delete;
    FrustumCullerT& operator=(const FrustumCullerT&) = delete;

    // returns the number of map points in view
    int isInFrustum(Frame& F, const std::vector<MapPoint*>& v, float viewingCosLimit, std::vector<char>& inView) {
        const int n = (int)v.size();
        inView.assign(n, 0);
        if (n == 0) return 0;
        const hvo_frustum_cam cam = camOf(F);
        std::vector<hvo_map_point> pts(n);
        for (int i = 0; i < n; ++i) {
            const cv::Mat P = v[i]->GetWorldPos(), N = v[i]->GetNormal();
            for (int k = 0; k < 3; ++k) { pts[i].pos[k] = P.template at<float>(k); pts[i].normal[k] = N.template at<float>(k); }
            pts[i].min_distance = v[i]->mfMinDistance; pts[i].max_distance = v[i]->mfMaxDistance;
        }
        std::vector<hvo_track_point> out(n);
        if (hvo_proj_frustum_points(hp_, &cam, pts.data(), n, viewingCosLimit, out.data()) != HVO_OK) {
            std::fprintf(stderr, "FrustumCuller: %s\n", hvo_last_error());
            return 0;
        }
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            v[i]->mbTrackInView = out[i].in_view != 0;
            if (!out[i].in_view) continue;
            v[i]->mTrackProjX = out[i].u; v[i]->mTrackProjXR = out[i].ur; v[i]->mTrackProjY = out[i].v;
            v[i]->mnTrackScaleLevel = out[i].level; v[i]->mTrackViewCos = out[i].view_cos;
            inView[i] = 1; ++cnt;
        }
        return cnt;
    }
    int isInFrustum(Frame& F, const std::vector<MapLine*>& v, float viewingCosLimit, std::vector<char>& inView) {
        const int n = (int)v.size();
        inView.assign(n, 0);
        if (n == 0) return 0;
        const hvo_frustum_cam cam = camOf(F);
        std::vector<hvo_map_line> ml(n);
        for (int i = 0; i < n; ++i) {
            const auto P = v[i]->GetWorldPos();
            const auto N = v[i]->GetNormal();
            for (int k = 0; k < 6; ++k) ml[i].pos[k] = P(k);
            for (int k = 0; k < 3; ++k) { ml[i].normal[k] = N(k); ml[i].dir[k] = 0.0; }
            ml[i].min_distance = v[i]->mfMinDistance; ml[i].max_distance = v[i]->mfMaxDistance;
        }
        std::vector<hvo_track_line> out(n);
        if (hvo_lproj_frustum_lines(hl_, &cam, ml.data(), n, viewingCosLimit, out.data()) != HVO_OK) {
            std::fprintf(stderr, "FrustumCuller: %s\n", hvo_last_error());
            return 0;
        }
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            v[i]->mbTrackInView = out[i].in_view != 0;
            if (!out[i].in_view) continue;
            v[i]->mTrackProjX1 = out[i].x1; v[i]->mTrackProjY1 = out[i].y1; v[i]->mTrackProjX2 = out[i].x2; v[i]->mTrackProjY2 = out[i].y2;
            v[i]->mnTrackScaleLevel = out[i].level; v[i]->mTrackViewCos = out[i].view_cos;
            inView[i] = 1; ++cnt;
        }
        return cnt;
    }

private:
    static hvo_frustum_cam camOf(const Frame& F) {
        hvo_frustum_cam c;
        for (int i = 0; i < 3; ++i) {
            for (int k = 0; k < 3; ++k) c.Rcw[3 * i + k] = F.mRcw.template at<float>(i, k);
            c.tcw[i] = F.mtcw.template at<float>(i); c.Ow[i] = F.mOw.template at<float>(i);
        }
        c.fx = Frame::fx; c.fy = Frame::fy; c.cx = Frame::cx; c.cy = Frame::cy; c.bf = F.mbf;
        c.min_x = Frame::mnMinX; c.min_y = Frame::mnMinY; c.max_x = Frame::mnMaxX; c.max_y = Frame::mnMaxY;
        c.log_scale_factor = F.mfLogScaleFactor; c.n_levels = F.mnScaleLevels;
        return c;
    }
    hvo_proj* hp_;
    hvo_lproj* hl_;
};

}  // namespace hvo_shim

#endif
