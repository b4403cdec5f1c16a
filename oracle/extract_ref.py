#!/usr/bin/env python
"""TEST INFRASTRUCTURE — pulls line ranges out of the reference's own sources, at BUILD time, into oracle/_ref/gen/*.inc so
that they can be compiled (unmodified, where they lie being impossible: the files as a whole need the full OpenCV / PCL / g2o
APIs) into the oracle/_ref harnesses.  Nothing extracted is committed: oracle/_ref/ is git-ignored.

    python oracle/extract_ref.py [/root/reference] [oracle/_ref/gen]

Every range carries an anchor: a substring its first line must contain, so that a drifted reference fails loudly instead of
compiling something else."""
import os
import sys

# name -> list of (file, first line, last line, anchor in the first line); 1-based inclusive, concatenated in order
RANGES = {
    # LBD of the vendored line_descriptor: includes/defines/namespace + combinations table + Params ctor; create + ctor; dtor ..
    # binaryConversion; compute + computeImpl; computeLBD.  (The rest of the file is the EDLines detector, never called.)
    'lbd': [
        ('Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp', 42, 116, '#include "precomp_custom.hpp"'),
        ('Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp', 206, 259, 'BinaryDescriptor::createBinaryDescriptor()'),
        ('Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp', 302, 412, 'BinaryDescriptor::~BinaryDescriptor()'),
        ('Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp', 523, 687, '/* requires descriptors computation (only one image) */'),
        ('Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp', 1026, 1372, 'int BinaryDescriptor::computeLBD('),
    ],
    # sort_lines_by_response
    'auxiliar': [('include/auxiliar.h', 47, 52, 'struct sort_lines_by_response')],
    # LINEextractor::operator()
    'line_extractor': [('src/LineExtractor.cpp', 329, 380, 'void LINEextractor::operator()')],
    # Frame::cullingLine, PointLineDistance, TwoLineAngle, MergeTwoLines
    'frame_cull': [('src/Frame.cc', 952, 1203, 'void Frame::cullingLine(')],
    # Frame grids: AssignFeaturesToGrid, AssignFeaturesToGridForLine; GetFeaturesInArea (the two '#pragma GCC' lines around it
    # are left out), GetFeaturesInAreaForLine; PosInGrid
    'frame_grid': [('src/Frame.cc', 832, 872, 'void Frame::AssignFeaturesToGrid()'),
                   ('src/Frame.cc', 1502, 1555, 'vector<size_t> Frame::GetFeaturesInArea('),
                   ('src/Frame.cc', 1557, 1631, 'vector<size_t> Frame::GetFeaturesInAreaForLine('),
                   ('src/Frame.cc', 1680, 1690, 'bool Frame::PosInGrid(')],
    # ORBmatcher: constants + ctor, SearchByProjection(F, MapPoints, th) + RadiusByViewingCos, SearchByProjection(Cur, Last, th,
    # mono), ComputeThreeMaxima + DescriptorDistance
    'orbmatcher': [('src/ORBmatcher.cc', 37, 43, 'const int ORBmatcher::TH_HIGH'),
                   ('src/ORBmatcher.cc', 45, 140, 'int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints'),
                   ('src/ORBmatcher.cc', 1353, 1497, 'int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame'),
                   ('src/ORBmatcher.cc', 1630, 1692, 'void ORBmatcher::ComputeThreeMaxima(')],
    # LSDmatcher: constants + ctor + computeAngle2D, SearchByProjection(Cur, Last, th), SearchByProjection(F, MapLines, ...),
    # DescriptorDistance, RadiusByViewingCos
    'lsdmatcher': [('src/LSDmatcher.cpp', 12, 34, 'const int LSDmatcher::TH_HIGH'),
                   ('src/LSDmatcher.cpp', 561, 664, 'int LSDmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame'),
                   ('src/LSDmatcher.cpp', 709, 801, 'int LSDmatcher::SearchByProjection(Frame &F, const std::vector<MapLine *> &vpMapLines'),
                   ('src/LSDmatcher.cpp', 1137, 1153, 'int LSDmatcher::DescriptorDistance('),
                   ('src/LSDmatcher.cpp', 1436, 1442, 'float LSDmatcher::RadiusByViewingCos(')],
    # Frame::isInFrustum(MapPoint*, float) and (MapLine*, float); compiled under the name isInFrustumRef (a #define around the include)
    # because the windowed line search above keeps its documented stand-in of the same name
    'frame_frustum': [('src/Frame.cc', 1371, 1499, 'bool Frame::isInFrustum(MapPoint *pMP, float viewingCosLimit)')],
    # MapPoint::Get{Min,Max}DistanceInvariance + PredictScale(dist, Frame*); MapLine::Get{Min,Max}DistanceInvariance + PredictScale
    'map_scale': [('src/MapPoint.cc', 371, 381, 'float MapPoint::GetMinDistanceInvariance()'),
                  ('src/MapPoint.cc', 400, 415, 'int MapPoint::PredictScale(const float &currentDist, Frame *pF)'),
                  ('src/MapLine.cpp', 537, 558, 'float MapLine::GetMinDistanceInvariance()')],
    # ORBmatcher::SearchForInitialization (whole, with its rotation histogram)
    'orb_init': [('src/ORBmatcher.cc', 412, 529, 'int ORBmatcher::SearchForInitialization(')],
    # ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches)
    'orb_bow': [('src/ORBmatcher.cc', 162, 293, 'int ORBmatcher::SearchByBoW(KeyFrame* pKF,Frame &F')],
    # ORBmatcher::CheckDistEpipolarLine, SearchForTriangulation
    'orb_tri': [('src/ORBmatcher.cc', 143, 160, 'bool ORBmatcher::CheckDistEpipolarLine('),
                ('src/ORBmatcher.cc', 668, 836, 'int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2')],
    # ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th); KeyFrame::GetFeaturesInArea, IsInImage; MapPoint::PredictScale(dist, KeyFrame*)
    'orb_fuse': [('src/ORBmatcher.cc', 838, 994, 'int ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th)')],
    'keyframe_area': [('src/KeyFrame.cc', 627, 666, 'vector<size_t> KeyFrame::GetFeaturesInArea(const float &x, const float &y, const float &r) const'),
                      ('src/KeyFrame.cc', 780, 783, 'bool KeyFrame::IsInImage(')],
    'mappoint_scale_kf': [('src/MapPoint.cc', 383, 398, 'int MapPoint::PredictScale(const float &currentDist, KeyFrame *pKF)')],
    # ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (loop detection); SearchByBoW(pKF1, pKF2, vpMatches12);
    # Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint); SearchBySim3; SearchByProjection(CurrentFrame, KeyFrame*, sAlreadyFound, th, ORBdist)
    'orb_kf_scw': [('src/ORBmatcher.cc', 295, 410, 'int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw')],
    'orb_bow_kf': [('src/ORBmatcher.cc', 531, 666, 'int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2')],
    'orb_fuse_scw': [('src/ORBmatcher.cc', 996, 1121, 'int ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw')],
    'orb_sim3': [('src/ORBmatcher.cc', 1123, 1351, 'int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2')],
    'orb_reloc': [('src/ORBmatcher.cc', 1499, 1628, 'int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound')],
    # KeyFrame::GetMapPoints; MapPoint::GetIndexInKeyFrame
    'keyframe_mps': [('src/KeyFrame.cc', 254, 267, 'set<MapPoint*> KeyFrame::GetMapPoints()')],
    'mappoint_index': [('src/MapPoint.cc', 313, 320, 'int MapPoint::GetIndexInKeyFrame(KeyFrame *pKF)')],
    # the two MAD comparators; LSDmatcher::SearchByDescriptor; matchNNR, match, SearchDouble x2, FrameBFMatch; LSDmatcher::lineDescriptorMAD;
    # Frame::lineDescriptorMAD (what SearchByDescriptor calls)
    'lsd_bf': [('include/auxiliar.h', 26, 38, 'struct compare_descriptor_by_NN_dist'),
               ('src/LSDmatcher.cpp', 522, 559, 'int LSDmatcher::SearchByDescriptor('),
               ('src/LSDmatcher.cpp', 803, 966, 'int LSDmatcher::matchNNR('),
               ('src/LSDmatcher.cpp', 1110, 1135, 'void LSDmatcher::lineDescriptorMAD('),
               ('src/Frame.cc', 1331, 1355, 'void Frame::lineDescriptorMAD(')],
    # LSDmatcher::SearchForTriangulation(pKF1, pKF2, vector<pair>&) and (pKF1, pKF2, vector<int>&, isDouble)
    'lsd_tri': [('src/LSDmatcher.cpp', 1155, 1231, 'int LSDmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2,')],
    # Manhattan::computeNormalsLPVO, removeMatRow, removeMatCol
    'manhattan_lpvo': [('src/Manhattan.cpp', 237, 393, 'void Manhattan::computeNormalsLPVO('),
                       ('src/Manhattan.cpp', 395, 493, 'void Manhattan::removeMatRow(')],
    # MapPoint::ComputeDistinctiveDescriptors, MapLine::ComputeDistinctiveDescriptors
    'distinctive': [('src/MapPoint.cc', 240, 305, 'void MapPoint::ComputeDistinctiveDescriptors()'),
                    ('src/MapLine.cpp', 331, 396, 'void MapLine::ComputeDistinctiveDescriptors()')],
    # sort_descriptor_by_queryIdx; LSDmatcher::FrameBFMatchNew + mutualOverlap
    'lsd_bfnew': [('include/auxiliar.h', 40, 45, 'struct sort_descriptor_by_queryIdx'),
                  ('src/LSDmatcher.cpp', 968, 1108, 'void LSDmatcher::FrameBFMatchNew(')],
}


def extract(ref, out):
    os.makedirs(out, exist_ok=True)
    for name, parts in RANGES.items():
        chunks = []
        for path, a, b, anchor in parts:
            lines = open(os.path.join(ref, path), encoding='utf-8', errors='replace').read().split('\n')
            if anchor and anchor not in lines[a - 1]:
                raise SystemExit(f'extract_ref: {path}:{a} does not contain {anchor!r} (reference drifted?): {lines[a - 1]!r}')
            chunks.append(f'// ---- {path}:{a}-{b} (extracted at build time, not committed) ----')
            chunks.append(f'#line {a} "{os.path.join(ref, path)}"')
            chunks += lines[a - 1:b]
        with open(os.path.join(out, name + '.inc'), 'w') as f:
            f.write('\n'.join(chunks) + '\n')
    print(f'extract_ref: wrote {len(RANGES)} files to {out}')


if __name__ == '__main__':
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'gen')
    extract(ref, out)
