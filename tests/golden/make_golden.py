"""Generates the committed golden fixtures.  Run in the build container (needs cv2 4.13.0 and, for the ORB
fixture, oracle/_ref/ref_orb = the reference's own src/ORBextractor.cc built by oracle/Makefile):

    python tests/golden/make_golden.py

prims_cv2.npz  inputs + outputs of the OpenCV primitives the reference calls (cv2 4.13.0 is the pin)
orb_ref.npz    inputs + outputs of the reference's ORBextractor.cc on small frames
lpvo_cv2.npz   inputs + outputs of the two cv2 primitives inside Manhattan::computeNormalsLPVO (Manhattan.cpp:312-326, :381):
               cv2.integral (CV_32F -> CV_64F) and cv2.normalize of 3x1 double vectors
lsd_cv2.npz    inputs + outputs of cv2.createLineSegmentDetector().detect (what LSDDetector_custom.cpp:149,158 calls) and of
               the two cv2 primitives inside it (GaussianBlur 7x7 sigma 0.75, resize 0.8 INTER_LINEAR_EXACT)
peac_ref.npz   outputs of the reference's own plane extractor (src/PlaneExtractor.cpp + include/peac/*.hpp compiled unmodified
               into oracle/_ref/ref_peac) on synthetic depth frames: initial 10x10 blocks, extracted planes, membership image.
               Inputs are the seeded synth frames (cfg, index); a CRC of every input depth image is stored with them.
lines_ref.npz  outputs of the reference's own line front-end (oracle/_ref/ref_lines: vendored LSDDetector_custom.cpp whole, the
               LBD functions of binary_descriptor_custom.cpp, LINEextractor::operator() and Frame::cullingLine extracted at build
               time) on synthetic frames: KeyLines + LBD + line functions before and after cullingLine.
bow_ref.npz    a vocabulary built by the reference's own DBoW2 (TemplatedVocabulary::create, k = 10, L = 4, seeded) from synthetic ORB
               descriptors, and its transform(features, BowVector, FeatureVector, levelsup) of six frames (oracle/_ref/ref_bow)
track_ref.npz  outputs of the reference's own Frame::isInFrustum x2, ORBmatcher::SearchForInitialization and LSDmatcher::FrameBFMatchNew
               (oracle/_ref/ref_match ops 4-7) on the scenes of tests/test_track.py
lpvo_ref.npz   outputs of the reference's own Manhattan::computeNormalsLPVO (oracle/_ref/ref_lpvo, cv::Rect body of removeMatRow / removeMatCol)
kf_ref.npz     outputs of the reference's own key-frame matchers (ref_match ops 12-18: ORBmatcher::SearchByProjection(KF, Scw, ...) / (Cur, KF, ...),
               Fuse(KF, Scw, ...), SearchBySim3, SearchByBoW(KF1, KF2), LSDmatcher::FrameBFMatch / match / SearchDouble x2 / SearchByDescriptor)
               on the scenes of tests/test_ref_kf.py
match_ref.npz  outputs of the reference's own windowed matchers (oracle/_ref/ref_match: Frame grids + GetFeaturesInArea*, the two
               ORBmatcher::SearchByProjection and the two LSDmatcher::SearchByProjection, extracted at build time) on the scenes of
               tests/test_ref_match.py: grids, candidate lists, final assignments, match counts.
"""
import os
import zlib
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hvo_b200  # noqa: E402
from hvo_b200 import synth  # noqa: E402
import oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def prims():
    assert cv2.__version__ == '4.13.0', cv2.__version__
    cv2.setNumThreads(1)
    img = synth.noise_frame(160, 120, 11)
    out = dict(img=img, cv2_version=np.array(cv2.__version__))
    sizes = [(133, 100), (111, 83), (93, 69), (77, 58), (150, 113), (81, 61)]
    cur = img
    for i, (w, h) in enumerate(sizes[:4]):  # cascade, as the pyramid does
        cur = cv2.resize(cur, (w, h), interpolation=cv2.INTER_LINEAR)
        out[f'resize_cascade_{i}'] = cur
    for i, (w, h) in enumerate(sizes[4:]):
        out[f'resize_single_{i}'] = cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)
    out['blur7'] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    out['blur5'] = cv2.GaussianBlur(img, (5, 5), 1, 1, borderType=cv2.BORDER_REFLECT_101)
    out['sobel_dx'] = cv2.Sobel(img, cv2.CV_16S, 1, 0, ksize=3)
    out['sobel_dy'] = cv2.Sobel(img, cv2.CV_16S, 0, 1, ksize=3)
    r = np.random.RandomState(5)
    rois = [(0, 0, 160, 120)] + [(r.randint(0, 110), r.randint(0, 80), r.randint(7, 44), r.randint(7, 40)) for _ in range(40)]
    out['fast_rois'] = np.array(rois, np.int32)
    for thr in (20, 7):
        det = cv2.FastFeatureDetector_create(threshold=thr, nonmaxSuppression=True)
        allk, offs = [], [0]
        for (x, y, w, h) in rois:
            ks = det.detect(img[y:y + h, x:x + w])
            allk += [[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in ks]
            offs.append(len(allk))
        out[f'fast_{thr}_kps'] = np.array(allk, np.int32).reshape(-1, 3)
        out[f'fast_{thr}_offs'] = np.array(offs, np.int32)
    yy = (r.randn(4000) * 500).astype(np.float32)
    xx = (r.randn(4000) * 500).astype(np.float32)
    yy[:8] = [0, 0, 1, -1, 5, -5, 0, 3]
    xx[:8] = [0, 1, 0, 0, 5, -5, -2, -3]
    out['atan2_y'], out['atan2_x'] = yy, xx
    out['atan2'] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(yy, xx)], np.float32)
    # brute-force Hamming knn-2 (what LSDmatcher uses: src/LSDmatcher.cpp:811-812), with planted ties
    q = r.randint(0, 256, (60, 32)).astype(np.uint8)
    t = r.randint(0, 256, (300, 32)).astype(np.uint8)
    t[17] = t[5]; t[200] = q[3]; t[201] = q[3]; t[77] = q[9]
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(q, t, k=2)
    out['knn_q'], out['knn_t'] = q, t
    out['knn_idx'] = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    out['knn_dist'] = np.array([[a.distance, b.distance] for a, b in m], np.float32)
    np.savez_compressed(os.path.join(OUT, 'prims_cv2.npz'), **out)
    print('prims_cv2.npz written')


def orb():
    if oracle.ref_orb_path() is None:
        print('oracle/_ref/ref_orb missing: run make -C oracle first')
        return
    cases = []
    g1, _ = synth.frame('S1', 0)
    g2, _ = synth.frame('S2', 1)
    cases.append(('s1_crop', g1[100:340, 150:470].copy(), dict(nfeatures=300, scale_factor=1.2, nlevels=6, ini_th=20, min_th=7)))
    cases.append(('s2_crop', g2[120:360, 200:520].copy(), dict(nfeatures=300, scale_factor=1.2, nlevels=6, ini_th=20, min_th=7)))
    cases.append(('noise', synth.noise_frame(256, 192, 3), dict(nfeatures=500, scale_factor=1.2, nlevels=5, ini_th=20, min_th=7)))
    out = {}
    for name, img, p in cases:
        (kps, desc), = oracle.ref_orb_extract(img[None], **p)
        out[name + '_img'] = img
        out[name + '_params'] = np.array([p['nfeatures'], p['nlevels'], p['ini_th'], p['min_th']], np.int32)
        out[name + '_scale'] = np.float32(p['scale_factor'])
        out[name + '_kps'] = kps
        out[name + '_desc'] = desc
        print(name, len(kps))
    np.savez_compressed(os.path.join(OUT, 'orb_ref.npz'), **out)
    print('orb_ref.npz written')


def lsd():
    assert cv2.__version__ == '4.13.0', cv2.__version__
    cv2.setNumThreads(1)
    det = cv2.createLineSegmentDetector()
    g1, _ = synth.frame('S1', 3)
    g2, _ = synth.frame('S2', 4)
    cases = [('s1_crop', g1[60:300, 100:420].copy()), ('s2_crop', g2[100:340, 160:480].copy()),
             ('s1_odd', g1[31:232, 17:350].copy()), ('noise', synth.noise_frame(160, 120, 7)),
             ('flat', np.full((120, 160), 77, np.uint8))]
    out = dict(cv2_version=np.array(cv2.__version__))
    for name, img in cases:
        seg = det.detect(img)[0]
        seg = np.zeros((0, 4), np.float32) if seg is None else seg.reshape(-1, 4)
        out[name + '_img'] = img
        out[name + '_segments'] = seg
        out[name + '_scaled'] = cv2.resize(cv2.GaussianBlur(img, (7, 7), 0.75), None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT)
        print(name, img.shape, len(seg))
    np.savez_compressed(os.path.join(OUT, 'lsd_cv2.npz'), **out)
    print('lsd_cv2.npz written')


def lpvo():
    assert cv2.__version__ == '4.13.0', cv2.__version__
    r = np.random.RandomState(21)
    img = (r.randn(48, 64) * r.choice([1e-3, 1.0, 50.0], (48, 64))).astype(np.float32)
    img[r.rand(48, 64) < 0.3] = 0
    vec = r.randn(256, 3) * r.choice([1e-9, 1e-3, 1.0], (256, 1))
    vec[0] = 0
    out = dict(img=img, integral=cv2.integral(img)[1:, 1:], vec=vec,
               normalized=np.stack([cv2.normalize(v.reshape(3, 1), None).ravel() for v in vec]), cv2_version=np.array(cv2.__version__))
    assert out['integral'].dtype == np.float64
    np.savez_compressed(os.path.join(OUT, 'lpvo_cv2.npz'), **out)
    print('lpvo_cv2.npz written')


PEAC_CASES = [('S1', 0), ('S1', 9), ('S2', 3), ('S3', 1)]


def peac():
    if oracle.ref_bin('ref_peac') is None:
        print('oracle/_ref/ref_peac missing: run make -C oracle first')
        return
    out = dict(cases=np.array([f'{c}:{i}' for c, i in PEAC_CASES]))
    for cfg, idx in PEAC_CASES:
        c = synth.CONFIGS[cfg]
        _, d = synth.frame(cfg, idx)
        (r,) = oracle.ref_peac(d[None], np.float32(1.0 / c['factor']), c['fx'], c['fy'], c['cx'], c['cy'])
        k = f'{cfg}_{idx}_'
        out[k + 'depth_crc'] = np.uint32(zlib.crc32(d.tobytes()))
        # block statistics: validity / N for every block, the double-precision fit for every 4th one (fixture size)
        out[k + 'block_N'] = r['blocks']['N'].astype(np.int16)
        out[k + 'block_nouse'] = r['blocks']['nouse'].astype(np.int8)
        out[k + 'blocks_every4'] = r['blocks'][::4]
        out[k + 'planes'] = r['planes']
        # the reference leaves its visit counters (-2 .. -6) in unassigned pixels (AHCPlaneFitter.hpp:445, 468-472): consumers only
        # read labels >= 0 (refineDetails :362-368), so the fixture keeps max(label, -1)
        out[k + 'membership'] = np.maximum(r['membership'], -1).astype(np.int8).reshape(d.shape)
        print(cfg, idx, len(r['planes']), r['planes']['N'])
    np.savez_compressed(os.path.join(OUT, 'peac_ref.npz'), **out)
    print('peac_ref.npz written')


LINE_CASES = [('S1', 3), ('S1', 4), ('S2', 2), ('S3', 0)]     # S1:3 and S1:4 hold equal responses at the sort (std::sort tie order)


def lines():
    if oracle.ref_bin('ref_lines') is None:
        print('oracle/_ref/ref_lines missing: run make -C oracle first')
        return
    out = dict(cases=np.array([f'{c}:{i}' for c, i in LINE_CASES]))
    for cfg, idx in LINE_CASES:
        g, _ = synth.frame(cfg, idx)
        (r,) = oracle.ref_lines(g[None], n_features=200, cull=True)
        k = f'{cfg}_{idx}_'
        out[k + 'gray_crc'] = np.uint32(zlib.crc32(g.tobytes()))
        for name in ('keylines', 'desc', 'linevec', 'keylines2', 'desc2', 'linevec2'):
            out[k + name] = r[name]
        print(cfg, idx, len(r['keylines']), len(r['keylines2']))
    np.savez_compressed(os.path.join(OUT, 'lines_ref.npz'), **out)
    print('lines_ref.npz written')


def match():
    if oracle.ref_bin('ref_match') is None:
        print('oracle/_ref/ref_match missing: run make -C oracle first')
        return
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import test_ref_match as t
    t._golden = None
    for check in (t._check_search_by_projection, t._check_search_last, t._check_line_search, t._check_line_search_last):
        check(hvo_b200, synth, gpu=False)          # CPU leg: the mirrors on the oracle must already equal the live reference
    np.savez_compressed(os.path.join(OUT, 'match_ref.npz'), **t.RECORD)
    print('match_ref.npz written:', len(t.RECORD), 'arrays')


def track():
    if oracle.ref_bin('ref_match') is None:
        print('oracle/_ref/ref_match missing: run make -C oracle first')
        return
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import test_track as t
    t._golden = None
    t.test_oracle_frustum_equals_reference()
    t.test_oracle_search_initialization_equals_reference(synth)
    t.test_oracle_lines_epipolar_equals_reference(synth)
    t.test_oracle_distinctive_equals_reference(synth)
    t.test_oracle_search_by_bow_equals_reference(hvo_b200, synth)
    t.test_oracle_search_for_triangulation_equals_reference(hvo_b200, synth)
    t.test_oracle_fuse_equals_reference(hvo_b200, synth)
    np.savez_compressed(os.path.join(OUT, 'track_ref.npz'), **t.RECORD)
    print('track_ref.npz written:', len(t.RECORD), 'arrays')


def lpvo_ref():
    if oracle.ref_bin('ref_lpvo') is None:
        print('oracle/_ref/ref_lpvo missing: run make -C oracle first')
        return
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import test_lpvo as t
    out = {}
    for cfg, idx in t.REF_CASES:
        for k in (0, 1):
            cam = t._cam(synth, cfg)
            n, z = oracle.ref_lpvo(t._holes(synth.frame(cfg, idx)[1], k), **cam)
            out[f'{cfg}_{idx}_{k}_n'] = n; out[f'{cfg}_{idx}_{k}_z'] = z
    np.savez_compressed(os.path.join(OUT, 'lpvo_ref.npz'), **out)
    print('lpvo_ref.npz written:', len(out), 'arrays')


def kf():
    if oracle.ref_bin('ref_match') is None:
        print('oracle/_ref/ref_match missing: run make -C oracle first')
        return
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import test_ref_kf as t
    t._golden = None
    for check in t.CHECKS:
        check(hvo_b200, synth, gpu=False)          # CPU leg: the mirrors on the oracle must already equal the live reference
    np.savez_compressed(os.path.join(OUT, 'kf_ref.npz'), **t.RECORD)
    print('kf_ref.npz written:', len(t.RECORD), 'arrays')


def bow():
    if oracle.ref_bin('ref_bow') is None:
        print('oracle/_ref/ref_bow missing: run make -C oracle first')
        return
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import test_bow as t
    out = t.make_golden(synth)
    print('bow_ref.npz written:', len(out), 'arrays')


if __name__ == '__main__':
    which = sys.argv[1:] or ['prims', 'orb', 'lsd', 'lpvo', 'peac', 'lines', 'match', 'track', 'bow', 'kf', 'lpvo_ref']
    if 'bow' in which:
        bow()
    if 'kf' in which:
        kf()
    if 'lpvo_ref' in which:
        lpvo_ref()
    if 'track' in which:
        track()
    if 'match' in which:
        match()
    if 'lines' in which:
        lines()
    if 'peac' in which:
        peac()
    if 'lpvo' in which:
        lpvo()
    if 'prims' in which:
        prims()
    if 'orb' in which:
        orb()
    if 'lsd' in which:
        lsd()
