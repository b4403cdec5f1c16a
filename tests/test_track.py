"""Tracking-time projection and two matcher loops, pinned BY EXECUTION of the reference's own functions (oracle/_ref/ref_match
ops 4-7: src/Frame.cc:1371-1499 isInFrustum x2, src/MapPoint.cc:371-415 / src/MapLine.cpp:537-558 PredictScale and distance
invariance, src/ORBmatcher.cc:412-529 SearchForInitialization, src/LSDmatcher.cpp:968-1108 FrameBFMatchNew + mutualOverlap; pulled out
at build time by oracle/extract_ref.py, compiled against stand-in Frame / MapPoint / MapLine types):

  CPU: oracle (oracle/track_oracle.cpp) == executed reference (live, or the committed fixture tests/golden/track_ref.npz);
       the PredictScale thresholds the library derives from the host's logf reproduce ceil(logf(r) / L) exactly;
       the OpenCV float-matrix rules the restatements assume are checked against cv2 when it is importable.
  GPU: hvo_proj_frustum_points / hvo_lproj_frustum_lines / hvo_proj_search_initialization / hvo_match_lines_epipolar and the fused
       hvo_*_search_local_map == oracle == reference.

Bar: bit-exact (floats compared by their bytes)."""
import os

import numpy as np
import pytest

import oracle

BOUNDS = (0.0, 0.0, 640.0, 480.0)
SF = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'track_ref.npz')
_golden = np.load(GOLDEN) if os.path.exists(GOLDEN) else None
RECORD = {}


def _ref(key, run):
    """Executed reference's result for `key`: live when oracle/_ref/ref_match exists (and checked against the fixture), else the fixture."""
    live = run() if (oracle.MATCH_EXE[0] or oracle.ref_bin('ref_match') is not None) else None
    if live is not None:
        live = live if isinstance(live, tuple) else (live,)
        for j, v in enumerate(live):
            v = np.asarray(v)
            RECORD[f'{key}_{j}'] = v.view(np.uint8) if v.dtype.names else v
            if _golden is not None and f'{key}_{j}' in _golden:
                assert RECORD[f'{key}_{j}'].tobytes() == _golden[f'{key}_{j}'].tobytes(), f'fixture differs from the live reference: {key}_{j}'
        return live
    if _golden is None or f'{key}_0' not in _golden:
        pytest.skip('neither oracle/_ref/ref_match nor tests/golden/track_ref.npz is available')
    out = []
    j = 0
    while f'{key}_{j}' in _golden:
        out.append(_golden[f'{key}_{j}']); j += 1
    return tuple(out)


def _same(a, b):
    return np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()


# ---------------------------------------------------------------------------------------------------------------------------
# scenes
# ---------------------------------------------------------------------------------------------------------------------------
def _pose(rng):
    a = rng.uniform(-0.15, 0.15, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
    return (Rz @ Ry @ Rx).astype(np.float32), rng.uniform(-0.3, 0.3, 3).astype(np.float32)


def _cam(hvo_or_none, rng, fy_sign=1.0):
    R, t = _pose(rng)
    import hvo_b200
    return hvo_b200.frustum_cam(R, t, 535.4, fy_sign * 539.2, 320.1, 247.6, 40.0, BOUNDS, 1.2, 8), R, t


def _world_points(rng, R, t, n):
    """points spread so that every exit of isInFrustum is taken: behind the camera, outside the image, too near / far, grazing view"""
    z = rng.uniform(-1.0, 8.0, n)
    u = rng.uniform(-150, 790, n); v = rng.uniform(-120, 600, n)
    Pc = np.stack([(u - 320.1) * z / 535.4, (v - 247.6) * z / 539.2, z], 1)
    Pw = (Pc - t.astype(np.float64)) @ R.astype(np.float64)        # R^T (Pc - t)
    return Pw


def _point_batch(rng, R, t, n):
    Pw = _world_points(rng, R, t, n)
    Ow = -(R.astype(np.float64).T @ t.astype(np.float64))
    view = Pw - Ow
    d = np.linalg.norm(view, axis=1) + 1e-9
    nrm = view / d[:, None] + rng.normal(0, 0.6, (n, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    pts = np.zeros(n, oracle.MAP_POINT_DTYPE)
    pts['pos'] = Pw.astype(np.float32); pts['normal'] = nrm.astype(np.float32)
    mx = d * rng.uniform(0.6, 5.0, n)
    pts['max_distance'] = mx.astype(np.float32)
    pts['min_distance'] = (mx / SF[7] * rng.uniform(0.5, 1.5, n)).astype(np.float32)
    # exact powers of the scale factor: ratios that sit on the PredictScale boundaries
    k = rng.randint(0, 8, n // 8)
    pts['max_distance'][:n // 8] = (d[:n // 8].astype(np.float32) * SF[k]).astype(np.float32)
    return pts


def _line_batch(rng, R, t, n):
    a = _world_points(rng, R, t, n)
    b = a + rng.normal(0, 0.4, (n, 3))
    Ow = -(R.astype(np.float64).T @ t.astype(np.float64))
    mid = 0.5 * (a + b) - Ow
    d = np.linalg.norm(mid, axis=1) + 1e-9
    nrm = mid / d[:, None] + rng.normal(0, 0.6, (n, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    ml = np.zeros(n, oracle.MAP_LINE_DTYPE)
    ml['pos'][:, :3] = a; ml['pos'][:, 3:] = b; ml['normal'] = nrm
    dirv = b - a
    ml['dir'] = dirv / np.linalg.norm(dirv, axis=1)[:, None]
    mx = d * rng.uniform(0.6, 5.0, n)
    ml['max_distance'] = mx.astype(np.float32)
    ml['min_distance'] = (mx / SF[7] * rng.uniform(0.5, 1.5, n)).astype(np.float32)
    return ml


def _init_scene(synth, seed):
    rng = np.random.RandomState(seed)
    k1, d1 = oracle.OrbOracle().extract(synth.frame('S1', 0)[0])
    k2, d2 = oracle.OrbOracle().extract(synth.frame('S1', 1 + seed)[0])
    F1 = dict(keys_un=k1, desc=d1, bounds=BOUNDS)
    F2 = dict(keys_un=k2, desc=d2, bounds=BOUNDS)
    prev = np.stack([k1['x'], k1['y']], 1).astype(np.float32) + rng.normal(0, 2.0, (len(k1), 2)).astype(np.float32)
    return F1, F2, prev


def _line_scene(synth, seed):
    rng = np.random.RandomState(seed)
    kl1, ld1, _ = oracle.line_extract(synth.frame('S1', 0)[0], n_features=200)
    kl2, ld2, lv2 = oracle.line_extract(synth.frame('S1', 1 + seed)[0], n_features=200)
    # a fundamental matrix of a small motion (row-major, float): F = K^-T [t]x R K^-1
    R, t = _pose(rng)
    K = np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1]])
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]], np.float64)
    F = np.linalg.inv(K).T @ tx @ R.astype(np.float64) @ np.linalg.inv(K)
    F = (F / np.abs(F).max()).astype(np.float32)
    # a few planted near-duplicates so that the ratio and TH gates are both exercised
    ld1 = ld1.copy()
    m = min(len(ld1), len(ld2), 40)
    ld1[:m] = ld2[:m]
    flips = rng.randint(0, 256, (m, 3))
    for i in range(m):
        for b in flips[i][: rng.randint(0, 4)]:
            ld1[i, b // 8] ^= np.uint8(1 << (b % 8))
    kl1 = kl1.copy()
    for f in ('startPointX', 'startPointY', 'endPointX', 'endPointY'):
        kl1[f][:m] = kl2[f][:m] + rng.normal(0, 1.5, m).astype(np.float32)
    return kl1, ld1, kl2, ld2, lv2, F


# ---------------------------------------------------------------------------------------------------------------------------
# CPU: oracle == executed reference
# ---------------------------------------------------------------------------------------------------------------------------
def _frustum_cases():
    for seed, sign, limit in ((0, 1.0, 0.5), (1, -1.0, 0.5), (2, 1.0, 0.2)):
        rng = np.random.RandomState(10 + seed)
        cam, R, t = _cam(None, rng, sign)
        yield seed, cam, _point_batch(rng, R, t, 4000), _line_batch(rng, R, t, 3000), limit


def test_oracle_frustum_equals_reference():
    for seed, cam, pts, ml, limit in _frustum_cases():
        (rp,) = _ref(f'fpt{seed}', lambda: oracle.ref_frustum_points(cam, pts, limit))
        op = oracle.frustum_points(cam, pts, limit)
        rp = np.frombuffer(np.ascontiguousarray(rp).tobytes(), oracle.TRACK_POINT_DTYPE)
        assert 300 < rp['in_view'].sum() < len(pts) - 300, 'scene does not exercise both outcomes'
        assert np.array_equal(op['in_view'], rp['in_view'])
        v = rp['in_view'] != 0
        assert _same(op[v], rp[v])
        assert len(np.unique(rp['level'][v])) == 8
        (rl,) = _ref(f'fln{seed}', lambda: oracle.ref_frustum_lines(cam, ml, limit))
        ol = oracle.frustum_lines(cam, ml, limit)
        rl = np.frombuffer(np.ascontiguousarray(rl).tobytes(), oracle.TRACK_LINE_DTYPE)
        assert 100 < rl['in_view'].sum() < len(ml) - 100
        assert np.array_equal(ol['in_view'], rl['in_view'])
        v = rl['in_view'] != 0
        assert _same(ol[v], rl[v])


def test_oracle_search_initialization_equals_reference(synth):
    for seed, window, ratio, ori in ((0, 100, 0.9, True), (1, 30, 0.9, True), (2, 100, 0.7, False)):
        F1, F2, prev = _init_scene(synth, seed)
        nm_r, m12_r, prev_r = _ref(f'init{seed}', lambda: oracle.ref_search_initialization(F1, F2, prev, window, ratio, ori))
        nm, m12, prev_o, acc = oracle.search_initialization(F1, F2, prev, window, ratio, ori)
        assert int(nm_r) == nm and np.array_equal(m12_r, m12) and _same(prev_r, prev_o)
        assert nm > 20
        assert (acc >= 0).sum() > (m12 >= 0).sum() or not ori      # take-overs / histogram culling happened


def test_oracle_lines_epipolar_equals_reference(synth):
    total = 0
    for seed, TH, ratio in ((0, 50.0, 0.95), (1, 80.0, 0.8)):
        kl1, ld1, kl2, ld2, lv2, F = _line_scene(synth, seed)
        (ref,) = _ref(f'epi{seed}', lambda: oracle.ref_lines_epipolar(ld1, kl1, ld2, kl2, lv2, F, TH, ratio))
        got = oracle.lines_epipolar(ld1, kl1, ld2, kl2, lv2, F, TH, ratio)
        assert np.array_equal(ref, got)
        total += int((got >= 0).sum())
    assert total > 5


def test_oracle_distinctive_equals_reference(synth):
    """MapPoint / MapLine::ComputeDistinctiveDescriptors (src/MapPoint.cc:240-305, src/MapLine.cpp:331-396) executed: the kept descriptor and,
    where it is unique inside the element, its index."""
    from test_match import _distinctive_groups
    desc, off = _distinctive_groups(synth, ngroups=200)
    for lines in (False, True):
        ri, rd = _ref(f'dist{int(lines)}', lambda: oracle.ref_distinctive(desc, off, lines))
        bi, bm = oracle.distinctive(desc, off)
        for g in range(len(off) - 1):
            if off[g + 1] == off[g]:
                assert bi[g] == -1 and ri[g] == -1
            else:
                assert np.array_equal(desc[off[g] + bi[g]], rd[g]), g
                assert ri[g] == bi[g] or np.array_equal(desc[off[g] + ri[g]], desc[off[g] + bi[g]])   # duplicates: the first identical row


def test_oracle_search_by_bow_equals_reference(hvo, synth):
    """ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (src/ORBmatcher.cc:162-293) executed with Thirdparty/DBoW2's own FeatureVector: the
    mirror (query order, rotation histogram) on top of the oracle's candidate search must give the reference's matches."""
    from test_projection import _bow_scenario
    import test_ref_match as trm

    class OraclePMc(trm.OraclePM):
        def search_candidates(self, q, t, off, cand, th, ratio):
            return oracle.search_candidates(q, t, off, cand, th, ratio)
    for seed, ratio, ori in ((0, 0.7, True), (1, 0.9, True), (2, 0.75, False)):
        KF, F = _bow_scenario(synth, seed)
        nm_r, match_r = _ref(f'bow{seed}', lambda: oracle.ref_search_by_bow(KF, F, ratio, ori))
        m = hvo.ORBmatcher.__new__(hvo.ORBmatcher)
        m.mfNNratio, m.mbCheckOrientation, m._pm = float(ratio), bool(ori), OraclePMc()
        nm, match = m.SearchByBoW(KF, F)
        assert nm == int(nm_r) and np.array_equal(match, match_r) and nm > 50


def _tri_scene(synth, seed):
    from test_projection import _bow_scenario, _fundamental
    KFa, Fb = _bow_scenario(synth, seed=2)
    rng = np.random.RandomState(12 + seed)
    k1, k2 = KFa['keys_un'], Fb['keys']
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32); sg = (sf * sf).astype(np.float32)
    ur1 = np.where(rng.rand(len(k1)) < 0.6, k1['x'] - 20, -1).astype(np.float32)
    ur2 = np.where(rng.rand(len(k2)) < 0.6, k2['x'] - 20, -1).astype(np.float32)
    KF1 = dict(desc=KFa['desc'], keys_un=k1, uright=ur1, featvec=KFa['featvec'], has_mappoint=rng.rand(len(k1)) < 0.4, scale_factors=sf, level_sigma2=sg)
    KF2 = dict(desc=Fb['desc'], keys_un=k2, uright=ur2, featvec=Fb['featvec'], has_mappoint=rng.rand(len(k2)) < 0.3, scale_factors=sf, level_sigma2=sg)
    F12 = np.asarray(_fundamental(0.004, 0.0005), np.float32)
    cam2 = np.float32([535.4, 539.2, 320.1, 247.6])
    # pKF1's camera centre seen from pKF2 (R2w = I): far to the side (epipole outside the image) or in front (epipole inside)
    Cw1 = np.float32([[-3.0, 0.01, 0.02], [0.02, -0.01, 1.0]][seed % 2])
    t2w = np.zeros(3, np.float32)
    C2 = (Cw1 + t2w).astype(np.float32)                       # R2w * Cw + t2w with R2w = I (cv::gemm: exact products, float sums)
    invz = np.float32(1.0) / C2[2]
    ex = np.float32(np.float32(np.float32(cam2[0] * C2[0]) * invz) + cam2[2]); ey = np.float32(np.float32(np.float32(cam2[1] * C2[1]) * invz) + cam2[3])
    return KF1, KF2, F12, Cw1, np.eye(3, dtype=np.float32), t2w, cam2, (ex, ey)


def _check_triangulation(hvo, synth, gpu):
    import test_ref_match as trm

    class OraclePMt(trm.OraclePM):
        def search_triangulation(self, *a):
            return oracle.search_triangulation(*a)
    total = 0
    for seed, only_stereo, ori in ((0, False, True), (1, False, True), (2, True, False)):
        KF1, KF2, F12, Cw1, R2w, t2w, cam2, epi = _tri_scene(synth, seed)
        nm_r, pairs_r = _ref(f'tri{seed}', lambda: oracle.ref_search_for_triangulation(KF1, KF2, F12, Cw1, R2w, t2w, cam2, only_stereo, ori, 0.6))
        if gpu:
            m = hvo.ORBmatcher(0.6, ori)
        else:
            m = hvo.ORBmatcher.__new__(hvo.ORBmatcher)
            m.mfNNratio, m.mbCheckOrientation, m._pm = 0.6, bool(ori), OraclePMt()
        nm, pairs = m.SearchForTriangulation(KF1, KF2, F12, epi, only_stereo)
        assert nm == int(nm_r) and np.array_equal(pairs, np.asarray(pairs_r).reshape(-1, 2))
        total += nm
    assert total > 20


def test_oracle_search_for_triangulation_equals_reference(hvo, synth):
    """ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:668-836) + CheckDistEpipolarLine (:143-160) executed."""
    _check_triangulation(hvo, synth, gpu=False)


def _fuse_project(cam, pts, n_levels=8):
    """The projection tests of ORBmatcher::Fuse (src/ORBmatcher.cc:866-900) with the reference's float rules (cv::gemm float sums, cv::norm /
    Mat::dot in double); PredictScale(dist, pKF) through the library's logf thresholds.  Returns (ok, u, v, ur, level)."""
    import hvo_b200
    f32 = np.float32
    R = cam['Rcw'].reshape(3, 3); t = cam['tcw']; O = cam['Ow']
    P = pts['pos']
    Pc = np.zeros((len(P), 3), f32)
    for i in range(3):
        acc = np.zeros(len(P), f32)
        for k in range(3):
            acc = (acc + (R[i, k] * P[:, k]).astype(f32)).astype(f32)
        Pc[:, i] = (acc + t[i]).astype(f32)
    with np.errstate(divide='ignore', invalid='ignore'):
        invz = (f32(1) / Pc[:, 2]).astype(f32)
        x = (Pc[:, 0] * invz).astype(f32); y = (Pc[:, 1] * invz).astype(f32)
        u = ((cam['fx'] * x).astype(f32) + cam['cx']).astype(f32); v = ((cam['fy'] * y).astype(f32) + cam['cy']).astype(f32)
        ur = (u - (cam['bf'] * invz).astype(f32)).astype(f32)
        ok = ~(Pc[:, 2] < 0) & (u >= cam['min_x']) & (u < cam['max_x']) & (v >= cam['min_y']) & (v < cam['max_y'])
        PO = (P - O[None, :]).astype(f32)
        d64 = np.sqrt(PO[:, 0].astype(np.float64) ** 2 + PO[:, 1].astype(np.float64) ** 2 + PO[:, 2].astype(np.float64) ** 2)
        dist = d64.astype(f32)
        ok &= ~(dist < (f32(0.8) * pts['min_distance']).astype(f32)) & ~(dist > (f32(1.2) * pts['max_distance']).astype(f32))
        Pn = pts['normal']
        dot = (PO[:, 0].astype(np.float64) * Pn[:, 0] + PO[:, 1].astype(np.float64) * Pn[:, 1]) + PO[:, 2].astype(np.float64) * Pn[:, 2]
        ok &= ~(dot < 0.5 * dist.astype(np.float64))
        ratio = (pts['max_distance'] / dist).astype(f32)
    thr = hvo_b200.predict_scale_thresholds(cam['log_scale_factor'], 0, n_levels - 1)
    level = (ratio[:, None] >= thr[None, :]).sum(1).astype(np.int32)
    return ok, u, v, ur, level


def _check_fuse(hvo, synth, gpu):
    import test_ref_match as trm
    from test_ref_match import _point_scene

    class OraclePMf(trm.OraclePM):
        def set_level_sigma(self, s):
            self.sig = np.ascontiguousarray(s, np.float32)

        def search(self, q, qd, claimed, mode, th, ratio):
            assert mode == 2
            return self.with_origin(lambda: oracle.search_fuse(self.k, self.ur, self.d, self.b, q, qd, self.sig, th))
    total = 0
    for seed, th in ((0, 3.0), (1, 5.0), (2, 4.0)):
        rng = np.random.RandomState(80 + seed)
        F, _, _, _ = _point_scene(synth, seed)            # key frame = the point frame of the projection tests
        k0, n0 = F['keys_un'], len(F['keys_un'])
        cam, R, t = _cam(hvo, rng)
        bounds = BOUNDS
        if seed == 2:
            # a distorted camera's image bounds are not integers: the key frame truncates them (include/KeyFrame.h:249-252) for IsInImage and
            # for the origin of its window lookups (src/KeyFrame.cc:627-666), in a grid its Frame assigned with the float bounds
            bounds = (-13.6, -9.3, 655.2, 492.7)
            F = dict(F); F['bounds'] = bounds
            cam = np.array(cam, copy=True)
            for k, b in zip(('min_x', 'min_y', 'max_x', 'max_y'), bounds):
                cam[k] = np.float32(int(b))
        z = rng.uniform(0.8, 5.0, n0)
        Pc = np.stack([(k0['x'] + rng.normal(0, 1, n0) - 320.1) * z / 535.4, (k0['y'] + rng.normal(0, 1, n0) - 247.6) * z / 539.2, z], 1)
        pts = np.concatenate([_point_batch(rng, R, t, n0), _point_batch(rng, R, t, 400)])
        pts['pos'][:n0] = ((Pc - t.astype(np.float64)) @ R.astype(np.float64)).astype(np.float32)
        Ow = -(R.astype(np.float64).T @ t.astype(np.float64))
        view = pts['pos'][:n0].astype(np.float64) - Ow
        dist = np.linalg.norm(view, axis=1)
        pts['normal'][:n0] = (view / dist[:, None]).astype(np.float32)
        pts['max_distance'][:n0] = (dist * SF[np.clip(k0['octave'], 0, 7)]).astype(np.float32)
        pts['min_distance'][:n0] = pts['max_distance'][:n0] / SF[7]
        M = len(pts)
        pdesc = np.concatenate([F['desc'], rng.randint(0, 256, (400, 32)).astype(np.uint8)])
        flips = rng.randint(0, 256, M)
        pdesc[np.arange(M), flips // 8] ^= (1 << (flips % 8)).astype(np.uint8)
        present = rng.rand(M) > 0.05; bad = rng.rand(M) < 0.05; nobs = rng.randint(0, 6, M); in_kf = rng.rand(M) < 0.1
        kf_obs = rng.randint(0, 6, n0)
        inv = (np.float32(1.0) / (SF * SF)).astype(np.float32)
        cam = cam.reshape(())
        ref = _ref(f'fuse{seed}', lambda: oracle.ref_fuse(F, kf_obs, [cam['fx'], cam['fy'], cam['cx'], cam['cy'], cam['bf']], cam['Rcw'], cam['tcw'], cam['Ow'],
                                                         inv, cam['log_scale_factor'], 8, th, pts, pdesc, present, bad, nobs, in_kf))
        nf_r, ev = int(ref[0]), np.asarray(ref[1]).reshape(-1, 3)
        ok, u, v, ur, level = _fuse_project(cam, pts)
        use = ok & present & ~bad & ~in_kf
        sel = np.nonzero(use)[0]
        KF = dict(keys_un=k0, uright=F['uright'], desc=F['desc'], bounds=bounds, scale_factors=SF, inv_level_sigma2=inv)
        MPs = dict(u=u[sel], v=v[sel], ur=ur[sel], level=level[sel], desc=pdesc[sel])
        if gpu:
            m = hvo.ORBmatcher(0.6, True)
        else:
            m = hvo.ORBmatcher.__new__(hvo.ORBmatcher)
            m.mfNNratio, m.mbCheckOrientation, m._pm = 0.6, True, OraclePMf()
        nf, best = m.Fuse(KF, MPs, th)
        full = np.full(M, -1, np.int32); full[sel] = best
        want = np.full(M, -1, np.int32); want[ev[:, 0]] = ev[:, 1]
        assert nf == nf_r == len(ev) and np.array_equal(full, want)
        assert len(set(ev[:, 2])) == 3            # AddObservation and both Replace directions occurred
        total += nf
    assert total > 300


def test_oracle_fuse_equals_reference(hvo, synth):
    """ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th) (src/ORBmatcher.cc:838-994) executed with KeyFrame::GetFeaturesInArea / IsInImage
    (src/KeyFrame.cc:627-666, 780-783) and MapPoint::PredictScale(dist, KeyFrame*): the key-frame keypoint every map point is fused into."""
    _check_fuse(hvo, synth, gpu=False)


def test_predict_scale_thresholds_reproduce_logf(hvo):
    """level(ratio) from the thresholds == ceil(logf(ratio) / L) of the host libm for random ratios and for both float neighbours of
    every threshold (the boundaries), clamped (points) and unclamped (lines)."""
    import ctypes as C
    libm = C.CDLL('libm.so.6')
    libm.logf.restype = C.c_float; libm.logf.argtypes = [C.c_float]
    for sf in (1.2, 1.1, 2.0):
        L = np.float32(np.log(np.float64(np.float32(sf))))
        lo, n = -32, 96
        thr = hvo.predict_scale_thresholds(L, lo, n)
        assert np.all(np.diff(thr.astype(np.float64)) > 0)

        def direct(r):
            return int(np.ceil(np.float32(np.float32(libm.logf(float(r))) / L)))
        rng = np.random.RandomState(3)
        rs = np.concatenate([np.exp(rng.uniform(np.log(sf) * (lo + 1), np.log(sf) * (lo + n - 1), 3000)).astype(np.float32),
                             thr, np.nextafter(thr, np.float32(0)), np.nextafter(thr, np.float32(np.inf))])
        for r in rs:
            want = direct(r)
            if lo < want < lo + n:
                assert lo + int((r >= thr).sum()) == want, (sf, float(r), want)
        t8 = hvo.predict_scale_thresholds(L, 0, 7)
        for r in rs:
            want = min(max(direct(r), 0), 7)
            assert int((r >= t8).sum()) == want


def test_opencv_float_matrix_rules():
    """What the restatements assume about cv::Mat arithmetic on CV_32F, checked against cv2 itself."""
    cv2 = pytest.importorskip('cv2')
    rng = np.random.RandomState(0)
    f32 = np.float32
    for _ in range(3000):
        R = rng.randn(3, 3).astype(f32); P = (rng.randn(3, 1) * 5).astype(f32); t = rng.randn(3, 1).astype(f32)
        d = cv2.gemm(R, P, 1.0, t, 1.0)
        for i in range(3):
            s = f32(f32(f32(R[i, 0] * P[0, 0]) + f32(R[i, 1] * P[1, 0])) + f32(R[i, 2] * P[2, 0]))
            assert d[i, 0] == f32(np.float64(s) + np.float64(t[i, 0])) == f32(s + t[i, 0])
        v = (rng.randn(3, 1) * 3).astype(f32)
        assert cv2.norm(v) == np.sqrt(np.float64(v[0, 0]) ** 2 + np.float64(v[1, 0]) ** 2 + np.float64(v[2, 0]) ** 2)
        a = (rng.randn(3, 1) * 10 ** rng.uniform(-3, 3)).astype(f32); b = (rng.randn(3, 1) * 10 ** rng.uniform(-3, 3)).astype(f32)
        assert np.array_equal(cv2.addWeighted(a, 0.5, b, 0.5, 0), (a * f32(0.5) + b * f32(0.5)).astype(f32))
        # -R.t() * t (ORBmatcher.cc:308, 1009, 1366, 1505) = one gemm(R, t, -1, GEMM_1_T): off the small-matrix path, double accumulation
        g = cv2.gemm(R, t, -1.0, None, 0.0, flags=cv2.GEMM_1_T)
        for i in range(3):
            acc = 0.0
            for k in range(3):
                acc += float(R[k, i]) * float(t[k, 0])
            assert g[i, 0] == f32(acc * -1.0)
        # (A / s and s * A, ORBmatcher.cc:306-307, 1140-1141, are Mat::convertTo(alpha): a float product with (float)alpha in OpenCV's sources;
        #  convertTo has no Python binding to check against)


# ---------------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_frustum_equals_oracle_and_reference(hvo):
    pm, lpm = hvo.ProjectionMatcher(), hvo.LineProjectionMatcher()
    for seed, cam, pts, ml, limit in _frustum_cases():
        gp = pm.frustum_points(cam, pts, limit)
        op = oracle.frustum_points(cam, pts, limit)
        assert np.array_equal(gp['in_view'], op['in_view'])
        v = op['in_view'] != 0
        assert _same(gp[v], op[v])
        (rp,) = _ref(f'fpt{seed}', lambda: oracle.ref_frustum_points(cam, pts, limit))
        rp = np.frombuffer(np.ascontiguousarray(rp).tobytes(), oracle.TRACK_POINT_DTYPE)
        assert np.array_equal(gp['in_view'], rp['in_view']) and _same(gp[v], rp[v])
        gl = lpm.frustum_lines(cam, ml, limit)
        ol = oracle.frustum_lines(cam, ml, limit)
        assert np.array_equal(gl['in_view'], ol['in_view'])
        v = ol['in_view'] != 0
        assert _same(gl[v], ol[v])
        (rl,) = _ref(f'fln{seed}', lambda: oracle.ref_frustum_lines(cam, ml, limit))
        rl = np.frombuffer(np.ascontiguousarray(rl).tobytes(), oracle.TRACK_LINE_DTYPE)
        assert np.array_equal(gl['in_view'], rl['in_view']) and _same(gl[v], rl[v])


@pytest.mark.gpu
def test_gpu_search_initialization_equals_reference(hvo, synth):
    for seed, window, ratio, ori in ((0, 100, 0.9, True), (1, 30, 0.9, True), (2, 100, 0.7, False)):
        F1, F2, prev = _init_scene(synth, seed)
        m = hvo.ORBmatcher(ratio, ori)
        nm, m12, prev_g = m.SearchForInitialization(F1, F2, prev, window)
        nm_o, m12_o, prev_o, acc_o = oracle.search_initialization(F1, F2, prev, window, ratio, ori)
        assert nm == nm_o and np.array_equal(m12, m12_o) and _same(prev_g, prev_o)
        nm_r, m12_r, prev_r = _ref(f'init{seed}', lambda: oracle.ref_search_initialization(F1, F2, prev, window, ratio, ori))
        assert int(nm_r) == nm and np.array_equal(m12_r, m12) and _same(prev_r, prev_g)


@pytest.mark.gpu
def test_gpu_lines_epipolar_equals_reference(hvo, synth):
    for seed, TH, ratio in ((0, 50.0, 0.95), (1, 80.0, 0.8)):
        kl1, ld1, kl2, ld2, lv2, F = _line_scene(synth, seed)
        got = hvo.LSDmatcher(ratio).FrameBFMatchNew(ld1, ld2, kl1, kl2, lv2, F, TH)
        assert np.array_equal(got, oracle.lines_epipolar(ld1, kl1, ld2, kl2, lv2, F, TH, ratio))
        (ref,) = _ref(f'epi{seed}', lambda: oracle.ref_lines_epipolar(ld1, kl1, ld2, kl2, lv2, F, TH, ratio))
        assert np.array_equal(ref, got)
    # degenerate sizes: no train lines / one train line -> nothing is accepted (the reference's inner loop has no iteration)
    kl1, ld1, kl2, ld2, lv2, F = _line_scene(synth, 0)
    assert np.all(hvo.LSDmatcher(0.95).FrameBFMatchNew(ld1, ld2[:1], kl1, kl2[:1], lv2[:1], F, 50.0) == -1)
    assert len(hvo.LSDmatcher(0.95).FrameBFMatchNew(ld1[:0], ld2, kl1[:0], kl2, lv2, F, 50.0)) == 0


@pytest.mark.gpu
def test_gpu_search_local_points_equals_two_step_reference_path(hvo, synth):
    """Fused isInFrustum + SearchByProjection (hvo_proj_search_local_map) == the oracle's isInFrustum followed by the oracle's sequential
    SearchByProjection on the projected fields (both pinned to the executed reference separately)."""
    k0, d0 = oracle.OrbOracle().extract(synth.frame('S1', 0)[0])
    for seed, th in ((0, 1.0), (1, 3.0)):
        rng = np.random.RandomState(40 + seed)
        cam, R, t = _cam(hvo, rng)
        n0 = len(k0)
        # map points that project onto the frame's keypoints (plus noise) and random ones
        z = rng.uniform(0.8, 5.0, n0)
        Pc = np.stack([(k0['x'] + rng.normal(0, 1, n0) - 320.1) * z / 535.4, (k0['y'] + rng.normal(0, 1, n0) - 247.6) * z / 539.2, z], 1)
        Pw = (Pc - t.astype(np.float64)) @ R.astype(np.float64)
        pts = np.concatenate([_point_batch(rng, R, t, n0), _point_batch(rng, R, t, 600)])
        pts['pos'][:n0] = Pw.astype(np.float32)
        Ow = -(R.astype(np.float64).T @ t.astype(np.float64))
        view = Pw - Ow
        dist = np.linalg.norm(view, axis=1)
        pts['normal'][:n0] = (view / dist[:, None]).astype(np.float32)
        pts['max_distance'][:n0] = (dist * SF[np.clip(k0['octave'], 0, 7)]).astype(np.float32)
        pts['min_distance'][:n0] = pts['max_distance'][:n0] / SF[7]
        M = len(pts)
        pdesc = np.concatenate([d0, rng.randint(0, 256, (600, 32)).astype(np.uint8)])
        skip = rng.rand(M) < 0.1
        has_obs = rng.rand(M) > 0.3
        F = dict(keys_un=k0, desc=d0, bounds=BOUNDS, scale_factors=SF, uright=np.where(rng.rand(n0) > 0.4, k0['x'] - rng.uniform(5, 40, n0), -1).astype(np.float32),
                 claimed=(rng.rand(n0) < 0.1), mappoint=np.full(n0, -1, np.int32))
        F2 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in F.items()}
        m = hvo.ORBmatcher(0.8)
        track, nm, match = m.SearchLocalPoints(F, cam, pts, pdesc, skip, has_obs, th)
        ot = oracle.frustum_points(cam, pts, 0.5)
        ot['in_view'][skip] = 0
        assert np.array_equal(track['in_view'], ot['in_view'])
        v = ot['in_view'] != 0
        assert _same(track[v], ot[v]) and v.sum() > 300
        MPs = dict(proj_x=ot['u'], proj_y=ot['v'], proj_xr=ot['ur'], view_cos=ot['view_cos'], level=ot['level'], in_view=v, bad=np.zeros(M, bool),
                   has_obs=has_obs, desc=pdesc)
        from test_ref_match import _orb_matcher
        nm_o, match_o = _orb_matcher(hvo, False, 0.8).SearchByProjection(F2, MPs, th)
        assert nm == nm_o and np.array_equal(match, match_o) and np.array_equal(F['mappoint'], F2['mappoint'])
        assert nm > 200


@pytest.mark.gpu
def test_gpu_search_local_lines_equals_two_step_reference_path(hvo, synth):
    import test_ref_match as trm
    for seed, th in ((0, 1.0), (1, 3.0)):
        F, MLs, *_ = trm._line_scene(synth, seed) if hasattr(trm, '_line_scene') else (None, None)
        if F is None:
            pytest.skip('line scene helper missing')
        rng = np.random.RandomState(60 + seed)
        cam, R, t = _cam(hvo, rng)
        ml = _line_batch(rng, R, t, 800)
        # lines whose projections are the frame's own key lines (plus noise)
        kl = F['keylines_un']
        n0 = min(len(kl), 150)
        for j, (fx, fy) in enumerate((('startPointX', 'startPointY'), ('endPointX', 'endPointY'))):
            z = rng.uniform(1.0, 4.0, n0)
            Pc = np.stack([(kl[fx][:n0] + rng.normal(0, 0.5, n0) - 320.1) * z / 535.4, (kl[fy][:n0] + rng.normal(0, 0.5, n0) - 247.6) * z / 539.2, z], 1)
            ml['pos'][:n0, 3 * j:3 * j + 3] = (Pc - t.astype(np.float64)) @ R.astype(np.float64)
        Ow = -(R.astype(np.float64).T @ t.astype(np.float64))
        mid = 0.5 * (ml['pos'][:n0, :3] + ml['pos'][:n0, 3:]) - Ow
        d = np.linalg.norm(mid, axis=1)
        ml['normal'][:n0] = mid / d[:, None]
        ml['max_distance'][:n0] = (d * 2).astype(np.float32); ml['min_distance'][:n0] = (d / 3).astype(np.float32)
        if F.get('lines3d') is not None:
            l3 = np.asarray(F['lines3d'], np.float64).reshape(-1, 6)
            ml['dir'][:n0] = l3[:n0, :3] - l3[:n0, 3:]
        M = len(ml)
        ldesc = rng.randint(0, 256, (M, 32)).astype(np.uint8)
        ldesc[:n0] = F['ldesc'][:n0]
        skip = rng.rand(M) < 0.1
        has_obs = rng.rand(M) > 0.3
        F2 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in F.items()}
        m = hvo.LSDmatcher(0.95)
        track, nm, match = m.SearchLocalLines(F, cam, ml, ldesc, skip, has_obs, th)
        ot = oracle.frustum_lines(cam, ml, 0.5)
        ot['in_view'][skip] = 0
        assert np.array_equal(track['in_view'], ot['in_view'])
        v = ot['in_view'] != 0
        assert _same(track[v], ot[v]) and v.sum() > 50
        MLs2 = dict(proj_x1=ot['x1'], proj_y1=ot['y1'], proj_x2=ot['x2'], proj_y2=ot['y2'], view_cos=ot['view_cos'], level=ot['level'], in_view=v,
                    bad=np.zeros(M, bool), has_obs=has_obs, desc=ldesc, world_vector=ml['dir'])
        nm_o, match_o = trm._lsd_matcher(hvo, False, 0.95).SearchByProjection(F2, MLs2, True, th)
        assert nm == nm_o and np.array_equal(match, match_o) and np.array_equal(F['mapline'], F2['mapline'])


@pytest.mark.gpu
def test_gpu_distinctive_and_search_by_bow_equal_reference(hvo, synth):
    from test_match import _distinctive_groups
    from test_projection import _bow_scenario
    desc, off = _distinctive_groups(synth, ngroups=200)
    bf = hvo.BFMatcherHamming()
    bi, bm = bf.distinctive(desc, off)
    for lines in (False, True):
        ri, rd = _ref(f'dist{int(lines)}', lambda: oracle.ref_distinctive(desc, off, lines))
        for g in range(len(off) - 1):
            if off[g + 1] == off[g]:
                assert bi[g] == -1
            else:
                assert np.array_equal(desc[off[g] + bi[g]], rd[g]), g
    for seed, ratio, ori in ((0, 0.7, True), (1, 0.9, True), (2, 0.75, False)):
        KF, F = _bow_scenario(synth, seed)
        nm_r, match_r = _ref(f'bow{seed}', lambda: oracle.ref_search_by_bow(KF, F, ratio, ori))
        nm, match = hvo.ORBmatcher(ratio, ori).SearchByBoW(KF, F)
        assert nm == int(nm_r) and np.array_equal(match, match_r)


@pytest.mark.gpu
def test_gpu_search_for_triangulation_equals_reference(hvo, synth):
    _check_triangulation(hvo, synth, gpu=True)


@pytest.mark.gpu
def test_gpu_fuse_equals_reference(hvo, synth):
    _check_fuse(hvo, synth, gpu=True)
