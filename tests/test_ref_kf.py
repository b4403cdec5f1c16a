"""The key-frame side of ORBmatcher / LSDmatcher pinned by EXECUTING the reference (oracle/_ref/ref_match ops 12-18; fixture
tests/golden/kf_ref.npz, written by tests/golden/make_golden.py kf):

  ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)              src/ORBmatcher.cc:295-410    (loop detection)
  ORBmatcher::Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)                       src/ORBmatcher.cc:996-1121   (loop closing)
  ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)                 src/ORBmatcher.cc:1123-1351
  ORBmatcher::SearchByProjection(CurrentFrame, KeyFrame*, sAlreadyFound, th, ORBdist)  src/ORBmatcher.cc:1499-1628  (relocalisation)
  ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12)                                     src/ORBmatcher.cc:531-666
  LSDmatcher::FrameBFMatch / match / SearchDouble x2 / SearchByDescriptor              src/LSDmatcher.cpp:522-559, 803-966, 1110-1135
  with KeyFrame::GetFeaturesInArea / IsInImage (src/KeyFrame.cc:627-666, 780-783: INTEGER image origin), KeyFrame::GetMapPoints,
  MapPoint::PredictScale / GetIndexInKeyFrame.

CPU leg: the mirrors (hvo.ORBmatcher / hvo.LSDmatcher) on top of the CPU oracle, with the projection tests restated in numpy under the
reference's cv::Mat float rules, must equal the executed reference.  GPU leg: the same mirrors on the CUDA library."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle  # noqa: E402
import test_ref_match as trm  # noqa: E402
import test_track as tt  # noqa: E402
from test_track import SF, _cam, _point_batch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'kf_ref.npz')
_golden = np.load(GOLDEN) if os.path.exists(GOLDEN) else None
RECORD = {}
f32 = np.float32
# image bounds: the undistorted TUM3 image, and a distorted camera's (non-integer: the key frame truncates them, include/KeyFrame.h:249-252)
BOUNDS_INT = (0.0, 0.0, 640.0, 480.0)
BOUNDS_FRAC = (-13.6, -9.3, 655.2, 492.7)


def _ref(key, run):
    """Executed reference's result for `key`: live when oracle/_ref/ref_match exists (and checked against the fixture), else the fixture."""
    live = run() if (oracle.MATCH_EXE[0] or oracle.ref_bin('ref_match') is not None) else None
    if live is not None:
        live = live if isinstance(live, tuple) else (live,)
        for j, v in enumerate(live):
            RECORD[f'{key}_{j}'] = np.asarray(v)
            if _golden is not None and f'{key}_{j}' in _golden:
                assert RECORD[f'{key}_{j}'].tobytes() == _golden[f'{key}_{j}'].tobytes(), f'fixture differs from the live reference: {key}_{j}'
        return live
    if _golden is None or f'{key}_0' not in _golden:
        pytest.skip('neither oracle/_ref/ref_match nor tests/golden/kf_ref.npz is available')
    out, j = [], 0
    while f'{key}_{j}' in _golden:
        out.append(_golden[f'{key}_{j}']); j += 1
    return tuple(out)


class OraclePMk(trm.OraclePM):
    """hvo.ProjectionMatcher served by the CPU oracle (with the key frame's integer window origin, see trm.OraclePM)"""

    def search_candidates(self, q, t, off, cand, th, ratio):
        return oracle.search_candidates(q, t, off, cand, th, ratio)


def _matcher(hvo, gpu, ratio, ori=True):
    if gpu:
        return hvo.ORBmatcher(ratio, ori)
    m = hvo.ORBmatcher.__new__(hvo.ORBmatcher)
    m.mfNNratio, m.mbCheckOrientation, m._pm = float(ratio), bool(ori), OraclePMk()
    return m


# ---------------------------------------------------------------------------------------------------------------------------
# the reference's cv::Mat float rules, in numpy
# ---------------------------------------------------------------------------------------------------------------------------
def _gemm_add(R, P, t):
    """R * p + t per row of P: cv::gemm's small-matrix path, float products and sums in k order, then the float sum with t"""
    out = np.zeros((len(P), 3), f32)
    for i in range(3):
        acc = np.zeros(len(P), f32)
        for k in range(3):
            acc = (acc + (R[i, k] * P[:, k]).astype(f32)).astype(f32)
        out[:, i] = (acc + t[i]).astype(f32)
    return out


def _neg_rt_t(R, t):
    """-R.t() * t: ONE cv::gemm(R, t, -1, GEMM_1_T): products and sums in double, (float)(sum * -1)  (oracle/cvshim MatTExpr, checked against cv2)"""
    out = np.zeros(3, f32)
    for i in range(3):
        s = 0.0
        for k in range(3):
            s += float(R[k, i]) * float(t[k])
        out[i] = f32(s * -1.0)
    return out


def _norm(P):
    return np.sqrt(P[:, 0].astype(np.float64) ** 2 + P[:, 1].astype(np.float64) ** 2 + P[:, 2].astype(np.float64) ** 2).astype(f32)


def _sim3_camera(cam, Scw):
    """ORBmatcher.cc:303-308 / 1004-1009: scw = sqrt(row0 . row0) (double dot), Rcw = sRcw / scw, tcw = Scw[:3, 3] / scw (float products with
    (float)(1. / scw)), Ow = -Rcw.t() * tcw; image bounds = the key frame's integers."""
    sR = Scw[:3, :3].astype(f32)
    dot = 0.0
    for k in range(3):
        dot += float(sR[0, k]) * float(sR[0, k])
    scw = f32(np.sqrt(dot))
    f = f32(1.0 / float(scw))
    c = np.array(cam, copy=True).reshape(())
    Rcw = (sR * f).astype(f32); tcw = (Scw[:3, 3].astype(f32) * f).astype(f32)
    c['Rcw'] = Rcw.reshape(9); c['tcw'] = tcw; c['Ow'] = _neg_rt_t(Rcw, tcw)
    for k in ('min_x', 'min_y', 'max_x', 'max_y'):
        c[k] = f32(int(c[k]))
    return c


def _levels(hvo_mod, ratio, log_sf, n_levels):
    thr = hvo_mod.predict_scale_thresholds(log_sf, 0, n_levels - 1)
    return (ratio[:, None] >= thr[None, :]).sum(1).astype(np.int32)


# ---------------------------------------------------------------------------------------------------------------------------
# scenes
# ---------------------------------------------------------------------------------------------------------------------------
def _frustum_cam(rng, bounds):
    import hvo_b200
    R, t = tt._pose(rng)
    return hvo_b200.frustum_cam(R, t, 535.4, 539.2, 320.1, 247.6, 40.0, bounds, 1.2, 8).reshape(()), R, t


def _points_at(rng, keys, R, t, extra):
    """map points that project (through R, t) near the given keypoints, plus `extra` random ones; invariance distances from the octave"""
    n0 = len(keys)
    z = rng.uniform(0.8, 5.0, n0)
    Pc = np.stack([(keys['x'] + rng.normal(0, 1, n0) - 320.1) * z / 535.4, (keys['y'] + rng.normal(0, 1, n0) - 247.6) * z / 539.2, z], 1)
    pts = np.concatenate([_point_batch(rng, R, t, n0), _point_batch(rng, R, t, extra)]) if extra else _point_batch(rng, R, t, n0)
    pts['pos'][:n0] = ((Pc - t.astype(np.float64)) @ R.astype(np.float64)).astype(f32)
    Ow = -(R.astype(np.float64).T @ t.astype(np.float64))
    view = pts['pos'][:n0].astype(np.float64) - Ow
    dist = np.linalg.norm(view, axis=1)
    pts['normal'][:n0] = (view / dist[:, None]).astype(f32)
    pts['max_distance'][:n0] = (dist * SF[np.clip(keys['octave'], 0, 7)]).astype(f32)
    pts['min_distance'][:n0] = pts['max_distance'][:n0] / SF[7]
    return pts


def _noisy_desc(rng, desc, extra):
    d = np.concatenate([desc, rng.randint(0, 256, (extra, 32)).astype(np.uint8)]) if extra else desc.copy()
    flips = rng.randint(0, 256, len(d))
    d[np.arange(len(d)), flips // 8] ^= (1 << (flips % 8)).astype(np.uint8)
    return d


def _sim3_scene(synth, seed):
    rng = np.random.RandomState(300 + seed)
    bounds = BOUNDS_FRAC if seed % 2 else BOUNDS_INT
    F, _, k1, d1 = trm._point_scene(synth, seed)
    F = dict(F); F['bounds'] = bounds
    cam, R, t = _frustum_cam(rng, bounds)
    s = (1.0, 1.7, 0.6)[seed % 3]
    Scw = np.eye(4, dtype=f32)
    Scw[:3, :3] = (s * R.astype(np.float64)).astype(f32); Scw[:3, 3] = (s * t.astype(np.float64)).astype(f32)
    # candidates = the points seen around the NEXT frame's keypoints (so they land near, not on, the key frame's) + random ones
    pts = _points_at(rng, k1, R, t, 300)
    pdesc = _noisy_desc(rng, d1, 300)
    return F, cam, Scw, pts, pdesc, rng


def _project_sim3(hvo_mod, cam, Scw, pts, viewing_angle=True):
    c = _sim3_camera(cam, Scw)
    ok, u, v, ur, level = tt._fuse_project(c, pts)     # same tests as Fuse(pKF, vpMapPoints, th): depth, image, distance, viewing angle, level
    assert viewing_angle
    return ok, u, v, level


# ---------------------------------------------------------------------------------------------------------------------------
# op 12: SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)
# ---------------------------------------------------------------------------------------------------------------------------
def _check_projection_scw(hvo, synth, gpu):
    import hvo_b200
    total = 0
    for seed, th in ((0, 10), (1, 10), (2, 4)):
        F, cam, Scw, pts, pdesc, rng = _sim3_scene(synth, seed)
        M, N = len(pts), len(F['keys_un'])
        bad = rng.rand(M) < 0.05
        found_slot = np.full(M, -1, np.int32)
        pick = rng.choice(M, 40, replace=False)
        found_slot[pick] = rng.choice(N, 40, replace=False)
        nm_r, ids = _ref(f'scw{seed}', lambda: oracle.ref_search_by_projection_scw(F, cam, Scw, th, pts, pdesc, bad, found_slot))
        if oracle.MATCH_EXE[0]:
            continue
        ok, u, v, level = _project_sim3(hvo_b200, cam, Scw, pts)
        sel = np.nonzero(ok & ~bad & (found_slot < 0))[0]
        matched = np.asarray(F['claimed'], bool).copy()
        matched[found_slot[found_slot >= 0]] = True
        preset = matched.copy()
        KF = dict(keys_un=F['keys_un'], desc=F['desc'], bounds=F['bounds'], scale_factors=SF)
        m = _matcher(hvo, gpu, 0.75)
        nm, idx = m.SearchByProjectionSim3(KF, dict(u=u[sel], v=v[sel], level=level[sel], desc=pdesc[sel]), matched, f32(th))
        want = {int(slot): int(i) for slot, i in enumerate(ids) if i >= 0 and not preset[slot]}
        got = {int(i): int(sel[k]) for k, i in enumerate(idx) if i >= 0}
        assert nm == int(nm_r) == len(want) and got == want
        # the pre-set entries are untouched
        assert all(ids[s] == -100 - s for s in np.nonzero(np.asarray(F['claimed'], bool) & ~np.isin(np.arange(N), found_slot))[0])
        total += nm
    assert oracle.MATCH_EXE[0] or total > 150


# ---------------------------------------------------------------------------------------------------------------------------
# op 13: Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)
# ---------------------------------------------------------------------------------------------------------------------------
def _check_fuse_scw(hvo, synth, gpu):
    import hvo_b200
    total = 0
    for seed, th in ((0, 4.0), (1, 4.0), (2, 7.5)):
        F, cam, Scw, pts, pdesc, rng = _sim3_scene(synth, seed)
        M, N = len(pts), len(F['keys_un'])
        F['claimed'] = rng.rand(N) < 0.5                       # half of the key frame's keypoints hold a map point
        kf_bad = rng.rand(N) < 0.1
        bad = rng.rand(M) < 0.05
        held_slot = np.full(M, -1, np.int32)
        pick = rng.choice(M, 40, replace=False)
        held_slot[pick] = rng.choice(N, 40, replace=False)      # 40 candidates ARE map points of the key frame (spAlreadyFound)
        nf_r, rep, ev, slots = _ref(f'fscw{seed}', lambda: oracle.ref_fuse_scw(F, kf_bad, cam, Scw, th, pts, pdesc, bad, held_slot))
        if oracle.MATCH_EXE[0]:
            continue
        ok, u, v, level = _project_sim3(hvo_b200, cam, Scw, pts)
        # KeyFrame::GetMapPoints: the good map points the key frame holds (a bad candidate sitting in a slot is not in the set, but is bad anyway)
        sel = np.nonzero(ok & ~bad & (held_slot < 0))[0]
        KF = dict(keys_un=F['keys_un'], desc=F['desc'], bounds=F['bounds'], scale_factors=SF)
        m = _matcher(hvo, gpu, 0.8)
        nf, best = m.FuseSim3(KF, dict(u=u[sel], v=v[sel], level=level[sel], desc=pdesc[sel]), f32(th))
        # the bookkeeping the drop-in applies in order (:1093-1112)
        slot_id = np.where(np.asarray(F['claimed'], bool), -100 - np.arange(N), -1)
        slot_bad = kf_bad.copy()
        for i in np.nonzero(held_slot >= 0)[0]:
            slot_id[held_slot[i]] = i; slot_bad[held_slot[i]] = bad[i]
        want_rep = np.full(M, -1, np.int32); want_ev = []
        for k, b in enumerate(best):
            if b < 0:
                continue
            if slot_id[b] != -1:
                if not slot_bad[b]:
                    want_rep[sel[k]] = slot_id[b]
            else:
                want_ev.append((sel[k], b)); slot_id[b] = sel[k]; slot_bad[b] = False
        assert nf == int(nf_r) == int((best >= 0).sum())
        assert np.array_equal(want_rep, rep) and np.array_equal(np.asarray(want_ev, np.int32).reshape(-1, 2), ev) and np.array_equal(slot_id, slots)
        assert (rep >= 0).any() and (rep < -1).any() and len(ev) > 0    # replaced by an added point, by a held point, and plain additions
        total += nf
    assert oracle.MATCH_EXE[0] or total > 150


# ---------------------------------------------------------------------------------------------------------------------------
# op 14: SearchBySim3
# ---------------------------------------------------------------------------------------------------------------------------
def _sim3_pair_scene(synth, seed):
    rng = np.random.RandomState(400 + seed)
    bounds = BOUNDS_FRAC if seed % 2 else BOUNDS_INT
    F, _, kn, dn = trm._point_scene(synth, seed)
    k0, d0 = F['keys_un'], F['desc']
    # key frame 2 sees most of key frame 1's features again (moved by a pixel or two, a few descriptor bits flipped, other order) + new ones
    perm = rng.permutation(len(k0))[:int(0.7 * len(k0))]
    k1 = np.concatenate([k0[perm], kn[:300]]); d1 = np.concatenate([_noisy_desc(rng, d0[perm], 0), dn[:300]])
    k1['x'][:len(perm)] += rng.normal(0, 1.5, len(perm)).astype(f32); k1['y'][:len(perm)] += rng.normal(0, 1.5, len(perm)).astype(f32)
    k1['x'] = np.clip(k1['x'], 1, 638); k1['y'] = np.clip(k1['y'], 1, 478)
    cam1, R1, t1 = _frustum_cam(rng, bounds)
    # camera 2 = camera 1 moved a little; the Sim3 given to the search is the true relative pose, slightly off, with a scale
    a = rng.uniform(-0.01, 0.01, 3)
    dR = np.array([[1, -a[2], a[1]], [a[2], 1, -a[0]], [-a[1], a[0], 1]])
    U, _, Vt = np.linalg.svd(dR)
    dR = U @ Vt                                                # the nearest rotation
    R2 = (dR @ R1.astype(np.float64)).astype(f32); t2 = (t1 + rng.uniform(-0.02, 0.02, 3)).astype(f32)
    import hvo_b200
    cam2 = hvo_b200.frustum_cam(R2, t2, 535.4, 539.2, 320.1, 247.6, 40.0, bounds, 1.2, 8).reshape(())
    KF1 = dict(keys_un=k0, desc=d0, bounds=bounds, scale_factors=SF, uright=np.full(len(k0), -1, f32))
    KF2 = dict(keys_un=k1, desc=d1, bounds=bounds, scale_factors=SF, uright=np.full(len(k1), -1, f32))
    mp1 = dict(pts=_points_at(rng, k0, R1, t1, 0), desc=_noisy_desc(rng, d0, 0), has=rng.rand(len(k0)) < 0.8, bad=rng.rand(len(k0)) < 0.05)
    mp2 = dict(pts=_points_at(rng, k1, R2, t2, 0), desc=_noisy_desc(rng, d1, 0), has=rng.rand(len(k1)) < 0.8, bad=rng.rand(len(k1)) < 0.05)
    s12 = f32((1.0, 1.03, 0.97)[seed % 3])
    R12 = (R1.astype(np.float64) @ R2.astype(np.float64).T).astype(f32)
    t12 = (t1.astype(np.float64) - R12.astype(np.float64) @ t2.astype(np.float64) + rng.uniform(-0.005, 0.005, 3)).astype(f32)
    m12 = np.full(len(k0), -1, np.int32)
    cand = np.nonzero(mp2['has'])[0]
    pick = rng.choice(len(k0), 30, replace=False)
    m12[pick] = rng.choice(cand, 30, replace=False)           # 30 matches already known (vpMatches12 on entry)
    return KF1, cam1, mp1, KF2, cam2, mp2, m12, s12, R12, t12


def _sim3_direction(hvo_mod, mp, already, Raw, taw, sRba, tba, cam_into):
    """ORBmatcher.cc:1166-1196 / 1243-1273 for one direction; returns (selected feature indices, u, v, level)"""
    n = len(mp['pts'])
    P = mp['pts']['pos']
    Pa = _gemm_add(Raw.reshape(3, 3), P, taw)
    Pb = _gemm_add(sRba, Pa, tba)
    with np.errstate(divide='ignore', invalid='ignore'):
        invz = (1.0 / Pb[:, 2].astype(np.float64)).astype(f32)
        x = (Pb[:, 0] * invz).astype(f32); y = (Pb[:, 1] * invz).astype(f32)
        u = ((cam_into['fx'] * x).astype(f32) + cam_into['cx']).astype(f32); v = ((cam_into['fy'] * y).astype(f32) + cam_into['cy']).astype(f32)
        b = [f32(int(cam_into[k])) for k in ('min_x', 'min_y', 'max_x', 'max_y')]
        ok = np.asarray(mp['has'], bool) & ~already & ~np.asarray(mp['bad'], bool) & ~(Pb[:, 2] < 0)
        ok &= (u >= b[0]) & (u < b[2]) & (v >= b[1]) & (v < b[3])
        dist = _norm(Pb)
        ok &= ~(dist < (f32(0.8) * mp['pts']['min_distance']).astype(f32)) & ~(dist > (f32(1.2) * mp['pts']['max_distance']).astype(f32))
        ratio = (mp['pts']['max_distance'] / dist).astype(f32)
    level = _levels(hvo_mod, ratio, cam_into['log_scale_factor'], int(cam_into['n_levels']))
    sel = np.nonzero(ok)[0]
    return sel, u[sel], v[sel], level[sel]


def _check_search_by_sim3(hvo, synth, gpu):
    import hvo_b200
    total = 0
    for seed, th in ((0, 7.5), (1, 7.5), (2, 3.0)):
        KF1, cam1, mp1, KF2, cam2, mp2, m12, s12, R12, t12 = _sim3_pair_scene(synth, seed)
        nf_r, ids = _ref(f'sim3_{seed}', lambda: oracle.ref_search_by_sim3(KF1, cam1, mp1, KF2, cam2, mp2, m12, s12, R12, t12, th))
        if oracle.MATCH_EXE[0]:
            continue
        N1, N2 = len(KF1['keys_un']), len(KF2['keys_un'])
        sR12 = (R12 * f32(s12)).astype(f32)                                  # s12 * R12: convertTo(alpha), a float product
        sR21 = (R12.T * f32(1.0 / float(s12))).astype(f32)                    # (1.0 / s12) * R12.t()
        t21 = (-_gemm_add(sR21, t12[None, :], np.zeros(3, f32))[0]).astype(f32)   # -sR21 * t12: gemm(alpha = -1) on the small-matrix path
        a1 = m12 >= 0
        a2 = np.zeros(N2, bool); a2[m12[a1]] = True                           # GetIndexInKeyFrame(pKF2) of the matched KF2 map points
        s1, u1, v1, l1 = _sim3_direction(hvo_b200, mp1, a1, cam1['Rcw'], cam1['tcw'], sR21, t21, cam2)
        s2, u2, v2, l2 = _sim3_direction(hvo_b200, mp2, a2, cam2['Rcw'], cam2['tcw'], sR12, t12, cam1)
        m = _matcher(hvo, gpu, 0.75)
        nf, pairs = m.SearchBySim3(KF1, KF2, dict(index=s1, u=u1, v=v1, level=l1, desc=mp1['desc'][s1]),
                                   dict(index=s2, u=u2, v=v2, level=l2, desc=mp2['desc'][s2]), f32(th))
        want = m12.copy()
        want[pairs[:, 0]] = pairs[:, 1]
        assert nf == int(nf_r) and np.array_equal(want, ids)
        total += nf
    assert oracle.MATCH_EXE[0] or total > 100


# ---------------------------------------------------------------------------------------------------------------------------
# op 15: SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)
# ---------------------------------------------------------------------------------------------------------------------------
def _check_reloc(hvo, synth, gpu):
    import hvo_b200
    total = 0
    for seed, th, orb_dist, ori in ((0, 10.0, 100, True), (1, 3.0, 64, True), (2, 10.0, 100, False)):
        rng = np.random.RandomState(500 + seed)
        bounds = BOUNDS_FRAC if seed % 2 else BOUNDS_INT
        F, _, k1, d1 = trm._point_scene(synth, seed)
        Cur = dict(F); Cur['bounds'] = bounds
        cam, R, t = _frustum_cam(rng, bounds)
        Tcw = np.eye(4, dtype=f32); Tcw[:3, :3] = R; Tcw[:3, 3] = t
        n1 = len(k1)
        mp = dict(pts=_points_at(rng, k1, R, t, 0), desc=_noisy_desc(rng, d1, 0), has=rng.rand(n1) < 0.85, bad=rng.rand(n1) < 0.05, found=rng.rand(n1) < 0.1)
        # a few points behind the camera: the reference has no depth test here, only the image bounds of the (then mirrored) projection
        back = rng.choice(n1, 20, replace=False)
        Pc_back = mp['pts']['pos'][back].astype(np.float64) @ R.astype(np.float64).T + t
        mp['pts']['pos'][back] = ((-Pc_back - t) @ R.astype(np.float64)).astype(f32)       # R^T (-Pc - t): same (u, v), negative depth
        nm_r, ids = _ref(f'reloc{seed}', lambda: oracle.ref_search_by_projection_reloc(Cur, [535.4, 539.2, 320.1, 247.6], Tcw, cam['log_scale_factor'], 8, th,
                                                                                      orb_dist, ori, k1, mp))
        if oracle.MATCH_EXE[0]:
            continue
        P = mp['pts']['pos']
        Ow = _neg_rt_t(R, t)
        Pc = _gemm_add(R, P, t)
        with np.errstate(divide='ignore', invalid='ignore'):
            invz = (1.0 / Pc[:, 2].astype(np.float64)).astype(f32)
            u = (((cam['fx'] * Pc[:, 0]).astype(f32) * invz).astype(f32) + cam['cx']).astype(f32)      # fx * xc * invzc + cx, left to right
            v = (((cam['fy'] * Pc[:, 1]).astype(f32) * invz).astype(f32) + cam['cy']).astype(f32)
            b = [f32(x) for x in bounds]
            ok = np.asarray(mp['has'], bool) & ~np.asarray(mp['bad'], bool) & ~np.asarray(mp['found'], bool)
            ok &= ~(u < b[0]) & ~(u > b[2]) & ~(v < b[1]) & ~(v > b[3])
            dist = _norm((P - Ow[None, :]).astype(f32))
            ok &= ~(dist < (f32(0.8) * mp['pts']['min_distance']).astype(f32)) & ~(dist > (f32(1.2) * mp['pts']['max_distance']).astype(f32))
            ratio = (mp['pts']['max_distance'] / dist).astype(f32)
        level = _levels(hvo_b200, ratio, cam['log_scale_factor'], 8)
        sel = np.nonzero(ok)[0]
        Cur['mappoint'] = np.where(np.asarray(Cur['claimed'], bool), -2, -1).astype(np.int32)
        Cur['claimed'] = np.asarray(Cur['claimed'], bool).copy()
        m = _matcher(hvo, gpu, 0.9, ori)
        nm, idx = m.SearchByProjectionKF(Cur, dict(u=u[sel], v=v[sel], level=level[sel], angle=k1['angle'][sel], desc=mp['desc'][sel]), f32(th), orb_dist)
        got = Cur['mappoint'].copy()
        new = got >= 0
        got[new] = sel[got[new]]
        assert nm == int(nm_r) and np.array_equal(got, ids)
        total += nm
    assert oracle.MATCH_EXE[0] or total > 150


# ---------------------------------------------------------------------------------------------------------------------------
# op 16: SearchByBoW(pKF1, pKF2, vpMatches12)
# ---------------------------------------------------------------------------------------------------------------------------
def _check_bow_kf(hvo, synth, gpu):
    from test_projection import _bow_scenario
    total = 0
    for seed, ratio, ori in ((0, 0.75, True), (1, 0.9, True), (2, 0.8, False)):
        KFa, Fb = _bow_scenario(synth, seed)
        rng = np.random.RandomState(600 + seed)
        n1, n2 = len(KFa['desc']), len(Fb['desc'])
        KF1 = dict(keys_un=KFa['keys_un'], desc=KFa['desc'], featvec=KFa['featvec'], has_mappoint=rng.rand(n1) < 0.8, bad=rng.rand(n1) < 0.05)
        KF2 = dict(keys_un=Fb['keys'], desc=Fb['desc'], featvec=Fb['featvec'], has_mappoint=rng.rand(n2) < 0.8, bad=rng.rand(n2) < 0.05)
        nm_r, m12_r = _ref(f'bowkf{seed}', lambda: oracle.ref_search_by_bow_kf(KF1, KF2, ratio, ori))
        if oracle.MATCH_EXE[0]:
            continue
        g1 = dict(KF1); g1['has_mappoint'] = KF1['has_mappoint'] & ~KF1['bad']
        g2 = dict(KF2); g2['has_mappoint'] = KF2['has_mappoint'] & ~KF2['bad']
        nm, m12 = _matcher(hvo, gpu, ratio, ori).SearchByBoWKF(g1, g2)
        assert nm == int(nm_r) and np.array_equal(m12, m12_r)
        total += nm
    assert oracle.MATCH_EXE[0] or total > 100


# ---------------------------------------------------------------------------------------------------------------------------
# ops 17, 18: the brute-force line matchers
# ---------------------------------------------------------------------------------------------------------------------------
class OracleBF:
    def knnMatch2(self, q, t):
        idx, dist = oracle.knn2(q, t)
        return idx, dist


def _lsd(hvo, gpu, ratio):
    if gpu:
        return hvo.LSDmatcher(ratio, True)
    m = hvo.LSDmatcher.__new__(hvo.LSDmatcher)
    m.mfNNratio, m.mbCheckOrientation, m._bf = f32(ratio), True, OracleBF()
    return m


def _check_line_bf(hvo, synth, gpu):
    total = 0
    for seed, TH, ratio, nnr in ((0, 50.0, 0.95, 0.8), (1, 80.0, 0.8, 0.6)):
        kl1, ld1, kl2, ld2, lv2, Fm = tt._line_scene(synth, seed)
        rng = np.random.RandomState(700 + seed)
        lm_r, n_nnr_r, m12_r, n_dbl_r, dbl_r = _ref(f'lbf{seed}', lambda: oracle.ref_line_bf(ld1, ld2, TH, ratio, nnr))
        has = rng.rand(len(ld1)) < 0.7
        n1_r, by_r, nd_r, dk_r = _ref(f'lbd{seed}', lambda: oracle.ref_line_by_descriptor(ld1, has, ld2, ratio))
        if oracle.MATCH_EXE[0]:
            continue
        m = _lsd(hvo, gpu, ratio)
        assert np.array_equal(m.FrameBFMatch(ld1, ld2, TH), lm_r)
        n_nnr, m12 = m.match(ld1, ld2, nnr)
        assert n_nnr == int(n_nnr_r) and np.array_equal(m12, m12_r)
        n_dbl, dbl = m.SearchDouble(ld1, ld2)
        assert n_dbl == int(n_dbl_r) and np.array_equal(dbl, dbl_r)
        n1, by = m.SearchByDescriptor(ld1, ld2, has)
        assert n1 == int(n1_r) and np.array_equal(by, by_r)
        nd, dk = m.SearchDoubleKF(ld1, has, ld2)
        assert nd == int(nd_r) and np.array_equal(dk, dk_r)
        has2 = rng.rand(len(ld2)) < 0.3
        n0_r, pairs_r, ns_r, ms_r, ndb_r, mdb_r = _ref(f'ltri{seed}', lambda: oracle.ref_line_triangulation(ld1, has, ld2, has2, ratio))
        n0, m0 = m.SearchForTriangulation(ld1, has, ld2, has2)
        i0 = np.nonzero(m0 >= 0)[0]
        assert n0 == int(n0_r) and np.array_equal(np.stack([i0, m0[i0]], 1).reshape(-1, 2), np.asarray(pairs_r).reshape(-1, 2))
        for dbl, nr, mr in ((False, ns_r, ms_r), (True, ndb_r, mdb_r)):
            n, mm = m.SearchForTriangulation(ld1, has, ld2, has2, th=m.TH_HIGH, is_double=dbl)
            assert n == int(nr) and np.array_equal(mm, mr)
        total += n_dbl + n1 + nd
    assert oracle.MATCH_EXE[0] or total > 60


CHECKS = (_check_projection_scw, _check_fuse_scw, _check_search_by_sim3, _check_reloc, _check_bow_kf, _check_line_bf)


@pytest.mark.parametrize('check', CHECKS, ids=lambda c: c.__name__[7:])
def test_oracle_equals_reference(check, hvo, synth):
    check(hvo, synth, gpu=False)


@pytest.mark.gpu
@pytest.mark.parametrize('check', CHECKS, ids=lambda c: c.__name__[7:])
def test_gpu_equals_reference(check, hvo, synth):
    check(hvo, synth, gpu=True)


def test_integer_window_origin_cannot_change_a_candidate_list():
    """The claim of include/hvo_capi.h (hvo_proj_set_window_origin): cells are assigned by round(), window cell ranges by floor() / ceil(), so
    locating the window from the key frame's truncated origin (a shift below one pixel, a tenth of a cell) only adds or drops cells without a
    keypoint within r: KeyFrame::GetFeaturesInArea and Frame::GetFeaturesInArea return the same list in the same order."""
    rng = np.random.RandomState(9)
    total = 0
    for trial in range(40):
        n = 1500
        bounds = (float(rng.uniform(-30, 30)), float(rng.uniform(-30, 30)), float(640 + rng.uniform(-30, 30)), float(480 + rng.uniform(-30, 30)))
        keys = np.zeros(n, oracle.KP_DTYPE)
        keys['x'] = rng.uniform(bounds[0], bounds[2], n).astype(f32); keys['y'] = rng.uniform(bounds[1], bounds[3], n).astype(f32)
        keys['octave'] = rng.randint(0, 8, n)
        # every fifth keypoint exactly on a cell border of the float grid, where round() flips
        cw = (bounds[2] - bounds[0]) / 64.0
        keys['x'][::5] = (bounds[0] + (rng.randint(0, 64, len(keys[::5])) + 0.5) * cw).astype(f32)
        for _ in range(60):
            x, y = f32(rng.uniform(bounds[0] - 20, bounds[2] + 20)), f32(rng.uniform(bounds[1] - 20, bounds[3] + 20))
            r = f32(rng.choice([1.5, 4.0, 7.5, 10.0, 25.0, 60.0]))
            lv = (-1, -1) if rng.rand() < 0.5 else (int(rng.randint(0, 4)), int(rng.randint(3, 8)))
            a = oracle.features_in_area(keys, bounds, x, y, r, *lv)
            oracle.window_origin((int(bounds[0]), int(bounds[1])))
            try:
                b = oracle.features_in_area(keys, bounds, x, y, r, *lv)
            finally:
                oracle.window_origin(None)
            assert np.array_equal(a, b)
            total += len(a)
    assert total > 5000
