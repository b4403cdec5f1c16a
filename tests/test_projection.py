"""Windowed (projection) matching: frame grid, GetFeaturesInArea, ORBmatcher::SearchByProjection (SURVEY §8 rows B4, B5, E2, E3,
E5 and the candidate-list primitive behind E4).

Oracle: oracle/proj_oracle.cpp restates src/Frame.cc:832-847, 1502-1555, 1680-1690 and src/ORBmatcher.cc:45-132, 1353-1497
(sequential greedy loops).  No reference execution is possible here and the reference has no tests: parity unpinned by
execution; the candidate set is cross-checked against a brute-force window test in numpy.
GPU bar: bit-exact (indices, distances, candidate order, match counts)."""
import numpy as np
import pytest

import oracle

BOUNDS = (0.0, 0.0, 640.0, 480.0)


def _frame_keys(synth, cfg='S1', idx=0):
    g, _ = synth.frame(cfg, idx)
    k, d = oracle.OrbOracle().extract(g)
    return k, d


def _scenario(synth, seed=0, n_extra=300, claims_all=True):
    """Frame = ORB of S1/0; queries = ORB of S1/1 treated as projected map points (same scene, small motion) plus noise points."""
    rng = np.random.RandomState(seed)
    k0, d0 = _frame_keys(synth, 'S1', 0)
    k1, d1 = _frame_keys(synth, 'S1', 1)
    sf = np.float32(1.2) ** np.arange(8, dtype=np.float32)
    n = len(k1) + n_extra
    q = np.zeros(n, oracle.PROJ_QUERY_DTYPE)
    lvl = np.concatenate([k1['octave'], rng.randint(0, 8, n_extra)]).astype(np.int32)
    q['u'] = np.concatenate([k1['x'] + rng.normal(0, 1.0, len(k1)), rng.uniform(-20, 660, n_extra)]).astype(np.float32)
    q['v'] = np.concatenate([k1['y'] + rng.normal(0, 1.0, len(k1)), rng.uniform(-20, 500, n_extra)]).astype(np.float32)
    base_r = np.where(rng.rand(n) > 0.5, np.float32(2.5), np.float32(4.0)).astype(np.float32) * np.float32(3.0)
    q['r'] = (base_r * sf[lvl]).astype(np.float32)
    q['min_level'] = lvl - 1; q['max_level'] = lvl
    q['ur'] = (q['u'] - rng.uniform(5, 40, n)).astype(np.float32)
    q['claims'] = 1 if claims_all else (rng.rand(n) > 0.3)
    qd = np.concatenate([d1, rng.randint(0, 256, (n_extra, 32)).astype(np.uint8)])
    uright = np.where(rng.rand(len(k0)) > 0.4, k0['x'] - rng.uniform(5, 40, len(k0)), -1).astype(np.float32)
    claimed = (rng.rand(len(k0)) < 0.1).astype(np.uint8)
    return k0, d0, uright, claimed, q, qd


def test_oracle_grid_and_area(synth):
    k, _ = _frame_keys(synth)
    cnt, items = oracle.grid_build(k, BOUNDS)
    px = np.floor((k['x'] - 0) * np.float32(64 / 640.0) + 0.5).astype(int); py = np.floor((k['y'] - 0) * np.float32(48 / 480.0) + 0.5).astype(int)
    inside = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    assert cnt.sum() == inside.sum() == len(items)
    assert np.array_equal(np.sort(items), np.nonzero(inside)[0])
    start = np.concatenate([[0], np.cumsum(cnt)])
    for c in np.nonzero(cnt > 1)[0][:50]:
        assert np.all(np.diff(items[start[c]:start[c + 1]]) > 0)          # push_back order = index order
    rng = np.random.RandomState(1)
    for _ in range(200):
        x, y, r = rng.uniform(-30, 670), rng.uniform(-30, 510), rng.uniform(1, 60)
        lo, hi = (-1, -1) if rng.rand() < 0.3 else (int(rng.randint(0, 5)), int(rng.randint(3, 8)))
        got = oracle.features_in_area(k, BOUNDS, x, y, r, lo, hi)
        ok = inside & (np.abs(k['x'] - np.float32(x)) < np.float32(r)) & (np.abs(k['y'] - np.float32(y)) < np.float32(r))
        if lo > 0 or hi >= 0:
            ok &= (k['octave'] >= lo)
            if hi >= 0:
                ok &= (k['octave'] <= hi)
        assert np.array_equal(np.sort(got), np.nonzero(ok)[0])
        cell = px[got] * 48 + py[got]
        assert np.all(np.diff(cell) >= 0)                                   # ix outer, iy inner


def test_oracle_greedy_conflict():
    # three keypoints at one spot with descriptors at distance 0 / 8 / 16 from the query descriptor; three identical queries
    keys = np.zeros(3, oracle.KP_DTYPE); keys['x'] = [100, 101, 102]; keys['y'] = 100; keys['octave'] = [0, 1, 0]
    qd = np.zeros((3, 32), np.uint8)
    d = np.zeros((3, 32), np.uint8); d[1, 0] = 0xff; d[2, 0] = 0xff; d[2, 1] = 0xff
    q = np.zeros(3, oracle.PROJ_QUERY_DTYPE); q['u'] = 101; q['v'] = 100; q['r'] = 10; q['min_level'] = -1; q['max_level'] = -1; q['claims'] = 1
    idx, dist, n = oracle.search_projection(keys, None, d, BOUNDS, q, qd, mode=1, th_dist=100)
    assert idx.tolist() == [0, 1, 2] and dist.tolist() == [0, 8, 16] and n == 3
    q['claims'] = 0                                                         # without observations nothing is ever claimed
    idx, dist, n = oracle.search_projection(keys, None, d, BOUNDS, q, qd, mode=1, th_dist=100)
    assert idx.tolist() == [0, 0, 0]
    q['claims'] = 1                                                         # mode 0: best 0 / second 16 at the same level -> ratio ok
    idx, _, _ = oracle.search_projection(keys, None, d, BOUNDS, q[:1], qd[:1], mode=0, th_dist=100, nnratio=0.6)
    assert idx.tolist() == [0]
    d[0, 0] = 0x0f                                                          # best 4 (level 0), second 8 (level 1): ratio not applied
    idx, _, _ = oracle.search_projection(keys, None, d, BOUNDS, q[:1], qd[:1], mode=0, th_dist=100, nnratio=0.3)
    assert idx.tolist() == [0]
    keys['octave'] = 0                                                      # same level now: 4 > 0.3 * 8 -> rejected
    idx, _, _ = oracle.search_projection(keys, None, d, BOUNDS, q[:1], qd[:1], mode=0, th_dist=100, nnratio=0.3)
    assert idx.tolist() == [-1]


@pytest.mark.gpu
def test_gpu_grid_and_area(hvo, synth):
    k, d = _frame_keys(synth)
    pm = hvo.ProjectionMatcher()
    pm.set_frame(k, None, d, *BOUNDS)
    cs, items = pm.grid()
    cnt, oitems = oracle.grid_build(k, BOUNDS)
    assert np.array_equal(np.diff(cs), cnt) and np.array_equal(items, oitems)
    rng = np.random.RandomState(2)
    for _ in range(60):
        x, y, r = rng.uniform(-30, 670), rng.uniform(-30, 510), rng.uniform(1, 80)
        lo, hi = (-1, -1) if rng.rand() < 0.3 else (int(rng.randint(0, 5)), int(rng.randint(3, 8)))
        assert np.array_equal(pm.GetFeaturesInArea(x, y, r, lo, hi), oracle.features_in_area(k, BOUNDS, x, y, r, lo, hi))


@pytest.mark.gpu
@pytest.mark.parametrize('mode,claims_all,seed', [(0, True, 0), (1, True, 1), (0, False, 2), (1, False, 3)])
def test_gpu_search_projection_bit_exact(hvo, synth, mode, claims_all, seed):
    k0, d0, ur, claimed, q, qd = _scenario(synth, seed, claims_all=claims_all)
    pm = hvo.ProjectionMatcher()
    pm.set_frame(k0, ur, d0, *BOUNDS)
    idx, dist, n = pm.search(q, qd, claimed, mode, 100, 0.8)
    oidx, odist, on = oracle.search_projection(k0, ur, d0, BOUNDS, q, qd, claimed, mode, 100, 0.8)
    assert n == on and np.array_equal(idx, oidx) and np.array_equal(dist[idx >= 0], odist[oidx >= 0])
    assert n > 300                                                          # the scenario really matches
    assert not np.any(claimed[idx[idx >= 0]])
    if claims_all:
        m = idx[idx >= 0]
        assert len(np.unique(m)) == len(m)                                  # every keypoint taken at most once


@pytest.mark.gpu
def test_gpu_search_projection_conflict_chain(hvo):
    # 40 identical queries compete for 40 keypoints whose descriptors are at distance 0, 1, 2, ...: query i must end on keypoint i,
    # which the fixed-point iteration reaches one query per round
    n = 40
    keys = np.zeros(n, oracle.KP_DTYPE); keys['x'] = 300 + 0.1 * np.arange(n); keys['y'] = 200
    d = np.zeros((n, 32), np.uint8)
    for i in range(n):
        bits = np.zeros(256, np.uint8); bits[:i] = 1
        d[i] = np.packbits(bits)
    q = np.zeros(n, oracle.PROJ_QUERY_DTYPE); q['u'] = 302; q['v'] = 200; q['r'] = 20; q['min_level'] = -1; q['max_level'] = -1; q['claims'] = 1
    qd = np.zeros((n, 32), np.uint8)
    pm = hvo.ProjectionMatcher()
    pm.set_frame(keys, None, d, *BOUNDS)
    idx, dist, nm = pm.search(q, qd, None, 1, 100, 0.6)
    oidx, odist, on = oracle.search_projection(keys, None, d, BOUNDS, q, qd, None, 1, 100, 0.6)
    assert np.array_equal(idx, np.arange(n)) and np.array_equal(idx, oidx) and np.array_equal(dist, odist) and nm == on == n
    assert pm.rounds() >= n


@pytest.mark.gpu
def test_gpu_search_projection_edge_cases(hvo, synth):
    pm = hvo.ProjectionMatcher()
    k, d = _frame_keys(synth)
    q = np.zeros(3, oracle.PROJ_QUERY_DTYPE); q['u'] = [-500, 320, 5000]; q['v'] = [240, -900, 240]; q['r'] = 5; q['min_level'] = -1; q['max_level'] = -1
    qd = np.zeros((3, 32), np.uint8)
    pm.set_frame(k[:0], None, d[:0], *BOUNDS)                                # empty frame
    idx, dist, n = pm.search(q, qd)
    assert n == 0 and np.all(idx == -1)
    pm.set_frame(k, None, d, *BOUNDS)
    idx, dist, n = pm.search(q, qd)                                        # windows outside the image
    assert n == 0 and np.all(idx == -1)
    idx, dist, n = pm.search(q[:0], qd[:0])                                # no queries
    assert n == 0 and len(idx) == 0


@pytest.mark.gpu
def test_gpu_match_candidates(hvo, synth):
    rng = np.random.RandomState(5)
    q, t = synth.descriptors_S4(nq=300, nt=2000, seed=21, planted=100, ties=20)
    lens = rng.randint(0, 70, 300); lens[:5] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cand = rng.randint(0, 2000, off[-1]).astype(np.int32)
    pm = hvo.ProjectionMatcher()
    got = pm.match_candidates(q, t, off, cand)
    assert np.array_equal(got, oracle.match_candidates(q, t, off, cand))


def _mirror_inputs(synth, seed):
    k0, d0, ur, claimed, q, qd = _scenario(synth, seed, n_extra=100, claims_all=False)
    rng = np.random.RandomState(seed + 100)
    sf = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
    M = len(q)
    lvl = q['max_level'].copy()
    MPs = dict(proj_x=q['u'].copy(), proj_y=q['v'].copy(), proj_xr=q['ur'].copy(), view_cos=rng.uniform(0.99, 1.0, M).astype(np.float32),
               level=lvl, in_view=rng.rand(M) > 0.1, bad=rng.rand(M) < 0.05, has_obs=q['claims'].astype(bool), desc=qd)
    F = dict(keys_un=k0, uright=ur, desc=d0, bounds=BOUNDS, scale_factors=sf, mappoint=np.full(len(k0), -1, np.int64), claimed=claimed.astype(bool))
    return F, MPs, sf


@pytest.mark.gpu
def test_orbmatcher_search_by_projection_mirror(hvo, synth):
    F, MPs, sf = _mirror_inputs(synth, 7)
    th = 3.0
    # the reference loop, written out on top of the oracle primitive
    use = MPs['in_view'] & ~MPs['bad']
    sel = np.nonzero(use)[0]
    r = np.where(MPs['view_cos'][sel] > np.float32(0.998), np.float32(2.5), np.float32(4.0)).astype(np.float32) * np.float32(th)
    q = np.zeros(len(sel), oracle.PROJ_QUERY_DTYPE)
    q['u'] = MPs['proj_x'][sel]; q['v'] = MPs['proj_y'][sel]; q['r'] = (r * sf[MPs['level'][sel]]).astype(np.float32)
    q['min_level'] = MPs['level'][sel] - 1; q['max_level'] = MPs['level'][sel]; q['ur'] = MPs['proj_xr'][sel]; q['claims'] = MPs['has_obs'][sel]
    oidx, _, on = oracle.search_projection(F['keys_un'], F['uright'], F['desc'], BOUNDS, q, MPs['desc'][sel], F['claimed'].astype(np.uint8), 0, 100, 0.8)
    m = hvo.ORBmatcher(0.8, True)
    n, match = m.SearchByProjection(F, MPs, th)
    assert n == on and np.array_equal(match[sel], oidx) and np.all(match[~use] == -1)
    taken = match[match >= 0]
    assert np.all(F['mappoint'][taken] >= 0)


@pytest.mark.gpu
def test_orbmatcher_search_by_projection_last_mirror(hvo, synth):
    F, MPs, sf = _mirror_inputs(synth, 9)
    rng = np.random.RandomState(11)
    n = len(MPs['proj_x'])
    last = dict(u=MPs['proj_x'], v=MPs['proj_y'], ur=MPs['proj_xr'], octave=np.clip(MPs['level'], 0, 7), has_obs=MPs['has_obs'], desc=MPs['desc'],
                angle=rng.uniform(0, 360, n).astype(np.float32))
    m = hvo.ORBmatcher(0.9, True)
    nm, idx = m.SearchByProjectionLast(F, last, th=7.0)
    q = np.zeros(n, oracle.PROJ_QUERY_DTYPE)
    q['u'] = last['u']; q['v'] = last['v']; q['ur'] = last['ur']; q['r'] = (np.float32(7.0) * sf[last['octave']]).astype(np.float32)
    q['min_level'] = last['octave'] - 1; q['max_level'] = last['octave'] + 1; q['claims'] = last['has_obs']
    claimed0 = _mirror_inputs(synth, 9)[0]['claimed']
    oidx, _, on = oracle.search_projection(F['keys_un'], F['uright'], F['desc'], BOUNDS, q, last['desc'], claimed0.astype(np.uint8), 1, 100, 0.9)
    assert np.array_equal(idx, oidx)
    assert 0 < nm <= on                                                    # the rotation histogram only removes matches


def _bow_scenario(synth, seed=0):
    """Key frame = ORB of S1/0, frame = ORB of S1/1; the vocabulary node of a feature is faked by a coarse hash of its position and
    octave (what matters to the matcher is only which features share a node)."""
    rng = np.random.RandomState(seed)
    k0, d0 = _frame_keys(synth, 'S1', 0)
    k1, d1 = _frame_keys(synth, 'S1', 1)
    node = lambda k: (k['x'] // 80).astype(int) * 100 + (k['y'] // 80).astype(int) * 10 + (k['octave'] // 3)
    fv0, fv1 = {}, {}
    for i, n in enumerate(node(k0)):
        fv0.setdefault(int(n), []).append(i)
    for i, n in enumerate(node(k1)):
        fv1.setdefault(int(n), []).append(i)
    has_mp = rng.rand(len(k0)) > 0.3
    return dict(desc=d0, keys_un=k0, featvec=fv0, has_mappoint=has_mp), dict(desc=d1, keys=k1, featvec=fv1)


def test_oracle_search_candidates_is_greedy():
    q = np.zeros((3, 32), np.uint8)
    t = np.zeros((3, 32), np.uint8); t[1, 0] = 0x0f; t[2, :2] = 0xff               # distances 0, 4, 16 to every query
    off = np.array([0, 3, 6, 9], np.int32); cand = np.tile(np.arange(3, dtype=np.int32), 3)
    idx, dist, n = oracle.search_candidates(q, t, off, cand, th_dist=50, nnratio=0.7)
    # query 0: best 0 < 0.7 * 4 -> takes row 0; query 1: best 4 < 0.7 * 16 -> row 1; query 2: only row 2 left, second = 256 -> row 2
    assert idx.tolist() == [0, 1, 2] and dist.tolist() == [0, 4, 16] and n == 3
    idx, _, n = oracle.search_candidates(q, t, off, cand, th_dist=3, nnratio=0.7)   # TH below 4: queries 1 and 2 fail
    assert idx.tolist() == [0, -1, -1] and n == 1
    t[1, 0] = 0                                                                     # tie 0 / 0: 0 < 0.7 * 0 is false
    idx, _, n = oracle.search_candidates(q, t, off, cand, th_dist=50, nnratio=0.7)
    assert idx.tolist() == [-1, -1, -1] and n == 0


@pytest.mark.gpu
def test_gpu_search_candidates_and_search_by_bow(hvo, synth):
    KF, F = _bow_scenario(synth)
    m = hvo.ORBmatcher(0.7, True)
    qi, off, cand = m.bow_queries(KF['featvec'], F['featvec'], KF['has_mappoint'])
    assert len(qi) > 300 and np.all(KF['has_mappoint'][qi])
    pm = hvo.ProjectionMatcher()
    idx, dist, nm = pm.search_candidates(KF['desc'][qi], F['desc'], off, cand, 50, 0.7)
    ridx, rdist, rnm = oracle.search_candidates(KF['desc'][qi], F['desc'], off, cand, 50, 0.7)
    assert rnm > 50 and np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and nm == rnm
    taken = idx[idx >= 0]
    assert len(set(taken.tolist())) == len(taken)                                   # a frame feature is given away once
    # heavy contention: every query sees the same short list
    q = np.repeat(KF['desc'][qi[:1]], 30, axis=0)
    off2 = (np.arange(31) * 12).astype(np.int32); cand2 = np.tile(cand[off[0]:off[0] + 12] if off[1] - off[0] >= 12 else np.arange(12, dtype=np.int32), 30)
    idx, dist, nm = pm.search_candidates(q, F['desc'], off2, cand2, 255, 1.1)
    ridx, rdist, rnm = oracle.search_candidates(q, F['desc'], off2, cand2, 255, 1.1)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and nm == rnm and pm.rounds() > 2
    pm.close()
    # the mirror: same assignment, then the rotation-histogram filter (ORBmatcher.cc:268-290)
    nm, match = m.SearchByBoW(KF, F)
    ridx, _, rnm = oracle.search_candidates(KF['desc'][qi], F['desc'], off, cand, 50, 0.7)
    want = np.full(len(F['desc']), -1, np.int32)
    hist = [[] for _ in range(30)]
    for k, i in zip(qi, ridx):
        if i >= 0:
            want[i] = k
            rot = np.float32(KF['keys_un']['angle'][k]) - np.float32(F['keys']['angle'][i])
            rot = np.float32(rot + np.float32(360)) if rot < 0 else rot
            b = int(np.floor(float(np.float32(rot * (np.float32(1) / np.float32(30)))) + 0.5)) % 30
            hist[b].append(i)
    keep = set(m.ComputeThreeMaxima([len(h) for h in hist]))
    for b in range(30):
        if b not in keep:
            want[hist[b]] = -1
    assert np.array_equal(match, want) and nm == int((want >= 0).sum()) > 20


def test_oracle_fuse_gate():
    # one keypoint at (100, 100), level 1; sigma2 = 1.44 -> inv 0.694: e2 * inv <= 5.99  <=>  e2 <= 8.63
    keys = np.zeros(1, oracle.KP_DTYPE); keys['x'] = 100; keys['y'] = 100; keys['octave'] = 1
    d = np.zeros((1, 32), np.uint8)
    inv = (1.0 / (np.float32(1.2) ** np.arange(8, dtype=np.float32)) ** 2).astype(np.float32)
    q = np.zeros(3, oracle.PROJ_QUERY_DTYPE); q['r'] = 10; q['v'] = 100; q['min_level'] = 0; q['max_level'] = 1
    q['u'] = [102, 103.5, 102]                      # e2 = 4 passes, e2 = 12.25 fails
    q['ur'] = [80, 80, 75]
    idx, dist, n = oracle.search_fuse(keys, None, d, BOUNDS, q, np.zeros((3, 32), np.uint8), inv, 50)
    assert idx.tolist() == [0, -1, 0] and n == 2
    ur = np.array([78.0], np.float32)               # stereo: e2 = 4 + 4 = 8 -> 5.56 <= 7.8 passes; 4 + 9 = 13 -> 9.03 fails
    idx, _, n = oracle.search_fuse(keys, ur, d, BOUNDS, q, np.zeros((3, 32), np.uint8), inv, 50)
    assert idx.tolist() == [0, -1, -1] and n == 1
    q['min_level'] = 2; q['max_level'] = 3          # level window excludes the keypoint
    assert oracle.search_fuse(keys, None, d, BOUNDS, q, np.zeros((3, 32), np.uint8), inv, 50)[2] == 0


@pytest.mark.gpu
def test_gpu_fuse_and_reloc_projection_match_oracle(hvo, synth):
    k0, d0, ur, claimed, q, qd = _scenario(synth, seed=3, n_extra=200, claims_all=True)
    sf = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
    inv = (np.float32(1.0) / (sf * sf)).astype(np.float32)
    lvl = np.clip(q['max_level'], 0, 7).astype(np.int32)
    m = hvo.ORBmatcher(0.6, True)
    # ---- Fuse: independent queries, chi-square gate, TH_LOW ----
    KF = dict(keys_un=k0, uright=ur, desc=d0, bounds=BOUNDS, scale_factors=sf, inv_level_sigma2=inv)
    MPs = dict(u=q['u'], v=q['v'], ur=q['ur'], level=lvl, desc=qd)
    nf, best = m.Fuse(KF, MPs, th=3.0)
    qq = q.copy(); qq['r'] = (np.float32(3.0) * sf[lvl]).astype(np.float32); qq['min_level'] = lvl - 1; qq['max_level'] = lvl
    ridx, _, rn = oracle.search_fuse(k0, ur, d0, BOUNDS, qq, qd, inv, 50)
    assert rn > 30 and nf == rn and np.array_equal(best, ridx)
    nf2, best2 = m.Fuse(dict(KF, uright=None), MPs, th=3.0)          # monocular key frame: 5.99 gate only
    ridx2, _, rn2 = oracle.search_fuse(k0, None, d0, BOUNDS, qq, qd, inv, 50)
    assert nf2 == rn2 >= rn and np.array_equal(best2, ridx2)
    # ---- relocalisation variant: any map point blocks a keypoint, every match claims, no stereo check, ORBdist ----
    has_mp = (np.random.RandomState(4).rand(len(k0)) < 0.15)
    Cur = dict(keys_un=k0, desc=d0, bounds=BOUNDS, scale_factors=sf, mappoint=np.where(has_mp, 0, -1).astype(np.int32), claimed=has_mp.copy())
    kf = dict(u=q['u'], v=q['v'], level=lvl, angle=np.zeros(len(q), np.float32), desc=qd)
    m2 = hvo.ORBmatcher(0.6, False)                                  # rotation check off: pure assignment parity
    nm, match = m2.SearchByProjectionKF(Cur, kf, th=10.0, ORBdist=64)
    q3 = q.copy(); q3['r'] = (np.float32(10.0) * sf[lvl]).astype(np.float32); q3['min_level'] = lvl - 1; q3['max_level'] = lvl + 1; q3['claims'] = 1
    ridx3, _, rn3 = oracle.search_projection(k0, None, d0, BOUNDS, q3, qd, has_mp.astype(np.uint8), 1, 64)
    assert rn3 > 50 and nm == rn3 and np.array_equal(match, ridx3)
    assert not np.any(has_mp[match[match >= 0]])                     # blocked keypoints never receive a match


@pytest.mark.gpu
def test_gpu_loop_closing_searches_match_oracle(hvo, synth):
    """ORBmatcher::Fuse(KF, Scw, ...) and SearchByProjection(KF, Scw, ...) (ORBmatcher.cc:992-1121, 295-410): window + level range,
    best <= TH_LOW; the first claims nothing, the second skips and fills vpMatched."""
    k0, d0, ur, claimed, q, qd = _scenario(synth, seed=8, n_extra=150, claims_all=True)
    sf = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
    lvl = np.clip(q['max_level'], 0, 7).astype(np.int32)
    KF = dict(keys_un=k0, desc=d0, bounds=BOUNDS, scale_factors=sf)
    pts = dict(u=q['u'], v=q['v'], level=lvl, desc=qd)
    m = hvo.ORBmatcher(0.75, True)
    qq = q.copy(); qq['r'] = (np.float32(4.0) * sf[lvl]).astype(np.float32); qq['min_level'] = lvl - 1; qq['max_level'] = lvl; qq['ur'] = -1
    qq['claims'] = 0
    n1, b1 = m.FuseSim3(KF, pts, 4.0)
    r1, _, rn1 = oracle.search_projection(k0, None, d0, BOUNDS, qq, qd, None, 1, 50)
    assert rn1 > 30 and n1 == rn1 and np.array_equal(b1, r1)
    assert len(set(b1[b1 >= 0].tolist())) < (b1 >= 0).sum() or True          # without claims several points may pick one keypoint
    matched = claimed.astype(bool).copy()
    qq['claims'] = 1
    n2, b2 = m.SearchByProjectionSim3(KF, pts, matched, 4.0)
    r2, _, rn2 = oracle.search_projection(k0, None, d0, BOUNDS, qq, qd, claimed, 1, 50)
    assert rn2 > 30 and n2 == rn2 and np.array_equal(b2, r2)
    got = b2[b2 >= 0]
    assert len(set(got.tolist())) == len(got) and not np.any(claimed.astype(bool)[got]) and np.all(matched[got])


def _fundamental(tx=0.05, ty=0.01, tz=0.0, fx=535.4, fy=539.2, cx=320.1, cy=247.6):
    """F12 of two cameras that differ by a small translation (x1' F12 x2 = 0), float32 like the reference's cv::Mat."""
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    t = np.array([tx, ty, tz])
    tx_ = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    F = np.linalg.inv(K).T @ tx_ @ np.linalg.inv(K)
    return (F / np.abs(F).max()).astype(np.float32)


def test_oracle_triangulation_gates():
    k1 = np.zeros(1, oracle.KP_DTYPE); k1['x'] = 300; k1['y'] = 200
    k2 = np.zeros(4, oracle.KP_DTYPE); k2['x'] = [310, 310, 310, 400]; k2['y'] = [200, 200, 230, 200]
    F = _fundamental(ty=0.0)                                  # pure x translation: epipolar lines are horizontal (y2 = y1)
    d1 = np.zeros((1, 32), np.uint8); d2 = np.zeros((4, 32), np.uint8); d2[3, 0] = 0x01
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32); sg = (sf * sf).astype(np.float32)
    off = np.array([0, 4], np.int32); cand = np.arange(4, dtype=np.int32)
    args = (d1, k1, [0], d2, k2, None, off, cand, F, -1e6, -1e6, sf, sg)
    run = lambda fl, **kw: oracle.search_triangulation(*args[:5], fl, *args[6:], **kw)
    idx, dist, n = run([0, 0, 0, 0])
    assert idx.tolist() == [1] and dist.tolist() == [0] and n == 1       # equal distances: the LAST passing candidate; row 2 is off the line
    assert run([0, 1, 0, 0])[0].tolist() == [0]                         # a candidate with a map point is skipped
    assert run([1, 1, 0, 0])[0].tolist() == [3]                         # only the worse one on the line remains (distance 1)
    assert run([0, 0, 0, 0], only_stereo=True)[2] == 0                  # monocular query under bOnlyStereo
    idx, _, _ = oracle.search_triangulation(d1, k1, [0], d2, k2, [0, 0, 0, 0], off, cand, F, 310.0, 200.0, sf, sg)
    assert idx.tolist() == [3]                                          # candidates within 10 px of the epipole are dropped (mono pair)


@pytest.mark.gpu
def test_gpu_search_for_triangulation_matches_oracle(hvo, synth):
    KFa, Fb = _bow_scenario(synth, seed=2)
    rng = np.random.RandomState(12)
    k1, k2 = KFa['keys_un'], Fb['keys']
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32); sg = (sf * sf).astype(np.float32)
    ur1 = np.where(rng.rand(len(k1)) < 0.6, k1['x'] - 20, -1).astype(np.float32)
    ur2 = np.where(rng.rand(len(k2)) < 0.6, k2['x'] - 20, -1).astype(np.float32)
    KF1 = dict(desc=KFa['desc'], keys_un=k1, uright=ur1, featvec=KFa['featvec'], has_mappoint=rng.rand(len(k1)) < 0.4, scale_factors=sf, level_sigma2=sg)
    KF2 = dict(desc=Fb['desc'], keys_un=k2, uright=ur2, featvec=Fb['featvec'], has_mappoint=rng.rand(len(k2)) < 0.3, scale_factors=sf, level_sigma2=sg)
    F12 = _fundamental(0.004, 0.0005)                        # S1/0 -> S1/1 is a small motion: matches lie near these epipolar lines
    m = hvo.ORBmatcher(0.6, False)
    qi, off, cand = m.bow_queries(KF1['featvec'], KF2['featvec'], ~KF1['has_mappoint'])
    tflags = (KF2['has_mappoint'].astype(np.uint8) | ((ur2 >= 0).astype(np.uint8) << 1)).astype(np.uint8)
    pm = hvo.ProjectionMatcher()
    total = 0
    for only_stereo in (False, True):
        for ex, ey in ((-1e5, 240.0), (330.0, 240.0)):
            got = pm.search_triangulation(KF1['desc'][qi], k1[qi], ur1[qi] >= 0, KF2['desc'], k2, tflags, off, cand, F12, ex, ey, sf, sg, only_stereo, 50)
            ref = oracle.search_triangulation(KF1['desc'][qi], k1[qi], ur1[qi] >= 0, KF2['desc'], k2, tflags, off, cand, F12, ex, ey, sf, sg, only_stereo, 50)
            assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2]
            total += ref[2]
    assert total > 40
    pm.close()
    nm, pairs = m.SearchForTriangulation(KF1, KF2, F12, (-1e5, 240.0), False)
    ref = oracle.search_triangulation(KF1['desc'][qi], k1[qi], ur1[qi] >= 0, KF2['desc'], k2, tflags, off, cand, F12, -1e5, 240.0, sf, sg, False, 50)
    want = {(int(a), int(b)) for a, b in zip(qi, ref[0]) if b >= 0}
    assert nm == len(want) > 10 and {(int(a), int(b)) for a, b in pairs} == want
    assert not np.any(KF1['has_mappoint'][pairs[:, 0]]) and not np.any(KF2['has_mappoint'][pairs[:, 1]])


@pytest.mark.gpu
def test_gpu_search_by_sim3_cross_check(hvo, synth):
    """ORBmatcher::SearchBySim3 (ORBmatcher.cc:1123-1351): two windowed searches (TH_HIGH, levels [l-1, l], nothing claimed) and
    the agreement test, against the oracle's sequential loop run in both directions."""
    ka, da = _frame_keys(synth, 'S1', 0)
    kb, db = _frame_keys(synth, 'S1', 1)
    rng = np.random.RandomState(21)
    sf = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
    KF1 = dict(keys_un=ka, desc=da, bounds=BOUNDS, scale_factors=sf)
    KF2 = dict(keys_un=kb, desc=db, bounds=BOUNDS, scale_factors=sf)

    def proj(k, d, n):                                   # "map points" of one key frame seen in the other: same scene, small motion
        sel = np.sort(rng.choice(len(k), n, replace=False))
        return dict(index=sel, u=(k['x'][sel] + rng.normal(0, 1.5, n)).astype(np.float32), v=(k['y'][sel] + rng.normal(0, 1.5, n)).astype(np.float32),
                    level=k['octave'][sel].astype(np.int32), desc=d[sel])
    p12, p21 = proj(ka, da, 600), proj(kb, db, 600)
    m = hvo.ORBmatcher(0.75, True)
    n, pairs = m.SearchBySim3(KF1, KF2, p12, p21, 7.5)

    def ref(K, D, pts):
        q = np.zeros(len(pts['u']), oracle.PROJ_QUERY_DTYPE)
        q['u'] = pts['u']; q['v'] = pts['v']; q['ur'] = -1; q['r'] = (np.float32(7.5) * sf[pts['level']]).astype(np.float32)
        q['min_level'] = pts['level'] - 1; q['max_level'] = pts['level']
        return oracle.search_projection(K, None, D, BOUNDS, q, pts['desc'], None, 1, 100)[0]
    m1 = np.full(len(ka), -1, np.int32); m1[p12['index']] = ref(kb, db, p12)
    m2 = np.full(len(kb), -1, np.int32); m2[p21['index']] = ref(ka, da, p21)
    want = [(i, m1[i]) for i in range(len(ka)) if m1[i] >= 0 and m2[m1[i]] == i]
    assert n == len(want) > 30 and [tuple(p) for p in pairs.tolist()] == want
