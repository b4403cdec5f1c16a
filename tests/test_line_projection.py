"""Windowed (projection) line matching: line grid, GetFeaturesInAreaForLine, LSDmatcher::SearchByProjection x2.

Oracle: oracle/lproj_oracle.cpp restates src/lineIterator.cpp, src/Frame.cc:849-872, 1557-1631 and src/LSDmatcher.cpp:561-664,
709-801 (sequential greedy loops).  No reference execution is possible here and the reference has no tests: parity unpinned
by execution; the candidate set is cross-checked against an independent numpy statement of the window test.
GPU bar: bit-exact (grid, candidate order, indices, distances, match counts)."""
import numpy as np
import pytest

import oracle

BOUNDS = (0.0, 0.0, 640.0, 480.0)


def _lines(synth, idx, cfg='S1'):
    g, _ = synth.frame(cfg, idx)
    kl, desc, lv = oracle.line_extract(g, n_features=200)
    return kl, desc, lv


def _scenario(synth, seed=0, n_extra=150, mode=0):
    """Frame = lines of S1/0; queries = lines of S1/1 (same scene, small motion) jittered, plus random segments."""
    rng = np.random.RandomState(seed)
    kl0, d0, lv0 = _lines(synth, 0)
    kl1, d1, _ = _lines(synth, 1)
    n1 = len(kl1)
    n = n1 + n_extra
    q = np.zeros(n, oracle.LPROJ_QUERY_DTYPE)
    jit = lambda a: (a + rng.normal(0, 1.0, len(a))).astype(np.float32)
    ex = rng.uniform(-20, 660, (n_extra, 2)); ey = rng.uniform(-20, 500, (n_extra, 2))
    q['x1'] = np.concatenate([jit(kl1['startPointX']), ex[:, 0]]); q['y1'] = np.concatenate([jit(kl1['startPointY']), ey[:, 0]])
    q['x2'] = np.concatenate([jit(kl1['endPointX']), ex[:, 1]]); q['y2'] = np.concatenate([jit(kl1['endPointY']), ey[:, 1]])
    qd = np.concatenate([d1, rng.randint(0, 256, (n_extra, 32)).astype(np.uint8)])
    # 3-D lines of the frame: a direction per line; map lines reuse the direction of the nearest frame line (or a random one)
    dirs0 = rng.normal(size=(len(kl0), 3))
    p0 = rng.uniform(-2, 2, (len(kl0), 3))
    lines3d = np.concatenate([p0 + dirs0, p0], axis=1)
    if mode == 0:
        q['r'] = np.where(rng.rand(n) > 0.5, np.float32(5.0), np.float32(8.0)) * np.float32(rng.choice([1.0, 3.0]))
        q['cos_th'] = np.float32(0.998)
        mid0 = np.stack([(kl0['startPointX'] + kl0['endPointX']) / 2, (kl0['startPointY'] + kl0['endPointY']) / 2], 1)
        midq = np.stack([(q['x1'] + q['x2']) / 2, (q['y1'] + q['y2']) / 2], 1)
        near = np.argmin(((midq[:, None, :] - mid0[None, :, :]) ** 2).sum(-1), axis=1)
        wv = dirs0[near] * rng.choice([-1.0, 1.0], (n, 1)) * rng.uniform(0.5, 2.0, (n, 1)) + rng.normal(0, 0.05, (n, 3))
        rnd = rng.rand(n) < 0.2
        wv[rnd] = rng.normal(size=(int(rnd.sum()), 3))
        q['dir'] = wv
    else:
        q['r'] = np.float32(15.0)
        q['cos_th'] = np.float32(0.96)
        dx = np.concatenate([kl1['ePointInOctaveX'] - kl1['sPointInOctaveX'], (ex[:, 1] - ex[:, 0]).astype(np.float32)])
        dy = np.concatenate([kl1['ePointInOctaveY'] - kl1['sPointInOctaveY'], (ey[:, 1] - ey[:, 0]).astype(np.float32)])
        q['dir'][:, 0] = dx.astype(np.float32); q['dir'][:, 1] = dy.astype(np.float32)
        q['length'] = np.concatenate([kl1['lineLength'], np.hypot(ex[:, 1] - ex[:, 0], ey[:, 1] - ey[:, 0]).astype(np.float32)])
    q['claims'] = rng.rand(n) > 0.25
    claimed = (rng.rand(len(kl0)) < 0.1).astype(np.uint8)
    return kl0, lv0, d0, lines3d, claimed, q, qd


def _bresenham_cells(kl, i):
    """independent statement of src/lineIterator.cpp on the grid coordinates of line i"""
    iw, ih = np.float32(64) / np.float32(640), np.float32(48) / np.float32(480)
    x1, y1 = float(kl['startPointX'][i] * iw), float(kl['startPointY'][i] * ih)
    x2, y2 = float(kl['endPointX'][i] * iw), float(kl['endPointY'][i] * ih)
    steep = abs(y2 - y1) > abs(x2 - x1)
    if steep:
        x1, y1, x2, y2 = y1, x1, y2, x2
    if x1 > x2:
        x1, x2, y1, y2 = x2, x1, y2, y1
    dx, dy = x2 - x1, abs(y2 - y1)
    err, ystep, x, y = dx / 2.0, (1 if y1 < y2 else -1), int(x1), int(y1)
    out = []
    while x <= int(x2):
        px, py = (y, x) if steep else (x, y)
        if 0 <= px < 64 and 0 <= py < 48:
            out.append(px * 48 + py)
        err -= dy
        if err < 0:
            y += ystep; err += dx
        x += 1
    return out


def test_oracle_line_grid_and_area(synth):
    kl, _, lv = _lines(synth, 0)
    assert len(kl) > 50
    cnt, items = oracle.line_grid_build(kl, BOUNDS)
    start = np.concatenate([[0], np.cumsum(cnt)])
    want = [[] for _ in range(64 * 48)]
    for i in range(len(kl)):
        for c in _bresenham_cells(kl, i):
            want[c].append(i)
    for c in range(64 * 48):
        assert items[start[c]:start[c + 1]].tolist() == want[c]
    # every line is in the cells of both endpoints' columns: at least ceil(len / cell) cells
    assert cnt.sum() >= len(kl)
    rng = np.random.RandomState(2)
    f32 = np.float32
    for t in range(300):
        j = rng.randint(len(kl))
        if t % 3 == 0:
            x1, y1, x2, y2 = rng.uniform(-30, 670), rng.uniform(-30, 510), rng.uniform(-30, 670), rng.uniform(-30, 510)
        else:                                                # near an existing line, so that the direction gate passes
            x1, y1 = kl['startPointX'][j] + rng.normal(0, 2), kl['startPointY'][j] + rng.normal(0, 2)
            x2, y2 = kl['endPointX'][j] + rng.normal(0, 2), kl['endPointY'][j] + rng.normal(0, 2)
        r, TH = f32(rng.choice([5, 8, 15, 24])), f32(rng.choice([0.998, 0.96]))
        got = oracle.line_features_in_area(kl, lv, BOUNDS, x1, y1, x2, y2, r, TH)
        assert len(set(got.tolist())) == len(got)
        # independent statement: a line is a candidate iff for some sample point it lies in a cell of the window and passes both tests
        x1, y1, x2, y2 = f32(x1), f32(y1), f32(x2), f32(y2)
        xs = [x1, f32((np.float64(x1 + x2)) / 2.0), x2]; ys = [y1, f32((np.float64(y1 + y2)) / 2.0), y2]
        d1 = np.array([x1 - x2, y1 - y2], f32); d1 = d1 / f32(np.sqrt(d1[0] * d1[0] + d1[1] * d1[1]))
        d2x = kl['startPointX'] - kl['endPointX']; d2y = kl['startPointY'] - kl['endPointY']
        n2 = np.sqrt(d2x * d2x + d2y * d2y).astype(f32)
        cs = np.abs(d1[0] * (d2x / n2) + d1[1] * (d2y / n2))
        expect = set()
        iw, ih = f32(64) / f32(640), f32(48) / f32(480)
        for x, y in zip(xs, ys):
            cx0, cx1 = max(0, int(np.floor((x - r) * iw))), min(63, int(np.ceil((x + r) * iw)))
            cy0, cy1 = max(0, int(np.floor((y - r) * ih))), min(47, int(np.ceil((y + r) * ih)))
            if cx0 >= 64 or cx1 < 0 or cy0 >= 48 or cy1 < 0:
                continue
            dist = (lv[:, 0] * np.float64(x) + lv[:, 1] * np.float64(y) + lv[:, 2]).astype(f32)
            for i in np.nonzero(~(cs < TH) & (np.abs(dist) < r))[0]:
                if any(cx0 <= c // 48 <= cx1 and cy0 <= c % 48 <= cy1 for c in _bresenham_cells(kl, i)):
                    expect.add(int(i))
        assert set(got.tolist()) == expect


def test_oracle_line_greedy_conflict():
    # two parallel frame lines 2 px apart, descriptors at distance 0 / 8 from the query descriptor; two identical queries
    kl = np.zeros(2, oracle.KL_DTYPE)
    kl['startPointX'] = 100; kl['endPointX'] = 300; kl['startPointY'] = [200, 202]; kl['endPointY'] = [200, 202]
    kl['sPointInOctaveX'] = 100; kl['ePointInOctaveX'] = 300; kl['sPointInOctaveY'] = kl['startPointY']; kl['ePointInOctaveY'] = kl['endPointY']
    kl['lineLength'] = 200
    lv = np.array([[0, 1, -200.0], [0, 1, -202.0]])
    d = np.zeros((2, 32), np.uint8); d[1, 0] = 0xff
    qd = np.zeros((2, 32), np.uint8)
    q = np.zeros(2, oracle.LPROJ_QUERY_DTYPE)
    q['x1'] = 100; q['x2'] = 300; q['y1'] = 201; q['y2'] = 201; q['r'] = 8; q['cos_th'] = 0.96
    q['dir'][:, 0] = 200; q['length'] = 200; q['claims'] = 1
    idx, dist, n = oracle.line_search_projection(kl, lv, d, None, BOUNDS, q, qd, mode=1)
    assert idx.tolist() == [0, 1] and dist.tolist() == [0, 8] and n == 2
    q['claims'] = 0                                            # without observations nothing is ever claimed
    idx, _, _ = oracle.line_search_projection(kl, lv, d, None, BOUNDS, q, qd, mode=1)
    assert idx.tolist() == [0, 0]
    q['length'] = 100                                          # length ratio 0.5 < 0.75: rejected
    idx, _, n = oracle.line_search_projection(kl, lv, d, None, BOUNDS, q, qd, mode=1)
    assert idx.tolist() == [-1, -1] and n == 0
    # mode 0: both lines are candidates at the same octave: best 0 vs second 8 passes the ratio, the second query (first line
    # taken) sees a single candidate (bestLevel2 = -1) and takes the other line
    l3 = np.array([[1, 0, 0, 0, 0, 0], [2, 0.01, 0, 0, 0, 0]], np.float64)
    q['claims'] = 1; q['dir'] = [1.0, 0, 0]; q['cos_th'] = 0.998
    idx, dist, n = oracle.line_search_projection(kl, lv, d, l3, BOUNDS, q, qd, mode=0, nnratio=0.95)
    assert idx.tolist() == [0, 1] and n == 2
    q['dir'] = [0, 1.0, 0]                                     # perpendicular in 3-D: the direction gate removes every candidate
    idx, _, n = oracle.line_search_projection(kl, lv, d, l3, BOUNDS, q, qd, mode=0)
    assert idx.tolist() == [-1, -1] and n == 0


@pytest.mark.gpu
def test_gpu_line_grid_and_area_match_oracle(hvo, synth):
    kl, d, lv = _lines(synth, 0)
    pm = hvo.LineProjectionMatcher()
    pm.set_frame(kl, lv, d, None, *BOUNDS)
    cnt, items = pm.grid()
    rcnt, ritems = oracle.line_grid_build(kl, BOUNDS)
    assert np.array_equal(cnt, rcnt) and np.array_equal(items, ritems)
    rng = np.random.RandomState(3)
    hits = 0
    for t in range(200):
        j = rng.randint(len(kl))
        if t % 4 == 0:
            x1, y1, x2, y2 = rng.uniform(-30, 670), rng.uniform(-30, 510), rng.uniform(-30, 670), rng.uniform(-30, 510)
        else:
            x1, y1 = kl['startPointX'][j] + rng.normal(0, 2), kl['startPointY'][j] + rng.normal(0, 2)
            x2, y2 = kl['endPointX'][j] + rng.normal(0, 2), kl['endPointY'][j] + rng.normal(0, 2)
        r, TH = float(rng.choice([5, 8, 15, 24])), float(rng.choice([0.998, 0.96]))
        got = pm.GetFeaturesInAreaForLine(x1, y1, x2, y2, r, TH=TH)
        ref = oracle.line_features_in_area(kl, lv, BOUNDS, x1, y1, x2, y2, r, TH)
        assert np.array_equal(got, ref)                      # same lines in the same order
        hits += len(ref) > 0
    assert hits > 50
    pm.close()


@pytest.mark.gpu
@pytest.mark.parametrize('mode', [0, 1])
def test_gpu_line_search_projection_matches_oracle(hvo, synth, mode):
    kl, lv, d, l3, claimed, q, qd = _scenario(synth, seed=mode, mode=mode)
    pm = hvo.LineProjectionMatcher()
    pm.set_frame(kl, lv, d, l3, *BOUNDS)
    for cl in (None, claimed):
        idx, dist, nm = pm.search(q, qd, cl, mode, 0.95)
        ridx, rdist, rnm = oracle.line_search_projection(kl, lv, d, l3, BOUNDS, q, qd, cl, mode, 0.95)
        assert rnm > 20
        assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and nm == rnm
        assert pm.rounds() >= 1
    # every query identical: the fixed point needs one round per contested line, and still equals the sequential result
    q2 = q[:40].copy(); q2[:] = q[0]; q2['claims'] = 1
    qd2 = np.repeat(qd[:1], 40, axis=0)
    q2['r'] = 24
    idx, dist, nm = pm.search(q2, qd2, None, mode, 0.95)
    ridx, rdist, rnm = oracle.line_search_projection(kl, lv, d, l3, BOUNDS, q2, qd2, None, mode, 0.95)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and nm == rnm
    pm.close()


@pytest.mark.gpu
def test_gpu_lsdmatcher_search_by_projection_mirror(hvo, synth):
    kl, lv, d, l3, claimed, q, qd = _scenario(synth, seed=5, mode=0)
    F = dict(keylines_un=kl, line_functions=lv, ldesc=d, lines3d=l3, bounds=BOUNDS, mapline=np.full(len(kl), -1, np.int32),
             claimed=claimed.astype(bool).copy())
    M = len(q)
    rng = np.random.RandomState(9)
    MLs = dict(proj_x1=q['x1'], proj_y1=q['y1'], proj_x2=q['x2'], proj_y2=q['y2'], view_cos=np.where(q['r'] % 5 == 0, 0.999, 0.9),
               in_view=rng.rand(M) > 0.1, bad=rng.rand(M) < 0.05, has_obs=q['claims'].astype(bool), world_vector=q['dir'], desc=qd)
    m = hvo.LSDmatcher(0.95)
    nm, match = m.SearchByProjection(F, MLs, True, th=3.0)
    sel = np.nonzero(MLs['in_view'] & ~MLs['bad'])[0]
    qq = q[sel].copy()
    qq['r'] = (np.where(np.asarray(MLs['view_cos'], np.float32)[sel] > np.float32(0.998), np.float32(5), np.float32(8)) * np.float32(3)).astype(np.float32)
    qq['cos_th'] = np.float32(0.998)
    ridx, _, rnm = oracle.line_search_projection(kl, lv, d, l3, BOUNDS, qq, qd[sel], claimed, 0, 0.95)
    assert nm == rnm > 5 and np.array_equal(match[sel], ridx) and np.all(match[np.setdiff1d(np.arange(M), sel)] == -1)
    for k, i in zip(sel, ridx):
        if i >= 0:
            assert F['mapline'][i] >= 0
    # last-frame variant
    kl1, d1, _ = _lines(synth, 1)
    last = dict(proj_x1=kl1['startPointX'], proj_y1=kl1['startPointY'], proj_x2=kl1['endPointX'], proj_y2=kl1['endPointY'], keylines=kl1,
                has_obs=np.ones(len(kl1), bool), desc=d1)
    Cur = dict(keylines_un=kl, line_functions=lv, ldesc=d, bounds=BOUNDS, mapline=np.full(len(kl), -1, np.int32), claimed=np.zeros(len(kl), bool))
    nm2, match2 = m.SearchByProjectionLast(Cur, last, 15.0)
    q1 = np.zeros(len(kl1), oracle.LPROJ_QUERY_DTYPE)
    q1['x1'] = kl1['startPointX']; q1['y1'] = kl1['startPointY']; q1['x2'] = kl1['endPointX']; q1['y2'] = kl1['endPointY']
    q1['r'] = 15; q1['cos_th'] = np.float32(0.96); q1['claims'] = 1; q1['length'] = kl1['lineLength']
    q1['dir'][:, 0] = kl1['ePointInOctaveX'] - kl1['sPointInOctaveX']; q1['dir'][:, 1] = kl1['ePointInOctaveY'] - kl1['sPointInOctaveY']
    ridx2, _, rnm2 = oracle.line_search_projection(kl, lv, d, None, BOUNDS, q1, d1, None, 1, 0.95)
    assert nm2 == rnm2 > 10 and np.array_equal(match2, ridx2)
    assert (Cur['mapline'] >= 0).sum() == len(set(ridx2[ridx2 >= 0].tolist()))
    m.close()
