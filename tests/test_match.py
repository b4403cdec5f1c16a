"""Descriptor matching: oracle vs cv2 golden (CPU) and CUDA vs oracle (GPU).  Bar: bit-exact indices."""
import numpy as np
import pytest

import oracle


def test_oracle_knn2_equals_cv2_bfmatcher_golden(golden_prims):
    g = golden_prims
    idx, dist = oracle.knn2(g['knn_q'], g['knn_t'])
    assert np.array_equal(idx, g['knn_idx'])                 # includes the planted ties: lower train index first
    assert np.array_equal(dist.astype(np.float32), g['knn_dist'])


def test_oracle_knn2_against_live_cv2(synth):
    cv2 = pytest.importorskip('cv2')
    q, t = synth.descriptors_S4(nq=150, nt=900, seed=11, planted=40, ties=10)
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(q, t, k=2)
    ref = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    idx, _ = oracle.knn2(q, t)
    assert np.array_equal(idx, ref)


def test_swar_distance_is_popcount():
    r = np.random.RandomState(0)
    a, b = r.randint(0, 256, (2, 50, 32)).astype(np.uint8)
    for x, y in zip(a, b):
        assert oracle.hamming(x, y) == int(np.unpackbits(x ^ y).sum())


def test_host_hamming_helper_needs_no_gpu(hvo):
    r = np.random.RandomState(1)
    a, b = r.randint(0, 256, (2, 32)).astype(np.uint8)
    assert hvo.LSDmatcher.DescriptorDistance(a, b) == oracle.hamming(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize('nq,nt', [(1, 2), (7, 3), (200, 200), (129, 513), (300, 5000), (2000, 50000)])
def test_gpu_knn2_bit_exact(hvo, synth, nq, nt):
    q, t = synth.descriptors_S4(nq=nq, nt=nt, seed=nq + nt, planted=min(nq, nt, 50) // 2, ties=min(nq, nt, 20) // 4)
    bf = hvo.BFMatcherHamming()
    idx, dist = bf.knnMatch2(q, t)
    oidx, odist = oracle.knn2(q, t)
    assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
    bf.close()


@pytest.mark.gpu
def test_gpu_knn2_equals_cv2_golden_and_edge_cases(hvo, golden_prims):
    g = golden_prims
    bf = hvo.BFMatcherHamming()
    idx, dist = bf.knnMatch2(g['knn_q'], g['knn_t'])
    assert np.array_equal(idx, g['knn_idx']) and np.array_equal(dist.astype(np.float32), g['knn_dist'])
    # all-identical train rows: best = 0, second = 1
    t = np.tile(g['knn_q'][:1], (40, 1))
    idx, dist = bf.knnMatch2(g['knn_q'][:1], t)
    assert idx.tolist() == [[0, 1]] and dist.tolist() == [[0, 0]]
    # one train row: no second neighbour; empty sets
    idx, dist = bf.knnMatch2(g['knn_q'][:3], g['knn_t'][:1])
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == -1).all() and (dist[:, 1] == -1).all()
    idx, dist = bf.knnMatch2(g['knn_q'][:3], np.empty((0, 32), np.uint8))
    assert (idx == -1).all()
    idx, dist = bf.knnMatch2(np.empty((0, 32), np.uint8), g['knn_t'])
    assert idx.shape == (0, 2)
    bf.close()


@pytest.mark.gpu
def test_lsdmatcher_match_and_frame_bf_match(hvo, synth):
    q, t = synth.descriptors_S4(nq=200, nt=5000, seed=5, planted=120, ties=10)
    m = hvo.LSDmatcher(0.95, True)
    for nnr in (0.9, 0.95, 0.8):
        n, m12 = m.match(q, t, nnr)
        on, om12 = oracle.match_nnr(q, t, nnr)
        assert n == on and np.array_equal(m12, om12)
        assert n >= 100
    q2, t2 = synth.descriptors_S4(nq=180, nt=200, seed=6, planted=100, ties=5)
    for TH in (80, 50):
        assert np.array_equal(m.FrameBFMatch(q2, t2, TH), oracle.frame_bf_match(q2, t2, 0.95, TH))
    with pytest.raises(hvo.HvoError):
        m.matchNNR(q, t[:1], 0.9)
    m.close()


@pytest.mark.gpu
def test_knn2_query_sharding_property(hvo, synth):
    """Size-independent property used by the multi-GPU stress: sharding the queries and concatenating equals
    the full run; splitting the train set and merging (best, second) in index order equals the full run."""
    q, t = synth.descriptors_S4(nq=500, nt=20000, seed=9)
    bf = hvo.BFMatcherHamming()
    idx, dist = bf.knnMatch2(q, t)
    parts = [bf.knnMatch2(q[i::3], t) for i in range(3)]
    for i in range(3):
        assert np.array_equal(parts[i][0], idx[i::3]) and np.array_equal(parts[i][1], dist[i::3])
    bf.close()


@pytest.mark.gpu
def test_lsdmatcher_search_double_and_by_descriptor(hvo, synth):
    """E7 / E9: two-way FrameBFMatch with cross-check (src/LSDmatcher.cpp:903-940) and SearchByDescriptor (:522-559)."""
    a, b = synth.descriptors_S4(nq=180, nt=200, seed=31, planted=120, ties=6)
    m = hvo.LSDmatcher(0.95, True)
    n, lm = m.SearchDouble(a, b)
    m12 = oracle.frame_bf_match(a, b, 0.95, 50)
    m21 = oracle.frame_bf_match(b, a, 0.95, 50)
    ref = np.array([j if j >= 0 and m21[j] == i else -1 for i, j in enumerate(m12)], np.int32)
    assert np.array_equal(lm, ref) and n == int((ref >= 0).sum()) and n > 50
    # two key-frame lines land on the same current-frame line; the second one holds no MapLine and must not overwrite the first
    a = a.copy(); a[101] = a[100]
    has = np.ones(len(a), bool); has[101] = False; has[::7] = False
    nm, got = m.SearchByDescriptor(a, b, has)
    idx, dist = oracle.knn2(a, b)
    exp = np.full(len(b), -1, np.int32)
    en = 0
    for q in range(len(a)):   # src/LSDmatcher.cpp:541-556
        if np.float32(dist[q, 0]) / np.float32(dist[q, 1]) < np.float32(1.0) / np.float32(1.5) and has[q]:
            exp[idx[q, 0]] = q
            en += 1
    assert np.array_equal(got, exp) and nm == en and (exp >= 0).sum() > 40
    assert idx[100, 0] == idx[101, 0] and got[idx[100, 0]] in (100, -1) and got[idx[100, 0]] != 101
    assert m.SearchDouble(a[:0], b)[0] == 0


def _distinctive_groups(synth, seed=3, ngroups=400):
    """groups of 'observations': noisy copies of a base descriptor, sizes 0..120 (most small, like real map points)"""
    rng = np.random.RandomState(seed)
    sizes = np.minimum(rng.geometric(0.15, ngroups), 120)
    sizes[:6] = [0, 1, 2, 3, 33, 120]
    base = rng.randint(0, 256, (ngroups, 32)).astype(np.uint8)
    descs = []
    for g, n in enumerate(sizes):
        d = np.repeat(base[g:g + 1], n, axis=0)
        flips = rng.rand(n, 256) < rng.uniform(0.0, 0.12)
        d ^= np.packbits(flips, axis=1)
        if n > 3 and g % 5 == 0:
            d[1] = d[0]                                  # exact duplicates: equal medians, the first row must win
        descs.append(d)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    return np.concatenate(descs) if off[-1] else np.zeros((0, 32), np.uint8), off


def test_oracle_distinctive_matches_a_numpy_statement(synth):
    desc, off = _distinctive_groups(synth, ngroups=60)
    bi, bm = oracle.distinctive(desc, off)
    for g in range(len(off) - 1):
        d = desc[off[g]:off[g + 1]]
        if len(d) == 0:
            assert bi[g] == -1
            continue
        D = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(2)
        med = np.sort(D, axis=1)[:, int(0.5 * (len(d) - 1))]
        assert bi[g] == int(np.argmin(med)) and bm[g] == med.min()


@pytest.mark.gpu
def test_gpu_distinctive_descriptors_bit_exact(hvo, synth):
    desc, off = _distinctive_groups(synth)
    bf = hvo.BFMatcherHamming()
    bi, bm = bf.distinctive(desc, off)
    ri, rm = oracle.distinctive(desc, off)
    assert np.array_equal(bi, ri) and np.array_equal(bm, rm)
    bi0, _ = bf.distinctive(np.zeros((0, 32), np.uint8), np.zeros(3, np.int32))   # only empty groups
    assert bi0.tolist() == [-1, -1]
    bf.close()
