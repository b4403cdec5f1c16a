"""CPU: the oracle's ORB restatement against (a) golden outputs of the reference's own ORBextractor.cc
(tests/golden/orb_ref.npz, produced by oracle/_ref/ref_orb) and (b) that binary run live when present."""
import numpy as np
import pytest

import oracle


def _params(g, name):
    nf, nl, ini, mn = (int(v) for v in g[name + '_params'])
    return dict(nfeatures=nf, scale_factor=float(g[name + '_scale']), nlevels=nl, ini_th=ini, min_th=mn)


@pytest.mark.parametrize('name', ['s1_crop', 's2_crop', 'noise'])
def test_oracle_orb_equals_reference_golden(golden_orb, name):
    g = golden_orb
    o = oracle.OrbOracle(**_params(g, name))
    kps, desc = o.extract(g[name + '_img'])
    assert len(kps) == len(g[name + '_kps'])
    assert kps.tobytes() == g[name + '_kps'].tobytes()       # position, size, angle, response, octave: bit-exact
    assert np.array_equal(desc, g[name + '_desc'])


def test_tables_match_reference_constants():
    sf, isf, nfeat, umax = oracle.OrbOracle().tables()
    assert list(nfeat) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(umax) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert sf[1] == np.float32(1.2) and sf[2] == np.float32(np.float32(1.2) * np.float64(np.float32(1.2)))
    sf2, _, nfeat2, _ = oracle.OrbOracle(nfeatures=2000).tables()
    assert list(nfeat2) == [434, 362, 302, 251, 209, 175, 145, 122]


def test_oracle_orb_equals_reference_binary_live(synth):
    if oracle.ref_orb_path() is None:
        pytest.skip('oracle/_ref/ref_orb not built (reference tree not mounted)')
    frames = np.stack([synth.frame('S1', 5)[0], synth.frame('S2', 2)[0], synth.noise_frame(640, 480, 9)])
    ref = oracle.ref_orb_extract(frames)
    o = oracle.OrbOracle()
    for f, (rk, rd) in zip(frames, ref):
        k, d = o.extract(f)
        assert k.tobytes() == rk.tobytes() and np.array_equal(d, rd)


def test_level_major_order_and_quota(synth):
    o = oracle.OrbOracle()
    kps, desc = o.extract(synth.frame('S1', 1)[0])
    assert np.all(np.diff(kps['octave']) >= 0)
    _, _, nfeat, _ = o.tables()
    for l in range(8):
        n = int((kps['octave'] == l).sum())
        assert n <= nfeat[l] + 3
    assert desc.shape == (len(kps), 32)


def test_empty_and_flat_images():
    o = oracle.OrbOracle()
    kps, desc = o.extract(np.full((480, 640), 127, np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)


def test_distribute_keeps_best_response_per_node_in_list_order():
    cand = np.array([[10, 10, 30], [11, 10, 50], [500, 400, 25], [501, 401, 25]], np.float32)
    out = oracle.distribute(cand, 16, 624, 16, 464, 2)
    # children are pushed to the list FRONT (TL first, BR last), so BR comes out first; equal responses keep
    # the earlier candidate (ORBextractor.cc:739-758)
    assert out.tolist() == [[500.0, 400.0, 25.0], [11.0, 10.0, 50.0]]


def test_distribute_stops_when_a_split_yields_a_single_child():
    # all keys fall in the top-left quadrant: the list size does not grow and the reference stops
    # (ORBextractor.cc:667-670) although the node holds several keys
    cand = np.array([[10, 10, 30], [11, 10, 50], [300, 200, 25], [301, 201, 25]], np.float32)
    out = oracle.distribute(cand, 16, 624, 16, 464, 2)
    assert out.tolist() == [[11.0, 10.0, 50.0]]
