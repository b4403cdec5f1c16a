"""GPU parity: the CUDA ORB path (through the C ABI) against the CPU oracle, stage by stage and end to end.
Bar: bit-exact (pyramid bytes, FAST candidates, keypoint order/position/angle/response/octave, descriptors)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _assert_frame_equal(kps, desc, okps, odesc, tag=''):
    assert len(kps) == len(okps), f'{tag}: {len(kps)} vs oracle {len(okps)} keypoints'
    for fld in ('octave', 'x', 'y', 'response', 'size', 'angle', 'class_id'):
        bad = np.nonzero(kps[fld] != okps[fld])[0]
        assert len(bad) == 0, f'{tag}: field {fld} differs at {bad[:8]}: {kps[fld][bad[:4]]} vs {okps[fld][bad[:4]]}'
    assert kps.tobytes() == okps.tobytes(), tag
    bad = np.nonzero((desc != odesc).any(axis=1))[0]
    assert len(bad) == 0, f'{tag}: {len(bad)} descriptors differ, first {bad[:8]}'


@pytest.fixture(scope='module')
def ex640(hvo):
    e = hvo.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=8)
    yield e
    e.close()


def test_pyramid_levels_bit_exact(ex640, synth):
    gray, _ = synth.frame('S1', 0)
    ex640(gray)
    o = oracle.OrbOracle()
    o.extract(gray)
    for l in range(8):
        ref = o.level(l)['img']
        got = ex640.pyramid_level(0, l)
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), f'level {l}: {(got != ref).sum()} bytes differ'


@pytest.mark.parametrize('cfg,idx', [('S1', 0), ('S2', 1)])
def test_fast_candidates_equal_as_sets(ex640, synth, cfg, idx):
    gray, _ = synth.frame(cfg, idx)
    ex640(gray)
    o = oracle.OrbOracle()
    o.extract(gray)
    for l in range(8):
        ref = o.level(l)['cand']  # x, y relative to the 16-px border
        got = ex640.candidates(0, l)
        ref_set = sorted((int(x) + 16, int(y) + 16, int(s)) for x, y, s in ref)
        got_set = sorted(map(tuple, got.tolist()))
        assert got_set == ref_set, f'{cfg} level {l}: {len(got_set)} vs {len(ref_set)}'


@pytest.mark.parametrize('cfg,idx', [('S1', 0), ('S1', 7), ('S2', 0), ('S2', 3)])
def test_extract_bit_exact_640(ex640, synth, cfg, idx):
    gray, _ = synth.frame(cfg, idx)
    kps, desc = ex640(gray)
    okps, odesc = oracle.OrbOracle().extract(gray)
    assert len(okps) > 20
    _assert_frame_equal(kps, desc, okps, odesc, f'{cfg}[{idx}]')


def test_extract_bit_exact_dense_noise(ex640, synth):
    """Dense corners: every level overflows its quota, many quadtree rounds and size ties."""
    for seed in (1, 2):
        gray = synth.noise_frame(640, 480, seed)
        kps, desc = ex640(gray)
        okps, odesc = oracle.OrbOracle().extract(gray)
        _assert_frame_equal(kps, desc, okps, odesc, f'noise{seed}')


def test_extract_bit_exact_1280x720_2000(hvo, synth):
    gray, _ = synth.frame('S3', 0)
    e = hvo.ORBextractor(2000, 1.2, 8, 20, 7)
    kps, desc = e(gray)
    okps, odesc = oracle.OrbOracle(nfeatures=2000).extract(gray)
    _assert_frame_equal(kps, desc, okps, odesc, 'S3')
    assert list(e.GetFeaturesPerLevel()) == [434, 362, 302, 251, 209, 175, 145, 122]
    e.close()


@pytest.mark.parametrize('name', ['s1_crop', 's2_crop', 'noise'])
def test_extract_equals_reference_golden(hvo, golden_orb, name):
    """Against outputs of the reference's own ORBextractor.cc (tests/golden/orb_ref.npz)."""
    g = golden_orb
    nf, nl, ini, mn = (int(v) for v in g[name + '_params'])
    e = hvo.ORBextractor(nf, float(g[name + '_scale']), nl, ini, mn)
    kps, desc = e(g[name + '_img'])
    _assert_frame_equal(kps, desc, g[name + '_kps'], g[name + '_desc'], name)
    e.close()


def test_odd_sizes_and_strided_input(hvo, synth):
    """Width not a multiple of 4/16, ragged level sizes, non-contiguous rows."""
    big = synth.noise_frame(700, 500, 5)
    for (w, h) in [(333, 251), (501, 377), (642, 479)]:
        view = big[7:7 + h, 13:13 + w]  # strided view
        e = hvo.ORBextractor(500, 1.2, 6, 20, 7)
        kps, desc = e(view)
        okps, odesc = oracle.OrbOracle(nfeatures=500, nlevels=6).extract(np.ascontiguousarray(view))
        _assert_frame_equal(kps, desc, okps, odesc, f'{w}x{h}')
        e.close()


def test_flat_and_empty_images(ex640, hvo):
    kps, desc = ex640(np.full((480, 640), 90, np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)
    kps, desc = ex640(np.empty((0, 0), np.uint8))   # silent return, ORBextractor.cc:1044-1045
    assert len(kps) == 0
    with pytest.raises(hvo.HvoError):
        ex640(np.zeros((480, 640), np.float32))      # assert(type == CV_8UC1), ORBextractor.cc:1048


def test_batch_equals_single_and_rgbd_epilogue(ex640, synth):
    gray, depth = synth.sequence('S1', 6, start=20)
    out = ex640.extract_batch(gray, depth16=depth, depth_factor=1.0 / 5000.0, bf=40.0)
    o = oracle.OrbOracle()
    for i in range(6):
        n = int(out['counts'][i])
        okps, odesc = o.extract(gray[i])
        _assert_frame_equal(out['kps'][i, :n], out['desc'][i, :n], okps, odesc, f'batch[{i}]')
        # Frame::ComputeStereoFromRGBD (Frame.cc:1940-1961): d = depth(int(v), int(u)) * factor, valid iff 0 < d < 7
        u, v = okps['x'].astype(np.int32), okps['y'].astype(np.int32)
        d = depth[i][v, u].astype(np.float32) * np.float32(1.0 / 5000.0)
        ok = (d > 0) & (d < 7.0)
        exp_d = np.where(ok, d, np.float32(-1))
        with np.errstate(divide='ignore'):
            exp_r = np.where(ok, okps['x'] - np.float32(40.0) / d, np.float32(-1)).astype(np.float32)
        assert np.array_equal(out['depth'][i, :n], exp_d)
        assert np.array_equal(out['uright'][i, :n], exp_r)


def test_getters_match_reference_tables(ex640):
    sf, isf, nfeat, _ = oracle.OrbOracle().tables()
    assert np.array_equal(ex640.GetScaleFactors(), sf)
    assert np.array_equal(ex640.GetInverseScaleFactors(), isf)
    assert np.array_equal(ex640.GetFeaturesPerLevel(), nfeat)
    assert np.array_equal(ex640.GetScaleSigmaSquares(), sf * sf)
    assert ex640.GetLevels() == 8 and abs(ex640.GetScaleFactor() - 1.2) < 1e-12


def test_repeatability_full_batch(ex640, synth):
    """Size-independent property: the same frames in any batch slot give identical bytes (no cross-frame state,
    no dependence on the unordered atomics inside the kernels)."""
    gray, _ = synth.sequence('S1', 8, start=40)
    a = ex640.extract_batch(gray)
    b = ex640.extract_batch(np.ascontiguousarray(gray[::-1]))
    for i in range(8):
        n = int(a['counts'][i])
        assert n == int(b['counts'][7 - i])
        assert a['kps'][i, :n].tobytes() == b['kps'][7 - i, :n].tobytes()
        assert np.array_equal(a['desc'][i, :n], b['desc'][7 - i, :n])
