"""Line segment detection + the full LINEextractor::operator() chain.

Oracle pin: oracle/lsd_oracle.cpp == cv2 4.13.0 createLineSegmentDetector().detect(), bit for bit (golden vectors
tests/golden/lsd_cv2.npz, and live when cv2 is importable).
GPU bar (north_star: "LSD endpoints within 0.5 px"): same number of segments in the same order, endpoints within
TOL_PX = 1e-3 px of the oracle (the CUDA path follows the same seed order and float32 accumulation; only
double-precision region sums are reduced in a different order), scaled image and seed order bit-exact; KeyLine
integer fields exact, LBD descriptors exact wherever the endpoints are bit-equal."""
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_PX = 1e-3
CASES = ['s1_crop', 's2_crop', 's1_odd', 'noise', 'flat']


@pytest.fixture(scope='module')
def golden_lsd():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'lsd_cv2.npz'))


@pytest.mark.parametrize('name', CASES)
def test_oracle_lsd_equals_cv2_golden(golden_lsd, name):
    seg, scaled = oracle.lsd_detect(golden_lsd[name + '_img'], want_scaled=True)
    assert np.array_equal(scaled, golden_lsd[name + '_scaled'])
    ref = golden_lsd[name + '_segments']
    assert seg.shape == ref.shape and np.array_equal(seg, ref)


@pytest.mark.parametrize('cfg,idx', [('S1', 0), ('S2', 1), ('S3', 2), ('S1', 9)])
def test_oracle_lsd_equals_cv2_live(synth, cfg, idx):
    cv2 = pytest.importorskip('cv2')
    cv2.setNumThreads(1)
    g, _ = synth.frame(cfg, idx)
    ref = cv2.createLineSegmentDetector().detect(g)[0]
    ref = np.zeros((0, 4), np.float32) if ref is None else ref.reshape(-1, 4)
    seg = oracle.lsd_detect(g)
    assert seg.shape == ref.shape and np.array_equal(seg, ref)
    assert len(seg) > 20


def test_oracle_segments_match_committed_lbd_fixture(synth):
    # tests/golden/lsd_cv2_segments.npz (used by the LBD tests) holds cv2 LSD output on S1/0 and S2/1
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'lsd_cv2_segments.npz'))
    for cfg, idx, key in [('S1', 0, 's1_segments'), ('S2', 1, 's2_segments')]:
        g, _ = synth.frame(cfg, idx)
        assert np.array_equal(oracle.lsd_detect(g), z[key])


def test_oracle_line_extract_truncates_by_response(synth):
    g, _ = synth.frame('S1', 0)
    kl, desc, lv = oracle.line_extract(g, n_features=200)
    assert len(kl) == 200 and desc.shape == (200, 32) and lv.shape == (200, 3)
    assert np.all(np.diff(kl['response']) <= 0) and np.all(kl['class_id'] == np.arange(200))
    assert np.allclose(np.hypot(lv[:, 0], lv[:, 1]), 1.0)
    kl2, _, _ = oracle.line_extract(g, n_features=10000)
    assert len(kl2) == len(oracle.lsd_detect(g)) and np.all(kl2['class_id'] == np.arange(len(kl2)))


def _seed_order_numpy(scaled):
    """ordered_points of the LSD restricted to defined pixels, from the scaled image (numpy restatement for the test)."""
    s = scaled.astype(np.int32)
    DA = s[1:, 1:] - s[:-1, :-1]
    BC = s[:-1, 1:] - s[1:, :-1]
    gx, gy = DA + BC, DA - BC
    norm = np.sqrt((gx * gx + gy * gy) / 4.0)
    rho = 2.0 / np.sin(np.pi * 22.5 / 180)
    defined = norm > rho
    if not defined.any():
        return np.zeros(0, np.uint32)
    coef = 1023.0 / norm[defined].max()
    bins = (norm * coef).astype(np.int32)
    ys, xs = np.nonzero(defined)
    idx = (ys * scaled.shape[1] + xs).astype(np.uint32)
    return idx[np.argsort(-bins[ys, xs], kind='stable')]


def _check_segments(got, ref):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if len(ref):
        assert np.abs(got - ref).max() <= TOL_PX, np.abs(got - ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_gpu_lsd_vs_cv2_golden(hvo, golden_lsd, name):
    img = golden_lsd[name + '_img']
    ex = hvo.LINEextractor(1, 1.2, 200, 0.125, width=img.shape[1], height=img.shape[0])
    (seg,) = ex.detect_segments(img)
    assert np.array_equal(ex.scaled_image(0), golden_lsd[name + '_scaled'])
    assert np.array_equal(ex.seed_order(0), _seed_order_numpy(golden_lsd[name + '_scaled']))
    _check_segments(seg, golden_lsd[name + '_segments'])


@pytest.mark.gpu
@pytest.mark.parametrize('cfg,idx', [('S1', 0), ('S2', 1), ('S3', 2), ('S1', 11), ('S2', 12)])
def test_gpu_lsd_vs_oracle_full_frames(hvo, synth, cfg, idx):
    g, _ = synth.frame(cfg, idx)
    ex = hvo.LINEextractor(1, 1.2, 200, 0.125, width=g.shape[1], height=g.shape[0])
    (seg,) = ex.detect_segments(g)
    ref, scaled = oracle.lsd_detect(g, want_scaled=True)
    assert np.array_equal(ex.scaled_image(0), scaled)
    assert np.array_equal(ex.seed_order(0), _seed_order_numpy(scaled))
    _check_segments(seg, ref)
    assert len(ref) > 20


@pytest.mark.gpu
def test_gpu_lsd_noise_and_flat(hvo, synth):
    ex = hvo.LINEextractor(1, 1.2, 200, 0.125, width=320, height=240, max_batch=3)
    frames = np.stack([synth.noise_frame(320, 240, 21), np.full((240, 320), 9, np.uint8), synth.noise_frame(320, 240, 22)])
    segs = ex.detect_segments(frames)
    for f, s in zip(frames, segs):
        _check_segments(s, oracle.lsd_detect(f))
    assert len(segs[1]) == 0


@pytest.mark.gpu
def test_gpu_line_extractor_matches_oracle(hvo, synth):
    """LINEextractor::operator(): keylines (all 17 fields), LBD descriptors and line functions."""
    for cfg, idx, nfeat in [('S1', 0, 200), ('S2', 1, 200), ('S1', 4, 50)]:
        g, _ = synth.frame(cfg, idx)
        ex = hvo.LINEextractor(1, 1.2, nfeat, 0.125)
        kl, desc, lv = ex(g)
        okl, odesc, olv = oracle.line_extract(g, n_features=nfeat)
        assert len(kl) == len(okl) and len(kl) > 10
        for name in ('class_id', 'octave', 'numOfPixels'):
            assert np.array_equal(kl[name], okl[name]), name
        for name in ('startPointX', 'startPointY', 'endPointX', 'endPointY', 'sPointInOctaveX', 'sPointInOctaveY',
                     'ePointInOctaveX', 'ePointInOctaveY', 'pt_x', 'pt_y'):
            assert np.abs(kl[name] - okl[name]).max() <= TOL_PX, name
        for name in ('angle', 'response', 'size', 'lineLength'):
            assert np.allclose(kl[name], okl[name], rtol=1e-5, atol=1e-4), name
        same = np.array([kl[i].tobytes() == okl[i].tobytes() for i in range(len(kl))])
        assert same.mean() > 0.9                      # in practice every keyline is bit-identical
        assert np.array_equal(desc[same], odesc[same])
        assert np.allclose(lv, olv, rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
def test_gpu_line_batch_equals_single(hvo, synth):
    frames = np.stack([synth.frame('S1', i)[0] for i in (0, 1)] + [synth.frame('S2', 2)[0]])
    ex = hvo.LINEextractor(1, 1.2, 200, 0.125, width=640, height=480, max_batch=3)
    out = ex.extract_batch(frames)
    ex1 = hvo.LINEextractor(1, 1.2, 200, 0.125)
    for i, f in enumerate(frames):
        kl, desc, lv = ex1(f)
        n = int(out['counts'][i])
        assert n == len(kl)
        assert out['keylines'][i, :n].tobytes() == kl.tobytes()
        assert np.array_equal(out['desc'][i, :n], desc)
        assert np.array_equal(out['linevec'][i, :n], lv)


@pytest.mark.gpu
def test_gpu_line_empty_image(hvo):
    ex = hvo.LINEextractor(1, 1.2, 200, 0.125)
    kl, desc, lv = ex(np.zeros((0, 0), np.uint8))
    assert len(kl) == 0 and desc.shape == (0, 32) and lv.shape == (0, 3)
