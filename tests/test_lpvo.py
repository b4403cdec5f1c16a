"""Manhattan::computeNormalsLPVO (reference src/Manhattan.cpp:237-393): oracle pins and CUDA parity.
Bar: the two OpenCV primitives inside it (cv::integral 32F->64F, cv::normalize) bit-exact vs cv2 4.13.0; the function itself EXECUTED
(oracle/_ref/ref_lpvo, fixture tests/golden/lpvo_ref.npz); oracle and CUDA return the reference's samples in its order with bit-identical
doubles (they add in cv::integral's order)."""
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cam(synth, cfg):
    c = synth.CONFIGS[cfg]
    return dict(factor=np.float32(1.0 / c['factor']), fx=c['fx'], fy=c['fy'], cx=c['cx'], cy=c['cy'])


def test_oracle_integral_and_normalize_match_cv2_golden():
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'lpvo_cv2.npz'))
    assert np.array_equal(oracle.integral_f32(g['img']), g['integral'])
    got = np.stack([oracle.normalize3(v) for v in g['vec']])
    assert np.array_equal(got, g['normalized'])
    assert np.all(got[0] == 0)                                   # zero vector: scale 0, not NaN


def test_oracle_integral_and_normalize_match_cv2_live():
    cv2 = pytest.importorskip('cv2')
    r = np.random.RandomState(5)
    img = (r.randn(480, 640) * 3).astype(np.float32)
    assert np.array_equal(oracle.integral_f32(img), cv2.integral(img)[1:, 1:])
    for v in r.randn(500, 3):
        assert np.array_equal(oracle.normalize3(v), cv2.normalize(v.reshape(3, 1), None).ravel())


def test_oracle_lpvo_on_the_synthetic_room(synth):
    cam = _cam(synth, 'S1')
    _, d = synth.frame('S1', 0)
    n, z, pix = oracle.lpvo_normals(d, **cam)
    assert 1000 < len(n) <= 32 * 42
    assert np.all(pix[:, 0] % 15 == 10) and np.all(pix[:, 1] % 15 == 10)
    order = pix[:, 1] * 1000 + pix[:, 0]
    assert np.all(np.diff(order) > 0)                            # row-major push_back order
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-12)
    assert np.all((z > 0.2) & (z < 7.0))
    # v x u of a surface seen by the camera points back at it: the back wall gives ~(0, 0, -1)
    assert np.mean(n[:, 2] < -0.9) > 0.3
    # an empty depth image has no valid tangent anywhere
    n0, _, _ = oracle.lpvo_normals(np.zeros_like(d), **cam)
    assert len(n0) == 0


GOLDEN_REF = os.path.join(ROOT, 'tests', 'golden', 'lpvo_ref.npz')
REF_CASES = [('S1', 0), ('S2', 3), ('S1', 5)]


def _holes(d, k):
    """case k > 0 of a frame: a hole, a strip at the range limits (z = 0.2 exactly; z = 8 > 7)"""
    d = d.copy()
    if k:
        d[100:300, 200:400] = 0
        d[:, :40] = 1000
        d[:, 600:] = 40000
    return d


def _ref_lpvo(synth, cfg, idx, k):
    """the executed reference's (normals, depth) for a case: live when oracle/_ref/ref_lpvo exists (and equal to the fixture), else the fixture"""
    cam = _cam(synth, cfg)
    d = _holes(synth.frame(cfg, idx)[1], k)
    key = f'{cfg}_{idx}_{k}'
    g = np.load(GOLDEN_REF) if os.path.exists(GOLDEN_REF) else None
    live = oracle.ref_lpvo(d, **cam)
    if live is not None:
        if g is not None and key + '_n' in g:
            assert live[0].tobytes() == g[key + '_n'].tobytes() and live[1].tobytes() == g[key + '_z'].tobytes(), 'fixture differs from the live reference'
        return d, cam, live[0], live[1]
    if g is None:
        pytest.skip('neither oracle/_ref/ref_lpvo nor tests/golden/lpvo_ref.npz is available')
    return d, cam, g[key + '_n'], g[key + '_z']


@pytest.mark.parametrize('cfg,idx', REF_CASES)
def test_oracle_lpvo_equals_executed_reference(synth, cfg, idx):
    """Manhattan::computeNormalsLPVO executed (src/Manhattan.cpp:237-393 + removeMatRow / removeMatCol, cv::Rect body; oracle/ref_lpvo_main.cpp):
    same samples in the same order, doubles bit for bit."""
    for k in (0, 1):
        d, cam, rn, rz = _ref_lpvo(synth, cfg, idx, k)
        n, z, _ = oracle.lpvo_normals(d, **cam)
        assert len(rn) > 300 and n.tobytes() == rn.tobytes() and z.tobytes() == rz.tobytes()


def test_reference_lpvo_as_built_corrupts_its_integral_images(synth):
    """Manhattan.cpp is built WITHOUT USE_CV_RECT (only src/Frame.cc:32 defines it): removeMatRow / removeMatCol then move
    width * sizeof(float) bytes per row of a CV_64F integral image.  Executed, that body turns most normals into NaN: the pin above
    uses the cv::Rect body, which does what the function's comment says."""
    cam = _cam(synth, 'S1')
    d = synth.frame('S1', 0)[1]
    good, bad = oracle.ref_lpvo(d, **cam), oracle.ref_lpvo(d, as_built=True, **cam)
    if good is None or bad is None:
        pytest.skip('oracle/_ref/ref_lpvo is not built (reference tree absent)')
    assert len(good[0]) == len(bad[0]) and not np.isnan(good[0]).any()
    assert np.isnan(bad[0]).mean() > 0.5


@pytest.mark.gpu
@pytest.mark.parametrize('cfg,idx', REF_CASES)
def test_gpu_lpvo_equals_executed_reference(hvo, synth, cfg, idx):
    K = None
    for k in (0, 1):
        d, cam, rn, rz = _ref_lpvo(synth, cfg, idx, k)
        if K is None:
            K = np.array([[cam['fx'], 0, cam['cx']], [0, cam['fy'], cam['cy']], [0, 0, 1]], np.float32)
            m = hvo.Manhattan(K, 640, 480, cam['factor'])
        n, z, _ = m.computeNormalsLPVO(d)
        assert n.tobytes() == rn.tobytes() and z.tobytes() == rz.tobytes()
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize('cfg,idx', [('S1', 0), ('S2', 3)])
def test_gpu_lpvo_matches_oracle_bit_for_bit(hvo, synth, cfg, idx):
    cam = _cam(synth, cfg)
    _, d = synth.frame(cfg, idx)
    K = np.array([[cam['fx'], 0, cam['cx']], [0, cam['fy'], cam['cy']], [0, 0, 1]], np.float32)
    m = hvo.Manhattan(K, 640, 480, cam['factor'])
    assert m.capacity == 32 * 42
    n, z, pix = m.computeNormalsLPVO(d)
    rn, rz, rpix = oracle.lpvo_normals(d, **cam)
    assert len(n) == len(rn) > 500
    assert np.array_equal(pix, rpix) and np.array_equal(z, rz)
    assert np.array_equal(n, rn)
    m.close()


@pytest.mark.gpu
def test_gpu_lpvo_batch_holes_and_range_limits(hvo, synth):
    cam = _cam(synth, 'S1')
    _, depths = synth.sequence('S1', 4, start=70)
    depths = depths.copy()
    depths[1] = 0                                                # nothing valid
    depths[2, 100:300, 200:400] = 0                              # a hole: samples whose 5-point stencil touches it disappear
    depths[3, :, :320] = 1000                                    # z = 0.2 exactly (factor 1/5000): inside the mask range, vertex map zero
    depths[3, :, 320:] = 40000                                   # z = 8 > 7
    K = np.array([[cam['fx'], 0, cam['cx']], [0, cam['fy'], cam['cy']], [0, 0, 1]], np.float32)
    m = hvo.Manhattan(K, 640, 480, cam['factor'], max_batch=4)
    got = m.computeNormalsLPVO_batch(depths)
    for f in range(4):
        rn, rz, rpix = oracle.lpvo_normals(depths[f], **cam)
        n, z, pix = got[f]
        assert len(n) == len(rn)
        assert np.array_equal(pix, rpix) and np.array_equal(z, rz) and np.array_equal(n, rn)
    assert len(got[1][0]) == 0 and 0 < len(got[2][0]) < len(got[0][0])
    with pytest.raises(hvo.HvoError):
        m.computeNormalsLPVO_batch(np.zeros((5, 480, 640), np.uint16))   # more frames than max_batch
    m.close()
