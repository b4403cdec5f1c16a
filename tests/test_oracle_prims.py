"""CPU: the oracle's restated OpenCV primitives against the committed cv2 4.13.0 golden vectors, and
against the live cv2 module when it is importable (it is in the build image)."""
import numpy as np
import pytest

import oracle


def test_resize_cascade_matches_cv2_golden(golden_prims):
    g = golden_prims
    cur = g['img']
    for i in range(4):
        ref = g[f'resize_cascade_{i}']
        out = oracle.resize_linear(cur, ref.shape[1], ref.shape[0])
        assert np.array_equal(out, ref), f'level {i}'
        cur = ref
    for i in range(2):
        ref = g[f'resize_single_{i}']
        assert np.array_equal(oracle.resize_linear(g['img'], ref.shape[1], ref.shape[0]), ref)


def test_blur_and_sobel_match_cv2_golden(golden_prims):
    g = golden_prims
    assert np.array_equal(oracle.blur7(g['img']), g['blur7'])
    assert np.array_equal(oracle.blur5(g['img']), g['blur5'])
    dx, dy = oracle.sobel3(g['img'])
    assert np.array_equal(dx, g['sobel_dx']) and np.array_equal(dy, g['sobel_dy'])


@pytest.mark.parametrize('thr', [20, 7])
def test_fast_matches_cv2_golden_including_order_and_response(golden_prims, thr):
    g = golden_prims
    img, rois, offs, kps = g['img'], g['fast_rois'], g[f'fast_{thr}_offs'], g[f'fast_{thr}_kps']
    assert len(kps) > 100
    for i, (x, y, w, h) in enumerate(rois):
        out = oracle.fast9(img[y:y + h, x:x + w], thr)
        assert np.array_equal(out, kps[offs[i]:offs[i + 1]]), f'roi {i}'


def test_fast_atan2_matches_cv2_golden(golden_prims):
    g = golden_prims
    assert np.array_equal(oracle.fast_atan2(g['atan2_y'], g['atan2_x']), g['atan2'])


def test_primitives_against_live_cv2(synth):
    cv2 = pytest.importorskip('cv2')
    img = synth.noise_frame(320, 240, 21)
    cur = img
    for (w, h) in [(267, 200), (222, 167), (185, 139)]:
        ref = cv2.resize(cur, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(oracle.resize_linear(cur, w, h), ref)
        cur = ref
    assert np.array_equal(oracle.blur7(img), cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
    assert np.array_equal(oracle.blur5(img), cv2.GaussianBlur(img, (5, 5), 1, 1, borderType=cv2.BORDER_REFLECT_101))
    for thr in (20, 7):
        det = cv2.FastFeatureDetector_create(threshold=thr, nonmaxSuppression=True)
        ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in det.detect(img)], np.int32).reshape(-1, 3)
        assert np.array_equal(oracle.fast9(img, thr), ref)
