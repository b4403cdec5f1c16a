"""LBD line descriptor: CUDA (through the C ABI) vs the oracle restatement of the vendored
binary_descriptor_custom.cpp.  Bar: gradients bit-exact; descriptors bit-exact (the kernel keeps the reference's
float32 operation order, so no binary test can flip)."""
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _segments(name):
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'lsd_cv2_segments.npz'))[name + '_segments']


def _random_segments(n, w, h, seed):
    r = np.random.RandomState(seed)
    p = r.rand(n, 2) * [w, h]
    ang = r.rand(n) * 2 * np.pi
    ln = 8 + r.rand(n) * 250
    q = p + np.stack([np.cos(ang), np.sin(ang)], 1) * ln[:, None]
    seg = np.concatenate([p, q], 1).astype(np.float32)
    seg[0] = [-5, 10, 700, 10]           # clamped by checkLineExtremes
    seg[1] = [100.5, 100.5, 100.5, 100.5]  # zero-length: numOfPixels = 1
    seg[2] = [0, 0, 639, 479]
    return seg


def test_oracle_lbd_gradients_match_cv2_chain(synth):
    cv2 = pytest.importorskip('cv2')
    g, _ = synth.frame('S1', 2)
    bl = cv2.GaussianBlur(g, (5, 5), 1)
    dx, dy = oracle.lbd_gradients(g)
    assert np.array_equal(dx, cv2.Sobel(bl, cv2.CV_16S, 1, 0, ksize=3))
    assert np.array_equal(dy, cv2.Sobel(bl, cv2.CV_16S, 0, 1, ksize=3))


def test_oracle_keyline_fields(synth):
    kl = oracle.keylines_from_segments(_random_segments(20, 640, 480, 1), 640, 480)
    assert kl['startPointX'][0] == 0 and kl['endPointX'][0] == 639          # clamp to w-1
    assert kl['numOfPixels'][1] == 1 and kl['numOfPixels'][2] == 640
    assert np.all(kl['class_id'] == np.arange(20)) and np.all(kl['octave'] == 0)
    assert np.allclose(kl['response'], kl['lineLength'] / 640.0)


def test_oracle_lbd_properties(synth):
    g, _ = synth.frame('S1', 0)
    kl = oracle.keylines_from_segments(_segments('s1'), 640, 480)
    desc, fd = oracle.lbd_compute(g, kl, want_float=True)
    assert desc.shape == (len(kl), 32) and len(kl) > 50
    norms = np.linalg.norm(fd, axis=1)
    assert np.allclose(norms[np.isfinite(norms)], 1.0, atol=1e-5)
    # reversing a segment rotates the support region by pi: a different descriptor, but still a valid one
    assert 5 < np.unpackbits(desc, axis=1).mean(axis=1).mean() * 256 < 250


@pytest.mark.gpu
def test_gpu_lbd_gradients_bit_exact(hvo, synth):
    for cfg, (w, h) in (('S1', (640, 480)), ('S3', (1280, 720))):
        g, _ = synth.frame(cfg, 1)
        bd = hvo.BinaryDescriptor(w, h, max_lines=16)
        kl = oracle.keylines_from_segments(_random_segments(8, w, h, 2), w, h)
        bd.compute(g, kl)
        dx, dy = bd.gradients()
        odx, ody = oracle.lbd_gradients(g)
        assert np.array_equal(dx, odx) and np.array_equal(dy, ody)
        bd.close()
    g = synth.noise_frame(333, 251, 4)   # ragged tile edges
    bd = hvo.BinaryDescriptor(333, 251, max_lines=16)
    bd.compute(g, oracle.keylines_from_segments(_random_segments(8, 333, 251, 3), 333, 251))
    dx, dy = bd.gradients()
    odx, ody = oracle.lbd_gradients(g)
    assert np.array_equal(dx, odx) and np.array_equal(dy, ody)
    bd.close()


@pytest.mark.gpu
@pytest.mark.parametrize('cfg,idx,name', [('S1', 0, 's1'), ('S2', 1, 's2')])
def test_gpu_lbd_bit_exact_on_lsd_segments(hvo, synth, cfg, idx, name):
    g, _ = synth.frame(cfg, idx)
    kl = oracle.keylines_from_segments(_segments(name), 640, 480)
    bd = hvo.BinaryDescriptor(640, 480, max_lines=len(kl))
    desc = bd.compute(g, kl)
    odesc = oracle.lbd_compute(g, kl)
    bad = np.nonzero((desc != odesc).any(axis=1))[0]
    assert len(bad) == 0, f'{len(bad)} of {len(kl)} descriptors differ, first {bad[:5]}'
    bd.close()


@pytest.mark.gpu
def test_gpu_lbd_random_segments_batch_and_float(hvo, synth):
    frames, _ = synth.sequence('S1', 3, start=11)
    bd = hvo.BinaryDescriptor(640, 480, max_lines=200, max_batch=3)
    kls = np.zeros((3, 200), hvo.KL_DTYPE)
    counts = np.array([200, 57, 0], np.int32)
    for f in range(3):
        kls[f, :counts[f]] = oracle.keylines_from_segments(_random_segments(200, 640, 480, 20 + f)[:counts[f]], 640, 480)
    desc, fd = bd.compute_batch(frames, kls, counts, want_float=True)
    for f in range(3):
        od, ofd = oracle.lbd_compute(frames[f], kls[f, :counts[f]], want_float=True)
        assert np.array_equal(desc[f, :counts[f]], od)
        assert np.array_equal(fd[f, :counts[f]].view(np.uint32), ofd.view(np.uint32))  # float LBD bit-identical, NaNs included
    assert bd.compute(frames[0], np.empty(0, hvo.KL_DTYPE)).shape == (0, 32)  # "keypoint list is empty"
    bd.close()
