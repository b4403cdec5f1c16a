"""Plane extraction (PEAC/AHC): oracle self-checks on CPU; CUDA block statistics + host graph stage vs the oracle
on GPU.  Bar: block validity/N exact, normals within 1e-3 rad (north-star tolerance; in practice bit-identical),
plane count, supports and pixel membership identical."""
import numpy as np
import pytest

import oracle


def _cam(synth, cfg):
    c = synth.CONFIGS[cfg]
    return dict(factor=np.float32(1.0 / c['factor']), fx=c['fx'], fy=c['fy'], cx=c['cx'], cy=c['cy'])


def _K(cam):
    return np.array([[cam['fx'], 0, cam['cx']], [0, cam['fy'], cam['cy']], [0, 0, 1]], np.float32)


def test_oracle_eig33_against_numpy():
    r = np.random.RandomState(0)
    for i in range(300):
        A = r.randn(3, 3) * (10 ** r.uniform(-3, 2))
        K = A @ A.T
        if i % 4 == 0:
            K = np.diag(r.rand(3)) + 1e-9 * K          # nearly diagonal
        s, V = oracle.eig33sym(K)
        w, U = np.linalg.eigh(K)
        assert np.allclose(s, w, rtol=1e-10, atol=1e-12 * np.abs(w).max())
        if (w[1] - w[0]) > 1e-6 * w[2]:                 # smallest eigenvector = plane normal
            assert min(np.linalg.norm(V[:, 0] - U[:, 0]), np.linalg.norm(V[:, 0] + U[:, 0])) < 1e-6


def test_oracle_smallest_eigenpair_against_numpy(synth):
    """eig33_smallest (what Stats::compute uses on both the oracle and the CUDA side) vs numpy.linalg.eigh."""
    r = np.random.RandomState(1)
    for i in range(400):
        A = r.randn(3, 3) * (10 ** r.uniform(-3, 2))
        K = A @ A.T
        if i % 4 == 0:
            K = np.diag(r.rand(3)) + 1e-9 * K
        lam, v = oracle.eig33_smallest(K)
        w, U = np.linalg.eigh(K)
        assert abs(lam - w[0]) <= 1e-10 * w[2]
        if (w[1] - w[0]) > 1e-3 * w[2]:
            assert np.arccos(min(1.0, abs(v @ U[:, 0]))) < 1e-6
    # covariances of real plane patches: thin in one direction
    c = synth.CONFIGS['S1']
    _, d = synth.frame('S1', 0)
    z = d.astype(np.float64) / c['factor']
    ys, xs = np.mgrid[0:480, 0:640]
    P = np.stack([(xs - c['cx']) * z / c['fx'], (ys - c['cy']) * z / c['fy'], z], -1)
    for _ in range(300):
        bh, bw = r.randint(1, 20, 2)
        by, bx = r.randint(0, 48 - bh + 1), r.randint(0, 64 - bw + 1)
        pts = P[by * 10:(by + bh) * 10, bx * 10:(bx + bw) * 10].reshape(-1, 3)
        if (pts[:, 2] == 0).any():
            continue
        K = pts.T @ pts - np.outer(pts.sum(0), pts.sum(0)) / len(pts)
        lam, v = oracle.eig33_smallest(K)
        w, U = np.linalg.eigh(K)
        assert abs(lam - w[0]) <= 1e-11 * w[2] and np.arccos(min(1.0, abs(v @ U[:, 0]))) < 1e-6


def test_oracle_finds_the_room_planes(synth):
    for cfg, idx in (('S1', 0), ('S2', 3)):
        cam = _cam(synth, cfg)
        _, d = synth.frame(cfg, idx)
        n, planes, mem = oracle.plane_detect(d, **cam)
        assert 3 <= n <= 8
        assert np.allclose(np.linalg.norm(planes[:, :3], axis=1), 1.0, atol=1e-9)
        assert np.all(np.sum(planes[:, :3] * planes[:, 3:6], axis=1) <= 0)       # normals face the camera
        assert np.all(np.diff(planes[:, 6]) <= 0) and planes[-1, 6] >= 3000      # sorted by support, minSupport
        assert (mem >= 0).mean() > 0.8 and mem.max() == n - 1
        assert abs(planes[0, 2]) > 0.99                                           # back wall dominates
    # invalid depth everywhere -> no planes
    n, planes, mem = oracle.plane_detect(np.zeros((480, 640), np.uint16), **_cam(synth, 'S1'))
    assert n == 0 and (mem == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize('cfg,idx', [('S1', 0), ('S2', 3), ('S3', 1)])
def test_gpu_block_statistics(hvo, synth, cfg, idx):
    cam = _cam(synth, cfg)
    _, d = synth.frame(cfg, idx)
    h, w = d.shape
    pd = hvo.PlaneDetection(w, h)
    assert pd.readDepthImage(d, _K(cam), cam['factor'])
    pd.runPlaneDetection(h, w)
    got = pd.blocks()
    ref = oracle.plane_blocks(d, **cam)
    assert np.array_equal(got[:, 0], ref[:, 0]) and np.array_equal(got[:, 1], ref[:, 1])    # queued flag and N
    ok = ref[:, 1] >= 4
    assert ok.sum() > 1000
    cosang = np.abs(np.sum(got[ok, 5:8] * ref[ok, 5:8], axis=1))
    assert np.all(np.arccos(np.clip(cosang, -1, 1)) < 1e-3)                                 # north-star tolerance
    assert np.array_equal(got[ok], ref[ok])                                                  # in fact bit-identical
    pd.close()


@pytest.mark.gpu
@pytest.mark.parametrize('cfg,idx', [('S1', 0), ('S1', 9), ('S2', 3), ('S3', 1)])
def test_gpu_plane_detection_equals_oracle(hvo, synth, cfg, idx):
    cam = _cam(synth, cfg)
    _, d = synth.frame(cfg, idx)
    h, w = d.shape
    pd = hvo.PlaneDetection(w, h)
    pd.readDepthImage(d, _K(cam), cam['factor'])
    n = pd.runPlaneDetection(h, w)
    on, oplanes, omem = oracle.plane_detect(d, **cam)
    assert n == on and n >= 3
    assert np.array_equal(pd.supports, oplanes[:, 6].astype(np.int64))
    ang = np.arccos(np.clip(np.sum(pd.normals * oplanes[:, :3], axis=1), -1, 1))
    assert np.all(ang < 1e-3)
    assert np.allclose(pd.centers, oplanes[:, 3:6], atol=1e-9)
    assert np.array_equal(pd.membership, omem)
    assert [len(v) for v in pd.plane_vertices_] == [int((omem == i).sum()) for i in range(n)]
    pd.close()


@pytest.mark.gpu
def test_gpu_plane_edge_cases_and_batch(hvo, synth):
    cam = _cam(synth, 'S1')
    pd = hvo.PlaneDetection(640, 480, max_batch=4)
    assert pd.readDepthImage(np.zeros((480, 640), np.float32), _K(cam), cam['factor']) is False    # wrong type: reference returns false
    assert pd.readDepthImage(np.zeros((480, 640), np.uint16), _K(cam), cam['factor'])
    assert pd.runPlaneDetection() == 0 and (pd.membership == -1).all()                            # no valid depth
    _, depths = synth.sequence('S1', 4, start=30)
    n, planes, mem = pd.detect_batch(depths)
    for f in range(4):
        on, op, om = oracle.plane_detect(depths[f], **cam)
        assert n[f] == on and np.array_equal(mem[f], om)
        assert np.allclose(planes[f, :on, :6], op[:, :6], atol=1e-9)
    pd.close()
