"""Plane extraction pinned BY EXECUTION of the reference: src/PlaneExtractor.cpp + include/peac/*.hpp are compiled unmodified
(oracle/Makefile -> oracle/_ref/ref_peac, OpenCV / Eigen stand-ins) and their outputs are the committed fixture
tests/golden/peac_ref.npz.

  CPU: the oracle restatement (oracle/plane_oracle.cpp) == the fixture; == the reference binary run live on more frames;
       the reference's result does not move when the eigen-solver stand-in is perturbed (the one substitution).
  GPU: the CUDA path == the fixture, i.e. against the reference's own output, not against the restatement.

Bars: block validity / N, plane count, supports and pixel membership identical; normals within 1e-3 rad (north-star); centres
within 1e-9 m.  The reference leaves visit counters (-2..-6) in unassigned membership pixels; only labels >= 0 are consumed
(AHCPlaneFitter.hpp:362-368), so membership is compared as max(label, -1)."""
import os
import zlib

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [('S1', 0), ('S1', 9), ('S2', 3), ('S3', 1)]


@pytest.fixture(scope='module')
def golden():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'peac_ref.npz'))


def _cam(synth, cfg):
    c = synth.CONFIGS[cfg]
    return dict(factor=np.float32(1.0 / c['factor']), fx=c['fx'], fy=c['fy'], cx=c['cx'], cy=c['cy'])


def _K(cam):
    return np.array([[cam['fx'], 0, cam['cx']], [0, cam['fy'], cam['cy']], [0, 0, 1]], np.float32)


def _depth(synth, golden, cfg, idx):
    _, d = synth.frame(cfg, idx)
    assert np.uint32(zlib.crc32(d.tobytes())) == golden[f'{cfg}_{idx}_depth_crc'], 'synthetic input drifted from the fixture'
    return d


def _angle(a, b):
    return np.arccos(np.clip(np.abs(np.sum(a * b, axis=-1)), 0, 1))


def _check_blocks(got, golden, k):
    """got: [nb, 9] = queued, N, center[3], normal[3], mse (oracle.plane_blocks / PlaneDetection.blocks layout)."""
    N = golden[k + 'block_N'].astype(np.int64)
    nouse = golden[k + 'block_nouse'].astype(bool)
    assert np.array_equal(got[:, 1].astype(np.int64), N)                       # INIT_STRICT validity + depth discontinuity rule
    assert not np.any(got[nouse, 0])                                           # a rejected block is never queued
    ref = golden[k + 'blocks_every4']
    g = got[::4]
    ok = ref['N'] >= 4
    assert ok.sum() > 200
    assert np.allclose(g[ok, 2:5], ref['center'][ok], rtol=0, atol=1e-12)
    assert np.all(_angle(g[ok, 5:8], ref['normal'][ok]) < 1e-3)                # north-star tolerance (measured: < 1e-6)
    assert np.allclose(g[ok, 8], ref['mse'][ok], rtol=1e-6, atol=1e-15)


def _check_planes(n, normals, centers, supports, membership, golden, k):
    ref = golden[k + 'planes']
    assert n == len(ref) and n >= 3
    assert np.array_equal(np.asarray(supports, np.int64), ref['N'].astype(np.int64))
    assert np.all(_angle(np.asarray(normals), ref['normal']) < 1e-3)
    assert np.all(np.sum(np.asarray(normals) * ref['normal'], axis=1) > 0)     # same orientation (towards the camera)
    assert np.allclose(centers, ref['center'], rtol=0, atol=1e-9)
    mem = golden[k + 'membership'].astype(np.int32).ravel()
    assert np.array_equal(np.asarray(membership).ravel(), mem)
    assert [int((mem == i).sum()) for i in range(n)] == ref['nvertices'].tolist()   # plane_vertices_ sizes


@pytest.mark.parametrize('cfg,idx', CASES)
def test_oracle_equals_reference_golden(synth, golden, cfg, idx):
    cam = _cam(synth, cfg)
    d = _depth(synth, golden, cfg, idx)
    k = f'{cfg}_{idx}_'
    _check_blocks(oracle.plane_blocks(d, **cam), golden, k)
    n, planes, mem = oracle.plane_detect(d, **cam)
    _check_planes(n, planes[:, :3], planes[:, 3:6], planes[:, 6], mem, golden, k)


def test_oracle_equals_reference_binary_live(synth):
    if oracle.ref_bin('ref_peac') is None:
        pytest.skip('oracle/_ref/ref_peac not built (reference tree not mounted)')
    for cfg, idxs in (('S1', range(20, 32)), ('S2', range(10, 18))):
        cam = _cam(synth, cfg)
        ds = np.stack([synth.frame(cfg, i)[1] for i in idxs])
        ds[0, 100:140, 200:260] = 0                                 # a hole: INIT_STRICT rejects the blocks it touches
        ds[1, :, 320:] = 0                                          # half the frame without depth
        for d, r in zip(ds, oracle.ref_peac(ds, **cam)):
            n, planes, mem = oracle.plane_detect(d, **cam)
            assert n == len(r['planes'])
            assert np.array_equal(planes[:, 6].astype(np.int64), r['planes']['N'])
            assert np.array_equal(mem, np.maximum(r['membership'], -1))
            if n:
                assert np.all(_angle(planes[:, :3], r['planes']['normal']) < 1e-6)
                assert np.allclose(planes[:, 3:6], r['planes']['center'], rtol=0, atol=1e-9)
            blk = oracle.plane_blocks(d, **cam)
            assert np.array_equal(blk[:, 1].astype(np.int64), r['blocks']['N'])


def test_reference_result_is_insensitive_to_the_eigen_solver(synth):
    """The one substitution in the reference build is Eigen::SelfAdjointEigenSolver (oracle/eigenshim).  Perturbing its
    eigenvalues and eigenvectors by 1e-12 .. 1e-10 relative (far above the differences between correct double-precision solvers)
    must not flip a single membership pixel, a support or the plane count."""
    if oracle.ref_bin('ref_peac_perturb') is None:
        pytest.skip('oracle/_ref/ref_peac_perturb not built (reference tree not mounted)')
    for cfg, idxs in (('S1', (0, 9)), ('S2', (3,)), ('S3', (1,))):
        cam = _cam(synth, cfg)
        ds = np.stack([synth.frame(cfg, i)[1] for i in idxs])
        base = oracle.ref_peac(ds, **cam)
        for eps in (1e-12, -1e-12, 1e-10):
            for a, b in zip(base, oracle.ref_peac(ds, perturb=eps, **cam)):
                assert len(a['planes']) == len(b['planes'])
                assert np.array_equal(a['planes']['N'], b['planes']['N'])
                assert np.array_equal(np.maximum(a['membership'], -1), np.maximum(b['membership'], -1))


@pytest.mark.gpu
@pytest.mark.parametrize('cfg,idx', CASES)
def test_gpu_equals_reference_golden(hvo, synth, golden, cfg, idx):
    cam = _cam(synth, cfg)
    d = _depth(synth, golden, cfg, idx)
    h, w = d.shape
    k = f'{cfg}_{idx}_'
    pd = hvo.PlaneDetection(w, h)
    assert pd.readDepthImage(d, _K(cam), cam['factor'])
    n = pd.runPlaneDetection(h, w)
    _check_blocks(pd.blocks(), golden, k)
    _check_planes(n, pd.normals, pd.centers, pd.supports, pd.membership, golden, k)
    assert [len(v) for v in pd.plane_vertices_] == golden[k + 'planes']['nvertices'].tolist()
    pd.close()
