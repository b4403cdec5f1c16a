"""Surface normals (PCL-style integral-image normals of Frame::ComputePlanes): CUDA vs the oracle restatement.
Bar: distance map and NaN pattern identical, positions bit-exact, normals within 1e-3 rad (north-star tolerance)."""
import numpy as np
import pytest

import oracle


def _cam(synth, cfg):
    c = synth.CONFIGS[cfg]
    return dict(factor=np.float32(1.0 / c['factor']), fx=c['fx'], fy=c['fy'], cx=c['cx'], cy=c['cy'])


def test_oracle_normals_are_unit_and_face_the_camera(synth):
    cam = _cam(synth, 'S1')
    _, d = synth.frame('S1', 0)
    out, dist = oracle.surface_normals(d, **cam, want_dist=True)
    assert out.shape == (80 * 107, 8) and dist.shape == (160, 214)
    ok = np.isfinite(out[:, 0])
    assert 0.5 < ok.mean() < 0.95                                   # 10-px border and depth edges are NaN
    n, p = out[ok, :3], out[ok, 3:6]
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    assert np.all(np.sum(n * p, axis=1) <= 1e-6)                    # flipped towards the origin
    assert np.array_equal(out[:, 6] % 6, np.full(len(out), 3)) and np.array_equal(out[:, 7] % 6, np.full(len(out), 3))
    assert dist.min() == 0 and {round(float(v), 3) for v in np.unique(dist[dist < 3])} <= {0.0, 1.0, 1.4, 2.0, 2.4, 2.8}
    # the back wall of the synthetic room dominates: most normals point along -z
    assert np.mean(n[:, 2] < -0.9) > 0.3


@pytest.mark.gpu
@pytest.mark.parametrize('cfg,idx', [('S1', 0), ('S2', 3), ('S3', 1)])
def test_gpu_surface_normals(hvo, synth, cfg, idx):
    cam = _cam(synth, cfg)
    _, d = synth.frame(cfg, idx)
    h, w = d.shape
    sn = hvo.SurfaceNormals(w, h, cam['fx'], cam['fy'], cam['cx'], cam['cy'], cam['factor'])
    got = sn.compute(d)
    ref, rdist = oracle.surface_normals(d, **cam, want_dist=True)
    assert np.array_equal(sn.distance_map(), rdist)                       # chamfer map: every float identical
    assert got.shape == ref.shape
    assert np.array_equal(got[:, 3:], ref[:, 3:])                          # positions bit-exact
    nan_g, nan_r = np.isnan(got[:, 0]), np.isnan(ref[:, 0])
    assert np.array_equal(nan_g, nan_r) and (~nan_r).sum() > 1000
    cosang = np.sum(got[~nan_r, :3] * ref[~nan_r, :3], axis=1)
    assert np.all(np.arccos(np.clip(cosang, -1, 1)) < 1e-3)
    sn.close()


@pytest.mark.gpu
def test_gpu_surface_normals_batch_and_invalid_depth(hvo, synth):
    cam = _cam(synth, 'S1')
    _, depths = synth.sequence('S1', 3, start=50)
    depths[2] = 0                                                        # all invalid: points at the origin, zero gradients
    sn = hvo.SurfaceNormals(640, 480, cam['fx'], cam['fy'], cam['cx'], cam['cy'], cam['factor'], max_batch=3)
    got = sn.compute(depths)
    for f in range(3):
        ref = oracle.surface_normals(depths[f], **cam)
        assert np.array_equal(np.isnan(got[f, :, 0]), np.isnan(ref[:, 0]))
        ok = ~np.isnan(ref[:, 0])
        assert np.allclose(got[f][ok], ref[ok], atol=1e-5)
    assert np.isnan(got[2, :, 0]).all()                                  # |gy x gx| == 0 -> NaN, as PCL
    sn.close()
