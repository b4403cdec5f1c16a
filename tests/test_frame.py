"""Frame-level front-end (hvo_frame_*): one call runs ORB + lines + planes + normals of a batch on three CUDA streams.
Bar: every output equals what the standalone extractors produce (which are checked against the oracle in their own
tests), and ORB / planes / normals also directly against the oracle here."""
import numpy as np
import pytest

import oracle


def _cam(synth, cfg):
    c = synth.CONFIGS[cfg]
    return c['fx'], c['fy'], c['cx'], c['cy'], 1.0 / c['factor']


@pytest.mark.gpu
def test_gpu_frame_front_end_matches_oracle_and_standalone(hvo, synth):
    fx, fy, cx, cy, df = _cam(synth, 'S1')
    gray, depth = synth.sequence('S1', 3, start=40)
    fe = hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, bf=40.0, max_batch=3)
    out = fe.extract_batch(gray, depth)
    ex = hvo.LINEextractor(1, 1.2, 200, 0.125)
    for f in range(3):
        # ORB: bit-exact vs the oracle
        ok, od = oracle.OrbOracle().extract(gray[f])
        n = int(out['kp_counts'][f])
        assert n == len(ok) and out['kps'][f, :n].tobytes() == ok.tobytes() and np.array_equal(out['desc'][f, :n], od)
        # ComputeStereoFromRGBD (Frame.cc:1940-1961)
        u, v = ok['x'].astype(np.int32), ok['y'].astype(np.int32)
        d = depth[f][v, u].astype(np.float32) * np.float32(df)
        valid = d > 0
        assert np.array_equal(out['kp_depth'][f, :n][valid], d[valid]) and np.all(out['kp_depth'][f, :n][~valid] == -1)
        # lines: same as the standalone extractor
        kl, desc, lv = ex(gray[f])
        nl = int(out['line_counts'][f])
        assert nl == len(kl) and out['keylines'][f, :nl].tobytes() == kl.tobytes()
        assert np.array_equal(out['line_desc'][f, :nl], desc) and np.array_equal(out['linevec3'][f, :nl], lv)
        # planes: vs the oracle
        on, op, om = oracle.plane_detect(depth[f], np.float32(df), fx, fy, cx, cy)
        assert int(out['n_planes'][f]) == on and np.array_equal(out['membership'][f], om)
        assert np.allclose(out['planes7'][f, :on, :6], op[:, :6], atol=1e-9)
        # normals: vs the oracle (NaN-aware)
        ref = oracle.surface_normals(depth[f], np.float32(df), fx, fy, cx, cy)
        got = out['normals8'][f]
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        assert np.allclose(np.nan_to_num(got), np.nan_to_num(ref), atol=2e-5)
    fe.close()


@pytest.mark.gpu
def test_gpu_frame_stage_subsets_and_errors(hvo, synth):
    fx, fy, cx, cy, df = _cam(synth, 'S2')
    gray, depth = synth.sequence('S2', 2, start=5)
    full = hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, max_batch=2).extract_batch(gray, depth)
    for stages, keys in ((hvo.STAGE_ORB, ('kps', 'desc', 'kp_counts')), (hvo.STAGE_LINES | hvo.STAGE_NORMALS, ('keylines', 'normals8')),
                         (hvo.STAGE_PLANES, ('n_planes', 'membership'))):
        fe = hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, stages=stages, max_batch=2)
        out = fe.extract_batch(gray, depth)
        for k in keys:
            if k in ('kps', 'desc'):
                for f in range(2):
                    n = int(out['kp_counts'][f])
                    assert out[k][f, :n].tobytes() == full[k][f, :n].tobytes()
            elif k == 'keylines':
                for f in range(2):
                    n = int(out['line_counts'][f])
                    assert n == int(full['line_counts'][f]) and out[k][f, :n].tobytes() == full[k][f, :n].tobytes()
            else:
                assert np.array_equal(out[k], full[k], equal_nan=True)
        with pytest.raises(hvo.HvoError):   # the device-resident call is limited to max_batch frames
            fe.extract_batch_device(1, 1, 3, {k: 1 for k in fe.output_shapes(3)})
        fe.close()
    with pytest.raises(hvo.HvoError):
        hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, stages=0)


def _same_outputs(hvo, a, b, n):
    for f in range(n):
        nk, nl = int(a['kp_counts'][f]), int(a['line_counts'][f])
        assert nk == int(b['kp_counts'][f]) and nl == int(b['line_counts'][f])
        for k in ('kps', 'desc', 'kp_depth', 'kp_uright'):
            assert a[k][f, :nk].tobytes() == b[k][f, :nk].tobytes(), k
        for k in ('keylines', 'line_desc', 'linevec3'):
            assert a[k][f, :nl].tobytes() == b[k][f, :nl].tobytes(), k
        npl = int(a['n_planes'][f])
        assert npl == int(b['n_planes'][f]) and np.array_equal(a['planes7'][f, :npl], b['planes7'][f, :npl])
        assert np.array_equal(a['normals8'][f], b['normals8'][f], equal_nan=True)


@pytest.mark.gpu
def test_gpu_frame_lanes_stream_chunks_and_byte_membership(hvo, synth):
    """The host call streams chunks of max_batch / lanes frames round-robin through the lanes, so it takes more frames
    than max_batch; every lane count gives the same bytes, and the one-byte labels equal the int32 membership image."""
    fx, fy, cx, cy, df = _cam(synth, 'S1')
    gray, depth = synth.sequence('S1', 7, start=11)
    one = hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, max_batch=7, lanes=1, line_cull=True, membership='both')
    assert (one.lanes, one.chunk) == (1, 7)
    ref = one.extract_batch(gray, depth)
    one.close()
    lab = ref['membership']
    assert np.array_equal(ref['membership8'], np.where(lab < 0, 255, lab).astype(np.uint8))
    for lanes, max_batch in ((2, 4), (3, 3), (2, 7)):
        fe = hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, max_batch=max_batch, lanes=lanes, line_cull=True, membership='u8')
        assert fe.lanes == lanes and fe.chunk == -(-max_batch // lanes)
        out = fe.extract_batch(gray, depth)       # 7 frames: several chunks per lane when max_batch < 7
        assert 'membership' not in out
        _same_outputs(hvo, ref, out, 7)
        assert np.array_equal(out['membership8'], ref['membership8'])
        out2 = fe.extract_batch(gray[:2], depth[:2])   # fewer frames than lanes * chunk
        _same_outputs(hvo, ref, out2, 2)
        fe.close()


@pytest.mark.gpu
def test_gpu_frame_async_calls_overlap_and_equal_blocking(hvo, synth):
    """hvo_frame_extract_batch_async: calls are queued back to back (the uploads of call k+1 overlap the tail of call k); after
    sync() every output set holds what the blocking call returns."""
    import torch
    fx, fy, cx, cy, df = _cam(synth, 'S1')
    gray, depth = synth.sequence('S1', 5, start=23)
    fe = hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, max_batch=4, lanes=2, line_cull=True, membership='u8')
    ref = fe.extract_batch(gray, depth)

    def pinned_like(a):
        t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
        v = t.numpy().view(a.dtype).reshape(a.shape)
        return t, v
    keep = []
    hg, g = pinned_like(gray); hd, d = pinned_like(depth)
    g[:] = gray; d[:] = depth
    outs = []
    for _ in range(3):
        o = {}
        for k, v in ref.items():
            t, a = pinned_like(v)
            a.view(np.uint8)[...] = 0xAB
            keep.append(t); o[k] = a
        outs.append(o)
    for o in outs:
        fe.extract_batch(g, d, out=o, wait=False)
    fe.sync()
    for o in outs:
        _same_outputs(hvo, ref, o, 5)
        assert np.array_equal(o['membership8'], ref['membership8'])
    fe.close()


_SERIAL_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import hvo_b200 as hvo
from hvo_b200 import synth
c = synth.CONFIGS['S1']
gray, depth = synth.sequence('S1', 5, start=23)
fe = hvo.FrameFrontEnd(640, 480, c['fx'], c['fy'], c['cx'], c['cy'], 1.0 / c['factor'], bf=40.0, max_batch=4, lanes=2, line_cull=True, membership='u8')
a = fe.extract_batch(gray, depth)          # 5 frames through 2 lanes of 2: three chunks, lanes alternate
b = fe.extract_batch(gray[:3], depth[:3])  # the next call starts on the other lane
fe.close()
np.savez(sys.argv[2], **{'a_' + k: v for k, v in a.items()}, **{'b_' + k: v for k, v in b.items()})
"""


@pytest.mark.gpu
def test_gpu_frame_serial_schedule_equals_side_by_side(hvo, synth, tmp_path):
    """Chunks of >= 2048 frames run their pipelines one after the other (planes, lines, ORB, normals) with the downloads on the lane's
    own streams and the depth uploaded ahead of the gray image; HVO_FRAME_SERIAL=1 forces that schedule on a small batch (the switch
    is read once per process, hence the child process).  Every output must equal the side-by-side schedule's."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode in ('0', '1'):
        out = str(tmp_path / f'serial{mode}.npz')
        env = dict(os.environ, HVO_FRAME_SERIAL=mode)
        subprocess.run([sys.executable, '-c', _SERIAL_SCRIPT, root, out], check=True, env=env, timeout=600)
        res[mode] = np.load(out)
    for pre, n in (('a_', 5), ('b_', 3)):
        a = {k[2:]: res['0'][k] for k in res['0'].files if k.startswith(pre)}
        b = {k[2:]: res['1'][k] for k in res['1'].files if k.startswith(pre)}
        _same_outputs(hvo, a, b, n)
        assert np.array_equal(a['membership8'], b['membership8'])
    fx, fy, cx, cy, df = _cam(synth, 'S1')
    gray, depth = synth.sequence('S1', 5, start=23)
    ok, od = oracle.OrbOracle().extract(gray[4])   # and the child's outputs are the real thing, not two equal failures
    n = int(res['1']['a_kp_counts'][4])
    assert n == len(ok) and np.array_equal(res['1']['a_desc'][4, :n], od)


@pytest.mark.gpu
def test_gpu_frame_reports_device_side_overflow(hvo, synth, monkeypatch):
    """The pipelines' fixed-capacity buffers cannot overflow by construction; HVO_DEBUG_* shrinks two of them so that the
    device-side fault flags fire.  hvo_frame_* must return HVO_ERR_OVERFLOW naming the frame and the pipeline, through
    extract_batch, through a queued call + sync, and recover on the next call (the record is reset when read)."""
    fx, fy, cx, cy, df = _cam(synth, 'S1')
    gray, depth = synth.sequence('S1', 3, start=11)
    flat = np.full((480, 640), 90, np.uint8)
    gray[0] = flat                                            # frame 0 has no line segments: the first faulty frame is 1
    # LSD segment buffer
    monkeypatch.setenv('HVO_DEBUG_LINE_SEGCAP', '20')
    fe = hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, max_batch=3, stages=hvo.STAGE_LINES | hvo.STAGE_ORB)
    monkeypatch.delenv('HVO_DEBUG_LINE_SEGCAP')
    with pytest.raises(hvo.HvoError) as e:
        fe.extract_batch(gray, depth)
    assert e.value.status == hvo.HVO_ERR_OVERFLOW and 'frame 1' in str(e.value) and 'LSD segment buffer' in str(e.value)
    out = fe.extract_batch(np.stack([flat] * 3), depth)      # no segments at all: no fault, and the old one does not linger
    assert (out['line_counts'] == 0).all()
    fe.extract_batch(gray, depth, wait=False)                 # queued call: the fault is reported by the sync
    with pytest.raises(hvo.HvoError) as e:
        fe.sync()
    assert e.value.status == hvo.HVO_ERR_OVERFLOW
    fe.close()
    # the standalone line extractor reports it too
    monkeypatch.setenv('HVO_DEBUG_LINE_SEGCAP', '20')
    ex = hvo.LINEextractor(1, 1.2, 200, 0.125, width=640, height=480)
    monkeypatch.delenv('HVO_DEBUG_LINE_SEGCAP')
    with pytest.raises(hvo.HvoError) as e:
        ex(gray[1])
    assert e.value.status == hvo.HVO_ERR_OVERFLOW
    ex.close()
    # plane refinement queue
    monkeypatch.setenv('HVO_DEBUG_PLANE_QCAP', '256')
    fe = hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, max_batch=3, stages=hvo.STAGE_PLANES)
    monkeypatch.delenv('HVO_DEBUG_PLANE_QCAP')
    with pytest.raises(hvo.HvoError) as e:
        fe.extract_batch(gray, depth)
    assert e.value.status == hvo.HVO_ERR_OVERFLOW and 'frame 0' in str(e.value) and 'plane refinement queue' in str(e.value)
    fe.close()
    # and with the real capacities the same frames are fine
    fe = hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, max_batch=3)
    fe.extract_batch(gray, depth)
    fe.close()


@pytest.mark.gpu
def test_gpu_rgbd_epilogue_of_a_distorted_camera(hvo, synth):
    """TUM1.yaml has k1 != 0: mvuRight is kpU.pt.x - bf / d on the UNDISTORTED keypoint (Frame.cc:1944, 1957).  With `distorted`
    the device leaves kp_uright at -1 and hvo_stereo_uright_from_depth forms it from cv::undistortPoints' output."""
    cv2 = pytest.importorskip('cv2')
    K = np.array([[517.306408, 0, 318.643040], [0, 516.469215, 255.313989], [0, 0, 1]], np.float32)      # Examples/RGB-D/TUM1.yaml
    dist = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)
    gray, depth = synth.sequence('S1', 2, start=3)
    ex = hvo.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=2)
    plain = ex.extract_batch(gray, depth16=depth, depth_factor=1.0 / 5000.0, bf=40.0)
    out = ex.extract_batch(gray, depth16=depth, depth_factor=1.0 / 5000.0, bf=40.0, distorted=True)
    assert np.array_equal(out['depth'], plain['depth'])                      # mvDepth is sampled at the distorted keypoint either way
    for f in range(2):
        n = int(out['counts'][f])
        assert (out['uright'][f, :n] == -1).all()
        kps = out['kps'][f, :n].copy()
        pts = np.stack([kps['x'], kps['y']], 1).reshape(-1, 1, 2).astype(np.float32)
        un = cv2.undistortPoints(pts, K, dist, None, K).reshape(-1, 2)    # Frame::UndistortKeyPoints (Frame.cc:1701-1731)
        kun = kps.copy(); kun['x'] = un[:, 0]; kun['y'] = un[:, 1]
        ur = hvo.ORBextractor.stereo_uright_from_depth(kun, out['depth'][f, :n], 40.0)
        d = out['depth'][f, :n]
        with np.errstate(divide='ignore'):
            exp = np.where(d > 0, kun['x'] - np.float32(40.0) / d, np.float32(-1)).astype(np.float32)
        assert np.array_equal(ur, exp)
        moved = np.abs(kun['x'] - kps['x']) > 0.5
        assert moved.any() and np.any(ur[moved & (d > 0)] != plain['uright'][f, :n][moved & (d > 0)])   # the distortion matters
    ex.close()


@pytest.mark.gpu
def test_gpu_frame_compact_host_outputs(hvo, synth):
    """membership4 (two pixels per byte) and normals3 (normal only) are the same results in 256 KB instead of 581 KB per frame:
    expanded on the host they equal the full outputs bit for bit."""
    fx, fy, cx, cy, df = _cam(synth, 'S1')
    gray, depth = synth.sequence('S1', 3, start=60)
    full = hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, max_batch=3, max_planes=15, membership='both')
    a = full.extract_batch(gray, depth)
    full.close()
    fe = hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, max_batch=3, max_planes=15, membership='u4', normals='n3')
    b = fe.extract_batch(gray, depth)
    assert 'membership' not in b and 'normals8' not in b and b['membership4'].shape == (3, 640 * 480 // 2)
    assert np.array_equal(fe.membership4_expand(b['membership4']), a['membership'])
    for f in range(3):
        n8 = fe.normals3_expand(b['normals3'][f], depth[f])
        assert n8.tobytes() == a['normals8'][f].tobytes()              # NaN rows included
    assert np.array_equal(a['n_planes'], b['n_planes']) and np.array_equal(a['kp_counts'], b['kp_counts'])
    for f in range(3):
        n, nl, npl = int(a['kp_counts'][f]), int(a['line_counts'][f]), int(a['n_planes'][f])
        assert a['kps'][f, :n].tobytes() == b['kps'][f, :n].tobytes() and np.array_equal(a['desc'][f, :n], b['desc'][f, :n])
        assert a['keylines'][f, :nl].tobytes() == b['keylines'][f, :nl].tobytes() and np.array_equal(a['planes7'][f, :npl], b['planes7'][f, :npl])
    fe.close()
    with pytest.raises(hvo.HvoError):                                     # 16 planes do not fit 4 bits
        hvo.FrameFrontEnd(640, 480, fx, fy, cx, cy, df, max_batch=1, max_planes=16, membership='u4').extract_batch(gray[:1], depth[:1])
