"""The reference-facing C++ shim (shim/ORBextractor.h): compiles against the reference's call pattern and, on a
GPU, produces the same bytes as the oracle / the reference's own ORBextractor.cc through the same driver."""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'a-low-texture-robust-hybrid-feature-based-visual-odometry_b200')
EXE = os.path.join(ROOT, 'tests', 'cpp', 'shim_orb')


def _build():
    src = os.path.join(ROOT, 'tests', 'cpp', 'shim_orb_main.cpp')
    if not os.path.exists(os.path.join(PKG, 'libhvofront.so')):
        import __graft_entry__
        __graft_entry__.build()
    subprocess.check_call(['g++', '-O2', '-std=c++14', '-I' + os.path.join(ROOT, 'oracle', 'cvshim'), '-I' + os.path.join(PKG, 'shim'),
                           '-I' + os.path.join(ROOT, 'include'), src, '-o', EXE, '-L' + PKG, '-lhvofront',
                           '-Wl,-rpath,' + PKG])
    return EXE


def test_shim_compiles_against_reference_call_pattern():
    exe = _build()
    assert os.path.exists(exe)


def _run(exe, frames, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7):
    n, h, w = frames.shape
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, 'in.bin'), os.path.join(td, 'out.bin')
        with open(fi, 'wb') as f:
            f.write(struct.pack('<8i', 0x4f524231, w, h, n, nfeatures, nlevels, ini, mn) + struct.pack('<f', scale))
            f.write(np.ascontiguousarray(frames).tobytes())
        subprocess.check_call([exe, fi, fo])
        return open(fo, 'rb').read()


@pytest.mark.gpu
def test_shim_binary_equals_oracle_bytes(synth):
    exe = EXE if os.path.exists(EXE) else _build()
    frames = np.stack([synth.frame('S1', 3)[0], synth.frame('S2', 4)[0], synth.noise_frame(640, 480, 17)])
    raw = _run(exe, frames)
    o = oracle.OrbOracle()
    exp = b''
    for f in frames:
        k, d = o.extract(f)
        exp += struct.pack('<i', len(k)) + k.tobytes() + d.tobytes()
    assert raw == exp


# ---- lines / planes / brute-force matcher bridges -----------------------------------------------------------------
EXE2 = os.path.join(ROOT, 'tests', 'cpp', 'shim_front')


def _build_front():
    src = os.path.join(ROOT, 'tests', 'cpp', 'shim_front_main.cpp')
    if not os.path.exists(os.path.join(PKG, 'libhvofront.so')):
        import __graft_entry__
        __graft_entry__.build()
    subprocess.check_call(['g++', '-O2', '-std=c++14', '-I' + os.path.join(ROOT, 'oracle', 'cvshim'), '-I' + os.path.join(ROOT, 'tests', 'cpp', 'standins'),
                           '-I' + os.path.join(PKG, 'shim'), '-I' + os.path.join(ROOT, 'include'), src, '-o', EXE2, '-L' + PKG, '-lhvofront',
                           '-Wl,-rpath,' + PKG])
    return EXE2


def test_front_shims_compile_against_reference_call_pattern():
    assert os.path.exists(_build_front())


@pytest.mark.gpu
def test_front_shims_equal_python_mirror(hvo, synth):
    exe = EXE2 if os.path.exists(EXE2) else _build_front()
    gray, depth = synth.frame('S1', 6)
    c = synth.CONFIGS['S1']
    cam = np.array([c['fx'], c['fy'], c['cx'], c['cy'], 1.0 / c['factor']], np.float32)
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, 'in.bin'), os.path.join(td, 'out.bin')
        with open(fi, 'wb') as f:
            f.write(struct.pack('<3i', 0x46524e54, 640, 480) + gray.tobytes() + depth.tobytes() + cam.tobytes())
        subprocess.check_call([exe, fi, fo])
        raw = open(fo, 'rb').read()
    kl, desc, lv = hvo.LINEextractor(1, 1.2, 200, 0.125)(gray)
    nl = struct.unpack_from('<i', raw, 0)[0]
    off = 4
    assert nl == len(kl) and raw[off:off + 68 * nl] == kl.tobytes()
    off += 68 * nl
    assert raw[off:off + 32 * nl] == desc.tobytes()
    off += 32 * nl
    assert raw[off:off + 24 * nl] == lv.tobytes()
    off += 24 * nl
    assert raw[off:off + 32 * nl] == desc.tobytes()          # LBD recomputed on the same keylines (Frame.cc:1094-1096 path)
    off += 32 * nl
    pd = hvo.PlaneDetection(640, 480)
    K = np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], np.float32)
    pd.readDepthImage(depth, K, cam[4])
    n = pd.runPlaneDetection(480, 640)
    npl = struct.unpack_from('<i', raw, off)[0]
    off += 4
    assert npl == n and n >= 3
    for i in range(n):
        normal = np.frombuffer(raw, np.float64, 3, off); center = np.frombuffer(raw, np.float64, 3, off + 24)
        N, nv = struct.unpack_from('<2i', raw, off + 48)
        xyz = np.frombuffer(raw, np.float64, 3, off + 56)
        off += 80
        assert np.array_equal(normal, pd.normals[i]) and np.array_equal(center, pd.centers[i]) and N == pd.supports[i]
        assert nv == len(pd.plane_vertices_[i])
        j = int(pd.plane_vertices_[i][0])
        z = float(depth.ravel()[j]) * float(cam[4])
        assert abs(xyz[2] - z) < 1e-12 and abs(xyz[0] - ((j % 640) - float(cam[2])) * z / float(cam[0])) < 1e-12
    mem = np.frombuffer(raw, np.int32, 640 * 480, off)
    off += 4 * 640 * 480
    assert np.array_equal(mem, pd.membership)
    m12 = np.frombuffer(raw, np.int32, nl, off)
    cnt, ref = oracle.match_nnr(desc, desc[::-1].copy(), 0.95)
    assert np.array_equal(m12, ref)
    off += 4 * nl
    # LineWindowMatcher: the frame's own lines matched against itself, LSDmatcher::SearchByProjection(Cur, Last, 15) shape
    lw = np.frombuffer(raw, np.int32, nl, off)
    off += 4 * nl
    q = np.zeros(nl, oracle.LPROJ_QUERY_DTYPE)
    q['x1'] = kl['startPointX']; q['y1'] = kl['startPointY']; q['x2'] = kl['endPointX']; q['y2'] = kl['endPointY']
    q['r'] = 15; q['cos_th'] = np.float32(0.96); q['claims'] = 1; q['length'] = kl['lineLength']
    q['dir'][:, 0] = kl['ePointInOctaveX'] - kl['sPointInOctaveX']; q['dir'][:, 1] = kl['ePointInOctaveY'] - kl['sPointInOctaveY']
    ridx, _, rnm = oracle.line_search_projection(kl, lv, desc, None, (0.0, 0.0, 640.0, 480.0), q, desc, None, 1, 0.95)
    assert np.array_equal(lw, ridx) and rnm > nl // 2
    # LpvoNormals: Manhattan::computeNormalsLPVO
    nn = struct.unpack_from('<i', raw, off)[0]
    off += 4
    rn, rz, rpix = oracle.lpvo_normals(depth, np.float32(cam[4]), cam[0], cam[1], cam[2], cam[3])
    assert nn == len(rn) > 500
    assert np.array_equal(np.frombuffer(raw, np.float64, 3 * nn, off).reshape(nn, 3), rn)
    off += 24 * nn
    assert np.array_equal(np.frombuffer(raw, np.float32, nn, off), rz)
    off += 4 * nn
    assert np.array_equal(np.frombuffer(raw, np.int32, 2 * nn, off).reshape(nn, 2), rpix)
    off += 8 * nn
    assert off == len(raw)


# ---- drop-in matcher classes (shim/ORBmatcher.h, shim/LSDmatcher.h, shim/FrustumGPU.h) ----------------------------------------
EXE3 = os.path.join(ROOT, 'tests', 'cpp', 'shim_match')


def _build_match():
    """tests/cpp/shim_match = oracle/ref_match_main.cpp (the executed reference's driver and stand-in Frame / MapPoint / MapLine types)
    compiled with REF_MATCH_USE_SHIM: ORBmatcher / LSDmatcher are the drop-in class templates, nothing of the reference is compiled."""
    src = os.path.join(ROOT, 'tests', 'cpp', 'shim_match_main.cpp')
    if not os.path.exists(os.path.join(PKG, 'libhvofront.so')):
        import __graft_entry__
        __graft_entry__.build()
    subprocess.check_call(['g++', '-O2', '-std=c++14', '-I' + os.path.join(ROOT, 'oracle', 'cvshim'), '-I' + os.path.join(ROOT, 'oracle', 'eigenshim'),
                           '-I' + os.path.join(ROOT, 'tests', 'cpp', 'standins'), '-I' + os.path.join(PKG, 'shim'), '-I' + os.path.join(ROOT, 'include'),
                           src, '-o', EXE3, '-L' + PKG, '-lhvofront', '-Wl,-rpath,' + PKG])
    return EXE3


def test_dropin_matcher_classes_compile_against_reference_types():
    assert os.path.exists(_build_match())


@pytest.mark.gpu
def test_dropin_matcher_classes_equal_executed_reference(hvo, synth):
    """The same in.bin the reference's own binary was fed (scenes of tests/test_ref_match.py / tests/test_track.py) through the drop-in
    classes: all twelve public methods of ORBmatcher (SearchByProjection x4, SearchByBoW x2, SearchForInitialization, SearchForTriangulation,
    SearchBySim3, Fuse x2), LSDmatcher::SearchByProjection x2, FrameBFMatch, FrameBFMatchNew, match, SearchDouble x2, SearchByDescriptor and
    Frame::isInFrustum x2 (batched) must write what the reference wrote (committed fixtures match_ref.npz / track_ref.npz / kf_ref.npz)."""
    import oracle
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import test_ref_match as trm
    import test_track as tt
    assert trm._golden is not None and tt._golden is not None
    oracle.MATCH_EXE[0] = _build_match()
    try:
        for check in (trm._check_search_by_projection, trm._check_search_last, trm._check_line_search, trm._check_line_search_last):
            check(hvo, synth, gpu=True)
        for seed, cam, pts, ml, limit in tt._frustum_cases():
            (rp,) = tt._ref(f'fpt{seed}', lambda: oracle.ref_frustum_points(cam, pts, limit))
            (rl,) = tt._ref(f'fln{seed}', lambda: oracle.ref_frustum_lines(cam, ml, limit))
        for seed, window, ratio, ori in ((0, 100, 0.9, True), (1, 30, 0.9, True), (2, 100, 0.7, False)):
            F1, F2, prev = tt._init_scene(synth, seed)
            tt._ref(f'init{seed}', lambda: oracle.ref_search_initialization(F1, F2, prev, window, ratio, ori))
        for seed, TH, ratio in ((0, 50.0, 0.95), (1, 80.0, 0.8)):
            kl1, ld1, kl2, ld2, lv2, F = tt._line_scene(synth, seed)
            tt._ref(f'epi{seed}', lambda: oracle.ref_lines_epipolar(ld1, kl1, ld2, kl2, lv2, F, TH, ratio))
        from test_projection import _bow_scenario
        for seed, ratio, ori in ((0, 0.7, True), (1, 0.9, True), (2, 0.75, False)):
            KF, F = _bow_scenario(synth, seed)
            tt._ref(f'bow{seed}', lambda: oracle.ref_search_by_bow(KF, F, ratio, ori))
        # the key-frame side (tests/test_ref_kf.py, fixture kf_ref.npz): SearchByProjection(pKF, Scw, ...) / (Cur, pKF, found, th, ORBdist),
        # Fuse(pKF, Scw, ...), SearchBySim3, SearchByBoW(pKF1, pKF2), LSDmatcher::FrameBFMatch / match / SearchDouble x2 / SearchByDescriptor,
        # and SearchForTriangulation / Fuse(pKF, vpMapPoints, th) (fixture track_ref.npz): every public method of the two matcher classes
        import test_ref_kf as tk
        assert tk._golden is not None
        for check in tk.CHECKS:
            check(hvo, synth, gpu=True)
        tt._check_triangulation(hvo, synth, gpu=True)
        tt._check_fuse(hvo, synth, gpu=True)
    finally:
        oracle.MATCH_EXE[0] = None
