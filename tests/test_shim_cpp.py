"""The reference-facing C++ shim (shim/ORBextractor.h): compiles against the reference's call pattern and, on a
GPU, produces the same bytes as the oracle / the reference's own ORBextractor.cc through the same driver."""
import os
import struct
import subprocess
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
