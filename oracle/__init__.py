"""TEST INFRASTRUCTURE — ctypes access to the CPU oracle (oracle/_build/liboracle.so) and to the
reference's own ORBextractor.cc built against the OpenCV stand-in (oracle/_ref/ref_orb).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  Nothing here is on the product path."""
import ctypes as C
import os
import struct
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([('x', '<f4'), ('y', '<f4'), ('size', '<f4'), ('angle', '<f4'), ('response', '<f4'),
                     ('octave', '<i4'), ('class_id', '<i4')])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    so = os.path.join(_HERE, '_build', 'liboracle.so')
    if force or not os.path.exists(so):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_orb_create.restype = C.c_void_p
        _LIB.orc_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _LIB.orc_orb_destroy.argtypes = [C.c_void_p]
        _LIB.orc_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _LIB.orc_orb_tables.argtypes = [C.c_void_p] * 5
        _LIB.orc_orb_level_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _LIB.orc_orb_level_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _LIB.orc_orb_level_cand.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _LIB.orc_orb_level_kps.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ref_orb_path():
    p = os.path.join(_HERE, '_ref', 'ref_orb')
    return p if os.path.exists(p) else None


# ---- primitives -------------------------------------------------------------------------------------
def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], _p(dst), dw, dh)
    return dst


def blur7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_blur7(_p(src), src.shape[1], src.shape[0], _p(dst))
    return dst


def blur5(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_blur5(_p(src), src.shape[1], src.shape[0], _p(dst))
    return dst


def sobel3(src):
    src = np.ascontiguousarray(src, np.uint8)
    dx = np.empty(src.shape, np.int16)
    dy = np.empty(src.shape, np.int16)
    lib().orc_sobel3(_p(src), src.shape[1], src.shape[0], _p(dx), _p(dy))
    return dx, dy


def fast9(img, thr):
    """FAST-9/16 + NMS on a (possibly strided) 2-D uint8 view; returns int32 [n,3] = x, y, score."""
    assert img.dtype == np.uint8 and img.strides[1] == 1
    h, w = img.shape
    cap = max(16, w * h // 4)
    out = np.empty((cap, 3), np.int32)
    n = lib().orc_fast9(C.c_void_p(img.ctypes.data), w, h, img.strides[0], thr, _p(out), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    lib().orc_fast_atan2(_p(y), _p(x), _p(out), y.size)
    return out


# ---- ORB --------------------------------------------------------------------------------------------
class OrbOracle:
    """CPU restatement of ORB_SLAM2::ORBextractor (reference src/ORBextractor.cc)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.params = (nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nlevels = nlevels
        self.cap = nfeatures + 4 * nlevels + 64
        self._h = lib().orc_orb_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, '_h', None):
            lib().orc_orb_destroy(self._h)
            self._h = None

    def tables(self):
        sf = np.empty(self.nlevels, np.float32)
        isf = np.empty(self.nlevels, np.float32)
        nf = np.empty(self.nlevels, np.int32)
        um = np.empty(16, np.int32)
        lib().orc_orb_tables(self._h, _p(sf), _p(isf), _p(nf), _p(um))
        return sf, isf, nf, um

    def extract(self, gray):
        assert gray.dtype == np.uint8 and gray.ndim == 2 and gray.strides[1] == 1
        kps = np.empty(self.cap, KP_DTYPE)
        desc = np.empty((self.cap, 32), np.uint8)
        n = lib().orc_orb_extract(self._h, C.c_void_p(gray.ctypes.data), gray.shape[1], gray.shape[0],
                                  gray.strides[0], _p(kps), _p(desc), self.cap)
        assert n <= self.cap
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        """Intermediates of the last extract(): dict(img, blurred|None, cand [n,3], kps)."""
        info = np.empty(4, np.int32)
        lib().orc_orb_level_info(self._h, l, _p(info))
        w, h, nc, nk = (int(v) for v in info)
        img = np.empty((h, w), np.uint8)
        lib().orc_orb_level_image(self._h, l, _p(img), 0)
        blurred = None
        if nk:
            blurred = np.empty((h, w), np.uint8)
            lib().orc_orb_level_image(self._h, l, _p(blurred), 1)
        cand = np.empty((nc, 3), np.float32)
        if nc:
            lib().orc_orb_level_cand(self._h, l, _p(cand))
        kps = np.empty(nk, KP_DTYPE)
        if nk:
            lib().orc_orb_level_kps(self._h, l, _p(kps))
        return dict(img=img, blurred=blurred, cand=cand, kps=kps)


def distribute(cand, minX, maxX, minY, maxY, N):
    cand = np.ascontiguousarray(cand, np.float32).reshape(-1, 3)
    out = np.empty((len(cand) + 8, 3), np.float32)
    n = lib().orc_orb_distribute(_p(cand), len(cand), minX, maxX, minY, maxY, N, _p(out))
    return out[:n].copy()


def orb_extract_batch(frames, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, nthreads=1,
                      want_output=False):
    """Frame-parallel oracle run (CPU baseline). frames: [n,h,w] uint8."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    counts = np.zeros(n, np.int32)
    cap = nfeatures + 4 * nlevels + 64
    kps = np.empty((n, cap), KP_DTYPE) if want_output else None
    desc = np.empty((n, cap, 32), np.uint8) if want_output else None
    lib().orc_orb_extract_batch(C.c_int(nfeatures), C.c_float(scale_factor), C.c_int(nlevels), C.c_int(ini_th),
                                C.c_int(min_th), _p(frames), C.c_int(n), C.c_int(w), C.c_int(h), C.c_int(nthreads),
                                _p(counts), _p(kps) if want_output else None, _p(desc) if want_output else None,
                                C.c_int(cap))
    if want_output:
        return counts, kps, desc
    return counts


def ref_orb_extract(frames, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """Run the reference's own ORBextractor.cc (oracle/_ref/ref_orb) on [n,h,w] uint8 frames.
    Returns a list of (kps, desc) or None when the binary is not available."""
    exe = ref_orb_path()
    if exe is None:
        return None
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, 'in.bin'), os.path.join(td, 'out.bin')
        with open(fi, 'wb') as f:
            f.write(struct.pack('<8i', 0x4f524231, w, h, n, nfeatures, nlevels, ini_th, min_th))
            f.write(struct.pack('<f', scale_factor))
            f.write(frames.tobytes())
        subprocess.check_call([exe, fi, fo])
        raw = open(fo, 'rb').read()
    out, off = [], 0
    for _ in range(n):
        (k,) = struct.unpack_from('<i', raw, off)
        off += 4
        kps = np.frombuffer(raw, KP_DTYPE, k, off).copy()
        off += 28 * k
        desc = np.frombuffer(raw, np.uint8, 32 * k, off).reshape(k, 32).copy()
        off += 32 * k
        out.append((kps, desc))
    return out


# ---- matching ------------------------------------------------------------------------------------------
def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return int(lib().orc_hamming(_p(a), _p(b)))


def knn2(q, t):
    """BFMatcher(NORM_HAMMING).knnMatch(k=2): (idx [nq,2], dist [nq,2]) with ties -> lower train index."""
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    idx = np.empty((len(q), 2), np.int32); dist = np.empty((len(q), 2), np.int32)
    lib().orc_knn2(_p(q), len(q), _p(t), len(t), _p(idx), _p(dist))
    return idx, dist


def match_nnr(q, t, nnr):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    m = np.empty(len(q), np.int32)
    n = lib().orc_match_nnr(_p(q), len(q), _p(t), len(t), C.c_float(nnr), _p(m))
    return int(n), m


def frame_bf_match(q, t, nnratio, TH):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    m = np.empty(len(q), np.int32)
    lib().orc_frame_bf_match(_p(q), len(q), _p(t), len(t), C.c_float(nnratio), C.c_float(TH), _p(m))
    return m


PROJ_QUERY_DTYPE = np.dtype([('u', '<f4'), ('v', '<f4'), ('r', '<f4'), ('min_level', '<i4'), ('max_level', '<i4'), ('ur', '<f4'),
                             ('claims', '<i4'), ('reserved', '<i4')])


def grid_build(keys, bounds):
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    cnt = np.zeros(64 * 48, np.int32); items = np.zeros(max(len(keys), 1), np.int32)
    f = lib().orc_grid_build
    f.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    f(_p(keys), len(keys), *[float(b) for b in bounds], _p(cnt), _p(items))
    return cnt, items[:cnt.sum()].copy()


def window_origin(origin=None):
    """KeyFrame::GetFeaturesInArea's integer origin (src/KeyFrame.cc:627-666, include/KeyFrame.h:249-252) for the window lookups of the
    calls that follow on this thread; None restores the Frame rule (origin = the float bounds)."""
    f = lib().orc_window_origin
    f.argtypes = [C.c_int, C.c_float, C.c_float]
    f.restype = None
    if origin is None:
        f(0, 0.0, 0.0)
    else:
        f(1, float(origin[0]), float(origin[1]))


def features_in_area(keys, bounds, x, y, r, min_level=-1, max_level=-1):
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    out = np.zeros(max(len(keys), 1), np.int32)
    f = lib().orc_features_in_area
    f.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 7 + [C.c_int, C.c_int, C.c_void_p, C.c_int]
    n = f(_p(keys), len(keys), *[float(b) for b in bounds], float(x), float(y), float(r), int(min_level), int(max_level), _p(out), len(out))
    return out[:n].copy()


def search_projection(keys, uright, desc, bounds, queries, qdesc, claimed=None, mode=0, th_dist=100, nnratio=0.6):
    """ORBmatcher::SearchByProjection's sequential greedy loop (src/ORBmatcher.cc:45-132 mode 0, :1353-1497 mode 1)."""
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
    qd = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
    ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
    cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
    idx = np.full(max(len(q), 1), -1, np.int32); dist = np.full(max(len(q), 1), 256, np.int32)
    f = lib().orc_search_projection
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                                                                    C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    n = f(_p(keys), _p(ur) if ur is not None else None, _p(desc), len(keys), *[float(b) for b in bounds], _p(q), _p(qd), len(q),
          _p(cl) if cl is not None else None, int(mode), int(th_dist), float(nnratio), _p(idx), _p(dist))
    return idx[:len(q)], dist[:len(q)], int(n)


def search_fuse(keys, uright, desc, bounds, queries, qdesc, inv_sigma2, th_dist=50):
    """the candidate loop of ORBmatcher::Fuse(pKF, vpMapPoints, th) (src/ORBmatcher.cc:896-990)."""
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
    qd = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
    ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
    sg = np.ascontiguousarray(inv_sigma2, np.float32)
    idx = np.full(max(len(q), 1), -1, np.int32); dist = np.full(max(len(q), 1), 256, np.int32)
    f = lib().orc_search_fuse
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                                                                    C.c_void_p, C.c_void_p]
    n = f(_p(keys), _p(ur) if ur is not None else None, _p(desc), len(keys), *[float(b) for b in bounds], _p(q), _p(qd), len(q), _p(sg),
          int(th_dist), _p(idx), _p(dist))
    return idx[:len(q)], dist[:len(q)], int(n)


def search_triangulation(qdesc, qkeys, qstereo, tdesc, tkeys, tflags, offsets, cand, F12, ex, ey, scale_factors, level_sigma2,
                         only_stereo=False, th_low=50):
    """the candidate loop of ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:668-836, CheckDistEpipolarLine :143-160)."""
    qd = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32); qk = np.ascontiguousarray(qkeys, KP_DTYPE)
    qs = np.ascontiguousarray(qstereo, np.uint8)
    td = np.ascontiguousarray(tdesc, np.uint8).reshape(-1, 32); tk = np.ascontiguousarray(tkeys, KP_DTYPE)
    tf = np.ascontiguousarray(tflags, np.uint8)
    off = np.ascontiguousarray(offsets, np.int32); cd = np.ascontiguousarray(cand, np.int32)
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    sf = np.ascontiguousarray(scale_factors, np.float32); sg = np.ascontiguousarray(level_sigma2, np.float32)
    idx = np.full(max(len(qd), 1), -1, np.int32); dist = np.full(max(len(qd), 1), 256, np.int32)
    f = lib().orc_search_triangulation
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                  C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n = f(_p(qd), _p(qk), _p(qs), len(qd), _p(td), _p(tk), _p(tf), _p(off), _p(cd), _p(F), float(np.float32(ex)), float(np.float32(ey)), _p(sf),
          _p(sg), int(bool(only_stereo)), int(th_low), _p(idx), _p(dist))
    return idx[:len(qd)], dist[:len(qd)], int(n)


def match_candidates(q, t, offsets, cand):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    off = np.ascontiguousarray(offsets, np.int32); cd = np.ascontiguousarray(cand, np.int32)
    best4 = np.zeros((max(len(q), 1), 4), np.int32)
    f = lib().orc_match_candidates
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(_p(q), len(q), _p(t), _p(off), _p(cd), _p(best4))
    return best4[:len(q)]


def distinctive(desc, offsets):
    """MapPoint / MapLine::ComputeDistinctiveDescriptors (src/MapPoint.cc:240-300, src/MapLine.cpp:331-400) per group."""
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    off = np.ascontiguousarray(offsets, np.int32)
    n = len(off) - 1
    bi = np.full(max(n, 1), -1, np.int32); bm = np.full(max(n, 1), -1, np.int32)
    lib().orc_distinctive(_p(d), _p(off), C.c_int(n), _p(bi), _p(bm))
    return bi[:n], bm[:n]


def search_candidates(q, t, offsets, cand, th_dist=50, nnratio=0.7):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    off = np.ascontiguousarray(offsets, np.int32); cd = np.ascontiguousarray(cand, np.int32)
    idx = np.full(max(len(q), 1), -1, np.int32); dist = np.full(max(len(q), 1), 256, np.int32)
    f = lib().orc_search_candidates
    f.restype = C.c_int
    nm = f(_p(q), C.c_int(len(q)), _p(t), C.c_int(len(t)), _p(off), _p(cd), C.c_int(th_dist), C.c_float(nnratio), _p(idx), _p(dist))
    return idx[:len(q)], dist[:len(q)], nm


# ---- lines: LSD wrapper fields + LBD ---------------------------------------------------------------------
KL_DTYPE = np.dtype([('angle', '<f4'), ('class_id', '<i4'), ('octave', '<i4'), ('pt_x', '<f4'), ('pt_y', '<f4'),
                     ('response', '<f4'), ('size', '<f4'), ('startPointX', '<f4'), ('startPointY', '<f4'),
                     ('endPointX', '<f4'), ('endPointY', '<f4'), ('sPointInOctaveX', '<f4'), ('sPointInOctaveY', '<f4'),
                     ('ePointInOctaveX', '<f4'), ('ePointInOctaveY', '<f4'), ('lineLength', '<f4'), ('numOfPixels', '<i4')])
assert KL_DTYPE.itemsize == 68


def keylines_from_segments(seg, w, h):
    """LSDDetector::detectImpl keyline fields (single octave) from [n,4] float32 segments x1,y1,x2,y2."""
    seg = np.ascontiguousarray(seg, np.float32).reshape(-1, 4)
    out = np.empty(len(seg), KL_DTYPE)
    lib().orc_keylines_from_segments(_p(seg), len(seg), w, h, _p(out))
    return out


def lsd_detect(gray, want_scaled=False):
    """cv::createLineSegmentDetector().detect(gray) restated (oracle/lsd_oracle.cpp): [n,4] float32 x1,y1,x2,y2."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    cap = 16384
    seg = np.empty((cap, 4), np.float32)
    sw, sh = C.c_int(0), C.c_int(0)
    sc = np.empty(h * w, np.uint8)
    n = lib().orc_lsd_detect(_p(gray), w, h, _p(seg), cap, _p(sc), C.byref(sw), C.byref(sh))
    assert n <= cap
    if want_scaled:
        return seg[:n].copy(), sc[:sw.value * sh.value].reshape(sh.value, sw.value).copy()
    return seg[:n].copy()


def sort_by_response(resp):
    """Permutation std::sort(.., sort_lines_by_response()) (include/auxiliar.h:47-52) leaves: descending, ties in libstdc++'s order."""
    r = np.ascontiguousarray(resp, np.float32)
    idx = np.empty(len(r), np.int32)
    lib().orc_sort_by_response(_p(r), len(r), _p(idx))
    return idx


def line_extract(gray, n_features=200):
    """LINEextractor::operator() restated: LSD -> keylines -> (if > n_features) response sort, truncate, renumber ->
    LBD -> line functions.  Returns (keylines, desc, linevec [n,3] float64)."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    kl = keylines_from_segments(lsd_detect(gray), w, h)
    if len(kl) > n_features:  # std::sort with sort_lines_by_response: ties in libstdc++'s (unstable) order
        order = sort_by_response(kl['response'])[:n_features]
        kl = kl[order].copy()
        kl['class_id'] = np.arange(n_features, dtype=np.int32)
    desc = lbd_compute(gray, kl) if len(kl) else np.empty((0, 32), np.uint8)
    sx, sy = kl['startPointX'].astype(np.float64), kl['startPointY'].astype(np.float64)
    ex, ey = kl['endPointX'].astype(np.float64), kl['endPointY'].astype(np.float64)
    l0, l1, l2 = sy - ey, ex - sx, sx * ey - sy * ex
    nn = np.sqrt(l0 * l0 + l1 * l1)
    with np.errstate(invalid='ignore', divide='ignore'):
        lv = np.stack([l0 / nn, l1 / nn, l2 / nn], axis=1)
    return kl, desc, lv


def line_functions(kl):
    """mvKeyLineFunctions: sp x ep normalised by the norm of its first two components (LineExtractor.cpp:365-377)."""
    sx, sy = kl['startPointX'].astype(np.float64), kl['startPointY'].astype(np.float64)
    ex, ey = kl['endPointX'].astype(np.float64), kl['endPointY'].astype(np.float64)
    l0, l1, l2 = sy - ey, ex - sx, sx * ey - sy * ex
    nn = np.sqrt(l0 * l0 + l1 * l1)
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.stack([l0 / nn, l1 / nn, l2 / nn], axis=1)


def clip_line(w, h, pt1, pt2):
    """cv::clipLine restatement: (inside, pt1, pt2)."""
    a = np.array([pt1[0], pt1[1], pt2[0], pt2[1]], np.int64)
    r = lib().orc_clip_line(int(w), int(h), _p(a))
    return bool(r), (int(a[0]), int(a[1])), (int(a[2]), int(a[3]))


def cull_lines(keylines, linevec, w, h, dis=5.0, angle=2.5, endpoint_dis=15.0, want_groups=False):
    """Frame::cullingLine steps 1-3 (src/Frame.cc:952-1092): merged, rebuilt, response-sorted KeyLines."""
    kl = np.ascontiguousarray(keylines, KL_DTYPE)
    lv = np.ascontiguousarray(linevec, np.float64).reshape(-1, 3)
    out = np.zeros(max(len(kl), 1), KL_DTYPE)
    grp = np.full(max(len(kl), 1), -1, np.int32)
    f = lib().orc_cull_lines
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    m = f(_p(kl), _p(lv), len(kl), int(w), int(h), dis, angle, endpoint_dis, _p(out), _p(grp))
    return (out[:m].copy(), grp[:len(kl)].copy()) if want_groups else out[:m].copy()


def line_extract_culled(gray, n_features=200):
    """Frame::ExtractLSD up to cullingLine (src/Frame.cc:895-947): LINEextractor::operator() then cullingLine(im, 5, 2.5, 15, 30)
    with its second LBD pass and rebuilt line functions.  Returns (keylines, desc, linevec)."""
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    kl, _, lv = line_extract(gray, n_features)
    kl2 = cull_lines(kl, lv, w, h)
    desc2 = lbd_compute(gray, kl2) if len(kl2) else np.empty((0, 32), np.uint8)
    return kl2, desc2, line_functions(kl2)


def lbd_gradients(gray):
    gray = np.ascontiguousarray(gray, np.uint8)
    dx = np.empty(gray.shape, np.int16); dy = np.empty(gray.shape, np.int16)
    lib().orc_lbd_gradients(_p(gray), gray.shape[1], gray.shape[0], _p(dx), _p(dy))
    return dx, dy


def lbd_compute(gray, keylines, want_float=False):
    gray = np.ascontiguousarray(gray, np.uint8)
    kl = np.ascontiguousarray(keylines, KL_DTYPE)
    desc = np.empty((len(kl), 32), np.uint8)
    fdesc = np.empty((len(kl), 72), np.float32) if want_float else None
    lib().orc_lbd_compute(_p(gray), gray.shape[1], gray.shape[0], _p(kl), len(kl), _p(desc), _p(fdesc) if want_float else None)
    return (desc, fdesc) if want_float else desc


# ---- planes (PEAC) ----------------------------------------------------------------------------------------
def plane_blocks(depth16, factor, fx, fy, cx, cy):
    d = np.ascontiguousarray(depth16, np.uint16)
    h, w = d.shape
    out = np.empty(((h // 10) * (w // 10), 9), np.float64)
    lib().orc_plane_blocks(_p(d), w, h, C.c_float(factor), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(out))
    return out


def plane_detect(depth16, factor, fx, fy, cx, cy, max_planes=64):
    """(n_planes, planes [n,7] = normal, center, N, membership [h*w])"""
    d = np.ascontiguousarray(depth16, np.uint16)
    h, w = d.shape
    planes = np.zeros((max_planes, 7), np.float64)
    mem = np.empty(h * w, np.int32)
    n = lib().orc_plane_detect(_p(d), w, h, C.c_float(factor), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                               _p(planes), max_planes, _p(mem))
    return int(n), planes[:n].copy(), mem


def eig33_smallest(K):
    K = np.ascontiguousarray(K, np.float64)
    lam = np.zeros(1, np.float64); v = np.empty(3, np.float64)
    lib().orc_eig33_smallest(_p(K), _p(lam), _p(v))
    return float(lam[0]), v


def eig33sym(K):
    K = np.ascontiguousarray(K, np.float64)
    s = np.empty(3, np.float64); V = np.empty((3, 3), np.float64)
    lib().orc_eig33sym(_p(K), _p(s), _p(V))
    return s, V


def ref_bin(name):
    """Path of a binary built from the reference's own sources (oracle/_ref/<name>), or None."""
    p = os.path.join(_HERE, '_ref', name)
    return p if os.path.exists(p) else None


def ref_peac(depth_frames, factor, fx, fy, cx, cy, perturb=None):
    """Run the reference's own PlaneExtractor.cpp + include/peac/*.hpp (oracle/_ref/ref_peac) on [n,h,w] uint16 frames.
    Returns a list of dicts: blocks [nb,10] (N, nouse, center, normal, mse, curvature), planes [np,9] (N, nvertices, normal,
    center, mse), membership [h*w] int32.  None when the binary is not available.  perturb: eigen-solver perturbation
    (ref_peac_perturb build)."""
    exe = ref_bin('ref_peac_perturb' if perturb is not None else 'ref_peac')
    if exe is None:
        return None
    d = np.ascontiguousarray(depth_frames, np.uint16)
    n, h, w = d.shape
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, 'in.bin'), os.path.join(td, 'out.bin')
        with open(fi, 'wb') as f:
            f.write(struct.pack('<4i', 0x50454143, w, h, n))
            f.write(np.array([fx, fy, cx, cy, factor], np.float32).tobytes())
            f.write(d.tobytes())
        subprocess.check_call([exe, fi, fo] + ([repr(float(perturb))] if perturb is not None else []))
        raw = open(fo, 'rb').read()
    out, off = [], 0
    bdt = np.dtype([('N', '<i4'), ('nouse', '<i4'), ('center', '<f8', (3,)), ('normal', '<f8', (3,)), ('mse', '<f8'), ('curvature', '<f8')])
    pdt = np.dtype([('N', '<i4'), ('nvertices', '<i4'), ('normal', '<f8', (3,)), ('center', '<f8', (3,)), ('mse', '<f8')])
    for _ in range(n):
        (nb,) = struct.unpack_from('<i', raw, off); off += 4
        blocks = np.frombuffer(raw, bdt, nb, off).copy(); off += bdt.itemsize * nb
        (npl,) = struct.unpack_from('<i', raw, off); off += 4
        planes = np.frombuffer(raw, pdt, npl, off).copy(); off += pdt.itemsize * npl
        mem = np.frombuffer(raw, np.int32, h * w, off).copy(); off += 4 * h * w
        out.append(dict(blocks=blocks, planes=planes, membership=mem))
    assert off == len(raw)
    return out


def ref_lines(frames, n_features=200, cull=True):
    """Run the reference's own line front-end (oracle/_ref/ref_lines: LSDDetector_custom.cpp whole, the LBD functions of
    binary_descriptor_custom.cpp, LINEextractor::operator() and Frame::cullingLine extracted at build time) on [n,h,w] uint8
    frames.  Returns a list of dicts: keylines/desc/fdesc/linevec (after LINEextractor::operator()) and, with cull,
    keylines2/desc2/linevec2 (after Frame::cullingLine).  None when the binary is not available."""
    exe = ref_bin('ref_lines')
    if exe is None:
        return None
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, 'in.bin'), os.path.join(td, 'out.bin')
        with open(fi, 'wb') as f:
            f.write(struct.pack('<6i', 0x4c494e45, w, h, n, n_features, 1 if cull else 0))
            f.write(frames.tobytes())
        subprocess.check_call([exe, fi, fo], stdout=subprocess.DEVNULL)
        raw = open(fo, 'rb').read()
    out, off = [], 0

    def take(with_float):
        nonlocal off
        (k,) = struct.unpack_from('<i', raw, off); off += 4
        kl = np.frombuffer(raw, KL_DTYPE, k, off).copy(); off += 68 * k
        d = np.frombuffer(raw, np.uint8, 32 * k, off).reshape(k, 32).copy(); off += 32 * k
        fd = None
        if with_float:
            fd = np.frombuffer(raw, np.float32, 72 * k, off).reshape(k, 72).copy(); off += 288 * k
        lv = np.frombuffer(raw, np.float64, 3 * k, off).reshape(k, 3).copy(); off += 24 * k
        return kl, d, fd, lv
    for _ in range(n):
        kl, d, fd, lv = take(True)
        r = dict(keylines=kl, desc=d, fdesc=fd, linevec=lv)
        if cull:
            kl2, d2, _, lv2 = take(False)
            r.update(keylines2=kl2, desc2=d2, linevec2=lv2)
        out.append(r)
    assert off == len(raw)
    return out


# ---- the reference's own windowed matchers, executed (oracle/_ref/ref_match; oracle/ref_match_main.cpp) ----------------
# tests/test_shim_cpp.py runs the SAME payloads through tests/cpp/shim_match (the drop-in matcher classes behind the reference
# harness' driver): MATCH_EXE[0] names that binary; its output has no grid / candidate-list sections.
MATCH_EXE = [None]


def _run_ref_match(payload):
    exe = MATCH_EXE[0] or ref_bin('ref_match')
    if exe is None:
        return None
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, 'in.bin'), os.path.join(td, 'out.bin')
        with open(fi, 'wb') as f:
            f.write(payload)
        subprocess.check_call([exe, fi, fo], stdout=subprocess.DEVNULL)
        return open(fo, 'rb').read()


def _f32(a):
    return np.ascontiguousarray(a, np.float32).tobytes()


def _point_frame_bytes(F):
    k = np.ascontiguousarray(F['keys_un'], KP_DTYPE)
    n = len(k)
    ur = np.full(n, -1, np.float32) if F.get('uright') is None else F['uright']
    cl = np.zeros(n, np.uint8) if F.get('claimed') is None else np.asarray(F['claimed'], np.uint8)
    sf = np.zeros(8, np.float32); sf[:len(F['scale_factors'])] = F['scale_factors']
    return struct.pack('<i', n) + k.tobytes() + _f32(ur) + np.ascontiguousarray(F['desc'], np.uint8).tobytes() + cl.tobytes() + sf.tobytes()


def _line_frame_bytes(F):
    kl = np.ascontiguousarray(F['keylines_un'], KL_DTYPE)
    n = len(kl)
    l3 = np.zeros((n, 6)) if F.get('lines3d') is None else F['lines3d']
    cl = np.zeros(n, np.uint8) if F.get('claimed') is None else np.asarray(F['claimed'], np.uint8)
    return (struct.pack('<i', n) + kl.tobytes() + np.ascontiguousarray(F['line_functions'], np.float64).tobytes()
            + np.ascontiguousarray(F['ldesc'], np.uint8).tobytes() + np.ascontiguousarray(l3, np.float64).tobytes() + cl.tobytes())


def _read_grid(raw, off):
    cnt = np.empty(64 * 48, np.int32); items = []
    for c in range(64 * 48):
        (k,) = struct.unpack_from('<i', raw, off); off += 4
        cnt[c] = k
        items.append(np.frombuffer(raw, np.int32, k, off)); off += 4 * k
    return cnt, (np.concatenate(items) if items else np.zeros(0, np.int32)), off


def _read_lists(raw, off, n):
    out = []
    for _ in range(n):
        (k,) = struct.unpack_from('<i', raw, off); off += 4
        out.append(np.frombuffer(raw, np.int32, k, off).copy()); off += 4 * k
    return out, off


def ref_search_by_projection(F, MPs, th, nnratio, windows=()):
    """The reference's ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) executed on the mirror's dict layout
    (hvo.ORBmatcher docstring).  windows: [(x, y, r, minLevel, maxLevel)] answered by its Frame::GetFeaturesInArea.
    Returns dict(nmatches, assign [N] map point index / -1 / -2 = held before the call, grid (cnt, items), areas)."""
    M = len(MPs['proj_x'])
    b = struct.pack('<2i', 0x4d544348, 0) + _f32(F['bounds']) + _point_frame_bytes(F) + struct.pack('<iff', M, th, nnratio)
    for i in range(M):
        b += struct.pack('<3fif4B', MPs['proj_x'][i], MPs['proj_y'][i], MPs['proj_xr'][i], int(MPs['level'][i]), MPs['view_cos'][i],
                         int(MPs['in_view'][i]), int(MPs['bad'][i]), int(MPs['has_obs'][i]), 0) + np.asarray(MPs['desc'][i], np.uint8).tobytes()
    w = np.asarray(windows, np.float32).reshape(-1, 5)
    b += struct.pack('<i', len(w)) + w.tobytes()
    raw = _run_ref_match(b)
    if raw is None:
        return None
    if MATCH_EXE[0]:
        (nm,) = struct.unpack_from('<i', raw, 0)
        return dict(nmatches=nm, assign=np.frombuffer(raw, np.int32, len(F['keys_un']), 4).copy())
    cnt, items, off = _read_grid(raw, 0)
    areas, off = _read_lists(raw, off, len(w))
    (nm,) = struct.unpack_from('<i', raw, off); off += 4
    n = len(F['keys_un'])
    return dict(nmatches=nm, assign=np.frombuffer(raw, np.int32, n, off).copy(), grid=(cnt, items), areas=areas)


def ref_search_by_projection_last(Cur, Last, cam, Tcw_cur, Tcw_last, th, mono=False, check_ori=True):
    """The reference's ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) executed.  cam = (fx, fy, cx, cy, mb, mbf);
    Last = dict(keys KP_DTYPE [N], has_mp, outlier, has_obs, world_pos [N,3] float32, desc [N,32]).
    Returns dict(nmatches, assign [Ncur] last-frame feature index / -1 / -2)."""
    n = len(Last['keys'])
    b = struct.pack('<2i', 0x4d544348, 1) + _f32(Cur['bounds']) + _point_frame_bytes(Cur) + _f32(cam) + _f32(Tcw_cur) + _f32(Tcw_last)
    b += struct.pack('<f3i', th, int(mono), int(check_ori), n) + np.ascontiguousarray(Last['keys'], KP_DTYPE).tobytes()
    for i in range(n):
        b += struct.pack('<4B', int(Last['has_mp'][i]), int(Last['outlier'][i]), int(Last['has_obs'][i]), 0) + _f32(Last['world_pos'][i]) \
            + np.asarray(Last['desc'][i], np.uint8).tobytes()
    raw = _run_ref_match(b)
    if raw is None:
        return None
    (nm,) = struct.unpack_from('<i', raw, 0)
    return dict(nmatches=nm, assign=np.frombuffer(raw, np.int32, len(Cur['keys_un']), 4).copy())


def ref_line_search_by_projection(F, MLs, th, nnratio, windows=()):
    """The reference's LSDmatcher::SearchByProjection(Frame&, vector<MapLine*>&, eval_orient, th) executed on the mirror's dict layout
    (hvo.LSDmatcher docstring; MLs also carries level).  windows: [(x1, y1, x2, y2, r, TH)] for Frame::GetFeaturesInAreaForLine."""
    M = len(MLs['proj_x1'])
    b = struct.pack('<2i', 0x4d544348, 2) + _f32(F['bounds']) + _line_frame_bytes(F) + struct.pack('<iff', M, th, nnratio)
    for i in range(M):
        b += struct.pack('<4fif4B', MLs['proj_x1'][i], MLs['proj_y1'][i], MLs['proj_x2'][i], MLs['proj_y2'][i], int(MLs['level'][i]), MLs['view_cos'][i],
                         int(MLs['in_view'][i]), int(MLs['bad'][i]), int(MLs['has_obs'][i]), 0) \
            + np.asarray(MLs['world_vector'][i], np.float64).tobytes() + np.asarray(MLs['desc'][i], np.uint8).tobytes()
    w = np.asarray(windows, np.float32).reshape(-1, 6)
    b += struct.pack('<i', len(w)) + w.tobytes()
    raw = _run_ref_match(b)
    if raw is None:
        return None
    if MATCH_EXE[0]:
        (nm,) = struct.unpack_from('<i', raw, 0)
        return dict(nmatches=nm, assign=np.frombuffer(raw, np.int32, len(F['keylines_un']), 4).copy())
    cnt, items, off = _read_grid(raw, 0)
    areas, off = _read_lists(raw, off, len(w))
    (nm,) = struct.unpack_from('<i', raw, off); off += 4
    return dict(nmatches=nm, assign=np.frombuffer(raw, np.int32, len(F['keylines_un']), off).copy(), grid=(cnt, items), areas=areas)


def ref_line_search_by_projection_last(Cur, Last, th):
    """The reference's LSDmatcher::SearchByProjection(CurrentFrame, LastFrame, th) executed.  Last = dict(keylines KL_DTYPE [N], has_ml,
    outlier, has_obs, in_frustum, proj_x1, proj_y1, proj_x2, proj_y2, level, desc [N,32]) (isInFrustum is the harness's one stand-in)."""
    n = len(Last['keylines'])
    eye = np.eye(4, dtype=np.float32)
    b = struct.pack('<2i', 0x4d544348, 3) + _f32(Cur['bounds']) + _line_frame_bytes(Cur) + _f32(eye) + _f32(eye) + struct.pack('<fi', th, n)
    b += np.ascontiguousarray(Last['keylines'], KL_DTYPE).tobytes()
    for i in range(n):
        b += struct.pack('<4B4fi', int(Last['has_ml'][i]), int(Last['outlier'][i]), int(Last['has_obs'][i]), int(Last['in_frustum'][i]),
                         Last['proj_x1'][i], Last['proj_y1'][i], Last['proj_x2'][i], Last['proj_y2'][i], int(Last['level'][i])) \
            + np.asarray(Last['desc'][i], np.uint8).tobytes()
    raw = _run_ref_match(b)
    if raw is None:
        return None
    (nm,) = struct.unpack_from('<i', raw, 0)
    return dict(nmatches=nm, assign=np.frombuffer(raw, np.int32, len(Cur['keylines_un']), 4).copy())


# ---- surface normals (PCL integral-image style, Frame.cc:2155-2212) -------------------------------------------

# ---- tracking-time projection and two matcher loops (oracle/track_oracle.cpp) ------------------------------------------------
FRUSTUM_CAM_DTYPE = np.dtype([('Rcw', '<f4', (9,)), ('tcw', '<f4', (3,)), ('Ow', '<f4', (3,)), ('fx', '<f4'), ('fy', '<f4'), ('cx', '<f4'),
                              ('cy', '<f4'), ('bf', '<f4'), ('min_x', '<f4'), ('min_y', '<f4'), ('max_x', '<f4'), ('max_y', '<f4'),
                              ('log_scale_factor', '<f4'), ('n_levels', '<i4')])
MAP_POINT_DTYPE = np.dtype([('pos', '<f4', (3,)), ('normal', '<f4', (3,)), ('min_distance', '<f4'), ('max_distance', '<f4')])
TRACK_POINT_DTYPE = np.dtype([('u', '<f4'), ('v', '<f4'), ('ur', '<f4'), ('level', '<i4'), ('view_cos', '<f4'), ('in_view', '<i4')])
MAP_LINE_DTYPE = np.dtype([('pos', '<f8', (6,)), ('normal', '<f8', (3,)), ('dir', '<f8', (3,)), ('min_distance', '<f4'), ('max_distance', '<f4')])
TRACK_LINE_DTYPE = np.dtype([('x1', '<f4'), ('y1', '<f4'), ('x2', '<f4'), ('y2', '<f4'), ('level', '<i4'), ('view_cos', '<f4'), ('in_view', '<i4')])


def frustum_points(cam, pts, limit=0.5):
    """Frame::isInFrustum(MapPoint*, limit) (src/Frame.cc:1371-1436)."""
    cam = np.ascontiguousarray(cam, FRUSTUM_CAM_DTYPE); pts = np.ascontiguousarray(pts, MAP_POINT_DTYPE)
    out = np.zeros(max(len(pts), 1), TRACK_POINT_DTYPE)
    f = lib().orc_frustum_points
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]; f.restype = None
    f(_p(cam), _p(pts), len(pts), float(limit), _p(out))
    return out[:len(pts)]


def frustum_lines(cam, lines, limit=0.5):
    """Frame::isInFrustum(MapLine*, limit) (src/Frame.cc:1438-1499)."""
    cam = np.ascontiguousarray(cam, FRUSTUM_CAM_DTYPE); ml = np.ascontiguousarray(lines, MAP_LINE_DTYPE)
    out = np.zeros(max(len(ml), 1), TRACK_LINE_DTYPE)
    f = lib().orc_frustum_lines
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]; f.restype = None
    f(_p(cam), _p(ml), len(ml), float(limit), _p(out))
    return out[:len(ml)]


def three_maxima(sizes):
    """ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1630-1671)."""
    max1 = max2 = max3 = 0
    ind1 = ind2 = ind3 = -1
    for i, s in enumerate(sizes):
        if s > max1:
            max3, max2, max1 = max2, max1, s
            ind3, ind2, ind1 = ind2, ind1, i
        elif s > max2:
            max3, max2 = max2, s
            ind3, ind2 = ind2, i
        elif s > max3:
            max3, ind3 = s, i
    if max2 < np.float32(0.1) * np.float32(max1):
        ind2 = ind3 = -1
    elif max3 < np.float32(0.1) * np.float32(max1):
        ind3 = -1
    return ind1, ind2, ind3


def search_initialization(F1, F2, prev_matched, window=100, nnratio=0.9, check_ori=True, th_low=50):
    """ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:412-529), whole: Fi = dict(keys_un, desc, bounds).
    Returns (nmatches, vnMatches12, vbPrevMatched, accepted12)."""
    k1 = np.ascontiguousarray(F1['keys_un'], KP_DTYPE); k2 = np.ascontiguousarray(F2['keys_un'], KP_DTYPE)
    d1 = np.ascontiguousarray(F1['desc'], np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(F2['desc'], np.uint8).reshape(-1, 32)
    prev = np.array(prev_matched, np.float32).reshape(-1, 2)
    oc = np.ascontiguousarray(k1['octave'], np.int32)
    b = np.ascontiguousarray(F2['bounds'], np.float32)
    m12 = np.full(max(len(k1), 1), -1, np.int32); acc = np.full(max(len(k1), 1), -1, np.int32)
    f = lib().orc_search_initialization
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                  C.c_void_p]
    nm = f(_p(k2), _p(d2), len(k2), _p(b), _p(prev), _p(oc), _p(d1), len(k1), int(window), int(th_low), float(nnratio), _p(m12), _p(acc))
    m12 = m12[:len(k1)]; acc = acc[:len(k1)]
    if check_ori:
        hist = [[] for _ in range(30)]
        factor = np.float32(1.0) / np.float32(30)
        for i1, i2 in enumerate(acc):
            if i2 < 0:
                continue
            rot = np.float32(k1['angle'][i1] - k2['angle'][i2])
            if rot < 0.0:
                rot = np.float32(rot + np.float32(360.0))
            bin_ = int(np.floor(float(np.float32(rot * factor)) + 0.5))
            if bin_ == 30:
                bin_ = 0
            hist[bin_].append(i1)
        a, b2, c = three_maxima([len(h) for h in hist])
        for i in range(30):
            if i not in (a, b2, c):
                for i1 in hist[i]:
                    if m12[i1] >= 0:
                        m12[i1] = -1
                        nm -= 1
    for i1 in np.nonzero(m12 >= 0)[0]:
        prev[i1, 0] = k2['x'][m12[i1]]; prev[i1, 1] = k2['y'][m12[i1]]
    return int(nm), m12, prev, acc


def lines_epipolar(ldesc1, kls1, ldesc2, kls2, kls2func, F, TH, nnratio):
    """LSDmatcher::FrameBFMatchNew (src/LSDmatcher.cpp:968-1031)."""
    d1 = np.ascontiguousarray(ldesc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(ldesc2, np.uint8).reshape(-1, 32)
    k1 = np.ascontiguousarray(kls1, KL_DTYPE); k2 = np.ascontiguousarray(kls2, KL_DTYPE)
    f2 = np.ascontiguousarray(kls2func, np.float64).reshape(-1, 3); Fm = np.ascontiguousarray(F, np.float32).reshape(9)
    out = np.full(max(len(d1), 1), -1, np.int32)
    f = lib().orc_lines_epipolar
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    f.restype = None
    f(_p(d1), _p(k1), len(d1), _p(d2), _p(k2), _p(f2), len(d2), _p(Fm), float(TH), float(nnratio), _p(out))
    return out[:len(d1)]


# the reference's own functions executed (oracle/_ref/ref_match ops 4-7)
def _frustum_payload(op, cam, limit, body, n):
    cam = np.ascontiguousarray(cam, FRUSTUM_CAM_DTYPE).reshape(())
    b = struct.pack('<2i', 0x4d544348, op) + _f32([cam['min_x'], cam['min_y'], cam['max_x'], cam['max_y']])
    b += _f32([cam['fx'], cam['fy'], cam['cx'], cam['cy'], cam['bf']]) + _f32(cam['Rcw']) + _f32(cam['tcw']) + _f32(cam['Ow'])
    b += struct.pack('<fifi', float(cam['log_scale_factor']), int(cam['n_levels']), float(limit), n)
    return b + body


def ref_frustum_points(cam, pts, limit=0.5):
    """The reference's Frame::isInFrustum(MapPoint*, float) executed over a batch.  None when oracle/_ref/ref_match is absent."""
    pts = np.ascontiguousarray(pts, MAP_POINT_DTYPE)
    raw = _run_ref_match(_frustum_payload(4, cam, limit, pts.tobytes(), len(pts)))
    return None if raw is None else np.frombuffer(raw, TRACK_POINT_DTYPE, len(pts)).copy()


def ref_frustum_lines(cam, lines, limit=0.5):
    """The reference's Frame::isInFrustum(MapLine*, float) executed over a batch."""
    ml = np.ascontiguousarray(lines, MAP_LINE_DTYPE)
    raw = _run_ref_match(_frustum_payload(5, cam, limit, ml.tobytes(), len(ml)))
    return None if raw is None else np.frombuffer(raw, TRACK_LINE_DTYPE, len(ml)).copy()


def ref_search_initialization(F1, F2, prev_matched, window=100, nnratio=0.9, check_ori=True):
    """The reference's ORBmatcher::SearchForInitialization executed.  Returns (nmatches, vnMatches12, vbPrevMatched) or None."""
    k1 = np.ascontiguousarray(F1['keys_un'], KP_DTYPE)
    prev = np.ascontiguousarray(prev_matched, np.float32).reshape(-1, 2)
    F2f = dict(F2); F2f.setdefault('scale_factors', np.ones(8, np.float32))
    b = struct.pack('<2i', 0x4d544348, 6) + _f32(F2['bounds']) + _point_frame_bytes(F2f)
    b += struct.pack('<i', len(k1)) + k1.tobytes() + np.ascontiguousarray(F1['desc'], np.uint8).tobytes() + prev.tobytes()
    b += struct.pack('<ifi', int(window), float(nnratio), int(check_ori))
    raw = _run_ref_match(b)
    if raw is None:
        return None
    n = len(k1)
    (nm,) = struct.unpack_from('<i', raw, 0)
    return nm, np.frombuffer(raw, np.int32, n, 4).copy(), np.frombuffer(raw, np.float32, 2 * n, 4 + 4 * n).reshape(n, 2).copy()


def ref_lines_epipolar(ldesc1, kls1, ldesc2, kls2, kls2func, F, TH, nnratio):
    """The reference's LSDmatcher::FrameBFMatchNew executed.  Returns LineMatches or None."""
    k1 = np.ascontiguousarray(kls1, KL_DTYPE); k2 = np.ascontiguousarray(kls2, KL_DTYPE)
    b = struct.pack('<2i', 0x4d544348, 7) + _f32([0, 0, 640, 480])
    b += struct.pack('<i', len(k1)) + k1.tobytes() + np.ascontiguousarray(ldesc1, np.uint8).tobytes()
    b += struct.pack('<i', len(k2)) + k2.tobytes() + np.ascontiguousarray(ldesc2, np.uint8).tobytes()
    b += np.ascontiguousarray(kls2func, np.float64).tobytes() + _f32(F) + struct.pack('<2f', float(TH), float(nnratio))
    raw = _run_ref_match(b)
    return None if raw is None else np.frombuffer(raw, np.int32, len(k1)).copy()



# ---- bag of words (oracle/bow_oracle.cpp; the reference's own DBoW2 executed: oracle/_ref/ref_bow) --------------------------------
def bow_transform(voc, desc, levelsup=4):
    """DBoW2 transform(features, BowVector, FeatureVector, levelsup) for one frame.  Returns ((words, values), fv dict, word_of, node_of)."""
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(d)
    cs = np.ascontiguousarray(voc['child_start'], np.int32); ci = np.ascontiguousarray(voc['child_ids'], np.int32)
    nd = np.ascontiguousarray(voc['node_desc'], np.uint8); nw = np.ascontiguousarray(voc['node_weight'], np.float64)
    wd = np.ascontiguousarray(voc['node_word'], np.int32)
    word = np.full(max(n, 1), -1, np.int32); node = np.zeros(max(n, 1), np.int32)
    bw = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1), np.float64); fo = np.zeros(max(n, 1), np.int32)
    k1, k2 = C.c_int(0), C.c_int(0)
    f = lib().orc_bow_transform
    f.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]
    f.restype = None
    f(_p(cs), _p(ci), _p(nd), _p(nw), _p(wd), int(voc['L']), _p(d), n, int(levelsup), _p(word), _p(node), _p(bw), _p(bv), C.byref(k1), _p(fo), C.byref(k2))
    fv = {}
    for i in fo[:k2.value]:
        fv.setdefault(int(node[i]), []).append(int(i))
    return (bw[:k1.value].copy(), bv[:k1.value].copy()), fv, word[:n].copy(), node[:n].copy()


def ref_bow(training, frames, k=10, L=4, seed=1, levelsup=2, exe_name='ref_bow'):
    """The reference's own DBoW2 executed (oracle/_ref/ref_bow): builds a vocabulary with TemplatedVocabulary::create from `training`
    (list of [n,32] descriptor arrays) and transforms `frames`.  Returns (voc dict, [((words, values), fv dict)]) or None.
    exe_name='shim_bow': the same driver and DBoW2 vocabulary object with the transform routed through shim/ORBVocabularyGPU.h (needs a GPU)."""
    exe = ref_bin(exe_name)
    if exe is None:
        return None
    b = struct.pack('<6i', 0x424f5756, k, L, seed, levelsup, len(training))
    for t in training:
        t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        b += struct.pack('<i', len(t)) + t.tobytes()
    b += struct.pack('<i', len(frames))
    for t in frames:
        t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        b += struct.pack('<i', len(t)) + t.tobytes()
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, 'in.bin'), os.path.join(td, 'out.bin')
        with open(fi, 'wb') as f:
            f.write(b)
        subprocess.check_call([exe, fi, fo], stdout=subprocess.DEVNULL)
        raw = open(fo, 'rb').read()
    off = 0
    (nn,) = struct.unpack_from('<i', raw, off); off += 4
    cs, ci, nd, nw, wd = [0], [], np.zeros((nn, 32), np.uint8), np.zeros(nn), np.zeros(nn, np.int32)
    for i in range(nn):
        parent, word, nc = struct.unpack_from('<3i', raw, off); off += 12
        ci.extend(struct.unpack_from(f'<{nc}i', raw, off)); off += 4 * nc
        cs.append(len(ci))
        nd[i] = np.frombuffer(raw, np.uint8, 32, off); off += 32
        (nw[i],) = struct.unpack_from('<d', raw, off); off += 8
        wd[i] = word
    voc = dict(child_start=np.asarray(cs, np.int32), child_ids=np.asarray(ci, np.int32), node_desc=nd, node_weight=nw, node_word=wd, L=L)
    out = []
    for _ in frames:
        (m,) = struct.unpack_from('<i', raw, off); off += 4
        rec = np.frombuffer(raw, np.dtype([('w', '<i4'), ('v', '<f8')]), m, off); off += 12 * m
        (nn2,) = struct.unpack_from('<i', raw, off); off += 4
        fv = {}
        for _ in range(nn2):
            nid, cnt = struct.unpack_from('<2i', raw, off); off += 8
            fv[nid] = list(struct.unpack_from(f'<{cnt}i', raw, off)); off += 4 * cnt
        out.append(((rec['w'].copy(), rec['v'].copy()), fv))
    assert off == len(raw)
    return voc, out



def ref_distinctive(desc, offsets, lines=False):
    """The reference's MapPoint / MapLine::ComputeDistinctiveDescriptors executed for a batch of map elements (oracle/_ref/ref_match op 8).
    Returns (best_idx [ngroups] (-1 for an element without observations), best descriptors [ngroups, 32]) or None."""
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); off = np.ascontiguousarray(offsets, np.int32)
    ng = len(off) - 1
    b = struct.pack('<2i', 0x4d544348, 8) + _f32([0, 0, 640, 480]) + struct.pack('<2i', int(lines), ng) + off.tobytes() + d.tobytes()
    raw = _run_ref_match(b)
    if raw is None:
        return None
    rec = np.frombuffer(raw, np.dtype([('idx', '<i4'), ('desc', 'u1', (32,))]), ng)
    return rec['idx'].copy(), rec['desc'].copy()


def _featvec_bytes(fv):
    b = struct.pack('<i', len(fv))
    for node in sorted(fv):
        idx = np.asarray(fv[node], np.int32)
        b += struct.pack('<2i', int(node), len(idx)) + idx.tobytes()
    return b


def ref_search_by_bow(KF, F, nnratio=0.7, check_ori=True, kf_bad=None):
    """The reference's ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) executed (op 9) on the mirror's dict layout
    (hvo.ORBmatcher.SearchByBoW).  Returns (nmatches, match [M]) or None."""
    k0 = np.ascontiguousarray(KF['keys_un'], KP_DTYPE); k1 = np.ascontiguousarray(F['keys'], KP_DTYPE)
    n0 = len(k0)
    bad = np.zeros(n0, np.uint8) if kf_bad is None else np.asarray(kf_bad, np.uint8)
    b = struct.pack('<2i', 0x4d544348, 9) + _f32([0, 0, 640, 480])
    b += struct.pack('<i', n0) + k0.tobytes() + np.ascontiguousarray(KF['desc'], np.uint8).tobytes() + np.asarray(KF['has_mappoint'], np.uint8).tobytes() + bad.tobytes()
    b += _featvec_bytes(KF['featvec'])
    b += struct.pack('<i', len(k1)) + k1.tobytes() + np.ascontiguousarray(F['desc'], np.uint8).tobytes() + _featvec_bytes(F['featvec'])
    b += struct.pack('<fi', float(nnratio), int(check_ori))
    raw = _run_ref_match(b)
    if raw is None:
        return None
    (nm,) = struct.unpack_from('<i', raw, 0)
    return nm, np.frombuffer(raw, np.int32, len(k1), 4).copy()



def ref_search_for_triangulation(KF1, KF2, F12, Cw1, R2w, t2w, cam2, only_stereo=False, check_ori=True, nnratio=0.6):
    """The reference's ORBmatcher::SearchForTriangulation executed (op 10) on the mirror's dict layout.  Cw1 = pKF1->GetCameraCenter(),
    R2w / t2w = pKF2's pose, cam2 = (fx, fy, cx, cy) of pKF2.  Returns (nmatches, pairs [n, 2]) or None."""
    def kf(K):
        k = np.ascontiguousarray(K['keys_un'], KP_DTYPE)
        return (struct.pack('<i', len(k)) + k.tobytes() + _f32(K['uright']) + np.ascontiguousarray(K['desc'], np.uint8).tobytes()
                + np.asarray(K['has_mappoint'], np.uint8).tobytes() + _featvec_bytes(K['featvec']))
    sf = np.zeros(8, np.float32); sf[:len(KF2['scale_factors'])] = KF2['scale_factors']
    sg = np.zeros(8, np.float32); sg[:len(KF2['level_sigma2'])] = KF2['level_sigma2']
    b = struct.pack('<2i', 0x4d544348, 10) + _f32([0, 0, 640, 480]) + kf(KF1) + _f32(Cw1) + kf(KF2) + _f32(R2w) + _f32(t2w) + _f32(cam2)
    b += sf.tobytes() + sg.tobytes() + _f32(F12) + struct.pack('<2if', int(only_stereo), int(check_ori), float(nnratio))
    raw = _run_ref_match(b)
    if raw is None:
        return None
    nm, npairs = struct.unpack_from('<2i', raw, 0)
    return nm, np.frombuffer(raw, np.int32, 2 * npairs, 8).reshape(npairs, 2).copy()



def ref_fuse(KF, kf_obs, cam5, Rcw, tcw, Ow, inv_sigma2, log_scale_factor, n_levels, th, pts, pdesc, present, bad, nobs, in_kf):
    """The reference's ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th) executed (op 11).  KF = point-frame dict (keys_un, uright, desc, bounds,
    scale_factors, claimed = the keypoint holds a map point); kf_obs = Observations() of those map points; pts MAP_POINT_DTYPE.
    Returns (nFused, events [n, 3] = (map point, key-frame keypoint, action)) or None."""
    pts = np.ascontiguousarray(pts, MAP_POINT_DTYPE)
    M = len(pts)
    b = struct.pack('<2i', 0x4d544348, 11) + _f32(KF['bounds']) + _point_frame_bytes(KF) + np.asarray(kf_obs, np.uint8).tobytes()
    b += _f32(cam5) + _f32(Rcw) + _f32(tcw) + _f32(Ow) + _f32(inv_sigma2) + struct.pack('<fifi', float(log_scale_factor), int(n_levels), float(th), M)
    pd = np.ascontiguousarray(pdesc, np.uint8).reshape(-1, 32)
    for i in range(M):
        b += pts[i].tobytes() + pd[i].tobytes() + struct.pack('<4B', int(present[i]), int(bad[i]), int(nobs[i]), int(in_kf[i]))
    raw = _run_ref_match(b)
    if raw is None:
        return None
    nf, ne = struct.unpack_from('<2i', raw, 0)
    return nf, np.frombuffer(raw, np.int32, 3 * ne, 8).reshape(ne, 3).copy()


def _kf_camera_bytes(cam):
    """fx fy cx cy bf, Rcw, tcw, Ow, mfLogScaleFactor, mnScaleLevels from a FRUSTUM_CAM record (read_kf_camera of the harness)"""
    c = np.asarray(cam).reshape(())
    return (_f32([c['fx'], c['fy'], c['cx'], c['cy'], c['bf']]) + _f32(c['Rcw']) + _f32(c['tcw']) + _f32(c['Ow'])
            + struct.pack('<fi', float(c['log_scale_factor']), int(c['n_levels'])))


def _map_point_records(pts, pdesc, flags, slots=None):
    """hvo_map_point + descriptor + 4 flag bytes + int32 slot per map point (read_map_point of the harness)"""
    pts = np.ascontiguousarray(pts, MAP_POINT_DTYPE)
    pd = np.ascontiguousarray(pdesc, np.uint8).reshape(-1, 32)
    fl = np.ascontiguousarray(flags, np.uint8).reshape(-1, 4)
    sl = np.full(len(pts), -1, np.int32) if slots is None else np.asarray(slots, np.int32)
    rec = np.zeros(len(pts), np.dtype([('p', MAP_POINT_DTYPE), ('d', 'u1', (32,)), ('f', 'u1', (4,)), ('s', '<i4')]))
    rec['p'] = pts; rec['d'] = pd; rec['f'] = fl; rec['s'] = sl
    return rec.tobytes()


def ref_search_by_projection_scw(KF, cam, Scw, th, pts, pdesc, bad, found_slot):
    """The reference's ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) executed (op 12).  KF = point-frame dict whose
    `claimed` marks the entries of vpMatched that are set at call time; found_slot[i] >= 0: point i itself sits in vpMatched[found_slot[i]].
    Returns (nmatches, vpMatched ids [N]: candidate index, -100 - i for the pre-set entry of slot i, -1 empty) or None."""
    M = len(pts)
    fl = np.zeros((M, 4), np.uint8); fl[:, 1] = bad; fl[:, 3] = np.asarray(found_slot) >= 0
    b = struct.pack('<2i', 0x4d544348, 12) + _f32(KF['bounds']) + _point_frame_bytes(KF) + _kf_camera_bytes(cam) + _f32(Scw)
    b += struct.pack('<2i', int(th), M) + _map_point_records(pts, pdesc, fl, found_slot)
    raw = _run_ref_match(b)
    if raw is None:
        return None
    (nm,) = struct.unpack_from('<i', raw, 0)
    return nm, np.frombuffer(raw, np.int32, len(KF['keys_un']), 4).copy()


def ref_fuse_scw(KF, kf_bad, cam, Scw, th, pts, pdesc, bad, held_slot):
    """The reference's ORBmatcher::Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) executed (op 13).  KF['claimed'] = the keypoint holds a map
    point (kf_bad: that point isBad()); held_slot[i] >= 0: candidate i is the map point of that slot.  Returns (nFused, replace ids [M],
    AddObservation events [n, 2] = (map point, keypoint), key-frame slots after the call [N]) or None."""
    M, N = len(pts), len(KF['keys_un'])
    fl = np.zeros((M, 4), np.uint8); fl[:, 1] = bad; fl[:, 3] = np.asarray(held_slot) >= 0
    b = struct.pack('<2i', 0x4d544348, 13) + _f32(KF['bounds']) + _point_frame_bytes(KF) + np.asarray(kf_bad, np.uint8).tobytes()
    b += _kf_camera_bytes(cam) + _f32(Scw) + struct.pack('<fi', float(th), M) + _map_point_records(pts, pdesc, fl, held_slot)
    raw = _run_ref_match(b)
    if raw is None:
        return None
    (nf,) = struct.unpack_from('<i', raw, 0)
    rep = np.frombuffer(raw, np.int32, M, 4).copy()
    (ne,) = struct.unpack_from('<i', raw, 4 + 4 * M)
    ev = np.frombuffer(raw, np.int32, 2 * ne, 8 + 4 * M).reshape(ne, 2).copy()
    slots = np.frombuffer(raw, np.int32, N, 8 + 4 * M + 8 * ne).copy()
    return nf, rep, ev, slots


def _kf_with_points_bytes(KF, cam, mp):
    """point frame + camera + the map point of every keypoint (mp = dict(pts, desc, has, bad[, found]))"""
    n = len(KF['keys_un'])
    fl = np.zeros((n, 4), np.uint8); fl[:, 0] = mp['has']; fl[:, 1] = mp['bad']
    if 'found' in mp:
        fl[:, 2] = mp['found']
    return _map_point_records(mp['pts'], mp['desc'], fl)


def ref_search_by_sim3(KF1, cam1, mp1, KF2, cam2, mp2, matches12, s12, R12, t12, th):
    """The reference's ORBmatcher::SearchBySim3 executed (op 14).  matches12 [N1] = index of the KF2 map point already matched, or -1.
    Returns (nFound, vpMatches12 ids [N1]) or None."""
    b = struct.pack('<2i', 0x4d544348, 14) + _f32(KF1['bounds'])
    b += _point_frame_bytes(KF1) + _kf_camera_bytes(cam1) + _kf_with_points_bytes(KF1, cam1, mp1)
    b += _point_frame_bytes(KF2) + _kf_camera_bytes(cam2) + _kf_with_points_bytes(KF2, cam2, mp2)
    b += np.asarray(matches12, np.int32).tobytes() + struct.pack('<f', float(s12)) + _f32(R12) + _f32(t12) + struct.pack('<f', float(th))
    raw = _run_ref_match(b)
    if raw is None:
        return None
    (nf,) = struct.unpack_from('<i', raw, 0)
    return nf, np.frombuffer(raw, np.int32, len(KF1['keys_un']), 4).copy()


def ref_search_by_projection_reloc(Cur, cam4, Tcw, log_scale_factor, n_levels, th, orb_dist, check_ori, kf_keys, mp):
    """The reference's ORBmatcher::SearchByProjection(CurrentFrame, KeyFrame*, sAlreadyFound, th, ORBdist) executed (op 15).  Cur['claimed'] =
    the frame's keypoint holds a map point; mp = the key frame's map points per keypoint (pts, desc, has, bad, found).
    Returns (nmatches, ids [N]: key-frame keypoint whose map point the frame keypoint received, -2 kept its own, -1 none) or None."""
    k = np.ascontiguousarray(kf_keys, KP_DTYPE)
    b = struct.pack('<2i', 0x4d544348, 15) + _f32(Cur['bounds']) + _point_frame_bytes(Cur) + _f32(cam4) + _f32(Tcw)
    b += struct.pack('<fifii', float(log_scale_factor), int(n_levels), float(th), int(orb_dist), int(check_ori))
    b += struct.pack('<i', len(k)) + k.tobytes() + _kf_with_points_bytes(dict(keys_un=k), None, mp)
    raw = _run_ref_match(b)
    if raw is None:
        return None
    (nm,) = struct.unpack_from('<i', raw, 0)
    return nm, np.frombuffer(raw, np.int32, len(Cur['keys_un']), 4).copy()


def ref_search_by_bow_kf(KF1, KF2, nnratio=0.75, check_ori=True):
    """The reference's ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) executed (op 16).  KFi = dict(keys_un, desc, has_mappoint, bad, featvec).
    Returns (nmatches, vpMatches12 [N1] = keypoint of pKF2 whose map point was matched, or -1) or None."""
    def kf(K):
        k = np.ascontiguousarray(K['keys_un'], KP_DTYPE)
        bad = np.zeros(len(k), np.uint8) if K.get('bad') is None else np.asarray(K['bad'], np.uint8)
        return (struct.pack('<i', len(k)) + k.tobytes() + np.ascontiguousarray(K['desc'], np.uint8).tobytes()
                + np.asarray(K['has_mappoint'], np.uint8).tobytes() + bad.tobytes() + _featvec_bytes(K['featvec']))
    b = struct.pack('<2i', 0x4d544348, 16) + _f32([0, 0, 640, 480]) + kf(KF1) + kf(KF2) + struct.pack('<fi', float(nnratio), int(check_ori))
    raw = _run_ref_match(b)
    if raw is None:
        return None
    (nm,) = struct.unpack_from('<i', raw, 0)
    return nm, np.frombuffer(raw, np.int32, len(KF1['keys_un']), 4).copy()


def ref_line_bf(ldesc1, ldesc2, TH, nnratio, nnr):
    """The reference's LSDmatcher::FrameBFMatch(ldesc1, ldesc2, LineMatches, TH), match(desc1, desc2, nnr, matches_12) and
    SearchDouble(InitialFrame, CurrentFrame, LineMatches) executed (op 17).  Returns (LineMatches [n1], n_nnr, matches_12 [n1], n_double,
    LineMatches of SearchDouble [n1]) or None."""
    d1 = np.ascontiguousarray(ldesc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(ldesc2, np.uint8).reshape(-1, 32)
    n1 = len(d1)
    b = struct.pack('<2i', 0x4d544348, 17) + _f32([0, 0, 640, 480]) + struct.pack('<i', n1) + d1.tobytes() + struct.pack('<i', len(d2)) + d2.tobytes()
    b += struct.pack('<3f', float(TH), float(nnratio), float(nnr))
    raw = _run_ref_match(b)
    if raw is None:
        return None
    lm = np.frombuffer(raw, np.int32, n1, 0).copy()
    (n_nnr,) = struct.unpack_from('<i', raw, 4 * n1)
    m12 = np.frombuffer(raw, np.int32, n1, 4 * n1 + 4).copy()
    (n_dbl,) = struct.unpack_from('<i', raw, 8 * n1 + 4)
    dbl = np.frombuffer(raw, np.int32, n1, 8 * n1 + 8).copy()
    return lm, n_nnr, m12, n_dbl, dbl


def ref_line_by_descriptor(kf_ldesc, kf_has_mapline, ldesc, nnratio):
    """The reference's LSDmatcher::SearchByDescriptor(pKF, currentF, vpMapLineMatches) and SearchDouble(KF, CurrentFrame) executed (op 18).
    Returns (nmatches, ids [NL] = key-frame line whose MapLine the frame line received, n_double, ids of SearchDouble [NL]) or None."""
    d1 = np.ascontiguousarray(kf_ldesc, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(ldesc, np.uint8).reshape(-1, 32)
    n2 = len(d2)
    b = struct.pack('<2i', 0x4d544348, 18) + _f32([0, 0, 640, 480]) + struct.pack('<i', len(d1)) + d1.tobytes() + np.asarray(kf_has_mapline, np.uint8).tobytes()
    b += struct.pack('<i', n2) + d2.tobytes() + struct.pack('<f', float(nnratio))
    raw = _run_ref_match(b)
    if raw is None:
        return None
    (n1,) = struct.unpack_from('<i', raw, 0)
    by = np.frombuffer(raw, np.int32, n2, 4).copy()
    (nd,) = struct.unpack_from('<i', raw, 4 + 4 * n2)
    dbl = np.frombuffer(raw, np.int32, n2, 8 + 4 * n2).copy()
    return n1, by, nd, dbl


def ref_line_triangulation(ldesc1, has1, ldesc2, has2, nnratio):
    """The reference's LSDmatcher::SearchForTriangulation(pKF1, pKF2, vector<pair>&) and (pKF1, pKF2, vector<int>&, isDouble) for isDouble =
    false / true executed (op 19).  Returns (n, pairs [n,2], n_single, match_single [NL1], n_double, match_double [NL1]) or None."""
    d1 = np.ascontiguousarray(ldesc1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(ldesc2, np.uint8).reshape(-1, 32)
    n1 = len(d1)
    b = struct.pack('<2i', 0x4d544348, 19) + _f32([0, 0, 640, 480])
    b += struct.pack('<i', n1) + d1.tobytes() + np.asarray(has1, np.uint8).tobytes()
    b += struct.pack('<i', len(d2)) + d2.tobytes() + np.asarray(has2, np.uint8).tobytes() + struct.pack('<f', float(nnratio))
    raw = _run_ref_match(b)
    if raw is None:
        return None
    n0, npairs = struct.unpack_from('<2i', raw, 0)
    pairs = np.frombuffer(raw, np.int32, 2 * npairs, 8).reshape(npairs, 2).copy()
    off = 8 + 8 * npairs
    out = [n0, pairs]
    for _ in range(2):
        (n,) = struct.unpack_from('<i', raw, off)
        out += [n, np.frombuffer(raw, np.int32, n1, off + 4).copy()]
        off += 4 + 4 * n1
    return tuple(out)


def surface_normals(depth16, factor, fx, fy, cx, cy, max_depth_change=0.05, smoothing=10.0, want_dist=False):
    d = np.ascontiguousarray(depth16, np.uint16)
    h, w = d.shape
    cw, ch = -(-w // 3), -(-h // 3)
    out = np.empty(((ch // 2) * (cw // 2), 8), np.float32)
    dist = np.empty((ch, cw), np.float32)
    n = lib().orc_surface_normals(_p(d), w, h, C.c_float(factor), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                                  C.c_float(max_depth_change), C.c_float(smoothing), _p(out), _p(dist))
    assert n == len(out)
    return (out, dist) if want_dist else out


# ---- windowed line matchers (Frame.cc:849-872, 1557-1631; LSDmatcher.cpp:561-664, 709-801) -----------------------------
LPROJ_QUERY_DTYPE = np.dtype([('x1', '<f4'), ('y1', '<f4'), ('x2', '<f4'), ('y2', '<f4'), ('r', '<f4'), ('cos_th', '<f4'), ('dir', '<f8', (3,)),
                              ('length', '<f4'), ('claims', '<i4'), ('reserved', '<i4', (2,))])
assert LPROJ_QUERY_DTYPE.itemsize == 64


def line_grid_build(keylines, bounds):
    kl = np.ascontiguousarray(keylines, KL_DTYPE)
    cnt = np.empty(64 * 48, np.int32)
    f = C.c_float
    lib().orc_line_grid_build.restype = C.c_int
    m = lib().orc_line_grid_build(_p(kl), C.c_int(len(kl)), f(bounds[0]), f(bounds[1]), f(bounds[2]), f(bounds[3]), _p(cnt), None)
    items = np.empty(max(m, 1), np.int32)
    lib().orc_line_grid_build(_p(kl), C.c_int(len(kl)), f(bounds[0]), f(bounds[1]), f(bounds[2]), f(bounds[3]), _p(cnt), _p(items))
    return cnt, items[:m]


def line_features_in_area(keylines, func3, bounds, x1, y1, x2, y2, r, TH=0.998):
    kl = np.ascontiguousarray(keylines, KL_DTYPE)
    fn = np.ascontiguousarray(func3, np.float64)
    out = np.empty(max(len(kl), 1), np.int32)
    f = C.c_float
    lib().orc_line_features_in_area.restype = C.c_int
    n = lib().orc_line_features_in_area(_p(kl), _p(fn), C.c_int(len(kl)), f(bounds[0]), f(bounds[1]), f(bounds[2]), f(bounds[3]), f(x1), f(y1),
                                        f(x2), f(y2), f(r), f(TH), _p(out), C.c_int(len(out)))
    return out[:n].copy()


def line_search_projection(keylines, func3, desc, lines3d, bounds, queries, qdesc, claimed=None, mode=0, nnratio=0.95):
    kl = np.ascontiguousarray(keylines, KL_DTYPE)
    fn = np.ascontiguousarray(func3, np.float64)
    d = np.ascontiguousarray(desc, np.uint8)
    l3 = np.zeros((len(kl), 6)) if lines3d is None else np.ascontiguousarray(lines3d, np.float64)
    q = np.ascontiguousarray(queries, LPROJ_QUERY_DTYPE)
    qd = np.ascontiguousarray(qdesc, np.uint8)
    cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
    idx = np.full(max(len(q), 1), -1, np.int32); dist = np.full(max(len(q), 1), 256, np.int32)
    f = C.c_float
    lib().orc_line_search_projection.restype = C.c_int
    nm = lib().orc_line_search_projection(_p(kl), _p(fn), _p(d), _p(l3), C.c_int(len(kl)), f(bounds[0]), f(bounds[1]), f(bounds[2]), f(bounds[3]),
                                          _p(q), _p(qd), C.c_int(len(q)), _p(cl) if cl is not None else None, C.c_int(mode), f(nnratio),
                                          _p(idx), _p(dist))
    return idx[:len(q)], dist[:len(q)], nm


# ---- LPVO normals (Manhattan::computeNormalsLPVO, Manhattan.cpp:237-393) ------------------------------------------
def integral_f32(img):
    """cv::integral(CV_32F -> CV_64F) without the leading zero row / column."""
    a = np.ascontiguousarray(img, np.float32)
    out = np.empty(a.shape, np.float64)
    lib().orc_integral_f32(_p(a), C.c_int(a.shape[1]), C.c_int(a.shape[0]), _p(out))
    return out


def normalize3(v):
    """cv::normalize of a 3-vector of doubles (NORM_L2)."""
    a = np.ascontiguousarray(v, np.float64).reshape(3)
    out = np.empty(3, np.float64)
    lib().orc_normalize3(_p(a), _p(out))
    return out


def lpvo_normals(depth16, factor, fx, fy, cx, cy):
    """-> (normals [n,3] float64, depth [n] float32, pixel (u, v) [n,2] int32) in the reference's push_back order."""
    d = np.ascontiguousarray(depth16, np.uint16)
    h, w = d.shape
    cap = ((h - 1 - 10 + 14) // 15) * ((w - 1 - 10 + 14) // 15)
    nrm = np.empty((cap, 3), np.float64); dep = np.empty(cap, np.float32); pix = np.empty((cap, 2), np.int32)
    lib().orc_lpvo_normals.restype = C.c_int
    n = lib().orc_lpvo_normals(_p(d), C.c_int(w), C.c_int(h), C.c_float(factor), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                               _p(nrm), _p(dep), _p(pix), C.c_int(cap))
    assert 0 <= n <= cap
    return nrm[:n], dep[:n], pix[:n]


def ref_lpvo(depth16, factor, fx, fy, cx, cy, as_built=False):
    """The reference's own Manhattan::computeNormalsLPVO executed (oracle/_ref/ref_lpvo: src/Manhattan.cpp:237-393 + removeMatRow / removeMatCol
    compiled with their cv::Rect body; as_built=True: the memcpy body Manhattan.cpp is built with, which mis-sizes the CV_64F rows).
    Returns (normals [n,3] float64, depth [n] float32) in push_back order, or None when the binary is absent."""
    exe = ref_bin('ref_lpvo_asbuilt' if as_built else 'ref_lpvo')
    if exe is None:
        return None
    d = np.ascontiguousarray(depth16, np.uint16)
    h, w = d.shape
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, 'in.bin'), os.path.join(td, 'out.bin')
        with open(fi, 'wb') as f:
            f.write(struct.pack('<2i5f', w, h, float(factor), float(fx), float(fy), float(cx), float(cy)) + d.tobytes())
        subprocess.check_call([exe, fi, fo], stdout=subprocess.DEVNULL)
        raw = open(fo, 'rb').read()
    (n,) = struct.unpack_from('<i', raw, 0)
    return np.frombuffer(raw, np.float64, 3 * n, 4).reshape(n, 3).copy(), np.frombuffer(raw, np.float32, n, 4 + 24 * n).copy()


# ---- whole front-end (bench.py CPU arm) ----------------------------------------------------------------------
def frontend_batch(gray, depth, cam, stages=15, nthreads=1, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7,
                   nlines=200):
    """Frame-parallel oracle run of ORB (1) | lines (2) | planes (4) | normals (8) over [n,h,w] frames.
    cam = (depth_factor, fx, fy, cx, cy).  Returns counts [n,4] = keypoints, lines, planes, normals."""
    gray = np.ascontiguousarray(gray, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    n, h, w = gray.shape
    counts = np.zeros((n, 4), np.int32)
    f = C.c_float
    lib().orc_frontend_batch(_p(gray), _p(depth), C.c_int(n), C.c_int(w), C.c_int(h), C.c_int(nthreads), C.c_int(stages),
                             C.c_int(nfeatures), f(scale_factor), C.c_int(nlevels), C.c_int(ini_th), C.c_int(min_th), C.c_int(nlines),
                             f(cam[0]), f(cam[1]), f(cam[2]), f(cam[3]), f(cam[4]), _p(counts))
    return counts
