"""Debug aid for the grow kernel (HVO_LSD_GUARD builds): segment count and segments against the cv2 golden, and the time of a batch."""
import sys, os, time, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hvo_b200 as hvo
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/golden/lsd_cv2.npz'))
for key in ('s1_crop', 's2_crop', 's1_odd', 'noise', 'flat'):
    if key + '_img' not in g.files:
        continue
    img = g[key + '_img']
    ex = hvo.LINEextractor(1, 1.2, 200, 0.125, width=img.shape[1], height=img.shape[0])
    frames = np.ascontiguousarray(img[None])
    cap = ex.segment_capacity
    seg = np.zeros((1, cap, 4), np.float32)
    counts = np.zeros(1, np.int32)
    t = time.time()
    st = hvo.lib().hvo_line_detect_batch(ex._h, frames.ctypes.data_as(C.c_void_p), 1, seg.ctypes.data_as(C.c_void_p), cap, counts.ctypes.data_as(C.c_void_p))
    gold = g[key + '_segments'].reshape(-1, 4)
    n = int(counts[0])
    same = n == len(gold) and np.abs(seg[0, :n] - gold).max() <= 1e-3
    print(key, 'status', st, hvo.lib().hvo_last_error().decode() if st else '', 'count', n, 'golden', len(gold), 'same', same, 'ms', round((time.time() - t) * 1e3, 2), flush=True)
    if n < 0:
        print('  guard tripped: m=%08x remaining=%08x' % tuple(seg[0, 0, :2].view(np.uint32)))
