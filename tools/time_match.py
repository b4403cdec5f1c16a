"""Profiling aid: brute-force Hamming knn-2 at the C4 stress sizes (BASELINE.json configs[3]): 2000 ORB queries x 50 000 map descriptors
and 200 LBD queries x 5 000, device-resident, next to cv2.BFMatcher on the host cores when cv2 is importable."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hvo_b200 as hvo
from hvo_b200 import synth

dev = torch.device('cuda', 0)
bf = hvo.BFMatcherHamming()
out = {}
for name, nq, nt, seed in (('orb_2000x50000', 2000, 50000, 4), ('lbd_200x5000', 200, 5000, 5)):
    q, t = synth.descriptors_S4(nq=nq, nt=nt, seed=seed)
    dq, dt = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
    di = torch.empty((nq, 2), dtype=torch.int32, device=dev); dd = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    for _ in range(3):
        bf.knn2_device(dq.data_ptr(), nq, dt.data_ptr(), nt, di.data_ptr(), dd.data_ptr())
    bf.sync()
    reps = 50
    bf.timer_start()
    for _ in range(reps):
        bf.knn2_device(dq.data_ptr(), nq, dt.data_ptr(), nt, di.data_ptr(), dd.data_ptr())
    ms = bf.timer_stop() / reps
    popc = nq * nt * 8
    r = dict(ms=ms, pairs_per_s=nq * nt / ms * 1e3, popc32_per_s=popc / ms * 1e3)
    t0 = time.perf_counter()
    idx, dist = bf.knnMatch2(q, t)          # host buffers through the C ABI
    r['host_call_ms'] = 1e3 * (time.perf_counter() - t0)
    try:
        import cv2
        m = cv2.BFMatcher(cv2.NORM_HAMMING, False)
        t0 = time.perf_counter()
        mm = m.knnMatch(q, t, k=2)
        r['cv2_ms'] = 1e3 * (time.perf_counter() - t0)
        r['cv2_threads'] = cv2.getNumThreads()
        r['same_as_cv2'] = bool(np.array_equal(idx, np.array([[a.trainIdx, b.trainIdx] for a, b in mm], np.int32)))
    except ImportError:
        pass
    out[name] = r
print(json.dumps(out))
