"""Profiling aid: device-resident per-stage timing of the ORB extractor on a batch of synthetic frames."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hvo_b200 as hvo
from hvo_b200 import synth
from bench import make_frames

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=1184)
ap.add_argument('--cfg', default='S1')
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--nfeatures', type=int, default=1000)
a = ap.parse_args()
B = a.batch
gray, depth = make_frames(min(B, 256), cfg=a.cfg)
if B > len(gray):
    gray = np.concatenate([gray] * ((B + len(gray) - 1) // len(gray)))[:B]
H, W = gray.shape[1:]
ex = hvo.ORBextractor(a.nfeatures, 1.2, 8, 20, 7, width=W, height=H, max_batch=B)
dev = torch.device('cuda', 0)
d_gray = torch.from_numpy(gray).to(dev)
cap = ex.capacity
d_kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev)
d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
d_cnt = torch.empty((B,), dtype=torch.int32, device=dev)
torch.cuda.synchronize()
def step():
    ex.extract_batch_device(d_gray.data_ptr(), B, d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr())
for _ in range(2):
    step()
ex.sync()
ex.timer_start()
for _ in range(a.reps):
    step()
ms = ex.timer_stop() / a.reps
ex.set_profiling(True)
acc = {}
for _ in range(a.reps):
    step(); ex.sync()
    for k, v in ex.stage_times().items():
        acc[k] = acc.get(k, 0.0) + v / a.reps
px = {'S1': 950532, 'S2': 950532}.get(a.cfg)
print(json.dumps(dict(batch=B, cfg=a.cfg, orb_ms=ms, orb_fps=B / ms * 1e3, stage_ms=acc, mean_kps=float(d_cnt.float().mean().item()),
                      fast_frac_hbm=(px * B / (acc['fast'] * 1e-3) / 1e9 / 6556.2) if px else None)))
