"""Profiling aid: device-resident timing of the line extractor (LSD + keylines + LBD) on a batch of synthetic frames.
    python tools/time_lines.py [--batch B] [--cfg S1] [--reps R]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hvo_b200 as hvo
from bench import make_frames

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--cfg', default='S1')
ap.add_argument('--reps', type=int, default=5)
a = ap.parse_args()
B = a.batch
gray, depth = make_frames(min(B, 256), cfg=a.cfg)
if B > len(gray):
    gray = np.concatenate([gray] * ((B + len(gray) - 1) // len(gray)))[:B]
H, W = gray.shape[1:]
ex = hvo.LINEextractor(1, 1.2, 200, 0.125, width=W, height=H, max_batch=B)
dev = torch.device('cuda', 0)
d_gray = torch.from_numpy(gray).to(dev)
ml = ex.max_lines
d_kl = torch.empty((B, ml, 17), dtype=torch.float32, device=dev)
d_desc = torch.empty((B, ml, 32), dtype=torch.uint8, device=dev)
d_lv = torch.empty((B, ml, 3), dtype=torch.float64, device=dev)
d_cnt = torch.empty((B,), dtype=torch.int32, device=dev)
torch.cuda.synchronize()
def step():
    ex.extract_batch_device(d_gray.data_ptr(), B, d_kl.data_ptr(), d_desc.data_ptr(), d_lv.data_ptr(), d_cnt.data_ptr())
for _ in range(2):
    step()
ex.sync()
ex.timer_start()
for _ in range(a.reps):
    step()
ms = ex.timer_stop() / a.reps
ex.set_profiling(True)
step(); ex.sync()
st = ex.stage_times()
print(json.dumps(dict(batch=B, cfg=a.cfg, ms_per_batch=ms, frames_per_s=B / ms * 1e3, stage_ms=st,
                      mean_lines=float(d_cnt.float().mean().item()))))
