"""Profiling aid: device-resident timing of plane extraction + surface normals on a batch of synthetic depth frames."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hvo_b200 as hvo
from hvo_b200 import synth
from bench import make_frames

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--cfg', default='S1')
ap.add_argument('--reps', type=int, default=3)
a = ap.parse_args()
B = a.batch
gray, depth = make_frames(min(B, 256), cfg=a.cfg)
if B > len(depth):
    depth = np.concatenate([depth] * ((B + len(depth) - 1) // len(depth)))[:B]
H, W = depth.shape[1:]
c = synth.CONFIGS[a.cfg]
K = np.array([[c['fx'], 0, c['cx']], [0, c['fy'], c['cy']], [0, 0, 1]], np.float32)
pd = hvo.PlaneDetection(W, H, max_batch=B)
pd.readDepthImage(depth[0], K, np.float32(1.0 / c['factor']))
dev = torch.device('cuda', 0)
d_depth = torch.from_numpy(depth.view(np.int16)).to(dev)
d_n = torch.empty((B,), dtype=torch.int32, device=dev)
d_pl = torch.empty((B, 64, 7), dtype=torch.float64, device=dev)
d_mem = torch.empty((B, H * W), dtype=torch.int32, device=dev)
torch.cuda.synchronize()
def step():
    pd.detect_batch_device(d_depth.data_ptr(), B, d_n.data_ptr(), d_pl.data_ptr(), 64, d_mem.data_ptr())
step(); pd.sync()
pd.timer_start()
for _ in range(a.reps):
    step()
ms = pd.timer_stop() / a.reps
pd.timer_start()
for _ in range(a.reps):
    pd.blocks_device(d_depth.data_ptr(), B)
ms_b = pd.timer_stop() / a.reps
sn = hvo.SurfaceNormals(W, H, c['fx'], c['fy'], c['cx'], c['cy'], 1.0 / c['factor'], max_batch=B)
d_out = torch.empty((B, sn.count, 8), dtype=torch.float32, device=dev)
sn.compute_device(d_depth.data_ptr(), B, d_out.data_ptr()); sn.sync()
sn.timer_start()
for _ in range(a.reps):
    sn.compute_device(d_depth.data_ptr(), B, d_out.data_ptr())
ms_n = sn.timer_stop() / a.reps
print(json.dumps(dict(batch=B, cfg=a.cfg, planes_ms=ms, planes_fps=B / ms * 1e3, blocks_ms=ms_b, normals_ms=ms_n, normals_fps=B / ms_n * 1e3,
                      mean_planes=float(d_n.float().mean().item()), phase_cycles=pd.phase_cycles(0))))
