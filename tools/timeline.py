"""Profiling aid: per-stream kernel start times of one front-end step (hvo_timeline_*), to see how the pipelines overlap."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hvo_b200 as hvo
from bench import make_frames, CAM, DEPTH_FACTOR, BF, NLINES

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=1024)
ap.add_argument('--lanes', type=int, default=0)
ap.add_argument('--stages', type=int, default=15)
ap.add_argument('--steps', type=int, default=2)
ap.add_argument('--dist', action='store_true', help='under torchrun: NCCL process group + barrier first (how bench.py runs at N > 1)')
a = ap.parse_args()
B = a.batch
gray, depth = make_frames(min(B, 512))
rep = (B + len(gray) - 1) // len(gray)
gray = np.concatenate([gray] * rep)[:B]; depth = np.concatenate([depth] * rep)[:B]
rank, world, lrank = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
dev = torch.device('cuda', lrank if a.dist else 0)
if a.dist:
    import torch.distributed as dist
    torch.cuda.set_device(lrank)
    dist.init_process_group('nccl', device_id=dev)
fe = hvo.FrameFrontEnd(640, 480, CAM['fx'], CAM['fy'], CAM['cx'], CAM['cy'], DEPTH_FACTOR, bf=BF, n_lines=NLINES, stages=a.stages, line_cull=True,
                       lanes=a.lanes, membership='u8', max_batch=B, device=(int(os.environ.get('LOCAL_RANK', 0)) if a.dist else 0))
d_gray = torch.from_numpy(gray).to(dev); d_depth = torch.from_numpy(depth.view(np.int16)).to(dev)
d_out = {k: torch.empty(int(np.prod(sh)) * np.dtype(dt).itemsize, dtype=torch.uint8, device=dev) for k, (sh, dt) in fe.output_shapes(B).items()}
ptrs = {k: v.data_ptr() for k, v in d_out.items()}
for _ in range(3):
    fe.extract_batch_device(d_gray.data_ptr(), d_depth.data_ptr(), B, ptrs)
fe.sync()
if a.dist:
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
hvo.timeline(True)
fe.timer_start()
for _ in range(a.steps):
    fe.extract_batch_device(d_gray.data_ptr(), d_depth.data_ptr(), B, ptrs)
ms = fe.timer_stop()
tl = hvo.timeline_dump()
hvo.timeline(False)
print(f'rank {rank}/{world} batch {B} lanes {fe.lanes} stages {a.stages}: {ms / a.steps:.2f} ms/step')
by = {}
for t, s, n in tl:
    by.setdefault(s, []).append((t, n))
for s in sorted(by):
    ev = by[s]
    print(f'stream {s}: ' + '  '.join(f'{n}@{t:.1f}' for t, n in ev))

if a.dist:
    fe.close(); dist.barrier(); dist.destroy_process_group()
