#!/usr/bin/env python
"""bench.py — front-end frames/s at 640x480 (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of B synthetic TUM-fr3-shaped 640x480 RGB-D frames per
GPU (config C1 of BASELINE.json).  Frames are independent, so ranks shard them with no data-path collective
("scaling": "weak"); torch.distributed is used only for the barrier and the max-over-ranks of the time.

  value  whole-job frames/s with the inputs already resident in HBM (CUDA events on the library's stream)
  e2e    the same metric through the reference-facing C-ABI call with HOST (pinned) buffers: host->device
         copy of gray+depth and device->host read of keypoints/descriptors inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md section "Measurement"

--impl reference times the reference's CPU implementation of the same path on the host cores: the oracle port
(oracle/, a restatement that is bit-identical to the reference's own ORBextractor.cc compiled in oracle/_ref;
the reference binary itself cannot be built: OpenCV/PCL/Eigen/Pangolin are absent from this image).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'front-end frames/s @640x480'
ORB = dict(nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7)  # TUM3.yaml:41-54
DEPTH_FACTOR, BF = 1.0 / 5000.0, 40.0
# algorithmic bytes per 640x480 frame (SURVEY.md section 8d / BASELINE.md section 5)
BYTES_PYRAMID, BYTES_FAST, BYTES_BLUR, BYTES_DESCRIBE_PATCH, BYTES_OUT = 1569878, 950532, 1901064, 1922000, 60000


def _gen(args):
    from hvo_b200 import synth
    cfg, i = args
    return synth.frame(cfg, i)


def make_frames(n, start=0, cfg='S1'):
    """n distinct synthetic frames (gray u8, depth u16); generated in parallel, cached under /tmp."""
    cache = f'/tmp/hvo_bench_{cfg}_{start}_{n}.npz'
    if os.path.exists(cache):
        z = np.load(cache)
        return z['gray'], z['depth']
    import hvo_b200  # noqa: F401  (registers the package alias for the workers)
    procs = max(1, min(32, (os.cpu_count() or 2) - 1))
    with mp.get_context('fork').Pool(procs) as pool:
        res = pool.map(_gen, [(cfg, start + i) for i in range(n)], chunksize=max(1, n // (4 * procs)))
    gray = np.stack([r[0] for r in res])
    depth = np.stack([r[1] for r in res])
    try:
        np.savez(cache, gray=gray, depth=depth)
    except OSError:
        pass
    return gray, depth


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.lines:
            p = [x.strip() for x in line.split(',')]
            if len(p) < 9:
                continue
            try:
                smax = float(p[2])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(p[1]))
                    for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[5:9]):
                        if v.lower().startswith('active'):
                            reasons.add(name)
            except ValueError:
                continue
        if not sm:  # region shorter than the sampling period: take every sample we have
            for ts, line in self.lines:
                p = [x.strip() for x in line.split(',')]
                try:
                    sm.append(float(p[1]))
                except (ValueError, IndexError):
                    pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def cpu_baseline(gray, target_s=12.0):
    """The oracle port timed on the host cores: frame-parallel over all cores (frames are independent)."""
    import oracle
    cores = os.cpu_count() or 1
    oracle.orb_extract_batch(gray[:cores], nthreads=cores, **ORB)  # warm-up / page-in
    n, dt, counts = 0, 0.0, []
    t0 = time.perf_counter()
    while dt < target_s:  # bounded sample: whole passes over the same frames until ~target_s of CPU work
        counts.append(oracle.orb_extract_batch(gray, nthreads=cores, **ORB))
        n += len(gray)
        dt = time.perf_counter() - t0
    return dict(value=n / dt, unit='frames/s', cores=cores, kind='port',
                sample=f'{n} frames ({n // len(gray)} passes over the same {len(gray)} synthetic 640x480 frames), ORB stage, oracle '
                       f'C++ port (-O3), one frame per thread on {cores} threads, {dt:.1f} s, '
                       f'mean {float(np.mean(counts)):.0f} keypoints/frame')


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) on all host cores; rank 0 only."""
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    sample = max(cores, min(args.batch, 4 * cores))
    gray, _ = make_frames(sample)
    for _ in range(args.warmup):
        oracle.orb_extract_batch(gray, nthreads=cores, **ORB)
    t = time.perf_counter()
    for _ in range(args.steps):
        oracle.orb_extract_batch(gray, nthreads=cores, **ORB)
    dt = time.perf_counter() - t
    v = sample * args.steps / dt
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': f'C1: synthetic TUM-fr3-shaped 640x480, ORB 1000 features 8 levels x1.2; bounded sample of {sample} frames/step',
                   'stages': ['orb'], 'note': 'reference CPU path = oracle port (bit-identical to the reference ORBextractor.cc built in oracle/_ref); '
                                              'the reference binary needs OpenCV/PCL/Eigen/Pangolin, absent here'},
        'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': f'{sample} frames x {args.steps} steps'},
        'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='frames per GPU per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--device-only', action='store_true', help='profiling aid: only the device-resident loop')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import hvo_b200 as hvo
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    B, W, H = args.batch, 640, 480
    gray, depth = make_frames(B, start=rank * B)  # every rank gets its own frames (frame-sharded, weak scaling)
    ex = hvo.ORBextractor(ORB['nfeatures'], ORB['scale_factor'], ORB['nlevels'], ORB['ini_th'], ORB['min_th'],
                          width=W, height=H, max_batch=B, device=local_rank)
    cap = ex.capacity
    dev = torch.device('cuda', local_rank)
    d_gray = torch.from_numpy(gray).to(dev)
    d_depth = torch.from_numpy(depth.view(np.int16)).to(dev)
    d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_counts = torch.empty((B,), dtype=torch.int32, device=dev)
    d_kd = torch.empty((B, cap), dtype=torch.float32, device=dev)
    d_ku = torch.empty((B, cap), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    def step_device():
        ex.extract_batch_device(d_gray.data_ptr(), B, d_kps.data_ptr(), d_desc.data_ptr(), d_counts.data_ptr(),
                                d_depth.data_ptr(), DEPTH_FACTOR, BF, d_kd.data_ptr(), d_ku.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        step_device()
    ex.sync()
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.25)
    barrier()
    t0 = time.time()
    ex.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = ex.timer_stop()
    t1 = time.time()
    barrier()
    clk = clocks.stop(t0, t1)
    ms = max_over_ranks(ms)
    launches_per_step = ex.last_launches()
    value = world * B * args.steps / (ms * 1e-3)
    mean_kp = float(d_counts.float().mean().item())

    if args.device_only:
        print(json.dumps({'metric': METRIC, 'value': value, 'unit': 'frames/s', 'ms_per_step': ms / args.steps, 'device_only': True}))
        return

    # ---- per-stage device times (separate, profiled pass; events between stages) ----
    ex.set_profiling(True)
    stage = {k: 0.0 for k in ('pyramid', 'fast', 'octree', 'blur', 'describe')}
    reps = 5
    for _ in range(reps):
        step_device()
        ex.sync()
        for k, v in ex.stage_times().items():
            stage[k] += v / reps
    ex.set_profiling(False)
    alg = {'pyramid': BYTES_PYRAMID, 'fast': BYTES_FAST, 'blur': BYTES_BLUR, 'describe': BYTES_DESCRIBE_PATCH + BYTES_OUT}
    dom = max(('pyramid', 'fast', 'blur', 'describe'), key=lambda k: stage[k])
    peak, peak_src = measured_peak()
    achieved = alg[dom] * B / (stage[dom] * 1e-3) / 1e9
    roofline = dict(bound='hbm', kernel={'pyramid': 'k_resize x7', 'fast': 'k_fast_cells', 'blur': 'k_blur', 'describe': 'k_describe'}[dom],
                    achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak, traffic=None, peak_source=peak_src,
                    algorithmic_bytes_per_launch=alg[dom] * B,
                    stage_ms={k: round(v, 4) for k, v in stage.items()},
                    stage_frac_of_hbm={k: round(alg[k] * B / (stage[k] * 1e-3) / 1e9 / peak, 4) for k in alg})

    # ---- e2e: host (pinned) buffers through the C-ABI call, copies inside the timed region ----
    def pinned(shape, dtype):
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        return t, t.numpy()
    _, h_gray = pinned((B, H, W), torch.uint8)
    _, h_depth_i16 = pinned((B, H, W), torch.int16)
    h_depth = h_depth_i16.view(np.uint16)
    h_gray[:] = gray
    h_depth[:] = depth
    out = dict(counts=pinned((B,), torch.int32)[1], kps=pinned((B, cap, 7), torch.float32)[1].view(hvo.KP_DTYPE).reshape(B, cap),
               desc=pinned((B, cap, 32), torch.uint8)[1], depth=pinned((B, cap), torch.float32)[1],
               uright=pinned((B, cap), torch.float32)[1])
    for _ in range(args.warmup):
        ex.extract_batch(h_gray, depth16=h_depth, depth_factor=DEPTH_FACTOR, bf=BF, out=out)
    barrier()
    ex.timer_start()
    for _ in range(args.steps):
        ex.extract_batch(h_gray, depth16=h_depth, depth_factor=DEPTH_FACTOR, bf=BF, out=out)
    e2e_ms = max_over_ranks(ex.timer_stop())
    barrier()
    e2e = dict(value=world * B * args.steps / (e2e_ms * 1e-3), unit='frames/s',
               h2d_bytes_per_step=int(h_gray.nbytes + h_depth.nbytes),
               d2h_bytes_per_step=int(sum(v.nbytes for v in out.values())), ms_per_step=e2e_ms / args.steps)

    # ---- single-frame latency through operator() (p50) ----
    ex1 = hvo.ORBextractor(ORB['nfeatures'], ORB['scale_factor'], ORB['nlevels'], ORB['ini_th'], ORB['min_th'],
                           width=W, height=H, max_batch=1, device=local_rank)
    lat = []
    for i in range(40):
        t = time.perf_counter()
        ex1(gray[i % B])
        lat.append(1e3 * (time.perf_counter() - t))
    p50 = float(np.median(lat[8:]))
    ex1.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(gray)

    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
            'data': 'synthetic',
            'config': {'workload': f'C1: synthetic TUM-fr3-shaped 640x480 RGB-D frames, ORB 1000 features 8 levels x1.2 (TUM3.yaml), '
                                   f'{B} distinct frames per GPU per step, frame-sharded over {world} GPU(s)',
                       'stages': ['orb: pyramid + per-cell FAST + quadtree + IC_Angle + blur + rBRIEF + RGB-D depth lookup'],
                       'batch_per_gpu': B, 'mean_keypoints_per_frame': mean_kp,
                       'l2': f'inputs larger than L2: per-step working set {B} frames x ~2.2 MB (gray, pyramid, candidates) >> 126 MB'},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches_per_step * args.steps,
            'gpu_launches_per_step': launches_per_step, 'clocks': clk, 'p50_latency_ms_single_frame': p50,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
