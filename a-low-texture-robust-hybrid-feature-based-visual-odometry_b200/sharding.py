"""Frame sharding for offline sequences (BASELINE.json config C5).

Frames are independent units: rank r of `world` processes frames [lo, hi) of the sequence, no data-path collective.
torch.distributed is used only to agree on timing (max over ranks) and to gather per-frame summaries on rank 0."""
import numpy as np


def shard_range(n_frames, rank, world):
    """Contiguous, balanced partition: the first n_frames % world ranks get one extra frame."""
    base, rem = divmod(int(n_frames), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, dist=None, device='cpu'):
    """Device-timed step length of the whole job = the slowest rank's."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_frame_rows(local_rows, n_frames, dist=None, device='cpu'):
    """Host gather in frame order: local_rows is this rank's [hi-lo, k] int32 array; returns [n_frames, k] on every rank."""
    local_rows = np.ascontiguousarray(local_rows, np.int32)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_rows
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    k = local_rows.shape[1]
    sizes = [shard_range(n_frames, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((cap, k), dtype=torch.int32, device=device)
    buf[:len(local_rows)] = torch.from_numpy(local_rows).to(device)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return np.concatenate([outs[r][:hi - lo].cpu().numpy() for r, (lo, hi) in enumerate(sizes)], axis=0)
