"""Multi-process path on CPU (gloo, world_size 2): frame sharding, max-over-ranks timing and the host gather.
The per-frame work in this test is the CPU oracle (the CUDA product cannot run here); the sharding logic is the
product's (hvo_b200.sharding)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_is_a_partition():
    from hvo_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 8192, 8191):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import hvo_b200  # noqa: F401
    from hvo_b200 import sharding, synth
    import oracle
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n = 5
    lo, hi = sharding.shard_range(n, rank, world)
    rows = []
    for i in range(lo, hi):
        g, _ = synth.frame('S2', 100 + i)
        kps, desc = oracle.OrbOracle(300, 1.2, 4, 20, 7).extract(g[100:340, 100:420])
        rows.append([i, len(kps), int(desc.sum())])
    allrows = sharding.gather_frame_rows(np.array(rows, np.int32).reshape(-1, 3), n, dist)
    t = sharding.max_over_ranks(10.0 + rank, dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, allrows.tolist(), t))


def test_two_rank_gloo_frame_sharding():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import hvo_b200  # noqa: F401
    from hvo_b200 import synth
    import oracle
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = []
    for i in range(5):
        g, _ = synth.frame('S2', 100 + i)
        kps, desc = oracle.OrbOracle(300, 1.2, 4, 20, 7).extract(g[100:340, 100:420])
        expect.append([i, len(kps), int(desc.sum())])
    for rank, rows, t in res:
        assert rows == expect          # every rank sees all frames, in frame order, identical to the unsharded run
        assert t == 11.0               # max over ranks
