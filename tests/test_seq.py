"""hvo_seq_*: offline sequences partitioned across the GPUs of one box from one process (no NCCL).  CPU: the partition.  GPU: the
gathered result equals the single-handle result frame for frame (on one GPU the device list names it twice: two handles, two host
threads, the same code path as two GPUs)."""
import numpy as np
import pytest


def test_shard_is_a_contiguous_partition(hvo):
    for n in (1, 7, 8, 1024, 8192, 8195):
        for g in (1, 2, 3, 4, 8):
            ranges = [hvo.FrameSequence.shard(n, g, d) for d in range(g)]
            assert ranges[0][0] == 0 and sum(c for _, c in ranges) == n
            for (a, ca), (b, _) in zip(ranges, ranges[1:]):
                assert a + ca == b                                          # contiguous, in device order
            assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 1


@pytest.mark.gpu
def test_gpu_sequence_equals_single_handle(hvo, synth):
    c = synth.CONFIGS['S1']
    cam = (c['fx'], c['fy'], c['cx'], c['cy'], 1.0 / c['factor'])
    gray, depth = synth.sequence('S1', 7, start=70)
    ndev = hvo.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0]
    seq = hvo.FrameSequence(640, 480, *cam, devices=devices, frames_per_call=2)   # 2 frames per queued call: several calls per device
    got = seq.extract(gray, depth)
    assert seq.last_ms() > 0
    seq.close()
    fe = hvo.FrameFrontEnd(640, 480, *cam, max_batch=7, max_planes=15, line_cull=True, membership='u4', normals='n3')
    ref = fe.extract_batch(gray, depth)
    fe.close()
    for k in ('kp_counts', 'line_counts', 'n_planes', 'membership4'):
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(np.isnan(got['normals3']), np.isnan(ref['normals3'])) and np.array_equal(np.nan_to_num(got['normals3']), np.nan_to_num(ref['normals3']))
    for f in range(7):
        n, nl, npl = int(ref['kp_counts'][f]), int(ref['line_counts'][f]), int(ref['n_planes'][f])
        assert got['kps'][f, :n].tobytes() == ref['kps'][f, :n].tobytes() and np.array_equal(got['desc'][f, :n], ref['desc'][f, :n])
        assert np.array_equal(got['kp_depth'][f, :n], ref['kp_depth'][f, :n])
        assert got['keylines'][f, :nl].tobytes() == ref['keylines'][f, :nl].tobytes() and np.array_equal(got['line_desc'][f, :nl], ref['line_desc'][f, :nl])
        assert np.array_equal(got['planes7'][f, :npl], ref['planes7'][f, :npl])
