"""CPU, world_size 2 over gloo: the N>1 path's host logic -- round-robin file
sharding + the single sum-reduce of hourly histograms (batch.reduce_hist) --
gives the same merged [hours x 2] counts as the unsharded oracle."""
import datetime
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meteor_scatter_b200.batch import reduce_hist, shard_indices
from meteor_scatter_b200.pipeline import hour_index

N_FILES, N_HOURS = 9, 2
T0 = datetime.datetime(2025, 6, 1, 23, 20, 0)


def _events_for_file(i):
    """Deterministic pseudo detections of file i as (start_block, stop_block)."""
    rng = np.random.default_rng(i)
    s = np.sort(rng.choice(1400, size=int(rng.integers(0, 9)), replace=False))
    return [(int(a), int(a) + int(rng.integers(1, 8))) for a in s]


def _local_hist(indices):
    from oracle import detector_a as oa
    hour0 = T0.replace(minute=0, second=0)
    hist = torch.zeros((N_HOURS, 2), dtype=torch.int32)
    for i in indices:
        start = T0 + datetime.timedelta(seconds=300 * i)
        dets = [oa.OutputDetection(a * 0.2, b * 0.2, b * 0.2 - a * 0.2, 0.0, start + datetime.timedelta(seconds=a * 0.2))
                for a, b in _events_for_file(i)]
        for h, c in oa.hourly_counts(dets).items():
            k = hour_index(h) - hour_index(hour0)
            hist[k, 0] += c[0]
            hist[k, 1] += c[1]
    return hist


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hist = _local_hist(shard_indices(N_FILES, rank, world))
        reduce_hist(hist)
        if rank == 0:
            np.save(out_path, hist.numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_histogram_equals_unsharded(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "hist.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    merged = np.load(out)
    full = _local_hist(range(N_FILES)).numpy()
    assert np.array_equal(merged, full)
    assert merged[:, 0].sum() == sum(len(_events_for_file(i)) for i in range(N_FILES)) > 0
