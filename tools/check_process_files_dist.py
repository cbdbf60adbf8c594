#!/usr/bin/env python
"""Functional check of the N>1 product path on real GPUs (run under torchrun):
batch.process_files shards WAV files round-robin over the ranks, every rank runs detector A on its share,
ONE NCCL sum-reduce merges the hourly histograms, rank 0 writes the day CSVs.  Rank 0 then recomputes
everything with the oracle and compares events per file and hourly counts."""
import datetime
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200.batch import process_files          # noqa: E402
from meteor_scatter_b200.synth import synth_file             # noqa: E402
from meteor_scatter_b200.wavio import write_wav_pcm16        # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    root = os.path.join(tempfile.gettempdir(), "ms_dist_check")
    t0 = datetime.datetime(2025, 6, 25, 22, 50, 0)
    durs = [300.0, 300.0, 150.5, 300.0, 300.0, 90.0, 300.0]
    paths = [os.path.join(root, f"expoFull_gqrx_{(t0 + datetime.timedelta(seconds=300 * i)).strftime('%Y%m%d_%H%M%S')}"
                                f"_49969000.wav") for i in range(len(durs))]
    if rank == 0:
        os.makedirs(os.path.join(root, "csv"), exist_ok=True)
        for i, (p, d) in enumerate(zip(paths, durs)):
            write_wav_pcm16(p, 6000, synth_file(300 + i, dur_s=d, rate_per_hour=300.0))
    dist.barrier()
    out = process_files(paths, csv_folder=os.path.join(root, "csv"))
    mine = sorted(out["detections"])
    assert mine == list(range(rank, len(paths), world)), (rank, mine)
    got_pairs = {i: [(d.t_start, d.t_stop) for d in out["detections"][i]] for i in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, got_pairs)
    if rank == 0:
        from oracle import detector_a as oa            # checker only
        ref_hist = {}
        merged = {}
        for g in gathered:
            merged.update(g)
        total = 0
        for i, d in enumerate(durs):
            x = synth_file(300 + i, dur_s=d, rate_per_hour=300.0)
            start = t0 + datetime.timedelta(seconds=300 * i)
            r = oa.detect_wav(x, 6000, 0.2, (993, 1013), (690, 710), 512, 4, wav_start_date_time=start)
            assert merged[i] == [(dd.t_start, dd.t_stop) for dd in r["detections"]], f"file {i} differs"
            total += len(r["detections"])
            for h, c in oa.hourly_counts(r["detections"]).items():
                a = ref_hist.setdefault(h, [0, 0]); a[0] += c[0]; a[1] += c[1]
        hist, hour0 = out["hist"], out["hour0"]
        for k in range(out["n_hours"]):
            assert list(hist[k]) == ref_hist.get(hour0 + datetime.timedelta(hours=k), [0, 0]), k
        assert int(hist[:, 0].sum()) == total > 0
        rows = sum(len(open(f).read().splitlines()) - 1 for f in out["csv_files"])
        print(f"OK world={world}: {len(paths)} files, {total} events, {len(out['csv_files'])} day files, {rows} hourly rows")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
