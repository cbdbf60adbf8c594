#!/usr/bin/env python
"""BASELINE configs[2]: a multi-day synthetic archive (288 five-minute files per day) sharded by day
across the ranks of one node, one NCCL sum-reduce of the [hours x 2] histogram to rank 0, day files
``YYYYMMDD.csv`` (Timestamp;Anzahl;Kritisch) written by rank 0.

    python tools/archive_run.py --days 30 [--out DIR]                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/archive_run.py --days 30

Day d belongs to rank d % world.  Each day is one 1.04 GB batch generated on the device, processed through
PassPipeline (detect of day i under the band-power kernel of day i+1) and accumulated into the rank's
archive histogram.  Prints one JSON line with throughput and consistency checks."""
import argparse
import datetime
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200 import csvout                                               # noqa: E402
from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams, PassPipeline, datetime_to_us  # noqa: E402
from meteor_scatter_b200.synth import synth_batch_torch                              # noqa: E402

FS, FILE_S, FILES_PER_DAY = 6000, 300, 288
T0 = datetime.datetime(2025, 6, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--days", type=int, default=30)
    ap.add_argument("--out", default=None)
    ap.add_argument("--distinct", type=int, default=2, help="distinct synthetic days kept in HBM and cycled")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    det = DetectorA(DetectorAParams(), impl="tc")
    spf = FS * FILE_S
    n_hours = args.days * 24
    my_days = list(range(rank, args.days, world))
    pool = [synth_batch_torch(FILES_PER_DAY, spf, seed=4242 + 17 * rank + i, device=dev) for i in range(args.distinct)]
    pipe = PassPipeline(det, FILES_PER_DAY, spf, n_hours, dev, depth=2)
    archive = torch.zeros((n_hours, 2), dtype=torch.int32, device=dev)
    total_events = torch.zeros((), dtype=torch.int64, device=dev)
    us_of_day = lambda d: torch.tensor(                                                # noqa: E731
        [datetime_to_us(T0 + datetime.timedelta(days=d, seconds=FILE_S * i)) for i in range(FILES_PER_DAY)],
        dtype=torch.int64, device=dev)
    starts = {d: us_of_day(d) for d in my_days}
    for _ in range(3):                                   # untimed warm-up: plan upload, lazy module load
        pipe.wait(pipe.submit(pool[0], us_of_day(0), T0))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prev = None
    for j, d in enumerate(my_days):
        slot = pipe.submit(pool[j % args.distinct], starts[d], T0)
        if prev is not None:
            r, h = pipe.wait(prev)
            archive += h
            total_events += r.det.counts.sum()
        prev = slot
    if prev is not None:
        r, h = pipe.wait(prev)
        archive += h
        total_events += r.det.counts.sum()
        r.det.check_capacity()
    if world > 1:
        dist.reduce(archive, dst=0, op=dist.ReduceOp.SUM)      # the one collective of the path
        dist.reduce(total_events, dst=0, op=dist.ReduceOp.SUM)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        hist = archive.cpu().numpy()
        out = args.out or tempfile.mkdtemp(prefix="ms_archive_csv_")
        os.makedirs(out, exist_ok=True)
        files = csvout.write_day_files(out, csvout.hourly_rows(hist, T0), merge=False)
        rows = sum(len(open(f).read().splitlines()) - 1 for f in files)
        samples = args.days * FILES_PER_DAY * spf
        print(json.dumps({
            "config": f"{args.days}-day synthetic archive = {args.days * FILES_PER_DAY} five-minute files, sharded by "
                      f"day over {world} GPU(s), PassPipeline + one NCCL reduce of the [{n_hours} x 2] histogram",
            "n_gpus": world, "ms_total": ms, "Msamples_per_s": samples / (ms * 1e-3) / 1e6,
            "events_total": int(total_events.item()), "anzahl_total": int(hist[:, 0].sum()),
            "kritisch_total": int(hist[:, 1].sum()), "csv_files": len(files), "csv_rows": rows,
            "checks": {"anzahl_equals_events": int(hist[:, 0].sum()) == int(total_events.item()),
                       "one_file_per_day": len(files) == args.days, "one_row_per_hour": rows == n_hours,
                       "kritisch_le_anzahl": bool((hist[:, 1] <= hist[:, 0]).all())},
            "csv_dir": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
