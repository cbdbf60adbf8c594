#!/usr/bin/env python
"""End to end from WAV FILES: parse + threaded read into pinned memory + H2D + kernels + D2H + day CSVs
("end-to-end (WAV parse + H2D + kernels + D2H + CSV) reported separately").
Writes --files synthetic 5-minute PCM16 recordings (page cache warm after writing: this measures the software path,
not the disk), then times batch.process_files for several reader-thread counts."""
import argparse
import datetime
import json
import os
import shutil
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200.batch import process_files          # noqa: E402
from meteor_scatter_b200.synth import synth_file             # noqa: E402
from meteor_scatter_b200.wavio import write_wav_pcm16        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=288)
    ap.add_argument("--chunk-files", type=int, default=24)
    args = ap.parse_args()
    root = tempfile.mkdtemp(prefix="ms_ingest_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    t0 = datetime.datetime(2025, 6, 25, 0, 0, 0)
    base = [synth_file(900 + i, dur_s=300.0, rate_per_hour=120.0) for i in range(8)]
    paths = []
    for i in range(args.files):
        t = t0 + datetime.timedelta(seconds=300 * i)
        p = os.path.join(root, f"expoFull_gqrx_{t.strftime('%Y%m%d_%H%M%S')}_49969000.wav")
        write_wav_pcm16(p, 6000, base[i % len(base)])
        paths.append(p)
    os.makedirs(os.path.join(root, "csv"))
    samples = args.files * len(base[0])
    process_files(paths[:8], csv_folder=None)                  # warm-up: library load, plans, allocator
    out = {"files": args.files, "samples": samples, "chunk_files": args.chunk_files, "host_cpus": os.cpu_count(), "runs": []}
    for threads in (1, 4, 8, 16, 32):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = process_files(paths, csv_folder=os.path.join(root, "csv"), chunk_files=args.chunk_files, io_threads=threads)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        out["runs"].append({"io_threads": threads, "seconds": round(dt, 4), "Msamples_per_s": round(samples / dt / 1e6, 1),
                            "events": int(r["hist"][:, 0].sum()), "csv_files": len(r["csv_files"])})
    shutil.rmtree(root, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
