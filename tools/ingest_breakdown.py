#!/usr/bin/env python
"""Where the time of the WAV-files -> day-CSV path goes (one B200, page cache warm): header parse, threaded readinto
into pinned rows, H2D, kernels, result unpacking.  Diagnostic for batch.process_files."""
import datetime
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200 import batch                         # noqa: E402
from meteor_scatter_b200.pipeline import DetectorA            # noqa: E402
from meteor_scatter_b200.synth import synth_file              # noqa: E402
from meteor_scatter_b200.wavio import wav_info, write_wav_pcm16  # noqa: E402


def main():
    n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 288
    root = tempfile.mkdtemp(prefix="ms_ingest_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    t0 = datetime.datetime(2025, 6, 25)
    base = [synth_file(900 + i, dur_s=300.0) for i in range(8)]
    paths = []
    for i in range(n_files):
        t = t0 + datetime.timedelta(seconds=300 * i)
        p = os.path.join(root, f"expoFull_gqrx_{t.strftime('%Y%m%d_%H%M%S')}_49969000.wav")
        write_wav_pcm16(p, 6000, base[i % 8])
        paths.append(p)
    out = {"files": n_files, "cpus": len(os.sched_getaffinity(0)), "where": root}
    tic = time.perf_counter
    t = tic(); infos = [wav_info(p) for p in paths]; out["wav_info_all_s"] = tic() - t
    batch.stage_files(paths[:8])
    for th in (8, 16, 32, 64):
        t = tic(); host, lens = batch.stage_files(paths, io_threads=th); dt = tic() - t
        out[f"stage_files_{th}thr_s"] = round(dt, 4)
        out[f"stage_files_{th}thr_GBs"] = round(host.numel() * 2 / dt / 1e9, 2)
    # raw readinto scaling without the staging function (one pre-allocated pinned buffer)
    hv = host.numpy()
    from concurrent.futures import ThreadPoolExecutor
    from meteor_scatter_b200.wavio import read_wav_into
    for th in (8, 16, 32):
        with ThreadPoolExecutor(th) as pool:
            t = tic(); list(pool.map(lambda i: read_wav_into(paths[i], infos[i], hv[i]), range(n_files))); dt = tic() - t
        out[f"readinto_only_{th}thr_GBs"] = round(host.numel() * 2 / dt / 1e9, 2)
    torch.cuda.synchronize()
    t = tic(); x = host.to("cuda", non_blocking=True); torch.cuda.synchronize(); out["h2d_s"] = tic() - t
    det = DetectorA()
    det.run(x); torch.cuda.synchronize()
    t = tic(); res = det.run(x); torch.cuda.synchronize(); out["kernels_s"] = tic() - t
    t = tic(); d = [res.detections(j, t0) for j in range(n_files)]; out["detections_s"] = tic() - t
    for th in (8, 32):
        torch.cuda.synchronize()
        t = tic(); batch.process_files(paths, csv_folder=None, chunk_files=48, io_threads=th); torch.cuda.synchronize()
        out[f"process_files_{th}thr_s"] = round(tic() - t, 4)
    shutil.rmtree(root, ignore_errors=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
