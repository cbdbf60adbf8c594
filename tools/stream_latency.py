#!/usr/bin/env python
"""configs[4]: streaming real-time mode.  Continuous SDR audio arrives in 1 s chunks
(detector B semantics: 4 kHz, 5 blocks of 800 samples per chunk); for every chunk
measure the wall time from "chunk available in host memory" to "its detections are
on the host" (H2D + Welch band kernel + state-machine kernel + D2H), p50/p99."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200.dsp.src.live.backend.aggregates import ConfigDetection  # noqa: E402
from meteor_scatter_b200.dsp.src.live.backend.processor import LiveDetector      # noqa: E402
from meteor_scatter_b200.synth import synth_file                                  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=10000)
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--graph", action="store_true", help="LiveDetector.push_host: the chunk step as one CUDA graph")
    args = ap.parse_args()
    fs, chunk = 4000, 4000
    base = synth_file(5, fs=fs, dur_s=600.0, carrier_hz=1020.0, rate_per_hour=900.0)
    n_base = len(base) // chunk
    cfg = ConfigDetection(proc_block_sec=0.2, n_fft=4096, detection_db_over_noise_mean_min=1,
                          detection_dur_min_sec=0.5, signal_freq=1020)
    det = LiveDetector(cfg, fs=fs, n_streams=args.streams, device="cuda")
    pinned = torch.empty((args.streams, chunk), dtype=torch.int16).pin_memory()
    dev = torch.empty((args.streams, chunk), dtype=torch.int16, device="cuda")
    lat = []
    n_det = 0
    for i in range(args.chunks + 50):
        src = torch.from_numpy(base[(i % n_base) * chunk:(i % n_base + 1) * chunk])
        pinned.copy_(src.unsqueeze(0).expand(args.streams, -1))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.graph:
            new = det.push_host(pinned)          # pinned staging copy + graph replay + stream sync
        else:
            dev.copy_(pinned, non_blocking=True)
            new = det.push(dev)                  # ends with a D2H of the detection counters (synchronises)
        t1 = time.perf_counter()
        n_det += len(new)
        if i >= 50:
            lat.append((t1 - t0) * 1e6)
    lat = np.sort(np.array(lat))
    out = {"mode": "streaming 1 s chunks, detector B (4 kHz, 5 x 800-sample blocks, Welch nfft 4096)",
           "path": "push_host (CUDA graph)" if args.graph else "push", "streams": args.streams, "chunks": args.chunks, "detections": n_det,
           "latency_us": {"p50": float(lat[len(lat) // 2]), "p90": float(lat[int(len(lat) * 0.9)]),
                          "p99": float(lat[int(len(lat) * 0.99)]), "max": float(lat[-1])},
           "realtime_factor_p99": 1e6 / float(lat[int(len(lat) * 0.99)])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
