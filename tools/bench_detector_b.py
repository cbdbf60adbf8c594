#!/usr/bin/env python
"""Detector B batch throughput: n_streams x 10 min of 4 kHz PCM16 through the Welch band kernel and the state
machine (the reference needs 2.3 s per 5-minute file on one core)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200 import ops                                                     # noqa: E402
from meteor_scatter_b200.dsp.src.live.backend.aggregates import ConfigDetection          # noqa: E402
from meteor_scatter_b200.dsp.src.live.backend.processor import band_bins, live_config    # noqa: E402
from meteor_scatter_b200.synth import synth_batch_torch                                  # noqa: E402

n_streams, n = 256, 4000 * 600
impl = sys.argv[1] if len(sys.argv) > 1 else "auto"     # auto | tc | qf | fft
x = synth_batch_torch(n_streams, n, fs=4000, carrier_hz=1020.0, rate_per_hour=600.0, seed=9, device="cuda")
cfg = ConfigDetection(proc_block_sec=0.2, n_fft=4096, detection_db_over_noise_mean_min=1, detection_dur_min_sec=0.5,
                      signal_freq=1020)
_, bands = band_bins(cfg, 4000)
lc = live_config(cfg, 4000, 800)


def run():
    band = ops.welch_band_db(x, 800, 4096, bands, 4000.0, impl=impl)
    st = ops.LiveStates(n_streams, "cuda")
    ops.live_state_step(st, lc, band[:, :, 3])
    return st


for _ in range(2):
    run()
torch.cuda.synchronize()
reps = 10
w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0.record()
for _ in range(reps):      # the Welch stage alone, back to back (1.2 GB of input: larger than L2)
    ops.welch_band_db(x, 800, 4096, bands, 4000.0, impl=impl)
w1.record()
torch.cuda.synchronize()
welch_rep_ms = w0.elapsed_time(w1) / reps
band = ops.welch_band_db(x, 800, 4096, bands, 4000.0, impl=impl)
state_ms = []
for _ in range(5):         # the state stage alone: fresh (pre-allocated) states, events around the one FFI call
    st = ops.LiveStates(n_streams, "cuda")
    torch.cuda.synchronize()
    b, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    ops.live_state_step(st, lc, band[:, :, 3])
    c.record()
    torch.cuda.synchronize()
    state_ms.append(b.elapsed_time(c))
state_ms = sorted(state_ms)[len(state_ms) // 2]
total_ms = welch_rep_ms + state_ms
print(json.dumps({"impl": impl, "welch_ms": welch_rep_ms, "welch_hbm_GBps": n_streams * n * 2 / (welch_rep_ms * 1e-3) / 1e9,
                  "state_ms": state_ms, "streams": n_streams, "samples": n_streams * n,
                  "Msamples_per_s": n_streams * n / (total_ms * 1e-3) / 1e6,
                  "detections": int(st.det_count.sum().item())}))
