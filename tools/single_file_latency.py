#!/usr/bin/env python
"""Latency of the drop-in calls on ONE recording, as a user of the reference would run them:
proc_wav_file (detector A, 5 min @ 6 kHz, mb_files parameters) and wav_file_process (detector B, 5 min @ 4 kHz):
WAV read + H2D + kernels + D2H + label/CSV files.  The reference needs ~0.13 s (A) and ~2.3 s (B) per file on one core
(SURVEY.md section 6)."""
import contextlib
import datetime
import io
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200.dsp.src.main import proc_wav_file                                   # noqa: E402
from meteor_scatter_b200.dsp.src.live.backend import processor                               # noqa: E402
from meteor_scatter_b200.dsp.src.live.backend.aggregates import (ConfigDetection, ConfigSpecExport,   # noqa: E402
                                                                 ConfigVisualization)
from meteor_scatter_b200.synth import synth_file                                              # noqa: E402
from meteor_scatter_b200.wavio import write_wav_pcm16                                         # noqa: E402

root = tempfile.mkdtemp(prefix="ms_single_")
pa = os.path.join(root, "expoFull_gqrx_20250625_120000_49969000.wav")
pb = os.path.join(root, "live.wav")
write_wav_pcm16(pa, 6000, synth_file(1, dur_s=300.0, rate_per_hour=240.0))
write_wav_pcm16(pb, 4000, synth_file(2, fs=4000, dur_s=300.0, carrier_hz=1020.0, rate_per_hour=600.0))


def run_a():
    proc_wav_file(pa, 0.2, (993, 1013), (690, 710), 512, 4, out_audacity_lbl_file=os.path.join(root, "lbl.txt"),
                  out_csv_file=os.path.join(root, "out.csv"), wav_start_date_time=datetime.datetime(2025, 6, 25, 12, 0, 0),
                  disable_show_and_write=True, quiet=True)


def run_b():
    cfg = ConfigDetection(proc_block_sec=0.2, n_fft=4096, detection_db_over_noise_mean_min=1, detection_dur_min_sec=0.5,
                          signal_freq=1020)
    with contextlib.redirect_stdout(io.StringIO()):
        processor.wav_file_process(pb, cfg, ConfigVisualization(enable_ui_plots=False), ConfigSpecExport(output_dir=""))


out = {}
for name, fn in (("proc_wav_file_ms", run_a), ("wav_file_process_ms", run_b)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    ts.sort()
    out[name] = {"p50": round(ts[len(ts) // 2], 3), "min": round(ts[0], 3), "max": round(ts[-1], 3)}
print(json.dumps(out))
