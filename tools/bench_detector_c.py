#!/usr/bin/env python
"""Detector C numeric stage (C-stft, prime_detection.py:67-92) in batch: n_seg 30-second segments of 5 kHz PCM16 ->
one-sided PSD rows 800-1200 Hz [164 x 145] per segment + the 250-800 Hz noise-band sum (specgram NFFT 2048, noverlap
1024, np.hanning).  Reports time, segments/s and the fraction of the HBM roofline (input + output bytes)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200 import ops                               # noqa: E402
from meteor_scatter_b200.synth import synth_batch_torch           # noqa: E402

n_seg, n, fs, nfft = 2048, 150_000, 5000, 2048
x = synth_batch_torch(n_seg, n, fs=fs, carrier_hz=1000.0, rate_per_hour=600.0, seed=3, device="cuda")
freqs = np.fft.rfftfreq(nfft, 1 / fs)
rows = np.nonzero((freqs >= 800) & (freqs <= 1200))[0]
nk = np.nonzero((freqs >= 250) & (freqs <= 800))[0]
w = np.hanning(nfft)


def run():
    return ops.psd_spectrogram(x, float(fs), nfft, nfft // 2, w, int(rows[0]), int(rows[-1]), int(nk[0]), int(nk[-1]))


for _ in range(3):
    psd, noise = run()
torch.cuda.synchronize()
reps = 10
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    psd, noise = run()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
bytes_algo = x.numel() * 2 + psd.numel() * 4
peak = 6553.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
print(json.dumps({"kernel": "stft_kernel (K1)" if os.environ.get("MS_PSD_IMPL") == "fft" else "psd_warp_kernel (K1W)",
                  "segments": n_seg, "samples": int(x.numel()), "frames_per_segment": int(psd.shape[2]),
                  "rows_per_frame": int(psd.shape[1]), "ms": ms, "segments_per_s": n_seg / (ms * 1e-3),
                  "Msamples_per_s": x.numel() / (ms * 1e-3) / 1e6, "algorithmic_bytes": int(bytes_algo),
                  "hbm_frac": bytes_algo / (ms * 1e-3) / 1e9 / peak}))
