#!/usr/bin/env python
"""One point (or a comma-separated list nfft:overlap:impl,...) of configs[3] at 24 h scale with one implementation:
the command profiled for the general tensor-core kernel.   python tools/sweep_point.py 2048 0.5 seg [n_iter]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200 import ops                      # noqa: E402
from meteor_scatter_b200.synth import synth_batch_torch  # noqa: E402


def main():
    fs, n_files, spf = 6000, 288, 1_800_000
    x = synth_batch_torch(n_files, spf, seed=3, device="cuda")
    if ":" in sys.argv[1]:          # several points in one process: nfft:overlap:impl,nfft:overlap:impl,...
        for item in sys.argv[1].split(","):
            nfft, ov, impl = item.split(":")
            one(x, int(nfft), float(ov), impl, 5)
        return
    one(x, int(sys.argv[1]), float(sys.argv[2]), sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 5)


def one(x, nfft, ov, impl, n_iter):
    fs = 6000
    n_files, spf = x.shape
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(nfft) / nfft)
    freqs = np.fft.rfftfreq(nfft, 1 / fs)
    sig = np.nonzero((freqs >= 993) & (freqs <= 1013))[0]
    noi = np.nonzero((freqs >= 690) & (freqs <= 710))[0]
    hop = max(8, int(round(nfft * (1 - ov) / 8)) * 8)
    spec = ops.BandSpec.stft(nfft, hop, w, sig, noi, fs=fs)
    for _ in range(2):
        ops.band_power(x, spec, impl=impl)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n_iter):
        ops.band_power(x, spec, impl=impl)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n_iter
    nfr = spec.n_blocks(spf)
    print(f"nfft {nfft} overlap {ov} hop {hop} impl {impl}: {ms:.4f} ms, "
          f"hbm_frac(unique) {n_files * (spf * 2 + nfr * 8) / (ms * 1e-3) / 1e9 / 6553:.4f}")


if __name__ == "__main__":
    main()
