#!/usr/bin/env python
"""BASELINE configs[3]: FFT size / hop sweep (nfft 1024..16384, 50/75/90 % overlap), periodic Hann window,
signal band = carrier +/- 10 Hz, noise band 690-710 Hz, on 2 h of synthetic 6 kHz PCM16 audio resident in HBM.
For every point: the FFT kernel (K1) and, where the frame fits its basis (nfft = 1024), the tensor-core
restricted-DFT kernel (K2).  Reports frames/s, Msamples/s of *unique* audio, and the algorithmic HBM fraction
(unique input bytes + 8 B per frame)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200 import ops                      # noqa: E402
from meteor_scatter_b200.synth import synth_batch_torch  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    fs = 6000
    n_files, spf = 24, 1_800_000                        # 2 h
    x = synth_batch_torch(n_files, spf, seed=3, device="cuda")
    peak = 6553.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rows = []
    for nfft in (1024, 2048, 4096, 8192, 16384):
        w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(nfft) / nfft)
        freqs = np.fft.rfftfreq(nfft, 1 / fs)
        sig = np.nonzero((freqs >= 993) & (freqs <= 1013))[0]
        noi = np.nonzero((freqs >= 690) & (freqs <= 710))[0]
        if nfft == 1024:
            noi = noi[:max(1, 8 - len(sig))]
        for ov in (0.5, 0.75, 0.9):
            hop = max(8, int(round(nfft * (1 - ov) / 8)) * 8)
            spec = ops.BandSpec.stft(nfft, hop, w, sig, noi, fs=fs)
            nfr = spec.n_blocks(spf)
            unique_bytes = n_files * (spf * 2 + nfr * 8)
            for impl in ("fft", "tc"):
                if impl == "tc" and not ops.tc_supported(x, spec):
                    continue
                ms = timeit(lambda: ops.band_power(x, spec, impl=impl))
                rows.append(dict(nfft=nfft, overlap=ov, hop=hop, impl=impl, frames=n_files * nfr, ms=round(ms, 4),
                                 Mframes_per_s=round(n_files * nfr / ms / 1e3, 1),
                                 Msamples_per_s=round(n_files * spf / ms / 1e3, 0),
                                 hbm_frac=round(unique_bytes / (ms * 1e-3) / 1e9 / peak, 4)))
                print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
