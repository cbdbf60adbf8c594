#!/usr/bin/env python
"""Randomised parity sweep of the PSD spectrogram entry point (detector C numeric stage / sweep config) against
scipy.signal.spectrogram in the call form the reference uses (dsp/src/main.py:52-54, prime_detection.py:70): random
nfft (or --nfft 2048: the warp-per-frame kernel only), overlap, row and noise-band ranges, int16 and float32 input.
PSD rows within 1e-4 relative + 1e-8 of the frame peak, noise-band sum within 1e-4; the values outside the strict
1e-4 of themselves are counted separately (``strict_outside`` of ``values``: the fp32 rounding floor of a frame with a
strong line).  One JSON line; exit code 1 on a mismatch."""
import argparse
import json
import os
import sys

import numpy as np
import torch
from scipy.signal import spectrogram

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200 import ops                     # noqa: E402
from meteor_scatter_b200.synth import synth_file         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=23)
    ap.add_argument("--nfft", type=int, default=0, help="0 = random power of two in 256..8192")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    out = {"cases": args.cases, "nfft": args.nfft or "random", "seed": args.seed, "frames": 0, "values": 0,
           "strict_outside": 0, "strict_max_rel": 0.0, "ok": 0, "mismatch": 0, "worst_excess": 0.0, "failures": []}
    for c in range(args.cases):
        fs = int(rng.choice([4000, 5000, 6000]))
        nfft = args.nfft or int(rng.choice([256, 512, 1024, 2048, 4096, 8192]))
        noverlap = int(rng.choice([0, nfft // 2, (3 * nfft) // 4, nfft - nfft // 8]))
        x = synth_file(int(rng.integers(1, 1 << 30)), fs=fs, dur_s=float(rng.uniform(2 * nfft / fs + 0.1, 20.0)),
                       carrier_hz=float(rng.uniform(300, fs / 2 - 300)), rate_per_hour=float(rng.choice([0, 1200, 6000])),
                       noise_sigma=float(rng.choice([30, 300, 3000])))
        as_float = bool(rng.integers(0, 2))
        xin = (x.astype(np.float32) / np.float32(32768.0)) if as_float else x
        k_lo = int(rng.integers(0, nfft // 2 - 8))
        k_hi = int(min(nfft // 2, k_lo + rng.integers(1, 200)))
        n_lo = int(rng.integers(0, nfft // 2 - 8))
        n_hi = int(min(nfft // 2, n_lo + rng.integers(1, 300)))
        w = np.hanning(nfft)
        f, t, ref = spectrogram(xin.astype(np.float64), fs, window=w, nperseg=nfft, noverlap=noverlap, detrend=False,
                                scaling="density", mode="psd")
        psd, noise = ops.psd_spectrogram(torch.from_numpy(np.ascontiguousarray(xin)).cuda(), float(fs), nfft, noverlap, w,
                                         k_lo, k_hi, n_lo, n_hi)
        got = psd[0].cpu().numpy().astype(np.float64)
        want = ref[k_lo:k_hi + 1]
        peak = ref.max(axis=0, keepdims=True)
        excess = float(np.max(np.abs(got - want) / (1e-4 * want + 1e-8 * peak + 1e-300)))
        nsum = float(ref[n_lo:n_hi + 1].sum())
        nerr = abs(float(noise[0].item()) - nsum) / nsum / 1e-4
        rel = np.abs(got - want) / np.maximum(want, 1e-300)
        out["values"] += int(want.size)
        out["strict_outside"] += int((rel > 1e-4).sum())
        out["strict_max_rel"] = max(out["strict_max_rel"], float(rel.max()))
        out["frames"] += got.shape[1]
        out["worst_excess"] = max(out["worst_excess"], excess, nerr)
        if got.shape == want.shape and excess <= 1.0 and nerr <= 1.0:
            out["ok"] += 1
        else:
            out["mismatch"] += 1
            out["failures"].append(dict(case=c, fs=fs, nfft=nfft, noverlap=noverlap, float=as_float, excess=excess,
                                        noise_excess=nerr, shape=[list(got.shape), list(want.shape)]))
    out["failures"] = out["failures"][:6]
    print(json.dumps(out))
    return 1 if out["mismatch"] else 0


if __name__ == "__main__":
    sys.exit(main())
