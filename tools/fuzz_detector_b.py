#!/usr/bin/env python
"""Randomised end-to-end parity sweep of detector B (Welch bands on the GPU + state machine) against the oracle (CPU,
checker only; scipy.signal.welch per block as processor.py:206 does): random signal frequency, channel width, noise
channel offset, n_fft, block length and detection parameters on synthetic 4 kHz PCM16 audio.
Band levels must agree within the 1e-4 energy budget; detections must be identical unless some block of the oracle sits
within 1e-3 dB of its threshold (reported separately).  One JSON line; exit code 1 on a real mismatch."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200.dsp.src.live.backend.aggregates import ConfigDetection     # noqa: E402
from meteor_scatter_b200.dsp.src.live.backend.processor import LiveDetector         # noqa: E402
from meteor_scatter_b200.synth import synth_file                                     # noqa: E402
from oracle import detector_b as ob                                                  # noqa: E402  (checker only)

DB_TOL = 10 * np.log10(1 + 1e-4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=5)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    out = {"cases": args.cases, "blocks": 0, "detections": 0, "identical": 0, "near_threshold_only": 0, "mismatch": 0,
           "max_band_db_err": 0.0, "failures": []}
    for c in range(args.cases):
        kw = dict(proc_block_sec=float(rng.choice([0.2, 0.25])), n_fft=int(rng.choice([2048, 4096, 8192])),
                  signal_freq=int(rng.choice([800, 1020, 1200, 1500])), channel_width=int(rng.choice([50, 100, 150])),
                  noise_channel_offset=int(rng.choice([200, 300])), avg_win_sec=float(rng.choice([4, 8])),
                  init_detection_wait_sec=float(rng.choice([2, 8])), after_tracking_wait_sec=float(rng.choice([4, 12])),
                  threshold_std_factor=float(rng.choice([3, 4, 5])),
                  detection_db_over_noise_mean_min=float(rng.choice([-1, 1])),
                  detection_dur_min_sec=float(rng.choice([-1, 0.5])))
        dur = float(rng.uniform(20, 90))
        x = synth_file(int(rng.integers(1, 1 << 30)), fs=4000, dur_s=dur, carrier_hz=kw["signal_freq"] + float(rng.uniform(-5, 5)),
                       rate_per_hour=float(rng.choice([300, 900, 2400])), noise_sigma=float(rng.choice([100, 300, 1500])))
        ref = ob.process(x.astype(np.float64) / 32768.0, 4000, ob.ConfigDetection(**kw))
        det = LiveDetector(ConfigDetection(**kw), fs=4000, n_streams=1, device="cuda")
        block = det.block
        nb = len(x) // block
        new, band, thr = det.push(torch.from_numpy(np.ascontiguousarray(x[:nb * block])).cuda(), want_series=True)
        band = band.cpu().numpy()[0].astype(np.float64)
        err = float(max(np.max(np.abs(band[:, 0] - ref["ms_db"])), np.max(np.abs(band[:, 1] - ref["n1_db"])),
                        np.max(np.abs(band[:, 2] - ref["n2_db"]))))
        out["max_band_db_err"] = max(out["max_band_db_err"], err)
        out["blocks"] += nb
        out["detections"] += len(ref["detections"])
        got = [(m.time_start, m.time_stop) for _, m in new]
        want = [(m.time_start, m.time_stop) for m in ref["detections"]]
        with np.errstate(invalid="ignore"):
            near = bool(np.any(np.abs(ref["db2"] - ref["thresholds"]) < 1e-3))
        if got == want and err <= DB_TOL + 1e-5:
            out["identical"] += 1
        elif near and err <= DB_TOL + 1e-5:
            out["near_threshold_only"] += 1
        else:
            out["mismatch"] += 1
            out["failures"].append(dict(case=c, cfg=kw, band_db_err=err, got=len(got), ref=len(want)))
    print(json.dumps(out))
    return 1 if out["mismatch"] else 0


if __name__ == "__main__":
    sys.exit(main())
