#!/usr/bin/env python
"""Randomised parity sweep of the live detector's threshold history + state machine (B-state) on the GPU against the
oracle (CPU, checker only): random db2 series (quiet, bursty, drifting, with NaN-producing single-value histories),
random detection parameters and random call sizes (>= 32 blocks: event-jumping batch kernel; < 32: per-block kernel;
mixed).  Thresholds must be bit-identical, detections identical in time, dB statistics within 1e-9.
Prints one JSON line; exit code 1 on any mismatch."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200 import _lib, ops       # noqa: E402
from oracle import detector_b as ob             # noqa: E402  (checker only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=150)
    ap.add_argument("--seed", type=int, default=11)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    out = {"cases": args.cases, "blocks": 0, "detections": 0, "identical": 0, "mismatch": 0, "failures": []}
    for c in range(args.cases):
        n = int(rng.integers(1, 2500))
        bsec = float(rng.choice([0.1, 0.2, 0.5]))
        fs = 4000
        block = int(bsec * fs)
        cfg = ob.ConfigDetection(proc_block_sec=bsec, avg_win_sec=float(rng.choice([bsec, 1, 4, 8, 20])),   # >= one block: avg_win 1..200
                                 init_detection_wait_sec=float(rng.choice([0, 1, 8, 30])),
                                 after_tracking_wait_sec=float(rng.choice([0, 2, 12])),
                                 threshold_std_factor=float(rng.choice([1.5, 3, 4, 6])),
                                 detection_db_over_noise_mean_min=float(rng.choice([-1, 1, 5])),
                                 detection_dur_min_sec=float(rng.choice([-1, 0.5, 2])))
        db2 = rng.normal(0, float(rng.choice([0.2, 1, 3])), size=n) + float(rng.uniform(-5, 5))
        db2 += np.cumsum(rng.normal(0, 0.02, size=n))                         # slow drift
        for _ in range(int(rng.integers(0, 12))):
            a = int(rng.integers(0, n))
            db2[a:a + int(rng.integers(1, 80))] += rng.uniform(3, 30)
        db2 = db2.astype(np.float32)
        dets_ref, thr_ref = ob.live_state_machine(db2.astype(np.float64), cfg, fs, block)
        lc = _lib.LiveConfig(block_samples=block, fs=float(fs), k_std=cfg.threshold_std_factor,
                             init_wait_sec=cfg.init_detection_wait_sec, after_wait_sec=cfg.after_tracking_wait_sec,
                             mean_min_db=cfg.detection_db_over_noise_mean_min, dur_min_sec=cfg.detection_dur_min_sec,
                             avg_win=int(cfg.avg_win_sec / cfg.proc_block_sec))
        st = ops.LiveStates(1, "cuda")
        d = torch.from_numpy(db2).cuda().reshape(1, -1)
        mode = int(rng.integers(0, 3))
        parts, i = [], 0
        while i < n:
            size = n if mode == 0 else int(rng.integers(32, 400)) if mode == 1 else int(rng.integers(1, 90))
            parts.append(ops.live_state_step(st, lc, d[:, i:i + size].contiguous(), want_thresholds=True))
            i += size
        thr = torch.cat(parts, dim=1).cpu().numpy()[0]
        k = int(st.det_count[0].item())
        got = st.det[0, :k].cpu().numpy()
        ref = np.array([[m.time_start, m.time_stop, m.duration, m.db_min, m.db_max, m.db_mean, m.db_std]
                        for m in dets_ref]).reshape(-1, 7)
        ok = (np.array_equal(thr, np.asarray(thr_ref, dtype=np.float64), equal_nan=True) and k == len(dets_ref)
              and np.array_equal(got[:, :5], ref[:, :5]) and np.allclose(got[:, 5:], ref[:, 5:], rtol=0, atol=1e-9))
        out["blocks"] += n
        out["detections"] += len(dets_ref)
        if ok:
            out["identical"] += 1
        else:
            out["mismatch"] += 1
            out["failures"].append(dict(case=c, n=n, mode=mode, got=k, ref=len(dets_ref),
                                        thr_equal=bool(np.array_equal(thr, np.asarray(thr_ref), equal_nan=True))))
    print(json.dumps(out))
    return 1 if out["mismatch"] else 0


if __name__ == "__main__":
    sys.exit(main())
