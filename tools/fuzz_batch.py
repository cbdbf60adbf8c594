#!/usr/bin/env python
"""Randomised parity sweep of the BATCH path of detector A against the oracle (CPU, checker only): several recordings
of different lengths in one padded batch with per-file block counts, random recording start times (fractions of a
second, hour and day boundaries, before the Unix epoch), fused hourly [Anzahl, Kritisch] histogram.
Per file the event list must equal the oracle's, and the histogram must equal the oracle's hour bucketing of all
files.  One JSON line; exit code 1 on any mismatch."""
import argparse
import datetime
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200.batch import hour_span                                                # noqa: E402
from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams, datetime_to_us, hour_index  # noqa: E402
from meteor_scatter_b200.synth import synth_file                                                # noqa: E402
from oracle import detector_a as oa                                                             # noqa: E402  (checker only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=40)
    ap.add_argument("--seed", type=int, default=17)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    out = {"batches": args.batches, "files": 0, "events": 0, "identical_batches": 0, "mismatch": 0, "failures": []}
    p = DetectorAParams()
    for b in range(args.batches):
        n_files = int(rng.integers(1, 10))
        base = datetime.datetime(int(rng.choice([1969, 1970, 2025])), int(rng.integers(1, 13)), int(rng.integers(1, 28)),
                                 int(rng.integers(0, 24)), int(rng.integers(0, 60)), int(rng.integers(0, 60)),
                                 int(rng.integers(0, 1000000)))
        xs, starts = [], []
        t = base
        for f in range(n_files):
            dur = float(rng.choice([0.1, 5.0, 60.0, 123.4, 300.0]))
            n = int(6000 * dur) - int(rng.integers(0, 700)) * int(rng.integers(0, 2))
            x = synth_file(int(rng.integers(1, 1 << 30)), dur_s=max(dur, 0.2), rate_per_hour=float(rng.choice([0, 600, 2400])))
            xs.append(x[:max(n, 0)])
            starts.append(t)
            t = t + datetime.timedelta(seconds=float(rng.choice([300, 3599.75, 86400.5, 12.000001])))
        lens = np.array([len(x) for x in xs], dtype=np.int64)
        max_len = int(max(lens.max(), 8))
        max_len += (-max_len) % 8
        host = np.zeros((n_files, max_len), dtype=np.int16)
        for i, x in enumerate(xs):
            host[i, :len(x)] = x
        hour0, n_hours = hour_span(starts, [n / 6000 for n in lens])
        det = DetectorA(p, impl="auto", max_events=2048)
        dev = torch.device("cuda")
        hist = torch.zeros((n_hours, 2), dtype=torch.int32, device=dev)
        nbpf = torch.from_numpy((lens // det.spec.block_size).astype(np.int32)).to(dev)
        us = torch.tensor([datetime_to_us(s) for s in starts], dtype=torch.int64, device=dev)
        res = det.run(torch.from_numpy(host).to(dev), n_blocks_per_file=nbpf,
                      hourly=dict(file_start_us=us, hour0=hour_index(hour0), n_hours=n_hours, out=hist))
        ok = True
        ref_hist = {}
        for i, x in enumerate(xs):
            if len(x) >= det.spec.block_size:
                r = oa.detect_wav(x, 6000, p.block_duration_sec, p.freq_band, p.noise_band, p.n_fft,
                                  p.threshold_std_factor, wav_start_date_time=starts[i])
                want = r["pairs"]
                for h, cnt in oa.hourly_counts(r["detections"]).items():
                    a = ref_hist.setdefault(h, [0, 0])
                    a[0] += cnt[0]
                    a[1] += cnt[1]
            else:
                want = []
            got = res.pairs(i)
            out["events"] += len(want)
            if got != want:
                ok = False
                out["failures"].append(dict(batch=b, file=i, n=int(lens[i]), got=len(got), ref=len(want)))
        hh = hist.cpu().numpy()
        for k in range(n_hours):
            want_row = ref_hist.pop(hour0 + datetime.timedelta(hours=k), [0, 0])
            if list(hh[k]) != want_row:
                ok = False
                out["failures"].append(dict(batch=b, hour=k, got=[int(v) for v in hh[k]], ref=want_row))
        if ref_hist:
            ok = False
            out["failures"].append(dict(batch=b, why="oracle hours outside the histogram span", n=len(ref_hist)))
        out["files"] += n_files
        out["identical_batches" if ok else "mismatch"] += 1
    out["failures"] = out["failures"][:8]
    print(json.dumps(out))
    return 1 if out["mismatch"] else 0


if __name__ == "__main__":
    sys.exit(main())
