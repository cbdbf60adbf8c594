#!/usr/bin/env python
"""Randomised parity sweep of the general tensor-core band-power kernel (csrc/ms_dft_seg.cu) against numpy's fp64 STFT:
random transform sizes (including non powers of two), frame lengths, hops (overlapping, abutting and gapped frames),
windows, band positions and widths (1..80 bins: one to three column groups), one to three files with ragged lengths.
The 1e-4 relative budget is applied strictly to every band value and every exception is COUNTED and classified: the
integer kernels carry the basis with 24 bits relative to its peak, so their error floor is ~2e-8 of the FRAME's
amplitude -- a band that sits in a spectral null more than ~55 dB under the frame's energy (a single bin next to a
strong carrier) can miss 1e-4 of ITSELF.  Those are reported as `outside_in_deep_nulls` (band energy below 1e-5 of
the frame's windowed energy); anything else outside the budget is a failure (exit code 1)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200 import ops                      # noqa: E402
from meteor_scatter_b200.synth import synth_file         # noqa: E402


def ref_energy(x, nfft, hop, w, sig, noi, n_frames):
    # n_frames: BandSpec.n_blocks -- len // hop for abutting / gapped frames (the reference drops the tail block,
    # dsp/src/main.py:356), (len - frame) // hop + 1 for overlapping ones (scipy / mlab framing)
    if n_frames <= 0:
        return np.zeros(0), np.zeros(0), np.zeros(0)
    idx = np.arange(len(w))[None, :] + hop * np.arange(n_frames)[:, None]
    xw = x[idx].astype(np.float64) * w[None, :]
    spec = np.fft.rfft(xw, n=nfft, axis=1)
    p = np.abs(spec) ** 2
    # Parseval: sum over all bins of |X|^2 = nfft * sum (x w)^2 -> the frame's total spectral energy
    return p[:, sig].sum(axis=1), p[:, noi].sum(axis=1), nfft * np.sum(xw * xw, axis=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=150)
    ap.add_argument("--seed", type=int, default=31)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    out = {"cases": args.cases, "frames": 0, "values": 0, "values_outside_1e-4_rel": 0, "outside_in_deep_nulls": 0,
           "outside_other": 0, "max_rel_err": 0.0, "max_rel_err_outside_nulls": 0.0,
           "segment_form": 0, "rows_are_frames_form": 0, "multi_group": 0, "non_pow2": 0, "failures": []}
    for c in range(args.cases):
        nfft = int(rng.choice([256, 500, 512, 1000, 1024, 1536, 2048, 3000, 4096, 8192]))
        frame = nfft if rng.integers(0, 3) else int(rng.integers(nfft // 2, nfft + 1))
        kind = int(rng.integers(0, 4))
        if kind == 0:      # abutting / gapped frames
            hop = int(rng.integers(frame // 8 + 1, frame // 4 + 40)) * 8
            hop = max(hop, (frame + 7) // 8 * 8)
        else:              # overlapping
            hop = max(8, int(round(frame * float(rng.choice([0.5, 0.25, 0.1, 0.37, 0.0625])) / 8)) * 8)
        wk = int(rng.integers(0, 5))
        n = np.arange(frame)
        w = [np.hanning(frame), 0.5 - 0.5 * np.cos(2 * np.pi * n / frame), rng.uniform(0.05, 1.0, frame),
             0.54 - 0.46 * np.cos(2 * np.pi * n / frame),
             0.42 - 0.5 * np.cos(2 * np.pi * n / frame) + 0.08 * np.cos(4 * np.pi * n / frame)][wk]
        divs = [d for d in (2, 4, 8, 16) if nfft % d == 0 and (nfft // d) % 8 == 0]
        if wk in (1, 3, 4) and kind != 0 and divs and rng.integers(0, 2):   # exercise the frequency-domain-window form
            frame = nfft
            n = np.arange(frame)
            w = [None, 0.5 - 0.5 * np.cos(2 * np.pi * n / frame), None, 0.54 - 0.46 * np.cos(2 * np.pi * n / frame),
                 0.42 - 0.5 * np.cos(2 * np.pi * n / frame) + 0.08 * np.cos(4 * np.pi * n / frame)][wk]
            hop = frame // int(rng.choice(divs))
        nb_sig, nb_noi = int(rng.choice([1, 3, 7, 20, 40])), int(rng.choice([0, 1, 4, 14, 40]))
        k0 = int(rng.integers(1, nfft // 2 - nb_sig - nb_noi - 2))
        sig = np.arange(k0, k0 + nb_sig)
        noi = np.arange(k0 + nb_sig + 1, k0 + nb_sig + 1 + nb_noi)
        n_files = int(rng.integers(1, 4))
        dur = float(rng.uniform(2.0, 20.0))
        base = 6000 * dur
        spf = int(base) // 8 * 8 - 8 * int(rng.integers(0, 40))
        files = [synth_file(int(rng.integers(1, 1 << 30)), dur_s=dur + 1, carrier_hz=6000.0 * (k0 + 1) / nfft,
                            rate_per_hour=float(rng.choice([0, 900, 3000])), noise_sigma=float(rng.choice([50, 300, 2000])))[:spf]
                 for _ in range(n_files)]
        spec = ops.BandSpec.stft(nfft, hop, w, sig, noi, fs=6000)
        xd = torch.from_numpy(np.stack(files)).cuda()
        if not ops.seg_supported(xd, spec):
            out["failures"].append(dict(case=c, why="seg_supported is False", nfft=nfft, frame=frame, hop=hop))
            continue
        out["segment_form" if hop < frame else "rows_are_frames_form"] += 1
        out["multi_group"] += int(nb_sig + nb_noi > 32)
        out["non_pow2"] += int(nfft & (nfft - 1) != 0)
        impl = "seg"
        if ops.rot_supported(xd, spec) and rng.integers(0, 4) > 0:
            impl = "rot"
            out["frequency_domain_window_form"] = out.get("frequency_domain_window_form", 0) + 1
        _, _, be, ne = ops.band_power(xd, spec, impl=impl, want_energy=True)
        bdb, ndb = ops.band_power(xd, spec, impl=impl)
        bad_case = 0
        for i, x in enumerate(files):
            eb, en, etot = ref_energy(x, nfft, hop, w, sig, noi, spec.n_blocks(len(x)))
            assert be.shape[1] == len(eb), (be.shape, len(eb))
            for got, ref, gdb in ((be[i], eb, bdb[i]), (ne[i], en, ndb[i])):
                if len(ref) == 0 or (ref is en and nb_noi == 0):
                    continue
                g = got.cpu().numpy().astype(np.float64)
                rel = np.abs(g - ref) / np.maximum(ref, 1e-300)
                # the energy outputs are fp32: allow their rounding on top of the budget
                db_err = np.abs(gdb.cpu().numpy().astype(np.float64) - 10 * np.log10(ref + 1e-12))
                miss = (rel > 1e-4 + 1.2e-7) | (db_err > 10 * np.log10(1 + 1e-4) + 2e-5)
                null = ref < 1e-5 * etot
                out["values"] += len(ref)
                out["values_outside_1e-4_rel"] += int(miss.sum())
                out["outside_in_deep_nulls"] += int((miss & null).sum())
                out["outside_other"] += int((miss & ~null).sum())
                out["max_rel_err"] = max(out["max_rel_err"], float(rel.max()))
                if (~null).any():
                    out["max_rel_err_outside_nulls"] = max(out["max_rel_err_outside_nulls"], float(rel[~null].max()))
                bad_case += int((miss & ~null).sum())
            out["frames"] += len(eb)
        if bad_case:
            out["failures"].append(dict(case=c, impl=impl, nfft=nfft, frame=frame, hop=hop, window=wk, bins=(nb_sig, nb_noi),
                                        n_files=n_files, spf=spf, outside=bad_case))
    print(json.dumps(out))
    return 1 if out["failures"] else 0


if __name__ == "__main__":
    sys.exit(main())
