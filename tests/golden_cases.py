"""Shared description of the golden fixtures (inputs are re-synthesised from seeds)."""
import datetime
import hashlib
import os

import numpy as np

from meteor_scatter_b200.synth import synth_file

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
START = datetime.datetime(2025, 6, 25, 7, 51, 41)
MB = dict(block_duration_sec=0.2, freq_band=(993, 1013), noise_band=(690, 710), n_fft=512,
          threshold_std_factor=4)
TL = dict(block_duration_sec=0.2, freq_band=(996, 1016), noise_band=(940, 960), n_fft=512,
          threshold_std_factor=3.5)

# name -> (seed, dur_s, params, adaptive, slice(start_sec, end_sec) | None, synth kwargs)
A_CASES = {
    "a_mb_s1": (1, 300.0, MB, True, None, {}),
    "a_mb_s2": (2, 300.0, MB, True, None, {}),
    "a_mb_s3_busy": (3, 300.0, MB, True, None, dict(rate_per_hour=300.0)),
    "a_mb_s4_long": (4, 1500.0, MB, True, None, dict(rate_per_hour=240.0)),
    "a_mb_s5_global": (5, 300.0, MB, False, None, dict(rate_per_hour=150.0)),
    "a_tl_s6": (6, 300.0, TL, True, None, dict(carrier_hz=1006.0, rate_per_hour=200.0)),
    "a_mb_s7_slice": (7, 300.0, MB, True, (30, 200.5), dict(rate_per_hour=250.0)),
    "a_mb_s8_f32": (8, 120.0, MB, True, None, dict(rate_per_hour=300.0, dtype=np.float32)),
    "a_mb_s9_ragged": (9, 61.37, MB, True, None, dict(rate_per_hour=400.0)),
}

B_CASES = {
    "b_live1_s11": (11, 240.0, dict(proc_block_sec=0.20, n_fft=4096, detection_db_over_noise_mean_min=1,
                                    detection_dur_min_sec=0.5, signal_freq=1020), dict(carrier_hz=1020.0)),
    "b_live2_s12": (12, 240.0, dict(proc_block_sec=0.20, n_fft=4096, detection_db_over_noise_mean_min=1,
                                    detection_dur_min_sec=0.5, signal_freq=1025), dict(carrier_hz=1025.0)),
    "b_default_s13": (13, 120.0, dict(), dict(carrier_hz=1000.0)),
}


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def a_input(name):
    """Re-synthesise the WAV samples of an A case, checking the stored sha256."""
    seed, dur, params, adaptive, sl, skw = A_CASES[name]
    x = synth_file(seed, fs=6000, dur_s=dur, **skw)
    g = load(name)
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() == str(g["input_sha256"]), \
        "synthetic generator no longer reproduces the golden input"
    return x, g


def a_sliced(name):
    """Input after the reference's wav_start_sec/wav_end_sec slicing (main.py:251-265)."""
    x, g = a_input(name)
    sl = A_CASES[name][4]
    if sl is not None:
        x = x[int(sl[0] * 6000):int(sl[1] * 6000)]
    return x, g


def b_input(name):
    seed, dur, cfg, skw = B_CASES[name]
    x = synth_file(seed, fs=4000, dur_s=dur, rate_per_hour=900.0, **skw)
    g = load(name)
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() == str(g["input_sha256"])
    return x, g
