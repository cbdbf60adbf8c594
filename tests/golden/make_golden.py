"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference)
on seeded synthetic audio.  Run in the build container only:

    python tests/golden/make_golden.py

Outputs ``tests/golden/*.npz`` (committed).  Inputs are not stored: they are
re-synthesised from the seed by ``meteor_scatter_b200.synth`` and guarded by a
sha256 of the PCM bytes kept in each fixture.
"""
from __future__ import annotations

import datetime
import hashlib
import os
import sys
import tempfile

import numpy as np
import scipy.io.wavfile as wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from meteor_scatter_b200.synth import synth_file  # noqa: E402
from oracle import ref_harness  # noqa: E402

MB = dict(block_duration_sec=0.2, freq_band=(993, 1013), noise_band=(690, 710), n_fft=512,
          threshold_std_factor=4)
TL = dict(block_duration_sec=0.2, freq_band=(996, 1016), noise_band=(940, 960), n_fft=512,
          threshold_std_factor=3.5)
START = datetime.datetime(2025, 6, 25, 7, 51, 41)

A_CASES = [
    # name, seed, dur_s, params, extra kwargs, synth kwargs
    ("a_mb_s1", 1, 300.0, MB, dict(flag_adaptive_threshold=True), {}),
    ("a_mb_s2", 2, 300.0, MB, dict(flag_adaptive_threshold=True), {}),
    ("a_mb_s3_busy", 3, 300.0, MB, dict(flag_adaptive_threshold=True), dict(rate_per_hour=300.0)),
    ("a_mb_s4_long", 4, 1500.0, MB, dict(flag_adaptive_threshold=True), dict(rate_per_hour=240.0)),
    ("a_mb_s5_global", 5, 300.0, MB, dict(flag_adaptive_threshold=False), dict(rate_per_hour=150.0)),
    ("a_tl_s6", 6, 300.0, TL, dict(flag_adaptive_threshold=True), dict(carrier_hz=1006.0, rate_per_hour=200.0)),
    ("a_mb_s7_slice", 7, 300.0, MB, dict(flag_adaptive_threshold=True, wav_start_sec=30, wav_end_sec=200.5),
     dict(rate_per_hour=250.0)),
    ("a_mb_s8_f32", 8, 120.0, MB, dict(flag_adaptive_threshold=True), dict(rate_per_hour=300.0, dtype=np.float32)),
    ("a_mb_s9_ragged", 9, 61.37, MB, dict(flag_adaptive_threshold=True), dict(rate_per_hour=400.0)),
]

B_CASES = [
    ("b_live1_s11", 11, 240.0, dict(proc_block_sec=0.20, n_fft=4096, detection_db_over_noise_mean_min=1,
                                     detection_dur_min_sec=0.5, signal_freq=1020), dict(carrier_hz=1020.0)),
    ("b_live2_s12", 12, 240.0, dict(proc_block_sec=0.20, n_fft=4096, detection_db_over_noise_mean_min=1,
                                     detection_dur_min_sec=0.5, signal_freq=1025), dict(carrier_hz=1025.0)),
    ("b_default_s13", 13, 120.0, dict(), dict(carrier_hz=1000.0)),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gen_a(tmp):
    for name, seed, dur, params, extra, skw in A_CASES:
        x = synth_file(seed, fs=6000, dur_s=dur, **skw)
        path = os.path.join(tmp, f"{name}.wav")
        wavfile.write(path, 6000, x)
        lbl = os.path.join(tmp, f"{name}.txt")
        csvp = os.path.join(tmp, f"{name}.csv")
        cap = ref_harness.run_reference_a(path, wav_start_date_time=START, out_audacity_lbl_file=lbl,
                                          out_csv_file=csvp, **params, **extra)
        dets = cap["t_out_det"]
        thr = cap["t_threshold"]
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"),
            input_sha256=sha(x), seed=seed, dur_s=dur,
            band_power=np.asarray(cap["band_power"], dtype=np.float64),
            noise_power=np.asarray(cap["noise_power"], dtype=np.float64),
            delta_power=np.asarray(cap["delta_power"], dtype=np.float64),
            thresholds=np.atleast_1d(np.asarray(thr, dtype=np.float64)),
            t_start=np.array([d.t_start for d in dets], dtype=np.float64),
            t_stop=np.array([d.t_stop for d in dets], dtype=np.float64),
            dur=np.array([d.dur_s for d in dets], dtype=np.float64),
            dB=np.array([d.dB for d in dets], dtype=np.float64),
            utc_start=np.array([d.utc_start.isoformat() for d in dets]),
            utc_stop=np.array([d.utc_stop.isoformat() for d in dets]),
            label_text=open(lbl).read(), csv_text=open(csvp, newline="").read())
        print(f"{name}: {len(cap['delta_power'])} blocks, {len(dets)} detections")


def gen_b(tmp):
    for name, seed, dur, cfg, skw in B_CASES:
        x = synth_file(seed, fs=4000, dur_s=dur, rate_per_hour=900.0, **skw)
        path = os.path.join(tmp, f"{name}.wav")
        wavfile.write(path, 4000, x)
        cap = ref_harness.run_reference_b(path, cfg)
        dets = cap["local_out_res_detections"]
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"),
            input_sha256=sha(x), seed=seed, dur_s=dur,
            ms_db=np.asarray(cap["local_data_abs_meas_sig"], dtype=np.float64),
            n1_db=np.asarray(cap["local_data_abs_meas_noise_1"], dtype=np.float64),
            n2_db=np.asarray(cap["local_data_abs_meas_noise_2"], dtype=np.float64),
            db2=np.asarray(cap["local_data_over_noise_sig"], dtype=np.float64),
            thresholds=np.asarray(cap["local_data_over_noise_sig_threshold"], dtype=np.float64),
            det=np.array([[d.time_start, d.time_stop, d.duration, d.db_min, d.db_max, d.db_mean, d.db_std]
                          for d in dets], dtype=np.float64).reshape(-1, 7))
        print(f"{name}: {len(cap['local_data_over_noise_sig'])} blocks, {len(dets)} detections")


if __name__ == "__main__":
    assert ref_harness.reference_available(), "run in the build container (needs /root/reference)"
    with tempfile.TemporaryDirectory() as tmp:
        gen_a(tmp)
        gen_b(tmp)
