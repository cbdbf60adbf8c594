"""ORACLE (test infrastructure, NOT product code) -- detector B, causal "live" path.

numpy/scipy restatement of ``wav_file_process``,
dsp/src/live/backend/processor.py:14-543 (numeric part only: lines 32-75,
176-207, 349-414, 444-510).  Pinned against the unmodified reference through
``oracle/ref_harness.py`` -> ``tests/golden``.

Third-party arithmetic: ``scipy.signal.welch`` (reference pins scipy==1.13.1,
dsp/src/requirements.txt:16; call site processor.py:206).  ``welch_psd`` below
restates its published algorithm for the reference's call form
(window='hann' periodic, nperseg=256, noverlap=128, detrend='constant',
scaling='density', one-sided) and is checked against scipy itself in tests.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class DetectedMeteor:
    """Mirror of dsp/src/live/backend/aggregates.py:66-74."""
    time_start: float
    time_stop: float
    duration: float
    db_min: float
    db_max: float
    db_mean: float
    db_std: float


@dataclass
class ConfigDetection:
    """Mirror of dsp/src/live/backend/aggregates.py:32-44."""
    proc_block_sec: float = 0.2
    n_fft: int = 4096
    signal_freq: int = 1000
    channel_width: int = 100
    noise_channel_offset: int = 300
    avg_win_sec: float = 8
    init_detection_wait_sec: float = 8 * 1.0
    after_tracking_wait_sec: float = 8 * 1.5
    threshold_std_factor: float = 4
    detection_db_over_noise_mean_min: float = -1
    detection_dur_min_sec: float = -1


def band_edges(cfg: ConfigDetection):
    """processor.py:32-44 -> three inclusive (lo, hi) bands: signal, noise1, noise2."""
    hw = cfg.channel_width / 2
    return ((cfg.signal_freq - hw, cfg.signal_freq + hw),
            ((cfg.signal_freq - cfg.noise_channel_offset) - hw, (cfg.signal_freq - cfg.noise_channel_offset) + hw),
            ((cfg.signal_freq + cfg.noise_channel_offset) - hw, (cfg.signal_freq + cfg.noise_channel_offset) + hw))


def welch_psd(block: np.ndarray, fs: float, nfft: int, nperseg: int = 256):
    """Published algorithm of scipy.signal.welch(block, fs, nfft=nfft) as called
    at processor.py:206: hann(nperseg) periodic, 50 % overlap, per-segment mean
    removal, zero-padded rfft, |X|^2/(fs*sum(w^2)), doubled except DC/Nyquist,
    mean over segments."""
    block = np.asarray(block, dtype=np.float64)
    nperseg = min(nperseg, len(block))
    hop = nperseg - nperseg // 2
    n = np.arange(nperseg)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / nperseg)
    nseg = (len(block) - (nperseg // 2)) // hop
    acc = np.zeros(nfft // 2 + 1)
    for s in range(nseg):
        seg = block[s * hop:s * hop + nperseg]
        seg = seg - seg.mean()
        acc += np.abs(np.fft.rfft(seg * w, n=nfft)) ** 2
    psd = acc / nseg / (fs * np.sum(w * w))
    if nfft % 2 == 0:
        psd[1:-1] *= 2
    else:
        psd[1:] *= 2
    return np.fft.rfftfreq(nfft, 1 / fs), psd


def block_band_db(block: np.ndarray, fs: float, cfg: ConfigDetection, use_scipy: bool = True):
    """processor.py:206, 349-367, 393: (ms_dB, noise1_dB, noise2_dB, db2)."""
    if use_scipy:
        from scipy.signal import welch
        freqs, psd = welch(block, fs, nfft=cfg.n_fft)                                   # processor.py:206
    else:
        freqs, psd = welch_psd(block, fs, cfg.n_fft)
    out = []
    for lo, hi in band_edges(cfg):
        m = (freqs >= lo) & (freqs <= hi)
        p = np.sum(psd[m])
        out.append(10 * np.log10(p) if p > 0 else -np.inf)                               # processor.py:352
    db2 = out[0] - np.mean([out[1], out[2]])                                             # processor.py:393
    return out[0], out[1], out[2], db2


def live_state_machine(db2_series, cfg: ConfigDetection, fs: int, block_size: int):
    """Threshold history + 3-state machine, processor.py:393-510.

    Returns (detections, thresholds).  States: 0 Init, 1 Detection, 2 Tracking.
    """
    avg_win = int(cfg.avg_win_sec / cfg.proc_block_sec)                                  # processor.py:56
    over_noise = []
    thresholds = []
    dets = []
    state = 0
    locked = -1.0
    lock_until = -1.0
    t0 = 0.0
    hist = []
    for j, db2 in enumerate(db2_series):
        start_idx = j * block_size
        ts = start_idx / fs                                                              # processor.py:181
        te = (start_idx + block_size) / fs                                               # processor.py:182
        history = over_noise[-avg_win:]                                                  # processor.py:394
        over_noise.append(db2)
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                h_mean = np.mean(history)                                                # processor.py:399
                h_std = np.std(history)
        thr = h_mean + cfg.threshold_std_factor * h_std                                  # processor.py:404
        if state == 2:
            thr = locked                                                                 # processor.py:408
        elif state == 1 and lock_until > te:
            thr = locked                                                                 # processor.py:411-412
        thresholds.append(thr)
        if state == 0:
            if ts >= cfg.init_detection_wait_sec:                                        # processor.py:455
                state, locked, lock_until = 1, -1.0, -1.0
        elif state == 1:
            if db2 > thr:                                                                # processor.py:463
                state, locked, t0, hist = 2, thr + 0 * h_std, ts, []
        else:
            hist.append(db2)                                                             # processor.py:477
            if db2 < thr:                                                                # processor.py:478
                dur = ts - t0
                m = np.mean(hist)
                if m >= cfg.detection_db_over_noise_mean_min and dur >= cfg.detection_dur_min_sec:
                    dets.append(DetectedMeteor(time_start=t0, time_stop=ts, duration=dur, db_min=min(hist),
                                               db_max=max(hist), db_mean=np.mean(hist), db_std=np.std(hist)))
                state, lock_until = 1, ts + cfg.after_tracking_wait_sec                  # processor.py:501-504
    return dets, thresholds


def process(file_data: np.ndarray, fs: int, cfg: ConfigDetection, use_scipy: bool = True):
    """Whole-file restatement: float samples in [-1,1) (soundfile convention,
    processor.py:65-75) -> per-block dB series, thresholds, detections."""
    block_size = int(cfg.proc_block_sec * fs)                                            # processor.py:75
    ms, n1, n2, db2 = [], [], [], []
    for s in range(0, len(file_data) - block_size + 1, block_size):                      # processor.py:176
        a, b, c, d = block_band_db(file_data[s:s + block_size], fs, cfg, use_scipy)
        ms.append(a); n1.append(b); n2.append(c); db2.append(d)
    dets, thr = live_state_machine(db2, cfg, fs, block_size)
    return dict(ms_db=np.array(ms), n1_db=np.array(n1), n2_db=np.array(n2), db2=np.array(db2),
                thresholds=np.array(thr, dtype=np.float64), detections=dets)
