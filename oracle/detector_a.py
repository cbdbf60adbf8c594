"""ORACLE (test infrastructure, NOT product code) -- detector A, batch DSP path.

CPU restatement in numpy of the reference's ``proc_wav_file`` hot path,
dsp/src/main.py:352-527 and its writers 626-658 / hour bucketing 687-700.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; the product path never does.

Parity status: the reference ships no golden vectors, so this
restatement is pinned against the *reference itself* run in the build
container (``oracle/ref_harness.py`` imports /root/reference unmodified and
``tests/golden/make_golden.py`` stores its outputs as fixtures);
``tests/test_oracle_vs_golden.py`` checks every function below against them.

The FFT itself lives in numpy (``np.fft.rfft``, pocketfft; the reference pins
numpy==2.0.2 at dsp/src/requirements.txt:8) -- we call the same routine.
"""
from __future__ import annotations

import datetime
from dataclasses import dataclass

import numpy as np


@dataclass
class OutputDetection:
    """Mirror of dsp/src/main.py:30-37."""
    t_start: float
    t_stop: float
    dur_s: float
    dB: float
    utc_start: datetime.datetime = None
    utc_stop: datetime.datetime = None


def block_geometry(n_samples: int, fs: int, block_duration_sec: float, n_fft: int):
    """dsp/src/main.py:352-363: n_fft doubled, block size truncated, tail dropped."""
    n_fft_real = n_fft * 2                                   # main.py:353
    block_size = int(fs * block_duration_sec)                # main.py:355
    num_blocks = n_samples // block_size                     # main.py:356
    freqs = np.fft.rfftfreq(n_fft_real, d=1 / fs)            # main.py:363
    return n_fft_real, block_size, num_blocks, freqs


def band_bins(freqs: np.ndarray, band) -> np.ndarray:
    """Inclusive mask on rfftfreq, dsp/src/main.py:382,386 -> bin indices."""
    return np.nonzero((freqs >= band[0]) & (freqs <= band[1]))[0]


def stft_band_power(wav_data: np.ndarray, fs: int, block_duration_sec: float, freq_band, noise_band,
                    n_fft: int):
    """Literal restatement of the STFT loop dsp/src/main.py:376-388.

    Same library calls per block as the reference (np.hanning recomputed,
    np.fft.rfft(n=n_fft) cropping/zero-padding the windowed block), so it is
    also the honest CPU-cost model used as ``cpu_baseline`` kind "port".
    Returns (band_power_dB, noise_power_dB) as float64 arrays.
    """
    n_fft_real, block_size, num_blocks, freqs = block_geometry(len(wav_data), fs, block_duration_sec, n_fft)
    band_power = []
    noise_power = []
    for i in range(num_blocks):
        block = wav_data[i * block_size:(i + 1) * block_size]
        fft_block = np.fft.rfft(block * np.hanning(len(block)), n=n_fft_real)      # main.py:379
        power_spectrum = np.abs(fft_block) ** 2                                     # main.py:380
        band_mask = (freqs >= freq_band[0]) & (freqs <= freq_band[1])               # main.py:382
        band_energy = np.sum(power_spectrum[band_mask]) + 1e-12                     # main.py:383
        band_power.append(10 * np.log10(band_energy))                               # main.py:384
        noise_mask = (freqs >= noise_band[0]) & (freqs <= noise_band[1])            # main.py:386
        noise_energy = np.sum(power_spectrum[noise_mask]) + 1e-12                   # main.py:387
        noise_power.append(10 * np.log10(noise_energy))                             # main.py:388
    return np.array(band_power, dtype=np.float64), np.array(noise_power, dtype=np.float64)


def stft_band_energy_vec(wav_data: np.ndarray, fs: int, block_duration_sec: float, freq_band, noise_band,
                         n_fft: int):
    """Vectorised equivalent of the loop above returning *linear* energies
    (before +1e-12 and log10); used by tests for relative-error checks."""
    n_fft_real, block_size, num_blocks, freqs = block_geometry(len(wav_data), fs, block_duration_sec, n_fft)
    blocks = np.asarray(wav_data[:num_blocks * block_size]).reshape(num_blocks, block_size)
    spec = np.fft.rfft(blocks * np.hanning(block_size)[None, :], n=n_fft_real, axis=1)
    p = np.abs(spec) ** 2
    return p[:, band_bins(freqs, freq_band)].sum(axis=1), p[:, band_bins(freqs, noise_band)].sum(axis=1)


def _mk_detection(delta_power, start, stop, block_duration_sec, wav_start_date_time):
    db_mean = np.mean(delta_power[start:stop])
    t_start = start * block_duration_sec
    t_stop = stop * block_duration_sec
    t_dur = t_stop - t_start
    u0 = u1 = None
    if wav_start_date_time is not None:
        u0 = wav_start_date_time + datetime.timedelta(seconds=t_start)
        u1 = wav_start_date_time + datetime.timedelta(seconds=t_stop)
    return OutputDetection(t_start=t_start, t_stop=t_stop, dur_s=t_dur, dB=db_mean, utc_start=u0, utc_stop=u1)


def get_detections(delta_power: np.ndarray, threshold_std_factor: float, block_duration_sec: float,
                   wav_start_date_time=None):
    """Global-threshold detector, dsp/src/main.py:396-448.

    Returns (detections, threshold, index_pairs).  Reproduces the reference's
    quirks: zip() truncation (main.py:420), stop = len-1 for an event open at
    EOF (main.py:414-415) and the AssertionError for a zero-length event
    (main.py:437).
    """
    t_threshold = np.mean(delta_power) + threshold_std_factor * np.std(delta_power)    # main.py:399-400
    above = delta_power > t_threshold                                                    # main.py:405
    d = np.diff(above.astype(int))
    starts = np.where(d == 1)[0] + 1                                                     # main.py:408
    stops = np.where(d == -1)[0] + 1                                                     # main.py:409
    if above[0]:
        starts = np.insert(starts, 0, 0)                                                 # main.py:412-413
    if above[-1]:
        stops = np.append(stops, len(delta_power) - 1)                                   # main.py:414-415
    out, pairs = [], []
    for start, stop in zip(starts, stops):                                               # main.py:420
        det = _mk_detection(delta_power, int(start), int(stop), block_duration_sec, wav_start_date_time)
        if wav_start_date_time is not None:
            assert det.utc_start < det.utc_stop, "UTC start time must be before stop time"   # main.py:435
        assert det.dur_s > 0, "Detection duration must be greater than 0"                # main.py:437
        out.append(det)
        pairs.append((int(start), int(stop)))
    return out, t_threshold, pairs


def adaptive_params(block_duration_sec, window_sec=120, before_sec=3, after_sec=20, fixed_sec=10):
    """int() truncations of dsp/src/main.py:458-461 (int(0.6/0.2) == 2!)."""
    return (int(window_sec / block_duration_sec), int(before_sec / block_duration_sec),
            int(after_sec / block_duration_sec), int(fixed_sec / block_duration_sec))


def get_detections_adaptive(delta_power: np.ndarray, threshold_std_factor: float, block_duration_sec: float,
                            wav_start_date_time=None, threshold_estimation_window_sec=120,
                            threshold_freeze_before_detection_sec=3, threshold_freeze_after_detection_sec=20,
                            fixed_threshold_duration_sec=10):
    """Adaptive-threshold detector, dsp/src/main.py:450-522 (literal O(N*W) loop).

    Returns (detections, thresholds[list], index_pairs[(start, stop_exclusive)]).
    """
    num_blocks = len(delta_power)
    detections = []
    freeze_until_idx = -1
    thresholds = []
    window_blocks, freeze_blocks_before, freeze_blocks_after, fixed_threshold_blocks = adaptive_params(
        block_duration_sec, threshold_estimation_window_sec, threshold_freeze_before_detection_sec,
        threshold_freeze_after_detection_sec, fixed_threshold_duration_sec)
    global_mean = np.mean(delta_power)                                                   # main.py:464
    global_std = np.std(delta_power)                                                     # main.py:465
    fixed_threshold = global_mean + threshold_std_factor * global_std                    # main.py:466
    threshold = fixed_threshold
    for i in range(num_blocks):                                                          # main.py:470
        if i < fixed_threshold_blocks:
            threshold = fixed_threshold
        elif i > freeze_until_idx:
            window_delta = delta_power[max(0, i - window_blocks):i]                      # main.py:475-477
            threshold = np.mean(window_delta) + threshold_std_factor * np.std(window_delta)
        thresholds.append(threshold)
        if delta_power[i] > threshold:                                                   # main.py:485
            if not detections or i > detections[-1]['stop'] + 1:
                detections.append({'start': i, 'stop': i})
            else:
                detections[-1]['stop'] = i
            freeze_until_idx = i + freeze_blocks_after                                   # main.py:491
            freeze_start_idx = max(0, i - freeze_blocks_before)
            freeze_until_idx = max(freeze_until_idx, freeze_start_idx)                   # main.py:493 (no-op)
    out, pairs = [], []
    for d in detections:                                                                 # main.py:497
        start, stop = d['start'], d['stop'] + 1
        out.append(_mk_detection(delta_power, start, stop, block_duration_sec, wav_start_date_time))
        pairs.append((start, stop))
    return out, thresholds, pairs


def detect_wav(wav_data, fs, block_duration_sec, freq_band, noise_band, n_fft, threshold_std_factor,
               wav_start_date_time=None, flag_adaptive_threshold=True, **adaptive_kw):
    """STFT loop + delta (main.py:393) + detector selection (main.py:524-527)."""
    band, noise = stft_band_power(wav_data, fs, block_duration_sec, freq_band, noise_band, n_fft)
    delta_power = band - noise                                                           # main.py:393
    if not flag_adaptive_threshold:
        dets, thr, pairs = get_detections(delta_power, threshold_std_factor, block_duration_sec,
                                          wav_start_date_time)
    else:
        dets, thr, pairs = get_detections_adaptive(delta_power, threshold_std_factor, block_duration_sec,
                                                   wav_start_date_time, **adaptive_kw)
    return dict(band_power=band, noise_power=noise, delta_power=delta_power, detections=dets,
                threshold=thr, pairs=pairs)


def audacity_label_text(dets) -> str:
    """dsp/src/main.py:630-635."""
    return "".join(f"{d.t_start:.2f}\t{d.t_stop:.2f}\tM\n" for d in dets)


def event_csv_rows(dets):
    """Row dicts of the per-event CSV, dsp/src/main.py:640-656."""
    return [{'t_start': d.t_start, 't_stop': d.t_stop, 'dur_s': d.dur_s, 'dB': d.dB,
             'utc_start': d.utc_start.isoformat() if d.utc_start else None,
             'utc_stop': d.utc_stop.isoformat() if d.utc_stop else None} for d in dets]


def hourly_counts(dets, critical_min_dur_s: float = 0.5):
    """Hour bucketing by utc_start (dsp/src/main.py:690-696, main_analyze.py:70-73)
    plus the critical rule "duration >= 0.5 s"
    (meteor_detect_class/detector_and_classification.py:50 and README.md:75-76).
    Returns {hour_datetime: [Anzahl, Kritisch]}.
    """
    out = {}
    for d in dets:
        h = d.utc_start.replace(minute=0, second=0, microsecond=0)
        c = out.setdefault(h, [0, 0])
        c[0] += 1
        if d.dur_s >= critical_min_dur_s:
            c[1] += 1
    return out
