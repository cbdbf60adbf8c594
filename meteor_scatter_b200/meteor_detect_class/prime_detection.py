"""Numeric stage and CSV contract of the reference's detector C
(meteor_detect_class/prime_detection.py) on the GPU.

* ``plot_spectrogram`` keeps the reference signature (:65) and computes what its
  body computes before rendering: the one-sided PSD spectrogram
  ``plt.specgram(x, Fs, NFFT=2048, noverlap=1024)`` (:70-71), the noise-band
  power density in dB/Hz (:73-84), the adaptive ``vmin`` (:85-91) and the
  800-1200 Hz rows in dB that the image shows (:88, :100).  The band-limited dB
  matrix is returned as a device tensor and, with ``out_path``, rendered to the
  JPG that ``detect_and_cluster_bursts`` consumes (``render.py``; Pillow, off the
  hot path, no matplotlib).
* ``HourlyCsv`` reproduces the only writer of ``Timestamp;Anzahl;Kritisch``
  (:132-146 day file creation, :229-252 hourly row, :255-270 day roll).

Unlike the reference module, importing this one has no side effects (no Twitch
grabber, no path assertions at import time).
"""
from __future__ import annotations

import datetime
import os

import numpy as np
import torch

from meteor_scatter_b200 import csvout, ops

C_MS_SPEC_CUT_FACTOR = 12          # prime_detection.py:22
C_SAMPLE_RATE = 5000               # :31
C_SEG_LEN = 30                     # :32
NFFT = 2048                        # :68


C_FILE_PATH_SPEC = "/tmp/spectrogram2.jpg"  # :27 (where the reference saves the rendered segment)


def plot_spectrogram(iq_segment, fs, display=True, vmin=10, vmax=30, *, device="cuda", out_path=None):
    """iq_segment: ``[n, 1]`` (or ``[n]``) int16/float32 samples, numpy or CUDA tensor.
    Returns dict(pxx_db_band [164, T] CUDA float32, freqs_band, bins, density_db_hz, vmin, vmax=40[, image_path]).
    ``out_path`` writes the JPG the reference saves at :105 (``imshow(Pxx_db, vmin=temp_vmin, vmax=40)``,
    ``ylim(800, 1200)``, axes off, 496 x 370 px) from the device tensor, so ``detect_and_cluster_bursts(out_path)``
    runs on a GPU-produced spectrogram without matplotlib (``render.save_spectrogram_jpg``)."""
    if isinstance(iq_segment, np.ndarray):
        x = iq_segment[:, 0] if iq_segment.ndim == 2 else iq_segment          # :70 iq_segment[:, 0]
        if x.dtype not in (np.int16, np.float32):
            x = x.astype(np.float32)
        x = torch.from_numpy(np.ascontiguousarray(x)).to(device)
    else:
        x = iq_segment[:, 0] if iq_segment.dim() == 2 else iq_segment
        x = x.contiguous()
    delta_f = fs / NFFT                                                         # :69
    freqs = np.fft.rfftfreq(NFFT, 1 / fs)
    noise_band = (freqs >= 250) & (freqs <= 800)                                # :73-75
    nk = np.nonzero(noise_band)[0]
    rows = np.nonzero((freqs >= 800) & (freqs <= 1200))[0]                      # :100 plt.ylim(800, 1200)
    psd, noise_sum = ops.psd_spectrogram(x, float(fs), NFFT, NFFT // 2, np.hanning(NFFT), int(rows[0]),
                                         int(rows[-1]), int(nk[0]), int(nk[-1]))
    bandwidth = np.sum(noise_band) * delta_f                                    # :77
    band_power = float(noise_sum[0].item())                                     # :83
    power_density_db_hz = 10 * np.log10(band_power / bandwidth)                 # :84
    factor = 40 / 23                                                            # :85
    temp_vmin = power_density_db_hz / factor + C_MS_SPEC_CUT_FACTOR             # :91
    pxx_db = 10.0 * torch.log10(psd[0])                                         # :88 (log of 0 -> -inf, as :89)
    n_frames = psd.shape[2]
    bins = (np.arange(n_frames) * (NFFT // 2) + NFFT / 2) / fs
    out = dict(pxx_db_band=pxx_db, freqs_band=freqs[rows], bins=bins, density_db_hz=power_density_db_hz,
               vmin=temp_vmin, vmax=40)
    if out_path is not None:
        from meteor_scatter_b200 import render
        out["image_path"] = render.save_spectrogram_jpg(out_path, pxx_db, temp_vmin, 40)     # :96-105
    return out


class HourlyCsv:
    """Accumulates burst counts and appends one ``Timestamp;Anzahl;Kritisch`` row per
    elapsed interval to ``<folder>/YYYYMMDD.csv`` (prime_detection.py:129-146, 229-270).
    One deliberate difference: at the day roll the reference re-creates the new day's file even if it exists
    (:262-264, truncating earlier rows of that day); here an existing day file is kept and appended to, so a
    restarted service does not lose the hours it already wrote."""

    def __init__(self, folder: str, now: datetime.datetime | None = None,
                 save_interval: datetime.timedelta = datetime.timedelta(minutes=59.8)):
        assert os.path.exists(folder), f"Path not found: {folder}"              # :35
        self.folder = folder
        self.save_interval = save_interval                                      # :129
        now = now or datetime.datetime.now()
        self.start_time = now
        self.previous_date = now.strftime('%Y-%m-%d')
        self.n_critical = 0
        self.n_non_critical = 0
        self.file_name = self._ensure_day_file(now)

    def _ensure_day_file(self, now):
        path = os.path.join(self.folder, csvout.day_file_name(now.date()))
        if not os.path.exists(path):                                            # :141-146
            with open(path, "w", newline="") as f:
                f.write(csvout.HEADER + "\n")
        return path

    def add(self, n_critical: int, n_non_critical: int, now: datetime.datetime | None = None):
        """One processed segment (:215-216), then the hourly / daily bookkeeping (:229-270)."""
        now = now or datetime.datetime.now()
        self.n_critical += int(n_critical)
        self.n_non_critical += int(n_non_critical)
        wrote = None
        if now - self.start_time >= self.save_interval:                         # :229
            row = csvout.format_row(self.start_time, self.n_critical + self.n_non_critical, self.n_critical)
            with open(self.file_name, "a", newline="") as f:
                f.write(row + "\n")
            wrote = row
            self.n_critical = 0
            self.n_non_critical = 0
            self.start_time = now
        current_date = now.strftime('%Y-%m-%d')
        if current_date != self.previous_date:                                  # :255-270
            self.previous_date = current_date
            self.file_name = self._ensure_day_file(now)
            self.n_critical = 0
            self.n_non_critical = 0
        return wrote
