"""ORACLE (test infrastructure, NOT product code) -- detector C numeric stage and
the dashboard CSV contract.

Restates meteor_detect_class/prime_detection.py:67-92 (specgram + noise-floor
scalar + vmin) and :132-146 / :229-270 (the only writer of
``Timestamp;Anzahl;Kritisch``).

Third-party arithmetic absent from /root/reference: ``matplotlib.mlab.specgram``
(reference pins matplotlib==3.9.4, dsp/src/requirements.txt:6; call site
prime_detection.py:70).  matplotlib is not installed here, so the reference's
``plot_spectrogram`` cannot run: **parity unpinned** for this function.  We
restate mlab's published PSD algorithm (np.hanning(NFFT) window, no detrend,
one-sided, x2 except DC/Nyquist, /Fs, /sum(w^2)) and cross-check it against
``scipy.signal.spectrogram`` in the call form the reference itself uses at
dsp/src/main.py:52-54.
"""
from __future__ import annotations

import numpy as np

C_MS_SPEC_CUT_FACTOR = 12          # prime_detection.py:22


def specgram_psd(x: np.ndarray, fs: float, nfft: int = 2048, noverlap: int = 1024, window=None):
    """mlab.specgram(mode='psd') restated: returns (Pxx[F, T], freqs, bins)."""
    x = np.asarray(x, dtype=np.float64)
    if window is None:
        window = np.hanning(nfft)                      # mlab.window_hanning
    hop = nfft - noverlap
    nframes = (len(x) - noverlap) // hop
    frames = np.stack([x[i * hop:i * hop + nfft] for i in range(nframes)], axis=1)
    spec = np.fft.rfft(frames * window[:, None], n=nfft, axis=0)
    pxx = np.abs(spec) ** 2 / (fs * np.sum(window ** 2))
    if nfft % 2 == 0:
        pxx[1:-1] *= 2
    else:
        pxx[1:] *= 2
    freqs = np.fft.rfftfreq(nfft, 1 / fs)
    bins = (np.arange(nframes) * hop + nfft / 2) / fs
    return pxx, freqs, bins


def noise_floor_vmin(pxx: np.ndarray, freqs: np.ndarray, fs: float, nfft: int = 2048,
                     lower_freq: float = 250, upper_freq: float = 800):
    """prime_detection.py:73-91: noise-band density in dB/Hz and the adaptive vmin."""
    delta_f = fs / nfft                                                  # :69
    noise_band = (freqs >= lower_freq) & (freqs <= upper_freq)           # :75
    bandwidth = np.sum(noise_band) * delta_f                             # :77
    band_power = np.sum(pxx[noise_band])                                 # :83 (over time too)
    power_density_db_hz = 10 * np.log10(band_power / bandwidth)          # :84
    factor = 40 / 23                                                     # :85
    temp_vmin = power_density_db_hz / factor + C_MS_SPEC_CUT_FACTOR      # :91
    return power_density_db_hz, temp_vmin


def plot_spectrogram_numeric(iq_segment: np.ndarray, fs: float, f_lo: float = 800, f_hi: float = 1200):
    """Numeric content of plot_spectrogram (prime_detection.py:65-105) without
    rendering: band-limited Pxx in dB (rows 800-1200 Hz), density, vmin."""
    x = iq_segment[:, 0] if iq_segment.ndim == 2 else iq_segment
    pxx, freqs, bins = specgram_psd(x, fs)
    dens, vmin = noise_floor_vmin(pxx, freqs, fs)
    with np.errstate(divide="ignore"):
        pxx_db = 10 * np.log10(pxx)                                      # :88
    rows = (freqs >= f_lo) & (freqs <= f_hi)
    return dict(pxx=pxx, pxx_db_band=pxx_db[rows], freqs=freqs, bins=bins, rows=np.nonzero(rows)[0],
                density_db_hz=dens, vmin=vmin, vmax=40)


def hourly_csv_text(rows) -> str:
    """Text of a day file as pandas writes it at prime_detection.py:138, 245:
    ``to_csv(sep=';', index=False)`` of columns Timestamp;Anzahl;Kritisch,
    '\\n' line ends.  ``rows`` = iterable of (timestamp_str, anzahl, kritisch)."""
    out = ["Timestamp;Anzahl;Kritisch"]
    for ts, a, k in rows:
        out.append(f"{ts};{int(a)};{int(k)}")
    return "\n".join(out) + "\n"
