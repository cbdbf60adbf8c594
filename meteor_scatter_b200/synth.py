"""Seeded synthetic BRAMS-beacon audio (SURVEY.md §8(d) "Synthetic input").

Mono PCM16: Gaussian receiver noise plus meteor pings at the beacon's audio
carrier (1003 Hz for the reference's ``mb_files`` parameter set,
dsp/src/main.py:827).  Underdense pings decay exponentially, overdense pings
have a plateau followed by a decay.  Two generators share one recipe:

* ``synth_file`` -- numpy, bit-reproducible from ``seed`` (tests, goldens,
  the CPU baseline sample);
* ``synth_batch_torch`` -- the same recipe as torch ops so a 24 h / 30 day
  batch can be produced directly in HBM for the benchmark.

This module is input generation only; it is not part of the detection path.
"""
from __future__ import annotations

import math

import numpy as np

# durations (seconds) drawn uniformly; <=0.4 s underdense, >=0.8 s overdense
PING_DURATIONS = (0.1, 0.2, 0.4, 0.8, 2.0, 6.0)


def ping_schedule(seed: int, dur_s: float, rate_per_hour: float = 120.0):
    """Poisson ping list for one file: (t0, duration, amplitude, phase)."""
    rng = np.random.default_rng([seed, 0x9E3779B9])
    n = int(rng.poisson(rate_per_hour * dur_s / 3600.0))
    t0 = np.sort(rng.uniform(0.0, dur_s, n))
    dur = rng.choice(np.asarray(PING_DURATIONS), n)
    amp = rng.uniform(150.0, 4000.0, n)
    ph = rng.uniform(0.0, 2.0 * math.pi, n)
    return t0, dur, amp, ph


def ping_envelope(t: np.ndarray, t0: float, dur: float) -> np.ndarray:
    """Amplitude envelope of one ping evaluated at times ``t`` (seconds)."""
    rel = t - t0
    env = np.zeros_like(rel)
    if dur <= 0.4:  # underdense: exponential decay, tau = dur/3
        m = (rel >= 0) & (rel < dur)
        env[m] = np.exp(-rel[m] / (dur / 3.0))
    else:  # overdense: plateau for 70 % then decay
        plateau = 0.7 * dur
        m1 = (rel >= 0) & (rel < plateau)
        env[m1] = 1.0
        m2 = (rel >= plateau) & (rel < dur)
        env[m2] = np.exp(-(rel[m2] - plateau) / ((dur - plateau) / 3.0))
    return env


def synth_file(seed: int, fs: int = 6000, dur_s: float = 300.0, carrier_hz: float = 1003.0,
               noise_sigma: float = 300.0, rate_per_hour: float = 120.0,
               dtype=np.int16) -> np.ndarray:
    """One synthetic recording, PCM16 by default (float32 in [-1,1) if asked)."""
    n = int(round(fs * dur_s))
    rng = np.random.default_rng([seed, 0x51ED270B])
    x = rng.standard_normal(n, dtype=np.float32).astype(np.float64) * noise_sigma
    t0s, durs, amps, phs = ping_schedule(seed, dur_s, rate_per_hour)
    for t0, d, a, ph in zip(t0s, durs, amps, phs):
        i0 = max(0, int(t0 * fs))
        i1 = min(n, int((t0 + d) * fs) + 1)
        if i1 <= i0:
            continue
        t = np.arange(i0, i1) / fs
        x[i0:i1] += a * ping_envelope(t, t0, d) * np.sin(2.0 * math.pi * carrier_hz * t + ph)
    pcm = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    if dtype == np.int16:
        return pcm
    if dtype == np.float32:
        return (pcm.astype(np.float32) / 32768.0).astype(np.float32)
    raise ValueError(f"unsupported dtype {dtype}")


def synth_batch_torch(n_files: int, samples_per_file: int, fs: int = 6000, carrier_hz: float = 1003.0,
                      noise_sigma: float = 300.0, rate_per_hour: float = 120.0, seed: int = 0,
                      device="cuda", files_per_chunk: int = 32):
    """``[n_files, samples_per_file]`` int16 tensor generated on ``device``.

    Same recipe as ``synth_file`` (noise + enveloped carrier bursts) but with
    torch's generator, so the values differ from the numpy version; parity
    checks always run both arms on the *same* tensor.
    """
    import torch

    dev = torch.device(device)
    out = torch.empty((n_files, samples_per_file), dtype=torch.int16, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    dur_s = samples_per_file / fs
    tt = torch.arange(samples_per_file, device=dev, dtype=torch.float32) / fs
    for f0 in range(0, n_files, files_per_chunk):
        f1 = min(n_files, f0 + files_per_chunk)
        x = torch.randn((f1 - f0, samples_per_file), generator=gen, device=dev, dtype=torch.float32)
        x.mul_(noise_sigma)
        for f in range(f0, f1):
            t0s, durs, amps, phs = ping_schedule(seed * 1_000_003 + f, dur_s, rate_per_hour)
            for t0, d, a, ph in zip(t0s, durs, amps, phs):
                i0 = max(0, int(t0 * fs))
                i1 = min(samples_per_file, int((t0 + d) * fs) + 1)
                if i1 <= i0:
                    continue
                t = tt[i0:i1]
                rel = t - float(t0)
                if d <= 0.4:
                    env = torch.exp(-rel / (d / 3.0))
                else:
                    plateau = 0.7 * d
                    env = torch.where(rel < plateau, torch.ones_like(rel),
                                      torch.exp(-(rel - plateau) / ((d - plateau) / 3.0)))
                env = env * ((rel >= 0) & (rel < d))
                x[f - f0, i0:i1] += float(a) * env * torch.sin(2.0 * math.pi * carrier_hz * t + float(ph))
        out[f0:f1] = x.round_().clamp_(-32768, 32767).to(torch.int16)
    return out
