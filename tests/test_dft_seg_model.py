"""CPU model of the general tensor-core band-power kernel (csrc/ms_dft_seg.cu) and of the frequency-domain-window form
(ms_dft_seg_projections_i16 + ms_window_combine), checked against numpy's fp64 windowed STFT.  Runs without a GPU, so
the arithmetic the kernels implement is pinned independently of the hardware:

* hop segments + shifted products: frame f = sum over shifts j of (segment row f+j) x (slice j of the basis), with the
  last slice zero beyond the frame; PCM16 as (lo, hi + 128) unsigned bytes; the basis normalised to its peak and split
  into three balanced s8 digits; four int32 slice sums per column; exact 64-bit recombination with ONE combined
  offset-binary correction per column;
* unwindowed per-segment partial sums, phase rotation, and the cosine-series window applied in the frequency domain.
"""
import numpy as np
import scipy.signal as ss

from meteor_scatter_b200 import ops
from meteor_scatter_b200.synth import synth_file


def digits3(v):
    q3 = ((v + 128) & 255) - 128
    v1 = (v - q3) // 256
    q2 = ((v1 + 128) & 255) - 128
    q1 = (v1 - q2) // 256
    return q1, q2, q3


def int_projections(rows, basis):
    """What the kernel computes for segment rows ``rows`` [n, H] (int16 values) against ``basis`` [H, C] (floats in
    [-1, 1], already a slice of the plan's normalised basis): exact integer projections (int64) and the scale."""
    v = np.rint(basis).astype(np.int64)                                  # basis is passed in already scaled
    q1, q2, q3 = digits3(v)
    assert np.all(q1 * 65536 + q2 * 256 + q3 == v) and np.abs(q1).max() <= 127
    rows = rows.astype(np.int64)
    lo, hi_u = rows & 255, (rows >> 8) + 128                             # the bytes the tensor core sees (u8)
    s0, s1, s2, s3 = hi_u @ q1, hi_u @ q2 + lo @ q1, hi_u @ q3 + lo @ q2, lo @ q3
    for s in (s0, s1, s2, s3):
        assert np.abs(s).max() < 2 ** 31                                 # int32 TMEM accumulators
    return s0, s1, s2, s3, (q1, q2, q3)


def test_segment_form_equals_numpy_stft():
    fs, nfft, hop = 6000, 1024, 104                                      # 90 % overlap: 10 shifts, the last one partial
    x = synth_file(3, fs=fs, dur_s=4.0, rate_per_hour=3000.0)
    frame = nfft
    w = np.hanning(frame)
    bins = [170, 171, 172, 118, 119, 120, 121]
    n = np.arange(frame)
    basis = np.stack([f(2 * np.pi * k * n / nfft) * w for k in bins for f in (np.cos, np.sin)], axis=1)
    scale = 0.99 * float(1 << 23) / np.abs(basis).max()                  # ms_dft_seg_plan_build
    R = -(-frame // hop)
    padded = np.zeros((R * hop, basis.shape[1]))
    padded[:frame] = basis * scale                                       # zero beyond the frame (last slice partial)
    n_rows = len(x) // hop
    rows = x[:n_rows * hop].reshape(n_rows, hop)
    n_frames = n_rows - R + 1
    acc = [np.zeros((n_frames, basis.shape[1]), dtype=np.int64) for _ in range(4)]
    off64 = np.zeros(basis.shape[1], dtype=np.int64)
    for j in range(R):                                                   # the shifted products accumulate in TMEM
        s0, s1, s2, s3, (q1, q2, q3) = int_projections(rows[j:j + n_frames], padded[j * hop:(j + 1) * hop])
        for a, s in zip(acc, (s0, s1, s2, s3)):
            a += s
            assert np.abs(a).max() < 2 ** 31
        off64 += 128 * (q1.sum(axis=0) * (1 << 24) + q2.sum(axis=0) * (1 << 16) + q3.sum(axis=0) * (1 << 8))
    V = (acc[0] << 24) + (acc[1] << 16) + (acc[2] << 8) + acc[3] - off64  # the epilogue's 64-bit recombination
    assert np.abs(V).max() < 2 ** 53
    X = V.astype(np.float64) / scale
    e = (X * X).reshape(n_frames, len(bins), 2).sum(axis=2)
    idx = np.arange(frame)[None, :] + hop * np.arange(n_frames)[:, None]
    ref = np.abs(np.fft.rfft(x[idx].astype(np.float64) * w[None, :], n=nfft, axis=1)) ** 2
    np.testing.assert_allclose(e, ref[:, bins], rtol=1e-5)
    np.testing.assert_allclose(e[:, :3].sum(axis=1), ref[:, bins[:3]].sum(axis=1), rtol=2e-6)


def test_frequency_domain_window_form_equals_numpy_stft():
    fs, nfft = 6000, 2048
    x = synth_file(4, fs=fs, dur_s=6.0, rate_per_hour=3000.0)
    for wname, hop, band in (("hann", 512, (339, 345)), ("blackman", 256, (0, 3)), ("hamming", 1024, (1020, 1024))):
        w = ss.get_window(wname, nfft)
        order, coef = ops.cosine_series(w)
        ext = list(range(band[0] - order, band[1] + order + 1))          # band bins +- window order (may leave 0..nfft/2)
        i = np.arange(hop)
        basis = np.stack([f(2 * np.pi * ((k * i) % nfft) / nfft) for k in ext for f in (np.cos, np.sin)], axis=1)
        scale = 0.99 * float(1 << 23) / np.abs(basis).max()
        n_rows = len(x) // hop
        rows = x[:n_rows * hop].reshape(n_rows, hop)
        s0, s1, s2, s3, (q1, q2, q3) = int_projections(rows, basis * scale)
        off64 = 128 * (q1.sum(axis=0) * (1 << 24) + q2.sum(axis=0) * (1 << 16) + q3.sum(axis=0) * (1 << 8))
        P = (((s0 << 24) + (s1 << 16) + (s2 << 8) + s3 - off64).astype(np.float64) / scale).reshape(n_rows, len(ext), 2)
        P = P[:, :, 0] - 1j * P[:, :, 1]                                 # unwindowed partial sums of every segment
        R = nfft // hop
        n_frames = n_rows - R + 1
        rect = np.zeros((n_frames, len(ext)), dtype=np.complex128)
        for j in range(R):                                               # ms_window_combine: rotate and add
            rot = np.exp(-2j * np.pi * ((np.array(ext) * j * hop) % nfft) / nfft)
            rect += P[j:j + n_frames] * rot[None, :]
        c = np.arange(order, len(ext) - order)
        X = coef[0] * rect[:, c]
        for m in range(1, order + 1):
            X = X + coef[m] * (rect[:, c - m] + rect[:, c + m])
        idx = np.arange(nfft)[None, :] + hop * np.arange(n_frames)[:, None]
        ref = np.fft.rfft(x[idx].astype(np.float64) * w[None, :], axis=1)[:, band[0]:band[1] + 1]
        np.testing.assert_allclose(np.abs(X) ** 2, np.abs(ref) ** 2, rtol=2e-5, atol=1e-9 * float((np.abs(ref) ** 2).max()),
                                   err_msg=wname)
