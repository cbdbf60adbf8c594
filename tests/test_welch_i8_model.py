"""CPU model of the tensor-core Welch kernel's number format (csrc/ms_welch_i8.cu): three balanced s8 digits of the
normalised eigenvector columns, offset-binary hi bytes, three accumulated slices with the lo*q3 slice dropped.
Checked against scipy.signal.welch band sums (what processor.py:206, 349-367 computes) before the kernel ever runs."""
import numpy as np
import pytest
from scipy.signal import welch

from meteor_scatter_b200 import ops

BANDS = [(994, 1095), (687, 788), (1301, 1402)]     # f0 = 1020 Hz at nfft 4096, fs 4000 (SURVEY 8a B-band)


def digits3(v):
    q3 = ((v + 128) & 255) - 128
    v1 = (v - q3) // 256
    q2 = ((v1 + 128) & 255) - 128
    q1 = (v1 - q2) // 256
    return q1, q2, q3


def model_band_db(x16, qf, block=800, nperseg=256):
    hop = nperseg // 2
    n_sub = (block - nperseg // 2) // hop
    nb = len(x16) // block
    cols = qf._cols64                                   # [3][26][nperseg]
    out = np.zeros((nb, 3))
    lo = (x16.astype(np.int64) & 255)
    hi_u = ((x16.astype(np.int64) >> 8) + 128)          # offset binary
    for g in range(3):
        e = np.zeros(nb)
        for c in range(cols.shape[1]):
            col = cols[g, c]
            peak = np.abs(col).max()
            if peak == 0:
                continue
            v = np.rint(col * (0.99 / peak) * 2.0 ** 23).astype(np.int64)
            q1, q2, q3 = digits3(v)
            assert np.all(q1 * 65536 + q2 * 256 + q3 == v) and np.abs(q1).max() <= 127
            cs = np.float32(peak / (0.99 * 2.0 ** 15))
            for b in range(nb):
                for s in range(n_sub):
                    sl = slice(b * block + s * hop, b * block + s * hop + nperseg)
                    s0 = int(np.sum(hi_u[sl] * q1)) - 128 * int(q1.sum())
                    s1 = int(np.sum(hi_u[sl] * q2 + lo[sl] * q1)) - 128 * int(q2.sum())
                    s2 = int(np.sum(hi_u[sl] * q3 + lo[sl] * q2)) - 128 * int(q3.sum())
                    assert max(abs(s0), abs(s1), abs(s2)) < 2 ** 24
                    V = np.float32(s0) * np.float32(65536) + (np.float32(s1) * np.float32(256) + np.float32(s2))
                    pr = np.float32(V) * cs
                    e[b] += float(pr) ** 2
        out[:, g] = e * (qf.group_scale[g] / 32768.0 ** 2)
    return 10 * np.log10(out)


@pytest.mark.parametrize("kind", ["noise", "tone"])
def test_number_format_matches_welch(kind):
    rng = np.random.default_rng(3)
    n = np.arange(800 * 3)
    if kind == "noise":
        x = 300 * rng.standard_normal(len(n)) + 2500 * np.sin(2 * np.pi * 1021.7 * n / 4000) * (n > 900)
    else:   # full-scale carrier in the signal band, noise channels ~80 dB below: leakage of the quantised basis
        x = 30000 * np.sin(2 * np.pi * 1020.3 * n / 4000) + 2.0 * rng.standard_normal(len(n))
    x16 = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    qf = ops.WelchQuadform(256, 4096, BANDS, 4000.0, 5, "cpu")
    assert qf.ranks == [26, 26, 26] and qf._tc_tail <= qf.TC_TAIL
    got = model_band_db(x16, qf)
    for b in range(3):
        f, psd = welch(x16[b * 800:(b + 1) * 800].astype(np.float64) / 32768.0, 4000, nfft=4096)
        ref = np.array([10 * np.log10(psd[lo:hi + 1].sum()) for lo, hi in BANDS])
        assert abs(got[b][0] - ref[0]) < 4.4e-4, (b, got[b], ref)            # 1e-4 relative in energy
        # channels 80 dB under a full-scale carrier sit at the rank cut (1e-10 of the carrier's power): 0.01 dB there
        tol = 4.4e-4 if kind == "noise" else 1e-2
        assert np.max(np.abs(got[b][1:] - ref[1:])) < tol, (b, got[b], ref)
