"""CPU model of the tensor-core band-power kernel's number format (csrc/ms_dft_i8.cu), checked against the golden
band powers of the unmodified reference: PCM16 as (lo, hi) bytes with offset-binary hi bytes, the basis normalised to
its peak and split into three balanced s8 digits, four int32 slice sums, exact recombination.  Runs without a GPU, so
the arithmetic the kernel implements is pinned independently of the hardware."""
import os

import numpy as np

from meteor_scatter_b200 import ops
from tests.golden_cases import A_CASES, MB, a_sliced

DB_TOL = 10 * np.log10(1 + 1e-4)


def _spec(params, fs=6000):
    return ops.BandSpec.from_reference_args(fs, params["block_duration_sec"], params["freq_band"], params["noise_band"],
                                            params["n_fft"])


def digits3(v):
    q3 = ((v + 128) & 255) - 128
    v1 = (v - q3) // 256
    q2 = ((v1 + 128) & 255) - 128
    q1 = (v1 - q2) // 256
    return q1, q2, q3


def test_number_format_reproduces_reference_band_power():
    x, g = a_sliced("a_mb_s1")
    spec = _spec(MB)
    basis, group = ops.DftI8Plan.basis_for(spec)
    scale = 0.99 * float(1 << 23) / np.abs(basis).max()                  # ms_dft_i8_plan_build
    v = np.rint(basis * scale).astype(np.int64)
    q1, q2, q3 = digits3(v)
    assert np.all(q1 * 65536 + q2 * 256 + q3 == v)
    assert max(np.abs(q1).max(), np.abs(q2).max(), np.abs(q3).max()) <= 128 and q1.max() <= 127 and q2.max() <= 127
    nb = spec.n_blocks(len(x))
    blocks = x[:nb * spec.block_size].reshape(nb, spec.block_size)[:, :spec.win_len].astype(np.int64)
    lo, hi_u = blocks & 255, (blocks >> 8) + 128                         # the bytes the tensor core sees (u8)
    assert lo.min() >= 0 and lo.max() <= 255 and hi_u.min() >= 0 and hi_u.max() <= 255
    s0 = hi_u @ q1 - 128 * q1.sum(axis=0)
    s1 = hi_u @ q2 + lo @ q1 - 128 * q2.sum(axis=0)
    s2 = hi_u @ q3 + lo @ q2 - 128 * q3.sum(axis=0)
    s3 = lo @ q3
    for s in (hi_u @ q1, hi_u @ q2 + lo @ q1, hi_u @ q3 + lo @ q2, s3):  # raw accumulators fit int32
        assert np.abs(s).max() < 2 ** 31
    V = ((s0 * 256 + s1) * 256 + s2) * 256 + s3
    assert np.array_equal(V, blocks @ v) and np.abs(V).max() < 2 ** 53   # exact recombination, exact in fp64
    X = V.astype(np.float64) * (1.0 / scale)
    e = X * X
    band = 10 * np.log10(e[:, group == 0].sum(axis=1) + 1e-12)           # main.py:383-384
    noise = 10 * np.log10(e[:, group == 1].sum(axis=1) + 1e-12)          # main.py:387-388
    assert np.max(np.abs(band - g["band_power"])) < DB_TOL
    assert np.max(np.abs(noise - g["noise_power"])) < DB_TOL
    # and far inside it: the only approximation is the 2^-24 quantisation of the basis
    assert np.max(np.abs(band - g["band_power"])) < 1e-5
