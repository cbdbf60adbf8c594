"""CPU model of the warp-per-frame PSD kernel's decomposition (csrc/ms_fft_warp.cuh): the index maps the kernel
relies on, checked against numpy's rfft.  No GPU, no extension: this pins the algebra the CUDA code implements --

* a 2048-sample real frame packed as 1024 complex points z[m] = x[2m] + i x[2m+1];
* 1024 = 32 x 32: lane n2 holds z[32 n1 + n2]; a 32-point transform over n1, the twiddle W^(n2 k1), a 32 x 32
  transpose, a 32-point transform over n2; register k2 of lane k1 then holds Z[k1 + 32 k2];
* each 32-point transform as one radix-2 level around two 16-point transforms that run in lockstep (the two halves
  of the packed fp32 registers): decimation in time in stage 1, decimation in frequency in stage 2;
* the real split X[k] = E + w^k O with the partner Z[1024 - k] in lane (32 - k1) & 31, register 31 - k2 (lane 0:
  (32 - k2) & 31), and the one-sided PSD scaling with the factors 1/2 folded into the scale.
"""
import numpy as np


def fft32_dit(v):
    """32-point DFT along axis 0 as the kernel's stage 1 does it: 16-point transforms of the even and odd inputs
    (the lockstep pair), then X[k] = E[k] + w32^k O[k], X[k + 16] = E[k] - w32^k O[k]."""
    e = np.fft.fft(v[0::2], axis=0)
    o = np.fft.fft(v[1::2], axis=0)
    w = np.exp(-2j * np.pi * np.arange(16) / 32).reshape((16,) + (1,) * (v.ndim - 1))
    return np.concatenate([e + w * o, e - w * o], axis=0)


def fft32_dif(v):
    """32-point DFT along axis 0 as stage 2 does it: a[j] = x[j] + x[j+16], b[j] = (x[j] - x[j+16]) w32^j, then the
    lockstep 16-point transforms give X[2k] = A[k], X[2k+1] = B[k]."""
    w = np.exp(-2j * np.pi * np.arange(16) / 32).reshape((16,) + (1,) * (v.ndim - 1))
    a = v[:16] + v[16:]
    b = (v[:16] - v[16:]) * w
    out = np.empty_like(v)
    out[0::2] = np.fft.fft(a, axis=0)
    out[1::2] = np.fft.fft(b, axis=0)
    return out


def warp_frame_model(x, window):
    """One frame through the kernel's data flow; returns regs[k2, k1] = Z[k1 + 32 k2] (register k2 of lane k1)."""
    xw = x.astype(np.float64) * window
    z = xw[0::2] + 1j * xw[1::2]                       # 1024 packed points
    v = z.reshape(32, 32)                               # v[n1, n2] = z[32 n1 + n2]: lane n2, register n1
    y = fft32_dit(v)                                    # y[k1, n2]
    k1 = np.arange(32)[:, None]
    n2 = np.arange(32)[None, :]
    y = y * np.exp(-2j * np.pi * k1 * n2 / 1024)        # the shared-memory twiddle table [k1][n2]
    t = y.T                                             # transpose: lane k1 now holds t[n2, k1] in register n2
    return fft32_dif(t)                                 # regs[k2, k1]


def real_split_model(regs, k_lo, k_hi):
    """|X[k]|^2 for k_lo <= k <= k_hi < 1024 from the distributed Z with the kernel's partner addressing."""
    out = np.zeros(k_hi - k_lo + 1)
    for k in range(k_lo, k_hi + 1):
        lane, k2 = k & 31, k >> 5
        zk = regs[k2, lane]
        p_lane = (32 - lane) & 31
        p_reg = (32 - k2) & 31 if lane == 0 else 31 - k2          # what lane `p_lane` supplies
        zn = regs[p_reg, p_lane]
        assert p_lane + 32 * p_reg == (1024 - k) % 1024           # ... is Z[1024 - k]
        e2 = zk + np.conj(zn)                                     # 2E
        o2 = -1j * (zk - np.conj(zn))                             # 2O
        x2 = e2 + np.exp(-1j * np.pi * k / 1024) * o2             # 2X
        out[k - k_lo] = 0.25 * abs(x2) ** 2
    return out


def test_register_layout_is_the_packed_transform():
    rng = np.random.default_rng(1)
    x = rng.integers(-32768, 32767, 2048).astype(np.float64)
    w = np.hanning(2048)
    regs = warp_frame_model(x, w)
    z = (x * w)[0::2] + 1j * (x * w)[1::2]
    ref = np.fft.fft(z)
    k1 = np.arange(32)[None, :]
    k2 = np.arange(32)[:, None]
    np.testing.assert_allclose(regs, ref[k1 + 32 * k2], rtol=0, atol=1e-6 * np.abs(ref).max())


def test_radix2_wrappers_equal_a_32_point_dft():
    rng = np.random.default_rng(2)
    v = rng.standard_normal((32, 5)) + 1j * rng.standard_normal((32, 5))
    ref = np.fft.fft(v, axis=0)
    np.testing.assert_allclose(fft32_dit(v), ref, atol=1e-12)
    np.testing.assert_allclose(fft32_dif(v), ref, atol=1e-12)


def test_real_split_and_psd_scaling_match_rfft():
    rng = np.random.default_rng(3)
    fs, nfft = 5000.0, 2048
    n = np.arange(nfft)
    x = np.round(3000 * np.sin(2 * np.pi * 1000.3 * n / fs) + 200 * rng.standard_normal(nfft))
    w = np.hanning(nfft)
    regs = warp_frame_model(x, w)
    for k_lo, k_hi in ((328, 491), (103, 327), (0, 40), (480, 700), (1000, 1023)):
        p = real_split_model(regs, k_lo, k_hi)
        ref = np.abs(np.fft.rfft(x * w)) ** 2
        np.testing.assert_allclose(p, ref[k_lo:k_hi + 1], rtol=1e-9, atol=1e-9 * ref.max())
        # one-sided density as the kernel writes it: |2X|^2 * (scale / 2), DC alone gets half of that
        scale = 1.0 / (fs * np.sum(w * w))
        k = np.arange(k_lo, k_hi + 1)
        psd = 4.0 * p * (0.5 * scale) * np.where(k == 0, 0.5, 1.0)
        want = ref[k_lo:k_hi + 1] * scale * np.where(k == 0, 1.0, 2.0)
        np.testing.assert_allclose(psd, want, rtol=1e-9, atol=1e-9 * want.max())


def test_epilogue_group_selection_covers_exactly_the_wanted_bins():
    """The epilogue walks pairs of 32-bin groups (bins 64 q + lane and 64 q + 32 + lane) and skips a pair when
    64 q + 63 < kmin or 64 q > kmax; every wanted bin must sit in a visited pair, for both unroll widths."""
    for k_lo, k_hi, n_lo, n_hi in ((328, 491, 103, 327), (0, 0, 5, 4), (500, 520, 600, 1023), (1, 1022, 2, 3)):
        have_noise = n_lo <= n_hi
        kmin = min(k_lo, n_lo) if have_noise else k_lo
        kmax = max(k_hi, n_hi) if have_noise else k_hi
        k2max = 16 if kmax < 512 else 32
        visited = set()
        for q in range(k2max // 2):
            if 64 * q + 63 < kmin or 64 * q > kmax:
                continue
            visited.update(range(64 * q, 64 * q + 64))
        wanted = set(range(k_lo, k_hi + 1)) | (set(range(n_lo, n_hi + 1)) if have_noise else set())
        assert wanted <= visited
