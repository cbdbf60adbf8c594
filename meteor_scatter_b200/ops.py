"""Torch-facing wrappers over the C-ABI (device memory + streams are torch's,
the arithmetic is the hand-written kernels in csrc/).  Everything here expects
CUDA tensors and raises otherwise: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import MsError, MsUnsupported, check, current_stream, ptr


def _cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the detection path has no CPU fallback)")
    return t.contiguous()


# --------------------------------------------------------------------------- geometry
@dataclass(frozen=True)
class BandSpec:
    """Block geometry and band bins derived with the reference's own float
    expressions (dsp/src/main.py:352-363, 382, 386)."""
    fs: int
    block_duration_sec: float
    n_fft_real: int
    block_size: int
    win_len: int
    sig_bins: tuple
    noise_bins: tuple
    window: np.ndarray = field(compare=False, repr=False)

    @staticmethod
    def from_reference_args(fs, block_duration_sec, freq_band, noise_band, n_fft) -> "BandSpec":
        n_fft_real = n_fft * 2                                        # main.py:353
        block_size = int(fs * block_duration_sec)                     # main.py:355
        freqs = np.fft.rfftfreq(n_fft_real, d=1 / fs)                 # main.py:363
        sig = np.nonzero((freqs >= freq_band[0]) & (freqs <= freq_band[1]))[0]
        noi = np.nonzero((freqs >= noise_band[0]) & (freqs <= noise_band[1]))[0]
        win_len = min(block_size, n_fft_real)                         # rfft(n=) crops or zero-pads
        window = np.hanning(block_size)[:win_len].astype(np.float64)  # main.py:379
        return BandSpec(int(fs), float(block_duration_sec), int(n_fft_real), int(block_size), int(win_len),
                        tuple(int(k) for k in sig), tuple(int(k) for k in noi), window)

    @staticmethod
    def _range(bins):
        if len(bins) == 0:
            return 1, 0               # empty mask: the sum is 0, like numpy
        assert bins[-1] - bins[0] + 1 == len(bins)
        return bins[0], bins[-1]

    @property
    def sig_range(self):
        return self._range(self.sig_bins)

    @property
    def noise_range(self):
        return self._range(self.noise_bins)

    def n_blocks(self, n_samples: int) -> int:
        if self.win_len <= self.block_size:
            return n_samples // self.block_size                       # main.py:356 (tail dropped)
        # overlapping frames (hop = block_size < window): scipy/mlab framing
        return 0 if n_samples < self.win_len else (n_samples - self.win_len) // self.block_size + 1

    @staticmethod
    def stft(nfft: int, hop: int, window, sig_bins, noise_bins, fs: int = 0) -> "BandSpec":
        """General STFT geometry (sweep config): frames of len(window) <= nfft samples every ``hop``."""
        w = np.asarray(window, dtype=np.float64)
        assert len(w) <= nfft
        return BandSpec(int(fs), hop / fs if fs else 0.0, int(nfft), int(hop), len(w),
                        tuple(int(k) for k in sig_bins), tuple(int(k) for k in noise_bins), w)

    def window_key(self):
        return hash(self.window.tobytes())


_WINDOW_CACHE = {}
_PLAN_CACHE = {}


def _window_dev(spec: BandSpec, device) -> torch.Tensor:
    key = (spec.block_size, spec.win_len, spec.window_key(), str(device))
    w = _WINDOW_CACHE.get(key)
    if w is None:
        w = torch.from_numpy(spec.window.astype(np.float32)).to(device)
        _WINDOW_CACHE[key] = w
    return w


# --------------------------------------------------------------------------- K2 plan
class DftI8Plan:
    """Device-resident digit-sliced basis for ms_band_power_i16_tc."""

    MAX_COLS = 16

    @staticmethod
    def basis_for(spec: BandSpec, part: str = "both"):
        """Windowed cos/sin columns of the band bins (fp64, host): ``basis[n, 2i] = w[n] cos(2 pi k_i n / nfft)``,
        ``basis[n, 2i+1] = w[n] sin(...)`` and the band (0 signal, 1 noise) of every column."""
        sig = list(spec.sig_bins) if part in ("both", "sig") else []
        noi = list(spec.noise_bins) if part in ("both", "noise") else []
        bins = sig + noi
        groups = [0] * len(sig) + [1] * len(noi)
        n_cols = 2 * len(bins)
        if n_cols == 0 or n_cols > DftI8Plan.MAX_COLS:
            raise MsUnsupported(-2, f"tensor-core path supports 1..8 band bins, got {len(bins)}")
        n = np.arange(spec.win_len, dtype=np.float64)
        basis = np.empty((spec.win_len, n_cols), dtype=np.float64)
        col_group = np.empty(n_cols, dtype=np.int32)
        for i, (k, g) in enumerate(zip(bins, groups)):
            ang = 2.0 * np.pi * ((k * n) % spec.n_fft_real) / spec.n_fft_real
            basis[:, 2 * i] = spec.window * np.cos(ang)
            basis[:, 2 * i + 1] = spec.window * np.sin(ang)
            col_group[2 * i] = col_group[2 * i + 1] = g
        return basis, col_group

    def __init__(self, spec: BandSpec, device, part: str = "both"):
        basis, col_group = self.basis_for(spec, part)
        n_cols = basis.shape[1]
        self.k_samples = spec.win_len
        self.n_cols = n_cols
        self.basis = basis
        self.col_group = col_group
        lib = _lib.load()
        nbytes = lib.ms_dft_i8_plan_bytes(self.k_samples, n_cols)
        if nbytes <= 0:
            raise MsUnsupported(-2, "ms_dft_i8_plan_bytes rejected the plan shape")
        self.blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
        basis_c = np.ascontiguousarray(basis)
        check(lib.ms_dft_i8_plan_build(basis_c.ctypes.data_as(C.c_void_p), col_group.ctypes.data_as(C.c_void_p),
                                       self.k_samples, n_cols, ptr(self.blob), current_stream()))

    @staticmethod
    def get(spec: BandSpec, device, part: str = "both") -> "DftI8Plan":
        key = (spec.win_len, spec.n_fft_real, spec.sig_bins, spec.noise_bins, spec.window_key(), str(device), part)
        p = _PLAN_CACHE.get(key)
        if p is None:
            p = DftI8Plan(spec, device, part)
            _PLAN_CACHE[key] = p
        return p


def k2_supported(x: torch.Tensor, spec: BandSpec) -> bool:
    """Can the resident-basis kernel (csrc/ms_dft_i8.cu) run this geometry?"""
    n_bins = len(spec.sig_bins) + len(spec.noise_bins)
    # one launch holds the cos/sin columns of up to 8 bins; wider parameter sets run the signal band and the
    # noise band as two launches as long as each band alone fits
    if not (x.dtype == torch.int16 and n_bins >= 1 and len(spec.sig_bins) <= 8 and len(spec.noise_bins) <= 8
            and (spec.block_size * 2) % 16 == 0
            and spec.win_len <= 1408):     # basis (8 KiB per 64 samples) + at least 3 stages must fit 227 KiB of smem
        return False
    flat = spec.win_len <= spec.block_size and x.dim() == 2 and x.shape[1] % spec.block_size == 0
    if x.dim() == 2 and x.shape[0] > 1 and not flat and (x.shape[1] * 2) % 16 != 0:
        return False                       # per-file launches: every file must still start 16-byte aligned (TMA)
    return x.data_ptr() % 16 == 0


def seg_supported(x: torch.Tensor, spec: BandSpec) -> bool:
    """Can the general tensor-core kernel (csrc/ms_dft_seg.cu: streamed basis, column groups, overlapping frames
    read once) run this geometry?  PCM16, hop a multiple of 8 samples, 16-byte aligned files, overlap <= 129 hops."""
    n_bins = len(spec.sig_bins) + len(spec.noise_bins)
    if not (x.dtype == torch.int16 and n_bins >= 1 and (spec.block_size * 2) % 16 == 0 and x.data_ptr() % 16 == 0):
        return False
    if x.dim() == 2 and x.shape[0] > 1 and (x.shape[1] * 2) % 16 != 0:
        return False
    return -(-spec.win_len // spec.block_size) <= 129


def float_as_pcm16(x: torch.Tensor):
    """float32 samples that are exactly PCM16 / 32768 (what ``soundfile`` and every float WAV made from 16-bit audio
    hold) as an int16 tensor, else None.  The check is exact: scaling by 2^15 does not round in fp32.  One extra pass
    over the data and one host synchronisation -- the price of putting float recordings on the exact-integer
    tensor-core path instead of the fp32 FFT kernel."""
    y = x * 32768.0
    ok = bool(((y == torch.floor(y)) & (y >= -32768.0) & (y <= 32767.0)).all().item())
    return y.to(torch.int16) if ok else None


def tc_preferred(x: torch.Tensor, spec: BandSpec) -> bool:
    """Cost model of "auto": the restricted DFT costs one pass over the audio per group of 32 bins, the FFT a fixed
    N log N.  Measured on B200 at 24 h scale (bench.py `sweep`): 109 bins (4 groups) of 16384-sample frames take 1.5 ms
    against 3.1 ms for the FFT kernel, 7 bins of 1024-sample frames 0.30 against 1.80 ms; the break-even is around
    6-8 groups."""
    return k2_supported(x, spec) or len(spec.sig_bins) + len(spec.noise_bins) <= 160


def tc_supported(x: torch.Tensor, spec: BandSpec) -> bool:
    """Is there a tensor-core (tcgen05 kind::i8) band-power path for this input?"""
    return k2_supported(x, spec) or seg_supported(x, spec)


class DftSegPlan:
    """Device-resident digit-sliced basis for ms_band_power_i16_seg: one column group (<= 32 bins) of the band
    bins, laid out per (K slab, shift) piece for hop segments of ``seg`` samples."""

    MAX_BINS = 32

    def __init__(self, spec: BandSpec, device, bins, groups, seg: int, n_shift: int):
        n = np.arange(spec.win_len, dtype=np.float64)
        basis = np.empty((spec.win_len, 2 * len(bins)), dtype=np.float64)
        col_group = np.empty(2 * len(bins), dtype=np.int32)
        for i, (k, g) in enumerate(zip(bins, groups)):
            ang = 2.0 * np.pi * ((k * n) % spec.n_fft_real) / spec.n_fft_real
            basis[:, 2 * i] = spec.window * np.cos(ang)
            basis[:, 2 * i + 1] = spec.window * np.sin(ang)
            col_group[2 * i] = col_group[2 * i + 1] = g
        self.n_frame, self.seg, self.n_shift, self.n_cols = spec.win_len, int(seg), int(n_shift), basis.shape[1]
        lib = _lib.load()
        nbytes = lib.ms_dft_seg_plan_bytes(self.n_frame, self.seg, self.n_shift, self.n_cols)
        if nbytes <= 0:
            raise MsUnsupported(-2, f"ms_dft_seg_plan_bytes rejected frame {self.n_frame}, segment {seg} x {n_shift}")
        self.blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
        basis = np.ascontiguousarray(basis)
        check(lib.ms_dft_seg_plan_build(basis.ctypes.data_as(C.c_void_p), col_group.ctypes.data_as(C.c_void_p),
                                        self.n_frame, self.seg, self.n_shift, self.n_cols, ptr(self.blob),
                                        current_stream()))

    @classmethod
    def from_basis(cls, basis: np.ndarray, col_group: np.ndarray, seg: int, n_shift: int, device) -> "DftSegPlan":
        """Plan for an explicit ``[n_frame, n_cols]`` basis (values in [-1, 1])."""
        self = cls.__new__(cls)
        self.n_frame, self.seg, self.n_shift, self.n_cols = basis.shape[0], int(seg), int(n_shift), basis.shape[1]
        lib = _lib.load()
        nbytes = lib.ms_dft_seg_plan_bytes(self.n_frame, self.seg, self.n_shift, self.n_cols)
        if nbytes <= 0:
            raise MsUnsupported(-2, f"ms_dft_seg_plan_bytes rejected frame {self.n_frame}, segment {seg} x {n_shift}")
        self.blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
        basis = np.ascontiguousarray(basis, dtype=np.float64)
        col_group = np.ascontiguousarray(col_group, dtype=np.int32)
        check(lib.ms_dft_seg_plan_build(basis.ctypes.data_as(C.c_void_p), col_group.ctypes.data_as(C.c_void_p),
                                        self.n_frame, self.seg, self.n_shift, self.n_cols, ptr(self.blob),
                                        current_stream()))
        return self

    @staticmethod
    def get(spec: BandSpec, device, bins, groups, seg: int, n_shift: int) -> "DftSegPlan":
        key = ("seg", spec.win_len, spec.n_fft_real, tuple(bins), tuple(groups), spec.window_key(), str(device),
               int(seg), int(n_shift))
        p = _PLAN_CACHE.get(key)
        if p is None:
            p = DftSegPlan(spec, device, bins, groups, seg, n_shift)
            _PLAN_CACHE[key] = p
        return p


def cosine_series(window: np.ndarray, max_order: int = 2, tol: float = 1e-12):
    """``(order, [W_0, W_1, W_2])`` if ``window[n] == sum_{|m|<=order} W_|m| e^{2 pi i m n / L}`` with real W (scipy's
    periodic 'boxcar' / 'hann' / 'hamming' / 'blackman'), else None."""
    w = np.asarray(window, dtype=np.float64)
    L = len(w)
    if L < 8:
        return None
    W = np.fft.fft(w) / L
    scale = float(np.max(np.abs(W)))
    if scale == 0.0:
        return None
    big = np.nonzero(np.abs(W) > tol * scale)[0]
    order = int(max((min(k, L - k) for k in big), default=0))
    if order > max_order or np.max(np.abs(W.imag)) > tol * scale:
        return None
    return order, [float(W[m].real) if m <= order else 0.0 for m in range(3)]


def rot_supported(x: torch.Tensor, spec: BandSpec) -> bool:
    """Overlapping frames whose window is a cosine series of the frame length == nfft and whose hop divides the frame:
    one unwindowed tensor-core product per hop segment + ms_window_combine."""
    if not (seg_supported(x, spec) and spec.block_size < spec.win_len == spec.n_fft_real
            and spec.win_len % spec.block_size == 0):
        return False
    cs = cosine_series(spec.window)
    if cs is None:
        return False
    n_ext = sum(len(b) + 2 * cs[0] for b in (spec.sig_bins, spec.noise_bins) if len(b))
    return 0 < n_ext <= 160


def _band_power_rot(lib, x, spec: BandSpec, nb: int, band_db, noise_db, be, ne, st):
    """Launch plan of the frequency-domain-window form (see ms_b200.h: ms_dft_seg_projections_i16 / ms_window_combine)."""
    n_files, spf = x.shape
    H, L, nfft = spec.block_size, spec.win_len, spec.n_fft_real
    R = L // H
    order, coef = cosine_series(spec.window)
    ext, ranges = [], []
    for b in (spec.sig_bins, spec.noise_bins):
        if len(b):
            ranges.append((len(ext) + order, len(b)))
            ext += list(range(b[0] - order, b[-1] + order + 1))
        else:
            ranges.append((order, 0))
    E = len(ext)
    rows = spf // H
    assert nb == rows - R + 1
    key = ("rot", H, nfft, tuple(ext), R, str(x.device))
    cached = _PLAN_CACHE.get(key)
    if cached is None:
        i = np.arange(H, dtype=np.float64)
        plans = []
        for g0 in range(0, E, DftSegPlan.MAX_BINS):
            ks = ext[g0:g0 + DftSegPlan.MAX_BINS]
            basis = np.empty((H, 2 * len(ks)), dtype=np.float64)
            for c, k in enumerate(ks):
                ang = 2.0 * np.pi * ((k * i) % nfft) / nfft
                basis[:, 2 * c] = np.cos(ang)
                basis[:, 2 * c + 1] = np.sin(ang)
            plans.append((2 * g0, DftSegPlan.from_basis(basis, np.zeros(2 * len(ks), dtype=np.int32), H, 1, x.device)))
        rot = np.empty((R, E, 2), dtype=np.float64)
        for j in range(R):
            for e, k in enumerate(ext):
                ang = 2.0 * np.pi * ((k * j * H) % nfft) / nfft
                rot[j, e] = (np.cos(ang), np.sin(ang))
        cached = _PLAN_CACHE[key] = (plans, torch.from_numpy(rot).to(x.device))
    plans, rot_d = cached
    proj = torch.empty((n_files, rows, 2 * E), dtype=torch.float64, device=x.device)
    fstride = spf * 2 if n_files > 1 else 16
    for col0, plan in plans:
        check(lib.ms_dft_seg_projections_i16(ptr(x), n_files, fstride, rows, H * 2, ptr(plan.blob), H, plan.n_cols, rows,
                                             ptr(proj), 2 * E, col0, st))
    hc = (C.c_double * 3)(*coef)
    check(lib.ms_window_combine(ptr(proj), ptr(rot_d), n_files, rows, nb, E, R, order, hc, ranges[0][0], ranges[0][1],
                                ranges[1][0], ranges[1][1], nb, ptr(band_db), ptr(noise_db), ptr(be), ptr(ne), st))


def _band_power_seg(lib, x, spec: BandSpec, nb: int, band_db, noise_db, st, want_energy: bool = False):
    """Launch plan of the general tensor-core kernel for ``x`` [n_files, spf] PCM16 (see ms_b200.h)."""
    n_files, spf = x.shape
    hop, frame = spec.block_size, spec.win_len
    bins = list(spec.sig_bins) + list(spec.noise_bins)
    groups = [0] * len(spec.sig_bins) + [1] * len(spec.noise_bins)
    chunks = [(bins[i:i + DftSegPlan.MAX_BINS], groups[i:i + DftSegPlan.MAX_BINS])
              for i in range(0, len(bins), DftSegPlan.MAX_BINS)]
    acc_b = acc_n = None
    if len(chunks) > 1 or want_energy:   # bands wider than one launch: fp64 energies accumulate across column groups
        acc_b = torch.empty((n_files, nb), dtype=torch.float64, device=x.device)
        acc_n = torch.empty((n_files, nb), dtype=torch.float64, device=x.device)
    fstride = spf * 2 if n_files > 1 else 16
    es = x.element_size()
    # frames [0, n_seg): hop segments read once (every segment row they touch is a whole row inside the file);
    # frames [n_seg, nb): the ragged tail (and everything when hop >= frame) as rows = frames
    n_shift = -(-frame // hop)
    n_seg = 0
    if hop < frame:
        n_seg = max(0, min(nb, spf // hop - n_shift + 1))
    for ci, (cb, cg) in enumerate(chunks):
        first, last = int(ci == 0), int(ci == len(chunks) - 1)
        if n_seg > 0:
            plan = DftSegPlan.get(spec, x.device, cb, cg, hop, n_shift)
            check(lib.ms_band_power_i16_seg(ptr(x), n_files, fstride, spf // hop, hop * 2, n_seg, ptr(plan.blob),
                                            frame, hop, n_shift, plan.n_cols, nb, 0, ptr(band_db), ptr(noise_db),
                                            ptr(acc_b), ptr(acc_n), first, last, st))
        if n_seg < nb:
            plan = DftSegPlan.get(spec, x.device, cb, cg, frame, 1)
            base = C.c_void_p(x.data_ptr() + n_seg * hop * es)
            check(lib.ms_band_power_i16_seg(base, n_files, fstride, nb - n_seg, hop * 2, nb - n_seg, ptr(plan.blob),
                                            frame, frame, 1, plan.n_cols, nb, n_seg, ptr(band_db), ptr(noise_db),
                                            ptr(acc_b), ptr(acc_n), first, last, st))
    return acc_b, acc_n


# --------------------------------------------------------------------------- A-stft
def band_power(x: torch.Tensor, spec: BandSpec, impl: str = "auto", want_energy: bool = False, out=None):
    """STFT band power of a batch of recordings.

    x: ``[n_files, samples_per_file]`` int16 or float32 CUDA tensor.
    Returns (band_db, noise_db[, band_energy, noise_energy]) float32 ``[n_files, n_blocks]``.
    impl: "fft" (K1), "tc" (tcgen05 kind::i8: "k2" = resident basis, csrc/ms_dft_i8.cu, for non-overlapping frames
    of <= 1408 samples and <= 8 bins per band; "seg" = the general kernel, csrc/ms_dft_seg.cu, otherwise) or "auto"
    (tc when supported and the bands are narrow enough for the restricted DFT to beat the FFT).
    """
    lib = _lib.load()
    x = _cuda(x, "x")
    if x.dim() == 1:
        x = x.unsqueeze(0)
    if x.dtype not in (torch.int16, torch.float32):
        raise ValueError(f"unsupported sample dtype {x.dtype}; expected int16 or float32")
    n_files, spf = x.shape
    nb = spec.n_blocks(spf)
    dev = x.device
    if out is not None:      # reuse caller-owned [n_files, n_blocks] float32 buffers (steady-state batches)
        band_db, noise_db = out
        assert band_db.shape == (n_files, nb) and noise_db.shape == (n_files, nb) and band_db.is_contiguous()
    else:
        band_db = torch.empty((n_files, nb), dtype=torch.float32, device=dev)
        noise_db = torch.empty((n_files, nb), dtype=torch.float32, device=dev)
    be = torch.empty((n_files, nb), dtype=torch.float32, device=dev) if want_energy else None
    ne = torch.empty((n_files, nb), dtype=torch.float32, device=dev) if want_energy else None
    ret = (band_db, noise_db, be, ne) if want_energy else (band_db, noise_db)
    if n_files == 0 or nb == 0:
        return ret
    if x.dtype == torch.float32 and impl in ("auto", "tc", "k2", "seg", "rot"):
        # float recordings: exact PCM16 / 32768 values run on the integer tensor-core path with the 2^-15 folded into
        # the window (the plan normalises its basis to the peak, so no precision is lost); anything else -> FFT kernel
        xi = float_as_pcm16(x)
        if xi is not None:
            import dataclasses
            spec_i = dataclasses.replace(spec, window=spec.window / 32768.0)
            if tc_supported(xi, spec_i) and (impl != "auto" or tc_preferred(xi, spec_i)):
                return band_power(xi, spec_i, impl="tc" if impl == "auto" else impl, want_energy=want_energy, out=out)
        if impl != "auto":
            raise MsUnsupported(-2, "tensor-core band power of float32 input needs samples that are exactly "
                                    "PCM16 / 32768 (and a TMA-addressable geometry); use impl='fft'")
        impl = "fft"
    if impl == "auto":
        impl = "tc" if tc_supported(x, spec) and tc_preferred(x, spec) else "fft"
    st = current_stream()
    if impl == "tc":         # tensor cores: the resident-basis kernel where it fits, else the general ones
        if k2_supported(x, spec) and (spec.win_len <= spec.block_size or not seg_supported(x, spec)):
            impl = "k2"
        elif rot_supported(x, spec) and spec.win_len // spec.block_size >= 4:
            # measured (24 h sweep): from 75 % overlap on, one unwindowed product per hop segment + the frequency-domain
            # window beats the shifted products (0.41 vs 0.53 ms at nfft 1024, 1.6 vs 2.5 ms at 16384); at 50 % the
            # shifted form is as fast or faster (0.27 vs 0.34 ms at nfft 2048)
            impl = "rot"
        elif seg_supported(x, spec):
            impl = "seg"
        else:
            raise MsUnsupported(-2, "tensor-core band power needs 16-byte aligned int16 input and a hop that is a "
                                    "multiple of 8 samples")
    if impl == "rot":
        if not rot_supported(x, spec):
            raise MsUnsupported(-2, "frequency-domain-window form: overlapping frames, frame length == nfft, hop | frame, "
                                    "window = periodic boxcar / hann / hamming / blackman, at most 160 extended bins")
        _band_power_rot(lib, x, spec, nb, band_db, noise_db, be, ne, st)
        return ret
    if impl == "seg":
        if not seg_supported(x, spec):
            raise MsUnsupported(-2, "general tensor-core band power needs 16-byte aligned int16 input, a hop that "
                                    "is a multiple of 8 samples and at most 129 hops per frame")
        acc = _band_power_seg(lib, x, spec, nb, band_db, noise_db, st, want_energy)
        if want_energy:      # the fp64 energy accumulators of the kernel, narrowed like the other paths' outputs
            be.copy_(acc[0])
            ne.copy_(acc[1])
        return ret
    if impl == "k2":
        if not k2_supported(x, spec):
            raise MsUnsupported(-2, "resident-basis tensor-core band power needs int16 input, <= 8 bins per band, "
                                    "frames of at most 1408 samples and a block size that is a multiple of 8 samples")
        stride_b = spec.block_size * 2
        flat = spf == nb * spec.block_size and spec.win_len <= spec.block_size
        if not flat and not (n_files == 1 or (spf * 2) % 16 == 0):
            raise MsUnsupported(-2, "samples_per_file must be a multiple of 8 for the tensor-core path")

        def launch(plan, o_band, o_noise, o_be, o_ne):
            if flat:
                check(lib.ms_band_power_i16_tc(ptr(x), n_files * nb, stride_b, ptr(plan.blob), plan.k_samples,
                                               plan.n_cols, ptr(o_band), ptr(o_noise), ptr(o_be), ptr(o_ne), st))
            else:   # ragged tail or overlapping frames: rank-3 tensor map [file][frame][bytes], one launch
                check(lib.ms_band_power_i16_tc_batched(ptr(x), n_files, spf * 2 if n_files > 1 else 16, nb, stride_b,
                                                       ptr(plan.blob), plan.k_samples, plan.n_cols, nb, ptr(o_band),
                                                       ptr(o_noise), ptr(o_be), ptr(o_ne), st))

        if len(spec.sig_bins) + len(spec.noise_bins) <= 8:
            launch(DftI8Plan.get(spec, dev), band_db, noise_db, be, ne)
        else:
            # two launches, each writing its own band; the other band's output of a launch goes to scratch
            scratch = torch.empty_like(band_db)
            if len(spec.sig_bins):
                launch(DftI8Plan.get(spec, dev, "sig"), band_db, scratch, be, None)
            else:
                band_db.fill_(-120.0)
                if be is not None:
                    be.zero_()
            if len(spec.noise_bins):
                launch(DftI8Plan.get(spec, dev, "noise"), scratch, noise_db, None, ne)
            else:
                noise_db.fill_(-120.0)
                if ne is not None:
                    ne.zero_()
        return ret
    if impl != "fft":
        raise ValueError(f"unknown impl {impl!r}")
    fn = lib.ms_band_power_i16 if x.dtype == torch.int16 else lib.ms_band_power_f32
    (slo, shi), (nlo, nhi) = spec.sig_range, spec.noise_range
    check(fn(ptr(x), n_files, spf, nb, spec.block_size, spec.win_len, ptr(_window_dev(spec, dev)), spec.n_fft_real,
             slo, shi, nlo, nhi, nb, ptr(band_db), ptr(noise_db), ptr(be), ptr(ne), st))
    return ret


def ingest_rows(host: torch.Tensor, block_size: int, used: int, out: torch.Tensor, stream=None):
    """Host -> device copy of the first ``used`` samples of every ``block_size``-sample block.

    host: pinned CPU tensor ``[n_files, n_blocks*block_size]`` (int16 or float32);
    out:  CUDA tensor ``[n_files, n_blocks*used]`` of the same dtype (dense rows).
    One strided DMA; returns ``out``."""
    lib = _lib.load()
    assert not host.is_cuda and out.is_cuda and host.dtype == out.dtype and host.is_contiguous() and out.is_contiguous()
    n_files, spf = host.shape
    nb = spf // block_size
    assert spf == nb * block_size and out.shape == (n_files, nb * used) and used <= block_size
    es = host.element_size()
    check(lib.ms_ingest_rows_h2d(C.c_void_p(host.data_ptr()), n_files * nb, block_size * es, used * es, ptr(out),
                                 used * es, current_stream() if stream is None else C.c_void_p(stream.cuda_stream)))
    return out


# --------------------------------------------------------------------------- A-thr
@dataclass
class DetectResult:
    events: torch.Tensor          # [n_files, max_events, 2] int32 (start, stop_exclusive)
    event_db: torch.Tensor        # [n_files, max_events] float64
    counts: torch.Tensor          # [n_files] int32
    thresholds: torch.Tensor | None
    near: torch.Tensor | None

    def check_capacity(self):
        """Raise if any file produced more events than ``max_events`` (the kernel keeps counting but stores,
        and counts into the hourly histogram, only the first ``max_events`` of a file).  Synchronises."""
        cap = self.events.shape[1]
        if self.counts.numel() and int(self.counts.max().item()) > cap:
            raise RuntimeError(f"event capacity exceeded: a file produced {int(self.counts.max().item())} events, "
                               f"max_events={cap}; re-run with a larger max_events")


def detect(band_db: torch.Tensor, noise_db: torch.Tensor, k_std: float, adaptive: bool = True,
           window_blocks: int = 600, before_blocks: int = 15, after_blocks: int = 100, fixed_blocks: int = 50,
           n_blocks_per_file: torch.Tensor | None = None, max_events: int = 256, want_thresholds: bool = False,
           want_near: bool = False, eps_db: float = 1e-3, workspace: torch.Tensor | None = None,
           out: DetectResult | None = None, hourly: dict | None = None) -> DetectResult:
    """``hourly`` = dict(file_start_us=int64 tensor, block_duration_sec, hour0, n_hours, out=[n_hours,2] int32
    tensor, crit_min_dur_sec=0.5) counts every event into ``out`` in the same launch (adaptive detector) or
    with one extra small kernel (global detector).  ``out`` reuses a previous DetectResult's buffers."""
    lib = _lib.load()
    band_db = _cuda(band_db, "band_db")
    noise_db = _cuda(noise_db, "noise_db")
    if band_db.dtype != torch.float32 or noise_db.dtype != torch.float32 or band_db.shape != noise_db.shape:
        raise ValueError("band_db / noise_db must be float32 tensors of the same [n_files, n_blocks] shape")
    if band_db.dim() == 1:
        band_db, noise_db = band_db.unsqueeze(0), noise_db.unsqueeze(0)
    n_files, nb = band_db.shape
    dev = band_db.device
    if out is not None and out.events.shape == (n_files, max_events, 2):
        events, event_db, counts = out.events, out.event_db, out.counts
    else:   # only entries < counts[f] are meaningful; the kernel writes counts for every file
        events = torch.empty((n_files, max_events, 2), dtype=torch.int32, device=dev)
        event_db = torch.empty((n_files, max_events), dtype=torch.float64, device=dev)
        counts = torch.empty((n_files,), dtype=torch.int32, device=dev)
    thr = torch.full((n_files, nb), float("nan"), dtype=torch.float64, device=dev) if want_thresholds else None
    near = torch.zeros((n_files, nb), dtype=torch.uint8, device=dev) if want_near else None
    need = lib.ms_detect_workspace_bytes(n_files, nb)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    npf = None
    if n_blocks_per_file is not None:
        npf = _cuda(n_blocks_per_file.to(torch.int32), "n_blocks_per_file")
    st = current_stream()
    if n_files == 0 or nb == 0:      # nothing to scan: no events (empty tensors have no device pointer)
        counts.zero_()
        return DetectResult(events, event_db, counts, thr, near)
    if hourly is not None:
        h_us = _cuda(hourly["file_start_us"], "file_start_us")
        assert h_us.dtype == torch.int64 and h_us.numel() == n_files
        h_out = hourly["out"]
        assert h_out.is_cuda and h_out.dtype == torch.int32 and h_out.shape == (hourly["n_hours"], 2)
    if adaptive and hourly is not None:
        check(lib.ms_detect_adaptive_hourly(
            ptr(band_db), ptr(noise_db), n_files, nb, nb, ptr(npf), float(k_std), int(window_blocks),
            int(before_blocks), int(after_blocks), int(fixed_blocks), int(max_events), ptr(events), ptr(event_db),
            ptr(counts), ptr(thr), ptr(near), float(eps_db), ptr(workspace), workspace.numel(), ptr(h_us),
            float(hourly["block_duration_sec"]), float(hourly.get("crit_min_dur_sec", 0.5)), int(hourly["hour0"]),
            int(hourly["n_hours"]), ptr(h_out), 1 if hourly.get("small_footprint") else 0, st))
    elif adaptive:
        check(lib.ms_detect_adaptive(ptr(band_db), ptr(noise_db), n_files, nb, nb, ptr(npf), float(k_std),
                                     int(window_blocks), int(before_blocks), int(after_blocks), int(fixed_blocks),
                                     int(max_events), ptr(events), ptr(event_db), ptr(counts), ptr(thr), ptr(near),
                                     float(eps_db), ptr(workspace), workspace.numel(), st))
    else:
        check(lib.ms_detect_global(ptr(band_db), ptr(noise_db), n_files, nb, nb, ptr(npf), float(k_std),
                                   int(max_events), ptr(events), ptr(event_db), ptr(counts), ptr(thr), ptr(near),
                                   float(eps_db), ptr(workspace), workspace.numel(), st))
        if hourly is not None:
            check(lib.ms_hourly_counts(ptr(events), ptr(counts), n_files, max_events, ptr(h_us),
                                       float(hourly["block_duration_sec"]), float(hourly.get("crit_min_dur_sec", 0.5)),
                                       int(hourly["hour0"]), int(hourly["n_hours"]), ptr(h_out), st))
    return DetectResult(events, event_db, counts, thr, near)


def hourly_counts(events: torch.Tensor, counts: torch.Tensor, file_start_us: torch.Tensor,
                  block_duration_sec: float, hour0: int, n_hours: int, crit_min_dur_sec: float = 0.5,
                  out: torch.Tensor | None = None) -> torch.Tensor:
    """[n_hours, 2] int32 (Anzahl, Kritisch); accumulates into ``out`` when given."""
    lib = _lib.load()
    events = _cuda(events, "events")
    counts = _cuda(counts, "counts")
    file_start_us = _cuda(file_start_us.to(torch.int64), "file_start_us")
    n_files, max_events, _ = events.shape
    if out is None:
        out = torch.zeros((n_hours, 2), dtype=torch.int32, device=events.device)
    check(lib.ms_hourly_counts(ptr(events), ptr(counts), n_files, max_events, ptr(file_start_us),
                               float(block_duration_sec), float(crit_min_dur_sec), int(hour0), int(n_hours),
                               ptr(out), current_stream()))
    return out


# --------------------------------------------------------------------------- B
_QF_CACHE = {}


class WelchQuadform:
    """Low-rank basis of the Welch band powers (see csrc/ms_welch_qf.cu): for each of the three bands the leading
    eigenpairs of P Q P, with Q[n,m] = w_n w_m sum_k c_k cos(2 pi k (n-m)/nfft) over the band's bins (c_k = 2 except
    DC/Nyquist, the one-sided PSD doubling) and P the per-segment mean removal of scipy's detrend='constant'.
    Eigenvalues below ``cut`` * lambda_max are dropped (1e-10: 26 columns for the reference's 102-bin bands)."""

    MAX_COLS = 32
    TC_COLS = 26        # columns per band of the tensor-core kernel (csrc/ms_welch_i8.cu)
    TC_TAIL = 1e-8      # largest dropped eigenvalue (relative) the tensor-core form accepts

    def __init__(self, nperseg: int, nfft: int, bands, fs: float, n_sub: int, device, cut: float = 1e-10):
        n = np.arange(nperseg, dtype=np.float64)
        w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / nperseg)                 # scipy get_window('hann') periodic
        scale = 1.0 / (fs * float(np.sum(w * w)))
        P = np.eye(nperseg) - np.ones((nperseg, nperseg)) / nperseg
        d = n[:, None] - n[None, :]
        basis = np.zeros((nperseg, self.MAX_COLS, 4), dtype=np.float32)
        self.group_scale = (C.c_double * 3)()
        self.ranks = []
        self.nperseg, self.device = nperseg, device
        self._cols64 = np.zeros((3, self.TC_COLS, nperseg), dtype=np.float64)   # [band][column][sample]
        self._tc_tail = 0.0
        self._tc_plan = None
        for g, (lo, hi) in enumerate(bands):
            if hi < lo:                                                   # empty mask: power 0 -> -inf dB
                self.group_scale[g] = 0.0
                self.ranks.append(0)
                continue
            k = np.arange(lo, hi + 1, dtype=np.float64)
            ck = np.where((k == 0) | (k == nfft // 2), 1.0, 2.0)
            q = np.zeros((nperseg, nperseg))
            for kk, cc in zip(k, ck):                                     # O(bins * K^2) once per parameter set
                q += cc * np.cos(2.0 * np.pi * kk * d / nfft)
            q = P @ (q * np.outer(w, w)) @ P
            lam, u = np.linalg.eigh(q)
            lam, u = lam[::-1], u[:, ::-1]
            keep = int(np.sum(lam > cut * lam[0]))
            if keep > self.MAX_COLS:
                raise MsUnsupported(-2, f"band {g} needs {keep} quadratic-form columns (> {self.MAX_COLS}); "
                                        "use the FFT path")
            basis[:, :keep, g] = (u[:, :keep] * np.sqrt(lam[:keep] / lam[0])).astype(np.float32)
            self.group_scale[g] = float(lam[0]) * scale / n_sub
            self.ranks.append(keep)
            kt = min(keep, self.TC_COLS)
            self._cols64[g, :kt] = (u[:, :kt] * np.sqrt(lam[:kt] / lam[0])).T
            if nperseg > kt:
                self._tc_tail = max(self._tc_tail, float(lam[kt] / lam[0]) if keep > kt else 0.0)
        self.basis = torch.from_numpy(basis.reshape(nperseg, self.MAX_COLS * 4)).to(device)

    def tc_ok(self) -> bool:
        """Can the tensor-core kernel represent this parameter set (26 columns per band are enough)?"""
        return self._tc_tail <= self.TC_TAIL and self.nperseg % 64 == 0 and \
            _lib.load().ms_welch_i8_plan_bytes(self.nperseg) > 0

    def tc_plan(self) -> torch.Tensor:
        """Device image for ms_welch_band_db_i8_i16 (built once per parameter set)."""
        if self._tc_plan is None:
            lib = _lib.load()
            blob = torch.empty(lib.ms_welch_i8_plan_bytes(self.nperseg), dtype=torch.uint8, device=self.device)
            cols = np.ascontiguousarray(self._cols64)
            check(lib.ms_welch_i8_plan_build(cols.ctypes.data_as(C.c_void_p), self.nperseg, self.TC_COLS, ptr(blob),
                                             current_stream()))
            self._tc_plan = blob
        return self._tc_plan

    @staticmethod
    def get(nperseg, nfft, bands, fs, n_sub, device) -> "WelchQuadform":
        key = (nperseg, nfft, tuple(tuple(b) for b in bands), float(fs), n_sub, str(device))
        q = _QF_CACHE.get(key)
        if q is None:
            q = WelchQuadform(nperseg, nfft, bands, fs, n_sub, device)
            _QF_CACHE[key] = q
        return q


def welch_band_db(x: torch.Tensor, block: int, nfft: int, bands, fs: float, nperseg: int = 256, rows=None,
                  impl: str = "auto"):
    """Per-block Welch band dB: returns ``[n_streams, n_blocks, 4]`` float32
    (ms_dB, noise1_dB, noise2_dB, db2).  int16 input is scaled by 1/32768 first
    (soundfile semantics, processor.py:65-71).  ``rows=(k_lo, k_hi)`` additionally returns the per-bin PSD
    in dB of those bins, ``[n_streams, n_blocks, k_hi-k_lo+1]`` (the reference's waterfall rows).
    impl: "tc" = low-rank quadratic form on the tensor cores (csrc/ms_welch_i8.cu; PCM16), "qf" = the same form on
    CUDA cores (csrc/ms_welch_qf.cu), "fft" = K1 in Welch mode; "auto" = tc, else qf, else fft (waterfall rows need
    per-bin PSDs and always take the FFT path)."""
    lib = _lib.load()
    x = _cuda(x, "x")
    if x.dim() == 1:
        x = x.unsqueeze(0)
    n_streams, n = x.shape
    nb = 0 if n < block else (n - block) // block + 1                # processor.py:176
    nperseg = min(nperseg, block)
    out = torch.empty((n_streams, nb, 4), dtype=torch.float32, device=x.device)
    out_rows = None
    if rows is not None:
        out_rows = torch.empty((n_streams, nb, rows[1] - rows[0] + 1), dtype=torch.float32, device=x.device)
    if nb == 0 or n_streams == 0:
        return out if rows is None else (out, out_rows)
    if x.dtype not in (torch.int16, torch.float32):
        raise ValueError(f"unsupported sample dtype {x.dtype}")
    hop = nperseg - nperseg // 2
    n_sub = (block - nperseg // 2) // hop
    qf_ok = rows is None and nperseg % 4 == 0 and hop % 4 == 0 and 1 <= n_sub <= 8
    # tensor-core form: PCM16, TMA-addressable geometry (16-byte multiples), shared-memory-sized basis
    # (a single stream needs no aligned stream stride: only its nb whole blocks are addressed)
    tc_stride = n if n_streams > 1 else nb * block
    tc_geom = (qf_ok and x.dtype == torch.int16 and x.is_contiguous() and nperseg % 64 == 0 and (hop * 2) % 16 == 0
               and (block * 2) % 16 == 0 and (tc_stride * 2) % 16 == 0 and x.data_ptr() % 16 == 0)
    if impl == "auto":
        impl = "fft"
        if qf_ok:
            try:        # bands too wide for the low-rank form (> 32 columns) stay on the FFT path
                impl = "tc" if tc_geom and WelchQuadform.get(nperseg, nfft, bands, fs, n_sub, x.device).tc_ok() else "qf"
            except MsUnsupported:
                impl = "fft"
    if impl == "tc":
        qf = WelchQuadform.get(nperseg, nfft, bands, fs, n_sub, x.device) if tc_geom else None
        if qf is None or not qf.tc_ok():
            raise MsUnsupported(-2, "tensor-core Welch kernel: contiguous PCM16, nperseg % 64 == 0, hop/block/stream "
                                    "stride multiples of 16 bytes, <= 26 quadratic-form columns per band")
        check(lib.ms_welch_band_db_i8_i16(ptr(x), n_streams, tc_stride, nb, int(block), int(nperseg), ptr(qf.tc_plan()),
                                          qf.group_scale, ptr(out), current_stream()))
        return out
    if impl == "qf":
        if not qf_ok:
            raise MsUnsupported(-2, "quadratic-form Welch kernel: no waterfall rows, nperseg % 8 == 0, <= 8 segments")
        qf = WelchQuadform.get(nperseg, nfft, bands, fs, n_sub, x.device)
        fnq = lib.ms_welch_band_db_qf_i16 if x.dtype == torch.int16 else lib.ms_welch_band_db_qf_f32
        check(fnq(ptr(x), n_streams, n, nb, int(block), int(nperseg), ptr(qf.basis), qf.group_scale, ptr(out),
                  current_stream()))
        return out
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(nperseg) / nperseg)   # scipy get_window('hann') periodic
    scale = 1.0 / (fs * float(np.sum(w * w)))
    wd = _WINDOW_CACHE.get(("welch", nperseg, str(x.device)))            # cached: no H2D copy inside a graph capture
    if wd is None:
        wd = _WINDOW_CACHE[("welch", nperseg, str(x.device))] = torch.from_numpy(w.astype(np.float32)).to(x.device)
    hb = (C.c_int32 * 6)(*[int(v) for pair in bands for v in pair])
    fn = {torch.int16: lib.ms_welch_band_db_i16, torch.float32: lib.ms_welch_band_db_f32}.get(x.dtype)
    check(fn(ptr(x), n_streams, n, nb, int(block), int(nperseg), ptr(wd), int(nfft), hb, scale, ptr(out),
             int(rows[0]) if rows else 0, int(rows[1]) if rows else 0, ptr(out_rows), current_stream()))
    return out if rows is None else (out, out_rows)


class LiveStates:
    """Device-resident per-stream state of the live detector (resumable)."""

    def __init__(self, n_streams: int, device, max_det: int = 4096):
        self.n_streams = n_streams
        self.max_det = max_det
        self.buf = torch.zeros((n_streams, C.sizeof(_lib.LiveState)), dtype=torch.uint8, device=device)
        self.det = torch.zeros((n_streams, max_det, 7), dtype=torch.float64, device=device)
        self.det_count = torch.zeros((n_streams,), dtype=torch.int32, device=device)


def live_state_step(states: LiveStates, cfg: "_lib.LiveConfig", db2: torch.Tensor, want_thresholds: bool = False):
    """Advance every stream by ``db2.shape[-1]`` blocks.  db2: [n_streams, n] float32
    (any element stride, e.g. column 3 of welch_band_db's output)."""
    lib = _lib.load()
    if not db2.is_cuda or db2.dtype != torch.float32:
        raise ValueError("db2 must be a float32 CUDA tensor")
    if db2.dim() == 1:
        db2 = db2.unsqueeze(0)
    n_streams, n = db2.shape
    assert n_streams == states.n_streams
    s0 = db2.stride(0) if n_streams > 1 else max(db2.stride(0), 0)
    thr = torch.empty((n_streams, n), dtype=torch.float64, device=db2.device) if want_thresholds else None
    need = int(lib.ms_live_state_workspace_bytes(n_streams, n))     # batch form only; torch's allocator caches it
    ws = torch.empty(need, dtype=torch.uint8, device=db2.device) if need else None
    check(lib.ms_live_state_step_ws(ptr(states.buf), C.byref(cfg), n_streams, ptr(db2), int(s0), int(db2.stride(1)),
                                    n, states.max_det, ptr(states.det), ptr(states.det_count), ptr(thr), ptr(ws),
                                    need, current_stream()))
    return thr


# --------------------------------------------------------------------------- C
def psd_spectrogram(x: torch.Tensor, fs: float, nfft: int, noverlap: int, window: np.ndarray, k_lo: int, k_hi: int,
                    k_noise_lo: int, k_noise_hi: int):
    """One-sided PSD rows ``[n_seg, k_hi-k_lo+1, n_frames]`` (float32) and the
    noise-band PSD sum over time and frequency ``[n_seg]`` (float64)."""
    lib = _lib.load()
    x = _cuda(x, "x")
    if x.dim() == 1:
        x = x.unsqueeze(0)
    n_seg, n = x.shape
    hop = nfft - noverlap
    n_frames = 0 if n < nfft else (n - noverlap) // hop
    w = np.asarray(window, dtype=np.float64)
    assert len(w) == nfft
    w32 = np.ascontiguousarray(w, dtype=np.float32)
    key = ("psd", nfft, w32.tobytes(), str(x.device))                   # same window again: no H2D copy per call
    hit = _WINDOW_CACHE.get(key)
    if hit is None:
        if len(_WINDOW_CACHE) > 64:
            _WINDOW_CACHE.clear()
        hit = _WINDOW_CACHE[key] = (torch.from_numpy(w32).to(x.device), float(np.sum(w * w)))
    wd, scale = hit[0], 1.0 / (fs * hit[1])
    out = torch.empty((n_seg, k_hi - k_lo + 1, n_frames), dtype=torch.float32, device=x.device)
    noise = torch.zeros((n_seg,), dtype=torch.float64, device=x.device)
    if n_frames == 0 or n_seg == 0:
        return out, noise
    fn = {torch.int16: lib.ms_psd_spectrogram_i16, torch.float32: lib.ms_psd_spectrogram_f32}.get(x.dtype)
    if fn is None:
        raise ValueError(f"unsupported sample dtype {x.dtype}")
    check(fn(ptr(x), n_seg, n, n_frames, hop, nfft, ptr(wd), scale, int(k_lo), int(k_hi), int(k_noise_lo),
             int(k_noise_hi), ptr(out), ptr(noise), current_stream()))
    return out, noise
