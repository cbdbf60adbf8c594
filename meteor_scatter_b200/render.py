"""Image export of GPU-produced spectrograms, off the hot path.

The reference renders with matplotlib, which is not needed here: this module
turns a device dB matrix into the same kind of picture with numpy + Pillow.

* ``spectrogram_image`` / ``save_spectrogram_jpg`` -- what detector C feeds to
  ``detect_and_cluster_bursts``: ``plt.imshow(Pxx_db, aspect='auto',
  origin='lower', vmin=temp_vmin, vmax=40)``, ``plt.ylim(800, 1200)``, axes off,
  ``savefig(.., format='jpg', bbox_inches='tight', pad_inches=0)``
  (meteor_detect_class/prime_detection.py:96-105).  matplotlib's default figure
  (6.4 x 4.8 in at 100 dpi, axes 0.775 x 0.77 of it) makes that a 496 x 370 px
  image -- the pixel grid ``detect_and_cluster_bursts`` assumes in its own tick
  labels (detector_and_classification.py:76-80: x 0..495, y 0..365) and in its
  "critical = bounding box >= 5 px wide" rule.
* ``event_figure`` -- detector A's per-event ``spec_and_psd`` picture
  (dsp/src/main.py:40-124, 721-806): spectrogram crop on the left 70 %, Welch PSD
  of the crop on the right 30 %, red dashed marker lines at the event limits.

Colours follow matplotlib's default colormap (viridis) through a degree-6
polynomial fit (within 0.02 of the 256-entry table per channel).
"""
from __future__ import annotations

import numpy as np

C_IMAGE_SIZE = (496, 370)           # (width, height) of the reference's saved axes area

_VIRIDIS = np.array([
    [0.2777273272234177, 0.005407344544966578, 0.3340998053353061],
    [0.1050930431085774, 1.404613529898575, 1.384590162594685],
    [-0.3308618287255563, 0.214847559468213, 0.09509516302823659],
    [-4.634230498983486, -5.799100973351585, -19.33244095627987],
    [6.228269936347081, 14.17993336680509, 56.69055260068105],
    [4.776384997670288, -13.74514537774601, -65.35303263337234],
    [-5.435455855934631, 4.645852612178535, 26.3124352495832]])


def viridis(t: np.ndarray) -> np.ndarray:
    """``t`` in [0, 1] (any shape) -> uint8 RGB array ``t.shape + (3,)``."""
    t = np.clip(np.asarray(t, dtype=np.float64), 0.0, 1.0)[..., None]
    acc = np.zeros(t.shape[:-1] + (3,))
    for c in _VIRIDIS[::-1]:
        acc = acc * t + c
    return np.clip(np.rint(acc * 255.0), 0, 255).astype(np.uint8)


def _to_numpy(a):
    if hasattr(a, "detach"):
        a = a.detach().float().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def spectrogram_image(db, vmin: float, vmax: float, size=C_IMAGE_SIZE):
    """dB matrix ``[n_freq, n_time]`` (row 0 = lowest frequency, numpy or CUDA tensor) -> ``PIL.Image`` of ``size``.

    Normalisation is matplotlib's ``Normalize(vmin, vmax)`` with clipping at both ends (``-inf`` maps to the bottom
    colour); ``origin='lower'`` puts row 0 at the bottom; ``aspect='auto'`` stretches the matrix over the whole axes,
    here with bilinear resampling."""
    from PIL import Image
    a = _to_numpy(db)
    assert a.ndim == 2 and a.shape[0] > 0 and a.shape[1] > 0
    t = (np.nan_to_num(a, nan=vmin, neginf=vmin, posinf=vmax) - vmin) / max(vmax - vmin, 1e-30)
    rgb = viridis(t)[::-1]                                     # origin='lower'
    return Image.fromarray(rgb, mode="RGB").resize(tuple(size), Image.BILINEAR)


def save_spectrogram_jpg(path: str, db, vmin: float, vmax: float, size=C_IMAGE_SIZE, quality: int = 95):
    """The JPG the reference saves at prime_detection.py:105 (input of ``detect_and_cluster_bursts``)."""
    spectrogram_image(db, vmin, vmax, size).save(path, format="JPEG", quality=quality)
    return path


def event_figure(path: str, crop: dict, title: str | None = None, size=(1400, 500)):
    """Per-event ``spec_and_psd`` picture of detector A (dsp/src/main.py:40-124): ``crop`` is one item of
    ``pipeline.event_crops`` (spectrogram dB ``sxx_db`` [n_freq, n_time] with ``t``/``f``, Welch PSD ``pxx_db`` over
    ``f_psd``, the event limits ``t_min``/``t_max`` inside the crop).  PNG, spectrogram 70 % / PSD 30 %."""
    from PIL import Image, ImageDraw
    w, h = size
    top = 40 if title else 10
    img = Image.new("RGB", (w, h), "white")
    draw = ImageDraw.Draw(img)
    if title:
        draw.text((10, 5), title.replace("\n", "   "), fill="black")
    sx0, sx1, sy0, sy1 = 60, int(w * 0.66), top, h - 40
    sxx = _to_numpy(crop["sxx_db"])
    lo, hi = float(np.min(sxx)), float(np.max(sxx))                       # pcolormesh autoscales to the data range
    img.paste(spectrogram_image(sxx, lo, hi, (sx1 - sx0, sy1 - sy0)), (sx0, sy0))
    t = np.asarray(crop["t"], dtype=np.float64)
    t_lo, t_hi = (float(t[0]), float(t[-1])) if len(t) > 1 else (0.0, 1.0)
    for tm in (crop.get("t_min"), crop.get("t_max")):                     # Marker(color='red', t_min=.., t_max=..)
        if tm is not None and t_hi > t_lo:
            x = sx0 + (float(tm) - t_lo) / (t_hi - t_lo) * (sx1 - sx0)
            if sx0 <= x <= sx1:
                for y in range(sy0, sy1, 8):
                    draw.line([(x, y), (x, min(y + 4, sy1))], fill="red", width=1)
    draw.rectangle([sx0, sy0, sx1, sy1], outline="black")
    draw.text((sx0, sy1 + 5), f"{t_lo:.2f} s", fill="black")
    draw.text((sx1 - 50, sy1 + 5), f"{t_hi:.2f} s", fill="black")
    f = np.asarray(crop["f"], dtype=np.float64)
    draw.text((5, sy1 - 10), f"{f[0]:.0f} Hz", fill="black")
    draw.text((5, sy0), f"{f[-1]:.0f} Hz", fill="black")
    px0, px1 = int(w * 0.72), w - 20
    pxx = _to_numpy(crop["pxx_db"])
    fp = np.asarray(crop["f_psd"], dtype=np.float64)
    draw.rectangle([px0, sy0, px1, sy1], outline="black")
    if len(pxx) > 1 and np.isfinite(pxx).any():
        plo, phi = float(np.nanmin(pxx)), float(np.nanmax(pxx))
        span = max(phi - plo, 1e-9)
        xs = px0 + (fp - fp[0]) / max(fp[-1] - fp[0], 1e-9) * (px1 - px0)
        ys = sy1 - (pxx - plo) / span * (sy1 - sy0)
        draw.line(list(zip(xs.tolist(), ys.tolist())), fill=(31, 119, 180), width=1)
        draw.text((px0, sy1 + 5), f"{fp[0]:.0f} Hz", fill="black")
        draw.text((px1 - 50, sy1 + 5), f"{fp[-1]:.0f} Hz", fill="black")
        draw.text((px0 + 4, sy0 + 2), f"{phi:.1f} dB", fill="black")
        draw.text((px0 + 4, sy1 - 12), f"{plo:.1f} dB", fill="black")
    img.save(path, format="PNG")
    return path
