"""ORACLE tooling (test infrastructure, NOT product code).

Runs the *unmodified* reference (/root/reference, read-only) in the build
container to pin the oracle: GUI-only imports (matplotlib, plotly) are stubbed
with MagicMock, ``soundfile`` is shimmed over scipy.io.wavfile (a stand-in for
a third-party reader, not for reference arithmetic), and the locals of the
reference functions are captured at return with ``sys.setprofile`` because
both detectors return None.

/root/reference does not exist on the GPU box.  ``oracle/stage_ref.py`` (run by
``__graft_entry__.build()`` in the build container) stages byte-identical copies
of the hot-path reference files under the git-ignored ``oracle/_ref/``, which
travel to the GPU box; this module falls back to them, so ``bench.py --impl
reference`` / ``cpu_baseline`` time the reference's own code there.  Other users:
``tests/golden/make_golden.py`` (run here) and the CPU tests that pin the oracle.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types
from unittest import mock

import numpy as np

_STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("MS_REFERENCE_ROOT", "/root/reference")
if not os.path.exists(os.path.join(REFERENCE_ROOT, "dsp/src/main.py")):
    REFERENCE_ROOT = _STAGED_ROOT        # the copy staged by oracle/stage_ref.py (GPU box)


def reference_available() -> bool:
    return os.path.exists(os.path.join(REFERENCE_ROOT, "dsp/src/main.py"))


def _stub_gui_modules():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.ticker",
                 "plotly", "plotly.graph_objects"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)


def _soundfile_shim():
    """soundfile.read(path, start=, stop=) -> (float64 in [-1,1), fs) for PCM16."""
    import scipy.io.wavfile as wavfile

    m = types.ModuleType("soundfile")

    def read(path, start=0, stop=None, **kw):
        fs, data = wavfile.read(path)
        if data.dtype == np.int16:
            data = data.astype(np.float64) / 32768.0
        elif data.dtype == np.int32:
            data = data.astype(np.float64) / 2147483648.0
        else:
            data = data.astype(np.float64)
        start = int(start)
        stop = None if stop is None else int(stop)
        return data[start:stop], fs

    m.read = read
    return m


def load_reference_main():
    """Import dsp/src/main.py as a module without running its __main__ block."""
    _stub_gui_modules()
    path = os.path.join(REFERENCE_ROOT, "dsp/src/main.py")
    spec = importlib.util.spec_from_file_location("ref_dsp_main", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.tqdm = lambda it, *a, **k: it      # progress bar only
    return mod


def load_reference_processor():
    _stub_gui_modules()
    sys.modules["soundfile"] = _soundfile_shim()
    live = os.path.join(REFERENCE_ROOT, "dsp/src/live")
    if live not in sys.path:
        sys.path.insert(0, live)
    for k in [k for k in sys.modules if k == "backend" or k.startswith("backend.")]:
        del sys.modules[k]
    from backend import processor, aggregates  # noqa
    processor.tqdm = lambda it, *a, **k: it
    return processor, aggregates


def _capture_locals(func, wanted, *args, **kwargs):
    """Call func and return {name: value} of its locals at return time."""
    captured = {}
    code = func.__code__

    def prof(frame, event, arg):
        if event == "return" and frame.f_code is code:
            for w in wanted:
                if w in frame.f_locals:
                    captured[w] = frame.f_locals[w]

    sys.setprofile(prof)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            func(*args, **kwargs)
    finally:
        sys.setprofile(None)
    return captured


def run_reference_a(wav_path, **kwargs):
    """proc_wav_file unmodified -> dict(band_power, noise_power, delta_power,
    t_threshold, t_out_det)."""
    mod = load_reference_main()
    kwargs.setdefault("disable_show_and_write", True)
    return _capture_locals(mod.proc_wav_file,
                           ("band_power", "noise_power", "delta_power", "t_threshold", "t_out_det"),
                           wav_path, **kwargs)


def run_reference_b(wav_path, cfg_kwargs):
    """wav_file_process unmodified -> captured per-block series and detections."""
    processor, agg = load_reference_processor()
    cap = _capture_locals(
        processor.wav_file_process,
        ("local_out_res_detections", "local_data_over_noise_sig", "local_data_over_noise_sig_threshold",
         "local_data_abs_meas_sig", "local_data_abs_meas_noise_1", "local_data_abs_meas_noise_2"),
        wav_path, agg.ConfigDetection(**cfg_kwargs), agg.ConfigVisualization(enable_ui_plots=False),
        agg.ConfigSpecExport(output_dir=""))
    return cap
