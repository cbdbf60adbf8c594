"""CPU: the C-ABI library builds, loads, and exports every symbol that
include/ms_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from meteor_scatter_b200 import build
    return build.build()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "ms_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(ms_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for n in ("ms_band_power_i16", "ms_band_power_f32", "ms_band_power_i16_tc", "ms_detect_global",
              "ms_detect_adaptive", "ms_hourly_counts", "ms_welch_band_db_f32", "ms_live_state_step",
              "ms_psd_spectrogram_i16"):
        assert n in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for n in declared_functions():
        assert hasattr(lib, n), f"{n} declared in ms_b200.h but not exported by libms_b200.so"
    assert lib.ms_abi_version() == 1


def test_binding_table_matches_header(lib_path):
    from meteor_scatter_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()
    _lib.load()


def test_struct_layouts_match_header():
    from meteor_scatter_b200 import _lib
    # ms_live_state: 8 + 4*4 + 8*8 + 256*8 bytes; ms_live_config: 8 + 6*8 + 2*4
    assert ctypes.sizeof(_lib.LiveState) == 8 + 16 + 64 + 2048
    assert ctypes.sizeof(_lib.LiveConfig) == 64


def test_workspace_and_plan_size_queries_run_without_gpu(lib_path):
    from meteor_scatter_b200 import _lib
    lib = _lib.load()
    assert lib.ms_detect_workspace_bytes(288, 1500) > 288 * 1500 * 24
    assert lib.ms_dft_i8_plan_bytes(1024, 14) == 1024 + 16 * 8192
    assert lib.ms_dft_i8_plan_bytes(1024, 17) == 0


def test_no_cpu_fallback_in_product_path():
    """ops refuse CPU tensors instead of silently computing elsewhere."""
    import torch
    from meteor_scatter_b200 import ops
    spec = ops.BandSpec.from_reference_args(6000, 0.2, (993, 1013), (690, 710), 512)
    assert spec.sig_bins == (170, 171, 172) and spec.noise_bins == (118, 119, 120, 121)
    assert spec.block_size == 1200 and spec.win_len == 1024 and spec.n_fft_real == 1024
    with pytest.raises(ValueError):
        ops.band_power(torch.zeros(1, 2400, dtype=torch.int16), spec)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "meteor_scatter_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{fn} imports the oracle"


def test_integration_guide_names_every_entry_point():
    """INTEGRATION.md maps each declared C-ABI function to the reference code it replaces; a new entry point must be
    added there too."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "ms_b200.h")).read()
    guide = open(os.path.join(root, "INTEGRATION.md")).read()
    names = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(ms_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert len(names) >= 25
    missing = sorted(n for n in names if n not in guide)
    assert not missing, missing


def test_argument_validation_needs_no_gpu(lib_path):
    """Every entry point validates its arguments before touching CUDA: null pointers and impossible geometry are
    reported as MS_ERR_INVALID_ARG / MS_ERR_UNSUPPORTED with a message, on a machine without a GPU too."""
    import ctypes as C
    from meteor_scatter_b200 import _lib
    lib = _lib.load()
    lib.ms_last_error.restype = C.c_char_p
    rc = lib.ms_detector_a_pass_overlapped_i16(None, 1, 1, 1200, None, 1024, 14, 4.0, 600, 15, 100, 50, 16, None, None,
                                               None, None, None, None, 0, None, 0.2, 0.5, 0, 24, None, None, None,
                                               None, None, None, None)
    assert rc == -1 and b"null pointer" in lib.ms_last_error()
    rc = lib.ms_welch_band_db_i8_i16(None, 1, 800, 1, 800, 256, None, None, None, None)
    assert rc == -1
    cfg = _lib.LiveConfig(block_samples=800, fs=4000.0, k_std=4.0, init_wait_sec=8.0, after_wait_sec=12.0,
                          mean_min_db=1.0, dur_min_sec=0.5, avg_win=0)
    dummy = (C.c_char * 64)()
    rc = lib.ms_live_state_step_ws(dummy, C.byref(cfg), 1, dummy, 8, 1, 8, 4, dummy, dummy, None, None, 0, None)
    assert rc == -2 and b"avg_win" in lib.ms_last_error()          # the reference's avg_win == 0 quirk is refused
    assert lib.ms_live_state_workspace_bytes(256, 3000) == 256 * 3000 * 32 + 256 * 94 * 4
    assert lib.ms_live_state_workspace_bytes(256, 5) == 0
    assert lib.ms_welch_i8_plan_bytes(256) == 2048 + 4 * 240 * 128 and lib.ms_welch_i8_plan_bytes(100) == 0
