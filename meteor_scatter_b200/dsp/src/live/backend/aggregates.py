"""Interface types of the live detector.

Drop-in requirement: callers of the reference construct these records by
keyword (dsp/src/live/main.py:22-67), so class names, field names, field order
and defaults must equal the reference's dsp/src/live/backend/aggregates.py
(states :9-24, configs :32-63, DetectedMeteor :66-74).  They are declared here
from one table -- (field, type, default, meaning) -- instead of hand-written
class bodies; ``tests/test_host_logic.py`` pins names, order and defaults.

On the GPU path the *state* records are not used for computation (the state
lives in the POD ``ms_live_state`` of include/ms_b200.h); they are provided so
code that imports or type-checks against them keeps working.
"""
from __future__ import annotations

from dataclasses import MISSING, field, make_dataclass

_REQUIRED = MISSING


def _record(name, rows, base=None, doc=""):
    fields = []
    for fname, ftype, default, _meaning in rows:
        fields.append((fname, ftype) if default is _REQUIRED else (fname, ftype, field(default=default)))
    cls = make_dataclass(name, fields, bases=(base,) if base else ())
    cls.__doc__ = doc + "\n\n" + "\n".join(f"{r[0]}: {r[3]}" for r in rows)
    cls.__module__ = __name__
    return cls


State = make_dataclass("State", [])
State.__module__ = __name__
Config = make_dataclass("Config", [])
Config.__module__ = __name__

StateInitialization = _record("StateInitialization", [
    ("history_channel_dB", list, _REQUIRED, "mean PSD level of every block seen while initialising"),
], State, "Waiting for init_detection_wait_sec of audio before thresholds are trusted.")

StateDetection = _record("StateDetection", [
    ("locked_threshold", float, -1.0, "threshold kept from the last tracked event"),
    ("use_locked_threshold_until_secs", float, -1.0, "the locked threshold applies to blocks ending before this time"),
], State, "Searching: a block above the threshold starts tracking.")

StateTracking = _record("StateTracking", [
    ("locked_threshold", float, _REQUIRED, "threshold frozen when the event started"),
    ("time_start_detection", float, _REQUIRED, "start time of the block that triggered"),
    ("history_over_noise_sig_dB", list, _REQUIRED, "signal-over-noise level of every tracked block"),
], State, "Inside an event: ends with the first block below the locked threshold.")

ConfigDetection = _record("ConfigDetection", [
    ("proc_block_sec", float, 0.2, "block length in seconds"),
    ("n_fft", int, 4096, "zero-padded transform length of the per-block Welch PSD"),
    ("signal_freq", int, 1000, "expected beacon tone in the audio, Hz"),
    ("channel_width", int, 100, "width of the signal channel and of each noise channel, Hz"),
    ("noise_channel_offset", int, 300, "noise channels sit this far below and above signal_freq, Hz"),
    ("avg_win_sec", float, 8, "length of the history used for mean and std of the level, s"),
    ("init_detection_wait_sec", float, 8 * 1.0, "initialisation period, s"),
    ("after_tracking_wait_sec", float, 8 * 1.5, "how long the locked threshold stays in force after an event, s"),
    ("threshold_std_factor", float, 4, "k in threshold = mean + k * std"),
    ("detection_db_over_noise_mean_min", float, -1, "discard events whose mean level is below this, dB"),
    ("detection_dur_min_sec", float, -1, "discard events shorter than this, s"),
], Config, "Numeric parameters of the live detector.")

ConfigVisualization = _record("ConfigVisualization", [
    ("enable_ui_plots", bool, True, "matplotlib UI of the reference; accepted, nothing is rendered here"),
    ("realtime_factor", float, 16, "UI replay speed (unused)"),
    ("flag_realtime_animation", bool, True, "UI pacing (unused)"),
    ("max_range_sec", int, 60, "length of the waterfall ring, s"),
    ("limit_freq_offset_wf2_and_export", int, 100, "half width of the exported band around signal_freq, Hz"),
    ("wf_offset_vmin", int, 20, "colour range below the init level, dB (rendering only)"),
    ("wf_offset_vmax", int, 20, "colour range above the init level, dB (rendering only)"),
    ("enable_debug_logs", bool, False, "verbose per-block prints of the reference (unused)"),
], Config, "Display parameters; only max_range_sec and the export band matter off-screen.")

ConfigSpecExport = _record("ConfigSpecExport", [
    ("output_dir", str, "", "JPG export directory of the reference; '' disables it"),
    ("time_before_meteor_sec", int, 3, "margin before the event in an exported crop, s"),
    ("time_after_meteor_sec", int, 3, "margin after the event in an exported crop, s"),
], Config, "Spectrogram crop export around each detection.")

DetectedMeteor = _record("DetectedMeteor", [
    ("time_start", float, _REQUIRED, "start of the triggering block, s"),
    ("time_stop", float, _REQUIRED, "start of the first block back below the threshold, s"),
    ("duration", float, _REQUIRED, "time_stop - time_start"),
    ("db_min", float, _REQUIRED, "minimum level over the tracked blocks, dB over noise"),
    ("db_max", float, _REQUIRED, "maximum level"),
    ("db_mean", float, _REQUIRED, "mean level"),
    ("db_std", float, _REQUIRED, "population standard deviation of the level"),
], None, "One detected meteor echo.")
