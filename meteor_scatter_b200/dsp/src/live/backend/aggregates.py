"""Interface types of the live detector, field-for-field compatible with the
reference's dsp/src/live/backend/aggregates.py (states :9-24, configs :32-63,
DetectedMeteor :66-74) so existing callers can construct and pass them as is."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class State:
    pass


@dataclass
class StateInitialization(State):
    history_channel_dB: list


@dataclass
class StateDetection(State):
    locked_threshold: float = -1.0
    use_locked_threshold_until_secs: float = -1.0


@dataclass
class StateTracking(State):
    locked_threshold: float
    time_start_detection: float
    history_over_noise_sig_dB: list


@dataclass
class Config:
    pass


@dataclass
class ConfigDetection(Config):
    proc_block_sec: float = 0.2             # block length [s]
    n_fft: int = 4096                       # zero-padded transform length of the per-block PSD
    signal_freq: int = 1000                 # expected beacon tone [Hz]
    channel_width: int = 100                # width of the signal / noise channels [Hz]
    noise_channel_offset: int = 300         # noise channels sit at signal_freq -/+ this offset [Hz]
    avg_win_sec: float = 8                  # history length for mean / std [s]
    init_detection_wait_sec: float = 8 * 1.0
    after_tracking_wait_sec: float = 8 * 1.5
    threshold_std_factor: float = 4
    detection_db_over_noise_mean_min: float = -1
    detection_dur_min_sec: float = -1


@dataclass
class ConfigVisualization(Config):
    enable_ui_plots: bool = True            # accepted; the GPU path renders nothing
    realtime_factor: float = 16
    flag_realtime_animation: bool = True
    max_range_sec: int = 60
    limit_freq_offset_wf2_and_export: int = 100
    wf_offset_vmin: int = 20
    wf_offset_vmax: int = 20
    enable_debug_logs: bool = False


@dataclass
class ConfigSpecExport(Config):
    output_dir: str = ""                    # "" disables the (unsupported) JPG export
    time_before_meteor_sec: int = 3
    time_after_meteor_sec: int = 3


@dataclass
class DetectedMeteor:
    time_start: float
    time_stop: float
    duration: float
    db_min: float
    db_max: float
    db_mean: float
    db_std: float
