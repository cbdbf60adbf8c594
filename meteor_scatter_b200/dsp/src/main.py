"""Drop-in for the reference's ``dsp/src/main.py``: same ``proc_wav_file``
signature, assertions, stdout lines, Audacity label file and per-event CSV
(dsp/src/main.py:207-229, 230-268, 352-363, 626-658) with the numeric path
(STFT, band power, thresholds, events) on the B200.

The per-detection ``spec_and_psd`` figures (dsp/src/main.py:721-806) are cut out
on the GPU (``pipeline.event_crops``) and written as PNG files into the export
directory (``render.event_figure``, numpy + Pillow, no matplotlib).  Out of scope
(SURVEY.md section 8): the interactive matplotlib/plotly debug figures; the
``debug_plot_*`` flags are accepted for signature compatibility and skipped with
a notice.
"""
from __future__ import annotations

import csv
import datetime
import os
from collections import Counter

import numpy as np
import torch

from ...pipeline import DetectorA, DetectorAParams, OutputDetection, event_crops  # noqa: F401  (re-exported type)
from ...wavio import read_wav


def proc_wav_file(file_path,
                  block_duration_sec,
                  freq_band,
                  noise_band,
                  n_fft,
                  threshold_std_factor,
                  wav_start_sec=None,
                  wav_end_sec=None,
                  debug_plot_whole=False,
                  debug_plot_config=False,
                  debug_plot_output=False,
                  debug_plot_output_interactive=False,
                  outfile_path=None,
                  out_audacity_lbl_file=None,
                  out_csv_file=None,
                  wav_start_date_time=None,
                  disable_show_and_write=False,
                  flag_adaptive_threshold=True,
                  threshold_estimation_window_sec=120,
                  threshold_freeze_before_detection_sec=3,
                  threshold_freeze_after_detection_sec=20,
                  threshold_fixed_init_duration_sec=10,
                  *, device=None, impl="auto", quiet=False):
    """See module docstring.  Extra keyword-only arguments: ``device`` (CUDA
    device, default current), ``impl`` ("auto" | "fft" | "tc"), ``quiet``.
    Supported inputs: PCM16 WAVs (and float32 WAVs holding PCM16 / 32768 values) with any ``n_fft`` and band width on
    the tensor-core kernels, provided ``int(fs * block_duration_sec)`` is a multiple of 8 samples; everything else
    (other float data, int32 / uint8 PCM converted to float32 -- 24-bit mantissa, ~6e-8 relative per sample) runs on
    the FFT kernel, which needs ``2 * n_fft`` to be a power of two in 256...16384 and raises MsUnsupported otherwise.
    Returns a dict with the detections and device tensors (the reference
    returns None; its callers ignore the value)."""
    say = (lambda *a, **k: None) if quiet else print
    assert os.path.exists(file_path), f"File does not exist: {file_path}"

    if outfile_path is not None:
        assert os.path.exists(
            os.path.dirname(outfile_path)), f"Output directory does not exist: {os.path.dirname(outfile_path)}"
        now = datetime.datetime.now()                                    # main.py:235-237: per-run export directory
        outfile_path = f"{outfile_path}/{now.strftime('%Y%m%d_%H%M%S')}/"
        os.makedirs(outfile_path, exist_ok=False)
    if out_audacity_lbl_file is not None:
        assert os.path.exists(os.path.dirname(out_audacity_lbl_file)), \
            f"Output directory does not exist: {os.path.dirname(out_audacity_lbl_file)}"
    if out_csv_file is not None:
        assert os.path.exists(
            os.path.dirname(out_csv_file)), f"Output directory does not exist: {os.path.dirname(out_csv_file)}"

    wav_sample_rate, wav_data = read_wav(file_path)                       # main.py:249

    if wav_start_sec is not None or wav_end_sec is not None:              # main.py:251-265
        if wav_start_sec is None:
            wav_start_sec = 0
        if wav_end_sec is None:
            wav_end_sec = len(wav_data) / wav_sample_rate
        start_sample = int(wav_start_sec * wav_sample_rate)
        end_sample = int(wav_end_sec * wav_sample_rate)
        assert start_sample < end_sample, "Start sample must be less than end sample"
        assert end_sample <= len(wav_data), "End sample exceeds length of audio data"
        wav_data = wav_data[start_sample:end_sample]

    assert wav_sample_rate == 6000, f"Sample rate must be 6000 Hz, but got {wav_sample_rate} Hz"
    assert len(wav_data.shape) == 1, f"Data must be mono or stereo, but got shape {wav_data.shape}"

    say("Wav duration [sec]:", len(wav_data) / wav_sample_rate)
    if debug_plot_whole or debug_plot_config or debug_plot_output or debug_plot_output_interactive:
        say("[ms_b200] debug plots are out of scope of the GPU path and were skipped")

    params = DetectorAParams(
        block_duration_sec=block_duration_sec, freq_band=tuple(freq_band), noise_band=tuple(noise_band), n_fft=n_fft,
        threshold_std_factor=threshold_std_factor, flag_adaptive_threshold=flag_adaptive_threshold,
        threshold_estimation_window_sec=threshold_estimation_window_sec,
        threshold_freeze_before_detection_sec=threshold_freeze_before_detection_sec,
        threshold_freeze_after_detection_sec=threshold_freeze_after_detection_sec,
        threshold_fixed_init_duration_sec=threshold_fixed_init_duration_sec, fs=wav_sample_rate)
    detector = DetectorA(params, impl=impl)      # event slots sized from the recording: no limit, like the reference
    spec = detector.spec
    num_blocks = spec.n_blocks(len(wav_data))

    say("n_fft [real]:", n_fft)
    say("Set n_fft to:", spec.n_fft_real, "samples")
    say("Wav block size in samples:", spec.block_size)
    say("Number of wav blocks:", num_blocks)
    freqs = np.fft.rfftfreq(spec.n_fft_real, d=1 / wav_sample_rate)
    say("Num of freq bins:", len(freqs))
    say("Bandwidth per freq bin [Hz]:", freqs[1] - freqs[0])
    say("Min Frequency [Hz]:", freqs[0])
    say("Max Frequency [Hz]:", freqs[-1])
    say("Power Band bandwidth [Hz]:", freq_band[1] - freq_band[0])
    say("Noise Band bandwidth [Hz]:", noise_band[1] - noise_band[0])

    if not torch.cuda.is_available():
        raise RuntimeError("proc_wav_file needs a CUDA device: the B200 detection path has no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())

    # host -> device: the only copy of the samples (int16 stays int16, float32 stays float32)
    used = np.ascontiguousarray(wav_data[:num_blocks * spec.block_size])
    if used.dtype == np.int16 or used.dtype == np.float32:
        host = torch.from_numpy(used.copy() if not used.flags.writeable else used)
    else:  # int32 / uint8 PCM: the reference would multiply by a float window anyway
        host = torch.from_numpy(used.astype(np.float32))
    x = host.to(dev, non_blocking=False).reshape(1, -1)

    if num_blocks == 0:
        # np.mean of an empty delta: the reference either finds nothing (adaptive) or raises IndexError (global)
        if not flag_adaptive_threshold:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")
        t_out_det, res = [], None
    else:
        res = detector.run(x, want_thresholds=True)
        t_out_det = res.detections(0, wav_start_date_time)
        if not flag_adaptive_threshold:
            say("Threshold for delta power detection [dB]:", float(res.det.thresholds[0, 0].item()))

    for det in t_out_det:                                                 # main.py:626-628
        say(f"Detection from {det.t_start:.2f} to {det.t_stop:.2f} seconds, dB: {det.dB:.2f} dB, "
            f"duration: {det.dur_s:.2f} seconds UTC_START: {det.utc_start}, UTC_STOP: {det.utc_stop}")

    if out_audacity_lbl_file is not None:                                 # main.py:630-638
        content_file_audacity = ""
        for det in t_out_det:
            content_file_audacity += f"{det.t_start:.2f}\t{det.t_stop:.2f}\tM\n"
        with open(out_audacity_lbl_file, 'w') as f:
            f.write(content_file_audacity)
        say("Write Pre-Lbl File to:", out_audacity_lbl_file)
        say("Wrote Items", len(t_out_det), "to Audacity LBL file")

    if out_csv_file is not None:                                          # main.py:640-658
        with open(out_csv_file, 'w', newline='') as csvfile:
            fieldnames = ['t_start', 't_stop', 'dur_s', 'dB', 'utc_start', 'utc_stop']
            writer = csv.DictWriter(csvfile, fieldnames=fieldnames)
            writer.writeheader()
            for det in t_out_det:
                writer.writerow({
                    't_start': det.t_start,
                    't_stop': det.t_stop,
                    'dur_s': det.dur_s,
                    'dB': det.dB,
                    'utc_start': det.utc_start.isoformat() if det.utc_start else None,
                    'utc_stop': det.utc_stop.isoformat() if det.utc_stop else None
                })
        say("Wrote Items", len(t_out_det), "to CSV file:", out_csv_file)

    crops = None
    if not disable_show_and_write and len(t_out_det):                    # main.py:721-806
        # per-detection spectrogram + PSD crops from the GPU; written as PNG when an export directory was given
        # (the reference shows them interactively otherwise, which a headless service cannot)
        crops = event_crops(x, wav_sample_rate, t_out_det, freq_band)
        if outfile_path is not None:
            from ... import render
            for det, crop in zip(t_out_det, crops):
                if crop["sxx_db"] is None or crop["pxx_db"] is None:
                    say(f"Error processing detection: cut of {crop['duration']:.2f} s is shorter than one frame")
                    continue
                title = (f"Detection from {det.t_start:.2f}s to {det.t_stop:.2f}s\n"
                         f"Wav duration: {crop['duration']:.2f}s, n_fft: {crop['n_fft']}\n"
                         f"Marker duration: {det.dur_s:.2f}s / dB: {det.dB:.2f}")
                render.event_figure(f"{outfile_path}spec_and_psd_{det.t_start:.2f}_{det.t_stop:.2f}.png", crop, title)

    count_per_hour = None
    if wav_start_date_time is not None:                                   # main.py:690-696
        count_per_hour = Counter(det.utc_start.replace(minute=0, second=0, microsecond=0) for det in t_out_det)

    return dict(detections=t_out_det, result=res, count_per_hour=count_per_hour, num_blocks=num_blocks, crops=crops)
