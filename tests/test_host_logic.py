"""CPU: host-side logic of the drop-in (no GPU, no compute calls)."""
import datetime
import inspect
import os

import numpy as np
import pytest

from meteor_scatter_b200 import csvout, wavio
from meteor_scatter_b200.pipeline import DetectorAParams, datetime_to_us, hour_index


def test_signatures_match_reference():
    from meteor_scatter_b200.dsp.src.main import proc_wav_file
    from meteor_scatter_b200.dsp.src.live.backend.processor import wav_file_process
    from meteor_scatter_b200.meteor_detect_class.detector_and_classification import detect_and_cluster_bursts
    from meteor_scatter_b200.meteor_detect_class.prime_detection import plot_spectrogram
    ref = ["file_path", "block_duration_sec", "freq_band", "noise_band", "n_fft", "threshold_std_factor",
           "wav_start_sec", "wav_end_sec", "debug_plot_whole", "debug_plot_config", "debug_plot_output",
           "debug_plot_output_interactive", "outfile_path", "out_audacity_lbl_file", "out_csv_file",
           "wav_start_date_time", "disable_show_and_write", "flag_adaptive_threshold",
           "threshold_estimation_window_sec", "threshold_freeze_before_detection_sec",
           "threshold_freeze_after_detection_sec", "threshold_fixed_init_duration_sec"]     # main.py:207-229
    sig = inspect.signature(proc_wav_file)
    pos = [n for n, p in sig.parameters.items() if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert pos == ref
    d = {n: p.default for n, p in sig.parameters.items()}
    assert (d["flag_adaptive_threshold"], d["threshold_estimation_window_sec"],
            d["threshold_freeze_before_detection_sec"], d["threshold_freeze_after_detection_sec"],
            d["threshold_fixed_init_duration_sec"], d["disable_show_and_write"]) == (True, 120, 3, 20, 10, False)
    pos = [n for n, p in inspect.signature(wav_file_process).parameters.items() if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert pos == ["wav_file_path", "config_detection", "config_visualization", "config_spec_export",
                   "wav_file_start_sec", "wav_file_stop_sec"]                                   # processor.py:14-21
    assert list(inspect.signature(detect_and_cluster_bursts).parameters) == \
        ["image_path", "eps", "min_samples", "display", "output_path"]                        # d_a_c.py:7
    pos = [n for n, p in inspect.signature(plot_spectrogram).parameters.items() if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert pos == ["iq_segment", "fs", "display", "vmin", "vmax"]                             # prime_detection.py:65


def test_aggregates_defaults_match_reference():
    from meteor_scatter_b200.dsp.src.live.backend import aggregates as ag
    c = ag.ConfigDetection()
    assert (c.proc_block_sec, c.n_fft, c.signal_freq, c.channel_width, c.noise_channel_offset, c.avg_win_sec,
            c.init_detection_wait_sec, c.after_tracking_wait_sec, c.threshold_std_factor,
            c.detection_db_over_noise_mean_min, c.detection_dur_min_sec) == (0.2, 4096, 1000, 100, 300, 8, 8.0, 12.0,
                                                                              4, -1, -1)
    v = ag.ConfigVisualization()
    assert (v.enable_ui_plots, v.realtime_factor, v.max_range_sec, v.enable_debug_logs) == (True, 16, 60, False)
    assert ag.ConfigSpecExport().output_dir == ""
    assert [f for f in ag.DetectedMeteor.__dataclass_fields__] == ["time_start", "time_stop", "duration", "db_min",
                                                                  "db_max", "db_mean", "db_std"]


def test_band_bins_of_live_parameter_sets():
    from meteor_scatter_b200.dsp.src.live.backend import aggregates as ag
    from meteor_scatter_b200.dsp.src.live.backend.processor import band_bins
    _, b = band_bins(ag.ConfigDetection(signal_freq=1020), 4000)      # SURVEY appendix A, B/live #1
    assert b == [(994, 1095), (687, 788), (1301, 1402)]
    _, b = band_bins(ag.ConfigDetection(signal_freq=1025), 4000)      # B/live #2
    assert b == [(999, 1100), (692, 793), (1306, 1408)]
    _, b = band_bins(ag.ConfigDetection(), 4000)                      # defaults
    assert b == [(973, 1075), (666, 768), (1280, 1382)]


def test_block_counts_truncate_like_the_reference():
    assert DetectorAParams().block_counts() == (600, 15, 100, 50)
    assert DetectorAParams(threshold_freeze_before_detection_sec=0.6).block_counts()[1] == 2   # int(0.6/0.2) == 2


def test_wav_roundtrip_and_name_parsing(tmp_path):
    x = (np.arange(5000) % 300 - 150).astype(np.int16)
    p = tmp_path / "expoFull_gqrx_20250625_075141_49969000.wav"
    wavio.write_wav_pcm16(str(p), 6000, x)
    fs, d = wavio.read_wav(str(p))
    assert fs == 6000 and d.dtype == np.int16 and np.array_equal(d, x)
    import scipy.io.wavfile as sw
    fs2, d2 = sw.read(str(p))
    assert fs2 == 6000 and np.array_equal(d2, x)
    xf = (x / 32768.0).astype(np.float32)
    pf = tmp_path / "f.wav"
    wavio.write_wav_pcm16(str(pf), 4000, xf)
    fs, d = wavio.read_wav(str(pf))
    assert fs == 4000 and d.dtype == np.float32 and np.array_equal(d, xf)
    assert wavio.start_time_from_name(str(p)) == datetime.datetime(2025, 6, 25, 7, 51, 41)       # main.py:859-862
    assert wavio.start_time_from_name("/x/expoFull_Brams_250607_23MESZ.wav") == \
        datetime.datetime(2025, 6, 7, 21, 0, 0)                                                 # main.py:917-923
    assert wavio.start_time_from_name("/x/whatever.wav") is None
    with pytest.raises(wavio.WavFormatError):
        bad = tmp_path / "bad.wav"
        bad.write_bytes(b"not a wav file at all")
        wavio.read_wav(str(bad))


def test_hourly_day_files_match_dashboard_format(tmp_path):
    hour0 = datetime.datetime(2024, 11, 29, 22, 0, 0)
    hist = np.array([[3, 1], [0, 0], [101, 0], [8, 3]])
    rows = csvout.hourly_rows(hist, hour0)
    files = csvout.write_day_files(str(tmp_path), rows)
    assert [os.path.basename(f) for f in files] == ["20241129.csv", "20241130.csv"]
    assert len(os.path.basename(files[0])) == 12                                               # database.py:247
    txt = open(files[1], newline="").read()
    assert txt == "Timestamp;Anzahl;Kritisch\n2024-11-30 00:00:00;101;0\n2024-11-30 01:00:00;8;3\n"
    # the consumer's reader (database.py:95) parses it
    import pandas as pd
    df = pd.read_csv(files[1], sep=";")
    assert list(df.columns) == ["Timestamp", "Anzahl", "Kritisch"] and int(df["Anzahl"].sum()) == 109
    assert pd.to_datetime(df["Timestamp"]).dt.hour.tolist() == [0, 1]
    # idempotent re-run, merge with an updated hour
    csvout.write_day_files(str(tmp_path), [(datetime.datetime(2024, 11, 30, 1), 9, 4)])
    assert open(files[1]).read().splitlines()[-1] == "2024-11-30 01:00:00;9;4"
    # byte-identical to the reference's pandas writer (prime_detection.py:138, 245)
    ref = tmp_path / "ref.csv"
    pd.DataFrame([{"Timestamp": "2024-11-30 00:00:00", "Anzahl": 101, "Kritisch": 0},
                  {"Timestamp": "2024-11-30 01:00:00", "Anzahl": 9, "Kritisch": 4}]).to_csv(ref, sep=";", index=False)
    assert open(ref, newline="").read() == open(files[1], newline="").read()


def test_hourly_csv_accumulator_follows_reference_loop(tmp_path):
    from meteor_scatter_b200.meteor_detect_class.prime_detection import HourlyCsv
    t0 = datetime.datetime(2024, 11, 30, 22, 30, 0)
    acc = HourlyCsv(str(tmp_path), now=t0)
    assert open(tmp_path / "20241130.csv").read() == "Timestamp;Anzahl;Kritisch\n"
    assert acc.add(1, 2, now=t0 + datetime.timedelta(minutes=30)) is None
    row = acc.add(0, 1, now=t0 + datetime.timedelta(minutes=59, seconds=50))
    assert row == "2024-11-30 22:30:00;4;1"
    acc.add(5, 5, now=t0 + datetime.timedelta(minutes=100))       # date rolls over: new file, counters reset
    assert os.path.exists(tmp_path / "20241201.csv")
    assert acc.n_critical == 0 and acc.n_non_critical == 0


def test_time_helpers():
    t = datetime.datetime(2025, 6, 1, 13, 59, 59, 999999)
    assert datetime_to_us(t) % 1_000_000 == 999999
    assert hour_index(t) + 1 == hour_index(t + datetime.timedelta(microseconds=1))


def test_classification_rule():
    from meteor_scatter_b200.meteor_detect_class.detector_and_classification import classify_events
    crit, non = classify_events([0.2, 0.4, 0.6000000000000001, 0.5, 2.0])
    assert crit == [2, 3, 4] and non == [0, 1]


def test_sharding_and_hour_span():
    from meteor_scatter_b200.batch import hour_span, shard_indices
    assert shard_indices(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((shard_indices(10, r, 4) for r in range(4)), [])) == list(range(10))
    t0 = datetime.datetime(2025, 6, 1, 22, 35)
    h0, n = hour_span([t0, t0 + datetime.timedelta(minutes=90)], [300.0, 300.0])
    assert h0 == datetime.datetime(2025, 6, 1, 22) and n == 3


def test_stage_files_threaded_matches_sequential(tmp_path):
    """A-io: ragged PCM16 files staged by several reader threads == staged one by one; padding is zero and every row
    starts 16-byte aligned (what the tensor-map path needs)."""
    import numpy as np
    from meteor_scatter_b200.batch import stage_files
    from meteor_scatter_b200.wavio import write_wav_pcm16
    rng = np.random.default_rng(0)
    paths, xs = [], []
    for i, n in enumerate([6000 * 7 + 13, 6000 * 3, 1, 6000 * 7 + 14, 4801]):
        x = rng.integers(-32768, 32767, size=n, dtype=np.int16)
        p = tmp_path / f"f{i}.wav"
        write_wav_pcm16(str(p), 6000, x)
        paths.append(str(p))
        xs.append(x)
    h1, l1 = stage_files(paths, 6000, io_threads=1)
    h4, l4 = stage_files(paths, 6000, io_threads=4)
    assert np.array_equal(l1, [len(x) for x in xs]) and np.array_equal(l1, l4)
    assert h1.shape == h4.shape and h1.shape[1] % 8 == 0 and h1.shape[1] >= max(l1)
    assert np.array_equal(h1.numpy(), h4.numpy())
    for i, x in enumerate(xs):
        assert np.array_equal(h4.numpy()[i, :len(x)], x) and not h4.numpy()[i, len(x):].any()
    empty, lens = stage_files([], 6000)
    assert empty.numel() == 0 and len(lens) == 0


def test_native_file_reader_fills_pinned_rows(tmp_path):
    """ms_read_files (native pread threads) behind batch.stage_files: ragged lengths, an empty file, zero padding,
    a float32 file converted through the memory-map path, and a missing file raising instead of returning garbage."""
    import numpy as np
    import pytest
    from meteor_scatter_b200 import batch
    from meteor_scatter_b200._lib import MsError
    from meteor_scatter_b200.wavio import write_wav_pcm16
    rng = np.random.default_rng(0)
    arrs = [rng.integers(-3000, 3000, size=n).astype(np.int16) for n in (1000, 1003, 7, 0, 5000)]
    paths = []
    for i, a in enumerate(arrs):
        p = str(tmp_path / f"f{i}.wav")
        write_wav_pcm16(p, 6000, a)
        paths.append(p)
    for threads in (1, 3, 16):
        host, lens = batch.stage_files(paths, io_threads=threads)
        assert host.shape == (5, 5000) and list(lens) == [len(a) for a in arrs]
        for i, a in enumerate(arrs):
            assert np.array_equal(host[i, :len(a)].numpy(), a) and not host[i, len(a):].any()
    f32 = str(tmp_path / "g.wav")
    write_wav_pcm16(f32, 6000, (arrs[0].astype(np.float32) / 32768.0))
    host, lens = batch.stage_files([paths[1], f32])
    assert host.dtype.is_floating_point and np.array_equal(host[1, :1000].numpy(), arrs[0].astype(np.float32) / 32768.0)
    assert np.array_equal(host[0, :1003].numpy(), arrs[1].astype(np.float32))
    from meteor_scatter_b200.wavio import wav_info
    infos = [wav_info(p) for p in paths[:2]]
    os.remove(paths[1])
    with pytest.raises(MsError, match="cannot open"):
        batch._fill_rows(paths[:2], infos, np.array([1000, 1003]), np.dtype(np.int16),
                         np.zeros((2, 1008), dtype=np.int16), 2)


def test_uncovered_hours_are_gaps_not_zero_rows():
    """process_files writes no row for an hour without audio (the reference writes rows only while it runs)."""
    from meteor_scatter_b200 import batch
    t0 = datetime.datetime(2025, 6, 25, 22, 50, 0)
    starts = [t0, t0 + datetime.timedelta(minutes=5), t0 + datetime.timedelta(hours=3, minutes=58)]
    durs = [300.0, 300.0, 300.0]                      # 22:50-23:00, ..., 02:48-02:53 next day
    hour0, n_hours = batch.hour_span(starts, durs)
    assert hour0 == datetime.datetime(2025, 6, 25, 22) and n_hours == 5
    assert batch.uncovered_hours(starts, durs, hour0, n_hours) == [1, 2, 3]
    # a recording that ends exactly on the hour does not cover the next hour
    assert batch.uncovered_hours([t0 + datetime.timedelta(minutes=5)], [300.0], hour0, 2) == [1]
    rows = csvout.hourly_rows(np.zeros((5, 2), dtype=int), hour0, skip_empty_hours=[1, 2, 3])
    assert [r[0].hour for r in rows] == [22, 2]


def test_cosine_series_recognises_the_periodic_scipy_windows():
    """ops.cosine_series decides whether the frequency-domain-window form applies: scipy's periodic 'boxcar', 'hann',
    'hamming', 'blackman' are cosine series of the frame length (orders 0, 1, 1, 2); the symmetric np.hanning and an
    arbitrary window are not."""
    import scipy.signal as ss
    from meteor_scatter_b200 import ops
    for name, order, coef in (("boxcar", 0, [1.0, 0.0, 0.0]), ("hann", 1, [0.5, -0.25, 0.0]),
                              ("hamming", 1, [0.54, -0.23, 0.0]), ("blackman", 2, [0.42, -0.25, 0.04])):
        got = ops.cosine_series(ss.get_window(name, 1024))
        assert got is not None and got[0] == order
        np.testing.assert_allclose(got[1], coef, atol=1e-12)
        # the series reproduces the window
        n = np.arange(1024)
        w = got[1][0] + sum(2 * got[1][m] * np.cos(2 * np.pi * m * n / 1024) for m in (1, 2))
        np.testing.assert_allclose(w, ss.get_window(name, 1024), atol=1e-12)
    assert ops.cosine_series(np.hanning(1024)) is None
    assert ops.cosine_series(np.random.default_rng(0).uniform(0.1, 1.0, 512)) is None
    assert ops.cosine_series(np.zeros(64)) is None
