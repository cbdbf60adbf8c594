"""Drop-in for the reference's causal "live" detector,
dsp/src/live/backend/processor.py:14-543: same ``wav_file_process`` signature,
assertions and "Detected Meteor: ..." output, with the per-block Welch band
power (processor.py:206, 349-367, 393) and the threshold/state machine
(:393-414, 444-510) on the GPU.

Out of scope: the matplotlib UI, waterfall and JPG export.
``enable_ui_plots`` / ``ConfigSpecExport.output_dir`` are accepted and skipped
with a notice.  ``LiveDetector`` below is the streaming form (1 s chunks keep a
device-resident state between calls).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from meteor_scatter_b200 import ops
from meteor_scatter_b200._lib import LiveConfig
from meteor_scatter_b200.wavio import read_wav
from .aggregates import ConfigDetection, ConfigSpecExport, ConfigVisualization, DetectedMeteor


def band_bins(cfg: ConfigDetection, fs: float):
    """Inclusive rfft bin ranges of the signal channel and the two noise channels
    (float edge arithmetic of processor.py:32-44, masks of :349-362)."""
    hw = cfg.channel_width / 2
    edges = ((cfg.signal_freq - hw, cfg.signal_freq + hw),
             ((cfg.signal_freq - cfg.noise_channel_offset) - hw, (cfg.signal_freq - cfg.noise_channel_offset) + hw),
             ((cfg.signal_freq + cfg.noise_channel_offset) - hw, (cfg.signal_freq + cfg.noise_channel_offset) + hw))
    freqs = np.fft.rfftfreq(cfg.n_fft, 1 / fs)
    out = []
    for lo, hi in edges:
        k = np.nonzero((freqs >= lo) & (freqs <= hi))[0]
        out.append((int(k[0]), int(k[-1])) if len(k) else (1, 0))
    return edges, out


def live_config(cfg: ConfigDetection, fs: float, block_size: int) -> LiveConfig:
    """Kernel-side view of ``ConfigDetection``.  The threshold history lives in a 256-deep device ring
    (``ms_live_state.hist``), so ``int(avg_win_sec / proc_block_sec)`` must be in [1, 256] (51 s at the reference's
    0.2 s blocks; the reference's own configurations use 40).  The reference also accepts 0 (``[-0:]`` = the whole
    history so far) and longer windows; those raise here instead of running with a different estimator."""
    avg_win = int(cfg.avg_win_sec / cfg.proc_block_sec)                           # processor.py:56
    if not 1 <= avg_win <= 256:
        from meteor_scatter_b200._lib import MsUnsupported
        raise MsUnsupported(-2, f"avg_win_sec / proc_block_sec = {avg_win} blocks is outside the supported 1..256 "
                                "(device history ring of the live state machine)")
    return LiveConfig(block_samples=block_size, fs=float(fs), k_std=float(cfg.threshold_std_factor),
                      init_wait_sec=float(cfg.init_detection_wait_sec),
                      after_wait_sec=float(cfg.after_tracking_wait_sec),
                      mean_min_db=float(cfg.detection_db_over_noise_mean_min),
                      dur_min_sec=float(cfg.detection_dur_min_sec),
                      avg_win=int(cfg.avg_win_sec / cfg.proc_block_sec))        # processor.py:56


class LiveDetector:
    """Streaming detector B for ``n_streams`` independent audio streams.

    ``push(chunk)`` consumes whole blocks (``[n_streams, k*block]`` samples,
    PCM16 or float32 in [-1, 1)) and returns the detections that completed in
    this chunk; the threshold history and state machine stay in HBM."""

    def __init__(self, cfg: ConfigDetection, fs: int = 4000, n_streams: int = 1, device="cuda", max_det: int = 4096,
                 waterfall: ConfigVisualization | None = None, export: ConfigSpecExport | None = None):
        """``waterfall``: keep the reference's waterfall ring (processor.py:223-229) on the device -- the last
        ``max_range_sec`` of per-block PSD rows (dB) for the bins within ``limit_freq_offset_wf2_and_export`` of
        the signal frequency -- so spectrogram crops around each detection (processor.py:295-343) can be cut
        out as tensors with ``export_ready()``; ``export`` gives the margins before/after the meteor."""
        self.cfg, self.fs = cfg, fs
        self.block = int(cfg.proc_block_sec * fs)                                # processor.py:75
        self.edges, self.bands = band_bins(cfg, fs)
        self.lc = live_config(cfg, fs, self.block)
        self.states = ops.LiveStates(n_streams, device, max_det=max_det)
        self._seen = np.zeros(n_streams, dtype=np.int64)
        self.n_blocks = 0                                                        # blocks consumed per stream
        self.viz, self.export = waterfall, export or ConfigSpecExport()
        self.rows = None
        if waterfall is not None:
            freqs = np.fft.rfftfreq(cfg.n_fft, 1 / fs)
            k = np.nonzero((freqs >= cfg.signal_freq - waterfall.limit_freq_offset_wf2_and_export) &
                           (freqs <= cfg.signal_freq + waterfall.limit_freq_offset_wf2_and_export))[0]
            self.rows = (int(k[0]), int(k[-1]))
            self.row_freqs = freqs[k]
            self.ring_len = int(waterfall.max_range_sec / cfg.proc_block_sec)    # processor.py:55
            self.ring = torch.full((n_streams, self.ring_len, len(k)), float("-inf"), dtype=torch.float32,
                                   device=device)
            self._ring_base = torch.zeros((), dtype=torch.int64, device=device)  # blocks written so far (device side)
            self._ring_idx = {}                                                  # per chunk length: arange on device
            self._pending = []                                                   # detections not exported yet

    def _ring_update(self, rows: torch.Tensor):
        """Write this chunk's PSD rows into the waterfall ring.  The write position lives on the device, so the
        same ops are valid eagerly and inside the CUDA graph of ``push_host`` (no host-dependent indices)."""
        nbk = rows.shape[1]
        idx = self._ring_idx.get(nbk)
        if idx is None:
            idx = self._ring_idx[nbk] = torch.arange(nbk, dtype=torch.int64, device=rows.device)
        pos = torch.remainder(self._ring_base + idx, self.ring_len)
        if nbk >= self.ring_len:
            rows, pos = rows[:, -self.ring_len:], pos[-self.ring_len:]
        self.ring.index_copy_(1, pos, rows)
        self._ring_base += nbk

    def push(self, chunk: torch.Tensor, want_series: bool = False):
        if chunk.dim() == 1:
            chunk = chunk.unsqueeze(0)
        assert chunk.shape[1] % self.block == 0, "push() takes whole blocks"
        if self.rows is None:
            band = ops.welch_band_db(chunk, self.block, self.cfg.n_fft, self.bands, float(self.fs))
        else:
            band, rows = ops.welch_band_db(chunk, self.block, self.cfg.n_fft, self.bands, float(self.fs),
                                           rows=self.rows)
            self._ring_update(rows)
        self.n_blocks += band.shape[1]
        thr = ops.live_state_step(self.states, self.lc, band[:, :, 3], want_thresholds=want_series)
        new = self._collect(self.states.det_count.cpu().numpy().astype(np.int64))
        return (new, band, thr) if want_series else new

    def _collect(self, counts):
        """Fetch the detections that completed since the last call (counts = per-stream totals, on the host)."""
        if int(counts.max(initial=0)) > self.states.max_det:
            raise RuntimeError("detection capacity exceeded; raise max_det")
        new = []
        fresh = np.nonzero(counts > self._seen)[0]
        if len(fresh):                                      # one gather + one D2H of just the new rows
            si = np.concatenate([np.full(counts[s] - self._seen[s], s) for s in fresh])
            ei = np.concatenate([np.arange(self._seen[s], counts[s]) for s in fresh])
            dev = self.states.det.device
            rows = self.states.det[torch.from_numpy(si).to(dev), torch.from_numpy(ei).to(dev)].cpu().numpy()
            for s, r in zip(si, rows):
                new.append((int(s), DetectedMeteor(*[float(v) for v in r])))
        self._seen = counts
        if self.rows is not None:
            self._pending += new
        return new

    def push_host(self, host_chunk: torch.Tensor):
        """Low-latency form of ``push`` for a fixed chunk shape arriving in HOST memory (``[n_streams, k*block]``
        PCM16): the H2D copy, the Welch band kernel, the state-machine kernel and the D2H of the detection counters
        are captured once in a CUDA graph and replayed per chunk, so a chunk costs one graph launch and one stream
        synchronisation instead of a dozen framework calls.  With a waterfall ring the per-bin PSD rows and the ring
        update (device-side write position) are part of the same graph."""
        if host_chunk.dim() == 1:
            host_chunk = host_chunk.unsqueeze(0)
        assert host_chunk.dtype == torch.int16 and not host_chunk.is_cuda
        assert host_chunk.shape[1] % self.block == 0, "push_host() takes whole blocks"
        g = getattr(self, "_graph", None)
        if g is None or g["shape"] != tuple(host_chunk.shape):
            g = self._capture(tuple(host_chunk.shape))
        g["host_in"].copy_(host_chunk)
        g["graph"].replay()                                   # launches on the current stream
        torch.cuda.current_stream(g["dev_in"].device).synchronize()
        self.n_blocks += host_chunk.shape[1] // self.block
        return self._collect(g["host_counts"].numpy().astype(np.int64))

    def _capture(self, shape):
        dev = self.states.buf.device
        host_in = torch.empty(shape, dtype=torch.int16).pin_memory()
        host_counts = torch.zeros((shape[0],), dtype=torch.int32).pin_memory()
        dev_in = torch.empty(shape, dtype=torch.int16, device=dev)
        stream = torch.cuda.Stream(device=dev)
        real = self.states

        def body():
            dev_in.copy_(host_in, non_blocking=True)
            if self.rows is None:
                band = ops.welch_band_db(dev_in, self.block, self.cfg.n_fft, self.bands, float(self.fs))
            else:
                band, rows = ops.welch_band_db(dev_in, self.block, self.cfg.n_fft, self.bands, float(self.fs),
                                               rows=self.rows)
                self._ring_update(rows)
            ops.live_state_step(self.states, self.lc, band[:, :, 3])
            host_counts.copy_(self.states.det_count, non_blocking=True)

        # warm up (plans, kernel attributes, allocator) on a scratch state so the real stream state does not advance
        self.states = ops.LiveStates(real.n_streams, dev, max_det=real.max_det)
        if self.rows is not None:                           # ... and on a scratch ring
            real_ring, real_base = self.ring, self._ring_base
            self.ring, self._ring_base = torch.empty_like(real_ring), torch.zeros_like(real_base)
        host_in.zero_()
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            for _ in range(2):
                body()
        stream.synchronize()
        self.states = real
        if self.rows is not None:
            self.ring, self._ring_base = real_ring, real_base
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            body()
        self._graph = dict(shape=shape, graph=graph, stream=stream, host_in=host_in, host_counts=host_counts,
                           dev_in=dev_in)
        return self._graph

    def export_ready(self):
        """Spectrogram crops of detections whose window [t_start - before, t_stop + after] now lies inside the
        waterfall ring (the reference's export condition, processor.py:304): list of dicts with the stream,
        the DetectedMeteor, ``db`` = [n_bins, n_cols] PSD in dB (CUDA tensor), ``times`` (block end times of
        the columns) and ``freqs``.  Each detection is returned once."""
        assert self.rows is not None, "construct LiveDetector with waterfall=ConfigVisualization(...)"
        bs = self.cfg.proc_block_sec
        n_in_ring = min(self.n_blocks, self.ring_len)
        first_blk = self.n_blocks - n_in_ring                     # oldest block still in the ring
        times = (np.arange(first_blk, self.n_blocks) * self.block + self.block) / self.fs   # processor.py:182
        out, keep = [], []
        for s, dm in self._pending:
            t0 = dm.time_start - self.export.time_before_meteor_sec
            t1 = dm.time_stop + self.export.time_after_meteor_sec
            if n_in_ring and times[0] <= t0 <= times[-1] and times[0] <= t1 <= times[-1]:
                sel = np.nonzero((times >= t0) & (times <= t1))[0]
                pos = torch.from_numpy(((first_blk + sel) % self.ring_len).astype(np.int64)).to(self.ring.device)
                out.append(dict(stream=s, meteor=dm, db=self.ring[s].index_select(0, pos).t().contiguous(),
                                times=times[sel], freqs=self.row_freqs))
            else:
                keep.append((s, dm))
        self._pending = keep
        return out


def wav_file_process(
        wav_file_path: str,
        config_detection: ConfigDetection,
        config_visualization: ConfigVisualization,
        config_spec_export: ConfigSpecExport,
        wav_file_start_sec: float = 0,
        wav_file_stop_sec: float = -1,
        *, device="cuda", quiet: bool = False):
    say = (lambda *a, **k: None) if quiet else print
    assert os.path.exists(wav_file_path), f"File not found: {wav_file_path}"
    if config_spec_export.output_dir != "":
        assert os.path.exists(
            config_spec_export.output_dir), f"Output Directory not found: {config_spec_export.output_dir}"

    edges, bands = band_bins(config_detection, 4000)
    say("Freq MS Min: ", edges[0][0])
    say("Freq MS Max: ", edges[0][1])
    say("Freq Noise 1 Min: ", edges[1][0])
    say("Freq Noise 1 Max: ", edges[1][1])
    say("Freq Noise 2 Min: ", edges[2][0])
    say("Freq Noise 2 Max: ", edges[2][1])
    say("Waterfall Win Size: ", int(config_visualization.max_range_sec / config_detection.proc_block_sec))
    say("Avg Win Size: ", int(config_detection.avg_win_sec / config_detection.proc_block_sec))

    file_sample_rate, file_data = read_wav(wav_file_path)
    assert file_sample_rate == 4000, f"Invalid Sample Rate: {file_sample_rate}"          # processor.py:66
    start = int(wav_file_start_sec * file_sample_rate)                                    # processor.py:68-71
    stop = None if wav_file_stop_sec == -1 else int(wav_file_stop_sec * file_sample_rate)
    file_data = file_data[start:stop]
    if len(file_data.shape) > 1:
        say("WARNING: Multichannel file detected. Using first channel only.")
        file_data = file_data[:, 0]
    file_block_size = int(config_detection.proc_block_sec * file_sample_rate)
    say("File Samplerate: ", file_sample_rate)
    say("File Blockgröße: ", file_block_size)
    say("File Dauer: ", len(file_data) / file_sample_rate)
    if config_visualization.enable_ui_plots or config_spec_export.output_dir != "":
        say("[ms_b200] the matplotlib UI / JPG export is out of scope of the GPU path and was skipped")

    if not torch.cuda.is_available():
        raise RuntimeError("wav_file_process needs a CUDA device: the B200 detection path has no CPU fallback")
    n_blocks = 0 if len(file_data) < file_block_size else (len(file_data) - file_block_size) // file_block_size + 1
    used = np.ascontiguousarray(file_data[:n_blocks * file_block_size])
    if used.dtype == np.int16 or used.dtype == np.float32:
        host = torch.from_numpy(used.copy())
    elif used.dtype == np.int32:                      # soundfile would scale by 2^-31
        host = torch.from_numpy((used.astype(np.float64) / 2147483648.0).astype(np.float32))
    else:
        host = torch.from_numpy(used.astype(np.float32))
    det = LiveDetector(config_detection, fs=file_sample_rate, n_streams=1, device=device)
    out = []
    if n_blocks > 0:
        for _, dm in det.push(host.to(device).reshape(1, -1)):
            out.append(dm)
            say("Detected Meteor:", dm, "Now Detected Meteors:", len(out))               # processor.py:492
    return out
