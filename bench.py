#!/usr/bin/env python
"""Benchmark of the meteor-scatter detection hot path on B200 (BASELINE.json metric).

One "step" = one pass of STFT band power -> delta -> adaptive threshold ->
events -> hourly [Anzahl, Kritisch] histogram over one batch of synthetic
beacon audio (configs[1]: 24 h = 288 five-minute 6 kHz PCM16 files per GPU;
weak scaling: rank r owns day r, hourly counts are merged with one NCCL reduce
after the last step, inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
The timed region of K steps is repeated ``--reps`` times (each repetition
bracketed by a barrier + synchronize); ``ms_per_step`` is the median over the
repetitions of the max over ranks, and every repetition is listed.
"""
from __future__ import annotations

import argparse
import contextlib
import datetime
import hashlib
import json
import os
import shutil
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 6000
FILE_SECONDS = 300
SAMPLES_PER_FILE = FS * FILE_SECONDS          # 1 800 000
FILES_PER_GPU = 288                           # 24 h
BLOCK = 1200
ALGO_BYTES_PER_BLOCK = 1024 * 2 + 8           # min(nfft, block)*2 B read + 2 fp32 written
METRIC = "Msamples/s through STFT+band-power+detect at 1/2/4/8 B200; % HBM roofline"
T0 = datetime.datetime(2025, 6, 1, 0, 0, 0)
EPOCH = datetime.datetime(1970, 1, 1)
DAY_SEED = 1234                               # day d of the synthetic archive is generated from seed DAY_SEED + d
ARCHIVE_DAYS = 30
# reference mb_files parameter set (dsp/src/main.py:865-899)
REF_KW = dict(block_duration_sec=0.2, freq_band=(993, 1013), noise_band=(690, 710), n_fft=512,
              threshold_std_factor=4, flag_adaptive_threshold=True, threshold_estimation_window_sec=120,
              threshold_freeze_before_detection_sec=3, threshold_freeze_after_detection_sec=20,
              threshold_fixed_init_duration_sec=10)


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def file_start_us(day: int, i: int) -> int:
    d = (T0 - EPOCH) + datetime.timedelta(days=day, seconds=i * FILE_SECONDS)
    return (d.days * 86400 + d.seconds) * 1_000_000 + d.microseconds


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons through NVML while the GPU works."""

    def __init__(self, index: int, period_s: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples = []          # (t, sm_mhz, reasons_bitmask)
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                try:
                    mem = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM))
                except Exception:
                    mem = None
                self.samples.append((time.perf_counter(), mhz, rs, pw, mem))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self, windows):
        """Median SM clock over samples inside any (t0, t1) window."""
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)]
        if not self.ok or not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        mhz = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = [n for b, n in names.items() if bits & b and n != "gpu_idle"]
        out = {"sm_mhz": mhz[len(mhz) // 2], "sm_min_mhz": mhz[0], "sm_max_mhz": self.max_mhz, "reasons": reasons,
               "samples": len(inside)}
        pw = [s[3] for s in inside if len(s) > 3 and s[3] is not None]
        if pw:
            out["power_w_max"] = max(pw)
        mem = sorted(s[4] for s in inside if len(s) > 4 and s[4] is not None)
        if mem:
            out["mem_mhz"] = mem[len(mem) // 2]
            out["mem_min_mhz"] = mem[0]
        # SM clock over time: first and last third of the samples (the repetitions slow down by ~5 % after ~0.1 s)
        third = max(1, len(inside) // 3)
        out["sm_mhz_first_third"] = sorted(s[1] for s in inside[:third])[third // 2]
        out["sm_mhz_last_third"] = sorted(s[1] for s in inside[-third:])[third // 2]
        return out


# --------------------------------------------------------------------------- CPU arm (reference / oracle port)
def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def reference_staged() -> bool:
    """Is the unmodified reference (oracle/_ref, staged by oracle/stage_ref.py, or /root/reference) importable?"""
    from oracle import ref_harness
    return ref_harness.reference_available()


_REF_MOD = {}


def _reference_worker(args):
    """One worker process = the reference's own single-process path: the UNMODIFIED ``proc_wav_file``
    (dsp/src/main.py:207) on WAV files, mb_files parameters, event CSV out.  Returns per file
    (index pairs, {hour: [Anzahl, Kritisch]}) parsed from the CSV the reference wrote."""
    import csv
    from oracle import ref_harness
    paths, start_us = args
    mod = _REF_MOD.get("mod")
    if mod is None:      # (the pool's untimed spin-up call lands here: the import is not part of the timed region)
        mod = _REF_MOD["mod"] = ref_harness.load_reference_main()
    if not paths:
        return []
    out = []
    bd = REF_KW["block_duration_sec"]
    with open(os.devnull, "w") as devnull:
        for p, us in zip(paths, start_us):
            t = EPOCH + datetime.timedelta(microseconds=int(us))
            csv_path = p + ".events.csv"
            with contextlib.redirect_stdout(devnull):
                mod.proc_wav_file(p, wav_start_date_time=t, disable_show_and_write=True, out_csv_file=csv_path,
                                  **REF_KW)
            pairs, hours = [], {}
            with open(csv_path, newline="") as f:
                for row in csv.DictReader(f):
                    pairs.append((int(round(float(row["t_start"]) / bd)), int(round(float(row["t_stop"]) / bd))))
                    h = datetime.datetime.fromisoformat(row["utc_start"]).replace(minute=0, second=0, microsecond=0)
                    c = hours.setdefault(h, [0, 0])
                    c[0] += 1
                    c[1] += 1 if float(row["dur_s"]) >= 0.5 else 0
            os.unlink(csv_path)
            out.append((pairs, hours))
    return out


def _port_worker(args):
    """One worker = the oracle port of the reference's single-process algorithm (used only when the
    reference itself is not staged, and for the fp64 band powers of the parity report)."""
    from oracle import detector_a as oa
    files, start_us, want_band = args
    out = []
    for x, us in zip(files, start_us):
        t = EPOCH + datetime.timedelta(microseconds=int(us))
        r = oa.detect_wav(x, FS, 0.2, (993, 1013), (690, 710), 512, 4, wav_start_date_time=t)
        rec = (r["pairs"], oa.hourly_counts(r["detections"]))
        if want_band:
            rec = rec + (r["band_power"], r["noise_power"])
        out.append(rec)
    return out


def run_pool(worker, items, start_us, procs, extra=()):
    """Fan `items` out round-robin over `procs` worker processes; returns (seconds, per-item results)."""
    import multiprocessing as mp
    chunks = [[] for _ in range(procs)]
    cus = [[] for _ in range(procs)]
    for i, (x, u) in enumerate(zip(items, start_us)):
        chunks[i % procs].append(x)
        cus[i % procs].append(u)
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(worker, [([], []) + tuple(extra)] * procs)      # spin the workers up outside the timed region
        t0 = time.perf_counter()
        res = pool.map(worker, [(c, u) + tuple(extra) for c, u in zip(chunks, cus)])
        dt = time.perf_counter() - t0
    flat = [None] * len(items)
    for p in range(procs):
        for j, r in enumerate(res[p]):
            flat[p + j * procs] = r
    return dt, flat


def write_wavs(arrays, prefix="ms_bench_wav_"):
    """PCM16 WAV files for the reference arm (it reads files: scipy.io.wavfile.read, main.py:249)."""
    import scipy.io.wavfile as wavfile
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 3 * sum(
        a.nbytes for a in arrays) else None
    d = tempfile.mkdtemp(prefix=prefix, dir=base)
    paths = []
    for i, a in enumerate(arrays):
        p = os.path.join(d, f"f{i:04d}.wav")
        wavfile.write(p, FS, a)
        paths.append(p)
    return d, paths


def cpu_arm(files, start_us, cores):
    """Time the CPU implementation over `files` (numpy int16 arrays) on `cores` processes.
    Returns (seconds, results, kind, description)."""
    if reference_staged():
        d, paths = write_wavs(files)
        try:
            dt, res = run_pool(_reference_worker, paths, start_us, cores)
        finally:
            shutil.rmtree(d, ignore_errors=True)
        return dt, res, "reference", "unmodified dsp/src/main.py:proc_wav_file on WAV files"
    dt, res = run_pool(_port_worker, files, start_us, cores, extra=(False,))
    return dt, res, "port", "oracle port of dsp/src/main.py:352-527"


def _synth_worker(args):
    from meteor_scatter_b200.synth import synth_file
    seeds, _ = args
    return [synth_file(s, fs=FS, dur_s=FILE_SECONDS) for s in seeds]


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (unmodified proc_wav_file from
    oracle/_ref when staged, else the oracle port) on all host cores, one step = the 288 files of configs[1]."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(host_cores(), 64)
    n_files = args.files
    _, files = run_pool(_synth_worker, [1000 + i for i in range(n_files)], [0] * n_files, cores)
    start_us = [file_start_us(0, i) for i in range(n_files)]
    kind = "reference" if reference_staged() else "port"
    if kind == "reference":
        d, paths = write_wavs(files)
        items, worker, extra = paths, _reference_worker, ()
    else:
        d, items, worker, extra = None, files, _port_worker, (False,)
    try:
        for _ in range(args.warmup):
            run_pool(worker, items[:cores], start_us[:cores], cores, extra)
        tot = 0.0
        for _ in range(args.steps):
            dt, res = run_pool(worker, items, start_us, cores, extra)
            tot += dt
    finally:
        if d:
            shutil.rmtree(d, ignore_errors=True)
    ms = tot / args.steps * 1e3
    value = n_files * SAMPLES_PER_FILE / (tot / args.steps) / 1e6
    what = ("unmodified dsp/src/main.py:proc_wav_file (oracle/_ref) on WAV files" if kind == "reference"
            else "oracle port of dsp/src/main.py")
    sample = f"{n_files} five-minute files per step over {cores} worker processes ({what})"
    anz = sum(c[0] for r in res for c in r[1].values())
    krit = sum(c[1] for r in res for c in r[1].values())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, "cpu", n_files),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "hourly_counts": {"anzahl_total": anz, "kritisch_total": krit},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, impl, n_files=FILES_PER_GPU):
    return {"workload": "configs[1]: 24 h of synthetic beacon audio = 288 five-minute 6 kHz mono PCM16 files per GPU, "
                        "reference mb_files parameters (block 0.2 s = 1200 samples, rfft 1024, bands 993-1013 / "
                        "690-710 Hz, k=4, adaptive threshold 120/3/20/10 s)",
            "files_per_gpu": n_files, "samples_per_file": SAMPLES_PER_FILE, "block": BLOCK, "nfft": 1024,
            "parallelism": f"files sharded per GPU x{n_gpus}, no data-path collective, one final NCCL reduce of the hourly counts",
            "band_power_impl": impl, "l2": "inputs (1.04 GB per GPU) are larger than the 126 MB L2; no flush needed"}


# --------------------------------------------------------------------------- extra measurements (our arm)
def band_parity_report(res, band_db, noise_db, near_total, counts_host, events_host, n_s):
    """Band-power parity of the first ``n_s`` files against the fp64 oracle, reported as COUNTS (north_star:
    band power within 1e-4 relative; frames within eps of the threshold are reported separately)."""
    import numpy as np
    bad = 0
    worst = 0.0
    blocks = 0
    mism = 0
    for i in range(n_s):
        pairs, _, band, noise = res[i][0], res[i][1], res[i][2], res[i][3]
        for ours, ref in ((band_db[i], band), (noise_db[i], noise)):
            # dB difference -> relative energy difference
            rel = np.abs(np.expm1((ours.astype(np.float64) - ref) * (np.log(10.0) / 10.0)))
            bad += int((rel > 1e-4).sum())
            worst = max(worst, float(rel.max()))
            blocks += len(ref)
        if [tuple(int(v) for v in p) for p in events_host[i, :counts_host[i]]] != pairs:
            mism += 1
    return {"files_checked": n_s, "files_with_different_events": mism,
            "events_checked": int(sum(len(r[0]) for r in res[:n_s])),
            "band_values_checked": blocks, "band_values_rel_err_gt_1e-4": bad, "band_max_rel_err": worst,
            "frames_within_1e-3_dB_of_threshold_whole_batch": near_total,
            "oracle": "fp64 port of dsp/src/main.py:376-522 (pinned bit-exactly to the reference's goldens)"}


def sweep_field(x, torch, ops, peak):
    """configs[3] at 24 h scale: nfft 1024..16384 x overlap 50/75/90 %, periodic Hann, signal band = carrier
    +/- 10 Hz, noise band 690-710 Hz, every supported implementation; algorithmic bytes = unique input + 8 B/frame."""
    import numpy as np
    n_files, spf = x.shape
    rows = []
    for nfft in (1024, 2048, 4096, 8192, 16384):
        w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(nfft) / nfft)
        freqs = np.fft.rfftfreq(nfft, 1 / FS)
        sig = np.nonzero((freqs >= 993) & (freqs <= 1013))[0]
        noi = np.nonzero((freqs >= 690) & (freqs <= 710))[0]
        for ov in (0.5, 0.75, 0.9):
            hop = max(8, int(round(nfft * (1 - ov) / 8)) * 8)
            spec = ops.BandSpec.stft(nfft, hop, w, sig, noi, fs=FS)
            nfr = spec.n_blocks(spf)
            unique_bytes = n_files * (spf * 2 + nfr * 8)
            # K1; K2 (resident basis, frames re-read); K2S (hop segments, shifted products); the frequency-domain-window
            # form (one unwindowed product per hop segment + rotation/window combine) where hop | frame
            for impl in ("fft", "k2", "seg", "rot"):
                if (impl == "k2" and not ops.k2_supported(x, spec)) or (impl == "seg" and not ops.seg_supported(x, spec)) \
                        or (impl == "rot" and not ops.rot_supported(x, spec)):
                    continue
                fn = lambda: ops.band_power(x, spec, impl=impl)          # noqa: E731
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 3
                a.record()
                for _ in range(n):
                    fn()
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / n
                rows.append({"nfft": nfft, "overlap": ov, "hop": hop, "impl": impl, "bins": len(sig) + len(noi),
                             "frames": n_files * nfr, "ms": round(ms, 4),
                             "Msamples_per_s": round(n_files * spf / ms / 1e3, 0),
                             "hbm_frac_unique_bytes": round(unique_bytes / (ms * 1e-3) / 1e9 / peak, 4)})
    return {"workload": "configs[3] on this rank's 24 h batch (288 files resident in HBM)", "points": rows}


def streaming_field(torch, chunks=10000):
    """configs[4]: 1 s chunks of detector-B audio (4 kHz, 5 blocks of 800 samples) arriving in pinned host memory;
    wall time from "chunk in host memory" to "its detections on the host", p50/p99."""
    import numpy as np
    from meteor_scatter_b200.dsp.src.live.backend.aggregates import ConfigDetection, ConfigVisualization
    from meteor_scatter_b200.dsp.src.live.backend.processor import LiveDetector
    from meteor_scatter_b200.synth import synth_file
    fs, chunk = 4000, 4000
    base = synth_file(5, fs=fs, dur_s=600.0, carrier_hz=1020.0, rate_per_hour=900.0)
    n_base = len(base) // chunk
    cfg = ConfigDetection(proc_block_sec=0.2, n_fft=4096, detection_db_over_noise_mean_min=1,
                          detection_dur_min_sec=0.5, signal_freq=1020)
    out = []
    for streams, ring in ((1, False), (64, False), (1, True), (64, True)):
        det = LiveDetector(cfg, fs=fs, n_streams=streams, device="cuda",
                           waterfall=ConfigVisualization(enable_ui_plots=False) if ring else None)
        pinned = torch.empty((streams, chunk), dtype=torch.int16).pin_memory()
        lat = []
        n_det = 0
        for i in range(chunks + 50):
            src = torch.from_numpy(base[(i % n_base) * chunk:(i % n_base + 1) * chunk])
            pinned.copy_(src.unsqueeze(0).expand(streams, -1))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            new = det.push_host(pinned)
            t1 = time.perf_counter()
            n_det += len(new)
            if i >= 50:
                lat.append((t1 - t0) * 1e6)
        lat = np.sort(np.array(lat))
        out.append({"streams": streams, "waterfall_ring": ring, "chunks": chunks, "detections": n_det,
                    "p50_us": round(float(lat[len(lat) // 2]), 1), "p99_us": round(float(lat[int(len(lat) * 0.99)]), 1),
                    "max_us": round(float(lat[-1]), 1)})
    return {"workload": "configs[4]: 1 s chunks, detector B (4 kHz, 5 x 800-sample blocks, Welch nfft 4096), "
                        "LiveDetector.push_host (one CUDA graph per chunk: H2D, Welch bands, state machine, "
                        "[waterfall ring], D2H of the detection counters)", "cases": out}


def pipelined_field(det, x, start_us, hour0, n_hours, torch, steps=50, reps=7):
    """Steady-state form of the pass (what the archive path runs, pipeline.PassPipeline): batch i's detect + hourly
    kernel on a side stream under batch i+1's band-power kernel.  Same work per step as `value`; reported beside it."""
    from meteor_scatter_b200.pipeline import PassPipeline
    n_files, spf = x.shape
    pipe = PassPipeline(det, n_files, spf, n_hours, x.device, depth=2)
    for _ in range(5):
        slot = pipe.submit(x, start_us, hour0)
    pipe.drain()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            slot = pipe.submit(x, start_us, hour0)
        pipe.drain()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b) / steps)
    h = pipe.wait(slot)[1]
    ref = torch.zeros((n_hours, 2), dtype=torch.int32, device=x.device)
    det.run_pass(x, start_us, hour0, n_hours, ref)           # the in-line pass on the same batch
    torch.cuda.synchronize()
    same = bool(torch.equal(h.cpu(), ref.cpu()))
    ms = sorted(times)[len(times) // 2]
    return {"what": "PassPipeline: detect(i) on a side stream under band power(i+1), same work per step as `value`",
            "ms_per_step": round(ms, 6), "value": round(n_files * spf / (ms * 1e-3) / 1e6, 1), "unit": "Msamples/s",
            "steps": steps, "reps": reps, "ms_per_step_each_rep": [round(t, 6) for t in times],
            "histogram_equals_inline_pass": same}


def detector_c_field(torch, ops, peak, n_seg=2048):
    """Detector C numeric stage in batch (SURVEY 8 C-stft, prime_detection.py:67-92): n_seg 30 s segments of 5 kHz PCM16
    -> one-sided PSD rows 800-1200 Hz [164 x 145] + the 250-800 Hz noise-band sum (specgram NFFT 2048, noverlap 1024,
    np.hanning) through ops.psd_spectrogram; roofline fraction on unique input + output bytes."""
    import numpy as np
    from meteor_scatter_b200.synth import synth_batch_torch
    n, fs, nfft = 150_000, 5000, 2048
    xc = synth_batch_torch(n_seg, n, fs=fs, carrier_hz=1000.0, rate_per_hour=600.0, seed=3, device="cuda")
    freqs = np.fft.rfftfreq(nfft, 1 / fs)
    rows = np.nonzero((freqs >= 800) & (freqs <= 1200))[0]
    nk = np.nonzero((freqs >= 250) & (freqs <= 800))[0]
    w = np.hanning(nfft)

    def run():
        return ops.psd_spectrogram(xc, float(fs), nfft, nfft // 2, w, int(rows[0]), int(rows[-1]), int(nk[0]), int(nk[-1]))

    for _ in range(3):
        psd, noise = run()
    times = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            psd, noise = run()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b) / 10)
    ms = sorted(times)[len(times) // 2]
    # spot check of one segment against numpy in fp64 (same framing and scaling as mlab.specgram / scipy 'density')
    seg = xc[7].cpu().numpy().astype(np.float64)
    nf = (n - nfft // 2) // (nfft // 2)
    fr = np.lib.stride_tricks.sliding_window_view(seg, nfft)[::nfft // 2][:nf] * w
    ref = (np.abs(np.fft.rfft(fr, axis=1)) ** 2).T / (fs * np.sum(w * w))
    ref[1:-1] *= 2
    got = psd[7].cpu().numpy().astype(np.float64)
    rel = np.abs(got - ref[rows]) / ref[rows]
    bytes_algo = xc.numel() * 2 + psd.numel() * 4
    del xc, psd
    return {"workload": f"{n_seg} segments x {n} samples of 5 kHz PCM16, nfft 2048, 50 % overlap, 164 PSD rows + noise band",
            "kernel": "psd_warp_kernel" if os.environ.get("MS_PSD_IMPL") != "fft" else "stft_kernel",
            "ms": round(ms, 4), "segments_per_s": round(n_seg / (ms * 1e-3), 1),
            "Msamples_per_s": round(n_seg * n / (ms * 1e-3) / 1e6, 1), "algorithmic_bytes": int(bytes_algo),
            "hbm_frac": round(bytes_algo / (ms * 1e-3) / 1e9 / peak, 4),
            "spot_check": {"values": int(rel.size), "outside_1e-4": int((rel > 1e-4).sum()), "max_rel_err": float(rel.max())}}


def ingest_field(x, torch, n_files):
    """from WAV FILES on disk (page cache warm) to the day CSVs through batch.process_files -- header
    parse, reader threads into the pinned ring, H2D, kernels, D2H of the event lists, YYYYMMDD.csv."""
    from meteor_scatter_b200.batch import process_files
    from meteor_scatter_b200.wavio import write_wav_pcm16
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 3 * x.numel() * 2 else None
    root = tempfile.mkdtemp(prefix="ms_bench_ingest_", dir=base)
    try:
        paths = []
        for i in range(n_files):
            t = T0 + datetime.timedelta(seconds=FILE_SECONDS * i)
            p = os.path.join(root, f"expoFull_gqrx_{t.strftime('%Y%m%d_%H%M%S')}_49969000.wav")
            write_wav_pcm16(p, FS, x[i].cpu().numpy())
            paths.append(p)
        os.makedirs(os.path.join(root, "csv"))
        process_files(paths, csv_folder=None)                 # warm-up: ring allocation (pinning), plans
        runs = []
        for _ in range(5):
            torch.cuda.synchronize()
            t = time.perf_counter()
            r = process_files(paths, csv_folder=os.path.join(root, "csv"))
            torch.cuda.synchronize()
            runs.append(time.perf_counter() - t)
        runs.sort()
        med = runs[len(runs) // 2]
        return {"workload": f"{n_files} five-minute PCM16 WAV files ({n_files * SAMPLES_PER_FILE * 2} B, page cache warm, "
                            f"{'tmpfs' if base else 'tmp dir'}) -> batch.process_files -> day CSVs",
                "value": n_files * SAMPLES_PER_FILE / med / 1e6, "unit": "Msamples/s", "seconds_median": med,
                "seconds_each_run": [round(v, 4) for v in runs], "file_bytes_per_s_GB": n_files * SAMPLES_PER_FILE * 2 / med / 1e9,
                "events": int(r["hist"][:, 0].sum()), "csv_files": len(r["csv_files"]), "host_cpus": host_cores(),
                "how": "headers parsed up front; ms_read_files (16 native pread threads) fills a persistent ring of 3 pinned "
                       "slots of 24 files; H2D on a copy stream; band power + detect + hourly counts per chunk; event "
                       "lists of chunk k-1 unpacked while chunk k runs"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def archive_field(det, dev, rank, world, torch, dist, days=ARCHIVE_DAYS):
    """configs[2]: a 30-day archive (8640 files), day d generated from seed DAY_SEED + d and owned by rank
    d % world; every day is one pass accumulated into the rank's [720 x 2] histogram, merged with ONE NCCL
    sum-reduce.  The sha256 of the merged histogram must not depend on the number of GPUs."""
    from meteor_scatter_b200.synth import synth_batch_torch
    n_hours = days * 24
    archive = torch.zeros((n_hours, 2), dtype=torch.int32, device=dev)
    day_hist = torch.zeros_like(archive)
    ms = 0.0
    gen_s = 0.0
    my_days = list(range(rank, days, world))
    for d in my_days:
        tg = time.perf_counter()
        xd = synth_batch_torch(FILES_PER_GPU, SAMPLES_PER_FILE, fs=FS, seed=DAY_SEED + d, device=dev)
        us = torch.tensor([file_start_us(d, i) for i in range(FILES_PER_GPU)], dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        gen_s += time.perf_counter() - tg
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = det.run_pass(xd, us, T0, n_hours, day_hist)
        archive += day_hist
        b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
        r.det.check_capacity()
        del xd
    red_ms = 0.0
    if world > 1:
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dist.reduce(archive, dst=0, op=dist.ReduceOp.SUM)          # the one collective of the path
        b.record()
        torch.cuda.synchronize()
        red_ms = a.elapsed_time(b)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank != 0:
        return None
    h = archive.cpu().numpy()
    total = days * FILES_PER_GPU * SAMPLES_PER_FILE
    return {"workload": f"configs[2]: {days}-day synthetic archive = {days * FILES_PER_GPU} five-minute files, day d "
                        f"(seed {DAY_SEED}+d) on rank d % {world}, one NCCL sum-reduce of the [{n_hours} x 2] histogram",
            "files": days * FILES_PER_GPU, "n_gpus": world, "ms_compute_max_over_ranks": ms, "reduce_ms": red_ms,
            "value": total / ((ms + red_ms) * 1e-3) / 1e6, "unit": "Msamples/s",
            "anzahl_total": int(h[:, 0].sum()), "kritisch_total": int(h[:, 1].sum()),
            "hist_sha256": hashlib.sha256(h.astype("<i4").tobytes()).hexdigest(),
            "note": "days are generated on the device one at a time (generation is outside the timed passes: "
                    f"{gen_s:.1f} s on rank 0); the hash is identical for every N by construction of the seeds"}


def h2d_ceiling(host_pcm, dev_buf, torch, dist, world, reps=5):
    """Plain cudaMemcpyAsync of the same pinned bytes the e2e path moves, all ranks at once: the H2D ceiling."""
    s = torch.cuda.current_stream()
    flat = host_pcm.view(-1)[:dev_buf.numel()]
    for _ in range(2):
        dev_buf.copy_(flat, non_blocking=True)
    s.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        dev_buf.copy_(flat, non_blocking=True)
    b.record()
    s.synchronize()
    ms = a.elapsed_time(b) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev_buf.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def traffic_from_profile():
    """ncu DRAM bytes per launch of the dominant kernel (profiles/traffic.json), with a staleness flag: the file
    is stamped with the hash of the sources the measured library was built from."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(tp) as f:
            t = json.load(f)
    except Exception:
        return None, None
    from meteor_scatter_b200 import build as _build
    stale = t.get("source_hash") != _build.source_hash(_build.K2_SOURCES)
    return t.get("dram_bytes_per_launch"), {"file": "profiles/traffic.json", "source_hash": t.get("source_hash"),
                                            "stale": stale}


# --------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--reps", type=int, default=15, help="repetitions of the K-step timed region (median reported)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "fft", "tc"])
    ap.add_argument("--files", type=int, default=FILES_PER_GPU, help="files per GPU (default: 24 h)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline-depth", type=int, default=2, help="batches in flight with --pipeline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the archive / sweep / streaming / ingest fields")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--pipeline", action="store_true",
                    help="run each batch's detect stage on a side stream under the next batch's STFT (PassPipeline, "
                         "ms_detector_a_pass_overlapped_i16); opt-in because it times the 4-fix-up-warp instantiation "
                         "of the band-power kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from meteor_scatter_b200 import _lib, ops
    from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams, PassPipeline, hour_index
    from meteor_scatter_b200.synth import synth_batch_torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.load()
    from meteor_scatter_b200.batch import bind_host_to_gpu
    numa_bound = False
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_files = args.files
    impl = "tc" if args.impl == "ours" else args.impl
    params = DetectorAParams()
    det = DetectorA(params, impl=impl)
    nb = det.spec.n_blocks(SAMPLES_PER_FILE)
    n_hours_local = (n_files * FILE_SECONDS + 3599) // 3600
    n_hours = n_hours_local * world
    hour0 = T0
    # rank r owns day r (seed DAY_SEED + r): rank 0's day is the N=1 day at every N, files are contiguous from
    # T0 + r days so hours fill exactly (12 files/hour)
    day = rank
    start_us_host = [file_start_us(0, day * n_files + i) for i in range(n_files)]
    start_us = torch.tensor(start_us_host, dtype=torch.int64, device=dev)

    x = synth_batch_torch(n_files, SAMPLES_PER_FILE, fs=FS, seed=DAY_SEED + day, device=dev)
    torch.cuda.synchronize()
    hist = torch.zeros((n_hours, 2), dtype=torch.int32, device=dev)
    warm = torch.zeros_like(hist)               # N>1: target of the communicator warm-up reduce
    ev_k2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev_k2:          # create the CUDA events now so their handles can cross the C-ABI
        a.record()
        b.record()
    hourly = dict(file_start_us=start_us, hour0=hour_index(hour0), n_hours=n_hours, out=hist)

    # --pipeline (opt-in): overlap detect(i) with band power(i+1) through PassPipeline
    pipe = PassPipeline(det, n_files, SAMPLES_PER_FILE, n_hours, dev, depth=args.pipeline_depth) if (impl == "tc" and args.pipeline) else None
    last = {"mode": "pass", "slot": 0}

    def step(i=None):
        """One pass over the batch.  tc: ONE C-ABI call enqueueing the band-power kernel (which also clears the
        histogram) and the detect+hourly kernel.  Ranks never talk during a step; N>1: one NCCL sum-reduce of the
        [hours x 2] histogram to rank 0 after the last step (inside the timed region).  On every 8th timed step
        CUDA events are recorded inside the C-ABI right around the band-power kernel: the roofline samples."""
        sampled = i is not None and i % 8 == 0
        if impl == "tc" and pipe is not None:
            evs = ev_k2[i] if sampled else (None, None)    # sampled: submit() lets the previous detect finish first
            last["slot"], last["mode"] = pipe.submit(x, start_us, hour0, ev_begin=evs[0], ev_end=evs[1]), "pipe"
            return None
        last["mode"] = "pass"
        if impl == "tc":
            evs = ev_k2[i] if sampled else (None, None)
            return det.run_pass(x, start_us, hour0, n_hours, hist, ev_begin=evs[0], ev_end=evs[1]).det
        bufs = det._buffers(n_files, nb, dev)
        hist.zero_()
        if i is not None:
            ev_k2[i][0].record()
        band_db, noise_db = ops.band_power(x, det.spec, impl=impl, out=(bufs["band"], bufs["noise"]))
        if i is not None:
            ev_k2[i][1].record()
        W, before, after, fixed = params.block_counts()
        return ops.detect(band_db, noise_db, params.threshold_std_factor, adaptive=True, window_blocks=W,
                          before_blocks=before, after_blocks=after, fixed_blocks=fixed, max_events=det.cap(nb),
                          workspace=det._ws, out=bufs["det"],
                          hourly=dict(hourly, block_duration_sec=params.block_duration_sec))

    def drain():
        if pipe is not None:
            pipe.drain()

    det._buffers(n_files, nb, dev)
    # clocks are reported for rank 0's GPU; the other ranks do not poll NVML (eight pollers on one host take the
    # driver's locks 4000 times a second next to eight enqueue loops)
    sampler = ClockSampler(local_rank, period_s=0.004)
    if rank != 0:
        sampler.ok = False
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
        if world > 1:      # warm the communicator too (NCCL sets channels up lazily on the first collective)
            dist.reduce(warm, dst=0, op=dist.ReduceOp.SUM)
    drain()

    # ---- the timed region: K steps (+ the one reduce at N>1), repeated `reps` times ----
    reps = max(1, args.reps)
    rep_ms, red_ms, k2_samples, enq_ms = [], [], [], []
    own_hist = None
    windows = []
    for r in range(reps):
        barrier()
        tw0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_enq = time.perf_counter()
        for i in range(args.steps):
            d_last = step(i)
        enq_ms.append((time.perf_counter() - t_enq) * 1e3)      # host time to ENQUEUE the K steps (no synchronisation)
        drain()
        final_hist = pipe.wait(last["slot"])[1] if (pipe is not None and last["mode"] == "pipe") else hist
        if world > 1:
            if r == 0:
                own_hist = final_hist.clone()            # (first repetition only, 5.8 KB) this rank's own counts
            # the path's one exchange step (north_star: "one final NCCL gather merges per-hour counts"): every
            # rank's [hours x 2] histogram is summed onto rank 0 once, inside the timed region
            r0.record()
            dist.reduce(final_hist, dst=0, op=dist.ReduceOp.SUM)
            r1.record()
        e1.record()
        barrier()
        windows.append((tw0, time.perf_counter()))
        rep_ms.append(e0.elapsed_time(e1))
        red_ms.append(r0.elapsed_time(r1) if world > 1 else 0.0)
        # the band-power kernel's own duration: every sampled launch of every repetition (the event pairs are reused)
        k2_samples += [ev_k2[i][0].elapsed_time(ev_k2[i][1]) for i in range(args.steps) if impl != "tc" or i % 8 == 0]
    t_wall0, t_wall1 = windows[0][0], windows[-1][1]
    times = torch.tensor([rep_ms, red_ms, enq_ms], dtype=torch.float64, device=dev)  # [3, reps]
    if world > 1:
        allt = [torch.empty_like(times) for _ in range(world)]
        dist.all_gather(allt, times)
        allt = torch.stack(allt).cpu().numpy()                                       # [world, 2, reps]
    else:
        allt = times.cpu().numpy()[None]
    per_rep = allt[:, 0, :].max(axis=0)                                              # max over ranks, per repetition
    order = np.argsort(per_rep)
    med = int(order[len(order) // 2])
    elapsed_ms = float(per_rep[med])
    k2_ms = sum(k2_samples) / len(k2_samples)
    if pipe is not None and last["mode"] == "pipe":
        res_last, hist_last = pipe.wait(last["slot"])
        d_last = res_last.det
        hist_host = hist_last.cpu().numpy().copy()
    else:
        hist_host = hist.cpu().numpy().copy()
    d_last.check_capacity()
    counts_host = d_last.counts.cpu().numpy()
    events_host = d_last.events.cpu().numpy()

    # ---- N>1: the merged histogram must be the sum of the per-rank ones, and rank 0's day the N=1 day ----
    multi = None
    if world > 1:
        own = [torch.empty_like(own_hist) for _ in range(world)]
        dist.all_gather(own, own_hist)
        own = torch.stack(own).cpu().numpy().astype(np.int64)                        # [world, hours, 2]
        if rank == 0:
            assert np.array_equal(own.sum(axis=0), hist_host.astype(np.int64)), \
                "NCCL-reduced histogram differs from the sum of the per-rank histograms"
            for r in range(world):      # a rank's counts fall only into its own day's hours
                outside = own[r].sum() - own[r, r * n_hours_local:(r + 1) * n_hours_local].sum()
                assert outside == 0, f"rank {r} counted events outside its day"
            multi = {"per_rank_day_totals": [[int(own[r, :, 0].sum()), int(own[r, :, 1].sum())] for r in range(world)],
                     "reduced_equals_sum_of_ranks": True,
                     "rank0_day_totals": [int(own[0, :, 0].sum()), int(own[0, :, 1].sum())],
                     "note": "rank 0's day (seed 1234) is the N=1 workload: rank0_day_totals must equal the N=1 "
                             "line's hourly_counts at every N"}

    # ---- the same pass on the dense resident layout the ingest path produces (extra field, not `value`) ----
    dense = None
    if impl == "tc" and world == 1 and not args.no_extras:
        det_d = det.dense_variant()
        wl = det.spec.win_len
        x_d = x.view(n_files, nb, BLOCK)[:, :, :wl].contiguous().view(n_files, nb * wl)
        hist_d = torch.zeros_like(hist)
        evd = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in evd:
            a.record()
            b.record()
        for _ in range(args.warmup):
            det_d.run_pass(x_d, start_us, hour0, n_hours, hist_d)
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for i in range(args.steps):
            if i % 8 == 0:
                det_d.run_pass(x_d, start_us, hour0, n_hours, hist_d, ev_begin=evd[i][0], ev_end=evd[i][1])
            else:
                det_d.run_pass(x_d, start_us, hour0, n_hours, hist_d)
        d1.record()
        torch.cuda.synchronize()
        ms_d = d0.elapsed_time(d1) / args.steps
        tm = [evd[i] for i in range(args.steps) if i % 8 == 0]
        k_d = sum(a.elapsed_time(b) for a, b in tm) / len(tm)
        assert np.array_equal(hist_d.cpu().numpy(), hist_host), "dense-layout pass disagrees with the PCM-layout pass"
        dense = {"layout": f"[files][blocks][{wl}] PCM16 (only the samples the transform reads are kept resident)",
                 "value": n_files * SAMPLES_PER_FILE / (ms_d * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_d,
                 "kernel_ms": k_d, "hbm_bytes_resident": int(x_d.numel() * 2),
                 "roofline_frac": n_files * nb * ALGO_BYTES_PER_BLOCK / (k_d * 1e-3) / 1e9 / measured_hbm_peak()[0]}
        del x_d

    # ---- end to end through the public API: pinned host PCM -> H2D -> kernels -> D2H results ----
    e2e = None
    e2e_windows = []
    if not args.no_e2e:
        chunk_files = 24
        # the pinned "recordings" are allocated on the NUMA node next to this rank's GPU; the affinity is restored
        # afterwards so the cpu_baseline pool still sees every core
        cpus_before = os.sched_getaffinity(0)
        numa_bound = bind_host_to_gpu(local_rank) if not args.no_numa else False
        host_pcm = torch.empty((n_files, SAMPLES_PER_FILE), dtype=torch.int16).pin_memory()
        os.sched_setaffinity(0, cpus_before)
        host_pcm.copy_(x)                      # (setup) the "recordings" now live in host memory
        out_host = {}

        def e2e_step():
            """Public API: DetectorA.run_host (pinned host PCM -> strided DMA of the used samples,
            double buffered and overlapped with the band-power kernel -> detect -> D2H of the results)."""
            out_host.update(det.run_host(host_pcm, start_us, hour0, n_hours, chunk_files=chunk_files))

        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(2):
            e2e_step()
        barrier()
        tw0 = time.perf_counter()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(e2e_steps):
            e2e_step()
        if world > 1:
            dist.reduce(det._host_state["hist"], dst=0, op=dist.ReduceOp.SUM)
            out_host["hist"].copy_(det._host_state["hist"], non_blocking=True)
        a1.record()
        barrier()
        e2e_windows.append((tw0, time.perf_counter()))
        e2e_ms = a0.elapsed_time(a1) / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        assert np.array_equal(out_host["counts"].numpy(), counts_host), "e2e path and resident path disagree"
        if rank == 0:
            assert np.array_equal(out_host["hist"].numpy(), hist_host), "e2e histogram differs from the resident path"
        h2d_bytes = int(n_files * nb * det.spec.win_len * 2)
        # the ceiling next to the number: a plain cudaMemcpyAsync of the same byte count from the same pinned buffer,
        # all ranks at once
        ceil_buf = torch.empty(h2d_bytes // 2, dtype=torch.int16, device=dev)
        ceil_ms = h2d_ceiling(host_pcm, ceil_buf, torch, dist, world)
        del ceil_buf
        e2e = {"value": world * n_files * SAMPLES_PER_FILE / (e2e_ms * 1e-3) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": int(sum(t.numel() * t.element_size() for t in out_host.values())),
               "ms_per_step": e2e_ms, "steps": e2e_steps,
               "host_numa_bound": numa_bound,
               "h2d_ceiling": {"ms": ceil_ms, "gbs_per_gpu": h2d_bytes / (ceil_ms * 1e-3) / 1e9,
                               "gbs_all_gpus": world * h2d_bytes / (ceil_ms * 1e-3) / 1e9,
                               "e2e_frac_of_ceiling": ceil_ms / e2e_ms,
                               "how": "one contiguous cudaMemcpyAsync of h2d_bytes_per_step from the same pinned "
                                      "buffer on every rank at once, max over ranks"},
               "how": f"DetectorA.run_host: pinned host PCM16 ({n_files * SAMPLES_PER_FILE * 2} B), strided DMA of the "
                      f"{det.spec.win_len} samples per {BLOCK}-sample block the transform reads, {chunk_files}-file "
                      f"chunks double-buffered and overlapped with the band-power kernel, results copied back"}
        del host_pcm

    sampler.stop()
    sampler.join(timeout=1.0)
    clocks = sampler.summary(windows)
    if clocks["samples"] < 3 and e2e_windows:
        clocks = sampler.summary(windows + e2e_windows)
        clocks["note"] = "timed region shorter than the NVML poll; samples include the e2e loop"

    # ---- parity on EVERY rank at every N, and the CPU baseline (rank 0, N=1 only) ----
    cpu_baseline = None
    cores = min(host_cores(), 64)
    res_w = det.run(x, want_near=True)                       # whole-batch near-threshold flags (eps 1e-3 dB)
    near_total = int(res_w.det.near.sum().item())
    band_h = res_w.band_db.cpu().numpy()
    noise_h = res_w.noise_db.cpu().numpy()
    del res_w
    n_par = min(n_files, (4 * cores) if world == 1 else max(2, min(8, cores // world)))
    if args.no_cpu_baseline:
        n_par = min(n_par, 2)
    files_np = [x[i].cpu().numpy() for i in range(n_par)]
    procs = max(1, min(cores if world == 1 else max(1, cores // world), n_par))
    _, res_p = run_pool(_port_worker, files_np, start_us_host[:n_par], procs, extra=(True,))
    parity = band_parity_report(res_p, band_h, noise_h, near_total, counts_host, events_host, n_par)
    if world > 1:
        agg = torch.tensor([parity["files_checked"], parity["files_with_different_events"], parity["events_checked"],
                            parity["band_values_checked"], parity["band_values_rel_err_gt_1e-4"],
                            parity["frames_within_1e-3_dB_of_threshold_whole_batch"]], dtype=torch.int64, device=dev)
        worst = torch.tensor([parity["band_max_rel_err"]], dtype=torch.float64, device=dev)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        a = agg.cpu().tolist()
        parity.update({"files_checked": a[0], "files_with_different_events": a[1], "events_checked": a[2],
                       "band_values_checked": a[3], "band_values_rel_err_gt_1e-4": a[4],
                       "frames_within_1e-3_dB_of_threshold_whole_batch": a[5], "band_max_rel_err": float(worst.item()),
                       "ranks_checked": world})
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s = min(n_files, 6 * cores)          # ~0.15-0.2 s of CPU per file -> roughly 15-20 s of CPU work
        files_cpu = files_np + [x[i].cpu().numpy() for i in range(n_par, n_s)]
        dt, res_c, kind, what = cpu_arm(files_cpu, start_us_host[:n_s], cores)
        cpu_baseline = {"value": n_s * SAMPLES_PER_FILE / dt / 1e6, "unit": "Msamples/s", "cores": cores,
                        "kind": kind,
                        "sample": f"first {n_s} of the {n_files} files, {cores} worker processes, {what} ({dt:.2f} s)"}
        mism = sum(1 for i in range(n_s)
                   if [tuple(int(v) for v in p) for p in events_host[i, :counts_host[i]]] != res_c[i][0])
        hours_ref = {}
        for r in res_c:
            for h, c in r[1].items():
                k = int((h - T0).total_seconds() // 3600)
                hours_ref[k] = [hours_ref.get(k, [0, 0])[0] + c[0], hours_ref.get(k, [0, 0])[1] + c[1]]
        whole = [k for k in hours_ref if (k + 1) * 12 <= n_s]       # hours whose 12 files were all in the sample
        hour_mism = sum(1 for k in whole if hours_ref[k] != [int(hist_host[k, 0]), int(hist_host[k, 1])])
        parity["vs_" + kind] = {"files_checked": n_s, "files_with_different_events": mism,
                                "events_checked": int(sum(len(r[0]) for r in res_c)),
                                "whole_hours_checked": len(whole), "hours_with_different_counts": hour_mism}

    # ---- extra fields: configs[2] archive (every N), configs[3] sweep and configs[4] streaming (N=1) ----
    archive = sweep = streaming = ingest = detector_c = pipelined = None
    if impl == "tc" and not args.no_extras:
        if world == 1 and pipe is None:
            pipelined = pipelined_field(det, x, start_us, hour0, n_hours, torch)
            assert pipelined["histogram_equals_inline_pass"], "pipelined pass and in-line pass disagree"
        del hist, warm
        archive = archive_field(det, dev, rank, world, torch, dist)
        if world == 1:
            ingest = ingest_field(x, torch, n_files)
            assert ingest["events"] == int(hist_host[:, 0].sum()), "file ingest path and resident path disagree"
            sweep = sweep_field(x, torch, ops, measured_hbm_peak()[0])
            del x
            torch.cuda.empty_cache()
            detector_c = detector_c_field(torch, ops, measured_hbm_peak()[0])
            torch.cuda.empty_cache()
            streaming = streaming_field(torch)

    if rank == 0:
        total_samples = world * n_files * SAMPLES_PER_FILE
        ms_per_step = elapsed_ms / args.steps
        peak, peak_src = measured_hbm_peak()
        achieved = n_files * nb * ALGO_BYTES_PER_BLOCK / (k2_ms * 1e-3) / 1e9
        traffic, traffic_src = traffic_from_profile()
        line = {
            "metric": METRIC, "value": total_samples / (ms_per_step * 1e-3) / 1e6, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i8" if impl == "tc" else "f32",
            "dtype_note": ("u8 x s8 -> s32 tensor-core accumulation (exact), fp64 epilogue, fp32 dB out; thresholds fp64"
                           if impl == "tc" else "fp32 FFT, fp32 dB out; thresholds fp64"),
            "data": "synthetic", "config": workload_config(world, impl, n_files),
            "repetitions": {"reps": reps, "statistic": "median over repetitions of (max over ranks of the K-step region)",
                            "ms_per_step_each_rep": [round(float(v) / args.steps, 6) for v in per_rep],
                            "per_rank_ms_per_step_median_rep": [round(float(v) / args.steps, 6) for v in allt[:, 0, med]],
                            "reduce_ms_each_rep_max_over_ranks": [round(float(v), 4) for v in allt[:, 1, :].max(axis=0)],
                            "host_enqueue_ms_per_step_median_rep_per_rank": [round(float(v) / args.steps, 6) for v in allt[:, 2, med]]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "dft_i8_kernel" if impl == "tc" else "stft_kernel",
                         "kernel_ms": k2_ms, "kernel_ms_samples": len(k2_samples),
                         "kernel_ms_min": min(k2_samples), "kernel_ms_max": max(k2_samples),
                         "algorithmic_bytes_per_launch": n_files * nb * ALGO_BYTES_PER_BLOCK,
                         "step_frac": n_files * nb * (ALGO_BYTES_PER_BLOCK + 12) / (ms_per_step * 1e-3) / 1e9 / peak,
                         "peak_source": peak_src},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": 2 * args.steps * reps, "clocks": clocks,
            "parity_sample": parity, "multi_gpu_check": multi, "dense_layout": dense,
            "hourly_counts": {"anzahl_total": int(hist_host[:, 0].sum()), "kritisch_total": int(hist_host[:, 1].sum()),
                              "hours": int(n_hours)},
            "archive": archive, "ingest": ingest, "sweep": sweep, "streaming": streaming,
            "detector_c": detector_c, "pipelined_pass": pipelined,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
