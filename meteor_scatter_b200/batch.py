"""Batch / multi-GPU front end of detector A: many recordings per launch,
file-level sharding across ranks, and the dashboard's hourly day files.

The reference processes one file per Python process (dsp/src/main.py:809-948)
and only detector C writes ``Timestamp;Anzahl;Kritisch`` (prime_detection.py:
229-245).  Here files are independent units: every
rank runs STFT -> detect -> hourly histogram on its share and ONE collective
(sum-reduce of the ``[n_hours, 2]`` int32 histogram to rank 0) merges the
result; there is no collective on the data path.
"""
from __future__ import annotations

import datetime
import os

import numpy as np
import torch

from . import csvout, ops
from .pipeline import DetectorA, DetectorAParams, datetime_to_us, hour_index
from .wavio import read_wav, read_wav_into, start_time_from_name, wav_info


def bind_host_to_gpu(device_index: int) -> bool:
    """Pin the calling thread to the CPUs NVML reports as nearest to `device_index`, so that the pinned staging
    buffers allocated afterwards land on that GPU's NUMA node (with 8 ranks per box the H2D ingest otherwise crosses
    the socket interconnect).  Host-side placement only; returns False when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(device_index))
        return True
    except Exception:
        return False


def shard_indices(n_items: int, rank: int, world: int):
    """Round-robin file sharding: item i belongs to rank i % world."""
    return list(range(rank, n_items, world))


def hour_span(file_starts, durations_s):
    """(hour0, n_hours) covering every recording."""
    first = min(file_starts)
    last = max(s + datetime.timedelta(seconds=float(d)) for s, d in zip(file_starts, durations_s))
    hour0 = first.replace(minute=0, second=0, microsecond=0)
    n_hours = int((last - hour0).total_seconds() // 3600) + 1
    return hour0, n_hours


def uncovered_hours(file_starts, durations_s, hour0: datetime.datetime, n_hours: int):
    """Hour indices (relative to ``hour0``) that no recording overlaps: the reference writes a row only for the
    hours its detector was running (prime_detection.py:229-245), so the dashboard shows a gap there, not a zero."""
    covered = set()
    for s, d in zip(file_starts, durations_s):
        if d <= 0:
            continue
        a = int((s - hour0).total_seconds() // 3600)
        b = int(((s + datetime.timedelta(seconds=float(d))) - hour0 - datetime.timedelta(microseconds=1)).total_seconds() // 3600)
        covered.update(range(max(a, 0), min(b, n_hours - 1) + 1))
    return [h for h in range(n_hours) if h not in covered]


def reduce_hist(hist: torch.Tensor, group=None, dst: int = 0) -> torch.Tensor:
    """The one collective of the path: sum per-rank hourly histograms onto ``dst``
    (NCCL for CUDA tensors, gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(hist, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return hist


def _pinned_empty(shape, dtype):
    """Page-locked host tensor allocated directly (``torch.empty(..).pin_memory()`` would allocate pageable memory
    and then COPY it into a pinned block: 50 ms per GB on the critical path)."""
    return torch.empty(shape, dtype=dtype, pin_memory=torch.cuda.is_available())


def _check_info(info, fs_expected):
    assert info[0] == fs_expected, f"Sample rate must be {fs_expected} Hz, but got {info[0]} Hz"
    assert info[2] == 1, f"Data must be mono or stereo, but got shape ({info[3]}, {info[2]})"


def _fill_rows(paths, infos, lens, dt, hv, io_threads, pool=None):
    """Read every file's samples straight into its row of ``hv``.  Files already in the row dtype go through the
    native reader (``ms_read_files``: pread on ``io_threads`` native threads, no interpreter lock -- Python reader
    threads lose most of their time re-acquiring the GIL from a busy main thread); the rest are converted through a
    memory map.  With ``pool`` the call is submitted to it and a list of futures is returned."""
    import ctypes as C
    from . import _lib
    native = [i for i in range(len(paths)) if infos[i][1] == dt]
    other = [i for i in range(len(paths)) if infos[i][1] != dt]
    row_bytes = hv.shape[1] * hv.itemsize if len(paths) else 0

    def run():
        if native:
            n = len(native)
            c_paths = (C.c_char_p * n)(*[os.fsencode(paths[i]) for i in native])
            offs = (C.c_int64 * n)(*[int(infos[i][4]) for i in native])
            nbytes = (C.c_int64 * n)(*[int(lens[i]) * hv.itemsize for i in native])
            dst = (C.c_void_p * n)(*[hv[i].ctypes.data for i in native])
            cap = (C.c_int64 * n)(*([row_bytes] * n))
            _lib.check(_lib.load().ms_read_files(c_paths, offs, nbytes, dst, cap, n, max(1, int(io_threads))))
        for i in other:
            k = int(lens[i])
            hv[i, :k] = read_wav(paths[i])[1].astype(dt)
            hv[i, k:] = 0

    if pool is not None:
        return [pool.submit(run)]
    run()
    return []


def stage_files(paths, fs_expected: int = 6000, io_threads: int = 16, out: torch.Tensor | None = None):
    """Read WAVs into one pinned ``[n_files, max_len]`` buffer (zero padded) + per-file lengths.

    Only the RIFF headers are parsed up front; the samples are then read on ``io_threads`` threads straight into the
    pinned rows.  Files that are not PCM16 are converted to float32 through a memory map.  ``out`` = a pinned buffer to
    fill instead of allocating one (steady-state ingest re-uses a ring of them)."""
    infos = [wav_info(p) for p in paths]
    for info in infos:
        _check_info(info, fs_expected)
    if not infos:
        return torch.empty((0, 0), dtype=torch.int16), np.zeros(0, dtype=np.int64)
    dt = np.dtype(np.int16) if all(i[1] == np.dtype("<i2") for i in infos) else np.dtype(np.float32)
    lens = np.array([i[3] for i in infos], dtype=np.int64)
    max_len = int(lens.max())
    max_len += (-max_len) % 8                       # keep every file 16-byte aligned for TMA
    tdt = torch.int16 if dt == np.int16 else torch.float32
    if out is not None and out.dtype == tdt and out.shape[0] >= len(infos) and out.shape[1] == max_len:
        host = out[:len(infos)]
    else:
        host = _pinned_empty((len(infos), max_len), tdt)
    _fill_rows(paths, infos, lens, dt, host.numpy(), io_threads)
    return host, lens


_RINGS = {}


def _ring(chunk_files: int, max_len: int, tdt, dev, depth: int = 3):
    """Persistent staging ring: ``depth`` pinned host slots + device slots of one chunk each, allocated once per shape
    (pinning 1 GB costs ~0.45 s: it must not happen per call)."""
    key = (chunk_files, max_len, tdt, str(dev), depth)
    r = _RINGS.get(key)
    if r is None:
        _RINGS.clear()
        r = dict(host=[_pinned_empty((chunk_files, max_len), tdt) for _ in range(depth)],
                 dev=[torch.empty((chunk_files, max_len), dtype=tdt, device=dev) for _ in range(depth)],
                 copied=[torch.cuda.Event() for _ in range(depth)], used=[torch.cuda.Event() for _ in range(depth)],
                 copy_stream=torch.cuda.Stream(device=dev))
        _RINGS[key] = r
    return r


def process_files(paths, params: DetectorAParams | None = None, file_starts=None, csv_folder: str | None = None,
                  device=None, impl: str = "auto", max_events: int | None = None, group=None, chunk_files: int = 24,
                  io_threads: int = 16):
    """Run detector A over ``paths`` (this rank's share when torch.distributed is
    initialised), return per-file detections and the merged hourly histogram, and
    optionally write the dashboard day files on rank 0.

    file_starts: naive-UTC datetimes; default = parsed from the file names
    (dsp/src/main.py:859-862, 917-923).
    The rank's files are processed ``chunk_files`` at a time through a persistent ring of three pinned host slots and
    three device slots: ``io_threads`` reader threads fill slot k+2 while chunk k+1 crosses PCIe and chunk k runs on
    the GPU, so host and device memory stay bounded for archives of any length and nothing is allocated or pinned per
    call.  All chunks accumulate into one hourly histogram.
    ``max_events`` = event slots per file; None sizes them from the recording length (cannot overflow).  With an
    explicit cap an overflow is raised only AFTER the collective, so the other ranks never hang in the reduce.
    """
    import torch.distributed as dist
    from concurrent.futures import ThreadPoolExecutor
    params = params or DetectorAParams()
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if file_starts is None:
        file_starts = [start_time_from_name(p) for p in paths]
        assert all(t is not None for t in file_starts), "cannot parse a start time from every file name"
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    mine = shard_indices(len(paths), rank, world)
    det = DetectorA(params, impl=impl, max_events=max_events)
    # every rank needs the same hour grid: lengths of ALL files from their headers (no samples read); parsed once
    all_infos = [wav_info(p) for p in paths]
    durations = [info[3] / params.fs for info in all_infos]
    hour0, n_hours = hour_span(file_starts, durations)
    hist = torch.zeros((n_hours, 2), dtype=torch.int32, device=dev)
    results = {}
    overflow = None
    chunk_files = max(1, int(chunk_files))
    chunks = [mine[i:i + chunk_files] for i in range(0, len(mine), chunk_files)]
    infos = {i: all_infos[i] for i in mine}
    for i in mine:
        _check_info(infos[i], params.fs)
    pcm16 = all(infos[i][1] == np.dtype("<i2") for i in mine)
    dt = np.dtype(np.int16) if pcm16 else np.dtype(np.float32)
    tdt = torch.int16 if pcm16 else torch.float32
    max_len = max((infos[i][3] for i in mine), default=0)
    max_len += (-max_len) % 8
    depth = 3
    ring = _ring(min(chunk_files, max(1, len(mine))), max_len, tdt, dev, depth) if mine else None
    main = torch.cuda.current_stream(dev)

    res_ring = {}

    def start_d2h(k, res):
        """Queue the D2H of one chunk's compact results (counts, events, event dB) into pinned buffers right behind its
        kernels.  A later ``.cpu()`` would be ordered behind everything enqueued since -- i.e. behind the NEXT chunk's
        H2D copy -- and serialise the pipeline (measured: 1.6-2.3 ms of waiting per chunk)."""
        slot = k % depth
        d = res.det
        shapes = (tuple(d.counts.shape), tuple(d.events.shape), tuple(d.event_db.shape))
        buf = res_ring.get(slot)
        if buf is None or buf["shapes"] != shapes:
            buf = res_ring[slot] = dict(shapes=shapes, counts=_pinned_empty(shapes[0], torch.int32),
                                        events=_pinned_empty(shapes[1], torch.int32),
                                        event_db=_pinned_empty(shapes[2], torch.float64), done=torch.cuda.Event())
        buf["counts"].copy_(d.counts, non_blocking=True)
        buf["events"].copy_(d.events, non_blocking=True)
        buf["event_db"].copy_(d.event_db, non_blocking=True)
        buf["done"].record(main)
        return buf

    def collect(c, res, buf):
        """Unpack one chunk's event lists on the host (waits only for that chunk's own D2H copies)."""
        nonlocal overflow
        buf["done"].synchronize()
        counts = buf["counts"].numpy()
        cap = buf["events"].shape[1]
        if counts.size and int(counts.max()) > cap:      # explicit max_events exceeded: finish the collective first
            overflow = overflow or RuntimeError(f"event capacity exceeded: a file produced {int(counts.max())} events, "
                                                f"max_events={cap}; re-run with a larger max_events")
            return
        res._host = dict(counts=counts.copy(), events=buf["events"].numpy().copy(), event_db=buf["event_db"].numpy().copy())
        for j, i in enumerate(c):
            results[i] = res.detections(j, file_starts[i])

    # Three-deep software pipeline over chunks: reader threads fill pinned slot k+2 while chunk k+1 crosses PCIe on the
    # copy stream and chunk k runs on the GPU; the event lists of chunk k-1 are unpacked on the host meanwhile.
    with ThreadPoolExecutor(max_workers=2) as readers:       # one helper thread per chunk in flight drives the native readers
        def start_read(k):
            c, slot = chunks[k], k % depth
            ring["used"][slot].synchronize()            # the H2D copy that last read this pinned slot has finished
            lens = np.array([infos[i][3] for i in c], dtype=np.int64)
            futs = _fill_rows([paths[i] for i in c], [infos[i] for i in c], lens, dt,
                              ring["host"][slot].numpy()[:len(c)], io_threads, pool=readers)
            return lens, futs

        reads = {k: start_read(k) for k in range(min(depth - 1, len(chunks)))}
        prev = None
        trace = [] if os.environ.get("MS_INGEST_TRACE") else None      # diagnostic: host timeline of the pipeline
        import time as _time
        for k, c in enumerate(chunks):
            t_a = _time.perf_counter()
            lens, futs = reads.pop(k)
            for f in futs:
                f.result()
            t_b = _time.perf_counter()
            slot = k % depth
            cs = ring["copy_stream"]
            cs.wait_stream(main)                         # the kernels that last read this device slot are enqueued before
            with torch.cuda.stream(cs):
                ring["dev"][slot][:len(c)].copy_(ring["host"][slot][:len(c)], non_blocking=True)
                ring["used"][slot].record(cs)
                ring["copied"][slot].record(cs)
            if k + depth - 1 < len(chunks):
                reads[k + depth - 1] = start_read(k + depth - 1)
            main.wait_event(ring["copied"][slot])
            x = ring["dev"][slot][:len(c)]
            nbpf = torch.from_numpy((lens // det.spec.block_size).astype(np.int32)).to(dev, non_blocking=True)
            us = torch.tensor([datetime_to_us(file_starts[i]) for i in c], dtype=torch.int64).to(dev, non_blocking=True)
            part = torch.zeros_like(hist)
            res = det.run(x, n_blocks_per_file=nbpf, hourly=dict(file_start_us=us, hour0=hour_index(hour0),
                                                                 n_hours=n_hours, out=part))
            hist += part
            buf = start_d2h(k, res)
            t_c = _time.perf_counter()
            if prev is not None:
                collect(*prev)
            prev = (c, res, buf)
            if trace is not None:
                trace.append((k, round((t_b - t_a) * 1e3, 3), round((t_c - t_b) * 1e3, 3),
                              round((_time.perf_counter() - t_c) * 1e3, 3)))
        if prev is not None:
            collect(*prev)
        if trace is not None:
            print("ingest trace (chunk, ms waiting for its read, ms enqueueing H2D + kernels, ms unpacking the previous "
                  "chunk):", trace)
    reduce_hist(hist, group=group)
    if overflow is not None:
        raise overflow
    hist_host = hist.cpu().numpy()
    written = []
    if rank == 0 and csv_folder is not None:
        skip = uncovered_hours(file_starts, durations, hour0, n_hours)
        written = csvout.write_day_files(csv_folder, csvout.hourly_rows(hist_host, hour0, skip_empty_hours=skip))
    return dict(detections=results, hist=hist_host, hour0=hour0, n_hours=n_hours, csv_files=written, rank=rank,
                world=world)
