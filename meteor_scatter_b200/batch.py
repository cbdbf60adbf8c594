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


def reduce_hist(hist: torch.Tensor, group=None, dst: int = 0) -> torch.Tensor:
    """The one collective of the path: sum per-rank hourly histograms onto ``dst``
    (NCCL for CUDA tensors, gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(hist, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return hist


def stage_files(paths, fs_expected: int = 6000, io_threads: int = 8):
    """Read WAVs into one pinned ``[n_files, max_len]`` buffer (zero padded) + per-file lengths.

    Only the RIFF headers are parsed up front; the samples are then read on ``io_threads`` threads straight into the
    pinned rows (``readinto``: one kernel copy from the page cache, releases the GIL).  Files that are not PCM16 are
    converted to float32 through a memory map."""
    infos = []
    for p in paths:
        info = wav_info(p)
        assert info[0] == fs_expected, f"Sample rate must be {fs_expected} Hz, but got {info[0]} Hz"
        assert info[2] == 1, f"Data must be mono or stereo, but got shape ({info[3]}, {info[2]})"
        infos.append(info)
    if not infos:
        return torch.empty((0, 0), dtype=torch.int16), np.zeros(0, dtype=np.int64)
    dt = np.dtype(np.int16) if all(i[1] == np.dtype("<i2") for i in infos) else np.dtype(np.float32)
    lens = np.array([i[3] for i in infos], dtype=np.int64)
    max_len = int(lens.max())
    max_len += (-max_len) % 8                       # keep every file 16-byte aligned for TMA
    host = torch.empty((len(infos), max_len), dtype=torch.int16 if dt == np.int16 else torch.float32)
    if torch.cuda.is_available():
        host = host.pin_memory()
    hv = host.numpy()

    def copy_one(i):
        n = int(lens[i])
        if infos[i][1] == dt:
            read_wav_into(paths[i], infos[i], hv[i])
        else:
            hv[i, :n] = read_wav(paths[i])[1].astype(dt)
        hv[i, n:] = 0

    if io_threads > 1 and len(infos) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(io_threads, len(infos))) as pool:
            list(pool.map(copy_one, range(len(infos))))
    else:
        for i in range(len(infos)):
            copy_one(i)
    return host, lens


def process_files(paths, params: DetectorAParams | None = None, file_starts=None, csv_folder: str | None = None,
                  device=None, impl: str = "auto", max_events: int | None = None, group=None, chunk_files: int = 288,
                  io_threads: int = 8):
    """Run detector A over ``paths`` (this rank's share when torch.distributed is
    initialised), return per-file detections and the merged hourly histogram, and
    optionally write the dashboard day files on rank 0.

    file_starts: naive-UTC datetimes; default = parsed from the file names
    (dsp/src/main.py:859-862, 917-923).
    The rank's files are processed ``chunk_files`` at a time (one day of 5-minute recordings by default): while the
    GPU works on a chunk, the next chunk is read into a second pinned buffer by ``io_threads`` reader threads, so
    host and device memory stay bounded for archives of any length.  All chunks accumulate into one hourly histogram.
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
    # every rank needs the same hour grid: lengths of ALL files from their headers (memory-mapped, no samples read)
    durations = [wav_info(p)[3] / params.fs for p in paths]
    hour0, n_hours = hour_span(file_starts, durations)
    hist = torch.zeros((n_hours, 2), dtype=torch.int32, device=dev)
    results = {}
    overflow = None
    chunk_files = max(1, int(chunk_files))
    chunks = [mine[i:i + chunk_files] for i in range(0, len(mine), chunk_files)]
    with ThreadPoolExecutor(max_workers=1) as prefetch:
        def stage(c):
            return stage_files([paths[i] for i in c], params.fs, io_threads)
        pending = prefetch.submit(stage, chunks[0]) if chunks else None
        for k, c in enumerate(chunks):
            host, lens = pending.result()
            pending = prefetch.submit(stage, chunks[k + 1]) if k + 1 < len(chunks) else None
            x = host.to(dev, non_blocking=True)
            nbpf = torch.from_numpy((lens // det.spec.block_size).astype(np.int32)).to(dev)
            us = torch.tensor([datetime_to_us(file_starts[i]) for i in c], dtype=torch.int64, device=dev)
            part = torch.zeros_like(hist)
            res = det.run(x, n_blocks_per_file=nbpf, hourly=dict(file_start_us=us, hour0=hour_index(hour0),
                                                                 n_hours=n_hours, out=part))
            hist += part
            try:
                for j, i in enumerate(c):      # D2H of the event lists: also keeps `host` alive until the copy is done
                    results[i] = res.detections(j, file_starts[i])
            except RuntimeError as e:          # explicit max_events exceeded: finish the collective first
                overflow = overflow or e
    reduce_hist(hist, group=group)
    if overflow is not None:
        raise overflow
    hist_host = hist.cpu().numpy()
    written = []
    if rank == 0 and csv_folder is not None:
        written = csvout.write_day_files(csv_folder, csvout.hourly_rows(hist_host, hour0))
    return dict(detections=results, hist=hist_host, hour0=hour0, n_hours=n_hours, csv_files=written, rank=rank,
                world=world)
