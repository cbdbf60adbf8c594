"""WAV ingest for the detection path: RIFF/WAVE header parse and recording
start time from the file name.

Stands where the reference calls ``scipy.io.wavfile.read`` (dsp/src/main.py:249)
and ``soundfile.read`` (dsp/src/live/backend/processor.py:65-71), and parses the
two file-name schemes of dsp/src/main.py:859-862 (gqrx) and 917-923 (Brams MESZ).
Samples are returned as a numpy view of the file (np.memmap) so they can be
staged to pinned memory without an intermediate copy.
"""
from __future__ import annotations

import datetime
import os
import struct

import numpy as np

WAVE_FORMAT_PCM = 1
WAVE_FORMAT_IEEE_FLOAT = 3
WAVE_FORMAT_EXTENSIBLE = 0xFFFE


class WavFormatError(ValueError):
    pass


def wav_info(path: str):
    """Parse the RIFF header only: (sample_rate, numpy dtype, channels, frames, byte offset of the sample data)."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise WavFormatError(f"{path}: not a RIFF/WAVE file")
        fmt = None
        data_off = data_len = None
        pos = 12
        while pos + 8 <= size:
            f.seek(pos)
            cid, clen = struct.unpack("<4sI", f.read(8))
            if cid == b"fmt ":
                raw = f.read(min(clen, 40))
                tag, ch, rate, _, align, bits = struct.unpack("<HHIIHH", raw[:16])
                if tag == WAVE_FORMAT_EXTENSIBLE and len(raw) >= 26:
                    tag = struct.unpack("<H", raw[24:26])[0]
                fmt = (tag, ch, rate, align, bits)
            elif cid == b"data":
                data_off, data_len = pos + 8, min(clen, size - pos - 8)
                break
            pos += 8 + clen + (clen & 1)
    if fmt is None or data_off is None:
        raise WavFormatError(f"{path}: missing fmt or data chunk")
    tag, ch, rate, align, bits = fmt
    if tag == WAVE_FORMAT_PCM and bits == 16:
        dt = np.dtype("<i2")
    elif tag == WAVE_FORMAT_PCM and bits == 32:
        dt = np.dtype("<i4")
    elif tag == WAVE_FORMAT_PCM and bits == 8:
        dt = np.dtype("u1")
    elif tag == WAVE_FORMAT_IEEE_FLOAT and bits == 32:
        dt = np.dtype("<f4")
    else:
        raise WavFormatError(f"{path}: unsupported WAV encoding (format tag {tag}, {bits} bit)")
    return rate, dt, ch, data_len // (dt.itemsize * ch), data_off


def read_wav(path: str, mmap: bool = True):
    """Return (sample_rate, data) like scipy.io.wavfile.read: int16 / int32 /
    float32 / uint8 ndarray, shape [n] (mono) or [n, channels]."""
    rate, dt, ch, n, data_off = wav_info(path)
    if mmap and n > 0:
        data = np.memmap(path, dtype=dt, mode="r", offset=data_off, shape=(n * ch,))
    else:
        with open(path, "rb") as f:
            f.seek(data_off)
            data = np.frombuffer(f.read(n * ch * dt.itemsize), dtype=dt)
    if ch > 1:
        data = data.reshape(n, ch)
    return rate, data


def read_wav_into(path: str, info, out: np.ndarray):
    """Read the samples of a mono file straight into ``out`` (same dtype, at least ``frames`` long) with readinto():
    one kernel copy from the page cache into the caller's (pinned) buffer, no page faults on a mapping."""
    rate, dt, ch, n, data_off = info
    assert ch == 1 and out.dtype == dt and out.flags.c_contiguous and len(out) >= n
    with open(path, "rb", buffering=0) as f:
        f.seek(data_off)
        view = memoryview(out[:n]).cast("B")
        got = 0
        while got < len(view):
            k = f.readinto(view[got:])
            if not k:
                raise WavFormatError(f"{path}: file shorter than its data chunk")
            got += k


def write_wav_pcm16(path: str, rate: int, data: np.ndarray):
    """Minimal PCM16 / float32 writer (tests and synthetic archives)."""
    data = np.ascontiguousarray(data)
    if data.dtype == np.int16:
        tag, bits = WAVE_FORMAT_PCM, 16
    elif data.dtype == np.float32:
        tag, bits = WAVE_FORMAT_IEEE_FLOAT, 32
    else:
        raise ValueError("write_wav_pcm16 supports int16 and float32")
    ch = 1 if data.ndim == 1 else data.shape[1]
    payload = data.tobytes()
    align = ch * bits // 8
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(payload)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, tag, ch, rate, rate * align, align, bits))
        f.write(b"data" + struct.pack("<I", len(payload)))
        f.write(payload)


def start_time_from_name(path: str) -> datetime.datetime | None:
    """Recording start (naive UTC) from the reference's two naming schemes:
    ``*_gqrx_YYYYMMDD_HHMMSS_<freq>.wav`` (dsp/src/main.py:859-862) and
    ``*_Brams_YYMMDD_HHMESZ.wav`` (local summer time, UTC+2; main.py:917-923)."""
    parts = os.path.basename(path).split("_")
    try:
        if len(parts) == 5:
            return datetime.datetime.strptime(parts[2] + "-" + parts[3], "%Y%m%d-%H%M%S")
        if len(parts) == 4 and parts[3].endswith("MESZ.wav"):
            t = datetime.datetime.strptime(parts[2] + "-" + parts[3].replace("MESZ.wav", ""), "%y%m%d-%H")
            return t - datetime.timedelta(hours=2)
    except ValueError:
        return None
    return None
