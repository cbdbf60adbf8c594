"""Hourly ``Timestamp;Anzahl;Kritisch`` day files -- the contract the reference's
Flask dashboard consumes (database.py:61-106 reads ``YYYYMMDD.csv`` with
``sep=';'``; README.md:52-63 shows the format).

The only writer in the reference is meteor_detect_class/prime_detection.py
(:132-146 creates the day file with its header, :229-245 appends one row per
hour via pandas ``to_csv(sep=';', index=False)``, :255-270 rolls the file at
midnight).  This module emits byte-identical text without pandas.
"""
from __future__ import annotations

import datetime
import os

HEADER = "Timestamp;Anzahl;Kritisch"


def day_file_name(day: datetime.date) -> str:
    return day.strftime("%Y%m%d") + ".csv"          # prime_detection.py:134-137; database.py:77-81


def format_row(hour: datetime.datetime, anzahl: int, kritisch: int) -> str:
    return f"{hour.strftime('%Y-%m-%d %H:%M:%S')};{int(anzahl)};{int(kritisch)}"   # prime_detection.py:232-233


def hourly_rows(hist, hour0: datetime.datetime, skip_empty_hours=None):
    """(hour, Anzahl, Kritisch) rows from a ``[n_hours, 2]`` integer histogram
    whose row 0 is ``hour0``.  ``skip_empty_hours`` = iterable of hour indices
    with no audio coverage (no row is written for them, like a detector that
    was not running)."""
    skip = set(skip_empty_hours or ())
    rows = []
    for i in range(len(hist)):
        if i in skip:
            continue
        rows.append((hour0 + datetime.timedelta(hours=i), int(hist[i][0]), int(hist[i][1])))
    return rows


def write_day_files(folder: str, rows, merge: bool = True):
    """Write/extend ``folder/YYYYMMDD.csv``.  Existing rows with the same
    Timestamp are replaced (re-running a batch is idempotent); others are kept
    (prime_detection.py:141-146 keeps an existing day file)."""
    assert os.path.exists(folder), f"Path not found: {folder}"           # prime_detection.py:35
    by_day = {}
    for hour, a, k in rows:
        by_day.setdefault(hour.date(), []).append((hour, a, k))
    written = []
    for day, drows in sorted(by_day.items()):
        path = os.path.join(folder, day_file_name(day))
        existing = {}
        if merge and os.path.exists(path):
            with open(path, "r", newline="") as f:
                lines = f.read().splitlines()
            for ln in lines[1:]:
                if ln.strip():
                    ts = ln.split(";")[0]
                    existing[ts] = ln
        for hour, a, k in drows:
            existing[hour.strftime("%Y-%m-%d %H:%M:%S")] = format_row(hour, a, k)
        with open(path, "w", newline="") as f:
            f.write(HEADER + "\n")
            for ts in sorted(existing):
                f.write(existing[ts] + "\n")
        written.append(path)
    return written
