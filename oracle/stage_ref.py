#!/usr/bin/env python
"""ORACLE tooling (test infrastructure, NOT product code): stage the unmodified reference for the CPU arm.

The reference is pure Python (no build step).  This recipe copies, byte for byte, the few reference source files
of the hot path from ``/root/reference`` into ``oracle/_ref/`` (git-ignored, so nothing of the reference enters the
history; NOT gpurun-ignored, so the staged files travel to the GPU box like our own built ``.so``):

    dsp/src/main.py                         detector A  (proc_wav_file, main.py:207)
    dsp/src/live/backend/processor.py       detector B  (wav_file_process, processor.py:14)
    dsp/src/live/backend/aggregates.py      detector B config / result types

``oracle/ref_harness.py`` imports them from there when ``/root/reference`` does not exist, which lets
``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` time the reference's OWN code on the GPU box's host
cores (``cpu_baseline.kind == "reference"``).  ``__graft_entry__.build()`` runs this whenever ``/root/reference``
is present; on the GPU box the prebuilt copy is used as is.

    python oracle/stage_ref.py            # prints the staged files and their sha256
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

SRC_ROOT = os.environ.get("MS_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("dsp/src/main.py", "dsp/src/live/backend/__init__.py", "dsp/src/live/backend/processor.py",
         "dsp/src/live/backend/aggregates.py")


def staged() -> bool:
    return all(os.path.exists(os.path.join(DST_ROOT, f)) for f in FILES)


def stage(verbose: bool = False) -> dict:
    """Copy the reference files (unchanged) and write a manifest of their hashes.  Returns the manifest."""
    if not os.path.exists(os.path.join(SRC_ROOT, FILES[0])):
        raise FileNotFoundError(f"reference tree not found at {SRC_ROOT}")
    manifest = {"source_root": SRC_ROOT, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest["files"][rel] = hashlib.sha256(f.read()).hexdigest()
        if verbose:
            print(f"{rel}  sha256={manifest['files'][rel][:16]}")
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


if __name__ == "__main__":
    stage(verbose=True)
