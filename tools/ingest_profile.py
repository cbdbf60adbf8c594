#!/usr/bin/env python
"""cProfile of batch.process_files over 288 warm-cache WAV files (diagnostic)."""
import cProfile
import datetime
import os
import pstats
import shutil
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200.batch import process_files          # noqa: E402
from meteor_scatter_b200.synth import synth_file             # noqa: E402
from meteor_scatter_b200.wavio import write_wav_pcm16        # noqa: E402

root = tempfile.mkdtemp(prefix="ms_ingest_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
base = [synth_file(900 + i, dur_s=300.0) for i in range(8)]
t0 = datetime.datetime(2025, 6, 25)
paths = []
for i in range(288):
    t = t0 + datetime.timedelta(seconds=300 * i)
    p = os.path.join(root, "expoFull_gqrx_" + t.strftime("%Y%m%d_%H%M%S") + "_49969000.wav")
    write_wav_pcm16(p, 6000, base[i % 8])
    paths.append(p)
process_files(paths)
process_files(paths)
pr = cProfile.Profile()
pr.enable()
process_files(paths)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
shutil.rmtree(root)
