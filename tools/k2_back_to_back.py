"""How much does the kernel boundary cost?  Times N back-to-back band-power launches (no detect stage) against the
event-timed single launch.  Run on a B200: python tools/k2_back_to_back.py"""
import json
import sys

import torch

sys.path.insert(0, ".")
from meteor_scatter_b200 import ops  # noqa: E402
from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams  # noqa: E402
from meteor_scatter_b200.synth import synth_batch_torch  # noqa: E402

dev = torch.device("cuda", 0)
det = DetectorA(DetectorAParams(), impl="tc", max_events=256)
n_files, spf = 288, 1_800_000
x = synth_batch_torch(n_files, spf, device=dev)
nb = det.spec.n_blocks(spf)
outs = [(torch.empty((n_files, nb), dtype=torch.float32, device=dev), torch.empty((n_files, nb), dtype=torch.float32, device=dev))
        for _ in range(2)]
for _ in range(10):
    ops.band_power(x, det.spec, impl="tc", out=outs[0])
torch.cuda.synchronize()
single = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    ops.band_power(x, det.spec, impl="tc", out=outs[0])
    b.record()
    torch.cuda.synchronize()
    single.append(a.elapsed_time(b))
n = 300
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(n):
    ops.band_power(x, det.spec, impl="tc", out=outs[i & 1])
b.record()
torch.cuda.synchronize()
print(json.dumps({"single_ms_median": sorted(single)[len(single) // 2], "back_to_back_ms": a.elapsed_time(b) / n}))
