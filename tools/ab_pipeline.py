#!/usr/bin/env python
"""Same-process A/B of the in-stream pass (K2 then K3 as its programmatic dependent) against the overlapped pass
(PassPipeline: K3 of batch i on the SMs that K2 of batch i+1 leaves free), alternating rounds of 50 steps on the same
24 h batch so that box-to-box and power-state differences cancel.  MS_OVL_SPARE_SMS selects the spare-SM count."""
import datetime
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams, PassPipeline, datetime_to_us  # noqa: E402
from meteor_scatter_b200.synth import synth_batch_torch                                              # noqa: E402

n_files, spf, steps, rounds = 288, 1_800_000, 50, 12
T0 = datetime.datetime(2025, 6, 1)
dev = torch.device("cuda")
x = synth_batch_torch(n_files, spf, seed=1234, device=dev)
us = torch.tensor([datetime_to_us(T0 + datetime.timedelta(seconds=300 * i)) for i in range(n_files)], dtype=torch.int64,
                  device=dev)
det = DetectorA(DetectorAParams(), impl="tc")
hist = torch.zeros((24, 2), dtype=torch.int32, device=dev)
pipe = PassPipeline(det, n_files, spf, 24, dev, depth=2)


def run_inline():
    for _ in range(steps):
        det.run_pass(x, us, T0, 24, hist)


def run_pipe():
    for _ in range(steps):
        pipe.submit(x, us, T0)
    pipe.drain()


for f in (run_inline, run_pipe):
    f()
torch.cuda.synchronize()
res = {"inline": [], "pipelined": []}
for r in range(rounds):
    for name, f in (("inline", run_inline), ("pipelined", run_pipe)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f()
        b.record()
        torch.cuda.synchronize()
        res[name].append(a.elapsed_time(b) / steps)
ref = hist.cpu().numpy()
_, h = pipe.wait(0)
assert np.array_equal(h.cpu().numpy(), ref), "overlapped pass disagrees with the in-stream pass"
out = {k: {"median_ms": float(np.median(v)), "min_ms": float(np.min(v)), "all": [round(t, 5) for t in v]} for k, v in res.items()}
out["spare_sms"] = os.environ.get("MS_OVL_SPARE_SMS", "default")
out["gain"] = 1.0 - out["pipelined"]["median_ms"] / out["inline"]["median_ms"]
print(json.dumps(out))
