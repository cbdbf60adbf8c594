"""Timeline of the overlapped pass (PassPipeline) from CUDA events: when does each band-power kernel start/end on the
main stream and when does the detect stage of the same batch finish on the side stream?  Run on a B200."""
import datetime
import json
import sys

import torch

sys.path.insert(0, ".")
from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams, PassPipeline, datetime_to_us  # noqa: E402
from meteor_scatter_b200.synth import synth_batch_torch  # noqa: E402

dev = torch.device("cuda", 0)
n_files, spf, n_steps = 288, 1_800_000, 24
det = DetectorA(DetectorAParams(), impl="tc", max_events=256)
x = synth_batch_torch(n_files, spf, device=dev)
t0 = datetime.datetime(2026, 1, 1)
start_us = torch.tensor([datetime_to_us(t0 + datetime.timedelta(seconds=300 * i)) for i in range(n_files)],
                        dtype=torch.int64, device=dev)
pipe = PassPipeline(det, n_files, spf, 24, dev, depth=n_steps)
for s in pipe.slots:   # timing-enabled events, one slot per step so nothing is overwritten
    s["k2_done"] = torch.cuda.Event(enable_timing=True)
    s["k3_done"] = torch.cuda.Event(enable_timing=True)
    s["k2_done"].record()
    s["k3_done"].record()
warm = PassPipeline(det, n_files, spf, 24, dev, depth=2)
for _ in range(10):
    warm.submit(x, start_us, t0)
warm.drain()
torch.cuda.synchronize()
begins = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps)]
ends = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps)]
for e in begins + ends:      # torch only lets elapsed_time() read events it has seen recorded; the library re-records them
    e.record()
torch.cuda.synchronize()
for i in range(n_steps):
    pipe.submit(x, start_us, t0, ev_begin=begins[i], ev_end=ends[i], isolate=False)
torch.cuda.synchronize()
rows = []
for i in range(n_steps):
    s = pipe.slots[i]
    rows.append({"step": i, "k2_begin_us": round(1e3 * begins[0].elapsed_time(begins[i]), 1),
                 "k2_end_us": round(1e3 * begins[0].elapsed_time(ends[i]), 1),
                 "k3_end_us": round(1e3 * begins[0].elapsed_time(s["k3_done"]), 1)})
for r in rows:
    r["k2_us"] = round(r["k2_end_us"] - r["k2_begin_us"], 1)
    r["k3_after_k2_us"] = round(r["k3_end_us"] - r["k2_end_us"], 1)
    print(json.dumps(r))
print("mean step us", (rows[-1]["k2_begin_us"] - rows[4]["k2_begin_us"]) / (n_steps - 5))
