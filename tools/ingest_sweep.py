import datetime, json, os, shutil, sys, tempfile, time
import torch
sys.path.insert(0, os.getcwd())
from meteor_scatter_b200.batch import process_files
from meteor_scatter_b200.synth import synth_file
from meteor_scatter_b200.wavio import write_wav_pcm16
root = tempfile.mkdtemp(prefix="ms_ingest_", dir="/dev/shm")
t0 = datetime.datetime(2025, 6, 25)
base = [synth_file(900 + i, dur_s=300.0, rate_per_hour=120.0) for i in range(8)]
paths = []
for i in range(288):
    t = t0 + datetime.timedelta(seconds=300 * i)
    p = os.path.join(root, f"expoFull_gqrx_{t.strftime('%Y%m%d_%H%M%S')}_49969000.wav")
    write_wav_pcm16(p, 6000, base[i % 8]); paths.append(p)
samples = 288 * len(base[0])
for chunk in (12, 24, 48):
    for threads in (16, 32, 64):
        process_files(paths, csv_folder=None, chunk_files=chunk, io_threads=threads)
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); t = time.perf_counter()
            process_files(paths, csv_folder=None, chunk_files=chunk, io_threads=threads)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
        ts.sort()
        print(json.dumps({"chunk_files": chunk, "io_threads": threads, "median_s": round(ts[2], 4), "min_s": round(ts[0], 4), "Gsamples_per_s": round(samples / ts[2] / 1e9, 2)}))
shutil.rmtree(root, ignore_errors=True)
