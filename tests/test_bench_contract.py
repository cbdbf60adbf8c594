"""CPU: bench.py's reference arm runs without a GPU and prints ONE JSON line carrying the contract keys
(the same keys, plus roofline/clocks/gpu_launches, are produced by the GPU arm on the B200 box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--files", "8"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("Msamples/s through STFT+band-power+detect")
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0 and d["value"] > 0
    # the unmodified reference is staged under oracle/_ref by __graft_entry__.build() (kind "reference");
    # without it the arm falls back to the oracle port
    from oracle import ref_harness
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_harness.reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["hourly_counts"]["anzahl_total"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and not [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode != 0 and "no CPU fallback" in (p.stdout + p.stderr)


def test_reference_worker_equals_oracle_port(tmp_path):
    """The CPU arm's two forms agree: the unmodified proc_wav_file driven through WAV files + its event CSV gives
    the same event index pairs and hourly counts as the oracle port on the same samples."""
    import pytest
    sys.path.insert(0, ROOT)
    import bench
    from oracle import ref_harness
    if not ref_harness.reference_available():
        pytest.skip("reference not staged (run __graft_entry__.build() where /root/reference exists)")
    from meteor_scatter_b200.synth import synth_file
    files = [synth_file(77 + i, fs=6000, dur_s=300.0, rate_per_hour=400.0) for i in range(2)]
    us = [bench.file_start_us(0, 11 + i) for i in range(2)]          # second file starts exactly on an hour boundary
    d, paths = bench.write_wavs(files, prefix="ms_test_wav_")
    try:
        ref = bench._reference_worker((paths, us))
    finally:
        import shutil
        shutil.rmtree(d, ignore_errors=True)
    port = bench._port_worker((files, us, False))
    assert sum(len(r[0]) for r in ref) > 0
    for r, p in zip(ref, port):
        assert r[0] == p[0] and r[1] == p[1]
