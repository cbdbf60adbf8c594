"""GPU parity tests proper: the CUDA path (through the C-ABI) against the oracle
and against the golden fixtures produced by the unmodified reference.

Tolerances (BASELINE.json north_star): band power within 1e-4 relative (fp32
FFT / exact-integer tensor-core DFT vs the reference's fp64 FFT), i.e.
|dB error| <= 10*log10(1+1e-4) = 4.35e-4 dB; event indices and hourly counts
bit-exact except frames within 1e-3 dB of the threshold (reported separately).
"""
import csv
import datetime
import io
import os

import numpy as np
import pytest
import torch

from oracle import detector_a as oa
from oracle import detector_b as ob
from oracle import detector_c as oc
from tests.golden_cases import A_CASES, B_CASES, MB, START, a_sliced, b_input

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4
DB_TOL = 10 * np.log10(1 + REL_TOL)


def _spec(params, fs=6000):
    from meteor_scatter_b200 import ops
    return ops.BandSpec.from_reference_args(fs, params["block_duration_sec"], params["freq_band"],
                                            params["noise_band"], params["n_fft"])


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


_PARITY_COUNTS = {}


def assert_rel_counted(got, ref, what, rtol=REL_TOL, max_outside=0, floor=None):
    """The 1e-4 relative budget, with every exception COUNTED and reported (north_star: "reported separately"), never
    absorbed by a silent absolute floor: at most ``max_outside`` values may miss ``rtol``; those must still be within
    ``floor`` (an array/scalar of absolute bounds stating WHY they miss, e.g. the fp32 rounding floor of a frame).
    Counts go to gpurun_out/r02_parity_counts.json."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    err = np.abs(got - ref)
    bad = err > rtol * np.abs(ref)
    n_bad = int(bad.sum())
    _PARITY_COUNTS[what] = {"values": int(ref.size), "outside_rel_tol": n_bad, "allowed": int(max_outside),
                            "max_rel_err": float(np.max(err / np.maximum(np.abs(ref), 1e-300))) if ref.size else 0.0}
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        import json
        with open("gpurun_out/r02_parity_counts.json", "w") as f:
            json.dump(_PARITY_COUNTS, f, indent=1, sort_keys=True)
    except OSError:
        pass
    assert n_bad <= max_outside, f"{what}: {n_bad} of {ref.size} values outside {rtol} relative (allowed {max_outside})"
    if n_bad and floor is not None:
        assert np.all(err[bad] <= np.broadcast_to(floor, err.shape)[bad]), f"{what}: an exception exceeds its stated floor"


@pytest.mark.parametrize("impl", ["fft", "tc"])
@pytest.mark.parametrize("name", sorted(A_CASES))
def test_band_power_matches_reference(name, impl):
    from meteor_scatter_b200 import ops
    seed, dur, params, adaptive, sl, skw = A_CASES[name]
    x, g = a_sliced(name)
    spec = _spec(params)
    xd = _dev(x).reshape(1, -1)        # (the float32 case holds PCM16 / 32768 values: it runs on the integer path too)
    band_db, noise_db, be, ne = ops.band_power(xd, spec, impl=impl, want_energy=True)
    torch.cuda.synchronize()
    eb_ref, en_ref = oa.stft_band_energy_vec(x, 6000, params["block_duration_sec"], params["freq_band"],
                                             params["noise_band"], params["n_fft"])
    np.testing.assert_allclose(be.cpu().numpy()[0], eb_ref, rtol=REL_TOL)
    np.testing.assert_allclose(ne.cpu().numpy()[0], en_ref, rtol=REL_TOL)
    np.testing.assert_allclose(band_db.cpu().numpy()[0], g["band_power"], rtol=0, atol=DB_TOL + 1e-5)
    np.testing.assert_allclose(noise_db.cpu().numpy()[0], g["noise_power"], rtol=0, atol=DB_TOL + 1e-5)


def test_tc_is_exact_integer_dft():
    """K2 computes the quantised-basis DFT in exact integer arithmetic: compare
    with the same integers evaluated by numpy (int64), to float32 rounding."""
    from meteor_scatter_b200 import ops
    x, g = a_sliced("a_mb_s1")
    spec = _spec(MB)
    xd = _dev(x).reshape(1, -1)
    plan = ops.DftI8Plan.get(spec, xd.device)
    _, _, be, ne = ops.band_power(xd, spec, impl="tc", want_energy=True)
    nb = spec.n_blocks(len(x))
    blocks = x[:nb * spec.block_size].reshape(nb, spec.block_size)[:, :spec.win_len].astype(np.int64)
    scale = 0.99 * float(1 << 23) / np.abs(plan.basis).max()        # ms_dft_i8_plan_build: peak -> 0.99 * 2^23
    v = np.rint(plan.basis * scale).astype(np.int64)                # [K, n_cols]
    assert np.abs(v).max() < (1 << 23)
    X = (blocks @ v).astype(np.float64) * (1.0 / scale)             # exact in int64, one rounding for the scaling
    e = X * X
    eb = e[:, plan.col_group == 0].sum(axis=1)
    en = e[:, plan.col_group == 1].sum(axis=1)
    assert np.array_equal(be.cpu().numpy()[0], eb.astype(np.float32))
    assert np.array_equal(ne.cpu().numpy()[0], en.astype(np.float32))


def test_band_power_edge_cases():
    from meteor_scatter_b200 import ops
    spec = _spec(MB)
    # digital silence -> 10*log10(1e-12) = -120 dB in both bands, full-scale square wave stays finite
    z = torch.zeros((2, 1200 * 130), dtype=torch.int16, device="cuda")
    z[1] = 32767
    z[1, ::2] = -32768
    for impl in ("fft", "tc"):
        b, n = ops.band_power(z, spec, impl=impl)
        assert torch.all(b[0] == -120.0) and torch.all(n[0] == -120.0)
        assert torch.isfinite(b[1]).all() and torch.isfinite(n[1]).all()
    # fewer samples than one block -> zero blocks
    b, n = ops.band_power(torch.zeros((3, 1199), dtype=torch.int16, device="cuda"), spec)
    assert b.shape == (3, 0)
    # batch of files == the files one by one (tile boundaries cross file boundaries in the flat layout)
    rng = np.random.default_rng(5)
    xs = torch.from_numpy(rng.integers(-3000, 3000, size=(5, 1200 * 37), dtype=np.int16)).cuda()
    for impl in ("fft", "tc"):
        b_all, _ = ops.band_power(xs, spec, impl=impl)
        for f in range(5):
            b_one, _ = ops.band_power(xs[f:f + 1].clone(), spec, impl=impl)
            assert torch.equal(b_all[f], b_one[0])
    # ragged per-file tail (samples_per_file not a multiple of the block) on both paths
    xr = torch.from_numpy(rng.integers(-3000, 3000, size=(3, 1200 * 9 + 408), dtype=np.int16)).cuda()
    b_f, n_f = ops.band_power(xr, spec, impl="fft")
    b_t, n_t = ops.band_power(xr, spec, impl="tc")
    np.testing.assert_allclose(b_f.cpu().numpy(), b_t.cpu().numpy(), atol=2 * DB_TOL)


def _pairs_from_result(res, f=0):
    return res.pairs(f)


@pytest.mark.parametrize("name", sorted(A_CASES))
def test_detect_kernel_equals_oracle_on_same_delta(name):
    """K3 alone: feed the golden delta (rounded to the float32 the kernel
    consumes) to both the kernel and the oracle -> identical event indices,
    thresholds to 1e-9."""
    from meteor_scatter_b200 import ops
    seed, dur, params, adaptive, sl, skw = A_CASES[name]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    band32 = g["band_power"].astype(np.float32)
    noise32 = g["noise_power"].astype(np.float32)
    delta = band32.astype(np.float64) - noise32.astype(np.float64)
    k = params["threshold_std_factor"]
    if adaptive:
        dets, thr_ref, pairs_ref = oa.get_detections_adaptive(delta, k, 0.2)
    else:
        dets, thr_ref, pairs_ref = oa.get_detections(delta, k, 0.2)
    res = ops.detect(_dev(band32).reshape(1, -1), _dev(noise32).reshape(1, -1), k, adaptive=adaptive,
                     want_thresholds=True, want_near=True)
    n = int(res.counts[0].item())
    pairs = [tuple(int(v) for v in p) for p in res.events[0, :n].cpu().numpy()]
    assert pairs == pairs_ref
    thr = res.thresholds[0].cpu().numpy()
    if adaptive:
        np.testing.assert_allclose(thr, np.asarray(thr_ref), rtol=0, atol=1e-9)
    else:
        assert abs(thr[0] - thr_ref) < 1e-9
    np.testing.assert_allclose(res.event_db[0, :n].cpu().numpy(), [d.dB for d in dets], rtol=0, atol=1e-9)


@pytest.mark.parametrize("seed", range(6))
def test_detect_kernel_adversarial_parameters(seed):
    from meteor_scatter_b200 import ops
    rng = np.random.default_rng(100 + seed)
    n_files, N = 7, int(rng.integers(40, 900))
    delta = (rng.standard_normal((n_files, N)) * 3.0).astype(np.float32)
    for f in range(n_files):
        for _ in range(int(rng.integers(0, 10))):
            a = int(rng.integers(0, N))
            delta[f, a:a + int(rng.integers(1, 40))] += rng.uniform(3, 25)
    W = int(rng.choice([3, 17, 50, 600]))
    before = int(rng.choice([0, 2, 15]))
    after = int(rng.choice([0, 1, 5, 31, 32, 33, 100]))
    fixed = int(rng.choice([0, 1, 7, 50, 64]))
    k = float(rng.choice([1.0, 2.0, 4.0]))
    lens = rng.integers(1, N + 1, size=n_files).astype(np.int32)
    lens[0] = N
    zeros = torch.zeros((n_files, N), dtype=torch.float32, device="cuda")
    res = ops.detect(_dev(delta), zeros, k, adaptive=True, window_blocks=W, before_blocks=before,
                     after_blocks=after, fixed_blocks=fixed, n_blocks_per_file=torch.from_numpy(lens).cuda(),
                     want_thresholds=True)
    res_g = ops.detect(_dev(delta), zeros, k, adaptive=False, n_blocks_per_file=torch.from_numpy(lens).cuda())
    ev, cnt = res.events.cpu().numpy(), res.counts.cpu().numpy()
    evg, cntg = res_g.events.cpu().numpy(), res_g.counts.cpu().numpy()
    bd = 0.2
    import warnings
    for f in range(n_files):
        d = delta[f, :lens[f]].astype(np.float64)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _, thr_ref, pairs_ref = oa.get_detections_adaptive(d, k, bd, None, (W + .5) * bd, (before + .5) * bd,
                                                               (after + .5) * bd, (fixed + .5) * bd)
        assert [tuple(p) for p in ev[f, :cnt[f]]] == pairs_ref
        np.testing.assert_allclose(res.thresholds[f, :lens[f]].cpu().numpy(), np.asarray(thr_ref, dtype=np.float64),
                                   rtol=0, atol=2e-5, equal_nan=True)
        # global detector incl. its open-at-EOF quirk; zero-length events make the reference assert
        above = d > (d.mean() + k * d.std())
        dd = np.diff(above.astype(int))
        starts = list(np.where(dd == 1)[0] + 1)
        stops = list(np.where(dd == -1)[0] + 1)
        if above[0]:
            starts.insert(0, 0)
        if above[-1]:
            stops.append(len(d) - 1)
        assert [tuple(p) for p in evg[f, :cntg[f]]] == list(zip(starts, stops))


@pytest.mark.parametrize("impl", ["fft", "tc"])
@pytest.mark.parametrize("name", sorted(A_CASES))
def test_end_to_end_events_match_reference(name, impl):
    """Audio -> events on the GPU vs the unmodified reference's events."""
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams
    seed, dur, params, adaptive, sl, skw = A_CASES[name]
    x, g = a_sliced(name)
    p = DetectorAParams(flag_adaptive_threshold=adaptive, **params)
    det = DetectorA(p, impl=impl)
    xd = _dev(x).reshape(1, -1)
    res = det.run(xd, want_thresholds=True, want_near=True, eps_db=1e-3)
    ref_pairs = [(int(round(a / 0.2)), int(round(b / 0.2))) for a, b in zip(g["t_start"], g["t_stop"])]
    n_near = int(res.det.near.sum().item())
    if n_near == 0:
        assert res.pairs(0) == ref_pairs
        dets = res.detections(0, START)
        assert [d.t_start for d in dets] == list(g["t_start"])
        assert [d.t_stop for d in dets] == list(g["t_stop"])
        assert [d.dur_s for d in dets] == list(g["dur"])
        assert [d.utc_start.isoformat() for d in dets] == list(g["utc_start"])
        np.testing.assert_allclose([d.dB for d in dets], g["dB"], rtol=0, atol=1e-3)
        if adaptive:
            np.testing.assert_allclose(res.det.thresholds[0].cpu().numpy(), g["thresholds"], rtol=0, atol=1e-3)
    else:  # frames within eps of the threshold are reported separately (north_star)
        print(f"{name}/{impl}: {n_near} frame(s) within 1e-3 dB of the threshold; exactness not asserted")
    delta = (res.band_db[0].double() - res.noise_db[0].double()).cpu().numpy()
    np.testing.assert_allclose(delta, g["delta_power"], rtol=0, atol=2 * DB_TOL + 2e-5)


def test_hourly_counts_match_oracle():
    from meteor_scatter_b200 import ops
    rng = np.random.default_rng(3)
    n_files, cap = 40, 16
    counts = rng.integers(0, cap + 1, size=n_files).astype(np.int32)
    events = np.zeros((n_files, cap, 2), dtype=np.int32)
    starts = []
    t0 = datetime.datetime(2025, 6, 24, 22, 13, 7)
    dets = []
    for f in range(n_files):
        fs = t0 + datetime.timedelta(seconds=300 * f + int(rng.integers(0, 3)))
        starts.append(fs)
        s = np.sort(rng.choice(1500, size=counts[f], replace=False))
        for e in range(counts[f]):
            ln = int(rng.integers(0, 6))
            events[f, e] = (s[e], s[e] + ln)
            ts, te = s[e] * 0.2, (s[e] + ln) * 0.2
            dets.append(oa.OutputDetection(ts, te, te - ts, 0.0, fs + datetime.timedelta(seconds=ts), None))
    ref = oa.hourly_counts(dets)
    hour0 = t0.replace(minute=0, second=0, microsecond=0)
    from meteor_scatter_b200.pipeline import datetime_to_us, hour_index
    us = torch.tensor([datetime_to_us(t) for t in starts], dtype=torch.int64, device="cuda")
    hist = ops.hourly_counts(_dev(events), _dev(counts), us, 0.2, hour_index(hour0), 8).cpu().numpy()
    for i in range(8):
        h = hour0 + datetime.timedelta(hours=i)
        assert list(hist[i]) == ref.get(h, [0, 0])
    assert hist[:, 0].sum() == counts.sum()


def test_proc_wav_file_drop_in(tmp_path):
    from meteor_scatter_b200.dsp.src.main import proc_wav_file
    from meteor_scatter_b200.wavio import write_wav_pcm16
    for name in ("a_mb_s1", "a_mb_s7_slice", "a_mb_s8_f32", "a_mb_s5_global"):
        seed, dur, params, adaptive, sl, skw = A_CASES[name]
        from tests.golden_cases import a_input
        x, g = a_input(name)
        wav = tmp_path / f"{name}.wav"
        write_wav_pcm16(str(wav), 6000, x)
        lbl, csvp = tmp_path / f"{name}.txt", tmp_path / f"{name}.csv"
        kw = dict(wav_start_sec=sl[0], wav_end_sec=sl[1]) if sl else {}
        out = proc_wav_file(str(wav), wav_start_date_time=START, out_audacity_lbl_file=str(lbl),
                            out_csv_file=str(csvp), disable_show_and_write=True, flag_adaptive_threshold=adaptive,
                            quiet=True, **params, **kw)
        assert open(lbl).read() == str(g["label_text"])
        rows = list(csv.DictReader(io.StringIO(open(csvp, newline="").read())))
        ref = list(csv.DictReader(io.StringIO(str(g["csv_text"]))))
        assert len(rows) == len(ref)
        for a, b in zip(rows, ref):
            for col in ("t_start", "t_stop", "dur_s", "utc_start", "utc_stop"):
                assert a[col] == b[col]
            assert abs(float(a["dB"]) - float(b["dB"])) < 1e-3
    with pytest.raises(AssertionError):
        proc_wav_file(str(tmp_path / "missing.wav"), 0.2, (993, 1013), (690, 710), 512, 4)
    bad = tmp_path / "fs4000.wav"
    write_wav_pcm16(str(bad), 4000, np.zeros(8000, dtype=np.int16))
    with pytest.raises(AssertionError, match="Sample rate must be 6000"):
        proc_wav_file(str(bad), 0.2, (993, 1013), (690, 710), 512, 4, quiet=True)


# ----------------------------------------------------------------------------- detector B
@pytest.mark.parametrize("name", sorted(B_CASES))
def test_welch_band_db_matches_reference(name):
    from meteor_scatter_b200 import ops
    seed, dur, cfgkw, skw = B_CASES[name]
    x, g = b_input(name)
    cfg = ob.ConfigDetection(**cfgkw)
    freqs = np.fft.rfftfreq(cfg.n_fft, 1 / 4000)
    bands = []
    for lo, hi in ob.band_edges(cfg):
        k = np.nonzero((freqs >= lo) & (freqs <= hi))[0]
        bands.append((int(k[0]), int(k[-1])))
    for xin in (_dev(x), _dev(x.astype(np.float32) / 32768.0)):
        out = ops.welch_band_db(xin, 800, cfg.n_fft, bands, 4000.0).cpu().numpy()[0]
        assert out.shape[0] == len(g["db2"])
        np.testing.assert_allclose(out[:, 0], g["ms_db"], rtol=0, atol=DB_TOL + 1e-5)
        np.testing.assert_allclose(out[:, 1], g["n1_db"], rtol=0, atol=DB_TOL + 1e-5)
        np.testing.assert_allclose(out[:, 2], g["n2_db"], rtol=0, atol=DB_TOL + 1e-5)
        np.testing.assert_allclose(out[:, 3], g["db2"], rtol=0, atol=2 * DB_TOL + 2e-5)


@pytest.mark.parametrize("name", sorted(B_CASES))
@pytest.mark.parametrize("chunk", [0, 5])
def test_live_state_machine_equals_oracle(name, chunk):
    """B-state alone on the same (float32-rounded) db2 series: thresholds bit-exact
    (numpy pairwise summation reproduced), detections exact in time, 1e-9 in dB stats;
    identical when streamed in 5-block (1 s) chunks."""
    from meteor_scatter_b200 import _lib, ops
    seed, dur, cfgkw, skw = B_CASES[name]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    cfg = ob.ConfigDetection(**cfgkw)
    db2 = g["db2"].astype(np.float32)
    dets_ref, thr_ref = ob.live_state_machine(db2.astype(np.float64), cfg, 4000, 800)
    lc = _lib.LiveConfig(block_samples=800, fs=4000.0, k_std=cfg.threshold_std_factor,
                         init_wait_sec=cfg.init_detection_wait_sec, after_wait_sec=cfg.after_tracking_wait_sec,
                         mean_min_db=cfg.detection_db_over_noise_mean_min, dur_min_sec=cfg.detection_dur_min_sec,
                         avg_win=int(cfg.avg_win_sec / cfg.proc_block_sec))
    st = ops.LiveStates(1, "cuda")
    d = _dev(db2).reshape(1, -1)
    if chunk == 0:
        thr = ops.live_state_step(st, lc, d, want_thresholds=True).cpu().numpy()[0]
    else:
        parts = [ops.live_state_step(st, lc, d[:, i:i + chunk].contiguous(), want_thresholds=True)
                 for i in range(0, d.shape[1], chunk)]
        thr = torch.cat(parts, dim=1).cpu().numpy()[0]
    assert np.array_equal(thr, np.asarray(thr_ref, dtype=np.float64), equal_nan=True)
    n = int(st.det_count[0].item())
    assert n == len(dets_ref)
    got = st.det[0, :n].cpu().numpy()
    ref = np.array([[m.time_start, m.time_stop, m.duration, m.db_min, m.db_max, m.db_mean, m.db_std]
                    for m in dets_ref]).reshape(-1, 7)
    assert np.array_equal(got[:, :5], ref[:, :5])
    np.testing.assert_allclose(got[:, 5:], ref[:, 5:], rtol=0, atol=1e-9)


def test_band_power_tc_in_place_form_still_exact():
    """MS_K2_TS=0 selects the band-power kernel's earlier form (hi bytes rewritten in place in shared memory, A operand
    from shared memory).  The knob is read once per process, so the check runs in a child process: both forms must give
    the same bits (the integer accumulation is exact in either)."""
    import subprocess
    import sys
    code = (
        "import numpy as np, torch, sys\n"
        "from meteor_scatter_b200 import ops\n"
        "from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams\n"
        "from meteor_scatter_b200.synth import synth_file\n"
        "x = torch.from_numpy(np.stack([synth_file(s, fs=6000, dur_s=60.0, rate_per_hour=900.0) for s in (5, 6, 7)])).cuda()\n"
        "r = DetectorA(DetectorAParams(), impl='tc').run(x)\n"
        "torch.cuda.synchronize()\n"
        "np.save(sys.argv[1], np.stack([r.band_db.cpu().numpy(), r.noise_db.cpu().numpy()]))\n")
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for ts in ("1", "0"):
            path = os.path.join(d, f"out{ts}.npy")
            env = dict(os.environ, MS_K2_TS=ts)
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), timeout=300)
            outs.append(np.load(path))
    assert np.array_equal(outs[0], outs[1])
    assert np.isfinite(outs[0]).all() and outs[0].shape == (2, 3, 300)


# ----------------------------------------------------------------------------- detector C
def test_psd_spectrogram_matches_oracle():
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.synth import synth_file
    x = synth_file(21, fs=5000, dur_s=30.0, carrier_hz=1000.0, rate_per_hour=1200.0)
    ref = oc.plot_spectrogram_numeric(x, 5000.0)
    rows = ref["rows"]
    freqs = ref["freqs"]
    nk = np.nonzero((freqs >= 250) & (freqs <= 800))[0]
    psd, noise = ops.psd_spectrogram(_dev(x), 5000.0, 2048, 1024, np.hanning(2048), int(rows[0]), int(rows[-1]),
                                     int(nk[0]), int(nk[-1]))
    assert psd.shape == (1, 164, 145)
    pr = ref["pxx"][rows]
    # K1 is an fp32 FFT: its rounding error scales with the strongest component of a frame, so a bin ~50 dB below a
    # strong ping cannot be held to 1e-4 of itself.  Such bins are counted (measured: 0-1 of the 23 780 bins) and
    # must stay within 1e-8 of their frame's peak power.
    assert_rel_counted(psd.cpu().numpy()[0], pr, "psd_spectrogram_fft_vs_oracle", max_outside=3,
                       floor=1e-8 * pr.max(axis=0, keepdims=True))
    bandwidth = len(nk) * 5000.0 / 2048
    dens = 10 * np.log10(noise.cpu().numpy()[0] / bandwidth)
    assert abs(dens - ref["density_db_hz"]) < DB_TOL


@pytest.mark.parametrize("case", ["low_rows_i16", "high_rows_i16", "f32", "hop256", "hop2048", "no_noise_band",
                                  "dc_row", "nyquist_fallback", "odd_offset_fallback", "batch"])
def test_psd_warp_kernel_branches(case):
    """nfft 2048 runs on the warp-per-frame kernel (csrc/ms_fft_warp.cuh): both epilogue widths (bins below / above
    512), float32 input, other hops, an empty noise band, the DC bin; a band that reaches the Nyquist bin and data
    that is not pair aligned fall back to the block-cooperative kernel.  Reference: scipy.signal.spectrogram in fp64
    (the pinned stand-in for mlab.specgram, prime_detection.py:70-71)."""
    from scipy.signal import spectrogram
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.synth import synth_file
    fs, nfft = 5000.0, 2048
    x = synth_file(77, fs=5000, dur_s=12.0, carrier_hz=1000.0, rate_per_hour=2400.0)
    k_lo, k_hi, n_lo, n_hi, noverlap, n_seg = 328, 491, 103, 327, 1024, 1
    if case == "high_rows_i16":
        k_lo, k_hi, n_lo, n_hi = 480, 700, 900, 1023
    elif case == "f32":
        x = (x.astype(np.float32) * np.float32(0.37)).astype(np.float32)       # not PCM16 values
    elif case == "hop256":
        noverlap = nfft - 256
    elif case == "hop2048":
        noverlap = 0
    elif case == "no_noise_band":
        n_lo, n_hi = 5, 4
    elif case == "dc_row":
        k_lo, k_hi, n_lo, n_hi = 0, 40, 0, 3
    elif case == "nyquist_fallback":
        k_lo, k_hi = 1000, 1024
    elif case == "odd_offset_fallback":
        x = x[1:]
    elif case == "batch":
        n_seg = 7
    w = np.hanning(nfft)
    if case == "odd_offset_fallback":
        xd = _dev(np.concatenate([[0], x]).astype(x.dtype))[1:].reshape(1, -1)   # a view that starts on an odd sample
        xs = [x]
    elif n_seg > 1:
        n = 20000
        xs = [x[i * 3000:i * 3000 + n] for i in range(n_seg)]
        xd = _dev(np.stack(xs))
    else:
        xd, xs = _dev(x).reshape(1, -1), [x]
    psd, noise = ops.psd_spectrogram(xd, fs, nfft, noverlap, w, k_lo, k_hi, n_lo, n_hi)
    got = psd.cpu().numpy()
    for i, xi in enumerate(xs):
        _, _, ref = spectrogram(xi.astype(np.float64), fs, window=w, nperseg=nfft, noverlap=noverlap, detrend=False,
                                scaling="density", mode="psd")
        assert got.shape[1:] == (k_hi - k_lo + 1, ref.shape[1])
        pr = ref[k_lo:k_hi + 1]
        assert_rel_counted(got[i], pr, f"psd_warp_{case}_{i}", max_outside=3,
                           floor=1e-8 * ref.max(axis=0, keepdims=True))
        want = ref[n_lo:n_hi + 1].sum() if n_lo <= n_hi else 0.0
        assert abs(noise[i].item() - want) <= 2e-5 * max(want, 1e-300)


def test_one_call_pass_equals_separate_calls():
    """ms_detector_a_pass_i16 (one FFI call, hourly fused into detect) == band_power + detect + hourly_counts."""
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams, datetime_to_us, hour_index
    from meteor_scatter_b200.synth import synth_file
    xs = np.stack([synth_file(50 + i, dur_s=300.0, rate_per_hour=200.0) for i in range(14)])
    x = _dev(xs)
    t0 = datetime.datetime(2025, 6, 1, 22, 35, 0)
    starts = [t0 + datetime.timedelta(seconds=300 * i) for i in range(14)]
    us = torch.tensor([datetime_to_us(t) for t in starts], dtype=torch.int64, device="cuda")
    hour0 = t0.replace(minute=0, second=0)
    det = DetectorA(DetectorAParams(), impl="tc")
    hist = torch.full((3, 2), 77, dtype=torch.int32, device="cuda")
    r1 = det.run_pass(x, us, hour0, 3, hist)
    pairs1 = [r1.pairs(f) for f in range(14)]
    hist1 = hist.cpu().numpy().copy()
    r2 = DetectorA(DetectorAParams(), impl="tc").run(x)
    assert [r2.pairs(f) for f in range(14)] == pairs1
    hist2 = ops.hourly_counts(r2.det.events, r2.det.counts, us, 0.2, hour_index(hour0), 3).cpu().numpy()
    assert np.array_equal(hist1, hist2)
    assert hist1[:, 0].sum() == sum(len(p) for p in pairs1) > 0
    # and against the oracle, file by file
    ref = {}
    for f in range(14):
        r = oa.detect_wav(xs[f], 6000, 0.2, (993, 1013), (690, 710), 512, 4, wav_start_date_time=starts[f])
        assert r["pairs"] == pairs1[f]
        for h, c in oa.hourly_counts(r["detections"]).items():
            a = ref.setdefault(h, [0, 0])
            a[0] += c[0]
            a[1] += c[1]
    for i in range(3):
        assert list(hist1[i]) == ref.get(hour0 + datetime.timedelta(hours=i), [0, 0])
    # CUDA-graph capture of the pass replays to the same result
    graph, res, ghist = det.capture(x, us, hour0, 3)
    graph.replay()
    torch.cuda.synchronize()
    assert np.array_equal(ghist.cpu().numpy(), hist1)


def test_long_recording_uses_global_workspace_path():
    """> 2048 blocks per file: per-block arrays move from shared memory to the workspace."""
    from meteor_scatter_b200 import ops
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "a_mb_s4_long.npz"))
    assert len(g["delta_power"]) == 7500
    band32, noise32 = g["band_power"].astype(np.float32), g["noise_power"].astype(np.float32)
    delta = band32.astype(np.float64) - noise32.astype(np.float64)
    _, thr_ref, pairs_ref = oa.get_detections_adaptive(delta, 4, 0.2)
    res = ops.detect(_dev(np.stack([band32, band32])), _dev(np.stack([noise32, noise32])), 4, want_thresholds=True)
    for f in range(2):
        n = int(res.counts[f].item())
        assert [tuple(int(v) for v in p) for p in res.events[f, :n].cpu().numpy()] == pairs_ref
    np.testing.assert_allclose(res.thresholds[1].cpu().numpy(), np.asarray(thr_ref), rtol=0, atol=1e-9)


def test_wav_file_process_drop_in(tmp_path, capsys):
    """Detector B end to end (WAV -> Welch bands -> state machine) vs the unmodified reference."""
    from meteor_scatter_b200.dsp.src.live.backend import aggregates as ag
    from meteor_scatter_b200.dsp.src.live.backend.processor import LiveDetector, wav_file_process
    from meteor_scatter_b200.wavio import write_wav_pcm16
    for name in sorted(B_CASES):
        seed, dur, cfgkw, skw = B_CASES[name]
        x, g = b_input(name)
        wav = tmp_path / f"{name}.wav"
        write_wav_pcm16(str(wav), 4000, x)
        cfg = ag.ConfigDetection(**cfgkw)
        dets = wav_file_process(str(wav), cfg, ag.ConfigVisualization(enable_ui_plots=False),
                                ag.ConfigSpecExport(output_dir=""))
        out = capsys.readouterr().out
        assert out.count("Detected Meteor:") == len(dets)
        ref = g["det"]
        # fp32 band power -> db2 differs by ~1e-5 dB from the fp64 reference; decisions sit far from
        # the thresholds in the fixtures, so the event list must be identical
        assert len(dets) == len(ref)
        got = np.array([[d.time_start, d.time_stop, d.duration, d.db_min, d.db_max, d.db_mean, d.db_std]
                        for d in dets]).reshape(-1, 7)
        assert np.array_equal(got[:, :3], ref[:, :3])
        np.testing.assert_allclose(got[:, 3:], ref[:, 3:], rtol=0, atol=2e-3)
        # streaming in 1 s chunks (5 blocks) yields the same detections
        ld = LiveDetector(cfg, fs=4000, n_streams=1)
        xs = _dev(x)
        streamed = []
        for i in range(0, (len(x) // 4000) * 4000, 4000):
            streamed += [d for _, d in ld.push(xs[i:i + 4000])]
        assert [(d.time_start, d.time_stop) for d in streamed] == [(d.time_start, d.time_stop) for d in dets]
    with pytest.raises(AssertionError, match="Invalid Sample Rate"):
        bad = tmp_path / "fs6000.wav"
        write_wav_pcm16(str(bad), 6000, np.zeros(6000, dtype=np.int16))
        wav_file_process(str(bad), ag.ConfigDetection(), ag.ConfigVisualization(enable_ui_plots=False),
                         ag.ConfigSpecExport(), quiet=True)


def test_plot_spectrogram_numeric_stage():
    from meteor_scatter_b200.meteor_detect_class.prime_detection import plot_spectrogram
    from meteor_scatter_b200.synth import synth_file
    x = synth_file(22, fs=5000, dur_s=30.0, carrier_hz=1000.0, rate_per_hour=1200.0)
    ref = oc.plot_spectrogram_numeric(x.reshape(-1, 1), 5000)
    got = plot_spectrogram(x.reshape(-1, 1), 5000, display=False)
    assert got["pxx_db_band"].shape == (164, 145)
    # per-bin tolerance: 1e-4 relative, plus 1e-8 of the frame's peak power for bins in deep spectral nulls
    # (an fp32 FFT's rounding error scales with the strongest component of the frame, not with the bin itself)
    lin = 10.0 ** (got["pxx_db_band"].double().cpu().numpy() / 10.0)
    lin_ref = 10.0 ** (ref["pxx_db_band"] / 10.0)
    assert_rel_counted(lin, lin_ref, "plot_spectrogram_fft_vs_oracle", rtol=REL_TOL + 3e-6,     # + the fp32 dB round trip
                       max_outside=5, floor=1e-8 * lin_ref.max(axis=0, keepdims=True))
    assert np.mean(np.abs(got["pxx_db_band"].cpu().numpy() - ref["pxx_db_band"]) <= DB_TOL) > 0.999
    assert abs(got["density_db_hz"] - ref["density_db_hz"]) < DB_TOL
    assert abs(got["vmin"] - ref["vmin"]) < DB_TOL and got["vmax"] == 40
    np.testing.assert_allclose(got["bins"], ref["bins"])


def test_process_files_writes_dashboard_csv(tmp_path):
    """Batch front end: ragged files, names -> UTC, events per file == oracle, day files == oracle counts."""
    from meteor_scatter_b200.batch import process_files
    from meteor_scatter_b200.synth import synth_file
    from meteor_scatter_b200.wavio import write_wav_pcm16
    paths, starts, xs = [], [], []
    t0 = datetime.datetime(2025, 6, 25, 23, 40, 0)
    for i, dur in enumerate([300.0, 300.0, 181.3, 300.0, 64.0]):
        t = t0 + datetime.timedelta(seconds=300 * i)
        p = tmp_path / f"expoFull_gqrx_{t.strftime('%Y%m%d_%H%M%S')}_49969000.wav"
        x = synth_file(70 + i, dur_s=dur, rate_per_hour=240.0)
        write_wav_pcm16(str(p), 6000, x)
        paths.append(str(p)); starts.append(t); xs.append(x)
    csv_dir = tmp_path / "csv"
    csv_dir.mkdir()
    out = process_files(paths, csv_folder=str(csv_dir))
    ref_hist = {}
    for i, x in enumerate(xs):
        r = oa.detect_wav(x, 6000, 0.2, (993, 1013), (690, 710), 512, 4, wav_start_date_time=starts[i])
        got = out["detections"][i]
        assert [(d.t_start, d.t_stop) for d in got] == [(d.t_start, d.t_stop) for d in r["detections"]]
        assert [d.utc_start for d in got] == [d.utc_start for d in r["detections"]]
        for h, c in oa.hourly_counts(r["detections"]).items():
            a = ref_hist.setdefault(h, [0, 0]); a[0] += c[0]; a[1] += c[1]
    assert sum(v[0] for v in ref_hist.values()) > 0
    rows = {}
    for f in out["csv_files"]:
        lines = open(f).read().splitlines()
        assert lines[0] == "Timestamp;Anzahl;Kritisch"
        for ln in lines[1:]:
            ts, a, k = ln.split(";")
            rows[datetime.datetime.strptime(ts, "%Y-%m-%d %H:%M:%S")] = [int(a), int(k)]
    assert sorted(os.path.basename(f) for f in out["csv_files"]) == ["20250625.csv", "20250626.csv"]
    for h, c in ref_hist.items():
        assert rows[h] == c
    assert sum(v[0] for v in rows.values()) == sum(v[0] for v in ref_hist.values())
    # chunked + prefetched ingest (2 files per chunk, 3 chunks) gives the same events and histogram
    out2 = process_files(paths, chunk_files=2, io_threads=2)
    assert np.array_equal(out2["hist"], out["hist"]) and out2["hour0"] == out["hour0"]
    for i in range(len(paths)):
        assert [(d.t_start, d.t_stop, d.dB) for d in out2["detections"][i]] == \
               [(d.t_start, d.t_stop, d.dB) for d in out["detections"][i]]


def _np_stft_band_energy(x, nfft, hop, window, sig, noise):
    n_frames = (len(x) - len(window)) // hop + 1
    idx = np.arange(len(window))[None, :] + hop * np.arange(n_frames)[:, None]
    spec = np.fft.rfft(x[idx].astype(np.float64) * window[None, :], n=nfft, axis=1)
    p = np.abs(spec) ** 2
    return p[:, sig].sum(axis=1), p[:, noise].sum(axis=1)


@pytest.mark.parametrize("nfft", [1024, 2048, 4096, 8192, 16384])
@pytest.mark.parametrize("overlap", [0.5, 0.75, 0.9])
def test_fft_size_and_overlap_sweep(nfft, overlap):
    """configs[3]: nfft 1024..16384 x 50/75/90 % overlap, periodic Hann (scipy 'hann'), band = carrier +/- 10 Hz,
    noise band 300 Hz below; K1 and the general tensor-core kernel (hop segments read once) for every size, K2
    (resident basis, frames re-read) where the frame fits (nfft = 1024)."""
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.synth import synth_file
    fs = 6000
    x = synth_file(31, fs=fs, dur_s=60.0, rate_per_hour=1200.0)
    hop = max(8, int(round(nfft * (1 - overlap) / 8)) * 8)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(nfft) / nfft)
    freqs = np.fft.rfftfreq(nfft, 1 / fs)
    sig = np.nonzero((freqs >= 993) & (freqs <= 1013))[0]
    noi = np.nonzero((freqs >= 690) & (freqs <= 710))[0][:max(1, 8 - len(sig))] if nfft == 1024 else \
        np.nonzero((freqs >= 690) & (freqs <= 710))[0]
    spec = ops.BandSpec.stft(nfft, hop, w, sig, noi, fs=fs)
    eb_ref, en_ref = _np_stft_band_energy(x, nfft, hop, w, sig, noi)
    xd = _dev(x).reshape(1, -1)
    assert ops.tc_supported(xd, spec)            # every point of the sweep has a tensor-core path
    impls = ["fft", "tc", "seg"] + (["k2"] if ops.k2_supported(xd, spec) else []) + \
        (["rot"] if ops.rot_supported(xd, spec) else [])
    assert ("k2" in impls) == (nfft == 1024)
    assert ("rot" in impls) == (overlap in (0.5, 0.75))     # hop | frame: the window is applied in the frequency domain
    for impl in impls:
        _, _, be, ne = ops.band_power(xd, spec, impl=impl, want_energy=True)
        assert be.shape == (1, len(eb_ref))
        np.testing.assert_allclose(be.cpu().numpy()[0], eb_ref, rtol=REL_TOL, err_msg=impl)
        np.testing.assert_allclose(ne.cpu().numpy()[0], en_ref, rtol=REL_TOL, err_msg=impl)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    dict(bd=0.5, n_fft=1024, fband=(993, 1013), nband=(690, 710)),        # 2048-sample frames in 3000-sample blocks
    dict(bd=0.26, n_fft=1024, fband=(993, 1013), nband=(690, 710)),       # 1560-sample window (> K2's 1408)
    dict(bd=0.2, n_fft=512, fband=(900, 1100), nband=(600, 800)),         # 35 + 35 bins: three column groups
    dict(bd=1.0, n_fft=4096, fband=(1000, 1006), nband=(700, 703)),       # 6000-sample frames, 8192-point transform
    dict(bd=0.2, n_fft=512, fband=(993, 1013), nband=(690, 710)),         # the reference geometry itself
])
def test_general_tensor_core_kernel_matches_oracle(case):
    """csrc/ms_dft_seg.cu in its rows = frames form (hop >= frame): streamed basis, column groups, several files with
    a ragged tail -- energies within 1e-4 of the fp64 oracle, events identical."""
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.synth import synth_file
    bd, n_fft, fband, nband = case["bd"], case["n_fft"], case["fband"], case["nband"]
    files = [synth_file(51 + i, dur_s=50.0, rate_per_hour=3000.0)[:6000 * 50 - 8 * 37] for i in range(3)]
    spec = ops.BandSpec.from_reference_args(6000, bd, fband, nband, n_fft)
    xd = _dev(np.stack(files))
    assert ops.seg_supported(xd, spec)
    bdb, ndb, be, ne = ops.band_power(xd, spec, impl="seg", want_energy=True)
    for i, x in enumerate(files):
        eb_ref, en_ref = oa.stft_band_energy_vec(x, 6000, bd, fband, nband, n_fft)
        np.testing.assert_allclose(be.cpu().numpy()[i], eb_ref, rtol=REL_TOL)
        np.testing.assert_allclose(ne.cpu().numpy()[i], en_ref, rtol=REL_TOL)
        np.testing.assert_allclose(bdb.cpu().numpy()[i], 10 * np.log10(eb_ref + 1e-12), atol=5e-4)
    # without the energy accumulators (single column group writes dB straight from the epilogue)
    bdb2, ndb2 = ops.band_power(xd, spec, impl="seg")
    assert torch.equal(bdb2, bdb) and torch.equal(ndb2, ndb)


@pytest.mark.gpu
@pytest.mark.parametrize("wname", ["boxcar", "hann", "hamming", "blackman"])
def test_frequency_domain_window_form_matches_numpy(wname):
    """Overlapping frames under scipy's periodic cosine-series windows: unwindowed per-hop partial sums on the tensor
    cores + phase rotation + the window applied in the frequency domain (ms_dft_seg_projections_i16 +
    ms_window_combine) equal numpy's windowed STFT; bands at DC and at Nyquist (extended bins below 0 / above nfft/2),
    several files, hop = frame/2, /4 and /8, a 72-bin band (three projection launches)."""
    import scipy.signal as ss
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.synth import synth_file
    fs = 6000
    for nfft, hop, sig, noi, n_files in ((1024, 512, (170, 172), (118, 121), 2), (2048, 512, (0, 3), (1020, 1024), 1),
                                         (512, 64, (80, 151), (20, 22), 3), (4096, 2048, (678, 691), (471, 484), 1)):
        w = ss.get_window(wname, nfft)
        sig_b, noi_b = np.arange(sig[0], sig[1] + 1), np.arange(noi[0], noi[1] + 1)
        spec = ops.BandSpec.stft(nfft, hop, w, sig_b, noi_b, fs=fs)
        files = [synth_file(71 + i, fs=fs, dur_s=20.0, rate_per_hour=2400.0)[:fs * 20 - 8 * (3 + i)] for i in range(n_files)]
        files = [f[:len(files[-1])] for f in files]
        xd = _dev(np.stack(files))
        assert ops.rot_supported(xd, spec), (wname, nfft, hop)
        bdb, ndb, be, ne = ops.band_power(xd, spec, impl="rot", want_energy=True)
        b2, n2 = ops.band_power(xd, spec, impl="tc")            # "tc" picks this form from 75 % overlap on
        if nfft // hop >= 4:
            assert torch.equal(b2, bdb) and torch.equal(n2, ndb)
        else:
            assert float((b2 - bdb).abs().max()) < 2 * DB_TOL
        for i, x in enumerate(files):
            eb_ref, en_ref = _np_stft_band_energy(x, nfft, hop, w, sig_b, noi_b)
            assert be.shape[1] == len(eb_ref)
            tag = f"rot_{wname}_{nfft}_{hop}_file{i}"
            assert_rel_counted(be.cpu().numpy()[i], eb_ref, tag + "_band", rtol=REL_TOL + 2e-7)
            assert_rel_counted(ne.cpu().numpy()[i], en_ref, tag + "_noise", rtol=REL_TOL + 2e-7)


@pytest.mark.gpu
def test_general_tensor_core_kernel_overlap_many_files_and_ragged_tail():
    """Segment form (hop < frame) over several files whose length is not a multiple of the hop: the frames that would
    need a row past the last whole segment run in the rows = frames form; every frame equals numpy's STFT."""
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.synth import synth_file
    fs, nfft = 6000, 2048
    for hop, n_files in ((208, 3), (1024, 2), (56, 1)):
        files = [synth_file(61 + i, fs=fs, dur_s=30.0, rate_per_hour=2400.0)[:fs * 30 - 8 * (5 + i)] for i in range(n_files)]
        files = [f[:len(files[-1])] for f in files]
        w = np.hanning(nfft)
        freqs = np.fft.rfftfreq(nfft, 1 / fs)
        sig = np.nonzero((freqs >= 993) & (freqs <= 1013))[0]
        noi = np.nonzero((freqs >= 690) & (freqs <= 710))[0]
        spec = ops.BandSpec.stft(nfft, hop, w, sig, noi, fs=fs)
        xd = _dev(np.stack(files))
        _, _, be, ne = ops.band_power(xd, spec, impl="seg", want_energy=True)
        for i, x in enumerate(files):
            eb_ref, en_ref = _np_stft_band_energy(x, nfft, hop, w, sig, noi)
            assert be.shape[1] == len(eb_ref)
            np.testing.assert_allclose(be.cpu().numpy()[i], eb_ref, rtol=REL_TOL, err_msg=f"hop {hop} file {i}")
            np.testing.assert_allclose(ne.cpu().numpy()[i], en_ref, rtol=REL_TOL, err_msg=f"hop {hop} file {i}")


def test_full_size_properties_24h():
    """BASELINE configs[1] at full size (288 files x 1.8 M samples = 432 000 blocks, 1.04 GB):
    size-independent properties instead of an oracle run --
    (a) K2 (exact integer DFT) and K1 (fp32 FFT) agree on every block within the 1e-4 budget;
    (b) the whole batch in one launch == the same files processed one by one (tile boundaries cross files);
    (c) doubling the samples adds 20*log10(2) dB to both bands, so delta, thresholds and events are unchanged;
    (d) every event satisfies 0 <= start < stop <= n_blocks, events are ordered and non-adjacent, and the
        hourly histogram sums to the number of events; (e) oracle check on a sample of files."""
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams, datetime_to_us
    from meteor_scatter_b200.synth import synth_batch_torch
    n_files, spf = 288, 1_800_000
    x = synth_batch_torch(n_files, spf, seed=7, device="cuda")
    det = DetectorA(DetectorAParams(), impl="tc")
    t0 = datetime.datetime(2025, 6, 1)
    us = torch.tensor([datetime_to_us(t0 + datetime.timedelta(seconds=300 * i)) for i in range(n_files)],
                      dtype=torch.int64, device="cuda")
    hist = torch.zeros((24, 2), dtype=torch.int32, device="cuda")
    res = det.run_pass(x, us, t0, 24, hist)
    band, noise = res.band_db.clone(), res.noise_db.clone()
    counts = res.det.counts.cpu().numpy().copy()
    events = res.det.events.cpu().numpy().copy()
    assert band.shape == (n_files, 1500)
    # (a)
    bf, nf = ops.band_power(x, det.spec, impl="fft")
    assert float((band - bf).abs().max()) < 2 * DB_TOL and float((noise - nf).abs().max()) < 2 * DB_TOL
    # (b)
    for f in (0, 143, 287):
        b1, n1 = ops.band_power(x[f:f + 1].clone(), det.spec, impl="tc")
        assert torch.equal(b1[0], band[f]) and torch.equal(n1[0], noise[f])
    # (c) exact in integer arithmetic: X doubles, energies x4
    half = (x[:16] // 2) * 1
    b2, n2 = ops.band_power(half * 2, det.spec, impl="tc")
    b1, n1 = ops.band_power(half, det.spec, impl="tc")
    step = 20 * np.log10(2.0)
    assert float((b2 - b1 - step).abs().max()) < 1e-4 and float((n2 - n1 - step).abs().max()) < 1e-4
    r1 = DetectorA(DetectorAParams(), impl="tc").run(half)
    r2 = DetectorA(DetectorAParams(), impl="tc").run(half * 2)
    assert [r1.pairs(f) for f in range(16)] == [r2.pairs(f) for f in range(16)]
    # (d)
    total = 0
    for f in range(n_files):
        ev = events[f, :counts[f]]
        total += len(ev)
        if len(ev):
            assert ev[:, 0].min() >= 0 and ev[:, 1].max() <= 1500 and np.all(ev[:, 0] < ev[:, 1])
            assert np.all(ev[1:, 0] > ev[:-1, 1])          # ordered, separated by >= 1 quiet block
    h = hist.cpu().numpy()
    assert total > 0 and int(h[:, 0].sum()) == total and np.all(h[:, 1] <= h[:, 0])
    # (e)
    for f in (3, 100, 250):
        r = oa.detect_wav(x[f].cpu().numpy(), 6000, 0.2, (993, 1013), (690, 710), 512, 4)
        assert [tuple(int(v) for v in p) for p in events[f, :counts[f]]] == r["pairs"]


def test_run_host_equals_resident_pass():
    """End-to-end API (pinned host PCM -> strided DMA of the used samples -> kernels -> host results)
    gives the same events and hourly counts as the HBM-resident pass; odd chunking included."""
    from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams, datetime_to_us
    from meteor_scatter_b200.synth import synth_file
    xs = np.stack([synth_file(90 + i, dur_s=300.0, rate_per_hour=200.0) for i in range(11)])
    t0 = datetime.datetime(2025, 6, 1, 23, 10, 0)
    us = torch.tensor([datetime_to_us(t0 + datetime.timedelta(seconds=300 * i)) for i in range(11)],
                      dtype=torch.int64, device="cuda")
    hour0 = t0.replace(minute=0, second=0)
    det = DetectorA(DetectorAParams(), impl="tc")
    hist = torch.zeros((2, 2), dtype=torch.int32, device="cuda")
    r = det.run_pass(_dev(xs), us, hour0, 2, hist)
    torch.cuda.synchronize()
    counts, events, h = r.det.counts.cpu().numpy().copy(), r.det.events.cpu().numpy().copy(), hist.cpu().numpy().copy()
    host = torch.from_numpy(xs).pin_memory()
    for chunk in (4, 24):
        out = DetectorA(DetectorAParams(), impl="tc").run_host(host, us, hour0, 2, chunk_files=chunk)
        torch.cuda.synchronize()
        assert np.array_equal(out["counts"].numpy(), counts)
        assert np.array_equal(out["hist"].numpy(), h)
        for f in range(11):
            assert np.array_equal(out["events"].numpy()[f, :counts[f]], events[f, :counts[f]])
    assert h[:, 0].sum() == counts.sum() > 0


def test_one_24h_recording_detect_matches_oracle():
    """Maximum realistic size for one file: a 24 h recording = 432 000 blocks (per-block arrays live in the
    global workspace, prefix scan runs over 211 tiles).  K3 vs the oracle's literal O(N*W) loop."""
    from meteor_scatter_b200 import ops
    rng = np.random.default_rng(11)
    N = 432_000
    delta = (rng.standard_normal(N) * 2.8 - 1.2).astype(np.float32)
    for a in rng.integers(0, N - 60, size=2500):
        delta[a:a + int(rng.integers(1, 30))] += rng.uniform(8, 35)
    d64 = delta.astype(np.float64)
    _, thr_ref, pairs_ref = oa.get_detections_adaptive(d64, 4, 0.2)
    zeros = torch.zeros((1, N), dtype=torch.float32, device="cuda")
    res = ops.detect(_dev(delta).reshape(1, -1), zeros, 4, max_events=8192, want_thresholds=True)
    n = int(res.counts[0].item())
    assert n == len(pairs_ref) > 1000
    assert [tuple(int(v) for v in p) for p in res.events[0, :n].cpu().numpy()] == pairs_ref
    np.testing.assert_allclose(res.thresholds[0].cpu().numpy(), np.asarray(thr_ref), rtol=0, atol=1e-8)


def test_empty_and_degenerate_batches():
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams
    det = DetectorA(DetectorAParams(), impl="auto")
    r = det.run(torch.zeros((0, 1200 * 10), dtype=torch.int16, device="cuda"))       # no files
    assert r.band_db.shape == (0, 10) and r.det.counts.numel() == 0
    r = det.run(torch.zeros((2, 1199), dtype=torch.int16, device="cuda"))            # files shorter than one block
    assert r.band_db.shape == (2, 0) and r.det.counts.cpu().tolist() == [0, 0]
    r = det.run(torch.zeros((2, 1200), dtype=torch.int16, device="cuda"))            # exactly one silent block
    assert r.det.counts.cpu().tolist() == [0, 0]
    # constant (digital silence) input: delta == 0 everywhere, std == 0, nothing exceeds the threshold
    r = det.run(torch.zeros((1, 1200 * 700), dtype=torch.int16, device="cuda"), want_thresholds=True)
    assert r.det.counts.cpu().tolist() == [0] and float(r.det.thresholds.abs().max()) == 0.0


def test_pipelined_batches_equal_sequential():
    """PassPipeline (detect of batch i on a side stream under the band-power kernel of batch i+1,
    small-footprint detect launch) == run_pass batch by batch, on distinct batches."""
    from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams, PassPipeline, datetime_to_us
    from meteor_scatter_b200.synth import synth_batch_torch
    n_files, spf = 40, 1_800_000
    t0 = datetime.datetime(2025, 6, 1)
    us = torch.tensor([datetime_to_us(t0 + datetime.timedelta(seconds=300 * i)) for i in range(n_files)],
                      dtype=torch.int64, device="cuda")
    batches = [synth_batch_torch(n_files, spf, seed=100 + b, device="cuda") for b in range(5)]
    det = DetectorA(DetectorAParams(), impl="tc")
    ref = []
    hist = torch.zeros((4, 2), dtype=torch.int32, device="cuda")
    for x in batches:
        r = det.run_pass(x, us, t0, 4, hist)
        torch.cuda.synchronize()
        ref.append((r.det.counts.cpu().numpy().copy(), r.det.events.cpu().numpy().copy(),
                    r.det.event_db.cpu().numpy().copy(), hist.cpu().numpy().copy()))
    pipe = PassPipeline(DetectorA(DetectorAParams(), impl="tc"), n_files, spf, 4, "cuda")
    got = []
    slots = []
    for b, x in enumerate(batches):
        slots.append(pipe.submit(x, us, t0))
        if b >= 1:                       # consume batch b-1 while batch b is in flight
            r, h = pipe.wait(slots[b - 1])
            got.append((r.det.counts.cpu().numpy().copy(), r.det.events.cpu().numpy().copy(),
                        r.det.event_db.cpu().numpy().copy(), h.cpu().numpy().copy()))
    r, h = pipe.wait(slots[-1])
    got.append((r.det.counts.cpu().numpy().copy(), r.det.events.cpu().numpy().copy(),
                r.det.event_db.cpu().numpy().copy(), h.cpu().numpy().copy()))
    assert sum(int(c[0].sum()) for c in ref) > 0
    for (c0, e0, d0, h0), (c1, e1, d1, h1) in zip(ref, got):
        assert np.array_equal(c0, c1) and np.array_equal(h0, h1)
        for f in range(n_files):
            assert np.array_equal(e0[f, :c0[f]], e1[f, :c0[f]])
            np.testing.assert_allclose(d0[f, :c0[f]], d1[f, :c0[f]], rtol=0, atol=1e-9)


@pytest.mark.parametrize("bd,n_fft,fband,nband", [
    (0.1, 512, (993, 1013), (690, 710)),      # block 600 < nfft 1024: zero-padded frame, partial last K slab
    (0.16, 512, (993, 1013), (690, 710)),     # block 960 < 1024: zero padding, 15 slabs
    (0.2, 128, (980, 1030), (680, 730)),      # nfft 256 << block: deep crop, 4 slabs, wide bins (23.4 Hz)
    (0.2, 256, (995, 1010), (695, 705)),      # nfft 512: one signal bin, one noise bin -> 4 of 16 columns used
    (0.3, 512, (993, 1013), (690, 710)),      # block 1800, stride 3600 B
    (0.2, 1024, (1000, 1008), (698, 706)),    # nfft 2048 > block 1200: 1200-sample frames, 19 K slabs, 4 stages
    (0.2, 512, (2990, 3000), (0, 6)),         # Nyquist and DC bins (sin columns are identically zero)
    (0.2, 512, (985, 1025), (690, 730)),      # 7 + 7 bins: signal and noise bands run as two K2 launches
    (0.2, 1024, (993, 1013), (690, 710)),     # nfft 2048 > block: 7 + 7 bins, 1200-sample frames, two launches
])
def test_tc_geometries_match_oracle(bd, n_fft, fband, nband):
    """K2 on unusual frame geometries vs the oracle (and vs K1)."""
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.synth import synth_file
    x = synth_file(41, dur_s=40.0, rate_per_hour=3000.0)
    spec = ops.BandSpec.from_reference_args(6000, bd, fband, nband, n_fft)
    xd = _dev(x).reshape(1, -1)
    assert ops.tc_supported(xd, spec), (spec.block_size, spec.win_len, spec.sig_bins, spec.noise_bins)
    eb_ref, en_ref = oa.stft_band_energy_vec(x, 6000, bd, fband, nband, n_fft)
    for impl in ("tc", "fft"):
        bdb, ndb, be, ne = ops.band_power(xd, spec, impl=impl, want_energy=True)
        # bands at DC / Nyquist hold almost no energy for zero-mean audio: their values may miss 1e-4 of themselves;
        # they are counted (none allowed for the ordinary bands) and bounded by 1e-6 of the largest band energy
        dc = fband[0] >= 2990 or nband[1] <= 6
        tag = f"tc_geometry_{impl}_{bd}_{n_fft}_{fband[0]}_{nband[0]}"
        assert_rel_counted(be.cpu().numpy()[0], eb_ref, tag + "_band", max_outside=len(eb_ref) if dc else 0,
                           floor=1e-6 * float(eb_ref.max()))
        assert_rel_counted(ne.cpu().numpy()[0], en_ref, tag + "_noise", max_outside=len(en_ref) if dc else 0,
                           floor=1e-6 * float(max(en_ref.max(), eb_ref.max())))
    ref = oa.detect_wav(x, 6000, bd, fband, nband, n_fft, 4)
    from meteor_scatter_b200.pipeline import DetectorA, DetectorAParams
    r = DetectorA(DetectorAParams(block_duration_sec=bd, freq_band=fband, noise_band=nband, n_fft=n_fft), impl="tc").run(
        xd, want_near=True)
    if int(r.det.near.sum().item()) == 0:
        assert r.pairs(0) == ref["pairs"]


def test_tc_unsupported_geometries_fall_back_or_raise():
    from meteor_scatter_b200 import ops
    x = torch.zeros((1, 6000 * 10), dtype=torch.int16, device="cuda")
    wide = ops.BandSpec.from_reference_args(6000, 0.2, (900, 1100), (600, 800), 512)        # 35 + 35 bins > 8 per band
    odd = ops.BandSpec.from_reference_args(6000, 0.15, (993, 1013), (690, 710), 512)        # 900-sample blocks: 1800 B rows
    big = ops.BandSpec.from_reference_args(6000, 0.26, (993, 1013), (690, 710), 1024)       # 1560-sample window > 1408
    for spec in (wide, big):         # beyond the resident-basis kernel, inside the general one
        assert not ops.k2_supported(x, spec) and ops.seg_supported(x, spec) and ops.tc_supported(x, spec)
        with pytest.raises(ops.MsUnsupported):
            ops.band_power(x, spec, impl="k2")
        b, n = ops.band_power(x, spec, impl="tc")
        assert torch.all(b == -120.0)
    assert not ops.tc_supported(x, odd)
    with pytest.raises(ops.MsUnsupported):
        ops.band_power(x, odd, impl="tc")
    b, n = ops.band_power(x, odd, impl="auto")               # falls back to the FFT kernel
    assert torch.all(b == -120.0)
    # float32: only values that are exactly PCM16 / 32768 can take the integer tensor-core path
    mb = ops.BandSpec.from_reference_args(6000, 0.2, (993, 1013), (690, 710), 512)
    assert not ops.tc_supported(x.float(), mb)
    xf = torch.rand((1, 6000 * 10), device="cuda") * 0.1
    with pytest.raises(ops.MsUnsupported):
        ops.band_power(xf, mb, impl="tc")
    b_auto, _ = ops.band_power(xf, mb, impl="auto")
    b_fft, _ = ops.band_power(xf, mb, impl="fft")
    assert torch.equal(b_auto, b_fft)
    xq = (torch.randint(-2000, 2000, (1, 6000 * 10), device="cuda").float() / 32768.0)
    b_tc, _ = ops.band_power(xq, mb, impl="tc")
    b_fft, _ = ops.band_power(xq, mb, impl="fft")
    assert float((b_tc - b_fft).abs().max()) < 2e-3


def test_int16_non_power_of_two_transform_runs_on_the_tensor_cores():
    """n_fft = 500 (a 1000-point transform) is legal for the reference (np.fft.rfft takes any n); the FFT kernel
    only has power-of-two sizes, the general tensor-core kernel takes any."""
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.synth import synth_file
    x = synth_file(43, dur_s=30.0, rate_per_hour=3000.0)
    spec = ops.BandSpec.from_reference_args(6000, 0.2, (993, 1013), (690, 710), 500)
    assert spec.n_fft_real == 1000 and spec.win_len == 1000
    eb_ref, en_ref = oa.stft_band_energy_vec(x, 6000, 0.2, (993, 1013), (690, 710), 500)
    for impl in ("auto", "tc"):
        _, _, be, ne = ops.band_power(_dev(x).reshape(1, -1), spec, impl=impl, want_energy=True)
        assert_rel_counted(be.cpu().numpy()[0], eb_ref, f"nfft1000_{impl}_band")
        assert_rel_counted(ne.cpu().numpy()[0], en_ref, f"nfft1000_{impl}_noise")


@pytest.mark.gpu
def test_frames_within_epsilon_of_threshold_are_flagged():
    """north_star: event indices are bit-exact except frames within a stated epsilon (1e-3 dB) of the
    threshold, which are reported separately -> out_near marks exactly those frames."""
    from meteor_scatter_b200 import ops
    rng = np.random.default_rng(21)
    N = 1500
    delta = (rng.standard_normal(N) * 2.5).astype(np.float32)
    delta[700:712] += 30.0
    d64 = delta.astype(np.float64)
    _, thr, _ = oa.get_detections_adaptive(d64, 4, 0.2)
    thr = np.asarray(thr)
    # put three quiet, unfrozen frames just inside / outside the epsilon band around their own threshold
    # (a frame's threshold does not depend on the frame itself: the window excludes it)
    picks = {300: +4e-4, 400: -6e-4, 500: +5e-3}
    for i, off in picks.items():
        delta[i] = np.float32(thr[i] + off)
    d64 = delta.astype(np.float64)
    _, thr2, pairs_ref = oa.get_detections_adaptive(d64, 4, 0.2)
    thr2 = np.asarray(thr2)
    zeros = torch.zeros((1, N), dtype=torch.float32, device="cuda")
    res = ops.detect(_dev(delta).reshape(1, -1), zeros, 4, want_thresholds=True, want_near=True, eps_db=1e-3)
    near = res.near[0].cpu().numpy().astype(bool)
    expect = np.abs(d64 - thr2) < 1e-3
    # allow disagreement only where |delta - thr| is within float rounding of the epsilon itself
    fuzzy = np.abs(np.abs(d64 - thr2) - 1e-3) < 1e-6
    assert np.array_equal(near[~fuzzy], expect[~fuzzy])
    assert near.sum() >= 1 and not near[500]
    n = int(res.counts[0].item())
    assert [tuple(int(v) for v in p) for p in res.events[0, :n].cpu().numpy()] == pairs_ref


def test_live_waterfall_rows_and_crops():
    """(f)-3: the live detector's waterfall ring on the device.  Rows == 10*log10(welch PSD) of the reference
    (processor.py:206-207) for the +/-100 Hz display bins; crops follow the reference's export window rule."""
    from scipy.signal import welch
    from meteor_scatter_b200.dsp.src.live.backend import aggregates as ag
    from meteor_scatter_b200.dsp.src.live.backend.processor import LiveDetector
    name = "b_live1_s11"
    seed, dur, cfgkw, skw = B_CASES[name]
    x, g = b_input(name)
    cfg = ag.ConfigDetection(**cfgkw)
    viz = ag.ConfigVisualization(enable_ui_plots=False, max_range_sec=60)
    ld = LiveDetector(cfg, fs=4000, n_streams=1, waterfall=viz, export=ag.ConfigSpecExport(output_dir=""))
    xs = _dev(x)
    crops, dets = [], []
    for i in range(0, (len(x) // 4000) * 4000, 4000):
        dets += ld.push(xs[i:i + 4000])
        crops += ld.export_ready()
    assert ld.rows == (922, 1147) or ld.rows[1] - ld.rows[0] + 1 == len(ld.row_freqs)
    assert len(ld.row_freqs) > 200 and abs(ld.row_freqs[0] - 920) < 2 and abs(ld.row_freqs[-1] - 1120) < 2
    # ring rows vs scipy for the last 300 blocks (60 s)
    xf = x.astype(np.float64) / 32768.0
    nblk = len(x) // 800
    ring = ld.ring[0].cpu().numpy()
    worst = 0.0
    for b in range(nblk - 300, nblk, 37):
        f, psd = welch(xf[b * 800:(b + 1) * 800], 4000, nfft=4096)
        ref = 10 * np.log10(psd[ld.rows[0]:ld.rows[1] + 1])
        got = ring[b % ld.ring_len]
        lin, lin_ref = 10.0 ** (got / 10.0), 10.0 ** (ref / 10.0)
        assert np.all(np.abs(lin - lin_ref) <= REL_TOL * lin_ref + 1e-7 * lin_ref.max())
        worst = max(worst, float(np.max(np.abs(got - ref)[ref > ref.max() - 40])))
    assert worst < 5e-3
    # every detection whose +/-3 s window fitted in the ring before the stream ended was exported exactly once
    assert len(dets) == len(g["det"])
    assert len(crops) > 0 and len({(c["meteor"].time_start, c["meteor"].time_stop) for c in crops}) == len(crops)
    for c in crops:
        m = c["meteor"]
        assert c["db"].shape == (len(ld.row_freqs), len(c["times"]))
        assert c["times"][0] >= m.time_start - 3 - 1e-9 and c["times"][-1] <= m.time_stop + 3 + 1e-9
        assert len(c["times"]) >= int(round((m.duration + 6) / 0.2)) - 1
        # the crop's loudest cell sits inside the event and near the carrier
        db = c["db"].cpu().numpy()
        kmax, tmax = np.unravel_index(np.argmax(db), db.shape)
        assert abs(c["freqs"][kmax] - 1020) < 30


def test_tc_full_scale_inputs_do_not_overflow():
    """int32 TMEM accumulators at the extremes: uniformly random full-range PCM16, all -32768, all +32767 and a
    full-scale tone at the signal bin, against the fp64 oracle (the slice sums stay below 2^31 by construction:
    1024 * (255*128 + 255*64) < 2^26)."""
    from meteor_scatter_b200 import ops
    rng = np.random.default_rng(77)
    nb = 300
    n = np.arange(nb * 1200)
    xs = np.stack([
        rng.integers(-32768, 32768, size=nb * 1200, dtype=np.int64).astype(np.int16),
        np.full(nb * 1200, -32768, dtype=np.int16),
        np.full(nb * 1200, 32767, dtype=np.int16),
        np.clip(np.rint(32767 * np.sin(2 * np.pi * 1001.953125 * n / 6000)), -32768, 32767).astype(np.int16),
        np.where((n // 3) % 2 == 0, 32767, -32768).astype(np.int16),
    ])
    spec = _spec(MB)
    for impl in ("tc", "fft"):
        _, _, be, ne = ops.band_power(_dev(xs), spec, impl=impl, want_energy=True)
        for f in range(xs.shape[0]):
            eb, en = oa.stft_band_energy_vec(xs[f], 6000, 0.2, (993, 1013), (690, 710), 512)
            tol = REL_TOL if impl == "tc" else 10 * REL_TOL      # fp32 FFT: leakage bins of a full-scale tone
            scale = max(float(eb.max()), float(en.max()))
            # constant full-scale inputs (files 1, 2) put nothing but window leakage into the bands: those values are
            # counted as exceptions bounded by 1e-9 of the input's largest band energy; every other file allows none
            const = f in (1, 2)
            assert_rel_counted(be[f].cpu().numpy(), eb, f"full_scale_{impl}_file{f}_band", rtol=tol,
                               max_outside=len(eb) if const else 0, floor=1e-9 * scale)
            assert_rel_counted(ne[f].cpu().numpy(), en, f"full_scale_{impl}_file{f}_noise", rtol=tol,
                               max_outside=len(en) if const else 0, floor=1e-9 * scale)


@pytest.mark.parametrize("name", sorted(B_CASES))
def test_welch_quadform_matches_reference(name):
    """B-psd/B-band as a low-rank quadratic form (no FFT): same levels as the unmodified reference
    (scipy.signal.welch + band sums), within the 1e-4 energy budget, for PCM16 and float input."""
    from meteor_scatter_b200 import ops
    seed, dur, cfgkw, skw = B_CASES[name]
    x, g = b_input(name)
    cfg = ob.ConfigDetection(**cfgkw)
    freqs = np.fft.rfftfreq(cfg.n_fft, 1 / 4000)
    bands = []
    for lo, hi in ob.band_edges(cfg):
        k = np.nonzero((freqs >= lo) & (freqs <= hi))[0]
        bands.append((int(k[0]), int(k[-1])))
    for xin in (_dev(x), _dev(x.astype(np.float32) / 32768.0)):
        out = ops.welch_band_db(xin, 800, cfg.n_fft, bands, 4000.0, impl="qf").cpu().numpy()[0]
        np.testing.assert_allclose(out[:, 0], g["ms_db"], rtol=0, atol=DB_TOL + 1e-5)
        np.testing.assert_allclose(out[:, 1], g["n1_db"], rtol=0, atol=DB_TOL + 1e-5)
        np.testing.assert_allclose(out[:, 2], g["n2_db"], rtol=0, atol=DB_TOL + 1e-5)
        np.testing.assert_allclose(out[:, 3], g["db2"], rtol=0, atol=2 * DB_TOL + 2e-5)
        fft = ops.welch_band_db(xin, 800, cfg.n_fft, bands, 4000.0, impl="fft").cpu().numpy()[0]
        np.testing.assert_allclose(out, fft, rtol=0, atol=2 * DB_TOL)


def test_welch_quadform_strong_out_of_band_tone():
    """Truncation of the quadratic form must not leak a strong tone into a quiet band: full-scale carrier in the
    signal channel, noise channels 90 dB below."""
    from scipy.signal import welch
    from meteor_scatter_b200 import ops
    rng = np.random.default_rng(5)
    n = np.arange(4000 * 20)
    x = (0.9 * np.sin(2 * np.pi * 1020.3 * n / 4000) + 3e-5 * rng.standard_normal(len(n))).astype(np.float32)
    cfg = ob.ConfigDetection(signal_freq=1020)
    freqs = np.fft.rfftfreq(4096, 1 / 4000)
    bands = []
    for lo, hi in ob.band_edges(cfg):
        k = np.nonzero((freqs >= lo) & (freqs <= hi))[0]
        bands.append((int(k[0]), int(k[-1])))
    out = ops.welch_band_db(_dev(x), 800, 4096, bands, 4000.0, impl="qf").cpu().numpy()[0]
    for b in (0, 7, 50, 99):
        f, psd = welch(x[b * 800:(b + 1) * 800].astype(np.float64), 4000, nfft=4096)
        ref = [10 * np.log10(psd[lo:hi + 1].sum()) for lo, hi in bands]
        # fp32 arithmetic: the noise channels sit ~85 dB under the carrier, i.e. at the float32 leakage floor of the
        # projections; they must still agree to a fraction of a dB and the signal channel to the usual budget
        assert abs(out[b, 0] - ref[0]) < DB_TOL + 1e-5
        assert abs(out[b, 1] - ref[1]) < 0.5 and abs(out[b, 2] - ref[2]) < 0.5


@pytest.mark.parametrize("name", sorted(B_CASES))
def test_welch_tensor_core_matches_reference(name):
    """B-psd/B-band on the tensor cores (tcgen05 kind::i8 quadratic form, csrc/ms_welch_i8.cu): same levels as the
    unmodified reference (scipy.signal.welch + band sums) within the 1e-4 energy budget, and it is what "auto" picks
    for PCM16."""
    from meteor_scatter_b200 import ops
    seed, dur, cfgkw, skw = B_CASES[name]
    x, g = b_input(name)
    cfg = ob.ConfigDetection(**cfgkw)
    freqs = np.fft.rfftfreq(cfg.n_fft, 1 / 4000)
    bands = []
    for lo, hi in ob.band_edges(cfg):
        k = np.nonzero((freqs >= lo) & (freqs <= hi))[0]
        bands.append((int(k[0]), int(k[-1])))
    n = len(x)                                              # a single stream: any length, whole blocks are used
    xin = _dev(x[:n])
    nb = n // 800
    out = ops.welch_band_db(xin, 800, cfg.n_fft, bands, 4000.0, impl="tc").cpu().numpy()[0]
    assert out.shape == (nb, 4)
    np.testing.assert_allclose(out[:, 0], g["ms_db"][:nb], rtol=0, atol=DB_TOL + 1e-5)
    np.testing.assert_allclose(out[:, 1], g["n1_db"][:nb], rtol=0, atol=DB_TOL + 1e-5)
    np.testing.assert_allclose(out[:, 2], g["n2_db"][:nb], rtol=0, atol=DB_TOL + 1e-5)
    np.testing.assert_allclose(out[:, 3], g["db2"][:nb], rtol=0, atol=2 * DB_TOL + 2e-5)
    auto = ops.welch_band_db(xin, 800, cfg.n_fft, bands, 4000.0).cpu().numpy()[0]
    assert np.array_equal(auto, out)


def test_welch_tensor_core_many_streams_ragged_tiles():
    """Several streams whose block count is not a multiple of the 25 blocks a tile holds, more tiles than SMs:
    every (stream, block) must equal the CUDA-core quadratic form to float rounding."""
    from meteor_scatter_b200 import ops
    rng = np.random.default_rng(11)
    n_streams, nb = 7, 613
    x = (400 * rng.standard_normal((n_streams, nb * 800))).astype(np.int16)
    t = np.arange(nb * 800)
    x[3] += (3000 * np.sin(2 * np.pi * 1019.0 * t / 4000) * ((t // 4000) % 3 == 0)).astype(np.int16)
    bands = [(994, 1095), (687, 788), (1301, 1402)]
    xd = _dev(x)
    tc = ops.welch_band_db(xd, 800, 4096, bands, 4000.0, impl="tc").cpu().numpy()
    qf = ops.welch_band_db(xd, 800, 4096, bands, 4000.0, impl="qf").cpu().numpy()
    assert tc.shape == (n_streams, nb, 4)
    np.testing.assert_allclose(tc, qf, rtol=0, atol=2 * DB_TOL)


def test_welch_tensor_core_rejects_unsupported_geometry():
    from meteor_scatter_b200 import ops
    one = _dev(np.zeros(800 * 4 + 4, dtype=np.int16))     # a single stream may have any length (whole blocks are used)
    assert ops.welch_band_db(one, 800, 4096, [(994, 1095), (687, 788), (1301, 1402)], 4000.0, impl="tc").shape == (1, 4, 4)
    x = _dev(np.zeros((2, 800 * 4 + 4), dtype=np.int16))  # stream stride not a multiple of 16 bytes
    with pytest.raises(ops.MsUnsupported):
        ops.welch_band_db(x, 800, 4096, [(994, 1095), (687, 788), (1301, 1402)], 4000.0, impl="tc")
    xf = _dev(np.zeros(800 * 4, dtype=np.float32))
    with pytest.raises(ops.MsUnsupported):
        ops.welch_band_db(xf, 800, 4096, [(994, 1095), (687, 788), (1301, 1402)], 4000.0, impl="tc")


@pytest.mark.parametrize("block,nperseg,nfft,bands", [
    (1024, 256, 4096, [(994, 1095), (687, 788), (1301, 1402)]),      # 7 segments per block -> 18 blocks per tile
    (800, 128, 2048, [(497, 547), (343, 394), (650, 701)]),          # 2 K slabs, 11 segments -> unsupported (> 8)
    (640, 128, 2048, [(497, 547), (343, 394), (650, 701)]),          # 2 K slabs, 9 segments -> unsupported
    (512, 128, 2048, [(497, 547), (343, 394), (650, 701)]),          # 2 K slabs, 7 segments
    (800, 320, 4096, [(994, 1095), (687, 788), (1301, 1402)]),       # needs 30 columns per band -> unsupported
    (800, 320, 4096, [(994, 1060), (687, 750), (1301, 1360)]),       # 5 K slabs, 4 segments, hop 160
])
def test_welch_tensor_core_other_geometries(block, nperseg, nfft, bands):
    """Segment counts, K depths and hops other than the reference's 5 x 256: the tensor-core form must agree with
    scipy.signal.welch band sums wherever it accepts the geometry, and refuse it loudly otherwise."""
    from scipy.signal import welch
    from meteor_scatter_b200 import ops
    rng = np.random.default_rng(21)
    nb = 57
    t = np.arange(nb * block)
    x = (500 * rng.standard_normal((3, nb * block)) + 2000 * np.sin(2 * np.pi * 1019.0 * t / 4000)[None, :] *
         ((t // 2000) % 2 == 0)[None, :]).astype(np.int16)
    hop = nperseg - nperseg // 2
    n_sub = (block - nperseg // 2) // hop
    xd = _dev(x)
    if n_sub > 8 or not ops.WelchQuadform.get(nperseg, nfft, bands, 4000.0, n_sub, xd.device).tc_ok():
        assert n_sub > 8 or (nperseg, bands[0]) == (320, (994, 1095))
        with pytest.raises(ops.MsUnsupported):
            ops.welch_band_db(xd, block, nfft, bands, 4000.0, nperseg=nperseg, impl="tc")
        return
    out = ops.welch_band_db(xd, block, nfft, bands, 4000.0, nperseg=nperseg, impl="tc").cpu().numpy()
    assert out.shape == (3, nb, 4)
    for s_i, b in ((0, 0), (1, 17), (2, nb - 1), (2, 18), (1, 36)):
        f, psd = welch(x[s_i, b * block:(b + 1) * block].astype(np.float64) / 32768.0, 4000, nperseg=nperseg, nfft=nfft)
        ref = np.array([10 * np.log10(psd[lo:hi + 1].sum()) for lo, hi in bands])
        np.testing.assert_allclose(out[s_i, b, :3], ref, rtol=0, atol=DB_TOL + 1e-5)
        assert abs(out[s_i, b, 3] - (ref[0] - 0.5 * (ref[1] + ref[2]))) < 2 * DB_TOL + 2e-5


def test_live_detector_graph_step_matches_push():
    """configs[4] low-latency form: LiveDetector.push_host (H2D + Welch + state machine + D2H counters replayed as one
    CUDA graph) yields exactly the detections of push() on the same chunks."""
    from meteor_scatter_b200.dsp.src.live.backend.aggregates import ConfigDetection
    from meteor_scatter_b200.dsp.src.live.backend.processor import LiveDetector
    from meteor_scatter_b200.synth import synth_file
    x = synth_file(5, fs=4000, dur_s=120.0, carrier_hz=1020.0, rate_per_hour=900.0)
    cfg = ConfigDetection(proc_block_sec=0.2, n_fft=4096, detection_db_over_noise_mean_min=1,
                          detection_dur_min_sec=0.5, signal_freq=1020)
    a = LiveDetector(cfg, fs=4000, n_streams=2, device="cuda")
    b = LiveDetector(cfg, fs=4000, n_streams=2, device="cuda")
    got_a, got_b = [], []
    for i in range(len(x) // 4000):
        c = torch.from_numpy(np.stack([x[i * 4000:(i + 1) * 4000], x[::-1][i * 4000:(i + 1) * 4000].copy()]))
        got_a += a.push(c.cuda())
        got_b += b.push_host(c)
    assert len(got_a) > 0 and got_a == got_b
    assert a.n_blocks == b.n_blocks


def test_live_state_jump_kernel_resumes_like_the_sequential_one():
    """B-state batch form (event-jumping kernel, >= 32 blocks per call) against the per-block kernel (5-block calls) on
    many streams, with call boundaries falling inside Init, locked Detection and Tracking stretches: thresholds
    bit-identical, detections identical in time and within 1e-9 in their dB statistics, same carried state."""
    from meteor_scatter_b200 import _lib, ops
    rng = np.random.default_rng(77)
    n_streams, n = 19, 1777
    db2 = rng.normal(0.0, 1.0, size=(n_streams, n))
    for s_i in range(n_streams):                       # bursts of different lengths; stream 0 stays quiet
        for _ in range(s_i * 2):
            a = int(rng.integers(45, n - 60))
            db2[s_i, a:a + int(rng.integers(1, 40))] += rng.uniform(4, 25)
    db2[3, 700:] = 30.0                                # never falls below the locked threshold: open event at the end
    db2 = db2.astype(np.float32)
    lc = _lib.LiveConfig(block_samples=800, fs=4000.0, k_std=4.0, init_wait_sec=8.0, after_wait_sec=12.0,
                         mean_min_db=1.0, dur_min_sec=0.5, avg_win=40)
    d = _dev(db2)
    seq = ops.LiveStates(n_streams, "cuda")
    thr_seq = torch.cat([ops.live_state_step(seq, lc, d[:, i:i + 5].contiguous(), want_thresholds=True)
                         for i in range(0, n, 5)], dim=1).cpu().numpy()
    jmp = ops.LiveStates(n_streams, "cuda")
    parts, i = [], 0
    for size in [64, 33, 700, 41, 32, 500, 10 ** 9]:
        j = min(n, i + size)
        parts.append(ops.live_state_step(jmp, lc, d[:, i:j].contiguous(), want_thresholds=True))
        i = j
        if i >= n:
            break
    thr_jmp = torch.cat(parts, dim=1).cpu().numpy()
    assert np.array_equal(thr_seq, thr_jmp, equal_nan=True)
    c_seq, c_jmp = seq.det_count.cpu().numpy(), jmp.det_count.cpu().numpy()
    assert np.array_equal(c_seq, c_jmp) and c_seq.sum() > 20 and c_seq[0] == 0
    for s_i in range(n_streams):
        a, b = seq.det[s_i, :c_seq[s_i]].cpu().numpy(), jmp.det[s_i, :c_jmp[s_i]].cpu().numpy()
        assert np.array_equal(a[:, :5], b[:, :5])
        np.testing.assert_allclose(a[:, 5:], b[:, 5:], rtol=0, atol=1e-9)
    # carried state: same machine state, lock, history ring and tracked-event accumulators
    import ctypes as C
    sa = np.frombuffer(seq.buf.cpu().numpy().tobytes(), dtype=np.uint8).reshape(n_streams, -1)
    sb = np.frombuffer(jmp.buf.cpu().numpy().tobytes(), dtype=np.uint8).reshape(n_streams, -1)
    for s_i in range(n_streams):
        A = _lib.LiveState.from_buffer_copy(sa[s_i].tobytes())
        B = _lib.LiveState.from_buffer_copy(sb[s_i].tobytes())
        for f in ("block_index", "state", "hist_len", "hist_pos", "trk_n", "trk_t0", "trk_min", "trk_max"):
            assert getattr(A, f) == getattr(B, f) or (getattr(A, "state") != 2 and f.startswith("trk")), (s_i, f)
        assert (A.locked_threshold == B.locked_threshold) or (np.isnan(A.locked_threshold) and np.isnan(B.locked_threshold))
        assert A.lock_until_sec == B.lock_until_sec
        if A.state == 2:
            assert abs(A.trk_sum - B.trk_sum) < 1e-9 and A.trk_mean_run == B.trk_mean_run
            assert abs(A.trk_m2_run - B.trk_m2_run) < 1e-6
        ha = [A.hist[(A.hist_pos - 1 - k) % 256] for k in range(A.hist_len)]
        hb = [B.hist[(B.hist_pos - 1 - k) % 256] for k in range(B.hist_len)]
        assert ha == hb
    assert _lib.LiveState.from_buffer_copy(sb[3].tobytes()).state == 2


def test_welch_auto_falls_back_to_fft_for_wide_bands():
    """A band that needs more than 32 quadratic-form columns (150 Hz of a 8192-point Welch at 256-sample segments) is
    served by the FFT form when the caller leaves the choice to "auto" (found by tools/fuzz_detector_b.py)."""
    from scipy.signal import welch
    from meteor_scatter_b200 import ops
    rng = np.random.default_rng(8)
    x = (300 * rng.standard_normal(1000 * 6)).astype(np.int16)
    bands = [(2919, 3225), (2304, 2611), (3533, 3840)]
    with pytest.raises(ops.MsUnsupported):
        ops.welch_band_db(_dev(x), 1000, 8192, bands, 4000.0, impl="qf")
    out = ops.welch_band_db(_dev(x), 1000, 8192, bands, 4000.0).cpu().numpy()[0]
    for b in range(6):
        f, psd = welch(x[b * 1000:(b + 1) * 1000].astype(np.float64) / 32768.0, 4000, nfft=8192)
        ref = np.array([10 * np.log10(psd[lo:hi + 1].sum()) for lo, hi in bands])
        np.testing.assert_allclose(out[b, :3], ref, rtol=0, atol=DB_TOL + 1e-5)


def test_spectrogram_jpg_feeds_detect_and_cluster_bursts(tmp_path):
    """the JPG rendered from the GPU spectrogram (render.save_spectrogram_jpg: imshow(vmin, vmax=40),
    ylim(800, 1200), axes off, 496 x 370 px) is a valid input of the reference's image stage: ORB + DBSCAN find the
    injected bursts, a noise-only segment yields no critical cluster, and the counts are the same when the same
    renderer is fed by the CPU oracle's spectrogram (i.e. the GPU numerics do not move a single cluster)."""
    cv2 = pytest.importorskip("cv2")
    pytest.importorskip("sklearn")
    from meteor_scatter_b200 import render
    from meteor_scatter_b200.meteor_detect_class.detector_and_classification import detect_and_cluster_bursts
    from meteor_scatter_b200.meteor_detect_class.prime_detection import plot_spectrogram
    from oracle import detector_c as oc
    fs, n = 5000, 150000
    rng = np.random.default_rng(5)
    t = np.arange(n) / fs
    noise = rng.standard_normal(n) * 200.0
    bursts = np.zeros(n)
    for t0, dur, amp in ((4.0, 3.0, 6000.0), (12.0, 0.15, 9000.0), (20.0, 2.0, 5000.0)):
        m = (t >= t0) & (t < t0 + dur)
        bursts[m] += amp * np.sin(2 * np.pi * 1000.0 * t[m])
    counts = {}
    for name, sig in (("bursts", noise + bursts), ("noise", noise)):
        x = np.clip(np.rint(sig), -32768, 32767).astype(np.int16).reshape(-1, 1)
        out = plot_spectrogram(x, fs, display=False, out_path=str(tmp_path / f"{name}.jpg"))
        img = cv2.imread(out["image_path"], cv2.IMREAD_COLOR)
        assert img.shape == (370, 496, 3)
        gpu = detect_and_cluster_bursts(out["image_path"], display=False)
        ref = oc.plot_spectrogram_numeric(x, fs)
        p_cpu = str(tmp_path / f"{name}_cpu.jpg")
        render.save_spectrogram_jpg(p_cpu, ref["pxx_db_band"], ref["vmin"], 40)
        cpu = detect_and_cluster_bursts(p_cpu, display=False)
        counts[name] = (len(gpu[3]), len(gpu[4]))
        assert counts[name] == (len(cpu[3]), len(cpu[4])), name
    assert counts["bursts"][0] >= 2 and counts["noise"][0] == 0, counts


def test_event_crops_match_scipy(tmp_path):
    """Detector A's per-event spec_and_psd crops (dsp/src/main.py:721-806, 40-124) from the GPU against
    scipy.signal.spectrogram / welch in the reference's call form; PNG files written by proc_wav_file."""
    import scipy.signal as ss
    from meteor_scatter_b200.dsp.src.main import proc_wav_file
    from meteor_scatter_b200.synth import synth_file
    from meteor_scatter_b200.wavio import write_wav_pcm16
    x = synth_file(5, dur_s=90.0, rate_per_hour=600.0)
    wav = str(tmp_path / "a.wav")
    write_wav_pcm16(wav, 6000, x)
    os.makedirs(tmp_path / "out")
    r = proc_wav_file(wav, 0.2, (993, 1013), (690, 710), 512, 4, outfile_path=str(tmp_path / "out") + "/", quiet=True)
    assert len(r["detections"]) >= 2 and len(r["crops"]) == len(r["detections"])
    export_dir = [d for d in os.listdir(tmp_path / "out")][0]
    pngs = os.listdir(tmp_path / "out" / export_dir)
    assert len(pngs) == len(r["detections"]) and all(p.startswith("spec_and_psd_") and p.endswith(".png") for p in pngs)
    for det, crop in zip(r["detections"], r["crops"]):
        cut0 = max(det.t_start - 3, 0)
        cut = x[int(cut0 * 6000):int(min(det.t_stop + 3, len(x) / 6000) * 6000)].astype(np.float64)
        n = crop["n_fft"]
        assert n == (2048 if len(cut) / 6000 > 8 else 1024)
        f, t, sxx = ss.spectrogram(cut, fs=6000, window="hann", nperseg=n, noverlap=n // 2, nfft=n, scaling="density",
                                   mode="psd")
        m = (f >= 943) & (f <= 1063)
        np.testing.assert_allclose(crop["f"], f[m])
        np.testing.assert_allclose(crop["t"], t)
        np.testing.assert_allclose(crop["sxx_db"].cpu().numpy(), 10 * np.log10(sxx[m] + 1e-10), atol=2e-3)
        fp, pxx = ss.welch(cut, fs=6000, window="hann", nperseg=4096, noverlap=2048, nfft=4096, scaling="density")
        mp = (fp >= 943) & (fp <= 1063)
        np.testing.assert_allclose(crop["pxx_db"].cpu().numpy(), 10 * np.log10(pxx[mp] + 1e-10), atol=2e-3)


def test_two_rank_nccl_product_path_equals_oracle():
    """N>1 on hardware: two NCCL ranks run batch.process_files (round-robin file shards, ONE sum-reduce of the hourly
    histogram, rank 0 writes the day CSVs) and rank 0 checks events per file and hourly counts against the oracle
    (tools/check_process_files_dist.py).  Needs two visible GPUs; the driver's single-GPU box skips it."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tools", "check_process_files_dist.py")],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    assert "OK world=2" in p.stdout
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "r02_two_rank_check.log"), "w") as f:
        f.write(p.stdout)


def test_general_tensor_core_kernel_edge_cases():
    """Edge sizes of csrc/ms_dft_seg.cu: no frame, one frame, fewer frames than a row tile, a single file whose length
    is not a multiple of 8 samples, gapped frames (hop > frame), 16384-sample abutting frames (basis streamed in 256
    K slabs), and the largest supported overlap (129 hops per frame)."""
    from meteor_scatter_b200 import ops
    from meteor_scatter_b200.synth import synth_file
    fs = 6000
    x_all = synth_file(91, fs=fs, dur_s=40.0, rate_per_hour=2400.0)

    def check(x, nfft, frame, hop, sig, noi, impl="seg", window=None):
        w = np.hanning(frame) if window is None else window
        spec = ops.BandSpec.stft(nfft, hop, w, np.arange(*sig), np.arange(*noi), fs=fs)
        xd = _dev(x).reshape(1, -1)
        nb = spec.n_blocks(len(x))
        bdb, ndb, be, ne = ops.band_power(xd, spec, impl=impl, want_energy=True)
        assert be.shape == (1, nb)
        if nb == 0:
            return
        idx = np.arange(frame)[None, :] + hop * np.arange(nb)[:, None]
        p2 = np.abs(np.fft.rfft(x[idx].astype(np.float64) * w[None, :], n=nfft, axis=1)) ** 2
        tag = f"seg_edge_{nfft}_{frame}_{hop}_{len(x)}"
        assert_rel_counted(be.cpu().numpy()[0], p2[:, sig[0]:sig[1]].sum(axis=1), tag + "_band", rtol=REL_TOL + 2e-7)
        assert_rel_counted(ne.cpu().numpy()[0], p2[:, noi[0]:noi[1]].sum(axis=1), tag + "_noise", rtol=REL_TOL + 2e-7)

    check(x_all[:1000], 1024, 1024, 256, (170, 173), (118, 122))            # shorter than one frame: nothing
    check(x_all[:1024], 1024, 1024, 256, (170, 173), (118, 122))            # exactly one frame
    check(x_all[:1024 + 256 * 40], 1024, 1024, 256, (170, 173), (118, 122))  # 41 frames < one 128-row tile
    check(x_all[:50003], 2048, 2048, 512, (339, 346), (236, 243))           # odd length, single file
    check(x_all[:200000], 1024, 1024, 1600, (170, 173), (118, 122))         # gapped frames: hop > frame
    check(x_all, 16384, 16384, 16384, (2713, 2768), (1885, 1940))           # long abutting frames, 55-bin bands
    check(x_all[:60000], 1032, 1032, 8, (171, 174), (119, 123))             # 129 hops per frame
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(2048) / 2048)
    check(x_all[:2048 * 3], 2048, 2048, 128, (339, 346), (236, 243), impl="rot", window=w)   # 16 segments per frame
    spec = ops.BandSpec.stft(1040, 8, np.hanning(1040), np.arange(171, 174), np.arange(119, 123), fs=fs)
    assert not ops.seg_supported(_dev(x_all).reshape(1, -1), spec)          # 130 hops per frame: refused, FFT fallback
