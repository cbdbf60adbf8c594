"""CPU model of the event-jumping form of the live state machine (csrc/ms_live.cu, live_state_jump_kernel): the same
control flow -- Init search, mask search while unlocked, bounded scans while locked or tracking -- written with numpy
searches instead of warp ballots, checked against the oracle's literal per-block loop (processor.py:393-510) on random
series, including calls that end in the middle of any state."""
import numpy as np
import pytest

from oracle import detector_b as ob


class JumpModel:
    def __init__(self, cfg, fs, block):
        self.cfg, self.fs, self.block = cfg, fs, block
        self.A = int(cfg.avg_win_sec / cfg.proc_block_sec)
        self.bi = 0
        self.state, self.locked, self.lock_until, self.t0 = 0, -1.0, -1.0, 0.0
        self.hist = []                     # last A db2 values
        self.trk = []                      # tracked values of the open event
        self.dets, self.thr = [], []

    def step(self, v):
        v = np.asarray(v, dtype=np.float64)
        n, cfg = len(v), self.cfg
        series = np.concatenate([np.asarray(self.hist, dtype=np.float64), v])
        off = len(self.hist)
        ts = np.array([(self.bi + j) * self.block / self.fs for j in range(n)])
        te = np.array([((self.bi + j) * self.block + self.block) / self.fs for j in range(n)])
        h_thr, h_std = np.empty(n), np.empty(n)
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for j in range(n):         # the parallel pre-pass: history threshold of every block
                    h = series[max(0, off + j - self.A):off + j]
                    m, s = np.mean(h), np.std(h)
                    h_thr[j], h_std[j] = m + cfg.threshold_std_factor * s, s
        U = v > h_thr                       # NaN threshold -> False
        thr_out = np.empty(n)

        def first(mask, frm):
            idx = np.nonzero(mask[frm:])[0]
            return frm + int(idx[0]) if len(idx) else n

        cur = 0
        while cur < n:
            if self.state == 0:
                j = first(ts >= cfg.init_detection_wait_sec, cur)
                thr_out[cur:min(j + 1, n)] = h_thr[cur:min(j + 1, n)]
                if j >= n:
                    break
                self.state, self.locked, self.lock_until = 1, -1.0, -1.0
                cur = j + 1
            elif self.state == 1:
                if self.lock_until > te[cur]:
                    j = first(~(self.lock_until > te) | (v > self.locked), cur)
                    if j < n and self.lock_until > te[j]:
                        thr_out[cur:j + 1] = self.locked
                        self._start(j, self.locked, h_std[j], ts[j])
                        cur = j + 1
                    else:
                        thr_out[cur:j] = self.locked
                        cur = j
                else:
                    j = first(U, cur)
                    thr_out[cur:min(j + 1, n)] = h_thr[cur:min(j + 1, n)]
                    if j >= n:
                        break
                    self._start(j, h_thr[j], h_std[j], ts[j])
                    cur = j + 1
            else:
                j_end = first(v < self.locked, cur)
                last = j_end if j_end < n else n - 1
                thr_out[cur:last + 1] = self.locked
                self.trk += list(v[cur:last + 1])
                if j_end >= n:
                    break
                dur = ts[j_end] - self.t0
                m = float(np.mean(self.trk))
                if m >= cfg.detection_db_over_noise_mean_min and dur >= cfg.detection_dur_min_sec:
                    self.dets.append((self.t0, ts[j_end], dur, min(self.trk), max(self.trk), m, float(np.std(self.trk))))
                self.state, self.lock_until = 1, ts[j_end] + cfg.after_tracking_wait_sec
                cur = j_end + 1
        self.thr += list(thr_out)
        self.hist = list(series[-self.A:]) if self.A > 0 else []
        self.bi += n

    def _start(self, j, thr, std, ts):
        with np.errstate(invalid="ignore"):
            self.state, self.locked, self.t0, self.trk = 2, thr + 0 * std, ts, []


@pytest.mark.parametrize("seed", range(12))
def test_jump_model_equals_per_block_loop(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(50, 1500))
    bsec = float(rng.choice([0.1, 0.2, 0.5]))
    cfg = ob.ConfigDetection(proc_block_sec=bsec, avg_win_sec=float(rng.choice([bsec, 1, 8])),
                             init_detection_wait_sec=float(rng.choice([0, 8, 30])),
                             after_tracking_wait_sec=float(rng.choice([0, 2, 12])),
                             threshold_std_factor=float(rng.choice([1.5, 4])),
                             detection_db_over_noise_mean_min=float(rng.choice([-1, 2])),
                             detection_dur_min_sec=float(rng.choice([-1, 1])))
    db2 = rng.normal(0, 1, size=n)
    for _ in range(int(rng.integers(0, 10))):
        a = int(rng.integers(0, n))
        db2[a:a + int(rng.integers(1, 60))] += rng.uniform(3, 25)
    db2 = db2.astype(np.float32).astype(np.float64)
    block = int(bsec * 4000)
    dets_ref, thr_ref = ob.live_state_machine(db2, cfg, 4000, block)
    m = JumpModel(cfg, 4000, block)
    i = 0
    while i < n:                           # calls of random size: boundaries fall inside every kind of stretch
        size = int(rng.integers(1, 300))
        m.step(db2[i:i + size])
        i += size
    assert np.array_equal(np.asarray(m.thr), np.asarray(thr_ref, dtype=np.float64), equal_nan=True)
    assert len(m.dets) == len(dets_ref)
    for got, ref in zip(m.dets, dets_ref):
        assert got[:5] == (ref.time_start, ref.time_stop, ref.duration, ref.db_min, ref.db_max)
        assert abs(got[5] - ref.db_mean) < 1e-12 and abs(got[6] - ref.db_std) < 1e-12
