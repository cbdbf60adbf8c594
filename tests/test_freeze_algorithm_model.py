"""CPU: lane-by-lane Python model of the warp-ballot formulation used by
csrc/ms_detect.cu (adaptive freeze logic resolved 32 blocks per step) checked
against the oracle's literal loop on adversarial random inputs, including tiny
window/after/before/fixed values.  Guards the *algorithm*; the CUDA kernel is
checked against the oracle in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import detector_a as oa


def model_adaptive(delta, k, W, before, after, fixed):
    N = len(delta)
    mean = delta.mean()
    c = delta - mean
    S1 = np.concatenate([[0.0], np.cumsum(c)])
    S2 = np.concatenate([[0.0], np.cumsum(c * c)])
    g = mean + k * np.sqrt(np.mean(c * c))
    T = np.empty(N)
    for i in range(N):
        if i < fixed:
            T[i] = g
        else:
            w0 = max(0, i - W)
            cnt = i - w0
            if cnt <= 0:
                T[i] = np.nan
            else:
                m = (S1[i] - S1[w0]) / cnt
                v = max((S2[i] - S2[w0]) / cnt - m * m, 0.0) if cnt > 1 else 0.0
                T[i] = (mean + m) + k * np.sqrt(v)
    det = np.zeros(N, dtype=bool)
    thr = np.zeros(N)
    F, H = -1, g
    for base in range(0, N, 32):
        lim = min(base + 32, N)
        lanes = np.arange(base, base + 32)
        valid = lanes < N
        cur = base
        while cur < lim:
            u_cur = (cur < fixed) or (cur > F)
            if u_cur:
                u_p = valid & (lanes >= cur) & ((lanes < fixed) | (lanes > F))
                run_end = cur
                while run_end < base + 32 and u_p[run_end - base]:
                    run_end += 1
                in_run = (lanes >= cur) & (lanes < run_end)
                dm = [int(l) for l in lanes[in_run] if delta[l] > T[l]]
                if not dm:
                    thr[cur:run_end] = T[cur:run_end]
                    H = T[run_end - 1]
                    cur = run_end
                else:
                    q = dm[0]
                    thr[cur:q + 1] = T[cur:q + 1]
                    det[q] = True
                    H = T[q]
                    F = max(q + after, max(0, q - before))
                    cur = q + 1
            else:
                end = min(F, lim - 1)
                idx = np.arange(cur, end + 1)
                hit = idx[delta[idx] > H]
                thr[cur:end + 1] = H
                det[hit] = True
                if len(hit):
                    q = int(hit[-1])
                    F = max(q + after, max(0, q - before))
                cur = end + 1
    # runs -> events
    pairs = []
    i = 0
    while i < N:
        if det[i]:
            j = i
            while j + 1 < N and det[j + 1]:
                j += 1
            pairs.append((i, j + 1))
            i = j + 1
        else:
            i += 1
    return pairs, thr


@pytest.mark.parametrize("seed", range(40))
def test_chunked_freeze_equals_literal_loop(seed):
    rng = np.random.default_rng(seed)
    N = int(rng.integers(1, 400))
    delta = rng.standard_normal(N) * 3.0
    nb = int(rng.integers(0, 12))
    for _ in range(nb):
        a = int(rng.integers(0, N))
        delta[a:a + int(rng.integers(1, 40))] += rng.uniform(3, 25)
    W = int(rng.choice([1, 3, 17, 50, 600]))
    before = int(rng.choice([0, 2, 15, 200]))
    after = int(rng.choice([0, 1, 5, 31, 32, 33, 100]))
    fixed = int(rng.choice([0, 1, 7, 50, 64]))
    k = float(rng.choice([0.5, 1.0, 2.0, 4.0]))
    bd = 0.2
    # drive the oracle with seconds that truncate to the chosen block counts
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dets, thr_ref, pairs_ref = oa.get_detections_adaptive(
            delta, k, bd, None, (W + 0.5) * bd, (before + 0.5) * bd, (after + 0.5) * bd, (fixed + 0.5) * bd)
    assert oa.adaptive_params(bd, (W + 0.5) * bd, (before + 0.5) * bd, (after + 0.5) * bd, (fixed + 0.5) * bd) == \
        (W, before, after, fixed)
    pairs, thr = model_adaptive(delta, k, W, before, after, fixed)
    assert pairs == pairs_ref
    # prefix-sum variance vs numpy's two-pass std: identical to ~1e-13 except for degenerate
    # zero-variance windows (W=1), where sqrt(cancellation) ~ 1e-7 dB -- far inside the 1e-3 dB epsilon
    np.testing.assert_allclose(thr, np.asarray(thr_ref, dtype=np.float64), rtol=0, atol=2e-5, equal_nan=True)


def model_adaptive_v2(delta, k, W, before, after, fixed):
    """Model of the mask-jump formulation in csrc/ms_detect.cu: U = detections under the 'unfrozen'
    hypothesis for every block (parallel), then per burst: jump to the next U bit, evaluate the frozen
    stretch word by word against the held threshold, extend while it keeps detecting."""
    N = len(delta)
    s0 = delta[0]
    c = delta - s0
    S1 = np.concatenate([[0.0], np.cumsum(c)])
    S2 = np.concatenate([[0.0], np.cumsum(c * c)])
    m0 = S1[N] / N
    g = (s0 + m0) + k * np.sqrt(max(S2[N] / N - m0 * m0, 0.0))
    T = np.empty(N)
    for i in range(N):
        if i < fixed:
            T[i] = g
        else:
            w0 = max(0, i - W)
            cnt = i - w0
            if cnt <= 0:
                T[i] = np.nan
            else:
                m = (S1[i] - S1[w0]) / cnt
                v = max((S2[i] - S2[w0]) / cnt - m * m, 0.0) if cnt > 1 else 0.0
                T[i] = (s0 + m) + k * np.sqrt(v)
    U = delta > T
    det = np.zeros(N, dtype=bool)
    thr = T.copy()

    def new_f(q):
        return max(q + after, max(0, q - before))

    F, H = -1, g
    nf = min(fixed, N)
    det[:nf] = U[:nf]                       # fixed region: thr = g regardless of freezing
    hits = np.nonzero(det[:nf])[0]
    if len(hits):
        F = new_f(int(hits[-1]))
    cur = nf
    while cur < N:
        if cur > F:
            nxt = np.nonzero(U[cur:])[0]
            if not len(nxt):
                break
            q = cur + int(nxt[0])
            det[q] = True
            H = T[q]
            F = new_f(q)
            cur = q + 1
        else:
            end = min(F, N - 1)
            idx = np.arange(cur, end + 1)
            hit = idx[delta[idx] > H]
            thr[cur:end + 1] = H
            det[hit] = True
            if len(hit):
                F = new_f(int(hit[-1]))
            cur = end + 1
    pairs = []
    i = 0
    while i < N:
        if det[i]:
            j = i
            while j + 1 < N and det[j + 1]:
                j += 1
            pairs.append((i, j + 1))
            i = j + 1
        else:
            i += 1
    return pairs, thr


@pytest.mark.parametrize("seed", range(60))
def test_mask_jump_freeze_equals_literal_loop(seed):
    rng = np.random.default_rng(1000 + seed)
    N = int(rng.integers(1, 500))
    delta = rng.standard_normal(N) * 3.0
    for _ in range(int(rng.integers(0, 12))):
        a = int(rng.integers(0, N))
        delta[a:a + int(rng.integers(1, 40))] += rng.uniform(3, 25)
    W = int(rng.choice([1, 3, 17, 50, 600]))
    before = int(rng.choice([0, 2, 15, 200]))
    after = int(rng.choice([0, 1, 5, 31, 32, 33, 100]))
    fixed = int(rng.choice([0, 1, 7, 50, 64, 1000]))
    k = float(rng.choice([0.5, 1.0, 2.0, 4.0]))
    bd = 0.2
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dets, thr_ref, pairs_ref = oa.get_detections_adaptive(
            delta, k, bd, None, (W + 0.5) * bd, (before + 0.5) * bd, (after + 0.5) * bd, (fixed + 0.5) * bd)
    pairs, thr = model_adaptive_v2(delta, k, W, before, after, fixed)
    assert pairs == pairs_ref
    np.testing.assert_allclose(thr, np.asarray(thr_ref, dtype=np.float64), rtol=0, atol=2e-5, equal_nan=True)
