// B-state: trailing-history threshold + Init/Detection/Tracking machine of the
// reference's causal "live" detector, resumable across calls (streaming).
//
// Behavioural spec: dsp/src/live/backend/processor.py:393-414 (history, threshold,
// lock override) and 444-510 (state machine); state types
// dsp/src/live/backend/aggregates.py:9-24.  One thread owns one stream: the
// machine is strictly sequential in time, streams are independent.
#include <stdlib.h>

#include "ms_common.cuh"

namespace ms {
namespace {

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
// pairwise_sum) so that np.mean / np.std of the <=256-entry history are
// reproduced bit for bit.  `get(i)` returns element i in list order.
// Non-recursive: n <= 256 means at most one split into two <=128 blocks.
template <typename F>
__device__ __forceinline__ double np_block_sum(F get, int off, int n) {   // n <= 128
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += get(off + i);
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = get(off + j);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += get(off + i + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += get(off + i);
    return res;
}

template <typename F>
__device__ __forceinline__ double np_pairwise_sum(F get, int n) {          // n <= 256
    if (n <= 128) return np_block_sum(get, 0, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_block_sum(get, 0, n2) + np_block_sum(get, n2, n - n2);
}

// Batch form: the history threshold of block j depends only on the db2 series (the avg_win values before j, some of
// them from the previous call's ring), not on the state machine -> one thread per (stream, block) computes
// mean + k*std (and std, needed for the nan-propagating lock) for every block up front.
struct __align__(32) ms_live_pre {
    double thr, std, ts, te;   // history threshold, history std, block start / end time in seconds
};

__global__ void live_thresholds_kernel(const ms_live_state* states, ms_live_config cfg, int64_t n_streams,
                                       const float* db2, int64_t db2_stride, int db2_elem, int64_t n,
                                       ms_live_pre* out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_streams * n) return;
    const int64_t sidx = idx / n, j = idx - sidx * n;
    const ms_live_state* gs = states + sidx;
    const float* in = db2 + sidx * db2_stride;
    const int A = cfg.avg_win;
    const int len0 = gs->hist_len, pos0 = gs->hist_pos;
    int64_t hl = (int64_t)len0 + j;
    if (hl > A) hl = A;
    // element i (chronological) of the history = series index j - hl + i; negative indices live in the ring
    auto hget = [&](int i) -> double {
        const int64_t g = j - hl + i;
        if (g >= 0) return (double)in[g * db2_elem];
        return gs->hist[(pos0 + (int)g + 2 * MS_LIVE_HIST_MAX) % MS_LIVE_HIST_MAX];
    };
    double thr, h_std;
    if (hl == 0) {
        thr = nan("");
        h_std = nan("");
    } else {
        const double h_mean = np_pairwise_sum(hget, (int)hl) / (double)hl;            // processor.py:399
        auto dget = [&](int i) -> double {
            const double d = hget(i) - h_mean;
            return __dmul_rn(d, d);
        };
        h_std = sqrt(np_pairwise_sum(dget, (int)hl) / (double)hl);                      // processor.py:400
        thr = __dadd_rn(h_mean, __dmul_rn(cfg.k_std, h_std));                           // processor.py:404
    }
    const int64_t bi = gs->block_index + j;                                             // processor.py:181-182
    ms_live_pre r;
    r.thr = thr;
    r.std = h_std;
    r.ts = (double)(bi * cfg.block_samples) / cfg.fs;
    r.te = (double)(bi * cfg.block_samples + cfg.block_samples) / cfg.fs;
    out[idx] = r;
}

__global__ void live_state_kernel(ms_live_state* states, ms_live_config cfg, int64_t n_streams, const float* db2,
                                  int64_t db2_stride, int db2_elem, int64_t n, int max_det, double* out_det,
                                  int32_t* out_det_count, double* out_thresholds, const ms_live_pre* pre_blocks) {
    // one thread per stream (one warp per stream was measured slower: 2.05 vs 1.63 ms for 256 streams x 3000 blocks)
    const int64_t sidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= n_streams) return;
    ms_live_state* gs = states + sidx;
    double* hist = gs->hist;   // ring of the last avg_win db2 values, stays in global memory (L1-cached)
    struct {
        int64_t block_index;
        int32_t state, hist_len, hist_pos, trk_n;
        double locked_threshold, lock_until_sec, trk_t0, trk_sum, trk_min, trk_max, trk_mean_run, trk_m2_run;
    } st;
    st.block_index = gs->block_index;
    st.state = gs->state;
    st.hist_len = gs->hist_len;
    st.hist_pos = gs->hist_pos;
    st.trk_n = gs->trk_n;
    st.locked_threshold = gs->locked_threshold;
    st.lock_until_sec = gs->lock_until_sec;
    st.trk_t0 = gs->trk_t0;
    st.trk_sum = gs->trk_sum;
    st.trk_min = gs->trk_min;
    st.trk_max = gs->trk_max;
    st.trk_mean_run = gs->trk_mean_run;
    st.trk_m2_run = gs->trk_m2_run;
    const float* in = db2 + sidx * db2_stride;
    const int A = cfg.avg_win;
    int n_det = out_det_count[sidx];

    // one block of the state machine; `pre`: threshold/std/times come from live_thresholds_kernel (batch form)
    auto step = [&](const int64_t j, const double v, const bool pre, const double thr_pre, const double std_pre,
                    const double ts_pre, const double te_pre) {
        const int64_t bi = st.block_index;
        // processor.py:181-182 (block_start_idx is an exact integer multiple of the block size)
        const double ts = pre ? ts_pre : (double)(bi * cfg.block_samples) / cfg.fs;
        const double te = pre ? te_pre : (double)(bi * cfg.block_samples + cfg.block_samples) / cfg.fs;

        // history = last A values BEFORE appending the current one (processor.py:394-395)
        const int hl = st.hist_len;
        const int first = (st.hist_pos - hl + 2 * MS_LIVE_HIST_MAX) % MS_LIVE_HIST_MAX;
        auto hget = [&](int i) -> double { return hist[(first + i) % MS_LIVE_HIST_MAX]; };
        double thr;
        double h_std = 0.0;
        if (pre) {
            thr = thr_pre;
            h_std = std_pre;
        } else if (hl == 0) {
            thr = nan("");  // np.mean([]) -> nan
            h_std = nan("");
        } else {
            const double h_mean = np_pairwise_sum(hget, hl) / (double)hl;          // processor.py:399
            auto dget = [&](int i) -> double {
                const double d = hget(i) - h_mean;
                return __dmul_rn(d, d);   // no FMA contraction: numpy squares, then sums
            };
            h_std = sqrt(np_pairwise_sum(dget, hl) / (double)hl);                    // processor.py:400
            thr = __dadd_rn(h_mean, __dmul_rn(cfg.k_std, h_std));                       // processor.py:404
        }
        // append current (ring of capacity A); the batch form only needs the ring at the end of the call
        if (!pre || j + A >= n) hist[st.hist_pos] = v;
        st.hist_pos = (st.hist_pos + 1) % MS_LIVE_HIST_MAX;
        if (st.hist_len < A) st.hist_len++;

        if (st.state == 2) {
            thr = st.locked_threshold;                                                  // processor.py:408
        } else if (st.state == 1 && st.lock_until_sec > te) {
            thr = st.locked_threshold;                                                  // processor.py:411-412
        }
        if (out_thresholds) out_thresholds[sidx * n + j] = thr;

        if (st.state == 0) {
            if (ts >= cfg.init_wait_sec) {                                              // processor.py:455
                st.state = 1;
                st.locked_threshold = -1.0;
                st.lock_until_sec = -1.0;
            }
        } else if (st.state == 1) {
            if (v > thr) {                                                              // processor.py:463
                st.state = 2;
                st.locked_threshold = __dadd_rn(thr, __dmul_rn(0.0, h_std));            // processor.py:466 (nan-propagating)
                st.trk_t0 = ts;
                st.trk_n = 0;
                st.trk_sum = 0.0;
                st.trk_min = INFINITY;
                st.trk_max = -INFINITY;
                st.trk_mean_run = 0.0;
                st.trk_m2_run = 0.0;
            }
        } else {
            // Tracking: append first (processor.py:477), then test (478)
            st.trk_n++;
            st.trk_sum += v;
            st.trk_min = fmin(st.trk_min, v);
            st.trk_max = fmax(st.trk_max, v);
            // std of the tracked values (np.std at processor.py:486) from sums shifted by the first value: no
            // division on the per-block path; dB values of one event span a few tens of dB, so fp64 cancellation
            // stays near 1e-13
            if (st.trk_n == 1) st.trk_mean_run = v;
            const double dlt = v - st.trk_mean_run;
            st.trk_m2_run += dlt * dlt;
            if (v < thr) {
                const double dur = ts - st.trk_t0;
                const double m = st.trk_sum / (double)st.trk_n;
                if (m >= cfg.mean_min_db && dur >= cfg.dur_min_sec) {                   // processor.py:481-482
                    if (n_det < max_det) {
                        double* o = out_det + (sidx * max_det + n_det) * 7;
                        o[0] = st.trk_t0;
                        o[1] = ts;
                        o[2] = dur;
                        o[3] = st.trk_min;
                        o[4] = st.trk_max;
                        o[5] = m;
                        const double dbar = m - st.trk_mean_run;
                        const double var = st.trk_m2_run / (double)st.trk_n - dbar * dbar;
                        o[6] = var > 0.0 ? sqrt(var) : 0.0;
                    }
                    ++n_det;
                }
                st.state = 1;                                                           // processor.py:501-504
                st.lock_until_sec = ts + cfg.after_wait_sec;
            }
        }
        st.block_index = bi + 1;
    };

    if (pre_blocks != nullptr) {
        // batch form: the operands of 8 blocks are fetched together (independent loads in flight), then consumed in order
        constexpr int kPf = 8;
        const ms_live_pre* pb = pre_blocks + sidx * n;
        for (int64_t j0 = 0; j0 < n; j0 += kPf) {
            double4 pf[kPf];
            float pv[kPf];
#pragma unroll
            for (int q = 0; q < kPf; ++q) {
                const int64_t jq = j0 + q < n ? j0 + q : n - 1;
                pf[q] = *reinterpret_cast<const double4*>(&pb[jq]);
                pv[q] = in[jq * db2_elem];
            }
#pragma unroll
            for (int q = 0; q < kPf; ++q)
                if (j0 + q < n) step(j0 + q, (double)pv[q], true, pf[q].x, pf[q].y, pf[q].z, pf[q].w);
        }
    } else {
        for (int64_t j = 0; j < n; ++j) step(j, (double)in[j * db2_elem], false, 0.0, 0.0, 0.0, 0.0);
    }
    out_det_count[sidx] = n_det;
    gs->block_index = st.block_index;
    gs->state = st.state;
    gs->hist_len = st.hist_len;
    gs->hist_pos = st.hist_pos;
    gs->trk_n = st.trk_n;
    gs->locked_threshold = st.locked_threshold;
    gs->lock_until_sec = st.lock_until_sec;
    gs->trk_t0 = st.trk_t0;
    gs->trk_sum = st.trk_sum;
    gs->trk_min = st.trk_min;
    gs->trk_max = st.trk_max;
    gs->trk_mean_run = st.trk_mean_run;
    gs->trk_m2_run = st.trk_m2_run;
}

// Batch form of the state machine: one WARP per stream, jumping from event to event instead of walking every block.
// Same semantics as live_state_kernel (processor.py:393-414, 444-510), restated per state:
//   Init      -> the first block with block_start >= init_wait switches to Detection (no detection test in it);
//   Detection -> while lock_until > block_end the threshold is the locked one, otherwise the history threshold;
//                the first block with db2 > threshold starts Tracking (its own db2 is not tracked);
//   Tracking  -> every block is appended, the first one with db2 < locked threshold ends the event.
// Unlocked Detection stretches are searched in a bit mask U (db2 > history threshold, built by the warp up front);
// locked stretches and Tracking are scanned 32 blocks per ballot; the tracked statistics are warp reductions (sum
// order differs from the sequential kernel by rounding only).
__global__ void __launch_bounds__(128)
live_state_jump_kernel(ms_live_state* states, ms_live_config cfg, int64_t n_streams, const float* db2,
                       int64_t db2_stride, int db2_elem, int64_t n, int max_det, double* out_det,
                       int32_t* out_det_count, double* out_thresholds, const ms_live_pre* pre, uint32_t* umask,
                       int64_t words, int use_smem) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t sidx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (sidx >= n_streams) return;     // whole warps leave together
    ms_live_state* gs = states + sidx;
    const ms_live_pre* pb = pre + sidx * n;
    const float* in = db2 + sidx * db2_stride;
    uint32_t* uw = umask + sidx * words;
    double* thr_out = out_thresholds ? out_thresholds + sidx * n : nullptr;
    // the db2 series of this warp's stream is staged in shared memory when it fits (the scans below are chains of
    // dependent reads: ~30 cycles from shared memory instead of an L2 round trip each)
    // (round 2) ... and so are the history thresholds, the mask U and a "history std is not finite" mask: every step of
    // an episode used to be a dependent L2 round trip (U word, pre[j].thr, pre[j].std: ~9 k cycles per episode)
    extern __shared__ double sm_all[];
    const size_t per_warp = (size_t)n * 12 + (size_t)words * 8;       // bytes: n doubles, n floats, 2 x words uint32
    unsigned char* wbase = reinterpret_cast<unsigned char*>(sm_all) + (size_t)(threadIdx.x >> 5) * ((per_warp + 15) & ~(size_t)15);
    double* thr_s = use_smem ? reinterpret_cast<double*>(wbase) : nullptr;
    float* sv = use_smem ? reinterpret_cast<float*>(wbase + (size_t)n * 8) : nullptr;
    uint32_t* u_s = use_smem ? reinterpret_cast<uint32_t*>(wbase + (size_t)n * 12) : nullptr;
    uint32_t* bad_s = use_smem ? u_s + words : nullptr;
    auto V = [&](int64_t j) -> double { return use_smem ? (double)sv[j] : (double)in[j * db2_elem]; };
    auto THR = [&](int64_t j) -> double { return use_smem ? thr_s[j] : pb[j].thr; };
    // processor.py:466: locked = thr + 0 * std (nan-propagating): thr itself unless the history std is nan / inf
    auto LOCK = [&](int64_t j, double thr) -> double {
        if (use_smem) return ((bad_s[j >> 5] >> (j & 31)) & 1u) ? nan("") : thr;
        return __dadd_rn(thr, __dmul_rn(0.0, pb[j].std));
    };
    auto UW = [&](int64_t w) -> unsigned { return use_smem ? u_s[w] : uw[w]; };
    // block start / end times: the same expressions as live_thresholds_kernel (processor.py:181-182), recomputed
    // instead of loaded
    const int64_t bi0 = gs->block_index;
    auto TS = [&](int64_t j) -> double { return (double)((bi0 + j) * cfg.block_samples) / cfg.fs; };
    auto TE = [&](int64_t j) -> double { return (double)((bi0 + j) * cfg.block_samples + cfg.block_samples) / cfg.fs; };

    int state = gs->state;
    double locked = gs->locked_threshold, lock_until = gs->lock_until_sec, t0 = gs->trk_t0;
    int trk_n = gs->trk_n;
    double trk_sum = gs->trk_sum, trk_min = gs->trk_min, trk_max = gs->trk_max, trk_v0 = gs->trk_mean_run,
           trk_sq = gs->trk_m2_run;
    int n_det = out_det_count[sidx];

    // U: block detects under the history threshold (NaN threshold -> false, as in the reference)
    for (int64_t base = 0; base < n; base += 128) {       // four independent 32-block groups per step (loads in flight)
        float v[4];
        double t[4], sd[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int64_t j = base + g * 32 + lane;
            v[g] = j < n ? in[j * db2_elem] : 0.0f;
            t[g] = j < n ? pb[j].thr : 0.0;
            sd[g] = j < n ? pb[j].std : 0.0;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int64_t j = base + g * 32 + lane;
            if (use_smem && j < n) {
                sv[j] = v[g];
                thr_s[j] = t[g];
            }
            const unsigned m = __ballot_sync(full, j < n && (double)v[g] > t[g]);
            const unsigned bad = __ballot_sync(full, j < n && !(fabs(sd[g]) <= 1.7976931348623157e308));   // nan or inf
            if (lane == 0 && base + g * 32 < n) {
                if (use_smem) {
                    u_s[(base >> 5) + g] = m;
                    bad_s[(base >> 5) + g] = bad;
                } else {
                    uw[(base >> 5) + g] = m;
                }
            }
        }
    }
    __syncwarp();

    // first j in [from, n) with pred(j), or n; pred is evaluated for 32 blocks per step
    auto find_first = [&](int64_t from, auto pred) -> int64_t {
        for (int64_t base = from & ~(int64_t)31; base < n; base += 32) {
            const int64_t j = base + lane;
            const unsigned m = __ballot_sync(full, j >= from && j < n && pred(j));
            if (m) return base + __ffs(m) - 1;
        }
        return n;
    };
    auto fill_thr = [&](int64_t a, int64_t b_excl, bool use_hist, double val) {   // thresholds of blocks [a, b_excl)
        if (!thr_out) return;
        for (int64_t j = a + lane; j < b_excl; j += 32) thr_out[j] = use_hist ? THR(j) : val;
    };
    auto start_tracking = [&](int64_t j, double thr) {
        state = 2;
        locked = LOCK(j, thr);                                                          // processor.py:466 (nan-propagating)
        t0 = TS(j);
        trk_n = 0;
        trk_sum = 0.0;
        trk_min = INFINITY;
        trk_max = -INFINITY;
        trk_v0 = 0.0;
        trk_sq = 0.0;
    };

    int64_t cur = 0;
    while (cur < n) {
        if (state == 0) {
            const int64_t j = find_first(cur, [&](int64_t q) { return TS(q) >= cfg.init_wait_sec; });   // :455
            fill_thr(cur, j < n ? j + 1 : n, true, 0.0);
            if (j >= n) break;
            state = 1;
            locked = -1.0;
            lock_until = -1.0;
            cur = j + 1;
        } else if (state == 1) {
            if (lock_until > TE(cur)) {                                                // processor.py:411-412
                // locked stretch: ends at the first block whose end time reaches lock_until, or at a detection
                const int64_t j = find_first(cur, [&](int64_t q) { return !(lock_until > TE(q)) || V(q) > locked; });
                if (j < n && lock_until > TE(j)) {                                     // detection under the lock
                    fill_thr(cur, j + 1, false, locked);
                    start_tracking(j, locked);
                    cur = j + 1;
                } else {                                                                  // lock expired (or chunk ended)
                    fill_thr(cur, j, false, locked);
                    cur = j;
                }
            } else {
                // unlocked: next set bit of U at or after cur (block end times grow, so the lock stays expired)
                int64_t j = n;
                for (int64_t wb = cur >> 5; wb < words; wb += 32) {
                    const int64_t w = wb + lane;
                    unsigned m = (w < words) ? UW(w) : 0u;
                    if (w == (cur >> 5)) m &= ~0u << (cur & 31);
                    const unsigned bal = __ballot_sync(full, m != 0u);
                    if (bal) {
                        const int L = __ffs(bal) - 1;
                        const unsigned mm = __shfl_sync(full, m, L);
                        j = (wb + L) * 32 + __ffs(mm) - 1;
                        break;
                    }
                }
                fill_thr(cur, j < n ? j + 1 : n, true, 0.0);
                if (j >= n) break;
                start_tracking(j, THR(j));                                                // processor.py:463-466
                cur = j + 1;
            }
        } else {
            // Tracking: append, then test (processor.py:477-478); the event ends at the first db2 < locked
            const int64_t j_end = find_first(cur, [&](int64_t q) { return V(q) < locked; });
            const int64_t last = j_end < n ? j_end : n - 1;     // last block appended in this call
            fill_thr(cur, last + 1, false, locked);
            if (trk_n == 0) trk_v0 = V(cur);
            double a_sum = 0.0, a_sq = 0.0, a_min = INFINITY, a_max = -INFINITY;
            for (int64_t j = cur + lane; j <= last; j += 32) {
                const double v = V(j), d = v - trk_v0;
                a_sum += v;
                a_sq += d * d;
                a_min = fmin(a_min, v);
                a_max = fmax(a_max, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a_sum += __shfl_xor_sync(full, a_sum, o);
                a_sq += __shfl_xor_sync(full, a_sq, o);
                a_min = fmin(a_min, __shfl_xor_sync(full, a_min, o));
                a_max = fmax(a_max, __shfl_xor_sync(full, a_max, o));
            }
            trk_n += (int)(last - cur + 1);
            trk_sum += a_sum;
            trk_sq += a_sq;
            trk_min = fmin(trk_min, a_min);
            trk_max = fmax(trk_max, a_max);
            if (j_end >= n) break;
            const double ts = TS(j_end);
            const double dur = ts - t0;
            const double m = trk_sum / (double)trk_n;
            if (m >= cfg.mean_min_db && dur >= cfg.dur_min_sec) {                         // processor.py:481-482
                if (n_det < max_det && lane == 0) {
                    double* o = out_det + (sidx * max_det + n_det) * 7;
                    const double dbar = m - trk_v0;
                    const double var = trk_sq / (double)trk_n - dbar * dbar;
                    o[0] = t0;
                    o[1] = ts;
                    o[2] = dur;
                    o[3] = trk_min;
                    o[4] = trk_max;
                    o[5] = m;
                    o[6] = var > 0.0 ? sqrt(var) : 0.0;
                }
                ++n_det;
            }
            state = 1;                                                                    // processor.py:501-504
            lock_until = ts + cfg.after_wait_sec;
            cur = j_end + 1;
        }
    }

    // history ring: value j of this call sits at (pos0 + j) % MS_LIVE_HIST_MAX; only the last avg_win matter
    const int A = cfg.avg_win;
    const int pos0 = gs->hist_pos, len0 = gs->hist_len;
    __syncwarp();
    for (int64_t j = (n > A ? n - A : 0) + lane; j < n; j += 32)
        gs->hist[(int)((pos0 + j) % MS_LIVE_HIST_MAX)] = V(j);
    if (lane == 0) {
        out_det_count[sidx] = n_det;
        gs->block_index = bi0 + n;
        gs->state = state;
        gs->hist_len = (int32_t)((int64_t)len0 + n > A ? A : len0 + n);
        gs->hist_pos = (int32_t)((pos0 + n) % MS_LIVE_HIST_MAX);
        gs->trk_n = trk_n;
        gs->locked_threshold = locked;
        gs->lock_until_sec = lock_until;
        gs->trk_t0 = t0;
        gs->trk_sum = trk_sum;
        gs->trk_min = trk_min;
        gs->trk_max = trk_max;
        gs->trk_mean_run = trk_v0;
        gs->trk_m2_run = trk_sq;
    }
}

}  // namespace
}  // namespace ms

namespace ms {
namespace {
inline size_t live_pre_bytes(int64_t n_streams, int64_t n) { return (size_t)n_streams * (size_t)n * sizeof(ms_live_pre); }
inline size_t live_mask_bytes(int64_t n_streams, int64_t n) {
    return (size_t)n_streams * (size_t)((n + 31) / 32) * sizeof(uint32_t);
}
}  // namespace
}  // namespace ms

extern "C" int64_t ms_live_state_workspace_bytes(int64_t n_streams, int64_t n) {
    if (n_streams <= 0 || n < 32) return 0;     // streaming-sized calls need no scratch
    return (int64_t)(ms::live_pre_bytes(n_streams, n) + ms::live_mask_bytes(n_streams, n));
}

extern "C" int ms_live_state_step_ws(ms_live_state* states, const ms_live_config* h_cfg, int64_t n_streams,
                                     const float* db2, int64_t db2_stride, int32_t db2_elem, int64_t n,
                                     int32_t max_det, double* out_det, int32_t* out_det_count,
                                     double* out_thresholds, void* workspace, int64_t workspace_bytes, void* stream) {
    MS_REQUIRE(states && h_cfg && db2 && out_det && out_det_count, MS_ERR_INVALID_ARG,
               "ms_live_state_step: null pointer");
    MS_REQUIRE(h_cfg->avg_win >= 1 && h_cfg->avg_win <= MS_LIVE_HIST_MAX, MS_ERR_UNSUPPORTED,
               "ms_live_state_step: avg_win=%d outside [1, %d]", h_cfg->avg_win, MS_LIVE_HIST_MAX);
    MS_REQUIRE(n_streams >= 0 && n >= 0 && max_det > 0 && db2_elem > 0, MS_ERR_INVALID_ARG,
               "ms_live_state_step: bad sizes");
    if (n_streams == 0 || n == 0) return MS_OK;
    const int threads = 32;
    const int64_t blocks = (n_streams + threads - 1) / threads;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // batch form (n >= 32): thresholds of all blocks in parallel, then the event-jumping state machine, one warp per
    // stream.  MS_LIVE_SEQUENTIAL=1 keeps the per-block kernel (cross-check in the tests).
    static const bool force_seq = [] {
        const char* e = getenv("MS_LIVE_SEQUENTIAL");
        return e && atoi(e) != 0;
    }();
    if (n >= 32) {
        const size_t cnt = (size_t)n_streams * (size_t)n;
        const int64_t words = (n + 31) / 32;
        const size_t pre_bytes = ms::live_pre_bytes(n_streams, n);
        MS_REQUIRE(workspace != nullptr && workspace_bytes >= ms_live_state_workspace_bytes(n_streams, n) &&
                       (reinterpret_cast<uintptr_t>(workspace) & 31) == 0,
                   MS_ERR_WORKSPACE, "ms_live_state_step_ws: needs %lld bytes of 32-byte aligned workspace",
                   (long long)ms_live_state_workspace_bytes(n_streams, n));
        char* scratch = static_cast<char*>(workspace);
        ms::ms_live_pre* pre = reinterpret_cast<ms::ms_live_pre*>(scratch);
        const int64_t tb = ((int64_t)cnt + 255) / 256;
        ms::live_thresholds_kernel<<<(unsigned)tb, 256, 0, st>>>(states, *h_cfg, n_streams, db2, db2_stride, db2_elem, n,
                                                                 pre);
        MS_CUDA_OK(cudaGetLastError());
        if (force_seq) {
            ms::live_state_kernel<<<(unsigned)blocks, threads, 0, st>>>(
                states, *h_cfg, n_streams, db2, db2_stride, db2_elem, n, max_det, out_det, out_det_count, out_thresholds, pre);
        } else {
            // one warp per stream; 4 warps per CTA while their db2 series fit into shared memory together, else fewer
            // per warp: the db2 series (float), the history thresholds (double), the mask U and the bad-std mask
            const size_t per_warp = (((size_t)n * 12 + (size_t)words * 8) + 15) & ~(size_t)15;
            int wpc = 4;
            while (wpc > 1 && (size_t)wpc * per_warp > (size_t)160 * 1024) wpc >>= 1;
            const size_t sm_bytes = (size_t)wpc * per_warp;
            const int use_smem = sm_bytes <= (size_t)200 * 1024 ? 1 : 0;
            if (use_smem && sm_bytes > 40 * 1024)
                MS_CUDA_OK(cudaFuncSetAttribute(ms::live_state_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)sm_bytes));
            const int64_t jb = (n_streams + wpc - 1) / wpc;
            ms::live_state_jump_kernel<<<(unsigned)jb, 32 * wpc, use_smem ? sm_bytes : 0, st>>>(
                states, *h_cfg, n_streams, db2, db2_stride, db2_elem, n, max_det, out_det, out_det_count, out_thresholds, pre,
                reinterpret_cast<uint32_t*>(scratch + pre_bytes), words, use_smem);
        }
        MS_CUDA_OK(cudaGetLastError());
        return MS_OK;
    }
    ms::live_state_kernel<<<(unsigned)blocks, threads, 0, st>>>(
        states, *h_cfg, n_streams, db2, db2_stride, db2_elem, n, max_det, out_det, out_det_count, out_thresholds, nullptr);
    MS_CUDA_OK(cudaGetLastError());
    return MS_OK;
}

// Convenience form that brings its own scratch (stream-ordered allocation) for callers without a device allocator.
extern "C" int ms_live_state_step(ms_live_state* states, const ms_live_config* h_cfg, int64_t n_streams,
                                  const float* db2, int64_t db2_stride, int32_t db2_elem, int64_t n, int32_t max_det,
                                  double* out_det, int32_t* out_det_count, double* out_thresholds, void* stream) {
    const int64_t need = ms_live_state_workspace_bytes(n_streams, n);
    if (need == 0)
        return ms_live_state_step_ws(states, h_cfg, n_streams, db2, db2_stride, db2_elem, n, max_det, out_det,
                                     out_det_count, out_thresholds, nullptr, 0, stream);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {   // keep the stream-ordered scratch cached across calls: by default the pool hands freed memory back to the
        // driver at every synchronisation and the next call would pay 2-3 ms for a fresh physical allocation.
        // This raises the release threshold of the device's DEFAULT memory pool (a process-wide setting); callers
        // that mind use ms_live_state_step_ws with their own workspace.
        static bool pool_ready[64] = {};
        int dev = 0;
        MS_CUDA_OK(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64 && !pool_ready[dev]) {
            cudaMemPool_t pool;
            MS_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
            uint64_t keep = UINT64_MAX;
            MS_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
            pool_ready[dev] = true;
        }
    }
    void* scratch = nullptr;
    MS_CUDA_OK(cudaMallocAsync(&scratch, (size_t)need, st));
    const int rc = ms_live_state_step_ws(states, h_cfg, n_streams, db2, db2_stride, db2_elem, n, max_det, out_det,
                                         out_det_count, out_thresholds, scratch, need, stream);
    MS_CUDA_OK(cudaFreeAsync(scratch, st));
    return rc;
}
