// B-state: trailing-history threshold + Init/Detection/Tracking machine of the
// reference's causal "live" detector, resumable across calls (streaming).
//
// Behavioural spec: dsp/src/live/backend/processor.py:393-414 (history, threshold,
// lock override) and 444-510 (state machine); state types
// dsp/src/live/backend/aggregates.py:9-24.  One thread owns one stream: the
// machine is strictly sequential in time, streams are independent.
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

}  // namespace
}  // namespace ms

extern "C" int ms_live_state_step(ms_live_state* states, const ms_live_config* h_cfg, int64_t n_streams,
                                  const float* db2, int64_t db2_stride, int32_t db2_elem, int64_t n, int32_t max_det,
                                  double* out_det, int32_t* out_det_count, double* out_thresholds, void* stream) {
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
    ms::ms_live_pre* pre = nullptr;
    if (n >= 32) {
        // batch form: thresholds of all blocks in parallel (stream-ordered scratch), then the sequential state logic
        const size_t cnt = (size_t)n_streams * (size_t)n;
        MS_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&pre), cnt * sizeof(ms::ms_live_pre), st));
        const int64_t tb = ((int64_t)cnt + 255) / 256;
        ms::live_thresholds_kernel<<<(unsigned)tb, 256, 0, st>>>(states, *h_cfg, n_streams, db2, db2_stride, db2_elem, n,
                                                                 pre);
        MS_CUDA_OK(cudaGetLastError());
    }
    ms::live_state_kernel<<<(unsigned)blocks, threads, 0, st>>>(
        states, *h_cfg, n_streams, db2, db2_stride, db2_elem, n, max_det, out_det, out_det_count, out_thresholds, pre);
    if (pre) {
        MS_CUDA_OK(cudaGetLastError());
        MS_CUDA_OK(cudaFreeAsync(pre, st));
    }
    MS_CUDA_OK(cudaGetLastError());
    return MS_OK;
}
