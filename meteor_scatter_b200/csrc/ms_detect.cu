// K3: delta -> threshold (global / adaptive with freeze) -> events, and the
// hourly [Anzahl, Kritisch] histogram.
//
// Behavioural spec: dsp/src/main.py:393 (delta), 396-448 (get_detections),
// 450-522 (get_detections_adaptive), 690-696 (hour bucketing) of the reference.
// The reference evaluates the adaptive threshold with an O(N*W) Python loop; here
// one CTA per file computes fp64 prefix sums of the centred delta (O(N)), all
// candidate thresholds in parallel, and one warp resolves the data-dependent
// freeze logic 32 blocks at a time with ballots.  Events are compacted with
// popc prefix sums, so the per-file event list comes out ordered.
#include "ms_common.cuh"

namespace ms {
namespace {

constexpr int kThreads = 256;       // default CTA size
constexpr int kThreadsSmall = 128;  // small-footprint launch that co-resides with the persistent band-power CTAs
constexpr int kItems = 8;  // prefix-sum items per thread per tile (one 2048-block tile covers a 5-minute file)
constexpr int kSmemBlocks = 2048;  // files up to this many blocks keep all per-block arrays in shared memory
constexpr size_t kSmemBytes = (size_t)(4 * (kSmemBlocks + 1)) * sizeof(double);

struct DetectParams {
    const float* band;
    const float* noise;
    const int32_t* n_blocks_per_file;
    int64_t stride;
    int64_t n_blocks;
    double k_std;
    int32_t window, before, after, fixed;
    int32_t max_events;
    int32_t* out_events;
    double* out_event_db;
    int32_t* out_counts;
    double* out_thresholds;
    uint8_t* out_near;
    double eps_db;
    char* workspace;
    int64_t ws_per_file;
    // optional fused A-hour stage (out_hist == nullptr: skip)
    const int64_t* file_start_us;
    double block_duration_sec, crit_min_dur_sec;
    int64_t hour0;
    int32_t n_hours;
    int32_t* out_hist;
    int32_t use_smem;   // 1: delta/S1/S2/T live in dynamic shared memory (n_blocks <= kSmemBlocks)
    int32_t pdl;        // 1: launched with programmatic stream serialization: wait for the producer grid first
    int32_t n_files;
};

__host__ __device__ inline int64_t align16(int64_t v) { return (v + 15) & ~int64_t(15); }

__host__ __device__ inline int64_t ws_per_file_bytes(int64_t stride) {
    // S1[stride+1], S2[stride+1], T[stride+1], delta[stride+1] doubles + two bit-mask word arrays
    return align16(4 * (stride + 1) * 8) + 2 * align16((stride / 32 + 2) * 4);
}

// Inclusive block scan of a pair of doubles; returns exclusive prefix for this
// thread and the block total (both pairs).  `sh` = shared double[2][9].
template <int THREADS>
__device__ __forceinline__ void block_excl_scan2(double v1, double v2, double& ex1, double& ex2, double& tot1,
                                                 double& tot2, double (*sh)[9]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double i1 = v1, i2 = v2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double a = __shfl_up_sync(0xffffffffu, i1, o);
        double b = __shfl_up_sync(0xffffffffu, i2, o);
        if (lane >= o) {
            i1 += a;
            i2 += b;
        }
    }
    __syncthreads();
    if (lane == 31) {
        sh[0][warp] = i1;
        sh[1][warp] = i2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < THREADS / 32; ++w) {
            double ta = sh[0][w], tb = sh[1][w];
            sh[0][w] = a;
            sh[1][w] = b;
            a += ta;
            b += tb;
        }
        sh[0][8] = a;
        sh[1][8] = b;
    }
    __syncthreads();
    ex1 = sh[0][warp] + (i1 - v1);
    ex2 = sh[1][warp] + (i2 - v2);
    tot1 = sh[0][8];
    tot2 = sh[1][8];
}

template <int THREADS>
__device__ __forceinline__ void block_excl_scan2i(int v1, int v2, int& ex1, int& ex2, int& tot1, int& tot2,
                                                  int (*sh)[9]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int i1 = v1, i2 = v2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int a = __shfl_up_sync(0xffffffffu, i1, o);
        int b = __shfl_up_sync(0xffffffffu, i2, o);
        if (lane >= o) {
            i1 += a;
            i2 += b;
        }
    }
    __syncthreads();
    if (lane == 31) {
        sh[0][warp] = i1;
        sh[1][warp] = i2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = 0, b = 0;
        for (int w = 0; w < THREADS / 32; ++w) {
            int ta = sh[0][w], tb = sh[1][w];
            sh[0][w] = a;
            sh[1][w] = b;
            a += ta;
            b += tb;
        }
        sh[0][8] = a;
        sh[1][8] = b;
    }
    __syncthreads();
    ex1 = sh[0][warp] + (i1 - v1);
    ex2 = sh[1][warp] + (i2 - v2);
    tot1 = sh[0][8];
    tot2 = sh[1][8];
}

// Python's timedelta(seconds=float) -> integer microseconds: the integral part
// is exact, the fractional part is rounded half-to-even after *1e6
// (CPython Modules/_datetimemodule.c accum()).
__device__ __forceinline__ int64_t py_seconds_to_us(double sec) {
    double ip;
    const double fp = modf(sec, &ip);
    return (int64_t)ip * 1000000ll + (int64_t)rint(fp * 1e6);
}

// A-hour + Kritisch for one event (dsp/src/main.py:690-696; detector_and_classification.py:50).
__device__ __forceinline__ void hourly_add(int start, int stop, int64_t file_start_us, double bd, double crit_min,
                                           int64_t hour0, int n_hours, int32_t* out_hist) {
    const double t_start = start * bd, t_stop = stop * bd;  // main.py:424-425, 503-504
    const double dur = t_stop - t_start;                    // main.py:426, 505
    const int64_t us = file_start_us + py_seconds_to_us(t_start);  // utc_start, main.py:432, 510
    int64_t hour = us / 3600000000ll;                       // floor division (times before the epoch are legal)
    if (us % 3600000000ll < 0) --hour;
    const int64_t h = hour - hour0;
    if (h < 0 || h >= n_hours) return;
    atomicAdd(&out_hist[h * 2 + 0], 1);                        // Anzahl
    if (dur >= crit_min) atomicAdd(&out_hist[h * 2 + 1], 1);   // Kritisch
}

template <bool ADAPTIVE, int THREADS>
__device__ __forceinline__ void detect_file(const DetectParams& p, const int f) {
    constexpr int kThreads = THREADS;   // shadows the namespace default inside this function
    __shared__ double shd[2][9];
    __shared__ int shi[2][9];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t n64 = p.n_blocks_per_file ? (int64_t)p.n_blocks_per_file[f] : p.n_blocks;
    if (n64 > p.stride) n64 = p.stride;
    if (n64 < 0) n64 = 0;
    const int N = (int)n64;
    const float* band = p.band + (int64_t)f * p.stride;
    const float* noise = p.noise + (int64_t)f * p.stride;
    char* ws = p.workspace + (int64_t)f * p.ws_per_file;
    extern __shared__ __align__(16) double dyn_smem[];
    // per-block arrays: shared memory for ordinary files (the freeze scan below is latency bound),
    // the global workspace for very long recordings
    double* arr = p.use_smem ? dyn_smem : reinterpret_cast<double*>(ws);
    const int64_t astride = p.use_smem ? (int64_t)(kSmemBlocks + 1) : (p.stride + 1);
    double* S1 = arr;
    double* S2 = S1 + astride;
    double* T = S2 + astride;
    double* dl = T + astride;
    __shared__ uint32_t sUw[kSmemBlocks / 32 + 1], sDw[kSmemBlocks / 32 + 1];
    uint32_t* gwords = reinterpret_cast<uint32_t*>(ws + align16(4 * (p.stride + 1) * 8));
    uint32_t* Dw = p.use_smem ? sDw : gwords;                                  // detected blocks, 1 bit each
    uint32_t* Uw = p.use_smem ? sUw : gwords + align16((p.stride / 32 + 2) * 4) / 4;  // unfrozen-hypothesis mask
    const int words = (N + 31) / 32;

    if (N == 0) {
        if (tid == 0) p.out_counts[f] = 0;
        return;
    }
    auto delta = [&](int i) -> double { return dl[i]; };

    // ---- one pass: delta = band - noise (main.py:393) and exclusive prefix sums of c = delta - s0 and c^2.
    // Windowed mean/variance are shift invariant, so any fixed shift s0 works; the scan totals also give the
    // whole-file mean and population std (main.py:399-400, 464-466) without separate reduction passes.
    const double s0 = (double)band[0] - (double)noise[0];
    if (tid == 0) {
        S1[0] = 0.0;
        S2[0] = 0.0;
    }
    double carry1 = 0.0, carry2 = 0.0;
    for (int base = 0; base < N; base += kThreads * kItems) {
        const int i0 = base + tid * kItems;
        float bv[kItems], nv[kItems];
#pragma unroll
        for (int j = 0; j < kItems; ++j) {   // independent loads first: their latencies overlap
            const bool ok = i0 + j < N;
            bv[j] = ok ? band[i0 + j] : 0.0f;
            nv[j] = ok ? noise[i0 + j] : 0.0f;
        }
        double c[kItems];
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            const double d = (double)bv[j] - (double)nv[j];
            const bool ok = i0 + j < N;
            if (ok) dl[i0 + j] = d;
            c[j] = ok ? d - s0 : 0.0;
            t1 += c[j];
            t2 += c[j] * c[j];
        }
        double ex1, ex2, tot1, tot2;
        block_excl_scan2<THREADS>(t1, t2, ex1, ex2, tot1, tot2, shd);
        double r1 = carry1 + ex1, r2 = carry2 + ex2;
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            r1 += c[j];
            r2 += c[j] * c[j];
            if (i0 + j < N) {
                S1[i0 + j + 1] = r1;
                S2[i0 + j + 1] = r2;
            }
        }
        carry1 += tot1;
        carry2 += tot2;
    }
    const double m0 = carry1 / (double)N;
    const double mean = s0 + m0;
    double var = carry2 / (double)N - m0 * m0;
    var = var > 0.0 ? var : 0.0;
    const double g = mean + p.k_std * sqrt(var);
    __syncthreads();

    const unsigned full = 0xffffffffu;
    if (ADAPTIVE) {
        // ---- all blocks in parallel: candidate threshold of the trailing window [max(0,i-W), i)
        // (main.py:475-482) and U = "detects if the estimator is not frozen" (delta > candidate) ----
        for (int i = tid; i < words * 32; i += kThreads) {
            bool u = false;
            if (i < N) {
                double t;
                if (i < p.fixed) {
                    t = g;
                } else {
                    const int w0 = max(0, i - p.window);
                    const int cnt = i - w0;
                    if (cnt <= 0) {
                        t = nan("");  // np.mean of an empty slice
                    } else {
                        const double m = (S1[i] - S1[w0]) / (double)cnt;
                        double v = (S2[i] - S2[w0]) / (double)cnt - m * m;
                        v = (v > 0.0 && cnt > 1) ? v : 0.0;  // a one-sample window has std == 0 exactly
                        t = (s0 + m) + p.k_std * sqrt(v);
                    }
                }
                T[i] = t;
                u = delta(i) > t;
            }
            const unsigned m = __ballot_sync(full, u);
            if (lane == 0) {
                Uw[i >> 5] = m;
                Dw[i >> 5] = 0u;
            }
        }
        __syncthreads();

        // ---- the data-dependent freeze logic (main.py:470-493), one warp, a few steps per burst:
        //  * unfrozen: the next detection is the next set bit of U -> jump there with word operations;
        //  * frozen (after a detection the threshold is held for `after` blocks): test the whole stretch
        //    against the held threshold with independent ballots, extend while it keeps detecting. ----
        if (warp == 0) {
            const bool want_thr = (p.out_thresholds != nullptr) || (p.out_near != nullptr);
            auto emit_thr = [&](int a, int b_incl, bool use_T, double Hval) {
                if (!want_thr) return;
                for (int q = a + lane; q <= b_incl; q += 32) {
                    const double th = use_T ? T[q] : Hval;
                    if (p.out_thresholds) p.out_thresholds[(int64_t)f * p.stride + q] = th;
                    if (p.out_near) p.out_near[(int64_t)f * p.stride + q] = (fabs(delta(q) - th) < p.eps_db) ? 1 : 0;
                }
            };
            auto new_freeze = [&](int q) { return max(q + p.after, max(0, q - p.before)); };  // main.py:491-493
            int F = -1;    // freeze_until_idx
            double H = g;  // threshold carried while frozen
            // fixed initial period: threshold = g whatever the freeze state (main.py:471-472)
            const int nf = min(p.fixed, N);
            int last_hit = -1;
            for (int w = lane; w * 32 < nf; w += 32) {
                unsigned m = Uw[w];
                const int hi = nf - w * 32;
                if (hi < 32) m &= (1u << hi) - 1u;
                Dw[w] = m;
                if (m) last_hit = max(last_hit, w * 32 + 31 - __clz(m));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) last_hit = max(last_hit, __shfl_xor_sync(full, last_hit, o));
            if (last_hit >= 0) F = new_freeze(last_hit);
            __syncwarp();
            emit_thr(0, nf - 1, true, 0.0);
            int cur = nf;
            while (cur < N) {
                if (cur > F) {
                    int q = -1;
                    for (int wb = cur >> 5; wb < words; wb += 32) {
                        const int w = wb + lane;
                        unsigned m = (w < words) ? Uw[w] : 0u;
                        if (w == (cur >> 5)) m &= ~0u << (cur & 31);
                        const unsigned bal = __ballot_sync(full, m != 0u);
                        if (bal) {
                            const int L = __ffs(bal) - 1;
                            const unsigned mm = __shfl_sync(full, m, L);
                            q = (wb + L) * 32 + __ffs(mm) - 1;
                            break;
                        }
                    }
                    if (q < 0) {  // no further detection: the rest of the file runs unfrozen
                        emit_thr(cur, N - 1, true, 0.0);
                        break;
                    }
                    emit_thr(cur, q, true, 0.0);
                    if (lane == 0) Dw[q >> 5] |= 1u << (q & 31);
                    H = T[q];
                    F = new_freeze(q);
                    cur = q + 1;
                } else {
                    const int end = min(F, N - 1);
                    int lastq = -1;
#pragma unroll 4
                    for (int w = cur >> 5; w <= (end >> 5); ++w) {
                        const int pos = w * 32 + lane;
                        const bool hit = pos >= cur && pos <= end && delta(pos) > H;
                        const unsigned m = __ballot_sync(full, hit);
                        if (m) {
                            if (lane == 0) Dw[w] |= m;
                            lastq = w * 32 + 31 - __clz(m);
                        }
                    }
                    emit_thr(cur, end, false, H);
                    if (lastq >= 0) F = new_freeze(lastq);
                    cur = end + 1;
                }
            }
        }
    } else {
        // ---- global threshold mask (main.py:405) ----
        for (int w = warp; w < words; w += kThreads / 32) {
            const int pos = w * 32 + lane;
            const bool valid = pos < N;
            const double d = valid ? delta(pos) : 0.0;
            const unsigned m = __ballot_sync(full, valid && (d > g));
            if (lane == 0) Dw[w] = m;
            if (valid && p.out_near) p.out_near[(int64_t)f * p.stride + pos] = (fabs(d - g) < p.eps_db) ? 1 : 0;
        }
        if (tid == 0 && p.out_thresholds) p.out_thresholds[(int64_t)f * p.stride] = g;
    }
    __syncthreads();

    // ---- runs of detected blocks -> ordered (start, stop) pairs ----
    int carry_s = 0, carry_e = 0;
    for (int wbase = 0; wbase < words; wbase += kThreads) {
        const int w = wbase + tid;
        unsigned d = 0, prev_msb = 0, next_lsb = 0;
        if (w < words) {
            d = Dw[w];
            if (w > 0) prev_msb = Dw[w - 1] >> 31;
            if (w + 1 < words) next_lsb = Dw[w + 1] & 1u;
        }
        unsigned starts = d & ~((d << 1) | prev_msb);
        unsigned ends = d & ~((d >> 1) | (next_lsb << 31));
        int exs, exe, tots, tote;
        block_excl_scan2i<THREADS>(__popc(starts), __popc(ends), exs, exe, tots, tote, shi);
        int is = carry_s + exs, ie = carry_e + exe;
        while (starts) {
            const int b = __ffs(starts) - 1;
            starts &= starts - 1;
            if (is < p.max_events) p.out_events[((int64_t)f * p.max_events + is) * 2 + 0] = w * 32 + b;
            ++is;
        }
        while (ends) {
            const int b = __ffs(ends) - 1;
            ends &= ends - 1;
            const int last = w * 32 + b;  // last detected block of the run
            int stop = last + 1;
            if (!ADAPTIVE && last == N - 1) stop = N - 1;  // open-at-EOF quirk, main.py:414-415
            if (ie < p.max_events) p.out_events[((int64_t)f * p.max_events + ie) * 2 + 1] = stop;
            ++ie;
        }
        carry_s += tots;
        carry_e += tote;
    }
    if (tid == 0) p.out_counts[f] = carry_s;
    __syncthreads();

    // ---- per-event mean dB = mean(delta[start:stop]) (main.py:422-423, 501-502) ----
    const int n_ev = min(carry_s, p.max_events);
    for (int e = tid; e < n_ev; e += kThreads) {
        const int64_t o = (int64_t)f * p.max_events + e;
        const int start = p.out_events[o * 2 + 0], stop = p.out_events[o * 2 + 1];
        p.out_event_db[o] = (stop > start) ? (S1[stop] - S1[start]) / (double)(stop - start) + s0 : nan("");
        if (p.out_hist) hourly_add(start, stop, p.file_start_us[f], p.block_duration_sec, p.crit_min_dur_sec, p.hour0,
                                   p.n_hours, p.out_hist);
    }
}

template <bool ADAPTIVE, int THREADS>
__global__ void __launch_bounds__(THREADS) detect_kernel(DetectParams p) {
    // programmatic dependent launch: this grid may be scheduled while the band-power kernel drains;
    // its results must not be read before that grid has completed and flushed
    if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    // one CTA per file normally; the small-footprint launch uses at most one CTA per SM and loops
    for (int f = blockIdx.x; f < p.n_files; f += gridDim.x) {
        detect_file<ADAPTIVE, THREADS>(p, f);
        __syncthreads();
    }
}

__global__ void hourly_kernel(const int32_t* __restrict__ events, const int32_t* __restrict__ counts, int64_t n_files,
                              int max_events, const int64_t* __restrict__ file_start_us, double bd, double crit_min,
                              int64_t hour0, int n_hours, int32_t* out_hist) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_files * (int64_t)max_events) return;
    const int64_t f = idx / max_events;
    const int e = (int)(idx % max_events);
    if (e >= counts[f]) return;
    hourly_add(events[idx * 2 + 0], events[idx * 2 + 1], file_start_us[f], bd, crit_min, hour0, n_hours, out_hist);
}

int launch_detect(bool adaptive, const float* band_db, const float* noise_db, int64_t n_files, int64_t stride,
                  int64_t n_blocks, const int32_t* n_blocks_per_file, double k_std, int32_t window, int32_t before,
                  int32_t after, int32_t fixed, int32_t max_events, int32_t* out_events, double* out_event_db,
                  int32_t* out_counts, double* out_thresholds, uint8_t* out_near, double eps_db, void* workspace,
                  int64_t workspace_bytes, void* stream, const int64_t* file_start_us = nullptr,
                  double block_duration_sec = 0.0, double crit_min_dur_sec = 0.0, int64_t hour0 = 0,
                  int32_t n_hours = 0, int32_t* out_hist = nullptr, bool pdl = false, bool small = false) {
    MS_REQUIRE(band_db && noise_db && out_events && out_event_db && out_counts, MS_ERR_INVALID_ARG,
               "ms_detect: null pointer argument");
    MS_REQUIRE(n_files >= 0 && stride >= 0 && n_blocks >= 0 && n_blocks <= stride, MS_ERR_INVALID_ARG,
               "ms_detect: bad sizes n_files=%lld stride=%lld n_blocks=%lld", (long long)n_files, (long long)stride,
               (long long)n_blocks);
    MS_REQUIRE(stride < (int64_t)1 << 30, MS_ERR_UNSUPPORTED, "ms_detect: more than 2^30 blocks per file");
    MS_REQUIRE(n_files < (int64_t)1 << 31, MS_ERR_UNSUPPORTED, "ms_detect: too many files");
    MS_REQUIRE(max_events > 0, MS_ERR_INVALID_ARG, "ms_detect: max_events must be positive");
    MS_REQUIRE(workspace && workspace_bytes >= ms_detect_workspace_bytes(n_files, stride), MS_ERR_WORKSPACE,
               "ms_detect: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
               (long long)ms_detect_workspace_bytes(n_files, stride));
    if (n_files == 0) return MS_OK;
    DetectParams p;
    p.band = band_db;
    p.noise = noise_db;
    p.n_blocks_per_file = n_blocks_per_file;
    p.stride = stride;
    p.n_blocks = n_blocks;
    p.k_std = k_std;
    p.window = window;
    p.before = before;
    p.after = after;
    p.fixed = fixed;
    p.max_events = max_events;
    p.out_events = out_events;
    p.out_event_db = out_event_db;
    p.out_counts = out_counts;
    p.out_thresholds = out_thresholds;
    p.out_near = out_near;
    p.eps_db = eps_db;
    p.workspace = static_cast<char*>(workspace);
    p.ws_per_file = ws_per_file_bytes(stride);
    p.file_start_us = file_start_us;
    p.block_duration_sec = block_duration_sec;
    p.crit_min_dur_sec = crit_min_dur_sec;
    p.hour0 = hour0;
    p.n_hours = n_hours;
    p.out_hist = out_hist;
    if (out_hist) MS_REQUIRE(file_start_us && n_hours > 0, MS_ERR_INVALID_ARG, "ms_detect: hourly stage needs file_start_us and n_hours");
    p.use_smem = (stride <= kSmemBlocks && !small) ? 1 : 0;
    const size_t smem = p.use_smem ? kSmemBytes : 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    p.pdl = pdl ? 1 : 0;
    p.n_files = (int32_t)n_files;
    if (adaptive && small) {
        // 128 threads, no dynamic shared memory, per-block arrays in the (L2-resident) workspace: fits next to
        // a persistent band-power CTA on the same SM, so it can run under the next batch's STFT
        int64_t grid = num_sms();
        if (grid < 1) grid = 1;
        if (grid > n_files) grid = n_files;
        // same shared-memory carveout as the band-power kernel: CTAs of kernels that prefer different L1/shared
        // splits do not share an SM, and this launch exists to run beside those CTAs
        MS_CUDA_OK(cudaFuncSetAttribute(detect_kernel<true, kThreadsSmall>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        (int)cudaSharedmemCarveoutMaxShared));
        detect_kernel<true, kThreadsSmall><<<(unsigned)grid, kThreadsSmall, 0, st>>>(p);
    } else if (adaptive && pdl) {
        if (smem) MS_CUDA_OK(cudaFuncSetAttribute(detect_kernel<true, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_files);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MS_CUDA_OK(cudaLaunchKernelEx(&cfg, detect_kernel<true, kThreads>, p));
    } else if (adaptive) {
        if (smem) MS_CUDA_OK(cudaFuncSetAttribute(detect_kernel<true, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        detect_kernel<true, kThreads><<<(unsigned)n_files, kThreads, smem, st>>>(p);
    } else {
        if (smem) MS_CUDA_OK(cudaFuncSetAttribute(detect_kernel<false, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        detect_kernel<false, kThreads><<<(unsigned)n_files, kThreads, smem, st>>>(p);
    }
    MS_CUDA_OK(cudaGetLastError());
    return MS_OK;
}

}  // namespace

// used by the one-call pass (ms_pipeline.cu): adaptive detect + hourly, launched as a programmatic dependent
int detect_adaptive_hourly_pdl(const float* band_db, const float* noise_db, int64_t n_files, int64_t n_blocks,
                               double k_std, int32_t window, int32_t before, int32_t after, int32_t fixed,
                               int32_t max_events, int32_t* out_events, double* out_event_db, int32_t* out_counts,
                               void* workspace, int64_t workspace_bytes, const int64_t* file_start_us,
                               double block_duration_sec, double crit_min_dur_sec, int64_t hour0, int32_t n_hours,
                               int32_t* out_hist, void* stream) {
    return launch_detect(true, band_db, noise_db, n_files, n_blocks, n_blocks, nullptr, k_std, window, before, after,
                         fixed, max_events, out_events, out_event_db, out_counts, nullptr, nullptr, 0.0, workspace,
                         workspace_bytes, stream, file_start_us, block_duration_sec, crit_min_dur_sec, hour0, n_hours,
                         out_hist, true);
}
}  // namespace ms

extern "C" {

int64_t ms_detect_workspace_bytes(int64_t n_files, int64_t stride) {
    if (n_files < 0 || stride < 0) return 0;
    return n_files * ms::ws_per_file_bytes(stride) + 256;
}

int ms_detect_global(const float* band_db, const float* noise_db, int64_t n_files, int64_t stride, int64_t n_blocks,
                     const int32_t* n_blocks_per_file, double k_std, int32_t max_events, int32_t* out_events,
                     double* out_event_db, int32_t* out_counts, double* out_thresholds, uint8_t* out_near,
                     double eps_db, void* workspace, int64_t workspace_bytes, void* stream) {
    return ms::launch_detect(false, band_db, noise_db, n_files, stride, n_blocks, n_blocks_per_file, k_std, 0, 0, 0, 0,
                             max_events, out_events, out_event_db, out_counts, out_thresholds, out_near, eps_db,
                             workspace, workspace_bytes, stream);
}

int ms_detect_adaptive(const float* band_db, const float* noise_db, int64_t n_files, int64_t stride, int64_t n_blocks,
                       const int32_t* n_blocks_per_file, double k_std, int32_t window_blocks,
                       int32_t freeze_before_blocks, int32_t freeze_after_blocks, int32_t fixed_blocks,
                       int32_t max_events, int32_t* out_events, double* out_event_db, int32_t* out_counts,
                       double* out_thresholds, uint8_t* out_near, double eps_db, void* workspace,
                       int64_t workspace_bytes, void* stream) {
    MS_REQUIRE(window_blocks >= 0 && fixed_blocks >= 0, MS_ERR_INVALID_ARG,
               "ms_detect_adaptive: negative window/fixed block count");
    return ms::launch_detect(true, band_db, noise_db, n_files, stride, n_blocks, n_blocks_per_file, k_std,
                             window_blocks, freeze_before_blocks, freeze_after_blocks, fixed_blocks, max_events,
                             out_events, out_event_db, out_counts, out_thresholds, out_near, eps_db, workspace,
                             workspace_bytes, stream);
}

int ms_detect_adaptive_hourly(const float* band_db, const float* noise_db, int64_t n_files, int64_t stride,
                              int64_t n_blocks, const int32_t* n_blocks_per_file, double k_std, int32_t window_blocks,
                              int32_t freeze_before_blocks, int32_t freeze_after_blocks, int32_t fixed_blocks,
                              int32_t max_events, int32_t* out_events, double* out_event_db, int32_t* out_counts,
                              double* out_thresholds, uint8_t* out_near, double eps_db, void* workspace,
                              int64_t workspace_bytes, const int64_t* file_start_us, double block_duration_sec,
                              double crit_min_dur_sec, int64_t hour0, int32_t n_hours, int32_t* out_hist,
                              uint32_t flags, void* stream) {
    MS_REQUIRE(window_blocks >= 0 && fixed_blocks >= 0, MS_ERR_INVALID_ARG,
               "ms_detect_adaptive_hourly: negative window/fixed block count");
    MS_REQUIRE(out_hist && file_start_us, MS_ERR_INVALID_ARG, "ms_detect_adaptive_hourly: null histogram arguments");
    return ms::launch_detect(true, band_db, noise_db, n_files, stride, n_blocks, n_blocks_per_file, k_std,
                             window_blocks, freeze_before_blocks, freeze_after_blocks, fixed_blocks, max_events,
                             out_events, out_event_db, out_counts, out_thresholds, out_near, eps_db, workspace,
                             workspace_bytes, stream, file_start_us, block_duration_sec, crit_min_dur_sec, hour0,
                             n_hours, out_hist, false, (flags & MS_DETECT_SMALL_FOOTPRINT) != 0);
}

int ms_hourly_counts(const int32_t* events, const int32_t* counts, int64_t n_files, int32_t max_events,
                     const int64_t* file_start_us, double block_duration_sec, double crit_min_dur_sec, int64_t hour0,
                     int32_t n_hours, int32_t* out_hist, void* stream) {
    MS_REQUIRE(events && counts && file_start_us && out_hist, MS_ERR_INVALID_ARG, "ms_hourly_counts: null pointer");
    MS_REQUIRE(n_files >= 0 && max_events > 0 && n_hours > 0, MS_ERR_INVALID_ARG, "ms_hourly_counts: bad sizes");
    if (n_files == 0) return MS_OK;
    const int64_t total = n_files * (int64_t)max_events;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    ms::hourly_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        events, counts, n_files, max_events, file_start_us, block_duration_sec, crit_min_dur_sec, hour0, n_hours,
        out_hist);
    MS_CUDA_OK(cudaGetLastError());
    return MS_OK;
}

}  // extern "C"
