// K1: framing + window + real FFT fused with |X|^2, band-bin selection and dB,
// so only band-limited results reach HBM.
//
// The real frame is packed as a half-length complex sequence z[m] = x[2m] + i x[2m+1]
// and transformed with a mixed-radix (8,8,..,[4|2]) Stockham FFT: every thread keeps
// one radix-8 butterfly (8 complex values) in registers per pass, passes exchange
// data through shared memory as float2 (64-bit accesses) padded by one element per 8
// to avoid bank conflicts, the transform size is a template parameter so all padded
// indices fold to constants, and several frames share a CTA so a barrier is amortised
// over 4-16 frames.  The first pass reads the samples straight from global memory
// (one 32-bit load per PCM16 pair converted with a mantissa trick instead of I2F,
// 64-bit per float pair) and applies mean removal and the window.
//
// One kernel template serves three reference call sites:
//   MODE_BAND  dsp/src/main.py:376-388        np.fft.rfft(block*np.hanning, n) -> 2 band sums -> dB
//   MODE_WELCH dsp/src/live/backend/processor.py:206, 349-367, 393
//              scipy.signal.welch (5 mean-removed Hann segments) -> 3 band sums -> dB, db2
//   MODE_PSD   meteor_detect_class/prime_detection.py:67-92 / dsp/src/main.py:52-54
//              one-sided PSD rows k_lo..k_hi + noise-band sum over time and frequency
//
// This is the general path (any power-of-two nfft, hop, band, float or PCM16).  For
// PCM16 input with a narrow band the tensor-core kernel in ms_dft_i8.cu is the fast path.
#include "ms_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace ms {
namespace {

enum { MODE_BAND = 0, MODE_WELCH = 1, MODE_PSD = 2 };
constexpr int kK1Threads = 256;
constexpr int kRowAccMax = 1024;   // waterfall bins (all frame slots of a CTA together)

struct StftParams {
    const void* x;
    int64_t n_outer;        // files / streams / segments
    int64_t outer_stride;   // samples between outer units
    int64_t n_frames;       // frames (blocks) per outer unit
    int32_t hop;            // samples between frames
    int32_t n_sub;          // sub-segments per frame (Welch), 1 otherwise
    int32_t sub_hop;
    int32_t win_len;        // samples entering the transform (<= nfft)
    const float* window;    // [win_len]
    int32_t log2_nc;        // nfft = 2 << log2_nc
    int32_t detrend;        // remove the mean of the win_len samples first
    float in_scale;         // multiply samples (1/32768 for soundfile-style PCM16)
    int32_t band_lo[3], band_hi[3];
    double scale;           // PSD scale 1/(fs*sum w^2) (WELCH / PSD modes)
    int64_t out_stride;
    float* out0;            // BAND: band dB     WELCH: out_db[.][4]   PSD: out_psd
    float* out1;            // BAND: noise dB
    float* out2;            // BAND: band energy (optional)
    float* out3;            // BAND: noise energy (optional)
    double* out_noise_sum;  // PSD
    float* out_rows;        // WELCH (optional): per-bin PSD in dB for bins row_lo..row_hi, [outer][frame][n_rows]
    int32_t row_lo, row_hi;
    int32_t pair_loads;     // 1: every frame starts on an even sample and x is pair aligned
};

// shared-memory index padding: one float2 per 16 (= one 128-byte bank row) keeps a half-warp's
// consecutive 64-bit accesses gap free and the stride-8 Stockham writes of the first pass conflict free
__device__ __forceinline__ int padi(int i) { return i + (i >> 4); }

__device__ __forceinline__ float load_sample(const int16_t* x, int64_t i) { return (float)x[i]; }
__device__ __forceinline__ float load_sample(const float* x, int64_t i) { return x[i]; }
// PCM16 pair -> two floats without the slow I2F path: (s ^ 0x8000) placed in the mantissa of 2^23
// gives 8388608 + 32768 + s exactly, and subtracting that constant is exact too.
constexpr float kI16Bias = 8421376.0f;
__device__ __forceinline__ void load_pair(const int16_t* x, int64_t i, float& a, float& b) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(x + i) ^ 0x80008000u;
    a = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)) - kI16Bias;
    b = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)) - kI16Bias;
}
__device__ __forceinline__ void load_pair(const float* x, int64_t i, float& a, float& b) {
    const float2 w = *reinterpret_cast<const float2*>(x + i);
    a = w.x;
    b = w.y;
}

// ---- register butterflies (forward transform, natural-order outputs) ----
__device__ __forceinline__ void bfly2(float2* v) {
    const float2 a = v[0], b = v[1];
    v[0] = make_float2(a.x + b.x, a.y + b.y);
    v[1] = make_float2(a.x - b.x, a.y - b.y);
}
__device__ __forceinline__ void bfly4(float2* v) {
    const float t0r = v[0].x + v[2].x, t0i = v[0].y + v[2].y, t1r = v[0].x - v[2].x, t1i = v[0].y - v[2].y;
    const float t2r = v[1].x + v[3].x, t2i = v[1].y + v[3].y;
    const float t3r = v[1].y - v[3].y, t3i = -(v[1].x - v[3].x);   // -i * (v1 - v3)
    v[0] = make_float2(t0r + t2r, t0i + t2i);
    v[1] = make_float2(t1r + t3r, t1i + t3i);
    v[2] = make_float2(t0r - t2r, t0i - t2i);
    v[3] = make_float2(t1r - t3r, t1i - t3i);
}
__device__ __forceinline__ void bfly8(float2* v) {
    const float S = 0.70710678118654752440f;
    const float a0r = v[0].x + v[4].x, a0i = v[0].y + v[4].y, a4r = v[0].x - v[4].x, a4i = v[0].y - v[4].y;
    const float a1r = v[1].x + v[5].x, a1i = v[1].y + v[5].y;
    float a5r = v[1].x - v[5].x, a5i = v[1].y - v[5].y;
    const float a2r = v[2].x + v[6].x, a2i = v[2].y + v[6].y;
    float a6r = v[2].x - v[6].x, a6i = v[2].y - v[6].y;
    const float a3r = v[3].x + v[7].x, a3i = v[3].y + v[7].y;
    float a7r = v[3].x - v[7].x, a7i = v[3].y - v[7].y;
    float t;
    t = S * (a5r + a5i); a5i = S * (a5i - a5r); a5r = t;          // * (1 - i)/sqrt2
    t = a6i; a6i = -a6r; a6r = t;                                  // * (-i)
    t = S * (a7i - a7r); a7i = -S * (a7r + a7i); a7r = t;          // * (-1 - i)/sqrt2
    const float b0r = a0r + a2r, b0i = a0i + a2i, b2r = a0r - a2r, b2i = a0i - a2i;
    const float b1r = a1r + a3r, b1i = a1i + a3i;
    const float b3r = a1i - a3i, b3i = -(a1r - a3r);               // -i * (a1 - a3)
    const float b4r = a4r + a6r, b4i = a4i + a6i, b6r = a4r - a6r, b6i = a4i - a6i;
    const float b5r = a5r + a7r, b5i = a5i + a7i;
    const float b7r = a5i - a7i, b7i = -(a5r - a7r);               // -i * (a5 - a7)
    v[0] = make_float2(b0r + b1r, b0i + b1i);
    v[1] = make_float2(b4r + b5r, b4i + b5i);
    v[2] = make_float2(b2r + b3r, b2i + b3i);
    v[3] = make_float2(b6r + b7r, b6i + b7i);
    v[4] = make_float2(b0r - b1r, b0i - b1i);
    v[5] = make_float2(b4r - b5r, b4i - b5i);
    v[6] = make_float2(b2r - b3r, b2i - b3i);
    v[7] = make_float2(b6r - b7r, b6i - b7i);
}
template <int R>
__device__ __forceinline__ void bfly(float2* v) {
    if (R == 8) bfly8(v);
    if (R == 4) bfly4(v);
    if (R == 2) bfly2(v);
}

// One Stockham pass of radix R over a frame held in a padded float2 shared array.
// NC, NS and TPF are compile-time so every padded index folds to base + constant.
// FIRST: inputs come from global memory (mean removed, windowed, packed pairs).
template <int R, int NC, int NS, int TPF, bool FIRST, typename T>
__device__ __forceinline__ void fft_pass(const StftParams& p, const T* __restrict__ x, int64_t base, float mean,
                                         const float2* __restrict__ src, float2* __restrict__ dst,
                                         const float2* __restrict__ tw, int t) {
    constexpr int Q = NC / R;
#pragma unroll
    for (int b0 = 0; b0 < Q; b0 += TPF) {
        const int b = b0 + t;
        float2 v[R];
        const int k = b & (NS - 1);
        if (FIRST) {
            if (p.pair_loads && p.win_len >= 2 * NC) {
                // whole frame in range: issue every load first (no per-element branches), then convert
                float a[R], c[R];
                float2 w[R];
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const int i0 = 2 * (b + j * Q);
                    load_pair(x, base + i0, a[j], c[j]);
                    w[j] = __ldg(reinterpret_cast<const float2*>(p.window + i0));
                }
#pragma unroll
                for (int j = 0; j < R; ++j)
                    v[j] = make_float2(fmaf(a[j], p.in_scale, -mean) * w[j].x, fmaf(c[j], p.in_scale, -mean) * w[j].y);
            } else {
#pragma unroll
                for (int j = 0; j < R; ++j) {   // zero-padded or unaligned frames
                    const int i0 = 2 * (b + j * Q);
                    float re = 0.0f, im = 0.0f;
                    if (i0 < p.win_len) re = (load_sample(x, base + i0) * p.in_scale - mean) * p.window[i0];
                    if (i0 + 1 < p.win_len) im = (load_sample(x, base + i0 + 1) * p.in_scale - mean) * p.window[i0 + 1];
                    v[j] = make_float2(re, im);
                }
            }
        } else {
            // this pass's twiddles exp(-2 pi i j k / (R NS)) are stored [k][j-1]: odd stride, conflict free
            const float2* twk = tw + k * (R - 1);
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const float2 a = src[padi(b + j * Q)];
                if (j == 0) {
                    v[0] = a;
                } else {
                    const float2 w = twk[j - 1];
                    v[j] = make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
                }
            }
        }
        bfly<R>(v);
        const int o = (b - k) * R + k;
#pragma unroll
        for (int j = 0; j < R; ++j) dst[padi(o + j * NS)] = v[j];
    }
}

// All passes of an NC-point transform: radix 8 while possible, then one radix-4 or radix-2 pass.
template <int NC, int NS, int TPF, typename T>
__device__ __forceinline__ const float2* fft_all(const StftParams& p, const T* x, int64_t base, float mean, bool active,
                                                 float2* a, float2* b, const float2* tw, int t) {
    constexpr int REM = NC / NS;
    constexpr int R = REM >= 8 ? 8 : REM;
    if (active) fft_pass<R, NC, NS, TPF, NS == 1, T>(p, x, base, mean, a, b, tw, t);
    __syncthreads();
    if constexpr (NS * R < NC) {
        // the next pass's table follows this one's (the first pass has none)
        return fft_all<NC, NS * R, TPF, T>(p, x, base, mean, active, b, a, tw + (NS == 1 ? 0 : NS * (R - 1)), t);
    } else {
        return b;
    }
}

// |X[k]|^2 of the length-2NC real transform from the NC-point packed transform Z.
__device__ __forceinline__ float real_bin_power(const float2* Z, int k, int NC) {
    const float2 zk = Z[padi(k & (NC - 1))];
    const float2 zn = Z[padi((NC - k) & (NC - 1))];
    // E = (Zk + conj(Zn))/2 ; O = -i (Zk - conj(Zn))/2 ; X = E + w^k O, w = exp(-i pi / NC)
    const float ex = 0.5f * (zk.x + zn.x), ey = 0.5f * (zk.y - zn.y);
    const float dx = 0.5f * (zk.x - zn.x), dy = 0.5f * (zk.y + zn.y);
    const float ox = dy, oy = -dx;
    float s, c;
    sincospif(-(float)k / (float)NC, &s, &c);
    const float xr = ex + (c * ox - s * oy);
    const float xi = ey + (c * oy + s * ox);
    return xr * xr + xi * xi;
}

template <typename T, int MODE, int LOG2NC>
__global__ void __launch_bounds__(kK1Threads) stft_kernel(StftParams p) {
    extern __shared__ __align__(16) float2 smem_f2[];
    __shared__ float red[kK1Threads / 32][3];
    __shared__ float row_acc[kRowAccMax];   // WELCH waterfall rows: per-bin power summed over the sub-segments
    constexpr int NC = 1 << LOG2NC;
    constexpr int PN = NC + (NC >> 4) + 1;
    constexpr int TPF = (NC / 8 < kK1Threads) ? NC / 8 : kK1Threads;   // threads per frame
    constexpr int FR = kK1Threads / TPF;                               // frames handled concurrently by a CTA
    float2* tw = smem_f2;                            // per-pass twiddle tables (< NC entries in total)
    float2* bufs = tw + NC;                          // FR x 2 x PN

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = tid / TPF, t = tid - slot * TPF;
    {   // per-pass twiddle tables, pass with sub-transform length NS and radix R: [k < NS][j = 1..R-1]
        int off = 0;
        for (int ns = 8; ns < NC; ns *= 8) {      // passes after the first; radix 8 while ns*8 <= NC
            const int rem = NC / ns;
            const int R = rem >= 8 ? 8 : rem;
            for (int e = tid; e < ns * (R - 1); e += kK1Threads) {
                const int k = e / (R - 1), j = e - k * (R - 1) + 1;
                float sn, cs;
                sincospif(-2.0f * (float)(j * k) / (float)(R * ns), &sn, &cs);
                tw[off + e] = make_float2(cs, sn);
            }
            off += ns * (R - 1);
        }
    }
    __syncthreads();

    float2* buf_a = bufs + (size_t)slot * 2 * PN;
    float2* buf_b = buf_a + PN;

    // sum over the threads of one frame slot, result broadcast to all of them
    auto slot_sum3 = [&](float& v0, float& v1, float& v2) {
        constexpr int W = TPF < 32 ? TPF : 32;
#pragma unroll
        for (int o = W >> 1; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
            v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        }
        if (TPF > 32) {
            __syncthreads();
            if (lane == 0) {
                red[warp][0] = v0;
                red[warp][1] = v1;
                red[warp][2] = v2;
            }
            __syncthreads();
            constexpr int WPS = TPF > 32 ? TPF / 32 : 1;
            v0 = v1 = v2 = 0.0f;
#pragma unroll
            for (int ww = 0; ww < WPS; ++ww) {
                v0 += red[slot * WPS + ww][0];
                v1 += red[slot * WPS + ww][1];
                v2 += red[slot * WPS + ww][2];
            }
        }
    };

    const T* x = static_cast<const T*>(p.x);
    const int64_t total = p.n_outer * p.n_frames;
    const int n_bands = (MODE == MODE_WELCH) ? 3 : 2;
    const int nfft_half = NC;  // Nyquist bin index

    for (int64_t u0 = (int64_t)blockIdx.x * FR; u0 < total; u0 += (int64_t)gridDim.x * FR) {
        const int64_t u = u0 + slot;
        const bool active = u < total;
        int64_t outer = 0, frame = 0;
        if (active) {
            if (total < 0x7fffffffll) {   // 32-bit division is several times cheaper than the 64-bit one
                outer = (uint32_t)u / (uint32_t)p.n_frames;
                frame = (uint32_t)u - (uint32_t)outer * (uint32_t)p.n_frames;
            } else {
                outer = u / p.n_frames;
                frame = u - outer * p.n_frames;
            }
        }
        float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;
        const int n_rows_out = (MODE == MODE_WELCH && p.out_rows) ? (p.row_hi - p.row_lo + 1) : 0;
        if (n_rows_out > 0) {
            for (int i = t; i < n_rows_out; i += TPF) row_acc[slot * n_rows_out + i] = 0.0f;
        }
        for (int sub = 0; sub < p.n_sub; ++sub) {
            const int64_t base = outer * p.outer_stride + frame * (int64_t)p.hop + (int64_t)sub * p.sub_hop;
            float mean = 0.0f;
            if (p.detrend) {
                float s = 0.0f, d1 = 0.0f, d2 = 0.0f;
                if (active)
                    for (int i = t; i < p.win_len; i += TPF) s += load_sample(x, base + i) * p.in_scale;
                slot_sum3(s, d1, d2);
                mean = s / (float)p.win_len;
            }
            const float2* Z = fft_all<NC, 1, TPF, T>(p, x, base, mean, active, buf_b, buf_a, tw, t);

            if (MODE == MODE_PSD && active) {
                const int nb = p.band_hi[0] - p.band_lo[0] + 1;
                float* out = p.out0 + (outer * (int64_t)nb) * p.n_frames + frame;
                for (int i = t; i < nb; i += TPF) {
                    const int k = p.band_lo[0] + i;
                    float pw = real_bin_power(Z, k, NC) * (float)p.scale;
                    if (k != 0 && k != nfft_half) pw *= 2.0f;
                    out[(int64_t)i * p.n_frames] = pw;
                }
            }
            if (n_rows_out > 0 && active) {   // waterfall rows: each bin is owned by one thread, no races
                for (int i = t; i < n_rows_out; i += TPF) {
                    const int k = p.row_lo + i;
                    float pw = real_bin_power(Z, k, NC);
                    if (k != 0 && k != nfft_half) pw *= 2.0f;
                    row_acc[slot * n_rows_out + i] += pw;
                }
            }
            // band sums (deterministic order: per-thread -> shuffle tree -> fixed-order partials)
            float part[3] = {0.0f, 0.0f, 0.0f};
            if (active) {
                const int b_first = (MODE == MODE_PSD) ? 1 : 0;
                for (int b = b_first; b < n_bands; ++b) {
                    const int lo = p.band_lo[b], hi = p.band_hi[b];
                    for (int k = lo + t; k <= hi; k += TPF) {
                        float pw = real_bin_power(Z, k, NC);
                        if (MODE != MODE_BAND && k != 0 && k != nfft_half) pw *= 2.0f;
                        part[b] += pw;
                    }
                }
            }
            slot_sum3(part[0], part[1], part[2]);
            acc0 += part[0];
            acc1 += part[1];
            acc2 += part[2];
            __syncthreads();   // everyone is done with this transform before the buffers are reused
        }
        if (n_rows_out > 0 && active) {
            const float sc = (float)(p.scale / (double)p.n_sub);
            float* o = p.out_rows + (outer * p.n_frames + frame) * (int64_t)n_rows_out;
            for (int i = t; i < n_rows_out; i += TPF) {
                const float pw = row_acc[slot * n_rows_out + i] * sc;
                o[i] = pw > 0.0f ? 10.0f * log10f(pw) : -INFINITY;     // processor.py:207 block_psd_db
            }
        }
        if (active && t == 0) {
            if (MODE == MODE_BAND) {
                const int64_t o = outer * p.out_stride + frame;
                p.out0[o] = 10.0f * log10f(acc0 + 1e-12f);   // main.py:383-384
                p.out1[o] = 10.0f * log10f(acc1 + 1e-12f);   // main.py:387-388
                if (p.out2) p.out2[o] = acc0;
                if (p.out3) p.out3[o] = acc1;
            } else if (MODE == MODE_WELCH) {
                const float sc = (float)(p.scale / (double)p.n_sub);
                const float a[3] = {acc0, acc1, acc2};
                float db[3];
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    const float pw = a[b] * sc;
                    db[b] = pw > 0.0f ? 10.0f * log10f(pw) : -INFINITY;   // processor.py:352
                }
                float* o = p.out0 + (outer * p.n_frames + frame) * 4;
                o[0] = db[0];
                o[1] = db[1];
                o[2] = db[2];
                o[3] = db[0] - 0.5f * (db[1] + db[2]);   // processor.py:393
            } else {
                atomicAdd(&p.out_noise_sum[outer], (double)acc1 * p.scale);   // prime_detection.py:83
            }
        }
    }
}

}  // namespace
}  // namespace ms
#include "ms_fft_warp.cuh"
namespace ms {
namespace {

int log2_exact(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return ((1 << l) == v) ? l : -1;
}

template <typename T, int MODE, int LOG2NC>
int launch_stft_sized(StftParams& p, cudaStream_t st) {
    constexpr int NC = 1 << LOG2NC;
    constexpr int PN = NC + (NC >> 4) + 1;
    constexpr int TPF = (NC / 8 < kK1Threads) ? NC / 8 : kK1Threads;
    constexpr int FR = kK1Threads / TPF;
    const size_t smem = sizeof(float2) * ((size_t)NC + (size_t)FR * 2 * PN);
    auto kern = stft_kernel<T, MODE, LOG2NC>;
    // static + dynamic shared memory above 48 KiB needs the opt-in
    if (smem > 40 * 1024) MS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kK1Threads, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t total = p.n_outer * p.n_frames;
    if (total == 0) return MS_OK;
    const int64_t groups = (total + FR - 1) / FR;
    int64_t grid = (int64_t)num_sms() * per_sm;
    if (grid > groups) grid = groups;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kK1Threads, smem, st>>>(p);
    MS_CUDA_OK(cudaGetLastError());
    return MS_OK;
}

template <typename T, int MODE>
int launch_stft(StftParams& p, int nfft, cudaStream_t st) {
    const int l2 = log2_exact(nfft);
    MS_REQUIRE(l2 >= 8 && l2 <= 14, MS_ERR_UNSUPPORTED, "nfft=%d must be a power of two in [256, 16384]", nfft);
    p.log2_nc = l2 - 1;
    const size_t elem = sizeof(T);
    p.pair_loads = ((reinterpret_cast<uintptr_t>(p.x) % (2 * elem)) == 0 && p.outer_stride % 2 == 0 && p.hop % 2 == 0 &&
                    p.sub_hop % 2 == 0 && (reinterpret_cast<uintptr_t>(p.window) % 8) == 0)
                       ? 1
                       : 0;
    switch (p.log2_nc) {
        case 7: return launch_stft_sized<T, MODE, 7>(p, st);
        case 8: return launch_stft_sized<T, MODE, 8>(p, st);
        case 9: return launch_stft_sized<T, MODE, 9>(p, st);
        case 10: return launch_stft_sized<T, MODE, 10>(p, st);
        case 11: return launch_stft_sized<T, MODE, 11>(p, st);
        case 12: return launch_stft_sized<T, MODE, 12>(p, st);
        default: return launch_stft_sized<T, MODE, 13>(p, st);
    }
}

template <typename T>
int band_power(const T* x, int64_t n_files, int64_t file_stride, int64_t n_frames, int32_t hop, int32_t win_len,
               const float* window, int32_t nfft, int32_t k_sig_lo, int32_t k_sig_hi, int32_t k_noise_lo,
               int32_t k_noise_hi, int64_t out_stride, float* out_band_db, float* out_noise_db,
               float* out_band_energy, float* out_noise_energy, void* stream) {
    MS_REQUIRE(x && window && out_band_db && out_noise_db, MS_ERR_INVALID_ARG, "ms_band_power: null pointer");
    MS_REQUIRE(n_files >= 0 && n_frames >= 0 && hop > 0 && win_len > 0 && win_len <= nfft, MS_ERR_INVALID_ARG,
               "ms_band_power: bad geometry (hop=%d win_len=%d nfft=%d)", hop, win_len, nfft);
    MS_REQUIRE(n_frames == 0 || (n_frames - 1) * (int64_t)hop + win_len <= file_stride, MS_ERR_INVALID_ARG,
               "ms_band_power: frames exceed file_stride");
    MS_REQUIRE(out_stride >= n_frames, MS_ERR_INVALID_ARG, "ms_band_power: out_stride < n_frames");
    const int nb = nfft / 2;
    MS_REQUIRE(k_sig_lo >= 0 && k_sig_hi <= nb && k_noise_lo >= 0 && k_noise_hi <= nb, MS_ERR_INVALID_ARG,
               "ms_band_power: bin range outside [0, nfft/2]");
    StftParams p = {};
    p.x = x;
    p.n_outer = n_files;
    p.outer_stride = file_stride;
    p.n_frames = n_frames;
    p.hop = hop;
    p.n_sub = 1;
    p.sub_hop = 0;
    p.win_len = win_len;
    p.window = window;
    p.detrend = 0;
    p.in_scale = 1.0f;
    p.band_lo[0] = k_sig_lo;
    p.band_hi[0] = k_sig_hi;     // an empty mask (lo > hi) sums to 0 like numpy
    p.band_lo[1] = k_noise_lo;
    p.band_hi[1] = k_noise_hi;
    p.band_lo[2] = 1;
    p.band_hi[2] = 0;
    p.scale = 1.0;
    p.out_stride = out_stride;
    p.out0 = out_band_db;
    p.out1 = out_noise_db;
    p.out2 = out_band_energy;
    p.out3 = out_noise_energy;
    return launch_stft<T, MODE_BAND>(p, nfft, static_cast<cudaStream_t>(stream));
}

template <typename T>
int welch_band_db(const T* x, float in_scale, int64_t n_streams, int64_t stream_stride, int64_t n_blocks,
                  int32_t block, int32_t nperseg, const float* window, int32_t nfft, const int32_t* h_bands,
                  double scale, float* out_db, int32_t row_lo, int32_t row_hi, float* out_rows, void* stream) {
    MS_REQUIRE(x && window && h_bands && out_db, MS_ERR_INVALID_ARG, "ms_welch_band_db: null pointer");
    MS_REQUIRE(block > 0 && nperseg > 1 && nperseg <= block && nperseg <= nfft, MS_ERR_INVALID_ARG,
               "ms_welch_band_db: need 1 < nperseg <= block and nperseg <= nfft");
    MS_REQUIRE(n_blocks == 0 || n_blocks * (int64_t)block <= stream_stride, MS_ERR_INVALID_ARG,
               "ms_welch_band_db: blocks exceed stream_stride");
    const int hop = nperseg - nperseg / 2;
    StftParams p = {};
    p.x = x;
    p.n_outer = n_streams;
    p.outer_stride = stream_stride;
    p.n_frames = n_blocks;
    p.hop = block;
    p.n_sub = (block - nperseg / 2) / hop;   // scipy: (n - noverlap) // step
    p.sub_hop = hop;
    p.win_len = nperseg;
    p.window = window;
    p.detrend = 1;
    p.in_scale = in_scale;
    for (int b = 0; b < 3; ++b) {
        p.band_lo[b] = h_bands[2 * b];
        p.band_hi[b] = h_bands[2 * b + 1];
        MS_REQUIRE(p.band_lo[b] >= 0 && p.band_hi[b] <= nfft / 2, MS_ERR_INVALID_ARG,
                   "ms_welch_band_db: bin range outside [0, nfft/2]");
    }
    p.scale = scale;
    p.out0 = out_db;
    if (out_rows) {
        const int frames_per_cta = (nfft / 16 < kK1Threads) ? kK1Threads / (nfft / 16) : 1;
        MS_REQUIRE(row_lo >= 0 && row_hi >= row_lo && row_hi <= nfft / 2 &&
                       (row_hi - row_lo + 1) * frames_per_cta <= kRowAccMax,
                   MS_ERR_INVALID_ARG, "ms_welch_band_db: waterfall bin range [%d, %d] invalid or wider than %d bins",
                   row_lo, row_hi, kRowAccMax / frames_per_cta);
        p.out_rows = out_rows;
        p.row_lo = row_lo;
        p.row_hi = row_hi;
    }
    return launch_stft<T, MODE_WELCH>(p, nfft, static_cast<cudaStream_t>(stream));
}

template <typename T>
int psd_spectrogram(const T* x, int64_t n_segments, int64_t seg_stride, int64_t n_frames, int32_t hop, int32_t nfft,
                    const float* window, double scale, int32_t k_lo, int32_t k_hi, int32_t k_noise_lo,
                    int32_t k_noise_hi, float* out_psd, double* out_noise_sum, void* stream) {
    MS_REQUIRE(x && window && out_psd && out_noise_sum, MS_ERR_INVALID_ARG, "ms_psd_spectrogram: null pointer");
    MS_REQUIRE(hop > 0 && (n_frames == 0 || (n_frames - 1) * (int64_t)hop + nfft <= seg_stride), MS_ERR_INVALID_ARG,
               "ms_psd_spectrogram: frames exceed seg_stride");
    MS_REQUIRE(k_lo >= 0 && k_hi <= nfft / 2 && k_lo <= k_hi && k_noise_lo >= 0 && k_noise_hi <= nfft / 2,
               MS_ERR_INVALID_ARG, "ms_psd_spectrogram: bad bin range");
    {   // nfft 2048 on pair-aligned data below the Nyquist bin: the warp-per-frame kernel (ms_fft_warp.cuh)
        static const bool force_k1 = [] {
            const char* e = getenv("MS_PSD_IMPL");
            return e && strcmp(e, "fft") == 0;
        }();
        const bool aligned = reinterpret_cast<uintptr_t>(x) % (2 * sizeof(T)) == 0 && seg_stride % 2 == 0 &&
                             hop % 2 == 0 && reinterpret_cast<uintptr_t>(window) % 8 == 0;
        if (!force_k1 && nfft == 2048 && aligned && k_hi < 1024 && (k_noise_lo > k_noise_hi || k_noise_hi < 1024)) {
            PsdWarpParams w = {};
            w.x = x;
            w.n_outer = n_segments;
            w.outer_stride = seg_stride;
            w.n_frames = n_frames;
            w.hop = hop;
            w.window = window;
            w.k_lo = k_lo;
            w.k_hi = k_hi;
            w.n_lo = k_noise_lo;
            w.n_hi = k_noise_hi;
            w.scale = scale;
            w.out = out_psd;
            w.out_noise = out_noise_sum;
            return launch_psd_warp<T>(w, static_cast<cudaStream_t>(stream));
        }
    }
    StftParams p = {};
    p.x = x;
    p.n_outer = n_segments;
    p.outer_stride = seg_stride;
    p.n_frames = n_frames;
    p.hop = hop;
    p.n_sub = 1;
    p.win_len = nfft;
    p.window = window;
    p.detrend = 0;
    p.in_scale = 1.0f;
    p.band_lo[0] = k_lo;
    p.band_hi[0] = k_hi;
    p.band_lo[1] = k_noise_lo;
    p.band_hi[1] = k_noise_hi;
    p.band_lo[2] = 1;
    p.band_hi[2] = 0;
    p.scale = scale;
    p.out0 = out_psd;
    p.out_noise_sum = out_noise_sum;
    return launch_stft<T, MODE_PSD>(p, nfft, static_cast<cudaStream_t>(stream));
}

}  // namespace
}  // namespace ms

extern "C" {

int ms_band_power_i16(const int16_t* x, int64_t n_files, int64_t file_stride, int64_t n_frames, int32_t hop,
                      int32_t win_len, const float* window, int32_t nfft, int32_t k_sig_lo, int32_t k_sig_hi,
                      int32_t k_noise_lo, int32_t k_noise_hi, int64_t out_stride, float* out_band_db,
                      float* out_noise_db, float* out_band_energy, float* out_noise_energy, void* stream) {
    return ms::band_power<int16_t>(x, n_files, file_stride, n_frames, hop, win_len, window, nfft, k_sig_lo, k_sig_hi,
                                   k_noise_lo, k_noise_hi, out_stride, out_band_db, out_noise_db, out_band_energy,
                                   out_noise_energy, stream);
}

int ms_band_power_f32(const float* x, int64_t n_files, int64_t file_stride, int64_t n_frames, int32_t hop,
                      int32_t win_len, const float* window, int32_t nfft, int32_t k_sig_lo, int32_t k_sig_hi,
                      int32_t k_noise_lo, int32_t k_noise_hi, int64_t out_stride, float* out_band_db,
                      float* out_noise_db, float* out_band_energy, float* out_noise_energy, void* stream) {
    return ms::band_power<float>(x, n_files, file_stride, n_frames, hop, win_len, window, nfft, k_sig_lo, k_sig_hi,
                                 k_noise_lo, k_noise_hi, out_stride, out_band_db, out_noise_db, out_band_energy,
                                 out_noise_energy, stream);
}

int ms_welch_band_db_f32(const float* x, int64_t n_streams, int64_t stream_stride, int64_t n_blocks, int32_t block,
                         int32_t nperseg, const float* window, int32_t nfft, const int32_t* h_bands, double scale,
                         float* out_db, int32_t row_lo, int32_t row_hi, float* out_rows, void* stream) {
    return ms::welch_band_db<float>(x, 1.0f, n_streams, stream_stride, n_blocks, block, nperseg, window, nfft,
                                    h_bands, scale, out_db, row_lo, row_hi, out_rows, stream);
}

int ms_welch_band_db_i16(const int16_t* x, int64_t n_streams, int64_t stream_stride, int64_t n_blocks, int32_t block,
                         int32_t nperseg, const float* window, int32_t nfft, const int32_t* h_bands, double scale,
                         float* out_db, int32_t row_lo, int32_t row_hi, float* out_rows, void* stream) {
    return ms::welch_band_db<int16_t>(x, 1.0f / 32768.0f, n_streams, stream_stride, n_blocks, block, nperseg, window,
                                      nfft, h_bands, scale, out_db, row_lo, row_hi, out_rows, stream);
}

int ms_psd_spectrogram_i16(const int16_t* x, int64_t n_segments, int64_t seg_stride, int64_t n_frames, int32_t hop,
                           int32_t nfft, const float* window, double scale, int32_t k_lo, int32_t k_hi,
                           int32_t k_noise_lo, int32_t k_noise_hi, float* out_psd, double* out_noise_sum,
                           void* stream) {
    return ms::psd_spectrogram<int16_t>(x, n_segments, seg_stride, n_frames, hop, nfft, window, scale, k_lo, k_hi,
                                        k_noise_lo, k_noise_hi, out_psd, out_noise_sum, stream);
}

int ms_psd_spectrogram_f32(const float* x, int64_t n_segments, int64_t seg_stride, int64_t n_frames, int32_t hop,
                           int32_t nfft, const float* window, double scale, int32_t k_lo, int32_t k_hi,
                           int32_t k_noise_lo, int32_t k_noise_hi, float* out_psd, double* out_noise_sum,
                           void* stream) {
    return ms::psd_spectrogram<float>(x, n_segments, seg_stride, n_frames, hop, nfft, window, scale, k_lo, k_hi,
                                      k_noise_lo, k_noise_hi, out_psd, out_noise_sum, stream);
}

}  // extern "C"
