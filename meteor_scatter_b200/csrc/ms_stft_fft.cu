// K1: framing + window + real FFT (Stockham radix-4 in shared memory, real
// input packed as a half-length complex transform) fused with |X|^2, band-bin
// selection and dB, so only band-limited results reach HBM.
//
// One kernel template serves three reference call sites:
//   MODE_BAND  dsp/src/main.py:376-388        np.fft.rfft(block*np.hanning, n) -> 2 band sums -> dB
//   MODE_WELCH dsp/src/live/backend/processor.py:206, 349-367, 393
//              scipy.signal.welch (5 mean-removed Hann segments) -> 3 band sums -> dB, db2
//   MODE_PSD   meteor_detect_class/prime_detection.py:67-92 / dsp/src/main.py:52-54
//              one-sided PSD rows k_lo..k_hi + noise-band sum over time and frequency
//
// This is the general path (any power-of-two nfft, hop, band).  For PCM16 input
// with a narrow band the tensor-core kernel in ms_dft_i8.cu is the fast path.
#include "ms_common.cuh"

namespace ms {
namespace {

enum { MODE_BAND = 0, MODE_WELCH = 1, MODE_PSD = 2 };

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
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ float load_sample(const int16_t* x, int64_t i) { return (float)x[i]; }
__device__ __forceinline__ float load_sample(const float* x, int64_t i) { return x[i]; }

// In-place-of-pair Stockham autosort FFT of NC complex points held in shared
// memory; returns the buffer holding the natural-order result.
__device__ float2* fft_stockham(float2* a, float2* b, const float2* tw, int log2_nc) {
    const int NC = 1 << log2_nc;
    const int tid = threadIdx.x, nth = blockDim.x;
    float2* src = a;
    float2* dst = b;
    int ns = 1;
    int rem = log2_nc;
    while (rem >= 2) {
        const int q = NC >> 2;
        const int tmul = NC / (4 * ns);
        for (int j = tid; j < q; j += nth) {
            const int k = j & (ns - 1);
            const int ts = k * tmul;
            const float2 v0 = src[j];
            const float2 v1 = cmul(src[j + q], tw[ts]);
            const float2 v2 = cmul(src[j + 2 * q], tw[2 * ts]);
            const float2 v3 = cmul(src[j + 3 * q], tw[3 * ts]);
            const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y);
            const float2 t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
            const float2 t3 = make_float2(v1.y - v3.y, -(v1.x - v3.x));  // -i * (v1 - v3)
            const int idx = ((j - k) << 2) + k;
            dst[idx] = make_float2(t0.x + t2.x, t0.y + t2.y);
            dst[idx + ns] = make_float2(t1.x + t3.x, t1.y + t3.y);
            dst[idx + 2 * ns] = make_float2(t0.x - t2.x, t0.y - t2.y);
            dst[idx + 3 * ns] = make_float2(t1.x - t3.x, t1.y - t3.y);
        }
        __syncthreads();
        float2* t = src;
        src = dst;
        dst = t;
        ns <<= 2;
        rem -= 2;
    }
    if (rem == 1) {
        const int h = NC >> 1;
        const int tmul = NC / (2 * ns);
        for (int j = tid; j < h; j += nth) {
            const int k = j & (ns - 1);
            const float2 v0 = src[j];
            const float2 v1 = cmul(src[j + h], tw[k * tmul]);
            const int idx = ((j - k) << 1) + k;
            dst[idx] = make_float2(v0.x + v1.x, v0.y + v1.y);
            dst[idx + ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
        }
        __syncthreads();
        float2* t = src;
        src = dst;
        dst = t;
    }
    return src;
}

// |X[k]|^2 of the length-2NC real transform from the NC-point packed transform Z.
__device__ __forceinline__ float real_bin_power(const float2* Z, int k, int NC) {
    const float2 zk = Z[k & (NC - 1)];
    const float2 zn = Z[(NC - k) & (NC - 1)];
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

template <typename T, int MODE>
__global__ void __launch_bounds__(256) stft_kernel(StftParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NC = 1 << p.log2_nc;
    float2* bufa = reinterpret_cast<float2*>(smem_raw);
    float2* bufb = bufa + NC;
    float2* tw = bufb + NC;               // NC entries (3/4 used)
    __shared__ float red[8][3];
    __shared__ float sh_mean;
    __shared__ float acc[3];

    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5;
    for (int m = tid; m < NC; m += nth) {
        float s, c;
        sincospif(-2.0f * (float)m / (float)NC, &s, &c);
        tw[m] = make_float2(c, s);
    }
    __syncthreads();

    const T* x = static_cast<const T*>(p.x);
    const int64_t total = p.n_outer * p.n_frames;
    const int n_bands = (MODE == MODE_WELCH) ? 3 : 2;
    const int nfft_half = NC;  // Nyquist bin index

    for (int64_t u = blockIdx.x; u < total; u += gridDim.x) {
        const int64_t outer = u / p.n_frames, frame = u % p.n_frames;
        if (tid < 3) acc[tid] = 0.0f;
        for (int sub = 0; sub < p.n_sub; ++sub) {
            const int64_t base = outer * p.outer_stride + frame * (int64_t)p.hop + (int64_t)sub * p.sub_hop;
            float mean = 0.0f;
            if (p.detrend) {
                float s = 0.0f;
                for (int i = tid; i < p.win_len; i += nth) s += load_sample(x, base + i) * p.in_scale;
                s = warp_sum(s);
                __syncthreads();
                if (lane == 0) red[warp][0] = s;
                __syncthreads();
                if (tid == 0) {
                    float t = 0.0f;
                    for (int w = 0; w < (nth + 31) / 32; ++w) t += red[w][0];
                    sh_mean = t / (float)p.win_len;
                }
                __syncthreads();
                mean = sh_mean;
            }
            // pack: z[m] = w[2m] x[2m] + i w[2m+1] x[2m+1], zero padded (np.fft.rfft(n=nfft))
            for (int m = tid; m < NC; m += nth) {
                const int i0 = 2 * m, i1 = 2 * m + 1;
                float re = 0.0f, im = 0.0f;
                if (i0 < p.win_len) re = (load_sample(x, base + i0) * p.in_scale - mean) * p.window[i0];
                if (i1 < p.win_len) im = (load_sample(x, base + i1) * p.in_scale - mean) * p.window[i1];
                bufa[m] = make_float2(re, im);
            }
            __syncthreads();
            const float2* Z = fft_stockham(bufa, bufb, tw, p.log2_nc);

            if (MODE == MODE_PSD) {
                const int nb = p.band_hi[0] - p.band_lo[0] + 1;
                float* out = p.out0 + (outer * (int64_t)nb) * p.n_frames + frame;
                for (int i = tid; i < nb; i += nth) {
                    const int k = p.band_lo[0] + i;
                    float pw = real_bin_power(Z, k, NC) * (float)p.scale;
                    if (k != 0 && k != nfft_half) pw *= 2.0f;
                    out[(int64_t)i * p.n_frames] = pw;
                }
            }
            // band sums (deterministic: per-thread -> warp shuffle -> fixed-order smem sum)
            float part[3] = {0.0f, 0.0f, 0.0f};
            const int b_first = (MODE == MODE_PSD) ? 1 : 0;
            for (int b = b_first; b < n_bands; ++b) {
                const int lo = p.band_lo[b], hi = p.band_hi[b];
                for (int k = lo + tid; k <= hi; k += nth) {
                    float pw = real_bin_power(Z, k, NC);
                    if (MODE != MODE_BAND && k != 0 && k != nfft_half) pw *= 2.0f;
                    part[b] += pw;
                }
            }
#pragma unroll
            for (int b = 0; b < 3; ++b) part[b] = warp_sum(part[b]);
            __syncthreads();  // everyone is done reading Z before the next sub-segment overwrites it
            if (lane == 0) {
                red[warp][0] = part[0];
                red[warp][1] = part[1];
                red[warp][2] = part[2];
            }
            __syncthreads();
            if (tid < 3) {
                float t = 0.0f;
                for (int w = 0; w < (nth + 31) / 32; ++w) t += red[w][tid];
                acc[tid] += t;
            }
            __syncthreads();
        }
        if (tid == 0) {
            if (MODE == MODE_BAND) {
                const float eb = acc[0], en = acc[1];
                const int64_t o = outer * p.out_stride + frame;
                p.out0[o] = 10.0f * log10f(eb + 1e-12f);   // main.py:383-384
                p.out1[o] = 10.0f * log10f(en + 1e-12f);   // main.py:387-388
                if (p.out2) p.out2[o] = eb;
                if (p.out3) p.out3[o] = en;
            } else if (MODE == MODE_WELCH) {
                const float sc = (float)(p.scale / (double)p.n_sub);
                float db[3];
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    const float pw = acc[b] * sc;
                    db[b] = pw > 0.0f ? 10.0f * log10f(pw) : -INFINITY;   // processor.py:352
                }
                float* o = p.out0 + (outer * p.n_frames + frame) * 4;
                o[0] = db[0];
                o[1] = db[1];
                o[2] = db[2];
                o[3] = db[0] - 0.5f * (db[1] + db[2]);   // processor.py:393
            } else {
                atomicAdd(&p.out_noise_sum[outer], (double)acc[1] * p.scale);   // prime_detection.py:83
            }
        }
        __syncthreads();
    }
}

int log2_exact(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return ((1 << l) == v) ? l : -1;
}

template <typename T, int MODE>
int launch_stft(StftParams& p, int nfft, cudaStream_t st) {
    const int l2 = log2_exact(nfft);
    MS_REQUIRE(l2 >= 8 && l2 <= 14, MS_ERR_UNSUPPORTED, "nfft=%d must be a power of two in [256, 16384]", nfft);
    p.log2_nc = l2 - 1;
    const int NC = nfft / 2;
    const size_t smem = (size_t)NC * sizeof(float2) * 3;
    int threads = NC / 4;
    if (threads > 256) threads = 256;
    if (threads < 64) threads = 64;
    auto kern = stft_kernel<T, MODE>;
    // static + dynamic shared memory above 48 KiB needs the opt-in (nfft = 4096 is exactly 48 KiB dynamic)
    if (smem > 40 * 1024) MS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t total = p.n_outer * p.n_frames;
    if (total == 0) return MS_OK;
    int64_t grid = (int64_t)num_sms() * per_sm;
    if (grid > total) grid = total;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, threads, smem, st>>>(p);
    MS_CUDA_OK(cudaGetLastError());
    return MS_OK;
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
                  double scale, float* out_db, void* stream) {
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
                         float* out_db, void* stream) {
    return ms::welch_band_db<float>(x, 1.0f, n_streams, stream_stride, n_blocks, block, nperseg, window, nfft,
                                    h_bands, scale, out_db, stream);
}

int ms_welch_band_db_i16(const int16_t* x, int64_t n_streams, int64_t stream_stride, int64_t n_blocks, int32_t block,
                         int32_t nperseg, const float* window, int32_t nfft, const int32_t* h_bands, double scale,
                         float* out_db, void* stream) {
    return ms::welch_band_db<int16_t>(x, 1.0f / 32768.0f, n_streams, stream_stride, n_blocks, block, nperseg, window,
                                      nfft, h_bands, scale, out_db, stream);
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
