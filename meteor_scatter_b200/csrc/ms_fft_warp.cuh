// K1W: one warp per frame, 2048-point real transform as a 32 x 32 complex transform held in registers.
//
// Included by ms_stft_fft.cu (uses its bfly4/bfly8/load_pair).  It serves the PSD mode of detector C
// (meteor_detect_class/prime_detection.py:67-92: specgram NFFT 2048, noverlap 1024) and detector A's
// per-event spectrogram crops (dsp/src/main.py:52-54) when nfft == 2048.
//
// Why another FFT kernel: the block-cooperative K1 spends 3650 warp instructions per 2048-sample frame,
// a third of them in the per-bin real-split epilogue (sincospif per bin) and another third in index
// arithmetic, shared-memory traffic and block barriers around four Stockham passes.  Here a frame is
// packed as z[m] = x[2m] + i x[2m+1] (1024 complex points) and factorised 1024 = 32 x 32:
//   stage 1  lane n2 holds z[32 n1 + n2], n1 = 0..31, and runs a 32-point transform over n1 in registers;
//   twiddle  Y[k1][n2] *= exp(-2 pi i n2 k1 / 1024), table in shared memory ([k1][n2], conflict free);
//   exchange one 32 x 32 transpose through a per-warp padded buffer (row pitch 33 float2), __syncwarp only;
//   stage 2  lane k1 runs the 32-point transform over n2: register k2 holds Z[k1 + 32 k2];
//   epilogue X[k] = E + w^k O needs Z[k] and Z[1024 - k]: the partner sits in lane (32 - k1) & 31,
//            register 31 - k2 (lane 0: (32 - k2) & 31), one shuffle pair per 32 bins; w^k = base(lane) * step(k2).
// No block barrier after the tables are built, 128 shared wavefronts of exchange per frame instead of ~650.
#pragma once

namespace ms {
namespace {

constexpr int kPwWarps = 20;   // 20 warps x 96 registers: the most an SM holds with the 64-register working set
constexpr int kPwThreads = kPwWarps * 32;
constexpr int kPwPitch = 33;   // float2 per transposed row

struct PsdWarpParams {
    const void* x;
    int64_t n_outer, outer_stride, n_frames;
    int32_t hop;
    const float* window;   // [2048]
    int32_t k_lo, k_hi;    // PSD rows
    int32_t n_lo, n_hi;    // noise band (empty when n_lo > n_hi)
    double scale;
    float* out;            // [outer][k_hi-k_lo+1][n_frames]
    double* out_noise;     // [outer]
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// a * exp(-2 pi i m / 32); m is a compile-time constant after unrolling, so the switch folds away
__device__ __forceinline__ float2 mul_w32(float2 a, int m) {
    constexpr float S = 0.70710678118654752440f;
    constexpr float C1 = 0.98078528040323044913f, S1 = 0.19509032201612826785f;   // pi/16
    constexpr float C2 = 0.92387953251128675613f, S2 = 0.38268343236508977173f;   // pi/8
    constexpr float C3 = 0.83146961230254523708f, S3 = 0.55557023301960222474f;   // 3 pi/16
    float c, s;   // exp(-i phi) = c - i s
    switch (m) {
        case 0: return a;
        case 4: return make_float2(S * (a.x + a.y), S * (a.y - a.x));
        case 8: return make_float2(a.y, -a.x);
        case 12: return make_float2(S * (a.y - a.x), -S * (a.x + a.y));
        case 16: return make_float2(-a.x, -a.y);
        case 1: c = C1; s = S1; break;
        case 2: c = C2; s = S2; break;
        case 3: c = C3; s = S3; break;
        case 5: c = S3; s = C3; break;
        case 6: c = S2; s = C2; break;
        case 7: c = S1; s = C1; break;
        case 9: c = -S1; s = C1; break;
        case 10: c = -S2; s = C2; break;
        case 14: c = -C2; s = S2; break;
        case 15: c = -C1; s = S1; break;
        case 18: c = -C2; s = -S2; break;
        case 21: c = -S3; s = -C3; break;
        default: c = 1.0f; s = 0.0f; break;   // not reached: m = b*c with b < 8, c < 4
    }
    return make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
}

// forward 32-point transform in registers, natural order in and out: 32 = 4 (a) x 8 (b), n = 8a + b, k = c + 4d
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        float2 t[4] = {v[b], v[8 + b], v[16 + b], v[24 + b]};
        bfly4(t);
#pragma unroll
        for (int c = 0; c < 4; ++c) v[8 * c + b] = mul_w32(t[c], b * c);
    }
    float2 r[32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float2 u[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) u[b] = v[8 * c + b];
        bfly8(u);
#pragma unroll
        for (int d = 0; d < 8; ++d) r[c + 4 * d] = u[d];
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = r[i];
}

// packed samples of one lane's 32 pairs as they come from memory (converted after the previous frame's epilogue)
template <typename T> struct RawPair;
template <> struct RawPair<int16_t> { typedef uint32_t type; };
template <> struct RawPair<float> { typedef float2 type; };
__device__ __forceinline__ float2 raw_to_f2(uint32_t w) {
    w ^= 0x80008000u;   // see load_pair: (s ^ 0x8000) in the mantissa of 2^23, minus the bias, is exact
    return make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)) - kI16Bias,
                       __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)) - kI16Bias);
}
__device__ __forceinline__ float2 raw_to_f2(float2 w) { return w; }

// K2MAX: number of 32-bin groups the epilogue is unrolled over (16 when every wanted bin is below 512)
template <typename T, int K2MAX, bool PF>
__global__ void __launch_bounds__(kPwThreads, 1) psd_warp_kernel(PsdWarpParams p) {
    typedef typename RawPair<T>::type Raw;
    constexpr bool kPrefetch = PF && sizeof(T) == 2;   // 32 spare registers exist for PCM16 words, not for 64 floats
    extern __shared__ __align__(16) float2 pw_smem[];
    float2* tw = pw_smem;              // [k1][n2] exp(-2 pi i n2 k1 / 1024)
    float2* win = tw + 1024;           // window pairs (w[2m], w[2m+1])
    float2* rtw = win + 1024;          // [k2][lane] exp(-i pi (lane + 32 k2) / 1024), the real-split twiddle of bin k
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* buf = rtw + 32 * K2MAX + warp * (32 * kPwPitch);

    for (int e = tid; e < 1024; e += kPwThreads) {
        float sn, cs;
        sincospif(-2.0f * (float)((e >> 5) * (e & 31)) / 1024.0f, &sn, &cs);
        tw[e] = make_float2(cs, sn);
        win[e] = __ldg(reinterpret_cast<const float2*>(p.window) + e);
        if (e < 32 * K2MAX) {
            sincospif(-(float)e / 1024.0f, &sn, &cs);
            rtw[e] = make_float2(cs, sn);
        }
    }
    __syncthreads();

    const T* x = static_cast<const T*>(p.x);
    const int64_t total = p.n_outer * p.n_frames;
    const uint32_t nb1 = (uint32_t)(p.k_hi - p.k_lo);              // rows - 1
    const bool have_noise = p.n_lo <= p.n_hi;
    const uint32_t nn1 = have_noise ? (uint32_t)(p.n_hi - p.n_lo) : 0u;
    const int kmin = have_noise ? min(p.k_lo, p.n_lo) : p.k_lo;
    const int kmax = have_noise ? max(p.k_hi, p.n_hi) : p.k_hi;
    const float scale4 = 0.25f * (float)p.scale;   // the halves of E and O are folded in here (exact)
    const int partner = (32 - lane) & 31;
    const int64_t stride = (int64_t)gridDim.x * kPwWarps;

    auto frame_ptr = [&](int64_t u, int64_t& outer, int64_t& frame) {
        if (total < 0x7fffffffll) {
            outer = (uint32_t)u / (uint32_t)p.n_frames;
            frame = (uint32_t)u - (uint32_t)outer * (uint32_t)p.n_frames;
        } else {
            outer = u / p.n_frames;
            frame = u - outer * p.n_frames;
        }
        return reinterpret_cast<const Raw*>(x + outer * p.outer_stride + frame * (int64_t)p.hop) + lane;
    };

    int64_t u = (int64_t)blockIdx.x * kPwWarps + warp;
    Raw raw[32];
    int64_t outer = 0, frame = 0;
    if (kPrefetch && u < total) {
        const Raw* xf = frame_ptr(u, outer, frame);
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) raw[n1] = xf[32 * n1];
    }
    for (; u < total; u += stride) {
        float2 v[32];
        if (!kPrefetch) {
            const Raw* xf = frame_ptr(u, outer, frame);
            if (u + stride < total) {   // the next frame's 32 (PCM16) or 64 (float) lines go to L1 while this one is computed
                int64_t o2, f2;
                const char* nx = reinterpret_cast<const char*>(frame_ptr(u + stride, o2, f2) - lane) + 128 * lane;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
                if (sizeof(Raw) == 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(nx + 4096));
            }
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) raw[n1] = xf[32 * n1];
        }
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const float2 a = raw_to_f2(raw[n1]);
            const float2 w = win[32 * n1 + lane];
            v[n1] = make_float2(a.x * w.x, a.y * w.y);
        }
        const int64_t cur_outer = outer, cur_frame = frame;
#pragma unroll
        for (int stage = 0; stage < 2; ++stage) {   // rolled on purpose: one copy of fft32 in the instruction cache
            fft32(v);
            if (stage == 0) {
#pragma unroll
                for (int k1 = 1; k1 < 32; ++k1) v[k1] = cmul(v[k1], tw[32 * k1 + lane]);
#pragma unroll
                for (int k1 = 0; k1 < 32; ++k1) buf[kPwPitch * k1 + lane] = v[k1];
                __syncwarp();
#pragma unroll
                for (int n2 = 0; n2 < 32; ++n2) v[n2] = buf[kPwPitch * lane + n2];
                __syncwarp();
            }
        }
        // v[k2] = Z[lane + 32 k2]; the next frame's samples travel while the epilogue runs
        if (kPrefetch && u + stride < total) {
            const Raw* xf = frame_ptr(u + stride, outer, frame);
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) raw[n1] = xf[32 * n1];
        }

        float noise_acc = 0.0f;
        const uint32_t row0 = (uint32_t)(lane - p.k_lo), nrow0 = (uint32_t)(lane - p.n_lo);
        float* out = p.out + (cur_outer * (int64_t)(nb1 + 1)) * p.n_frames + cur_frame;
#pragma unroll
        for (int k2 = 0; k2 < K2MAX; ++k2) {
            if (32 * k2 + 31 < kmin || 32 * k2 > kmax) continue;   // warp uniform
            const float2 zk = v[k2];
            const float2 sup = (lane == 0) ? v[(32 - k2) & 31] : v[31 - k2];
            float2 zn;
            zn.x = __shfl_sync(0xffffffffu, sup.x, partner);
            zn.y = __shfl_sync(0xffffffffu, sup.y, partner);
            // 2E = Zk + conj(Zn) ; 2O = -i (Zk - conj(Zn)) ; 2X = 2E + w^k 2O, w = exp(-i pi / 1024)
            const float ex = zk.x + zn.x, ey = zk.y - zn.y;
            const float ox = zk.y + zn.y, oy = zn.x - zk.x;
            const float2 w = rtw[32 * k2 + lane];
            const float xr = ex + (w.x * ox - w.y * oy);
            const float xi = ey + (w.x * oy + w.y * ox);
            const float pw = xr * xr + xi * xi;
            const uint32_t row = row0 + 32u * k2;                  // k - k_lo, huge when k < k_lo
            if (row <= nb1) {
                float o = pw * scale4;
                if (k2 != 0 || lane != 0) o *= 2.0f;
                out[(int64_t)row * p.n_frames] = o;
            }
            if (have_noise && nrow0 + 32u * k2 <= nn1) noise_acc += (k2 != 0 || lane != 0) ? 2.0f * pw : pw;
        }
        if (have_noise) {
            noise_acc = warp_sum(noise_acc);
            if (lane == 0) atomicAdd(&p.out_noise[cur_outer], (double)(0.25f * noise_acc) * p.scale);   // prime_detection.py:83
        }
    }
}

template <typename T, int K2MAX, bool PF>
int launch_psd_warp_sized(const PsdWarpParams& p, cudaStream_t st) {
    auto kern = psd_warp_kernel<T, K2MAX, PF>;
    static thread_local int attr_dev = -1;
    int dev = 0;
    MS_CUDA_OK(cudaGetDevice(&dev));
    const size_t smem = sizeof(float2) * (size_t)(1024 + 1024 + 32 * K2MAX + kPwWarps * 32 * kPwPitch);
    if (attr_dev != dev) {
        MS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_dev = dev;
    }
    const int64_t total = p.n_outer * p.n_frames;
    if (total == 0) return MS_OK;
    int64_t grid = (int64_t)num_sms();
    const int64_t groups = (total + kPwWarps - 1) / kPwWarps;
    if (grid > groups) grid = groups;
    kern<<<(unsigned)grid, kPwThreads, smem, st>>>(p);
    MS_CUDA_OK(cudaGetLastError());
    return MS_OK;
}

template <typename T>
int launch_psd_warp(const PsdWarpParams& p, cudaStream_t st) {
    const int kmax = (p.n_lo <= p.n_hi && p.n_hi > p.k_hi) ? p.n_hi : p.k_hi;
    static const bool pf = [] {
        const char* e = getenv("MS_PSD_PREFETCH");
        return e && e[0] == '1';
    }();
    if (pf) return kmax < 512 ? launch_psd_warp_sized<T, 16, true>(p, st) : launch_psd_warp_sized<T, 32, true>(p, st);
    return kmax < 512 ? launch_psd_warp_sized<T, 16, false>(p, st) : launch_psd_warp_sized<T, 32, false>(p, st);
}

}  // namespace
}  // namespace ms
