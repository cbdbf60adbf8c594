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
//
// The kernel is bound by instruction issue, so the 32-point transforms use Blackwell's packed fp32 pipe
// (add/sub/mul/fma.f32x2 = FADD2/FMUL2/FFMA2: two IEEE fp32 operations per issue slot, measured at the same
// flops per clock as the scalar forms, tools/microbench/f32x2_rate.cu): a 32-point transform is split by one
// radix-2 level into two 16-point transforms that run in lockstep in the two halves of 64-bit registers
// (stage 1 decimation in time: even/odd inputs first, scalar combine last; stage 2 decimation in frequency:
// scalar split first, lockstep transforms last), equal constants are immediates of the packed instructions.
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

// ---- packed fp32 pairs (sm_100a): two lockstep transforms live in the halves of a 64-bit register ----
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk(float lo, float hi) {
    pk2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(pk2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ pk2 add2(pk2 a, pk2 b) {
    pk2 r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pk2 sub2(pk2 a, pk2 b) {
    pk2 r;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pk2 mul2(pk2 a, pk2 b) {
    pk2 r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pk2 fma2(pk2 a, pk2 b, pk2 c) {
    pk2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ pk2 bc(float c) { return pk(c, c); }   // equal halves fold into an immediate

// cos / sin of pi m / 16 for the constant twiddles exp(-2 pi i m / 32)
__device__ __forceinline__ void w32_cs(int m, float& c, float& s) {
    constexpr float S = 0.70710678118654752440f;
    constexpr float C1 = 0.98078528040323044913f, S1 = 0.19509032201612826785f;   // pi/16
    constexpr float C2 = 0.92387953251128675613f, S2 = 0.38268343236508977173f;   // pi/8
    constexpr float C3 = 0.83146961230254523708f, S3 = 0.55557023301960222474f;   // 3 pi/16
    switch (m) {   // m is a compile-time constant after unrolling
        case 0: c = 1.0f; s = 0.0f; break;
        case 1: c = C1; s = S1; break;
        case 2: c = C2; s = S2; break;
        case 3: c = C3; s = S3; break;
        case 4: c = S; s = S; break;
        case 5: c = S3; s = C3; break;
        case 6: c = S2; s = C2; break;
        case 7: c = S1; s = C1; break;
        case 8: c = 0.0f; s = 1.0f; break;
        case 9: c = -S1; s = C1; break;
        case 10: c = -S2; s = C2; break;
        case 11: c = -S3; s = C3; break;
        case 12: c = -S; s = S; break;
        case 13: c = -C3; s = S3; break;
        case 14: c = -C2; s = S2; break;
        case 15: c = -C1; s = S1; break;
        case 16: c = -1.0f; s = 0.0f; break;
        default: c = -C2; s = -S2; break;   // m = 18 (= 2 * 9), the only other product used
    }
}

// a * exp(-2 pi i m / 32), scalar
__device__ __forceinline__ float2 mul_w32(float2 a, int m) {
    constexpr float S = 0.70710678118654752440f;
    if (m == 0) return a;
    if (m == 4) return make_float2(S * (a.x + a.y), S * (a.y - a.x));
    if (m == 8) return make_float2(a.y, -a.x);
    if (m == 12) return make_float2(S * (a.y - a.x), -S * (a.x + a.y));
    float c, s;
    w32_cs(m, c, s);
    return make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
}

// (r, i) * exp(-2 pi i m / 16) on both halves
__device__ __forceinline__ void mul_w16p(pk2& r, pk2& i, int m) {
    constexpr float S = 0.70710678118654752440f;
    if (m == 0) return;
    if (m == 2) {
        const pk2 a = add2(r, i), d = sub2(i, r);
        r = mul2(a, bc(S));
        i = mul2(d, bc(S));
    } else if (m == 4) {
        const pk2 t = r;
        r = i;
        i = mul2(t, bc(-1.0f));
    } else if (m == 6) {
        const pk2 a = add2(r, i), d = sub2(i, r);
        r = mul2(d, bc(S));
        i = mul2(a, bc(-S));
    } else {
        float c, s;
        w32_cs(2 * m, c, s);
        const pk2 nr = fma2(i, bc(s), mul2(r, bc(c)));
        i = fma2(r, bc(-s), mul2(i, bc(c)));
        r = nr;
    }
}

// forward radix-4 butterfly on both halves, natural order
__device__ __forceinline__ void bfly4p(pk2& r0, pk2& i0, pk2& r1, pk2& i1, pk2& r2, pk2& i2, pk2& r3, pk2& i3) {
    const pk2 t0r = add2(r0, r2), t0i = add2(i0, i2), t1r = sub2(r0, r2), t1i = sub2(i0, i2);
    const pk2 t2r = add2(r1, r3), t2i = add2(i1, i3), dr = sub2(r1, r3), di = sub2(i1, i3);
    r0 = add2(t0r, t2r);
    i0 = add2(t0i, t2i);
    r2 = sub2(t0r, t2r);
    i2 = sub2(t0i, t2i);
    r1 = add2(t1r, di);   // t1 - i d
    i1 = sub2(t1i, dr);
    r3 = sub2(t1r, di);   // t1 + i d
    i3 = add2(t1i, dr);
}

// two forward 16-point transforms in lockstep (halves of R/I), natural order in and out: n = 4a + b, k = c + 4d
__device__ __forceinline__ void fft16p(pk2 (&R)[16], pk2 (&I)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        bfly4p(R[b], I[b], R[4 + b], I[4 + b], R[8 + b], I[8 + b], R[12 + b], I[12 + b]);
#pragma unroll
        for (int c = 1; c < 4; ++c) mul_w16p(R[4 * c + b], I[4 * c + b], b * c);
    }
    pk2 r[16], i[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        bfly4p(R[4 * c], I[4 * c], R[4 * c + 1], I[4 * c + 1], R[4 * c + 2], I[4 * c + 2], R[4 * c + 3], I[4 * c + 3]);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            r[c + 4 * d] = R[4 * c + d];
            i[c + 4 * d] = I[4 * c + d];
        }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        R[k] = r[k];
        I[k] = i[k];
    }
}

// 32-point transform, decimation in time: halves hold the even / odd inputs (R[j] = (x[2j].re, x[2j+1].re));
// X[k] = E[k] + w32^k O[k], X[k+16] = E[k] - w32^k O[k] come out as scalars
__device__ __forceinline__ void fft32_dit(pk2 (&R)[16], pk2 (&I)[16], float2 (&v)[32]) {
    fft16p(R, I);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        float2 e, o;
        upk(R[k], e.x, o.x);
        upk(I[k], e.y, o.y);
        o = mul_w32(o, k);
        v[k] = make_float2(e.x + o.x, e.y + o.y);
        v[k + 16] = make_float2(e.x - o.x, e.y - o.y);
    }
}

// 32-point transform, decimation in frequency: a[j] = x[j] + x[j+16] and b[j] = (x[j] - x[j+16]) w32^j go to the
// halves, the lockstep 16-point transforms give X[2k] and X[2k+1]
__device__ __forceinline__ void fft32_dif(float2 (&v)[32]) {
    pk2 R[16], I[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 a = make_float2(v[j].x + v[j + 16].x, v[j].y + v[j + 16].y);
        const float2 b = mul_w32(make_float2(v[j].x - v[j + 16].x, v[j].y - v[j + 16].y), j);
        R[j] = pk(a.x, b.x);
        I[j] = pk(a.y, b.y);
    }
    fft16p(R, I);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        upk(R[k], v[2 * k].x, v[2 * k + 1].x);
        upk(I[k], v[2 * k].y, v[2 * k + 1].y);
    }
}

// raw sample pair of one complex point -> floats (PCM16: the mantissa trick of load_pair, bias still on; the
// one-instruction I2F.S16 .H0/.H1 form was measured 9 % slower end to end: the XU pipe runs at a quarter rate)
__device__ __forceinline__ float2 raw_biased(uint32_t w) {
    w ^= 0x80008000u;
    return make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)),
                       __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)));
}
template <typename T> struct RawPair;
template <> struct RawPair<int16_t> { typedef uint32_t type; };
template <> struct RawPair<float> { typedef float2 type; };

// K2MAX: number of 32-bin groups the epilogue is unrolled over (16 when every wanted bin is below 512)
template <typename T, int K2MAX>
__global__ void __launch_bounds__(kPwThreads, 1) psd_warp_kernel(PsdWarpParams p) {
    typedef typename RawPair<T>::type Raw;
    extern __shared__ __align__(16) float2 pw_smem[];
    float2* tw = pw_smem;              // [k1][n2] exp(-2 pi i n2 k1 / 1024)
    float4* win4 = reinterpret_cast<float4*>(tw + 1024);   // [j][lane]: window of points 32(2j)+lane and 32(2j+1)+lane: (re_e, re_o, im_e, im_o)
    float4* rtw4 = reinterpret_cast<float4*>(tw + 2048);   // [q][lane]: real-split twiddles exp(-i pi k / 1024) of bins k = 64 q + lane (a) and k + 32 (b): (re_a, re_b, im_a, im_b)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* buf = tw + 2048 + 32 * K2MAX + warp * (32 * kPwPitch);

    for (int e = tid; e < 1024; e += kPwThreads) {
        float sn, cs;
        sincospif(-2.0f * (float)((e >> 5) * (e & 31)) / 1024.0f, &sn, &cs);
        tw[e] = make_float2(cs, sn);
        if (e < 512) {
            const int j = e >> 5, l = e & 31;
            const float2 we = __ldg(reinterpret_cast<const float2*>(p.window) + 32 * (2 * j) + l);
            const float2 wo = __ldg(reinterpret_cast<const float2*>(p.window) + 32 * (2 * j + 1) + l);
            win4[e] = make_float4(we.x, wo.x, we.y, wo.y);
        }
        if (e < 16 * K2MAX) {
            const int ka = 64 * (e >> 5) + (e & 31);
            float sb, cb;
            sincospif(-(float)ka / 1024.0f, &sn, &cs);
            sincospif(-(float)(ka + 32) / 1024.0f, &sb, &cb);
            rtw4[e] = make_float4(cs, cb, sn, sb);
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
        return reinterpret_cast<const Raw*>(x + outer * p.outer_stride + frame * (int64_t)p.hop);
    };

    for (int64_t u = (int64_t)blockIdx.x * kPwWarps + warp; u < total; u += stride) {
        int64_t outer, frame;
        const Raw* xf = frame_ptr(u, outer, frame) + lane;
        if (u + stride < total) {   // the next frame's 32 (PCM16) or 64 (float) lines go to L1 while this one is computed
            int64_t o2, f2;
            const char* nx = reinterpret_cast<const char*>(frame_ptr(u + stride, o2, f2)) + 128 * lane;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
            if (sizeof(Raw) == 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(nx + 4096));
        }
        float2 v[32];
        {
            Raw raw[32];
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) raw[n1] = xf[32 * n1];
            pk2 R[16], I[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 w = win4[32 * j + lane];
                if constexpr (sizeof(T) == 2) {
                    const float2 e = raw_biased(raw[2 * j]), o = raw_biased(raw[2 * j + 1]);
                    R[j] = mul2(sub2(pk(e.x, o.x), bc(kI16Bias)), pk(w.x, w.y));
                    I[j] = mul2(sub2(pk(e.y, o.y), bc(kI16Bias)), pk(w.z, w.w));
                } else {
                    R[j] = mul2(pk(raw[2 * j].x, raw[2 * j + 1].x), pk(w.x, w.y));
                    I[j] = mul2(pk(raw[2 * j].y, raw[2 * j + 1].y), pk(w.z, w.w));
                }
            }
            fft32_dit(R, I, v);
        }
#pragma unroll
        for (int k1 = 1; k1 < 32; ++k1) v[k1] = cmul(v[k1], tw[32 * k1 + lane]);
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) buf[kPwPitch * k1 + lane] = v[k1];
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) v[n2] = buf[kPwPitch * lane + n2];
        __syncwarp();
        fft32_dif(v);   // v[k2] = Z[lane + 32 k2]

        float noise_acc = 0.0f;
        const uint32_t row0 = (uint32_t)(lane - p.k_lo), nrow0 = (uint32_t)(lane - p.n_lo);
        float* out = p.out + (outer * (int64_t)(nb1 + 1)) * p.n_frames + frame;
#pragma unroll
        for (int q = 0; q < K2MAX / 2; ++q) {   // bins 64 q + lane and 64 q + 32 + lane, both halves of packed registers
            if (64 * q + 63 < kmin || 64 * q > kmax) continue;   // warp uniform
            const int ka = 2 * q, kb = 2 * q + 1;
            const float2 sa = (lane == 0) ? v[(32 - ka) & 31] : v[31 - ka];
            const float2 sb = (lane == 0) ? v[(32 - kb) & 31] : v[31 - kb];
            const pk2 znx = pk(__shfl_sync(0xffffffffu, sa.x, partner), __shfl_sync(0xffffffffu, sb.x, partner));
            const pk2 zny = pk(__shfl_sync(0xffffffffu, sa.y, partner), __shfl_sync(0xffffffffu, sb.y, partner));
            const pk2 zkx = pk(v[ka].x, v[kb].x), zky = pk(v[ka].y, v[kb].y);
            // 2E = Zk + conj(Zn) ; 2O = -i (Zk - conj(Zn)) ; 2X = 2E + w^k 2O, w = exp(-i pi / 1024)
            const pk2 ex = add2(zkx, znx), ey = sub2(zky, zny), ox = add2(zky, zny), oy = sub2(znx, zkx);
            const float4 w = rtw4[32 * q + lane];   // (w.re of bin a, of bin b, w.im of bin a, of bin b)
            const pk2 wx = pk(w.x, w.y), wy = pk(w.z, w.w);
            const pk2 xr = fma2(wx, ox, sub2(ex, mul2(wy, oy)));
            const pk2 xi = fma2(wx, oy, fma2(wy, ox, ey));
            const pk2 pw2 = fma2(xi, xi, mul2(xr, xr));
            float pwv[2];
            upk(pw2, pwv[0], pwv[1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k2 = 2 * q + h;
                const float pw = pwv[h];
                const uint32_t row = row0 + 32u * k2;                  // k - k_lo, huge when k < k_lo
                if (row <= nb1) {
                    float o = pw * scale4;
                    if (k2 != 0 || lane != 0) o *= 2.0f;
                    out[(int64_t)row * p.n_frames] = o;
                }
                if (have_noise && nrow0 + 32u * k2 <= nn1) noise_acc += (k2 != 0 || lane != 0) ? 2.0f * pw : pw;
            }
        }
        if (have_noise) {
            noise_acc = warp_sum(noise_acc);
            if (lane == 0) atomicAdd(&p.out_noise[outer], (double)(0.25f * noise_acc) * p.scale);   // prime_detection.py:83
        }
    }
}

template <typename T, int K2MAX>
int launch_psd_warp_sized(const PsdWarpParams& p, cudaStream_t st) {
    auto kern = psd_warp_kernel<T, K2MAX>;
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
    return kmax < 512 ? launch_psd_warp_sized<T, 16>(p, st) : launch_psd_warp_sized<T, 32>(p, st);
}

}  // namespace
}  // namespace ms
