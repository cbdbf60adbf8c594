// B-psd + B-band without FFTs: the Welch band power of a block as a low-rank quadratic form.
//
// Reference semantics (dsp/src/live/backend/processor.py:206, 349-367, 393):
//   welch(block, fs, nfft): n_sub segments of `nperseg` samples (hop nperseg/2), each mean-removed, Hann-windowed,
//   zero-padded to nfft, |X[k]|^2 * scale (x2 except DC/Nyquist), averaged over the segments; then the sum over an
//   inclusive bin range per band -> 10*log10.
// For one segment x (length K = nperseg) and one band:
//   sum_k c_k |sum_n (x_n - mean(x)) w_n e^{-2 pi i k n / nfft}|^2  =  x^T (P Q P) x,
//   Q[n,m] = w_n w_m sum_k c_k cos(2 pi k (n-m)/nfft),  P = I - 11^T/K  (the mean removal).
// P Q P is symmetric PSD with numerical rank ~ 2*K*bandwidth/nfft + O(log) (26 for 102 of 4096 bins at K = 256), so
// with its leading eigenpairs  band power = lambda_max * sum_r (b_r . x)^2,  b_r = sqrt(lambda_r/lambda_max) u_r.
// The host builds the basis (ops.WelchQuadform); this kernel evaluates the projections:
//   one warp per block; the block's samples are staged in shared memory; lane l owns column l of EACH band
//   (3 columns, up to 32 per band) and accumulates them for all n_sub segments, so no cross-lane traffic is
//   needed until the final per-band sums.  Five zero-padded 4096-point FFTs per block become 256 x 78 FMAs x 5.
#include "ms_common.cuh"

namespace ms {
namespace {

constexpr int kQfThreads = 512;       // 16 warps share one basis copy in shared memory; one block of audio per warp
constexpr int kQfMaxSub = 8;          // segments per block
constexpr int kQfBands = 3;
// (32 basis columns per band: one per lane)

struct WelchQfParams {
    const void* x;
    int64_t n_streams, stream_stride, n_blocks;
    int32_t block, nperseg, hop, n_sub;
    float in_scale;                   // 1/32768 for PCM16 (soundfile semantics)
    const float4* basis;              // [nperseg][32] float4: (band0, band1, band2, 0) column of lane l at sample n
    float group_scale[kQfBands];      // lambda_max * scale / n_sub per band
    float* out_db;                    // [n_streams][n_blocks][4]
};

__device__ __forceinline__ float qf_load(const int16_t* x, int64_t i) { return (float)x[i]; }
__device__ __forceinline__ float qf_load(const float* x, int64_t i) { return x[i]; }

template <typename T>
__global__ void __launch_bounds__(kQfThreads) welch_qf_kernel(WelchQfParams p) {
    extern __shared__ __align__(16) unsigned char qf_smem[];
    float4* sb = reinterpret_cast<float4*>(qf_smem);                        // basis [nperseg][32]
    float* sx = reinterpret_cast<float*>(sb + (size_t)p.nperseg * 32);      // 4 warps x block_pad samples
    const int block_pad = (p.block + 3) & ~3;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < p.nperseg * 32; i += kQfThreads) sb[i] = p.basis[i];
    __syncthreads();

    const T* x = static_cast<const T*>(p.x);
    float* xs = sx + (size_t)warp * block_pad;
    const int64_t total = p.n_streams * p.n_blocks;
    const int64_t wstride = (int64_t)gridDim.x * (kQfThreads / 32);
    for (int64_t u = (int64_t)blockIdx.x * (kQfThreads / 32) + warp; u < total; u += wstride) {
        const int64_t s = u / p.n_blocks, b = u - s * p.n_blocks;
        const int64_t base = s * p.stream_stride + b * (int64_t)p.block;
        __syncwarp();
        for (int i = lane; i < p.block; i += 32) xs[i] = qf_load(x, base + i) * p.in_scale;
        __syncwarp();

        float a0[kQfMaxSub], a1[kQfMaxSub], a2[kQfMaxSub];
        const float4* xq[kQfMaxSub];      // segment q of this block (16-byte aligned: hop % 4 == 0)
#pragma unroll
        for (int q = 0; q < kQfMaxSub; ++q) {
            a0[q] = a1[q] = a2[q] = 0.0f;
            xq[q] = reinterpret_cast<const float4*>(xs + (q < p.n_sub ? q : 0) * p.hop);
        }
        const float4* bl = sb + lane;
        const int n4 = p.nperseg >> 2;
#pragma unroll 2
        for (int n = 0; n < n4; ++n) {
            const float4 b0 = bl[(4 * n + 0) * 32], b1 = bl[(4 * n + 1) * 32];
            const float4 b2 = bl[(4 * n + 2) * 32], b3 = bl[(4 * n + 3) * 32];
#pragma unroll
            for (int q = 0; q < kQfMaxSub; ++q) {
                if (q < p.n_sub) {
                    const float4 xv = xq[q][n];   // broadcast
                    a0[q] = fmaf(xv.x, b0.x, fmaf(xv.y, b1.x, fmaf(xv.z, b2.x, fmaf(xv.w, b3.x, a0[q]))));
                    a1[q] = fmaf(xv.x, b0.y, fmaf(xv.y, b1.y, fmaf(xv.z, b2.y, fmaf(xv.w, b3.y, a1[q]))));
                    a2[q] = fmaf(xv.x, b0.z, fmaf(xv.y, b1.z, fmaf(xv.z, b2.z, fmaf(xv.w, b3.z, a2[q]))));
                }
            }
        }
        float e0 = 0.0f, e1 = 0.0f, e2 = 0.0f;
#pragma unroll
        for (int q = 0; q < kQfMaxSub; ++q) {
            e0 = fmaf(a0[q], a0[q], e0);
            e1 = fmaf(a1[q], a1[q], e1);
            e2 = fmaf(a2[q], a2[q], e2);
        }
        e0 = warp_sum(e0);
        e1 = warp_sum(e1);
        e2 = warp_sum(e2);
        if (lane == 0) {
            const float pw[3] = {e0 * p.group_scale[0], e1 * p.group_scale[1], e2 * p.group_scale[2]};
            float db[3];
#pragma unroll
            for (int g = 0; g < 3; ++g) db[g] = pw[g] > 0.0f ? 10.0f * log10f(pw[g]) : -INFINITY;   // processor.py:352
            float* o = p.out_db + u * 4;
            o[0] = db[0];
            o[1] = db[1];
            o[2] = db[2];
            o[3] = db[0] - 0.5f * (db[1] + db[2]);                                                     // processor.py:393
        }
    }
}

template <typename T>
int launch_welch_qf(const T* x, float in_scale, int64_t n_streams, int64_t stream_stride, int64_t n_blocks, int32_t block,
                    int32_t nperseg, const float* d_basis, const double* h_group_scale, float* out_db, void* stream) {
    MS_REQUIRE(x && d_basis && h_group_scale && out_db, MS_ERR_INVALID_ARG, "ms_welch_band_db_qf: null pointer");
    MS_REQUIRE(block > 0 && nperseg >= 4 && nperseg % 4 == 0 && nperseg <= block, MS_ERR_UNSUPPORTED,
               "ms_welch_band_db_qf: nperseg must be a multiple of 4 and <= block");
    const int hop = nperseg - nperseg / 2;
    const int n_sub = (block - nperseg / 2) / hop;
    MS_REQUIRE(n_sub >= 1 && n_sub <= kQfMaxSub && hop % 4 == 0, MS_ERR_UNSUPPORTED,
               "ms_welch_band_db_qf: %d segments per block (max %d), hop must be a multiple of 4", n_sub, kQfMaxSub);
    MS_REQUIRE(n_blocks == 0 || n_blocks * (int64_t)block <= stream_stride, MS_ERR_INVALID_ARG,
               "ms_welch_band_db_qf: blocks exceed stream_stride");
    MS_REQUIRE((reinterpret_cast<uintptr_t>(d_basis) & 15) == 0, MS_ERR_INVALID_ARG, "ms_welch_band_db_qf: basis alignment");
    const int64_t total = n_streams * n_blocks;
    if (total == 0) return MS_OK;
    WelchQfParams p = {};
    p.x = x;
    p.n_streams = n_streams;
    p.stream_stride = stream_stride;
    p.n_blocks = n_blocks;
    p.block = block;
    p.nperseg = nperseg;
    p.hop = hop;
    p.n_sub = n_sub;
    p.in_scale = in_scale;
    p.basis = reinterpret_cast<const float4*>(d_basis);
    for (int g = 0; g < kQfBands; ++g) p.group_scale[g] = (float)h_group_scale[g];
    p.out_db = out_db;
    const int block_pad = (block + 3) & ~3;
    const size_t smem = (size_t)nperseg * 32 * sizeof(float4) + (size_t)(kQfThreads / 32) * block_pad * sizeof(float);
    MS_REQUIRE(smem <= 220 * 1024, MS_ERR_UNSUPPORTED, "ms_welch_band_db_qf: nperseg/block too large for shared memory");
    auto kern = welch_qf_kernel<T>;
    if (smem > 40 * 1024) MS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kQfThreads, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)num_sms() * per_sm;
    const int64_t need = (total + kQfThreads / 32 - 1) / (kQfThreads / 32);
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, kQfThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    MS_CUDA_OK(cudaGetLastError());
    return MS_OK;
}

}  // namespace
}  // namespace ms

extern "C" {

int ms_welch_band_db_qf_i16(const int16_t* x, int64_t n_streams, int64_t stream_stride, int64_t n_blocks, int32_t block,
                            int32_t nperseg, const float* d_basis, const double* h_group_scale, float* out_db,
                            void* stream) {
    return ms::launch_welch_qf<int16_t>(x, 1.0f / 32768.0f, n_streams, stream_stride, n_blocks, block, nperseg, d_basis,
                                        h_group_scale, out_db, stream);
}

int ms_welch_band_db_qf_f32(const float* x, int64_t n_streams, int64_t stream_stride, int64_t n_blocks, int32_t block,
                            int32_t nperseg, const float* d_basis, const double* h_group_scale, float* out_db,
                            void* stream) {
    return ms::launch_welch_qf<float>(x, 1.0f, n_streams, stream_stride, n_blocks, block, nperseg, d_basis, h_group_scale,
                                      out_db, stream);
}

}  // extern "C"
