// Microbenchmark: how fast does one SM retire tcgen05.mma.kind::i8 (M = 128, K = 32 bytes, SWIZZLE_128B K-major operands
// in shared memory) as a function of N, of how the instructions are spread over accumulators, and of whether the A
// operand start address moves (adjacent 32-byte K chunks / different row tiles)?  One CTA per SM, one issuing thread,
// `iters` x `unroll` instructions, one commit + wait at the end; prints cycles per instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_rate umma_i8_rate.cu && ./umma_i8_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_i8(int n, int m) {
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

// mode 0: same A chunk, same accumulator; 1: A walks the 4 K chunks of a slab, same accumulator;
// 2: A walks K chunks, accumulator alternates between 2; 3: A alternates between 2 row tiles + K chunks, 2 accumulators;
// 4: like 1 but A also walks 8 different tiles (128 KiB of A in rotation); 5: the same arithmetic confined to 2 tiles;
// 6: 4 tiles (64 KiB) in rotation; 7: 2 tiles at 96 / 112 KiB; 8: 2 tiles 112 KiB apart; 9: 6 tiles in rotation
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int mode, int iters, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    unsigned char* a = smem;                    // 4 stages x 2 tiles x 16 KiB
    unsigned char* b = smem + 8 * 16384;        // 32 KiB
    for (int i = threadIdx.x; i < (8 * 16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_i8(N, 128);
        const uint32_t a0 = smem_u32(a), b0 = smem_u32(b);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                uint32_t aa = a0, dd = tm;
                if (mode >= 1) aa += (u & 3) * 32;
                if (mode == 2) dd += (u & 1) * 256;
                if (mode == 3) {
                    aa += ((u >> 2) & 1) * 16384;
                    dd += ((u >> 2) & 1) * 256;
                }
                if (mode == 4) aa += ((it & 3) * 2 + (u >> 2)) * 16384;
                if (mode == 5) aa += (((it & 3) * 2 + (u >> 2)) & 1) * 16384;     // mode 4's address arithmetic, 2 tiles only
                if (mode == 6) aa += ((it & 1) * 2 + (u >> 2)) * 16384;          // 4 tiles = 64 KiB of A in rotation
                if (mode == 7) aa += (6 + (u >> 2)) * 16384;                     // 2 tiles at 96 / 112 KiB
                if (mode == 8) aa += (u >> 2) * 7 * 16384;                       // 2 tiles, 0 and 112 KiB apart
                if (mode == 9) aa += (((it & 3) * 2 + (u >> 2)) % 6) * 16384;    // 6 tiles in rotation
                mma(dd, desc_sw128(aa), desc_sw128(b0 + (u & 3) * 32), idesc, (it | u) ? 1u : 0u);
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    const int smem = 8 * 16384 + 32768;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int grid : {sms}) {
        for (int N : {64, 96, 128, 256}) {
            for (int mode = 0; mode <= 9; ++mode) {
                long long h[2];
                for (int rep = 0; rep < 2; ++rep) {
                    rate_kernel<<<grid, 128, smem>>>(N, mode, iters, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) {
                        printf("error: %s\n", cudaGetErrorString(e));
                        return 1;
                    }
                }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("grid %3d N %3d mode %d: issue %.1f cyc/mma, complete %.1f cyc/mma\n", grid, N, mode,
                       (double)h[0] / (iters * 8), (double)h[1] / (iters * 8));
            }
        }
    }
    return 0;
}
