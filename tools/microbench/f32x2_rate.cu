// Issue/throughput of packed fp32 (FADD2/FMUL2/FFMA2, sm_100a) against the scalar forms.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu && ./f32x2_rate
// Each thread runs ITER iterations of 8 independent dependency chains; 148 CTAs x 1024 threads.
// Reported: SM cycles per warp instruction per scheduler (1.0 = one instruction per clock per SMSP).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm volatile("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd1(float a, float b) { float r; asm volatile("add.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint32_t lop(uint32_t a, uint32_t b) { uint32_t r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

constexpr int ITER = 4096;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, float seed) {
    u64 p[8];
    float s[8];
    uint32_t q[8];
    for (int i = 0; i < 8; ++i) {
        float2 f = make_float2(seed + i + threadIdx.x, seed - i);
        p[i] = *reinterpret_cast<u64*>(&f);
        s[i] = seed + i + threadIdx.x;
        q[i] = threadIdx.x * 7 + i;
    }
    float2 mf = make_float2(0.999f, 1.001f);
    const u64 m = *reinterpret_cast<u64*>(&mf);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) s[i] = ffma1(s[i], 0.999f, 0.001f);
            if (MODE == 1) p[i] = ffma2(p[i], m, m);
            if (MODE == 2) s[i] = fadd1(s[i], 0.001f);
            if (MODE == 3) p[i] = fadd2(p[i], m);
            if (MODE == 4) { p[i] = ffma2(p[i], m, m); q[i] = lop(q[i], 0x5a5a5a5au + i); }            // packed + ALU 1:1
            if (MODE == 5) { s[i] = ffma1(s[i], 0.999f, 0.001f); q[i] = lop(q[i], 0x5a5a5a5au + i); }   // scalar + ALU 1:1
            if (MODE == 6) { p[i] = ffma2(p[i], m, m); s[i] = ffma1(s[i], 0.999f, 0.001f); }            // packed + scalar 1:1
        }
    }
    const long long t1 = clock64();
    float acc = 0.0f;
    for (int i = 0; i < 8; ++i) {
        float2 f = *reinterpret_cast<float2*>(&p[i]);
        acc += f.x + f.y + s[i] + (float)q[i];
    }
    out[blockIdx.x * 1024 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    k<MODE><<<148, 1024>>>(out, cyc, 1.0f);
    k<MODE><<<148, 1024>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    // 32 warps per SM = 8 per scheduler; instructions per scheduler = 8 warps * ITER * per_iter
    printf("%-28s %8.3f cycles per warp instruction per scheduler\n", name, avg / (8.0 * ITER * per_iter));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("FFMA", 8);
    run<1>("FFMA2", 8);
    run<2>("FADD", 8);
    run<3>("FADD2", 8);
    run<4>("FFMA2 + LOP3 (1:1)", 16);
    run<5>("FFMA + LOP3 (1:1)", 16);
    run<6>("FFMA2 + FFMA (1:1)", 16);
    return 0;
}
