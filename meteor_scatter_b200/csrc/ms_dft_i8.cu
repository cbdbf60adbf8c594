// K2: band energies of the STFT as a restricted DFT on the 5th-gen tensor cores,
// evaluated in EXACT integer arithmetic (tcgen05.mma kind::i8, int32 TMEM
// accumulators).  Same reference semantics as K1: dsp/src/main.py:376-388
// (np.fft.rfft(block*np.hanning(len(block)), n=n_fft) -> |X|^2 -> masked band
// sums -> 10*log10(.+1e-12)); only the handful of bins inside the two bands are
// ever needed, so X[k] = sum_n x[n]*w[n]*exp(-2*pi*i*k*n/nfft) is a
// [rows x K] x [K x 16] product.
//
// Number format.  PCM16 sample x = 256*hi + lo (hi signed byte, lo unsigned
// byte) sits in HBM as the byte pair (lo, hi).  A frame of K samples is read as
// a row of 2K unsigned bytes (TMA, 128B-swizzled); the hi bytes are turned into
// offset binary (hi+128) by an XOR pass on the way to tensor memory (default) or
// in place (MS_K2_TS=0), so the row is a valid u8 operand.  Each basis value b = w[n]*cos/sin is quantised to v = round(b*scale)
// with scale = 0.99*2^23 / max|b| (the plan's largest value uses the full range)
// and split into three balanced base-256 digits v = q1*2^16 + q2*2^8 + q3
// (s8).  With four 16-column slices the MMA accumulates
//   S0 = sum hi*q1, S1 = sum hi*q2+lo*q1, S2 = sum hi*q3+lo*q2, S3 = sum lo*q3
// and X*scale = S0*2^24 + S1*2^16 + S2*2^8 + S3 exactly (|.| < 2^53, so the fp64
// epilogue is exact up to the scaling and the final squares/sum/log10).  The only
// approximation is the 2^-24 (relative to the basis peak) quantisation of the basis.
//
// Pipeline per CTA (persistent, one CTA per SM):
//   warp 0      TMA producer: [128 rows x 128 B] boxes -> smem stage (6 stages at K=1024), mbarrier tx
//   warps 2-9   fix-up: landed stage -> registers -> XOR 0x80 into the hi bytes -> tcgen05.st into a 12-slot ring of
//               A slabs in tensor memory (one thread per row; the overlapped pass runs the same with 4 warps)
//   warp 1      MMA issuer: 4 x UTCIMMA (M128 N64 K32) per 128-byte K slab, A from tensor memory
//   last 4      epilogue: tcgen05.ld 64 columns/row -> fp64 combine -> dB -> HBM
// The 64x(2K)-byte basis lives in shared memory for the whole kernel.
#include <cuda.h>

#include <stdlib.h>

#include <vector>

#include "ms_async.cuh"
#include "ms_common.cuh"
#include "ms_umma.cuh"

namespace ms {
namespace {

constexpr int kTileRows = 128;          // UMMA M
constexpr int kSlabBytes = 128;         // K bytes per pipeline stage (one swizzle atom)
constexpr int kN = 64;                  // UMMA N: 4 digit slices x 16 columns
constexpr int kCols = 16;               // basis columns (cos/sin pairs of up to 8 bins)
constexpr int kStageBytes = kTileRows * kSlabBytes;   // 16 KiB
constexpr int kBSlabBytes = kN * kSlabBytes;          // 8 KiB
constexpr int kMaxStages = 8;             // operand pipeline depth is chosen at launch: as many 16 KiB stages as fit
constexpr int kMinStages = 3;
constexpr int kTmemCols = 128;          // two 64-column accumulators
constexpr int kASlots = 12;             // TS form: ring of A slabs in tensor memory, 32 columns (128 rows x 128 B) each
constexpr int kTmemColsTS = 512;        // 128 accumulator columns + 12 x 32 operand columns
constexpr int kFixWarpsDefault = 8;        // warps turning hi bytes into offset binary (template parameter: 4 or 8)
constexpr uint32_t kPlanMagic = 0x4d534938u;  // "MSI8"
constexpr int kPlanHeaderBytes = 1024;
constexpr int kFracBits = 23;             // basis digits: 3 x s8 -> |v| < 2^23
constexpr double kBasisPeak = 0.99;       // the largest |basis value| maps to 0.99 * 2^23 (leading digit <= 127)

struct PlanHeader {
    uint32_t magic;
    int32_t k_samples;     // samples per frame entering the transform
    int32_t n_cols;
    int32_t n_slabs;       // ceil(2*k_samples / 128)
    double inv_scale;      // combined slices * inv_scale = X (scale = 0.99 * 2^23 / max|basis|)
    int32_t offs[kN];      // 128 * sum_n digit (offset-binary correction per accumulator column)
    int32_t group[kCols];  // 0 signal band, 1 noise band, -1 unused
};
static_assert(sizeof(PlanHeader) <= kPlanHeaderBytes, "plan header too large");

// kind::i8: D=S32, A=u8, B=s8, both K-major, N=64, M=128 (helpers: ms_umma.cuh / ms_async.cuh)
__device__ __forceinline__ uint32_t k2_idesc() { return umma_idesc_i8(kN, kTileRows); }

// Slab order within a tile, rotated per CTA: the integer accumulation is order-independent, and CTAs that run in
// lockstep then do not all request the same 128-byte column of their 2 KiB rows (same DRAM channel bits) at once.
__device__ __forceinline__ int slab_at(int s0, int rot, int n_slabs) {
    const int s = s0 + rot;
    return s >= n_slabs ? s - n_slabs : s;
}

struct SmemLayout {
    static constexpr int kBarBytes = 512;
    __host__ __device__ static size_t bytes(int n_slabs, int n_stages) {
        return (size_t)n_slabs * kBSlabBytes + (size_t)n_stages * kStageBytes + kBarBytes +
               sizeof(PlanHeader);
    }
    // deepest pipeline that fits next to the resident basis (bytes in flight bound the HBM throughput)
    __host__ static int stages_for(int n_slabs) {
        int n = kMaxStages;
        while (n >= kMinStages && bytes(n_slabs, n) > (size_t)227 * 1024) --n;
        return n;   // < kMinStages: does not fit
    }
};

// TS = true (default; MS_K2_TS=0 selects the other form): the fix-up warps do not rewrite the
// landed stage in shared memory; they load it (one thread per row, conflict free through the 128-byte swizzle), flip
// the hi bytes in registers and store it to a ring in TENSOR memory (tcgen05.st 32x32b.x16, lane = row), and the MMAs
// take A from there (tcgen05.mma [d], [a_tmem], b_desc).  The stage is free again as soon as it has been read, and
// neither the fix-up's write-back nor the MMA's A fetch touch shared memory: 2 instead of 4 shared-memory passes per
// operand byte.  Measured on the same box: 0.1557 -> 0.1491 ms per 24 h batch (0.87 -> 0.91 of the HBM roofline).
template <int FIX_WARPS, bool TS>
__global__ void __launch_bounds__(64 + 32 * FIX_WARPS + 128, 1)
dft_i8_kernel(const __grid_constant__ CUtensorMap tmap, const unsigned char* __restrict__ plan, int64_t n_rows,
              int64_t n_files, int64_t out_stride, int n_slabs, float* __restrict__ out_band_db, float* __restrict__ out_noise_db,
              float* __restrict__ out_band_e, float* __restrict__ out_noise_e, int32_t* __restrict__ zero_buf,
              int zero_count, int n_stages, int slab_rot, int a_slots) {
    constexpr int fix_warps = FIX_WARPS;
    // The swizzled operand slabs need 1 KiB alignment.  The kernel has no static shared memory, so the dynamic window
    // starts at the CTA's (1 KiB aligned) base; no slack is requested -- those bytes are what lets a small-footprint
    // detect CTA share the SM -- and a misaligned base stops the kernel instead of corrupting operands.
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) {
        if (threadIdx.x == 0) printf("dft_i8_kernel: dynamic shared memory base is not 1 KiB aligned\n");
        __trap();
    }
    unsigned char* smem_b = smem;                                        // n_slabs x 8 KiB
    unsigned char* smem_a = smem_b + (size_t)n_slabs * kBSlabBytes;      // n_stages x 16 KiB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)n_stages * kStageBytes);
    uint64_t* full = bars;                    // TMA landed
    uint64_t* ready = bars + kMaxStages;      // fix-up done
    uint64_t* empty = bars + 2 * kMaxStages;  // MMAs that read the stage retired
    uint64_t* tfull = bars + 3 * kMaxStages;  // accumulator complete [2]
    uint64_t* tempty = tfull + 2;             // accumulator drained [2]
    uint64_t* bbar = tempty + 2;              // basis landed
    uint64_t* aempty = bbar + 1;              // TS: MMAs that read the tensor-memory slot retired [kASlots]
    uint64_t* aready = aempty + kASlots;      // TS: tensor-memory slot filled [kASlots]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aready + kASlots);
    PlanHeader* hdr = reinterpret_cast<PlanHeader*>(reinterpret_cast<unsigned char*>(bars) + SmemLayout::kBarBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rot = (int)((blockIdx.x * (unsigned)slab_rot) % (unsigned)n_slabs);
    // n_files == 0: one flat row space (2-D map, n_rows rows).  n_files > 0: rank-3 map [file][row][bytes] with
    // n_rows rows per file (frames may overlap or files may have gaps); a tile never crosses a file.
    const int64_t tiles_per_file = (n_rows + kTileRows - 1) / kTileRows;
    const int64_t n_tiles = (n_files > 0) ? n_files * tiles_per_file : tiles_per_file;

    // one-call pass: clear the hourly histogram that the detect kernel (next in the stream) accumulates into
    if (zero_buf != nullptr && blockIdx.x == 0)
        for (int i = threadIdx.x; i < zero_count; i += (int)blockDim.x) zero_buf[i] = 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&ready[s], (uint32_t)fix_warps);
            mbar_init(&empty[s], TS ? (uint32_t)fix_warps : 1u);
        }
        for (int s = 0; s < kASlots; ++s) {
            mbar_init(&aempty[s], 1);
            mbar_init(&aready[s], (uint32_t)fix_warps);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);
        }
        mbar_init(bbar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {  // TMEM allocation is warp-collective; this warp also frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TS ? kTmemColsTS : kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < (int)(sizeof(PlanHeader) / 4); i += (int)blockDim.x)
        reinterpret_cast<uint32_t*>(hdr)[i] = reinterpret_cast<const uint32_t*>(plan)[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            // basis image: one bulk copy per 8 KiB slab, all on one barrier
            mbar_arrive_expect_tx(bbar, (uint32_t)n_slabs * kBSlabBytes);
            for (int s = 0; s < n_slabs; ++s)
                bulk_load_1d(smem_b + (size_t)s * kBSlabBytes, plan + kPlanHeaderBytes + (size_t)s * kBSlabBytes,
                             kBSlabBytes, bbar);
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int s0 = 0; s0 < n_slabs; ++s0) {
                    const int s = slab_at(s0, rot, n_slabs);
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], kStageBytes);
                    if (n_files > 0) {
                        const int64_t f = tile / tiles_per_file, t0 = (tile - f * tiles_per_file) * kTileRows;
                        tma_load_3d(smem_a + (size_t)stage * kStageBytes, &tmap, s * kSlabBytes, (int)t0, (int)f,
                                    &full[stage]);
                    } else {
                        tma_load_2d(smem_a + (size_t)stage * kStageBytes, &tmap, s * kSlabBytes,
                                    (int)(tile * kTileRows), &full[stage]);
                    }
                    if (++stage == n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        mbar_wait(bbar, 0);
        const uint32_t idesc = k2_idesc();
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        int slot = 0;
        uint32_t slot_phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kN);
            for (int s0 = 0; s0 < n_slabs; ++s0) {
                const int s = slab_at(s0, rot, n_slabs);
                if constexpr (TS) {
                    mbar_wait(&aready[slot], slot_phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_tmem = tmem_base + (uint32_t)(2 * kN + slot * 32);
                        const uint32_t b_addr = smem_u32(smem_b + (size_t)s * kBSlabBytes);
#pragma unroll
                        for (int k = 0; k < kSlabBytes / 32; ++k)
                            umma_i8_ts(d_tmem, a_tmem + k * 8, umma_desc_sw128(b_addr + k * 32), idesc,
                                       (s0 > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&aempty[slot]);
                        if (s0 == n_slabs - 1) umma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++slot == a_slots) {
                        slot = 0;
                        slot_phase ^= 1;
                    }
                    continue;
                }
                mbar_wait(&ready[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(smem_a + (size_t)stage * kStageBytes);
                    const uint32_t b_addr = smem_u32(smem_b + (size_t)s * kBSlabBytes);
#pragma unroll
                    for (int k = 0; k < kSlabBytes / 32; ++k)
                        umma_i8(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                                (s0 > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty[stage]);
                    if (s0 == n_slabs - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++stage == n_stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else if (warp < 2 + fix_warps) {
        // ===================== fix-up: hi byte -> offset binary =====================
        const int t = threadIdx.x - 64;  // 0..32*fix_warps-1
        constexpr int fix_threads = FIX_WARPS * 32;
        int stage = 0;
        uint32_t phase = 0;
        int slot = 0;
        uint32_t slot_phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int s = 0; s < n_slabs; ++s) {
                mbar_wait(&full[stage], phase);
                if constexpr (TS) {
                    // this warp may touch TMEM lanes 32 (warp % 4) ...; two warps share a lane quarter and split the slab
                    // eight warps: two share a lane quarter and split the slab; four warps: one per quarter, both halves
                    constexpr int kHalves = FIX_WARPS == 8 ? 1 : 2;
                    const int q = warp & 3, h0 = FIX_WARPS == 8 ? ((warp - 2) >> 2) : 0, r = q * 32 + lane;
                    const unsigned char* row = smem_a + (size_t)stage * kStageBytes + (size_t)r * kSlabBytes;
                    uint32_t v[kHalves][16];
#pragma unroll
                    for (int hh = 0; hh < kHalves; ++hh)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {   // logical 16-byte chunk 4h + c of row r sits at chunk ^ (r & 7)
                            const uint4 w = *reinterpret_cast<const uint4*>(row + (((4 * (h0 + hh) + c) ^ (r & 7)) << 4));
                            v[hh][4 * c + 0] = w.x ^ 0x80008000u;
                            v[hh][4 * c + 1] = w.y ^ 0x80008000u;
                            v[hh][4 * c + 2] = w.z ^ 0x80008000u;
                            v[hh][4 * c + 3] = w.w ^ 0x80008000u;
                        }
                    mbar_wait(&aempty[slot], slot_phase ^ 1);
                    tc_fence_after();
#pragma unroll
                    for (int hh = 0; hh < kHalves; ++hh)
                        tmem_st16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * kN + slot * 32 + (h0 + hh) * 16),
                                  v[hh]);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&empty[stage]);   // the stage has been read (its values went through the XORs above)
                        mbar_arrive(&aready[slot]);
                    }
                    if (++stage == n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                    if (++slot == a_slots) {
                        slot = 0;
                        slot_phase ^= 1;
                    }
                    continue;
                }
                uint4* base = reinterpret_cast<uint4*>(smem_a + (size_t)stage * kStageBytes);
#pragma unroll
                for (int i = 0; i < kStageBytes / 16 / fix_threads; ++i) {
                    uint4 v = base[i * fix_threads + t];
                    v.x ^= 0x80008000u;
                    v.y ^= 0x80008000u;
                    v.z ^= 0x80008000u;
                    v.w ^= 0x80008000u;
                    base[i * fix_threads + t] = v;
                }
                fence_proxy_async();  // make the generic-proxy writes visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[stage]);
                if (++stage == n_stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kN);
            // two halves of 8 basis columns x 4 digit slices: 32 live accumulator registers instead of 64,
            // which keeps the kernel small enough for a detect CTA of the previous batch to share the SM
            double eb = 0.0, en = 0.0;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                int32_t v[32];
                tmem_ld8(taddr + 0 + half * 8, v + 0);
                tmem_ld8(taddr + 16 + half * 8, v + 8);
                tmem_ld8(taddr + 32 + half * 8, v + 16);
                tmem_ld8(taddr + 48 + half * 8, v + 24);
                tmem_ld_wait();
                if (half == 1) {   // the whole accumulator is in registers: release it to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);
                }
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const int c = half * 8 + c8;
                    const int g = hdr->group[c];
                    // exact: |V| < 2^53
                    double V = (double)(v[c8] - hdr->offs[c]);
                    V = V * 256.0 + (double)(v[8 + c8] - hdr->offs[16 + c]);
                    V = V * 256.0 + (double)(v[16 + c8] - hdr->offs[32 + c]);
                    V = V * 256.0 + (double)(v[24 + c8] - hdr->offs[48 + c]);
                    const double X = V * hdr->inv_scale;
                    const double p2 = X * X;
                    if (g == 0) eb += p2;
                    if (g == 1) en += p2;
                }
            }
            int64_t row, orow;
            if (n_files > 0) {
                const int64_t f = tile / tiles_per_file;
                row = (tile - f * tiles_per_file) * kTileRows + q * 32 + lane;
                orow = f * out_stride + row;
            } else {
                row = tile * kTileRows + q * 32 + lane;
                orow = row;
            }
            if (row < n_rows) {
                out_band_db[orow] = (float)(10.0 * log10(eb + 1e-12));    // main.py:383-384
                out_noise_db[orow] = (float)(10.0 * log10(en + 1e-12));   // main.py:387-388
                if (out_band_e) out_band_e[orow] = (float)eb;
                if (out_noise_e) out_noise_e[orow] = (float)en;
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TS ? kTmemColsTS : kTmemCols)
                     : "memory");
    }
}

// ------------------------------------------------------------------ host side
inline int n_slabs_for(int k_samples) { return (2 * k_samples + kSlabBytes - 1) / kSlabBytes; }

}  // namespace
}  // namespace ms

namespace ms {
int band_power_i16_tc_impl(const int16_t* x, int64_t n_rows, int64_t row_stride_bytes, const void* d_plan,
                           int32_t k_samples, int32_t n_cols, float* out_band_db, float* out_noise_db,
                           float* out_band_energy, float* out_noise_energy, int32_t* zero_buf, int32_t zero_count,
                           void* stream, int64_t n_files = 0, int64_t file_stride_bytes = 0, int64_t out_stride = 0,
                           int fix_warps_req = 0, int grid_req = 0);
}

extern "C" {

int64_t ms_dft_i8_plan_bytes(int32_t k_samples, int32_t n_cols) {
    if (k_samples <= 0 || n_cols <= 0 || n_cols > ms::kCols) return 0;
    return ms::kPlanHeaderBytes + (int64_t)ms::n_slabs_for(k_samples) * ms::kBSlabBytes;
}

int ms_dft_i8_plan_build(const double* h_basis, const int32_t* h_col_group, int32_t k_samples, int32_t n_cols,
                         void* d_plan, void* stream) {
    using namespace ms;
    MS_REQUIRE(h_basis && h_col_group && d_plan, MS_ERR_INVALID_ARG, "ms_dft_i8_plan_build: null pointer");
    MS_REQUIRE(k_samples > 0 && n_cols > 0 && n_cols <= kCols, MS_ERR_UNSUPPORTED,
               "ms_dft_i8_plan_build: need 1 <= n_cols <= %d (got %d)", kCols, n_cols);
    const int n_slabs = n_slabs_for(k_samples);
    MS_REQUIRE(SmemLayout::stages_for(n_slabs) >= kMinStages, MS_ERR_UNSUPPORTED,
               "ms_dft_i8_plan_build: k_samples=%d needs %zu bytes of shared memory (> 227 KiB)", k_samples,
               SmemLayout::bytes(n_slabs, kMinStages));
    const int64_t total = ms_dft_i8_plan_bytes(k_samples, n_cols);
    std::vector<unsigned char> img((size_t)total, 0);
    PlanHeader* h = reinterpret_cast<PlanHeader*>(img.data());
    h->magic = kPlanMagic;
    h->k_samples = k_samples;
    h->n_cols = n_cols;
    h->n_slabs = n_slabs;
    for (int c = 0; c < kCols; ++c) h->group[c] = (c < n_cols) ? h_col_group[c] : -1;
    int64_t dsum[3][kCols] = {};
    unsigned char* B = img.data() + kPlanHeaderBytes;
    // element (accumulator column j = slice*16 + c, K byte kb) of the K-major SWIZZLE_128B image:
    //   slab = kb/128, within a slab row j sits at (j/8)*1024 + (j%8)*128 and its 16-byte chunk
    //   (kb%128)/16 is stored at chunk ^ (j%8)
    auto put = [&](int j, int kb, int8_t val) {
        const int slab = kb / kSlabBytes, kin = kb % kSlabBytes;
        const int chunk = (kin >> 4) ^ (j & 7);
        B[(size_t)slab * kBSlabBytes + (size_t)(j >> 3) * 1024 + (size_t)(j & 7) * 128 + chunk * 16 + (kin & 15)] =
            (unsigned char)val;
    };
    // the basis is normalised so that its largest magnitude uses the full three-digit range: a window that is
    // cropped far below 1 (block much longer than nfft) keeps the same relative quantisation as a full-scale one
    double peak = 0.0;
    for (size_t i = 0; i < (size_t)k_samples * n_cols; ++i) {
        MS_REQUIRE(isfinite(h_basis[i]), MS_ERR_INVALID_ARG, "ms_dft_i8_plan_build: non-finite basis value");
        peak = fmax(peak, fabs(h_basis[i]));
    }
    MS_REQUIRE(peak <= 1.0, MS_ERR_INVALID_ARG, "ms_dft_i8_plan_build: basis value %g outside [-1, 1]", peak);
    const double scale = peak > 0.0 ? kBasisPeak * (double)(1 << kFracBits) / peak : 1.0;
    h->inv_scale = 1.0 / scale;
    for (int n = 0; n < k_samples; ++n) {
        for (int c = 0; c < n_cols; ++c) {
            const double b = h_basis[(size_t)n * n_cols + c];
            const long long v = llrint(b * scale);
            const int q3 = (int)(((v + 128) & 255) - 128);
            const long long v1 = (v - q3) / 256;
            const int q2 = (int)(((v1 + 128) & 255) - 128);
            const int q1 = (int)((v1 - q2) / 256);
            // lo byte (K index 2n): slices 1,2,3 ; hi byte (2n+1): slices 0,1,2
            put(1 * 16 + c, 2 * n, (int8_t)q1);
            put(2 * 16 + c, 2 * n, (int8_t)q2);
            put(3 * 16 + c, 2 * n, (int8_t)q3);
            put(0 * 16 + c, 2 * n + 1, (int8_t)q1);
            put(1 * 16 + c, 2 * n + 1, (int8_t)q2);
            put(2 * 16 + c, 2 * n + 1, (int8_t)q3);
            dsum[0][c] += q1;
            dsum[1][c] += q2;
            dsum[2][c] += q3;
        }
    }
    for (int c = 0; c < kCols; ++c) {
        h->offs[0 * 16 + c] = (int32_t)(128 * dsum[0][c]);
        h->offs[1 * 16 + c] = (int32_t)(128 * dsum[1][c]);
        h->offs[2 * 16 + c] = (int32_t)(128 * dsum[2][c]);
        h->offs[3 * 16 + c] = 0;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MS_CUDA_OK(cudaMemcpyAsync(d_plan, img.data(), (size_t)total, cudaMemcpyHostToDevice, st));
    MS_CUDA_OK(cudaStreamSynchronize(st));  // img goes out of scope
    return MS_OK;
}

int ms_band_power_i16_tc(const int16_t* x, int64_t n_rows, int64_t row_stride_bytes, const void* d_plan,
                         int32_t k_samples, int32_t n_cols, float* out_band_db, float* out_noise_db,
                         float* out_band_energy, float* out_noise_energy, void* stream) {
    return ms::band_power_i16_tc_impl(x, n_rows, row_stride_bytes, d_plan, k_samples, n_cols, out_band_db,
                                      out_noise_db, out_band_energy, out_noise_energy, nullptr, 0, stream);
}

int ms_band_power_i16_tc_batched(const int16_t* x, int64_t n_files, int64_t file_stride_bytes, int64_t n_frames,
                                 int64_t row_stride_bytes, const void* d_plan, int32_t k_samples, int32_t n_cols,
                                 int64_t out_stride, float* out_band_db, float* out_noise_db, float* out_band_energy,
                                 float* out_noise_energy, void* stream) {
    MS_REQUIRE(n_files > 0, MS_ERR_INVALID_ARG, "ms_band_power_i16_tc_batched: n_files must be positive");
    return ms::band_power_i16_tc_impl(x, n_frames, row_stride_bytes, d_plan, k_samples, n_cols, out_band_db,
                                      out_noise_db, out_band_energy, out_noise_energy, nullptr, 0, stream, n_files,
                                      file_stride_bytes, out_stride);
}

}  // extern "C"

namespace ms {
int band_power_i16_tc_impl(const int16_t* x, int64_t n_rows, int64_t row_stride_bytes, const void* d_plan,
                           int32_t k_samples, int32_t n_cols, float* out_band_db, float* out_noise_db,
                           float* out_band_energy, float* out_noise_energy, int32_t* zero_buf, int32_t zero_count,
                           void* stream, int64_t n_files, int64_t file_stride_bytes, int64_t out_stride, int fix_warps_req,
                           int grid_req) {
    MS_REQUIRE(x && d_plan && out_band_db && out_noise_db, MS_ERR_INVALID_ARG, "ms_band_power_i16_tc: null pointer");
    MS_REQUIRE(n_files >= 0 && n_files < ((int64_t)1 << 31), MS_ERR_INVALID_ARG, "ms_band_power_i16_tc: bad n_files");
    if (n_files > 0)
        MS_REQUIRE(file_stride_bytes > 0 && file_stride_bytes % 16 == 0 && out_stride >= n_rows, MS_ERR_UNSUPPORTED,
                   "ms_band_power_i16_tc_batched: file stride %lld bytes must be a positive multiple of 16 and "
                   "out_stride >= frames per file", (long long)file_stride_bytes);
    MS_REQUIRE(n_rows >= 0 && n_rows < ((int64_t)1 << 31), MS_ERR_INVALID_ARG, "ms_band_power_i16_tc: bad n_rows");
    MS_REQUIRE(row_stride_bytes > 0 && row_stride_bytes % 16 == 0, MS_ERR_UNSUPPORTED,
               "ms_band_power_i16_tc: row stride %lld bytes is not a multiple of 16 (TMA); use ms_band_power_i16",
               (long long)row_stride_bytes);
    MS_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, MS_ERR_UNSUPPORTED,
               "ms_band_power_i16_tc: x must be 16-byte aligned");
    MS_REQUIRE(k_samples > 0 && n_cols > 0 && n_cols <= kCols, MS_ERR_UNSUPPORTED, "ms_band_power_i16_tc: bad plan shape");
    if (n_rows == 0) {
        if (zero_buf && zero_count > 0)
            MS_CUDA_OK(cudaMemsetAsync(zero_buf, 0, sizeof(int32_t) * (size_t)zero_count, static_cast<cudaStream_t>(stream)));
        return MS_OK;
    }
    const int n_slabs = n_slabs_for(k_samples);
    const int n_stages = SmemLayout::stages_for(n_slabs);
    MS_REQUIRE(n_stages >= kMinStages, MS_ERR_UNSUPPORTED, "ms_band_power_i16_tc: k_samples too large for shared memory");
    const size_t smem = SmemLayout::bytes(n_slabs, n_stages);

    EncodeTiledFn encode = get_encode_fn();
    MS_REQUIRE(encode != nullptr, MS_ERR_CUDA, "ms_band_power_i16_tc: cuTensorMapEncodeTiled unavailable");
    CUtensorMap tmap;
    // inner extent = the frame's real bytes; rows may overlap (hop < frame) or leave gaps (hop > frame),
    // and whatever a 128-byte box covers beyond this extent is zero-filled by TMA
    const int64_t row_bytes = (int64_t)k_samples * 2;
    const cuuint64_t gdim[3] = {(cuuint64_t)row_bytes, (cuuint64_t)n_rows, (cuuint64_t)(n_files > 0 ? n_files : 1)};
    const cuuint64_t gstride[2] = {(cuuint64_t)row_stride_bytes, (cuuint64_t)(n_files > 0 ? file_stride_bytes : 16)};
    const cuuint32_t box[3] = {(cuuint32_t)kSlabBytes, (cuuint32_t)kTileRows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t rank = n_files > 0 ? 3 : 2;
    static const int slab_rot = [] {  // tuning knob: MS_K2_SLAB_ROT = per-CTA rotation step of the slab order (0 = none)
        const char* e = getenv("MS_K2_SLAB_ROT");
        return e ? atoi(e) : 5;
    }();
    static const int fix_warps_default = [] {  // tuning knob: MS_K2_FIX_WARPS = 4 or 8
        const char* e = getenv("MS_K2_FIX_WARPS");
        const int v = e ? atoi(e) : kFixWarpsDefault;
        return v == 8 ? 8 : 4;
    }();
    // 8 warps shorten the time a landed stage waits for its fix-up (+2 % on the dense layout); the overlapped pass asks
    // for 4 so that a small-footprint detect CTA still fits into the register file next to this kernel
    const int fix_warps = fix_warps_req == 4 || fix_warps_req == 8 ? fix_warps_req : fix_warps_default;
    static const int l2promo = [] {   // tuning knob: MS_TMA_L2PROMO = 0 none, 1 64B, 2 128B, 3 256B
        const char* e = getenv("MS_TMA_L2PROMO");
        return e ? atoi(e) : 2;
    }();
    const CUtensorMapL2promotion promo = l2promo == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                         : l2promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : l2promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                        : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    // steady-state batches re-use the same buffer and geometry: the encoded map of the last call is kept per thread
    struct MapKey {
        const void* x;
        int64_t n_rows, row_stride, n_files, file_stride;
        int32_t k_samples, promo;
        bool operator==(const MapKey& o) const {
            return x == o.x && n_rows == o.n_rows && row_stride == o.row_stride && n_files == o.n_files &&
                   file_stride == o.file_stride && k_samples == o.k_samples && promo == o.promo;
        }
    };
    static thread_local MapKey cached_key = {nullptr, 0, 0, 0, 0, 0, 0};
    static thread_local CUtensorMap cached_map;
    const MapKey key = {x, n_rows, row_stride_bytes, n_files, file_stride_bytes, k_samples, l2promo};
    if (cached_key == key) {
        tmap = cached_map;
    } else {
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, const_cast<int16_t*>(x), gdim, gstride, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MS_REQUIRE(r == CUDA_SUCCESS, MS_ERR_CUDA, "ms_band_power_i16_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
        cached_map = tmap;
        cached_key = key;
    }

    // once per device: the limit is the per-SM maximum, every launch passes its own size
    static thread_local int attr_dev = -1;
    int cur_dev = 0;
    MS_CUDA_OK(cudaGetDevice(&cur_dev));
    if (attr_dev != cur_dev) {
        MS_CUDA_OK(cudaFuncSetAttribute(dft_i8_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MS_CUDA_OK(cudaFuncSetAttribute(dft_i8_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MS_CUDA_OK(cudaFuncSetAttribute(dft_i8_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MS_CUDA_OK(cudaFuncSetAttribute(dft_i8_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_dev = cur_dev;
    }
    const int64_t n_tiles = ((n_rows + kTileRows - 1) / kTileRows) * (n_files > 0 ? n_files : 1);
    int64_t grid = num_sms();
    if (grid_req > 0 && grid_req < grid) grid = grid_req;   // overlapped pass: leave a few SMs to the detect kernel
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) grid = 1;
    static const bool ts_form = [] {   // tuning knob: MS_K2_TS=0 = A operand from shared memory (rewritten in place)
        const char* e = getenv("MS_K2_TS");
        return !(e && e[0] == '0');
    }();
    static const int a_slots = [] {   // tuning knob: MS_K2_TS_SLOTS = depth of the tensor-memory operand ring (2..12)
        const char* e = getenv("MS_K2_TS_SLOTS");
        const int v = e ? atoi(e) : kASlots;
        return v < 2 ? 2 : (v > kASlots ? kASlots : v);
    }();
    if (fix_warps == 8 && ts_form)
        dft_i8_kernel<8, true><<<(unsigned)grid, 64 + 32 * 8 + 128, smem, static_cast<cudaStream_t>(stream)>>>(
            tmap, static_cast<const unsigned char*>(d_plan), n_rows, n_files, out_stride, n_slabs, out_band_db, out_noise_db,
            out_band_energy, out_noise_energy, zero_buf, zero_count, n_stages, slab_rot, a_slots);
    else if (fix_warps == 8)
        dft_i8_kernel<8, false><<<(unsigned)grid, 64 + 32 * 8 + 128, smem, static_cast<cudaStream_t>(stream)>>>(
            tmap, static_cast<const unsigned char*>(d_plan), n_rows, n_files, out_stride, n_slabs, out_band_db, out_noise_db,
            out_band_energy, out_noise_energy, zero_buf, zero_count, n_stages, slab_rot, a_slots);
    else if (ts_form)
        dft_i8_kernel<4, true><<<(unsigned)grid, 64 + 32 * 4 + 128, smem, static_cast<cudaStream_t>(stream)>>>(
            tmap, static_cast<const unsigned char*>(d_plan), n_rows, n_files, out_stride, n_slabs, out_band_db, out_noise_db,
            out_band_energy, out_noise_energy, zero_buf, zero_count, n_stages, slab_rot, a_slots);
    else
        dft_i8_kernel<4, false><<<(unsigned)grid, 64 + 32 * 4 + 128, smem, static_cast<cudaStream_t>(stream)>>>(
            tmap, static_cast<const unsigned char*>(d_plan), n_rows, n_files, out_stride, n_slabs, out_band_db, out_noise_db,
            out_band_energy, out_noise_energy, zero_buf, zero_count, n_stages, slab_rot, a_slots);
    MS_CUDA_OK(cudaGetLastError());
    return MS_OK;
}
}  // namespace ms
