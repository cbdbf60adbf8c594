// K2S: the general form of the tensor-core restricted DFT (tcgen05.mma kind::i8, exact integer arithmetic; number
// format as in ms_dft_i8.cu) for what K2 cannot hold: frames longer than its resident basis, bands of up to 32 bins
// per launch (64 cos/sin columns, UMMA N up to 256), and OVERLAPPING frames whose samples are read from HBM once.
// Reference semantics unchanged: dsp/src/main.py:376-388 (windowed rfft -> |X|^2 -> band sums -> 10*log10(.+1e-12))
// and the spectrogram call form of dsp/src/main.py:52-54, 132-133 (nperseg = nfft, noverlap > 0).
//
// Segment form.  With hop H and frame length L = (R-1)*H + r (0 < r <= H), frame f is the concatenation of R hop
// segments f, f+1, .., f+R-1 (the last one partial).  Its DFT bin is
//     X_f[k] = sum_{j<R} sum_{i<H} x[(f+j)*H + i] * b_j[i][k],     b_j[i][k] = w[j*H+i] * e^{-2 pi i k (j*H+i)/nfft}
// (b = 0 beyond the frame).  The audio is addressed as the matrix A[s][i] of NON-overlapping hop segments (each
// sample is fetched once), and the j-th term is the product of A *shifted down by j rows* with the j-th slice of the
// basis: the accumulator row f collects sum_j A[f+j] * B_j inside TMEM.  The shift costs nothing: a tile is staged
// as 128*T + R-1 consecutive segment rows (128-byte swizzled, TMA), and the UMMA shared-memory descriptor of shift j
// simply starts j rows (j*128 bytes) further down.  Measured on B200: the 128-byte swizzle is a function of the
// ABSOLUTE shared-memory address bits, so the shifted view is consistent with what TMA wrote and the descriptor's
// base-offset field must stay 0 (with the PTX ISA's (addr >> 7) & 7 in it every shifted product came out wrong).  R = 1 (hop >= frame, or `direct` mode: rows = frames) degenerates to K2's scheme with a streamed basis.
//
// Pipeline per CTA (persistent, one CTA per SM); a "pass" = T row tiles of 128 frames sharing every basis piece:
//   warp 0      A producer: per K slab (128 bytes of a row) T boxes of 128 rows + one halo box -> A ring
//   warp 2      B producer: basis pieces (slab s, shift j), [N x 128 B] each -> B ring (or all resident, loaded once)
//   warps 4-11  fix-up: XOR 0x80 into the hi bytes of the landed A stage (two's complement -> offset binary)
//   warp 1      MMA issuer: per piece T x 4 UTCIMMA (M128, N = 4*nc, K32) into T accumulators
//   warps 12-19 epilogue (two per TMEM lane quarter): tcgen05.ld, exact recombination of the 4 digit slices, |X|^2,
//               band sums, dB
// Bands wider than 32 bins run as several launches (column groups) accumulating fp64 energies (`first`/`last`).
#include <cuda.h>

#include <stdlib.h>

#include <vector>

#include "ms_async.cuh"
#include "ms_common.cuh"
#include "ms_umma.cuh"

namespace ms {
namespace {

constexpr int kTileRows = 128;
constexpr int kSlabBytes = 128;
constexpr int kMaxNc = 64;                 // basis columns per launch (cos/sin pairs of up to 32 bins)
constexpr int kMaxAStages = 6;
constexpr int kMaxBStages = 8;
constexpr int kBarBytes = 512;
constexpr int kPartBytes = 2 * kTileRows * 16;   // band-sum partials of the second epilogue warp of a lane quarter [parity][row]
constexpr int kHdrBytes = 2048;
constexpr uint32_t kMagic = 0x4d535347u;   // "MSSG"
constexpr int kFracBits = 23;
constexpr double kBasisPeak = 0.99;
constexpr int kThreads = 640;              // A producer, MMA, B producer, (idle), 8 fix-up warps, 8 epilogue warps
constexpr int kFixWarps = 8;
constexpr size_t kSmemBudget = (size_t)227 * 1024;

struct SegHeader {
    uint32_t magic;
    int32_t n_frame;      // samples per frame entering the transform
    int32_t seg_samples;  // H (== n_frame in direct mode)
    int32_t n_shift;      // R
    int32_t n_slabs;      // ceil(2*H / 128)
    int32_t nc;           // padded basis columns (multiple of 8, 16..64); UMMA N = 4*nc
    int32_t n_cols;       // real basis columns
    int32_t pad;
    double inv_scale;
    int32_t group[kMaxNc];      // 0 signal band, 1 noise band, -1 unused
    int64_t off64[kMaxNc];      // offset-binary correction of a column's four digit slices, already combined:
                                // 128 * (sum q1 * 2^24 + sum q2 * 2^16 + sum q3 * 2^8)
};
static_assert(sizeof(SegHeader) <= kHdrBytes, "header too large");

// tcgen05.mma kind::i8 with both descriptors given as (low word, shared high word): the low word carries the
// 16-byte-granular start address, so advancing an operand is one 32-bit add.
__device__ __forceinline__ void umma_i8_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                             uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

struct Cfg {
    int T, n_a, n_b, resident, halo_rows, a_rows, n_acc, tmem_cols;
    size_t smem;
};

__global__ void __launch_bounds__(kThreads, 1)
dft_seg_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_halo,
               const unsigned char* __restrict__ plan, int64_t n_rows, int64_t n_files, int64_t out_stride,
               int64_t out_offset, float* __restrict__ out_band_db, float* __restrict__ out_noise_db,
               double* __restrict__ acc_band, double* __restrict__ acc_noise, int first, int last, int T, int n_a,
               int n_b, int resident, int halo_rows, int n_acc, int tmem_cols, double* __restrict__ out_raw,
               int raw_cols, int raw_col0) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) {
        if (threadIdx.x == 0) printf("dft_seg_kernel: dynamic shared memory base is not 1 KiB aligned\n");
        __trap();
    }
    const SegHeader* ghdr = reinterpret_cast<const SegHeader*>(plan);
    const int n_slabs = ghdr->n_slabs, R = ghdr->n_shift, nc = ghdr->nc;
    const int N = 4 * nc;
    const int piece_bytes = N * kSlabBytes;
    const int n_pieces = n_slabs * R;
    const int a_rows = T * kTileRows + halo_rows;
    const int a_stage_bytes = a_rows * kSlabBytes;
    unsigned char* smem_b = smem;
    unsigned char* smem_a = smem_b + (size_t)(resident ? n_pieces : n_b) * piece_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)n_a * a_stage_bytes);
    uint64_t* a_full = bars;
    uint64_t* a_ready = a_full + kMaxAStages;
    uint64_t* a_empty = a_ready + kMaxAStages;
    uint64_t* b_full = a_empty + kMaxAStages;
    uint64_t* b_empty = b_full + kMaxBStages;
    uint64_t* tfull = b_empty + kMaxBStages;
    uint64_t* tempty = tfull + 2;
    uint64_t* bres = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres + 1);
    SegHeader* hdr = reinterpret_cast<SegHeader*>(reinterpret_cast<unsigned char*>(bars) + kBarBytes);
    double2* part = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(hdr) + kHdrBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t rows_per_pass = (int64_t)T * kTileRows;
    const int64_t ppf = (n_rows + rows_per_pass - 1) / rows_per_pass;   // passes per file
    const int64_t n_pass = n_files * ppf;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_a; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_ready[s], kFixWarps);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < n_b; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 8);
        }
        mbar_init(bres, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < (int)(sizeof(SegHeader) / 4); i += (int)blockDim.x)
        reinterpret_cast<uint32_t*>(hdr)[i] = reinterpret_cast<const uint32_t*>(plan)[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== A producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t p = blockIdx.x; p < n_pass; p += gridDim.x) {
                const int f = (int)(p / ppf);
                const int row0 = (int)((p - (int64_t)f * ppf) * rows_per_pass);
                for (int s = 0; s < n_slabs; ++s) {
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&a_full[stage], (uint32_t)a_stage_bytes);
                    unsigned char* dst = smem_a + (size_t)stage * a_stage_bytes;
                    for (int t = 0; t < T; ++t)
                        tma_load_3d(dst + (size_t)t * kTileRows * kSlabBytes, &tmap, s * kSlabBytes,
                                    row0 + t * kTileRows, f, &a_full[stage]);
                    if (halo_rows > 0)
                        tma_load_3d(dst + (size_t)T * kTileRows * kSlabBytes, &tmap_halo, s * kSlabBytes,
                                    row0 + T * kTileRows, f, &a_full[stage]);
                    if (++stage == n_a) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== B producer =====================
        if (lane == 0) {
            const unsigned char* src = plan + kHdrBytes;
            if (resident) {
                mbar_arrive_expect_tx(bres, (uint32_t)n_pieces * (uint32_t)piece_bytes);
                for (int i = 0; i < n_pieces; ++i)
                    bulk_load_1d(smem_b + (size_t)i * piece_bytes, src + (size_t)i * piece_bytes, (uint32_t)piece_bytes,
                                 bres);
            } else {
                int stage = 0;
                uint32_t phase = 0;
                for (int64_t p = blockIdx.x; p < n_pass; p += gridDim.x) {
                    for (int i = 0; i < n_pieces; ++i) {
                        mbar_wait(&b_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&b_full[stage], (uint32_t)piece_bytes);
                        bulk_load_1d(smem_b + (size_t)stage * piece_bytes, src + (size_t)i * piece_bytes,
                                     (uint32_t)piece_bytes, &b_full[stage]);
                        if (++stage == n_b) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp runs the loop nest with warp-uniform values (descriptors, stages, TMEM addresses live in
        // uniform registers) and one elected lane issues; every descriptor is a 32-bit add on a precomputed low word.
        // With R*T*4 instructions per K slab the issue loop itself is the critical path (first version: ~22
        // instructions and several R2UR moves per UTCIMMA; the tensor pipe was 12 % busy and HBM at 26 %).
        {
            if (resident) mbar_wait(bres, 0);
            const bool leader = elect_one();
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t idesc = umma_idesc_i8(N, kTileRows);
            const uint64_t d0 = umma_desc_sw128(0);
            const uint32_t desc_hi = (uint32_t)(d0 >> 32);
            const uint32_t a_lo0 = (uint32_t)d0 + ((smem_u32(smem_a) & 0x3FFFFu) >> 4);
            const uint32_t b_lo0 = (uint32_t)d0 + ((smem_u32(smem_b) & 0x3FFFFu) >> 4);
            const uint32_t a_stage16 = (uint32_t)a_stage_bytes >> 4, piece16 = (uint32_t)piece_bytes >> 4;
            const uint32_t tile16 = (uint32_t)(kTileRows * kSlabBytes) >> 4, row16 = (uint32_t)kSlabBytes >> 4;
            const int last_kc = (2 * ghdr->seg_samples - (n_slabs - 1) * kSlabBytes + 31) / 32;
            int astage = 0, bstage = 0, acc = 0;
            uint32_t aphase = 0, bphase = 0, acc_phase = 0;
            for (int64_t p = blockIdx.x; p < n_pass; p += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_u + (uint32_t)(acc * T * N);
                uint32_t b_res = b_lo0;                      // resident basis: pieces in (slab, shift) order
                for (int s = 0; s < n_slabs; ++s) {
                    mbar_wait(&a_ready[astage], aphase);
                    tc_fence_after();
                    const uint32_t a_s = a_lo0 + (uint32_t)astage * a_stage16;
                    const int kc = (s == n_slabs - 1) ? last_kc : 4;   // 32-byte K chunks holding real row bytes
                    for (int j = 0; j < R; ++j) {
                        uint32_t b_p;
                        if (resident) {
                            b_p = b_res;
                            b_res += piece16;
                        } else {
                            mbar_wait(&b_full[bstage], bphase);
                            tc_fence_after();
                            b_p = b_lo0 + (uint32_t)bstage * piece16;
                        }
                        const uint32_t accf = (s | j) != 0 ? 1u : 0u;
                        uint32_t a_t = a_s + (uint32_t)j * row16;
                        uint32_t d_t = d_tmem;
                        for (int t = 0; t < T; ++t) {
                            if (leader) {
                                umma_i8_lohi(d_t, a_t, b_p, desc_hi, idesc, accf);
                                if (kc > 1) umma_i8_lohi(d_t, a_t + 2, b_p + 2, desc_hi, idesc, 1u);
                                if (kc > 2) umma_i8_lohi(d_t, a_t + 4, b_p + 4, desc_hi, idesc, 1u);
                                if (kc > 3) umma_i8_lohi(d_t, a_t + 6, b_p + 6, desc_hi, idesc, 1u);
                            }
                            a_t += tile16;
                            d_t += (uint32_t)N;
                        }
                        if (!resident) {
                            if (leader) umma_commit(&b_empty[bstage]);
                            if (++bstage == n_b) {
                                bstage = 0;
                                bphase ^= 1;
                            }
                        }
                    }
                    if (leader) {
                        umma_commit(&a_empty[astage]);
                        if (s == n_slabs - 1) umma_commit(&tfull[acc]);
                    }
                    if (++astage == n_a) {
                        astage = 0;
                        aphase ^= 1;
                    }
                }
                if (++acc == n_acc) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4 && warp < 4 + kFixWarps) {
        // ===================== fix-up: hi byte -> offset binary =====================
        const int t = threadIdx.x - 4 * 32;
        const int n16 = a_stage_bytes / 16;
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t p = blockIdx.x; p < n_pass; p += gridDim.x) {
            for (int s = 0; s < n_slabs; ++s) {
                mbar_wait(&a_full[stage], phase);
                uint4* base = reinterpret_cast<uint4*>(smem_a + (size_t)stage * a_stage_bytes);
#pragma unroll 4
                for (int i = t; i < n16; i += kFixWarps * 32) {
                    uint4 v = base[i];
                    v.x ^= 0x80008000u;
                    v.y ^= 0x80008000u;
                    v.z ^= 0x80008000u;
                    v.w ^= 0x80008000u;
                    base[i] = v;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[stage]);
                if (++stage == n_a) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp >= 12) {
        // ===================== epilogue =====================
        // Eight warps: two per TMEM lane quarter, `half` 0 takes the column chunks 0, 2, 4, .. (8 columns each) and
        // `half` 1 the chunks 1, 3, 5, ..  A row's work is a dependent chain (TMEM load -> integer recombination ->
        // one conversion -> square), so with passes of only a few K slabs the epilogue, not the MMAs, set the pace;
        // splitting the columns over two warps halves that chain.  Band mode: half 1 hands its partial band sums to
        // half 0 through shared memory (one named barrier per row tile, buffers alternate with the tile parity).
        const int q = warp & 3, half = (warp - 12) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0, tile_par = 0;
        const double inv_scale = hdr->inv_scale;
        const int n_chunks = nc / 8;
        for (int64_t p = blockIdx.x; p < n_pass; p += gridDim.x) {
            const int64_t f = p / ppf;
            const int64_t row0 = (p - f * ppf) * rows_per_pass;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            for (int t = 0; t < T; ++t) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * T * N + t * N);
                const int64_t row = row0 + (int64_t)t * kTileRows + q * 32 + lane;
                double eb = 0.0, en = 0.0, x_even = 0.0;
                const int last_chunk = ((n_chunks - 1 - half) & ~1) + half;     // this warp's last chunk index (may be < half)
                for (int ch = half; ch < n_chunks; ch += 2) {
                    const int c0 = ch * 8;
                    int32_t v[32];
                    tmem_ld8(taddr + 0 * nc + c0, v + 0);
                    tmem_ld8(taddr + 1 * nc + c0, v + 8);
                    tmem_ld8(taddr + 2 * nc + c0, v + 16);
                    tmem_ld8(taddr + 3 * nc + c0, v + 24);
                    tmem_ld_wait();
                    if (t == T - 1 && ch == last_chunk) {   // this warp's last read of the pass
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[acc]);
                    }
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {
                        const int c = c0 + c8;
                        const int g = hdr->group[c];
                        // the four digit slices recombine exactly in 64-bit integers (|V| < 2^53), one conversion per
                        // column
                        const long long V = ((long long)v[c8] << 24) + ((long long)v[8 + c8] << 16) +
                                            ((long long)v[16 + c8] << 8) + (long long)v[24 + c8] - hdr->off64[c];
                        const double X = (double)V * inv_scale;
                        if (out_raw != nullptr) {      // projection mode: the column values themselves (fp64) leave,
                            // as (cos, sin) pairs: one 16-byte store per extended bin
                            if (c8 & 1) {
                                if (row < n_rows && c < hdr->n_cols)
                                    *reinterpret_cast<double2*>(
                                        &out_raw[((f * out_stride + out_offset + row) * raw_cols) + raw_col0 + c - 1]) =
                                        make_double2(x_even, X);
                            } else {
                                x_even = X;
                            }
                            continue;
                        }
                        const double p2 = X * X;
                        if (g == 0) eb += p2;
                        if (g == 1) en += p2;
                    }
                }
                if (t == T - 1 && last_chunk < half) {      // fewer chunks than warps: nothing to read, still hand back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);
                }
                if (out_raw == nullptr) {
                    double2* px = part + (size_t)tile_par * kTileRows;
                    if (half == 1) px[q * 32 + lane] = make_double2(eb, en);
                    asm volatile("bar.sync 1, 256;" ::: "memory");           // the eight epilogue warps
                    if (half == 0 && row < n_rows) {
                        const double2 o = px[q * 32 + lane];
                        eb += o.x;
                        en += o.y;
                        const int64_t orow = f * out_stride + out_offset + row;
                        if (acc_band != nullptr) {   // one of several column groups: energies accumulate in fp64
                            if (!first) {
                                eb += acc_band[orow];
                                en += acc_noise[orow];
                            }
                            acc_band[orow] = eb;     // (after the last group: the linear energies, for callers that want them)
                            acc_noise[orow] = en;
                        }
                        if (last) {
                            out_band_db[orow] = (float)(10.0 * log10(eb + 1e-12));    // main.py:383-384
                            out_noise_db[orow] = (float)(10.0 * log10(en + 1e-12));   // main.py:387-388
                        }
                    }
                    tile_par ^= 1;
                }
            }
            if (++acc == n_acc) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------ host side
inline int n_slabs_for(int seg_samples) { return (2 * seg_samples + kSlabBytes - 1) / kSlabBytes; }
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Pick the tiling for one plan shape.  Returns false when nothing fits.
bool choose_cfg(int seg_samples, int n_shift, int nc, Cfg* c) {
    const int N = 4 * nc;
    const size_t piece = (size_t)N * kSlabBytes;
    const size_t n_pieces = (size_t)n_slabs_for(seg_samples) * n_shift;
    const size_t fixed = kBarBytes + kHdrBytes + kPartBytes;
    auto a_stage = [&](int T) { return (size_t)round_up(T * kTileRows + n_shift - 1, 8) * kSlabBytes; };
    if (n_shift - 1 > 128) return false;
    // basis resident (small frames): one row tile per pass, as many A stages as fit
    if (n_pieces * piece + 3 * a_stage(1) + fixed <= kSmemBudget) {
        c->T = 1;
        c->resident = 1;
        c->n_b = 1;
        c->n_a = (int)((kSmemBudget - fixed - n_pieces * piece) / a_stage(1));
        if (c->n_a > kMaxAStages) c->n_a = kMaxAStages;
    } else {
        // streamed basis: T row tiles share every piece (TMEM holds T accumulators of N columns)
        // measured (24 h sweep): two tiles per pass beat four (N = 128: 0.27 vs 0.31 ms at nfft 2048, 50 % overlap) --
        // a deeper ring of smaller A stages matters more than halving the basis traffic again
        int T = 512 / N;
        if (T > 2) T = 2;
        static const int t_req = [] {      // tuning knob: MS_SEG_T = row tiles per pass of the streamed-basis form
            const char* e = getenv("MS_SEG_T");
            return e ? atoi(e) : 0;
        }();
        if (t_req >= 1 && t_req < T) T = t_req;
        for (; T >= 1; --T)
            if (3 * piece + 2 * a_stage(T) + fixed <= kSmemBudget) break;
        if (T < 1) return false;
        c->T = T;
        c->resident = 0;
        c->n_a = 2;
        size_t left = kSmemBudget - fixed - 2 * a_stage(T);
        c->n_b = (int)(left / piece);
        if (c->n_b > kMaxBStages) c->n_b = kMaxBStages;
        // prefer a third A stage over more than 4 basis pieces in flight
        if (c->n_b > 4 && (size_t)4 * piece + 3 * a_stage(T) + fixed <= kSmemBudget) {
            c->n_a = 3;
            c->n_b = (int)((kSmemBudget - fixed - 3 * a_stage(T)) / piece);
            if (c->n_b > kMaxBStages) c->n_b = kMaxBStages;
        }
    }
    c->a_rows = round_up(c->T * kTileRows + n_shift - 1, 8);
    c->halo_rows = c->a_rows - c->T * kTileRows;
    c->n_acc = (2 * c->T * N <= 512) ? 2 : 1;
    int cols = 32;
    while (cols < c->n_acc * c->T * N) cols *= 2;
    c->tmem_cols = cols;
    c->smem = (c->resident ? n_pieces : (size_t)c->n_b) * piece + (size_t)c->n_a * c->a_rows * kSlabBytes + fixed;
    return c->smem <= kSmemBudget;
}

}  // namespace
}  // namespace ms

extern "C" {

int64_t ms_dft_seg_plan_bytes(int32_t n_frame, int32_t seg_samples, int32_t n_shift, int32_t n_cols) {
    using namespace ms;
    if (n_frame <= 0 || seg_samples <= 0 || n_shift <= 0 || n_cols <= 0 || n_cols > kMaxNc) return 0;
    if ((int64_t)seg_samples * n_shift < n_frame || (int64_t)seg_samples * (n_shift - 1) >= n_frame) return 0;
    const int nc = round_up(n_cols < 16 ? 16 : n_cols, 8);
    Cfg c;
    if (!choose_cfg(seg_samples, n_shift, nc, &c)) return 0;
    return kHdrBytes + (int64_t)n_slabs_for(seg_samples) * n_shift * (4 * nc) * kSlabBytes;
}

int ms_dft_seg_plan_build(const double* h_basis, const int32_t* h_col_group, int32_t n_frame, int32_t seg_samples,
                          int32_t n_shift, int32_t n_cols, void* d_plan, void* stream) {
    using namespace ms;
    MS_REQUIRE(h_basis && h_col_group && d_plan, MS_ERR_INVALID_ARG, "ms_dft_seg_plan_build: null pointer");
    const int64_t total = ms_dft_seg_plan_bytes(n_frame, seg_samples, n_shift, n_cols);
    MS_REQUIRE(total > 0, MS_ERR_UNSUPPORTED,
               "ms_dft_seg_plan_build: unsupported shape (frame %d, segment %d x %d, %d columns; need 1..%d columns, "
               "(n_shift-1)*segment < frame <= n_shift*segment, n_shift <= 129)",
               n_frame, seg_samples, n_shift, n_cols, kMaxNc);
    const int nc = round_up(n_cols < 16 ? 16 : n_cols, 8);
    const int N = 4 * nc;
    const int n_slabs = n_slabs_for(seg_samples);
    const size_t piece = (size_t)N * kSlabBytes;
    std::vector<unsigned char> img((size_t)total, 0);
    SegHeader* h = reinterpret_cast<SegHeader*>(img.data());
    h->magic = kMagic;
    h->n_frame = n_frame;
    h->seg_samples = seg_samples;
    h->n_shift = n_shift;
    h->n_slabs = n_slabs;
    h->nc = nc;
    h->n_cols = n_cols;
    for (int c = 0; c < kMaxNc; ++c) h->group[c] = (c < n_cols) ? h_col_group[c] : -1;
    double peak = 0.0;
    for (size_t i = 0; i < (size_t)n_frame * n_cols; ++i) {
        MS_REQUIRE(isfinite(h_basis[i]), MS_ERR_INVALID_ARG, "ms_dft_seg_plan_build: non-finite basis value");
        peak = fmax(peak, fabs(h_basis[i]));
    }
    MS_REQUIRE(peak <= 1.0, MS_ERR_INVALID_ARG, "ms_dft_seg_plan_build: basis value %g outside [-1, 1]", peak);
    const double scale = peak > 0.0 ? kBasisPeak * (double)(1 << kFracBits) / peak : 1.0;
    h->inv_scale = 1.0 / scale;
    unsigned char* B = img.data() + kHdrBytes;
    // accumulator column jc = slice*nc + c of piece (slab, shift) is row jc of a K-major SWIZZLE_128B [N x 128 B]
    // image: 8-row groups of 1024 B, 16-byte chunk index XOR (row % 8)
    auto put = [&](int shift, int jc, int kb, int8_t val) {
        const int slab = kb / kSlabBytes, kin = kb % kSlabBytes;
        const int chunk = (kin >> 4) ^ (jc & 7);
        B[((size_t)slab * n_shift + shift) * piece + (size_t)(jc >> 3) * 1024 + (size_t)(jc & 7) * 128 + chunk * 16 +
          (kin & 15)] = (unsigned char)val;
    };
    std::vector<int64_t> dsum((size_t)3 * kMaxNc, 0);
    for (int n = 0; n < n_frame; ++n) {
        const int j = n / seg_samples, i = n % seg_samples;
        for (int c = 0; c < n_cols; ++c) {
            const long long v = llrint(h_basis[(size_t)n * n_cols + c] * scale);
            const int q3 = (int)(((v + 128) & 255) - 128);
            const long long v1 = (v - q3) / 256;
            const int q2 = (int)(((v1 + 128) & 255) - 128);
            const int q1 = (int)((v1 - q2) / 256);
            // lo byte (K index 2i): slices 1,2,3 ; hi byte (2i+1): slices 0,1,2
            put(j, 1 * nc + c, 2 * i, (int8_t)q1);
            put(j, 2 * nc + c, 2 * i, (int8_t)q2);
            put(j, 3 * nc + c, 2 * i, (int8_t)q3);
            put(j, 0 * nc + c, 2 * i + 1, (int8_t)q1);
            put(j, 1 * nc + c, 2 * i + 1, (int8_t)q2);
            put(j, 2 * nc + c, 2 * i + 1, (int8_t)q3);
            dsum[0 * kMaxNc + c] += q1;
            dsum[1 * kMaxNc + c] += q2;
            dsum[2 * kMaxNc + c] += q3;
        }
    }
    for (int c = 0; c < kMaxNc; ++c)
        h->off64[c] = 128 * (dsum[0 * kMaxNc + c] * (1ll << 24) + dsum[1 * kMaxNc + c] * (1ll << 16) +
                             dsum[2 * kMaxNc + c] * (1ll << 8));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MS_CUDA_OK(cudaMemcpyAsync(d_plan, img.data(), (size_t)total, cudaMemcpyHostToDevice, st));
    MS_CUDA_OK(cudaStreamSynchronize(st));
    return MS_OK;
}

static int seg_launch(const int16_t* x, int64_t n_files, int64_t file_stride_bytes, int64_t n_tensor_rows,
                      int64_t row_stride_bytes, int64_t n_frames, const void* d_plan, int32_t n_frame,
                      int32_t seg_samples, int32_t n_shift, int32_t n_cols, int64_t out_stride, int64_t out_offset,
                      float* out_band_db, float* out_noise_db, double* acc_band, double* acc_noise, int32_t first,
                      int32_t last, void* stream, double* out_raw, int32_t raw_cols, int32_t raw_col0) {
    using namespace ms;
    MS_REQUIRE(x && d_plan && (out_raw || (out_band_db && out_noise_db)), MS_ERR_INVALID_ARG,
               "ms_band_power_i16_seg: null pointer");
    MS_REQUIRE(n_files > 0 && n_files < ((int64_t)1 << 31) && n_frames >= 0 && n_frames < ((int64_t)1 << 31) &&
                   n_tensor_rows >= 0 && n_tensor_rows < ((int64_t)1 << 31),
               MS_ERR_INVALID_ARG, "ms_band_power_i16_seg: bad extents");
    MS_REQUIRE(row_stride_bytes > 0 && row_stride_bytes % 16 == 0 && (n_files == 1 || (file_stride_bytes > 0 &&
                   file_stride_bytes % 16 == 0)) && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
               MS_ERR_UNSUPPORTED, "ms_band_power_i16_seg: x, the row stride and the file stride must be 16-byte multiples");
    MS_REQUIRE(out_stride >= out_offset + n_frames || n_files == 1, MS_ERR_INVALID_ARG,
               "ms_band_power_i16_seg: out_stride too small");
    MS_REQUIRE((acc_band == nullptr) == (acc_noise == nullptr) && (acc_band != nullptr || (first && last)),
               MS_ERR_INVALID_ARG, "ms_band_power_i16_seg: column groups need both energy accumulators");
    MS_REQUIRE(ms_dft_seg_plan_bytes(n_frame, seg_samples, n_shift, n_cols) > 0, MS_ERR_UNSUPPORTED,
               "ms_band_power_i16_seg: unsupported plan shape");
    if (n_frames == 0) return MS_OK;
    MS_REQUIRE(n_shift == 1 || row_stride_bytes == (int64_t)seg_samples * 2, MS_ERR_INVALID_ARG,
               "ms_band_power_i16_seg: segment mode needs row_stride_bytes == 2 * seg_samples");
    const int nc = round_up(n_cols < 16 ? 16 : n_cols, 8);
    Cfg c;
    MS_REQUIRE(choose_cfg(seg_samples, n_shift, nc, &c), MS_ERR_UNSUPPORTED, "ms_band_power_i16_seg: does not fit");

    EncodeTiledFn encode = get_encode_fn();
    MS_REQUIRE(encode != nullptr, MS_ERR_CUDA, "ms_band_power_i16_seg: cuTensorMapEncodeTiled unavailable");
    // inner extent = the real bytes of a row (a hop segment, or a whole frame in direct mode); TMA zero-fills what a
    // 128-byte box covers beyond it and every row beyond n_tensor_rows
    const cuuint64_t gdim[3] = {(cuuint64_t)seg_samples * 2, (cuuint64_t)n_tensor_rows, (cuuint64_t)n_files};
    const cuuint64_t gstride[2] = {(cuuint64_t)row_stride_bytes, (cuuint64_t)(n_files > 1 ? file_stride_bytes : 16)};
    const cuuint32_t estr[3] = {1, 1, 1};
    static const int l2promo = [] {
        const char* e = getenv("MS_TMA_L2PROMO");
        return e ? atoi(e) : 3;
    }();
    const CUtensorMapL2promotion promo = l2promo == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                         : l2promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : l2promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                        : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUtensorMap tmap, tmap_halo;
    const cuuint32_t box[3] = {(cuuint32_t)kSlabBytes, (cuuint32_t)kTileRows, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int16_t*>(x), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MS_REQUIRE(r == CUDA_SUCCESS, MS_ERR_CUDA, "ms_band_power_i16_seg: cuTensorMapEncodeTiled failed (%d)", (int)r);
    const cuuint32_t hbox[3] = {(cuuint32_t)kSlabBytes, (cuuint32_t)(c.halo_rows > 0 ? c.halo_rows : 8), 1};
    r = encode(&tmap_halo, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int16_t*>(x), gdim, gstride, hbox, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MS_REQUIRE(r == CUDA_SUCCESS, MS_ERR_CUDA, "ms_band_power_i16_seg: halo tensor map failed (%d)", (int)r);

    static thread_local int attr_dev = -1;      // once per device (and host thread)
    int cur_dev = 0;
    MS_CUDA_OK(cudaGetDevice(&cur_dev));
    if (attr_dev != cur_dev) {
        MS_CUDA_OK(cudaFuncSetAttribute(dft_seg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        attr_dev = cur_dev;
    }
    const int64_t rows_per_pass = (int64_t)c.T * kTileRows;
    const int64_t n_pass = n_files * ((n_frames + rows_per_pass - 1) / rows_per_pass);
    int64_t grid = num_sms();
    if (grid > n_pass) grid = n_pass;
    if (grid < 1) grid = 1;
    dft_seg_kernel<<<(unsigned)grid, kThreads, c.smem, static_cast<cudaStream_t>(stream)>>>(
        tmap, tmap_halo, static_cast<const unsigned char*>(d_plan), n_frames, n_files, out_stride, out_offset,
        out_band_db, out_noise_db, acc_band, acc_noise, first, last, c.T, c.n_a, c.n_b, c.resident, c.halo_rows,
        c.n_acc, c.tmem_cols, out_raw, raw_cols, raw_col0);
    MS_CUDA_OK(cudaGetLastError());
    return MS_OK;
}

int ms_band_power_i16_seg(const int16_t* x, int64_t n_files, int64_t file_stride_bytes, int64_t n_tensor_rows,
                          int64_t row_stride_bytes, int64_t n_frames, const void* d_plan, int32_t n_frame,
                          int32_t seg_samples, int32_t n_shift, int32_t n_cols, int64_t out_stride, int64_t out_offset,
                          float* out_band_db, float* out_noise_db, double* acc_band, double* acc_noise, int32_t first,
                          int32_t last, void* stream) {
    return seg_launch(x, n_files, file_stride_bytes, n_tensor_rows, row_stride_bytes, n_frames, d_plan, n_frame,
                      seg_samples, n_shift, n_cols, out_stride, out_offset, out_band_db, out_noise_db, acc_band,
                      acc_noise, first, last, stream, nullptr, 0, 0);
}

int ms_dft_seg_projections_i16(const int16_t* x, int64_t n_files, int64_t file_stride_bytes, int64_t n_rows,
                               int64_t row_stride_bytes, const void* d_plan, int32_t seg_samples, int32_t n_cols,
                               int64_t out_row_stride, double* out_raw, int32_t raw_cols, int32_t raw_col0,
                               void* stream) {
    MS_REQUIRE(out_raw && raw_cols > 0 && raw_col0 >= 0 && raw_col0 + n_cols <= raw_cols, MS_ERR_INVALID_ARG,
               "ms_dft_seg_projections_i16: bad output columns");
    return seg_launch(x, n_files, file_stride_bytes, n_rows, row_stride_bytes, n_rows, d_plan, seg_samples, seg_samples, 1,
                      n_cols, out_row_stride, 0, nullptr, nullptr, nullptr, nullptr, 1, 1, stream, out_raw, raw_cols,
                      raw_col0);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Cosine-series windows (scipy 'hann' / 'hamming' / 'blackman' in their periodic form, the spectrogram default) with
// frame length == nfft and hop | frame: w[n] = sum_m a_m cos(2 pi m n / L) turns the windowed bin into
//     X_f[k] = a_0 R_f[k] + sum_{m>=1} (a_m / 2) (R_f[k-m] + R_f[k+m]),   R_f[k'] = sum_j rot_j[k'] P_{f+j}[k'],
// with P_s[k'] = sum_i x[sH+i] e^{-2 pi i k' i / nfft} the UNWINDOWED partial sum of hop segment s and
// rot_j[k'] = e^{-2 pi i k' j H / nfft}.  P is one product per segment (R = 1: a quarter / half of the MMA
// instructions of the shifted form at 75 / 50 % overlap), computed by dft_seg_kernel in projection mode; this kernel
// adds the phase-rotated partial sums of the frame's segments in fp64 and applies the window in the frequency domain.
namespace ms {
namespace {

constexpr int kCombMaxExt = 160;    // extended bins (band bins +- window order) per call

struct CombineParams {
    const double* proj;      // [file][segment row][2 * n_ext]  (re = sum x cos, im' = sum x sin; P = re - i im')
    const double* rot;       // [n_shift][n_ext][2]  cos, sin of 2 pi k' j H / nfft
    int64_t n_files, rows_per_file, n_frames;
    int32_t n_ext, n_shift, order;           // order = M (1 for Hann / Hamming, 2 for Blackman)
    double coef[3];                          // a_0, a_1 / 2, a_2 / 2
    int32_t sig_lo, sig_n, noise_lo, noise_n;   // bands as index ranges into the extended-bin list (centre positions)
    int64_t out_stride;
    float* out_band_db;
    float* out_noise_db;
    float* out_band_e;
    float* out_noise_e;
};

// One CTA = `blockDim.x` consecutive frames of one file.  Their blockDim.x + n_shift - 1 segment rows are one contiguous
// span of `proj`: it is copied into shared memory with coalesced 16-byte loads (row pitch padded to an odd number of
// 16-byte units, so the per-frame walks below are bank-conflict free), then one thread per frame forms its bins.
__global__ void __launch_bounds__(128)
window_combine_kernel(const CombineParams p, int pitch2) {    // pitch2 = shared-memory row pitch in double2 units
    extern __shared__ double2 comb_s[];                       // [n_shift][n_ext] rotations, then the staged rows
    double2* rot_s = comb_s;
    double2* rows_s = comb_s + (size_t)p.n_shift * p.n_ext;
    const int E = p.n_ext, M = p.order, FR = (int)blockDim.x;
    const int64_t chunks_per_file = (p.n_frames + FR - 1) / FR;
    const int64_t f = blockIdx.x / chunks_per_file;
    const int64_t fr0 = (blockIdx.x - f * chunks_per_file) * FR;
    const int n_fr = (int)((p.n_frames - fr0 < FR) ? p.n_frames - fr0 : FR);
    const int n_rows = n_fr + p.n_shift - 1;
    for (int i = threadIdx.x; i < p.n_shift * E; i += FR) rot_s[i] = reinterpret_cast<const double2*>(p.rot)[i];
    const double2* src = reinterpret_cast<const double2*>(p.proj) + (f * p.rows_per_file + fr0) * E;
    for (int i = threadIdx.x; i < n_rows * E; i += FR) {
        const int r = i / E, e = i - r * E;
        rows_s[(size_t)r * pitch2 + e] = src[i];
    }
    __syncthreads();
    if ((int)threadIdx.x >= n_fr) return;
    const double2* base = rows_s + (size_t)threadIdx.x * pitch2;
    // rectangular-window bin of this frame at extended bin e: the rotated partial sums of its n_shift hop segments
    auto rect = [&](int e, double& re, double& im) {
        re = 0.0;
        im = 0.0;
#pragma unroll 4
        for (int j = 0; j < p.n_shift; ++j) {
            const double2 cs = base[(size_t)j * pitch2 + e];             // (sum x cos, sum x sin): P = c - i s
            const double2 r = rot_s[j * E + e];
            re += cs.x * r.x - cs.y * r.y;                               // (c - i s)(cr - i sr)
            im -= cs.x * r.y + cs.y * r.x;
        }
    };
    // a band walks its extended bins once, keeping the last 2M+1 rectangular bins in a register ring
    auto band = [&](int lo, int n) {
        if (n <= 0) return 0.0;
        double r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0, i0 = 0, i1 = 0, i2 = 0, i3 = 0, i4 = 0;
        double acc = 0.0;
        for (int e = lo - M; e < lo + n + M; ++e) {
            r0 = r1; r1 = r2; r2 = r3; r3 = r4;
            i0 = i1; i1 = i2; i2 = i3; i3 = i4;
            rect(e, r4, i4);
            if (e >= lo + M) {                       // the ring holds bins e-4 .. e; the centre bin is e - M
                double xr, xi;
                if (M == 0) {
                    xr = p.coef[0] * r4;
                    xi = p.coef[0] * i4;
                } else if (M == 1) {
                    xr = p.coef[0] * r3 + p.coef[1] * (r2 + r4);
                    xi = p.coef[0] * i3 + p.coef[1] * (i2 + i4);
                } else {
                    xr = p.coef[0] * r2 + p.coef[1] * (r1 + r3) + p.coef[2] * (r0 + r4);
                    xi = p.coef[0] * i2 + p.coef[1] * (i1 + i3) + p.coef[2] * (i0 + i4);
                }
                acc += xr * xr + xi * xi;
            }
        }
        return acc;
    };
    const double eb = band(p.sig_lo, p.sig_n), en = band(p.noise_lo, p.noise_n);
    const int64_t o = f * p.out_stride + fr0 + threadIdx.x;
    p.out_band_db[o] = (float)(10.0 * log10(eb + 1e-12));
    p.out_noise_db[o] = (float)(10.0 * log10(en + 1e-12));
    if (p.out_band_e) p.out_band_e[o] = (float)eb;
    if (p.out_noise_e) p.out_noise_e[o] = (float)en;
}

}  // namespace
}  // namespace ms

extern "C" int ms_window_combine(const double* proj, const double* rot, int64_t n_files, int64_t rows_per_file,
                                 int64_t n_frames, int32_t n_ext, int32_t n_shift, int32_t order, const double* h_coef,
                                 int32_t sig_lo, int32_t sig_n, int32_t noise_lo, int32_t noise_n, int64_t out_stride,
                                 float* out_band_db, float* out_noise_db, float* out_band_energy,
                                 float* out_noise_energy, void* stream) {
    using namespace ms;
    MS_REQUIRE(proj && rot && h_coef && out_band_db && out_noise_db, MS_ERR_INVALID_ARG, "ms_window_combine: null pointer");
    MS_REQUIRE(n_ext > 0 && n_ext <= kCombMaxExt && n_shift > 0 && order >= 0 && order <= 2, MS_ERR_UNSUPPORTED,
               "ms_window_combine: 1..%d extended bins, window order 0..2", kCombMaxExt);
    MS_REQUIRE(sig_n >= 0 && noise_n >= 0 && (sig_n == 0 || (sig_lo - order >= 0 && sig_lo + sig_n + order <= n_ext)) &&
                   (noise_n == 0 || (noise_lo - order >= 0 && noise_lo + noise_n + order <= n_ext)),
               MS_ERR_INVALID_ARG, "ms_window_combine: band ranges must leave `order` extended bins on both sides");
    MS_REQUIRE(n_frames >= 0 && n_frames + n_shift - 1 <= rows_per_file && out_stride >= n_frames, MS_ERR_INVALID_ARG,
               "ms_window_combine: frames need n_shift segment rows each");
    if (n_files <= 0 || n_frames == 0) return MS_OK;
    CombineParams p = {};
    p.proj = proj;
    p.rot = rot;
    p.n_files = n_files;
    p.rows_per_file = rows_per_file;
    p.n_frames = n_frames;
    p.n_ext = n_ext;
    p.n_shift = n_shift;
    p.order = order;
    for (int i = 0; i < 3; ++i) p.coef[i] = i <= order ? h_coef[i] : 0.0;
    p.sig_lo = sig_lo;
    p.sig_n = sig_n;
    p.noise_lo = noise_lo;
    p.noise_n = noise_n;
    p.out_stride = out_stride;
    p.out_band_db = out_band_db;
    p.out_noise_db = out_noise_db;
    p.out_band_e = out_band_energy;
    p.out_noise_e = out_noise_energy;
    const int pitch2 = n_ext | 1;                                   // odd number of 16-byte units per staged row
    int fr = 128;                                                   // frames per CTA: as many as 96 KiB of staging hold
    auto smem_for = [&](int frames) {
        return ((size_t)n_shift * n_ext + (size_t)(frames + n_shift - 1) * pitch2) * sizeof(double2);
    };
    while (fr > 32 && smem_for(fr) > 96 * 1024) fr >>= 1;
    const size_t sm = smem_for(fr);
    MS_REQUIRE(sm <= 200 * 1024, MS_ERR_UNSUPPORTED, "ms_window_combine: %d extended bins x %d segments do not fit", n_ext, n_shift);
    static thread_local int attr_dev = -1;      // once per device (and host thread)
    int cur_dev = 0;
    MS_CUDA_OK(cudaGetDevice(&cur_dev));
    if (attr_dev != cur_dev) {
        MS_CUDA_OK(cudaFuncSetAttribute(window_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_dev = cur_dev;
    }
    const int64_t chunks = (n_frames + fr - 1) / fr;
    window_combine_kernel<<<(unsigned)(n_files * chunks), fr, sm, static_cast<cudaStream_t>(stream)>>>(p, pitch2);
    MS_CUDA_OK(cudaGetLastError());
    return MS_OK;
}
