// One-call pass of detector A over a batch resident in HBM:
//   zero histogram -> STFT band power (tensor-core K2) -> delta/adaptive threshold/events + hourly counts.
// A convenience composition of the three entry points above it in ms_b200.h, so a
// host binding pays one FFI call per batch (dsp/src/main.py:352-527, 690-696).
#include <stdlib.h>

#include "ms_common.cuh"

namespace ms {
int band_power_i16_tc_impl(const int16_t* x, int64_t n_rows, int64_t row_stride_bytes, const void* d_plan,
                           int32_t k_samples, int32_t n_cols, float* out_band_db, float* out_noise_db,
                           float* out_band_energy, float* out_noise_energy, int32_t* zero_buf, int32_t zero_count,
                           void* stream, int64_t n_files = 0, int64_t file_stride_bytes = 0, int64_t out_stride = 0,
                           int fix_warps_req = 0, int grid_req = 0);
int detect_adaptive_hourly_pdl(const float* band_db, const float* noise_db, int64_t n_files, int64_t n_blocks,
                               double k_std, int32_t window, int32_t before, int32_t after, int32_t fixed,
                               int32_t max_events, int32_t* out_events, double* out_event_db, int32_t* out_counts,
                               void* workspace, int64_t workspace_bytes, const int64_t* file_start_us,
                               double block_duration_sec, double crit_min_dur_sec, int64_t hour0, int32_t n_hours,
                               int32_t* out_hist, void* stream);
}  // namespace ms

extern "C" int ms_detector_a_pass_i16(const int16_t* x, int64_t n_files, int64_t n_blocks, int32_t block_size,
                                      const void* d_plan, int32_t k_samples, int32_t n_cols, double k_std,
                                      int32_t window_blocks, int32_t freeze_before_blocks,
                                      int32_t freeze_after_blocks, int32_t fixed_blocks, int32_t max_events,
                                      float* band_db, float* noise_db, int32_t* out_events, double* out_event_db,
                                      int32_t* out_counts, void* workspace, int64_t workspace_bytes,
                                      const int64_t* file_start_us, double block_duration_sec,
                                      double crit_min_dur_sec, int64_t hour0, int32_t n_hours, int32_t* out_hist,
                                      void* ev_stft_begin, void* ev_stft_end, void* stream) {
    MS_REQUIRE(x && d_plan && band_db && noise_db && out_hist, MS_ERR_INVALID_ARG, "ms_detector_a_pass_i16: null pointer");
    MS_REQUIRE(n_files >= 0 && n_blocks >= 0 && block_size > 0 && n_hours > 0, MS_ERR_INVALID_ARG,
               "ms_detector_a_pass_i16: bad sizes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ev_stft_begin) MS_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(ev_stft_begin), st));
    // files are back to back ([n_files][n_blocks*block_size] samples): one flat row space; the
    // band-power kernel also clears the histogram (no separate memset in the stream)
    int rc = ms::band_power_i16_tc_impl(x, n_files * n_blocks, (int64_t)block_size * 2, d_plan, k_samples, n_cols,
                                        band_db, noise_db, nullptr, nullptr, out_hist, 2 * n_hours, stream);
    if (rc != MS_OK) return rc;
    if (ev_stft_end) MS_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(ev_stft_end), st));
    // the profiling hook serialises the stream anyway; without it the detect kernel is a programmatic
    // dependent of the band-power kernel so its launch overlaps that kernel's tail
    if (ev_stft_end)
        return ms_detect_adaptive_hourly(band_db, noise_db, n_files, n_blocks, n_blocks, nullptr, k_std,
                                         window_blocks, freeze_before_blocks, freeze_after_blocks, fixed_blocks,
                                         max_events, out_events, out_event_db, out_counts, nullptr, nullptr, 0.0,
                                         workspace, workspace_bytes, file_start_us, block_duration_sec,
                                         crit_min_dur_sec, hour0, n_hours, out_hist, 0u, stream);
    return ms::detect_adaptive_hourly_pdl(band_db, noise_db, n_files, n_blocks, k_std, window_blocks,
                                          freeze_before_blocks, freeze_after_blocks, fixed_blocks, max_events,
                                          out_events, out_event_db, out_counts, workspace, workspace_bytes,
                                          file_start_us, block_duration_sec, crit_min_dur_sec, hour0, n_hours,
                                          out_hist, stream);
}

// The same pass for back-to-back batches: the detect stage of THIS batch runs on `side_stream`, under the band-power
// kernel of the NEXT batch (small-footprint detect CTAs fit beside the persistent band-power CTAs).  The caller owns a
// slot (band/noise/event/histogram buffers + the two events) per batch in flight; reusing a slot is ordered by
// ev_detect_done, which the next call on that slot makes `stream` wait for.  Stream and event plumbing only.
extern "C" int ms_detector_a_pass_overlapped_i16(
    const int16_t* x, int64_t n_files, int64_t n_blocks, int32_t block_size, const void* d_plan, int32_t k_samples,
    int32_t n_cols, double k_std, int32_t window_blocks, int32_t freeze_before_blocks, int32_t freeze_after_blocks,
    int32_t fixed_blocks, int32_t max_events, float* band_db, float* noise_db, int32_t* out_events,
    double* out_event_db, int32_t* out_counts, void* workspace, int64_t workspace_bytes, const int64_t* file_start_us,
    double block_duration_sec, double crit_min_dur_sec, int64_t hour0, int32_t n_hours, int32_t* out_hist,
    void* ev_stft_begin, void* ev_stft_end, void* stream, void* side_stream, void* ev_stft_done, void* ev_detect_done) {
    MS_REQUIRE(x && d_plan && band_db && noise_db && out_hist, MS_ERR_INVALID_ARG,
               "ms_detector_a_pass_overlapped_i16: null pointer");
    MS_REQUIRE(side_stream && ev_stft_done && ev_detect_done && side_stream != stream, MS_ERR_INVALID_ARG,
               "ms_detector_a_pass_overlapped_i16: needs a second stream and two events");
    MS_REQUIRE(n_files >= 0 && n_blocks >= 0 && block_size > 0 && n_hours > 0, MS_ERR_INVALID_ARG,
               "ms_detector_a_pass_overlapped_i16: bad sizes");
    cudaStream_t st = static_cast<cudaStream_t>(stream), side = static_cast<cudaStream_t>(side_stream);
    cudaEvent_t k2_done = static_cast<cudaEvent_t>(ev_stft_done), k3_done = static_cast<cudaEvent_t>(ev_detect_done);
    MS_CUDA_OK(cudaStreamWaitEvent(st, k3_done, 0));   // the slot's previous batch has been consumed (no-op if never recorded)
    if (ev_stft_begin) MS_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(ev_stft_begin), st));
    // Two ways to run the detect kernel of the previous batch under this band-power kernel:
    //  spare = 0 (default): all SMs run band-power CTAs (4 fix-up warps, 144 registers) and small-footprint detect
    //             CTAs share them;
    //  spare > 0: the persistent band-power grid leaves `spare` SMs free and the ordinary detect kernel (288 CTAs,
    //             ~12 us chain each) runs there; the band-power kernel keeps its 8 fix-up warps.
    // Same-process A/B on one B200 (tools/ab_pipeline.py, 12 alternating rounds of 50 steps): against the in-stream
    // pass the co-resident form gains 3.1 %, leaving 8 / 12 / 16 SMs free LOSES 1.9 / 3.3 / 6.2 % (the band-power
    // kernel slows in proportion to the SMs it gives up).
    static const int spare = [] {      // tuning knob: MS_OVL_SPARE_SMS
        const char* e = getenv("MS_OVL_SPARE_SMS");
        return e ? atoi(e) : 0;
    }();
    const int sms = ms::num_sms();
    const bool split = spare > 0 && spare < sms;
    int rc = ms::band_power_i16_tc_impl(x, n_files * n_blocks, (int64_t)block_size * 2, d_plan, k_samples, n_cols,
                                        band_db, noise_db, nullptr, nullptr, out_hist, 2 * n_hours, stream, 0, 0, 0,
                                        /*fix_warps_req=*/split ? 8 : 4, /*grid_req=*/split ? sms - spare : 0);
    if (rc != MS_OK) return rc;
    if (ev_stft_end) MS_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(ev_stft_end), st));
    MS_CUDA_OK(cudaEventRecord(k2_done, st));
    MS_CUDA_OK(cudaStreamWaitEvent(side, k2_done, 0));
    rc = ms_detect_adaptive_hourly(band_db, noise_db, n_files, n_blocks, n_blocks, nullptr, k_std, window_blocks,
                                   freeze_before_blocks, freeze_after_blocks, fixed_blocks, max_events, out_events,
                                   out_event_db, out_counts, nullptr, nullptr, 0.0, workspace, workspace_bytes,
                                   file_start_us, block_duration_sec, crit_min_dur_sec, hour0, n_hours, out_hist,
                                   split ? 0u : MS_DETECT_SMALL_FOOTPRINT, side_stream);
    if (rc != MS_OK) return rc;
    MS_CUDA_OK(cudaEventRecord(k3_done, side));
    return MS_OK;
}

// A-io on the fast path: host PCM -> device, copying only the samples the transform reads.
// The reference loads the whole file (dsp/src/main.py:249) but np.fft.rfft(n=n_fft) crops every
// windowed block to its first min(n_fft, block) samples (main.py:379), so the tail of each block
// never has to cross PCIe.  One strided DMA (cudaMemcpy2DAsync) per call; h_src should be pinned.
extern "C" int ms_ingest_rows_h2d(const void* h_src, int64_t n_rows, int64_t src_row_stride_bytes,
                                  int64_t row_bytes, void* d_dst, int64_t dst_row_stride_bytes, void* stream) {
    MS_REQUIRE(h_src && d_dst, MS_ERR_INVALID_ARG, "ms_ingest_rows_h2d: null pointer");
    MS_REQUIRE(n_rows >= 0 && row_bytes > 0 && src_row_stride_bytes >= row_bytes && dst_row_stride_bytes >= row_bytes,
               MS_ERR_INVALID_ARG, "ms_ingest_rows_h2d: bad geometry");
    if (n_rows == 0) return MS_OK;
    MS_CUDA_OK(cudaMemcpy2DAsync(d_dst, (size_t)dst_row_stride_bytes, h_src, (size_t)src_row_stride_bytes,
                                 (size_t)row_bytes, (size_t)n_rows, cudaMemcpyHostToDevice,
                                 static_cast<cudaStream_t>(stream)));
    return MS_OK;
}
