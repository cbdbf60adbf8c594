/*
 * ms_b200.h -- C-ABI of the B200-native meteor-scatter detection hot path.
 *
 * The reference (th-nuernberg/meteor-scatter) is pure Python and has no FFI;
 * its boundary for this path is three Python callables plus one CSV format
 * (SURVEY.md section 8(b)).  This header is what a ctypes/cffi binding on the
 * reference side binds instead of the numpy/scipy calls cited per function.
 *
 * Conventions
 *  - plain pointers and sizes only; every data pointer is a DEVICE pointer
 *    owned by the caller unless its name starts with h_ (host pointer);
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing
 *    synchronises unless stated;
 *  - return value: 0 (MS_OK) or a negative MS_ERR_* code; the message of the
 *    last error on the calling thread is returned by ms_last_error();
 *  - no global mutable state, re-entrant; there is NO CPU fallback: without a
 *    CUDA device every compute entry point returns MS_ERR_CUDA.
 */
#ifndef MS_B200_H
#define MS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MS_OK 0
#define MS_ERR_INVALID_ARG (-1)
#define MS_ERR_UNSUPPORTED (-2)
#define MS_ERR_CUDA (-3)
#define MS_ERR_WORKSPACE (-4)

#define MS_ABI_VERSION 1

int ms_abi_version(void);
const char* ms_last_error(void);

/* ------------------------------------------------------------------------
 * A-stft: framing + window + real FFT + |X|^2 + band sums + dB.
 * Replaces the STFT loop dsp/src/main.py:376-388 (np.hanning, np.fft.rfft,
 * masked sums, 10*log10(.+1e-12)).
 *
 *   x              [n_files][file_stride] samples (int16 PCM or float32)
 *   n_frames       frames (reference: "blocks") per file, frame j starts at
 *                  sample j*hop of its file
 *   win_len        samples of each frame that enter the transform
 *                  (= min(block_size, nfft): rfft(n=nfft) crops, main.py:379)
 *   window         [win_len] float32 window (np.hanning(block_size)[:win_len])
 *   nfft           transform length, power of two in [256, 16384]
 *   k_*_lo/hi      inclusive rfft bin ranges of the signal and noise bands
 *                  (the inclusive masks of main.py:382,386)
 *   out_band_db, out_noise_db      [n_files][out_stride] float32, 10*log10(E+1e-12)
 *   out_band_energy, out_noise_energy   optional (may be NULL) linear energies
 * ---------------------------------------------------------------------- */
int ms_band_power_i16(const int16_t* x, int64_t n_files, int64_t file_stride, int64_t n_frames,
                      int32_t hop, int32_t win_len, const float* window, int32_t nfft,
                      int32_t k_sig_lo, int32_t k_sig_hi, int32_t k_noise_lo, int32_t k_noise_hi,
                      int64_t out_stride, float* out_band_db, float* out_noise_db,
                      float* out_band_energy, float* out_noise_energy, void* stream);

int ms_band_power_f32(const float* x, int64_t n_files, int64_t file_stride, int64_t n_frames,
                      int32_t hop, int32_t win_len, const float* window, int32_t nfft,
                      int32_t k_sig_lo, int32_t k_sig_hi, int32_t k_noise_lo, int32_t k_noise_hi,
                      int64_t out_stride, float* out_band_db, float* out_noise_db,
                      float* out_band_energy, float* out_noise_energy, void* stream);

/* ------------------------------------------------------------------------
 * A-stft on the tensor cores: the same band energies as a restricted DFT
 * evaluated in exact integer arithmetic with tcgen05.mma kind::i8
 * (raw PCM16 bytes x a 3-digit base-256 fixed-point window*twiddle basis,
 * int32 accumulators in TMEM).  Same reference lines as above.
 *
 *   plan           opaque device blob built by ms_dft_i8_plan_build()
 *   rows           n_files * n_frames frames, frame r starts at byte
 *                  r * row_stride_bytes of x (row_stride_bytes % 16 == 0)
 * ---------------------------------------------------------------------- */
int64_t ms_dft_i8_plan_bytes(int32_t k_samples, int32_t n_cols);
/* h_basis: host double [k_samples][n_cols], column c = window[n]*cos/sin term;
 * col_group: host int32 [n_cols], 0 = signal band, 1 = noise band, -1 = unused;
 * d_plan: device buffer of ms_dft_i8_plan_bytes() bytes (filled via `stream`). */
int ms_dft_i8_plan_build(const double* h_basis, const int32_t* h_col_group, int32_t k_samples,
                         int32_t n_cols, void* d_plan, void* stream);
int ms_band_power_i16_tc(const int16_t* x, int64_t n_rows, int64_t row_stride_bytes,
                         const void* d_plan, int32_t k_samples, int32_t n_cols,
                         float* out_band_db, float* out_noise_db,
                         float* out_band_energy, float* out_noise_energy, void* stream);

/* Same kernel over a batch of files whose frames are not one flat row space
 * (overlapping frames, hop != frame length, or a ragged tail per file): a rank-3
 * TMA tensor map [file][frame][bytes]; frame j of file f starts at byte
 * f*file_stride_bytes + j*row_stride_bytes (both multiples of 16); results go to
 * out[f*out_stride + j].  One persistent launch for the whole batch. */
int ms_band_power_i16_tc_batched(const int16_t* x, int64_t n_files, int64_t file_stride_bytes, int64_t n_frames,
                                 int64_t row_stride_bytes, const void* d_plan, int32_t k_samples, int32_t n_cols,
                                 int64_t out_stride, float* out_band_db, float* out_noise_db,
                                 float* out_band_energy, float* out_noise_energy, void* stream);

/* ------------------------------------------------------------------------
 * General tensor-core restricted DFT (csrc/ms_dft_seg.cu): frames of any
 * length (the basis is streamed through shared memory when it does not fit),
 * up to 64 cos/sin columns (32 bins) per launch, and OVERLAPPING frames
 * (hop < frame) read from HBM once: the audio is addressed as non-overlapping
 * hop segments and frame f accumulates segments f .. f+n_shift-1 against the
 * matching slices of the basis inside TMEM.  Same reference lines as A-stft;
 * the overlapped call form is dsp/src/main.py:52-54, 132-133 (spectrogram with
 * nperseg = nfft, noverlap > 0) and BASELINE configs[3].
 *
 *   segment mode: seg_samples = hop, n_shift = ceil(n_frame / hop),
 *                 row_stride_bytes = 2*hop, n_tensor_rows = samples_per_file / hop
 *                 (whole segments only; frames must not need a later row)
 *   direct mode:  seg_samples = n_frame, n_shift = 1, rows = frames at any
 *                 16-byte-multiple stride, n_tensor_rows = n_frames
 *   h_basis       host double [n_frame][n_cols] (window * cos / sin terms)
 *   acc_band/acc_noise  fp64 [n_files*out_stride] energy accumulators, needed
 *                 when a band is split over several launches (column groups):
 *                 first != 0 starts them, last != 0 writes the dB outputs.
 *   results       out[f*out_stride + out_offset + frame]
 * ---------------------------------------------------------------------- */
int64_t ms_dft_seg_plan_bytes(int32_t n_frame, int32_t seg_samples, int32_t n_shift, int32_t n_cols);
int ms_dft_seg_plan_build(const double* h_basis, const int32_t* h_col_group, int32_t n_frame,
                          int32_t seg_samples, int32_t n_shift, int32_t n_cols, void* d_plan, void* stream);
int ms_band_power_i16_seg(const int16_t* x, int64_t n_files, int64_t file_stride_bytes, int64_t n_tensor_rows,
                          int64_t row_stride_bytes, int64_t n_frames, const void* d_plan, int32_t n_frame,
                          int32_t seg_samples, int32_t n_shift, int32_t n_cols, int64_t out_stride,
                          int64_t out_offset, float* out_band_db, float* out_noise_db, double* acc_band,
                          double* acc_noise, int32_t first, int32_t last, void* stream);

/* Overlapping frames under a cosine-series window (scipy 'hann' / 'hamming' / 'blackman' in their periodic form -- the
 * window of the reference's spectrogram calls, dsp/src/main.py:52-54, 132-133) with frame length == nfft and
 * hop | frame: the window acts in the frequency domain, X[k] = a0 R[k] + sum_m (a_m/2)(R[k-m] + R[k+m]), and the
 * rectangular-window bin R_f[k'] of frame f is the phase-rotated sum of the UNWINDOWED partial sums
 * P_s[k'] = sum_i x[sH+i] e^{-2 pi i k' i / nfft} of its hop segments.  ms_dft_seg_projections_i16 computes P for
 * every segment once on the tensor cores (exact integer arithmetic, columns = cos / sin pairs of the extended bins,
 * fp64 out [file][row][raw_cols], this launch writing columns raw_col0 .. raw_col0 + n_cols - 1);
 * ms_window_combine adds the n_shift rotated partial sums of each frame (rot = host-built device table
 * [n_shift][n_ext][2] of cos / sin(2 pi k' j H / nfft)), applies the window coefficients h_coef = {a0, a1/2, a2/2},
 * sums |X|^2 over the two bands (index ranges into the extended-bin list) and writes dB (and energies). */
int ms_dft_seg_projections_i16(const int16_t* x, int64_t n_files, int64_t file_stride_bytes, int64_t n_rows,
                               int64_t row_stride_bytes, const void* d_plan, int32_t seg_samples, int32_t n_cols,
                               int64_t out_row_stride, double* out_raw, int32_t raw_cols, int32_t raw_col0,
                               void* stream);
int ms_window_combine(const double* proj, const double* rot, int64_t n_files, int64_t rows_per_file, int64_t n_frames,
                      int32_t n_ext, int32_t n_shift, int32_t order, const double* h_coef, int32_t sig_lo,
                      int32_t sig_n, int32_t noise_lo, int32_t noise_n, int64_t out_stride, float* out_band_db,
                      float* out_noise_db, float* out_band_energy, float* out_noise_energy, void* stream);

/* ------------------------------------------------------------------------
 * A-delta + A-thr-global / A-thr-adapt + event extraction.
 * Replaces dsp/src/main.py:393 and get_detections (396-448) /
 * get_detections_adaptive (450-522).
 *
 *   band_db, noise_db   [n_files][stride] float32 (delta = band - noise, fp64)
 *   n_blocks_per_file   optional int32 [n_files]; NULL = n_blocks for all
 *   k_std               threshold_std_factor
 *   window/after/before/fixed   block counts, already truncated with the
 *                       reference's int(sec/block_duration_sec) (main.py:458-461)
 *   max_events          capacity per file of the event arrays
 *   out_events          [n_files][max_events][2] int32 (start, stop_exclusive)
 *   out_event_db        [n_files][max_events] float64 mean(delta[start:stop])
 *   out_counts          [n_files] int32 events found (may exceed max_events:
 *                       the caller must treat that as an overflow)
 *   out_thresholds      optional [n_files][stride] float64 per-block threshold
 *                       (global detector: element 0 of each file only)
 *   out_near            optional [n_files][stride] uint8, 1 where
 *                       |delta - threshold| < eps_db (reported separately)
 *   workspace           device scratch of ms_detect_workspace_bytes() bytes
 * ---------------------------------------------------------------------- */
int64_t ms_detect_workspace_bytes(int64_t n_files, int64_t stride);

int ms_detect_global(const float* band_db, const float* noise_db, int64_t n_files, int64_t stride,
                     int64_t n_blocks, const int32_t* n_blocks_per_file, double k_std,
                     int32_t max_events, int32_t* out_events, double* out_event_db, int32_t* out_counts,
                     double* out_thresholds, uint8_t* out_near, double eps_db,
                     void* workspace, int64_t workspace_bytes, void* stream);

int ms_detect_adaptive(const float* band_db, const float* noise_db, int64_t n_files, int64_t stride,
                       int64_t n_blocks, const int32_t* n_blocks_per_file, double k_std,
                       int32_t window_blocks, int32_t freeze_before_blocks, int32_t freeze_after_blocks,
                       int32_t fixed_blocks,
                       int32_t max_events, int32_t* out_events, double* out_event_db, int32_t* out_counts,
                       double* out_thresholds, uint8_t* out_near, double eps_db,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* ms_detect_adaptive with the A-hour stage (see ms_hourly_counts below) fused
 * into the same launch: every event found is also counted into out_hist.
 * flags: MS_DETECT_SMALL_FOOTPRINT launches 128-thread CTAs without dynamic shared
 * memory (per-block arrays in the workspace) so the kernel can run on a second
 * stream underneath the persistent band-power kernel of the next batch. */
#define MS_DETECT_SMALL_FOOTPRINT 1u
int ms_detect_adaptive_hourly(const float* band_db, const float* noise_db, int64_t n_files, int64_t stride,
                              int64_t n_blocks, const int32_t* n_blocks_per_file, double k_std,
                              int32_t window_blocks, int32_t freeze_before_blocks, int32_t freeze_after_blocks,
                              int32_t fixed_blocks,
                              int32_t max_events, int32_t* out_events, double* out_event_db, int32_t* out_counts,
                              double* out_thresholds, uint8_t* out_near, double eps_db,
                              void* workspace, int64_t workspace_bytes,
                              const int64_t* file_start_us, double block_duration_sec, double crit_min_dur_sec,
                              int64_t hour0, int32_t n_hours, int32_t* out_hist, uint32_t flags, void* stream);

/* ------------------------------------------------------------------------
 * A-hour + Kritisch rule: events -> hourly [Anzahl, Kritisch] histogram.
 * Replaces the hour bucketing of dsp/src/main.py:690-696 /
 * dsp/src/main_analyze.py:70-73 and the critical rule "duration >= 0.5 s" of
 * meteor_detect_class/detector_and_classification.py:50.
 *
 *   file_start_us   [n_files] int64 microseconds since the Unix epoch (naive UTC)
 *   hour0           hours since the epoch of histogram row 0
 *   out_hist        [n_hours][2] int32, ACCUMULATED into (zero it first)
 * utc_start = file_start + timedelta(seconds=start*block_duration_sec) is
 * evaluated with Python's timedelta rounding (modf + round-half-even to us).
 * ---------------------------------------------------------------------- */
int ms_hourly_counts(const int32_t* events, const int32_t* counts, int64_t n_files, int32_t max_events,
                     const int64_t* file_start_us, double block_duration_sec, double crit_min_dur_sec,
                     int64_t hour0, int32_t n_hours, int32_t* out_hist, void* stream);

/* ------------------------------------------------------------------------
 * B-psd + B-band: per-block Welch PSD band sums in dB.
 * Replaces scipy.signal.welch(block, fs, nfft) at
 * dsp/src/live/backend/processor.py:206 and the band sums :349-367, :393.
 *
 *   x          [n_streams][stream_stride] float32 samples in [-1, 1)
 *   block      samples per block (int(proc_block_sec*fs)); segments of
 *              `nperseg` with hop nperseg/2, periodic Hann, mean removed
 *   window     [nperseg] float32 periodic Hann (scipy get_window('hann', nperseg))
 *   h_bands    host int32 [3][2] inclusive bin ranges (signal, noise1, noise2)
 *   scale      1 / (fs * sum(w^2)); bins other than DC/Nyquist are doubled
 *   out_db     [n_streams][n_blocks][4] float32: ms_dB, n1_dB, n2_dB, db2
 *   out_rows   optional (NULL = off): the block's PSD in dB (processor.py:207
 *              block_psd_db) for bins row_lo..row_hi, [n_streams][n_blocks][row_hi-row_lo+1]
 *              -- the rows of the reference's waterfall (processor.py:223-229)
 * ---------------------------------------------------------------------- */
int ms_welch_band_db_f32(const float* x, int64_t n_streams, int64_t stream_stride, int64_t n_blocks,
                         int32_t block, int32_t nperseg, const float* window, int32_t nfft,
                         const int32_t* h_bands, double scale, float* out_db,
                         int32_t row_lo, int32_t row_hi, float* out_rows, void* stream);
/* PCM16 input, scaled by 1/32768 first (what soundfile.read does, processor.py:65-71) */
int ms_welch_band_db_i16(const int16_t* x, int64_t n_streams, int64_t stream_stride, int64_t n_blocks,
                         int32_t block, int32_t nperseg, const float* window, int32_t nfft,
                         const int32_t* h_bands, double scale, float* out_db,
                         int32_t row_lo, int32_t row_hi, float* out_rows, void* stream);

/* The same band levels WITHOUT FFTs: the Welch band power of a segment is the quadratic form
 * x^T (P Q P) x (P = mean removal, Q = windowed band kernel), which has numerical rank ~26 for the
 * reference's 102-of-4096-bin bands; d_basis holds sqrt(lambda_r/lambda_max) * u_r for up to 32
 * eigenvectors per band, laid out [nperseg][32] float4 = (band0, band1, band2, 0) for column l at sample n;
 * h_group_scale[g] = lambda_max_g * scale / n_segments (host doubles).  Built by ops.WelchQuadform. */
int ms_welch_band_db_qf_i16(const int16_t* x, int64_t n_streams, int64_t stream_stride, int64_t n_blocks,
                            int32_t block, int32_t nperseg, const float* d_basis, const double* h_group_scale,
                            float* out_db, void* stream);
int ms_welch_band_db_qf_f32(const float* x, int64_t n_streams, int64_t stream_stride, int64_t n_blocks,
                            int32_t block, int32_t nperseg, const float* d_basis, const double* h_group_scale,
                            float* out_db, void* stream);

/* The same quadratic form on the tensor cores (tcgen05 kind::i8, exact integer accumulation; csrc/ms_welch_i8.cu):
 * PCM16 only, nperseg a multiple of 64, at most 26 eigenvector columns per band, hop/block/stream stride multiples of
 * 16 bytes.  h_basis is host double [3][cols_per_band][nperseg] (band-major; sqrt(lambda_r/lambda_max) * u_r);
 * the plan (ms_welch_i8_plan_bytes bytes of device memory) holds the two-digit s8 image of the normalised columns.
 * h_group_scale as for the _qf_ entry points.  Anything outside these limits returns MS_ERR_UNSUPPORTED: use
 * ms_welch_band_db_qf_i16. */
int64_t ms_welch_i8_plan_bytes(int32_t nperseg);
int ms_welch_i8_plan_build(const double* h_basis, int32_t nperseg, int32_t cols_per_band, void* d_plan, void* stream);
int ms_welch_band_db_i8_i16(const int16_t* x, int64_t n_streams, int64_t stream_stride, int64_t n_blocks,
                            int32_t block, int32_t nperseg, const void* d_plan, const double* h_group_scale,
                            float* out_db, void* stream);

/* ------------------------------------------------------------------------
 * B-state: threshold history + Init/Detection/Tracking machine, resumable.
 * Replaces dsp/src/live/backend/processor.py:393-414, 444-510 and the state
 * dataclasses dsp/src/live/backend/aggregates.py:9-24.
 * ---------------------------------------------------------------------- */
#define MS_LIVE_HIST_MAX 256

typedef struct ms_live_config {
    int64_t block_samples;         /* int(proc_block_sec * fs) */
    double fs;                     /* sample rate */
    double k_std;                  /* threshold_std_factor */
    double init_wait_sec;          /* init_detection_wait_sec */
    double after_wait_sec;         /* after_tracking_wait_sec */
    double mean_min_db;            /* detection_db_over_noise_mean_min */
    double dur_min_sec;            /* detection_dur_min_sec */
    int32_t avg_win;               /* int(avg_win_sec/proc_block_sec), 1..MS_LIVE_HIST_MAX */
    int32_t reserved;
} ms_live_config;

typedef struct ms_live_state {
    int64_t block_index;           /* blocks consumed so far */
    int32_t state;                 /* 0 Init, 1 Detection, 2 Tracking */
    int32_t hist_len;              /* valid entries of hist (<= avg_win) */
    int32_t hist_pos;              /* ring write position */
    int32_t trk_n;                 /* Tracking: samples accumulated */
    double locked_threshold;
    double lock_until_sec;
    double trk_t0;                 /* Tracking: time_start_detection */
    double trk_sum, trk_min, trk_max;   /* Tracking statistics */
    double trk_mean_run, trk_m2_run;    /* std accumulators: first tracked value (shift), sum of squared deviations from it */
    double hist[MS_LIVE_HIST_MAX]; /* last avg_win db2 values */
} ms_live_state;

/* db2: [n_streams][n] float32 (stride db2_stride, element stride db2_elem);
 * states: [n_streams]; out_det: [n_streams][max_det][7] float64
 * (time_start, time_stop, duration, db_min, db_max, db_mean, db_std);
 * out_det_count: [n_streams] int32 (ACCUMULATED); out_thresholds optional [n_streams][n]. */
int ms_live_state_step(ms_live_state* states, const ms_live_config* h_cfg, int64_t n_streams,
                       const float* db2, int64_t db2_stride, int32_t db2_elem, int64_t n,
                       int32_t max_det, double* out_det, int32_t* out_det_count,
                       double* out_thresholds, void* stream);
/* The same with caller-owned scratch.  Calls of n >= 32 blocks (the batch form: parallel threshold pre-pass + one
 * warp per stream jumping from event to event) need ms_live_state_workspace_bytes(n_streams, n) bytes of 32-byte
 * aligned device memory; shorter (streaming) calls need none.  ms_live_state_step allocates that scratch itself with
 * cudaMallocAsync and, to keep it cached, raises the release threshold of the device's default memory pool. */
int64_t ms_live_state_workspace_bytes(int64_t n_streams, int64_t n);
int ms_live_state_step_ws(ms_live_state* states, const ms_live_config* h_cfg, int64_t n_streams,
                          const float* db2, int64_t db2_stride, int32_t db2_elem, int64_t n,
                          int32_t max_det, double* out_det, int32_t* out_det_count,
                          double* out_thresholds, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * C-stft / sweep: one-sided PSD spectrogram rows + noise-band density.
 * Replaces plt.specgram(x, Fs, NFFT=2048, noverlap=1024) and the noise-floor
 * scalar of meteor_detect_class/prime_detection.py:67-92, and
 * scipy.signal.spectrogram(..., scaling='density', mode='psd') at
 * dsp/src/main.py:52-54.
 *
 *   x           [n_segments][seg_stride] samples
 *   window      [nfft] float32; scale = 1/(fs*sum(w^2))
 *   k_lo..k_hi  inclusive bin range written to out_psd
 *   out_psd     [n_segments][k_hi-k_lo+1][n_frames] float32 (F x T like mlab)
 *   out_noise_sum  [n_segments] float64 sum over time AND bins of PSD in
 *               k_noise_lo..k_noise_hi (prime_detection.py:83), ACCUMULATED
 * nfft == 2048 on pair-aligned data with all bins below nfft/2 runs on the
 * warp-per-frame kernel (csrc/ms_fft_warp.cuh); every other call on the
 * block-cooperative FFT kernel.  MS_PSD_IMPL=fft forces the latter.
 * ---------------------------------------------------------------------- */
int ms_psd_spectrogram_i16(const int16_t* x, int64_t n_segments, int64_t seg_stride, int64_t n_frames,
                           int32_t hop, int32_t nfft, const float* window, double scale,
                           int32_t k_lo, int32_t k_hi, int32_t k_noise_lo, int32_t k_noise_hi,
                           float* out_psd, double* out_noise_sum, void* stream);
int ms_psd_spectrogram_f32(const float* x, int64_t n_segments, int64_t seg_stride, int64_t n_frames,
                           int32_t hop, int32_t nfft, const float* window, double scale,
                           int32_t k_lo, int32_t k_hi, int32_t k_noise_lo, int32_t k_noise_hi,
                           float* out_psd, double* out_noise_sum, void* stream);

/* ------------------------------------------------------------------------
 * One call = one pass of detector A over a batch resident in HBM
 * (dsp/src/main.py:352-527 + 690-696): zero out_hist, tensor-core band power,
 * adaptive detection, hourly counts.  x is [n_files][n_blocks*block_size] PCM16
 * (files back to back).  ev_stft_begin / ev_stft_end are optional cudaEvent_t
 * recorded around the STFT kernel (profiling hook).  Other arguments as in the
 * functions it composes.
 * ---------------------------------------------------------------------- */
int ms_detector_a_pass_i16(const int16_t* x, int64_t n_files, int64_t n_blocks, int32_t block_size,
                           const void* d_plan, int32_t k_samples, int32_t n_cols, double k_std,
                           int32_t window_blocks, int32_t freeze_before_blocks, int32_t freeze_after_blocks,
                           int32_t fixed_blocks, int32_t max_events,
                           float* band_db, float* noise_db, int32_t* out_events, double* out_event_db,
                           int32_t* out_counts, void* workspace, int64_t workspace_bytes,
                           const int64_t* file_start_us, double block_duration_sec, double crit_min_dur_sec,
                           int64_t hour0, int32_t n_hours, int32_t* out_hist,
                           void* ev_stft_begin, void* ev_stft_end, void* stream);

/* ------------------------------------------------------------------------
 * The same pass for a sequence of batches (a 30-day archive is 30 of them per
 * GPU): the band-power kernel runs on `stream`, the detect + hourly stage of
 * the same batch on `side_stream` (MS_DETECT_SMALL_FOOTPRINT), i.e. under the
 * band-power kernel of the next call.  Each batch in flight owns a slot: its
 * output buffers, workspace, histogram and the two cudaEvent_t handles.
 *   ev_stft_done    recorded on `stream` after the band-power kernel
 *   ev_detect_done  recorded on `side_stream` after the detect kernel; the call
 *                   first makes `stream` wait for it, which orders the reuse of
 *                   a slot (a never-recorded event does not block)
 * Results of a slot are valid once ev_detect_done has completed.
 * ---------------------------------------------------------------------- */
int ms_detector_a_pass_overlapped_i16(const int16_t* x, int64_t n_files, int64_t n_blocks, int32_t block_size,
                           const void* d_plan, int32_t k_samples, int32_t n_cols, double k_std,
                           int32_t window_blocks, int32_t freeze_before_blocks, int32_t freeze_after_blocks,
                           int32_t fixed_blocks, int32_t max_events,
                           float* band_db, float* noise_db, int32_t* out_events, double* out_event_db,
                           int32_t* out_counts, void* workspace, int64_t workspace_bytes,
                           const int64_t* file_start_us, double block_duration_sec, double crit_min_dur_sec,
                           int64_t hour0, int32_t n_hours, int32_t* out_hist,
                           void* ev_stft_begin, void* ev_stft_end, void* stream,
                           void* side_stream, void* ev_stft_done, void* ev_detect_done);

/* A-io, file side (dsp/src/main.py:249, scipy.io.wavfile.read once per file): read n_bytes[i] bytes at
 * offsets[i] of paths[i] into dst[i] for n_files files with n_threads native threads (pread; no interpreter lock,
 * no intermediate buffer); bytes between n_bytes[i] and dst_capacity[i] are zeroed (ragged batches).  Host memory
 * only (dst should be page-locked rows); blocks until every file is read. */
int ms_read_files(const char* const* paths, const int64_t* offsets, const int64_t* n_bytes, void* const* dst,
                  const int64_t* dst_capacity, int32_t n_files, int32_t n_threads);

/* ------------------------------------------------------------------------
 * A-io fast path: strided host->device copy of only the samples the transform
 * reads.  Replaces "load the whole WAV" (dsp/src/main.py:249) for batch ingest:
 * rfft(n=n_fft) crops each windowed block to its first min(n_fft, block) samples
 * (main.py:379), so row r copies row_bytes from h_src + r*src_row_stride_bytes
 * to d_dst + r*dst_row_stride_bytes.  h_src is a HOST pointer (pinned for
 * asynchronous DMA); one cudaMemcpy2DAsync on `stream`.
 * ---------------------------------------------------------------------- */
int ms_ingest_rows_h2d(const void* h_src, int64_t n_rows, int64_t src_row_stride_bytes, int64_t row_bytes,
                       void* d_dst, int64_t dst_row_stride_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MS_B200_H */
