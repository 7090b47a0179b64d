/* adt_b200.h — C ABI of the B200-native FFT filter / EQ path of pyAudioDspTools.
 *
 * The reference (ArjaanAuinger/pyaudiodsptools) has no FFI: its boundary for
 * this path is the Python device protocol
 *     config.initialize(sampling_rate, chunk_size)         pyAudioDspTools/config.py:31
 *     CreateHighCutFilter(cutoff).apply(chunk)              pyAudioDspTools/EffectFFTFilter.py:18,49
 *     CreateLowCutFilter(cutoff).apply(chunk)               pyAudioDspTools/EffectFFTFilter.py:91,125
 *     CreateEQ3BandFFT(6 scalars).apply(chunk)              pyAudioDspTools/EffectEQ3BandFFT.py:47,156
 *     CreateEQ3Band(6 scalars).apply{low,mid,high}band(x)   pyAudioDspTools/EffectEQ3Band.py:29,90,121,152
 * The entry points below are what a ctypes binding for exactly those methods
 * needs (see INTEGRATION.md for the binding); pyaudiodsptools_b200/ is that
 * binding.  Plain pointers and sizes only — no torch / cupy types.
 *
 * Conventions: every function returns an adt_status (0 = OK, negative =
 * error) and never throws or aborts; adt_last_error(ctx) gives the message of
 * the last failure on that context.  Audio buffers are planar float32
 * [n_rows][pitch] (one row per mono channel).  Host buffers are borrowed for
 * the duration of the call only.  One adt_ctx / adt_fir / adt_biquad is not
 * thread-safe (same as one reference device object); distinct objects are
 * independent.  Every entry point that does compute requires a CUDA device —
 * there is no CPU fallback.
 */
#ifndef ADT_B200_H
#define ADT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ADT_OK = 0,
    ADT_ERR_INVALID = -1,      /* bad argument (state is left untouched) */
    ADT_ERR_CUDA = -2,         /* CUDA runtime failure */
    ADT_ERR_NO_DEVICE = -3,    /* no usable CUDA device */
    ADT_ERR_UNSUPPORTED = -4,  /* geometry outside what the kernels implement */
    ADT_ERR_NCCL = -5,         /* NCCL failure or libnccl not loadable */
    ADT_ERR_NOMEM = -6
} adt_status;

typedef struct adt_ctx adt_ctx;
typedef struct adt_fir adt_fir;
typedef struct adt_biquad adt_biquad;
typedef struct adt_comm adt_comm;
typedef struct adt_event adt_event;

/* ---- library / device -------------------------------------------------- */
const char* adt_version(void);
const char* adt_status_string(int status);
int adt_device_count(int* count); /* ADT_OK with *count = 0 when no driver/GPU */

/* One context = one device + one CUDA stream all its work is ordered on. */
int adt_ctx_create(int device, adt_ctx** out);
int adt_ctx_destroy(adt_ctx* ctx);
const char* adt_last_error(adt_ctx* ctx);
int adt_ctx_sync(adt_ctx* ctx);
/* number of this library's kernels launched on ctx since creation */
int adt_ctx_launch_count(adt_ctx* ctx, uint64_t* count);
int adt_ctx_device_name(adt_ctx* ctx, char* buf, size_t len);
/* "domain:bus:device.function" of the context's GPU (len >= 13): lets the host side find the GPU's NUMA
 * node under /sys/bus/pci/devices/ and keep pinned staging buffers and the calling thread next to it */
int adt_ctx_pci_bus_id(adt_ctx* ctx, char* buf, size_t len);

/* ---- memory (no torch / cupy to lean on) --------------------------------- */
int adt_malloc(adt_ctx* ctx, size_t bytes, void** dptr);
int adt_free(adt_ctx* ctx, void* dptr);
int adt_malloc_host(adt_ctx* ctx, size_t bytes, void** hptr); /* pinned */
int adt_free_host(adt_ctx* ctx, void* hptr);
int adt_memset(adt_ctx* ctx, void* dptr, int value, size_t bytes);
int adt_memcpy_h2d(adt_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);  /* sync on return */
int adt_memcpy_d2h(adt_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);  /* sync on return */
int adt_memcpy_d2d(adt_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);   /* async */
/* one H2D and one D2H copy running concurrently on two copy streams; returns when both are complete.  The
 * platform's transfer ceiling for the *_host entry points (measurement aid for bench.py). */
int adt_copy_roundtrip_host(adt_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes_h2d, void* dst_host,
                            const void* src_dev, size_t bytes_d2h);

/* ---- CUDA-event timing on the context's stream (for bench.py) ------------ */
int adt_event_create(adt_ctx* ctx, adt_event** out);
int adt_event_destroy(adt_event* ev);
int adt_event_record(adt_event* ev);
int adt_event_elapsed_ms(adt_event* start, adt_event* stop, float* ms); /* syncs on stop */

/* ---- the FFT FIR engine ----------------------------------------------------
 * Computes, per row,   y[m] = sum_k h[k] * x[m - D - k]   (x outside [0,n_in) = 0)
 * by overlap-save blocks: block b reads the fft_size-sample window starting at
 * stream index b*hop - back, circularly convolves it with the filter whose
 * fft_size-point spectrum is `mask`, and keeps circular indices
 * [n0, n0 + hop).  The caller (pyaudiodsptools_b200/design.py) derives
 * (fft_size, hop, n0, back, mask) from the reference's taps; this replaces
 * EffectFFTFilter.py:143-151 / EffectEQ3BandFFT.py:175-211.
 */
typedef struct {
    int32_t fft_size;     /* N: 4096, 8192, 16384 (or 32768) */
    int32_t hop;          /* 1 <= hop, n0 + hop <= N */
    int32_t n0;           /* 0 <= n0 */
    int32_t back;         /* >= 0 */
    int32_t mask_is_real; /* 1: only Re(mask) is used (zero-phase filter) */
    int32_t chunk;        /* streaming chunk size C (0: no streaming state) */
    int32_t n_channels;   /* rows of streaming state (0: none) */
    int32_t reserved;
} adt_fir_desc;

/* mask: fft_size complex64 values (re,im interleaved), natural bin order,
 * without the 1/N factor.  Copied; the pointer is not retained. */
int adt_fir_create(adt_ctx* ctx, const adt_fir_desc* desc, const float* mask, adt_fir** out);
int adt_fir_destroy(adt_fir* fir);
/* Filters too long for one transform (the reference accepts any chunk size, e.g. Example4.py's 88200) are
 * PARTITIONED in time: taps = [h_0 | h_1 | ...], y = sum_s (h_s * x) delayed by the segment offset.  Each
 * segment is an ordinary block plan of its own (its delay folded into `back`); segment 0 stores, the others
 * accumulate into the same output in segment order.  chunk / n_channels must agree.  Plain float32 I/O only
 * (no int16 mode, no store epilogue) when n_segments > 1.  adt_fir_create == one segment. */
int adt_fir_create_segmented(adt_ctx* ctx, int32_t n_segments, const adt_fir_desc* descs, const float* const* masks,
                             adt_fir** out);

/* Whole-buffer mode on DEVICE buffers (async on the context stream): rows of
 * n_in valid samples in, n_out samples out; equals ceil(n/C) successive
 * reference .apply() calls on the zero-padded stream. */
int adt_fir_process_dev(adt_fir* fir, const float* x_dev, int64_t in_pitch, int64_t n_in, float* y_dev,
                        int64_t out_pitch, int64_t n_out, int32_t n_rows);
/* Same on HOST buffers: H2D, kernel and D2H are pipelined over row groups
 * (pinned buffers from adt_malloc_host overlap fully); returns when y is ready. */
int adt_fir_process_host(adt_fir* fir, const float* x_host, int64_t in_pitch, int64_t n_in, float* y_host,
                         int64_t out_pitch, int64_t n_out, int32_t n_rows);

/* 16-bit PCM variants: the WAV conversions either side of the path are fused into
 * the kernel's load and store — x = int16/32768 (Utility.py:236-237), y = (int16)(y*32767) with C
 * truncation (Utility.py:306) — which halves HBM and PCIe bytes.  Reported as a separate mode. */
int adt_fir_process_dev_i16(adt_fir* fir, const int16_t* x_dev, int64_t in_pitch, int64_t n_in, int16_t* y_dev,
                            int64_t out_pitch, int64_t n_out, int32_t n_rows);
int adt_fir_process_host_i16(adt_fir* fir, const int16_t* x_host, int64_t in_pitch, int64_t n_in, int16_t* y_host,
                             int64_t out_pitch, int64_t n_out, int32_t n_rows);

/* Streaming step == one reference .apply(): in/out are [n_channels][chunk]
 * contiguous; updates the device-resident history (2 buffers of back+chunk
 * samples per channel).  Output i corresponds to input i-1 (latency = chunk). */
int adt_fir_apply_host(adt_fir* fir, const float* in_host, float* out_host);
int adt_fir_apply_dev(adt_fir* fir, const float* in_dev, float* out_dev); /* async */
int adt_fir_reset(adt_fir* fir);                                          /* history := 0 */

/* ---- in-repo consumers of the path ------------------------------------------
 * Pointwise wave-shapers, float32 arithmetic in the reference's operation order:
 *   kind 1 = CreateSaturator.apply   (EffectSaturator.py:41-48)  params = {s, 1-s, (s+1)/2, gain, mode(1|2)}
 *            with s = 10^(threshold_dB/20), gain = 10^(makeup_dB/20)                    (bit-exact)
 *   kind 2 = CreateSoftClipper.apply (EffectSoftClipper.py:37-44) params = {drive+1, 0, 0, 0, 0}  (powf ulps)
 * adt_fir_set_epilogue fuses one into the FIR kernel's store (kind 0 detaches); the fused form replaces the
 * IEEE divisions / powf by reciprocal-multiply, __fdividef and __powf (result within ~2 ulp / 2e-6). */
int adt_fir_set_epilogue(adt_fir* fir, int kind, const float* params);
int adt_shape_apply_dev(adt_ctx* ctx, int kind, const float* params, const float* x_dev, float* y_dev, int64_t n);
int adt_shape_apply_host(adt_ctx* ctx, int kind, const float* params, const float* x_host, float* y_host, int64_t n);

/* CreateDelay.apply (EffectDelay.py:31-74): y = x + sum_k ramp[k] * x[n - delay*(k+1)] accumulated in a
 * float32 delay line in the reference's order (bit-exact); wet != 0 returns only the delay line.
 * Buffers are [n_channels][n] contiguous, n <= 2*delay_samples per call. */
typedef struct adt_delay adt_delay;
int adt_delay_create(adt_ctx* ctx, int64_t delay_samples, int32_t feedback_loops, const float* ramp, int32_t wet,
                     int32_t n_channels, adt_delay** out);
int adt_delay_destroy(adt_delay* dl);
int adt_delay_apply_dev(adt_delay* dl, const float* x_dev, float* y_dev, int64_t n);
int adt_delay_apply_host(adt_delay* dl, const float* x_host, float* y_host, int64_t n);

/* ---- the streaming biquad of CreateEQ3Band (EffectEQ3Band.py:90-180) -------
 * y[n] = c0*x[n-1] + c1*x[n-2] + c2*x[n-3] - c3*y[n-1] - c4*y[n-2]
 * (the reference's one-sample numerator delay is part of the contract),
 * float64 arithmetic per step, outputs and feedback state rounded to float32
 * when f64 == 0.  coef = {b0/a0, b1/a0, b2/a0, a1/a0, a2/a0} as doubles.
 * One thread per channel; state (3 inputs, 2 outputs) lives on the device. */
int adt_biquad_create(adt_ctx* ctx, const double coef[5], int32_t n_channels, int32_t f64, adt_biquad** out);
int adt_biquad_destroy(adt_biquad* bq);
int adt_biquad_reset(adt_biquad* bq);
/* rows of n samples; element type float (f64 == 0) or double (f64 == 1) */
int adt_biquad_apply_dev(adt_biquad* bq, const void* x_dev, void* y_dev, int64_t pitch, int64_t n);
int adt_biquad_apply_host(adt_biquad* bq, const void* x_host, void* y_host, int64_t pitch, int64_t n);
/* applyhighband(applymidband(applylowband(x))) in ONE launch: the three bands run as a software pipeline over
 * 32-sample time tiles (one warp per band, tiles handed over in shared memory), so the sequential recurrence
 * is walked once instead of three times.  Same individually rounded arithmetic per band -> bit-identical to
 * the three separate calls; each band's own state is read and updated, so both styles can be mixed. */
int adt_biquad_chain_apply_dev(adt_biquad* low, adt_biquad* mid, adt_biquad* high, const void* x_dev, void* y_dev,
                               int64_t pitch, int64_t n);
int adt_biquad_chain_apply_host(adt_biquad* low, adt_biquad* mid, adt_biquad* high, const void* x_host, void* y_host,
                                int64_t pitch, int64_t n);

/* ---- channel sharding across the GPUs of one box (NCCL over NVLink) ---------
 * Channels are independent, so the only exchanges are a scatter of input rows
 * from a root rank and a gather of output rows back.  libnccl
 * is dlopen()ed on first use. */
#define ADT_NCCL_UNIQUE_ID_BYTES 128
int adt_comm_unique_id(unsigned char id[ADT_NCCL_UNIQUE_ID_BYTES]);
int adt_comm_create(adt_ctx* ctx, const unsigned char id[ADT_NCCL_UNIQUE_ID_BYTES], int32_t rank, int32_t world,
                    adt_comm** out);
int adt_comm_destroy(adt_comm* comm);
/* rows [rank*rows_per_rank, (rank+1)*rows_per_rank) of the root's [world*rows_per_rank][pitch] matrix */
int adt_comm_scatter_rows(adt_comm* comm, const float* full_dev_on_root, float* shard_dev, int64_t rows_per_rank,
                          int64_t pitch, int32_t root);
int adt_comm_gather_rows(adt_comm* comm, const float* shard_dev, float* full_dev_on_root, int64_t rows_per_rank,
                         int64_t pitch, int32_t root);
/* uneven shards (what sharding.channel_range produces when channels % (2*world) != 0): rank r owns
 * row_counts[r] rows, stored back to back in the root's matrix; row_counts has `world` entries and must be
 * identical on every rank.  Ranks with zero rows take no part in the exchange. */
int adt_comm_scatterv_rows(adt_comm* comm, const float* full_dev_on_root, float* shard_dev, const int64_t* row_counts,
                           int64_t pitch, int32_t root);
int adt_comm_gatherv_rows(adt_comm* comm, const float* shard_dev, float* full_dev_on_root, const int64_t* row_counts,
                          int64_t pitch, int32_t root);
int adt_comm_broadcast(adt_comm* comm, void* buf_dev, size_t bytes, int32_t root);
int adt_comm_barrier(adt_comm* comm);

#ifdef __cplusplus
}
#endif
#endif /* ADT_B200_H */
