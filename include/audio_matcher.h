/* audio_matcher.h -- C ABI of libaudio_matcher_b200.so
 *
 * B200-native (sm_100a) drop-in for the one data-parallel hot path of
 * NilsJochem/audio-matcher: slide a snippet over a long stream with block-wise
 * FFT cross-correlation and pick the peak offsets.  The reference has no FFI of
 * its own; the seams these entry points replace are (paths relative to the
 * reference checkout):
 *
 *   LibConvolve::new                              src/matcher/audio_matcher.rs:289
 *   CorrelateAlgo::inverse_sample_auto_correlation  src/matcher/audio_matcher.rs:66,321-329
 *   CorrelateAlgo::correlate_with_sample          src/matcher/audio_matcher.rs:67-72,331-343
 *   calc_chunks                                   src/matcher/audio_matcher.rs:88-141
 *   Config / PeakConfig                           src/matcher/audio_matcher.rs:24-53
 *   find_peaks::Peak<f32>                         src/matcher/audio_matcher.rs:97,124-127
 *   PCM scale + stereo downmix                    src/matcher/mp3_reader.rs:12,33-36
 *
 * INTEGRATION.md shows the Rust binding (`extern "C"` block + CudaConvolve shim)
 * a maintainer of the reference would add.
 *
 * Conventions: every call returns an am_status; am_last_error() returns a
 * thread-local message for the last failure on the calling thread.  Handles are
 * opaque and internally locked (a handle may be shared between threads, calls on
 * it serialise).  All output buffers are caller-allocated: capacity in, count
 * out.  No C++ types, no exceptions, no torch types cross this boundary.
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with AM_ERR_CUDA.
 */
#ifndef AUDIO_MATCHER_B200_H
#define AUDIO_MATCHER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AM_ABI_VERSION 3

typedef enum {
    AM_OK = 0,
    AM_ERR_INVALID = 1,     /* bad argument */
    AM_ERR_CUDA = 2,        /* CUDA runtime / launch failure, or no device */
    AM_ERR_CAPACITY = 3,    /* caller buffer or an internal per-chunk peak list too small */
    AM_ERR_NOMEM = 4,
    AM_ERR_UNSUPPORTED = 5  /* e.g. snippet longer than the largest FFT block */
} am_status;

/* Mode, src/matcher/audio_matcher.rs:54-59 */
typedef enum { AM_MODE_FULL = 0, AM_MODE_SAME = 1, AM_MODE_VALID = 2 } am_mode;

/* stream / snippet sample formats.  I16 variants apply the reference's scale
 * (l + r) * 0.5 * (1/65535) on load (mono: l = r = s), mp3_reader.rs:12,35 */
typedef enum { AM_FMT_F32_MONO = 0, AM_FMT_I16_MONO = 1, AM_FMT_I16_STEREO = 2 } am_sample_fmt;

typedef enum { AM_MEM_HOST = 0, AM_MEM_DEVICE = 1 } am_mem;

/* Config + PeakConfig (audio_matcher.rs:24-53) as plain data */
typedef struct {
    double chunk_size_s;   /* args.rs:71, default 60 */
    double overlap_s;      /* snippet duration (audio_matcher.rs:41); < 0 => m / sr */
    double distance_s;     /* args.rs:75, default 480; find_peaks uses whole seconds (:228) */
    float prominence;      /* already divided by 100 (audio_matcher.rs:44), default 0.13 */
    uint32_t fft_log2;     /* overlap-save block length 2^fft_log2; 0 = choose from the snippet length */
    uint32_t max_peaks_per_chunk;  /* capacity of the per-chunk list before min-distance suppression; 0 = 1024 */
    uint32_t reserved;
} am_config;

/* find_peaks::Peak<f32> with global positions (after offset_range, lib.rs:8-10) */
typedef struct {
    uint64_t start, end;   /* position: Range<usize>, end exclusive */
    float height;
    float prominence;
    float left_diff, right_diff;
    uint32_t snippet_id;   /* always 0 for a single-snippet matcher */
    uint32_t chunk;        /* logical chunk index that produced the peak */
} am_peak;

/* counters of the work the last am_correlate / am_calc_chunks* call did */
typedef struct {
    uint64_t kernel_launches;   /* CUDA kernels launched by this library */
    uint64_t fft_blocks;        /* overlap-save blocks transformed */
    uint64_t frames;            /* stream frames covered */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t fft_log2, log2_n1, log2_n2;   /* block length and its four-step split (n1 = 1 => single pass) */
    uint32_t chunks;
    uint32_t summary_mode;      /* 0 dense correlation, 1 run summaries, 2 run summaries + some chunks repeated densely */
    uint32_t dense_chunks;      /* summary mode: logical chunks whose peak search was repeated on a dense correlation */
} am_stats;

/* kernel classes for the optional per-kernel device timing */
enum {
    AM_K_COL_FWD = 0, AM_K_ROW = 1, AM_K_COL_INV = 2, AM_K_SMALL = 3, AM_K_DIRECT = 4, AM_K_TILE_MINMAX = 5,
    AM_K_CHUNK_PEAKS = 6, AM_K_SPECTRUM = 7, AM_KERNEL_CLASSES = 8
};
typedef struct {
    int kernel_class;
    uint64_t launches;
    double total_ms;       /* sum of cudaEvent-bracketed launch durations on the matcher's stream */
    char name[24];
} am_kernel_time;

typedef struct am_matcher am_matcher;

/* Optional progress callback, the C shape of the two Once callbacks per chunk the reference feeds its progress
 * bar with (audio_matcher.rs:102-117,129).  Fired from the thread that called am_calc_chunks*: phase 0 when the
 * work of logical chunks [first_chunk, first_chunk + n_chunks) has been submitted to the GPU (one call per
 * segment), phase 1 once when the peaks of the whole call are final.  Must not call back into the library. */
typedef void (*am_progress_fn)(void *user, int phase, size_t first_chunk, size_t n_chunks);

/* Multi-GPU communicator (one process per GPU): wraps an NCCL communicator that am_calc_chunks_sharded uses for
 * the one exchange step of the path, the merge of per-shard peak candidates (audio_matcher.rs:132-139).  NCCL is
 * loaded at run time (libnccl.so.2, the copy already in the process if there is one). */
typedef struct am_comm am_comm;
#define AM_COMM_ID_BYTES 128   /* sizeof(ncclUniqueId) */

const char *am_last_error(void);
int am_abi_version(void);
/* number of visible CUDA devices (0 without a driver); never fails */
int am_device_count(void);

void am_config_default(am_config *cfg);

/* LibConvolve::new (audio_matcher.rs:289): takes a copy of the snippet, uploads it to the
 * CURRENT CUDA device, computes sum(s^2) and the conjugate snippet spectrum once. */
am_status am_matcher_create(const float *snippet, size_t m, uint32_t sr, const am_config *cfg, am_matcher **out);
/* same from 16-bit PCM (channels 1 or 2), scaled/downmixed like mp3_reader.rs:35 */
am_status am_matcher_create_pcm16(const int16_t *pcm, size_t frames, int channels, uint32_t sr,
                                  const am_config *cfg, am_matcher **out);
/* Batch of n_snippets snippets of m samples each (row-major [n_snippets][m]) matched against the same
 * stream: the stream-side transforms are shared, multiply + inverse + peak search run per snippet.
 * The reference matches one snippet per run (src/matcher/mod.rs:29-34); results are defined as
 * n_snippets independent runs, reported snippet-major with am_peak.snippet_id set. */
am_status am_matcher_create_batch(const float *snippets, size_t m, size_t n_snippets, uint32_t sr,
                                  const am_config *cfg, am_matcher **out);
size_t am_matcher_snippet_count(const am_matcher *h);
/* snippet used by am_correlate / am_inverse_sample_auto_correlation (default 0) */
am_status am_matcher_select_snippet(am_matcher *h, size_t snippet_id);
void am_matcher_destroy(am_matcher *h);

/* run this matcher's work on a caller-owned cudaStream_t (NULL = the legacy default stream) */
am_status am_matcher_set_stream(am_matcher *h, void *cuda_stream);
am_status am_matcher_set_config(am_matcher *h, const am_config *cfg);
am_status am_matcher_get_stats(const am_matcher *h, am_stats *out);
/* on != 0: bracket every kernel launch with cudaEvents on the matcher's stream and accumulate
 * per-class totals (reset by this call); read them with am_matcher_get_kernel_times */
am_status am_matcher_set_profiling(am_matcher *h, int on);
am_status am_matcher_get_kernel_times(const am_matcher *h, am_kernel_time *out, size_t cap, size_t *n_out);

/* progress callback (NULL = none, the default) */
am_status am_matcher_set_progress(am_matcher *h, am_progress_fn fn, void *user);

/* CorrelateAlgo::inverse_sample_auto_correlation (audio_matcher.rs:66,321-329): 1 / sum(s^2) */
am_status am_inverse_sample_auto_correlation(am_matcher *h, float *out);

/* output length of a correlation of n stream samples with m snippet samples */
size_t am_out_len(size_t n, size_t m, am_mode mode);
/* the Valid case, V = n - m + 1 (0 when n < m): the outputs one logical chunk of n samples contributes
 * (audio_matcher.rs:121) */
size_t am_valid_len(size_t n, size_t m);

/* CorrelateAlgo::correlate_with_sample (audio_matcher.rs:67-72): writes the correlation
 * (scaled by 1/sum(s^2) when scale != 0, audio_matcher.rs:73-75,306-308) to `out`.
 * `within` and `out` live in within_mem / out_mem. */
am_status am_correlate(am_matcher *h, const void *within, size_t n, am_sample_fmt fmt, am_mem within_mem,
                       am_mode mode, int scale, float *out, size_t cap, am_mem out_mem, size_t *out_len);

/* number of logical chunks calc_chunks makes of a stream (audio_matcher.rs:99-104) */
size_t am_num_chunks(const am_matcher *h, size_t frames);

/* chunk geometry in samples as calc_chunks derives it (audio_matcher.rs:99-100): chunk = round(chunk_size * sr),
 * overlap = round(overlap_length * sr), f64::round = half away from zero */
am_status am_chunk_geometry(const am_matcher *h, size_t *chunk, size_t *overlap);
/* frames [*lo, *hi) of a stream of total_frames frames that logical chunks [first_chunk, first_chunk + num_chunks)
 * read: the chunks themselves plus the overlap halo (the shard a rank has to hold) */
am_status am_shard_frames(const am_matcher *h, size_t total_frames, size_t first_chunk, size_t num_chunks, size_t *lo,
                          size_t *hi);

/* calc_chunks (audio_matcher.rs:88-141): the fused path -- correlation, per-chunk peak
 * finding (min prominence, min distance), global sort and neighbour filter.  Peaks come
 * back sorted by start. */
am_status am_calc_chunks(am_matcher *h, const void *stream, size_t frames, am_sample_fmt fmt, am_mem mem,
                         int scale, am_peak *out, size_t cap, size_t *n_out);

/* Shard entry point: logical chunks [first_chunk, first_chunk + num_chunks) of a stream of
 * total_frames frames.  `stream` holds frames [buf_first_frame, buf_first_frame + buf_frames)
 * of that stream (frames outside the buffer but inside the stream are an error if a chunk
 * needs them).  final_filter == 0 returns the per-chunk peaks before the global sort and
 * neighbour filter (audio_matcher.rs:132-139) so that shards can be merged with am_merge_peaks. */
am_status am_calc_chunks_range(am_matcher *h, const void *stream, size_t buf_first_frame, size_t buf_frames,
                               size_t total_frames, am_sample_fmt fmt, am_mem mem, int scale,
                               size_t first_chunk, size_t num_chunks, int final_filter,
                               am_peak *out, size_t cap, size_t *n_out);

/* global stable sort by start + filter_surrounding/is_overshadowed (audio_matcher.rs:135-160)
 * over peaks gathered from all shards; host-only, `peaks` is reordered in place. */
am_status am_merge_peaks(am_peak *peaks, size_t n, uint32_t sr, double distance_s, am_peak *out, size_t cap,
                         size_t *n_out);

/* Several files per call: the loop over args.within of the reference (src/matcher/mod.rs:42-99, one calc_chunks per
 * file against the same snippet).  streams[f] holds the frames[f] frames of file f (all files in `fmt` / `mem`).
 * The work of all files is queued before the first result is read back, so the upload of file f+1 overlaps the
 * kernels of file f and short files keep the GPU busy.  Peaks of file f are out[sum(n_out[0..f-1]) ...), n_out[f]
 * of them, each list exactly what am_calc_chunks returns for that file; n_out has n_files entries and is filled
 * even when the total exceeds cap (AM_ERR_CAPACITY). */
am_status am_calc_chunks_files(am_matcher *h, size_t n_files, const void *const *streams, const size_t *frames,
                               am_sample_fmt fmt, am_mem mem, int scale, am_peak *out, size_t cap, size_t *n_out);

/* ---- push session: calc_chunks for a stream that arrives piece by piece --------------------------------------
 * The reference hands calc_chunks a lazy iterator of decoded frames with a claimed length (mp3_reader.rs:13-66,
 * matcher/mod.rs:71-83).  A decoder thread calls am_stream_push with whatever it has decoded (host memory of any
 * kind, any piece size; the buffer may be reused as soon as the call returns); the library collects the frames in a
 * pinned ring, uploads them with one asynchronous copy per 8 MB and launches the transforms + peak search of a
 * segment of logical chunks as soon as its frames are complete, so matching and upload overlap decoding.
 * max_frames: upper bound of the stream length (the claimed length, mod.rs:78); the stream's true length is what has
 * been pushed when am_stream_finish is called.  Results equal am_calc_chunks on the concatenated frames.
 * Between begin and finish/abort the matcher's other compute entry points fail with AM_ERR_INVALID. */
typedef struct am_stream_session am_stream_session;
am_status am_stream_begin(am_matcher *h, size_t max_frames, am_sample_fmt fmt, int scale, am_stream_session **out);
am_status am_stream_push(am_stream_session *s, const void *pcm, size_t frames);
/* runs the last (partial) segment, the global sort + neighbour filter, and closes the session (also on error) */
am_status am_stream_finish(am_stream_session *s, am_peak *out, size_t cap, size_t *n_out);
void am_stream_abort(am_stream_session *s);

/* ---- multi-GPU: one process per GPU, chunk-range shards, one all-gather of peak candidates ----------------
 * am_comm_get_unique_id: rank 0 creates the 128-byte id and hands it to the other ranks by any means (file, MPI,
 * torch.distributed ...); am_comm_init is collective over the nranks processes and binds the communicator to the
 * CURRENT CUDA device. */
am_status am_comm_get_unique_id(void *id_out /* AM_COMM_ID_BYTES */);
am_status am_comm_init(int nranks, int rank, const void *nccl_unique_id, am_comm **out);
void am_comm_destroy(am_comm *comm);
int am_comm_rank(const am_comm *comm);
int am_comm_size(const am_comm *comm);
/* calc_chunks over a stream sharded by logical-chunk ranges: this rank runs chunks [first_chunk, first_chunk +
 * num_chunks) on the frames it holds (am_shard_frames: its chunks + the overlap halo, no halo exchange); the
 * device-resident peak lists and counts of all ranks are exchanged with ONE ncclAllGather of fixed-size records
 * (a second one only if a rank overflows the record), and every rank applies the global sort + neighbour filter
 * (am_merge_peaks) and gets the complete result.  Collective: every rank of `comm` must call it. */
am_status am_calc_chunks_sharded(am_matcher *h, am_comm *comm, const void *stream, size_t buf_first_frame,
                                 size_t buf_frames, size_t total_frames, am_sample_fmt fmt, am_mem mem, int scale,
                                 size_t first_chunk, size_t num_chunks, am_peak *out, size_t cap, size_t *n_out);

/* is_overshadowed (audio_matcher.rs:143-160); other == NULL is None */
int am_is_overshadowed(const am_peak *element, const am_peak *other, uint32_t sr, double max_distance_s);

/* Test hook: the per-chunk peak kernels on a caller-supplied correlation (one segment starting at chunk 0, chunk
 * geometry from the matcher's config and snippet length).  summary != 0 uses the run-record path with every run the
 * transform kernels would not have stored poisoned; *mode_out: 0 dense, 1 run records, 2 records rejected -> dense.
 * Peaks are returned before the global sort / neighbour filter, ordered by (chunk, height descending). */
am_status am_debug_peaks_from_correlation(am_matcher *h, const float *c_host, size_t n, int summary, am_peak *out,
                                          size_t cap, size_t *n_out, uint32_t *mode_out);

/* ---- synthetic workload generator (bench/test utility) -------------------
 * Device-side, integer-only, bit-identical to the oracle's generator.
 * out[i] = int16((hash64(seed, first + i) >> 50) - 8192) */
am_status am_synth_pcm16_device(uint64_t seed, uint64_t first, size_t count, int16_t *dev_out, void *cuda_stream);
/* coloured (low-pass) noise for the loud / coloured-programme bench legs:
 * out[i] = sat16((sum_{k < taps} (int16((hash64(seed, first + i - k) >> 50) - 8192)) * mul) >> shift) */
am_status am_synth_coloured_pcm16_device(uint64_t seed, uint64_t first, size_t count, int taps, int mul, int shift,
                                         int16_t *dev_out, void *cuda_stream);
/* x[(offset + j)*channels + c] = sat16((x >> 1) + (snip[j] >> shift)), j < m, frame < frames */
am_status am_synth_plant_device(int16_t *dev_pcm, size_t frames, int channels, const int16_t *dev_snip, size_t m,
                                uint64_t offset, int shift, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* AUDIO_MATCHER_B200_H */
