/* capi_smoke.c -- the C ABI used from plain C (no Python, no C++): what a cgo/FFI caller sees.
 * Built and run by tests/test_gpu_parity.py::test_c_abi_from_plain_c.
 *   gcc -std=c11 -I include tests/capi_smoke.c -L audio_matcher_b200 -laudio_matcher_b200 -lm
 * Reference KAT: src/matcher/audio_matcher.rs:489-517 (Valid, unscaled, [-10,10) vs [1,2,3]). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "audio_matcher.h"

static void progress_cb(void *user, int phase, size_t first_chunk, size_t n_chunks) {
    (void)first_chunk;
    if (phase == 1) *(size_t *)user += n_chunks;             /* phase 1: the peaks of the call are final */
}

int main(void) {
    if (am_abi_version() != AM_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 2; }
    if (am_device_count() < 1) { fprintf(stderr, "no CUDA device: %s\n", "cannot run"); return 77; }
    float snippet[3] = {1.f, 2.f, 3.f}, within[20], out[32];
    for (int i = 0; i < 20; ++i) within[i] = (float)(i - 10);
    am_config cfg;
    am_config_default(&cfg);
    am_matcher *h = NULL;
    if (am_matcher_create(snippet, 3, 1, &cfg, &h) != AM_OK) { fprintf(stderr, "create: %s\n", am_last_error()); return 1; }
    size_t n = 0;
    if (am_correlate(h, within, 20, AM_FMT_F32_MONO, AM_MEM_HOST, AM_MODE_VALID, 0, out, 32, AM_MEM_HOST, &n) != AM_OK) {
        fprintf(stderr, "correlate: %s\n", am_last_error());
        return 1;
    }
    if (n != 18) { fprintf(stderr, "expected 18 outputs, got %zu\n", n); return 1; }
    for (size_t k = 0; k < n; ++k)
        if (fabsf(out[k] - (float)(-52 + 6 * (int)k)) > 1.2e-5f) { fprintf(stderr, "out[%zu] = %g\n", k, out[k]); return 1; }
    float inv = 0.f;
    am_inverse_sample_auto_correlation(h, &inv);
    if (fabsf(inv - 1.0f / 14.0f) > 1e-8f) { fprintf(stderr, "inverse autocorrelation %g\n", inv); return 1; }
    /* error path: capacity too small must fail loudly, not truncate */
    if (am_correlate(h, within, 20, AM_FMT_F32_MONO, AM_MEM_HOST, AM_MODE_VALID, 0, out, 4, AM_MEM_HOST, &n) != AM_ERR_CAPACITY) {
        fprintf(stderr, "capacity error not reported\n");
        return 1;
    }
    /* calc_chunks through the ABI: 1-tap snippet, the find_peaks KAT of audio_matcher.rs:167-185 */
    float one = 1.f, y[7] = {0.f, 0.7f, 0.5f, 1.0f, 0.5f, 0.8f, 0.f};
    am_matcher *g = NULL;
    cfg.chunk_size_s = 7.0; cfg.overlap_s = 0.0; cfg.distance_s = 0.0; cfg.prominence = 0.0f;
    if (am_matcher_create(&one, 1, 1, &cfg, &g) != AM_OK) { fprintf(stderr, "create: %s\n", am_last_error()); return 1; }
    am_peak peaks[8];
    if (am_calc_chunks(g, y, 7, AM_FMT_F32_MONO, AM_MEM_HOST, 1, peaks, 8, &n) != AM_OK) {
        fprintf(stderr, "calc_chunks: %s\n", am_last_error());
        return 1;
    }
    if (n != 3 || peaks[0].start != 1 || peaks[1].start != 3 || peaks[2].start != 5 ||
        fabsf(peaks[1].prominence - 1.0f) > 1e-6f || fabsf(peaks[2].prominence - 0.3f) > 1e-6f ||
        fabsf(peaks[0].prominence - 0.2f) > 1e-6f) {
        fprintf(stderr, "find_peaks KAT failed (%zu peaks)\n", n);
        return 1;
    }
    /* the same through a push session (the lazy sample iterator of mp3_reader.rs:13-66): three pieces, claimed length 9 */
    static size_t seen_chunks;
    seen_chunks = 0;
    if (am_matcher_set_progress(g, progress_cb, &seen_chunks) != AM_OK) { fprintf(stderr, "set_progress: %s\n", am_last_error()); return 1; }
    am_stream_session *s = NULL;
    if (am_stream_begin(g, 9, AM_FMT_F32_MONO, 1, &s) != AM_OK) { fprintf(stderr, "stream_begin: %s\n", am_last_error()); return 1; }
    if (am_calc_chunks(g, y, 7, AM_FMT_F32_MONO, AM_MEM_HOST, 1, peaks, 8, &n) != AM_ERR_INVALID) { fprintf(stderr, "busy handle not reported\n"); return 1; }
    if (am_stream_push(s, y, 2) != AM_OK || am_stream_push(s, y + 2, 4) != AM_OK || am_stream_push(s, y + 6, 1) != AM_OK) {
        fprintf(stderr, "stream_push: %s\n", am_last_error());
        return 1;
    }
    am_peak streamed[8];
    size_t ns = 0;
    if (am_stream_finish(s, streamed, 8, &ns) != AM_OK) { fprintf(stderr, "stream_finish: %s\n", am_last_error()); return 1; }
    if (ns != 3 || streamed[0].start != 1 || streamed[1].start != 3 || streamed[2].start != 5 || streamed[1].prominence != peaks[1].prominence) {
        fprintf(stderr, "push session differs from the one-shot call (%zu peaks)\n", ns);
        return 1;
    }
    if (seen_chunks != 1) { fprintf(stderr, "progress callback saw %zu chunks\n", seen_chunks); return 1; }
    /* chunk geometry and shard frames as the library rounds them (audio_matcher.rs:99-100) */
    size_t C = 0, ov = 0, lo = 0, hi = 0;
    if (am_chunk_geometry(g, &C, &ov) != AM_OK || C != 7 || ov != 0) { fprintf(stderr, "chunk geometry %zu %zu\n", C, ov); return 1; }
    if (am_shard_frames(g, 100, 3, 2, &lo, &hi) != AM_OK || lo != 21 || hi != 35) { fprintf(stderr, "shard frames %zu %zu\n", lo, hi); return 1; }
    /* one rank is a valid communicator-less call of the sharded entry point */
    if (am_calc_chunks_sharded(g, NULL, y, 0, 7, 7, AM_FMT_F32_MONO, AM_MEM_HOST, 1, 0, 1, peaks, 8, &n) != AM_OK || n != 3) {
        fprintf(stderr, "calc_chunks_sharded: %s\n", am_last_error());
        return 1;
    }
    /* three "files" in one call (the loop over args.within, src/matcher/mod.rs:42): the KAT, an empty file, the KAT again */
    {
        const void *files[3] = {y, NULL, y};
        size_t frames[3] = {7, 0, 7}, counts[3] = {9, 9, 9};
        am_peak multi[8];
        if (am_calc_chunks_files(g, 3, files, frames, AM_FMT_F32_MONO, AM_MEM_HOST, 1, multi, 8, counts) != AM_OK) {
            fprintf(stderr, "calc_chunks_files: %s\n", am_last_error());
            return 1;
        }
        if (counts[0] != 3 || counts[1] != 0 || counts[2] != 3 || multi[0].start != 1 || multi[2].start != 5 || multi[3].start != 1 ||
            multi[4].prominence != streamed[1].prominence) {
            fprintf(stderr, "calc_chunks_files differs from the one-file call (%zu %zu %zu)\n", counts[0], counts[1], counts[2]);
            return 1;
        }
        if (am_calc_chunks_files(g, 3, files, frames, AM_FMT_F32_MONO, AM_MEM_HOST, 1, multi, 4, counts) != AM_ERR_CAPACITY || counts[2] != 3) {
            fprintf(stderr, "calc_chunks_files: capacity overflow not reported\n");
            return 1;
        }
    }
    am_matcher_destroy(g);
    am_matcher_destroy(h);
    printf("capi_smoke ok\n");
    return 0;
}
