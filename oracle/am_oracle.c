/* am_oracle.c -- CPU restatement of audio-matcher's snippet-vs-stream hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by
 * or executed from the product library (audio_matcher_b200/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's shared object, and only as the checker or as the timed
 * CPU baseline.
 *
 * The Rust reference cannot be compiled here (no cargo/rustc; four private git
 * dependencies, Cargo.toml:9-12), so this is a "port"-kind oracle.  Every
 * function cites the reference lines it restates (paths relative to
 * /root/reference).  Arithmetic that lives in third-party crates absent from
 * the reference tree is restated from their published behaviour:
 *   fftconvolve 0.1  fftcorrelate(): reverse the kernel, zero-pad both inputs to
 *                    n+m-1, complex FFT, multiply, inverse FFT, 1/N, crop
 *                    (call site src/matcher/audio_matcher.rs:305)
 *   realfft 3.3 / rustfft: unnormalised DFT of arbitrary length (orc_fft.inc)
 *   find_peaks 0.1   PeakFinder: scipy-style local maxima with plateaus,
 *                    prominence, min-distance, output sorted by height
 *                    (call site src/matcher/audio_matcher.rs:221-230)
 *   common (private) chunked(window, step), filter_surrounding(), with_size()
 *                    (call sites src/matcher/audio_matcher.rs:104,136)
 *
 * PINNED by the reference's own tests (tests/test_oracle_kat.py):
 *   correlate KAT   audio_matcher.rs:489-517   [-52,-46,...,50]
 *   find_peaks KAT  audio_matcher.rs:167-185   order (3,5,1), prom (1,.3,.2)
 *   overshadow KATs audio_matcher.rs:187-218
 * UNPINNED (no reference test exists; this file defines the behaviour, see
 * DESIGN.md "parity unpinned" list): min-distance rule, plateau midpoint,
 * array-edge handling, chunked() tail windows, filter_surrounding neighbours,
 * Duration::from_secs_f64 truncation, tie order among equal heights.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define REAL float
#define SUF f
#include "orc_fft.inc"
#undef REAL
#undef SUF
#define REAL double
#define SUF d
#include "orc_fft.inc"
#undef REAL
#undef SUF

/* ---- types shared with the ctypes binding (oracle/am_oracle.py) ---------- */

enum { ORC_MODE_FULL = 0, ORC_MODE_SAME = 1, ORC_MODE_VALID = 2 }; /* audio_matcher.rs:54-59 */

typedef struct {           /* find_peaks::Peak<f32>, positions global after calc_chunks */
    uint64_t start, end;   /* position: Range<usize>, end exclusive */
    float height, prominence, left_diff, right_diff;
    uint32_t chunk;        /* logical chunk that produced it (diagnostic) */
    uint32_t _pad;
} orc_peak;

typedef struct {           /* Config + PeakConfig, audio_matcher.rs:24-53 */
    double chunk_size_s;   /* args.rs:71 default 60 */
    double overlap_s;      /* = snippet duration, audio_matcher.rs:41 */
    double distance_s;     /* args.rs:75 default 480 */
    float prominence;      /* args.prominence / 100, audio_matcher.rs:44 */
    float _pad;
} orc_config;

/* ---- PCM scale / downmix: src/matcher/mp3_reader.rs:12,33-36 -------------- */

void orc_pcm16_to_f32(const int16_t *pcm, size_t frames, int channels, float *out) {
    const float pcm_factor = 1.0f / (float)((1 << 16) - 1);          /* :12 */
    if (channels == 2) {
        for (size_t i = 0; i < frames; ++i)                           /* :35 */
            out[i] = ((float)pcm[2 * i] + (float)pcm[2 * i + 1]) * 0.5f * pcm_factor;
    } else {                                                          /* mono == (l = r = s) */
        for (size_t i = 0; i < frames; ++i)
            out[i] = ((float)pcm[i] + (float)pcm[i]) * 0.5f * pcm_factor;
    }
}

/* ---- correlation: LibConvolve::correlate audio_matcher.rs:297-310 and the
 *      normative maths of MyConvolve::correlate :414-464 ---------------------- */

size_t orc_out_len(size_t n, size_t m, int mode) {
    if (n == 0 || m == 0) return 0;
    switch (mode) {
    case ORC_MODE_FULL: return n + m - 1;
    case ORC_MODE_SAME: return n;
    default: return n >= m ? n - m + 1 : 0;       /* n < m: no valid outputs (unpinned) */
    }
}

static size_t crop_start(size_t n, size_t m, int mode) {   /* centered(): :460-464 */
    size_t full = n + m - 1;
    return (full - orc_out_len(n, m, mode)) / 2;            /* Valid: m-1, Same: (m-1)/2 */
}

#define DEF_CORRELATE(REAL, SUF, CPX, PLAN)                                                       \
    int orc_correlate_##SUF(const REAL *within, size_t n, const REAL *sample, size_t m, int mode,  \
                            REAL *out) {                                                           \
        size_t olen = orc_out_len(n, m, mode);                                                     \
        if (olen == 0) return 0;                                                                   \
        size_t N = n + m - 1;                       /* pad_len, :421 */                            \
        PLAN *p = plan_new_##SUF(N);                                                               \
        if (!p) return -1;                                                                         \
        CPX *a = (CPX *)calloc(N, sizeof(CPX)), *b = (CPX *)calloc(N, sizeof(CPX));                 \
        CPX *w = (CPX *)malloc(2 * p->m * sizeof(CPX));                                            \
        if (!a || !b || !w) { free(a); free(b); free(w); plan_free_##SUF(p); return -1; }          \
        /* fftcorrelate == convolve with the reversed kernel: within padded at the back,           \
         * reversed sample padded at the back, plain product (the "reverse mult" variant           \
         * :424-426 is the same thing and the conj variant :432-434 equals it, pinned :489-506) */ \
        for (size_t i = 0; i < n; ++i) a[i].re = within[i];                                        \
        for (size_t j = 0; j < m; ++j) b[j].re = sample[m - 1 - j];                                \
        fft_any_##SUF(p, a, 0, w);                                                                 \
        fft_any_##SUF(p, b, 0, w);                                                                 \
        for (size_t k = 0; k < N; ++k) a[k] = cmul_##SUF(a[k], b[k]);                              \
        fft_any_##SUF(p, a, 1, w);                                                                 \
        REAL inv = (REAL)(1.0 / (double)N);          /* :442 */                                    \
        size_t s0 = crop_start(n, m, mode);                                                        \
        for (size_t k = 0; k < olen; ++k) out[k] = a[s0 + k].re * inv;                             \
        free(a); free(b); free(w); plan_free_##SUF(p);                                             \
        return 0;                                                                                  \
    }
DEF_CORRELATE(float, f, cpx_f, plan_f)
DEF_CORRELATE(double, d, cpx_d, plan_d)

/* O(n*m) definition of the same thing, double accumulation: the ground truth
 * for small cases (full[k] = sum_j xpad[k+j] s[j], xpad = x with m-1 zeros on
 * both sides; Valid/Same are crops of Full). */
int orc_correlate_direct(const float *within, size_t n, const float *sample, size_t m, int mode,
                         double *out) {
    size_t olen = orc_out_len(n, m, mode);
    size_t s0 = olen ? crop_start(n, m, mode) : 0;
    for (size_t k = 0; k < olen; ++k) {
        double acc = 0.0;
        size_t f = s0 + k;                           /* index into Full */
        for (size_t j = 0; j < m; ++j) {
            /* xpad index f + j  ->  x index f + j - (m-1) */
            size_t xi = f + j;
            if (xi < m - 1) continue;
            xi -= m - 1;
            if (xi >= n) break;
            acc += (double)within[xi] * (double)sample[j];
        }
        out[k] = acc;
    }
    return 0;
}


/* "Optimised CPU" variant for the baseline discussion (second CPU line): same semantics as
 * orc_correlate_f but with power-of-two transforms and a snippet spectrum computed once, so that the GPU speed-up
 * is not inflated by the reference's exact-length (often prime) transforms and per-chunk snippet FFT
 * (audio_matcher.rs:305).  Valid mode only. */
typedef struct { size_t N; cpx_f *spec; plan_f *plan; } orc_pow2_cache;

static orc_pow2_cache *pow2_cache_new(const float *sample, size_t m, size_t N) {
    orc_pow2_cache *c = (orc_pow2_cache *)calloc(1, sizeof *c);
    c->N = N;
    c->plan = plan_new_f(N);
    c->spec = (cpx_f *)calloc(N, sizeof(cpx_f));
    cpx_f *w = (cpx_f *)malloc(2 * N * sizeof(cpx_f));
    for (size_t j = 0; j < m; ++j) c->spec[j].re = sample[j];
    fft_any_f(c->plan, c->spec, 0, w);
    for (size_t k = 0; k < N; ++k) c->spec[k].im = -c->spec[k].im;       /* conj: correlation */
    free(w);
    return c;
}
static void pow2_cache_free(orc_pow2_cache *c) {
    if (!c) return;
    plan_free_f(c->plan); free(c->spec); free(c);
}
static void correlate_pow2_valid(const orc_pow2_cache *c, const float *within, size_t n, size_t m, float *out) {
    size_t N = c->N, V = n - m + 1;
    cpx_f *a = (cpx_f *)calloc(N, sizeof(cpx_f)), *w = (cpx_f *)malloc(2 * N * sizeof(cpx_f));
    for (size_t i = 0; i < n; ++i) a[i].re = within[i];
    fft_any_f(c->plan, a, 0, w);
    for (size_t k = 0; k < N; ++k) a[k] = cmul_f(a[k], c->spec[k]);
    fft_any_f(c->plan, a, 1, w);
    float inv = (float)(1.0 / (double)N);
    for (size_t k = 0; k < V; ++k) out[k] = a[k].re * inv;               /* N >= n + m - 1: no wrap-around */
    free(a); free(w);
}

/* inverse_sample_auto_correlation: audio_matcher.rs:321-329
 * 1 / fftcorrelate(sample, sample, Valid)[0]  (f32 transform like the reference) */
float orc_inv_autocorr_f32(const float *sample, size_t m) {
    float v = 0.f;
    if (orc_correlate_f(sample, m, sample, m, ORC_MODE_VALID, &v) != 0) return NAN;
    return 1.0f / v;
}
/* same quantity from the definition sum s^2 in double */
double orc_inv_autocorr_exact(const float *sample, size_t m) {
    double acc = 0.0;
    for (size_t j = 0; j < m; ++j) acc += (double)sample[j] * (double)sample[j];
    return 1.0 / acc;
}

/* CorrelateAlgo::scale: audio_matcher.rs:73-75, scale_slice :246-252 */
void orc_scale_f32(float *data, size_t len, float factor) {
    for (size_t i = 0; i < len; ++i) data[i] = data[i] * factor;
}

/* ---- find_peaks 0.1 PeakFinder, as configured at audio_matcher.rs:221-230 -- */

static int cmp_height_desc(const void *pa, const void *pb) {
    const orc_peak *a = (const orc_peak *)pa, *b = (const orc_peak *)pb;
    if (a->height > b->height) return -1;
    if (a->height < b->height) return 1;
    return (a->start > b->start) - (a->start < b->start);   /* tie: lower position first (unpinned) */
}

static inline uint64_t peak_mid(const orc_peak *p) { return (p->start + p->end) / 2; }

/* Returns the number of peaks found (may exceed cap; only cap are written).
 * use_prom == 0 skips prominence entirely (not used by the reference). */
size_t orc_find_peaks(const float *y, size_t len, int use_prom, float min_prom, size_t min_dist,
                      orc_peak *out, size_t cap) {
    if (len < 3) return 0;                          /* endpoints are never peaks (unpinned) */
    size_t np = 0, acap = 64;
    orc_peak *pk = (orc_peak *)malloc(acap * sizeof(orc_peak));
    size_t i = 1, imax = len - 1;
    while (i < imax) {
        if (y[i - 1] < y[i]) {
            size_t ahead = i + 1;
            while (ahead < imax && y[ahead] == y[i]) ++ahead;       /* plateau */
            if (y[ahead] < y[i]) {
                float h = y[i];
                orc_peak p;
                memset(&p, 0, sizeof p);
                p.start = i; p.end = ahead;                          /* Range: end exclusive */
                p.height = h;
                p.left_diff = h - y[i - 1];
                p.right_diff = h - y[ahead];
                int keep = 1;
                if (use_prom) {
                    /* prominence = height - max(lowest point on the way to a strictly
                     * higher sample on each side); pinned 1.0/0.3/0.2 at :167-185 */
                    float lmin = h, rmin = h;
                    for (size_t l = i; l-- > 0;) { if (y[l] > h) break; if (y[l] < lmin) lmin = y[l]; }
                    for (size_t r = ahead; r < len; ++r) { if (y[r] > h) break; if (y[r] < rmin) rmin = y[r]; }
                    p.prominence = h - (lmin > rmin ? lmin : rmin);
                    keep = p.prominence >= min_prom;                 /* with_min_prominence :227 */
                }
                if (keep) {
                    if (np == acap) { acap *= 2; pk = (orc_peak *)realloc(pk, acap * sizeof(orc_peak)); }
                    pk[np++] = p;
                }
                i = ahead;
            }
        }
        ++i;
    }
    qsort(pk, np, sizeof(orc_peak), cmp_height_desc);   /* output ordered by height, pinned :172-182 */
    size_t nk = 0;
    if (min_dist > 0) {                                 /* with_min_distance :228 (greedy, unpinned) */
        for (size_t a = 0; a < np; ++a) {
            int ok = 1;
            uint64_t ma = peak_mid(&pk[a]);
            for (size_t b = 0; b < nk && ok; ++b) {
                uint64_t mb = peak_mid(&pk[b]);
                uint64_t d = ma > mb ? ma - mb : mb - ma;
                if (d < min_dist) ok = 0;
            }
            if (ok) pk[nk++] = pk[a];
        }
    } else nk = np;
    for (size_t a = 0; a < nk && a < cap; ++a) out[a] = pk[a];
    free(pk);
    return nk;
}

/* ---- Duration::from_secs_f64(start / sr): src/matcher/mod.rs:127-129 ------
 * exact truncation of the f64 quotient to whole nanoseconds (unpinned: Rust
 * >= 1.63 truncates; 1.60-1.62 rounded to nearest). */
static uint64_t secs_f64_to_ns(double v) {
    if (!(v > 0.0)) return 0;
    int e;
    double fr = frexp(v, &e);                        /* v = fr * 2^e, fr in [0.5,1) */
    uint64_t mant = (uint64_t)ldexp(fr, 53);         /* exact 53-bit integer */
    e -= 53;                                         /* v = mant * 2^e */
    unsigned __int128 t = (unsigned __int128)mant * 1000000000ull;
    if (e >= 0) return (uint64_t)(t << e);
    if (-e >= 127) return 0;
    return (uint64_t)(t >> (-e));
}
uint64_t orc_start_ns(uint64_t start, uint32_t sr) { return secs_f64_to_ns((double)start / (double)sr); }
static uint64_t duration_ns(double secs) { return secs <= 0 ? 0 : (uint64_t)llround(secs * 1e9); }

/* is_overshadowed: audio_matcher.rs:143-160 (other == NULL is None) */
int orc_is_overshadowed(const orc_peak *element, const orc_peak *other, uint32_t sr, double max_distance_s) {
    if (!other) return 0;
    uint64_t e = orc_start_ns(element->start, sr), b = orc_start_ns(other->start, sr);
    if (e < b) { uint64_t t = e; e = b; b = t; }
    return ((e - b) < duration_ns(max_distance_s)) && (other->prominence > element->prominence);
}

/* ---- calc_chunks: audio_matcher.rs:88-141 --------------------------------- */

static int cmp_start(const void *pa, const void *pb) {   /* stable via (start, chunk, seq in chunk) */
    const orc_peak *a = (const orc_peak *)pa, *b = (const orc_peak *)pb;
    if (a->start != b->start) return a->start < b->start ? -1 : 1;
    if (a->chunk != b->chunk) return a->chunk < b->chunk ? -1 : 1;
    return a->_pad < b->_pad ? -1 : (a->_pad > b->_pad);
}

size_t orc_num_chunks(size_t L, uint32_t sr, const orc_config *cfg) {
    size_t C = (size_t)llround(cfg->chunk_size_s * (double)sr);
    if (C == 0 || L == 0) return 0;
    return (L + C - 1) / C;                          /* chunked(C+ov, C): a window starts at every C*i < L */
}

/* precision: 32 -> f32 exact-length FFT like the reference (also the timed CPU
 * baseline); 64 -> f64 FFT, result rounded to f32 before peak finding;
 * 0 -> direct O(n*m) sums in double (small cases only).
 * first_chunk/num_chunks restrict the work (bounded baseline samples, shards);
 * final_filter == 0 returns the per-chunk peaks before sort + filter_surrounding. */
size_t orc_calc_chunks_range(const float *stream, size_t L, const float *snippet, size_t m, uint32_t sr,
                             const orc_config *cfg, int scale, int precision, int threads,
                             size_t first_chunk, size_t num_chunks, int final_filter,
                             orc_peak *out, size_t cap) {
    size_t ov = (size_t)llround(cfg->overlap_s * (double)sr);       /* :99 */
    size_t C = (size_t)llround(cfg->chunk_size_s * (double)sr);     /* :100 */
    size_t total = orc_num_chunks(L, sr, cfg);
    if (first_chunk >= total) return 0;
    if (num_chunks > total - first_chunk) num_chunks = total - first_chunk;
    size_t min_dist = (size_t)((uint64_t)cfg->distance_s) * (size_t)sr;  /* as_secs() truncates, :228 */
    float inv_ac = 1.0f;
    if (scale) inv_ac = (precision == 32) ? orc_inv_autocorr_f32(snippet, m)
                                          : (float)orc_inv_autocorr_exact(snippet, m);
    /* precision 33: "optimised CPU" -- power-of-two f32 transforms, snippet spectrum cached for full windows */
    orc_pow2_cache *pcache = NULL;
    if (precision == 33) {
        size_t N = 1;
        while (N < C + ov + m - 1) N <<= 1;
        pcache = pow2_cache_new(snippet, m, N);
    }

    orc_peak **lists = (orc_peak **)calloc(num_chunks, sizeof(orc_peak *));
    size_t *counts = (size_t *)calloc(num_chunks, sizeof(size_t));
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (size_t ci = 0; ci < num_chunks; ++ci) {                    /* rayon par_bridge :114 */
        size_t i = first_chunk + ci;
        size_t offset = C * i;                                      /* :119 */
        size_t n = L - offset < C + ov ? L - offset : C + ov;       /* chunked window */
        size_t V = orc_out_len(n, m, ORC_MODE_VALID);
        if (V == 0) continue;
        float *c = (float *)malloc(V * sizeof(float));
        if (precision == 32) {
            orc_correlate_f(stream + offset, n, snippet, m, ORC_MODE_VALID, c);
        } else if (precision == 33) {
            correlate_pow2_valid(pcache, stream + offset, n, m, c);          /* short tail windows reuse the big plan */
        } else {
            double *cd = (double *)malloc(V * sizeof(double));
            if (precision == 64) {
                double *wd = (double *)malloc(n * sizeof(double)), *sd = (double *)malloc(m * sizeof(double));
                for (size_t k = 0; k < n; ++k) wd[k] = stream[offset + k];
                for (size_t k = 0; k < m; ++k) sd[k] = snippet[k];
                orc_correlate_d(wd, n, sd, m, ORC_MODE_VALID, cd);
                free(wd); free(sd);
            } else {
                orc_correlate_direct(stream + offset, n, snippet, m, ORC_MODE_VALID, cd);
            }
            for (size_t k = 0; k < V; ++k) c[k] = (float)cd[k];
            free(cd);
        }
        if (scale) orc_scale_f32(c, V, inv_ac);                     /* :306-308 */
        size_t pcap = V / 2 + 1;
        orc_peak *pk = (orc_peak *)malloc(pcap * sizeof(orc_peak));
        size_t np = orc_find_peaks(c, V, 1, cfg->prominence, min_dist, pk, pcap);   /* :124 */
        for (size_t k = 0; k < np; ++k) {                           /* offset_range, lib.rs:8-10 */
            pk[k].start += offset; pk[k].end += offset; pk[k].chunk = (uint32_t)i; pk[k]._pad = (uint32_t)k;
        }
        lists[ci] = pk; counts[ci] = np;
        free(c);
    }
    size_t tot = 0;
    for (size_t ci = 0; ci < num_chunks; ++ci) tot += counts[ci];
    orc_peak *all = (orc_peak *)malloc((tot ? tot : 1) * sizeof(orc_peak));
    size_t w = 0;
    for (size_t ci = 0; ci < num_chunks; ++ci) {
        if (counts[ci]) memcpy(all + w, lists[ci], counts[ci] * sizeof(orc_peak));
        w += counts[ci];
        free(lists[ci]);
    }
    free(lists); free(counts);
    pow2_cache_free(pcache);
    size_t nout = 0;
    if (!final_filter) {
        for (size_t k = 0; k < tot; ++k) { if (nout < cap) out[nout] = all[k]; ++nout; }
    } else {
        qsort(all, tot, sizeof(orc_peak), cmp_start);               /* sorted_by start, stable :135 */
        for (size_t k = 0; k < tot; ++k) {                          /* filter_surrounding :136-139 */
            const orc_peak *before = k > 0 ? &all[k - 1] : NULL;
            const orc_peak *after = k + 1 < tot ? &all[k + 1] : NULL;
            if (orc_is_overshadowed(&all[k], before, sr, cfg->distance_s) ||
                orc_is_overshadowed(&all[k], after, sr, cfg->distance_s))
                continue;
            if (nout < cap) out[nout] = all[k];
            ++nout;
        }
    }
    free(all);
    return nout;
}

size_t orc_calc_chunks(const float *stream, size_t L, const float *snippet, size_t m, uint32_t sr,
                       const orc_config *cfg, int scale, int precision, int threads,
                       orc_peak *out, size_t cap) {
    return orc_calc_chunks_range(stream, L, snippet, m, sr, cfg, scale, precision, threads,
                                 0, (size_t)-1, 1, out, cap);
}

/* final sort + filter_surrounding over an arbitrary peak list (merge of shards) */
size_t orc_merge_peaks(orc_peak *all, size_t tot, uint32_t sr, double distance_s, orc_peak *out, size_t cap) {
    qsort(all, tot, sizeof(orc_peak), cmp_start);
    size_t nout = 0;
    for (size_t k = 0; k < tot; ++k) {
        const orc_peak *before = k > 0 ? &all[k - 1] : NULL;
        const orc_peak *after = k + 1 < tot ? &all[k + 1] : NULL;
        if (orc_is_overshadowed(&all[k], before, sr, distance_s) ||
            orc_is_overshadowed(&all[k], after, sr, distance_s))
            continue;
        if (nout < cap) out[nout] = all[k];
        ++nout;
    }
    return nout;
}

/* ---- synthetic inputs (integer-only so every
 *      implementation produces identical PCM) -------------------------------- */

static inline uint64_t hash64(uint64_t seed, uint64_t n) {  /* SplitMix64 finaliser */
    uint64_t z = seed + n * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
uint64_t orc_hash64(uint64_t seed, uint64_t n) { return hash64(seed, n); }

/* out[i] = int16((hash64(seed, first + i) >> 50) - 8192), i < count */
void orc_synth_pcm16(uint64_t seed, uint64_t first, size_t count, int16_t *out) {
    for (size_t i = 0; i < count; ++i) out[i] = (int16_t)((int)(hash64(seed, first + i) >> 50) - 8192);
}

/* plant: x[o + j] = sat16((x[o + j] >> 1) + (s[j] >> shift)) over mono pcm or,
 * for channels == 2, over both channels of frame o + j (s is mono). */
void orc_synth_plant(int16_t *pcm, size_t frames, int channels, const int16_t *snip, size_t m,
                     uint64_t offset, int shift) {
    for (size_t j = 0; j < m; ++j) {
        size_t f = offset + j;
        if (f >= frames) break;
        for (int ch = 0; ch < channels; ++ch) {
            int v = (pcm[f * channels + ch] >> 1) + (snip[j] >> shift);
            if (v > 32767) v = 32767;
            if (v < -32768) v = -32768;
            pcm[f * channels + ch] = (int16_t)v;
        }
    }
}

/* planted offsets: o_k = k*P + hash64(seed_plant, k) mod J, with o_3 = 3C-5 and
 * o_4 = 4C forced (chunk-boundary cases); gain shift cycles 0,1,2,3 */
uint64_t orc_plant_offset(uint64_t seed_plant, uint64_t k, uint64_t P, uint64_t J, uint64_t C) {
    if (k == 3) return 3 * C - 5;
    if (k == 4) return 4 * C;
    return k * P + (J ? hash64(seed_plant, k) % J : 0);
}

int orc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
