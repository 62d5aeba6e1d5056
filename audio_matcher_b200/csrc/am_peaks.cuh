// am_peaks.cuh -- per-logical-chunk peak extraction on the device.
//
// Restates what calc_chunks does with each chunk's correlation
// (src/matcher/audio_matcher.rs:124 -> find_peaks :221-230 -> find_peaks::PeakFinder):
// local maxima (plateau aware, endpoints excluded), prominence >= min_prominence,
// greedy min-distance suppression in descending height order.  The chunk geometry
// (window C + ov stepped by C, V = n - m + 1 outputs, audio_matcher.rs:99-104,119)
// scopes the prominence walks, so it is reproduced here index for index.
//
//   k_tile_minmax   min / max of every 1024-sample tile of every chunk (dense correlation)
//   k_tile_from_runs  the same from the per-run records of k_col_inv's summary epilogue (1/4 of the bytes)
//   k_chunk_peaks   one CTA per chunk: chunk minimum -> exact-safe candidate filter
//                   (prominence <= height - chunk_min) -> warp-cooperative prominence
//                   walks that skip whole tiles through the min/max summaries ->
//                   min-distance suppression (block argmax loop) -> append to the output.
//                   Candidates are taken in descending bands of height: a band holds at most
//                   max_peaks_per_chunk candidates, and the descent stops as soon as the kept peaks
//                   cover the chunk under the minimum distance (nothing lower can survive), so loud
//                   or tonal material with millions of local maxima per chunk costs a few bands.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace amp {

constexpr int TP_LOG2 = 10;
constexpr int TP = 1 << TP_LOG2;       // samples per summary tile

struct DevPeak {                        // mirrors am_peak (include/audio_matcher.h)
    unsigned long long start, end;
    float height, prominence, left_diff, right_diff;
    unsigned snippet_id, chunk;
};

struct ChunkGeom {
    long long c_stride;                 // floats between the correlation buffers of consecutive snippets (grid z)
    long long C, ov, m;                 // samples
    long long total;                    // virtual stream length (with lead/tail padding)
    long long first_chunk;              // global index of the segment's first chunk
    long long c_g0;                     // global offset of c[0]
    int tiles_stride;                   // tiles allocated per chunk in tmin/tmax
    long long seg_end;                  // summary mode: index (in c) one past the segment's last output
};

__device__ __forceinline__ long long chunk_valid_len(const ChunkGeom &g, long long chunk) {
    long long off = g.C * chunk;
    if (off >= g.total) return 0;
    long long n = g.total - off;
    if (n > g.C + g.ov) n = g.C + g.ov;
    return n >= g.m ? n - g.m + 1 : 0;
}

__device__ __forceinline__ float warp_min(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// grid (ceil(tiles_stride / 8), chunks in segment, snippets), 256 threads: one 1024-sample tile per warp,
// eight independent float4 loads in flight per lane
__global__ void __launch_bounds__(256)
k_tile_minmax(const float *__restrict__ c, ChunkGeom g, float *__restrict__ tmin, float *__restrict__ tmax) {
    const long long chunk = g.first_chunk + blockIdx.y;
    const long long V = chunk_valid_len(g, chunk);
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long k0 = tile << TP_LOG2;
    if (k0 >= V) return;
    const float *y = c + blockIdx.z * g.c_stride + (g.C * chunk - g.c_g0) + k0;
    const long long left = V - k0;
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    if (left >= TP && (((size_t)y) & 15) == 0) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg((const float4 *)y + i * 32 + lane);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            mn = fminf(mn, fminf(fminf(v[i].x, v[i].y), fminf(v[i].z, v[i].w)));
            mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
        }
    } else {
        for (int k = lane; k < TP && k < left; k += 32) {
            float v = __ldg(y + k);
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) {
        size_t o = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * g.tiles_stride + tile;
        tmin[o] = mn;
        tmax[o] = mx;
    }
}


struct PeakOut {
    DevPeak *peaks;                 // global output list
    unsigned long long cap;
    unsigned long long *count;      // appended entries (may exceed cap: overflow is detected by the host)
    unsigned *flags;                // FLAG_OVERFLOW | FLAG_NEED_DENSE (some chunk is marked in redo[])
    // summary mode: redo[snippet * redo_stride + chunk - redo_first] = 1 marks a chunk whose peaks the run records
    // cannot give exactly; it emits nothing and the host repeats exactly those chunks on a dense correlation
    unsigned char *redo;
    long long redo_first, redo_stride;
};
__device__ __forceinline__ long long redo_index(const PeakOut &o, unsigned snippet, long long chunk) {
    return (long long)snippet * o.redo_stride + (chunk - o.redo_first);
}

// ---- summary mode -------------------------------------------------------------------------
// k_col_inv's summary epilogue leaves, for every aligned run of 16 outputs, a record {min, max, first, last}
// and the 16 values themselves only where max >= theta.  A chunk's outputs start run-aligned (C % 16 == 0);
// its last run may be partial: if the chunk ends where the segment ends the record already covers only the
// valid outputs, otherwise exactly one output may be valid (ov == m: V = C + 1) and `first` is that output.
// Any other geometry sets FLAG_NEED_DENSE and the host repeats the call with the dense correlation.
constexpr unsigned FLAG_OVERFLOW = 1u, FLAG_NEED_DENSE = 2u;

// One 16-byte record per run: rec[r] = {min, max, first, last} of the outputs [16 r, 16 r + 16) of the segment
// buffer.  (Measured on B200: splitting it into {min, max} / {first, last} arrays so that the tile pass reads 8
// bytes per run costs the inverse column kernel two scattered 8-byte stores per run instead of one 16-byte store:
// k_col_inv 5.7 -> 6.8 ms per 24 h, tile pass 0.70 -> 0.82 ms.  One array it is.)
struct RunRecs {
    float4 *rec;
    __host__ __device__ RunRecs offset(long long runs) const { return RunRecs{rec + runs}; }
};

struct RunView {
    const float4 *rec;      // records of this chunk, [0] = run of the chunk's first output
    long long nr_full;      // full runs
    int pv;                 // valid outputs of run nr_full (0: there is no partial run)
    bool masked;            // the partial run's record covers only the valid outputs
    __device__ __forceinline__ long long total() const { return nr_full + (pv ? 1 : 0); }
    __device__ __forceinline__ int count(long long r) const { return r == nr_full ? pv : 16; }
    __device__ __forceinline__ void minmax(long long r, float &mn, float &mx) const {
        const float4 q = __ldg(rec + r);
        if (r == nr_full && !masked) { mn = q.z; mx = q.z; } else { mn = q.x; mx = q.y; }
    }
    __device__ __forceinline__ float first(long long r) const { return __ldg(rec + r).z; }
    __device__ __forceinline__ float last(long long r) const { return __ldg(rec + r).w; }
};
__device__ __forceinline__ RunView make_run_view(const RunRecs &rsum, const ChunkGeom &g, long long chunk, long long V,
                                                 unsigned snippet) {
    RunView rv;
    const long long cs = g.C * chunk - g.c_g0;
    rv.rec = rsum.rec + ((snippet * g.c_stride + cs) >> 4);
    rv.nr_full = V >> 4;
    rv.pv = (int)(V & 15);
    rv.masked = (cs + V == g.seg_end);
    return rv;
}

// grid (ceil(tiles_stride / 8), chunks in segment, snippets), 256 threads: one tile (64 runs) per warp
__global__ void __launch_bounds__(256)
k_tile_from_runs(RunRecs rsum, ChunkGeom g, float *__restrict__ tmin, float *__restrict__ tmax,
                 PeakOut out) {
    const long long chunk = g.first_chunk + blockIdx.y;
    const long long V = chunk_valid_len(g, chunk);
    const int lane = threadIdx.x & 31;
    const long long tile = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if ((tile << TP_LOG2) >= V) return;
    const RunView rv = make_run_view(rsum, g, chunk, V, blockIdx.z);
    if (rv.pv > 1 && !rv.masked && lane == 0 && tile == 0) {     // a partial last run the records cannot represent
        out.redo[redo_index(out, blockIdx.z, chunk)] = 1;
        atomicOr(out.flags, FLAG_NEED_DENSE);
    }
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const long long r = (tile << 6) + i * 32 + lane;
        if (r < rv.total()) {
            float a, b;
            rv.minmax(r, a, b);
            mn = fminf(mn, a);
            mx = fmaxf(mx, b);
        }
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) {
        size_t o = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * g.tiles_stride + tile;
        tmin[o] = mn;
        tmax[o] = mx;
    }
}

// walk_min for summary mode: tiles -> runs -> the one stored run that holds the first sample > h
template <int DIR>
__device__ float walk_min_sum(const float *__restrict__ y, const RunView &rv, long long V, const float *__restrict__ tmin,
                              const float *__restrict__ tmax, long long from, float h) {
    const int lane = threadIdx.x & 31;
    float m = h;
    bool found = false;
    // samples [16 r + s_lo, 16 r + s_hi) of a stored run, towards DIR
    auto raw_run = [&](long long r, int s_lo, int s_hi) {
        const int s = (DIR < 0) ? (s_hi - 1 - lane) : (s_lo + lane);
        const bool valid = lane < (s_hi - s_lo);
        const float v = valid ? __ldg(y + (r << 4) + s) : CUDART_INF_F;
        const unsigned hi = __ballot_sync(0xffffffffu, valid && v > h);
        if (hi) {
            const int first = __ffs(hi) - 1;
            m = fminf(m, warp_min(lane < first ? v : CUDART_INF_F));
            found = true;
        } else m = fminf(m, warp_min(v));
    };
    // runs [ra, rb) towards DIR, 32 per step
    auto runs = [&](long long ra, long long rb) {
        long long done = 0;
        const long long len = rb - ra;
        while (done < len && !found) {
            const long long r = (DIR < 0) ? (rb - 1 - done - lane) : (ra + done + lane);
            const bool valid = (DIR < 0) ? (r >= ra) : (r < rb);
            float mn = CUDART_INF_F, mx = -CUDART_INF_F;
            if (valid) rv.minmax(r, mn, mx);
            const unsigned hi = __ballot_sync(0xffffffffu, valid && mx > h);
            if (hi) {
                const int first = __ffs(hi) - 1;
                m = fminf(m, warp_min(lane < first ? mn : CUDART_INF_F));
                const long long rr = (DIR < 0) ? (rb - 1 - done - first) : (ra + done + first);
                raw_run(rr, 0, rv.count(rr));               // max > h >= theta: this run is stored
                found = true;
            } else {
                m = fminf(m, warp_min(mn));
                done += 32;
            }
        }
    };
    const long long total_runs = rv.total();
    const long long ntiles = (V + TP - 1) >> TP_LOG2;
    if (DIR < 0) {
        const long long r0 = from >> 4;
        const int s0 = (int)(from & 15);
        if (s0) raw_run(r0, 0, s0);                         // rest of the peak's own (stored) run
        if (!found) runs((r0 >> 6) << 6, r0);
        long long t_hi = r0 >> 6;
        while (!found && t_hi > 0) {
            const long long tt = t_hi - 1 - lane;
            const bool valid = tt >= 0;
            const float mx = valid ? tmax[tt] : -CUDART_INF_F, mn = valid ? tmin[tt] : CUDART_INF_F;
            const unsigned hi = __ballot_sync(0xffffffffu, valid && mx > h);
            if (hi) {
                const int first = __ffs(hi) - 1;
                m = fminf(m, warp_min(lane < first ? mn : CUDART_INF_F));
                const long long te = t_hi - 1 - first;
                runs(te << 6, (te + 1) << 6);
                found = true;
            } else {
                m = fminf(m, warp_min(mn));
                t_hi -= 32;
            }
        }
    } else {
        long long rn = from >> 4;                           // next run to examine through its record
        const int s0 = (int)(from & 15);
        if (s0) {                                           // rest of the run that holds the plateau's last sample
            const int cnt = rv.count(rn);
            if (s0 < cnt) raw_run(rn, s0, cnt);
            ++rn;
        }
        const long long t = from >> TP_LOG2;                // tile of `from`
        long long tile_hi = (t + 1) << 6;
        if (tile_hi > total_runs) tile_hi = total_runs;
        if (!found && rn < tile_hi) runs(rn, tile_hi);
        long long t_lo = t + 1;
        while (!found && t_lo < ntiles) {
            const long long tt = t_lo + lane;
            const bool valid = tt < ntiles;
            const float mx = valid ? tmax[tt] : -CUDART_INF_F, mn = valid ? tmin[tt] : CUDART_INF_F;
            const unsigned hi = __ballot_sync(0xffffffffu, valid && mx > h);
            if (hi) {
                const int first = __ffs(hi) - 1;
                m = fminf(m, warp_min(lane < first ? mn : CUDART_INF_F));
                const long long te = t_lo + first;
                long long e = (te + 1) << 6;
                if (e > total_runs) e = total_runs;
                runs(te << 6, e);
                found = true;
            } else {
                m = fminf(m, warp_min(mn));
                t_lo += 32;
            }
        }
    }
    return m;
}

// Walk from the peak towards lower (DIR = -1) or higher (DIR = +1) indices until a sample
// strictly greater than h; returns the minimum of the samples passed (h if none).  The whole
// warp calls this with identical arguments.  [lo, hi) is the range still to scan.
template <int DIR>
__device__ float walk_min(const float *__restrict__ y, long long V, const float *__restrict__ tmin,
                          const float *__restrict__ tmax, long long from, float h) {
    const int lane = threadIdx.x & 31;
    float m = h;
    bool found = false;
    // raw scan of [a, b): towards DIR, 32 samples per step; stops at the first sample > h
    auto raw = [&](long long a, long long b) {
        long long done = 0, len = b - a;
        while (done < len && !found) {
            long long k = (DIR < 0) ? (b - 1 - done - lane) : (a + done + lane);
            bool valid = (DIR < 0) ? (k >= a) : (k < b);
            float v = valid ? __ldg(y + k) : CUDART_INF_F;
            unsigned hi = __ballot_sync(0xffffffffu, valid && v > h);
            if (hi) {
                int first = __ffs(hi) - 1;
                m = fminf(m, warp_min(lane < first ? v : CUDART_INF_F));
                found = true;
            } else {
                m = fminf(m, warp_min(v));
                done += 32;
            }
        }
    };
    const long long ntiles = (V + TP - 1) >> TP_LOG2;
    if (DIR < 0) {
        // samples [0, from) remain; first the part inside from's own tile
        long long t = from >> TP_LOG2;                 // tile containing index from (or == ntiles boundary)
        long long lo = t << TP_LOG2;
        if (from > lo) raw(lo, from);
        long long t_hi = t;                            // tiles [0, t_hi) remain
        while (!found && t_hi > 0) {
            long long tt = t_hi - 1 - lane;
            bool valid = tt >= 0;
            float mx = valid ? tmax[tt] : -CUDART_INF_F, mn = valid ? tmin[tt] : CUDART_INF_F;
            unsigned hi = __ballot_sync(0xffffffffu, valid && mx > h);
            if (hi) {
                int first = __ffs(hi) - 1;
                m = fminf(m, warp_min(lane < first ? mn : CUDART_INF_F));
                long long te = t_hi - 1 - first;
                raw(te << TP_LOG2, (te + 1) << TP_LOG2);
            } else {
                m = fminf(m, warp_min(mn));
                t_hi -= 32;
            }
        }
    } else {
        // samples [from, V) remain
        long long t = from >> TP_LOG2;
        long long hi_end = (t + 1) << TP_LOG2;
        if (hi_end > V) hi_end = V;
        if (from < hi_end) raw(from, hi_end);
        long long t_lo = t + 1;                        // tiles [t_lo, ntiles) remain
        while (!found && t_lo < ntiles) {
            long long tt = t_lo + lane;
            bool valid = tt < ntiles;
            float mx = valid ? tmax[tt] : -CUDART_INF_F, mn = valid ? tmin[tt] : CUDART_INF_F;
            unsigned hi = __ballot_sync(0xffffffffu, valid && mx > h);
            if (hi) {
                int first = __ffs(hi) - 1;
                m = fminf(m, warp_min(lane < first ? mn : CUDART_INF_F));
                long long te = t_lo + first;
                long long e = (te + 1) << TP_LOG2;
                if (e > V) e = V;
                raw(te << TP_LOG2, e);
            } else {
                m = fminf(m, warp_min(mn));
                t_lo += 32;
            }
        }
    }
    return m;
}

// dynamic shared memory: [2 * sm_tiles floats: the chunk's tile summaries, when they fit] +
// pk_cap * (2*u32 + 4*f32 + u8) candidates of the current band + pk_cap * (2*u32 + 4*f32) kept peaks.
// grid = (chunks in segment, snippets), 256 threads
// SUM: summary mode -- `rsum` holds the run records, only runs with max >= theta are present in c.
// only_flagged: dense repeat -- work only on the chunks marked in out.redo[].
constexpr int MAX_BANDS = 64;
__host__ __device__ inline size_t chunk_peaks_smem(int pk_cap, int sm_tiles) {
    return (size_t)sm_tiles * 8 + (size_t)pk_cap * (2 * 4 + 4 * 4) * 2 + (size_t)pk_cap + 16;
}
template <bool SUM>
__global__ void __launch_bounds__(256)
k_chunk_peaks(const float *__restrict__ c, RunRecs rsum, float theta, ChunkGeom g,
              const float *__restrict__ tmin_all, const float *__restrict__ tmax_all, float min_prom,
              unsigned long long min_dist, int pk_cap, int sm_tiles, PeakOut out, int only_flagged) {
    extern __shared__ unsigned char smraw[];
    float *s_tmin = (float *)smraw, *s_tmax = s_tmin + sm_tiles;
    unsigned *p_start = (unsigned *)(s_tmax + sm_tiles);
    unsigned *p_end = p_start + pk_cap;
    float *p_h = (float *)(p_end + pk_cap);
    float *p_prom = p_h + pk_cap;
    float *p_ld = p_prom + pk_cap;
    float *p_rd = p_ld + pk_cap;
    unsigned *k_start = (unsigned *)(p_rd + pk_cap);                // kept peaks (emitted at the end)
    unsigned *k_end = k_start + pk_cap;
    float *k_h = (float *)(k_end + pk_cap);
    float *k_prom = k_h + pk_cap;
    float *k_ld = k_prom + pk_cap;
    float *k_rd = k_ld + pk_cap;
    unsigned char *p_alive = (unsigned char *)(k_rd + pk_cap);
    __shared__ float s_red[8], s_red2[8];
    __shared__ int s_redi[8];
    __shared__ int s_ncand, s_win, s_nkept;
    __shared__ unsigned long long s_base;
    __shared__ float s_cmin, s_gmax;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long chunk = g.first_chunk + blockIdx.x;
    const long long V = chunk_valid_len(g, chunk);
    if (V < 3) return;                                              // endpoints are never peaks
    const unsigned snippet_id = blockIdx.y;
    const long long ridx = out.redo ? redo_index(out, snippet_id, chunk) : 0;
    if (only_flagged ? !out.redo[ridx] : (SUM && out.redo[ridx])) return;   // dense repeat: marked chunks only; summary: skip chunks k_tile_from_runs rejected
    const float *y = c + snippet_id * g.c_stride + (g.C * chunk - g.c_g0);
    const float *tmin = tmin_all + ((size_t)snippet_id * gridDim.x + blockIdx.x) * g.tiles_stride;
    const float *tmax = tmax_all + ((size_t)snippet_id * gridDim.x + blockIdx.x) * g.tiles_stride;
    const long long ntiles = (V + TP - 1) >> TP_LOG2;
    if (ntiles <= sm_tiles) {                                       // stage the summaries: the walks re-read them
        for (int t = tid; t < ntiles; t += 256) { s_tmin[t] = tmin[t]; s_tmax[t] = tmax[t]; }
        __syncthreads();
        tmin = s_tmin;
        tmax = s_tmax;
    }

    // (a) chunk minimum and maximum
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    for (long long t = tid; t < ntiles; t += 256) { mn = fminf(mn, tmin[t]); mx = fmaxf(mx, tmax[t]); }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) { s_red[warp] = mn; s_red2[warp] = mx; }
    if (tid == 0) { s_ncand = 0; s_nkept = 0; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) { mn = fminf(mn, s_red[w]); mx = fmaxf(mx, s_red2[w]); }
        s_cmin = mn;
        s_gmax = mx;
    }
    __syncthreads();
    const float cmin = s_cmin, gmax = s_gmax;
    if (!(gmax - cmin >= min_prom)) return;                         // no sample can reach the prominence: no peaks
    // tiles that can hold a candidate: many of them means loud / coloured material where almost every local maximum
    // passes the bound -- the descent then starts with a thin band under the chunk maximum instead of everything
    int nq = 0;
    for (long long t = tid; t < ntiles; t += 256) nq += (tmax[t] - cmin >= min_prom) ? 1 : 0;
    const int many_tiles = __syncthreads_count(nq > 0 ? 1 : 0) > 16;   // (threads that saw a qualifying tile: > 16 <=> > 16 tiles)
    auto mark_dense = [&]() {                                       // this chunk is repeated on a dense correlation; emits nothing here
        if (tid == 0) { out.redo[ridx] = 1; atomicOr(out.flags, FLAG_NEED_DENSE); }
    };
    RunView rv;
    // Summary mode is exact while every run that can hold a candidate (max - cmin >= min_prom) or stop a walk
    // (max > h) is stored, i.e. theta <= min_prom + cmin.  Otherwise (stream much louder than the snippet, tonal
    // material) only candidates of height >= theta can be examined: the result is still exact if the peaks kept
    // among those cover the whole chunk under the minimum distance -- anything lower is then suppressed whatever
    // its prominence.  If they do not, the chunk is marked for the dense repeat.
    bool may_abort = false;
    if constexpr (SUM) {
        rv = make_run_view(rsum, g, chunk, V, snippet_id);
        if (!(min_prom + cmin >= theta)) {
            may_abort = true;
            if (min_dist == 0 || gmax < theta) { mark_dense(); return; }
        }
    }
    const float lo_floor = may_abort ? theta : -CUDART_INF_F;       // lowest height the descent has to reach
    const float b0 = may_abort ? theta : cmin;                      // finite stand-in for the bisection

    // (b) candidates of the band [lo, hi): local maxima whose upper bound h - chunk_min on the prominence passes.
    // prominence = h - max(lmin, rmin) <= h - chunk_min (fp subtraction is monotone), so no peak that find_peaks
    // would keep is dropped here.  Stops early once the list has overflowed (the caller then narrows the band).
    auto enumerate = [&](float lo, float hi) {
        for (long long t = warp; t < ntiles; t += 8) {
            if (*(volatile int *)&s_ncand > pk_cap) break;
            const float tm = tmax[t];
            if (!(tm - cmin >= min_prom) || !(tm >= lo)) continue;  // warp-uniform
            const long long k0 = t << TP_LOG2;
            if constexpr (SUM) {
                // qualifying runs of the tile (two per lane); a lane scans its run serially
                auto val = [&](long long a) { return (a & 15) == 0 ? rv.first(a >> 4) : __ldg(y + a); };
                for (int i = 0; i < 2; ++i) {
                    const long long r = (t << 6) + i * 32 + lane;
                    if (r >= rv.total()) continue;
                    float rmn, rmx;
                    rv.minmax(r, rmn, rmx);
                    if (!(rmx - cmin >= min_prom) || !(rmx >= lo)) continue;
                    const int cnt = rv.count(r);
                    for (int sidx = 0; sidx < cnt; ++sidx) {
                        const long long k = (r << 4) + sidx;
                        if (k < 1 || k >= V - 1) continue;
                        const float yk = __ldg(y + k);
                        if (!(yk >= lo) || !(yk < hi) || !(yk - cmin >= min_prom)) continue;
                        const float prev = sidx ? __ldg(y + k - 1) : rv.last(r - 1);
                        if (!(prev < yk)) continue;
                        long long a = k + 1;
                        while (a < V - 1 && val(a) == yk) ++a;      // plateau (its runs have max >= yk >= theta)
                        if (val(a) < yk) {
                            int slot = atomicAdd(&s_ncand, 1);
                            if (slot < pk_cap) {
                                p_start[slot] = (unsigned)k;
                                p_end[slot] = (unsigned)a;
                                p_h[slot] = yk;
                            }
                        }
                    }
                }
                continue;
            }
            for (int it = 0; it < TP / 32; ++it) {
                long long k = k0 + it * 32 + lane;
                if (k >= 1 && k < V - 1) {
                    float yk = __ldg(y + k);
                    if (yk >= lo && yk < hi && __ldg(y + k - 1) < yk && yk - cmin >= min_prom) {
                        long long a = k + 1;
                        while (a < V - 1 && __ldg(y + a) == yk) ++a;    // plateau
                        if (__ldg(y + a) < yk) {
                            int slot = atomicAdd(&s_ncand, 1);
                            if (slot < pk_cap) {
                                p_start[slot] = (unsigned)k;
                                p_end[slot] = (unsigned)a;
                                p_h[slot] = yk;
                            }
                        }
                    }
                }
            }
        }
    };
    auto mid_of = [](unsigned a, unsigned b) { return ((unsigned long long)a + b) / 2; };
    // do the kept peaks leave no position of [1, V-2] at distance >= min_dist from all of them?
    auto covered = [&]() -> bool {
        const int nk = s_nkept;
        bool ok = true;
        for (int i = tid; i < nk; i += 256) {
            const unsigned long long mi = mid_of(k_start[i], k_end[i]);
            unsigned long long next = ~0ull;
            bool has_prev = false;
            for (int j = 0; j < nk; ++j) {
                const unsigned long long mj = mid_of(k_start[j], k_end[j]);
                if (mj > mi && mj < next) next = mj;
                if (mj < mi) has_prev = true;
            }
            if (!has_prev && mi - 1 >= min_dist && mi >= 1 + min_dist) ok = false;            // position 1 is out of reach
            if (next == ~0ull) { if ((unsigned long long)(V - 2) >= mi + min_dist) ok = false; }
            else if (next - mi >= 2 * min_dist) ok = false;
        }
        return __syncthreads_and(ok) && nk > 0;
    };

    float hi = CUDART_INF_F, width = 0.f;
    bool finished = false;
    for (int band = 0; band < MAX_BANDS && !finished; ++band) {
        // choose the band's lower edge: the floor if the candidates fit, else bisect towards the band's top
        const float top = (hi < CUDART_INF_F) ? hi : gmax;
        float lo = lo_floor;
        if (band > 0 && width > 0.f && hi - 2.f * width > b0) lo = hi - 2.f * width;
        if (band == 0 && many_tiles && min_dist > 0) {
            const float thin = gmax - (gmax - fmaxf(b0, cmin + min_prom)) * (1.0f / 64.0f);
            if (thin > b0 && thin < gmax) lo = thin;
        }
        int ncand = 0;
        for (;;) {
            __syncthreads();
            if (tid == 0) s_ncand = 0;
            __syncthreads();
            enumerate(lo, hi);
            __syncthreads();
            ncand = s_ncand;
            if (ncand <= pk_cap) break;
            const float from = (lo > -CUDART_INF_F) ? lo : b0;
            const float nlo = from + 0.5f * (top - from);
            if (!(nlo > from) || !(nlo <= top) || nlo == lo) {      // cannot narrow further: > pk_cap equal-height maxima
                if (tid == 0) atomicOr(out.flags, FLAG_OVERFLOW);
                return;
            }
            lo = nlo;
        }
        const bool reached_floor = !(lo > lo_floor);

        // (c) exact prominence, one warp per candidate
        for (int i = warp; i < ncand; i += 8) {
            const float h = p_h[i];
            const long long s = p_start[i], e = p_end[i];
            float lmin, rmin;
            if constexpr (SUM) {
                lmin = walk_min_sum<-1>(y, rv, V, tmin, tmax, s, h);
                rmin = walk_min_sum<+1>(y, rv, V, tmin, tmax, e, h);
            } else {
                lmin = walk_min<-1>(y, V, tmin, tmax, s, h);
                rmin = walk_min<+1>(y, V, tmin, tmax, e, h);
            }
            if (lane == 0) {
                float prom = h - fmaxf(lmin, rmin);
                p_prom[i] = prom;
                if constexpr (SUM) {
                    p_ld[i] = h - ((s & 15) ? __ldg(y + s - 1) : rv.last((s >> 4) - 1));
                    p_rd[i] = h - ((e & 15) ? __ldg(y + e) : rv.first(e >> 4));
                } else {
                    p_ld[i] = h - __ldg(y + s - 1);
                    p_rd[i] = h - __ldg(y + e);
                }
                p_alive[i] = prom >= min_prom;                      // with_min_prominence, :227
            }
        }
        __syncthreads();

        auto keep = [&](int i) {                                    // one thread; returns false when the kept list is full
            const int k = s_nkept;
            if (k >= pk_cap) return false;
            k_start[k] = p_start[i]; k_end[k] = p_end[i]; k_h[k] = p_h[i];
            k_prom[k] = p_prom[i]; k_ld[k] = p_ld[i]; k_rd[k] = p_rd[i];
            s_nkept = k + 1;
            return true;
        };
        // (d) with_min_distance (:228): greedy in descending height, ties by lower position.  Peaks kept in
        // earlier (higher) bands come first in that order, so they suppress this band's candidates up front.
        if (min_dist == 0) {
            // no suppression: every alive candidate is a result.  (may_abort chunks never get here.)  Flush the
            // kept list to the output whenever it fills up.
            for (int i0 = 0; i0 < ncand; i0 += 256) {
                const int i = i0 + tid;
                const bool a = i < ncand && p_alive[i];
                const unsigned long long slot = a ? atomicAdd(out.count, 1ull) : 0;
                if (a && slot < out.cap) {
                    DevPeak p;
                    const unsigned long long off = (unsigned long long)(g.C * chunk);
                    p.start = off + p_start[i]; p.end = off + p_end[i]; p.height = p_h[i]; p.prominence = p_prom[i];
                    p.left_diff = p_ld[i]; p.right_diff = p_rd[i]; p.snippet_id = snippet_id; p.chunk = (unsigned)chunk;
                    out.peaks[slot] = p;
                }
            }
        } else {
            const int nk0 = s_nkept;
            for (int i = tid; i < ncand; i += 256) {
                if (!p_alive[i]) continue;
                const unsigned long long mi = mid_of(p_start[i], p_end[i]);
                for (int k = 0; k < nk0; ++k) {
                    const unsigned long long mk = mid_of(k_start[k], k_end[k]);
                    if ((mi > mk ? mi - mk : mk - mi) < min_dist) { p_alive[i] = 0; break; }
                }
            }
            __syncthreads();
            for (;;) {
                float bh = -CUDART_INF_F;
                int bi = -1;
                for (int i = tid; i < ncand; i += 256) {
                    if (p_alive[i]) {
                        float h = p_h[i];
                        if (bi < 0 || h > bh || (h == bh && p_start[i] < p_start[bi])) { bh = h; bi = i; }
                    }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    float oh = __shfl_xor_sync(0xffffffffu, bh, o);
                    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (oi >= 0 && (bi < 0 || oh > bh || (oh == bh && p_start[oi] < p_start[bi]))) { bh = oh; bi = oi; }
                }
                if (lane == 0) { s_red[warp] = bh; s_redi[warp] = bi; }
                __syncthreads();
                if (tid == 0) {
                    for (int w = 1; w < 8; ++w) {
                        float oh = s_red[w];
                        int oi = s_redi[w];
                        if (oi >= 0 && (bi < 0 || oh > bh || (oh == bh && p_start[oi] < p_start[bi]))) { bh = oh; bi = oi; }
                    }
                    if (bi >= 0 && !keep(bi)) bi = -2;              // kept list full
                    s_win = bi;
                }
                __syncthreads();
                const int win = s_win;
                if (win == -2) {
                    if (tid == 0) atomicOr(out.flags, FLAG_OVERFLOW);
                    return;
                }
                if (win < 0) break;
                const unsigned long long mw = mid_of(p_start[win], p_end[win]);
                for (int i = tid; i < ncand; i += 256) {
                    if (p_alive[i]) {
                        unsigned long long mi = mid_of(p_start[i], p_end[i]);
                        unsigned long long d = mi > mw ? mi - mw : mw - mi;
                        if (i == win || d < min_dist) p_alive[i] = 0;
                    }
                }
                __syncthreads();
            }
        }

        // done when the descent has reached its floor, or when nothing lower can survive the minimum distance
        if (reached_floor) {
            if (may_abort && !covered()) { mark_dense(); return; }  // peaks below theta may exist: not decidable from the records
            finished = true;
        } else if (min_dist > 0 && covered()) {
            finished = true;
        } else {
            width = top - lo;
            hi = lo;
        }
    }
    if (!finished) {                                                // too many bands: the kept peaks never covered the chunk
        if (tid == 0) atomicOr(out.flags, FLAG_OVERFLOW);
        return;
    }

    // emit the kept peaks: one reservation per chunk
    const int nk = s_nkept;
    if (nk == 0) return;
    if (tid == 0) s_base = atomicAdd(out.count, (unsigned long long)nk);
    __syncthreads();
    const unsigned long long base = s_base;
    const unsigned long long off = (unsigned long long)(g.C * chunk);   // offset_range, lib.rs:8-10
    for (int i = tid; i < nk; i += 256) {
        if (base + i < out.cap) {
            DevPeak p;
            p.start = off + k_start[i];
            p.end = off + k_end[i];
            p.height = k_h[i];
            p.prominence = k_prom[i];
            p.left_diff = k_ld[i];
            p.right_diff = k_rd[i];
            p.snippet_id = snippet_id;
            p.chunk = (unsigned)chunk;
            out.peaks[base + i] = p;
        }
    }
}

// Test hook: build the run records the summary epilogue of k_col_inv would have written for a correlation that is
// already in memory, and poison (NaN) every run it would not have stored, so that any read the summary-mode
// kernels are not entitled to shows up as a wrong result.
__global__ void k_debug_make_runs(float *__restrict__ c, long long n, float theta, RunRecs rsum) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if ((r << 4) >= n) return;
    const long long left = n - (r << 4);
    const int valid = left < 16 ? (int)left : 16;
    float *p = c + (r << 4);
    float mn = p[0], mx = p[0], last = p[0];
    for (int i = 1; i < valid; ++i) { mn = fminf(mn, p[i]); mx = fmaxf(mx, p[i]); last = p[i]; }
    rsum.rec[r] = make_float4(mn, mx, p[0], last);
    if (!(mx >= theta))
        for (int i = 0; i < valid; ++i) p[i] = __int_as_float(0x7fc00000);
}

}  // namespace amp
