// am_capi.cu -- C ABI (include/audio_matcher.h) and host-side scheduling of the matcher.
//
// Host responsibilities (everything numeric runs in the CUDA kernels):
//   * plan: overlap-save block length N, four-step split N1 x N2, segments of logical chunks
//   * stage host-resident PCM to the device through two buffers on a copy stream
//   * launch the column / row / column kernels (or the single-pass kernel) per group of block pairs
//     (TMA descriptors for the persistent forward column kernel are encoded here per launch), then
//     the per-chunk peak kernels; one D2H of the peak list at the end
//   * calc_chunks' tail: stable sort by start + filter_surrounding/is_overshadowed
//     (src/matcher/audio_matcher.rs:132-160) over the handful of surviving peaks
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <chrono>
#include <condition_variable>

#include <cuda.h>
#include <dlfcn.h>

#include "../../include/audio_matcher.h"
#include "am_kernels.cuh"
#include "am_peaks.cuh"

static_assert(sizeof(am_peak) == sizeof(amp::DevPeak), "am_peak layout");

namespace {

thread_local char g_err[512] = "";

am_status fail(am_status st, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return st;
}

#define CU(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? AM_ERR_NOMEM : AM_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
    } while (0)
#define TRY(expr)                       \
    do {                                \
        am_status s_ = (expr);          \
        if (s_ != AM_OK) return s_;     \
    } while (0)

size_t env_mb(const char *name, size_t dflt_mb) {
    const char *v = getenv(name);
    if (v && *v) {
        long long x = atoll(v);
        if (x > 0) return (size_t)x;
    }
    return dflt_mb;
}

int ceil_log2(unsigned long long x) {
    int l = 0;
    while ((1ull << l) < x) ++l;
    return l;
}

template <class T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    am_status reserve(size_t want) {
        if (want <= n) return AM_OK;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        CU(cudaMalloc((void **)&p, want * sizeof(T)));
        n = want;
        return AM_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

}  // namespace

// Upload of PAGEABLE host memory (what a Vec / numpy caller or a decoder thread hands over).  cudaMemcpyAsync stages
// such memory inside the driver at ~11 GB/s; here a persistent pool of host threads copies 8 MB chunks into a small
// ring of pinned buffers and each chunk goes out with ONE asynchronous copy while the next one is being filled.
// Pinned / registered memory handed to the one-shot entry points never comes here.
struct HostStager {
    static constexpr size_t CHUNK = (size_t)32 << 20;
    static constexpr int SLOTS = 4;
    void *pinned[SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    bool busy[SLOTS] = {false, false, false, false};
    int slot = 0;                            // next slot to fill
    // worker pool: one job at a time, cut into pieces that the workers and the calling thread pull
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    char *job_dst = nullptr;
    const char *job_src = nullptr;
    size_t job_n = 0, job_piece = 0, job_pieces = 0, next_piece = 0, pieces_done = 0;
    unsigned long long generation = 0;
    std::atomic<unsigned long long> generation_pub{0};   // copy of `generation` the workers may poll without the lock
    bool stop = false;

    // bytes per slot fill / asynchronous copy: 8 MB fills run at ~71 GB/s, 32 MB fills at ~45 (the ring then no longer
    // stays cache-resident); AM_STAGE_CHUNK_MB overrides (1..32)
    static size_t fill_bytes() {
        static const size_t piece = [] { const char *v = getenv("AM_STAGE_CHUNK_MB"); size_t mb = v && *v ? (size_t)atoi(v) : 8; return std::min<size_t>(std::max<size_t>(mb, 1), 32) << 20; }();
        return piece;
    }
    static int default_threads() {
        static const int env = [] { const char *v = getenv("AM_STAGE_THREADS"); return v && *v ? atoi(v) : -1; }();
        if (env >= 0) return env;
        const int hw = (int)std::thread::hardware_concurrency();
        return std::max(1, std::min(12, hw / 2));
    }
    bool ensure(int threads) {
        for (int i = 0; i < SLOTS; ++i) {
            if (!pinned[i] && cudaHostAlloc(&pinned[i], CHUNK, cudaHostAllocDefault) != cudaSuccess) { pinned[i] = nullptr; cudaGetLastError(); return false; }
            if (!ev[i] && cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) { ev[i] = nullptr; cudaGetLastError(); return false; }
        }
        while ((int)workers.size() + 1 < threads) workers.emplace_back([this] { work(); });
        return true;
    }
    bool take_piece(size_t &i) {            // mu held
        if (next_piece >= job_pieces) return false;
        i = next_piece++;
        return true;
    }
    void copy_piece(size_t i) {
        const size_t o = i * job_piece;
        memcpy(job_dst + o, job_src + o, std::min(job_piece, job_n - o));
    }
    void work() {
        std::unique_lock<std::mutex> lk(mu);
        unsigned long long seen = 0;
        for (;;) {
            if (!stop && generation == seen) {
                // a decoder pushes block after block: poll for the next job for a moment before going to sleep
                // (a condition-variable wake-up costs about as much as copying a 4 MB push)
                lk.unlock();
                const auto t0 = std::chrono::steady_clock::now();
                while (generation_pub.load(std::memory_order_acquire) == seen &&
                       std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(150)) {
#if defined(__x86_64__) || defined(__i386__)
                    __builtin_ia32_pause();
#endif
                }
                lk.lock();
            }
            cv_work.wait(lk, [&] { return stop || (generation != seen && next_piece < job_pieces); });
            if (stop) return;
            seen = generation;
            size_t i;
            while (take_piece(i)) {
                lk.unlock();
                copy_piece(i);
                lk.lock();
                if (++pieces_done == job_pieces) cv_done.notify_all();
            }
        }
    }
    void parallel_copy(void *dst, const void *src, size_t n) {
        if (workers.empty() || n < ((size_t)1 << 20)) { memcpy(dst, src, n); return; }
        std::unique_lock<std::mutex> lk(mu);
        job_dst = (char *)dst; job_src = (const char *)src; job_n = n;
        // two pieces per thread, between 128 KB and 1 MB each: a 4 MB push still keeps the whole pool busy
        job_piece = std::min<size_t>((size_t)1 << 20, std::max<size_t>((size_t)128 << 10, (n / (2 * (workers.size() + 1)) + 4095) & ~(size_t)4095));
        job_pieces = (n + job_piece - 1) / job_piece;
        next_piece = 0; pieces_done = 0;
        ++generation;
        generation_pub.store(generation, std::memory_order_release);
        cv_work.notify_all();
        size_t i;
        while (take_piece(i)) {
            lk.unlock();
            copy_piece(i);
            lk.lock();
            ++pieces_done;
        }
        cv_done.wait(lk, [&] { return pieces_done == job_pieces; });
    }
    // wait until the slot's previous asynchronous copy has left it
    cudaError_t acquire(int sl) {
        if (busy[sl]) {
            cudaError_t e = cudaEventSynchronize(ev[sl]);
            if (e != cudaSuccess) return e;
            busy[sl] = false;
        }
        return cudaSuccess;
    }
    cudaError_t send(int sl, void *dst_dev, size_t n, cudaStream_t s) {
        cudaError_t e = cudaMemcpyAsync(dst_dev, pinned[sl], n, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return e;
        if ((e = cudaEventRecord(ev[sl], s)) != cudaSuccess) return e;
        busy[sl] = true;
        return cudaSuccess;
    }
    cudaError_t upload(void *dst_dev, const void *src, size_t bytes, cudaStream_t s) {
        static const bool dbg = getenv("AM_STAGE_DEBUG") != nullptr;
        const size_t piece = fill_bytes();
        double t_acq = 0, t_copy = 0, t_send = 0;
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        for (size_t o = 0; o < bytes; o += piece, slot = (slot + 1) % SLOTS) {
            const size_t n = std::min(piece, bytes - o);
            cudaError_t e;
            const double t0 = dbg ? now() : 0;
            if ((e = acquire(slot)) != cudaSuccess) return e;
            const double t1 = dbg ? now() : 0;
            parallel_copy(pinned[slot], (const char *)src + o, n);
            const double t2 = dbg ? now() : 0;
            if ((e = send(slot, (char *)dst_dev + o, n, s)) != cudaSuccess) return e;
            if (dbg) { const double t3 = now(); t_acq += t1 - t0; t_copy += t2 - t1; t_send += t3 - t2; }
        }
        if (dbg) fprintf(stderr, "[stager] %zu MB: acquire %.2f ms, copy %.2f ms (%.1f GB/s), send %.2f ms\n", bytes >> 20, t_acq, t_copy, bytes / t_copy / 1e6, t_send);
        return cudaSuccess;
    }
    void release() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto &t : workers) t.join();
        workers.clear();
        stop = false;
        for (int i = 0; i < SLOTS; ++i) {
            if (ev[i]) { if (busy[i]) cudaEventSynchronize(ev[i]); cudaEventDestroy(ev[i]); ev[i] = nullptr; }
            if (pinned[i]) { cudaFreeHost(pinned[i]); pinned[i] = nullptr; }
            busy[i] = false;
        }
    }
};

struct am_matcher {
    std::mutex mu;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    am_config cfg;
    uint32_t sr = 0;
    size_t m = 0;
    size_t S = 1;                            // snippets in the batch (all of m samples)
    size_t active = 0;                       // snippet am_correlate / am_inverse_sample_auto_correlation use
    DevBuf<float> d_snip;                    // [S][m]
    std::vector<double> sumsq;
    std::vector<float> inv_ac;
    DevBuf<float2> d_tw;
    DevBuf<int> d_sched;                // next-tile counter of the persistent kernels
    std::map<int, float2 *> spectra;         // log2n -> [S][N] conjugate spectra
    DevBuf<float2> d_A, d_B;
    DevBuf<float> d_c, d_tmin, d_tmax;
    DevBuf<float4> d_rsum;                   // summary mode: one {min, max, first, last} record per aligned run of 16 outputs
    DevBuf<amp::DevPeak> d_peaks;
    DevBuf<unsigned char> d_redo;           // summary mode: chunks marked for the dense repeat, [snippet][chunk of the call]
    DevBuf<unsigned long long> d_count;   // [0] = count, [1] low 32 bits = flags
    DevBuf<unsigned char> d_stage[2];
    HostStager stager;                      // pageable host streams only
    DevBuf<unsigned char> d_gsend, d_grecv; // multi-GPU: fixed-size peak records for the all-gather
    // am_calc_chunks_files: per-file peak lists, counters and dense-repeat marks, so that the work of all files is
    // queued before the first result is read back
    DevBuf<amp::DevPeak> d_fpeaks;
    DevBuf<unsigned long long> d_fcount;
    DevBuf<unsigned char> d_fredo;
    bool in_session = false;                // a push session (am_stream_begin) owns the calc_chunks buffers
    am_progress_fn progress = nullptr;      // optional, fired from the calling thread
    void *progress_user = nullptr;
    am_stats stats;
    // optional per-kernel-class device timing (cudaEvent pairs around every launch)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<int> ev_cls;                 // class of event pair i (events 2i, 2i+1)
    double cls_ms[AM_KERNEL_CLASSES] = {0};
    uint64_t cls_launches[AM_KERNEL_CLASSES] = {0};
};

namespace {

size_t fmt_bytes(int fmt) { return fmt == AM_FMT_F32_MONO ? 4 : (fmt == AM_FMT_I16_MONO ? 2 : 4); }

template <class K> am_status set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return AM_OK;
}

cudaEvent_t prof_event(am_matcher *h) {
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        h->ev_pool.push_back(e);
    }
    return h->ev_pool[h->ev_used++];
}
void prof_begin(am_matcher *h, int cls) {
    if (!h->profiling) return;
    h->ev_cls.push_back(cls);
    cudaEventRecord(prof_event(h), h->stream);
}
void prof_end(am_matcher *h) {
    if (!h->profiling) return;
    cudaEventRecord(prof_event(h), h->stream);
}
// after the stream has been synchronised: fold the recorded event pairs into the class totals
void prof_collect(am_matcher *h) {
    for (size_t i = 0; i < h->ev_cls.size(); ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]) == cudaSuccess) {
            h->cls_ms[h->ev_cls[i]] += ms;
            h->cls_launches[h->ev_cls[i]]++;
        }
    }
    h->ev_cls.clear();
    h->ev_used = 0;
}

#define LAUNCH(h, cls, ...)                    \
    do {                                       \
        prof_begin((h), (cls));                \
        __VA_ARGS__;                           \
        prof_end((h));                         \
        (h)->stats.kernel_launches++;          \
        CU(cudaGetLastError());                \
    } while (0)

// ---- kernel dispatch ----------------------------------------------------------------
template <int LOG2N> am_status launch_small_t(am_matcher *h, const amk::BlockGroup &g, const float2 *spec) {
    typedef amk::SmallCfg<LOG2N> Cfg;
    TRY(set_smem(amk::k_small<LOG2N>, Cfg::SMEM));
    int pairs = (g.nblocks + 1) / 2;
    int grid = (pairs + Cfg::G - 1) / Cfg::G;
    LAUNCH(h, AM_K_SMALL, amk::k_small<LOG2N><<<grid, Cfg::THREADS, Cfg::SMEM, h->stream>>>(g, spec, h->d_tw.p));
    return AM_OK;
}
am_status launch_small(am_matcher *h, int log2n, const amk::BlockGroup &g, const float2 *spec) {
    switch (log2n) {
#define C_(L) case L: return launch_small_t<L>(h, g, spec);
        C_(4) C_(5) C_(6) C_(7) C_(8) C_(9) C_(10) C_(11) C_(12) C_(13)
#undef C_
    }
    return fail(AM_ERR_UNSUPPORTED, "single-pass length 2^%d not built", log2n);
}

// TMA descriptor over a resident mono int16 window: rows of N2 frames at a pitch of N2 frames, 2 N2 columns wide (see
// k_col_fwd_stream).  The driver entry point is looked up once; false = not available (the caller keeps the LDG kernel).
typedef CUresult (*encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_t tensor_map_encoder() {
    static const encode_t encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
            fn = nullptr;
        cudaGetLastError();
        return (encode_t)fn;
    }();
    return encode;
}
static_assert(sizeof(amk::TensorMap) == sizeof(CUtensorMap), "tensor map size");
bool encode_pcm_tensor_map(amk::TensorMap *out, const amk::StreamView &sv, int log2n2, int box_cols, int box_rows, int frame_bytes) {
    const encode_t encode = tensor_map_encoder();
    if (!encode || (((size_t)sv.x) & 15) != 0 || sv.buf_frames < (1ll << log2n2)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)2 << log2n2, (cuuint64_t)(sv.buf_frames >> log2n2) + 1};
    const cuuint64_t strides[1] = {(cuuint64_t)frame_bytes << log2n2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    return encode((CUtensorMap *)out, frame_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2,
                  const_cast<void *>(sv.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// true = launched; false with last_status() == AM_OK = not applicable (the caller falls back to k_col_fwd)
thread_local am_status tl_status = AM_OK;
am_status last_status() { return tl_status; }
template <int L1, int LT, int E, int FMT, int L2C = -1>
bool launch_col_stream(am_matcher *h, const amk::BlockGroup &g, int l2, float2 *A, dim3 grid) {
    if constexpr (L2C < 0 && ((L1 == 9 && LT == 4 && E == 32 && FMT == amk::FMT_I16_MONO) ||
                              (L1 == 10 && LT == 3 && E == 32 && FMT == amk::FMT_I16_MONO) ||
                              (L1 == 7 && LT == 4 && E == 16 && FMT == amk::FMT_I16_STEREO)))
        if (l2 == 13) return launch_col_stream<L1, LT, E, FMT, 13>(h, g, l2, A, grid);
    if constexpr (L2C < 0 && L1 == 9 && LT == 4 && E == 32 && FMT == amk::FMT_I16_MONO)
        if (l2 == 14) return launch_col_stream<L1, LT, E, FMT, 14>(h, g, l2, A, grid);     // N = 2^23 = 512 x 16384
    typedef amk::ColStreamCfg<L1, LT, E, FMT> SC;
    tl_status = AM_OK;
    if constexpr (!SC::OK) return false;
    else {
        amk::TensorMap tm;
        if (!encode_pcm_tensor_map(&tm, g.sv, l2, SC::T, SC::BOX_ROWS, (int)sizeof(typename SC::Frame))) return false;
        static const int ctas = [] {
            const char *v = getenv("AM_COL_STREAM_CTAS");
            if (v && *v) return atoi(v);
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaFuncSetAttribute(amk::k_col_fwd_stream<L1, LT, E, FMT, L2C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC::SMEM);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, amk::k_col_fwd_stream<L1, LT, E, FMT, L2C>, SC::THREADS, SC::SMEM);
            return sms * (per_sm > 0 ? per_sm : 1);
        }();
        const int ntiles = (int)(grid.x * grid.y), nctas = ntiles < ctas ? ntiles : ctas;
        auto body = [&]() -> am_status {
            TRY(set_smem(amk::k_col_fwd_stream<L1, LT, E, FMT, L2C>, SC::SMEM));
            TRY(h->d_sched.reserve(1));
            CU(cudaMemsetAsync(h->d_sched.p, 0, sizeof(int), h->stream));
            LAUNCH(h, AM_K_COL_FWD, amk::k_col_fwd_stream<L1, LT, E, FMT, L2C><<<nctas, SC::THREADS, SC::SMEM, h->stream>>>(tm, g, l2, A, h->d_tw.p, ntiles, h->d_sched.p));
            return AM_OK;
        };
        tl_status = body();
        return tl_status == AM_OK;
    }
}

template <int L1, int LT, int E, bool INV> am_status launch_col_t(am_matcher *h, const amk::BlockGroup &g, int l2, float2 *A) {
    typedef amk::ColCfg<L1, LT, E> Cfg;
    int pairs = (g.nblocks + 1) / 2;
    dim3 grid((1u << l2) >> Cfg::LT, pairs);
    if (INV) {
        if constexpr ((L1 == 9 && LT == 4 && E == 32) || (L1 == 10 && LT == 3 && E == 16) || (L1 == 7 && LT == 4 && E == 16)) {
            if (l2 == 13) {                   // N = 2^22 / 2^23 / 2^20 with 8192-point rows: column pitch known at compile time
                TRY(set_smem(amk::k_col_inv<L1, LT, E, 13>, Cfg::SMEM_INV));
                LAUNCH(h, AM_K_COL_INV, amk::k_col_inv<L1, LT, E, 13><<<grid, Cfg::THREADS, Cfg::SMEM_INV, h->stream>>>(g, l2, A, h->d_tw.p));
                return AM_OK;
            }
        }
        if constexpr (L1 == 9 && LT == 4 && E == 32) {
            if (l2 == 14) {                   // N = 2^23 = 512 x 16384
                TRY(set_smem(amk::k_col_inv<L1, LT, E, 14>, Cfg::SMEM_INV));
                LAUNCH(h, AM_K_COL_INV, amk::k_col_inv<L1, LT, E, 14><<<grid, Cfg::THREADS, Cfg::SMEM_INV, h->stream>>>(g, l2, A, h->d_tw.p));
                return AM_OK;
            }
        }
        TRY(set_smem(amk::k_col_inv<L1, LT, E>, Cfg::SMEM_INV));
        LAUNCH(h, AM_K_COL_INV, amk::k_col_inv<L1, LT, E><<<grid, Cfg::THREADS, Cfg::SMEM_INV, h->stream>>>(g, l2, A, h->d_tw.p));
    } else {
        // resident window in one of the three sample formats: persistent kernel whose next tile is fetched by TMA
        static const bool stream_on = [] { const char *v = getenv("AM_COL_STREAM"); return !(v && *v == '0'); }();
        if (stream_on) {
            switch (g.sv.fmt) {
            case amk::FMT_I16_MONO: if (launch_col_stream<L1, LT, E, amk::FMT_I16_MONO>(h, g, l2, A, grid)) return AM_OK; break;
            case amk::FMT_I16_STEREO: if (launch_col_stream<L1, LT, E, amk::FMT_I16_STEREO>(h, g, l2, A, grid)) return AM_OK; break;
            case amk::FMT_F32_MONO: if (launch_col_stream<L1, LT, E, amk::FMT_F32_MONO>(h, g, l2, A, grid)) return AM_OK; break;
            }
            if (last_status() != AM_OK) return last_status();
        }
        TRY(set_smem(amk::k_col_fwd<L1, LT, E>, Cfg::SMEM));
        LAUNCH(h, AM_K_COL_FWD, amk::k_col_fwd<L1, LT, E><<<grid, Cfg::THREADS, Cfg::SMEM, h->stream>>>(g, l2, A, h->d_tw.p));
    }
    return AM_OK;
}
template <bool INV> am_status launch_col(am_matcher *h, int l1, const amk::BlockGroup &g, int l2, float2 *A) {
    static const int lt_env = [] { const char *v = getenv("AM_COL_LT"); return v && *v ? atoi(v) : 0; }();
    // elements per thread: 16 for the forward tiles (more warps hide the PCM staging), 32 for the inverse
    // tiles (one exchange, 80 registers, 3 CTAs per SM) -- measured; AM_COL_EPT_FWD / AM_COL_EPT_INV override
    static const int ept_fwd = [] { const char *v = getenv("AM_COL_EPT_FWD"); return v && *v ? atoi(v) : 16; }();
    static const int ept_inv = [] { const char *v = getenv("AM_COL_EPT_INV"); return v && *v ? atoi(v) : 32; }();
    static const bool ept_fwd_set = [] { const char *v = getenv("AM_COL_EPT_FWD"); return v && *v; }();
    const int ept = INV ? ept_inv : ept_fwd;
    switch (l1) {
#define C_(L) case L: return launch_col_t<L, amk::col_default_lt(L), 16, INV>(h, g, l2, A);
        C_(4) C_(5) C_(6) C_(7) C_(11)
#undef C_
    case 8:
        if (lt_env == 5) return launch_col_t<8, 5, 16, INV>(h, g, l2, A);
        return launch_col_t<8, 4, 16, INV>(h, g, l2, A);
    case 9:                                   // tuning knobs: 8 or 16 columns per tile, 16 or 32 elements per thread
        // forward tiles take the TMA-fed persistent kernel, which is fastest with 32 elements per thread
        if (!INV && !ept_fwd_set) return launch_col_t<9, 4, 32, INV>(h, g, l2, A);
        if (ept == 32) return launch_col_t<9, 4, 32, INV>(h, g, l2, A);
        if (lt_env == 3) return launch_col_t<9, 3, 16, INV>(h, g, l2, A);
        return launch_col_t<9, 4, 16, INV>(h, g, l2, A);
    case 10:                                  // 1024-point columns: inverse tiles are faster with 16 elements per thread
        if (!INV && !ept_fwd_set) return launch_col_t<10, 3, 32, INV>(h, g, l2, A);     // TMA-fed forward: 8.3 vs 11.4 ms (cfg 4, 30 h)
        if (ept == 32 && getenv("AM_COL10_EPT32")) {
            if (atoi(getenv("AM_COL10_EPT32")) == 3) return launch_col_t<10, 3, 32, INV>(h, g, l2, A);
            return launch_col_t<10, 4, 32, INV>(h, g, l2, A);
        }
        if (lt_env == 4) return launch_col_t<10, 4, 16, INV>(h, g, l2, A);
        return launch_col_t<10, 3, 16, INV>(h, g, l2, A);
    }
    return fail(AM_ERR_UNSUPPORTED, "column length 2^%d not built", l1);
}

// does launch_col<true> pick a 16-column-tile kernel (the ones with the run-summary epilogue) for this column length?
bool col_inv_writes_runs(int l1) {
    const char *v = getenv("AM_COL_LT");
    const int lt_env = v && *v ? atoi(v) : 0;
    v = getenv("AM_COL_EPT_INV");
    const int ept_inv = v && *v ? atoi(v) : 32;
    if (l1 == 7) return true;
    if (l1 == 8) return lt_env != 5;
    if (l1 == 9) return ept_inv == 32 || lt_env != 3;
    return false;
}

// nsn > 1 (ROW_INVERSE only): nsn snippets in one launch of the persistent kernel, see k_row32_stream
template <int L2, int MODE>
am_status launch_row_t(am_matcher *h, float2 *A, const float2 *spec, float2 *B, int l1, int rows, int nsn = 1,
                       size_t spec_stride = 0, size_t b_stride = 0) {
    typedef amk::RowCfg<L2> Cfg;
    static const int ept = [] { const char *v = getenv("AM_ROW_EPT"); return v && *v ? atoi(v) : 32; }();
    if constexpr (L2 == 13 || L2 == 12 || L2 == 14) {
        if (ept == 32) {
            typedef amk::Row32Cfg<L2> C32;
            if constexpr ((MODE == amk::ROW_FUSED || MODE == amk::ROW_INVERSE) && L2 == 13) {
                // persistent kernel fed by bulk asynchronous copies.  Fused: hides the load waits but measured no
                // faster (12.0 vs 11.6 ms per 24 h; the transforms themselves bound the row pass) -- opt-in,
                // AM_ROW_STREAM=1.  Inverse-only (batch mode): the plain kernel is load-latency bound -- default,
                // AM_ROW_STREAM=0 turns it off.
                static const int stream_env = [] { const char *v = getenv("AM_ROW_STREAM"); return v && *v ? atoi(v) : -1; }();
                const bool stream_on = MODE == amk::ROW_FUSED ? stream_env == 1 : stream_env != 0;
                if (stream_on) {
                    static const int ctas = [] {
                        const char *v = getenv("AM_ROW_STREAM_CTAS");
                        if (v && *v) return atoi(v);
                        int dev = 0, sms = 0;
                        cudaGetDevice(&dev);
                        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                        return sms * 2;
                    }();
                    TRY(set_smem(amk::k_row32_stream<L2, MODE>, C32::SMEM));
                    TRY(h->d_sched.reserve(1));
                    CU(cudaMemsetAsync(h->d_sched.p, 0, sizeof(int), h->stream));
                    const long long tickets = (long long)rows * nsn;
                    LAUNCH(h, AM_K_ROW, amk::k_row32_stream<L2, MODE><<<(unsigned)(tickets < ctas ? tickets : ctas), C32::THREADS, C32::SMEM, h->stream>>>(
                                            A, spec, B, l1, rows, h->d_tw.p, h->d_sched.p, nsn, spec_stride, b_stride));
                    return AM_OK;
                }
            }
            if (nsn > 1) return fail(AM_ERR_UNSUPPORTED, "multi-snippet row launch needs the persistent row kernel");
            TRY(set_smem(amk::k_row32<L2, MODE>, C32::SMEM));
            LAUNCH(h, AM_K_ROW, amk::k_row32<L2, MODE><<<rows, C32::THREADS, C32::SMEM, h->stream>>>(A, spec, B, l1, rows, h->d_tw.p));
            return AM_OK;
        }
    }
    if (nsn > 1) return fail(AM_ERR_UNSUPPORTED, "multi-snippet row launch needs the persistent row kernel");
    TRY(set_smem(amk::k_row<L2, MODE>, Cfg::SMEM));
    int grid = (rows + Cfg::G - 1) / Cfg::G;
    LAUNCH(h, AM_K_ROW, amk::k_row<L2, MODE><<<grid, Cfg::THREADS, Cfg::SMEM, h->stream>>>(A, spec, B, l1, rows, h->d_tw.p));
    return AM_OK;
}
template <int MODE> am_status launch_row(am_matcher *h, int l2, float2 *A, const float2 *spec, float2 *B, int l1, int rows,
                                         int nsn = 1, size_t spec_stride = 0, size_t b_stride = 0) {
    switch (l2) {
#define C_(L) case L: return launch_row_t<L, MODE>(h, A, spec, B, l1, rows, nsn, spec_stride, b_stride);
        C_(10) C_(11) C_(12) C_(13) C_(14)
#undef C_
    }
    return fail(AM_ERR_UNSUPPORTED, "row length 2^%d not built", l2);
}

// snippets per inverse-row launch in batch mode: the persistent kernel exists for 8192-point rows with 32 elements
// per thread; AM_BATCH_SNIPPETS overrides (1 = one launch per snippet)
int batch_snippets_per_launch(int l2, size_t ns) {
    // measured on cfg 3 (64 snippets vs 24 h): 1 -> 23.6, 4 -> 25.1, 8 -> 25.0, 16 -> 25.9 stream-h/s
    static const int env = [] { const char *v = getenv("AM_BATCH_SNIPPETS"); return v && *v ? atoi(v) : 16; }();
    static const int row_stream = [] { const char *v = getenv("AM_ROW_STREAM"); return v && *v ? atoi(v) : -1; }();
    static const int row_ept = [] { const char *v = getenv("AM_ROW_EPT"); return v && *v ? atoi(v) : 32; }();
    if (l2 != 13 || row_stream == 0 || row_ept != 32 || env < 1) return 1;
    return (int)std::min<size_t>((size_t)env, ns);
}

// ---- planning -----------------------------------------------------------------------
constexpr int SMALL_MAX_LOG2 = 13;
constexpr int MAX_LOG2 = 24;

void split(int log2n, int &l1, int &l2) {
    if (log2n <= SMALL_MAX_LOG2) { l1 = 0; l2 = log2n; return; }
    l2 = std::min(13, log2n - 4);
    // 2^23 as 512 x 16384: the 512-point column kernels (16-column tiles, 32 elements per thread, run summaries) gain
    // more than the 16384-point rows (one 512-thread CTA per SM) lose: 31.0 vs 35.7 ms per 30 h at cfg 4
    if (log2n == 23) l2 = 14;
    const char *v = getenv("AM_ROW_LOG2");
    if (v && *v) {
        int x = atoi(v);
        if (x >= 10 && x <= 14 && log2n - x >= 4 && log2n - x <= 11) l2 = x;
    }
    l1 = log2n - l2;
    if (l1 > 11) { l1 = 11; l2 = log2n - 11; }
}

am_status choose_log2n(const am_matcher *h, unsigned long long outputs, int &log2n) {
    const unsigned long long m = h->m;
    if (h->cfg.fft_log2) {
        log2n = (int)h->cfg.fft_log2;
        if (log2n < 4 || log2n > MAX_LOG2) return fail(AM_ERR_INVALID, "fft_log2 %d outside [4, %d]", log2n, MAX_LOG2);
        if ((1ull << log2n) < 2 * m) return fail(AM_ERR_INVALID, "fft_log2 %d too small for a snippet of %llu samples", log2n, m);
        return AM_OK;
    }
    int l = std::max(8, ceil_log2(8 * m));
    if (l > 23) l = std::max(23, ceil_log2(2 * m));
    if (l > MAX_LOG2) return fail(AM_ERR_UNSUPPORTED, "snippet of %llu samples needs an FFT block > 2^%d", m, MAX_LOG2);
    while (l > 4 && (1ull << (l - 1)) >= 2 * m && (1ull << (l - 1)) - m + 1 >= outputs) --l;
    log2n = l;
    return AM_OK;
}

amk::StreamView snippet_view(const am_matcher *h, size_t snippet = 0) {
    amk::StreamView sv;
    sv.x = h->d_snip.p + snippet * h->m;
    sv.fmt = amk::FMT_F32_MONO;
    sv.buf_first = 0;
    sv.buf_frames = (long long)h->m;
    sv.total = (long long)h->m;
    sv.lead = 0;
    return sv;
}

am_status ensure_workspace(am_matcher *h, int log2n, unsigned long long pairs_total, unsigned long long &pairs_per_group) {
    // launch groups of 256 block pairs at N = 2^22 (8 GB): larger groups amortise the launch tails and gaps (64 pairs
    // 22.8, 128 pairs 22.5, 256 pairs 22.2 ms per 24 h); a batch keeps 2 GB groups, its second workspace holds
    // several snippets per group
    // The defaults assume a B200 to ourselves; when the device cannot give that much (other tenants, other
    // matchers) the group is halved until it fits -- smaller launch groups cost a few percent, not the call.
    size_t budget = env_mb("AM_WORKSPACE_MB", h->S > 1 ? 2048 : 8192) << 20;
    const size_t per_pair = sizeof(float2) << log2n;
    int l1, l2;
    split(log2n, l1, l2);
    for (;;) {
        unsigned long long g = std::max<size_t>(1, budget / per_pair);
        g = std::min<unsigned long long>(g, pairs_total);
        g = std::min<unsigned long long>(g, 32768);
        am_status st = h->d_A.reserve((size_t)g << log2n);
        if (st == AM_OK && h->S > 1) st = h->d_B.reserve(((size_t)g << log2n) * (size_t)batch_snippets_per_launch(l2, h->S));
        if (st == AM_OK) { pairs_per_group = g; return AM_OK; }
        if (st != AM_ERR_NOMEM || g <= 1) return st;
        cudaGetLastError();
        h->d_A.release(); h->d_B.release();
        budget = std::max(per_pair, (size_t)(g / 2) * per_pair);
    }
}

// conjugate spectrum of the zero-padded snippet for block length 2^log2n, computed once per
// matcher in double precision on the device and stored as correctly rounded fp32 in the order
// the row kernel consumes ([k1][k2] for the four-step split)
am_status get_spectrum(am_matcher *h, int log2n, float2 **out) {
    auto it = h->spectra.find(log2n);
    if (it != h->spectra.end()) { *out = it->second; return AM_OK; }
    const long long n = 1ll << log2n;
    int l1, l2;
    split(log2n, l1, l2);
    float2 *spec = nullptr;
    double2 *buf = nullptr;
    CU(cudaMalloc((void **)&spec, h->S * (sizeof(float2) << log2n)));
    cudaError_t e = cudaMalloc((void **)&buf, 2 * (sizeof(double2) << log2n));
    if (e != cudaSuccess) { cudaFree(spec); return fail(AM_ERR_NOMEM, "spectrum scratch: %s", cudaGetErrorString(e)); }
    const unsigned grid_n = (unsigned)((n + 255) / 256), grid_h = (unsigned)((n / 2 + 255) / 256);
    for (size_t sn = 0; sn < h->S; ++sn) {
        double2 *a = buf, *b = buf + n;
        prof_begin(h, AM_K_SPECTRUM);
        amk::k_spec64_load<<<grid_n, 256, 0, h->stream>>>(snippet_view(h, sn), n, a);
        h->stats.kernel_launches++;
        for (long long ns = 1; ns < n; ns <<= 1) {
            amk::k_spec64_pass<<<grid_h, 256, 0, h->stream>>>(a, b, n / 2, ns);
            h->stats.kernel_launches++;
            std::swap(a, b);
        }
        amk::k_spec64_store<<<grid_n, 256, 0, h->stream>>>(a, n, l1, l1 ? l2 : 0, spec + sn * (size_t)n);
        h->stats.kernel_launches++;
        prof_end(h);
    }
    e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(buf);
    if (e != cudaSuccess) { cudaFree(spec); return fail(AM_ERR_CUDA, "snippet spectrum: %s", cudaGetErrorString(e)); }
    h->spectra[log2n] = spec;
    *out = spec;
    return AM_OK;
}

// correlation outputs [g0, g1) (virtual offsets) of snippets [s0, s0 + ns) -> c[j * c_stride + g - c_g0].
// With several snippets the stream-side work (column pass + forward row pass) is done once per block
// group and only the multiply + inverse passes run per snippet.
am_status run_correlation(am_matcher *h, const amk::StreamView &sv, long long g0, long long g1, int log2n, int scale,
                          float *c, size_t c_stride, long long c_g0, size_t s0, size_t ns, amp::RunRecs rsum = amp::RunRecs{nullptr},
                          float theta = 0.f) {
    if (g1 <= g0 || ns == 0) return AM_OK;
    const float inv_n = (float)(1.0 / (double)(1ull << log2n));
    auto scalar_of = [&](size_t sn) { return inv_n * (scale ? h->inv_ac[sn] : 1.0f); };
    if (h->m <= (size_t)amk::DIRECT_MAX_M && !getenv("AM_NO_DIRECT")) {
        // very short snippet: direct sums
        const unsigned long long nb = (unsigned long long)(g1 - g0 + 255) / 256;
        if (nb > 0x7fffffffull) return fail(AM_ERR_UNSUPPORTED, "direct path: too many outputs");
        for (size_t j = 0; j < ns; ++j)
            LAUNCH(h, AM_K_DIRECT, amk::k_direct<<<(unsigned)nb, 256, 0, h->stream>>>(
                                       sv, h->d_snip.p + (s0 + j) * h->m, (int)h->m, g0, g1, c + j * c_stride, c_g0,
                                       scale ? h->inv_ac[s0 + j] : 1.0f));
        h->stats.fft_log2 = 0; h->stats.log2_n1 = 0; h->stats.log2_n2 = 0;
        return AM_OK;
    }
    float2 *spec_all;
    TRY(get_spectrum(h, log2n, &spec_all));
    long long N = 1ll << log2n, VN = N - (long long)h->m + 1;
    if (VN >= 4096) VN &= ~31ll;        // blocks advance by a multiple of 32 frames: 16-byte aligned PCM rows, 128-byte aligned output rows
    const unsigned long long nblocks = (unsigned long long)((g1 - g0 + VN - 1) / VN);
    const unsigned long long pairs_total = (nblocks + 1) / 2;
    int l1, l2;
    split(log2n, l1, l2);
    h->stats.fft_log2 = log2n; h->stats.log2_n1 = l1; h->stats.log2_n2 = l2;
    h->stats.fft_blocks += nblocks;
    amk::BlockGroup g;
    g.sv = sv; g.g_end = g1; g.VN = VN; g.c = c; g.c_g0 = c_g0; g.scalar = 0.f; g.rsum = rsum; g.theta = theta;
    unsigned long long ppg = 1u << 20;     // pairs per launch
    if (l1 != 0) TRY(ensure_workspace(h, log2n, pairs_total, ppg));
    for (unsigned long long p0 = 0; p0 < pairs_total; p0 += ppg) {
        unsigned long long np = std::min(ppg, pairs_total - p0);
        g.g0 = g0 + (long long)(2 * p0) * VN;
        g.nblocks = (int)std::min<unsigned long long>(2 * np, nblocks - 2 * p0);
        const int rows = (int)(np << l1);
        if (l1 == 0) {
            for (size_t j = 0; j < ns; ++j) {
                g.c = c + j * c_stride; g.scalar = scalar_of(s0 + j);
                TRY(launch_small(h, log2n, g, spec_all + (s0 + j) * (size_t)N));
            }
        } else if (ns == 1) {
            g.c = c; g.scalar = scalar_of(s0);
            TRY(launch_col<false>(h, l1, g, l2, h->d_A.p));
            TRY(launch_row<amk::ROW_FUSED>(h, l2, h->d_A.p, spec_all + s0 * (size_t)N, nullptr, l1, rows));
            TRY(launch_col<true>(h, l1, g, l2, h->d_A.p));
        } else {
            TRY(launch_col<false>(h, l1, g, l2, h->d_A.p));
            TRY(launch_row<amk::ROW_FORWARD>(h, l2, h->d_A.p, nullptr, nullptr, l1, rows));
            // multiply + inverse rows for sb snippets per launch (the forward rows come from DRAM once per sb
            // snippets instead of once per snippet), then the inverse column pass per snippet
            const size_t sb = (size_t)batch_snippets_per_launch(l2, ns), b_stride = (size_t)ppg << log2n;
            for (size_t j0 = 0; j0 < ns; j0 += sb) {
                const size_t nj = std::min(sb, ns - j0);
                TRY(launch_row<amk::ROW_INVERSE>(h, l2, h->d_A.p, spec_all + (s0 + j0) * (size_t)N, h->d_B.p, l1, rows, (int)nj,
                                                 (size_t)N, b_stride));
                for (size_t j = j0; j < j0 + nj; ++j) {
                    g.c = c + j * c_stride; g.scalar = scalar_of(s0 + j);
                    if (rsum.rec) g.rsum = rsum.offset((long long)(j * (c_stride >> 4)));
                    TRY(launch_col<true>(h, l1, g, l2, h->d_B.p + (j - j0) * b_stride));
                }
            }
        }
    }
    return AM_OK;
}

am_status init_common(am_matcher *h, uint32_t sr, const am_config *cfg) {
    CU(cudaGetDevice(&h->device));
    h->sr = sr;
    if (cfg) h->cfg = *cfg; else am_config_default(&h->cfg);
    if (h->cfg.overlap_s < 0) h->cfg.overlap_s = (double)h->m / (double)sr;
    if (!(h->cfg.chunk_size_s > 0)) return fail(AM_ERR_INVALID, "chunk_size_s must be > 0");
    memset(&h->stats, 0, sizeof h->stats);
    CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&h->ev_up[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
    // master twiddle table W_{2^14}^j from double precision
    std::vector<float2> tw(amfft::TW_N);
    for (int j = 0; j < amfft::TW_N; ++j) {
        double a = -2.0 * M_PI * (double)j / (double)amfft::TW_N;
        tw[j] = make_float2((float)cos(a), (float)sin(a));
    }
    TRY(h->d_tw.reserve(amfft::TW_N));
    CU(cudaMemcpy(h->d_tw.p, tw.data(), sizeof(float2) * amfft::TW_N, cudaMemcpyHostToDevice));
    TRY(h->d_count.reserve(2));
    // sum s^2 per snippet on the device (double accumulation)
    h->sumsq.assign(h->S, 0.0);
    h->inv_ac.assign(h->S, 0.f);
    double *d_acc = (double *)h->d_count.p;
    for (size_t sn = 0; sn < h->S; ++sn) {
        CU(cudaMemsetAsync(d_acc, 0, sizeof(double), h->stream));
        LAUNCH(h, AM_K_SPECTRUM, amk::k_sumsq<<<(unsigned)std::min<size_t>(1024, (h->m + 255) / 256), 256, 0, h->stream>>>(
                                     snippet_view(h, sn), (long long)h->m, d_acc));
        CU(cudaMemcpyAsync(&h->sumsq[sn], d_acc, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        h->inv_ac[sn] = (float)(1.0 / h->sumsq[sn]);
    }
    return AM_OK;
}

// Duration::from_secs_f64(start / sr) truncated to whole nanoseconds (src/matcher/mod.rs:127-129)
uint64_t start_ns(uint64_t start, uint32_t sr) {
    double v = (double)start / (double)sr;
    if (!(v > 0.0)) return 0;
    int e;
    double fr = frexp(v, &e);
    uint64_t mant = (uint64_t)ldexp(fr, 53);
    e -= 53;
    unsigned __int128 t = (unsigned __int128)mant * 1000000000ull;
    if (e >= 0) return (uint64_t)(t << e);
    if (-e >= 127) return 0;
    return (uint64_t)(t >> (-e));
}
uint64_t duration_ns(double secs) { return secs <= 0 ? 0 : (uint64_t)llround(secs * 1e9); }

// ---- synthetic workload generator (bench/test utility) -------------------
__device__ __forceinline__ unsigned long long hash64(unsigned long long seed, unsigned long long n) {
    unsigned long long z = seed + n * 0x9E3779B97F4A7C15ull;     // SplitMix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void k_synth_pcm16(unsigned long long seed, unsigned long long first, size_t count, short *out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (short)((int)(hash64(seed, first + i) >> 50) - 8192);
}
// coloured noise: a boxcar of `taps` white samples, scaled by mul / 2^shift (integer arithmetic only)
__global__ void k_synth_coloured16(unsigned long long seed, unsigned long long first, size_t count, int taps, int mul, int shift,
                                   short *out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        long long acc = 0;
        for (int k = 0; k < taps; ++k) acc += (int)(hash64(seed, first + i - (unsigned long long)k) >> 50) - 8192;
        long long v = (acc * mul) >> shift;
        out[i] = (short)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
    }
}
__global__ void k_synth_plant(short *pcm, size_t frames, int channels, const short *snip, size_t m,
                              unsigned long long offset, int shift) {
    size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= m) return;
    size_t f = offset + j;
    if (f >= frames) return;
    for (int ch = 0; ch < channels; ++ch) {
        int v = (pcm[f * channels + ch] >> 1) + (snip[j] >> shift);
        v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
        pcm[f * channels + ch] = (short)v;
    }
}

}  // namespace

extern "C" {

const char *am_last_error(void) { return g_err; }
int am_abi_version(void) { return AM_ABI_VERSION; }
int am_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void am_config_default(am_config *cfg) {
    if (!cfg) return;
    cfg->chunk_size_s = 60.0;       // src/matcher/args.rs:71
    cfg->overlap_s = -1.0;          // snippet duration, audio_matcher.rs:41
    cfg->distance_s = 480.0;        // src/matcher/args.rs:75
    cfg->prominence = 0.13f;        // 13 / 100, args.rs:19 + audio_matcher.rs:44
    cfg->fft_log2 = 0;
    cfg->max_peaks_per_chunk = 0;
    cfg->reserved = 0;
}

size_t am_out_len(size_t n, size_t m, am_mode mode) {
    if (n == 0 || m == 0) return 0;
    switch (mode) {
    case AM_MODE_FULL: return n + m - 1;
    case AM_MODE_SAME: return n;
    default: return n >= m ? n - m + 1 : 0;
    }
}

size_t am_valid_len(size_t n, size_t m) { return am_out_len(n, m, AM_MODE_VALID); }

static am_status create_impl(const void *data, size_t m, size_t S, int fmt, uint32_t sr, const am_config *cfg,
                             am_matcher **out) {
    if (!out) return fail(AM_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!data || m == 0 || S == 0) return fail(AM_ERR_INVALID, "empty snippet");
    if (S > 4096) return fail(AM_ERR_INVALID, "too many snippets in one batch (%zu)", S);
    if (sr == 0) return fail(AM_ERR_INVALID, "sample rate 0");
    if (m > (1ull << (MAX_LOG2 - 1))) return fail(AM_ERR_UNSUPPORTED, "snippet of %zu samples exceeds 2^%d", m, MAX_LOG2 - 1);
    if (am_device_count() == 0) return fail(AM_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
    am_matcher *h = new (std::nothrow) am_matcher();
    if (!h) return fail(AM_ERR_NOMEM, "out of host memory");
    h->m = m;
    h->S = S;
    am_status st = h->d_snip.reserve(m * S);
    if (st == AM_OK) {
        if (fmt == AM_FMT_F32_MONO) {
            cudaError_t e = cudaMemcpy(h->d_snip.p, data, m * S * sizeof(float), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) st = fail(AM_ERR_CUDA, "snippet upload: %s", cudaGetErrorString(e));
        } else {
            // scale / downmix on the device with the same load path the stream uses
            void *raw = nullptr;
            size_t bytes = m * fmt_bytes(fmt);
            cudaError_t e = cudaMalloc(&raw, bytes);
            if (e == cudaSuccess) e = cudaMemcpy(raw, data, bytes, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) {
                amk::StreamView sv{raw, fmt, 0, (long long)m, (long long)m, 0};
                amk::k_to_f32<<<(unsigned)((m + 255) / 256), 256>>>(sv, 0, (long long)m, h->d_snip.p);
                e = cudaDeviceSynchronize();
            }
            if (raw) cudaFree(raw);
            if (e != cudaSuccess) st = fail(AM_ERR_CUDA, "snippet upload: %s", cudaGetErrorString(e));
        }
    }
    if (st == AM_OK) st = init_common(h, sr, cfg);
    if (st != AM_OK) { am_matcher_destroy(h); return st; }
    *out = h;
    return AM_OK;
}

am_status am_matcher_create(const float *snippet, size_t m, uint32_t sr, const am_config *cfg, am_matcher **out) {
    return create_impl(snippet, m, 1, AM_FMT_F32_MONO, sr, cfg, out);
}
am_status am_matcher_create_batch(const float *snippets, size_t m, size_t n_snippets, uint32_t sr, const am_config *cfg,
                                  am_matcher **out) {
    return create_impl(snippets, m, n_snippets, AM_FMT_F32_MONO, sr, cfg, out);
}
size_t am_matcher_snippet_count(const am_matcher *h) { return h ? h->S : 0; }
am_status am_matcher_select_snippet(am_matcher *h, size_t snippet_id) {
    if (!h) return fail(AM_ERR_INVALID, "NULL handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (snippet_id >= h->S) return fail(AM_ERR_INVALID, "snippet %zu of %zu", snippet_id, h->S);
    h->active = snippet_id;
    return AM_OK;
}
am_status am_matcher_create_pcm16(const int16_t *pcm, size_t frames, int channels, uint32_t sr, const am_config *cfg,
                                  am_matcher **out) {
    if (channels != 1 && channels != 2) return fail(AM_ERR_INVALID, "channels must be 1 or 2");
    return create_impl(pcm, frames, 1, channels == 2 ? AM_FMT_I16_STEREO : AM_FMT_I16_MONO, sr, cfg, out);
}

void am_matcher_destroy(am_matcher *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto &kv : h->spectra) cudaFree(kv.second);
    h->d_snip.release(); h->d_tw.release(); h->d_A.release(); h->d_B.release(); h->d_c.release(); h->d_tmin.release();
    h->d_tmax.release(); h->d_rsum.release(); h->d_peaks.release(); h->d_redo.release(); h->d_count.release(); h->d_sched.release();
    h->d_stage[0].release(); h->d_stage[1].release(); h->stager.release(); h->d_gsend.release(); h->d_grecv.release();
    h->d_fpeaks.release(); h->d_fcount.release(); h->d_fredo.release();
    for (int i = 0; i < 2; ++i) {
        if (h->ev_up[i]) cudaEventDestroy(h->ev_up[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    delete h;
}

am_status am_matcher_set_stream(am_matcher *h, void *cuda_stream) {
    if (!h) return fail(AM_ERR_INVALID, "NULL handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->stream = (cudaStream_t)cuda_stream;
    return AM_OK;
}
am_status am_matcher_set_config(am_matcher *h, const am_config *cfg) {
    if (!h || !cfg) return fail(AM_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    if (!(cfg->chunk_size_s > 0)) return fail(AM_ERR_INVALID, "chunk_size_s must be > 0");
    h->cfg = *cfg;
    if (h->cfg.overlap_s < 0) h->cfg.overlap_s = (double)h->m / (double)h->sr;
    return AM_OK;
}
am_status am_matcher_get_stats(const am_matcher *h, am_stats *out) {
    if (!h || !out) return fail(AM_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(const_cast<am_matcher *>(h)->mu);
    *out = h->stats;
    return AM_OK;
}
am_status am_matcher_set_progress(am_matcher *h, am_progress_fn fn, void *user) {
    if (!h) return fail(AM_ERR_INVALID, "NULL handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->progress = fn;
    h->progress_user = user;
    return AM_OK;
}

am_status am_matcher_set_profiling(am_matcher *h, int on) {
    if (!h) return fail(AM_ERR_INVALID, "NULL handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->profiling = on != 0;
    for (int i = 0; i < AM_KERNEL_CLASSES; ++i) { h->cls_ms[i] = 0; h->cls_launches[i] = 0; }
    return AM_OK;
}
am_status am_matcher_get_kernel_times(const am_matcher *h, am_kernel_time *out, size_t cap, size_t *n_out) {
    static const char *names[AM_KERNEL_CLASSES] = {"k_col_fwd", "k_row", "k_col_inv", "k_small", "k_direct",
                                                   "k_tile_minmax", "k_chunk_peaks", "spectrum"};
    if (!h || !n_out || (!out && cap)) return fail(AM_ERR_INVALID, "NULL argument");
    size_t k = 0;
    for (int i = 0; i < AM_KERNEL_CLASSES; ++i) {
        if (!h->cls_launches[i]) continue;
        if (k < cap) {
            out[k].kernel_class = i;
            out[k].launches = h->cls_launches[i];
            out[k].total_ms = h->cls_ms[i];
            strncpy(out[k].name, names[i], sizeof out[k].name - 1);
            out[k].name[sizeof out[k].name - 1] = 0;
        }
        ++k;
    }
    *n_out = k;
    return AM_OK;
}

am_status am_inverse_sample_auto_correlation(am_matcher *h, float *out) {
    if (!h || !out) return fail(AM_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    *out = h->inv_ac[h->active];
    return AM_OK;
}

am_status am_correlate(am_matcher *h, const void *within, size_t n, am_sample_fmt fmt, am_mem within_mem,
                       am_mode mode, int scale, float *out, size_t cap, am_mem out_mem, size_t *out_len) {
    if (!h || !out_len) return fail(AM_ERR_INVALID, "NULL argument");
    if ((int)fmt < 0 || (int)fmt > 2 || (int)mode < 0 || (int)mode > 2) return fail(AM_ERR_INVALID, "bad fmt/mode");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->in_session) return fail(AM_ERR_INVALID, "a push session is open on this matcher");
    CU(cudaSetDevice(h->device));
    memset(&h->stats, 0, sizeof h->stats);
    const size_t olen = am_out_len(n, h->m, mode);
    *out_len = olen;
    if (olen == 0) return AM_OK;
    if (!within || !out) return fail(AM_ERR_INVALID, "NULL buffer");
    if (olen > cap) return fail(AM_ERR_CAPACITY, "output needs %zu floats, capacity %zu", olen, cap);
    amk::StreamView sv;
    sv.fmt = (int)fmt; sv.buf_first = 0; sv.buf_frames = (long long)n; sv.total = (long long)n;
    const long long full = (long long)(n + h->m - 1);
    sv.lead = (long long)(h->m - 1) - (full - (long long)olen) / 2;      // centered(): audio_matcher.rs:460-464
    if (within_mem == AM_MEM_HOST) {
        size_t bytes = n * fmt_bytes(fmt);
        TRY(h->d_stage[0].reserve(bytes));
        CU(cudaMemcpyAsync(h->d_stage[0].p, within, bytes, cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += bytes;
        sv.x = h->d_stage[0].p;
    } else sv.x = within;
    float *d_out = out;
    if (out_mem == AM_MEM_HOST) {
        TRY(h->d_c.reserve(olen));
        d_out = h->d_c.p;
    }
    int log2n;
    TRY(choose_log2n(h, olen, log2n));
    TRY(run_correlation(h, sv, 0, (long long)olen, log2n, scale, d_out, 0, 0, h->active, 1));
    if (out_mem == AM_MEM_HOST) {
        CU(cudaMemcpyAsync(out, d_out, olen * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        h->stats.d2h_bytes += olen * sizeof(float);
    }
    CU(cudaStreamSynchronize(h->stream));
    prof_collect(h);
    h->stats.frames = n;
    return AM_OK;
}

static void chunk_params(const am_matcher *h, long long &C, long long &ov) {
    ov = llround(h->cfg.overlap_s * (double)h->sr);      // audio_matcher.rs:99
    C = llround(h->cfg.chunk_size_s * (double)h->sr);    // audio_matcher.rs:100
}

size_t am_num_chunks(const am_matcher *h, size_t frames) {
    if (!h) return 0;
    long long C, ov;
    chunk_params(h, C, ov);
    if (C <= 0 || frames == 0) return 0;
    return (frames + (size_t)C - 1) / (size_t)C;          // chunked(C + ov, C), audio_matcher.rs:104
}

am_status am_chunk_geometry(const am_matcher *h, size_t *chunk, size_t *overlap) {
    if (!h) return fail(AM_ERR_INVALID, "NULL handle");
    long long C, ov;
    chunk_params(h, C, ov);
    if (chunk) *chunk = (size_t)std::max<long long>(C, 0);
    if (overlap) *overlap = (size_t)std::max<long long>(ov, 0);
    return AM_OK;
}

am_status am_shard_frames(const am_matcher *h, size_t total_frames, size_t first_chunk, size_t num_chunks, size_t *lo,
                          size_t *hi) {
    if (!h || !lo || !hi) return fail(AM_ERR_INVALID, "NULL argument");
    long long C, ov;
    chunk_params(h, C, ov);
    if (C <= 0) return fail(AM_ERR_INVALID, "chunk size rounds to 0 samples");
    const size_t total_chunks = am_num_chunks(h, total_frames);
    if (first_chunk >= total_chunks || num_chunks == 0) { *lo = *hi = std::min((size_t)C * first_chunk, total_frames); return AM_OK; }
    num_chunks = std::min(num_chunks, total_chunks - first_chunk);
    *lo = (size_t)C * first_chunk;
    *hi = std::min(total_frames, (size_t)C * (first_chunk + num_chunks) + (size_t)std::max<long long>(ov, 0));
    return AM_OK;
}

int am_is_overshadowed(const am_peak *element, const am_peak *other, uint32_t sr, double max_distance_s) {
    if (!element || !other) return 0;                      // None never overshadows, audio_matcher.rs:149
    uint64_t e = start_ns(element->start, sr), b = start_ns(other->start, sr);
    if (e < b) std::swap(e, b);
    return ((e - b) < duration_ns(max_distance_s)) && (other->prominence > element->prominence);
}

am_status am_merge_peaks(am_peak *peaks, size_t n, uint32_t sr, double distance_s, am_peak *out, size_t cap,
                         size_t *n_out) {
    if (!n_out || (n && (!peaks || !out))) return fail(AM_ERR_INVALID, "NULL argument");
    if (sr == 0) return fail(AM_ERR_INVALID, "sample rate 0");
    // sorted_by position.start (stable, audio_matcher.rs:135); equal starts keep chunk order
    // (a batch is handled as independent runs: snippet-major, neighbours only within a snippet)
    std::stable_sort(peaks, peaks + n, [](const am_peak &a, const am_peak &b) {
        if (a.snippet_id != b.snippet_id) return a.snippet_id < b.snippet_id;
        if (a.start != b.start) return a.start < b.start;
        return a.chunk < b.chunk;
    });
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {                       // filter_surrounding, audio_matcher.rs:136-139
        const am_peak *before = (i > 0 && peaks[i - 1].snippet_id == peaks[i].snippet_id) ? &peaks[i - 1] : nullptr;
        const am_peak *after = (i + 1 < n && peaks[i + 1].snippet_id == peaks[i].snippet_id) ? &peaks[i + 1] : nullptr;
        if (am_is_overshadowed(&peaks[i], before, sr, distance_s) || am_is_overshadowed(&peaks[i], after, sr, distance_s))
            continue;
        if (k < cap) out[k] = peaks[i];
        ++k;
    }
    *n_out = k;
    if (k > cap) return fail(AM_ERR_CAPACITY, "%zu peaks, capacity %zu", k, cap);
    return AM_OK;
}

// ---- calc_chunks on the device -----------------------------------------------------------------------
// RangePlan: geometry, buffers and kernel parameters of one calc_chunks call over logical chunks
// [c_first, c_last) of a stream of L frames.  plan_range() sizes everything, compute_segment() launches the
// transforms + peak kernels of one segment whose frames are resident on the device, finish_range() collects the
// counters and repeats marked chunks densely.  The one-shot entry points and the push session share them.
struct RangePlan {
    long long C = 0, ov = 0, L = 0, m = 0, c_first = 0, c_last = 0, K = 1, seg_c_len = 0, tiles_stride = 0;
    size_t S = 1, fb = 2, pk_smem = 0, dev_cap = 0, num_chunks = 0;
    int log2n = 0, pk_cap = 1024, sm_tiles = 0, scale = 1, fmt = 0;
    bool summary = false;
    unsigned long long min_dist = 0;
    float theta = 0.f;
    amp::PeakOut po;
    amp::RunRecs recs{nullptr};
    long long chunk_end(long long i) const {                 // one past the last global offset chunk i evaluates
        long long n = std::min(C + ov, L - C * i);
        return n >= m ? C * i + n - m + 1 : C * i;
    }
    long long seg_out_end(long long i0, long long i1) const {   // one past the last output of chunks [i0, i1); C i0 if none
        for (long long i = i1 - 1; i >= i0; --i)
            if (chunk_end(i) > C * i) return chunk_end(i);
        return C * i0;
    }
};

// `mem_host`: host streams take 512 MB segments (a segment is also the unit of the double-buffered upload and the
// first one is not overlapped with compute), resident streams 4 GB.
static am_status plan_range(am_matcher *h, size_t total_frames, am_sample_fmt fmt, bool mem_host, int scale, size_t first_chunk,
                            size_t num_chunks, RangePlan &pl, bool *nothing_to_do) {
    *nothing_to_do = true;
    if ((int)fmt < 0 || (int)fmt > 2) return fail(AM_ERR_INVALID, "bad sample format");
    CU(cudaSetDevice(h->device));
    memset(&h->stats, 0, sizeof h->stats);
    long long C, ov;
    chunk_params(h, C, ov);
    const long long L = (long long)total_frames, m = (long long)h->m;
    if (C <= 0) return fail(AM_ERR_INVALID, "chunk size rounds to 0 samples");
    if (C + ov >= (1ll << 31)) return fail(AM_ERR_UNSUPPORTED, "chunk window of %lld samples exceeds 2^31", C + ov);
    const size_t total_chunks = am_num_chunks(h, total_frames);
    if (first_chunk >= total_chunks || num_chunks == 0) return AM_OK;
    num_chunks = std::min(num_chunks, total_chunks - first_chunk);
    pl.C = C; pl.ov = ov; pl.L = L; pl.m = m; pl.S = h->S; pl.fb = fmt_bytes(fmt); pl.scale = scale; pl.fmt = (int)fmt;
    pl.c_first = (long long)first_chunk; pl.c_last = pl.c_first + (long long)num_chunks; pl.num_chunks = num_chunks;
    const long long out_end = pl.seg_out_end(pl.c_first, pl.c_last);
    if (out_end <= C * pl.c_first) return AM_OK;             // every window shorter than the snippet
    TRY(choose_log2n(h, (unsigned long long)(out_end - C * pl.c_first), pl.log2n));

    // Summary mode: the inverse column kernel writes one {min, max, first, last} record per aligned run of 16
    // outputs and the outputs themselves only where the run maximum reaches theta = prominence / 2; the peak
    // kernels work from the records.  Needs 16-column tiles, run-aligned chunks whose full windows end on a run
    // boundary or one output after it (V mod 16 <= 1: the usual ov = m gives V = C + 1), and a positive threshold.
    // A chunk the records cannot decide exactly (a chunk minimum below theta - prominence whose kept peaks do not
    // cover it, an odd-length window in the middle of a segment) is marked on the device and exactly those chunks
    // are repeated on a dense correlation afterwards.
    int sp_l1, sp_l2;
    split(pl.log2n, sp_l1, sp_l2);
    const long long v_full = C + ov >= m ? C + ov - m + 1 : 0;
    pl.summary = col_inv_writes_runs(sp_l1) && h->m > (size_t)amk::DIRECT_MAX_M && (C % 16) == 0 && (v_full & 15) <= 1 &&
                 h->cfg.prominence > 0.f && std::isfinite(h->cfg.prominence);
    {
        static const int env = [] { const char *v = getenv("AM_SUMMARY"); return v && *v ? atoi(v) : 1; }();
        if (!env) pl.summary = false;
    }

    // segments of K logical chunks share one correlation buffer per snippet; a batch keeps at
    // least ~48 M outputs per snippet per segment so that a segment still spans several block pairs
    const size_t S = h->S;
    // Resident streams take 16 GB segments (~1480 chunks: the per-chunk peak kernel fills the GPU, fewer ragged launch
    // groups; 512 MB 25.7, 4 GB 22.8, 8 GB 22.5, 16 GB 22.2 ms per 24 h).
    size_t seg_mb = mem_host ? 512 : 16384;
    if (S > 1 && !getenv("AM_SEGMENT_MB")) {
        // a batch shares the budget between its S correlation buffers; short segments mean short launch groups
        // (measured 17.2 vs 13.5 us per snippet and block pair in the inverse row pass), so take up to a quarter of
        // the free device memory, at most 32 GB
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) seg_mb = std::max<size_t>(seg_mb, std::min<size_t>(free_b >> 22, 32768));
    }
    size_t seg_floats = (env_mb("AM_SEGMENT_MB", seg_mb) << 20) / sizeof(float) / S;
    if (S > 1) seg_floats = std::max<size_t>(seg_floats, (size_t)48 << 20);
    pl.tiles_stride = ((C + std::max<long long>(ov - m + 1, 1)) + amp::TP - 1) / amp::TP + 1;
    for (;;) {
        // the segment buffers are sized for a B200 to ourselves; when the device cannot give that much the segment is
        // halved (down to one logical chunk) until they fit
        pl.K = std::max<long long>(1, (long long)(seg_floats / (size_t)C));
        pl.K = std::min<long long>(pl.K, (long long)num_chunks);
        pl.seg_c_len = ((pl.K * C + std::max<long long>(ov - m + 1, 0) + 1 + 15) / 16) * 16;   // float4 / run alignment per snippet
        am_status st = h->d_c.reserve((size_t)pl.seg_c_len * S);
        if (st == AM_OK && pl.summary) st = h->d_rsum.reserve(((size_t)pl.seg_c_len * S) >> 4);
        if (st == AM_OK) st = h->d_tmin.reserve((size_t)(pl.K * pl.tiles_stride) * S);
        if (st == AM_OK) st = h->d_tmax.reserve((size_t)(pl.K * pl.tiles_stride) * S);
        if (st == AM_OK) break;
        if (st != AM_ERR_NOMEM || pl.K <= 1) return st;
        cudaGetLastError();
        h->d_c.release(); h->d_rsum.release(); h->d_tmin.release(); h->d_tmax.release();
        seg_floats = (size_t)(pl.K / 2) * (size_t)C;
    }
    pl.recs = amp::RunRecs{h->d_rsum.p};
    pl.pk_cap = h->cfg.max_peaks_per_chunk ? (int)h->cfg.max_peaks_per_chunk : 1024;
    if (amp::chunk_peaks_smem(pl.pk_cap, 0) > 200 * 1024) return fail(AM_ERR_INVALID, "max_peaks_per_chunk %d too large", pl.pk_cap);
    pl.sm_tiles = 0;                                         // tile summaries staged in shared memory when they fit
    if (amp::chunk_peaks_smem(pl.pk_cap, (int)pl.tiles_stride) <= 96 * 1024) pl.sm_tiles = (int)pl.tiles_stride;
    pl.pk_smem = amp::chunk_peaks_smem(pl.pk_cap, pl.sm_tiles);
    TRY(set_smem(amp::k_chunk_peaks<false>, pl.pk_smem));
    TRY(set_smem(amp::k_chunk_peaks<true>, pl.pk_smem));
    pl.dev_cap = std::min<size_t>((size_t)num_chunks * (size_t)pl.pk_cap * S, (size_t)1 << 22);
    TRY(h->d_peaks.reserve(pl.dev_cap));
    TRY(h->d_redo.reserve((size_t)num_chunks * S));
    pl.po.peaks = h->d_peaks.p; pl.po.cap = pl.dev_cap; pl.po.count = h->d_count.p; pl.po.flags = (unsigned *)(h->d_count.p + 1);
    pl.po.redo = h->d_redo.p; pl.po.redo_first = pl.c_first; pl.po.redo_stride = (long long)num_chunks;
    pl.min_dist = (unsigned long long)h->cfg.distance_s * (unsigned long long)h->sr;  // as_secs(), :228
    pl.theta = 0.5f * h->cfg.prominence;
    CU(cudaMemsetAsync(h->d_count.p, 0, 2 * sizeof(unsigned long long), h->stream));
    if (pl.summary) CU(cudaMemsetAsync(h->d_redo.p, 0, (size_t)num_chunks * S, h->stream));
    h->stats.summary_mode = pl.summary ? 1 : 0;
    *nothing_to_do = false;
    return AM_OK;
}

// Transforms + peak kernels of logical chunks [i0, i1) (at most K of them) on h->stream; `sv` must hold frames
// [C i0, min(L, seg_out_end + m - 1)).  redo: the dense repeat of the chunks the summary pass marked.
static am_status compute_segment(am_matcher *h, const RangePlan &pl, const amk::StreamView &sv, long long i0, long long i1,
                                 bool sum, bool redo) {
    const long long g0 = pl.C * i0, g1 = pl.seg_out_end(i0, i1);
    if (g1 <= g0) return AM_OK;
    const amp::RunRecs no_recs{nullptr};
    amp::ChunkGeom cg;
    cg.C = pl.C; cg.ov = pl.ov; cg.m = pl.m; cg.total = pl.L; cg.first_chunk = i0; cg.c_g0 = g0; cg.tiles_stride = (int)pl.tiles_stride;
    cg.c_stride = pl.seg_c_len; cg.seg_end = g1 - g0;
    TRY(run_correlation(h, sv, g0, g1, pl.log2n, pl.scale, h->d_c.p, (size_t)pl.seg_c_len, g0, 0, pl.S, sum ? pl.recs : no_recs, pl.theta));
    dim3 tgrid3((unsigned)((pl.tiles_stride + 7) / 8), (unsigned)(i1 - i0), (unsigned)pl.S), pgrid((unsigned)(i1 - i0), (unsigned)pl.S);
    if (sum) {
        LAUNCH(h, AM_K_TILE_MINMAX, amp::k_tile_from_runs<<<tgrid3, 256, 0, h->stream>>>(pl.recs, cg, h->d_tmin.p, h->d_tmax.p, pl.po));
        LAUNCH(h, AM_K_CHUNK_PEAKS, amp::k_chunk_peaks<true><<<pgrid, 256, pl.pk_smem, h->stream>>>(
                                        h->d_c.p, pl.recs, pl.theta, cg, h->d_tmin.p, h->d_tmax.p, h->cfg.prominence, pl.min_dist,
                                        pl.pk_cap, pl.sm_tiles, pl.po, 0));
    } else {
        LAUNCH(h, AM_K_TILE_MINMAX, amp::k_tile_minmax<<<tgrid3, 256, 0, h->stream>>>(h->d_c.p, cg, h->d_tmin.p, h->d_tmax.p));
        LAUNCH(h, AM_K_CHUNK_PEAKS, amp::k_chunk_peaks<false><<<pgrid, 256, pl.pk_smem, h->stream>>>(
                                        h->d_c.p, no_recs, 0.f, cg, h->d_tmin.p, h->d_tmax.p, h->cfg.prominence, pl.min_dist,
                                        pl.pk_cap, pl.sm_tiles, pl.po, redo ? 1 : 0));
    }
    return AM_OK;
}

// Upload frames [f_lo, f_hi) of a host stream into staging buffer b (copy stream; pageable memory through the pinned
// ring) and make h->stream wait for it.
static am_status upload_segment(am_matcher *h, const RangePlan &pl, const void *stream, size_t buf_first_frame, int b,
                                long long f_lo, long long f_hi, int stage_threads, amk::StreamView &sv) {
    const size_t bytes = (size_t)(f_hi - f_lo) * pl.fb;
    CU(cudaStreamWaitEvent(h->copy_stream, h->ev_done[b], 0));   // kernels that read this buffer are done
    TRY(h->d_stage[b].reserve(bytes));
    const unsigned char *src = (const unsigned char *)stream + (size_t)(f_lo - (long long)buf_first_frame) * pl.fb;
    if (stage_threads > 0 && bytes >= ((size_t)1 << 20) && h->stager.ensure(stage_threads))
        CU(h->stager.upload(h->d_stage[b].p, src, bytes, h->copy_stream));
    else
        CU(cudaMemcpyAsync(h->d_stage[b].p, src, bytes, cudaMemcpyHostToDevice, h->copy_stream));
    CU(cudaEventRecord(h->ev_up[b], h->copy_stream));
    CU(cudaStreamWaitEvent(h->stream, h->ev_up[b], 0));
    h->stats.h2d_bytes += bytes;
    sv.x = h->d_stage[b].p; sv.buf_first = f_lo; sv.buf_frames = f_hi - f_lo;
    return AM_OK;
}

// chunks (relative to c_first) the summary pass marked in [r0, r1), any snippet
static am_status marked_chunks(am_matcher *h, const RangePlan &pl, long long r0, long long r1, std::vector<unsigned char> &any) {
    any.assign((size_t)(r1 - r0), 0);
    std::vector<unsigned char> redo((size_t)(r1 - r0));
    for (size_t sn = 0; sn < pl.S; ++sn) {
        CU(cudaMemcpy(redo.data(), h->d_redo.p + sn * pl.num_chunks + (size_t)r0, redo.size(), cudaMemcpyDeviceToHost));
        h->stats.d2h_bytes += redo.size();
        for (size_t i = 0; i < redo.size(); ++i) any[i] |= redo[i];
    }
    return AM_OK;
}

static am_status fetch_counters(am_matcher *h, unsigned long long (&cnt)[2]) {
    CU(cudaMemcpyAsync(cnt, h->d_count.p, sizeof cnt, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->stats.d2h_bytes += sizeof cnt;
    return AM_OK;
}

static am_status check_counters(am_matcher *h, const RangePlan &pl, const unsigned long long (&cnt)[2], unsigned long long *count_out) {
    if ((unsigned)cnt[1] & amp::FLAG_OVERFLOW)
        return fail(AM_ERR_CAPACITY, "a chunk kept more than max_peaks_per_chunk = %d peaks, or holds more than that many local maxima of "
                                     "exactly equal height (raise it, the prominence or the distance)", pl.pk_cap);
    if (cnt[0] > pl.dev_cap) return fail(AM_ERR_CAPACITY, "%llu peaks exceed the device list capacity %zu", cnt[0], pl.dev_cap);
    *count_out = cnt[0];
    return AM_OK;
}

// The per-rank part of calc_chunks: correlation + per-chunk peak kernels for logical chunks
// [first_chunk, first_chunk + num_chunks).  The peaks stay in h->d_peaks (unordered), *count_out of them.
// The caller holds h->mu.
static am_status range_pass_device(am_matcher *h, const void *stream, size_t buf_first_frame, size_t buf_frames,
                                   size_t total_frames, am_sample_fmt fmt, am_mem mem, int scale, size_t first_chunk,
                                   size_t num_chunks, unsigned long long *count_out) {
    *count_out = 0;
    RangePlan pl;
    bool nothing = true;
    TRY(plan_range(h, total_frames, fmt, mem == AM_MEM_HOST, scale, first_chunk, num_chunks, pl, &nothing));
    if (nothing) return AM_OK;
    // frames the range reads: [C c_first, min(L, C (c_last-1) + C + ov))
    const long long need_lo = pl.C * pl.c_first, need_hi = std::min(pl.L, pl.C * pl.c_last + pl.ov);
    if (!stream || (long long)buf_first_frame > need_lo || (long long)(buf_first_frame + buf_frames) < need_hi)
        return fail(AM_ERR_INVALID, "stream buffer [%zu, %zu) does not cover frames [%lld, %lld) of chunks [%lld, %lld)",
                    buf_first_frame, buf_first_frame + buf_frames, need_lo, need_hi, pl.c_first, pl.c_last);

    // pageable host stream: copy through our own pinned ring with a pool of threads (AM_STAGE_THREADS, 0 = leave it to the driver)
    int stage_threads = 0;
    if (mem == AM_MEM_HOST) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, stream) == cudaSuccess && pa.type == cudaMemoryTypeUnregistered) stage_threads = HostStager::default_threads();
        cudaGetLastError();
    }
    int seg_idx = 0;
    auto run_segment = [&](long long i0, long long i1, bool sum, bool redo) -> am_status {
        const long long g0 = pl.C * i0, g1 = pl.seg_out_end(i0, i1);
        if (g1 <= g0) return AM_OK;
        amk::StreamView sv;
        sv.fmt = (int)fmt; sv.total = pl.L; sv.lead = 0;
        if (mem == AM_MEM_HOST) {
            TRY(upload_segment(h, pl, stream, buf_first_frame, seg_idx & 1, g0, std::min(pl.L, g1 + pl.m - 1), stage_threads, sv));
        } else {
            sv.x = stream; sv.buf_first = (long long)buf_first_frame; sv.buf_frames = (long long)buf_frames;
        }
        TRY(compute_segment(h, pl, sv, i0, i1, sum, redo));
        if (mem == AM_MEM_HOST) CU(cudaEventRecord(h->ev_done[seg_idx & 1], h->stream));
        ++seg_idx;
        if (h->progress && !redo) h->progress(h->progress_user, 0, (size_t)i0, (size_t)(i1 - i0));
        return AM_OK;
    };
    unsigned long long cnt[2] = {0, 0};
    for (long long i0 = pl.c_first; i0 < pl.c_last; i0 += pl.K) TRY(run_segment(i0, std::min(pl.c_last, i0 + pl.K), pl.summary, false));
    TRY(fetch_counters(h, cnt));
    if (pl.summary && ((unsigned)cnt[1] & amp::FLAG_NEED_DENSE) && !((unsigned)cnt[1] & amp::FLAG_OVERFLOW)) {
        // dense repeat of exactly the marked chunks: consecutive marked chunks share a segment
        std::vector<unsigned char> any;
        TRY(marked_chunks(h, pl, 0, (long long)pl.num_chunks, any));
        h->stats.summary_mode = 2;
        for (long long i = 0; i < (long long)pl.num_chunks;) {
            if (!any[(size_t)i]) { ++i; continue; }
            long long j = i;
            while (j < (long long)pl.num_chunks && any[(size_t)j] && j - i < pl.K) ++j;
            h->stats.dense_chunks += (uint32_t)(j - i);
            TRY(run_segment(pl.c_first + i, pl.c_first + j, false, true));
            i = j;
        }
        TRY(fetch_counters(h, cnt));
    }
    prof_collect(h);
    h->stats.frames = (uint64_t)(need_hi - need_lo);
    h->stats.chunks = (uint32_t)pl.num_chunks;
    TRY(check_counters(h, pl, cnt, count_out));
    if (h->progress) h->progress(h->progress_user, 1, first_chunk, pl.num_chunks);
    return AM_OK;
}

am_status am_calc_chunks_range(am_matcher *h, const void *stream, size_t buf_first_frame, size_t buf_frames,
                               size_t total_frames, am_sample_fmt fmt, am_mem mem, int scale, size_t first_chunk,
                               size_t num_chunks, int final_filter, am_peak *out, size_t cap, size_t *n_out) {
    if (!h || !n_out) return fail(AM_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    *n_out = 0;
    if (h->in_session) return fail(AM_ERR_INVALID, "a push session is open on this matcher");
    unsigned long long count = 0;
    TRY(range_pass_device(h, stream, buf_first_frame, buf_frames, total_frames, fmt, mem, scale, first_chunk, num_chunks, &count));
    std::vector<am_peak> all((size_t)count);
    if (count) {
        CU(cudaMemcpy(all.data(), h->d_peaks.p, (size_t)count * sizeof(am_peak), cudaMemcpyDeviceToHost));
        h->stats.d2h_bytes += (size_t)count * sizeof(am_peak);
    }
    if (!final_filter) {
        std::stable_sort(all.begin(), all.end(), [](const am_peak &a, const am_peak &b) {
            if (a.snippet_id != b.snippet_id) return a.snippet_id < b.snippet_id;
            if (a.chunk != b.chunk) return a.chunk < b.chunk;
            return a.height > b.height || (a.height == b.height && a.start < b.start);
        });
        *n_out = all.size();
        if (all.size() > cap) return fail(AM_ERR_CAPACITY, "%zu peaks, capacity %zu", all.size(), cap);
        if (!all.empty()) {
            if (!out) return fail(AM_ERR_INVALID, "NULL output buffer");
            memcpy(out, all.data(), all.size() * sizeof(am_peak));
        }
        return AM_OK;
    }
    std::vector<am_peak> kept(all.size());
    size_t nk = 0;
    TRY(am_merge_peaks(all.data(), all.size(), h->sr, h->cfg.distance_s, kept.data(), kept.size(), &nk));
    *n_out = nk;
    if (nk > cap) return fail(AM_ERR_CAPACITY, "%zu peaks, capacity %zu", nk, cap);
    if (nk) {
        if (!out) return fail(AM_ERR_INVALID, "NULL output buffer");
        memcpy(out, kept.data(), nk * sizeof(am_peak));
    }
    return AM_OK;
}

// ---- several files per call ------------------------------------------------------------------------------
// The reference matches its snippet against every file of args.within in turn (src/matcher/mod.rs:42-99), one
// calc_chunks per file.  Here the transforms + peak kernels of ALL files are queued first -- each file with its own
// peak list, counter and dense-repeat marks on the device -- and read back after one synchronisation: the upload of
// file k+1 (copy stream, the other staging buffer) overlaps the kernels of file k, and short files do not drain
// the GPU between calls.  A file whose run records cannot decide some chunk, or that outgrows its share of the
// peak list, is repeated through the one-file path.  Results per file equal am_calc_chunks on that file.
am_status am_calc_chunks_files(am_matcher *h, size_t n_files, const void *const *streams, const size_t *frames,
                               am_sample_fmt fmt, am_mem mem, int scale, am_peak *out, size_t cap, size_t *n_out) {
    if (!h || !n_out || (n_files && (!streams || !frames))) return fail(AM_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    for (size_t f = 0; f < n_files; ++f) n_out[f] = 0;
    if (h->in_session) return fail(AM_ERR_INVALID, "a push session is open on this matcher");
    if (n_files == 0) return AM_OK;
    const size_t F = n_files, S = h->S;
    size_t fmax = 0;
    for (size_t f = 0; f < F; ++f) {
        if (frames[f] && !streams[f]) return fail(AM_ERR_INVALID, "file %zu: NULL stream", f);
        if (frames[f] > frames[fmax]) fmax = f;
    }
    // the longest file sizes the shared buffers once (a buffer that had to grow later would synchronise the device)
    RangePlan plmax;
    bool nothing = true;
    TRY(plan_range(h, frames[fmax], fmt, mem == AM_MEM_HOST, scale, 0, (size_t)-1, plmax, &nothing));
    if (!nothing && mem == AM_MEM_HOST) {
        const size_t seg_bytes = (size_t)std::min<long long>(plmax.L, plmax.K * plmax.C + plmax.ov) * plmax.fb;
        TRY(h->d_stage[0].reserve(seg_bytes));
        if (F > 1 || (long long)plmax.num_chunks > plmax.K) TRY(h->d_stage[1].reserve(seg_bytes));
    }
    std::vector<size_t> pk_off(F + 1, 0), redo_off(F + 1, 0), pk_cap(F, 0);
    const size_t per_chunk = h->cfg.max_peaks_per_chunk ? h->cfg.max_peaks_per_chunk : 1024;
    const size_t min_cap = getenv("AM_FILES_MIN_CAP") ? env_mb("AM_FILES_MIN_CAP", 1) : 0;   // test knob: a file's share of the peak list
    for (size_t f = 0; f < F; ++f) {
        const size_t nc = am_num_chunks(h, frames[f]);
        pk_cap[f] = std::min(nc * per_chunk * S, min_cap ? min_cap : std::max<size_t>(4096, 16 * nc * S));
        pk_off[f + 1] = pk_off[f] + pk_cap[f];
        redo_off[f + 1] = redo_off[f] + nc * S;
    }
    TRY(h->d_fpeaks.reserve(std::max<size_t>(pk_off[F], 1)));
    TRY(h->d_fcount.reserve(2 * F));
    TRY(h->d_fredo.reserve(std::max<size_t>(redo_off[F], 1)));
    CU(cudaMemsetAsync(h->d_fcount.p, 0, 2 * F * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->d_fredo.p, 0, std::max<size_t>(redo_off[F], 1), h->stream));
    // pageable files go through the pinned ring, pinned / registered ones straight to the copy engine (per file)
    auto stage_threads_of = [&](const void *p) {
        int t = 0;
        if (mem == AM_MEM_HOST && p) {
            cudaPointerAttributes pa;
            if (cudaPointerGetAttributes(&pa, p) == cudaSuccess && pa.type == cudaMemoryTypeUnregistered) t = HostStager::default_threads();
            cudaGetLastError();
        }
        return t;
    };

    // queue every file
    am_stats acc;
    memset(&acc, 0, sizeof acc);
    std::vector<RangePlan> pls(F);
    std::vector<char> queued(F, 0);
    int seg_idx = 0;
    for (size_t f = 0; f < F; ++f) {
        RangePlan &pl = pls[f];
        TRY(plan_range(h, frames[f], fmt, mem == AM_MEM_HOST, scale, 0, (size_t)-1, pl, &nothing));
        if (nothing) continue;
        queued[f] = 1;
        const int stage_threads = stage_threads_of(streams[f]);
        pl.po.peaks = h->d_fpeaks.p + pk_off[f]; pl.po.cap = pk_cap[f]; pl.dev_cap = pk_cap[f];
        pl.po.count = h->d_fcount.p + 2 * f; pl.po.flags = (unsigned *)(h->d_fcount.p + 2 * f + 1);
        pl.po.redo = h->d_fredo.p + redo_off[f];
        for (long long i0 = pl.c_first; i0 < pl.c_last; i0 += pl.K) {
            const long long i1 = std::min(pl.c_last, i0 + pl.K), g0 = pl.C * i0, g1 = pl.seg_out_end(i0, i1);
            if (g1 <= g0) continue;
            amk::StreamView sv;
            sv.fmt = (int)fmt; sv.total = pl.L; sv.lead = 0;
            if (mem == AM_MEM_HOST) {
                TRY(upload_segment(h, pl, streams[f], 0, seg_idx & 1, g0, std::min(pl.L, g1 + pl.m - 1), stage_threads, sv));
            } else {
                sv.x = streams[f]; sv.buf_first = 0; sv.buf_frames = (long long)frames[f];
            }
            TRY(compute_segment(h, pl, sv, i0, i1, pl.summary, false));
            if (mem == AM_MEM_HOST) CU(cudaEventRecord(h->ev_done[seg_idx & 1], h->stream));
            ++seg_idx;
            if (h->progress) h->progress(h->progress_user, 0, (size_t)i0, (size_t)(i1 - i0));
        }
        acc.kernel_launches += h->stats.kernel_launches; acc.fft_blocks += h->stats.fft_blocks;
        acc.h2d_bytes += h->stats.h2d_bytes; acc.frames += frames[f]; acc.chunks += (uint32_t)pl.num_chunks;
        acc.fft_log2 = h->stats.fft_log2; acc.log2_n1 = h->stats.log2_n1; acc.log2_n2 = h->stats.log2_n2;
        acc.summary_mode = std::max(acc.summary_mode, h->stats.summary_mode);
    }

    // one synchronisation, then the per-file host tails
    std::vector<unsigned long long> cnt(2 * F, 0);
    CU(cudaMemcpyAsync(cnt.data(), h->d_fcount.p, 2 * F * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    acc.d2h_bytes += 2 * F * sizeof(unsigned long long);
    prof_collect(h);
    size_t total = 0;
    std::vector<am_peak> all, kept;
    for (size_t f = 0; f < F; ++f) {
        if (!queued[f]) { if (h->progress) h->progress(h->progress_user, 1, 0, 0); continue; }
        const RangePlan &pl = pls[f];
        const unsigned flags = (unsigned)cnt[2 * f + 1];
        unsigned long long count = cnt[2 * f];
        const amp::DevPeak *src = h->d_fpeaks.p + pk_off[f];
        if ((flags & amp::FLAG_OVERFLOW) && count <= pk_cap[f])
            return fail(AM_ERR_CAPACITY, "file %zu: a chunk kept more than max_peaks_per_chunk = %d peaks, or holds more than that many local "
                                         "maxima of exactly equal height (raise it, the prominence or the distance)", f, pl.pk_cap);
        if ((flags & (amp::FLAG_NEED_DENSE | amp::FLAG_OVERFLOW)) || count > pk_cap[f]) {
            // rare: this file through the one-file path (summary pass + dense repeat of the marked chunks, full-size peak list)
            const am_stats keep = h->stats;
            TRY(range_pass_device(h, streams[f], 0, frames[f], frames[f], fmt, mem, scale, 0, (size_t)-1, &count));
            acc.kernel_launches += h->stats.kernel_launches; acc.fft_blocks += h->stats.fft_blocks; acc.h2d_bytes += h->stats.h2d_bytes;
            acc.d2h_bytes += h->stats.d2h_bytes; acc.dense_chunks += h->stats.dense_chunks;
            acc.summary_mode = std::max(acc.summary_mode, h->stats.summary_mode);
            h->stats = keep;
            src = h->d_peaks.p;
        } else if (h->progress) {
            h->progress(h->progress_user, 1, 0, pl.num_chunks);
        }
        all.resize((size_t)count);
        if (count) {
            CU(cudaMemcpy(all.data(), src, (size_t)count * sizeof(am_peak), cudaMemcpyDeviceToHost));
            acc.d2h_bytes += (size_t)count * sizeof(am_peak);
        }
        kept.resize(all.size());
        size_t nk = 0;
        TRY(am_merge_peaks(all.data(), all.size(), h->sr, h->cfg.distance_s, kept.data(), kept.size(), &nk));
        n_out[f] = nk;
        if (total + nk <= cap && nk) {
            if (!out) return fail(AM_ERR_INVALID, "NULL output buffer");
            memcpy(out + total, kept.data(), nk * sizeof(am_peak));
        }
        total += nk;
    }
    h->stats = acc;
    if (total > cap) return fail(AM_ERR_CAPACITY, "%zu peaks, capacity %zu", total, cap);
    return AM_OK;
}


// ---- push session: calc_chunks for a decoder that produces the stream piece by piece -------------------
// The reference feeds calc_chunks a lazy iterator of decoded frames (mp3_reader.rs:13-66, matcher/mod.rs:71-83).
// Here the decoder thread pushes whatever it has; frames collect in the pinned ring (one asynchronous copy per
// 8 MB, or per segment end) and land in one of two device segment buffers.  As soon as the frames of a segment of
// K logical chunks are complete its transforms + peak kernels are launched, so matching overlaps decoding and the
// uploads; only the last, partial segment is left for am_stream_finish.
struct am_stream_session {
    am_matcher *h = nullptr;
    RangePlan pl;
    long long max_frames = 0, pushed = 0;
    long long seg_i0 = 0;            // first logical chunk of the segment being filled
    int b = 0;                       // device staging buffer of that segment
    long long seg_cap = 0;           // frames of a full segment: K C + ov
    int slot = -1;                   // pinned slot being filled, -1 = none
    size_t slot_fill = 0;            // bytes in it
    long long slot_first = 0;        // stream frame of its first byte
    bool have_prev = false;          // the previous launch still awaits its marked-chunk check
    long long prev_i0 = 0, prev_i1 = 0, prev_origin = 0;   // its chunks, and the first chunk held by its buffer
    int prev_b = 0;
    bool failed = false;
};

namespace {
am_status session_flush_slot(am_stream_session *s) {
    am_matcher *h = s->h;
    if (s->slot < 0 || s->slot_fill == 0) return AM_OK;
    const long long seg_first = s->pl.C * s->seg_i0;
    void *dst = h->d_stage[s->b].p + (size_t)(s->slot_first - seg_first) * s->pl.fb;
    CU(h->stager.send(s->slot, dst, s->slot_fill, h->copy_stream));
    h->stats.h2d_bytes += s->slot_fill;
    h->stager.slot = (s->slot + 1) % HostStager::SLOTS;
    s->slot = -1;
    s->slot_fill = 0;
    return AM_OK;
}
// device buffer b holds frames [C origin, frames_end)
amk::StreamView session_view(const am_stream_session *s, int b, long long origin, long long frames_end) {
    amk::StreamView sv;
    sv.fmt = s->pl.fmt; sv.total = s->pl.L; sv.lead = 0;
    sv.x = s->h->d_stage[b].p; sv.buf_first = s->pl.C * origin; sv.buf_frames = frames_end - sv.buf_first;
    return sv;
}
// dense repeat of the chunks of the previous segment that the summary pass marked (its frames are still in prev_b)
am_status session_settle_prev(am_stream_session *s) {
    am_matcher *h = s->h;
    if (!s->have_prev) return AM_OK;
    s->have_prev = false;
    if (!s->pl.summary) return AM_OK;
    unsigned long long cnt[2];
    TRY(fetch_counters(h, cnt));
    if (!((unsigned)cnt[1] & amp::FLAG_NEED_DENSE) || ((unsigned)cnt[1] & amp::FLAG_OVERFLOW)) return AM_OK;
    std::vector<unsigned char> any;
    TRY(marked_chunks(h, s->pl, s->prev_i0 - s->pl.c_first, s->prev_i1 - s->pl.c_first, any));
    const long long frames_end = std::min(s->pl.L, s->pl.C * s->prev_i1 + s->pl.ov);
    for (long long i = 0; i < (long long)any.size();) {
        if (!any[(size_t)i]) { ++i; continue; }
        long long j = i;
        while (j < (long long)any.size() && any[(size_t)j]) ++j;
        h->stats.summary_mode = 2;
        h->stats.dense_chunks += (uint32_t)(j - i);
        TRY(compute_segment(h, s->pl, session_view(s, s->prev_b, s->prev_origin, frames_end), s->prev_i0 + i, s->prev_i0 + j, false, true));
        i = j;
    }
    CU(cudaEventRecord(h->ev_done[s->prev_b], h->stream));
    return AM_OK;
}
// all frames of chunks [i0, i1) (at most K of them, i0 >= seg_i0) are on their way into the current buffer: launch them
am_status session_launch(am_stream_session *s, long long i0, long long i1, long long frames_end) {
    am_matcher *h = s->h;
    TRY(session_flush_slot(s));
    CU(cudaEventRecord(h->ev_up[s->b], h->copy_stream));
    CU(cudaStreamWaitEvent(h->stream, h->ev_up[s->b], 0));
    TRY(session_settle_prev(s));
    TRY(compute_segment(h, s->pl, session_view(s, s->b, s->seg_i0, frames_end), i0, i1, s->pl.summary, false));
    CU(cudaEventRecord(h->ev_done[s->b], h->stream));
    if (h->progress) h->progress(h->progress_user, 0, (size_t)i0, (size_t)(i1 - i0));
    s->have_prev = true; s->prev_i0 = i0; s->prev_i1 = i1; s->prev_origin = s->seg_i0; s->prev_b = s->b;
    return AM_OK;
}
}  // namespace

am_status am_stream_begin(am_matcher *h, size_t max_frames, am_sample_fmt fmt, int scale, am_stream_session **out) {
    if (!h || !out) return fail(AM_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (max_frames == 0) return fail(AM_ERR_INVALID, "max_frames is 0 (pass the claimed stream length, mod.rs:78)");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->in_session) return fail(AM_ERR_INVALID, "a push session is already open on this matcher");
    am_stream_session *s = new (std::nothrow) am_stream_session();
    if (!s) return fail(AM_ERR_NOMEM, "out of host memory");
    s->h = h;
    s->max_frames = (long long)max_frames;
    bool nothing = true;
    am_status st = plan_range(h, max_frames, fmt, true, scale, 0, (size_t)-1, s->pl, &nothing);
    if (st == AM_OK && !nothing) {
        s->seg_cap = s->pl.K * s->pl.C + s->pl.ov;
        for (int b = 0; b < 2 && st == AM_OK; ++b) st = h->d_stage[b].reserve((size_t)std::min(s->seg_cap, s->max_frames) * s->pl.fb);
        if (st == AM_OK && !h->stager.ensure(HostStager::default_threads())) st = fail(AM_ERR_NOMEM, "pinned staging ring");
    } else if (st == AM_OK) {
        // the stream cannot hold a single output (shorter than the snippet): accept the frames, report no peaks
        s->pl.fb = fmt_bytes(fmt);
        s->seg_cap = 0;
    }
    if (st != AM_OK) { delete s; return st; }
    h->in_session = true;
    *out = s;
    return AM_OK;
}

am_status am_stream_push(am_stream_session *s, const void *pcm, size_t frames) {
    if (!s) return fail(AM_ERR_INVALID, "NULL session");
    if (frames == 0) return AM_OK;
    if (!pcm) return fail(AM_ERR_INVALID, "NULL buffer");
    am_matcher *h = s->h;
    std::lock_guard<std::mutex> lk(h->mu);
    if (s->failed) return fail(AM_ERR_INVALID, "the session has failed; call am_stream_abort");
    if (s->pushed + (long long)frames > s->max_frames)
        return fail(AM_ERR_INVALID, "%lld frames pushed, more than the %lld announced to am_stream_begin", s->pushed + (long long)frames, s->max_frames);
    if (s->seg_cap == 0) { s->pushed += (long long)frames; return AM_OK; }
    CU(cudaSetDevice(h->device));
    auto body = [&]() -> am_status {
        const unsigned char *src = (const unsigned char *)pcm;
        long long left = (long long)frames;
        const size_t fb = s->pl.fb;
        while (left > 0) {
            const long long seg_first = s->pl.C * s->seg_i0;
            const long long room = seg_first + s->seg_cap - s->pushed;       // frames until the segment is complete
            if (s->slot < 0) {
                s->slot = h->stager.slot;
                CU(h->stager.acquire(s->slot));
                if (s->pushed == seg_first + (s->seg_i0 > 0 ? s->pl.ov : 0))   // first upload into this buffer: its previous kernels must be done
                    CU(cudaStreamWaitEvent(h->copy_stream, h->ev_done[s->b], 0));
                s->slot_fill = 0;
                s->slot_first = s->pushed;
            }
            const long long slot_room = (long long)((HostStager::fill_bytes() - s->slot_fill) / fb);
            const long long take = std::min(left, std::min(room, slot_room));
            h->stager.parallel_copy((char *)h->stager.pinned[s->slot] + s->slot_fill, src, (size_t)take * fb);
            s->slot_fill += (size_t)take * fb;
            s->pushed += take;
            src += (size_t)take * fb;
            left -= take;
            if (take == slot_room) TRY(session_flush_slot(s));
            if (take == room) {
                // segment complete: launch it, then seed the next buffer with the ov-frame halo (device to device)
                const long long i1 = s->seg_i0 + s->pl.K;
                TRY(session_launch(s, s->seg_i0, i1, s->pushed));
                const int nb = s->b ^ 1;
                CU(cudaStreamWaitEvent(h->copy_stream, h->ev_done[nb], 0));
                if (s->pl.ov > 0)
                    CU(cudaMemcpyAsync(h->d_stage[nb].p, h->d_stage[s->b].p + (size_t)(s->pl.K * s->pl.C) * fb, (size_t)s->pl.ov * fb,
                                       cudaMemcpyDeviceToDevice, h->copy_stream));
                s->seg_i0 = i1;
                s->b = nb;
            }
        }
        return AM_OK;
    };
    const am_status st = body();
    if (st != AM_OK) s->failed = true;
    return st;
}

static void session_close(am_stream_session *s) {
    am_matcher *h = s->h;
    cudaStreamSynchronize(h->copy_stream);
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < HostStager::SLOTS; ++i) h->stager.busy[i] = false;
    h->in_session = false;
    delete s;
}

void am_stream_abort(am_stream_session *s) {
    if (!s) return;
    std::lock_guard<std::mutex> lk(s->h->mu);
    cudaSetDevice(s->h->device);
    session_close(s);
}

am_status am_stream_finish(am_stream_session *s, am_peak *out, size_t cap, size_t *n_out) {
    if (!s || !n_out) return fail(AM_ERR_INVALID, "NULL argument");
    am_matcher *h = s->h;
    std::lock_guard<std::mutex> lk(h->mu);
    *n_out = 0;
    auto body = [&]() -> am_status {
        if (s->failed) return fail(AM_ERR_INVALID, "the session has failed");
        CU(cudaSetDevice(h->device));
        if (s->seg_cap == 0) return AM_OK;
        // the stream's true length is what was pushed (<= the announced maximum): the tail chunks get their real windows
        RangePlan &pl = s->pl;
        pl.L = s->pushed;
        const long long c_last = (long long)am_num_chunks(h, (size_t)s->pushed);
        unsigned long long cnt[2] = {0, 0};
        // (a tail that stops just short of a full segment can span a few chunks more than K: ov frames past the last full window)
        TRY(session_flush_slot(s));
        for (long long i0 = s->seg_i0; i0 < c_last; i0 += pl.K) TRY(session_launch(s, i0, std::min(c_last, i0 + pl.K), s->pushed));
        TRY(session_settle_prev(s));
        TRY(fetch_counters(h, cnt));
        CU(cudaStreamSynchronize(h->copy_stream));
        prof_collect(h);
        h->stats.frames = (uint64_t)s->pushed;
        h->stats.chunks = (uint32_t)c_last;
        unsigned long long count = 0;
        TRY(check_counters(h, pl, cnt, &count));
        std::vector<am_peak> all((size_t)count), kept((size_t)count);
        if (count) {
            CU(cudaMemcpy(all.data(), h->d_peaks.p, (size_t)count * sizeof(am_peak), cudaMemcpyDeviceToHost));
            h->stats.d2h_bytes += (size_t)count * sizeof(am_peak);
        }
        size_t nk = 0;
        TRY(am_merge_peaks(all.data(), all.size(), h->sr, h->cfg.distance_s, kept.data(), kept.size(), &nk));
        *n_out = nk;
        if (nk > cap) return fail(AM_ERR_CAPACITY, "%zu peaks, capacity %zu", nk, cap);
        if (nk) {
            if (!out) return fail(AM_ERR_INVALID, "NULL output buffer");
            memcpy(out, kept.data(), nk * sizeof(am_peak));
        }
        if (h->progress) h->progress(h->progress_user, 1, 0, (size_t)c_last);
        return AM_OK;
    };
    const am_status st = body();
    char msg[sizeof g_err];
    memcpy(msg, g_err, sizeof g_err);
    session_close(s);
    memcpy(g_err, msg, sizeof g_err);
    return st;
}

// Test hook (tests/test_gpu_parity.py): run the per-chunk peak kernels on a correlation supplied by the caller
// instead of one computed from a stream.  `c_host` holds the n outputs of one segment that starts at chunk 0; the
// chunk geometry comes from the matcher's config and snippet length.  summary != 0 exercises the run-record path
// (records built from c, runs below theta poisoned with NaN); peaks come back before the global filter.
am_status am_debug_peaks_from_correlation(am_matcher *h, const float *c_host, size_t n, int summary, am_peak *out,
                                          size_t cap, size_t *n_out, uint32_t *mode_out) {
    if (!h || !c_host || !n_out || n == 0) return fail(AM_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    long long C, ov;
    chunk_params(h, C, ov);
    const long long m = (long long)h->m, L = (long long)n + m - 1;
    if (C <= 0 || C + ov >= (1ll << 31)) return fail(AM_ERR_INVALID, "bad chunk geometry");
    if (summary && (C % 16) != 0) return fail(AM_ERR_INVALID, "summary mode needs a chunk size that is a multiple of 16 samples");
    const long long nchunks = (L + C - 1) / C;
    const long long seg_c_len = (((long long)n + 15) / 16) * 16;
    TRY(h->d_c.reserve((size_t)seg_c_len));
    TRY(h->d_rsum.reserve((size_t)seg_c_len >> 4));
    const amp::RunRecs recs{h->d_rsum.p}, no_recs{nullptr};
    const long long tiles_stride = ((C + std::max<long long>(ov - m + 1, 1)) + amp::TP - 1) / amp::TP + 1;
    TRY(h->d_tmin.reserve((size_t)(nchunks * tiles_stride)));
    TRY(h->d_tmax.reserve((size_t)(nchunks * tiles_stride)));
    const int pk_cap = h->cfg.max_peaks_per_chunk ? (int)h->cfg.max_peaks_per_chunk : 1024;
    if (amp::chunk_peaks_smem(pk_cap, 0) > 200 * 1024) return fail(AM_ERR_INVALID, "max_peaks_per_chunk too large");
    int sm_tiles = 0;
    if (amp::chunk_peaks_smem(pk_cap, (int)tiles_stride) <= 96 * 1024) sm_tiles = (int)tiles_stride;
    const size_t pk_smem = amp::chunk_peaks_smem(pk_cap, sm_tiles);
    TRY(set_smem(amp::k_chunk_peaks<false>, pk_smem));
    TRY(set_smem(amp::k_chunk_peaks<true>, pk_smem));
    const size_t dev_cap = std::min<size_t>((size_t)nchunks * (size_t)pk_cap, (size_t)1 << 22);
    TRY(h->d_peaks.reserve(dev_cap));
    TRY(h->d_redo.reserve((size_t)nchunks));
    amp::PeakOut po;
    po.peaks = h->d_peaks.p; po.cap = dev_cap; po.count = h->d_count.p; po.flags = (unsigned *)(h->d_count.p + 1);
    po.redo = h->d_redo.p; po.redo_first = 0; po.redo_stride = nchunks;
    const unsigned long long min_dist = (unsigned long long)h->cfg.distance_s * (unsigned long long)h->sr;
    const float theta = 0.5f * h->cfg.prominence;
    amp::ChunkGeom cg;
    cg.C = C; cg.ov = ov; cg.m = m; cg.total = L; cg.first_chunk = 0; cg.c_g0 = 0; cg.tiles_stride = (int)tiles_stride;
    cg.c_stride = seg_c_len; cg.seg_end = (long long)n;
    dim3 tgrid3((unsigned)((tiles_stride + 7) / 8), (unsigned)nchunks, 1), pgrid((unsigned)nchunks, 1);
    unsigned long long cnt[2] = {0, 0};
    uint32_t mode = summary ? 1 : 0;
    CU(cudaMemsetAsync(h->d_count.p, 0, 2 * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->d_redo.p, 0, (size_t)nchunks, h->stream));
    for (int pass = 0; pass < 2; ++pass) {
        const bool sum = summary && pass == 0;
        if (pass == 1 && !(summary && ((unsigned)cnt[1] & amp::FLAG_NEED_DENSE))) break;
        if (pass == 1) mode = 2;                             // the marked chunks are repeated on the dense correlation
        CU(cudaMemcpyAsync(h->d_c.p, c_host, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        if (sum) {
            amp::k_debug_make_runs<<<(unsigned)(((seg_c_len >> 4) + 255) / 256), 256, 0, h->stream>>>(h->d_c.p, (long long)n, theta, recs);
            amp::k_tile_from_runs<<<tgrid3, 256, 0, h->stream>>>(recs, cg, h->d_tmin.p, h->d_tmax.p, po);
            amp::k_chunk_peaks<true><<<pgrid, 256, pk_smem, h->stream>>>(h->d_c.p, recs, theta, cg, h->d_tmin.p, h->d_tmax.p,
                                                                       h->cfg.prominence, min_dist, pk_cap, sm_tiles, po, 0);
        } else {
            amp::k_tile_minmax<<<tgrid3, 256, 0, h->stream>>>(h->d_c.p, cg, h->d_tmin.p, h->d_tmax.p);
            amp::k_chunk_peaks<false><<<pgrid, 256, pk_smem, h->stream>>>(h->d_c.p, no_recs, 0.f, cg, h->d_tmin.p, h->d_tmax.p,
                                                                        h->cfg.prominence, min_dist, pk_cap, sm_tiles, po, pass);
        }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(cnt, h->d_count.p, sizeof cnt, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    if (mode_out) *mode_out = mode;
    if ((unsigned)cnt[1] & amp::FLAG_OVERFLOW) return fail(AM_ERR_CAPACITY, "more than max_peaks_per_chunk = %d candidates in a chunk", pk_cap);
    if (cnt[0] > dev_cap) return fail(AM_ERR_CAPACITY, "%llu peaks exceed the device list capacity %zu", cnt[0], dev_cap);
    std::vector<am_peak> all((size_t)cnt[0]);
    if (cnt[0]) CU(cudaMemcpy(all.data(), h->d_peaks.p, (size_t)cnt[0] * sizeof(am_peak), cudaMemcpyDeviceToHost));
    std::stable_sort(all.begin(), all.end(), [](const am_peak &a, const am_peak &b) {
        if (a.chunk != b.chunk) return a.chunk < b.chunk;
        return a.height > b.height || (a.height == b.height && a.start < b.start);
    });
    *n_out = all.size();
    if (all.size() > cap) return fail(AM_ERR_CAPACITY, "%zu peaks, capacity %zu", all.size(), cap);
    if (!all.empty()) memcpy(out, all.data(), all.size() * sizeof(am_peak));
    return AM_OK;
}

// ---- multi-GPU merge over NCCL ------------------------------------------------------------------------
// NCCL is bound at run time: the handful of entry points used here have had the same C signatures through
// NCCL 2.x, and a process that already holds a copy (torch.distributed) must share it rather than get a second one.
namespace {
struct NcclId { char internal[AM_COMM_ID_BYTES]; };
typedef struct ncclComm *nccl_comm_t;
struct NcclApi {
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, NcclId, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    char why[256] = "";
};
NcclApi &nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        const char *names[] = {getenv("AM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        void *lib = nullptr;
        for (const char *n : names) {
            if (!n || !*n) continue;
            // RTLD_NOLOAD first: reuse the copy the process already mapped (e.g. the one torch bundles)
            lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
            if (!lib) lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) { snprintf(a.why, sizeof a.why, "libnccl.so.2 not found (%s)", dlerror()); return a; }
        a.GetUniqueId = (int (*)(NcclId *))dlsym(lib, "ncclGetUniqueId");
        a.CommInitRank = (int (*)(nccl_comm_t *, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
        a.CommDestroy = (int (*)(nccl_comm_t))dlsym(lib, "ncclCommDestroy");
        a.AllGather = (int (*)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t))dlsym(lib, "ncclAllGather");
        a.GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
        if (!a.ok) snprintf(a.why, sizeof a.why, "libnccl.so.2 lacks an expected symbol");
        return a;
    }();
    return api;
}
#define NC(expr)                                                                                           \
    do {                                                                                                   \
        int r_ = (expr);                                                                                   \
        if (r_ != 0) return fail(AM_ERR_CUDA, "%s: %s", #expr, nccl_api().GetErrorString(r_));             \
    } while (0)
}  // namespace

struct am_comm {
    nccl_comm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    size_t record_peaks = 4096;       // peaks per rank in one all-gather record (grows if a rank ever overflows it)
};

am_status am_comm_get_unique_id(void *id_out) {
    if (!id_out) return fail(AM_ERR_INVALID, "NULL argument");
    NcclApi &api = nccl_api();
    if (!api.ok) return fail(AM_ERR_UNSUPPORTED, "NCCL unavailable: %s", api.why);
    NcclId id;
    NC(api.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return AM_OK;
}

am_status am_comm_init(int nranks, int rank, const void *nccl_unique_id, am_comm **out) {
    if (!out) return fail(AM_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks || !nccl_unique_id) return fail(AM_ERR_INVALID, "bad rank %d of %d", rank, nranks);
    NcclApi &api = nccl_api();
    if (!api.ok) return fail(AM_ERR_UNSUPPORTED, "NCCL unavailable: %s", api.why);
    am_comm *c = new (std::nothrow) am_comm();
    if (!c) return fail(AM_ERR_NOMEM, "out of host memory");
    c->nranks = nranks;
    c->rank = rank;
    if (const char *v = getenv("AM_GATHER_RECORD_PEAKS")) {      // tests: a tiny record forces the second all-gather round
        const long long x = atoll(v);
        if (x > 0) c->record_peaks = (size_t)x;
    }
    cudaError_t e = cudaGetDevice(&c->device);
    if (e != cudaSuccess) { delete c; return fail(AM_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e)); }
    NcclId id;
    memcpy(&id, nccl_unique_id, sizeof id);
    int r = api.CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) { delete c; return fail(AM_ERR_CUDA, "ncclCommInitRank: %s", api.GetErrorString(r)); }
    *out = c;
    return AM_OK;
}
void am_comm_destroy(am_comm *comm) {
    if (!comm) return;
    if (comm->comm) nccl_api().CommDestroy(comm->comm);
    delete comm;
}
int am_comm_rank(const am_comm *comm) { return comm ? comm->rank : 0; }
int am_comm_size(const am_comm *comm) { return comm ? comm->nranks : 1; }

am_status am_calc_chunks_sharded(am_matcher *h, am_comm *comm, const void *stream, size_t buf_first_frame,
                                 size_t buf_frames, size_t total_frames, am_sample_fmt fmt, am_mem mem, int scale,
                                 size_t first_chunk, size_t num_chunks, am_peak *out, size_t cap, size_t *n_out) {
    if (!h || !n_out) return fail(AM_ERR_INVALID, "NULL argument");
    if (!comm || comm->nranks == 1)          // one rank: the shard is the stream
        return am_calc_chunks_range(h, stream, buf_first_frame, buf_frames, total_frames, fmt, mem, scale, first_chunk,
                                    num_chunks, 1, out, cap, n_out);
    std::lock_guard<std::mutex> lk(h->mu);
    *n_out = 0;
    if (h->in_session) return fail(AM_ERR_INVALID, "a push session is open on this matcher");
    if (h->device != comm->device) return fail(AM_ERR_INVALID, "matcher lives on device %d, communicator on %d", h->device, comm->device);
    NcclApi &api = nccl_api();
    unsigned long long count = 0;
    // A rank that fails locally still has to take part in the collective: it contributes count = ~0 and all ranks fail.
    const am_status local = range_pass_device(h, stream, buf_first_frame, buf_frames, total_frames, fmt, mem, scale,
                                              first_chunk, num_chunks, &count);
    char local_err[sizeof g_err];
    memcpy(local_err, g_err, sizeof g_err);
    CU(cudaSetDevice(h->device));
    const int R = comm->nranks;
    std::vector<unsigned char> host;
    std::vector<unsigned long long> counts((size_t)R);
    for (int round = 0; round < 2; ++round) {
        const size_t G = comm->record_peaks, rec = 16 + G * sizeof(am_peak);
        TRY(h->d_gsend.reserve(rec));
        TRY(h->d_grecv.reserve(rec * (size_t)R));
        // record = [count u64][status u64][G peaks]; count and peaks come straight from the device buffers
        unsigned long long head[2] = {local == AM_OK ? count : ~0ull, (unsigned long long)local};
        CU(cudaMemcpyAsync(h->d_gsend.p, head, sizeof head, cudaMemcpyHostToDevice, h->stream));
        const size_t n_here = local == AM_OK ? (size_t)std::min<unsigned long long>(count, G) : 0;
        if (n_here) CU(cudaMemcpyAsync(h->d_gsend.p + 16, h->d_peaks.p, n_here * sizeof(am_peak), cudaMemcpyDeviceToDevice, h->stream));
        NC(api.AllGather(h->d_gsend.p, h->d_grecv.p, rec, /* ncclChar */ 0, comm->comm, h->stream));
        host.resize(rec * (size_t)R);
        CU(cudaMemcpyAsync(host.data(), h->d_grecv.p, host.size(), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        h->stats.d2h_bytes += host.size();
        unsigned long long worst = 0;
        for (int r = 0; r < R; ++r) {
            memcpy(&counts[(size_t)r], host.data() + rec * (size_t)r, 8);
            if (counts[(size_t)r] == ~0ull) {
                if (local != AM_OK) { memcpy(g_err, local_err, sizeof g_err); return local; }
                return fail(AM_ERR_CUDA, "rank %d failed in its shard of the sharded calc_chunks", r);
            }
            worst = std::max(worst, counts[(size_t)r]);
        }
        if (worst <= G) break;                       // every rank's peaks fit the record
        comm->record_peaks = (size_t)worst;          // all ranks see the same counts and grow alike; gather once more
        if (round == 1) return fail(AM_ERR_CAPACITY, "peak records grew during the exchange");
    }
    const size_t rec = 16 + comm->record_peaks * sizeof(am_peak);
    size_t total = 0;
    for (int r = 0; r < R; ++r) total += (size_t)counts[(size_t)r];
    std::vector<am_peak> all(total), kept(total);
    size_t o = 0;
    for (int r = 0; r < R; ++r) {
        if (counts[(size_t)r]) memcpy(all.data() + o, host.data() + rec * (size_t)r + 16, (size_t)counts[(size_t)r] * sizeof(am_peak));
        o += (size_t)counts[(size_t)r];
    }
    size_t nk = 0;
    TRY(am_merge_peaks(all.data(), all.size(), h->sr, h->cfg.distance_s, kept.data(), kept.size(), &nk));
    *n_out = nk;
    if (nk > cap) return fail(AM_ERR_CAPACITY, "%zu peaks, capacity %zu", nk, cap);
    if (nk) {
        if (!out) return fail(AM_ERR_INVALID, "NULL output buffer");
        memcpy(out, kept.data(), nk * sizeof(am_peak));
    }
    return AM_OK;
}

am_status am_calc_chunks(am_matcher *h, const void *stream, size_t frames, am_sample_fmt fmt, am_mem mem, int scale,
                         am_peak *out, size_t cap, size_t *n_out) {
    return am_calc_chunks_range(h, stream, 0, frames, frames, fmt, mem, scale, 0, (size_t)-1, 1, out, cap, n_out);
}

am_status am_synth_pcm16_device(uint64_t seed, uint64_t first, size_t count, int16_t *dev_out, void *cuda_stream) {
    if (count == 0) return AM_OK;
    if (!dev_out) return fail(AM_ERR_INVALID, "NULL buffer");
    unsigned grid = (unsigned)std::min<size_t>((count + 255) / 256, 148 * 32);
    k_synth_pcm16<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(seed, first, count, dev_out);
    CU(cudaGetLastError());
    return AM_OK;
}
am_status am_synth_coloured_pcm16_device(uint64_t seed, uint64_t first, size_t count, int taps, int mul, int shift,
                                         int16_t *dev_out, void *cuda_stream) {
    if (count == 0) return AM_OK;
    if (!dev_out || taps < 1 || taps > 4096 || shift < 0 || shift > 30) return fail(AM_ERR_INVALID, "bad argument");
    unsigned grid = (unsigned)std::min<size_t>((count + 255) / 256, 148 * 32);
    k_synth_coloured16<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(seed, first, count, taps, mul, shift, dev_out);
    CU(cudaGetLastError());
    return AM_OK;
}
am_status am_synth_plant_device(int16_t *dev_pcm, size_t frames, int channels, const int16_t *dev_snip, size_t m,
                                uint64_t offset, int shift, void *cuda_stream) {
    if (m == 0) return AM_OK;
    if (!dev_pcm || !dev_snip || (channels != 1 && channels != 2) || shift < 0 || shift > 15)
        return fail(AM_ERR_INVALID, "bad argument");
    k_synth_plant<<<(unsigned)((m + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(dev_pcm, frames, channels, dev_snip, m,
                                                                                      offset, shift);
    CU(cudaGetLastError());
    return AM_OK;
}

}  // extern "C"
