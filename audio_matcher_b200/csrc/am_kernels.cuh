// am_kernels.cuh -- CUDA kernels of the snippet-vs-stream matcher (sm_100a).
//
// Overlap-save FFT cross-correlation.  The stream is cut into blocks of N = 2^k
// frames that advance by V_N = N - m + 1 frames; two real blocks are packed as the
// real and imaginary part of one complex signal (correlation with a real snippet is
// real-linear, so Re/Im of IFFT(FFT(z) conj(S)) are the two blocks' correlations).
//
//   N <= 2^13 : k_small   one pass, whole transform in one thread group
//   N >  2^13 : four-step N = N1 x N2 (element n = n1*N2 + n2, spectrum bin k = k1 + N1*k2)
//       k_col_fwd_stream / k_col_fwd
//                  PCM -> f32 (scale/downmix fused into the load, mp3_reader.rs:12,35),
//                  length-N1 column transforms over a tile of T columns, twiddle, store A[k1][n2];
//                  the _stream version is persistent and gets the next tile's frames by TMA
//       k_row32 / k_row (/ k_row32_stream)
//                  per row k1: length-N2 transform, multiply by the conjugate snippet
//                  spectrum, inverse transform, all in registers; in-place on A (for a batch of
//                  snippets the forward transform runs once and the multiply+inverse per snippet)
//       k_col_inv  conjugate twiddle (carrying the 1/(N sum s^2) scale), inverse column transforms,
//                  crop to the valid outputs, store the correlation or its run summaries
//
// Replaces fftconvolve::fftcorrelate as called at src/matcher/audio_matcher.rs:305 and the
// maths of MyConvolve::correlate (:414-457).
#pragma once
#include "am_fft.cuh"
#include "am_peaks.cuh"

namespace amk {

using amfft::EPT;
using amfft::RegFFT;

enum { FMT_F32_MONO = 0, FMT_I16_MONO = 1, FMT_I16_STEREO = 2 };

// A window of the (virtual) stream resident in device memory.  Virtual frame v maps to
// stream frame v - lead (lead = number of zeros prepended for Mode::Full / Mode::Same);
// frames outside [0, total) read as zero; frames inside must lie in the buffer.
struct StreamView {
    const void *x;
    int fmt;
    long long buf_first;   // stream frame index of x[0]
    long long buf_frames;  // frames held in x
    long long total;       // stream length in frames
    long long lead;
};

// exact int16 -> fp32 on the integer/FMA pipes (I2F.S16 runs on the slow conversion unit)
__device__ __forceinline__ float s16_to_f32(int s) { return __int_as_float(0x4B400000 + s) - 12582912.0f; }   // |s| < 2^22
// The reference's scale / downmix, mp3_reader.rs:12,35: fl(fl(fl(l + r) * 0.5) * (1/65535)), mono: l = r.
// For 16-bit samples l + r and the halving are exact in fp32, so the value is the single rounding
// fl((l + r) * (1/65535) / 2): one integer add, the exact int->float trick and one multiply.
__device__ __forceinline__ float pcm_scale_sum(int l_plus_r) { return __fmul_rn(s16_to_f32(l_plus_r), 0.5f / 65535.0f); }
template <int FMT> __device__ __forceinline__ float load_raw(const void *x, long long f) {
    if (FMT == FMT_F32_MONO) return __ldg((const float *)x + f);
    if (FMT == FMT_I16_MONO) return __fmul_rn(s16_to_f32((int)__ldg((const short *)x + f)), 1.0f / 65535.0f);
    short2 lr = __ldg((const short2 *)x + f);
    return pcm_scale_sum((int)lr.x + (int)lr.y);
}

__device__ __forceinline__ float load_frame(const StreamView &s, long long v) {
    long long f = v - s.lead;
    if (f < 0 || f >= s.total) return 0.f;
    f -= s.buf_first;
    if (f < 0 || f >= s.buf_frames) return 0.f;
    if (s.fmt == FMT_F32_MONO) return load_raw<FMT_F32_MONO>(s.x, f);
    if (s.fmt == FMT_I16_MONO) return load_raw<FMT_I16_MONO>(s.x, f);
    return load_raw<FMT_I16_STEREO>(s.x, f);
}
// true when both blocks of `pair` lie entirely inside the resident buffer (CTA-uniform):
// the loads then need no per-element range checks
__device__ __forceinline__ bool pair_in_range(const StreamView &s, long long v0, long long VN, long long n, bool two) {
    long long lo = v0 - s.lead, hi = v0 + (two ? VN : 0) + n - s.lead;
    long long lim = s.buf_first + s.buf_frames;
    if (lim > s.total) lim = s.total;
    return two && lo >= s.buf_first && hi <= lim;
}

// One launch group of overlap-save blocks.
struct BlockGroup {
    StreamView sv;
    long long g0;       // virtual output offset of block 0
    long long g_end;    // outputs at or beyond g_end are dropped
    long long VN;       // N - m + 1
    int nblocks;        // real blocks (pairs = (nblocks + 1) / 2)
    float *c;           // correlation out, c[g - c_g0]
    long long c_g0;
    float scalar;       // 1/N or 1/(N sum s^2)
    // Summary mode (k_col_inv with 16-column tiles only): for every aligned run of 16 outputs the kernel writes
    // rsum.rec[(g - c_g0) >> 4] = {min, max, first, last} and stores the 16 values themselves only when max >= theta.
    // theta = -inf keeps the correlation dense.  rec == nullptr: no records, plain dense stores.
    amp::RunRecs rsum;
    float theta;
};

template <bool SCALE = true>
__device__ __forceinline__ void store_pair(const BlockGroup &g, int pair, long long n, float2 val) {
    if (n >= g.VN) return;
    const float sc = SCALE ? g.scalar : 1.0f;
    long long o = g.g0 + (long long)(2 * pair) * g.VN + n;
    if (o < g.g_end) g.c[o - g.c_g0] = val.x * sc;
    if (2 * pair + 1 < g.nblocks) {
        o += g.VN;
        if (o < g.g_end) g.c[o - g.c_g0] = val.y * sc;
    }
}
__device__ __forceinline__ float2 load_pair(const BlockGroup &g, int pair, long long n) {
    long long v = g.g0 + (long long)(2 * pair) * g.VN + n;
    float re = load_frame(g.sv, v);
    float im = (2 * pair + 1 < g.nblocks) ? load_frame(g.sv, v + g.VN) : 0.f;
    return make_float2(re, im);
}

// ---- single-pass path -------------------------------------------------------------
template <int LOG2N> struct SmallCfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int GT = N / EPT;                       // threads per transform
    static constexpr int THREADS = GT < 128 ? 128 : GT;
    static constexpr int G = THREADS / GT;                   // transforms per CTA
    static constexpr size_t SMEM = (size_t)G * RegFFT<LOG2N, 0, false>::SMEM_ELEMS * sizeof(float2);
};

// one thread group per block pair: load, transform, multiply by the snippet spectrum, inverse, crop
template <int LOG2N>
__global__ void __launch_bounds__(SmallCfg<LOG2N>::THREADS)
k_small(BlockGroup g, const float2 *__restrict__ spec, const float2 *__restrict__ tw) {
    typedef RegFFT<LOG2N, 0, false> F;
    typedef RegFFT<LOG2N, 0, true> I;
    typedef SmallCfg<LOG2N> Cfg;
    extern __shared__ float2 sm_all[];
    const int grp = threadIdx.x / Cfg::GT, gtid = threadIdx.x % Cfg::GT;
    float2 *sm = sm_all + (size_t)grp * F::SMEM_ELEMS;
    const int npairs = (g.nblocks + 1) / 2;
    int pair = blockIdx.x * Cfg::G + grp;
    const bool active = pair < npairs;
    if (!active) pair = npairs - 1;                          // keep every thread in the barriers

    float2 v[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        int idx, t;
        F::template in_coord<0>(gtid, j, idx, t);
        v[j] = load_pair(g, pair, idx);
    }
    F::run(v, sm, gtid, tw);
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        int idx, t;
        F::out_coord(gtid, j, idx, t);
        v[j] = amfft::cmul(v[j], __ldg(&spec[idx]));
    }
    __syncthreads();                                         // exchange buffer is reused by the inverse
    I::run(v, sm, gtid, tw);
    if (active) {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            int idx, t;
            I::out_coord(gtid, j, idx, t);
            store_pair(g, pair, idx, v[j]);
        }
    }
}

// ---- four-step path ---------------------------------------------------------------
// default tile width (log2 columns): N1*T >= 2048 elements and <= 64 KB of exchange buffer
constexpr int col_default_lt(int l1) { return l1 >= 10 ? 13 - l1 : (l1 >= 7 ? 4 : 11 - l1); }
template <int L1, int LT_ = col_default_lt(L1), int E_ = 16> struct ColCfg {
    static constexpr int E = E_;
    static constexpr int LT = LT_;
    static constexpr int T = 1 << LT;
    static constexpr int THREADS = ((1 << L1) * T) / E;
    static constexpr int MINB = THREADS >= 1024 ? 1 : (1024 * 16 / E) / THREADS;   // <= 64 (E = 16) / 128 (E = 32) registers
    // the inverse kernel fits 80 registers without spilling: a third resident CTA hides more load latency
    static constexpr int MINB_INV = (E == 32 && THREADS == 256) ? 3 : MINB;
    static constexpr int MINB_FWD = MINB;        // the forward kernel spills at 80 registers (measured slower)
    static constexpr size_t SMEM = (size_t)RegFFT<L1, LT, false, E>::SMEM_ELEMS * sizeof(float2);
    static constexpr size_t SMEM_INV = SMEM;     // the transposing epilogue reuses the exchange buffer (16 float2 per row, XOR swizzle)
};

// exp(-+ 2 pi i p / N), p < N <= 2^24 (p and 2/N exact in fp32)
__device__ __forceinline__ float2 twiddle_big(unsigned p, float two_over_n, bool inverse) {
    float s, c;
    sincospif((float)p * two_over_n, &s, &c);
    return make_float2(c, inverse ? s : -s);
}

// v[s] *= base * step^s, s < R (product tree of depth log2 R)
template <int R> __device__ __forceinline__ void twiddle_geo(float2 *v, float2 base, float2 step) {
    using amfft::cmul;
    float2 w[R];
    w[0] = base;
    if (R >= 2) w[1] = cmul(base, step);
    float2 sp = step;
#pragma unroll
    for (int h = 2; h < R; h <<= 1) {
        sp = cmul(sp, sp);                         // step^h
#pragma unroll
        for (int i = 0; i < h; ++i) w[h + i] = cmul(w[i], sp);
    }
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = cmul(v[i], w[i]);
}

// Four-step twiddles of one thread: column n2, rows k1 = q0 + l dq + s NB dq (l < NB butterflies, s < R), i.e.
// W_N^{n2 k1} = base[l] step^s with base[l] = W^{n2 q0} z^l, step = z^NB, z = W^{n2 dq}: two sincospif per thread,
// the rest are products.
template <int NB> __device__ __forceinline__ void fourstep_bases(float2 (&base)[NB], float2 &step, unsigned n2, unsigned q0,
                                                                unsigned dq, float two_over_n, bool inverse) {
    using amfft::cmul;
    base[0] = twiddle_big(n2 * q0, two_over_n, inverse);
    const float2 z = twiddle_big(n2 * dq, two_over_n, inverse);
    step = z;
#pragma unroll
    for (int l = 1; l < NB; ++l) base[l] = cmul(base[l - 1], z);
#pragma unroll
    for (int h = 1; h < NB; h <<= 1) step = cmul(step, step);
}

template <int FMT, class F, int LT>
__device__ __forceinline__ void load_tile_fast(float2 (&v)[F::EPT], const BlockGroup &g, long long v0, int log2n2, int n2_0, int tid) {
    const long long f0 = v0 - g.sv.lead - g.sv.buf_first;
#pragma unroll
    for (int j = 0; j < F::EPT; ++j) {
        int idx, t;
        F::template in_coord<0>(tid, j, idx, t);
        long long f = f0 + ((long long)idx << log2n2) + n2_0 + t;
        v[j] = make_float2(load_raw<FMT>(g.sv.x, f), load_raw<FMT>(g.sv.x, f + g.VN));
    }
}


// Mono int16 tile through shared memory with 128-bit global loads: the tile's rows are 16 samples
// (32 bytes) of each block; every (row, half-row) unit is fetched with one LDG.128 per block, the two
// blocks are interleaved as (re, im) int16 pairs in one 32-bit word per element, and the threads then
// pick their stage-0 inputs with conflict-free LDS.32.  Needs 16-byte aligned rows (block advance V_N
// and segment starts multiples of 8 frames), which the host arranges; anything else takes the scalar path.
template <class F, int LT, int L1, int THREADS>
__device__ __forceinline__ void load_tile_i16_staged(float2 (&v)[F::EPT], const BlockGroup &g, long long f0, int log2n2,
                                                      int tid, unsigned *sraw) {
    static_assert(LT == 4, "16-column tiles");
    const short *x = (const short *)g.sv.x + f0;                // f0 already includes the tile's column offset
#pragma unroll
    for (int u = tid; u < (2 << L1); u += THREADS) {
        const int n1 = u >> 1, half = u & 1;
        const short *p = x + ((long long)n1 << log2n2) + half * 8;
        const uint4 re = __ldg((const uint4 *)p);
        const uint4 im = __ldg((const uint4 *)(p + g.VN));
        uint4 a, b;
        a.x = __byte_perm(re.x, im.x, 0x5410); a.y = __byte_perm(re.x, im.x, 0x7632);
        a.z = __byte_perm(re.y, im.y, 0x5410); a.w = __byte_perm(re.y, im.y, 0x7632);
        b.x = __byte_perm(re.z, im.z, 0x5410); b.y = __byte_perm(re.z, im.z, 0x7632);
        b.z = __byte_perm(re.w, im.w, 0x5410); b.w = __byte_perm(re.w, im.w, 0x7632);
        uint4 *dst = (uint4 *)(sraw + n1 * 16 + half * 8);
        dst[0] = a;
        dst[1] = b;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < F::EPT; ++j) {
        int idx, t;
        F::template in_coord<0>(tid, j, idx, t);
        const unsigned w = sraw[idx * 16 + t];
        const float re = __fmul_rn(s16_to_f32((int)(short)(w & 0xffffu)), 1.0f / 65535.0f);
        const float im = __fmul_rn(s16_to_f32((int)w >> 16), 1.0f / 65535.0f);
        v[j] = make_float2(re, im);
    }
    __syncthreads();                                            // the buffer becomes the exchange buffer
}

// column transform of one tile held in registers, four-step twiddle, store to A[pair][k1][n2]
template <int L1, int LT, int E, bool TAIL_SYNC = true>
__device__ __forceinline__ void col_fwd_finish(float2 (&v)[E], float2 *sm_all, int tid, const float2 *__restrict__ tw,
                                               int log2n2, int n2_0, float2 *__restrict__ Ap) {
    typedef RegFFT<L1, LT, false, E> F;
    F::template run<0, TAIL_SYNC>(v, sm_all, tid, tw);
    AM_TL(2);
    const float two_over_n = 2.0f / (float)(1u << (L1 + log2n2));
    constexpr int RB = F::bits_at(F::NST - 1), R = 1 << RB, NB = E / R;
    const int t = tid & ((1 << LT) - 1), q = tid >> LT;      // q < N1 / E
    const unsigned n2 = n2_0 + t;
    static_assert((F::GT >> LT) * NB == ((1 << L1) >> RB), "k1 = q + l N1/E + s N1/R");
    float2 base[NB], step;                                    // k1 = q_l + s (N1 / R): W_N^{n2 k1} = W_N^{n2 q_l} (W_N^{n2 N1 / R})^s
    fourstep_bases<NB>(base, step, n2, (unsigned)q, (unsigned)(F::GT >> LT), two_over_n, false);
#pragma unroll
    for (int l = 0; l < NB; ++l) twiddle_geo<R>(&v[l * R], base[l], step);
    AM_TL(3);
#pragma unroll
    for (int l = 0; l < NB; ++l)
#pragma unroll
        for (int s = 0; s < R; ++s)
            Ap[((size_t)(q + l * (F::GT >> LT) + s * ((1 << L1) >> RB)) << log2n2) + n2] = v[l * R + s];
    AM_TL(4);
}

// grid (N2 / T, pairs).  A[pair][k1][n2] = W_N^{n2 k1} * sum_{n1} z[n1 N2 + n2] W_N1^{n1 k1}
template <int L1, int LT, int E>
__global__ void __launch_bounds__(ColCfg<L1, LT, E>::THREADS, ColCfg<L1, LT, E>::MINB_FWD)
k_col_fwd(BlockGroup g, int log2n2, float2 *__restrict__ A, const float2 *__restrict__ tw) {
    typedef ColCfg<L1, LT, E> Cfg;
    typedef RegFFT<L1, Cfg::LT, false, E> F;
    constexpr int EPT = E;
    extern __shared__ float2 sm_all[];
    const int tid = threadIdx.x, pair = blockIdx.y;
    const int n2_0 = blockIdx.x << Cfg::LT;
    float2 v[EPT];
    AM_TL(0);
    const long long v0 = g.g0 + (long long)(2 * pair) * g.VN;
    if (pair_in_range(g.sv, v0, g.VN, 1ll << (L1 + log2n2), 2 * pair + 1 < g.nblocks)) {
        if (g.sv.fmt == FMT_I16_MONO) {
            const long long f0 = v0 - g.sv.lead - g.sv.buf_first + n2_0;
            bool staged = false;
            if constexpr (Cfg::LT == 4 && L1 >= 5)
                if (((f0 | g.VN) & 7) == 0 && (((size_t)g.sv.x) & 15) == 0) {
                    load_tile_i16_staged<F, Cfg::LT, L1, Cfg::THREADS>(v, g, f0, log2n2, tid, (unsigned *)sm_all);
                    staged = true;
                }
            if (!staged) load_tile_fast<FMT_I16_MONO, F, Cfg::LT>(v, g, v0, log2n2, n2_0, tid);
        } else if (g.sv.fmt == FMT_I16_STEREO) load_tile_fast<FMT_I16_STEREO, F, Cfg::LT>(v, g, v0, log2n2, n2_0, tid);
        else load_tile_fast<FMT_F32_MONO, F, Cfg::LT>(v, g, v0, log2n2, n2_0, tid);
    } else {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            int idx, t;
            F::template in_coord<0>(tid, j, idx, t);
            v[j] = load_pair(g, pair, ((long long)idx << log2n2) + n2_0 + t);
        }
    }
    AM_TL(1);
    col_fwd_finish<L1, Cfg::LT, E>(v, sm_all, tid, tw, log2n2, n2_0, A + ((size_t)pair << (L1 + log2n2)));
}

// Persistent variant of k_col_fwd: grid = CTAs resident on the GPU, tiles (tile = pair * (N2 / T) + column tile) are
// handed out dynamically.  The raw frames of the NEXT tile (2 blocks x N1 rows x T frames) are fetched by TMA
// (cp.async.bulk.tensor.2d, boxes of T columns x 256 rows) into their own shared-memory buffer while the current
// tile is transformed.  Why TMA: a tile is 1024 scattered
// 32-byte pieces; as LDG / cp.async they occupy the SM's load-miss tracking (measured: 2 CTAs/SM reach 2.3 / 4.4 TB/s
// on this pattern and the issuing warps block meanwhile, 58 % of a CTA's life in k_col_fwd), while one thread hands
// TMA the four boxes and every warp goes on to the transform.
// The tensor map views the PCM buffer as rows of N2 frames with a row pitch of N2 frames but 2 N2 columns (rows
// overlap), so that a tile starting at frame f = r N2 + c (c < N2) is the box at (c, r) even when c + 16 > N2.
// Tiles that are not fully resident take the checked scalar loads.
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned smem_dst, const void *tmap, unsigned bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
struct alignas(64) TensorMap { unsigned long long opaque[16]; };     // CUtensorMap (128 bytes), encoded by the host
template <int FMT> struct RawFrame { typedef short type; };                       // FMT_I16_MONO
template <> struct RawFrame<FMT_I16_STEREO> { typedef short2 type; };
template <> struct RawFrame<FMT_F32_MONO> { typedef float type; };
template <int FMT> __device__ __forceinline__ float frame_to_f32(typename RawFrame<FMT>::type x);
template <> __device__ __forceinline__ float frame_to_f32<FMT_I16_MONO>(short x) { return __fmul_rn(s16_to_f32((int)x), 1.0f / 65535.0f); }
template <> __device__ __forceinline__ float frame_to_f32<FMT_I16_STEREO>(short2 x) { return pcm_scale_sum((int)x.x + (int)x.y); }
template <> __device__ __forceinline__ float frame_to_f32<FMT_F32_MONO>(float x) { return x; }

template <int L1, int LT, int E, int FMT> struct ColStreamCfg {
    typedef ColCfg<L1, LT, E> Cfg;
    typedef typename RawFrame<FMT>::type Frame;
    static constexpr int THREADS = Cfg::THREADS;
    static constexpr int T = 1 << LT;
    static constexpr int BOX_ROWS = (1 << L1) < 256 ? (1 << L1) : 256;
    static constexpr size_t XCHG = Cfg::SMEM;                          // exchange buffer
    static constexpr size_t RAW = (size_t)2 * (1 << L1) * T * sizeof(Frame);
    static constexpr size_t SMEM = XCHG + RAW;
    // TMA wants 16-byte box rows and a 128-byte aligned destination; two CTAs per SM must still fit
    static constexpr bool OK = XCHG % 128 == 0 && (T * sizeof(Frame)) % 16 == 0 && SMEM <= 110 * 1024;
};
template <int L1, int LT, int E, int FMT, int L2C = -1>
__global__ void __launch_bounds__(ColStreamCfg<L1, LT, E, FMT>::THREADS, ColCfg<L1, LT, E>::MINB_FWD)
k_col_fwd_stream(const __grid_constant__ TensorMap tm, BlockGroup g, int log2n2_arg, float2 *__restrict__ A,
                 const float2 *__restrict__ tw, int ntiles, int *__restrict__ next_tile) {
    const int log2n2 = L2C >= 0 ? L2C : log2n2_arg;
    typedef ColStreamCfg<L1, LT, E, FMT> SC;
    typedef typename SC::Frame Frame;
    typedef RegFFT<L1, LT, false, E> F;
    constexpr int N1 = 1 << L1, T = SC::T;
    extern __shared__ __align__(128) float2 sm_all[];
    __shared__ __align__(8) unsigned long long bar_store;
    __shared__ int s_next;
    Frame *raw = (Frame *)((char *)sm_all + SC::XCHG);
    const unsigned raw_a = (unsigned)__cvta_generic_to_shared(raw), bar = (unsigned)__cvta_generic_to_shared(&bar_store);
    const int tid = threadIdx.x;
    const int ltx = log2n2 - LT;                                       // log2 of column tiles per pair
    // frame offset of the tile in the buffer, or -1 when the tile needs the checked loads (CTA-uniform)
    // (a last pair that holds a single block is fetched as one block; its imaginary half is zero)
    auto tile_f0 = [&](int tile) -> long long {
        const int pair = tile >> ltx, n2_0 = (tile & ((1 << ltx) - 1)) << LT;
        const long long v0 = g.g0 + (long long)(2 * pair) * g.VN;
        const bool two = 2 * pair + 1 < g.nblocks;
        const long long lo = v0 - g.sv.lead, hi = lo + (two ? g.VN : 0) + (1ll << (L1 + log2n2));
        long long lim = g.sv.buf_first + g.sv.buf_frames;
        if (lim > g.sv.total) lim = g.sv.total;
        if (lo < g.sv.buf_first || hi > lim) return -1;
        return lo - g.sv.buf_first + n2_0;
    };
    auto prefetch = [&](int tile) {                                    // thread 0 only
        const long long f0 = tile_f0(tile);
        if (f0 < 0) return;
        const int nblk = 2 * (tile >> ltx) + 1 < g.nblocks ? 2 : 1;
        mbar_expect_tx(bar, (unsigned)(SC::RAW / 2) * nblk);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            if (blk == nblk) break;
            const long long fb = f0 + blk * g.VN;
            const int c = (int)(fb & ((1ll << log2n2) - 1)), r = (int)(fb >> log2n2);
#pragma unroll
            for (int h = 0; h < N1 / SC::BOX_ROWS; ++h)
                tma_load_2d(raw_a + (unsigned)((blk * N1 + h * SC::BOX_ROWS) * T * sizeof(Frame)), &tm, bar, c, r + h * SC::BOX_ROWS);
        }
    };
    // Tiles are handed out dynamically (*next_tile counts the tiles claimed beyond the first gridDim.x): the two CTAs of an SM do not progress at
    // the same rate (measured 10 k vs 21 k cycles per tile), a static stride would leave the faster one idle.
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    int tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) prefetch(tile);
    unsigned parity = 0;
    while (tile < ntiles) {
        const int pair = tile >> ltx, n2_0 = (tile & ((1 << ltx) - 1)) << LT;
        float2 v[E];
        AM_TL_SET(tile);
        AM_TL(0);
        int claimed = 0;
        if (tid == 0) claimed = (int)gridDim.x + atomicAdd(next_tile, 1);    // in flight while the tile is unpacked
        if (tile_f0(tile) >= 0) {
            mbar_wait(bar, parity);
            parity ^= 1;
            AM_TL(5);
#pragma unroll
            for (int j = 0; j < E; ++j) {
                int idx, t;
                F::template in_coord<0>(tid, j, idx, t);
                v[j] = make_float2(frame_to_f32<FMT>(raw[idx * T + t]), frame_to_f32<FMT>(raw[(N1 + idx) * T + t]));
            }
            if (2 * pair + 1 >= g.nblocks) {                       // single block: the second half of the buffer is stale
#pragma unroll
                for (int j = 0; j < E; ++j) v[j].y = 0.f;
            }
        } else {
            AM_TL(5);
#pragma unroll
            for (int j = 0; j < E; ++j) {
                int idx, t;
                F::template in_coord<0>(tid, j, idx, t);
                v[j] = load_pair(g, pair, ((long long)idx << log2n2) + n2_0 + t);
            }
        }
        if (tid == 0) s_next = claimed;
        __syncthreads();                       // raw tile consumed by everyone; previous tile's exchange reads done
        const int next = s_next;
        if (tid == 0 && next < ntiles) prefetch(next);
        AM_TL(1);
        // (the barrier above separates this tile's exchange writes from the previous tile's last exchange reads)
        col_fwd_finish<L1, LT, E, false>(v, sm_all, tid, tw, log2n2, n2_0, A + ((size_t)pair << (L1 + log2n2)));
        tile = next;
    }
}

// grid (N2 / T, pairs).  y[n1 N2 + n2] = sum_{k1} W_N1^{-n1 k1} W_N^{-n2 k1} B[k1][n2]
// L2C >= 0: log2 N2 known at compile time (the hot shapes; folds the tile addressing into immediates)
template <int L1, int LT, int E, int L2C = -1>
__global__ void __launch_bounds__(ColCfg<L1, LT, E>::THREADS, ColCfg<L1, LT, E>::MINB_INV)
k_col_inv(BlockGroup g, int log2n2_arg, const float2 *__restrict__ A, const float2 *__restrict__ tw) {
    typedef ColCfg<L1, LT, E> Cfg;
    typedef RegFFT<L1, Cfg::LT, true, E> I;
    constexpr int EPT = E;
    const int log2n2 = L2C >= 0 ? L2C : log2n2_arg;
    extern __shared__ float2 sm_all[];
    const int tid = threadIdx.x, pair = blockIdx.y;
    const int n2_0 = blockIdx.x << Cfg::LT;
    const float two_over_n = 2.0f / (float)(1u << (L1 + log2n2));
    const float2 *Ap = A + ((size_t)pair << (L1 + log2n2));
    float2 v[EPT];
    constexpr int RB = I::bits_at(0), R = 1 << RB, NB = EPT / R;
    // conjugate four-step twiddles W_N^{-n2 k1}, k1 = q_l + r N1 / R: per butterfly a geometric sequence base_l step^r
    // that rides on the first butterflies (butterfly0_geo); the column n2 = n2_0 + t is the same for every l
    static_assert(I::GT % Cfg::T == 0, "tile columns divide the thread count");
    static_assert((I::GT >> Cfg::LT) * NB == ((1 << L1) >> RB), "k1 = q + l N1/E + r N1/R");
    float2 base[NB], step;
    const unsigned n2 = n2_0 + (tid & (Cfg::T - 1));
#pragma unroll
    for (int l = 0; l < NB; ++l) {
        const int q = (tid + l * I::GT) >> Cfg::LT;
#pragma unroll
        for (int r = 0; r < R; ++r) v[l * R + r] = Ap[((size_t)(q + r * ((1 << L1) >> RB)) << log2n2) + n2];
    }
    fourstep_bases<NB>(base, step, n2, (unsigned)(tid >> Cfg::LT), (unsigned)(I::GT >> Cfg::LT), two_over_n, true);
    base[0].x *= g.scalar;                                          // the output scale rides on the twiddles
    base[0].y *= g.scalar;
#pragma unroll
    for (int l = 1; l < NB; ++l) { base[l].x *= g.scalar; base[l].y *= g.scalar; }
    I::butterfly0_geo(v, base, step);
    I::template run<0, true, false, true>(v, sm_all, tid, tw);
    const long long o0 = g.g0 + (long long)(2 * pair) * g.VN;
    if constexpr (Cfg::LT == 4) {
        if (g.rsum.rec != nullptr) {
            // Summary epilogue: transpose the tile through the (now idle) exchange buffer so that every thread owns
            // whole rows = aligned runs of 16 outputs of both blocks, then write one {min, max, first, last} record
            // per run and the run itself (four 128-bit stores) only if its maximum reaches theta.
            // rows of 16 float2 = eight 16-byte chunks; chunk c of row r lives at position c ^ (r & 7), which keeps
            // both the column-wise float2 writes and the row-wise float4 reads free of bank conflicts
            const int vn = (int)g.VN;
            const bool has_im = 2 * pair + 1 < g.nblocks;
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                int n1, t;
                I::out_coord(tid, j, n1, t);
                sm_all[n1 * 16 + ((((t >> 1) ^ (n1 & 7)) << 1) | (t & 1))] = v[j];
            }
            __syncthreads();
            for (int row = tid; row < (1 << L1); row += Cfg::THREADS) {
                const int nb = (row << log2n2) + n2_0;
                if (nb >= vn) continue;                              // cropped: V_N and nb are multiples of 16
                float re[16], im[16];
                const float4 *src = (const float4 *)(sm_all + row * 16);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 q = src[i ^ (row & 7)];
                    re[2 * i] = q.x; im[2 * i] = q.y; re[2 * i + 1] = q.z; im[2 * i + 1] = q.w;
                }
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    if (blk == 1 && !has_im) break;
                    const float *vals = blk ? im : re;
                    const long long o = o0 + (long long)blk * vn + nb;
                    const long long left = g.g_end - o;
                    if (left <= 0) continue;
                    const int valid = left < 16 ? (int)left : 16;
                    float mn = vals[0], mx = vals[0], last = vals[0];
                    if (valid == 16) {
#pragma unroll
                        for (int i = 1; i < 16; ++i) { mn = fminf(mn, vals[i]); mx = fmaxf(mx, vals[i]); }
                        last = vals[15];
                    } else {
#pragma unroll
                        for (int i = 1; i < 16; ++i)
                            if (i < valid) { mn = fminf(mn, vals[i]); mx = fmaxf(mx, vals[i]); last = vals[i]; }
                    }
                    const long long ci = o - g.c_g0;
                    g.rsum.rec[ci >> 4] = make_float4(mn, mx, vals[0], last);
                    if (mx >= g.theta) {
                        float *dst = g.c + ci;
                        if (valid == 16) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                ((float4 *)dst)[i] = make_float4(vals[4 * i], vals[4 * i + 1], vals[4 * i + 2], vals[4 * i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (i < valid) dst[i] = vals[i];
                        }
                    }
                }
            }
            return;
        }
    }
    if (2 * pair + 1 < g.nblocks && o0 + 2 * g.VN <= g.g_end) {
        // both blocks entirely inside the requested output range (CTA-uniform): 32-bit crop test only
        float *c0 = g.c + (o0 - g.c_g0);
        const int vn = (int)g.VN;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            int n1, t;
            I::out_coord(tid, j, n1, t);
            const int n = (n1 << log2n2) + n2_0 + t;
            if (n < vn) {
                c0[n] = v[j].x;
                c0[n + vn] = v[j].y;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            int n1, t;
            I::out_coord(tid, j, n1, t);
            store_pair<false>(g, pair, ((long long)n1 << log2n2) + n2_0 + t, v[j]);
        }
    }
}

template <int L2> struct RowCfg {
    static constexpr int N = 1 << L2;
    static constexpr int GT = N / EPT;
    static constexpr int THREADS = GT < 128 ? 128 : GT;
    static constexpr int G = THREADS / GT;
    static constexpr int MINB = THREADS >= 1024 ? 1 : (THREADS >= 512 ? 2 : 1024 / THREADS);
    static constexpr size_t SMEM = (size_t)G * RegFFT<L2, 0, false>::SMEM_ELEMS * sizeof(float2);
};

// rows = pairs * N1, row r of A holds A[pair][k1][0..N2).
//   ROW_FUSED    A[row] <- IFFT(FFT(A[row]) * spec[k1])            one snippet, in place
//   ROW_FORWARD  A[row] <- FFT(A[row])                             many snippets: shared stream spectrum
//   ROW_INVERSE  Bout[row] <- IFFT(A[row] * spec[k1])              ... then once per snippet
enum { ROW_FUSED = 0, ROW_FORWARD = 2, ROW_INVERSE = 3 };
template <int L2, int MODE>
__global__ void __launch_bounds__(RowCfg<L2>::THREADS, RowCfg<L2>::MINB)
k_row(float2 *__restrict__ A, const float2 *__restrict__ spec, float2 *__restrict__ Bout, int log2n1, int rows,
      const float2 *__restrict__ tw) {
    typedef RegFFT<L2, 0, false> F;
    typedef RegFFT<L2, 0, true> I;
    typedef RowCfg<L2> Cfg;
    extern __shared__ float2 sm_all[];
    const int grp = threadIdx.x / Cfg::GT, gtid = threadIdx.x % Cfg::GT;
    float2 *sm = sm_all + (size_t)grp * F::SMEM_ELEMS;
    int row = blockIdx.x * Cfg::G + grp;
    const bool active = row < rows;
    if (!active) row = rows - 1;
    float2 *Ar = A + ((size_t)row << L2);
    float2 v[EPT];
    if constexpr (MODE != ROW_INVERSE) {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            int idx, t;
            F::template in_coord<0>(gtid, j, idx, t);
            v[j] = Ar[idx];
        }
        F::run(v, sm, gtid, tw);
        if constexpr (MODE == ROW_FORWARD) {
            if (active) {
#pragma unroll
                for (int j = 0; j < EPT; ++j) {
                    int idx, t;
                    F::out_coord(gtid, j, idx, t);
                    Ar[idx] = v[j];
                }
            }
            return;
        }
    }
    const int k1 = row & ((1 << log2n1) - 1);
    const float2 *Sr = spec + ((size_t)k1 << L2);
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        int idx, t;
        I::template in_coord<0>(gtid, j, idx, t);            // == F::out_coord: no exchange across the multiply
        if constexpr (MODE == ROW_INVERSE) v[j] = Ar[idx];
        v[j] = amfft::cmul(v[j], __ldg(&Sr[idx]));
    }
    if constexpr (MODE == ROW_FUSED) __syncthreads();        // exchange buffer is reused by the inverse
    I::run(v, sm, gtid, tw);
    float2 *Or = (MODE == ROW_INVERSE) ? Bout + ((size_t)row << L2) : Ar;
    if (active) {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            int idx, t;
            I::out_coord(gtid, j, idx, t);
            Or[idx] = v[j];
        }
    }
}



// Fused row kernel with 32 elements per thread: radix 32-16-16 for 8192 points, i.e. TWO exchanges per
// transform instead of three.  k_row is bound by the shared-memory pipe (768 KB of exchange traffic per
// row at 128 B/clk) and FP issue, so a third less exchange traffic and one twiddled stage less is what
// this variant buys; it pays with 128 registers per thread (256 threads per row, 2 CTAs per SM).
template <int L2> struct Row32Cfg {
    static constexpr int THREADS = (1 << L2) / 32;
    static constexpr int MINB = THREADS >= 512 ? 1 : 512 / THREADS;      // 128 registers per thread
    static constexpr size_t SMEM = (size_t)RegFFT<L2, 0, false, 32>::SMEM_ELEMS * sizeof(float2);
};
template <int L2, int MODE>
__global__ void __launch_bounds__(Row32Cfg<L2>::THREADS, Row32Cfg<L2>::MINB)
k_row32(float2 *__restrict__ A, const float2 *__restrict__ spec, float2 *__restrict__ Bout, int log2n1, int rows,
        const float2 *__restrict__ tw) {
    typedef RegFFT<L2, 0, false, 32> F;
    typedef RegFFT<L2, 0, true, 32> I;
    extern __shared__ float2 sm[];
    const int gtid = threadIdx.x;
    // k1-major order: consecutive CTAs take row k1 of every block pair of the group, so a spectrum row is fetched
    // from DRAM once per launch group instead of once per pair (it does not survive 64 MB of row traffic in L2)
    const int np = rows >> log2n1;
    const int row = (int)(((blockIdx.x % np) << log2n1) + blockIdx.x / np);
    float2 *Ar = A + ((size_t)row << L2);
    float2 v[32];
    if constexpr (MODE != ROW_INVERSE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int idx, t;
            F::template in_coord<0>(gtid, j, idx, t);
            v[j] = Ar[idx];
        }
        AM_TL(0);
        AM_TL_WAIT(v, 32);
        AM_TL(1);
        if constexpr (MODE == ROW_FUSED) F::template run<0, false>(v, sm, gtid, tw);   // the inverse's lead barrier covers it
        else F::run(v, sm, gtid, tw);
        AM_TL(2);
        if constexpr (MODE == ROW_FORWARD) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int idx, t;
                F::out_coord(gtid, j, idx, t);
                Ar[idx] = v[j];
            }
            return;
        }
    }
    const float2 *Sr = spec + ((size_t)(row & ((1 << log2n1) - 1)) << L2);
    if constexpr (MODE == ROW_INVERSE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int idx, t;
            I::template in_coord<0>(gtid, j, idx, t);
            v[j] = Ar[idx];
        }
    }
    // the product with the conjugate snippet spectrum rides on the first inverse butterflies; register j holds
    // element I::in_coord<0>(j) == F::out_coord(j): no exchange across the multiply
    I::butterfly0_mul(v, [&](int j) {
        int idx, t;
        I::template in_coord<0>(gtid, j, idx, t);
        return __ldg(&Sr[idx]);
    });
    AM_TL_WAIT(v, 32);
    AM_TL(3);
    // fused: the forward transform's last exchange reads must be over before the inverse writes (barrier after the
    // first inverse butterflies); nothing touches the buffer after the inverse
    I::template run<0, false, MODE == ROW_FUSED, true>(v, sm, gtid, tw);
    AM_TL(4);
    float2 *Or = (MODE == ROW_INVERSE) ? Bout + ((size_t)row << L2) : Ar;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        int idx, t;
        I::out_coord(gtid, j, idx, t);
        Or[idx] = v[j];
    }
    AM_TL(5);
}

// Persistent variant of k_row32<L2, ROW_FUSED>: grid = resident CTAs, rows handed out dynamically.  The row kernel
// spends ~40 % of a CTA's life waiting (launch + row load, spectrum load, store drain; timeline in DESIGN.md) and two
// CTAs of 8 warps cannot fill the issue slots while one of them waits.  Here the exchange buffer doubles as the
// landing zone of bulk asynchronous copies (cp.async.bulk, 64 KB each) issued by one thread at the two points where
// the buffer is idle: the snippet-spectrum row arrives during the last forward butterflies, the NEXT row of A during
// the last inverse butterflies and the stores, so both loads are off the critical path and cost no registers.
__device__ __forceinline__ void bulk_load(unsigned smem_dst, const void *gsrc, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
// MODE = ROW_FUSED (in place) or ROW_INVERSE (batch: A holds forward spectra, output to Bout; the spectrum row is an
// L2 hit thanks to the k1-major order and is read with plain loads while the row of A waits in shared memory).
// Batch mode (ROW_INVERSE) takes nsn snippets per launch: ticket = (row ticket) * nsn + snippet, so the nsn CTAs that
// need the same row of A run back to back (one DRAM fetch, the rest are L2 hits) while the k1-major row order keeps
// the nsn spectrum rows of the current k1 in L2 for all block pairs of the group.  Snippet j reads spec + j * spec_stride
// and writes Bout + j * b_stride.
template <int L2, int MODE>
__global__ void __launch_bounds__(Row32Cfg<L2>::THREADS, 2)
k_row32_stream(float2 *__restrict__ A, const float2 *__restrict__ spec, float2 *__restrict__ Bout, int log2n1, int rows,
               const float2 *__restrict__ tw, int *__restrict__ next_row, int nsn, size_t spec_stride, size_t b_stride) {
    typedef RegFFT<L2, 0, false, 32> F;
    typedef RegFFT<L2, 0, true, 32> I;
    static_assert(MODE == ROW_FUSED || MODE == ROW_INVERSE, "fused or inverse-only");
    constexpr unsigned ROW_BYTES = (unsigned)sizeof(float2) << L2, PIECE = 16384;
    extern __shared__ __align__(128) float2 sm[];
    __shared__ __align__(8) unsigned long long bar_store[2];
    __shared__ int s_next;
    const int gtid = threadIdx.x;
    const unsigned sm_a = (unsigned)__cvta_generic_to_shared(sm);
    const unsigned bar_row = (unsigned)__cvta_generic_to_shared(&bar_store[0]);
    const unsigned bar_spec = (unsigned)__cvta_generic_to_shared(&bar_store[1]);
    const int np = rows >> log2n1;
    const int tickets = rows * nsn;
    auto row_of = [&](int ticket) { const int r = ticket / nsn; return ((r % np) << log2n1) + r / np; };   // k1-major, see k_row32
    auto fetch = [&](const float2 *src, unsigned bar) {                 // one thread; the buffer must be idle
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, ROW_BYTES);
#pragma unroll
        for (unsigned o = 0; o < ROW_BYTES; o += PIECE) bulk_load(sm_a + o, (const char *)src + o, PIECE, bar);
    };
    if (gtid == 0) {
        mbar_init(bar_row, 1);
        mbar_init(bar_spec, 1);
    }
    __syncthreads();
    int ticket = blockIdx.x;
    if (gtid == 0 && ticket < tickets) fetch(A + ((size_t)row_of(ticket) << L2), bar_row);
    unsigned parity = 0;
    while (ticket < tickets) {
        const int row = row_of(ticket), sn = ticket % nsn;
        int claimed = 0;
        if (gtid == 0) claimed = (int)gridDim.x + atomicAdd(next_row, 1);
        const float2 *Sr = spec + (size_t)sn * spec_stride + ((size_t)(row & ((1 << log2n1) - 1)) << L2);
        float2 v[32];
        AM_TL_SET(row);
        AM_TL(0);
        if constexpr (MODE == ROW_FUSED) {
            mbar_wait(bar_row, parity);
            AM_TL(1);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int idx, t;
                F::template in_coord<0>(gtid, j, idx, t);
                v[j] = sm[idx];
            }
            if (gtid == 0) s_next = claimed;
            __syncthreads();                                           // row copied out; the buffer becomes the exchange buffer
            AM_TL(2);
            F::run_hook(v, sm, gtid, tw, [&] { if (gtid == 0) fetch(Sr, bar_spec); });
            AM_TL(3);
            mbar_wait(bar_spec, parity);
            AM_TL(4);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int idx, t;
                I::template in_coord<0>(gtid, j, idx, t);              // == F::out_coord: no exchange across the multiply
                v[j] = amfft::cmul(v[j], sm[idx]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {                             // spectrum row: in flight during the wait
                int idx, t;
                I::template in_coord<0>(gtid, j, idx, t);
                v[j] = __ldg(&Sr[idx]);
            }
            mbar_wait(bar_row, parity);
            AM_TL(4);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int idx, t;
                I::template in_coord<0>(gtid, j, idx, t);
                v[j] = amfft::cmul(sm[idx], v[j]);
            }
            if (gtid == 0) s_next = claimed;
        }
        __syncthreads();
        AM_TL(5);
        const int next = s_next;
        I::run_hook(v, sm, gtid, tw, [&] { if (gtid == 0 && next < tickets) fetch(A + ((size_t)row_of(next) << L2), bar_row); });
        AM_TL(6);
        float2 *Or = (MODE == ROW_INVERSE ? Bout + (size_t)sn * b_stride : A) + ((size_t)row << L2);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int idx, t;
            I::out_coord(gtid, j, idx, t);
            Or[idx] = v[j];
        }
        AM_TL(7);
        parity ^= 1;
        ticket = next;
    }
}

// sum of squares of the snippet in double (inverse_sample_auto_correlation, audio_matcher.rs:321-329)
__global__ void k_sumsq(StreamView sv, long long m, double *out) {
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        double s = (double)load_frame(sv, i);
        acc += s * s;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) atomicAdd(out, acc);
    }
}

// materialise a window of the stream as f32 (only used to hand the caller's PCM snippet back)
__global__ void k_to_f32(StreamView sv, long long first, long long count, float *out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < count) out[i] = load_frame(sv, first + i);
}

// ---- snippet spectrum in double precision ----------------------------------------------
// Runs once per matcher and block length, so the conjugate snippet spectrum the hot kernels
// multiply with is correctly rounded fp32 instead of carrying an fp32 transform's error.
// Plain Stockham radix-2 passes in global memory (ping-pong), natural-order output.
__global__ void k_spec64_load(StreamView sv, long long n, double2 *out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2((double)load_frame(sv, i), 0.0);
}
__global__ void k_spec64_pass(const double2 *__restrict__ in, double2 *__restrict__ out, long long half, long long ns) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= half) return;
    long long k = j & (ns - 1);
    double s, c;
    sincospi(-(double)k / (double)ns, &s, &c);          // exp(-2 pi i k / (2 ns))
    double2 a = in[j], b = in[j + half];
    double2 bw = make_double2(b.x * c - b.y * s, b.x * s + b.y * c);
    long long o = ((j - k) << 1) + k;
    out[o] = make_double2(a.x + bw.x, a.y + bw.y);
    out[o + ns] = make_double2(a.x - bw.x, a.y - bw.y);
}
// spec[(k & (N1-1)) * N2 + (k >> log2n1)] = conj(X[k])   (four-step layout; log2n1 = 0: natural)
__global__ void k_spec64_store(const double2 *__restrict__ X, long long n, int log2n1, int log2n2, float2 *spec) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    long long pos = ((k & ((1ll << log2n1) - 1)) << log2n2) + (k >> log2n1);
    double2 x = X[k];
    spec[pos] = make_float2((float)x.x, (float)(-x.y));
}

// ---- direct form for very short snippets -------------------------------------------------
// m <= DIRECT_MAX_M: c[g] = scalar * sum_j x[g + j] s[j], one output per thread, snippet in
// shared memory.  Cheaper than block transforms at this size and exact to fp32 rounding.
constexpr int DIRECT_MAX_M = 64;
__global__ void __launch_bounds__(256)
k_direct(StreamView sv, const float *__restrict__ snip, int m, long long g0, long long g1, float *c, long long c_g0,
         float scalar) {
    __shared__ float s_sn[DIRECT_MAX_M];
    __shared__ float s_x[256 + DIRECT_MAX_M];
    const long long base = g0 + (long long)blockIdx.x * 256;
    if (threadIdx.x < m) s_sn[threadIdx.x] = snip[threadIdx.x];
    for (int i = threadIdx.x; i < 256 + m - 1; i += 256) s_x[i] = load_frame(sv, base + i);
    __syncthreads();
    const long long g = base + threadIdx.x;
    if (g >= g1) return;
    float acc = 0.f;
    for (int j = 0; j < m; ++j) acc = fmaf(s_x[threadIdx.x + j], s_sn[j], acc);
    c[g - c_g0] = acc * scalar;
}

}  // namespace amk
