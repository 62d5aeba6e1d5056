// am_fft.cuh -- register-resident Stockham FFT building block (sm_100a).
//
// Every thread keeps 16 complex values in registers for the whole transform;
// shared memory is only the exchange medium between stages (one write + one
// read per stage boundary).  Radices are 2/4/8/16 chosen so that a length
// 2^LOG2N transform takes ceil(LOG2N/4) stages.  The same code serves
//   * row transforms      (BATCH = 1, one contiguous row per thread group)
//   * column-tile transforms (BATCH = T columns, element (i,t) at i*T+t)
// The inverse transform runs the forward plan backwards, so the registers a
// thread holds after the forward transform's last stage are exactly the inputs
// of the inverse transform's first stage (no exchange across the spectrum
// multiply) and the inverse's outputs land on the indices the thread loaded.
//
// The phase functions are __host__ __device__ so tests/host_emul.cu can run the
// identical index arithmetic on the CPU (this container has no GPU).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef AM_HD
#define AM_HD __host__ __device__ __forceinline__
#endif

// AM_TL(i): per-CTA phase timestamps for the micro-benchmarks under build/mb (compiled out of the library)
#ifndef AM_TL
#define AM_TL(i)
#define AM_TL_SET(tile)
#define AM_TL_WAIT(v, n)
#endif

namespace amfft {

constexpr int EPT = 16;           // complex elements per thread
constexpr int TW_LOG2 = 14;       // master twiddle table: W_{2^14}^j, j < 2^14
constexpr int TW_N = 1 << TW_LOG2;

AM_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
AM_HD float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a*conj(b)
AM_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
AM_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> AM_HD float2 mul_mi(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
// multiply by compile-time twiddle exp(-+ 2 pi i j/16) given as (c, s) = (cos, sin) of the angle magnitude
template <bool INV> AM_HD float2 mul_w(float2 a, float c, float s) {
    // forward: (c - i s); inverse: (c + i s)
    return INV ? make_float2(a.x * c - a.y * s, a.y * c + a.x * s) : make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
}

#define AM_C8 0.70710678118654752440f
#define AM_C16_1 0.92387953251128675613f
#define AM_S16_1 0.38268343236508977173f

template <bool INV> AM_HD void dft2(float2 &a, float2 &b) {
    float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}
// in-place natural-order 4-point DFT
template <bool INV> AM_HD void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    float2 s0 = cadd(a0, a2), d0 = csub(a0, a2), s1 = cadd(a1, a3), d1 = mul_mi<INV>(csub(a1, a3));
    a0 = cadd(s0, s1);
    a1 = cadd(d0, d1);
    a2 = csub(s0, s1);
    a3 = csub(d0, d1);
}
// 8 = 4 x 2: n = 2 n1 + n2, k = k1 + 4 k2
template <bool INV> AM_HD void dft8(float2 *v) {
    dft4<INV>(v[0], v[2], v[4], v[6]);    // n2 = 0 -> A[k1][0] in v[0],v[2],v[4],v[6]
    dft4<INV>(v[1], v[3], v[5], v[7]);    // n2 = 1 -> A[k1][1] in v[1],v[3],v[5],v[7]
    v[3] = mul_w<INV>(v[3], AM_C8, AM_C8);            // W8^1
    v[5] = mul_mi<INV>(v[5]);                          // W8^2
    v[7] = mul_w<INV>(v[7], -AM_C8, AM_C8);           // W8^3
    float2 x0 = cadd(v[0], v[1]), x4 = csub(v[0], v[1]);
    float2 x1 = cadd(v[2], v[3]), x5 = csub(v[2], v[3]);
    float2 x2 = cadd(v[4], v[5]), x6 = csub(v[4], v[5]);
    float2 x3 = cadd(v[6], v[7]), x7 = csub(v[6], v[7]);
    v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6; v[7] = x7;
}
// 16 = 4 x 4: n = 4 n1 + n2, k = k1 + 4 k2
template <bool INV> AM_HD void dft16(float2 *v) {
    dft4<INV>(v[0], v[4], v[8], v[12]);   // n2 = 0: A[k1][0] at v[4 k1 + 0]
    dft4<INV>(v[1], v[5], v[9], v[13]);
    dft4<INV>(v[2], v[6], v[10], v[14]);
    dft4<INV>(v[3], v[7], v[11], v[15]);
    // twiddle A[k1][n2] *= W16^{n2 k1}
    v[5] = mul_w<INV>(v[5], AM_C16_1, AM_S16_1);      // 1
    v[6] = mul_w<INV>(v[6], AM_C8, AM_C8);            // 2
    v[7] = mul_w<INV>(v[7], AM_S16_1, AM_C16_1);      // 3
    v[9] = mul_w<INV>(v[9], AM_C8, AM_C8);            // 2
    v[10] = mul_mi<INV>(v[10]);                        // 4
    v[11] = mul_w<INV>(v[11], -AM_C8, AM_C8);         // 6
    v[13] = mul_w<INV>(v[13], AM_S16_1, AM_C16_1);    // 3
    v[14] = mul_w<INV>(v[14], -AM_C8, AM_C8);         // 6
    v[15] = mul_w<INV>(v[15], -AM_C16_1, -AM_S16_1);  // 9: cos(9pi/8) = -c1, sin(9pi/8) = -s1
    // second step: for each k1, DFT4 over n2: X[k1 + 4 k2]
    dft4<INV>(v[0], v[1], v[2], v[3]);
    dft4<INV>(v[4], v[5], v[6], v[7]);
    dft4<INV>(v[8], v[9], v[10], v[11]);
    dft4<INV>(v[12], v[13], v[14], v[15]);
    // v[4 k1 + k2] holds X[k1 + 4 k2]: transpose to natural order
    float2 t;
#define AM_SWAP(a, b) t = v[a]; v[a] = v[b]; v[b] = t;
    AM_SWAP(1, 4) AM_SWAP(2, 8) AM_SWAP(3, 12) AM_SWAP(6, 9) AM_SWAP(7, 13) AM_SWAP(11, 14)
#undef AM_SWAP
}
// 32 = 2 x 16: even / odd halves through dft16, W32^k on the odd half, radix-2 combine
template <bool INV> AM_HD void dft32(float2 *v) {
    float2 e[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    dft16<INV>(e);
    dft16<INV>(o);
    // cos / sin of pi k / 16, k = 0..15
    constexpr float C[16] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
                             0.f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                             -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f};
    constexpr float S[16] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                             0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
                             1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        float2 t = (k == 0) ? o[0] : (k == 8 ? mul_mi<INV>(o[8]) : mul_w<INV>(o[k], C[k], S[k]));
        v[k] = cadd(e[k], t);
        v[k + 16] = csub(e[k], t);
    }
}
template <int R, bool INV> AM_HD void dft(float2 *v) {
    if constexpr (R == 2) dft2<INV>(v[0], v[1]);
    else if constexpr (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else if constexpr (R == 8) dft8<INV>(v);
    else if constexpr (R == 16) dft16<INV>(v);
    else dft32<INV>(v);
}

// ---- FMA-fused butterflies ---------------------------------------------------------------------
// A twiddle multiply that feeds an add/sub pair costs 8 instructions when done separately (4 for the
// product, 4 for the pair).  Folding the product into the sum (a + w b: 4 FMA) and taking the
// difference as 2a - (a + w b) (2 FMA) makes it 6.  The radix-16 / radix-32 butterflies below apply
// this to the inter-stage twiddles (first dft4 layer), to the constant W16 twiddles between the two
// dft4 layers and to the radix-2 combine of the radix-32 butterfly: ~10 % fewer FP instructions per
// transform.  AM_FFT_LEGACY keeps the unfused butterflies for A/B measurements.
AM_HD float2 cfma(float2 a, float2 w, float2 b) {           // a + w b
    return make_float2(fmaf(w.x, b.x, fmaf(-w.y, b.y, a.x)), fmaf(w.x, b.y, fmaf(w.y, b.x, a.y)));
}
AM_HD float2 cdbl_sub(float2 a, float2 s) { return make_float2(fmaf(2.0f, a.x, -s.x), fmaf(2.0f, a.y, -s.y)); }   // 2a - s
template <bool INV> AM_HD float2 cw(float c, float s) { return make_float2(c, INV ? s : -s); }   // exp(-+ i phi), (c, s) = (cos, sin) phi
// dft4 of (a0, w1 a1, w2 a2, w3 a3), natural order, in place: 24 instructions (28 unfused)
template <bool INV> AM_HD void dft4_tw(float2 &a0, float2 &a1, float2 &a2, float2 &a3, float2 w1, float2 w2, float2 w3) {
    const float2 s0 = cfma(a0, w2, a2), d0 = cdbl_sub(a0, s0);
    const float2 t1 = cmul(a1, w1);
    const float2 s1 = cfma(t1, w3, a3), d1 = mul_mi<INV>(cdbl_sub(t1, s1));
    a0 = cadd(s0, s1);
    a1 = cadd(d0, d1);
    a2 = csub(s0, s1);
    a3 = csub(d0, d1);
}
// second dft4 layer of the 16-point butterfly with the constant twiddles W16^{n2 k1} folded in, then the
// transposition to natural order
template <bool INV> AM_HD void dft16_layer2(float2 *v) {
    dft4<INV>(v[0], v[1], v[2], v[3]);
    dft4_tw<INV>(v[4], v[5], v[6], v[7], cw<INV>(AM_C16_1, AM_S16_1), cw<INV>(AM_C8, AM_C8), cw<INV>(AM_S16_1, AM_C16_1));
    {   // k1 = 2: twiddles W16^2, W16^4 = -+i, W16^6
        const float2 m = mul_mi<INV>(v[10]);
        const float2 s0 = cadd(v[8], m), d0 = csub(v[8], m);
        const float2 t1 = cmul(v[9], cw<INV>(AM_C8, AM_C8));
        const float2 s1 = cfma(t1, cw<INV>(-AM_C8, AM_C8), v[11]), d1 = mul_mi<INV>(cdbl_sub(t1, s1));
        v[8] = cadd(s0, s1);
        v[9] = cadd(d0, d1);
        v[10] = csub(s0, s1);
        v[11] = csub(d0, d1);
    }
    dft4_tw<INV>(v[12], v[13], v[14], v[15], cw<INV>(AM_S16_1, AM_C16_1), cw<INV>(-AM_C8, AM_C8), cw<INV>(-AM_C16_1, -AM_S16_1));
    float2 t;
#define AM_SWAP(a, b) t = v[a]; v[a] = v[b]; v[b] = t;
    AM_SWAP(1, 4) AM_SWAP(2, 8) AM_SWAP(3, 12) AM_SWAP(6, 9) AM_SWAP(7, 13) AM_SWAP(11, 14)
#undef AM_SWAP
}
// NBF 16-point butterflies (v[16 f + r]) whose inputs carry the same geometric twiddles:
//   TW = 0: none;  TW = 1: v[r] *= u^r;  TW = 2: v[r] *= b u^r.
// The twiddles of one first-layer group (r = n2 + 4 n1) are generated right before they are used
// (t0 = b u^n2, t0 u^4, t0 u^8, t0 u^12), so few of them are live at a time.
template <bool INV, int TW, int NBF> AM_HD void dft16_f(float2 *v, float2 b, float2 u) {
    if constexpr (TW == 0) {
#pragma unroll
        for (int f = 0; f < NBF; ++f) {
            float2 *x = v + 16 * f;
            dft4<INV>(x[0], x[4], x[8], x[12]);
            dft4<INV>(x[1], x[5], x[9], x[13]);
            dft4<INV>(x[2], x[6], x[10], x[14]);
            dft4<INV>(x[3], x[7], x[11], x[15]);
        }
    } else {
        const float2 u2 = cmul(u, u), u3 = cmul(u2, u), u4 = cmul(u2, u2), u8 = cmul(u4, u4), u12 = cmul(u8, u4);
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            float2 t0 = n2 == 0 ? make_float2(1.f, 0.f) : (n2 == 1 ? u : (n2 == 2 ? u2 : u3));
            if constexpr (TW == 2) t0 = n2 == 0 ? b : cmul(b, t0);
            const bool unit = (TW == 1 && n2 == 0);
            const float2 t1 = unit ? u4 : cmul(t0, u4), t2 = unit ? u8 : cmul(t0, u8), t3 = unit ? u12 : cmul(t0, u12);
#pragma unroll
            for (int f = 0; f < NBF; ++f) {
                float2 *x = v + 16 * f;
                if (!unit) x[n2] = cmul(x[n2], t0);
                dft4_tw<INV>(x[n2], x[4 + n2], x[8 + n2], x[12 + n2], t1, t2, t3);
            }
        }
    }
#pragma unroll
    for (int f = 0; f < NBF; ++f) dft16_layer2<INV>(v + 16 * f);
}
// 32-point butterfly, optionally with input twiddles v[r] *= w^r (TW)
template <bool INV, bool TW> AM_HD void dft32_f(float2 *v, float2 w) {
    float2 e[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    if constexpr (TW) {
        const float2 u = cmul(w, w);
        dft16_f<INV, 1, 1>(e, u, u);
        dft16_f<INV, 2, 1>(o, w, u);
    } else {
        dft16_f<INV, 0, 1>(e, w, w);
        dft16_f<INV, 0, 1>(o, w, w);
    }
    constexpr float C[16] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
                             0.f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                             -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f};
    constexpr float S[16] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                             0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
                             1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k == 0 || k == 8) {
            const float2 t = k == 0 ? o[0] : mul_mi<INV>(o[8]);
            v[k] = cadd(e[k], t);
            v[k + 16] = csub(e[k], t);
        } else {
            v[k] = cfma(e[k], cw<INV>(C[k], S[k]), o[k]);
            v[k + 16] = cdbl_sub(e[k], v[k]);
        }
    }
}

// v[r] *= w^r, r = 1..R-1, with w^r built by a depth-log2(R) product tree
template <int R> AM_HD void apply_twiddle_powers(float2 *v, float2 w1) {
    v[1] = cmul(v[1], w1);
    if constexpr (R >= 4) {
        float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        if constexpr (R >= 8) {
            float2 w4 = cmul(w2, w2), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
            v[4] = cmul(v[4], w4); v[5] = cmul(v[5], w5); v[6] = cmul(v[6], w6); v[7] = cmul(v[7], w7);
            if constexpr (R >= 16) {
                float2 w8 = cmul(w4, w4);
                v[8] = cmul(v[8], w8);
                v[9] = cmul(v[9], cmul(w8, w1));   v[10] = cmul(v[10], cmul(w8, w2));
                v[11] = cmul(v[11], cmul(w8, w3)); v[12] = cmul(v[12], cmul(w8, w4));
                v[13] = cmul(v[13], cmul(w8, w5)); v[14] = cmul(v[14], cmul(w8, w6));
                v[15] = cmul(v[15], cmul(w8, w7));
                if constexpr (R >= 32) {
                    const float2 w16 = cmul(w8, w8);
                    const float2 lo[8] = {make_float2(1.f, 0.f), w1, w2, w3, w4, w5, w6, w7};
                    v[16] = cmul(v[16], w16);
#pragma unroll
                    for (int i = 1; i < 8; ++i) v[16 + i] = cmul(v[16 + i], cmul(w16, lo[i]));
                    const float2 w24 = cmul(w16, w8);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[24 + i] = cmul(v[24 + i], i ? cmul(w24, lo[i]) : w24);
                }
            }
        }
    }
}

// Transform plan: LOG2N = length, LOG2B = batch (columns interleaved), INV = direction.
// GT = threads in the group that owns the N*B elements (N*B == 16*GT).
// E = complex elements per thread (16, or 32 for long rows: radix-32 first stage, one exchange fewer).
template <int LOG2N, int LOG2B, bool INV, int E = EPT> struct RegFFT {
    static constexpr int N = 1 << LOG2N, B = 1 << LOG2B;
    static constexpr int EPT = E;                    // shadows amfft::EPT inside the plan
    static constexpr int MAXB = (E == 32) ? 5 : 4;   // largest radix = E
    static constexpr int GT = (N * B) / EPT;
    static constexpr int NST = (LOG2N + MAXB - 1) / MAXB;
    static_assert(E == 16 || E == 32, "16 or 32 elements per thread");
    static_assert(LOG2N >= MAXB && LOG2N <= TW_LOG2, "unsupported length");
    static_assert(N * B >= EPT, "group too small");
    // radix bits of stage st in execution order (inverse = forward plan reversed)
    static constexpr int bits_at(int st) {
        int s = INV ? NST - 1 - st : st;
        return LOG2N / NST + (s < LOG2N % NST ? 1 : 0);
    }
    static constexpr int logns_at(int st) {          // log2 of the product of the radices before stage st
        int a = 0;
        for (int i = 0; i < st; ++i) a += bits_at(i);
        return a;
    }
    // padded shared-memory slot of flat element index i (i = idx*B + t)
    static AM_HD int slot(int i) { return B >= 16 ? i : i + (i >> 4); }
    static constexpr int SMEM_ELEMS = (B >= 16) ? N * B : N * B + ((N * B) >> 4);

    // element index (within the length-N transform) that register j holds BEFORE stage st,
    // and the batch column: used by callers for the first stage's loads.
    template <int ST> static AM_HD void in_coord(int gtid, int j, int &idx, int &t) {
        constexpr int RB = bits_at(ST), R = 1 << RB;
        int l = j >> RB, r = j & (R - 1);
        int id = gtid + l * GT;
        t = id & (B - 1);
        idx = (id >> LOG2B) + r * (N >> RB);
    }
    // element index register j holds AFTER the last stage (natural order output index)
    static AM_HD void out_coord(int gtid, int j, int &idx, int &t) {
        constexpr int RB = bits_at(NST - 1), R = 1 << RB;
        int l = j >> RB, s = j & (R - 1);
        int id = gtid + l * GT;
        t = id & (B - 1);
        idx = (id >> LOG2B) + s * (N >> RB);
    }

    // twiddle W_{Ns R}^k of butterfly l of stage ST (conjugated for the inverse)
    template <int ST> static AM_HD float2 stage_twiddle(int gtid, int l, const float2 *__restrict__ tw) {
        constexpr int RB = bits_at(ST), LOGNS = logns_at(ST);
        const int q = (gtid + l * GT) >> LOG2B;
        const int k = q & ((1 << LOGNS) - 1);
        float2 w = tw[k << (TW_LOG2 - LOGNS - RB)];
        if (INV) w.y = -w.y;
        return w;
    }
    // twiddle + butterflies of stage ST on the registers
    template <int ST> static AM_HD void butterfly(float2 (&v)[EPT], int gtid, const float2 *__restrict__ tw) {
        constexpr int RB = bits_at(ST), R = 1 << RB, NB = EPT / R, LOGNS = logns_at(ST);
#ifndef AM_FFT_LEGACY
        if constexpr (R == 16 || R == 32) {
            if constexpr (ST == 0) {
                if constexpr (R == 16) dft16_f<INV, 0, NB>(v, make_float2(1.f, 0.f), make_float2(1.f, 0.f));
                else {
#pragma unroll
                    for (int l = 0; l < NB; ++l) dft32_f<INV, false>(&v[l * R], make_float2(1.f, 0.f));
                }
            } else if constexpr (R == 16 && NB > 1 && ((GT >> LOG2B) % (1 << LOGNS)) == 0) {
                // every butterfly of the thread has the same k: one set of twiddle powers serves all of them
                const float2 w = stage_twiddle<ST>(gtid, 0, tw);
                dft16_f<INV, 1, NB>(v, w, w);
            } else {
#pragma unroll
                for (int l = 0; l < NB; ++l) {
                    const float2 w = stage_twiddle<ST>(gtid, l, tw);
                    if constexpr (R == 16) dft16_f<INV, 1, 1>(&v[l * R], w, w);
                    else dft32_f<INV, true>(&v[l * R], w);
                }
            }
            return;
        }
#endif
#pragma unroll
        for (int l = 0; l < NB; ++l) {
            if constexpr (ST > 0) apply_twiddle_powers<R>(&v[l * R], stage_twiddle<ST>(gtid, l, tw));
            dft<R, INV>(&v[l * R]);
        }
    }
    // Stage 0 with an arbitrary multiplier per register, v[j] *= mul(j) (the snippet-spectrum product of the row pass
    // rides on the first inverse butterflies: 28 instead of 32 instructions per 4-point group, and the multipliers
    // are consumed group by group while later ones are still in flight)
    template <class Mul> static AM_HD void butterfly0_mul(float2 (&v)[EPT], Mul &&mul) {
        constexpr int RB = bits_at(0), R = 1 << RB, NB = EPT / R;
#ifndef AM_FFT_LEGACY
        if constexpr (R == 16) {
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
#pragma unroll
                for (int f = 0; f < NB; ++f) {
                    float2 *x = &v[16 * f];
                    x[n2] = cmul(x[n2], mul(16 * f + n2));
                    dft4_tw<INV>(x[n2], x[4 + n2], x[8 + n2], x[12 + n2], mul(16 * f + 4 + n2), mul(16 * f + 8 + n2), mul(16 * f + 12 + n2));
                }
            }
#pragma unroll
            for (int f = 0; f < NB; ++f) dft16_layer2<INV>(&v[16 * f]);
            return;
        }
#endif
#pragma unroll
        for (int j = 0; j < EPT; ++j) v[j] = cmul(v[j], mul(j));
#pragma unroll
        for (int l = 0; l < NB; ++l) dft<R, INV>(&v[l * R]);
    }
    // Stage 0 with geometric input twiddles v[l R + r] *= base[l] step^r (the four-step twiddles of the inverse
    // column pass ride on the first butterflies)
    static AM_HD void butterfly0_geo(float2 (&v)[EPT], const float2 *base, float2 step) {
        constexpr int RB = bits_at(0), R = 1 << RB, NB = EPT / R;
#pragma unroll
        for (int l = 0; l < NB; ++l) {
#ifndef AM_FFT_LEGACY
            if constexpr (R == 16) { dft16_f<INV, 2, 1>(&v[l * R], base[l], step); continue; }
#endif
            float2 w[R];
            w[0] = base[l];
            if (R >= 2) w[1] = cmul(base[l], step);
            float2 sp = step;
#pragma unroll
            for (int h = 2; h < R; h <<= 1) {
                sp = cmul(sp, sp);
#pragma unroll
                for (int i = 0; i < h; ++i) w[h + i] = cmul(w[i], sp);
            }
#pragma unroll
            for (int i = 0; i < R; ++i) v[l * R + i] = cmul(v[l * R + i], w[i]);
            dft<R, INV>(&v[l * R]);
        }
    }
    // scatter stage ST's outputs into the exchange buffer (Stockham autosort index).
    // The padded slot of output s is slot(A) + a compile-time offset: A = j*Ns*R + k with k < Ns, so
    // adding s*Ns never carries out of the low four bits that the padding term (i >> 4) drops.
    template <int ST> static AM_HD void xchg_write(const float2 (&v)[EPT], float2 *sm, int gtid) {
        constexpr int RB = bits_at(ST), R = 1 << RB, NB = EPT / R, LOGNS = logns_at(ST), NS = 1 << LOGNS;
#pragma unroll
        for (int l = 0; l < NB; ++l) {
            int id = gtid + l * GT;
            int t = id & (B - 1), q = id >> LOG2B;
            int k = q & (NS - 1);
            int base = ((q - k) << RB) + k;
            if constexpr (B == 1) {
                const int sb = base + (base >> 4);
#pragma unroll
                for (int s = 0; s < R; ++s) sm[sb + s * NS + ((s * NS) >> 4)] = v[l * R + s];
            } else if constexpr (B >= 16) {
                const int sb = (base << LOG2B) + t;
#pragma unroll
                for (int s = 0; s < R; ++s) sm[sb + ((s * NS) << LOG2B)] = v[l * R + s];
            } else {
#pragma unroll
                for (int s = 0; s < R; ++s) sm[slot(((base + (s << LOGNS)) << LOG2B) + t)] = v[l * R + s];
            }
        }
    }
    // gather the inputs of stage ST (ST >= 1) from the exchange buffer
    template <int ST> static AM_HD void xchg_read(float2 (&v)[EPT], const float2 *sm, int gtid) {
        constexpr int RB = bits_at(ST), R = 1 << RB, NB = EPT / R, M = N >> RB;
#pragma unroll
        for (int l = 0; l < NB; ++l) {
            int id = gtid + l * GT;
            int t = id & (B - 1), q = id >> LOG2B;
            if constexpr (B == 1 && M >= 16) {
                const int sb = q + (q >> 4);
#pragma unroll
                for (int r = 0; r < R; ++r) v[l * R + r] = sm[sb + r * (M + M / 16)];
            } else if constexpr (B >= 16) {
                const int sb = (q << LOG2B) + t;
#pragma unroll
                for (int r = 0; r < R; ++r) v[l * R + r] = sm[sb + ((r * M) << LOG2B)];
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) v[l * R + r] = sm[slot(((q + r * M) << LOG2B) + t)];
            }
        }
    }

#ifdef __CUDACC__
    // Full transform on the device.  On entry v holds the stage-0 inputs (see in_coord<0>),
    // on exit the natural-order outputs (see out_coord).  All threads of the CTA must call.
    // TAIL_SYNC = false leaves out the barrier after the last exchange read: the caller then guarantees a
    // __syncthreads() before the buffer is written again.
    // LEAD_SYNC = true puts a barrier in front of the first exchange write (after the first butterflies) for a
    // caller whose threads may still be reading the buffer when they enter.
    // SKIP_B0 = true: the caller has already done stage 0's butterflies (butterfly0_geo).
    template <int ST = 0, bool TAIL_SYNC = true, bool LEAD_SYNC = false, bool SKIP_B0 = false>
    static __device__ __forceinline__ void run(float2 (&v)[EPT], float2 *sm, int gtid, const float2 *__restrict__ tw) {
        if constexpr (!(ST == 0 && SKIP_B0)) butterfly<ST>(v, gtid, tw);
        AM_TL(8 + (INV ? 9 : 0) + ST * 3);
        if constexpr (ST + 1 < NST) {
            if constexpr (ST == 0 && LEAD_SYNC) __syncthreads();
            xchg_write<ST>(v, sm, gtid);
            __syncthreads();
            AM_TL(8 + (INV ? 9 : 0) + ST * 3 + 1);
            xchg_read<ST + 1>(v, sm, gtid);
            if constexpr (TAIL_SYNC || ST + 2 < NST) __syncthreads();
            AM_TL(8 + (INV ? 9 : 0) + ST * 3 + 2);
            run<ST + 1, TAIL_SYNC, LEAD_SYNC, SKIP_B0>(v, sm, gtid, tw);
        }
    }
    // Same, calling hook() once when the exchange buffer has been read for the last time (before the last stage's
    // butterflies): from there on the buffer is free, e.g. as the landing zone of an asynchronous copy.
    template <int ST = 0, class Hook>
    static __device__ __forceinline__ void run_hook(float2 (&v)[EPT], float2 *sm, int gtid, const float2 *__restrict__ tw,
                                                    Hook &&hook) {
        if constexpr (ST + 1 == NST) hook();
        butterfly<ST>(v, gtid, tw);
        if constexpr (ST + 1 < NST) {
            xchg_write<ST>(v, sm, gtid);
            __syncthreads();
            xchg_read<ST + 1>(v, sm, gtid);
            __syncthreads();
            run_hook<ST + 1>(v, sm, gtid, tw, hook);
        }
    }
#endif
};

}  // namespace amfft
