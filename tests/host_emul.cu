// host_emul.cu -- runs the register-FFT phase functions of am_fft.cuh on the CPU
// (threads become loop iterations, __syncthreads becomes a phase boundary) and
// checks them against a double-precision DFT.  Built and run by
// tests/test_host_emul.py with `nvcc -x cu` as a host-only program: this is how
// the index arithmetic of the CUDA kernels is verified in a container with no GPU.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../audio_matcher_b200/csrc/am_fft.cuh"

using namespace amfft;
typedef std::complex<double> cd;

static void ref_fft(std::vector<cd> &a, bool inv) {   // recursive radix-2, double
    size_t n = a.size();
    if (n == 1) return;
    std::vector<cd> e(n / 2), o(n / 2);
    for (size_t i = 0; i < n / 2; ++i) { e[i] = a[2 * i]; o[i] = a[2 * i + 1]; }
    ref_fft(e, inv); ref_fft(o, inv);
    for (size_t k = 0; k < n / 2; ++k) {
        double ang = (inv ? 2.0 : -2.0) * M_PI * (double)k / (double)n;
        cd w(cos(ang), sin(ang));
        a[k] = e[k] + w * o[k];
        a[k + n / 2] = e[k] - w * o[k];
    }
}

static std::vector<float2> g_tw;

template <class F, int ST> struct Runner {
    static constexpr int EPT = F::EPT;
    static void go(std::vector<float2> &regs, std::vector<float2> &sm) {
        constexpr int GT = F::GT;
        for (int t = 0; t < GT; ++t) F::template butterfly<ST>(*(float2(*)[EPT]) & regs[t * EPT], t, g_tw.data());
        if constexpr (ST + 1 < F::NST) {
            for (int t = 0; t < GT; ++t) F::template xchg_write<ST>(*(const float2(*)[EPT]) & regs[t * EPT], sm.data(), t);
            for (int t = 0; t < GT; ++t) F::template xchg_read<ST + 1>(*(float2(*)[EPT]) & regs[t * EPT], sm.data(), t);
            Runner<F, ST + 1>::go(regs, sm);
        }
    }
};

template <int LOG2N, int LOG2B, bool INV, int E = 16> static double check() {
    typedef RegFFT<LOG2N, LOG2B, INV, E> F;
    constexpr int N = F::N, B = F::B, GT = F::GT, EPT = E;
    std::vector<float2> x(N * B);
    srand(1234 + LOG2N * 31 + LOG2B);
    for (auto &e : x) e = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    std::vector<float2> regs(GT * EPT), sm(F::SMEM_ELEMS, make_float2(NAN, NAN));
    for (int t = 0; t < GT; ++t)
        for (int j = 0; j < EPT; ++j) {
            int idx, c;
            F::template in_coord<0>(t, j, idx, c);
            regs[t * EPT + j] = x[idx * B + c];
        }
    Runner<F, 0>::go(regs, sm);
    std::vector<float2> y(N * B, make_float2(NAN, NAN));
    for (int t = 0; t < GT; ++t)
        for (int j = 0; j < EPT; ++j) {
            int idx, c;
            F::out_coord(t, j, idx, c);
            y[idx * B + c] = regs[t * EPT + j];
        }
    double maxerr = 0, maxref = 0;
    for (int c = 0; c < B; ++c) {
        std::vector<cd> a(N);
        for (int i = 0; i < N; ++i) a[i] = cd(x[i * B + c].x, x[i * B + c].y);
        ref_fft(a, INV);
        for (int i = 0; i < N; ++i) {
            double e = std::abs(a[i] - cd(y[i * B + c].x, y[i * B + c].y));
            if (!(e <= maxerr)) maxerr = e;          // NaN-propagating max
            maxref = std::max(maxref, std::abs(a[i]));
        }
    }
    double rel = maxerr / maxref;
    printf("fft log2n=%d log2b=%d inv=%d ept=%d stages=%d  max_rel_err=%.3g %s\n", LOG2N, LOG2B, (int)INV, E, F::NST, rel,
           rel < 2e-6 ? "OK" : "FAIL");
    return rel;
}

// the forward-then-inverse register hand-over used by the row kernel: after the forward
// transform the thread multiplies its registers and feeds them to the inverse unchanged.
template <int LOG2N, int E = 16> static double check_roundtrip() {
    typedef RegFFT<LOG2N, 0, false, E> F;
    typedef RegFFT<LOG2N, 0, true, E> I;
    constexpr int N = F::N, GT = F::GT, EPT = E;
    std::vector<float2> x(N);
    for (auto &e : x) e = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    std::vector<float2> regs(GT * EPT), sm(F::SMEM_ELEMS);
    double bad = 0;
    for (int t = 0; t < GT; ++t)
        for (int j = 0; j < EPT; ++j) {
            int idx, c, idx2, c2;
            F::template in_coord<0>(t, j, idx, c);
            regs[t * EPT + j] = x[idx];
            // coordinates must line up: fwd out == inv in, inv out == fwd in
            F::out_coord(t, j, idx, c); I::template in_coord<0>(t, j, idx2, c2);
            if (idx != idx2) bad = 1;
            F::template in_coord<0>(t, j, idx, c); I::out_coord(t, j, idx2, c2);
            if (idx != idx2) bad = 1;
        }
    Runner<F, 0>::go(regs, sm);
    Runner<I, 0>::go(regs, sm);
    double maxerr = 0;
    for (int t = 0; t < GT; ++t)
        for (int j = 0; j < EPT; ++j) {
            int idx, c;
            I::out_coord(t, j, idx, c);
            double ex = regs[t * EPT + j].x / N - x[idx].x, ey = regs[t * EPT + j].y / N - x[idx].y;
            maxerr = std::max(maxerr, std::sqrt(ex * ex + ey * ey));
        }
    printf("roundtrip log2n=%d coord_mismatch=%d max_err=%.3g %s\n", LOG2N, (int)bad, maxerr,
           (bad == 0 && maxerr < 2e-6) ? "OK" : "FAIL");
    return bad ? 1.0 : maxerr;
}

// stage 0 with geometric input twiddles (RegFFT::butterfly0_geo, the inverse column pass): element r of
// butterfly l of thread t is multiplied by base[t][l] * step[t]^r before the transform
template <class F, int ST> struct RunnerFrom1 {
    static constexpr int EPT = F::EPT;
    static void go(std::vector<float2> &regs, std::vector<float2> &sm) {
        constexpr int GT = F::GT;
        if constexpr (ST > 0)
            for (int t = 0; t < GT; ++t) F::template butterfly<ST>(*(float2(*)[EPT]) & regs[t * EPT], t, g_tw.data());
        if constexpr (ST + 1 < F::NST) {
            for (int t = 0; t < GT; ++t) F::template xchg_write<ST>(*(const float2(*)[EPT]) & regs[t * EPT], sm.data(), t);
            for (int t = 0; t < GT; ++t) F::template xchg_read<ST + 1>(*(float2(*)[EPT]) & regs[t * EPT], sm.data(), t);
            RunnerFrom1<F, ST + 1>::go(regs, sm);
        }
    }
};
template <int LOG2N, int LOG2B, int E> static double check_geo() {
    typedef RegFFT<LOG2N, LOG2B, true, E> F;
    constexpr int N = F::N, B = F::B, GT = F::GT, EPT = E, RB = F::bits_at(0), R = 1 << RB, NB = EPT / R;
    std::vector<float2> x(N * B);
    for (auto &e : x) e = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    std::vector<cd> xt(N * B);                       // twiddled inputs in double
    std::vector<float2> regs(GT * EPT), sm(F::SMEM_ELEMS, make_float2(NAN, NAN));
    for (int t = 0; t < GT; ++t) {
        float2 base[NB];
        const double as = 0.37 * (t % 7 + 1) / R;
        const float2 step = make_float2((float)cos(as), (float)sin(as));
        for (int l = 0; l < NB; ++l) {
            const double ab = 0.11 * (t + 3 * l);
            base[l] = make_float2((float)(0.5 * cos(ab)), (float)(0.5 * sin(ab)));
            for (int r = 0; r < R; ++r) {
                int idx, c;
                F::template in_coord<0>(t, l * R + r, idx, c);
                regs[t * EPT + l * R + r] = x[idx * B + c];
                xt[idx * B + c] = cd(x[idx * B + c].x, x[idx * B + c].y) * cd(base[l].x, base[l].y) * std::pow(cd(step.x, step.y), r);
            }
        }
        F::butterfly0_geo(*(float2(*)[EPT]) & regs[t * EPT], base, step);
    }
    RunnerFrom1<F, 0>::go(regs, sm);
    double maxerr = 0, maxref = 0;
    for (int c = 0; c < B; ++c) {
        std::vector<cd> a(N);
        for (int i = 0; i < N; ++i) a[i] = xt[i * B + c];
        ref_fft(a, true);
        for (int t = 0; t < GT; ++t)
            for (int j = 0; j < EPT; ++j) {
                int idx, cc;
                F::out_coord(t, j, idx, cc);
                if (cc != c) continue;
                double e = std::abs(a[idx] - cd(regs[t * EPT + j].x, regs[t * EPT + j].y));
                if (!(e <= maxerr)) maxerr = e;
                maxref = std::max(maxref, std::abs(a[idx]));
            }
    }
    double rel = maxerr / maxref;
    printf("geo-twiddled stage 0: log2n=%d log2b=%d ept=%d radix0=%d  max_rel_err=%.3g %s\n", LOG2N, LOG2B, E, R, rel, rel < 2e-6 ? "OK" : "FAIL");
    return rel;
}

// stage 0 with one multiplier per register (RegFFT::butterfly0_mul, the spectrum product of the row pass)
template <int LOG2N, int E> static double check_mul() {
    typedef RegFFT<LOG2N, 0, true, E> F;
    constexpr int N = F::N, GT = F::GT, EPT = E;
    std::vector<float2> x(N), w(N);
    for (auto &e : x) e = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    for (auto &e : w) e = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    std::vector<float2> regs(GT * EPT), sm(F::SMEM_ELEMS, make_float2(NAN, NAN));
    for (int t = 0; t < GT; ++t) {
        for (int j = 0; j < EPT; ++j) {
            int idx, c;
            F::template in_coord<0>(t, j, idx, c);
            regs[t * EPT + j] = x[idx];
        }
        F::butterfly0_mul(*(float2(*)[EPT]) & regs[t * EPT], [&](int j) {
            int idx, c;
            F::template in_coord<0>(t, j, idx, c);
            return w[idx];
        });
    }
    RunnerFrom1<F, 0>::go(regs, sm);
    std::vector<cd> a(N);
    for (int i = 0; i < N; ++i) a[i] = cd(x[i].x, x[i].y) * cd(w[i].x, w[i].y);
    ref_fft(a, true);
    double maxerr = 0, maxref = 0;
    for (int t = 0; t < GT; ++t)
        for (int j = 0; j < EPT; ++j) {
            int idx, c;
            F::out_coord(t, j, idx, c);
            double e = std::abs(a[idx] - cd(regs[t * EPT + j].x, regs[t * EPT + j].y));
            if (!(e <= maxerr)) maxerr = e;
            maxref = std::max(maxref, std::abs(a[idx]));
        }
    double rel = maxerr / maxref;
    printf("multiplied stage 0: log2n=%d ept=%d  max_rel_err=%.3g %s\n", LOG2N, E, rel, rel < 2e-6 ? "OK" : "FAIL");
    return rel;
}

int main() {
    g_tw.resize(TW_N);
    for (int j = 0; j < TW_N; ++j) {
        double a = -2.0 * M_PI * j / TW_N;
        g_tw[j] = make_float2((float)cos(a), (float)sin(a));
    }
    double worst = 0;
#define CHK(n, b) worst = std::max(worst, check<n, b, false>()); worst = std::max(worst, check<n, b, true>());
    CHK(4, 0) CHK(5, 0) CHK(6, 0) CHK(7, 0) CHK(8, 0) CHK(9, 0) CHK(10, 0) CHK(11, 0) CHK(12, 0) CHK(13, 0) CHK(14, 0)
    CHK(4, 4) CHK(5, 4) CHK(6, 4) CHK(7, 4) CHK(8, 4) CHK(9, 4) CHK(10, 4) CHK(4, 2) CHK(6, 3) CHK(9, 5)
    worst = std::max(worst, check_roundtrip<4>());
    worst = std::max(worst, check_roundtrip<9>());
    worst = std::max(worst, check_roundtrip<13>());
#define CHK32(n) worst = std::max(worst, check<n, 0, false, 32>()); worst = std::max(worst, check<n, 0, true, 32>());
    CHK32(5) CHK32(9) CHK32(10) CHK32(12) CHK32(13) CHK32(14)
    worst = std::max(worst, check<9, 4, false, 32>()); worst = std::max(worst, check<9, 4, true, 32>());
    worst = std::max(worst, check<10, 4, false, 32>()); worst = std::max(worst, check<10, 4, true, 32>());
    worst = std::max(worst, check_roundtrip<12, 32>());
    worst = std::max(worst, check_roundtrip<13, 32>());
    worst = std::max(worst, check_roundtrip<14, 32>());
    worst = std::max(worst, check_mul<13, 32>());
    worst = std::max(worst, check_mul<14, 32>());
    worst = std::max(worst, check_mul<12, 32>());
    worst = std::max(worst, check_mul<10, 16>());
    worst = std::max(worst, check_geo<9, 4, 32>());
    worst = std::max(worst, check_geo<9, 4, 16>());
    worst = std::max(worst, check_geo<10, 3, 16>());
    worst = std::max(worst, check_geo<7, 4, 16>());
    worst = std::max(worst, check_geo<8, 4, 16>());
    worst = std::max(worst, check_geo<5, 6, 16>());
    printf(worst < 2e-6 ? "ALL OK\n" : "SOME FAILED\n");
    return worst < 2e-6 ? 0 : 1;
}
