// Timeline micro-benchmark of k_row32<13, ROW_FUSED>: per-CTA phase timestamps (clock64 of warp 0)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__device__ long long g_tl[16384 * 32];
__device__ int g_flag;
__device__ __forceinline__ void tl_rec(int i) {
    if (threadIdx.x == 0) {
        long long *p = g_tl + (size_t)blockIdx.x * 32;
        p[i] = clock64();
        if (i == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); p[31] = s; }
    }
}
#define AM_TL(i) tl_rec(i)
#define AM_TL_SET(tile)
#ifdef NOWAIT
#define AM_TL_WAIT(v, n)
#else
#define AM_TL_WAIT(v, n) do { float a_ = 0.f; _Pragma("unroll") for (int q_ = 0; q_ < (n); ++q_) a_ += (v)[q_].x + (v)[q_].y; if (a_ == 1.2345e38f) g_flag = 1; } while (0)
#endif
#include "../../audio_matcher_b200/csrc/am_kernels.cuh"
using namespace amk;
int main() {
    const int L1 = 9, L2 = 13, pairs = 32;
    const size_t N = (size_t)1 << 22;
    float2 *A, *S; cudaMalloc(&A, pairs * N * 8); cudaMalloc(&S, N * 8);
    cudaMemset(A, 0, pairs * N * 8); cudaMemset(S, 0, N * 8);
    std::vector<float2> tw(amfft::TW_N);
    for (int j = 0; j < amfft::TW_N; ++j) { double a = -2.0 * M_PI * j / amfft::TW_N; tw[j] = make_float2((float)cos(a), (float)sin(a)); }
    float2 *dtw; cudaMalloc(&dtw, tw.size() * 8); cudaMemcpy(dtw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice);
    typedef Row32Cfg<L2> Cfg;
    cudaFuncSetAttribute(k_row32<L2, ROW_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    const int rows = pairs << L1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 2; ++it) k_row32<L2, ROW_FUSED><<<rows, Cfg::THREADS, Cfg::SMEM>>>(A, S, nullptr, L1, rows, dtw);
    cudaEventRecord(e0);
    k_row32<L2, ROW_FUSED><<<rows, Cfg::THREADS, Cfg::SMEM>>>(A, S, nullptr, L1, rows, dtw);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("k_row32 %.1f us for %d rows (%s), smem %zu\n", ms * 1e3, rows, cudaGetErrorString(cudaGetLastError()), (size_t)Cfg::SMEM);
    std::vector<long long> tl(16384 * 32);
    cudaMemcpyFromSymbol(tl.data(), g_tl, tl.size() * 8);
    const int n = rows;
    double ph[5] = {0, 0, 0, 0, 0}, life = 0;
    for (int i = 0; i < n; ++i) { for (int k = 0; k < 5; ++k) ph[k] += tl[i * 32 + k + 1] - tl[i * 32 + k]; life += tl[i * 32 + 5] - tl[i * 32]; }
    printf("avg cycles per CTA (after load issue): load-wait %.0f  fwd %.0f  spectrum+mul %.0f  inv %.0f  store-issue %.0f  total %.0f\n", ph[0] / n, ph[1] / n, ph[2] / n, ph[3] / n, ph[4] / n, life / n);
    {   // stage detail: forward = slots 8.., inverse = slots 17..
        const char *nm[] = {"bfly0", "write0+sync", "read1+sync", "bfly1", "write1+sync", "read2+sync", "bfly2"};
        for (int dir = 0; dir < 2; ++dir) {
            printf(dir ? "  inverse:" : "  forward:");
            for (int k = 0; k < 7; ++k) { double a = 0; for (int i = 0; i < n; ++i) { long long prev = k == 0 ? tl[i * 32 + (dir ? 3 : 1)] : tl[i * 32 + 8 + dir * 9 + k - 1]; a += tl[i * 32 + 8 + dir * 9 + k] - prev; }
                printf(" %s %.0f", nm[k], a / n); }
            printf("\n");
        }
    }
    std::vector<std::pair<long long, int>> sm0;
    for (int i = 0; i < n; ++i) if (tl[i * 32 + 31] == 0) sm0.push_back({tl[i * 32], i});
    std::sort(sm0.begin(), sm0.end());
    for (size_t k = 0; k < sm0.size() && k < 12; ++k) { int i = sm0[k].second; long long b = sm0[0].first;
        printf("  SM0 cta %5d: issued %7lld  loaded %7lld  fwd %7lld  mul %7lld  inv %7lld  stored %7lld\n", i, tl[i*32]-b, tl[i*32+1]-b, tl[i*32+2]-b, tl[i*32+3]-b, tl[i*32+4]-b, tl[i*32+5]-b); }
    return 0;
}
