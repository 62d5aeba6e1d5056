// Timeline micro-benchmark of k_row32_stream<13>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__device__ long long g_tl[16384 * 32];
__device__ int g_cur[2048];
__device__ __forceinline__ void tl_rec(int i) {
    if (threadIdx.x == 0 && i < 8) {
        long long *p = g_tl + (size_t)g_cur[blockIdx.x] * 32;
        p[i] = clock64();
        if (i == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); p[31] = s; p[30] = blockIdx.x; }
    }
}
#define AM_TL(i) tl_rec(i)
#define AM_TL_SET(tile) do { if (threadIdx.x == 0) g_cur[blockIdx.x] = (tile); } while (0)
#define AM_TL_WAIT(v, n)
#include "../../audio_matcher_b200/csrc/am_kernels.cuh"
using namespace amk;
int main(int argc, char **argv) {
    const int L1 = 9, L2 = 13, pairs = 32;
    const size_t N = (size_t)1 << 22;
    float2 *A, *S; cudaMalloc(&A, pairs * N * 8); cudaMalloc(&S, N * 8);
    cudaMemset(A, 0, pairs * N * 8); cudaMemset(S, 0, N * 8);
    std::vector<float2> tw(amfft::TW_N);
    for (int j = 0; j < amfft::TW_N; ++j) { double a = -2.0 * M_PI * j / amfft::TW_N; tw[j] = make_float2((float)cos(a), (float)sin(a)); }
    float2 *dtw; cudaMalloc(&dtw, tw.size() * 8); cudaMemcpy(dtw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice);
    int *ctr; cudaMalloc(&ctr, 4);
    typedef Row32Cfg<L2> Cfg;
    cudaFuncSetAttribute(k_row32_stream<L2, ROW_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    const int rows = pairs << L1, ctas = argc > 1 ? atoi(argv[1]) : 296;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 2; ++it) { cudaMemset(ctr, 0, 4); k_row32_stream<L2, ROW_FUSED><<<ctas, Cfg::THREADS, Cfg::SMEM>>>(A, S, nullptr, L1, rows, dtw, ctr, 1, 0, 0); }
    cudaMemset(ctr, 0, 4);
    cudaEventRecord(e0);
    k_row32_stream<L2, ROW_FUSED><<<ctas, Cfg::THREADS, Cfg::SMEM>>>(A, S, nullptr, L1, rows, dtw, ctr, 1, 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("k_row32_stream %.1f us for %d rows (%s)\n", ms * 1e3, rows, cudaGetErrorString(cudaGetLastError()));
    std::vector<long long> tl(16384 * 32);
    cudaMemcpyFromSymbol(tl.data(), g_tl, tl.size() * 8);
    const int n = rows;
    const char *nm[] = {"wait-row", "copy-out+sync", "forward", "wait-spec", "mul+sync", "inverse", "store-issue"};
    double tot = 0;
    for (int k = 0; k < 7; ++k) { double a = 0; for (int i = 0; i < n; ++i) a += tl[i * 32 + k + 1] - tl[i * 32 + k]; printf("  %s %.0f", nm[k], a / n); tot += a / n; }
    printf("  total %.0f\n", tot);
    // rows per CTA distribution
    std::vector<int> per(ctas, 0); for (int i = 0; i < n; ++i) per[tl[i * 32 + 30]]++;
    std::sort(per.begin(), per.end());
    printf("rows per CTA: min %d median %d max %d\n", per[0], per[ctas / 2], per[ctas - 1]);
    std::vector<std::pair<long long, int>> sm0;
    for (int i = 0; i < n; ++i) if (tl[i * 32 + 31] == 0) sm0.push_back({tl[i * 32], i});
    std::sort(sm0.begin(), sm0.end());
    for (size_t k = 0; k < sm0.size() && k < 12; ++k) { int i = sm0[k].second; long long b = sm0[0].first;
        printf("  SM0 cta %3lld row %5d:", tl[i*32+30], i); for (int q = 0; q < 8; ++q) printf(" %7lld", tl[i*32+q]-b); printf("\n"); }
    float2 *Bbuf; cudaMalloc(&Bbuf, pairs * N * 8);
    
    cudaFuncSetAttribute(k_row32_stream<L2, ROW_INVERSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    
    
    for (int it = 0; it < 2; ++it) { cudaMemset(ctr, 0, 4); k_row32_stream<L2, ROW_INVERSE><<<ctas, Cfg::THREADS, Cfg::SMEM>>>(A, S, Bbuf, L1, rows, dtw, ctr, 1, 0, 0); }
    cudaMemset(ctr, 0, 4);
    cudaEventRecord(e0);
    k_row32_stream<L2, ROW_INVERSE><<<ctas, Cfg::THREADS, Cfg::SMEM>>>(A, S, Bbuf, L1, rows, dtw, ctr, 1, 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
     cudaEventElapsedTime(&ms, e0, e1);
    printf("k_row32_stream<ROW_INVERSE> %.1f us for %d rows (%s)\n", ms * 1e3, rows, cudaGetErrorString(cudaGetLastError()));
    
    cudaMemcpyFromSymbol(tl.data(), g_tl, tl.size() * 8);
    
    const char *nm2[] = {"wait-row", "copy-out+sync", "forward", "wait-spec", "mul+sync", "inverse", "store-issue"};
    tot = 0;
    for (int k = 0; k < 7; ++k) { double a = 0; for (int i = 0; i < n; ++i) a += tl[i * 32 + k + 1] - tl[i * 32 + k]; printf("  %s %.0f", nm2[k], a / n); tot += a / n; }
    printf("  total %.0f\n", tot);
    // rows per CTA distribution
    per.assign(ctas, 0); for (int i = 0; i < n; ++i) per[tl[i * 32 + 30]]++;
    std::sort(per.begin(), per.end());
    printf("rows per CTA: min %d median %d max %d\n", per[0], per[ctas / 2], per[ctas - 1]);
    sm0.clear();
    for (int i = 0; i < n; ++i) if (tl[i * 32 + 31] == 0) sm0.push_back({tl[i * 32], i});
    std::sort(sm0.begin(), sm0.end());
    for (size_t k = 0; k < sm0.size() && k < 12; ++k) { int i = sm0[k].second; long long b = sm0[0].first;
        printf("  SM0 cta %3lld row %5d:", tl[i*32+30], i); for (int q = 0; q < 8; ++q) printf(" %7lld", tl[i*32+q]-b); printf("\n"); }
    return 0;
}
