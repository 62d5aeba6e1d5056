// Timeline micro-benchmark of k_col_fwd<9,4,16> and (with an argument = CTA count, e.g. 296) of the persistent
// k_col_fwd_stream<9,4,32,mono16>: per-CTA / per-tile phase timestamps (clock64 of warp 0) and SM ids.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ long long g_tl[16384 * 8];
__device__ int g_cur[2048];
__device__ int g_stream_mode;
#define AM_TL_SET(tile) do { if (threadIdx.x == 0) g_cur[blockIdx.x] = (tile); } while (0)
__device__ __forceinline__ void tl_rec(int i) {
    if (threadIdx.x == 0 && i < 6) {
        long long *p = g_tl + (size_t)(g_stream_mode ? g_cur[blockIdx.x] : blockIdx.y * gridDim.x + blockIdx.x) * 8;
        p[i] = clock64();
        if (i == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); p[7] = s; unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p[6] = (long long)t; }
    }
}
#define AM_TL(i) tl_rec(i)
#define AM_TL_WAIT(v, n)
#include "../../audio_matcher_b200/csrc/am_kernels.cuh"
using namespace amk;
int main(int argc, char **argv) {
    constexpr int L1 = 9, L2 = 13, pairs = 32, ES = 32;     // ES: elements per thread of the streaming kernel
    const long long N = 1ll << 22, m = 480000, VN = ((N - m + 1) / 32) * 32;
    const long long frames = 2 * pairs * VN + N + 64;
    short *pcm; cudaMalloc(&pcm, frames * 2); cudaMemset(pcm, 3, frames * 2);
    float2 *A; cudaMalloc(&A, (size_t)pairs * N * 8);
    std::vector<float2> tw(amfft::TW_N);
    for (int j = 0; j < amfft::TW_N; ++j) { double a = -2.0 * M_PI * j / amfft::TW_N; tw[j] = make_float2((float)cos(a), (float)sin(a)); }
    float2 *dtw; cudaMalloc(&dtw, tw.size() * 8); cudaMemcpy(dtw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice);
    BlockGroup g{};
    g.sv.x = pcm; g.sv.fmt = FMT_I16_MONO; g.sv.buf_first = 0; g.sv.buf_frames = frames; g.sv.total = frames; g.sv.lead = 0;
    g.g0 = 0; g.g_end = 2 * pairs * VN; g.VN = VN; g.nblocks = 2 * pairs; g.c = nullptr; g.c_g0 = 0; g.scalar = 1.f; g.rsum = nullptr; g.theta = 0.f;
    typedef ColCfg<L1, 4, 16> Cfg;
    cudaFuncSetAttribute(k_col_fwd<L1, 4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    dim3 grid((1 << L2) >> 4, pairs);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 2; ++it) k_col_fwd<L1, 4, 16><<<grid, Cfg::THREADS, Cfg::SMEM>>>(g, L2, A, dtw);
    cudaEventRecord(e0);
    k_col_fwd<L1, 4, 16><<<grid, Cfg::THREADS, Cfg::SMEM>>>(g, L2, A, dtw);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("kernel %.1f us (%s), smem %zu\n", ms * 1e3, cudaGetErrorString(cudaGetLastError()), (size_t)Cfg::SMEM);
    if (argc > 1) {
        int one = 1; cudaMemcpyToSymbol(g_stream_mode, &one, 4);
        typedef ColStreamCfg<L1, 4, ES, FMT_I16_MONO> SC;
        cudaFuncSetAttribute(k_col_fwd_stream<L1, 4, ES, FMT_I16_MONO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC::SMEM);
        const int ntiles = grid.x * grid.y, ctas = atoi(argv[1]);
        typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        enc_t enc = nullptr; cudaDriverEntryPointQueryResult qr;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qr);
        TensorMap tm; int *dctr; cudaMalloc(&dctr, 4);
        cuuint64_t dims[2] = {(cuuint64_t)2 << L2, (cuuint64_t)(frames >> L2) + 1}; cuuint64_t strides[1] = {(cuuint64_t)2 << L2};
        cuuint32_t box[2] = {16, 256}, es[2] = {1, 1};
        CUresult r = enc((CUtensorMap *)&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, pcm, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d\n", (int)r);
        for (int it = 0; it < 2; ++it) (cudaMemset(dctr, 0, 4), k_col_fwd_stream<L1, 4, ES, FMT_I16_MONO><<<ctas, SC::THREADS, SC::SMEM>>>(tm, g, L2, A, dtw, ntiles, dctr));
        cudaMemset(dctr, 0, 4);
        cudaEventRecord(e0);
        (cudaMemset(dctr, 0, 4), k_col_fwd_stream<L1, 4, ES, FMT_I16_MONO><<<ctas, SC::THREADS, SC::SMEM>>>(tm, g, L2, A, dtw, ntiles, dctr));
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("stream kernel %.1f us (%s), smem %zu, ctas %d\n", ms * 1e3, cudaGetErrorString(cudaGetLastError()), (size_t)SC::SMEM, ctas);
    }
    std::vector<long long> tl(16384 * 8);
    cudaMemcpyFromSymbol(tl.data(), g_tl, tl.size() * 8);
    const int n = grid.x * grid.y;
    double ph[4] = {0, 0, 0, 0}, life = 0;
    for (int i = 0; i < n; ++i) { for (int k = 0; k < 4; ++k) ph[k] += tl[i * 8 + k + 1] - tl[i * 8 + k]; life += tl[i * 8 + 4] - tl[i * 8]; }
    if (argc > 1) { double w = 0, c = 0; for (int i = 0; i < n; ++i) { w += tl[i * 8 + 5] - tl[i * 8]; c += tl[i * 8 + 1] - tl[i * 8 + 5]; }
        printf("stream: wait %.0f  consume+prefetch-issue %.0f\n", w / n, c / n); }
    printf("avg cycles per CTA: load %.0f  fft %.0f  twiddle %.0f  store-issue %.0f  total %.0f\n", ph[0] / n, ph[1] / n, ph[2] / n, ph[3] / n, life / n);
    // per-SM: CTAs, busy span by globaltimer
    long long t0 = tl[6]; for (int i = 0; i < n; ++i) t0 = std::min(t0, tl[i * 8 + 6]);
    long long tmax = 0; for (int i = 0; i < n; ++i) tmax = std::max(tmax, tl[i * 8 + 6]);
    printf("first-to-last CTA start: %.1f us; CTAs per SM ~ %.1f\n", (tmax - t0) / 1e3, n / 148.0);
    // start-time histogram of CTA starts in 5 us bins
    std::vector<int> hist((tmax - t0) / 5000 + 1, 0);
    for (int i = 0; i < n; ++i) hist[(tl[i * 8 + 6] - t0) / 5000]++;
    printf("CTA starts per 5 us: "); for (size_t b = 0; b < hist.size() && b < 40; ++b) printf("%d ", hist[b]); printf("\n");
    // timeline of SM 0: starts and phase ends relative to first start (cycles)
    std::vector<std::pair<long long, int>> sm0;
    for (int i = 0; i < n; ++i) if (tl[i * 8 + 7] == 0) sm0.push_back({tl[i * 8], i});
    std::sort(sm0.begin(), sm0.end());
    for (size_t k = 0; k < sm0.size() && k < 16; ++k) { int i = sm0[k].second; long long b = sm0[0].first;
        printf("  SM0 cta %5d: start %7lld  loaded %7lld  fft %7lld  tw %7lld  stored %7lld\n", i, tl[i*8]-b, tl[i*8+1]-b, tl[i*8+2]-b, tl[i*8+3]-b, tl[i*8+4]-b); }
    return 0;
}
