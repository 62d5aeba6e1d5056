// micro-benchmark: which load flavour keeps the most 32-byte pieces in flight when shared memory crowds out L1
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __device__ __forceinline__ uint4 ld(const uint4 *p) {
    uint4 r;
    if (MODE == 0) r = __ldg(p);
    else if (MODE == 1) r = __ldcg(p);
    else if (MODE == 2) r = __ldcs(p);
    else if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (MODE == 4) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <int PIECE, int MODE>
__global__ void __launch_bounds__(512, 2) k_read(const uint4 *__restrict__ x, size_t row_stride_u4, int tiles_x, unsigned *sink) {
    constexpr int U4_PER_PIECE = PIECE / 16;
    constexpr int ROWS = 32768 / PIECE;
    extern __shared__ uint4 smx[];
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const uint4 *base = x + (size_t)ty * ROWS * row_stride_u4 + (size_t)tx * U4_PER_PIECE;
    unsigned acc = 0;
    if (MODE == 6) {
#pragma unroll
        for (int u = threadIdx.x; u < 2048; u += 512) {
            const int row = u / U4_PER_PIECE, c = u % U4_PER_PIECE;
            unsigned d = (unsigned)__cvta_generic_to_shared(smx + u);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(base + (size_t)row * row_stride_u4 + c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        uint4 v = smx[threadIdx.x];
        acc = v.x ^ v.y ^ v.z ^ v.w;
    } else {
#pragma unroll
        for (int u = threadIdx.x; u < 2048; u += 512) {
            const int row = u / U4_PER_PIECE, c = u % U4_PER_PIECE;
            uint4 v = ld<MODE>(base + (size_t)row * row_stride_u4 + c);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
        if (tiles_x < 0) smx[threadIdx.x].x = acc;
    }
    if (acc == 0x12345678u) *sink = acc;
}
template <int PIECE, int MODE> void run(const uint4 *x, size_t bytes, unsigned *sink, int smem) {
    cudaFuncSetAttribute(k_read<PIECE, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const size_t row_stride = 16384;
    const int tiles_x = row_stride / PIECE;
    const int rows = 32768 / PIECE;
    const size_t band = row_stride * rows;
    const int tiles_y = bytes / band;
    const int grid = tiles_x * tiles_y;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 2; ++it) k_read<PIECE, MODE><<<grid, 512, smem>>>(x, row_stride / 16, tiles_x, sink);
    cudaEventRecord(e0);
    for (int it = 0; it < 5; ++it) k_read<PIECE, MODE><<<grid, 512, smem>>>(x, row_stride / 16, tiles_x, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const char *names[] = {"ldg(nc)", "ldcg", "ldcs", "nc.no_allocate", "no_allocate", "volatile", "cp.async.cg"};
    printf("smem %6d piece %5d B %-15s: %.1f GB/s  (%s)\n", smem, PIECE, names[MODE], 5.0 * grid * 32768.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
template <int PIECE> void all(const uint4 *x, size_t bytes, unsigned *sink, int smem) {
    run<PIECE, 0>(x, bytes, sink, smem); run<PIECE, 1>(x, bytes, sink, smem); run<PIECE, 2>(x, bytes, sink, smem);
    run<PIECE, 3>(x, bytes, sink, smem); run<PIECE, 4>(x, bytes, sink, smem); run<PIECE, 5>(x, bytes, sink, smem);
    run<PIECE, 6>(x, bytes, sink, smem);
}
int main() {
    size_t bytes = (size_t)4 << 30;
    uint4 *x; unsigned *sink;
    cudaMalloc(&x, bytes); cudaMalloc(&sink, 4); cudaMemset(x, 1, bytes);
    for (int smem : {100000}) { all<32>(x, bytes, sink, smem); all<128>(x, bytes, sink, smem); }
    all<32>(x, bytes, sink, 40000);
    return 0;
}
