// micro-benchmark: TMA 2D tile loads of 32-byte-wide pieces (16 int16 columns x 1024 rows per CTA)
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 2) k_tma(const __grid_constant__ CUtensorMap tm, int tiles_x, unsigned *sink, int boxes, int box_rows) {
    extern __shared__ __align__(128) unsigned char smx[];
    __shared__ __align__(8) unsigned long long bar;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(boxes * box_rows * 32) : "memory");
        for (int b = 0; b < boxes; ++b) {
            unsigned dst = (unsigned)__cvta_generic_to_shared(smx + b * box_rows * 32);
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(dst), "l"(&tm), "r"(bar_a), "r"(tx * 16), "r"(ty * boxes * box_rows + b * box_rows) : "memory");
        }
    }
    unsigned ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar_a) : "memory");
    uint4 v = ((uint4 *)smx)[threadIdx.x];
    unsigned acc = v.x ^ v.y ^ v.z ^ v.w;
    if (acc == 0x12345678u) *sink = acc;
}
int main() {
    size_t bytes = (size_t)4 << 30;
    void *x; unsigned *sink;
    cudaMalloc(&x, bytes); cudaMalloc(&sink, 4); cudaMemset(x, 1, bytes);
    typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    enc_t enc = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qr);
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const cuuint64_t rows = bytes / 16384;
    for (int promo = 0; promo < 4; ++promo) for (int smem : {40000, 100000}) for (int box_rows : {256, 64}) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {8192, rows}; cuuint64_t strides[1] = {16384}; cuuint32_t box[2] = {16, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int boxes = 1024 / box_rows;
        cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const int tiles_x = 512, tiles_y = rows / 1024, grid = tiles_x * tiles_y;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int it = 0; it < 2; ++it) k_tma<<<grid, 512, smem>>>(tm, tiles_x, sink, boxes, box_rows);
        cudaEventRecord(e0);
        for (int it = 0; it < 5; ++it) k_tma<<<grid, 512, smem>>>(tm, tiles_x, sink, boxes, box_rows);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("TMA l2promo %d smem %6d box_rows %3d: %.1f GB/s (%s)\n", promo, smem, box_rows, 5.0 * grid * 32768.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
