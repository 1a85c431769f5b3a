// tma_swizzle_probe.cu -- where does a SWIZZLE_128B tensor-map box land when its shared-memory destination is
// 128-byte but not 1024-byte aligned, and are negative box coordinates zero-filled?  (Decides the layout of the
// frame-major half-band input tiles.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_swizzle_probe tools/tma_swizzle_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*PFN_enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                            const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, int k, int row0, uint32_t *out) {
    extern __shared__ __align__(1024) uint32_t sm[];
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 0xdeadbeefu;
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;");
    if (threadIdx.x == 0) {
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm) + 128u * k;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4 * 128));
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(&map), "r"(0), "r"(row0), "r"(b) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = sm[i];
}

int main() {
    PFN_enc enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &q);
    const int rows = 64, cols = 32;
    uint32_t *g, *o;
    cudaMalloc(&g, rows * cols * 4);
    cudaMalloc(&o, 2048 * 4);
    uint32_t h[rows * cols];
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) h[r * cols + c] = (r + 1) * 100 + c;  // row r: (r+1)*100 + col
    cudaMemcpy(g, h, sizeof(h), cudaMemcpyHostToDevice);
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 4};
    cuuint32_t box[2] = {32, 4};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    uint32_t res[2048];
    for (int k : {0, 1, 2, 3, 5, 9}) {
        for (int row0 : {8, -1}) {
            probe<<<1, 128, 8192>>>(m, k, row0, o);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("k=%d row0=%d: %s\n", k, row0, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(res, o, sizeof(res), cudaMemcpyDeviceToHost);
            printf("dst = base + %d*128, box rows %d..%d:\n", k, row0, row0 + 3);
            for (int line = k; line < k + 4; line++) {  // 128-byte line `line` of shared memory
                printf("  smem line %2d (addr bits 7-9 = %d): chunks hold source (row, col/4):", line, line & 7);
                for (int c = 0; c < 8; c++) {
                    uint32_t v = res[line * 32 + c * 4];
                    if (v == 0xdeadbeefu) printf("  ----");
                    else printf("  %2d,%d", (int)(v / 100) - 1, (int)(v % 100) / 4);
                }
                printf("\n");
            }
        }
    }
    return 0;
}
