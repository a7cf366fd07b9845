// Probe: which cp.async.bulk.tensor window shapes does the B200 accept? (tools/jobs/r4d.sh)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template<int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float *out, int bytes, int c0, int c1, int c2, int c3)
{
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        if (RANK == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(smem_u32(sm)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(sm)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char **argv)
{
    int only = argc > 1 ? atoi(argv[1]) : -1, idx = -1;
    void *p = 0; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiled enc = (EncodeTiled)p;
    const int pitch = 64, rows = 20, planes = 22, members = 24;
    size_t n = (size_t)pitch * rows * planes * members;
    float *h = (float *)malloc(n * 4), *d, *out;
    for (size_t i = 0; i < n; ++i) h[i] = (float)i;
    cudaMalloc(&d, n * 4); cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice); cudaMalloc(&out, 1 << 16);
    // x0 = first element of the window along x: 31 * 4 bytes is not a multiple of 16
    struct { int rank, bx, by, x0; CUtensorMapDataType t; const char *name; } cases[] = {
        {4, 36, 14, 31, CU_TENSOR_MAP_DATA_TYPE_INT32, "4d 36x14 i32 x0=31"}, {4, 36, 14, 28, CU_TENSOR_MAP_DATA_TYPE_INT32, "4d 36x14 i32 x0=28"},
        {4, 40, 14, 28, CU_TENSOR_MAP_DATA_TYPE_INT32, "4d 40x14 i32 x0=28"}, {4, 40, 14, 30, CU_TENSOR_MAP_DATA_TYPE_INT32, "4d 40x14 i32 x0=30"},
        {3, 32, 14, 31, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, "3d 32x14 f32 x0=31"}, {3, 32, 14, 28, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, "3d 32x14 f32 x0=28"}};
    for (auto& c : cases) {
        if (++idx != only && only >= 0) continue;
        CUtensorMap map;
        cuuint64_t dims[4] = {pitch, rows, planes, members};
        cuuint64_t strides[3] = {pitch * 4, pitch * rows * 4, (cuuint64_t)pitch * rows * planes * 4};
        cuuint32_t box[4] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, 1, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&map, c.t, c.rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int bytes = c.bx * c.by * 4;
        if (c.rank == 4) probe<4><<<1, 128, 16384>>>(map, out, bytes, c.x0, 1, 2, 3);
        else probe<3><<<1, 128, 16384>>>(map, out, bytes, c.x0, 1, 2, 0);
        cudaError_t e = cudaDeviceSynchronize();
        float v[2] = {0, 0};
        if (e == cudaSuccess) cudaMemcpy(v, out, 8, cudaMemcpyDeviceToHost);
        printf("%s: encode %d, run %s, first %.0f (want %.0f)\n", c.name, (int)r, cudaGetErrorString(e), v[0],
               (float)(c.x0 + 1 * pitch + 2 * pitch * rows + (c.rank == 4 ? 3.0 * pitch * rows * planes : 0)));
        if (e != cudaSuccess) { printf("context lost, stopping\n"); return 1; }
    }
    return 0;
}
