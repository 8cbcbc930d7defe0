// tma_probe.cu — stand-alone check of the tensor-map configurations k_inter_tma uses (one configuration per process: an illegal
// instruction poisons the context).  usage: tma_probe <cfg>   cfg 0: 2-D box 32x21, 1: 3-D luma map, 2: 4-D chroma map, 3: 3-D with expect_tx after issue
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_probe(const __grid_constant__ CUtensorMap map, int rank, int x, int y, int z, int w, uint8_t *out, int bytes) {
    __shared__ alignas(128) uint8_t tile[2048];
    __shared__ unsigned long long bar;
    const uint32_t b = saddr(&bar), d = saddr(tile);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" :: "r"(b), "r"(bytes) : "memory");
        const unsigned long long m = (unsigned long long)&map;
        if (rank == 2) asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" :: "r"(d), "l"(m), "r"(x), "r"(y), "r"(b) : "memory");
        if (rank == 3) asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" :: "r"(d), "l"(m), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
        if (rank == 4) asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" :: "r"(d), "l"(m), "r"(x), "r"(y), "r"(z), "r"(w), "r"(b) : "memory");
    }
    int ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.s32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char **argv) {
    const int cfg = argc > 1 ? atoi(argv[1]) : 0;
    const int xarg = argc > 2 ? atoi(argv[2]) : -1;
    const uint64_t W = 1920, H = 1088, NS = 34, FB = W * H * 3 / 2;
    uint8_t *buf; cudaMalloc(&buf, FB * NS + 4096);
    std::vector<uint8_t> h(FB * NS);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 24);
    cudaMemcpy(buf, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    PFN_cuTensorMapEncodeTiled enc = (PFN_cuTensorMapEncodeTiled)fn;
    CUtensorMap m; CUresult r; int rank, bytes; int x = 37, y = 55, z = 3, w = 0;
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (cfg == 0) { const cuuint64_t d[2] = {W, H * NS}, s[1] = {W}; const cuuint32_t b[2] = {32, 21}; rank = 2; bytes = 672;
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    else if (cfg == 1 || cfg == 3) { const cuuint64_t d[3] = {W, H, NS}, s[2] = {W, FB}; const cuuint32_t b[3] = {32, 21, 1}; rank = 3; bytes = 672;
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, buf, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    else if (cfg == 10) { const cuuint64_t d[2] = {W, H * NS}, s[1] = {W}; const cuuint32_t b[2] = {64, 16}; rank = 2; bytes = 1024;
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    else if (cfg == 11) { const cuuint64_t d[2] = {W / 4, H * NS}, s[1] = {W}; const cuuint32_t b[2] = {16, 16}; rank = 2; bytes = 1024; x = 8;
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    else if (cfg == 12) { const cuuint64_t d[2] = {W, H * NS}, s[1] = {W}; const cuuint32_t b[2] = {32, 21}; rank = 2; bytes = 672; x = 32; y = 0;
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    else { const cuuint64_t d[4] = {W / 2, H / 2, 2, NS}, s[3] = {W / 2, W / 2 * H / 2, FB}; const cuuint32_t b[4] = {16, 9, 2, 1}; rank = 4; bytes = 288; z = 0; w = 3;
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, buf + W * H, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
    if (xarg >= 0) x = xarg;
    { cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0); int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt); printf("device %s cc %d.%d driver %d runtime %d\n", pr.name, pr.major, pr.minor, drv, rt); }
    printf("cfg %d encode %d\n", cfg, (int)r);
    uint8_t *out; cudaMalloc(&out, 4096);
    k_probe<<<1, 64>>>(m, rank, x, y, z, w, out, bytes);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cfg %d kernel: %s\n", cfg, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint8_t> o(bytes); cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        if (rank != 4) { for (int r2 = 0; r2 < 21; r2++) for (int c = 0; c < 32; c++) { size_t src = (rank == 2 ? 0 : (size_t)z * FB) + (size_t)(y + r2) * W + x + c; bad += o[r2 * 32 + c] != h[src]; } }
        else { for (int p = 0; p < 2; p++) for (int r2 = 0; r2 < 9; r2++) for (int c = 0; c < 16; c++) { size_t src = (size_t)w * FB + W * H + (size_t)p * (W / 2) * (H / 2) + (size_t)(y + r2) * (W / 2) + x + c; bad += o[(p * 9 + r2) * 16 + c] != h[src]; } }
        printf("cfg %d mismatches %d\n", cfg, bad);
    }
    return 0;
}
