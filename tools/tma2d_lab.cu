// Lab: one 2-D tensor-map TMA box load of FP64 data, checked against the host.  nvcc -arch=sm_100a tma2d_lab.cu -o tma2d_lab
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void lab(const __grid_constant__ CUtensorMap map, int c0, int c1, int n, double* out, uint32_t* info) {
    extern __shared__ __align__(128) double sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + n);
    if (threadIdx.x == 0) {
        info[0] = smem_u32(sm);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(sm)), "l"(&map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
    const int rows = 1024, cols = 256, ld = 1024;
    const int BR = argc > 1 ? atoi(argv[1]) : 132, BC = 16;
    const int c0 = argc > 2 ? atoi(argv[2]) : 128, c1 = argc > 3 ? atoi(argv[3]) : 32;
    std::vector<double> h((size_t)ld * cols);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
    double *d, *out; uint32_t* info;
    cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    const int n = BR * BC;
    cudaMalloc(&out, n * 8); cudaMalloc(&info, 16);
    PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)cols}; cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)BR, (cuuint32_t)BC}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d BR=%d c0=%d c1=%d\n", (int)r, BR, c0, c1);
    const int smem = n * 8 + 64;
    cudaFuncSetAttribute(lab, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    lab<<<1, 128, smem>>>(map, c0, c1, n, out, info);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e) return 1;
    std::vector<double> ho(n); uint32_t hi[4];
    cudaMemcpy(ho.data(), out, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hi, info, 16, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < BC; ++c) for (int rr = 0; rr < BR; ++rr) {
        const int gr = c0 + rr, gc = c1 + c;
        const double want = (gr < rows && gc < cols) ? h[(size_t)gc * ld + gr] : 0.0;
        if (ho[c * BR + rr] != want) { if (bad < 5) printf("mismatch c=%d r=%d got %g want %g\n", c, rr, ho[c * BR + rr], want); ++bad; }
    }
    printf("smem base 0x%x, mismatches %d\n", hi[0], bad);
    return 0;
}
