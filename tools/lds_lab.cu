// Lab: shared-memory load throughput per SM for broadcast-style reads (what the factorisation kernels do):
// cycles per warp-level LDS.32 / LDS.64 / LDS.128 with 1..16 warps resident, all lanes reading 16 distinct or 2
// distinct addresses.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 lds_lab.cu -o lds_lab.bin
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int MODE> __global__ void lds(double* out, long long* cyc, int iters) {
  extern __shared__ __align__(16) unsigned char smraw[];
  T* sm = reinterpret_cast<T*>(smraw);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<float*>(smraw)[i] = 1.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int base = (MODE == 0) ? (lane & 15) : (lane >> 4);   // 16 distinct consecutive / 2 distinct addresses
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      T v = sm[base + 16 * u + (it & 1)];
      if constexpr (sizeof(T) == 16) acc += reinterpret_cast<double2&>(v).x; else if constexpr (sizeof(T) == 8) acc += reinterpret_cast<double&>(v); else acc += reinterpret_cast<float&>(v);
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 256;
  for (int warps : {1, 2, 4, 8, 16}) {
    long long c[6];
    lds<float, 0><<<1, warps * 32, 16384>>>(out, cyc, iters); cudaMemcpy(&c[0], cyc, 8, cudaMemcpyDeviceToHost);
    lds<double, 0><<<1, warps * 32, 16384>>>(out, cyc, iters); cudaMemcpy(&c[1], cyc, 8, cudaMemcpyDeviceToHost);
    lds<double2, 0><<<1, warps * 32, 16384>>>(out, cyc, iters); cudaMemcpy(&c[2], cyc, 8, cudaMemcpyDeviceToHost);
    lds<double, 1><<<1, warps * 32, 16384>>>(out, cyc, iters); cudaMemcpy(&c[3], cyc, 8, cudaMemcpyDeviceToHost);
    lds<double2, 1><<<1, warps * 32, 16384>>>(out, cyc, iters); cudaMemcpy(&c[4], cyc, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 16 * warps;   // warp-level load instructions issued by the CTA
    printf("%2d warps: cycles per warp-LDS (SM-wide)  LDS.32 %.2f | LDS.64 %.2f | LDS.128 %.2f | LDS.64 2-addr %.2f | LDS.128 2-addr %.2f   (%s)\n",
           warps, c[0] / n, c[1] / n, c[2] / n, c[3] / n, c[4] / n, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
