// Issue/latency probes for the FP64 pipes of one SM (B200): DFMA and DMMA.8x8x4 throughput per scheduler as a function
// of the warps and independent accumulators in flight, dependent-chain latencies, shuffle / LDS round trips.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC> __global__ void dmma_tp(double* out, int iters, long long* cyc) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  const double a = 1e-3 + threadIdx.x * 1e-9, b = 1e-3;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int NACC> __global__ void dfma_tp(double* out, int iters, long long* cyc) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x + i;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-3;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void misc_lat(double* out, int iters, long long* cyc) {
  __shared__ double buf[64];
  double x = 1.5 + threadIdx.x * 1e-6;
  buf[threadIdx.x] = x; buf[32 + threadIdx.x] = x;
  __syncwarp();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) + 1e-9;     // shfl64 + dadd
  long long t1 = clock64();
  for (int i = 0; i < iters; ++i) { double y0; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x)); x = y0 + 1.0; }   // MUFU.RSQ64H + dadd
  long long t2 = clock64();
  for (int i = 0; i < iters; ++i) { buf[threadIdx.x] = x; __syncwarp(); x = buf[(threadIdx.x + 1) & 31] + 1e-9; __syncwarp(); }  // STS, LDS, dadd
  long long t3 = clock64();
  int idx = threadIdx.x;
  for (int i = 0; i < iters; ++i) { idx = (int)buf[idx & 31] & 31; }   // dependent LDS + cvt
  long long t4 = clock64();
  out[threadIdx.x] = x + idx;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64);
  long long h[8]; const int it = 2000;
#define RUN(K, NACC, T) { K<NACC><<<1, T>>>(out, it, cyc); cudaDeviceSynchronize(); K<NACC><<<1, T>>>(out, it, cyc); cudaDeviceSynchronize(); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-8s warps/SM %2d (per scheduler %d)  independent %2d : %.1f cycles per instruction per warp, %.2f cycles per instruction per scheduler\n", #K, T / 32, (T / 32 + 3) / 4, NACC, (double)h[0] / it / NACC, (double)h[0] / it / NACC / ((T / 32 + 3) / 4)); }
  RUN(dmma_tp, 1, 32) RUN(dmma_tp, 2, 32) RUN(dmma_tp, 4, 32) RUN(dmma_tp, 8, 32) RUN(dmma_tp, 16, 32)
  RUN(dmma_tp, 4, 128) RUN(dmma_tp, 8, 128) RUN(dmma_tp, 4, 256) RUN(dmma_tp, 8, 256) RUN(dmma_tp, 16, 256) RUN(dmma_tp, 4, 512) RUN(dmma_tp, 8, 512)
  RUN(dfma_tp, 1, 32) RUN(dfma_tp, 4, 32) RUN(dfma_tp, 16, 32) RUN(dfma_tp, 16, 128) RUN(dfma_tp, 16, 256) RUN(dfma_tp, 16, 512)
  misc_lat<<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize(); misc_lat<<<1, 32>>>(out, it, cyc); cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
  printf("shfl64+dadd %.1f   MUFU.RSQ64H+dadd %.1f   STS+syncwarp+LDS+dadd+syncwarp %.1f   dependent LDS+cvt %.1f cycles\n", (double)h[0] / it, (double)h[1] / it, (double)h[2] / it, (double)h[3] / it);
  return 0;
}
