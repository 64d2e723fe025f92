// Ceiling of the inner loop: DMMA.8x8x4 with distinct operand registers, with and without the shared-memory
// fragment loads of the 32x32 warp tile (no global traffic, no barriers).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MT, int NT, bool LDS>
__global__ void __launch_bounds__(256, 2) mix(double* out, int iters) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g4 = lane >> 2, tq = lane & 3;
  for (int i = tid; i < 16 * 132 + 16 * 132; i += 256) sm[i] = 1e-3 * (i % 17);
  __syncthreads();
  double acc[MT][NT][2];
#pragma unroll
  for (int a = 0; a < MT; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  double a[MT], b[NT];
#pragma unroll
  for (int m = 0; m < MT; ++m) a[m] = 1e-3 * (m + lane);
#pragma unroll
  for (int n = 0; n < NT; ++n) b[n] = 1e-3 * (n - lane);
  const double* sA = sm; const double* sB = sm + 16 * 132;
  const int wi = warp & 3, wj = warp >> 2;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (LDS) {
#pragma unroll
        for (int m = 0; m < MT; ++m) a[m] = sB[(kk * 4 + tq) * 132 + ((wj * 8 * MT + m * 8 + g4) & 127)];
#pragma unroll
        for (int n = 0; n < NT; ++n) b[n] = sA[(kk * 4 + tq) * 132 + ((wi * 8 * NT + n * 8 + g4) & 127)];
      }
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) dmma(acc[m][n][0], acc[m][n][1], a[m], b[n]);
    }
  }
  double s = 0;
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) s += acc[m][n][0] + acc[m][n][1];
  out[blockIdx.x * 256 + tid] = s;
}
template <int MT, int NT, bool LDS> void run(const char* name, int ctas_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 2 * 256);
  const int iters = 4000; const int smem = 2 * 16 * 132 * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); mix<MT, NT, LDS><<<sms * ctas_per_sm, 256, smem>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  double fl = 2.0 * 256 * MT * NT * 4.0 * iters * 8 * sms * ctas_per_sm;
  printf("%-34s ctas/sm=%d  %.2f TFLOP/s\n", name, ctas_per_sm, fl / best * 1e-9);
  cudaFree(out);
}
int main() {
  run<4, 4, false>("4x4 regs only", 1); run<4, 4, false>("4x4 regs only", 2);
  run<4, 4, true>("4x4 + LDS frags (32x32 warp tile)", 1); run<4, 4, true>("4x4 + LDS frags (32x32 warp tile)", 2);
  run<8, 4, true>("8x4 + LDS frags (64x32 warp tile)", 1);
  run<4, 2, true>("4x2 + LDS frags", 2);
  return 0;
}
