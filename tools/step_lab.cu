// How expensive is one "publish -> barrier -> consume" step on one SM?  Synthetic version of the potrf_diag inner
// step: 128 dependent steps of { LDS broadcast, (rsqrt), NMUL dmul, NFMA dfma, STS publish, __syncthreads }.
#include <cstdio>
#include <cuda_runtime.h>
template <int NFMA, int NLDS, bool RSQRT>
__global__ void __launch_bounds__(1024, 1) lab(double* out, int steps) {
  __shared__ double buf[2][256];
  const int tid = threadIdx.x;
  double m[NFMA > 0 ? NFMA : 1];
#pragma unroll
  for (int i = 0; i < (NFMA > 0 ? NFMA : 1); ++i) m[i] = 1.0 + 1e-3 * (tid + i);
  if (tid < 256) { buf[0][tid] = 2.0 + tid * 1e-3; buf[1][tid] = 2.0; }
  __syncthreads();
  long long t0 = clock64();
  for (int j = 0; j < steps; ++j) {
    const double* b = buf[j & 1];
    double d = b[j & 127];
    double inv = RSQRT ? rsqrt(d) : d * 0.5;
    double l[NLDS > 0 ? NLDS : 1];
#pragma unroll
    for (int i = 0; i < NLDS; ++i) l[i] = b[(tid + 16 * i) & 255] * inv;
#pragma unroll
    for (int i = 0; i < NFMA; ++i) m[i] = fma(-l[i % (NLDS > 0 ? NLDS : 1)], inv, m[i]);
    if ((tid >> 4) == (j & 15)) buf[(j + 1) & 1][(tid & 15) + 16 * (j & 7)] = m[0] * 1e-6 + 2.0;
    __syncthreads();
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < (NFMA > 0 ? NFMA : 1); ++i) s += m[i];
  out[tid] = s;
  if (tid == 0) out[1024] = (double)(t1 - t0) / steps;
}
template <int NFMA, int NLDS, bool RSQRT> void run(int threads) {
  double* out; cudaMalloc(&out, 1025 * 8);
  lab<NFMA, NLDS, RSQRT><<<1, threads>>>(out, 128); cudaDeviceSynchronize();
  lab<NFMA, NLDS, RSQRT><<<1, threads>>>(out, 128); cudaDeviceSynchronize();
  double cyc; cudaMemcpy(&cyc, out + 1024, 8, cudaMemcpyDeviceToHost);
  printf("threads=%4d NFMA=%2d NLDS=%2d rsqrt=%d : %.0f cycles/step\n", threads, NFMA, NLDS, (int)RSQRT, cyc);
  cudaFree(out);
}
int main() {
  for (int t : {32, 128, 256, 512}) { run<0, 0, false>(t); }
  for (int t : {128, 256, 512}) { run<0, 0, true>(t); run<0, 8, true>(t); run<0, 16, true>(t); run<16, 16, true>(t); run<36, 16, true>(t); run<36, 16, false>(t); }
  return 0;
}
