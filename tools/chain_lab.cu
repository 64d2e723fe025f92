// Dependent-chain cost of one pivot step of potrf_diag7's chain warp, piece by piece (one warp, cycles per iteration).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rs(double x) {
  double y0; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double t = y0 * y0; const double e = fma(-t, x, 1.0); const double p = fma(e, 0.375, 0.5); const double q = y0 * e;
  return fma(p, q, y0);
}
template <int V> __global__ void k(double* out, int n, long long* cyc) {
  __shared__ double buf[64];
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) { __syncthreads(); return; }   // the other warps of the CTA wait at the barrier
  double a0 = 2.0 + lane * 1e-3, a1 = 3.0 + lane * 1e-3, s = 0.7, d = 2.0;
  buf[lane] = 0.001; buf[32 + lane] = 0.001;
  __syncwarp();
  long long t0 = clock64();
  for (int j = 0; j < n; ++j) {
    const double l = a0 * s;
    const double dn = fma(-l, l, a1);
    double dnext;
    if (V == 0) dnext = __shfl_sync(0xffffffffu, dn, (j + 1) & 31);                 // full step
    else if (V == 1) dnext = dn;                                                     // no shuffle
    else if (V == 2) dnext = __shfl_sync(0xffffffffu, dn, (j + 1) & 31);            // shuffle, cheap rsqrt below
    else if (V == 3) dnext = __shfl_sync(0xffffffffu, dn, (j + 1) & 31);
    else dnext = dn;
    if (V == 3 || V == 0) { buf[lane] = l; __syncwarp(); }
    if (V == 2) { double y0; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(dnext)); s = y0; }
    else if (V == 4) s = dnext * 0.25;                                               // only mul + fma + mul
    else s = rs(dnext);
    d = dnext;
    if (V == 3 || V == 0) { const double c = buf[(lane + 1) & 31]; a0 = fma(-l, c, a1); a1 = fma(-l, c, 3.0); }
    else { a0 = a1 * 0.999 + 1.0; }
  }
  long long t1 = clock64();
  out[lane] = a0 + a1 + s + d;
  if (lane == 0) *cyc = t1 - t0;
  __syncthreads();
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 8); long long h; const int n = 4096;
  const char* names[] = {"mul fma shfl64 sts/lds rsqrt(full)", "mul fma rsqrt(full), no shuffle", "mul fma shfl64 MUFU only", "same as 0", "mul fma mul (no shuffle, no rsqrt)"};
#define RUN(V) k<V><<<1, 32>>>(out, n, cyc); cudaDeviceSynchronize(); k<V><<<1, 32>>>(out, n, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-40s %.1f cycles per step\n", names[V], (double)h / n);
  RUN(0) RUN(1) RUN(2) RUN(4)
#undef RUN
#define RUN(V) k<V><<<1, 288>>>(out, n, cyc); cudaDeviceSynchronize(); k<V><<<1, 288>>>(out, n, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("9 warps, 8 at the barrier: %-40s %.1f cycles per step\n", names[V], (double)h / n);
  RUN(0) RUN(1) RUN(2) RUN(4)
  return 0;
}
