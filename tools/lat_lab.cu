// Dependent-chain latencies (cycles) of the FP64 ops on the potrf critical path, one warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void chain(double* out, double x0, int n) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.000001;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (OP == 0) x = fma(x, y, 1e-9);
    else if (OP == 1) x = x * y;
    else if (OP == 2) x = rsqrt(x) + 1.0;
    else if (OP == 3) x = 1.0 / x + 0.5;
    else if (OP == 4) x = sqrt(x) + 1.0;
    else if (OP == 5) x = __drcp_rn(x) + 0.5;
    else if (OP == 6) { float f = rsqrtf((float)x); x = (double)f + 1.0; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / n;
}
__global__ void smem_roundtrip(double* out, int n) {
  __shared__ double buf[64];
  double x = threadIdx.x;
  buf[threadIdx.x] = x; __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { buf[threadIdx.x] = x; __syncthreads(); x = buf[(threadIdx.x + 1) & 31] + 1.0; __syncthreads(); }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / n;
}
int main() {
  double* out; cudaMalloc(&out, 65 * 8); double c;
  const char* names[] = {"dfma", "dmul", "rsqrt(double)+add", "1.0/x+add", "sqrt+add", "__drcp_rn+add", "rsqrtf via float +cvt+add"};
  for (int op = 0; op < 7; ++op) {
    for (int r = 0; r < 2; ++r) {
      switch (op) { case 0: chain<0><<<1,32>>>(out, 1.5, 4096); break; case 1: chain<1><<<1,32>>>(out, 1.5, 4096); break; case 2: chain<2><<<1,32>>>(out, 1.5, 4096); break;
        case 3: chain<3><<<1,32>>>(out, 1.5, 4096); break; case 4: chain<4><<<1,32>>>(out, 1.5, 4096); break; case 5: chain<5><<<1,32>>>(out, 1.5, 4096); break; case 6: chain<6><<<1,32>>>(out, 1.5, 4096); break; }
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&c, out + 64, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.1f cycles per dependent op\n", names[op], c);
  }
  for (int t : {32, 256}) { smem_roundtrip<<<1, t>>>(out, 1024); cudaDeviceSynchronize(); smem_roundtrip<<<1, t>>>(out, 1024); cudaDeviceSynchronize();
    cudaMemcpy(&c, out + 64, 8, cudaMemcpyDeviceToHost); printf("STS+bar+LDS+dadd+bar (%d thr)   %.1f cycles\n", t, c); }
  return 0;
}
