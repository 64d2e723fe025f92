// FP64 peak probe for B200 (sm_100a): DMMA.8x8x4 / DFMA issue rates, exp() rate, and the
// library comparators (cuBLAS dgemm/dsyrk, cuSOLVER dpotrf) that serve as roofline denominators.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu -lcublas -lcusolver
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_rate(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_rate(double* out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x + i;
  double a = a0, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void exp_rate(double* out, int iters, double x0) {
  double x = x0 - threadIdx.x * 1e-3, s = 0;
  for (int it = 0; it < iters; ++it) { s += exp(x); x -= 1e-4; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void clock_chain(long long* out, int iters) {
  double c0 = 1.0, c1 = 2.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) dmma(c0, c1, 1e-3, 1e-3);
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)(c0 + c1); }
}

template <typename F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 32 * 1024));
  const int iters = 20000;
  // DMMA with different resident warp counts (threads per block, blocks per SM=1)
  int tpb[] = {128, 256, 512, 1024};
  for (int t : tpb) {
    float ms = time_ms([&] { dmma_rate<8><<<sms, t>>>(out, iters, 1e-3, 1e-3); }, 3);
    double fl = 2.0 * 256 * 8 * (double)iters * (t / 32) * sms;
    printf(" \"dmma_tflops_warps%d\": %.3f,\n", t / 32, fl / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { dmma_rate<16><<<sms * 2, 256>>>(out, iters, 1e-3, 1e-3); }, 3);
    double fl = 2.0 * 256 * 16 * (double)iters * 8 * sms * 2;
    printf(" \"dmma_tflops_2cta_x8warps_16acc\": %.3f,\n", fl / ms * 1e-9);
  }
  for (int t : tpb) {
    float ms = time_ms([&] { dfma_rate<16><<<sms, t>>>(out, iters, 0.999, 1e-3); }, 3);
    double fl = 2.0 * 16 * (double)iters * t * sms;
    printf(" \"dfma_tflops_warps%d\": %.3f,\n", t / 32, fl / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { exp_rate<<<sms * 8, 256>>>(out, 4000, -1.0); }, 3);
    printf(" \"exp_gops\": %.3f,\n", 4000.0 * 256 * sms * 8 / ms * 1e-6);
  }
  {
    long long* d; CK(cudaMalloc(&d, 16)); long long h[2];
    clock_chain<<<1, 32>>>(d, 4096); CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf(" \"dmma_dependent_latency_cycles\": %.2f,\n", (double)h[0] / 4096);
  }
  // Library comparators
  cublasHandle_t hb; cublasCreate(&hb);
  {
    int n = 8192; double *A, *B, *C; size_t sz = sizeof(double) * n * n;
    CK(cudaMalloc(&A, sz)); CK(cudaMalloc(&B, sz)); CK(cudaMalloc(&C, sz));
    CK(cudaMemset(A, 0, sz)); CK(cudaMemset(B, 0, sz)); CK(cudaMemset(C, 0, sz));
    double al = 1, be = 0;
    float ms = time_ms([&] { cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, A, n, B, n, &be, C, n); }, 5);
    printf(" \"cublas_dgemm_nt_8192_tflops\": %.3f,\n", 2.0 * n * n * n / ms * 1e-9);
    // sustained: 20 back to back
    float ms2 = time_ms([&] { for (int i = 0; i < 20; ++i) cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, A, n, B, n, &be, C, n); }, 1);
    printf(" \"cublas_dgemm_nt_8192_sustained_tflops\": %.3f,\n", 20 * 2.0 * n * n * n / ms2 * 1e-9);
    // syrk rank-128 update of a 9088 matrix (the shape of one trailing update)
    int m = 8192, k = 128; al = -1; be = 1;
    ms = time_ms([&] { cublasDsyrk(hb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, m, k, &al, A, m, &be, C, m); }, 5);
    printf(" \"cublas_dsyrk_8192_k128_tflops\": %.3f,\n", 1.0 * m * m * k / ms * 1e-9);
    k = 256;
    ms = time_ms([&] { cublasDsyrk(hb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, m, k, &al, A, m, &be, C, m); }, 5);
    printf(" \"cublas_dsyrk_8192_k256_tflops\": %.3f,\n", 1.0 * m * m * k / ms * 1e-9);
    ms = time_ms([&] { cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, m, m, 128, &al, A, m, B, m, &be, C, m); }, 5);
    printf(" \"cublas_dgemm_nt_8192_k128_tflops\": %.3f,\n", 2.0 * m * m * 128 / ms * 1e-9);
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  {
    cusolverDnHandle_t hs; cusolverDnCreate(&hs);
    int ns[] = {2048, 4096, 9000, 16384};
    for (int n : ns) {
      double* A; size_t sz = sizeof(double) * (size_t)n * n; CK(cudaMalloc(&A, sz));
      std::vector<double> h((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) h[(size_t)j * n + i] = (i == j) ? 2.0 : 1.0 / (1.0 + (i - j) * (i - j));
      int lw; cusolverDnDpotrf_bufferSize(hs, CUBLAS_FILL_MODE_LOWER, n, A, n, &lw);
      double* w; CK(cudaMalloc(&w, sizeof(double) * lw)); int* info; CK(cudaMalloc(&info, 4));
      float best = 1e30f;
      for (int r = 0; r < 3; ++r) {
        CK(cudaMemcpy(A, h.data(), sz, cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        cusolverDnDpotrf(hs, CUBLAS_FILL_MODE_LOWER, n, A, n, w, lw, info);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      int hi; cudaMemcpy(&hi, info, 4, cudaMemcpyDeviceToHost);
      printf(" \"cusolver_dpotrf_%d_ms\": %.3f, \"cusolver_dpotrf_%d_tflops\": %.3f, \"cusolver_info_%d\": %d,\n", n, best, n,
             (double)n * n * n / 3.0 / best * 1e-9, n, hi);
      cudaFree(A); cudaFree(w); cudaFree(info);
    }
  }
  printf(" \"done\": 1}\n");
  return 0;
}
