// Lab: potrf_diag5 (blocked) against potrf_diag3 and a host long-double reference on one 128x128 SPD block.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 potrf5_lab.cu -o potrf5_lab.bin
#define PSOAP_POTRF_TRACE 1
#include <math_constants.h>
#include <cmath>
#include <cstdio>
#include <vector>
#include "../psoap_b200/csrc/chol.cuh"
using namespace psoap;
typedef void (*kern_t)(const double*, int64_t, int, int, double*, double*, double*, double*, int*, const int*, int, double*);
struct Out { std::vector<double> Linv, y; double acc[8], res[4]; float us; };
static Out run(kern_t k, int smem, const std::vector<double>& h, const std::vector<double>& hr) {
  const int n = 128;
  double *W, *Linv, *r, *y, *acc, *res; int* info;
  cudaMalloc(&W, n * n * 8); cudaMalloc(&Linv, n * n * 8); cudaMalloc(&r, n * 8); cudaMalloc(&y, n * 8); cudaMalloc(&acc, 64); cudaMalloc(&res, 32); cudaMalloc(&info, 8);
  cudaMemcpy(W, h.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(r, hr.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemset(acc, 0, 64); cudaMemset(info, 0, 8); cudaMemset(Linv, 0xff, n * n * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<1, 256, smem>>>(W, n, 0, 0, Linv, r, y, acc, info, nullptr, 1, res);
  Out o; o.Linv.resize(n * n); o.y.resize(n);
  cudaError_t e = cudaDeviceSynchronize();
  if (e) { printf("kernel error: %s\n", cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(o.Linv.data(), Linv, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(o.y.data(), y, n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(o.acc, acc, 64, cudaMemcpyDeviceToHost); cudaMemcpy(o.res, res, 32, cudaMemcpyDeviceToHost);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) k<<<1, 256, smem>>>(W, n, 0, 0, Linv, r, y, acc, info, nullptr, 1, res);
  cudaEventRecord(e0);
  for (int w = 0; w < 50; ++w) k<<<1, 256, smem>>>(W, n, 0, 0, Linv, r, y, acc, info, nullptr, 1, res);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); o.us = ms * 1000 / 50;
  return o;
}
int main() {
  const int n = 128; std::vector<double> h(n * n), hr(n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j)
    h[i + j * n] = (i == j ? 1e-3 : 0.0) + 0.01 * exp(-0.5 * (i - j) * (i - j) / 9.0) + 0.0025 * exp(-0.5 * (i - j) * (i - j) / 25.0);
  for (int i = 0; i < n; ++i) hr[i] = sin(0.37 * i) * 0.05;
  // host reference in long double
  std::vector<long double> L(n * n, 0.0L), X(n * n, 0.0L), yy(n);
  long double logdet = 0;
  for (int j = 0; j < n; ++j) {
    long double d = h[j + j * n];
    for (int k = 0; k < j; ++k) d -= L[j + k * n] * L[j + k * n];
    logdet += logl(d);
    L[j + j * n] = sqrtl(d);
    for (int i = j + 1; i < n; ++i) { long double s = h[i + j * n]; for (int k = 0; k < j; ++k) s -= L[i + k * n] * L[j + k * n]; L[i + j * n] = s / L[j + j * n]; }
  }
  for (int c = 0; c < n; ++c) for (int i = c; i < n; ++i) { long double s = (i == c) ? 1.0L : 0.0L; for (int k = c; k < i; ++k) s -= L[i + k * n] * X[k + c * n]; X[i + c * n] = s / L[i + i * n]; }
  long double quad = 0;
  for (int i = 0; i < n; ++i) { long double s = 0; for (int c = 0; c <= i; ++c) s += X[i + c * n] * hr[c]; yy[i] = s; quad += s * s; }
  Out o3 = run(potrf_diag3_kernel, POTRF_SMEM, h, hr);
  Out o5 = run(potrf_diag5_kernel, POTRF5_SMEM, h, hr);
  for (auto* o : {&o3, &o5}) {
    double ex = 0, ey = 0, xmax = 0;
    for (int i = 0; i < n * n; ++i) { ex = fmax(ex, fabs(o->Linv[i] - (double)X[i])); xmax = fmax(xmax, fabs((double)X[i])); }
    for (int i = 0; i < n; ++i) ey = fmax(ey, fabs(o->y[i] - (double)yy[i]));
    printf("%s: %.2f us  max|X err| %.3e (max|X| %.3e)  max|y err| %.3e  logdet %.15g (ref %.15Lg)  quad %.15g (ref %.15Lg) lnlike %.15g info %g\n",
           o == &o3 ? "potrf_diag3" : "potrf_diag5", o->us, ex, xmax, ey, o->acc[0], logdet, o->acc[2], quad, o->res[0], o->res[3]);
  }
  static long long t[8][8][8];
  cudaMemcpyFromSymbol(t, g_potrf5_trace, sizeof(t));
  for (int b = 0; b < 8; ++b) {
    printf("block %d:", b);
    for (int w : {0, 3, 7}) printf("  w%d: extract %4lld bar %4lld | A %5lld bar %4lld | B %4lld bar %4lld | C %5lld", w, t[w][b][1] - t[w][b][0], t[w][b][2] - t[w][b][1], t[w][b][3] - t[w][b][2], t[w][b][4] - t[w][b][3], t[w][b][5] - t[w][b][4], t[w][b][6] - t[w][b][5], t[w][b][7] - t[w][b][6]);
    printf("\n");
  }
  printf("blocks total: %lld cycles\n", t[0][7][7] - t[0][0][0]);
  return 0;
}
