// Phase timestamps of potrf_diag3 on one 128x128 SPD block (lab build: -DPSOAP_POTRF_TRACE).
#define PSOAP_POTRF_TRACE 1
#include <math_constants.h>
#include <cstdio>
#include <vector>
#include "../psoap_b200/csrc/chol.cuh"
using namespace psoap;
int main() {
  const int n = 128; std::vector<double> h(n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) h[i + j * n] = (i == j ? 2.0 : 0.0) + 1.0 / (1.0 + (i - j) * (i - j));
  double *W, *Linv, *r, *y, *acc, *res; int* info;
  cudaMalloc(&W, n * n * 8); cudaMalloc(&Linv, n * n * 8); cudaMalloc(&r, n * 8); cudaMalloc(&y, n * 8); cudaMalloc(&acc, 64); cudaMalloc(&res, 32); cudaMalloc(&info, 8);
  cudaMemcpy(W, h.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemset(r, 0, n * 8); cudaMemset(acc, 0, 64); cudaMemset(info, 0, 8);
  cudaFuncSetAttribute(potrf_diag3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) potrf_diag3_kernel<<<1, 256, POTRF_SMEM>>>(W, n, 0, 0, Linv, r, y, acc, info, nullptr, 1, res);
  cudaEventRecord(e0);
  for (int w = 0; w < 20; ++w) potrf_diag3_kernel<<<1, 256, POTRF_SMEM>>>(W, n, 0, 0, Linv, r, y, acc, info, nullptr, 1, res);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); printf("potrf_diag3 avg %.2f us (%s)\n", ms * 1000 / 20, cudaGetErrorString(cudaGetLastError()));
  static long long t[8][128][8];
  cudaMemcpyFromSymbol(t, g_potrf_trace, sizeof(t));
  for (int j : {5, 40, 64, 100, 120}) {
    printf("step %3d:", j);
    for (int w : {0, 3, 7}) printf("  warp%d: pub %4lld | bar %4lld | xrow %4lld li %4lld w %4lld | fma %4lld | total %4lld", w, t[w][j][1] - t[w][j][0], t[w][j][2] - t[w][j][1], t[w][j][5] - t[w][j][2], t[w][j][6] - t[w][j][5], t[w][j][3] - t[w][j][6], t[w][j][4] - t[w][j][3], t[w][j + 1][0] - t[w][j][0]);
    printf("\n");
  }
  printf("whole loop: %lld cycles for 127 steps\n", t[0][127][0] - t[0][0][0]);
  return 0;
}
