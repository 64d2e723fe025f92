// Lab: cycles of one 16x16 in-warp LDL^T/Cholesky (phase A of potrf_diag5) in several formulations, one warp.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ void load_row(const double* D, int r, double (&m)[16]) {
#pragma unroll
  for (int c = 0; c < 16; ++c) m[c] = (c <= r) ? D[r * 17 + c] : 0.0;
}
// V0: shuffles, rotating registers (current)
__device__ __noinline__ void va0(const double* __restrict__ D, double* __restrict__ Ld, double* __restrict__ inv16) {
  const int lane = threadIdx.x & 31, r = lane & 15; double m[16]; load_row(D, r, m);
#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    const double d = __shfl_sync(0xffffffffu, m[0], j, 16);
    const double is = rsqrt(d);
    double uc[16];
#pragma unroll
    for (int t = 1; t < 16; ++t) uc[t] = __shfl_sync(0xffffffffu, m[0], j + t, 16);
    const double l = m[0] * is, my = l * is;
    if (lane < 16 && r >= j) Ld[r * 17 + j] = l;
    if (lane == j) inv16[j] = is;
#pragma unroll
    for (int t = 1; t < 16; ++t) m[t - 1] = fma(-my, uc[t], m[t]);
    m[15] = 0.0;
  }
}
// V1: column broadcast through shared memory (STS + syncwarp + 8 LDS.128), rotating registers
__device__ __noinline__ void va1(const double* __restrict__ D, double* __restrict__ Ld, double* __restrict__ inv16, double* col) {
  const int lane = threadIdx.x & 31, r = lane & 15; double m[16]; load_row(D, r, m);
#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    double* cb = col + (j & 1) * 32;          // [32] doubles: entries j..j+15 live at cb[0..15] after rotation
    if (lane < 16) cb[(r - j) & 15] = m[0];   // lane r's column-j entry at slot r-j (lanes r<j write junk slots)
    __syncwarp();
    const double2* c2 = reinterpret_cast<const double2*>(cb);
    double u[16];
#pragma unroll
    for (int t = 0; t < 8; ++t) { double2 v = c2[t]; u[2 * t] = v.x; u[2 * t + 1] = v.y; }
    const double d = u[0];
    const double is = rsqrt(d);
    const double l = m[0] * is, my = l * is;
    if (lane < 16 && r >= j) Ld[r * 17 + j] = l;
    if (lane == j) inv16[j] = is;
#pragma unroll
    for (int t = 1; t < 16; ++t) m[t - 1] = fma(-my, u[t], m[t]);
    m[15] = 0.0;
  }
}
// V2: as V0 but the pivot lane alone takes the rsqrt, float seed + Newton (no MUFU.RSQ64H slow-path branch)
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y = (double)rsqrtf((float)d);
  double e = fma(-d * y, y, 1.0);
  y = fma(y * e, fma(e, 0.375, 0.5), y);
  e = fma(-d * y, y, 1.0);
  y = fma(y * e, fma(e, 0.375, 0.5), y);
  return y;
}
__device__ __noinline__ void va2(const double* __restrict__ D, double* __restrict__ Ld, double* __restrict__ inv16) {
  const int lane = threadIdx.x & 31, r = lane & 15; double m[16]; load_row(D, r, m);
#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    const double d = __shfl_sync(0xffffffffu, m[0], j, 16);
    const double is = fast_rsqrt(d);
    double uc[16];
#pragma unroll
    for (int t = 1; t < 16; ++t) uc[t] = __shfl_sync(0xffffffffu, m[0], j + t, 16);
    const double l = m[0] * is, my = l * is;
    if (lane < 16 && r >= j) Ld[r * 17 + j] = l;
    if (lane == j) inv16[j] = is;
#pragma unroll
    for (int t = 1; t < 16; ++t) m[t - 1] = fma(-my, uc[t], m[t]);
    m[15] = 0.0;
  }
}
// V3: V1 with fast_rsqrt
__device__ __noinline__ void va3(const double* __restrict__ D, double* __restrict__ Ld, double* __restrict__ inv16, double* col) {
  const int lane = threadIdx.x & 31, r = lane & 15; double m[16]; load_row(D, r, m);
#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    double* cb = col + (j & 1) * 32;
    if (lane < 16) cb[(r - j) & 15] = m[0];
    __syncwarp();
    const double2* c2 = reinterpret_cast<const double2*>(cb);
    double u[16];
#pragma unroll
    for (int t = 0; t < 8; ++t) { double2 v = c2[t]; u[2 * t] = v.x; u[2 * t + 1] = v.y; }
    const double is = fast_rsqrt(u[0]);
    const double l = m[0] * is, my = l * is;
    if (lane < 16 && r >= j) Ld[r * 17 + j] = l;
    if (lane == j) inv16[j] = is;
#pragma unroll
    for (int t = 1; t < 16; ++t) m[t - 1] = fma(-my, u[t], m[t]);
    m[15] = 0.0;
  }
}

__device__ long long g_st[2][16][8];
#define ST(v, k) do { if (threadIdx.x == 0) g_st[v][j][k] = clock64(); } while (0)
__device__ __noinline__ void va0t(const double* __restrict__ D, double* __restrict__ Ld, double* __restrict__ inv16) {
  const int lane = threadIdx.x & 31, r = lane & 15; double m[16]; load_row(D, r, m);
#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    ST(0, 0);
    const double d = __shfl_sync(0xffffffffu, m[0], j, 16);
    if (d == 12345.0) inv16[0] = d;
    ST(0, 1);
    const double is = rsqrt(d);
    if (is == 12345.0) inv16[0] = is;
    ST(0, 2);
    double uc[16];
#pragma unroll
    for (int t = 1; t < 16; ++t) uc[t] = __shfl_sync(0xffffffffu, m[0], j + t, 16);
    if (uc[15] == 12345.0) inv16[0] = uc[15];
    ST(0, 3);
    const double l = m[0] * is, my = l * is;
    if (my == 12345.0) inv16[0] = my;
    ST(0, 4);
    if (lane < 16 && r >= j) Ld[r * 17 + j] = l;
    if (lane == j) inv16[j] = is;
    ST(0, 5);
#pragma unroll
    for (int t = 1; t < 16; ++t) m[t - 1] = fma(-my, uc[t], m[t]);
    m[15] = 0.0;
    if (m[14] == 12345.0) inv16[0] = m[14];
    ST(0, 6);
  }
}
__global__ void lab(const double* Dg, double* out, long long* cyc) {
  __shared__ double D[16 * 17], Ld[4][16 * 17], inv16[16];
  __shared__ __align__(16) double col[64];
  for (int i = threadIdx.x; i < 16 * 17; i += 32) D[i] = Dg[i];
  __syncwarp();
  long long t[5];
  va0(D, Ld[0], inv16); __syncwarp();
  t[0] = clock64(); va0(D, Ld[0], inv16); __syncwarp();
  t[1] = clock64(); va1(D, Ld[1], inv16, col); __syncwarp();
  t[2] = clock64(); va2(D, Ld[2], inv16); __syncwarp();
  t[3] = clock64(); va3(D, Ld[3], inv16, col); __syncwarp();
  t[4] = clock64();
  // second pass (instruction cache warm)
  long long s[5];
  s[0] = clock64(); va0(D, Ld[0], inv16); __syncwarp();
  s[1] = clock64(); va1(D, Ld[1], inv16, col); __syncwarp();
  s[2] = clock64(); va2(D, Ld[2], inv16); __syncwarp();
  s[3] = clock64(); va3(D, Ld[3], inv16, col); __syncwarp();
  s[4] = clock64();
  va0t(D, Ld[0], inv16); __syncwarp(); va0t(D, Ld[0], inv16); __syncwarp();
  if (threadIdx.x == 0) for (int v = 0; v < 4; ++v) { cyc[v] = t[v + 1] - t[v]; cyc[4 + v] = s[v + 1] - s[v]; }
  for (int v = 0; v < 4; ++v) for (int i = threadIdx.x; i < 16 * 17; i += 32) out[v * 16 * 17 + i] = Ld[v][i];
}
int main() {
  double h[16 * 17] = {0}, ref[16 * 17] = {0};
  for (int i = 0; i < 16; ++i) for (int j = 0; j <= i; ++j) h[i * 17 + j] = (i == j ? 1e-3 : 0.0) + 0.01 * exp(-0.5 * (i - j) * (i - j) / 9.0);
  for (int j = 0; j < 16; ++j) {
    double d = h[j * 17 + j]; for (int k = 0; k < j; ++k) d -= ref[j * 17 + k] * ref[j * 17 + k];
    ref[j * 17 + j] = sqrt(d);
    for (int i = j + 1; i < 16; ++i) { double s = h[i * 17 + j]; for (int k = 0; k < j; ++k) s -= ref[i * 17 + k] * ref[j * 17 + k]; ref[i * 17 + j] = s / ref[j * 17 + j]; }
  }
  double *D, *out; long long* cyc; cudaMalloc(&D, sizeof(h)); cudaMalloc(&out, 4 * sizeof(h)); cudaMalloc(&cyc, 64);
  cudaMemcpy(D, h, sizeof(h), cudaMemcpyHostToDevice);
  lab<<<1, 32>>>(D, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  double ho[4][16 * 17]; long long hc[8]; cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 64, cudaMemcpyDeviceToHost);
  const char* nm[] = {"V0 shfl + rsqrt", "V1 smem bcast + rsqrt", "V2 shfl + fast_rsqrt", "V3 smem bcast + fast_rsqrt"};
  for (int v = 0; v < 4; ++v) {
    double e = 0; for (int i = 0; i < 16; ++i) for (int j = 0; j <= i; ++j) e = fmax(e, fabs(ho[v][i * 17 + j] - ref[i * 17 + j]) / fabs(ref[i * 17 + j]));
    printf("%-28s first %6lld cycles, warm %6lld cycles (%.0f / step), max rel err %.2e\n", nm[v], hc[v], hc[4 + v], hc[4 + v] / 16.0, e);
  }
  static long long st[2][16][8]; cudaMemcpyFromSymbol(st, g_st, sizeof(st));
  for (int j : {2, 8, 13}) printf("V0 step %2d: shfl d %lld | rsqrt %lld | 15 shfl %lld | l,my %lld | stores %lld | fma %lld | loop %lld\n", j, st[0][j][1]-st[0][j][0], st[0][j][2]-st[0][j][1], st[0][j][3]-st[0][j][2], st[0][j][4]-st[0][j][3], st[0][j][5]-st[0][j][4], st[0][j][6]-st[0][j][5], st[0][j+1][0]-st[0][j][6]);
  return 0;
}
