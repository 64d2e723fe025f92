// potrf_diag7 + trsm7 (csrc/chain.cuh) against a long-double host factorisation, their times beside potrf_diag3,
// and the phase stamps of potrf_diag7 (lab build: -DPSOAP_P7_TRACE).
#define PSOAP_P7_TRACE 1
#include <math_constants.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cudaTypedefs.h>
#include <vector>
#include "../psoap_b200/csrc/chain.cuh"
using namespace psoap;

static PFN_cuTensorMapEncodeTiled g_enc = nullptr;
static CUtensorMap tmap(const double* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  if (!g_enc) { void* fn = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q); g_enc = (PFN_cuTensorMapEncodeTiled)fn; }
  CUtensorMap m; const cuuint64_t gd[2] = {rows, cols}; const cuuint64_t gs[1] = {ld * 8}; const cuuint32_t bx[2] = {box_rows, box_cols}; const cuuint32_t es[2] = {1, 1};
  CUresult r = g_enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed %d\n", (int)r);
  return m;
}

static void host_chol(const std::vector<double>& A, int lda, int n, std::vector<long double>& L) {
  L.assign(n * n, 0.0L);
  for (int j = 0; j < n; ++j) {
    long double d = A[j + (size_t)j * lda];
    for (int k = 0; k < j; ++k) d -= L[j + k * n] * L[j + k * n];
    L[j + j * n] = sqrtl(d);
    for (int i = j + 1; i < n; ++i) {
      long double v = A[i + (size_t)j * lda];
      for (int k = 0; k < j; ++k) v -= L[i + k * n] * L[j + k * n];
      L[i + j * n] = v / L[j + j * n];
    }
  }
}

int main() {
  const int n = 128, R = 3, Nt = (R + 1) * n;
  double *W, *Lfac, *Xd, *r, *y, *acc, *res, *P; int* info;
  cudaMalloc(&W, (size_t)Nt * Nt * 8); cudaMalloc(&Lfac, n * n * 8); cudaMalloc(&Xd, 4 * XD_BLOCK * 8); cudaMalloc(&r, Nt * 8);
  cudaMalloc(&y, Nt * 8); cudaMalloc(&acc, 64); cudaMalloc(&res, 32); cudaMalloc(&info, 8); cudaMalloc(&P, (size_t)Nt * n * 8);
  cudaFuncSetAttribute(potrf_diag3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM);
  cudaFuncSetAttribute(potrf_diag7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF7_SMEM);
  cudaFuncSetAttribute(trsm7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM7_SMEM);
  printf("POTRF7_SMEM = %d bytes, TRSM7_SMEM = %d bytes\n", POTRF7_SMEM, TRSM7_SMEM);
  const CUtensorMap mWblk = tmap(W, Nt, Nt, Nt, P7_LD, n), mWt = tmap(W, Nt, Nt, Nt, T7_RS, n);
  const CUtensorMap mL0 = tmap(Lfac, n, n, n, T7_LS0, 32), mL1 = tmap(Lfac, n, n, n, T7_LS1, 32), mL2 = tmap(Lfac, n, n, n, T7_LS2, 32);
  for (int tc = 0; tc < 4; ++tc) {
    std::vector<double> h((size_t)Nt * Nt, NAN), hr(Nt, 0.0);   // NaN wherever the kernels must not read
    for (int i = 0; i < Nt; ++i) {
      hr[i] = 0.3 * sin(0.37 * i) - 0.05;
      for (int j = 0; j <= i && j < n; ++j) {
        double v;
        const double dv = 2.8 * (i - j);
        if (tc == 0) v = (i == j ? 2.0 : 0.0) + 1.0 / (1.0 + (double)(i - j) * (i - j));
        else if (tc == 1 || tc == 3) v = 0.25 * exp(-0.5 * dv * dv / 25.0) + (i == j ? 6.9e-4 : 0.0);   // package defaults
        else {
          if (i < 40 || j < 40) v = (i == j) ? 1.0 : 0.0;   // front padding
          else v = 0.01 * exp(-0.5 * dv * dv / 25.0) + 0.0025 * exp(-0.5 * dv * dv / 49.0) + (i == j ? 6.9e-4 : 0.0);
        }
        if (i >= n && tc != 2) v += 1e-3 * sin(0.11 * i + 0.7 * j);   // rows below: not just a decaying kernel
        h[i + (size_t)j * Nt] = v;
      }
      if (tc == 2 && i < 40) hr[i] = 0.0;
    }
    if (tc == 3) h[70 + (size_t)70 * Nt] = -1.0;   // not positive definite: pivot 71 fails
    // diagonal sub-block upper halves inside the 128 x 128 block hold valid mirrored numbers in the product
    for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) if ((j & ~1) <= i) h[i + (size_t)j * Nt] = h[j + (size_t)i * Nt];
    std::vector<long double> L;
    host_chol(h, Nt, n, L);
    // host: y = L^-1 r, Pref = W_below L^-T, X_bb
    std::vector<long double> yr(n), Pref((size_t)Nt * n, 0.0L);
    long double ld = 0, qd = 0;
    for (int i = 0; i < n; ++i) { long double v = hr[i]; for (int k = 0; k < i; ++k) v -= L[i + k * n] * yr[k]; yr[i] = v / L[i + i * n]; qd += yr[i] * yr[i]; ld += 2 * logl(L[i + i * n]); }
    for (int i = n; i < Nt; ++i) for (int j = 0; j < n; ++j) { long double v = h[i + (size_t)j * Nt]; for (int k = 0; k < j; ++k) v -= Pref[i + (size_t)k * Nt] * L[j + k * n]; Pref[i + (size_t)j * Nt] = v / L[j + j * n]; }
    cudaMemcpy(W, h.data(), (size_t)Nt * Nt * 8, cudaMemcpyHostToDevice); cudaMemcpy(r, hr.data(), Nt * 8, cudaMemcpyHostToDevice);
    cudaMemset(acc, 0, 64); cudaMemset(info, 0, 8); cudaMemset(Lfac, 0xff, n * n * 8); cudaMemset(y, 0xff, Nt * 8); cudaMemset(P, 0xff, (size_t)Nt * n * 8);
    cudaMemset(Xd, 0xff, 4 * XD_BLOCK * 8);
    potrf_diag7_kernel<<<1, P7_THREADS, POTRF7_SMEM>>>(W, Nt, 0, tc == 2 ? 40 : 0, Lfac, Xd, r, y, acc, info, nullptr, 1, res, mWblk);
    Trsm7Args a; a.W = W; a.ld = Nt; a.kb = 0; a.Lfac = Lfac; a.Xd = Xd; a.P = P; a.ldp = Nt; a.ntiles = 4 * R;
    trsm7_kernel<<<tc == 1 ? 5 : 4 * R, T7_THREADS, TRSM7_SMEM>>>(a, mWt, mL0, mL1, mL2);   // case 1: fewer CTAs than tiles (persistent loop)
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<double> gL(n * n), gy(n), gP((size_t)Nt * n), gX(4 * XD_BLOCK); double gacc[8], gres[4]; int ginfo[2];
    cudaMemcpy(gL.data(), Lfac, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(gy.data(), y, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(gP.data(), P, (size_t)Nt * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(gX.data(), Xd, 4 * XD_BLOCK * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(gacc, acc, 64, cudaMemcpyDeviceToHost); cudaMemcpy(gres, res, 32, cudaMemcpyDeviceToHost); cudaMemcpy(ginfo, info, 8, cudaMemcpyDeviceToHost);
    double el = 0, ey = 0, ep = 0, pmax = 0, ex = 0, xmax = 0;
    for (int c = 0; c < n; ++c) for (int i = c; i < n; ++i) el = fmax(el, fabs(gL[i + c * n] - (double)L[i + c * n]));
    for (int i = 0; i < n; ++i) ey = fmax(ey, fabs(gy[i] - (double)yr[i]));
    for (int i = n; i < Nt; ++i) for (int j = 0; j < n; ++j) { ep = fmax(ep, fabs(gP[i + (size_t)j * Nt] - (double)Pref[i + (size_t)j * Nt])); pmax = fmax(pmax, fabs((double)Pref[i + (size_t)j * Nt])); }
    for (int b = 0; b < 4; ++b) {   // X_bb = inverse of the diagonal sub-block of L
      std::vector<long double> X(32 * 32, 0.0L);
      for (int c = 0; c < 32; ++c) for (int i = c; i < 32; ++i) { long double v = (i == c); for (int k = c; k < i; ++k) v -= L[(32 * b + i) + (32 * b + k) * n] * X[k + c * 32]; X[i + c * 32] = v / L[(32 * b + i) + (32 * b + i) * n]; }
      for (int i = 0; i < 32; ++i) for (int c = 0; c < 32; ++c) { ex = fmax(ex, fabs(gX[b * XD_BLOCK + i * P7_XLD + c] - (double)X[i + c * 32])); xmax = fmax(xmax, fabs((double)X[i + c * 32])); }
    }
    printf("case %d (%s): L err %.3e  X_bb err %.3e (max %.3e)  y err %.3e  P err %.3e (max |P| %.3e)  logdet %.15g (ref %.15Lg)  quad %.15g (ref %.15Lg)  info %d  lnlike %.15g\n",
           tc, cudaGetErrorString(e), el, ex, xmax, ey, ep, pmax, gacc[0], ld, gacc[2], qd, ginfo[0], gres[0]);
  }
  // timing, chained launches on one stream (launch gaps included)
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  double* Linv; cudaMalloc(&Linv, n * n * 8);
  for (int ver : {3, 7}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      for (int w = 0; w < 50; ++w) {
        if (ver == 3) potrf_diag3_kernel<<<1, 256, POTRF_SMEM>>>(W, Nt, 0, 0, Linv, r, y, acc, info, nullptr, 1, res);
        else potrf_diag7_kernel<<<1, P7_THREADS, POTRF7_SMEM>>>(W, Nt, 0, 0, Lfac, Xd, r, y, acc, info, nullptr, 1, res, mWblk);
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("potrf_diag%d: %.2f us per launch (50 back-to-back) (%s)\n", ver, ms * 1000 / 50, cudaGetErrorString(cudaGetLastError()));
  }
  long long t[16];
  cudaMemcpyFromSymbol(t, g_p7_trace, sizeof(t));
  printf("  loaded %lld | pipeline end %lld | kernel end %lld cycles\n", t[1] - t[0], t[9] - t[0], t[10] - t[0]);
  static long long tw[4][2][12];
  cudaMemcpyFromSymbol(tw, g_p7_warp, sizeof(tw));
  for (int b = 0; b < 4; ++b) {
    printf("  block %d, cycles since load: chain done %6lld | store issued %6lld | followers: follow done", b, tw[b][0][0] - t[1], tw[b][0][8] - t[1]);
    for (int w : {1, 2, 3, 5, 6, 7, 9, 10, 11}) printf(" %6lld", tw[b][0][w] - t[1]);
    if (b < 3) { printf(" | update done"); for (int w : {1, 2, 3, 5, 6, 7, 9, 10, 11}) printf(" %6lld", tw[b][1][w] - t[1]); }
    printf("\n");
  }
  { static long long fs[4][4]; cudaMemcpyFromSymbol(fs, g_p7_fsync, sizeof(fs));
    for (int b = 0; b < 3; ++b) printf("  block %d follower 0 (since load): follow done %lld, fenced %lld, barrier passed %lld, critical units done %lld, all critical done %lld\n", b, tw[b][0][1] - t[1], fs[b][0] - t[1], fs[b][1] - t[1], fs[b][2] - t[1], fs[b][3] - t[1]); }
  static long long tf[4][8][4], tcn[4][8];
  cudaMemcpyFromSymbol(tf, g_p7_fol, sizeof(tf)); cudaMemcpyFromSymbol(tcn, g_p7_chn, sizeof(tcn));
  for (int b = 0; b < 2; ++b) {
    const long long base = t[1];
    printf("  block %d micro-steps (cycles since load): chain arrive | follower warp 1: wait-begin, wait-end, step-end\n", b);
    for (int m = 0; m < 8; ++m) printf("    m=%d  chain %6lld | %6lld %6lld (+%lld loads) %6lld\n", m, tcn[b][m] - base, tf[b][m][0] - base, tf[b][m][1] - base, tf[b][m][3] - tf[b][m][1], tf[b][m][2] - base);
  }
  // trsm7 alone on a bigger panel: R row blocks below
  for (int Rb : {15, 37, 46}) {
    const int Ntb = (Rb + 1) * n;
    double *Wb, *Pb; cudaMalloc(&Wb, (size_t)Ntb * n * 8); cudaMalloc(&Pb, (size_t)Ntb * n * 8);
    cudaMemset(Wb, 0, (size_t)Ntb * n * 8);
    Trsm7Args a; a.W = Wb; a.ld = Ntb; a.kb = 0; a.Lfac = Lfac; a.Xd = Xd; a.P = Pb; a.ldp = Ntb; a.ntiles = 4 * Rb;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      const CUtensorMap mWtb = tmap(Wb, Ntb, n, Ntb, T7_RS, n);
      for (int w = 0; w < 50; ++w) trsm7_kernel<<<4 * Rb < 148 ? 4 * Rb : 148, T7_THREADS, TRSM7_SMEM>>>(a, mWtb, mL0, mL1, mL2);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("trsm7 R=%d (%d tiles): %.2f us per launch (%s)\n", Rb, 4 * Rb, ms * 1000 / 50, cudaGetErrorString(cudaGetLastError()));
    { long long tt[8]; cudaMemcpyFromSymbol(tt, g_t7_trace, sizeof(tt));
      printf("   CTA 0: operands landed +%lld, solve +%lld, tile stored +%lld cycles\n", tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2]); }
    cudaFree(Wb); cudaFree(Pb);
  }
  return 0;
}
