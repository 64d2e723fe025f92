// chol.cuh — blocked FP64 Cholesky with the forward solve, log-determinant and quadratic form fused in
// (replaces scipy.linalg.cho_factor / cho_solve = LAPACK dpotrf / dpotrs, psoap/covariance.py:325-331).
//
// Right-looking, panel width NB = 128, on a column-major lower-triangular workspace W [Np, ld]:
//   for kb = 0 .. T-1
//     potrf_diag : one CTA factors the 128x128 diagonal block in registers, builds L_kk^-1 alongside
//                  (Gauss-Jordan), y_k = L_kk^-1 r_k, logdet += sum log d_j, quad += |y_k|^2
//     trsm       : P[i, :] = W[i, kb-panel] * L_kk^-T for the rows below, as a DMMA GEMM with L_kk^-1 (gemm.cuh)
//     syrk       : W[I, J] -= P_I P_J^T on the trailing lower triangle (DMMA), plus r_I -= P_I y_k (gemm.cuh)
// The factor itself is never needed by the likelihood, so the panel P lives in a small (L2-resident)
// ping-pong buffer and is not written back.
#pragma once
#include "common.cuh"

namespace psoap {

// accumulators (doubles): [0] logdet, [1] logdet compensation, [2] quad, [3] quad compensation
__device__ __forceinline__ void kahan_add(double* sum, double* comp, double x) {
    double y = __dsub_rn(x, *comp);
    double t = __dadd_rn(*sum, y);
    *comp = __dsub_rn(__dsub_rn(t, *sum), y);
    *sum = t;
}

// ------------------------------------------------------------------------------------------------------
// potrf_diag3: 256 threads, thread (ti = tid%16, tc = tid/16) owns the block-cyclic entries
// (i = ti + 16p, c = tc + 16q), p >= q, in registers.  At step j an entry with c > j still holds the
// partially updated matrix, an entry with c < j (row i > j) holds the running right-hand side of
// L X = I, so one rank-1 update per step advances the factorisation and the inverse together.
// One __syncthreads per step; the published column/row are double buffered.
// ------------------------------------------------------------------------------------------------------
constexpr int XS = NB + 1;
constexpr int POTRF_SMEM = (NB * XS + 2 * NB + 2 * NB + NB + NB + 512 + 4) * 8;

// Common tail of the diagonal-block kernels: logdet and pivot check, y_k = X r_k and its norm, L_kk^-1 to
// global memory, Kahan accumulation across panels, and the final result record on the last panel.
template <int NT = 256>
__device__ __forceinline__ void potrf_epilogue(int tid, int kb, int pad, const double* Xs, const double* dval,
                                               const double* rs, double* red, double* __restrict__ Linv,
                                               double* __restrict__ yk, double* __restrict__ acc,
                                               int* __restrict__ info, const int* __restrict__ sentinel, int is_last,
                                               double* __restrict__ result) {
    // logdet contribution and pivot check
    double lg = 0.0;
    int bad = 0x7fffffff;
    if (tid < NB) {
        const double d = dval[tid];
        lg = log(d);  // = 2 log L_jj (covariance.py:329)
        if (!(d > 0.0)) bad = tid;
    }
    red[tid] = lg;
    __syncthreads();
    for (int s = NT / 2; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    const double lgsum = red[0];
    __syncthreads();
    double y = 0.0;
    if (tid < NB) {
        for (int c = 0; c <= tid; ++c) y = fma(Xs[c * XS + tid], rs[c], y);
        yk[tid] = y;
    }
    red[tid] = y * y;
    __syncthreads();
    for (int s = NT / 2; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    const double qsum = red[0];
    __syncthreads();
    int* redi = reinterpret_cast<int*>(red);
    redi[tid] = bad;
    __syncthreads();
    for (int s = NT / 2; s > 0; s >>= 1) {
        if (tid < s) redi[tid] = min(redi[tid], redi[tid + s]);
        __syncthreads();
    }
    for (int e = tid; e < NB * NB; e += NT) {
        const int r = e & (NB - 1), c = e >> 7;
        Linv[e] = (r >= c) ? Xs[c * XS + r] : 0.0;
    }
    if (tid == 0) {
        if (redi[0] != 0x7fffffff && info[0] == 0) info[0] = kb * NB + redi[0] - pad + 1;
        kahan_add(&acc[0], &acc[1], lgsum);
        kahan_add(&acc[2], &acc[3], qsum);
        if (is_last) {
            const int inf = info[0];
            const bool flagged = (inf != 0) || (sentinel != nullptr && sentinel[0] != 0);
            result[0] = flagged ? -CUDART_INF : -0.5 * (acc[2] + acc[0]);  // covariance.py:331
            result[1] = acc[0];
            result[2] = acc[2];
            result[3] = (double)inf;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// The per-step instruction count is what bounds this kernel (the step is issue bound: 8 warps x ~130 instructions on
// one SM), so the owners of column j scale it BEFORE
// publishing (the pivot travels to them by warp shuffle, they sit in one half-warp), rows that are already
// finished are published as zeros, so a consumer's step is 16 shared loads, <= 36 DFMA and a handful of
// multiplies for the inverse part; no thread but the 16 column owners evaluates the reciprocal square root.
// ------------------------------------------------------------------------------------------------------
#ifdef PSOAP_POTRF_TRACE
__device__ long long g_potrf_trace[8][128][8];   // [warp][step][phase] clock64 stamps (lab builds only)
#define PSOAP_TRACE(ph) do { if ((tid & 31) == 0) g_potrf_trace[tid >> 5][j][ph] = clock64(); } while (0)
#else
#define PSOAP_TRACE(ph) do { } while (0)
#endif

template <int JQ>
__device__ __forceinline__ void potrf3_block_steps(double (&M)[8][8], int ti, int tc, int tid, double* Xs, double* colb,
                                                   double* rowb, double* scal, double* dval) {
#pragma unroll 1
    for (int jr = 0; jr < 16; ++jr) {
        const int j = 16 * JQ + jr;
        double* cb = colb + (j & 1) * NB;
        double* rb = rowb + (j & 1) * NB;
        PSOAP_TRACE(0);
        // ---- publish: scaled column j (zeros for the finished rows), raw row j of the inverse part, pivot
        if (tc == jr) {
            const unsigned hmask = 0xFFFFu << (16 * (tid >> 4 & 1));
            const double d = __shfl_sync(hmask, M[JQ][JQ], (tid & 16) + jr);
            const double inv = rsqrt(d);
#pragma unroll
            for (int p = JQ; p < 8; ++p) {
                const bool below = (p > JQ) || (ti > jr);
                cb[ti + 16 * p] = below ? M[p][JQ] * inv : 0.0;
            }
            if (ti == jr) { scal[(j & 1) * 2] = d; scal[(j & 1) * 2 + 1] = inv; dval[j] = d; }
        }
        if (ti == jr) {
#pragma unroll
            for (int q = 0; q <= JQ; ++q) rb[tc + 16 * q] = M[JQ][q];
        }
        PSOAP_TRACE(1);
        __syncthreads();
        PSOAP_TRACE(2);
        // ---- consume
        const double inv = scal[(j & 1) * 2 + 1];
        if (tid < NB) {  // row j of X = L^-1 is final
            const int c = tid;
            Xs[c * XS + j] = (c < j) ? rb[c] * inv : ((c == j) ? inv : 0.0);
        }
        PSOAP_TRACE(5);
        double li[8], w[8];
#pragma unroll
        for (int p = JQ; p < 8; ++p) li[p] = cb[ti + 16 * p];
        PSOAP_TRACE(6);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = tc + 16 * q;
            if (q > JQ) w[q] = cb[c];
            else if (q < JQ) w[q] = rb[c] * inv;
            else w[q] = (tc > jr) ? cb[c] : ((tc < jr) ? rb[c] * inv : inv);
        }
        const bool special = (tc == jr);
        PSOAP_TRACE(3);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
#pragma unroll
            for (int p = (q > JQ ? q : JQ); p < 8; ++p) {
                const double base = (special && q == JQ) ? 0.0 : M[p][q];
                M[p][q] = fma(-li[p], w[q], base);
            }
        }
        PSOAP_TRACE(4);
    }
}

__global__ void __launch_bounds__(256, 1)
potrf_diag3_kernel(const double* __restrict__ W, int64_t ld, int kb, int pad, double* __restrict__ Linv,
                   double* __restrict__ rvec, double* __restrict__ yk, double* __restrict__ acc,
                   int* __restrict__ info, const int* __restrict__ sentinel, int is_last,
                   double* __restrict__ result) {
    extern __shared__ double sm[];
    double* Xs = sm;                  // Xs[c*XS + r] = X[r][c], X = L^-1
    double* colb = Xs + NB * XS;      // [2][NB]
    double* rowb = colb + 2 * NB;     // [2][NB]
    double* dval = rowb + 2 * NB;     // [NB] pivots d_j
    double* rs = dval + NB;           // [NB] residual segment
    double* red = rs + NB;            // [256]
    double* scal = red + 512;         // [2][2]: pivot d_j and d_j^-1/2
    const int tid = threadIdx.x;
    const int ti = tid & 15, tc = tid >> 4;
    const double* A = W + (int64_t)kb * NB + (int64_t)kb * NB * ld;
    TL_IN();
    pdl_trigger();   // one CTA: the panel solve may become resident on the other SMs while this block is factored
    pdl_wait();
    TL_GO(5);

    double M[8][8];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            if (p < q) { M[p][q] = 0.0; continue; }
            const int i = ti + 16 * p, c = tc + 16 * q;
            M[p][q] = (i >= c) ? A[i + (int64_t)c * ld] : 0.0;
        }
    if (tid < NB) rs[tid] = rvec[kb * NB + tid];

    potrf3_block_steps<0>(M, ti, tc, tid, Xs, colb, rowb, scal, dval);
    potrf3_block_steps<1>(M, ti, tc, tid, Xs, colb, rowb, scal, dval);
    potrf3_block_steps<2>(M, ti, tc, tid, Xs, colb, rowb, scal, dval);
    potrf3_block_steps<3>(M, ti, tc, tid, Xs, colb, rowb, scal, dval);
    potrf3_block_steps<4>(M, ti, tc, tid, Xs, colb, rowb, scal, dval);
    potrf3_block_steps<5>(M, ti, tc, tid, Xs, colb, rowb, scal, dval);
    potrf3_block_steps<6>(M, ti, tc, tid, Xs, colb, rowb, scal, dval);
    potrf3_block_steps<7>(M, ti, tc, tid, Xs, colb, rowb, scal, dval);
    __syncthreads();
    potrf_epilogue(tid, kb, pad, Xs, dval, rs, red, Linv, yk, acc, info, sentinel, is_last, result);
    TL_OUT();
}

// Register-resident DMMA loop: the FP64 tensor-pipe peak used as the roofline denominator.
__global__ void dmma_peak_kernel(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    const double a = 1e-3 + threadIdx.x * 1e-9, b = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma_8x8x4(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace psoap
