// chol.cuh — blocked FP64 Cholesky with the forward solve, log-determinant and quadratic form fused in
// (replaces scipy.linalg.cho_factor / cho_solve = LAPACK dpotrf / dpotrs, psoap/covariance.py:325-331).
//
// Right-looking, panel width NB = 128, on a column-major lower-triangular workspace W [Np, ld]:
//   for kb = 0 .. T-1
//     potrf_diag : one CTA factors the 128x128 diagonal block in registers, builds L_kk^-1 alongside
//                  (Gauss-Jordan), y_k = L_kk^-1 r_k, logdet += sum log d_j, quad += |y_k|^2
//     trsm       : P[i, :] = W[i, kb-panel] * L_kk^-T for the rows below, as a DMMA GEMM with L_kk^-1
//     syrk       : W[I, J] -= P_I P_J^T on the trailing lower triangle (DMMA), plus r_I -= P_I y_k
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
// potrf_diag: 256 threads, thread (ti = tid%16, tc = tid/16) owns the block-cyclic entries
// (i = ti + 16p, c = tc + 16q), p >= q, in registers.  At step j an entry with c > j still holds the
// partially updated matrix, an entry with c < j (row i > j) holds the running right-hand side of
// L X = I, so one rank-1 update per step advances the factorisation and the inverse together.
// One __syncthreads per step; the published column/row are double buffered.
// ------------------------------------------------------------------------------------------------------
constexpr int XS = NB + 1;
constexpr int POTRF_SMEM = (NB * XS + 2 * NB + 2 * NB + NB + NB + 256) * 8;

__global__ void __launch_bounds__(256, 1)
potrf_diag_kernel(const double* __restrict__ W, int64_t ld, int kb, int pad, double* __restrict__ Linv,
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
    const int tid = threadIdx.x;
    const int ti = tid & 15, tc = tid >> 4;
    const double* A = W + (int64_t)kb * NB + (int64_t)kb * NB * ld;

    double M[8][8];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            if (p < q) continue;
            const int i = ti + 16 * p, c = tc + 16 * q;
            M[p][q] = (i >= c) ? A[i + (int64_t)c * ld] : 0.0;
        }
    if (tid < NB) rs[tid] = rvec[kb * NB + tid];

#pragma unroll
    for (int jq = 0; jq < 8; ++jq) {
        for (int jr = 0; jr < 16; ++jr) {
            const int j = 16 * jq + jr;
            double* cb = colb + (j & 1) * NB;
            double* rb = rowb + (j & 1) * NB;
            if (tc == jr) {
#pragma unroll
                for (int p = jq; p < 8; ++p) cb[ti + 16 * p] = M[p][jq];
            }
            if (ti == jr) {
#pragma unroll
                for (int q = 0; q <= jq; ++q) rb[tc + 16 * q] = M[jq][q];
            }
            __syncthreads();
            const double d = cb[j];
            const double inv = rsqrt(d);
            if (tid == 0) dval[j] = d;
            if (tid < NB) {  // row j of X = L^-1 is final
                const int c = tid;
                Xs[c * XS + j] = (c < j) ? rb[c] * inv : ((c == j) ? inv : 0.0);
            }
            double li[8];
#pragma unroll
            for (int p = jq; p < 8; ++p) {
                const int i = ti + 16 * p;
                li[p] = (i > j) ? cb[i] * inv : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int c = tc + 16 * q;
                double w;
                bool special = false;
                if (q > jq) {
                    w = cb[c] * inv;
                } else if (q < jq) {
                    w = rb[c] * inv;
                } else {
                    w = (c > j) ? cb[c] * inv : ((c < j) ? rb[c] * inv : inv);
                    special = (c == j);
                }
#pragma unroll
                for (int p = (q > jq ? q : jq); p < 8; ++p) {
                    const double base = special ? 0.0 : M[p][q];
                    M[p][q] = fma(-li[p], w, base);
                }
            }
        }
    }
    __syncthreads();
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
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    const double lgsum = red[0];
    __syncthreads();
    // y_k = X r_k, quad contribution
    double y = 0.0;
    if (tid < NB) {
        for (int c = 0; c <= tid; ++c) y = fma(Xs[c * XS + tid], rs[c], y);
        yk[tid] = y;
    }
    red[tid] = y * y;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    const double qsum = red[0];
    // first failing pivot in this block
    __syncthreads();
    int* redi = reinterpret_cast<int*>(red);
    redi[tid] = bad;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) redi[tid] = min(redi[tid], redi[tid + s]);
        __syncthreads();
    }
    // L_kk^-1 for the TRSM (column-major 128x128, zeros above the diagonal)
    for (int e = tid; e < NB * NB; e += 256) {
        const int r = e & (NB - 1), c = e >> 7;
        Linv[e] = Xs[c * XS + r];
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
// DMMA tile GEMM  acc[i][j] = sum_{k in [kbeg,kend)} Ai[i,k] * Bj[j,k]   (128 x 64 tile, 256 threads)
//   Ai, Bj column-major (row index contiguous).  cp.async 16 B, 4 stages of BK = 16, padded shared rows
//   (stride = rows + 4 doubles) so the m8n8k4 fragment loads are bank-conflict free.
//   mma M <-> j, mma N <-> i, so each thread ends up with two consecutive rows i of a column j
//   (double2 epilogue on the column-major output).  2 CTAs per SM: one CTA's epilogue (HBM read-modify-
//   write of its tile) overlaps the other's main loop.
// ------------------------------------------------------------------------------------------------------
constexpr int BI = 128, BJ = 64, BK = 16, STAGES = 4;
constexpr int SA = BI + 4, SB = BJ + 4;
constexpr int STAGE_DOUBLES = BK * SA + BK * SB;
constexpr int GEMM_SMEM = STAGES * STAGE_DOUBLES * 8;

struct GemmTile {
    const double* Ai;
    int64_t lda;
    const double* Bj;
    int64_t ldb;
    int kbeg, kend;
    double* C;
    int64_t ldc;
};

template <int MODE>  // 0: C = acc, 1: C -= acc
__device__ __forceinline__ void gemm_tile(const GemmTile& t, double* sm) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int wi = warp & 3, wj = warp >> 2;
    const int KT = (t.kend - t.kbeg) / BK;

    auto load_stage = [&](int stage, int kt) {
        double* sA = sm + stage * STAGE_DOUBLES;
        double* sB = sA + BK * SA;
        const int k0 = t.kbeg + kt * BK;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int q = tid + 256 * r;
            const int col = q >> 6, row2 = q & 63;
            cp_async16(sA + col * SA + 2 * row2, t.Ai + 2 * row2 + (int64_t)(k0 + col) * t.lda);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int q = tid + 256 * r;
            const int col = q >> 5, row2 = q & 31;
            cp_async16(sB + col * SB + 2 * row2, t.Bj + 2 * row2 + (int64_t)(k0 + col) * t.ldb);
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = kt + STAGES - 1;
        if (nk < KT) load_stage(nk % STAGES, nk);
        cp_async_commit();
        const double* sA = sm + (kt % STAGES) * STAGE_DOUBLES;
        const double* sB = sA + BK * SA;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int mj = 0; mj < 4; ++mj) a[mj] = sB[(kk * 4 + tq) * SB + wj * 32 + mj * 8 + g];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = sA[(kk * 4 + tq) * SA + wi * 32 + ni * 8 + g];
#pragma unroll
            for (int mj = 0; mj < 4; ++mj)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma_8x8x4(acc[mj][ni][0], acc[mj][ni][1], a[mj], b[ni]);
        }
    }
    cp_async_wait<0>();

    if (MODE == 0) {
#pragma unroll
        for (int mj = 0; mj < 4; ++mj)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int i = wi * 32 + ni * 8 + tq * 2, j = wj * 32 + mj * 8 + g;
                *reinterpret_cast<double2*>(t.C + i + (int64_t)j * t.ldc) = make_double2(acc[mj][ni][0], acc[mj][ni][1]);
            }
    } else {
        double2 cv[4][4];
#pragma unroll
        for (int mj = 0; mj < 4; ++mj)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int i = wi * 32 + ni * 8 + tq * 2, j = wj * 32 + mj * 8 + g;
                cv[mj][ni] = *reinterpret_cast<const double2*>(t.C + i + (int64_t)j * t.ldc);
            }
#pragma unroll
        for (int mj = 0; mj < 4; ++mj)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int i = wi * 32 + ni * 8 + tq * 2, j = wj * 32 + mj * 8 + g;
                cv[mj][ni].x -= acc[mj][ni][0];
                cv[mj][ni].y -= acc[mj][ni][1];
                *reinterpret_cast<double2*>(t.C + i + (int64_t)j * t.ldc) = cv[mj][ni];
            }
    }
}

// trsm: rows below panel kb.  Block t -> row tile I = kb+1 + t/2, column half jh = t%2.
//   P[I*128 + :, jh*64 + :] = W[I*128 + :, kb*128 + (0..(jh+1)*64)] * Linv[jh*64 + :, same]^T
// nrows_tiles counts 128-row tiles below the panel (border rows of a Schur problem included).
__global__ void __launch_bounds__(256, 2)
trsm_kernel(const double* __restrict__ W, int64_t ld, int kb, int kbeg, const double* __restrict__ Linv,
            double* __restrict__ P, int64_t ldp) {
    extern __shared__ double sm[];
    const int t = blockIdx.x;
    const int I = kb + 1 + (t >> 1), jh = t & 1;
    GemmTile gt;
    gt.Ai = W + (int64_t)I * NB + (int64_t)kb * NB * ld;
    gt.lda = ld;
    gt.Bj = Linv + jh * BJ;
    gt.ldb = NB;
    gt.kbeg = kbeg;
    gt.kend = (jh + 1) * BJ;
    gt.C = P + (int64_t)I * NB + (int64_t)jh * BJ * ldp;
    gt.ldc = ldp;
    gemm_tile<0>(gt, sm);
}

// syrk: trailing update with panel kb, rows/cols of tiles I in [kb+1, T).  Blocks [0, ntiles) are 128x64
// tiles of the lower triangle (row r = I-kb-1 has 2(r+1) tiles); blocks [ntiles, ntiles + nrow) update the
// residual r_I -= P_I y_k (deterministic two-half reduction).  jlimit: last tile column that needs
// updating (T for the likelihood; Schur problems pass the full border).
__global__ void __launch_bounds__(256, 2)
syrk_kernel(double* __restrict__ W, int64_t ld, int kb, int kbeg, const double* __restrict__ P, int64_t ldp,
            int ntiles, const double* __restrict__ yk, double* __restrict__ rvec) {
    extern __shared__ double sm[];
    const int t = blockIdx.x;
    if (t < ntiles) {
        int r = (int)((sqrt(4.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((r + 1) * (r + 2) <= t) ++r;
        while (r * (r + 1) > t) --r;
        const int I = kb + 1 + r;
        const int J64 = 2 * (kb + 1) + (t - r * (r + 1));
        GemmTile gt;
        gt.Ai = P + (int64_t)I * NB;
        gt.lda = ldp;
        gt.Bj = P + (int64_t)J64 * BJ;
        gt.ldb = ldp;
        gt.kbeg = kbeg;
        gt.kend = NB;
        gt.C = W + (int64_t)I * NB + (int64_t)J64 * BJ * ld;
        gt.ldc = ld;
        gemm_tile<1>(gt, sm);
    } else {
        const int I = kb + 1 + (t - ntiles);
        const int tid = threadIdx.x;
        const int row = tid & (NB - 1), half = tid >> 7;
        const double* p = P + (int64_t)I * NB + row + (int64_t)half * 64 * ldp;
        double s = 0.0;
#pragma unroll 8
        for (int c = 0; c < 64; ++c) s = fma(p[(int64_t)c * ldp], yk[half * 64 + c], s);
        sm[tid] = s;
        __syncthreads();
        if (half == 0) rvec[I * NB + row] -= (sm[row] + sm[NB + row]);
    }
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
