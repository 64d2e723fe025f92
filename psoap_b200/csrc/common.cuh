// common.cuh — shared device helpers for the sm_100a PSOAP kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace psoap {

constexpr int NB = 128;                 // panel width / diagonal block / tile rows
constexpr double C_KMS = 2.99792458e5;  // psoap/constants.py:13, matrix_functions.pyx:16
constexpr double C_KMS2 = C_KMS * C_KMS;

// GP hyper-parameters: immediate values (operator surface) or a device vector [amp_f,l_f,amp_g,l_g,amp_h,l_h]
// (chunk farm: the CUDA graph stays valid while the sampler changes the values).
struct GpParams {
    double amp[3];
    double l[3];
    const double* dev;
};

// amp^2 and p2 = -0.5 c^2 / l^2 exactly as matrix_functions.pyx:28-29,:108-112 (no FMA contraction).
__device__ __forceinline__ void gp_coeffs(const GpParams& gp, int c, double& amp2, double& p2) {
    double a = gp.dev ? gp.dev[2 * c] : gp.amp[c];
    double l = gp.dev ? gp.dev[2 * c + 1] : gp.l[c];
    amp2 = __dmul_rn(a, a);
    p2 = __ddiv_rn(-0.5 * C_KMS2, __dmul_rn(l, l));
}

// exp(x) for x <= 0 without a branch (the squared-exponential argument is never positive).  libdevice's exp takes a
// slow path below -708 — which is where most covariance entries live — and its branches keep the compiler from
// interleaving the independent evaluations of an unrolled fill loop.  Cody-Waite reduction x = n ln2 + r,
// |r| <= ln2/2, degree-12 Taylor polynomial (truncation 1.7e-16), result scaled by 2^n in two steps so that the
// subnormal range rounds once.  Maximum relative error 3.2e-16 against glibc over [-745, 0] (2e7 samples, host copy
// of this code); exactly 0 below -745.2 and NaN for NaN, like exp.
__device__ __forceinline__ double exp_neg(double x) {
    x = (x < -750.0) ? -750.0 : x;
    const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52: the integer n lands in the low word of t
    const double t = fma(x, 1.4426950408889634, SHIFT);
    const double n = t - SHIFT;
    double r = fma(n, -6.93147180369123816490e-01, x);
    r = fma(n, -1.90821492927058770002e-10, r);
    double p = 1.0 / 479001600.0;
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int ni = __double2loint(t);
    const int n1 = ni >> 1, n2 = ni - n1;
    const double s1 = __hiloint2double((n1 + 1023) << 20, 0);
    const double s2 = __hiloint2double((n2 + 1023) << 20, 0);
    return (p * s1) * s2;
}

// One covariance term amp2 * exp((p2 * r) * r), r = zj - zi (matrix_functions.pyx:47-49).
__device__ __forceinline__ double se_term(double amp2, double p2, double zi, double zj) {
    double r = __dsub_rn(zj, zi);
    return __dmul_rn(amp2, exp_neg(__dmul_rn(__dmul_rn(p2, r), r)));
}

// Where the ln-wavelength of component c at data index i comes from: either per-component vectors
// (operator surface, already Doppler shifted by the caller) or the base vector shifted on the fly by that
// pixel's epoch velocity: lwl + (-v)/c_kms  (data.py:37,:61).
struct ZSource {
    const double* lwl[3];
    const int32_t* epoch;
    const double* vel;  // [ncomp, n_epochs]
    int n_epochs;
    int shift;
};

__device__ __forceinline__ double z_at(const ZSource& zs, int c, int64_t i) {
    if (!zs.shift) return zs.lwl[c][i];
    double v = zs.vel[(int64_t)c * zs.n_epochs + zs.epoch[i]];
    return __dadd_rn(zs.lwl[0][i], __ddiv_rn(-v, C_KMS));
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// become resident while its predecessor in the stream is still running; it must not touch the predecessor's
// output before pdl_wait().  pdl_trigger() in the predecessor lets the dependent start early.  Both are no-ops
// for kernels launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace psoap
