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

// 2^(j/64), j = 0..63, correctly rounded (generated with 60-digit decimal arithmetic).
__device__ const double EXP2_64_TAB[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

// Every kernel that evaluates covariance terms keeps a copy of the table in shared memory (lanes index it with
// different j): call before the kernel's first __syncthreads().  ncu counts 41 % of the fill's shared-memory
// wavefronts as bank conflicts of these lookups; sixteen interleaved copies (one per lane of a half-warp, conflict
// free) were measured and are NOT faster (fill_lower at N = 6000: 64.8 us against 62.6): the lookup is not what bounds
// the kernel, the FP64 pipe is (66 % active).
__device__ __forceinline__ void load_exp_table(double* tab_sm) {
    if (threadIdx.x < 64) tab_sm[threadIdx.x] = EXP2_64_TAB[threadIdx.x];
}

// exp(x) for x <= 0 without a branch (the squared-exponential argument is never positive).  libdevice's exp takes a
// slow path below -708 — which is where most covariance entries live — and its branches keep the compiler from
// interleaving the independent evaluations of an unrolled fill loop.  Table-driven: x = (64 e + j) ln2/64 + r with
// |r| <= ln2/128 (Cody-Waite, two-word ln2/64), exp(x) = 2^e * T[j] * (1 + r + ... + r^5/120) (truncation 3.5e-17),
// scaled by 2^e in two steps so that the subnormal range rounds once: 13 FP64 operations against libdevice's ~21.
// Maximum relative error 2.2e-16 against glibc over [-745, 0] (3e7 samples, host copy of this code); exactly 0
// below -745.2 and NaN for NaN, like exp.  (Scaling by an exact addition to the exponent field when the result is normal,
// with the two multiplications behind a warp-uniform branch for the subnormal range, was measured: the branch ends the
// basic block that lets the unrolled evaluations interleave, fill_lower at N = 6000 71 us against 62.)
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ tab_sm) {
    x = (x < -750.0) ? -750.0 : x;
    const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52: the integer 64 e + j lands in the low word of t
    const double t = fma(x, 92.33248261689365676830, SHIFT);            // 64 / ln 2
    const double n = t - SHIFT;
    double r = fma(n, -1.083042469326755963266e-02, x);                  // ln2/64, high word (21 trailing zero bits)
    r = fma(n, -2.981585826985293281284e-12, r);                         // ln2/64, low word
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    const double p = q * r;                                              // exp(r) - 1
    const int ni = __double2loint(t);
    const double T = tab_sm[ni & 63];
    const double v = fma(T, p, T);
    const int e = ni >> 6, e1 = e >> 1, e2 = e - e1;
    const double s1 = __hiloint2double((e1 + 1023) << 20, 0);
    const double s2 = __hiloint2double((e2 + 1023) << 20, 0);
    return (v * s1) * s2;
}

// One covariance term amp2 * exp((p2 * r) * r), r = zj - zi (matrix_functions.pyx:47-49).
__device__ __forceinline__ double se_term(double amp2, double p2, double zi, double zj, const double* __restrict__ tab_sm) {
    double r = __dsub_rn(zj, zi);
    return __dmul_rn(amp2, exp_neg(__dmul_rn(__dmul_rn(p2, r), r), tab_sm));
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

// ---- mbarrier and 1-D bulk copy (TMA engine, SASS UBLKCP) ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
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

// Lab builds only (-DPSOAP_TIMELINE, tools/timeline.py): every CTA leaves (entry, start after the programmatic-launch
// wait, end) stamps of the GPU's global timer, so that the overlap of the streams of one factorisation can be drawn.
#ifdef PSOAP_TIMELINE
struct TlRec { unsigned long long t_in, t_go, t_out; int kernel, block, nblocks, smid; };
constexpr unsigned TL_CAP = 1u << 20;
__device__ TlRec g_tl[TL_CAP];
__device__ unsigned int g_tl_n;
__device__ __forceinline__ unsigned long long tl_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TL_IN() unsigned tl_i = 0xffffffffu; unsigned long long tl_tin = 0; if (threadIdx.x == 0) tl_tin = psoap::tl_now()
#define TL_GO(kid) do { if (threadIdx.x == 0) { tl_i = atomicAdd(&psoap::g_tl_n, 1u); if (tl_i < psoap::TL_CAP) { \
    unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); psoap::TlRec& r_ = psoap::g_tl[tl_i]; r_.t_in = tl_tin; \
    r_.t_go = psoap::tl_now(); r_.t_out = 0; r_.kernel = (kid); r_.block = blockIdx.x; r_.nblocks = gridDim.x; r_.smid = (int)sm_; } } } while (0)
#define TL_OUT() do { if (threadIdx.x == 0 && tl_i < psoap::TL_CAP) psoap::g_tl[tl_i].t_out = psoap::tl_now(); } while (0)
#else
#define TL_IN() do { } while (0)
#define TL_GO(kid) do { } while (0)
#define TL_OUT() do { } while (0)
#endif

}  // namespace psoap
