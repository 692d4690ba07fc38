// common.cuh -- shared device helpers for libsmcb200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/smcb200.h"

namespace smcb {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs
constexpr double kNegInf = -__builtin_huge_val();

// ---- host-side error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> SMCB_ERR_CUDA

#define SMCB_REQUIRE(cond, msg)                       \
    do {                                              \
        if (!(cond)) {                                \
            ::smcb::set_error("%s: %s", __func__, msg); \
            return SMCB_ERR_ARG;                      \
        }                                             \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- runtime options (smcb_set_option; defaults from the SMCB_* environment variables) ----
// Tuning / test aids: they select between kernel variants that compute the same step, so the
// tests can pin every variant the dispatcher can take.  -1 = automatic.
struct Options {
    int no_tma;            // 1: never take the TMA-staged wide identity kernel
    int strict_proposal;   // 1: wide identity kernel computes sqrt(s) L z with fp64 FMAs (no tf32 mma)
    int linear_cfg;        // -1 auto; 0..6 force a (threads, particles/thread, dynamic) configuration
    int persist;           // -1 auto; 0/1 one tile per block / persistent blocks
    int no_dmma;           // 1: second covariance pass on the FP64 vector pipe
    int cov_two_pass;      // 1: never take the one-pass shifted covariance
    int no_coop;           // 1: launch-chain bisection / two-launch reweight
    int coop_tpb, coop_bps;// cooperative kernels: threads per block, blocks per SM
    int bisect_plain;      // 1: evaluate every scipy midpoint (no monotone bracketing)
    int spring_ppt;        // -1 auto; 1/2 particles per thread in the RK4 kernel
    int no_mvn3;           // 1: never take the static-layout MVNormal kernel (K_LINEAR_MVN3)
    int no_pref;           // 1: never take the cp.async-prefetching variant of it
    int xchg_timeout_s;    // peer-exchange time-out inside kernels (seconds; 0 = default 30)
    int no_fused_resample; // 1: separate uniforms / search / count / gather kernels (native mode)
    int spring_form;       // spring-mass kernels: 1 (default) RK4 propagator of the linear ODE, 0 RK4 stages at every sub-step
    int no_pdl;            // 1: smcb_mh_steps does not chain its launches by programmatic dependent launch
};
Options& options();
void note_mh_variant(const char* text);

// grid sized in multiples of the SM count (persistent-style grid-stride kernels)
inline int grid_for(int64_t n, int threads, int items_per_thread, int max_blocks_per_sm) {
    int64_t per_block = (int64_t)threads * items_per_thread;
    int64_t want = (n + per_block - 1) / per_block;
    int64_t cap = (int64_t)kNumSMs * max_blocks_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// workspace layout (bytes from the start of `ws`); the caller zero-initialises once
constexpr size_t kWsTickets = 0;            // 64 x u32 tickets / counters
constexpr size_t kWsPartials = 1024;        // block partials: up to 4096 blocks x 8 doubles
constexpr size_t kWsPartialsBytes = 4u << 20;  // 4 MiB
constexpr size_t kWsSmall = kWsPartials + kWsPartialsBytes;   // 4 KB small scalars
constexpr size_t kWsLarge = kWsSmall + 4096;                  // scan tile state / flags

// ---- warp / block reductions ---------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_ull(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; result valid in thread 0 (deterministic order). `sh` >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (wid == 0) v = warp_sum(v);
    return v;
}

// ---- log-sum-exp triple (max, S1 = sum e^(v-max), S2 = sum e^(2(v-max))) --------------
struct Lse {
    double m, s1, s2;
};
__device__ __forceinline__ Lse lse_empty() { return Lse{kNegInf, 0.0, 0.0}; }
__device__ __forceinline__ Lse lse_merge(const Lse& a, const Lse& b) {
    Lse r;
    r.m = fmax(a.m, b.m);
    if (r.m == kNegInf) {
        r.s1 = 0.0;
        r.s2 = 0.0;
        return r;
    }
    const double ea = (a.m == kNegInf) ? 0.0 : exp(a.m - r.m);
    const double eb = (b.m == kNegInf) ? 0.0 : exp(b.m - r.m);
    r.s1 = a.s1 * ea + b.s1 * eb;
    r.s2 = a.s2 * (ea * ea) + b.s2 * (eb * eb);
    return r;
}
// Block merge, max first: the block maximum is an exact fmax reduction, every thread then
// rescales its own sums with ONE independent exp, and the sums are plain ordered reductions
// (a tree of pairwise lse_merge calls would chain ~20 dependent exps).  Result valid in
// thread 0; sh >= 32 doubles.
__device__ __forceinline__ double block_max_all(double v, double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double m = kNegInf;
    for (int k = 0; k < nw; ++k) m = fmax(m, sh[k]);      // same value in every thread
    return m;
}
__device__ __forceinline__ Lse block_lse(Lse v, double* sh) {
    const double bm = block_max_all(v.m, sh);
    const double e = (v.m == kNegInf) ? 0.0 : exp(v.m - bm);
    Lse r;
    r.m = bm;
    r.s1 = block_sum(v.s1 * e, sh);
    r.s2 = block_sum(v.s2 * (e * e), sh);
    return r;
}
// Merge of n partial triples stored as (m, s1, s2) rows by a whole block (all threads call; result
// valid in thread 0): every thread
// takes rows t, t + blockDim, ... so the L2 round trips overlap instead of chaining ten deep in
// one lane (the one-warp version cost ~9 us for 296 rows).  Rows are read with ld.global.cg:
// they were written by other SMs.  Fixed reduction order: deterministic, identical in every block.
__device__ __forceinline__ Lse block_merge_partials(const double* parts, int n, double* sh) {
    double m = kNegInf;
    for (int b = threadIdx.x; b < n; b += blockDim.x) m = fmax(m, __ldcg(parts + 3 * b));
    const double bm = block_max_all(m, sh);
    double s1 = 0.0, s2 = 0.0;
    if (bm != kNegInf) {
        for (int b = threadIdx.x; b < n; b += blockDim.x) {
            const double mb = __ldcg(parts + 3 * b);
            const double e = (mb == kNegInf) ? 0.0 : exp(mb - bm);
            s1 += __ldcg(parts + 3 * b + 1) * e;
            s2 += __ldcg(parts + 3 * b + 2) * (e * e);
        }
    }
    Lse r;
    r.m = bm;
    r.s1 = block_sum(s1, sh);
    r.s2 = block_sum(s2, sh);
    return r;
}
// exp(x) for x <= 0 (or -inf) as the streaming log-sum-exp loops need it: 2^(k/32) table + degree-6
// polynomial on |r| <= ln2/64 (truncation 3.5e-18), no special-case paths; arguments below -708
// (results under 1e-307 next to a maximum term of 1) return 0.  ~1 ulp; 11 FP64 instructions
// against ~17 + select chains in the CUDA library exp.
static __device__ const double kExpTab[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};
__device__ __forceinline__ double exp_nonpos(double x) {
    const double kMagic = 6755399441055744.0;                      // 1.5 * 2^52: round-to-nearest-integer trick
    double kd = fma(x, 0x1.71547652b82fep+5, kMagic);              // x * 32/ln2
    const int ki = __double2loint(kd);
    kd -= kMagic;
    double r = fma(kd, -0x1.62e42fefa39efp-6, x);                  // ln2/32, hi + lo
    r = fma(kd, -0x1.abc9e3b39803fp-61, r);
    double q = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    q = fma(q, r, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    const double t = __ldg(&kExpTab[ki & 31]);
    const double sres = fma(r * r, q, r);                          // exp(r) - 1
    const double v = fma(t, sres, t);
    const double scaled = __hiloint2double(__double2hiint(v) + ((ki >> 5) << 20), __double2loint(v));
    return (x >= -708.0) ? scaled : 0.0;
}
// add one value with a running max (rescale only when the max moves)
__device__ __forceinline__ void lse_push4(Lse& acc, double v0, double v1, double v2, double v3) {
    const double tm = fmax(fmax(v0, v1), fmax(v2, v3));
    if (tm > acc.m) {
        if (acc.m != kNegInf) {
            const double e = exp_nonpos(acc.m - tm);
            acc.s1 *= e;
            acc.s2 *= e * e;
        }
        acc.m = tm;
    }
    if (acc.m == kNegInf) return;
    const double e0 = exp_nonpos(v0 - acc.m), e1 = exp_nonpos(v1 - acc.m), e2 = exp_nonpos(v2 - acc.m),
                 e3 = exp_nonpos(v3 - acc.m);
    acc.s1 += (e0 + e1) + (e2 + e3);
    acc.s2 += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
}

// ---- GeometricPath arithmetic (smcpy/paths.py:81-105) ---------------------------------
struct PathCoef {
    double phi, pe, qe;        // exponents at phi
    double phi0, pe0, qe0;     // exponents at phi_prev
};
inline PathCoef make_path_coef(double phi, double phi_prev, double lambda) {
    PathCoef c;
    c.phi = phi;
    c.pe = fmin(1.0, phi / lambda);
    c.qe = fmax(0.0, (lambda - phi) / lambda);
    c.phi0 = phi_prev;
    c.pe0 = fmin(1.0, phi_prev / lambda);
    c.qe0 = fmax(0.0, (lambda - phi_prev) / lambda);
    return c;
}
// row sum in the reference's column order, no FMA contraction, NaN -> -inf (paths.py:55-61)
__device__ __forceinline__ double path_target(double ll, double lp, double lq, double phi, double pe,
                                              double qe) {
    double t = __dmul_rn(ll, phi);
    t = __dadd_rn(t, pe > 0.0 ? __dmul_rn(lp, pe) : 0.0);
    t = __dadd_rn(t, qe > 0.0 ? __dmul_rn(lq, qe) : 0.0);
    return (t != t) ? kNegInf : t;
}
__device__ __forceinline__ double path_inc(double ll, double lp, double lq, const PathCoef& c) {
    double t = __dmul_rn(ll, c.phi);
    t = __dadd_rn(t, c.pe > 0.0 ? __dmul_rn(lp, c.pe) : 0.0);
    t = __dadd_rn(t, c.qe > 0.0 ? __dmul_rn(lq, c.qe) : 0.0);
    t = __dadd_rn(t, -__dmul_rn(ll, c.phi0));
    t = __dadd_rn(t, c.pe0 > 0.0 ? -__dmul_rn(lp, c.pe0) : -0.0);
    t = __dadd_rn(t, c.qe0 > 0.0 ? -__dmul_rn(lq, c.qe0) : -0.0);
    return (t != t) ? kNegInf : t;
}

// ---- un-normalised log-weight of one particle (updater.py:122-133) -------------------------------
struct RwIn {
    const double* ll;
    const double* lp;
    const double* lq;      // nullptr -> lp
    const double* lw;      // nullptr -> uniform
    double lw_uniform;     // log(1/n_global)
    double lp_const;       // used when lp == nullptr (flat priors: the summed log-prior is one constant)
    PathCoef c;
};

__device__ __forceinline__ double rw_value(const RwIn& in, int64_t i) {
    const double ll = in.ll[i];
    const double lp = in.lp ? in.lp[i] : in.lp_const;
    const double lq = in.lq ? in.lq[i] : lp;
    const double inc = path_inc(ll, lp, lq, in.c);
    const double lw = in.lw ? in.lw[i] : in.lw_uniform;
    return __dadd_rn(lw, inc);
}

// incremental weight only (adaptive search: previous weights are ignored, samplers.py:264-276)
__device__ __forceinline__ double inc_value(const RwIn& in, int64_t i) {
    const double ll = in.ll[i];
    const double lp = in.lp ? in.lp[i] : in.lp_const;
    const double lq = in.lq ? in.lq[i] : lp;
    return path_inc(ll, lp, lq, in.c);
}

inline RwIn make_rw_in(const double* ll, const double* lp, const double* lq, const double* lw, int64_t n_global,
                       const smcb_path_t* path, double lp_const = 0.0) {
    RwIn in;
    in.ll = ll;
    in.lp = lp;
    in.lq = lq;
    in.lw = lw;
    in.lw_uniform = log(1.0 / (double)n_global);
    in.lp_const = lp_const;
    in.c = make_path_coef(path->phi, path->phi_prev, path->lambda);
    return in;
}

// ---- Philox4x32-10 counter-based generator ---------------------------------------------
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint64_t id, uint32_t call, uint32_t stream) const {
        uint32_t c0 = (uint32_t)id, c1 = (uint32_t)(id >> 32), c2 = call, c3 = stream;
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
            c0 = n0;
            c1 = lo1;
            c2 = n2;
            c3 = lo0;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
// Same generator with the 10 round keys precomputed on the host (they are the same for every
// thread): the key schedule becomes constant-bank operands instead of ~20 adds per call.
struct PhiloxKeys {
    uint32_t k[20];
};
inline PhiloxKeys make_philox_keys(uint64_t seed) {
    PhiloxKeys r;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) {
        r.k[2 * i] = a;
        r.k[2 * i + 1] = b;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    return r;
}
__device__ __forceinline__ uint4 philox_k(const PhiloxKeys& rk, uint64_t id, uint32_t call, uint32_t stream) {
    uint32_t c0 = (uint32_t)id, c1 = (uint32_t)(id >> 32), c2 = call, c3 = stream;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ rk.k[2 * r], n2 = hi0 ^ c3 ^ rk.k[2 * r + 1];
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}
// The same generator with the key schedule computed in registers (two adds per round): for the
// out-of-line RNG blocks of the wide kernels, where a pointer to the precomputed keys would turn
// every key into a generic load from the parameter space (ncu: 20 LD.E per call).
__device__ __forceinline__ uint4 philox_s(uint32_t k0, uint32_t k1, uint64_t id, uint32_t call, uint32_t stream) {
    uint32_t c0 = (uint32_t)id, c1 = (uint32_t)(id >> 32), c2 = call, c3 = stream;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ (k0 + (uint32_t)r * 0x9E3779B9u), n2 = hi0 ^ c3 ^ (k1 + (uint32_t)r * 0xBB67AE85u);
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}
// 53-bit uniform in [0,1) from two 32-bit words (same construction as numpy's next_double)
__device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {
    const uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
// two standard normals from two 32-bit words: Box-Muller in fp32 on the special-function unit
// (MUFU.LG2 / SQRT / SIN / COS through the .approx.ftz PTX forms: the library wrappers __log2f, rsqrtf,
// __frsqrt_rn add denormal scaling and Newton refinement -- 34 SASS instructions per pair, 13 here; the
// arguments are never denormal).  Proposals only need a symmetric, well-distributed z (the Metropolis
// ratio does not involve the proposal density) -- see DESIGN.md "RNG".
__device__ __forceinline__ float mufu_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sin(float x) {
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_cos(float x) {
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void normal_pair_f(uint32_t w0, uint32_t w1, float& z0, float& z1) {
    const float v = (float)(w0 >> 8) + 0.5f;                       // 2^24 u1, u1 in (0, 1]: 24 bits
    const float th = ((float)(int32_t)w1) * (3.14159265358979f / 2147483648.0f);   // [-pi, pi)
    // -2 ln u1 = 2 ln 2 (24 - lg2 v); v can round to 2^24 and the approximation can overshoot: clamp at 0
    const float x = fmaxf(fmaf(mufu_lg2(v), -1.38629436111989f, 24.0f * 1.38629436111989f), 0.0f);
    const float r = mufu_sqrt(x);
    z0 = r * mufu_cos(th);
    z1 = r * mufu_sin(th);
}
__device__ __forceinline__ void normal_pair(uint32_t w0, uint32_t w1, double& z0, double& z1) {
    float a, b;
    normal_pair_f(w0, w1, a, b);
    z0 = (double)a;
    z1 = (double)b;
}

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a): 1-D cp.async.bulk global -> shared ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// bytes and both addresses must be multiples of 16
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// per-thread asynchronous 8-byte copy global -> shared (LDGSTS): each thread prefetches ITS OWN next-tile values
// into thread-private shared slots, so no block barrier is needed -- cp_async_wait_all() makes them visible to
// the issuing thread
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace smcb
