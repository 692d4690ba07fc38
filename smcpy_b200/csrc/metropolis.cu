// metropolis.cu -- kernel (d): the fused VectorMCMC.smc_metropolis step.
//
// One launch = one Metropolis step for every particle (smcpy/mcmc/vector_mcmc.py:168-183):
//   Gaussian random-walk proposal x' = x + sqrt(s) L z          (:124-129, cov scale rule :55-60)
//   prior log-pdfs of x'                                         (:98-111)
//   model + log-likelihood where the prior is finite             (:205-214, log_likelihoods.py)
//   tempered acceptance ratio exp(target(x') - target(x))        (:131-146, paths.py:81-84)
//   reject iff ratio < u, select                                 (:148-151, :176-181)
// in one pass over structure-of-arrays particles.  Persistent blocks (a multiple of the SM
// count) stage the Cholesky factor and the shared data vector in shared memory once, then loop
// over tiles of TPB*PPT particles: thread t owns particles tile*TPB*PPT + p*TPB + t (coalesced
// column loads), keeps proposal, model outputs and residuals in registers, and writes back only
// accepted rows.  Algorithmic HBM bytes per particle-step: 2 * 8 * (d + 2).
//
// The same kernels in other modes evaluate log-priors and log-likelihood of incoming
// particles (vector_mcmc.py:162-166, initializer.py:66-73), priors only, or the proposal only.
#include <stdlib.h>

#include "common.cuh"

namespace smcb {

constexpr int kMhThreads = 128;
constexpr int kChunk = 1024;          // data points staged per shared-memory chunk
constexpr int kMvnStats = SMCB_MVN_STATS;   // MVNormal: snapshot mean (3) + scatter matrix (3 x 3)

// K_LINEAR_MVN3: the MVNormal layout of config C4 -- 3 features, all six covariance terms sampled in the
// trailing columns 2..7 (and, if present, one 3 x 3 ImproperCov prior over the same columns): every column
// index is a compile-time constant, so the parameter registers never spill to an indexable local array
// (profiles/r01_c3_c4_kernels.md: STL.64 in the hot loop of the generic kind).
enum Kind { K_LINEAR_NORMAL = 0, K_SPRING_NORMAL = 1, K_IDENTITY_NORMAL = 2, K_LINEAR_MVN = 3, K_LINEAR_MVN3 = 4 };
__host__ __device__ constexpr bool is_mvn(int kind) { return kind == K_LINEAR_MVN || kind == K_LINEAR_MVN3; }

// per-column view of the prior table (host-built from smcb_problem_t::priors).  Scalar priors
// share one branch-free support test:  t = x - a;  lo <= t <= hi.
//   scipy uniform(loc, scale): a = loc, [0, scale] -- the reference tests (x-loc)/scale in [0,1];
//     with t = fl(x-loc) that is exactly 0 <= t <= scale (DESIGN.md "priors"), no division;
//   ImproperUniform(lo, hi):   a = 0 (t = x exactly), [lo, hi];
//   columns of a matrix prior: a = 0, [-inf, inf].
// One record per column (32 B): the d columns a kernel tests are contiguous in the constant bank -- as
// separate arrays (a[32], lo[32], hi[32]) they were 256 B apart and, with the rest of the 3.7 KB argument
// block, kept evicting one another from the per-SM constant cache (profiles/r02_kernels.md).
struct ColPrior {
    double a, lo, hi, logp;
};
struct ColPriors {
    ColPrior col[SMCB_MAX_PARAMS];
    double logp_total;                      // sum of logp in column order
    int8_t prior_of_col[SMCB_MAX_PARAMS];   // -1: column belongs to a matrix prior
    int32_t n_external;                     // number of EXTERNAL (host-evaluated) priors
    int32_t n_cov;                          // number of IMPROPER_COV priors (<= 2)
    int32_t cov_c0[2], cov_ncols[2], cov_prior[2];
};

// Kernel argument block (3.7 KB of constant bank).  What every tile of the hot loop reads comes first and
// close together (pointers, sizes, path, generator, round keys, the prior records); the problem description,
// of which a kernel reads a few scalars, goes last.
struct MhArgs {
    double* params;
    int64_t stride, n;
    double n_global;
    double* ll;
    double* lp;
    double* lq;           // proposal log-pdf column (nullptr: no proposal, lq := lp)
    uint8_t* moved;
    const double* chol;
    unsigned long long* accept_counts;
    int32_t step;
    int32_t spring_form;  // spring-mass kernels: 0 = RK4 stages evaluated at every sub-step, 1 = the RK4 propagator
    double phi, pe, qe;
    smcb_rng_t rng;
    int32_t* flags;
    double* lp_cols;      // init / priors only (optional)
    double* new_params;   // propose only
    double* new_lp;       // propose only
    PhiloxKeys rk;        // Philox round keys of rng.seed
    ColPriors cp;
    smcb_xchg_t x;        // world <= 1: no cross-GPU exchange
    unsigned long long xchg_timeout_ns;
    smcb_problem_t prob;
};

// covariance scale after the first `step` Metropolis steps (vector_mcmc.py:55-60):
// cov/5 below 30 % acceptance, cov*2 above 70 %; chol(s C) = sqrt(s) chol(C).
__device__ __forceinline__ double cov_scale_sqrt(const unsigned long long* counts, int step, double n_global) {
    double s = 1.0;
    for (int k = 0; k < step; ++k) {
        const double acc = (double)counts[k];
        if (acc < n_global * 0.3) s = s * 1 / 5;
        if (acc > n_global * 0.7) s = s * 2;
    }
    return sqrt(s);
}

// ---- cross-GPU accept-count exchange (fused collective over NVLink peer memory) ----------------
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Slot buffer of a rank (symmetric memory, smcb_mh_xchg_slot_words() words): FINAL counts [step][rank], then
// EARLY counts [step][rank].  A word is (generation << 32 | count).
//   final: the rank's accept count of the step, stored by the last block of its launch;
//   early: a LOWER BOUND of it, stored (dynamic-tile kernels) as soon as all but `slack` of the rank's particles
//          are counted, so count <= final <= count + slack.
// The covariance-scale rule (vector_mcmc.py:55-60) only asks whether the GLOBAL count is below 0.3 N, above 0.7 N
// or in between.  With every rank's count known to within its slack that question usually has ONE answer long
// before the slowest rank finishes the step (C2: ~45 % acceptance, decided at 85 % progress), so the next step
// does not wait for the peers' kernels to end plus an NVLink round trip: the ranks run the K steps of a mutation
// free of each other and meet again at the next reweight.  When the bounds straddle a threshold the reader keeps
// polling until the finals arrive -- the decision is always the one the exact count gives.
constexpr unsigned long long kCountMask = 0xffffffffull;
__device__ __forceinline__ unsigned long long xchg_early_slack(double n_global, int world) {
    return (unsigned long long)(0.15 * n_global / (double)world);
}
// a count with the same scale decision as the global accept count of step j (the exact count whenever every
// rank's final slot had arrived): local cache for earlier steps, else from the peers' slots
__device__ __forceinline__ double xchg_global_count(const smcb_xchg_t& x, int j, bool latest, double n_global,
                                                    unsigned long long timeout_ns, int32_t* flags) {
    if (!latest) return (double)x.global_counts[j];
    const unsigned long long* fin =
        reinterpret_cast<const unsigned long long*>(x.peer_slots[x.rank]) + (size_t)j * SMCB_MAX_RANKS;
    const unsigned long long* early = fin + (size_t)SMCB_MAX_MCMC_STEPS * SMCB_MAX_RANKS;
    const unsigned long long slack = xchg_early_slack(n_global, x.world);
    unsigned long long rep = 0;
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    for (;;) {
        unsigned long long lo = 0, hi = 0;
        bool ready = true, exact = true;
        // all 2 x world slot words are requested before the first one is looked at: one load latency per poll
        // instead of one per rank (a dependent chain of `world` system-scope loads was +0.6 us per rank on every
        // launch of every block: 0.1587 -> 0.1635 ms per C2 step on 8 GPUs)
        unsigned long long vf[SMCB_MAX_RANKS], ve[SMCB_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < SMCB_MAX_RANKS; ++r) {
            vf[r] = 0ull;
            ve[r] = 0ull;
            if (r < x.world) {
                vf[r] = ld_sys(fin + r);
                ve[r] = ld_sys(early + r);
            }
        }
#pragma unroll
        for (int r = 0; r < SMCB_MAX_RANKS; ++r) {
            if (r >= x.world || !ready) continue;
            const unsigned long long v = vf[r];
            if ((v >> 32) == x.gen) {
                lo += v & kCountMask;
                hi += v & kCountMask;
                continue;
            }
            const unsigned long long e = ve[r];
            if ((e >> 32) != x.gen) {
                ready = false;
                continue;
            }
            lo += e & kCountMask;
            hi += (e & kCountMask) + slack;
            exact = false;
        }
        if (ready) {
            const double dlo = (double)lo, dhi = (double)hi;
            if (exact || dlo > n_global * 0.7 || (dlo >= n_global * 0.3 && dhi <= n_global * 0.7)) {
                rep = lo;
                break;
            }
            if (dhi < n_global * 0.3) {
                rep = hi;
                break;
            }
        }
        __nanosleep(64);
        // a peer that never answers (died, or never launched the step): give up after the exchange time-out, flag it
        // (the host raises on every rank) and carry on with what is known -- the result is discarded anyway
        if ((++spins & 1023u) == 0u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > timeout_ns) {
                atomicOr(flags, SMCB_FLAG_XCHG_TIMEOUT);
                rep = lo;
                break;
            }
        }
    }
    x.global_counts[j] = (long long)rep;     // blocks may store different, decision-equivalent values
    return (double)rep;
}
// producer side (dynamic-tile kernels, thread 0 of a block that finished a tile of `cnt` particles of which `acc`
// were accepted).  One 64-bit atomicAdd per tile on the launch's progress word
//     (particles counted << 32 | proposals accepted among them)
// returns a CONSISTENT (counted, accepted) pair, so no fence is needed: the block whose tile takes `counted`
// across n - slack stores the early slot, the block whose tile completes n stores the final slot (and clears
// the word for the next launch) -- no end-of-launch ticket round either.
__device__ __forceinline__ void xchg_tile_done(const smcb_xchg_t& x, int step, long long n, double n_global,
                                               unsigned int cnt, unsigned int acc) {
    unsigned long long* prog = reinterpret_cast<unsigned long long*>(x.ticket + 2);
    const unsigned long long add = ((unsigned long long)cnt << 32) | acc;
    const unsigned long long before = atomicAdd(prog, add), after = before + add;
    const unsigned long long d0 = before >> 32, d1 = after >> 32;
    const unsigned long long slack = xchg_early_slack(n_global, x.world);
    const unsigned long long thr = (unsigned long long)n > slack ? (unsigned long long)n - slack : 0ull;
    const bool early = d1 >= thr && (d0 < thr || (thr == 0ull && d0 == 0ull));
    const bool fin = d1 == (unsigned long long)n;
    if (early || fin) {
        const unsigned long long word = (x.gen << 32) | (after & kCountMask);
        for (int r = 0; r < x.world; ++r) {
            unsigned long long* base = reinterpret_cast<unsigned long long*>(x.peer_slots[r]) + x.rank;
            if (early) st_sys(base + (size_t)(SMCB_MAX_MCMC_STEPS + step) * SMCB_MAX_RANKS, word);
            if (fin) st_sys(base + (size_t)step * SMCB_MAX_RANKS, word);
        }
        if (fin) *prog = 0ull;                      // every particle is counted: nobody adds again in this launch
    }
}
// The same scale computed by the whole block: thread k classifies step k (the accept-count loads of all
// previous steps are in flight together instead of one dependent load after another on thread 0: up to 19 x 0.6 us
// at the 20th step of a mutate call), the two class counts are reduced with barrier counts, and every thread
// applies them.  s = 2^up / 5^down with one rounding per division: multiplying by 2 is exact, so the order of the
// operations does not matter and the result is bit-identical to the sequential rule above.
// Must be called by every thread of the block.
template <int TPB>
__device__ __forceinline__ double block_cov_scale_sqrt(const MhArgs& a) {
    int up = 0, down = 0;
    for (int base = 0; base < a.step; base += TPB) {
        const int k = base + (int)threadIdx.x;
        bool u = false, dn = false;
        if (k < a.step) {
            const double acc = (a.x.world > 1) ? xchg_global_count(a.x, k, k == a.step - 1, a.n_global, a.xchg_timeout_ns, a.flags)
                                               : (double)a.accept_counts[k];
            dn = acc < a.n_global * 0.3;
            u = acc > a.n_global * 0.7;
        }
        up += __syncthreads_count(u);
        down += __syncthreads_count(dn);
    }
    double s = 1.0;
    for (int i = 0; i < down && s > 0.0; ++i) s = s * 1 / 5;
    return sqrt(scalbn(s, up));
}

// ---- priors -----------------------------------------------------------------------------------
__device__ __noinline__ double gen_uniform53(uint32_t k0, uint32_t k1, uint64_t pid, uint32_t call, uint32_t stream) {
    const uint4 r = philox_s(k0, k1, pid, call, stream);
    return u01_53(r.x, r.y);
}

// Out-of-line on purpose: the wide kernels need DMAX normals per particle and inlined copies push the
// tile body past the 32 KB instruction cache (ncu: 21 % of stall samples were `no_inst`,
// profiles/r01_mh_kernels.md).
// Sixteen normals from four Philox calls (counters call .. call + 3) in one out-of-line block: the
// four 10-round chains are independent, so the scheduler interleaves them (one call per block left
// the four warps of a scheduler waiting on one dependent IMAD chain each: 33 % `wait` stalls;
// 1.044 -> 0.999 ms with two chains, 0.984 ms with four).
struct Float16 {
    float4 a, b, c, d;
};
__device__ __noinline__ Float16 gen_normals16(uint32_t k0, uint32_t k1, uint64_t pid, uint32_t call, uint32_t stream) {
    const uint4 r0 = philox_s(k0, k1, pid, call, stream);
    const uint4 r1 = philox_s(k0, k1, pid, call + 1u, stream);
    const uint4 r2 = philox_s(k0, k1, pid, call + 2u, stream);
    const uint4 r3 = philox_s(k0, k1, pid, call + 3u, stream);
    Float16 z;
    normal_pair_f(r0.x, r0.y, z.a.x, z.a.y);
    normal_pair_f(r1.x, r1.y, z.b.x, z.b.y);
    normal_pair_f(r2.x, r2.y, z.c.x, z.c.y);
    normal_pair_f(r3.x, r3.y, z.d.x, z.d.y);
    normal_pair_f(r0.z, r0.w, z.a.z, z.a.w);
    normal_pair_f(r1.z, r1.w, z.b.z, z.b.w);
    normal_pair_f(r2.z, r2.w, z.c.z, z.c.w);
    normal_pair_f(r3.z, r3.w, z.d.z, z.d.w);
    return z;
}

// Cholesky pivots and inverse diagonal of a symmetric 3 x 3 matrix (unblocked, LAPACK dpotf2 failure rule:
// pivot <= 0 or NaN).  One rsqrt per pivot, no division: l_ij = c_ij * (1 / sqrt(pivot)).
struct Chol3 {
    double i00, i11, i22, l10, l20, l21, p00, p11, p22;
    bool pd;
};
__device__ __forceinline__ Chol3 chol3(const double (&c)[3][3]) {
    Chol3 f;
    f.p00 = c[0][0];
    f.i00 = rsqrt(f.p00);
    f.l10 = c[1][0] * f.i00;
    f.l20 = c[2][0] * f.i00;
    f.p11 = c[1][1] - f.l10 * f.l10;
    f.i11 = rsqrt(f.p11);
    f.l21 = (c[2][1] - f.l20 * f.l10) * f.i11;
    f.p22 = c[2][2] - f.l20 * f.l20 - f.l21 * f.l21;
    f.i22 = rsqrt(f.p22);
    f.pd = (f.p00 > 0.0) && (f.p11 > 0.0) && (f.p22 > 0.0);
    return f;
}

template <int KIND, int DMAX, bool EXACT>
__device__ __forceinline__ double eval_priors(const MhArgs& a, const double (&th)[DMAX], double* lp_cols_row,
                                              int64_t col_stride) {
    const int d = a.prob.n_params;
    bool ok = true;
#pragma unroll
    for (int c = 0; c < DMAX; ++c) {
        if (EXACT || c < d) {
            const double t = th[c] - a.cp.col[c].a;
            ok = ok && (t >= a.cp.col[c].lo) && (t <= a.cp.col[c].hi);
        }
    }
    double lps = ok ? a.cp.logp_total : kNegInf;
    if (lp_cols_row) {                      // un-summed per-prior values (host-facing API only)
        for (int c = 0; c < d; ++c) {
            if (a.cp.prior_of_col[c] < 0) continue;
            double x = 0.0;
#pragma unroll
            for (int k = 0; k < DMAX; ++k) x = (k == c) ? th[k] : x;
            const double t = x - a.cp.col[c].a;
            lp_cols_row[(int64_t)a.cp.prior_of_col[c] * col_stride] =
                (t >= a.cp.col[c].lo && t <= a.cp.col[c].hi) ? a.cp.col[c].logp : kNegInf;
        }
    }
    // matrix priors: packed upper-triangular columns must form a Cholesky-PD matrix.
    // Compiled for the MVNormal kind only: the runtime column selects would otherwise push
    // th[] into local memory for every kernel.
    if constexpr (is_mvn(KIND))
    for (int q = 0; q < a.cp.n_cov; ++q) {
        const int c0 = a.cp.cov_c0[q], nc = a.cp.cov_ncols[q];
        double m[3][3];
        int e = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = r; s < 3; ++s) {
                double v = (r == s) ? 1.0 : 0.0;
                if constexpr (KIND == K_LINEAR_MVN3) {
                    v = th[(2 + e < DMAX) ? 2 + e : 0];
                    ++e;
                } else if (r < nc && s < nc) {
                    v = 0.0;
#pragma unroll
                    for (int c = 0; c < DMAX; ++c) v = (c == c0 + e) ? th[c] : v;
                    ++e;
                }
                m[r][s] = v;
                m[s][r] = v;
            }
        const bool pd = chol3(m).pd;
        const double v = pd ? 0.0 : kNegInf;
        lps += v;
        if (lp_cols_row) lp_cols_row[(int64_t)a.cp.cov_prior[q] * col_stride] = v;
    }
    return lps;
}

// proposal log-pdf log q(x) = sum of independent components (smcpy/proposals.py:27-30 over
// scipy.stats frozen norm / uniform: norm.logpdf = -y^2/2 - log(sqrt(2 pi)) - log(scale))
template <int DMAX, bool EXACT>
__device__ __forceinline__ double eval_proposal(const MhArgs& a, const double (&th)[DMAX]) {
    const int d = a.prob.n_params;
    double lq = 0.0;
#pragma unroll
    for (int c = 0; c < DMAX; ++c) {
        if (EXACT || c < d) {
            const smcb_proposal_col_t& q = a.prob.proposal[c];
            const double y = (th[c] - q.a) / q.b;
            double v;
            if (q.kind == SMCB_PROPOSAL_NORMAL) v = (-(y * y) / 2.0 - 0.91893853320467274178) - log(q.b);
            else v = (y >= 0.0 && y <= 1.0) ? -log(q.b) : kNegInf;
            lq += v;
        }
    }
    return lq;
}

// ---- models + likelihoods ------------------------------------------------------------------------
// All threads of the block call these.  When the data vector fits one chunk it is staged once
// per block (`resident`); otherwise it is re-staged chunk by chunk for every tile.

__device__ __forceinline__ double normal_ll(double ssq, double sigma, int m) {
    // log_likelihoods.py:42-49: term1 = -log(2 pi var) * (M/2); term2 = -1/2 * ssqe / var
    const double var = sigma * sigma;
    const double term1 = -log(2 * 3.141592653589793 * var) * ((double)m / 2.0);
    const double term2 = -1 / 2.0 * ssq / var;
    return term1 + term2;
}

template <int KIND, bool CENTERED = false>
__device__ __forceinline__ void stage_data(const MhArgs& a, double* sdata, int base, int cnt) {
    if (KIND == K_LINEAR_NORMAL) {
        double2* sxy = reinterpret_cast<double2*>(sdata);
        if (CENTERED) {
            // (x_j - xbar, -(y_j - ybar)) pairs prepared once per problem by smcb_linear_center
            const double2* cen = reinterpret_cast<const double2*>(a.prob.lin_centered + 2) + base;
            const int bd = blockDim.x;
            for (int j = threadIdx.x; j < cnt; j += 4 * bd) {       // four independent loads in flight per thread
                double2 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = (j + k * bd < cnt) ? cen[j + k * bd] : make_double2(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (j + k * bd < cnt) sxy[j + k * bd] = v[k];
            }
        } else {
            for (int j = threadIdx.x; j < cnt; j += blockDim.x)
                sxy[j] = make_double2(a.prob.xgrid[base + j], a.prob.data[base + j]);
        }
    } else if (is_mvn(KIND)) {
        // [0..2] mean of the snapshots, [3..11] their scatter matrix (3 x 3, zero padded), [12..14] X
        const int F = a.prob.n_data;
        for (int j = threadIdx.x; j < kMvnStats; j += blockDim.x) sdata[j] = a.prob.mvn_stats[j];
        for (int j = threadIdx.x; j < 3; j += blockDim.x) sdata[kMvnStats + j] = (j < F) ? a.prob.xgrid[j] : 0.0;
    } else {
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) sdata[j] = a.prob.data[base + j];
        if (KIND == K_IDENTITY_NORMAL && threadIdx.x == 0) {       // constants of normal_ll, slots past DMAX <= 32
            const double var = a.prob.sigma * a.prob.sigma;
            sdata[SMCB_MAX_PARAMS] = -log(2 * 3.141592653589793 * var) * ((double)a.prob.n_data / 2.0);
            sdata[SMCB_MAX_PARAMS + 1] = var;
        }
    }
}

// Linear model + Normal likelihood for PPT particles per thread: every staged data point
// (one broadcast LDS.128) feeds 2*PPT FP64 instructions, which keeps the shared-memory pipe
// off the critical path (profiles/r01_mh_kernels.md).
// The residuals are accumulated about the data means: with u_j = x_j - xbar, v_j = y_j - ybar (sum u = sum v = 0)
//     sum_j (a x_j + b - y_j)^2 = sum_j (a u_j - v_j)^2 + M (a xbar + b - ybar)^2,
// the cross term vanishing identically.  The model is still evaluated at every data point and the squared
// residuals are still reduced over the whole vector, but the intercept is applied once instead of M times:
// two FP64 instructions per point (fma, fma) instead of three (fma, sub, fma), and no cancellation between
// b and y_j.  Within a few ulp of the reference's direct form (smcpy/log_likelihoods.py:42-49); pinned at
// 1e-10 by the replay tests.
template <int DMAX, int PPT>
__device__ __forceinline__ void ll_linear_normal(const MhArgs& a, const double (&th)[PPT][DMAX], bool run_any,
                                                 bool resident, double* sdata, double (&ll)[PPT]) {
    const double2* sxy = reinterpret_cast<const double2*>(sdata);      // (u_j, -v_j)
    const int m = a.prob.n_data;
    double s0[PPT], s1[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) s0[p] = s1[p] = 0.0;
    for (int base = 0; base < m; base += kChunk) {
        const int cnt = (m - base < kChunk) ? (m - base) : kChunk;
        if (!resident) {
            __syncthreads();
            stage_data<K_LINEAR_NORMAL, true>(a, sdata, base, cnt);
            __syncthreads();
        }
        if (run_any) {
            int j = 0;
#pragma unroll 8
            for (; j + 2 <= cnt; j += 2) {
                const double2 p0 = sxy[j], p1 = sxy[j + 1];
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const double d0 = fma(th[p][0], p0.x, p0.y);
                    const double d1 = fma(th[p][0], p1.x, p1.y);
                    s0[p] = fma(d0, d0, s0[p]);
                    s1[p] = fma(d1, d1, s1[p]);
                }
            }
            for (; j < cnt; ++j) {
                const double2 q = sxy[j];
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const double dd = fma(th[p][0], q.x, q.y);
                    s0[p] = fma(dd, dd, s0[p]);
                }
            }
        }
    }
    const double xbar = a.prob.lin_centered[0], ybar = a.prob.lin_centered[1];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const double sigma = a.prob.sigma_is_param ? th[p][DMAX > 2 ? 2 : 0] : a.prob.sigma;
        const double c = fma(th[p][0], xbar, th[p][1]) - ybar;
        ll[p] = normal_ll(fma((double)m * c, c, s0[p] + s1[p]), sigma, m);
    }
}

// Spring-mass ODE (classical RK4 on (x, v), nsub sub-steps per output point) + Normal likelihood for PPT
// particles per thread: the sub-steps of one particle are a dependent DFMA/DADD chain, so PPT independent
// chains are interleaved to keep the FP64 pipe fed (profiles/r01_c3_c4_kernels.md: 35 % `math` stalls at
// PPT = 1).  Per particle the arithmetic (operations and their order) does not depend on PPT.
template <int DMAX, int PPT>
__device__ __forceinline__ void ll_spring_normal(const MhArgs& a, const double (&th)[PPT][DMAX], bool run_any,
                                                 bool resident, double* sy, double (&ll)[PPT]) {
    const int m = a.prob.n_data;
    double K[PPT], g[PPT], x[PPT], v[PPT], ssq[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        K[p] = th[p][0];
        g[p] = th[p][1];
        x[p] = a.prob.model_args[0];
        v[p] = a.prob.model_args[1];
        ssq[p] = 0.0;
    }
    const int nsub = (int)a.prob.model_args[3];
    const double h = a.prob.model_args[2] / (double)nsub;
    const double hh = 0.5 * h, h6 = h / 6.0;
    // RK4 propagator form (a.spring_form == 1).  The ODE x'' = -K x + g is linear with constant coefficients,
    // y' = A y + b, y = (x, v), A = [[0, 1], [-K, 0]], b = (0, g), so one classical RK4 step of size h IS the affine map
    //     y <- P(hA) y + h Q(hA) b,   P(z) = 1 + z + z^2/2 + z^3/6 + z^4/24,   Q(z) = 1 + z/2 + z^2/6 + z^3/24
    // (what the four stages add up to, exactly), and with (hA)^2 = -w I, w = h^2 K:
    //     P(hA) = p I + q hA,  p = 1 - w/2 + w^2/24,  q = 1 - w/6;   h Q(hA) b = h ((1/2 - w/24) h g,  q g).
    // Every power of P(hA) is again alpha I + beta hA, so the nsub sub-steps between two observations compose into
    //     x <- alpha x + bh v + c0,   v <- alpha v - K bh x + c1
    // with five coefficients per particle (2 x 2 complex-like products, nsub of them, once per particle and step).
    // The trajectory is the RK4 trajectory -- same method, same step schedule, every observation still visited in
    // turn -- at 4 FMAs per observation instead of 18 FP64 instructions per sub-step; it differs from the staged
    // evaluation by rounding only (1.5e-13 relative over the 1996 sub-steps of config C3; the replay tests hold it
    // to the same 1e-10 against the oracle's staged NumPy RK4).
    if (a.spring_form == 1) {
        double al[PPT], bh[PPT], kbh[PPT], c0[PPT], c1[PPT];
        // (every product and sum below is an explicitly rounded mul / fma, so that the coefficients -- and with them
        // every log-likelihood -- do not depend on how the compiler contracts the expressions of an instantiation)
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const double w = __dmul_rn(__dmul_rn(h, h), K[p]);
            const double p1 = fma(__dmul_rn(w, w), 1.0 / 24.0, fma(-0.5, w, 1.0));        // one step: p1 I + q1 hA
            const double q1 = fma(-1.0 / 6.0, w, 1.0);
            const double qh = __dmul_rn(q1, h), qhk = __dmul_rn(qh, K[p]);
            const double d0 = __dmul_rn(__dmul_rn(__dmul_rn(fma(-1.0 / 24.0, w, 0.5), h), g[p]), h);
            const double d1 = __dmul_rn(qh, g[p]);
            double A = p1, B = q1, e0 = d0, e1 = d1;
            for (int s = 1; s < nsub; ++s) {
                // constant term first (it uses the ONE-step map): c <- P1 c + d;  then (A + B z)(p1 + q1 z), z^2 = -w
                const double n0 = fma(p1, e0, fma(qh, e1, d0));
                const double n1 = fma(p1, e1, fma(-qhk, e0, d1));
                e0 = n0;
                e1 = n1;
                const double An = fma(A, p1, -__dmul_rn(w, __dmul_rn(B, q1)));
                const double Bn = fma(A, q1, __dmul_rn(B, p1));
                A = An;
                B = Bn;
            }
            al[p] = A;
            bh[p] = __dmul_rn(B, h);
            kbh[p] = -__dmul_rn(bh[p], K[p]);
            c0[p] = e0;
            c1[p] = e1;
        }
        for (int base = 0; base < m; base += kChunk) {
            const int cnt = (m - base < kChunk) ? (m - base) : kChunk;
            if (!resident) {
                __syncthreads();
                stage_data<K_SPRING_NORMAL>(a, sy, base, cnt);
                __syncthreads();
            }
            if (run_any) {
                int j = 0;
                if (base == 0) {                  // the first observation is the initial state: no step before it
                    const double y = sy[0];
#pragma unroll
                    for (int p = 0; p < PPT; ++p) {
                        const double dd = __dsub_rn(x[p], y);
                        ssq[p] = fma(dd, dd, ssq[p]);
                    }
                    j = 1;
                }
#pragma unroll 4
                for (; j < cnt; ++j) {
                    const double y = sy[j];
#pragma unroll
                    for (int p = 0; p < PPT; ++p) {
                        const double xn = fma(al[p], x[p], fma(bh[p], v[p], c0[p]));
                        const double vn = fma(al[p], v[p], fma(kbh[p], x[p], c1[p]));
                        x[p] = xn;
                        v[p] = vn;
                        const double dd = __dsub_rn(xn, y);
                        ssq[p] = fma(dd, dd, ssq[p]);
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const double sigma = a.prob.sigma_is_param ? th[p][DMAX > 2 ? 2 : 0] : a.prob.sigma;
            ll[p] = normal_ll(ssq[p], sigma, m);
        }
        return;
    }
    for (int base = 0; base < m; base += kChunk) {
        const int cnt = (m - base < kChunk) ? (m - base) : kChunk;
        if (!resident) {
            __syncthreads();
            stage_data<K_SPRING_NORMAL>(a, sy, base, cnt);
            __syncthreads();
        }
        if (run_any) {
            for (int j = 0; j < cnt; ++j) {
                if (base + j > 0) {
                    for (int s = 0; s < nsub; ++s) {
#pragma unroll
                        for (int p = 0; p < PPT; ++p) {
                            const double k1x = v[p], k1v = g[p] - K[p] * x[p];
                            const double x2 = x[p] + hh * k1x, v2 = v[p] + hh * k1v;
                            const double k2x = v2, k2v = g[p] - K[p] * x2;
                            const double x3 = x[p] + hh * k2x, v3 = v[p] + hh * k2v;
                            const double k3x = v3, k3v = g[p] - K[p] * x3;
                            const double x4 = x[p] + h * k3x, v4 = v[p] + h * k3v;
                            const double k4x = v4, k4v = g[p] - K[p] * x4;
                            x[p] = x[p] + h6 * (k1x + 2.0 * k2x + 2.0 * k3x + k4x);
                            v[p] = v[p] + h6 * (k1v + 2.0 * k2v + 2.0 * k3v + k4v);
                        }
                    }
                }
                const double y = sy[j];
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const double dd = x[p] - y;
                    ssq[p] = fma(dd, dd, ssq[p]);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const double sigma = a.prob.sigma_is_param ? th[p][DMAX > 2 ? 2 : 0] : a.prob.sigma;
        ll[p] = normal_ll(ssq[p], sigma, m);
    }
}

// sy[m] = term1 = -log(2 pi var) * (M/2), sy[m+1] = -1/(2 var): computed once per block by
// stage_data (the identity model always has a fixed sigma), same arithmetic as normal_ll
template <int DMAX, bool EXACT>
__device__ __forceinline__ double ll_identity_normal(const MhArgs& a, const double (&th)[DMAX], bool run,
                                                     const double* sy) {
    const int m = a.prob.n_data;          // == n_model_params <= DMAX
    double s0 = 0.0, s1 = 0.0;
    if (run) {
#pragma unroll
        for (int j = 0; j < DMAX; j += 2) {
            if (EXACT || j < m) { const double dd = th[j] - sy[j]; s0 = fma(dd, dd, s0); }
            if (j + 1 < DMAX && (EXACT || j + 1 < m)) { const double dd = th[j + 1] - sy[j + 1]; s1 = fma(dd, dd, s1); }
        }
    }
    return sy[SMCB_MAX_PARAMS] + (-1 / 2.0 * (s0 + s1)) / sy[SMCB_MAX_PARAMS + 1];
}

// MVNormal (log_likelihoods.py:150-192) with one Cholesky per particle:
//   sum_s [ -F/2 log 2pi - 1/2 log det C - 1/2 e_s^T C^-1 e_s ],  e_s = model - data_s.
// The model output is the same for every snapshot, so with ybar = mean_s data_s and the scatter matrix
// W = sum_s (data_s - ybar)(data_s - ybar)^T (smcb_mvn_scatter, once per problem)
//   sum_s e_s^T C^-1 e_s = S (model - ybar)^T C^-1 (model - ybar) + tr(C^-1 W):
// O(F^2) per particle instead of O(S F^2), whatever the model.
template <int KIND, int DMAX>
__device__ __forceinline__ double ll_linear_mvn(const MhArgs& a, const double (&th)[DMAX], bool run,
                                                const double* sdat, bool& not_pd) {
    const int F = (KIND == K_LINEAR_MVN3) ? 3 : a.prob.n_data, S = a.prob.n_snapshots;
    if (!run) return 0.0;
    const double* W = sdat + 3;
    const double* X = sdat + kMvnStats;
    double c[3][3];
    int e = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = r; s < 3; ++s) {
            double v = (r == s) ? 1.0 : 0.0;
            if constexpr (KIND == K_LINEAR_MVN3) {
                v = th[(2 + e < DMAX) ? 2 + e : 0];
                ++e;
            } else if (r < F && s < F) {
                v = a.prob.cov_fixed[e];
                const int col = a.prob.cov_param_col[e];
#pragma unroll
                for (int k = 0; k < DMAX; ++k) v = (k == col) ? th[k] : v;
                ++e;
            }
            c[r][s] = v;
            c[s][r] = v;
        }
    const Chol3 f = chol3(c);
    if (!f.pd) not_pd = true;
    const double i00 = f.i00, i11 = f.i11, i22 = f.i22, l10 = f.l10, l20 = f.l20, l21 = f.l21;
    // log det C = log of the product of the pivots (one log; the sum of three when the product leaves the
    // normal range)
    const double pprod = f.p00 * f.p11 * f.p22;
    const double logdet = (pprod > 1e-280 && pprod < 1e280) ? log(pprod) : (log(f.p00) + log(f.p11) + log(f.p22));
    // S (model - ybar)^T C^-1 (model - ybar): one forward substitution
    const double e0 = fma(th[0], X[0], th[1]) - sdat[0];
    const double e1 = (F > 1) ? fma(th[0], X[1], th[1]) - sdat[1] : 0.0;
    const double e2 = (F > 2) ? fma(th[0], X[2], th[1]) - sdat[2] : 0.0;
    const double y0 = e0 * i00;
    const double y1 = (e1 - l10 * y0) * i11;
    const double y2 = (e2 - l20 * y0 - l21 * y1) * i22;
    double quad = (double)S * (y0 * y0 + y1 * y1 + y2 * y2);
    // tr(C^-1 W) = tr(L^-1 W L^-T): Z = L^-1 W column by column, then the diagonal of L^-1 Z^T
    double z[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        z[0][k] = W[k] * i00;
        z[1][k] = (W[3 + k] - l10 * z[0][k]) * i11;
        z[2][k] = (W[6 + k] - l20 * z[0][k] - l21 * z[1][k]) * i22;
    }
    const double m00 = z[0][0] * i00;
    const double v10 = z[1][0] * i00;
    const double m11 = (z[1][1] - l10 * v10) * i11;
    const double v20 = z[2][0] * i00;
    const double v21 = (z[2][1] - l10 * v20) * i11;
    const double m22 = (z[2][2] - l20 * v20 - l21 * v21) * i22;
    quad += m00 + m11 + m22;
    const double term1 = -(double)F / 2 * log(2 * 3.141592653589793);
    return (double)S * (term1 - 0.5 * logdet) - 0.5 * quad;
}

// register budget per thread -> minimum resident blocks per SM (keeps ptxas from trading
// occupancy for unrolling: profiles/r01_mh_kernels.md)
#ifndef SMCB_PPT4_REGS
#define SMCB_PPT4_REGS 84
#endif
__host__ __device__ constexpr int mh_min_blocks(int dmax, int tpb, int ppt, bool tma = false, int kind = -1) {
#ifdef SMCB_TMA_REGS
    const int tma_regs = SMCB_TMA_REGS;
#else
    const int tma_regs = 128;
#endif
    // MVNormal (3 x 3 factorisation + solves) and two interleaved RK4 chains need ~80 registers to stay
    // out of local memory; both are FP64-bound, so 6 blocks of 128 threads per SM are plenty
    const bool wide = (is_mvn(kind) && dmax >= 8) || (kind == K_SPRING_NORMAL && ppt >= 2);
    const int regs = dmax >= 32 ? (tma ? tma_regs : 128) : (dmax >= 16 ? 96 : (wide ? 80 : (ppt >= 4 ? SMCB_PPT4_REGS : 64)));
    return 65536 / (tpb * regs) > 0 ? 65536 / (tpb * regs) : 1;
}

// MultiSourceNormal (log_likelihoods.py:52-118) for the three built-in models: consecutive data
// segments, each with its own (fixed or sampled) std-dev; data resident in shared memory.
// Generic one-point-at-a-time loop: this likelihood is a coverage feature, not a tuned path.
template <int KIND, int DMAX>
__device__ __forceinline__ double ll_multisource(const MhArgs& a, const double (&th)[DMAX], bool run,
                                                 const double* sdata) {
    if (!run) return 0.0;
    const double2* sxy = reinterpret_cast<const double2*>(sdata);
    double x = a.prob.model_args[0], v = a.prob.model_args[1];          // spring-mass state
    const int nsub = (int)a.prob.model_args[3];
    const double h = (KIND == K_SPRING_NORMAL) ? a.prob.model_args[2] / (double)nsub : 0.0;
    double ll = 0.0;
    int j = 0;
    for (int sgm = 0; sgm < a.prob.n_segments; ++sgm) {
        const int len = a.prob.seg_len[sgm];
        if (len == 0) continue;
        double sigma = a.prob.seg_sigma[sgm];
        const int col = a.prob.seg_sigma_col[sgm];
#pragma unroll
        for (int k = 0; k < DMAX; ++k) sigma = (k == col) ? th[k] : sigma;
        double ssq = 0.0;
        for (int e = j + len; j < e; ++j) {
            double out;
            if (KIND == K_LINEAR_NORMAL) {
                out = fma(th[0], sxy[j].x, th[1]);
            } else if (KIND == K_SPRING_NORMAL) {
                if (j > 0) {
                    for (int s = 0; s < nsub; ++s) {
                        const double K = th[0], g = th[1], hh = 0.5 * h, h6 = h / 6.0;
                        const double k1x = v, k1v = g - K * x;
                        const double x2 = x + hh * k1x, v2 = v + hh * k1v;
                        const double k2x = v2, k2v = g - K * x2;
                        const double x3 = x + hh * k2x, v3 = v + hh * k2v;
                        const double k3x = v3, k3v = g - K * x3;
                        const double x4 = x + h * k3x, v4 = v + h * k3v;
                        const double k4x = v4, k4v = g - K * x4;
                        x = x + h6 * (k1x + 2.0 * k2x + 2.0 * k3x + k4x);
                        v = v + h6 * (k1v + 2.0 * k2v + 2.0 * k3v + k4v);
                    }
                }
                out = x;
            } else {
                out = 0.0;
#pragma unroll
                for (int k = 0; k < DMAX; ++k) out = (k == j) ? th[k] : out;
            }
            const double y = (KIND == K_LINEAR_NORMAL) ? sxy[j].y : sdata[j];
            const double dd = out - y;
            ssq = fma(dd, dd, ssq);
        }
        ll += normal_ll(ssq, sigma, len);
    }
    return ll;
}

// mode: 0 = init (evaluate incoming particles), 1 = fused Metropolis step, 2 = propose only,
//       3 = log-priors only.  EXACT: n_params == DMAX (drops every per-column bounds check).
//       TMA: the next tile's d+2 column segments are bulk-copied (cp.async.bulk + mbarrier) into
//       a shared-memory stage by one elected thread while the block computes the current tile.
//       PROP: the path has a proposal q: its log-pdf is a third resident column next to ll, lp.
//       DYN: persistent blocks pull tiles from an atomic counter (a.flags[1]; a.flags[2] is the
//       ticket with which the last block resets both) -- removes the tail of the FP64-bound kernels.
//       MULTI: MultiSourceNormal likelihood (segment table in the problem).
//       STRICT (with TMA): the increment sqrt(s) L z is accumulated with fp64 FMAs instead of the tf32
//       tensor-core product (option "strict_proposal"); same normals, same Philox counters.
//       PREF (persistent static tiles, PPT = 1, MODE 1): every thread prefetches its own d + 2 values of the
//       block's NEXT tile into thread-private shared slots with cp.async (LDGSTS) while it computes the
//       current tile -- hides the column-load latency of the narrow bandwidth-bound kernels (C4) without a
//       block barrier.
template <int KIND, int DMAX, int MODE, int TPB, int PPT, bool EXACT, bool TMA = false, bool PROP = false,
          bool DYN = false, bool MULTI = false, bool STRICT = false, bool PREF = false>
__global__ void __launch_bounds__(TPB, mh_min_blocks(DMAX, TPB, PPT, TMA, KIND))
mh_kernel(const __grid_constant__ MhArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t mbar, mbar_empty;
    const int d = EXACT ? DMAX : a.prob.n_params;
    // Cholesky factor transposed and zero-padded to DMAX: sLt[j * DMAX + r] = L[r][j], so column j
    // is contiguous (LDS.128) and the unrolled update needs no bounds checks
    double* sLt = smem;
    double* sdata = smem + DMAX * DMAX;      // 16-byte aligned data staging area

    const bool resident = (KIND == K_IDENTITY_NORMAL) || is_mvn(KIND) || (a.prob.n_data <= kChunk);
    if (MODE == 1) {
        // Programmatic dependent launch (smcb_mh_steps chains its K launches with it): let the next step's grid be
        // scheduled as this one's blocks retire, and stage the problem data -- constant for the whole run, written
        // by nothing on the stream -- BEFORE waiting for the previous step's grid to complete and flush.  The next
        // launch's latency, block start-up and data staging then overlap the previous step's tail.  Both
        // instructions are no-ops in a launch without the attribute.
        asm volatile("griddepcontrol.launch_dependents;");
        if (resident) stage_data<KIND, !MULTI>(a, sdata, 0, a.prob.n_data);
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (MODE == 1 || MODE == 2) {
        for (int e = threadIdx.x; e < DMAX * DMAX; e += TPB) {
            const int j = e / DMAX, r = e - j * DMAX;
            const double v = (r < d && j <= r) ? a.chol[r * d + j] : 0.0;
            if (!TMA || STRICT) sLt[e] = v;
        }
        if (TMA && !STRICT) {
            // TMA variant: the increment L z goes through the tensor cores (mma.sync m16n8k8, tf32
            // inputs, fp32 accumulation).  B fragments of the lower-triangular 8x8 blocks of L, in
            // fragment order: block q = (nb, kb <= nb), lane (g, t) holds L[nb*8+g][kb*8+t], [..][kb*8+t+4]
            float* sB = reinterpret_cast<float*>(sLt);
            constexpr int NB = DMAX >= 8 ? DMAX / 8 : 1;
            for (int e = threadIdx.x; e < NB * (NB + 1) / 2 * 64; e += TPB) {
                const int q = e >> 6, ln = (e >> 1) & 31, hi = e & 1;
                int nb = 0, rem = q;
                while (rem > nb) { rem -= nb + 1; ++nb; }          // q = nb (nb + 1) / 2 + kb
                const int kb = rem, g = ln >> 2, t = ln & 3;
                const int r = nb * 8 + g, j = kb * 8 + t + 4 * hi;
                sB[e] = (r < d && j <= r) ? (float)a.chol[r * d + j] : 0.0f;
            }
        }
    }
    if (MODE == 0 && resident) stage_data<KIND, !MULTI>(a, sdata, 0, a.prob.n_data);
    double sc = 0.0;
    if (MODE == 1 || MODE == 2) sc = block_cov_scale_sqrt<TPB>(a);
    __shared__ long long sh_tile;                  // DYN: the tile the block works on (-1: none yet)
    __shared__ unsigned int sh_acc;                // DYN, several GPUs: proposals accepted in that tile
    if (DYN && threadIdx.x == 0) {
        sh_tile = -1;
        sh_acc = 0u;
    }
    __syncthreads();

    const int64_t per_tile = (int64_t)TPB * PPT;
    const int64_t n_tiles = (a.n + per_tile - 1) / per_tile;
    int flag = 0;
    int n_acc = 0;
    // TMA stage: (d + 2) column segments of TPB doubles, after the data staging area
    double* stage = sdata + 64;
    uint32_t tma_phase = 0;
    auto tma_issue = [&](int64_t tile) {          // one elected thread
        const int64_t base = tile * per_tile;
        const uint32_t cnt = (uint32_t)((a.n - base < per_tile) ? (a.n - base) : per_tile);
        const uint32_t bytes = cnt * 8u;
        fence_proxy_async();                       // generic-proxy reads of the stage are done
        mbar_expect_tx(&mbar, bytes * (uint32_t)(d + 2));
#pragma unroll 1
        for (int c = 0; c < d; ++c) tma_load_1d(stage + c * TPB, a.params + (int64_t)c * a.stride + base, bytes, &mbar);
        tma_load_1d(stage + d * TPB, a.ll + base, bytes, &mbar);
        tma_load_1d(stage + (d + 1) * TPB, a.lp + base, bytes, &mbar);
    };
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&mbar, 1);
            mbar_init(&mbar_empty, TPB);
        }
        __syncthreads();
        if (threadIdx.x == 0 && (int64_t)blockIdx.x < n_tiles) tma_issue(blockIdx.x);
    }
    // PREF stage: (d + 2) x TPB doubles after the data staging area, slot [c][threadIdx.x] is private to the thread
    double* pstage = sdata + 64;
    auto pref_issue = [&](int64_t tile) {
        int64_t i = tile * per_tile + threadIdx.x;
        if (i >= a.n) i = 0;                                   // ragged end: any valid row (the result is not used)
        const double* pc = a.params + i;
#pragma unroll
        for (int c = 0; c < DMAX; ++c) {
            if (EXACT || c < d) cp_async8(pstage + c * TPB + threadIdx.x, pc);
            pc += a.stride;
        }
        cp_async8(pstage + DMAX * TPB + threadIdx.x, a.ll + i);
        cp_async8(pstage + (DMAX + 1) * TPB + threadIdx.x, a.lp + i);
        cp_async_commit();
    };
    if (PREF && (int64_t)blockIdx.x < n_tiles) pref_issue(blockIdx.x);
    // (fetching the next tile at the START of the current one, to hide the atomic's latency, was measured 15 %
    // slower for C2 and is not done)
    for (int64_t tile = blockIdx.x;; tile += gridDim.x) {
        if (DYN) {
            if (MODE == 1 && a.x.world > 1) {
                // several GPUs: count the finished tile's accepted proposals now (not at the end of the launch),
                // so that thread 0 can report the tile below with its block total
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
                if ((threadIdx.x & 31) == 0 && n_acc) {
                    atomicAdd(a.accept_counts + a.step, (unsigned long long)n_acc);
                    atomicAdd(&sh_acc, (unsigned int)n_acc);
                }
                n_acc = 0;
            }
            __syncthreads();                       // previous tile fully consumed sh_tile
            if (threadIdx.x == 0) {
                // the tile fetch first: the two atomics are independent and travel together
                const long long next = (long long)atomicAdd(reinterpret_cast<unsigned int*>(a.flags) + 1, 1u);
                if (MODE == 1 && a.x.world > 1 && sh_tile >= 0) {       // sh_tile still names the tile just finished
                    const int64_t left = a.n - (int64_t)sh_tile * per_tile;
                    xchg_tile_done(a.x, a.step, a.n, a.n_global, (unsigned int)(left < per_tile ? left : per_tile), sh_acc);
                    sh_acc = 0u;
                }
                sh_tile = next;
            }
            __syncthreads();
            tile = sh_tile;
        }
        if (tile >= n_tiles) break;
        int64_t idx[PPT];
        bool valid[PPT];
        double th[PPT][DMAX];
        double ll_old[PPT], lp_old[PPT], lp_new[PPT], ll_new[PPT];
        double lq_old[PPT], lq_new[PPT];
        bool prior_ok[PPT];
        // linear model, several particles per thread: the current ll / lp are only needed by the acceptance test --
        // loading them after the pass over the data keeps 4 PPT registers out of the FP64 loop (with them live ptxas
        // schedules every residual right before its square: a dependent DFMA pair per point)
        constexpr bool kLateOld = (KIND == K_LINEAR_NORMAL && PPT >= 4 && MODE == 1 && !TMA && !PREF && !PROP);
        if (TMA) {
            // native RNG only.  The proposal increment sqrt(s) L z is a (particles x d) x (d x d) product:
            // tensor cores, tf32 inputs (the mma ignores the low mantissa bits: truncation, exactly
            // odd in z, so the proposal stays symmetric -- all the Metropolis ratio needs), fp32
            // accumulation.  Done BEFORE the tile is read from the TMA stage, so the stage wait hides
            // behind it.
            const int64_t i = tile * per_tile + threadIdx.x;
            valid[0] = i < a.n;
            idx[0] = valid[0] ? i : 0;
            const uint64_t pid = (uint64_t)(a.rng.id_offset + idx[0]);
            float delta[STRICT ? 1 : DMAX];                              // first sqrt(s) z, then sqrt(s) L z
            if constexpr (!STRICT) {
            const float* sB = reinterpret_cast<const float*>(sLt);      // tf32 B fragments of L (see above)
            constexpr int NB = DMAX >= 8 ? DMAX / 8 : 1;
            constexpr int kZLd = DMAX + 4;                               // padded row: conflict-free LDS.128 / fragment loads
            float* zb = reinterpret_cast<float*>(stage + (DMAX + 2) * TPB) + (threadIdx.x >> 5) * 16 * kZLd;
            const int lane = threadIdx.x & 31, fg = lane >> 2, ft = lane & 3;
            const float scf = (float)sc;
#pragma unroll
            for (int c = 0; c + 15 < DMAX; c += 16) {
                const Float16 zz = gen_normals16(a.rk.k[0], a.rk.k[1], pid, (uint32_t)(c >> 2), (uint32_t)a.rng.stream);
                const float4 q4[4] = {zz.a, zz.b, zz.c, zz.d};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    delta[c + 4 * q + 0] = scf * q4[q].x;
                    delta[c + 4 * q + 1] = scf * q4[q].y;
                    delta[c + 4 * q + 2] = scf * q4[q].z;
                    delta[c + 4 * q + 3] = scf * q4[q].w;
                }
            }
            // Each warp multiplies its 32 rows of z by L^T as two 16-row M-tiles.  The rows pass
            // through a 16-row shared buffer per warp to get from one-particle-per-thread into the
            // mma fragment layout and back: z -> A fragments, D fragments -> increments.
            float dfr[2][NB][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                if ((lane >> 4) == mt) {
#pragma unroll
                    for (int c = 0; c < DMAX; c += 4)
                        *reinterpret_cast<float4*>(&zb[(lane & 15) * kZLd + c]) =
                            make_float4(delta[c], delta[c + 1], delta[c + 2], delta[c + 3]);
                }
                __syncwarp();
                uint32_t afr[NB][4];
#pragma unroll
                for (int kb = 0; kb < NB; ++kb) {
                    afr[kb][0] = __float_as_uint(zb[fg * kZLd + kb * 8 + ft]);
                    afr[kb][1] = __float_as_uint(zb[(fg + 8) * kZLd + kb * 8 + ft]);
                    afr[kb][2] = __float_as_uint(zb[fg * kZLd + kb * 8 + ft + 4]);
                    afr[kb][3] = __float_as_uint(zb[(fg + 8) * kZLd + kb * 8 + ft + 4]);
                }
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
#pragma unroll
                    for (int kb = 0; kb <= nb; ++kb) {                    // blocks above the diagonal of L are zero
                        const float2 b = *reinterpret_cast<const float2*>(&sB[(nb * (nb + 1) / 2 + kb) * 64 + lane * 2]);
                        asm volatile(
                            "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                            "{%0,%1,%2,%3};"
                            : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                            : "r"(afr[kb][0]), "r"(afr[kb][1]), "r"(afr[kb][2]), "r"(afr[kb][3]),
                              "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)));
                    }
                    dfr[mt][nb][0] = c0;
                    dfr[mt][nb][1] = c1;
                    dfr[mt][nb][2] = c2;
                    dfr[mt][nb][3] = c3;
                }
                __syncwarp();                                             // fragments read: the buffer is free
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    *reinterpret_cast<float2*>(&zb[fg * kZLd + nb * 8 + 2 * ft]) = make_float2(dfr[mt][nb][0], dfr[mt][nb][1]);
                    *reinterpret_cast<float2*>(&zb[(fg + 8) * kZLd + nb * 8 + 2 * ft]) =
                        make_float2(dfr[mt][nb][2], dfr[mt][nb][3]);
                }
                __syncwarp();
                if ((lane >> 4) == mt) {
#pragma unroll
                    for (int c = 0; c < DMAX; c += 4) {
                        const float4 v = *reinterpret_cast<const float4*>(&zb[(lane & 15) * kZLd + c]);
                        delta[c + 0] = v.x;
                        delta[c + 1] = v.y;
                        delta[c + 2] = v.z;
                        delta[c + 3] = v.w;
                    }
                }
                __syncwarp();
            }
            }  // !STRICT
            mbar_wait(&mbar, tma_phase);
            tma_phase ^= 1u;
#pragma unroll
            for (int c = 0; c < DMAX; ++c)
                th[0][c] = stage[c * TPB + threadIdx.x] + (STRICT ? 0.0 : (double)delta[STRICT ? 0 : c]);
            ll_old[0] = stage[d * TPB + threadIdx.x];
            lp_old[0] = stage[(d + 1) * TPB + threadIdx.x];
            // every thread reports that it has drained the stage; only the issuing thread waits for
            // the whole block (a block-wide barrier here cost 11 % of the stall samples)
            mbar_arrive(&mbar_empty);
            if (threadIdx.x == 0) {
                mbar_wait(&mbar_empty, tma_phase ^ 1u);
                if (tile + gridDim.x < n_tiles) tma_issue(tile + gridDim.x);
            }
            if constexpr (STRICT) {
                // fp64 accumulation of sqrt(s) L z, column-outer, on the same normals the tf32 variant draws
#pragma unroll
                for (int c = 0; c + 15 < DMAX; c += 16) {
                    const Float16 zz = gen_normals16(a.rk.k[0], a.rk.k[1], pid, (uint32_t)(c >> 2), (uint32_t)a.rng.stream);
                    const float zq[16] = {zz.a.x, zz.a.y, zz.a.z, zz.a.w, zz.b.x, zz.b.y, zz.b.z, zz.b.w,
                                          zz.c.x, zz.c.y, zz.c.z, zz.c.w, zz.d.x, zz.d.y, zz.d.z, zz.d.w};
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int j = c + k;
                        const double zs = sc * (double)zq[k];
#pragma unroll
                        for (int r2 = (j & ~1); r2 < DMAX; r2 += 2) {       // above the diagonal: 0
                            const double2 l2 = *reinterpret_cast<const double2*>(&sLt[j * DMAX + r2]);
                            th[0][r2] = fma(l2.x, zs, th[0][r2]);
                            if (r2 + 1 < DMAX) th[0][r2 + 1] = fma(l2.y, zs, th[0][r2 + 1]);
                        }
                    }
                }
            }
        } else if constexpr (PREF) {
            const int64_t i = tile * per_tile + threadIdx.x;
            valid[0] = i < a.n;
            idx[0] = valid[0] ? i : 0;
            cp_async_wait_all();                               // this thread's prefetched values have landed
#pragma unroll
            for (int c = 0; c < DMAX; ++c) th[0][c] = (EXACT || c < d) ? pstage[c * TPB + threadIdx.x] : 0.0;
            ll_old[0] = pstage[DMAX * TPB + threadIdx.x];
            lp_old[0] = pstage[(DMAX + 1) * TPB + threadIdx.x];
            if (tile + gridDim.x < n_tiles) pref_issue(tile + gridDim.x);     // slots are free: values are in registers
        } else {
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const int64_t i = tile * per_tile + (int64_t)p * TPB + threadIdx.x;
                valid[p] = i < a.n;
                idx[p] = valid[p] ? i : 0;
                const double* pc = a.params + idx[p];
#pragma unroll
                for (int c = 0; c < DMAX; ++c) {
                    th[p][c] = (EXACT || c < d) ? *pc : 0.0;
                    pc += a.stride;
                }
                ll_old[p] = lp_old[p] = 0.0;
                if (MODE == 1 && !kLateOld) {
                    ll_old[p] = a.ll[idx[p]];
                    lp_old[p] = a.lp[idx[p]];
                }
            }
        }
#pragma unroll
        for (int p = 0; p < PPT; ++p) lq_old[p] = (PROP && MODE == 1) ? a.lq[idx[p]] : 0.0;

        if ((MODE == 1 || MODE == 2) && !TMA) {
            // proposal x' = x + sqrt(s) L z, column-outer: z_j updates rows j..d-1, which are
            // independent accumulators (the parameter registers themselves)
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const uint64_t pid = (uint64_t)(a.rng.id_offset + idx[p]);
                const double* zt = (a.rng.mode == 1) ? a.rng.z + idx[p] * d : nullptr;
#pragma unroll
                for (int c = 0; c < DMAX; c += 4) {
                    if (EXACT || c < d) {
                        double z4[4];
                        if (a.rng.mode == 1) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) z4[q] = (c + q < d) ? zt[c + q] : 0.0;
                        } else {
                            const uint4 r = philox_k(a.rk, pid, (uint32_t)(c >> 2), (uint32_t)a.rng.stream);
                            normal_pair(r.x, r.y, z4[0], z4[1]);
                            normal_pair(r.z, r.w, z4[2], z4[3]);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int j = c + q;
                            if (j < DMAX) {
                                const double zs = sc * z4[q];
                                if (DMAX == 1) {
                                    th[p][0] = fma(sLt[0], zs, th[p][0]);
                                } else {
#pragma unroll
                                    for (int r2 = (j & ~1); r2 < DMAX; r2 += 2) {   // above the diagonal: 0
                                        const double2 l2 = *reinterpret_cast<const double2*>(&sLt[j * DMAX + r2]);
                                        th[p][r2] = fma(l2.x, zs, th[p][r2]);
                                        if (r2 + 1 < DMAX) th[p][r2 + 1] = fma(l2.y, zs, th[p][r2 + 1]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }

#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            double* lp_cols_row =
                ((MODE == 0 || MODE == 3) && a.lp_cols && valid[p]) ? a.lp_cols + idx[p] : nullptr;
            lp_new[p] = eval_priors<KIND, DMAX, EXACT>(a, th[p], lp_cols_row, a.stride);
            prior_ok[p] = lp_new[p] != kNegInf;
            lq_new[p] = PROP ? eval_proposal<DMAX, EXACT>(a, th[p]) : 0.0;
        }

        if (MODE == 3) {
#pragma unroll
            for (int p = 0; p < PPT; ++p)
                if (valid[p]) a.lp[idx[p]] = lp_new[p];
            continue;
        }
        if (MODE == 2) {
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                if (valid[p]) {
                    double* pc = a.new_params + idx[p];
#pragma unroll
                    for (int c = 0; c < DMAX; ++c) {
                        if (EXACT || c < d) *pc = th[p][c];
                        pc += a.stride;
                    }
                    a.new_lp[idx[p]] = lp_new[p];
                }
            }
            continue;
        }

        // proposals with a -inf prior are not evaluated (quirk Q6, vector_mcmc.py:205-214);
        // incoming particles (MODE 0) always are (vector_mcmc.py:116-118 has no prior test)
        bool run[PPT], not_pd[PPT];
        bool run_any = false;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            run[p] = valid[p] && (MODE == 0 || prior_ok[p]);
            not_pd[p] = false;
            run_any = run_any || run[p];
        }
        if (MULTI) {
#pragma unroll
            for (int p = 0; p < PPT; ++p) ll_new[p] = ll_multisource<KIND, DMAX>(a, th[p], run[p], sdata);
        } else if (KIND == K_LINEAR_NORMAL) {
            ll_linear_normal<DMAX, PPT>(a, th, run_any, resident, sdata, ll_new);
        } else if (KIND == K_SPRING_NORMAL) {
            ll_spring_normal<DMAX, PPT>(a, th, run_any, resident, sdata, ll_new);
        } else {
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                if (KIND == K_IDENTITY_NORMAL) ll_new[p] = ll_identity_normal<DMAX, EXACT>(a, th[p], run[p], sdata);
                else ll_new[p] = ll_linear_mvn<KIND, DMAX>(a, th[p], run[p], sdata, not_pd[p]);
            }
        }
        if constexpr (kLateOld) {
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                ll_old[p] = a.ll[idx[p]];
                lp_old[p] = a.lp[idx[p]];
            }
        }
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            if (!run[p]) ll_new[p] = 0.0;       // placeholder (vector_mcmc.py:207)
            // a particle outside the prior support may legitimately have no valid likelihood
            if (run[p] && prior_ok[p] && ll_new[p] != ll_new[p]) flag |= SMCB_FLAG_MODEL_NAN;
            if (run[p] && prior_ok[p] && not_pd[p]) flag |= SMCB_FLAG_COV_NOT_PD;
            if (MODE == 0) {
                if (valid[p]) {
                    a.ll[idx[p]] = ll_new[p];
                    a.lp[idx[p]] = lp_new[p];
                    if (PROP) a.lq[idx[p]] = lq_new[p];
                    if (!prior_ok[p]) flag |= SMCB_FLAG_PRIOR_OOB;
                }
                continue;
            }
            // acceptance (vector_mcmc.py:131-151 with evaluate_log_posterior = path.logpdf)
            if (valid[p]) {
                const int64_t i = idx[p];
                const double t_old = path_target(ll_old[p], lp_old[p], PROP ? lq_old[p] : lp_old[p], a.phi, a.pe, a.qe);
                const double t_new = path_target(ll_new[p], lp_new[p], PROP ? lq_new[p] : lp_new[p], a.phi, a.pe, a.qe);
                const double ratio = exp(t_new - t_old);
                double u;
                if (a.rng.mode == 1) {
                    u = a.rng.u[i];
                } else {
                    u = gen_uniform53(a.rk.k[0], a.rk.k[1], (uint64_t)(a.rng.id_offset + i), 0xFFFFu, (uint32_t)a.rng.stream);
                }
                const bool rejected = ratio < u;     // NaN compares false -> accept, as numpy does
                if (!rejected && !(run[p] && not_pd[p])) {
                    double* pc = a.params + i;
#pragma unroll
                    for (int c = 0; c < DMAX; ++c) {
                        if (EXACT || c < d) *pc = th[p][c];
                        pc += a.stride;
                    }
                    a.ll[i] = ll_new[p];
                    a.lp[i] = lp_new[p];
                    if (PROP) a.lq[i] = lq_new[p];
                    a.moved[i] = 1;
                    ++n_acc;
                }
            }
        }
    }
    if (flag) atomicOr(a.flags, flag);
    if (DYN) {
        // the last block to run out of tiles resets the tile counter and the ticket
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int* ctr = reinterpret_cast<unsigned int*>(a.flags);
            __threadfence();
            if (atomicAdd(ctr + 2, 1u) == gridDim.x - 1) {
                ctr[1] = 0u;
                ctr[2] = 0u;
            }
        }
    }
    if (MODE != 1) return;
    // accept count: warp shuffle reduction, one atomic per warp per launch
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
    if ((threadIdx.x & 31) == 0 && n_acc) atomicAdd(a.accept_counts + a.step, (unsigned long long)n_acc);
    if (a.x.world > 1 && !DYN) {                    // (dynamic-tile kernels: xchg_tile_done stores the final slot)
        // the last block of the grid publishes this rank's count to every peer (P2P stores)
        __shared__ bool is_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) is_last = (atomicAdd(a.x.ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (is_last && threadIdx.x < a.x.world) {
            __threadfence();
            const unsigned long long local = atomicAdd(a.accept_counts + a.step, 0ull);
            const unsigned long long word = (a.x.gen << 32) | (local & 0xffffffffull);
            st_sys(reinterpret_cast<unsigned long long*>(a.x.peer_slots[threadIdx.x]) +
                       (size_t)a.step * SMCB_MAX_RANKS + a.x.rank,
                   word);
            __threadfence_system();
            if (threadIdx.x == 0) a.x.ticket[0] = 0u;
        }
    }
}

// ---- un-fused pieces (torch-callable user models) ---------------------------------------------------
// row-wise Normal log-likelihood of model outputs (n, m) row-major: one warp per row
__global__ void __launch_bounds__(256) normal_ll_rows_kernel(const double* __restrict__ out, int64_t n, int m,
                                                             const double* __restrict__ data,
                                                             const double* sigma_col, double sigma,
                                                             const double* lp, double* ll, int* flags) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += n_warps) {
        if (lp && lp[r] == kNegInf) {
            if (lane == 0) ll[r] = 0.0;
            continue;
        }
        const double* row = out + r * m;
        double s = 0.0;
        for (int j = lane; j < m; j += 32) {
            const double dd = row[j] - data[j];
            s = fma(dd, dd, s);
        }
        s = warp_sum(s);
        if (lane == 0) {
            const double v = normal_ll(s, sigma_col ? sigma_col[r] : sigma, m);
            ll[r] = v;
            if (v != v) atomicOr(flags, SMCB_FLAG_MODEL_NAN);
        }
    }
}

struct AcceptArgs {
    double* params;
    const double* new_params;
    int64_t stride, n;
    int32_t d;
    double* ll;
    const double* new_ll;
    double* lp;
    const double* new_lp;
    uint8_t* moved;
    unsigned long long* accept_counts;
    int32_t step;
    double phi, pe, qe;
    smcb_rng_t rng;
};

__global__ void __launch_bounds__(256) mh_accept_kernel(const AcceptArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool accepted = false;
    if (i < a.n) {
        const double t_old = path_target(a.ll[i], a.lp[i], a.lp[i], a.phi, a.pe, a.qe);
        const double t_new = path_target(a.new_ll[i], a.new_lp[i], a.new_lp[i], a.phi, a.pe, a.qe);
        const double ratio = exp(t_new - t_old);
        double u;
        if (a.rng.mode == 1) {
            u = a.rng.u[i];
        } else {
            const Philox rng(a.rng.seed);
            const uint4 r = rng((uint64_t)(a.rng.id_offset + i), 0xFFFFu, (uint32_t)a.rng.stream);
            u = u01_53(r.x, r.y);
        }
        accepted = !(ratio < u);
        if (accepted) {
            for (int c = 0; c < a.d; ++c)
                a.params[(int64_t)c * a.stride + i] = a.new_params[(int64_t)c * a.stride + i];
            a.ll[i] = a.new_ll[i];
            a.lp[i] = a.new_lp[i];
            a.moved[i] = 1;
        }
    }
    const int n_acc = __syncthreads_count(accepted ? 1 : 0);
    if (threadIdx.x == 0 && n_acc) atomicAdd(a.accept_counts + a.step, (unsigned long long)n_acc);
}

// initial particles from uniform priors: x = loc + scale * U  (scipy uniform.rvs)
// two standard normals in full fp64 (Box-Muller on 53-bit uniforms): prior draws, not proposals
__device__ __forceinline__ void normal_pair53(const uint4 r, double& z0, double& z1) {
    const double u1 = 1.0 - u01_53(r.x, r.y);              // (0, 1]
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u01_53(r.z, r.w), &sn, &cs);
    z0 = rad * cs;
    z1 = rad * sn;
}

// Initial particles from the priors (vector_mcmc.py:91-96 sample_from_priors on the device): scipy
// uniform(loc, scale).rvs = loc + scale * U; ImproperCov.rvs = inverse-Wishart(dof, S) draws
// (smcpy/priors.py:118-123) by the Bartlett decomposition -- A lower triangular with
// A_ii = sqrt(chi2(dof - i)) (integer dof: a sum of squared normals), A_ij ~ N(0, 1), W = A A^T ~
// Wishart(dof, I), X = C W^-1 C^T with S = C C^T -- written as packed upper-triangular row-major columns.
struct IwArgs {
    int n_cov;
    int c0[2], ncols[2], dof[2];
    double chol[2][6];
};
__global__ void __launch_bounds__(256) sample_priors_kernel(ColPriors cp, IwArgs iw, int d, double* params, int64_t stride,
                                                            int64_t n, uint64_t seed, uint32_t stream_id,
                                                            int64_t id_offset) {
    const Philox rng(seed);
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
        const uint64_t pid = (uint64_t)(id_offset + i);
        for (int c = 0; c < d; c += 2) {
            const uint4 r = rng(pid, (uint32_t)(c >> 1), stream_id);
            if (cp.prior_of_col[c] >= 0) params[(int64_t)c * stride + i] = cp.col[c].a + cp.col[c].hi * u01_53(r.x, r.y);
            if (c + 1 < d && cp.prior_of_col[c + 1] >= 0)
                params[(int64_t)(c + 1) * stride + i] = cp.col[c + 1].a + cp.col[c + 1].hi * u01_53(r.z, r.w);
        }
        for (int q = 0; q < iw.n_cov; ++q) {
            const int p = iw.ncols[q];
            uint32_t call = 64u + 64u * (uint32_t)q;
            double spare = 0.0;
            bool have = false;
            auto next_normal = [&]() {
                if (have) { have = false; return spare; }
                double z0;
                normal_pair53(rng(pid, call++, stream_id), z0, spare);
                have = true;
                return z0;
            };
            double A[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
            for (int r = 0; r < p; ++r) {
                for (int c = 0; c < r; ++c) A[r][c] = next_normal();
                double chi = 0.0;
                for (int k = 0; k < iw.dof[q] - r; ++k) { const double z = next_normal(); chi = fma(z, z, chi); }
                A[r][r] = sqrt(chi);
            }
            // B = A^-1 (lower), M = C B^T ... X = C (A A^T)^-1 C^T = (C A^-T)(C A^-T)^T; A^-T = B^T upper
            const double b00 = 1.0 / A[0][0], b11 = 1.0 / A[1][1], b22 = 1.0 / A[2][2];
            const double b10 = -A[1][0] * b00 * b11;
            const double b21 = -A[2][1] * b11 * b22;
            const double b20 = -(A[2][0] * b00 + A[2][1] * b10) * b22;
            const double B[3][3] = {{b00, 0, 0}, {b10, b11, 0}, {b20, b21, b22}};
            const double* L = iw.chol[q];
            const double C[3][3] = {{L[0], 0, 0}, {L[1], L[2], 0}, {L[3], L[4], L[5]}};
            double M[3][3];                                   // M = C B^T  (M[r][c] = sum_k C[r][k] B[c][k])
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    double v = 0.0;
                    for (int k = 0; k < 3; ++k) v = fma(C[r][k], B[c][k], v);
                    M[r][c] = v;
                }
            int e = 0;
            for (int r = 0; r < p; ++r)
                for (int c = r; c < p; ++c) {
                    double v = 0.0;
                    for (int k = 0; k < 3; ++k) v = fma(M[r][k], M[c][k], v);
                    params[(int64_t)(iw.c0[q] + e) * stride + i] = v;
                    ++e;
                }
        }
    }
}

// centred data of the linear model + Normal likelihood: out[0] = xbar, out[1] = ybar (fixed-order tree sums),
// out[2 + 2 j] = x_j - xbar, out[3 + 2 j] = -(y_j - ybar)
__global__ void __launch_bounds__(256) linear_center_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                            int m, double* out) {
    __shared__ double sx[256], sy[256];
    double ax = 0.0, ay = 0.0;
    for (int j = threadIdx.x; j < m; j += 256) {
        ax += x[j];
        ay += y[j];
    }
    sx[threadIdx.x] = ax;
    sy[threadIdx.x] = ay;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sx[threadIdx.x] += sx[threadIdx.x + o];
            sy[threadIdx.x] += sy[threadIdx.x + o];
        }
        __syncthreads();
    }
    const double xbar = sx[0] / (double)m, ybar = sy[0] / (double)m;
    if (threadIdx.x == 0) {
        out[0] = xbar;
        out[1] = ybar;
    }
    for (int j = threadIdx.x; j < m; j += 256) {
        out[2 + 2 * j] = x[j] - xbar;
        out[3 + 2 * j] = -(y[j] - ybar);
    }
}

// snapshot mean and scatter matrix of the MVNormal data: one thread per output entry, sequential sums
// over the snapshots (deterministic; S is small)
__global__ void mvn_scatter_kernel(const double* __restrict__ data, int S, int F, double* out) {
    __shared__ double mean[3];
    const int t = threadIdx.x;
    if (t < 3) {
        double m = 0.0;
        if (t < F) {
            for (int s = 0; s < S; ++s) m += data[s * F + t];
            m /= (double)S;
        }
        mean[t] = m;
        out[t] = m;
    }
    __syncthreads();
    if (t < 9) {
        const int j = t / 3, k = t - 3 * j;
        double w = 0.0;
        if (j < F && k < F)
            for (int s = 0; s < S; ++s) w = fma(data[s * F + j] - mean[j], data[s * F + k] - mean[k], w);
        out[3 + t] = w;
    }
}

// ---- host side ------------------------------------------------------------------------------------------
static int build_col_priors(const smcb_problem_t* p, ColPriors* cp) {
    memset(cp, 0, sizeof(*cp));
    const double inf = __builtin_huge_val();
    for (int c = 0; c < SMCB_MAX_PARAMS; ++c) {
        cp->prior_of_col[c] = -1;
        cp->col[c].lo = -inf;
        cp->col[c].hi = inf;
    }
    int col = 0;
    double total = 0.0;
    for (int q = 0; q < p->n_priors; ++q) {
        const smcb_prior_t& pr = p->priors[q];
        if (pr.kind == SMCB_PRIOR_UNIFORM || pr.kind == SMCB_PRIOR_IMPROPER_UNIFORM) {
            if (col >= SMCB_MAX_PARAMS) return -1;
            cp->prior_of_col[col] = (int8_t)q;
            if (pr.kind == SMCB_PRIOR_UNIFORM) {
                cp->col[col].a = pr.a;
                cp->col[col].lo = 0.0;
                cp->col[col].hi = pr.b;
            } else {
                cp->col[col].a = 0.0;
                cp->col[col].lo = pr.a;
                cp->col[col].hi = pr.b;
            }
            cp->col[col].logp = pr.logp_in;
            total += pr.logp_in;
            col += 1;
        } else if (pr.kind == SMCB_PRIOR_IMPROPER_COV) {
            if (cp->n_cov >= 2 || pr.ncols < 1 || pr.ncols > 3) return -1;
            cp->cov_c0[cp->n_cov] = col;
            cp->cov_ncols[cp->n_cov] = pr.ncols;
            cp->cov_prior[cp->n_cov] = q;
            cp->n_cov += 1;
            col += pr.ncols * (pr.ncols + 1) / 2;
            if (col > SMCB_MAX_PARAMS) return -1;
        } else if (pr.kind == SMCB_PRIOR_EXTERNAL) {
            // host-evaluated prior: flat (0) on the device over its ncols columns
            if (pr.ncols < 1 || col + pr.ncols > SMCB_MAX_PARAMS) return -1;
            for (int k = 0; k < pr.ncols; ++k) {
                cp->prior_of_col[col + k] = (int8_t)q;
                cp->col[col + k].a = 0.0;
            }
            col += pr.ncols;
            cp->n_external += 1;
        } else {
            return -1;
        }
    }
    cp->logp_total = total;
    return col;
}

static size_t mh_smem_bytes(const smcb_problem_t* p) {
    const int d = p->n_params;
    size_t data = 0;
    if (p->loglike_id == SMCB_LL_MVNORMAL) data = (size_t)(kMvnStats + 3 + 1) * 8;
    else if (p->model_id == SMCB_MODEL_LINEAR) data = (size_t)kChunk * 16;
    else if (p->model_id == SMCB_MODEL_SPRING_MASS) data = (size_t)kChunk * 8;
    else data = (size_t)(SMCB_MAX_PARAMS + 2) * 8;
    (void)d;
    return data + 16;      // + DMAX*DMAX doubles for the Cholesky factor, added per instantiation
}

// set by smcb_mh_steps around the launches of steps k >= 1 (the calling thread's launches only)
static thread_local bool tl_pdl_launch = false;

template <int KIND, int DMAX, int MODE, int TPB, int PPT, bool EXACT, bool TMA = false, bool PROP = false,
          bool DYN = false, bool MULTI = false, bool STRICT = false, bool PREF = false>
static int launch_cfg2(const MhArgs& a, cudaStream_t s) {
    static_assert(!PREF || (PPT == 1 && MODE == 1 && !TMA && !DYN && !PROP), "PREF: persistent static tiles, one particle per thread");
    const size_t smem = mh_smem_bytes(&a.prob) + (size_t)DMAX * DMAX * sizeof(double) +
                        (TMA ? (size_t)(64 + (DMAX + 2) * TPB) * sizeof(double) +
                                   (size_t)(TPB / 32) * 16 * (DMAX + 4) * sizeof(float) : 0) +
                        (PREF ? (size_t)(64 + (DMAX + 2) * TPB) * sizeof(double) : 0);
    if (TMA) {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(mh_kernel<KIND, DMAX, MODE, TPB, PPT, EXACT, TMA, PROP, DYN, MULTI, STRICT, PREF>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            attr = true;
        }
    }
    // persistent grid: resident blocks per SM (occupancy query, cached) x SM count
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mh_kernel<KIND, DMAX, MODE, TPB, PPT, EXACT, TMA, PROP, DYN, MULTI, STRICT, PREF>, TPB,
                                                          smem) != cudaSuccess || nb < 1)
            nb = 1;
        blocks_per_sm = nb;
    }
    const int64_t per_tile = (int64_t)TPB * PPT;
    const int64_t n_tiles = (a.n + per_tile - 1) / per_tile;
    const int persist_env = options().persist;
    // persistent blocks (stage L / data once, loop over tiles) for the kernels with little work per particle
    // (identity model, MVNormal through its scatter matrix: C4 0.188 ms persistent against 0.277 ms with one
    // 128-particle tile per block, and on several GPUs every block also ends with a fence + ticket); one tile
    // per block for the FP64-bound models with long tiles, where the hardware block scheduler balances the
    // tail better than a static tile loop (C3: 10.6 ms against 11.2 ms).  SMCB_PERSIST overrides: tuning aid
    const bool persist = PREF || DYN || (persist_env >= 0 ? (persist_env != 0) : (KIND == K_IDENTITY_NORMAL || is_mvn(KIND)));
    const int64_t cap = persist ? (int64_t)kNumSMs * blocks_per_sm : n_tiles;
    const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
    if (MODE == 1) {
        char txt[200];
        snprintf(txt, sizeof(txt),
                 "mh_kernel<kind=%d,dmax=%d,tpb=%d,ppt=%d,exact=%d,tma=%d,prop=%d,dyn=%d,multi=%d,strict=%d,pref=%d> grid=%u "
                 "rng=%d n=%lld%s",
                 KIND, DMAX, TPB, PPT, (int)EXACT, (int)TMA, (int)PROP, (int)DYN, (int)MULTI, (int)STRICT, (int)PREF, grid,
                 (int)a.rng.mode, (long long)a.n,
                 KIND == K_SPRING_NORMAL ? (a.spring_form == 1 && !MULTI ? " rk4=propagator" : " rk4=stages") : "");
        note_mh_variant(txt);
    }
    if (MODE == 1 && tl_pdl_launch) {
        // step k >= 1 of smcb_mh_steps: programmatic stream serialisation against the previous step's launch
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(TPB);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const cudaError_t e =
            cudaLaunchKernelEx(&cfg, mh_kernel<KIND, DMAX, MODE, TPB, PPT, EXACT, TMA, PROP, DYN, MULTI, STRICT, PREF>, a);
        if (e != cudaSuccess) {
            set_error("mh_kernel (dependent launch): %s", cudaGetErrorString(e));
            return SMCB_ERR_CUDA;
        }
        return SMCB_OK;
    }
    mh_kernel<KIND, DMAX, MODE, TPB, PPT, EXACT, TMA, PROP, DYN, MULTI, STRICT, PREF><<<grid, TPB, smem, s>>>(a);
    return check_launch("mh_kernel");
}

template <int KIND, int DMAX, int MODE, int TPB, int PPT>
static int launch_cfg(const MhArgs& a, cudaStream_t s) {
    // exact-width instantiations (no per-column bounds checks) where they pay
    constexpr bool kHasExact = (KIND == K_IDENTITY_NORMAL && DMAX >= 8) || (!is_mvn(KIND) && DMAX == 2) ||
                               (KIND == K_LINEAR_MVN3);
    if constexpr (kHasExact) {
        if (a.prob.n_params == DMAX) {
            // TMA-staged variant for the bandwidth-bound wide identity kernels: needs 16-byte
            // aligned column segments (even n and stride) and enough tiles to pipeline
            if constexpr (KIND == K_IDENTITY_NORMAL && MODE == 1 && DMAX >= 16 && PPT == 1) {
                const bool no_tma = options().no_tma != 0;                       // tuning / test aid
                const bool aligned = ((((uintptr_t)a.params | (uintptr_t)a.ll | (uintptr_t)a.lp) & 15) == 0) &&
                                     (a.stride % 2 == 0) && (a.n % 2 == 0);
                if (!no_tma && aligned && a.rng.mode == 0 && a.n >= (int64_t)kNumSMs * 8 * TPB) {
                    if (options().strict_proposal)
                        return launch_cfg2<KIND, DMAX, MODE, TPB, PPT, true, true, false, false, false, true>(a, s);
                    return launch_cfg2<KIND, DMAX, MODE, TPB, PPT, true, true>(a, s);
                }
            }
            if constexpr (KIND == K_LINEAR_MVN3 && MODE == 1 && PPT == 1) {
                // narrow bandwidth-bound kernel: per-thread cp.async prefetch of the next tile (option no_pref: off)
                if (!options().no_pref && a.n >= (int64_t)kNumSMs * 8 * TPB)
                    return launch_cfg2<KIND, DMAX, MODE, TPB, PPT, true, false, false, false, false, false, true>(a, s);
            }
            return launch_cfg2<KIND, DMAX, MODE, TPB, PPT, true>(a, s);
        }
    }
    return launch_cfg2<KIND, DMAX, MODE, TPB, PPT, false>(a, s);
}

template <int KIND, int DMAX, int MODE>
static int launch_one(const MhArgs& a, cudaStream_t s) {
    if constexpr (MODE == 0 || MODE == 1) {
        if constexpr (!is_mvn(KIND)) {
            if (a.prob.loglike_id == SMCB_LL_MULTISOURCE) {
                if (a.prob.has_proposal) {
                    set_error("MultiSourceNormal with a path proposal has no device kernel");
                    return SMCB_ERR_UNSUPPORTED;
                }
                return launch_cfg2<KIND, DMAX, MODE, kMhThreads, 1, false, false, false, false, true>(a, s);
            }
        }
        // paths with a proposal: generic 1-particle-per-thread variant with the log q column
        if (a.prob.has_proposal) return launch_cfg2<KIND, DMAX, MODE, kMhThreads, 1, false, false, true>(a, s);
    }
    if constexpr (KIND == K_LINEAR_NORMAL && (MODE == 0 || MODE == 1)) {
        // several particles per thread amortise the shared-memory data reads; keep at least
        // ~4 tiles per SM so small problems still fill the chip
        const int cfg = options().linear_cfg;                                                      // tuning / test aid
        if (cfg == 0) return launch_cfg<KIND, DMAX, MODE, kMhThreads, 1>(a, s);
        if (cfg == 1) return launch_cfg<KIND, DMAX, MODE, 128, 2>(a, s);
        if (cfg == 2) return launch_cfg<KIND, DMAX, MODE, 128, 4>(a, s);
        if (cfg == 3) return launch_cfg<KIND, DMAX, MODE, 256, 4>(a, s);
        if constexpr (MODE == 1 && DMAX == 2) {
            // persistent blocks + dynamic 512-particle tiles (128 threads x 4 particles): measured
            // 0.2198 ms vs 0.2385 ms for static 1024-particle blocks (C2, same box)
            const bool dyn_ok = a.prob.n_params == 2 && a.n >= (int64_t)kNumSMs * 16 * 256;
            if (cfg == 5) return launch_cfg2<KIND, DMAX, MODE, 64, 4, true, false, false, true>(a, s);
            if (cfg == 6 || (cfg < 0 && dyn_ok))
                return launch_cfg2<KIND, DMAX, MODE, 128, 4, true, false, false, true>(a, s);
        }
        if (a.n >= (int64_t)kNumSMs * 4 * 1024) return launch_cfg<KIND, DMAX, MODE, 256, 4>(a, s);
        if (a.n >= (int64_t)kNumSMs * 4 * 512) return launch_cfg<KIND, DMAX, MODE, 128, 4>(a, s);
        if (a.n >= (int64_t)kNumSMs * 4 * 256) return launch_cfg<KIND, DMAX, MODE, 128, 2>(a, s);
    }
    if constexpr (KIND == K_SPRING_NORMAL && (MODE == 0 || MODE == 1)) {
        // two interleaved RK4 chains per thread once there are enough tiles to fill the chip
        const int ppt = options().spring_ppt;                                                      // tuning / test aid
        if (ppt == 2 || (ppt < 0 && a.n >= (int64_t)kNumSMs * 8 * 256)) return launch_cfg<KIND, DMAX, MODE, 128, 2>(a, s);
    }
    return launch_cfg<KIND, DMAX, MODE, kMhThreads, 1>(a, s);
}

template <int MODE>
static int dispatch(const MhArgs& a, cudaStream_t s) {
    const smcb_problem_t& p = a.prob;
    const int d = p.n_params;
#define SMCB_CASE(KIND, DMAX) return launch_one<KIND, DMAX, MODE>(a, s)
    if constexpr (MODE == 3) {           // priors only: matrix priors need the MVN instantiation
        if (a.cp.n_cov > 0) {
            if (d <= 4) SMCB_CASE(K_LINEAR_MVN, 4);
            if (d <= 8) SMCB_CASE(K_LINEAR_MVN, 8);
        } else {
            if (d <= 2) SMCB_CASE(K_IDENTITY_NORMAL, 2);
            if (d <= 4) SMCB_CASE(K_IDENTITY_NORMAL, 4);
            if (d <= 8) SMCB_CASE(K_IDENTITY_NORMAL, 8);
            if (d <= 16) SMCB_CASE(K_IDENTITY_NORMAL, 16);
            if (d <= 32) SMCB_CASE(K_IDENTITY_NORMAL, 32);
        }
    } else if constexpr (MODE == 2) {    // the proposal does not depend on the model
        if (d <= 2) SMCB_CASE(K_IDENTITY_NORMAL, 2);
        if (d <= 4) SMCB_CASE(K_IDENTITY_NORMAL, 4);
        if (d <= 8) SMCB_CASE(K_IDENTITY_NORMAL, 8);
        if (d <= 16) SMCB_CASE(K_IDENTITY_NORMAL, 16);
        if (d <= 32) SMCB_CASE(K_IDENTITY_NORMAL, 32);
    } else if (p.loglike_id == SMCB_LL_MULTISOURCE && p.model_id == SMCB_MODEL_LINEAR) {
        if (d <= 4) SMCB_CASE(K_LINEAR_NORMAL, 4);
        if (d <= 8) SMCB_CASE(K_LINEAR_NORMAL, 8);
    } else if (p.loglike_id == SMCB_LL_MULTISOURCE && p.model_id == SMCB_MODEL_SPRING_MASS) {
        if (d <= 4) SMCB_CASE(K_SPRING_NORMAL, 4);
        if (d <= 8) SMCB_CASE(K_SPRING_NORMAL, 8);
    } else if (p.loglike_id == SMCB_LL_NORMAL && p.model_id == SMCB_MODEL_LINEAR) {
        if (d == 2) SMCB_CASE(K_LINEAR_NORMAL, 2);
        if (d == 3) SMCB_CASE(K_LINEAR_NORMAL, 4);
    } else if (p.loglike_id == SMCB_LL_NORMAL && p.model_id == SMCB_MODEL_SPRING_MASS) {
        if (d == 2) SMCB_CASE(K_SPRING_NORMAL, 2);
        if (d == 3) SMCB_CASE(K_SPRING_NORMAL, 4);
    } else if ((p.loglike_id == SMCB_LL_NORMAL || p.loglike_id == SMCB_LL_MULTISOURCE) &&
               p.model_id == SMCB_MODEL_IDENTITY) {
        if (d <= 2) SMCB_CASE(K_IDENTITY_NORMAL, 2);
        if (d <= 4) SMCB_CASE(K_IDENTITY_NORMAL, 4);
        if (d <= 8) SMCB_CASE(K_IDENTITY_NORMAL, 8);
        if (d <= 16) SMCB_CASE(K_IDENTITY_NORMAL, 16);
        if (d <= 32) SMCB_CASE(K_IDENTITY_NORMAL, 32);
    } else if (p.loglike_id == SMCB_LL_MVNORMAL && p.model_id == SMCB_MODEL_LINEAR) {
        if constexpr (MODE == 0 || MODE == 1) {
            bool std3 = d == 8 && p.n_data == 3 && p.n_model_params == 2 && !p.has_proposal && !options().no_mvn3 &&
                        (a.cp.n_cov == 0 || (a.cp.n_cov == 1 && a.cp.cov_c0[0] == 2 && a.cp.cov_ncols[0] == 3));
            for (int e = 0; e < 6; ++e) std3 = std3 && p.cov_param_col[e] == 2 + e;
            if (std3) SMCB_CASE(K_LINEAR_MVN3, 8);
        }
        if (d <= 4) SMCB_CASE(K_LINEAR_MVN, 4);
        if (d <= 8) SMCB_CASE(K_LINEAR_MVN, 8);
    }
#undef SMCB_CASE
    set_error("no fused kernel for model %d / loglike %d / d=%d", p.model_id, p.loglike_id, d);
    return SMCB_ERR_UNSUPPORTED;
}

static int validate_problem(const smcb_problem_t* p, ColPriors* cp) {
    if (!p) { set_error("problem is NULL"); return SMCB_ERR_ARG; }
    if (p->n_params < 1 || p->n_params > SMCB_MAX_PARAMS) { set_error("n_params out of range"); return SMCB_ERR_ARG; }
    if (p->n_priors < 1 || p->n_priors > SMCB_MAX_PRIORS) { set_error("n_priors out of range"); return SMCB_ERR_ARG; }
    const int cols = build_col_priors(p, cp);
    if (cols != p->n_params) {
        set_error("Num prior distributions != num input params (%d prior columns, %d params)", cols, p->n_params);
        return SMCB_ERR_ARG;
    }
    if (!p->data) { set_error("data pointer is NULL"); return SMCB_ERR_ARG; }
    if (cp->n_external > 0) {
        set_error("host-evaluated (EXTERNAL) priors go through the un-fused route (smcb_mh_propose / smcb_mh_accept)");
        return SMCB_ERR_UNSUPPORTED;
    }
    if (p->has_proposal) {
        for (int c = 0; c < p->n_params; ++c) {
            const smcb_proposal_col_t& q = p->proposal[c];
            if ((q.kind != SMCB_PROPOSAL_NORMAL && q.kind != SMCB_PROPOSAL_UNIFORM) || !(q.b > 0.0)) {
                set_error("proposal column %d: kind must be normal/uniform with scale > 0", c);
                return SMCB_ERR_ARG;
            }
        }
    }
    if (p->loglike_id == SMCB_LL_NORMAL) {
        const int extra = p->sigma_is_param ? 1 : 0;
        if (p->n_model_params + extra != p->n_params) { set_error("Normal: n_model_params + sigma != n_params"); return SMCB_ERR_ARG; }
        if (p->model_id == SMCB_MODEL_LINEAR && (!p->xgrid || p->n_model_params != 2)) { set_error("linear model needs xgrid and 2 params"); return SMCB_ERR_ARG; }
        if (p->model_id == SMCB_MODEL_LINEAR && p->loglike_id == SMCB_LL_NORMAL && !p->lin_centered) { set_error("linear model + Normal: lin_centered is NULL (fill it with smcb_linear_center)"); return SMCB_ERR_ARG; }
        if (p->model_id == SMCB_MODEL_SPRING_MASS && (p->n_model_params != 2 || p->model_args[3] < 1)) { set_error("spring-mass model needs 2 params and nsub >= 1"); return SMCB_ERR_ARG; }
        if (p->model_id == SMCB_MODEL_IDENTITY && p->n_model_params != p->n_data) { set_error("identity model needs n_data == n_model_params"); return SMCB_ERR_ARG; }
        if (p->model_id == SMCB_MODEL_IDENTITY && p->sigma_is_param) { set_error("identity model: fused kernels need a fixed sigma"); return SMCB_ERR_UNSUPPORTED; }
        if (cp->n_cov > 0) { set_error("ImproperCov priors are fused for the MVNormal likelihood only"); return SMCB_ERR_UNSUPPORTED; }
    } else if (p->loglike_id == SMCB_LL_MULTISOURCE) {
        if (p->n_segments < 1 || p->n_segments > SMCB_MAX_SEGMENTS) { set_error("MultiSourceNormal: 1..8 segments"); return SMCB_ERR_ARG; }
        int total = 0, n_sampled = 0;
        for (int k = 0; k < p->n_segments; ++k) {
            if (p->seg_len[k] < 0) { set_error("negative segment length"); return SMCB_ERR_ARG; }
            total += p->seg_len[k];
            if (p->seg_sigma_col[k] >= 0) {
                ++n_sampled;
                if (p->seg_sigma_col[k] >= p->n_params) { set_error("segment sigma column out of range"); return SMCB_ERR_ARG; }
            }
        }
        if (total != p->n_data) { set_error("data segments in args[0] must sum to dim of data"); return SMCB_ERR_ARG; }
        if (p->n_model_params + n_sampled != p->n_params) { set_error("MultiSourceNormal: n_model_params + sampled sigmas != n_params"); return SMCB_ERR_ARG; }
        if (p->n_data > kChunk) { set_error("MultiSourceNormal: at most %d data points", kChunk); return SMCB_ERR_UNSUPPORTED; }
        if (p->model_id == SMCB_MODEL_LINEAR && (!p->xgrid || p->n_model_params != 2)) { set_error("linear model needs xgrid and 2 params"); return SMCB_ERR_ARG; }
        if (p->model_id == SMCB_MODEL_SPRING_MASS && (p->n_model_params != 2 || p->model_args[3] < 1)) { set_error("spring-mass model needs 2 params and nsub >= 1"); return SMCB_ERR_ARG; }
        if (p->model_id == SMCB_MODEL_IDENTITY && p->n_model_params != p->n_data) { set_error("identity model needs n_data == n_model_params"); return SMCB_ERR_ARG; }
        if (cp->n_cov > 0) { set_error("ImproperCov priors are fused for the MVNormal likelihood only"); return SMCB_ERR_UNSUPPORTED; }
    } else if (p->loglike_id == SMCB_LL_MVNORMAL) {
        if (p->model_id != SMCB_MODEL_LINEAR || !p->xgrid || p->n_model_params != 2) { set_error("MVNormal is built in for the linear model only"); return SMCB_ERR_ARG; }
        if (p->n_data < 1 || p->n_data > 3 || p->n_snapshots < 1) { set_error("MVNormal: 1..3 features"); return SMCB_ERR_ARG; }
        if (!p->mvn_stats) { set_error("MVNormal: mvn_stats is NULL (fill it with smcb_mvn_scatter)"); return SMCB_ERR_ARG; }
    } else {
        set_error("unknown loglike_id");
        return SMCB_ERR_ARG;
    }
    return SMCB_OK;
}

}  // namespace smcb

using namespace smcb;

extern "C" int smcb_linear_center(const double* x, const double* y, int32_t m, double* out, void* stream) {
    if (!x || !y || !out || m < 1) {
        set_error("smcb_linear_center: bad argument");
        return SMCB_ERR_ARG;
    }
    linear_center_kernel<<<1, 256, 0, as_stream(stream)>>>(x, y, m, out);
    return check_launch("linear_center_kernel");
}

extern "C" int smcb_mvn_scatter(const double* data, int32_t n_snapshots, int32_t n_features, double* stats_out,
                                void* stream) {
    if (!data || !stats_out || n_snapshots < 1 || n_features < 1 || n_features > 3) {
        set_error("smcb_mvn_scatter: bad argument (1..3 features)");
        return SMCB_ERR_ARG;
    }
    mvn_scatter_kernel<<<1, 32, 0, as_stream(stream)>>>(data, n_snapshots, n_features, stats_out);
    return check_launch("mvn_scatter_kernel");
}

extern "C" int smcb_mh_init(const smcb_problem_t* prob, const double* params, int64_t stride, int64_t n,
                            double* ll_out, double* lp_out, double* lp_cols_out, double* lq_out, int32_t* flags,
                            void* stream) {
    MhArgs a;
    memset(&a, 0, sizeof(a));
    int rc = validate_problem(prob, &a.cp);
    if (rc) return rc;
    SMCB_REQUIRE(params && ll_out && lp_out && flags && n > 0 && stride >= n, "bad argument");
    a.prob = *prob;
    a.params = const_cast<double*>(params);
    a.stride = stride;
    a.n = n;
    a.ll = ll_out;
    SMCB_REQUIRE(!prob->has_proposal || lq_out, "lq_out required when the problem has a proposal");
    a.lp = lp_out;
    a.lq = lq_out;
    a.lp_cols = lp_cols_out;
    a.flags = flags;
    a.spring_form = options().spring_form;
    return dispatch<0>(a, as_stream(stream));
}

static void fill_step_args(MhArgs& a, const smcb_problem_t* prob, const smcb_path_t* path, int64_t stride, int64_t n,
                           int64_t n_global, const double* chol, int64_t* accept_counts, int32_t step,
                           const smcb_rng_t* rng) {
    a.prob = *prob;
    a.stride = stride;
    a.n = n;
    a.n_global = (double)n_global;
    a.chol = chol;
    a.accept_counts = reinterpret_cast<unsigned long long*>(accept_counts);
    a.step = step;
    a.spring_form = options().spring_form;
    a.rng = *rng;
    a.rk = make_philox_keys(rng->seed);
    if (path) {
        const PathCoef c = make_path_coef(path->phi, path->phi_prev, path->lambda);
        a.phi = c.phi;
        a.pe = c.pe;
        a.qe = c.qe;
    }
}

extern "C" int smcb_mh_step(const smcb_problem_t* prob, const smcb_path_t* path, double* params, int64_t stride,
                            int64_t n, int64_t n_global, double* ll, double* lp, double* lq, uint8_t* moved,
                            const double* chol,
                            int64_t* accept_counts, int32_t step, const smcb_rng_t* rng, int32_t* flags,
                            void* stream) {
    MhArgs a;
    memset(&a, 0, sizeof(a));
    int rc = validate_problem(prob, &a.cp);
    if (rc) return rc;
    SMCB_REQUIRE(path && params && ll && lp && moved && chol && accept_counts && rng && flags, "NULL argument");
    SMCB_REQUIRE(n > 0 && stride >= n && n_global >= n, "bad sizes");
    SMCB_REQUIRE(step >= 0 && step < SMCB_MAX_MCMC_STEPS, "step out of range");
    SMCB_REQUIRE(rng->mode == 0 || (rng->z && rng->u), "replay mode needs z and u tapes");
    SMCB_REQUIRE(!prob->has_proposal || lq, "lq required when the problem has a proposal");
    fill_step_args(a, prob, path, stride, n, n_global, chol, accept_counts, step, rng);
    a.params = params;
    a.ll = ll;
    a.lp = lp;
    a.lq = lq;
    a.moved = moved;
    a.flags = flags;
    return dispatch<1>(a, as_stream(stream));
}

extern "C" int smcb_mh_xchg_slot_words(void) { return 2 * SMCB_MAX_MCMC_STEPS * SMCB_MAX_RANKS; }

extern "C" int smcb_mh_step_xchg(const smcb_problem_t* prob, const smcb_path_t* path, double* params,
                                 int64_t stride, int64_t n, int64_t n_global, double* ll, double* lp, double* lq,
                                 uint8_t* moved, const double* chol, int64_t* accept_counts, int32_t step,
                                 const smcb_rng_t* rng, int32_t* flags, const smcb_xchg_t* x, void* stream) {
    MhArgs a;
    memset(&a, 0, sizeof(a));
    int rc = validate_problem(prob, &a.cp);
    if (rc) return rc;
    SMCB_REQUIRE(path && params && ll && lp && moved && chol && accept_counts && rng && flags && x, "NULL argument");
    SMCB_REQUIRE(n > 0 && stride >= n && n_global >= n, "bad sizes");
    SMCB_REQUIRE(step >= 0 && step < SMCB_MAX_MCMC_STEPS, "step out of range");
    SMCB_REQUIRE(rng->mode == 0 || (rng->z && rng->u), "replay mode needs z and u tapes");
    SMCB_REQUIRE(x->world >= 1 && x->world <= SMCB_MAX_RANKS && x->rank >= 0 && x->rank < x->world, "bad ranks");
    SMCB_REQUIRE(x->world == 1 || (x->global_counts && x->ticket && x->gen > 0 && x->gen < (1ull << 32)),
                 "exchange buffers required");
    for (int r = 0; r < x->world && x->world > 1; ++r) SMCB_REQUIRE(x->peer_slots[r] != nullptr, "peer slot pointer is NULL");
    fill_step_args(a, prob, path, stride, n, n_global, chol, accept_counts, step, rng);
    SMCB_REQUIRE(!prob->has_proposal || lq, "lq required when the problem has a proposal");
    a.params = params;
    a.ll = ll;
    a.lp = lp;
    a.lq = lq;
    a.moved = moved;
    a.flags = flags;
    a.x = *x;
    {
        const int tmo_s = options().xchg_timeout_s;
        a.xchg_timeout_ns = (unsigned long long)(tmo_s > 0 ? tmo_s : 30) * 1000000000ull;
    }
    return dispatch<1>(a, as_stream(stream));
}

extern "C" int smcb_mh_steps(const smcb_problem_t* prob, const smcb_path_t* path, double* params, int64_t stride,
                             int64_t n, int64_t n_global, double* ll, double* lp, double* lq, uint8_t* moved,
                             const double* chol, int64_t* accept_counts, int32_t first_step, int32_t n_steps,
                             const smcb_rng_t* rng, int32_t* flags, const smcb_xchg_t* x, void* stream) {
    SMCB_REQUIRE(rng && n_steps >= 0 && first_step >= 0 && first_step + n_steps <= SMCB_MAX_MCMC_STEPS, "bad step range");
    SMCB_REQUIRE(rng->mode == 0, "smcb_mh_steps: native generator only (replay tapes are per step)");
    smcb_rng_t r = *rng;
    const bool pdl = !options().no_pdl;
    for (int32_t k = 0; k < n_steps; ++k) {
        r.stream = rng->stream + (uint64_t)k;
        // steps after the first follow a launch of the same kernel: chained by programmatic dependent launch
        // (the first step follows the resampling / covariance / Cholesky kernels, which know nothing of it)
        tl_pdl_launch = pdl && k > 0;
        const int rc = (x && x->world > 1)
                           ? smcb_mh_step_xchg(prob, path, params, stride, n, n_global, ll, lp, lq, moved, chol, accept_counts,
                                               first_step + k, &r, flags, x, stream)
                           : smcb_mh_step(prob, path, params, stride, n, n_global, ll, lp, lq, moved, chol, accept_counts,
                                          first_step + k, &r, flags, stream);
        tl_pdl_launch = false;
        if (rc) return rc;
    }
    return SMCB_OK;
}

extern "C" int smcb_mh_propose(const smcb_problem_t* prob, const double* params, int64_t stride, int64_t n,
                               int64_t n_global, const double* chol, const int64_t* accept_counts, int32_t step,
                               const smcb_rng_t* rng, double* new_params, double* new_lp, void* stream) {
    MhArgs a;
    memset(&a, 0, sizeof(a));
    SMCB_REQUIRE(prob && params && chol && accept_counts && rng && new_params && new_lp, "NULL argument");
    SMCB_REQUIRE(n > 0 && stride >= n && prob->n_params >= 1 && prob->n_params <= SMCB_MAX_PARAMS, "bad sizes");
    const int cols = build_col_priors(prob, &a.cp);
    if (cols != prob->n_params) {
        set_error("Num prior distributions != num input params");
        return SMCB_ERR_ARG;
    }
    fill_step_args(a, prob, nullptr, stride, n, n_global, chol, const_cast<int64_t*>(accept_counts), step, rng);
    a.params = const_cast<double*>(params);
    a.new_params = new_params;
    a.new_lp = new_lp;
    return dispatch<2>(a, as_stream(stream));
}

extern "C" int smcb_log_priors(const smcb_problem_t* prob, const double* params, int64_t stride, int64_t n,
                               double* lp_out, double* lp_cols_out, void* stream) {
    MhArgs a;
    memset(&a, 0, sizeof(a));
    SMCB_REQUIRE(prob && params && lp_out && n > 0 && stride >= n, "bad argument");
    SMCB_REQUIRE(prob->n_params >= 1 && prob->n_params <= SMCB_MAX_PARAMS, "n_params out of range");
    const int cols = build_col_priors(prob, &a.cp);
    if (cols != prob->n_params) {
        set_error("Num prior distributions != num input params");
        return SMCB_ERR_ARG;
    }
    a.prob = *prob;
    a.params = const_cast<double*>(params);
    a.stride = stride;
    a.n = n;
    a.lp = lp_out;
    a.lp_cols = lp_cols_out;
    return dispatch<3>(a, as_stream(stream));
}

extern "C" int smcb_normal_loglike_rows(const double* outputs, int64_t n, int32_t m, const double* data,
                                        const double* sigma_col, double sigma, const double* lp, double* ll_out,
                                        int32_t* flags, void* stream) {
    SMCB_REQUIRE(outputs && data && ll_out && flags && n > 0 && m > 0, "bad argument");
    const int grid = grid_for(n * 32, 256, 1, 8);
    normal_ll_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(outputs, n, m, data, sigma_col, sigma, lp, ll_out, flags);
    return check_launch("normal_ll_rows_kernel");
}

extern "C" int smcb_mh_accept(const smcb_path_t* path, double* params, const double* new_params, int64_t stride,
                              int32_t d, int64_t n, double* ll, const double* new_ll, double* lp, const double* new_lp,
                              uint8_t* moved, int64_t* accept_counts, int32_t step, const smcb_rng_t* rng,
                              void* stream) {
    SMCB_REQUIRE(path && params && new_params && ll && new_ll && lp && new_lp && moved && accept_counts && rng,
                 "NULL argument");
    SMCB_REQUIRE(n > 0 && d > 0 && stride >= n && step >= 0 && step < SMCB_MAX_MCMC_STEPS, "bad sizes");
    AcceptArgs a;
    a.params = params;
    a.new_params = new_params;
    a.stride = stride;
    a.n = n;
    a.d = d;
    a.ll = ll;
    a.new_ll = new_ll;
    a.lp = lp;
    a.new_lp = new_lp;
    a.moved = moved;
    a.accept_counts = reinterpret_cast<unsigned long long*>(accept_counts);
    a.step = step;
    const PathCoef c = make_path_coef(path->phi, path->phi_prev, path->lambda);
    a.phi = c.phi;
    a.pe = c.pe;
    a.qe = c.qe;
    a.rng = *rng;
    mh_accept_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(a);
    return check_launch("mh_accept_kernel");
}

extern "C" int smcb_sample_priors(const smcb_problem_t* prob, double* params, int64_t stride, int64_t n,
                                  uint64_t seed, uint64_t stream_id, int64_t id_offset, void* stream) {
    SMCB_REQUIRE(prob && params && n > 0 && stride >= n, "bad argument");
    ColPriors cp;
    const int cols = build_col_priors(prob, &cp);
    if (cols != prob->n_params) {
        set_error("Num prior distributions != num input params");
        return SMCB_ERR_ARG;
    }
    IwArgs iw;
    memset(&iw, 0, sizeof(iw));
    for (int q = 0; q < prob->n_priors; ++q) {
        const smcb_prior_t& pr = prob->priors[q];
        if (pr.kind == SMCB_PRIOR_UNIFORM) continue;
        if (pr.kind == SMCB_PRIOR_IMPROPER_COV) {
            const int k = iw.n_cov;
            const double dof = pr.a;
            if (!(dof >= pr.ncols && dof <= 256 && dof == (double)(int)dof)) {
                set_error("device inverse-Wishart sampling needs an integer dof in [ncols, 256] (prior %d)", q);
                return SMCB_ERR_UNSUPPORTED;
            }
            iw.c0[k] = cp.cov_c0[k];
            iw.ncols[k] = pr.ncols;
            iw.dof[k] = (int)dof;
            for (int e = 0; e < 6; ++e) iw.chol[k][e] = prob->iw_chol[k][e];
            if (pr.ncols < 3) iw.chol[k][5] = 1.0;                      // padding rows / columns: identity
            if (pr.ncols < 2) iw.chol[k][2] = 1.0;
            if (!(iw.chol[k][0] > 0.0)) {
                set_error("device inverse-Wishart sampling: iw_chol of prior %d is not set", q);
                return SMCB_ERR_ARG;
            }
            iw.n_cov = k + 1;
            continue;
        }
        set_error("device prior sampling supports scipy-uniform and ImproperCov priors only (prior %d)", q);
        return SMCB_ERR_UNSUPPORTED;
    }
    sample_priors_kernel<<<grid_for(n, 256, 1, 8), 256, 0, as_stream(stream)>>>(cp, iw, cols, params, stride, n, seed,
                                                                                  (uint32_t)stream_id, id_offset);
    return check_launch("sample_priors_kernel");
}
