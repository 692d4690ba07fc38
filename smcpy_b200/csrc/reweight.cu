// reweight.cu -- kernel (a): fused tempered reweighting + log-sum-exp normalisation + ESS,
// and kernel (b): device-side search for the adaptive phi step, ESS(phi) = target.
//
// Reference: smcpy/paths.py:86-105 (inc_log_weights), smcpy/smc/updater.py:97-133,
// smcpy/smc/particles.py:141-162,200-205, smcpy/smc/samplers.py:250-301.
//
// (a) ONE cooperative launch: every block streams its share of (ll [, lp, lq, log_w]) into an online
//     (max, S1, S2) triple -> grid barrier -> every block merges the block partials in block order
//     (deterministic) [several GPUs: block 0 exchanges the rank triple with every peer over NVLink
//     inside the kernel] -> the normalisation scalars are known to every block -> if the step will NOT
//     resample (ESS >= threshold) the same launch writes log_w and w, walking back to front so that it
//     starts on what the first phase left in L2; if it will resample, nothing else is written: the
//     CDF scan of resample.cu recomputes the weights from ll on the fly.
//     Algorithmic bytes per particle: 8 (ll; flat priors, uniform incoming weights) .. 32 (ll, lp, lq,
//     log_w) read once, + 16 written only when the weights are materialised.
// (b) ONE cooperative launch for the whole search.  ESS(phi) is non-increasing in phi (mean of an
//     exponential family), so the sign scipy's bisection would compute at a midpoint is KNOWN without
//     evaluating it whenever a point with a clearly positive margin lies above it or one with a
//     clearly negative margin below it.  The kernel first closes a bracket around the root with a
//     safeguarded Newton iteration on log ESS in log(phi - phi_old) (each evaluation also returns
//     d ESS / d phi), then replays scipy/optimize/Zeros/bisect.c (xtol 2e-12, rtol 8.88e-16) midpoint
//     by midpoint, evaluating only the midpoints that fall inside the bracket: ~8-12 passes over the
//     particles instead of ~43, same decisions, same phi.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "search.cuh"

namespace smcb {

constexpr int kRwThreads = 256;
constexpr int kRwItems = 4;
constexpr int kLseThreads = 512;    // streaming log-sum-exp passes: 32 warps/SM at 2 blocks hide the load latency
constexpr int kLseBlocksPerSM = 2;

// One thread's share of a streaming log-sum-exp pass: grid-stride, eight independent loads in
// flight, then the (max, S1, S2) update four values at a time; the ragged end is predicated.
template <bool kIncOnly>
__device__ __forceinline__ double lse_val(const RwIn& in, int64_t i, int64_t n) {
    if (i >= n) return kNegInf;
    return kIncOnly ? inc_value(in, i) : rw_value(in, i);
}
template <bool kIncOnly, bool kWide = kIncOnly>
__device__ __forceinline__ void lse_stream(const RwIn& in, int64_t n, Lse& acc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (kWide) {             // (the 64-register reweight form reads up to four arrays: four wide is its budget)
        for (; i + 7 * stride < n; i += 8 * stride) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = kIncOnly ? inc_value(in, i + k * stride) : rw_value(in, i + k * stride);
            lse_push4(acc, v[0], v[1], v[2], v[3]);
            lse_push4(acc, v[4], v[5], v[6], v[7]);
        }
    }
    for (; i < n; i += 4 * stride) {
        const double v0 = lse_val<kIncOnly>(in, i, n), v1 = lse_val<kIncOnly>(in, i + stride, n),
                     v2 = lse_val<kIncOnly>(in, i + 2 * stride, n), v3 = lse_val<kIncOnly>(in, i + 3 * stride, n);
        lse_push4(acc, v0, v1, v2, v3);
    }
}

// ---- the same pass with the derivative sums T1 = sum s e^(v-max), T2 = sum s e^(2(v-max)), s = d v / d phi ----
// (adaptive search only; s = ll for a path without proposal: inc = ll * (phi - phi_old) + const.)
// The (max, S1, S2) arithmetic is EXACTLY lse_push4's, so a margin computed here equals the one the
// plain pass computes.
__device__ __forceinline__ Lse5 lse5_empty() { return Lse5{kNegInf, 0.0, 0.0, 0.0, 0.0}; }
__device__ __forceinline__ void lse5_push4(Lse5& acc, const double (&v)[4], const double (&s)[4]) {
    const double tm = fmax(fmax(v[0], v[1]), fmax(v[2], v[3]));
    if (tm > acc.m) {
        if (acc.m != kNegInf) {
            const double e = exp_nonpos(acc.m - tm);
            acc.s1 *= e;
            acc.s2 *= e * e;
            acc.t1 *= e;
            acc.t2 *= e * e;
        }
        acc.m = tm;
    }
    if (acc.m == kNegInf) return;
    const double e0 = exp_nonpos(v[0] - acc.m), e1 = exp_nonpos(v[1] - acc.m), e2 = exp_nonpos(v[2] - acc.m),
                 e3 = exp_nonpos(v[3] - acc.m);
    acc.s1 += (e0 + e1) + (e2 + e3);
    acc.s2 += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
    // a particle with zero weight contributes nothing, whatever its slope (-inf * 0 would be NaN)
    const double a0 = e0 > 0.0 ? s[0] * e0 : 0.0, a1 = e1 > 0.0 ? s[1] * e1 : 0.0, a2 = e2 > 0.0 ? s[2] * e2 : 0.0,
                 a3 = e3 > 0.0 ? s[3] * e3 : 0.0;
    acc.t1 += (a0 + a1) + (a2 + a3);
    acc.t2 += (a0 * e0 + a1 * e1) + (a2 * e2 + a3 * e3);
}
__device__ __forceinline__ void lse5_stream(const RwIn& in, int64_t n, Lse5& acc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // same grouping as lse_stream<true>: 8-wide main loop (two pushes of four), then predicated fours
    for (; i + 7 * stride < n; i += 8 * stride) {
        double v[8], s[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s[k] = in.ll[i + k * stride];
            v[k] = inc_value(in, i + k * stride);
        }
        const double va[4] = {v[0], v[1], v[2], v[3]}, sa[4] = {s[0], s[1], s[2], s[3]};
        const double vb[4] = {v[4], v[5], v[6], v[7]}, sb[4] = {s[4], s[5], s[6], s[7]};
        lse5_push4(acc, va, sa);
        lse5_push4(acc, vb, sb);
    }
    for (; i < n; i += 4 * stride) {
        double v[4], s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t j = i + k * stride;
            s[k] = (j < n) ? in.ll[j] : 0.0;
            v[k] = lse_val<true>(in, j, n);
        }
        lse5_push4(acc, v, s);
    }
}
__device__ __forceinline__ Lse5 block_lse5(Lse5 v, double* sh) {
    const double bm = block_max_all(v.m, sh);
    const double e = (v.m == kNegInf) ? 0.0 : exp(v.m - bm);
    Lse5 r;
    r.m = bm;
    r.s1 = block_sum(v.s1 * e, sh);
    r.s2 = block_sum(v.s2 * (e * e), sh);
    r.t1 = block_sum(v.t1 * e, sh);
    r.t2 = block_sum(v.t2 * (e * e), sh);
    return r;
}
// merge of n rows (m, s1, s2, t1, t2) by a whole block; same order and arithmetic for (m, s1, s2) as
// block_merge_partials on 3-wide rows
__device__ __forceinline__ Lse5 block_merge_partials5(const double* parts, int n, double* sh) {
    double m = kNegInf;
    for (int b = threadIdx.x; b < n; b += blockDim.x) m = fmax(m, __ldcg(parts + 5 * b));
    const double bm = block_max_all(m, sh);
    double s1 = 0.0, s2 = 0.0, t1 = 0.0, t2 = 0.0;
    if (bm != kNegInf) {
        for (int b = threadIdx.x; b < n; b += blockDim.x) {
            const double mb = __ldcg(parts + 5 * b);
            const double e = (mb == kNegInf) ? 0.0 : exp(mb - bm);
            s1 += __ldcg(parts + 5 * b + 1) * e;
            s2 += __ldcg(parts + 5 * b + 2) * (e * e);
            t1 += __ldcg(parts + 5 * b + 3) * e;
            t2 += __ldcg(parts + 5 * b + 4) * (e * e);
        }
    }
    Lse5 r;
    r.m = bm;
    r.s1 = block_sum(s1, sh);
    r.s2 = block_sum(s2, sh);
    r.t1 = block_sum(t1, sh);
    r.t2 = block_sum(t2, sh);
    return r;
}
__device__ __forceinline__ Lse5 lse5_merge(const Lse5& a, const Lse5& b) {
    Lse5 r;
    r.m = fmax(a.m, b.m);
    if (r.m == kNegInf) {
        r.s1 = r.s2 = r.t1 = r.t2 = 0.0;
        return r;
    }
    const double ea = (a.m == kNegInf) ? 0.0 : exp(a.m - r.m);
    const double eb = (b.m == kNegInf) ? 0.0 : exp(b.m - r.m);
    r.s1 = a.s1 * ea + b.s1 * eb;
    r.s2 = a.s2 * (ea * ea) + b.s2 * (eb * eb);
    r.t1 = a.t1 * ea + b.t1 * eb;
    r.t2 = a.t2 * (ea * ea) + b.t2 * (eb * eb);
    return r;
}

// Per-block (max, S1, S2) partials; the last block to finish merges them in block order
// (deterministic) and writes the merged triple to `out3`.
template <bool kIncOnly>
__device__ __forceinline__ bool lse_partials_body(const RwIn& in, int64_t n, double* partials,
                                                  unsigned int* ticket, double* out3, Lse& merged) {
    __shared__ double sh[96];
    __shared__ bool is_last;
    Lse acc = lse_empty();
    lse_stream<kIncOnly>(in, n, acc);
    acc = block_lse(acc, sh);
    if (threadIdx.x == 0) {
        partials[3 * blockIdx.x + 0] = acc.m;
        partials[3 * blockIdx.x + 1] = acc.s1;
        partials[3 * blockIdx.x + 2] = acc.s2;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // the last block merges the block partials (max first, then independent rescales, fixed order)
    const Lse a = block_merge_partials(partials, (int)gridDim.x, sh);
    if (threadIdx.x == 0) {
        out3[0] = a.m;
        out3[1] = a.s1;
        out3[2] = a.s2;
        *ticket = 0u;
        merged = a;
        return true;          // the one finalizer thread of the grid
    }
    return false;
}

// scalars: [0] max, [1] S1, [2] S2, [3] total_unnorm_log_weight, [4] ESS, [5] log S1,
//          [6] 1 if log_w / w were written by the launch that produced the scalars (cooperative form)
__device__ __forceinline__ void write_scalars(const Lse& a, double* sc) {
    sc[0] = a.m;
    sc[1] = a.s1;
    sc[2] = a.s2;
    const double ls = log(a.s1);
    sc[3] = a.m + ls;                       // = Z0 + log(1 + sum exp(Z - Z0)), particles.py:200-205
    sc[4] = (a.s1 * a.s1) / a.s2;           // = 1 / sum w^2, particles.py:158-162
    sc[5] = ls;
}

// `scalars` != nullptr (single GPU): the finalizer thread also writes the normalisation scalars,
// which saves the separate merge launch.
__global__ void __launch_bounds__(kLseThreads, kLseBlocksPerSM) rw_partials_kernel(RwIn in, int64_t n, double* partials,
                                                                                  unsigned int* ticket, double* out3,
                                                                                  double* scalars) {
    Lse merged;
    if (lse_partials_body<false>(in, n, partials, ticket, out3, merged) && scalars) {
        write_scalars(merged, scalars);
        scalars[6] = 0.0;
    }
}

__global__ void lse_merge_kernel(const double* parts, int n_parts, double* sc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Lse a = lse_empty();
    for (int r = 0; r < n_parts; ++r) a = lse_merge(a, Lse{parts[3 * r], parts[3 * r + 1], parts[3 * r + 2]});
    write_scalars(a, sc);
    sc[6] = 0.0;
}

// normalised log-weights and weights (particles.py:141-150,130-131):
// log_w = (v - max) - log S1, w = exp(log_w); weights that underflow get log_w = -inf.
// Back to front: the partials pass streamed the inputs front to back, so its tail is what the L2
// still holds.
__device__ __forceinline__ void rw_finalize_body(const RwIn& in, int64_t n, double m, double ls, double* log_w,
                                                 double* w) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = n - 1 - ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    for (; i - 3 * stride >= 0; i -= 4 * stride) {          // four independent load/exp/store chains
        double v[4], lw[4], wi[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = rw_value(in, i - k * stride);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            lw[k] = (v[k] - m) - ls;
            wi[k] = exp(lw[k]);
            if (!(wi[k] > 0.0)) lw[k] = (wi[k] != wi[k]) ? wi[k] : kNegInf;   // underflow / -inf / NaN
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            log_w[i - k * stride] = lw[k];
            w[i - k * stride] = wi[k];
        }
    }
    for (; i >= 0; i -= stride) {
        const double v = rw_value(in, i);
        double lw = (v - m) - ls;
        double wi = exp(lw);
        if (!(wi > 0.0)) lw = (wi != wi) ? wi : kNegInf;
        log_w[i] = lw;
        w[i] = wi;
    }
}
__global__ void __launch_bounds__(kRwThreads) rw_finalize_kernel(RwIn in, int64_t n, const double* sc,
                                                                 double* log_w, double* w) {
    rw_finalize_body(in, n, sc[0], sc[5], log_w, w);
}

// elementwise path values for host-facing wrappers: mode 0 = logpdf(phi), 1 = inc_log_weights
__global__ void __launch_bounds__(kRwThreads) path_eval_kernel(RwIn in, int64_t n, int mode, double* out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ll = in.ll[i], lp = in.lp[i];
        const double lq = in.lq ? in.lq[i] : lp;
        out[i] = mode ? path_inc(ll, lp, lq, in.c) : path_target(ll, lp, lq, in.c.phi, in.c.pe, in.c.qe);
    }
}

// ---- exchange of small rank records over NVLink peer memory (multi-GPU, inside a kernel) -----------------
// Each rank owns a symmetric slot buffer of kXSlotWords uint64 mapped by every peer:
// [call parity][round parity][rank][kXRec] = (kXRec - 1 payload words, flag).  ONE thread of the rank
// (thread 0 of block 0) stores its record into slot [..][rank] of EVERY rank, fences (system scope),
// stores the flag gen * 256 + round + 1, then waits for all `world` flags in its OWN memory and reads
// the records in rank order.  A peer can be at most one round ahead (it needs this rank's next record
// to go further), so two round parities suffice; the call parity keeps a slow peer's last slot of the
// previous call intact.  Time-out (option / default 30 s of %globaltimer): returns false.
constexpr int kXRec = 8;
constexpr int kXSlotWords = 2 * 2 * SMCB_MAX_RANKS * kXRec;
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// mine / all: kXRec - 1 doubles per rank.  Called by ONE thread.
__device__ __noinline__ bool xchg_allgather(const smcb_xchg_t& x, int round, const double* mine, int n_vals, double* all,
                               unsigned long long timeout_ns) {
    const size_t off = (size_t)(x.gen & 1) * 2 * SMCB_MAX_RANKS * kXRec + (size_t)(round & 1) * SMCB_MAX_RANKS * kXRec;
    const unsigned long long flag = x.gen * 256ull + (unsigned long long)(round & 255) + 1ull;
    for (int r = 0; r < x.world; ++r) {
        unsigned long long* slot = reinterpret_cast<unsigned long long*>(x.peer_slots[r]) + off + (size_t)kXRec * x.rank;
        for (int k = 0; k < n_vals; ++k) st_sys_u64(slot + k, (unsigned long long)__double_as_longlong(mine[k]));
    }
    __threadfence_system();
    for (int r = 0; r < x.world; ++r)
        st_sys_u64(reinterpret_cast<unsigned long long*>(x.peer_slots[r]) + off + (size_t)kXRec * x.rank + (kXRec - 1), flag);
    const unsigned long long* own = reinterpret_cast<const unsigned long long*>(x.peer_slots[x.rank]) + off;
    const unsigned long long t0 = global_ns();
    // every rank's flag is requested before the first is looked at: a poll costs one load latency, not `world`
    for (unsigned int spins = 0;;) {
        bool all_in = true;
        unsigned long long f[SMCB_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < SMCB_MAX_RANKS; ++r) f[r] = (r < x.world) ? ld_sys_u64(own + (size_t)kXRec * r + (kXRec - 1)) : flag;
#pragma unroll
        for (int r = 0; r < SMCB_MAX_RANKS; ++r) all_in = all_in && (f[r] == flag);
        if (all_in) break;
        __nanosleep(40);
        if ((++spins & 1023u) == 0u && global_ns() - t0 > timeout_ns) return false;
    }
    __threadfence_system();
    for (int r = 0; r < x.world; ++r)
        for (int k = 0; k < n_vals; ++k)
            all[r * n_vals + k] = __longlong_as_double((long long)ld_sys_u64(own + (size_t)kXRec * r + k));
    return true;
}

// one thread: exchange this rank's sums, merge all ranks' in rank order (the order of the all-gather
// fallback), publish to the other blocks through gbuf ([3] / [5] = -1: a peer never answered)
__device__ __noinline__ void xchg_merge3(const smcb_xchg_t& x, const Lse& a, double* gbuf, unsigned long long timeout_ns) {
    const double mine[3] = {a.m, a.s1, a.s2};
    double all[3 * SMCB_MAX_RANKS];
    const bool ok = xchg_allgather(x, 0, mine, 3, all, timeout_ns);
    Lse g = lse_empty();
    if (ok)
        for (int r = 0; r < x.world; ++r) g = lse_merge(g, Lse{all[3 * r], all[3 * r + 1], all[3 * r + 2]});
    gbuf[0] = g.m;
    gbuf[1] = g.s1;
    gbuf[2] = g.s2;
    gbuf[3] = ok ? 1.0 : -1.0;
    __threadfence();
}
__device__ __noinline__ void xchg_merge5(const smcb_xchg_t& x, int round, const Lse5& m, double* gb,
                                         unsigned long long timeout_ns) {
    const double mine[5] = {m.m, m.s1, m.s2, m.t1, m.t2};
    double all[5 * SMCB_MAX_RANKS];
    const bool ok = xchg_allgather(x, round, mine, 5, all, timeout_ns);
    Lse5 g = lse5_empty();
    if (ok)
        for (int r = 0; r < x.world; ++r)
            g = lse5_merge(g, Lse5{all[5 * r], all[5 * r + 1], all[5 * r + 2], all[5 * r + 3], all[5 * r + 4]});
    gb[0] = g.m;
    gb[1] = g.s1;
    gb[2] = g.s2;
    gb[3] = g.t1;
    gb[4] = g.t2;
    gb[5] = ok ? 1.0 : -1.0;
    __threadfence();
}

// ---- (a) cooperative reweight ---------------------------------------------------------------------------
// ess_threshold_n < 0: always write the weights.  Multi-GPU (x.world > 1): `gbuf` (device, 8 doubles)
// carries the merged record from block 0 to the other blocks across a second grid barrier.
__global__ void __launch_bounds__(kLseThreads, 1)
rw_coop_kernel(RwIn in, int64_t n, double* partials, double* scalars, double* log_w, double* w,
               double ess_threshold_n, smcb_xchg_t x, double* gbuf, unsigned long long timeout_ns) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[96];
    __shared__ double sh_m, sh_ls, sh_ess;
    Lse acc = lse_empty();
    lse_stream<false, true>(in, n, acc);
    acc = block_lse(acc, sh);
    if (threadIdx.x == 0) {
        partials[3 * blockIdx.x + 0] = acc.m;
        partials[3 * blockIdx.x + 1] = acc.s1;
        partials[3 * blockIdx.x + 2] = acc.s2;
    }
    grid.sync();
    Lse a = block_merge_partials(partials, (int)gridDim.x, sh);
    if (x.world > 1) {
        if (blockIdx.x == 0 && threadIdx.x == 0) xchg_merge3(x, a, gbuf, timeout_ns);
        grid.sync();
        if (threadIdx.x == 0) {
            a.m = __ldcg(gbuf + 0);
            a.s1 = __ldcg(gbuf + 1);
            a.s2 = __ldcg(gbuf + 2);
        }
    }
    if (threadIdx.x == 0) {
        sh_m = a.m;
        sh_ls = log(a.s1);
        sh_ess = (a.s1 * a.s1) / a.s2;
    }
    __syncthreads();
    const bool write_w = !(sh_ess < ess_threshold_n);          // updater.py:106: resample iff ess < threshold * N
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        write_scalars(a, scalars);
        scalars[6] = write_w ? 1.0 : 0.0;
        scalars[7] = (x.world > 1) ? __ldcg(gbuf + 3) : 1.0;  // -1: a peer never answered
    }
    if (write_w) rw_finalize_body(in, n, sh_m, sh_ls, log_w, w);
}

// ---- (b) adaptive phi search ------------------------------------------------------------------------------
// state (device, 16 doubles, smcb200.h): [0] result, [1] #function evaluations scipy would have made,
// [2] done (1; -1 = a peer never answered), [3] xa, [4] dm, [5] fa, [6] x being evaluated, [7] last margin,
// [8] full-step margin, [9] phase, [10] #passes over the particles, [11..12] bracket around the root.

__global__ void bisect_begin_kernel(double phi_old, double* st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    st[0] = 1.0;
    st[1] = 0.0;
    st[2] = 0.0;
    st[3] = phi_old;
    st[4] = 1.0 - phi_old;
    st[5] = 0.0;
    st[6] = 1.0;      // first evaluation point: full step
    st[7] = 0.0;
    st[8] = 0.0;
    st[9] = 0.0;
    st[10] = 0.0;
}

// plain form: every midpoint is evaluated (launch-chain fallback, proposal paths, option bisect_plain)
__device__ void bisect_advance(double* st, const Lse& a, double phi_old, double target, double n_global) {
    if (st[2] != 0.0) return;
    const double f = margin_from(a, n_global, target);
    st[7] = f;
    st[1] += 1.0;
    st[10] += 1.0;
    const int phase = (int)st[9];
    if (phase == 0) {
        st[8] = f;
        if (f > 0.0) {                 // full step meets the target
            st[0] = 1.0;
            st[2] = 1.0;
            return;
        }
        // scipy: fa = f(xa) with delta_phi = 0 -> ESS = N; fb = f(1) = f
        const double fa = n_global - n_global * target;
        st[1] += 2.0;                  // scipy evaluates f(a) and f(b)
        st[5] = fa;
        if (fa == 0.0) { st[0] = phi_old; st[2] = 1.0; return; }
        if (f == 0.0) { st[0] = 1.0; st[2] = 1.0; return; }
        // first midpoint
        st[4] *= 0.5;
        st[6] = st[3] + st[4];
        st[9] = 1.0;
        return;
    }
    const double xm = st[6], fa = st[5];
    if (f * fa >= 0.0) st[3] = xm;
    if (f == 0.0 || fabs(st[4]) < kXtol + kRtol * fabs(xm)) {
        st[0] = xm;
        st[2] = 1.0;
        return;
    }
    st[4] *= 0.5;
    st[6] = st[3] + st[4];
}

// one ESS evaluation at x = state[6]; the last block advances the bisection (single GPU)
__global__ void __launch_bounds__(kLseThreads, kLseBlocksPerSM) ess_eval_kernel(RwIn in, int64_t n, double phi_old,
                                                             double lambda, double target, double n_global,
                                                             double* st, double* partials,
                                                             unsigned int* ticket, double* out3,
                                                             int advance) {
    if (st[2] != 0.0) return;          // search finished: remaining launches are no-ops
    const double x = st[6];
    in.c.phi = x;
    in.c.pe = fmin(1.0, x / lambda);
    in.c.qe = fmax(0.0, (lambda - x) / lambda);
    Lse merged;
    const bool fin = lse_partials_body<true>(in, n, partials, ticket, out3, merged);
    // single GPU: the finalizer thread advances the search; every other block has already
    // read st[] (it took its ticket after its loads), the next launch is stream-ordered
    if (fin && advance) bisect_advance(st, merged, phi_old, target, n_global);
}

__global__ void bisect_update_kernel(const double* parts, int n_parts, double phi_old, double target,
                                     double n_global, double* st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Lse a = lse_empty();
    for (int r = 0; r < n_parts; ++r) a = lse_merge(a, Lse{parts[3 * r], parts[3 * r + 1], parts[3 * r + 2]});
    bisect_advance(st, a, phi_old, target, n_global);
}

// The whole search in ONE cooperative launch: persistent blocks, a grid-wide barrier after each pass;
// every block merges the same block partials in the same order and advances an identical copy of the
// search, so no block ever disagrees about the next evaluation point.  Partials are double-buffered
// by pass parity and read with ld.global.cg (L2) loads.  x.world > 1: after the local merge block 0
// exchanges the rank record with every peer over NVLink (xchg_allgather) and hands the merged sums to
// the other blocks through `gbuf` across a second barrier; a time-out is seen by all blocks at once.
__global__ void __launch_bounds__(kLseThreads, 1)
ess_search_coop_kernel(RwIn in, int64_t n, double phi_old, double lambda, double target, double n_global, double hint,
                       int plain, double* st_out, double* partials, int max_passes, smcb_xchg_t x, double* gbuf,
                       unsigned long long timeout_ns) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[96];
    __shared__ Search s;
    __shared__ int status;
    if (threadIdx.x == 0) {
        search_init(s, phi_old, target, n_global, hint, plain);
        status = 0;
    }
    __syncthreads();
    for (int it = 0; it < max_passes; ++it) {
        if (s.done || status < 0) break;               // identical in every block of every rank
        const double xq = s.x_eval;
        in.c.phi = xq;
        in.c.pe = fmin(1.0, xq / lambda);
        in.c.qe = fmax(0.0, (lambda - xq) / lambda);
        Lse5 acc = lse5_empty();
        lse5_stream(in, n, acc);
        acc = block_lse5(acc, sh);
        double* buf = partials + (size_t)(it & 1) * 5 * gridDim.x;
        if (threadIdx.x == 0) {
            buf[5 * blockIdx.x + 0] = acc.m;
            buf[5 * blockIdx.x + 1] = acc.s1;
            buf[5 * blockIdx.x + 2] = acc.s2;
            buf[5 * blockIdx.x + 3] = acc.t1;
            buf[5 * blockIdx.x + 4] = acc.t2;
        }
        grid.sync();
        Lse5 m = block_merge_partials5(buf, (int)gridDim.x, sh);
        if (x.world > 1) {
            double* gb = gbuf + (size_t)(it & 1) * 8;
            if (blockIdx.x == 0 && threadIdx.x == 0) xchg_merge5(x, it, m, gb, timeout_ns);
            grid.sync();
            if (threadIdx.x == 0) {
                m.m = __ldcg(gb + 0); m.s1 = __ldcg(gb + 1); m.s2 = __ldcg(gb + 2);
                m.t1 = __ldcg(gb + 3); m.t2 = __ldcg(gb + 4);
                if (__ldcg(gb + 5) < 0.0) status = -1;
            }
        }
        if (threadIdx.x == 0 && status == 0) search_advance(s, m);
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) search_export(s, st_out, status);
}

static RwIn make_in(const double* ll, const double* lp, const double* lq, const double* lw,
                    int64_t n_global, const smcb_path_t* path) {
    return make_rw_in(ll, lp, lq, lw, n_global, path);
}

// co-resident grid of the cooperative kernels (0: cooperative launch unsupported / disabled)
static int coop_blocks_for(const void* kernel, int tpb, int want_per_sm) {
    int dev = 0, coop = 0, per_sm = 0, sms = 0;
    if (options().no_coop) return 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, tpb, 0) != cudaSuccess) per_sm = 0;
    return (coop && per_sm > 0) ? sms * (per_sm < want_per_sm ? per_sm : want_per_sm) : 0;
}

static unsigned long long xchg_timeout_ns() {
    const int s = options().xchg_timeout_s;
    return (unsigned long long)(s > 0 ? s : 30) * 1000000000ull;
}

static int check_xchg(const smcb_xchg_t* x, const char* who) {
    if (!x) return SMCB_OK;
    if (!(x->world >= 2 && x->world <= SMCB_MAX_RANKS && x->rank >= 0 && x->rank < x->world && x->gen > 0)) {
        set_error("%s: bad exchange descriptor", who);
        return SMCB_ERR_ARG;
    }
    for (int r = 0; r < x->world; ++r)
        if (!x->peer_slots[r]) {
            set_error("%s: peer slot pointer is NULL", who);
            return SMCB_ERR_ARG;
        }
    return SMCB_OK;
}

}  // namespace smcb

using namespace smcb;

extern "C" int smcb_path_eval(int64_t n, const double* ll, const double* lp, const double* lq,
                              const smcb_path_t* path, int32_t mode, double* out, void* stream) {
    SMCB_REQUIRE(n > 0 && ll && lp && path && out && (mode == 0 || mode == 1), "bad argument");
    RwIn in = make_in(ll, lp, lq, nullptr, 1, path);
    path_eval_kernel<<<grid_for(n, kRwThreads, 1, 8), kRwThreads, 0, as_stream(stream)>>>(in, n, mode, out);
    return check_launch("path_eval_kernel");
}

extern "C" int smcb_reweight_partials(int64_t n, const double* ll, const double* lp, const double* lq,
                                      const double* log_w_in, int64_t n_global, const smcb_path_t* path,
                                      double* partial_out, void* ws, void* stream) {
    SMCB_REQUIRE(n > 0 && ll && lp && path && partial_out && ws, "bad argument");
    RwIn in = make_in(ll, lp, lq, log_w_in, n_global, path);
    const int grid = grid_for(n, kLseThreads, kRwItems, kLseBlocksPerSM);
    char* w = (char*)ws;
    rw_partials_kernel<<<grid, kLseThreads, 0, as_stream(stream)>>>(
        in, n, (double*)(w + kWsPartials), (unsigned int*)(w + kWsTickets), partial_out, nullptr);
    return check_launch("rw_partials_kernel");
}

extern "C" int smcb_lse_merge(const double* partials, int32_t n_parts, double* scalars_out, void* stream) {
    SMCB_REQUIRE(partials && n_parts > 0 && scalars_out, "bad argument");
    lse_merge_kernel<<<1, 32, 0, as_stream(stream)>>>(partials, n_parts, scalars_out);
    return check_launch("lse_merge_kernel");
}

extern "C" int smcb_reweight_finalize(int64_t n, const double* ll, const double* lp, const double* lq,
                                      const double* log_w_in, int64_t n_global, const smcb_path_t* path,
                                      const double* scalars, double* log_w_out, double* w_out,
                                      void* stream) {
    SMCB_REQUIRE(n > 0 && ll && path && scalars && log_w_out && w_out, "bad argument");
    RwIn in = make_in(ll, lp, lq, log_w_in, n_global, path);
    const int grid = grid_for(n, kRwThreads, kRwItems, 8);
    rw_finalize_kernel<<<grid, kRwThreads, 0, as_stream(stream)>>>(in, n, scalars, log_w_out, w_out);
    return check_launch("rw_finalize_kernel");
}

extern "C" int smcb_reweight(int64_t n, const double* ll, const double* lp, const double* lq,
                             const double* log_w_in, int64_t n_global, const smcb_path_t* path,
                             double* log_w_out, double* w_out, double* scalars_out, void* ws,
                             void* stream) {
    SMCB_REQUIRE(n > 0 && ll && lp && path && scalars_out && ws, "bad argument");
    char* w = (char*)ws;
    RwIn in = make_in(ll, lp, lq, log_w_in, n_global, path);
    const int grid = grid_for(n, kLseThreads, kRwItems, kLseBlocksPerSM);
    rw_partials_kernel<<<grid, kLseThreads, 0, as_stream(stream)>>>(
        in, n, (double*)(w + kWsPartials), (unsigned int*)(w + kWsTickets), (double*)(w + kWsSmall), scalars_out);
    int rc = check_launch("rw_partials_kernel");
    if (rc) return rc;
    return smcb_reweight_finalize(n, ll, lp, lq, log_w_in, n_global, path, scalars_out, log_w_out, w_out,
                                  stream);
}

extern "C" int smcb_reweight_fused(int64_t n, const double* ll, const double* lp, double lp_const, const double* lq,
                                   const double* log_w_in, int64_t n_global, const smcb_path_t* path,
                                   double ess_threshold, double* log_w_out, double* w_out, double* scalars_out,
                                   void* ws, const smcb_xchg_t* x, void* stream) {
    SMCB_REQUIRE(n > 0 && ll && path && scalars_out && ws && log_w_out && w_out, "bad argument");
    int rc = check_xchg(x, "smcb_reweight_fused");
    if (rc) return rc;
    char* w = (char*)ws;
    RwIn in = make_in(ll, lp, lq, log_w_in, n_global, path);
    in.lp_const = lp_const;
    static int coop_blocks = -1;
    if (coop_blocks < 0) coop_blocks = coop_blocks_for((const void*)rw_coop_kernel, kLseThreads, 1);
    if (coop_blocks <= 0 || options().no_coop) {
        set_error("smcb_reweight_fused needs cooperative launch");
        return SMCB_ERR_UNSUPPORTED;
    }
    int grid = grid_for(n, kLseThreads, kRwItems, 1);
    if (grid > coop_blocks) grid = coop_blocks;
    double* partials = (double*)(w + kWsPartials);
    double* gbuf = (double*)(w + kWsSmall + 256);
    double thr_n = ess_threshold < 0.0 ? -1.0 : ess_threshold * (double)n_global;
    smcb_xchg_t xv;
    memset(&xv, 0, sizeof(xv));
    if (x) xv = *x;
    unsigned long long tmo = xchg_timeout_ns();
    void* args[] = {&in, &n, &partials, &scalars_out, &log_w_out, &w_out, &thr_n, &xv, &gbuf, &tmo};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)rw_coop_kernel, dim3(grid), dim3(kLseThreads), args,
                                                      0, as_stream(stream));
    if (e != cudaSuccess) {
        set_error("rw_coop_kernel: %s", cudaGetErrorString(e));
        return SMCB_ERR_CUDA;
    }
    return SMCB_OK;
}

extern "C" int smcb_ess_bisect_max_iters(void) { return kBisectMaxIters; }

extern "C" int smcb_ess_bisect_begin(double phi_old, double* state, void* stream) {
    SMCB_REQUIRE(state, "bad argument");
    bisect_begin_kernel<<<1, 32, 0, as_stream(stream)>>>(phi_old, state);
    return check_launch("bisect_begin_kernel");
}

extern "C" int smcb_ess_partial(int64_t n, const double* ll, const double* lp, const double* lq,
                                double lp_const, double phi_old, double lambda, const double* state,
                                double* partial_out, void* ws, void* stream) {
    SMCB_REQUIRE(n > 0 && ll && state && partial_out && ws, "bad argument");
    smcb_path_t p{phi_old, phi_old, lambda};
    RwIn in = make_in(ll, lp, lq, nullptr, 1, &p);
    in.lp_const = lp_const;
    const int grid = grid_for(n, kLseThreads, kRwItems, kLseBlocksPerSM);
    char* w = (char*)ws;
    ess_eval_kernel<<<grid, kLseThreads, 0, as_stream(stream)>>>(
        in, n, phi_old, lambda, 0.0, 0.0, const_cast<double*>(state), (double*)(w + kWsPartials),
        (unsigned int*)(w + kWsTickets), partial_out, 0);
    return check_launch("ess_eval_kernel");
}

extern "C" int smcb_ess_bisect_update(const double* partials, int32_t n_parts, double phi_old,
                                      double target_ess, int64_t n_global, double* state, void* stream) {
    SMCB_REQUIRE(partials && n_parts > 0 && state, "bad argument");
    bisect_update_kernel<<<1, 32, 0, as_stream(stream)>>>(partials, n_parts, phi_old, target_ess,
                                                          (double)n_global, state);
    return check_launch("bisect_update_kernel");
}

static int bisect_iters(double phi_old) {
    // 1 full-step evaluation + halvings until dm < xtol (bisect.c stopping rule), + slack
    const double span = 1.0 - phi_old;
    int k = 3;
    if (span > kXtol) k = 3 + (int)ceil(log2(span / kXtol));
    return k < kBisectMaxIters ? k : kBisectMaxIters;
}

static int search_launch(int64_t n, const double* ll, const double* lp, const double* lq, double lp_const,
                         double phi_old, double lambda, double target_ess, int64_t n_global, double hint,
                         double* state, void* ws, const smcb_xchg_t* x, void* stream) {
    smcb_path_t p{phi_old, phi_old, lambda};
    RwIn in = make_in(ll, lp, lq, nullptr, 1, &p);
    in.lp_const = lp_const;
    char* w = (char*)ws;
    static int coop_blocks = -1;
    static int coop_tpb = 0;
    if (coop_blocks < 0) {
        coop_tpb = (options().coop_tpb > 0 && options().coop_tpb <= kLseThreads) ? options().coop_tpb : kLseThreads;
        // one fat block per SM: the grid barrier + partial merge per evaluation scale with the block count
        coop_blocks = coop_blocks_for((const void*)ess_search_coop_kernel, coop_tpb,
                                      options().coop_bps > 0 ? options().coop_bps : 1);
    }
    if (coop_blocks <= 0 || options().no_coop) return -1;       // caller falls back
    int grid = grid_for(n, coop_tpb, kRwItems, 8);
    if (grid > coop_blocks) grid = coop_blocks;
    double* partials = (double*)(w + kWsPartials);
    double* gbuf = (double*)(w + kWsSmall + 512);
    double ng = (double)n_global;
    // the monotone bracket needs inc = ll * (phi - phi_old) + const: no proposal column, no kink at lambda
    int plain = (options().bisect_plain || lq != nullptr) ? 1 : 0;
    int max_passes = plain ? bisect_iters(phi_old) : kSearchMaxPasses;
    smcb_xchg_t xv;
    memset(&xv, 0, sizeof(xv));
    if (x) xv = *x;
    unsigned long long tmo = xchg_timeout_ns();
    void* args[] = {&in, &n, &phi_old, &lambda, &target_ess, &ng, &hint, &plain, &state, &partials, &max_passes,
                    &xv, &gbuf, &tmo};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)ess_search_coop_kernel, dim3(grid), dim3(coop_tpb),
                                                      args, 0, as_stream(stream));
    if (e != cudaSuccess) {
        set_error("ess_search_coop_kernel: %s", cudaGetErrorString(e));
        return SMCB_ERR_CUDA;
    }
    return SMCB_OK;
}

extern "C" int smcb_ess_bisect(int64_t n, const double* ll, const double* lp, const double* lq,
                               double lp_const, double phi_old, double lambda, double target_ess,
                               int64_t n_global, double* state, void* ws, void* stream) {
    return smcb_ess_search(n, ll, lp, lq, lp_const, phi_old, lambda, target_ess, n_global, 0.0, state, ws, nullptr,
                           stream);
}

extern "C" int smcb_ess_search(int64_t n, const double* ll, const double* lp, const double* lq, double lp_const,
                               double phi_old, double lambda, double target_ess, int64_t n_global, double hint_dphi,
                               double* state, void* ws, const smcb_xchg_t* x, void* stream) {
    SMCB_REQUIRE(n > 0 && ll && state && ws, "bad argument");
    SMCB_REQUIRE(phi_old >= 0.0 && phi_old < 1.0, "phi_old must be in [0, 1)");
    SMCB_REQUIRE(target_ess > 0.0 && target_ess <= 1.0 && lambda > 0.0, "target_ess in (0, 1], lambda > 0");
    int rc = check_xchg(x, "smcb_ess_search");
    if (rc) return rc;
    rc = search_launch(n, ll, lp, lq, lp_const, phi_old, lambda, target_ess, n_global,
                       hint_dphi > 0.0 ? hint_dphi : 0.0, state, ws, x, stream);
    if (rc >= 0) return rc;
    if (x) {
        set_error("smcb_ess_search: the multi-GPU form needs cooperative launch");
        return SMCB_ERR_UNSUPPORTED;
    }
    // launch-chain fallback (no cooperative launch): every scipy midpoint is evaluated
    char* w = (char*)ws;
    smcb_path_t p{phi_old, phi_old, lambda};
    RwIn in = make_in(ll, lp, lq, nullptr, 1, &p);
    in.lp_const = lp_const;
    rc = smcb_ess_bisect_begin(phi_old, state, stream);
    if (rc) return rc;
    const int grid = grid_for(n, kLseThreads, kRwItems, kLseBlocksPerSM);
    const int iters = bisect_iters(phi_old);
    for (int it = 0; it < iters; ++it) {
        ess_eval_kernel<<<grid, kLseThreads, 0, as_stream(stream)>>>(
            in, n, phi_old, lambda, target_ess, (double)n_global, state, (double*)(w + kWsPartials),
            (unsigned int*)(w + kWsTickets), (double*)(w + kWsSmall), 1);
    }
    return check_launch("ess_eval_kernel");
}

extern "C" int smcb_xchg_slot_words(void) { return kXSlotWords; }
extern "C" int smcb_ess_bisect_xchg_slot_words(void) { return kXSlotWords; }

extern "C" int smcb_ess_bisect_xchg(int64_t n, const double* ll, const double* lp, const double* lq,
                                    double lp_const, double phi_old, double lambda, double target_ess,
                                    int64_t n_global, double* state, void* ws, const smcb_xchg_t* x,
                                    void* stream) {
    SMCB_REQUIRE(x, "bad argument");
    return smcb_ess_search(n, ll, lp, lq, lp_const, phi_old, lambda, target_ess, n_global, 0.0, state, ws, x, stream);
}

// ---- host-side TEST HOOK for the search state machine (see search.cuh) --------------------------------
// Sums of one evaluation computed on the CPU (plain loops, libm exp), then the SAME search_init /
// search_advance code as the kernel.  `ll` is a HOST array.  Used by tests/test_search_cpu.py to compare
// the decisions with scipy.optimize.bisect on the margin function exported below; never by the product.
static Lse5 host_sums(const double* ll, int64_t n, double phi_old, double x) {
    Lse5 a{-__builtin_huge_val(), 0.0, 0.0, 0.0, 0.0};
    for (int64_t i = 0; i < n; ++i) {
        double v = ll[i] * x - ll[i] * phi_old;
        if (v != v) v = -__builtin_huge_val();
        if (v > a.m) a.m = v;
    }
    if (a.m == -__builtin_huge_val()) return a;
    for (int64_t i = 0; i < n; ++i) {
        double v = ll[i] * x - ll[i] * phi_old;
        if (v != v) v = -__builtin_huge_val();
        const double e = exp(v - a.m);
        a.s1 += e;
        a.s2 += e * e;
        if (e > 0.0) {
            a.t1 += ll[i] * e;
            a.t2 += ll[i] * e * e;
        }
    }
    return a;
}

extern "C" double smcb_test_search_margin(const double* ll_host, int64_t n, double phi_old, double x, double target) {
    const Lse5 a = host_sums(ll_host, n, phi_old, x);
    return margin_from(Lse{a.m, a.s1, a.s2}, (double)n, target);
}

extern "C" int smcb_test_search_host(const double* ll_host, int64_t n, double phi_old, double target, double hint,
                                     int32_t plain, double* state_out, double* trace_out, int32_t trace_len) {
    if (!ll_host || n < 1 || !state_out) {
        set_error("smcb_test_search_host: bad argument");
        return SMCB_ERR_ARG;
    }
    Search s;
    search_init(s, phi_old, target, (double)n, hint, plain);
    int it = 0;
    for (; it < kSearchMaxPasses && !s.done; ++it) {
        if (trace_out && it < trace_len) trace_out[it] = s.x_eval;
        const Lse5 a = host_sums(ll_host, n, phi_old, s.x_eval);
        search_advance(s, a);
    }
    search_export(s, state_out, 0);
    return SMCB_OK;
}
