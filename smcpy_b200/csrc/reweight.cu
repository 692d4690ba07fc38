// reweight.cu -- kernel (a): fused tempered reweighting + log-sum-exp normalisation + ESS,
// and kernel (b): device-side bisection of ESS(phi) = target.
//
// Reference: smcpy/paths.py:86-105 (inc_log_weights), smcpy/smc/updater.py:122-133,
// smcpy/smc/particles.py:141-162,200-205, smcpy/smc/samplers.py:250-301.
//
// Bytes per particle (algorithmic): partial pass reads ll, lp, log_w (24 B; 16 B when the
// weights are uniform), finalize pass reads them again and writes log_w, w (16 B).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace smcb {

constexpr int kRwThreads = 256;
constexpr int kRwItems = 4;
constexpr int kLseThreads = 512;    // streaming log-sum-exp passes: 32 warps/SM at 2 blocks hide the load latency
constexpr int kLseBlocksPerSM = 2;

struct RwIn {
    const double* ll;
    const double* lp;
    const double* lq;      // nullptr -> lp
    const double* lw;      // nullptr -> uniform
    double lw_uniform;     // log(1/n_global)
    double lp_const;       // used when lp == nullptr (bisection fast path)
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

// incremental weight only (bisection: previous weights are ignored, samplers.py:264-276)
__device__ __forceinline__ double inc_value(const RwIn& in, int64_t i) {
    const double ll = in.ll[i];
    const double lp = in.lp ? in.lp[i] : in.lp_const;
    const double lq = in.lq ? in.lq[i] : lp;
    return path_inc(ll, lp, lq, in.c);
}

// One thread's share of a streaming log-sum-exp pass: grid-stride, eight independent loads in
// flight, then the (max, S1, S2) update four values at a time; the ragged end is predicated.
template <bool kIncOnly>
__device__ __forceinline__ double lse_val(const RwIn& in, int64_t i, int64_t n) {
    if (i >= n) return kNegInf;
    return kIncOnly ? inc_value(in, i) : rw_value(in, i);
}
template <bool kIncOnly>
__device__ __forceinline__ void lse_stream(const RwIn& in, int64_t n, Lse& acc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (kIncOnly) {          // (the reweight form reads up to four arrays: four wide is its register budget)
        for (; i + 7 * stride < n; i += 8 * stride) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = inc_value(in, i + k * stride);
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

// scalars: [0] max, [1] S1, [2] S2, [3] total_unnorm_log_weight, [4] ESS, [5] log S1
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
    if (lse_partials_body<false>(in, n, partials, ticket, out3, merged) && scalars) write_scalars(merged, scalars);
}

__global__ void lse_merge_kernel(const double* parts, int n_parts, double* sc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Lse a = lse_empty();
    for (int r = 0; r < n_parts; ++r) a = lse_merge(a, Lse{parts[3 * r], parts[3 * r + 1], parts[3 * r + 2]});
    write_scalars(a, sc);
}

// normalised log-weights and weights (particles.py:141-150,130-131):
// log_w = (v - max) - log S1, w = exp(log_w); weights that underflow get log_w = -inf.
__global__ void __launch_bounds__(kRwThreads) rw_finalize_kernel(RwIn in, int64_t n, const double* sc,
                                                                 double* log_w, double* w) {
    const double m = sc[0], ls = sc[5];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // back to front: the partials pass streamed the inputs front to back, so its tail is what
    // the L2 still holds
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

// elementwise path values for host-facing wrappers: mode 0 = logpdf(phi), 1 = inc_log_weights
__global__ void __launch_bounds__(kRwThreads) path_eval_kernel(RwIn in, int64_t n, int mode, double* out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ll = in.ll[i], lp = in.lp[i];
        const double lq = in.lq ? in.lq[i] : lp;
        out[i] = mode ? path_inc(ll, lp, lq, in.c) : path_target(ll, lp, lq, in.c.phi, in.c.pe, in.c.qe);
    }
}

// ---- (b) bisection ----------------------------------------------------------------------
// state: [0] result, [1] n_evals, [2] done, [3] xa, [4] dm, [5] fa, [6] x, [7] margin,
//        [8] full-step margin, [9] phase (0: f(a)... see below)
// Phases mirror AdaptiveSampler.optimize_step (samplers.py:250-259) followed by
// scipy/optimize/Zeros/bisect.c:
//   eval 0: x = 1     -> full-step test (_full_step_meets_target): margin > 0 -> phi = 1, done
//   eval 1: x = phi_old (fa; delta_phi = 0 -> beta = 0 -> ESS = N exactly)  [computed in closed form]
//   eval 2: x = 1 (fb)                                                      [reuses eval 0]
//   eval k: midpoints
constexpr double kXtol = 2e-12;
constexpr double kRtol = 8.881784197001252e-16;
constexpr int kBisectMaxIters = 48;   // (1-0)/2^41 < 2e-12: scipy stops after <= 40 halvings

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
}

__device__ __forceinline__ double margin_from(const Lse& a, double n_global, double target) {
    // predict_ess_margin (samplers.py:264-276): ESS = exp(2 LSE(b) - LSE(2b)) = S1^2/S2,
    // ESS = 0 when either log-sum is -inf
    double ess = 0.0;
    if (a.m != kNegInf && a.s1 > 0.0 && a.s2 > 0.0) ess = (a.s1 * a.s1) / a.s2;
    return ess - n_global * target;
}

__device__ void bisect_advance(double* st, const Lse& a, double phi_old, double target, double n_global) {
    if (st[2] != 0.0) return;
    const double f = margin_from(a, n_global, target);
    st[7] = f;
    st[1] += 1.0;
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

// The whole search in ONE cooperative launch (single GPU): persistent blocks, a grid-wide
// barrier after each evaluation; every block merges the same block partials in the same order
// and advances an identical copy of the scipy state machine, so no block ever disagrees about
// the next evaluation point.  Partials are double-buffered by iteration parity and read with
// ld.global.cg (L2) loads.
__global__ void __launch_bounds__(1024, 1) ess_bisect_coop_kernel(RwIn in, int64_t n, double phi_old,
                                                                    double lambda, double target, double n_global,
                                                                    double* st_out, double* partials, int max_iters) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[96];
    __shared__ double st[16];
    if (threadIdx.x == 0) {
        st[0] = 1.0; st[1] = 0.0; st[2] = 0.0; st[3] = phi_old; st[4] = 1.0 - phi_old;
        st[5] = 0.0; st[6] = 1.0; st[7] = 0.0; st[8] = 0.0; st[9] = 0.0;
    }
    __syncthreads();
    for (int it = 0; it < max_iters; ++it) {
        if (st[2] != 0.0) break;                       // identical in every block
        const double x = st[6];
        in.c.phi = x;
        in.c.pe = fmin(1.0, x / lambda);
        in.c.qe = fmax(0.0, (lambda - x) / lambda);
        Lse acc = lse_empty();
        lse_stream<true>(in, n, acc);
        acc = block_lse(acc, sh);
        double* buf = partials + (size_t)(it & 1) * 3 * gridDim.x;
        if (threadIdx.x == 0) {
            buf[3 * blockIdx.x + 0] = acc.m;
            buf[3 * blockIdx.x + 1] = acc.s1;
            buf[3 * blockIdx.x + 2] = acc.s2;
        }
        grid.sync();
        const Lse m = block_merge_partials(buf, (int)gridDim.x, sh);
        if (threadIdx.x == 0) bisect_advance(st, m, phi_old, target, n_global);
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x < 10) st_out[threadIdx.x] = st[threadIdx.x];
}

// ---- multi-GPU: the same single launch, partial exchange over NVLink peer memory ----------------
// Each rank owns a symmetric slot buffer of kBisSlotWords uint64 mapped by every peer:
// [call parity][iteration parity][rank][4] = (max, S1, S2 bit patterns, flag).  After the local
// grid barrier block 0 stores the rank's merged triple into slot [..][rank] of EVERY rank, fences
// (system scope) and then stores the flag gen * 64 + iteration + 1; thread 0 of every block waits
// for all `world` flags in its OWN memory, merges the triples in rank order (the order of the
// all-gather fallback, bisect_update_kernel) and advances its copy of the search.  A peer can be
// at most one evaluation ahead (it needs this rank's next triple to go further), so two
// iteration parities suffice; the call parity keeps a slow peer's last slot of the previous
// search intact.
constexpr int kBisSlotWords = 2 * 2 * SMCB_MAX_RANKS * 4;
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(1024, 1) ess_bisect_coop_xchg_kernel(RwIn in, int64_t n, double phi_old,
                                                                      double lambda, double target, double n_global,
                                                                      double* st_out, double* partials, int max_iters,
                                                                      smcb_xchg_t x) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[96];
    __shared__ double st[16];
    if (threadIdx.x == 0) {
        st[0] = 1.0; st[1] = 0.0; st[2] = 0.0; st[3] = phi_old; st[4] = 1.0 - phi_old;
        st[5] = 0.0; st[6] = 1.0; st[7] = 0.0; st[8] = 0.0; st[9] = 0.0;
    }
    __syncthreads();
    const size_t call_off = (size_t)(x.gen & 1) * 2 * SMCB_MAX_RANKS * 4;
    for (int it = 0; it < max_iters; ++it) {
        if (st[2] != 0.0) break;                       // identical in every block of every rank
        const double xq = st[6];
        in.c.phi = xq;
        in.c.pe = fmin(1.0, xq / lambda);
        in.c.qe = fmax(0.0, (lambda - xq) / lambda);
        Lse acc = lse_empty();
        lse_stream<true>(in, n, acc);
        acc = block_lse(acc, sh);
        double* buf = partials + (size_t)(it & 1) * 3 * gridDim.x;
        if (threadIdx.x == 0) {
            buf[3 * blockIdx.x + 0] = acc.m;
            buf[3 * blockIdx.x + 1] = acc.s1;
            buf[3 * blockIdx.x + 2] = acc.s2;
        }
        grid.sync();
        const Lse mine = block_merge_partials(buf, (int)gridDim.x, sh);
        if (threadIdx.x == 0) {
            const unsigned long long flag = x.gen * 64ull + (unsigned long long)it + 1ull;
            const size_t off = call_off + (size_t)(it & 1) * SMCB_MAX_RANKS * 4;
            if (blockIdx.x == 0) {
                for (int r = 0; r < x.world; ++r) {
                    unsigned long long* slot = reinterpret_cast<unsigned long long*>(x.peer_slots[r]) + off + 4 * x.rank;
                    st_sys_u64(slot + 0, (unsigned long long)__double_as_longlong(mine.m));
                    st_sys_u64(slot + 1, (unsigned long long)__double_as_longlong(mine.s1));
                    st_sys_u64(slot + 2, (unsigned long long)__double_as_longlong(mine.s2));
                }
                __threadfence_system();
                for (int r = 0; r < x.world; ++r)
                    st_sys_u64(reinterpret_cast<unsigned long long*>(x.peer_slots[r]) + off + 4 * x.rank + 3, flag);
            }
            const unsigned long long* own = reinterpret_cast<const unsigned long long*>(x.peer_slots[x.rank]) + off;
            bool ok = true;
            for (int r = 0; r < x.world && ok; ++r) {
                long long spins = 0;
                while (ld_sys_u64(own + 4 * r + 3) != flag) {
                    __nanosleep(32);
                    if (++spins > (1ll << 25)) { ok = false; break; }      // a peer is gone: give up (~seconds)
                }
            }
            if (!ok) {
                st[2] = -1.0;
            } else {
                __threadfence_system();
                Lse a = lse_empty();
                for (int r = 0; r < x.world; ++r) {
                    Lse b;
                    b.m = __longlong_as_double((long long)ld_sys_u64(own + 4 * r + 0));
                    b.s1 = __longlong_as_double((long long)ld_sys_u64(own + 4 * r + 1));
                    b.s2 = __longlong_as_double((long long)ld_sys_u64(own + 4 * r + 2));
                    a = lse_merge(a, b);
                }
                bisect_advance(st, a, phi_old, target, n_global);
            }
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x < 10) st_out[threadIdx.x] = st[threadIdx.x];
}

static RwIn make_in(const double* ll, const double* lp, const double* lq, const double* lw,
                    int64_t n_global, const smcb_path_t* path) {
    RwIn in;
    in.ll = ll;
    in.lp = lp;
    in.lq = lq;
    in.lw = lw;
    in.lw_uniform = log(1.0 / (double)n_global);
    in.lp_const = 0.0;
    in.c = make_path_coef(path->phi, path->phi_prev, path->lambda);
    return in;
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
    SMCB_REQUIRE(n > 0 && ll && lp && path && scalars && log_w_out && w_out, "bad argument");
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

// co-resident grid of the cooperative kernels (0: cooperative launch unsupported / disabled)
static int coop_config(const void* kernel, int* tpb_out) {
    static int coop_tpb = kLseThreads;
    static int want = -1;
    int dev = 0, coop = 0, per_sm = 0, sms = 0;
    if (want < 0) {
        want = 1;      // one fat block per SM: the grid barrier + partial merge per evaluation scale with the block count
        if (const char* e = getenv("SMCB_COOP_TPB")) coop_tpb = atoi(e);   // tuning aids: threads per block, blocks per SM
        if (const char* e = getenv("SMCB_COOP_BPS")) want = atoi(e);
        if (getenv("SMCB_NO_COOP")) want = 0;                              // launch-chain version
    }
    *tpb_out = coop_tpb;
    if (want == 0) return 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, coop_tpb, 0);
    return (coop && per_sm > 0) ? sms * (per_sm < want ? per_sm : want) : 0;
}

extern "C" int smcb_ess_bisect(int64_t n, const double* ll, const double* lp, const double* lq,
                               double lp_const, double phi_old, double lambda, double target_ess,
                               int64_t n_global, double* state, void* ws, void* stream) {
    SMCB_REQUIRE(n > 0 && ll && state && ws, "bad argument");
    SMCB_REQUIRE(phi_old >= 0.0 && phi_old < 1.0, "phi_old must be in [0, 1)");
    smcb_path_t p{phi_old, phi_old, lambda};
    RwIn in = make_in(ll, lp, lq, nullptr, 1, &p);
    in.lp_const = lp_const;
    char* w = (char*)ws;
    int iters = bisect_iters(phi_old);
    static int coop_blocks = -1, coop_tpb = 0;
    if (coop_blocks < 0) coop_blocks = coop_config((const void*)ess_bisect_coop_kernel, &coop_tpb);
    if (coop_blocks > 0) {
        int grid = grid_for(n, coop_tpb, kRwItems, 8);
        if (grid > coop_blocks) grid = coop_blocks;
        double* partials = (double*)(w + kWsPartials);
        double ng = (double)n_global;
        void* args[] = {&in, &n, &phi_old, &lambda, &target_ess, &ng, &state, &partials, &iters};
        const cudaError_t e = cudaLaunchCooperativeKernel((const void*)ess_bisect_coop_kernel, dim3(grid),
                                                          dim3(coop_tpb), args, 0, as_stream(stream));
        if (e != cudaSuccess) {
            set_error("ess_bisect_coop_kernel: %s", cudaGetErrorString(e));
            return SMCB_ERR_CUDA;
        }
        return SMCB_OK;
    }
    int rc = smcb_ess_bisect_begin(phi_old, state, stream);
    if (rc) return rc;
    const int grid = grid_for(n, kLseThreads, kRwItems, kLseBlocksPerSM);
    for (int it = 0; it < iters; ++it) {
        ess_eval_kernel<<<grid, kLseThreads, 0, as_stream(stream)>>>(
            in, n, phi_old, lambda, target_ess, (double)n_global, state, (double*)(w + kWsPartials),
            (unsigned int*)(w + kWsTickets), (double*)(w + kWsSmall), 1);
    }
    return check_launch("ess_eval_kernel");
}

extern "C" int smcb_ess_bisect_xchg_slot_words(void) { return kBisSlotWords; }

extern "C" int smcb_ess_bisect_xchg(int64_t n, const double* ll, const double* lp, const double* lq,
                                    double lp_const, double phi_old, double lambda, double target_ess,
                                    int64_t n_global, double* state, void* ws, const smcb_xchg_t* x,
                                    void* stream) {
    SMCB_REQUIRE(n > 0 && ll && state && ws && x, "bad argument");
    SMCB_REQUIRE(phi_old >= 0.0 && phi_old < 1.0, "phi_old must be in [0, 1)");
    SMCB_REQUIRE(x->world >= 2 && x->world <= SMCB_MAX_RANKS && x->rank >= 0 && x->rank < x->world && x->gen > 0,
                 "bad exchange descriptor");
    smcb_path_t p{phi_old, phi_old, lambda};
    RwIn in = make_in(ll, lp, lq, nullptr, 1, &p);
    in.lp_const = lp_const;
    char* w = (char*)ws;
    int iters = bisect_iters(phi_old);
    static int coop_blocks = -1, coop_tpb = 0;
    if (coop_blocks < 0) coop_blocks = coop_config((const void*)ess_bisect_coop_xchg_kernel, &coop_tpb);
    if (coop_blocks <= 0) {
        set_error("smcb_ess_bisect_xchg needs cooperative launch");
        return SMCB_ERR_UNSUPPORTED;
    }
    // every rank must take the same number of local blocks only for its own merge: grids may differ
    int grid = grid_for(n, coop_tpb, kRwItems, 8);
    if (grid > coop_blocks) grid = coop_blocks;
    double* partials = (double*)(w + kWsPartials);
    double ng = (double)n_global;
    smcb_xchg_t xv = *x;
    void* args[] = {&in, &n, &phi_old, &lambda, &target_ess, &ng, &state, &partials, &iters, &xv};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)ess_bisect_coop_xchg_kernel, dim3(grid),
                                                      dim3(coop_tpb), args, 0, as_stream(stream));
    if (e != cudaSuccess) {
        set_error("ess_bisect_coop_xchg_kernel: %s", cudaGetErrorString(e));
        return SMCB_ERR_CUDA;
    }
    return SMCB_OK;
}
