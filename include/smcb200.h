/*
 * smcb200.h -- C ABI of libsmcb200.so: the B200 (sm_100a) implementation of
 * nasa/SMCPy's data-parallel particle hot path
 *     reweight -> ESS / phi bisection -> resample -> covariance -> K x Metropolis.
 *
 * The reference (pure Python, v0.1.9) has no FFI; each entry point below names the
 * reference function(s) it replaces (file:line relative to the reference checkout).
 * Conventions
 *   - plain pointers and sizes only; every array pointer is a DEVICE pointer unless
 *     the parameter is documented as host; `stream` is a cudaStream_t passed as void*;
 *   - nothing is allocated inside: the caller owns all buffers, including one scratch
 *     `ws` of at least smcb_workspace_bytes(n, d) bytes, zero-initialised once
 *     (every kernel leaves its tickets / counters reset);
 *   - all calls are asynchronous on `stream`; results the host needs are written to
 *     device scalars and read at step boundaries;
 *   - return value 0 = ok, otherwise an SMCB_ERR_* code; smcb_last_error() gives text;
 *   - particles are structure-of-arrays: parameter c of particle i is
 *     params[c * stride + i]; indices are int64.
 *   - precision: every quantity the reference computes -- priors, model, log-likelihood, path
 *     arithmetic, acceptance ratio, weights, ESS, phi, CDF, covariance, Cholesky -- is IEEE fp64.
 *     The only narrower arithmetic is inside the NATIVE proposal generator (rng.mode 0), which has no
 *     counterpart in the reference (its draws come from numpy): standard normals are Box-Muller in
 *     fp32 (24-bit radius uniform, |z| <= 5.9), and the wide identity kernels (d >= 16) form the
 *     increment sqrt(s) L z on the tensor cores with tf32 inputs / fp32 accumulation unless the
 *     option "strict_proposal" selects fp64 FMAs.  A Metropolis proposal only has to be symmetric,
 *     which both keep exactly; replay mode (rng.mode 1) uses the host's fp64 normals throughout.
 */
#ifndef SMCB200_H
#define SMCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMCB_VERSION 100            /* 0.1.0 */
#define SMCB_MAX_PARAMS 32          /* fused Metropolis kernels: d <= 32 */
#define SMCB_MAX_PRIORS 32
#define SMCB_MAX_COV_ARGS 6         /* MVNormal: F <= 3 features -> F(F+1)/2 packed terms */
#define SMCB_MAX_MCMC_STEPS 4096    /* accept-count history length per mutate call */
#define SMCB_MVN_STATS 12           /* MVNormal snapshot statistics: mean[3] + scatter[3][3] (smcb_mvn_scatter) */

enum { SMCB_OK = 0, SMCB_ERR_ARG = 1, SMCB_ERR_CUDA = 2, SMCB_ERR_UNSUPPORTED = 3 };

/* built-in device models: a model maps (N, n_model_params) -> (N, n_data) */
enum {
    SMCB_MODEL_LINEAR = 1,       /* y_j = a * x_j + b           (README.md:64-79) */
    SMCB_MODEL_SPRING_MASS = 2,  /* x'' = -K x + g, RK4 device integrator
                                    (examples/spring_mass/spring_mass_model.py:8-52) */
    SMCB_MODEL_IDENTITY = 3      /* y_j = theta_j               (config C5) */
};
/* log-likelihoods (smcpy/log_likelihoods.py) */
enum { SMCB_LL_NORMAL = 1 /* :22-49 */, SMCB_LL_MVNORMAL = 2 /* :121-192 */, SMCB_LL_MULTISOURCE = 3 /* :52-118 */ };
#define SMCB_MAX_SEGMENTS 8
/* priors (smcpy/mcmc/vector_mcmc.py:98-114) */
enum {
    SMCB_PRIOR_UNIFORM = 1,           /* scipy.stats.uniform(loc=a, scale=b) */
    SMCB_PRIOR_IMPROPER_UNIFORM = 2,  /* smcpy/priors.py:6-35, bounds [a, b] inclusive */
    SMCB_PRIOR_IMPROPER_COV = 3,      /* smcpy/priors.py:93-147, ncols x ncols PD test; a = inverse-Wishart dof of
                                         its rvs (:118-123), scale factor in smcb_problem_t::iw_chol */
    SMCB_PRIOR_EXTERNAL = 4           /* any other prior object (InvWishart :38-90, ImproperConstrainedUniform
                                         :150-247, user classes): `ncols` parameter columns whose log-pdf the HOST
                                         evaluates and adds to the log-prior column of the un-fused route
                                         (smcb_mh_propose / smcb_mh_accept); the device treats it as 0 and the
                                         fused kernels refuse it */
};
/* device error-flag bits (checked by the host at step boundaries) */
enum {
    SMCB_FLAG_PRIOR_OOB = 1,   /* incoming particle has zero prior (vector_mcmc.py:199-203) */
    SMCB_FLAG_MODEL_NAN = 2,   /* NaN in model output (log_likelihoods.py:14-15) */
    SMCB_FLAG_COV_NOT_PD = 4,  /* MVNormal covariance not PD where the likelihood needed it */
    SMCB_FLAG_CHOL_FAIL = 8,   /* proposal covariance not PD (vector_mcmc.py:224-236) */
    SMCB_FLAG_COV_RESHIFT = 16, /* one-pass covariance: the shift was too far from the mean, redo in two passes */
    SMCB_FLAG_XCHG_TIMEOUT = 32 /* smcb_mh_step_xchg: a peer's accept count never arrived (option xchg_timeout_s) */
};
/* resampling uniforms (smcpy/resampler_rngs.py:4-9; systematic is an extension) */
enum { SMCB_RESAMPLE_STANDARD = 0, SMCB_RESAMPLE_STRATIFIED = 1, SMCB_RESAMPLE_SYSTEMATIC = 2 };
/* CDF scan modes */
enum {
    SMCB_SCAN_SEQUENTIAL = 0,  /* one fp64 add chain, bit-equal to np.cumsum (parity mode) */
    SMCB_SCAN_LOOKBACK = 1     /* single-pass decoupled look-back, deterministic (fast mode) */
};

/* initial-particle proposal q(x) of GeometricPath(proposal=...): independent components
 * (smcpy/proposals.py:4-34 MultivarIndependent over scipy.stats distributions) */
enum { SMCB_PROPOSAL_NORMAL = 1 /* norm(loc=a, scale=b) */, SMCB_PROPOSAL_UNIFORM = 2 /* uniform(loc=a, scale=b) */ };
typedef struct smcb_proposal_col {
    int32_t kind;    /* SMCB_PROPOSAL_* */
    int32_t pad;
    double a, b;     /* loc, scale */
} smcb_proposal_col_t;

typedef struct smcb_prior {
    int32_t kind;    /* SMCB_PRIOR_* */
    int32_t ncols;   /* IMPROPER_COV: matrix size; else 1 */
    double a, b;     /* UNIFORM: loc, scale; IMPROPER_UNIFORM: lower, upper */
    double logp_in;  /* log-density inside the support (-log(scale), or 0) */
} smcb_prior_t;

/* model + likelihood + priors of one calibration problem; plain data, passed by value
 * to the kernels.  Replaces the VectorMCMC constructor arguments
 * (smcpy/mcmc/vector_mcmc.py:11-34). */
typedef struct smcb_problem {
    int32_t model_id;        /* SMCB_MODEL_* */
    int32_t loglike_id;      /* SMCB_LL_* */
    int32_t n_params;        /* d: all sampled columns, incl. sigma / covariance terms */
    int32_t n_model_params;  /* leading columns handed to the model */
    int32_t n_data;          /* Normal: M data points; MVNormal: F features */
    int32_t n_snapshots;     /* MVNormal: S rows of data; else 1 */
    int32_t sigma_is_param;  /* Normal: 1 -> last column is the std-dev (args=None) */
    int32_t n_priors;
    double sigma;            /* Normal: fixed std-dev when sigma_is_param == 0 */
    double model_args[8];    /* SPRING_MASS: x0, v0, dt, nsub */
    const double* data;      /* device: y[M] or data[S][F] row-major */
    const double* xgrid;     /* device: LINEAR x[M] (MVNormal: X[F]); else NULL */
    double cov_fixed[SMCB_MAX_COV_ARGS];      /* MVNormal packed upper-tri, fixed values */
    int32_t cov_param_col[SMCB_MAX_COV_ARGS]; /* -1 = fixed, else parameter column */
    smcb_prior_t priors[SMCB_MAX_PRIORS];
    int32_t has_proposal;    /* 1: the path has a proposal; one entry per parameter column below */
    int32_t pad;
    smcb_proposal_col_t proposal[SMCB_MAX_PARAMS];
    /* MultiSourceNormal: consecutive data segments with their own std-dev */
    int32_t n_segments;
    int32_t seg_len[SMCB_MAX_SEGMENTS];
    int32_t seg_sigma_col[SMCB_MAX_SEGMENTS]; /* -1 = fixed (seg_sigma), else parameter column */
    int32_t pad2;
    double seg_sigma[SMCB_MAX_SEGMENTS];
    /* MVNormal: device, SMCB_MVN_STATS doubles written by smcb_mvn_scatter(data, S, F, ...) once per problem */
    const double* mvn_stats;
    /* IMPROPER_COV priors, in prior order (at most 2): lower Cholesky factor of the inverse-Wishart scale
     * matrix of their rvs, packed row-major (l00, l10, l11, l20, l21, l22); used by smcb_sample_priors only */
    double iw_chol[2][6];
    /* LINEAR model + NORMAL likelihood: device, 2 + 2 M doubles written by smcb_linear_center(xgrid, data, M, ...)
     * once per problem: xbar, ybar, then the pairs (x_j - xbar, -(y_j - ybar)) */
    const double* lin_centered;
} smcb_problem_t;

/* GeometricPath state (smcpy/paths.py:64-112): exponents are derived inside the
 * library exactly as paths.py:94-95 does. */
typedef struct smcb_path {
    double phi, phi_prev, lambda;
} smcb_path_t;

/* random numbers for one Metropolis step.  mode 0: Philox4x32-10 keyed by
 * (seed; global particle id, stream) -- results do not depend on the GPU count;
 * mode 1: replay a host tape drawn in the reference's order:
 * z = normal(0,1,(N,d)) row-major, u = uniform(0,1,(N,1)). */
typedef struct smcb_rng {
    int32_t mode;
    uint64_t seed;
    uint64_t stream;     /* unique per use (host-incremented counter) */
    int64_t id_offset;   /* global id of local particle 0 (multi-GPU shards) */
    const double* z;     /* replay: device (N, d) row-major */
    const double* u;     /* replay: device (N) */
} smcb_rng_t;

/* cross-GPU accept-count exchange fused into the Metropolis kernel (multi-GPU only).  Each rank
 * owns a symmetric `slots` buffer of smcb_mh_xchg_slot_words() uint64 (two tables of SMCB_MAX_MCMC_STEPS x
 * SMCB_MAX_RANKS words: FINAL counts, then EARLY counts) mapped by every peer.  The last block of step k
 * stores (gen << 32 | local count) into final slot [k][rank] of EVERY rank over NVLink; the dynamic-tile
 * kernels also store a running count into early slot [k][rank] once all but 15 % / world of N are counted.
 * The prologue of step k+1 reads the `world` slots of step k in its own memory and applies the
 * covariance-scale rule (vector_mcmc.py:55-60) as soon as the (bounds on the) global count leave one
 * answer -- no separate collective launch, and usually no wait for the slowest rank's kernel to end.
 * `gen` increases with every mutate call. */
#define SMCB_MAX_RANKS 8
typedef struct smcb_xchg {
    int32_t world, rank;
    uint64_t gen;
    uint64_t* peer_slots[SMCB_MAX_RANKS];   /* device pointers to each rank's slots buffer */
    int64_t* global_counts;                 /* device, SMCB_MAX_MCMC_STEPS: per step a count with the scale decision of the
                                             * global count (the global count itself when every final slot had arrived) */
    uint32_t* ticket;                       /* device, 4 zero-initialised words, 8-byte aligned: [0] block ticket,
                                             * [2..3] progress word (particles counted << 32 | accepted) */
} smcb_xchg_t;

/* ---- library --------------------------------------------------------------------- */
int smcb_version(void);
const char* smcb_last_error(void);
size_t smcb_workspace_bytes(int64_t n, int32_t d);
/* Runtime options (tuning and test aids; defaults come from the SMCB_<NAME> environment variables):
 * they choose between kernel variants that compute the same step, so that the parity tests can pin
 * every variant the dispatcher may take.  Names: "no_tma", "strict_proposal" (wide identity
 * Metropolis kernel: fp64 FMAs for sqrt(s) L z instead of the tf32 tensor-core product),
 * "linear_cfg", "persist", "no_dmma", "cov_two_pass", "no_coop", "coop_tpb", "coop_bps",
 * "bisect_plain", "spring_ppt" (-1 auto, 1 / 2 interleaved RK4 chains per thread), "no_mvn3" (generic MVNormal
 * kernel instead of the static 3-feature layout), "no_pref" (that kernel without its cp.async prefetch), "no_fused_resample", "xchg_timeout_s", "no_pdl" (smcb_mh_steps without
 * programmatic dependent launches), "spring_form" (spring-mass kernels: 1 = RK4 propagator of the linear ODE, the default;
 * 0 = the RK4 stages evaluated at every sub-step).  smcb_get_option returns INT64_MIN for an unknown name. */
int smcb_set_option(const char* name, int64_t value);
int64_t smcb_get_option(const char* name);
/* Text description of the Metropolis kernel instantiation the calling thread launched last
 * (template arguments, grid, RNG mode): lets a test assert WHICH variant it compared with the oracle. */
const char* smcb_last_mh_variant(void);

/* measurement aid: `blocks` x 256 threads each run 8 independent chains of `iters` FP64 FMAs
 * (16 * iters flops per thread); out holds blocks*256 doubles. */
int smcb_fp64_probe(double* out, int32_t blocks, int32_t iters, void* stream);
/* the same for the FP64 tensor-core path: every warp runs 16 * iters mma.sync.m8n8k4.f64 (512 flops each)
 * on 4 independent accumulator pairs; out holds blocks*256 doubles */
int smcb_dmma_probe(double* out, int32_t blocks, int32_t iters, void* stream);
/* out[i] = the streaming log-sum-exp loops' exp for x[i] <= 0 (table + degree-6 polynomial, ~1 ulp;
 * 0 below -708): accuracy probe for the tests, not part of the reference's path */
int smcb_exp_probe(const double* x, double* out, int64_t n, void* stream);

/* ---- layout / plumbing ----------------------------------------------------------- */
/* (N,d) row-major <-> SoA (d, stride); kernel_base.py:53-66 dict<->array conversions */
int smcb_aos_to_soa(const double* aos, double* soa, int64_t n, int32_t d, int64_t stride, void* stream);
int smcb_soa_to_aos(const double* soa, double* aos, int64_t n, int32_t d, int64_t stride, void* stream);
int smcb_fill_f64(double* dst, int64_t n, double value, void* stream);

/* ---- (a) tempered reweighting + LSE normalisation + ESS -------------------------- */
/* un_lw = log_w_in + [target(phi) - target(phi_prev)]   (paths.py:86-105, updater.py:122-133)
 * then normalise (particles.py:141-150), total_unnorm_log_weight (particles.py:81,200-205),
 * ESS (particles.py:158-162).  lq = proposal logpdf or NULL (-> prior, paths.py:107-112).
 * log_w_in NULL = uniform log(1/n_global).
 * scalars_out (device, 8 doubles): [0] max un_lw, [1] S1 = sum exp(un_lw-max),
 * [2] S2 = sum exp(2(un_lw-max)), [3] total_unnorm_log_weight, [4] ESS, [5] log S1. */
int smcb_reweight(int64_t n, const double* ll, const double* lp, const double* lq,
                  const double* log_w_in, int64_t n_global, const smcb_path_t* path,
                  double* log_w_out, double* w_out, double* scalars_out, void* ws, void* stream);
/* The same update in ONE cooperative launch (updater.py:97-133 incl. the resampling test :106):
 * partial sums -> grid barrier -> [world > 1: rank triples exchanged over NVLink peer memory inside
 * the kernel, x as for smcb_ess_search] -> scalars_out; log_w_out / w_out are written by the same
 * launch ONLY IF the step will not resample (ESS >= ess_threshold * n_global; ess_threshold < 0:
 * always) -- a resampling step takes its weights from smcb_cdf_scan_weights instead and never
 * materialises them.  lp may be NULL when the summed log-prior is one constant (lp_const).
 * scalars_out[6] = 1 if the weights were written, [7] = -1 if a peer never answered.
 * Algorithmic bytes per particle: 8 (ll) + 8 per optional input column, + 16 when weights are written. */
int smcb_reweight_fused(int64_t n, const double* ll, const double* lp, double lp_const, const double* lq,
                        const double* log_w_in, int64_t n_global, const smcb_path_t* path, double ess_threshold,
                        double* log_w_out, double* w_out, double* scalars_out, void* ws,
                        const smcb_xchg_t* x /* NULL: single GPU */, void* stream);
/* elementwise path values (host-facing GeometricPath wrappers): mode 0 = logpdf at phi
 * (paths.py:81-84), mode 1 = inc_log_weights phi_prev -> phi (paths.py:86-91) */
int smcb_path_eval(int64_t n, const double* ll, const double* lp, const double* lq,
                   const smcb_path_t* path, int32_t mode, double* out, void* stream);
/* multi-GPU split of the same: per-rank (max, S1, S2) -> [all-gather] -> merge -> finalize */
int smcb_reweight_partials(int64_t n, const double* ll, const double* lp, const double* lq,
                           const double* log_w_in, int64_t n_global, const smcb_path_t* path,
                           double* partial_out /*3*/, void* ws, void* stream);
int smcb_lse_merge(const double* partials /*n_parts x 3*/, int32_t n_parts, double* scalars_out,
                   void* stream);
int smcb_reweight_finalize(int64_t n, const double* ll, const double* lp, const double* lq,
                           const double* log_w_in, int64_t n_global, const smcb_path_t* path,
                           const double* scalars, double* log_w_out, double* w_out, void* stream);

/* ---- (b) adaptive phi: device bisection of ESS(phi) = target ---------------------- */
/* AdaptiveSampler.optimize_step / predict_ess_margin (samplers.py:250-301) with the
 * decision sequence of scipy.optimize.bisect (xtol 2e-12, rtol 8.88e-16, maxiter 100).
 * No host round-trip: the whole search is enqueued on `stream`.
 * state (device, 16 doubles): [0] result phi, [1] #function evals, [2] done flag,
 *   [3] xa, [4] dm, [5] fa, [6] current x, [7] last margin, [8] full-step margin.
 * lp may be NULL when the summed log-prior is the same for every particle (lp_const). */
int smcb_ess_bisect(int64_t n, const double* ll, const double* lp, const double* lq,
                    double lp_const, double phi_old, double lambda, double target_ess,
                    int64_t n_global, double* state, void* ws, void* stream);
/* The search as the samplers call it.  Same result as smcb_ess_bisect -- scipy's decision sequence --
 * in ~8-12 passes over the particles instead of ~43: ESS(phi) is non-increasing in phi, so the sign
 * scipy would compute at a midpoint is known without evaluating it whenever an evaluated point with a
 * clearly positive margin (> 1e-12 target N) lies above it, or one with a clearly negative margin
 * below it.  The kernel first closes such a bracket around the root (safeguarded Newton on log ESS in
 * log(phi - phi_old), started from hint_dphi = the previous step's phi increment when > 0, then two
 * probes either side of the converged estimate), then replays bisect.c midpoint by midpoint and
 * evaluates only the midpoints inside the bracket.  Paths with a proposal column (lq != NULL: the
 * increment is not one exponential family in phi) and option "bisect_plain" evaluate every midpoint.
 * x != NULL: multi-GPU form -- after every pass the ranks exchange their partial sums over NVLink
 * peer memory inside the one cooperative launch (x->peer_slots[r]: rank r's symmetric buffer of
 * smcb_xchg_slot_words() uint64, zero-initialised once; x->gen must increase with every call that uses
 * the buffer; global_counts / ticket unused).  state[2] = -1 if a peer never answered (time-out option
 * "xchg_timeout_s", default 30 s); state[10] = passes over the particles; state[11..12] = the bracket. */
int smcb_ess_search(int64_t n, const double* ll, const double* lp, const double* lq, double lp_const,
                    double phi_old, double lambda, double target_ess, int64_t n_global, double hint_dphi,
                    double* state, void* ws, const smcb_xchg_t* x, void* stream);
int smcb_xchg_slot_words(void);
/* TEST HOOKS (no product path calls them): the search's state machine -- the same C++ the kernel runs --
 * driven on the HOST with sums computed by plain CPU loops over a host array ll, so that its decisions
 * can be compared with scipy.optimize.bisect on smcb_test_search_margin without a GPU.
 * trace_out (optional, host): the phi evaluated in each pass. */
double smcb_test_search_margin(const double* ll_host, int64_t n, double phi_old, double x, double target);
int smcb_test_search_host(const double* ll_host, int64_t n, double phi_old, double target, double hint,
                          int32_t plain, double* state_out, double* trace_out, int32_t trace_len);
/* multi-GPU pieces: begin -> { partial -> [all-gather] -> update } x smcb_ess_bisect_max_iters */
int smcb_ess_bisect_begin(double phi_old, double* state, void* stream);
int smcb_ess_partial(int64_t n, const double* ll, const double* lp, const double* lq,
                     double lp_const, double phi_old, double lambda, const double* state,
                     double* partial_out /*3*/, void* ws, void* stream);
int smcb_ess_bisect_update(const double* partials, int32_t n_parts, double phi_old,
                           double target_ess, int64_t n_global, double* state, void* stream);
int smcb_ess_bisect_max_iters(void);
/* multi-GPU form of smcb_ess_bisect (replaces the MPI-free but collective-per-evaluation loop
 * smcb_ess_partial -> all-gather -> smcb_ess_bisect_update): ONE cooperative launch per rank; after
 * every evaluation the ranks exchange their (max, S1, S2) over NVLink peer memory inside the kernel.
 * x->peer_slots[r]: rank r's symmetric slot buffer of smcb_ess_bisect_xchg_slot_words() uint64
 * (zero-initialised once, mapped by every peer); x->gen must increase with every call;
 * x->global_counts / x->ticket are unused.  state[2] = -1 if a peer never answered. */
int smcb_ess_bisect_xchg_slot_words(void);
int smcb_ess_bisect_xchg(int64_t n, const double* ll, const double* lp, const double* lq,
                         double lp_const, double phi_old, double lambda, double target_ess,
                         int64_t n_global, double* state, void* ws, const smcb_xchg_t* x, void* stream);

/* ---- (c) resampling ---------------------------------------------------------------- */
/* u for `count` global slots starting at `first` out of n_global (resampler_rngs.py:4-9) */
int smcb_resample_uniforms(int32_t kind, int64_t n_global, int64_t first, int64_t count,
                           uint64_t seed, uint64_t stream_id, double* u_out, void* stream);
/* `standard` uniforms (resampler_rngs.py:4-5) in SORTED order by exponential spacings, for sharded
 * particles: U_(k) = S_k / S_(N+1), S = running sum of i.i.d. unit exponentials.
 *   1. smcb_resample_exponentials: e_out[i] = E / (2 N) of global slot first + i (Philox keyed by the
 *      slot; the scale keeps every running sum below 1 for the fixed-point scan carries);
 *   2. smcb_cdf_scan over e (local running sums t); the caller sums the rank totals in rank order;
 *   3. smcb_resample_spacings_finish: u_out[i] = (offset + t[i]) / (total + E_(N+1)), offset = sum of
 *      the lower ranks' totals, total = sum of all ranks' totals (device scalars).
 * Same distribution as sorting N i.i.d. uniforms; no sort, and routing becomes a contiguous split. */
int smcb_resample_exponentials(int64_t first, int64_t count, int64_t n_global, uint64_t seed,
                               uint64_t stream_id, double* e_out, void* stream);
int smcb_resample_spacings_finish(const double* t, int64_t n, const double* offset, const double* total,
                                  int64_t n_global, uint64_t seed, uint64_t stream_id, double* u_out,
                                  void* stream);
/* running sums of the same exponentials without the intermediate array: t_out[i] = sum_{j<=i} e_j over
 * this rank's `count` slots (steps 1 + 2 above in one scan whose inputs are generated in registers) */
int smcb_resample_exponential_sums(int64_t first, int64_t count, int64_t n_global, uint64_t seed,
                                   uint64_t stream_id, double* t_out, double* total_out /* or NULL: t_out[count-1] */,
                                   void* ws, void* stream);
/* CDF of the normalised weights of a step that resamples, straight from the resident columns: the
 * look-back scan recomputes w_i = exp((log_w_in_i + inc_i - max) - log S1) (arguments as for
 * smcb_reweight_fused, scalars = its scalars_out) instead of reading a weight array
 * (particles.py:141-150 + np.cumsum, updater.py:152) */
int smcb_cdf_scan_weights(int64_t n, const double* ll, const double* lp, double lp_const, const double* lq,
                          const double* log_w_in, int64_t n_global, const smcb_path_t* path,
                          const double* scalars, double* cdf, double* total_out /* or NULL: cdf[n-1] */, void* ws,
                          void* stream);
/* Resampling with sorted uniforms in one pass (updater.py:135-165 for the native generator): uniforms of
 * the global output slots evaluated on the fly (resampler_rngs.py:4-9; STANDARD: U_(k) = S_k / S_(N+1)
 * from the running sums t), ancestor search confined to the CDF window of each 1024-slot tile
 * (merge-path style partition), gather of all n_cols columns, distinct-ancestor count -- no uniform,
 * index or flag arrays.  Sharded particles (world > 1, equal contiguous blocks): this rank OWNS the
 * slots whose uniform falls into its CDF segment [O_r, O_r+1) -- a contiguous slot range it derives from
 * `totals` -- and writes the selected rows straight into the SoA buffer of the rank holding each slot
 * (dst_peer[q], peer-mapped over NVLink); STANDARD reads the running sums of the slot's rank through
 * t_peer[q].  The caller all-gathers `totals` before and synchronises the ranks after. */
typedef struct smcb_resample {
    int32_t kind;                           /* SMCB_RESAMPLE_* */
    int32_t world, rank, n_cols;
    int64_t n_local, n_global;              /* n_global == world * n_local */
    uint64_t seed, stream_id;
    const double* cdf;                      /* local inclusive scan of the globally normalised weights */
    const double* totals;                   /* device, world x 2, rank order: (cdf[n_local-1], t[n_local-1]); may be
                                               NULL when world == 1 */
    const double* t_peer[SMCB_MAX_RANKS];   /* STANDARD: every rank's running sums of the exponentials */
    const double* src;                      /* local SoA columns */
    int64_t src_stride;
    double* dst_peer[SMCB_MAX_RANKS];       /* every rank's destination SoA buffer */
    int64_t dst_stride;
    int64_t* idx_out;                       /* optional, world == 1: ancestor of every slot */
    int64_t* n_unique_out;                  /* device: distinct local ancestors selected by this rank */
} smcb_resample_t;
int smcb_resample_fused(const smcb_resample_t* r, void* ws, void* stream);
/* inclusive prefix sum of normalised weights -> CDF (np.cumsum, updater.py:152) */
int smcb_cdf_scan(const double* w, double* cdf, int64_t n, int32_t mode, void* ws, void* stream);
/* idx_k = #{j : cdf_j <= u_k}  (np.digitize / searchsorted side='right', updater.py:152),
 * clamped to n_cdf-1 (documented divergence from the reference's IndexError, quirk Q4) */
int smcb_searchsorted(const double* cdf, int64_t n_cdf, const double* u, int64_t n_u,
                      int64_t* idx_out, void* stream);
/* same against a shard of a global CDF: idx_k = #{j : *offset + cdf_j <= u_k} (multi-GPU:
 * offset = sum of the weight totals of the lower ranks, device scalar) */
int smcb_searchsorted_offset(const double* cdf, int64_t n_cdf, const double* offset, const double* u,
                             int64_t n_u, int64_t* idx_out, void* stream);
/* dst[c][k] = src[c][idx[k]] for n_cols SoA columns (updater.py:140-143) */
int smcb_gather(const double* src, int64_t src_stride, double* dst, int64_t dst_stride,
                const int64_t* idx, int64_t n_out, int32_t n_cols, void* stream);
/* fused gather + all-to-all over NVLink peer memory (multi-GPU resampling): request j of this
 * rank's n_req incoming requests came from rank q = the block with block_off[q] <= j <
 * block_off[q+1]; the selected row idx[j] is written column by column straight into THAT rank's
 * SoA buffer dst[q] (a peer-mapped pointer) at row slot_off[q] + (j - block_off[q]).
 * block_off (world+1) and slot_off (world) are device int64 arrays, dst is a HOST array of
 * `world` device pointers.  The caller synchronises the ranks afterwards. */
int smcb_gather_p2p(const double* src, int64_t src_stride, const int64_t* idx, int64_t n_req, int32_t n_cols,
                    int32_t world, const int64_t* block_off, const int64_t* slot_off, double* const* dst,
                    int64_t dst_stride, void* stream);
/* number of distinct ancestors (updater.py:154-165); flags = n_src bytes of scratch */
int smcb_count_unique(const int64_t* idx, int64_t n_idx, int64_t n_src, int64_t* count_out,
                      void* ws, void* stream);

/* ---- small rank records over NVLink peer memory (multi-GPU plumbing of the path) ---------------- */
/* All-gather (reduce = 0: dst[r * n_vals + k] = rank r's src[k]) or all-reduce (reduce = 1: dst[k] = sum over
 * ranks in rank order, bit-identical on every rank) of n_vals <= smcb_xchg_max_vals() doubles in ONE
 * single-block kernel: every rank stores its record into a slot of every peer's symmetric buffer
 * (x->peer_slots[r]: smcb_xchg_big_slot_words() uint64 each, zero-initialised once), fences, stores the
 * flag x->gen (must increase by one with every call on this buffer) and waits for the `world` flags in
 * its own memory.  Replaces the MPI allgather / bcast of the reference's parallel path
 * (smcpy/mcmc/parallel_vector_mcmc.py:36-47, smcpy/utils/mpi_utils.py) and the small NCCL collectives for
 * weight totals, covariance partial sums, counts and error flags.  *status |= 1 on time-out. */
int smcb_xchg_big_slot_words(void);
int smcb_xchg_max_vals(void);
int smcb_xchg_allgather(const smcb_xchg_t* x, const double* src, int32_t n_vals, double* dst, int32_t reduce,
                        int32_t* status, void* stream);

/* ---- weighted mean / covariance (particles.py:189-198, np.cov ddof=0 aweights) ---- */
/* mean_out[d], cov_out[d*d] (device).  sums_out (device, 1+d+d*d doubles) receives the raw
 * partial sums [sum w, sum w x, sum w (x-m)(x-m)^T] for the multi-GPU all-reduce. */
int smcb_weighted_cov(const double* params, int64_t stride, const double* w, int64_t n, int32_t d,
                      double* mean_out, double* cov_out, double* sums_out, void* ws, void* stream);
/* the same for particles sharded over several GPUs, enqueued by one call: both moment sums are all-reduced over
 * the ranks through smcb_xchg_allgather on generations x->gen and x->gen + 1 (the caller advances its counter by
 * two); xchg_scratch: device, max(1 + d, d * d) doubles; *status |= 1 on time-out.  Replaces np.cov over the
 * gathered particles of the reference's parallel path (particles.py:189-198). */
int smcb_weighted_cov_xchg(const double* params, int64_t stride, const double* w, int64_t n, int32_t d,
                           double* mean_out, double* cov_out, double* sums_out, void* ws, const smcb_xchg_t* x,
                           double* xchg_scratch, int32_t* status, void* stream);
int smcb_weighted_sums1(const double* params, int64_t stride, const double* w, int64_t n, int32_t d,
                        double* sums_out /*1+d*/, void* ws, void* stream);
int smcb_mean_from_sums(const double* sums /*1+d*/, int32_t d, double* mean_out, void* stream);
int smcb_weighted_sums2(const double* params, int64_t stride, const double* w, int64_t n, int32_t d,
                        const double* mean, double* sums_out /*d*d*/, void* ws, void* stream);
int smcb_cov_from_sums(const double* wsum /*1*/, const double* sums2 /*d*d*/, int32_t d,
                       double* cov_out, void* stream);
/* the same covariance (particles.py:189-198) in ONE pass over the particles, for d > 8: moments about
 * a shift c close to the mean (the previous step's mean).  sums_out = [W, S1 (d), S2 (d x d)] with
 * S1 = sum w (x - c), S2 = sum w (x - c)(x - c)^T (all-reduce them across ranks), then
 * smcb_cov_from_shifted_sums: mean = c + S1 / W, cov = S2 / W - (S1 / W)(S1 / W)^T, c <- mean; sets
 * SMCB_FLAG_COV_RESHIFT when |mean - c|^2 > 1e4 var for some parameter (cancellation: the caller then
 * recomputes with the two-pass entry points).  Returns SMCB_ERR_UNSUPPORTED for d <= 8, unaligned
 * columns or fewer than 148 x 128 particles. */
int smcb_weighted_moments_shifted(const double* params, int64_t stride, const double* w, int64_t n,
                                  int32_t d, const double* shift, double* sums_out, void* ws, void* stream);
int smcb_cov_from_shifted_sums(const double* sums, double* shift, int32_t d, double* mean_out,
                               double* cov_out, int32_t* flags, void* stream);
/* lower Cholesky factor (np.linalg.cholesky, vector_mcmc.py:225); sets SMCB_FLAG_CHOL_FAIL */
int smcb_cholesky(const double* cov, int32_t d, double* chol_out, int32_t* flags, void* stream);

/* ---- (d) vectorised Metropolis ------------------------------------------------------ */
/* Linear model + Normal likelihood (README.md:64-79, smcpy/log_likelihoods.py:31-49): the kernels accumulate the
 * squared residuals about the data means -- sum_j (a x_j + b - y_j)^2 = sum_j (a u_j - v_j)^2 + M (a xbar + b - ybar)^2
 * with u = x - xbar, v = y - ybar -- two FP64 instructions per data point instead of three, still one pass over
 * the whole data vector per particle.  out (device, 2 + 2 M doubles) = xbar, ybar, (u_j, -v_j) pairs; set
 * smcb_problem_t::lin_centered to it. */
int smcb_linear_center(const double* xgrid, const double* data, int32_t n_data, double* out, void* stream);
/* MVNormal (smcpy/log_likelihoods.py:150-192): the model output is the same for every snapshot, so the
 * kernels need the S x F data only through its mean and scatter matrix.  stats_out (device,
 * SMCB_MVN_STATS doubles) = mean[3], sum_s (y_s - mean)(y_s - mean)^T [3][3] (zero padded for F < 3);
 * set smcb_problem_t::mvn_stats to it.  data: device, S x F row-major, F <= 3. */
int smcb_mvn_scatter(const double* data, int32_t n_snapshots, int32_t n_features, double* stats_out, void* stream);
/* log-priors + log-likelihood of incoming particles
 * (vector_mcmc.py:162-166 _initialize_probabilities, initializer.py:66-73).
 * lp_cols_out (optional, n_priors x stride) = un-summed log priors (vector_mcmc.py:98-111).
 * lq_out (required iff prob->has_proposal) = proposal log-pdf (paths.py:107-112). */
int smcb_mh_init(const smcb_problem_t* prob, const double* params, int64_t stride, int64_t n,
                 double* ll_out, double* lp_out, double* lp_cols_out, double* lq_out, int32_t* flags,
                 void* stream);
/* log-priors only: lp_out[n] summed (vector_mcmc_kernel.py:33-36), lp_cols_out optional
 * (n_priors x stride) un-summed (vector_mcmc.py:98-111) */
int smcb_log_priors(const smcb_problem_t* prob, const double* params, int64_t stride, int64_t n,
                    double* lp_out, double* lp_cols_out, void* stream);
/* one fused smc_metropolis step (vector_mcmc.py:168-183 + :55-60): Gaussian random-walk
 * proposal x + sqrt(s) L z, priors, model, likelihood, tempered ratio, accept/reject, in place.
 * accept_counts (device, int64[SMCB_MAX_MCMC_STEPS]): slot `step` receives this step's number
 * of accepted particles; slots < step (already all-reduced across GPUs by the caller) give
 * the covariance scale s (x1/5 below 30 %, x2 above 70 %). `moved` (n bytes) |= accepted.
 * lq (required iff prob->has_proposal): resident proposal log-pdf column, updated in place.
 * flags: device int32[4], zero-initialised once: [0] SMCB_FLAG_* bits, [1] and [2] are the
 * library's dynamic tile counter and ticket (left at zero by every launch). */
int smcb_mh_step(const smcb_problem_t* prob, const smcb_path_t* path, double* params, int64_t stride,
                 int64_t n, int64_t n_global, double* ll, double* lp, double* lq, uint8_t* moved,
                 const double* chol, int64_t* accept_counts, int32_t step, const smcb_rng_t* rng,
                 int32_t* flags, void* stream);
/* same step with the cross-GPU accept-count exchange fused in (x->world > 1); accept_counts
 * then holds the LOCAL counts, x->global_counts decision-equivalent global ones (see smcb_xchg_t). */
int smcb_mh_xchg_slot_words(void);
int smcb_mh_step_xchg(const smcb_problem_t* prob, const smcb_path_t* path, double* params, int64_t stride,
                      int64_t n, int64_t n_global, double* ll, double* lp, double* lq, uint8_t* moved,
                      const double* chol, int64_t* accept_counts, int32_t step, const smcb_rng_t* rng,
                      int32_t* flags, const smcb_xchg_t* x, void* stream);
/* `n_steps` consecutive fused Metropolis steps (first_step .. first_step + n_steps - 1 of the mutate call, generator
 * streams rng->stream .. rng->stream + n_steps - 1) enqueued by ONE call: the K-step loop of
 * VectorMCMC.smc_metropolis (vector_mcmc.py:52-60) without a host round trip per step -- what makes the small
 * configurations (C1: 500 particles) launch-bound is the per-step binding overhead, not the kernels.  The launches
 * of steps after the first are chained by programmatic dependent launch (the next grid is scheduled while the
 * previous one drains and stages the constant problem data before its griddepcontrol.wait; option "no_pdl" = 1
 * serialises them plainly; results are bit-identical).  Native generator only; x != NULL && x->world > 1 selects
 * the multi-GPU form. */
int smcb_mh_steps(const smcb_problem_t* prob, const smcb_path_t* path, double* params, int64_t stride, int64_t n,
                  int64_t n_global, double* ll, double* lp, double* lq, uint8_t* moved, const double* chol,
                  int64_t* accept_counts, int32_t first_step, int32_t n_steps, const smcb_rng_t* rng, int32_t* flags,
                  const smcb_xchg_t* x, void* stream);
/* fraction of particles that moved (mutator.py:69-71): count of non-zero bytes */
int smcb_count_moved(const uint8_t* moved, int64_t n, int64_t* count_out, void* ws, void* stream);
/* initial particles from uniform priors on the device (vector_mcmc.py:91-96) */
int smcb_sample_priors(const smcb_problem_t* prob, double* params, int64_t stride, int64_t n,
                       uint64_t seed, uint64_t stream_id, int64_t id_offset, void* stream);

/* un-fused pieces for user models given as torch callables on device tensors:
 * propose -> [user model] -> row-wise Normal log-likelihood -> accept */
int smcb_mh_propose(const smcb_problem_t* prob, const double* params, int64_t stride, int64_t n,
                    int64_t n_global, const double* chol, const int64_t* accept_counts, int32_t step,
                    const smcb_rng_t* rng, double* new_params /*SoA*/, double* new_lp, void* stream);
int smcb_normal_loglike_rows(const double* outputs /*(n, m) row-major*/, int64_t n, int32_t m,
                             const double* data, const double* sigma_col /*or NULL*/, double sigma,
                             const double* lp /*or NULL: rows with -inf keep 0*/, double* ll_out,
                             int32_t* flags, void* stream);
int smcb_mh_accept(const smcb_path_t* path, double* params, const double* new_params, int64_t stride,
                   int32_t d, int64_t n, double* ll, const double* new_ll, double* lp,
                   const double* new_lp, uint8_t* moved, int64_t* accept_counts, int32_t step,
                   const smcb_rng_t* rng, void* stream);

/* user model registered as CUDA source (north_star: "a registered __device__ model through the C-ABI").
 * `source` defines
 *     __device__ double smcb_user_model(const double* theta, int j, double x, double* state)
 * = model output at data point j for the model parameters theta[0..n_model_params) (x = xgrid[j] or 0;
 * state = 8 doubles per particle, zero at j = 0, persisting over j).  It replaces the user's
 * `model(inputs)` callable of VectorMCMC.evaluate_model (smcpy/mcmc/vector_mcmc.py:88-89); NVRTC
 * compiles it for sm_100a with a kernel that also evaluates the Normal log-likelihood
 * (smcpy/log_likelihoods.py:22-49), so model outputs never reach HBM.  Host pointers; handle >= 1. */
int smcb_model_register_nvrtc(const char* source, int32_t n_model_params, int32_t* handle_out);
/* ll_out[i] = Normal log-likelihood of the registered model at SoA particle i (replaces
 * "user model + smcb_normal_loglike_rows" in the un-fused route above) */
int smcb_user_loglike(int32_t handle, const double* params, int64_t stride, int64_t n, int32_t m,
                      const double* data, const double* xgrid /*or NULL*/, const double* sigma_col /*or NULL*/,
                      double sigma, const double* lp /*or NULL: rows with -inf keep 0*/, double* ll_out,
                      int32_t* flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMCB200_H */
