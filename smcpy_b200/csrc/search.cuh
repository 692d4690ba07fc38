// search.cuh -- the adaptive phi search's state machine (kernel (b); smcpy/smc/samplers.py:250-276 +
// scipy/optimize/Zeros/bisect.c).  Plain C++ shared by the cooperative kernel (one identical copy per
// block, advanced by thread 0) and by smcb_test_search_host, a host-side TEST HOOK that drives the same
// code with sums computed on the CPU so that its decisions can be compared with scipy.optimize.bisect
// without a GPU.  No product path calls the host form.
#pragma once

#include <math.h>

#include "common.cuh"

#define SMCB_HD __host__ __device__
#ifdef __CUDA_ARCH__
#define SMCB_NOINLINE __noinline__
#else
#define SMCB_NOINLINE
#endif

namespace smcb {

constexpr double kXtol = 2e-12;
constexpr double kRtol = 8.881784197001252e-16;
constexpr int kBisectMaxIters = 48;   // (1-0)/2^41 < 2e-12: scipy stops after <= 40 halvings
constexpr int kZoneMaxEvals = 16;     // bracket-closing evaluations before falling back to plain midpoints
constexpr int kSearchMaxPasses = kBisectMaxIters + kZoneMaxEvals + 4;

// (max, S1, S2) plus the derivative sums T1 = sum s e^(v-max), T2 = sum s e^(2(v-max)), s = d v / d phi
struct Lse5 {
    double m, s1, s2, t1, t2;
};

SMCB_HD inline double margin_from(const Lse& a, double n_global, double target) {
    // predict_ess_margin (samplers.py:264-276): ESS = exp(2 LSE(b) - LSE(2b)) = S1^2/S2,
    // ESS = 0 when either log-sum is -inf
    double ess = 0.0;
    if (a.m != -__builtin_huge_val() && a.s1 > 0.0 && a.s2 > 0.0) ess = (a.s1 * a.s1) / a.s2;
    return ess - n_global * target;
}


// Accelerated form.  Search state (one identical copy per block, advanced by thread 0):
struct Search {
    double phi_old, target, target_n, n_global, tol_f, hint;
    // scipy/optimize/Zeros/bisect.c replay: bracket [xa, xa + dm], fa = f(xa) > 0, pending midpoint xm
    double xa, dm, fa, xm;
    double result, f_full, f_last;
    int done, scipy_evals, passes;
    // points with a CLEARLY signed margin (|f| > tol_f): ps = largest x with f > 0, qs = smallest with f < 0
    double ps, qs;
    int ps_ok, qs_ok;
    // points with any sign (bracket of the Newton iteration)
    double pa, qa;
    // latest bracket-closing evaluation: x, g = log(ESS / (target N)), dg/dx
    double xl, gl, dl;
    int zone_evals;
    int phase;            // 0 full step, 1 Newton, 2 low probe pending, 3 high probe pending, 4 midpoints only
    double d_star, eta;   // converged root estimate (as phi - phi_old) and probe half-width
    double x_eval;
    int eval_is_mid;
    int plain;
};

SMCB_HD inline bool search_term(const Search& s) { return fabs(s.dm) < kXtol + kRtol * fabs(s.xm); }

// next point for the bracket-closing iteration, or NaN if it has nothing useful to offer
SMCB_HD inline double search_zone_candidate(Search& s) {
    const double dlo = s.pa - s.phi_old, dhi = s.qa - s.phi_old;
    double cand = nan("");
    const double dl = s.xl - s.phi_old;
    const double gp = s.dl * dl;                          // d g / d log(delta)
    if (s.dl < 0.0 && gp < 0.0 && isfinite(s.gl) && dl > 0.0) {
        double step = -s.gl / gp;                         // Newton on g in log(delta)
        if (fabs(step) <= 1e-9) {                         // converged: close the bracket with two probes
            s.d_star = dl * exp(step);
            double eta = 4e-12 / fabs(gp);                // margin there ~ ESS * |gp| * eta = 4e-12 ESS > tol_f
            eta = fmin(fmax(eta, 4e-12), 1e-3);
            s.eta = eta;
            s.phase = 2;
            return s.phi_old + s.d_star * (1.0 - eta);
        }
        step = fmin(fmax(step, -2.5), 2.5);
        cand = s.phi_old + dl * exp(step);
    }
    if (!(cand > s.pa && cand < s.qa)) {                  // outside the bracket (or no usable slope): split it
        if (!(dlo > 0.0)) cand = s.phi_old + dhi * 0.0625;
        else if (dhi > 4.0 * dlo) cand = s.phi_old + sqrt(dlo * dhi);
        else cand = s.phi_old + 0.5 * (dlo + dhi);
    }
    if (!(cand > s.pa && cand < s.qa)) return nan("");
    return cand;
}

SMCB_HD inline void search_init(Search& s, double phi_old, double target, double n_global, double hint, int plain) {
    s.phi_old = phi_old;
    s.n_global = n_global;
    s.target = target;
    s.target_n = n_global * target;
    s.tol_f = 1e-12 * s.target_n;
    s.hint = hint;
    s.xa = phi_old;
    s.dm = 1.0 - phi_old;
    s.fa = 0.0;
    s.xm = 1.0;
    s.result = 1.0;
    s.f_full = s.f_last = 0.0;
    s.done = 0;
    s.scipy_evals = 0;
    s.passes = 0;
    s.ps = s.pa = phi_old;
    s.qs = s.qa = 1.0;
    s.ps_ok = s.qs_ok = 0;
    s.xl = s.gl = s.dl = 0.0;
    s.zone_evals = 0;
    s.phase = 0;
    s.d_star = s.eta = 0.0;
    s.x_eval = 1.0;
    s.eval_is_mid = 0;
    s.plain = plain;
}

// consumes the sums of the evaluation at s.x_eval and sets the next s.x_eval (or s.done)
SMCB_HD SMCB_NOINLINE void search_advance(Search& s, const Lse5& a) {
    if (s.done) return;
    const double f = margin_from(Lse{a.m, a.s1, a.s2}, s.n_global, s.target);
    const double x = s.x_eval;
    s.f_last = f;
    s.passes += 1;
    // g = log(ESS / (target N)), dg/dx = 2 (T1/S1 - T2/S2)
    double g = -__builtin_huge_val(), dg = 0.0;
    if (a.m != -__builtin_huge_val() && a.s1 > 0.0 && a.s2 > 0.0) {
        g = log(((a.s1 * a.s1) / a.s2) / s.target_n);
        dg = 2.0 * (a.t1 / a.s1 - a.t2 / a.s2);
    }
    if (s.phase == 0) {
        // AdaptiveSampler.optimize_step: full step first (samplers.py:251-252), then scipy evaluates f(a), f(b)
        s.f_full = f;
        s.scipy_evals = 1;
        if (f > 0.0) { s.result = 1.0; s.done = 1; return; }
        s.fa = s.n_global - s.target_n;               // delta_phi = 0: ESS = N exactly (samplers.py:266-267)
        s.scipy_evals = 3;
        if (s.fa == 0.0) { s.result = s.phi_old; s.done = 1; return; }
        if (f == 0.0) { s.result = 1.0; s.done = 1; return; }
        s.ps_ok = s.fa > s.tol_f;
        s.qs_ok = f < -s.tol_f;
        s.xl = 1.0;
        s.gl = g;
        s.dl = dg;
        s.phase = s.plain ? 4 : 1;
        s.dm *= 0.5;                                  // first midpoint
        s.xm = s.xa + s.dm;
    } else {
        if (f > 0.0) { if (x > s.pa) s.pa = x; } else if (f < 0.0) { if (x < s.qa) s.qa = x; }
        if (f > s.tol_f) { if (!s.ps_ok || x > s.ps) { s.ps = x; s.ps_ok = 1; } }
        if (f < -s.tol_f) { if (!s.qs_ok || x < s.qs) { s.qs = x; s.qs_ok = 1; } }
        if (s.eval_is_mid) {
            s.scipy_evals += 1;
            if (f * s.fa >= 0.0) s.xa = s.xm;
            if (f == 0.0 || search_term(s)) { s.result = s.xm; s.done = 1; return; }
            s.dm *= 0.5;
            s.xm = s.xa + s.dm;
        } else {
            s.xl = x;
            s.gl = g;
            s.dl = dg;
            s.zone_evals += 1;
        }
    }
    // replay scipy's decisions for every midpoint whose sign is known from monotonicity
    while (true) {
        int sign = 0;
        if (s.ps_ok && s.xm <= s.ps) sign = 1;
        else if (s.qs_ok && s.xm >= s.qs) sign = -1;
        if (sign == 0) break;
        s.scipy_evals += 1;
        if (sign > 0) s.xa = s.xm;                    // f * fa >= 0  (fa > 0)
        if (search_term(s)) { s.result = s.xm; s.done = 1; return; }
        s.dm *= 0.5;
        s.xm = s.xa + s.dm;
    }
    // the pending midpoint lies inside the bracket: either tighten the bracket or evaluate it
    double cand = nan("");
    if (s.phase == 1) {
        if (s.zone_evals >= kZoneMaxEvals) s.phase = 4;
        else if (s.zone_evals == 0 && s.hint > 0.0 && s.phi_old + 1.3 * s.hint < s.qa) cand = s.phi_old + 1.3 * s.hint;
        else cand = search_zone_candidate(s);         // may move to phase 2
        if (s.phase == 1 && !(cand == cand)) s.phase = 4;
    } else if (s.phase == 2) {                        // low probe done
        if (f > s.tol_f) {                            // as expected: now the high probe
            s.phase = 3;
            cand = s.phi_old + s.d_star * (1.0 + s.eta);
            if (!(cand < 1.0)) { cand = nan(""); s.phase = 4; }
        } else if (s.zone_evals < kZoneMaxEvals) {    // the estimate was off: iterate again from the probe
            s.phase = 1;
            cand = search_zone_candidate(s);
            if (!(cand == cand)) s.phase = 4;
        } else {
            s.phase = 4;
        }
    } else if (s.phase == 3) {
        if (f < -s.tol_f || s.zone_evals >= kZoneMaxEvals) {
            s.phase = 4;
        } else {
            s.phase = 1;
            cand = search_zone_candidate(s);
            if (!(cand == cand)) s.phase = 4;
        }
    }
    if (cand == cand && cand > s.phi_old && cand < 1.0 && s.phase != 4) {
        s.x_eval = cand;
        s.eval_is_mid = 0;
    } else {
        s.x_eval = s.xm;
        s.eval_is_mid = 1;
    }
}

SMCB_HD inline void search_export(const Search& s, double* st_out, int status) {
    st_out[0] = s.result;
    st_out[1] = (double)s.scipy_evals;
    st_out[2] = (status < 0) ? -1.0 : (s.done ? 1.0 : 0.0);
    st_out[3] = s.xa;
    st_out[4] = s.dm;
    st_out[5] = s.fa;
    st_out[6] = s.x_eval;
    st_out[7] = s.f_last;
    st_out[8] = s.f_full;
    st_out[9] = (double)s.phase;
    st_out[10] = (double)s.passes;
    st_out[11] = s.ps_ok ? s.ps : nan("");
    st_out[12] = s.qs_ok ? s.qs : nan("");
}

}  // namespace smcb
