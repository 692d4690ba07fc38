"""
Small, seeded versions of the BASELINE.json configs (C1..C5) used by the golden
fixtures, the oracle tests and the GPU parity tests.  Pure data: builds the
oracle-style problem dict (see oracle/smc_oracle.py) plus sampler arguments.
Nothing here reads /root/reference.
"""

import numpy as np

TRUE_COV = np.array([[0.5, 0.25, 0.005], [0.25, 0.25, 0.04], [0.005, 0.04, 1.0]])


def noisy(clean, std, seed):
    """Same draw as smcpy/utils/noise_generator.py:4-10 (normal(0,std,data.T.shape).T)."""
    clean = np.atleast_2d(clean)
    return np.random.default_rng(seed).normal(0, std, clean.T.shape).T + clean


def linear_problem(m, sigma=0.5, data_seed=0, unknown_sigma=False):
    x = np.linspace(0.5, 2.5, m)
    data = noisy(2.0 * x + 3.5, 0.5, data_seed)[0]
    priors = [("uniform", -6.0, 12.0), ("uniform", -6.0, 12.0)]
    if unknown_sigma:
        priors.append(("uniform", 0.0, 10.0))
    return {
        "model": ("linear", x),
        "data": data,
        "loglike": ("normal", None if unknown_sigma else sigma),
        "priors": priors,
        "names": ["a", "b"] + (["std"] if unknown_sigma else []),
    }


def spring_problem(n_out=100, nsub=4, data_seed=3):
    t_end = 5.0
    dt = t_end / (n_out - 1)
    args = dict(x0=0.0, v0=0.0, t0=0.0, dt=dt, n_out=n_out, nsub=nsub)
    t = np.arange(n_out) * dt
    K, g = 1.67, 4.62
    clean = (g / K) * (1.0 - np.cos(np.sqrt(K) * t))
    data = noisy(clean, 0.5, data_seed)[0]
    return {
        "model": ("spring_mass", args),
        "data": data,
        "loglike": ("normal", None),
        "priors": [("uniform", 0.0, 10.0)] * 3,
        "names": ["K", "g", "std"],
    }


def mvn_problem(n_snap=50, data_seed=200):
    X = np.arange(3, dtype=float)
    rs = np.random.RandomState(data_seed)
    data = np.tile(2.0 * X + 3.5, (n_snap, 1)) + rs.multivariate_normal([0, 0, 0], TRUE_COV, n_snap)
    return {
        "model": ("linear_features", X),
        "data": data,
        "loglike": ("mvnormal", [None] * 6),
        "priors": [("uniform", 0.0, 6.0), ("uniform", 0.0, 6.0), ("improper_cov", 3)],
        "names": ["a", "b"] + ["cov%d" % i for i in range(6)],
    }


def gauss_problem(d, data_seed=5):
    data = np.random.default_rng(data_seed).normal(0, 1, d)
    return {
        "model": ("identity", d),
        "data": data,
        "loglike": ("normal", 1.0),
        "priors": [("uniform", -10.0, 20.0)] * d,
        "names": ["x%d" % i for i in range(d)],
    }


def proposal_problem(required_phi=1):
    """C1 with an initial-particle proposal q = N(2,1) x N(3.5,1.5) (examples/proposal_example)."""
    p = linear_problem(100)
    p["proposal"] = [("normal", 2.0, 1.0), ("normal", 3.5, 1.5)]
    p["required_phi"] = required_phi
    return p


def multisource_problem(data_seed=12):
    """Linear model, three data sources with different noise levels; the middle one's std-dev is
    sampled (examples/likelihood_functions/multisource_example)."""
    x = np.linspace(0.5, 2.5, 90)
    rng = np.random.default_rng(data_seed)
    lens, true_std = (40, 20, 30), (0.3, 1.0, 0.6)
    noise = np.concatenate([rng.normal(0, s, n) for n, s in zip(lens, true_std)])
    return {
        "model": ("linear", x),
        "data": 2.0 * x + 3.5 + noise,
        "loglike": ("multisource", [lens, (0.3, None, 0.6)]),
        "priors": [("uniform", -6.0, 12.0), ("uniform", -6.0, 12.0), ("uniform", 0.0, 5.0)],
        "names": ["a", "b", "std1"],
    }


def invwishart_problem(n_snap=50):
    """C4 with a proper InvWishart(dof=5, scale=I) prior on the covariance terms (smcpy/priors.py:38-90)
    instead of ImproperCov."""
    p = mvn_problem(n_snap)
    p["priors"] = [("uniform", 0.0, 6.0), ("uniform", 0.0, 6.0), ("invwishart", 3, 5, np.eye(3))]
    return p


def sum_below(x):
    """Constraint of the constrained-prior case: a + b < 5.6 (binds: the posterior sits at a + b ~ 5.52)."""
    return x[:, 0] + x[:, 1] < 5.6


def constrained_problem():
    """C1 under one 2-column ImproperConstrainedUniform prior (smcpy/priors.py:150-247,
    examples/priors/run_example_constrained_prior.py)."""
    p = linear_problem(100)
    p["priors"] = [("constrained", 2, sum_below, np.array([[-6.0, -6.0], [6.0, 6.0]]))]
    return p


def impcov_normal_problem():
    """Linear model, Normal likelihood with the std-dev sampled, its prior a 1 x 1 ImproperCov (positive)."""
    p = linear_problem(60, data_seed=8, unknown_sigma=True)
    p["priors"] = [("uniform", -6.0, 12.0), ("uniform", -6.0, 12.0), ("improper_cov", 1)]
    return p


def constrained_init_params(n, seed):
    """ImproperConstrainedUniform.rvs (smcpy/priors.py:199-212): uniform candidates inside the bounds, pruned by
    the constraint, repeated until n are found."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        cand = rng.uniform([-6.0, -6.0], [6.0, 6.0], (n, 2))
        out += list(cand[sum_below(cand)])
    return np.array(out)[:n]


def impcov_normal_init_params(n, seed):
    rng = np.random.default_rng(seed)
    return np.column_stack([rng.uniform(-6, 6, n), rng.uniform(-6, 6, n), rng.uniform(0.05, 5.0, n)])


# name -> (problem builder, sampler kind, sampler kwargs, rng seed)
CASES = {
    "c1_adaptive": (lambda: linear_problem(100), "adaptive",
                    dict(n=500, K=10, target_ess=0.8, resampler="standard"), 1),
    "c2_fixed_strat": (lambda: linear_problem(200, data_seed=7), "fixed",
                       dict(n=600, K=5, phi=np.linspace(0, 1, 12), ess_threshold=0.8,
                            resampler="stratified"), 2),
    "c2_fixed_unknown_std": (lambda: linear_problem(60, data_seed=8, unknown_sigma=True), "fixed",
                             dict(n=400, K=6, phi=np.linspace(0, 1, 15) ** 2, ess_threshold=0.6,
                                  resampler="standard"), 3),
    "c3_spring": (lambda: spring_problem(100, 4), "adaptive",
                  dict(n=300, K=3, target_ess=0.8, resampler="standard"), 4),
    "c4_mvn": (lambda: mvn_problem(50), "adaptive",
               dict(n=300, K=5, target_ess=0.8, resampler="standard", iw_seed=11), 5),
    "c5_gauss8": (lambda: gauss_problem(8), "adaptive",
                  dict(n=400, K=4, target_ess=0.8, resampler="stratified"), 6),
    "c5_gauss32": (lambda: gauss_problem(32), "adaptive",
                   dict(n=200, K=3, target_ess=0.85, resampler="standard"), 7),
    "c2_multisource": (lambda: multisource_problem(), "fixed",
                       dict(n=400, K=5, phi=np.linspace(0, 1, 10) ** 2, ess_threshold=0.7,
                            resampler="standard"), 10),
    "c1_proposal": (lambda: proposal_problem(), "adaptive",
                    dict(n=400, K=5, target_ess=0.8, resampler="standard"), 8),
    "c1_proposal_reqphi": (lambda: proposal_problem([0.2, 0.6]), "adaptive",
                           dict(n=400, K=4, target_ess=0.75, resampler="stratified"), 9),
    # priors without a device log-pdf (host-evaluated column of the un-fused route)
    "c4_invwishart": (lambda: invwishart_problem(50), "adaptive",
                      dict(n=200, K=3, target_ess=0.8, resampler="standard", iw_seed=13), 11),
    "c1_constrained": (lambda: constrained_problem(), "adaptive",
                       dict(n=300, K=4, target_ess=0.8, resampler="standard", init="constrained", init_seed=14), 12),
    "c2_impcov_normal": (lambda: impcov_normal_problem(), "fixed",
                         dict(n=300, K=4, phi=np.linspace(0, 1, 12) ** 2, ess_threshold=0.7, resampler="standard",
                              init="impcov_normal", init_seed=15), 13),
}


def init_params_for(kw):
    """Host-side initial particles of the cases whose priors the oracle / device do not sample."""
    if "iw_seed" in kw:
        return mvn_init_params(kw["n"], kw["iw_seed"])
    if kw.get("init") == "constrained":
        return constrained_init_params(kw["n"], kw["init_seed"])
    if kw.get("init") == "impcov_normal":
        return impcov_normal_init_params(kw["n"], kw["init_seed"])
    return None


def mvn_init_params(n, seed):
    """Host-side prior draw for the MVNormal case (uniform a,b + inverse-Wishart
    covariance columns), as the reference's ImproperCov.rvs does it
    (smcpy/priors.py:118-123); stored in the fixture so the box needs no scipy
    stream parity."""
    from scipy.stats import invwishart, uniform
    rng = np.random.default_rng(seed)
    a = uniform(0.0, 6.0).rvs(n, random_state=rng)
    b = uniform(0.0, 6.0).rvs(n, random_state=rng)
    cov = invwishart(5, np.eye(3)).rvs(n, random_state=rng)
    i1, i2 = np.triu_indices(3)
    return np.column_stack([a, b, cov[:, i1, i2]])
