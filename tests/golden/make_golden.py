"""
Generates tests/golden/*.npz by running the UNMODIFIED reference (nasa/SMCPy,
imported read-only from /root/reference) in the build container.  The fixtures
are data, not code; nothing from the reference is copied.  Re-run with

    python tests/golden/make_golden.py

Only this script reads /root/reference; tests, smoke() and bench.py read the
committed .npz files.
"""

import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))     # only HDF5Storage needs it
sys.path.insert(0, os.environ.get("SMCPY_REFERENCE", "/root/reference"))

import smcpy                                                  # noqa: E402
from smcpy import (AdaptiveSampler, FixedPhiSampler, VectorMCMC, VectorMCMCKernel,  # noqa: E402
                   ImproperUniform, ImproperCov, MVNormal)
from smcpy.priors import InvWishart, ImproperConstrainedUniform  # noqa: E402
from smcpy.log_likelihoods import MultiSourceNormal, Normal   # noqa: E402
from smcpy.paths import GeometricPath                         # noqa: E402
from smcpy.proposals import MultivarIndependent               # noqa: E402
from smcpy.resampler_rngs import standard, stratified         # noqa: E402
from smcpy.smc.particles import Particles                     # noqa: E402
from smcpy.smc.updater import Updater                         # noqa: E402
from scipy import stats                                       # noqa: E402

import configs                                                # noqa: E402
from oracle import smc_oracle as orc                          # noqa: E402  (numpy models only)


def ref_objects(problem):
    """Reference-side model / priors / likelihood for an oracle-style problem."""
    kind, arg = problem["model"]
    if kind in ("linear", "linear_features"):
        x = np.asarray(arg, float)
        model = lambda th: orc.model_linear(th, x)
    elif kind == "spring_mass":
        model = lambda th: orc.model_spring_mass_rk4(th, **arg)
    elif kind == "identity":
        model = lambda th: th.copy()
    else:
        raise ValueError(kind)
    priors = []
    for p in problem["priors"]:
        if p[0] == "uniform":
            priors.append(stats.uniform(p[1], p[2]))
        elif p[0] == "improper_uniform":
            priors.append(ImproperUniform(p[1], p[2]))
        elif p[0] == "improper_cov":
            priors.append(ImproperCov(int(p[1]), dof=5, S=np.eye(int(p[1]))))
        elif p[0] == "invwishart":
            priors.append(InvWishart(p[2], np.asarray(p[3], float)))
        elif p[0] == "constrained":
            priors.append(ImproperConstrainedUniform(p[2], bounds=np.asarray(p[3], float)))
    lk, largs = problem["loglike"]
    ll_cls = {"normal": Normal, "mvnormal": MVNormal, "multisource": MultiSourceNormal}[lk]
    if lk == "multisource":
        largs = [tuple(largs[0]), tuple(largs[1])]
    return model, priors, largs, ll_cls


class Recorder:
    """Wraps a few reference methods to record per-step internals."""

    def __init__(self):
        self.ess, self.resampled, self.idx, self.cov, self.accepts = [], [], [], [], []
        self._idx_calls = 0
        self._acc_step = []

    def install(self):
        rec = self
        self._orig = (Updater.resample_if_needed, Updater._generate_resample_indices,
                      Particles.compute_covariance, VectorMCMC.get_rejections,
                      VectorMCMC.smc_metropolis)

        def resample_if_needed(upd, particles):
            out = rec._orig[0](upd, particles)
            rec.ess.append(upd.ess)
            rec.resampled.append(upd.resampled)
            if not upd.resampled:
                rec.idx.append(np.zeros(0, dtype=np.int64))
            return out

        def gen_idx(upd, particles):
            out = rec._orig[1](upd, particles)
            if rec._idx_calls % 2 == 0:                       # first draw selects (quirk Q1)
                rec.idx.append(np.asarray(out, dtype=np.int64))
            rec._idx_calls += 1
            return out

        def cov(p):
            out = rec._orig[2](p)
            rec.cov.append(np.array(out))
            return out

        def rej(m, ratios):
            out = rec._orig[3](m, ratios)
            rec._acc_step.append(int(ratios.shape[0] - np.sum(out)))
            return out

        def smc_met(m, inputs, k, c):
            rec._acc_step = []
            out = rec._orig[4](m, inputs, k, c)
            rec.accepts.append(list(rec._acc_step))
            return out

        Updater.resample_if_needed = resample_if_needed
        Updater._generate_resample_indices = gen_idx
        Particles.compute_covariance = cov
        VectorMCMC.get_rejections = rej
        VectorMCMC.smc_metropolis = smc_met

    def remove(self):
        (Updater.resample_if_needed, Updater._generate_resample_indices,
         Particles.compute_covariance, VectorMCMC.get_rejections,
         VectorMCMC.smc_metropolis) = self._orig


def run_case(name):
    build, kind, kw, seed = configs.CASES[name]
    problem = build()
    model, priors, largs, ll_cls = ref_objects(problem)
    rng = np.random.default_rng(seed)
    mcmc = VectorMCMC(model, problem["data"], priors, largs, ll_cls)
    path = None
    if problem.get("proposal"):
        dists = [stats.norm(loc, sc) if k == "normal" else stats.uniform(loc, sc) for k, loc, sc in problem["proposal"]]
        path = GeometricPath(proposal=MultivarIndependent(*dists), required_phi=problem.get("required_phi", 1))
    kernel = VectorMCMCKernel(mcmc, param_order=problem["names"], rng=rng, path=path)
    init = configs.init_params_for(kw)
    if init is not None:
        kernel.sample_from_prior = lambda n: dict(zip(kernel.param_order, init.T))
    res = {"standard": standard, "stratified": stratified}[kw["resampler"]]
    rec = Recorder()
    rec.install()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if kind == "adaptive":
                smc = AdaptiveSampler(kernel, show_progress_bar=False)
                steps, mll = smc.sample(kw["n"], kw["K"], target_ess=kw["target_ess"],
                                        resample_rng=res)
            else:
                smc = FixedPhiSampler(kernel, show_progress_bar=False)
                steps, mll = smc.sample(kw["n"], kw["K"], kw["phi"], kw["ess_threshold"],
                                        resample_rng=res)
    finally:
        rec.remove()
    steps = list(steps)
    out = {
        "phi": np.array([s.attrs["phi"] for s in steps], float),
        "mll": np.asarray(mll, float),
        "total_unnorm_log_weight": np.array([s.total_unnorm_log_weight for s in steps]),
        "mutation_ratio": np.array([s.attrs["mutation_ratio"] for s in steps], float),
        "ess": np.array(rec.ess, float),
        "resampled": np.array(rec.resampled, bool),
        "cov": np.array(rec.cov),
        "accepts": np.array(rec.accepts, dtype=np.int64),
        "mean": np.array([np.sum(s.params * s.weights, axis=0) for s in steps]),
        "log_like_sum": np.array([np.sum(s.log_likes) for s in steps]),
        "params_last": steps[-1].params,
        "log_likes_last": steps[-1].log_likes,
        "log_weights_last": steps[-1].log_weights,
        "params_0": steps[0].params,
        "log_likes_0": steps[0].log_likes,
        "params_1": steps[1].params,
        "log_likes_1": steps[1].log_likes,
        "log_weights_1": steps[1].log_weights,
        "idx_1": rec.idx[0],
        "idx_last": rec.idx[-1],
        "rng_check": rng.uniform(0, 1, 4),                    # tape position after the run
    }
    if init is not None:
        out["init_params"] = init
    if kind == "adaptive":
        out["req_phi_index"] = np.array(smc.req_phi_index, dtype=np.int64)
    if steps[0].params.size * len(steps) <= 40000:            # small cases: every step
        out["params_all"] = np.array([s.params for s in steps])
        out["log_likes_all"] = np.array([s.log_likes for s in steps])
        out["log_weights_all"] = np.array([s.log_weights for s in steps])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "steps", len(steps), "mll", mll[-1], "phi[1]", out["phi"][1])


def unit_vectors():
    """Function-level vectors: reference outputs on seeded random inputs."""
    rng = np.random.default_rng(1234)
    out = {}

    # particles: normalise / logsum / ess / cov / mean       (particles.py)
    n = 257
    lw = rng.normal(-3, 4, n)
    lw[5] = -np.inf
    lw[17] = -800.0
    prm = {"a": rng.normal(2, 1, n), "b": rng.normal(-1, 3, n), "c": rng.uniform(0, 1, n)}
    p = Particles(prm, rng.normal(0, 1, n), lw)
    out.update(p_lw_in=lw, p_params=p.params, p_log_w=p.log_weights, p_w=p.weights,
               p_total=p.total_unnorm_log_weight, p_ess=p.compute_ess(),
               p_cov=p.compute_covariance(), p_mean=p.compute_mean(package=False),
               p_var=p.compute_variance(package=False))

    # path: inc weights / logpdf, plain, with required phi and with proposal (paths.py)
    ll = rng.normal(-50, 20, (n, 1))
    lp = rng.normal(-3, 1, (n, 1))
    lp[3] = -np.inf
    ll[9] = -np.inf
    lq = rng.normal(-2, 1, (n, 1))
    out.update(path_ll=ll, path_lp=lp, path_lq=lq)

    class FakeProposal:
        def logpdf(self, x):
            return lq

    for tag, path in (("plain", GeometricPath()),
                      ("req", GeometricPath(required_phi=[0.3, 0.6])),
                      ("prop", GeometricPath(proposal=FakeProposal())),
                      ("propreq", GeometricPath(proposal=FakeProposal(), required_phi=0.4))):
        phis = [0.05, 0.2, 0.5, 1.0]
        incs, lpdf = [], []
        for ph in phis:
            path.phi = ph
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                incs.append(path.inc_log_weights(None, ll, lp))
                lpdf.append(path.logpdf(None, ll, lp))
        out["path_%s_inc" % tag] = np.array(incs)
        out["path_%s_logpdf" % tag] = np.array(lpdf)
    out["path_phis"] = np.array([0.0, 0.05, 0.2, 0.5, 1.0])

    # resampling: digitize(u, cumsum(w))                      (updater.py:150-152)
    w = p.weights.flatten()
    u = rng.uniform(0, 1, 1000) * np.cumsum(w)[-1] * (1 - 1e-12)
    out.update(rs_w=w, rs_u=u, rs_idx=np.digitize(u, np.cumsum(w)))

    # likelihoods                                             (log_likelihoods.py)
    th = np.column_stack([rng.uniform(1, 3, 40), rng.uniform(2, 5, 40), rng.uniform(0.2, 2, 40)])
    x = np.linspace(0.5, 2.5, 33)
    data = 2 * x + 3.5 + rng.normal(0, 0.5, 33)
    out.update(nl_theta=th, nl_x=x, nl_data=data,
               nl_fixed=Normal(lambda t: orc.model_linear(t, x), data, 0.7)(th[:, :2]),
               nl_unknown=Normal(lambda t: orc.model_linear(t, x), data, None)(th))
    X = np.arange(3.0)
    mdata = np.tile(2 * X + 3.5, (7, 1)) + rng.multivariate_normal([0, 0, 0], configs.TRUE_COV, 7)
    A = rng.normal(0, 1, (40, 3, 3))
    covs = A @ np.transpose(A, (0, 2, 1)) + 0.3 * np.eye(3)
    i1, i2 = np.triu_indices(3)
    mth = np.column_stack([th[:, :2], covs[:, i1, i2]])
    out.update(mv_theta=mth, mv_data=mdata,
               mv_all=MVNormal(lambda t: orc.model_linear(t, X), mdata, [None] * 6)(mth).flatten())
    args_mixed = [0.5, None, 0.005, None, 0.04, 1.0]
    mth2 = np.column_stack([th[:, :2], rng.uniform(-0.1, 0.1, 40), rng.uniform(0.3, 0.6, 40)])
    out.update(mv_theta_mixed=mth2,
               mv_mixed=MVNormal(lambda t: orc.model_linear(t, X), mdata, args_mixed)(mth2).flatten())

    # MultiSourceNormal incl. an empty segment and two sampled std-devs  (log_likelihoods.py:52-118)
    msx = np.linspace(0, 1, 12)
    msd = 1.5 * msx + 0.5 + rng.normal(0, 0.3, 12)
    msth = np.column_stack([rng.uniform(1, 2, 30), rng.uniform(0, 1, 30), rng.uniform(0.2, 1, 30), rng.uniform(0.2, 1, 30)])
    msargs = [(5, 0, 3, 4), (None, 0.4, 0.25, None)]
    out.update(ms_x=msx, ms_data=msd, ms_theta=msth,
               ms_ll=MultiSourceNormal(lambda t: orc.model_linear(t, msx), msd, msargs)(msth))

    # priors                                                   (priors.py, scipy.stats.uniform)
    xs = np.concatenate([rng.uniform(-8, 8, 50), [-6.0, 6.0, np.nextafter(6.0, 7.0), np.nextafter(-6.0, -7)]])
    out.update(pr_x=xs, pr_uniform=stats.uniform(-6, 12).logpdf(xs),
               pr_improper=ImproperUniform(-6, 6).logpdf(xs))
    cc = rng.normal(0, 1, (60, 6))
    cc[:30] = covs[:30][:, i1, i2]
    out.update(pr_cov_x=cc, pr_cov=ImproperCov(3).logpdf(cc).flatten())

    # adaptive phi: margin + scipy bisect via the reference sampler (samplers.py:250-301)
    from smcpy.mcmc.kernel_base import KernelBase

    class StubKernel(KernelBase):
        def __init__(self):
            self.path = GeometricPath()

        def get_log_priors(self, pd):
            return np.full((len(pd["a"]), 1), -np.log(144.0))

        mutate_particles = sample_from_prior = sample_from_proposal = None
        get_log_likelihoods = set_mcmc_rng = None

    KernelBase.__abstractmethods__ = frozenset()
    phis, margins = [], []
    lls = []
    for scale in (1.0, 30.0, 3000.0):
        n2 = 1000
        l2 = rng.normal(-10, 1, n2) * scale
        pp = Particles({"a": rng.normal(0, 1, n2), "b": rng.normal(0, 1, n2)}, l2,
                       np.full(n2, -np.log(n2)))
        smc = AdaptiveSampler(StubKernel(), show_progress_bar=False)
        row_p, row_m = [], []
        for phi_old in (0.0, 1e-4, 0.3):
            smc._mcmc_kernel.path = GeometricPath()
            if phi_old > 0:
                smc._mcmc_kernel.path.phi = phi_old
            row_p.append(smc.optimize_step(pp, phi_old, 0.8))
            row_m.append(smc.predict_ess_margin(min(1.0, phi_old + 0.01), phi_old, pp, 0.8))
        phis.append(row_p)
        margins.append(row_m)
        lls.append(l2)
    out.update(ad_ll=np.array(lls), ad_phi=np.array(phis, float), ad_margin=np.array(margins, float))

    np.savez_compressed(os.path.join(HERE, "unit_vectors.npz"), **out)
    print("unit_vectors", len(out), "arrays")


def metropolis_fixture():
    """Stand-alone VectorMCMC.metropolis chain with covariance adaptation (vector_mcmc.py:64-86)."""
    problem = configs.linear_problem(40, unknown_sigma=True)
    rng0 = np.random.default_rng(2)
    theta = np.column_stack([rng0.uniform(1.5, 2.5, 64), rng0.uniform(3, 4, 64), rng0.uniform(0.3, 0.8, 64)])
    cov = np.diag([0.01, 0.02, 0.005])
    x = problem["model"][1]
    m = VectorMCMC(lambda t: orc.model_linear(t, x), problem["data"],
                   [stats.uniform(-6, 12)] * 2 + [stats.uniform(0, 10)], None)
    m.rng = np.random.default_rng(8)
    chain = m.metropolis(theta, 12, cov, adapt_interval=4, adapt_delay=2)
    np.savez_compressed(os.path.join(HERE, "metropolis_chain.npz"), theta=theta, cov=cov, chain=chain)
    print("metropolis_chain", chain.shape)


def hierarch_fixture():
    """ApproxHierarch + MVNHierarchModel (hierarch_log_likelihoods.py) on seeded inputs."""
    from smcpy.hierarch_log_likelihoods import ApproxHierarch, MVNHierarchModel
    rng = np.random.default_rng(77)
    R, ns, p, N = 3, 40, 2, 25
    data = rng.normal(0, 1, (R, ns, p)) + rng.normal(0, 2, (R, 1, p))
    mlls = rng.normal(-20, 3, R)
    lpr = rng.normal(-3, 0.5, (R, ns))
    A = rng.normal(0, 1, (N, p, p))
    covs = A @ np.transpose(A, (0, 2, 1)) + 0.5 * np.eye(p)
    i1, i2 = np.triu_indices(p)
    theta = np.column_stack([rng.normal(0, 2, (N, p)), covs[:, i1, i2]])
    ll = ApproxHierarch(MVNHierarchModel, data, [mlls, lpr])(theta)
    np.savez_compressed(os.path.join(HERE, "hierarch.npz"), data=data, mlls=mlls, lpr=lpr, theta=theta, ll=ll)
    print("hierarch", ll.shape)


if __name__ == "__main__":
    print("reference smcpy", smcpy.__version__, "numpy", np.__version__)
    only = sys.argv[1:]
    if not only:
        unit_vectors()
        metropolis_fixture()
        hierarch_fixture()
    elif "hierarch" in only:
        hierarch_fixture()
    for name in configs.CASES:
        if not only or name in only:
            run_case(name)
