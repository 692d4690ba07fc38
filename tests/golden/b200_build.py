"""
Builds the product objects (B200VectorMCMC + VectorMCMCKernel) for a problem dict of
tests/golden/configs.py.  Pure product + scipy: shared by the tests and by the timing drivers under
tools/, which must not import the oracle.
"""

import numpy as np


def make_b200(problem, rng=None, path=None):
    """B200VectorMCMC + VectorMCMCKernel for an oracle-style problem dict."""
    from scipy import stats
    import smcpy_b200 as sb
    kind, arg = problem["model"]
    if kind in ("linear", "linear_features"):
        model = sb.LinearModel(arg)
    elif kind == "spring_mass":
        model = sb.SpringMassModel(arg["dt"], arg["n_out"], arg["nsub"], arg["x0"], arg["v0"])
    elif kind == "identity":
        model = sb.IdentityModel(arg)
    else:
        raise ValueError(kind)
    priors = []
    for p in problem["priors"]:
        if p[0] == "uniform":
            priors.append(stats.uniform(p[1], p[2]))
        elif p[0] == "improper_uniform":
            priors.append(sb.ImproperUniform(p[1], p[2]))
        elif p[0] == "improper_cov":
            priors.append(sb.ImproperCov(int(p[1]), dof=5, S=np.eye(int(p[1]))))
        elif p[0] == "invwishart":
            priors.append(sb.InvWishart(p[2], np.asarray(p[3], dtype=float)))
        elif p[0] == "constrained":
            priors.append(sb.ImproperConstrainedUniform(p[2], bounds=np.asarray(p[3], dtype=float)))
    lk, largs = problem["loglike"]
    cls = {"normal": sb.Normal, "mvnormal": sb.MVNormal, "multisource": sb.MultiSourceNormal}[lk]
    mcmc = sb.B200VectorMCMC(model, problem["data"], priors, largs, cls)
    if path is None and problem.get("proposal"):
        dists = [stats.norm(loc, sc) if k == "normal" else stats.uniform(loc, sc)
                 for k, loc, sc in problem["proposal"]]
        path = sb.GeometricPath(proposal=sb.MultivarIndependent(*dists),
                                required_phi=problem.get("required_phi", 1))
    kernel = sb.VectorMCMCKernel(mcmc, param_order=problem["names"], rng=rng, path=path)
    return mcmc, kernel
