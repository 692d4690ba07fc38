"""
Host-side logic that needs no GPU: the phi policies of FixedTimeSampler / MaxStepSampler pinned to
the known answers of the reference's own unit tests (tests/unit/test_samplers.py:343-631, which
mock the adaptive search), adaptive-step helpers, proposals, host resampling uniforms, the
native generator's stream bookkeeping.
"""

import numpy as np
import pytest
from scipy import stats

import smcpy_b200 as sb
from smcpy_b200 import samplers


@pytest.fixture
def kernel():
    mc = sb.B200VectorMCMC(sb.LinearModel(np.arange(3.0)), np.zeros(3), [stats.uniform(0, 1)] * 2, 1.0)
    return sb.VectorMCMCKernel(mc, ["a", "b"], rng=np.random.default_rng(0))


# ---- FixedTimeSampler (reference tests/unit/test_samplers.py:343-498) -------------------------
@pytest.mark.parametrize("thr, knock", ((1, 0), (0, 1), (0.5, 0.5)))
def test_fixed_time_initialised(kernel, thr, knock):
    smc = sb.FixedTimeSampler(kernel, time=10, norm_time_threshold=thr, time_knockdown_factor=knock)
    assert smc.buffer_time == thr * 10 and smc.final_time == knock * 10
    d = sb.FixedTimeSampler(kernel, time=1)
    assert (d.buffer_time, d.final_time, d._time_history, d._buffer_phi, d._start_time) == (0.8, 0.95, [0], None, None)


def test_fixed_time_tracks_elapsed_time_and_buffer_phi(kernel, monkeypatch):
    monkeypatch.setattr(samplers.SamplerBase, "_do_smc_step", lambda self, phi, num_mcmc_samples: None)
    ticks = iter([2, 4, 7])
    monkeypatch.setattr(samplers.time, "time", lambda: next(ticks))
    smc = sb.FixedTimeSampler(kernel, time=100)
    smc._start_time = 1
    for want_hist, want_prev in (([0, 1], 1), ([0, 1, 3], 2), ([0, 1, 3, 6], 3)):
        smc._do_smc_step(phi=1, num_mcmc_samples=1)
        assert smc._time_history == want_hist and smc.prev_time_step == want_prev
    ticks = iter([2, 51, 52])
    smc = sb.FixedTimeSampler(kernel, time=100, norm_time_threshold=0.5, time_knockdown_factor=0.8)
    smc._start_time = 1
    got = []
    for phi in (0.1, 0.6, 0.9):
        smc._do_smc_step(phi=phi, num_mcmc_samples=1)
        got.append(smc._buffer_phi)
    assert got == [None, 0.6, 0.6]


def test_fixed_time_interpolates_phi_in_log_space(kernel, monkeypatch):
    monkeypatch.setattr(samplers.AdaptiveSampler, "optimize_step", lambda self, particles, phi_old, target_ess=1: 1e-100)
    smc = sb.FixedTimeSampler(kernel, time=100, norm_time_threshold=0.5, time_knockdown_factor=1)
    smc._time_history.extend([40, 50])
    smc._buffer_phi = 1e-5
    assert smc.optimize_step(particles=1, phi_old=1) == pytest.approx(1e-4, rel=1e-12)


@pytest.mark.parametrize("overshot", (30, 50))
def test_fixed_time_jumps_to_one_when_out_of_time(kernel, monkeypatch, overshot):
    monkeypatch.setattr(samplers.AdaptiveSampler, "optimize_step", lambda self, particles, phi_old, target_ess=1: 0.999)
    smc = sb.FixedTimeSampler(kernel, time=100, time_knockdown_factor=0.9)
    smc._time_history.append(overshot)
    assert smc.optimize_step(particles=1, phi_old=1) == 1


@pytest.mark.parametrize("opt_phi, expected", ((0.1, 0.5), (0.9, 0.9)))
def test_fixed_time_takes_the_larger_phi(kernel, monkeypatch, opt_phi, expected):
    monkeypatch.setattr(samplers.AdaptiveSampler, "optimize_step", lambda self, particles, phi_old, target_ess=1: opt_phi)
    smc = sb.FixedTimeSampler(kernel, time=100, norm_time_threshold=0.0, time_knockdown_factor=0.8)
    smc._buffer_phi = 0.5
    smc._time_history.append(0)
    assert smc.optimize_step(particles=1, phi_old=1) == pytest.approx(expected)


# ---- MaxStepSampler (reference tests/unit/test_samplers.py:501-631) ------------------------------
def test_max_step_policy(kernel, monkeypatch):
    smc = sb.MaxStepSampler(kernel, max_steps=10)
    assert (smc.buffer_step, smc.final_step, smc._step_count, smc._buffer_phi) == (7, 9, 0, None)
    monkeypatch.setattr(samplers.SamplerBase, "_do_smc_step", lambda self, phi, num_mcmc_samples: None)
    phis = []
    for i in range(9):
        smc._do_smc_step(phi=0.01 * (i + 1), num_mcmc_samples=1)
        phis.append(smc._buffer_phi)
    assert phis[:7] == [None] * 7 and phis[7] == pytest.approx(0.08) and phis[8] == pytest.approx(0.08)
    assert smc._step_count == 9
    monkeypatch.setattr(samplers.AdaptiveSampler, "optimize_step", lambda self, particles, phi_old, target_ess=1: 1e-6)
    assert smc.optimize_step(particles=1, phi_old=0.09) == 1.0                  # last allowed step
    smc._step_count = 8
    # halfway (in log10) between the buffer phi and 1
    assert smc.optimize_step(particles=1, phi_old=0.09) == pytest.approx(10 ** (0.5 * np.log10(0.08)))
    smc._buffer_phi = None
    smc._step_count = 3
    assert smc.optimize_step(particles=1, phi_old=0.01) == 1e-6


# ---- adaptive-step helpers (samplers.py:290-301) ----------------------------------------------------
def test_select_phi_and_min_dphi(kernel):
    smc = sb.AdaptiveSampler(kernel, show_progress_bar=False)
    assert smc._select_phi([0.2, 0.5, 0.35, 1.0], 0.3) == 0.35
    smc._phi_sequence = [0, 0.4]
    assert smc._verify_min_dphi(0.41, 0.05) == pytest.approx(0.45)
    assert smc._verify_min_dphi(0.5, 0.05) == 0.5
    assert smc._verify_min_dphi(0.41, None) == 0.41


# ---- proposals (proposals.py:4-34) ------------------------------------------------------------------------
def test_multivar_independent_and_its_device_table():
    from smcpy_b200.proposals import describe_proposal
    from smcpy_b200 import _lib
    q = sb.MultivarIndependent(stats.norm(2.0, 1.0), stats.uniform(-1.0, 4.0), stats.norm(loc=0.5, scale=2.0))
    x = q.rvs(7, random_state=np.random.default_rng(3))
    assert x.shape == (7, 3)
    ref = np.random.default_rng(3)
    want = np.column_stack([stats.norm(2.0, 1.0).rvs(7, random_state=ref), stats.uniform(-1.0, 4.0).rvs(7, random_state=ref),
                            stats.norm(0.5, 2.0).rvs(7, random_state=ref)])
    np.testing.assert_array_equal(x, want)
    lp = q.logpdf(x)
    assert lp.shape == (7, 1)
    np.testing.assert_allclose(lp[:, 0], stats.norm(2, 1).logpdf(x[:, 0]) + stats.uniform(-1, 4).logpdf(x[:, 1])
                               + stats.norm(0.5, 2).logpdf(x[:, 2]))
    assert describe_proposal(q) == [(_lib.PROPOSAL_NORMAL, 2.0, 1.0), (_lib.PROPOSAL_UNIFORM, -1.0, 4.0),
                                    (_lib.PROPOSAL_NORMAL, 0.5, 2.0)]
    assert describe_proposal(sb.MultivarIndependent(stats.gamma(2.0))) is None
    assert describe_proposal(object()) is None
    mv = sb.MultivarIndependent(stats.multivariate_normal([0, 0], np.eye(2)), stats.norm(0, 1))
    assert mv.rvs(4, random_state=np.random.default_rng(0)).shape == (4, 3)
    assert mv.logpdf(np.zeros((4, 3))).shape == (4, 1)


# ---- host resampling uniforms (resampler_rngs.py:4-9) -----------------------------------------------------
def test_host_resampler_uniforms(kernel):
    kernel.rng = np.random.default_rng(11)
    u = sb.standard(kernel, 6)
    np.testing.assert_array_equal(u, np.random.default_rng(11).uniform(0, 1, 6))
    kernel.rng = np.random.default_rng(12)
    s = sb.stratified(kernel, 5)
    np.testing.assert_array_equal(s, 1 / 5 * (np.arange(5) + np.random.default_rng(12).uniform(0, 1, 5)))
    assert np.all(np.diff(s) > 0) and np.all((s >= np.arange(5) / 5) & (s < (np.arange(5) + 1) / 5))
    kernel.rng = np.random.default_rng(13)
    y = sb.systematic(kernel, 4)
    np.testing.assert_allclose(np.diff(y), 0.25)
    assert (sb.standard.kind, sb.stratified.kind, sb.systematic.kind) == (0, 1, 2)
    native = sb.VectorMCMCKernel(kernel._mcmc, ["a", "b"], rng=5) if hasattr(kernel, "_mcmc") else None
    if native is not None:
        with pytest.raises(TypeError):
            sb.standard(native, 3)


def test_native_generator_streams():
    r = sb.PhiloxRNG(42)
    a, b = r.next_stream(), r.next_stream()
    assert b == a + 1 and r.seed == 42
    assert sb.PhiloxRNG(42).next_stream() == a


# ---- host-evaluated priors (smcpy/priors.py:38-90,150-247): pure host objects, no GPU needed -------------
def test_invwishart_prior_host_logpdf_and_rvs():
    from oracle import smc_oracle as orc
    scale = np.array([[2.0, 0.3], [0.3, 1.0]])
    p = sb.InvWishart(4, scale)
    assert p.dim == 3
    x = p.rvs(7, random_state=np.random.default_rng(3))
    assert x.shape == (7, 3)
    assert p.rvs(1, random_state=np.random.default_rng(3)).shape == (1, 3)
    lp = p.logpdf(x)
    assert lp.shape == (7, 1)
    np.testing.assert_allclose(lp[:, 0], orc.logpdf_invwishart(x, 2, 4, scale), rtol=1e-13)
    covs = np.array([[[a, b], [b, c]] for a, b, c in x])
    np.testing.assert_allclose(lp[:, 0], stats.invwishart(4, scale).logpdf(np.transpose(covs, (1, 2, 0))), rtol=1e-13)
    np.testing.assert_allclose(p.pdf(x), np.exp(lp), rtol=1e-12)
    # one matrix of the batch not positive definite: the reference's quirk gives the whole batch -inf / 0
    bad = x.copy()
    bad[2] = [1.0, 5.0, 1.0]
    assert np.isneginf(p.logpdf(bad)).all() and (p.pdf(bad) == 0).all()


def test_improper_constrained_uniform_prior():
    bounds = np.array([[-1.0, 0.0], [1.0, 2.0]])
    cons = lambda x: x[:, 0] < x[:, 1]
    p = sb.ImproperConstrainedUniform(cons, bounds=bounds)
    assert p.dim == 2
    x = np.array([[0.0, 1.0], [0.5, 0.2], [-1.0, 0.0], [0.0, 2.5], [1.0, 2.0], [-1.5, 1.0]])
    np.testing.assert_array_equal(p.logpdf(x)[:, 0], [0.0, -np.inf, 0.0, -np.inf, 0.0, -np.inf])
    s = p.rvs(200, random_state=np.random.default_rng(1))
    assert s.shape == (200, 2) and cons(s).all() and (s >= bounds[0]).all() and (s <= bounds[1]).all()
    assert sb.ImproperConstrainedUniform(cons, dim=2).dim == 2
    np.testing.assert_array_equal(sb.ImproperConstrainedUniform(cons, dim=2).logpdf(x)[:, 0],
                                  [0.0, -np.inf, 0.0, 0.0, 0.0, 0.0])
    with pytest.raises(ValueError):
        sb.ImproperConstrainedUniform(cons)
    with pytest.raises(NotImplementedError):
        sb.ImproperConstrainedUniform(cons, dim=2).rvs(3)
    with pytest.raises(TypeError):
        p.rvs(3, random_state=np.random.RandomState(1))
    with pytest.raises(ValueError, match="Max rvs tries"):
        sb.ImproperConstrainedUniform(lambda x: x[:, 0] > 5, bounds=bounds, max_rvs_tries=3).rvs(2, np.random.default_rng(0))


def test_prior_table_descriptors():
    from smcpy_b200 import _lib
    from smcpy_b200.priors import describe_prior
    assert describe_prior(stats.uniform(-6, 12))[:4] == (_lib.PRIOR_UNIFORM, 1, -6.0, 12.0)
    assert describe_prior(sb.ImproperUniform(0, 3))[:4] == (_lib.PRIOR_IMPROPER_UNIFORM, 1, 0.0, 3.0)
    assert describe_prior(sb.ImproperCov(3, dof=5))[:3] == (_lib.PRIOR_IMPROPER_COV, 3, 5.0)
    assert describe_prior(sb.InvWishart(5, np.eye(3)))[:2] == (_lib.PRIOR_EXTERNAL, 6)
    assert describe_prior(stats.norm(0, 1))[:2] == (_lib.PRIOR_EXTERNAL, 1)       # any object with logpdf
    with pytest.raises(TypeError):
        describe_prior(object())
