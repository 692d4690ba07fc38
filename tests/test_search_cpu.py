"""
The adaptive phi search's state machine (smcpy_b200/csrc/search.cuh -- the code the cooperative kernel
runs) driven on the host through the library's test hook, against scipy.optimize.bisect on the same
margin function: AdaptiveSampler.optimize_step (smcpy/smc/samplers.py:250-262) must return the SAME phi,
bit for bit, while evaluating far fewer points.  No GPU needed.
"""

import ctypes as C

import numpy as np
import pytest
from scipy.optimize import bisect

import smcpy_b200._lib as L


def _run(ll, phi_old, target, hint=0.0, plain=0):
    ll = np.ascontiguousarray(ll, dtype=np.float64)
    state = np.zeros(16)
    trace = np.full(128, np.nan)
    L.call("smcb_test_search_host", ll.ctypes.data_as(C.c_void_p), ll.size, float(phi_old), float(target),
           float(hint), int(plain), state.ctypes.data_as(C.c_void_p), trace.ctypes.data_as(C.c_void_p), 128,
           launches=0)
    return state, trace[~np.isnan(trace)]


def _scipy_phi(ll, phi_old, target):
    ll = np.ascontiguousarray(ll, dtype=np.float64)
    p = ll.ctypes.data_as(C.c_void_p)
    calls = [0]

    def f(x):
        calls[0] += 1
        return L.lib.smcb_test_search_margin(p, ll.size, float(phi_old), float(x), float(target))
    if f(1.0) > 0:                              # samplers.py:251-252
        return 1.0, calls[0]
    # samplers.py:266-267: delta_phi = 0 gives ESS = N exactly
    g = lambda x: (ll.size - ll.size * target) if x == phi_old else f(x)
    return bisect(g, phi_old, 1.0), calls[0] + 1


def _cases():
    rng = np.random.default_rng(0)
    x = np.linspace(0.5, 2.5, 100)
    y = 2 * x + 3.5 + rng.normal(0, 0.5, 100)
    th = rng.uniform(-6, 6, (2000, 2))
    ll_lin = -0.5 * (((th[:, :1] * x + th[:, 1:]) - y) ** 2).sum(1) / 0.25     # C1-like: huge spread
    out = [("c1_first_step", ll_lin, 0.0, 0.8)]
    g32 = -0.5 * (rng.uniform(-10, 10, (5000, 32)) ** 2).sum(1)               # C5-like prior cloud
    out.append(("c5_first_step", g32, 0.0, 0.8))
    post = -0.5 * rng.chisquare(32, 5000)                                      # late steps: narrow spread
    out += [("late_%g" % p0, post, p0, 0.8) for p0 in (0.05, 0.3, 0.7, 0.93)]
    out.append(("full_step_ok", post * 1e-3, 0.9, 0.8))                        # phi = 1 meets the target
    out.append(("target_0.5", ll_lin, 1e-4, 0.5))
    out.append(("target_0.99", g32, 0.0, 0.99))
    out.append(("two_particles", np.array([-1.0, -3.0]), 0.0, 0.8))
    out.append(("with_neg_inf", np.concatenate([post, [-np.inf] * 50]), 0.2, 0.8))
    out.append(("many_neg_inf", np.concatenate([post[:100], [-np.inf] * 400]), 0.2, 0.8))   # ESS(0+) < target N
    out.append(("all_equal", np.full(100, -3.0), 0.1, 0.8))
    out.append(("heavy_tail", -np.abs(rng.standard_cauchy(4000)) * 50, 0.0, 0.8))
    out.append(("bimodal", np.concatenate([rng.normal(-5, 1, 3000), rng.normal(-500, 1, 3000)]), 0.0, 0.8))
    for k in range(6):                                                         # random problems
        n = int(rng.integers(50, 3000))
        out.append(("rand%d" % k, rng.normal(-100, 10 ** rng.uniform(-1, 3), n), float(rng.uniform(0, 0.9)),
                    float(rng.uniform(0.3, 0.95))))
    return out


@pytest.mark.parametrize("name, ll, phi_old, target", _cases(), ids=[c[0] for c in _cases()])
def test_search_reproduces_scipy_bisect(name, ll, phi_old, target):
    want, scipy_calls = _scipy_phi(ll, phi_old, target)
    for hint in (0.0, 0.37 * (want - phi_old), 3.0 * (want - phi_old), 50.0 * (want - phi_old)):
        st, trace = _run(ll, phi_old, target, hint=hint)
        assert st[2] == 1.0, (name, st)
        assert st[0] == want, (name, hint, st[0], want, st[0] - want)           # bit-equal to scipy's root
        assert st[10] == len(trace)
        if want < 1.0 and name not in ("many_neg_inf", "all_equal"):
            assert st[10] <= 22, (name, hint, st[10], trace)                    # (scipy: ~40-45 evaluations)
    st_plain, trace_plain = _run(ll, phi_old, target, plain=1)
    # scipy_calls = full-step test + f(a) + f(b) + midpoints; the plain form evaluates the full step and the midpoints
    assert st_plain[0] == want and st_plain[10] == (1 if want == 1.0 else scipy_calls - 2)
    assert st_plain[1] == (1 if want == 1.0 else scipy_calls)


def test_search_pass_counts_are_small():
    """Typical adaptive sequence on a C5-like problem: every step in <= 14 passes with the previous
    step as the hint (the kernel's pass count is what the device timing scales with)."""
    rng = np.random.default_rng(3)
    ll = -0.5 * (rng.uniform(-10, 10, (20000, 32)) ** 2).sum(1)
    phi, hint, passes = 0.0, 0.0, []
    for _ in range(40):
        st, _ = _run(ll, phi, 0.8, hint=hint)
        want, _ = _scipy_phi(ll, phi, 0.8)
        assert st[0] == want
        passes.append(int(st[10]))
        hint, phi = st[0] - phi, st[0]
        if phi >= 1.0:
            break
    assert max(passes[1:]) <= 14 and np.mean(passes) <= 12, passes
