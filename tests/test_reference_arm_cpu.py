"""The workloads with which `bench.py --impl reference` drives the UNMODIFIED reference (baseline/_ref) on the host:
each builder runs through the reference's own public API at a tiny particle count.  Skipped when baseline/_ref is
absent; nothing here reads /root/reference and nothing of the product (smcpy_b200) is on this path."""

import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fc():
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "smcpy")):
        pytest.skip("reference not installed under baseline/_ref")
    from baseline import fork_comm
    if fork_comm.import_smcpy() is None:
        pytest.skip("reference not importable")
    return fork_comm


@pytest.mark.parametrize("name, n, K, lo, hi", [("_c1_build", 60, 3, -90.0, -60.0), ("_c3_build", 24, 2, -480.0, -330.0),
                                                ("_c4_build", 48, 4, -230.0, -140.0), ("_c5_build", 64, 2, -140.0, -60.0)])
def test_reference_arm_builders_run_through_the_reference(fc, name, n, K, lo, hi):
    """AdaptiveSampler of the reference to phi = 1 on each configuration's model / data / priors / likelihood class
    (C4: MVNormal + ImproperCov, C3: odeint spring-mass with the noise std as third parameter): finite log-evidence in
    the range the configuration's data allows, at least two SMC steps."""
    import bench
    dt, mll, steps = fc.adaptive_run(getattr(bench, name), n, K, 0.8, 1, 1)
    assert dt > 0.0 and steps >= 2
    assert np.isfinite(mll) and lo < mll < hi, mll


def test_reference_c2_pass_on_two_fork_ranks(fc):
    """The C2 FixedPhi pass through ParallelVectorMCMC over the fork-based communicator (two host processes): the
    same log-evidence as the single-process run of the same seed -- rank 0's draws are the ones every rank uses."""
    import bench
    x, data = bench.linear_data(100)
    phi = np.linspace(0.0, 1.0, 6)
    _, mll1 = fc.c2_fixed_run(x, data, 64, 2, phi, 0.8, 3, 1)
    _, mll2 = fc.c2_fixed_run(x, data, 64, 2, phi, 0.8, 3, 2)
    assert np.isfinite(mll1) and abs(mll1 - mll2) < 1e-9 * abs(mll1)
