"""
GPU tests of the fused step kernels (run on the B200 box: `pytest -m gpu`), through the C ABI:
  * cooperative reweight (smcb_reweight_fused) incl. the unmaterialised-weights route of a resampling step;
  * the accelerated phi search (smcb_ess_search) against scipy.optimize.bisect, the plain search and the
    host test hook;
  * one-pass resampling for sorted uniforms (smcb_resample_fused) against the separate
    uniforms / scan / search / count / gather kernels, bit for bit;
  * the SHARDED form of the same kernel with `world` virtual ranks laid out in one GPU's memory, against the
    single-shard result (the peer pointers are plain device pointers here, so the driver's one-GPU test
    lease exercises the owner-computes slot ranges, the remote running sums and the cross-shard writes).
Reference semantics: smcpy/smc/updater.py:97-165, smcpy/smc/samplers.py:250-276, smcpy/resampler_rngs.py:4-9.
"""

import ctypes as C
import warnings

import numpy as np
import pytest

from oracle import smc_oracle as orc
from tests.golden import configs
from tests.helpers import make_b200

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope="module")
def sb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import smcpy_b200
    return smcpy_b200


def _state(n, d):
    from smcpy_b200 import engine
    return engine.DeviceState(n, n, d)


# ------------------------------------------------------------------ (a) cooperative reweight
@pytest.mark.parametrize("n", [3, 1000, 100003, (1 << 21) + 5])
@pytest.mark.parametrize("uniform_in, const_lp", [(True, True), (False, False), (True, False)])
def test_reweight_fused_matches_oracle_and_two_launch_form(sb, n, uniform_in, const_lp):
    from smcpy_b200 import engine, _lib
    rng = np.random.default_rng(n)
    ll = rng.normal(-40, 25, n)
    lpc = -2 * np.log(12.0)
    lp = np.full(n, lpc)
    lw = np.full(n, np.log(1.0 / n)) if uniform_in else rng.normal(-3, 2, n)
    if n > 10:
        ll[3] = -np.inf
        if not const_lp:
            lp[5] = -np.inf
    st = _state(n, 2)
    st.ll.copy_(engine.dev(ll))
    st.lp.copy_(engine.dev(lp))
    st.log_w.copy_(engine.dev(lw))
    st.uniform_weights = uniform_in
    st.lp_const = lpc if const_lp else None
    path = engine.make_path(0.35, 0.2, 1.0)
    un = lw.reshape(-1, 1) + orc.inc_log_weights(ll.reshape(-1, 1), lp.reshape(-1, 1), 0.35, 0.2)
    o_lw, o_w = orc.normalize_log_weights(un)
    o_ess = orc.compute_ess(o_w)
    # (i) always materialise
    ess, tot = engine.reweight(st, path, -1.0)
    assert st.pending is None
    g_lw, g_w = st.log_w.cpu().numpy(), st.w.cpu().numpy()
    fin = np.isfinite(o_lw[:, 0])
    np.testing.assert_array_equal(np.isfinite(g_lw), fin)
    np.testing.assert_allclose(g_lw[fin], o_lw[fin, 0], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(g_w, o_w[:, 0], rtol=RTOL, atol=1e-300)
    assert ess == pytest.approx(o_ess, rel=RTOL) and tot == pytest.approx(orc.logsum(un), rel=RTOL)
    # two-launch form on the same inputs: same scalars and weights to the last bits of the merge order
    st2 = _state(n, 2)
    st2.ll.copy_(st.ll); st2.lp.copy_(st.lp); st2.log_w.copy_(engine.dev(lw))
    lw_in = None if uniform_in else st2.log_w
    _lib.call("smcb_reweight", n, _lib.ptr(st2.ll), _lib.ptr(st2.lp), None, _lib.ptr(lw_in), n, C.byref(path),
              _lib.ptr(st2.log_w_alt), _lib.ptr(st2.w_alt), _lib.ptr(st2.scalars), _lib.ptr(st2.ws), _lib.stream_ptr())
    np.testing.assert_allclose(st2.w_alt.cpu().numpy(), g_w, rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(st2.scalars[:6].cpu().numpy(), st.scalars[:6].cpu().numpy(), rtol=1e-13)
    # (ii) with the resampling test: ESS below threshold -> nothing written, CDF straight from the columns
    st.log_w.copy_(engine.dev(lw))
    st.uniform_weights = uniform_in
    before = (st.log_w_alt.clone(), st.w_alt.clone())
    ess2, _ = engine.reweight(st, path, 1.0)                   # threshold 1: always below
    assert ess2 == ess and st.pending is not None
    assert engine.torch.equal(st.log_w_alt, before[0]) and engine.torch.equal(st.w_alt, before[1])
    engine.weights_cdf(st, total_out=st.rs_tot[0:1])
    cdf_pending = st.cdf.cpu().numpy().copy()
    st.pending = None
    st.w.copy_(engine.dev(g_w))
    engine.weights_cdf(st)
    np.testing.assert_array_equal(cdf_pending, st.cdf.cpu().numpy())       # same weights, same scan: bit-equal
    assert float(st.rs_tot[0].item()) == cdf_pending[-1]
    # (iii) threshold 0: never below -> weights written by the same launch
    st.log_w.copy_(engine.dev(lw))
    st.uniform_weights = uniform_in
    engine.reweight(st, path, 0.0)
    assert st.pending is None
    np.testing.assert_array_equal(st.w.cpu().numpy(), g_w)


# ------------------------------------------------------------------ (b) accelerated search
def _scipy_root(ll, lpc, phi_old, target):
    from scipy.optimize import bisect
    llc, lp = ll.reshape(-1, 1), np.full((ll.size, 1), lpc)
    f = lambda x: orc.ess_margin(llc, lp, x, phi_old, target)
    if f(1.0) > 0:
        return 1.0
    return bisect(f, phi_old, 1.0)


@pytest.mark.parametrize("n", [500, 70001, 1 << 21])
def test_search_matches_scipy_plain_search_and_host_hook(sb, n):
    from smcpy_b200 import engine, _lib
    rng = np.random.default_rng(n)
    x = np.linspace(0.5, 2.5, 100)
    y = 2 * x + 3.5 + rng.normal(0, 0.5, 100)
    th = rng.uniform(-6, 6, (n, 2))
    ll = -0.5 * (((th[:, :1] * x + th[:, 1:]) - y) ** 2).sum(1) / 0.25 - 50 * np.log(2 * np.pi * 0.25)
    lpc = -np.log(144.0)
    st = _state(n, 2)
    st.ll.copy_(engine.dev(ll))
    st.lp.fill_(lpc)

    class Spec:
        lp_const = lpc
    phi_old = 0.0
    for step in range(6):
        st.last_dphi = 0.0 if step % 2 == 0 else st.last_dphi
        phi = engine.find_phi(st, Spec, phi_old, 0.8)
        passes, scipy_evals = st.search_info
        with _lib.option("bisect_plain", 1):
            st.last_dphi = 0.0
            phi_plain = engine.find_phi(st, Spec, phi_old, 0.8)
            passes_plain, scipy_evals_plain = st.search_info
        assert phi == phi_plain, (phi, phi_plain, phi - phi_plain)        # same decisions, same root, bit for bit
        assert scipy_evals == scipy_evals_plain
        if phi < 1.0:
            assert passes <= 16 < passes_plain, (passes, passes_plain)
        if n <= 70001:
            assert phi == pytest.approx(_scipy_root(ll, lpc, phi_old, 0.8), rel=RTOL)
        if phi >= 1.0:
            break
        phi_old = phi
        # thin the cloud towards the mode like a real run would: keep the search non-trivial at the next phi
        ll = ll * 1.0
    assert phi_old > 0


def test_search_with_proposal_column_takes_plain_route(sb):
    """A path with a proposal column is not one exponential family in phi: every midpoint is evaluated."""
    from smcpy_b200 import engine
    rng = np.random.default_rng(5)
    n = 4000
    st = engine.DeviceState(n, n, 2, with_lq=True)
    ll, lp, lq = rng.normal(-30, 8, n), np.full(n, -np.log(144.0)), rng.normal(-3, 1, n)
    st.ll.copy_(engine.dev(ll)); st.lp.copy_(engine.dev(lp)); st.lq.copy_(engine.dev(lq))
    phi = engine.find_phi(st, None, 0.0, 0.8, lam=0.4)
    want = orc.optimize_phi(ll.reshape(-1, 1), lp.reshape(-1, 1), 0.0, 0.8, lam=0.4, lq=lq.reshape(-1, 1))
    assert phi == pytest.approx(want, rel=RTOL)
    assert st.search_info[0] >= 30


# ------------------------------------------------------------------ (c) fused resampling
def _weights(n, seed, kind="exp"):
    rng = np.random.default_rng(seed)
    if kind == "exp":
        w = rng.exponential(1.0, n)
        w[rng.integers(0, n, max(1, n // 10))] = 0.0
    elif kind == "spiky":                      # a few ancestors take nearly everything: long runs of equal indices
        w = rng.exponential(1.0, n) * 1e-6
        w[rng.integers(0, n, 5)] = 1.0
    else:                                      # empty stretches: whole tiles of zero weight
        w = np.zeros(n)
        w[n // 3: n // 3 + max(1, n // 50)] = rng.exponential(1.0, max(1, n // 50))
    return w / w.sum()


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("n, wkind", [(7, "exp"), (1024, "exp"), (1025, "spiky"), (300007, "exp"), (1 << 20, "spiky"),
                                      (200000, "gap")])
def test_fused_resample_equals_separate_kernels(sb, kind, n, wkind):
    from smcpy_b200 import engine, _lib
    w = _weights(n, n + kind, wkind)
    rng = np.random.default_rng(1)
    cols = rng.normal(0, 1, (5, n))
    res = {}
    for fused in (1, 0):
        st = _state(n, 3)
        st.w.copy_(engine.dev(w))
        st.cols.copy_(engine.dev(cols))

        class K:
            pass
        K.rng = engine.PhiloxRNG(99)
        rs = [sb.standard, sb.stratified, sb.systematic][kind]
        with _lib.option("no_fused_resample", 0 if fused else 1), warnings.catch_warnings(record=True) as wl:
            warnings.simplefilter("always")
            idx = engine.resample(st, K, rs, 0.01)
        res[fused] = (idx.cpu().numpy().copy(), st.cols.cpu().numpy().copy(), len(wl), st.cdf.cpu().numpy().copy())
    np.testing.assert_array_equal(res[1][3], res[0][3])
    np.testing.assert_array_equal(res[1][0], res[0][0])                 # ancestors bit-exact
    np.testing.assert_array_equal(res[1][1], res[0][1])                 # gathered rows
    assert res[1][2] == res[0][2]                                       # same collapse warning decision
    idx = res[1][0]
    assert np.all(np.diff(idx) >= 0) and idx.min() >= 0 and idx.max() < n
    np.testing.assert_array_equal(res[1][1], cols[:, idx])


def test_fused_resample_unique_count(sb):
    from smcpy_b200 import engine
    n = 300007
    w = _weights(n, 3, "spiky")
    st = _state(n, 2)
    st.w.copy_(engine.dev(w))

    class K:
        rng = engine.PhiloxRNG(5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        idx = engine.resample(st, K, sb.stratified, 0.01).cpu().numpy()
    assert int(st.counts[0].item()) == len(np.unique(idx))


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("world, n_local, wkind", [(2, 5000, "exp"), (4, 70001, "exp"), (8, 1 << 17, "spiky"),
                                                   (4, 50000, "gap"), (8, 16, "exp")])
def test_sharded_fused_resample_with_virtual_ranks(sb, kind, world, n_local, wkind):
    """`world` shards in ONE GPU's memory: per-shard CDF / running sums, gathered totals, then every
    virtual rank runs the owner-computes kernel and writes into the shard that holds each output slot.
    Against the single-shard kernel on the concatenated arrays: the same ancestors except where the
    last bit of a segment offset moves a slot across a neighbour boundary (sharded sums are taken in a
    different order), every output slot written exactly once, distinct-ancestor counts consistent."""
    import torch
    from smcpy_b200 import engine, _lib
    n = world * n_local
    ncol = 4
    w = _weights(n, 17 * world + kind, wkind)
    rng = np.random.default_rng(2)
    cols = np.vstack([np.arange(n, dtype=np.float64)[None, :], rng.normal(0, 1, (ncol - 1, n))])   # row 0 = global id
    seed, stream_id = 4242, 9
    dv = engine.Runtime.get().device
    ws = engine.Runtime.get().workspace(n, ncol)
    s = _lib.stream_ptr()
    # single shard
    w_all = engine.dev(w)
    cdf_all = torch.empty(n, dtype=torch.float64, device=dv)
    t_all = torch.empty(n, dtype=torch.float64, device=dv)
    src_all = engine.dev(cols)
    dst_all = torch.full((ncol, n), float("nan"), dtype=torch.float64, device=dv)
    cnt = torch.zeros(4, dtype=torch.int64, device=dv)
    idx_all = torch.empty(n, dtype=torch.int64, device=dv)
    _lib.call("smcb_cdf_scan", _lib.ptr(w_all), _lib.ptr(cdf_all), n, _lib.SCAN_LOOKBACK, _lib.ptr(ws), s)
    _lib.call("smcb_resample_exponential_sums", 0, n, n, seed, stream_id, _lib.ptr(t_all), None, _lib.ptr(ws), s)
    r = _lib.Resample()
    r.kind, r.world, r.rank, r.n_cols, r.n_local, r.n_global, r.seed, r.stream_id = kind, 1, 0, ncol, n, n, seed, stream_id
    r.cdf, r.totals, r.src, r.src_stride = cdf_all.data_ptr(), None, src_all.data_ptr(), n
    r.t_peer[0], r.dst_peer[0], r.dst_stride = t_all.data_ptr(), dst_all.data_ptr(), n
    r.idx_out, r.n_unique_out = idx_all.data_ptr(), cnt.data_ptr()
    _lib.call("smcb_resample_fused", C.byref(r), _lib.ptr(ws), s)
    want = dst_all.cpu().numpy()
    want_unique = int(cnt[0].item())
    np.testing.assert_array_equal(want[0], idx_all.cpu().numpy().astype(np.float64))
    # shards
    w_sh = [engine.dev(w[q * n_local:(q + 1) * n_local]) for q in range(world)]
    cdf_sh = [torch.empty(n_local, dtype=torch.float64, device=dv) for _ in range(world)]
    t_sh = [torch.empty(n_local, dtype=torch.float64, device=dv) for _ in range(world)]
    src_sh = [engine.dev(cols[:, q * n_local:(q + 1) * n_local]) for q in range(world)]
    dst_sh = [torch.full((ncol, n_local), float("nan"), dtype=torch.float64, device=dv) for _ in range(world)]
    totals = torch.zeros((world, 2), dtype=torch.float64, device=dv)
    for q in range(world):
        _lib.call("smcb_cdf_scan", _lib.ptr(w_sh[q]), _lib.ptr(cdf_sh[q]), n_local, _lib.SCAN_LOOKBACK, _lib.ptr(ws), s)
        _lib.call("smcb_resample_exponential_sums", q * n_local, n_local, n, seed, stream_id, _lib.ptr(t_sh[q]),
                  _lib.ptr(totals[q, 1:2]), _lib.ptr(ws), s)
        totals[q, 0] = cdf_sh[q][n_local - 1]
    uniq = []
    for q in range(world):
        r = _lib.Resample()
        r.kind, r.world, r.rank, r.n_cols, r.n_local, r.n_global, r.seed, r.stream_id = (kind, world, q, ncol, n_local, n,
                                                                                         seed, stream_id)
        r.cdf, r.totals, r.src, r.src_stride = cdf_sh[q].data_ptr(), totals.data_ptr(), src_sh[q].data_ptr(), n_local
        for p in range(world):
            r.t_peer[p], r.dst_peer[p] = t_sh[p].data_ptr(), dst_sh[p].data_ptr()
        r.dst_stride, r.idx_out, r.n_unique_out = n_local, None, cnt.data_ptr()
        _lib.call("smcb_resample_fused", C.byref(r), _lib.ptr(ws), s)
        uniq.append(int(cnt[0].item()))
    got = np.hstack([d.cpu().numpy() for d in dst_sh])
    assert not np.isnan(got).any()                                       # every output slot written
    anc, anc_want = got[0].astype(np.int64), want[0].astype(np.int64)
    assert np.all(np.diff(anc) >= 0)                                     # sorted uniforms -> sorted ancestors
    diff = anc != anc_want
    assert diff.mean() < 2e-4, diff.mean()
    assert np.all(np.abs(anc[diff] - anc_want[diff]) <= 2) or wkind == "gap"
    np.testing.assert_array_equal(got[1:], cols[1:, anc])                # rows travel with their ancestor id
    assert sum(uniq) == len(np.unique(anc))
    assert abs(sum(uniq) - want_unique) <= max(4, int(2e-4 * n))
    # offspring counts follow the weights (stratified / systematic: within the scheme's bound)
    if kind:
        counts = np.bincount(anc, minlength=n)
        assert np.max(np.abs(counts - n * w)) < 2.0 + 1e-6


# ------------------------------------------------------------------ whole steps through the samplers
def test_native_sampler_runs_fused_and_separate_paths_agree(sb):
    """FixedPhi and Adaptive runs of the C1 model with the fused step kernels against the same runs with
    every fused piece switched off (separate reweight launches, every midpoint evaluated, separate resampling
    kernels): identical Philox counters, so the phi sequence, the ancestors and the particles agree."""
    from smcpy_b200 import _lib
    problem = configs.linear_problem(100)
    out = {}
    for mode in ("fused", "separate"):
        opts = {} if mode == "fused" else {"no_coop": 1, "bisect_plain": 1, "no_fused_resample": 1}
        saved = [(k, _lib.set_option(k, v)) for k, v in opts.items()]
        try:
            res = []
            for sampler in ("adaptive", "fixed"):
                mcmc, kernel = make_b200(problem, rng=31)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    if sampler == "adaptive":
                        smc = sb.AdaptiveSampler(kernel, show_progress_bar=False)
                        steps, mll = smc.sample(30000, 4, target_ess=0.8, resample_rng=sb.stratified)
                    else:
                        smc = sb.FixedPhiSampler(kernel, show_progress_bar=False)
                        steps, mll = smc.sample(30000, 4, np.linspace(0, 1, 9) ** 3, 0.7, resample_rng=sb.standard)
                res.append((np.asarray(smc.phi_sequence, dtype=float), mll.copy(), steps[-1].params.copy(),
                            steps[-1].log_weights.copy()))
            out[mode] = res
        finally:
            for k, v in saved:
                _lib.set_option(k, v)
    for a, b in zip(out["fused"], out["separate"]):
        np.testing.assert_allclose(a[0], b[0], rtol=1e-10)   # phi sequence: same decisions (sums merge in another order)
        np.testing.assert_allclose(a[1], b[1], rtol=1e-12)   # mll (block-merge order of the sums differs)
        np.testing.assert_allclose(a[2], b[2], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(a[3], b[3], rtol=1e-9, atol=1e-12)
