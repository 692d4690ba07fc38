"""
Multi-process tests.
  * world_size-2 gloo tests on CPU (run everywhere): the host-side bookkeeping of the sharded
    resampler (segment offsets, routing counts, all-to-all split exchange, row conservation).
  * world_size-2 NCCL test on GPUs (marked gpu, skipped with < 2 devices): a sharded
    FixedPhiSampler run agrees statistically with the single-GPU run and conserves particles.
"""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from smcpy_b200 import engine, sharded
    assert engine.dist_info() == (rank, world)
    n_global = 1001
    lo, hi = engine.shard_bounds(n_global, rank, world)
    n = hi - lo
    rng = np.random.default_rng(5)
    w_all = rng.exponential(1.0, n_global)
    w_all /= w_all.sum()
    w = torch.tensor(w_all[lo:hi])
    # 2: rank totals -> identical segment offsets on every rank
    tot = torch.empty(world, dtype=torch.float64)
    dist.all_gather_into_tensor(tot, torch.cumsum(w, 0)[-1:].clone())
    off = sharded.segment_offsets(tot)
    assert off[0] == 0 and abs(float(off[-1]) - 1.0) < 1e-12
    # 3: stratified uniforms of my slots, routed by segment
    u = (torch.arange(lo, hi, dtype=torch.float64) + torch.tensor(np.random.default_rng(rank).uniform(0, 1, n))) / n_global
    bounds = torch.cat([off[1:world], torch.tensor([float("inf")], dtype=torch.float64)])
    dest = torch.searchsorted(bounds, u, right=True)
    send = sharded.route_counts(dest, world)
    assert torch.equal(send, sharded.route_counts_sorted(u, bounds))
    recv = sharded.exchange_counts(send)
    assert int(send.sum()) == n
    u_in = torch.empty(int(recv.sum()), dtype=torch.float64)
    dist.all_to_all_single(u_in, u.contiguous(), output_split_sizes=recv.tolist(), input_split_sizes=send.tolist())
    # every request that reached me lies in my CDF segment
    assert bool(((u_in >= off[rank]) & (u_in < bounds[rank])).all())
    # 4-5: answer with (global ancestor id) rows and send them back
    cdf = torch.cumsum(w, 0)
    idx = torch.clamp(torch.searchsorted(cdf + off[rank], u_in, right=True), max=n - 1)
    rows = (idx + lo).to(torch.float64).reshape(-1, 1).repeat(1, 3)
    back = torch.empty((n, 3), dtype=torch.float64)
    dist.all_to_all_single(back, rows.contiguous(), output_split_sizes=send.tolist(), input_split_sizes=recv.tolist())
    # compare with the single-process answer on the full arrays
    want = np.minimum(np.searchsorted(np.cumsum(w_all), u.numpy(), side="right"), n_global - 1)
    got = back[:, 0].numpy().astype(np.int64)
    mism = int((got != want).sum())
    tot_req = torch.tensor([float(recv.sum())])
    dist.all_reduce(tot_req)
    # sorted multinomial uniforms by exponential spacings, sharded (sharded.sorted_uniforms with the
    # kernels replaced by torch ops): per-slot exponentials -> local running sums -> rank offsets
    e_all = torch.tensor(np.random.default_rng(99).exponential(1.0, n_global + 1)) * (0.5 / n_global)
    t_loc = torch.cumsum(e_all[lo:hi], 0)
    tots = torch.empty(world, dtype=torch.float64)
    dist.all_gather_into_tensor(tots, t_loc[-1:].clone())
    offs = sharded.segment_offsets(tots)
    u_sorted = (offs[rank] + t_loc) / (offs[world] + e_all[n_global])
    n_max = -(-n_global // world)                       # equal-size slots for the gather, trimmed after
    padded = torch.full((n_max,), float("nan"), dtype=torch.float64)
    padded[:n] = u_sorted
    allp = torch.empty(world * n_max, dtype=torch.float64)
    dist.all_gather_into_tensor(allp, padded)
    sizes = [engine.shard_bounds(n_global, r, world)[1] - engine.shard_bounds(n_global, r, world)[0] for r in range(world)]
    u_glob = np.concatenate([allp[r * n_max:r * n_max + sizes[r]].numpy() for r in range(world)])
    want_u = np.cumsum(e_all.numpy()[:n_global]) / e_all.numpy().sum()
    assert np.all(np.diff(u_glob) >= 0) and u_glob[0] > 0 and u_glob[-1] < 1
    np.testing.assert_allclose(u_glob, want_u, rtol=1e-12)
    out.put((rank, n, mism, float(tot_req.item())))
    dist.destroy_process_group()


def test_sharded_resampling_bookkeeping_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[1] for r in res) == 1001
    for rank, n, mism, tot_req in res:
        assert tot_req == 1001.0                 # every slot is answered exactly once
        assert mism <= 1                         # sharded offsets may move an edge by one ulp


def _fixed_time_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import time
    from scipy import stats
    import smcpy_b200 as sb
    from smcpy_b200 import samplers
    mc = sb.B200VectorMCMC(sb.LinearModel(np.arange(3.0)), np.zeros(3), [stats.uniform(0, 1)] * 2, 1.0)
    kernel = sb.VectorMCMCKernel(mc, ["a", "b"], rng=np.random.default_rng(0))
    smc = sb.FixedTimeSampler(kernel, time=0.5, show_progress_bar=False)
    # the device work is replaced by sleeps of different length per rank: local clocks disagree
    samplers.AdaptiveSampler._do_smc_step = lambda self, phi, num_mcmc_samples: time.sleep(0.03 * (1 + 2 * rank))
    samplers.AdaptiveSampler.optimize_step = lambda self, particles, phi_old, target_ess=1: min(1.0, phi_old + 0.05)
    smc._start_time = time.time()
    phis = [0.0]
    while phis[-1] < 1 and len(phis) < 100:
        phi = smc.optimize_step(None, phis[-1], 0.8)
        smc._do_smc_step(phi, 1)
        phis.append(phi)
    assert samplers.rank_zero_value(rank) == 0
    out.put((rank, phis, list(smc._time_history), smc._buffer_phi))
    dist.destroy_process_group()


def test_fixed_time_sampler_ranks_follow_rank_zero_clock_gloo():
    """FixedTimeSampler decides phi from wall-clock time; under one process per GPU every rank must
    take rank 0's decisions (smcpy/smc/samplers.py:304-379 + utils/mpi_utils.py rank_zero_output_only)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fixed_time_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict((r[0], r[1:]) for r in (q.get(timeout=120) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][0] == res[1][0] and res[0][0][-1] == 1        # same phi sequence, ended by the time budget
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]    # same time history and buffer phi
    assert len(res[0][0]) < 21                                  # jumped to 1 before the 0.05 steps got there


def _nccl_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import warnings
    import smcpy_b200 as sb
    from tests.golden import configs
    from tests.helpers import make_b200
    res = {}
    for resampler in (sb.stratified, sb.standard, sb.systematic):
        problem = configs.linear_problem(100)
        mcmc, kernel = make_b200(problem, rng=7)
        smc = sb.AdaptiveSampler(kernel, show_progress_bar=False)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            steps, mll = smc.sample(20000, 8, target_ess=0.8, resample_rng=resampler)
        last = steps[-1]
        res[resampler.__name__] = (float(mll[-1]), last.compute_mean(package=False), last.num_particles,
                                   len(steps), float(last.compute_ess()))
    # fused bisection (partials exchanged over NVLink inside one cooperative launch) against the
    # all-gather-per-evaluation fallback: same search, per-rank partials merged in the same order
    phis = {}
    for mode in ("fused", "nccl"):
        if mode == "nccl":
            os.environ["SMCB_NO_BISECT_XCHG"] = "1"
        mcmc, kernel = make_b200(configs.linear_problem(100), rng=11)
        smc = sb.AdaptiveSampler(kernel, show_progress_bar=False)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            steps, mll = smc.sample(30000, 3, target_ess=0.8)
        phis[mode] = (np.asarray(smc.phi_sequence, dtype=np.float64), float(mll[-1]))
    os.environ.pop("SMCB_NO_BISECT_XCHG", None)
    res["bisect"] = phis
    out.put((rank, res))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_sampler_two_gpus_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in (0, 1):
        f, g = res[r]["bisect"]["fused"], res[r]["bisect"]["nccl"]
        assert len(f[0]) == len(g[0])
        assert abs(f[0][1] - g[0][1]) < 1e-11                 # first step: identical particles, same root
        np.testing.assert_allclose(f[0], g[0], rtol=1e-3)
        assert abs(f[1] - g[1]) < 0.05
    np.testing.assert_array_equal(res[0]["bisect"]["fused"][0], res[1]["bisect"]["fused"][0])
    for name in ("stratified", "standard", "systematic"):
        a, b = res[0][name], res[1][name]
        assert a[0] == b[0] and a[3] == b[3]                  # identical scalars on every rank
        np.testing.assert_array_equal(a[1], b[1])
        assert a[2] + b[2] == 20000                           # particles conserved
        assert a[0] == pytest.approx(-77.357884, abs=0.3)     # SURVEY appendix B anchors
        np.testing.assert_allclose(a[1], [2.0363965, 3.4859536], atol=0.01)
