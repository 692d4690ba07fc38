"""
Reference arm plumbing for bench.py: runs the UNMODIFIED nasa/SMCPy package (pip-installed under
baseline/_ref, git-ignored, travels to the GPU box) through its own public API --
smcpy.ParallelVectorMCMC + VectorMCMCKernel + FixedPhiSampler / AdaptiveSampler -- on the host cores.

mpi4py / mpirun are not in the image, so the communicator ParallelVectorMCMC and the rank_zero_* decorators
need (smcpy/mcmc/parallel_vector_mcmc.py:20-47, smcpy/utils/mpi_utils.py: Get_size, Get_rank, scatter,
allgather, bcast, Barrier) is provided over forked processes: control messages and small objects travel
pickled on pipes through rank 0 (as mpi4py's lower-case methods pickle Python objects); the one large
payload of the path -- the all-gather of the model outputs, (N, M) doubles per evaluation -- goes through
a shared-memory arena so that it costs one local copy per rank, as a shared-memory MPI transport would.
Every rank runs the whole sampler (SPMD, as under mpirun); only the model evaluations are split.
Reported as "fork shim", never as mpi4py.  Nothing here is used by the GPU arm.
"""

import multiprocessing as mp
import os
import sys
import time
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def import_smcpy():
    """The installed reference package, or None.  h5py (absent from the image) is only needed by
    HDF5Storage and is stubbed."""
    if not os.path.isdir(os.path.join(REF, "smcpy")):
        return None
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import smcpy
    return smcpy


ARENA_DOUBLES = 1 << 24          # 128 MiB of shared doubles for the all-gather of model outputs


class ForkComm:
    def __init__(self, rank, size, conns, arena=None):
        self._rank, self._size, self._conns = rank, size, conns      # rank 0: conns[r] to rank r; else conns[0] to rank 0
        self._arena = None if arena is None else np.frombuffer(arena, dtype=np.float64)

    def Get_rank(self):
        return self._rank

    def Get_size(self):
        return self._size

    def scatter(self, sendobj, root=0):
        if self._rank == 0:
            for r in range(1, self._size):
                self._conns[r].send(sendobj[r])
            return sendobj[0]
        return self._conns[0].recv()

    def _small_allgather(self, obj):
        if self._rank == 0:
            objs = [obj] + [self._conns[r].recv() for r in range(1, self._size)]
            for r in range(1, self._size):
                self._conns[r].send(objs)
            return objs
        self._conns[0].send(obj)
        return self._conns[0].recv()

    def allgather(self, obj):
        if self._arena is not None and isinstance(obj, np.ndarray) and obj.dtype == np.float64:
            # shapes through the pipes (this round also fences the previous use of the arena), payload
            # through shared memory, a second small round as the "all written" barrier
            shapes = self._small_allgather(obj.shape)
            sizes = [int(np.prod(sh)) for sh in shapes]
            if sum(sizes) <= self._arena.size:
                offs = np.concatenate([[0], np.cumsum(sizes)])
                lo = int(offs[self._rank])
                self._arena[lo:lo + sizes[self._rank]] = obj.reshape(-1)
                self._small_allgather(None)
                return [self._arena[int(offs[r]):int(offs[r]) + sizes[r]].reshape(shapes[r]) for r in range(self._size)]
            self._small_allgather(None)
        if self._rank == 0:
            objs = [obj] + [self._conns[r].recv() for r in range(1, self._size)]
            for r in range(1, self._size):
                self._conns[r].send(objs)
            return objs
        self._conns[0].send(obj)
        return self._conns[0].recv()

    def bcast(self, obj, root=0):
        if self._rank == 0:
            for r in range(1, self._size):
                self._conns[r].send(obj)
            return obj
        return self._conns[0].recv()

    def Barrier(self):
        self.allgather(None)


def _child(rank, size, conn, arena, fn, args):
    try:
        fn(ForkComm(rank, size, {0: conn}, arena), *args)
    finally:
        conn.close()
        os._exit(0)


def run_spmd(fn, size, *args):
    """fn(comm, *args) on `size` forked ranks; returns rank 0's result."""
    if size <= 1:
        from smcpy.utils.single_rank_comm import SingleRankComm
        return fn(SingleRankComm(), *args)
    ctx = mp.get_context("fork")
    conns, procs = {}, []
    arena = ctx.RawArray("d", ARENA_DOUBLES)
    for r in range(1, size):
        a, b = ctx.Pipe()
        conns[r] = a
        p = ctx.Process(target=_child, args=(r, size, b, arena, fn, args), daemon=True)
        p.start()
        b.close()
        procs.append(p)
    try:
        return fn(ForkComm(0, size, conns, arena), *args)
    finally:
        for c in conns.values():
            c.close()
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()


# ---- workloads (same model / data / priors / sampler arguments as the GPU arm) -------------------------------
def _c2_fixed(comm, x, data, n, K, phi, ess_threshold, seed):
    import smcpy
    from scipy import stats
    from smcpy.resampler_rngs import standard
    model = lambda th: th[:, 0:1] * x[None, :] + th[:, 1:2]
    priors = [stats.uniform(-6.0, 12.0), stats.uniform(-6.0, 12.0)]
    if comm.Get_size() > 1 or type(comm).__name__ == "ForkComm":
        mcmc = smcpy.ParallelVectorMCMC(model, data, priors, comm, 0.5)
    else:
        mcmc = smcpy.VectorMCMC(model, data, priors, 0.5)
    kernel = smcpy.VectorMCMCKernel(mcmc, param_order=("a", "b"), rng=np.random.default_rng(seed))
    smc = smcpy.FixedPhiSampler(kernel, show_progress_bar=False)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steps, mll = smc.sample(n, K, phi, ess_threshold, resample_rng=standard)
    return time.perf_counter() - t0, float(mll[-1])


def c2_fixed_run(x, data, n, K, phi, ess_threshold, seed, ranks):
    """One FixedPhi pass of config C2 at n particles on `ranks` host processes -> (seconds, mll)."""
    return run_spmd(_c2_fixed, ranks, np.asarray(x), np.asarray(data), int(n), int(K), np.asarray(phi),
                    float(ess_threshold), int(seed))


def _adaptive(comm, build, n, K, target, seed):
    import smcpy
    from smcpy.resampler_rngs import standard
    built = build()
    model, data, priors, std, names = built[:5]
    # optional sixth item: the name of a reference log-likelihood class (default Normal); priors given as
    # ("improper_cov", n) are built from the reference's own prior classes here
    extra = {}
    if len(built) > 5:
        from smcpy import log_likelihoods
        extra["log_like_func"] = getattr(log_likelihoods, built[5])
    from smcpy.priors import ImproperCov
    priors = [ImproperCov(p[1], dof=5, S=np.eye(p[1])) if isinstance(p, tuple) and p[0] == "improper_cov" else p
              for p in priors]
    if type(comm).__name__ == "ForkComm":
        mcmc = smcpy.ParallelVectorMCMC(model, data, priors, comm, std, **extra)
    else:
        mcmc = smcpy.VectorMCMC(model, data, priors, std, **extra)
    kernel = smcpy.VectorMCMCKernel(mcmc, param_order=names, rng=np.random.default_rng(seed))
    smc = smcpy.AdaptiveSampler(kernel, show_progress_bar=False)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steps, mll = smc.sample(n, K, target, resample_rng=standard)
    return time.perf_counter() - t0, float(mll[-1]), len(smc.phi_sequence) - 1


def adaptive_run(build, n, K, target, seed, ranks):
    """AdaptiveSampler.sample to phi = 1 -> (seconds, mll, SMC steps); build() -> (model, data, priors, std, names
    [, name of the reference log-likelihood class])."""
    return run_spmd(_adaptive, ranks, build, int(n), int(K), float(target), int(seed))
