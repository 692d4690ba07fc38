"""
Device-resident engine: the host side of the hot path
    reweight -> ESS / phi bisection -> resample -> covariance -> K x Metropolis
driving the C-ABI kernels of libsmcb200.so on torch device tensors (PyTorch is plumbing:
memory, streams, torch.distributed).  Every numerical step is a call into the library.

Particle store (structure of arrays, fp64): one (d+2, N) tensor `cols` per shard whose rows are
the d parameter columns, the log-likelihood and the summed log-prior, so a resample is one
gather over d+2 columns into the ping-pong twin.
"""

import ctypes as C
import math
import os
import warnings

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


# ---------------------------------------------------------------------------------------------
# runtime: device, workspace, process group
# ---------------------------------------------------------------------------------------------
class Runtime:
    _inst = None

    def __init__(self):
        _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._ws = None
        self._ws_bytes = 0

    @classmethod
    def get(cls):
        if cls._inst is None or cls._inst.device.index != torch.cuda.current_device():
            cls._inst = Runtime()
        return cls._inst

    def workspace(self, n, d):
        need = int(_lib.lib.smcb_workspace_bytes(int(n), int(d)))
        if need > self._ws_bytes:
            self._ws = torch.zeros(need, dtype=torch.uint8, device=self.device)
            self._ws_bytes = need
        return self._ws


class PeerExchange:
    """Symmetric-memory buffers for the collectives that are fused into kernels (one per process):
    * `slots`     -- accept-count exchange inside the Metropolis kernel (smcb_mh_step_xchg);
    * `bis_slots` -- rank sums exchanged inside the cooperative phi search and reweight kernels
                     (smcb_ess_search, smcb_reweight_fused);
    * `big_slots` -- small all-gather / all-reduce kernel (smcb_xchg_allgather: weight totals,
                     covariance sums, counts, error flags).
    get() returns None when symmetric memory cannot be set up (then NCCL collectives are used)."""
    _inst = None
    _failed = False

    def __init__(self):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        rank, world = dist_info()
        dv = Runtime.get().device
        self.rank, self.world = rank, world

        def sym(words):
            t = symm_mem.empty(int(words), dtype=torch.int64, device=dv)
            t.zero_()
            h = symm_mem.rendezvous(t, dist.group.WORLD)
            x = _lib.Xchg()
            x.world, x.rank = world, rank
            for r, p in enumerate(peer_pointers(t, h)):
                x.peer_slots[r] = p
            return t, h, x
        self.slots, self.handle, self.x = sym(_lib.lib.smcb_mh_xchg_slot_words())
        self.global_counts = torch.zeros(_lib.MAX_MCMC_STEPS, dtype=torch.int64, device=dv)
        self.ticket = torch.zeros(4, dtype=torch.int32, device=dv)
        self.x.global_counts = self.global_counts.data_ptr()
        self.x.ticket = self.ticket.data_ptr()
        self.gen = 0
        self.bis_slots, self.bis_handle, self.xb = sym(_lib.lib.smcb_xchg_slot_words())
        self.bis_gen = 0
        self.big_slots, self.big_handle, self.xg = sym(_lib.lib.smcb_xchg_big_slot_words())
        self.big_gen = 0
        self.max_vals = int(_lib.lib.smcb_xchg_max_vals())
        self.status = torch.zeros(1, dtype=torch.int32, device=dv)
        torch.cuda.synchronize()
        dist.barrier()                       # every rank's slots are zeroed before anyone publishes

    @classmethod
    def get(cls):
        if cls._failed or os.environ.get("SMCB_NO_XCHG"):
            return None
        _, world = dist_info()
        if world < 2 or world > _lib.MAX_RANKS:
            return None
        if cls._inst is None:
            try:
                cls._inst = PeerExchange()
            except Exception as e:            # pragma: no cover - depends on the box
                warnings.warn("symmetric-memory exchange unavailable (%s); using NCCL collectives" % e)
                cls._failed = True
                return None
        return cls._inst

    def next_bisect_gen(self):
        self.bis_gen += 1
        if self.bis_gen >= (1 << 48):
            raise RuntimeError("exchange generation overflow")
        self.xb.gen = self.bis_gen
        return self.xb

    def next_gen(self):
        self.gen += 1
        if self.gen >= (1 << 32) - 1:
            raise RuntimeError("accept-count exchange generation overflow")
        self.x.gen = self.gen
        return self.x

    def gather(self, src, reduce=False, out=None):
        """All-gather (rows in rank order) or all-reduce (rank-order sum) of a small float64 device
        vector through smcb_xchg_allgather; asynchronous on the current stream."""
        n = int(src.numel())
        if n > self.max_vals:
            raise ValueError("record too large for the peer exchange")
        if out is None:
            out = torch.empty(n if reduce else (self.world, n), dtype=torch.float64, device=src.device)
        self.big_gen += 1
        self.xg.gen = self.big_gen
        call("smcb_xchg_allgather", C.byref(self.xg), ptr(src), n, ptr(out), 1 if reduce else 0, ptr(self.status),
             stream_ptr())
        return out

    def check(self):
        if int(self.status.item()) != 0:
            raise RuntimeError("peer exchange: a rank never answered")


AcceptExchange = PeerExchange


def peer_pointers(t, handle):
    """Device pointers of tensor `t` in every rank's address space: buffer_ptrs are the bases of
    the symmetric allocation, the tensor may sit at an offset inside it."""
    rank, _ = dist_info()
    base = [int(p) for p in handle.buffer_ptrs]
    off = int(t.data_ptr()) - base[rank]
    if off < 0:
        raise RuntimeError("symmetric tensor lies outside its allocation")
    return [b + off for b in base]


class SymmetricPool:
    """Peer-mapped (symmetric-memory) particle buffers for the multi-GPU resampler, cached per
    shape so that repeated sample() calls do not rendezvous again.  get() returns
    [(tensor, [peer pointers], handle)] x 2 (the ping-pong pair) or None when symmetric memory is
    unavailable (then the NCCL all-to-all path is used).  A pair is LEASED to one live DeviceState:
    a second state of the same shape gets its own pair (every rank allocates in the same order)."""
    _cache = {}
    _failed = False

    @classmethod
    def get(cls, rows, n_max, owner=None):
        import weakref
        if cls._failed or os.environ.get("SMCB_NO_P2P"):
            return None
        key = (int(rows), int(n_max))
        leases = cls._cache.setdefault(key, [])
        for lease in leases:
            holder = lease[1]() if lease[1] is not None else None
            if holder is None or holder is owner:
                lease[1] = weakref.ref(owner) if owner is not None else None
                return lease[0]
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            dv = Runtime.get().device
            pair = []
            for _ in range(2):
                t = symm_mem.empty((rows, n_max), dtype=torch.float64, device=dv)
                h = symm_mem.rendezvous(t, dist.group.WORLD)
                pair.append((t, peer_pointers(t, h), h))
        except Exception as e:            # pragma: no cover - depends on the box
            warnings.warn("symmetric particle buffers unavailable (%s); using NCCL all-to-all" % e)
            cls._failed = True
            return None
        leases.append([pair, weakref.ref(owner) if owner is not None else None])
        return pair


_SOLO = False


class solo:
    """with solo(): ... -- run as a single-GPU job inside a multi-rank process (bench.py's N-GPU
    self-check compares the sharded run with a one-GPU run of the same global problem)."""

    def __enter__(self):
        global _SOLO
        self.old, _SOLO = _SOLO, True

    def __exit__(self, *exc):
        global _SOLO
        _SOLO = self.old


def dist_info():
    import torch.distributed as dist
    if not _SOLO and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_global, rank, world):
    """Contiguous block partition of global particle ids (np.array_split sizes, as
    smcpy/mcmc/parallel_vector_mcmc.py:38 splits the model inputs)."""
    base, rem = divmod(int(n_global), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PhiloxRNG:
    """Native counter-based generator: (seed, per-use stream counter).  Streams are consumed
    identically on every rank so multi-GPU runs stay in lock step."""

    def __init__(self, seed=None):
        if seed is None:
            seed = int(np.random.SeedSequence().entropy) & ((1 << 63) - 1)
            rank, world = dist_info()
            if world > 1:
                # one seed for the whole job: the sharded resampler and the global particle ids assume it
                import torch.distributed as dist
                box = [seed]
                dist.broadcast_object_list(box, src=0)
                seed = int(box[0])
        self.seed = int(seed) & ((1 << 64) - 1)
        self._stream = 0

    def next_stream(self):
        self._stream += 1
        return self._stream


def dev(a, dtype=torch.float64):
    """Host array -> device tensor (contiguous)."""
    rt = Runtime.get()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(rt.device, non_blocking=False)


def host(t):
    """Device tensor -> numpy array.  Arrays of 1-128 MB go through page-locked memory (torch caches
    the pinned blocks) at PCIe speed; the returned array owns that block.  Small reads and very
    large ones take the plain synchronous copy."""
    nbytes = t.numel() * t.element_size()
    if nbytes < (1 << 20) or nbytes > (128 << 20):      # very large reads stay pageable: locked pages are scarce
        return t.cpu().numpy()
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out.numpy()


def make_path(phi, phi_prev, lam):
    return _lib.Path(float(phi), float(phi_prev), float(lam))


# ---------------------------------------------------------------------------------------------
# problem specification
# ---------------------------------------------------------------------------------------------
class ProblemSpec:
    """Model + data + priors + likelihood bound to device buffers (the VectorMCMC constructor
    arguments, smcpy/mcmc/vector_mcmc.py:11-34, as an smcb_problem_t).

    fused = True: built-in device model, device priors, Normal / MultiSourceNormal / (linear, <= 3
    features) MVNormal -- the whole Metropolis step is one kernel.  Otherwise the un-fused route
    (callable_route.py): proposal kernel -> [host-evaluated priors] -> model / log-likelihood as torch
    callables on device tensors (or a registered CUDA model) -> accept kernel."""

    def __init__(self, model, data, priors, log_like_args, kind, loglike=None):
        from .models import DeviceModel, LinearModel, NvrtcModel
        from .priors import describe_prior
        rt = Runtime.get()
        self.model = model
        self.kind = kind
        self.priors = list(priors)
        self.loglike = loglike
        c = _lib.Problem()
        descs = [describe_prior(p) for p in self.priors]
        if len(descs) > _lib.MAX_PRIORS:
            raise NotImplementedError("at most %d priors" % _lib.MAX_PRIORS)
        d = 0
        n_cov = 0
        self.external = []                      # (prior object, first column, columns): host-evaluated log-pdf
        for i, (k, ncols, a, b, lpin) in enumerate(descs):
            c.priors[i] = _lib.Prior(k, ncols, a, b, lpin)
            if k == _lib.PRIOR_IMPROPER_COV:
                if n_cov < 2:
                    # lower Cholesky factor of the inverse-Wishart scale of ImproperCov.rvs (priors.py:118-123)
                    L = np.eye(3)
                    L[:ncols, :ncols] = np.linalg.cholesky(np.asarray(self.priors[i]._iw_scale, dtype=np.float64))
                    for e, (r, q) in enumerate(((0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2))):
                        c.iw_chol[n_cov][e] = float(L[r, q])
                n_cov += 1
                d += ncols * (ncols + 1) // 2
            elif k == _lib.PRIOR_EXTERNAL:
                self.external.append((self.priors[i], d, ncols))
                d += ncols
            else:
                d += 1
        if d > _lib.MAX_PARAMS:
            raise NotImplementedError("at most %d parameters" % _lib.MAX_PARAMS)
        c.n_priors = len(descs)
        c.n_params = d
        self.d = d
        self.n_priors = len(descs)
        self.n_cov = n_cov
        # summed log-prior of any particle inside the support when every prior is flat there (one add
        # chain in prior order, as build_col_priors does on the library side); None with host-evaluated priors
        self.lp_const = None
        if not self.external:
            self.lp_const = 0.0
            for (k, _, _, _, lpin) in descs:
                if k != _lib.PRIOR_IMPROPER_COV:
                    self.lp_const += float(lpin)
        data = np.asarray(data, dtype=np.float64)
        self._data_dev = dev(data.reshape(-1))
        c.data = self._data_dev.data_ptr()
        self._x_dev = None
        dev_model = isinstance(model, DeviceModel)
        fusable = dev_model and not self.external
        if kind == "normal":
            fusable = fusable and n_cov == 0
            c.loglike_id = _lib.LL_NORMAL
            c.n_data = int(data.reshape(-1).shape[0])
            c.n_snapshots = 1
            c.sigma_is_param = 1 if log_like_args is None else 0
            c.sigma = 0.0 if log_like_args is None else float(np.asarray(log_like_args).reshape(-1)[0])
            c.n_model_params = d - c.sigma_is_param
        elif kind == "multisource":
            if not (fusable and n_cov == 0):
                raise NotImplementedError("MultiSourceNormal needs a built-in device model and device priors")
            lens, stds = list(log_like_args[0]), list(log_like_args[1])
            if len(lens) != len(stds) or len(lens) > _lib.MAX_SEGMENTS:
                raise NotImplementedError("MultiSourceNormal on the device supports up to %d segments" % _lib.MAX_SEGMENTS)
            if sum(lens) != data.reshape(-1).shape[0]:
                raise ValueError("data segments in args[0] must sum to dim of data")
            c.loglike_id = _lib.LL_MULTISOURCE
            c.n_data = int(data.reshape(-1).shape[0])
            c.n_snapshots = 1
            n_none = stds.count(None)
            c.n_segments = len(lens)
            j = 0
            for i, (ln, sd) in enumerate(zip(lens, stds)):
                c.seg_len[i] = int(ln)
                if sd is None:
                    c.seg_sigma_col[i] = d - n_none + j
                    j += 1
                else:
                    c.seg_sigma_col[i] = -1
                    c.seg_sigma[i] = float(sd)
            c.n_model_params = d - n_none
        elif kind == "mvnormal":
            if data.ndim != 2:
                raise ValueError("MVNormal data must be (snapshots, features)")
            S, F = int(data.shape[0]), int(data.shape[1])
            args = list(log_like_args)
            if len(args) != F * (F + 1) // 2:
                raise ValueError("MVNormal args must hold F(F+1)/2 covariance terms")
            n_none = args.count(None)
            c.n_model_params = d - n_none
            fusable = fusable and isinstance(model, LinearModel) and F <= 3
            self.mvn_args, self.mvn_shape = args, (S, F)
            self._data2_dev = self._data_dev.reshape(S, F)
            if fusable:
                c.loglike_id = _lib.LL_MVNORMAL
                c.n_snapshots, c.n_data = S, F
                j = 0
                for i, a in enumerate(args):
                    if a is None:
                        c.cov_param_col[i] = d - n_none + j
                        j += 1
                    else:
                        c.cov_param_col[i] = -1
                        c.cov_fixed[i] = float(a)
                # snapshot mean + scatter matrix: all the kernels need of the S x F data
                self._mvn_stats = torch.zeros(_lib.MVN_STATS, dtype=torch.float64, device=rt.device)
                call("smcb_mvn_scatter", ptr(self._data_dev), S, F, ptr(self._mvn_stats), stream_ptr())
                c.mvn_stats = self._mvn_stats.data_ptr()
        elif kind == "torch":
            if loglike is None or not callable(getattr(loglike, "torch_loglike", None)):
                raise TypeError("a user log-likelihood needs a torch_loglike(params) method")
            fusable = False
            c.n_model_params = d
        else:
            raise NotImplementedError(kind)
        self.fused = bool(fusable)
        self.model_fn = None                    # un-fused route: (N, d_model) device tensor -> (N, M) device tensor
        if self.fused:
            c.model_id = model.model_id
            if model.n_model_params != c.n_model_params:
                raise ValueError("model takes %d parameters but the priors/likelihood leave %d"
                                 % (model.n_model_params, c.n_model_params))
            if model.model_id == _lib.MODEL_LINEAR:
                if model.n_out != c.n_data:
                    raise ValueError("linear model grid and data length differ")
                self._x_dev = dev(model.x)
                c.xgrid = self._x_dev.data_ptr()
                if kind == "normal":
                    # data centred about its means: what the linear + Normal kernels stage (smcb_linear_center)
                    self._lin_centered = torch.empty(2 + 2 * c.n_data, dtype=torch.float64, device=rt.device)
                    call("smcb_linear_center", ptr(self._x_dev), ptr(self._data_dev), c.n_data, ptr(self._lin_centered),
                         stream_ptr())
                    c.lin_centered = self._lin_centered.data_ptr()
            elif model.model_id == _lib.MODEL_SPRING_MASS:
                if model.n_out != c.n_data:
                    raise ValueError("spring-mass output count and data length differ")
                c.model_args[0], c.model_args[1] = model.x0, model.v0
                c.model_args[2], c.model_args[3] = model.dt, float(model.nsub)
            elif model.model_id == _lib.MODEL_IDENTITY:
                if model.n_out != c.n_data:
                    raise ValueError("identity model size and data length differ")
        else:
            if kind != "torch" and not (callable(model) or isinstance(model, (NvrtcModel, DeviceModel))):
                raise TypeError("model must be a smcpy_b200.models.DeviceModel, NvrtcModel or a torch callable")
            if isinstance(model, NvrtcModel):
                if kind != "normal":
                    raise NotImplementedError("registered CUDA models support the Normal likelihood only")
                if model.n_model_params != c.n_model_params:
                    raise ValueError("model takes %d parameters but the priors/likelihood leave %d"
                                     % (model.n_model_params, c.n_model_params))
                if model.n_out != int(data.reshape(-1).shape[0]):
                    raise ValueError("model output count and data length differ")
            elif dev_model:
                if model.n_model_params != c.n_model_params:
                    raise ValueError("model takes %d parameters but the priors/likelihood leave %d"
                                     % (model.n_model_params, c.n_model_params))
                self.model_fn = model.torch_eval
            else:
                self.model_fn = model
            if kind != "normal":
                # the kernels of the un-fused route only see the prior table; the likelihood is a torch callable
                c.loglike_id = _lib.LL_NORMAL
                c.n_data = 1
                c.n_snapshots = 1
            c.model_id = _lib.MODEL_IDENTITY      # unused by the un-fused kernels
        self.c = c
        self.device = rt.device

    @property
    def n_model_params(self):
        return int(self.c.n_model_params)

    @property
    def has_device_proposal(self):
        return bool(self.c.has_proposal)

    def set_proposal(self, proposal):
        """Binds the path's proposal (GeometricPath(proposal=...)): fills the device log q table
        when the proposal is describable, else leaves has_proposal = 0."""
        from .proposals import describe_proposal
        cols = describe_proposal(proposal) if proposal else None
        self.c.has_proposal = 0
        if cols is not None:
            if len(cols) != self.d:
                raise ValueError("proposal has %d components but there are %d parameters" % (len(cols), self.d))
            if not self.fused:
                raise NotImplementedError("proposals with torch-callable models are not on the device")
            for i, (k, a, b) in enumerate(cols):
                self.c.proposal[i] = _lib.ProposalCol(k, 0, a, b)
            self.c.has_proposal = 1


# ---------------------------------------------------------------------------------------------
# particle state
# ---------------------------------------------------------------------------------------------
class DeviceState:
    """Working set of one particle shard."""

    def __init__(self, n_local, n_global, d, id_offset=0, with_lq=False, symmetric=False):
        rt = Runtime.get()
        dv = rt.device
        self.n, self.n_global, self.d, self.id_offset = int(n_local), int(n_global), int(d), int(id_offset)
        f64 = torch.float64
        self.with_lq = bool(with_lq)            # third resident column: proposal log-pdf
        rows = d + 3 if with_lq else d + 2
        self.peer_cols = self.peer_cols_alt = None      # peer pointers of cols / cols_alt (multi-GPU)
        self.extra = self.extra_alt = None              # spare symmetric row after the columns (resampler scratch)
        self.peer_extra = self.peer_extra_alt = None
        self.col_stride = self.n
        pool = None
        if symmetric:
            _, world = dist_info()
            n_max = -(-self.n_global // world)
            pool = SymmetricPool.get(rows + 1, n_max, owner=self)
        if pool is not None:
            (full, peers, _), (full_alt, peers_alt, _) = pool
            self.col_stride = int(full.shape[1])
            if self.col_stride != self.n:
                raise NotImplementedError("multi-GPU runs need the particle count to be a multiple of the GPU count")
            self.cols, self.cols_alt = full[:rows], full_alt[:rows]
            self.extra, self.extra_alt = full[rows], full_alt[rows]
            self.peer_cols, self.peer_cols_alt = list(peers), list(peers_alt)
            off = rows * self.col_stride * 8
            self.peer_extra = [q + off for q in peers]
            self.peer_extra_alt = [q + off for q in peers_alt]
        else:
            self.cols = torch.empty((rows, self.n), dtype=f64, device=dv)
            self.cols_alt = torch.empty((rows, self.n), dtype=f64, device=dv)
        self.log_w = torch.empty(self.n, dtype=f64, device=dv)
        self.w = torch.empty(self.n, dtype=f64, device=dv)
        self.log_w_alt = torch.empty(self.n, dtype=f64, device=dv)
        self.w_alt = torch.empty(self.n, dtype=f64, device=dv)
        self._lq_host = None                    # host-evaluated proposal log-pdf (reweight only)
        self.cdf = torch.empty(self.n, dtype=f64, device=dv)
        self.u = torch.empty(self.n, dtype=f64, device=dv)
        self.idx = torch.empty(self.n, dtype=torch.int64, device=dv)
        self.moved = torch.zeros(self.n, dtype=torch.uint8, device=dv)
        self.accept_counts = torch.zeros(_lib.MAX_MCMC_STEPS, dtype=torch.int64, device=dv)
        self.flags = torch.zeros(4, dtype=torch.int32, device=dv)   # [0] error bits, [1..2] library tile counter/ticket
        self.scalars = torch.zeros(16, dtype=f64, device=dv)
        self.bis = torch.zeros(16, dtype=f64, device=dv)
        self.counts = torch.zeros(4, dtype=torch.int64, device=dv)
        self.zero = torch.zeros(1, dtype=f64, device=dv)
        self.mean = torch.zeros(d, dtype=f64, device=dv)
        self.cov = torch.zeros((d, d), dtype=f64, device=dv)
        self.chol = torch.zeros((d, d), dtype=f64, device=dv)
        self.sums = torch.zeros(1 + d + d * d, dtype=f64, device=dv)
        self.ws = rt.workspace(self.n, d)
        self.uniform_weights = False
        self.total_unnorm_log_weight = None
        self.shift = None                       # previous mean: shift of the one-pass covariance (d > 8)
        self.rs_tot = torch.zeros(2, dtype=f64, device=dv)   # running-sum totals of the resampler's scans
        # reweight of a step that resamples: the new weights exist only as (inputs, scalars) until the CDF
        # scan recomputes them; log_w / w still hold the PREVIOUS weights then
        self.pending = None
        self.lp_const = None                    # summed log-prior when it is one constant over the shard
        self.last_dphi = 0.0                    # previous adaptive step (start of the next phi search)
        self.deferred = []                      # device scalars to check at the end of the step

    @property
    def params(self):
        return self.cols[:self.d]

    @property
    def ll(self):
        return self.cols[self.d]

    @property
    def lp(self):
        return self.cols[self.d + 1]

    @property
    def lq(self):
        return self.cols[self.d + 2] if self.with_lq else self._lq_host

    @lq.setter
    def lq(self, value):
        self._lq_host = value

    def swap_cols(self):
        self.cols, self.cols_alt = self.cols_alt, self.cols
        self.peer_cols, self.peer_cols_alt = self.peer_cols_alt, self.peer_cols
        self.extra, self.extra_alt = self.extra_alt, self.extra
        self.peer_extra, self.peer_extra_alt = self.peer_extra_alt, self.peer_extra

    def set_params_host(self, params):
        """(n, d) host array -> SoA columns."""
        a = dev(np.asarray(params, dtype=np.float64))
        call("smcb_aos_to_soa", ptr(a), ptr(self.cols), self.n, self.d, self.n, stream_ptr())

    def params_host(self):
        out = torch.empty((self.n, self.d), dtype=torch.float64, device=self.cols.device)
        call("smcb_soa_to_aos", ptr(self.cols), ptr(out), self.n, self.d, self.n, stream_ptr())
        return host(out)

    def set_uniform_weights(self):
        """Weights after initialisation / resampling: log(1/N) normalised the reference's way
        (particles.py:141-150: exp(log((1/N)...)))."""
        n = self.n_global
        lw = float(np.log(1.0 / n))
        w = float(np.exp(lw))
        call("smcb_fill_f64", ptr(self.log_w), self.n, lw, stream_ptr())
        call("smcb_fill_f64", ptr(self.w), self.n, w, stream_ptr())
        self.uniform_weights = True
        self.pending = None

    def check_flags(self, where, allow_oob=False):
        f = int(self.flags[0].item())
        if f == 0:
            return 0
        self.flags[0] = 0
        if (f & _lib.FLAG_PRIOR_OOB) and not allow_oob:
            raise ValueError("Initial inputs are out of bounds; prior log prob = -inf (%s)" % where)
        if f & _lib.FLAG_MODEL_NAN:
            raise ValueError("nan in model output.")
        if f & _lib.FLAG_COV_NOT_PD:
            raise ValueError("MVNormal covariance is not positive definite for a particle with "
                             "finite prior (%s); add an ImproperCov prior" % where)
        if f & _lib.FLAG_XCHG_TIMEOUT:
            raise RuntimeError("a peer GPU never published its accept count (%s): the Metropolis exchange timed out" % where)
        return f


# ---------------------------------------------------------------------------------------------
# collectives (only when a process group exists)
# ---------------------------------------------------------------------------------------------
def _all_gather_rows(t):
    import torch.distributed as dist
    rank, world = dist_info()
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous())
    return out


def _all_reduce_sum(t):
    """In-place sum over ranks.  Small float64 records go through the peer-memory exchange kernel
    (rank-order add chain: bit-identical on every rank), everything else through NCCL."""
    px = PeerExchange.get()
    if px is not None and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= px.max_vals:
        src = t.clone()
        px.gather(src, reduce=True, out=t)
        return t
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


# ---------------------------------------------------------------------------------------------
# (a) reweight
# ---------------------------------------------------------------------------------------------
def _rw_inputs(st):
    """(lp pointer, lp_const, lw_in) of the reweight / search kernels: the summed log-prior column is
    skipped when it is one constant over the shard and the path has no proposal column."""
    lw_in = None if st.uniform_weights else st.log_w
    if st.lp_const is not None and st.lq is None:
        return None, float(st.lp_const), lw_in
    return st.lp, 0.0, lw_in


def reweight(st, path, ess_threshold=-1.0):
    """Updater._compute_new_weights + Particles normalisation (updater.py:97-133,
    particles.py:141-162) in one cooperative launch.  Returns (ess, total_unnorm_log_weight) after
    one host sync.  With ess_threshold >= 0 a step whose ESS falls below threshold * N does not
    materialise its weights (st.pending; resample() scans them straight from the columns)."""
    _, world = dist_info()
    lp, lpc, lw_in = _rw_inputs(st)
    s = stream_ptr()
    px = PeerExchange.get() if world > 1 else None
    fused = (world == 1 or px is not None) and not _lib.lib.smcb_get_option(b"no_coop")
    st.pending = None
    if fused:
        with _timed("reweight"):
            x = C.byref(px.next_bisect_gen()) if px is not None else None
            call("smcb_reweight_fused", st.n, ptr(st.ll), ptr(lp), lpc, ptr(st.lq), ptr(lw_in), st.n_global,
                 C.byref(path), float(ess_threshold), ptr(st.log_w_alt), ptr(st.w_alt), ptr(st.scalars), ptr(st.ws),
                 x, s)
    elif world == 1:
        with _timed("reweight"):
            call("smcb_reweight", st.n, ptr(st.ll), ptr(st.lp), ptr(st.lq), ptr(lw_in), st.n_global,
                 C.byref(path), ptr(st.log_w_alt), ptr(st.w_alt), ptr(st.scalars), ptr(st.ws), s)
    else:
        with _timed("reweight"):
            part = st.scalars[8:11]
            call("smcb_reweight_partials", st.n, ptr(st.ll), ptr(st.lp), ptr(st.lq), ptr(lw_in), st.n_global,
                 C.byref(path), ptr(part), ptr(st.ws), s)
            parts = _all_gather_rows(part)                       # rank-ordered (max, S1, S2)
            call("smcb_lse_merge", ptr(parts), world, ptr(st.scalars), s)
            call("smcb_reweight_finalize", st.n, ptr(st.ll), ptr(st.lp), ptr(st.lq), ptr(lw_in), st.n_global,
                 C.byref(path), ptr(st.scalars), ptr(st.log_w_alt), ptr(st.w_alt), s)
    sc = st.scalars[:8].cpu().numpy()                        # step-boundary sync
    if fused and sc[7] < 0:
        raise RuntimeError("fused multi-GPU reweight: a peer rank never answered")
    if fused and sc[6] != 1.0:
        st.pending = dict(path=path, lp=lp, lp_const=lpc, lw_in=lw_in)
    else:
        st.log_w, st.log_w_alt = st.log_w_alt, st.log_w
        st.w, st.w_alt = st.w_alt, st.w
        st.uniform_weights = False
    st.total_unnorm_log_weight = float(sc[3])
    # ll + every optional input column read once; log_w and w written only when materialised
    cols_in = 1 + (lp is not None if fused else 1) + (st.lq is not None) + (lw_in is not None)
    _note_bytes("reweight", st.n * (8.0 * cols_in * (1 if fused else 2) + (16.0 if st.pending is None else 0.0)))
    return float(sc[4]), float(sc[3])


def materialize_weights(st):
    """Writes the weights of a pending reweight (the step did not resample after all): the same
    launch again with the resampling test switched off.  Collective on several GPUs."""
    p = st.pending
    if p is None:
        return
    reweight(st, p["path"], ess_threshold=-1.0)


# ---------------------------------------------------------------------------------------------
# (b) adaptive phi
# ---------------------------------------------------------------------------------------------
def bisect_iters(phi_old):
    span = 1.0 - phi_old
    k = 3
    if span > 2e-12:
        k = 3 + int(math.ceil(math.log2(span / 2e-12)))
    return min(k, int(_lib.lib.smcb_ess_bisect_max_iters()))


def find_phi(st, spec, phi_old, target_ess, lam=1.0):
    """AdaptiveSampler.optimize_step's root search (samplers.py:250-259) on the device: one
    cooperative launch (smcb_ess_search), on several GPUs with the rank sums exchanged over NVLink
    inside it.  st.search_info = (passes over the particles, evaluations scipy would have made)."""
    _, world = dist_info()
    s = stream_ptr()
    flat = spec is not None and spec.lp_const is not None and st.lq is None
    lp = None if flat else st.lp
    lp_const = spec.lp_const if flat else 0.0
    px = PeerExchange.get() if world > 1 else None
    if world == 1 or (px is not None and not os.environ.get("SMCB_NO_BISECT_XCHG")):
        with _timed("bisect"):
            x = C.byref(px.next_bisect_gen()) if px is not None else None
            call("smcb_ess_search", st.n, ptr(st.ll), ptr(lp), ptr(st.lq), lp_const, float(phi_old), float(lam),
                 float(target_ess), st.n_global, float(st.last_dphi), ptr(st.bis), ptr(st.ws), x, s, launches=1)
    else:
        with _timed("bisect"):
            call("smcb_ess_bisect_begin", float(phi_old), ptr(st.bis), s)
            part = st.scalars[8:11]
            for _ in range(bisect_iters(phi_old)):
                call("smcb_ess_partial", st.n, ptr(st.ll), ptr(lp), ptr(st.lq), lp_const, float(phi_old), float(lam),
                     ptr(st.bis), ptr(part), ptr(st.ws), s)
                parts = _all_gather_rows(part)
                call("smcb_ess_bisect_update", ptr(parts), world, float(phi_old), float(target_ess), st.n_global,
                     ptr(st.bis), s)
    b = st.bis[:13].cpu().numpy()
    if b[2] < 0.0:
        raise RuntimeError("fused multi-GPU phi search: a peer rank never answered")
    if b[2] == 0.0:
        raise RuntimeError("device phi search did not converge")
    st.search_info = (int(b[10]), int(b[1]))
    _note_bytes("bisect", st.n * 8.0 * (1 + (lp is not None) + (st.lq is not None)) * max(1, int(b[10])))
    st.last_dphi = float(b[0]) - float(phi_old)
    return float(b[0])


def ess_margin(st, spec, phi_new, phi_old, target_ess, lam=1.0):
    """AdaptiveSampler.predict_ess_margin (samplers.py:264-276) for one candidate phi."""
    _, world = dist_info()
    s = stream_ptr()
    flat = spec is not None and spec.lp_const is not None and st.lq is None
    lp = None if flat else st.lp
    lp_const = spec.lp_const if flat else 0.0
    st.bis.zero_()
    st.bis[6] = float(phi_new)
    part = st.scalars[8:11]
    call("smcb_ess_partial", st.n, ptr(st.ll), ptr(lp), ptr(st.lq), lp_const, float(phi_old), float(lam),
         ptr(st.bis), ptr(part), ptr(st.ws), s)
    parts = _all_gather_rows(part) if world > 1 else part.reshape(1, 3)
    call("smcb_lse_merge", ptr(parts.contiguous()), world, ptr(st.scalars), s)
    sc = st.scalars[:3].cpu().numpy()
    ess = 0.0
    if np.isfinite(sc[0]) and sc[1] > 0 and sc[2] > 0:
        ess = sc[1] * sc[1] / sc[2]
    return ess - st.n_global * target_ess


# ---------------------------------------------------------------------------------------------
# (c) resample
# ---------------------------------------------------------------------------------------------
def resample_indices(st, u_dev, scan_mode):
    """idx = digitize(u, cumsum(w)) for the local shard (single GPU)."""
    s = stream_ptr()
    call("smcb_cdf_scan", ptr(st.w), ptr(st.cdf), st.n, int(scan_mode), ptr(st.ws), s,
         launches=1 if scan_mode == _lib.SCAN_SEQUENTIAL else 2)
    call("smcb_searchsorted", ptr(st.cdf), st.n, ptr(u_dev), st.n, ptr(st.idx), s)
    return st.idx


def count_unique(st, idx):
    call("smcb_count_unique", ptr(idx), st.n, st.n, ptr(st.counts), ptr(st.ws), stream_ptr())
    return st.counts[0]


def gather_particles(st, idx):
    call("smcb_gather", ptr(st.cols), st.n, ptr(st.cols_alt), st.n, ptr(idx), st.n, int(st.cols.shape[0]),
         stream_ptr())
    st.swap_cols()
    st.set_uniform_weights()


def weights_cdf(st, total_out=None):
    """st.cdf <- inclusive scan of the step's normalised weights: recomputed from the resident columns
    when the reweight left them pending (a resampling step never writes log_w / w), else from st.w."""
    s = stream_ptr()
    p = st.pending
    if p is not None:
        call("smcb_cdf_scan_weights", st.n, ptr(st.ll), ptr(p["lp"]), p["lp_const"], ptr(st.lq), ptr(p["lw_in"]),
             st.n_global, C.byref(p["path"]), ptr(st.scalars), ptr(st.cdf), ptr(total_out), ptr(st.ws), s)
    else:
        call("smcb_cdf_scan", ptr(st.w), ptr(st.cdf), st.n, _lib.SCAN_LOOKBACK, ptr(st.ws), s)
        if total_out is not None:
            total_out.copy_(st.cdf[st.n - 1:st.n])


def warn_collapsed(n_unique, n_global, warn_threshold):
    if n_unique <= max(1, warn_threshold * n_global):
        warnings.warn(f"Resampled to less than {warn_threshold * 100}% of particles; ", UserWarning)


def run_deferred(st):
    """Host-side checks whose device scalars were left unread during the step (one copy)."""
    if not st.deferred:
        return
    items, st.deferred = st.deferred, []
    vals = torch.stack([t.reshape(()).to(torch.float64) for t, _ in items]).cpu().numpy()
    for v, (_, fn) in zip(vals, items):
        fn(float(v))


def resample(st, kernel, resample_rng, warn_threshold, defer=False):
    """Updater._resample (updater.py:135-165).  Replay mode (numpy Generator): both uniform
    draws of the reference (quirk Q1) are made on the host, the sequential scan gives
    bit-exact indices.  Native mode: all built-in resamplers have sorted uniforms, so ONE fused pass
    evaluates them, searches, gathers and counts the distinct ancestors (smcb_resample_fused).
    defer=True leaves the collapse warning to run_deferred() (no host sync here)."""
    _, world = dist_info()
    host_rng = isinstance(kernel.rng, np.random.Generator)
    kind = getattr(resample_rng, "kind", None)
    if world > 1:
        from .sharded import resample_sharded
        with _timed("resample"):
            return resample_sharded(st, kernel, resample_rng, warn_threshold, defer)
    n = st.n
    with _timed("resample"):
        if host_rng or kind is None or _lib.lib.smcb_get_option(b"no_fused_resample"):
            materialize_weights(st)
            return _resample_local(st, kernel, resample_rng, warn_threshold, n, host_rng, kind)
        return _resample_fused(st, kernel, kind, warn_threshold, defer)


def _resample_fused(st, kernel, kind, warn_threshold, defer):
    rng = kernel.rng
    s = stream_ptr()
    stream_id = rng.next_stream()
    weights_cdf(st)
    r = _lib.Resample()
    if kind == _lib.RESAMPLE_STANDARD:
        # multinomial uniforms in sorted order: running sums of per-slot exponentials (st.u)
        call("smcb_resample_exponential_sums", st.id_offset, st.n, st.n_global, rng.seed, stream_id, ptr(st.u),
             None, ptr(st.ws), s)
        r.t_peer[0] = st.u.data_ptr()
    r.kind, r.world, r.rank, r.n_cols = int(kind), 1, 0, int(st.cols.shape[0])
    r.n_local, r.n_global, r.seed, r.stream_id = st.n, st.n_global, rng.seed, stream_id
    r.cdf, r.totals = st.cdf.data_ptr(), None
    r.src, r.src_stride = st.cols.data_ptr(), st.col_stride
    r.dst_peer[0], r.dst_stride = st.cols_alt.data_ptr(), st.col_stride
    r.idx_out, r.n_unique_out = st.idx.data_ptr(), st.counts.data_ptr()
    call("smcb_resample_fused", C.byref(r), ptr(st.ws), s)
    st.swap_cols()
    st.set_uniform_weights()
    n_global = st.n_global
    if defer:
        st.deferred.append((st.counts[0:1].clone(), lambda v: warn_collapsed(v, n_global, warn_threshold)))
    else:
        warn_collapsed(int(st.counts[0].item()), n_global, warn_threshold)
    return st.idx


def _resample_local(st, kernel, resample_rng, warn_threshold, n, host_rng, kind):
    if host_rng or kind is None:
        u1 = np.asarray(resample_rng(kernel, n), dtype=np.float64)
        idx = resample_indices(st, dev(u1), _lib.SCAN_SEQUENTIAL)
        idx_sel = idx.clone()
        u2 = np.asarray(resample_rng(kernel, n), dtype=np.float64)
        u2_dev = dev(u2)
        call("smcb_searchsorted", ptr(st.cdf), n, ptr(u2_dev), n, ptr(st.idx), stream_ptr())
        n_unique = count_unique(st, st.idx)
        idx = idx_sel
    else:
        rng = kernel.rng
        if kind == _lib.RESAMPLE_STANDARD and os.environ.get("SMCB_SORTED_STANDARD", "1") != "0":
            # multinomial uniforms generated in sorted order (exponential spacings): the ancestor
            # indices come out non-decreasing, so the CDF search and the row gather walk memory
            # front to back instead of at random
            from .sharded import sorted_uniforms
            sorted_uniforms(st, rng.seed, rng.next_stream())
        else:
            call("smcb_resample_uniforms", int(kind), st.n_global, 0, n, rng.seed, rng.next_stream(), ptr(st.u),
                 stream_ptr())
        idx = resample_indices(st, st.u, _lib.SCAN_LOOKBACK)
        n_unique = count_unique(st, idx)
    gather_particles(st, idx)
    warn_collapsed(int(n_unique.item()), st.n_global, warn_threshold)
    return idx


# ---------------------------------------------------------------------------------------------
# covariance + Cholesky
# ---------------------------------------------------------------------------------------------
def _host_chol_psd(cov):
    """vector_mcmc.py:216-236 fallback for a non-PD covariance (quirk Q9), on the host."""
    warnings.warn("Covariance matrix is not positive semi-definite; forcing negative eigenvalues "
                  "to zero and rebuilding covariance matrix.")
    eigval, eigvec = np.linalg.eigh(cov)
    eigval[eigval < 0] = 0
    cov = (eigvec @ np.diag(eigval)) @ eigvec.T
    cov += 1e-14 * np.eye(len(eigval))
    return np.linalg.cholesky(cov)


def covariance(st, two_pass=False):
    """Particles.compute_covariance (particles.py:189-198) + its Cholesky factor; leaves
    st.cov / st.chol on the device and returns the device covariance.  For d > 8 and a state that
    already has a mean (every SMC step after the first) the moments are taken in one pass about
    the previous mean (smcb_weighted_moments_shifted); cholesky() falls back to the two passes
    if the kernel reports that the shift was too far off."""
    _, world = dist_info()
    s = stream_ptr()
    d = st.d
    st._cov_one_pass = False
    if getattr(st, "pending", None) is not None:
        materialize_weights(st)
    if (not two_pass and d > 8 and st.n >= 148 * 128 and getattr(st, "shift", None) is not None
            and not _lib.lib.smcb_get_option(b"cov_two_pass")):
        try:
            with _timed("cov"):
                call("smcb_weighted_moments_shifted", ptr(st.params), st.col_stride if hasattr(st, "col_stride") else st.n,
                     ptr(st.w), st.n, d, ptr(st.shift), ptr(st.sums), ptr(st.ws), s)
                if world > 1:
                    _all_reduce_sum(st.sums)
                call("smcb_cov_from_shifted_sums", ptr(st.sums), ptr(st.shift), d, ptr(st.mean), ptr(st.cov),
                     ptr(st.flags), s)
            st._cov_one_pass = True
            return st.cov
        except NotImplementedError:
            pass
    stride = st.col_stride if hasattr(st, "col_stride") else st.n
    if world == 1:
        with _timed("cov"):
            call("smcb_weighted_cov", ptr(st.params), stride, ptr(st.w), st.n, d, ptr(st.mean), ptr(st.cov),
                 ptr(st.sums), ptr(st.ws), s)
    elif (PeerExchange.get() is not None and d * d <= PeerExchange.get().max_vals):
        # sharded: the two passes and their two all-reduces over NVLink peer memory, enqueued by one library call
        px = PeerExchange.get()
        if getattr(px, "cov_scratch", None) is None:
            px.cov_scratch = torch.empty(_lib.MAX_PARAMS * _lib.MAX_PARAMS + 1, dtype=torch.float64, device=st.cols.device)
        px.big_gen += 2
        px.xg.gen = px.big_gen - 1
        with _timed("cov"):
            call("smcb_weighted_cov_xchg", ptr(st.params), stride, ptr(st.w), st.n, d, ptr(st.mean), ptr(st.cov),
                 ptr(st.sums), ptr(st.ws), C.byref(px.xg), ptr(px.cov_scratch), ptr(px.status), s)
    else:
        with _timed("cov"):
            call("smcb_weighted_sums1", ptr(st.params), stride, ptr(st.w), st.n, d, ptr(st.sums), ptr(st.ws), s)
            _all_reduce_sum(st.sums[:1 + d])
            call("smcb_mean_from_sums", ptr(st.sums), d, ptr(st.mean), s)
            s2 = st.sums[1 + d:]
            call("smcb_weighted_sums2", ptr(st.params), stride, ptr(st.w), st.n, d, ptr(st.mean), ptr(s2),
                 ptr(st.ws), s)
            _all_reduce_sum(s2)
            call("smcb_cov_from_sums", ptr(st.sums), ptr(s2), d, ptr(st.cov), s)
    if hasattr(st, "flags") and d > 8:
        st.shift = st.mean.clone()           # the next call can take the one-pass route
    return st.cov


def cholesky(st, cov_dev=None, lazy=False):
    """Cholesky factor of the proposal covariance.  lazy=True (native generator): no host sync here --
    a failed factorisation leaves L = 0, under which the Metropolis steps are the identity, and
    mutate() repairs and repeats when it reads the flags at the end of the step."""
    call("smcb_cholesky", ptr(st.cov if cov_dev is None else cov_dev), st.d, ptr(st.chol), ptr(st.flags),
         stream_ptr())
    if lazy:
        return st.chol
    f = int(st.flags[0].item())
    if f & _lib.FLAG_COV_RESHIFT:
        # one-pass covariance taken about a shift too far from the mean: redo it in two passes
        st.flags[0] = f & ~(_lib.FLAG_COV_RESHIFT | _lib.FLAG_CHOL_FAIL)
        if cov_dev is None and getattr(st, "_cov_one_pass", False):
            covariance(st, two_pass=True)
            return cholesky(st)
        f &= ~_lib.FLAG_COV_RESHIFT
    if f & _lib.FLAG_CHOL_FAIL:
        st.flags[0] = f & ~_lib.FLAG_CHOL_FAIL
        cov = (st.cov if cov_dev is None else cov_dev).cpu().numpy()
        st.chol.copy_(dev(_host_chol_psd(cov)))
    return st.chol


# ---------------------------------------------------------------------------------------------
# (d) Metropolis
# ---------------------------------------------------------------------------------------------
def eval_incoming(st, spec, lp_cols=None):
    """log-priors + log-likelihood of the particles in st.cols (vector_mcmc.py:162-166)."""
    if spec.fused:
        if spec.has_device_proposal and not st.with_lq:
            raise ValueError("particle state has no proposal log-pdf column")
        call("smcb_mh_init", C.byref(spec.c), ptr(st.cols), st.n, st.n, ptr(st.ll), ptr(st.lp), ptr(lp_cols),
             ptr(st.lq) if spec.has_device_proposal else None, ptr(st.flags), stream_ptr())
    else:
        from .callable_route import eval_incoming_callable
        eval_incoming_callable(st, spec, lp_cols)


# optional kernel timing (bench.py): name -> list of (start, end) CUDA events on the launch stream;
# PROFILE_BYTES: name -> algorithmic HBM bytes of each of those calls on this rank (include/smcb200.h)
PROFILE = None
PROFILE_BYTES = {}


def _note_bytes(name, nbytes):
    if PROFILE is not None:
        PROFILE_BYTES.setdefault(name, []).append(float(nbytes))


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.b.record()
            PROFILE.setdefault(self.name, []).append((self.a, self.b))


def mutate(st, spec, num_samples, path, rng, check_incoming=False, lazy_chol=False):
    """VectorMCMC.smc_metropolis (vector_mcmc.py:46-62): `num_samples` fused Metropolis steps in
    place, proposal covariance rescaled from the global accept count after every step.
    st.chol must hold the Cholesky factor of the proposal covariance.  Returns the mutation ratio
    after the ONE host sync of the call, which also reads the error flags, the checks deferred during
    the step (st.deferred) and, with lazy_chol, the outcome of the factorisation."""
    if num_samples > _lib.MAX_MCMC_STEPS:
        raise ValueError("num_mcmc_samples > %d" % _lib.MAX_MCMC_STEPS)
    _, world = dist_info()
    s = stream_ptr()
    st.accept_counts[:max(1, num_samples)].zero_()
    st.moved.zero_()
    host_rng = isinstance(rng, np.random.Generator)
    r = _lib.Rng()
    r.id_offset = st.id_offset
    keep = []
    lq = ptr(st.lq) if (spec.fused and spec.has_device_proposal) else None
    xchg = PeerExchange.get() if (world > 1 and spec.fused) else None
    xs = xchg.next_gen() if xchg is not None else None
    if spec.fused and not host_rng and PROFILE is None and num_samples > 0:
        # native generator: the K launches are enqueued by ONE library call (smcb_mh_steps); per-step calls are kept
        # for the replay tapes and for per-launch timing
        r.mode, r.seed, r.stream = 0, rng.seed, rng._stream + 1
        rng._stream += num_samples
        call("smcb_mh_steps", C.byref(spec.c), C.byref(path), ptr(st.cols), st.n, st.n, st.n_global,
             ptr(st.ll), ptr(st.lp), lq, ptr(st.moved), ptr(st.chol), ptr(st.accept_counts), 0, num_samples,
             C.byref(r), ptr(st.flags), C.byref(xs) if xs is not None else None, s, launches=num_samples)
        num_steps_left = 0
    else:
        num_steps_left = num_samples
    for k in range(num_steps_left):
        if host_rng:
            z = dev(rng.normal(0, 1, (st.n, st.d)))
            u = dev(rng.uniform(0, 1, (st.n, 1)).reshape(-1))
            keep.append((z, u))
            r.mode, r.z, r.u = 1, z.data_ptr(), u.data_ptr()
        else:
            r.mode, r.seed, r.stream = 0, rng.seed, rng.next_stream()
        if spec.fused and xs is not None:
            with _timed("mh_step"):
                call("smcb_mh_step_xchg", C.byref(spec.c), C.byref(path), ptr(st.cols), st.n, st.n, st.n_global,
                     ptr(st.ll), ptr(st.lp), lq, ptr(st.moved), ptr(st.chol), ptr(st.accept_counts), k,
                     C.byref(r), ptr(st.flags), C.byref(xs), s)
            continue
        elif spec.fused:
            with _timed("mh_step"):
                call("smcb_mh_step", C.byref(spec.c), C.byref(path), ptr(st.cols), st.n, st.n, st.n_global,
                     ptr(st.ll), ptr(st.lp), lq, ptr(st.moved), ptr(st.chol), ptr(st.accept_counts), k,
                     C.byref(r), ptr(st.flags), s)
        else:
            from .callable_route import mh_step_callable
            mh_step_callable(st, spec, path, k, r)
        if world > 1:
            _all_reduce_sum(st.accept_counts[k:k + 1])
    call("smcb_count_moved", ptr(st.moved), st.n, ptr(st.counts[1:]), ptr(st.ws), s)
    # one record for the host: [moved, flag bits 0..4, deferred scalars...]; summed over the ranks so
    # that every rank raises (or none does)
    items, st.deferred = st.deferred, []
    extra = [t.reshape(1).to(torch.float64) for t, _ in items]
    if world > 1:
        bits = (st.flags[0].to(torch.int64) >> torch.arange(6, device=st.cols.device)) & 1
        pack = torch.cat([st.counts[1:2].to(torch.float64), bits.to(torch.float64)] + extra)
        _all_reduce_sum(pack)
        packed = pack.cpu().numpy()                          # the step's host sync
        flags = int(sum((1 << b) for b in range(6) if packed[1 + b] > 0))
        rest = packed[7:]
    else:
        packed = torch.cat([st.counts[1:2].to(torch.float64), st.flags[0:1].to(torch.float64)] + extra).cpu().numpy()
        flags = int(packed[1])
        rest = packed[2:]
    for v, (_, fn) in zip(rest, items):
        fn(float(v))
    if lazy_chol and (flags & (_lib.FLAG_CHOL_FAIL | _lib.FLAG_COV_RESHIFT)):
        st.flags[0] = flags & ~(_lib.FLAG_CHOL_FAIL | _lib.FLAG_COV_RESHIFT)
        if flags & _lib.FLAG_CHOL_FAIL:
            # the steps above ran with L = 0 (identity): factorise properly and repeat them
            if getattr(st, "_cov_one_pass", False):
                covariance(st, two_pass=True)
            cholesky(st)
            return mutate(st, spec, num_samples, path, rng, check_incoming, lazy_chol=False)
        st.shift = None                                      # one-pass covariance drifted: two passes next time
        flags &= ~(_lib.FLAG_CHOL_FAIL | _lib.FLAG_COV_RESHIFT)
    if flags:
        st.flags[0] = flags
        st.check_flags("smc_metropolis")
    return float(packed[0]) / st.n_global


# ---------------------------------------------------------------------------------------------
# host-array wrappers of the path arithmetic (GeometricPath.logpdf / inc_log_weights)
# ---------------------------------------------------------------------------------------------
def _path_eval(ll, lp, lq, path, mode):
    n = ll.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=Runtime.get().device)
    a, b = dev(ll.reshape(-1)), dev(lp.reshape(-1))
    c = None if lq is None else dev(lq.reshape(-1))
    call("smcb_path_eval", n, ptr(a), ptr(b), ptr(c), C.byref(path), mode, ptr(out), stream_ptr())
    return host(out).reshape(-1, 1)


def host_path_logpdf(ll, lp, lq, phi, lam):
    return _path_eval(ll, lp, lq, make_path(phi, phi, lam), 0)


def host_inc_log_weights(ll, lp, lq, phi, phi_prev, lam):
    return _path_eval(ll, lp, lq, make_path(phi, phi_prev, lam), 1)
