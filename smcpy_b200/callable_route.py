"""
Un-fused Metropolis route (BASELINE.json north_star: "User models plug in either as a torch callable on
device tensors or as a registered __device__ model"), also taken whenever something on the path has no
fused kernel:
  * torch-callable user models, and built-in device models next to host-evaluated priors;
  * priors evaluated by their own logpdf on the host (InvWishart, ImproperConstrainedUniform, any object
    with logpdf: SMCB_PRIOR_EXTERNAL) and ImproperCov next to a Normal likelihood;
  * MVNormal beyond the fused case (any model, any number of features), ApproxHierarch and user
    likelihood classes with a torch_loglike(params) method.
Per step: smcb_mh_propose -> [smcb_log_priors + host prior columns] -> model / log-likelihood ->
smcb_mh_accept.  A torch-callable model pays 8*M bytes per particle for its outputs, which the fused
built-ins avoid; a registered CUDA source (models.NvrtcModel) keeps them in registers
(smcb_user_loglike).  The proposal, the acceptance and the device priors are the library's kernels; the
torch callables are the user's code (and, for the general MVNormal / hierarchical likelihoods, batched
torch.linalg calls) -- a coverage route, not the measured hot path.
"""

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .models import NvrtcModel


def _check_output(out, n, m):
    if not (torch.is_tensor(out) and out.is_cuda and out.dtype == torch.float64):
        raise TypeError("torch-callable models must return a float64 CUDA tensor (N, M)")
    out = out.contiguous()
    if out.shape != (n, m):
        raise ValueError("model output shape %s != (%d, %d)" % (tuple(out.shape), n, m))
    return out


def mvnormal_torch(spec, aos):
    """MVNormal.__call__ (smcpy/log_likelihoods.py:150-192) for any model and feature count, batched over
    the particles: per-particle covariance from the packed upper-triangular terms (fixed, or sampled as the
    trailing columns), one Cholesky each, and the snapshot sum through the mean / scatter matrix of the
    data (as the fused kernel does).  Returns (ll (N,), not_pd (N,) bool)."""
    S, F = spec.mvn_shape
    args = spec.mvn_args
    n, d = aos.shape
    n_none = args.count(None)
    terms = torch.empty((n, len(args)), dtype=torch.float64, device=aos.device)
    j = 0
    for i, a in enumerate(args):
        if a is None:
            terms[:, i] = aos[:, d - n_none + j]
            j += 1
        else:
            terms[:, i] = float(a)
    r, c = torch.triu_indices(F, F, device=aos.device)
    cov = torch.zeros((n, F, F), dtype=torch.float64, device=aos.device)
    cov[:, r, c] = terms
    cov[:, c, r] = terms
    out = _check_output(spec.model_fn(aos[:, :d - n_none].contiguous()), n, F)
    data = spec._data2_dev
    ybar = data.mean(dim=0)
    dev_ = data - ybar
    W = dev_.t() @ dev_
    L, info = torch.linalg.cholesky_ex(cov)
    not_pd = info > 0
    L = torch.where(not_pd[:, None, None], torch.eye(F, dtype=torch.float64, device=aos.device), L)
    y = torch.linalg.solve_triangular(L, (out - ybar).unsqueeze(-1), upper=False).squeeze(-1)
    quad = S * (y * y).sum(dim=1) + torch.cholesky_solve(W.expand(n, F, F), L).diagonal(dim1=1, dim2=2).sum(dim=1)
    logdet = 2.0 * torch.log(L.diagonal(dim1=1, dim2=2)).sum(dim=1)
    ll = S * (-F / 2.0 * math.log(2 * math.pi) - 0.5 * logdet) - 0.5 * quad
    return ll, not_pd


def _model_ll(st, spec, soa_params, lp, ll_out):
    """ll_out[i] = log-likelihood of params_i; rows with lp = -inf keep the placeholder 0 and are not
    counted in the NaN check (vector_mcmc.py:205-214)."""
    c = spec.c
    n, d = st.n, st.d
    if spec.kind == "normal":
        dm = int(c.n_model_params)
        sigma_col = soa_params[d - 1] if c.sigma_is_param else None
        if isinstance(spec.model, NvrtcModel):                     # registered __device__ model: one kernel
            m = spec.model
            call("smcb_user_loglike", m.handle, ptr(soa_params), soa_params.stride(0), n, int(c.n_data),
                 C.c_void_p(int(c.data)), ptr(m.x_dev()), ptr(sigma_col), float(c.sigma), ptr(lp), ptr(ll_out),
                 ptr(st.flags), stream_ptr())
            return
        aos = soa_params[:dm].t().contiguous()                     # (n, d_model) for the user callable
        out = _check_output(spec.model_fn(aos), n, int(c.n_data))
        call("smcb_normal_loglike_rows", ptr(out), n, int(c.n_data), C.c_void_p(int(c.data)), ptr(sigma_col),
             float(c.sigma), ptr(lp), ptr(ll_out), ptr(st.flags), stream_ptr())
        return
    aos = soa_params.t().contiguous()                              # (n, d): all parameter columns
    live = ~(torch.isinf(lp) & (lp < 0))
    if spec.kind == "mvnormal":
        ll, not_pd = mvnormal_torch(spec, aos)
        if bool((not_pd & live).any().item()):
            st.flags[0] |= _lib.FLAG_COV_NOT_PD
    else:
        ll = spec.loglike.torch_loglike(aos)
        if not (torch.is_tensor(ll) and ll.is_cuda and ll.dtype == torch.float64 and ll.numel() == n):
            raise TypeError("torch_loglike must return a float64 CUDA tensor with one value per particle")
    ll = ll.reshape(-1)
    if bool((torch.isnan(ll) & live).any().item()):
        st.flags[0] |= _lib.FLAG_MODEL_NAN
    ll_out.copy_(torch.where(live, ll, torch.zeros_like(ll)))


def log_priors_into(st, spec, soa_params, lp_out, lp_cols=None):
    """lp_out <- summed log-priors of the rows in soa_params: device prior table (smcb_log_priors), plus
    the host-evaluated priors (one D2H of their columns, the object's own logpdf, one H2D)."""
    call("smcb_log_priors", C.byref(spec.c), ptr(soa_params), soa_params.stride(0), st.n, ptr(lp_out), ptr(lp_cols),
         stream_ptr())
    for prior, c0, dim in spec.external:
        block = soa_params[c0:c0 + dim].t().contiguous().cpu().numpy()
        v = np.asarray(prior.logpdf(block), dtype=np.float64).reshape(-1)
        if v.shape[0] != st.n:
            raise ValueError("prior %r returned %d log-pdf values for %d rows" % (prior, v.shape[0], st.n))
        lp_out.add_(torch.as_tensor(v, device=lp_out.device))


def eval_incoming_callable(st, spec, lp_cols=None):
    log_priors_into(st, spec, st.params, st.lp, lp_cols)
    if bool(torch.isinf(st.lp).any().item()):
        st.flags[0] |= _lib.FLAG_PRIOR_OOB
    # incoming particles are always evaluated (vector_mcmc.py:116-118 has no prior test)
    _model_ll(st, spec, st.params, torch.zeros_like(st.lp), st.ll)


def mh_step_callable(st, spec, path, k, rng_struct):
    if not hasattr(st, "_prop"):
        st._prop = torch.empty((st.d + 2, st.n), dtype=torch.float64, device=st.cols.device)
    new = st._prop
    new_params, new_ll, new_lp = new[:st.d], new[st.d], new[st.d + 1]
    s = stream_ptr()
    call("smcb_mh_propose", C.byref(spec.c), ptr(st.cols), st.n, st.n, st.n_global, ptr(st.chol),
         ptr(st.accept_counts), k, C.byref(rng_struct), ptr(new_params), ptr(new_lp), s)
    if spec.external or spec.n_cov:
        # the proposal kernel knows the scalar device priors only: matrix and host-evaluated priors are
        # (re-)evaluated on the proposed rows
        log_priors_into(st, spec, new_params, new_lp)
    _model_ll(st, spec, new_params, new_lp, new_ll)
    call("smcb_mh_accept", C.byref(path), ptr(st.cols), ptr(new_params), st.n, st.d, st.n, ptr(st.ll),
         ptr(new_ll), ptr(st.lp), ptr(new_lp), ptr(st.moved), ptr(st.accept_counts), k, C.byref(rng_struct), s)
