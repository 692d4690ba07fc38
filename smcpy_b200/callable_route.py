"""
Un-fused Metropolis route for user models (BASELINE.json north_star: "User models plug in
either as a torch callable on device tensors or as a registered __device__ model").
Torch callable, per step: smcb_mh_propose -> user model -> smcb_normal_loglike_rows ->
smcb_mh_accept; pays 8*M bytes per particle for the model outputs, which the fused built-ins
avoid.  Registered CUDA source (models.NvrtcModel): smcb_mh_propose -> smcb_user_loglike ->
smcb_mh_accept; outputs stay in registers.
"""

import ctypes as C

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .models import NvrtcModel


def _model_ll(st, spec, soa_params, lp, ll_out):
    """ll_out[i] = Normal log-like of model(params_i); rows with lp = -inf keep 0."""
    c = spec.c
    n, d = st.n, st.d
    dm = int(c.n_model_params)
    sigma_col = soa_params[d - 1] if c.sigma_is_param else None
    if isinstance(spec.model, NvrtcModel):                     # registered __device__ model: one kernel
        m = spec.model
        call("smcb_user_loglike", m.handle, ptr(soa_params), soa_params.stride(0), n, int(c.n_data),
             C.c_void_p(int(c.data)), ptr(m.x_dev()), ptr(sigma_col), float(c.sigma), ptr(lp), ptr(ll_out),
             ptr(st.flags), stream_ptr())
        return
    aos = soa_params[:dm].t().contiguous()                     # (n, d_model) for the user callable
    out = spec.model(aos)
    if not (torch.is_tensor(out) and out.is_cuda and out.dtype == torch.float64):
        raise TypeError("torch-callable models must return a float64 CUDA tensor (N, M)")
    out = out.contiguous()
    if out.shape != (n, int(c.n_data)):
        raise ValueError("model output shape %s != (%d, %d)" % (tuple(out.shape), n, int(c.n_data)))
    call("smcb_normal_loglike_rows", ptr(out), n, int(c.n_data), C.c_void_p(int(c.data)), ptr(sigma_col),
         float(c.sigma), ptr(lp), ptr(ll_out), ptr(st.flags), stream_ptr())


def eval_incoming_callable(st, spec, lp_cols=None):
    call("smcb_log_priors", C.byref(spec.c), ptr(st.cols), st.n, st.n, ptr(st.lp), ptr(lp_cols), stream_ptr())
    if bool(torch.isinf(st.lp).any().item()):
        st.flags[0] = _lib.FLAG_PRIOR_OOB
    _model_ll(st, spec, st.params, st.lp, st.ll)


def mh_step_callable(st, spec, path, k, rng_struct):
    if not hasattr(st, "_prop"):
        st._prop = torch.empty((st.d + 2, st.n), dtype=torch.float64, device=st.cols.device)
    new = st._prop
    new_params, new_ll, new_lp = new[:st.d], new[st.d], new[st.d + 1]
    s = stream_ptr()
    call("smcb_mh_propose", C.byref(spec.c), ptr(st.cols), st.n, st.n, st.n_global, ptr(st.chol),
         ptr(st.accept_counts), k, C.byref(rng_struct), ptr(new_params), ptr(new_lp), s)
    _model_ll(st, spec, new_params, new_lp, new_ll)
    call("smcb_mh_accept", C.byref(path), ptr(st.cols), ptr(new_params), st.n, st.d, st.n, ptr(st.ll),
         ptr(new_ll), ptr(st.lp), ptr(new_lp), ptr(st.moved), ptr(st.accept_counts), k, C.byref(rng_struct), s)
