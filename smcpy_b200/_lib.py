"""
ctypes binding of libsmcb200.so (C ABI in include/smcb200.h).

There is no CPU fallback: importing this module without the built library, or
using it without a CUDA device, raises.  Build with `python -m smcpy_b200.build`.
"""

import ctypes as C
import os

import torch

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMCB_LIB", os.path.join(PKG, "libsmcb200.so"))   # SMCB_LIB: tuning aid

MAX_PARAMS = 32
MAX_PRIORS = 32
MAX_COV_ARGS = 6
MAX_MCMC_STEPS = 4096
MVN_STATS = 12

MODEL_LINEAR, MODEL_SPRING_MASS, MODEL_IDENTITY = 1, 2, 3
LL_NORMAL, LL_MVNORMAL, LL_MULTISOURCE = 1, 2, 3
MAX_SEGMENTS = 8
PRIOR_UNIFORM, PRIOR_IMPROPER_UNIFORM, PRIOR_IMPROPER_COV, PRIOR_EXTERNAL = 1, 2, 3, 4
FLAG_PRIOR_OOB, FLAG_MODEL_NAN, FLAG_COV_NOT_PD, FLAG_CHOL_FAIL, FLAG_COV_RESHIFT, FLAG_XCHG_TIMEOUT = 1, 2, 4, 8, 16, 32
RESAMPLE_STANDARD, RESAMPLE_STRATIFIED, RESAMPLE_SYSTEMATIC = 0, 1, 2
SCAN_SEQUENTIAL, SCAN_LOOKBACK = 0, 1


class Prior(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ncols", C.c_int32), ("a", C.c_double), ("b", C.c_double),
                ("logp_in", C.c_double)]


PROPOSAL_NORMAL, PROPOSAL_UNIFORM = 1, 2


class ProposalCol(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("a", C.c_double), ("b", C.c_double)]


class Problem(C.Structure):
    _fields_ = [("model_id", C.c_int32), ("loglike_id", C.c_int32), ("n_params", C.c_int32),
                ("n_model_params", C.c_int32), ("n_data", C.c_int32), ("n_snapshots", C.c_int32),
                ("sigma_is_param", C.c_int32), ("n_priors", C.c_int32), ("sigma", C.c_double),
                ("model_args", C.c_double * 8), ("data", C.c_void_p), ("xgrid", C.c_void_p),
                ("cov_fixed", C.c_double * MAX_COV_ARGS), ("cov_param_col", C.c_int32 * MAX_COV_ARGS),
                ("priors", Prior * MAX_PRIORS), ("has_proposal", C.c_int32), ("pad", C.c_int32),
                ("proposal", ProposalCol * MAX_PARAMS), ("n_segments", C.c_int32),
                ("seg_len", C.c_int32 * MAX_SEGMENTS), ("seg_sigma_col", C.c_int32 * MAX_SEGMENTS),
                ("pad2", C.c_int32), ("seg_sigma", C.c_double * MAX_SEGMENTS), ("mvn_stats", C.c_void_p),
                ("iw_chol", (C.c_double * 6) * 2), ("lin_centered", C.c_void_p)]


class Path(C.Structure):
    _fields_ = [("phi", C.c_double), ("phi_prev", C.c_double), ("lam", C.c_double)]


MAX_RANKS = 8


class Xchg(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("gen", C.c_uint64),
                ("peer_slots", C.c_void_p * MAX_RANKS), ("global_counts", C.c_void_p), ("ticket", C.c_void_p)]


class Resample(C.Structure):
    _fields_ = [("kind", C.c_int32), ("world", C.c_int32), ("rank", C.c_int32), ("n_cols", C.c_int32),
                ("n_local", C.c_int64), ("n_global", C.c_int64), ("seed", C.c_uint64), ("stream_id", C.c_uint64),
                ("cdf", C.c_void_p), ("totals", C.c_void_p), ("t_peer", C.c_void_p * MAX_RANKS),
                ("src", C.c_void_p), ("src_stride", C.c_int64), ("dst_peer", C.c_void_p * MAX_RANKS),
                ("dst_stride", C.c_int64), ("idx_out", C.c_void_p), ("n_unique_out", C.c_void_p)]


class Rng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("seed", C.c_uint64), ("stream", C.c_uint64),
                ("id_offset", C.c_int64), ("z", C.c_void_p), ("u", C.c_void_p)]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_U64 = C.c_uint64
_D = C.c_double
_PP = C.POINTER

# name -> argtypes; every function returns int except where noted below
SIGNATURES = {
    "smcb_aos_to_soa": [_P, _P, _I64, _I32, _I64, _P],
    "smcb_soa_to_aos": [_P, _P, _I64, _I32, _I64, _P],
    "smcb_fp64_probe": [_P, _I32, _I32, _P],
    "smcb_mvn_scatter": [_P, _I32, _I32, _P, _P],
    "smcb_linear_center": [_P, _P, _I32, _P, _P],
    "smcb_dmma_probe": [_P, _I32, _I32, _P],
    "smcb_fill_f64": [_P, _I64, _D, _P],
    "smcb_reweight": [_I64, _P, _P, _P, _P, _I64, _PP(Path), _P, _P, _P, _P, _P],
    "smcb_path_eval": [_I64, _P, _P, _P, _PP(Path), _I32, _P, _P],
    "smcb_reweight_partials": [_I64, _P, _P, _P, _P, _I64, _PP(Path), _P, _P, _P],
    "smcb_lse_merge": [_P, _I32, _P, _P],
    "smcb_reweight_finalize": [_I64, _P, _P, _P, _P, _I64, _PP(Path), _P, _P, _P, _P],
    "smcb_ess_bisect": [_I64, _P, _P, _P, _D, _D, _D, _D, _I64, _P, _P, _P],
    "smcb_ess_bisect_begin": [_D, _P, _P],
    "smcb_ess_partial": [_I64, _P, _P, _P, _D, _D, _D, _P, _P, _P, _P],
    "smcb_ess_bisect_update": [_P, _I32, _D, _D, _I64, _P, _P],
    "smcb_ess_bisect_max_iters": [],
    "smcb_ess_bisect_xchg_slot_words": [],
    "smcb_resample_uniforms": [_I32, _I64, _I64, _I64, _U64, _U64, _P, _P],
    "smcb_resample_exponentials": [_I64, _I64, _I64, _U64, _U64, _P, _P],
    "smcb_resample_spacings_finish": [_P, _I64, _P, _P, _I64, _U64, _U64, _P, _P],
    "smcb_cdf_scan": [_P, _P, _I64, _I32, _P, _P],
    "smcb_searchsorted": [_P, _I64, _P, _I64, _P, _P],
    "smcb_searchsorted_offset": [_P, _I64, _P, _P, _I64, _P, _P],
    "smcb_gather": [_P, _I64, _P, _I64, _P, _I64, _I32, _P],
    "smcb_gather_p2p": [_P, _I64, _P, _I64, _I32, _I32, _P, _P, C.POINTER(C.c_void_p), _I64, _P],
    "smcb_count_unique": [_P, _I64, _I64, _P, _P, _P],
    "smcb_weighted_cov": [_P, _I64, _P, _I64, _I32, _P, _P, _P, _P, _P],
    "smcb_weighted_cov_xchg": [_P, _I64, _P, _I64, _I32, _P, _P, _P, _P, _PP(Xchg), _P, _P, _P],
    "smcb_weighted_sums1": [_P, _I64, _P, _I64, _I32, _P, _P, _P],
    "smcb_mean_from_sums": [_P, _I32, _P, _P],
    "smcb_weighted_sums2": [_P, _I64, _P, _I64, _I32, _P, _P, _P, _P],
    "smcb_cov_from_sums": [_P, _P, _I32, _P, _P],
    "smcb_weighted_moments_shifted": [_P, _I64, _P, _I64, _I32, _P, _P, _P, _P],
    "smcb_cov_from_shifted_sums": [_P, _P, _I32, _P, _P, _P, _P],
    "smcb_cholesky": [_P, _I32, _P, _P, _P],
    "smcb_mh_init": [_PP(Problem), _P, _I64, _I64, _P, _P, _P, _P, _P, _P],
    "smcb_log_priors": [_PP(Problem), _P, _I64, _I64, _P, _P, _P],
    "smcb_mh_step": [_PP(Problem), _PP(Path), _P, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _I32,
                     _PP(Rng), _P, _P],
    "smcb_mh_step_xchg": [_PP(Problem), _PP(Path), _P, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _I32,
                          _PP(Rng), _P, _PP(Xchg), _P],
    "smcb_mh_steps": [_PP(Problem), _PP(Path), _P, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _I32, _I32,
                      _PP(Rng), _P, _PP(Xchg), _P],
    "smcb_count_moved": [_P, _I64, _P, _P, _P],
    "smcb_sample_priors": [_PP(Problem), _P, _I64, _I64, _U64, _U64, _I64, _P],
    "smcb_mh_propose": [_PP(Problem), _P, _I64, _I64, _I64, _P, _P, _I32, _PP(Rng), _P, _P, _P],
    "smcb_normal_loglike_rows": [_P, _I64, _I32, _P, _P, _D, _P, _P, _P, _P],
    "smcb_mh_accept": [_PP(Path), _P, _P, _I64, _I32, _I64, _P, _P, _P, _P, _P, _P, _I32, _PP(Rng), _P],
    "smcb_ess_bisect_xchg": [_I64, _P, _P, _P, _D, _D, _D, _D, _I64, _P, _P, _PP(Xchg), _P],
    "smcb_exp_probe": [_P, _P, _I64, _P],
    "smcb_reweight_fused": [_I64, _P, _P, _D, _P, _P, _I64, _PP(Path), _D, _P, _P, _P, _P, _PP(Xchg), _P],
    "smcb_ess_search": [_I64, _P, _P, _P, _D, _D, _D, _D, _I64, _D, _P, _P, _PP(Xchg), _P],
    "smcb_xchg_slot_words": [],
    "smcb_mh_xchg_slot_words": [],
    "smcb_test_search_host": [_P, _I64, _D, _D, _D, _I32, _P, _P, _I32],
    "smcb_resample_exponential_sums": [_I64, _I64, _I64, _U64, _U64, _P, _P, _P, _P],
    "smcb_cdf_scan_weights": [_I64, _P, _P, _D, _P, _P, _I64, _PP(Path), _P, _P, _P, _P, _P],
    "smcb_xchg_big_slot_words": [],
    "smcb_xchg_max_vals": [],
    "smcb_xchg_allgather": [_PP(Xchg), _P, _I32, _P, _I32, _P, _P],
    "smcb_resample_fused": [_PP(Resample), _P, _P],
    "smcb_model_register_nvrtc": [C.c_char_p, _I32, C.POINTER(C.c_int32)],
    "smcb_user_loglike": [_I32, _P, _I64, _I64, _I32, _P, _P, _P, _D, _P, _P, _P, _P],
}
OTHER_SYMBOLS = ["smcb_version", "smcb_last_error", "smcb_workspace_bytes", "smcb_set_option", "smcb_get_option",
                 "smcb_last_mh_variant", "smcb_test_search_margin"]


class SmcbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libsmcb200.so is not built (%s). Run `python -m smcpy_b200.build`; smcpy_b200 has no "
            "CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError("libsmcb200.so is stale (missing %s): rebuild with "
                              "`python -m smcpy_b200.build`" % name) from e
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.smcb_version.restype = C.c_int
    lib.smcb_last_error.restype = C.c_char_p
    lib.smcb_workspace_bytes.argtypes = [_I64, _I32]
    lib.smcb_workspace_bytes.restype = C.c_size_t
    lib.smcb_set_option.argtypes = [C.c_char_p, _I64]
    lib.smcb_set_option.restype = C.c_int
    lib.smcb_get_option.argtypes = [C.c_char_p]
    lib.smcb_get_option.restype = _I64
    lib.smcb_last_mh_variant.restype = C.c_char_p
    lib.smcb_test_search_margin.argtypes = [_P, _I64, _D, _D, _D]
    lib.smcb_test_search_margin.restype = _D
    return lib


class _LazyLib:
    """Loads libsmcb200.so on first use, so that importing the package (e.g. to run
    `python -m smcpy_b200.build`) does not need a built library."""
    _handle = None

    def __getattr__(self, name):
        if _LazyLib._handle is None:
            _LazyLib._handle = _load()
        return getattr(_LazyLib._handle, name)


lib = _LazyLib()
# kernels launched per C-ABI call (bench.py reports the sum as gpu_launches)
LAUNCHES = {"smcb_reweight": 2, "smcb_cdf_scan": 2, "smcb_cdf_scan_weights": 2, "smcb_resample_exponential_sums": 2,
            "smcb_resample_fused": 3, "smcb_xchg_slot_words": 0, "smcb_mh_xchg_slot_words": 0, "smcb_xchg_big_slot_words": 0, "smcb_xchg_max_vals": 0, "smcb_count_unique": 2, "smcb_weighted_cov": 7, "smcb_weighted_cov_xchg": 9,
            "smcb_weighted_sums1": 2, "smcb_weighted_moments_shifted": 3, "smcb_weighted_sums2": 3, "smcb_ess_bisect_max_iters": 0,
            "smcb_ess_bisect_xchg_slot_words": 0}
launch_count = 0


def call(name, *args, launches=None):
    """Invoke a C-ABI entry point; raises SmcbError with the library's message."""
    global launch_count
    launch_count += LAUNCHES.get(name, 1) if launches is None else launches
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.smcb_last_error().decode()
        if rc == 3:
            raise NotImplementedError(msg)
        raise SmcbError("%s failed (%d): %s" % (name, rc, msg))


def set_option(name, value):
    """Runtime option of the library (smcb_set_option); returns the previous value."""
    old = int(lib.smcb_get_option(name.encode()))
    if lib.smcb_set_option(name.encode(), int(value)) != 0:
        raise SmcbError(lib.smcb_last_error().decode())
    return old


class option:
    """with option("no_tma", 1): ...  -- scoped runtime option."""

    def __init__(self, name, value):
        self.name, self.value = name, value

    def __enter__(self):
        self.old = set_option(self.name, self.value)

    def __exit__(self, *exc):
        set_option(self.name, self.old)


def last_mh_variant():
    return lib.smcb_last_mh_variant().decode()


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("smcpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device, as the void* the C ABI takes.  Called once per
    library call: torch.cuda.current_stream() costs ~14 us of Python, the raw getter ~1 us."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
