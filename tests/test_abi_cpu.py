"""CPU-side checks: the C-ABI library loads and exports every symbol include/smcb200.h
declares; host logic that needs no GPU (shard bounds, path bookkeeping, validation)."""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "smcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smcb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import smcpy_b200._lib as L
    lib = ctypes.CDLL(L.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    bound = set(L.SIGNATURES) | set(L.OTHER_SYMBOLS)
    assert set(names) == bound, set(names) ^ bound
    assert L.lib.smcb_version() == 100
    assert L.lib.smcb_workspace_bytes(1 << 20, 2) > (1 << 20)


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of every struct of include/smcb200.h as gcc lays them out == the ctypes mirrors."""
    import subprocess
    import smcpy_b200._lib as L
    structs = {"smcb_prior_t": (L.Prior, ["kind", "ncols", "a", "b", "logp_in"]),
               "smcb_proposal_col_t": (L.ProposalCol, ["kind", "a", "b"]),
               "smcb_path_t": (L.Path, ["phi", "phi_prev"]),
               "smcb_rng_t": (L.Rng, ["mode", "seed", "stream", "id_offset", "z", "u"]),
               "smcb_xchg_t": (L.Xchg, ["world", "rank", "gen", "peer_slots", "global_counts", "ticket"]),
               "smcb_resample_t": (L.Resample, ["kind", "n_local", "seed", "cdf", "totals", "t_peer", "src", "src_stride",
                                                "dst_peer", "dst_stride", "idx_out", "n_unique_out"]),
               "smcb_problem_t": (L.Problem, ["model_id", "n_priors", "sigma", "model_args", "data", "xgrid", "cov_fixed",
                                              "cov_param_col", "priors", "has_proposal", "proposal", "n_segments",
                                              "seg_len", "seg_sigma_col", "seg_sigma", "mvn_stats", "iw_chol", "lin_centered"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "smcb200.h"', "int main(void) {"]
    for cname, (_, fields) in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f in fields:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
    lines.append("return 0; }")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, (cls, fields) in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for f in fields:
            assert int(got["%s.%s" % (cname, f)]) == getattr(cls, f).offset, (cname, f)
    assert L.MVN_STATS == 12 and L.MAX_RANKS == 8


def test_argument_errors_are_reported_without_a_gpu():
    import smcpy_b200._lib as L
    with pytest.raises(L.SmcbError, match="bad argument"):
        L.call("smcb_fill_f64", None, 0, 1.0, None)
    p = L.Problem()
    with pytest.raises(L.SmcbError):
        L.call("smcb_mh_init", ctypes.byref(p), None, 0, 0, None, None, None, None, None, None)


def test_no_cpu_fallback():
    import torch
    import smcpy_b200 as sb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sb.Particles({"a": [1.0, 2.0]}, [0.0, 0.0], [0.0, 0.0])


def test_shard_bounds_partition():
    from smcpy_b200.engine import shard_bounds
    for n, w in ((10, 3), (1 << 20, 8), (7, 8), (64, 1)):
        b = [shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert sizes == [len(x) for x in np.array_split(np.arange(n), w)]


def test_path_bookkeeping():
    # smcpy/paths.py:9-79 semantics (tests/unit/test_paths.py in the reference)
    from smcpy_b200 import GeometricPath
    p = GeometricPath(required_phi=[0.3, 0.1, 1.0, 2.0])
    assert p.required_phi_list == [0.1, 0.3] and p.lam == 0.1
    p.phi = 0.2
    assert p.previous_phi == 0 and p.delta_phi == 0.2
    with pytest.raises(ValueError):
        p.phi = 0.2
    p.undo_phi_set()
    assert p.phi == 0 and p.previous_phi is None
    lst = p.required_phi_list
    lst.append(5)
    assert p.required_phi_list == [0.1, 0.3]


def test_validation_errors_match_reference():
    import smcpy_b200 as sb
    from smcpy_b200.updater import Updater
    from smcpy_b200.mutator import Mutator
    from smcpy_b200.initializer import Initializer
    with pytest.raises(ValueError):
        Updater(1.1, None, sb.standard)
    with pytest.raises(ValueError):
        Updater(0.5, None, sb.standard, particles_warn_threshold=-1)
    with pytest.raises(TypeError):
        Updater(0.5, None, resample_rng="standard")
    with pytest.raises(TypeError):
        Mutator(object())
    with pytest.raises(TypeError):
        Initializer(object())
    mc = sb.B200VectorMCMC(sb.LinearModel(np.arange(3.0)), np.zeros(3), [], 1.0)
    with pytest.raises(TypeError):
        mc.rng = 5
    k = sb.VectorMCMCKernel(mc, ["a", "b"], rng=np.random.default_rng(0))
    assert mc.rng is k.rng and k.param_order == ("a", "b")
    assert mc.evaluate_log_posterior.__self__ is k.path
    with pytest.raises(ValueError):
        sb.AdaptiveSampler(k, show_progress_bar=False).sample(10, 1, target_ess=1.0)


def test_prior_descriptions():
    from scipy import stats
    import smcpy_b200 as sb
    from smcpy_b200.priors import describe_prior
    k, nc, a, b, lp = describe_prior(stats.uniform(-6, 12))
    assert (k, nc, a, b) == (1, 1, -6.0, 12.0) and lp == -np.log(12.0)
    assert describe_prior(stats.uniform(loc=1.0, scale=2.0))[2:4] == (1.0, 2.0)
    assert describe_prior(sb.ImproperUniform(None, 3))[2:4] == (-np.inf, 3.0)
    assert describe_prior(sb.ImproperCov(3))[:2] == (3, 3)
    assert describe_prior(stats.norm(0, 1))[:2] == (4, 1)        # any object with logpdf: host-evaluated column
    with pytest.raises(TypeError):
        describe_prior(3.0)


def test_every_documented_runtime_option_exists():
    """The option names the header documents for smcb_set_option / smcb_get_option are the ones the library knows
    (an unknown name answers INT64_MIN), and a scoped change is undone."""
    import smcpy_b200._lib as L
    text = open(os.path.join(ROOT, "include", "smcb200.h")).read()
    doc = text[text.index("/* Runtime options"):text.index("int smcb_set_option")]
    names = re.findall(r'"([a-z0-9_]+)"', doc)
    assert len(names) >= 17 and "no_pdl" in names and "spring_form" in names, names
    for n in names:
        assert L.lib.smcb_get_option(n.encode()) != -(1 << 63), n
    assert L.lib.smcb_get_option(b"no_such_option") == -(1 << 63)
    before = L.lib.smcb_get_option(b"spring_form")
    with L.option("spring_form", 0):
        assert L.lib.smcb_get_option(b"spring_form") == 0
    assert L.lib.smcb_get_option(b"spring_form") == before == 1
