"""
File-backed result storage (smcpy_b200/storage.py; reference: smcpy/utils/storage.py:173-277 and
tests/unit/test_storage.py): host-side bookkeeping with stand-in steps, no GPU.
"""

import pickle
import types

import numpy as np
import pytest

from smcpy_b200.storage import ContextManager, HDF5Storage, InMemoryStorage, PickleStorage


def _fake_step(phi, total, n=5, seed=0):
    rng = np.random.default_rng(seed)
    return types.SimpleNamespace(
        param_dict={"a": rng.normal(size=n), "b": rng.normal(size=n)}, log_likes=rng.normal(size=(n, 1)),
        log_weights=np.full((n, 1), np.log(1 / n)), attrs={"phi": phi, "mutation_ratio": 0.5 + phi / 4},
        total_unnorm_log_weight=total)


@pytest.mark.parametrize("mode", ["r", "x", "ab"])
def test_pickle_storage_rejects_unknown_modes(tmp_path, mode):
    with pytest.raises(ValueError, match="Available modes"):
        PickleStorage(tmp_path / "s.pkl", mode=mode)


def test_pickle_storage_append_restart_overwrite(tmp_path):
    fn = tmp_path / "steps.pkl"
    st = PickleStorage(fn)
    assert not st.is_restart and len(st) == 0
    for i, (phi, tot) in enumerate([(0.0, 0.0), (0.3, -4.0), (1.0, -2.5)]):
        st.save_step(_fake_step(phi, tot, seed=i))
    assert len(st) == 3 and st.phi_sequence == [0.0, 0.3, 1.0]
    np.testing.assert_allclose(st.mut_ratio_sequence, [0.5, 0.575, 0.75])
    np.testing.assert_allclose(st.estimate_marginal_log_likelihoods(), [0.0, 0.0, -4.0, -6.5])
    # the file is a plain stream of pickled dicts
    with open(fn, "rb") as f:
        first = pickle.load(f)
        second = pickle.load(f)
    assert list(first["params"]) == ["a", "b"] and second["attrs"]["phi"] == 0.3
    np.testing.assert_array_equal(second["params"]["a"], _fake_step(0.3, -4.0, seed=1).param_dict["a"])
    # reopen: restart, same bookkeeping; append continues
    again = PickleStorage(fn)
    assert again.is_restart and len(again) == 3 and again.phi_sequence == [0.0, 0.3, 1.0]
    again.save_step(_fake_step(1.0, -1.0, seed=9))
    assert len(PickleStorage(fn)) == 4
    with pytest.raises(IndexError):
        again._index(4)
    assert again._index(-1) == 3
    # mode 'w' starts over
    fresh = PickleStorage(fn, mode="w")
    assert not fresh.is_restart and len(fresh) == 0


def test_storage_context_stack():
    with pytest.raises(TypeError):
        ContextManager.get_context()
    a, b = InMemoryStorage(), InMemoryStorage()
    with a:
        assert ContextManager.get_context() is a
        with b:
            assert ContextManager.get_context() is b
        assert ContextManager.get_context() is a
    with pytest.raises(TypeError):
        ContextManager.get_context()


def test_hdf5_storage_needs_h5py(tmp_path):
    try:
        import h5py
        if hasattr(h5py, "File"):
            pytest.skip("h5py present")
    except ImportError:
        pass
    with pytest.raises(ImportError, match="needs h5py"):
        HDF5Storage(tmp_path / "s.h5")
    with pytest.raises(ValueError, match="Available modes"):
        HDF5Storage(tmp_path / "s.h5", mode="r")


# ---- HDF5 layout through an h5py stand-in (tests/h5py_stub.py): written here, read by the reference ---------
@pytest.fixture
def h5stub(monkeypatch):
    import sys
    from tests import h5py_stub
    mod = h5py_stub.module()
    monkeypatch.setitem(sys.modules, "h5py", mod)
    return mod


def test_hdf5_storage_writes_the_reference_layout(tmp_path, h5stub):
    """utils/storage.py:108-122: one group per step named by its index; attrs phi, total_unnorm_log_weight,
    mutation_ratio; datasets log_likes, log_weights; group params with one dataset per parameter in order."""
    fn = tmp_path / "run.h5"
    st = HDF5Storage(fn, mode="w")
    steps = [_fake_step(phi, tot, seed=i) for i, (phi, tot) in enumerate([(0.0, 0.0), (0.4, -3.0), (1.0, -1.5)])]
    for s in steps:
        st.save_step(s)
    assert len(st) == 3 and st.phi_sequence == [0.0, 0.4, 1.0]
    np.testing.assert_allclose(st.mut_ratio_sequence, [0.5, 0.6, 0.75])
    np.testing.assert_allclose(st.estimate_marginal_log_likelihoods(), [0.0, 0.0, -3.0, -4.5])
    with h5stub.File(fn, "r") as h5:
        assert list(h5.keys()) == ["0", "1", "2"]
        g = h5["1"]
        assert list(g.attrs) == ["phi", "total_unnorm_log_weight", "mutation_ratio"]
        assert sorted(g.keys()) == ["log_likes", "log_weights", "params"]
        assert list(g["params"].keys()) == ["a", "b"]
        np.testing.assert_array_equal(g["params"]["b"][:], steps[1].param_dict["b"])
        np.testing.assert_array_equal(g["log_likes"][:], steps[1].log_likes)
        assert g["log_weights"][:].shape == (5, 1)
    # restart: reopening in append mode finds the stored steps and continues the numbering
    again = HDF5Storage(fn)
    assert again.is_restart and len(again) == 3 and again.phi_sequence == [0.0, 0.4, 1.0]
    again.save_step(_fake_step(1.0, -0.5, seed=7))
    with h5stub.File(fn, "r") as h5:
        assert list(h5.keys()) == ["0", "1", "2", "3"]
    assert not HDF5Storage(tmp_path / "other.h5").is_restart
    # mode 'w' truncates on the first write
    fresh = HDF5Storage(fn, mode="w")
    assert len(fresh) == 0 and not fresh.is_restart
    fresh.save_step(steps[0])
    with h5stub.File(fn, "r") as h5:
        assert list(h5.keys()) == ["0"]


def test_hdf5_files_are_interchangeable_with_the_reference(tmp_path, h5stub):
    """A file written by smcpy_b200.HDF5Storage is read by the UNMODIFIED reference's HDF5Storage (installed
    under baseline/_ref; skipped when absent) and vice versa, through the same h5py stand-in."""
    import os
    import sys
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "smcpy")):
        pytest.skip("reference not installed under baseline/_ref")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import smcpy.utils.storage as rs
    old = rs.h5py
    rs.h5py = h5stub
    try:
        fn = tmp_path / "ours.h5"
        ours = HDF5Storage(fn, mode="w")
        steps = [_fake_step(phi, tot, seed=10 + i) for i, (phi, tot) in enumerate([(0.0, 0.0), (0.5, -2.0), (1.0, -1.0)])]
        for s in steps:
            ours.save_step(s)
        theirs = rs.HDF5Storage(fn, mode="a")
        assert theirs.is_restart and len(theirs) == 3
        assert theirs.phi_sequence == [0.0, 0.5, 1.0]
        np.testing.assert_allclose(theirs.estimate_marginal_log_likelihoods(), [0.0, 0.0, -2.0, -3.0])
        p = theirs[1]                                           # the reference's own Particles (numpy)
        assert p.param_names == ("a", "b") and p.attrs["phi"] == 0.5
        np.testing.assert_array_equal(p.params, np.column_stack([steps[1].param_dict["a"], steps[1].param_dict["b"]]))
        np.testing.assert_array_equal(p.log_likes, steps[1].log_likes)
        # the other way round: the reference writes, our store reads the bookkeeping and the raw groups
        fn2 = tmp_path / "theirs.h5"
        w = rs.HDF5Storage(fn2, mode="w")
        for s in steps:
            s.params = np.column_stack(list(s.param_dict.values()))
            w.save_step(s)
        mine = HDF5Storage(fn2)
        assert mine.is_restart and len(mine) == 3 and mine.phi_sequence == [0.0, 0.5, 1.0]
        np.testing.assert_allclose(mine.estimate_marginal_log_likelihoods(), [0.0, 0.0, -2.0, -3.0])
        with h5stub.File(fn2, "r") as h5:
            np.testing.assert_array_equal(h5["2"]["params"]["a"][:], steps[2].param_dict["a"])
    finally:
        rs.h5py = old
