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
