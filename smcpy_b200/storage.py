"""
Result storage: the in-memory container and the marginal-log-likelihood estimate of
smcpy/utils/storage.py:14-76 (estimate_marginal_log_likelihoods :18-21 is on the hot path's
output side).  File-backed storage (HDF5 / pickle) is out of scope:
feed the reference's classes with `Particles` host views if checkpoint files are needed.
Steps hold device snapshots; `retain` bounds how many are kept (all by default, as the
reference keeps every step).
"""

import threading

import numpy as np


class ContextManager:
    """Thread-local stack of active storage contexts (smcpy/utils/context_manager.py:5-24)."""
    _contexts = threading.local()

    def __enter__(self):
        type(self).get_contexts().append(self)
        return self

    def __exit__(self, typ, value, traceback):
        type(self).get_contexts().pop()

    @classmethod
    def get_contexts(cls):
        if not hasattr(cls._contexts, "stack"):
            cls._contexts.stack = []
        return cls._contexts.stack

    @classmethod
    def get_context(cls):
        try:
            return cls.get_contexts()[-1]
        except IndexError:
            raise TypeError("No context on context stack")


class InMemoryStorage(ContextManager):
    def __init__(self, retain=None):
        self.is_restart = False
        self._step_list = []
        self._phi_sequence = []
        self._mut_ratio_sequence = []
        self._totals = []
        self._retain = retain

    @property
    def phi_sequence(self):
        return self._phi_sequence

    @property
    def mut_ratio_sequence(self):
        return self._mut_ratio_sequence

    def __getitem__(self, idx):
        return self._step_list[idx]

    def __iter__(self):
        return iter(self._step_list)

    def __len__(self):
        return len(self._step_list)

    def save_step(self, step):
        self._step_list.append(step)
        self._phi_sequence.append(step.attrs["phi"])
        self._mut_ratio_sequence.append(step.attrs["mutation_ratio"])
        self._totals.append(step.total_unnorm_log_weight)
        if self._retain is not None and len(self._step_list) > self._retain:
            del self._step_list[0]

    def estimate_marginal_log_likelihoods(self):
        """cumsum([0] + per-step total_unnorm_log_weight)  (utils/storage.py:18-21)."""
        return np.cumsum([0] + list(self._totals))
