"""
Result storage with the interface of smcpy/utils/storage.py: the in-memory container (:43-76), the
marginal-log-likelihood estimate (:18-21, the output side of the hot path) and the two file-backed
stores (HDF5Storage :79-170, PickleStorage :173-277).

Steps are device snapshots (`Particles`).  InMemoryStorage keeps them in HBM (`retain` bounds how
many; the reference keeps every step, 17 GB per step at 64M x 32).  The file-backed stores stream
every step out ASYNCHRONOUSLY (StepStreamer): save_step() enqueues a device-to-host copy of the
snapshot's SoA rows into page-locked memory on a dedicated copy stream, ordered after the producing
kernels by an event, and a writer thread appends the record to the file once the copy has landed --
the sampler goes on with the next SMC step meanwhile.  Nothing stays on the device; reading a step
back (`storage[i]`, iteration, restart) drains the queue first and uploads the step again.

On-disk formats:
  * HDF5Storage writes the reference's layout -- one group per step named by its index, attrs `phi`,
    `total_unnorm_log_weight`, `mutation_ratio`, datasets `log_likes`, `log_weights`, group `params`
    with one dataset per parameter in order -- so files are interchangeable with smcpy's
    (needs h5py, which this image does not ship: the class raises ImportError without it).
  * PickleStorage appends one pickle per step of a plain dict {params, log_likes, log_weights,
    attrs} (numpy arrays).  The reference pickles its own Particles class, which only smcpy can
    load; the plain dict loads anywhere.

Restart (samplers `_initialize`, samplers.py:87-92): a store opened in mode 'a' on an existing file
sets `is_restart`; the sampler then resumes from the last stored step and phi sequence.
"""

import os
import pickle
import threading

import numpy as np

_local = threading.local()


def _stack():
    if not hasattr(_local, "stack"):
        _local.stack = []
    return _local.stack


class ContextManager:
    """`with storage:` makes `storage` the result container of samplers built inside the block
    (smcpy/utils/context_manager.py); one stack per thread."""

    def __enter__(self):
        _stack().append(self)
        return self

    def __exit__(self, exc_type, exc, tb):
        _stack().pop()

    @classmethod
    def get_contexts(cls):
        return _stack()

    @classmethod
    def get_context(cls):
        s = _stack()
        if not s:
            raise TypeError("No context on context stack")
        return s[-1]


class StepStreamer:
    """Asynchronous device -> host -> file pipeline of the file-backed stores.
    submit(step, sink) returns at once; sink(record) runs on the writer thread, in submission order.
    The structure-of-arrays layout makes every parameter one contiguous row, so the record's arrays are
    views of the pinned buffer -- no transposition on the way out."""

    def __init__(self, depth=4):
        import queue
        self._q = queue.Queue(maxsize=depth)         # bounds the page-locked memory in flight
        self._err = None
        self._thread = None
        self._stream = None

    def _worker(self):
        while True:
            item = self._q.get()
            try:
                if item is None:
                    return
                done, rec, sink, _keep = item
                if self._err is None:
                    if done is not None:
                        done.synchronize()                # the copy of this step has landed
                    sink(rec)
            except BaseException as e:                     # surfaced by the next submit / drain
                self._err = e
            finally:
                self._q.task_done()

    def _start(self):
        if self._thread is None:
            self._thread = threading.Thread(target=self._worker, name="smcb-step-writer", daemon=True)
            self._thread.start()

    def submit(self, step, sink):
        self._raise()
        cols = getattr(step, "_cols", None)
        if cols is None or not getattr(cols, "is_cuda", False):
            self.drain()
            sink(_host_record(step))                            # host-side step object: nothing to overlap
            return
        self._start()
        import torch
        d, n = step._d, step._num_particles
        if self._stream is None:
            self._stream = torch.cuda.Stream()
        ready = torch.cuda.Event()
        ready.record()                                      # after everything that produced the snapshot
        buf = torch.empty((d + 2, n), dtype=torch.float64, pin_memory=True)
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)
            buf[:d + 1].copy_(cols[:d + 1], non_blocking=True)     # parameters + log-likelihood rows
            buf[d + 1].copy_(step._log_w, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        arr = buf.numpy()
        rec = {"params": {k: arr[i] for i, k in enumerate(step.param_names)},
               "log_likes": arr[d].reshape(-1, 1), "log_weights": arr[d + 1].reshape(-1, 1),
               "attrs": _step_attrs(step)}
        self._q.put((done, rec, sink, step))                # the snapshot stays referenced until its copy has landed

    def drain(self):
        if self._thread is not None:
            self._q.join()
        self._raise()

    def _raise(self):
        if self._err is not None:
            e, self._err = self._err, None
            raise e


def _step_attrs(step):
    a = {"phi": step.attrs["phi"], "mutation_ratio": step.attrs["mutation_ratio"],
         "total_unnorm_log_weight": step.total_unnorm_log_weight}
    if "rng_stream" in step.attrs:          # native generator: stream counter, restored on restart
        a["rng_stream"] = int(step.attrs["rng_stream"])
    return a


class BaseStorage(ContextManager):
    # False: the store never hands out a step other than the most recent one, so the samplers may give it zero-copy
    # views of the live working set instead of a device-to-device snapshot per SMC step
    wants_snapshots = True

    def __init__(self):
        self.is_restart = False

    def estimate_marginal_log_likelihoods(self):
        """cumsum([0] + per-step total_unnorm_log_weight)  (utils/storage.py:18-21)."""
        return np.cumsum([0] + [float(t) for t in self._step_totals()])

    def _step_totals(self):
        return [p.total_unnorm_log_weight for p in self]

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def _index(self, idx):
        n = len(self)
        if idx < 0:
            idx += n
        if idx < 0 or idx >= n:
            raise IndexError(f"index {idx} out of range.")
        return idx


class InMemoryStorage(BaseStorage):
    def __init__(self, retain=None):
        super().__init__()
        self._step_list = []
        self._phi_sequence = []
        self._mut_ratio_sequence = []
        self._totals = []
        self._retain = retain
        # retain=1: only the latest step is ever visible; at C5 scale a snapshot is 2.3 GB read + written per step
        self.wants_snapshots = retain != 1

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

    def _step_totals(self):
        return self._totals

    def save_step(self, step):
        self._step_list.append(step)
        self._phi_sequence.append(step.attrs["phi"])
        self._mut_ratio_sequence.append(step.attrs["mutation_ratio"])
        self._totals.append(step.total_unnorm_log_weight)
        if self._retain is not None and len(self._step_list) > self._retain:
            del self._step_list[0]


def _check_mode(mode):
    if mode != "a" and mode != "w":
        raise ValueError("Available modes are 'a' (default, read/append) and 'w' (overwrite).")


def _rank_filename(filename):
    """One file per rank when particles are sharded over several GPUs (every rank stores its own
    shard; the reference, where every MPI rank holds all particles, writes on rank 0 only)."""
    from . import engine
    rank, world = engine.dist_info()
    return str(filename) if world == 1 else "%s.rank%d" % (filename, rank)


def _host_record(step):
    """Host copy of one step (synchronous): what both file formats store."""
    return {"params": {k: np.array(v) for k, v in step.param_dict.items()},
            "log_likes": np.array(step.log_likes), "log_weights": np.array(step.log_weights),
            "attrs": _step_attrs(step)}


def _particles_from_record(rec):
    from .particles import Particles
    p = Particles(rec["params"], rec["log_likes"], rec["log_weights"])
    p.attrs = dict(rec["attrs"])
    return p


class PickleStorage(BaseStorage):
    """Append-only file of one pickle per step (utils/storage.py:173-277)."""

    def __init__(self, filename, mode="a"):
        _check_mode(mode)
        super().__init__()
        self._filename = _rank_filename(filename)
        self._offsets = []                   # byte offset of every stored step (appended by the writer thread)
        self._meta = []                      # (phi, mutation_ratio, total_unnorm_log_weight) per step
        self._streamer = StepStreamer()
        if mode == "w" and os.path.exists(self._filename):
            os.remove(self._filename)
        if os.path.exists(self._filename):
            self._scan()
            self.is_restart = True

    def _scan(self):
        self._offsets, self._meta = [], []
        with open(self._filename, "rb") as f:
            while True:
                pos = f.tell()
                try:
                    rec = pickle.load(f)
                except EOFError:
                    break
                self._offsets.append(pos)
                a = rec["attrs"]
                self._meta.append((a["phi"], a["mutation_ratio"], a["total_unnorm_log_weight"]))

    @property
    def phi_sequence(self):
        return [m[0] for m in self._meta]

    @property
    def mut_ratio_sequence(self):
        return [m[1] for m in self._meta]

    def _step_totals(self):
        return [m[2] for m in self._meta]

    def save_step(self, step):
        a = _step_attrs(step)
        self._meta.append((a["phi"], a["mutation_ratio"], a["total_unnorm_log_weight"]))   # known at once
        self._streamer.submit(step, self._write)

    def _write(self, rec):
        pos = os.path.getsize(self._filename) if os.path.exists(self._filename) else 0
        with open(self._filename, "ab") as f:
            pickle.dump(rec, f, pickle.HIGHEST_PROTOCOL)
        self._offsets.append(pos)

    def flush(self):
        """Blocks until every submitted step is in the file."""
        self._streamer.drain()

    def __len__(self):
        return len(self._meta)

    def __getitem__(self, idx):
        self.flush()
        idx = self._index(idx)
        with open(self._filename, "rb") as f:
            f.seek(self._offsets[idx])
            rec = pickle.load(f)
        return _particles_from_record(rec)


class HDF5Storage(BaseStorage):
    """The reference's HDF5 layout (utils/storage.py:79-170); needs h5py."""

    def __init__(self, filename, mode="a"):
        _check_mode(mode)
        super().__init__()
        try:
            import h5py
            if not hasattr(h5py, "File"):
                raise ImportError("h5py stub without File")
        except ImportError as e:              # h5py is not in this image
            raise ImportError("HDF5Storage needs h5py; PickleStorage and InMemoryStorage do not") from e
        self._h5py = h5py
        self._filename = _rank_filename(filename)
        self._mode = mode
        self._len = 0
        self._streamer = StepStreamer()
        if os.path.exists(self._filename) and mode == "a":
            with self._open("r") as h5:
                self._len = len(h5.keys())
            self.is_restart = True

    def _open(self, mode):
        return self._h5py.File(self._filename, mode, track_order=True)

    def flush(self):
        """Blocks until every submitted step is in the file."""
        self._streamer.drain()

    def _attr_list(self, name):
        self.flush()
        with self._open("r") as h5:
            return [h5[str(i)].attrs[name] for i in range(len(h5.keys()))]

    @property
    def phi_sequence(self):
        return self._attr_list("phi")

    @property
    def mut_ratio_sequence(self):
        return self._attr_list("mutation_ratio")

    def _step_totals(self):
        return self._attr_list("total_unnorm_log_weight")

    def __len__(self):
        return self._len

    def save_step(self, step):
        self._len += 1
        self._streamer.submit(step, self._write)

    def _write(self, rec):
        with self._open(self._mode) as h5:
            self._mode = "a"
            grp = h5.create_group(str(len(h5.keys())))
            for k in ("phi", "total_unnorm_log_weight", "mutation_ratio"):     # the reference's three, its order
                grp.attrs[k] = rec["attrs"][k]
            if "rng_stream" in rec["attrs"]:
                grp.attrs["rng_stream"] = rec["attrs"]["rng_stream"]
            grp.create_dataset("log_likes", data=rec["log_likes"])
            grp.create_dataset("log_weights", data=rec["log_weights"])
            pg = grp.create_group("params", track_order=True)
            for k, v in rec["params"].items():
                pg.create_dataset(k, data=v)

    def __getitem__(self, idx):
        self.flush()
        idx = self._index(idx)
        with self._open("r") as h5:
            grp = h5[str(idx)]
            rec = {"params": {k: v[:] for k, v in grp["params"].items()}, "log_likes": grp["log_likes"][:],
                   "log_weights": grp["log_weights"][:], "attrs": {k: v for k, v in grp.attrs.items()}}
        return _particles_from_record(rec)
