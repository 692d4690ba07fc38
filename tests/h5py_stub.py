"""
Minimal stand-in for h5py (absent from the image) covering what HDF5Storage -- ours and the reference's
(smcpy/utils/storage.py:79-170) -- uses: File(filename, mode, track_order=...) as a context manager or with
close(), nested groups with ordered keys()/items(), attrs, create_group, create_dataset(name, data=...),
dataset[:] reads.  A "file" is a pickle of the group tree, written on close.  Test infrastructure only: it
lets the HDF5 layout be written by one implementation and read by the other without the real library.
"""

import os
import pickle
import types

import numpy as np


class Dataset:
    def __init__(self, data):
        self._a = np.array(data)

    def __getitem__(self, key):
        return self._a[key]

    @property
    def shape(self):
        return self._a.shape


class Group:
    def __init__(self):
        self.attrs = {}
        self._items = {}

    def create_group(self, name, track_order=False):
        if name in self._items:
            raise ValueError("unable to create group (name already exists)")
        g = Group()
        self._items[name] = g
        return g

    def create_dataset(self, name, data=None):
        if name in self._items:
            raise ValueError("unable to create dataset (name already exists)")
        d = Dataset(data)
        self._items[name] = d
        return d

    def keys(self):
        return self._items.keys()

    def items(self):
        return self._items.items()

    def __getitem__(self, name):
        return self._items[name]

    def __contains__(self, name):
        return name in self._items

    def __len__(self):
        return len(self._items)


class File(Group):
    def __init__(self, filename, mode="r", track_order=False):
        super().__init__()
        self._filename, self._mode = str(filename), mode
        exists = os.path.exists(self._filename)
        if mode == "r" and not exists:
            raise FileNotFoundError(self._filename)
        if mode in ("r", "a", "r+") and exists:
            with open(self._filename, "rb") as f:
                root = pickle.load(f)
            self.attrs, self._items = root.attrs, root._items
        elif mode not in ("w", "a"):
            raise ValueError("mode %r" % mode)

    def close(self):
        if self._mode != "r":
            root = Group()
            root.attrs, root._items = self.attrs, self._items
            with open(self._filename, "wb") as f:
                pickle.dump(root, f)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def module():
    m = types.ModuleType("h5py")
    m.File, m.Group, m.Dataset = File, Group, Dataset
    m.__stub__ = True
    return m
