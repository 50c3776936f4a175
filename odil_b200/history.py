"""Training history: column store flushed to train.csv (reference src/odil/history.py)."""
import pickle

import numpy as np


class History:

    def __init__(self, csvpath=None, warmup=0):
        """warmup: rows are written only once more than `warmup` entries exist (late columns)."""
        self.data = dict()
        self.count = 0
        self.warmup = warmup
        self.csvcount = 0
        self.csvpath = csvpath
        self.csvkeys = None
        self.csvfile = open(csvpath, "w") if csvpath is not None else None

    @staticmethod
    def _blank(value):
        if value is None:
            return None
        if isinstance(value, (float, np.floating)):
            return 0.0
        if isinstance(value, (int, np.integer)):
            return 0
        raise ValueError("Unknown type: " + str(type(value)))

    def append(self, key, value=None):
        if hasattr(value, "__array__") and not isinstance(value, (np.ndarray, np.generic)):
            value = np.asarray(value)  # device scalars convert here
        assert value is None or isinstance(value, (int, float, str, np.integer, np.floating, np.ndarray)), \
            "Unexpected type: " + str(type(value))
        if isinstance(value, np.ndarray):
            assert value.shape == (1,) or value.ndim == 0
            value = value.item()
        if key not in self.data:
            assert value is not None
            self.data[key] = [self._blank(value)] * self.count
        if value is None:
            assert len(self.data[key]) > 0, "Expected non-empty column " + key
            value = self._blank(self.data[key][-1])
        self.data[key].append(value)

    def commit(self):
        longest = max(len(v) for v in self.data.values())
        missing = [k for k, v in self.data.items() if len(v) < longest]
        if missing:
            raise RuntimeError("Missing values for columns: " + ",".join(missing))
        self.count += 1

    def get(self, key, default=None):
        return self.data.get(key, default)

    def append_dict(self, newdict):
        for k, v in newdict.items():
            self.append(k, v)

    def write(self, nocommit=False):
        if not nocommit:
            self.commit()
        if self.count <= self.warmup or self.csvfile is None:
            return
        if self.csvkeys is not None and len(self.data) != len(self.csvkeys):
            raise RuntimeError("Unexpected keys in history: {:}".format(list(set(self.data) - set(self.csvkeys))))
        if self.csvcount == 0:
            self.csvkeys = list(self.data.keys())
            self.csvfile.write(",".join(self.csvkeys) + "\n")
        while self.csvcount < self.count:
            self.csvfile.write(",".join(str(self.data[k][self.csvcount]) for k in self.data) + "\n")
            self.csvcount += 1
        self.csvfile.flush()

    def save(self, path):
        with open(path, "wb") as f:
            pickle.dump(self.data, f)

    def load(self, path):
        with open(path, "rb") as f:
            self.data = pickle.load(f)
        self.csvkeys = list(self.data.keys())
        self.count = len(next(iter(self.data.values())))
        self.write(nocommit=True)

    def close(self):
        if self.csvfile:
            self.csvfile.close()
