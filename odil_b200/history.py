"""
Training history behind `make_callback`: named scalar series, one entry per reported epoch, streamed to `train.csv`
and picklable as a plain `{name: [values]}` dict.  Same public behaviour and file bytes as the reference's `History`
(src/odil/history.py:9-123: `append` :35-56, `commit` :58-69, `write` :78-102, `save` / `load` :104-120) --
`tests/test_io_cpu.py::test_history_matches_reference_bytes` replays a scripted session through both.

Differences that matter on this backend: values that live on the device (the lazily fetched loss / norm scalars of
`Problem.eval_loss_grad`) are accepted and converted when they are appended, so a callback that does not record
history never synchronises.
"""
import numbers
import pickle

import numpy as np


def _filler(sample):
    """What a series is padded with for entries it did not take part in: a zero of the kind of `sample`."""
    if sample is None:
        return None
    if isinstance(sample, (float, np.floating)):
        return 0.0
    if isinstance(sample, numbers.Integral):
        return 0
    raise ValueError("Unknown type: " + str(type(sample)))


def _scalar(value):
    """Plain Python / NumPy scalar (or str / None) from whatever the caller recorded."""
    if value is None or isinstance(value, (str, numbers.Real, np.generic)):
        return value
    if not isinstance(value, np.ndarray):
        if not hasattr(value, "__array__"):
            raise AssertionError("Unexpected type: " + str(type(value)))
        value = np.asarray(value)  # device-resident scalar: this is where it reaches the host
    if value.ndim != 0 and value.shape != (1,):
        raise AssertionError("Expected a scalar, got shape " + str(value.shape))
    return value.item()


class _CsvStream:
    """Append-only CSV file: the header is fixed by the first row that is written."""

    def __init__(self, path):
        self.path = path
        self.file = open(path, "w") if path is not None else None
        self.columns = None
        self.rows = 0

    def emit(self, series, upto):
        names = list(series)
        if self.columns is not None and len(names) != len(self.columns):
            raise RuntimeError("Unexpected keys in history: {:}".format(list(set(names) - set(self.columns))))
        if self.rows == 0:
            self.columns = names
            self.file.write(",".join(names) + "\n")
        for i in range(self.rows, upto):
            self.file.write(",".join(str(series[name][i]) for name in names) + "\n")
        self.rows = max(self.rows, upto)
        self.file.flush()


class History:

    def __init__(self, csvpath=None, warmup=0):
        """
        csvpath: file the entries are streamed to (None: keep them in memory only).
        warmup: nothing is written until more than `warmup` entries are complete, so that series which first
        appear in the second entry still get a column.
        """
        self.data = {}    # name -> list of values, all of length `count` between entries
        self.count = 0    # complete entries
        self.warmup = warmup
        self._csv = _CsvStream(csvpath)

    # names the reference exposes
    csvpath = property(lambda self: self._csv.path)
    csvfile = property(lambda self: self._csv.file)
    csvkeys = property(lambda self: self._csv.columns)
    csvcount = property(lambda self: self._csv.rows)

    def append(self, key, value=None):
        """Records `value` under `key` in the entry being assembled.  A new key is back-filled with zeros for the
        entries it missed; `value=None` repeats a zero of the series' kind."""
        value = _scalar(value)
        series = self.data.get(key)
        if series is None:
            assert value is not None
            series = self.data[key] = [_filler(value)] * self.count
        if value is None:
            assert series, "Expected non-empty column " + key
            value = _filler(series[-1])
        series.append(value)

    def append_dict(self, newdict):
        for key, value in newdict.items():
            self.append(key, value)

    def commit(self):
        """Closes the entry: every series must have taken part in it."""
        lengths = {key: len(series) for key, series in self.data.items()}
        full = max(lengths.values())
        behind = "".join(key + "," for key, n in lengths.items() if n < full)
        if behind:
            raise RuntimeError("Missing values for columns: " + behind)
        self.count += 1

    def get(self, key, default=None):
        return self.data.get(key, default)

    def write(self, nocommit=False):
        """Closes the entry (unless `nocommit`) and streams the entries not yet in the CSV file."""
        if not nocommit:
            self.commit()
        if self.count > self.warmup and self._csv.file is not None:
            self._csv.emit(self.data, self.count)

    def save(self, path):
        with open(path, "wb") as f:
            pickle.dump(self.data, f)

    def load(self, path):
        """Replaces the series by a saved dict and streams them out."""
        with open(path, "rb") as f:
            self.data = pickle.load(f)
        self._csv.columns = list(self.data)
        self.count = len(next(iter(self.data.values())))
        self.write(nocommit=True)

    def close(self):
        if self._csv.file:
            self._csv.file.close()
