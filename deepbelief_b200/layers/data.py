"""In-memory dataset with the reference's epoch/batch contract, fed from the built-in HDF5 reader.

Contract kept from deepbelief/layers/data.py (reference):
  * a batch is `batch_size` rows picked through an index permutation, and the picked row numbers
    are SORTED before the gather (data.py:92-93);
  * when the next window would run past the end, the partial tail is dropped, `epoch` increments and
    the permutation is reshuffled with the global `np.random` state (data.py:84-90);
  * the first epoch uses the identity permutation unless `shuffle_first` (data.py:64-66);
  * `has_labels` makes `next_batch()` return `(datapoints, labels)`.
Added for the GPU path: `Data.from_arrays(...)` builds the same object from in-memory arrays (kept in
pinned host memory, so the trainer's prefetching feeder can copy host->device asynchronously without a
staging pass), and batches whose sorted row numbers are contiguous are returned as zero-copy views.
"""
import json
import logging
import os

import numpy as np

from .. import hdf5


class _EpochCursor:
    """Produces the sorted row indices of successive batches."""

    def __init__(self, num_rows, batch_size, shuffle_first):
        self.order = np.arange(num_rows)
        self.num_rows, self.batch_size = num_rows, batch_size
        self.start = 0        # first position of the NEXT window inside `order`
        self.epoch = 0
        self.random = np.random    # the reference reshuffles with the GLOBAL numpy state (data.py:84-90)
        if shuffle_first:
            np.random.shuffle(self.order)

    def advance(self):
        lo, hi = self.start, self.start + self.batch_size
        wrapped = hi > self.num_rows
        if wrapped:           # drop the tail, start a new shuffled epoch
            self.epoch += 1
            lo, hi = 0, self.batch_size
            self.random.shuffle(self.order)
        self.start = hi
        return np.sort(self.order[lo:hi]), wrapped


class Data:
    """Args as in the reference (data.py:28-34): data_fname, has_labels=False, shuffle_first=False,
    batch_size=1, log_epochs=True, name='Data'."""

    def __init__(self, data_fname, has_labels=False, shuffle_first=False, batch_size=1, log_epochs=True,
                 name='Data'):
        if not os.path.isfile(data_fname):
            raise AssertionError('The file does not exist: %s' % data_fname)
        self.name = name
        self.logger = logging.getLogger(name)
        self.has_labels = bool(has_labels)
        self.batch_size = int(batch_size)
        self.log_epochs = log_epochs

        contents = hdf5.File(data_fname)
        self._setup(contents['datapoints'], contents['labels'] if self.has_labels else None, shuffle_first)

    @classmethod
    def from_arrays(cls, datapoints, labels=None, shuffle_first=False, batch_size=1, log_epochs=True, name='Data',
                    pin=True):
        """Same contract as the file-backed constructor, over arrays already in memory. With `pin` the
        datapoints are moved once into page-locked host memory (needs a CUDA runtime; skipped without)."""
        self = cls.__new__(cls)
        self.name = name
        self.logger = logging.getLogger(name)
        self.has_labels = labels is not None
        self.batch_size = int(batch_size)
        self.log_epochs = log_epochs
        datapoints = np.ascontiguousarray(datapoints, dtype=np.float32)
        self._pinned_store = None
        if pin:
            try:
                import torch
                if torch.cuda.is_available():
                    self._pinned_store = torch.from_numpy(datapoints).pin_memory()
                    datapoints = self._pinned_store.numpy()
            except Exception:       # pinning is an optimisation only
                self._pinned_store = None
        self._setup(datapoints, None if labels is None else np.asarray(labels, dtype=np.float32), shuffle_first)
        return self

    def _setup(self, datapoints, labels, shuffle_first):
        self.datapoints = datapoints
        self.labels = labels
        self.num_datapoints, self.datapoint_size = self.datapoints.shape
        if self.batch_size > self.num_datapoints:
            raise AssertionError('batch_size %d exceeds the %d datapoints' % (self.batch_size, self.num_datapoints))

        self._cursor = _EpochCursor(self.num_datapoints, self.batch_size, shuffle_first)
        self._pinned = None
        self._u8 = None             # uint8 copy of a {0,1} data set (binary_bytes), built on first request
        self._u8_checked = False
        self._shard = None          # (rank, world, local rows) once shard() has been called
        self.logger.debug(self)

    # ---- data-parallel view (no counterpart in the reference, which is single-process) -----------------------------
    def shard(self, rank, world, shuffle_seed=None):
        """Row-shard every batch over `world` data-parallel ranks: the GLOBAL batch is still `batch_size` rows chosen by
        the reference's rule (sorted rows of the permutation window, tail dropped, reshuffle per epoch), and this rank
        keeps rows [rank * b, (rank + 1) * b) of it, b = batch_size / world — so N ranks train on exactly the batches
        one process would see. Every rank must walk the same permutations: epoch reshuffles switch from the global
        numpy state to a private RandomState(shuffle_seed) (RBM.train broadcasts rank 0's seed)."""
        rank, world = int(rank), int(world)
        assert 0 <= rank < world
        assert self.batch_size % world == 0, 'batch_size %d is not divisible by %d ranks' % (self.batch_size, world)
        assert not self.has_labels, 'sharding is implemented for unlabelled training data'
        self._shard = (rank, world, self.batch_size // world)
        if shuffle_seed is not None:
            self._cursor.random = np.random.RandomState(int(shuffle_seed) & 0x7fffffff)
        return self

    def _local(self, rows):
        if self._shard is None:
            return rows
        rank, _, b = self._shard
        return rows[rank * b:(rank + 1) * b]

    @staticmethod
    def _take(arr, rows):
        # sorted row numbers (data.py:92-93) that happen to be contiguous need no gather
        if len(rows) and int(rows[-1]) - int(rows[0]) + 1 == len(rows):
            return arr[int(rows[0]):int(rows[-1]) + 1]
        return arr[rows]

    # attributes the reference exposes
    @property
    def epoch(self):
        return self._cursor.epoch

    @property
    def next_batch_start(self):
        return self._cursor.start

    @property
    def _index_map(self):
        return self._cursor.order

    def batch_shape(self):
        return (self._shard[2] if self._shard else self.batch_size), self.datapoint_size

    def next_batch(self):
        rows, wrapped = self._cursor.advance()
        rows = self._local(rows)
        if wrapped and self.log_epochs:
            self.logger.info('Epoch: %d', self._cursor.epoch)
        if self.has_labels:
            return self._take(self.datapoints, rows), self._take(self.labels, rows)
        return self._take(self.datapoints, rows)

    def binary_bytes(self):
        """uint8 copy of the datapoints when every value is exactly 0 or 1 (the binary data sets the RBM examples train
        on), else None. Built once, in page-locked memory when a CUDA runtime is present: the trainer's feeder then moves
        1 byte per pixel to the device instead of 4 and the tensor-core chain reads the bytes directly (db_cd_tc_desc_t.v0_u8).
        The float32 `datapoints` array the reference exposes is unchanged."""
        if not self._u8_checked:
            self._u8_checked = True
            d = self.datapoints
            if d.size and float(d.min()) >= 0.0 and float(d.max()) <= 1.0:
                u8 = d.astype(np.uint8)
                if np.array_equal(u8, d):
                    try:
                        import torch
                        if torch.cuda.is_available():
                            self._u8_store = torch.from_numpy(u8).pin_memory()
                            u8 = self._u8_store.numpy()
                    except Exception:       # pinning is an optimisation only
                        pass
                    self._u8 = u8
        return self._u8

    def next_batch_into(self, out, as_bytes=False):
        """The datapoints of next_batch() without the intermediate copy: a zero-copy view when the sorted
        rows are contiguous, otherwise gathered by the native multi-threaded row gather
        (db_host_gather_rows) into `out` (a [batch_size, datapoint_size] array, normally the feeder's pinned staging
        buffer: float32, or uint8 with as_bytes, which reads the binary_bytes() copy; or a callable returning that array,
        called only when a gather is needed). Returns the array that holds the batch."""
        rows, wrapped = self._cursor.advance()
        rows = self._local(rows)
        if wrapped and self.log_epochs:
            self.logger.info('Epoch: %d', self._cursor.epoch)
        src = self.binary_bytes() if as_bytes else self.datapoints
        assert src is not None, 'as_bytes needs a {0,1} data set'
        if len(rows) and int(rows[-1]) - int(rows[0]) + 1 == len(rows):
            return src[int(rows[0]):int(rows[-1]) + 1]
        if callable(out):
            out = out()
        if (src.dtype != out.dtype or src.dtype not in (np.float32, np.uint8) or not src.flags['C_CONTIGUOUS'] or
                not out.flags['C_CONTIGUOUS']):
            np.take(src, rows, axis=0, out=out)
            return out
        import ctypes
        from .. import _lib
        rows64 = np.ascontiguousarray(rows, dtype=np.int64)
        rc = _lib.lib().db_host_gather_rows(ctypes.c_void_p(src.ctypes.data), src.shape[1] * src.itemsize,
                                            ctypes.c_void_p(rows64.ctypes.data), len(rows64), src.shape[0],
                                            ctypes.c_void_p(out.ctypes.data), 0)
        if rc != 0:
            raise RuntimeError('db_host_gather_rows failed (%d)' % rc)
        return out

    def __str__(self):
        def stats(arr, tag):
            return {tag + '_array_dtype': str(arr.dtype), tag + '_array_min': float(arr.min()),
                    tag + '_array_max': float(arr.max())}
        d = dict(batch_size=self.batch_size, num_datapoints=self.num_datapoints, datapoint_size=self.datapoint_size)
        d.update(stats(self.datapoints, 'datapoint'))
        if self.has_labels:
            d.update(stats(self.labels, 'label'))
        return json.dumps(d, indent=4, sort_keys=True)
