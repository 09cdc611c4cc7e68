"""Oracle restatement of Data.next_batch (deepbelief/layers/data.py:75-100) on an in-memory array
(TEST INFRASTRUCTURE — see oracle/__init__.py)."""
import numpy as np


class DataOracle:
    def __init__(self, datapoints, labels=None, shuffle_first=False, batch_size=1):
        self.datapoints = datapoints
        self.labels = labels
        self.num_datapoints = datapoints.shape[0]
        self.batch_size = batch_size
        self.epoch = 0
        self.next_batch_start = 0
        assert batch_size <= self.num_datapoints                      # data.py:59
        self._index_map = np.arange(self.num_datapoints)
        if shuffle_first:
            np.random.shuffle(self._index_map)                        # data.py:64-66 (global numpy RNG)

    def next_batch(self):
        batch_start = self.next_batch_start
        self.next_batch_start += self.batch_size
        if self.next_batch_start > self.num_datapoints:               # data.py:84-90: tail dropped, reshuffle
            self.epoch += 1
            batch_start = 0
            self.next_batch_start = self.batch_size
            np.random.shuffle(self._index_map)
        idx = sorted(list(self._index_map[batch_start:self.next_batch_start]))   # data.py:92-93
        if self.labels is not None:
            return self.datapoints[idx], self.labels[idx]
        return self.datapoints[idx]
