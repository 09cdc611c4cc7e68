"""HDF5 reader/writer and the Data batching contract against the oracle restatement."""
import os

import numpy as np
import pytest

from conftest import DATASETS
from deepbelief_b200 import hdf5
from deepbelief_b200.layers.data import Data
from oracle.data import DataOracle


def test_reader_on_bundled_files():
    f = hdf5.File(os.path.join(DATASETS, 'weizmann_horses_training_328x1024_binary.h5'))
    d = f['datapoints']
    assert d.shape == (328, 1024) and d.dtype == np.float32
    assert set(np.unique(d)) == {0.0, 1.0}
    np.testing.assert_allclose(d.mean(), 0.3995, atol=1e-4)          # SURVEY §8c fixture facts
    assert int((d.std(axis=0) == 0).sum()) == 13                     # 13 constant columns
    m = hdf5.File(os.path.join(DATASETS, 'mnist_test_30x728_equal_classes_binary.h5'))
    assert m['datapoints'].shape == (30, 784)
    assert m['labels'].ravel().astype(int).tolist() == [7, 2, 1, 0, 4, 1, 4, 9, 5, 9, 0, 6, 9, 0, 1, 5, 7, 3, 4, 6,
                                                         6, 5, 7, 3, 3, 2, 2, 8, 8, 8]
    t = hdf5.File(os.path.join(DATASETS, 'weizmann_horses_eslami_test_14x1024_binary.h5'))
    assert t['datapoints'].shape == (14, 1024)
    # the parser walks the structure: raw bytes at the documented offset agree
    raw = np.fromfile(os.path.join(DATASETS, 'weizmann_horses_training_328x1024_binary.h5'), dtype='<f4',
                      count=328 * 1024, offset=2144).reshape(328, 1024)
    np.testing.assert_array_equal(raw, d)


def test_writer_round_trip(tmp_path):
    rs = np.random.RandomState(0)
    a = rs.rand(37, 11).astype(np.float32)
    lab = rs.randint(0, 10, size=(37, 1)).astype(np.float32)
    p = str(tmp_path / 'x.h5')
    hdf5.write(p, {'datapoints': a, 'labels': lab})
    f = hdf5.File(p)
    assert sorted(f.keys()) == ['datapoints', 'labels']
    np.testing.assert_array_equal(f['datapoints'], a)
    np.testing.assert_array_equal(f['labels'], lab)
    with pytest.raises(KeyError):
        f['nope']


def test_rejects_non_hdf5(tmp_path):
    p = tmp_path / 'bad.h5'
    p.write_bytes(b'not an hdf5 file at all')
    with pytest.raises(hdf5.HDF5FormatError):
        hdf5.File(str(p))


@pytest.mark.parametrize('batch,shuffle_first', [(100, False), (328, False), (50, True), (1, False)])
def test_next_batch_matches_oracle(batch, shuffle_first):
    path = os.path.join(DATASETS, 'weizmann_horses_training_328x1024_binary.h5')
    arr = hdf5.File(path)['datapoints']
    # both draw from the GLOBAL numpy RNG (data.py:66,90), so they are run one after the other
    np.random.seed(123)
    d = Data(path, shuffle_first=shuffle_first, batch_size=batch, log_epochs=False)
    assert d.batch_shape() == (batch, 1024)
    got = [(d.next_batch(), d.epoch, d.next_batch_start) for _ in range(12)]
    np.random.seed(123)
    o = DataOracle(arr, shuffle_first=shuffle_first, batch_size=batch)
    for b, epoch, start in got:
        np.testing.assert_array_equal(b, o.next_batch())
        assert epoch == o.epoch and start == o.next_batch_start


def test_labels_and_edge_cases():
    path = os.path.join(DATASETS, 'mnist_test_30x728_equal_classes_binary.h5')
    d = Data(path, has_labels=True, batch_size=7, log_epochs=False)
    x, y = d.next_batch()
    assert x.shape == (7, 784) and y.shape == (7, 1)
    with pytest.raises(AssertionError):
        Data(path, batch_size=31)                      # data.py:59
    with pytest.raises(AssertionError):
        Data('/nonexistent.h5')
    # the tail is dropped: 30 rows, batch 7 -> 4 batches per epoch, 5th call starts epoch 1 at row 0
    d = Data(path, batch_size=7, log_epochs=False)
    for _ in range(4):
        d.next_batch()
    assert d.epoch == 0
    d.next_batch()
    assert d.epoch == 1 and d.next_batch_start == 7
