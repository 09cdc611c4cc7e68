import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
DATASETS = os.path.join(GOLDEN, 'datasets')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


class Golden:
    """npz archive whose keys use '/' (stored as '__')."""

    def __init__(self, name):
        self._z = np.load(os.path.join(GOLDEN, name))

    def __getitem__(self, key):
        return self._z[key.replace('/', '__')]

    def __contains__(self, key):
        return key.replace('/', '__') in self._z.files


@pytest.fixture(scope='session')
def rbm_golden():
    return Golden('rbm_golden.npz')


@pytest.fixture(scope='session')
def gp_golden():
    return Golden('gp_golden.npz')


@pytest.fixture(scope='session')
def misc_golden():
    return Golden('misc_golden.npz')


@pytest.fixture(scope='session')
def horses():
    from deepbelief_b200 import hdf5
    return hdf5.File(os.path.join(DATASETS, 'weizmann_horses_training_328x1024_binary.h5'))['datapoints']


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
