"""Shared settings of the example scripts: dataset paths, output directory, iteration override."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATASETS = os.path.join(ROOT, 'tests', 'golden', 'datasets')
TRAINING_DATA = os.path.join(DATASETS, 'weizmann_horses_training_328x1024_binary.h5')
TEST_DATA = os.path.join(DATASETS, 'weizmann_horses_eslami_test_14x1024_binary.h5')
IMG_SHAPE = (32, 32)


def iterations(default):
    return int(os.environ.get('EXAMPLE_ITERS', default))


def out_dir(name):
    path = os.path.join(os.environ.get('EXAMPLE_OUT', os.path.join(os.getcwd(), 'example_out')), name)
    os.makedirs(path, exist_ok=True)
    return path
