"""Host-side logic that needs no GPU: the tf compat shim, optimizers' descriptors, preprocessing, the
deepbelief alias package, Philox oracle self-checks, and the oracle's optimizer rules."""
import numpy as np
import pytest

from deepbelief_b200 import compat, preprocessing
from oracle import philox, rbm as orbm


def test_compat_session_and_lazy():
    s = compat.Session(config=compat.ConfigProto(gpu_options=compat.GPUOptions(allow_growth=True)))
    calls = []
    lazy = compat.Lazy(lambda: calls.append(1) or np.float32(3.0))
    assert s.run(lazy) == 3.0 and s.run([lazy, 2])[1] == 2
    assert len(calls) == 2                                   # re-evaluated on every run, like a graph node
    ph = compat.placeholder(shape=(None, 2))
    s.run(1, feed_dict={ph: np.ones((3, 2))})
    assert ph.value.shape == (3, 2)
    s.close()


def test_optimizer_descriptors_follow_tf_rules():
    adam = compat.train.AdamOptimizer(learning_rate=1e-3)
    d1, d2 = adam.descriptor(), adam.descriptor()
    assert (d1.step, d2.step) == (1, 2) and abs(d1.beta2 - 0.999) < 1e-7 and abs(d1.eps - 1e-8) < 1e-12
    assert adam.state_multiple() == 2
    adam.reset()
    assert adam.descriptor().step == 1                       # slots re-initialised per train() (rbm.py:107-117)
    assert compat.train.GradientDescentOptimizer(0.1).state_multiple() == 0
    assert compat.train.MomentumOptimizer(0.1, momentum=0.8, weight_decay=1e-4).descriptor().momentum == pytest.approx(0.8)


def test_oracle_optimizers():
    th, g = np.array([1.0, -2.0]), np.array([0.5, 0.25])
    np.testing.assert_allclose(orbm.SGD(0.1).step(th, g), th - 0.1 * g)
    m = orbm.Momentum(0.1, 0.9)
    t1 = m.step(th, g)
    t2 = m.step(t1, g)
    np.testing.assert_allclose(t2, t1 - 0.1 * (0.9 * g + g))
    a = orbm.Adam(1e-3)
    t1 = a.step(th, g)
    # first Adam step moves every coordinate by ~lr in the direction of -sign(g)
    np.testing.assert_allclose(t1, th - 1e-3 * np.sign(g), rtol=1e-6)


def test_preprocessing_matches_reference_goldens(misc_golden):
    A = misc_golden['pre/A']
    np.testing.assert_allclose(preprocessing.standardize(A), misc_golden['pre/std'], rtol=1e-5, atol=1e-6)
    p = preprocessing.PCA(A, 2)
    for col in range(2):
        s = np.sign(np.dot(p[:, col], misc_golden['pre/pca2'][:, col]))
        np.testing.assert_allclose(s * p[:, col], misc_golden['pre/pca2'][:, col], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(preprocessing.inverse_softplus(np.array([0.3, 1.0, 2.5])), misc_golden['pre/inv_softplus'])
    # the reference's own test properties (tests/test_preprocessing.py:12-48)
    arr = np.random.RandomState(0).uniform(0, 255, size=(5, 5))
    b = preprocessing.binarize(arr, 127)
    assert set(np.unique(b)) <= {0.0, 1.0} and b.dtype == np.float32
    s = preprocessing.scale255to1(arr)
    assert s.min() >= 0 and s.max() <= 1
    st = preprocessing.standardize(arr)
    np.testing.assert_allclose(st.mean(0), 0, atol=1e-6)
    np.testing.assert_allclose(st.std(0), 1, atol=1e-5)
    with pytest.raises(AssertionError):
        preprocessing.scale255to1(arr * 2)


def test_alias_package_maps_reference_module_paths():
    import deepbelief  # noqa: F401
    import deepbelief.dist as d
    import deepbelief.layers.data as data
    assert d.__name__ == 'deepbelief_b200.dist' and data.Data.__module__ == 'deepbelief_b200.layers.data'
    # every layer module of the reference (deepbelief/layers/*.py except the unfinished conv_rbm) resolves
    import importlib
    for name, cls in (('rbm', 'RBM'), ('bgrbm', 'BGRBM'), ('gbrbm', 'GBRBM'), ('ggrbm', 'GGRBM'), ('bgrbmsv', 'BGRBMSV'),
                      ('sbm_lower', 'SBM_Lower'), ('gplvm', 'GPLVM'), ('gprbm', 'GPRBM')):
        mod = importlib.import_module('deepbelief.layers.' + name)
        assert hasattr(mod, cls) and mod.__name__ == 'deepbelief_b200.layers.' + name


def test_philox_known_answer_and_row_sharding():
    # Random123 known-answer test for philox4x32-10: counter = key = 0
    out = philox.philox4x32_10(np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    ones = 0xFFFFFFFF
    out = philox.philox4x32_10(np.uint32(ones), np.uint32(ones), np.uint32(ones), np.uint32(ones), ones, ones)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    u = philox.uniform(64, 40, seed=7, draw=3)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    # a row shard with row_offset reproduces the rows of the full draw (G-invariance)
    np.testing.assert_array_equal(philox.uniform(32, 40, seed=7, draw=3, row_offset=32), u[32:])
    n = philox.normal(512, 64, seed=1, draw=0)
    assert abs(n.mean()) < 0.02 and abs(n.std() - 1.0) < 0.02


def test_data_shard_tiles_the_global_batch():
    """Data.shard (data-parallel view): the ranks' batches are the consecutive row blocks of the batch one process
    draws (sorted rows, tail dropped, reshuffle per epoch — data.py:84-94), through several epochs."""
    from deepbelief_b200.layers.data import Data
    rs = np.random.RandomState(3)
    arr = rs.normal(size=(50, 7)).astype(np.float32)
    world, bs = 4, 16
    one = Data.from_arrays(arr, batch_size=bs, log_epochs=False, pin=False)
    one._cursor.random = np.random.RandomState(77)
    shards = [Data.from_arrays(arr, batch_size=bs, log_epochs=False, pin=False).shard(r, world, shuffle_seed=77) for r in range(world)]
    assert shards[0].batch_shape() == (bs // world, 7) and one.batch_shape() == (bs, 7)
    staging = np.empty((bs // world, 7), np.float32)
    for it in range(11):                                       # 3 batches per epoch: crosses epoch boundaries
        want = one.next_batch()
        got = [s.next_batch() if it % 2 else np.array(s.next_batch_into(staging)) for s in shards]
        np.testing.assert_array_equal(np.concatenate(got), want)
    assert one.epoch == shards[0].epoch == 3
    try:
        Data.from_arrays(arr, batch_size=10, pin=False).shard(0, 4)
        assert False
    except AssertionError:
        pass


def test_layers_get_distinct_default_streams():
    """layers built without seed= must not share one Philox stream; explicit seeds are kept."""
    from deepbelief_b200.layers import rbm
    a, b = rbm.default_stream('BBRBM-200'), rbm.default_stream('BBRBM-200-100')
    c = rbm.default_stream('BBRBM-200')
    assert len({a, b, c}) == 3 and (a >> 32) == (c >> 32) != (b >> 32)


def test_data_binary_bytes_and_byte_batches():
    """Data.binary_bytes(): a uint8 copy exists exactly for {0,1} data sets, and next_batch_into(as_bytes=True) walks the
    same rows as next_batch() (data.py:78-105 batching rule) for contiguous windows and gathered ones."""
    import numpy as np
    from deepbelief_b200.layers.data import Data
    rs = np.random.RandomState(0)
    X = (rs.uniform(size=(37, 24)) < 0.5).astype(np.float32)
    for shuffle in (False, True):
        np.random.seed(4)                                            # the epoch reshuffles draw from the global numpy state
        a = Data.from_arrays(X, batch_size=8, shuffle_first=shuffle, log_epochs=False, pin=False)
        wanted = [np.array(a.next_batch()) for _ in range(12)]       # crosses epoch boundaries (reshuffles)
        np.random.seed(4)
        b = Data.from_arrays(X, batch_size=8, shuffle_first=shuffle, log_epochs=False, pin=False)
        assert b.binary_bytes() is not None and b.binary_bytes().dtype == np.uint8
        assert b.datapoints.dtype == np.float32                     # what the reference exposes is untouched
        out = np.empty((8, 24), dtype=np.uint8)
        for want in wanted:
            got = b.next_batch_into(out, as_bytes=True)
            assert got.dtype == np.uint8 and np.array_equal(got.astype(np.float32), want)
    c = Data.from_arrays(X * 0.5, batch_size=8, log_epochs=False, pin=False)
    assert c.binary_bytes() is None
    d = Data.from_arrays(X - 1.0, batch_size=8, log_epochs=False, pin=False)
    assert d.binary_bytes() is None
