"""GPU parity of SBM_Lower (patch-shared weights, sbm_lower.py:59-117) against the oracle, the goldens produced by the
reference class itself, and — for training — the oracle's Adam trajectory with injected draws."""
import os

import numpy as np
import pytest
import torch

from conftest import has_cuda
from oracle import sbm as osbm, rbm as orbm

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')]
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def golden():
    z = np.load(os.path.join(HERE, 'golden', 'sbm_golden.npz'))
    return {k.replace('__', '/'): z[k] for k in z.files}


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def _layer(side, ov, H, W, b, **kw):
    from deepbelief_b200.layers.sbm_lower import SBM_Lower
    lay = SBM_Lower(side, ov, H, **kw)
    lay.set_weights(W=W.astype(np.float32), b=b.astype(np.float32))
    return lay


@pytest.mark.parametrize('tag', ['s8', 's12'])
def test_forward_quantities_match_reference_goldens(golden, tag):
    pre = 'sbm/%s/' % tag
    side, ov, H, N = [int(x) for x in golden[pre + 'dims']]
    lay = _layer(side, ov, H, golden[pre + 'W'], golden[pre + 'b'])
    assert tuple(lay.W.shape) == ((side // 2 + ov // 2) ** 2, H // 4)          # the shared matrix is what the API exposes
    v0, vk, h = golden[pre + 'v0'], golden[pre + 'vk'], golden[pre + 'h']
    q = pre + 'float64/'
    np.testing.assert_allclose(_np(lay.h_net_input(v0)), golden[q + 'h_net'], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(_np(lay.v_net_input(h)), golden[q + 'v_net'], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(_np(lay.h_probs(v0)), golden[q + 'h_probs'], rtol=1e-5)
    np.testing.assert_allclose(_np(lay.v_probs(h)), golden[q + 'v_probs'], rtol=1e-5)
    np.testing.assert_allclose(_np(lay.free_energy(v0)), golden[q + 'free_energy'], rtol=1e-5)
    np.testing.assert_allclose(float(lay.cd_loss(v0, vk)), float(golden[q + 'cd_loss']), rtol=1e-4, atol=1e-5)
    # gradient of cd_loss wrt the shared parameters against finite differences of the reference's own loss
    g = _np(lay.cd_gradient(v0, vk))
    nw = lay.W.numel()
    gW, gb, gc = g[:nw].reshape(tuple(lay.W.shape)), g[nw:nw + side * side], g[nw + side * side:]
    assert np.max(np.abs(gW - golden[pre + 'fd_gW'])) / np.max(np.abs(golden[pre + 'fd_gW'])) < 1e-4
    assert np.max(np.abs(gb - golden[pre + 'fd_gb'])) / np.max(np.abs(golden[pre + 'fd_gb'])) < 1e-4
    assert not gc.any()


def test_sampling_with_injected_draws_is_bit_exact(golden):
    pre = 'sbm/s8/'
    side, ov, H, N = [int(x) for x in golden[pre + 'dims']]
    lay = _layer(side, ov, H, golden[pre + 'W'], golden[pre + 'b'])
    o = osbm.SBMLowerOracle(golden[pre + 'W'].astype(np.float32), golden[pre + 'b'].astype(np.float32), side, ov, H)
    rs = np.random.RandomState(5)
    v0 = golden[pre + 'v0']
    u = rs.rand(N, H).astype(np.float32)
    p64 = o.h_probs(v0)
    got = _np(lay.sample_h(lay.h_probs(v0), draw=torch.from_numpy(u).cuda()))
    want = (p64 > u).astype(np.float64)
    safe = np.abs(p64 - u) > 1e-6                                 # tie band
    assert np.array_equal(got[safe], want[safe])
    assert set(np.unique(got)) <= {0.0, 1.0}


def test_cd_training_follows_oracle_adam_on_the_shared_parameters():
    """Three CD-1 Adam updates with injected uniforms: shared W and b track the oracle, c stays zero, and the dense
    kernel weight is always the scatter of the shared matrix."""
    from deepbelief_b200 import compat
    side, ov, H, N, k = 8, 2, 16, 12, 1
    V = side * side
    rs = np.random.RandomState(11)
    P = side // 2 + ov // 2
    W = rs.normal(size=(P * P, H // 4)) * 0.2
    b = rs.normal(size=V) * 0.1
    lay = _layer(side, ov, H, W, b)
    adam = compat.AdamOptimizer(learning_rate=1e-2)
    state = lay.new_optimizer_state(adam)
    oW, ob = W.astype(np.float32).astype(np.float64), b.astype(np.float32).astype(np.float64)
    oadam = orbm.Adam(1e-2)
    for it in range(3):
        v0 = rs.binomial(1, 0.4, size=(N, V)).astype(np.float32)
        draws = rs.rand(N * H + k * (N * V + N * H)).astype(np.float32)
        grad, vk, _ = lay.cd_step(torch.from_numpy(v0).cuda(), k, injected=torch.from_numpy(draws).cuda(), force_simt=True)
        lay.apply_gradient(adam, grad, state)
        o = osbm.SBMLowerOracle(oW, ob, side, ov, H)
        gW, gb, gc = o.cd_grad_shared(v0, _np(vk))                 # gradient at the GPU's own chain end
        theta = np.concatenate([oW.ravel(), ob.ravel(), np.zeros(H)])
        theta = oadam.step(theta, np.concatenate([gW.ravel(), gb.ravel(), gc.ravel()]))
        oW, ob = theta[:oW.size].reshape(oW.shape), theta[oW.size:oW.size + V]
        np.testing.assert_allclose(_np(lay.W), oW, rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(_np(lay.b).ravel(), ob, rtol=2e-4, atol=2e-6)
        assert not _np(lay.c).any()
        np.testing.assert_array_equal(_np(lay.Wk), osbm.dense_weights(_np(lay.W), side, ov, H))


def test_train_loop_checkpoint_and_tensor_core_path(tmp_path):
    """RBM.train drives the layer (side 16 -> V = 256, H = 64, batch 64 takes the tcgen05 CD path on the dense weight);
    the checkpoint stores the shared parameters and restores the same layer."""
    from deepbelief_b200 import compat
    from deepbelief_b200.layers.data import Data
    from deepbelief_b200.layers.sbm_lower import SBM_Lower
    rs = np.random.RandomState(2)
    X = (rs.rand(256, 256) < 0.3).astype(np.float32)
    lay = SBM_Lower(16, 4, 64, W_std=0.05, seed=3)
    w_before = _np(lay.W).copy()
    lay.train(compat.AdamOptimizer(learning_rate=1e-3), Data.from_arrays(X, batch_size=64), None, num_gibbs_steps=2, pcd=False,
              max_iterations=5, ckpt_interval=5, ckpt_dir=str(tmp_path))
    assert not np.array_equal(_np(lay.W), w_before) and np.isfinite(_np(lay.W)).all()
    assert not _np(lay.c).any()
    np.testing.assert_array_equal(_np(lay.Wk), osbm.dense_weights(_np(lay.W), 16, 4, 64))
    other = SBM_Lower(16, 4, 64, name=lay.name)
    other.restore(os.path.join(str(tmp_path), '%s-5' % lay.name))
    np.testing.assert_array_equal(_np(other.W), _np(lay.W))
    np.testing.assert_array_equal(_np(other.Wk), _np(lay.Wk))
