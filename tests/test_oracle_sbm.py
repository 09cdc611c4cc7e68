"""The SBM_Lower oracle against the goldens produced by the reference's own sbm_lower.py (CPU, no GPU needed)."""
import os

import numpy as np
import pytest

from oracle import sbm as osbm

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def golden():
    z = np.load(os.path.join(HERE, 'golden', 'sbm_golden.npz'))
    return {k.replace('__', '/'): z[k] for k in z.files}


@pytest.mark.parametrize('tag', ['s8', 's12'])
def test_oracle_matches_reference_outputs(golden, tag):
    pre = 'sbm/%s/' % tag
    side, ov, H, N = [int(x) for x in golden[pre + 'dims']]
    W, b = golden[pre + 'W'], golden[pre + 'b']
    v0, vk, h = golden[pre + 'v0'], golden[pre + 'vk'], golden[pre + 'h']
    for dname, tol in (('float64', 1e-12), ('float32', 2e-5)):
        o = osbm.SBMLowerOracle(W, b, side, ov, H, dtype=np.dtype(dname).type)
        q = pre + dname + '/'
        np.testing.assert_allclose(o.h_net_input(v0), golden[q + 'h_net'], rtol=tol, atol=tol)
        np.testing.assert_allclose(o.v_net_input(h), golden[q + 'v_net'], rtol=tol, atol=tol)
        np.testing.assert_allclose(o.h_probs(v0), golden[q + 'h_probs'], rtol=tol, atol=tol)
        np.testing.assert_allclose(o.v_probs(h), golden[q + 'v_probs'], rtol=tol, atol=tol)
        np.testing.assert_allclose(o.free_energy(v0), golden[q + 'free_energy'], rtol=tol, atol=tol)
        np.testing.assert_allclose(o.cd_loss(v0, vk), golden[q + 'cd_loss'], rtol=max(tol, 1e-10), atol=10 * tol)


@pytest.mark.parametrize('tag', ['s8', 's12'])
def test_shared_gradient_matches_finite_differences_of_the_reference_loss(golden, tag):
    pre = 'sbm/%s/' % tag
    side, ov, H, N = [int(x) for x in golden[pre + 'dims']]
    o = osbm.SBMLowerOracle(golden[pre + 'W'], golden[pre + 'b'], side, ov, H)
    gW, gb, gc = o.cd_grad_shared(golden[pre + 'v0'], golden[pre + 'vk'])
    np.testing.assert_allclose(gW, golden[pre + 'fd_gW'], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(gb.ravel(), golden[pre + 'fd_gb'], rtol=1e-6, atol=1e-8)
    assert not gc.any()


def test_dense_weights_place_every_patch():
    W = np.arange(1, 26, dtype=np.float64).reshape(25, 1) * np.ones((1, 3))
    dense = osbm.dense_weights(W, 8, 2, 12)
    assert dense.shape == (64, 12)
    assert np.count_nonzero(dense[:, 0]) == 25 and np.count_nonzero(dense) == 4 * 25 * 3
    assert dense[0, 0] == 1 and dense[3, 3] == 1 and dense[3 * 8 + 3, 6] == 1 and dense[3 * 8, 9] == 1   # patch origins
