"""oracle.rbm.propagate_loss_and_grads (RBM.backprop's loss and gradient) against goldens from the reference's own
rbm.py / dist.py: forward value, and the analytic gradient against directional finite differences of that loss."""
import numpy as np
import pytest

from conftest import Golden
from oracle import rbm as orbm

CASES = {'concrete': (0.25, True), 'prob': (None, False)}


@pytest.fixture(scope='module')
def bp_golden():
    return Golden('backprop_golden.npz')


def stack_from(g, T):
    P = {k: g['backprop/P/' + k] for k in ('W1', 'b1', 'c1', 'W2', 'b2', 'c2')}
    return [dict(W=P['W1'], b=P['b1'], c=P['c1'], temperature=T), dict(W=P['W2'], b=P['b2'], c=P['c2'], temperature=T)], P


def get_draws(g):
    return [g['backprop/draw%d' % i] for i in range(4 * int(g['backprop/k']))]


@pytest.mark.parametrize('cname', list(CASES))
def test_backprop_loss_and_gradient(bp_golden, cname):
    g = bp_golden
    T, sample = CASES[cname]
    layers, P = stack_from(g, T)
    k = int(g['backprop/k']) if (T is not None or sample) else 1
    loss, grads, v1 = orbm.propagate_loss_and_grads(layers, g['backprop/v'], k, get_draws(g), sample=sample)
    np.testing.assert_allclose(loss, float(g['backprop/%s/loss' % cname]), rtol=1e-12)
    np.testing.assert_allclose(v1, g['backprop/%s/v1' % cname], rtol=1e-10, atol=1e-14)
    flat = {'W1': grads[0]['W'], 'b1': grads[0]['b'], 'c1': grads[0]['c'], 'W2': grads[1]['W'], 'b2': grads[1]['b'],
            'c2': grads[1]['c']}
    for key, gr in flat.items():
        d = g['backprop/%s/dir/%s' % (cname, key)]
        fd = float(g['backprop/%s/fd/%s' % (cname, key)])
        an = float(np.sum(gr.reshape(d.shape) * d))
        assert abs(an - fd) <= 1e-6 * max(abs(fd), 1.0), (key, an, fd)
