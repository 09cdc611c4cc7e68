"""The GPDBN oracle (oracle/gprbm.py) against goldens produced by the reference's own gprbm.py on the numpy TF
shim (tests/golden/make_golden.py:gpdbn_goldens): forward quantities, and the analytic gradient against directional
finite differences of the REFERENCE's loss."""
import numpy as np
import pytest

from conftest import Golden
from oracle import gprbm as ogprbm

CASES = {'fixed': dict(trainable=(False, False, False), fixed_noise=0.05),
         'trainable': dict(trainable=(True, True, True), fixed_noise=0.0)}


@pytest.fixture(scope='module')
def gpdbn_golden():
    return Golden('gpdbn_golden.npz')


def make_oracle(g, cname, dtype=np.float64, params=None):
    P = {k: g['gpdbn/P/' + k] for k in ('W0', 'b0', 'c0', 'W1', 'b1', 'c1', 'Wt', 'bt', 'ct', 'X')}
    if params:
        P.update(params)
    hyp = g['gpdbn/%s/hyp_opt' % cname]
    D = P['Wt'].shape[1]
    lower = [dict(W=P['W0'], b=P['b0'], c=P['c0']), dict(W=P['W1'], b=P['b1'], c=P['c1'])]
    top = dict(W=P['Wt'], b=P['bt'], c=P['ct'], opt_alpha_h=np.full(D, np.log(np.e - 1.0)))
    gp = dict(X=P['X'], opt_alpha=hyp[0], opt_gamma=hyp[1], opt_noise=hyp[2], jitter=1e-6, **CASES[cname])
    return ogprbm.GPDBNOracle(lower, top, gp, g['gpdbn/V'], float(g['gpdbn/T']), dtype=dtype)


def get_draws(g):
    d = {k: g['gpdbn/draws/' + k] for k in ('up0', 'up1', 'eps1', 'eps2', 'down_top', 'down1', 'down0')}
    return {'up': [d['up0'], d['up1']], 'eps1': d['eps1'], 'eps2': d['eps2'], 'down_top': d['down_top'],
            'down': [d['down1'], d['down0']]}


@pytest.mark.parametrize('cname', list(CASES))
def test_forward_matches_reference(gpdbn_golden, cname):
    g = gpdbn_golden
    st = make_oracle(g, cname).forward(get_draws(g))
    pre = 'gpdbn/%s/float64/' % cname
    np.testing.assert_allclose(st['pv'], g[pre + 'pred_var'], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(st['sigma_h'], g[pre + 'sigma_h'], rtol=1e-9)
    np.testing.assert_allclose(st['H1'], g[pre + 'H1'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st['H2'], g[pre + 'H2'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st['logdet'], g[pre + 'log_det'], rtol=1e-11)
    np.testing.assert_allclose(st['gp_loss'], g[pre + 'gp_loss'], rtol=1e-11)
    np.testing.assert_allclose(st['loss'], g[pre + 'loss'], rtol=1e-11)
    # fp32 restatement vs the reference run in fp32 (config.float_type default)
    st32 = make_oracle(g, cname, dtype=np.float32).forward(get_draws(g))
    np.testing.assert_allclose(st32['loss'], g['gpdbn/%s/float32/loss' % cname], rtol=2e-4)


@pytest.mark.parametrize('cname', list(CASES))
def test_gradient_matches_reference_finite_differences(gpdbn_golden, cname):
    g = gpdbn_golden
    _, grads = make_oracle(g, cname).grads(get_draws(g))
    flat = {'W0': grads['lower'][0]['W'], 'b0': grads['lower'][0]['b'], 'c0': grads['lower'][0]['c'],
            'W1': grads['lower'][1]['W'], 'b1': grads['lower'][1]['b'], 'c1': grads['lower'][1]['c'],
            'Wt': grads['top']['W'], 'bt': grads['top']['b'], 'ct': grads['top']['c'], 'X': grads['X']}
    for key, gr in flat.items():
        d = g['gpdbn/%s/dir/%s' % (cname, key)]
        fd = float(g['gpdbn/%s/fd/%s' % (cname, key)])
        an = float(np.sum(gr * d))
        assert abs(an - fd) <= 1e-5 * max(abs(fd), 1.0), (key, an, fd)
    if cname == 'trainable':
        for key in ('opt_alpha', 'opt_gamma', 'opt_noise'):
            fd = float(g['gpdbn/trainable/fd/' + key])
            assert abs(float(grads[key]) - fd) <= 1e-5 * max(abs(fd), 1.0), (key, float(grads[key]), fd)


def test_gradient_opt_alpha_h_by_own_finite_difference(gpdbn_golden):
    """opt_alpha_h is created inside the reference constructor (gprbm.py:79-83) and cannot be injected there; its
    gradient is checked against central differences of the oracle's own (reference-pinned) forward."""
    g = gpdbn_golden
    o = make_oracle(g, 'fixed')
    draws = get_draws(g)
    _, grads = o.grads(draws)
    rs = np.random.RandomState(3)
    d = rs.normal(size=o.top['opt_alpha_h'].shape)
    base = o.top['opt_alpha_h'].copy()
    eps = 1e-4
    o.top['opt_alpha_h'] = base + eps * d
    lp = o.loss(draws)
    o.top['opt_alpha_h'] = base - eps * d
    lm = o.loss(draws)
    fd = (lp - lm) / (2 * eps)
    an = float(np.sum(grads['top']['opt_alpha_h'] * d))
    assert abs(an - fd) <= 1e-5 * max(abs(fd), 1.0)
