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


@pytest.fixture(scope='module')
def latent_golden():
    return Golden('latent_golden.npz')


def latent_draws(lg, pre):
    return {'eps': lg[pre + 'draws/eps'], 'down_top': lg[pre + 'draws/down_top'],
            'down': [lg[pre + 'draws/down1'], lg[pre + 'draws/down0']]}


def test_held_out_loss_matches_reference(gpdbn_golden, latent_golden):
    """oracle.held_out_loss against the mse the reference's own build_sample_mean / downprop / dist.mse produce in eval
    mode (gprbm.py:440-453), fp64 and fp32."""
    g, lg = gpdbn_golden, latent_golden
    for dname, rtol in (('float64', 1e-11), ('float32', 2e-4)):
        o = make_oracle(g, 'trainable', dtype=np.dtype(dname).type)
        got = ogprbm.held_out_loss(o, lg['latent/H2'], lg['latent/H1_mean'], lg['heldout/x'], lg['heldout/test_data'],
                                   latent_draws(lg, 'heldout/'))
        np.testing.assert_allclose(got, lg['heldout/%s/mse' % dname], rtol=rtol)


def test_interp_pieces_match_reference(gpdbn_golden, latent_golden):
    """oracle.interp_images / interp_objective against create_observed_prediction_with_variance run from the reference
    source (gprbm.py:564-612)."""
    g, lg = gpdbn_golden, latent_golden
    o = make_oracle(g, 'trainable')
    Xi, S = lg['interp/Xi'], int(lg['interp/S'])
    imgs, pv = ogprbm.interp_images(o, lg['latent/H2'], lg['latent/H1_mean'], Xi, S, latent_draws(lg, 'interp/'))
    np.testing.assert_allclose(pv, lg['interp/float64/pred_var'], rtol=1e-10)
    np.testing.assert_allclose(imgs, lg['interp/float64/images'], rtol=1e-9, atol=1e-13)
    obj, pv2 = ogprbm.interp_objective(o, Xi[1:-1], Xi[0], Xi[-1], 0.1)
    np.testing.assert_allclose(pv2, lg['interp/float64/pred_var'][1:-1], rtol=1e-10)
    seg = np.diff(Xi, axis=0)
    want = np.mean(np.sum(seg ** 2, axis=1)) + 0.1 * np.mean(np.log(lg['interp/float64/pred_var'][1:-1]))
    np.testing.assert_allclose(obj, want, rtol=1e-10)
    # the linear initial path of gprbm.py:623-637
    path = ogprbm.interp_path(Xi[0], Xi[-1], 3)
    np.testing.assert_allclose(path, Xi[0] + np.linspace(0, 1, 5)[:, None] * (Xi[-1] - Xi[0]), rtol=1e-14)
