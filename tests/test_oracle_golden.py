"""Pins the CPU oracle (oracle/) against golden vectors produced by executing the reference's own
source files (tests/golden/make_golden.py) and against the fp64 anchors of SURVEY.md §8(c)."""
import numpy as np
import pytest

from oracle import rbm as orbm, gp as ogp, dist as odist

KINDS = ['rbm', 'bgrbm', 'gbrbm', 'ggrbm']
TOL = {'float32': dict(rtol=2e-5, atol=2e-6), 'float64': dict(rtol=1e-12, atol=1e-13)}


def _model(g, tag, kind, dname, temperature=None):
    dt = np.dtype(dname).type
    return orbm.RBMOracle(g[tag + '/W'], g[tag + '/b'], g[tag + '/c'], kind=kind, opt_sigma_h=g[tag + '/opt_sigma'],
                          temperature=temperature, dtype=dt)


@pytest.mark.parametrize('dname', ['float32', 'float64'])
@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('tag', ['toy', 'mid'])
def test_rbm_family_forward(rbm_golden, tag, kind, dname):
    g = rbm_golden
    m = _model(g, tag, kind, dname)
    v = g[tag + ('/v_real' if m.gaussian_visible else '/v_bin')]
    h = g[tag + ('/h_real' if m.gaussian_hidden else '/h_bin')]
    pre = '%s/%s/%s/' % (tag, kind, dname)
    tol = TOL[dname]
    np.testing.assert_allclose(m.h_net_input(v), g[pre + 'h_net_input'], **tol)
    np.testing.assert_allclose(m.v_net_input(h), g[pre + 'v_net_input'], **tol)
    hp, vp = m.h_probs(v), m.v_probs(h)
    np.testing.assert_allclose(hp, g[pre + 'h_probs'], **tol)
    np.testing.assert_allclose(vp, g[pre + 'v_probs'], **tol)
    # sampling is applied to the golden probabilities so a 1-ulp difference upstream cannot flip a bit
    sh = m.sample_h(g[pre + 'h_probs'], g[tag + ('/E_h' if m.gaussian_hidden else '/U_h')])
    sv = m.sample_v(g[pre + 'v_probs'], g[tag + '/U_v'])
    if m.gaussian_hidden:
        np.testing.assert_allclose(sh, g[pre + 'sample_h'], **tol)
    else:
        np.testing.assert_array_equal(sh, g[pre + 'sample_h'])
    if m.gaussian_visible:
        np.testing.assert_allclose(sv, g[pre + 'sample_v'], **tol)
    else:
        np.testing.assert_array_equal(sv, g[pre + 'sample_v'])
    np.testing.assert_allclose(m.free_energy(v), g[pre + 'free_energy'], rtol=tol['rtol'] * 5, atol=tol['atol'] * 20)
    vk = (0.7 * g[tag + '/v_real'][::-1] + 0.05) if m.gaussian_visible else g[tag + '/vk_bin']
    np.testing.assert_allclose(m.cd_loss(v, vk), g[pre + 'cd_loss'], rtol=1e-3 if dname == 'float32' else 1e-10,
                               atol=1e-4 if dname == 'float32' else 1e-12)


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('tag', ['toy', 'mid'])
def test_gibbs_chain_fp64(rbm_golden, tag, kind):
    g = rbm_golden
    m = _model(g, tag, kind, 'float64')
    h = g[tag + ('/h_real' if m.gaussian_hidden else '/h_bin')]
    k = int(g[tag + '/k'])
    draws = [(g[tag + '/chain_v'][s], (g[tag + '/chain_h_e'] if m.gaussian_hidden else g[tag + '/chain_h_u'])[s])
             for s in range(k)]
    vp, v, hp, hh = m.gibbs_sample(h, k, draws)
    pre = '%s/%s/float64/' % (tag, kind)
    np.testing.assert_allclose(vp, g[pre + 'gibbs_v_probs'], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(hp, g[pre + 'gibbs_h_probs'], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(v, g[pre + 'gibbs_v'], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(hh, g[pre + 'gibbs_h'], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize('kind', ['rbm', 'bgrbm', 'gbrbm'])
def test_concrete_units(rbm_golden, kind):
    g = rbm_golden
    for dname in ('float32', 'float64'):
        m = _model(g, 'mid', kind, dname, temperature=0.1)
        pre = 'mid/%s/%s/' % (kind, dname)
        tol = dict(rtol=5e-4, atol=1e-5) if dname == 'float32' else dict(rtol=1e-10, atol=1e-12)
        if not m.gaussian_hidden:
            np.testing.assert_allclose(m.sample_h(g[pre + 'h_probs'], g['mid/U_h']), g[pre + 'concrete_h'], **tol)
        if not m.gaussian_visible:
            np.testing.assert_allclose(m.sample_v(g[pre + 'v_probs'], g['mid/U_v']), g[pre + 'concrete_v'], **tol)


@pytest.mark.parametrize('kind', KINDS)
def test_cd_gradient_closed_form_vs_reference_finite_differences(rbm_golden, kind):
    """The closed forms against central differences of the REFERENCE's cd_loss."""
    g = rbm_golden
    m = _model(g, 'toy', kind, 'float64')
    v = g['toy/v_real'] if m.gaussian_visible else g['toy/v_bin']
    vk = (0.7 * g['toy/v_real'][::-1] + 0.05) if m.gaussian_visible else g['toy/vk_bin']
    gW, gb, gc, gs = m.cd_grad(v, vk)
    np.testing.assert_allclose(gW, g['toy/%s/fd_grad_W' % kind], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(gb.reshape(-1), g['toy/%s/fd_grad_b' % kind], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(gc.reshape(-1), g['toy/%s/fd_grad_c' % kind], rtol=1e-6, atol=1e-8)
    if m.gaussian_hidden:
        np.testing.assert_allclose(gs.reshape(-1), g['toy/%s/fd_grad_opt_sigma' % kind], rtol=1e-6, atol=1e-8)


def test_stack_upprop_downprop(rbm_golden):
    g = rbm_golden
    l1 = orbm.RBMOracle(g['stack/W1'], g['stack/b1'], g['stack/c1'])
    l2 = orbm.RBMOracle(g['stack/W2'], g['stack/b2'], g['stack/c2'])
    h1 = l1.sample_h(l1.h_probs(g['stack/v']), g['stack/u1'])
    up = l2.sample_h(l2.h_probs(h1), g['stack/u2'])
    np.testing.assert_array_equal(up, g['stack/up'])
    v1 = l2.sample_v(l2.v_probs(up), g['stack/u3'])
    down = l1.sample_v(l1.v_probs(v1), g['stack/u4'])
    np.testing.assert_array_equal(down, g['stack/down'])


@pytest.mark.parametrize('dname', ['float32', 'float64'])
def test_se_kernel(gp_golden, dname):
    g = gp_golden
    dt = np.dtype(dname).type
    tol = dict(rtol=1e-5, atol=1e-6) if dname == 'float32' else dict(rtol=1e-12, atol=1e-14)
    for tag in ('k83', 'k30'):
        X, Y = g[tag + '/X'].astype(dt), g[tag + '/Y'].astype(dt)
        for an, (a, ga) in {'unit': (1.0, 1.0), 'ag': (1.7, 0.6)}.items():
            pre = '%s/%s/%s/' % (tag, dname, an)
            np.testing.assert_allclose(ogp.se_kernel(X, None, dt(a), dt(ga)), g[pre + 'K_X'], **tol)
            np.testing.assert_allclose(ogp.se_kernel(X, Y, dt(a), dt(ga)), g[pre + 'K_XY'], **tol)
            np.testing.assert_allclose(ogp.se_kernel(X, None, dt(a), dt(ga), diag=True), g[pre + 'K_X_diag'], **tol)
            if (pre + 'K_XY_diag') in g:
                np.testing.assert_allclose(ogp.se_kernel(X, Y, dt(a), dt(ga), diag=True), g[pre + 'K_XY_diag'], **tol)


def _gplvm_oracle(g, cname, dname='float64'):
    dt = np.dtype(dname).type
    oa, og, on = g['gplvm/%s/opt' % cname]
    trainable = (True, True, True) if cname == 'trainable' else (False, False, False)
    return ogp.GPLVMOracle(g['gplvm/%s/%s/Y_centered' % (cname, dname)], g['gplvm/%s/%s/X0' % (cname, dname)],
                           opt_alpha=oa, opt_gamma=og, opt_noise=on, fixed_noise=0.01, jitter=1e-6,
                           trainable=trainable, dtype=dt)


@pytest.mark.parametrize('cname', ['trainable', 'gpdbn'])
def test_gplvm_loss_and_prediction(gp_golden, cname):
    g = gp_golden
    m = _gplvm_oracle(g, cname)
    pre = 'gplvm/%s/float64/' % cname
    loss, logdet, ssq, L, S = m.loss()
    np.testing.assert_allclose(m.K(), g[pre + 'K'], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(L, g[pre + 'L'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(loss, g[pre + 'loss'], rtol=1e-12)
    np.testing.assert_allclose(logdet, g[pre + 'log_det'], rtol=1e-11)
    mean, var = m.predict(g['gplvm/xstar'], g[pre + 'Y_mean'])
    np.testing.assert_allclose(mean, g[pre + 'pred_mean'], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, g[pre + 'pred_var'], rtol=1e-8, atol=1e-11)
    # fp32 restatement stays within fp32 distance of the reference's fp32 run
    m32 = _gplvm_oracle(g, cname, 'float32')
    np.testing.assert_allclose(m32.loss()[0], g['gplvm/%s/float32/loss' % cname], rtol=5e-5)


@pytest.mark.parametrize('cname', ['trainable', 'gpdbn'])
def test_gplvm_gradient_closed_form_vs_reference_finite_differences(gp_golden, cname):
    g = gp_golden
    m = _gplvm_oracle(g, cname)
    dX, da, dg, dn = m.grad()
    np.testing.assert_allclose(dX, g['gplvm/%s/fd_dX' % cname], rtol=2e-5, atol=1e-6)
    if cname == 'trainable':
        np.testing.assert_allclose([da, dg, dn], g['gplvm/%s/fd_dopt' % cname], rtol=2e-5, atol=1e-6)


def test_pca_standardize_init_matches_reference(gp_golden, misc_golden):
    A = misc_golden['pre/A']
    p = ogp.pca(A, 2)
    ref = misc_golden['pre/pca2']
    for col in range(2):   # eigenvector signs are arbitrary
        s = np.sign(np.dot(p[:, col], ref[:, col]))
        np.testing.assert_allclose(s * p[:, col], ref[:, col], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(ogp.standardize(A, np.float64), misc_golden['pre/std'], rtol=1e-12)
    np.testing.assert_allclose(ogp.inverse_softplus(np.array([0.3, 1.0, 2.5])), misc_golden['pre/inv_softplus'], rtol=1e-14)


def test_dist(misc_golden):
    p, t = misc_golden['dist/p'], misc_golden['dist/t']
    np.testing.assert_allclose(odist.mse(p, t), misc_golden['dist/mse'], rtol=1e-13)
    np.testing.assert_allclose(odist.cross_entropy(p, t), misc_golden['dist/ce_mean'], rtol=1e-13)
    np.testing.assert_allclose(odist.cross_entropy(p, t, reduce_mean=False), misc_golden['dist/ce_sum'], rtol=1e-13)


def test_survey_anchors_on_horses(horses, gp_golden):
    """SURVEY.md §8(c) known answers (fp64) on the bundled dataset."""
    v = horses.astype(np.float64)
    V, H = 1024, 500
    zero = orbm.RBMOracle(np.zeros((V, H)), np.zeros(V), np.zeros(H))
    np.testing.assert_allclose(zero.free_energy(v), -H * np.log(2.0), rtol=1e-14)          # anchor (1)
    assert np.all(zero.h_probs(v) == 0.5) and zero.cd_loss(v, 1 - v) == 0.0
    i, j = np.arange(V)[:, None], np.arange(H)[None, :]
    m = orbm.RBMOracle(0.01 * np.cos(0.37 * i + 0.11 * j), 0.05 * np.sin(0.3 * np.arange(V)), 0.02 * np.cos(0.7 * np.arange(H)))
    F = m.free_energy(v)                                                                    # anchor (2)
    np.testing.assert_allclose(np.mean(F), -347.28000091799646, rtol=1e-12)
    np.testing.assert_allclose(F[0, 0], -348.6856458065582, rtol=1e-12)
    np.testing.assert_allclose(np.sum(m.h_probs(v)), 81967.46732612466, rtol=1e-12)
    gW = m.cd_grad(v, 1 - v)[0]
    np.testing.assert_allclose(np.linalg.norm(gW), 240.9141144337796, rtol=1e-10)
    np.testing.assert_allclose(gW[0, 0], 0.45235972577049083, rtol=1e-10)
    np.testing.assert_allclose(gW[512, 250], 0.48199651515066666, rtol=1e-10)
    bg = orbm.RBMOracle(m.W, m.b, m.c, kind='bgrbm')
    np.testing.assert_allclose(np.mean(bg.free_energy(v)), -3.632522481346598, rtol=1e-11)
    # anchor (3): GPLVM at construction; the reference-code golden and the surveyor's scalar agree to ~1e-9
    X0 = gp_golden['horses/float64/X0']
    Yc = v - v.mean(0, keepdims=True)
    gp = ogp.GPLVMOracle(Yc, X0)
    loss, logdet, ssq, _, _ = gp.loss()
    np.testing.assert_allclose(loss, gp_golden['horses/float64/loss'], rtol=1e-12)
    np.testing.assert_allclose(loss, 342948.9536974635, rtol=1e-8)
    np.testing.assert_allclose(logdet, 36.850482771829924, rtol=1e-7)
    dX, da, dg, dn = gp.grad()
    np.testing.assert_allclose(np.linalg.norm(dX), 891.6048745762016, rtol=1e-5)
    np.testing.assert_allclose([da, dg, dn], [4347.456017283226, -9954.42826462195, 92050.97401113895], rtol=1e-5)


def test_latent_search_loss_matches_reference(gp_golden):
    """oracle.gp.testgen_loss against the loss the reference's own graph pieces produce (gplvm.py:478-483)."""
    g = gp_golden
    pre = 'gplvm/trainable/float64/'
    oa, og_, on = g['gplvm/trainable/opt']
    m = ogp.GPLVMOracle(g[pre + 'Y_centered'], g[pre + 'X0'], opt_alpha=oa, opt_gamma=og_, opt_noise=on)
    got = ogp.testgen_loss(m, g['gplvm/xstar'], g['testgen/target'], 0.1, y_mean=g[pre + 'Y_mean'])
    np.testing.assert_allclose(got, g['testgen/loss'], rtol=1e-12)


def test_ard_kernel_and_gplvm_match_reference():
    """SEKernel(ARD=True) (gplvm.py:24-28) through the oracle against the reference's own gplvm.py: K, L, loss,
    prediction, and the closed-form gradient against central differences of the reference loss."""
    from conftest import Golden
    g = Golden("ard_golden.npz")
    opt = g['ard/opt']
    Q = g['ard/X0'].shape[1]
    m = ogp.GPLVMOracle(g['ard/Y'] - g['ard/Y'].mean(axis=0, keepdims=True), g['ard/X0'], opt_alpha=opt[0],
                        opt_gamma=opt[1:1 + Q], opt_noise=opt[1 + Q])
    np.testing.assert_allclose(m.gamma, g['ard/gamma'], rtol=1e-12)
    np.testing.assert_allclose(ogp.se_kernel(m.X, None, m.alpha, m.gamma), g['ard/float64/K_X'], rtol=1e-12)
    np.testing.assert_allclose(ogp.se_kernel(m.X, g['ard/xstar'], m.alpha, m.gamma), g['ard/float64/K_Xx'], rtol=1e-12)
    np.testing.assert_allclose(m.K(), g['ard/float64/K'], rtol=1e-12)
    loss, logdet, _, L, _ = m.loss()
    np.testing.assert_allclose(L, g['ard/float64/L'], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(loss, g['ard/float64/loss'], rtol=1e-12)
    np.testing.assert_allclose(logdet, g['ard/float64/log_det'], rtol=1e-12)
    mean, var = m.predict(g['ard/xstar'], y_mean=g['ard/Y'].mean(axis=0, keepdims=True))
    np.testing.assert_allclose(mean, g['ard/float64/pred_mean'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(var, g['ard/float64/pred_var'], rtol=1e-9)
    dX, da, dg, dn = m.grad()
    assert np.max(np.abs(dX - g['ard/fd_dX'])) / np.max(np.abs(g['ard/fd_dX'])) < 1e-6
    np.testing.assert_allclose(dg, g['ard/fd_dopt_gamma'], rtol=1e-6)
    np.testing.assert_allclose(da, g['ard/fd_dopt_alpha'], rtol=1e-6)
    np.testing.assert_allclose(dn, g['ard/fd_dopt_noise'], rtol=1e-6)
