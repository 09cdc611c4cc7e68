"""GPU parity tests of the GP path (Gram, Cholesky, triangular solve, LML + analytic gradient, prediction,
GPLVM.optimize) against the fp64 CPU oracle and the goldens produced by the reference's own code.

Tolerance: losses / log-likelihoods 1e-4 relative; gradient arrays 1e-4 norm-relative, or
the error of the fp32 restatement of the reference formula where that is larger (ill-conditioned K at
sigma^2 = 0.01, SURVEY.md §7 hard part 3)."""
import numpy as np
import pytest
import scipy.linalg
import torch

from conftest import has_cuda
from oracle import gp as ogp, rbm as orbm

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')]

LOSS_RTOL = 1e-4
GRAD_NORMREL = 1e-4


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def _cu(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _normrel(a, b):
    return np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)


def test_se_kernel_reference_unit_tests_and_goldens(gp_golden):
    """tests/test_gplvm.py of the reference (K(X)==K(X,X), diag variants) + values from the reference's code."""
    from deepbelief_b200.layers.gplvm import SEKernel
    g = gp_golden
    for an, (a, ga) in {'unit': (1.0, 1.0), 'ag': (1.7, 0.6)}.items():
        kern = SEKernel(alpha=a, gamma=ga)
        for tag in ('k83', 'k30'):
            X, Y = g[tag + '/X'], g[tag + '/Y']
            pre = '%s/float64/%s/' % (tag, an)
            kx, kxx, kxy = _np(kern.K(X)), _np(kern.K(X, X)), _np(kern.K(X, Y))
            np.testing.assert_allclose(kx, kxx, rtol=1e-6)
            np.testing.assert_allclose(kx, g[pre + 'K_X'], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(kxy, g[pre + 'K_XY'], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(_np(kern.K(X, diag=True)), g[pre + 'K_X_diag'], rtol=1e-6)
            if X.shape == Y.shape:
                np.testing.assert_allclose(_np(kern.K(X, Y, diag=True)), np.diag(kxy), rtol=1e-6)
                np.testing.assert_allclose(_np(kern.K(X, Y, diag=True)), g[pre + 'K_XY_diag'], rtol=1e-5, atol=1e-6)
    assert abs(float(kern.alpha) - 1.7) < 1e-6 and abs(float(kern.gamma) - 0.6) < 1e-6
    fixed = SEKernel(alpha=False, gamma=False)
    assert float(fixed.alpha) == 1.0 and float(fixed.gamma) == 1.0       # gplvm.py:31-32,41-42


@pytest.mark.parametrize('N', [5, 64, 130, 700, 1101])
def test_cholesky_and_triangular_solves(N):
    from deepbelief_b200 import ops
    rs = np.random.RandomState(N)
    X = rs.normal(size=(N, 2))
    K = ogp.se_kernel(X, None, 1.3, 0.9) + 0.3 * np.eye(N)
    A = _cu(K)
    info = ops.potrf_lower(A)
    assert int(info.item()) == 0
    Lref = np.linalg.cholesky(K)
    assert _normrel(np.tril(_np(A)), Lref) < 1e-5          # fp32 factor of a cond(K) ~ 3e3 matrix (bound ~ cond * eps)
    assert np.allclose(np.triu(_np(A), 1), np.triu(K, 1), rtol=1e-6)       # strict upper triangle untouched
    for nrhs in (1, 37, 300):
        B = rs.normal(size=(N, nrhs))
        fwd = _np(ops.trsm_lower(A, _cu(B), trans=False))
        bwd = _np(ops.trsm_lower(A, _cu(B), trans=True))
        assert _normrel(fwd, scipy.linalg.solve_triangular(Lref, B, lower=True)) < 5e-5
        assert _normrel(bwd, scipy.linalg.solve_triangular(Lref, B, lower=True, trans='T')) < 5e-5


def test_cholesky_reports_non_positive_definite_input():
    from deepbelief_b200 import ops
    N = 100
    rs = np.random.RandomState(0)
    M = rs.normal(size=(N, N))
    K = M @ M.T + N * np.eye(N)
    K[70, 70] = -5.0                          # first failing pivot at (1-based) 71
    info = ops.potrf_lower(_cu(K))
    assert int(info.item()) == 71


def _gplvm_pair(g, cname):
    from deepbelief_b200.layers.gplvm import GPLVM, SEKernel
    trainable = cname == 'trainable'
    kw = dict(noise_variance=0.7, fixed_noise_variance=False, alpha=1.3, gamma=0.8) if trainable else \
        dict(noise_variance=0.01, fixed_noise_variance=True, alpha=False, gamma=False)
    pre = 'gplvm/%s/float64/' % cname
    Yraw = g['gplvm/Y']
    kern = SEKernel(alpha=kw['alpha'], gamma=kw['gamma'])
    lay = GPLVM(Yraw.astype(np.float32), 2, X0=g[pre + 'X0'].astype(np.float32), kern=kern,
                noise_variance=kw['noise_variance'], fixed_noise_variance=kw['fixed_noise_variance'])
    lay.build_model()
    oa, og, on = g['gplvm/%s/opt' % cname]
    orc = ogp.GPLVMOracle(g[pre + 'Y_centered'], g[pre + 'X0'], opt_alpha=oa, opt_gamma=og, opt_noise=on, fixed_noise=0.01,
                          jitter=1e-6, trainable=(trainable,) * 3)
    return lay, orc, pre


@pytest.mark.parametrize('cname', ['trainable', 'gpdbn'])
def test_gplvm_loss_gradient_prediction_vs_oracle_and_reference_goldens(gp_golden, cname):
    g = gp_golden
    lay, orc, pre = _gplvm_pair(g, cname)
    from deepbelief_b200 import compat
    sess = compat.Session()
    np.testing.assert_allclose(_np(lay.Y), g[pre + 'Y_centered'], atol=1e-6)
    np.testing.assert_allclose(sess.run(lay.K), g[pre + 'K'], rtol=1e-5, atol=1e-6)
    assert _normrel(sess.run(lay.L), g[pre + 'L']) < 2e-5
    np.testing.assert_allclose(sess.run(lay.loss), g[pre + 'loss'], rtol=LOSS_RTOL)
    np.testing.assert_allclose(sess.run(lay.log_det), g[pre + 'log_det'], rtol=LOSS_RTOL, atol=1e-4)
    loss, dX, dh = lay.loss_and_grad()
    w_dX, w_da, w_dg, w_dn = orc.grad()
    # error budget: 1e-4, or the error of the fp32 restatement of the reference formula if that is larger
    o32 = ogp.GPLVMOracle(orc.Y, orc.X, opt_alpha=orc.opt_alpha, opt_gamma=orc.opt_gamma, opt_noise=orc.opt_noise,
                          fixed_noise=0.01, jitter=1e-6, trainable=orc.trainable, dtype=np.float32)
    budget = max(GRAD_NORMREL, 2 * _normrel(o32.grad()[0].astype(np.float64), w_dX))
    assert _normrel(_np(dX), w_dX) < budget
    assert _normrel(_np(dX), g['gplvm/%s/fd_dX' % cname]) < budget + 1e-4          # reference-code finite differences
    if cname == 'trainable':
        np.testing.assert_allclose(_np(dh), [w_da, w_dg, w_dn], rtol=GRAD_NORMREL * 3)
        np.testing.assert_allclose(_np(dh), g['gplvm/trainable/fd_dopt'], rtol=5e-4)
    else:
        assert np.all(_np(dh) == 0)
    xs = g['gplvm/xstar']
    mean = lay.predict_mean(xs)
    var = lay.predict_variance(xs)
    assert var.shape == (xs.shape[0], 1)
    m_budget = 1e-4 if cname == 'trainable' else 2e-3          # sigma^2 = 0.01: cond(K) ~ 1e4 in fp32
    assert _normrel(mean, g[pre + 'pred_mean']) < m_budget
    np.testing.assert_allclose(var[:, 0], g[pre + 'pred_var'], rtol=m_budget * 10, atol=m_budget)


def test_gplvm_on_horses_matches_anchors(horses, gp_golden):
    """GPLVM at construction on the bundled horses data (SURVEY §8c anchor (3); golden from the reference's code)."""
    from deepbelief_b200.layers.gplvm import GPLVM
    lay = GPLVM(horses, 2, noise_variance=1.0)
    assert (lay.N, lay.D, lay.Q) == (328, 1024, 2)
    X0 = _np(lay.X)
    ref_X0 = gp_golden['horses/float64/X0']
    for q in range(2):                    # PCA eigenvector signs are arbitrary
        s = np.sign(np.dot(X0[:, q], ref_X0[:, q]))
        np.testing.assert_allclose(s * X0[:, q], ref_X0[:, q], atol=2e-3)
    loss, dX, dh = lay.loss_and_grad()
    np.testing.assert_allclose(float(loss), float(gp_golden['horses/float64/loss']), rtol=LOSS_RTOL)
    np.testing.assert_allclose(float(lay.build_log_det()), 36.850482771829924, rtol=LOSS_RTOL)
    np.testing.assert_allclose(np.linalg.norm(_np(dX)), 891.6048745762016, rtol=2e-4)
    np.testing.assert_allclose(_np(dh), [4347.456017283226, -9954.42826462195, 92050.97401113895], rtol=2e-4)


def test_gplvm_optimize_follows_oracle_adam(gp_golden):
    from deepbelief_b200 import compat as tf
    g = gp_golden
    lay, orc, pre = _gplvm_pair(g, 'trainable')
    iters = 5
    lay.optimize(tf.train.AdamOptimizer(learning_rate=1e-2), num_iterations=iters)
    theta = np.concatenate([orc.X.ravel(), [orc.opt_alpha, orc.opt_gamma, orc.opt_noise]])
    adam = orbm.Adam(1e-2)
    NQ = orc.N * orc.Q
    for _ in range(iters):
        o = ogp.GPLVMOracle(orc.Y, theta[:NQ].reshape(orc.N, orc.Q), opt_alpha=theta[NQ], opt_gamma=theta[NQ + 1],
                            opt_noise=theta[NQ + 2], jitter=1e-6)
        dX, da, dg, dn = o.grad()
        theta = adam.step(theta, np.concatenate([dX.ravel(), [da, dg, dn]]))
    got = _np(lay._params)
    assert np.max(np.abs(got - theta)) < 2e-4             # 5 steps of size <= 1e-2
    final = ogp.GPLVMOracle(orc.Y, theta[:NQ].reshape(orc.N, orc.Q), opt_alpha=theta[NQ], opt_gamma=theta[NQ + 1],
                            opt_noise=theta[NQ + 2], jitter=1e-6).loss()[0]
    np.testing.assert_allclose(float(lay.build_loss()), final, rtol=LOSS_RTOL)
    assert final < orc.loss()[0]                          # and the loss went down


def test_non_pd_and_negative_variance_are_reported():
    from deepbelief_b200 import ops
    from deepbelief_b200.layers.gplvm import GPLVM
    rs = np.random.RandomState(0)
    Y = rs.normal(size=(40, 6)).astype(np.float32)
    X0 = np.zeros((40, 2), dtype=np.float32)              # all latents equal + zero noise -> singular K
    lay = GPLVM(Y, 2, X0=X0, noise_variance=0.0, fixed_noise_variance=True)
    with pytest.raises(RuntimeError):
        lay.build_model()
        lay.predict_mean(np.zeros((1, 2)))
    with pytest.raises(AssertionError):
        GPLVM([[1.0]], 2)                                  # gplvm.py:110


@pytest.mark.parametrize('N,D', [(1500, 96)])
def test_mid_size_against_oracle(N, D):
    """A size where the fp64 oracle still finishes in seconds."""
    from deepbelief_b200 import ops
    rs = np.random.RandomState(1)
    X = rs.normal(size=(N, 2))
    Y = rs.normal(size=(N, D)); Y -= Y.mean(0)
    orc = ogp.GPLVMOracle(Y, X, jitter=1e-6)
    plan = ops.GplvmPlan(N, 2, D)
    hyp = _cu([orc.opt_alpha, orc.opt_gamma, orc.opt_noise])
    scal, dX, info = plan.evaluate(_cu(X), _cu(Y), hyp, 7, jitter=1e-6)
    assert int(info.item()) == 0
    loss, logdet, ssq, _, _ = orc.loss()
    np.testing.assert_allclose(float(scal[0]), loss, rtol=LOSS_RTOL)
    w_dX, w_da, w_dg, w_dn = orc.grad()
    assert _normrel(_np(dX), w_dX) < GRAD_NORMREL
    np.testing.assert_allclose(_np(scal[3:6]), [w_da, w_dg, w_dn], rtol=3e-4)


def test_full_size_n5000_properties():
    """BASELINE 'GPLVM at N=5000' (D=784, Q=2) through size-independent properties: L L^T reproduces K, the
    triangular solve inverts L, loss pieces are consistent with the factor (the gradient and the loss against the fp64
    oracle: test_full_size_n5000_vs_fp64_oracle)."""
    from deepbelief_b200 import ops
    N, D, Q = 5000, 784, 2
    gen = torch.Generator(device='cuda').manual_seed(0)
    X = torch.randn(N, Q, device='cuda', generator=gen)
    Y = torch.randn(N, D, device='cuda', generator=gen)
    Y = Y - Y.mean(0, keepdim=True)
    hyp_opt = _cu([0.5413, 0.5413, 0.5413])
    plan = ops.GplvmPlan(N, Q, D)
    Lf = torch.zeros(N, N, device='cuda')
    scal, dX, info = plan.evaluate(X, Y, hyp_opt, 7, jitter=1e-6, L_out=Lf)
    assert int(info.item()) == 0
    hyp = ops.gp_effective_hyp(hyp_opt, 7, 0.0, 1e-6)
    K = ops.gram_rbf(X, None, hyp, add_noise=True)
    assert torch.equal(K, K.t())                                              # symmetric, exact
    Lt = torch.tril(Lf).double()
    resid = (Lt @ Lt.t() - K.double()).abs().max() / K.double().abs().max()   # torch fp64 only as the checker
    assert float(resid) < 5e-6            # fp32 factor; the wide trailing updates run as 3xTF32 tensor-core GEMMs
    S = ops.trsm_lower(Lf, Y.clone())
    assert float((Lt @ S.double() - Y.double()).abs().max()) < 5e-4
    np.testing.assert_allclose(float(scal[2]), float((S.double() ** 2).sum()), rtol=1e-5)
    np.testing.assert_allclose(float(scal[1]), float(2 * torch.log(torch.diagonal(Lt)).sum()), rtol=1e-5)


def test_full_size_n5000_vs_fp64_oracle():
    """BASELINE metric 2 at its own size (N=5000, D=784, Q=2; binary Bernoulli(0.13) targets as in bench.py): loss, log
    determinant, ||S||^2 and the FULL gradient (dX and the three hyper-parameter derivatives) against the fp64 oracle
    (oracle/gp.py restating gplvm.py:210-244 and the autodiff of gplvm.py:425-428), at the stated 1e-4. The budget per
    quantity is max(1e-4, error of the reference's own formulas evaluated in fp32) (SURVEY.md section 7.3 rule)."""
    from deepbelief_b200 import ops
    N, D, Q = 5000, 784, 2
    rs = np.random.RandomState(5000)
    X = rs.normal(size=(N, Q)).astype(np.float32)
    Y = (rs.uniform(size=(N, D)) < 0.13).astype(np.float32)
    Y -= Y.mean(axis=0, keepdims=True)
    orc = ogp.GPLVMOracle(Y, X, jitter=1e-6)              # alpha = gamma = 1, noise = 1 + 1e-6, all trainable
    loss, logdet, ssq, _, _ = orc.loss()
    w_dX, w_da, w_dg, w_dn = orc.grad()
    o32 = ogp.GPLVMOracle(Y, X, jitter=1e-6, dtype=np.float32)
    l32 = o32.loss()
    g32 = o32.grad()
    plan = ops.GplvmPlan(N, Q, D)
    hyp = _cu([orc.opt_alpha, orc.opt_gamma, orc.opt_noise])
    for rep in range(3):                                  # eager launch, graph capture, graph replay: all three must agree
        scal, dX, info = plan.evaluate(_cu(X), _cu(Y), hyp, 7, jitter=1e-6)
        assert int(info.item()) == 0
        budget = lambda want, ref32: max(1e-4, 2.0 * abs(float(ref32) - float(want)) / abs(float(want)))  # noqa: E731
        rel = lambda got, want: abs(float(got) - float(want)) / abs(float(want))  # noqa: E731
        err_dX = _normrel(_np(dX), w_dX)
        errs = {'loss': (rel(scal[0], loss), budget(loss, l32[0])), 'logdet': (rel(scal[1], logdet), budget(logdet, l32[1])),
                'ssq': (rel(scal[2], ssq), budget(ssq, l32[2])),
                'dX': (err_dX, max(1e-4, 2.0 * _normrel(g32[0].astype(np.float64), w_dX)))}
        for name, got, want, r32 in zip(('d_opt_alpha', 'd_opt_gamma', 'd_opt_noise'), _np(scal[3:6]), (w_da, w_dg, w_dn), g32[1:]):
            errs[name] = (rel(got, want), budget(want, r32))
        print('N=5000 parity (pass %d): ' % rep + ', '.join('%s %.1e (budget %.1e)' % (k, e, b) for k, (e, b) in errs.items()))
        for name, (e, b) in errs.items():
            assert e < b, (name, e, b)


def test_latent_search_loss_and_gradient(gp_golden):
    """GPLVM.test_generalisation's per-restart loss (gplvm.py:478-483) against the reference's own evaluation (golden),
    its gradient wrt the test latents against central differences of the fp64 oracle, and the driver loop."""
    import pickle
    from deepbelief_b200 import compat as tf
    g = gp_golden
    lay, orc, pre = _gplvm_pair(g, 'trainable')
    xs, target = g['gplvm/xstar'], g['testgen/target']
    loss, g_x, info, pm_norm, pm_clip = lay.latent_search_loss(xs, target, 0.1, want_outputs=True)
    assert int(info.item()) == 0
    # the normalised prediction contains an exact 1 (its arg-max), clipped to 1 - 1e-6: fp32 resolves that to 1 - 1.013e-6,
    # which moves log(p/(1-p)) by 0.0132 for that one element (the reference's fp32 run does the same)
    np.testing.assert_allclose(_np(loss), g['testgen/loss'], rtol=1e-4, atol=0.015 / lay.D)
    assert _normrel(_np(pm_norm), g['testgen/pm_norm']) < 1e-4
    assert _normrel(_np(pm_clip), g['testgen/pm_clip']) < 1e-4
    y_mean = g[pre + 'Y_mean']
    eps = 1e-5
    fd = np.zeros_like(xs)
    for q in range(xs.shape[1]):
        d = np.zeros_like(xs); d[:, q] = eps
        fd[:, q] = (ogp.testgen_loss(orc, xs + d, target, 0.1, y_mean) - ogp.testgen_loss(orc, xs - d, target, 0.1, y_mean)) / (2 * eps)
    assert _normrel(_np(g_x), fd) < 2e-4
    # the search itself: SGD from every training latent, best restart kept; the distance can only improve on the start
    start = _np(lay.latent_search_loss(_np(lay.X), target, 0.1)[0]).min()
    import tempfile, os
    path = os.path.join(tempfile.mkdtemp(), 'table.pkl')
    table = lay.test_generalisation(np.vstack([target, 1 - target]), path, tf.train.GradientDescentOptimizer(learning_rate=0.01),
                                    log_var_scaling=0.1, num_iterations=25, num_runs=lay.N)
    assert len(table) == 2 and table[0][0] == 0 and table[0][2].shape == (1, lay.Q) and table[0][4].shape == (1, lay.D)
    assert table[0][1] <= start + 1e-6
    with open(path, 'rb') as f:
        assert len(pickle.load(f)) == 2


@pytest.mark.parametrize('N,D', [(1024, 64), (1540, 128)])
def test_tensor_core_gp_path_vs_oracle(N, D):
    """N >= 512 with N, D multiples of 4 takes the 3xTF32 tcgen05 path (Cholesky trailing updates, K^-1, K^-1 Y,
    A A^T): loss, log-determinant and every gradient against the fp64 oracle."""
    from deepbelief_b200 import ops
    rs = np.random.RandomState(N)
    X = rs.normal(size=(N, 2)) * 1.5
    Y = rs.normal(size=(N, D))
    Y -= Y.mean(axis=0, keepdims=True)
    oa, og_, on = np.log(np.expm1(1.3)), np.log(np.expm1(0.8)), np.log(np.expm1(0.7))
    orc = ogp.GPLVMOracle(Y, X, opt_alpha=oa, opt_gamma=og_, opt_noise=on)
    plan = ops.GplvmPlan(N, 2, D)
    Lf = torch.zeros(N, N, device='cuda')
    scal, dX, info = plan.evaluate(_cu(X), _cu(Y), _cu([oa, og_, on]), 7, jitter=1e-6, L_out=Lf)
    assert int(info.item()) == 0
    loss, logdet, ssq, Lref, _ = orc.loss()
    w_dX, w_da, w_dg, w_dn = orc.grad()
    np.testing.assert_allclose(float(scal[0]), float(loss), rtol=LOSS_RTOL)
    np.testing.assert_allclose(float(scal[1]), float(logdet), rtol=LOSS_RTOL)
    np.testing.assert_allclose(float(scal[2]), float(ssq), rtol=LOSS_RTOL)
    assert _normrel(np.tril(_np(Lf)), Lref) < 1e-5
    assert _normrel(_np(dX), w_dX) < GRAD_NORMREL
    np.testing.assert_allclose(_np(scal[3:6]), [w_da, w_dg, w_dn], rtol=3e-4)


def test_ard_kernel_and_gplvm_vs_reference(tmp_path):
    """SEKernel(ARD=True) (gplvm.py:24-28, one length-scale per latent dimension): K, loss, prediction against the
    values the reference's own gplvm.py produces (tests/golden/ard_golden.npz); gradients with respect to X and every
    optimisation variable against central differences of the reference loss; optimize() moves all of them; checkpoint
    round trip of the vector length-scales."""
    from conftest import Golden
    from deepbelief_b200 import compat as tf
    from deepbelief_b200.layers.gplvm import GPLVM, SEKernel
    g = Golden('ard_golden.npz')
    Q = g['ard/X0'].shape[1]
    kern = SEKernel(alpha=float(g['ard/alpha']), gamma=g['ard/gamma'], ARD=True, Q=Q)
    np.testing.assert_allclose(_np(kern.gamma), g['ard/gamma'], rtol=1e-6)
    X0, xs = g['ard/X0'], g['ard/xstar']
    np.testing.assert_allclose(_np(kern.K(X0)), g['ard/float64/K_X'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_np(kern.K(X0, xs)), g['ard/float64/K_Xx'], rtol=1e-5, atol=1e-6)
    lay = GPLVM(g['ard/Y'].astype(np.float32), Q, X0=X0, kern=kern, noise_variance=float(g['ard/noise']), name='gplvm_ard')
    lay.build_model()
    np.testing.assert_allclose(_np(lay.K.eval()), g['ard/float64/K'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(lay.loss.eval()), g['ard/float64/loss'], rtol=LOSS_RTOL)
    np.testing.assert_allclose(float(lay.log_det.eval()), g['ard/float64/log_det'], rtol=LOSS_RTOL)
    assert _normrel(_np(lay.build_pred_mean(xs)), g['ard/float64/pred_mean']) < 1e-4
    np.testing.assert_allclose(_np(lay.build_pred_var(xs)), g['ard/float64/pred_var'], rtol=1e-3, atol=1e-6)
    loss, dX, dopt = lay.loss_and_grad()
    dopt = _np(dopt)
    assert dopt.shape == (Q + 2,)
    assert _normrel(_np(dX), g['ard/fd_dX']) < GRAD_NORMREL
    np.testing.assert_allclose(dopt[0], g['ard/fd_dopt_alpha'], rtol=1e-4)
    np.testing.assert_allclose(dopt[1:1 + Q], g['ard/fd_dopt_gamma'], rtol=1e-4, atol=1e-4 * np.max(np.abs(g['ard/fd_dopt_gamma'])))
    np.testing.assert_allclose(dopt[1 + Q], g['ard/fd_dopt_noise'], rtol=1e-4)
    # latent search (gplvm.py:478-483) in the unscaled coordinates: loss and x-gradient against the fp64 oracle
    opt = g['ard/opt']
    y_mean = g['ard/Y'].mean(axis=0, keepdims=True)
    orc = ogp.GPLVMOracle(g['ard/Y'] - y_mean, X0, opt_alpha=opt[0], opt_gamma=opt[1:1 + Q], opt_noise=opt[1 + Q])
    target = (np.random.RandomState(3).uniform(size=(1, lay.D)) > 0.5).astype(np.float64)
    x = xs[:3]
    ls, g_x, _, _, _ = lay.latent_search_loss(x, target, 0.1)
    np.testing.assert_allclose(_np(ls), ogp.testgen_loss(orc, x, target, 0.1, y_mean), rtol=1e-4, atol=0.015 / lay.D)
    fd = np.zeros_like(x)
    for q in range(Q):
        d = np.zeros_like(x); d[:, q] = 1e-6
        fd[:, q] = (ogp.testgen_loss(orc, x + d, target, 0.1, y_mean) - ogp.testgen_loss(orc, x - d, target, 0.1, y_mean)) / 2e-6
    assert _normrel(_np(g_x), fd) < 1e-3
    # optimisation touches X, alpha, every length-scale and the noise, and lowers the loss
    before = (float(lay.loss.eval()), _np(kern.gamma).copy(), _np(lay.X).copy())
    lay.optimize(tf.train.AdamOptimizer(learning_rate=0.01), num_iterations=30)
    assert float(lay.loss.eval()) < before[0]
    assert np.all(np.abs(_np(kern.gamma) - before[1]) > 1e-4) and np.max(np.abs(_np(lay.X) - before[2])) > 1e-3
    ck = kern.save(1, str(tmp_path))
    k2 = SEKernel(alpha=1.0, gamma=1.0, ARD=True, Q=Q)
    k2.restore(ck)
    np.testing.assert_allclose(_np(k2.gamma), _np(kern.gamma), rtol=0, atol=0)


@pytest.mark.parametrize('N,M,Q', [(328, None, 2), (70, None, 1), (1030, None, 3), (300, 45, 4), (129, 700, 5), (5000, None, 2),
                                   (4100, 4100, 7)])
def test_gram_kernel_every_latent_dimension(N, M, Q):
    """K3 is compiled per latent dimension (Q = 1..4) with a run-time-Q fallback, and tiles 32 or 128 rows per CTA
    depending on the matrix size: square (noise on the diagonal, exact alpha + noise there) and rectangular K_Xx forms
    against float64 (gplvm.py:61-86), ragged edges included."""
    from deepbelief_b200 import ops
    rs = np.random.RandomState(N + Q)
    X = rs.normal(size=(N, Q)).astype(np.float32)
    Y = None if M is None else rs.normal(size=(M, Q)).astype(np.float32)
    alpha, gamma, noise = 1.3, 0.8, 0.05
    hyp = torch.tensor([alpha, gamma, noise], device='cuda')
    got = ops.gram_rbf(torch.as_tensor(X).cuda(), None if Y is None else torch.as_tensor(Y).cuda(), hyp, add_noise=True)
    A = X.astype(np.float64) / gamma
    B = A if Y is None else Y.astype(np.float64) / gamma
    d2 = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1) if N * B.shape[0] <= 2_000_000 else None
    if d2 is None:                                     # big cases: row blocks
        got_np = got.cpu().numpy()
        for r0 in range(0, N, 997):
            blk = ((A[r0:r0 + 64, None, :] - B[None, :, :]) ** 2).sum(-1)
            want = alpha * np.exp(-0.5 * blk)
            if Y is None:
                want[np.arange(want.shape[0]), r0 + np.arange(want.shape[0])] += noise
            np.testing.assert_allclose(got_np[r0:r0 + 64], want, rtol=2e-5, atol=1e-6)
        return
    want = alpha * np.exp(-0.5 * d2)
    if Y is None:
        want[np.arange(N), np.arange(N)] += noise
        assert np.all(got.diagonal().cpu().numpy() == np.float32(np.float32(alpha) + np.float32(noise)))
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')
@pytest.mark.parametrize('form', ['f16x3', 'tf32x3'])
@pytest.mark.parametrize('M,N,K', [(512, 384, 1000), (784, 1300, 2500), (130, 70, 96)])
def test_split_precision_tensor_gemms_vs_fp64(form, M, N, K):
    """The two fp32-accurate tensor-core products of the GP chain (db_gemm_f16x3_nt: scaled fp16 hi/lo pairs, three
    kind::f16 passes; db_gemm_tf32x3_nt: 3xTF32) against fp64, on Gaussian operands and on operands whose entries span six
    decades (L^-1, K^-1 have such rows): error at most 2e-6 of the largest entry of the result, i.e. at the level of an fp32
    accumulation (tf.matmul in the reference, gplvm.py:235-244), far below the 1e-4 the losses and gradients must hold."""
    from deepbelief_b200 import ops
    rs = np.random.RandomState(M + N + K)
    if K % 8:                                     # the 3xTF32 form needs a 16-byte row pitch: pad K with zeros for both
        K = (K + 7) // 8 * 8
    for wide in (False, True):
        X = rs.normal(size=(M, K))
        Y = rs.normal(size=(N, K))
        if wide:
            X *= 10.0 ** rs.uniform(-6, 0, size=(M, K))
            Y *= 10.0 ** rs.uniform(-6, 0, size=(N, K))
        X32, Y32 = X.astype(np.float32), Y.astype(np.float32)
        want = X32.astype(np.float64) @ Y32.astype(np.float64).T
        fn = ops.gemm_f16x3_nt if form == 'f16x3' else ops.gemm_tf32x3_nt
        got = _np(fn(_cu(X32), _cu(Y32), alpha=1.0))
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err < 2e-6, (form, wide, err)
        C0 = rs.normal(size=(M, N)).astype(np.float32)
        got = _np(fn(_cu(X32), _cu(Y32), alpha=-0.5, beta=1.0, out=_cu(C0).clone()))
        err = np.abs(got - (C0 - 0.5 * want)).max() / np.abs(want).max()
        assert err < 2e-6, (form, wide, 'beta', err)
    if M == N:
        return
    S = rs.normal(size=(M, K)).astype(np.float32)                     # lower_only on a square product
    want = np.tril(S.astype(np.float64) @ S.astype(np.float64).T)
    fn = ops.gemm_f16x3_nt if form == 'f16x3' else ops.gemm_tf32x3_nt
    got = np.tril(_np(fn(_cu(S), _cu(S), lower_only=True)))
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-6
