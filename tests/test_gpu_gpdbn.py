"""GPDBN stack (two Concrete RBMs + GPRBM top layer) on the GPU against the reference-pinned oracle
(oracle/gprbm.py, goldens from the reference's own gprbm.py): loss pieces and every gradient array with the same
injected draws; a short Adam run against the oracle's Adam; the BASELINE config-4 size through properties."""
import numpy as np
import pytest

from conftest import Golden, has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')]

import torch  # noqa: E402

from oracle import gprbm as ogprbm, rbm as orbm  # noqa: E402
from test_oracle_gpdbn import CASES, get_draws, make_oracle  # noqa: E402

GRAD_NORMREL = 2e-4      # fp32 kernels vs fp64 oracle through Concrete units at T = 0.1 (steep sigmoids)


@pytest.fixture(scope='module')
def gpdbn_golden():
    return Golden('gpdbn_golden.npz')


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def _normrel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / (np.max(np.abs(np.asarray(b))) + 1e-300))


def build_stack(g, cname):
    from deepbelief_b200.layers.rbm import RBM
    from deepbelief_b200.layers.gprbm import GPRBM
    from deepbelief_b200.layers.gplvm import SEKernel
    P = {k: g['gpdbn/P/' + k] for k in ('W0', 'b0', 'c0', 'W1', 'b1', 'c1', 'Wt', 'bt', 'ct', 'X')}
    T = float(g['gpdbn/T'])
    V = g['gpdbn/V']
    N, V0 = V.shape
    Ha, Hb, D = P['W0'].shape[1], P['W1'].shape[1], P['Wt'].shape[1]
    l0 = RBM(V0, Ha, temperature=T, name='l0', seed=1)
    l0.set_weights(P['W0'], P['b0'], P['c0'])
    l1 = RBM(Ha, Hb, temperature=T, bottom=l0, name='l1', seed=2)
    l1.set_weights(P['W1'], P['b1'], P['c1'])
    if cname == 'fixed':
        kern = SEKernel(alpha=False, gamma=False)
        top = GPRBM(Hb, D, 2, N, False, X0=P['X'], V=V, kern=kern, noise_variance=0.05, fixed_noise_variance=True,
                    temperature=T, bottom=l1, name='top', seed=3)
    else:
        kern = SEKernel(alpha=1.4, gamma=0.9)
        top = GPRBM(Hb, D, 2, N, False, X0=P['X'], V=V, kern=kern, noise_variance=0.3, fixed_noise_variance=False,
                    temperature=T, bottom=l1, name='top', seed=3)
    top.set_weights(P['Wt'], P['bt'], P['ct'])
    return top, (l0, l1)


@pytest.mark.parametrize('cname', list(CASES))
def test_loss_and_gradients_vs_oracle(gpdbn_golden, cname):
    g = gpdbn_golden
    top, (l0, l1) = build_stack(g, cname)
    draws = get_draws(g)
    loss, pieces = top.loss_and_grads(draws=draws)
    st, og = make_oracle(g, cname).grads(draws)
    np.testing.assert_allclose(_np(pieces['pred_var']), st['pv'], rtol=2e-4)
    np.testing.assert_allclose(float(pieces['log_det']), st['logdet'], rtol=1e-4)
    assert _normrel(_np(pieces['H1']), st['H1']) < 1e-4
    assert _normrel(_np(pieces['Hm']), st['Hm']) < 1e-4
    assert _normrel(_np(pieces['V_sample']), st['V_sample']) < 1e-3      # Concrete at T = 0.1 amplifies by ~1/T
    np.testing.assert_allclose(float(pieces['gp_loss']), st['gp_loss'], rtol=1e-4)
    # the cross entropy of saturated Concrete samples clips at 1 - 1e-6, where fp32 resolves 1 - p to ~6 %: the
    # reference's own fp32 run (config.float_type) differs from its fp64 run by this much (SURVEY.md §7.3 rule)
    ref64, ref32 = float(g['gpdbn/%s/float64/loss' % cname]), float(g['gpdbn/%s/float32/loss' % cname])
    budget = max(2e-4, 2.0 * abs(ref32 - ref64) / abs(ref64))
    np.testing.assert_allclose(float(loss), st['loss'], rtol=budget)
    np.testing.assert_allclose(float(loss), ref64, rtol=budget)   # the reference itself
    G = top._grads
    for li, lay in enumerate((l0, l1)):
        gW, gb, gc = top._views(lay, G['lower'][li])
        assert _normrel(_np(gW), og['lower'][li]['W']) < GRAD_NORMREL, 'W%d' % li
        assert _normrel(_np(gb), og['lower'][li]['b']) < GRAD_NORMREL, 'b%d' % li
        assert _normrel(_np(gc), og['lower'][li]['c']) < GRAD_NORMREL, 'c%d' % li
    gW, gb, gc = top._views(top, G['top'])
    assert _normrel(_np(gW), og['top']['W']) < GRAD_NORMREL
    assert _normrel(_np(gb), og['top']['b']) < GRAD_NORMREL
    assert _normrel(_np(gc), og['top']['c']) < GRAD_NORMREL
    n = top.num_v * top.num_h + top.num_v + top.num_h
    assert _normrel(_np(G['top'][n:]), og['top']['opt_alpha_h']) < GRAD_NORMREL
    assert _normrel(_np(G['X']).reshape(og['X'].shape), og['X']) < GRAD_NORMREL
    want_h = np.array([float(og['opt_alpha']), float(og['opt_gamma']), float(og['opt_noise'])])
    if cname == 'trainable':
        np.testing.assert_allclose(_np(G['hyp']), want_h, rtol=5e-4)
    else:
        assert np.all(_np(G['hyp']) == 0)
    # and against the directional finite differences of the REFERENCE's loss
    flat = {'W0': top._views(l0, G['lower'][0])[0], 'c1': top._views(l1, G['lower'][1])[2], 'Wt': gW,
            'X': G['X'].view(top.N, top.Q)}
    for key, arr in flat.items():
        d = g['gpdbn/%s/dir/%s' % (cname, key)]
        fd = float(g['gpdbn/%s/fd/%s' % (cname, key)])
        assert abs(float(np.sum(_np(arr) * d)) - fd) <= 5e-4 * max(abs(fd), 1.0), key


def test_optimize_follows_oracle_adam(gpdbn_golden):
    """3 Adam iterations with injected draws reproduce the oracle's parameters (same draws every iteration)."""
    from deepbelief_b200 import compat as tf, ops
    g = gpdbn_golden
    top, (l0, l1) = build_stack(g, 'trainable')
    draws = get_draws(g)
    opt = tf.train.AdamOptimizer(learning_rate=1e-3)
    opt.reset()
    states = {}
    o = make_oracle(g, 'trainable')
    adams = {}

    def ostep(key, theta, grad):
        if key not in adams:
            adams[key] = orbm.Adam(1e-3)
        return adams[key].step(theta, grad)
    for it in range(3):
        top.loss_and_grads(draws=draws)
        desc = opt.descriptor()
        for params, grad, lay in top._trainable_buffers(False):
            st = states.setdefault(params.data_ptr(), torch.zeros(2 * params.numel(), device='cuda'))
            ops.optimizer_step(desc, params, grad, st)
            if lay is not None:
                lay._weights_version += 1
        top._touch()
        _, og = o.grads(draws)
        for li in range(2):
            for k in ('W', 'b', 'c'):
                o.lower[li][k] = ostep(('l', li, k), o.lower[li][k], og['lower'][li][k])
        for k in ('W', 'b', 'c', 'opt_alpha_h'):
            o.top[k] = ostep(('t', k), o.top[k], og['top'][k])
        o.gp['X'] = ostep('X', o.gp['X'], og['X'])
        for k in ('opt_alpha', 'opt_gamma', 'opt_noise'):
            o.gp[k] = ostep(k, np.float64(o.gp[k]), np.float64(og[k]))
    # Adam's first steps are ~ lr * sign(g): compare with an absolute budget of a fraction of the total movement
    np.testing.assert_allclose(_np(top.X), o.gp['X'], atol=3e-4)
    np.testing.assert_allclose(_np(l0.W), o.lower[0]['W'], atol=3e-4)
    np.testing.assert_allclose(_np(top.W), o.top['W'], atol=3e-4)
    np.testing.assert_allclose(_np(top._hyp), [o.gp['opt_alpha'], o.gp['opt_gamma'], o.gp['opt_noise']], atol=3e-4)


def test_optimize_api_and_checkpoint(gpdbn_golden, tmp_path):
    """GPRBM.optimize with Philox draws: loss falls on average, fixed-X phase leaves X untouched, save/restore."""
    from deepbelief_b200 import compat as tf
    g = gpdbn_golden
    top, _ = build_stack(g, 'fixed')
    X0 = top.X.clone()
    l_start = np.mean([float(top.loss_and_grads(want_grads=False)[0]) for _ in range(8)])
    top.optimize(tf.train.AdamOptimizer(learning_rate=5e-3), num_iterations=10, fixed_X_num_iterations=10)
    assert torch.equal(top.X, X0)
    top.optimize(tf.train.AdamOptimizer(learning_rate=5e-3), num_iterations=60, ckpt_interval=60, ckpt_dir=str(tmp_path))
    assert not torch.equal(top.X, X0)
    l_end = np.mean([float(top.loss_and_grads(want_grads=False)[0]) for _ in range(8)])
    assert np.isfinite(l_end) and l_end < l_start
    other, _ = build_stack(g, 'fixed')
    other.restore(str(tmp_path / 'top-60'))
    assert torch.equal(other._params, top._params) and torch.equal(other.X, top.X)
    # eval-mode style generation from the snapshots
    img = top.downprop_latent(np.zeros(2, dtype=np.float32), num_samples_to_average=16)
    assert img.shape == (top.datapoint_len(),) and np.all(np.isfinite(img))
    assert top.predict_variance(np.zeros((3, 2))).shape == (3, 1)


def _config4_stack(N, seed=0):
    from deepbelief_b200.layers.rbm import RBM
    from deepbelief_b200.layers.gprbm import GPRBM
    from deepbelief_b200.layers.gplvm import SEKernel
    rs = np.random.RandomState(seed)
    V0, Ha, Hb, D, T = 784, 200, 100, 50, 0.1
    V = (rs.rand(N, V0) < 0.13).astype(np.float32)
    X0 = rs.normal(size=(N, 2)).astype(np.float32)
    np.random.seed(1)
    l0 = RBM(V0, Ha, W_std=0.05, temperature=T, name='c4l0', seed=11)
    l1 = RBM(Ha, Hb, W_std=0.05, temperature=T, bottom=l0, name='c4l1', seed=12)
    top = GPRBM(Hb, D, 2, N, False, X0=X0, V=V, kern=SEKernel(alpha=False, gamma=False), noise_variance=0.01,
                fixed_noise_variance=True, W_std=0.05, temperature=T, bottom=l1, name='c4top', seed=13)
    gen = torch.Generator(device='cuda').manual_seed(5)
    U = lambda r, c: torch.rand(r, c, device='cuda', generator=gen).clamp_(1e-4, 1 - 1e-4)  # noqa: E731
    draws = {'up': [U(N, Ha), U(N, Hb)], 'eps1': torch.randn(N, D, device='cuda', generator=gen),
             'eps2': torch.randn(N, D, device='cuda', generator=gen), 'down_top': U(N, Hb), 'down': [U(N, Ha), U(N, V0)]}
    return top, (l0, l1), V, draws


def test_config4_size_properties():
    """BASELINE config 4 size: synthetic binary 5000x784 (Bernoulli 0.13), stack 784->200->100->GPRBM(50), Q=2, T=0.1,
    noise 0.01 fixed: pred_var in (0, noise], finite loss and gradients, loss = CE + gp_loss, Philox run works."""
    N = 5000
    top, _, V, draws = _config4_stack(N)
    loss, pieces = top.loss_and_grads(draws=draws)
    pv = _np(pieces['pred_var'])
    assert np.all(pv > 0) and np.all(pv <= 0.01 * (1 + 1e-5))
    assert np.isfinite(float(loss))
    np.testing.assert_allclose(float(loss), float(pieces['ce']) + float(pieces['gp_loss']), rtol=1e-6)
    G = top._grads
    for t in [G['top'], G['X'], G['hyp']] + G['lower']:
        assert bool(torch.isfinite(t).all())
    assert float(G['X'].abs().max()) > 0 and float(G['lower'][0].abs().max()) > 0
    loss2, _ = top.loss_and_grads()              # Philox draws
    assert np.isfinite(float(loss2))


@pytest.mark.parametrize('N', [1200, 5000])
def test_config4_layer_sizes_vs_oracle(N):
    """The config-4 stack (784->200->100->50, T=0.1, noise 0.01 fixed: cond(K) ~ 1e4-1e5) at N=1200 and at the FULL
    BASELINE size N=5000 (the fp64 numpy oracle needs ~20 s there): loss pieces and every gradient array against
    the oracle."""
    top, (l0, l1), V, draws = _config4_stack(N, seed=3)
    loss, pieces = top.loss_and_grads(draws=draws)
    cpu = lambda t: t.detach().cpu().numpy().astype(np.float64)  # noqa: E731
    lower = [dict(W=cpu(l.W), b=cpu(l.b).ravel(), c=cpu(l.c).ravel()) for l in (l0, l1)]
    topd = dict(W=cpu(top.W), b=cpu(top.b).ravel(), c=cpu(top.c).ravel(), opt_alpha_h=cpu(top.opt_alpha_h).ravel())
    gp = dict(X=cpu(top.X), opt_alpha=0.0, opt_gamma=0.0, opt_noise=0.0, trainable=(False, False, False), fixed_noise=0.01)
    o = ogprbm.GPDBNOracle(lower, topd, gp, V, 0.1)
    nd = {k: ([cpu(a) for a in v] if isinstance(v, list) else cpu(v)) for k, v in draws.items()}
    vs = pieces['V_sample']
    mask = ((vs >= 1.0e-6) & (vs <= torch.tensor(1.0, device='cuda') - 1.0e-6)).cpu().numpy()   # the kernel's fp32 mask
    st, og = o.grads(nd, clip_mask=mask)
    assert np.mean(mask != st['clip_mask']) < 2e-3        # tie band of the clip step (see oracle.grads)
    # sigma^2 = 0.01 without jitter: cond(K) ~ 1e5 at N = 1200, so fp32 carries ~cond * eps. The budget per array is
    # max(2e-3, 2 x the error of the reference's own formulas evaluated in fp32) — SURVEY.md §7.3 rule
    o32 = ogprbm.GPDBNOracle(lower, topd, gp, V, 0.1, dtype=np.float32)
    try:
        _, og32 = o32.grads(nd, clip_mask=mask)
    except FloatingPointError:          # the fp32 reference formula may trip its own assert_non_negative here
        og32 = None

    def budget(got64, key_path):
        if og32 is None:
            return 5e-3
        ref32 = og32
        for k in key_path:
            ref32 = ref32[k]
        return max(2e-3, 2.0 * _normrel(np.asarray(ref32, dtype=np.float64), got64))
    np.testing.assert_allclose(_np(pieces['pred_var']), st['pv'], rtol=5e-3, atol=3e-6)   # s - s^2 K^-1_ii cancels ~60x
    np.testing.assert_allclose(float(pieces['gp_loss']), st['gp_loss'], rtol=1e-4)
    np.testing.assert_allclose(float(loss), st['loss'], rtol=1e-3)
    G = top._grads
    for li, lay in enumerate((l0, l1)):
        gW, gb, gc = top._views(lay, G['lower'][li])
        assert _normrel(_np(gW), og['lower'][li]['W']) < budget(og['lower'][li]['W'], ('lower', li, 'W'))
        assert _normrel(_np(gb), og['lower'][li]['b']) < budget(og['lower'][li]['b'], ('lower', li, 'b'))
        assert _normrel(_np(gc), og['lower'][li]['c']) < budget(og['lower'][li]['c'], ('lower', li, 'c'))
    gW, gb, gc = top._views(top, G['top'])
    assert _normrel(_np(gW), og['top']['W']) < budget(og['top']['W'], ('top', 'W'))
    assert _normrel(_np(gc), og['top']['c']) < budget(og['top']['c'], ('top', 'c'))
    assert _normrel(_np(G['X']).reshape(og['X'].shape), og['X']) < budget(og['X'], ('X',))


def test_latent_search_loss_and_gradient_vs_oracle(gpdbn_golden, tmp_path):
    """GPRBM.test_generalisation (gprbm.py:468-562): per-restart loss against the oracle's line-by-line restatement of
    gprbm.py:490-499 with the same injected draws, the x-gradient against central differences of that oracle, and the
    batched driver loop (table format, distance no worse than the best start)."""
    from deepbelief_b200 import compat as tf
    g = gpdbn_golden
    top, (l0, l1) = build_stack(g, 'trainable')
    o = make_oracle(g, 'trainable')
    rs = np.random.RandomState(11)
    N, D, Q = top.N, top.D, top.Q
    H2 = rs.normal(size=(N, D)).astype(np.float32)
    H1_mean = rs.normal(size=D).astype(np.float32) * 0.3
    top.H2_var.copy_(torch.as_tensor(H2).cuda())
    top.H1_mean_var.copy_(torch.as_tensor(H1_mean).cuda())
    R, S = 5, 3
    x = rs.normal(size=(R, Q)).astype(np.float32)
    Vt, Va, V0 = top.num_v, l1.num_v, l0.num_v
    draws = {'eps': rs.normal(size=(R * S, D)).astype(np.float32),
             'down_top': rs.uniform(0.05, 0.95, size=(R * S, Vt)).astype(np.float32),
             'down': [rs.uniform(0.05, 0.95, size=(R * S, Va)).astype(np.float32),
                      rs.uniform(0.05, 0.95, size=(R * S, V0)).astype(np.float32)]}
    target = g['gpdbn/V'][:1].astype(np.float32)
    loss, g_x, info, mp, pv = top.latent_search_loss(x, target, 0.1, S, draws=draws, want_outputs=True)
    assert int(info.item()) == 0
    want = ogprbm.latent_search_loss(o, H2, H1_mean, x, target, 0.1, S, draws)
    np.testing.assert_allclose(_np(loss), want, rtol=2e-4, atol=0.015 / V0)
    eps = 1e-4
    fd = np.zeros((R, Q))
    for q in range(Q):
        d = np.zeros((R, Q)); d[:, q] = eps
        fd[:, q] = (ogprbm.latent_search_loss(o, H2, H1_mean, x + d, target, 0.1, S, draws) -
                    ogprbm.latent_search_loss(o, H2, H1_mean, x - d, target, 0.1, S, draws)) / (2 * eps)
    # entries of the sample mean that sit at the 1e-6 clip of dist.cross_entropy switch their gradient on/off (tie band,
    # see oracle.grads): most components agree to ~1e-4, a restart with such an element to ~1e-2
    err = np.abs(_np(g_x) - fd) / np.max(np.abs(fd))
    assert np.max(err) < 1e-2 and np.median(err) < 3e-4
    # driver: the snapshots are refreshed from a fresh upprop, all training latents are tried (method 'all_training')
    table = top.test_generalisation(g['gpdbn/V'][:2], str(tmp_path / 't.pkl'), tf.train.GradientDescentOptimizer(learning_rate=0.01),
                                    num_iterations=10, log_var_scaling=0.1, method='all_training', num_runs=N,
                                    num_samples_to_average=4)
    assert len(table) == 2 and table[0][2].shape == (1, Q) and table[0][4].shape == (V0,) and np.isfinite(table[0][1])
    t2 = top.test_generalisation(g['gpdbn/V'][:1], str(tmp_path / 't2.pkl'), tf.train.GradientDescentOptimizer(learning_rate=0.01),
                                 num_iterations=3, method='random_training', num_runs=6, num_samples_to_average=2)
    assert len(t2) == 1


def _latent_draws(lg, pre):
    f = lambda a: np.asarray(a, dtype=np.float32)  # noqa: E731
    return {'eps': f(lg[pre + 'draws/eps']), 'down_top': f(lg[pre + 'draws/down_top']),
            'down': [f(lg[pre + 'draws/down1']), f(lg[pre + 'draws/down0'])]}


def test_held_out_data_vs_reference_and_oracle(gpdbn_golden, tmp_path):
    """GPRBM.test_held_out_data (gprbm.py:440-466): the objective at given x against the value the reference's own
    graph pieces produce (tests/golden/latent_golden.npz) and against the oracle, its x-gradient against central
    differences of the oracle, and the driver loop (saved x, loss not worse than at the start on fixed draws)."""
    from deepbelief_b200 import compat as tf
    from conftest import Golden
    g, lg = gpdbn_golden, Golden('latent_golden.npz')
    top, _ = build_stack(g, 'trainable')
    o = make_oracle(g, 'trainable')
    H2, H1_mean = lg['latent/H2'], lg['latent/H1_mean']
    top.H2_var.copy_(torch.as_tensor(H2, dtype=torch.float32).cuda())
    top.H1_mean_var.copy_(torch.as_tensor(H1_mean, dtype=torch.float32).cuda().reshape(top.H1_mean_var.shape))
    x, T = lg['heldout/x'].astype(np.float32), lg['heldout/test_data'].astype(np.float32)
    draws = _latent_draws(lg, 'heldout/')
    loss, g_x, info = top.held_out_loss(x, T, draws=draws)
    assert int(info.item()) == 0
    np.testing.assert_allclose(_np(loss), lg['heldout/float64/mse'], rtol=1e-4)          # the reference's own value
    x64 = x.astype(np.float64)
    np.testing.assert_allclose(_np(loss), ogprbm.held_out_loss(o, H2, H1_mean, x64, T, draws), rtol=1e-4)
    eps = 1e-4
    fd = np.zeros_like(x64)
    for q in range(x.shape[1]):
        d = np.zeros_like(x64); d[:, q] = eps
        fd[:, q] = (ogprbm.held_out_loss(o, H2, H1_mean, x64 + d, T, draws) -
                    ogprbm.held_out_loss(o, H2, H1_mean, x64 - d, T, draws)) / (2 * eps)
    assert _normrel(_np(g_x), fd) < 1e-3
    # driver: eval-mode call pattern (x_test_var from the constructor), SGD on fresh draws, result saved with np.save
    top.eval_flag = True
    top.x_test = tf.get_variable(name='x_test', initializer=x[:1])
    out = str(tmp_path / 'x_held_out.npy')
    x_opt = top.test_held_out_data(T, out, tf.train.AdamOptimizer(learning_rate=0.05), num_iterations=60)
    assert x_opt.shape == (1, top.Q) and np.array_equal(np.load(out), x_opt)
    l0 = np.mean([float(top.held_out_loss(x[:1], T)[0][0]) for _ in range(40)])
    l1 = np.mean([float(top.held_out_loss(x_opt, T)[0][0]) for _ in range(40)])
    assert np.isfinite(l1) and l1 < l0 + 0.05 * abs(l0)


def test_interp_test_vs_reference_and_oracle(gpdbn_golden):
    """GPRBM.interp_test (gprbm.py:564-689): images and predictive variances along a path against
    create_observed_prediction_with_variance run from the reference source, the path objective and its gradient against
    the oracle, and the driver (objective decreases from the straight line; output shapes)."""
    from conftest import Golden
    g, lg = gpdbn_golden, Golden('latent_golden.npz')
    top, _ = build_stack(g, 'trainable')
    o = make_oracle(g, 'trainable')
    H2, H1_mean = lg['latent/H2'], lg['latent/H1_mean']
    top.H2_var.copy_(torch.as_tensor(H2, dtype=torch.float32).cuda())
    top.H1_mean_var.copy_(torch.as_tensor(H1_mean, dtype=torch.float32).cuda().reshape(top.H1_mean_var.shape))
    Xi, S = lg['interp/Xi'], int(lg['interp/S'])
    imgs, pv = top.create_observed_prediction_with_variance(Xi.astype(np.float32), S, draws=_latent_draws(lg, 'interp/'))
    np.testing.assert_allclose(_np(pv), lg['interp/float64/pred_var'], rtol=2e-4)
    assert _normrel(_np(imgs), lg['interp/float64/images']) < 1e-4
    lam = 0.1
    obj, g_x, info = top.interp_objective(Xi[1:-1].astype(np.float32), Xi[0].astype(np.float32), Xi[-1].astype(np.float32), lam)
    want, _ = ogprbm.interp_objective(o, Xi[1:-1], Xi[0], Xi[-1], lam)
    assert int(info.item()) == 0
    np.testing.assert_allclose(float(obj), want, rtol=1e-4)
    eps = 1e-5
    fd = np.zeros_like(Xi[1:-1])
    for i in range(fd.shape[0]):
        for q in range(fd.shape[1]):
            d = np.zeros_like(fd); d[i, q] = eps
            fd[i, q] = (ogprbm.interp_objective(o, Xi[1:-1] + d, Xi[0], Xi[-1], lam)[0] -
                        ogprbm.interp_objective(o, Xi[1:-1] - d, Xi[0], Xi[-1], lam)[0]) / (2 * eps)
    assert _normrel(_np(g_x), fd) < 1e-3
    # driver: the same path start as the reference builds, objective not worse after Adam
    path0 = ogprbm.interp_path(_np(top.X)[0], _np(top.X)[5], 4).astype(np.float32)
    start = float(top.interp_objective(path0, path0[0], path0[-1], lam)[0])
    lin, geo = top.interp_test(0, 5, num_points=4, interp_lambda=lam, learning_rate=0.01, num_iterations=50,
                               num_sample_to_average=8)
    V0 = top.datapoint_len()
    assert lin.shape == (6, V0) and geo.shape == (6, V0) and np.isfinite(geo).all()
    end = float(top.interp_objective(top.interp_latents, path0[0], path0[-1], lam)[0])
    assert end <= start + 1e-6


@pytest.mark.parametrize('opt_name', ['adam', 'momentum'])
def test_optimize_as_cuda_graph_matches_the_eager_loop(monkeypatch, opt_name):
    """GPRBM.optimize replays one captured CUDA graph per iteration (replay.ReplayedStep: Philox draw ids and the optimizer
    step number come from a device clock) once a mode has run eagerly — across the fixed-X / free-X switch and with
    evaluation steps (which draw eagerly) in between. The host counters must end identical, and every trainable of the stack
    must agree with the eager loop as closely as two eager runs agree with each other (the loss reductions use atomics, so
    the eager loop itself repeats only to ~1e-6..1e-5 after 23 iterations; one wrong draw id would move parameters by
    ~lr per iteration, 1e-3 and more by the end)."""
    from deepbelief_b200 import compat as tf
    from deepbelief_b200.layers.gprbm import GPRBM
    from deepbelief_b200 import replay

    def run(graph):
        monkeypatch.setenv('DEEPBELIEF_B200_GRAPH', '1' if graph else '0')
        monkeypatch.setattr(GPRBM, 'GRAPH_MIN_ITERATIONS', 4)
        top, lower, _, _ = _config4_stack(192, seed=3)
        opt = (tf.train.AdamOptimizer(learning_rate=2e-3) if opt_name == 'adam'
               else tf.train.MomentumOptimizer(learning_rate=1e-4, momentum=0.9))
        created = []
        orig = replay.ReplayedStep.__init__

        def spy(self, *a, **k):
            created.append(1)
            orig(self, *a, **k)
        monkeypatch.setattr(replay.ReplayedStep, '__init__', spy)
        top.optimize(opt, num_iterations=23, fixed_X_num_iterations=9, eval_interval=5)
        monkeypatch.setattr(replay.ReplayedStep, '__init__', orig)
        assert len(created) == (2 if graph else 0)                       # one graph per mode (X fixed / X free)
        return [top._params.clone(), top.X.clone(), top._hyp.clone()] + [l._params.clone() for l in lower], \
            [l.rng.draw for l in top.layers], opt.step_count

    p_graph, d_graph, s_graph = run(True)
    p_eager, d_eager, s_eager = run(False)
    p_again, _, _ = run(False)
    assert (d_graph, s_graph) == (d_eager, s_eager)
    noise = max(float((a - b).abs().max()) for a, b in zip(p_eager, p_again))
    diff = max(float((a - b).abs().max()) for a, b in zip(p_graph, p_eager))
    assert all(bool(torch.isfinite(a).all()) for a in p_graph)
    assert diff <= max(10.0 * noise, 2e-4), (diff, noise)     # (measured: noise 6e-6..1.3e-5, diff 5e-6..2.1e-5)
