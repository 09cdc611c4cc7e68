"""GPU parity tests of the RBM / CD path: CUDA (through the layer API and the C ABI) vs the CPU oracle
on the same seeded inputs and the same injected random draws.

Tolerances: sampled binary states bit-exact (uniforms are injected with a guaranteed
margin |p - u| > 1e-4, or mismatches must be confined to rows holding a near-tie); probabilities
1e-5 relative; gradients 1e-4 norm-relative, all against the fp64 oracle.
"""
import numpy as np
import pytest
import torch

from conftest import has_cuda
from oracle import philox, rbm as orbm

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')]

KINDS = ['rbm', 'bgrbm', 'gbrbm', 'ggrbm']
PROB_RTOL = 1e-5
GRAD_NORMREL = 1e-4


def _layer(kind, V, H, **kw):
    from deepbelief_b200.layers import rbm, bgrbm, gbrbm, ggrbm
    cls = {'rbm': rbm.RBM, 'bgrbm': bgrbm.BGRBM, 'gbrbm': gbrbm.GBRBM, 'ggrbm': ggrbm.GGRBM}[kind]
    return cls(num_v=V, num_h=H, **kw)


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def _normrel(a, b):
    return np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)


def _pair(g, tag, kind, temperature=None):
    """(CUDA layer, fp64 oracle) sharing the golden parameters."""
    V, H = g[tag + '/W'].shape
    kw = {} if kind == 'ggrbm' else {'temperature': temperature}
    lay = _layer(kind, V, H, **kw)
    if lay.layer_kind in (1, 3):
        lay.set_weights(g[tag + '/W'], g[tag + '/b'], g[tag + '/c'], opt_sigma_h=g[tag + '/opt_sigma'])
    else:
        lay.set_weights(g[tag + '/W'], g[tag + '/b'], g[tag + '/c'])
    orc = orbm.RBMOracle(g[tag + '/W'], g[tag + '/b'], g[tag + '/c'], kind=kind, opt_sigma_h=g[tag + '/opt_sigma'],
                         temperature=temperature)
    return lay, orc


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('tag', ['toy', 'mid'])
def test_layer_forward_vs_oracle_and_reference_goldens(rbm_golden, tag, kind):
    g = rbm_golden
    lay, orc = _pair(g, tag, kind)
    v = g[tag + ('/v_real' if orc.gaussian_visible else '/v_bin')]
    h = g[tag + ('/h_real' if orc.gaussian_hidden else '/h_bin')]
    pre = '%s/%s/float64/' % (tag, kind)
    for got, want in ((lay.h_net_input(v), orc.h_net_input(v)), (lay.v_net_input(h), orc.v_net_input(h)),
                      (lay.h_probs(v), orc.h_probs(v)), (lay.v_probs(h), orc.v_probs(h))):
        np.testing.assert_allclose(_np(got), want, rtol=PROB_RTOL, atol=2e-6)
    # against the goldens produced by the reference's own code
    np.testing.assert_allclose(_np(lay.h_probs(v)), g[pre + 'h_probs'], rtol=PROB_RTOL, atol=2e-6)
    np.testing.assert_allclose(_np(lay.v_probs(h)), g[pre + 'v_probs'], rtol=PROB_RTOL, atol=2e-6)
    np.testing.assert_allclose(_np(lay.free_energy(v)), g[pre + 'free_energy'], rtol=2e-5, atol=2e-5)
    vk = (0.7 * g[tag + "/v_real"][::-1] + 0.05) if orc.gaussian_visible else g[tag + '/vk_bin']
    np.testing.assert_allclose(float(lay.cd_loss(v, vk)), float(g[pre + 'cd_loss']), rtol=1e-3, atol=2e-4)


@pytest.mark.parametrize('kind', KINDS)
def test_sampling_with_injected_draws(rbm_golden, kind):
    g = rbm_golden
    lay, orc = _pair(g, 'mid', kind)
    pre = 'mid/%s/float64/' % kind
    hp, vp = g[pre + 'h_probs'], g[pre + 'v_probs']
    draw_h = g['mid/E_h'] if orc.gaussian_hidden else g['mid/U_h']
    sh = _np(lay.sample_h(hp.astype(np.float32), draw=torch.as_tensor(draw_h.astype(np.float32)).cuda()))
    sv = _np(lay.sample_v(vp.astype(np.float32), draw=torch.as_tensor(g['mid/U_v'].astype(np.float32)).cuda()))
    if orc.gaussian_hidden:
        np.testing.assert_allclose(sh, g[pre + 'sample_h'], rtol=1e-5, atol=1e-6)
    else:
        safe = np.abs(hp - draw_h) > 1e-6
        np.testing.assert_array_equal(sh[safe], g[pre + 'sample_h'][safe])          # bit-exact away from ties
        assert set(np.unique(sh)) <= {0.0, 1.0}
    if orc.gaussian_visible:
        np.testing.assert_allclose(sv, g[pre + 'sample_v'], rtol=1e-5, atol=1e-6)   # the "sample" is the mean
    else:
        safe = np.abs(vp - g['mid/U_v']) > 1e-6
        np.testing.assert_array_equal(sv[safe], g[pre + 'sample_v'][safe])


@pytest.mark.parametrize('kind', ['rbm', 'bgrbm', 'gbrbm'])
def test_concrete_units(rbm_golden, kind):
    g = rbm_golden
    lay, orc = _pair(g, 'mid', kind, temperature=0.1)
    pre = 'mid/%s/float64/' % kind
    if not orc.gaussian_hidden:
        got = _np(lay.sample_h(g[pre + 'h_probs'].astype(np.float32), draw=torch.as_tensor(g['mid/U_h'].astype(np.float32)).cuda()))
        np.testing.assert_allclose(got, g[pre + 'concrete_h'], rtol=2e-4, atol=2e-5)
    if not orc.gaussian_visible:
        got = _np(lay.sample_v(g[pre + 'v_probs'].astype(np.float32), draw=torch.as_tensor(g['mid/U_v'].astype(np.float32)).cuda()))
        np.testing.assert_allclose(got, g[pre + 'concrete_v'], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('tag', ['toy', 'mid'])
def test_cd_gradient_vs_oracle(rbm_golden, tag, kind):
    g = rbm_golden
    lay, orc = _pair(g, tag, kind)
    v = g[tag + ('/v_real' if orc.gaussian_visible else '/v_bin')]
    vk = (0.7 * g[tag + "/v_real"][::-1] + 0.05) if orc.gaussian_visible else g[tag + '/vk_bin']
    got = _np(lay.cd_gradient(v, vk))
    want = orc.cd_grad_packed(v, vk)
    V, H = orc.V, orc.H
    for lo, hi in ((0, V * H), (V * H, V * H + V), (V * H + V, V * H + V + H), (V * H + V + H, got.size)):
        if hi > lo:
            assert _normrel(got[lo:hi], want[lo:hi]) < GRAD_NORMREL
    if tag == 'toy':      # and against finite differences of the REFERENCE's cd_loss
        assert _normrel(got[:V * H].reshape(V, H), g['toy/%s/fd_grad_W' % kind]) < 1e-4


def test_gibbs_chain_and_stack_traversal(rbm_golden):
    g = rbm_golden
    lay, orc = _pair(g, 'mid', 'rbm')
    k = int(g['mid/k'])
    cu = lambda a: torch.as_tensor(a.astype(np.float32)).cuda()  # noqa: E731
    draws = [(cu(g['mid/chain_v'][s]), cu(g['mid/chain_h_u'][s])) for s in range(k)]
    vp, v, hp, h = lay.gibbs_sample(g['mid/h_bin'], k=k, draws=draws)
    pre = 'mid/rbm/float64/'
    np.testing.assert_array_equal(_np(v), g[pre + 'gibbs_v'])
    np.testing.assert_array_equal(_np(h), g[pre + 'gibbs_h'])
    np.testing.assert_allclose(_np(vp), g[pre + 'gibbs_v_probs'], rtol=PROB_RTOL, atol=1e-6)
    np.testing.assert_allclose(_np(hp), g[pre + 'gibbs_h_probs'], rtol=PROB_RTOL, atol=1e-6)
    # two-layer stack (rbm.py:277-325)
    from deepbelief_b200.layers.rbm import RBM
    l1 = RBM(12, 8, name='l1')
    l1.set_weights(g['stack/W1'], g['stack/b1'], g['stack/c1'])
    l2 = RBM(8, 5, bottom=l1, name='l2')
    l2.set_weights(g['stack/W2'], g['stack/b2'], g['stack/c2'])
    up = l2.upprop(g['stack/v'], draws=[cu(g['stack/u1']), cu(g['stack/u2'])])
    np.testing.assert_array_equal(_np(up), g['stack/up'])
    down = l2.downprop(up, draws=[cu(g['stack/u3']), cu(g['stack/u4'])])
    np.testing.assert_array_equal(_np(down), g['stack/down'])
    assert l2.layers == [l2, l1] and l2.datapoint_len() == 12


def test_reference_unit_test_properties():
    """Re-expression of the reference's tests/test_rbm.py, test_bgrbm.py, test_gbrbm.py (toy 6x5, batch 20)."""
    rs = np.random.RandomState(0)
    rbm = _layer('rbm', 6, 5, W_std=0.01)
    v = rs.binomial(1, 0.5, size=(20, 6)).astype(np.float32)
    h = rs.binomial(1, 0.5, size=(20, 5)).astype(np.float32)
    sh, sv = _np(rbm.sample_h(rbm.h_probs(v))), _np(rbm.sample_v(rbm.v_probs(h)))
    assert sh.shape == (20, 5) and sv.shape == (20, 6)
    assert np.all(np.isin(sh, [0.0, 1.0])) and np.all(np.isin(sv, [0.0, 1.0]))
    assert tuple(rbm.free_energy(v).shape) == (20, 1)
    bg = _layer('bgrbm', 6, 5, W_std=0.01)
    m = _np(bg.sample_h(bg.h_probs(v), sample=False))
    assert m.shape == (20, 5)
    np.testing.assert_allclose(m.mean(0), 0, atol=0.1)
    np.testing.assert_allclose(m.std(0), 0, atol=0.1)
    gb = _layer('gbrbm', 6, 5, W_std=0.01)
    r = _np(gb.sample_v(gb.v_probs(h)))
    assert r.shape == (20, 6)
    np.testing.assert_allclose(r.mean(0), 0, atol=0.1)
    np.testing.assert_allclose(r.std(0), 0, atol=0.1)
    with pytest.raises(AssertionError):
        _layer('rbm', 1, 5)                        # rbm.py:44-45
    with pytest.raises(AssertionError):
        rbm.gibbs_sample(h, k=0)                   # rbm.py:364


def test_horses_anchors(horses):
    """SURVEY.md §8(c) anchor (2) on the bundled dataset, V=1024, H=500 (the SIMT path: 500 % 8 != 0)."""
    V, H = 1024, 500
    i, j = np.arange(V)[:, None], np.arange(H)[None, :]
    W, b, c = 0.01 * np.cos(0.37 * i + 0.11 * j), 0.05 * np.sin(0.3 * np.arange(V)), 0.02 * np.cos(0.7 * np.arange(H))
    lay = _layer('rbm', V, H)
    lay.set_weights(W, b, c)
    F = _np(lay.free_energy(horses))
    np.testing.assert_allclose(F.mean(), -347.28000091799646, rtol=1e-5)
    np.testing.assert_allclose(F[0, 0], -348.6856458065582, rtol=1e-5)
    np.testing.assert_allclose(_np(lay.h_probs(horses)).sum(), 81967.46732612466, rtol=1e-5)
    gW = _np(lay.cd_gradient(horses, 1.0 - horses))[:V * H].reshape(V, H)
    np.testing.assert_allclose(np.linalg.norm(gW), 240.9141144337796, rtol=1e-4)
    np.testing.assert_allclose(gW[0, 0], 0.45235972577049083, rtol=1e-4)
    np.testing.assert_allclose(gW[512, 250], 0.48199651515066666, rtol=1e-4)
    zero = _layer('rbm', V, H)
    zero.set_weights(np.zeros((V, H)), np.zeros(V), np.zeros(H))
    np.testing.assert_allclose(_np(zero.free_energy(horses)), -H * np.log(2.0), rtol=1e-6)     # anchor (1)
    assert np.all(_np(zero.h_probs(horses)) == 0.5) and abs(float(zero.cd_loss(horses, 1 - horses))) < 1e-4


# ---- the fused tcgen05 chain ---------------------------------------------------------------------------------------
def _margin_uniforms(orc, v0, k, rs, margin=1e-4):
    """Uniform draws (fp32-representable) for a CD-k chain with |p - u| > margin at every step of the fp64
    oracle chain, so correct fp32 implementations must agree bit for bit (SURVEY.md §7 hard part 2)."""
    def fix(p, shape):
        u = rs.uniform(size=shape).astype(np.float32).astype(np.float64)
        near = np.abs(p - u) < margin
        u[near] = np.where(p[near] + 4 * margin < 1.0, p[near] + 4 * margin, p[near] - 4 * margin)
        return u.astype(np.float32).astype(np.float64)
    N = v0.shape[0]
    p0 = orc.h_probs(v0)
    U = [fix(p0, (N, orc.H))]
    h = orc.sample_h(p0, U[0])
    chain = []
    for _ in range(k):
        pv = orc.v_probs(h)
        uv = fix(pv, (N, orc.V))
        v = orc.sample_v(pv, uv)
        ph = orc.h_probs(v)
        uh = fix(ph, (N, orc.H))
        h = orc.sample_h(ph, uh)
        chain.append((uv, uh))
    flat = np.concatenate([U[0].ravel()] + [np.concatenate([a.ravel(), b.ravel()]) for a, b in chain])
    return U[0], chain, flat.astype(np.float32)


# the last case has 160 up-GEMM tiles on 148 SMs: its sparse last wave runs as 128-column half tiles (ragged edges)
@pytest.mark.parametrize('N,V,H,k', [(64, 64, 64, 1), (328, 1024, 496, 3), (200, 264, 136, 2), (2504, 136, 2040, 1),
                                     (328, 1024, 500, 2), (104, 76, 50, 1)])
def test_tc_chain_bit_exact_with_injected_uniforms(horses, N, V, H, k):
    from deepbelief_b200 import ops
    rs = np.random.RandomState(N + V + H)
    W = rs.normal(size=(V, H)) * 0.05
    b, c = rs.normal(size=V) * 0.1, rs.normal(size=H) * 0.1
    v0 = horses[:N, :V].astype(np.float64) if (N <= 328 and V <= 1024) else rs.binomial(1, 0.4, size=(N, V)).astype(np.float64)
    orc = orbm.RBMOracle(W, b, c)
    u0, chain, flat = _margin_uniforms(orc, v0, k, rs)
    ref = orbm.cd_k(orc, v0, k, u0, chain)
    params = torch.as_tensor(orc.pack().astype(np.float32)).cuda()
    plan = ops.CdTcPlan(N, V, H, k)
    plan.refresh_weights(params)
    grad = torch.zeros_like(params)
    vk, hk = torch.empty(N, V, device='cuda'), torch.empty(N, H, device='cuda')
    p0, pk = torch.empty(N, H, device='cuda'), torch.empty(N, H, device='cuda')
    plan.step(torch.as_tensor(v0.astype(np.float32)).cuda(), params, grad, injected=torch.as_tensor(flat).cuda(),
              vk_out=vk, hk_out=hk, p0_out=p0, pk_out=pk)
    np.testing.assert_array_equal(_np(vk), ref['vk'])                                    # bit-exact states
    np.testing.assert_array_equal(_np(hk), ref['hk'])
    np.testing.assert_allclose(_np(p0), ref['h0_probs'], rtol=PROB_RTOL)
    np.testing.assert_allclose(_np(pk), ref['hk_probs'], rtol=PROB_RTOL)
    got, want = _np(grad), ref['grad']
    for lo, hi in ((0, V * H), (V * H, V * H + V), (V * H + V, V * H + V + H)):
        assert _normrel(got[lo:hi], want[lo:hi]) < GRAD_NORMREL


def test_tc_chain_philox_mode_matches_oracle_philox():
    """Counter-based mode: the kernels' Philox draws are reproduced by oracle/philox.py; any mismatch must
    sit in a row that holds a near-tie somewhere along the chain."""
    from deepbelief_b200 import ops
    N, V, H, k, seed = 256, 512, 384, 2, 1234
    rs = np.random.RandomState(9)
    W, b, c = rs.normal(size=(V, H)) * 0.05, rs.normal(size=V) * 0.1, rs.normal(size=H) * 0.1
    v0 = rs.binomial(1, 0.4, size=(N, V)).astype(np.float64)
    orc = orbm.RBMOracle(W, b, c)
    u0 = philox.uniform(N, H, seed, 0).astype(np.float64)
    chain = [(philox.uniform(N, V, seed, 2 * s - 1).astype(np.float64), philox.uniform(N, H, seed, 2 * s).astype(np.float64))
             for s in range(1, k + 1)]
    ref = orbm.cd_k(orc, v0, k, u0, chain)
    # rows with a near-tie anywhere in the oracle chain
    tie = np.abs(ref['h0_probs'] - u0).min(axis=1) < 1e-6
    h = ref['h0']
    for uv, uh in chain:
        pv = orc.v_probs(h); tie |= np.abs(pv - uv).min(axis=1) < 1e-6
        v = orc.sample_v(pv, uv)
        ph = orc.h_probs(v); tie |= np.abs(ph - uh).min(axis=1) < 1e-6
        h = orc.sample_h(ph, uh)
    params = torch.as_tensor(orc.pack().astype(np.float32)).cuda()
    plan = ops.CdTcPlan(N, V, H, k)
    plan.refresh_weights(params)
    grad = torch.zeros_like(params)
    vk, hk = torch.empty(N, V, device='cuda'), torch.empty(N, H, device='cuda')
    rng = ops.Rng(seed=seed)
    plan.step(torch.as_tensor(v0.astype(np.float32)).cuda(), params, grad, rng=rng, vk_out=vk, hk_out=hk)
    assert rng.draw == 2 * k + 1
    bad_rows = np.where((_np(vk) != ref['vk']).any(axis=1) | (_np(hk) != ref['hk']).any(axis=1))[0]
    assert set(bad_rows) <= set(np.where(tie)[0]), 'mismatching rows %s hold no near-tie' % bad_rows
    # the SIMT sampler draws the same numbers (same counter convention)
    got = _np(ops.sample(torch.as_tensor(ref['h0_probs'].astype(np.float32)).cuda(), 1, rng=ops.Rng(seed=seed)))
    safe = np.abs(ref['h0_probs'] - u0) > 1e-6
    np.testing.assert_array_equal(got[safe], ref['h0'][safe])


def test_tc_row_sharding_is_invariant():
    """Two half-batches with row_offset reproduce the full batch (Philox keyed by global row)."""
    from deepbelief_b200 import ops
    N, V, H, k = 512, 256, 128, 2
    rs = np.random.RandomState(3)
    params = torch.as_tensor(np.concatenate([rs.normal(size=V * H) * 0.05, rs.normal(size=V + H) * 0.1]).astype(np.float32)).cuda()
    v0 = torch.as_tensor(rs.binomial(1, 0.4, size=(N, V)).astype(np.float32)).cuda()

    def run(rows, offset, n_global):
        plan = ops.CdTcPlan(rows.shape[0], V, H, k, n_global=n_global)
        plan.refresh_weights(params)
        grad = torch.zeros_like(params)
        vk = torch.empty(rows.shape[0], V, device='cuda')
        plan.step(rows.contiguous(), params, grad, rng=ops.Rng(seed=5, row_offset=offset), vk_out=vk)
        return vk, grad
    vk_full, g_full = run(v0, 0, N)
    vk_a, g_a = run(v0[:N // 2], 0, N)
    vk_b, g_b = run(v0[N // 2:], N // 2, N)
    assert torch.equal(torch.cat([vk_a, vk_b]), vk_full)
    assert _normrel(_np(g_a + g_b), _np(g_full)) < 1e-5        # sum of shard gradients == full-batch gradient


def test_training_loop_simt_and_tc_agree_with_oracle(horses, tmp_path):
    """RBM.train on the bundled horses data, CD-1, SGD: 3 iterations reproduce the oracle's parameters when the
    same Philox draws are used (general path, H=500) and the checkpoint round-trips. (SGD, because Adam's
    first steps are +-lr * sign(g) and amplify a sign flip of a ~0 gradient entry; Adam itself is checked
    element-wise in test_optimizer_kernels_vs_oracle.)"""
    import os
    from conftest import DATASETS
    from deepbelief_b200 import compat as tf
    from deepbelief_b200.layers.data import Data
    from deepbelief_b200.layers.rbm import RBM
    path = os.path.join(DATASETS, 'weizmann_horses_training_328x1024_binary.h5')
    test_path = os.path.join(DATASETS, 'weizmann_horses_eslami_test_14x1024_binary.h5')
    V, H, seed, iters = 1024, 500, 77, 3
    rs = np.random.RandomState(0)
    W0 = (rs.normal(size=(V, H)) * 0.01).astype(np.float32)
    lay = RBM(V, H, name='BBRBM-test', seed=seed)
    lay.set_weights(W0, np.zeros(V), np.zeros(H))
    lay.train(tf.train.GradientDescentOptimizer(learning_rate=0.05), Data(path, batch_size=328, log_epochs=False),
              Data(test_path, batch_size=8, log_epochs=False), num_gibbs_steps=1, pcd=False, max_iterations=iters,
              eval_interval=2, ckpt_interval=iters, ckpt_dir=str(tmp_path))
    # oracle replay: draws 0 (h0), 1 (v1), 2 (h1) per iteration; eval steps consume draws too, so replay the counter
    theta = np.concatenate([W0.ravel().astype(np.float64), np.zeros(V + H)])
    sgd = orbm.SGD(0.05)
    draw = 0
    v0 = horses.astype(np.float64)
    for it in range(1, iters + 1):
        if it == 1:
            draw += 2                        # evaluation before training: sample_h + downprop sample_v (rbm.py:684-686)
        orc = orbm.RBMOracle(theta[:V * H].reshape(V, H), theta[V * H:V * H + V], theta[V * H + V:])
        u0 = philox.uniform(328, H, seed, draw).astype(np.float64)
        uv = philox.uniform(328, V, seed, draw + 1).astype(np.float64)
        uh = philox.uniform(328, H, seed, draw + 2).astype(np.float64)
        draw += 3
        out = orbm.cd_k(orc, v0, 1, u0, [(uv, uh)])
        theta = sgd.step(theta, out['grad'])
        if it % 2 == 0:
            draw += 2
    got = _np(lay._params)
    assert np.max(np.abs(got - theta)) < 1e-6
    assert len(lay.loss_y_vals) == 2 and all(np.isfinite(lay.loss_y_vals))
    ck = os.path.join(str(tmp_path), 'BBRBM-test-%d' % iters)
    assert os.path.exists(ck)
    other = RBM(V, H, name='BBRBM-test')
    other.restore(ck)
    assert torch.equal(other._params, lay._params) and other.checkpoint_iter == 1 + iters


@pytest.mark.parametrize('kind', ['sgd', 'momentum', 'adam'])
def test_optimizer_kernels_vs_oracle(kind):
    from deepbelief_b200 import compat as tf, ops
    rs = np.random.RandomState(4)
    n = 10007
    theta0 = rs.normal(size=n)
    grads = [rs.normal(size=n) * 0.1 for _ in range(4)]
    if kind == 'sgd':
        opt, orc = tf.train.GradientDescentOptimizer(0.1), orbm.SGD(0.1)
    elif kind == 'momentum':
        opt, orc = tf.train.MomentumOptimizer(1e-3, momentum=0.9, weight_decay=1e-4), orbm.Momentum(1e-3, 0.9, 1e-4)
    else:
        opt, orc = tf.train.AdamOptimizer(1e-3), orbm.Adam(1e-3)
    p = torch.as_tensor(theta0.astype(np.float32)).cuda()
    state = torch.zeros(max(opt.state_multiple(), 1) * n, device='cuda')
    th = theta0.astype(np.float32).astype(np.float64)
    for g in grads:
        g32 = g.astype(np.float32)
        ops.optimizer_step(opt.descriptor(), p, torch.as_tensor(g32).cuda(), state)
        th = orc.step(th, g32.astype(np.float64))
    np.testing.assert_allclose(_np(p), th, rtol=2e-6, atol=2e-7)


def test_full_size_config5_properties():
    """BASELINE config 5 size (V=4096, H=8192, N=16384) through size-independent properties: with W = 0 the
    chain factorises, so dW[v,h] == db[v] * sigmoid(c[h]) exactly, dc == 0, and the sample means follow
    sigmoid(b) / sigmoid(c)."""
    from deepbelief_b200 import ops
    N, V, H, k = 16384, 4096, 8192, 1
    rs = np.random.RandomState(1)
    b = rs.uniform(-2, 2, size=V)
    c = rs.uniform(-2, 2, size=H)
    params = torch.as_tensor(np.concatenate([np.zeros(V * H), b, c]).astype(np.float32)).cuda()
    v0 = (torch.rand(N, V, device='cuda', generator=torch.Generator(device='cuda').manual_seed(0)) < 0.3).float()
    plan = ops.CdTcPlan(N, V, H, k)
    plan.refresh_weights(params)
    grad = torch.zeros_like(params)
    vk, hk = torch.empty(N, V, device='cuda'), torch.empty(N, H, device='cuda')
    plan.step(v0, params, grad, rng=ops.Rng(seed=2), vk_out=vk, hk_out=hk)
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))  # noqa: E731
    vmean, hmean = _np(vk.mean(0)), _np(hk.mean(0))
    assert np.max(np.abs(vmean - sig(b))) < 6 * 0.5 / np.sqrt(N)
    assert np.max(np.abs(hmean - sig(c))) < 6 * 0.5 / np.sqrt(N)
    assert torch.all((vk == 0) | (vk == 1)) and torch.all((hk == 0) | (hk == 1))
    g = _np(grad)
    gW, gb, gc = g[:V * H].reshape(V, H), g[V * H:V * H + V], g[V * H + V:]
    np.testing.assert_allclose(gb, -(_np(v0.mean(0)) - vmean), rtol=1e-4, atol=1e-7)
    assert np.max(np.abs(gc)) < 1e-6
    # the dW GEMM runs K = 2N = 32768 deep; the tensor core's fp32 accumulate truncates, so the K axis is cut
    # into <= 4096-column chunks summed in fp32 (csrc/cd_chain.cu): 1.2e-4 unsplit, bound here 3e-5
    assert _normrel(gW, gb[:, None] * sig(c.astype(np.float32).astype(np.float64))[None, :]) < 3e-5


def test_full_size_config5_sampled_rows_vs_fp64_oracle():
    """BASELINE config 5 at its own size (N=16384, V=4096, H=8192) with W ~ N(0, 0.01^2) — the accumulation depth
    (K = 4096 / 8192 chain, K = 2N = 32768 dW), the 2^e weight scaling and the fp16 hi/lo split the metric runs on.
    Rows of the chain are independent given (W, b, c) (rbm.py:628-646), so a SAMPLED set of 128 rows is followed
    through the whole CD-1 chain by the fp64 oracle with margin uniforms injected for exactly those rows (every other
    row draws plain random uniforms): h0 probabilities 1e-5, v1 / h1 states bit-exact, hk probabilities 1e-5. dW is
    checked on a sampled 64 x 64 block (4096 entries) computed in fp64 from the chain ends of ALL 16384 rows
    (K = 32768 deep) at 1e-4 of max|dW|, and db / dc in full."""
    from deepbelief_b200 import ops
    N, V, H, k = 16384, 4096, 8192, 1
    R = 128
    rs = np.random.RandomState(55)
    W = (rs.normal(size=(V, H)) * 0.01).astype(np.float32)
    b = (rs.normal(size=V) * 0.1).astype(np.float32)
    c = (rs.normal(size=H) * 0.1).astype(np.float32)
    gen = torch.Generator(device='cuda').manual_seed(0)
    v0 = (torch.rand(N, V, device='cuda', generator=gen) < 0.5).float()
    rows = np.sort(rs.choice(N, size=R, replace=False))
    rows_t = torch.as_tensor(rows, device='cuda')
    v0_rows = v0[rows_t].cpu().numpy().astype(np.float64)
    orc = orbm.RBMOracle(W, b, c)                                   # fp64 ground truth of the sampled rows
    u0, chain, _ = _margin_uniforms(orc, v0_rows, k, rs)
    ref = orbm.cd_k(orc, v0_rows, k, u0, chain)
    # injected draws for the whole batch: [U_h0 | U_v1 | U_h1], sampled rows overwritten with the margin uniforms
    NH, NV = N * H, N * V
    inj = torch.rand(NH + k * (NV + NH), device='cuda', generator=gen)
    inj[:NH].view(N, H)[rows_t] = torch.as_tensor(u0.astype(np.float32)).cuda()
    inj[NH:NH + NV].view(N, V)[rows_t] = torch.as_tensor(chain[0][0].astype(np.float32)).cuda()
    inj[NH + NV:].view(N, H)[rows_t] = torch.as_tensor(chain[0][1].astype(np.float32)).cuda()
    params = torch.as_tensor(np.concatenate([W.ravel(), b, c])).cuda()
    plan = ops.CdTcPlan(N, V, H, k)
    plan.refresh_weights(params)
    grad = torch.zeros_like(params)
    vk, hk = torch.empty(N, V, device='cuda'), torch.empty(N, H, device='cuda')
    p0, pk = torch.empty(N, H, device='cuda'), torch.empty(N, H, device='cuda')
    plan.step(v0, params, grad, injected=inj, vk_out=vk, hk_out=hk, p0_out=p0, pk_out=pk)
    torch.cuda.synchronize()
    del inj
    # ---- the sampled rows against the oracle ----
    np.testing.assert_allclose(_np(p0[rows_t]), ref['h0_probs'], rtol=PROB_RTOL)
    np.testing.assert_array_equal(_np(vk[rows_t]), ref['vk'])            # bit-exact states (margin 1e-4)
    np.testing.assert_array_equal(_np(hk[rows_t]), ref['hk'])
    np.testing.assert_allclose(_np(pk[rows_t]), ref['hk_probs'], rtol=PROB_RTOL)
    assert torch.all((vk == 0) | (vk == 1)) and torch.all((hk == 0) | (hk == 1))
    # ---- gradient: sampled dW block in fp64 over all N rows (chain ends taken from the device, verified above on
    # the sampled rows), db and dc in full ----
    vs = np.sort(rs.choice(V, size=64, replace=False))
    hs = np.sort(rs.choice(H, size=64, replace=False))
    vs_t, hs_t = torch.as_tensor(vs, device='cuda'), torch.as_tensor(hs, device='cuda')
    v0s, vks = _np(v0[:, vs_t]), _np(vk[:, vs_t])
    p0s, pks = _np(p0[:, hs_t]), _np(pk[:, hs_t])
    want_dW = -(v0s.T @ p0s - vks.T @ pks) / N
    g = grad
    got_dW = _np(g[:V * H].view(V, H)[vs_t][:, hs_t])
    scale = float(g[:V * H].abs().max())
    err_dW = np.max(np.abs(got_dW - want_dW)) / scale
    assert err_dW < GRAD_NORMREL, err_dW
    want_db = -(_np(v0.double().mean(0)) - _np(vk.double().mean(0)))
    want_dc = -(_np(p0.double().mean(0)) - _np(pk.double().mean(0)))
    assert _normrel(_np(g[V * H:V * H + V]), want_db) < GRAD_NORMREL
    assert _normrel(_np(g[V * H + V:]), want_dc) < GRAD_NORMREL
    print('config-5 parity: p0 %.2e pk %.2e dW(4096 entries) %.2e of max|dW| %.3g' % (
        np.max(np.abs(_np(p0[rows_t]) - ref['h0_probs']) / ref['h0_probs']),
        np.max(np.abs(_np(pk[rows_t]) - ref['hk_probs']) / ref['hk_probs']), err_dW, scale))


@pytest.mark.parametrize('cname', ['concrete', 'prob'])
def test_backprop_gradients_vs_oracle(cname):
    """RBM.backprop's loss and gradient (rbm.py:739-758) on a two-layer stack against the reference-pinned oracle."""
    from conftest import Golden
    from deepbelief_b200.layers.rbm import RBM
    g = Golden('backprop_golden.npz')
    T, sample = {'concrete': (0.25, True), 'prob': (None, False)}[cname]
    P = {k: g['backprop/P/' + k] for k in ('W1', 'b1', 'c1', 'W2', 'b2', 'c2')}
    v = g['backprop/v']
    V, H1 = P['W1'].shape
    H2 = P['W2'].shape[1]
    k = int(g['backprop/k']) if (T is not None or sample) else 1
    draws = [g['backprop/draw%d' % i] for i in range(4 * int(g['backprop/k']))]
    l1 = RBM(V, H1, temperature=T, name='bp1')
    l1.set_weights(P['W1'], P['b1'], P['c1'])
    l2 = RBM(H1, H2, temperature=T, bottom=l1, name='bp2')
    l2.set_weights(P['W2'], P['b2'], P['c2'])
    loss, grads, v1 = l2.backprop_gradients(v, k=k, sample=sample, draws=draws if T is not None else None)
    layers = [dict(W=P['W1'], b=P['b1'], c=P['c1'], temperature=T), dict(W=P['W2'], b=P['b2'], c=P['c2'], temperature=T)]
    oloss, ograds, ov1 = orbm.propagate_loss_and_grads(layers, v, k, draws, sample=sample)
    # saturated Concrete samples are clipped at 1 - 1e-6, where fp32 resolves 1 - p to ~6 %: the budget is the error of
    # the reference formulas evaluated in fp32 (SURVEY.md §7.3 rule)
    oloss32 = orbm.propagate_loss_and_grads(layers, v, k, draws, sample=sample, dtype=np.float32)[0]
    budget = max(1e-5, 2.0 * abs(oloss32 - oloss) / abs(oloss))
    np.testing.assert_allclose(float(loss), oloss, rtol=budget)
    np.testing.assert_allclose(float(loss), float(g['backprop/%s/loss' % cname]), rtol=budget)      # the reference itself
    assert _normrel(_np(v1), ov1) < 1e-4
    for gl, og, (nv, nh) in zip(grads, ograds, ((V, H1), (H1, H2))):
        got = _np(gl)
        assert _normrel(got[:nv * nh].reshape(nv, nh), og['W']) < GRAD_NORMREL
        assert _normrel(got[nv * nh:nv * nh + nv], og['b']) < GRAD_NORMREL
        assert _normrel(got[nv * nh + nv:], og['c']) < GRAD_NORMREL


def test_backprop_loop_reduces_reconstruction_loss(horses):
    """RBM.backprop on horses (probabilities propagated, SGD): the test cross entropy falls."""
    import os
    from conftest import DATASETS
    from deepbelief_b200.layers.data import Data
    from deepbelief_b200.layers.rbm import RBM
    np.random.seed(0)
    lay = RBM(1024, 64, W_std=0.01, name='bp-horses', seed=3)
    path = os.path.join(DATASETS, 'weizmann_horses_training_328x1024_binary.h5')
    lay.backprop(Data(path, batch_size=328, log_epochs=False), Data(path, batch_size=328, log_epochs=False),
                 training_sampling=False, learning_rate=0.05, max_iterations=60, eval_interval=30)
    assert len(lay.loss_y_vals) == 3 and lay.loss_y_vals[-1] < lay.loss_y_vals[0]


@pytest.mark.parametrize('case', ['rbm_adam_eval', 'gbrbm_momentum_pcd', 'stack_concrete', 'rbm_tensor_chain'])
def test_training_iteration_as_cuda_graph_is_bit_identical(horses, monkeypatch, case):
    """RBM.train replays one captured CUDA graph per iteration (CdGraphStep: draw ids and the Adam step number come from
    a device clock): the parameters after the loop must be bit-identical to the eager loop — Adam with evaluation steps
    in between (which consume draws eagerly), Momentum + weight decay with PCD on a Gaussian-visible layer, a Concrete
    layer trained on top of a re-sampled bottom layer, and the tcgen05 chain (narrow 64-row tiles at this size, Philox
    draws keyed by the device clock, db_cd_tc_apply_update with Adam's step size formed on the device)."""
    from deepbelief_b200 import compat as tf
    from deepbelief_b200.layers.data import Data
    from deepbelief_b200.layers.rbm import RBM, CdGraphStep
    from deepbelief_b200.layers.gbrbm import GBRBM
    rs = np.random.RandomState(5)
    X = horses[:300].astype(np.float32)
    test = horses[300:328].astype(np.float32)

    def run(graph):
        monkeypatch.setenv('DEEPBELIEF_B200_GRAPH', '1' if graph else '0')
        monkeypatch.setattr(RBM, 'GRAPH_MIN_ITERATIONS', 8)          # the loops here are 17 iterations long
        np.random.seed(0)                           # Data reshuffles with the global numpy stream on every epoch wrap
        rs2 = np.random.RandomState(11)
        kw = {}
        if case == 'rbm_tensor_chain':
            monkeypatch.setattr(RBM, 'TC_MIN_WORK', 0)
            lay = RBM(1024, 104, name='g3', seed=8)
            lay.set_weights((rs2.normal(size=(1024, 104)) * 0.01).astype(np.float32), np.zeros(1024), np.zeros(104))
            opt, data = tf.train.AdamOptimizer(learning_rate=1e-3), X[:296]
            kw = dict(num_gibbs_steps=2, pcd=False, eval_interval=5)
        elif case == 'rbm_adam_eval':
            lay = RBM(1024, 100, name='g0', seed=3)
            lay.set_weights((rs2.normal(size=(1024, 100)) * 0.01).astype(np.float32), np.zeros(1024), np.zeros(100))
            opt, data = tf.train.AdamOptimizer(learning_rate=1e-3), X
            kw = dict(num_gibbs_steps=2, pcd=False, eval_interval=5)
        elif case == 'gbrbm_momentum_pcd':
            lay = GBRBM(1024, 60, name='g1', seed=4)
            lay.set_weights((rs2.normal(size=(1024, 60)) * 0.001).astype(np.float32), np.zeros(1024), np.zeros(60))
            opt = tf.train.MomentumOptimizer(learning_rate=1e-3, momentum=0.9, weight_decay=1e-4)
            data = ((X - X.mean(0)) / (X.std(0) + 1e-3)).astype(np.float32)
            kw = dict(num_gibbs_steps=1, pcd=True)
        else:
            bottom = RBM(1024, 64, temperature=0.1, name='g2b', seed=5)
            bottom.set_weights((rs2.normal(size=(1024, 64)) * 0.05).astype(np.float32), np.zeros(1024), np.zeros(64))
            lay = RBM(64, 32, temperature=0.1, bottom=bottom, name='g2', seed=6)
            lay.set_weights((rs2.normal(size=(64, 32)) * 0.05).astype(np.float32), np.zeros(64), np.zeros(32))
            opt, data = tf.train.AdamOptimizer(learning_rate=1e-3), X
            kw = dict(num_gibbs_steps=3, pcd=False)
        created = []
        orig = CdGraphStep.__init__

        def spy(self, *a, **k):
            created.append(1)
            orig(self, *a, **k)
        monkeypatch.setattr(CdGraphStep, '__init__', spy)
        lay.train(opt, Data.from_arrays(data, batch_size=104 if case == 'rbm_tensor_chain' else 100, log_epochs=False),
                  Data.from_arrays(test, batch_size=14, log_epochs=False), max_iterations=17, **kw)
        monkeypatch.setattr(CdGraphStep, '__init__', orig)
        assert bool(created) == graph
        assert (lay._tc_plan is not None) == (case == 'rbm_tensor_chain')
        return _np(lay._params), lay.rng.draw, opt.step_count

    p_graph, d_graph, s_graph = run(True)
    p_eager, d_eager, s_eager = run(False)
    assert (d_graph, s_graph) == (d_eager, s_eager)
    assert np.array_equal(p_graph, p_eager)
    assert np.isfinite(p_graph).all() and np.abs(p_graph).max() > 0


@pytest.mark.parametrize('M,N,K,trans', [(328, 500, 1024, False), (328, 1024, 500, True), (70, 130, 1000, False),
                                         (100, 60, 1024, False), (33, 64, 257, True), (5000, 100, 200, False)])
def test_layer_forward_cluster_split_k(M, N, K, trans):
    """Launches with few output tiles run as thread-block clusters that split K and reduce through distributed shared
    memory (simt_gemm.cuh): every epilogue (scale, bias, sigmoid, Bernoulli with injected uniforms, Gaussian noise,
    a_div prologue) against a float64 evaluation, ragged edges included, and bit-identical on repetition."""
    from deepbelief_b200 import ops, _lib as L
    rs = np.random.RandomState(M + N + K)
    A = rs.rand(M, K).astype(np.float32)
    W = (rs.normal(size=(N, K) if trans else (K, N)) * 0.05).astype(np.float32)
    bias = rs.normal(size=N).astype(np.float32) * 0.1
    scale = (0.5 + rs.rand(1, N)).astype(np.float32)
    div = (0.5 + rs.rand(1, K)).astype(np.float32)
    cu = lambda a: torch.as_tensor(a).cuda()  # noqa: E731
    x = (A.astype(np.float64) / div) @ (W.T if trans else W).astype(np.float64) * scale + bias
    p = 1.0 / (1.0 + np.exp(-x))
    # margin-guaranteed uniforms: |p - u| > 1e-3
    u = rs.rand(M, N)
    u = np.where(np.abs(p - u) < 1e-3, np.clip(p + 0.01, 0, 1), u).astype(np.float32)
    act, s = ops.layer_forward(cu(A), cu(W), cu(bias), trans, L.ACT_SIGMOID, L.SAMPLE_BERNOULLI, a_div=cu(div),
                               col_scale=cu(scale), injected=cu(u))
    np.testing.assert_allclose(_np(act), p, rtol=PROB_RTOL, atol=1e-7)
    np.testing.assert_array_equal(_np(s), (p > u).astype(np.float64))
    act2, s2 = ops.layer_forward(cu(A), cu(W), cu(bias), trans, L.ACT_SIGMOID, L.SAMPLE_BERNOULLI, a_div=cu(div),
                                 col_scale=cu(scale), injected=cu(u))
    assert torch.equal(act, act2) and torch.equal(s, s2)
    eps = rs.normal(size=(M, N)).astype(np.float32)
    lin, g = ops.layer_forward(cu(A), cu(W), cu(bias), trans, L.ACT_LINEAR, L.SAMPLE_GAUSSIAN, a_div=cu(div),
                               col_scale=cu(scale), noise_scale=cu(scale), injected=cu(eps))
    assert _normrel(_np(lin), x) < 2e-6
    assert _normrel(_np(g), x + scale * eps) < 2e-6


@pytest.mark.parametrize('kind', ['rbm', 'gbrbm'])
def test_cd_gradient_from_chain_probabilities_is_bit_identical(horses, kind):
    """db_cd_gradients_hidden takes the hidden probabilities the chain sampled from instead of repeating the two
    free-energy GEMMs (rbm.py:388-390): same bits as db_cd_gradients, and both within tolerance of the oracle."""
    from deepbelief_b200 import ops
    rs = np.random.RandomState(3)
    V, H = 1024, 500
    lay = _layer(kind, V, H, name='q' + kind)
    W = (rs.normal(size=(V, H)) * 0.02).astype(np.float32)
    b, c = rs.normal(size=V).astype(np.float32) * 0.1, rs.normal(size=H).astype(np.float32) * 0.1
    lay.set_weights(W, b, c)
    v0 = horses.astype(np.float32) if kind == 'rbm' else rs.normal(size=(328, V)).astype(np.float32)
    vk = (rs.rand(328, V) < 0.4).astype(np.float32) if kind == 'rbm' else rs.normal(size=(328, V)).astype(np.float32)
    p0, pk = lay.h_probs(v0), lay.h_probs(vk)
    g_plain = lay.cd_gradient(v0, vk)
    g_hidden = ops.cd_gradients(lay.layer_kind, torch.as_tensor(v0).cuda(), torch.as_tensor(vk).cuda(), lay._params, V, H,
                                hidden_probs=(p0, pk))
    assert torch.equal(g_plain, g_hidden)
    orc = orbm.RBMOracle(W, b, c, kind=kind)
    want = orc.cd_grad_packed(v0.astype(np.float64), vk.astype(np.float64))
    got = _np(g_hidden)
    for lo, hi in ((0, V * H), (V * H, V * H + V), (V * H + V, V * H + V + H)):
        assert _normrel(got[lo:hi], want[lo:hi]) < GRAD_NORMREL


def test_bgrbmsv_shared_sigma(horses):
    """BGRBMSV (bgrbmsv.py): one shared sigma_h = softplus(1) at construction, forward / gradients equal to the BGRBM oracle
    with that value in every unit, the shared scalar's gradient the sum of the per-unit gradients, and training keeps the
    H packed entries equal; a fed sigma_h (the placeholder form) is used as given and not trained."""
    from deepbelief_b200 import compat as tf
    from deepbelief_b200.layers.bgrbmsv import BGRBMSV
    from deepbelief_b200.layers.data import Data
    rs = np.random.RandomState(9)
    V, H, N = 1024, 40, 96
    lay = BGRBMSV(V, H, name='sv', seed=2)
    W = (rs.normal(size=(V, H)) * 0.01).astype(np.float32)
    b, c = rs.normal(size=V).astype(np.float32) * 0.05, rs.normal(size=H).astype(np.float32) * 0.05
    lay.set_weights(W, b, c)
    sig = np.log1p(np.exp(1.0))
    np.testing.assert_allclose(_np(lay.sigma_h), np.full((1, H), sig), rtol=1e-6)
    assert lay.params['sigma_h']().shape == (1,)
    orc = orbm.RBMOracle(W, b, c, kind='bgrbm', opt_sigma_h=np.full((1, H), 1.0))
    v0 = horses[:N].astype(np.float64)
    vk = (rs.rand(N, V) < 0.4).astype(np.float64)
    np.testing.assert_allclose(_np(lay.h_probs(v0)), orc.h_probs(v0), rtol=PROB_RTOL, atol=1e-6)
    want = orc.cd_grad_packed(v0, vk)
    grad = lay.cd_gradient(v0, vk)
    off = V * H + V + H
    assert _normrel(_np(grad)[:off], want[:off]) < GRAD_NORMREL
    _, shared = lay._trainable_pair(grad)
    got_sigma = _np(shared)[off:]
    assert np.all(got_sigma == got_sigma[0]) and abs(got_sigma[0] - want[off:].sum()) < GRAD_NORMREL * np.abs(want[off:]).sum()
    lay.train(tf.train.AdamOptimizer(learning_rate=1e-3), Data.from_arrays(horses[:N].astype(np.float32), batch_size=N, log_epochs=False),
              None, num_gibbs_steps=1, pcd=False, max_iterations=6)
    os_ = _np(lay.opt_sigma_h).ravel()
    assert np.all(os_ == os_[0]) and os_[0] != 1.0 and np.isfinite(os_[0])
    fed = BGRBMSV(V, H, sigma_h_ph=0.7, name='svfed', seed=2)
    fed.set_weights(W, b, c)
    np.testing.assert_allclose(_np(fed.sigma_h), np.full((1, H), 0.7), rtol=1e-6)
    fed.train(tf.train.GradientDescentOptimizer(learning_rate=1e-3), Data.from_arrays(horses[:N].astype(np.float32), batch_size=N, log_epochs=False),
              None, num_gibbs_steps=1, pcd=False, max_iterations=3)
    np.testing.assert_allclose(_np(fed.sigma_h), np.full((1, H), 0.7), rtol=1e-6)
    assert not np.array_equal(_np(fed.W), W)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')
@pytest.mark.parametrize('kind', ['sgd', 'momentum', 'adam'])
def test_fused_update_is_bit_identical_to_step_plus_optimizer(kind):
    """North-star K2 (db_cd_tc_step_update): the dW kernel whose epilogue applies the optimizer rule and rewrites the
    split-weight cache, and the bias kernel that updates b | c, against db_cd_tc_step + db_optimizer_step +
    db_cd_tc_refresh_weights (rbm.py:664-670) on the same Philox draws: parameters, optimizer slots, packed gradient
    and the fp16 cache bytes must agree BIT FOR BIT over several steps. Ragged tile edges (V, H, N not multiples of the
    128 x 256 tile / 64-row block) and a learning rate that moves max|W| across powers of two (the re-scale path)."""
    from deepbelief_b200 import compat as tf, ops
    N, V, H, k = 520, 2056, 1416, 2
    rs = np.random.RandomState(9)
    theta0 = np.concatenate([(rs.normal(size=V * H) * 0.01), rs.normal(size=V) * 0.1, rs.normal(size=H) * 0.1]).astype(np.float32)
    v0 = torch.as_tensor((rs.uniform(size=(N, V)) < 0.4).astype(np.float32)).cuda()

    def make():
        if kind == 'sgd':
            return tf.train.GradientDescentOptimizer(3.0)
        if kind == 'momentum':
            return tf.train.MomentumOptimizer(1.0, momentum=0.9, weight_decay=1e-4)
        return tf.train.AdamOptimizer(2e-2)

    runs = []
    for fused in (False, True):
        opt = make()
        p = torch.as_tensor(theta0).cuda().clone()
        state = torch.zeros(max(opt.state_multiple(), 1) * p.numel(), device='cuda')
        grad = torch.zeros_like(p)
        plan = ops.CdTcPlan(N, V, H, k)
        assert plan.update_fusable()
        plan.refresh_weights(p)
        rng = ops.Rng(seed=21)
        scale0 = float(plan.cache[:4].clone().view(torch.float32)[0])       # the scale chosen from the initial max|W|
        snaps = []
        for it in range(4):
            if fused:
                plan.step_update(v0, p, grad, opt.descriptor(), state, rng=rng)
            else:
                plan.step(v0, p, grad, rng=rng)
                ops.optimizer_step(opt.descriptor(), p, grad, state)
                plan.refresh_weights(p)
            torch.cuda.synchronize()
            mat = V * H * 2                                   # the four fp16 matrices sit at 256-byte aligned offsets behind the header
            stride = (mat + 255) // 256 * 256
            cache = torch.cat([plan.cache[256 + i * stride:256 + i * stride + mat] for i in range(4)])
            snaps.append((p.clone(), state.clone(), grad.clone(), cache, plan.cache[:8].clone()))
        runs.append(snaps)
    scales = {scale0}
    for it, (a, b) in enumerate(zip(*runs)):
        for name, x, y in zip(('params', 'optimizer state', 'gradient', 'split-weight cache', 'cache scale'), a, b):
            assert torch.equal(x, y), 'step %d: %s differs between the fused and the separate update' % (it, name)
        scales.add(float(a[4].view(torch.float32)[0]))
    assert np.all(np.isfinite(_np(runs[1][-1][0])))
    if kind != 'adam':
        assert len(scales) > 1, 'the test is meant to cross a power of two in max|W| (re-scale path)'


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')
@pytest.mark.parametrize('V,H', [(520, 392), (1024, 500), (76, 50)])
def test_apply_update_and_byte_batches_are_bit_identical(V, H):
    """The data-parallel tail db_cd_tc_apply_update (optimizer rule + split-weight cache rewrite in one pass, run after the
    gradient all-reduce) against db_optimizer_step + db_cd_tc_refresh_weights, and a uint8 {0,1} batch (what the feeder
    stages for binary data sets) against the same batch in fp32: identical bits in every output."""
    from deepbelief_b200 import compat as tf, ops
    N, k = 264, 1                                     # not fusable into the dW epilogue (few tiles): the stand-alone kernel's case;
    # widths that are not multiples of 8 (the reference's 500-unit layer): padded row pitches of the internal fp16 buffers
    rs = np.random.RandomState(3)
    theta0 = np.concatenate([(rs.normal(size=V * H) * 0.05), rs.normal(size=V) * 0.1, rs.normal(size=H) * 0.1]).astype(np.float32)
    v_f32 = torch.as_tensor((rs.uniform(size=(N, V)) < 0.3).astype(np.float32)).cuda()
    v_u8 = v_f32.to(torch.uint8)
    out = []
    for mode in ('separate', 'apply'):
        opt = tf.train.MomentumOptimizer(0.5, momentum=0.9, weight_decay=1e-4)
        p = torch.as_tensor(theta0).cuda().clone()
        state = torch.zeros(opt.state_multiple() * p.numel(), device='cuda')
        grad = torch.zeros_like(p)
        plan = ops.CdTcPlan(N, V, H, k)
        plan.refresh_weights(p)
        rng = ops.Rng(seed=5)
        for it in range(3):
            plan.step(v_u8 if (mode == 'apply' and it % 2 == 0) else v_f32, p, grad, rng=rng)
            if mode == 'apply':
                plan.apply_update(p, grad, opt.descriptor(), state)
            else:
                ops.optimizer_step(opt.descriptor(), p, grad, state)
                plan.refresh_weights(p)
        torch.cuda.synchronize()
        Hp, Vp = (H + 7) // 8 * 8, (V + 7) // 8 * 8
        sizes = [V * Hp * 2, V * Hp * 2, H * Vp * 2, H * Vp * 2]          # W hi, W lo [V, Hp]; W^T hi, lo [H, Vp]
        off, parts = 256, []
        for i, nbytes in enumerate(sizes):
            m = plan.cache[off:off + nbytes].view(torch.float16).view(V if i < 2 else H, Hp if i < 2 else Vp)
            parts.append(m[:, :H if i < 2 else V].contiguous().view(-1))    # the padding columns are never written
            off += (nbytes + 255) // 256 * 256
        out.append((p.clone(), state.clone(), grad.clone(), plan.cache[:8].clone(), torch.cat(parts)))
    for name, x, y in zip(('params', 'optimizer state', 'gradient', 'cache scale', 'split-weight cache'), *out):
        assert torch.equal(x, y), name
