"""The torch-CPU form of the oracle (oracle/torch_port.py: bench.py's CPU baseline) against the numpy
oracle it restates, on injected draws."""
import numpy as np
import torch

from oracle import gp as ogp, rbm as orbm, torch_port


def test_cd_step_matches_numpy_oracle():
    rs = np.random.RandomState(0)
    N, V, H, k = 24, 20, 12, 3
    W, b, c = rs.normal(size=(V, H)) * 0.3, rs.normal(size=V) * 0.1, rs.normal(size=H) * 0.1
    v0 = rs.binomial(1, 0.5, size=(N, V)).astype(np.float64)
    u0 = rs.uniform(size=(N, H))
    chain = [(rs.uniform(size=(N, V)), rs.uniform(size=(N, H))) for _ in range(k)]
    m = orbm.RBMOracle(W, b, c)
    adam = orbm.Adam(1e-3)
    theta = m.pack()
    port = torch_port.CdStepTorch(W, b, c, lr=1e-3, dtype=torch.float64)
    flat = [u0] + [d for pair in chain for d in pair]
    for _ in range(3):
        m = orbm.RBMOracle(theta[:V * H].reshape(V, H), theta[V * H:V * H + V], theta[V * H + V:])
        out = orbm.cd_k(m, v0, k, u0, chain)
        theta = adam.step(theta, out['grad'])
        gW, gb, gc = port.step(v0, k, draws=flat)
        np.testing.assert_allclose(np.concatenate([gW.numpy().ravel(), gb.numpy().ravel(), gc.numpy().ravel()]), out['grad'],
                                   rtol=1e-10, atol=1e-13)
    got = np.concatenate([port.W.numpy().ravel(), port.b.numpy().ravel(), port.c.numpy().ravel()])
    np.testing.assert_allclose(got, theta, rtol=1e-10, atol=1e-12)


def test_gplvm_autograd_matches_closed_form_oracle():
    rs = np.random.RandomState(1)
    N, D, Q = 60, 9, 2
    X = rs.normal(size=(N, Q))
    Y = rs.normal(size=(N, D)); Y -= Y.mean(0)
    oa, og, on = 0.3, 0.7, -0.2
    o = ogp.GPLVMOracle(Y, X, opt_alpha=oa, opt_gamma=og, opt_noise=on, jitter=1e-6)
    dX, da, dg, dn = o.grad()
    loss, tX, ta, tg, tn = torch_port.gplvm_loss_and_grad(X, Y, oa, og, on, jitter=1e-6, dtype=torch.float64)
    np.testing.assert_allclose(loss, float(o.loss()[0]), rtol=1e-12)
    np.testing.assert_allclose(tX.numpy(), dX, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose([ta, tg, tn], [da, dg, dn], rtol=1e-8)
