"""torch-CPU form of the oracle for the two benchmarked steps — TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).

SURVEY.md section 8(d) "CPU side-by-side": TensorFlow 1.8 cannot run here, so the CPU baseline is the fp32 restatement
of the reference graph expressed with torch CPU ops — `torch.matmul` (MKL / oneDNN sgemm) for the CD chain, and
`torch.linalg.cholesky`, `solve_triangular` and autograd for the GPLVM loss and gradient, mirroring how TF autodiff
produced it (gplvm.py:425-428) — on all host threads. Only bench.py's cpu_baseline / `--impl reference` legs and the
tests call it; tests/test_oracle_torch_port.py pins it to the numpy oracle (oracle/rbm.py, oracle/gp.py) on injected draws.
"""
import math

import torch


class CdStepTorch:
    """One CD-k iteration of RBM.train on a binary RBM (rbm.py:628-646 chain, rbm.py:664 cd_loss and its gradient in the
    closed form of SURVEY.md section 8 R9, rbm.py:666-670 Adam [TF rule restated, oracle/rbm.py:Adam])."""

    def __init__(self, W, b, c, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, dtype=torch.float32):
        self.dtype = dtype
        self.W = torch.as_tensor(W, dtype=dtype).clone()
        self.b = torch.as_tensor(b, dtype=dtype).reshape(1, -1).clone()
        self.c = torch.as_tensor(c, dtype=dtype).reshape(1, -1).clone()
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta1, beta2, eps, 0
        self.m = [torch.zeros_like(p) for p in (self.W, self.b, self.c)]
        self.v = [torch.zeros_like(p) for p in (self.W, self.b, self.c)]

    def chain(self, v0, k, draws=None, generator=None):
        """-> (p0, vk, pk); draws = [u_h0, (u_v1, u_h1), ...] injected uniforms or None (torch.rand, the counterpart of
        tf.random_uniform inside the graph, rbm.py:202,230). Strict `p > u` (rbm.py:207,235)."""
        def u(shape, i):
            if draws is not None:
                return torch.as_tensor(draws[i], dtype=self.dtype)
            return torch.rand(shape, dtype=self.dtype, generator=generator)
        p0 = torch.sigmoid(v0 @ self.W + self.c)                       # rbm.py:129-131, 160
        h = (p0 > u(p0.shape, 0)).to(self.dtype)
        v = ph = None
        for s in range(k):                                            # rbm.py:351-376
            pv = torch.sigmoid(h @ self.W.t() + self.b)                # rbm.py:143-147, 174
            v = (pv > u(pv.shape, 1 + 2 * s)).to(self.dtype)
            ph = torch.sigmoid(v @ self.W + self.c)
            h = (ph > u(ph.shape, 2 + 2 * s)).to(self.dtype)
        return p0, v, ph

    def gradients(self, v0, p0, vk, pk):
        n = float(v0.shape[0])
        gW = -(v0.t() @ p0 - vk.t() @ pk) / n
        gb = -(v0.sum(0, keepdim=True) - vk.sum(0, keepdim=True)) / n
        gc = -(p0.sum(0, keepdim=True) - pk.sum(0, keepdim=True)) / n
        return gW, gb, gc

    def adam(self, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for p, g, m, v in zip((self.W, self.b, self.c), grads, self.m, self.v):
            m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
            p.sub_(lr_t * m / (v.sqrt() + self.eps))

    def step(self, v0, k, draws=None, generator=None):
        v0 = torch.as_tensor(v0, dtype=self.dtype)
        p0, vk, pk = self.chain(v0, k, draws, generator)
        grads = self.gradients(v0, p0, vk, pk)
        self.adam(grads)
        return grads


def gplvm_loss_and_grad(X, Y, opt_alpha, opt_gamma, opt_noise, jitter=1e-6, dtype=torch.float32):
    """GPLVM loss (gplvm.py:210-244: K = alpha exp(-sqdist / 2) + (noise + jitter) I in the expanded-distance form of
    gplvm.py:61-86, L = cholesky(K), S = L^-1 Y, loss = (|S|^2 + D logdet + D N log 2 pi) / 2) and its gradient with
    respect to X and the three optimisation variables by autograd — what optimizer.minimize differentiates, gplvm.py:425-428.
    -> (loss, dX, d_opt_alpha, d_opt_gamma, d_opt_noise)"""
    X = torch.as_tensor(X, dtype=dtype).clone().requires_grad_(True)
    Y = torch.as_tensor(Y, dtype=dtype)
    hyp = [torch.tensor(float(h), dtype=dtype, requires_grad=True) for h in (opt_alpha, opt_gamma, opt_noise)]
    sp = torch.nn.functional.softplus
    alpha, gamma, noise = sp(hyp[0]), sp(hyp[1]), sp(hyp[2]) + jitter
    N, D = Y.shape
    xx = X / gamma
    sq = (xx * xx).sum(1, keepdim=True)
    K = alpha * torch.exp(-0.5 * (-2.0 * (xx @ xx.t()) + sq + sq.t())) + noise * torch.eye(N, dtype=dtype)
    Lf = torch.linalg.cholesky(K)
    S = torch.linalg.solve_triangular(Lf, Y, upper=False)
    logdet = 2.0 * torch.log(torch.diagonal(Lf)).sum()
    loss = 0.5 * ((S * S).sum() + D * logdet + D * N * math.log(2.0 * math.pi))
    loss.backward()
    return float(loss.detach()), X.grad, float(hyp[0].grad), float(hyp[1].grad), float(hyp[2].grad)
