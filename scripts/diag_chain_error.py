"""Diagnostic (GPU): error of the tcgen05 chain GEMM's pre-activation x = vW + c against fp64 at the config-5 depths
(K = 4096 and K = 8192), on sampled rows: mean / spread / max of (x_gpu - x64) / |x64| and of the probability error, split
by the sign of x (a truncating accumulator shrinks toward zero)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepbelief_b200 import ops

def run(N, V, H, tag):
    rs = np.random.RandomState(1)
    W = (rs.normal(size=(V, H)) * 0.01).astype(np.float32)
    b = (rs.normal(size=V) * 0.1).astype(np.float32)
    c = (rs.normal(size=H) * 0.1).astype(np.float32)
    gen = torch.Generator(device='cuda').manual_seed(0)
    v0 = (torch.rand(N, V, device='cuda', generator=gen) < 0.5).float()
    params = torch.as_tensor(np.concatenate([W.ravel(), b, c])).cuda()
    plan = ops.CdTcPlan(N, V, H, 1)
    plan.refresh_weights(params)
    grad = torch.zeros_like(params)
    p0 = torch.empty(N, H, device='cuda')
    plan.step(v0, params, grad, rng=ops.Rng(seed=3), p0_out=p0)
    torch.cuda.synchronize()
    rows = np.arange(0, N, max(1, N // 256))[:256]
    v = v0[rows].cpu().numpy().astype(np.float64)
    x64 = v @ W.astype(np.float64) + c.astype(np.float64)
    p64 = 1.0 / (1.0 + np.exp(-x64))
    p = p0[rows].cpu().numpy().astype(np.float64)
    x = np.log(p) - np.log1p(-p)
    rel = (x - x64) / np.abs(x64)
    big = np.abs(x64) > 0.5
    prel = np.abs(p - p64) / p64
    print('%s K=%d: x relerr (|x|>0.5): signed-toward-zero mean %.2e, std %.2e, max %.2e; prob relerr max %.2e mean %.2e; abs x err max %.2e'
          % (tag, V, np.mean(-np.sign(x64[big]) * (x - x64)[big] / np.abs(x64[big])), np.std(rel[big]), np.max(np.abs(rel[big])),
             prel.max(), prel.mean(), np.max(np.abs(x - x64))))

run(2048, 4096, 8192, 'up')
run(2048, 8192, 4096, 'deep')
run(2048, 1024, 512, 'small')
