"""CD-10 chain timing at the per-rank batch sizes of the 1/2/4/8-GPU runs (V=4096, H=8192)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepbelief_b200 import ops
V, H, k = 4096, 8192, 10
W = torch.randn(V, H, device='cuda') * 0.01
params = torch.cat([W.reshape(-1), torch.zeros(V + H, device='cuda')]).contiguous()
for N in [int(x) for x in os.environ.get('NS', '2048,4096,8192,16384').split(',')]:
    v0 = (torch.rand(N, V, device='cuda') < 0.5).float()
    plan = ops.CdTcPlan(N, V, H, k, n_global=16384)
    plan.refresh_weights(params)
    grad = torch.zeros_like(params)
    rng = ops.Rng(seed=1)
    for _ in range(2): plan.step(v0, params, grad, rng=rng)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    iters = 5
    e0.record()
    for _ in range(iters): plan.step(v0, params, grad, rng=rng)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = (2 * k + 3) * 2.0 * N * V * H
    print('N=%5d: %.3f ms/step  %.1f TFLOP/s algorithmic  (x%d = %.2f ms)' % (N, ms, fl / ms / 1e9, 16384 // N, ms * (16384 // N)))
    del plan, v0
