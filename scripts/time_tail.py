"""Per-rank pieces of the 8-GPU step at config 5 (N_local = 2048, V = 4096, H = 8192) on ONE GPU: chain + dW, the fused
apply_update pass, and the separate optimizer + refresh it replaces."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepbelief_b200 import ops, compat
V, H, k = 4096, 8192, 10
N = int(os.environ.get('N', 2048))
W = torch.randn(V, H, device='cuda') * 0.01
params = torch.cat([W.reshape(-1), torch.zeros(V + H, device='cuda')]).contiguous()
v0 = (torch.rand(N, V, device='cuda') < 0.5).float()
plan = ops.CdTcPlan(N, V, H, k, n_global=16384)
plan.refresh_weights(params)
grad = torch.zeros_like(params)
rng = ops.Rng(seed=1)
opt = compat.AdamOptimizer(1e-3)
state = torch.zeros(2 * params.numel(), device='cuda')

def timed(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

print('N=%d chain + dW          : %.3f ms' % (N, timed(lambda: plan.step(v0, params, grad, rng=rng))))
print('apply_update (one pass)  : %.3f ms' % timed(lambda: plan.apply_update(params, grad, opt.descriptor(), state)))
print('optimizer_step           : %.3f ms' % timed(lambda: ops.optimizer_step(opt.descriptor(), params, grad, state)))
print('refresh_weights          : %.3f ms' % timed(lambda: plan.refresh_weights(params)))
print('fused step_update        : %.3f ms' % timed(lambda: plan.step_update(v0, params, grad, opt.descriptor(), state, rng=rng)))
