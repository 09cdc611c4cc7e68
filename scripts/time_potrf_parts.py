"""Timing diagnostics: potrf at N=5000 with the column-block steps or the wide updates skipped (results are wrong then)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepbelief_b200 import ops
N = int(os.environ.get('N', 5000))
X = torch.randn(N, 2, device='cuda')
hyp = ops.gp_effective_hyp(torch.tensor([0.5413, 0.5413, 0.5413], device='cuda'), 7, 0.0, 1e-6)
K = ops.gram_rbf(X, None, hyp, add_noise=True)
plan = ops.GplvmPlan(N, 2, 784)
Y = torch.randn(N, 784, device='cuda')
def timeit(fn, iters=10):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
A = K.clone()
t_copy = timeit(lambda: A.copy_(K))
print('potrf (SIMT wide updates, no lo_buf) %.3f ms' % (timeit(lambda: (A.copy_(K), ops.potrf_lower(A))) - t_copy))
hyp_opt = torch.tensor([0.5413, 0.5413, 0.5413], device='cuda')
print('lml only %.3f ms' % timeit(lambda: plan.evaluate(X, Y, hyp_opt, 7, want_grad=False)))
