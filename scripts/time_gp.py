"""Timing of the GP chain stages at N=5000 (and 328): potrf alone, LML only, LML+grad."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepbelief_b200 import ops

def timeit(fn, iters=10):
    for _ in range(3): fn()        # the third call of a plan replays its CUDA graph
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for N, D in ((5000, 784), (328, 1024), (2000, 50)):
    Q = 2
    X = torch.randn(N, Q, device='cuda'); Y = torch.randn(N, D, device='cuda')
    hyp_opt = torch.tensor([0.5413, 0.5413, 0.5413], device='cuda')
    hyp = ops.gp_effective_hyp(hyp_opt, 7, 0.0, 1e-6)
    K = ops.gram_rbf(X, None, hyp, add_noise=True)
    A = K.clone()
    def potrf():
        A.copy_(K); return ops.potrf_lower(A)
    t_copy = timeit(lambda: A.copy_(K))
    t_potrf = timeit(potrf) - t_copy
    info = potrf(); torch.cuda.synchronize()
    Lr = torch.linalg.cholesky(K.double())
    err = float((torch.tril(A).double() - Lr).abs().max() / Lr.abs().max())
    plan = ops.GplvmPlan(N, Q, D)
    t_lml = timeit(lambda: plan.evaluate(X, Y, hyp_opt, 7, want_grad=False))
    t_all = timeit(lambda: plan.evaluate(X, Y, hyp_opt, 7, want_grad=True))
    print('N=%d D=%d: potrf %.3f ms (info %d, L err %.1e)  lml %.3f ms  lml+grad %.3f ms' % (N, D, t_potrf, int(info), err, t_lml, t_all))
