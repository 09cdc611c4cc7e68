"""Where the end-to-end GPLVM.optimize time goes: constructor, first (eager) evaluation, capture, replays."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepbelief_b200 import compat, ops
from deepbelief_b200.layers.gplvm import GPLVM
N, D, Q = 5000, 784, 2
rs = np.random.RandomState(0)
Y = (rs.rand(N, D) < 0.13).astype(np.float32); Y -= Y.mean(0, keepdims=True)
X = rs.randn(N, Q).astype(np.float32)
def run(iters):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = GPLVM(Y, Q, X0=X, name='diag', device='cuda:0')
    torch.cuda.synchronize(); t1 = time.perf_counter()
    m.optimize(compat.AdamOptimizer(learning_rate=1e-3), num_iterations=iters)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    loss = float(m.loss_and_grad()[0].item())
    t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, len(m._plan._graphs), m._plan.use_graph
for it in (1, 2, 3, 10, 100, 100):
    c, o, l, g, ug = run(it)
    print('iters %3d: constructor %.1f ms, optimize %.1f ms (%.3f ms/iter), final loss_and_grad %.1f ms, graphs %d use_graph %s' % (it, c, o, o / it, l, g, ug), flush=True)
