"""Where does the error of the noise-variance gradient (tr G, a ~50x cancelling difference) come from?
fp64 torch is the checker only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepbelief_b200 import ops
from oracle import gp as ogp

N, D = 1500, 96
rs = np.random.RandomState(1)
X = rs.normal(size=(N, 2)); Y = rs.normal(size=(N, D)); Y -= Y.mean(0)
orc = ogp.GPLVMOracle(Y, X, jitter=1e-6)
w_dX, w_da, w_dg, w_dn = orc.grad()
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()
plan = ops.GplvmPlan(N, 2, D)
hyp_opt = cu([orc.opt_alpha, orc.opt_gamma, orc.opt_noise])
Lf = torch.zeros(N, N, device='cuda')
scal, dX, info = plan.evaluate(cu(X), cu(Y), hyp_opt, 7, jitter=1e-6, L_out=Lf)
sig = 1.0 / (1.0 + np.exp(-orc.opt_noise))
print('oracle dn %.6f  gpu dn %.6f  rel %.3e' % (w_dn, float(scal[5]), abs(float(scal[5]) - w_dn) / abs(w_dn)))
hyp = ops.gp_effective_hyp(hyp_opt, 7, 0.0, 1e-6)
K32 = ops.gram_rbf(cu(X), None, hyp, add_noise=True)
Y64 = cu(Y).double()
def trG(L):
    Linv = torch.linalg.solve_triangular(L, torch.eye(N, device='cuda', dtype=torch.float64), upper=False)
    trK = (Linv ** 2).sum()
    A = Linv.t() @ (Linv @ Y64)
    return 0.5 * (D * trK - (A ** 2).sum()), trK, (A ** 2).sum()
t2 = trG(torch.linalg.cholesky(K32.double()))
t3 = trG(torch.tril(Lf).double())
print('trG oracle(fp64 K) %.6f' % (w_dn / sig))
print('trG fp64 from fp32 K %.6f  (D trKinv %.6f, |A|^2 %.6f)' % tuple(float(v) for v in (t2[0], D * t2[1], t2[2])))
print('trG fp64 from fp32 L %.6f  (D trKinv %.6f, |A|^2 %.6f)' % tuple(float(v) for v in (t3[0], D * t3[1], t3[2])))
print('trG gpu %.6f' % (float(scal[5]) / sig))
