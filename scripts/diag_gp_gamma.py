"""Diagnostic (GPU): which stage of the N=5000 LML+gradient loses d_gamma accuracy. fp64 re-evaluations of the closed-form
gradient with one device-computed ingredient at a time (Gram matrix, Cholesky factor)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.linalg
from deepbelief_b200 import ops
from oracle import gp as ogp

N, D, Q = int(os.environ.get('N', 5000)), 784, 2
rs = np.random.RandomState(5000)
X = rs.normal(size=(N, Q)).astype(np.float32)
Y = (rs.uniform(size=(N, D)) < 0.13).astype(np.float32)
Y -= Y.mean(axis=0, keepdims=True)
orc = ogp.GPLVMOracle(Y, X, jitter=1e-6)
w_dX, w_da, w_dg, w_dn = orc.grad()
X64, Y64 = X.astype(np.float64), Y.astype(np.float64)
alpha, gamma, noise = float(orc.alpha), float(orc.gamma), float(orc.noise)
Kf64 = ogp.se_kernel(X64, None, alpha, gamma)
r2 = ogp.sq_dist(X64, None, 1.0)
sg = float(ogp.sigmoid(np.asarray(orc.opt_gamma)))

def dgamma_from_Kinv(Kinv):
    A = Kinv @ Y64
    G = 0.5 * (D * Kinv - A @ A.T)
    return np.sum(G * Kf64 * r2) / gamma ** 3 * sg, np.sum(G * Kf64) / alpha

print('oracle d_gamma %.9g' % w_dg)
cu = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).cuda()
hyp = cu([orc.opt_alpha, orc.opt_gamma, orc.opt_noise])
plan = ops.GplvmPlan(N, Q, D)
Lout = torch.empty(N, N, device='cuda')
scal, dX, info = plan.evaluate(cu(X), cu(Y), hyp, 7, jitter=1e-6, L_out=Lout)
torch.cuda.synchronize()
print('gpu d_gamma %.9g relerr %.3e ; d_alpha relerr %.3e' % (float(scal[4]), abs(float(scal[4]) - w_dg) / abs(w_dg), abs(float(scal[3]) - w_da) / abs(w_da)))
hyp_eff = ops.gp_effective_hyp(hyp, 7, 0.0, 1e-6)
Kg = ops.gram_rbf(cu(X), None, hyp_eff, add_noise=True).cpu().numpy().astype(np.float64) if 'gram_rbf' in dir(ops) else None
K64 = Kf64 + noise * np.eye(N)
if Kg is not None:
    if abs(Kg[0, 0] - K64[0, 0]) > 0.5: Kg = Kg + noise * np.eye(N)
    rel = (Kg - K64) / K64
    big = Kf64 > 1e-3
    print('Gram: rel err mean %.3e rms %.3e max %.3e (entries > 1e-3); diag err %.3e' % (rel[big].mean(), np.sqrt((rel[big] ** 2).mean()), np.abs(rel[big]).max(), np.abs(np.diag(Kg) - np.diag(K64)).max()))
    dg, da = dgamma_from_Kinv(np.linalg.inv(Kg))
    print('fp64 pipeline on the device Gram: d_gamma relerr %.3e, d_alpha*sig relerr %.3e' % (abs(dg - w_dg) / abs(w_dg), abs(da * float(ogp.sigmoid(np.asarray(orc.opt_alpha))) - w_da) / abs(w_da)))
    # fp32-rounded exact K
    K32 = K64.astype(np.float32).astype(np.float64)
    dg, da = dgamma_from_Kinv(np.linalg.inv(K32))
    print('fp64 pipeline on the fp32-rounded exact Gram: d_gamma relerr %.3e' % (abs(dg - w_dg) / abs(w_dg)))
Lg = np.tril(Lout.cpu().numpy().astype(np.float64))
Li = scipy.linalg.solve_triangular(Lg, np.eye(N), lower=True)
dg, da = dgamma_from_Kinv(Li.T @ Li)
print('fp64 pipeline on the device Cholesky factor: d_gamma relerr %.3e' % (abs(dg - w_dg) / abs(w_dg)))
L64 = np.linalg.cholesky(Kg if Kg is not None else K64)
print('device factor vs fp64 factor of the device Gram: max abs %.3e, rel fro %.3e' % (np.abs(Lg - L64).max(), np.linalg.norm(Lg - L64) / np.linalg.norm(L64)))
