"""Two CD-10 steps at config 5 (V=4096, H=8192, N=16384) incl. optimizer + weight refresh (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepbelief_b200 import compat
from deepbelief_b200.layers.rbm import RBM
V, H, N, k = 4096, 8192, 16384, 10
np.random.seed(0)
rbm = RBM(V, H, W_std=0.01, name='cd_once', seed=7)
opt = compat.AdamOptimizer(1e-3)
st = rbm.new_optimizer_state(opt)
v0 = (torch.rand(N, V, device='cuda') < 0.5).float()
grad = torch.empty_like(rbm._params)
for _ in range(2):
    rbm.cd_step(v0, k, grad=grad)
    rbm.apply_gradient(opt, grad, st)
torch.cuda.synchronize()
print('ok', float(grad.abs().max()))
