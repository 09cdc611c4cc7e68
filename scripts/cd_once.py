"""Two CD-10 updates at config 5 (V=4096, H=8192, N=16384) + one stand-alone apply_update pass (for ncu captures)."""
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
for _ in range(2):                    # RBM.cd_update: chain + dW with the fused optimizer / cache rewrite (db_cd_tc_step_update)
    rbm.cd_update(v0, k, opt, st, grad)
plan = rbm._tc_plan                   # and the data-parallel tail as a stand-alone pass (db_cd_tc_apply_update), for its capture
plan.apply_update(rbm._params, grad, opt.descriptor(), st)
torch.cuda.synchronize()
print('ok', float(grad.abs().max()))
