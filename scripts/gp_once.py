"""One warm-up + one LML+gradient evaluation at N=5000, D=784 (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepbelief_b200 import ops
N, D, Q = (int(sys.argv[1]) if len(sys.argv) > 1 else 5000), (int(sys.argv[2]) if len(sys.argv) > 2 else 784), 2
X = torch.randn(N, Q, device='cuda'); Y = torch.randn(N, D, device='cuda')
hyp_opt = torch.tensor([0.5413, 0.5413, 0.5413], device='cuda')
plan = ops.GplvmPlan(N, Q, D)
for _ in range(2):
    plan.evaluate(X, Y, hyp_opt, 7, want_grad=True)
torch.cuda.synchronize()
print('info', int(plan.info.item()), 'loss', float(plan.scalars[0]))
