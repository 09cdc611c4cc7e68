"""Two warm-up + N_IT training iterations of the config-4 GPDBN stack (784-200-100-GPRBM(50), N=5000) for ncu launch lists / timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepbelief_b200 import compat
from deepbelief_b200.layers.rbm import RBM
from deepbelief_b200.layers.gprbm import GPRBM
from deepbelief_b200.layers.gplvm import SEKernel
N, D, Q = (int(sys.argv[1]) if len(sys.argv) > 1 else 5000), (int(sys.argv[2]) if len(sys.argv) > 2 else 784), 2    # `328 1024`: the gpdbn_horses sizes
n_it = int(os.environ.get('N_IT', 1))
dev = 'cuda:0'
rs = np.random.RandomState(0)
Vd = (rs.rand(N, D) < 0.13).astype(np.float32)
X0 = rs.normal(size=(N, Q)).astype(np.float32)
l0 = RBM(D, 200, W_std=0.05, temperature=0.1, name='gpdbn_l0', seed=11, device=dev)
l1 = RBM(200, 100, W_std=0.05, temperature=0.1, bottom=l0, name='gpdbn_l1', seed=12, device=dev)
top = GPRBM(100, 50, Q, N, False, X0=X0, V=Vd, kern=SEKernel(alpha=False, gamma=False, device=dev), noise_variance=0.01,
            fixed_noise_variance=True, W_std=0.05, temperature=0.1, bottom=l1, name='gpdbn_top', seed=13, device=dev)
adam = compat.AdamOptimizer(learning_rate=1e-3)
top.optimize(adam, num_iterations=2)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
top.optimize(adam, num_iterations=n_it)
e1.record(); torch.cuda.synchronize()
print('gpdbn iteration %.3f ms' % (e0.elapsed_time(e1) / n_it))
if os.environ.get('PROFILE'):
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    top.optimize(adam, num_iterations=2); torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
