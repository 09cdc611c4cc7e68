"""torchrun probe of the data-parallel CD step variants at config 5 (per-rank batch 16384 / world).
env MODE = plain | pipe ; COMM_SMS = SMs left to NCCL in pipe mode ; DW_CHUNKS."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
from deepbelief_b200 import compat
from deepbelief_b200.layers.rbm import RBM
rank, world, lr_ = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(lr_)
td.init_process_group('nccl', device_id=torch.device('cuda:%d' % lr_))
mode = os.environ.get('MODE', 'plain')
RBM.COMM_SMS = int(os.environ.get('COMM_SMS', '16'))
RBM.DW_CHUNKS = int(os.environ.get('DW_CHUNKS', '4'))
V, H, Ng, k = 4096, 8192, 16384, 10
n = Ng // world
np.random.seed(0)
r = RBM(V, H, W_std=0.01, name='probe', seed=7)
r.rng.row_offset = rank * n
opt = compat.AdamOptimizer(1e-3)
st = r.new_optimizer_state(opt)
g = torch.empty_like(r._params)
data = (torch.rand(n, V, device='cuda') < 0.5).float()
def step():
    if mode == 'plain':
        r.cd_step(data, k, grad=g, n_global=Ng); r.apply_gradient(opt, g, st)
    else:
        r.cd_update(data, k, opt, st, g, n_global=Ng)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    td.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device='cuda'); td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t)
t_step = timeit(step)
t_ar = timeit(lambda: td.all_reduce(g))
t_ar4 = timeit(lambda: [td.all_reduce(g[i * (g.numel() // 4):(i + 1) * (g.numel() // 4)]) for i in range(4)])
t_chain = timeit(lambda: r.cd_step(data, k, grad=g, n_global=Ng))
if rank == 0:
    print('MODE=%s COMM_SMS=%d CHUNKS=%d NCCL_MAX_CTAS=%s world=%d: step %.3f ms | allreduce(134MB) %.3f ms, 4 chunks %.3f ms | cd_step alone %.3f ms'
          % (mode, RBM.COMM_SMS, RBM.DW_CHUNKS, os.environ.get('NCCL_MAX_CTAS', '-'), world, t_step, t_ar, t_ar4, t_chain))
td.destroy_process_group()
