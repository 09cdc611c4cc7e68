"""torchrun check (N>=2 GPUs): RBM.cd_update (chunk-pipelined dW / all-reduce / optimizer) gives the same parameters
as cd_step + apply_gradient, and the replicas stay identical across ranks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
from deepbelief_b200 import compat
from deepbelief_b200.layers.rbm import RBM

rank, world, lr_ = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(lr_)
os.environ.setdefault('NCCL_MAX_CTAS', str(RBM.COMM_SMS))
td.init_process_group('nccl', device_id=torch.device('cuda:%d' % lr_))
V, H, Ng, k = 2048, 1024, 4096, 2
n = Ng // world
gen = torch.Generator(device='cuda').manual_seed(0)
W0 = torch.randn(V, H, device='cuda', generator=gen) * 0.02
data = (torch.rand(Ng, V, device='cuda', generator=gen) < 0.4).float()[rank * n:(rank + 1) * n].contiguous()
res = []
for mode in ('plain', 'pipelined'):
    np.random.seed(0)
    r = RBM(V, H, name='chk', seed=5)
    r.set_weights(W=W0)
    r.rng.row_offset = rank * n
    opt = compat.AdamOptimizer(1e-3)
    st = r.new_optimizer_state(opt)
    g = torch.empty_like(r._params)
    for it in range(3):
        if mode == 'plain':
            r.cd_step(data, k, grad=g, n_global=Ng)
            r.apply_gradient(opt, g, st)
        else:
            r.cd_update(data, k, opt, st, g, n_global=Ng)
    torch.cuda.synchronize()
    res.append(r._params.clone())
diff = float((res[0] - res[1]).abs().max())
ref = [torch.empty_like(res[1]) for _ in range(world)]
td.all_gather(ref, res[1])
same = all(torch.equal(ref[0], t) for t in ref)
print('rank %d: max |plain - pipelined| = %.3e, replicas identical: %s' % (rank, diff, same))
assert diff <= 1e-7 and same
td.destroy_process_group()
