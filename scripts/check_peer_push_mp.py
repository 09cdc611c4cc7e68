"""Multi-process check (torchrun, one rank per GPU) of the data-parallel step with the gradient reduce-scattered by PUSH from
the dW epilogue (db_cd_tc_step_push + db_cd_tc_apply_update_sharded) against the NCCL all-reduce path, on the same seeds:
the parameters after a few steps must agree to fp32 summation order, and every rank must hold bit-identical parameters."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
from deepbelief_b200 import compat
from deepbelief_b200.layers.rbm import RBM

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = 'cuda:%d' % local
td.init_process_group(backend='nccl', device_id=torch.device(dev))
V, H, k = int(os.environ.get('V', 4096)), int(os.environ.get('H', 4096)), 2
n_local = int(os.environ.get('NLOCAL', 512))
n_global = n_local * world
out = {}
for mode in ('push', 'nccl'):
    RBM.PEER_PUSH = mode == 'push'
    np.random.seed(0)
    rbm = RBM(V, H, W_std=0.01, name='pp_' + mode, seed=7, device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    rbm.set_weights(W=torch.randn(V, H, device=dev, generator=gen) * 0.01)
    rbm.rng.row_offset = rank * n_local
    # plain SGD: Adam's m / sqrt(v) turns summation-order noise on near-zero gradient entries into full-size steps
    opt = compat.GradientDescentOptimizer(learning_rate=0.1)
    st = rbm.new_optimizer_state(opt)
    grad = torch.zeros_like(rbm._params)
    g2 = torch.Generator(device=dev).manual_seed(100 + rank)
    batches = [(torch.rand(n_local, V, device=dev, generator=g2) < 0.5).float() for _ in range(2)]
    rbm.cd_update(batches[0], k, opt, st, grad, n_global=n_global)
    torch.cuda.synchronize()
    p1 = rbm._params.clone()            # after ONE update: the two paths differ by the summation order over ranks only
    for it in range(1, 4):
        rbm.cd_update(batches[it % 2], k, opt, st, grad, n_global=n_global)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(10):
        rbm.cd_update(batches[it % 2], k, opt, st, grad, n_global=n_global)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    used_push = bool(getattr(rbm, '_peer', None))
    p = rbm._params.clone()
    # replicas identical?
    chk = torch.stack([p.double().sum(), p.double().abs().sum(), (p.double() * torch.arange(p.numel(), device=dev) % 7).sum()])
    allchk = [torch.zeros_like(chk) for _ in range(world)]
    td.all_gather(allchk, chk)
    same = all(torch.equal(allchk[0], c) for c in allchk)
    out[mode] = (p, used_push, same, ms, p1)
if rank == 0:
    a, b = out['push'][4], out['nccl'][4]
    rel = float((a - b).abs().max() / b.abs().max())
    rel_end = float((out['push'][0] - out['nccl'][0]).abs().max() / out['nccl'][0].abs().max())
    print('after 14 updates (sampled chains diverge once one bit flips): %.2e' % rel_end)
    print('world %d V %d H %d n_local %d: push used %s / %s; replicas identical push %s nccl %s; max |dtheta| / max|theta| = %.2e; ms/step push %.3f nccl %.3f'
          % (world, V, H, n_local, out['push'][1], out['nccl'][1], out['push'][2], out['nccl'][2], rel, out['push'][3], out['nccl'][3]))
    assert out['push'][1] and not out['nccl'][1]
    assert out['push'][2] and out['nccl'][2]
    assert rel < 1e-6, 'after one update the two paths may differ by the fp32 summation order over ranks only'
    assert rel_end < 5e-3
    print('PEER PUSH OK')
td.destroy_process_group()
