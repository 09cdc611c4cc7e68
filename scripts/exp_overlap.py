"""A/B of the overlapped triangular inverse (DB_GP_OVERLAP=0/1/2, read once per process): times the GPLVM LML+grad
evaluation at N=5000, D=784 graph-replayed and eager, the GPDBN factor, and compares the outputs with the mode-0 run
(gpurun_out/overlap_ref.npz, written by the mode-0 process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepbelief_b200 import ops

mode = os.environ.get('DB_GP_OVERLAP', '1')
N, D, Q = (int(sys.argv[1]), int(sys.argv[2])) + (2,) if len(sys.argv) > 2 else (5000, 784, 2)
torch.manual_seed(0)
X = torch.randn(N, Q, device='cuda'); Y = (torch.rand(N, D, device='cuda') < 0.13).float()
Y = Y - Y.mean(0, keepdim=True)
hyp_opt = torch.tensor([0.5413, 0.5413, 0.5413], device='cuda')

def timeit(fn, iters=20):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

plan = ops.GplvmPlan(N, Q, D)
plan.use_graph = False
t_eager = timeit(lambda: plan.evaluate(X, Y, hyp_opt, 7, want_grad=True))
sc, dX, info = plan.evaluate(X, Y, hyp_opt, 7, want_grad=True)
torch.cuda.synchronize()
out = {'scalars': sc.cpu().numpy().copy(), 'dX': dX.cpu().numpy().copy(), 'info': info.cpu().numpy().copy()}
plan2 = ops.GplvmPlan(N, Q, D)
t_graph = timeit(lambda: plan2.evaluate(X, Y, hyp_opt, 7, want_grad=True))
sc2, dX2, _ = plan2.evaluate(X, Y, hyp_opt, 7, want_grad=True)
torch.cuda.synchronize()
g_vs_e = float((dX2 - torch.from_numpy(out['dX']).cuda()).abs().max() / np.abs(out['dX']).max())
g_sc = float(np.max(np.abs(sc2.cpu().numpy() - out['scalars']) / (np.abs(out['scalars']) + 1e-30)))
if os.environ.get('REPEAT'):
    # replays of the same graph must agree with each other too
    outs = []
    for _ in range(4):
        plan2.evaluate(X, Y, hyp_opt, 7, want_grad=True); torch.cuda.synchronize(); outs.append(plan2.dX.clone())
    print('replay spread', [float((o - outs[0]).abs().max() / outs[0].abs().max()) for o in outs[1:]], 'scalars g-vs-e %.1e' % g_sc)
msg = 'mode %s N=%d D=%d: eager %.3f ms  graph %.3f ms (graphs captured: %d)  info %d  graph-vs-eager dX %.1e' % (
    mode, N, D, t_eager, t_graph, len(plan2._graphs), int(out['info'][0]), g_vs_e)
ref_path = os.path.join(os.environ.get('OVERLAP_REF_DIR', 'gpurun_out'), 'overlap_ref_%d_%d.npz' % (N, D))
if mode == '0':
    np.savez(ref_path, **out)
elif os.path.exists(ref_path):
    ref = np.load(ref_path)
    msg += '  vs mode 0: scalars %.1e dX %.1e' % (
        float(np.max(np.abs(out['scalars'] - ref['scalars']) / (np.abs(ref['scalars']) + 1e-30))),
        float(np.abs(out['dX'] - ref['dX']).max() / np.abs(ref['dX']).max()))
print(msg, flush=True)
