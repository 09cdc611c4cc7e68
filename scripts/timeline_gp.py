"""Ordered kernel timeline (start offset, duration, stream, grid) of ONE GPLVM LML+gradient evaluation (CUPTI through
torch.profiler's chrome trace), to see how the factorisation and the side-stream triangular inverse share the GPU.
    N=5000 GRAPH=1 python scripts/timeline_gp.py > gpurun_out/timeline.txt"""
import os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from deepbelief_b200 import ops
N, D, Q = int(os.environ.get('N', 5000)), int(os.environ.get('D', 784)), 2
X = torch.randn(N, Q, device='cuda')
Y = (torch.rand(N, D, device='cuda') < 0.13).float(); Y -= Y.mean(0, keepdim=True)
isp = float(np.log(np.expm1(1.0)))
hyp = torch.tensor([isp, isp, isp], device='cuda')
plan = ops.GplvmPlan(N, Q, D)
plan.use_graph = os.environ.get('GRAPH', '1') != '0'
for _ in range(4): plan.evaluate(X, Y, hyp, 7, jitter=1e-6)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    plan.evaluate(X, Y, hyp, 7, jitter=1e-6)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), 'trace.json')
prof.export_chrome_trace(path)
evs = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
evs.sort(key=lambda e: e['ts'])
t0 = evs[0]['ts']
streams = {}
print('span %.1f us, %d device activities' % (max(e['ts'] + e['dur'] for e in evs) - t0, len(evs)))
for e in evs:
    a = e.get('args', {})
    sid = streams.setdefault(a.get('stream'), len(streams))
    name = e['name'].replace('(anonymous namespace)::', '').replace('db::', '').split('(')[0][-40:]
    print('%9.1f %8.1f  s%d  %-40s grid %s' % (e['ts'] - t0, e['dur'], sid, name, a.get('grid')))
