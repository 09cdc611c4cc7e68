"""Ordered kernel timeline of ONE GPDBN config-4 training iteration (CUPTI through torch.profiler's chrome trace)."""
import os, sys, json, tempfile, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from deepbelief_b200 import compat
from deepbelief_b200.layers.rbm import RBM
from deepbelief_b200.layers.gprbm import GPRBM
from deepbelief_b200.layers.gplvm import SEKernel
N, D, Q = int(os.environ.get('N', 5000)), int(os.environ.get('D', 784)), 2
dev = 'cuda:0'
rs = np.random.RandomState(0)
Vd = (rs.rand(N, D) < 0.13).astype(np.float32)
X0 = rs.normal(size=(N, Q)).astype(np.float32)
l0 = RBM(D, 200, W_std=0.05, temperature=0.1, name='gpdbn_l0', seed=11, device=dev)
l1 = RBM(200, 100, W_std=0.05, temperature=0.1, bottom=l0, name='gpdbn_l1', seed=12, device=dev)
top = GPRBM(100, 50, Q, N, False, X0=X0, V=Vd, kern=SEKernel(alpha=False, gamma=False, device=dev), noise_variance=0.01,
            fixed_noise_variance=True, W_std=0.05, temperature=0.1, bottom=l1, name='gpdbn_top', seed=13, device=dev)
adam = compat.AdamOptimizer(learning_rate=1e-3)
top.optimize(adam, num_iterations=8)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    top.optimize(adam, num_iterations=1)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), 'trace.json')
prof.export_chrome_trace(path)
evs = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
evs.sort(key=lambda e: e['ts'])
t0 = evs[0]['ts']
print('span %.1f us, %d device activities' % (max(e['ts'] + e['dur'] for e in evs) - t0, len(evs)))
agg = collections.defaultdict(lambda: [0, 0.0])
streams = {}
lines = []
for e in evs:
    a = e.get('args', {})
    sid = streams.setdefault(a.get('stream'), len(streams))
    name = e['name'].replace('(anonymous namespace)::', '').replace('db::', '').split('(')[0][-44:]
    agg[name][0] += 1; agg[name][1] += e['dur']
    lines.append('%9.1f %8.1f  s%d  %-44s grid %s' % (e['ts'] - t0, e['dur'], sid, name, a.get('grid')))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print('%-46s n=%4d %9.1f us' % (k, v[0], v[1]))
print('\n'.join(lines))
