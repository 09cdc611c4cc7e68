"""Kernel timeline (CUPTI via torch.profiler) of graph-replayed RBM.train iterations at the reference's layer sizes, tensor-core
chain (TC_MIN_WORK=0) or SIMT (TC=0): per-kernel durations of the last replayed iterations and the span per iteration."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from deepbelief_b200 import compat as tf
from deepbelief_b200.layers.data import Data
from deepbelief_b200.layers.rbm import RBM
N, V, H, k = [int(x) for x in os.environ.get('SHAPE', '328,1024,496,1').split(',')]
if os.environ.get('TC', '1') == '1':
    RBM.TC_MIN_WORK = 0
else:
    RBM.TC_MIN_WORK = RBM.TC_MIN_WORK_LONG_CHAIN = 1 << 62
rs = np.random.RandomState(0)
X = (rs.rand(N, V) < 0.4).astype(np.float32)
lay = RBM(V, H, name='p', seed=1)
opt = tf.train.AdamOptimizer(learning_rate=1e-3)
data = Data.from_arrays(X, batch_size=N, log_epochs=False)
lay.train(opt, data, None, num_gibbs_steps=k, pcd=False, max_iterations=80)
torch.cuda.synchronize()
lay.checkpoint_iter = 1
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    lay.train(opt, data, None, num_gibbs_steps=k, pcd=False, max_iterations=80)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if 'CUDA' in str(getattr(e, 'device_type', ''))], key=lambda e: e.time_range.start)
# the graph-replayed part: last 40 iterations
names = [e.name.replace('(anonymous namespace)::', '').replace('db::', '').split('(')[0][-44:] for e in ev]
per_iter = names.count('clock_advance_kernel')
idx = [i for i, n in enumerate(names) if n == 'clock_advance_kernel']
if len(idx) > 12:
    a, b = idx[-11] + 1, idx[-1] + 1
    sel = ev[a:b]
    span = (sel[-1].time_range.end - sel[0].time_range.start) / 10
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e, n in zip(sel, names[a:b]):
        agg[n][0] += 1; agg[n][1] += e.time_range.end - e.time_range.start
    print('shape %s tc=%s: %.1f us per replayed iteration, kernel time %.1f us, %d kernels per iteration' % (
        (N, V, H, k), os.environ.get('TC', '1'), span, sum(v[1] for v in agg.values()) / 10, len(sel) // 10))
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('  %-46s n=%3d  %7.1f us each' % (n, v[0] // 10, v[1] / v[0]))
else:
    print('no graph replay found', per_iter, len(ev))
