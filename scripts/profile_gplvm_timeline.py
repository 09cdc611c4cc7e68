"""Per-kernel durations and gaps of ONE GPLVM LML+gradient evaluation at N=5000 as it really runs (CUPTI through
torch.profiler, warm caches, kernels back to back) — unlike the ncu launch list, whose per-launch times are cold and serialised."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from deepbelief_b200 import ops
N, D, Q = int(os.environ.get('N', 5000)), 784, 2

X = torch.randn(N, Q, device='cuda')
Y = (torch.rand(N, D, device='cuda') < 0.13).float(); Y -= Y.mean(0, keepdim=True)
isp = float(np.log(np.expm1(1.0)))
hyp = torch.tensor([isp, isp, isp], device='cuda')
plan = ops.GplvmPlan(N, Q, D)
for _ in range(3): plan.evaluate(X, Y, hyp, 7, jitter=1e-6)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    plan.evaluate(X, Y, hyp, 7, jitter=1e-6)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if 'CUDA' in str(getattr(e, 'device_type', ''))], key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, ev[-1].time_range.end
agg = collections.defaultdict(lambda: [0, 0.0])
busy, last_end, gaps = 0.0, t0, 0.0
for e in ev:
    name = e.name.replace('(anonymous namespace)::', '').replace('db::', '').split('(')[0][-48:]
    d = e.time_range.end - e.time_range.start
    agg[name][0] += 1; agg[name][1] += d
    gaps += max(0.0, e.time_range.start - last_end)
    last_end = max(last_end, e.time_range.end)
print('span %.1f us, sum of kernel durations %.1f us, idle gaps %.1f us, %d kernels' % (t1 - t0, sum(v[1] for v in agg.values()), gaps, len(ev)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print('%-50s n=%4d %9.1f us' % (k, v[0], v[1]))
print('tf32x3 launches (us):', ' '.join('%.0f' % (e.time_range.end - e.time_range.start) for e in ev if 'tf32x3' in e.name))
