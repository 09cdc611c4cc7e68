"""Diagnostic: host->device bandwidth and where the time of the e2e (RBM.train) step goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
dev = 'cuda'
n, V = 16384, 4096
host = torch.empty((2 * n, V), dtype=torch.float32).pin_memory()
view = torch.from_numpy(host.numpy()[:n])
print('view pinned:', view.is_pinned(), 'host pinned:', host.is_pinned())
d = torch.empty((n, V), device=dev)
for src, name in ((host[:n], 'pinned tensor'), (view, 'numpy view of pinned'), (torch.empty((n, V)), 'pageable')):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        d.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print('%s: %.1f ms  %.1f GB/s' % (name, dt * 1e3, n * V * 4 / dt / 1e9))
# overlap test: copy on side stream while a long kernel chain runs
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
s2 = torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): a @ a
torch.cuda.synchronize(); t_mm = time.perf_counter() - t0
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s2):
    d.copy_(host[:n], non_blocking=True)
for _ in range(20): a @ a
torch.cuda.synchronize(); t_both = time.perf_counter() - t0
print('matmul alone %.1f ms, matmul + concurrent H2D %.1f ms' % (t_mm * 1e3, t_both * 1e3))
