"""Gram kernel (K3) bandwidth: algorithmic bytes 4 N^2 + 4 N Q per launch / CUDA-event time."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepbelief_b200 import ops
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
for N, Q in ((5000, 2), (8192, 2), (5000, 8)):
    X = torch.randn(N, Q, device='cuda')
    hyp = torch.tensor([1.0, 1.0, 0.5], device='cuda')
    for _ in range(3):
        ops.gram_rbf(X, None, hyp, add_noise=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 60
    e0.record()
    for i in range(reps):
        ops.gram_rbf(X, None, hyp, add_noise=True)          # the caching allocator cycles through the freed 100 MB blocks
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    gb = (4.0 * N * N + 4.0 * N * Q) / 1e9
    print('gram N=%d Q=%d: %.1f us, %.0f GB/s' % (N, Q, us, gb / (us * 1e-6)), flush=True)
# write-only ceiling of this box for the same footprint: a plain fill of N^2 floats (library kernel, reference only)
for N in (5000, 8192):
    bufs = [torch.empty(N, N, device='cuda') for _ in range(3)]
    for b in bufs: b.fill_(1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(60): bufs[i % 3].fill_(1.0)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 60 * 1e3
    print('fill  N=%d: %.1f us, %.0f GB/s (write-only reference)' % (N, us, 4.0 * N * N / 1e9 / (us * 1e-6)), flush=True)
print('peaks', {k: v for k, v in peaks.items() if 'hbm' in k.lower() or 'GB' in str(k)})
