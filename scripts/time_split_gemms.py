"""Kernel time (CUPTI) of the two fp32-accurate tensor-core GEMM forms at GP-chain shapes: 3xTF32 vs 3 x fp16 (scaled pairs)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from deepbelief_b200 import ops
for (M, N, K, lower) in ((4096, 4096, 4096, False), (5000, 5000, 5000, True), (784, 5000, 5000, False), (5000, 5000, 784, True), (4488, 4488, 512, True)):
    X = torch.randn(M, K, device='cuda'); Y = X if lower else torch.randn(N, K, device='cuda')
    res = {}
    for name, fn in (('tf32x3', ops.gemm_tf32x3_nt), ('f16x3', ops.gemm_f16x3_nt)):
        for _ in range(2): fn(X, Y, lower_only=lower)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3): fn(X, Y, lower_only=lower)
            torch.cuda.synchronize()
        d = [e.time_range.end - e.time_range.start for e in prof.events() if 'x3_gemm_kernel' in e.name]
        res[name] = sum(d) / len(d)
    flop = 2.0 * M * N * K * (0.5 if lower else 1.0)
    print('%5d x %5d x %5d %s: tf32x3 %.1f us (%.0f TFLOP/s useful)  f16x3 %.1f us (%.0f TFLOP/s useful)  x%.2f' % (
        M, N, K, 'lower' if lower else 'full ', res['tf32x3'], flop / res['tf32x3'] / 1e6, res['f16x3'], flop / res['f16x3'] / 1e6,
        res['tf32x3'] / res['f16x3']))
