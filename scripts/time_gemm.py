"""SIMT fp32 GEMM core efficiency (db_gemm) at a few shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepbelief_b200 import ops
def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (M, N, K, ta, tb) in ((4096, 4096, 4096, False, True), (4096, 4096, 4096, False, False), (4096, 4096, 4096, True, False),
                          (4096, 4096, 512, False, True), (4096, 4096, 64, False, True), (4992, 448, 64, False, True)):
    A = torch.randn((K, M) if ta else (M, K), device='cuda'); B = torch.randn((N, K) if tb else (K, N), device='cuda')
    C = torch.empty(M, N, device='cuda')
    ms = timeit(lambda: ops.gemm(A, B, trans_a=ta, trans_b=tb, out=C))
    print('M=%d N=%d K=%d ta=%d tb=%d: %.3f ms  %.1f TFLOP/s' % (M, N, K, ta, tb, ms, 2.0 * M * N * K / ms / 1e9))
    ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double()
    print('   relerr %.2e' % float((C.double() - ref).abs().max() / ref.abs().max()))
