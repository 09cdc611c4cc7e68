"""3xTF32 tcgen05 GEMM: accuracy vs fp64 and throughput."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepbelief_b200 import ops
torch.manual_seed(0)
def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (M, N, K, lower) in ((256, 128, 64, False), (1000, 904, 520, False), (4488, 4488, 512, True), (4096, 4096, 4096, False), (5000, 5000, 5000, True), (784, 5000, 5000, False)):
    X = torch.randn(M, K, device='cuda'); Y = X if (lower and M == N) else torch.randn(N, K, device='cuda')
    C0 = torch.randn(M, N, device='cuda')
    Xl = ops.tf32_split_lo(X); Yl = Xl if Y is X else ops.tf32_split_lo(Y)
    C = C0.clone()
    ops.gemm_tf32x3_nt(X, Y, alpha=-1.0, beta=1.0, out=C, lower_only=lower, X_lo=Xl, Y_lo=Yl)
    ref = C0.double() - X.double() @ Y.double().t()
    if lower:
        mask = torch.tril(torch.ones(M, N, device='cuda', dtype=torch.bool))
        err = float(((C.double() - ref).abs() * mask).max() / ref.abs().max())
        untouched = bool(torch.equal(C[~mask], C0[~mask]))
    else:
        err = float((C.double() - ref).abs().max() / ref.abs().max()); untouched = True
    ms = timeit(lambda: ops.gemm_tf32x3_nt(X, Y, alpha=-1.0, beta=1.0, out=C, lower_only=lower, X_lo=Xl, Y_lo=Yl))
    fl = 2.0 * M * N * K * (0.5 if lower else 1.0)
    print('M=%d N=%d K=%d lower=%d: relerr %.2e upper untouched %s | %.3f ms  %.1f TFLOP/s (useful)' % (M, N, K, lower, err, untouched, ms, fl / ms / 1e9))
