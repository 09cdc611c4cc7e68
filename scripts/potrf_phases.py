"""Phase breakdown of the Cholesky step kernel (clock64 in strip CTA 1): needs a library built with -DDB_POTRF_TIMING, e.g.
    DEEPBELIEF_B200_LIB=deepbelief_b200/libdeepbelief_b200_timing.so DB_POTRF_DEBUG=1 python scripts/potrf_phases.py
The dump of a factorisation is printed (stderr) when the next one starts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('DEEPBELIEF_B200_GRAPH', '0')
import torch
from deepbelief_b200 import ops
N = int(os.environ.get('N', 5000))
X = torch.randn(N, 2, device='cuda')
hyp = ops.gp_effective_hyp(torch.tensor([0.5413, 0.5413, 0.5413], device='cuda'), 7, 0.0, 1e-6)
K = ops.gram_rbf(X, None, hyp, add_noise=True)
A = K.clone()
for _ in range(3):
    A.copy_(K); ops.potrf_lower(A)
torch.cuda.synchronize()
