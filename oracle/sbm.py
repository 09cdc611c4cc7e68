"""TEST INFRASTRUCTURE — CPU restatement of deepbelief/layers/sbm_lower.py (Shape-Boltzmann lower layer).

Only tests/ may import this module; the product path never does.

The four overlapping patches of the side x side image share one weight matrix W [P*P, H/4], P = side/2 + overlap/2
(sbm_lower.py:28-41). h_net_input slices the four patches, multiplies each by W and concatenates (sbm_lower.py:59-81,
no hidden bias); v_net_input multiplies the four hidden groups by W^T, zero-pads each back to its place in the image,
adds them and the visible bias (sbm_lower.py:83-117). Both are linear in v / h, so the layer equals an RBM with the
dense weight `dense_weights(W)` and c = 0; `SBMLowerOracle` is that RBM, with the gradient projected back onto the
shared matrix (sum over the four blocks) and no gradient for c.
Pinned by tests/golden/sbm_golden.npz (outputs of the reference class itself, tests/golden/make_golden.py).
"""
import numpy as np

from .rbm import RBMOracle


def patch_offsets(side, side_overlap):
    """(row offset, column offset) of patches 1..4 (sbm_lower.py:63-67) and the patch side."""
    P = int(side / 2 + side_overlap / 2)
    pad = P - side_overlap
    return P, [(0, 0), (0, pad), (pad, pad), (pad, 0)]


def dense_weights(W, side, side_overlap, num_h, dtype=np.float64):
    P, offs = patch_offsets(side, side_overlap)
    h4 = num_h // 4
    W = np.asarray(W, dtype=dtype).reshape(P * P, h4)
    dense = np.zeros((side * side, num_h), dtype=dtype)
    for p, (ro, co) in enumerate(offs):
        for r in range(P):
            for c in range(P):
                dense[(r + ro) * side + (c + co), p * h4:(p + 1) * h4] = W[r * P + c]
    return dense


def gather_gradient(g_dense, side, side_overlap, num_h):
    P, offs = patch_offsets(side, side_overlap)
    h4 = num_h // 4
    g = np.zeros((P * P, h4), dtype=g_dense.dtype)
    for p, (ro, co) in enumerate(offs):
        for r in range(P):
            for c in range(P):
                g[r * P + c] += g_dense[(r + ro) * side + (c + co), p * h4:(p + 1) * h4]
    return g


class SBMLowerOracle(RBMOracle):
    def __init__(self, W_shared, b, side, side_overlap, num_h, temperature=None, dtype=np.float64):
        self.side, self.side_overlap = side, side_overlap
        self.W_shared = np.asarray(W_shared, dtype=dtype)
        super().__init__(dense_weights(W_shared, side, side_overlap, num_h, dtype), b, np.zeros(num_h), kind='rbm',
                         temperature=temperature, dtype=dtype)

    def cd_grad_shared(self, v0, vk):
        """(dW_shared, db, dc = 0): what TF autodiff yields for the reference's variables."""
        gW, gb, gc, _ = self.cd_grad(v0, vk)
        return gather_gradient(gW, self.side, self.side_overlap, self.H), gb, np.zeros_like(gc)
