"""Shape-Boltzmann-machine lower layer — counterpart of deepbelief/layers/sbm_lower.py.

Four overlapping patches of the side x side visible image share ONE weight matrix W [patch_side^2, num_h/4]
(sbm_lower.py:28-41); patch p feeds hidden units [p num_h/4, (p+1) num_h/4). The reference expresses this with
tf.slice / tf.matmul / tf.concat (h_net_input, :59-81) and tf.split / tf.pad / add (v_net_input, :83-117). Here the
layer is an RBM whose dense kernel weight is the shared matrix scattered four times (db_sbm_expand), so the CD chain,
sampling, free energy and the tensor-core path run unchanged; the optimizer sees the shared parameters and the sum of
the four scattered gradient blocks (db_sbm_gather).

Reference quirk kept (SURVEY.md §8 note 14): h_net_input never adds the hidden bias c (sbm_lower.py:74-81), so c takes
no part in any loss, TF computes no gradient for it and it stays at its initial zeros.
"""
import numpy as np
import torch

from .. import ops
from . import rbm
from . import layer
from .layer import to_device


class SBM_Lower(rbm.RBM):
    def __init__(self, side, side_overlap, num_h, W_std=0.001, lr=0.001, temperature=None, bottom=None,
                 loss_plotter=None, distr_plotter=None, filter_plotter=None, latent_sample_plotter=None,
                 latent_space_explorer=None, session=None, name='SBM_Lower', seed=None, device='cuda'):
        assert num_h % 4 == 0                                   # sbm_lower.py:24-26
        assert side % 2 == 0
        assert side_overlap % 2 == 0
        self.side_overlap = side_overlap
        self.side = side
        self.patch_side = int(side / 2 + side_overlap / 2)      # sbm_lower.py:32
        self._patch_side_squared = self.patch_side * self.patch_side
        super().__init__(num_v=side * side, num_h=num_h, W_std=W_std, lr=lr, temperature=temperature, bottom=bottom,
                         loss_plotter=loss_plotter, distr_plotter=distr_plotter, filter_plotter=filter_plotter,
                         latent_sample_plotter=latent_sample_plotter, latent_space_explorer=latent_space_explorer,
                         session=session, name=name, seed=seed, device=device)
        # shared (trainable) parameters [W | b | c]; the packed buffer of the base class stays the kernel operand
        h4 = num_h // 4
        nw = self._patch_side_squared * h4
        self._shared = torch.zeros(nw + self.num_v + self.num_h, dtype=torch.float32, device=device)
        self._shared_grad = torch.zeros_like(self._shared)
        self.W = self._shared[:nw].view(self._patch_side_squared, h4)
        self.b = self._shared[nw:nw + self.num_v].view(self.num_v, 1)
        self.c = self._shared[nw + self.num_v:].view(1, self.num_h)
        w0 = np.random.normal(size=(self._patch_side_squared, h4))          # truncated normal (sbm_lower.py:36-39)
        bad = np.abs(w0) > 2.0
        while bad.any():
            w0[bad] = np.random.normal(size=int(bad.sum()))
            bad = np.abs(w0) > 2.0
        self.W.copy_(torch.as_tensor((w0 * W_std).astype(np.float32)))
        self.params = {'W': self.W, 'b': self.b, 'c': self.c}
        self._expand()

    # ---- shared <-> dense ----------------------------------------------------------------------------------------
    def _expand(self):
        ops.sbm_expand(self.W, self.side, self.side_overlap, self.num_h, self.Wk)
        self.bk.copy_(self.b)
        self.ck.zero_()                                          # h_net_input has no hidden bias (sbm_lower.py:74-81)
        self._weights_version += 1

    def _trainable_pair(self, grad):
        vh = self.num_v * self.num_h
        nw = self.W.numel()
        ops.sbm_gather(grad[:vh].view(self.num_v, self.num_h), self.side, self.side_overlap, self.num_h,
                       self._shared_grad[:nw].view_as(self.W))
        self._shared_grad[nw:nw + self.num_v].copy_(grad[vh:vh + self.num_v])
        self._shared_grad[nw + self.num_v:].zero_()              # c is not part of the graph: no gradient
        return self._shared, self._shared_grad

    def _after_update(self):
        self._expand()

    def set_weights(self, W=None, b=None, c=None):
        if W is not None:
            self.W.copy_(to_device(W, self.device).reshape(self.W.shape))
        if b is not None:
            self.b.copy_(to_device(b, self.device).reshape(self.num_v, 1))
        if c is not None:
            self.c.copy_(to_device(c, self.device).reshape(1, self.num_h))
        self._expand()

    # ---- state ---------------------------------------------------------------------------------------------------
    def state_tensors(self):
        return {'params': self._shared}

    def save(self, i, ckpt_dir):
        import os
        assert os.path.isdir(ckpt_dir), 'The directory does not exist: %s' % ckpt_dir
        path = os.path.join(ckpt_dir, '%s-%d' % (self.name, i))
        layer.write_checkpoint({'params': self._shared.detach().cpu(), 'ckpt_iter': self.checkpoint_iter, 'lr': self.lr}, path)
        self.logger.info('State saved')
        return path

    def restore(self, ckpt_fname):
        state = torch.load(ckpt_fname, map_location='cpu')
        self._shared.copy_(state['params'].to(self._shared.device))
        self.checkpoint_iter = int(state.get('ckpt_iter', 1))
        self.lr = float(state.get('lr', self.lr))
        self._expand()

    def cd_gradient(self, v_0_batch, v_k_batch, inv_global_n=None):
        """Gradient of cd_loss with respect to the shared parameters, packed [W | b | c]."""
        dense = super().cd_gradient(v_0_batch, v_k_batch, inv_global_n)
        return self._trainable_pair(dense)[1].clone()
