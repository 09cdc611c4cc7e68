"""Bernoulli-visible / Gaussian-hidden RBM with ONE shared sigma_h — counterpart of deepbelief/layers/bgrbmsv.py.

The reference creates a single trainable opt_sigma_h of shape [1], initialised to 1.0 so that sigma_h = softplus(1)
(bgrbmsv.py:45-51), or takes sigma_h from a placeholder fed from outside (bgrbmsv.py:52-53); everything else is BGRBM with
that value broadcast over the hidden units. Here the packed parameter buffer keeps the BGRBM layout
[W | b | c | opt_sigma_h (H)] with all H entries EQUAL: the kernels need nothing new, and the gradient of the shared scalar
— the sum of the per-unit gradients, because every unit sees the same value — is written back to all H slots before the
optimizer runs, so identical values, identical slots and identical gradients keep the entries equal under SGD, Momentum and
Adam.
"""
import numpy as np
import torch

from .. import preprocessing
from . import bgrbm
from .layer import to_device


class BGRBMSV(bgrbm.BGRBM):
    def __init__(self, num_v, num_h, W_std=0.001, lr=0.001, sigma_h_ph=None, temperature=None, bottom=None,
                 loss_plotter=None, distr_plotter=None, filter_plotter=None, latent_sample_plotter=None,
                 latent_space_explorer=None, session=None, name='BGRBMSV', seed=None, device='cuda'):
        self._sigma_fed = sigma_h_ph is not None
        super().__init__(num_v=num_v, num_h=num_h, W_std=W_std, lr=lr, temperature=temperature, bottom=bottom,
                         loss_plotter=loss_plotter, distr_plotter=distr_plotter, filter_plotter=filter_plotter,
                         latent_sample_plotter=latent_sample_plotter, latent_space_explorer=latent_space_explorer,
                         session=session, name=name, seed=seed, device=device)
        if self._sigma_fed:
            self.feed_sigma_h(sigma_h_ph)
        else:
            self.opt_sigma_h.fill_(1.0)                     # tf.ones(shape=1) is the OPTIMISED variable (bgrbmsv.py:46-51)
            self._weights_version += 1
        self.params['sigma_h'] = lambda: self.sigma_h[:, :1].reshape(1)

    def feed_sigma_h(self, value):
        """The value a feed_dict would give the sigma_h placeholder (bgrbmsv.py:52-53): one positive scalar."""
        v = np.asarray(value.detach().cpu() if isinstance(value, torch.Tensor) else value, dtype=np.float64).reshape(-1)
        assert v.size == 1 and v[0] > 0, 'the shared sigma_h is one positive scalar'
        self.opt_sigma_h.fill_(float(preprocessing.inverse_softplus(v[0])))
        self._weights_version += 1

    def set_weights(self, W=None, b=None, c=None, opt_sigma_h=None):
        if opt_sigma_h is not None:
            v = to_device(opt_sigma_h, self.device).reshape(-1)
            assert v.numel() == 1, 'opt_sigma_h is a single shared value'
            opt_sigma_h = v.expand(self.num_h).contiguous()
        super().set_weights(W, b, c, opt_sigma_h=opt_sigma_h)

    def _trainable_pair(self, grad):
        """The optimizer's (parameters, gradient) pair; the sigma slots of a real gradient buffer are made those of the
        shared scalar first. (train() also calls this with the parameter buffer itself, to size and snapshot state.)"""
        if grad.data_ptr() != self._params.data_ptr():
            off = self.num_v * self.num_h + self.num_v + self.num_h
            seg = grad[off:off + self.num_h]
            if self._sigma_fed:
                seg.zero_()                                  # fed from outside: not a variable of this layer
            else:
                seg.copy_(seg.sum().expand(self.num_h))      # d/d(shared scalar) = sum of the per-unit gradients
        return self._params, grad
