"""Bernoulli-visible / Gaussian-hidden RBM — counterpart of deepbelief/layers/bgrbm.py.

Extra parameter opt_sigma_h [1,H] with sigma_h = softplus(opt_sigma_h), initialised so sigma_h = 1
(bgrbm.py:44-51). Hidden mean = sigma_h * (v W) + c (bgrbm.py:69-82), which is also `h_probs`
(bgrbm.py:94); hidden sample = mean + sigma_h * eps (bgrbm.py:96-114); visible input uses h / sigma_h
(bgrbm.py:116-131); free energy bgrbm.py:133-153. All of it runs in the same fused CUDA launches as
the binary layer, selected through scale operands of db_layer_forward.
"""
import numpy as np

from .. import _lib as L
from .. import ops, preprocessing
from . import rbm
from .layer import to_device


class BGRBM(rbm.RBM):
    layer_kind = L.LAYER_BGRBM
    hidden_act = L.ACT_LINEAR

    def __init__(self, num_v, num_h, W_std=0.001, lr=0.001, temperature=None, bottom=None, loss_plotter=None,
                 distr_plotter=None, filter_plotter=None, latent_sample_plotter=None, latent_space_explorer=None,
                 session=None, name='BGRBM', seed=None, device='cuda'):
        super().__init__(num_v=num_v, num_h=num_h, W_std=W_std, lr=lr, temperature=temperature, bottom=bottom,
                         loss_plotter=loss_plotter, distr_plotter=distr_plotter, filter_plotter=filter_plotter,
                         latent_sample_plotter=latent_sample_plotter, latent_space_explorer=latent_space_explorer,
                         session=session, name=name, seed=seed, device=device)
        self.params['sigma_h'] = lambda: self.sigma_h

    def _init_extra_params(self):
        off = self.num_v * self.num_h + self.num_v + self.num_h
        if not hasattr(self, '_sigma_override'):
            self._sigma_override = None
        self.opt_sigma_h = self._params[off:off + self.num_h].view(1, self.num_h)
        self.opt_sigma_h.fill_(float(preprocessing.inverse_softplus(np.float64(1.0))))
        self._sigma_cache = (None, -1)

    @property
    def sigma_h(self):
        """softplus(opt_sigma_h) [1,H] — or the externally supplied per-datapoint sigma [N,H] (GPRBM)."""
        if self._sigma_override is not None:
            return self._sigma_override
        cached, version = self._sigma_cache
        if version != self._weights_version or cached is None:
            cached = ops.softplus(self.opt_sigma_h)
            self._sigma_cache = (cached, self._weights_version)
        return cached

    def set_weights(self, W=None, b=None, c=None, opt_sigma_h=None):
        if opt_sigma_h is not None:
            self.opt_sigma_h.copy_(to_device(opt_sigma_h, self.device).reshape(1, self.num_h))
        super().set_weights(W, b, c)

    def _sigma_vector(self):
        return self.sigma_h

    def _h_scale(self):
        return self.sigma_h

    def _v_div(self):
        return self.sigma_h

    def _hidden_sampling(self, sample):
        # Gaussian hiddens ignore temperature; sample=False returns the mean (bgrbm.py:106-113)
        return (L.SAMPLE_GAUSSIAN, self.sigma_h) if sample else (L.SAMPLE_NONE, None)

    def sample_h(self, h_probs, sample=True, summary=True, *, draw=None):
        """mean + sigma_h * eps (bgrbm.py:96-114)."""
        return rbm.RBM.sample_h(self, h_probs, sample, summary, draw=draw)
