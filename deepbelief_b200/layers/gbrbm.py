"""Gaussian-visible (unit variance) / Bernoulli-hidden RBM — counterpart of deepbelief/layers/gbrbm.py.

`v_probs` is the linear mean h W^T + b (gbrbm.py:56-57) and `sample_v` returns that mean unchanged:
the reference adds no noise and ignores its `sample` argument (gbrbm.py:59-64).
"""
from .. import _lib as L
from . import rbm
from .layer import to_device


class GBRBM(rbm.RBM):
    layer_kind = L.LAYER_GBRBM
    visible_act = L.ACT_LINEAR

    def __init__(self, num_v, num_h, W_std=0.001, lr=0.001, temperature=None, bottom=None, loss_plotter=None,
                 distr_plotter=None, filter_plotter=None, latent_sample_plotter=None, latent_space_explorer=None,
                 session=None, name='GBRBM', seed=None, device='cuda'):
        super().__init__(num_v=num_v, num_h=num_h, W_std=W_std, lr=lr, temperature=temperature, bottom=bottom,
                         loss_plotter=loss_plotter, distr_plotter=distr_plotter, filter_plotter=filter_plotter,
                         latent_sample_plotter=latent_sample_plotter, latent_space_explorer=latent_space_explorer,
                         session=session, name=name, seed=seed, device=device)

    def _visible_sampling(self, sample):
        return L.SAMPLE_NONE, None

    def sample_v(self, v_probs, sample=None, summary=True, *, draw=None):
        return to_device(v_probs, self.device)
