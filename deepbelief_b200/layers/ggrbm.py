"""Gaussian-Gaussian RBM — counterpart of deepbelief/layers/ggrbm.py: linear visibles (GBRBM) and
Gaussian hiddens with the BGRBM energy, by the same multiple inheritance GGRBM(GBRBM, BGRBM)
(ggrbm.py:5). The reference constructor takes no `temperature` (ggrbm.py:26-38)."""
from .. import _lib as L
from . import bgrbm, gbrbm


class GGRBM(gbrbm.GBRBM, bgrbm.BGRBM):
    layer_kind = L.LAYER_GGRBM

    def __init__(self, num_v, num_h, W_std=0.001, lr=0.001, bottom=None, loss_plotter=None, distr_plotter=None,
                 filter_plotter=None, latent_sample_plotter=None, latent_space_explorer=None, session=None,
                 name='GGRBM', seed=None, device='cuda'):
        super().__init__(num_v=num_v, num_h=num_h, W_std=W_std, lr=lr, bottom=bottom, loss_plotter=loss_plotter,
                         distr_plotter=distr_plotter, filter_plotter=filter_plotter,
                         latent_sample_plotter=latent_sample_plotter, latent_space_explorer=latent_space_explorer,
                         session=session, name=name, seed=seed, device=device)
