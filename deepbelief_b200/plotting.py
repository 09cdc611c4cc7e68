"""Plotters are OUT OF SCOPE: the reference's matplotlib visualisation is not
part of the hot path and matplotlib is not available. The classes below accept the constructor
arguments the example scripts pass and record what they are given, so the scripts run headless;
`set_data` / `plot` write the arrays as .npy next to where the figures would have gone."""
import os

import numpy as np


class _Recorder:
    def __init__(self, output_dir=None, **kwargs):
        self.output_dir = output_dir
        self.options = kwargs
        self.last = None

    def _dump(self, name, **arrays):
        self.last = arrays
        if self.output_dir and os.path.isdir(self.output_dir):
            np.savez(os.path.join(self.output_dir, str(name) + '.npz'),
                     **{k: np.asarray(v) for k, v in arrays.items() if v is not None})


class GraphPlotter(_Recorder):
    def __init__(self, xlim=None, ylim=None, xlabel=None, ylabel=None, output_dir=None, **kw):
        super().__init__(output_dir, xlim=xlim, ylim=ylim, xlabel=xlabel, ylabel=ylabel, **kw)

    def set_data(self, x_vals, y_vals, fig_name):
        self._dump(fig_name, x=x_vals, y=y_vals)


class ImageRowPlotter(_Recorder):
    def __init__(self, num_images=None, image_shape=None, output_dir=None, **kw):
        super().__init__(output_dir, num_images=num_images, image_shape=image_shape, **kw)

    def set_data(self, images, fig_name):
        self._dump(fig_name, **{'image_%d' % i: im for i, im in enumerate(images)})


class ImageGridPlotter(_Recorder):
    def set_data(self, images, fig_name):
        self._dump(fig_name, images=images)


def _latent_grid(xlim, ylim, delta, yflipud=False):
    """The evaluation grid of the reference's latent-space plotters (plotting.py:174-184): every (x, y) with
    x in xlim, y in ylim on a `delta` lattice, x fastest; the sample plotter walks y from the top."""
    x = np.arange(xlim[0], xlim[1] + delta, delta)
    y = np.arange(ylim[0], ylim[1] + delta, delta)
    xx, yy = np.meshgrid(x, y)
    xx = xx.ravel()[:, None]
    yy = (np.flipud(yy.ravel()) if yflipud else yy.ravel())[:, None]
    return np.hstack([xx, yy]).astype(np.float32)


class LatentPointPlotter(_Recorder):
    """Constructor arguments as the reference (plotting.py:189-199): xlim, ylim, delta build the grid on which the layers
    evaluate the predictive variance; a ready-made `grid` may be passed instead."""

    def __init__(self, xlim=None, ylim=None, delta=None, grid=None, output_dir=None, **kw):
        super().__init__(output_dir, xlim=xlim, ylim=ylim, delta=delta, **kw)
        if grid is None and xlim is not None and ylim is not None and delta:
            grid = _latent_grid(xlim, ylim, delta)
        self.grid = grid

    def plot(self, X, log_var, fig_name, labels=None):
        self._dump(fig_name, X=X, log_var=log_var, labels=labels)


class LatentSamplePlotter(_Recorder):
    """plotting.py:462-483: grid on a `delta` lattice (default 0.5), y descending."""

    def __init__(self, image_shape=None, xlim=None, ylim=None, delta=0.5, grid=None, num_samples_to_average=100,
                 output_dir=None, **kw):
        super().__init__(output_dir, image_shape=image_shape, xlim=xlim, ylim=ylim, delta=delta, **kw)
        if grid is None and xlim is not None and ylim is not None and delta:
            grid = _latent_grid(xlim, ylim, delta, yflipud=True)
        self.grid = grid
        self.num_samples_to_average = num_samples_to_average

    def plot(self, v_vecs, fig_name):
        self._dump(fig_name, samples=v_vecs)
