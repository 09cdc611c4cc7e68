"""Host-side data preparation (reference: deepbelief/preprocessing.py). numpy, off the hot path;
only PCA / standardize / inverse_softplus feed the GPLVM initialisation (gplvm.py:128-132)."""
import numpy as np

from . import config as c


def binarize(array, threshold):
    """1 where array > threshold else 0, as float_type (preprocessing.py:7-22)."""
    binary = np.zeros_like(array, dtype=c.float_type)
    binary[array > threshold] = 1
    return binary


def scale255to1(array):
    """Scale [0, 255] values to [0, 1] (preprocessing.py:25-40)."""
    assert np.min(array) >= 0 and np.max(array) <= 255, 'The input array values should lie in range [0, 255].'
    return array.astype(c.float_type) / np.asarray(255.0, dtype=c.float_type)


def normalize(array):
    """In-place min-max normalisation to [0, 1]; a constant array becomes zeros (preprocessing.py:43-70)."""
    lo, hi = np.min(array), np.max(array)
    if hi - lo == 0:
        array -= hi
        return array
    array -= lo
    array *= 1.0 / (hi - lo)
    return array


def standardize(array, mean_fname=None, std_fname=None, return_means_stds=False):
    """Column-wise (x - mean) / std with population std; std <= 1e-8 is replaced by 1
    (preprocessing.py:73-119)."""
    assert array.ndim == 2, 'The array must be 2-dimensional'
    assert array.shape[0] > 1, 'The array should contain at least two rows'
    array = array.astype(c.float_type)
    means = array.mean(axis=0)
    stds = array.std(axis=0)
    stds[stds <= 1e-8] = 1.0
    array = (array - means) / stds
    if mean_fname:
        np.save(mean_fname, means)
    if std_fname:
        np.save(std_fname, stds)
    if return_means_stds:
        return array, means, stds
    return array


def unstandardize(tensor, mean_fname=None, std_fname=None, means=None, stds=None):
    """tensor * stds + means column-wise (preprocessing.py:122-154); evaluation-time only."""
    import torch
    assert type(means) is np.ndarray or mean_fname is not None
    assert type(stds) is np.ndarray or std_fname is not None
    np_means = np.load(mean_fname) if mean_fname else means
    np_stds = np.load(std_fname) if std_fname else stds
    t = tensor if isinstance(tensor, torch.Tensor) else torch.as_tensor(np.asarray(tensor, dtype=np.float32))
    return t * torch.as_tensor(np_stds, dtype=torch.float32, device=t.device) + \
        torch.as_tensor(np_means, dtype=torch.float32, device=t.device)


def PCA(X, num_components):
    """Projection on the leading eigenvectors of the scatter matrix (preprocessing.py:157-168)."""
    assert X.shape[1] >= num_components, \
        'The number of principal components must be less than or equal to the dimensionality of the input matrix'
    X_zero = X - X.mean(0)
    eigenvals, eigenvecs = np.linalg.eigh(X_zero.T.dot(X_zero))
    order = np.argsort(eigenvals)[::-1]
    return X_zero.dot(eigenvecs[:, order][:, :num_components])


def inverse_softplus(x):
    """log(exp(x) - 1) (preprocessing.py:171-172)."""
    return np.log(np.exp(x) - 1.0)
