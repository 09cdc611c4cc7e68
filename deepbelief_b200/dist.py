"""Reconstruction metrics (reference: deepbelief/dist.py), computed by CUDA reduction kernels."""
from . import ops
from .layers.layer import to_device


def mse(predictions, targets):
    """Mean over datapoints of the SUM of squared errors over dimensions (dist.py:5-21)."""
    return ops.mse(to_device(predictions), to_device(targets))


def cross_entropy(probs, targets, reduce_mean=True):
    """Sigmoid cross entropy of clipped probabilities (dist.py:24-35)."""
    return ops.cross_entropy(to_device(probs), to_device(targets), reduce_mean)
