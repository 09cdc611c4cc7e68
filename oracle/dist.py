"""Oracle restatement of deepbelief/dist.py (TEST INFRASTRUCTURE — see oracle/__init__.py)."""
import numpy as np


def mse(predictions, targets):
    """dist.py:5-21: mean over rows of the SUM over columns."""
    return np.mean(np.sum(np.square(predictions - targets), axis=1))


def cross_entropy(probs, targets, reduce_mean=True):
    """dist.py:24-35: clip to [1e-6, 1-1e-6], logit, sigmoid cross entropy with logits [TF]."""
    e0 = probs.dtype.type(10e-7)
    p = np.clip(probs, e0, 1 - e0)
    x = np.log(p / (1 - p))
    ce = np.maximum(x, 0) - x * targets + np.log1p(np.exp(-np.abs(x)))
    if reduce_mean:
        return np.mean(np.sum(ce, axis=1))
    return np.sum(ce)
