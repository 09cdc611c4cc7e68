"""deepbelief_b200 — B200-native (sm_100a) implementation of the two hot paths of zeis/deepbelief.

The public boundary is the reference's Python layer API (`deepbelief_b200.layers.*`, mirrored under
the alias package `deepbelief`); the arithmetic runs in hand-written CUDA behind the C ABI declared
in `include/deepbelief_b200.h`.
"""
from . import config  # noqa: F401

__version__ = '0.1.0'
