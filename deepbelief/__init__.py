"""Drop-in alias: `import deepbelief...` resolves to the B200-native implementation.

The reference's example scripts import `deepbelief.layers.rbm`, `deepbelief.layers.data`,
`deepbelief.plotting`, `deepbelief.util`, ... (e.g. examples/dbn_horses/train.py:1-6). This package
maps every one of those module paths onto `deepbelief_b200`, so a script only has to replace its
`import tensorflow as tf` line with `from deepbelief_b200 import compat as tf`.
"""
import importlib
import sys

_ALIASES = {
    'deepbelief.config': 'deepbelief_b200.config',
    'deepbelief.dist': 'deepbelief_b200.dist',
    'deepbelief.preprocessing': 'deepbelief_b200.preprocessing',
    'deepbelief.util': 'deepbelief_b200.util',
    'deepbelief.plotting': 'deepbelief_b200.plotting',
    'deepbelief.layers': 'deepbelief_b200.layers',
    'deepbelief.layers.layer': 'deepbelief_b200.layers.layer',
    'deepbelief.layers.data': 'deepbelief_b200.layers.data',
    'deepbelief.layers.rbm': 'deepbelief_b200.layers.rbm',
    'deepbelief.layers.bgrbm': 'deepbelief_b200.layers.bgrbm',
    'deepbelief.layers.gbrbm': 'deepbelief_b200.layers.gbrbm',
    'deepbelief.layers.ggrbm': 'deepbelief_b200.layers.ggrbm',
    'deepbelief.layers.bgrbmsv': 'deepbelief_b200.layers.bgrbmsv',
    'deepbelief.layers.sbm_lower': 'deepbelief_b200.layers.sbm_lower',
    'deepbelief.layers.gplvm': 'deepbelief_b200.layers.gplvm',
    'deepbelief.layers.gprbm': 'deepbelief_b200.layers.gprbm',
}


class _AliasFinder:
    """Resolves `deepbelief.<x>` lazily so importing the alias does not load CUDA-dependent modules."""

    @staticmethod
    def find_spec(name, path=None, target=None):
        real = _ALIASES.get(name)
        if real is None:
            return None
        module = importlib.import_module(real)
        sys.modules[name] = module
        return importlib.util.spec_from_loader(name, loader=_AliasLoader(module))


class _AliasLoader:
    def __init__(self, module):
        self._module = module

    def create_module(self, spec):
        return self._module

    def exec_module(self, module):
        pass


import importlib.util  # noqa: E402

sys.meta_path.insert(0, _AliasFinder)
