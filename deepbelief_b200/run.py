"""Run one of the reference's scripts UNMODIFIED on the B200-native implementation:

    cd examples/dbn_horses && python -m deepbelief_b200.run train.py [script args...]

The reference's scripts start with `import tensorflow as tf` and `from deepbelief.layers... import ...`
(examples/dbn_horses/train.py:1-6) and load `flags.py` from the working directory (train.py:9-13). This launcher
registers `deepbelief_b200.compat` as `tensorflow`, makes the `deepbelief.*` alias package importable and executes the
script with runpy as `__main__` from its own directory — the script text itself is not touched.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        sys.stderr.write(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)                 # the `deepbelief` alias package sits next to `deepbelief_b200`
    from deepbelief_b200 import compat
    compat.install_script_environment()
    sys.argv = [script] + argv[1:]
    sys.path.insert(0, os.path.dirname(script))  # what `python script.py` would have put there
    runpy.run_path(script, run_name='__main__')
    return 0


if __name__ == '__main__':
    sys.exit(main())
