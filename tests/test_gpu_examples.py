"""The example scripts (the reference's dbn / gplvm / gpdbn horses workflows written against the `deepbelief.*` module
paths) run end to end for a few iterations on the GPU and leave their checkpoints behind."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason='needs a CUDA device')]


@pytest.mark.parametrize('script,iters', [('dbn_horses_train.py', 6), ('gplvm_horses_train.py', 8), ('gpdbn_horses_train.py', 6)])
def test_example_runs(script, iters, tmp_path):
    env = dict(os.environ, EXAMPLE_ITERS=str(iters), EXAMPLE_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'examples', script)], cwd=os.path.join(ROOT, 'examples'), env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    produced = [f for _, _, files in os.walk(str(tmp_path)) for f in files]
    assert any('-%d' % iters in f for f in produced), produced          # <name>-<iteration> checkpoints (layer.py:61-72)
