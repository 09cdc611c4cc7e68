"""The GP chain with its triangular inverse on the side lanes (csrc/gp.cu, TrinvOverlap) against the same chain on one stream,
eager and graph-replayed. DB_GP_OVERLAP is read once per process, so every variant runs in its own interpreter
(scripts/exp_overlap.py). Covers the CUDA-graph finding documented in DESIGN.md (K5): a dependent launch right behind a
cross-stream event wait lost that edge once captured — replays then differed from the eager run."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(mode, n, d, ref_dir):
    env = dict(os.environ, DB_GP_OVERLAP=str(mode), REPEAT='1', OVERLAP_REF_DIR=str(ref_dir))
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'exp_overlap.py'), str(n), str(d)], env=env, cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


@pytest.mark.parametrize('n,d', [(2048, 256), (5000, 784), (6144, 64)])
def test_side_lanes_match_one_stream_eager_and_replayed(tmp_path, n, d):
    base = _run(0, n, d, tmp_path)
    assert 'info 0' in base
    for mode in (1, 2):
        text = _run(mode, n, d, tmp_path)
        assert 'info 0' in text, text
        g_vs_e = float(re.search(r'graph-vs-eager dX ([0-9.e+-]+)', text).group(1))
        sc, dx = re.search(r'vs mode 0: scalars ([0-9.e+-]+) dX ([0-9.e+-]+)', text).groups()
        spread = [float(x) for x in re.search(r'replay spread \[([^\]]*)\]', text).group(1).split(',')]
        # atomics in the reductions: two runs of the SAME launch sequence agree to ~5e-7; a lost dependency showed 1e-2
        assert g_vs_e < 5e-6, text
        assert max(spread) < 5e-6, text
        assert float(sc) < 1e-5 and float(dx) < 5e-6, text
