"""Data-parallel CD step on real GPUs (needs >= 2 devices on the box; skipped otherwise): the peer-push exchange
(db_cd_tc_step_push + db_cd_tc_apply_update_sharded over CUDA-IPC mapped buffers) against the NCCL all-reduce path,
run as `torch.distributed.run` with one rank per GPU (scripts/check_peer_push_mp.py holds the assertions:
replicas bit-identical, the two paths equal to fp32 summation order after one SGD update)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, has_cuda


def _device_count():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda() or _device_count() < 2, reason='needs at least two CUDA devices')
def test_peer_push_step_matches_nccl_step():
    world = min(_device_count(), 4)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                        '--master-addr', '127.0.0.1', '--master-port', '29571', os.path.join(ROOT, 'scripts', 'check_peer_push_mp.py')],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'PEER PUSH OK' in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
