"""Data-parallel CD host logic on CPU: world_size-2 gloo. Each rank computes the
gradient SUMS of its row shard scaled by 1/N_global (oracle arithmetic stands in for the kernels);
the sum-all-reduce must reproduce the full-batch mean gradient, and the Philox draws keyed by global
row must not depend on the number of ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from deepbelief_b200 import parallel
from oracle import philox, rbm as orbm


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    assert parallel.init_from_env(backend='gloo') == (rank, world)
    rs = np.random.RandomState(5)
    N, V, H = 16, 12, 8
    W, b, c = rs.normal(size=(V, H)) * 0.3, rs.normal(size=V) * 0.1, rs.normal(size=H) * 0.1
    v0 = rs.binomial(1, 0.5, size=(N, V)).astype(np.float64)
    m = orbm.RBMOracle(W, b, c)
    lo, hi = parallel.shard_bounds(N, rank, world)
    u = philox.uniform(hi - lo, H, seed=11, draw=0, row_offset=lo)              # this rank's rows only
    h0 = m.sample_h(m.h_probs(v0[lo:hi]), u)
    uv = philox.uniform(hi - lo, V, seed=11, draw=1, row_offset=lo)
    vk = m.sample_v(m.v_probs(h0), uv)
    local = m.cd_grad_packed(v0[lo:hi], vk) * ((hi - lo) / N)                   # local sums / N_global
    g = torch.from_numpy(local.copy())
    parallel.all_reduce_sum_(g)
    t = parallel.max_over_ranks(float(rank + 1))
    if rank == 0:
        results.put((g.numpy(), t))
    td.barrier()
    td.destroy_process_group()


def test_sharded_gradient_allreduce_equals_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    results = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, results)) for r in range(world)]
    for p in procs:
        p.start()
    got, tmax = results.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert tmax == 2.0
    # single-process reference over the full batch with the same global-row-keyed draws
    rs = np.random.RandomState(5)
    N, V, H = 16, 12, 8
    W, b, c = rs.normal(size=(V, H)) * 0.3, rs.normal(size=V) * 0.1, rs.normal(size=H) * 0.1
    v0 = rs.binomial(1, 0.5, size=(N, V)).astype(np.float64)
    m = orbm.RBMOracle(W, b, c)
    h0 = m.sample_h(m.h_probs(v0), philox.uniform(N, H, seed=11, draw=0))
    vk = m.sample_v(m.v_probs(h0), philox.uniform(N, V, seed=11, draw=1))
    np.testing.assert_allclose(got, m.cd_grad_packed(v0, vk), rtol=1e-12, atol=1e-14)


def test_shard_bounds():
    assert parallel.shard_bounds(16384, 3, 8) == (6144, 8192)
    assert parallel.dist_info() == (0, 1)
    try:
        parallel.shard_bounds(10, 0, 4)
        assert False
    except ValueError:
        pass
