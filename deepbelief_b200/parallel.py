"""Data-parallel contrastive divergence: host-side plumbing (new work — the reference is single
device).

Rows of the minibatch are independent given (W, b, c), so rank g owns rows
[g*N/G, (g+1)*N/G) of every batch; parameters and optimizer state are replicated. The only exchange
per iteration is a sum-all-reduce of the packed gradient [dW | db | dc], each rank having scaled its
local sums by 1/N_global, after which every rank applies the identical optimizer step (replicas stay
bit-identical). Philox counters are keyed by the GLOBAL row (`row_offset`), so the samples do not
depend on the number of GPUs.
"""
import os

import torch
import torch.distributed as td


def dist_info():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1 or (td.is_available() and td.is_initialized()):
        return dist_info()
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    td.init_process_group(backend=backend)
    return dist_info()


def shard_bounds(n_global, rank, world):
    """Row range [lo, hi) of `rank`: equal shards, N_global must divide evenly (keeps Philox rows aligned)."""
    if n_global % world:
        raise ValueError('global batch %d is not divisible by %d ranks' % (n_global, world))
    per = n_global // world
    return rank * per, (rank + 1) * per


def all_reduce_sum_(tensor):
    """In-place sum over the data-parallel group (NCCL on GPUs, gloo in the CPU tests)."""
    _, world = dist_info()
    if world > 1:
        td.all_reduce(tensor, op=td.ReduceOp.SUM)
    return tensor


def max_over_ranks(value, device=None):
    """Max of a python float over ranks (timing: a step is as slow as its slowest rank)."""
    _, world = dist_info()
    if world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device or ('cuda' if td.get_backend() == 'nccl' else 'cpu'))
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


class PeerBuffers:
    """One device tensor per rank of the node, every one of them mapped into THIS process (CUDA IPC), so that a kernel can store
    straight into a peer's memory over NVLink (the dW epilogue's reduce-scatter by push, csrc/gemm_tc.cu `peer_out`, and the
    sharded update's all-gather by push, `apply_shard_kernel`). All ranks must run on one node and see every GPU.

    .ptrs: ctypes array of the world device pointers (this rank's own tensor for r == rank), valid for kernels launched on this
    rank's device. The handles are opened through the library (db_ipc_import: in this rank's OWN device context with lazy peer
    access) — a tensor rebuilt by torch.multiprocessing lives in the exporting device's context and raw kernels of this
    device fault on it."""

    def __init__(self, tensor):
        """Collective over the default group. Raises RuntimeError on EVERY rank when the mapping failed on any of them (no
        peer path between two devices, IPC not permitted in this container, ...), so callers can fall back together."""
        import ctypes
        from . import _lib as L
        rank, world = dist_info()
        assert world > 1 and tensor.is_cuda and tensor.is_contiguous()
        self.tensor = tensor                                            # keeps the exported allocation alive
        handle = ctypes.create_string_buffer(64)
        offset = ctypes.c_uint64(0)
        err = None
        try:
            L.check(L.lib().db_ipc_export(L.handle(), ctypes.c_void_p(int(tensor.data_ptr())), handle, ctypes.byref(offset)),
                    'db_ipc_export')
        except Exception as e:                                          # still take part in the collectives below
            err = e
        gathered = [None] * world
        td.all_gather_object(gathered, (bytes(handle.raw), int(offset.value), int(tensor.device.index), err is None))
        ptrs = []
        if err is None and all(g[3] for g in gathered):
            try:
                probe = torch.empty(1, dtype=torch.float32, device=tensor.device)
                for r, (hb, off, dev_index, _) in enumerate(gathered):
                    if r == rank:
                        ptrs.append(int(tensor.data_ptr()))
                        continue
                    L.check(L.lib().db_enable_peer_access(L.handle(), dev_index), 'db_enable_peer_access')
                    out = ctypes.c_void_p()
                    L.check(L.lib().db_ipc_import(L.handle(), ctypes.create_string_buffer(hb, 64), ctypes.c_uint64(off),
                                                  ctypes.byref(out)), 'db_ipc_import')
                    ptrs.append(int(out.value))
                    if tensor.dtype == torch.float32:                   # one raw-kernel read of the peer pointer: fails loudly here
                        L.check(L.lib().db_softplus(L.handle(), L.stream(), ctypes.c_void_p(ptrs[-1]), 1, L.fptr(probe)), 'peer read')
                        torch.cuda.synchronize(tensor.device)
            except Exception as e:
                err = e
        elif err is None:
            err = RuntimeError('another rank could not export its buffer')
        ok = torch.tensor([0.0 if err is not None else 1.0], device=tensor.device)
        td.all_reduce(ok, op=td.ReduceOp.MIN)                           # also the barrier: every mapping is in place
        if float(ok.item()) < 1.0:
            raise RuntimeError('peer mapping of the data-parallel buffers failed on at least one rank: %s' % (err,))
        self.ptrs = (ctypes.c_void_p * world)(*ptrs)
