"""Parameter ownership, checkpointing and the bottom-linked layer stack.

Mirrors the conventions of deepbelief/layers/layer.py (reference): a layer owns named parameters,
`save(i, dir)` writes `<dir>/<name>-<i>` (layer.py:61-68), `restore(fname)` reads it back
(layer.py:70-72), `Layer.layers` lists the stack from this layer down (layer.py:84-90) and `save_all`
walks it (layer.py:98-100). Parameters here are fp32 CUDA `torch.Tensor`s owned by the layer (the
reference's were TF variables owned by a session); checkpoints are `torch.save` dictionaries.
"""
import json
import logging
import os

import numpy as np
import torch


def to_device(x, device='cuda'):
    """numpy / torch input -> contiguous fp32 CUDA tensor (the dtype contract of config.float_type)."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).to(device)


def to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def write_checkpoint(state, path):
    """torch.save through a temporary file + rename, from rank 0 only: under torch.distributed every rank holds the same
    replicated state and would otherwise write the same path concurrently (non-atomic)."""
    from ..parallel import dist_info
    if dist_info()[0] != 0:
        return path
    tmp = '%s.tmp.%d' % (path, os.getpid())
    torch.save(state, tmp)
    os.replace(tmp, path)
    return path


class Stateful:
    def __init__(self, params=None, session=None, name=None):
        self.params = params if params is not None else {}
        self.session = session
        self.name = name
        self.min_array_size_for_stats = 10
        self.logger = logging.getLogger(self.name)
        self.summary_writer = None
        self.summary = None

    # subclasses list every tensor that belongs in a checkpoint (trainable or not)
    def state_tensors(self):
        return {k: v for k, v in self.params.items() if isinstance(v, torch.Tensor)}

    def __str__(self):
        return json.dumps(self.param_dict(), indent=4, sort_keys=True)

    def param_dict(self):
        """Scalar parameters verbatim, arrays as min/max/avg/std statistics (layer.py:27-51)."""
        d = {}
        for key, val in self.params.items():
            arr = to_numpy(val() if callable(val) else val)
            if arr.size == 1:
                d[key] = float(arr.reshape(-1)[0])
            elif arr.size > self.min_array_size_for_stats:
                d[key + '_avg'] = float(np.mean(arr))
                d[key + '_min'] = float(np.min(arr))
                d[key + '_max'] = float(np.max(arr))
                d[key + '_std'] = float(np.std(arr))
                d[key + '_min_abs'] = float(np.min(np.abs(arr)))
            else:
                d[key] = arr.tolist()
        return d

    def init_variables(self):
        """Parameters are initialised at construction; kept for API parity (layer.py:53-59)."""

    def save(self, i, ckpt_dir):
        assert os.path.isdir(ckpt_dir), 'The directory does not exist: %s' % ckpt_dir
        ckpt_file = os.path.join(ckpt_dir, '%s-%d' % (self.name, i))
        write_checkpoint({k: v.detach().cpu() for k, v in self.state_tensors().items()}, ckpt_file)
        self.logger.info('State saved')
        return ckpt_file

    def restore(self, ckpt_fname):
        state = torch.load(ckpt_fname, map_location='cpu')
        mine = self.state_tensors()
        for k, v in state.items():
            if k in mine:
                mine[k].copy_(v.to(mine[k].device).reshape(mine[k].shape))
        self._after_restore()

    def _after_restore(self):
        pass


class Layer(Stateful):
    def __init__(self, params=None, bottom=None, session=None, name='Layer'):
        super().__init__(params=params, session=session, name=name)
        self.bottom = bottom
        layer = self
        self.layers = [layer]
        while layer.bottom:
            self.layers = self.layers + [layer.bottom]
            layer = layer.bottom

    def layer_zero(self):
        layer = self
        while layer.bottom:
            layer = layer.bottom
        return layer

    def save_all(self, i, ckpt_dir):
        for layer in self.layers:
            layer.save(i, ckpt_dir)
