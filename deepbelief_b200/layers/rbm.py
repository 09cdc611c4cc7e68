"""Binary RBM layer and the CD-k / PCD-k trainer — the B200 counterpart of deepbelief/layers/rbm.py.

Same constructor, method names, argument meaning and return shapes as the reference class
(rbm.py:29-42 and the method list in SURVEY.md §8b), but eager: methods take / return fp32 CUDA
tensors [N, units] and run hand-written CUDA through the C ABI:
  * h_probs / v_probs / sample_* / upprop / downprop / gibbs_sample -> db_layer_forward (one fused
    GEMM + bias + logistic + sampling launch per call; no materialised transposes, rbm.py:143-147);
  * train -> the tcgen05 fused chain db_cd_tc_step for binary layers whose sizes are multiples of 8,
    otherwise the general SIMT path (db_layer_forward chain + db_cd_gradients), then
    db_optimizer_step; under torch.distributed the packed gradient is sum-all-reduced over NCCL
    between the two (data-parallel CD).
Randomness: a counter-based Philox stream per layer (`seed`), or injected draws via the keyword-only
`draw=` arguments (parity tests).
"""
import os
import sys
import time
import pickle

import numpy as np
import torch

from .. import _lib as L
from .. import dist, ops, preprocessing
from . import layer
from .layer import to_device, to_numpy


from ..parallel import dist_info as _dist_info, all_reduce_sum_

import itertools
import zlib

_layer_counter = itertools.count(1)


def default_stream(name):
    """Philox key of a layer constructed without `seed=`: crc32(name) in the high word, the construction index in the low
    word. Two layers of one stack therefore draw from different streams (TF gives every random op its own stream,
    rbm.py:202,230), and every rank of a data-parallel job — same script, same construction order — derives the same keys."""
    return (zlib.crc32(str(name).encode()) << 32) | (next(_layer_counter) & 0xffffffff)


# page-locked staging and device batch buffers are kept across train() calls: cudaHostAlloc of a
# 256 MB buffer costs ~100 ms, far more than a CD step
_BUFFER_POOL = {}
_COPY_STREAMS = {}


def _pooled_buffers(device, shape, dtype=torch.float32):
    key = (str(device), tuple(shape), dtype)
    bufs = _BUFFER_POOL.get(key)
    if bufs is None:
        if len(_BUFFER_POOL) >= 4:                 # a handful of batch shapes at most; do not hoard pinned memory
            _BUFFER_POOL.clear()
        bufs = _BUFFER_POOL[key] = {
            'stage': [torch.empty(shape, dtype=dtype).pin_memory() for _ in range(2)],
            'dev': [torch.empty(shape, dtype=dtype, device=device) for _ in range(2)],
            # recorded after every copy OUT of a staging buffer, by whichever feeder made it: the next writer waits on it
            'stage_free': [torch.cuda.Event() for _ in range(2)]}
    return bufs


class BatchFeeder:
    """Host batches -> device, one batch ahead of the compute stream.

    `next()` hands out the device copy of the current batch; `prefetch()` — called by the trainer right
    after it has enqueued the step's kernels — assembles the following host batch (native threaded
    gather into pinned staging, or a zero-copy view of pinned storage) and starts its host->device
    copy on a side stream, so both overlap the CD chain of the current batch. Device and staging
    buffers are double-buffered. Exactly `count` batches are drawn from the data source — the
    reference's loop consumes one per iteration (rbm.py:695) and no more.
    allow_bytes: the consumer reads uint8 batches (the tensor-core CD chain); a {0,1} data set is then staged and
    copied as bytes (Data.binary_bytes), a quarter of the host->device traffic."""

    def __init__(self, data, device, count, allow_bytes=False):
        self.data, self.device, self.left = data, device, int(count)
        self.as_bytes = bool(allow_bytes and hasattr(data, 'binary_bytes') and hasattr(data, 'next_batch_into') and
                             data.binary_bytes() is not None)
        self.dtype = torch.uint8 if self.as_bytes else torch.float32
        self.copy_stream = _COPY_STREAMS.get(str(device))          # one side stream per device, kept across train() calls
        if self.copy_stream is None:
            self.copy_stream = _COPY_STREAMS[str(device)] = torch.cuda.Stream(device=device)
        # pooled device buffers may still be read by work a previous trainer enqueued on the compute stream
        self.copy_stream.wait_stream(torch.cuda.current_stream())
        self.dev = [None, None]
        self.stage = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]     # H2D of slot s has landed
        self.released = [None, None]                               # compute no longer reads slot s
        self.slot = 0
        self.h2d_bytes = 0
        self._pending = False
        self.prefetch()

    def _staging(self, s, shape):
        """Pinned staging buffer of slot s, safe to overwrite: the last host->device copy out of it (this feeder's or a
        previous one's) has finished. Only called when a batch really has to be assembled on the host — a full-stream
        synchronise here (as the first version did) stalled the launch queue for one step at the start of every train()."""
        pooled = _pooled_buffers(self.device, shape, self.dtype)
        self.stage[s] = pooled['stage'][s]
        self._stage_free = pooled['stage_free']
        pooled['stage_free'][s].synchronize()
        return self.stage[s]

    def prefetch(self):
        if self._pending or self.left <= 0:
            return
        self.left -= 1
        s = self.slot
        used_stage = [False]

        def staging_array():                       # evaluated by Data only when the rows are not one contiguous block
            used_stage[0] = True
            return self._staging(s, self.data.batch_shape()).numpy()
        if self.as_bytes:
            host = self.data.next_batch_into(staging_array, as_bytes=True)
        elif hasattr(self.data, 'next_batch_into'):
            host = self.data.next_batch_into(staging_array)
        else:
            host = self.data.next_batch()
            if isinstance(host, tuple):            # (datapoints, labels): the trainers use the datapoints only
                host = host[0]
        np_dtype = np.uint8 if self.as_bytes else np.float32
        src = host if isinstance(host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(host, dtype=np_dtype))
        self._pending = True
        if src.is_cuda:
            self.dev[s] = src.to(dtype=torch.float32)
            self.ready[s].record(torch.cuda.current_stream())
            return
        if self.dev[s] is None or self.dev[s].shape != src.shape:
            self.dev[s] = _pooled_buffers(self.device, src.shape, self.dtype)['dev'][s]
        if not src.is_pinned():
            stage = self._staging(s, src.shape)
            stage.copy_(src)
            src = stage
            used_stage[0] = True
        if self.released[s] is not None:
            self.copy_stream.wait_event(self.released[s])     # the step that read dev[s] has finished
        with torch.cuda.stream(self.copy_stream):
            self.dev[s].copy_(src, non_blocking=True)
            self.ready[s].record(self.copy_stream)
            if used_stage[0]:
                self._stage_free[s].record(self.copy_stream)
        self.h2d_bytes += src.numel() * src.element_size()

    def next(self):
        if not self._pending:
            self.prefetch()
        assert self._pending, 'BatchFeeder exhausted'
        s = self.slot
        main = torch.cuda.current_stream()
        # everything enqueued so far (the previous step, which read the OTHER slot) precedes this event
        ev = torch.cuda.Event()
        ev.record(main)
        self.released[s ^ 1] = ev
        main.wait_event(self.ready[s])
        self._pending = False
        self.slot ^= 1
        return self.dev[s]


class CdGraphStep:
    """One whole training iteration — lower layers re-sampled, CD-k chain (SIMT tiles, or the tcgen05 chain when the layer
    qualifies), gradient, optimizer — captured ONCE into a CUDA graph and replayed per batch. The reference-sized configurations
    (328 x 1024 x 500, CD-1) are launch-bound: ~10 kernels of a few microseconds each behind a Python call chain.
    Draw ids and the Adam step number cannot be baked into a graph, so the captured kernels add a device clock
    (db_rng_t.clock / db_opt_t.clock) that the last node of the graph advances: replay j uses exactly the draw ids
    and step number the j-th eager iteration would, and the results are bit-identical to the eager loop."""

    def __init__(self, lay, batch_shape, k, optimizer, opt_state, grad, n_global, lr, persistent):
        self.lay, self.k, self.optimizer = lay, k, optimizer
        dev = lay.device
        self.batch = torch.empty(tuple(batch_shape), dtype=torch.float32, device=dev)
        self.rngs = [l.rng for l in lay.layers]                      # this layer first, then the stack below it
        self.clocks = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in self.rngs]

        def iteration():
            v0 = lay.bottom.upprop(self.batch) if lay.bottom else self.batch
            lay.cd_update(v0, k, optimizer, opt_state, grad, n_global=n_global, lr=lr, persistent=persistent)

        # one eager iteration on a snapshot: loads every kernel, counts the draws each layer consumes, changes nothing
        snap = (lay._trainable_pair(lay._params)[0].clone(), lay._params.clone(), opt_state.clone(),
                None if persistent is None else persistent.clone(), [r.draw for r in self.rngs], optimizer.step_count,
                lay._weights_version)
        self.batch.zero_()
        iteration()
        self.draws = [r.draw - d0 for r, d0 in zip(self.rngs, snap[4])]

        def restore():
            lay._trainable_pair(lay._params)[0].copy_(snap[0])
            lay._params.copy_(snap[1])
            opt_state.copy_(snap[2])
            if persistent is not None:
                persistent.copy_(snap[3])
            for r, d0 in zip(self.rngs, snap[4]):
                r.draw = d0
            optimizer.step_count = snap[5]
            if lay._tc_plan is not None:                             # the split-weight cache follows the restored weights, and
                lay._tc_plan.refresh_weights(lay._params)            # the refresh stays OUT of the captured iteration
                lay._tc_plan.version = lay._weights_version
        restore()
        self.graph = torch.cuda.CUDAGraph()
        # capture_begin / capture_end on a side stream rather than the torch.cuda.graph context manager: that one runs
        # gc.collect() and empties the allocator cache first, 20-230 ms here against 0.1 ms per replayed iteration
        main = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(main)
        try:
            for r, c in zip(self.rngs, self.clocks):
                r.clock = c
            optimizer.clock = self.clocks[0]
            with torch.cuda.stream(side):
                self.graph.capture_begin()
                try:
                    iteration()
                    for c, n in zip(self.clocks, self.draws):
                        ops.clock_advance(c, n, 1)
                finally:
                    self.graph.capture_end()
        finally:
            for r in self.rngs:
                r.clock = None
            optimizer.clock = None
        main.wait_stream(side)
        restore()                                                    # capture executed nothing; host counters rewound
        self.base_draw, self.base_step = list(snap[4]), snap[5]       # what clock == 0 stands for
        self.synced = (list(snap[4]), snap[5])                       # host counters the device clocks currently match

    def run(self, batch):
        now = ([r.draw for r in self.rngs], self.optimizer.step_count)
        if now != self.synced:                                       # eager sampling in between (evaluation steps): re-base
            for c, d, b in zip(self.clocks, now[0], self.base_draw):
                c.copy_(torch.tensor([d - b, now[1] - self.base_step], dtype=torch.int64))
        self.batch.copy_(batch, non_blocking=True)
        self.graph.replay()
        for r, n in zip(self.rngs, self.draws):
            r.draw += n
        self.optimizer.step_count += 1
        self.lay._weights_version += 1
        if self.lay._tc_plan is not None:                            # the captured apply_update keeps the cache current
            self.lay._tc_plan.version = self.lay._weights_version
        self.synced = ([r.draw for r in self.rngs], self.optimizer.step_count)


class RBM(layer.Layer):
    """Binary-binary RBM. Args as the reference: num_v, num_h, W_std=0.1, lr=0.1, temperature=None
    (Concrete units when set), bottom=None, plotters (accepted, may be None), session (ignored), name.
    Extra keyword `seed` selects the Philox stream (default: a stream of its own per layer, see default_stream)."""

    layer_kind = L.LAYER_RBM
    visible_act = L.ACT_SIGMOID
    hidden_act = L.ACT_SIGMOID

    def __init__(self, num_v, num_h, W_std=0.1, lr=0.1, temperature=None, bottom=None, loss_plotter=None,
                 distr_plotter=None, filter_plotter=None, latent_sample_plotter=None, latent_space_explorer=None,
                 session=None, name='RBM', seed=None, device='cuda'):
        assert isinstance(num_v, int) and num_v > 1
        assert isinstance(num_h, int) and num_h > 1
        self.num_v, self.num_h = num_v, num_h
        self.temperature = temperature
        self.loss_plotter, self.distr_plotter, self.filter_plotter = loss_plotter, distr_plotter, filter_plotter
        self.ls_plotter = latent_sample_plotter
        self.latent_space_explorer = latent_space_explorer
        self.device = device
        self.rng = ops.Rng(seed=default_stream(name) if seed is None else seed)

        # one packed buffer [W | b | c | (opt_sigma_h)] so the optimizer / all-reduce see a single array
        n = ops.cd_param_count(self.layer_kind, num_v, num_h)
        self._params = torch.zeros(n, dtype=torch.float32, device=device)
        vh = num_v * num_h
        self.W = self._params[:vh].view(num_v, num_h)
        self.b = self._params[vh:vh + num_v].view(num_v, 1)
        self.c = self._params[vh + num_v:vh + num_v + num_h].view(1, num_h)
        # operands the kernels read; a layer with tied weights (SBM_Lower) exposes its shared parameters as W / b / c
        # and keeps the dense expansion here
        self.Wk, self.bk, self.ck = self.W, self.b, self.c
        # truncated normal, +-2 sigma redraw (rbm.py:59-62) — host RNG, parity runs inject W
        w0 = np.random.normal(size=(num_v, num_h))
        bad = np.abs(w0) > 2.0
        while bad.any():
            w0[bad] = np.random.normal(size=int(bad.sum()))
            bad = np.abs(w0) > 2.0
        self.W.copy_(torch.as_tensor((w0 * W_std).astype(np.float32)))
        self._init_extra_params()

        self.lr = float(lr)                  # rbm.py:74-78; only changed by lr_drop
        self.checkpoint_iter = 1             # rbm.py:80-85
        self.iter = 0
        self.loss_x_vals = self.loss_y_vals = None
        self._tc_plan = None
        self._weights_version = 0
        params = {'W': self.W, 'b': self.b, 'c': self.c}
        super().__init__(params, bottom, session, name)

    def _init_extra_params(self):
        pass

    # ---- state ----------------------------------------------------------------------------------------------
    def state_tensors(self):
        return {'params': self._params}

    def save(self, i, ckpt_dir):
        import os
        assert os.path.isdir(ckpt_dir), 'The directory does not exist: %s' % ckpt_dir
        path = os.path.join(ckpt_dir, '%s-%d' % (self.name, i))
        layer.write_checkpoint({'params': self._params.detach().cpu(), 'ckpt_iter': self.checkpoint_iter, 'lr': self.lr}, path)
        self.logger.info('State saved')
        return path

    def restore(self, ckpt_fname):
        state = torch.load(ckpt_fname, map_location='cpu')
        self._params.copy_(state['params'].to(self._params.device))
        self.checkpoint_iter = int(state.get('ckpt_iter', 1))
        self.lr = float(state.get('lr', self.lr))
        self._weights_version += 1

    def set_weights(self, W=None, b=None, c=None):
        """Overwrite parameters (parity runs inject them: TF's initial W cannot be seeded)."""
        if W is not None:
            self.W.copy_(to_device(W, self.device).reshape(self.num_v, self.num_h))
        if b is not None:
            self.b.copy_(to_device(b, self.device).reshape(self.num_v, 1))
        if c is not None:
            self.c.copy_(to_device(c, self.device).reshape(1, self.num_h))
        self._weights_version += 1

    # ---- hooks specialised by the Gaussian variants ------------------------------------------------------------
    def _h_scale(self):      # multiplies v@W before the bias (bgrbm.py:80)
        return None

    def _v_div(self):        # divides h before W@h^T (bgrbm.py:126)
        return None

    def _hidden_sampling(self, sample):
        """(sample_kind, noise_scale) for hidden units."""
        if self.temperature is not None:
            return L.SAMPLE_CONCRETE, None                                   # rbm.py:178-179
        return (L.SAMPLE_BERNOULLI if sample else L.SAMPLE_NONE), None

    def _visible_sampling(self, sample):
        if self.temperature is not None:
            return L.SAMPLE_CONCRETE, None                                   # rbm.py:183-184
        return (L.SAMPLE_BERNOULLI if sample else L.SAMPLE_NONE), None

    # ---- net inputs and probabilities ----------------------------------------------------------------------------
    def _up(self, v_batch, act, sample_kind=L.SAMPLE_NONE, noise=None, draw=None, want_act=True):
        return ops.layer_forward(to_device(v_batch, self.device), self.Wk, self.ck, False, act, sample_kind,
                                 self.temperature, col_scale=self._h_scale(), noise_scale=noise, injected=draw,
                                 rng=self.rng, want_act=want_act)

    def _down(self, h_batch, act, sample_kind=L.SAMPLE_NONE, noise=None, draw=None, want_act=True):
        return ops.layer_forward(to_device(h_batch, self.device), self.Wk, self.bk, True, act, sample_kind,
                                 self.temperature, a_div=self._v_div(), noise_scale=noise, injected=draw,
                                 rng=self.rng, want_act=want_act)

    def h_net_input(self, v_batch):
        """v @ W + c -> [N, H] (rbm.py:119-131)."""
        return self._up(v_batch, L.ACT_LINEAR)[0]

    def v_net_input(self, h_batch):
        """(W @ h^T + b)^T -> [N, V] (rbm.py:133-147)."""
        return self._down(h_batch, L.ACT_LINEAR)[0]

    def h_probs(self, v_batch):
        """sigmoid(h_net_input) (rbm.py:149-161)."""
        return self._up(v_batch, self.hidden_act)[0]

    def v_probs(self, h_batch):
        """sigmoid(v_net_input) (rbm.py:163-175)."""
        return self._down(h_batch, self.visible_act)[0]

    # ---- sampling ------------------------------------------------------------------------------------------------------
    def sample_h(self, h_probs, sample=True, summary=True, *, draw=None):
        """Bernoulli `p > u` (rbm.py:187-213), Concrete when temperature is set (rbm.py:243-258)."""
        kind, noise = self._hidden_sampling(sample)
        if kind == L.SAMPLE_NONE:
            return to_device(h_probs, self.device)
        return ops.sample(to_device(h_probs, self.device), kind, self.temperature, noise_scale=noise, injected=draw,
                          rng=self.rng)

    def sample_v(self, v_probs, sample=True, summary=True, *, draw=None):
        kind, noise = self._visible_sampling(sample)
        if kind == L.SAMPLE_NONE:
            return to_device(v_probs, self.device)
        return ops.sample(to_device(v_probs, self.device), kind, self.temperature, noise_scale=noise, injected=draw,
                          rng=self.rng)

    def _sample_h(self, h_probs, sample=True, summary=True, *, draw=None):
        return RBM.sample_h(self, h_probs, sample, summary, draw=draw)

    def _sample_v(self, v_probs, sample=True, summary=True, *, draw=None):
        return RBM.sample_v(self, v_probs, sample, summary, draw=draw)

    # fused probability + sample (one launch): what upprop / downprop / gibbs_sample use internally
    def _up_sample(self, v_batch, sample=True, draw=None, want_probs=False):
        kind, noise = self._hidden_sampling(sample)
        act, s = self._up(v_batch, self.hidden_act, kind, noise, draw, want_act=(want_probs or kind == L.SAMPLE_NONE))
        return act, (s if s is not None else act)

    def _down_sample(self, h_batch, sample=True, draw=None, want_probs=False):
        kind, noise = self._visible_sampling(sample)
        act, s = self._down(h_batch, self.visible_act, kind, noise, draw, want_act=(want_probs or kind == L.SAMPLE_NONE))
        return act, (s if s is not None else act)

    # ---- stack traversal ---------------------------------------------------------------------------------------------
    def upprop(self, v_batch, sample=True, summary=True, *, draws=None):
        """Sample of this layer's hidden units from a batch fed to the LOWEST layer (rbm.py:277-300).
        draws: optional list of injected draws, one per layer from the bottom up."""
        draws = list(draws) if draws is not None else None
        if self.bottom:
            v_batch = self.bottom.upprop(v_batch, sample, summary, draws=draws[:-1] if draws else None)
        return self._up_sample(v_batch, sample, draws[-1] if draws else None)[1]

    def downprop(self, h_batch, sample=True, summary=True, *, draws=None):
        """Sample of the LOWEST layer's visibles from this layer's hiddens (rbm.py:302-325).
        draws: optional injected draws, this layer first."""
        draws = list(draws) if draws is not None else None
        v_batch = self._down_sample(h_batch, sample, draws[0] if draws else None)[1]
        if self.bottom:
            v_batch = self.bottom.downprop(v_batch, sample, summary, draws=draws[1:] if draws else None)
        return v_batch

    def propagate(self, v_batch, k=1, sample=True, summary=True):
        """k cycles of upprop + downprop (rbm.py:327-349)."""
        assert k > 0, 'k must be > 0'
        for _ in range(k):
            v_batch = self.downprop(self.upprop(v_batch, sample, summary), sample, summary)
        return v_batch

    def gibbs_sample(self, h_batch, k=1, summary=True, *, draws=None):
        """k Gibbs steps; returns the LAST (v_probs, v, h_probs, h). Always samples — the reference
        passes `summary` into the `sample` slot (rbm.py:368,370; SURVEY §8 a'.1).
        draws: optional list of k (draw_v, draw_h)."""
        assert k > 0, 'k must be > 0'
        for s in range(k):
            dv, dh = draws[s] if draws is not None else (None, None)
            last = s == k - 1
            v_probs, v_batch = self._down_sample(h_batch, True, dv, want_probs=last)
            h_probs, h_batch = self._up_sample(v_batch, True, dh, want_probs=last)
        return v_probs, v_batch, h_probs, h_batch

    # ---- energies ------------------------------------------------------------------------------------------------------
    def free_energy(self, v_batch):
        """F(v) = -(v b + sum_j softplus((vW + c)_j)) -> [N, 1] (rbm.py:378-396)."""
        return ops.free_energy(self.layer_kind, to_device(v_batch, self.device), self.Wk, self.bk, self.ck,
                               self._sigma_vector())

    def _sigma_vector(self):
        return None

    def cd_loss(self, v_0_batch, v_k_batch):
        """mean F(v0) - mean F(vk) as a 0-d tensor (rbm.py:398-415)."""
        f0 = self.free_energy(v_0_batch)
        fk = self.free_energy(v_k_batch)
        return (f0.sum() - fk.sum()) / f0.shape[0]

    def cd_gradient(self, v_0_batch, v_k_batch, inv_global_n=None):
        """Packed gradient of cd_loss wrt (W | b | c [| opt_sigma_h]) — what TF autodiff of rbm.py:664
        produces."""
        return ops.cd_gradients(self.layer_kind, to_device(v_0_batch, self.device), to_device(v_k_batch, self.device),
                                self._params, self.num_v, self.num_h, inv_global_n)

    # ---- generation helpers ---------------------------------------------------------------------------------------------
    def datapoint_len(self):
        if self.bottom:
            return self.bottom.layers[-1].num_v
        return self.num_v

    def downprop_latent(self, x, num_samples_to_average=100):
        """Average of down-propagated samples from hidden configurations x (rbm.py:437-458)."""
        assert type(x) == np.ndarray
        assert x.ndim == 1 or x.ndim == 2
        if x.ndim == 1:
            x = x[np.newaxis, :]
        if x.shape[0] == 1:
            x_batch = np.tile(x, (num_samples_to_average, 1))
            return np.mean(to_numpy(self.downprop(x_batch, True)), axis=0)
        acc = torch.zeros((x.shape[0], self.datapoint_len()), dtype=torch.float32, device=self.device)
        xd = to_device(x, self.device)
        for _ in range(num_samples_to_average):
            acc += self.downprop(xd, True)
        return to_numpy(acc) / num_samples_to_average

    def plot_latent_samples(self, iteration=None):
        assert self.ls_plotter is not None
        v_vecs = self.downprop_latent(self.ls_plotter.grid, self.ls_plotter.num_samples_to_average)
        fig_name = 'latent_samples' + ('' if iteration is None else '-' + str(iteration))
        self.ls_plotter.plot(v_vecs, fig_name)

    def explore_latent_space(self):
        self.latent_space_explorer.set_sample_callback(self.downprop_latent)
        self.latent_space_explorer.plot()

    def test_generalisation(self, test_data, output_table_path, num_iterations=1000, num_runs=1,
                            num_samples_to_average=100):
        """Gibbs-chain reconstruction of each test point (rbm.py:470-521): a chain of
        `num_samples_to_average` copies is propagated up and down; the distance is the cross entropy
        between the chain average and the test point."""
        table = []
        for idx in range(test_data.shape[0]):
            test_point = np.reshape(test_data[idx], (1, test_data.shape[1])).astype(np.float32)
            tp = to_device(test_point, self.device)
            datapoint = to_device(np.tile(test_point, (num_samples_to_average, 1)), self.device)
            prev_dist = 10e7
            for run in range(num_runs):
                for _ in range(num_iterations):
                    datapoint = self.downprop(self.upprop(datapoint, True), True)
                model_point = to_numpy(self.downprop(self.upprop(datapoint, True), True)).mean(axis=0)
                distance = float(dist.cross_entropy(datapoint.mean(dim=0, keepdim=True), tp))
                if distance < prev_dist:
                    if table and table[-1][0] == idx:
                        del table[-1]
                    table.append((idx, distance, test_point, model_point))
                    prev_dist = distance
                self.logger.info('Test point: {}, Run: {}, Distance: {}'.format(idx, run, distance))
        with open(output_table_path, 'wb') as f:
            pickle.dump(table, f)

    # ---- evaluation ------------------------------------------------------------------------------------------------------
    def _eval_step(self, test_data, test_mean_fname=None, test_std_fname=None):
        """One-step reconstruction MSE on a test batch, pushed down to the pixels (rbm.py:648-662, 534-578)."""
        test_batch = to_device(test_data.next_batch(), self.device)
        v0 = self.bottom.upprop(test_batch) if self.bottom else test_batch
        h0 = self._up_sample(v0, True)[1]
        v1 = self.downprop(h0)
        test_loss = float(dist.mse(v1, test_batch))
        self.logger.info('Iteration: %d, Test loss: %f', self.iter, test_loss)
        self.logger.debug(self)
        self.loss_x_vals.append(self.iter)
        self.loss_y_vals.append(test_loss)
        if self.loss_plotter:
            self.loss_plotter.set_data(self.loss_x_vals, self.loss_y_vals, 'loss_' + self.name)
        if self.distr_plotter:
            data_distr = test_batch.mean(dim=0, keepdim=True)
            model_distr = v1.mean(dim=0, keepdim=True)
            if test_mean_fname and test_std_fname:
                data_distr = preprocessing.unstandardize(data_distr, test_mean_fname, test_std_fname)
                model_distr = preprocessing.unstandardize(model_distr, test_mean_fname, test_std_fname)
            self.distr_plotter.set_data((to_numpy(data_distr), to_numpy(model_distr)),
                                        'distr_' + self.name + '_' + str(self.iter))
        if self.filter_plotter:
            self.filter_plotter.set_data(to_numpy(self.W), 'filters_' + self.name + '_' + str(self.iter))
        return test_loss

    # ---- training ----------------------------------------------------------------------------------------------------------
    # rows x V x H from which the tensor-core chain is taken. Its launches cost ~10 us each however small the layer (64-row
    # tiles, programmatic dependent launch), against 22-32 us for a cluster split-K SIMT launch at the reference's sizes, but
    # an iteration also pays ~25 us of fixed work (batch conversion, one-pass update + cache rewrite). Measured in graph
    # replay (scripts/probe_tc_crossover.py, scripts/profile_small_cd_timeline.py): 328 x 1024 x 200 CD-10 0.27 ms tcgen05 vs
    # 0.45 SIMT, CD-1 0.15 vs 0.13; 328 x 1024 x 496 CD-10 0.31 vs 0.50, CD-1 0.10 vs 0.12; from 656 rows on tcgen05 wins 2-10x.
    TC_MIN_WORK = 150_000_000          # CD-1 .. CD-3
    TC_MIN_WORK_LONG_CHAIN = 50_000_000  # k >= 4: 2k+1 chain launches dominate the iteration

    def _tc_eligible(self, n_rows, k=None):
        floor = self.TC_MIN_WORK_LONG_CHAIN if (k is not None and k >= 4) else self.TC_MIN_WORK
        floor = min(floor, self.TC_MIN_WORK)
        return (self.layer_kind == L.LAYER_RBM and self.temperature is None and n_rows % 8 == 0 and
                self.num_v >= 8 and self.num_h >= 8 and n_rows >= 8 and
                n_rows * self.num_v * self.num_h >= floor)

    def cd_step(self, v_0_batch, k=1, persistent=None, *, grad=None, n_global=None, injected=None, force_simt=False):
        """Positive phase + k Gibbs steps + packed CD gradient for one (local) batch; returns
        (grad, v_k, h_k). `persistent`: PCD chain state to start from and update (fp32 [N,H] tensor,
        general path only — the tensor-core path keeps its own fp16 state in the plan)."""
        v0 = self._batch_on_device(v_0_batch)
        n = v0.shape[0]
        inv_n = 1.0 / float(n_global or n)
        if grad is None:
            grad = torch.empty_like(self._params)
        if v0.dtype == torch.uint8 and not (self._tc_eligible(n, k) and not force_simt and persistent is None):
            v0 = v0.float()
        if self._tc_eligible(n, k) and not force_simt and persistent is None:
            plan = self._tc_plan_for(n, k, n_global or n)
            plan.step(v0, self._params, grad, injected=injected, rng=self.rng)
            return grad, None, None
        # general path: one fused launch per half Gibbs step
        off = 0

        def take(count):
            nonlocal off
            if injected is None:
                return None
            d = injected[off:off + count].view(n, -1)
            off += count
            return d
        # Bernoulli-hidden layers: the probabilities h_0 and h_k are sampled from are the hidden terms of the gradient
        reuse = self.hidden_act == L.ACT_SIGMOID
        p0, h0 = self._up_sample(v0, True, take(n * self.num_h), want_probs=reuse)
        h = h0 if persistent is None else persistent
        draws = None
        if injected is not None:
            draws = []
            for _ in range(k):
                dv = take(n * self.num_v) if self._visible_sampling(True)[0] != L.SAMPLE_NONE else None
                draws.append((dv, take(n * self.num_h)))
        _, vk, pk, hk = self.gibbs_sample(h, k, draws=draws)
        if persistent is not None:
            persistent.copy_(hk)
        ops.cd_gradients(self.layer_kind, v0, vk, self._params, self.num_v, self.num_h, inv_n, grad,
                         hidden_probs=(p0, pk) if reuse else None)
        return grad, vk, hk

    def _batch_on_device(self, batch):
        """Device batch for the chain: fp32, or the feeder's uint8 {0,1} bytes, which the tensor-core chain reads directly."""
        if isinstance(batch, torch.Tensor) and batch.is_cuda and batch.dtype == torch.uint8:
            return batch.contiguous()
        return to_device(batch, self.device)

    def launches_per_cd_step(self, k):
        """Kernels of this library launched by one cd_step on the tensor-core path: batch conversion,
        2k+1 chain GEMMs, the dW GEMM and the bias-gradient reduction (csrc/cd_chain.cu)."""
        return 1 + (2 * k + 1) + 1 + 1

    # train(): iterations of the general path whose per-layer work (rows x V x H) is below this are launch-bound and run
    # as one captured CUDA graph (CdGraphStep) once the loop is at least GRAPH_MIN_ITERATIONS long
    GRAPH_MAX_WORK = 1 << 31
    GRAPH_MIN_ITERATIONS = 64

    # single-process tensor-core steps run the optimizer inside the dW kernel (False: separate optimizer + refresh launches)
    FUSED_UPDATE = True
    # data-parallel tensor-core steps on one node: the dW epilogue stores each rank's rows of the gradient straight into that
    # rank's memory over NVLink and the owner pushes the updated rows back (False, or DEEPBELIEF_B200_PEER_PUSH=0: NCCL all-reduce
    # of the whole gradient + replicated update)
    PEER_PUSH = True

    def _peer_state(self, plan, world):
        """Peer-mapped slot buffers and parameter vectors of all ranks (built once per layer), or None when the shape or the
        environment does not qualify."""
        st = getattr(self, '_peer', None)
        if st is not None and st.get('params_ptr', self._params.data_ptr()) == self._params.data_ptr():
            return st if st else None
        self._peer = {}
        if os.environ.get('DEEPBELIEF_B200_PEER_PUSH', '1') == '0' or not plan.push_ok(world):
            return None
        from ..parallel import PeerBuffers
        slots = torch.zeros(self.num_v * self.num_h, dtype=torch.float32, device=self.device)   # [world][V / world * H]
        try:
            sb, pb = PeerBuffers(slots), PeerBuffers(self._params)
        except RuntimeError as e:                # raised on every rank together: all of them take the NCCL path
            self.logger.warning('peer-push exchange unavailable (%s); using the NCCL all-reduce path', e)
            return None
        self._peer = {'slots': slots, 'slot_ptrs': sb.ptrs, 'param_ptrs': pb.ptrs, 'buffers': (sb, pb),
                      'params_ptr': self._params.data_ptr(),
                      'flag': torch.zeros(1, dtype=torch.float32, device=self.device)}
        return self._peer

    def cd_update(self, v_0_batch, k, optimizer, opt_state, grad, n_global=None, lr=None, persistent=None):
        """One full CD-k update: chain, gradient, (data-parallel) sum over ranks, optimizer step. On the tensor-core path
        the optimizer rule and the split-weight cache rewrite run inside the dW kernel (one process, deep batch) or as one
        pass over W after the NCCL all-reduce. Results are identical to cd_step + apply_gradient.
        (A pipeline of row-chunked dW launches with the all-reduce of chunk c under chunk c+1 was measured SLOWER on
        8 B200 — 4 chunked all-reduces 0.68 ms against 0.40 ms for one 134 MB call, and the dW GEMM loses the SMs left to
        NCCL: 4.85-5.00 ms per step against 4.77 ms — and was removed.)"""
        _, world = _dist_info()
        v0 = self._batch_on_device(v_0_batch)
        n = v0.shape[0]
        if (world == 1 and persistent is None and self._tc_eligible(n, k) and self.FUSED_UPDATE and
                self._trainable_pair(grad)[0] is self._params):
            # one process, tensor-core path: the dW kernel applies the optimizer rule and rewrites the split-weight cache
            # in its epilogue (db_cd_tc_step_update); bit-identical to cd_step + apply_gradient + refresh_weights
            plan = self._tc_plan_for(n, k, n_global or n)
            if plan.update_fusion_pays() and getattr(optimizer, 'clock', None) is None:
                plan.step_update(v0, self._params, grad, optimizer.descriptor(lr), opt_state, rng=self.rng)
                self._after_update()
                plan.version = self._weights_version          # the cache already holds the updated weights
                return
        if (world > 1 and persistent is None and self._tc_eligible(n, k) and self.FUSED_UPDATE and self.PEER_PUSH and
                self._trainable_pair(grad)[0] is self._params and getattr(optimizer, 'clock', None) is None):
            plan = self._tc_plan_for(n, k, n_global or n)
            peer = self._peer_state(plan, world)
            if peer is not None:
                # reduce-scatter by push from the dW epilogue, update of this rank's rows of W, all-gather by push; the two
                # small NCCL calls carry db | dc and order the ranks (every peer store has landed before it is read)
                import torch.distributed as td
                rank = _dist_info()[0]
                in_sync = getattr(self, '_replica_version', None) == self._weights_version
                nW = self.num_v * self.num_h
                plan.step_push(v0, self._params, grad, peer['slot_ptrs'], world, rank, rng=self.rng)
                td.all_reduce(grad[nW:], op=td.ReduceOp.SUM)
                plan.apply_update_sharded(peer['param_ptrs'], peer['slots'], grad, optimizer.descriptor(lr), opt_state, world, rank)
                td.all_reduce(peer['flag'], op=td.ReduceOp.SUM)
                self._after_update()
                plan.refresh_weights(self._params)
                plan.version = self._weights_version
                if in_sync:
                    self._replica_version = self._weights_version
                return
        self.cd_step(v0, k, persistent, grad=grad, n_global=n_global)
        if (persistent is None and self._tc_eligible(n, k) and self.FUSED_UPDATE and self._tc_plan is not None and
                self._trainable_pair(grad)[0] is self._params):
            # tensor-core path: all-reduce, then optimizer rule + split-weight cache rewrite in one pass over W
            in_sync = getattr(self, '_replica_version', None) == self._weights_version
            all_reduce_sum_(grad)
            plan = self._tc_plan
            plan.apply_update(self._params, grad, optimizer.descriptor(lr), opt_state)
            self._after_update()
            plan.version = self._weights_version
            if in_sync:
                self._replica_version = self._weights_version
            return
        in_sync = getattr(self, '_replica_version', None) == self._weights_version
        self.apply_gradient(optimizer, grad, opt_state, lr)
        if in_sync:
            self._replica_version = self._weights_version

    def sync_replicas(self, training_data=None):
        """Data-parallel start-up (no counterpart in the single-process reference): every layer of the stack takes rank
        0's parameters (the constructors draw W from each process's own numpy state, rbm.py:59-62), and a training set
        that has not been sharded yet is row-sharded with a shuffle seed broadcast from rank 0 (Data.shard), so that the
        ranks' batches tile the global batch a single process would have drawn."""
        import torch.distributed as td
        rank, world = _dist_info()
        if world == 1:
            return
        for lay in self.layers:
            # a layer whose parameters only moved through data-parallel updates since the last broadcast is still identical
            # on every rank (same all-reduced gradient, same rule): repeated train() calls do not re-send 134 MB
            if getattr(lay, '_replica_version', None) == lay._weights_version:
                continue
            td.broadcast(lay._params, 0)
            lay._after_update()
            lay._replica_version = lay._weights_version
        if training_data is not None and hasattr(training_data, 'shard') and getattr(training_data, '_shard', None) is None:
            seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64, device=self._params.device)
            td.broadcast(seed, 0)
            training_data.shard(rank, world, shuffle_seed=int(seed.item()))
            self.logger.info('training data sharded over %d ranks: %d of the %d rows of each batch on this rank',
                             world, training_data.batch_shape()[0], training_data.batch_size)

    def _tc_plan_for(self, n, k, n_global):
        plan = self._tc_plan
        if plan is None or (plan.N, plan.k, plan.n_global) != (n, k, n_global):
            plan = self._tc_plan = ops.CdTcPlan(n, self.num_v, self.num_h, k, pcd=False, n_global=n_global, device=self.device)
            plan.n_global = n_global
            plan.version = -1
            plan.num_sms = int(L.lib().db_num_sms(L.handle()))
        if plan.version != self._weights_version:
            plan.refresh_weights(self._params)
            plan.version = self._weights_version
        return plan

    # ---- what the optimizer sees: the packed kernel parameters themselves, unless a subclass ties weights ----
    def _trainable_pair(self, grad):
        """(trainable parameter buffer, its gradient) for the packed kernel-layout gradient `grad`."""
        return self._params, grad

    def _after_update(self):
        self._weights_version += 1

    def apply_gradient(self, optimizer, grad, opt_state, lr=None):
        """All-reduce (data parallel) + optimizer step on the packed parameters."""
        all_reduce_sum_(grad)
        params, g = self._trainable_pair(grad)
        ops.optimizer_step(optimizer.descriptor(lr), params, g, opt_state)
        self._after_update()

    def new_optimizer_state(self, optimizer):
        optimizer.reset()                                    # fresh slots per train() (rbm.py:107-117, 670)
        mult = optimizer.state_multiple()
        n = self._trainable_pair(self._params)[0].numel()
        return torch.zeros(max(mult, 1) * n, dtype=torch.float32, device=self.device)

    def train(self, optimizer, training_data, test_data, num_gibbs_steps=1, pcd=True, lr_interval=None,
              lr_drop=None, test_mean_fname=None, test_std_fname=None, max_iterations=1000, eval_interval=0,
              ckpt_interval=0, ckpt_dir=None, *, step_callback=None):
        """CD-k / PCD-k training loop (rbm.py:595-711). One iteration = next batch -> (lower layers
        re-sampled, rbm.py:621-623) -> chain -> gradient -> optimizer. Under torch.distributed every
        rank feeds its own row shard and the packed gradient is summed over NCCL.
        Extension: step_callback(iteration, grad) is called after every update with the packed
        (all-reduced) gradient still on the device — a monitoring hook; the reference has none. (With the peer-push
        exchange of a multi-rank tensor-core step, the dW part of `grad` holds the summed gradient for THIS rank's rows of W
        only; db | dc are complete on every rank.)"""
        rank, world = _dist_info()
        _tt = [time.perf_counter()] if os.environ.get('DEEPBELIEF_B200_TRAIN_TIMING') else None
        opt_state = self.new_optimizer_state(optimizer)
        if world > 1:
            self.sync_replicas(training_data)
        if _tt is not None:
            _tt.append(time.perf_counter())
        n_local = training_data.batch_shape()[0]
        n_global = n_local * world
        for lay in self.layers:                          # G-invariant draws: the lower layers re-sampled by
            lay.rng.row_offset = rank * n_local          # bottom.upprop key their draws by the global row as well
        grad = torch.empty_like(self._params)

        persistent = None
        if pcd:
            # the first batch only initialises the persistent chain and is discarded (rbm.py:631-641)
            first = training_data.next_batch()
            first = to_device(first[0] if isinstance(first, tuple) else first, self.device)
            v_first = self.bottom.upprop(first) if self.bottom else first
            persistent = self._up_sample(v_first, True)[1].clone()

        start_iter = self.checkpoint_iter
        max_iter = max_iterations + 1
        if eval_interval:
            self.loss_x_vals, self.loss_y_vals = [], []
            self.iter = start_iter - 1                     # first evaluation before any update (rbm.py:684-686)
            self._eval_step(test_data, test_mean_fname, test_std_fname)

        lr_now = optimizer.learning_rate
        graph_step = None
        width = self.layers[-1].num_v
        use_graph = (world == 1 and max_iter - start_iter >= self.GRAPH_MIN_ITERATIONS and
                     os.environ.get('DEEPBELIEF_B200_GRAPH', '1') != '0' and
                     n_local * max(l.num_v * l.num_h for l in self.layers) <= self.GRAPH_MAX_WORK)
        # binary data of a single tensor-core layer travels to the device as bytes; the chain reads them directly
        bytes_ok = self.bottom is None and not pcd and not use_graph and self._tc_eligible(n_local, num_gibbs_steps)
        feeder = BatchFeeder(training_data, self.device, max(0, max_iter - start_iter), allow_bytes=bytes_ok)
        self.last_feeder = feeder
        if _tt is not None:
            _tt.append(time.perf_counter())
        for self.iter in range(start_iter, max_iter):
            batch = feeder.next()
            if use_graph and graph_step is None:
                try:
                    graph_step = CdGraphStep(self, (n_local, width), num_gibbs_steps, optimizer, opt_state, grad, n_global,
                                             lr_now, persistent)
                except RuntimeError as e:                       # capture not possible here: plain launches
                    self.logger.warning('CUDA-graph capture of the training iteration failed (%s); eager launches', e)
                    use_graph = False
            if graph_step is not None:
                graph_step.run(batch)
            else:
                v0 = self.bottom.upprop(batch) if self.bottom else batch
                self.cd_update(v0, num_gibbs_steps, optimizer, opt_state, grad, n_global=n_global, lr=lr_now,
                               persistent=persistent)
            if _tt is not None and len(_tt) < 40:
                _tt.append(time.perf_counter())
            feeder.prefetch()              # host gather + H2D of the next batch overlap this step's kernels
            if _tt is not None and len(_tt) < 40:
                _tt.append(time.perf_counter())
            if step_callback is not None:
                step_callback(self.iter, grad)
            if _tt is not None and len(_tt) < 40:
                _tt.append(time.perf_counter())
            if lr_interval and lr_drop and self.iter % lr_interval == 0:
                self.lr *= lr_drop                         # rbm.py:672-673, 703-704: only layer.lr changes
            if eval_interval and self.iter % eval_interval == 0:
                self._eval_step(test_data, test_mean_fname, test_std_fname)
            if ckpt_dir and ckpt_interval and self.iter % ckpt_interval == 0:
                self.checkpoint_iter += ckpt_interval      # rbm.py:676, 709-711
                self.save(self.iter, ckpt_dir)
        if _tt is not None and rank == 0:                  # host-side phase times of this call (diagnostics)
            d = [1e3 * (b - a) for a, b in zip(_tt[:-1], _tt[1:])]
            sys.stderr.write('train() host ms: state+sync %.2f, setup+feeder %.2f, per step [update, prefetch, callback]: %s\n'
                             % (d[0], d[1], ' | '.join('%.2f %.2f %.2f' % tuple(d[i:i + 3]) for i in range(2, len(d) - 2, 3))))

    # ---- fine-tuning through propagate() --------------------------------------------------------------------------------
    def _stack_bottom_first(self):
        return list(reversed(self.layers))

    def propagate_with_tape(self, v_batch, k=1, sample=True, draws=None):
        """propagate() (rbm.py:327-349) that also records, per sampling op, (direction, layer, input, probabilities,
        output, mode) for the backward pass. draws: optional injected uniforms, one per sampling op in order."""
        assert k > 0, 'k must be > 0'
        stack = self._stack_bottom_first()
        for lay in stack:
            if lay.layer_kind != L.LAYER_RBM:
                raise NotImplementedError('backprop supports stacks of binary RBM layers (as the reference examples use)')
        tape, di = [], 0
        a = to_device(v_batch, self.device)

        def one(lay, x, up):
            nonlocal di
            if lay.temperature is not None:
                kind, mode = L.SAMPLE_CONCRETE, 'concrete'              # rbm.py:177-185: whatever `sample` says
            elif sample:
                kind, mode = L.SAMPLE_BERNOULLI, 'bernoulli'
            else:
                kind, mode = L.SAMPLE_NONE, 'prob'
            d = None
            if kind != L.SAMPLE_NONE and draws is not None:
                d = to_device(draws[di], self.device)
                di += 1
            fn = lay._up if up else lay._down
            p, s = fn(x, L.ACT_SIGMOID, kind, None, d, want_act=True)
            out = p if mode == 'prob' else s
            tape.append(('up' if up else 'down', lay, x, p, out, mode))
            return out
        for _ in range(k):
            for lay in stack:
                a = one(lay, a, True)
            for lay in reversed(stack):
                a = one(lay, a, False)
        return a, tape

    def backprop_gradients(self, v_batch, k=1, sample=True, draws=None):
        """loss = dist.cross_entropy(propagate(v, k, sample), v) (rbm.py:739-748) and its gradient wrt (W | b | c) of every
        layer of the stack, bottom first — what optimizer.minimize(training_loss) differentiates (rbm.py:755-758)."""
        v = to_device(v_batch, self.device)
        out, tape = self.propagate_with_tape(v, k, sample, draws)
        loss = dist.cross_entropy(out, v)
        stack = self._stack_bottom_first()
        grads = {id(lay): torch.zeros_like(lay._params) for lay in stack}
        g = ops.cross_entropy_grad(out, v, scale=1.0 / v.shape[0])
        for direction, lay, a_in, p, o, mode in reversed(tape):
            if mode == 'bernoulli':
                break                                                   # tf.cast(tf.greater(p, u)) carries no gradient
            dz = ops.concrete_backward(g, p, o, lay.temperature) if mode == 'concrete' else ops.sigmoid_backward(g, p)
            vh = lay.num_v * lay.num_h
            gl = grads[id(lay)]
            gW = gl[:vh].view(lay.num_v, lay.num_h)
            if direction == 'up':
                ops.gemm(a_in, dz, trans_a=True, beta=1.0, out=gW)
                ops.colsum(dz, out=gl[vh + lay.num_v:vh + lay.num_v + lay.num_h], beta=1.0)
                g = ops.gemm(dz, lay.Wk, trans_b=True)
            else:
                ops.gemm(dz, a_in, trans_a=True, beta=1.0, out=gW)
                ops.colsum(dz, out=gl[vh:vh + lay.num_v], beta=1.0)
                g = ops.gemm(dz, lay.Wk)
        return loss, [grads[id(lay)] for lay in stack], out

    def backprop(self, training_data, test_data, training_sampling=True, training_num_propagation_cycles=5,
                 test_num_propagation_cycles=5, learning_rate=0.1, max_iterations=1000, ignore_ckpt_iter=False,
                 eval_interval=0, ckpt_interval=0, ckpt_dir=None):
        """Cross-entropy fine-tuning of the whole stack through propagate() with plain gradient descent
        (rbm.py:713-789): per iteration one forward/backward sweep of fused GEMM+sample launches, db_gemm and
        db_concrete_backward / db_sigmoid_backward, then db_optimizer_step per layer."""
        from .. import compat
        if not training_sampling:
            training_num_propagation_cycles = 1                        # rbm.py:735-736
        stack = self._stack_bottom_first()
        opt = compat.GradientDescentOptimizer(learning_rate=learning_rate)
        start_iter = 1 if ignore_ckpt_iter else self.checkpoint_iter
        max_iter = max_iterations + 1

        def evaluate():
            tb = test_data.next_batch()
            tb = to_device(tb[0] if isinstance(tb, tuple) else tb, self.device)
            v1 = self.propagate(tb, k=test_num_propagation_cycles, sample=True)
            test_loss = float(dist.cross_entropy(v1, tb))
            self.logger.info('Iteration: %d, Test loss: %f', self.iter, test_loss)
            self.loss_x_vals.append(self.iter)
            self.loss_y_vals.append(test_loss)
            if self.loss_plotter:
                self.loss_plotter.set_data(self.loss_x_vals, self.loss_y_vals, 'loss_' + self.name)
            if self.distr_plotter:
                self.distr_plotter.set_data((to_numpy(tb.mean(dim=0, keepdim=True)), to_numpy(v1.mean(dim=0, keepdim=True))),
                                            'distr_' + self.name + '_' + str(self.iter))
        if eval_interval:
            self.loss_x_vals, self.loss_y_vals = [], []
            self.iter = start_iter - 1
            evaluate()
        feeder = BatchFeeder(training_data, self.device, max(0, max_iter - start_iter))
        for self.iter in range(start_iter, max_iter):
            batch = feeder.next()
            _, grads, _ = self.backprop_gradients(batch, training_num_propagation_cycles, training_sampling)
            feeder.prefetch()
            desc = opt.descriptor()
            for lay, gl in zip(stack, grads):
                params, g = lay._trainable_pair(gl)
                ops.optimizer_step(desc, params, g, None)
                lay._after_update()
            if eval_interval and self.iter % eval_interval == 0:
                evaluate()
            if ckpt_dir and ckpt_interval and self.iter % ckpt_interval == 0:
                self.checkpoint_iter += ckpt_interval
                self.save_all(self.iter, ckpt_dir)
