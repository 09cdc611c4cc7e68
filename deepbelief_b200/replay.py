"""Replay of one training iteration as a CUDA graph.

The reference runs one `session.run(min_step)` per iteration (gprbm.py:414, rbm.py:697). Here an iteration of the
horses-sized stacks is ~80 kernels of 5-30 us behind a Python call chain, i.e. bound by the host. `ReplayedStep` captures
the launches of one iteration once and replays them; what changes from one iteration to the next and cannot be baked into a
graph — the Philox draw ids and the optimizer's step number — is read by the kernels from a device clock
(`db_rng_t.clock` / `db_opt_t.clock`, include/deepbelief_b200.h) that the last nodes of the graph advance, so replay j
uses exactly the draw ids and step number the j-th eager iteration would, and the results are bit-identical to the eager loop.
"""
import torch

from . import ops


class ReplayedStep:
    """fn(): one iteration made of stream-ordered launches only (no device-to-host reads), already run eagerly at least once
    (kernels loaded, lazily created buffers exist). rngs: the ops.Rng streams fn draws from. optimizer: compat.Optimizer
    whose descriptor() fn uses (its step number comes from the clock of the first stream)."""

    def __init__(self, fn, rngs, optimizer, device):
        self.rngs, self.optimizer = list(rngs), optimizer
        assert self.rngs, 'an iteration without random draws needs no clock: capture it directly'
        self.clocks = [torch.zeros(2, dtype=torch.int64, device=device) for _ in self.rngs]
        base_draw, base_step = [r.draw for r in self.rngs], optimizer.step_count
        self.graph = torch.cuda.CUDAGraph()
        # capture_begin / capture_end on a side stream: the torch.cuda.graph context manager runs gc.collect() and empties
        # the allocator cache first (20-230 ms measured)
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
                    fn()
                    self.draws = [r.draw - d for r, d in zip(self.rngs, base_draw)]
                    self.steps = optimizer.step_count - base_step
                    for c, n in zip(self.clocks, self.draws):
                        ops.clock_advance(c, n, self.steps)
                finally:
                    self.graph.capture_end()
        finally:
            for r, d in zip(self.rngs, base_draw):       # capture executed nothing: rewind the host counters
                r.clock = None
                r.draw = d
            optimizer.clock = None
            optimizer.step_count = base_step
        main.wait_stream(side)
        self.base_draw, self.base_step = base_draw, base_step          # what clock == 0 stands for
        self.synced = (list(base_draw), base_step)                    # host counters the device clocks currently match

    def run(self):
        now = ([r.draw for r in self.rngs], self.optimizer.step_count)
        if now != self.synced:                            # eager sampling in between (evaluation steps): re-base the clocks
            for c, d, b in zip(self.clocks, now[0], self.base_draw):
                c.copy_(torch.tensor([d - b, now[1] - self.base_step], dtype=torch.int64))
        self.graph.replay()
        for r, n in zip(self.rngs, self.draws):
            r.draw += n
        self.optimizer.step_count += self.steps
        self.synced = ([r.draw for r in self.rngs], self.optimizer.step_count)
