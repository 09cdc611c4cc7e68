"""Tensor-level wrappers of the C ABI: torch tensors in, torch tensors out, CUDA kernels in between.

These are the calls the layer classes (deepbelief_b200.layers) are built from. Every function
launches hand-written sm_100a kernels through `libdeepbelief_b200.so`; torch only allocates.
"""
import ctypes
import os

import torch

from . import _lib as L


def _f32(t):
    assert isinstance(t, torch.Tensor) and t.is_cuda, 'expected a CUDA tensor'
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


def _bcast_mode(t, rows, cols):
    """DB_BCAST_* mode of an optional scale tensor shaped [cols], [1, cols] or [rows, cols]."""
    if t is None:
        return L.BCAST_NONE
    n = t.numel()
    if n == cols:
        return L.BCAST_ROW
    assert n == rows * cols, 'scale operand has %d elements, expected %d or %d' % (n, cols, rows * cols)
    return L.BCAST_FULL


class Rng:
    """Counter-based Philox stream: one `draw` id per sampling call (see csrc/common.cuh)."""

    def __init__(self, seed=0, row_offset=0):
        self.seed = int(seed)
        self.draw = 0
        self.row_offset = int(row_offset)
        self.clock = None          # device clock tensor (int64 [2]) while a training iteration is captured into a CUDA graph

    def next(self, count=1):
        r = L.make_rng(self.seed, self.draw, self.row_offset, self.clock.data_ptr() if self.clock is not None else None)
        self.draw += count
        return r


def layer_forward(A, W, bias, trans_w, act, sample_kind=L.SAMPLE_NONE, temperature=None, a_div=None,
                  col_scale=None, noise_scale=None, injected=None, rng=None, want_act=True, want_sample=None):
    """sample(act(col_scale * ((A / a_div) @ op(W)) + bias)) -> (act or None, sample or None)."""
    A = _f32(A)
    W = _f32(W)
    M, K = A.shape
    N = W.shape[0] if trans_w else W.shape[1]
    assert (W.shape[1] if trans_w else W.shape[0]) == K, 'inner dimensions differ: %s vs %s' % (A.shape, W.shape)
    if want_sample is None:
        want_sample = sample_kind != L.SAMPLE_NONE
    bias = _f32(bias).reshape(-1) if bias is not None else None
    a_div = _f32(a_div) if a_div is not None else None
    col_scale = _f32(col_scale) if col_scale is not None else None
    noise_scale = _f32(noise_scale) if noise_scale is not None else None
    injected = _f32(injected) if injected is not None else None
    out_act = torch.empty((M, N), device=A.device, dtype=torch.float32) if want_act else None
    out_s = torch.empty((M, N), device=A.device, dtype=torch.float32) if want_sample else None
    r = rng.next() if (rng is not None and injected is None and want_sample) else None
    rc = L.lib().db_layer_forward(
        L.handle(), L.stream(), L.fptr(A), K, L.fptr(a_div), _bcast_mode(a_div, M, K),
        L.fptr(W), W.shape[1], 1 if trans_w else 0, L.fptr(col_scale), _bcast_mode(col_scale, M, N),
        L.fptr(bias), M, N, K, act, sample_kind, float(temperature or 0.0),
        L.fptr(noise_scale), _bcast_mode(noise_scale, M, N), L.fptr(injected),
        ctypes.byref(r) if r is not None else None, L.fptr(out_act), L.fptr(out_s))
    L.check(rc, 'db_layer_forward')
    return out_act, out_s


def sample(P, sample_kind, temperature=None, noise_scale=None, injected=None, rng=None):
    P = _f32(P)
    M, N = P.shape
    noise_scale = _f32(noise_scale) if noise_scale is not None else None
    injected = _f32(injected) if injected is not None else None
    out = torch.empty_like(P)
    r = rng.next() if (rng is not None and injected is None) else None
    rc = L.lib().db_sample(L.handle(), L.stream(), L.fptr(P), M, N, sample_kind, float(temperature or 0.0),
                           L.fptr(noise_scale), _bcast_mode(noise_scale, M, N), L.fptr(injected),
                           ctypes.byref(r) if r is not None else None, L.fptr(out))
    L.check(rc, 'db_sample')
    return out


def free_energy(layer_kind, v, W, b, c, sigma_h=None):
    v = _f32(v)
    N, V = v.shape
    H = W.shape[1]
    ws = torch.empty((N, H), device=v.device, dtype=torch.float32)
    out = torch.empty((N, 1), device=v.device, dtype=torch.float32)
    rc = L.lib().db_free_energy(L.handle(), L.stream(), layer_kind, L.fptr(v), N, V, H, L.fptr(_f32(W)),
                                L.fptr(_f32(b).reshape(-1)), L.fptr(_f32(c).reshape(-1)),
                                L.fptr(_f32(sigma_h).reshape(-1)) if sigma_h is not None else None,
                                L.fptr(ws), L.fptr(out))
    L.check(rc, 'db_free_energy')
    return out


def cd_param_count(layer_kind, V, H):
    return int(L.lib().db_cd_param_count(layer_kind, V, H))


def cd_gradients(layer_kind, v0, vk, params, V, H, inv_global_n=None, grad=None, hidden_probs=None):
    """Packed CD gradient (W|b|c[|opt_sigma_h]) from the two ends of the chain. hidden_probs: (p0, pk), the hidden
    probabilities at v0 and vk when the chain has them already (Bernoulli-hidden layers)."""
    v0 = _f32(v0)
    vk = _f32(vk)
    N = v0.shape[0]
    if grad is None:
        grad = torch.empty_like(params)
    if hidden_probs is not None:
        p0, pk = _f32(hidden_probs[0]), _f32(hidden_probs[1])
        assert p0.shape == (N, H) and pk.shape == (N, H)
        rc = L.lib().db_cd_gradients_hidden(L.handle(), L.stream(), layer_kind, L.fptr(v0), L.fptr(vk), L.fptr(p0),
                                            L.fptr(pk), N, V, H, L.fptr(params),
                                            float(inv_global_n if inv_global_n is not None else 1.0 / N), L.fptr(grad))
        L.check(rc, 'db_cd_gradients_hidden')
        return grad
    ws = torch.empty((2, N, H), device=v0.device, dtype=torch.float32)
    rc = L.lib().db_cd_gradients(L.handle(), L.stream(), layer_kind, L.fptr(v0), L.fptr(vk), N, V, H,
                                 L.fptr(params), float(inv_global_n if inv_global_n is not None else 1.0 / N),
                                 L.fptr(ws), L.fptr(grad))
    L.check(rc, 'db_cd_gradients')
    return grad


def clock_advance(clock, draws, steps):
    """clock[0] += draws, clock[1] += steps on the stream (the last node of a captured training iteration)."""
    assert clock.dtype == torch.int64 and clock.numel() == 2 and clock.is_cuda
    L.check(L.lib().db_clock_advance(L.handle(), L.stream(), clock.data_ptr(), int(draws), int(steps)), 'db_clock_advance')


def optimizer_step(opt, params, grad, state):
    rc = L.lib().db_optimizer_step(L.handle(), L.stream(), ctypes.byref(opt), L.fptr(params), L.fptr(grad),
                                   L.fptr(state), params.numel())
    L.check(rc, 'db_optimizer_step')


def softplus(x):
    x = _f32(x)
    out = torch.empty_like(x)
    L.check(L.lib().db_softplus(L.handle(), L.stream(), L.fptr(x), x.numel(), L.fptr(out)), 'db_softplus')
    return out


def mse(a, b):
    a = _f32(a)
    b = _f32(b)
    out = torch.empty((), device=a.device, dtype=torch.float32)
    L.check(L.lib().db_mse(L.handle(), L.stream(), L.fptr(a), L.fptr(b), a.shape[0], a.shape[1], L.fptr(out)), 'db_mse')
    return out


def cross_entropy(probs, targets, reduce_mean=True):
    probs = _f32(probs)
    targets = _f32(targets)
    if targets.shape != probs.shape:
        targets = targets.expand_as(probs).contiguous()
    out = torch.empty((), device=probs.device, dtype=torch.float32)
    L.check(L.lib().db_cross_entropy(L.handle(), L.stream(), L.fptr(probs), L.fptr(targets), probs.shape[0],
                                     probs.shape[1], 1 if reduce_mean else 0, L.fptr(out)), 'db_cross_entropy')
    return out


# ---- tensor-core CD chain --------------------------------------------------------------------------
class CdTcPlan:
    """Owns the workspace / split-weight cache of the fused tcgen05 CD-k step for one (N, V, H, k)."""

    def __init__(self, N, V, H, k, pcd=False, n_global=None, device='cuda'):
        self.desc = L.db_cd_tc_desc_t(N, V, H, k, 1 if pcd else 0, 1.0 / float(n_global or N), 0)
        wb = int(L.lib().db_cd_tc_workspace_bytes(ctypes.byref(self.desc)))
        if wb == 0:
            raise ValueError('tensor-core CD path needs N a multiple of 8 and V, H >= 8 (got %d, %d, %d)' % (N, V, H))
        self.workspace = torch.empty(wb, dtype=torch.uint8, device=device)
        self.cache = torch.empty(int(L.lib().db_cd_tc_weight_cache_bytes(V, H)), dtype=torch.uint8, device=device)
        self.pcd_state = torch.zeros((N, (H + 7) // 8 * 8), dtype=torch.float16, device=device) if pcd else None   # row pitch pad8(H)
        self.N, self.V, self.H, self.k = N, V, H, k

    def refresh_weights(self, params):
        rc = L.lib().db_cd_tc_refresh_weights(L.handle(), L.stream(), L.fptr(params), self.V, self.H, L.ptr(self.cache))
        L.check(rc, 'db_cd_tc_refresh_weights')

    def _batch_ptr(self, v0):
        """v0 [N,V]: fp32, or uint8 {0,1} (binary data staged as bytes by the feeder: 4x less host->device traffic)."""
        assert v0.is_cuda and v0.is_contiguous() and tuple(v0.shape) == (self.N, self.V)
        assert v0.dtype in (torch.float32, torch.uint8), 'the chain reads fp32 or uint8 batches, got %s' % v0.dtype
        self.desc.v0_u8 = 1 if v0.dtype == torch.uint8 else 0
        return L.ptr(v0)

    def step(self, v0, params, grad, injected=None, rng=None, vk_out=None, hk_out=None, p0_out=None, pk_out=None):
        r = rng.next(2 * self.k + 1) if (rng is not None and injected is None) else None
        rc = L.lib().db_cd_tc_step(
            L.handle(), L.stream(), ctypes.byref(self.desc), self._batch_ptr(v0), L.fptr(params), L.ptr(self.cache),
            L.ptr(self.pcd_state), L.fptr(injected), ctypes.byref(r) if r is not None else None,
            L.ptr(self.workspace), L.fptr(grad), L.fptr(vk_out), L.fptr(hk_out), L.fptr(p0_out), L.fptr(pk_out))
        L.check(rc, 'db_cd_tc_step')
        return grad

    def update_fusable(self):
        """True when the dW kernel of this shape can carry the optimizer update in its epilogue (db_cd_tc_step_update)."""
        return int(L.lib().db_cd_tc_update_fusable(ctypes.byref(self.desc))) >= 1

    def update_fusion_pays(self):
        """True when that fused form is also the faster one (deep batches); else step() + apply_update()."""
        return int(L.lib().db_cd_tc_update_fusable(ctypes.byref(self.desc))) == 2

    def step_update(self, v0, params, grad, opt, opt_state, injected=None, rng=None):
        """step() + optimizer step + split-weight refresh in the chain's own launches (north-star K2): the dW epilogue
        updates W / the optimizer slots / the fp16 cache, the bias kernel updates b | c. `params` and `opt_state` are
        modified in place; `grad` receives the packed gradient as step() would."""
        r = rng.next(2 * self.k + 1) if (rng is not None and injected is None) else None
        rc = L.lib().db_cd_tc_step_update(
            L.handle(), L.stream(), ctypes.byref(self.desc), self._batch_ptr(v0), L.fptr(params), L.ptr(self.cache),
            L.ptr(self.pcd_state), L.fptr(injected), ctypes.byref(r) if r is not None else None,
            L.ptr(self.workspace), L.fptr(grad), ctypes.byref(opt), L.fptr(opt_state))
        L.check(rc, 'db_cd_tc_step_update')
        return grad

    def push_ok(self, world):
        """True when the data-parallel step can reduce-scatter dW by peer stores from the dW epilogue (db_cd_tc_step_push)."""
        return bool(L.lib().db_cd_tc_push_ok(ctypes.byref(self.desc), int(world)))

    def step_push(self, v0, params, grad, peer_slot_ptrs, world, rank, injected=None, rng=None):
        """step() whose dW epilogue stores the rows of dW owned by rank r into rank r's slot buffer (peer-mapped pointers);
        grad receives db | dc only."""
        r = rng.next(2 * self.k + 1) if (rng is not None and injected is None) else None
        rc = L.lib().db_cd_tc_step_push(
            L.handle(), L.stream(), ctypes.byref(self.desc), self._batch_ptr(v0), L.fptr(params), L.ptr(self.cache),
            L.ptr(self.pcd_state), L.fptr(injected), ctypes.byref(r) if r is not None else None,
            L.ptr(self.workspace), L.fptr(grad), peer_slot_ptrs, int(world), int(rank))
        L.check(rc, 'db_cd_tc_step_push')
        return grad

    def apply_update_sharded(self, peer_param_ptrs, slots, grad, opt, opt_state, world, rank):
        """Sum this rank's slots, apply the optimizer rule to its rows of W and store the new rows into every rank's
        parameter vector; b | c from the all-reduced bias gradients in grad (db_cd_tc_apply_update_sharded)."""
        rc = L.lib().db_cd_tc_apply_update_sharded(L.handle(), L.stream(), self.V, self.H, int(world), int(rank), peer_param_ptrs,
                                                   L.fptr(slots), L.fptr(grad), ctypes.byref(opt), L.fptr(opt_state))
        L.check(rc, 'db_cd_tc_apply_update_sharded')

    def apply_update(self, params, grad, opt, opt_state):
        """Optimizer step on the packed parameters + rewrite of the split-weight cache in one pass over W — what follows the
        gradient all-reduce of a data-parallel step (db_cd_tc_apply_update); bit-identical to optimizer_step +
        refresh_weights."""
        rc = L.lib().db_cd_tc_apply_update(L.handle(), L.stream(), self.V, self.H, L.fptr(params), L.fptr(grad),
                                           ctypes.byref(opt), L.fptr(opt_state), L.ptr(self.cache))
        L.check(rc, 'db_cd_tc_apply_update')


# ---- GP ---------------------------------------------------------------------------------------------
def gp_effective_hyp(hyp_opt, flags, fixed_noise, jitter):
    out = torch.empty(3, device=hyp_opt.device, dtype=torch.float32)
    L.check(L.lib().db_gp_effective_hyp(L.handle(), L.stream(), L.fptr(hyp_opt), flags, float(fixed_noise),
                                        float(jitter), L.fptr(out)), 'db_gp_effective_hyp')
    return out


def gram_rbf(X, Y, hyp, add_noise=False):
    X = _f32(X)
    N, Q = X.shape
    if Y is not None:
        Y = _f32(Y)
        assert Y.shape[1] == Q
    M = N if Y is None else Y.shape[0]
    out = torch.empty((N, M), device=X.device, dtype=torch.float32)
    L.check(L.lib().db_gram_rbf(L.handle(), L.stream(), L.fptr(X), L.fptr(Y), N, M, Q, L.fptr(hyp),
                                1 if add_noise else 0, L.fptr(out), M), 'db_gram_rbf')
    return out


def potrf_lower(A):
    """In-place lower Cholesky of A [N,N]; returns the device info scalar (0 = ok)."""
    N = A.shape[0]
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    L.check(L.lib().db_potrf_lower(L.handle(), L.stream(), L.fptr(A), N, N, L.ptr(info)), 'db_potrf_lower')
    return info


def trsm_lower(Lmat, B, trans=False):
    """In-place B <- L^-1 B (or L^-T B)."""
    N = Lmat.shape[0]
    L.check(L.lib().db_trsm_lower(L.handle(), L.stream(), L.fptr(Lmat), N, N, L.fptr(B), B.shape[1], B.shape[1],
                                  1 if trans else 0), 'db_trsm_lower')
    return B


class GplvmPlan:
    """Workspace owner for the fused LML + gradient evaluation at a fixed (N, Q, D).

    The evaluation is ~230 launches on three streams at N = 5000 (79 Cholesky steps; the triangular inverse runs beside them on
    two auxiliary streams of the library handle and joins before the K^-1 product), ~60 at N = 328. When the same argument
    set (pointers and scalars) comes back a second time — the optimisation loop of GPLVM.optimize updates X and the
    hyper-parameters in place — the launch sequence is captured into a CUDA graph and replayed from then on
    (DEEPBELIEF_B200_GRAPH=0 keeps the plain launches)."""

    MAX_GRAPHS = 8

    def __init__(self, N, Q, D, device='cuda'):
        self.N, self.Q, self.D = N, Q, D
        nbytes = int(L.lib().db_gplvm_workspace_bytes(N, Q, D))
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.scalars = torch.zeros(8, dtype=torch.float32, device=device)
        self.dX = torch.zeros((N, Q), dtype=torch.float32, device=device)
        self.info = torch.zeros(1, dtype=torch.int32, device=device)
        self.use_graph = os.environ.get('DEEPBELIEF_B200_GRAPH', '1') != '0'
        self._seen = set()
        self._graphs = {}

    def _launch(self, X, Y, hyp_opt, flags, fixed_noise, jitter, x_prior_weight, include_const, want_grad, L_out):
        rc = L.lib().db_gplvm_lml_grad(
            L.handle(), L.stream(), L.fptr(X), L.fptr(Y), self.N, self.Q, self.D, L.fptr(hyp_opt), int(flags),
            float(fixed_noise), float(jitter), float(x_prior_weight), 1 if include_const else 0,
            1 if want_grad else 0, L.ptr(self.workspace), L.fptr(self.scalars), L.fptr(self.dX), L.fptr(L_out),
            L.ptr(self.info))
        L.check(rc, 'db_gplvm_lml_grad')

    def evaluate(self, X, Y, hyp_opt, flags, fixed_noise=0.0, jitter=1e-6, x_prior_weight=0.0, include_const=True,
                 want_grad=True, L_out=None):
        args = (X, Y, hyp_opt, flags, fixed_noise, jitter, x_prior_weight, include_const, want_grad, L_out)
        if self.use_graph and not torch.cuda.is_current_stream_capturing():
            key = (X.data_ptr(), Y.data_ptr(), hyp_opt.data_ptr(), int(flags), float(fixed_noise), float(jitter),
                   float(x_prior_weight), bool(include_const), bool(want_grad), L_out.data_ptr() if L_out is not None else 0)
            graph = self._graphs.get(key)
            if graph is not None:
                graph.replay()
                return self.scalars, self.dX, self.info
            if key in self._seen and len(self._graphs) < self.MAX_GRAPHS:
                graph = torch.cuda.CUDAGraph()
                # capture_begin / capture_end on a side stream: the torch.cuda.graph context manager synchronises, runs
                # gc.collect() and empties the allocator cache first (20-230 ms measured, replay.py); nothing is allocated
                # inside the capture (workspace and outputs belong to the plan)
                main = torch.cuda.current_stream()
                side = torch.cuda.Stream()
                side.wait_stream(main)
                try:
                    with torch.cuda.stream(side):
                        graph.capture_begin()
                        try:
                            self._launch(*args)
                        finally:
                            graph.capture_end()
                    main.wait_stream(side)
                except Exception:                                  # capture not possible here: plain launches from now on
                    self.use_graph = False
                    torch.cuda.synchronize()
                    self._launch(*args)
                    return self.scalars, self.dX, self.info
                self._graphs[key] = graph
                self._keepalive = getattr(self, '_keepalive', []) + [args]      # captured pointers stay valid
                graph.replay()
                return self.scalars, self.dX, self.info
            self._seen.add(key)
        self._launch(*args)
        return self.scalars, self.dX, self.info


def gp_predict(X, Lmat, S, y_mean, xstar, hyp, want_mean=True, want_var=True):
    X = _f32(X)
    xstar = _f32(xstar)
    N, Q = X.shape
    M = xstar.shape[0]
    D = S.shape[1] if S is not None else 0
    ws = torch.empty((N, M), device=X.device, dtype=torch.float32)
    mean = torch.empty((M, D), device=X.device, dtype=torch.float32) if want_mean else None
    var = torch.empty((M,), device=X.device, dtype=torch.float32) if want_var else None
    info = torch.zeros(1, dtype=torch.int32, device=X.device)
    rc = L.lib().db_gp_predict(L.handle(), L.stream(), L.fptr(X), L.fptr(Lmat), L.fptr(S) if want_mean else None,
                               L.fptr(y_mean.reshape(-1)) if (y_mean is not None and want_mean) else None,
                               L.fptr(xstar), N, Q, D, M, L.fptr(hyp), L.fptr(ws), L.fptr(mean), L.fptr(var),
                               L.ptr(info))
    L.check(rc, 'db_gp_predict')
    return mean, var, info


# ---- backward building blocks / GPDBN ------------------------------------------------------------------
def gemm(A, B, trans_a=False, trans_b=False, alpha=1.0, beta=0.0, out=None):
    """out = alpha * op(A) @ op(B) + beta * out (fp32, db_gemm)."""
    A = _f32(A)
    B = _f32(B)
    M, K = (A.shape[1], A.shape[0]) if trans_a else (A.shape[0], A.shape[1])
    K2, N = (B.shape[1], B.shape[0]) if trans_b else (B.shape[0], B.shape[1])
    assert K == K2, 'inner dimensions differ: %s vs %s' % (tuple(A.shape), tuple(B.shape))
    if out is None:
        assert beta == 0.0
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    assert out.is_contiguous() and tuple(out.shape) == (M, N)
    rc = L.lib().db_gemm(L.handle(), L.stream(), 1 if trans_a else 0, 1 if trans_b else 0, M, N, K, float(alpha),
                         L.fptr(A), A.shape[1], L.fptr(B), B.shape[1], float(beta), L.fptr(out), N)
    L.check(rc, 'db_gemm')
    return out


def concrete_backward(g, p, a, temperature):
    g, p, a = _f32(g), _f32(p), _f32(a)
    out = torch.empty_like(p)
    L.check(L.lib().db_concrete_backward(L.handle(), L.stream(), L.fptr(g), L.fptr(p), L.fptr(a), p.shape[0], p.shape[1],
                                         float(temperature), L.fptr(out)), 'db_concrete_backward')
    return out


def sigmoid_backward(g, p):
    g, p = _f32(g), _f32(p)
    out = torch.empty_like(p)
    L.check(L.lib().db_sigmoid_backward(L.handle(), L.stream(), L.fptr(g), L.fptr(p), p.shape[0], p.shape[1], L.fptr(out)),
            'db_sigmoid_backward')
    return out


def cross_entropy_grad(probs, targets, scale=1.0):
    probs, targets = _f32(probs), _f32(targets)
    out = torch.empty_like(probs)
    L.check(L.lib().db_cross_entropy_grad(L.handle(), L.stream(), L.fptr(probs), L.fptr(targets), probs.shape[0],
                                          probs.shape[1], float(scale), L.fptr(out)), 'db_cross_entropy_grad')
    return out


def sbm_expand(W_shared, side, side_overlap, num_h, out):
    """Dense [side*side, num_h] equivalent of the patch-shared SBM_Lower weights (sbm_lower.py:59-117)."""
    L.check(L.lib().db_sbm_expand(L.handle(), L.stream(), L.fptr(_f32(W_shared)), int(side), int(side_overlap), int(num_h),
                                  L.fptr(out)), 'db_sbm_expand')
    return out


def sbm_gather(gW_eff, side, side_overlap, num_h, out):
    """Gradient of the shared SBM_Lower matrix from the gradient of its dense equivalent."""
    L.check(L.lib().db_sbm_gather(L.handle(), L.stream(), L.fptr(_f32(gW_eff)), int(side), int(side_overlap), int(num_h),
                                  L.fptr(out)), 'db_sbm_gather')
    return out


def colsum(A, out=None, scale=1.0, beta=0.0):
    A = _f32(A)
    if out is None:
        out = torch.empty(A.shape[1], device=A.device, dtype=torch.float32)
    L.check(L.lib().db_colsum(L.handle(), L.stream(), L.fptr(A), A.shape[0], A.shape[1], float(scale), float(beta),
                              L.fptr(out)), 'db_colsum')
    return out


def randn(rows, cols, rng, device='cuda'):
    """Standard normals [rows, cols] from the layer's Philox stream (Box-Muller, csrc/common.cuh)."""
    zeros = torch.zeros((rows, cols), device=device, dtype=torch.float32)
    ones = torch.ones(cols, device=device, dtype=torch.float32)
    return sample(zeros, L.SAMPLE_GAUSSIAN, noise_scale=ones, rng=rng)


class GpdbnPlan:
    """Workspace owner of the GPDBN GP block at a fixed (N, Q, D): K, L, K^-1, B = K^-1 H2, P = K^-1 g_Hm."""

    def __init__(self, N, Q, D, device='cuda'):
        self.N, self.Q, self.D = N, Q, D
        self.workspace = torch.empty(int(L.lib().db_gpdbn_workspace_bytes(N, Q, D)), dtype=torch.uint8, device=device)
        self.pred_var = torch.zeros(N, dtype=torch.float32, device=device)
        self.scalars = torch.zeros(8, dtype=torch.float32, device=device)
        self.info = torch.zeros(2, dtype=torch.int32, device=device)
        self.dX = torch.zeros((N, Q), dtype=torch.float32, device=device)
        self.dhyp = torch.zeros(3, dtype=torch.float32, device=device)

    def factor(self, X, hyp_opt, flags, fixed_noise, jitter, L_out=None):
        rc = L.lib().db_gpdbn_factor(L.handle(), L.stream(), L.fptr(X), self.N, self.Q, self.D, L.fptr(hyp_opt), int(flags),
                                     float(fixed_noise), float(jitter), L.ptr(self.workspace), L.fptr(self.pred_var),
                                     L.fptr(L_out), L.ptr(self.info))
        L.check(rc, 'db_gpdbn_factor')
        return self.pred_var

    def apply(self, H2, Hm=None):
        if Hm is None:
            Hm = torch.empty_like(H2)
        rc = L.lib().db_gpdbn_apply(L.handle(), L.stream(), L.fptr(H2), self.N, self.Q, self.D, L.ptr(self.workspace),
                                    L.fptr(Hm), L.fptr(self.scalars))
        L.check(rc, 'db_gpdbn_apply')
        return Hm, self.scalars

    def backward_h(self, g_Hm):
        g_H2 = torch.empty_like(g_Hm)
        rc = L.lib().db_gpdbn_backward_h(L.handle(), L.stream(), L.fptr(g_Hm), self.N, self.Q, self.D,
                                         L.ptr(self.workspace), L.fptr(g_H2))
        L.check(rc, 'db_gpdbn_backward_h')
        return g_H2

    def backward_x(self, X, g_pv, hyp_opt, flags, x_prior_weight=1.0):
        rc = L.lib().db_gpdbn_backward_x(L.handle(), L.stream(), L.fptr(X), L.fptr(g_pv), self.N, self.Q, self.D,
                                         L.fptr(hyp_opt), int(flags), float(x_prior_weight), L.ptr(self.workspace),
                                         L.fptr(self.dX), L.fptr(self.dhyp))
        L.check(rc, 'db_gpdbn_backward_x')
        return self.dX, self.dhyp


def gpdbn_top_forward_up(x, pred_var, alpha_h, c, eps1):
    N, D = x.shape
    H1 = torch.empty_like(x)
    H2 = torch.empty_like(x)
    mean = torch.empty(D, device=x.device, dtype=torch.float32)
    rc = L.lib().db_gpdbn_top_forward_up(L.handle(), L.stream(), L.fptr(x), L.fptr(pred_var), L.fptr(alpha_h), L.fptr(c),
                                         L.fptr(eps1), N, D, L.fptr(H1), L.fptr(mean), L.fptr(H2))
    L.check(rc, 'db_gpdbn_top_forward_up')
    return H1, mean, H2


def gpdbn_top_forward_mid(Hm, pred_var, alpha_h, H1_mean, eps2, want_H=False):
    N, D = Hm.shape
    H = torch.empty_like(Hm) if want_H else None
    y = torch.empty_like(Hm)
    rc = L.lib().db_gpdbn_top_forward_mid(L.handle(), L.stream(), L.fptr(Hm), L.fptr(pred_var), L.fptr(alpha_h),
                                          L.fptr(H1_mean), L.fptr(eps2), N, D, L.fptr(H), L.fptr(y))
    L.check(rc, 'db_gpdbn_top_forward_mid')
    return H, y


def gpdbn_top_backward1(g_y, y, Hm, pred_var, alpha_h, eps2):
    N, D = g_y.shape
    g_Hm = torch.empty_like(g_y)
    gsig = torch.empty_like(g_y)
    ga = torch.empty(D, device=g_y.device, dtype=torch.float32)
    gmean = torch.empty(D, device=g_y.device, dtype=torch.float32)
    rc = L.lib().db_gpdbn_top_backward1(L.handle(), L.stream(), L.fptr(g_y), L.fptr(y), L.fptr(Hm), L.fptr(pred_var),
                                        L.fptr(alpha_h), L.fptr(eps2), N, D, L.fptr(g_Hm), L.fptr(gsig), L.fptr(ga),
                                        L.fptr(gmean))
    L.check(rc, 'db_gpdbn_top_backward1')
    return g_Hm, gsig, ga, gmean


def gpdbn_top_backward2(g_H2, H2, x, eps1, pred_var, alpha_h, opt_alpha_h, ga, gmean, gsig, g_c, g_opt_alpha_h):
    """Returns (g_x [N,D], g_pv [N]); writes g_c [D] and g_opt_alpha_h [D] in place (views of the packed gradient)."""
    N, D = g_H2.shape
    g_x = torch.empty_like(g_H2)
    g_pv = torch.empty(N, device=g_H2.device, dtype=torch.float32)
    scratch = torch.empty(D, device=g_H2.device, dtype=torch.float32)
    rc = L.lib().db_gpdbn_top_backward2(L.handle(), L.stream(), L.fptr(g_H2), L.fptr(H2), L.fptr(x), L.fptr(eps1),
                                        L.fptr(pred_var), L.fptr(alpha_h), L.fptr(opt_alpha_h), L.fptr(ga), L.fptr(gmean),
                                        N, D, L.fptr(g_x), L.fptr(gsig), L.fptr(g_c), L.fptr(scratch),
                                        L.fptr(g_opt_alpha_h), L.fptr(g_pv))
    L.check(rc, 'db_gpdbn_top_backward2')
    return g_x, g_pv


# ---- test-time latent search (GPLVM.test_generalisation) -----------------------------------------------------
def testgen_eval(X, Kinv, A, y_mean, hyp, x, target, log_var_scaling, want_outputs=False):
    """Loss [R] and d loss / d x [R,Q] of the batched latent search (gplvm.py:478-483) at restarts x [R,Q].
    Kinv [N,N] and A = K^-1 Y [N,D] are cached by the caller (the reference refactorises K on every step)."""
    N, Q = X.shape
    R = x.shape[0]
    D = A.shape[1]
    dev = X.device
    Kxs = gram_rbf(X, x, hyp)                                                   # [N,R]
    pm = (y_mean.reshape(1, D).expand(R, D).contiguous() if y_mean is not None
          else torch.zeros((R, D), device=dev, dtype=torch.float32))
    gemm(Kxs, A, trans_a=True, beta=1.0, out=pm)                                # K_Xx^T K^-1 Y + Y_mean
    Cv = gemm(Kinv, Kxs)                                                        # K^-1 K_Xx
    pv, info = testgen_pred_var(Kxs, Cv, hyp)
    loss = torch.empty(R, device=dev, dtype=torch.float32)
    g_pm = torch.empty((R, D), device=dev, dtype=torch.float32)
    g_pv = torch.empty(R, device=dev, dtype=torch.float32)
    pm_norm = torch.empty((R, D), device=dev, dtype=torch.float32) if want_outputs else None
    pm_clip = torch.empty((R, D), device=dev, dtype=torch.float32) if want_outputs else None
    L.check(L.lib().db_testgen_rowloss(L.handle(), L.stream(), L.fptr(pm), L.fptr(target), L.fptr(pv), R, D,
                                       float(log_var_scaling), L.fptr(loss), L.fptr(pm_norm), L.fptr(pm_clip), L.fptr(g_pm),
                                       L.fptr(g_pv)), 'db_testgen_rowloss')
    gK = gemm(A, g_pm, trans_b=True)                                            # [N,R]
    g_x = testgen_grad_x(gK, Cv, Kxs, g_pv, X, x, hyp)
    return loss, g_x, info, pm_norm, pm_clip


def testgen_pred_var(Kxs, Cv, hyp):
    N, R = Kxs.shape
    pv = torch.empty(R, device=Kxs.device, dtype=torch.float32)
    info = torch.zeros(1, dtype=torch.int32, device=Kxs.device)
    L.check(L.lib().db_testgen_pred_var(L.handle(), L.stream(), L.fptr(Kxs), L.fptr(Cv), N, R, L.fptr(hyp), L.fptr(pv),
                                        L.ptr(info)), 'db_testgen_pred_var')
    return pv, info


def testgen_grad_x(gK, Cv, Kxs, g_pv, X, x, hyp):
    N, R = Kxs.shape
    Q = X.shape[1]
    g_x = torch.empty((R, Q), device=X.device, dtype=torch.float32)
    L.check(L.lib().db_testgen_grad_x(L.handle(), L.stream(), L.fptr(gK), L.fptr(Cv), L.fptr(Kxs), L.fptr(g_pv), L.fptr(X),
                                      L.fptr(x), N, R, Q, L.fptr(hyp), L.fptr(g_x)), 'db_testgen_grad_x')
    return g_x


def testgen_meanloss(Vs, target, pv, R, S, log_var_scaling, want_model_point=False):
    Dim = Vs.shape[1]
    loss = torch.empty(R, device=Vs.device, dtype=torch.float32)
    mp = torch.empty((R, Dim), device=Vs.device, dtype=torch.float32) if want_model_point else None
    g = torch.empty_like(Vs)
    L.check(L.lib().db_testgen_meanloss(L.handle(), L.stream(), L.fptr(Vs), L.fptr(target), L.fptr(pv), R, S, Dim,
                                        float(log_var_scaling), L.fptr(loss), L.fptr(mp), L.fptr(g)), 'db_testgen_meanloss')
    return loss, mp, g


def testgen_reduce_samples(gHm_rows, gsig_rows, alpha_h, pv, R, S, log_var_scaling):
    D = gHm_rows.shape[1]
    g_Hm = torch.empty((R, D), device=pv.device, dtype=torch.float32)
    g_pv = torch.empty(R, device=pv.device, dtype=torch.float32)
    L.check(L.lib().db_testgen_reduce_samples(L.handle(), L.stream(), L.fptr(gHm_rows), L.fptr(gsig_rows), L.fptr(alpha_h),
                                              L.fptr(pv), R, S, D, float(log_var_scaling), L.fptr(g_Hm), L.fptr(g_pv)),
            'db_testgen_reduce_samples')
    return g_Hm, g_pv


def heldout_mse(Vs, T):
    """loss [R] = mean_m sum_d (T[m,d] - Vs[r,d])^2 and d loss / d Vs (GPRBM.test_held_out_data, gprbm.py:453)."""
    Vs, T = _f32(Vs), _f32(T)
    R, Dim = Vs.shape
    loss = torch.empty(R, device=Vs.device, dtype=torch.float32)
    g = torch.empty_like(Vs)
    L.check(L.lib().db_heldout_mse(L.handle(), L.stream(), L.fptr(Vs), L.fptr(T), R, T.shape[0], Dim, L.fptr(loss), L.fptr(g)),
            'db_heldout_mse')
    return loss, g


def interp_objective(Xi, x_start, x_end, pv, interp_lambda):
    """(objective [1], g_x of the segment term [P,Q], g_pv [P]) of GPRBM.interp_test (gprbm.py:651-660)."""
    Xi = _f32(Xi)
    P, Q = Xi.shape
    obj = torch.empty(1, device=Xi.device, dtype=torch.float32)
    g_x = torch.empty_like(Xi)
    g_pv = torch.empty(P, device=Xi.device, dtype=torch.float32)
    L.check(L.lib().db_interp_objective(L.handle(), L.stream(), L.fptr(Xi), L.fptr(_f32(x_start)), L.fptr(_f32(x_end)), L.fptr(pv),
                                        P, Q, float(interp_lambda), L.fptr(obj), L.fptr(g_x), L.fptr(g_pv)), 'db_interp_objective')
    return obj, g_x, g_pv


# ---- 3xTF32 tensor-core GEMM -------------------------------------------------------------------------------------
def tf32_split_lo(x):
    """x - tf32(x): the remainder operand of the 3xTF32 GEMM (same shape)."""
    x = _f32(x)
    lo = torch.empty_like(x)
    L.check(L.lib().db_tf32_split_lo(L.handle(), L.stream(), L.fptr(x), x.shape[0], x.shape[1], x.shape[1], L.fptr(lo)),
            'db_tf32_split_lo')
    return lo


def gemm_tf32x3_nt(X, Y, alpha=1.0, beta=0.0, out=None, lower_only=False, X_lo=None, Y_lo=None):
    """out[M,N] = alpha * X[M,K] @ Y[N,K]^T + beta * out on the tensor cores with fp32-level accuracy."""
    X, Y = _f32(X), _f32(Y)
    M, K = X.shape
    N = Y.shape[0]
    assert Y.shape[1] == K
    X_lo = tf32_split_lo(X) if X_lo is None else X_lo
    Y_lo = (X_lo if Y is X else tf32_split_lo(Y)) if Y_lo is None else Y_lo
    if out is None:
        assert beta == 0.0
        out = torch.zeros((M, N), device=X.device, dtype=torch.float32) if lower_only else \
            torch.empty((M, N), device=X.device, dtype=torch.float32)
    rc = L.lib().db_gemm_tf32x3_nt(L.handle(), L.stream(), M, N, K, float(alpha), L.fptr(X), L.fptr(X_lo), K, L.fptr(Y),
                                   L.fptr(Y_lo), K, float(beta), L.fptr(out), N, 1 if lower_only else 0)
    L.check(rc, 'db_gemm_tf32x3_nt')
    return out


def gemm_f16x3_nt(X, Y, alpha=1.0, beta=0.0, out=None, lower_only=False):
    """out[M,N] = alpha * X[M,K] @ Y[N,K]^T + beta * out on the tensor cores from scaled fp16 hi/lo pairs (three kind::f16
    passes, fp32-level accuracy at twice the rate of the 3xTF32 form)."""
    X, Y = _f32(X), _f32(Y)
    M, K = X.shape
    N = Y.shape[0]
    assert Y.shape[1] == K
    if out is None:
        assert beta == 0.0
        out = torch.zeros((M, N), device=X.device, dtype=torch.float32) if lower_only else \
            torch.empty((M, N), device=X.device, dtype=torch.float32)
    scratch = torch.empty(int(L.lib().db_gemm_f16x3_scratch_bytes(M, N, K)), dtype=torch.uint8, device=X.device)
    rc = L.lib().db_gemm_f16x3_nt(L.handle(), L.stream(), M, N, K, float(alpha), L.fptr(X), K, L.fptr(Y), K, float(beta),
                                  L.fptr(out), N, 1 if lower_only else 0, L.ptr(scratch))
    L.check(rc, 'db_gemm_f16x3_nt')
    return out
