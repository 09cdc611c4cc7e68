"""GPRBM — the GPDBN top layer (BGRBM + GP-LVM prior over its hidden units): counterpart of
deepbelief/layers/gprbm.py.

Same constructor and method names as the reference (gprbm.py:12-39 and SURVEY.md §8b). The reference builds one TF
graph `loss = CE(downprop(sample_mean), V) + gp_loss` (gprbm.py:247-256) and lets TF autodiff + an optimizer
minimise it over EVERY trainable of the stack (gprbm.py:386-396). Here one optimisation step is an explicit
forward / backward sweep over hand-written CUDA kernels (C ABI, include/deepbelief_b200.h):
  * GP block  : db_gpdbn_factor (K, chol, K^-1, pred_var), db_gpdbn_apply (Hm, gp_loss),
                db_gpdbn_backward_h / _x (dH2, dX, d hyper-parameters) — training-mode shortcuts of SURVEY.md §8;
  * top layer : db_gpdbn_top_forward_up / _mid, db_gpdbn_top_backward1 / 2 (sigma_h = alpha_h sqrt(pred_var));
  * RBM stack : db_layer_forward (GEMM + sigmoid + Concrete sample, one launch), db_concrete_backward, db_gemm,
                db_colsum, db_cross_entropy / _grad;
  * update    : db_optimizer_step per parameter buffer (same Adam step t for all, as one TF minimize op).
Randomness: per-layer Philox streams, or injected draws (`draws=`) for the parity tests.
Eval-mode graphs: `downprop_latent`, `predict_variance` and `test_generalisation` (batched latent search with cached
K^-1) are built; `test_held_out_data` and `interp_test` are not (scope row f2 of SURVEY.md §8).
"""
import os

import numpy as np
import torch

from .. import _lib as L
from .. import compat, ops, preprocessing
from . import bgrbm, gplvm
from ..replay import ReplayedStep
from . import layer
from .layer import to_device, to_numpy

_FLAG_NOISE = 4


class GPRBM(bgrbm.BGRBM):
    def __init__(self, num_v, num_h, Q, N, eval_flag, X0=None, V=None, Y_labels=None, kern=None, noise_variance=1.0,
                 fixed_noise_variance=False, noise_jitter=1e-6, W_std=0.001, lr=0.001, temperature=None, bottom=None,
                 x_test_var=None, loss_plotter=None, distr_plotter=None, filter_plotter=None, latent_point_plotter=None,
                 latent_sample_plotter=None, latent_space_explorer=None, session=None, name='GPRBM', seed=None,
                 device='cuda'):
        self.Q, self.N, self.D = Q, N, num_h
        self.Y_labels = Y_labels
        self.x_test = x_test_var
        self.lp_plotter = latent_point_plotter
        self.eval_interval = None
        self.eval_flag = bool(eval_flag)
        self.V_data = None
        if not self.eval_flag:
            assert V is not None, 'V data must be provided at training time.'
            assert V.shape[0] == self.N
            self.V_data = to_device(V, device)
        if X0 is not None:
            assert X0.shape == (self.N, self.Q)
            X0 = np.asarray(X0, dtype=np.float32)
        elif not self.eval_flag:
            X0 = preprocessing.standardize(preprocessing.PCA(np.asarray(V, dtype=np.float32), self.Q))   # gprbm.py:57-58
        else:
            X0 = np.zeros((self.N, self.Q), dtype=np.float32)
        self._x_params = to_device(X0, device).reshape(-1).clone()
        self.X = self._x_params.view(self.N, self.Q)
        self._hyp = torch.zeros(3, dtype=torch.float32, device=device)
        self.fixed_noise = float(noise_variance)
        if fixed_noise_variance is True:
            self.noise_jitter, self._noise_flag = None, 0                      # gprbm.py:64-66: constant, no jitter
        else:
            self.noise_jitter, self._noise_flag = noise_jitter, _FLAG_NOISE
            self._hyp[2] = float(preprocessing.inverse_softplus(np.float64(noise_variance)))   # gprbm.py:68-73
        self.kern = kern or gplvm.SEKernel(session=session, Q=self.Q, device=device)
        if getattr(self.kern, 'ard', False):
            raise NotImplementedError('GPRBM takes a scalar length-scale (the GPDBN examples fix gamma, gpdbn_horses/flags.py:35-36); '
                                      'ARD kernels are supported by GPLVM')
        self.kern.bind(self._hyp)
        self._plan = ops.GpdbnPlan(self.N, self.Q, self.D, device=device)
        self._Lfac = None
        self._version = 0
        self._gp_version = -1

        super().__init__(num_v=num_v, num_h=num_h, W_std=W_std, lr=lr, temperature=temperature, bottom=bottom,
                         loss_plotter=loss_plotter, distr_plotter=distr_plotter, filter_plotter=filter_plotter,
                         latent_sample_plotter=latent_sample_plotter, latent_space_explorer=latent_space_explorer,
                         session=session, name=name, seed=seed, device=device)
        # the BGRBM sigma slot of the packed parameters holds opt_alpha_h (gprbm.py:79-83; sigma_h is derived)
        self.opt_alpha_h = self.opt_sigma_h
        del self.params['sigma_h']                                               # gprbm.py:132-135
        self.params['alpha_h'] = lambda: self.alpha_h
        self.params['X'] = self.X
        self.params['noise_variance'] = lambda: self.noise_variance
        # non-trainable snapshots the eval-mode graphs are rebuilt from (gprbm.py:85-91, 149-153)
        self.H1_var = torch.zeros((self.N, self.D), dtype=torch.float32, device=device)
        self.H2_var = torch.zeros((self.N, self.D), dtype=torch.float32, device=device)
        self.H1_mean_var = torch.zeros(self.D, dtype=torch.float32, device=device)
        self.Y_labels_var = torch.zeros((self.N, 1), dtype=torch.float32, device=device)
        self.loss = None
        self._grads = None
        self._opt_states = None

    # ---- parameters -------------------------------------------------------------------------------------------
    @property
    def flags(self):
        return self.kern.flags | self._noise_flag

    @property
    def alpha_h(self):
        return ops.softplus(self.opt_alpha_h)

    @property
    def noise_variance(self):
        if self._noise_flag:
            return ops.softplus(self._hyp[2:3])[0]
        return torch.tensor(self.fixed_noise, device=self.device)

    @property
    def sigma_h(self):
        """alpha_h [1,D] * sqrt(pred_var(X)) [N,1] (gprbm.py:114-115)."""
        pv = self._factor()
        return (self.alpha_h.reshape(1, self.D) * torch.sqrt(pv).reshape(self.N, 1)).contiguous()

    INFO_INTERVAL = 32          # optimize(): iterations between reads of the factorisation status
    GRAPH_MIN_ITERATIONS = 64   # optimize(): shortest run of iterations worth a graph capture (~3 ms)

    def _touch(self):
        self._version += 1
        self._weights_version += 1

    def set_latents(self, X):
        self.X.copy_(to_device(X, self.device))
        self._touch()

    def state_tensors(self):
        return {'params': self._params, 'X': self._x_params, 'hyp': self._hyp, 'H1': self.H1_var, 'H2': self.H2_var,
                'H1_mean': self.H1_mean_var, 'Y_labels': self.Y_labels_var}

    def save(self, i, ckpt_dir):
        assert os.path.isdir(ckpt_dir), 'The directory does not exist: %s' % ckpt_dir
        path = os.path.join(ckpt_dir, '%s-%d' % (self.name, i))
        layer.write_checkpoint({k: v.detach().cpu() for k, v in self.state_tensors().items()}, path)
        self.logger.info('State saved')
        return path

    def restore(self, ckpt_fname):
        state = torch.load(ckpt_fname, map_location='cpu')
        mine = self.state_tensors()
        for k, v in state.items():
            if k in mine:
                mine[k].copy_(v.to(mine[k].device).reshape(mine[k].shape))
        self._touch()

    def save_all(self, i, ckpt_dir):
        """gprbm.py:359-361: snapshot H1 first so eval-mode graphs can be rebuilt from the checkpoint."""
        if not self.eval_flag:
            self._snapshot_hidden()
        super().save_all(i, ckpt_dir)

    # ---- GP block -----------------------------------------------------------------------------------------------
    def _factor(self, need_L=False):
        """K, L, K^-1, pred_var at the current (X, hyper-parameters); cached until they change."""
        if self._gp_version != self._version or (need_L and self._Lfac is None):
            if need_L and self._Lfac is None:
                self._Lfac = torch.zeros((self.N, self.N), dtype=torch.float32, device=self.device)
            self._plan.factor(self.X, self._hyp, self.flags, self.fixed_noise, self.noise_jitter or 0.0, L_out=self._Lfac)
            self._gp_version = self._version
            if not getattr(self, '_defer_info', False):
                self._check_info()
        return self._plan.pred_var

    def _check_info(self):
        info = self._plan.info.tolist()
        if info[0]:
            raise RuntimeError('Cholesky decomposition was not successful: pivot %d is not positive' % info[0])
        if info[1]:
            raise RuntimeError('assert_non_negative(pred_var) failed at training point %d' % (info[1] - 1))

    def build_K(self):
        return ops.gram_rbf(self.X, None, ops.gp_effective_hyp(self._hyp, self.flags, self.fixed_noise,
                                                               self.noise_jitter or 0.0), add_noise=True)

    def build_X_prior(self):
        return (self.X * self.X).sum()

    def build_pred_var(self, x=None):
        """gprbm.py:208-218 — at the training latents (x None) or at test latents x [M,Q]."""
        if x is None:
            return self._factor().clone()
        self._factor(need_L=True)
        hyp = ops.gp_effective_hyp(self._hyp, self.flags, self.fixed_noise, self.noise_jitter or 0.0)
        _, var, info = ops.gp_predict(self.X, self._Lfac, None, None, to_device(np.array(to_numpy(x), ndmin=2), self.device),
                                      hyp, want_mean=False)
        bad = int(info.item())
        if bad:
            raise RuntimeError('assert_non_negative(pred_var) failed at test point %d' % (bad - 1))
        return var

    def predict_variance(self, x):
        return to_numpy(self.build_pred_var(np.array(x, ndmin=2)))[:, np.newaxis]

    # ---- model ---------------------------------------------------------------------------------------------------
    def build_model(self):
        """gprbm.py:162-191. Training mode binds `loss`; eval mode refreshes H2 / H1_mean from the restored H1."""
        if not self.eval_flag:
            self.loss = compat.Lazy(lambda: self.loss_and_grads(want_grads=False)[0])
            if self.Y_labels is not None:
                self.Y_labels_var.copy_(to_device(np.asarray(self.Y_labels, dtype=np.float32).reshape(self.N, 1), self.device))
        else:
            labels = to_numpy(self.Y_labels_var)
            self.Y_labels = labels if labels.any() else None
            self._refresh_eval_snapshots(self.H1_var)
        self.pred_var = compat.Lazy(self.build_pred_var)

    def _refresh_eval_snapshots(self, H1):
        mean = H1.mean(dim=0)
        self.H1_mean_var.copy_(mean)
        self.H2_var.copy_((H1 - mean) / self.alpha_h.reshape(1, self.D))

    def _snapshot_hidden(self):
        """update_H1_var (gprbm.py:142) + update_H2_var / update_H1_mean_var (gprbm.py:152-153): one fresh upprop."""
        H1 = self.upprop(self.V_data)
        self.H1_var.copy_(H1)
        self._refresh_eval_snapshots(H1)

    def _lower(self):
        return list(reversed(self.layers[1:]))           # bottom-most first

    def _sampling_kind(self, lay):
        return L.SAMPLE_CONCRETE if lay.temperature is not None else L.SAMPLE_BERNOULLI

    def loss_and_grads(self, draws=None, want_grads=True):
        """One evaluation of gprbm.py:247-256 and (optionally) its gradient wrt every trainable of the stack.
        Returns (loss 0-d tensor, pieces dict). Gradients land in self._grads:
          {'lower': [packed (W|b|c) per lower layer, bottom first], 'top': packed (W|b|c|opt_alpha_h), 'X', 'hyp'}.
        draws (optional, injected randomness): {'up': [U per lower layer], 'eps1', 'eps2', 'down_top',
        'down': [U per lower layer, top-most first]}."""
        assert not self.eval_flag, 'the loss is defined in training mode only (gprbm.py:163-165)'
        dev, N, D = self.device, self.N, self.D
        lower = self._lower()
        dr = draws or {}
        inj = lambda v: None if v is None else to_device(v, dev)  # noqa: E731
        pv = self._factor()
        alpha_h = self.alpha_h.reshape(-1)
        c_top = self.c.reshape(-1)
        # ---- H1 = upprop(V_data) ----
        a = self.V_data
        ups = []
        for i, lay in enumerate(lower):
            p, s = lay._up(a, lay.hidden_act, self._sampling_kind(lay), None, inj(dr['up'][i]) if 'up' in dr else None)
            ups.append((a, p, s))
            a = s
        x = ops.gemm(a, self.W)
        eps1 = inj(dr.get('eps1'))
        if eps1 is None:
            eps1 = ops.randn(N, D, self.rng, dev)
        H1, H1_mean, H2 = ops.gpdbn_top_forward_up(x, pv, alpha_h, c_top, eps1)
        Hm, scal = self._plan.apply(H2)
        eps2 = inj(dr.get('eps2'))
        if eps2 is None:
            eps2 = ops.randn(N, D, self.rng, dev)
        _, y = ops.gpdbn_top_forward_mid(Hm, pv, alpha_h, H1_mean, eps2)
        # ---- V_sample = downprop(H) ----
        kind_t = self._sampling_kind(self)
        q, d = ops.layer_forward(y, self.W, self.b, True, L.ACT_SIGMOID, kind_t, self.temperature,
                                 injected=inj(dr.get('down_top')), rng=self.rng)
        downs = [(y, q, d)]
        for i, lay in enumerate(reversed(lower)):
            q2, d2 = lay._down(d, lay.visible_act, self._sampling_kind(lay), None, inj(dr['down'][i]) if 'down' in dr else None)
            downs.append((d, q2, d2))
            d = d2
        ce = ops.cross_entropy(d, self.V_data, reduce_mean=False)
        loss = ce + scal[0]
        pieces = {'ce': ce, 'gp_loss': scal[0].clone(), 'log_det': scal[1].clone(), 'pred_var': pv, 'H1': H1, 'H2': H2,
                  'Hm': Hm, 'V_sample': d}
        self._last_H1 = H1
        if not want_grads:
            return loss, pieces
        self._backward(lower, ups, downs, x, eps1, eps2, pv, alpha_h, H2, Hm, d)
        return loss, pieces

    def _ensure_grads(self, lower):
        if self._grads is None:
            self._grads = {'lower': [torch.zeros_like(lay._params) for lay in lower],
                           'top': torch.zeros_like(self._params), 'X': torch.zeros_like(self._x_params),
                           'hyp': torch.zeros(3, dtype=torch.float32, device=self.device)}
        return self._grads

    @staticmethod
    def _views(lay, g):
        vh = lay.num_v * lay.num_h
        return g[:vh].view(lay.num_v, lay.num_h), g[vh:vh + lay.num_v], g[vh + lay.num_v:vh + lay.num_v + lay.num_h]

    def _backward(self, lower, ups, downs, x, eps1, eps2, pv, alpha_h, H2, Hm, v_sample):
        N, D = self.N, self.D
        G = self._ensure_grads(lower)
        g = ops.cross_entropy_grad(v_sample, self.V_data)
        # ---- down pass, lowest layer first; the dW of the down pass is WRITTEN (beta 0), the up pass adds to it ----
        for i in range(len(lower) - 1, -1, -1):
            lay = lower[len(lower) - 1 - i]
            gW, gb, gc = self._views(lay, G['lower'][len(lower) - 1 - i])
            d_in, q, d_out = downs[i + 1]
            if lay.temperature is None:
                raise NotImplementedError('GPDBN training needs Concrete units (temperature) in the lower layers: the '
                                          'Bernoulli samples of rbm.py:187-241 carry no gradient')
            dz = ops.concrete_backward(g, q, d_out, lay.temperature)          # [N, V_l]
            ops.gemm(dz, d_in, trans_a=True, out=gW)                          # dW = dz^T d_in
            ops.colsum(dz, out=gb)
            g = ops.gemm(dz, lay.Wk)                                          # [N, H_l]
        gWt, gbt, gct = self._views(self, G['top'])
        g_oah = G['top'][self.num_v * self.num_h + self.num_v + self.num_h:]
        y, q, d_out = downs[0]
        if self.temperature is None:
            raise NotImplementedError('GPDBN training needs Concrete units (temperature) in the top layer')
        dz = ops.concrete_backward(g, q, d_out, self.temperature)             # [N, V_t]
        ops.gemm(dz, y, trans_a=True, out=gWt)
        ops.colsum(dz, out=gbt)
        g_y = ops.gemm(dz, self.W)                                            # [N, D]
        g_Hm, gsig, ga, gmean = ops.gpdbn_top_backward1(g_y, y, Hm, pv, alpha_h, eps2)
        g_H2 = self._plan.backward_h(g_Hm)
        g_x, g_pv = ops.gpdbn_top_backward2(g_H2, H2, x, eps1, pv, alpha_h, self.opt_alpha_h.reshape(-1), ga, gmean, gsig,
                                            gct, g_oah)
        dX, dhyp = self._plan.backward_x(self.X, g_pv, self._hyp, self.flags, 1.0)
        G['X'].copy_(dX.view(-1))
        G['hyp'].copy_(dhyp)
        a_top = ups[-1][2] if ups else self.V_data
        ops.gemm(a_top, g_x, trans_a=True, beta=1.0, out=gWt)                 # + a^T g_x
        g_a = ops.gemm(g_x, self.W, trans_b=True) if ups else None
        # ---- up pass ----
        for l in range(len(lower) - 1, -1, -1):
            lay = lower[l]
            gW, gb, gc = self._views(lay, G['lower'][l])
            a_in, p, a_out = ups[l]
            dz = ops.concrete_backward(g_a, p, a_out, lay.temperature)        # [N, H_l]
            ops.gemm(a_in, dz, trans_a=True, beta=1.0, out=gW)
            ops.colsum(dz, out=gc)
            if l > 0:
                g_a = ops.gemm(dz, lay.Wk, trans_b=True)
        return G

    # ---- generation ---------------------------------------------------------------------------------------------------
    def datapoint_len(self):
        return self.layers[-1].num_v

    def build_sample_mean_at(self, x, num_samples=None):
        """gprbm.py:220-232 at test latents x [M,Q], from the H2 / H1_mean snapshots (eval-mode graphs)."""
        self._factor(need_L=True)
        hyp = ops.gp_effective_hyp(self._hyp, self.flags, self.fixed_noise, self.noise_jitter or 0.0)
        S = ops.trsm_lower(self._Lfac, self.H2_var.clone())
        xs = to_device(np.array(to_numpy(x), ndmin=2), self.device)
        mean, var, info = ops.gp_predict(self.X, self._Lfac, S, None, xs, hyp)
        if int(info.item()):
            raise RuntimeError('assert_non_negative(pred_var) failed at test point %d' % (int(info.item()) - 1))
        alpha_h = self.alpha_h.reshape(1, self.D)
        sigma = alpha_h * torch.sqrt(var).reshape(-1, 1)
        H = mean * alpha_h + self.H1_mean_var.reshape(1, self.D)
        rows = H.shape[0] if num_samples is None else num_samples
        noise = ops.randn(rows, self.D, self.rng, self.device)
        return H + sigma * noise, sigma

    def downprop_latent(self, x, num_samples_to_average=100):
        """gprbm.py:265-286."""
        assert type(x) == np.ndarray
        assert x.ndim == 1 or x.ndim == 2
        if x.ndim == 1:
            x = x[np.newaxis, :]

        def down(xb):
            H, sigma = self.build_sample_mean_at(xb)
            prev, self._sigma_override = self._sigma_override, sigma.contiguous()
            try:
                return self.downprop(H)
            finally:
                self._sigma_override = prev
        if x.shape[0] == 1:
            return to_numpy(down(np.tile(x, (num_samples_to_average, 1)))).mean(axis=0)
        acc = torch.zeros((x.shape[0], self.datapoint_len()), dtype=torch.float32, device=self.device)
        for _ in range(num_samples_to_average):
            acc += down(x)
        return to_numpy(acc) / num_samples_to_average

    # BGRBM hooks: sigma_h is per datapoint here
    def _sigma_vector(self):
        return self.sigma_h if self._sigma_override is None else self._sigma_override

    def _h_scale(self):
        return self._sigma_vector()

    def _v_div(self):
        return self._sigma_vector()

    def _hidden_sampling(self, sample):
        return (L.SAMPLE_GAUSSIAN, self._sigma_vector()) if sample else (L.SAMPLE_NONE, None)

    def free_energy(self, v_batch):
        raise NotImplementedError('free_energy with a per-datapoint sigma_h is not used by the reference')

    def plot_latent_points(self, iteration=None):
        assert self.lp_plotter is not None
        var_diag = to_numpy(self.build_pred_var(self.lp_plotter.grid))
        fig_name = 'latent_points' + ('' if iteration is None else '-' + str(iteration))
        self.lp_plotter.plot(to_numpy(self.X), np.log(var_diag), fig_name, self.Y_labels)

    # ---- optimisation -------------------------------------------------------------------------------------------------
    def _log_training_info(self, loss, i):
        self.logger.info('Iteration: %d, Loss: %f', i, loss)
        self.logger.debug(self)

    def _eval_step(self, i):
        loss = float(self.loss_and_grads(want_grads=False)[0])
        self._log_training_info(loss, i)
        if self.lp_plotter:
            self.plot_latent_points(i)
        if self.ls_plotter:
            self._snapshot_hidden()                       # gprbm.py:328-337 (eval_flag flips around the plot)
            self.plot_latent_samples(i)
        return loss

    def _trainable_buffers(self, fixed_X):
        lower = self._lower()
        G = self._ensure_grads(lower)
        bufs = [lay._trainable_pair(g) + (lay,) for lay, g in zip(lower, G['lower'])]
        bufs.append((self._params, G['top'], self))
        if not fixed_X:
            bufs.append((self._x_params, G['X'], None))
        bufs.append((self._hyp, G['hyp'], None))
        return bufs

    def optimize(self, optimizer, num_iterations=100, fixed_X_num_iterations=None, eval_interval=0, ckpt_interval=0,
                 ckpt_dir=None, summary_writer=None):
        """gprbm.py:375-438: Adam (or any compat optimizer) over every trainable of the stack; the first
        `fixed_X_num_iterations` iterations leave X out of the update (its slots stay untouched, as with TF's
        second minimize op over var_list without X)."""
        optimizer.reset()                                                  # fresh slots (gprbm.py:363-373, 398)
        states = {}

        def step(fixed_X):
            self.loss_and_grads(want_grads=True)
            desc = optimizer.descriptor()                                  # ONE step count for all buffers
            for params, grad, lay in self._trainable_buffers(fixed_X):
                key = params.data_ptr()
                if key not in states:
                    states[key] = torch.zeros(max(optimizer.state_multiple(), 1) * params.numel(), dtype=torch.float32,
                                              device=self.device)
                ops.optimizer_step(desc, params, grad, states[key])
                if lay is not None:
                    lay._after_update()
            self._touch()

        i = 0
        if eval_interval:
            self.eval_interval = eval_interval
            self._eval_step(i)
        # The factorisation status (pivot / negative pred_var, what TF raises inside session.run) is a device-to-host read
        # that would stall the launch queue once per iteration: inside the loop it is read every INFO_INTERVAL iterations
        # and after the last one — a failure aborts the training a few iterations late, with the same error.
        self._defer_info = True
        try:
            fixed = fixed_X_num_iterations or 0
            assert fixed <= num_iterations
            # the horses-sized stacks are bound by the host (~80 launches of 5-30 us per iteration): once a mode (X fixed /
            # free) has run eagerly it is captured and replayed as one CUDA graph (replay.ReplayedStep), bit-identical
            use_graph = os.environ.get('DEEPBELIEF_B200_GRAPH', '1') != '0'
            rngs = [l.rng for l in self.layers]
            warmed, replayed = set(), {}
            for i in range(1, num_iterations + 1):
                fx = i <= fixed
                left = (fixed if fx else num_iterations) - i + 1          # iterations left in this mode
                if use_graph and fx in warmed and fx not in replayed and left >= self.GRAPH_MIN_ITERATIONS:
                    self._touch()                                         # the captured iteration must re-factorise
                    try:
                        replayed[fx] = ReplayedStep(lambda: step(fixed_X=fx), rngs, optimizer, self.device)
                    except RuntimeError as e:
                        self.logger.warning('CUDA-graph capture of the training iteration failed (%s); eager launches', e)
                        use_graph = False
                if fx in replayed:
                    replayed[fx].run()
                    for l in self.layers[1:]:
                        l._weights_version += 1           # (their post-update launches are part of the graph)
                    self._touch()
                else:
                    step(fixed_X=fx)
                    warmed.add(fx)
                if i % self.INFO_INTERVAL == 0:
                    self._check_info()
                if eval_interval and i % eval_interval == 0:
                    self._eval_step(i)
                if ckpt_dir and ckpt_interval and i % ckpt_interval == 0:
                    self.save_all(i, ckpt_dir)
                    self.kern.save(i, ckpt_dir)
        except KeyboardInterrupt:
            self.logger.info('Training interrupted')
            if eval_interval:
                self._eval_step(i)
            if ckpt_dir and ckpt_interval:
                self.save_all(i, ckpt_dir)
                self.kern.save(i, ckpt_dir)
        finally:
            self._defer_info = False
        self._check_info()

    # ---- test-time latent search (gprbm.py:468-562) ------------------------------------------------------------------
    def _latent_search_cache(self):
        """L^-1 and B = K^-1 H2 for the frozen model (H2 / H1_mean are the snapshots of gprbm.py:149-153).
        The predictive variance is formed as the reference forms it, alpha - |L^-1 k|^2 (gprbm.py:208-214): a sum of squares.
        k^T (K^-1 k) through an explicit K^-1 cancels ~1/noise-fold and returned negative variances in fp32 at the horses
        settings (noise 0.01, pred_var ~1e-4 next to the training points)."""
        self._factor(need_L=True)
        eye = torch.eye(self.N, device=self.device, dtype=torch.float32)
        Linv = ops.trsm_lower(self._Lfac, eye)
        B = ops.gemm(Linv, ops.gemm(Linv, self.H2_var), trans_a=True)
        return Linv, B

    def _latent_forward(self, x, num_samples, cache=None, draws=None):
        """Frozen-model forward at test latents x [R,Q], S samples each (rows r * S + s): K_Xx, predictive variance,
        sample mean (gprbm.py:208-232 with sigma_h = alpha_h sqrt(pred_var), gprbm.py:114-115) and downprop through the
        Concrete stack. Returns the tape the x-gradient needs."""
        if self.temperature is None:
            raise NotImplementedError('the latent search differentiates through Concrete units (temperature)')
        Linv, B = cache or self._latent_search_cache()
        dev, D, S = self.device, self.D, int(num_samples)
        x = to_device(x, dev)
        R = x.shape[0]
        dr = draws or {}
        inj = lambda v: None if v is None else to_device(v, dev)  # noqa: E731
        hyp = ops.gp_effective_hyp(self._hyp, self.flags, self.fixed_noise, self.noise_jitter or 0.0)
        alpha_h = self.alpha_h.reshape(-1)
        Kxs = ops.gram_rbf(self.X, x, hyp)                                   # [N,R]
        Hm = ops.gemm(Kxs, B, trans_a=True)                                  # (L^-1 K_Xx)^T (L^-1 H2)
        V1 = ops.gemm(Linv, Kxs)                                             # L^-1 K_Xx
        pv, info = ops.testgen_pred_var(V1, V1, hyp)                         # alpha - colsum(V1^2)
        Cv = ops.gemm(Linv, V1, trans_a=True)                                # K^-1 K_Xx (the x-gradient of pred_var needs it)
        rows = torch.arange(R, device=dev).repeat_interleave(S)
        Hm_rows, pv_rows = Hm[rows].contiguous(), pv[rows].contiguous()
        eps = inj(dr.get('eps'))
        if eps is None:
            eps = ops.randn(R * S, D, self.rng, dev)
        _, y = ops.gpdbn_top_forward_mid(Hm_rows, pv_rows, alpha_h, self.H1_mean_var, eps)
        q, d = ops.layer_forward(y, self.W, self.b, True, L.ACT_SIGMOID, L.SAMPLE_CONCRETE, self.temperature,
                                 injected=inj(dr.get('down_top')), rng=self.rng)
        tape = [(self, q, d)]
        for i, lay in enumerate(self.layers[1:]):                              # top-most lower layer first
            if lay.temperature is None:
                raise NotImplementedError('the latent search differentiates through Concrete units (temperature)')
            q, d = lay._down(d, lay.visible_act, L.SAMPLE_CONCRETE, None, inj(dr['down'][i]) if 'down' in dr else None)
            tape.append((lay, q, d))
        return dict(x=x, R=R, S=S, B=B, hyp=hyp, alpha_h=alpha_h, Kxs=Kxs, Cv=Cv, pv=pv, info=info, Hm_rows=Hm_rows,
                    pv_rows=pv_rows, eps=eps, y=y, tape=tape, out=d)

    def _latent_backward(self, fw, g, log_var_scaling):
        """d loss / d x [R,Q] from g = d loss / d (down-propagated rows); adds log_var_scaling / pred_var to d / d pred_var."""
        for lay, q, d_out in reversed(fw['tape']):
            g = ops.gemm(ops.concrete_backward(g, q, d_out, lay.temperature), lay.Wk)
        gHm_rows, gsig_rows, _, _ = ops.gpdbn_top_backward1(g, fw['y'], fw['Hm_rows'], fw['pv_rows'], fw['alpha_h'], fw['eps'])
        g_Hm, g_pv = ops.testgen_reduce_samples(gHm_rows, gsig_rows, fw['alpha_h'], fw['pv'], fw['R'], fw['S'], log_var_scaling)
        return ops.testgen_grad_x(ops.gemm(fw['B'], g_Hm, trans_b=True), fw['Cv'], fw['Kxs'], g_pv, self.X, fw['x'], fw['hyp'])

    def latent_search_loss(self, x, test_point, log_var_scaling=0.1, num_samples=1, cache=None, draws=None,
                           want_outputs=False):
        """loss [R] and d loss / d x [R,Q] of gprbm.py:490-499 at restarts x [R,Q], each with `num_samples` Gaussian
        samples of the top hidden layer pushed down the Concrete stack (rows ordered r * S + s).
        draws (optional injected randomness): {'eps' [R*S,D], 'down_top' [R*S,Vt], 'down': [per lower layer, top first]}."""
        fw = self._latent_forward(x, num_samples, cache, draws)
        target = to_device(test_point, self.device).reshape(-1)
        loss, model_point, g = ops.testgen_meanloss(fw['out'], target, fw['pv'], fw['R'], fw['S'], log_var_scaling, want_outputs)
        g_x = self._latent_backward(fw, g, log_var_scaling)
        return loss, g_x, fw['info'], model_point, fw['pv']

    def held_out_loss(self, x, test_data, cache=None, draws=None):
        """The objective of test_held_out_data (gprbm.py:447-453) at x [R,Q] (the reference's x_test is one row): the mean
        over the test rows of the squared distance to ONE down-propagated sample per row of x, and its x-gradient."""
        fw = self._latent_forward(x, 1, cache, draws)
        loss, g = ops.heldout_mse(fw['out'], to_device(test_data, self.device))
        return loss, self._latent_backward(fw, g, 0.0), fw['info']

    def test_generalisation(self, test_data, output_table_path, optimizer, num_iterations=1000, log_var_scaling=0.1,
                            method='all_training', num_runs=1, num_samples_to_average=1):
        """gprbm.py:468-562. The restarts of one test point advance together as a batch [R,Q] with K^-1 and K^-1 H2
        cached (the reference runs them one after the other and refactorises K on every step); optimizer slots start
        fresh per test point."""
        import pickle
        test_data = np.asarray(test_data, dtype=np.float32)
        assert num_runs <= self.N
        if self.eval_flag is False:
            self._snapshot_hidden()                       # H2 / H1_mean from a fresh upprop, as _eval_step does before plots
        cache = self._latent_search_cache()
        rs = np.random.RandomState(0)
        table = []
        for test_point_index in range(test_data.shape[0]):
            test_point = test_data[test_point_index:test_point_index + 1]
            if method == 'all_training':
                assert num_runs == self.N
                x = self.X.clone().contiguous()
            elif method == 'random_normal':
                x = to_device(rs.normal(size=(num_runs, self.Q)), self.device)
            elif method == 'random_training':
                idx = rs.randint(0, max(self.N - 1, 1), size=num_runs)        # maxval N-1 exclusive (gprbm.py:516-521)
                x = self.X[torch.as_tensor(idx, device=self.device)].clone().contiguous()
            else:
                raise AssertionError('unknown method %r' % method)
            optimizer.reset()
            state = torch.zeros(max(optimizer.state_multiple(), 1) * x.numel(), dtype=torch.float32, device=self.device)
            for _ in range(num_iterations):
                _, g_x, _, _, _ = self.latent_search_loss(x, test_point, log_var_scaling, num_samples_to_average, cache)
                ops.optimizer_step(optimizer.descriptor(), x.view(-1), g_x.view(-1), state)
            loss, _, info, model_point, pv = self.latent_search_loss(x, test_point, log_var_scaling, num_samples_to_average,
                                                                    cache, want_outputs=True)
            bad = int(info.item())
            if bad:
                raise RuntimeError('assert_non_negative(pred_var) failed for restart %d' % (bad - 1))
            dist_np = to_numpy(loss)
            best = int(np.argmin(dist_np))
            for run in range(x.shape[0]):
                self.logger.info('Test point: {}, Run: {}, Distance: {}'.format(test_point_index, run, dist_np[run]))
            table.append((test_point_index, float(dist_np[best]), to_numpy(x[best:best + 1]), test_point,
                          to_numpy(model_point[best]), float(pv[best])))
        with open(output_table_path, 'wb') as f:
            pickle.dump(table, f)
        return table

    def test_held_out_data(self, test_data, output_x_fname, optimizer, num_iterations=1000):
        """gprbm.py:440-466: minimise dist.mse(test_data, downprop(sample_mean(x_test))) over the single test latent
        x_test [1,Q] (a fresh draw of the hidden noise and of the Concrete units every iteration, as session.run would
        give) and np.save the result. x_test starts from the constructor's x_test_var when one was passed, else from a
        standard-normal draw (the initializer gpdbn_horses/test_generalisation.py:36 uses)."""
        test_data = np.asarray(test_data, dtype=np.float32)
        assert test_data.ndim == 2 and test_data.shape[1] == self.datapoint_len()
        if self.eval_flag is False:
            self._snapshot_hidden()
        cache = self._latent_search_cache()
        x0 = self.x_test
        if x0 is None:
            x = ops.randn(1, self.Q, self.rng, self.device)
        else:
            x0 = x0.value if isinstance(x0, compat.Variable) else x0
            x = to_device(np.array(to_numpy(x0), dtype=np.float32, ndmin=2), self.device).clone().contiguous()
        assert x.shape == (1, self.Q)
        T = to_device(test_data, self.device)
        optimizer.reset()
        state = torch.zeros(max(optimizer.state_multiple(), 1) * x.numel(), dtype=torch.float32, device=self.device)
        for _ in range(num_iterations):
            _, g_x, _ = self.held_out_loss(x, T, cache)
            ops.optimizer_step(optimizer.descriptor(), x.view(-1), g_x.view(-1), state)
        x_np = to_numpy(x)
        if isinstance(self.x_test, compat.Variable):
            self.x_test.assign(x_np)                      # the reference optimises the variable in place
        np.save(output_x_fname, x_np)
        return x_np

    # ---- geodesic interpolation (gprbm.py:564-689) ---------------------------------------------------------------
    def create_observed_prediction_with_variance(self, tf_input, num_samples, cache=None, draws=None):
        """gprbm.py:564-612: for every latent point of tf_input [P,Q] the mean over `num_samples` down-propagated samples
        drawn with sigma_h = alpha_h sqrt(pred_var_p), and the predictive variances [P]."""
        fw = self._latent_forward(tf_input, num_samples, cache, draws)
        bad = int(fw['info'].item())
        if bad:
            raise RuntimeError('assert_non_negative(pred_var) failed at interpolation point %d' % (bad - 1))
        images = fw['out'].view(fw['R'], fw['S'], -1).mean(dim=1)          # tf.reduce_mean over the samples of a point
        return images, fw['pv']

    def interp_objective(self, Xi, x_start, x_end, interp_lambda, cache=None):
        """Objective of interp_test (gprbm.py:651-660) at the path Xi [P,Q] and its gradient with respect to Xi."""
        Linv, _ = cache or self._latent_search_cache()
        Xi = to_device(Xi, self.device)
        hyp = ops.gp_effective_hyp(self._hyp, self.flags, self.fixed_noise, self.noise_jitter or 0.0)
        Kxs = ops.gram_rbf(self.X, Xi, hyp)
        V1 = ops.gemm(Linv, Kxs)
        pv, info = ops.testgen_pred_var(V1, V1, hyp)
        Cv = ops.gemm(Linv, V1, trans_a=True)
        obj, g_seg, g_pv = ops.interp_objective(Xi, to_device(x_start, self.device).reshape(-1),
                                                to_device(x_end, self.device).reshape(-1), pv, interp_lambda)
        g_x = ops.testgen_grad_x(torch.zeros_like(Kxs), Cv, Kxs, g_pv, self.X, Xi, hyp)
        return obj, g_x + g_seg, info

    def interp_test(self, starting_point_index, end_point_index, num_points=20, interp_lambda=0.1, learning_rate=0.01,
                    num_iterations=500, num_sample_to_average=100):
        """gprbm.py:614-689: a path of num_points + 2 latents between two training latents, first the straight line, then
        optimised with Adam(learning_rate) towards short segments through low predictive variance. Returns
        (linear_points, geodesic_points): the mean down-propagated images along both paths [num_points + 2, V]."""
        if self.eval_flag is False:
            self._snapshot_hidden()
        cache = self._latent_search_cache()
        X = to_numpy(self.X)
        x_s = np.reshape(X[starting_point_index], (1, self.Q))
        x_e = np.reshape(X[end_point_index], (1, self.Q))
        a = np.linspace(0.0, 1.0, num_points + 2)[:, np.newaxis]             # 0, interior points, 1 (gprbm.py:627-634)
        xx = (x_s + np.dot(a, x_e - x_s)).astype(np.float32)
        Xi = to_device(xx, self.device).clone().contiguous()
        obs_init, _ = self.create_observed_prediction_with_variance(Xi, num_sample_to_average, cache)
        optimizer = compat.train.AdamOptimizer(learning_rate=learning_rate)
        optimizer.reset()
        state = torch.zeros(max(optimizer.state_multiple(), 1) * Xi.numel(), dtype=torch.float32, device=self.device)
        refresh_iter = int(np.ceil(num_iterations / 10))
        cost = float('nan')
        for i in range(num_iterations):
            obj, g_x, info = self.interp_objective(Xi, x_s, x_e, interp_lambda, cache)
            ops.optimizer_step(optimizer.descriptor(), Xi.view(-1), g_x.view(-1), state)
            if (i % refresh_iter) == 0:
                bad = int(info.item())
                if bad:
                    raise RuntimeError('assert_non_negative(pred_var) failed at interpolation point %d' % (bad - 1))
                cost = float(obj.item())
                print('  opt iter {:5}: {}'.format(i, cost))
        if num_iterations:
            print('Final iter {:5}: {}'.format(num_iterations - 1, float(obj.item())))
        self.interp_latents = to_numpy(Xi)
        geodesic, _ = self.create_observed_prediction_with_variance(Xi, num_sample_to_average, cache)
        return to_numpy(obs_init), to_numpy(geodesic)
