"""RBF kernel and GP-LVM layer — the B200 counterpart of deepbelief/layers/gplvm.py.

`SEKernel` and `GPLVM` keep the reference constructors and method names (gplvm.py:11-17, 91-108 and
SURVEY.md §8b). The arithmetic is the CUDA GP chain behind the C ABI: db_gram_rbf (K3),
db_potrf_lower (K4), db_trsm_lower (K5), db_gplvm_lml_grad (fused K3-K6: loss and its analytic
gradient, replacing TF autodiff of gplvm.py:425-428) and db_gp_predict.

Eager semantics: attributes the reference builds as graph tensors (`K`, `L`, `loss`, `pred_mean`,
`pred_var`, …) are `compat.Lazy` values here — re-evaluated whenever a `Session.run` (or `.eval()`)
asks for them, exactly as a TF graph node would be.
ARD length-scales (gplvm.py:24-28, gamma [1,Q]): K depends on x only through x / gamma, so the layers hand the CUDA
chain the latents in kernel coordinates (X / gamma, scalar gamma slot fixed at 1) and chain the gradients back:
dL/dX = dL/dXs / gamma, dL/dgamma_q = -sum_n dL/dXs[n,q] X[n,q] / gamma_q^2.
"""
import pickle

import numpy as np
import torch

from .. import compat, dist, ops, preprocessing
from . import layer
from .layer import to_device, to_numpy

_FLAG_ALPHA, _FLAG_GAMMA, _FLAG_NOISE = 1, 2, 4


def _inv_softplus(v):
    return float(preprocessing.inverse_softplus(np.float64(v)))


class SEKernel(layer.Stateful):
    """K(x, y) = alpha * exp(-0.5 * |x/gamma - y/gamma|^2); alpha / gamma are softplus-constrained
    optimisation variables, or the constant 1.0 when passed as False (gplvm.py:18-49)."""

    def __init__(self, alpha=1.0, gamma=1.0, ARD=False, Q=None, session=None, name='SEKernel', device='cuda'):
        self.Q = Q
        self.device = device
        self.ard = bool(ARD) and gamma is not False
        self.flags = (0 if alpha is False else _FLAG_ALPHA) | (0 if (gamma is False or self.ard) else _FLAG_GAMMA)
        # [opt_alpha, opt_gamma, (noise slot, owned by the GP layer)]
        self._hyp = torch.zeros(3, dtype=torch.float32, device=device)
        self._hyp[0] = _inv_softplus(alpha) if alpha is not False else 0.0
        self._opt_gamma_vec = None
        if self.ard:
            assert Q is not None, 'ARD needs the latent dimensionality Q'
            if type(gamma) == np.ndarray:
                assert gamma.shape == (1, Q)                                       # gplvm.py:26-27
            g = (gamma * np.ones((1, Q))).astype(np.float64)
            self._opt_gamma_vec = torch.as_tensor(preprocessing.inverse_softplus(g).reshape(-1).astype(np.float32)).to(device)
        else:
            self._hyp[1] = _inv_softplus(gamma) if gamma is not False else 0.0
        super().__init__(params={'alpha': lambda: self.alpha, 'gamma': lambda: self.gamma}, session=session, name=name)

    def bind(self, storage, ard_storage=None):
        """Move the optimisation variables into `storage` (a 3-float view of the GP layer's packed parameters; the ARD
        length-scales into `ard_storage` [Q])."""
        storage[:2].copy_(self._hyp[:2])
        self._hyp = storage
        if self.ard and ard_storage is not None:
            ard_storage.copy_(self._opt_gamma_vec)
            self._opt_gamma_vec = ard_storage

    @property
    def opt_alpha(self):
        return self._hyp[0]

    @property
    def opt_gamma(self):
        return self._opt_gamma_vec.view(1, -1) if self.ard else self._hyp[1]

    def gamma_vec(self):
        """ARD length-scales [Q] (softplus of the optimisation variables)."""
        return ops.softplus(self._opt_gamma_vec)

    def scale(self, x):
        """x in kernel coordinates: x / gamma per latent dimension with ARD (gplvm.py:67), x itself otherwise (the scalar
        gamma is applied inside the CUDA kernels)."""
        if not self.ard or x is None:
            return x
        return (x / self.gamma_vec().view(1, -1)).contiguous()

    def scale_into(self, x, out):
        """scale(x) written into a persistent buffer (ARD): the LML+gradient plan keys its CUDA graphs on the operand
        addresses, and a fresh temporary per iteration would never be replayed."""
        if not self.ard:
            return x
        torch.div(x, self.gamma_vec().view(1, -1), out=out)
        return out

    def chain_scale(self, dXs, X):
        """(dL/dX, dL/dopt_gamma [Q]) from the gradient with respect to the scaled latents."""
        gam = self.gamma_vec().view(1, -1)
        d_gamma = -(dXs * X).sum(dim=0) / (gam.view(-1) ** 2)
        return dXs / gam, d_gamma * torch.sigmoid(self._opt_gamma_vec)

    def effective(self, noise_flags=0, fixed_noise=0.0, jitter=0.0):
        """Device tensor [alpha, gamma, noise] after the softplus / constant rules."""
        return ops.gp_effective_hyp(self._hyp, self.flags | noise_flags, fixed_noise, jitter)

    @property
    def alpha(self):
        return self.effective()[0]

    @property
    def gamma(self):
        return self.gamma_vec().view(1, -1) if self.ard else self.effective()[1]

    def state_tensors(self):
        if self.ard:
            return {'hyp': self._hyp, 'opt_gamma_ard': self._opt_gamma_vec}
        return {'hyp': self._hyp}

    def K(self, x, y=None, diag=False):
        """gplvm.py:84-86. diag=True with y=None is exactly alpha (gplvm.py:64-66)."""
        x = self.scale(to_device(x, self.device))
        hyp = self.effective()
        if diag and y is None:
            return hyp[0].expand(x.shape[0]).contiguous()
        full = ops.gram_rbf(x, self.scale(to_device(y, self.device)) if y is not None else None, hyp, add_noise=False)
        if diag:
            return torch.diagonal(full).contiguous()
        return full


class GPLVM(layer.Layer):
    def __init__(self, Y, Q, Y_labels=None, X0=None, kern=None, noise_variance=1.0, fixed_noise_variance=False,
                 noise_jitter=1e-6, bottom=None, x_test_var=None, subtract_Y_mean=True, uppropagate_Y=True,
                 latent_point_plotter=None, latent_sample_plotter=None, latent_space_explorer=None, session=None,
                 name='GPLVM', device='cuda'):
        assert type(Y) == np.ndarray, 'Y must be of type numpy.ndarray'
        assert Y.ndim == 2, 'Y must be 2-dimensional numpy.ndarray'
        self.Q = Q
        self.Y_labels = Y_labels
        self.subtract_Y_mean = subtract_Y_mean
        self.device = device

        if uppropagate_Y and bottom is not None:
            # the reference keeps this as a stochastic graph tensor re-sampled per run (gplvm.py:117-118,
            # SURVEY §8 a'.10, exercised by no example); here it is one draw
            Y_dev = to_device(bottom.upprop(Y.astype(np.float32)), device)
        else:
            Y_dev = to_device(Y, device)          # one host->device copy of the caller's array (no host-side passes over it)
        if self.subtract_Y_mean:
            # gplvm.py:122-124, on the device: the host version cost three passes over Y and 13-63 ms of a constructor call
            self.Y_mean = Y_dev.mean(dim=0, keepdim=True)
            Y_dev = Y_dev - self.Y_mean
        else:
            self.Y_mean = None
        self.Y = Y_dev
        self._Y_np = None

        if X0 is not None:
            self.X0 = np.asarray(X0, dtype=np.float32)
        else:
            self.X0 = preprocessing.standardize(preprocessing.PCA(self.Y_np, self.Q))   # gplvm.py:131-132
        assert self.X0.shape[0] == self.Y.shape[0]
        assert self.X0.shape[1] == self.Q
        self.N, self.D = self.X0.shape[0], int(self.Y.shape[1])

        # packed trainables [X | opt_alpha, opt_gamma, opt_noise | ARD opt_gamma [Q]]
        nq = self.N * self.Q
        n_ard = self.Q if (kern is not None and getattr(kern, 'ard', False)) else 0
        self._params = torch.zeros(nq + 3 + n_ard, dtype=torch.float32, device=device)
        self.X = self._params[:nq].view(self.N, self.Q)
        self.X.copy_(to_device(self.X0, device))
        self._hyp = self._params[nq:nq + 3]
        self.fixed_noise = float(noise_variance)
        if fixed_noise_variance is True:
            self.noise_jitter = None                                   # gplvm.py:141-143: constant, no jitter
            self._noise_flag = 0
        else:
            self.noise_jitter = noise_jitter
            self._noise_flag = _FLAG_NOISE
            self._hyp[2] = _inv_softplus(noise_variance)               # gplvm.py:145-152

        self.kern = kern or SEKernel(session=session, Q=self.Q, device=device)
        self.kern.bind(self._hyp, self._params[nq + 3:] if n_ard else None)

        if x_test_var is not None:
            self.x_test = x_test_var
            self.x_ph = x_test_var
        else:
            self.x_test = None
            self.x_ph = compat.Placeholder(shape=(None, self.Q), name='x_ph')

        self.lp_plotter = latent_point_plotter
        self.ls_plotter = latent_sample_plotter
        self.latent_space_explorer = latent_space_explorer
        self.eval_interval = None
        self.loss_values_fname = None
        self.loss_values = None

        self._plan = ops.GplvmPlan(self.N, self.Q, self.D, device=device)
        self._L = torch.zeros((self.N, self.N), dtype=torch.float32, device=device)
        self._S = None
        self._version = 0
        self._factor_version = -1
        params = {'X': self.X, 'noise_variance': lambda: self.noise_variance}
        super().__init__(params=params, bottom=bottom, session=session, name=name)

    @property
    def Y_np(self):
        """The (centred) observations as a host array, fetched on first use (PCA initialisation, diagnostics)."""
        if self._Y_np is None:
            self._Y_np = to_numpy(self.Y)
        return self._Y_np

    # ---- hyper-parameters -------------------------------------------------------------------------------------
    @property
    def flags(self):
        return self.kern.flags | self._noise_flag

    def _effective_hyp(self):
        return ops.gp_effective_hyp(self._hyp, self.flags, self.fixed_noise, self.noise_jitter or 0.0)

    @property
    def noise_variance(self):
        """softplus(opt_noise_variance), or the fixed constant (gplvm.py:143,152); jitter is added in build_K only."""
        if self._noise_flag:
            return ops.softplus(self._hyp[2:3])[0]
        return torch.tensor(self.fixed_noise, device=self.device)

    @property
    def opt_noise_variance(self):
        return self._hyp[2]

    def state_tensors(self):
        return {'params': self._params}

    def _after_restore(self):
        self._version += 1

    def set_latents(self, X):
        self.X.copy_(to_device(X, self.device))
        self._version += 1

    def param_dict(self):
        d = super().param_dict()
        d[self.kern.name] = self.kern.param_dict()
        return d

    # ---- model ---------------------------------------------------------------------------------------------------
    def _evaluate(self, want_grad=False, need_factor=False):
        Xs = self.X
        if self.kern.ard:
            if getattr(self, '_Xs', None) is None:
                self._Xs = torch.empty_like(self.X)
            Xs = self.kern.scale_into(self.X, self._Xs)
        scal, dX, info = self._plan.evaluate(Xs, self.Y, self._hyp, self.flags, fixed_noise=self.fixed_noise,
                                             jitter=self.noise_jitter or 0.0, x_prior_weight=0.0, include_const=True,
                                             want_grad=want_grad, L_out=self._L if need_factor else None)
        return scal, dX, info

    def _factor(self):
        """Cholesky factor and S = L^-1 Y for the current parameters (cached until they change)."""
        if self._factor_version != self._version:
            _, _, info = self._evaluate(want_grad=False, need_factor=True)
            pivot = int(info.item())
            if pivot:
                raise RuntimeError('Cholesky decomposition was not successful: pivot %d is not positive' % pivot)
            self._S = ops.trsm_lower(self._L, self.Y.clone())
            self._factor_version = self._version
        return self._L, self._S

    def build_model(self):
        """gplvm.py:210-216 — binds the lazily evaluated model quantities."""
        self.K = compat.Lazy(self.build_K)
        self.L = compat.Lazy(lambda: torch.tril(self._factor()[0]))
        self.log_det = compat.Lazy(self.build_log_det)
        self.loss = compat.Lazy(self.build_loss)
        self.pred_mean = compat.Lazy(self.build_pred_mean)
        self.pred_var = compat.Lazy(self.build_pred_var)

    def build_K(self):
        """K_f(X) + (noise [+ jitter]) I (gplvm.py:218-224)."""
        return ops.gram_rbf(self.kern.scale(self.X), None, self._effective_hyp(), add_noise=True)

    def build_log_det(self):
        """2 sum log L_ii (gplvm.py:226-228)."""
        return self._evaluate()[0][1].clone()

    def build_X_prior(self):
        return (self.X * self.X).sum()

    def build_loss(self, summary=True):
        """0.5 (|L^-1 Y|^2 + D logdet + D N log 2pi) (gplvm.py:233-244)."""
        scal, _, info = self._evaluate()
        return scal[0].clone()

    def loss_and_grad(self):
        """(loss, dX [N,Q], d/d(opt_alpha, opt_gamma, opt_noise) [3]) — K3-K6 in one call."""
        scal, dX, info = self._evaluate(want_grad=True)
        if self.kern.ard:                                          # (loss, dX, d/d(opt_alpha, opt_gamma [Q], opt_noise))
            dX, d_gam = self.kern.chain_scale(dX, self.X)
            return scal[0].clone(), dX, torch.cat([scal[3:4], d_gam, scal[5:6]])
        return scal[0].clone(), dX.clone(), scal[3:6].clone()     # the plan's buffers are reused by the next call

    def _xstar(self, x=None):
        if x is None:
            x = self.x_ph.value
        assert x is not None, 'no test latent points: feed x_ph or pass x_test_var'
        return self.kern.scale(to_device(np.array(to_numpy(x), ndmin=2) if not isinstance(x, torch.Tensor) else x, self.device))

    def build_pred_mean(self, x=None):
        """(L^-1 K_Xx)^T (L^-1 Y) + Y_mean (gplvm.py:246-254)."""
        Lf, S = self._factor()
        mean, _, _ = ops.gp_predict(self.kern.scale(self.X), Lf, S, self.Y_mean, self._xstar(x), self._effective_hyp(),
                                    want_var=False)
        return mean

    def build_pred_var(self, x=None):
        """alpha - colsum((L^-1 K_Xx)^2); noise not added back; asserts >= 0 (gplvm.py:278-288)."""
        Lf, S = self._factor()
        _, var, info = ops.gp_predict(self.kern.scale(self.X), Lf, None, None, self._xstar(x), self._effective_hyp(),
                                      want_mean=False)
        bad = int(info.item())
        if bad:
            raise RuntimeError('assert_non_negative(pred_var) failed at test point %d' % (bad - 1))
        return var

    def build_pred_mean_thresholded(self, pred_mean):
        return torch.tanh(pred_mean)

    def build_pred_mean_normalized(self, pred_mean):
        lo, hi = pred_mean.min(), pred_mean.max()
        return (pred_mean - lo) / (hi - lo)

    def predict_mean(self, x):
        return to_numpy(self.build_pred_mean(np.array(x, ndmin=2)))

    def predict_variance(self, x):
        return to_numpy(self.build_pred_var(np.array(x, ndmin=2)))[:, np.newaxis]

    def datapoint_len(self):
        if self.bottom:
            return self.bottom.layers[-1].num_v
        return self.D

    def downprop_latent(self, x, num_samples_to_average=100):
        """Down-propagate the predictive mean and average (gplvm.py:314-338)."""
        assert type(x) == np.ndarray
        assert x.ndim == 1 or x.ndim == 2
        if x.ndim == 1:
            x = x[np.newaxis, :]
        mean = self.build_pred_mean(x)
        if not self.bottom:
            return to_numpy(mean)[0] if x.shape[0] == 1 else to_numpy(mean)
        if x.shape[0] == 1:
            tiled = mean.expand(num_samples_to_average, -1).contiguous()
            return to_numpy(self.bottom.downprop(tiled)).mean(axis=0)
        acc = torch.zeros((x.shape[0], self.datapoint_len()), dtype=torch.float32, device=self.device)
        for _ in range(num_samples_to_average):
            acc += self.bottom.downprop(mean)
        return to_numpy(acc) / num_samples_to_average

    # ---- plotting hooks (accepted, off the hot path) ----------------------------------------------------------------
    def plot_latent_points(self, iteration=None):
        assert self.lp_plotter is not None
        var_diag = to_numpy(self.build_pred_var(self.lp_plotter.grid))
        fig_name = 'latent_points' + ('' if iteration is None else '-' + str(iteration))
        self.lp_plotter.plot(to_numpy(self.X), np.log(var_diag), fig_name, self.Y_labels)

    def plot_latent_samples(self, iteration=None):
        assert self.ls_plotter is not None
        v_vecs = self.downprop_latent(self.ls_plotter.grid, self.ls_plotter.num_samples_to_average)
        fig_name = 'latent_samples' + ('' if iteration is None else '-' + str(iteration))
        self.ls_plotter.plot(v_vecs, fig_name)

    def explore_latent_space(self):
        assert self.latent_space_explorer is not None
        self.latent_space_explorer.set_sample_callback(self.downprop_latent)
        self.latent_space_explorer.set_variance_callback(self.predict_variance)
        self.latent_space_explorer.plot()

    def save_loss_values(self, loss, iteration):
        idx = iteration // self.eval_interval
        self.loss_values[idx] = (iteration, loss)
        np.save(self.loss_values_fname, self.loss_values[:idx + 1])

    def _eval_step(self, i):
        loss = float(self.build_loss())
        self.logger.info('Iteration: %d, Loss: %f', i, loss)
        self.logger.debug(self)
        if self.loss_values_fname:
            self.save_loss_values(loss, i)
        if self.lp_plotter:
            self.plot_latent_points(i)
        if self.ls_plotter:
            self.plot_latent_samples(iteration=i)
        return loss

    # ---- optimisation ---------------------------------------------------------------------------------------------------
    def optimize(self, optimizer, num_iterations=100, eval_interval=0, ckpt_interval=0, ckpt_dir=None,
                 summary_writer=None, loss_values_fname=None):
        """Gradient loop over {X, opt_noise, opt_alpha, opt_gamma} (gplvm.py:414-468): each iteration is
        one fused LML+gradient evaluation and one optimizer kernel on the packed parameters."""
        optimizer.reset()                                   # fresh slots (gplvm.py:386-396, 430)
        state = torch.zeros(max(optimizer.state_multiple(), 1) * self._params.numel(), dtype=torch.float32,
                            device=self.device)
        grad = torch.zeros_like(self._params)
        nq = self.N * self.Q
        i = 0
        if eval_interval:
            self.eval_interval = eval_interval
            if loss_values_fname:
                self.loss_values_fname = loss_values_fname
                self.loss_values = np.zeros(shape=(int(np.floor(num_iterations / eval_interval) + 1), 2))
            self._eval_step(i)
        try:
            for i in range(1, num_iterations + 1):
                scal, dX, info = self._evaluate(want_grad=True)
                if self.kern.ard:
                    dX, d_gam = self.kern.chain_scale(dX, self.X)
                    grad[nq + 3:].copy_(d_gam)
                grad[:nq].copy_(dX.reshape(-1))
                grad[nq:nq + 3].copy_(scal[3:6])
                ops.optimizer_step(optimizer.descriptor(), self._params, grad, state)
                self._version += 1
                at_eval = eval_interval and i % eval_interval == 0
                at_ckpt = ckpt_dir and ckpt_interval and i % ckpt_interval == 0
                # the factorisation status is a device scalar: reading it stalls the launch queue, so it is read every
                # INFO_INTERVAL iterations and before anything is logged or saved (the reference's tf.cholesky raises
                # InvalidArgumentError at the failing iteration, gplvm.py:212)
                if at_eval or at_ckpt or i % self.INFO_INTERVAL == 0:
                    self._raise_if_not_pd()
                if at_eval:
                    self._eval_step(i)
                if at_ckpt:
                    self.save_all(i, ckpt_dir)
                    self.kern.save(i, ckpt_dir)
        except KeyboardInterrupt:
            self.logger.info('Training interrupted')
            if eval_interval:
                self._eval_step(i)
            if ckpt_dir and ckpt_interval:
                self.save_all(i, ckpt_dir)
                self.kern.save(i, ckpt_dir)
        self._raise_if_not_pd()

    INFO_INTERVAL = 32

    def _raise_if_not_pd(self):
        pivot = int(self._plan.info.item())
        if pivot:
            raise RuntimeError('Cholesky decomposition was not successful: pivot %d is not positive' % pivot)

    def _latent_search_cache(self):
        """K^-1 and A = K^-1 Y for the current (frozen) parameters — the reference re-runs tf.cholesky inside every
        one of its ~num_points * num_runs * num_iterations session.run calls."""
        Lf, S = self._factor()
        eye = torch.eye(self.N, device=self.device, dtype=torch.float32)
        Kinv = ops.trsm_lower(Lf, ops.trsm_lower(Lf, eye), trans=True)
        A = ops.trsm_lower(Lf, S.clone(), trans=True)
        return Kinv, A

    def latent_search_loss(self, x, test_point, log_var_scaling=0.1, cache=None, want_outputs=False):
        """(loss [R], d loss / d x [R,Q], ...) of gplvm.py:478-483 at restarts x [R,Q] for one test point."""
        Kinv, A = cache or self._latent_search_cache()
        out = ops.testgen_eval(self.kern.scale(self.X), Kinv, A, self.Y_mean, self._effective_hyp(),
                               self.kern.scale(to_device(x, self.device)), to_device(test_point, self.device).reshape(-1),
                               log_var_scaling, want_outputs)
        if self.kern.ard:                                          # d/dx = d/d(x / gamma) / gamma
            out = (out[0], out[1] / self.kern.gamma_vec().view(1, -1)) + tuple(out[2:])
        return out

    def test_generalisation(self, test_data, output_table_path, optimizer, log_var_scaling=0.1, num_iterations=1000,
                            num_runs=1):
        """Latent-point search for test images (gplvm.py:470-541). For every test point, `num_runs` restarts (all
        training latents when num_runs == N, else random training latents) are optimised for `num_iterations` steps
        on loss = CE(normalised(pred_mean(x)), test_point) / Dim + log_var_scaling * log(pred_var(x)); the best run is
        kept. The restarts of one test point are advanced TOGETHER as a batch [R,Q] with K^-1 and K^-1 Y cached —
        the arithmetic per restart is the reference's; optimizer slots start fresh per test point (the reference
        keeps one set of slots across runs, which only matters for stateful optimizers)."""
        test_data = np.asarray(test_data, dtype=np.float32)
        assert num_runs <= self.N
        cache = self._latent_search_cache()
        rs = np.random.RandomState(0)                              # tf.random_uniform(seed=0) [TF stream not reproduced]
        table = []
        for test_point_index in range(test_data.shape[0]):
            test_point = test_data[test_point_index:test_point_index + 1]
            if num_runs == self.N:
                idx = np.arange(self.N)
            else:
                idx = rs.randint(0, max(self.N - 1, 1), size=num_runs)        # maxval = N - 1, exclusive (gplvm.py:497-499)
            x = self.X[torch.as_tensor(idx, device=self.device)].clone().contiguous()
            optimizer.reset()
            state = torch.zeros(max(optimizer.state_multiple(), 1) * x.numel(), dtype=torch.float32, device=self.device)
            for _ in range(num_iterations):
                _, g_x, _, _, _ = self.latent_search_loss(x, test_point, log_var_scaling, cache)
                ops.optimizer_step(optimizer.descriptor(), x.view(-1), g_x.view(-1), state)
            loss, _, info, pm_norm, pm_clip = self.latent_search_loss(x, test_point, log_var_scaling, cache, want_outputs=True)
            bad = int(info.item())
            if bad:
                raise RuntimeError('assert_non_negative(pred_var) failed for restart %d' % (bad - 1))
            dist_np = to_numpy(loss)
            best = int(np.argmin(dist_np))                          # first minimum = the reference's strict '<' sweep
            for run in range(len(idx)):
                self.logger.info('Test point: {}, Run: {}, Distance: {}'.format(test_point_index, run, dist_np[run]))
            table.append((test_point_index, float(dist_np[best]), to_numpy(x[best:best + 1]), test_point,
                          to_numpy(pm_norm[best:best + 1]), to_numpy(pm_clip[best:best + 1])))
        with open(output_table_path, 'wb') as f:
            pickle.dump(table, f)
        return table
