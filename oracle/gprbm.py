"""Oracle restatement of the GPDBN top layer (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows /root/reference/deepbelief/layers/gprbm.py in TRAINING mode (eval_flag False, x_ph = X,
gprbm.py:106) with every random draw injected:
  forward  = gprbm.py:108-115 (K, L, pred_var, sigma_h), :140-147 (H1 = upprop(V), H1_mean, H2),
             :220-232 (build_sample_mean), :234-256 (gp loss + cross-entropy of the down-propagated sample),
             through a stack of lower RBMs with Concrete units (rbm.py:243-275, 277-325) or Bernoulli units;
  backward = the gradient TF autodiff produces for optimizer.minimize(loss) (gprbm.py:386-396), written
             as explicit VJPs in numpy (paragraph after GP14) and pinned by directional finite
             differences of the REFERENCE's own loss (tests/golden/make_golden.py -> gprbm/* goldens).
The forward follows the reference's formulas literally (triangular solves, expanded squared distances);
the training-mode shortcuts used by the CUDA path are NOT used here.
"""
import numpy as np
import scipy.linalg

from .rbm import EPS0, sigmoid, softplus
from .gp import se_kernel, sq_dist


def _logit_eps(p, e0):
    return np.log(p + e0) - np.log(1.0 - p + e0)


def concrete(p, u, temperature):
    """rbm.py:243-258."""
    e0 = p.dtype.type(EPS0)
    return sigmoid((_logit_eps(p, e0) + _logit_eps(u, e0)) / p.dtype.type(temperature))


def concrete_vjp(g, p, a, temperature):
    """d concrete / d p, times g. a = concrete(p, u, T)."""
    e0 = p.dtype.type(EPS0)
    return g * a * (1.0 - a) / p.dtype.type(temperature) * (1.0 / (p + e0) + 1.0 / (1.0 - p + e0))


class GPDBNOracle:
    """lower: list of dicts {W [V,H], b [V], c [H]} from the bottom up (Concrete units with `temperature`);
    top: dict {W [Vt,D], b [Vt], c [D], opt_alpha_h [D]}; gp: dict {X [N,Q], opt_alpha, opt_gamma, opt_noise,
    trainable (3 bools), fixed_noise, jitter}."""

    def __init__(self, lower, top, gp, V_data, temperature, dtype=np.float64):
        self.dt = dtype
        c = lambda a: np.asarray(a, dtype=dtype)  # noqa: E731
        self.lower = [{k: c(v) for k, v in l.items()} for l in lower]
        self.top = {k: c(v) for k, v in top.items()}
        self.gp = dict(gp)
        self.gp['X'] = c(gp['X'])
        self.V = c(V_data)
        self.T = temperature
        self.N = self.V.shape[0]
        self.D = self.top['W'].shape[1]

    # ---- hyper-parameters (gplvm.py:31-49, gprbm.py:62-73, 194-196) -----------------------------------------
    def hyp(self):
        g = self.gp
        tr = g.get('trainable', (False, False, False))
        alpha = softplus(np.asarray(g['opt_alpha'], dtype=self.dt)) if tr[0] else self.dt(1.0)
        gamma = softplus(np.asarray(g['opt_gamma'], dtype=self.dt)) if tr[1] else self.dt(1.0)
        if tr[2]:
            noise = softplus(np.asarray(g['opt_noise'], dtype=self.dt)) + self.dt(g.get('jitter', 1e-6))
        else:
            noise = self.dt(g['fixed_noise'])
        return alpha, gamma, noise

    def forward(self, draws):
        """draws: dict with 'up' (list of U per lower layer), 'eps1', 'eps2' [N,D], 'down_top' U [N,Vt],
        'down' (list of U per lower layer, TOP-most lower layer first — the order downprop visits them)."""
        dt = self.dt
        T = self.T
        alpha, gamma, noise = self.hyp()
        X = self.gp['X']
        N, D = self.N, self.D
        st = {}
        Kf = se_kernel(X, None, alpha, gamma)                         # gplvm.py:84-86 via gprbm.py:197
        K = Kf + noise * np.eye(N, dtype=dt)                          # gprbm.py:193-199
        L = np.linalg.cholesky(K)                                     # gprbm.py:109
        KXx = se_kernel(X, X, alpha, gamma)                           # x_ph = X (gprbm.py:106, 209)
        sys1 = scipy.linalg.solve_triangular(L, KXx, lower=True)      # gprbm.py:210
        pv = se_kernel(X, None, alpha, gamma, diag=True) - np.sum(np.square(sys1), axis=0)   # gprbm.py:211-214
        if np.any(pv < 0):
            raise FloatingPointError('assert_non_negative(pred_var) failed')                  # gprbm.py:215
        r = np.sqrt(pv)
        alpha_h = softplus(self.top['opt_alpha_h']).reshape(1, D)     # gprbm.py:83
        sigma_h = alpha_h * r[:, None]                                # gprbm.py:114-115
        # ---- H1 = upprop(V_data) (gprbm.py:140; rbm.py:277-300) ----
        a = self.V
        ups = []
        for l, lay in enumerate(self.lower):
            p = sigmoid(a @ lay['W'] + lay['c'].reshape(1, -1))
            if T is not None:
                nxt = concrete(p, np.asarray(draws['up'][l], dtype=dt), T)
            else:
                nxt = (p > np.asarray(draws['up'][l], dtype=dt)).astype(dt)
            ups.append((a, p, nxt))
            a = nxt
        x = a @ self.top['W']
        eps1 = np.asarray(draws['eps1'], dtype=dt)
        H1 = sigma_h * x + self.top['c'].reshape(1, D) + sigma_h * eps1   # bgrbm.py:79-82, 107-110
        H1_mean = np.mean(H1, axis=0, keepdims=True)                  # gprbm.py:146
        H2 = (H1 - H1_mean) / alpha_h                                 # gprbm.py:147
        # ---- sample mean (gprbm.py:220-232) ----
        sys2 = scipy.linalg.solve_triangular(L, H2, lower=True)
        Hm = sys1.T @ sys2
        eps2 = np.asarray(draws['eps2'], dtype=dt)
        H = Hm * alpha_h + H1_mean + sigma_h * eps2
        # ---- downprop (bgrbm.py:116-131, rbm.py:302-325) ----
        y = H / sigma_h
        zt = (self.top['W'] @ y.T + self.top['b'].reshape(-1, 1)).T
        q = sigmoid(zt)
        if T is not None:
            d = concrete(q, np.asarray(draws['down_top'], dtype=dt), T)
        else:
            d = (q > np.asarray(draws['down_top'], dtype=dt)).astype(dt)
        downs = [(y, q, d)]
        for i, lay in enumerate(reversed(self.lower)):
            z = (lay['W'] @ d.T + lay['b'].reshape(-1, 1)).T
            q = sigmoid(z)
            if T is not None:
                nd = concrete(q, np.asarray(draws['down'][i], dtype=dt), T)
            else:
                nd = (q > np.asarray(draws['down'][i], dtype=dt)).astype(dt)
            downs.append((d, q, nd))
            d = nd
        V_sample = d
        # ---- losses (dist.py:24-35 reduce_mean=False; gprbm.py:234-256) ----
        e0 = dt(EPS0)
        pc = np.clip(V_sample, e0, 1 - e0)
        logits = np.log(pc / (1 - pc))
        ce = np.sum(np.maximum(logits, 0) - logits * self.V + np.log1p(np.exp(-np.abs(logits))))
        logdet = 2.0 * np.sum(np.log(np.diag(L)))
        ssq = np.sum(np.square(sys2))
        gp_loss = 0.5 * (ssq + D * logdet) + np.sum(np.square(X))
        st.update(dict(alpha=alpha, gamma=gamma, noise=noise, Kf=Kf, K=K, L=L, pv=pv, r=r, alpha_h=alpha_h,
                       sigma_h=sigma_h, ups=ups, x=x, eps1=eps1, eps2=eps2, H1=H1, H1_mean=H1_mean, H2=H2, Hm=Hm,
                       H=H, downs=downs, V_sample=V_sample, pc=pc, ce=ce, gp_loss=gp_loss, logdet=logdet, ssq=ssq,
                       loss=ce + gp_loss))
        return st

    def loss(self, draws):
        return float(self.forward(draws)['loss'])

    def grads(self, draws, clip_mask=None):
        """Gradient of `loss` wrt every trainable: returns (state, dict) with keys lower[l].{W,b,c},
        top.{W,b,c,opt_alpha_h}, X, opt_alpha, opt_gamma, opt_noise.
        clip_mask (optional bool [N,V]): which elements of V_sample count as strictly inside the clip interval of
        dist.py:25-26. The mask is a step function of V_sample at 1e-6 / 1 - 1e-6, where fp32 resolves only ~16
        values per 1e-6: a parity run passes the mask its fp32 evaluation produced (a tie band, like |p - u| for
        Bernoulli samples) and checks separately that the two masks differ in few elements."""
        st = self.forward(draws)
        dt, T, N, D = self.dt, self.T, self.N, self.D
        X = self.gp['X']
        e0 = dt(EPS0)
        alpha_h, sigma_h, r = st['alpha_h'], st['sigma_h'], st['r']
        out = {'lower': [dict(W=np.zeros_like(l['W']), b=np.zeros_like(l['b']), c=np.zeros_like(l['c']))
                         for l in self.lower]}
        # cross entropy wrt the clipped sample; clip passes gradient only strictly inside (tf.clip_by_value)
        pc, Vs = st['pc'], st['V_sample']
        inside = ((Vs >= e0) & (Vs <= 1 - e0)) if clip_mask is None else np.asarray(clip_mask, dtype=bool)
        st['clip_mask'] = (Vs >= e0) & (Vs <= 1 - e0)
        g = (-self.V / pc + (1.0 - self.V) / (1.0 - pc)) * inside
        # ---- down pass, lowest layer first ----
        downs = st['downs']
        for i in range(len(self.lower) - 1, -1, -1):
            lay_index = len(self.lower) - 1 - i              # downs[i+1] belongs to lower[::-1][i]
            lay = self.lower[lay_index]
            d_in, q, d_out = downs[i + 1]
            if T is None:
                g = np.zeros_like(q)
                continue
            dz = concrete_vjp(g, q, d_out, T) * q * (1 - q)
            out['lower'][lay_index]['W'] += dz.T @ d_in
            out['lower'][lay_index]['b'] += dz.sum(axis=0)
            g = dz @ lay['W']
        y, q, d_out = downs[0]
        gt = {'W': np.zeros_like(self.top['W']), 'b': np.zeros_like(self.top['b']), 'c': np.zeros_like(self.top['c'])}
        if T is not None:
            dz = concrete_vjp(g, q, d_out, T) * q * (1 - q)
        else:
            dz = np.zeros_like(q)
        gt['W'] += dz.T @ y
        gt['b'] += dz.sum(axis=0)
        g_y = dz @ self.top['W']
        g_H = g_y / sigma_h
        g_sig = -g_y * y / sigma_h
        # H = Hm * alpha_h + H1_mean + sigma_h * eps2
        g_Hm = g_H * alpha_h
        g_alpha_h = np.sum(g_H * st['Hm'], axis=0, keepdims=True)
        g_H1mean = np.sum(g_H, axis=0, keepdims=True)
        g_sig = g_sig + g_H * st['eps2']
        # ---- GP block, part 2: (K, noise, H2) -> Hm, gp_loss ; exact forms with Kinv ----
        alpha, gamma, noise = st['alpha'], st['gamma'], st['noise']
        Kinv = np.linalg.inv(st['K'])
        Kf = st['Kf']
        C = Kinv @ Kf                                   # K^-1 K_Xx
        B = Kinv @ st['H2']
        g_Kf = g_Hm @ B.T                               # Hm = Kf^T K^-1 H2 (Kf symmetric: the K_Xx slot)
        g_Bv = Kf @ g_Hm
        g_H2 = Kinv @ g_Bv + B                          # + d(0.5 * H2^T K^-1 H2)
        g_K = -(Kinv @ g_Bv) @ B.T + 0.5 * (D * Kinv - B @ B.T)
        # H2 = (H1 - mean) / alpha_h
        g_alpha_h = g_alpha_h - np.sum(g_H2 * st['H2'], axis=0, keepdims=True) / alpha_h
        g_H1 = g_H2 / alpha_h
        g_H1mean = g_H1mean - np.sum(g_H1, axis=0, keepdims=True)
        g_H1 = g_H1 + g_H1mean / N
        # H1 = sigma_h * (x + eps1) + c
        g_x = g_H1 * sigma_h
        g_sig = g_sig + g_H1 * (st['x'] + st['eps1'])
        gt['c'] += g_H1.sum(axis=0)
        a_top = st['ups'][-1][2] if self.lower else self.V
        gt['W'] += a_top.T @ g_x
        g_a = g_x @ self.top['W'].T
        # sigma_h = alpha_h * r
        g_alpha_h = g_alpha_h + np.sum(g_sig * r[:, None], axis=0, keepdims=True)
        g_r = np.sum(g_sig * alpha_h, axis=1)
        g_pv = g_r / (2.0 * r)
        gt['opt_alpha_h'] = (g_alpha_h * sigmoid(self.top['opt_alpha_h'].reshape(1, D))).reshape(-1)
        # ---- GP block, part 1: pv = alpha - colsum((L^-1 K_Xx)^2) = alpha - diag(Kf K^-1 Kf) ----
        Dg = np.diag(g_pv)
        g_Kf = g_Kf - Dg @ C.T - C @ Dg                 # d/dKf of -diag(Kf^T K^-1 Kf), both Kf slots
        g_K = g_K + C @ Dg @ C.T
        g_alpha_direct = np.sum(g_pv)                   # Kxx = alpha
        g_Kf = g_Kf + g_K
        g_noise = np.trace(g_K)
        S = g_Kf + g_Kf.T                               # each K_ij depends on x_i and x_j
        Wm = S * Kf
        g2 = gamma * gamma
        out['X'] = -(np.sum(Wm, axis=1, keepdims=True) * X - Wm @ X) / g2 + 2.0 * X
        diff2 = sq_dist(X, None, 1.0)
        tr = self.gp.get('trainable', (False, False, False))
        out['opt_alpha'] = ((np.sum(g_Kf * Kf) / alpha + g_alpha_direct) * sigmoid(np.asarray(self.gp['opt_alpha'], dtype=dt))
                            if tr[0] else dt(0))
        out['opt_gamma'] = (np.sum(g_Kf * Kf * diff2) / gamma ** 3 * sigmoid(np.asarray(self.gp['opt_gamma'], dtype=dt))
                            if tr[1] else dt(0))
        out['opt_noise'] = g_noise * sigmoid(np.asarray(self.gp['opt_noise'], dtype=dt)) if tr[2] else dt(0)
        # ---- up pass ----
        for l in range(len(self.lower) - 1, -1, -1):
            a_in, p, a_out = st['ups'][l]
            if T is None:
                break
            dz = concrete_vjp(g_a, p, a_out, T) * p * (1 - p)
            out['lower'][l]['W'] += a_in.T @ dz
            out['lower'][l]['c'] += dz.sum(axis=0)
            g_a = dz @ self.lower[l]['W'].T
        out['top'] = gt
        return st, out


def latent_search_loss(oracle, H2, H1_mean, x, target, log_var_scaling, num_samples, draws):
    """Per-restart loss of GPRBM.test_generalisation (gprbm.py:490-499) at test latents x [R,Q], each row its own
    M=1 problem as in the reference: build_sample_mean(H2_var, H1_mean_var, num_samples) (gprbm.py:220-232) with
    sigma_h = alpha_h * sqrt(pred_var(x)) (gprbm.py:114-115), downprop through the Concrete stack, cross entropy of
    the sample MEAN against the test point / Dim + log_var_scaling * log(pred_var).
    draws: {'eps' [R*S,D], 'down_top' [R*S,Vt], 'down': [per lower layer, top-most first]}, rows ordered r*S + s."""
    o, dt, T = oracle, oracle.dt, oracle.T
    alpha, gamma, noise = o.hyp()
    X = o.gp['X']
    x = np.asarray(x, dtype=dt)
    R, S, D = x.shape[0], int(num_samples), o.D
    H2 = np.asarray(H2, dtype=dt)
    H1_mean = np.asarray(H1_mean, dtype=dt).reshape(1, D)
    target = np.asarray(target, dtype=dt).reshape(1, -1)
    K = se_kernel(X, None, alpha, gamma) + noise * np.eye(o.N, dtype=dt)
    L = np.linalg.cholesky(K)
    alpha_h = softplus(o.top['opt_alpha_h']).reshape(1, D)
    sys2 = scipy.linalg.solve_triangular(L, H2, lower=True)
    out = np.zeros(R, dtype=dt)
    e0 = dt(EPS0)
    for r in range(R):
        KXx = se_kernel(X, x[r:r + 1], alpha, gamma)
        sys1 = scipy.linalg.solve_triangular(L, KXx, lower=True)
        pv = se_kernel(x[r:r + 1], None, alpha, gamma, diag=True) - np.sum(np.square(sys1), axis=0)     # [1]
        sigma_h = alpha_h * np.sqrt(pv)[:, None]                                                      # [1,D]
        H = (sys1.T @ sys2) * alpha_h + H1_mean                                                       # [1,D]
        rows = slice(r * S, (r + 1) * S)
        H = H + sigma_h * np.asarray(draws['eps'], dtype=dt)[rows]                                    # [S,D]
        y = H / sigma_h
        q = sigmoid((o.top['W'] @ y.T + o.top['b'].reshape(-1, 1)).T)
        d = concrete(q, np.asarray(draws['down_top'], dtype=dt)[rows], T)
        for i, lay in enumerate(reversed(o.lower)):
            q = sigmoid((lay['W'] @ d.T + lay['b'].reshape(-1, 1)).T)
            d = concrete(q, np.asarray(draws['down'][i], dtype=dt)[rows], T)
        m = np.mean(d, axis=0, keepdims=True)
        pc = np.clip(m, e0, 1 - e0)
        lg = np.log(pc / (1 - pc))
        ce = np.sum(np.maximum(lg, 0) - lg * target + np.log1p(np.exp(-np.abs(lg))))
        out[r] = ce / target.shape[1] + log_var_scaling * np.log(pv[0])
    return out


def _latent_down(o, L, sys2, H1_mean, alpha_h, x_r, rows, draws):
    """One test latent x_r [1,Q]: pred_var (gprbm.py:208-218), sigma_h = alpha_h sqrt(pred_var) (gprbm.py:114-115),
    build_sample_mean with the noise rows `rows` of draws['eps'] (gprbm.py:220-232) and downprop through the Concrete
    stack (rbm.py:302-325). Returns (down-propagated rows [S,V0], pred_var [1])."""
    dt, T = o.dt, o.T
    alpha, gamma, _ = o.hyp()
    X = o.gp['X']
    KXx = se_kernel(X, x_r, alpha, gamma)
    sys1 = scipy.linalg.solve_triangular(L, KXx, lower=True)
    pv = se_kernel(x_r, None, alpha, gamma, diag=True) - np.sum(np.square(sys1), axis=0)
    sigma_h = alpha_h * np.sqrt(pv)[:, None]
    H = (sys1.T @ sys2) * alpha_h + H1_mean
    H = H + sigma_h * np.asarray(draws['eps'], dtype=dt)[rows]
    y = H / sigma_h
    q = sigmoid((o.top['W'] @ y.T + o.top['b'].reshape(-1, 1)).T)
    d = concrete(q, np.asarray(draws['down_top'], dtype=dt)[rows], T)
    for i, lay in enumerate(reversed(o.lower)):
        q = sigmoid((lay['W'] @ d.T + lay['b'].reshape(-1, 1)).T)
        d = concrete(q, np.asarray(draws['down'][i], dtype=dt)[rows], T)
    return d, pv


def _frozen_gp(o, H2, H1_mean):
    dt = o.dt
    alpha, gamma, noise = o.hyp()
    X = o.gp['X']
    K = se_kernel(X, None, alpha, gamma) + noise * np.eye(o.N, dtype=dt)
    L = np.linalg.cholesky(K)
    alpha_h = softplus(o.top['opt_alpha_h']).reshape(1, o.D)
    sys2 = scipy.linalg.solve_triangular(L, np.asarray(H2, dtype=dt), lower=True)
    return L, sys2, np.asarray(H1_mean, dtype=dt).reshape(1, o.D), alpha_h


def held_out_loss(oracle, H2, H1_mean, x, test_data, draws):
    """The objective of GPRBM.test_held_out_data (gprbm.py:440-466): dist.mse(test_data, downprop(sample_mean(x_test)))
    with the [M,V] test set as `predictions` and the ONE down-propagated row (num_samples None, x_test [1,Q]) broadcast
    as `targets` (dist.py:5-21): mean over the M test rows of sum_v (t_mv - d_v)^2. x may hold R independent rows;
    row r uses row r of every draw. Returns [R]."""
    o, dt = oracle, oracle.dt
    L, sys2, H1m, alpha_h = _frozen_gp(o, H2, H1_mean)
    x = np.asarray(x, dtype=dt)
    Tt = np.asarray(test_data, dtype=dt)
    out = np.zeros(x.shape[0], dtype=dt)
    for r in range(x.shape[0]):
        d, _ = _latent_down(o, L, sys2, H1m, alpha_h, x[r:r + 1], slice(r, r + 1), draws)
        out[r] = np.mean(np.sum(np.square(Tt - d), axis=1))
    return out


def interp_path(x_start, x_end, num_points):
    """The initial (linear) interpolation latents of GPRBM.interp_test (gprbm.py:623-637): num_points interior points
    plus both end points, [num_points + 2, Q]."""
    x_s = np.reshape(x_start, (1, -1))
    x_e = np.reshape(x_end, (1, -1))
    a = np.linspace(0.0, 1.0, num_points + 2)[1:-1, np.newaxis]
    a = np.insert(a, 0, 0)
    a = np.append(a, 1.0)
    a = np.reshape(a, (num_points + 2, 1))
    return x_s + np.dot(a, x_e - x_s)


def interp_objective(oracle, Xi, x_start, x_end, interp_lambda):
    """gprbm.py:651-660: mean_i |segment_i|^2 over the P+1 segments start -> Xi[0] -> ... -> Xi[-1] -> end, plus
    lambda * mean(log pred_var(Xi)) with pred_var as build_pred_var_interp (gprbm.py:569-579). Returns (objective,
    pred_var [P])."""
    o, dt = oracle, oracle.dt
    alpha, gamma, noise = o.hyp()
    X = o.gp['X']
    Xi = np.asarray(Xi, dtype=dt)
    x_s = np.asarray(x_start, dtype=dt).reshape(1, -1)
    x_e = np.asarray(x_end, dtype=dt).reshape(1, -1)
    K = se_kernel(X, None, alpha, gamma) + noise * np.eye(o.N, dtype=dt)
    L = np.linalg.cholesky(K)
    KXx = se_kernel(X, Xi, alpha, gamma)
    sys1 = scipy.linalg.solve_triangular(L, KXx, lower=True)
    pv = se_kernel(Xi, None, alpha, gamma, diag=True) - np.sum(np.square(sys1), axis=0)
    seg = np.concatenate([Xi[0:1] - x_s, Xi[1:] - Xi[:-1], x_e - Xi[-1:]], axis=0)
    dist_term = np.mean(np.sum(np.square(seg), axis=1))
    return dist_term + dt(interp_lambda) * np.mean(np.log(pv)), pv


def interp_images(oracle, H2, H1_mean, Xi, num_samples, draws):
    """create_observed_prediction_with_variance (gprbm.py:564-612): per interpolation point, sigma_h = alpha_h *
    sqrt(pred_var_i), `num_samples` noise rows (rows i*S + s of the draws), downprop, mean over the samples.
    Returns (images [P,V0], pred_var [P])."""
    o, dt = oracle, oracle.dt
    L, sys2, H1m, alpha_h = _frozen_gp(o, H2, H1_mean)
    Xi = np.asarray(Xi, dtype=dt)
    S = int(num_samples)
    imgs, pvs = [], []
    for i in range(Xi.shape[0]):
        d, pv = _latent_down(o, L, sys2, H1m, alpha_h, Xi[i:i + 1], slice(i * S, (i + 1) * S), draws)
        imgs.append(np.mean(d, axis=0, keepdims=True))
        pvs.append(pv[0])
    return np.concatenate(imgs, axis=0), np.asarray(pvs, dtype=dt)
