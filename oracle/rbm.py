"""Oracle restatement of the RBM family (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows /root/reference/deepbelief/layers/rbm.py, bgrbm.py, gbrbm.py, ggrbm.py line by line, in numpy,
with a dtype switch and all randomness injected. Layer kinds: 'rbm', 'bgrbm', 'gbrbm', 'ggrbm'.
"""
import numpy as np

EPS0 = 10e-7   # rbm.py:249 / dist.py:25 — note: 1e-6, not 1e-7


def sigmoid(x):
    # numerically stable logistic in the array's own dtype (tf.sigmoid)
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    ex = np.exp(x[~pos])
    out[~pos] = ex / (1.0 + ex)
    return out


def softplus(x):
    # tf.nn.softplus = log(1 + exp(x)), stable form
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def inverse_softplus(x):
    # preprocessing.py:171-172
    return np.log(np.exp(x) - 1.0)


class RBMOracle:
    """Parameters W[V,H], b[V,1], c[1,H] (rbm.py:56-72) and, for Gaussian hiddens, opt_sigma_h[1,H]
    with sigma_h = softplus(opt_sigma_h) (bgrbm.py:44-51)."""

    def __init__(self, W, b, c, kind='rbm', opt_sigma_h=None, temperature=None, dtype=np.float64, sigma_h=None):
        self.dtype = dtype
        self.kind = kind
        self.W = np.asarray(W, dtype=dtype)
        self.V, self.H = self.W.shape
        self.b = np.asarray(b, dtype=dtype).reshape(self.V, 1)
        self.c = np.asarray(c, dtype=dtype).reshape(1, self.H)
        self.temperature = temperature
        self.gaussian_hidden = kind in ('bgrbm', 'ggrbm')
        self.gaussian_visible = kind in ('gbrbm', 'ggrbm')
        if self.gaussian_hidden:
            if sigma_h is not None:          # GPRBM feeds a per-datapoint sigma_h [N,H] (gprbm.py:114-115)
                self.opt_sigma_h = None
                self.sigma_h = np.asarray(sigma_h, dtype=dtype)
            else:
                if opt_sigma_h is None:
                    opt_sigma_h = inverse_softplus(np.ones((1, self.H)))
                self.opt_sigma_h = np.asarray(opt_sigma_h, dtype=dtype).reshape(1, self.H)
                self.sigma_h = softplus(self.opt_sigma_h)

    # ---- net inputs / probabilities ---------------------------------------------------------------
    def h_net_input(self, v):
        v = np.asarray(v, dtype=self.dtype)
        if self.gaussian_hidden:                       # bgrbm.py:79-82
            return self.sigma_h * (v @ self.W) + self.c
        return v @ self.W + self.c                     # rbm.py:129-131

    def v_net_input(self, h):
        h = np.asarray(h, dtype=self.dtype)
        if self.gaussian_hidden:                       # bgrbm.py:126-131
            h = h / self.sigma_h
        return (self.W @ h.T + self.b).T               # rbm.py:143-147

    def h_probs(self, v):
        x = self.h_net_input(v)
        return x if self.gaussian_hidden else sigmoid(x)        # bgrbm.py:94 / rbm.py:160

    def v_probs(self, h):
        x = self.v_net_input(h)
        return x if self.gaussian_visible else sigmoid(x)       # gbrbm.py:56-57 / rbm.py:174

    # ---- sampling -------------------------------------------------------------------------------------
    def _bernoulli(self, p, u, sample=True):
        if not sample:                                 # rbm.py:209-210
            return p
        return (p > np.asarray(u, dtype=self.dtype)).astype(self.dtype)    # strict '>' rbm.py:206-208

    def _concrete(self, p, u):                         # rbm.py:243-258
        u = np.asarray(u, dtype=self.dtype)
        e0 = self.dtype(EPS0)
        pre = np.log(p + e0) - np.log(1.0 - p + e0) + np.log(u + e0) - np.log(1.0 - u + e0)
        return sigmoid(self.dtype(1.0 / self.temperature) * pre)

    def sample_h(self, h_probs, draw=None, sample=True):
        """draw: uniforms (Bernoulli / Concrete) or standard normals (Gaussian hiddens)."""
        if self.gaussian_hidden:                       # bgrbm.py:96-114
            if not sample:
                return h_probs
            return h_probs + self.sigma_h * np.asarray(draw, dtype=self.dtype)
        if self.temperature is not None:               # rbm.py:177-180
            return self._concrete(h_probs, draw)
        return self._bernoulli(h_probs, draw, sample)

    def sample_v(self, v_probs, draw=None, sample=True):
        if self.gaussian_visible:                      # gbrbm.py:59-64: the "sample" is the mean
            return v_probs
        if self.temperature is not None:               # rbm.py:182-185
            return self._concrete(v_probs, draw)
        return self._bernoulli(v_probs, draw, sample)

    def gibbs_sample(self, h, k, draws):
        """rbm.py:351-376. draws: list of k (draw_v, draw_h) pairs; always samples (quirk a'.1)."""
        assert k > 0
        for s in range(k):
            dv, dh = draws[s]
            v_probs = self.v_probs(h)
            v = self.sample_v(v_probs, dv)
            h_probs = self.h_probs(v)
            h = self.sample_h(h_probs, dh)
        return v_probs, v, h_probs, h

    # ---- energies / CD ------------------------------------------------------------------------------------
    def free_energy(self, v):
        v = np.asarray(v, dtype=self.dtype)
        if self.gaussian_hidden:                       # bgrbm.py:133-153
            x = v @ self.W
            t = 0.5 * np.square(x) + (self.c / self.sigma_h) * x
            return -(v @ self.b + np.sum(t, axis=1, keepdims=True))
        x = v @ self.W + self.c                        # rbm.py:388-396
        return -(v @ self.b + np.sum(softplus(x), axis=1, keepdims=True))

    def cd_loss(self, v0, vk):                         # rbm.py:398-415
        return np.mean(self.free_energy(v0)) - np.mean(self.free_energy(vk))

    def cd_grad(self, v0, vk):
        """Closed-form gradient of cd_loss wrt (W, b, c[, opt_sigma_h]) — SURVEY.md §8 R9 / B2.
        Cross-checked against central differences of cd_loss in tests/test_oracle.py."""
        v0 = np.asarray(v0, dtype=self.dtype)
        vk = np.asarray(vk, dtype=self.dtype)
        n = self.dtype(v0.shape[0])
        if self.gaussian_hidden:
            x0, xk = v0 @ self.W, vk @ self.W
            q0, qk = x0 + self.c / self.sigma_h, xk + self.c / self.sigma_h
            d = (np.sum(x0, axis=0, keepdims=True) - np.sum(xk, axis=0, keepdims=True)) / n
            gc = -d / self.sigma_h
            gsig = (self.c / np.square(self.sigma_h)) * d
            gopt = gsig * sigmoid(self.opt_sigma_h) if self.opt_sigma_h is not None else gsig
        else:
            q0, qk = sigmoid(v0 @ self.W + self.c), sigmoid(vk @ self.W + self.c)
            gc = -(np.sum(q0, axis=0, keepdims=True) - np.sum(qk, axis=0, keepdims=True)) / n
            gopt = None
        gW = -(v0.T @ q0 - vk.T @ qk) / n
        gb = -((np.sum(v0, axis=0) - np.sum(vk, axis=0)) / n).reshape(self.V, 1)
        return gW, gb, gc, gopt

    def pack(self):
        parts = [self.W.reshape(-1), self.b.reshape(-1), self.c.reshape(-1)]
        if self.gaussian_hidden and self.opt_sigma_h is not None:
            parts.append(self.opt_sigma_h.reshape(-1))
        return np.concatenate(parts)

    def cd_grad_packed(self, v0, vk):
        gW, gb, gc, gopt = self.cd_grad(v0, vk)
        parts = [gW.reshape(-1), gb.reshape(-1), gc.reshape(-1)]
        if gopt is not None:
            parts.append(gopt.reshape(-1))
        return np.concatenate(parts)


def cd_k(model, v0, k, draw_h0, chain_draws, persistent_h=None):
    """One CD-k / PCD-k step of RBM.train (rbm.py:628-646, 664): returns dict with the chain ends and
    the packed gradient. persistent_h (PCD, rbm.py:631-643) replaces h0 as the chain start."""
    h0_probs = model.h_probs(v0)
    h0 = model.sample_h(h0_probs, draw_h0)
    start = h0 if persistent_h is None else np.asarray(persistent_h, dtype=model.dtype)
    vk_probs, vk, hk_probs, hk = model.gibbs_sample(start, k, chain_draws)
    return {'h0_probs': h0_probs, 'h0': h0, 'vk_probs': vk_probs, 'vk': vk, 'hk_probs': hk_probs, 'hk': hk,
            'grad': model.cd_grad_packed(v0, vk), 'loss': model.cd_loss(v0, vk)}


# ---- optimizers [TF, not vendored]: documented update rules ------------------------------
class SGD:
    def __init__(self, lr):
        self.lr = lr

    def step(self, theta, g):
        return theta - self.lr * g


class Momentum:
    """tf.train.MomentumOptimizer: a <- mu a + g; theta <- theta - lr a. weight_decay adds wd*theta to g."""

    def __init__(self, lr, momentum=0.9, weight_decay=0.0):
        self.lr, self.mu, self.wd, self.a = lr, momentum, weight_decay, None

    def step(self, theta, g):
        g = g + self.wd * theta
        self.a = g if self.a is None else self.mu * self.a + g
        return theta - self.lr * self.a


class Adam:
    """tf.train.AdamOptimizer: lr_t = lr sqrt(1-b2^t)/(1-b1^t); eps outside the root."""

    def __init__(self, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
        self.lr, self.b1, self.b2, self.eps, self.wd = lr, beta1, beta2, eps, weight_decay
        self.m = self.v = None
        self.t = 0

    def step(self, theta, g):
        g = g + self.wd * theta
        if self.m is None:
            self.m, self.v = np.zeros_like(theta), np.zeros_like(theta)
        self.t += 1
        self.m = self.b1 * self.m + (1 - self.b1) * g
        self.v = self.b2 * self.v + (1 - self.b2) * g * g
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        return theta - lr_t * self.m / (np.sqrt(self.v) + self.eps)


# ---- fine-tuning through propagate(): RBM.backprop (rbm.py:713-789) -------------------------------------------
def _concrete(p, u, temperature, dtype):
    e0 = dtype(EPS0)
    pre = np.log(p + e0) - np.log(1.0 - p + e0) + np.log(u + e0) - np.log(1.0 - u + e0)
    return sigmoid(dtype(1.0 / temperature) * pre)


def propagate_loss_and_grads(layers, v, k, draws, sample=True, dtype=np.float64):
    """loss = dist.cross_entropy(propagate(v, k, sample), v) (rbm.py:739-748, dist.py:24-35 reduce_mean=True) for a
    stack of binary RBMs given bottom-first as dicts {W, b, c, temperature}, and its gradient wrt every W, b, c
    (what optimizer.minimize(training_loss) differentiates, rbm.py:755-758).
    draws: one uniform array per sampling op in execution order (per cycle: upprop bottom->top, downprop top->bottom).
    Units with a temperature are Concrete (rbm.py:177-185: chosen whenever temperature is set, whatever `sample`
    says); without one they are Bernoulli samples (no gradient) or, with sample=False, probabilities."""
    L = [{key: (np.asarray(val, dtype=dtype) if key in ('W', 'b', 'c') else val) for key, val in lay.items()} for lay in layers]
    v = np.asarray(v, dtype=dtype)
    tape, di, a = [], 0, v
    for _ in range(k):
        for li, lay in enumerate(L):                              # upprop (rbm.py:277-300)
            p = sigmoid(a @ lay['W'] + lay['c'].reshape(1, -1))
            T = lay.get('temperature')
            if T is not None:
                out, mode = _concrete(p, np.asarray(draws[di], dtype=dtype), T, dtype), 'concrete'; di += 1
            elif sample:
                out, mode = (p > np.asarray(draws[di], dtype=dtype)).astype(dtype), 'bernoulli'; di += 1
            else:
                out, mode = p, 'prob'
            tape.append(('up', li, a, p, out, mode))
            a = out
        for li in range(len(L) - 1, -1, -1):                      # downprop (rbm.py:302-325)
            lay = L[li]
            p = sigmoid(a @ lay['W'].T + lay['b'].reshape(1, -1))
            T = lay.get('temperature')
            if T is not None:
                out, mode = _concrete(p, np.asarray(draws[di], dtype=dtype), T, dtype), 'concrete'; di += 1
            elif sample:
                out, mode = (p > np.asarray(draws[di], dtype=dtype)).astype(dtype), 'bernoulli'; di += 1
            else:
                out, mode = p, 'prob'
            tape.append(('down', li, a, p, out, mode))
            a = out
    e0 = dtype(EPS0)
    pc = np.clip(a, e0, 1 - e0)
    logits = np.log(pc / (1 - pc))
    ce = np.maximum(logits, 0) - logits * v + np.log1p(np.exp(-np.abs(logits)))
    N = v.shape[0]
    loss = np.mean(np.sum(ce, axis=1))
    grads = [dict(W=np.zeros_like(l['W']), b=np.zeros_like(l['b']), c=np.zeros_like(l['c'])) for l in L]
    g = (-v / pc + (1 - v) / (1 - pc)) * ((a >= e0) & (a <= 1 - e0)) / N
    for direction, li, a_in, p, out, mode in reversed(tape):
        lay = L[li]
        if mode == 'bernoulli':
            break                                                   # tf.cast(tf.greater(...)) carries no gradient
        if mode == 'concrete':
            T = dtype(lay['temperature'])
            dz = g * out * (1 - out) / T * (1 / (p + e0) + 1 / (1 - p + e0)) * p * (1 - p)
        else:
            dz = g * p * (1 - p)
        if direction == 'up':
            grads[li]['W'] += a_in.T @ dz
            grads[li]['c'] += dz.sum(axis=0).reshape(grads[li]['c'].shape)
            g = dz @ lay['W'].T
        else:
            grads[li]['W'] += dz.T @ a_in
            grads[li]['b'] += dz.sum(axis=0).reshape(grads[li]['b'].shape)
            g = dz @ lay['W']
    return float(loss), grads, a
