"""Oracle restatement of the GP path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows /root/reference/deepbelief/layers/gplvm.py (SEKernel, GPLVM), gprbm.py (GP part of GPRBM) and
preprocessing.py (PCA, standardize, inverse_softplus) in numpy with a dtype switch.
"""
import numpy as np
import scipy.linalg

from .rbm import softplus, sigmoid, inverse_softplus  # noqa: F401


def sq_dist(x, y=None, gamma=1.0, diag=False):
    """gplvm.py:61-82 — expanded form, exactly as the reference evaluates it."""
    if diag and y is None:
        return np.zeros((x.shape[0],), dtype=x.dtype)            # gplvm.py:64-66
    keepdims = not diag
    xx = x / gamma
    sum_xx = np.sum(np.square(xx), 1, keepdims=keepdims)
    if y is None:
        yy, sum_yy = xx, sum_xx
    else:
        yy = y / gamma
        sum_yy = np.sum(np.square(yy), 1, keepdims=keepdims)
    if diag:
        prod = 2.0 * np.sum(xx * yy, 1)
    else:
        prod = 2.0 * (xx @ yy.T)
        sum_yy = sum_yy.T
    return -prod + sum_xx + sum_yy


def se_kernel(x, y=None, alpha=1.0, gamma=1.0, diag=False):
    """gplvm.py:84-86: K = alpha * exp(-0.5 * sqdist)."""
    return alpha * np.exp(-0.5 * sq_dist(x, y, gamma, diag))


def pca(X, num_components):
    """preprocessing.py:157-168 (eigh of the scatter matrix in the input dtype)."""
    x_mean = X.mean(0)
    X_zero = X - x_mean
    cov = X_zero.T.dot(X_zero)
    eigenvals, eigenvecs = np.linalg.eigh(cov)
    idx = np.argsort(eigenvals)[::-1]
    V = eigenvecs[:, idx][:, :num_components]
    return X_zero.dot(V)


def standardize(array, dtype=np.float32):
    """preprocessing.py:73-119: population std, std <= 1e-8 -> 1."""
    array = array.astype(dtype)
    means = array.mean(axis=0)
    stds = array.std(axis=0)
    stds[stds <= 1e-8] = 1.0
    return (array - means) / stds


class GPLVMOracle:
    """hyper-parameters are the OPTIMISATION variables (pre-softplus); `trainable` = (alpha, gamma, noise)
    flags; untrainable alpha/gamma are the constant 1 (gplvm.py:31-32,41-42), untrainable noise is the
    fixed value with NO jitter (gplvm.py:141-152, 219-221)."""

    def __init__(self, Y, X, opt_alpha=None, opt_gamma=None, opt_noise=None, fixed_noise=1.0, jitter=1e-6,
                 trainable=(True, True, True), dtype=np.float64):
        self.dtype = dtype
        self.Y = np.asarray(Y, dtype=dtype)
        self.X = np.asarray(X, dtype=dtype)
        self.N, self.D = self.Y.shape
        self.Q = self.X.shape[1]
        one = inverse_softplus(np.float64(1.0))
        self.opt_alpha = dtype(one if opt_alpha is None else opt_alpha)
        if opt_gamma is not None and np.ndim(opt_gamma) > 0:          # ARD: one length-scale per latent dimension (gplvm.py:24-28)
            self.opt_gamma = np.asarray(opt_gamma, dtype=dtype).reshape(1, self.Q)
        else:
            self.opt_gamma = dtype(one if opt_gamma is None else opt_gamma)
        self.opt_noise = dtype(one if opt_noise is None else opt_noise)
        self.fixed_noise = dtype(fixed_noise)
        self.jitter = dtype(jitter)
        self.trainable = trainable

    # effective hyper-parameters
    @property
    def alpha(self):
        return softplus(np.asarray(self.opt_alpha)) if self.trainable[0] else self.dtype(1.0)

    @property
    def gamma(self):
        return softplus(np.asarray(self.opt_gamma)) if self.trainable[1] else self.dtype(1.0)

    @property
    def noise(self):
        if self.trainable[2]:
            return softplus(np.asarray(self.opt_noise)) + self.jitter
        return self.fixed_noise

    def K(self):                                                    # gplvm.py:218-224
        Kf = se_kernel(self.X, None, self.alpha, self.gamma)
        return Kf + self.noise * np.eye(self.N, dtype=self.dtype)

    def chol(self):
        return np.linalg.cholesky(self.K())                         # gplvm.py:212

    def loss(self, x_prior_weight=0.0, include_const=True):
        """gplvm.py:233-244 (include_const) / gprbm.py:234-245 (x_prior_weight=1, no constant)."""
        L = self.chol()
        S = scipy.linalg.solve_triangular(L, self.Y, lower=True)
        logdet = 2.0 * np.sum(np.log(np.diag(L)))                   # gplvm.py:226-228
        ssq = np.sum(np.square(S))
        loss = 0.5 * (ssq + self.D * logdet)
        if include_const:
            loss = loss + 0.5 * self.D * self.N * np.log(2.0 * np.pi)
        loss = loss + x_prior_weight * np.sum(np.square(self.X))
        return self.dtype(loss), self.dtype(logdet), self.dtype(ssq), L, S

    def grad(self, x_prior_weight=0.0):
        """Closed forms of SURVEY.md §8 GP7 (what TF autodiff of gplvm.py:425-428 yields):
        returns dX, d/dopt_alpha, d/dopt_gamma, d/dopt_noise."""
        Kf = se_kernel(self.X, None, self.alpha, self.gamma)
        K = Kf + self.noise * np.eye(self.N, dtype=self.dtype)
        Kinv = np.linalg.inv(K)
        A = Kinv @ self.Y
        G = 0.5 * (self.D * Kinv - A @ A.T)
        Wm = G * Kf
        g2 = self.gamma * self.gamma                                # scalar, or [1,Q] with ARD (broadcasts over the columns of X)
        dX = -2.0 * (np.sum(Wm, axis=1, keepdims=True) * self.X - Wm @ self.X) / g2
        dX = dX + 2.0 * x_prior_weight * self.X
        d_alpha = np.sum(Wm) / self.alpha
        if np.ndim(self.opt_gamma) > 0:                             # ARD: d/dgamma_q = sum_ij Wm_ij (x_iq - x_jq)^2 / gamma_q^3
            d_gamma = np.array([np.sum(Wm * sq_dist(self.X[:, q:q + 1], None, 1.0)) for q in range(self.Q)]).reshape(1, self.Q)
            d_gamma = d_gamma / (self.gamma ** 3)
        else:
            diff2 = sq_dist(self.X, None, 1.0)                      # unscaled |xi-xj|^2
            d_gamma = np.sum(Wm * diff2) / (self.gamma ** 3)
        d_noise = np.trace(G)
        out = [dX.astype(self.dtype)]
        out.append(self.dtype(d_alpha * sigmoid(np.asarray(self.opt_alpha))) if self.trainable[0] else self.dtype(0))
        if np.ndim(self.opt_gamma) > 0:
            out.append((d_gamma * sigmoid(self.opt_gamma)).reshape(-1).astype(self.dtype))
        else:
            out.append(self.dtype(d_gamma * sigmoid(np.asarray(self.opt_gamma))) if self.trainable[1] else self.dtype(0))
        out.append(self.dtype(d_noise * sigmoid(np.asarray(self.opt_noise))) if self.trainable[2] else self.dtype(0))
        return tuple(out)

    def predict(self, xstar, y_mean=None):
        """gplvm.py:246-254 and 278-288: mean [M,D], var [M] (noise NOT added back)."""
        xstar = np.asarray(xstar, dtype=self.dtype)
        L = self.chol()
        KXx = se_kernel(self.X, xstar, self.alpha, self.gamma)
        sys1 = scipy.linalg.solve_triangular(L, KXx, lower=True)
        sys2 = scipy.linalg.solve_triangular(L, self.Y, lower=True)
        mean = sys1.T @ sys2
        if y_mean is not None:
            mean = mean + np.asarray(y_mean, dtype=self.dtype).reshape(1, -1)
        var = se_kernel(xstar, None, self.alpha, self.gamma, diag=True) - np.sum(np.square(sys1), axis=0)
        return mean, var


def testgen_loss(model, x, target, log_var_scaling=0.1, y_mean=None):
    """Per-restart loss of GPLVM.test_generalisation (gplvm.py:478-483) at test latents x [R,Q], each row treated as
    its own M=1 problem: CE(normalised(pred_mean), target) / Dim + log_var_scaling * log(pred_var)."""
    x = np.asarray(x, dtype=model.dtype)
    target = np.asarray(target, dtype=model.dtype).reshape(1, -1)
    out = np.zeros(x.shape[0], dtype=model.dtype)
    e0 = model.dtype(10e-7)
    for r in range(x.shape[0]):
        mean, var = model.predict(x[r:r + 1], y_mean)
        pn = (mean - mean.min()) / (mean.max() - mean.min())                     # gplvm.py:293-296
        pc = np.clip(pn, e0, 1 - e0)
        lg = np.log(pc / (1 - pc))
        ce = np.sum(np.maximum(lg, 0) - lg * target + np.log1p(np.exp(-np.abs(lg))))   # dist.py:24-35, one row
        out[r] = ce / target.shape[1] + log_var_scaling * np.log(var[0])
    return out
