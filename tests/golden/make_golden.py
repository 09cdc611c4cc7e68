"""Generate the golden vectors in tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules are imported unmodified from /root/reference on top of tf_numpy_shim (a numpy
stand-in for the `tensorflow` module, see its docstring), in both fp32 (deepbelief/config.py default)
and fp64 (config.float_type patched). All random draws are injected. Gradients, which the reference
obtains from TF autodiff, are recorded as central finite differences (fp64) of the reference's own
loss functions.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim as shim  # noqa: E402

tf = shim.install()
REFERENCE_ROOT = '/root/reference'
sys.path.insert(0, REFERENCE_ROOT)
import deepbelief.config as cfg  # noqa: E402
from deepbelief import dist, preprocessing  # noqa: E402
from deepbelief.layers import rbm as ref_rbm, bgrbm as ref_bgrbm, gbrbm as ref_gbrbm, ggrbm as ref_ggrbm  # noqa: E402
from deepbelief.layers import gplvm as ref_gplvm  # noqa: E402
from deepbelief.layers import gprbm as ref_gprbm  # noqa: E402
from deepbelief.layers import sbm_lower as ref_sbm  # noqa: E402

sys.path.insert(0, os.path.join(HERE, '..', '..'))
from deepbelief_b200 import hdf5  # noqa: E402  (only the HDF5 reader; h5py is unavailable)

KINDS = {'rbm': ref_rbm.RBM, 'bgrbm': ref_bgrbm.BGRBM, 'gbrbm': ref_gbrbm.GBRBM, 'ggrbm': ref_ggrbm.GGRBM}


def set_dtype(name):
    cfg.float_type = name


def make_layer(kind, V, H, W, b, c, opt_sigma, temperature=None, bottom=None, name='L'):
    sess = tf.Session()
    kw = dict(num_v=V, num_h=H, session=sess, name=name, bottom=bottom)
    if kind != 'ggrbm':
        kw['temperature'] = temperature
    layer = KINDS[kind](**kw)
    dt = np.dtype(cfg.float_type)
    layer.W = W.astype(dt)
    layer.b = b.astype(dt).reshape(V, 1)
    layer.c = c.astype(dt).reshape(1, H)
    if kind in ('bgrbm', 'ggrbm'):
        layer.opt_sigma_h = opt_sigma.astype(dt).reshape(1, H)
        layer.sigma_h = tf.nn.softplus(layer.opt_sigma_h)      # bgrbm.py:51
    return layer


def rbm_goldens(out):
    rs = np.random.RandomState(1234)
    for tag, (V, H, N) in {'toy': (6, 5, 20), 'mid': (40, 24, 16)}.items():
        W = rs.normal(size=(V, H)) * 0.3
        b = rs.normal(size=V) * 0.2
        c = rs.normal(size=H) * 0.2
        opt_sigma = rs.normal(size=H) * 0.3 + 0.5
        v_bin = rs.binomial(1, 0.5, size=(N, V)).astype(np.float64)
        v_real = rs.normal(size=(N, V))
        h_bin = rs.binomial(1, 0.5, size=(N, H)).astype(np.float64)
        h_real = rs.normal(size=(N, H))
        vk_bin = rs.binomial(1, 0.5, size=(N, V)).astype(np.float64)
        k = 3
        U_h = rs.uniform(size=(N, H))
        U_v = rs.uniform(size=(N, V))
        # bgrbm.py:108 draws float32 normals whatever config.float_type is: keep them fp32-representable
        E_h = rs.normal(size=(N, H)).astype(np.float32).astype(np.float64)
        chain_v = rs.uniform(size=(k, N, V))
        chain_h_u = rs.uniform(size=(k, N, H))
        chain_h_e = rs.normal(size=(k, N, H)).astype(np.float32).astype(np.float64)
        base = {'W': W, 'b': b, 'c': c, 'opt_sigma': opt_sigma, 'v_bin': v_bin, 'v_real': v_real, 'h_bin': h_bin,
                'h_real': h_real, 'vk_bin': vk_bin, 'U_h': U_h, 'U_v': U_v, 'E_h': E_h, 'chain_v': chain_v,
                'chain_h_u': chain_h_u, 'chain_h_e': chain_h_e, 'k': np.array(k)}
        for kk, vv in base.items():
            out['%s/%s' % (tag, kk)] = vv
        for kind in KINDS:
            gv = kind in ('gbrbm', 'ggrbm')
            gh = kind in ('bgrbm', 'ggrbm')
            for dname in ('float32', 'float64'):
                set_dtype(dname)
                dt = np.dtype(dname)
                layer = make_layer(kind, V, H, W, b, c, opt_sigma)
                v = (v_real if gv else v_bin).astype(dt)
                h = (h_real if gh else h_bin).astype(dt)
                vk = (0.7 * v_real[::-1] + 0.05 if gv else vk_bin).astype(dt)
                pre = '%s/%s/%s/' % (tag, kind, dname)
                hp = layer.h_probs(v)
                vp = layer.v_probs(h)
                out[pre + 'h_net_input'] = layer.h_net_input(v)
                out[pre + 'v_net_input'] = layer.v_net_input(h)
                out[pre + 'h_probs'] = hp
                out[pre + 'v_probs'] = vp
                shim.RANDOM_QUEUE[:] = [E_h if gh else U_h]
                out[pre + 'sample_h'] = layer.sample_h(hp)
                shim.RANDOM_QUEUE[:] = [] if gv else [U_v]
                out[pre + 'sample_v'] = layer.sample_v(vp)
                out[pre + 'free_energy'] = layer.free_energy(v)
                out[pre + 'cd_loss'] = np.asarray(layer.cd_loss(v, vk))
                # k-step chain (rbm.py:351-376): draws consumed in the order v, h, v, h, ...
                q = []
                for s in range(k):
                    if not gv:
                        q.append(chain_v[s])
                    q.append(chain_h_e[s] if gh else chain_h_u[s])
                shim.RANDOM_QUEUE[:] = q
                gvp, gvs, ghp, ghs = layer.gibbs_sample(h, k=k)
                assert not shim.RANDOM_QUEUE
                out[pre + 'gibbs_v_probs'], out[pre + 'gibbs_v'] = gvp, gvs
                out[pre + 'gibbs_h_probs'], out[pre + 'gibbs_h'] = ghp, ghs
                # Concrete units (rbm.py:243-275); GGRBM takes no temperature
                if kind != 'ggrbm':
                    lc = make_layer(kind, V, H, W, b, c, opt_sigma, temperature=0.1)
                    if not gh:
                        shim.RANDOM_QUEUE[:] = [U_h]
                        out[pre + 'concrete_h'] = lc.sample_h(hp)
                    if not gv:
                        shim.RANDOM_QUEUE[:] = [U_v]
                        out[pre + 'concrete_v'] = lc.sample_v(vp)
                shim.RANDOM_QUEUE[:] = []
            # finite-difference gradient of the reference's cd_loss (fp64), toy shape only
            if tag == 'toy':
                set_dtype('float64')
                v = (v_real if gv else v_bin)
                vk = (0.7 * v_real[::-1] + 0.05 if gv else vk_bin)
                theta = [W.copy(), b.copy(), c.copy()] + ([opt_sigma.copy()] if gh else [])

                def f(parts):
                    lay = make_layer(kind, V, H, parts[0], parts[1], parts[2], parts[3] if gh else opt_sigma)
                    return float(lay.cd_loss(v, vk))
                grads = []
                eps = 1e-6
                for pi, par in enumerate(theta):
                    g = np.zeros_like(par)
                    it = np.nditer(par, flags=['multi_index'])
                    for _ in it:
                        idx = it.multi_index
                        old = par[idx]
                        par[idx] = old + eps
                        fp = f(theta)
                        par[idx] = old - eps
                        fm = f(theta)
                        par[idx] = old
                        g[idx] = (fp - fm) / (2 * eps)
                    grads.append(g)
                out['%s/%s/fd_grad_W' % (tag, kind)] = grads[0]
                out['%s/%s/fd_grad_b' % (tag, kind)] = grads[1]
                out['%s/%s/fd_grad_c' % (tag, kind)] = grads[2]
                if gh:
                    out['%s/%s/fd_grad_opt_sigma' % (tag, kind)] = grads[3]
    # two-layer stack: upprop / downprop recursion (rbm.py:277-325)
    set_dtype('float64')
    rs = np.random.RandomState(77)
    V, H1, H2, N = 12, 8, 5, 9
    W1, b1, c1 = rs.normal(size=(V, H1)) * 0.4, rs.normal(size=V) * 0.1, rs.normal(size=H1) * 0.1
    W2, b2, c2 = rs.normal(size=(H1, H2)) * 0.4, rs.normal(size=H1) * 0.1, rs.normal(size=H2) * 0.1
    l1 = make_layer('rbm', V, H1, W1, b1, c1, None, name='l1')
    l2 = make_layer('rbm', H1, H2, W2, b2, c2, None, bottom=l1, name='l2')
    v = rs.binomial(1, 0.5, size=(N, V)).astype(np.float64)
    u1, u2, u3, u4 = rs.uniform(size=(N, H1)), rs.uniform(size=(N, H2)), rs.uniform(size=(N, H1)), rs.uniform(size=(N, V))
    shim.RANDOM_QUEUE[:] = [u1, u2]
    up = l2.upprop(v)
    shim.RANDOM_QUEUE[:] = [u3, u4]
    down = l2.downprop(up)
    for kk, vv in dict(W1=W1, b1=b1, c1=c1, W2=W2, b2=b2, c2=c2, v=v, u1=u1, u2=u2, u3=u3, u4=u4, up=up, down=down).items():
        out['stack/' + kk] = vv


def backprop_goldens(out):
    """RBM.backprop's training loss (rbm.py:739-748): cross entropy of propagate(v, k) against v on a two-layer stack,
    by the reference's own rbm.py / dist.py; its gradient as directional finite differences (fp64)."""
    set_dtype('float64')
    rs = np.random.RandomState(314)
    V, H1, H2, N, k = 14, 9, 6, 11, 2
    P = {'W1': rs.normal(size=(V, H1)) * 0.4, 'b1': rs.normal(size=V) * 0.1, 'c1': rs.normal(size=H1) * 0.1,
         'W2': rs.normal(size=(H1, H2)) * 0.4, 'b2': rs.normal(size=H1) * 0.1, 'c2': rs.normal(size=H2) * 0.1}
    v = rs.binomial(1, 0.5, size=(N, V)).astype(np.float64)
    draws = []
    for _ in range(k):
        draws += [rs.uniform(size=(N, H1)), rs.uniform(size=(N, H2)), rs.uniform(size=(N, H1)), rs.uniform(size=(N, V))]
    for kk, vv in P.items():
        out['backprop/P/' + kk] = vv
    out['backprop/v'], out['backprop/k'] = v, np.int64(k)
    for i, d in enumerate(draws):
        out['backprop/draw%d' % i] = d

    def loss_at(params, T, sample):
        l1 = make_layer('rbm', V, H1, params['W1'], params['b1'], params['c1'], None, temperature=T, name='b1')
        l2 = make_layer('rbm', H1, H2, params['W2'], params['b2'], params['c2'], None, temperature=T, bottom=l1, name='b2')
        shim.RANDOM_QUEUE[:] = [d for d in draws] if (T is not None or sample) else []
        v1 = l2.propagate(v, k=k if (T is not None or sample) else 1, sample=sample)
        shim.RANDOM_QUEUE[:] = []
        return float(dist.cross_entropy(v1, v)), v1
    for cname, (T, sample) in {'concrete': (0.25, True), 'prob': (None, False)}.items():
        loss, v1 = loss_at(P, T, sample)
        out['backprop/%s/loss' % cname], out['backprop/%s/v1' % cname] = np.float64(loss), v1
        drs = np.random.RandomState(8)
        eps = 1e-5
        for key in P:
            d = drs.normal(size=P[key].shape); d /= np.linalg.norm(d)
            Pp, Pm = dict(P), dict(P)
            Pp[key] = P[key] + eps * d; Pm[key] = P[key] - eps * d
            out['backprop/%s/dir/%s' % (cname, key)] = d
            out['backprop/%s/fd/%s' % (cname, key)] = np.float64((loss_at(Pp, T, sample)[0] - loss_at(Pm, T, sample)[0]) / (2 * eps))


def gp_goldens(out):
    rs = np.random.RandomState(4321)
    # SEKernel (tests/test_gplvm.py shapes + a 2-D case)
    for tag, (n, m, q) in {'k83': (8, 8, 3), 'k30': (30, 11, 2)}.items():
        X, Y = rs.uniform(size=(n, q)), rs.uniform(size=(m, q))
        out['%s/X' % tag], out['%s/Y' % tag] = X, Y
        for dname in ('float32', 'float64'):
            set_dtype(dname)
            for an, (alpha, gamma) in {'unit': (1.0, 1.0), 'ag': (1.7, 0.6)}.items():
                kern = ref_gplvm.SEKernel(alpha=alpha, gamma=gamma, session=tf.Session())
                dt = np.dtype(dname)
                pre = '%s/%s/%s/' % (tag, dname, an)
                out[pre + 'K_X'] = kern.K(X.astype(dt))
                out[pre + 'K_XY'] = kern.K(X.astype(dt), Y.astype(dt))
                out[pre + 'K_X_diag'] = kern.K(X.astype(dt), diag=True)
                if n == m:
                    out[pre + 'K_XY_diag'] = kern.K(X.astype(dt), Y.astype(dt), diag=True)
    # GPLVM: loss, K, L, prediction; finite-difference gradient
    cases = {'trainable': dict(noise_variance=0.7, fixed_noise_variance=False, alpha=1.3, gamma=0.8),
             'gpdbn': dict(noise_variance=0.01, fixed_noise_variance=True, alpha=False, gamma=False)}
    N, D, Q, M = 30, 12, 2, 7
    Yd = rs.normal(size=(N, D))
    xs = rs.normal(size=(M, Q))
    out['gplvm/Y'], out['gplvm/xstar'] = Yd, xs
    for cname, kw in cases.items():
        for dname in ('float32', 'float64'):
            set_dtype(dname)
            dt = np.dtype(dname)
            sess = tf.Session()
            kern = ref_gplvm.SEKernel(alpha=kw['alpha'], gamma=kw['gamma'], session=sess)
            g = ref_gplvm.GPLVM(Yd.astype(dt), Q, kern=kern, noise_variance=kw['noise_variance'],
                                fixed_noise_variance=kw['fixed_noise_variance'], x_test_var=xs.astype(dt), session=sess)
            g.build_model()
            pre = 'gplvm/%s/%s/' % (cname, dname)
            out[pre + 'X0'] = np.asarray(g.X)
            out[pre + 'Y_centered'] = np.asarray(g.Y)
            out[pre + 'Y_mean'] = np.asarray(g.Y_mean)
            out[pre + 'K'], out[pre + 'L'] = g.K, g.L
            out[pre + 'loss'], out[pre + 'log_det'] = np.asarray(g.loss), np.asarray(g.log_det)
            out[pre + 'pred_mean'], out[pre + 'pred_var'] = g.pred_mean, g.pred_var
        # fd gradient in fp64 wrt X and the optimisation variables
        set_dtype('float64')
        sess = tf.Session()
        X0 = out['gplvm/%s/float64/X0' % cname].copy()
        isp = lambda v: np.log(np.exp(v) - 1.0)  # noqa: E731

        def loss_at(X, oa, og, on):
            kern = ref_gplvm.SEKernel(alpha=kw['alpha'], gamma=kw['gamma'], session=sess)
            g = ref_gplvm.GPLVM(Yd, Q, X0=X, kern=kern, noise_variance=kw['noise_variance'],
                                fixed_noise_variance=kw['fixed_noise_variance'], x_test_var=xs, session=sess)
            if kw['alpha'] is not False:
                kern.opt_alpha = np.float64(oa); kern.alpha = tf.nn.softplus(kern.opt_alpha)
            if kw['gamma'] is not False:
                kern.opt_gamma = np.float64(og); kern.gamma = tf.nn.softplus(kern.opt_gamma)
            if not kw['fixed_noise_variance']:
                g.opt_noise_variance = np.float64(on); g.noise_variance = tf.nn.softplus(g.opt_noise_variance)
            g.K = g.build_K(); g.L = tf.cholesky(g.K); g.log_det = g.build_log_det()
            return float(g.build_loss())
        oa = isp(kw['alpha']) if kw['alpha'] is not False else 0.0
        og = isp(kw['gamma']) if kw['gamma'] is not False else 0.0
        on = isp(kw['noise_variance'])
        eps = 1e-6
        gX = np.zeros_like(X0)
        for i in range(N):
            for qd in range(Q):
                Xp, Xm = X0.copy(), X0.copy()
                Xp[i, qd] += eps; Xm[i, qd] -= eps
                gX[i, qd] = (loss_at(Xp, oa, og, on) - loss_at(Xm, oa, og, on)) / (2 * eps)
        out['gplvm/%s/fd_dX' % cname] = gX
        out['gplvm/%s/fd_dopt' % cname] = np.array([
            (loss_at(X0, oa + eps, og, on) - loss_at(X0, oa - eps, og, on)) / (2 * eps),
            (loss_at(X0, oa, og + eps, on) - loss_at(X0, oa, og - eps, on)) / (2 * eps),
            (loss_at(X0, oa, og, on + eps) - loss_at(X0, oa, og, on - eps)) / (2 * eps)])
        out['gplvm/%s/opt' % cname] = np.array([oa, og, on])
    # test-time latent search loss (gplvm.py:478-483) at single test latents, by the reference's own graph pieces
    set_dtype('float64')
    tp = rs.binomial(1, 0.5, size=(1, D)).astype(np.float64)
    out['testgen/target'] = tp
    kw = cases['trainable']
    losses, norms, clips = [], [], []
    for m in range(M):
        sess = tf.Session()
        kern = ref_gplvm.SEKernel(alpha=kw['alpha'], gamma=kw['gamma'], session=sess)
        g = ref_gplvm.GPLVM(Yd, Q, kern=kern, noise_variance=kw['noise_variance'],
                            fixed_noise_variance=kw['fixed_noise_variance'], x_test_var=xs[m:m + 1], session=sess)
        g.build_model()
        pmn = g.build_pred_mean_normalized(g.pred_mean)
        loss = 1 / D * dist.cross_entropy(pmn, tp) + 0.1 * tf.log(g.pred_var)
        losses.append(np.asarray(loss).reshape(-1)[0]); norms.append(np.asarray(pmn)[0])
        clips.append(np.asarray(tf.clip_by_value(g.pred_mean, 0.0, 1.0))[0])
    out['testgen/loss'], out['testgen/pm_norm'], out['testgen/pm_clip'] = np.array(losses), np.array(norms), np.array(clips)
    # GPLVM at construction on the bundled horses data (anchor (3) of SURVEY.md §8c), reference code, fp64 + fp32
    horses = hdf5.File(os.path.join(HERE, 'datasets', 'weizmann_horses_training_328x1024_binary.h5'))['datapoints']
    for dname in ('float32', 'float64'):
        set_dtype(dname)
        sess = tf.Session()
        g = ref_gplvm.GPLVM(horses.astype(np.dtype(dname)), 2, noise_variance=1.0, session=sess,
                            x_test_var=np.zeros((1, 2), dtype=np.dtype(dname)))
        g.build_model()
        out['horses/%s/loss' % dname] = np.asarray(g.loss)
        out['horses/%s/log_det' % dname] = np.asarray(g.log_det)
        out['horses/%s/X0' % dname] = np.asarray(g.X)


def gpdbn_goldens(out):
    """GPDBN stack (gpdbn_horses/train.py:49-94 in miniature): two Concrete RBMs + the GPRBM top layer, built and
    evaluated by the reference's own gprbm.py / bgrbm.py / rbm.py / gplvm.py / dist.py. Parameters and every random
    draw are injected (initial W through tf.truncated_normal, top b/c through tf.zeros, see the shim); the
    gradient TF autodiff would give is pinned as directional finite differences (fp64) of that loss."""
    rs = np.random.RandomState(2024)
    N, V0, Ha, Hb, D, Q, T = 24, 30, 16, 12, 6, 2, 0.1
    Vd = rs.binomial(1, 0.4, size=(N, V0)).astype(np.float64)
    P = {'W0': rs.normal(size=(V0, Ha)) * 0.3, 'b0': rs.normal(size=V0) * 0.2, 'c0': rs.normal(size=Ha) * 0.2,
         'W1': rs.normal(size=(Ha, Hb)) * 0.3, 'b1': rs.normal(size=Ha) * 0.2, 'c1': rs.normal(size=Hb) * 0.2,
         'Wt': rs.normal(size=(Hb, D)) * 0.3, 'bt': rs.normal(size=Hb) * 0.2, 'ct': rs.normal(size=D) * 0.2,
         'X': rs.normal(size=(N, Q))}
    draws = {'up0': rs.uniform(size=(N, Ha)), 'up1': rs.uniform(size=(N, Hb)),
             'eps1': rs.normal(size=(N, D)).astype(np.float32).astype(np.float64),
             'eps2': rs.normal(size=(N, D)).astype(np.float32).astype(np.float64),
             'down_top': rs.uniform(size=(N, Hb)), 'down1': rs.uniform(size=(N, Ha)), 'down0': rs.uniform(size=(N, V0))}
    cases = {'fixed': dict(noise_variance=0.05, fixed_noise_variance=True, alpha=False, gamma=False),
             'trainable': dict(noise_variance=0.3, fixed_noise_variance=False, alpha=1.4, gamma=0.9)}
    for kk, vv in P.items():
        out['gpdbn/P/' + kk] = vv
    for kk, vv in draws.items():
        out['gpdbn/draws/' + kk] = vv
    out['gpdbn/V'] = Vd
    out['gpdbn/T'] = np.float64(T)

    def build(params, kw, dname, hyp=None):
        set_dtype(dname)
        dt = np.dtype(dname)
        sess = tf.Session()
        l0 = make_layer('rbm', V0, Ha, params['W0'], params['b0'], params['c0'], None, temperature=T, name='l0')
        l1 = make_layer('rbm', Ha, Hb, params['W1'], params['b1'], params['c1'], None, temperature=T, bottom=l0, name='l1')
        kern = ref_gplvm.SEKernel(alpha=kw['alpha'], gamma=kw['gamma'], session=sess)
        if hyp is not None:           # perturbed optimisation variables (finite differences)
            if kw['alpha'] is not False:
                kern.opt_alpha = np.asarray(hyp[0], dtype=dt); kern.alpha = tf.nn.softplus(kern.opt_alpha)
            if kw['gamma'] is not False:
                kern.opt_gamma = np.asarray(hyp[1], dtype=dt); kern.gamma = tf.nn.softplus(kern.opt_gamma)
        # the constructor itself evaluates K, L, pred_var and H1 = upprop(V) (gprbm.py:108-140): inject first
        shim.TRUNC_QUEUE[:] = [params['Wt']]
        shim.ZEROS_INJECT.clear()
        shim.ZEROS_INJECT[(Hb, 1)] = [params['bt']]
        shim.ZEROS_INJECT[(1, D)] = [params['ct']]
        shim.RANDOM_QUEUE[:] = [draws['up0'], draws['up1'], draws['eps1']]
        nv = kw['noise_variance']
        if hyp is not None and not kw['fixed_noise_variance']:
            nv = float(np.log1p(np.exp(hyp[2])))          # noise passed through its own inverse_softplus/softplus pair
        g = ref_gprbm.GPRBM(num_v=Hb, num_h=D, Q=Q, N=N, eval_flag=False, X0=params['X'].astype(dt), V=Vd.astype(dt),
                            kern=kern, noise_variance=nv, fixed_noise_variance=kw['fixed_noise_variance'],
                            temperature=T, bottom=l1, session=sess, name='top')
        if 'oah' in params:
            raise RuntimeError('opt_alpha_h is created inside the constructor; perturb it through alpha_h_hook')
        assert not shim.TRUNC_QUEUE and not shim.RANDOM_QUEUE
        shim.ZEROS_INJECT.clear()
        shim.RANDOM_QUEUE[:] = [draws['eps2'], draws['down_top'], draws['down1'], draws['down0']]
        g.build_model()
        assert not shim.RANDOM_QUEUE
        return g

    isp = lambda v: np.log(np.exp(v) - 1.0)  # noqa: E731
    for cname, kw in cases.items():
        for dname in ('float32', 'float64'):
            g = build(P, kw, dname)
            pre = 'gpdbn/%s/%s/' % (cname, dname)
            out[pre + 'loss'] = np.asarray(g.loss)
            out[pre + 'pred_var'] = np.asarray(g.pred_var)
            out[pre + 'sigma_h'] = np.asarray(g.sigma_h)
            out[pre + 'H1'] = np.asarray(g.H1)
            out[pre + 'H2'] = np.asarray(g.H2)
            out[pre + 'log_det'] = np.asarray(g.log_det)
            out[pre + 'gp_loss'] = np.asarray(g.build_gp_loss(summary=False))
        # directional finite differences of the reference loss (fp64): one random direction per parameter block
        hyp0 = np.array([isp(kw['alpha']) if kw['alpha'] is not False else 0.0,
                         isp(kw['gamma']) if kw['gamma'] is not False else 0.0, isp(kw['noise_variance'])])
        out['gpdbn/%s/hyp_opt' % cname] = hyp0
        eps = 1e-4     # the ~1e3-sized loss carries ~1e-10 evaluation noise: 1e-6 steps are noise-limited at 1e-4
        drs = np.random.RandomState(7)
        for key in ('W0', 'b0', 'c0', 'W1', 'b1', 'c1', 'Wt', 'bt', 'ct', 'X'):
            d = drs.normal(size=P[key].shape)
            d /= np.linalg.norm(d)
            Pp, Pm = dict(P), dict(P)
            Pp[key] = P[key] + eps * d
            Pm[key] = P[key] - eps * d
            fd = (float(build(Pp, kw, 'float64', hyp0).loss) - float(build(Pm, kw, 'float64', hyp0).loss)) / (2 * eps)
            out['gpdbn/%s/dir/%s' % (cname, key)] = d
            out['gpdbn/%s/fd/%s' % (cname, key)] = np.float64(fd)
        if cname == 'trainable':
            for j, key in enumerate(('opt_alpha', 'opt_gamma', 'opt_noise')):
                hp, hm = hyp0.copy(), hyp0.copy()
                hp[j] += eps; hm[j] -= eps
                fd = (float(build(P, kw, 'float64', hp).loss) - float(build(P, kw, 'float64', hm).loss)) / (2 * eps)
                out['gpdbn/%s/fd/%s' % (cname, key)] = np.float64(fd)


def latent_goldens(out):
    """GPRBM.test_held_out_data (gprbm.py:440-466) and the graph pieces of GPRBM.interp_test (gprbm.py:564-660) in
    EVAL mode (eval_flag True, x_ph = x_test), executed from the reference's own gprbm.py on the stack and parameters of
    gpdbn_golden.npz ('trainable' case). The optimizer loops themselves need TF autodiff and are not run; the objective
    values at given latents are what is pinned (their x-gradients are checked by finite differences in the tests)."""
    G = {k.replace('__', '/'): v for k, v in np.load(os.path.join(HERE, 'gpdbn_golden.npz')).items()}
    P = {k: G['gpdbn/P/' + k] for k in ('W0', 'b0', 'c0', 'W1', 'b1', 'c1', 'Wt', 'bt', 'ct', 'X')}
    Vd, T = G['gpdbn/V'], float(G['gpdbn/T'])
    N, V0 = Vd.shape
    Ha, Hb, D, Q = P['W0'].shape[1], P['W1'].shape[1], P['Wt'].shape[1], P['X'].shape[1]
    rs = np.random.RandomState(4242)
    H2 = rs.normal(size=(N, D))
    H1_mean = rs.normal(size=(1, D)) * 0.3
    out['latent/H2'], out['latent/H1_mean'] = H2, H1_mean

    def build(x_test, dname):
        set_dtype(dname)
        dt = np.dtype(dname)
        sess = tf.Session()
        l0 = make_layer('rbm', V0, Ha, P['W0'], P['b0'], P['c0'], None, temperature=T, name='l0')
        l1 = make_layer('rbm', Ha, Hb, P['W1'], P['b1'], P['c1'], None, temperature=T, bottom=l0, name='l1')
        kern = ref_gplvm.SEKernel(alpha=1.4, gamma=0.9, session=sess)
        shim.TRUNC_QUEUE[:] = [P['Wt']]
        shim.ZEROS_INJECT.clear()
        shim.ZEROS_INJECT[(Hb, 1)] = [P['bt']]
        shim.ZEROS_INJECT[(1, D)] = [P['ct']]
        shim.RANDOM_QUEUE[:] = []
        g = ref_gprbm.GPRBM(num_v=Hb, num_h=D, Q=Q, N=N, eval_flag=True, X0=P['X'].astype(dt), kern=kern,
                            noise_variance=0.3, fixed_noise_variance=False, temperature=T, bottom=l1,
                            x_test_var=np.asarray(x_test, dtype=dt), session=sess, name='top')
        assert not shim.TRUNC_QUEUE
        shim.ZEROS_INJECT.clear()
        g.H2_var = H2.astype(dt)                   # what restore() / update_H2_var leave in the snapshots
        g.H1_mean_var = H1_mean.astype(dt)
        return g

    # ---- test_held_out_data: mse(test_data, downprop(build_sample_mean(H2_var, H1_mean_var))) ----
    R, M = 3, 5
    xs = rs.normal(size=(R, Q))
    test_data = rs.binomial(1, 0.4, size=(M, V0)).astype(np.float64)
    hd = {'eps': rs.normal(size=(R, D)).astype(np.float32).astype(np.float64),
          'down_top': rs.uniform(0.05, 0.95, size=(R, Hb)), 'down1': rs.uniform(0.05, 0.95, size=(R, Ha)),
          'down0': rs.uniform(0.05, 0.95, size=(R, V0))}
    out['heldout/x'], out['heldout/test_data'] = xs, test_data
    for k, v in hd.items():
        out['heldout/draws/' + k] = v
    for dname in ('float32', 'float64'):
        vals = []
        for r in range(R):
            g = build(xs[r:r + 1], dname)
            shim.RANDOM_QUEUE[:] = [hd['eps'][r:r + 1], hd['down_top'][r:r + 1], hd['down1'][r:r + 1], hd['down0'][r:r + 1]]
            pm = g.build_sample_mean(g.H2_var, g.H1_mean_var)
            dm = g.downprop(pm)
            assert not shim.RANDOM_QUEUE
            vals.append(float(dist.mse(test_data.astype(dname), dm)))
        out['heldout/%s/mse' % dname] = np.array(vals)

    # ---- interp_test pieces: create_observed_prediction_with_variance at given interpolation latents ----
    Pn, S = 4, 3
    Xi = rs.normal(size=(Pn + 2, Q)) * 0.8
    idr = {'eps': rs.normal(size=((Pn + 2) * S, D)).astype(np.float32).astype(np.float64),
           'down_top': rs.uniform(0.05, 0.95, size=((Pn + 2) * S, Hb)),
           'down1': rs.uniform(0.05, 0.95, size=((Pn + 2) * S, Ha)),
           'down0': rs.uniform(0.05, 0.95, size=((Pn + 2) * S, V0))}
    out['interp/Xi'] = Xi
    out['interp/S'] = np.int64(S)
    for k, v in idr.items():
        out['interp/draws/' + k] = v
    for dname in ('float32', 'float64'):
        g = build(xs[:1], dname)
        q = []
        for i in range(Pn + 2):
            rows = slice(i * S, (i + 1) * S)
            # build_sample_mean_interp draws a [1,D] normal first and discards it when num_samples is given (gprbm.py:587-589)
            q += [np.zeros((1, D)), idr['eps'][rows], idr['down_top'][rows], idr['down1'][rows], idr['down0'][rows]]
        shim.RANDOM_QUEUE[:] = q
        imgs, var = g.create_observed_prediction_with_variance(tf_input=Xi.astype(dname), num_samples=S)
        assert not shim.RANDOM_QUEUE
        out['interp/%s/images' % dname] = np.asarray(imgs)
        out['interp/%s/pred_var' % dname] = np.asarray(var)


def ard_goldens(out):
    """SEKernel(ARD=True) (gplvm.py:24-28: one length-scale per latent dimension, gamma [1,Q]) and a GPLVM on it, run
    from the reference source; gradients as central differences (fp64) of the reference loss with respect to X and
    to each optimisation variable (opt_alpha, opt_gamma[q], opt_noise)."""
    rs = np.random.RandomState(777)
    N, D, Q, M = 26, 9, 3, 5
    Yd = rs.normal(size=(N, D))
    xs = rs.normal(size=(M, Q))
    X0 = rs.normal(size=(N, Q))
    gam = np.array([[0.6, 1.1, 1.9]])
    alpha, noise = 1.3, 0.4
    isp = lambda v: np.log(np.exp(v) - 1.0)  # noqa: E731
    out['ard/Y'], out['ard/xstar'], out['ard/X0'], out['ard/gamma'] = Yd, xs, X0, gam
    out['ard/alpha'], out['ard/noise'] = np.float64(alpha), np.float64(noise)

    def build(dname, X, oa, og, on):
        set_dtype(dname)
        dt = np.dtype(dname)
        sess = tf.Session()
        kern = ref_gplvm.SEKernel(alpha=alpha, gamma=1.0, ARD=True, Q=Q, session=sess)     # gamma -> ones [1,Q]
        assert np.asarray(kern.gamma).shape == (1, Q)
        kern.opt_alpha = np.asarray(oa, dtype=dt); kern.alpha = tf.nn.softplus(kern.opt_alpha)
        kern.opt_gamma = np.asarray(og, dtype=dt).reshape(1, Q); kern.gamma = tf.nn.softplus(kern.opt_gamma)
        g = ref_gplvm.GPLVM(Yd.astype(dt), Q, X0=X.astype(dt), kern=kern, noise_variance=noise, x_test_var=xs.astype(dt),
                            session=sess)
        g.opt_noise_variance = np.asarray(on, dtype=dt); g.noise_variance = tf.nn.softplus(g.opt_noise_variance)
        g.build_model()
        return g, kern

    oa, og, on = isp(alpha), isp(gam), isp(noise)
    out['ard/opt'] = np.concatenate([[oa], og.reshape(-1), [on]])
    for dname in ('float32', 'float64'):
        g, kern = build(dname, X0, oa, og, on)
        pre = 'ard/%s/' % dname
        dt = np.dtype(dname)
        out[pre + 'K_X'] = kern.K(X0.astype(dt))
        out[pre + 'K_Xx'] = kern.K(X0.astype(dt), xs.astype(dt))
        out[pre + 'K'], out[pre + 'L'] = g.K, g.L
        out[pre + 'loss'], out[pre + 'log_det'] = np.asarray(g.loss), np.asarray(g.log_det)
        out[pre + 'pred_mean'], out[pre + 'pred_var'] = g.pred_mean, g.pred_var
    eps = 1e-6
    f = lambda X, a, gm, n: float(build('float64', X, a, gm, n)[0].loss)  # noqa: E731
    gX = np.zeros_like(X0)
    for i in range(N):
        for qd in range(Q):
            Xp, Xm = X0.copy(), X0.copy()
            Xp[i, qd] += eps; Xm[i, qd] -= eps
            gX[i, qd] = (f(Xp, oa, og, on) - f(Xm, oa, og, on)) / (2 * eps)
    out['ard/fd_dX'] = gX
    dg = np.zeros(Q)
    for qd in range(Q):
        gp, gm = og.copy(), og.copy()
        gp[0, qd] += eps; gm[0, qd] -= eps
        dg[qd] = (f(X0, oa, gp, on) - f(X0, oa, gm, on)) / (2 * eps)
    out['ard/fd_dopt_gamma'] = dg
    out['ard/fd_dopt_alpha'] = np.float64((f(X0, oa + eps, og, on) - f(X0, oa - eps, og, on)) / (2 * eps))
    out['ard/fd_dopt_noise'] = np.float64((f(X0, oa, og, on + eps) - f(X0, oa, og, on - eps)) / (2 * eps))


def sbm_goldens(out):
    """SBM_Lower (sbm_lower.py:59-117): net inputs, probabilities, free energy, cd_loss of the reference class with injected
    shared weights, and the gradient of cd_loss with respect to the shared W and b as central differences (fp64)."""
    rs = np.random.RandomState(77)
    for tag, (side, ov, H, N) in {'s8': (8, 2, 12, 9), 's12': (12, 4, 8, 7)}.items():
        P = side // 2 + ov // 2
        W = rs.normal(size=(P * P, H // 4)) * 0.3
        b = rs.normal(size=side * side) * 0.2
        v0 = rs.binomial(1, 0.5, size=(N, side * side)).astype(np.float64)
        vk = rs.binomial(1, 0.5, size=(N, side * side)).astype(np.float64)
        h = rs.binomial(1, 0.5, size=(N, H)).astype(np.float64)
        pre = 'sbm/%s/' % tag
        out[pre + 'dims'] = np.array([side, ov, H, N])
        out[pre + 'W'], out[pre + 'b'], out[pre + 'v0'], out[pre + 'vk'], out[pre + 'h'] = W, b, v0, vk, h
        for dname in ('float64', 'float32'):
            set_dtype(dname)
            dt = np.dtype(dname)

            def build(Wm, bm, dt=dt):
                lay = ref_sbm.SBM_Lower(side=side, side_overlap=ov, num_h=H, session=tf.Session(), name='S' + tag + str(dt))
                lay.W = Wm.astype(dt)
                lay.b = bm.astype(dt).reshape(-1, 1)
                lay._b_transpose = np.transpose(lay.b)          # sbm_lower.py:57 (evaluated at construction)
                return lay
            lay = build(W, b)
            q = pre + dname + '/'
            out[q + 'h_net'] = lay.h_net_input(v0.astype(dt))
            out[q + 'v_net'] = lay.v_net_input(h.astype(dt))
            out[q + 'h_probs'] = lay.h_probs(v0.astype(dt))
            out[q + 'v_probs'] = lay.v_probs(h.astype(dt))
            out[q + 'free_energy'] = lay.free_energy(v0.astype(dt))
            out[q + 'cd_loss'] = lay.cd_loss(v0.astype(dt), vk.astype(dt))
        set_dtype('float64')

        def loss_at(Wm, bm):
            return float(build(Wm, bm, np.dtype('float64')).cd_loss(v0, vk))
        eps = 1e-6
        gW = np.zeros_like(W)
        for i in range(W.shape[0]):
            for j in range(W.shape[1]):
                d = np.zeros_like(W); d[i, j] = eps
                gW[i, j] = (loss_at(W + d, b) - loss_at(W - d, b)) / (2 * eps)
        gb = np.zeros_like(b)
        for i in range(b.size):
            d = np.zeros_like(b); d[i] = eps
            gb[i] = (loss_at(W, b + d) - loss_at(W, b - d)) / (2 * eps)
        out[pre + 'fd_gW'], out[pre + 'fd_gb'] = gW, gb
    set_dtype('float32')


def misc_goldens(out):
    rs = np.random.RandomState(99)
    set_dtype('float64')
    p = rs.uniform(size=(7, 9)); p[0, 0] = 0.0; p[1, 1] = 1.0
    t = rs.binomial(1, 0.5, size=(7, 9)).astype(np.float64)
    out['dist/p'], out['dist/t'] = p, t
    out['dist/mse'] = np.asarray(dist.mse(p, t))
    out['dist/ce_mean'] = np.asarray(dist.cross_entropy(p, t))
    out['dist/ce_sum'] = np.asarray(dist.cross_entropy(p, t, reduce_mean=False))
    A = rs.normal(size=(25, 6)) * np.array([3.0, 2.0, 1.0, 0.5, 0.2, 0.1])
    out['pre/A'] = A
    out['pre/pca2'] = preprocessing.PCA(A, 2)
    out['pre/std'] = preprocessing.standardize(A.copy())
    out['pre/inv_softplus'] = np.asarray(preprocessing.inverse_softplus(np.array([0.3, 1.0, 2.5])))


def reference_example_fixtures():
    """Byte copies of the reference's own example scripts (flags / train / test_generalisation of the three workflows
    SURVEY.md section 8b names), stored as `*.py.fixture`: tests/test_gpu_reference_scripts.py runs them UNMODIFIED
    through `python -m deepbelief_b200.run` (only iteration counts in the flags file are edited). Test inputs, like the
    bundled .h5 datasets — not product source."""
    import shutil
    for ex in ('dbn_horses', 'gplvm_horses', 'gpdbn_horses'):
        dst = os.path.join(HERE, 'reference_examples', ex)
        os.makedirs(dst, exist_ok=True)
        for f in ('flags', 'train', 'test_generalisation'):
            shutil.copyfile(os.path.join(REFERENCE_ROOT, 'examples', ex, f + '.py'), os.path.join(dst, f + '.py.fixture'))
        print('copied examples/%s/{flags,train,test_generalisation}.py' % ex)


def main():
    only = sys.argv[1:]
    if not only or 'reference_examples' in only:
        reference_example_fixtures()
    for fn, name in ((rbm_goldens, 'rbm_golden.npz'), (gp_goldens, 'gp_golden.npz'), (misc_goldens, 'misc_golden.npz'),
                     (gpdbn_goldens, 'gpdbn_golden.npz'), (backprop_goldens, 'backprop_golden.npz'), (sbm_goldens, 'sbm_golden.npz'),
                     (latent_goldens, 'latent_golden.npz'), (ard_goldens, 'ard_golden.npz')):
        if only and name not in only:
            continue
        out = {}
        fn(out)
        path = os.path.join(HERE, name)
        np.savez_compressed(path, **{k.replace('/', '__'): np.asarray(v) for k, v in out.items()})
        print('wrote %s: %d arrays, %d bytes' % (name, len(out), os.path.getsize(path)))


if __name__ == '__main__':
    main()
