"""First on-GPU sanity run of every kernel family against torch fp64 (checker only)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepbelief_b200 import ops, _lib as L

torch.manual_seed(0)
dev = 'cuda'

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300))

def check_cd_tc(N, V, H, k, pcd=False):
    W = (torch.randn(V, H, device=dev) * 0.05)
    b = torch.randn(V, device=dev) * 0.1
    c = torch.randn(H, device=dev) * 0.1
    params = torch.cat([W.reshape(-1), b, c]).contiguous()
    v0 = (torch.rand(N, V, device=dev) < 0.4).float()
    inj = torch.rand(N * H + k * (N * V + N * H), device=dev)
    plan = ops.CdTcPlan(N, V, H, k, pcd=pcd)
    plan.refresh_weights(params)
    grad = torch.zeros_like(params)
    vk = torch.empty(N, V, device=dev); hk = torch.empty(N, H, device=dev)
    p0 = torch.empty(N, H, device=dev); pk = torch.empty(N, H, device=dev)
    plan.step(v0, params, grad, injected=inj, vk_out=vk, hk_out=hk, p0_out=p0, pk_out=pk)
    torch.cuda.synchronize()
    # fp64 reference chain with the same uniforms
    Wd, bd, cd_ = W.double(), b.double(), c.double()
    p0r = torch.sigmoid(v0.double() @ Wd + cd_)
    h = (p0r > inj[:N * H].view(N, H).double()).double()
    off = N * H
    flips = 0
    for s in range(k):
        Uv = inj[off:off + N * V].view(N, V).double(); off += N * V
        Uh = inj[off:off + N * H].view(N, H).double(); off += N * H
        pv = torch.sigmoid(h @ Wd.t() + bd)
        v = (pv > Uv).double()
        ph = torch.sigmoid(v @ Wd + cd_)
        h = (ph > Uh).double()
    print(f'[cd_tc N={N} V={V} H={H} k={k}] p0 rel {rel(p0, p0r):.2e}  vk mismatches {int((vk.double() != v).sum())}/{N*V}'
          f'  hk mismatches {int((hk.double() != h).sum())}/{N*H}  pk rel {rel(pk, ph):.2e}')
    # gradient from the GPU's own chain ends (so sample flips do not pollute the check)
    vkd = vk.double(); pkd = torch.sigmoid(vkd @ Wd + cd_)
    gW = -(v0.double().t() @ p0r - vkd.t() @ pkd) / N
    gb = -(v0.double() - vkd).mean(0); gc = -(p0r - pkd).mean(0)
    g = grad.double()
    print(f'    dW rel {rel(g[:V*H].view(V, H), gW):.2e}  db rel {rel(g[V*H:V*H+V], gb):.2e}  dc rel {rel(g[V*H+V:], gc):.2e}')

def time_cd_tc(N, V, H, k, iters=3):
    W = (torch.randn(V, H, device=dev) * 0.01)
    params = torch.cat([W.reshape(-1), torch.zeros(V + H, device=dev)]).contiguous()
    v0 = (torch.rand(N, V, device=dev) < 0.5).float()
    plan = ops.CdTcPlan(N, V, H, k)
    plan.refresh_weights(params)
    grad = torch.zeros_like(params)
    rng = ops.Rng(seed=1)
    plan.step(v0, params, grad, rng=rng)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        plan.step(v0, params, grad, rng=rng)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = (2 * k + 3) * 2.0 * N * V * H
    print(f'[time cd_tc N={N} V={V} H={H} k={k}] {ms:.2f} ms/iter  {fl/ms/1e9:.1f} TFLOP/s algorithmic  {N*k/ms*1e3:.3e} sample-steps/s')

def check_simt():
    N, V, H = 200, 70, 45
    A = torch.randn(N, V, device=dev); W = torch.randn(V, H, device=dev) * 0.3; c = torch.randn(H, device=dev)
    U = torch.rand(N, H, device=dev)
    act, s = ops.layer_forward(A, W, c, False, L.ACT_SIGMOID, L.SAMPLE_BERNOULLI, injected=U)
    ref = torch.sigmoid(A.double() @ W.double() + c.double())
    print(f'[simt fwd] rel {rel(act, ref):.2e} sample mismatches {int((s.double() != (ref > U.double()).double()).sum())}')
    hb = torch.randn(N, H, device=dev); b = torch.randn(V, device=dev); sig = torch.rand(H, device=dev) + 0.5
    act, _ = ops.layer_forward(hb, W, b, True, L.ACT_LINEAR, a_div=sig)
    ref = (hb.double() / sig.double()) @ W.double().t() + b.double()
    print(f'[simt fwd transW a_div] rel {rel(act, ref):.2e}')
    # philox: moments
    rng = ops.Rng(seed=3)
    P = torch.full((512, 256), 0.3, device=dev)
    s = ops.sample(P, L.SAMPLE_BERNOULLI, rng=rng)
    g = ops.sample(torch.zeros(512, 256, device=dev), L.SAMPLE_GAUSSIAN, noise_scale=torch.ones(256, device=dev), rng=rng)
    print(f'[philox] bernoulli mean {float(s.mean()):.4f} (0.3)  gaussian mean {float(g.mean()):.4f} std {float(g.std()):.4f}')
    # cd grads general
    for kind in (L.LAYER_RBM, L.LAYER_BGRBM):
        n = ops.cd_param_count(kind, V, H)
        params = torch.randn(n, device=dev) * 0.1
        v0 = (torch.rand(N, V, device=dev) < 0.5).float(); vk = (torch.rand(N, V, device=dev) < 0.5).float()
        g = ops.cd_gradients(kind, v0, vk, params, V, H)
        pd = params.double().clone().requires_grad_(True)
        Wd = pd[:V*H].view(V, H); bd = pd[V*H:V*H+V]; cd_ = pd[V*H+V:V*H+V+H]
        def F(v):
            x = v.double() @ Wd
            if kind == L.LAYER_RBM:
                return -(v.double() @ bd + torch.nn.functional.softplus(x + cd_).sum(1))
            sg = torch.nn.functional.softplus(pd[V*H+V+H:])
            return -(v.double() @ bd + (0.5 * x * x + (cd_ / sg) * x).sum(1))
        loss = F(v0).mean() - F(vk).mean()
        loss.backward()
        print(f'[cd_gradients kind={kind}] rel {rel(g, pd.grad):.2e}')

def check_gp(N, D, Q=2, noise=1.0):
    X = torch.randn(N, Q, device=dev); Y = torch.randn(N, D, device=dev)
    Y = Y - Y.mean(0, keepdim=True)
    isp = lambda v: float(torch.log(torch.expm1(torch.tensor(v, dtype=torch.float64))))
    hyp_opt = torch.tensor([isp(1.0), isp(1.0), isp(noise)], device=dev, dtype=torch.float32)
    plan = ops.GplvmPlan(N, Q, D)
    Lout = torch.zeros(N, N, device=dev)
    t0 = time.time()
    sc, dX, info = plan.evaluate(X, Y, hyp_opt, 7, jitter=1e-6, L_out=Lout)
    torch.cuda.synchronize()
    sc = sc.cpu().double(); 
    Xd = X.double().clone().requires_grad_(True); ho = hyp_opt.double().clone().requires_grad_(True)
    sp = torch.nn.functional.softplus
    al, ga, nz = sp(ho[0]), sp(ho[1]), sp(ho[2]) + 1e-6
    xx = Xd / ga
    d2 = (xx * xx).sum(1)[:, None] + (xx * xx).sum(1)[None, :] - 2 * xx @ xx.t()
    K = al * torch.exp(-0.5 * d2) + nz * torch.eye(N, device=dev, dtype=torch.float64)
    Lr = torch.linalg.cholesky(K)
    S = torch.linalg.solve_triangular(Lr, Y.double(), upper=False)
    logdet = 2 * torch.log(torch.diagonal(Lr)).sum()
    loss = 0.5 * ((S * S).sum() + D * logdet + D * N * torch.log(torch.tensor(2 * torch.pi, dtype=torch.float64)))
    loss.backward()
    print(f'[gplvm N={N} D={D}] info {int(info)} loss rel {abs(float(sc[0]) - float(loss)) / abs(float(loss)):.2e} '
          f'logdet rel {abs(float(sc[1]) - float(logdet)) / abs(float(logdet)):.2e} L rel {rel(torch.tril(Lout), Lr):.2e} '
          f'dX normrel {float((dX.double() - Xd.grad).abs().max() / Xd.grad.abs().max()):.2e} '
          f'dhyp rel {[round(abs(float(sc[3+i]) - float(ho.grad[i])) / (abs(float(ho.grad[i])) + 1e-30), 7) for i in range(3)]}')
    # predict
    xs = torch.randn(37, Q, device=dev)
    hyp = ops.gp_effective_hyp(hyp_opt, 7, 0.0, 1e-6)
    S32 = Y.clone(); ops.trsm_lower(Lout, S32)
    mean, var, pinfo = ops.gp_predict(X, Lout, S32, None, xs, hyp)
    with torch.no_grad():
        xs_ = xs.double() / ga
        d2s = (xx * xx).sum(1)[:, None] + (xs_ * xs_).sum(1)[None, :] - 2 * xx @ xs_.t()
        Kxs = al * torch.exp(-0.5 * d2s)
        V1 = torch.linalg.solve_triangular(Lr, Kxs, upper=False)
        mr = V1.t() @ S; vr = al - (V1 * V1).sum(0)
    print(f'    trsm rel {rel(S32, S.detach()):.2e} pred mean rel {rel(mean, mr):.2e} var rel {rel(var, vr):.2e} info {int(pinfo)}')

def time_gp(N, D, Q=2, iters=3):
    X = torch.randn(N, Q, device=dev); Y = torch.randn(N, D, device=dev)
    hyp_opt = torch.tensor([0.5413, 0.5413, 0.5413], device=dev)
    plan = ops.GplvmPlan(N, Q, D)
    plan.evaluate(X, Y, hyp_opt, 7); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        plan.evaluate(X, Y, hyp_opt, 7)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f'[time gplvm N={N} D={D}] {ms:.2f} ms/eval  {1e3/ms:.1f} evals/s  {(N**3 + 3.0*N*N*D)/ms/1e9:.2f} TFLOP/s')

which = sys.argv[1:] or ['simt', 'gp', 'tc', 'time']
if 'simt' in which:
    check_simt()
if 'gp' in which:
    check_gp(328, 1024); check_gp(200, 50, noise=0.01); check_gp(1000, 64)
if 'tc' in which:
    check_cd_tc(256, 128, 256, 1)
    check_cd_tc(328, 1024, 496, 2)
    check_cd_tc(1024, 512, 768, 3, pcd=False)
if 'time' in which:
    time_gp(328, 1024); time_gp(5000, 784)
    time_cd_tc(2048, 4096, 8192, 1)
    time_cd_tc(16384, 4096, 8192, 10, iters=2)
print('done')
