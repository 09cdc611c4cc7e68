"""Throughput of the reference's real (small, latency-bound) configurations — BASELINE configs[0..2] and the
example flags — through the layer API, next to the numpy oracle port on the host cores. Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepbelief_b200 import compat as tf, hdf5
from deepbelief_b200.layers.data import Data
from deepbelief_b200.layers.rbm import RBM
from deepbelief_b200.layers.bgrbm import BGRBM
from deepbelief_b200.layers.gbrbm import GBRBM
from deepbelief_b200.layers.gplvm import GPLVM
from oracle import rbm as orbm, gp as ogp

HORSES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'datasets',
                      'weizmann_horses_training_328x1024_binary.h5')
out = {}

def time_train(layer, data, k, iters, opt, reps=7):
    """Median (and best) seconds per iteration over `reps` train() calls of `iters` iterations each: the loops are 50-250 ms
    long, so a single one also measures the clock ramp after the idle CPU phases in between."""
    layer.train(opt, data, None, num_gibbs_steps=k, pcd=False, max_iterations=100)        # warm-up incl. one graph capture
    times = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        layer.checkpoint_iter = 1
        layer.train(opt, data, None, num_gibbs_steps=k, pcd=False, max_iterations=iters)
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) / iters)
    times.sort()
    return times[len(times) // 2], times[0]

def cpu_cd(kind, V, H, k, v0, iters=5):
    rs = np.random.RandomState(0)
    m = orbm.RBMOracle((rs.randn(V, H) * 0.01).astype(np.float32), np.zeros(V, np.float32), np.zeros(H, np.float32), kind=kind,
                       dtype=np.float32)
    opt = orbm.Adam(1e-3); N = v0.shape[0]
    t0 = time.perf_counter()
    for _ in range(iters):
        gauss = kind in ('bgrbm', 'ggrbm')
        dh0 = (rs.standard_normal((N, H)) if gauss else rs.random_sample((N, H))).astype(np.float32)
        chain = [(rs.random_sample((N, V)).astype(np.float32),
                  (rs.standard_normal((N, H)) if gauss else rs.random_sample((N, H))).astype(np.float32)) for _ in range(k)]
        o = orbm.cd_k(m, v0, k, dh0, chain)
        th = opt.step(m.pack(), o['grad'].astype(np.float32))
        m.W = th[:V * H].reshape(V, H); m.b = th[V * H:V * H + V].reshape(V, 1); m.c = th[V * H + V:V * H + V + H].reshape(1, H)
    return (time.perf_counter() - t0) / iters

horses = hdf5.File(HORSES)['datapoints']
N = horses.shape[0]
cases = [('config1: RBM 1024x500 CD-1 Adam, horses 328 (full batch)', RBM, 'rbm', 500, 1, horses),
         ('dbn_horses layer 1: RBM 1024x200 CD-10 Adam', RBM, 'rbm', 200, 10, horses),
         ('config2: BGRBM 1024x500 CD-1 Adam, binary horses', BGRBM, 'bgrbm', 500, 1, horses),
         ('config2: GBRBM 1024x500 CD-1 Adam, standardized horses', GBRBM, 'gbrbm', 500, 1,
          ((horses - horses.mean(0)) / np.where(horses.std(0) > 1e-8, horses.std(0), 1.0)).astype(np.float32))]
for name, cls, kind, H, k, data_np in cases:
    np.random.seed(0)
    lay = cls(1024, H, W_std=0.01, name='small', seed=1)
    data = Data.from_arrays(data_np, batch_size=N, log_epochs=False)
    sec, best = time_train(lay, data, k, 500, tf.train.AdamOptimizer(1e-3))
    cpu = cpu_cd(kind, 1024, H, k, data_np)
    flop = (2 * k + 3) * 2.0 * N * 1024 * H
    out[name] = {'gpu_ms_per_iter': sec * 1e3, 'gpu_ms_per_iter_best': best * 1e3, 'gpu_sample_steps_per_s': N * k / sec, 'gpu_algorithmic_tflops': flop / sec / 1e12,
                 'cpu_ms_per_iter': cpu * 1e3, 'cpu_sample_steps_per_s': N * k / cpu, 'path': 'tcgen05' if lay._tc_eligible(N, k) else 'simt'}

# config 3: GPLVM on horses, Q=2, alpha=gamma=1, noise 1 trainable
g = GPLVM(horses, 2, noise_variance=1.0)
g.build_model()
for _ in range(5): g.loss_and_grad()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): g.loss_and_grad()
torch.cuda.synchronize(); ev = (time.perf_counter() - t0) / 200
g.optimize(tf.train.AdamOptimizer(1e-3), num_iterations=5)
torch.cuda.synchronize(); t0 = time.perf_counter()
g.optimize(tf.train.AdamOptimizer(1e-3), num_iterations=200)
torch.cuda.synchronize(); it = (time.perf_counter() - t0) / 200
o = ogp.GPLVMOracle(g.Y_np, g.X0, dtype=np.float32)
t0 = time.perf_counter()
for _ in range(5):
    o.loss(); o.grad()
cpu = (time.perf_counter() - t0) / 5
out['config3: GPLVM horses N=328 D=1024 Q=2'] = {'gpu_ms_per_eval': ev * 1e3, 'gpu_evals_per_s': 1 / ev, 'gpu_optimize_iters_per_s': 1 / it,
                                                 'cpu_ms_per_eval': cpu * 1e3, 'cpu_evals_per_s': 1 / cpu}
out['host_cores'] = os.cpu_count()
print(json.dumps(out, indent=1))
