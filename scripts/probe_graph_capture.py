"""How long does CdGraphStep construction (eager iteration + capture + instantiate) take, and how stable is replay?"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from deepbelief_b200 import compat as tf, hdf5
from deepbelief_b200.layers.data import Data
from deepbelief_b200.layers import rbm as R

X = hdf5.File('tests/golden/datasets/weizmann_horses_training_328x1024_binary.h5')['datapoints'].astype(np.float32)
orig = R.CdGraphStep.__init__
def timed(self, *a, **k):
    torch.cuda.synchronize(); t = time.perf_counter()
    orig(self, *a, **k)
    torch.cuda.synchronize(); print('  capture %.1f ms' % ((time.perf_counter() - t) * 1e3), flush=True)
R.CdGraphStep.__init__ = timed
for rep in range(6):
    lay = R.RBM(1024, 500, name='p', seed=1)
    opt = tf.train.AdamOptimizer(learning_rate=1e-3)
    data = Data.from_arrays(X, batch_size=328, log_epochs=False)
    for iters in (50, 500):
        lay.checkpoint_iter = 1
        torch.cuda.synchronize(); t = time.perf_counter()
        lay.train(opt, data, None, num_gibbs_steps=1, pcd=False, max_iterations=iters)
        torch.cuda.synchronize()
        print('rep %d: %d iterations %.3f ms/iter (whole call %.1f ms)' % (rep, iters, (time.perf_counter() - t) / iters * 1e3, (time.perf_counter() - t) * 1e3), flush=True)
