"""Times RBM.train on the reference-sized configurations with and without the captured training iteration."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from deepbelief_b200 import compat as tf, hdf5
from deepbelief_b200.layers.data import Data
from deepbelief_b200.layers.rbm import RBM
from deepbelief_b200.layers.bgrbm import BGRBM

X = hdf5.File('tests/golden/datasets/weizmann_horses_training_328x1024_binary.h5')['datapoints'].astype(np.float32)
for name, cls, H, k, pcd in (('rbm 1024x500 cd1', RBM, 500, 1, False), ('rbm 1024x200 cd10 pcd', RBM, 200, 10, True),
                             ('bgrbm 1024x500 cd1', BGRBM, 500, 1, False)):
    for graph in ('0', '1'):
        os.environ['DEEPBELIEF_B200_GRAPH'] = graph
        lay = cls(1024, H, name='p', seed=1)
        opt = tf.train.AdamOptimizer(learning_rate=1e-3)
        data = Data.from_arrays(X, batch_size=328, log_epochs=False)
        lay.train(opt, data, None, num_gibbs_steps=k, pcd=pcd, max_iterations=50)
        torch.cuda.synchronize()
        t = time.perf_counter()
        iters = 500
        lay.checkpoint_iter = 1
        lay.train(opt, data, None, num_gibbs_steps=k, pcd=pcd, max_iterations=iters)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        print('%-24s graph=%s  %.3f ms/iter' % (name, graph, dt / iters * 1e3), flush=True)
