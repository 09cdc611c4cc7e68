"""Where does the tcgen05 CD chain start to beat the general SIMT path (graph replay)? RBM.train, Adam, non-PCD."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from deepbelief_b200 import compat as tf
from deepbelief_b200.layers.data import Data
from deepbelief_b200.layers.rbm import RBM

rs = np.random.RandomState(0)
for (N, V, H) in ((328, 1024, 200), (328, 1024, 496), (656, 1024, 496), (1024, 1024, 512), (2048, 1024, 1024), (2048, 2048, 2048)):
    X = (rs.rand(N, V) < 0.4).astype(np.float32)
    for k in (1, 10):
        row = []
        for name, min_work in (('tc', 0), ('simt', 1 << 62)):
            RBM.TC_MIN_WORK = RBM.TC_MIN_WORK_LONG_CHAIN = min_work
            lay = RBM(V, H, name='p', seed=1)
            opt = tf.train.AdamOptimizer(learning_rate=1e-3)
            data = Data.from_arrays(X, batch_size=N, log_epochs=False)
            lay.train(opt, data, None, num_gibbs_steps=k, pcd=False, max_iterations=20)
            torch.cuda.synchronize()
            iters = 300 if N * V * H < (1 << 32) else 30
            lay.checkpoint_iter = 1
            t = time.perf_counter()
            lay.train(opt, data, None, num_gibbs_steps=k, pcd=False, max_iterations=iters)
            torch.cuda.synchronize()
            row.append('%s %.3f ms' % (name, (time.perf_counter() - t) / iters * 1e3))
        print('N=%d V=%d H=%d k=%d work=%.2e: %s' % (N, V, H, k, N * V * H, '  '.join(row)), flush=True)
