import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
os.environ.setdefault('DEEPBELIEF_B200_GRAPH', '1')
from deepbelief_b200 import compat as tf, hdf5
from deepbelief_b200.layers.data import Data
from deepbelief_b200.layers.rbm import RBM
X = hdf5.File('tests/golden/datasets/weizmann_horses_training_328x1024_binary.h5')['datapoints'].astype(np.float32)
lay = RBM(1024, 500, name='p', seed=1)
opt = tf.train.AdamOptimizer(learning_rate=1e-3)
data = Data.from_arrays(X, batch_size=328, log_epochs=False)
lay.train(opt, data, None, num_gibbs_steps=1, pcd=False, max_iterations=14)
torch.cuda.synchronize()
