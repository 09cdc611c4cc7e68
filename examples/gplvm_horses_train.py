"""GPLVM with an RBF kernel and a 2-D latent space on the Weizmann horses (reference examples/gplvm_horses/train.py):
log marginal likelihood + analytic gradient per iteration, Adam on the latents and the kernel / noise parameters."""
import _common as cfg
from deepbelief_b200 import compat as tf
from deepbelief.layers.data import Data
from deepbelief.layers import gplvm
from deepbelief.plotting import LatentPointPlotter
from deepbelief.util import init_logging

out = cfg.out_dir('gplvm_horses')
num_iterations = cfg.iterations(5000)
init_logging(out + '/training.log')

training_data = Data(cfg.TRAINING_DATA, shuffle_first=False, batch_size=328, log_epochs=False, name='TrainingData')
Y = training_data.next_batch()
session = tf.Session(config=tf.ConfigProto(gpu_options=tf.GPUOptions(allow_growth=True)))
lp_plotter = LatentPointPlotter(xlim=(-3, 3), ylim=(-3, 3), delta=0.1, output_dir=out, fig_ext='.png')

kern = gplvm.SEKernel(session=session, alpha=1.0, gamma=1.0, ARD=False, Q=2)
layer = gplvm.GPLVM(Y=Y, Q=2, kern=kern, noise_variance=1.0, fixed_noise_variance=False, latent_point_plotter=lp_plotter,
                    session=session, name='GPLVM')
layer.build_model()
optimizer = tf.train.AdamOptimizer(learning_rate=0.01)
layer.optimize(optimizer=optimizer, num_iterations=num_iterations, eval_interval=max(1, num_iterations // 10),
               ckpt_interval=num_iterations, ckpt_dir=out)
print('final loss %.3f' % float(session.run(layer.loss)))
session.close()
