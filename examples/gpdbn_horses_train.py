"""GPDBN on the Weizmann horses (reference examples/gpdbn_horses/train.py): two Concrete-unit RBM layers under a GPRBM
top layer with a 2-D latent space; one Adam step per iteration over every trainable of the stack."""
import _common as cfg
from deepbelief_b200 import compat as tf
from deepbelief.layers.data import Data
from deepbelief.layers.rbm import RBM
from deepbelief.layers.gprbm import GPRBM
from deepbelief.layers import gplvm
from deepbelief.plotting import LatentPointPlotter, LatentSamplePlotter
from deepbelief.util import init_logging

out = cfg.out_dir('gpdbn_horses')
num_iterations = cfg.iterations(20000)
init_logging(out + '/training.log')

lp_plotter = LatentPointPlotter(xlim=(-3, 3), ylim=(-3, 3), delta=0.1, output_dir=out, fig_ext='.png')
ls_plotter = LatentSamplePlotter(image_shape=cfg.IMG_SHAPE, xlim=(-3, 3), ylim=(-3, 3), delta=0.5, num_samples_to_average=100,
                                 output_dir=out)
training_data = Data(cfg.TRAINING_DATA, shuffle_first=False, batch_size=328, log_epochs=False, name='TrainingData')
Y = training_data.next_batch()
session = tf.Session(config=tf.ConfigProto(gpu_options=tf.GPUOptions(allow_growth=True)))

temperature = 0.1
layer1 = RBM(session=session, num_v=cfg.IMG_SHAPE[0] * cfg.IMG_SHAPE[1], num_h=200, W_std=0.01, temperature=temperature,
             name='BBRBM-200')
layer1.init_variables()
layer2 = RBM(session=session, num_v=200, num_h=100, W_std=0.01, temperature=temperature, bottom=layer1, name='BBRBM-200-100')
layer2.init_variables()
kern = gplvm.SEKernel(session=session, alpha=1.0, gamma=1.0, ARD=False, Q=2)
layer3 = GPRBM(num_v=100, num_h=50, W_std=0.01, Q=2, N=328, V=Y, eval_flag=False, kern=kern, noise_variance=0.01,
               fixed_noise_variance=True, temperature=temperature, bottom=layer2, latent_point_plotter=lp_plotter,
               latent_sample_plotter=ls_plotter, session=session, name='GPRBM-100-50')
layer3.build_model()

optimizer = tf.train.AdamOptimizer(learning_rate=0.001)
summary_writer = tf.summary.FileWriter(out + '/summary', session.graph)
layer3.optimize(optimizer=optimizer, num_iterations=num_iterations, fixed_X_num_iterations=min(100, num_iterations // 2),
                eval_interval=max(1, num_iterations // 10), ckpt_interval=num_iterations, ckpt_dir=out,
                summary_writer=summary_writer)
session.close()
print('trained, checkpoints in %s' % out)
