"""Greedy layer-wise training of a 1024-200-100-50 deep belief network on the Weizmann horses with CD-10 and Adam
(the workflow of the reference's examples/dbn_horses/train.py with its flags.py values)."""
import _common as cfg
from deepbelief_b200 import compat as tf                      # reference scripts: `import tensorflow as tf`
from deepbelief.layers.data import Data
from deepbelief.layers.rbm import RBM
from deepbelief.plotting import GraphPlotter, ImageRowPlotter
from deepbelief.util import init_logging

out = cfg.out_dir('dbn_horses')
num_iterations = cfg.iterations(20000)
init_logging(out + '/training.log')

loss_plotter = GraphPlotter(xlim=(0, num_iterations), ylim=(0, 600), xlabel='Iteration', ylabel='Loss', output_dir=out)
distr_plotter = ImageRowPlotter(num_images=2, image_shape=cfg.IMG_SHAPE, output_dir=out)
training_data = Data(cfg.TRAINING_DATA, shuffle_first=False, batch_size=328, log_epochs=False, name='TrainingData')
test_data = Data(cfg.TEST_DATA, shuffle_first=False, batch_size=8, log_epochs=False, name='TestData')

optimizer = tf.train.AdamOptimizer(learning_rate=0.001)
session = tf.Session(config=tf.ConfigProto(gpu_options=tf.GPUOptions(allow_growth=True)))

sizes = [cfg.IMG_SHAPE[0] * cfg.IMG_SHAPE[1], 200, 100, 50]
bottom = None
for i in range(3):
    layer = RBM(session=session, num_v=sizes[i], num_h=sizes[i + 1], lr=0.001, W_std=0.01, bottom=bottom,
                loss_plotter=loss_plotter, distr_plotter=distr_plotter, name='BBRBM-' + '-'.join(str(s) for s in sizes[1:i + 2]))
    layer.train(optimizer=optimizer, training_data=training_data, test_data=test_data, num_gibbs_steps=10, pcd=False,
                max_iterations=num_iterations, eval_interval=max(1, num_iterations // 20), ckpt_interval=num_iterations,
                ckpt_dir=out)
    bottom = layer
session.close()
print('trained %d layers, checkpoints in %s' % (3, out))
