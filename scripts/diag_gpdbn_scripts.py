"""Diagnostic (GPU): run the reference's gpdbn_horses train + test_generalisation scripts as the test does and report where
model_data.npy is not finite."""
import os, sys, pickle, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import test_gpu_reference_scripts as T

tmp = tempfile.mkdtemp()
d = T.lay_out(tmp, 'gpdbn_horses')
print(T.run_script(d, 'train.py')[-500:])
out = T.run_script(d, 'test_generalisation.py')
print(out[-1500:])
gen = os.path.join(d, 'output', 'test_plots', 'generalisation')
model = np.load(os.path.join(gen, 'model_data.npy'))
print('non-finite per row:', (~np.isfinite(model)).sum(axis=1))
with open(os.path.join(gen, 'table'), 'rb') as f:
    table = pickle.load(f)
for t in table:
    print(t[0], t[1], t[2], np.isfinite(t[4]).all(), t[5])
log = open(os.path.join(d, 'output', 'training.log')).read()
print(log[-1500:])
