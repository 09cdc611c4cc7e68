"""CPU oracle for the two hot paths of zeis/deepbelief — TEST INFRASTRUCTURE ONLY.

A plain numpy restatement of the reference's algorithms (fp64 = ground truth, fp32 = "what TF would
roughly produce"), with every random draw INJECTED as an array so the CUDA path can be compared on the
same inputs. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl
reference` legs may import this package; the product (`deepbelief_b200`) never does and fails loudly
when its CUDA library is missing.

Pinning: the reference cannot run here (TensorFlow 1.8 / h5py are absent and its own
tests hold no golden vectors or seeds), so the oracle is pinned in two ways:
  1. `tests/golden/make_golden.py` executes the reference's OWN source files from /root/reference
     (rbm.py, bgrbm.py, gbrbm.py, ggrbm.py, gplvm.py, dist.py, preprocessing.py) on top of a small
     numpy stand-in for the `tensorflow` module (`tests/golden/tf_numpy_shim.py`) and commits their
     outputs as `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks the oracle against them;
  2. the fp64 known-answer anchors of SURVEY.md §8(c) on the bundled horses dataset.
TensorFlow's own kernels (Eigen GEMM rounding, Philox stream) are NOT reproduced: parity is defined
with injected randomness and fp32 tolerances.

Modules: rbm (RBM/BGRBM/GBRBM/GGRBM chain, free energy, CD gradient, optimizers), gp (SE kernel,
GPLVM loss / gradient / prediction, PCA), dist, data (Data.next_batch semantics), philox (the
counter-based RNG of the CUDA kernels, for bit-exact sampling checks in Philox mode).
"""
