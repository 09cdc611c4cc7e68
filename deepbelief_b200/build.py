"""Builds libdeepbelief_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m deepbelief_b200.build [--force] [--verbose]

The shared library sits next to this file so it travels with the repository snapshot to the GPU box.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libdeepbelief_b200.so')
STAMP = os.path.join(HERE, 'csrc', '.build_stamp')

SOURCES = ['api.cu', 'host_io.cu', 'layer_simt.cu', 'gemm_tc.cu', 'gemm_tf32.cu', 'gemm_f16x3.cu', 'cd_chain.cu', 'gp.cu', 'gpdbn.cu']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC',
    '--expt-relaxed-constexpr',
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def _source_hash():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ['../../include/deepbelief_b200.h']
    for name in names:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path) and name.endswith(('.cu', '.cuh', '.h')):
            with open(path, 'rb') as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into one shared library. Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
              ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write('--- nvcc %s ---\n%s\n' % (src, out))
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed; see output above')
    link = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart', '-lpthread']
    subprocess.check_call(link)
    with open(STAMP, 'w') as f:
        f.write(_source_hash())
    return LIB_PATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)
