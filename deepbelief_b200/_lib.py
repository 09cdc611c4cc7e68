"""ctypes binding of include/deepbelief_b200.h (the C ABI of the CUDA library).

PyTorch is used only for tensor ownership: every call below receives raw device pointers of
contiguous fp32 CUDA tensors plus the current CUDA stream. There is no CPU fallback: if the shared
library is missing or no CUDA device is present, `lib()` / `handle()` raise.
"""
import ctypes
import os
import threading

import torch

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

c_float_p = ctypes.c_void_p
c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t


class db_rng_t(ctypes.Structure):
    _fields_ = [('seed', ctypes.c_uint64), ('draw', ctypes.c_uint64),
                ('row_offset', ctypes.c_uint32), ('reserved', ctypes.c_uint32), ('clock', ctypes.c_void_p)]


class db_opt_t(ctypes.Structure):
    _fields_ = [('kind', c_int), ('lr', c_float), ('momentum', c_float), ('beta1', c_float),
                ('beta2', c_float), ('eps', c_float), ('weight_decay', c_float), ('step', c_int), ('clock', ctypes.c_void_p)]


class db_cd_tc_desc_t(ctypes.Structure):
    _fields_ = [('N', c_int), ('V', c_int), ('H', c_int), ('k', c_int), ('pcd', c_int),
                ('inv_global_n', c_float), ('v0_u8', c_int)]


# enums (mirror the header)
ACT_LINEAR, ACT_SIGMOID = 0, 1
SAMPLE_NONE, SAMPLE_BERNOULLI, SAMPLE_CONCRETE, SAMPLE_GAUSSIAN = 0, 1, 2, 3
BCAST_NONE, BCAST_ROW, BCAST_FULL = 0, 1, 2
LAYER_RBM, LAYER_BGRBM, LAYER_GBRBM, LAYER_GGRBM = 0, 1, 2, 3
OPT_SGD, OPT_MOMENTUM, OPT_ADAM = 0, 1, 2
ERR_BAD_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = -1, -2, -3, -4

# name -> (restype, argtypes); the test-suite checks that every symbol of the header is exported
SIGNATURES = {
    'db_create': (c_int, [ctypes.POINTER(c_void_p), c_int]),
    'db_destroy': (c_int, [c_void_p]),
    'db_last_error': (ctypes.c_char_p, [c_void_p]),
    'db_version': (c_int, []),
    'db_num_sms': (c_int, [c_void_p]),
    'db_layer_forward': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_float_p, c_int, c_float_p, c_int, c_int,
                                 c_float_p, c_int, c_float_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                 c_float_p, c_int, c_float_p, ctypes.POINTER(db_rng_t), c_float_p, c_float_p]),
    'db_sample': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_int, c_int, c_float, c_float_p, c_int,
                          c_float_p, ctypes.POINTER(db_rng_t), c_float_p]),
    'db_free_energy': (c_int, [c_void_p, c_void_p, c_int, c_float_p, c_int, c_int, c_int, c_float_p, c_float_p,
                               c_float_p, c_float_p, c_float_p, c_float_p]),
    'db_cd_param_count': (c_size_t, [c_int, c_int, c_int]),
    'db_cd_gradients': (c_int, [c_void_p, c_void_p, c_int, c_float_p, c_float_p, c_int, c_int, c_int, c_float_p,
                                c_float, c_float_p, c_float_p]),
    'db_cd_gradients_hidden': (c_int, [c_void_p, c_void_p, c_int, c_float_p, c_float_p, c_float_p, c_float_p, c_int, c_int,
                                       c_int, c_float_p, c_float, c_float_p]),
    'db_clock_advance': (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_uint64, ctypes.c_uint64]),
    'db_optimizer_step': (c_int, [c_void_p, c_void_p, ctypes.POINTER(db_opt_t), c_float_p, c_float_p, c_float_p,
                                  c_size_t]),
    'db_cd_tc_workspace_bytes': (c_size_t, [ctypes.POINTER(db_cd_tc_desc_t)]),
    'db_cd_tc_weight_cache_bytes': (c_size_t, [c_int, c_int]),
    'db_cd_tc_refresh_weights': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_int, c_void_p]),
    'db_cd_tc_step': (c_int, [c_void_p, c_void_p, ctypes.POINTER(db_cd_tc_desc_t), c_float_p, c_float_p, c_void_p,
                              c_void_p, c_float_p, ctypes.POINTER(db_rng_t), c_void_p, c_float_p, c_float_p,
                              c_float_p, c_float_p, c_float_p]),
    'db_cd_tc_step_update': (c_int, [c_void_p, c_void_p, ctypes.POINTER(db_cd_tc_desc_t), c_float_p, c_float_p, c_void_p,
                                     c_void_p, c_float_p, ctypes.POINTER(db_rng_t), c_void_p, c_float_p,
                                     ctypes.POINTER(db_opt_t), c_float_p]),
    'db_cd_tc_apply_update': (c_int, [c_void_p, c_void_p, c_int, c_int, c_float_p, c_float_p, ctypes.POINTER(db_opt_t), c_float_p,
                                      c_void_p]),
    'db_gemm_f16x3_scratch_bytes': (c_size_t, [c_int, c_int, c_int]),
    'db_gemm_f16x3_nt': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float_p, c_int, c_float_p, c_int, c_float,
                                 c_float_p, c_int, c_int, c_void_p]),
    'db_enable_peer_access': (c_int, [c_void_p, c_int]),
    'db_ipc_export': (c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(ctypes.c_uint64)]),
    'db_ipc_import': (c_int, [c_void_p, c_void_p, ctypes.c_uint64, ctypes.POINTER(c_void_p)]),
    'db_cd_tc_push_ok': (c_int, [ctypes.POINTER(db_cd_tc_desc_t), c_int]),
    'db_cd_tc_step_push': (c_int, [c_void_p, c_void_p, ctypes.POINTER(db_cd_tc_desc_t), c_void_p, c_float_p, c_void_p,
                                   c_void_p, c_float_p, ctypes.POINTER(db_rng_t), c_void_p, c_float_p,
                                   ctypes.POINTER(c_void_p), c_int, c_int]),
    'db_cd_tc_apply_update_sharded': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, ctypes.POINTER(c_void_p), c_float_p,
                                              c_float_p, ctypes.POINTER(db_opt_t), c_float_p]),
    'db_cd_tc_update_fusable': (c_int, [ctypes.POINTER(db_cd_tc_desc_t)]),
    'db_gram_rbf': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_int, c_int, c_int, c_float_p, c_int,
                            c_float_p, c_int]),
    'db_potrf_lower': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_int, c_void_p]),
    'db_trsm_lower': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_int, c_float_p, c_int, c_int, c_int]),
    'db_gplvm_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'db_gplvm_lml_grad': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_int, c_int, c_int, c_float_p, c_int,
                                  c_float, c_float, c_float, c_int, c_int, c_void_p, c_float_p, c_float_p,
                                  c_float_p, c_void_p]),
    'db_gp_predict': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_int,
                              c_int, c_int, c_int, c_float_p, c_float_p, c_float_p, c_float_p, c_void_p]),
    'db_gp_effective_hyp': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_float, c_float, c_float_p]),
    'db_softplus': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_float_p]),
    'db_mse': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_int, c_int, c_float_p]),
    'db_cross_entropy': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_int, c_int, c_int, c_float_p]),
    'db_tf32_split_lo': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_int, c_int, c_float_p]),
    'db_gemm_tf32x3_nt': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float_p, c_float_p, c_int, c_float_p,
                                  c_float_p, c_int, c_float, c_float_p, c_int, c_int]),
    'db_gpdbn_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'db_gpdbn_factor': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_int, c_int, c_float_p, c_int, c_float, c_float,
                                c_void_p, c_float_p, c_float_p, c_void_p]),
    'db_gpdbn_apply': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_int, c_int, c_void_p, c_float_p, c_float_p]),
    'db_gpdbn_backward_h': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_int, c_int, c_void_p, c_float_p]),
    'db_gpdbn_backward_x': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_int, c_int, c_int, c_float_p, c_int,
                                    c_float, c_void_p, c_float_p, c_float_p]),
    'db_gpdbn_top_forward_up': (c_int, [c_void_p, c_void_p] + [c_float_p] * 5 + [c_int, c_int] + [c_float_p] * 3),
    'db_gpdbn_top_forward_mid': (c_int, [c_void_p, c_void_p] + [c_float_p] * 5 + [c_int, c_int] + [c_float_p] * 2),
    'db_gpdbn_top_backward1': (c_int, [c_void_p, c_void_p] + [c_float_p] * 6 + [c_int, c_int] + [c_float_p] * 4),
    'db_gpdbn_top_backward2': (c_int, [c_void_p, c_void_p] + [c_float_p] * 9 + [c_int, c_int] + [c_float_p] * 6),
    'db_testgen_rowloss': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_float_p, c_int, c_int, c_float] + [c_float_p] * 5),
    'db_testgen_pred_var': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_int, c_int, c_float_p, c_float_p, c_void_p]),
    'db_testgen_grad_x': (c_int, [c_void_p, c_void_p] + [c_float_p] * 6 + [c_int, c_int, c_int, c_float_p, c_float_p]),
    'db_testgen_meanloss': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_float_p, c_int, c_int, c_int, c_float,
                                    c_float_p, c_float_p, c_float_p]),
    'db_testgen_reduce_samples': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_float_p, c_float_p, c_int, c_int,
                                          c_int, c_float, c_float_p, c_float_p]),
    'db_heldout_mse': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_int, c_int, c_int, c_float_p, c_float_p]),
    'db_interp_objective': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_float_p, c_float_p, c_int, c_int, c_float,
                                    c_float_p, c_float_p, c_float_p]),
    'db_gemm': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float_p, c_int, c_float_p,
                        c_int, c_float, c_float_p, c_int]),
    'db_concrete_backward': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_float_p, c_int, c_int, c_float,
                                     c_float_p]),
    'db_sigmoid_backward': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_int, c_int, c_float_p]),
    'db_cross_entropy_grad': (c_int, [c_void_p, c_void_p, c_float_p, c_float_p, c_int, c_int, c_float, c_float_p]),
    'db_sbm_expand': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_int, c_int, c_float_p]),
    'db_sbm_gather': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_int, c_int, c_float_p]),
    'db_colsum': (c_int, [c_void_p, c_void_p, c_float_p, c_int, c_int, c_float, c_float, c_float_p]),
    'db_host_gather_rows': (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_void_p, c_int]),
}

_lock = threading.Lock()
_lib = None
_handles = {}


def lib_path():
    return _build.LIB_PATH


def lib():
    """Load (building first if the in-tree library is missing or stale) and return the ctypes library."""
    global _lib
    with _lock:
        if _lib is None:
            path = os.environ.get('DEEPBELIEF_B200_LIB')                        # override: instrumented debug builds
            if not path:
                # csrc/.build_stamp (untracked, written next to the .so) holds the hash of the sources the library was built
                # from: a library older than the sources would be called with newer struct layouts, so it is rebuilt
                path = _build.build() if _build.needs_build() else _build.LIB_PATH
            try:
                loaded = ctypes.CDLL(path)
            except OSError as e:   # no silent fallback
                raise RuntimeError('deepbelief_b200: cannot load the CUDA library %s: %s' % (path, e))
            for name, (restype, argtypes) in SIGNATURES.items():
                fn = getattr(loaded, name)
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = loaded
    return _lib


class DeepbeliefError(RuntimeError):
    pass


def handle(device=None):
    """Per-device library handle; raises if no CUDA device is available (there is no CPU path)."""
    if not torch.cuda.is_available():
        raise RuntimeError('deepbelief_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists')
    dev = torch.cuda.current_device() if device is None else torch.device(device).index or 0
    h = _handles.get(dev)
    if h is None:
        out = c_void_p()
        rc = lib().db_create(ctypes.byref(out), dev)
        if rc != 0:
            raise RuntimeError('db_create(device=%d) failed with %d' % (dev, rc))
        h = _handles[dev] = out
    return h


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), 'expected a contiguous CUDA tensor'
    return c_void_p(t.data_ptr())


def fptr(t):
    if t is None:
        return None
    assert t.dtype == torch.float32, 'expected float32 (config.float_type), got %s' % t.dtype
    return ptr(t)


def check(rc, what):
    """0 -> ok; <0 -> raise with the library's message; >0 -> returned to the caller (LAPACK-style info)."""
    if rc < 0:
        msg = lib().db_last_error(handle()).decode('utf-8', 'replace')
        if rc == ERR_BAD_ARG:
            raise AssertionError('%s: %s' % (what, msg))
        if rc == ERR_UNSUPPORTED:
            raise ValueError('%s: %s' % (what, msg))
        raise DeepbeliefError('%s failed (%d): %s' % (what, rc, msg))
    return rc


def make_rng(seed, draw, row_offset=0, clock=None):
    """clock: device address of a 2 x uint64 clock (db_clock_advance) or None."""
    return db_rng_t(int(seed) & 0xFFFFFFFFFFFFFFFF, int(draw) & 0xFFFFFFFFFFFFFFFF, int(row_offset), 0, clock)
