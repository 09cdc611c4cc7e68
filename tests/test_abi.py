"""The C-ABI shared library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute is called here (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_cuda


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'deepbelief_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(db_[a-z0-9_]+)\s*\(', text)))


def test_library_builds_and_exports_every_declared_symbol():
    from deepbelief_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), 'symbol %s declared in the header but not exported' % name
    # the ctypes binding covers the whole header and nothing else
    assert sorted(_lib.SIGNATURES.keys()) == declared
    assert lib.db_version() == 100


def test_sass_contains_blackwell_native_instructions():
    """tcgen05.mma / TMA / tcgen05.ld must be in the shipped cubin (UTC*MMA, UTMALDG, LDTM)."""
    import subprocess
    from deepbelief_b200 import build
    sass = subprocess.run(['cuobjdump', '-sass', build.build()], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'LDTM'):
        assert mnemonic in sass, mnemonic + ' missing from SASS'
    assert 'HMMA.' not in sass.replace('UTCHMMA', '')   # no legacy mma.sync path


@pytest.mark.skipif(has_cuda(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback_without_a_gpu():
    from deepbelief_b200 import _lib
    out = ctypes.c_void_p()
    assert _lib.lib().db_create(ctypes.byref(out), 0) < 0      # DB_ERR_CUDA: the product path fails loudly
    with pytest.raises(RuntimeError):
        _lib.handle()


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'deepbelief_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f
